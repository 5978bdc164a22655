"""Host-side mirror of LAMPrey's "chain" step, the immediate consumer of the seed hits
(SURVEY.md section 8f, next row 1):

* ``removeOverlappingPrimerAlignments``   /root/reference/lamprey.py:859-899
* ``extractAmpliconAroundTarget``         /root/reference/lamprey.py:903-998
* ``target_subreads``                     the stage-1 loop of processLamplicon, lamprey.py:1566-1589

Names, arguments and return values follow the reference.  These are O(hits) list walks and stay on the
host, as the north star prescribes; the overlap pruning additionally exists as a device pass
(``lsw_set_overlap``: every hit's ``reserved`` field is 1 if it survives), which ``SeedBatch.kept`` exposes
for whole-FASTQ batches.
"""


def removeOverlappingPrimerAlignments(debug_print, alignments, allowed_overlap):
    """Start-sorted hits in, pruned list out: a hit that starts more than `allowed_overlap` bases before the
    end of the current survivor competes with it; the strictly higher identity wins."""
    new_alignments = []
    last_alignment = alignments[0]
    for idx, alignment in enumerate(alignments):
        if idx == 0:
            continue
        if alignment.start < last_alignment.end - allowed_overlap:
            if debug_print:
                print("  - Found overlapping primer alignment: {}".format(
                    str(alignment.start) + " : " + str(alignment.end)), flush=True)
            if alignment.identity > last_alignment.identity:
                last_alignment = alignment
        else:
            new_alignments.append(last_alignment)
            last_alignment = alignment
    new_alignments.append(last_alignment)
    return new_alignments


def extractAmpliconAroundTarget(primer_set, alignments, target):
    """(amplicon_start, amplicon_end) around a target hit, following the expected primer order to the right
    and to the left; amplicon_end == -1 means "to the end of the read"."""
    amplicon_start = 0
    amplicon_end = 0
    fwd_strand = target.primer_name == "T"
    order = primer_set.fwd_primer_order if fwd_strand else primer_set.rev_primer_order
    target_index = [i for i, a in enumerate(alignments) if a is target][0]

    allowed_mismatches = 0
    primer_counter = order.index('T' if fwd_strand else 'Tc') + 1
    all_matched = True
    for i in range(target_index + 1, len(alignments)):
        if alignments[i].primer_name != order[primer_counter]:
            if allowed_mismatches > 0:
                allowed_mismatches -= 1
            else:
                amplicon_end = alignments[i - 1].end
                all_matched = False
                break
        else:
            primer_counter += 1
            if primer_counter == len(primer_set.fwd_primer_order):
                break
    if all_matched:
        amplicon_end = -1

    allowed_mismatches = 0
    primer_counter = order.index('T' if fwd_strand else 'Tc') - 1
    all_matched = True
    for i in range(target_index - 1, 0, -1):          # the reference never looks at index 0
        if alignments[i].primer_name != order[primer_counter]:
            if allowed_mismatches > 0:
                allowed_mismatches -= 1
            else:
                amplicon_start = alignments[i + 1].start
                all_matched = False
                break
        else:
            primer_counter -= 1
            if primer_counter < 0:
                break
    if all_matched:
        amplicon_start = 0
    return amplicon_start, amplicon_end


def target_subreads(primer_set, alignments, read_length, lamp_idx=0):
    """Stage 1 of processLamplicon (lamprey.py:1566-1589): one (start, end, id, direction) per target hit."""
    out = []
    target_idx = 0
    for alignment in alignments:
        if alignment.primer_name in ('T', 'Tc'):
            seq_start, seq_end = extractAmpliconAroundTarget(primer_set, alignments, alignment)
            if seq_end == -1:
                seq_end = read_length - 1
            out.append((seq_start, seq_end, "{}_{}".format(lamp_idx, target_idx),
                        'fwd' if alignment.primer_name == 'T' else 'rev'))
            target_idx += 1
    return out
