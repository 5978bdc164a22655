# cython: language_level=3
"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement of the two "chain" functions that consume the seed hits:

* ``remove_overlapping_primer_alignments``   /root/reference/lamprey.py:859-899
* ``extract_amplicon_around_target``         /root/reference/lamprey.py:903-998

Both are order-sensitive host code over the per-read hit list; restated here so the mirror in
lamprey_b200/chain.py can be checked.  PARITY UNPINNED (the reference ships no tests).
"""


def remove_overlapping_primer_alignments(alignments, allowed_overlap):
    """Walk the start-sorted hits; when a hit starts more than `allowed_overlap` bases before the end of the
    current survivor, keep whichever has the strictly higher identity (the earlier one on ties)."""
    kept = []
    current = alignments[0]
    for idx, aln in enumerate(alignments):
        if idx == 0:
            continue
        if aln.start < current.end - allowed_overlap:
            if aln.identity > current.identity:
                current = aln
        else:
            kept.append(current)
            current = aln
    kept.append(current)
    return kept


def extract_amplicon_around_target(primer_set, alignments, target):
    """Extend from the target hit to the right and to the left while the hits follow the expected primer
    order; returns (amplicon_start, amplicon_end) with -1 meaning "to the end of the read"."""
    amplicon_start = 0
    amplicon_end = 0
    fwd = target.primer_name == "T"
    order = primer_set.fwd_primer_order if fwd else primer_set.rev_primer_order
    t_index = alignments.index(target)

    # to the right
    allowed_mismatches = 0
    counter = order.index('T' if fwd else 'Tc') + 1
    all_matched = True
    for i in range(t_index + 1, len(alignments)):
        expected = order[counter]
        if alignments[i].primer_name != expected:
            if allowed_mismatches > 0:
                allowed_mismatches -= 1
            else:
                amplicon_end = alignments[i - 1].end
                all_matched = False
                break
        else:
            counter += 1
            if counter == len(primer_set.fwd_primer_order):
                break
    if all_matched:
        amplicon_end = -1

    # to the left (the reference's loop stops before index 0)
    allowed_mismatches = 0
    counter = order.index('T' if fwd else 'Tc') - 1
    all_matched = True
    for i in range(t_index - 1, 0, -1):
        expected = order[counter]
        if alignments[i].primer_name != expected:
            if allowed_mismatches > 0:
                allowed_mismatches -= 1
            else:
                amplicon_start = alignments[i + 1].start
                all_matched = False
                break
        else:
            counter -= 1
            if counter < 0:
                break
    if all_matched:
        amplicon_start = 0
    return amplicon_start, amplicon_end
