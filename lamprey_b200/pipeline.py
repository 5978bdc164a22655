"""``processLamplicon`` up to the end of stage 1, fed from a batched seeding result
(a lamprey.py-compatible shim that consumes batched hits).

The reference's ``processLamplicon`` (/root/reference/lamprey.py:1521-1605) starts every read with
``findAllPrimerAlignments`` (1533), ``removeOverlappingPrimerAlignments(..., 4)`` (1536), the target / ont counts
(1547-1553) and, per target hit, ``extractAmpliconAroundTarget`` + the sub-read record (1566-1601).  ``SeedStage`` is that
prefix with the alignments coming from ``seeding.seed_reads(batch, ..., allowed_overlap=4, subreads=True)``: one device
pass per FASTQ instead of ~65 align() calls per read.  Everything after it in the reference (minimap2, pysam, pileups,
voting) is host / external-tool code and out of scope."""
from . import seeding, skbio_ssw


class SeedStage(object):
    """What processLamplicon holds after stage 1 for one read."""

    def __init__(self, alignments, alignment_coverage, target_depth, num_ont_seqs, alignment_intervals,
                 alignment_interval_dict, subread_slices):
        self.alignments = alignments                          # result.alignments (after overlap pruning), lamprey.py:1539
        self.alignment_coverage = alignment_coverage
        self.target_depth = target_depth                      # lamprey.py:1547-1549
        self.num_ont_seqs = num_ont_seqs                      # lamprey.py:1551-1553
        self.alignment_intervals = alignment_intervals        # [(seq_start, seq_end, subread_id)], lamprey.py:1586
        self.alignment_interval_dict = alignment_interval_dict  # subread_id -> (seq_start, seq_end, direction), lamprey.py:1587
        self.subread_slices = subread_slices                  # [(subread_id, direction, lamplicon.seq[seq_start:seq_end])], 1590-1601


def seed_batch(reads, primers, threshold, aligner=None, device=0, args=None):
    """One device pass for a whole batch of reads: seeds, overlap pruning (4 bp, lamprey.py:1536) and stage-1 sub-reads."""
    return seeding.seed_reads(reads, primers, threshold, aligner=aligner, device=device, allowed_overlap=4, subreads=True, args=args)


def processLamplicon_seed_stage(seeds, k, lamp_idx, lamplicon_seq, primers, threshold):
    """lamprey.py:1533-1601 for read k of a batch seeded by seed_batch(): -> SeedStage.  lamplicon_seq: the read (str or Seq)."""
    alignment_coverage = float(seeds.coverage[k]) if seeds.read_lengths[k] else 0.0
    alignments = seeds.alignments(k, pruned=True)
    target_depth = 0
    num_ont_seqs = 0
    for alignment in alignments:
        if alignment.primer_name in ['T', 'Tc']:
            target_depth = target_depth + 1
        if alignment.primer_name.startswith('ont') and alignment.identity >= threshold:
            num_ont_seqs = num_ont_seqs + 1
    intervals, interval_dict, slices = [], {}, []
    if target_depth > 0:
        for seq_start, seq_end, subread_id, direction in seeds.target_subreads(k, lamp_idx):
            intervals.append((seq_start, seq_end, subread_id))
            interval_dict[subread_id] = (seq_start, seq_end, direction)
            slices.append((subread_id, direction, lamplicon_seq[seq_start:seq_end]))
    return SeedStage(alignments, alignment_coverage, target_depth, num_ont_seqs, intervals, interval_dict, slices)


cigarOpMap = {'M': 0, 'I': 1, 'D': 2}          # the pysam codes LAMPrey's cigarStringToTuples produces (lamprey.py:527-545)


def stage2_align_batch(reads, seeds, target_ref_str, device=0):
    """Stage 2 of processLamplicon for a whole seeded batch (lamprey.py:1637-1669): every candidate sub-read of every read --
    seeds.subreads, from seed_batch() -- against the target reference, in one device pass.  -> one dict per sub-read, in
    seeds.subreads order, with what the reference puts into its pysam.AlignedSegment: query_name parts, reference_start
    (target_begin), cigar string and pysam tuples, and query_sequence (aligned_query_sequence without '-')."""
    sub = seeds.subreads
    pairs = [(int(x["read"]), int(x["start"]), int(x["end"]), int(x["direction"])) for x in sub]
    als = skbio_ssw.align_subreads(reads, pairs, target_ref_str, device=device)
    out = []
    for x, a in zip(sub, als):
        out.append({"read": int(x["read"]), "target_idx": int(x["target_idx"]), "direction": 'fwd' if x["direction"] == 0 else 'rev',
                    "reference_start": a.target_begin, "cigar": a.cigar,
                    "cigartuples": [(cigarOpMap[op], n) for n, op in a._runs],
                    "query_sequence": a.aligned_query_sequence.replace('-', '') if a.target_begin >= 0 else "",
                    "score": a.optimal_alignment_score})
    return out
