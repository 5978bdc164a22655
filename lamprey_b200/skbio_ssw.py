"""Drop-in for ``from skbio.alignment import StripedSmithWaterman`` as LAMPrey uses it
(/root/reference/lamprey.py:17, 577-578, 1643-1647), backed by the batched GPU engine.

``StripedSmithWaterman(query)(target)`` returns an ``AlignmentStructure`` with the fields LAMPrey reads:
``optimal_alignment_score, query_begin, query_end, target_begin, target_end_optimal, cigar, aligned_query_sequence,
aligned_target_sequence`` (zero-based, inclusive ends: skbio's defaults).  Only skbio's default parameterisation is offered
(gap_open 5, gap_extend 2, match 2, mismatch -3, nucleotides); other keyword values raise NotImplementedError rather than
silently differ.  The arithmetic runs in liblsw.so (lsw_set_ssw + lsw_align_tasks); nothing here aligns on the CPU.

Two device paths sit behind the one class: a query of at most 63 bases (the seeding call: query = primer) runs on the
register-row kernel; a longer query (LAMPrey's stage-2 call: query = a whole sub-read against the target reference,
lamprey.py:1643-1647) runs on the stage-2 kernel (``lsw_ssw_subreads``).  ``align_many`` and ``align_subreads`` are the
batched forms."""
import numpy as np

from . import _native
from ._native import TASK_DT

_OPS = " MID"


class AlignmentStructure(object):
    def __init__(self, r, cigar_runs, query, target):
        aligned = r["r_pos"] >= 0
        self.optimal_alignment_score = int(r["score"])
        self.query_begin = int(r["q_pos"]) if aligned else -1
        self.query_end = int(r["q_end"]) - 1 if aligned else -1
        self.target_begin = int(r["r_pos"]) if aligned else -1
        self.target_end_optimal = int(r["r_end"]) - 1 if aligned else -1
        self.query_sequence = query
        self.target_sequence = target
        self.matches = int(r["matches"])
        self._runs = [(int(c >> 2), _OPS[int(c & 3)]) for c in cigar_runs]
        self.cigar = "".join("%d%s" % x for x in self._runs)

    def _aligned(self, seq, begin, end, gap_type):
        # skbio _get_aligned_sequence: its own gap type writes '-', everything else consumes the sequence; what the CIGAR did
        # not consume of [begin, end] is appended
        s = seq[begin:end + 1]
        out, index = [], 0
        for length, op in self._runs:
            if op == gap_type:
                out.append('-' * length)
            else:
                out.append(s[index:index + length])
                index += length
        out.append(s[index:end - begin + 1])
        return "".join(out)

    @property
    def aligned_query_sequence(self):
        return self._aligned(self.query_sequence, self.query_begin, self.query_end, 'D')

    @property
    def aligned_target_sequence(self):
        return self._aligned(self.target_sequence, self.target_begin, self.target_end_optimal, 'I')

    def __getitem__(self, key):
        return getattr(self, key)


def align_many(queries, targets, intervals=None, device=0):
    """Batched StripedSmithWaterman(query)(target[start:end]).  intervals: None -> every query against every whole target
    (target-major order), or (target, start, end, query) rows.  -> list of AlignmentStructure in task order."""
    from . import swalign as _sw
    queries = [str(q) for q in queries]
    targets = [str(t) for t in targets]
    eng = _sw.get_engine(device)
    eng.set_ssw(2, -3, 5, 2)
    eng.set_queries(queries)
    eng.upload_read_list(targets)
    if intervals is None:
        tasks = np.zeros(len(targets) * len(queries), TASK_DT)
        tasks["read"] = np.repeat(np.arange(len(targets)), len(queries))
        tasks["query"] = np.tile(np.arange(len(queries)), len(targets))
        tasks["end"] = np.repeat([len(t) for t in targets], len(queries))
    else:
        iv = np.asarray(intervals).reshape(-1, 4)
        tasks = np.zeros(len(iv), TASK_DT)
        tasks["read"], tasks["start"], tasks["end"], tasks["query"] = iv[:, 0], iv[:, 1], iv[:, 2], iv[:, 3]
    res, cig = eng.align_tasks(tasks)
    return [AlignmentStructure(r, cig[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]], queries[t["query"]],
                               targets[t["read"]][t["start"]:t["end"]]) for r, t in zip(res, tasks)]


class StripedSmithWaterman(object):
    def __init__(self, query_sequence, gap_open_penalty=5, gap_extend_penalty=2, score_size=2, mask_length=15, mask_auto=True,
                 score_only=False, score_filter=None, distance_filter=None, override_skip_babp=False, protein=False,
                 match_score=2, mismatch_score=-3, substitution_matrix=None, suppress_sequences=False, zero_index=True, device=0):
        if (gap_open_penalty, gap_extend_penalty, match_score, mismatch_score) != (5, 2, 2, -3) or protein or \
                substitution_matrix is not None or score_only or score_filter or distance_filter or not zero_index:
            raise NotImplementedError("only skbio's default StripedSmithWaterman parameters (the LAMPrey path) run on the GPU")
        self.query_sequence = str(query_sequence)
        self.device = device

    def __call__(self, target_sequence):
        if len(self.query_sequence) > _native.MAX_QUERY_LEN:
            n = len(self.query_sequence)
            return align_subreads([self.query_sequence], [(0, 0, n, 0)], str(target_sequence), device=self.device)[0]
        return align_many([self.query_sequence], [str(target_sequence)], device=self.device)[0]


_COMP = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}


def align_subreads(reads, pairs, target, device=0, engine=None):
    """Batched stage 2 of processLamplicon (lamprey.py:1637-1658): for every (read, start, end, revcomp) row, the sub-read
    read[start:end] -- reverse-complemented if revcomp -- is the QUERY of StripedSmithWaterman, `target` (the target reference
    string) the target.  reads: list of str, or None to use the batch already uploaded to `engine` (then the AlignmentStructure
    objects carry no sequences).  -> list of AlignmentStructure."""
    from . import swalign as _sw
    eng = engine or _sw.get_engine(device)
    if reads is not None:
        reads = [str(r) for r in reads]
        eng.upload_read_list(reads)
    p = np.zeros(len(pairs), _native.PAIR_DT)
    for k, row in enumerate(pairs):
        p[k] = tuple(int(v) for v in row)
    res, cig = eng.ssw_subreads(p, target)
    out = []
    for r, row in zip(res, p):
        q = ""
        if reads is not None:
            q = reads[row["read"]][row["start"]:row["end"]]
            if row["revcomp"]:
                q = ''.join(_COMP.get(b, b) for b in reversed(q))
        out.append(AlignmentStructure(r, cig[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]], q, str(target)))
    return out
