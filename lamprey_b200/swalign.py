"""Drop-in replacement for the third-party ``swalign`` module LAMPrey imports
(/root/reference/lamprey.py:14-16, constructed at 1860-1865, called at 563 and 672-694).

Same surface -- ``NucleotideScoringMatrix``, ``LocalAlignment(...).align(ref, query)``,
``Alignment`` with ``score q_pos q_end r_pos r_end cigar cigar_str matches mismatches identity
query ref dump()`` -- computed by the sm_100a kernels in liblsw.so, plus the batched entry points
``align_many`` and ``find_all_primer_alignments``.  Only LAMPrey's parameterisation runs on the
GPU (local mode, gap_penalty == gap_extension_penalty, no decay, no wildcard); anything else
raises ``NotImplementedError`` instead of silently differing.  There is no CPU fallback.
"""
import sys

import numpy as np

from . import _native
from ._native import TASK_DT

_OPS = " MID"

__all__ = ["NucleotideScoringMatrix", "LocalAlignment", "Alignment", "AlignmentBatch", "align_many",
           "find_all_primer_alignments"]


class NucleotideScoringMatrix(object):
    def __init__(self, match=1, mismatch=-1):
        self.match = match
        self.mismatch = mismatch

    def score(self, one, two, wildcard=None):
        if wildcard and (one in wildcard or two in wildcard):
            return self.match
        return self.match if one == two else self.mismatch


def _decode_cigar(runs):
    return [(int(c) >> 2, _OPS[int(c) & 3]) for c in runs]


class Alignment(object):
    """Result of one local alignment; field for field what swalign.Alignment exposes."""

    def __init__(self, query, ref, q_pos, r_pos, cigar, score, ref_name='', query_name='', rc=False,
                 globalalign=False, wildcard=None, _counts=None):
        self.q_pos = q_pos
        self.r_pos = r_pos
        self.cigar = cigar
        self.score = score
        self.r_name = ref_name
        self.q_name = query_name
        self.rc = rc
        self.globalalign = globalalign
        self.wildcard = wildcard
        self.r_offset = 0
        self.r_region = None
        self.orig_query = query
        self.query = query.upper()
        self.orig_ref = ref
        self.ref = ref.upper()

        q_len = sum(n for n, op in cigar if op in 'MI')
        r_len = sum(n for n, op in cigar if op in 'MD')
        if _counts is not None:                 # counted on the device during the traceback
            self.matches, self.mismatches = _counts
        else:
            self.matches = self.mismatches = 0
            i, j = r_pos, q_pos
            for n, op in cigar:
                if op == 'M':
                    for _ in range(n):
                        if self.query[j] == self.ref[i]:
                            self.matches += 1
                        else:
                            self.mismatches += 1
                        i += 1
                        j += 1
                elif op == 'I':
                    j += n
                    self.mismatches += n
                elif op == 'D':
                    i += n
                    self.mismatches += n
        self.q_end = q_pos + q_len
        self.r_end = r_pos + r_len
        if self.mismatches + self.matches > 0:
            self.identity = float(self.matches) / (self.mismatches + self.matches)
        else:
            self.identity = 0

    def set_ref_offset(self, ref, offset, region):
        self.r_name = ref
        self.r_offset = offset
        self.r_region = region

    @property
    def cigar_str(self):
        return ''.join('%s%s' % (n, op) for n, op in self.cigar)

    @property
    def extended_cigar_str(self):
        ext, qpos, rpos = [], 0, 0
        for n, op in self.cigar:
            if op == 'M':
                for k in range(n):
                    same = self.query[self.q_pos + qpos + k] == self.ref[self.r_pos + rpos + k]
                    ext.append('M' if same else 'X')
                qpos += n
                rpos += n
            elif op == 'I':
                qpos += n
                ext.extend('I' * n)
            elif op == 'D':
                rpos += n
                ext.extend('D' * n)
        out, last, cnt = [], None, 0
        for op in ext:
            if op != last and last is not None:
                out.append('%s%s' % (cnt, last))
                cnt = 0
            last = op
            cnt += 1
        if last is not None:
            out.append('%s%s' % (cnt, last))
        return ''.join(out)

    def dump(self, wrap=None, out=sys.stdout):
        i, j = self.r_pos, self.q_pos
        q = m = r = ''
        for n, op in self.cigar:
            if op == 'M':
                for _ in range(n):
                    q += self.orig_query[j]
                    r += self.orig_ref[i]
                    m += '|' if self.query[j] == self.ref[i] else '.'
                    i += 1
                    j += 1
            elif op == 'D':
                for _ in range(n):
                    q += '-'
                    r += self.orig_ref[i]
                    m += ' '
                    i += 1
            elif op == 'I':
                for _ in range(n):
                    q += self.orig_query[j]
                    r += '-'
                    m += ' '
                    j += 1
        if self.q_name:
            out.write('Query: %s%s (%s nt)\n' % (self.q_name, ' (reverse-compliment)' if self.rc else '', len(self.query)))
        if self.r_name:
            if self.r_region:
                out.write('Ref  : %s (%s)\n\n' % (self.r_name, self.r_region))
            else:
                out.write('Ref  : %s (%s nt)\n\n' % (self.r_name, len(self.ref)))
        width = max(len(str(v)) for v in (self.q_pos + 1, self.q_end + 1, self.r_pos + self.r_offset + 1,
                                          self.r_end + self.r_offset + 1))
        q_pre, r_pre, m_pre = 'Query: %%%ss ' % width, 'Ref  : %%%ss ' % width, ' ' * (8 + width)
        rpos = self.r_pos
        qpos = self.q_end if self.rc else self.q_pos
        while q and r and m:
            out.write(q_pre % (qpos if self.rc else qpos + 1))
            if wrap:
                qf, mf, rf = q[:wrap], m[:wrap], r[:wrap]
                q, m, r = q[wrap:], m[wrap:], r[wrap:]
            else:
                qf, mf, rf = q, m, r
                q = m = r = ''
            out.write(qf)
            for base in qf:
                if base != '-':
                    qpos += -1 if self.rc else 1
            out.write(' %s\n' % (qpos + 1 if self.rc else qpos))
            out.write(m_pre + mf + '\n')
            out.write(r_pre % (rpos + self.r_offset + 1))
            out.write(rf)
            for base in rf:
                if base != '-':
                    rpos += 1
            out.write(' %s\n\n' % (rpos + self.r_offset))
        out.write('Score: %s\n' % self.score)
        out.write('Matches: %s (%.1f%%)\n' % (self.matches, self.identity * 100))
        out.write('Mismatches: %s\n' % (self.mismatches,))
        out.write('CIGAR: %s\n' % self.cigar_str)


class AlignmentBatch(object):
    """Struct-of-arrays result of ``align_many``; ``batch[i]`` materialises one ``Alignment``."""

    def __init__(self, results, cigar, refs, queries, tasks):
        self.results = results
        self.cigar = cigar
        self._refs, self._queries, self._tasks = refs, queries, tasks
        for name in ("score", "q_pos", "q_end", "r_pos", "r_end", "matches", "mismatches", "n_cigar", "cigar_off"):
            setattr(self, name, results[name])

    def __len__(self):
        return len(self.results)

    @property
    def identity(self):
        tot = self.matches + self.mismatches
        out = np.zeros(len(tot), np.float64)
        np.divide(self.matches, tot, out=out, where=tot > 0)
        return out

    def cigar_of(self, i):
        r = self.results[i]
        return _decode_cigar(self.cigar[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]])

    def __getitem__(self, i):
        r, t = self.results[i], self._tasks[i]
        ref = self._refs[t["read"]][t["start"]:t["end"]]
        return Alignment(self._queries[t["query"]], ref, int(r["q_pos"]), int(r["r_pos"]), self.cigar_of(i),
                         int(r["score"]), _counts=(int(r["matches"]), int(r["mismatches"])))


_engines = {}


def get_engine(device=0):
    """One lazily created context per device (CUDA is initialised on first use, never at import,
    because LAMPrey builds its aligner inside forked workers: lamprey.py:1886)."""
    eng = _engines.get(device)
    if eng is None:
        eng = _engines[device] = _native.Engine(device)
    return eng


def _as_text(s):
    return s if isinstance(s, str) else bytes(s).decode("ascii")


def align_many(reads, primers, intervals=None, scoring=(2, -1), gap_penalty=-1, device=0, want_cigar=True):
    """Batched ``align(ref=read[start:end], query=primer)``.

    intervals: None -> every primer against every whole read (read-major order), or an array of
    (read, start, end, primer) rows.  Returns an ``AlignmentBatch``.
    """
    reads = [_as_text(r) for r in reads]
    primers = [_as_text(p) for p in primers]
    eng = get_engine(device)
    eng.set_scoring(scoring[0], scoring[1], gap_penalty, gap_penalty)
    eng.set_queries(primers)
    eng.upload_read_list(reads)
    if intervals is None:
        tasks = np.zeros(len(reads) * len(primers), TASK_DT)
        tasks["read"] = np.repeat(np.arange(len(reads)), len(primers))
        tasks["query"] = np.tile(np.arange(len(primers)), len(reads))
        tasks["end"] = np.repeat([len(r) for r in reads], len(primers))
    else:
        iv = np.asarray(intervals)
        if iv.dtype != TASK_DT:
            iv = iv.reshape(-1, 4)
            tasks = np.zeros(len(iv), TASK_DT)
            tasks["read"], tasks["start"], tasks["end"], tasks["query"] = iv[:, 0], iv[:, 1], iv[:, 2], iv[:, 3]
        else:
            tasks = iv
    res, cig = eng.align_tasks(tasks, want_cigar=want_cigar)
    return AlignmentBatch(res, cig, reads, primers, tasks)


class LocalAlignment(object):
    """swalign.LocalAlignment with the constructor signature LAMPrey relies on (lamprey.py:1865)."""

    def __init__(self, scoring_matrix, gap_penalty=-1, gap_extension_penalty=-1, gap_extension_decay=0.0,
                 prefer_gap_runs=True, verbose=False, globalalign=False, wildcard=None, full_query=False,
                 device=0):
        if globalalign or full_query:
            raise NotImplementedError("only local alignment (the LAMPrey path) runs on the GPU")
        if gap_extension_decay:
            raise NotImplementedError("gap_extension_decay is not implemented")
        if wildcard:
            raise NotImplementedError("wildcard is not implemented")
        if not prefer_gap_runs:
            raise NotImplementedError("prefer_gap_runs=False is not implemented")
        if gap_penalty != gap_extension_penalty:
            raise NotImplementedError("gap_penalty != gap_extension_penalty is not implemented")
        if not isinstance(scoring_matrix, NucleotideScoringMatrix) and not (
                hasattr(scoring_matrix, "match") and hasattr(scoring_matrix, "mismatch")):
            raise NotImplementedError("only NucleotideScoringMatrix(match, mismatch) scoring is implemented")
        self.scoring_matrix = scoring_matrix
        self.gap_penalty = gap_penalty
        self.gap_extension_penalty = gap_extension_penalty
        self.gap_extension_decay = gap_extension_decay
        self.prefer_gap_runs = prefer_gap_runs
        self.verbose = verbose
        self.globalalign = globalalign
        self.wildcard = wildcard
        self.full_query = full_query
        self.device = device

    def _scoring(self):
        return (self.scoring_matrix.match, self.scoring_matrix.mismatch)

    def align(self, ref, query, ref_name='', query_name='', rc=False):
        batch = align_many([ref], [query], scoring=self._scoring(), gap_penalty=self.gap_penalty, device=self.device)
        r = batch.results[0]
        return Alignment(query, ref, int(r["q_pos"]), int(r["r_pos"]), batch.cigar_of(0), int(r["score"]), ref_name,
                         query_name, rc, self.globalalign, self.wildcard,
                         _counts=(int(r["matches"]), int(r["mismatches"])))

    def align_many(self, reads, primers, intervals=None, want_cigar=True):
        return align_many(reads, primers, intervals, self._scoring(), self.gap_penalty, self.device, want_cigar)


def find_all_primer_alignments(reads, primer_set, threshold, aligner=None, device=0):
    """Batched findAllPrimerAlignments (lamprey.py:699-733): see lamprey_b200.seeding.seed_reads."""
    from .seeding import seed_reads
    return seed_reads(reads, primer_set, threshold, aligner=aligner, device=device)
