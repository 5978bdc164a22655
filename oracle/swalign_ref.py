# cython: language_level=3
"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement of the third-party module the reference's hot path runs:
``swalign == 0.3.6`` (upstream mbreese/swalign, pinned in
/root/reference/requirements.txt:1; loaded at /root/reference/lamprey.py:14-16,
constructed at lamprey.py:1860-1865, called at lamprey.py:563).  The package is
NOT vendored in /root/reference and is not installable here (no network), so
this file restates its published algorithm and is
anchored on the reference's call sites and the one known-answer vector the
reference copies from the upstream README (lamprey.py:1860-1865):

    align('ACACACTA', 'AGCACACA') -> score 12, CIGAR 1M1I5M1D1M, 7 matches,
    2 mismatches, identity 0.7778

PARITY UNPINNED: the reference ships no tests or golden vectors for this path;
apart from the README vector above nothing in /root/reference pins the results.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The same file compiled with
``cythonize -3`` (untyped, exactly the recipe /root/reference/README.md:18-25
prescribes for the accelerated fork) is the "Cython swalign" CPU baseline.

Surface restated (only what LAMPrey consumes, lamprey.py:563-566, 672-694):
``NucleotideScoringMatrix``, ``LocalAlignment(...).align(ref, query)``,
``Alignment`` with ``score q_pos q_end r_pos r_end cigar cigar_str matches
mismatches identity query ref dump()``.
"""
import sys


class NucleotideScoringMatrix(object):
    """match / mismatch scorer."""

    def __init__(self, match=1, mismatch=-1):
        self.match = match
        self.mismatch = mismatch

    def score(self, one, two, wildcard=None):
        if wildcard and (one in wildcard or two in wildcard):
            return self.match
        if one == two:
            return self.match
        return self.mismatch


def _reduce_cigar(operations):
    """Run-length encode a sequence of 'm'/'i'/'d' ops into [(count, 'M'|'I'|'D')]."""
    out = []
    count = 0
    last = None
    for op in operations:
        if last is not None and op != last:
            out.append((count, last.upper()))
            count = 0
        last = op
        count += 1
    if last is not None:
        out.append((count, last.upper()))
    return out


class LocalAlignment(object):
    """Single-matrix local aligner whose cells are (value, op, run) triples.

    Restates SURVEY.md Appendix A.1 in its general form (``gap_penalty`` and
    ``gap_extension_penalty`` kept separate, ``gap_extension_decay`` honoured) so
    a later correction of a remembered default is a one-constant change.
    Global / full_query modes are outside the hot path and raise.
    """

    def __init__(self, scoring_matrix, gap_penalty=-1, gap_extension_penalty=-1,
                 gap_extension_decay=0.0, prefer_gap_runs=True, verbose=False,
                 globalalign=False, wildcard=None, full_query=False):
        if globalalign or full_query:
            raise NotImplementedError("oracle restates local mode only (the LAMPrey path)")
        self.scoring_matrix = scoring_matrix
        self.gap_penalty = gap_penalty
        self.gap_extension_penalty = gap_extension_penalty
        self.gap_extension_decay = gap_extension_decay
        self.prefer_gap_runs = prefer_gap_runs
        self.verbose = verbose
        self.globalalign = globalalign
        self.wildcard = wildcard
        self.full_query = full_query

    def align(self, ref, query, ref_name='', query_name='', rc=False):
        orig_ref = ref
        orig_query = query
        ref = ref.upper()
        query = query.upper()

        nrows = len(query) + 1
        ncols = len(ref) + 1
        gap = self.gap_penalty
        gap_ext = self.gap_extension_penalty
        decay = self.gap_extension_decay
        prefer_runs = self.prefer_gap_runs
        score = self.scoring_matrix.score
        wildcard = self.wildcard

        # three parallel flat planes instead of a list of tuples: value, op, run
        size = nrows * ncols
        vals = [0] * size
        ops = [' '] * size
        runs = [0] * size
        for row in range(1, nrows):
            ops[row * ncols] = 'i'
        for col in range(1, ncols):
            ops[col] = 'd'

        max_val = 0
        max_row = 0
        max_col = 0

        for row in range(1, nrows):
            qch = query[row - 1]
            base = row * ncols
            above = base - ncols
            for col in range(1, ncols):
                mm_val = vals[above + col - 1] + score(qch, ref[col - 1], wildcard)

                ins_run = 0
                del_run = 0

                up_val = vals[above + col]
                if ops[above + col] == 'i':
                    ins_run = runs[above + col]
                    if up_val == 0:
                        ins_val = 0          # no penalty to start the alignment
                    elif not decay:
                        ins_val = up_val + gap_ext
                    else:
                        ins_val = up_val + min(0, gap_ext + ins_run * decay)
                else:
                    ins_val = up_val + gap

                left_val = vals[base + col - 1]
                if ops[base + col - 1] == 'd':
                    del_run = runs[base + col - 1]
                    if left_val == 0:
                        del_val = 0          # no penalty to start the alignment
                    elif not decay:
                        del_val = left_val + gap_ext
                    else:
                        del_val = left_val + min(0, gap_ext + del_run * decay)
                else:
                    del_val = left_val + gap

                cell_val = max(mm_val, del_val, ins_val, 0)

                if not prefer_runs:
                    ins_run = 0
                    del_run = 0

                if del_run and cell_val == del_val:
                    op = 'd'
                    run = del_run + 1
                elif ins_run and cell_val == ins_val:
                    op = 'i'
                    run = ins_run + 1
                elif cell_val == mm_val:
                    op = 'm'
                    run = 0
                elif cell_val == del_val:
                    op = 'd'
                    run = 1
                elif cell_val == ins_val:
                    op = 'i'
                    run = 1
                else:
                    cell_val = 0
                    op = 'x'
                    run = 0

                if cell_val >= max_val:
                    max_val = cell_val
                    max_row = row
                    max_col = col

                vals[base + col] = cell_val
                ops[base + col] = op
                runs[base + col] = run

        # traceback from the arg-max until the first non-positive cell
        row = max_row
        col = max_col
        aln = []
        while True:
            val = vals[row * ncols + col]
            op = ops[row * ncols + col]
            if val <= 0:
                break
            aln.append(op)
            if op == 'm':
                row -= 1
                col -= 1
            elif op == 'i':
                row -= 1
            elif op == 'd':
                col -= 1
            else:
                break
        aln.reverse()

        cigar = _reduce_cigar(aln)
        return Alignment(orig_query, orig_ref, row, col, cigar, max_val, ref_name,
                         query_name, rc, self.globalalign, self.wildcard)


class Alignment(object):
    """Result record."""

    def __init__(self, query, ref, q_pos, r_pos, cigar, score, ref_name='',
                 query_name='', rc=False, globalalign=False, wildcard=None):
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

        q_len = 0
        r_len = 0
        self.matches = 0
        self.mismatches = 0

        i = self.r_pos
        j = self.q_pos
        for count, op in self.cigar:
            if op == 'M':
                q_len += count
                r_len += count
                for _ in range(count):
                    if self.query[j] == self.ref[i]:
                        self.matches += 1
                    else:
                        self.mismatches += 1
                    i += 1
                    j += 1
            elif op == 'I':
                q_len += count
                j += count
                self.mismatches += count
            elif op == 'D':
                r_len += count
                i += count
                self.mismatches += count

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
        qpos = 0
        rpos = 0
        ext = []
        for count, op in self.cigar:
            if op == 'M':
                for k in range(count):
                    same = self.query[self.q_pos + qpos + k] == self.ref[self.r_pos + rpos + k]
                    ext.append('M' if same else 'X')
                qpos += count
                rpos += count
            elif op == 'I':
                qpos += count
                ext.extend('I' * count)
            elif op == 'D':
                rpos += count
                ext.extend('D' * count)
        return ''.join('%s%s' % (n, op) for n, op in _reduce_cigar(ext))

    def dump(self, wrap=None, out=sys.stdout):
        i = self.r_pos
        j = self.q_pos
        q = ''
        m = ''
        r = ''
        for count, op in self.cigar:
            if op == 'M':
                for _ in range(count):
                    q += self.orig_query[j]
                    r += self.orig_ref[i]
                    m += '|' if self.query[j] == self.ref[i] else '.'
                    i += 1
                    j += 1
            elif op == 'D':
                for _ in range(count):
                    q += '-'
                    r += self.orig_ref[i]
                    m += ' '
                    i += 1
            elif op == 'I':
                for _ in range(count):
                    q += self.orig_query[j]
                    r += '-'
                    m += ' '
                    j += 1

        if self.q_name:
            out.write('Query: %s%s (%s nt)\n' % (
                self.q_name, ' (reverse-compliment)' if self.rc else '', len(self.query)))
        if self.r_name:
            if self.r_region:
                out.write('Ref  : %s (%s)\n\n' % (self.r_name, self.r_region))
            else:
                out.write('Ref  : %s (%s nt)\n\n' % (self.r_name, len(self.ref)))

        poslens = [self.q_pos + 1, self.q_end + 1,
                   self.r_pos + self.r_offset + 1, self.r_end + self.r_offset + 1]
        maxlen = max(len(str(x)) for x in poslens)
        q_pre = 'Query: %%%ss ' % maxlen
        r_pre = 'Ref  : %%%ss ' % maxlen
        m_pre = ' ' * (8 + maxlen)

        rpos = self.r_pos
        qpos = self.q_end if self.rc else self.q_pos

        while q and r and m:
            if not self.rc:
                out.write(q_pre % (qpos + 1))
            else:
                out.write(q_pre % qpos)
            if wrap:
                qfrag, mfrag, rfrag = q[:wrap], m[:wrap], r[:wrap]
                q, m, r = q[wrap:], m[wrap:], r[wrap:]
            else:
                qfrag, mfrag, rfrag = q, m, r
                q = m = r = ''

            out.write(qfrag)
            step = -1 if self.rc else 1
            for base in qfrag:
                if base != '-':
                    qpos += step
            if not self.rc:
                out.write(' %s\n' % qpos)
            else:
                out.write(' %s\n' % (qpos + 1))

            out.write(m_pre)
            out.write(mfrag)
            out.write('\n')
            out.write(r_pre % (rpos + self.r_offset + 1))
            out.write(rfrag)
            for base in rfrag:
                if base != '-':
                    rpos += 1
            out.write(' %s\n\n' % (rpos + self.r_offset))

        out.write('Score: %s\n' % self.score)
        out.write('Matches: %s (%.1f%%)\n' % (self.matches, self.identity * 100))
        out.write('Mismatches: %s\n' % (self.mismatches,))
        out.write('CIGAR: %s\n' % self.cigar_str)
