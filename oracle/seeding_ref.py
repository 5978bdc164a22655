# cython: language_level=3
"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement of LAMPrey's greedy recursive primer search ("seed" stage):

* ``rc``                       /root/reference/lamprey.py:649-659
* ``PrimerSet``                /root/reference/lamprey.py:81-136
* ``Alignment`` (hit record)   /root/reference/lamprey.py:138-161
* ``align_sequence``           /root/reference/lamprey.py:558-569  (alignSequence_swalign)
* ``find_primer_alignments``   /root/reference/lamprey.py:596-645  (findPrimerAlignments, --swalign branch)
* ``find_all_primer_alignments`` /root/reference/lamprey.py:699-733 (findAllPrimerAlignments)

lamprey.py itself cannot be imported here (skbio, mappy, bwapy, pysam, vcfpy,
statsmodels, Bio, progressbar, watchdog are all absent - lamprey.py:10-36), so
these functions are restated; behaviour (argument meaning, ordering, the Python
float threshold test, the pre-order hit list, the stable sort on ``start``)
follows the cited lines.  PARITY UNPINNED by reference tests (there are none).
"""

_COMPLEMENT = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}


def rc(seq):
    """Reverse complement; characters outside ACGT pass through unchanged."""
    return ''.join(_COMPLEMENT.get(ch, ch) for ch in reversed(seq))


class PrimerSet(object):
    """LAMP schema file: ``<name> <sequence>`` lines, ``#`` comments, two order lines.

    A name ending in ``c`` is stored reverse-complemented under the stripped name.
    """

    def __init__(self, primer_set_filename=None, lines=None):
        self.primer_dict = dict()
        self.fwd_primer_order = ['F3', 'F2', 'F1', 'T', 'B1c', 'B2c', 'B3c']
        self.rev_primer_order = ['B3', 'B2', 'B1', 'Tc', 'F1c', 'F2c', 'F3c']
        if lines is None:
            with open(primer_set_filename) as fp:
                lines = fp.readlines()
        for line in lines:
            line = line.strip()
            if not line or line.startswith('#'):
                continue
            fields = line.split(' ')
            name = fields[0]
            if name == 'fwd_primer_order':
                self.fwd_primer_order = list(fields[1:])
                continue
            if name == 'rev_primer_order':
                self.rev_primer_order = list(fields[1:])
                continue
            seq = fields[1]
            if name.endswith('c'):
                self.primer_dict[name[:-1]] = rc(seq)
            else:
                self.primer_dict[name] = seq

    def sequences(self):
        """The (name, sequence) list in the order findAllPrimerAlignments visits it."""
        out = [('T', self.primer_dict['T']), ('Tc', rc(self.primer_dict['T']))]
        for name, seq in self.primer_dict.items():
            if name == 'T':
                continue
            out.append((name, seq))
            out.append((name + 'c', rc(seq)))
        return out


class Alignment(object):
    """Seed hit: [start, end) on the read, identity, schema sequence name."""

    def __init__(self, start=0, end=0, identity=0.0, primer_name=''):
        self.start = start
        self.end = end
        self.identity = identity
        self.primer_name = primer_name

    def __repr__(self):
        return "{} at ({}, {}) identity: {}".format(
            self.primer_name, self.start, self.end, self.identity)

    __str__ = __repr__

    def __lt__(self, other):
        return self.start < other.start

    def __eq__(self, other):          # the reference's __eq__ returns None
        self.start == other.start

    __hash__ = object.__hash__

    def astuple(self):
        return (self.start, self.end, self.identity, self.primer_name)


class Stats(object):
    """Work counters the reference does not keep: align() calls and DP cells."""

    def __init__(self):
        self.calls = 0
        self.cells = 0
        self.max_depth = 0


def align_sequence(aligner, seq, primer, primer_name):
    """ref = read slice, query = primer; keep r_pos, r_end, identity."""
    a = aligner.align(str(seq), primer)
    return Alignment(a.r_pos, a.r_end, a.identity, primer_name)


def find_primer_alignments(aligner, seq, primer, primer_name, identity_threshold,
                           stats=None, _depth=0):
    """Best hit, then recurse on what lies left and right of it (pre-order list)."""
    if stats is not None:
        stats.calls += 1
        stats.cells += len(seq) * len(primer)
        if _depth > stats.max_depth:
            stats.max_depth = _depth
    best = align_sequence(aligner, seq, primer, primer_name)
    span = best.end - best.start

    hits = []
    if best.identity >= identity_threshold and span >= len(primer) * identity_threshold:
        hits.append(best)
        left = seq[:best.start]
        right = seq[best.end:]
        if len(left) >= len(primer):
            hits.extend(find_primer_alignments(aligner, left, primer, primer_name,
                                               identity_threshold, stats, _depth + 1))
        if len(right) >= len(primer):
            shifted = find_primer_alignments(aligner, right, primer, primer_name,
                                             identity_threshold, stats, _depth + 1)
            for h in shifted:
                h.start += best.end
                h.end += best.end
            hits.extend(shifted)
    return hits


def find_all_primer_alignments(aligner, seq, primers, identity_threshold, stats=None):
    """T, Tc, then every other schema entry forward and reverse-complemented;
    stable sort on start; coverage = covered bases / read length."""
    hits = []
    for name, pseq in primers.sequences():
        hits.extend(find_primer_alignments(aligner, seq, pseq, name,
                                           identity_threshold, stats))
    hits = sorted(hits)
    covered = 0
    for h in hits:
        covered += h.end - h.start
    coverage = float(covered) / float(len(seq))
    return hits, coverage
