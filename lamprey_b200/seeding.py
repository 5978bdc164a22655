"""Host-side mirror of LAMPrey's seed stage, backed by the batched GPU engine.

Mirrors, name for name and argument for argument (same ordering and error behaviour):

* ``rc``                         /root/reference/lamprey.py:649-659
* ``PrimerSet``                  /root/reference/lamprey.py:81-136
* ``Alignment``                  /root/reference/lamprey.py:138-161
* ``alignSequence_swalign``      /root/reference/lamprey.py:558-569
* ``alignSequence_skbio``        /root/reference/lamprey.py:572-593   (the default branch: scikit-bio's StripedSmithWaterman)
* ``findPrimerAlignments``       /root/reference/lamprey.py:596-645   (both branches: ``args.swalign`` decides, as in the reference)
* ``findAllPrimerAlignments``    /root/reference/lamprey.py:699-733

and adds the batched form the reference lacks: ``seed_reads(reads, primer_set, threshold)`` seeds a
whole FASTQ in one device pass (recursion included) and returns a ``SeedBatch`` (struct of arrays;
``batch[i]`` gives exactly what ``findAllPrimerAlignments`` returns for read ``i``); ``seed_fastq(path, ...)`` does the
same straight from a FASTQ file.
The arithmetic runs only in liblsw.so; nothing here computes an alignment on the CPU.
"""
import numpy as np

from . import _native
from ._native import HIT_DT

_COMPLEMENT = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}


def rc(seq):
    """Reverse complement; characters outside ACGT pass through unchanged."""
    return ''.join(_COMPLEMENT.get(b, b) for b in reversed(seq))


class PrimerSet(object):
    """LAMP schema: ``<name> <sequence>`` lines, ``#`` comments, fwd/rev primer-order lines.
    Names ending in ``c`` are stored reverse-complemented under the stripped name."""

    def __init__(self, primer_set_filename=None, lines=None, verbose=False):
        self.primer_dict = dict()
        self.fwd_primer_order = ['F3', 'F2', 'F1', 'T', 'B1c', 'B2c', 'B3c']
        self.rev_primer_order = ['B3', 'B2', 'B1', 'Tc', 'F1c', 'F2c', 'F3c']
        if lines is None:
            with open(primer_set_filename) as fp:
                lines = fp.readlines()
        for line in lines:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            fields = line.split(' ')
            name = fields[0]
            if name == 'fwd_primer_order':
                self.fwd_primer_order = list(fields[1:])
            elif name == 'rev_primer_order':
                self.rev_primer_order = list(fields[1:])
            elif name.endswith('c'):
                self.primer_dict[name[:-1]] = rc(fields[1])
            else:
                self.primer_dict[name] = fields[1]
        if verbose:                      # the reference prints both order lists on construction
            print(self.fwd_primer_order)
            print(self.rev_primer_order)

    def sequences(self):
        """[(name, sequence)] in the order findAllPrimerAlignments searches them:
        T, Tc, then every other entry forward and reverse-complemented (lamprey.py:703-719)."""
        out = [('T', self.primer_dict['T']), ('Tc', rc(self.primer_dict['T']))]
        for name, seq in self.primer_dict.items():
            if name != 'T':
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
        return "{} at ({}, {}) identity: {}".format(self.primer_name, self.start, self.end, self.identity)

    __str__ = __repr__

    def __lt__(self, other):
        return self.start < other.start

    def __eq__(self, other):              # as in the reference: evaluates and returns None
        self.start == other.start

    __hash__ = object.__hash__


def _read_arrays(reads):
    """list of str/bytes, or (uint8 concat, int64 offsets) -> (concat, offsets)."""
    if isinstance(reads, tuple) and len(reads) == 2 and isinstance(reads[0], np.ndarray):
        return np.ascontiguousarray(reads[0], np.uint8), np.ascontiguousarray(reads[1], np.int64)
    return _native.concat_sequences([str(r) for r in reads])


class SeedBatch(object):
    """Seeds of a read batch.  ``hits`` is a HIT_DT array ordered by (read, start, search order) --
    the order lamprey.py:722's stable sort produces; ``offsets[i]:offsets[i+1]`` are read i's hits."""

    def __init__(self, hits, n_reads, read_lengths, names, counters, qlen=None):
        self.hits = hits
        self.names = names
        self.counters = counters
        self.read_lengths = read_lengths
        self._qlen = qlen
        self._identity = None
        # where every read's hits start, and the bases they cover: one multi-threaded pass in the library (lsw_hits_index)
        self.offsets, covered = _native.hits_index(hits, n_reads)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.coverage = covered / read_lengths.astype(np.float64)

    @property
    def identity(self):
        """Per hit, in Python doubles as the reference computes it (computed on first use)."""
        if self._identity is None:
            hits = self.hits
            if self._qlen is not None:    # skbio path: float(matches)/float(len(primer)), lamprey.py:585
                self._identity = hits["matches"] / self._qlen[hits["query"]].astype(np.float64)
            else:
                tot = (hits["matches"].astype(np.int64) + hits["mismatches"])
                self._identity = hits["matches"] / np.maximum(tot, 1)        # tot > 0 for every hit
        return self._identity

    def __len__(self):
        return len(self.offsets) - 1

    subreads = None          # SUBREAD_DT array ordered by (read, target_idx), when seed_reads(..., subreads=True) asked for it
    subread_offsets = None

    def _set_subreads(self, sub):
        self.subreads = sub
        self.subread_offsets = np.searchsorted(sub["read"], np.arange(len(self) + 1)).astype(np.int64)

    def target_subreads(self, i, lamp_idx=None):
        """Stage 1 of processLamplicon for read i (lamprey.py:1566-1589): [(seq_start, seq_end, subread_id, direction)], computed
        on the device (lsw_chain_subreads) -- what lamprey_b200.chain.target_subreads returns for the pruned hit list."""
        if self.subreads is None:
            raise ValueError("seed the batch with seed_reads(..., allowed_overlap=4, subreads=True)")
        lo, hi = self.subread_offsets[i], self.subread_offsets[i + 1]
        k = i if lamp_idx is None else lamp_idx
        return [(int(x["start"]), int(x["end"]), "{}_{}".format(k, int(x["target_idx"])), 'fwd' if x["direction"] == 0 else 'rev')
                for x in self.subreads[lo:hi]]

    @property
    def kept(self):
        """Mask of the hits that survive removeOverlappingPrimerAlignments (needs seed_reads(allowed_overlap=...))."""
        return self.hits["reserved"] == 1

    def target_depth(self, pruned=True):
        """result.target_depth of processLamplicon for EVERY read of the batch at once (lamprey.py:1547-1549): the number of
        'T' / 'Tc' hits, after removeOverlappingPrimerAlignments when pruned (needs seed_reads(allowed_overlap=...))."""
        is_t = np.isin(np.arange(len(self.names)), [k for k, n in enumerate(self.names) if n in ('T', 'Tc')])
        m = is_t[self.hits["query"]]
        if pruned:
            m &= self.kept
        return np.bincount(self.hits["read"][m], minlength=len(self)).astype(np.int64)

    def num_ont_seqs(self, threshold, pruned=True):
        """num_ont_seqs of processLamplicon for every read at once (lamprey.py:1551-1553): hits of sequences whose name starts with
        'ont' and whose identity reaches the threshold."""
        is_ont = np.array([n.startswith('ont') for n in self.names], bool)
        m = is_ont[self.hits["query"]] & (self.identity >= threshold)
        if pruned:
            m &= self.kept
        return np.bincount(self.hits["read"][m], minlength=len(self)).astype(np.int64)

    def alignments(self, i, pruned=False):
        lo, hi = self.offsets[i], self.offsets[i + 1]
        h = self.hits[lo:hi]
        if self._identity is not None:
            ident = self._identity[lo:hi]
        elif self._qlen is not None:
            ident = h["matches"] / self._qlen[h["query"]].astype(np.float64)
        else:
            ident = h["matches"] / np.maximum(h["matches"].astype(np.int64) + h["mismatches"], 1)
        start, end, query, keep = h["start"].tolist(), h["end"].tolist(), h["query"].tolist(), h["reserved"].tolist()
        names = self.names
        return [Alignment(start[k], end[k], float(ident[k]), names[query[k]]) for k in range(hi - lo) if not pruned or keep[k] == 1]

    def __getitem__(self, i):
        """(sorted alignment list, coverage) -- findAllPrimerAlignments' return value for read i."""
        if self.read_lengths[i] == 0:
            raise ZeroDivisionError("float division by zero")      # lamprey.py:731 on an empty read
        return self.alignments(i), float(self.coverage[i])


_pools = {}


def _device_subreads(eng, primer_set, names):
    """extractAmpliconAroundTarget for every surviving target hit of the batch the engine just seeded (lsw_chain_subreads)."""
    index = {n: k for k, n in enumerate(names)}
    return eng.chain_subreads([index.get(n, -1) for n in primer_set.fwd_primer_order],
                              [index.get(n, -1) for n in primer_set.rev_primer_order], index["T"], index["Tc"])


def _scoring_of(aligner):
    if aligner is None:
        return (2, -1, -1, -1)                          # initAligners, lamprey.py:1860-1865
    return (aligner.scoring_matrix.match, aligner.scoring_matrix.mismatch, aligner.gap_penalty, aligner.gap_extension_penalty)


SKBIO_DEFAULTS = (2, -3, 5, 2)       # StripedSmithWaterman: match_score, mismatch_score, gap_open_penalty, gap_extend_penalty


def _use_skbio(args, method):
    """Which branch of lamprey.py:607-610 to take: `args.swalign` decides, as in the reference (its default is False: skbio);
    without an args object the `method` keyword does ("swalign" unless told otherwise)."""
    if method is not None:
        if method not in ("swalign", "skbio"):
            raise ValueError("method must be 'swalign' or 'skbio'")
        return method == "skbio"
    return args is not None and hasattr(args, "swalign") and not args.swalign


def _set_engine_scoring(eng, aligner, skbio):
    if skbio:
        eng.set_ssw(*SKBIO_DEFAULTS)
    else:
        eng.set_scoring(*_scoring_of(aligner))


def get_pool(devices, contexts_per_device=1):
    """One lazily created pool per (device list, contexts per device)."""
    key = (tuple(int(d) for d in devices), int(contexts_per_device))
    pool = _pools.get(key)
    if pool is None:
        pool = _pools[key] = _native.Pool(key[0], key[1])
    return pool


def _seed_on_pool(pool, concat, offs, hit_cap=None, hit_buffer=None):
    """One batch through a configured pool: submit, wait, enlarge the hit buffer once if it was too small.
    hit_buffer: a caller-owned array of pool.hit_dtype records the hits are downloaded into (page-locked memory --
    _native.PinnedArray -- makes the download run at PCIe speed); the returned compact hits are expanded into new memory,
    wide ones are a view of it."""
    cap = int(hit_cap or (len(concat) // 8 + 4096))
    if hit_buffer is not None and (hit_buffer.dtype != pool.hit_dtype or hit_buffer.ndim != 1 or not hit_buffer.flags["C_CONTIGUOUS"]):
        raise ValueError("hit_buffer must be a contiguous 1-d array of %d-byte hit records (%s)" % (pool.hit_dtype.itemsize, pool.hit_dtype.names,))
    while True:
        out = hit_buffer if hit_buffer is not None and len(hit_buffer) >= 1 else np.zeros(cap, pool.hit_dtype)
        try:
            hits, counters, times = pool.wait(pool.submit(concat, offs, out))
        except OverflowError as e:
            cap = e.needed + 1024
            hit_buffer = None                            # the caller's buffer was too small for this batch: a pageable one of the right size
            continue
        if pool.hit_format == _native.HITS_COMPACT:
            hits = _native.expand_hits(hits, *pool.scoring)
        counters["pool_times"] = times
        return hits, counters


def seed_reads(reads, primer_set, threshold, aligner=None, device=0, engine=None, allowed_overlap=None, devices=None,
               compact=True, method=None, args=None, subreads=False, hit_buffer=None):
    """findAllPrimerAlignments for every read of a batch in one device pass.  allowed_overlap (LAMPrey uses 4,
    lamprey.py:1536) also runs removeOverlappingPrimerAlignments on the device: see SeedBatch.kept.

    devices=[0, 1, ...]: the batch is split over those GPUs in contiguous read ranges balanced by bases, one host thread
    and context per GPU, and the hits are gathered on the host in global read order -- what the reference does with
    worker processes (lamprey.py:2182-2207).  The result is identical to the single-device one.

    method="skbio" (or an `args` whose .swalign is False, the reference's default) seeds with scikit-bio's
    StripedSmithWaterman semantics instead of swalign's (lamprey.py:572-593): hits then carry identity = matches / len(primer).

    With `devices`, the PCIe copies run at full speed when the read arrays and `hit_buffer` (records of _native.HIT16_DT with
    compact=True, the default, else _native.HIT_DT) lie in page-locked memory: fastq.read_fastq(path, pinned=True) and
    _native.PinnedArray(n, dtype).array provide both."""
    from . import swalign as _sw
    concat, offs = _read_arrays(reads)
    pairs = primer_set.sequences()
    names = [n for n, _ in pairs]
    skbio = _use_skbio(args, method)
    if skbio and devices is not None:
        if subreads:
            raise ValueError("subreads=True needs a single device")
        pool = get_pool(devices)
        pool.configure([s for _, s in pairs], threshold, allowed_overlap=allowed_overlap, compact=False, ssw=SKBIO_DEFAULTS)
        hits, counters = _seed_on_pool(pool, concat, offs, hit_buffer=hit_buffer)
        return SeedBatch(hits, len(offs) - 1, np.diff(offs), names, counters, qlen=np.array([len(s) for _, s in pairs], np.int64))
    if skbio:
        eng = engine or _sw.get_engine(device)
        eng.set_ssw(*SKBIO_DEFAULTS)
        eng.set_queries([s for _, s in pairs])
        eng.upload_reads(concat, offs)
        eng.set_overlap(allowed_overlap)
        try:
            hits = eng.find_all(threshold)
            sub = _device_subreads(eng, primer_set, names) if subreads else None
        finally:
            eng.set_overlap(None)
        batch = SeedBatch(hits, len(offs) - 1, np.diff(offs), names, eng.last_counters,
                          qlen=np.array([len(s) for _, s in pairs], np.int64))
        if sub is not None:
            batch._set_subreads(sub)
        return batch
    if subreads and (allowed_overlap is None or devices is not None):
        raise ValueError("subreads=True needs allowed_overlap (LAMPrey uses 4) and a single device")
    if devices is not None:
        pool = get_pool(devices)
        sc = _scoring_of(aligner)
        pool.configure([s for _, s in pairs], threshold, *sc, allowed_overlap=allowed_overlap, compact=compact)
        hits, counters = _seed_on_pool(pool, concat, offs, hit_buffer=hit_buffer)
        return SeedBatch(hits, len(offs) - 1, np.diff(offs), names, counters)
    eng = engine or _sw.get_engine(device)
    if aligner is not None:
        eng.set_scoring(aligner.scoring_matrix.match, aligner.scoring_matrix.mismatch, aligner.gap_penalty,
                        aligner.gap_extension_penalty)
    else:
        eng.set_scoring(2, -1, -1, -1)                  # initAligners, lamprey.py:1860-1865
    eng.set_queries([s for _, s in pairs])
    eng.upload_reads(concat, offs)
    eng.set_overlap(allowed_overlap)
    try:
        hits = eng.find_all(threshold)
        sub = _device_subreads(eng, primer_set, names) if subreads else None
    finally:
        eng.set_overlap(None)            # the engine is shared (swalign.get_engine): never leave the pruning switched on
    batch = SeedBatch(hits, len(offs) - 1, np.diff(offs), names, eng.last_counters)
    if sub is not None:
        batch._set_subreads(sub)
    return batch


def seed_stream(batches, primer_set, threshold, aligner=None, devices=(0,), allowed_overlap=None, compact=True, depth=2):
    """Seed a stream of read batches (e.g. the FASTQ files an online run delivers one by one, lamprey.py:451-507), yielding
    one SeedBatch per input batch, in order.  The library double-buffers: while one batch is being seeded, the next one is
    uploaded and the previous one's hits are downloaded (`depth` contexts per device take alternate batches)."""
    pairs = primer_set.sequences()
    names = [n for n, _ in pairs]
    pool = get_pool(devices, depth)
    pool.configure([s for _, s in pairs], threshold, *_scoring_of(aligner), allowed_overlap=allowed_overlap, compact=compact)
    inflight = []

    def finish(item):
        ticket, concat, offs, out = item
        try:
            hits, counters, times = pool.wait(ticket)
        except OverflowError as e:                       # rare: the default hit buffer was too small for this batch
            hits, counters = _seed_on_pool(pool, concat, offs, hit_cap=e.needed + 1024)
            return SeedBatch(hits, len(offs) - 1, np.diff(offs), names, counters)
        if pool.hit_format == _native.HITS_COMPACT:
            hits = _native.expand_hits(hits, *pool.scoring)
        counters["pool_times"] = times
        return SeedBatch(hits, len(offs) - 1, np.diff(offs), names, counters)

    for reads in batches:
        concat, offs = _read_arrays(reads)
        out = np.zeros(len(concat) // 8 + 4096, pool.hit_dtype)
        inflight.append((pool.submit(concat, offs, out), concat, offs, out))
        if len(inflight) >= depth:
            yield finish(inflight.pop(0))
    while inflight:
        yield finish(inflight.pop(0))


def _unpin(got):
    """(PinnedArray, PinnedArray, ...) of fastq.read_fastq(pinned=True) -> the arrays, plus the objects that keep them alive."""
    return (got[0].array, got[1].array), (got[0], got[1])


def seed_fastq(path, primer_set, threshold, aligner=None, device=0, engine=None, allowed_overlap=None, with_names=False,
               devices=None, method=None, args=None, pinned=False, hit_buffer=None):
    """Seed one FASTQ file as one batch -- what the online FastqHandler.process_file does per file
    (/root/reference/lamprey.py:451-507) and the offline loader per run (lamprey.py:2145-2155), without Biopython: the
    library's scanner cuts the file's bytes into the (text, offsets) pair the upload takes (fastq.read_fastq).
    pinned=True parses straight into page-locked memory.  -> SeedBatch [, read names]."""
    from . import fastq as _fastq
    got = _fastq.read_fastq(path, with_names=with_names, pinned=pinned)
    arrays, keep = _unpin(got) if pinned else ((got[0], got[1]), None)
    batch = seed_reads(arrays, primer_set, threshold, aligner=aligner, device=device, engine=engine,
                       allowed_overlap=allowed_overlap, devices=devices, method=method, args=args, hit_buffer=hit_buffer)
    if keep is not None:
        batch.read_lengths = np.array(batch.read_lengths)          # (np.diff made a copy already; nothing else refers to the pinned arrays)
        for k in keep:
            k.close()
    return (batch, got[2]) if with_names else batch


def seed_fastq_files(paths, primer_set, threshold, aligner=None, devices=(0,), allowed_overlap=None, compact=True, depth=2,
                     pinned=False):
    """A run of FASTQ files, one batch per file, in order -- the online mode's stream of delivered files (lamprey.py:451-507):
    file k + 1 is parsed (all host cores) and uploaded while file k is being seeded.  Yields (path, SeedBatch)."""
    from . import fastq as _fastq
    paths = list(paths)
    keep = []

    def batches():
        for p in paths:
            got = _fastq.read_fastq(p, pinned=pinned)
            if pinned:
                arrays, objs = _unpin(got)
                keep.append(objs)
                yield arrays
            else:
                yield got
    for k, batch in enumerate(seed_stream(batches(), primer_set, threshold, aligner=aligner, devices=devices,
                                          allowed_overlap=allowed_overlap, compact=compact, depth=depth)):
        if pinned:                                       # batch k is done: nothing reads its page-locked text any more
            for o in keep[k]:
                o.close()
            keep[k] = None
        yield paths[k], batch


def alignSequence_swalign(aligner, seq, primer, primer_name):
    """Aligns a primer to a DNA sequence -- drop-in for /root/reference/lamprey.py:558-569: one
    ``aligner.align(str(seq), primer)`` and the three fields LAMPrey keeps of it."""
    alignment_tmp = aligner.align(str(seq), primer)
    return Alignment(alignment_tmp.r_pos, alignment_tmp.r_end, alignment_tmp.identity, primer_name)


def alignSequence_skbio(seq, primer, primer_name):
    """Aligns two sequences with scikit-bio's StripedSmithWaterman semantics -- drop-in for
    /root/reference/lamprey.py:572-593 (the default seeding branch)."""
    from . import skbio_ssw
    alignment = skbio_ssw.StripedSmithWaterman(primer)(str(seq))
    aq, at = alignment.aligned_query_sequence, alignment.aligned_target_sequence
    matches = 0
    for i in range(0, len(aq)):
        if aq[i] == at[i]:
            matches = matches + 1
    identity = float(matches) / float(len(primer))
    return Alignment(alignment.target_begin, alignment.target_end_optimal + 1, identity, primer_name)


def _preorder(hits, identity, name):
    """Rebuild findPrimerAlignments' pre-order list from one sequence's hits sorted by start.
    Hits of one sequence never overlap and each knows its recursion depth, so the recursion tree
    is the Cartesian tree of (start-order, depth): root = the depth-minimum of a segment."""
    out = []
    stack = [(0, len(hits))]
    while stack:
        lo, hi = stack.pop()
        if lo >= hi:
            continue
        k = lo + int(np.argmin(hits["depth"][lo:hi]))
        out.append(Alignment(int(hits["start"][k]), int(hits["end"][k]), float(identity[k]), name))
        stack.append((k + 1, hi))      # right is visited after left: push it first
        stack.append((lo, k))
    return out


def findPrimerAlignments(aligner, seq, primer, primer_name, identity_threshold, args=None):
    """Greedy best-hit + left/right recursion for ONE primer on ONE sequence; returns the same
    pre-order list the reference builds.  ``args`` is accepted for signature compatibility
    (the reference only reads ``args.swalign`` from it)."""
    from . import swalign as _sw
    eng = _sw.get_engine(getattr(aligner, "device", 0))
    skbio = _use_skbio(args, None)
    _set_engine_scoring(eng, aligner, skbio)
    eng.set_queries([primer])
    eng.upload_read_list([str(seq)])
    hits = eng.find_all(identity_threshold)
    if skbio:
        return _preorder(hits, hits["matches"] / float(len(primer)), primer_name)
    tot = hits["matches"].astype(np.int64) + hits["mismatches"]
    return _preorder(hits, hits["matches"] / np.maximum(tot, 1), primer_name)


def findAllPrimerAlignments(aligner, seq, primers, identity_threshold, args=None):
    """(stable-sorted hit list, coverage) for one read -- drop-in for lamprey.py:699-733."""
    batch = seed_reads([str(seq)], primers, identity_threshold, aligner=aligner,
                       device=getattr(aligner, "device", 0), args=args)
    return batch[0]


def shard_reads(offsets, world_size):
    """Contiguous read ranges per rank, balanced by bases (the reference
    deals reads round-robin over processes, lamprey.py:2182-2187).  -> [(lo, hi)] * world_size."""
    offsets = np.asarray(offsets, np.int64)
    n = len(offsets) - 1
    total = offsets[-1] - offsets[0]
    cuts = [0]
    for r in range(1, world_size):
        target = offsets[0] + (total * r) // world_size
        cuts.append(int(min(n, max(cuts[-1], np.searchsorted(offsets, target, side="left")))))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]
