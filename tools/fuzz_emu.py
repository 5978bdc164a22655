"""CPU fuzzer: the host emulation of the kernels' lane logic (liblsw_emu.so: K1, K2 band / fast / full, reuse of round 0, K1-affine,
the SSW per-task kernel) against the C oracles, on inputs built to hit ties and odd shapes: two-letter alphabets, homopolymers,
tandem repeats of a sequence (LAMP concatemers), planted noisy copies, non-ACGT bytes, reads of 0..3000 bases, 1..63-base sequences,
several scorings and thresholds.   python tools/fuzz_emu.py [seconds [first seed]]     (run from the repo root; no GPU needed)"""
import random
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import emu_lib                            # noqa: E402
from conftest import hits_matrix          # noqa: E402
from lamprey_b200._native import TASK_DT  # noqa: E402
from oracle import c_oracle, ssw_oracle   # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def noisy(rng, s, rate):
    out = []
    for ch in s:
        u = rng.random()
        if u < rate / 3:
            continue
        out.append(rng.choice("ACGT") if u < 2 * rate / 3 else ch)
        if u > 1 - rate / 3:
            out.append(rng.choice("ACGT"))
    return "".join(out)


def make_case(rng):
    alphabet = rng.choice(["ACGT", "ACGT", "AC", "AT", "A", "ACG"])
    M, X, gap = rng.choice([(2, -1, -1), (2, -1, -1), (2, -1, -1), (1, -1, -1), (3, -2, -2), (2, -3, -1), (1, -2, -3), (2, -1, -2), (4, -3, -2)])
    qmax = min(63, (127 + gap) // M)
    seqs = ["".join(rng.choice(alphabet) for _ in range(rng.choice([1, 2, 5, 15, 16, 17, 20, 24, 25, 31, 32, 33, 40, 48, 57, 63]) % qmax + 1))
            for _ in range(rng.randrange(1, 7))]
    if rng.random() < 0.5:                   # a sequence and its reverse complement, as the schema has them
        seqs.append("".join(COMP[c] for c in reversed(seqs[0])))
    thr = rng.choice([0.5, 0.6, 0.7, 0.7, 0.75, 0.8, 0.9, 0.95])
    reads = []
    long_case = rng.random() < 0.08          # a few long reads: slices that span whole groups of blocks (second-level block bests)
    for _ in range(rng.randrange(1, 14) if not long_case else rng.randrange(1, 3)):
        n = rng.choice([0, 1, 5, 31, 32, 33, 63, 64, 65, 127, 128, 129, 200, 449, 700, 1024, 1500, 3000])
        if long_case:
            n = rng.choice([5000, 9000, 20000, 40000])
        kind = rng.random()
        if long_case and kind >= 0.6:
            kind = rng.random() * 0.6        # (random reads hold nothing to find)
        if kind < 0.4:                       # concatemer of noisy copies of the sequences
            r, rate = "", rng.choice([0.0, 0.03, 0.08, 0.12, 0.2])
            while len(r) < n:
                r += noisy(rng, rng.choice(seqs), rate)
                if rng.random() < 0.3:
                    r += "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 30)))
            r = r[:n]
        elif kind < 0.6:                     # exact tandem repeats (every copy ties with its neighbours)
            s = rng.choice(seqs)
            r = (s * (n // len(s) + 1))[:n]
        else:
            r = [rng.choice(alphabet) for _ in range(n)]
            for _ in range(rng.randrange(0, 8)):
                s = rng.choice(seqs)
                if n > len(s):
                    p = rng.randrange(0, n - len(s))
                    for k, ch in enumerate(s):
                        if rng.random() > 0.1:
                            r[p + k] = ch
            r = "".join(r)
        if n and rng.random() < 0.3:
            r = list(r)
            for _ in range(rng.randrange(1, 4)):
                r[rng.randrange(n)] = rng.choice("NnacgtRY-")
            r = "".join(r)
        reads.append(r)
    return (M, X, gap), seqs, thr, reads


def oracle_rows(find_all, reads, seqs, thr, *scoring):
    rows, calls, cells = [], 0, 0
    for i, r in enumerate(reads):
        h, st = find_all(r, seqs, thr, *scoring)
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]; cells += st["cells"]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"], x["score"], x["q_pos"], x["q_end"], x["depth"]) for x in h]
    return np.array(rows, np.int32).reshape(-1, 10), calls, cells


t_end = time.time() + budget
it = bad = too_deep = limited = 0
n_hits = 0
while time.time() < t_end:
    rng = random.Random(seed0 + it)
    (M, X, gap), seqs, thr, reads = make_case(rng)
    what = []
    # swalign path: find_all with and without the reuse of round 0, and with every candidate forced through the full-window kernel
    want, calls, cells = oracle_rows(c_oracle.find_all, reads, seqs, thr, M, X, gap, gap)
    if len(want) and want[:, 9].max() > 30000:
        # deeper than the engine goes (lsw_hit.depth is 16 bits; lsw_find_all stops at 30 000 rounds with an error):
        # a 1-base "primer" on a 40 kb homopolymer recurses once per base
        too_deep += 1
        it += 1
        continue
    n_hits += len(want)
    for kw in (dict(reuse=True), dict(reuse=False), dict(force_full=True)):
        hits, cnt = emu_lib.find_all(reads, seqs, thr, M, X, gap, **kw)
        if not (np.array_equal(hits_matrix(hits), want) and (cnt["tasks"], cnt["cells"]) == (calls, cells)):
            what.append("find_all %r" % kw)
    # align(): a few random slices, CIGARs included
    tasks, exp = [], []
    for _ in range(6):
        ri = rng.randrange(len(reads))
        a = rng.randrange(0, len(reads[ri]) + 1); b = rng.randrange(a, len(reads[ri]) + 1)
        q = rng.randrange(len(seqs))
        tasks.append((ri, a, b, q))
        exp.append(c_oracle.align(reads[ri][a:b].upper(), seqs[q], M, X, gap, gap))
    res, cig = emu_lib.align_tasks(reads, seqs, np.array(tasks, TASK_DT), M, X, gap)
    for k, d in enumerate(exp):
        r = res[k]
        runs = [(int(c >> 2), " MID"[int(c & 3)]) for c in cig[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]]]
        if tuple(int(r[f]) for f in ("score", "q_pos", "q_end", "r_pos", "r_end", "matches", "mismatches")) != \
                tuple(d[f] for f in ("score", "q_pos", "q_end", "r_pos", "r_end", "matches", "mismatches")) or runs != d["cigar"]:
            what.append("align task %r" % (tasks[k],))
    # skbio path (default scoring only fits the packed cells for match * len <= 127)
    if it % 2 == 0:
        rows, calls, cells = [], 0, 0
        for i, r in enumerate(reads):
            h, st = ssw_oracle.find_all(r.upper(), seqs, thr)      # the engine upper-cases reads (LAMPrey feeds upper-case FASTQ)
            h = h[np.argsort(h["start"], kind="stable")]
            calls += st["calls"]; cells += st["cells"]
            rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["score"], x["q_pos"], x["q_end"], x["depth"]) for x in h]
        want = np.array(rows, np.int32).reshape(-1, 9)
        for packed in (True, False):
            hits, cnt = emu_lib.ssw_find_all(reads, seqs, thr, packed=packed)
            got = hits_matrix(hits)[:, [0, 1, 2, 3, 4, 6, 7, 8, 9]]       # (the oracle's hit record has no mismatch count)
            if not (np.array_equal(got, want) and (cnt["tasks"], cnt["cells"]) == (calls, cells)):
                what.append("ssw_find_all packed=%r" % packed)
    # stage 2 (lamprey.py:1637-1658): random sub-reads of the batch (reverse-complemented or not) against one read as the reference
    refs = [r for r in reads if 50 <= len(r) <= 700]
    if it % 4 == 1 and refs:
        from lamprey_b200._native import PAIR_DT
        ref = refs[0].upper()
        pairs = []
        for _ in range(12):
            i = rng.randrange(len(reads))
            if reads[i]:
                a = rng.randrange(len(reads[i]))
                pairs.append((i, a, min(len(reads[i]), a + rng.randrange(1, 600)), rng.randrange(2)))
        if pairs:
            try:
                res, cig = emu_lib.ssw_subreads(reads, np.array(pairs, PAIR_DT), ref)
            except RuntimeError:                     # the band limit of the stage-2 kernel (reported, not guessed): counted
                res, cig, pairs = [], [], []
                limited += 1
            for (i, a, b, rcf), r in zip(pairs, res):
                q = reads[i][a:b].upper()
                if rcf:
                    q = "".join(COMP.get(ch, ch) for ch in reversed(q))
                d = ssw_oracle.align(ref, q)
                if d["score"] == 0:
                    ok = r["score"] == 0
                else:
                    runs = [(int(x >> 2), " MID"[int(x & 3)]) for x in cig[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]]]
                    # (matches: the engine never counts a non-ACGT character as a match, the restated wrapper compares raw characters --
                    # they differ only when the REFERENCE holds the same non-ACGT character; LAMPrey's references are ACGT and its
                    # stage 2 does not read the match count)
                    pure = set(ref) <= set("ACGT")
                    ok = (r["score"], r["r_pos"], r["r_end"] - 1, r["q_pos"], r["q_end"] - 1, r["matches"] if pure else 0, runs) == \
                        (d["score"], d["ref_begin"], d["ref_end"], d["read_begin"], d["read_end"], d["matches"] if pure else 0, d["cigar"])
                if not ok:
                    what.append("stage 2 pair %r" % ((i, a, b, rcf),))
    if what:
        bad += 1
        print("MISMATCH seed", seed0 + it, (M, X, gap), thr, [len(s) for s in seqs], [len(r) for r in reads], what, flush=True)
    it += 1
print("fuzz_emu: %d cases (seeds %d..%d), %d oracle hits, %d mismatching cases, %d skipped (recursion deeper than 30000), %d stage-2 batches beyond the band limit" % (it, seed0, seed0 + it - 1, n_hits, bad, too_deep, limited))
sys.exit(1 if bad else 0)
