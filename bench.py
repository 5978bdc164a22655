#!/usr/bin/env python
"""bench.py -- seeding throughput of the B200 engine on BASELINE.json's synthetic workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--workload NAME] [--thr T]
    python bench.py --impl reference ...        # the reference arm: CPU (Cython) swalign seeding

One "step" = one pass of the hot path (findAllPrimerAlignments for every read: all 46 schema
sequences, greedy recursion included) over one batch of synthetic reads.  At N=1 the workload is
BASELINE.json configs[2]: 10^6 ONT-like concatemer reads, 2 kb mean, 8 % error, H3.3 schema.
With N GPUs every rank seeds its own batch of R reads (weak scaling, no data-path collective --
reads are independent units; only timings and counts are reduced).

Printed JSON line (rank 0): metric reads/s (whole job), plus GCUPS, roofline (integer ALU bound,
SURVEY.md section 8d), cpu_baseline (Cython swalign on the host cores, N=1 only), e2e (host
buffers -> hits on the host through the C ABI, H2D/D2H inside the timed region), clocks.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCHEMA = os.path.join(ROOT, "tests", "golden", "H3.3_set1.primers")
ALG_OPS_PER_CELL = 6          # int32-lane ops per DP cell (algorithmic)


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def gather_max_and_sum(t_local, sums, device="cuda"):
    """max over ranks of a time, sum over ranks of counters (the only cross-rank traffic)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return (t_local,) + tuple(sums)
    t = torch.tensor([t_local], dtype=torch.float64, device=device)
    s = torch.tensor(list(sums), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return (float(t[0]),) + tuple(int(round(v)) for v in s.tolist())


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle's Cython build ("Cython swalign", README.md:18-25 of the reference) driven by
# the restated findAllPrimerAlignments, one process per host core (lamprey.py:2182-2196).
# ---------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(schema_path, n_names=None, pure_python=False):
    import importlib.util
    mods = {}
    if pure_python:      # the same two files, interpreted (SURVEY.md section 8d asks for this figure as well)
        from oracle import swalign_ref, seeding_ref
        mods = {"swalign_cy": swalign_ref, "seeding_cy": seeding_ref}
    else:
        from oracle import build_cython
        out = build_cython.build()
        for name in ("swalign_cy", "seeding_cy"):
            so = [f for f in os.listdir(out) if f.startswith(name) and f.endswith(".so")][0]
            spec = importlib.util.spec_from_file_location(name, os.path.join(out, so))
            mods[name] = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mods[name])
    _CPU["sw"] = mods["swalign_cy"].LocalAlignment(mods["swalign_cy"].NucleotideScoringMatrix(2, -1))
    _CPU["seed"] = mods["seeding_cy"]
    _CPU["ps"] = mods["seeding_cy"].PrimerSet(schema_path)
    if n_names:          # workloads that search only the first n named sequences of the schema
        _CPU["ps"].primer_dict = dict(list(_CPU["ps"].primer_dict.items())[:n_names])


def _cpu_seed(args):
    read, thr = args
    st = _CPU["seed"].Stats()
    hits, _ = _CPU["seed"].find_all_primer_alignments(_CPU["sw"], read, _CPU["ps"], thr, st)
    return len(hits), st.calls, st.cells


def cpu_seed_sample(reads, thr, cores, pool=None, n_names=None, pure_python=False):
    """-> (seconds, hits, calls, cells) for seeding `reads` on `cores` processes."""
    import multiprocessing as mp
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init, initargs=(SCHEMA, n_names, pure_python))
        pool.map(_cpu_seed, [("ACGT" * 8, thr)] * cores)        # start-up (imports) outside the timing
    t0 = time.perf_counter()
    res = pool.map(_cpu_seed, [(r, thr) for r in reads], chunksize=1)
    dt = time.perf_counter() - t0
    if own:
        pool.close()
        pool.join()
    return dt, sum(r[0] for r in res), sum(r[1] for r in res), sum(r[2] for r in res)


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([f.strip() for f in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_JSON_FD = None


def emit(line):
    """The JSON line goes to the real stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def ncu_traffic():
    """DRAM bytes of the dominant K1 launch from the committed `ncu --set full` capture (profiles/, a 32768-read run:
    per-launch sizes differ from the bench's, so it is reported beside `traffic`, not as it).  HBM is not the bound here."""
    import csv
    path = os.path.join(ROOT, "profiles", "r01_score_raw.csv")
    try:
        rows = list(csv.reader(open(path)))
        h = rows[0]
        r = max(rows[2:], key=lambda x: float(x[h.index("gpu__time_duration.sum")]))
        rd, wr = float(r[h.index("dram__bytes_read.sum")]), float(r[h.index("dram__bytes_write.sum")])
        unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[rows[1][h.index("dram__bytes_read.sum")]]
        ms = float(r[h.index("gpu__time_duration.sum")])
        return {"kernel": r[h.index("Kernel Name")].replace("void ", ""), "workload_reads": 32768,
                "dram_bytes_per_launch": (rd + wr) * unit, "launch_ms_under_ncu": ms,
                "hbm_gb_per_s": (rd + wr) * unit / 1e9 / (ms / 1e3), "source": "profiles/r01_score_raw.csv"}
    except Exception:
        return None


def workload_reads(args, rank, world, workers):
    from lamprey_b200 import synth
    kw = dict(synth.CONFIGS[args.workload])
    kw["seed"] = kw["seed"] + 1000 * rank            # every rank seeds its own batch (weak scaling)
    return synth.generate(args.reads, workers=workers, **kw)


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    from lamprey_b200 import synth
    cores = os.cpu_count() or 1
    sample = 3 * max(cores, 8)        # a few reads per core, so one long read does not set the step time
    kw = dict(synth.CONFIGS[args.workload])
    concat, offs = synth.generate(sample * (args.steps + args.warmup), chunk=sample, **kw)
    reads = synth.reads_as_strings(concat, offs)
    import multiprocessing as mp
    pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init, initargs=(SCHEMA, synth.SCHEMA_NAMES.get(args.workload)))
    pool.map(_cpu_seed, [("ACGT" * 8, args.thr)] * cores)
    k = 0
    for _ in range(args.warmup):
        cpu_seed_sample(reads[k:k + sample], args.thr, cores, pool)
        k += sample
    t = cells = 0
    for _ in range(args.steps):
        dt, _, _, c = cpu_seed_sample(reads[k:k + sample], args.thr, cores, pool)
        t += dt
        cells += c
        k += sample
    pool.close()
    pool.join()
    value = sample * args.steps / t
    line = {"impl": "reference", "metric": "reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "python-int",
            "data": "synthetic", "gcups": cells / t / 1e9,
            "config": {"workload": args.workload, "schema": "H3.3_set1 (46 sequences)", "threshold": args.thr,
                       "reads_per_step": sample},
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port",
                             "sample": "%d reads of %s per step, Cython build of oracle/swalign_ref.py + "
                                       "seeding_ref.py, one process per core" % (sample, args.workload)},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1000000, help="reads per GPU per step")
    ap.add_argument("--workload", default="synth_2kb_8pct")
    ap.add_argument("--thr", type=float, default=0.70)          # lamprey.py:2293 default
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-engines", type=int, default=2, help="contexts (CUDA streams) that take alternate e2e steps")
    args = ap.parse_args()
    # stdout carries the one JSON line and nothing else: libraries that print to fd 1 (NCCL's version
    # banner, forked workers) are pointed at stderr for the whole run.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    rank, local_rank, world = dist_env()
    cores = os.cpu_count() or 1
    workers = max(1, cores // max(1, world))

    # ---- everything that forks happens before CUDA is touched -------------------------------
    concat, offs = workload_reads(args, rank, world, workers)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from lamprey_b200 import synth
        # ~2 s per 2 kb read per core for the Cython port: size the sample for about --cpu-seconds
        per_read = 2.2 * (np.diff(offs).mean() / 2000.0) ** 1.0
        sample = int(max(cores, min(len(offs) - 1, cores * args.cpu_seconds / per_read)))
        reads = synth.reads_as_strings(concat[:offs[sample]], offs[:sample + 1])
        dt, hits_cpu, calls_cpu, cells_cpu = cpu_seed_sample(reads, args.thr, cores, n_names=synth.SCHEMA_NAMES.get(args.workload))
        cpu_baseline = {"value": sample / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                        "gcups": cells_cpu / dt / 1e9, "seconds": dt,
                        "sample": "first %d reads of the workload, Cython build of oracle/swalign_ref.py + "
                                  "seeding_ref.py (README.md:18-25 recipe), one process per core" % sample,
                        "_hits": hits_cpu, "_calls": calls_cpu, "_cells": cells_cpu, "_n": sample}
        # the interpreted figure beside it: one read per core
        dt_py, _, _, cells_py = cpu_seed_sample(reads[:cores], args.thr, cores, n_names=synth.SCHEMA_NAMES.get(args.workload),
                                                pure_python=True)
        cpu_baseline["pure_python"] = {"value": min(cores, len(reads)) / dt_py, "unit": "reads/s", "gcups": cells_py / dt_py / 1e9,
                                       "sample": "first %d reads, oracle/swalign_ref.py + seeding_ref.py interpreted" % min(cores, len(reads))}

    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from lamprey_b200 import _native, seeding

    eng = _native.Engine(local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_scoring(2, -1, -1, -1)
    ps = seeding.PrimerSet(SCHEMA)
    from lamprey_b200 import synth as _synth
    seqs = [s for _, s in _synth.schema_sequences(ps, args.workload)]
    eng.set_queries(seqs)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # pinned host buffers: the reference-facing call takes host memory
    h_ascii = torch.empty(len(concat), dtype=torch.uint8, pin_memory=True)
    h_ascii.numpy()[:] = concat
    h_offs = torch.empty(len(offs), dtype=torch.int64, pin_memory=True)
    h_offs.numpy()[:] = offs
    ascii_np, offs_np = h_ascii.numpy(), h_offs.numpy()

    # integer roofline of this box (not in MEASURED_PEAKS.json)
    peak = {}
    if rank == 0:
        for mode, name in ((0, "viaddmnmx_s16x2"), (1, "vimnmx3_s16x2"), (2, "iadd"), (3, "imad"), (4, "k1_mix"),
                           (5, "k1_mix_imad"), (6, "lop3")):
            ops, _ = eng.measure_int_peak(mode, 20000)
            peak[name] = ops / 1e9

    # ---- value: inputs resident in HBM ---------------------------------------------------------
    eng.upload_reads(ascii_np, offs_np)
    n_hits = 0
    for _ in range(args.warmup):
        n_hits = eng.find_all(args.thr, fetch=False)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = dict(ms_score=0.0, ms_trace=0.0, ms_other=0.0, ms_total=0.0, launches=0, score_launches=0)
    e0.record(stream)
    for _ in range(args.steps):
        n_hits = eng.find_all(args.thr, fetch=False)
        c = eng.last_counters
        for k in ("ms_score", "ms_trace", "ms_other", "ms_total"):
            acc[k] += c[k]
        acc["launches"] += c["score_launches"] + c["trace_launches"] + c["other_launches"]
        acc["score_launches"] += c["score_launches"]
    e1.record(stream)
    barrier()
    t_value = e0.elapsed_time(e1) / 1e3
    clocks = sampler.stop() if sampler else None
    cnt = dict(eng.last_counters)
    t_value, reads_total, cells_total, hits_total = gather_max_and_sum(
        t_value, [args.reads * args.steps, cnt["cells"] * args.steps, n_hits * args.steps])

    # parity spot check against the CPU baseline's own sample (same reads, same threshold)
    parity = None
    if cpu_baseline is not None:
        n = cpu_baseline.pop("_n")
        eng.upload_reads(ascii_np[:offs_np[n]], offs_np[:n + 1])
        sub = eng.find_all(args.thr)
        c2 = eng.last_counters
        parity = {"reads": n, "hits_match": int(len(sub)) == cpu_baseline.pop("_hits"),
                  "calls_match": c2["tasks"] == cpu_baseline.pop("_calls"),
                  "cells_match": c2["cells"] == cpu_baseline.pop("_cells")}
        eng.upload_reads(ascii_np, offs_np)

    # ---- e2e: host buffers in, hits on the host out, every step --------------------------------
    # The reference-facing call sequence (lsw_upload_reads -> lsw_find_all -> lsw_fetch_hits) from pinned host
    # memory.  With --e2e-engines 2 (default) two contexts, each on its own CUDA stream and driven by its own
    # host thread, take alternate steps, so one batch's PCIe copies overlap the other batch's kernels -- the
    # way a caller streaming FASTQ batches would use the library.  Every step still uploads its own input
    # and downloads its own hits inside the timed region (transfers are serialised per direction).
    e2e = None
    if not args.no_e2e:
        import threading
        n_eng = max(1, args.e2e_engines)
        engines = [eng]
        for _ in range(n_eng - 1):
            e2 = _native.Engine(local_rank)
            e2.set_scoring(2, -1, -1, -1)
            e2.set_queries(seqs)
            engines.append(e2)
        if n_eng > 1:
            eng.set_stream(0)                                        # back to the context's own non-blocking stream
        bufs = []
        for _ in engines:
            hb = torch.empty(max(1, int(n_hits * 1.05) + 1024) * 32, dtype=torch.uint8, pin_memory=True)
            bufs.append(hb.numpy().view(_native.HIT_DT))
        got = [0] * n_eng
        h2d_lock, d2h_lock = threading.Lock(), threading.Lock()       # one transfer per direction at a time
        # one find_all at a time: two contexts computing side by side share SMs and L2 and each runs at less than half
        # speed, while the copies of the waiting context overlap the running one's kernels either way
        gpu_lock = threading.Lock() if os.environ.get("LSW_E2E_COMPUTE_LOCK", "1") != "0" else contextlib.nullcontext()

        phase = {"upload": [], "find_all": [], "fetch": []}          # host wall clock of every call (s), all contexts

        def e2e_worker(k, my_steps):
            for _ in range(my_steps):
                with h2d_lock:
                    t0 = time.perf_counter()
                    engines[k].upload_reads(ascii_np, offs_np)       # H2D of this step's reads (+ re-coding)
                    t1 = time.perf_counter()
                with gpu_lock:
                    t1b = time.perf_counter()
                    engines[k].find_all(args.thr, fetch=False)       # overlaps the other context's copies
                    t2 = time.perf_counter()
                with d2h_lock:
                    t3 = time.perf_counter()
                    got[k] = len(engines[k].fetch_hits(out=bufs[k]))  # D2H of this step's hits
                    t4 = time.perf_counter()
                phase["upload"].append(t1 - t0); phase["find_all"].append(t2 - t1b); phase["fetch"].append(t4 - t3)

        def run_steps(total):
            per = [total // n_eng + (1 if k < total % n_eng else 0) for k in range(n_eng)]
            ths = [threading.Thread(target=e2e_worker, args=(k, per[k])) for k in range(n_eng) if per[k]]
            for t in ths:
                t.start()
            for t in ths:
                t.join()

        run_steps(max(n_eng, args.warmup - 2))
        for v in phase.values():
            del v[:]
        barrier()
        e0.record(stream)
        run_steps(args.steps)
        torch.cuda.synchronize()
        e1.record(stream)
        barrier()
        t_e2e = e0.elapsed_time(e1) / 1e3
        t_e2e, reads_e2e = gather_max_and_sum(t_e2e, [args.reads * args.steps])
        e2e = {"value": reads_e2e / t_e2e, "unit": "reads/s",
               "h2d_bytes_per_step": int(ascii_np.nbytes + offs_np.nbytes), "d2h_bytes_per_step": int(max(got) * 32),
               "ms_per_step": 1e3 * t_e2e / args.steps, "engines": n_eng, "hits_per_step": int(max(got)),
               "host_ms_per_call": {k: 1e3 * float(np.mean(v)) for k, v in phase.items() if v},
               "call": "per step: lsw_upload_reads + lsw_find_all + lsw_fetch_hits from/to pinned host buffers; %d contexts "
                       "(CUDA streams, one host thread each) take alternate steps so one step's copies overlap another's kernels; "
                       "one lsw_find_all runs at a time" % n_eng}
        for e2 in engines[1:]:
            e2.close()

    if rank == 0:
        gcups = cells_total / t_value / 1e9
        score_s = acc["ms_score"] / 1e3
        # cells K1 actually filled (child tasks reuse the round-0 blocks of their read, so this is below the
        # reference-equivalent count that `gcups` uses)
        k1_gcups = cnt["computed_cells"] * args.steps / score_s / 1e9 if score_s > 0 else None
        peak_ops = peak.get("viaddmnmx_s16x2", 0.0)      # G lane-ops / s of the packed DPX instruction K1 is built from
        # one packed (s16x2) instruction lane updates 2 cells' worth of one algorithmic op
        peak_gcups = peak_ops * 2.0 / ALG_OPS_PER_CELL if peak_ops else None
        roofline = {"bound": "int_alu", "kernel": "k_score<NR> (K1: DP fill + arg-max)",
                    "achieved": k1_gcups, "peak": peak_gcups, "unit": "GCUPS",
                    "frac": (k1_gcups / peak_gcups) if (k1_gcups and peak_gcups) else None,
                    "traffic": None, "traffic_ncu": ncu_traffic(),
                    # K1's real instruction mix: 5 packed DPX lane-ops per 4 cells (SASS-checked) -> the rate at which the
                    # DPX pipe alone would saturate, and the rate a register-only microbenchmark of the row mix reaches
                    "peak_dpx_only": peak_ops * 4.0 / 5.0 if peak_ops else None,
                    "frac_dpx_only": (k1_gcups / (peak_ops * 4.0 / 5.0)) if (k1_gcups and peak_ops) else None,
                    # (9 instructions per 4 cells at the lane-op rate the mixed DPX/add chain measures)
                    "peak_row_mix_microbench": peak.get("k1_mix", 0.0) * 4.0 / 9.0 if peak.get("k1_mix") else None,
                    "frac_row_mix_microbench": (k1_gcups / (peak["k1_mix"] * 4.0 / 9.0)) if (k1_gcups and peak.get("k1_mix")) else None,
                    "peak_source": "lsw_measure_int_peak on this box (MEASURED_PEAKS.json has no integer rate): "
                                   "VIADDMNMX.S16x2 lane-ops/s x 2 cells / %d algorithmic ops per cell" % ALG_OPS_PER_CELL,
                    "int_peak_glaneops": peak, "launches": acc["score_launches"],
                    "avg_launch_ms": acc["ms_score"] / max(1, acc["score_launches"]),
                    "k1_share_of_step": acc["ms_score"] / max(1e-9, acc["ms_total"]),
                    "computed_cells_per_step": cnt["computed_cells"], "reference_cells_per_step": cnt["cells"],
                    "k1_reference_equivalent_gcups": cnt["cells"] * args.steps / score_s / 1e9 if score_s > 0 else None,
                    "cells_per_lane_step": cnt["computed_cells"] / max(1, cnt["steps"])}
        line = {"metric": "reads_per_sec", "value": reads_total / t_value, "unit": "reads/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_value / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s16x2",
                "data": "synthetic", "gcups": gcups, "gcups_computed": cnt["computed_cells"] * args.steps * world / t_value / 1e9,
                "gcups_top_level": cnt["top_cells"] * args.steps * world / t_value / 1e9,
                "config": {"workload": args.workload, "reads_per_gpu": args.reads, "mean_len": float(np.diff(offs).mean()),
                           "schema": "H3.3_set1 (%d sequences, %d rows)" % (len(seqs), sum(len(q) for q in seqs)), "threshold": args.thr,
                           "l2": "inputs larger than L2 (%.2f GB of read codes + task lists per step)" % (len(concat) / 1e9),
                           "tasks_per_step": cnt["tasks"], "cells_per_step": cnt["cells"], "hits_per_step": n_hits,
                           "rounds": cnt["rounds"]},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
                "gpu_launches": acc["launches"],
                "breakdown_ms_per_step": {k: acc[k] / args.steps for k in ("ms_score", "ms_trace", "ms_other", "ms_total")},
                "parity_vs_cpu_sample": parity}
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
