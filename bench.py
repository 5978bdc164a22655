#!/usr/bin/env python
"""bench.py -- seeding throughput of the B200 engine on BASELINE.json's synthetic workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--workload NAME] [--thr T]
    python bench.py --impl reference ...        # the reference arm: CPU (Cython) swalign seeding

One "step" = one pass of the hot path (findAllPrimerAlignments for every read: all 46 schema
sequences, greedy recursion included) over one batch of synthetic reads.  At N=1 the workload is
BASELINE.json configs[2]: 10^6 ONT-like concatemer reads, 2 kb mean, 8 % error, H3.3 schema.
With N GPUs (one process per GPU) every rank seeds its own batch of R reads (weak scaling, no
data-path collective -- reads are independent units; only timings and counts are reduced).

Printed JSON line (rank 0): metric reads/s (whole job), plus
  gcups         reference cells per second (all recursion tasks);
  roofline      K1 against the DPX-pipe ceiling of its own instruction mix (integer ALU bound), the
                survey's 6-op model and the inner-loop microbenchmark beside it, ncu pipe counters and DRAM traffic of the
                10^6-read capture under profiles/;
  cpu_baseline  Cython swalign on the host cores (N=1), and a parity check of the GPU hits against it tuple by tuple;
  e2e           host buffers -> hits on the host through the library's own double-buffered pool (lsw_pool_*), a plain
                submit/wait loop, H2D and D2H inside the timed region;
  strong_scaling ONE fixed batch (rank 0's) split over all N GPUs by the pool of rank 0, host-side gather included;
  extra_configs BASELINE.json configs[3] and [4] (5 kb / 12 %; 20-50 kb, 40-sequence schema) with their own timing;
  host_side     FASTQ scanner / hit-record expansion / per-read index rates of the library's host threads on this box (no GPU work);
  clocks.
"""
import argparse
import csv
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
ALG_OPS_PER_CELL = 6          # int32-lane ops per DP cell (the survey's model)
K1_DPX_PER_4_CELLS = 5        # K1's real mix (lsw_score.cuh column2, SASS-checked): 5 packed DPX + 4 adds + 1 LDS.64 per 4 cells
NCU_SCORE_RAW = os.path.join(ROOT, "profiles", "r02_score24_1M_raw.csv")     # ncu --set full of round 0's k_score<24> at 10^6 reads


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
    """-> (hits as the reference returns them: [(start, end, identity, name)], align() calls, cells)."""
    read, thr = args
    st = _CPU["seed"].Stats()
    hits, _ = _CPU["seed"].find_all_primer_alignments(_CPU["sw"], read, _CPU["ps"], thr, st)
    return [(h.start, h.end, h.identity, h.primer_name) for h in hits], st.calls, st.cells


def cpu_seed_sample(reads, thr, cores, pool=None, n_names=None, pure_python=False):
    """-> (seconds, per-read hit lists, calls, cells) for seeding `reads` on `cores` processes."""
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
    return dt, [r[0] for r in res], sum(r[1] for r in res), sum(r[2] for r in res)


def _c_oracle_chunk(args):
    """Full hit records (all ten fields) of a slice of reads from the C restatement of the oracle."""
    from oracle import c_oracle
    reads, seqs, thr, first = args
    rows = []
    for i, r in enumerate(reads):
        h, _ = c_oracle.find_all(r, seqs, thr)
        h = h[np.argsort(h["start"], kind="stable")]
        rows += [(first + i, int(x["seq"]), int(x["start"]), int(x["end"]), int(x["matches"]), int(x["mismatches"]), int(x["score"]),
                  int(x["q_pos"]), int(x["q_end"]), int(x["depth"])) for x in h]
    return rows


def c_oracle_rows(reads, seqs, thr, cores):
    import multiprocessing as mp
    from oracle import c_oracle
    c_oracle.build()
    step = max(1, (len(reads) + cores - 1) // cores)
    with mp.get_context("fork").Pool(cores) as pool:
        parts = pool.map(_c_oracle_chunk, [(reads[k:k + step], seqs, thr, k) for k in range(0, len(reads), step)])
    return [r for p in parts for r in p]


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
_T0 = time.time()


def log(msg):
    """Progress on stderr (stdout carries the JSON line only)."""
    sys.stderr.write("[bench %6.1fs rank %s] %s\n" % (time.time() - _T0, os.environ.get("RANK", "0"), msg))
    sys.stderr.flush()


def emit(line):
    """The JSON line goes to the real stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def ncu_capture():
    """Counters of the round-0 k_score<24> launch AT THE 10^6-READ CONFIG from the committed `ncu --set full` capture
    (profiles/, produced by tools/profile_1M.sh): DRAM bytes of the launch, pipe utilisation, issue slots."""
    try:
        rows = list(csv.reader(open(NCU_SCORE_RAW)))
        h, units, r = rows[0], rows[1], rows[2]
        mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}

        def val(k):
            return float(r[h.index(k)].replace(",", ""))
        rd = val("dram__bytes_read.sum") * mult[units[h.index("dram__bytes_read.sum")]]
        wr = val("dram__bytes_write.sum") * mult[units[h.index("dram__bytes_write.sum")]]
        dur = val("gpu__time_duration.sum")
        du = units[h.index("gpu__time_duration.sum")]
        ms = dur / 1e6 if du in ("ns", "nsecond") else dur / 1e3 if du in ("us", "usecond") else dur if du in ("ms", "msecond") else dur * 1e3
        out = {"kernel": r[h.index("Kernel Name")].replace("void ", "") + " <24>, round 0 (32 sequences x 10^6 whole reads)",
               "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "launch_ms_under_ncu": ms,
               "hbm_gb_per_s": (rd + wr) / 1e9 / (ms / 1e3), "source": os.path.relpath(NCU_SCORE_RAW, ROOT)}
        for key, name in (("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
                          ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
                          ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                          ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct")):
            if key in h:
                out[name] = val(key)
        return out
    except Exception:
        return None


def workload_reads(workload, n_reads, rank, workers, world=1):
    from lamprey_b200 import synth
    kw = dict(synth.CONFIGS[workload])
    kw["seed"] = kw["seed"] + 1000 * rank            # every rank seeds its own batch (weak scaling)
    # all ranks of a box generate at the same time: each keeps its jobs in flight within its share of half the free host memory
    return synth.generate(n_reads, workers=workers, mem_share=0.5 / max(1, world), **kw)


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    from lamprey_b200 import synth
    cores = os.cpu_count() or 1
    sample = 3 * max(cores, 8)        # a few reads per core, so one long read does not set the step time
    kw = dict(synth.CONFIGS[args.workload])
    kw.pop("chunk", None)
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


# ---------------------------------------------------------------------------------------------
def timed_find_all(eng, torch, stream, thr, steps, warmup, barrier, sampler_gpu=None):
    """W warm-up + K timed passes over the batch resident in `eng` (CUDA events on the launching stream).
    -> (seconds, counters of the last pass, accumulated per-phase ms and launches, clocks, hits of a pass)."""
    n_hits = 0
    for _ in range(warmup):
        n_hits = eng.find_all(thr, fetch=False)
    sampler = ClockSampler(sampler_gpu) if sampler_gpu is not None else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = dict(ms_score=0.0, ms_score_round0=0.0, ms_trace=0.0, ms_other=0.0, ms_total=0.0, launches=0, score_launches=0)
    e0.record(stream)
    for _ in range(steps):
        n_hits = eng.find_all(thr, fetch=False)
        c = eng.last_counters
        for k in ("ms_score", "ms_score_round0", "ms_trace", "ms_other", "ms_total"):
            acc[k] += c[k]
        acc["launches"] += c["score_launches"] + c["trace_launches"] + c["other_launches"]
        acc["score_launches"] += c["score_launches"]
    e1.record(stream)
    barrier()
    seconds = e0.elapsed_time(e1) / 1e3
    clocks = sampler.stop() if sampler else None
    return seconds, dict(eng.last_counters), acc, clocks, n_hits


def k1_rates(cnt, acc, steps):
    """K1 GCUPS: all launches, round 0 (whole reads) and the child rounds (segments) separately."""
    s_all, s_r0 = acc["ms_score"] / 1e3, acc["ms_score_round0"] / 1e3
    s_ch = s_all - s_r0
    c_all, c_r0 = cnt["computed_cells"] * steps, cnt["computed_cells_round0"] * steps
    return {"all": c_all / s_all / 1e9 if s_all > 0 else None,
            "round0": c_r0 / s_r0 / 1e9 if s_r0 > 0 else None,
            "children": (c_all - c_r0) / s_ch / 1e9 if s_ch > 0 else None,
            "ms_round0_per_step": acc["ms_score_round0"] / steps, "ms_children_per_step": (acc["ms_score"] - acc["ms_score_round0"]) / steps}


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
    ap.add_argument("--no-extra", action="store_true", help="skip BASELINE.json configs[3] and [4]")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-batch strong-scaling leg")
    ap.add_argument("--extra-reads-5kb", type=int, default=400000)
    ap.add_argument("--extra-reads-long", type=int, default=100000)
    ap.add_argument("--skbio-reads", type=int, default=100000, help="reads of the headline batch seeded on the skbio path (extra config)")
    ap.add_argument("--strong-pieces", type=int, default=1, help="ranges per device in the strong-scaling leg")
    ap.add_argument("--e2e-pieces", type=int, default=1, help="ranges per device a batch is cut into inside the pool (lsw_pool_set_pieces)")
    ap.add_argument("--e2e-depth", type=int, default=2, help="contexts per device of the pool (batches in flight)")
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
    from lamprey_b200 import synth

    # ---- everything that forks happens before CUDA is touched -------------------------------
    concat, offs = workload_reads(args.workload, args.reads, rank, workers, world)
    extra_specs = [] if args.no_extra else [("synth_5kb_12pct", args.extra_reads_5kb), ("synth_20_50kb_12pct", args.extra_reads_long)]
    extra_reads = [(wl, n) + workload_reads(wl, n, rank, workers, world) for wl, n in extra_specs]
    from lamprey_b200 import seeding
    ps = seeding.PrimerSet(SCHEMA)
    seqs = [s for _, s in synth.schema_sequences(ps, args.workload)]
    names = [n for n, _ in synth.schema_sequences(ps, args.workload)]
    cpu_baseline, cpu_ref = None, None
    if rank == 0 and not args.no_cpu_baseline:
        n_names = synth.SCHEMA_NAMES.get(args.workload)
        if world == 1:
            # ~2 s per 2 kb read per core for the Cython port: size the sample for about --cpu-seconds
            per_read = 2.2 * (np.diff(offs).mean() / 2000.0) ** 1.0
            sample = int(max(cores, min(len(offs) - 1, cores * args.cpu_seconds / per_read)))
            reads = synth.reads_as_strings(concat[:offs[sample]], offs[:sample + 1])
            dt, hits_cpu, calls_cpu, cells_cpu = cpu_seed_sample(reads, args.thr, cores, n_names=n_names)
            cpu_baseline = {"value": sample / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                            "gcups": cells_cpu / dt / 1e9, "seconds": dt,
                            "sample": "first %d reads of the workload, Cython build of oracle/swalign_ref.py + "
                                      "seeding_ref.py (README.md:18-25 recipe), one process per core" % sample}
            # the interpreted figure beside it: one read per core
            dt_py, _, _, cells_py = cpu_seed_sample(reads[:cores], args.thr, cores, n_names=n_names, pure_python=True)
            cpu_baseline["pure_python"] = {"value": min(cores, len(reads)) / dt_py, "unit": "reads/s", "gcups": cells_py / dt_py / 1e9,
                                           "sample": "first %d reads, oracle/swalign_ref.py + seeding_ref.py interpreted" % min(cores, len(reads))}
            cpu_ref = {"n": sample, "py_hits": hits_cpu, "calls": calls_cpu, "cells": cells_cpu}
        else:
            # N > 1: no CPU timing (rank 0 shares the host with N - 1 other ranks), but the parity check still runs, against
            # the C restatement of the oracle on a small sample
            sample = min(len(offs) - 1, 8 * workers)
            reads = synth.reads_as_strings(concat[:offs[sample]], offs[:sample + 1])
            cpu_ref = {"n": sample, "py_hits": None, "calls": None, "cells": None}
        cpu_ref["c_rows"] = c_oracle_rows(reads, seqs, args.thr, workers if world > 1 else cores)

    # host side of a FASTQ -> hits run (FASTQ scanner, record expansion, per-read index: lsw_hostops.cpp) on this box's cores
    host_side = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_host
            n_hs = min(len(offs) - 1, 200000)
            host_side = bench_host.measure(concat[:offs[n_hs]], offs[:n_hs + 1], with_numpy=False)
        except Exception as e:                      # a side figure: never the reason a bench run fails
            host_side = {"error": repr(e)}

    log("inputs generated, CPU legs done" + ("; host side: %r" % (host_side.get("fastq") or host_side) if host_side else ""))
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # a second, CPU-side group: ranks that wait while rank 0 drives all GPUs (strong-scaling leg) must not spin a NCCL kernel on
        # the very GPUs being measured
        cpu_group = dist.new_group(backend="gloo")
    from lamprey_b200 import _native

    eng = _native.Engine(local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_scoring(2, -1, -1, -1)
    eng.set_queries(seqs)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # page-locked host buffers (lsw_host_alloc): the reference-facing calls take host memory
    pin_ascii = _native.PinnedArray(len(concat), np.uint8)
    pin_ascii.array[:] = concat
    pin_offs = _native.PinnedArray(len(offs), np.int64)
    pin_offs.array[:] = offs
    ascii_np, offs_np = pin_ascii.array, pin_offs.array

    # integer roofline of this box (not in MEASURED_PEAKS.json)
    peak = {}
    if rank == 0:
        for mode, name in ((0, "viaddmnmx_s16x2"), (1, "vimnmx3_s16x2"), (2, "iadd"), (3, "imad"), (8, "vimnmx_s16x2_2op"),
                           (9, "k1_inner_loop")):
            ops, _ = eng.measure_int_peak(mode, 20000 if mode != 9 else 4000)
            peak[name] = ops / 1e9

    # ---- value: inputs resident in HBM ---------------------------------------------------------
    t_up0 = time.perf_counter()
    eng.upload_reads(ascii_np, offs_np)
    upload_alone_ms = 1e3 * (time.perf_counter() - t_up0)
    t_value, cnt, acc, clocks, n_hits = timed_find_all(eng, torch, stream, args.thr, args.steps, args.warmup, barrier,
                                                       local_rank if rank == 0 else None)
    t_value, reads_total, cells_total, hits_total = gather_max_and_sum(
        t_value, [args.reads * args.steps, cnt["cells"] * args.steps, n_hits * args.steps])

    log("value leg done: %.1f ms/step" % (1e3 * t_value / args.steps))
    # parity against the CPU sample (same reads, same threshold): every hit, field by field
    parity = None
    if cpu_ref is not None:
        n = cpu_ref["n"]
        eng.upload_reads(ascii_np[:offs_np[n]], offs_np[:n + 1])
        sub = eng.find_all(args.thr)
        c2 = eng.last_counters
        got_rows = [tuple(int(x[k]) for k in ("read", "query", "start", "end", "matches", "mismatches", "score", "q_pos", "q_end", "depth"))
                    for x in sub]
        parity = {"reads": n, "hits": len(got_rows),
                  "hit_records_match_c_oracle": got_rows == [tuple(r) for r in cpu_ref["c_rows"]],
                  "fields": "read, query, start, end, matches, mismatches, score, q_pos, q_end, depth"}
        if cpu_ref["py_hits"] is not None:
            # what the reference itself returns per read: [(start, end, identity, primer_name)] in stable start order
            offs_h = np.searchsorted(sub["read"], np.arange(n + 1))
            ident = sub["matches"] / np.maximum(sub["matches"].astype(np.int64) + sub["mismatches"], 1)
            ok = True
            for i in range(n):
                mine = [(int(sub["start"][k]), int(sub["end"][k]), float(ident[k]), names[sub["query"][k]]) for k in range(offs_h[i], offs_h[i + 1])]
                ok = ok and mine == [tuple(h) for h in cpu_ref["py_hits"][i]]
            parity.update({"hit_tuples_match_cython_sample": ok, "calls_match": c2["tasks"] == cpu_ref["calls"],
                           "cells_match": c2["cells"] == cpu_ref["cells"]})
        eng.upload_reads(ascii_np, offs_np)

    # ---- extra configs (BASELINE.json configs[3], [4]): same engine, own timing ---------------------------------
    extra = []
    for wl, n_reads, x_concat, x_offs in extra_reads:
        x_seqs = [s for _, s in synth.schema_sequences(ps, wl)]
        eng.set_queries(x_seqs)
        eng.upload_reads(x_concat, x_offs)
        xs = max(2, min(args.steps, 3))
        t_x, c_x, a_x, clk_x, h_x = timed_find_all(eng, torch, stream, args.thr, xs, 3, barrier, local_rank if rank == 0 else None)
        t_x, r_x, cells_x = gather_max_and_sum(t_x, [n_reads * xs, c_x["cells"] * xs])
        if rank == 0:
            k1 = k1_rates(c_x, a_x, xs)
            pk = peak.get("viaddmnmx_s16x2", 0.0) * 4.0 / K1_DPX_PER_4_CELLS
            extra.append({"workload": wl, "reads_per_gpu": n_reads, "n_gpus": world, "mean_len": float(np.diff(x_offs).mean()),
                          "schema_sequences": len(x_seqs), "steps": xs, "warmup": 3, "value": r_x / t_x, "unit": "reads/s",
                          "ms_per_step": 1e3 * t_x / xs, "gcups": cells_x / t_x / 1e9,
                          "gcups_computed": c_x["computed_cells"] * xs * world / t_x / 1e9,
                          "tasks_per_step": c_x["tasks"], "hits_per_step": h_x, "rounds": c_x["rounds"], "fallbacks": c_x["fallbacks"],
                          "breakdown_ms_per_step": {k: a_x[k] / xs for k in ("ms_score", "ms_trace", "ms_other", "ms_total")},
                          "roofline": {"bound": "int_alu", "achieved": k1["all"], "peak": pk, "unit": "GCUPS",
                                       "frac": k1["all"] / pk if (pk and k1["all"]) else None, "k1_gcups": k1},
                          "device_bytes": c_x["device_bytes"], "reuse_dropped": c_x["reuse_dropped"], "clocks": clk_x})
    # the DEFAULT seeding path (scikit-bio StripedSmithWaterman semantics, lamprey.py:572-593) on a slice of the headline batch
    if not args.no_extra:
        n_sk = min(args.reads, args.skbio_reads)
        eng.set_queries(seqs)
        eng.set_ssw(2, -3, 5, 2)
        eng.upload_reads(ascii_np[:offs_np[n_sk]], offs_np[:n_sk + 1])
        xs = max(2, min(args.steps, 3))
        t_x, c_x, a_x, clk_x, h_x = timed_find_all(eng, torch, stream, args.thr, xs, 3, barrier, local_rank if rank == 0 else None)
        t_x, r_x, cells_x = gather_max_and_sum(t_x, [n_sk * xs, c_x["cells"] * xs])
        if rank == 0:
            extra.append({"workload": args.workload, "path": "skbio (StripedSmithWaterman semantics: lsw_set_ssw, lsw_ssw.cuh)",
                          "reads_per_gpu": n_sk, "n_gpus": world, "steps": xs, "warmup": 3, "value": r_x / t_x, "unit": "reads/s",
                          "ms_per_step": 1e3 * t_x / xs, "gcups": cells_x / t_x / 1e9, "tasks_per_step": c_x["tasks"],
                          "hits_per_step": h_x, "rounds": c_x["rounds"], "clocks": clk_x,
                          "note": "first CUDA version of this row: one thread per task, scalar int32 cells, no reuse of round 0"})
        eng.set_scoring(2, -1, -1, -1)
    del extra_reads
    eng.set_queries(seqs)
    eng.close()
    log("extra configs done")

    # ---- e2e: host buffers in, hits on the host out, every step, through the library's pool --------------------
    # lsw_pool_submit / lsw_pool_wait (include/lsw.h): the pool's worker threads run lsw_upload_reads -> lsw_find_all ->
    # lsw_fetch_hits_compact per batch; with two contexts per device one batch's PCIe copies overlap the next batch's kernels.
    # Every step uploads its own input and downloads its own hits inside the timed region.
    e2e = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not args.no_e2e:
        depth = max(1, args.e2e_depth)
        pool = _native.Pool([local_rank], depth)
        pool.configure(seqs, args.thr, compact=True, pieces=args.e2e_pieces)
        cap = int(n_hits * 1.05) + 4096
        outs = [_native.PinnedArray(cap, _native.HIT16_DT) for _ in range(depth)]

        def run_steps(total):
            tickets, last, tms, done_at = [], 0, [], []
            for k in range(total):
                tickets.append(pool.submit(ascii_np, offs_np, outs[k % depth].array))
                if len(tickets) >= depth:
                    h, _, tm = pool.wait(tickets.pop(0))
                    last = len(h); tms.append(tm); done_at.append(time.perf_counter())
            while tickets:
                h, _, tm = pool.wait(tickets.pop(0))
                last = len(h); tms.append(tm); done_at.append(time.perf_counter())
            return last, tms, done_at

        run_steps(max(depth, args.warmup - 1))
        barrier()
        e0.record(stream)
        t0 = time.perf_counter()
        got, tms, done_at = run_steps(args.steps)
        t_wall = time.perf_counter() - t0          # the pool's contexts run on their own streams: wall clock around the blocking waits
        e1.record(stream)
        barrier()
        t_e2e, reads_e2e = gather_max_and_sum(t_wall, [args.reads * args.steps])
        e2e = {"value": reads_e2e / t_e2e, "unit": "reads/s",
               "h2d_bytes_per_step": int(ascii_np.nbytes + offs_np.nbytes), "d2h_bytes_per_step": int(got * 16),
               "ms_per_step": 1e3 * t_e2e / args.steps, "contexts": depth, "pieces": args.e2e_pieces, "hits_per_step": int(got),
               "host_ms_per_call": {k: 1e3 * float(np.mean([tm[k + "_s"] for tm in tms])) for k in ("upload", "find", "fetch")},
               "first_upload_ms": upload_alone_ms,      # first lsw_upload_reads of the context: includes its device allocations
               # between consecutive completions the pipeline is full (the timed region above also pays for filling and draining
               # it: the first upload and the last download of the K steps overlap nothing)
               "steady_state_ms_per_step": 1e3 * (done_at[-1] - done_at[0]) / (len(done_at) - 1) if len(done_at) > 1 else None,
               "call": "per step: lsw_pool_submit + lsw_pool_wait on a pool of %d contexts of this GPU (the library's worker threads run "
                       "lsw_upload_reads -> lsw_find_all -> lsw_fetch_hits_compact; one batch's copies overlap the next batch's kernels); "
                       "read text and 16-byte hit records in page-locked host buffers" % depth}
        # the same loop with ONE context: no overlap at all (upload, seed, download one after the other)
        if rank == 0 and world == 1:
            pool.close()
            pool = _native.Pool([local_rank], 1)
            pool.configure(seqs, args.thr, compact=True)
            pool.wait(pool.submit(ascii_np, offs_np, outs[0].array))
            t0 = time.perf_counter()
            for _ in range(args.steps):
                pool.wait(pool.submit(ascii_np, offs_np, outs[0].array))
            t1 = time.perf_counter() - t0
            e2e["single_context"] = {"value": args.reads * args.steps / t1, "unit": "reads/s", "ms_per_step": 1e3 * t1 / args.steps}
        pool.close()
        for o in outs:
            o.close()

    # ---- strong scaling: ONE fixed batch (rank 0's) over all N GPUs, host-side gather included --------------------
    # The product path for several GPUs is one process with one thread + context per GPU (seeding.seed_reads(devices=...),
    # lsw_pool_*): rank 0 drives all N GPUs while the other ranks, whose engines are closed, wait at the barrier.
    log("e2e leg done")

    def cpu_barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier(group=cpu_group)

    strong = None
    if not args.no_strong:
        cpu_barrier()
        if rank == 0:
            ndev = min(world, torch.cuda.device_count())
            pool = _native.Pool(list(range(ndev)), 2)
            pool.configure(seqs, args.thr, compact=True, pieces=args.strong_pieces)
            cap = int(n_hits * 1.05) + 4096
            outs = [_native.PinnedArray(cap, _native.HIT16_DT) for _ in range(2)]
            for k in range(2):
                pool.wait(pool.submit(ascii_np, offs_np, outs[k].array))           # warm-up: buffers of every context
            lat = []
            for _ in range(2):                                                     # latency of one batch, nothing else in flight
                _, _, tm = pool.wait(pool.submit(ascii_np, offs_np, outs[0].array))
                lat.append(tm)
            t0 = time.perf_counter()
            tickets, tms = [], []
            for k in range(args.steps + 1):
                tickets.append(pool.submit(ascii_np, offs_np, outs[k % 2].array))
                if len(tickets) >= 2:
                    _, _, tm = pool.wait(tickets.pop(0)); tms.append(tm)
            while tickets:
                _, _, tm = pool.wait(tickets.pop(0)); tms.append(tm)
            t_s = time.perf_counter() - t0
            best = min(lat, key=lambda t: t["total_s"])
            strong = {"scaling": "strong", "n_gpus": ndev, "reads": args.reads, "workload": args.workload, "pieces": args.strong_pieces,
                      "value": args.reads * (args.steps + 1) / t_s, "unit": "reads/s", "ms_per_batch_pipelined": 1e3 * t_s / (args.steps + 1),
                      "ms_per_batch_latency": 1e3 * best["total_s"],
                      "latency_phases_ms": {k: 1e3 * best[k + "_s"] for k in ("upload", "find", "fetch", "gather_wait")},
                      "gather_ms": 1e3 * (best["fetch_s"] + best["gather_wait_s"]),
                      "gather_share_of_batch": (best["fetch_s"] + best["gather_wait_s"]) / best["total_s"] if best["total_s"] > 0 else None,
                      "reads_per_device": best["reads"][:ndev], "hits_per_device": best["hits"][:ndev],
                      "what": "one batch split over the GPUs in contiguous read ranges balanced by bases (lsw_pool_submit), seeded with "
                              "global read ids, hits copied from every GPU straight into their place in one host buffer (the gather: "
                              "a range waits for the hit counts of the ranges before it, then downloads); H2D and D2H inside the timing"}
            pool.close()
            for o in outs:
                o.close()
        cpu_barrier()
    log("strong-scaling leg done")

    if rank == 0:
        gcups = cells_total / t_value / 1e9
        k1 = k1_rates(cnt, acc, args.steps)
        peak_ops = peak.get("viaddmnmx_s16x2", 0.0)      # G lane-ops / s of the packed DPX instruction K1 is built from
        # the ceiling: K1 issues 5 packed DPX lane-ops per 4 cells and the DPX (ALU) pipe retires peak_ops of them per second
        peak_dpx = peak_ops * 4.0 / K1_DPX_PER_4_CELLS if peak_ops else None
        peak_survey = peak_ops * 2.0 / ALG_OPS_PER_CELL if peak_ops else None
        peak_inner = peak.get("k1_inner_loop", 0.0) * 4.0 / 9.0 if peak.get("k1_inner_loop") else None
        ncu = ncu_capture()
        roofline = {"bound": "int_alu", "kernel": "k_score<NR> (K1: DP fill + arg-max)",
                    "achieved": k1["all"], "peak": peak_dpx, "unit": "GCUPS",
                    "frac": (k1["all"] / peak_dpx) if (k1["all"] and peak_dpx) else None,
                    "peak_definition": "DPX-pipe ceiling of K1's own instruction mix: %d packed VIADDMNMX/VIMNMX3.S16x2 lane-ops per 4 cells "
                                       "at the VIADDMNMX.S16x2 rate lsw_measure_int_peak measures on this box (MEASURED_PEAKS.json has no "
                                       "integer rate)" % K1_DPX_PER_4_CELLS,
                    "traffic": ncu["dram_bytes_per_launch"] if ncu else None, "ncu": ncu,
                    "k1_gcups": k1,
                    "frac_round0": (k1["round0"] / peak_dpx) if (k1["round0"] and peak_dpx) else None,
                    "frac_children": (k1["children"] / peak_dpx) if (k1["children"] and peak_dpx) else None,
                    # SURVEY.md 8d's model (6 int32-lane ops per cell, packed x2): beaten by the real mix, so not a ceiling
                    "peak_survey_model": peak_survey, "frac_survey_model": (k1["all"] / peak_survey) if (k1["all"] and peak_survey) else None,
                    # K1's inner loop alone (column2 with its LDS.64 profile loads, K1's launch shape; intpeak mode 9)
                    "peak_inner_loop_microbench": peak_inner,
                    "frac_inner_loop_microbench": (k1["all"] / peak_inner) if (k1["all"] and peak_inner) else None,
                    "int_peak_glaneops": peak, "launches": acc["score_launches"],
                    "avg_launch_ms": acc["ms_score"] / max(1, acc["score_launches"]),
                    "k1_share_of_step": acc["ms_score"] / max(1e-9, acc["ms_total"]),
                    "computed_cells_per_step": cnt["computed_cells"], "reference_cells_per_step": cnt["cells"],
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
                           "rounds": cnt["rounds"], "fallbacks": cnt["fallbacks"]},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
                "gpu_launches": acc["launches"],
                "breakdown_ms_per_step": {k: acc[k] / args.steps for k in ("ms_score", "ms_trace", "ms_other", "ms_total")},
                "device_bytes": cnt["device_bytes"], "reuse_dropped": cnt["reuse_dropped"],
                "parity_vs_cpu_sample": parity, "strong_scaling": strong, "host_side": host_side, "extra_configs": extra}
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
