"""world_size-2 gloo worker for tests/test_host.py: shards a synthetic read batch exactly as bench.py
does, "seeds" each shard with the host emulation of the kernels, and checks that the gathered
result equals the unsharded one (reads are independent units; no data-path collective)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import emu_lib  # noqa: E402
from lamprey_b200 import seeding, synth  # noqa: E402
from bench import gather_max_and_sum  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, ws = dist.get_rank(), dist.get_world_size()
    concat, offs = synth.generate(24, seed=5, mean_len=600)
    reads = synth.reads_as_strings(concat, offs)
    seqs = [s for _, s in seeding.PrimerSet(os.path.join(ROOT, "tests", "golden", "H3.3_set1.primers")).sequences()]
    lo, hi = seeding.shard_reads(offs, ws)[rank]
    hits, cnt = emu_lib.find_all(reads[lo:hi], seqs, 0.7)
    tmax, total_hits, total_cells = gather_max_and_sum(float(rank + 1), [len(hits), cnt["cells"]], device="cpu")
    if rank == 0:
        all_hits, all_cnt = emu_lib.find_all(reads, seqs, 0.7)
        assert tmax == float(ws)
        assert total_hits == len(all_hits) and total_cells == all_cnt["cells"], (total_hits, len(all_hits))
        print("GATHER_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
