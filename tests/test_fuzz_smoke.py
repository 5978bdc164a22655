"""A few seconds of the CPU fuzzers in every run of the CPU tier (the long runs are recorded in profiles/r02_fuzz_emu.txt):
tools/fuzz_emu.py (lane logic of both seeding paths, align(), stage 2, against the C oracles) and tools/fuzz_emu_chain.py
(the chain step's device logic against the host mirror and the oracle), on fixed seeds."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.parametrize("tool,seconds,seed", [("fuzz_emu.py", 10, 424242), ("fuzz_emu_chain.py", 5, 31337)])
def test_cpu_fuzzers_find_nothing(tool, seconds, seed):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), str(seconds), str(seed)], cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = p.stdout.decode()
    assert p.returncode == 0 and " 0 mismatching cases" in out, out[-2000:]
