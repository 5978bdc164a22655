"""C-ABI library: loads, exports every symbol include/lsw.h declares, fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from lamprey_b200 import _native


def _declared():
    src = open(os.path.join(ROOT, "include", "lsw.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lsw_[a-z_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_native.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.lsw_version() == 1


def test_record_sizes_match_header():
    assert _native.TASK_DT.itemsize == 16
    assert _native.RESULT_DT.itemsize == 40
    assert _native.HIT_DT.itemsize == 32


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.Engine(0)
    from lamprey_b200 import swalign
    al = swalign.LocalAlignment(swalign.NucleotideScoringMatrix(2, -1))
    with pytest.raises(RuntimeError):
        al.align("ACGT", "ACG")


def test_product_never_imports_oracle():
    # the oracle is test infrastructure: nothing under lamprey_b200/ or lib/ may reference it
    for d in ("lamprey_b200", "lib"):
        for base, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                    txt = open(os.path.join(base, f)).read()
                    assert "import oracle" not in txt and "from oracle" not in txt and "sw_oracle" not in txt, f


def test_unsupported_options_raise():
    from lamprey_b200 import swalign
    sm = swalign.NucleotideScoringMatrix(2, -1)
    for kw in (dict(globalalign=True), dict(full_query=True), dict(gap_extension_decay=0.5), dict(wildcard="N"),
               dict(prefer_gap_runs=False), dict(gap_penalty=-2, gap_extension_penalty=-1)):
        with pytest.raises(NotImplementedError):
            swalign.LocalAlignment(sm, **kw)
