"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Builds oracle/_cy/{swalign_cy,seeding_cy}.*.so: the SAME files as swalign_ref.py and
seeding_ref.py compiled with untyped Cython (language_level 3), which is what
/root/reference/README.md:18-25 prescribes for the accelerated swalign fork.  This is the
"Cython swalign" CPU baseline bench.py reports.  Built artefacts are git-ignored but ship
to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_cy")


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    inc = sysconfig.get_paths()["include"]
    for src, mod in (("swalign_ref.py", "swalign_cy"), ("seeding_ref.py", "seeding_cy")):
        so = os.path.join(OUT, mod + ext)
        srcp = os.path.join(HERE, src)
        if not force and os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(srcp):
            continue
        pyx = os.path.join(OUT, mod + ".py")
        shutil.copyfile(srcp, pyx)
        c = os.path.join(OUT, mod + ".c")
        subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", c])
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fwrapv", "-I", inc, c, "-o", so])
        os.remove(pyx)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
