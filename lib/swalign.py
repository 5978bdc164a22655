"""`swalign` as LAMPrey imports it (/root/reference/lamprey.py:14-16 puts ./lib first on sys.path
and then does `import swalign`; /root/reference/README.md:18-25 tells users to drop the accelerated
build into lib/).  Copy or symlink this file into LAMPrey's lib/ with this repo on PYTHONPATH and
`--swalign` runs on the B200 engine.  See INTEGRATION.md."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from lamprey_b200.swalign import (Alignment, AlignmentBatch, LocalAlignment,  # noqa: E402,F401
                                  NucleotideScoringMatrix, align_many, find_all_primer_alignments)
