"""Module boundary of /root/reference/lamprey.py:17: ``from skbio.alignment import StripedSmithWaterman``."""
from lamprey_b200.skbio_ssw import StripedSmithWaterman, AlignmentStructure, align_many  # noqa: F401
