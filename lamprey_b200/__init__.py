"""lamprey_b200 -- B200-native batched Smith-Waterman seeding engine for LAMPrey's seed stage.

swalign   : drop-in for the third-party `swalign` module (LocalAlignment.align + batched forms)
seeding   : mirror of lamprey.py's PrimerSet / rc / findPrimerAlignments / findAllPrimerAlignments
_native   : ctypes binding of liblsw.so (C ABI in include/lsw.h)
synth     : synthetic ONT-like concatemer reads for tests and bench
"""
__all__ = ["swalign", "seeding", "synth"]
