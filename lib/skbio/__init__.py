"""``lib/skbio``: found before the real scikit-bio because LAMPrey puts ./lib first on sys.path (/root/reference/lamprey.py:14),
so that ``from skbio.alignment import StripedSmithWaterman`` (lamprey.py:17) resolves to the GPU engine.  Only the one name
LAMPrey imports from scikit-bio is provided."""
