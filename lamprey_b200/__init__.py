"""lamprey_b200 -- B200-native batched Smith-Waterman engine for LAMPrey's seed stage and the rows next to it.

swalign   : drop-in for the third-party `swalign` module (LocalAlignment.align + batched forms)
skbio_ssw : drop-in for `skbio.alignment.StripedSmithWaterman` as LAMPrey uses it (default seeding branch, stage 2)
seeding   : mirrors of lamprey.py's PrimerSet / rc / alignSequence_* / findPrimerAlignments / findAllPrimerAlignments,
            and the batched seed_reads (one or several GPUs), seed_stream, seed_fastq, seed_fastq_files
chain     : host mirrors of removeOverlappingPrimerAlignments / extractAmpliconAroundTarget
pipeline  : processLamplicon's seed stage and stage 2 as batched calls
fastq     : FASTQ -> (bases, offsets) without Biopython (the library's multi-threaded scanner)
_native   : ctypes binding of liblsw.so (C ABI in include/lsw.h): Engine (one context), Pool (several contexts / GPUs)
synth     : synthetic ONT-like concatemer reads for tests and bench
"""
__all__ = ["swalign", "skbio_ssw", "seeding", "chain", "pipeline", "fastq", "synth"]
