"""snpgenie_b200: B200-native engine for SNPGenie's within-/between-group dN/dS path.

Only what the hot path needs lives here: csrc/ (CUDA kernels + the C ABI),
_abi.py (ctypes binding), engine.py (host mirror of the reference's engine),
hostio.py (FASTA / GTF ingest with the reference's semantics), synth.py
(benchmark-shaped synthetic alignments).  The Perl command-line host is under
perl/.  Importing `snpgenie_b200.engine` requires the built shared library.
"""
__version__ = "0.1"
