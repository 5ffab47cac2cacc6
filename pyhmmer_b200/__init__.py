"""pyhmmer_b200 -- a Blackwell-native engine behind pyhmmer's search path.

``pyhmmer_b200.easel`` / ``pyhmmer_b200.plan7`` / ``pyhmmer_b200.hmmer`` mirror the parts of
the reference's modules that `hmmsearch` / `hmmscan` / `Pipeline.search_hmm` touch; all dynamic
programming runs in hand-written sm_100a CUDA kernels behind the C ABI of ``include/b2h.h``.
"""
from . import _lib, easel, plan7          # noqa: F401  (importing _lib fails loudly if libb2h.so is missing)

__version__ = "0.1.0"
