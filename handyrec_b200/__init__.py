"""handyrec_b200 -- B200 (sm_100a) implementation of HandyRec's data-parallel hot path.

Layout
  csrc/        hand-written CUDA kernels + the C ABI (include/hrb200.h) -> libhrb200.so
  _lib.py      ctypes binding generated from the header (raises if the library is missing)
  kernels.py   torch-tensor front end of the C ABI (device memory / streams only)
  features/, layers/   host-side mirror of handyrec.features / handyrec.layers
There is no CPU fallback anywhere in this package; the CPU oracle lives in /oracle and is
test infrastructure only.
"""
__version__ = "0.1.0"
