"""dfdclip_b200 — B200-native (sm_100a) implementation of the DFD-CLIP encoder + decoder hot path.

Layout:
  csrc/            hand-written CUDA kernels and the C ABI (include/dfdclip_b200.h)
  _native.py       ctypes binding of libdfdclip_b200.so (fails loudly if the library is missing)
  clip/            drop-in for the reference's ``src/clip`` loader API (``load``, ``available_models``)
  models.py        drop-in for the reference's ``src/models.py`` ``Detector`` / ``Decoder`` interface
  inference.py     video-level scoring driver (clip chunking, per-video mean, cross-rank gather)
  synthetic.py     seeded synthetic weights / clips shared by tests and bench
"""
__version__ = "0.1.0"
