"""shapemol_b200: B200-native (sm_100a) implementation of ShapeMol's reverse-diffusion denoising step.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/shapemol_b200.h)
  _lib.py          ctypes binding of libshapemol_b200.so (no fallback)
  engine.py        host-side driver: weight packing cache, workspace, CUDA-graph step loop
  dropin/models/   drop-in mirror of the reference's `models` package (same class names, same
                   state_dict, same forward / sample_diffusion signatures)
  distributed.py   molecule-sharded multi-GPU sampling (one process per GPU, one final gather)
"""
__version__ = '0.1.0'
