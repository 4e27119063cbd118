"""B200-native render hot path for nim-raytracer (hand-written sm_100a CUDA behind a C ABI).

Package layout (only what the path needs):
  csrc/     CUDA kernels + the C ABI (libnrt.so; include/nrt.h)
  host/     C++ header mirroring the reference's renderer interface over the C ABI
  api.py    Python mirror of the same interface (ctypes)
  loaders.py, scenes.py, linalg.py   fixtures: meshes and the reference's scene files as data
"""
from . import api, linalg, loaders, scenes  # noqa: F401
