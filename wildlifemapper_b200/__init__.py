"""wildlifemapper_b200 -- B200-native (sm_100a) implementation of WildlifeMapper's tile-detection hot path.

Layout:
  csrc/                 hand-written CUDA kernels + the C ABI (include/wm_b200.h) -> libwm_b200.so
  lib.py                ctypes binding of the C ABI (fails loudly when the library is missing)
  ops.py                torch.library registration (``torch.ops.wm_b200.*``) + thin tensor wrappers
  engine.py             weight preparation, workspaces and the kernel schedule of the forward pass
  postprocess.py        PostProcess / sigmoid-top-k / NMS front-ends
  dist.py               tile sharding + NCCL gather of packed detections
  segment_anything/     drop-in replacement of the reference package (same names, signatures, state_dict)
"""
__version__ = "0.1.0"
