"""efg_b200 — B200-native (sm_100a) implementation of V2AI/EFG's 3D-detection hot path.

Layers (bottom up):
  csrc/ + include/efgb200.h   hand-written CUDA kernels behind a flat C ABI (libefgb200.so)
  _lib / ops                  ctypes binding and tensor-level wrappers (no CPU fallback)
  _C                          the six callables of the reference's ``efg._C`` for this path
  operators                   ``efg.operators`` surface: Voxelization, dynamic_scatter, BoxAttnFunction
  spconv                      the ``spconv.pytorch`` names efg/modeling/backbones/sparse_net.py imports
  modeling / detectors        the models of the path (sparse backbones, Voxel-DETR, ConQueR, CenterPoint)
  compat                      installs the ``efg.*`` / ``spconv.*`` import aliases the playground uses
"""
__version__ = "0.1.0"
