"""gasfm_b200 -- B200-native (sm_100a) implementation of GASFM's graph-attention message passing.

Layout
  csrc/     hand-written CUDA kernels + the C ABI (include/gasfm_b200.h) -> lib/libgasfm_b200.so
  _lib.py   ctypes binding of the C ABI (fails loudly when the library is missing)
  index.py  per-scene CSR/CSC observation index and segment plans
  ops.py    torch.autograd wrappers of the kernels
  utils/    mirror of the reference's code/utils API for this path (SparseMat, M2sparse, graph wrappers)
  models/   mirror of the reference's code/models API (same classes, forward(data), state_dict names)
  dist.py   track-sharded multi-GPU execution (peer-memory exchange kernels; NCCL for set-up and scene-per-GPU training)
  scene.py / loss_functions.py / evaluation.py   sparse-first scene + sampling, ESFM loss, per-step metric
  plugin.py registers the mirror under the reference's module names (models.graph_attn_sfm, ...)
"""
__version__ = "0.1.0"
