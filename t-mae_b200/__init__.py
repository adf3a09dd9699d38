"""tmae_b200: B200-native (sm_100a) implementation of T-MAE's sparse-window voxel-encoder hot path.

Host side mirrors the pcdet module API (vfe / backbone_3d `forward(batch_dict)`, `get_loss()`);
all arithmetic on the path runs in hand-written CUDA kernels behind the C-ABI library
`csrc/libtmae_sm100.so` (see include/tmae_sm100.h).  There is no CPU fallback: importing the
compute modules without the built library raises.
"""
__version__ = "0.1.0"
