"""mcgra_b200 -- B200-native (sm_100a) hot path of MC-GRA's PGD graph-reconstruction attack.

Module names mirror the reference's `MC-GRA/` directory (topology_attack, base_attack, utils, hsic,
models.gcn, gcn_parameterized, main); their bodies call hand-written CUDA kernels through the C ABI in
include/mcgra.h.  There is no CPU / Triton / PyTorch fallback: without libmcgra_b200.so and a CUDA device
the hot-path entry points raise.
"""
__version__ = "0.1.0"
