"""dreamvla_b200 -- B200-native (sm_100a) implementation of the DreamVLA transformer forward/backward hot path.

Hot ops are hand-written CUDA kernels in libdvla_sm100.so (C ABI: include/dvla.h), driven from
torch.autograd.Function wrappers in dreamvla_b200.ops.  There is no CPU / PyTorch fallback for the kernels.
"""
__version__ = "0.1.0"
