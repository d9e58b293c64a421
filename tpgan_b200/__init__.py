"""tpgan_b200 — B200-native (sm_100a) implementation of the TP-GAN G+D training-step hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all compute goes through the C ABI of
libtpgan_b200.so (include/tpgan_b200.h): tcgen05/TMEM/TMA implicit-GEMM convolutions and fused HBM-bound kernels.
"""
__version__ = "0.1.0"
