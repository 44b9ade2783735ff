"""alpro_b200 — B200-native (sm_100a) forward/backward for the ALPRO video-language hot path.

Host side mirrors the reference's task-model interface (src/modeling/alpro_models.py); compute is hand-written CUDA
reached through the C-ABI in include/alpro_b200.h. No CPU fallback exists in the product path.
"""
__version__ = "0.1.0"
