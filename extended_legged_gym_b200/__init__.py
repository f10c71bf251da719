"""B200-native per-step hot path of extended_legged_gym.

Importing the package does not touch CUDA; the C-ABI library is loaded (and, if absent,
compiled with nvcc) the first time a host class needs it.  There is no CPU fallback.
"""
__version__ = "0.1.0"
