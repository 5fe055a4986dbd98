"""tpc_b200: B200-native rollout + A2C-update hot path of transformer-pointer-critic.

Python host code mirrors the reference's Agent / trainer / tester / env interfaces and calls hand-written
sm_100a CUDA kernels through the C-ABI in ``include/tpc.h`` (ctypes). PyTorch only allocates device memory.
"""
__version__ = "0.1.0"
