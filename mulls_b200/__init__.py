"""mulls_b200 — B200-native implementation of the MULLS per-scan registration hot path.

`mulls_b200.abi`   ctypes mirror of include/mulls_b200/abi.h + library loader (fails loudly if unbuilt)
`mulls_b200.registration`  host-side mirror of lo::CRegistration / constraint_t / cloudblock_t
`mulls_b200.synth` seeded synthetic KITTI-shape scan pairs (benchmark inputs)
`mulls_b200/csrc`  the CUDA kernels and the C-ABI library (libmulls_b200.so)
"""
__version__ = "0.1.0"
