"""toist_b200 — B200-native (sm_100a) implementation of the TOIST / MDETR training hot path.

Public surface mirrors the reference: `toist_b200.models.build_model(args)`, `toist_b200.util.misc.NestedTensor`.
The arithmetic lives in libtoist_b200.so (C ABI declared in include/toist_b200.h); nothing here falls back to
eager PyTorch or the CPU.
"""
__version__ = "0.1.0"
