"""
findex_b200 — the B200-native FM-index search path behind findex's operator API.

The product is `libfmgpu.so` (`findex_b200/csrc/`, C ABI in `include/fmgpu.h`); this package is its Python host side:

    fmindex   ctypes mirror with the reference's member names: GpuFMSearcher (= NaiveFMSearcher on the GPU), ReTree / ThompsonNFA
              (regex engines), LCPSearcher, pinned host buffers, device-resident calls
    dfa       mirror of the DFA engine (dfa.scala)
    sharded   multi-GPU plumbing: replicated index, sharded batch, device-resident exchange by peer stores
    synth     seeded synthetic workloads of the BASELINE configs
    build     nvcc build of libfmgpu.so for sm_100a

Nothing here computes a result on the CPU: every call goes through the library and fails loudly without it or without a CUDA device.
"""
from .fmindex import FmxError, GpuFMSearcher, ReTree, ThompsonNFA  # noqa: F401

__all__ = ["FmxError", "GpuFMSearcher", "ReTree", "ThompsonNFA"]
