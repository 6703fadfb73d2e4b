"""rdo_ptq_b200 -- B200-native (sm_100a) implementation of the RDO-PTQ hot path.

Layout: `csrc/` CUDA kernels + C ABI (include/b200lic.h) -> `lib/libb200lic.so`; `ops.py` tensor-level calls;
`quantization/` drop-in for task-oriented-PTQ/quantization; `quant_int/` drop-in for light-uniform-PTQ/quant_int;
`codec/` compressai-style layers, entropy models and the three BASELINE model graphs.
Importing the package does not need a GPU; running any op does (no CPU fallback).
"""
__version__ = "0.1.0"
