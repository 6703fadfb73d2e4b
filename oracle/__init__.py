"""CPU oracle for the RDO-PTQ hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain PyTorch-CPU fp32 arithmetic, the algorithm of the
reference's fake-quantized LIC forward/backward (Eric-qi/RDO-PTQ) and of the
`compressai==1.2.4` pieces it depends on (GDN, EntropyBottleneck,
GaussianConditional, MaskedConv2d, the Balle2018 / Minnen2018 / Cheng2020 graphs),
which are NOT vendored under /root/reference (requirements.txt:1).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.  The product package (`rdo_ptq_b200`) never
does: its ops call the sm_100a CUDA library through the C ABI and raise when the
library is missing.

Pinning status
--------------
* quantizer arithmetic (`oracle.quantizers`): PINNED against the reference's own
  `task-oriented-PTQ/quantization/quantizer.py` and
  `light-uniform-PTQ/quant_int/quantizer.py`, imported from /root/reference by
  `oracle/make_golden.py` (they depend on torch only); the comparison is bit-exact and
  the resulting vectors are committed under `tests/golden/`.
* wrappers and the calibration loop (`oracle.quant_wrap`, `oracle.calib`): PINNED against the
  reference's own `task-oriented-PTQ/quantization` and `light-uniform-PTQ/quant_int` packages.
  `oracle/_ref_shim.py` registers stand-ins for compressai / timm / pytorch_msssim (mapped to
  `oracle.codec`) in `sys.modules`, imports the UNMODIFIED reference files and
  `oracle/make_golden.py::wrap_vectors` asserts bit-exact agreement for the `QuantModel` rewrite,
  FP / W8 / W8A8 / W4-with-8-bit-head forwards of the three codecs (per wrapped module),
  `save_inp_oup_data`, `LossFunction`, `layer_reconstruction` / `block_reconstruction` walks
  (alpha and hardened weights of every unit) and LU's `QuantModel`; the reference's outputs are
  committed as `tests/golden/wrap_ref.pt`.
* compressai pieces (`oracle.codec`): the reference ships no tests or golden
  vectors for them and compressai is not installable here, so they
  are restated from the published 1.2.4 semantics and checked against closed forms
  (`tests/test_oracle_closed_forms.py`).  For these rows: "parity unpinned".
"""
