// Exact fp32 division by a divisor that is reused many times (the per-channel range of the dynamic activation quantiser,
// its level count): Markstein's FMA sequence on the correctly rounded reciprocal.  Shared by gdn_tc.cu and act_quant.cu;
// b200lic_selftest_fast_div (gdn_tc.cu) counts mismatches against __fdiv_rn.
#pragma once
#include <stdint.h>

namespace b200lic {

// a / b correctly rounded, given y = RN(1 / b): two residual corrections.  Requires that the significand of b is not all
// ones (div_rn_ok) and that a / b neither overflows nor matters when it underflows (the quantiser clamps and rounds it).
__device__ __forceinline__ float div_rn(float a, float b, float y) {
  float q = __fmul_rn(a, y);
  float r = __fmaf_rn(-b, q, a);
  q = __fmaf_rn(r, y, q);
  r = __fmaf_rn(-b, q, a);
  return __fmaf_rn(r, y, q);
}
__device__ __forceinline__ bool div_rn_ok(float b) {
  const uint32_t u = __float_as_uint(b);
  return (u & 0x7fffffu) != 0x7fffffu && b >= 1e-30f && b <= 1e30f;
}

}  // namespace b200lic
