// K9 (Gaussian-conditional likelihood) and K10 (factorised-prior likelihood), fused with latent rounding and the
// -log2 reduction that feeds bpp.  HBM-bound: float4 loads, warp-shuffle + one atomic per CTA.
//
// Semantics restated from compressai 1.2.4 (not vendored by the reference; reached from
// task-oriented-PTQ/models/nic_cvt.py:297-308):
//   quantize("dequantize"): y_hat = rint(y - mu) + mu
//   GaussianConditional:  lik = Phi((.5-a)/s) - Phi((-.5-a)/s), a = |y_hat - mu|, s = max(scale, 0.11),
//                         Phi(t) = 0.5*erfc(-t/sqrt(2)), lik >= 1e-9
//   EntropyBottleneck:    5-layer per-channel 1-3-3-3-3-1 softplus/tanh cumulative, lik = |sig(s*u) - sig(s*l)|.
#include "common.cuh"

namespace b200lic {

constexpr float kNegInvSqrt2 = -0.70710678118654752440f;  // float(-(2 ** -0.5))

__device__ __forceinline__ float gauss_one(float y, float mu, float sc, float scale_bound, float lik_bound,
                                           float* lik_out, float& bits) {
  const float yh = __fadd_rn(rintf(__fsub_rn(y, mu)), mu);
  const float a = fabsf(__fsub_rn(yh, mu));
  const float s = fmaxf(sc, scale_bound);
  const float up = 0.5f * erfcf(kNegInvSqrt2 * __fdiv_rn(0.5f - a, s));
  const float lo = 0.5f * erfcf(kNegInvSqrt2 * __fdiv_rn(-0.5f - a, s));
  const float lik = fmaxf(up - lo, lik_bound);
  if (lik_out) *lik_out = lik;
  bits -= log2f(lik);
  return yh;
}

// One CTA handles kChunkG consecutive elements of one sample (so the strided parameter views stay linear).
constexpr int kChunkG = 4096;

__global__ void __launch_bounds__(256)
    gaussian_lik_kernel(const float* __restrict__ y, const float* __restrict__ scales, const float* __restrict__ means,
                        int CHW, int chunks, long long pstride, float scale_bound, float lik_bound,
                        float* __restrict__ y_hat, float* __restrict__ lik, float* __restrict__ bits_out) {
  __shared__ float red[32];
  const int chunk = blockIdx.x % chunks, n = blockIdx.x / chunks;
  const size_t ybase = (size_t)n * CHW;
  const size_t pbase = (size_t)n * (size_t)pstride;
  const int beg = chunk * kChunkG, end = min(CHW, beg + kChunkG);
  float bits = 0.f;
  const bool vec = (CHW & 3) == 0 && (pstride & 3) == 0 &&
                   (((uintptr_t)y | (uintptr_t)scales | (uintptr_t)means | (uintptr_t)y_hat | (uintptr_t)lik) & 15) == 0;
  if (vec) {
    const float4* y4 = reinterpret_cast<const float4*>(y + ybase);
    const float4* s4 = reinterpret_cast<const float4*>(scales + pbase);
    const float4* m4 = means ? reinterpret_cast<const float4*>(means + pbase) : nullptr;
    float4* o4 = reinterpret_cast<float4*>(y_hat + ybase);
    float4* l4 = lik ? reinterpret_cast<float4*>(lik + ybase) : nullptr;
    for (int i = (beg >> 2) + threadIdx.x; i < (end >> 2); i += blockDim.x) {
      const float4 yv = __ldg(y4 + i), sv = __ldg(s4 + i);
      const float4 mv = m4 ? __ldg(m4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 o, l;
      o.x = gauss_one(yv.x, mv.x, sv.x, scale_bound, lik_bound, &l.x, bits);
      o.y = gauss_one(yv.y, mv.y, sv.y, scale_bound, lik_bound, &l.y, bits);
      o.z = gauss_one(yv.z, mv.z, sv.z, scale_bound, lik_bound, &l.z, bits);
      o.w = gauss_one(yv.w, mv.w, sv.w, scale_bound, lik_bound, &l.w, bits);
      o4[i] = o;
      if (l4) l4[i] = l;
    }
  } else {
    for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const float mu = means ? __ldg(means + pbase + i) : 0.f;
      float l;
      y_hat[ybase + i] = gauss_one(__ldg(y + ybase + i), mu, __ldg(scales + pbase + i), scale_bound, lik_bound, &l, bits);
      if (lik) lik[ybase + i] = l;
    }
  }
  if (bits_out) {
    const float tot = block_sum(bits, red);
    if (threadIdx.x == 0) atomicAdd(bits_out, tot);
  }
}

__global__ void __launch_bounds__(256) round_latent_kernel(const float* __restrict__ y, const float* __restrict__ means,
                                                            size_t n, float* __restrict__ y_hat) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float mu = means ? __ldg(means + i) : 0.f;
    y_hat[i] = __fadd_rn(rintf(__fsub_rn(__ldg(y + i), mu)), mu);
  }
}

// ---- K10 ------------------------------------------------------------------------------------------------
// params[c][58]: M0[3] M1[9] M2[9] M3[9] M4[3] | b0[3] b1[3] b2[3] b3[3] b4[1] | f0[3] f1[3] f2[3] f3[3]
struct FactorizedParams {
  float m[33];
  float b[13];
  float f[12];
};

__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__device__ __forceinline__ float logits_cumulative(const FactorizedParams& P, float v) {
  float l[3], t[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float u = P.m[k] * v + P.b[k];
    l[k] = u + P.f[k] * tanhf(u);
  }
#pragma unroll
  for (int layer = 1; layer <= 3; ++layer) {
    const float* M = P.m + 3 + 9 * (layer - 1);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float u = M[3 * k] * l[0] + M[3 * k + 1] * l[1] + M[3 * k + 2] * l[2] + P.b[3 * layer + k];
      t[k] = u + P.f[3 * layer + k] * tanhf(u);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) l[k] = t[k];
  }
  return P.m[30] * l[0] + P.m[31] * l[1] + P.m[32] * l[2] + P.b[12];
}

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int kChunkF = 2048;

__global__ void __launch_bounds__(256)
    factorized_lik_kernel(const float* __restrict__ z, const float* __restrict__ params,
                          const float* __restrict__ medians, int C, int HW, int chunks, float lik_bound,
                          float* __restrict__ z_hat, float* __restrict__ lik, float* __restrict__ bits_out) {
  __shared__ FactorizedParams P;
  __shared__ float red[32];
  const int chunk = blockIdx.x % chunks, plane = blockIdx.x / chunks, c = plane % C;
  if (threadIdx.x < 58) {
    const float raw = __ldg(params + (size_t)c * 58 + threadIdx.x);
    float* dst = reinterpret_cast<float*>(&P);
    float v = raw;
    if (threadIdx.x < 33) v = softplusf_(raw);
    else if (threadIdx.x >= 46) v = tanhf(raw);
    dst[threadIdx.x] = v;
  }
  __syncthreads();
  const float med = __ldg(medians + c);
  const size_t base = (size_t)plane * HW;
  const int beg = chunk * kChunkF, end = min(HW, beg + kChunkF);
  float bits = 0.f;
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const float zh = __fadd_rn(rintf(__fsub_rn(__ldg(z + base + i), med)), med);
    const float lo = logits_cumulative(P, zh - 0.5f);
    const float up = logits_cumulative(P, zh + 0.5f);
    const float sum = lo + up;
    const float sgn = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
    const float l = fmaxf(fabsf(sigm(sgn * up) - sigm(sgn * lo)), lik_bound);
    z_hat[base + i] = zh;
    if (lik) lik[base + i] = l;
    bits -= log2f(l);
  }
  if (bits_out) {
    const float tot = block_sum(bits, red);
    if (threadIdx.x == 0) atomicAdd(bits_out, tot);
  }
}

}  // namespace b200lic

using namespace b200lic;

extern "C" {

int b200lic_gaussian_lik_fwd(const float* y, const float* scales, const float* means, int N, int C, int HW,
                             long long param_batch_stride, float scale_bound, float lik_bound, float* y_hat,
                             float* lik, float* bits, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(y && scales && y_hat, "gaussian_lik_fwd: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && HW > 0, "gaussian_lik_fwd: bad shape (%d,%d,%d)", N, C, HW);
  const long long chw = (long long)C * HW;
  B200_REQUIRE(chw < 2147483647LL, "gaussian_lik_fwd: sample too large");
  B200_REQUIRE(param_batch_stride >= chw, "gaussian_lik_fwd: parameter batch stride %lld < C*HW", param_batch_stride);
  const int chunks = (int)((chw + kChunkG - 1) / kChunkG);
  gaussian_lik_kernel<<<(unsigned)(N * chunks), 256, 0, as_stream(stream)>>>(
      y, scales, means, (int)chw, chunks, param_batch_stride, scale_bound, lik_bound, y_hat, lik, bits);
  B200_LAUNCH_CHECK("gaussian_lik_kernel");
  return B200LIC_OK;
}

int b200lic_round_latent(const float* y, const float* means, size_t n, float* y_hat, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(y && y_hat, "round_latent: null pointer");
  if (n == 0) return B200LIC_OK;
  round_latent_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(y, means, n, y_hat);
  B200_LAUNCH_CHECK("round_latent_kernel");
  return B200LIC_OK;
}

int b200lic_factorized_lik_fwd(const float* z, const float* params, const float* medians, int N, int C, int HW,
                               float lik_bound, float* z_hat, float* lik, float* bits, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(z && params && medians && z_hat, "factorized_lik_fwd: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && HW > 0, "factorized_lik_fwd: bad shape (%d,%d,%d)", N, C, HW);
  const int chunks = (HW + kChunkF - 1) / kChunkF;
  factorized_lik_kernel<<<(unsigned)(N * C * chunks), 256, 0, as_stream(stream)>>>(z, params, medians, C, HW, chunks,
                                                                                  lik_bound, z_hat, lik, bits);
  B200_LAUNCH_CHECK("factorized_lik_kernel");
  return B200LIC_OK;
}

}  // extern "C"
