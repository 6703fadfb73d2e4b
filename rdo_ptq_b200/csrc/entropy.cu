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

constexpr float kInvSqrt2 = 0.70710678118654752440f;

// erfc(x) for x >= 0: Chebyshev fit t*exp(-x^2 + P(t)), t = 1/(1 + x/2), fractional error < 1.2e-7 everywhere
// (Numerical Recipes `erfcc`); 1 MUFU.RCP + 10 FMA + 1 MUFU.EX2 instead of libdevice's ~45-instruction erfcf, which is
// what keeps this kernel on the HBM roofline rather than the issue roofline.
__device__ __forceinline__ float erfc_pos(float x) {
  const float t = __fdividef(1.f, fmaf(0.5f, x, 1.f));
  float p = 0.17087277f;
  p = fmaf(p, t, -0.82215223f);
  p = fmaf(p, t, 1.48851587f);
  p = fmaf(p, t, -1.13520398f);
  p = fmaf(p, t, 0.27886807f);
  p = fmaf(p, t, -0.18628806f);
  p = fmaf(p, t, 0.09678418f);
  p = fmaf(p, t, 0.37409196f);
  p = fmaf(p, t, 1.00002368f);
  p = fmaf(p, t, -1.26551223f);
  return t * __expf(fmaf(-x, x, p));
}

__device__ __forceinline__ float gauss_one(float y, float mu, float sc, float scale_bound, float lik_bound,
                                           float* lik_out, float& bits) {
  const float yh = __fadd_rn(rintf(__fsub_rn(y, mu)), mu);      // latent symbol: bit-exact with round(y - mu) + mu
  const float a = fabsf(__fsub_rn(yh, mu));                      // a non-negative integer (up to rounding of mu)
  const float s = fmaxf(sc, scale_bound);
  const float inv = __fdividef(kInvSqrt2, s);
  // lik = Phi((.5-a)/s) - Phi((-.5-a)/s) = (erfc(u1) - erfc(u2)) / 2 with u1 = (a-.5)/(s sqrt2), u2 = (a+.5)/(s sqrt2) > 0
  const float u1 = (a - 0.5f) * inv, u2 = (a + 0.5f) * inv;
  const float e1 = erfc_pos(fabsf(u1)), e2 = erfc_pos(u2);
  const float E1 = u1 < 0.f ? 2.f - e1 : e1;
  const float lik = fmaxf(0.5f * (E1 - e2), lik_bound);
  if (lik_out) *lik_out = lik;
  bits -= __log2f(lik);
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

// likelihood of the symbol z_hat (a value med + k): |sig(s*c(z_hat+.5)) - sig(s*c(z_hat-.5))|
__device__ __forceinline__ float factorized_one(const FactorizedParams& P, float zh, float lik_bound) {
  const float lo = logits_cumulative(P, zh - 0.5f);
  const float up = logits_cumulative(P, zh + 0.5f);
  const float sum = lo + up;
  const float sgn = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
  return fmaxf(fabsf(sigm(sgn * up) - sigm(sgn * lo)), lik_bound);
}

// out-of-table symbols: kept out of line so the streaming loop stays at ~40 registers (4+ CTAs per SM)
__device__ __noinline__ float factorized_rare(const FactorizedParams* P, float zh, float lik_bound) {
  return factorized_one(*P, zh, lik_bound);
}

// The latent symbols of a channel are med + k for a few dozen integers k, so the 58-parameter CDF network is evaluated
// once per (channel, k) into a shared-memory table (|k| <= kTabR) and the element loop is two table reads: the kernel
// streams z at HBM speed instead of spending ~300 instructions per element.  Symbols outside the table (rare) take the
// direct path; both paths evaluate the same expression on the same operands, so results are identical.
constexpr int kTabR = 96;
constexpr int kTabN = 2 * kTabR + 1;

__global__ void __launch_bounds__(256, 4)
    factorized_lik_kernel(const float* __restrict__ z, const float* __restrict__ params,
                          const float* __restrict__ medians, const float* __restrict__ table, int N, int C, int HW,
                          int per_ch, float lik_bound, float* __restrict__ z_hat, float* __restrict__ lik,
                          float* __restrict__ bits_out) {
  __shared__ FactorizedParams P;
  __shared__ float red[32];
  __shared__ float t_lik[kTabN], t_bits[kTabN];
  // One resident wave, every CTA the same bytes.  per_ch > 0: gridDim.x = C * per_ch, CTA (c, r) takes the r-th share of
  // channel c's N planes (one symbol table per CTA).  per_ch == 0 (more channels than CTA slots): the N*C planes,
  // channel-major, are cut into gridDim.x equal contiguous ranges and the table changes with the channel.  The first
  // version launched C * splits CTAs of unequal residency (1.7 waves at C = 192): 46 % of the HBM peak (r1c).
  const long long planes = (long long)N * C;
  long long p_begin, p_end;
  if (per_ch > 0) {
    const int c = blockIdx.x / per_ch, r = blockIdx.x - c * per_ch;
    p_begin = (long long)c * N + (long long)N * r / per_ch;
    p_end = (long long)c * N + (long long)N * (r + 1) / per_ch;
  } else {
    p_begin = planes * blockIdx.x / gridDim.x;
    p_end = planes * (blockIdx.x + 1) / gridDim.x;
  }
  const bool vec = (HW & 3) == 0 && (((uintptr_t)z | (uintptr_t)z_hat | (uintptr_t)lik) & 15) == 0;
  const int q4 = HW >> 2;
  float bits = 0.f;
  long long pp = p_begin;
  while (pp < p_end) {
    const int c = (int)(pp / N), n0 = (int)(pp - (long long)c * N);
    const int n1 = (int)((long long)n0 + (p_end - pp) < (long long)N ? (long long)n0 + (p_end - pp) : (long long)N);
    pp += n1 - n0;
    {   // the first planes of the segment start their way from DRAM to L2 while the table is built
      const int lines = (HW * 4 + 127) >> 7, pre = (n1 - n0) < 24 ? (n1 - n0) : 24;
      for (int i = threadIdx.x; i < pre * lines; i += blockDim.x) {
        const int pl = i / lines, ln = i - pl * lines;
        const char* a = reinterpret_cast<const char*>(z + ((size_t)(n0 + pl) * C + c) * HW) + (size_t)ln * 128;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
      }
    }
    __syncthreads();                       // the previous segment's table is no longer read
    const float med = __ldg(medians + c);
    if (table != nullptr) {                // tables prepared by b200lic_factorized_table (a function of the parameters)
      const float* tc = table + (size_t)c * (2 * kTabN);
      for (int j = threadIdx.x; j < kTabN; j += blockDim.x) {
        t_lik[j] = __ldg(tc + j);
        t_bits[j] = __ldg(tc + kTabN + j);
      }
    }
    if (threadIdx.x < 58) {                // the direct path (fix-up pass) needs the reparametrised network either way
      const float raw = __ldg(params + (size_t)c * 58 + threadIdx.x);
      float* dst = reinterpret_cast<float*>(&P);
      float v = raw;
      if (threadIdx.x < 33) v = softplusf_(raw);
      else if (threadIdx.x >= 46) v = tanhf(raw);
      dst[threadIdx.x] = v;
    }
    __syncthreads();
    if (table == nullptr) {
      for (int j = threadIdx.x; j < kTabN; j += blockDim.x) {
        const float zh = __fadd_rn((float)(j - kTabR), med);
        const float l = factorized_rare(&P, zh, lik_bound);
        t_lik[j] = l;
        t_bits[j] = -log2f(l);
      }
      __syncthreads();
    }
    // Streaming pass: table symbols only.  A symbol outside the table (|k| > kTabR: rare) gets its z_hat here and
    // its likelihood in the fix-up pass below, so the hot loop carries no call and stays under 64 registers: four CTAs
    // (64 KB of loads in flight) per SM instead of three.
    bool rare = false;
    auto one = [&](float zv, float& zh, float& l) {
      const float k = rintf(__fsub_rn(zv, med));
      zh = __fadd_rn(k, med);
      const bool in = fabsf(k) <= (float)kTabR;
      const int j = in ? (int)k + kTabR : 0;
      l = t_lik[j];
      bits += in ? t_bits[j] : 0.f;
      rare |= !in;
    };
    const unsigned total = (unsigned)(n1 - n0) * (unsigned)q4, row4 = (unsigned)C * (unsigned)q4;
    const unsigned base4 = (unsigned)n0 * row4 + (unsigned)c * (unsigned)q4;      // float4 units; host checks numel < 2^34
    if (vec) {
      // planes n0..n1-1 of channel c as one flat float4 index space, four independent requests per thread in flight
      const float4* z4 = reinterpret_cast<const float4*>(z);
      for (unsigned i0 = threadIdx.x; i0 < total; i0 += 4u * blockDim.x) {
        float4 zv[4];
        unsigned off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned i = i0 + (unsigned)u * blockDim.x;
          const unsigned pl = i / (unsigned)q4;
          off[u] = base4 + pl * row4 + (i - pl * (unsigned)q4);
          if (i < total) zv[u] = __ldg(z4 + off[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (i0 + (unsigned)u * blockDim.x >= total) break;
          float4 o, l;
          one(zv[u].x, o.x, l.x);
          one(zv[u].y, o.y, l.y);
          one(zv[u].z, o.z, l.z);
          one(zv[u].w, o.w, l.w);
          reinterpret_cast<float4*>(z_hat)[off[u]] = o;
          if (lik) reinterpret_cast<float4*>(lik)[off[u]] = l;
        }
      }
    } else {
      for (int n = n0; n < n1; ++n) {
        const size_t base = ((size_t)n * C + c) * HW;
        for (int i = threadIdx.x; i < HW; i += blockDim.x) {
          float o, l;
          one(__ldg(z + base + i), o, l);
          z_hat[base + i] = o;
          if (lik) lik[base + i] = l;
        }
      }
    }
    if (__syncthreads_or(rare ? 1 : 0)) {
      // fix-up pass (some symbol of this segment lies outside the table): direct evaluation of those symbols only
      for (int n = n0; n < n1; ++n) {
        const size_t base = ((size_t)n * C + c) * HW;
        for (int i = threadIdx.x; i < HW; i += blockDim.x) {
          const float k = rintf(__fsub_rn(__ldg(z + base + i), med));
          if (fabsf(k) <= (float)kTabR) continue;
          const float l = factorized_rare(&P, __fadd_rn(k, med), lik_bound);
          bits -= log2f(l);
          if (lik) lik[base + i] = l;
        }
      }
    }
  }
  if (bits_out) {
    const float tot = block_sum(bits, red);
    if (threadIdx.x == 0) atomicAdd(bits_out, tot);
  }
}


// Symbol tables of every channel: table[c][0][k + R] = likelihood of the symbol median[c] + k, table[c][1][k + R] =
// -log2 of it, |k| <= R = kTabR.  A function of the prior's parameters only (frozen in PTQ), so callers build it once
// per model; the same expressions on the same operands as the in-kernel build, hence identical values.
__global__ void __launch_bounds__(256) factorized_table_kernel(const float* __restrict__ params,
                                                                const float* __restrict__ medians, float lik_bound,
                                                                float* __restrict__ table) {
  __shared__ FactorizedParams P;
  const int c = blockIdx.x;
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
  float* tc = table + (size_t)c * (2 * kTabN);
  for (int j = threadIdx.x; j < kTabN; j += blockDim.x) {
    const float zh = __fadd_rn((float)(j - kTabR), med);
    const float l = factorized_rare(&P, zh, lik_bound);
    tc[j] = l;
    tc[kTabN + j] = -log2f(l);
  }
}

// ---- K9 backward -------------------------------------------------------------------------------------------
// Gradient of the Gaussian-conditional likelihood (compressai autograd of GaussianConditional.forward in eval mode,
// reached from nic_cvt.py:300-308) for the rate term of RateDistortionLoss (losses/losses.py:20-28).
//   v = y_hat - mu, a = |v|, s = LowerBound(scale, 0.11), lik = LowerBound(Phi((.5-a)/s) - Phi((-.5-a)/s), 1e-9)
//   dlik/da = (phi(u_lo) - phi(u_up)) / s,  dlik/ds = (phi(u_lo) u_lo - phi(u_up) u_up) / s,  u = (+-.5 - a)/s
// upstream: g = g_lik[i] + g_bits * d(-log2 lik)/dlik.  LowerBound passes a gradient iff x >= bound or g < 0.
// ste == 0: torch.round has zero gradient, so only d_scales (and d_means = g_yhat through the "+ means") are non-zero.
// ste != 0: latent rounding is straight-through (round_ste of quantizer.py:64-68, as layer_opt.py:69 applies to y):
//           d_y = g_yhat + g dlik/dv, d_means = -g dlik/dv.
constexpr float kInvSqrt2Pi = 0.39894228040143267794f;
constexpr float kInvLn2 = 1.44269504088896340736f;

__device__ __forceinline__ void gauss_bwd_one(float yh, float mu, float sc, float g_lik, float g_bits, float g_yh,
                                              float scale_bound, float lik_bound, int ste, float& d_y, float& d_s,
                                              float& d_m) {
  const float v = __fsub_rn(yh, mu);
  const float a = fabsf(v);
  const float s = fmaxf(sc, scale_bound);
  const float rs = __fdividef(1.f, s);
  const float inv = kInvSqrt2 * rs;
  const float u1 = (a - 0.5f) * inv, u2 = (a + 0.5f) * inv;            // same operands as the forward kernel
  const float e1 = erfc_pos(fabsf(u1)), e2 = erfc_pos(u2);
  const float E1 = u1 < 0.f ? 2.f - e1 : e1;
  const float lik_raw = 0.5f * (E1 - e2);
  const float lik = fmaxf(lik_raw, lik_bound);
  float g = g_lik - g_bits * kInvLn2 * __fdividef(1.f, lik);
  if (!(lik_raw >= lik_bound || g < 0.f)) g = 0.f;
  const float up = (0.5f - a) * rs, lo = (-0.5f - a) * rs;
  const float p_up = kInvSqrt2Pi * __expf(-0.5f * up * up), p_lo = kInvSqrt2Pi * __expf(-0.5f * lo * lo);
  const float dl_da = (p_lo - p_up) * rs;
  const float dl_ds = (p_lo * lo - p_up * up) * rs;
  float ds = g * dl_ds;
  if (!(sc >= scale_bound || ds < 0.f)) ds = 0.f;
  d_s = ds;
  const float sgn = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
  const float dv = g * dl_da * sgn;
  if (ste) {
    d_y = g_yh + dv;
    d_m = -dv;
  } else {
    d_y = 0.f;
    d_m = g_yh;
  }
}

__global__ void __launch_bounds__(256)
    gaussian_lik_bwd_kernel(const float* __restrict__ y_hat, const float* __restrict__ scales,
                            const float* __restrict__ means, const float* __restrict__ g_lik,
                            const float* __restrict__ g_bits, const float* __restrict__ g_yhat, int CHW, int chunks,
                            long long pstride, long long gstride, float scale_bound, float lik_bound, int ste,
                            float* __restrict__ d_y, float* __restrict__ d_scales, float* __restrict__ d_means) {
  const int chunk = blockIdx.x % chunks, n = blockIdx.x / chunks;
  const size_t ybase = (size_t)n * CHW, pbase = (size_t)n * (size_t)pstride, gbase = (size_t)n * (size_t)gstride;
  const int beg = chunk * kChunkG, end = min(CHW, beg + kChunkG);
  const float gb = g_bits ? __ldg(g_bits) : 0.f;
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const float mu = means ? __ldg(means + pbase + i) : 0.f;
    float dy, ds, dm;
    gauss_bwd_one(__ldg(y_hat + ybase + i), mu, __ldg(scales + pbase + i), g_lik ? __ldg(g_lik + ybase + i) : 0.f, gb,
                  g_yhat ? __ldg(g_yhat + ybase + i) : 0.f, scale_bound, lik_bound, ste, dy, ds, dm);
    if (d_y) d_y[ybase + i] = dy;
    d_scales[gbase + i] = ds;
    if (d_means) d_means[gbase + i] = dm;
  }
}

// ---- K10 backward ------------------------------------------------------------------------------------------
// d lik / d z_hat of the factorised prior by forward-mode differentiation of the per-channel cumulative network
// (parameters are frozen during PTQ: only the latent receives a gradient, and only under straight-through rounding --
// with torch.round the gradient of EntropyBottleneck.forward w.r.t. z is identically zero).
__device__ __forceinline__ void logits_cumulative_jvp(const FactorizedParams& P, float v, float& c, float& dc) {
  float l[3], d[3], t[3], td[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float u = P.m[k] * v + P.b[k];
    const float th = tanhf(u);
    l[k] = u + P.f[k] * th;
    d[k] = P.m[k] * (1.f + P.f[k] * (1.f - th * th));
  }
#pragma unroll
  for (int layer = 1; layer <= 3; ++layer) {
    const float* M = P.m + 3 + 9 * (layer - 1);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float u = M[3 * k] * l[0] + M[3 * k + 1] * l[1] + M[3 * k + 2] * l[2] + P.b[3 * layer + k];
      const float du = M[3 * k] * d[0] + M[3 * k + 1] * d[1] + M[3 * k + 2] * d[2];
      const float th = tanhf(u);
      t[k] = u + P.f[3 * layer + k] * th;
      td[k] = du * (1.f + P.f[3 * layer + k] * (1.f - th * th));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      l[k] = t[k];
      d[k] = td[k];
    }
  }
  c = P.m[30] * l[0] + P.m[31] * l[1] + P.m[32] * l[2] + P.b[12];
  dc = P.m[30] * d[0] + P.m[31] * d[1] + P.m[32] * d[2];
}

// returns raw (un-bounded) likelihood and its derivative w.r.t. the symbol value
__device__ __noinline__ void factorized_grad_one(const FactorizedParams* P, float zh, float* lik_raw, float* dlik) {
  float lo, dlo, up, dup;
  logits_cumulative_jvp(*P, zh - 0.5f, lo, dlo);
  logits_cumulative_jvp(*P, zh + 0.5f, up, dup);
  const float sum = lo + up;
  const float sgn = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
  const float su = sigm(sgn * up), sl = sigm(sgn * lo);
  const float diff = su - sl;
  const float sd = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
  *lik_raw = fabsf(diff);
  *dlik = sd * sgn * (su * (1.f - su) * dup - sl * (1.f - sl) * dlo);
}

__global__ void __launch_bounds__(256)
    factorized_lik_bwd_kernel(const float* __restrict__ z_hat, const float* __restrict__ params,
                              const float* __restrict__ medians, const float* __restrict__ g_lik,
                              const float* __restrict__ g_bits, const float* __restrict__ g_zhat, int N, int C, int HW,
                              int splits, float lik_bound, int ste, float* __restrict__ d_z) {
  __shared__ FactorizedParams P;
  __shared__ float t_lik[kTabN], t_dl[kTabN];
  const int c = blockIdx.x % C, split = blockIdx.x / C;
  if (threadIdx.x < 58) {
    const float raw = __ldg(params + (size_t)c * 58 + threadIdx.x);
    float v = raw;
    if (threadIdx.x < 33) v = softplusf_(raw);
    else if (threadIdx.x >= 46) v = tanhf(raw);
    reinterpret_cast<float*>(&P)[threadIdx.x] = v;
  }
  __syncthreads();
  const float med = __ldg(medians + c);
  if (ste) {
    for (int j = threadIdx.x; j < kTabN; j += blockDim.x)
      factorized_grad_one(&P, __fadd_rn((float)(j - kTabR), med), &t_lik[j], &t_dl[j]);
  }
  __syncthreads();
  const float gb = g_bits ? __ldg(g_bits) : 0.f;
  for (int n = split; n < N; n += splits) {
    const size_t base = ((size_t)n * C + c) * HW;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      float out = 0.f;
      if (ste) {
        const float zh = __ldg(z_hat + base + i);
        const float k = rintf(__fsub_rn(zh, med));
        float lr, dl;
        if (fabsf(k) <= (float)kTabR) {
          lr = t_lik[(int)k + kTabR];
          dl = t_dl[(int)k + kTabR];
        } else {
          factorized_grad_one(&P, zh, &lr, &dl);
        }
        float g = (g_lik ? __ldg(g_lik + base + i) : 0.f) - gb * kInvLn2 * __fdividef(1.f, fmaxf(lr, lik_bound));
        if (!(lr >= lik_bound || g < 0.f)) g = 0.f;
        out = g * dl + (g_zhat ? __ldg(g_zhat + base + i) : 0.f);
      }
      d_z[base + i] = out;
    }
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

int b200lic_factorized_table_floats(void) { return 2 * kTabN; }

int b200lic_factorized_table(const float* params, const float* medians, int C, float lik_bound, float* table,
                             b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(params && medians && table, "factorized_table: null pointer");
  B200_REQUIRE(C > 0, "factorized_table: C=%d", C);
  factorized_table_kernel<<<(unsigned)C, 256, 0, as_stream(stream)>>>(params, medians, lik_bound, table);
  B200_LAUNCH_CHECK("factorized_table_kernel");
  return B200LIC_OK;
}

int b200lic_factorized_lik_fwd(const float* z, const float* params, const float* medians, const float* table, int N,
                               int C, int HW, float lik_bound, float* z_hat, float* lik, float* bits,
                               b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(z && params && medians && z_hat, "factorized_lik_fwd: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && HW > 0, "factorized_lik_fwd: bad shape (%d,%d,%d)", N, C, HW);
  B200_REQUIRE((long long)N * C * HW < (1LL << 34), "factorized_lik_fwd: tensor too large (%d,%d,%d)", N, C, HW);
  // one resident wave of 4 CTAs per SM: whole shares of a channel per CTA when the channels are fewer than the slots
  const long long slots = 4LL * num_sms(), planes = (long long)N * C;
  int per_ch = (int)(slots / C);
  if (per_ch > N) per_ch = N;
  long long grid = per_ch > 0 ? (long long)C * per_ch : (slots < planes ? slots : planes);
  factorized_lik_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(z, params, medians, table, N, C, HW, per_ch,
                                                                       lik_bound, z_hat, lik, bits);
  B200_LAUNCH_CHECK("factorized_lik_kernel");
  return B200LIC_OK;
}

int b200lic_gaussian_lik_bwd(const float* y_hat, const float* scales, const float* means, const float* g_lik,
                             const float* g_bits, const float* g_yhat, int N, int C, int HW,
                             long long param_batch_stride, long long grad_batch_stride, float scale_bound,
                             float lik_bound, int ste, float* d_y, float* d_scales, float* d_means,
                             b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(y_hat && scales && d_scales, "gaussian_lik_bwd: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && HW > 0, "gaussian_lik_bwd: bad shape (%d,%d,%d)", N, C, HW);
  B200_REQUIRE(!means == !d_means, "gaussian_lik_bwd: means and d_means go together");
  const long long chw = (long long)C * HW;
  B200_REQUIRE(chw < 2147483647LL, "gaussian_lik_bwd: sample too large");
  B200_REQUIRE(param_batch_stride >= chw && grad_batch_stride >= chw, "gaussian_lik_bwd: batch stride < C*HW");
  const int chunks = (int)((chw + kChunkG - 1) / kChunkG);
  gaussian_lik_bwd_kernel<<<(unsigned)(N * chunks), 256, 0, as_stream(stream)>>>(
      y_hat, scales, means, g_lik, g_bits, g_yhat, (int)chw, chunks, param_batch_stride, grad_batch_stride, scale_bound,
      lik_bound, ste, d_y, d_scales, d_means);
  B200_LAUNCH_CHECK("gaussian_lik_bwd_kernel");
  return B200LIC_OK;
}

int b200lic_factorized_lik_bwd(const float* z_hat, const float* params, const float* medians, const float* g_lik,
                               const float* g_bits, const float* g_zhat, int N, int C, int HW, float lik_bound, int ste,
                               float* d_z, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(z_hat && params && medians && d_z, "factorized_lik_bwd: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && HW > 0, "factorized_lik_bwd: bad shape (%d,%d,%d)", N, C, HW);
  int splits = (4 * num_sms() + C - 1) / C;
  if (splits > N) splits = N;
  if (splits < 1) splits = 1;
  factorized_lik_bwd_kernel<<<(unsigned)(C * splits), 256, 0, as_stream(stream)>>>(
      z_hat, params, medians, g_lik, g_bits, g_zhat, N, C, HW, splits, lik_bound, ste, d_z);
  B200_LAUNCH_CHECK("factorized_lik_bwd_kernel");
  return B200LIC_OK;
}

}  // extern "C"
