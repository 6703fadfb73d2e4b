// K7 (per-channel weight range + fake-quant) and K6 (AdaRound fwd / fused bwd + regulariser + Adam).
// Compiled with -fmad=false: the integer codes are a bit-exact contract, so every fp32 operation is
// issued exactly as the reference issues it (true division, rint, separate mul/add).
//
// Reference arithmetic restated (not translated): task-oriented-PTQ/quantization/quantizer.py
//   :281-298 range -> (delta, zero_point); :175-177 fake-quant; :437-452 AdaRound forward;
//   :454-466 alpha init; layer_opt.py:160-165 rounding regulariser; torch.optim.Adam.
#include <initializer_list>
#include <stdlib.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "prepared.cuh"

namespace b200lic {

constexpr float kGamma = -0.1f;   // AdaRound's rectified sigmoid stretches (0, 1) to (gamma, zeta) = (-0.1, 1.1)
constexpr float kStretch = 1.2f;  // fl32(zeta - gamma) = fl32(1.2000000000000002)

// ---- K7a: one CTA per quantisation channel; tensor viewed as [outer, ch, inner] ---------------------
__global__ void __launch_bounds__(256) wq_minmax_kernel(const float* __restrict__ w, int outer, int ch, int inner,
                                                         int n_bits, int scale_variant, int symmetric,
                                                         float* __restrict__ delta, float* __restrict__ zp) {
  const int c = blockIdx.x;
  float mn = INFINITY, mx = -INFINITY;
  const size_t per_outer = (size_t)ch * inner;
  const size_t total = (size_t)outer * inner;
  for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
    const size_t o = i / inner, k = i - o * inner;
    const float v = __ldg(w + o * per_outer + (size_t)c * inner + k);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  __shared__ float smn[8], smx[8];
  mn = warp_min(mn);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) {
    smn[threadIdx.x >> 5] = mn;
    smx[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
      mn = fminf(mn, smn[i]);
      mx = fmaxf(mx, smx[i]);
    }
    // float64 range arithmetic on the `.item()` values, then one rounding to fp32 (quantizer.py:282-297)
    double lo = fmin((double)mn, 0.0), hi = fmax((double)mx, 0.0);
    if (scale_variant) {
      lo = lo * (double)(n_bits + 2) / 8.0;
      hi = hi * (double)(n_bits + 2) / 8.0;
    }
    if (symmetric) {
      const double a = fmax(fabs(lo), hi);
      lo = lo < 0 ? -a : 0.0;
      hi = a;
    }
    const double levels = (double)((1 << n_bits) - 1);
    float d = (float)((hi - lo) / levels);
    d = fmaxf(d, 1e-8f);
    // (-x_min / delta) on a 0-dim tensor is delta.reciprocal() * (-x_min) in fp32 (Tensor.__rdiv__)
    const float z = rintf(__fmul_rn(__frcp_rn(d), (float)(-lo)));
    delta[c] = d;
    zp[c] = z;
  }
}

// ---- K7c: search-based / moment-based ranges (quantizer.py:300-370; LU quantizer.py:265-280) ---------------------
// One CTA per quantisation channel.  'mse' / 'l1' / 'l2': the slice's (min, max) are shrunk in n_steps steps of
// `shrink`; every candidate range is scored by mean |x - fake_quant(x)|^p and the first strictly best candidate wins
// (the reference's `if score < best_score`).  The fake-quant of a candidate is issued in the reference's fp32 order
// (quantizer.py:375-382); scores are accumulated in fp64 (the reference's fp32 mean differs in the last bits only,
// which matters for exact ties alone).  'gaussian': range = mean -+ 6 * var(unbiased) (:318-336).
struct CandRange { float d, z; };

__device__ __forceinline__ CandRange cand_range(float mx, float mn, float f, float levels) {
  const float nh = __fmul_rn(mx, f), nl = __fmul_rn(mn, f);
  CandRange r;
  r.d = fmaxf(__fdiv_rn(__fsub_rn(nh, nl), levels), 1e-8f);
  r.z = rintf(__fdiv_rn(-nl, r.d));
  return r;
}

__device__ __forceinline__ double block_sum_f64(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];   // fixed order, every thread gets the total
  return t;
}

constexpr int kSearchChunk = 8;

__global__ void __launch_bounds__(256) wq_search_kernel(const float* __restrict__ w, int outer, int ch, int inner,
                                                         int n_bits, int method, int n_steps, double shrink, float p,
                                                         int symmetric, float* __restrict__ delta,
                                                         float* __restrict__ zp) {
  const int c = blockIdx.x;
  const size_t per_outer = (size_t)ch * inner;
  const size_t total = (size_t)outer * inner;
  const float* base = w + (size_t)c * inner;
  const float levels = (float)((1 << n_bits) - 1);
  __shared__ float smn[8], smx[8];
  __shared__ double red[8];

  if (method == B200LIC_SCALE_GAUSSIAN) {
    double s = 0.0;
    for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
      const size_t o = i / inner, k = i - o * inner;
      s += (double)__ldg(base + o * per_outer + k);
    }
    const double mean = block_sum_f64(s, red) / (double)total;
    double q = 0.0;
    for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
      const size_t o = i / inner, k = i - o * inner;
      const double e = (double)__ldg(base + o * per_outer + k) - mean;
      q += e * e;
    }
    const double var = block_sum_f64(q, red) / (double)(total - 1);
    if (threadIdx.x == 0) {
      const float mu = (float)mean, six = __fmul_rn(6.f, (float)var);
      float lo = fminf(__fsub_rn(mu, six), 0.f), hi = fmaxf(__fadd_rn(mu, six), 0.f);
      if (symmetric) {
        const float a = fmaxf(fabsf(lo), hi);
        lo = lo < 0.f ? -a : 0.f;
        hi = a;
      }
      const float d = fmaxf(__fdiv_rn(__fsub_rn(hi, lo), levels), 1e-8f);
      delta[c] = d;
      zp[c] = rintf(__fdiv_rn(-lo, d));
    }
    return;
  }

  float mn = INFINITY, mx = -INFINITY;
  for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
    const size_t o = i / inner, k = i - o * inner;
    const float v = __ldg(base + o * per_outer + k);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) {
    smn[threadIdx.x >> 5] = mn;
    smx[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  mn = smn[0];
  mx = smx[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
    mn = fminf(mn, smn[i]);
    mx = fmaxf(mx, smx[i]);
  }
  const float top = (float)((1 << n_bits) - 1);
  double best = 1e10;
  float best_d = 0.f, best_z = 0.f;
  bool found = false;
  for (int s0 = 0; s0 < n_steps; s0 += kSearchChunk) {
    CandRange cr[kSearchChunk];
    double acc[kSearchChunk];
#pragma unroll
    for (int j = 0; j < kSearchChunk; ++j) {
      // `1.0 - i * shrink` is a Python float (fp64) cast to fp32 when it meets the fp32 tensor
      cr[j] = cand_range(mx, mn, (float)(1.0 - (double)(s0 + j) * shrink), levels);
      acc[j] = 0.0;
    }
    for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
      const size_t o = i / inner, k = i - o * inner;
      const float v = __ldg(base + o * per_outer + k);
#pragma unroll
      for (int j = 0; j < kSearchChunk; ++j) {
        float q = __fadd_rn(rintf(__fdiv_rn(v, cr[j].d)), cr[j].z);
        q = fminf(fmaxf(q, 0.f), top);
        const double e = (double)fabsf(__fsub_rn(v, __fmul_rn(__fsub_rn(q, cr[j].z), cr[j].d)));
        if (method == B200LIC_SCALE_L1) acc[j] += e;
        else if (method == B200LIC_SCALE_L2 || p == 2.f) acc[j] += e * e;
        else if (p == 3.5f) acc[j] += e * e * e * sqrt(e);
        else acc[j] += pow(e, (double)p);
      }
    }
#pragma unroll
    for (int j = 0; j < kSearchChunk; ++j) {
      const double score = block_sum_f64(acc[j], red) / (double)total;
      if (s0 + j < n_steps && score < best) {
        best = score;
        best_d = cr[j].d;
        best_z = cr[j].z;
        found = true;
      }
    }
  }
  if (threadIdx.x == 0) {
    // the reference leaves delta unset (and fails) when no candidate scores below 1e10; report NaN instead
    delta[c] = found ? best_d : __int_as_float(0x7fc00000);
    zp[c] = found ? best_z : __int_as_float(0x7fc00000);
  }
}

// ---- K7b: fake-quant -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wq_fake_quant_kernel(const float* __restrict__ w, const float* __restrict__ delta,
                                                             const float* __restrict__ zp, size_t n, int ch, int inner,
                                                             float top, float* __restrict__ w_dq,
                                                             float* __restrict__ codes, uint8_t* __restrict__ codes_u8) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / inner) % ch);
    const float d = __ldg(delta + c), z = __ldg(zp + c);
    float q = __fadd_rn(rintf(__fdiv_rn(w[i], d)), z);
    q = fminf(fmaxf(q, 0.f), top);
    if (codes) codes[i] = q;
    if (codes_u8) codes_u8[i] = (uint8_t)q;
    if (w_dq) w_dq[i] = __fmul_rn(__fsub_rn(q, z), d);
  }
}

__global__ void __launch_bounds__(256) wq_dequant_u8_kernel(const uint8_t* __restrict__ q, const float* __restrict__ delta,
                                                             const float* __restrict__ zp, size_t n, int ch, int inner,
                                                             float* __restrict__ w_dq) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / inner) % ch);
    w_dq[i] = __fmul_rn(__fsub_rn((float)q[i], __ldg(zp + c)), __ldg(delta + c));
  }
}

// ---- K6 ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float a) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-a))); }

__global__ void __launch_bounds__(256) adaround_init_kernel(const float* __restrict__ w, const float* __restrict__ delta,
                                                             size_t n, int ch, int inner, float* __restrict__ alpha) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / inner) % ch);
    const float t = __fdiv_rn(w[i], __ldg(delta + c));
    const float rest = __fsub_rn(t, floorf(t));
    // -log((zeta-gamma)/(rest-gamma) - 1); scalar/tensor is reciprocal()*scalar in torch
    const float r = __fmul_rn(__frcp_rn(__fsub_rn(rest, kGamma)), kStretch);
    alpha[i] = -logf(__fsub_rn(r, 1.f));
  }
}

// float4 helpers: 4 consecutive elements share a quantisation channel when inner % 4 == 0
template <bool VEC>
struct Pack {
  float v[VEC ? 4 : 1];
};
template <bool VEC>
__device__ __forceinline__ Pack<VEC> ldp(const float* p, size_t i) {
  Pack<VEC> r;
  if (VEC) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
    r.v[0] = t.x; r.v[1] = t.y; r.v[VEC ? 2 : 0] = t.z; r.v[VEC ? 3 : 0] = t.w;
  } else {
    r.v[0] = __ldg(p + i);
  }
  return r;
}
template <bool VEC>
__device__ __forceinline__ void stp(float* p, size_t i, const Pack<VEC>& r) {
  if (VEC) *reinterpret_cast<float4*>(p + i) = make_float4(r.v[0], r.v[1], r.v[VEC ? 2 : 0], r.v[VEC ? 3 : 0]);
  else p[i] = r.v[0];
}

template <bool VEC>
__global__ void __launch_bounds__(256) adaround_fwd_kernel(const float* __restrict__ w, const float* __restrict__ alpha,
                                                            const float* __restrict__ delta, const float* __restrict__ zp,
                                                            size_t n, int ch, int inner, float top, int soft,
                                                            float* __restrict__ w_q, float* __restrict__ codes) {
  constexpr int E = VEC ? 4 : 1;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * E; i < n; i += (size_t)gridDim.x * blockDim.x * E) {
    const int c = (int)((i / inner) % ch);
    const float d = __ldg(delta + c), z = __ldg(zp + c);
    const Pack<VEC> wv = ldp<VEC>(w, i), av = ldp<VEC>(alpha, i);
    Pack<VEC> qv, dq;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float base = floorf(__fdiv_rn(wv.v[e], d));
      const float a = av.v[e];
      float up;
      if (soft) {
        const float s = __fadd_rn(__fmul_rn(sigmoidf_(a), kStretch), kGamma);
        up = fminf(fmaxf(s, 0.f), 1.f);
      } else {
        up = a >= 0.f ? 1.f : 0.f;
      }
      float q = __fadd_rn(__fadd_rn(base, up), z);
      q = fminf(fmaxf(q, 0.f), top);
      qv.v[e] = q;
      dq.v[e] = __fmul_rn(__fsub_rn(q, z), d);
    }
    if (codes) stp<VEC>(codes, i, qv);
    if (w_q) stp<VEC>(w_q, i, dq);
  }
}

struct AdamArgs {
  float lr_over_bc1, inv_sqrt_bc2, beta1, beta2, eps;
};

// d loss / d alpha of one element: STE masks of the clamp and of the hard sigmoid (autograd of quantizer.py:437-452) on
// the upstream dL/dWq, plus the gradient of the rounding regulariser (layer_opt.py:160-165); reg_acc accumulates its value.
__device__ __forceinline__ float adaround_grad(float w, float a, float d, float z, float top, float d_wq, float grad_scale,
                                               float reg_weight, float reg_b, float& reg_acc) {
  const float sg = sigmoidf_(a);
  const float s = __fadd_rn(__fmul_rn(sg, kStretch), kGamma);
  const float h = fminf(fmaxf(s, 0.f), 1.f);
  // autograd of clamp passes the gradient on the closed interval
  const float dh_da = (s >= 0.f && s <= 1.f) ? kStretch * sg * (1.f - sg) : 0.f;
  const float xi = __fadd_rn(__fadd_rn(floorf(__fdiv_rn(w, d)), h), z);
  float g = 0.f;
  if (xi >= 0.f && xi <= top) g = grad_scale * d_wq * d * dh_da;
  if (reg_b > 0.f) {
    const float u = fabsf(h - 0.5f) * 2.f;          // |2h-1|
    const float ub1 = powf(u, reg_b - 1.f);
    reg_acc += 1.f - ub1 * u;
    const float sgn = (h > 0.5f) ? 1.f : ((h < 0.5f) ? -1.f : 0.f);
    g += -reg_weight * reg_b * ub1 * 2.f * sgn * dh_da;
  }
  return g;
}
// torch.optim.Adam (layer_opt.py:254,307) on one element; bias corrections pre-folded into `ad`.
__device__ __forceinline__ void adam_step(float a, float g, const AdamArgs& ad, float& m, float& v, float& a_new) {
  const float mi = ad.beta1 * m + (1.f - ad.beta1) * g;
  const float vi = ad.beta2 * v + (1.f - ad.beta2) * g * g;
  m = mi;
  v = vi;
  const float denom = sqrtf(vi) * ad.inv_sqrt_bc2 + ad.eps;
  a_new = a - ad.lr_over_bc1 * (mi / denom);
}

template <bool VEC>
__global__ void __launch_bounds__(256)
    adaround_bwd_adam_kernel(const float* __restrict__ w, float* __restrict__ alpha, const float* __restrict__ delta,
                             const float* __restrict__ zp, const float* __restrict__ d_wq, float* __restrict__ m,
                             float* __restrict__ v, size_t n, int ch, int inner, float top, AdamArgs ad,
                             float grad_scale, float reg_weight, float reg_b, float* __restrict__ reg_loss,
                             float* __restrict__ d_alpha_out, const b200lic_calib_sched* __restrict__ sched) {
  __shared__ float red[32];
  constexpr int E = VEC ? 4 : 1;
  float reg_acc = 0.f;
  if (sched != nullptr) {  // iteration-dependent scalars come from device memory (CUDA-graph replay)
    ad.lr_over_bc1 = __ldg(&sched->lr_over_bc1);
    ad.inv_sqrt_bc2 = __ldg(&sched->inv_sqrt_bc2);
    reg_b = __ldg(&sched->reg_b);
  }
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * E; i < n; i += (size_t)gridDim.x * blockDim.x * E) {
    const int c = (int)((i / inner) % ch);
    const float d = __ldg(delta + c), z = __ldg(zp + c);
    const Pack<VEC> wv = ldp<VEC>(w, i), av = ldp<VEC>(alpha, i), gv = ldp<VEC>(d_wq, i);
    Pack<VEC> mv, vv, ga, an;
    if (m != nullptr) {
      mv = ldp<VEC>(m, i);
      vv = ldp<VEC>(v, i);
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float g = adaround_grad(wv.v[e], av.v[e], d, z, top, gv.v[e], grad_scale, reg_weight, reg_b, reg_acc);
      ga.v[e] = g;
      if (m != nullptr) adam_step(av.v[e], g, ad, mv.v[e], vv.v[e], an.v[e]);
    }
    if (d_alpha_out) stp<VEC>(d_alpha_out, i, ga);
    if (m != nullptr) {  // else: gradient-only mode (autograd surface); no optimiser state touched
      stp<VEC>(m, i, mv);
      stp<VEC>(v, i, vv);
      stp<VEC>(alpha, i, an);
    }
  }
  if (reg_loss != nullptr && reg_b > 0.f) {
    const float tot = block_sum(reg_acc, red);
    if (threadIdx.x == 0) atomicAdd(reg_loss, reg_weight * tot);
  }
}

// ---- fused weight-gradient tail: split-K slab sum -> STE masks -> regulariser -> Adam --------------------------------------
// part[split][tap][cs][cb] are the per-split partial weight gradients of tc_wgrad_kernel (conv_tc_wgrad.cu).  One CTA owns
// 32 consecutive `cb` of one `cs`: phase 1 reads the slabs in their own order (128 B rows) and sums the splits in the
// fixed order of wgrad_reduce_kernel, phase 2 walks the same 32*T weights in WEIGHT order ([cs][cb][tap], contiguous), so
// w / alpha / m / v move in full lines.  Bit-identical to wgrad_reduce_kernel + adaround_bwd_adam_kernel.
constexpr int kTailCb = 32;
__global__ void __launch_bounds__(256)
    wgrad_reduce_adam_kernel(const float* __restrict__ part, int splits, int T, int Cs, int Cb, WgTail t, AdamArgs ad) {
  pdl_wait();                 // programmatic dependent launch (common.cuh): predecessor complete, memory visible
  extern __shared__ float tile[];                 // [kTailCb][T]
  __shared__ float red[32];
  const int cbb = (Cb + kTailCb - 1) / kTailCb;
  const int cs = blockIdx.x / cbb, cb0 = (blockIdx.x % cbb) * kTailCb;
  const int ncb = min(kTailCb, Cb - cb0);
  const size_t per = (size_t)T * Cs * Cb;
  float reg_b = 0.f;
  if (t.sched != nullptr) {
    ad.lr_over_bc1 = __ldg(&t.sched->lr_over_bc1);
    ad.inv_sqrt_bc2 = __ldg(&t.sched->inv_sqrt_bc2);
    reg_b = __ldg(&t.sched->reg_b);
  }
  // Phase 1: one item = 4 consecutive cb of one tap (a float4 of every slab).  All the slabs' loads of an item are
  // issued before the first add (groups of 8 independent 16-byte loads), so a thread pays two or three L2 round trips
  // instead of one per split; the additions keep wgrad_reduce_kernel's order: lane z%4 accumulates slab z, then
  // (a0 + a1) + (a2 + a3).
  const bool vec = (Cb & 3) == 0 && (((uintptr_t)part) & 15) == 0;
  if (vec) {
    for (int e = threadIdx.x; e < (kTailCb / 4) * T; e += blockDim.x) {
      const int tap = e / (kTailCb / 4), c4 = (e - tap * (kTailCb / 4)) * 4;
      float4 a[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c4 < ncb) {                               // ncb is a multiple of 4 here
        const float* p = part + ((size_t)tap * Cs + cs) * Cb + cb0 + c4;
        int z = 0;
        for (; z + 7 < splits; z += 8) {
          float4 v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(p + (size_t)(z + k) * per));
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            a[k & 3].x += v[k].x; a[k & 3].y += v[k].y; a[k & 3].z += v[k].z; a[k & 3].w += v[k].w;
          }
        }
        for (; z + 3 < splits; z += 4) {
          float4 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(p + (size_t)(z + k) * per));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            a[k].x += v[k].x; a[k].y += v[k].y; a[k].z += v[k].z; a[k].w += v[k].w;
          }
        }
        {
          float4 v[3];
          const int rem = splits - z;
#pragma unroll
          for (int k = 0; k < 3; ++k)
            v[k] = k < rem ? __ldg(reinterpret_cast<const float4*>(p + (size_t)(z + k) * per)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 3; ++k)
            if (k < rem) { a[0].x += v[k].x; a[0].y += v[k].y; a[0].z += v[k].z; a[0].w += v[k].w; }
        }
      }
      tile[(c4 + 0) * T + tap] = (a[0].x + a[1].x) + (a[2].x + a[3].x);
      tile[(c4 + 1) * T + tap] = (a[0].y + a[1].y) + (a[2].y + a[3].y);
      tile[(c4 + 2) * T + tap] = (a[0].z + a[1].z) + (a[2].z + a[3].z);
      tile[(c4 + 3) * T + tap] = (a[0].w + a[1].w) + (a[2].w + a[3].w);
    }
  } else {
    for (int e = threadIdx.x; e < kTailCb * T; e += blockDim.x) {
      const int tap = e / kTailCb, cbl = e - tap * kTailCb;
      float acc = 0.f;
      if (cbl < ncb) {
        const float* p = part + ((size_t)tap * Cs + cs) * Cb + cb0 + cbl;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int z = 0;
        for (; z + 3 < splits; z += 4) {
          a0 += __ldg(p + (size_t)z * per);
          a1 += __ldg(p + (size_t)(z + 1) * per);
          a2 += __ldg(p + (size_t)(z + 2) * per);
          a3 += __ldg(p + (size_t)(z + 3) * per);
        }
        for (; z < splits; ++z) a0 += __ldg(p + (size_t)z * per);
        acc = (a0 + a1) + (a2 + a3);
      }
      tile[cbl * T + tap] = acc;
    }
  }
  __syncthreads();
  // Phase 2: the CTA's ncb * T weights are contiguous in weight order and start 16-byte aligned (cb0 % 32 == 0): four per
  // thread, one 16-byte load / store per array, so the whole CTA is a single pass.  Channel index in 32-bit arithmetic.
  float reg_acc = 0.f;
  const unsigned base = (unsigned)(((size_t)cs * Cb + cb0) * T);
  const unsigned uinner = (unsigned)t.inner, uch = (unsigned)t.ch;
  auto elem = [&](unsigned i, float dwq, float w_, float a, float& mi, float& vi, float& an) {
    const unsigned c = (i / uinner) % uch;
    const float d = __ldg(t.delta + c), z = __ldg(t.zp + c);
    const float g = adaround_grad(w_, a, d, z, t.top, dwq, t.grad_scale, t.reg_weight, reg_b, reg_acc);
    adam_step(a, g, ad, mi, vi, an);
  };
  const int n_el = ncb * T;
  if (vec && (n_el & 3) == 0 && ((((uintptr_t)t.w) | ((uintptr_t)t.alpha) | ((uintptr_t)t.m) | ((uintptr_t)t.v) |
                                   ((uintptr_t)t.dw_out)) & 15) == 0) {
    for (int e = threadIdx.x * 4; e < n_el; e += blockDim.x * 4) {
      const unsigned i = base + (unsigned)e;
      const float4 wv = __ldg(reinterpret_cast<const float4*>(t.w + i));
      const float4 av = *reinterpret_cast<const float4*>(t.alpha + i);
      float4 mv = *reinterpret_cast<const float4*>(t.m + i), vv = *reinterpret_cast<const float4*>(t.v + i), an;
      const float4 gq = make_float4(tile[e], tile[e + 1], tile[e + 2], tile[e + 3]);
      if (t.dw_out) *reinterpret_cast<float4*>(t.dw_out + i) = gq;
      elem(i, gq.x, wv.x, av.x, mv.x, vv.x, an.x);
      elem(i + 1, gq.y, wv.y, av.y, mv.y, vv.y, an.y);
      elem(i + 2, gq.z, wv.z, av.z, mv.z, vv.z, an.z);
      elem(i + 3, gq.w, wv.w, av.w, mv.w, vv.w, an.w);
      *reinterpret_cast<float4*>(t.m + i) = mv;
      *reinterpret_cast<float4*>(t.v + i) = vv;
      *reinterpret_cast<float4*>(t.alpha + i) = an;
    }
  } else {
    for (int e = threadIdx.x; e < n_el; e += blockDim.x) {
      const unsigned i = base + (unsigned)e;
      const float dwq = tile[e];
      if (t.dw_out) t.dw_out[i] = dwq;
      float mi = t.m[i], vi = t.v[i], an;
      elem(i, dwq, __ldg(t.w + i), t.alpha[i], mi, vi, an);
      t.m[i] = mi;
      t.v[i] = vi;
      t.alpha[i] = an;
    }
  }
  if (t.reg_loss != nullptr && reg_b > 0.f) {
    const float tot = block_sum(reg_acc, red);
    if (threadIdx.x == 0) atomicAdd(t.reg_loss, t.reg_weight * tot);
  }
}

int launch_wgrad_reduce_adam(const float* part, int splits, int T, int Cs, int Cb, const WgTail* tail, cudaStream_t s) {
  AdamArgs ad{0.f, 0.f, tail->beta1, tail->beta2, tail->eps};
  const int cbb = (Cb + kTailCb - 1) / kTailCb;
  const size_t smem = (size_t)kTailCb * T * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("wgrad_reduce_adam: %d taps exceed the tile", T);
    return B200LIC_ERR_UNSUPPORTED;
  }
  launch_pdl(wgrad_reduce_adam_kernel, dim3(Cs * cbb), dim3(256), smem, s, part, splits, T, Cs, Cb, *tail, ad);
  B200_LAUNCH_CHECK("wgrad_reduce_adam_kernel");
  return B200LIC_OK;
}

// ---- multi-GPU tail: cross-GPU gradient reduction + Adam over NVLink peer memory, one kernel ------------------------------
// Data-parallel calibration all-reduces dL/dWq and then applies the same Adam step on every rank.  Here the
// optimiser state is sharded instead: rank r owns elements [lo, hi) of the layer.  It sums the ranks' local gradients of
// its shard straight out of their memory (peer loads over NVLink / NVSwitch, fixed rank order), applies the STE masks,
// the regulariser and Adam with its shard of m / v, and stores the new alpha into EVERY rank's alpha buffer (peer
// stores): the traffic of a ring all-reduce, but no collective launch, no second pass for Adam, 1/N of the Adam work per
// rank, and alpha is bit-identical everywhere by construction (each element is computed once).
// Two cross-GPU barriers on flags in the symmetric buffers: A (entry) -- every rank's gradient buffer is complete (the
// kernels that wrote it precede this one in stream order); B (exit, last CTA only) -- every rank's alpha stores have
// landed here before the kernel completes, so stream order protects the next reader of alpha.
// exit_barrier = 0 drops B: valid when, on every rank, another call of this kernel (any layer: its barrier A is a
// system-scope release / acquire over all ranks, cumulative over the stores fenced here) precedes in stream order the
// next reader of this layer's alpha and the next writer of its gradient buffer -- a sequential sweep over >= 2 units.
struct XgpuArgs {
  const float* const* grad_ptrs;    // [world] peer pointers: each rank's local dL/dWq of this layer
  float* const* alpha_ptrs;         // [world] peer pointers: each rank's alpha of this layer
  unsigned* const* flag_ptrs;       // [world] peer pointers: each rank's flag block, 2 * world words (A row, B row)
  unsigned* state;                  // local: [0] epoch, [1] CTAs finished
  int rank, world;
  size_t lo, hi;                    // this rank's shard
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// spin until every rank's flag in `row` of MY flag block has reached `epoch`; gives up after ~4 s (a peer died) rather
// than hanging the GPU
__device__ __forceinline__ void xgpu_wait(const unsigned* my_flags, int row, int world, unsigned epoch) {
  for (int p = 0; p < world; ++p) {
    long long spins = 0;
    while ((int)(ld_acquire_sys(my_flags + row * world + p) - epoch) < 0) {
      __nanosleep(100);
      if (++spins > 40000000LL) return;
    }
  }
}

__global__ void __launch_bounds__(256)
    xgpu_reduce_adam_kernel(XgpuArgs x, const float* __restrict__ w, const float* __restrict__ delta,
                            const float* __restrict__ zp, float* __restrict__ m, float* __restrict__ v, int ch, int inner,
                            float top, const b200lic_calib_sched* __restrict__ sched, AdamArgs ad, float grad_scale,
                            float reg_weight, float* __restrict__ reg_loss, int exit_barrier) {
  __shared__ float red[32];
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(x.state) + 1u;
  // ---- barrier A: tell everyone my gradient is ready, wait until everyone's is
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0)
      for (int p = 0; p < x.world; ++p) st_release_sys(x.flag_ptrs[p] + x.rank, epoch);
    xgpu_wait(x.flag_ptrs[x.rank], 0, x.world, epoch);
  }
  __syncthreads();
  ad.lr_over_bc1 = __ldg(&sched->lr_over_bc1);
  ad.inv_sqrt_bc2 = __ldg(&sched->inv_sqrt_bc2);
  const float reg_b = __ldg(&sched->reg_b);
  float reg_acc = 0.f;
  const unsigned uinner = (unsigned)inner, uch = (unsigned)ch;
  const size_t n4 = (x.hi - x.lo) >> 2;                       // shards are multiples of four elements, 16-byte aligned
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += (size_t)gridDim.x * blockDim.x) {
    const size_t i = x.lo + (k << 2);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < x.world; ++p) {                        // fixed rank order; volatile: peer memory, never cached here
      float4 t;
      asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                   : "l"(x.grad_ptrs[p] + i)
                   : "memory");
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + i));
    const float4 av = *reinterpret_cast<const float4*>(x.alpha_ptrs[x.rank] + i);
    float4 mv = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i), an;
    auto elem = [&](unsigned e, float dwq, float w_, float a, float& mi, float& vi, float& out) {
      const unsigned c = (e / uinner) % uch;
      const float d = __ldg(delta + c), z = __ldg(zp + c);
      const float gg = adaround_grad(w_, a, d, z, top, dwq, grad_scale, reg_weight, reg_b, reg_acc);
      adam_step(a, gg, ad, mi, vi, out);
    };
    elem((unsigned)i, g.x, wv.x, av.x, mv.x, vv.x, an.x);
    elem((unsigned)i + 1, g.y, wv.y, av.y, mv.y, vv.y, an.y);
    elem((unsigned)i + 2, g.z, wv.z, av.z, mv.z, vv.z, an.z);
    elem((unsigned)i + 3, g.w, wv.w, av.w, mv.w, vv.w, an.w);
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
    for (int p = 0; p < x.world; ++p) *reinterpret_cast<float4*>(x.alpha_ptrs[p] + i) = an;
  }
  if (reg_loss != nullptr && reg_b > 0.f) {
    const float tot = block_sum(reg_acc, red);
    if (threadIdx.x == 0) atomicAdd(reg_loss, reg_weight * tot * (float)x.world);   // this rank's shard, scaled up
  }
  // ---- barrier B: last CTA of the grid signals "my alpha stores are done" and waits for everyone's
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(x.state + 1, 1u);
    if (done == gridDim.x - 1) {
      __threadfence_system();
      if (exit_barrier) {
        for (int p = 0; p < x.world; ++p) st_release_sys(x.flag_ptrs[p] + x.world + x.rank, epoch);
        xgpu_wait(x.flag_ptrs[x.rank], 1, x.world, epoch);
      }
      x.state[1] = 0u;
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(x.state) = epoch;
    }
  }
}

// ---- fused weight quantiser -> packed tensor-core operand ----------------------------------------------------------------
// For every element of the packed operand [phase][co][tap][ci] (conv_tc.cu pack_weights_kernel's layout): fetch the
// source weight, fake-quantise it (nearest: quantizer.py:175-177; AdaRound soft / hard: :437-449 -- the arithmetic of
// wq_fake_quant_kernel / adaround_fwd_kernel above, bit for bit) and store its bf16 hi / lo split.  mode 1 stores the
// integer n = code - zero_point instead (exact in bf16, hi slab only): the operand of the two-pass forward.
__global__ void __launch_bounds__(256)
    quant_pack_kernel(PackDst g, const float* __restrict__ w, const float* __restrict__ alpha,
                      const float* __restrict__ delta, const float* __restrict__ zp, int ch, int inner, float top, int soft,
                      int mode, __nv_bfloat16* __restrict__ bh, __nv_bfloat16* __restrict__ bl, float* __restrict__ w_q) {
  const int phase = blockIdx.z;
  int r0 = 0, s0 = 0, KHp = g.KH, KWp = g.KW, rstep = 1;
  if (g.transposed) {
    const int st = g.stride, ph = phase / st, pw = phase % st;
    r0 = (ph + g.pad) % st;
    s0 = (pw + g.pad) % st;
    KHp = r0 < g.KH ? (g.KH - r0 + st - 1) / st : 0;
    KWp = s0 < g.KW ? (g.KW - s0 + st - 1) / st : 0;
    rstep = st;
  }
  const int T = KHp * KWp;
  const size_t Kmax = (size_t)g.Tmax * g.Cpad;
  const size_t per_phase = (size_t)g.CoutPad * Kmax;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < per_phase; e += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(e / Kmax);
    const int k = (int)(e - (size_t)co * Kmax);
    const int t = k / g.Cpad, ci = k - t * g.Cpad;
    float val = 0.f;
    if (co < g.Cout && ci < g.Cin && t < T) {
      const int i = t / KWp, j = t - i * KWp;
      const size_t src = (size_t)((long long)co * g.s_co + (long long)ci * g.s_ci + (r0 + i * rstep) * g.KW + (s0 + j * rstep));
      const int c = (int)((src / inner) % ch);
      const float d = __ldg(delta + c), z = __ldg(zp + c);
      const float tq = __fdiv_rn(__ldg(w + src), d);
      float q;
      if (alpha == nullptr) {
        q = __fadd_rn(rintf(tq), z);
      } else {
        const float a = __ldg(alpha + src);
        float up;
        if (soft) up = fminf(fmaxf(__fadd_rn(__fmul_rn(sigmoidf_(a), kStretch), kGamma), 0.f), 1.f);
        else up = a >= 0.f ? 1.f : 0.f;
        q = __fadd_rn(__fadd_rn(floorf(tq), up), z);
      }
      q = fminf(fmaxf(q, 0.f), top);
      const float n_ = __fsub_rn(q, z);
      val = mode ? n_ : __fmul_rn(n_, d);
      if (w_q) w_q[src] = val;
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(val);
    bh[(size_t)phase * per_phase + e] = h;
    if (!mode) bl[(size_t)phase * per_phase + e] = __float2bfloat16_rn(val - __bfloat162float(h));
  }
}

// The same, one CTA per (output channel, 32 input channels) with that slice staged in shared memory: the source is read in
// its own order (conv: one contiguous 32*KH*KW run; transposed conv: 32 runs of KH*KW) and the packed rows [tap][ci] are
// written in 64-byte runs -- the element-wise kernel above walks the PACKED index and so reads w / alpha with a stride
// of KH*KW floats (one useful word per 32-byte sector).
__global__ void __launch_bounds__(256)
    quant_pack_tile_kernel(PackDst g, const float* __restrict__ w, const float* __restrict__ alpha,
                           const float* __restrict__ delta, const float* __restrict__ zp, int ch, int inner, float top,
                           int soft, int mode, __nv_bfloat16* __restrict__ bh, __nv_bfloat16* __restrict__ bl,
                           float* __restrict__ w_q) {
  pdl_wait();                 // programmatic dependent launch (common.cuh): predecessor complete, memory visible
  extern __shared__ float wt[];                     // [KH*KW][32]: one output channel x 32 input channels
  // Index arithmetic once per CTA instead of once per element (the first version spent 200 instructions per weight, most
  // of them integer divisions by run-time KH*KW / inner / Tmax): the quantisation channel, delta and zero point of each
  // of the 32 input channels (a (co, ci) filter never straddles a channel of the quantiser: `inner` is a multiple of
  // KH*KW), and the source tap of every (phase, tap slot) of the packed rows.
  __shared__ float s_d[32], s_z[32];
  __shared__ unsigned s_src[32];
  __shared__ short s_rs[4 * 64];                    // [phase][t] -> r*KW + s, or -1 (phases <= 4... checked by the launcher)
  const int co = blockIdx.x, ci0 = blockIdx.y * 32;
  const int KK = g.KH * g.KW;
  const int nci = min(32, g.Cin - ci0);             // <= 0 for a chunk of pure padding channels
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 32) {
    const int cl = threadIdx.x;
    if (cl < nci) {
      const unsigned src = (unsigned)((long long)co * g.s_co + (long long)(ci0 + cl) * g.s_ci);
      const unsigned c = (src / (unsigned)inner) % (unsigned)ch;
      s_src[cl] = src;
      s_d[cl] = __ldg(delta + c);
      s_z[cl] = __ldg(zp + c);
    }
  }
  const int slots = g.phases * g.Tmax;
  for (int i = threadIdx.x; i < slots; i += blockDim.x) {
    const int phase = i / g.Tmax, t = i - phase * g.Tmax;
    int r0 = 0, s0 = 0, KHp = g.KH, KWp = g.KW, rstep = 1;
    if (g.transposed) {
      const int st = g.stride, ph = phase / st, pw = phase % st;
      r0 = (ph + g.pad) % st;
      s0 = (pw + g.pad) % st;
      KHp = r0 < g.KH ? (g.KH - r0 + st - 1) / st : 0;
      KWp = s0 < g.KW ? (g.KW - s0 + st - 1) / st : 0;
      rstep = st;
    }
    short rs = -1;
    if (t < KHp * KWp) {
      const int ii = t / KWp, jj = t - ii * KWp;
      rs = (short)((r0 + ii * rstep) * g.KW + (s0 + jj * rstep));
    }
    s_rs[i] = rs;
  }
  __syncthreads();
  for (int cl = warp; cl < nci; cl += 8) {          // one input channel per warp and round, lanes along the taps
    const float d = s_d[cl], z = s_z[cl];
    const unsigned base = s_src[cl];
    for (int rs = lane; rs < KK; rs += 32) {
      const unsigned src = base + (unsigned)rs;
      const float tq = __fdiv_rn(__ldg(w + src), d);
      float q;
      if (alpha == nullptr) {
        q = __fadd_rn(rintf(tq), z);
      } else {
        const float a = __ldg(alpha + src);
        float up;
        if (soft) up = fminf(fmaxf(__fadd_rn(__fmul_rn(sigmoidf_(a), kStretch), kGamma), 0.f), 1.f);
        else up = a >= 0.f ? 1.f : 0.f;
        q = __fadd_rn(__fadd_rn(floorf(tq), up), z);
      }
      q = fminf(fmaxf(q, 0.f), top);
      const float n_ = __fsub_rn(q, z);
      const float val = mode ? n_ : __fmul_rn(n_, d);
      if (w_q) w_q[src] = val;
      wt[rs * 32 + cl] = val;
    }
  }
  __syncthreads();
  const size_t Kmax = (size_t)g.Tmax * g.Cpad;
  const size_t per_phase = (size_t)g.CoutPad * Kmax;
  // items: (slot = phase * Tmax + t, pair of input channels): 64-byte runs of the packed row
  for (int it = threadIdx.x; it < slots * 16; it += blockDim.x) {
    const int c2 = (it & 15) * 2, slot = it >> 4;
    const int rs = s_rs[slot];
    float v0 = 0.f, v1 = 0.f;
    if (rs >= 0) {
      if (c2 < nci) v0 = wt[rs * 32 + c2];
      if (c2 + 1 < nci) v1 = wt[rs * 32 + c2 + 1];
    }
    // slot = phase * Tmax + t and per_phase = CoutPad * Tmax * Cpad: phase * per_phase + t * Cpad without a division
    const int phase = g.phases == 1 ? 0 : slot / g.Tmax;
    const int t = slot - phase * g.Tmax;
    const size_t o = (size_t)phase * per_phase + (size_t)co * Kmax + (size_t)t * g.Cpad + ci0 + c2;
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
    __nv_bfloat162 hv;
    hv.x = h0;
    hv.y = h1;
    *reinterpret_cast<__nv_bfloat162*>(bh + o) = hv;
    if (!mode) {
      __nv_bfloat162 lv;
      lv.x = __float2bfloat16_rn(v0 - __bfloat162float(h0));
      lv.y = __float2bfloat16_rn(v1 - __bfloat162float(h1));
      *reinterpret_cast<__nv_bfloat162*>(bl + o) = lv;
    }
  }
}

// zero rows of the padded output channels (co >= Cout) of a packed operand
__global__ void __launch_bounds__(256) pack_pad_rows_kernel(PackDst g, int mode, __nv_bfloat16* __restrict__ bh,
                                                             __nv_bfloat16* __restrict__ bl) {
  const size_t Kmax = (size_t)g.Tmax * g.Cpad;
  const size_t per_phase = (size_t)g.CoutPad * Kmax, pad = (size_t)(g.CoutPad - g.Cout) * Kmax;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < pad * g.phases; e += (size_t)gridDim.x * blockDim.x) {
    const size_t phase = e / pad, r = e - phase * pad;
    const size_t o = phase * per_phase + (size_t)g.Cout * Kmax + r;
    bh[o] = __float2bfloat16_rn(0.f);
    if (!mode) bl[o] = __float2bfloat16_rn(0.f);
  }
}

int launch_quant_pack(const PackDst& g, const float* w, const float* alpha, const float* delta, const float* zp, int ch,
                      int inner, int n_levels, int soft, int mode, void* packed, float* w_q, cudaStream_t s) {
  const size_t per_phase = (size_t)g.CoutPad * g.Tmax * g.Cpad;
  __nv_bfloat16* bh = reinterpret_cast<__nv_bfloat16*>(packed);
  __nv_bfloat16* bl = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(packed) + g.b_bytes);
  const size_t tile_bytes = (size_t)32 * g.KH * g.KW * sizeof(float);
  static const bool elementwise_only = getenv("B200LIC_QPACK_ELEMENTWISE") != nullptr;     // A/B experiments
  if (tile_bytes <= 40 * 1024 && g.Cpad / 32 <= 65535 && g.phases * g.Tmax <= 256 && (inner % (g.KH * g.KW)) == 0 &&
      !elementwise_only) {
    dim3 tgrid((unsigned)g.Cout, (unsigned)(g.Cpad / 32), 1);
    launch_pdl(quant_pack_tile_kernel, tgrid, dim3(256), tile_bytes, s, g, w, alpha, delta, zp, ch, inner, (float)(n_levels - 1), soft,
                                                          mode, bh, bl, w_q);
    B200_LAUNCH_CHECK("quant_pack_tile_kernel");
    if (g.CoutPad > g.Cout) {
      const size_t pad = (size_t)(g.CoutPad - g.Cout) * g.Tmax * g.Cpad * g.phases;
      pack_pad_rows_kernel<<<grid_for(pad, 256), 256, 0, s>>>(g, mode, bh, bl);
      B200_LAUNCH_CHECK("pack_pad_rows_kernel");
    }
    return B200LIC_OK;
  }
  dim3 grid((unsigned)((per_phase + 255) / 256 > 1184 ? 1184 : (per_phase + 255) / 256), 1, g.phases);
  quant_pack_kernel<<<grid, 256, 0, s>>>(g, w, alpha, delta, zp, ch, inner, (float)(n_levels - 1), soft, mode, bh, bl, w_q);
  B200_LAUNCH_CHECK("quant_pack_kernel");
  return B200LIC_OK;
}

// ---- integer weights n = code - zero_point (the operand of the two-pass forward, b200lic_conv_fwd_wq) ----------------
__global__ void __launch_bounds__(256) wq_int_weights_kernel(const float* __restrict__ w, const float* __restrict__ alpha,
                                                              const float* __restrict__ delta, const float* __restrict__ zp,
                                                              size_t n, int ch, int inner, float top,
                                                              float* __restrict__ w_int) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / inner) % ch);
    const float d = __ldg(delta + c), z = __ldg(zp + c);
    const float t = __fdiv_rn(w[i], d);
    float q;
    if (alpha == nullptr) q = __fadd_rn(rintf(t), z);                                     // quantizer.py:175
    else q = __fadd_rn(__fadd_rn(floorf(t), __ldg(alpha + i) >= 0.f ? 1.f : 0.f), z);     // quantizer.py:437-449, hard
    q = fminf(fmaxf(q, 0.f), top);
    w_int[i] = __fsub_rn(q, z);
  }
}

// ---- learned step size (LSQ) ------------------------------------------------------------------------------
// d loss / d delta[c] by the autograd of the fake-quant expressions with delta as the leaf (the reference keeps this
// option as commented-out code: quantizer.py:166-168, layer_opt.py:259-265, block_opt.py:254-266):
//   nearest (quantizer.py:175-177, x_int = round_ste(w/d) + zp):   d out/d d = (x_q - zp) - [0 <= x_int <= top] * w/d
//   AdaRound (quantizer.py:437-449, x_int = floor(w/d) + h + zp):   d out/d d = (x_q - zp)        (floor: zero gradient)
// One CTA per quantisation channel, fixed-order block reduction (deterministic), optional fused Adam step on delta.
__global__ void __launch_bounds__(256)
    lsq_delta_grad_kernel(const float* __restrict__ w, const float* __restrict__ alpha, float* __restrict__ delta,
                          const float* __restrict__ zp, const float* __restrict__ d_wq, int outer, int ch, int inner,
                          float top, int soft, float grad_scale, float* __restrict__ d_delta, float* __restrict__ m,
                          float* __restrict__ v, AdamArgs ad, const b200lic_calib_sched* __restrict__ sched,
                          float lr_scale) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  if (sched != nullptr) {  // CUDA-graph replay: bias corrections of this iteration come from device memory
    ad.lr_over_bc1 = lr_scale * __ldg(&sched->lr_over_bc1);
    ad.inv_sqrt_bc2 = __ldg(&sched->inv_sqrt_bc2);
  }
  const float d = delta[c], z = __ldg(zp + c);
  const size_t per_outer = (size_t)ch * inner, total = (size_t)outer * inner;
  float acc = 0.f;
  for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
    const size_t o = i / inner, k = i - o * inner;
    const size_t e = o * per_outer + (size_t)c * inner + k;
    const float t = __fdiv_rn(__ldg(w + e), d);
    float term;
    if (alpha == nullptr) {
      const float xi = __fadd_rn(rintf(t), z);
      const float xq = fminf(fmaxf(xi, 0.f), top);
      term = __fsub_rn(xq, z);
      if (xi >= 0.f && xi <= top) term = __fsub_rn(term, t);
    } else {
      const float a = __ldg(alpha + e);
      float up;
      if (soft) {
        const float sg = __fadd_rn(__fmul_rn(sigmoidf_(a), kStretch), kGamma);
        up = fminf(fmaxf(sg, 0.f), 1.f);
      } else {
        up = a >= 0.f ? 1.f : 0.f;
      }
      const float xq = fminf(fmaxf(__fadd_rn(__fadd_rn(floorf(t), up), z), 0.f), top);
      term = __fsub_rn(xq, z);
    }
    acc += __ldg(d_wq + e) * term;
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float g = grad_scale * tot;
    if (d_delta) d_delta[c] = g;
    if (m != nullptr) {
      const float mi = ad.beta1 * m[c] + (1.f - ad.beta1) * g;
      const float vi = ad.beta2 * v[c] + (1.f - ad.beta2) * g * g;
      m[c] = mi;
      v[c] = vi;
      delta[c] = fmaxf(d - ad.lr_over_bc1 * (mi / (sqrtf(vi) * ad.inv_sqrt_bc2 + ad.eps)), 1e-8f);
    }
  }
}

static inline bool vec4_ok(size_t n, int inner, std::initializer_list<const void*> ptrs) {
  if ((inner & 3) != 0 || (n & 3) != 0) return false;
  for (const void* p : ptrs)
    if (p && (((uintptr_t)p) & 15)) return false;
  return true;
}

// One thread advances the device-resident schedule (same arithmetic as the host path: fp64 bias corrections cast to
// fp32; LinearTempDecay of utils.py:37-54 evaluated in fp64 like Python does).
__global__ void calib_sched_tick_kernel(b200lic_calib_sched* s, int iters, double warmup, double b_start, double b_end,
                                        float lr, float beta1, float beta2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int step = s->step + 1;
  s->step = step;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  s->lr_over_bc1 = (float)((double)lr / bc1);
  s->inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const double start_decay = warmup * (double)iters;
  double b;
  if ((double)step < start_decay) {
    b = 0.0;                                   // layer_opt.py:156-158: no rounding loss during warm-up
  } else {
    const double rel = ((double)step - start_decay) / ((double)iters - start_decay);
    b = b_end + (b_start - b_end) * fmax(0.0, 1.0 - rel);
  }
  s->reg_b = (float)b;
}

}  // namespace b200lic

using namespace b200lic;

extern "C" {

int b200lic_wq_init_minmax(const float* w, int outer, int ch, int inner, int n_bits, int scale_variant,
                           int symmetric, float* delta, float* zero_point, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(w && delta && zero_point, "wq_init_minmax: null pointer");
  B200_REQUIRE(outer > 0 && ch > 0 && inner > 0, "wq_init_minmax: bad shape (%d,%d,%d)", outer, ch, inner);
  B200_REQUIRE(n_bits >= 2 && n_bits <= 16, "wq_init_minmax: n_bits=%d outside [2,16]", n_bits);
  wq_minmax_kernel<<<ch, 256, 0, as_stream(stream)>>>(w, outer, ch, inner, n_bits, scale_variant, symmetric, delta,
                                                      zero_point);
  B200_LAUNCH_CHECK("wq_minmax_kernel");
  return B200LIC_OK;
}

int b200lic_wq_init_search(const float* w, int outer, int ch, int inner, int n_bits, int method, int n_steps,
                           double shrink, float p, int symmetric, float* delta, float* zero_point,
                           b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(w && delta && zero_point, "wq_init_search: null pointer");
  B200_REQUIRE(outer > 0 && ch > 0 && inner > 0, "wq_init_search: bad shape (%d,%d,%d)", outer, ch, inner);
  B200_REQUIRE(n_bits >= 2 && n_bits <= 16, "wq_init_search: n_bits=%d outside [2,16]", n_bits);
  B200_REQUIRE(method >= B200LIC_SCALE_MSE && method <= B200LIC_SCALE_GAUSSIAN, "wq_init_search: method %d", method);
  B200_REQUIRE(method == B200LIC_SCALE_GAUSSIAN || (n_steps >= 1 && n_steps <= 1024 && shrink > 0.0 && p > 0.f),
               "wq_init_search: n_steps=%d shrink=%g p=%g", n_steps, shrink, (double)p);
  wq_search_kernel<<<ch, 256, 0, as_stream(stream)>>>(w, outer, ch, inner, n_bits, method, n_steps, shrink, p,
                                                      symmetric, delta, zero_point);
  B200_LAUNCH_CHECK("wq_search_kernel");
  return B200LIC_OK;
}

int b200lic_wq_fake_quant(const float* w, const float* delta, const float* zero_point, int outer, int ch, int inner,
                          int n_levels, float* w_dq, float* codes, uint8_t* codes_u8, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(w && delta && zero_point, "wq_fake_quant: null pointer");
  B200_REQUIRE(outer > 0 && ch > 0 && inner > 0 && n_levels >= 2, "wq_fake_quant: bad shape");
  B200_REQUIRE(!codes_u8 || n_levels <= 256, "wq_fake_quant: uint8 codes need n_levels <= 256");
  const size_t n = (size_t)outer * ch * inner;
  wq_fake_quant_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(w, delta, zero_point, n, ch, inner,
                                                                        (float)(n_levels - 1), w_dq, codes, codes_u8);
  B200_LAUNCH_CHECK("wq_fake_quant_kernel");
  return B200LIC_OK;
}

int b200lic_wq_dequant_u8(const uint8_t* codes_u8, const float* delta, const float* zero_point, int outer, int ch,
                          int inner, float* w_dq, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(codes_u8 && delta && zero_point && w_dq, "wq_dequant_u8: null pointer");
  const size_t n = (size_t)outer * ch * inner;
  wq_dequant_u8_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(codes_u8, delta, zero_point, n, ch, inner, w_dq);
  B200_LAUNCH_CHECK("wq_dequant_u8_kernel");
  return B200LIC_OK;
}

int b200lic_adaround_init_alpha(const float* w, const float* delta, int outer, int ch, int inner, float* alpha,
                                b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(w && delta && alpha, "adaround_init_alpha: null pointer");
  const size_t n = (size_t)outer * ch * inner;
  adaround_init_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(w, delta, n, ch, inner, alpha);
  B200_LAUNCH_CHECK("adaround_init_kernel");
  return B200LIC_OK;
}

int b200lic_adaround_fwd(const float* w, const float* alpha, const float* delta, const float* zero_point, int outer,
                         int ch, int inner, int n_levels, int soft, float* w_q, float* codes, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(w && alpha && delta && zero_point, "adaround_fwd: null pointer");
  const size_t n = (size_t)outer * ch * inner;
  if (vec4_ok(n, inner, {w, alpha, w_q, codes}))
    adaround_fwd_kernel<true><<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(
        w, alpha, delta, zero_point, n, ch, inner, (float)(n_levels - 1), soft, w_q, codes);
  else
    adaround_fwd_kernel<false><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
        w, alpha, delta, zero_point, n, ch, inner, (float)(n_levels - 1), soft, w_q, codes);
  B200_LAUNCH_CHECK("adaround_fwd_kernel");
  return B200LIC_OK;
}

int b200lic_adaround_bwd_adam(const float* w, float* alpha, const float* delta, const float* zero_point,
                              const float* d_wq, float* exp_avg, float* exp_avg_sq, int outer, int ch, int inner,
                              int n_levels, int step, float lr, float beta1, float beta2, float eps, float grad_scale,
                              float reg_weight, float reg_b, float* reg_loss, float* d_alpha_out,
                              b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(w && alpha && delta && zero_point && d_wq, "adaround_bwd_adam: null pointer");
  B200_REQUIRE((exp_avg && exp_avg_sq) || (!exp_avg && !exp_avg_sq && d_alpha_out),
               "adaround_bwd_adam: pass both Adam moments, or neither together with d_alpha_out (gradient-only)");
  B200_REQUIRE(step >= 1, "adaround_bwd_adam: step must be >= 1");
  const size_t n = (size_t)outer * ch * inner;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  AdamArgs ad{(float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2, eps};
  if (vec4_ok(n, inner, {w, alpha, d_wq, exp_avg, exp_avg_sq, d_alpha_out}))
    adaround_bwd_adam_kernel<true><<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(
        w, alpha, delta, zero_point, d_wq, exp_avg, exp_avg_sq, n, ch, inner, (float)(n_levels - 1), ad, grad_scale,
        reg_weight, reg_b, reg_loss, d_alpha_out, nullptr);
  else
    adaround_bwd_adam_kernel<false><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
        w, alpha, delta, zero_point, d_wq, exp_avg, exp_avg_sq, n, ch, inner, (float)(n_levels - 1), ad, grad_scale,
        reg_weight, reg_b, reg_loss, d_alpha_out, nullptr);
  B200_LAUNCH_CHECK("adaround_bwd_adam_kernel");
  return B200LIC_OK;
}

int b200lic_calib_sched_tick(b200lic_calib_sched* sched, int iters, double warmup, double b_start, double b_end, float lr,
                             float beta1, float beta2, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(sched, "calib_sched_tick: null pointer");
  B200_REQUIRE(iters >= 1 && warmup >= 0.0 && warmup < 1.0, "calib_sched_tick: bad schedule (iters=%d, warmup=%f)", iters,
               warmup);
  calib_sched_tick_kernel<<<1, 32, 0, as_stream(stream)>>>(sched, iters, warmup, b_start, b_end, lr, beta1, beta2);
  B200_LAUNCH_CHECK("calib_sched_tick_kernel");
  return B200LIC_OK;
}

int b200lic_adaround_bwd_adam_sched(const float* w, float* alpha, const float* delta, const float* zero_point,
                                    const float* d_wq, float* exp_avg, float* exp_avg_sq, int outer, int ch, int inner,
                                    int n_levels, const b200lic_calib_sched* sched, float beta1, float beta2, float eps,
                                    float grad_scale, float reg_weight, float* reg_loss, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(w && alpha && delta && zero_point && d_wq && exp_avg && exp_avg_sq && sched,
               "adaround_bwd_adam_sched: null pointer");
  const size_t n = (size_t)outer * ch * inner;
  AdamArgs ad{0.f, 0.f, beta1, beta2, eps};
  if (vec4_ok(n, inner, {w, alpha, d_wq, exp_avg, exp_avg_sq}))
    adaround_bwd_adam_kernel<true><<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(
        w, alpha, delta, zero_point, d_wq, exp_avg, exp_avg_sq, n, ch, inner, (float)(n_levels - 1), ad, grad_scale,
        reg_weight, 0.f, reg_loss, nullptr, sched);
  else
    adaround_bwd_adam_kernel<false><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
        w, alpha, delta, zero_point, d_wq, exp_avg, exp_avg_sq, n, ch, inner, (float)(n_levels - 1), ad, grad_scale,
        reg_weight, 0.f, reg_loss, nullptr, sched);
  B200_LAUNCH_CHECK("adaround_bwd_adam_kernel(sched)");
  return B200LIC_OK;
}

int b200lic_xgpu_reduce_adam_sched(const float* const* grad_ptrs, float* const* alpha_ptrs, unsigned* const* flag_ptrs,
                                   unsigned* state, int rank, int world, size_t shard_lo, size_t shard_hi, const float* w,
                                   const float* delta, const float* zero_point, float* exp_avg, float* exp_avg_sq,
                                   int outer, int ch, int inner, int n_levels, const b200lic_calib_sched* sched,
                                   float beta1, float beta2, float eps, float grad_scale, float reg_weight,
                                   float* reg_loss, int exit_barrier, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(grad_ptrs && alpha_ptrs && flag_ptrs && state && w && delta && zero_point && exp_avg && exp_avg_sq && sched,
               "xgpu_reduce_adam_sched: null pointer");
  B200_REQUIRE(world >= 2 && world <= 16 && rank >= 0 && rank < world, "xgpu_reduce_adam_sched: rank %d of %d", rank, world);
  const size_t n = (size_t)outer * ch * inner;
  B200_REQUIRE(shard_lo <= shard_hi && shard_hi <= n && (shard_lo & 3) == 0 && ((shard_hi - shard_lo) & 3) == 0,
               "xgpu_reduce_adam_sched: shard [%zu, %zu) must be 4-element aligned inside %zu", shard_lo, shard_hi, n);
  XgpuArgs x{grad_ptrs, alpha_ptrs, flag_ptrs, state, rank, world, shard_lo, shard_hi};
  AdamArgs ad{0.f, 0.f, beta1, beta2, eps};
  const size_t n4 = (shard_hi - shard_lo) >> 2;
  int grid = grid_for(n4 ? n4 : 1, 256, 4);
  xgpu_reduce_adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, w, delta, zero_point, exp_avg, exp_avg_sq, ch, inner,
                                                               (float)(n_levels - 1), sched, ad, grad_scale, reg_weight,
                                                               reg_loss, exit_barrier);
  B200_LAUNCH_CHECK("xgpu_reduce_adam_kernel");
  return B200LIC_OK;
}

int b200lic_wq_int_weights(const float* w, const float* alpha, const float* delta, const float* zero_point, int outer,
                           int ch, int inner, int n_levels, float* w_int, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(w && delta && zero_point && w_int, "wq_int_weights: null pointer");
  B200_REQUIRE(outer > 0 && ch > 0 && inner > 0 && n_levels >= 2, "wq_int_weights: bad shape");
  const size_t n = (size_t)outer * ch * inner;
  wq_int_weights_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(w, alpha, delta, zero_point, n, ch, inner,
                                                                         (float)(n_levels - 1), w_int);
  B200_LAUNCH_CHECK("wq_int_weights_kernel");
  return B200LIC_OK;
}

int b200lic_lsq_delta_grad(const float* w, const float* alpha, float* delta, const float* zero_point, const float* d_wq,
                           int outer, int ch, int inner, int n_levels, int soft, float grad_scale, float* d_delta,
                           float* exp_avg, float* exp_avg_sq, int step, float lr, float beta1, float beta2, float eps,
                           b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(w && delta && zero_point && d_wq, "lsq_delta_grad: null pointer");
  B200_REQUIRE(outer > 0 && ch > 0 && inner > 0 && n_levels >= 2, "lsq_delta_grad: bad shape");
  B200_REQUIRE((exp_avg && exp_avg_sq && step >= 1) || (!exp_avg && !exp_avg_sq && d_delta),
               "lsq_delta_grad: pass both Adam moments and step >= 1, or neither together with d_delta");
  AdamArgs ad{0.f, 0.f, beta1, beta2, eps};
  if (exp_avg) {
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    ad.lr_over_bc1 = (float)((double)lr / bc1);
    ad.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  }
  lsq_delta_grad_kernel<<<ch, 256, 0, as_stream(stream)>>>(w, alpha, delta, zero_point, d_wq, outer, ch, inner,
                                                           (float)(n_levels - 1), soft, grad_scale, d_delta, exp_avg,
                                                           exp_avg_sq, ad, nullptr, 1.f);
  B200_LAUNCH_CHECK("lsq_delta_grad_kernel");
  return B200LIC_OK;
}

int b200lic_lsq_delta_grad_sched(const float* w, const float* alpha, float* delta, const float* zero_point,
                                 const float* d_wq, int outer, int ch, int inner, int n_levels, int soft,
                                 float grad_scale, float* d_delta, float* exp_avg, float* exp_avg_sq,
                                 const b200lic_calib_sched* sched, float lr_scale, float beta1, float beta2, float eps,
                                 b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(w && delta && zero_point && d_wq && exp_avg && exp_avg_sq && sched, "lsq_delta_grad_sched: null pointer");
  B200_REQUIRE(outer > 0 && ch > 0 && inner > 0 && n_levels >= 2, "lsq_delta_grad_sched: bad shape");
  AdamArgs ad{0.f, 0.f, beta1, beta2, eps};
  lsq_delta_grad_kernel<<<ch, 256, 0, as_stream(stream)>>>(w, alpha, delta, zero_point, d_wq, outer, ch, inner,
                                                           (float)(n_levels - 1), soft, grad_scale, d_delta, exp_avg,
                                                           exp_avg_sq, ad, sched, lr_scale);
  B200_LAUNCH_CHECK("lsq_delta_grad_kernel(sched)");
  return B200LIC_OK;
}

}  // extern "C"
