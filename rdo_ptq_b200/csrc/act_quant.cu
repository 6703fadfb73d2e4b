// K8: activation quantisers.  HBM-bound, float4-vectorised, warp-shuffle reduced.
// Compiled with -fmad=false (activation codes are part of the bit-exact contract).
//
// Reference arithmetic restated: task-oriented-PTQ/quantization/quantizer.py:81-117 -- per channel c over
// (N,H,W): m = min, r = max(max - m, 1e-6), q = rint(clamp((x-m)/r, -1, 1) * L), out = (q/L)*r + m;
// light-uniform-PTQ/quant_int/quantizer.py:120-128 -- static Q(a_l).(a_r) fixed point.
#include <cuda_bf16.h>
#include "common.cuh"
#include "quant_math.cuh"

namespace b200lic {

constexpr int kChunk = 8192;  // elements of one (n,c) plane handled by one CTA

__global__ void actq_stats_init_kernel(unsigned* keys, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    keys[2 * i] = 0xffffffffu;  // min key
    keys[2 * i + 1] = 0u;       // max key
  }
}

// grid.x = N*C*chunks_per_plane
__global__ void __launch_bounds__(256) actq_stats_kernel(const float* __restrict__ x, int C, int HW, int chunks,
                                                          unsigned* __restrict__ keys) {
  const int chunk = blockIdx.x % chunks;
  const int plane = blockIdx.x / chunks;  // n*C + c
  const int c = plane % C;
  const float* p = x + (size_t)plane * HW;
  const int beg = chunk * kChunk, end = min(HW, beg + kChunk);
  float mn = INFINITY, mx = -INFINITY;
  if ((HW & 3) == 0 && ((uintptr_t)x & 15) == 0) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const int e4 = end >> 2, bd = blockDim.x;
    int i = (beg >> 2) + threadIdx.x;
    for (; i + 3 * bd < e4; i += 4 * bd) {           // four independent 16-byte loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(p4 + i + u * bd);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        mn = fminf(fminf(mn, v[u].x), fminf(v[u].y, fminf(v[u].z, v[u].w)));
        mx = fmaxf(fmaxf(mx, v[u].x), fmaxf(v[u].y, fmaxf(v[u].z, v[u].w)));
      }
    }
    for (; i < e4; i += bd) {
      const float4 v = __ldg(p4 + i);
      mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
      mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
    }
  } else {
    for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const float v = __ldg(p + i);
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  __shared__ float smn[8], smx[8];
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
    atomicMin(keys + 2 * c, f2key(mn));
    atomicMax(keys + 2 * c + 1, f2key(mx));
  }
}

__device__ __forceinline__ float actq_one(float v, float m, float r, float L, float* code) {
  float t = __fdiv_rn(__fsub_rn(v, m), r);
  t = fminf(fmaxf(t, -1.f), 1.f);
  const float q = rintf(__fmul_rn(t, L));
  if (code) *code = q;
  return __fadd_rn(__fmul_rn(__fdiv_rn(q, L), r), m);
}

// actq_one with both divisions by the FMA sequence (ry = RN(1/r), Ly = RN(1/L)); bit-identical results
__device__ __forceinline__ float actq_one_fast(float v, float m, float r, float ry, float L, float Ly) {
  float t = div_rn(__fsub_rn(v, m), r, ry);
  t = fminf(fmaxf(t, -1.f), 1.f);
  const float q = rintf(__fmul_rn(t, L));
  return __fadd_rn(__fmul_rn(div_rn(q, L, Ly), r), m);
}

__global__ void __launch_bounds__(256) actq_apply_kernel(const float* __restrict__ x, const unsigned* __restrict__ keys,
                                                          int C, int HW, int chunks, float L, float* __restrict__ out,
                                                          float* __restrict__ codes) {
  const int chunk = blockIdx.x % chunks;
  const int plane = blockIdx.x / chunks;
  const int c = plane % C;
  const float m = key2f(keys[2 * c]);
  const float r = fmaxf(__fsub_rn(key2f(keys[2 * c + 1]), m), 1e-6f);
  const size_t base = (size_t)plane * HW;
  const int beg = chunk * kChunk, end = min(HW, beg + kChunk);
  const bool vec = (HW & 3) == 0 && (((uintptr_t)x | (uintptr_t)out | (uintptr_t)codes) & 15) == 0;
  if (vec) {
    const float4* p4 = reinterpret_cast<const float4*>(x + base);
    float4* o4 = reinterpret_cast<float4*>(out + base);
    float4* c4 = codes ? reinterpret_cast<float4*>(codes + base) : nullptr;
    const int e4 = end >> 2, bd = blockDim.x;
    auto quant4 = [&](int i, const float4& v) {
      float4 o, q;
      o.x = actq_one(v.x, m, r, L, &q.x);
      o.y = actq_one(v.y, m, r, L, &q.y);
      o.z = actq_one(v.z, m, r, L, &q.z);
      o.w = actq_one(v.w, m, r, L, &q.w);
      o4[i] = o;
      if (c4) c4[i] = q;
    };
    int i = (beg >> 2) + threadIdx.x;
    for (; i + 3 * bd < e4; i += 4 * bd) {           // four independent 16-byte loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(p4 + i + u * bd);
#pragma unroll
      for (int u = 0; u < 4; ++u) quant4(i + u * bd, v[u]);
    }
    for (; i < e4; i += bd) quant4(i, __ldg(p4 + i));
  } else {
    for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
      float q;
      out[base + i] = actq_one(__ldg(x + base + i), m, r, L, &q);
      if (codes) codes[base + i] = q;
    }
  }
}

// ---- K8 fused: statistics + apply in ONE launch, one thread-block cluster per channel ---------------------------------
// The three-launch form (init, min/max with global atomics, apply) reads the activation twice from HBM (12 B/elem) and
// costs three launch latencies per layer -- the dominant share of the W8A8 evaluation forward's non-GEMM time
// (profiles/README.md r1d: 27 % of the Cheng2020 forward).  Here the CTAs of a cluster split the (N, H*W) elements of a
// channel, keep their slice in shared memory while reducing it, exchange the per-CTA (min, max) through distributed
// shared memory, and quantise the slice they still hold: 8 B/elem, no atomics, no initialisation pass.  Slices that
// do not fit in shared memory are re-read from global memory (L2) in the second pass.  Arithmetic is actq_one(), so
// codes and outputs are bit-identical to the three-launch path.
constexpr int kFuseThreads = 512;
constexpr int kFuseKeepBytes = 160 * 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_peer_f32(const float* local, uint32_t rank) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(local);
  uint32_t pa;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(pa) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(pa));
  return v;
}

__global__ void __launch_bounds__(kFuseThreads, 1)
    actq_cluster_kernel(const float* __restrict__ x, int N, int C, int HW, int S, int ps, int keep, float L,
                        float* __restrict__ out, float* __restrict__ codes) {
  extern __shared__ float4 kept4[];
  float* kept = reinterpret_cast<float*>(kept4);
  __shared__ float s_mn[kFuseThreads / 32], s_mx[kFuseThreads / 32];
  __shared__ float s_stat[2];
  const uint32_t rank = cluster_ctarank();
  const int c = blockIdx.x / S;
  // CTA `rank` owns elements [beg, end) of EVERY plane (n, c): no index division anywhere in the loops
  const int beg = (int)rank * ps, end = min(HW, beg + ps);
  const bool vec = (HW & 3) == 0 && (((uintptr_t)x | (uintptr_t)out | (uintptr_t)codes) & 15) == 0;   // ps % 4 == 0
  float mn = INFINITY, mx = -INFINITY;
  for (int n = 0; n < N; ++n) {
    const float* p = x + ((size_t)n * C + c) * HW;
    if (vec) {
      float4* k4 = kept4 + (size_t)n * (ps >> 2);
      for (int i = beg + 4 * (int)threadIdx.x; i < end; i += 4 * kFuseThreads) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + i));
        if (keep) k4[(i - beg) >> 2] = v;
        mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
        mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
      }
    } else {
      float* k1 = kept + (size_t)n * ps;
      for (int i = beg + (int)threadIdx.x; i < end; i += kFuseThreads) {
        const float v = __ldg(p + i);
        if (keep) k1[i - beg] = v;
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
      }
    }
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) {
    s_mn[threadIdx.x >> 5] = mn;
    s_mx[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < kFuseThreads / 32; ++i) {
      mn = fminf(mn, s_mn[i]);
      mx = fmaxf(mx, s_mx[i]);
    }
    s_stat[0] = mn;
    s_stat[1] = mx;
  }
  cluster_sync_all();                              // every CTA's (min, max) is visible cluster-wide
  float gmn = INFINITY, gmx = -INFINITY;
  for (int r = 0; r < S; ++r) {
    gmn = fminf(gmn, ld_peer_f32(&s_stat[0], (uint32_t)r));
    gmx = fmaxf(gmx, ld_peer_f32(&s_stat[1], (uint32_t)r));
  }
  const float m = gmn, rr = fmaxf(__fsub_rn(gmx, gmn), 1e-6f);
  for (int n = 0; n < N; ++n) {
    const size_t base = ((size_t)n * C + c) * HW;
    if (vec) {
      const float4* k4 = kept4 + (size_t)n * (ps >> 2);
      for (int i = beg + 4 * (int)threadIdx.x; i < end; i += 4 * kFuseThreads) {
        const float4 v = keep ? k4[(i - beg) >> 2] : __ldg(reinterpret_cast<const float4*>(x + base + i));
        float4 o, q;
        o.x = actq_one(v.x, m, rr, L, &q.x);
        o.y = actq_one(v.y, m, rr, L, &q.y);
        o.z = actq_one(v.z, m, rr, L, &q.z);
        o.w = actq_one(v.w, m, rr, L, &q.w);
        *reinterpret_cast<float4*>(out + base + i) = o;
        if (codes) *reinterpret_cast<float4*>(codes + base + i) = q;
      }
    } else {
      const float* k1 = kept + (size_t)n * ps;
      for (int i = beg + (int)threadIdx.x; i < end; i += kFuseThreads) {
        float q;
        out[base + i] = actq_one(keep ? k1[i - beg] : __ldg(x + base + i), m, rr, L, &q);
        if (codes) codes[base + i] = q;
      }
    }
  }
  cluster_sync_all();                              // no CTA leaves while a peer may still read its s_stat
}

__global__ void __launch_bounds__(256) fixed_point_kernel(const float* __restrict__ x, size_t n, float lo, float hi,
                                                           float mult, float* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((((uintptr_t)x | (uintptr_t)out) & 15) == 0) {
    const size_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (size_t i = tid; i < n4; i += stride) {
      float4 v = __ldg(x4 + i);
      v.x = __fdiv_rn(rintf(__fmul_rn(fminf(fmaxf(v.x, lo), hi), mult)), mult);
      v.y = __fdiv_rn(rintf(__fmul_rn(fminf(fmaxf(v.y, lo), hi), mult)), mult);
      v.z = __fdiv_rn(rintf(__fmul_rn(fminf(fmaxf(v.z, lo), hi), mult)), mult);
      v.w = __fdiv_rn(rintf(__fmul_rn(fminf(fmaxf(v.w, lo), hi), mult)), mult);
      o4[i] = v;
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += stride)
      out[i] = __fdiv_rn(rintf(__fmul_rn(fminf(fmaxf(x[i], lo), hi), mult)), mult);
  } else {
    for (size_t i = tid; i < n; i += stride)
      out[i] = __fdiv_rn(rintf(__fmul_rn(fminf(fmaxf(x[i], lo), hi), mult)), mult);
  }
}

// ---- K8 apply fused with the consumer's operand staging ---------------------------------------------------------------
// The quantised activation's next reader on the evaluation path is the following layer's tensor-core GEMM, whose operand
// is the split-bf16 NHWC copy of it (conv_tc.cu nhwc_split_kernel).  actq_apply + nhwc_split read the tensor twice and
// write it twice (16 B/elem); here one pass quantises (actq_one: same codes, bit for bit) and writes the staged operand
// (x, or x*x for a GDN consumer) -- and, only when someone needs it (GDN's epilogue), the fp32 quantised tensor.
__global__ void __launch_bounds__(256)
    actq_apply_stage_kernel(const float* __restrict__ x, const unsigned* __restrict__ keys, int C, int HW, int cpad, float L,
                            int square, __nv_bfloat16* __restrict__ xh, __nv_bfloat16* __restrict__ xl,
                            float* __restrict__ out) {
  __shared__ float t[64][65];
  __shared__ float s_m[64], s_r[64], s_ry[64];
  const int cblk = (cpad + 63) >> 6;
  const int n = blockIdx.z, c0 = (int)(blockIdx.x % cblk) * 64, p0 = (int)(blockIdx.x / cblk) * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t img = (size_t)n * C * HW;
  const bool vec2 = (HW & 1) == 0;
  // per-channel (min, range, RN(1/range)) once per CTA: the two divisions of actq_one become FMA sequences on a
  // precomputed reciprocal (quant_math.cuh: same correctly rounded quotient, a third of the instructions)
  if (threadIdx.x < 64) {
    const int c = c0 + threadIdx.x;
    float m = 0.f, r = 1.f;
    if (c < C) {
      m = key2f(keys[2 * c]);
      r = fmaxf(__fsub_rn(key2f(keys[2 * c + 1]), m), 1e-6f);
    }
    s_m[threadIdx.x] = m;
    s_r[threadIdx.x] = r;
    s_ry[threadIdx.x] = div_rn_ok(r) ? __frcp_rn(r) : 0.f;
  }
  __syncthreads();
  const float Ly = div_rn_ok(L) ? __frcp_rn(L) : 0.f;
  // All eight 8-byte loads of a thread are issued before the first one is used: with one load in flight per thread the
  // kernel ran at 55 % of the HBM rate (2048 threads x 8 B = 16 KB in flight per SM against ~35 KB that the latency needs).
  float2 raw[8];
  if (vec2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + warp + 8 * j, p = p0 + 2 * lane;
      raw[j] = (c < C && p < HW) ? __ldg(reinterpret_cast<const float2*>(x + img + (size_t)c * HW + p))
                                 : make_float2(0.f, 0.f);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int cl = warp + 8 * j, c = c0 + cl, p = p0 + 2 * lane;
    float v0 = 0.f, v1 = 0.f;
    if (c < C && p < HW) {
      const float m = s_m[cl], r = s_r[cl], ry = s_ry[cl];
      const bool fast = ry != 0.f && Ly != 0.f;                  // uniform over the warp (one channel per warp and j)
      const size_t e = img + (size_t)c * HW + p;
      const bool has1 = p + 1 < HW;
      if (vec2) {
        const float2 v = raw[j];
        v0 = fast ? actq_one_fast(v.x, m, r, ry, L, Ly) : actq_one(v.x, m, r, L, nullptr);
        v1 = fast ? actq_one_fast(v.y, m, r, ry, L, Ly) : actq_one(v.y, m, r, L, nullptr);
        if (out) *reinterpret_cast<float2*>(out + e) = make_float2(v0, v1);
      } else {
        const float a0 = __ldg(x + e);
        v0 = fast ? actq_one_fast(a0, m, r, ry, L, Ly) : actq_one(a0, m, r, L, nullptr);
        if (has1) {
          const float a1 = __ldg(x + e + 1);
          v1 = fast ? actq_one_fast(a1, m, r, ry, L, Ly) : actq_one(a1, m, r, L, nullptr);
        }
        if (out) {
          out[e] = v0;
          if (has1) out[e + 1] = v1;
        }
      }
    }
    if (square) {
      v0 = __fmul_rn(v0, v0);
      v1 = __fmul_rn(v1, v1);
    }
    t[cl][2 * lane] = v0;
    t[cl][2 * lane + 1] = v1;
  }
  __syncthreads();
  if (c0 + 2 * lane >= cpad) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int pl = warp + 8 * j, p = p0 + pl;
    if (p < HW) {
      const float a = t[2 * lane][pl], b = t[2 * lane + 1][pl];
      const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
      __nv_bfloat162 hv, lv;
      hv.x = ah;
      hv.y = bh;
      lv.x = __float2bfloat16_rn(__fsub_rn(a, __bfloat162float(ah)));
      lv.y = __float2bfloat16_rn(__fsub_rn(b, __bfloat162float(bh)));
      const size_t o = ((size_t)n * HW + p) * cpad + c0 + 2 * lane;
      *reinterpret_cast<__nv_bfloat162*>(xh + o) = hv;
      *reinterpret_cast<__nv_bfloat162*>(xl + o) = lv;
    }
  }
}

}  // namespace b200lic

using namespace b200lic;

extern "C" {

int b200lic_actq_stats_init(float* minmax, int C, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(minmax && C > 0, "actq_stats_init: bad arguments");
  actq_stats_init_kernel<<<(C + 255) / 256, 256, 0, as_stream(stream)>>>(reinterpret_cast<unsigned*>(minmax), C);
  B200_LAUNCH_CHECK("actq_stats_init_kernel");
  return B200LIC_OK;
}

int b200lic_actq_stats(const float* x, int N, int C, int HW, float* minmax, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && minmax && N > 0 && C > 0 && HW > 0, "actq_stats: bad arguments");
  const int chunks = (HW + kChunk - 1) / kChunk;
  const long long blocks = (long long)N * C * chunks;
  B200_REQUIRE(blocks < 2147483647LL, "actq_stats: tensor too large");
  actq_stats_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, C, HW, chunks,
                                                                      reinterpret_cast<unsigned*>(minmax));
  B200_LAUNCH_CHECK("actq_stats_kernel");
  return B200LIC_OK;
}

int b200lic_actq_apply_stage(const float* x, const float* minmax, int N, int C, int HW, int n_bits, int square,
                             void* x_hi, void* x_lo, int cpad, float* out, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && minmax && x_hi && x_lo && N > 0 && N <= 65535 && C > 0 && HW > 0, "actq_apply_stage: bad arguments");
  B200_REQUIRE(n_bits >= 2 && n_bits <= 16, "actq_apply_stage: n_bits=%d outside [2,16]", n_bits);
  B200_REQUIRE(cpad >= C && (cpad % 32) == 0, "actq_apply_stage: cpad=%d for %d channels", cpad, C);
  B200_REQUIRE(((((uintptr_t)x) | ((uintptr_t)out)) & 7) == 0, "actq_apply_stage: 8-byte alignment");
  dim3 grid((unsigned)(((cpad + 63) / 64) * ((HW + 63) / 64)), 1, (unsigned)N);
  actq_apply_stage_kernel<<<grid, 256, 0, as_stream(stream)>>>(
      x, reinterpret_cast<const unsigned*>(minmax), C, HW, cpad, (float)((1 << n_bits) - 1), square,
      reinterpret_cast<__nv_bfloat16*>(x_hi), reinterpret_cast<__nv_bfloat16*>(x_lo), out);
  B200_LAUNCH_CHECK("actq_apply_stage_kernel");
  return B200LIC_OK;
}

int b200lic_actq_apply(const float* x, const float* minmax, int N, int C, int HW, int n_bits, float* out, float* codes,
                       b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && minmax && out && N > 0 && C > 0 && HW > 0, "actq_apply: bad arguments");
  B200_REQUIRE(n_bits >= 2 && n_bits <= 16, "actq_apply: n_bits=%d outside [2,16]", n_bits);
  const int chunks = (HW + kChunk - 1) / kChunk;
  const long long blocks = (long long)N * C * chunks;
  B200_REQUIRE(blocks < 2147483647LL, "actq_apply: tensor too large");
  actq_apply_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      x, reinterpret_cast<const unsigned*>(minmax), C, HW, chunks, (float)((1 << n_bits) - 1), out, codes);
  B200_LAUNCH_CHECK("actq_apply_kernel");
  return B200LIC_OK;
}

int b200lic_actq_fused(const float* x, int N, int C, int HW, int n_bits, float* out, float* codes,
                       b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && out && N > 0 && C > 0 && HW > 0, "actq_fused: bad arguments");
  B200_REQUIRE(n_bits >= 2 && n_bits <= 16, "actq_fused: n_bits=%d outside [2,16]", n_bits);
  const long long bytes = (long long)N * HW * 4;
  // cluster size: enough CTAs to keep the channel's elements in shared memory, and enough to occupy the chip
  int S = 1;
  while (S < 8 && bytes > (long long)S * kFuseKeepBytes) S *= 2;
  while (S < 8 && (long long)C * S < 2LL * num_sms() && HW / (2 * S) >= 2048) S *= 2;
  const int ps = ((HW + S - 1) / S + 3) / 4 * 4;              // elements of every plane owned by one CTA
  const int keep = (long long)N * ps * 4 <= (long long)kFuseKeepBytes ? 1 : 0;
  const size_t smem = keep ? (size_t)N * ps * 4 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(actq_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuseKeepBytes);
    if (e != cudaSuccess) {
      set_error("actq_fused: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
      return B200LIC_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(C * S), 1, 1);
  cfg.blockDim = dim3(kFuseThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)S;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, actq_cluster_kernel, x, N, C, HW, S, ps, keep, (float)((1 << n_bits) - 1), out,
                                     codes);
  if (e != cudaSuccess) {
    set_error("actq_fused: launch failed: %s", cudaGetErrorString(e));
    return B200LIC_ERR_CUDA;
  }
  B200_LAUNCH_CHECK("actq_cluster_kernel");
  return B200LIC_OK;
}

int b200lic_fixed_point(const float* x, size_t n, int a_l, int a_r, float* out, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && out, "fixed_point: null pointer");
  B200_REQUIRE(a_l >= 1 && a_l <= 24 && a_r >= 0 && a_r <= 24, "fixed_point: bad format Q%d.%d", a_l, a_r);
  if (n == 0) return B200LIC_OK;
  const float hi = (float)(1 << (a_l - 1));
  fixed_point_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, as_stream(stream)>>>(x, n, -hi, hi, (float)(1 << a_r), out);
  B200_LAUNCH_CHECK("fixed_point_kernel");
  return B200LIC_OK;
}

}  // extern "C"
