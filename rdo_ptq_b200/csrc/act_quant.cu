// K8: activation quantisers.  HBM-bound, float4-vectorised, warp-shuffle reduced.
// Compiled with -fmad=false (activation codes are part of the bit-exact contract).
//
// Reference arithmetic restated: task-oriented-PTQ/quantization/quantizer.py:81-117 -- per channel c over
// (N,H,W): m = min, r = max(max - m, 1e-6), q = rint(clamp((x-m)/r, -1, 1) * L), out = (q/L)*r + m;
// light-uniform-PTQ/quant_int/quantizer.py:120-128 -- static Q(a_l).(a_r) fixed point.
#include "common.cuh"

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
    for (int i = (beg >> 2) + threadIdx.x; i < (end >> 2); i += blockDim.x) {
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
    for (int i = (beg >> 2) + threadIdx.x; i < (end >> 2); i += blockDim.x) {
      const float4 v = __ldg(p4 + i);
      float4 o, q;
      o.x = actq_one(v.x, m, r, L, &q.x);
      o.y = actq_one(v.y, m, r, L, &q.y);
      o.z = actq_one(v.z, m, r, L, &q.z);
      o.w = actq_one(v.w, m, r, L, &q.w);
      o4[i] = o;
      if (c4) c4[i] = q;
    }
  } else {
    for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
      float q;
      out[base + i] = actq_one(__ldg(x + base + i), m, r, L, &q);
      if (codes) codes[base + i] = q;
    }
  }
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
