// GDN / IGDN forward as ONE kernel over the raw fp32 NCHW tensor (B200LIC_ENGINE_TC, evaluation path).
//
//   y[n,c,p] = xq[n,c,p] * rsqrt(beta[c] + sum_k gamma[c,k] * xq[n,k,p]^2)          (IGDN: * sqrt)
//   xq = x, or the dynamic per-channel activation quantiser of x (the producer deferred it: ops.DEFER_ACTQ)
//
// compressai.layers.GDN.forward is F.conv2d(x**2, gamma, beta) + rsqrt + mul (five full-tensor passes, TO quant_layer.py
// wraps it as a QuantModule).  The conv engine (conv_tc2.cu gdn_mode) already fused the epilogue, but its operand came
// from a staging pass: x -> [quantise] -> x^2 -> split-bf16 NHWC (4 B/elem written, 4 B/elem read back) next to the fp32
// xq the epilogue multiplies (4 B/elem written by the quantiser, 4 B/elem read): 20 B/elem of HBM traffic for an op whose
// algorithmic traffic is 8 B/elem (read x, write y).  At 1536x2048 the two full-resolution GDN layers and their staging
// passes were 2.0 ms of an 8.4 ms forward (profiles/r2_launches_fwd_2k.csv).
//
// Here the operand is produced on chip:
//   warp 0      TMA: fp32 [32 ch][128 px] boxes of x (NCHW is already "pixel-contiguous per channel") -> X ring
//   warp 2      TMA: [C][32] hi|lo K-block of the packed gamma (tc2_weight_layout form, L2 resident) -> B ring
//   warps 3-10  converters: X ring -> quantise -> square -> split bf16 -> A ring.  The pixel axis is the M axis of the
//               GEMM and it is the contiguous one in shared memory, so A is fed MN-major (SWIZZLE_128B: a K row holds 64
//               pixels = 128 B; 16-byte chunk j of K row c lands at chunk j ^ (c & 7)) -- no transpose anywhere.  With a
//               deferred quantiser the integer codes of the tile (1 byte each) stay in shared memory for the epilogue.
//   warp 1      tcgen05.mma issuer: D[128 px, C] += A[128, 32] * B[C, 32]^T, three split-bf16 passes, two TMEM accumulators
//   warps 11-18 epilogue: TMEM -> + beta -> rsqrt/sqrt -> * xq -> coalesced fp32 stores.  xq is rebuilt from the code
//               tile (deferred quantiser), or x is re-read through L2 (the tile was fetched a few microseconds earlier).
// HBM traffic: 4 B/elem read + 4 B/elem written.
//
// The quantiser's two fp32 divisions ((v - min) / range and code / levels) have a per-channel resp. constant divisor:
// div_rn() below is Markstein's FMA sequence on the correctly rounded reciprocal -- correctly rounded for every divisor
// whose significand is not all ones (those channels take __fdiv_rn), so the codes stay those of b200lic_actq_apply bit
// for bit (b200lic_selftest_fast_div counts mismatches against __fdiv_rn; tests/test_gpu_prepared.py).
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "quant_math.cuh"

namespace b200lic {

bool tc_encode_map_ex(CUtensorMap* m, CUtensorMapDataType dt, CUtensorMapSwizzle sw, void* base, int rank,
                      const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box, const cuuint32_t* estr);
bool tc2_weight_layout(int Cin, int Cout, int KH, int KW, int stride, int transposed, int* Cpad, int* CoutPad, int* Tmax,
                       int* phases, size_t* b_bytes);

void tc2_stats_once(unsigned* keys);
unsigned* tc2_stats_peek();

namespace gd {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// same, for warps that wait long (converters, epilogue): the suspend-time hint keeps their polling off the issue ports
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr)
      : "memory");
}
// A: MN-major SWIZZLE_128B (64 pixels = 128 B per K row, 8 K rows per 1024 B atom; SBO = next 8 K rows, LBO = next 64 pixels)
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// B: K-major SWIZZLE_64B (rows of 32 bf16 = 64 B, 8-row groups of 512 B) -- the packed-weight form of conv_tc2.cu
__device__ __forceinline__ uint64_t make_kmajor_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// act_quant.cu's actq_one, restated with explicit round-to-nearest intrinsics (never contracted into FMAs): the codes
// this kernel produces are the codes b200lic_actq_apply produces, bit for bit.
//   code = rint(clamp((v - m) / r, -1, 1) * L): the division by the FMA sequence (FAST) or __fdiv_rn; rint by adding
//   1.5 * 2^23 (round-to-nearest-even like rintf; the integer then sits in the low significand bits: no F2I / FRND)
//   value = (code / L) * r + m: code / L from a table built with __fdiv_rn
constexpr float kRintMagic = 12582912.f;
template <bool FAST>
__device__ __forceinline__ uint32_t quant_bits(float v, float m, float r, float ry, float L) {
  const float a = __fsub_rn(v, m);
  float t = FAST ? div_rn(a, r, ry) : __fdiv_rn(a, r);
  t = fminf(fmaxf(t, -1.f), 1.f);
  return __float_as_uint(__fadd_rn(__fmul_rn(t, L), kRintMagic));        // low byte = the code
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// MUFU.RSQ alone: rsqrtf() without the denormal pre-scaling (the norm is >= beta > 0, far from denormal)
__device__ __forceinline__ float rsqrt_fast(float v) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
// MUFU.SQRT alone (IGDN): sqrtf() is the IEEE-rounded sequence with a slow-path call; the norm is a normal positive number
__device__ __forceinline__ float sqrt_fast(float v) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

}  // namespace gd

constexpr int kGdThreads = 608;            // 19 warps: X-TMA, MMA, B-TMA, 8 converters, 8 epilogue
constexpr int kGdConv = 256, kGdEpi = 256; // converter / epilogue threads
constexpr uint32_t kGdXStage = 32 * 128 * 4;   // fp32 [32 channels][128 pixels]
constexpr uint32_t kGdAStage = 2 * 128 * 32 * 2;   // bf16 hi | lo, each [2 pixel blocks][32 K rows][64 pixels]
constexpr int kGdMaxSX = 8, kGdMaxSA = 4;

struct GdnGeom {
  int N, C, HW;
  int nkb;              // K blocks of 32 channels
  int BN;               // accumulator columns = padded channel count of the packed gamma (multiple of 16, <= 256)
  int tiles_per_img, n_tiles;
  int SX, SA;           // ring depths
  int acc_cols;         // TMEM columns per accumulator (power of two >= BN)
  int inverse, has_q;
  float L;              // quantiser levels - 1
};

__global__ void __launch_bounds__(kGdThreads, 1)
    gdn_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_b, GdnGeom g,
                     const float* __restrict__ x, const unsigned* __restrict__ keys, const float* __restrict__ beta,
                     float* __restrict__ y, unsigned* __restrict__ stats) {
  pdl_wait();                 // programmatic dependent launch (common.cuh): predecessor complete, memory visible
  using namespace gd;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t b_stage = 2u * (uint32_t)g.BN * 64u;
  const uint32_t q_tile = g.has_q ? (uint32_t)g.nkb * 32u * 128u : 0u;     // one code byte per (channel, pixel)
  const uint32_t x_ring = smem_base;
  const uint32_t a_ring = x_ring + (uint32_t)g.SX * kGdXStage;
  const uint32_t b_ring = a_ring + (uint32_t)g.SA * kGdAStage;
  const uint32_t q_buf = b_ring + (uint32_t)g.SA * b_stage;
  const uint32_t par = q_buf + 2u * q_tile;
  float4* s_par = reinterpret_cast<float4*>(smem_gen + (par - smem_base));    // per channel: beta, min, range, RN(1/range)
  float* s_lut = reinterpret_cast<float*>(s_par + 256);                        // code / L
  const uint32_t par_a = par, lut_a = par + 4096u;                             // the same, as shared-window addresses
  // [256] min keys, [256] max keys of the OUTPUT channels (stats != nullptr): statistics for the dynamic quantiser of the
  // layer after this one, taken from the epilogue's registers (b200lic_conv_stats_once, see conv_tc2.cu)
  unsigned* s_stat = reinterpret_cast<unsigned*>(s_lut + 256);
  const uint32_t bars = par + 7u * 256u * 4u;
  const uint32_t x_full = bars, x_empty = x_full + 8u * kGdMaxSX;
  const uint32_t a_full = x_empty + 8u * kGdMaxSX, ab_empty = a_full + 8u * kGdMaxSA, b_full = ab_empty + 8u * kGdMaxSA;
  const uint32_t t_full = b_full + 8u * kGdMaxSA, t_empty = t_full + 16u;
  const uint32_t q_full = t_empty + 16u, q_empty = q_full + 16u;
  const uint32_t tmem_ptr_addr = q_empty + 16u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_addr - smem_base));

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.SX; ++s) {
      mbar_init(x_full + 8u * s, 1);
      mbar_init(x_empty + 8u * s, kGdConv);
    }
    for (int s = 0; s < g.SA; ++s) {
      mbar_init(a_full + 8u * s, kGdConv);
      mbar_init(ab_empty + 8u * s, 1);
      mbar_init(b_full + 8u * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full + 8u * s, 1);
      mbar_init(t_empty + 8u * s, kGdEpi);
      mbar_init(q_full + 8u * s, kGdConv);
      mbar_init(q_empty + 8u * s, kGdEpi);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const float L = g.L;
  for (int c = threadIdx.x; c < 256; c += kGdThreads) {
    const bool in = c < g.C;
    float m = 0.f, r = 1.f;
    if (in && g.has_q) {
      m = key2f(keys[2 * c]);
      r = fmaxf(__fsub_rn(key2f(keys[2 * c + 1]), m), 1e-6f);
    }
    s_par[c] = make_float4(in ? __ldg(beta + c) : 1.f, m, r, div_rn_ok(r) ? __frcp_rn(r) : 0.f);
    s_lut[c] = (float)c <= L ? __fdiv_rn((float)c, L) : 0.f;
    s_stat[c] = 0xffffffffu;
    s_stat[256 + c] = 0u;
  }
  if (warp == 1) {
    const uint32_t cols = 2u * (uint32_t)g.acc_cols;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_gen;

  const int first = blockIdx.x, step = gridDim.x;

  if (warp == 0) {
    // ===== TMA: x boxes =================================================================================================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = first; t < g.n_tiles; t += step) {
        const int n = t / g.tiles_per_img, p0 = (t - n * g.tiles_per_img) * 128;
        for (int kb = 0; kb < g.nkb; ++kb) {
          mbar_wait_long(x_empty + 8u * s, ph ^ 1u);
          mbar_expect_tx(x_full + 8u * s, kGdXStage);
          tma_load_3d(x_ring + (uint32_t)s * kGdXStage, &map_x, x_full + 8u * s, p0, kb * 32, n);
          if (++s == g.SX) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===== TMA: gamma K blocks (hi and lo slab in one instruction) =======================================================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = first; t < g.n_tiles; t += step) {
        for (int kb = 0; kb < g.nkb; ++kb) {
          mbar_wait_long(ab_empty + 8u * s, ph ^ 1u);
          mbar_expect_tx(b_full + 8u * s, b_stage);
          tma_load_4d(b_ring + (uint32_t)s * b_stage, &map_b, b_full + 8u * s, kb * 32, 0, 0, 0);
          if (++s == g.SA) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ===================================================================================================
    if (elect_one()) {
      // D = f32, A = B = bf16, A MN-major (bit 15), B K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(g.BN >> 3) << 17) |
                             ((128u >> 4) << 24);
      const uint64_t a_d0 = make_mnmajor_sw128_desc(a_ring, 4096u);
      const uint64_t b_d0 = make_kmajor_sw64_desc(b_ring);
      const uint64_t a_lo = (uint64_t)((kGdAStage / 2) >> 4), b_lo = (uint64_t)(((uint32_t)g.BN * 64u) >> 4);
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph[2] = {0u, 0u};
      for (int t = first; t < g.n_tiles; t += step) {
        mbar_wait(t_empty + 8u * acc, acc_ph[acc] ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + (uint32_t)(acc * g.acc_cols);
        for (int kb = 0; kb < g.nkb; ++kb) {
          mbar_wait(b_full + 8u * s, ph);
          mbar_wait(a_full + 8u * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t a0 = a_d0 + (uint64_t)((uint32_t)s * (kGdAStage >> 4));
          const uint64_t b0 = b_d0 + (uint64_t)((uint32_t)s * (b_stage >> 4));
#pragma unroll
          for (int k = 0; k < 2; ++k) {          // UMMA_K = 16 channels: 16 K rows of A = 2048 B, 32 B inside B's 64 B row
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {   // hi*hi, hi*lo, lo*hi
              const uint64_t ad = a0 + (pass == 2 ? a_lo : 0) + (uint64_t)(k * (2048 >> 4));
              const uint64_t bd = b0 + (pass == 1 ? b_lo : 0) + (uint64_t)(k * 2);
              umma_bf16(d, ad, bd, idesc, (kb != 0 || k != 0 || pass != 0) ? 1u : 0u);
            }
          }
          umma_commit(ab_empty + 8u * s);
          if (++s == g.SA) {
            s = 0;
            ph ^= 1u;
          }
        }
        umma_commit(t_full + 8u * acc);
        acc_ph[acc] ^= 1u;
        acc ^= 1;
      }
    }
  } else if (warp >= 3 && warp < 11) {
    // ===== converters: fp32 X stage -> [quantise] -> square -> split bf16, MN-major SW128 A stage ==========================
    const int tid = threadIdx.x - 96;            // 0..255
    const int pg = tid & 15;                     // 8-pixel group of the 128-pixel tile
    const int cb = (tid >> 4) & 7;               // channel within a group of 8
    const int half = tid >> 7;                   // chunk i handles channel cb + 8 * (2 * half + i)
    const uint32_t a_off = (uint32_t)(pg >> 3) * 4096u + (uint32_t)(2 * half) * 1024u + (uint32_t)cb * 128u +
                           (uint32_t)(((pg & 7) ^ cb) << 4);
    const uint32_t x_off = (uint32_t)(cb + 16 * half) * 512u + (uint32_t)pg * 32u;
    int sx = 0, sa = 0, buf = 0;
    uint32_t phx = 0, pha = 0;
    uint32_t buf_ph[2] = {0u, 0u};
    for (int t = first; t < g.n_tiles; t += step) {
      if (g.has_q) mbar_wait_long(q_empty + 8u * buf, buf_ph[buf] ^ 1u);   // the epilogue two tiles back is done with it
      for (int kb = 0; kb < g.nkb; ++kb) {
        mbar_wait_long(x_full + 8u * sx, phx);
        mbar_wait_long(ab_empty + 8u * sa, pha ^ 1u);
        const uint32_t xs = x_ring + (uint32_t)sx * kGdXStage + x_off;
        const uint32_t as = a_ring + (uint32_t)sa * kGdAStage + a_off;
#pragma unroll
        for (int i = 0; i < 2; ++i) {            // channel cl = cb + 8 (2 half + i): K row at (cl >> 3) * 1024 + (cl & 7) * 128
          float v[8];
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
                       : "r"(xs + (uint32_t)i * 4096u));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                       : "r"(xs + (uint32_t)i * 4096u + 16u));
          if (g.has_q) {
            const int c = kb * 32 + cb + 8 * (2 * half + i);
            const float4 pr = lds_f32x4(par_a + (uint32_t)c * 16u);
            uint32_t bits[8];
            if (pr.w != 0.f) {                   // uniform over the chunk: one branch per 8 elements
#pragma unroll
              for (int j = 0; j < 8; ++j) bits[j] = quant_bits<true>(v[j], pr.y, pr.z, pr.w, L);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) bits[j] = quant_bits<false>(v[j], pr.y, pr.z, pr.w, L);
            }
            // (pixels beyond the tensor are zero-filled and may fall below the channel minimum: their low byte is not a
            //  code, and nobody reads it back -- those accumulator rows are never stored)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __fadd_rn(__fmul_rn(lds_f32(lut_a + ((bits[j] & 0xffu) << 2)), pr.z), pr.y);
            const uint32_t w0 = __byte_perm(__byte_perm(bits[0], bits[1], 0x0040), __byte_perm(bits[2], bits[3], 0x0040), 0x5410);
            const uint32_t w1 = __byte_perm(__byte_perm(bits[4], bits[5], 0x0040), __byte_perm(bits[6], bits[7], 0x0040), 0x5410);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(q_buf + (uint32_t)buf * q_tile + (uint32_t)c * 128u +
                                                                   (uint32_t)pg * 8u),
                         "r"(w0), "r"(w1)
                         : "memory");
          }
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a = __fmul_rn(v[2 * j], v[2 * j]), b = __fmul_rn(v[2 * j + 1], v[2 * j + 1]);
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);        // one F2FP for the pair
            hi[j] = *reinterpret_cast<const uint32_t*>(&h2);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(__fsub_rn(a, __uint_as_float(hi[j] << 16)),
                                                            __fsub_rn(b, __uint_as_float(hi[j] & 0xffff0000u)));
            lo[j] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(as + (uint32_t)i * 1024u), "r"(hi[0]), "r"(hi[1]),
                       "r"(hi[2]), "r"(hi[3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(as + (uint32_t)i * 1024u + kGdAStage / 2), "r"(lo[0]),
                       "r"(lo[1]), "r"(lo[2]), "r"(lo[3])
                       : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
        mbar_arrive(a_full + 8u * sa);
        mbar_arrive(x_empty + 8u * sx);
        if (++sx == g.SX) {
          sx = 0;
          phx ^= 1u;
        }
        if (++sa == g.SA) {
          sa = 0;
          pha ^= 1u;
        }
      }
      if (g.has_q) {
        mbar_arrive(q_full + 8u * buf);
        buf_ph[buf] ^= 1u;
        buf ^= 1;
      }
    }
  } else if (warp >= 11) {
    // ===== epilogue =====================================================================================================
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;               // pixel of the tile
    const int eh = (warp - 11) >> 2;             // the two warps of a quarter take alternate 16-channel chunks
    int acc = 0;
    uint32_t acc_ph[2] = {0u, 0u};
    for (int t = first; t < g.n_tiles; t += step) {
      const int n = t / g.tiles_per_img, p0 = (t - n * g.tiles_per_img) * 128, p = p0 + row;
      const bool valid = p < g.HW;
      const bool full = p0 + 128 <= g.HW && g.C == g.BN;       // uniform: no predicates in the common case
      const size_t base = (size_t)n * g.C * g.HW + (size_t)(valid ? p : 0);
      float xn[16];
      if (!g.has_q) {                            // first chunk's x: in flight while the tile's MMAs finish
#pragma unroll
        for (int j = 0; j < 16; ++j)
          xn[j] = (valid && eh * 16 + j < g.C) ? __ldg(x + base + (size_t)(eh * 16 + j) * g.HW) : 0.f;
      } else {
        mbar_wait_long(q_full + 8u * acc, acc_ph[acc]);
      }
      mbar_wait_long(t_full + 8u * acc, acc_ph[acc]);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * g.acc_cols);
      const uint32_t qrow = q_buf + (uint32_t)acc * q_tile + (uint32_t)row;
      for (int c0 = eh * 16; c0 < g.BN; c0 += 32) {
        uint32_t v[16];
        float xv[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        if (g.has_q) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            uint32_t code;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(code) : "r"(qrow + (uint32_t)(c0 + j) * 128u));
            xv[j] = lds_f32(lut_a + (code << 2));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) xv[j] = xn[j];
          if (c0 + 32 < g.BN) {                  // next chunk's x: overlaps this chunk's arithmetic and stores
#pragma unroll
            for (int j = 0; j < 16; ++j)
              xn[j] = (valid && c0 + 32 + j < g.C) ? __ldg(x + base + (size_t)(c0 + 32 + j) * g.HW) : 0.f;
          }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float nrm[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 pr = lds_f32x4(par_a + (uint32_t)(c0 + j) * 16u);
          if (g.has_q) xv[j] = __fadd_rn(__fmul_rn(xv[j], pr.z), pr.y);
          nrm[j] = __uint_as_float(v[j]) + pr.x;
        }
        if (g.inverse) {
#pragma unroll
          for (int j = 0; j < 16; ++j) xv[j] *= sqrt_fast(nrm[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) xv[j] *= rsqrt_fast(nrm[j]);
        }
        if (stats != nullptr) {
          unsigned kmn, kmx;
          warp_channel_minmax16(xv, valid, lane, kmn, kmx);
          if (lane < 16 && c0 + lane < g.C) {
            atomicMin(s_stat + c0 + lane, kmn);
            atomicMax(s_stat + 256 + c0 + lane, kmx);
          }
        }
        float* yp = y + base + (size_t)c0 * g.HW;
        if (full) {
#pragma unroll
          for (int j = 0; j < 16; ++j) __stcs(yp + (size_t)j * g.HW, xv[j]);   // streaming: y is not read again here
        } else if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < g.C) __stcs(yp + (size_t)j * g.HW, xv[j]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(t_empty + 8u * acc);
      if (g.has_q) mbar_arrive(q_empty + 8u * acc);
      acc_ph[acc] ^= 1u;
      acc ^= 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (stats != nullptr) {
    for (int c = threadIdx.x; c < g.C; c += kGdThreads) {
      if (s_stat[c] != 0xffffffffu) atomicMin(stats + 2 * c, s_stat[c]);
      if (s_stat[256 + c] != 0u) atomicMax(stats + 2 * c + 1, s_stat[256 + c]);
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t cols = 2u * (uint32_t)g.acc_cols;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols) : "memory");
  }
}

// Counts (a, b) pairs for which div_rn differs from __fdiv_rn: a in [0, b] (the quantiser's first division, b = channel
// range) and the exhaustive code / levels table of every bit width (its second).
__global__ void fast_div_selftest_kernel(unsigned long long n, unsigned long long seed, unsigned long long* bad) {
  using namespace gd;
  unsigned long long local = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long h = (i + seed) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    const uint32_t u0 = (uint32_t)h, u1 = (uint32_t)(h >> 32);
    // b: random significand, exponent in [-20, 20]; a = b * fraction in [0, 1]
    const float b = __uint_as_float(((127u - 20u + (u1 >> 23) % 41u) << 23) | (u1 & 0x7fffffu));
    const float a = __fmul_rn(b, (float)(u0 >> 8) * (1.f / 16777216.f));
    if (!div_rn_ok(b)) continue;
    if (__float_as_uint(div_rn(a, b, __frcp_rn(b))) != __float_as_uint(__fdiv_rn(a, b))) ++local;
  }
  if (blockIdx.x == 0) {
    for (int bits = 2; bits <= 16; ++bits) {
      const float L = (float)((1 << bits) - 1);
      for (int q = threadIdx.x; q <= (1 << bits) - 1; q += blockDim.x)
        if (__float_as_uint(div_rn((float)q, L, __frcp_rn(L))) != __float_as_uint(__fdiv_rn((float)q, L))) ++local;
    }
  }
  if (local) atomicAdd(bad, local);
}

}  // namespace b200lic

using namespace b200lic;

extern "C" {

// 1 when b200lic_gdn_fwd_fused accepts the shape (the caller otherwise stages the operand and uses the conv engine)
int b200lic_gdn_fused_ok(int C, int HW) {
  int cpad, coutpad, tmax, phases;
  size_t bb;
  if (C < 1 || HW < 1 || (HW & 3)) return 0;
  if (!tc2_weight_layout(C, C, 1, 1, 1, 0, &cpad, &coutpad, &tmax, &phases, &bb)) return 0;
  return (coutpad <= 256 && cpad <= 256) ? 1 : 0;
}

int b200lic_gdn_fwd_fused(const float* x, const float* minmax, int n_bits, const void* packed_gamma, const float* beta,
                          int N, int C, int HW, int inverse, float* y, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && packed_gamma && beta && y && N > 0 && C > 0 && HW > 0, "gdn_fwd_fused: bad arguments");
  B200_REQUIRE(!minmax || (n_bits >= 2 && n_bits <= 8), "gdn_fwd_fused: on-chip quantiser holds 1-byte codes (n_bits=%d)", n_bits);
  if (!b200lic_gdn_fused_ok(C, HW) || ((((uintptr_t)x) | ((uintptr_t)y)) & 15) || (((uintptr_t)packed_gamma) & 127)) {
    set_error("gdn_fwd_fused: shape or alignment not eligible (C=%d HW=%d)", C, HW);
    return B200LIC_ERR_UNSUPPORTED;
  }
  int cpad, coutpad, tmax, phases;
  size_t bb;
  tc2_weight_layout(C, C, 1, 1, 1, 0, &cpad, &coutpad, &tmax, &phases, &bb);
  GdnGeom g{};
  g.N = N;
  g.C = C;
  g.HW = HW;
  g.nkb = cpad / 32;
  g.BN = coutpad;
  g.tiles_per_img = (HW + 127) / 128;
  const long long tiles = (long long)N * g.tiles_per_img;
  B200_REQUIRE(tiles < 2147483647LL, "gdn_fwd_fused: tensor too large");
  g.n_tiles = (int)tiles;
  g.inverse = inverse ? 1 : 0;
  g.has_q = minmax ? 1 : 0;
  const uint32_t b_stage = 2u * (uint32_t)g.BN * 64u;
  const size_t q_bytes = g.has_q ? 2 * (size_t)cpad * 128 : 0;      // two code tiles
  size_t fixed = 0;
  // two A/B stages are enough (a K block's six MMAs are much shorter than its conversion); the rest of the shared
  // memory goes to the X ring, which hides the HBM latency
  g.SA = 2;
  fixed = 1024 + (size_t)g.SA * (kGdAStage + b_stage) + q_bytes + 7 * 256 * 4 + 512;
  int sx = fixed < 227 * 1024 ? (int)((227 * 1024 - fixed) / kGdXStage) : 0;
  g.SX = sx > kGdMaxSX ? kGdMaxSX : sx;
  if (g.SX < 3) {
    set_error("gdn_fwd_fused: shared memory budget (C=%d)", C);
    return B200LIC_ERR_UNSUPPORTED;
  }
  g.acc_cols = 32;
  while (g.acc_cols < g.BN) g.acc_cols *= 2;
  g.L = (float)((1 << (minmax ? n_bits : 8)) - 1);
  const size_t smem = fixed + (size_t)g.SX * kGdXStage;

  CUtensorMap mx, mb;
  {
    cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t strides[2] = {(cuuint64_t)HW * 4, (cuuint64_t)HW * C * 4};
    cuuint32_t box[3] = {128, 32, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (!tc_encode_map_ex(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_NONE, const_cast<float*>(x), 3, dims,
                          strides, box, es))
      return B200LIC_ERR_CUDA;
    cuuint64_t bdims[4] = {(cuuint64_t)cpad, (cuuint64_t)coutpad, 1, 2};
    cuuint64_t bstrides[3] = {(cuuint64_t)cpad * 2, (cuuint64_t)cpad * 2 * coutpad, (cuuint64_t)bb};
    cuuint32_t bbox[4] = {32, (cuuint32_t)coutpad, 1, 2};
    cuuint32_t bes[4] = {1, 1, 1, 1};
    if (!tc_encode_map_ex(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B, const_cast<void*>(packed_gamma), 4,
                          bdims, bstrides, bbox, bes))
      return B200LIC_ERR_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gdn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("gdn_fwd_fused: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
      return B200LIC_ERR_CUDA;
    }
    attr_set = true;
  }
  const int sms = num_sms();
  const int grid = g.n_tiles < sms ? g.n_tiles : sms;
  unsigned* stats = tc2_stats_peek();            // a pending b200lic_conv_stats_once request: this launch honours it
  if (stats) tc2_stats_once(nullptr);
  launch_pdl(gdn_fused_kernel, dim3(grid), dim3(kGdThreads), smem, as_stream(stream), mx, mb, g, x, reinterpret_cast<const unsigned*>(minmax),
                                                                  beta, y, stats);
  B200_LAUNCH_CHECK("gdn_fused_kernel");
  return B200LIC_OK;
}

// Test hook: number of mismatches of the FMA division against __fdiv_rn over n pseudo-random pairs + all code tables.
int b200lic_selftest_fast_div(unsigned long long n, unsigned long long seed, unsigned long long* mismatches_dev,
                              b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(mismatches_dev, "selftest_fast_div: null pointer");
  fast_div_selftest_kernel<<<num_sms() * 8, 256, 0, as_stream(stream)>>>(n, seed, mismatches_dev);
  B200_LAUNCH_CHECK("fast_div_selftest_kernel");
  return B200LIC_OK;
}

}  // extern "C"
