// tcgen05 tensor-core implicit-GEMM engine (B200LIC_ENGINE_TC) for conv / transposed-conv forward and dgrad.
//
// Formulation (same gather geometry as conv_simt.cu, so both engines are interchangeable and cross-checked):
//   D[pixel, co] = sum_{tap, ci} A[pixel, (tap, ci)] * B[co, (tap, ci)]        M = 128 output pixels, N = BN channels
// * operands are staged ONCE per call into tensor-core friendly form by two small HBM-bound kernels:
//     activations  NCHW fp32 -> NHWC bf16 "hi" and "lo" slices (x = hi + lo + O(2^-17 |x|)), channels padded to 64
//     weights      -> [phase][Cout][tap][ci] bf16 hi/lo slices, K-major
//   three MMA passes per k-block (hi*hi + hi*lo + lo*hi) accumulate in fp32 in TMEM: ~2^-16 relative product error,
//   inside the 1e-4 per-layer bar while running on the bf16 tensor pipe (plain bf16/TF32 would not meet it);
// * the A tile of one filter tap is ONE 4-D TMA box (64 ch x BW x BH x BI pixels) over the NHWC tensor: conv stride
//   becomes the TMA element stride, padding becomes TMA out-of-bounds zero fill -- there is no im2col buffer;
// * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
//   warps 2..5 = epilogue (tcgen05.ld -> bias / GDN tail / activation / Q8.8 -> NCHW fp32 stores);
//   smem ring of S stages with full/empty mbarriers, accumulator hand-off through a tmem_full mbarrier.
// Replaces cuDNN under F.conv2d / F.conv_transpose2d (TO quant_layer.py:28,36,123) and their dgrad.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"

namespace b200lic {

// ---------------------------------------------------------------------------------------------------------------------
// operand staging kernels
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// x [N,C,HW] fp32 -> xh/xl [N,HW,Cpad] bf16 (pad channels zero).  One CTA transposes a 64-channel x 64-pixel tile
// through shared memory: reads are 256 B contiguous per warp (64 pixels of one channel), writes are one full 128 B
// NHWC row (64 channels x bf16) per warp for each of the hi and lo slices.
__global__ void __launch_bounds__(256) nhwc_split_kernel(const float* __restrict__ x, int C, int HW, int Cpad,
                                                          int square, __nv_bfloat16* __restrict__ xh,
                                                          __nv_bfloat16* __restrict__ xl) {
  __shared__ float t[64][65];
  // channel block fastest: the CTAs that complete one NHWC row (Cpad * 2 bytes) are co-resident, so the 128-byte pieces
  // they write merge into full rows in L2 instead of reaching DRAM one third of a row at a time
  const int cblk = (Cpad + 63) >> 6;
  const int n = blockIdx.z, c0 = (int)(blockIdx.x % cblk) * 64, p0 = (int)(blockIdx.x / cblk) * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 8 warps
  const float* xn = x + (size_t)n * C * HW;
  const bool vec2 = (HW & 1) == 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int cl = warp + 8 * j, c = c0 + cl, p = p0 + 2 * lane;
    float v0 = 0.f, v1 = 0.f;
    if (c < C) {
      const float* src = xn + (size_t)c * HW + p;
      if (vec2 && p + 1 < HW) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(src));
        v0 = v.x;
        v1 = v.y;
      } else {
        if (p < HW) v0 = __ldg(src);
        if (p + 1 < HW) v1 = __ldg(src + 1);
      }
    }
    if (square) {
      v0 *= v0;
      v1 *= v1;
    }
    t[cl][2 * lane] = v0;
    t[cl][2 * lane + 1] = v1;
  }
  __syncthreads();
  if (c0 + 2 * lane >= Cpad) return;               // Cpad is a multiple of 64, so this only trims nothing; kept for safety
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int pl = warp + 8 * j, p = p0 + pl;
    if (p < HW) {
      const float a = t[2 * lane][pl], b = t[2 * lane + 1][pl];
      __nv_bfloat16 ah, al, bh, bl;
      split_bf16(a, ah, al);
      split_bf16(b, bh, bl);
      const size_t o = ((size_t)n * HW + p) * Cpad + c0 + 2 * lane;
      __nv_bfloat162 hv, lv;
      hv.x = ah; hv.y = bh;
      lv.x = al; lv.y = bl;
      *reinterpret_cast<__nv_bfloat162*>(xh + o) = hv;
      *reinterpret_cast<__nv_bfloat162*>(xl + o) = lv;
    }
  }
}

// weights -> [phase][co][tap][ci] bf16 hi/lo.  Source element (co, ci, r, s) at co*s_co + ci*s_ci + r*KW + s.
struct PackGeom {
  int Cout, Cin, KH, KW, stride, pad, transposed;
  int CoutPad, Cpad, Tmax;  // Kmax = Tmax * Cpad
  long long s_co, s_ci;
};

__global__ void __launch_bounds__(256) pack_weights_kernel(PackGeom g, const float* __restrict__ w,
                                                            __nv_bfloat16* __restrict__ bh,
                                                            __nv_bfloat16* __restrict__ bl) {
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
    float v = 0.f;
    if (co < g.Cout && ci < g.Cin && t < T) {
      const int i = t / KWp, j = t - i * KWp;
      v = __ldg(w + (long long)co * g.s_co + (long long)ci * g.s_ci + (r0 + i * rstep) * g.KW + (s0 + j * rstep));
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    bh[(size_t)phase * per_phase + e] = h;
    bl[(size_t)phase * per_phase + e] = l;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
// same, for warps that wait through a whole main loop: sleep between probes instead of hammering the barrier
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(256);
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
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

// K-major, SWIZZLE_128B shared-memory matrix descriptor (8-row groups of 1024 B; rows of 128 B = 64 bf16).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

// ---------------------------------------------------------------------------------------------------------------------
// the GEMM kernel
// ---------------------------------------------------------------------------------------------------------------------
struct TcGeom {
  int N, H, W, Cpad;          // gathered NHWC tensor (channels padded to 64)
  int Cout, Ho, Wo;           // written NCHW tensor
  int KH, KW, stride, pad, transposed;
  int BW, BH, BI;             // pixel box of one M tile: BW*BH*BI == 128
  int BN, n_tiles;            // output-channel tile and their count
  int stages, tmem_cols;
  int act;
  float slope;
  int gdn_mode, fixed_point;
};

constexpr int kTcThreads = 192;
constexpr int kATileBytes = 128 * 128;  // 128 pixel rows x 64 bf16

__global__ void __launch_bounds__(kTcThreads, 1)
    tc_gather_gemm_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                          const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, TcGeom g,
                          const float* __restrict__ bias, const float* __restrict__ gdn_x, float* __restrict__ norm_out,
                          float* __restrict__ y) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- phase geometry (mirrors conv_simt.cu) --------------------------------------------------------------------
  int ph = 0, pw = 0, KHp = g.KH, KWp = g.KW, Pa = g.Ho, Pb = g.Wo;
  int in_step = g.stride, tap_step = 1, base_h = -g.pad, base_w = -g.pad, out_step = 1;
  if (g.transposed) {
    const int st = g.stride;
    ph = blockIdx.z / st;
    pw = blockIdx.z % st;
    const int r0 = (ph + g.pad) % st, s0 = (pw + g.pad) % st;
    KHp = r0 < g.KH ? (g.KH - r0 + st - 1) / st : 0;
    KWp = s0 < g.KW ? (g.KW - s0 + st - 1) / st : 0;
    Pa = ph < g.Ho ? (g.Ho - ph + st - 1) / st : 0;
    Pb = pw < g.Wo ? (g.Wo - pw + st - 1) / st : 0;
    in_step = 1;
    tap_step = -1;
    base_h = (ph + g.pad - r0) / st;
    base_w = (pw + g.pad - s0) / st;
    out_step = st;
  }
  const int T = KHp * KWp;
  const int cblocks = g.Cpad >> 6;
  const int num_kb = T * cblocks;
  const int tiles_w = (Pb + g.BW - 1) / g.BW, tiles_h = (Pa + g.BH - 1) / g.BH;
  const int tiles_n = (g.N + g.BI - 1) / g.BI;
  const int m_tile = blockIdx.x;
  if (Pa <= 0 || Pb <= 0 || m_tile >= tiles_w * tiles_h * tiles_n) return;   // uniform per CTA
  const int tw = m_tile % tiles_w, th = (m_tile / tiles_w) % tiles_h, tn = m_tile / (tiles_w * tiles_h);
  const int a0 = th * g.BH, b0 = tw * g.BW, n0 = tn * g.BI;
  const int n_tile = blockIdx.y;

  // ---- shared memory carve-up -------------------------------------------------------------------------------------
  const uint32_t b_tile_bytes = (uint32_t)g.BN * 128u;
  const uint32_t stage_bytes = 2u * kATileBytes + 2u * b_tile_bytes;
  const uint32_t bars = smem_base + (uint32_t)g.stages * stage_bytes;       // full[S], empty[S], tmem_full, tmem_ptr
  const uint32_t full_bar = bars, empty_bar = bars + 8u * g.stages, tmem_full_bar = bars + 16u * g.stages;
  const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(full_bar + 8u * s, 1);
      mbar_init(empty_bar + 8u * s, 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"((uint32_t)g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ===== TMA producer ===============================================================================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_ah)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_bh)) : "memory");
      const int w_base = b0 * in_step + base_w, h_base = a0 * in_step + base_h;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % g.stages;
        const uint32_t round = (uint32_t)(kb / g.stages);
        mbar_wait(empty_bar + 8u * s, (round & 1u) ^ 1u);
        const int t = kb / cblocks, cb = kb - t * cblocks;
        const int i = t / KWp, j = t - i * KWp;
        const uint32_t st_base = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t fb = full_bar + 8u * s;
        mbar_expect_tx(fb, stage_bytes);
        const int cw = w_base + j * tap_step, ch = h_base + i * tap_step;
        tma_load_4d(st_base, &map_ah, fb, cb * 64, cw, ch, n0);
        tma_load_4d(st_base + kATileBytes, &map_al, fb, cb * 64, cw, ch, n0);
        tma_load_3d(st_base + 2u * kATileBytes, &map_bh, fb, kb * 64, n_tile * g.BN, (int)blockIdx.z);
        tma_load_3d(st_base + 2u * kATileBytes + b_tile_bytes, &map_bl, fb, kb * 64, n_tile * g.BN, (int)blockIdx.z);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====================================================================================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.BN >> 3) << 17) | ((128u >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % g.stages;
        const uint32_t round = (uint32_t)(kb / g.stages);
        mbar_wait(full_bar + 8u * s, round & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st_base = smem_base + (uint32_t)s * stage_bytes;
        const uint64_t ah = make_kmajor_sw128_desc(st_base), al = make_kmajor_sw128_desc(st_base + kATileBytes);
        const uint64_t bh = make_kmajor_sw128_desc(st_base + 2u * kATileBytes);
        const uint64_t bl = make_kmajor_sw128_desc(st_base + 2u * kATileBytes + b_tile_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x UMMA_K(16) = 64; +32 B per step inside the 128 B swizzle row
          const uint64_t o = (uint64_t)(k * 2);
          umma_bf16(tmem_base, ah + o, bh + o, idesc, (kb | k) != 0);
          umma_bf16(tmem_base, ah + o, bl + o, idesc, 1u);
          umma_bf16(tmem_base, al + o, bh + o, idesc, 1u);
        }
        umma_commit(empty_bar + 8u * s);        // frees the smem stage when these MMAs retire
      }
      umma_commit(tmem_full_bar);               // accumulator complete
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =========================================================
    const int q = warp & 3;
    const int m = q * 32 + lane;                                    // row of the tile = pixel
    const int iw = m % g.BW, ih = (m / g.BW) % g.BH, ii = m / (g.BW * g.BH);
    const int a = a0 + ih, b = b0 + iw, n = n0 + ii;
    const bool valid = a < Pa && b < Pb && n < g.N;
    const int ho = a * out_step + ph, wo = b * out_step + pw;
    const long long pix = (long long)ho * g.Wo + wo;
    const long long plane = (long long)g.Ho * g.Wo;
    if (num_kb > 0) {
      mbar_wait_backoff(tmem_full_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int co_base = n_tile * g.BN;
    const long long obase = ((long long)n * g.Cout + co_base) * plane + pix;   // element (n, co_base, ho, wo)
    const bool gdn = g.gdn_mode != 0 && valid;
    // GDN operand x for the 16 channels of a chunk: all 16 loads are issued together (and one chunk ahead of their use),
    // so the epilogue keeps 16-32 requests per thread in flight instead of one dependent round trip per channel.
    float xn[16];
    auto load_x = [&](int c0) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        xn[j] = (gdn && co_base + c0 + j < g.Cout) ? __ldg(gdn_x + obase + (long long)(c0 + j) * plane) : 0.f;
    };
    if (gdn) load_x(0);
    for (int c0 = 0; c0 < g.BN; c0 += 16) {
      uint32_t v[16];
      float xc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) xc[j] = xn[j];
      if (gdn && c0 + 16 < g.BN) load_x(c0 + 16);
      if (num_kb > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int co = co_base + c0 + j;
          if (co < g.Cout) {
            const long long idx = obase + (long long)(c0 + j) * plane;
            float r = __uint_as_float(v[j]) + (bias ? __ldg(bias + co) : 0.f);
            if (g.gdn_mode) {
              if (norm_out) norm_out[idx] = r;
              r = g.gdn_mode == 1 ? xc[j] * rsqrtf(r) : xc[j] * sqrtf(r);
            }
            r = apply_act(r, g.act, g.slope);
            if (g.fixed_point) r = rintf(fminf(fmaxf(r, -128.f), 128.f) * 256.f) * (1.f / 256.f);
            y[idx] = r;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}
static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p *= 2;
  return p;
}

struct TcPlan {
  bool ok = false;
  int Cpad, CoutPad, Tmax, phases, BN, n_tiles, BW, BH, BI, stages, tmem_cols, m_tiles;
  size_t x_bytes, b_bytes, total_bytes, smem_bytes;
};

// written tensor [N,Cout,Ho,Wo]; gathered tensor [N,Cin,H,W]
static TcPlan make_plan(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride,
                        int transposed) {
  TcPlan p;
  if (Cin < 1 || Cout < 1) return p;
  const int st = transposed ? stride : 1;
  p.phases = st * st;
  p.Tmax = transposed ? ((KH + st - 1) / st) * ((KW + st - 1) / st) : KH * KW;
  if (p.Tmax < 1 || p.Tmax > 64) return p;
  p.Cpad = (Cin + 63) / 64 * 64;
  const int c16 = (Cout + 15) / 16 * 16;
  p.BN = 0;
  for (int bn = 256; bn >= 16; bn -= 16)
    if (c16 % bn == 0) {
      p.BN = bn;
      break;
    }
  if (c16 > 256 && p.BN < 96) {                              // awkward factorisation: pad to a multiple of 128 instead
    p.BN = 128;
    p.CoutPad = (Cout + 127) / 128 * 128;
  } else {
    p.CoutPad = c16;
  }
  p.n_tiles = p.CoutPad / p.BN;
  const int Pa = (Ho + st - 1) / st, Pb = (Wo + st - 1) / st;  // largest phase
  p.BW = Pb >= 16 ? 16 : pow2_ceil(Pb);
  const int es = transposed ? 1 : stride;
  if (p.BW * es > 256) return p;
  p.BH = 128 / p.BW;
  if (p.BH > pow2_ceil(Pa)) p.BH = pow2_ceil(Pa);
  if (p.BH * es > 256) return p;
  p.BI = 128 / (p.BW * p.BH);
  if (p.BI > 256) return p;
  p.m_tiles = ((Pb + p.BW - 1) / p.BW) * ((Pa + p.BH - 1) / p.BH) * ((N + p.BI - 1) / p.BI);
  const size_t stage = 2 * (size_t)kATileBytes + 2 * (size_t)p.BN * 128;
  p.stages = (int)((227 * 1024 - 2048) / stage);
  if (p.stages > 6) p.stages = 6;
  if (p.stages < 2) return p;
  p.smem_bytes = (size_t)p.stages * stage + 1024 /*align*/ + 256 /*barriers*/;
  p.tmem_cols = p.BN <= 32 ? 32 : pow2_ceil(p.BN);
  p.x_bytes = ((size_t)N * H * W * p.Cpad * 2 + 1023) / 1024 * 1024;
  p.b_bytes = ((size_t)p.phases * p.CoutPad * p.Tmax * p.Cpad * 2 + 1023) / 1024 * 1024;
  p.total_bytes = 2 * p.x_bytes + 2 * p.b_bytes + 1024;
  p.ok = true;
  return p;
}

static bool encode_map(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                       const cuuint32_t* box, const cuuint32_t* estr) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return false;
  }
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return false;
  }
  return true;
}

// second-generation engine (conv_tc2.cu); B200LIC_TC_V1=1 in the environment keeps the first one for A/B runs
size_t tc2_workspace_bytes(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride,
                           int transposed);
int tc2_launch(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride, int pad,
               int transposed, long long s_co, long long s_ci, int act, float slope, int in_square, int gdn_mode,
               int fixed_point, const float* x, const float* w, const float* bias, const float* gdn_x, float* norm_out,
               float* y, void* workspace, size_t workspace_bytes, cudaStream_t s, const char* name);
static bool use_v1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200LIC_TC_V1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// Generic launcher.  (N,Cin,H,W) gathered tensor, (Cout,Ho,Wo) written tensor, weight strides of the written /
// gathered channel axes.
static int tc_launch(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride, int pad,
                     int transposed, long long s_co, long long s_ci, int act, float slope, int in_square, int gdn_mode,
                     int fixed_point, const float* x, const float* w, const float* bias, const float* gdn_x,
                     float* norm_out, float* y, void* workspace, size_t workspace_bytes, cudaStream_t s,
                     const char* name) {
  if (!use_v1())
    return tc2_launch(N, Cin, H, W, Cout, Ho, Wo, KH, KW, stride, pad, transposed, s_co, s_ci, act, slope, in_square,
                      gdn_mode, fixed_point, x, w, bias, gdn_x, norm_out, y, workspace, workspace_bytes, s, name);
  TcPlan p = make_plan(N, Cin, H, W, Cout, Ho, Wo, KH, KW, stride, transposed);
  if (!p.ok) {
    set_error("%s: shape not eligible for the tcgen05 engine", name);
    return B200LIC_ERR_UNSUPPORTED;
  }
  if (!workspace || workspace_bytes < p.total_bytes) {
    set_error("%s: tcgen05 engine needs %zu workspace bytes (got %zu)", name, p.total_bytes, workspace_bytes);
    return B200LIC_ERR_UNSUPPORTED;
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* xl = reinterpret_cast<__nv_bfloat16*>(ws + p.x_bytes);
  __nv_bfloat16* bh = reinterpret_cast<__nv_bfloat16*>(ws + 2 * p.x_bytes);
  __nv_bfloat16* bl = reinterpret_cast<__nv_bfloat16*>(ws + 2 * p.x_bytes + p.b_bytes);

  // 1. stage operands
  {
    const int HW = H * W;
    dim3 grid(((p.Cpad + 63) / 64) * ((HW + 63) / 64), 1, N);
    nhwc_split_kernel<<<grid, 256, 0, s>>>(x, Cin, HW, p.Cpad, in_square, xh, xl);
    B200_LAUNCH_CHECK("nhwc_split_kernel");
    PackGeom pg{Cout, Cin, KH, KW, stride, pad, transposed, p.CoutPad, p.Cpad, p.Tmax, s_co, s_ci};
    const size_t per_phase = (size_t)p.CoutPad * p.Tmax * p.Cpad;
    dim3 pgrid((unsigned)((per_phase + 255) / 256 > 1184 ? 1184 : (per_phase + 255) / 256), 1, p.phases);
    pack_weights_kernel<<<pgrid, 256, 0, s>>>(pg, w, bh, bl);
    B200_LAUNCH_CHECK("pack_weights_kernel");
  }
  // 2. tensor maps
  CUtensorMap mah, mal, mbh, mbl;
  {
    const int es = transposed ? 1 : stride;
    cuuint64_t dims[4] = {(cuuint64_t)p.Cpad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)p.Cpad * 2, (cuuint64_t)W * p.Cpad * 2, (cuuint64_t)H * W * p.Cpad * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(p.BW * es), (cuuint32_t)(p.BH * es), (cuuint32_t)p.BI};
    cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
    if (!encode_map(&mah, xh, 4, dims, strides, box, estr) || !encode_map(&mal, xl, 4, dims, strides, box, estr))
      return B200LIC_ERR_CUDA;
    const cuuint64_t Kmax = (cuuint64_t)p.Tmax * p.Cpad;
    cuuint64_t bdims[3] = {Kmax, (cuuint64_t)p.CoutPad, (cuuint64_t)p.phases};
    cuuint64_t bstrides[2] = {Kmax * 2, Kmax * 2 * (cuuint64_t)p.CoutPad};
    cuuint32_t bbox[3] = {64, (cuuint32_t)p.BN, 1};
    cuuint32_t bestr[3] = {1, 1, 1};
    if (!encode_map(&mbh, bh, 3, bdims, bstrides, bbox, bestr) || !encode_map(&mbl, bl, 3, bdims, bstrides, bbox, bestr))
      return B200LIC_ERR_CUDA;
  }
  // 3. GEMM
  TcGeom g{N, H, W, p.Cpad, Cout, Ho, Wo, KH, KW, stride, pad, transposed, p.BW, p.BH, p.BI, p.BN, p.n_tiles,
           p.stages, p.tmem_cols, act, slope, gdn_mode, fixed_point};
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_gather_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("%s: cannot raise dynamic shared memory: %s", name, cudaGetErrorString(e));
      return B200LIC_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid(p.m_tiles, p.n_tiles, p.phases);
  tc_gather_gemm_kernel<<<grid, kTcThreads, p.smem_bytes, s>>>(mah, mal, mbh, mbl, g, bias, gdn_x, norm_out, y);
  B200_LAUNCH_CHECK(name);
  return B200LIC_OK;
}

// shared with conv_tc_wgrad.cu
int tc_stage_nhwc(const float* x, int N, int C, int HW, int Cpad, int square, void* xh, void* xl, cudaStream_t s) {
  dim3 grid(((Cpad + 63) / 64) * ((HW + 63) / 64), 1, N);
  nhwc_split_kernel<<<grid, 256, 0, s>>>(x, C, HW, Cpad, square, reinterpret_cast<__nv_bfloat16*>(xh),
                                         reinterpret_cast<__nv_bfloat16*>(xl));
  B200_LAUNCH_CHECK("nhwc_split_kernel");
  return B200LIC_OK;
}
bool tc_encode_map(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, const cuuint32_t* estr) {
  return encode_map(m, base, rank, dims, strides_bytes, box, estr);
}

int tc_pack_weights(const float* w, int Cout, int Cin, int KH, int KW, int stride, int pad, int transposed, int CoutPad,
                    int Cpad, int Tmax, int phases, long long s_co, long long s_ci, void* bh, void* bl, cudaStream_t s) {
  PackGeom pg{Cout, Cin, KH, KW, stride, pad, transposed, CoutPad, Cpad, Tmax, s_co, s_ci};
  const size_t per_phase = (size_t)CoutPad * Tmax * Cpad;
  dim3 pgrid((unsigned)((per_phase + 255) / 256 > 1184 ? 1184 : (per_phase + 255) / 256), 1, phases);
  pack_weights_kernel<<<pgrid, 256, 0, s>>>(pg, w, reinterpret_cast<__nv_bfloat16*>(bh),
                                            reinterpret_cast<__nv_bfloat16*>(bl));
  B200_LAUNCH_CHECK("pack_weights_kernel");
  return B200LIC_OK;
}
bool tc_encode_map_ex(CUtensorMap* m, CUtensorMapDataType dt, CUtensorMapSwizzle sw, void* base, int rank,
                      const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                      const cuuint32_t* estr) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return false;
  }
  CUresult r = fn(m, dt, (cuuint32_t)rank, base, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, box %u x %u)", (int)r, rank, box[0], box[1]);
    return false;
  }
  return true;
}

size_t tc_workspace_bytes(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride,
                          int transposed) {
  if (!use_v1()) return tc2_workspace_bytes(N, Cin, H, W, Cout, Ho, Wo, KH, KW, stride, transposed);
  TcPlan p = make_plan(N, Cin, H, W, Cout, Ho, Wo, KH, KW, stride, transposed);
  return p.ok ? p.total_bytes : 0;
}

size_t smallc_conv_fwd_ws(const b200lic_conv_desc* d);
size_t smallc_deconv_fwd_ws(const b200lic_conv_desc* d);
int smallc_conv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void* ws,
                    size_t ws_bytes, cudaStream_t s);
int smallc_deconv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void* ws,
                      size_t ws_bytes, cudaStream_t s);
// Where the forward call left the split-bf16 NHWC copy of x inside its workspace, if the wgrad engine can consume it
// as is (channels a multiple of 64, generic tap-by-tap path of the second-generation engine).
bool tc_staged_view(const b200lic_conv_desc* d, int transposed, void* fwd_ws, size_t ws_bytes, void** hi, void** lo) {
  *hi = *lo = nullptr;
  if (use_v1() || !fwd_ws || (d->Cin % 64) != 0) return false;
  if ((transposed ? smallc_deconv_fwd_ws(d) : smallc_conv_fwd_ws(d)) != 0) return false;
  const size_t need = tc2_workspace_bytes(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride,
                                          transposed);
  if (need == 0 || ws_bytes < need) return false;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)fwd_ws + 1023) & ~(uintptr_t)1023);
  const size_t x_bytes = ((size_t)d->N * d->H * d->W * d->Cin * 2 + 1023) / 1024 * 1024;
  *hi = base;
  *lo = base + x_bytes;
  return true;
}

size_t tc_conv_fwd_ws(const b200lic_conv_desc* d) {
  if (!use_v1())
    if (const size_t n = smallc_conv_fwd_ws(d)) return n;
  return tc_workspace_bytes(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, 0);
}
size_t tc_deconv_fwd_ws(const b200lic_conv_desc* d) {
  if (!use_v1())
    if (const size_t n = smallc_deconv_fwd_ws(d)) return n;
  return tc_workspace_bytes(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, 1);
}

int tc_conv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, const float* gdn_x,
                float* norm_out, float* y, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (!use_v1() && smallc_conv_fwd_ws(d) != 0) return smallc_conv_fwd(d, x, w, bias, y, ws, ws_bytes, s);
  return tc_launch(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 0,
                   (long long)d->Cin * d->KH * d->KW, (long long)d->KH * d->KW, d->act, d->act_slope, d->in_square,
                   d->gdn_mode, d->fixed_point, x, w, bias, gdn_x, norm_out, y, ws, ws_bytes, s, "conv_fwd(tc)");
}

int tc_deconv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void* ws,
                  size_t ws_bytes, cudaStream_t s) {
  if (!use_v1() && smallc_deconv_fwd_ws(d) != 0) return smallc_deconv_fwd(d, x, w, bias, y, ws, ws_bytes, s);
  return tc_launch(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 1,
                   (long long)d->KH * d->KW, (long long)d->Cout * d->KH * d->KW, d->act, d->act_slope, 0, 0,
                   d->fixed_point, x, w, bias, nullptr, nullptr, y, ws, ws_bytes, s, "deconv_fwd(tc)");
}

int tc2_launch_wq(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride, int pad,
                  int transposed, long long s_co, long long s_ci, int act, float slope, int in_square, int gdn_mode,
                  int fixed_point, const float* x, const float* w, const float* w_scale, const float* bias,
                  const float* gdn_x, float* norm_out, float* y, void* workspace, size_t workspace_bytes, cudaStream_t s,
                  const char* name);
// integer-valued weights + per-output-channel scale (two MMA passes); the folded-tap path of the 3-channel layers keeps
// the three-pass engine (its scale would have to be expanded per folded column)
int tc_conv_fwd_wq(const b200lic_conv_desc* d, const float* x, const float* w_int, const float* w_scale, const float* bias,
                   float* y, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (use_v1() || smallc_conv_fwd_ws(d) != 0) {
    set_error("conv_fwd_wq: shape runs on the folded-tap path");
    return B200LIC_ERR_UNSUPPORTED;
  }
  return tc2_launch_wq(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 0,
                       (long long)d->Cin * d->KH * d->KW, (long long)d->KH * d->KW, d->act, d->act_slope, 0, 0,
                       d->fixed_point, x, w_int, w_scale, bias, nullptr, nullptr, y, ws, ws_bytes, s, "conv_fwd_wq(tc)");
}
int tc_deconv_fwd_wq(const b200lic_conv_desc* d, const float* x, const float* w_int, const float* w_scale,
                     const float* bias, float* y, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (use_v1() || smallc_deconv_fwd_ws(d) != 0) {
    set_error("deconv_fwd_wq: shape runs on the folded-tap path");
    return B200LIC_ERR_UNSUPPORTED;
  }
  return tc2_launch_wq(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 1,
                       (long long)d->KH * d->KW, (long long)d->Cout * d->KH * d->KW, d->act, d->act_slope, 0, 0,
                       d->fixed_point, x, w_int, w_scale, bias, nullptr, nullptr, y, ws, ws_bytes, s,
                       "deconv_fwd_wq(tc)");
}

int tc_conv_dgrad(const b200lic_conv_desc* d, const float* dy, const float* w, float* dx, void* ws, size_t ws_bytes,
                  cudaStream_t s) {
  return tc_launch(d->N, d->Cout, d->Ho, d->Wo, d->Cin, d->H, d->W, d->KH, d->KW, d->stride, d->pad, 1,
                   (long long)d->KH * d->KW, (long long)d->Cin * d->KH * d->KW, 0, 0.f, 0, 0, 0, dy, w, nullptr, nullptr,
                   nullptr, dx, ws, ws_bytes, s, "conv_dgrad(tc)");
}

int tc_deconv_dgrad(const b200lic_conv_desc* d, const float* dy, const float* w, float* dx, void* ws, size_t ws_bytes,
                    cudaStream_t s) {
  return tc_launch(d->N, d->Cout, d->Ho, d->Wo, d->Cin, d->H, d->W, d->KH, d->KW, d->stride, d->pad, 0,
                   (long long)d->Cout * d->KH * d->KW, (long long)d->KH * d->KW, 0, 0.f, 0, 0, 0, dy, w, nullptr,
                   nullptr, nullptr, dx, ws, ws_bytes, s, "deconv_dgrad(tc)");
}

}  // namespace b200lic
