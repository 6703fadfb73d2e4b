// tcgen05 tensor-core wgrad for conv / transposed conv (B200LIC_ENGINE_TC).
//
//   dW[cs, cb, r, s] = sum_{n,h,w} small[n, cs, h, w] * big[n, cb, h*st - pad + r, w*st - pad + s]
//   (conv: small = dY, big = x;  transposed conv: small = x, big = dY -- same roles as conv_simt.cu's wgrad)
//
// GEMM view per filter tap t:  D_t[cb, cs] = sum_pixels  Big_t[pixel, cb] * Small[pixel, cs]
// The contraction runs over PIXELS, which are the rows of the NHWC tiles TMA delivers, so both operands are fed to
// tcgen05.mma as MN-major (channel-contiguous) SWIZZLE_128B tiles: a [64 pixel x 64 channel] bf16 box is exactly one
// 64-wide MN block with K = 64.  Two (tap, channel-block) boxes stacked 8 KB apart form one M = 128 operand; the small
// tensor's channel blocks stacked 8 KB apart form N = 64*nb.  A CTA owns NACC = 2 accumulators (4 (tap, cb) boxes) in
// TMEM and streams the pixel tiles of its split through a 2-stage TMA ring; split-K partial sums are merged with fp32
// atomics into the zero-initialised dW.  Split-bf16 operands, 3 passes (hi*hi + hi*lo + lo*hi), fp32 accumulation.
// Replaces cuDNN wgrad under autograd of F.conv2d / F.conv_transpose2d (TO layer_opt.py:298-307).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"
#include "prepared.cuh"

namespace b200lic {

int tc_stage_nhwc(const float* x, int N, int C, int HW, int Cpad, int square, void* xh, void* xl, cudaStream_t s);
bool tc_encode_map(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, const cuuint32_t* estr);

namespace wg {

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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// one lane of a converged warp (see conv_tc2.cu: with `lane == 0` every UTMALDG / UTCHMMA is wrapped in an ELECT loop)
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

// MN-major SWIZZLE_128B descriptor: 64-element (128 B) MN rows, 8 K-rows per 1024 B atom;
// SBO = next 8-row K group, LBO = next 64-wide MN block.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

}  // namespace wg

constexpr int kWgThreads = 192;
constexpr int kBoxBytes = 64 * 128;   // 64 pixels x 64 bf16 channels
constexpr int kNacc = 2;              // accumulators (TMEM) per CTA
constexpr int kBoxesA = 2 * kNacc;    // (tap, channel-block) boxes per CTA

struct WgGeom {
  int N;
  int Cs, Hs, Ws, CsPad;     // small tensor (indexed directly)
  int Cb, Hb, Wb, CbPad;     // big tensor (gathered)
  int KH, KW, stride, pad;
  int BW, BH, BI;            // pixel box: BW*BH*BI == 64
  int nb_max;                // 64-channel blocks of the small tensor per N tile (<= 4)
  int tiles_per_split;       // pixel tiles per split
  int tmem_cols;             // columns per accumulator (power of two >= 64*nb_max)
};

__global__ void __launch_bounds__(kWgThreads, 1)
    tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_s, WgGeom g,
                    float* __restrict__ part) {
  pdl_wait();                 // programmatic dependent launch (common.cuh): predecessor complete, memory visible
  using namespace wg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int T = g.KH * g.KW;
  const int cblocks = g.CbPad >> 6;
  const int TT = T * cblocks;                         // (tap, channel-block) boxes in total
  const int unit = blockIdx.x;                        // 4 consecutive boxes
  const int sblocks = g.CsPad >> 6;
  const int nb0 = blockIdx.y * g.nb_max;              // first 64-block of the small tensor handled here
  const int nb = min(g.nb_max, sblocks - nb0);
  const int BN = nb * 64;
  const int tiles_w = (g.Ws + g.BW - 1) / g.BW, tiles_h = (g.Hs + g.BH - 1) / g.BH, tiles_n = (g.N + g.BI - 1) / g.BI;
  const int ktiles = tiles_w * tiles_h * tiles_n;
  const int kt0 = blockIdx.z * g.tiles_per_split;
  const int kt1 = min(ktiles, kt0 + g.tiles_per_split);
  if (kt0 >= kt1 || nb <= 0) return;
  const int nkt = kt1 - kt0;

  constexpr int S = 2;
  const uint32_t a_bytes = 2u * kBoxesA * kBoxBytes;            // hi + lo
  const uint32_t b_bytes = 2u * (uint32_t)g.nb_max * kBoxBytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t bars = smem_base + S * stage_bytes;
  const uint32_t full_bar = bars, empty_bar = bars + 8u * S, tmem_full_bar = bars + 16u * S;
  const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar + 8u * s, 1);
      mbar_init(empty_bar + 8u * s, 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t total_cols = (uint32_t)(kNacc * g.tmem_cols);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(total_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ===== TMA producer =================================================================================================
    // Shared-memory stage: A region = kBoxesA x [hi box | lo box], B region = nb_max x [hi box | lo box].  One TMA
    // instruction fetches the hi and the lo slice of a box (outermost tensor-map dimension = the two slabs of the
    // staged operand): the TMA unit's cost is per instruction (~115-135 ns, scripts/probes/tma_probe.cu), and fourteen
    // single-slice loads per pixel tile paced this loop.
    if (elect_one()) {
      int bcb[kBoxesA], br[kBoxesA], bs[kBoxesA];
#pragma unroll
      for (int j = 0; j < kBoxesA; ++j) {
        int idx = unit * kBoxesA + j;
        if (idx >= TT) idx = TT - 1;                   // duplicate of the last box; masked in the epilogue
        const int bt = idx / cblocks;
        bcb[j] = idx - bt * cblocks;
        br[j] = bt / g.KW;
        bs[j] = bt - br[j] * g.KW;
      }
      const uint32_t tx_bytes = a_bytes + 2u * (uint32_t)nb * kBoxBytes;
      int tw = kt0 % tiles_w, th = (kt0 / tiles_w) % tiles_h, tn = kt0 / (tiles_w * tiles_h);
      int s = 0;
      uint32_t sphase = 0;
      for (int it = 0; it < nkt; ++it) {
        mbar_wait(empty_bar + 8u * s, sphase ^ 1u);
        const int w0 = tw * g.BW, h0 = th * g.BH, n0 = tn * g.BI;
        const uint32_t st_base = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t fb = full_bar + 8u * s;
        mbar_expect_tx(fb, tx_bytes);
#pragma unroll
        for (int j = 0; j < kBoxesA; ++j) {
          const int cw = w0 * g.stride - g.pad + bs[j], ch = h0 * g.stride - g.pad + br[j];
          tma_load_5d(st_base + (uint32_t)(2 * j) * kBoxBytes, &map_b, fb, bcb[j] * 64, cw, ch, n0, 0);
        }
        for (int j = 0; j < nb; ++j)
          tma_load_5d(st_base + a_bytes + (uint32_t)(2 * j) * kBoxBytes, &map_s, fb, (nb0 + j) * 64, w0, h0, n0, 0);
        if (++s == S) {
          s = 0;
          sphase ^= 1u;
        }
        if (++tw == tiles_w) {
          tw = 0;
          if (++th == tiles_h) {
            th = 0;
            ++tn;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ===================================================================================================
    if (elect_one()) {
      // D=f32, A=B=bf16, A and B MN-major (bits 15, 16), N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
      // Stage-0 descriptors; everything else is a compile-time offset or one add per stage (descriptor address fields
      // count 16-byte units), so the 24 MMAs of a pixel tile are straight-line code.  Consecutive 64-wide MN blocks of
      // one slice are 2 boxes apart (LBO), the lo slice of a box follows its hi slice.
      const uint64_t a_d0 = make_mnmajor_sw128_desc(smem_base, 2u * kBoxBytes);
      const uint64_t b_d0 = make_mnmajor_sw128_desc(smem_base + a_bytes, 2u * kBoxBytes);
      constexpr uint64_t kLo = kBoxBytes >> 4, kAcc = (4u * kBoxBytes) >> 4, kStep = 2048u >> 4;
      const uint32_t d0 = tmem_base, d1 = tmem_base + (uint32_t)g.tmem_cols;
      int s = 0;
      uint32_t sphase = 0;
      for (int it = 0; it < nkt; ++it) {
        mbar_wait(full_bar + 8u * s, sphase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t sd = (uint64_t)((uint32_t)s * (stage_bytes >> 4));
        const uint64_t a0 = a_d0 + sd, b0 = b_d0 + sd;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                   // 4 x UMMA_K(16 pixels) = 64; 16 K-rows = 2048 B
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {        // hi*hi, hi*lo, lo*hi
#pragma unroll
            for (int acc = 0; acc < kNacc; ++acc) {     // alternate the two accumulators
              const uint64_t ad = a0 + (uint64_t)acc * kAcc + (pass == 2 ? kLo : 0) + (uint64_t)k * kStep;
              const uint64_t bd = b0 + (pass == 1 ? kLo : 0) + (uint64_t)k * kStep;
              umma_bf16(acc == 0 ? d0 : d1, ad, bd, idesc, (it != 0 || k != 0 || pass != 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar + 8u * s);
        if (++s == S) {
          s = 0;
          sphase ^= 1u;
        }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ===== epilogue: TMEM -> this split's slab of partial sums, part[split][tap][cs][cb] ===================================
    // Plain stores, lanes along cb (128 B per warp instruction); wgrad_reduce_kernel sums the slabs into dW[cs][cb][tap].
    // (The first version added every split's tile into dW with one fp32 atomic per element: 14 M scattered L2 atomics
    // for the 192x192 5x5 layer, ~70 us of its 175 us, and a summation order that changed from run to run.)
    const int q = warp & 3;
    const int row = q * 32 + lane;                      // accumulator row = box (row / 64), channel (row % 64)
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* slab = part + (size_t)blockIdx.z * T * g.Cs * g.Cb;
    for (int acc = 0; acc < kNacc; ++acc) {
      const int idx = unit * kBoxesA + 2 * acc + (row >> 6);
      const int t = idx / cblocks, cb = (idx - t * cblocks) * 64 + (row & 63);
      const bool valid = idx < TT && cb < g.Cb;
      float* dst = slab + ((size_t)t * g.Cs) * g.Cb + cb;
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * g.tmem_cols + c0), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int cs = nb0 * 64 + c0 + j;
            if (cs < g.Cs) dst[(size_t)cs * g.Cb] = __uint_as_float(v[j]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(total_cols) : "memory");
  }
}

// dW[cs][cb][tap] = sum over splits of part[split][tap][cs][cb]; fixed summation order (deterministic): slab z goes
// to accumulator z % 4 (a trailing group of fewer than four slabs to accumulator 0), result (a0 + a1) + (a2 + a3).
// The slabs are L2-resident and the kernel is bound by load latency, so every thread takes four consecutive elements
// (one 16-byte load per slab) and issues the loads of eight slabs before the first add.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int splits, int T, int Cs, int Cb,
                                                            float* __restrict__ dw) {
  pdl_wait();                 // programmatic dependent launch (common.cuh): predecessor complete, memory visible
  const size_t per = (size_t)T * Cs * Cb;
  auto store = [&](size_t i, float acc) {
    const int cb = (int)(i % Cb);
    const size_t r = i / Cb;
    const int cs = (int)(r % Cs), t = (int)(r / Cs);
    dw[((size_t)cs * Cb + cb) * T + t] = acc;
  };
  if (splits > 32 && (per & 3) == 0 && (((uintptr_t)part) & 15) == 0) {
    // Many slabs of a small weight (GDN's 192 x 192 gamma: 147 slabs): one WARP per group of four elements, lane l sums
    // the slabs z = l, l + 32, ... in increasing order, the lanes are combined by a fixed butterfly (xor 16, 8, 4, 2, 1).
    // A thread per element would leave a few dozen CTAs walking 147 dependent loads each.
    const int lane = threadIdx.x & 31;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i4 < (per >> 2); i4 += warps) {
      const float* p = part + (i4 << 2);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int z = lane; z < splits; z += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)z * per));
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, o);
        a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
      }
      if (lane == 0) {
        const size_t i = i4 << 2;
        store(i, a.x);
        store(i + 1, a.y);
        store(i + 2, a.z);
        store(i + 3, a.w);
      }
    }
    return;
  }
  if ((per & 3) == 0 && (((uintptr_t)part) & 15) == 0) {
    for (size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < (per >> 2); i4 += (size_t)gridDim.x * blockDim.x) {
      const float* p = part + (i4 << 2);
      float4 a[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
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
      for (; z < splits; ++z) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)z * per));
        a[0].x += v.x; a[0].y += v.y; a[0].z += v.z; a[0].w += v.w;
      }
      const size_t i = i4 << 2;
      store(i, (a[0].x + a[1].x) + (a[2].x + a[3].x));
      store(i + 1, (a[0].y + a[1].y) + (a[2].y + a[3].y));
      store(i + 2, (a[0].z + a[1].z) + (a[2].z + a[3].z));
      store(i + 3, (a[0].w + a[1].w) + (a[2].w + a[3].w));
    }
    return;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int z = 0;
    for (; z + 3 < splits; z += 4) {
      a0 += __ldg(part + (size_t)z * per + i);
      a1 += __ldg(part + (size_t)(z + 1) * per + i);
      a2 += __ldg(part + (size_t)(z + 2) * per + i);
      a3 += __ldg(part + (size_t)(z + 3) * per + i);
    }
    for (; z < splits; ++z) a0 += __ldg(part + (size_t)z * per + i);
    store(i, (a0 + a1) + (a2 + a3));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
struct WgPlan {
  bool ok = false;
  int CsPad, CbPad, BW, BH, BI, nb_max, n_tiles, units, ktiles, splits, tiles_per_split, tmem_cols;
  size_t small_bytes, big_bytes, part_bytes, total_bytes, smem_bytes;
};

static int p2ceil(int v) {
  int p = 1;
  while (p < v) p *= 2;
  return p;
}

static WgPlan make_wg_plan(int N, int Cs, int Hs, int Ws, int Cb, int Hb, int Wb, int KH, int KW, int stride) {
  WgPlan p;
  if (Cs < 1 || Cb < 1 || KH * KW > 64) return p;
  p.CsPad = (Cs + 63) / 64 * 64;
  p.CbPad = (Cb + 63) / 64 * 64;
  p.BW = Ws >= 16 ? 16 : p2ceil(Ws);
  p.BH = 64 / p.BW;
  if (p.BH > p2ceil(Hs)) p.BH = p2ceil(Hs);
  p.BI = 64 / (p.BW * p.BH);
  if (p.BW * stride > 256 || p.BH * stride > 256 || p.BI > 256) return p;
  const int sblocks = p.CsPad / 64;
  p.nb_max = sblocks >= 3 ? 3 : sblocks;          // N <= 192 per MMA
  if (sblocks == 4) p.nb_max = 2;
  p.n_tiles = (sblocks + p.nb_max - 1) / p.nb_max;
  p.tmem_cols = p.nb_max * 64 <= 64 ? 64 : (p.nb_max * 64 <= 128 ? 128 : 256);
  const int TT = KH * KW * (p.CbPad / 64);
  p.units = (TT + kBoxesA - 1) / kBoxesA;
  p.ktiles = ((Ws + p.BW - 1) / p.BW) * ((Hs + p.BH - 1) / p.BH) * ((N + p.BI - 1) / p.BI);
  const int ctas = p.units * p.n_tiles;
  // split-K so that the grid fills the SMs: ONE round of CTAs when that uses >= 75 % of them, else two rounds without
  // spilling into a third, nearly empty one.  Every split pays a prologue (tensor-memory allocation, barriers) and an
  // epilogue (two 128 x 192 accumulators to its slab) and adds a slab for the reduction to read, so one round of longer
  // K ranges beats two rounds well below a full machine: threshold 0.95 -> 0.75 (19 (tap, block) units x 7 splits = 133
  // CTAs instead of 285) took the sequential sweep 3.36 -> 3.32 ms and the overlapped one 2.76 -> 2.67 ms (A/B on one box,
  // profiles/r2_ab_wgrad_rounds_*.json; 0.65 and "always one round" measure the same).
  const int sms_ = num_sms();
  int want = (2 * sms_) / ctas;
  static double one_round_frac = -1.0;
  if (one_round_frac < 0.0) {
    const char* e = getenv("B200LIC_WG_ONE_ROUND");           // experiments: SM fill from which one round is preferred
    one_round_frac = e ? atof(e) : 0.75;
  }
  if (sms_ / ctas >= 1 && (double)((sms_ / ctas) * ctas) >= one_round_frac * sms_) want = sms_ / ctas;
  int max_splits = (p.ktiles + 3) / 4;            // >= 4 pixel tiles per split
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  p.tiles_per_split = (p.ktiles + want - 1) / want;
  p.splits = (p.ktiles + p.tiles_per_split - 1) / p.tiles_per_split;
  const size_t stage = 2 * (size_t)kBoxesA * kBoxBytes + 2 * (size_t)p.nb_max * kBoxBytes;
  p.smem_bytes = 2 * stage + 1024 + 256;
  if (p.smem_bytes > 227 * 1024) return p;
  p.small_bytes = ((size_t)N * Hs * Ws * p.CsPad * 2 + 1023) / 1024 * 1024;
  p.big_bytes = ((size_t)N * Hb * Wb * p.CbPad * 2 + 1023) / 1024 * 1024;
  p.part_bytes = ((size_t)p.splits * KH * KW * Cs * Cb * sizeof(float) + 1023) / 1024 * 1024;   // per-split partial dW
  p.total_bytes = 2 * p.small_bytes + 2 * p.big_bytes + p.part_bytes + 1024;
  p.ok = true;
  return p;
}

size_t tc_wgrad_workspace_bytes(int N, int Cs, int Hs, int Ws, int Cb, int Hb, int Wb, int KH, int KW, int stride) {
  WgPlan p = make_wg_plan(N, Cs, Hs, Ws, Cb, Hb, Wb, KH, KW, stride);
  return p.ok ? p.total_bytes : 0;
}

// Fused tail (weight_quant.cu): instead of writing dW, the per-split slabs are summed (same fixed order as
// wgrad_reduce_kernel) straight into the AdaRound backward + Adam step of the layer's alpha.
size_t smallc_conv_wgrad_ws(const b200lic_conv_desc* d);
size_t smallc_deconv_wgrad_ws(const b200lic_conv_desc* d);
int launch_wgrad_reduce_adam(const float* part, int splits, int T, int Cs, int Cb, const WgTail* tail, cudaStream_t s);

// small [N,Cs,Hs,Ws], big [N,Cb,Hb,Wb] (fp32 NCHW) -> dw [Cs][Cb][KH][KW] (overwritten)
// Pre-staged operands: `small` / `big` may be nullptr when the matching split-bf16 NHWC operand already exists, either in
// this workspace's own slot (conv_tc_smallc.cu) or, with small_pre / big_pre, in a forward workspace (staged view).
int tc_wgrad_ex(int N, int Cs, int Hs, int Ws, int Cb, int Hb, int Wb, int KH, int KW, int stride, int pad, int big_square,
                const float* small, const float* big, const void* const* small_pre, const void* const* big_pre, float* dw,
                void* workspace, size_t workspace_bytes, cudaStream_t s, const char* name, const WgTail* tail = nullptr,
                int pre_pitch = 0);
int tc_wgrad(int N, int Cs, int Hs, int Ws, int Cb, int Hb, int Wb, int KH, int KW, int stride, int pad, int big_square,
             const float* small, const float* big, float* dw, void* workspace, size_t workspace_bytes, cudaStream_t s,
             const char* name) {
  return tc_wgrad_ex(N, Cs, Hs, Ws, Cb, Hb, Wb, KH, KW, stride, pad, big_square, small, big, nullptr, nullptr, dw,
                     workspace, workspace_bytes, s, name);
}
int tc_wgrad_ex(int N, int Cs, int Hs, int Ws, int Cb, int Hb, int Wb, int KH, int KW, int stride, int pad, int big_square,
                const float* small, const float* big, const void* const* small_pre, const void* const* big_pre, float* dw,
                void* workspace, size_t workspace_bytes, cudaStream_t s, const char* name, const WgTail* tail,
                int pre_pitch) {
  // pre_pitch: channel pitch of a PRE-staged operand when it is not this engine's own padding (a forward workspace pads
  // channels to 32, this engine to 64): the tensor map then describes the real extent and TMA zero-fills the rest of the
  // last 64-channel box
  WgPlan p = make_wg_plan(N, Cs, Hs, Ws, Cb, Hb, Wb, KH, KW, stride);
  if (!p.ok) {
    set_error("%s: shape not eligible for the tcgen05 engine", name);
    return B200LIC_ERR_UNSUPPORTED;
  }
  if (!workspace || workspace_bytes < p.total_bytes) {
    set_error("%s: tcgen05 engine needs %zu workspace bytes (got %zu)", name, p.total_bytes, workspace_bytes);
    return B200LIC_ERR_UNSUPPORTED;
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  void *sh = ws, *sl = ws + p.small_bytes, *bh = ws + 2 * p.small_bytes, *bl = ws + 2 * p.small_bytes + p.big_bytes;
  int rc = B200LIC_OK;
  if (small_pre) {
    sh = const_cast<void*>(small_pre[0]);
    sl = const_cast<void*>(small_pre[1]);
  } else if (small != nullptr) {
    rc = tc_stage_nhwc(small, N, Cs, Hs * Ws, p.CsPad, 0, sh, sl, s);
    if (rc != B200LIC_OK) return rc;
  }
  if (big_pre) {
    bh = const_cast<void*>(big_pre[0]);
    bl = const_cast<void*>(big_pre[1]);
  } else if (big != nullptr) {   // nullptr: the caller staged the gathered operand in its workspace slot (conv_tc_smallc.cu)
    rc = tc_stage_nhwc(big, N, Cb, Hb * Wb, p.CbPad, big_square, bh, bl, s);
    if (rc != B200LIC_OK) return rc;
  }
  float* part = reinterpret_cast<float*>(ws + 2 * p.small_bytes + 2 * p.big_bytes);
  // hi and lo slabs of each operand as the outermost dimension of one 5-D map (one TMA instruction per box pair)
  const long long big_slab = (const uint8_t*)bl - (const uint8_t*)bh, small_slab = (const uint8_t*)sl - (const uint8_t*)sh;
  if (big_slab <= 0 || small_slab <= 0 || (big_slab & 15) || (small_slab & 15)) {
    set_error("%s: staged hi/lo slabs must be ascending and 16-byte aligned", name);
    return B200LIC_ERR_UNSUPPORTED;
  }
  CUtensorMap mb, ms;
  {
    const cuuint64_t bp = (big_pre && pre_pitch > 0) ? (cuuint64_t)pre_pitch : (cuuint64_t)p.CbPad;
    const cuuint64_t sp = (small_pre && pre_pitch > 0) ? (cuuint64_t)pre_pitch : (cuuint64_t)p.CsPad;
    cuuint64_t dims[5] = {bp, (cuuint64_t)Wb, (cuuint64_t)Hb, (cuuint64_t)N, 2};
    cuuint64_t str[4] = {bp * 2, (cuuint64_t)Wb * bp * 2, (cuuint64_t)Hb * Wb * bp * 2, (cuuint64_t)big_slab};
    cuuint32_t box[5] = {64, (cuuint32_t)(p.BW * stride), (cuuint32_t)(p.BH * stride), (cuuint32_t)p.BI, 2};
    cuuint32_t es[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1, 1};
    if (!tc_encode_map(&mb, bh, 5, dims, str, box, es)) return B200LIC_ERR_CUDA;
    cuuint64_t sdims[5] = {sp, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)N, 2};
    cuuint64_t sstr[4] = {sp * 2, (cuuint64_t)Ws * sp * 2, (cuuint64_t)Hs * Ws * sp * 2, (cuuint64_t)small_slab};
    cuuint32_t sbox[5] = {64, (cuuint32_t)p.BW, (cuuint32_t)p.BH, (cuuint32_t)p.BI, 2};
    cuuint32_t ses[5] = {1, 1, 1, 1, 1};
    if (!tc_encode_map(&ms, sh, 5, sdims, sstr, sbox, ses)) return B200LIC_ERR_CUDA;
  }
  WgGeom g{N, Cs, Hs, Ws, p.CsPad, Cb, Hb, Wb, p.CbPad, KH, KW, stride, pad, p.BW, p.BH, p.BI, p.nb_max,
           p.tiles_per_split, p.tmem_cols};
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e2 = cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e2 != cudaSuccess) {
      set_error("%s: cannot raise dynamic shared memory: %s", name, cudaGetErrorString(e2));
      return B200LIC_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid(p.units, p.n_tiles, p.splits);
  launch_pdl(tc_wgrad_kernel, grid, dim3(kWgThreads), p.smem_bytes, s, mb, ms, g, part);
  B200_LAUNCH_CHECK(name);
  if (tail != nullptr) return launch_wgrad_reduce_adam(part, p.splits, KH * KW, Cs, Cb, tail, s);   // (any split count:
  // the fused tail always sums in the serial order; with > 32 slabs it differs from wgrad_reduce_kernel's butterfly in the
  // last bits only)
  const size_t per = (size_t)KH * KW * Cs * Cb;
  // one thread (splits <= 32) or one warp (more slabs) per four elements
  const size_t red_threads = ((per + 3) / 4) * (p.splits > 32 ? 32 : 1);
  launch_pdl(wgrad_reduce_kernel, dim3(grid_for(red_threads, 256)), dim3(256), 0, s, part, p.splits, KH * KW, Cs, Cb, dw);
  B200_LAUNCH_CHECK("wgrad_reduce_kernel");
  return B200LIC_OK;
}

// Slot of the weight-gradient workspace that holds the split-bf16 NHWC copy of dy ([N,Ho,Wo,cpad], channels padded to
// 64): the `small` operand of a conv wgrad, the `big` (gathered) operand of a transposed-conv wgrad.
bool tc_wgrad_dy_slot(const b200lic_conv_desc* d, int transposed, void* workspace, size_t workspace_bytes, void** hi,
                      void** lo, int* cpad) {
  if ((transposed ? smallc_deconv_wgrad_ws(d) : smallc_conv_wgrad_ws(d)) != 0) return false;
  WgPlan p = transposed ? make_wg_plan(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride)
                        : make_wg_plan(d->N, d->Cout, d->Ho, d->Wo, d->Cin, d->H, d->W, d->KH, d->KW, d->stride);
  if (!p.ok || !workspace || workspace_bytes < p.total_bytes) return false;
  uint8_t* ws = reinterpret_cast<uint8_t*>(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  if (transposed) {
    *hi = ws + 2 * p.small_bytes;
    *lo = ws + 2 * p.small_bytes + p.big_bytes;
    *cpad = p.CbPad;
  } else {
    *hi = ws;
    *lo = ws + p.small_bytes;
    *cpad = p.CsPad;
  }
  return true;
}

// x staged by the forward call, dy either fp32 (staged here) or nullptr (already staged in the dy slot), and the
// AdaRound backward + Adam fused behind the split-K reduction when `tail` is given.
int tc_wgrad_prepared(const b200lic_conv_desc* d, int transposed, const void* x_hi, const void* x_lo, int x_pitch,
                      const float* dy, float* dw, void* ws, size_t ws_bytes, const WgTail* tail, cudaStream_t s) {
  if (x_pitch < d->Cin || (x_pitch % 32) != 0) {
    set_error("conv_wgrad (prepared operands): staged x has channel pitch %d for %d channels", x_pitch, d->Cin);
    return B200LIC_ERR_ARG;
  }
  if ((transposed ? smallc_deconv_wgrad_ws(d) : smallc_conv_wgrad_ws(d)) != 0) {
    set_error("conv_wgrad (prepared operands): shape runs on the folded-tap path");
    return B200LIC_ERR_UNSUPPORTED;
  }
  const void* pre[2] = {x_hi, x_lo};
  if (transposed)
    return tc_wgrad_ex(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 0, nullptr, dy,
                       pre, nullptr, dw, ws, ws_bytes, s, "deconv_wgrad(tc, prepared)", tail, x_pitch);
  return tc_wgrad_ex(d->N, d->Cout, d->Ho, d->Wo, d->Cin, d->H, d->W, d->KH, d->KW, d->stride, d->pad, d->in_square, dy,
                     nullptr, nullptr, pre, dw, ws, ws_bytes, s, "conv_wgrad(tc, prepared)", tail, x_pitch);
}

size_t smallc_conv_wgrad_ws(const b200lic_conv_desc* d);
size_t smallc_deconv_wgrad_ws(const b200lic_conv_desc* d);
int smallc_conv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                      cudaStream_t s);
int smallc_deconv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                        cudaStream_t s);

int tc_conv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                  cudaStream_t s) {
  if (smallc_conv_wgrad_ws(d) != 0) return smallc_conv_wgrad(d, x, dy, dw, ws, ws_bytes, s);
  return tc_wgrad(d->N, d->Cout, d->Ho, d->Wo, d->Cin, d->H, d->W, d->KH, d->KW, d->stride, d->pad, d->in_square, dy, x,
                  dw, ws, ws_bytes, s, "conv_wgrad(tc)");
}

int tc_deconv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                    cudaStream_t s) {
  if (smallc_deconv_wgrad_ws(d) != 0) return smallc_deconv_wgrad(d, x, dy, dw, ws, ws_bytes, s);
  return tc_wgrad(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 0, x, dy, dw, ws,
                  ws_bytes, s, "deconv_wgrad(tc)");
}

// x already staged by the forward call (b200lic_conv_staged_view): skip its NHWC split
int tc_conv_wgrad_pre(const b200lic_conv_desc* d, const void* x_hi, const void* x_lo, const float* dy, float* dw, void* ws,
                      size_t ws_bytes, cudaStream_t s) {
  const void* pre[2] = {x_hi, x_lo};
  return tc_wgrad_ex(d->N, d->Cout, d->Ho, d->Wo, d->Cin, d->H, d->W, d->KH, d->KW, d->stride, d->pad, d->in_square, dy,
                     nullptr, nullptr, pre, dw, ws, ws_bytes, s, "conv_wgrad(tc, staged x)");
}
int tc_deconv_wgrad_pre(const b200lic_conv_desc* d, const void* x_hi, const void* x_lo, const float* dy, float* dw,
                        void* ws, size_t ws_bytes, cudaStream_t s) {
  const void* pre[2] = {x_hi, x_lo};
  return tc_wgrad_ex(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 0, nullptr, dy, pre,
                     nullptr, dw, ws, ws_bytes, s, "deconv_wgrad(tc, staged x)");
}

size_t tc_conv_wgrad_ws(const b200lic_conv_desc* d) {
  if (const size_t n = smallc_conv_wgrad_ws(d)) return n;
  return tc_wgrad_workspace_bytes(d->N, d->Cout, d->Ho, d->Wo, d->Cin, d->H, d->W, d->KH, d->KW, d->stride);
}
size_t tc_deconv_wgrad_ws(const b200lic_conv_desc* d) {
  if (const size_t n = smallc_deconv_wgrad_ws(d)) return n;
  return tc_wgrad_workspace_bytes(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride);
}

}  // namespace b200lic
