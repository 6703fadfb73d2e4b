// 1x1 convolutions with a short contraction (K = padded input channels <= 256) on tcgen05: the weight operand stays in
// shared memory for the whole kernel and the epilogue stores straight from registers.
//
// The generic engine (conv_tc2.cu) hides its epilogue behind the K loop of the next work item; with a handful of K blocks
// per item there is nothing to hide behind and its per-chunk staging round (TMEM -> shared -> TMA store, two named
// barriers and a proxy fence per 32 channels) paces the layer: the folded 3 -> N analysis conv ran 6 144 items x 10 us at
// 10 % tensor-pipe activity (profiles/README.md r2, item 24).  The layers concerned are exactly the 1x1 problems of the
// codec: the folded-tap ends (conv_tc_smallc.cu: K = 75 -> 96, or N = 75 -> 80), and the norm GEMM of a GDN unit during
// calibration (K = N = 192).  Here
//   warp 0     TMA: the packed weight operand [CoutPad][Cpad] hi | lo once per CTA (tc2_weight_layout form), then the
//              staged activation operand (split-bf16 NHWC, [128 px][32 ch] hi | lo boxes) through a ring
//   warp 1     tcgen05.mma issuer: D[128 px, CoutPad] += A . B^T, 6 MMAs per K block (4 with the bf16-exact integer
//              weight operand), two TMEM accumulators
//   warps 2-9  epilogue: TMEM -> (* w_scale) + bias -> activation -> fp32 NCHW, one pixel per thread and 16 channels per
//              TMEM load, lanes along the pixel axis: every store instruction writes 128 contiguous bytes
// Same arithmetic and accumulation order as the generic engine's single-chain form (K blocks in order; hi*hi, hi*lo, lo*hi).
// Replaces F.conv2d / F.conv_transpose2d of those layers (TO quant_layer.py:28,36,123) like conv_tc2.cu does.
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace b200lic {

bool tc_encode_map_ex(CUtensorMap* m, CUtensorMapDataType dt, CUtensorMapSwizzle sw, void* base, int rank,
                      const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box, const cuuint32_t* estr);

namespace g1 {

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
// K-major SWIZZLE_64B operand tile: rows of 64 B (32 bf16), 8-row groups of 512 B (the staged form of conv_tc2.cu)
__device__ __forceinline__ uint64_t make_kmajor_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

}  // namespace g1

constexpr int kG1Threads = 320;                 // TMA, MMA, 8 epilogue warps
constexpr uint32_t kG1AStage = 2 * 128 * 64;    // [128 px][32 ch] bf16, hi | lo
constexpr int kG1MaxSA = 8;

struct Gemm1x1Geom {
  long long M;          // pixels in total (N * H * W)
  int HW, Cout;
  int nkb;              // K blocks of 32 channels
  int BN;               // accumulator columns = CoutPad
  int n_tiles, SA, acc_cols;
  int w_exact;          // integer weight operand: hi slab only, two passes, scale in the epilogue
  int act;
  float slope;
};

__global__ void __launch_bounds__(kG1Threads, 1)
    gemm1x1_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Gemm1x1Geom g,
                   const float* __restrict__ bias, const float* __restrict__ w_scale, float* __restrict__ y,
                   unsigned* __restrict__ stats) {
  pdl_wait();                 // programmatic dependent launch (common.cuh): predecessor complete, memory visible
  using namespace g1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t b_slab = (uint32_t)g.BN * 64u;                       // one K block of one slab
  const uint32_t b_kb = (g.w_exact ? 1u : 2u) * b_slab;               // hi | lo of one K block
  const uint32_t b_area = smem_base;
  const uint32_t a_ring = b_area + (uint32_t)g.nkb * b_kb;
  const uint32_t par = a_ring + (uint32_t)g.SA * kG1AStage;
  float* s_bias = reinterpret_cast<float*>(smem_gen + (par - smem_base));
  float* s_scale = s_bias + 256;
  // [256] min keys, [256] max keys of the output channels (stats != nullptr): per-channel statistics for the next
  // layer's dynamic activation quantiser, taken from the registers of the epilogue (see conv_tc2.cu)
  unsigned* s_stat = reinterpret_cast<unsigned*>(s_scale + 256);
  const uint32_t bars = par + 4096u;
  const uint32_t a_full = bars, a_empty = a_full + 8u * kG1MaxSA, b_full = a_empty + 8u * kG1MaxSA;
  const uint32_t t_full = b_full + 8u, t_empty = t_full + 16u;
  const uint32_t tmem_ptr_addr = t_empty + 16u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_addr - smem_base));

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.SA; ++s) {
      mbar_init(a_full + 8u * s, 1);
      mbar_init(a_empty + 8u * s, 1);
    }
    mbar_init(b_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full + 8u * s, 1);
      mbar_init(t_empty + 8u * s, 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = threadIdx.x; c < 256; c += kG1Threads) {
    s_bias[c] = (bias && c < g.Cout) ? __ldg(bias + c) : 0.f;
    s_scale[c] = (w_scale && c < g.Cout) ? __ldg(w_scale + c) : 1.f;
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
    // ===== TMA: weights once, then the activation ring ====================================================================
    if (elect_one()) {
      mbar_expect_tx(b_full, (uint32_t)g.nkb * b_kb);
      for (int kb = 0; kb < g.nkb; ++kb) tma_load_3d(b_area + (uint32_t)kb * b_kb, &map_b, b_full, kb * 32, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int t = first; t < g.n_tiles; t += step) {
        for (int kb = 0; kb < g.nkb; ++kb) {
          mbar_wait(a_empty + 8u * s, ph ^ 1u);
          mbar_expect_tx(a_full + 8u * s, kG1AStage);
          tma_load_3d(a_ring + (uint32_t)s * kG1AStage, &map_a, a_full + 8u * s, kb * 32, t * 128, 0);
          if (++s == g.SA) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====================================================================================================
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.BN >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t a_d0 = make_kmajor_sw64_desc(a_ring);
      const uint64_t b_d0 = make_kmajor_sw64_desc(b_area);
      const uint64_t a_lo = (uint64_t)((kG1AStage / 2) >> 4), b_lo = (uint64_t)(b_slab >> 4);
      mbar_wait(b_full, 0);
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph[2] = {0u, 0u};
      for (int t = first; t < g.n_tiles; t += step) {
        mbar_wait(t_empty + 8u * acc, acc_ph[acc] ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + (uint32_t)(acc * g.acc_cols);
        for (int kb = 0; kb < g.nkb; ++kb) {
          mbar_wait(a_full + 8u * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t a0 = a_d0 + (uint64_t)((uint32_t)s * (kG1AStage >> 4));
          const uint64_t b0 = b_d0 + (uint64_t)((uint32_t)kb * (b_kb >> 4));
#pragma unroll
          for (int k = 0; k < 2; ++k) {          // UMMA_K = 16 channels = 32 B inside the 64 B swizzle row
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {   // hi*hi, hi*lo, lo*hi
              if (pass == 1 && g.w_exact) continue;  // exact weights have no lo slice
              const uint64_t ad = a0 + (pass == 2 ? a_lo : 0) + (uint64_t)(k * 2);
              const uint64_t bd = b0 + (pass == 1 ? b_lo : 0) + (uint64_t)(k * 2);
              umma_bf16(d, ad, bd, idesc, (kb != 0 || k != 0 || pass != 0) ? 1u : 0u);
            }
          }
          umma_commit(a_empty + 8u * s);
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
  } else {
    // ===== epilogue =======================================================================================================
    const int q = warp & 3;                      // TMEM lane quarter of this warp
    const int row = q * 32 + lane;
    const int eh = (warp - 2) >> 2;              // the two warps of a quarter take alternate 16-channel chunks
    int acc = 0;
    uint32_t acc_ph[2] = {0u, 0u};
    const int act = g.act;
    const float slope = g.slope;
    for (int t = first; t < g.n_tiles; t += step) {
      const long long p = (long long)t * 128 + row;
      const bool valid = p < g.M;
      const long long n = valid ? p / g.HW : 0;
      const long long hw = valid ? p - n * g.HW : 0;
      float* yp = y + (size_t)(n * g.Cout) * g.HW + hw;
      const bool full = (long long)t * 128 + 128 <= g.M && g.Cout == g.BN;
      mbar_wait(t_full + 8u * acc, acc_ph[acc]);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * g.acc_cols);
      for (int c0 = eh * 16; c0 < g.BN; c0 += 32) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        float bj[16], sj[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          bj[j] = s_bias[c0 + j];
          sj[j] = s_scale[c0 + j];
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float r[16];
        if (g.w_exact) {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = fmaf(__uint_as_float(v[j]), sj[j], bj[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = __uint_as_float(v[j]) + bj[j];
        }
        if (act == B200LIC_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = fmaxf(r[j], 0.f);
        } else if (act == B200LIC_ACT_LEAKY_RELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = r[j] > 0.f ? r[j] : r[j] * slope;
        }
        if (stats != nullptr) {
          unsigned kmn, kmx;
          warp_channel_minmax16(r, valid, lane, kmn, kmx);
          if (lane < 16 && c0 + lane < g.Cout) {
            atomicMin(s_stat + c0 + lane, kmn);
            atomicMax(s_stat + 256 + c0 + lane, kmx);
          }
        }
        float* yc = yp + (size_t)c0 * g.HW;
        if (full) {
#pragma unroll
          for (int j = 0; j < 16; ++j) yc[(size_t)j * g.HW] = r[j];
        } else if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < g.Cout) yc[(size_t)j * g.HW] = r[j];
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(t_empty + 8u * acc);
      acc_ph[acc] ^= 1u;
      acc ^= 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (stats != nullptr) {
    for (int c = threadIdx.x; c < g.Cout; c += kG1Threads) {
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

// 1 = off (the generic engine runs these layers), anything else = on.  B200LIC_TC_GEMM1X1=0 in the environment or
// b200lic_set_option("gemm1x1", 0) for A/B runs and tests.
static int g_gemm1x1_mode = -1;
void gemm1x1_set_mode(int v) { g_gemm1x1_mode = v ? 1 : 0; }
static bool gemm1x1_enabled() {
  if (g_gemm1x1_mode < 0) {
    const char* e = getenv("B200LIC_TC_GEMM1X1");
    g_gemm1x1_mode = (e && atoi(e) == 0) ? 0 : 1;
  }
  return g_gemm1x1_mode != 0;
}

static size_t gemm1x1_smem(int nkb, int BN, int w_exact, int SA) {
  return 1024 + (size_t)nkb * (w_exact ? 1 : 2) * BN * 64 + (size_t)SA * kG1AStage + 4096 + 512;
}

// Eligibility of a planned tc2 launch: 1x1, stride 1, no padding, plain epilogue, the whole weight operand and at least
// three activation stages in shared memory.
bool gemm1x1_eligible(int KH, int KW, int stride, int pad, int Cpad, int CoutPad, int n_out_tiles, int gdn_mode,
                      int fixed_point, int w_exact, int H, int W, int Ho, int Wo) {
  if (!gemm1x1_enabled()) return false;
  if (KH != 1 || KW != 1 || stride != 1 || pad != 0 || gdn_mode || fixed_point || n_out_tiles != 1) return false;
  if (H != Ho || W != Wo || Cpad > 256 || CoutPad > 256 || (CoutPad % 16) != 0) return false;
  return gemm1x1_smem(Cpad / 32, CoutPad, w_exact, 3) <= 227 * 1024;
}

// xh / xl: staged activation operand [M][Cpad] bf16 (lo slab x_bytes behind the hi slab); bh: packed weights
// [CoutPad][Cpad] hi, lo slab b_bytes behind it.
int gemm1x1_launch(long long M, int HW, int Cpad, int Cout, int CoutPad, void* xh, size_t x_bytes, void* bh, size_t b_bytes,
                   int w_exact, const float* w_scale, const float* bias, int act, float slope, float* y, unsigned* stats,
                   cudaStream_t s, const char* name) {
  Gemm1x1Geom g{};
  g.M = M;
  g.HW = HW;
  g.Cout = Cout;
  g.nkb = Cpad / 32;
  g.BN = CoutPad;
  const long long tiles = (M + 127) / 128;
  if (tiles > 2147483647LL / 128) {
    set_error("%s: tensor too large for the 1x1 engine", name);
    return B200LIC_ERR_UNSUPPORTED;
  }
  g.n_tiles = (int)tiles;
  g.w_exact = w_exact ? 1 : 0;
  g.act = act;
  g.slope = slope;
  g.acc_cols = 32;
  while (g.acc_cols < g.BN) g.acc_cols *= 2;
  g.SA = 3;
  while (g.SA < kG1MaxSA && gemm1x1_smem(g.nkb, g.BN, g.w_exact, g.SA + 1) <= 227 * 1024) ++g.SA;
  const size_t smem = gemm1x1_smem(g.nkb, g.BN, g.w_exact, g.SA);

  CUtensorMap ma, mb;
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cpad, (cuuint64_t)M, 2};
    cuuint64_t strides[2] = {(cuuint64_t)Cpad * 2, (cuuint64_t)x_bytes};
    cuuint32_t box[3] = {32, 128, 2};
    cuuint32_t es[3] = {1, 1, 1};
    if (!tc_encode_map_ex(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B, xh, 3, dims, strides, box, es))
      return B200LIC_ERR_CUDA;
    cuuint64_t bdims[3] = {(cuuint64_t)Cpad, (cuuint64_t)CoutPad, 2};
    cuuint64_t bstrides[2] = {(cuuint64_t)Cpad * 2, (cuuint64_t)b_bytes};
    cuuint32_t bbox[3] = {32, (cuuint32_t)CoutPad, (cuuint32_t)(w_exact ? 1 : 2)};
    if (!tc_encode_map_ex(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B, bh, 3, bdims, bstrides, bbox, es))
      return B200LIC_ERR_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm1x1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("%s: cannot raise dynamic shared memory: %s", name, cudaGetErrorString(e));
      return B200LIC_ERR_CUDA;
    }
    attr_set = true;
  }
  const int sms = num_sms();
  const int grid = g.n_tiles < sms ? g.n_tiles : sms;
  launch_pdl(gemm1x1_kernel, dim3(grid), dim3(kG1Threads), smem, s, ma, mb, g, bias, w_scale, y, stats);
  B200_LAUNCH_CHECK(name);
  return B200LIC_OK;
}

}  // namespace b200lic
