// Micro-benchmark: cost of one TMA tile load as a function of the box geometry (rows x row bytes, rank, element
// strides), single CTA and all SMs, source tensor L2-resident.  Evidence for the A-operand staging design of
// conv_tc2.cu (profiles/README.md).   nvcc -gencode arch=compute_100a,code=sm_100a -o build/tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

struct Cfg { int rank; int c[5]; int step_dim; int step; int wrap; uint32_t bytes; int depth; int iters; int prefetch; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap map, Cfg cfg, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bars = base + 6u * 32768u;       // up to 6 slots of 32 KB
  if (threadIdx.x == 0) {
    for (int i = 0; i < cfg.depth; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8u * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (cfg.prefetch) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map)) : "memory");
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    int issued = 0;
    uint32_t ph = 0;
    const uint32_t bar = bars;
    for (int it = 0; it < cfg.iters; ++it) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cfg.bytes * cfg.depth) : "memory");
      for (int s = 0; s < cfg.depth; ++s) {
        const uint32_t dst = base + 32768u * s;
        int c[5] = {cfg.c[0], cfg.c[1], cfg.c[2], cfg.c[3], cfg.c[4]};
        c[cfg.step_dim] += ((issued + (int)blockIdx.x * 7) % cfg.wrap) * cfg.step;
        if (cfg.rank == 2)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar), "r"(c[0]), "r"(c[1]) : "memory");
        else if (cfg.rank == 3)
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
        else if (cfg.rank == 4)
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                       ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
        ++issued;
      }
      while (!mbar_try(bar, ph)) {}
      ph ^= 1u;
    }
    unsigned long long t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    out[blockIdx.x] = t1 - t0;
  }
}

static bool encode(CUtensorMap* m, CUtensorMapSwizzle sw, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                   const cuuint32_t* box, const cuuint32_t* es) {
  CUresult r = cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("  encode failed %d\n", (int)r); return false; }
  return true;
}

int main() {
  cudaFree(0);
  const int N = 8, H = 128, W = 128, C = 192;      // NHWC bf16, 50 MB: L2-resident
  void* x;
  cudaMalloc(&x, (size_t)N * H * W * C * 2);
  cudaMemset(x, 0, (size_t)N * H * W * C * 2);
  unsigned long long* out;
  cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Case { const char* name; int rank; CUtensorMapSwizzle sw; cuuint64_t dims[5]; cuuint64_t str[4]; cuuint32_t box[5]; cuuint32_t es[5]; int step_dim, step, wrap; };
  const cuuint64_t pix = (cuuint64_t)C * 2;
  const cuuint64_t half = (cuuint64_t)N * H * W * C;      // bytes of one half of the buffer (two [N/2..] halves as "hi"/"lo")
  std::vector<Case> cases = {
    {"5D {32ch,16w,8h,1n,2} SW64 hi+lo in one load (256 rows x 64 B)", 5, CU_TENSOR_MAP_SWIZZLE_64B, {C, W, H, N / 2, 2}, {pix, pix * W, pix * W * H, half}, {32, 16, 8, 1, 2}, {1, 1, 1, 1, 1}, 0, 32, 6},
    {"5D {32ch,32w,16h,1n,2} SW64 es2 hi+lo in one load (256 rows x 64 B)", 5, CU_TENSOR_MAP_SWIZZLE_64B, {C, W, H, N / 2, 2}, {pix, pix * W, pix * W * H, half}, {32, 32, 16, 1, 2}, {1, 2, 2, 1, 1}, 0, 32, 6},
    {"3D {32k,192rows,2} SW64 weights hi+lo in one load (384 rows x 64 B)", 3, CU_TENSOR_MAP_SWIZZLE_64B, {4800, 2048, 2}, {9600, 9600 * 2048, 0}, {32, 192, 2}, {1, 1, 1}, 0, 32, 150},
    {"3D {32k,96rows,2} SW64 weights hi+lo (192 rows x 64 B)", 3, CU_TENSOR_MAP_SWIZZLE_64B, {4800, 2048, 2}, {9600, 9600 * 2048, 0}, {32, 96, 2}, {1, 1, 1}, 0, 32, 150},
    {"4D {32ch,16w,8h,1n} SW64 es1   (128 rows x 64 B)", 4, CU_TENSOR_MAP_SWIZZLE_64B, {C, W, H, N}, {pix, pix * W, pix * W * H}, {32, 16, 8, 1}, {1, 1, 1, 1}, 0, 32, 6},
    {"4D {32ch,32w,16h,1n} SW64 es2  (128 rows x 64 B, stride-2 conv)", 4, CU_TENSOR_MAP_SWIZZLE_64B, {C, W, H, N}, {pix, pix * W, pix * W * H}, {32, 32, 16, 1}, {1, 2, 2, 1}, 0, 32, 6},
    {"4D {32ch,128w,1h,1n} SW64 es1  (128 rows x 64 B, one image row)", 4, CU_TENSOR_MAP_SWIZZLE_64B, {C, W, H, N}, {pix, pix * W, pix * W * H}, {32, 128, 1, 1}, {1, 1, 1, 1}, 0, 32, 6},
    {"4D {64ch,16w,8h,1n} SW128 es1  (128 rows x 128 B)", 4, CU_TENSOR_MAP_SWIZZLE_128B, {C, W, H, N}, {pix, pix * W, pix * W * H}, {64, 16, 8, 1}, {1, 1, 1, 1}, 0, 64, 3},
    {"4D {64ch,32w,16h,1n} SW128 es2 (128 rows x 128 B, stride 2)", 4, CU_TENSOR_MAP_SWIZZLE_128B, {C, W, H, N}, {pix, pix * W, pix * W * H}, {64, 32, 16, 1}, {1, 2, 2, 1}, 0, 64, 3},
    {"4D {64ch,16w,4h,1n} SW128 es1  (64 rows x 128 B)", 4, CU_TENSOR_MAP_SWIZZLE_128B, {C, W, H, N}, {pix, pix * W, pix * W * H}, {64, 16, 4, 1}, {1, 1, 1, 1}, 0, 64, 3},
    {"4D {16ch,16w,8h,1n} SW32 es1   (128 rows x 32 B)", 4, CU_TENSOR_MAP_SWIZZLE_32B, {C, W, H, N}, {pix, pix * W, pix * W * H}, {16, 16, 8, 1}, {1, 1, 1, 1}, 0, 16, 12},
    {"2D {32ch,128px} SW64           (128 rows x 64 B)", 2, CU_TENSOR_MAP_SWIZZLE_64B, {C, (cuuint64_t)N * H * W, 1, 1}, {pix, 0, 0}, {32, 128, 1, 1}, {1, 1, 1, 1}, 0, 32, 6},
    {"2D {64ch,128px} SW128          (128 rows x 128 B)", 2, CU_TENSOR_MAP_SWIZZLE_128B, {C, (cuuint64_t)N * H * W, 1, 1}, {pix, 0, 0}, {64, 128, 1, 1}, {1, 1, 1, 1}, 0, 64, 3},
    {"2D {64ch,256px} SW128          (256 rows x 128 B)", 2, CU_TENSOR_MAP_SWIZZLE_128B, {C, (cuuint64_t)N * H * W, 1, 1}, {pix, 0, 0}, {64, 256, 1, 1}, {1, 1, 1, 1}, 0, 64, 3},
    {"2D {32ch,192rows} SW64 weights (192 rows x 64 B, row stride 9600 B)", 2, CU_TENSOR_MAP_SWIZZLE_64B, {4800, 4096, 1, 1}, {9600, 0, 0}, {32, 192, 1, 1}, {1, 1, 1, 1}, 0, 32, 150},
  };
  {
    std::vector<Case> extra;
    for (auto cs : cases) {
      if (cs.dims[0] != (cuuint64_t)C) continue;
      Case v = cs;
      v.step_dim = 1;
      v.step = (int)cs.box[1];
      v.wrap = cs.rank == 2 ? 200 : (int)(W / cs.box[1] > 0 ? W / cs.box[1] : 1);
      v.name = "  ^ same box, stepping along pixels instead of channels";
      extra.push_back(cs);
      extra.push_back(v);
    }
    extra.push_back(cases.back());
    cases = extra;
  }
  for (auto& cs : cases) {
    CUtensorMap m;
    if (!encode(&m, cs.sw, x, cs.rank, cs.dims, cs.str, cs.box, cs.es)) continue;
    uint32_t elems = 1;
    for (int i = 0; i < cs.rank; ++i) elems *= (cs.box[i] + cs.es[i] - 1) / cs.es[i];
    for (int grid : {1}) {
      for (int depth : {1, 2, 4}) {
        Cfg cfg{cs.rank, {0, 0, 0, 0, 0}, cs.step_dim, cs.step, cs.wrap, elems * 2, depth, 2000, depth != 4};
        probe<<<grid, 128, 227 * 1024, 0>>>(m, cfg, out);
        probe<<<grid, 128, 227 * 1024, 0>>>(m, cfg, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        unsigned long long h[148];
        cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        const double ns = (double)mx / cfg.iters;
        printf("%-70s grid %3d batch %d: %7.1f ns/batch  %6.1f GB/s/SM  (%u B per load)\n", cs.name, grid, depth, ns, cfg.bytes * depth / ns, cfg.bytes);
      }
    }
  }
  return 0;
}
