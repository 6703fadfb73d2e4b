// Shared plumbing of libb200lic: error reporting, arch gate, launch accounting, warp/block reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/b200lic.h"

namespace b200lic {

void set_error(const char* fmt, ...);
int check_arch();                       // B200LIC_OK iff current device is sm_100
void count_launch(int n = 1);
int num_sms();

#define B200_REQUIRE(cond, ...)                              \
  do {                                                       \
    if (!(cond)) {                                           \
      ::b200lic::set_error(__VA_ARGS__);                     \
      return B200LIC_ERR_ARG;                                \
    }                                                        \
  } while (0)

#define B200_ARCH_GATE()                                     \
  do {                                                       \
    int _a = ::b200lic::check_arch();                        \
    if (_a != B200LIC_OK) return _a;                         \
  } while (0)

#define B200_LAUNCH_CHECK(name)                                                        \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      ::b200lic::set_error("%s: launch failed: %s", name, cudaGetErrorString(_e));     \
      return B200LIC_ERR_CUDA;                                                         \
    }                                                                                  \
    ::b200lic::count_launch();                                                         \
  } while (0)

static inline cudaStream_t as_stream(b200lic_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Persistent-style 1-D grid: a multiple of the SM count, capped by the work available.
static inline int grid_for(size_t work_items, int per_block, int blocks_per_sm = 8) {
  size_t need = (work_items + per_block - 1) / per_block;
  size_t cap = (size_t)num_sms() * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; result valid in thread 0.  `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? red[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

// Order-preserving float <-> uint key (for atomicMin/atomicMax on floats).
__device__ __forceinline__ unsigned f2key(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Warp-wide fp32 min / max in one instruction (CREDUX.MIN/MAX.F32, sm_100a); NaN inputs are ignored like fminf / fmaxf.
__device__ __forceinline__ float warp_redux_min_f32(float v) {
  float r;
  asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float warp_redux_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
// Per-channel (min, max) of 16 channel values per lane over the lanes of a warp (one pixel per lane): lane j < 16 returns
// the ordered-integer keys (f2key) of channel j.  `valid`: this lane's pixel exists (ragged tiles).
__device__ __forceinline__ void warp_channel_minmax16(const float (&r)[16], bool valid, int lane, unsigned& kmn,
                                                      unsigned& kmx) {
  const bool allv = __all_sync(0xffffffffu, valid);
  float fmn = INFINITY, fmx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float lo = r[j], hi = r[j];
    if (!allv) {
      lo = valid ? lo : INFINITY;
      hi = valid ? hi : -INFINITY;
    }
    const float a = warp_redux_min_f32(lo), b = warp_redux_max_f32(hi);
    if (lane == j) {
      fmn = a;
      fmx = b;
    }
  }
  kmn = f2key(fmn);      // an all-invalid warp yields key(+inf) / key(-inf): neutral against any finite value
  kmx = f2key(fmx);
}

// ---- programmatic dependent launch ----------------------------------------------------------------------------------
// The kernels of the calibration sweep / evaluation forward are launched through launch_pdl(): such a kernel may be
// scheduled as soon as the CTAs of its predecessor in the stream have exited (the implicit trigger), without waiting
// for the predecessor's completion to be processed as a separate event; its first instruction, pdl_wait(), blocks until
// the predecessor's memory is visible.  Measured (bench.py --skip-cpu --skip-fwd --skip-extra, same box, twice each):
// overlapped schedule 54.2 k -> 56.5 k imgs/s, sequential unchanged (3.474 -> 3.469 ms).  An EXPLICIT early trigger
// (griddepcontrol.launch_dependents at kernel entry) was slower -- 3.458 -> 3.536 ms: the successors' CTAs sit on the SMs
// the predecessor's later waves need.  pdl_wait() is a no-op in a normally launched kernel; B200LIC_PDL=0 launches
// everything normally.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();

// cluster_x > 1: the grid is launched as thread-block clusters of that many CTAs (grid.x a multiple of it).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                             int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster_x;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args... args) {
  return launch_pdl_cluster(kernel, grid, block, smem, s, 1, args...);
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == B200LIC_ACT_RELU) return fmaxf(v, 0.f);
  if (act == B200LIC_ACT_LEAKY_RELU) return v > 0.f ? v : v * slope;
  return v;
}

}  // namespace b200lic
