// Host side and operand staging of the tcgen05 tensor-core engine (B200LIC_ENGINE_TC) for conv / transposed-conv forward
// and dgrad; the GEMM kernels are conv_tc2.cu (generic implicit GEMM) and gemm1x1_tc.cu (short-K 1x1 layers).
//
// Formulation (same gather geometry as conv_simt.cu, so both engines are interchangeable and cross-checked):
//   D[pixel, co] = sum_{tap, ci} A[pixel, (tap, ci)] * B[co, (tap, ci)]        M = 128 output pixels, N = BN channels
// * operands are staged into tensor-core friendly form by two small HBM-bound kernels -- unless the caller already holds
//   them in that form (prepared operands, prepared.cu):
//     activations  NCHW fp32 -> NHWC bf16 "hi" and "lo" slices (x = hi + lo + O(2^-17 |x|)), channels padded to 32
//     weights      -> [phase][Cout][tap][ci] bf16 hi/lo slices, K-major
//   three MMA passes per k-block (hi*hi + hi*lo + lo*hi) accumulate in fp32 in TMEM: ~2^-16 relative product error,
//   inside the 1e-4 per-layer bar while running on the bf16 tensor pipe (plain bf16/TF32 would not meet it).
// This file: the staging kernels, the tensor-map encoder, and the dispatch between the folded-tap path
// (conv_tc_smallc.cu) and the generic one.  Replaces cuDNN under F.conv2d / F.conv_transpose2d (TO quant_layer.py:28,36,123)
// and their dgrad.
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

// the GEMM engine lives in conv_tc2.cu (the first-generation kernel that used to sit in this file is gone: round 1's A/B
// runs are recorded in profiles/README.md)
size_t tc2_workspace_bytes(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride,
                           int transposed);
int tc2_launch(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride, int pad,
               int transposed, long long s_co, long long s_ci, int act, float slope, int in_square, int gdn_mode,
               int fixed_point, const float* x, const float* w, const float* bias, const float* gdn_x, float* norm_out,
               float* y, void* workspace, size_t workspace_bytes, cudaStream_t s, const char* name);

static int tc_launch(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride, int pad,
                     int transposed, long long s_co, long long s_ci, int act, float slope, int in_square, int gdn_mode,
                     int fixed_point, const float* x, const float* w, const float* bias, const float* gdn_x,
                     float* norm_out, float* y, void* workspace, size_t workspace_bytes, cudaStream_t s,
                     const char* name) {
  return tc2_launch(N, Cin, H, W, Cout, Ho, Wo, KH, KW, stride, pad, transposed, s_co, s_ci, act, slope, in_square,
                    gdn_mode, fixed_point, x, w, bias, gdn_x, norm_out, y, workspace, workspace_bytes, s, name);
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
  return tc2_workspace_bytes(N, Cin, H, W, Cout, Ho, Wo, KH, KW, stride, transposed);
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
  if (!fwd_ws || (d->Cin % 64) != 0) return false;
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
  if (const size_t n = smallc_conv_fwd_ws(d)) return n;
  return tc_workspace_bytes(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, 0);
}
size_t tc_deconv_fwd_ws(const b200lic_conv_desc* d) {
  if (const size_t n = smallc_deconv_fwd_ws(d)) return n;
  return tc_workspace_bytes(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, 1);
}

int tc_conv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, const float* gdn_x,
                float* norm_out, float* y, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (smallc_conv_fwd_ws(d) != 0) return smallc_conv_fwd(d, x, w, bias, y, ws, ws_bytes, s);
  return tc_launch(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 0,
                   (long long)d->Cin * d->KH * d->KW, (long long)d->KH * d->KW, d->act, d->act_slope, d->in_square,
                   d->gdn_mode, d->fixed_point, x, w, bias, gdn_x, norm_out, y, ws, ws_bytes, s, "conv_fwd(tc)");
}

int tc_deconv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void* ws,
                  size_t ws_bytes, cudaStream_t s) {
  if (smallc_deconv_fwd_ws(d) != 0) return smallc_deconv_fwd(d, x, w, bias, y, ws, ws_bytes, s);
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
  if (smallc_conv_fwd_ws(d) != 0) {
    set_error("conv_fwd_wq: shape runs on the folded-tap path");
    return B200LIC_ERR_UNSUPPORTED;
  }
  return tc2_launch_wq(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 0,
                       (long long)d->Cin * d->KH * d->KW, (long long)d->KH * d->KW, d->act, d->act_slope, 0, 0,
                       d->fixed_point, x, w_int, w_scale, bias, nullptr, nullptr, y, ws, ws_bytes, s, "conv_fwd_wq(tc)");
}
int tc_deconv_fwd_wq(const b200lic_conv_desc* d, const float* x, const float* w_int, const float* w_scale,
                     const float* bias, float* y, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (smallc_deconv_fwd_ws(d) != 0) {
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
