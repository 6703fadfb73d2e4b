// Small-channel ends of the codec (3 -> N analysis conv, N -> 3 synthesis transposed conv) on the tensor-core engine.
//
// With Cin = 3 (or Cout = 3) the generic tap-by-tap implicit GEMM wastes its K (or N) dimension on channel padding:
// 25 taps x 32 padded channels for 75 real products per pixel, or 16-wide N tiles that re-stream the whole input once
// per tap.  Folding the taps into the channel axis turns both layers into 1x1 problems of the same engine:
//   conv    fwd : A'[pixel, (ci,r,s)] = im2col(x) gathered straight into the split-bf16 NHWC operand (K = Cin*k*k),
//                 y = A' . W'^T with W'[co, (ci,r,s)] = w viewed as [Cout, Cin*k*k]                 (tc2 engine, 1x1)
//   conv  wgrad : dW'[co, (ci,r,s)] = sum_pixels dy[pixel, co] * A'[pixel, (ci,r,s)]                (wgrad engine, 1x1)
//   tconv   fwd : col[pixel_in, (co,r,s)] = x . W'' with W''[(co,r,s), ci] = w viewed as [Cin, Cout*k*k]^T (tc2, 1x1),
//                 y = bias + col2im(col) (+ activation / Q8.8)                                      (gather kernel)
//   tconv wgrad : dW[ci, (co,r,s)] = sum_pixels x[pixel, ci] * im2col(dy)[pixel, (co,r,s)]          (wgrad engine, 1x1)
// Replaces the same reference calls as conv_tc2.cu (F.conv2d / F.conv_transpose2d and their autograd for g_a.0 and
// g_s.6 of the Balle / Minnen graphs, TO quant_layer.py:28,36,123).
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace b200lic {

void tc2_stats_once(unsigned* keys);
unsigned* tc2_stats_peek();
int tc2_launch(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride, int pad,
               int transposed, long long s_co, long long s_ci, int act, float slope, int in_square, int gdn_mode,
               int fixed_point, const float* x, const float* w, const float* bias, const float* gdn_x, float* norm_out,
               float* y, void* workspace, size_t workspace_bytes, cudaStream_t s, const char* name);
size_t tc2_workspace_bytes(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride,
                           int transposed);
int tc_wgrad(int N, int Cs, int Hs, int Ws, int Cb, int Hb, int Wb, int KH, int KW, int stride, int pad, int big_square,
             const float* small, const float* big, float* dw, void* workspace, size_t workspace_bytes, cudaStream_t s,
             const char* name);
size_t tc_wgrad_workspace_bytes(int N, int Cs, int Hs, int Ws, int Cb, int Hb, int Wb, int KH, int KW, int stride);

constexpr int kSmallK = 128;     // fold the taps when (small channel count) * KH * KW <= kSmallK

static inline size_t align1k(size_t v) { return (v + 1023) / 1024 * 1024; }

// out[n, p = (ho, wo), k = (c, r, s)] = x[n, c, ho*st - pad + r, wo*st - pad + s] (0 outside / for k >= C*KH*KW),
// written as split-bf16 hi / lo rows of Cpad channels.  One thread = one bf16x2 pair: writes are fully coalesced,
// reads hit the (tiny) 3-channel tensor in L1/L2; the k -> (c, r, s) decode comes from a shared-memory table so the
// inner loop has one integer division (pixel -> row, column).
constexpr int kFoldMaxK = 256;
__global__ void __launch_bounds__(256)
    im2col_split_kernel(const float* __restrict__ x, int C, int H, int W, int KH, int KW, int st, int pad, int Ho, int Wo,
                        int Cpad, size_t pairs_per_image, __nv_bfloat16* __restrict__ xh, __nv_bfloat16* __restrict__ xl) {
  __shared__ int s_off[kFoldMaxK];     // c*H*W + r*W + s, or -1 for padding channels
  __shared__ short s_r[kFoldMaxK], s_s[kFoldMaxK];
  const int n = blockIdx.y;
  const int KK = KH * KW, Kreal = C * KK, half = Cpad >> 1;
  for (int k = threadIdx.x; k < Cpad; k += blockDim.x) {
    if (k < Kreal) {
      const int c = k / KK, rs = k - c * KK, r = rs / KW, s_ = rs - r * KW;
      s_off[k] = (c * H + r) * W + s_;
      s_r[k] = (short)r;
      s_s[k] = (short)s_;
    } else {
      s_off[k] = -1;
      s_r[k] = s_s[k] = 0;
    }
  }
  __syncthreads();
  const float* xn = x + (size_t)n * C * H * W;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < pairs_per_image; t += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(t / half), kp = (int)(t - (size_t)p * half);
    const int ho = p / Wo, wo = p - ho * Wo;
    const int h0 = ho * st - pad, w0 = wo * st - pad;
    const int base = h0 * W + w0;
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = 2 * kp + e;
      const int off = s_off[k];
      const int hi = h0 + s_r[k], wi = w0 + s_s[k];
      v[e] = (off >= 0 && hi >= 0 && hi < H && wi >= 0 && wi < W) ? __ldg(xn + base + off) : 0.f;
    }
    const __nv_bfloat16 h0b = __float2bfloat16_rn(v[0]), h1b = __float2bfloat16_rn(v[1]);
    const __nv_bfloat16 l0b = __float2bfloat16_rn(v[0] - __bfloat162float(h0b));
    const __nv_bfloat16 l1b = __float2bfloat16_rn(v[1] - __bfloat162float(h1b));
    __nv_bfloat162 hv, lv;
    hv.x = h0b; hv.y = h1b;
    lv.x = l0b; lv.y = l1b;
    const size_t o = ((size_t)n * Ho * Wo + p) * Cpad + 2 * kp;
    *reinterpret_cast<__nv_bfloat162*>(xh + o) = hv;
    *reinterpret_cast<__nv_bfloat162*>(xl + o) = lv;
  }
}

static int stage_im2col(const float* x, int N, int C, int H, int W, int KH, int KW, int st, int pad, int Ho, int Wo, int Cpad,
                        void* xh, void* xl, cudaStream_t s) {
  const size_t pairs = (size_t)Ho * Wo * (Cpad / 2);
  dim3 grid((unsigned)((pairs + 255) / 256 > 4736 ? 4736 : (pairs + 255) / 256), N);
  im2col_split_kernel<<<grid, 256, 0, s>>>(x, C, H, W, KH, KW, st, pad, Ho, Wo, Cpad, pairs,
                                           reinterpret_cast<__nv_bfloat16*>(xh), reinterpret_cast<__nv_bfloat16*>(xl));
  B200_LAUNCH_CHECK("im2col_split_kernel");
  return B200LIC_OK;
}

// y[n, co, ho, wo] = act(bias[co] + sum over taps (r, s) that hit (ho, wo) of col[n, (co, r, s), h, w]),
// h = (ho + pad - r) / st, w = (wo + pad - s) / st.  One thread owns the st x st output block (a, b) -> (st*a + i, st*b + j):
// every tap is read exactly once per block and reads are contiguous along b.
constexpr int kMaxSt = 4;
__global__ void __launch_bounds__(256)
    col2im_kernel(const float* __restrict__ col, const float* __restrict__ bias, int Cout, int H, int W, int KH, int KW,
                  int st, int pad, int Ho, int Wo, int Ab, int Bb, size_t n_blocks, int act, float slope, int fixed_point,
                  float* __restrict__ y) {
  const int KK = KH * KW;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_blocks; t += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(t % Bb);
    const size_t t1 = t / Bb;
    const int a = (int)(t1 % Ab);
    const size_t t2 = t1 / Ab;
    const int co = (int)(t2 % Cout);
    const size_t n = t2 / Cout;
    const float* cn = col + (n * Cout + co) * (size_t)KK * H * W;
    const float bv = bias ? __ldg(bias + co) : 0.f;
    float acc[kMaxSt][kMaxSt];
#pragma unroll
    for (int i = 0; i < kMaxSt; ++i)
#pragma unroll
      for (int j = 0; j < kMaxSt; ++j) acc[i][j] = bv;
    for (int r = 0; r < KH; ++r) {
      // output row st*a + i receives tap r from input row h with st*h - pad + r = st*a + i
      const int i = ((r - pad) % st + st) % st;
      const int h = a - (r - pad - i) / st;
      if (h < 0 || h >= H) continue;
      for (int s_ = 0; s_ < KW; ++s_) {
        const int j = ((s_ - pad) % st + st) % st;
        const int w = b - (s_ - pad - j) / st;
        if (w < 0 || w >= W) continue;
        const float v = __ldg(cn + ((size_t)(r * KW + s_) * H + h) * W + w);
#pragma unroll
        for (int ii = 0; ii < kMaxSt; ++ii)
#pragma unroll
          for (int jj = 0; jj < kMaxSt; ++jj)
            if (ii == i && jj == j) acc[ii][jj] += v;
      }
    }
    float* yn = y + (n * Cout + co) * (size_t)Ho * Wo;
#pragma unroll
    for (int i = 0; i < kMaxSt; ++i) {
      const int ho = st * a + i;
      if (i >= st || ho >= Ho) continue;
#pragma unroll
      for (int j = 0; j < kMaxSt; ++j) {
        const int wo = st * b + j;
        if (j >= st || wo >= Wo) continue;
        float v = apply_act(acc[i][j], act, slope);
        if (fixed_point) v = rintf(fminf(fmaxf(v, -128.f), 128.f) * 256.f) * (1.f / 256.f);
        yn[(size_t)ho * Wo + wo] = v;
      }
    }
  }
}

// The same gather with the geometry known at compile time (the synthesis tail of every codec here is k = 5 or 3,
// stride 2): the tap -> (output phase, input offset) arithmetic folds into constants and the 25 loads of a thread are
// issued back to back.  Same summation order as the generic kernel (taps in (r, s) order), so results are bit-identical.
template <int K, int ST, int PAD>
__global__ void __launch_bounds__(256)
    col2im_fixed_kernel(const float* __restrict__ col, const float* __restrict__ bias, int Cout, int H, int W, int Ho,
                        int Wo, int Ab, int Bb, size_t n_blocks, int act, float slope, int fixed_point,
                        float* __restrict__ y) {
  constexpr int KK = K * K;
  constexpr int pad = PAD;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_blocks; t += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(t % Bb);
    const size_t t1 = t / Bb;
    const int a = (int)(t1 % Ab);
    const size_t t2 = t1 / Ab;
    const int co = (int)(t2 % Cout);
    const size_t n = t2 / Cout;
    const float* cn = col + (n * Cout + co) * (size_t)KK * H * W;
    const float bv = bias ? __ldg(bias + co) : 0.f;
    float v[K][K];
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int i = ((r - pad) % ST + ST) % ST;
      const int h = a - (r - pad - i) / ST;
#pragma unroll
      for (int s_ = 0; s_ < K; ++s_) {
        const int j = ((s_ - pad) % ST + ST) % ST;
        const int w = b - (s_ - pad - j) / ST;
        const bool in = h >= 0 && h < H && w >= 0 && w < W;
        v[r][s_] = in ? __ldg(cn + ((size_t)(r * K + s_) * H + h) * W + w) : 0.f;
      }
    }
    float acc[ST][ST];
#pragma unroll
    for (int i = 0; i < ST; ++i)
#pragma unroll
      for (int j = 0; j < ST; ++j) acc[i][j] = bv;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int i = ((r - pad) % ST + ST) % ST;
      const int h = a - (r - pad - i) / ST;
#pragma unroll
      for (int s_ = 0; s_ < K; ++s_) {
        const int j = ((s_ - pad) % ST + ST) % ST;
        const int w = b - (s_ - pad - j) / ST;
        const bool in = h >= 0 && h < H && w >= 0 && w < W;      // out-of-range taps are skipped, not added as zeros
#pragma unroll
        for (int ii = 0; ii < ST; ++ii)
#pragma unroll
          for (int jj = 0; jj < ST; ++jj)
            if (ii == i && jj == j && in) acc[ii][jj] += v[r][s_];
      }
    }
    float* yn = y + (n * Cout + co) * (size_t)Ho * Wo;
#pragma unroll
    for (int i = 0; i < ST; ++i) {
      const int ho = ST * a + i;
      if (ho >= Ho) continue;
#pragma unroll
      for (int j = 0; j < ST; ++j) {
        const int wo = ST * b + j;
        if (wo >= Wo) continue;
        float o = apply_act(acc[i][j], act, slope);
        if (fixed_point) o = rintf(fminf(fmaxf(o, -128.f), 128.f) * 256.f) * (1.f / 256.f);
        yn[(size_t)ho * Wo + wo] = o;
      }
    }
  }
}

static void launch_col2im(const float* col, const float* bias, int N, int Cout, int H, int W, int KH, int KW, int st, int pad,
                          int Ho, int Wo, int act, float slope, int fixed_point, float* y, cudaStream_t s) {
  const int Ab = (Ho + st - 1) / st, Bb = (Wo + st - 1) / st;
  const size_t n_blocks = (size_t)N * Cout * Ab * Bb;
  const int grid = grid_for(n_blocks, 256, 8);
  if (KH == 5 && KW == 5 && st == 2 && pad == 2)
    col2im_fixed_kernel<5, 2, 2><<<grid, 256, 0, s>>>(col, bias, Cout, H, W, Ho, Wo, Ab, Bb, n_blocks, act, slope,
                                                   fixed_point, y);
  else if (KH == 3 && KW == 3 && st == 2 && pad == 1)
    col2im_fixed_kernel<3, 2, 1><<<grid, 256, 0, s>>>(col, bias, Cout, H, W, Ho, Wo, Ab, Bb, n_blocks, act, slope,
                                                   fixed_point, y);
  else
    col2im_kernel<<<grid, 256, 0, s>>>(col, bias, Cout, H, W, KH, KW, st, pad, Ho, Wo, Ab, Bb, n_blocks, act, slope,
                                       fixed_point, y);
}

// ---- eligibility -------------------------------------------------------------------------------------------------
static inline bool fold_conv(const b200lic_conv_desc* d) {      // fold taps into the INPUT channel axis
  const int KK = d->KH * d->KW;
  return KK > 1 && d->Cin * KK <= kSmallK && !d->in_square && !d->gdn_mode;
}
static inline bool fold_deconv(const b200lic_conv_desc* d) {    // fold taps into the OUTPUT channel axis
  const int KK = d->KH * d->KW;
  return KK > 1 && d->Cout * KK <= kSmallK && d->stride <= kMaxSt && !d->in_square && !d->gdn_mode;
}

// ---- conv forward ---------------------------------------------------------------------------------------------------
size_t smallc_conv_fwd_ws(const b200lic_conv_desc* d) {
  if (!fold_conv(d)) return 0;
  return tc2_workspace_bytes(d->N, d->Cin * d->KH * d->KW, d->Ho, d->Wo, d->Cout, d->Ho, d->Wo, 1, 1, 1, 0);
}
int smallc_conv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void* ws,
                    size_t ws_bytes, cudaStream_t s) {
  if (!fold_conv(d)) return B200LIC_ERR_UNSUPPORTED;
  const int K = d->Cin * d->KH * d->KW, Cpad = (K + 31) / 32 * 32;
  const size_t need = smallc_conv_fwd_ws(d);
  if (need == 0 || !ws || ws_bytes < need) {
    set_error("conv_fwd(tc, folded taps): needs %zu workspace bytes (got %zu)", need, ws_bytes);
    return B200LIC_ERR_UNSUPPORTED;
  }
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  const size_t x_bytes = align1k((size_t)d->N * d->Ho * d->Wo * Cpad * 2);
  int rc = stage_im2col(x, d->N, d->Cin, d->H, d->W, d->KH, d->KW, d->stride, d->pad, d->Ho, d->Wo, Cpad, base,
                        base + x_bytes, s);
  if (rc != B200LIC_OK) return rc;
  // x == nullptr: the activation operand is already staged at the head of the workspace
  return tc2_launch(d->N, K, d->Ho, d->Wo, d->Cout, d->Ho, d->Wo, 1, 1, 1, 0, 0, (long long)K, 1LL, d->act, d->act_slope,
                    0, 0, d->fixed_point, nullptr, w, bias, nullptr, nullptr, y, ws, ws_bytes, s,
                    "conv_fwd(tc, folded taps)");
}

// ---- conv wgrad: dW[co][(ci,r,s)] = sum dy[p, co] * im2col(x)[p, (ci,r,s)] --------------------------------------------
size_t smallc_conv_wgrad_ws(const b200lic_conv_desc* d) {
  if (!fold_conv(d)) return 0;
  return tc_wgrad_workspace_bytes(d->N, d->Cout, d->Ho, d->Wo, d->Cin * d->KH * d->KW, d->Ho, d->Wo, 1, 1, 1);
}
int smallc_conv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                      cudaStream_t s) {
  if (!fold_conv(d)) return B200LIC_ERR_UNSUPPORTED;
  const int K = d->Cin * d->KH * d->KW, CbPad = (K + 63) / 64 * 64, CsPad = (d->Cout + 63) / 64 * 64;
  const size_t need = smallc_conv_wgrad_ws(d);
  if (need == 0 || !ws || ws_bytes < need) {
    set_error("conv_wgrad(tc, folded taps): needs %zu workspace bytes (got %zu)", need, ws_bytes);
    return B200LIC_ERR_UNSUPPORTED;
  }
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  const size_t small_bytes = align1k((size_t)d->N * d->Ho * d->Wo * CsPad * 2);
  const size_t big_bytes = align1k((size_t)d->N * d->Ho * d->Wo * CbPad * 2);
  int rc = stage_im2col(x, d->N, d->Cin, d->H, d->W, d->KH, d->KW, d->stride, d->pad, d->Ho, d->Wo, CbPad,
                        base + 2 * small_bytes, base + 2 * small_bytes + big_bytes, s);
  if (rc != B200LIC_OK) return rc;
  // big == nullptr: the gathered operand is already staged in its workspace slot
  return tc_wgrad(d->N, d->Cout, d->Ho, d->Wo, K, d->Ho, d->Wo, 1, 1, 1, 0, 0, dy, nullptr, dw, ws, ws_bytes, s,
                  "conv_wgrad(tc, folded taps)");
}

// ---- transposed conv forward: col = x . W'' (1x1), y = col2im(col) ---------------------------------------------------
static inline size_t deconv_col_bytes(const b200lic_conv_desc* d) {
  return align1k((size_t)d->N * d->Cout * d->KH * d->KW * d->H * d->W * sizeof(float));
}
size_t smallc_deconv_fwd_ws(const b200lic_conv_desc* d) {
  if (!fold_deconv(d)) return 0;
  const size_t inner = tc2_workspace_bytes(d->N, d->Cin, d->H, d->W, d->Cout * d->KH * d->KW, d->H, d->W, 1, 1, 1, 0);
  return inner ? align1k(inner) + deconv_col_bytes(d) + 1024 : 0;
}
int smallc_deconv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void* ws,
                      size_t ws_bytes, cudaStream_t s) {
  if (!fold_deconv(d)) return B200LIC_ERR_UNSUPPORTED;
  const int Cc = d->Cout * d->KH * d->KW;
  const size_t need = smallc_deconv_fwd_ws(d);
  if (need == 0 || !ws || ws_bytes < need) {
    set_error("deconv_fwd(tc, folded taps): needs %zu workspace bytes (got %zu)", need, ws_bytes);
    return B200LIC_ERR_UNSUPPORTED;
  }
  const size_t inner = align1k(tc2_workspace_bytes(d->N, d->Cin, d->H, d->W, Cc, d->H, d->W, 1, 1, 1, 0));
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  float* col = reinterpret_cast<float*>(base + inner);
  // W''[(co,r,s), ci] = w[ci, co, r, s]: written-channel stride 1, gathered-channel stride Cout*KH*KW
  // (the GEMM writes `col`, not the layer output: a pending b200lic_conv_stats_once request stays pending)
  unsigned* pending_stats = tc2_stats_peek();
  tc2_stats_once(nullptr);
  int rc = tc2_launch(d->N, d->Cin, d->H, d->W, Cc, d->H, d->W, 1, 1, 1, 0, 0, 1LL, (long long)Cc, B200LIC_ACT_NONE, 0.f, 0,
                      0, 0, x, w, nullptr, nullptr, nullptr, col, base, inner, s, "deconv_fwd(tc, folded taps)");
  tc2_stats_once(pending_stats);
  if (rc != B200LIC_OK) return rc;
  launch_col2im(col, bias, d->N, d->Cout, d->H, d->W, d->KH, d->KW, d->stride, d->pad, d->Ho, d->Wo, d->act, d->act_slope,
                d->fixed_point, y, s);
  B200_LAUNCH_CHECK("col2im_kernel");
  return B200LIC_OK;
}

// ---- transposed conv wgrad: dW[ci][(co,r,s)] = sum x[p, ci] * im2col(dy)[p, (co,r,s)] ----------------------------------
size_t smallc_deconv_wgrad_ws(const b200lic_conv_desc* d) {
  if (!fold_deconv(d)) return 0;
  return tc_wgrad_workspace_bytes(d->N, d->Cin, d->H, d->W, d->Cout * d->KH * d->KW, d->H, d->W, 1, 1, 1);
}
int smallc_deconv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                        cudaStream_t s) {
  if (!fold_deconv(d)) return B200LIC_ERR_UNSUPPORTED;
  const int K = d->Cout * d->KH * d->KW, CbPad = (K + 63) / 64 * 64, CsPad = (d->Cin + 63) / 64 * 64;
  const size_t need = smallc_deconv_wgrad_ws(d);
  if (need == 0 || !ws || ws_bytes < need) {
    set_error("deconv_wgrad(tc, folded taps): needs %zu workspace bytes (got %zu)", need, ws_bytes);
    return B200LIC_ERR_UNSUPPORTED;
  }
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  const size_t small_bytes = align1k((size_t)d->N * d->H * d->W * CsPad * 2);
  const size_t big_bytes = align1k((size_t)d->N * d->H * d->W * CbPad * 2);
  // im2col of dy with the conv geometry whose "output" grid is the transposed conv's input grid (H x W)
  int rc = stage_im2col(dy, d->N, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, d->H, d->W, CbPad,
                        base + 2 * small_bytes, base + 2 * small_bytes + big_bytes, s);
  if (rc != B200LIC_OK) return rc;
  return tc_wgrad(d->N, d->Cin, d->H, d->W, K, d->H, d->W, 1, 1, 1, 0, 0, x, nullptr, dw, ws, ws_bytes, s,
                  "deconv_wgrad(tc, folded taps)");
}

}  // namespace b200lic

using namespace b200lic;

extern "C" {

// The two data movements of the folded-tap layers as entry points of their own, so that the fused AdaRound iteration can
// run those layers as 1x1 problems on prepared operands (recon.py FusedFolded): im2col straight into a staged operand
// slot, and the col2im gather that finishes a folded transposed convolution.
int b200lic_im2col_stage(const float* x, int N, int C, int H, int W, int KH, int KW, int stride, int pad, int Ho, int Wo,
                         void* x_hi, void* x_lo, int cpad, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && x_hi && x_lo && N > 0 && N <= 65535 && C > 0 && H > 0 && W > 0 && KH > 0 && KW > 0 && stride > 0 &&
                   Ho > 0 && Wo > 0,
               "im2col_stage: bad arguments");
  B200_REQUIRE(cpad % 32 == 0 && cpad <= kFoldMaxK && C * KH * KW <= cpad, "im2col_stage: cpad=%d for %d folded channels",
               cpad, C * KH * KW);
  return stage_im2col(x, N, C, H, W, KH, KW, stride, pad, Ho, Wo, cpad, x_hi, x_lo, as_stream(stream));
}

int b200lic_col2im(const float* col, const float* bias, int N, int Cout, int H, int W, int KH, int KW, int stride, int pad,
                   int Ho, int Wo, int act, float slope, int fixed_point, float* y, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(col && y && N > 0 && Cout > 0 && H > 0 && W > 0 && KH > 0 && KW > 0 && Ho > 0 && Wo > 0,
               "col2im: bad arguments");
  B200_REQUIRE(stride >= 1 && stride <= kMaxSt, "col2im: stride %d outside [1,%d]", stride, kMaxSt);
  launch_col2im(col, bias, N, Cout, H, W, KH, KW, stride, pad, Ho, Wo, act, slope, fixed_point, y, as_stream(stream));
  B200_LAUNCH_CHECK("col2im_kernel");
  return B200LIC_OK;
}

}  // extern "C"
