// MS-SSIM of two [N,C,H,W] fp32 images (evaluation metric of the reference's entry points: pytorch_msssim.ms_ssim at
// TO/losses/losses.py:26,31,49-52, LU/quantize.py:89, LU/quant.py:86, LU/dataset_test.py:60-61).
//
// One level = five 11-tap separable Gaussian (sigma 1.5) VALID filterings (x, y, x*x, y*y, x*y; height axis first like the
// package's gaussian_filter), the ssim / cs maps and their per-plane means.  ssim_level_kernel does all of that in one
// pass over the two images: a CTA loads a (16+10) x (32+10) patch of x and y into shared memory, filters the five
// quantities down the columns into shared memory, then along the rows in registers, forms the two maps and reduces them
// (fp64 partial sums, one atomicAdd pair per CTA).  8 B/pixel of HBM traffic per level: the kernel is bound by the
// 110 FMAs per pixel, ~0.1 ms for a 768x512 RGB image.  avg_pool2_kernel is the 2x2 mean between levels (padding =
// size % 2, zeros counted), msssim_combine_kernel the weighted product over the five levels.
#include "common.cuh"

namespace b200lic {

constexpr int kSsimWin = 11;
constexpr int kSsimTW = 32, kSsimTH = 16;
constexpr int kSsimPW = kSsimTW + kSsimWin - 1, kSsimPH = kSsimTH + kSsimWin - 1;      // 42 x 26 input patch

__global__ void __launch_bounds__(256)
    ssim_level_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ win, int H,
                      int W, float c1, float c2, double* __restrict__ sums) {
  __shared__ float sx[kSsimPH][kSsimPW + 1], sy[kSsimPH][kSsimPW + 1];
  __shared__ float v[5][kSsimTH][kSsimPW + 1];
  __shared__ float w[kSsimWin];
  __shared__ double red[2][8];
  const int plane = blockIdx.z;
  const int x0 = blockIdx.x * kSsimTW, y0 = blockIdx.y * kSsimTH;
  const int Ho = H - (kSsimWin - 1), Wo = W - (kSsimWin - 1);
  const float* xp = x + (size_t)plane * H * W;
  const float* yp = y + (size_t)plane * H * W;
  if (threadIdx.x < kSsimWin) w[threadIdx.x] = win[threadIdx.x];
  for (int i = threadIdx.x; i < kSsimPH * kSsimPW; i += 256) {
    const int r = i / kSsimPW, c = i - r * kSsimPW;
    const int gy = y0 + r, gx = x0 + c;
    const bool in = gy < H && gx < W;
    sx[r][c] = in ? __ldg(xp + (size_t)gy * W + gx) : 0.f;
    sy[r][c] = in ? __ldg(yp + (size_t)gy * W + gx) : 0.f;
  }
  __syncthreads();
  // height axis first
  for (int i = threadIdx.x; i < kSsimTH * kSsimPW; i += 256) {
    const int r = i / kSsimPW, c = i - r * kSsimPW;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
    for (int k = 0; k < kSsimWin; ++k) {
      const float wk = w[k], xv = sx[r + k][c], yv = sy[r + k][c];
      a0 = fmaf(wk, xv, a0);
      a1 = fmaf(wk, yv, a1);
      a2 = fmaf(wk, xv * xv, a2);
      a3 = fmaf(wk, yv * yv, a3);
      a4 = fmaf(wk, xv * yv, a4);
    }
    v[0][r][c] = a0;
    v[1][r][c] = a1;
    v[2][r][c] = a2;
    v[3][r][c] = a3;
    v[4][r][c] = a4;
  }
  __syncthreads();
  double s_ssim = 0.0, s_cs = 0.0;
  for (int i = threadIdx.x; i < kSsimTH * kSsimTW; i += 256) {
    const int r = i / kSsimTW, c = i - r * kSsimTW;
    if (y0 + r >= Ho || x0 + c >= Wo) continue;
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < kSsimWin; ++k) {
      const float wk = w[k];
      m1 = fmaf(wk, v[0][r][c + k], m1);
      m2 = fmaf(wk, v[1][r][c + k], m2);
      e11 = fmaf(wk, v[2][r][c + k], e11);
      e22 = fmaf(wk, v[3][r][c + k], e22);
      e12 = fmaf(wk, v[4][r][c + k], e12);
    }
    const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
    const float s1 = e11 - m11, s2 = e22 - m22, s12 = e12 - m12;
    const float cs = (2.f * s12 + c2) / (s1 + s2 + c2);
    const float ss = ((2.f * m12 + c1) / (m11 + m22 + c1)) * cs;
    s_ssim += (double)ss;
    s_cs += (double)cs;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_ssim += __shfl_xor_sync(0xffffffffu, s_ssim, o);
    s_cs += __shfl_xor_sync(0xffffffffu, s_cs, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    red[0][wid] = s_ssim;
    red[1][wid] = s_cs;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) {
      a += red[0][i];
      b += red[1][i];
    }
    atomicAdd(sums + 2 * plane, a);
    atomicAdd(sums + 2 * plane + 1, b);
  }
}

__global__ void __launch_bounds__(256)
    avg_pool2_kernel(const float* __restrict__ x, int H, int W, int ph, int pw, int Ho, int Wo, size_t total,
                     float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ow = (int)(i % Wo);
    const size_t t = i / Wo;
    const int oh = (int)(t % Ho);
    const size_t plane = t / Ho;
    const float* p = x + plane * (size_t)H * W;
    const int h0 = 2 * oh - ph, w0 = 2 * ow - pw;
    float s = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int hh = h0 + dy, ww = w0 + dx;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) s += __ldg(p + (size_t)hh * W + ww);
      }
    out[i] = s * 0.25f;
  }
}

// sums: [levels][planes][2] (ssim, cs) totals; inv_count[l] = 1 / (Ho_l * Wo_l).  per_plane[p] = prod_l relu(v_l)^w_l with
// v_l = cs mean for l < levels-1 and the ssim mean for the last level; mean[0] = mean over planes.
__global__ void __launch_bounds__(256)
    msssim_combine_kernel(const double* __restrict__ sums, const double* __restrict__ inv_count, int levels, int planes,
                          float* __restrict__ per_plane, float* __restrict__ mean) {
  const float wts[5] = {0.0448f, 0.2856f, 0.3001f, 0.2363f, 0.1333f};
  __shared__ float red[32];
  float acc = 0.f;
  for (int p = threadIdx.x; p < planes; p += blockDim.x) {
    float prod = 1.f;
    for (int l = 0; l < levels; ++l) {
      const double* s = sums + ((size_t)l * planes + p) * 2;
      const float val = (float)((l == levels - 1 ? s[0] : s[1]) * inv_count[l]);
      prod *= powf(fmaxf(val, 0.f), wts[l]);
    }
    per_plane[p] = prod;
    acc += prod;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) mean[0] = acc / (float)planes;
}

}  // namespace b200lic

using namespace b200lic;

extern "C" {

int b200lic_ssim_level(const float* x, const float* y, const float* win11, int planes, int H, int W, float c1, float c2,
                       double* sums, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && y && win11 && sums && planes > 0 && planes <= 65535, "ssim_level: bad arguments");
  B200_REQUIRE(H >= kSsimWin && W >= kSsimWin, "ssim_level: %dx%d is smaller than the %d-tap window", H, W, kSsimWin);
  const int Ho = H - (kSsimWin - 1), Wo = W - (kSsimWin - 1);
  dim3 grid((unsigned)((Wo + kSsimTW - 1) / kSsimTW), (unsigned)((Ho + kSsimTH - 1) / kSsimTH), (unsigned)planes);
  B200_REQUIRE(grid.y <= 65535, "ssim_level: image too tall");
  ssim_level_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, win11, H, W, c1, c2, sums);
  B200_LAUNCH_CHECK("ssim_level_kernel");
  return B200LIC_OK;
}

int b200lic_avg_pool2(const float* x, int planes, int H, int W, int pad_h, int pad_w, float* out,
                      b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && out && planes > 0 && H > 0 && W > 0, "avg_pool2: bad arguments");
  B200_REQUIRE((pad_h == 0 || pad_h == 1) && (pad_w == 0 || pad_w == 1), "avg_pool2: padding must be 0 or 1");
  const int Ho = (H + 2 * pad_h - 2) / 2 + 1, Wo = (W + 2 * pad_w - 2) / 2 + 1;
  const size_t total = (size_t)planes * Ho * Wo;
  avg_pool2_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(x, H, W, pad_h, pad_w, Ho, Wo, total, out);
  B200_LAUNCH_CHECK("avg_pool2_kernel");
  return B200LIC_OK;
}

int b200lic_msssim_combine(const double* sums, const double* inv_count, int levels, int planes, float* per_plane,
                           float* mean, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(sums && inv_count && per_plane && mean && planes > 0, "msssim_combine: bad arguments");
  B200_REQUIRE(levels >= 1 && levels <= 5, "msssim_combine: levels=%d outside [1,5]", levels);
  msssim_combine_kernel<<<1, 256, 0, as_stream(stream)>>>(sums, inv_count, levels, planes, per_plane, mean);
  B200_LAUNCH_CHECK("msssim_combine_kernel");
  return B200LIC_OK;
}

}  // extern "C"
