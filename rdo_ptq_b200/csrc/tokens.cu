// Token-major ([rows, C], channels contiguous) pieces of the reference's Linear / LayerNorm wrappers
// (TO quantization/quant_layer.py:38-49,117-121: QuantModule over nn.Linear and nn.LayerNorm, used by the Swin blocks of
// quant_block.py:330-641).  A Linear over tokens is a 1x1 convolution over rows "images" of one pixel:
// the token matrix IS the NHWC operand of the conv engine, so the GEMM runs on the tcgen05 engine unchanged and only the
// operand staging differs -- stage_tokens_kernel splits the fp32 rows into the bf16 hi / lo slabs in place of the NCHW ->
// NHWC transposition.  LayerNorm and GELU are one pass each.
#include <cuda_bf16.h>
#include "common.cuh"

namespace b200lic {

// x [rows, C] fp32 -> xh / xl [rows, cpad] bf16 (hi = bf16(x), lo = bf16(x - hi)); channels C..cpad-1 are zero.
__global__ void __launch_bounds__(256)
    stage_tokens_kernel(const float* __restrict__ x, size_t rows, int C, int cpad, __nv_bfloat16* __restrict__ xh,
                        __nv_bfloat16* __restrict__ xl) {
  const size_t pairs = rows * (size_t)(cpad >> 1);
  const int half = cpad >> 1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / half;
    const int c = (int)(i - r * half) * 2;
    const float a = c < C ? __ldg(x + r * C + c) : 0.f;
    const float b = c + 1 < C ? __ldg(x + r * C + c + 1) : 0.f;
    const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
    __nv_bfloat162 hv, lv;
    hv.x = ah;
    hv.y = bh;
    lv.x = __float2bfloat16_rn(a - __bfloat162float(ah));
    lv.y = __float2bfloat16_rn(b - __bfloat162float(bh));
    *reinterpret_cast<__nv_bfloat162*>(xh + r * cpad + c) = hv;
    *reinterpret_cast<__nv_bfloat162*>(xl + r * cpad + c) = lv;
  }
}

// F.layer_norm over the last axis: one warp per row, two passes over registers / L1 (mean, then centred variance).
__global__ void __launch_bounds__(256)
    layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         size_t rows, int C, float eps, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const float* xr = x + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __ldg(xr + c);
    const float mean = warp_sum(s) / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = __ldg(xr + c) - mean;
      v = fmaf(d, d, v);
    }
    const float rstd = 1.f / sqrtf(warp_sum(v) / (float)C + eps);
    float* yr = y + r * C;
    for (int c = lane; c < C; c += 32) {
      const float n = (__ldg(xr + c) - mean) * rstd;
      yr[c] = gamma ? fmaf(n, __ldg(gamma + c), beta ? __ldg(beta + c) : 0.f) : n;
    }
  }
}

// nn.GELU() (exact, erf form): 0.5 x (1 + erf(x / sqrt 2))
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const float* __restrict__ x, size_t n, float* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = __ldg(x + i);
    y[i] = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  }
}

// ---- dynamic per-channel activation quantiser on token-major tensors ------------------------------------------------------
// TO quantizer.py:81-121 `ActQuant` on a 3-D [B, L, C] tensor quantises per LAST-axis channel; for [rows, C] rows that is
// a reduction down the columns.  Same arithmetic as act_quant.cu (actq_one: IEEE division, rint, explicit _rn ops), same
// key layout; threads run along the channels, so every load is coalesced.
constexpr int kTokRows = 128;              // rows per CTA of the statistics pass

__global__ void actq_tokens_init_kernel(unsigned* keys, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    keys[2 * i] = 0xffffffffu;
    keys[2 * i + 1] = 0u;
  }
}

__global__ void __launch_bounds__(256)
    actq_tokens_stats_kernel(const float* __restrict__ x, size_t rows, int C, unsigned* __restrict__ keys) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const size_t r0 = (size_t)blockIdx.y * kTokRows;
  const size_t r1 = r0 + kTokRows < rows ? r0 + kTokRows : rows;
  float mn = INFINITY, mx = -INFINITY;
  for (size_t r = r0; r < r1; ++r) {
    const float v = __ldg(x + r * C + c);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  atomicMin(keys + 2 * c, f2key(mn));
  atomicMax(keys + 2 * c + 1, f2key(mx));
}

__global__ void __launch_bounds__(256)
    actq_tokens_apply_kernel(const float* __restrict__ x, const unsigned* __restrict__ keys, size_t n, int C, float L,
                             float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (size_t)C);
    const float m = key2f(keys[2 * c]);
    const float r = fmaxf(__fsub_rn(key2f(keys[2 * c + 1]), m), 1e-6f);
    float t = __fdiv_rn(__fsub_rn(__ldg(x + i), m), r);
    t = fminf(fmaxf(t, -1.f), 1.f);
    const float q = rintf(__fmul_rn(t, L));
    out[i] = __fadd_rn(__fmul_rn(__fdiv_rn(q, L), r), m);
  }
}

// ---- window attention core (TO models/layers.py:137-166, quant_block.py:383-418) -----------------------------------------
// qkv [B_, N, 3C] (the output of the qkv Linear: q | k | v, each [nH, hd] per token).  One CTA per (window b, head h).
// P[b,h,i,j] = softmax_j( (q_i * scale) . k_j + bias[h,i,j] + mask[b % nW, i, j] )
// out[b, i, h*hd + d] = sum_j P[b,h,i,j] * v[j,d]        (= (attn @ v).transpose(1, 2).reshape(B_, N, C))
// The two halves are separate kernels because the reference quantises P (dynamic, per head over ALL windows) in between.
constexpr int kAttnMaxN = 64, kAttnMaxD = 48;   // Lu2022: 8 x 8 windows, head dimension 12 ... 48 (192 channels on 4 heads)

__global__ void __launch_bounds__(256)
    window_attn_softmax_kernel(const float* __restrict__ qkv, const float* __restrict__ bias, const float* __restrict__ mask,
                               int N, int C, int nH, int nW, float scale, float* __restrict__ P) {
  __shared__ float qs[kAttnMaxN][kAttnMaxD + 1], ks[kAttnMaxN][kAttnMaxD + 1], S[kAttnMaxN][kAttnMaxN + 1];
  const int b = blockIdx.x / nH, h = blockIdx.x % nH, hd = C / nH;
  const float* base = qkv + (size_t)b * N * 3 * C + (size_t)h * hd;
  for (int e = threadIdx.x; e < N * hd; e += blockDim.x) {
    const int i = e / hd, d = e - i * hd;
    qs[i][d] = __ldg(base + (size_t)i * 3 * C + d) * scale;
    ks[i][d] = __ldg(base + (size_t)i * 3 * C + C + d);
  }
  __syncthreads();
  const float* bh = bias + (size_t)h * N * N;
  const float* mw = mask ? mask + (size_t)(b % nW) * N * N : nullptr;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int i = e / N, j = e - i * N;
    float acc = 0.f;
    for (int d = 0; d < hd; ++d) acc = fmaf(qs[i][d], ks[j][d], acc);
    acc += __ldg(bh + e);
    if (mw) acc += __ldg(mw + e);
    S[i][j] = acc;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* Pb = P + ((size_t)b * nH + h) * N * N;
  for (int i = warp; i < N; i += nwarps) {
    const float a0 = lane < N ? S[i][lane] : -INFINITY, a1 = lane + 32 < N ? S[i][lane + 32] : -INFINITY;
    const float mx = warp_max(fmaxf(a0, a1));
    const float e0 = lane < N ? expf(a0 - mx) : 0.f, e1 = lane + 32 < N ? expf(a1 - mx) : 0.f;
    const float sum = warp_sum(e0 + e1);
    if (lane < N) Pb[(size_t)i * N + lane] = e0 / sum;
    if (lane + 32 < N) Pb[(size_t)i * N + lane + 32] = e1 / sum;
  }
}

__global__ void __launch_bounds__(256)
    window_attn_apply_kernel(const float* __restrict__ P, const float* __restrict__ qkv, int N, int C, int nH,
                             float* __restrict__ out) {
  __shared__ float Ps[kAttnMaxN][kAttnMaxN + 1], vs[kAttnMaxN][kAttnMaxD + 1];
  const int b = blockIdx.x / nH, h = blockIdx.x % nH, hd = C / nH;
  const float* Pb = P + ((size_t)b * nH + h) * N * N;
  const float* vbase = qkv + (size_t)b * N * 3 * C + 2 * (size_t)C + (size_t)h * hd;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) Ps[e / N][e % N] = __ldg(Pb + e);
  for (int e = threadIdx.x; e < N * hd; e += blockDim.x) {
    const int j = e / hd, d = e - j * hd;
    vs[j][d] = __ldg(vbase + (size_t)j * 3 * C + d);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < N * hd; e += blockDim.x) {
    const int i = e / hd, d = e - i * hd;
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(Ps[i][j], vs[j][d], acc);
    out[((size_t)b * N + i) * C + (size_t)h * hd + d] = acc;
  }
}

}  // namespace b200lic

using namespace b200lic;

extern "C" {

int b200lic_stage_tokens(const float* x, size_t rows, int C, int cpad, void* x_hi, void* x_lo, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && x_hi && x_lo && rows > 0 && C > 0, "stage_tokens: bad arguments");
  B200_REQUIRE(cpad >= C && (cpad % 32) == 0, "stage_tokens: cpad=%d for %d channels", cpad, C);
  stage_tokens_kernel<<<grid_for(rows * (size_t)(cpad / 2), 256), 256, 0, as_stream(stream)>>>(
      x, rows, C, cpad, reinterpret_cast<__nv_bfloat16*>(x_hi), reinterpret_cast<__nv_bfloat16*>(x_lo));
  B200_LAUNCH_CHECK("stage_tokens_kernel");
  return B200LIC_OK;
}

int b200lic_layernorm_fwd(const float* x, const float* gamma, const float* beta, size_t rows, int C, float eps, float* y,
                          b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && y && rows > 0 && C > 0, "layernorm_fwd: bad arguments");
  B200_REQUIRE(eps >= 0.f, "layernorm_fwd: eps=%g", (double)eps);
  layernorm_fwd_kernel<<<grid_for(rows * 32, 256), 256, 0, as_stream(stream)>>>(x, gamma, beta, rows, C, eps, y);
  B200_LAUNCH_CHECK("layernorm_fwd_kernel");
  return B200LIC_OK;
}

int b200lic_actq_tokens(const float* x, size_t rows, int C, int n_bits, float* minmax, float* out, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && minmax && out && rows > 0 && C > 0, "actq_tokens: bad arguments");
  B200_REQUIRE(n_bits >= 2 && n_bits <= 16, "actq_tokens: n_bits=%d outside [2,16]", n_bits);
  const size_t row_blocks = (rows + kTokRows - 1) / kTokRows;
  B200_REQUIRE(row_blocks <= 65535, "actq_tokens: too many rows");
  unsigned* keys = reinterpret_cast<unsigned*>(minmax);
  actq_tokens_init_kernel<<<(C + 255) / 256, 256, 0, as_stream(stream)>>>(keys, C);
  B200_LAUNCH_CHECK("actq_tokens_init_kernel");
  actq_tokens_stats_kernel<<<dim3((unsigned)((C + 255) / 256), (unsigned)row_blocks), 256, 0, as_stream(stream)>>>(x, rows, C,
                                                                                                                   keys);
  B200_LAUNCH_CHECK("actq_tokens_stats_kernel");
  const size_t n = rows * (size_t)C;
  actq_tokens_apply_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, keys, n, C, (float)((1 << n_bits) - 1), out);
  B200_LAUNCH_CHECK("actq_tokens_apply_kernel");
  return B200LIC_OK;
}

static int window_attn_check(const char* name, int B_, int N, int C, int nH) {
  B200_REQUIRE(B_ > 0 && N > 0 && C > 0 && nH > 0 && C % nH == 0, "%s: bad shape", name);
  if (N > kAttnMaxN || C / nH > kAttnMaxD) {
    set_error("%s: window of %d tokens x head dimension %d exceeds %d x %d", name, N, C / nH, kAttnMaxN, kAttnMaxD);
    return B200LIC_ERR_UNSUPPORTED;
  }
  B200_REQUIRE((long long)B_ * nH < 2147483647LL, "%s: too many windows", name);
  return B200LIC_OK;
}

int b200lic_window_attn_softmax(const float* qkv, const float* bias, const float* mask, int B_, int N, int C, int nH, int nW,
                                float scale, float* P, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(qkv && bias && P, "window_attn_softmax: null pointer");
  int rc = window_attn_check("window_attn_softmax", B_, N, C, nH);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(!mask || (nW > 0 && B_ % nW == 0), "window_attn_softmax: %d windows do not tile a batch of %d", nW, B_);
  window_attn_softmax_kernel<<<(unsigned)(B_ * nH), 256, 0, as_stream(stream)>>>(qkv, bias, mask, N, C, nH, mask ? nW : 1,
                                                                                 scale, P);
  B200_LAUNCH_CHECK("window_attn_softmax_kernel");
  return B200LIC_OK;
}

int b200lic_window_attn_apply(const float* P, const float* qkv, int B_, int N, int C, int nH, float* out,
                              b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(P && qkv && out, "window_attn_apply: null pointer");
  int rc = window_attn_check("window_attn_apply", B_, N, C, nH);
  if (rc != B200LIC_OK) return rc;
  window_attn_apply_kernel<<<(unsigned)(B_ * nH), 256, 0, as_stream(stream)>>>(P, qkv, N, C, nH, out);
  B200_LAUNCH_CHECK("window_attn_apply_kernel");
  return B200LIC_OK;
}

int b200lic_gelu_fwd(const float* x, size_t n, float* y, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && y && n > 0, "gelu_fwd: bad arguments");
  gelu_fwd_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, n, y);
  B200_LAUNCH_CHECK("gelu_fwd_kernel");
  return B200LIC_OK;
}

}  // extern "C"
