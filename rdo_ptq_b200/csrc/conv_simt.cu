// fp32 CUDA-core implicit-GEMM engine for conv / transposed conv forward, dgrad and wgrad (B200LIC_ENGINE_SIMT).
//
// This is the exact-fp32 engine: the shapes tcgen05 does not take (Cin=3 / Cout=3 ends of the codec, tiny
// hyper-prior maps) and the cross-check for the tensor-core engine (conv_tc.cu).  One gather formulation
// covers all data-path ops:
//   conv fwd            y[n,co,ho,wo] = sum_{ci,r,s} w[co,ci,r,s] x[n,ci,ho*st-p+r, wo*st-p+s]
//   transposed-conv fwd is decomposed into st*st output phases; inside a phase it is a stride-1 gather
//                       with the taps r = r0 + st*i that hit that phase (no wasted zero taps)
//   conv dgrad          = transposed-conv fwd with the roles of the weight's two channel axes swapped
//   transposed dgrad    = conv fwd with the roles swapped
//   wgrad (both)        dW[cs,cb,r,s] = sum_pixels small[n,cs,h,w] * big[n,cb,h*st-p+r, w*st-p+s]
// Replaces cuDNN under F.conv2d / F.conv_transpose2d (TO quant_layer.py:28,36,123) and their autograd.
#include "common.cuh"

namespace b200lic {

constexpr int kMaxTaps = 128;

struct GatherGeom {
  int N, Cin, H, W;       // gathered tensor
  int Cout, Ho, Wo;       // written tensor
  int KH, KW, stride, pad;
  int transposed;         // 0: conv-style gather, 1: phase-decomposed transposed conv
  long long w_co_stride, w_ci_stride;  // weight element strides of the written / gathered channel axes
  int act;
  float slope;
  int in_square, gdn_mode, fixed_point;
};

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
    gather_gemm_kernel(GatherGeom g, const float* __restrict__ x, const float* __restrict__ w,
                       const float* __restrict__ bias, const float* __restrict__ gdn_x, float* __restrict__ norm_out,
                       float* __restrict__ y) {
  constexpr int THREADS = (BM / TM) * (BN / TN);
  constexpr int TXN = BN / TN;  // threads along the pixel axis
  static_assert(THREADS % BN == 0, "B-tile loader needs THREADS % BN == 0");
  constexpr int KSTEP_B = THREADS / BN;
  constexpr int LB = BK / KSTEP_B;
  static_assert(BK % KSTEP_B == 0, "BK must be a multiple of THREADS/BN");
  constexpr int LA = (BM * BK + THREADS - 1) / THREADS;

  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  __shared__ int tap_in[kMaxTaps];   // (dh << 16) | (dw & 0xffff) offsets into the gathered tensor
  __shared__ int tap_w[kMaxTaps];    // offset of the tap inside one [KH,KW] filter

  const int tid = threadIdx.x;
  // ---- phase geometry --------------------------------------------------------------------------------------
  int ph = 0, pw = 0, r0 = 0, s0 = 0, KHp = g.KH, KWp = g.KW, Pa = g.Ho, Pb = g.Wo;
  int in_step = g.stride, tap_step = 1, base_h = -g.pad, base_w = -g.pad, out_step = 1, r_step = 1;
  if (g.transposed) {
    const int st = g.stride;
    ph = blockIdx.z / st;
    pw = blockIdx.z % st;
    r0 = (ph + g.pad) % st;
    s0 = (pw + g.pad) % st;
    KHp = r0 < g.KH ? (g.KH - r0 + st - 1) / st : 0;
    KWp = s0 < g.KW ? (g.KW - s0 + st - 1) / st : 0;
    Pa = ph < g.Ho ? (g.Ho - ph + st - 1) / st : 0;
    Pb = pw < g.Wo ? (g.Wo - pw + st - 1) / st : 0;
    in_step = 1;
    tap_step = -1;
    base_h = (ph + g.pad - r0) / st;
    base_w = (pw + g.pad - s0) / st;
    out_step = st;
    r_step = st;
  }
  const int T = KHp * KWp;
  const int K = g.Cin * T;
  const long long P = (long long)g.N * Pa * Pb;
  const long long p0 = (long long)blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  if (p0 >= P) return;

  for (int t = tid; t < T; t += THREADS) {
    const int i = t / KWp, j = t % KWp;
    tap_in[t] = ((i * tap_step) << 16) | ((j * tap_step) & 0xffff);
    tap_w[t] = (r0 + i * r_step) * g.KW + (s0 + j * r_step);
  }
  __syncthreads();

  // ---- per-thread loader state ----------------------------------------------------------------------------------
  const int pl = tid % BN;
  const long long p = p0 + pl;
  const bool p_ok = p < P;
  int hbase = 0, wbase = 0;
  long long xoff_n = 0;
  if (p_ok) {
    const int n = (int)(p / ((long long)Pa * Pb));
    const int rem = (int)(p - (long long)n * Pa * Pb);
    const int a = rem / Pb, b = rem - a * Pb;
    hbase = a * in_step + base_h;
    wbase = b * in_step + base_w;
    xoff_n = (long long)n * g.Cin * g.H * g.W;
  }
  int b_ci[LB], b_t[LB];
#pragma unroll
  for (int j = 0; j < LB; ++j) {
    const int kl = tid / BN + j * KSTEP_B;
    b_ci[j] = T > 0 ? kl / T : 0;
    b_t[j] = T > 0 ? kl % T : 0;
  }
  int a_ci[LA], a_t[LA], a_m[LA], a_k[LA];
#pragma unroll
  for (int j = 0; j < LA; ++j) {
    const int e = tid + j * THREADS;
    a_k[j] = e % BK;
    a_m[j] = e / BK;  // may be >= BM when BM*BK < THREADS
    a_ci[j] = T > 0 ? a_k[j] / T : 0;
    a_t[j] = T > 0 ? a_k[j] % T : 0;
  }

  float ra[LA], rb[LB];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int j = 0; j < LB; ++j) {
      const int k = k0 + tid / BN + j * KSTEP_B;
      float v = 0.f;
      if (p_ok && k < K) {
        const int pk = tap_in[b_t[j]];
        const int hi = hbase + (pk >> 16);
        const int wi = wbase + (int)(short)(pk & 0xffff);
        if ((unsigned)hi < (unsigned)g.H && (unsigned)wi < (unsigned)g.W) {
          v = __ldg(x + xoff_n + ((long long)b_ci[j] * g.H + hi) * g.W + wi);
          if (g.in_square) v *= v;
        }
      }
      rb[j] = v;
      b_t[j] += BK;
      while (T > 0 && b_t[j] >= T) {
        b_t[j] -= T;
        ++b_ci[j];
      }
    }
#pragma unroll
    for (int j = 0; j < LA; ++j) {
      float v = 0.f;
      const int m = m0 + a_m[j];
      if (a_m[j] < BM && m < g.Cout && k0 + a_k[j] < K)
        v = __ldg(w + (long long)m * g.w_co_stride + (long long)a_ci[j] * g.w_ci_stride + tap_w[a_t[j]]);
      ra[j] = v;
      a_t[j] += BK;
      while (T > 0 && a_t[j] >= T) {
        a_t[j] -= T;
        ++a_ci[j];
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int j = 0; j < LB; ++j) Bs[buf][tid / BN + j * KSTEP_B][pl] = rb[j];
#pragma unroll
    for (int j = 0; j < LA; ++j)
      if (a_m[j] < BM) As[buf][a_k[j]][a_m[j]] = ra[j];
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int tx = tid % TXN, ty = tid / TXN;
  const int nk = (K + BK - 1) / BK;
  if (nk > 0) {
    load_tile(0);
    store_tile(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float af[TM], bf[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) af[i] = As[buf][kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bf[j] = Bs[buf][kk][tx + j * TXN];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: bias, GDN tail, activation, fixed point ----------------------------------------------------------
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const long long pp = p0 + tx + j * TXN;
    if (pp >= P) continue;
    const int n = (int)(pp / ((long long)Pa * Pb));
    const int rem = (int)(pp - (long long)n * Pa * Pb);
    const int a = rem / Pb, b = rem - a * Pb;
    const int ho = a * out_step + ph, wo = b * out_step + pw;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int co = m0 + ty * TM + i;
      if (co >= g.Cout) continue;
      const long long idx = (((long long)n * g.Cout + co) * g.Ho + ho) * g.Wo + wo;
      float v = acc[i][j] + (bias ? __ldg(bias + co) : 0.f);
      if (g.gdn_mode) {
        if (norm_out) norm_out[idx] = v;
        const float xv = __ldg(gdn_x + idx);
        v = g.gdn_mode == 1 ? xv * rsqrtf(v) : xv * sqrtf(v);
      }
      v = apply_act(v, g.act, g.slope);
      if (g.fixed_point) v = rintf(fminf(fmaxf(v, -128.f), 128.f) * 256.f) * (1.f / 256.f);
      y[idx] = v;
    }
  }
}

// ---- wgrad ----------------------------------------------------------------------------------------------------
struct WgradGeom {
  int N;
  int Cs, Hs, Ws;   // "small" tensor: indexed directly (conv: dy; transposed conv: x)
  int Cb, Hb, Wb;   // "big" tensor: gathered at (h*st - pad + r, w*st - pad + s) (conv: x; transposed: dy)
  int KH, KW, stride, pad;
  int big_square;   // GDN: the conv input is x*x
  int splits;       // split-K factor over pixels (atomic accumulation when > 1)
  long long chunk;  // pixels per split (multiple of BK)
};

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
    wgrad_kernel(WgradGeom g, const float* __restrict__ small, const float* __restrict__ big, float* __restrict__ dw) {
  constexpr int THREADS = (BM / TM) * (BN / TN);
  constexpr int TXN = BN / TN;
  static_assert(THREADS % BK == 0, "loader needs THREADS % BK == 0");
  constexpr int ROWS = THREADS / BK;       // tile rows (m or q) covered per pass
  constexpr int LA = BM / ROWS, LB = BN / ROWS;
  static_assert(BM % ROWS == 0 && BN % ROWS == 0, "tile must be a multiple of THREADS/BK");

  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int T = g.KH * g.KW;
  const int Nq = g.Cb * T;
  const long long Kp = (long long)g.N * g.Hs * g.Ws;
  const int m0 = blockIdx.y * BM, q0 = blockIdx.x * BN;
  const long long pbeg = (long long)blockIdx.z * g.chunk;
  const long long pend = pbeg + g.chunk < Kp ? pbeg + g.chunk : Kp;
  if (pbeg >= pend) return;

  const int pl = tid % BK, row = tid / BK;
  // pixel cursor of this thread (advances by BK per k-tile)
  long long pcur = pbeg + pl;
  int n = (int)(pcur / ((long long)g.Hs * g.Ws));
  int rem = (int)(pcur - (long long)n * g.Hs * g.Ws);
  int hs = rem / g.Ws, ws = rem - hs * g.Ws;

  int q_cb[LB], q_r[LB], q_s[LB];
  bool q_ok[LB];
#pragma unroll
  for (int j = 0; j < LB; ++j) {
    const int q = q0 + row + j * ROWS;
    q_ok[j] = q < Nq;
    const int qq = q_ok[j] ? q : 0;
    q_cb[j] = qq / T;
    const int t = qq - q_cb[j] * T;
    q_r[j] = t / g.KW - g.pad;
    q_s[j] = t % g.KW - g.pad;
  }

  float ra[LA], rb[LB];
  auto load_tile = [&]() {
    const bool ok = pcur < pend;
    const long long sbase = ((long long)n * g.Cs) * g.Hs * g.Ws + (long long)hs * g.Ws + ws;
#pragma unroll
    for (int j = 0; j < LA; ++j) {
      const int m = m0 + row + j * ROWS;
      ra[j] = (ok && m < g.Cs) ? __ldg(small + sbase + (long long)m * g.Hs * g.Ws) : 0.f;
    }
    const int hb0 = hs * g.stride, wb0 = ws * g.stride;
    const long long bbase = (long long)n * g.Cb * g.Hb * g.Wb;
#pragma unroll
    for (int j = 0; j < LB; ++j) {
      float v = 0.f;
      const int hb = hb0 + q_r[j], wb = wb0 + q_s[j];
      if (ok && q_ok[j] && (unsigned)hb < (unsigned)g.Hb && (unsigned)wb < (unsigned)g.Wb) {
        v = __ldg(big + bbase + ((long long)q_cb[j] * g.Hb + hb) * g.Wb + wb);
        if (g.big_square) v *= v;
      }
      rb[j] = v;
    }
    pcur += BK;
    ws += BK;
    while (ws >= g.Ws) {
      ws -= g.Ws;
      if (++hs >= g.Hs) {
        hs = 0;
        ++n;
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int j = 0; j < LA; ++j) As[buf][pl][row + j * ROWS] = ra[j];
#pragma unroll
    for (int j = 0; j < LB; ++j) Bs[buf][pl][row + j * ROWS] = rb[j];
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int tx = tid % TXN, ty = tid / TXN;
  const int nk = (int)((pend - pbeg + BK - 1) / BK);
  load_tile();
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float af[TM], bf[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) af[i] = As[buf][kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bf[j] = Bs[buf][kk][tx + j * TXN];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= g.Cs) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int q = q0 + tx + j * TXN;
      if (q >= Nq) continue;
      float* dst = dw + (long long)m * Nq + q;
      if (g.splits > 1) atomicAdd(dst, acc[i][j]);
      else *dst = acc[i][j];
    }
  }
}

// ---- host-side dispatch --------------------------------------------------------------------------------------------
static int launch_gather(const GatherGeom& g, const float* x, const float* w, const float* bias, const float* gdn_x,
                         float* norm_out, float* y, cudaStream_t s, const char* name) {
  const int st = g.transposed ? g.stride : 1;
  const int phases = st * st;
  const int Pa = (g.Ho + st - 1) / st, Pb = (g.Wo + st - 1) / st;  // largest phase
  const long long P = (long long)g.N * Pa * Pb;
  const int taps = g.transposed ? ((g.KH + st - 1) / st) * ((g.KW + st - 1) / st) : g.KH * g.KW;
  if (taps > kMaxTaps) {
    set_error("%s: %d filter taps exceed the engine limit %d", name, taps, kMaxTaps);
    return B200LIC_ERR_UNSUPPORTED;
  }
  if (g.Cout <= 16) {
    constexpr int BM = 8, BN = 256, BK = 8, TM = 8, TN = 1;
    dim3 grid((unsigned)((P + BN - 1) / BN), (g.Cout + BM - 1) / BM, phases);
    gather_gemm_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, s>>>(g, x, w, bias, gdn_x, norm_out, y);
  } else if (g.Cout <= 64 || P <= 64 * 148) {
    constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
    dim3 grid((unsigned)((P + BN - 1) / BN), (g.Cout + BM - 1) / BM, phases);
    gather_gemm_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, s>>>(g, x, w, bias, gdn_x, norm_out, y);
  } else {
    constexpr int BM = 128, BN = 128, BK = 8, TM = 8, TN = 8;
    dim3 grid((unsigned)((P + BN - 1) / BN), (g.Cout + BM - 1) / BM, phases);
    gather_gemm_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, s>>>(g, x, w, bias, gdn_x, norm_out, y);
  }
  B200_LAUNCH_CHECK(name);
  return B200LIC_OK;
}

static int launch_wgrad(WgradGeom g, const float* small, const float* big, float* dw, cudaStream_t s, const char* name) {
  const int Nq = g.Cb * g.KH * g.KW;
  const long long Kp = (long long)g.N * g.Hs * g.Ws;
  auto plan = [&](int BM, int BN, int BK) {
    const long long tiles = (long long)((g.Cs + BM - 1) / BM) * ((Nq + BN - 1) / BN);
    long long want = (2LL * num_sms() + tiles - 1) / tiles;          // aim for ~2 waves
    const long long max_splits = (Kp + 4LL * BK - 1) / (4LL * BK);   // >= 4 k-tiles per split
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    long long chunk = (Kp + want - 1) / want;
    chunk = (chunk + BK - 1) / BK * BK;
    g.splits = (int)((Kp + chunk - 1) / chunk);
    g.chunk = chunk;
  };
  if (g.Cs <= 64 || Nq <= 64) {
    constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
    plan(BM, BN, BK);
    if (g.splits > 1) {
      cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)g.Cs * Nq, s);
      if (e != cudaSuccess) {
        set_error("%s: memset failed: %s", name, cudaGetErrorString(e));
        return B200LIC_ERR_CUDA;
      }
    }
    dim3 grid((Nq + BN - 1) / BN, (g.Cs + BM - 1) / BM, g.splits);
    wgrad_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, s>>>(g, small, big, dw);
  } else {
    constexpr int BM = 128, BN = 128, BK = 8, TM = 8, TN = 8;
    plan(BM, BN, BK);
    if (g.splits > 1) {
      cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)g.Cs * Nq, s);
      if (e != cudaSuccess) {
        set_error("%s: memset failed: %s", name, cudaGetErrorString(e));
        return B200LIC_ERR_CUDA;
      }
    }
    dim3 grid((Nq + BN - 1) / BN, (g.Cs + BM - 1) / BM, g.splits);
    wgrad_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, s>>>(g, small, big, dw);
  }
  B200_LAUNCH_CHECK(name);
  return B200LIC_OK;
}

static int check_desc(const b200lic_conv_desc* d, const char* name, bool transposed) {
  if (!d) {
    set_error("%s: null descriptor", name);
    return B200LIC_ERR_ARG;
  }
  if (d->N <= 0 || d->Cin <= 0 || d->H <= 0 || d->W <= 0 || d->Cout <= 0 || d->Ho <= 0 || d->Wo <= 0 || d->KH <= 0 ||
      d->KW <= 0 || d->stride <= 0 || d->pad < 0) {
    set_error("%s: non-positive dimension in descriptor", name);
    return B200LIC_ERR_ARG;
  }
  if (!transposed) {
    const int ho = (d->H + 2 * d->pad - d->KH) / d->stride + 1, wo = (d->W + 2 * d->pad - d->KW) / d->stride + 1;
    if (ho != d->Ho || wo != d->Wo) {
      set_error("%s: output %dx%d inconsistent with input %dx%d k=%dx%d s=%d p=%d (expect %dx%d)", name, d->Ho, d->Wo,
                d->H, d->W, d->KH, d->KW, d->stride, d->pad, ho, wo);
      return B200LIC_ERR_ARG;
    }
  } else {
    const int hmin = (d->H - 1) * d->stride - 2 * d->pad + d->KH, wmin = (d->W - 1) * d->stride - 2 * d->pad + d->KW;
    if (d->Ho < hmin || d->Ho >= hmin + d->stride || d->Wo < wmin || d->Wo >= wmin + d->stride) {
      set_error("%s: output %dx%d not reachable by conv_transpose2d from %dx%d k=%dx%d s=%d p=%d", name, d->Ho, d->Wo,
                d->H, d->W, d->KH, d->KW, d->stride, d->pad);
      return B200LIC_ERR_ARG;
    }
  }
  return B200LIC_OK;
}

// Engine-internal entry points (also used by conv_tc.cu as the fallback for shapes tcgen05 does not take).
int simt_conv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, const float* gdn_x,
                  float* norm_out, float* y, cudaStream_t s) {
  GatherGeom g{d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 0,
               (long long)d->Cin * d->KH * d->KW, (long long)d->KH * d->KW, d->act, d->act_slope, d->in_square,
               d->gdn_mode, d->fixed_point};
  return launch_gather(g, x, w, bias, gdn_x, norm_out, y, s, "conv_fwd(simt)");
}

int simt_deconv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                    cudaStream_t s) {
  GatherGeom g{d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 1,
               (long long)d->KH * d->KW, (long long)d->Cout * d->KH * d->KW, d->act, d->act_slope, 0, 0,
               d->fixed_point};
  return launch_gather(g, x, w, bias, nullptr, nullptr, y, s, "deconv_fwd(simt)");
}

int simt_conv_dgrad(const b200lic_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t s) {
  // dx = conv_transpose2d(dy, w): gathered tensor dy [N,Cout,Ho,Wo], written tensor dx [N,Cin,H,W]
  GatherGeom g{d->N, d->Cout, d->Ho, d->Wo, d->Cin, d->H, d->W, d->KH, d->KW, d->stride, d->pad, 1,
               (long long)d->KH * d->KW, (long long)d->Cin * d->KH * d->KW, 0, 0.f, 0, 0, 0};
  return launch_gather(g, dy, w, nullptr, nullptr, nullptr, dx, s, "conv_dgrad(simt)");
}

int simt_deconv_dgrad(const b200lic_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t s) {
  // dx = conv2d(dy, w viewed as [out=Cin, in=Cout, KH, KW]): gathered dy [N,Cout,Ho,Wo], written dx [N,Cin,H,W]
  GatherGeom g{d->N, d->Cout, d->Ho, d->Wo, d->Cin, d->H, d->W, d->KH, d->KW, d->stride, d->pad, 0,
               (long long)d->Cout * d->KH * d->KW, (long long)d->KH * d->KW, 0, 0.f, 0, 0, 0};
  return launch_gather(g, dy, w, nullptr, nullptr, nullptr, dx, s, "deconv_dgrad(simt)");
}

int simt_conv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t s) {
  WgradGeom g{d->N, d->Cout, d->Ho, d->Wo, d->Cin, d->H, d->W, d->KH, d->KW, d->stride, d->pad, d->in_square, 1, 0};
  return launch_wgrad(g, dy, x, dw, s, "conv_wgrad(simt)");
}

int simt_deconv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t s) {
  WgradGeom g{d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 0, 1, 0};
  return launch_wgrad(g, x, dy, dw, s, "deconv_wgrad(simt)");
}

int conv_check_desc(const b200lic_conv_desc* d, const char* name, bool transposed) {
  return check_desc(d, name, transposed);
}

}  // namespace b200lic
