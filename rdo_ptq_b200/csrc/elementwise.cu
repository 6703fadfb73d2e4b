// K11 (loss reductions) and the small fused elementwise helpers around the conv kernels.
// All HBM-bound: float4 where alignment allows, grid a multiple of the SM count, one atomic per CTA.
//
// Reference lines restated: quantizer.py:71-79 (lp_loss), losses/losses.py:20-28, test_datasets.py:21-33,98,
// quant_block.py:219-328 (residual add / LeakyReLU placement), layer_opt.py:291-292 (QDrop mix),
// compressai GDN reparametrisation (via quant_layer.py:142-154).
#include "common.cuh"

namespace b200lic {

// Generic float4-vectorised elementwise launcher: F::apply(a,b,c) -> out, any of b/c may be unused.
template <typename F>
__global__ void __launch_bounds__(256) ew_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                  const float* __restrict__ c, float* __restrict__ out, size_t n, F f) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)out) & 15) == 0;
  if (vec) {
    const size_t n4 = n >> 2;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = tid; i < n4; i += stride) {
      const float4 av = __ldg(reinterpret_cast<const float4*>(a) + i);
      const float4 bv = b ? __ldg(reinterpret_cast<const float4*>(b) + i) : zero;
      const float4 cv = c ? __ldg(reinterpret_cast<const float4*>(c) + i) : zero;
      float4 o;
      o.x = f(av.x, bv.x, cv.x, 4 * i);
      o.y = f(av.y, bv.y, cv.y, 4 * i + 1);
      o.z = f(av.z, bv.z, cv.z, 4 * i + 2);
      o.w = f(av.w, bv.w, cv.w, 4 * i + 3);
      reinterpret_cast<float4*>(out)[i] = o;
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += stride) out[i] = f(a[i], b ? b[i] : 0.f, c ? c[i] : 0.f, i);
  } else {
    for (size_t i = tid; i < n; i += stride) out[i] = f(a[i], b ? b[i] : 0.f, c ? c[i] : 0.f, i);
  }
}

template <typename F>
static int launch_ew(const char* name, const float* a, const float* b, const float* c, float* out, size_t n, F f,
                     cudaStream_t s) {
  if (n == 0) return B200LIC_OK;
  ew_kernel<F><<<grid_for(n / 4 + 1, 256), 256, 0, s>>>(a, b, c, out, n, f);
  B200_LAUNCH_CHECK(name);
  return B200LIC_OK;
}

struct AddAct {
  int act;
  float slope;
  __device__ float operator()(float a, float b, float, size_t) const { return apply_act(a + b, act, slope); }
};
struct ActBwd {  // a = activation output y, b = d_out
  int act;
  float slope;
  __device__ float operator()(float y, float g, float, size_t) const {
    if (act == B200LIC_ACT_RELU) return y > 0.f ? g : 0.f;
    if (act == B200LIC_ACT_LEAKY_RELU) return y > 0.f ? g : g * slope;
    return g;
  }
};
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {  // splitmix64 finaliser
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
// Batch pick + QDrop mix in one pass: out[b, :] = keep ? q[idx[b], :] : fp[idx[b], :]
// keep(i) = mask[i] if a mask is given, else 16 bits of a counter-based hash (one splitmix64 word per 4 elements)
// compared with prob * 65536.  With a device schedule the batch-pick row and the seed follow sched->step.
struct GatherSched {
  const long long* idx_table;   // [table_rows][rows] or nullptr (identity)
  int table_rows, units, unit;
  const b200lic_calib_sched* sched;   // nullptr: use idx / seed as given
};

__device__ __forceinline__ bool qdrop_keep(unsigned long long word, unsigned lane4, unsigned thresh) {
  return (unsigned)((word >> (16u * lane4)) & 0xFFFFu) < thresh;
}

template <bool VEC>
__global__ void __launch_bounds__(256)
    gather_mix_kernel(const float* __restrict__ q, const float* __restrict__ fp, const long long* __restrict__ idx,
                      size_t rows, size_t row, float prob, unsigned long long seed, const uint8_t* __restrict__ mask,
                      GatherSched gs, float* __restrict__ out) {
  if (gs.sched != nullptr) {
    const unsigned long long k =
        (unsigned long long)(__ldg(&gs.sched->step) - 1) * (unsigned long long)gs.units + (unsigned long long)gs.unit;
    idx = gs.idx_table ? gs.idx_table + (size_t)(k % (unsigned long long)gs.table_rows) * rows : nullptr;
    seed = (seed + k) & 0xFFFFFFFFFFFFull;
  }
  const unsigned thresh = prob >= 1.f ? 65536u : (unsigned)(fminf(fmaxf(prob, 0.f), 1.f) * 65536.f);
  const bool all_q = (mask == nullptr && prob >= 1.f);
  const size_t n = rows * row;
  const size_t stride = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    for (size_t i4 = tid; i4 < (n >> 2); i4 += stride) {
      const size_t i = i4 << 2;
      const size_t b = i / row, e = i - b * row;
      const size_t src = (idx ? (size_t)idx[b] : b) * row + e;
      const float4 qv = __ldg(reinterpret_cast<const float4*>(q + src));
      float4 o = qv;
      if (!all_q) {
        const float4 fv = __ldg(reinterpret_cast<const float4*>(fp + src));
        bool k0, k1, k2, k3;
        if (mask) {
          const uchar4 mv = *reinterpret_cast<const uchar4*>(mask + i);
          k0 = mv.x != 0; k1 = mv.y != 0; k2 = mv.z != 0; k3 = mv.w != 0;
        } else {
          const unsigned long long word = mix64(seed ^ mix64((unsigned long long)i4));
          k0 = qdrop_keep(word, 0, thresh); k1 = qdrop_keep(word, 1, thresh);
          k2 = qdrop_keep(word, 2, thresh); k3 = qdrop_keep(word, 3, thresh);
        }
        o.x = k0 ? qv.x : fv.x; o.y = k1 ? qv.y : fv.y; o.z = k2 ? qv.z : fv.z; o.w = k3 ? qv.w : fv.w;
      }
      reinterpret_cast<float4*>(out)[i4] = o;
    }
  } else {
    for (size_t i = tid; i < n; i += stride) {
      const size_t b = i / row, e = i - b * row;
      const size_t src = (idx ? (size_t)idx[b] : b) * row + e;
      bool keep;
      if (mask) keep = mask[i] != 0;
      else if (all_q) keep = true;
      else keep = qdrop_keep(mix64(seed ^ mix64((unsigned long long)(i >> 2))), (unsigned)(i & 3), thresh);
      out[i] = keep ? __ldg(q + src) : __ldg(fp + src);
    }
  }
}

static int launch_gather_mix(const char* name, const float* q, const float* fp, const long long* idx, size_t rows,
                             size_t row_elems, float prob, unsigned long long seed, const uint8_t* mask,
                             GatherSched gs, float* out, cudaStream_t s) {
  const size_t n = rows * row_elems;
  if (n == 0) return B200LIC_OK;
  const bool vec = (row_elems & 3) == 0 &&
                   (((uintptr_t)q | (uintptr_t)fp | (uintptr_t)out | (uintptr_t)mask) & 15) == 0;
  if (vec)
    gather_mix_kernel<true><<<grid_for(n / 4, 256, 8), 256, 0, s>>>(q, fp, idx, rows, row_elems, prob, seed, mask, gs, out);
  else
    gather_mix_kernel<false><<<grid_for(n, 256, 8), 256, 0, s>>>(q, fp, idx, rows, row_elems, prob, seed, mask, gs, out);
  B200_LAUNCH_CHECK(name);
  return B200LIC_OK;
}
struct AttnGate {
  __device__ float operator()(float a, float b, float c, size_t) const { return a * (1.f / (1.f + expf(-b))) + c; }
};
struct Abs {
  __device__ float operator()(float a, float, float, size_t) const { return fabsf(a); }
};
struct ReparamFwd {
  float bound, pedestal;
  __device__ float operator()(float p, float, float, size_t) const {
    const float l = fmaxf(p, bound);
    return l * l - pedestal;
  }
};
struct ReparamBwd {  // a = p, b = d_out
  float bound;
  __device__ float operator()(float p, float g, float, size_t) const {
    const float gl = g * 2.f * fmaxf(p, bound);
    return (p >= bound || gl < 0.f) ? gl : 0.f;
  }
};
struct GdnBwdDnorm {  // a = x, b = norm, c = dy
  int inverse;
  __device__ float operator()(float x, float nrm, float dy, size_t) const {
    const float r = rsqrtf(nrm);
    return inverse ? dy * x * 0.5f * r : dy * x * (-0.5f) * r * r * r;
  }
};
struct GdnBwdDirect {  // a = norm, b = dy
  int inverse;
  __device__ float operator()(float nrm, float dy, float, size_t) const {
    return inverse ? dy * sqrtf(nrm) : dy * rsqrtf(nrm);
  }
};
struct GdnBwdFinish {  // a = x, b = t, c = dx_direct
  __device__ float operator()(float x, float t, float d, size_t) const { return d + 2.f * x * t; }
};

// ---- reductions ----------------------------------------------------------------------------------------------
struct LossPick {               // target rows picked from a cache by the device schedule (nullptr table: plain tgt)
  const long long* idx_table;
  int table_rows, units, unit;
  size_t rows, row;
  const b200lic_calib_sched* sched;
};

__global__ void __launch_bounds__(256) lp_loss_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                       size_t n, float p, float scale, float grad_scale,
                                                       float* __restrict__ loss, float* __restrict__ d_pred,
                                                       LossPick pk) {
  __shared__ float red[32];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  const long long* idx = nullptr;
  if (pk.idx_table != nullptr) {
    const unsigned long long k =
        (unsigned long long)(__ldg(&pk.sched->step) - 1) * (unsigned long long)pk.units + (unsigned long long)pk.unit;
    idx = pk.idx_table + (size_t)(k % (unsigned long long)pk.table_rows) * pk.rows;
  }
  const bool p2 = (p == 2.f);
  auto one = [&](float a, float b, float& g) {
    const float d = a - b;
    if (p2) {
      acc += d * d;
      g = grad_scale * 2.f * d;
    } else {
      const float ad = fabsf(d);
      const float pw = powf(ad, p - 1.f);
      acc += pw * ad;
      g = grad_scale * p * pw * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    }
  };
  if (idx != nullptr) {           // picked target rows; row % 4 == 0 and 16-byte alignment are checked by the host
    const size_t n4 = n >> 2;
    auto tgt4 = [&](size_t i) {
      const size_t e0 = i << 2, b_ = e0 / pk.row, e = e0 - b_ * pk.row;
      return __ldg(reinterpret_cast<const float4*>(tgt + (size_t)idx[b_] * pk.row + e));
    };
    size_t i = tid;
    for (; i + stride < n4; i += 2 * stride) {        // two independent float4 pairs in flight per thread
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(pred) + i), b0 = tgt4(i);
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(pred) + i + stride), b1 = tgt4(i + stride);
      float4 g0, g1;
      one(a0.x, b0.x, g0.x);
      one(a0.y, b0.y, g0.y);
      one(a0.z, b0.z, g0.z);
      one(a0.w, b0.w, g0.w);
      one(a1.x, b1.x, g1.x);
      one(a1.y, b1.y, g1.y);
      one(a1.z, b1.z, g1.z);
      one(a1.w, b1.w, g1.w);
      if (d_pred) {
        reinterpret_cast<float4*>(d_pred)[i] = g0;
        reinterpret_cast<float4*>(d_pred)[i + stride] = g1;
      }
    }
    for (; i < n4; i += stride) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(pred) + i), b = tgt4(i);
      float4 g;
      one(a.x, b.x, g.x);
      one(a.y, b.y, g.y);
      one(a.z, b.z, g.z);
      one(a.w, b.w, g.w);
      if (d_pred) reinterpret_cast<float4*>(d_pred)[i] = g;
    }
    if (loss) {
      const float tot = block_sum(acc, red);
      if (threadIdx.x == 0) atomicAdd(loss, scale * tot);
    }
    return;
  }
  const bool vec = (((uintptr_t)pred | (uintptr_t)tgt | (uintptr_t)d_pred) & 15) == 0;
  if (vec) {
    const size_t n4 = n >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(pred);
    const float4* t4 = reinterpret_cast<const float4*>(tgt);
    size_t i = tid;
    for (; i + stride < n4; i += 2 * stride) {        // two independent float4 pairs in flight per thread
      const float4 a0 = __ldg(p4 + i), b0 = __ldg(t4 + i), a1 = __ldg(p4 + i + stride), b1 = __ldg(t4 + i + stride);
      float4 g0, g1;
      one(a0.x, b0.x, g0.x);
      one(a0.y, b0.y, g0.y);
      one(a0.z, b0.z, g0.z);
      one(a0.w, b0.w, g0.w);
      one(a1.x, b1.x, g1.x);
      one(a1.y, b1.y, g1.y);
      one(a1.z, b1.z, g1.z);
      one(a1.w, b1.w, g1.w);
      if (d_pred) {
        reinterpret_cast<float4*>(d_pred)[i] = g0;
        reinterpret_cast<float4*>(d_pred)[i + stride] = g1;
      }
    }
    for (; i < n4; i += stride) {
      const float4 a = __ldg(p4 + i), b = __ldg(t4 + i);
      float4 g;
      one(a.x, b.x, g.x);
      one(a.y, b.y, g.y);
      one(a.z, b.z, g.z);
      one(a.w, b.w, g.w);
      if (d_pred) reinterpret_cast<float4*>(d_pred)[i] = g;
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += stride) {
      float g;
      one(pred[i], tgt[i], g);
      if (d_pred) d_pred[i] = g;
    }
  } else {
    for (size_t i = tid; i < n; i += stride) {
      float g;
      one(pred[i], tgt[i], g);
      if (d_pred) d_pred[i] = g;
    }
  }
  if (loss) {
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(loss, scale * tot);
  }
}

__global__ void __launch_bounds__(256) sq_err_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n,
                                                      float* __restrict__ out) {
  __shared__ float red[32];
  float s0 = 0.f, s1 = 0.f;
  auto one = [&](float av, float bv) {
    const float d0 = av - bv, d1 = fminf(fmaxf(av, 0.f), 1.f) - bv;
    s0 += d0 * d0;
    s1 += d1 * d1;
  };
  const size_t stride = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t done = 0;
  if ((((uintptr_t)a | (uintptr_t)b) & 15) == 0) {
    const size_t n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    size_t i = tid;
    for (; i + 3 * stride < n4; i += 4 * stride) {    // eight independent 16-byte loads in flight per thread
      float4 av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        av[u] = __ldg(a4 + i + u * stride);
        bv[u] = __ldg(b4 + i + u * stride);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        one(av[u].x, bv[u].x);
        one(av[u].y, bv[u].y);
        one(av[u].z, bv[u].z);
        one(av[u].w, bv[u].w);
      }
    }
    for (; i < n4; i += stride) {
      const float4 av = __ldg(a4 + i), bv = __ldg(b4 + i);
      one(av.x, bv.x);
      one(av.y, bv.y);
      one(av.z, bv.z);
      one(av.w, bv.w);
    }
    done = n4 << 2;
  }
  for (size_t i = done + tid; i < n; i += stride) one(__ldg(a + i), __ldg(b + i));
  const float t0 = block_sum(s0, red);
  const float t1 = block_sum(s1, red);
  if (threadIdx.x == 0) {
    atomicAdd(out, t0);
    atomicAdd(out + 1, t1);
  }
}

__global__ void __launch_bounds__(256) bits_sum_kernel(const float* __restrict__ lik, size_t n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t done = 0;
  if ((((uintptr_t)lik) & 15) == 0) {
    const size_t n4 = n >> 2;
    const float4* l4 = reinterpret_cast<const float4*>(lik);
    size_t i = tid;
    for (; i + 3 * stride < n4; i += 4 * stride) {    // four independent 16-byte loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(l4 + i + u * stride);
#pragma unroll
      for (int u = 0; u < 4; ++u) s -= log2f(v[u].x) + log2f(v[u].y) + log2f(v[u].z) + log2f(v[u].w);
    }
    for (; i < n4; i += stride) {
      const float4 v = __ldg(l4 + i);
      s -= log2f(v.x) + log2f(v.y) + log2f(v.z) + log2f(v.w);
    }
    done = n4 << 2;
  }
  for (size_t i = done + tid; i < n; i += stride) s -= log2f(__ldg(lik + i));
  const float t = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, t);
}

// ---- pixel shuffle (compressai subpel_conv3x3 tail; quant_layer.py:108-111 adds the LeakyReLU) -------------------
__global__ void __launch_bounds__(256) pixel_shuffle_kernel(const float* __restrict__ x, int C, int H, int W, int r,
                                                             size_t n, int act, float slope, int inverse,
                                                             float* __restrict__ out) {
  // forward: out[n,c,h*r+i,w*r+j] = act(x[n,c*r*r+i*r+j,h,w]); inverse: dx[...] = dy[...] (same index map)
  const int Ho = H * r, Wo = W * r;
  for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t)gridDim.x * blockDim.x) {
    const int wo = (int)(o % Wo);
    const int ho = (int)((o / Wo) % Ho);
    const int c = (int)((o / ((size_t)Wo * Ho)) % C);
    const int b = (int)(o / ((size_t)Wo * Ho * C));
    const int i = ho % r, j = wo % r, h = ho / r, w = wo / r;
    const size_t src = (((size_t)b * C * r * r + (size_t)c * r * r + i * r + j) * H + h) * W + w;
    if (inverse) out[src] = x[o];
    else out[o] = apply_act(__ldg(x + src), act, slope);
  }
}

}  // namespace b200lic

using namespace b200lic;

extern "C" {

int b200lic_lp_loss_fwd_bwd(const float* pred, const float* tgt, size_t n, float p, float scale, float grad_scale,
                            float* loss, float* d_pred, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(pred && tgt, "lp_loss_fwd_bwd: null pointer");
  B200_REQUIRE(p >= 1.f, "lp_loss_fwd_bwd: p=%f < 1", p);
  if (n == 0) return B200LIC_OK;
  lp_loss_kernel<<<grid_for(n / 8 + 1, 256, 4), 256, 0, as_stream(stream)>>>(pred, tgt, n, p, scale, grad_scale, loss,
                                                                             d_pred, LossPick{nullptr, 0, 0, 0, 0, 0, nullptr});
  B200_LAUNCH_CHECK("lp_loss_kernel");
  return B200LIC_OK;
}

int b200lic_lp_loss_fwd_bwd_sched(const float* pred, const float* tgt_cache, const long long* idx_table, int table_rows,
                                  size_t rows, size_t row_elems, int units, int unit, const b200lic_calib_sched* sched,
                                  float p, float scale, float grad_scale, float* loss, float* d_pred,
                                  b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(pred && tgt_cache && idx_table && sched, "lp_loss_fwd_bwd_sched: null pointer");
  B200_REQUIRE(p >= 1.f, "lp_loss_fwd_bwd_sched: p=%f < 1", p);
  B200_REQUIRE(table_rows >= 1 && units >= 1 && unit >= 0 && unit < units, "lp_loss_fwd_bwd_sched: bad pick arguments");
  B200_REQUIRE((row_elems & 3) == 0 && (((uintptr_t)pred | (uintptr_t)tgt_cache | (uintptr_t)d_pred) & 15) == 0,
               "lp_loss_fwd_bwd_sched: rows must be a multiple of 4 elements and 16-byte aligned");
  const size_t n = rows * row_elems;
  if (n == 0) return B200LIC_OK;
  lp_loss_kernel<<<grid_for(n / 8 + 1, 256, 4), 256, 0, as_stream(stream)>>>(
      pred, tgt_cache, n, p, scale, grad_scale, loss, d_pred,
      LossPick{idx_table, table_rows, units, unit, rows, row_elems, sched});
  B200_LAUNCH_CHECK("lp_loss_kernel(sched)");
  return B200LIC_OK;
}

int b200lic_sq_err_sum(const float* a, const float* b, size_t n, float* out, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(a && b && out, "sq_err_sum: null pointer");
  if (n == 0) return B200LIC_OK;
  sq_err_kernel<<<grid_for(n / 16 + 1, 256, 4), 256, 0, as_stream(stream)>>>(a, b, n, out);   // 60 registers: 4 CTAs/SM resident
  B200_LAUNCH_CHECK("sq_err_kernel");
  return B200LIC_OK;
}

int b200lic_bits_sum(const float* lik, size_t n, float* out, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(lik && out, "bits_sum: null pointer");
  if (n == 0) return B200LIC_OK;
  bits_sum_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, as_stream(stream)>>>(lik, n, out);
  B200_LAUNCH_CHECK("bits_sum_kernel");
  return B200LIC_OK;
}

int b200lic_gdn_reparam_fwd(const float* p, size_t n, float bound, float pedestal, float* out, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(p && out, "gdn_reparam_fwd: null pointer");
  return launch_ew("gdn_reparam_fwd", p, nullptr, nullptr, out, n, ReparamFwd{bound, pedestal}, as_stream(stream));
}

int b200lic_gdn_reparam_bwd(const float* p, const float* d_out, size_t n, float bound, float* d_p,
                            b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(p && d_out && d_p, "gdn_reparam_bwd: null pointer");
  return launch_ew("gdn_reparam_bwd", p, d_out, nullptr, d_p, n, ReparamBwd{bound}, as_stream(stream));
}

int b200lic_gdn_bwd_prep(const float* x, const float* norm, const float* dy, size_t n, int inverse, float* d_norm,
                         float* dx_direct, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && norm && dy && d_norm, "gdn_bwd_prep: null pointer");
  int rc = launch_ew("gdn_bwd_dnorm", x, norm, dy, d_norm, n, GdnBwdDnorm{inverse}, as_stream(stream));
  if (rc != B200LIC_OK || !dx_direct) return rc;
  return launch_ew("gdn_bwd_direct", norm, dy, nullptr, dx_direct, n, GdnBwdDirect{inverse}, as_stream(stream));
}

int b200lic_gdn_bwd_finish(const float* x, const float* t, const float* dx_direct, size_t n, float* dx,
                           b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && t && dx_direct && dx, "gdn_bwd_finish: null pointer");
  return launch_ew("gdn_bwd_finish", x, t, dx_direct, dx, n, GdnBwdFinish{}, as_stream(stream));
}

int b200lic_add_act(const float* a, const float* b, size_t n, int act, float slope, float* out, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(a && out, "add_act: null pointer");
  return launch_ew("add_act", a, b, nullptr, out, n, AddAct{act, slope}, as_stream(stream));
}

int b200lic_act_bwd(const float* y, const float* d_out, size_t n, int act, float slope, float* d_in,
                    b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(y && d_out && d_in, "act_bwd: null pointer");
  return launch_ew("act_bwd", y, d_out, nullptr, d_in, n, ActBwd{act, slope}, as_stream(stream));
}

int b200lic_gather_mix(const float* q, const float* fp, const long long* idx, size_t rows, size_t row_elems, float prob,
                       unsigned long long seed, const uint8_t* mask, float* out, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(q && fp && out, "gather_mix: null pointer");
  return launch_gather_mix("gather_mix_kernel", q, fp, idx, rows, row_elems, prob, seed, mask,
                           GatherSched{nullptr, 0, 0, 0, nullptr}, out, as_stream(stream));
}

int b200lic_gather_mix_sched(const float* q, const float* fp, const long long* idx_table, int table_rows, size_t rows,
                             size_t row_elems, float prob, unsigned long long seed_base, int units, int unit,
                             const b200lic_calib_sched* sched, float* out, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(q && fp && out && sched, "gather_mix_sched: null pointer");
  B200_REQUIRE(units >= 1 && unit >= 0 && unit < units, "gather_mix_sched: unit %d outside [0,%d)", unit, units);
  B200_REQUIRE(!idx_table || table_rows >= 1, "gather_mix_sched: empty index table");
  return launch_gather_mix("gather_mix_kernel(sched)", q, fp, nullptr, rows, row_elems, prob, seed_base, nullptr,
                           GatherSched{idx_table, table_rows, units, unit, sched}, out, as_stream(stream));
}

int b200lic_attn_gate(const float* a, const float* b, const float* c, size_t n, float* out, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(a && b && c && out, "attn_gate: null pointer");
  return launch_ew("attn_gate", a, b, c, out, n, AttnGate{}, as_stream(stream));
}

int b200lic_abs(const float* x, size_t n, float* out, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && out, "abs: null pointer");
  return launch_ew("abs", x, nullptr, nullptr, out, n, Abs{}, as_stream(stream));
}

int b200lic_pixel_shuffle(const float* x, int N, int C, int H, int W, int r, int act, float slope, float* out,
                          b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(x && out && N > 0 && C > 0 && H > 0 && W > 0 && r > 0, "pixel_shuffle: bad arguments");
  const size_t n = (size_t)N * C * H * W * r * r;
  pixel_shuffle_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, C, H, W, r, n, act, slope, 0, out);
  B200_LAUNCH_CHECK("pixel_shuffle_kernel");
  return B200LIC_OK;
}

int b200lic_pixel_unshuffle(const float* dy, int N, int C, int H, int W, int r, float* dx, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(dy && dx && N > 0 && C > 0 && H > 0 && W > 0 && r > 0, "pixel_unshuffle: bad arguments");
  const size_t n = (size_t)N * C * H * W * r * r;
  pixel_shuffle_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(dy, C, H, W, r, n, 0, 0.f, 1, dx);
  B200_LAUNCH_CHECK("pixel_shuffle_kernel(inverse)");
  return B200LIC_OK;
}

}  // extern "C"
