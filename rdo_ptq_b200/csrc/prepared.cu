// Prepared-operand path of the tensor-core conv engine: the calibration iteration and the evaluation forward hand the
// GEMM kernels operands that are ALREADY in tensor-core form, produced by the kernel that had the values in registers
// anyway, instead of running the two staging kernels (NHWC split, weight pack) in front of every GEMM:
//   * batch pick + QDrop mix (layer_opt.py:289-292)      -> split-bf16 NHWC activation operand   (stage_mix_kernel)
//   * weight quantiser (quantizer.py:175-177, :437-449)   -> packed bf16 hi/lo weight operand     (quant_pack_kernel)
//   * lp_loss value + gradient (quantizer.py:71-79)       -> split-bf16 NHWC dY operand of wgrad  (loss_stage_kernel)
//   * split-K sum of the weight gradient                  -> STE masks + regulariser + Adam       (wgrad_reduce_adam_kernel)
// Every activation byte the staging kernels re-read and re-wrote (8 B per element per GEMM operand) disappears, and a
// frozen layer's packed weights are built once and kept by the caller (evaluation re-quantised every weight on every
// forward, like the reference's QuantModule.forward does, quant_layer.py:113-115).
#include <cuda.h>
#include <cuda_bf16.h>
#include "prepared.cuh"

namespace b200lic {

int conv_check_desc(const b200lic_conv_desc* d, const char* name, bool transposed);
size_t smallc_conv_fwd_ws(const b200lic_conv_desc* d);
size_t smallc_deconv_fwd_ws(const b200lic_conv_desc* d);
int tc_pack_weights(const float* w, int Cout, int Cin, int KH, int KW, int stride, int pad, int transposed, int CoutPad,
                    int Cpad, int Tmax, int phases, long long s_co, long long s_ci, void* bh, void* bl, cudaStream_t s);
void tc2_limit_taps_once(int taps);
int tc2_launch_ex(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride, int pad,
                  int transposed, long long s_co, long long s_ci, int act, float slope, int in_square, int gdn_mode,
                  int fixed_point, const float* x, const float* w, const void* packed_w, const float* w_scale,
                  const float* bias, const float* gdn_x, float* norm_out, float* y, void* workspace,
                  size_t workspace_bytes, cudaStream_t s, const char* name);
bool tc2_x_slot(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride, int transposed,
                void* workspace, size_t workspace_bytes, void** hi, void** lo, int* cpad);
bool tc_wgrad_dy_slot(const b200lic_conv_desc* d, int transposed, void* workspace, size_t workspace_bytes, void** hi,
                      void** lo, int* cpad);
int tc_wgrad_prepared(const b200lic_conv_desc* d, int transposed, const void* x_hi, const void* x_lo, int x_pitch,
                      const float* dy, float* dw, void* ws, size_t ws_bytes, const WgTail* tail, cudaStream_t s);
int launch_quant_pack(const PackDst& g, const float* w, const float* alpha, const float* delta, const float* zp, int ch,
                      int inner, int n_levels, int soft, int mode, void* packed, float* w_q, cudaStream_t s);

__device__ __forceinline__ unsigned long long mix64p(unsigned long long x) {  // splitmix64 finaliser (elementwise.cu)
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
// the QDrop draw of gather_mix_kernel for flat element i of the [rows, row] batch: 16 bits of one hash word per 4 elements
__device__ __forceinline__ bool qdrop_keep_at(unsigned long long seed, size_t i, unsigned thresh) {
  const unsigned long long word = mix64p(seed ^ mix64p((unsigned long long)(i >> 2)));
  return (unsigned)((word >> (16u * (unsigned)(i & 3))) & 0xFFFFu) < thresh;
}
__device__ __forceinline__ void split2(float a, float b, __nv_bfloat162& hv, __nv_bfloat162& lv) {
  const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  hv.x = ah;
  hv.y = bh;
  lv.x = __float2bfloat16_rn(a - __bfloat162float(ah));
  lv.y = __float2bfloat16_rn(b - __bfloat162float(bh));
}

struct PickSched {                 // the device-schedule pick of gather_mix_sched / lp_loss_fwd_bwd_sched
  const long long* idx_table;      // [table_rows][rows] or nullptr (identity)
  int table_rows, units, unit;
  const b200lic_calib_sched* sched;
};

// out row b = keep ? q[idx[b]] : fp[idx[b]] (fp32 NCHW, [rows, C, HW]) -> xh / xl [rows, HW, cpad] bf16 (x, or x*x with
// `square`), and optionally the mixed fp32 row itself (GDN's epilogue needs x next to the x*x operand).
// One CTA transposes a 64-channel x 64-pixel tile through shared memory (reads: 256 B per warp from each source; writes:
// one 128 B NHWC row per warp and slice).  Same picks and the same QDrop draws as gather_mix_kernel.
__global__ void __launch_bounds__(256)
    stage_mix_kernel(const float* __restrict__ q, const float* __restrict__ fp, PickSched ps, int C, int HW, int cpad,
                     float prob, unsigned long long seed, int square, __nv_bfloat16* __restrict__ xh,
                     __nv_bfloat16* __restrict__ xl, float* __restrict__ out) {
  pdl_wait();                 // programmatic dependent launch (common.cuh): predecessor complete, memory visible
  __shared__ float t[64][65];
  const long long* idx = nullptr;
  if (ps.sched != nullptr) {
    const unsigned long long k =
        (unsigned long long)(__ldg(&ps.sched->step) - 1) * (unsigned long long)ps.units + (unsigned long long)ps.unit;
    idx = ps.idx_table ? ps.idx_table + (size_t)(k % (unsigned long long)ps.table_rows) * gridDim.z : nullptr;
    seed = (seed + k) & 0xFFFFFFFFFFFFull;
  }
  const unsigned thresh = prob >= 1.f ? 65536u : (unsigned)(fminf(fmaxf(prob, 0.f), 1.f) * 65536.f);
  const bool all_q = prob >= 1.f;
  const int cblk = (cpad + 63) >> 6;
  const int b = blockIdx.z, c0 = (int)(blockIdx.x % cblk) * 64, p0 = (int)(blockIdx.x / cblk) * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t row = (size_t)C * HW;
  const size_t src_row = (idx ? (size_t)idx[b] : (size_t)b) * row, dst_row = (size_t)b * row;
  const bool vec2 = (HW & 1) == 0;
  if (vec2) {
    // all 16 loads of the thread (8 channel rows x {q, fp}) are issued before the first use; one hash word serves both
    // pixels of a pair (element index even => same group of four)
    const int p = p0 + 2 * lane;
    float2 qv[8], fv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + warp + 8 * j;
      qv[j] = (c < C && p < HW) ? __ldg(reinterpret_cast<const float2*>(q + src_row + (size_t)c * HW + p)) : make_float2(0.f, 0.f);
    }
    if (!all_q) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + warp + 8 * j;
        fv[j] = (c < C && p < HW) ? __ldg(reinterpret_cast<const float2*>(fp + src_row + (size_t)c * HW + p)) : make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cl = warp + 8 * j, c = c0 + cl;
      float v0 = qv[j].x, v1 = qv[j].y;
      if (c < C && p < HW) {
        const size_t e = (size_t)c * HW + p;
        if (!all_q) {
          const size_t i = dst_row + e;
          const unsigned long long word = mix64p(seed ^ mix64p((unsigned long long)(i >> 2)));
          const unsigned sh = 16u * (unsigned)(i & 3);
          if (!((unsigned)((word >> sh) & 0xFFFFu) < thresh)) v0 = fv[j].x;
          if (!((unsigned)((word >> (sh + 16u)) & 0xFFFFu) < thresh)) v1 = fv[j].y;
        }
        if (out != nullptr) *reinterpret_cast<float2*>(out + dst_row + e) = make_float2(v0, v1);
      }
      if (square) {
        v0 *= v0;
        v1 *= v1;
      }
      t[cl][2 * lane] = v0;
      t[cl][2 * lane + 1] = v1;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cl = warp + 8 * j, c = c0 + cl, p = p0 + 2 * lane;
      float v0 = 0.f, v1 = 0.f;
      if (c < C && p < HW) {
        const size_t e = (size_t)c * HW + p;
        const bool has1 = p + 1 < HW;
        const float q0 = __ldg(q + src_row + e), q1 = has1 ? __ldg(q + src_row + e + 1) : 0.f;
        float f0 = 0.f, f1 = 0.f;
        if (!all_q) {
          f0 = __ldg(fp + src_row + e);
          if (has1) f1 = __ldg(fp + src_row + e + 1);
        }
        v0 = (all_q || qdrop_keep_at(seed, dst_row + e, thresh)) ? q0 : f0;
        if (has1) v1 = (all_q || qdrop_keep_at(seed, dst_row + e + 1, thresh)) ? q1 : f1;
        if (out != nullptr) {
          out[dst_row + e] = v0;
          if (has1) out[dst_row + e + 1] = v1;
        }
      }
      if (square) {
        v0 *= v0;
        v1 *= v1;
      }
      t[cl][2 * lane] = v0;
      t[cl][2 * lane + 1] = v1;
    }
  }
  __syncthreads();
  if (c0 + 2 * lane >= cpad) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int pl = warp + 8 * j, p = p0 + pl;
    if (p < HW) {
      __nv_bfloat162 hv, lv;
      split2(t[2 * lane][pl], t[2 * lane + 1][pl], hv, lv);
      const size_t o = ((size_t)b * HW + p) * cpad + c0 + 2 * lane;
      *reinterpret_cast<__nv_bfloat162*>(xh + o) = hv;
      *reinterpret_cast<__nv_bfloat162*>(xl + o) = lv;
    }
  }
}

// loss += scale * sum |pred - tgt|^p over the batch, with the gradient grad_scale * p |d|^(p-1) sign(d) leaving as the
// split-bf16 NHWC operand [rows, HW, cpad] of the weight-gradient GEMM (and optionally as fp32 NCHW for a dgrad).
// tgt row b = tgt_cache[idx[b]] (the pick of the device schedule) or tgt_cache[b].
// GDN instantiation: `pred` is a GDN unit's output -- given, or (pred == nullptr) recomputed as x * norm^-+1/2 exactly as
// the conv engine's GDN epilogue computes it -- and the gradient is carried on to the norm accumulator,
// d_norm = dy * x * d(norm^-+1/2)/dnorm (b200lic_gdn_bwd_prep's arithmetic), the dY operand of gamma's weight gradient.
// Loads are issued four channel rows at a time before their first use (registers: 4 CTAs of 256 threads per SM).
template <bool GDN>
__global__ void __launch_bounds__(256, 4)
    loss_stage_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, PickSched ps, int C, int HW, int cpad,
                      float p, float scale, float grad_scale, int act, float slope, const float* __restrict__ gdn_x,
                      const float* __restrict__ gdn_norm, int gdn_inverse, float* __restrict__ loss,
                      __nv_bfloat16* __restrict__ gh, __nv_bfloat16* __restrict__ gl, float* __restrict__ d_pred) {
  pdl_wait();                 // programmatic dependent launch (common.cuh): predecessor complete, memory visible
  __shared__ float t[64][65];
  __shared__ float red[32];
  const long long* idx = nullptr;
  if (ps.sched != nullptr && ps.idx_table != nullptr) {
    const unsigned long long k =
        (unsigned long long)(__ldg(&ps.sched->step) - 1) * (unsigned long long)ps.units + (unsigned long long)ps.unit;
    idx = ps.idx_table + (size_t)(k % (unsigned long long)ps.table_rows) * gridDim.z;
  }
  const int cblk = (cpad + 63) >> 6;
  const int b = blockIdx.z, c0 = (int)(blockIdx.x % cblk) * 64, p0 = (int)(blockIdx.x / cblk) * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t row = (size_t)C * HW;
  const size_t tgt_row = (idx ? (size_t)idx[b] : (size_t)b) * row, pred_row = (size_t)b * row;
  const bool vec2 = (HW & 1) == 0;
  const bool p2 = (p == 2.f);
  float acc = 0.f;
  // `pred` is the output of the layer's fused activation: its derivative (b200lic_act_bwd on the activation OUTPUT) is
  // applied here, so the gradient that leaves is the one at the pre-activation accumulator, the wgrad's dY
  const float neg = act == B200LIC_ACT_RELU ? 0.f : (act == B200LIC_ACT_LEAKY_RELU ? slope : 1.f);
  auto one = [&](float a, float bb) -> float {
    const float d = a - bb;
    float g;
    if (p2) {
      acc += d * d;
      g = grad_scale * 2.f * d;
    } else {
      const float ad = fabsf(d);
      const float pw = powf(ad, p - 1.f);
      acc += pw * ad;
      g = grad_scale * p * pw * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    }
    return a > 0.f ? g : g * neg;
  };
  auto gdn_y = [&](float x, float nrm) -> float { return gdn_inverse ? x * sqrtf(nrm) : x * rsqrtf(nrm); };
  auto dnorm = [&](float g, float x, float nrm) -> float {
    const float r = rsqrtf(nrm);
    return gdn_inverse ? g * x * 0.5f * r : g * x * (-0.5f) * r * r * r;
  };
  if (vec2) {
    const int px = p0 + 2 * lane;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float2 av[4], bv[4], xv[4], nv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + warp + 8 * (4 * half + j);
        const bool ok = c < C && px < HW;
        const size_t e = pred_row + (size_t)c * HW + px;
        bv[j] = ok ? __ldg(reinterpret_cast<const float2*>(tgt + tgt_row + (size_t)c * HW + px)) : make_float2(0.f, 0.f);
        if (GDN) {
          xv[j] = ok ? __ldg(reinterpret_cast<const float2*>(gdn_x + e)) : make_float2(0.f, 0.f);
          nv[j] = ok ? __ldg(reinterpret_cast<const float2*>(gdn_norm + e)) : make_float2(1.f, 1.f);
        }
        if (!GDN || pred != nullptr)
          av[j] = ok ? __ldg(reinterpret_cast<const float2*>(pred + e)) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cl = warp + 8 * (4 * half + j), c = c0 + cl;
        float g0 = 0.f, g1 = 0.f;
        if (c < C && px < HW) {
          float a0, a1;
          if (GDN && pred == nullptr) {
            a0 = gdn_y(xv[j].x, nv[j].x);
            a1 = gdn_y(xv[j].y, nv[j].y);
          } else {
            a0 = av[j].x;
            a1 = av[j].y;
          }
          g0 = one(a0, bv[j].x);
          g1 = one(a1, bv[j].y);
          if (GDN) {
            g0 = dnorm(g0, xv[j].x, nv[j].x);
            g1 = dnorm(g1, xv[j].y, nv[j].y);
          }
          if (d_pred != nullptr)
            *reinterpret_cast<float2*>(d_pred + pred_row + (size_t)c * HW + px) = make_float2(g0, g1);
        }
        t[cl][2 * lane] = g0;
        t[cl][2 * lane + 1] = g1;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cl = warp + 8 * j, c = c0 + cl, px = p0 + 2 * lane;
      float g0 = 0.f, g1 = 0.f;
      if (c < C && px < HW) {
        const size_t e = (size_t)c * HW + px;
        const bool has1 = px + 1 < HW;
        const float a0 = (GDN && pred == nullptr) ? gdn_y(__ldg(gdn_x + pred_row + e), __ldg(gdn_norm + pred_row + e))
                                                  : __ldg(pred + pred_row + e);
        g0 = one(a0, __ldg(tgt + tgt_row + e));
        if (has1) {
          const float a1 = (GDN && pred == nullptr)
                               ? gdn_y(__ldg(gdn_x + pred_row + e + 1), __ldg(gdn_norm + pred_row + e + 1))
                               : __ldg(pred + pred_row + e + 1);
          g1 = one(a1, __ldg(tgt + tgt_row + e + 1));
        }
        if (GDN) {
          g0 = dnorm(g0, __ldg(gdn_x + pred_row + e), __ldg(gdn_norm + pred_row + e));
          if (has1) g1 = dnorm(g1, __ldg(gdn_x + pred_row + e + 1), __ldg(gdn_norm + pred_row + e + 1));
        }
        if (d_pred != nullptr) {
          d_pred[pred_row + e] = g0;
          if (has1) d_pred[pred_row + e + 1] = g1;
        }
      }
      t[cl][2 * lane] = g0;
      t[cl][2 * lane + 1] = g1;
    }
  }
  if (loss != nullptr) {                         // block_sum synchronises: the tile is complete afterwards
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(loss, scale * tot);
  }
  __syncthreads();
  if (c0 + 2 * lane >= cpad) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int pl = warp + 8 * j, px = p0 + pl;
    if (px < HW) {
      __nv_bfloat162 hv, lv;
      split2(t[2 * lane][pl], t[2 * lane + 1][pl], hv, lv);
      const size_t o = ((size_t)b * HW + px) * cpad + c0 + 2 * lane;
      *reinterpret_cast<__nv_bfloat162*>(gh + o) = hv;
      *reinterpret_cast<__nv_bfloat162*>(gl + o) = lv;
    }
  }
}

static bool fwd_is_folded(const b200lic_conv_desc* d, int op) {
  return (op == B200LIC_OP_DECONV_FWD ? smallc_deconv_fwd_ws(d) : smallc_conv_fwd_ws(d)) != 0;
}

static bool pack_geometry(const b200lic_conv_desc* d, int op, PackDst* g) {
  const int tr = op == B200LIC_OP_DECONV_FWD ? 1 : 0;
  if (fwd_is_folded(d, op)) return false;
  int Cpad, CoutPad, Tmax, phases;
  size_t bb;
  if (!tc2_weight_layout(d->Cin, d->Cout, d->KH, d->KW, d->stride, tr, &Cpad, &CoutPad, &Tmax, &phases, &bb)) return false;
  g->Cout = d->Cout; g->Cin = d->Cin; g->KH = d->KH; g->KW = d->KW; g->stride = d->stride; g->pad = d->pad;
  g->transposed = tr;
  g->CoutPad = CoutPad; g->Cpad = Cpad; g->Tmax = Tmax; g->phases = phases;
  g->s_co = tr ? (long long)d->KH * d->KW : (long long)d->Cin * d->KH * d->KW;
  g->s_ci = tr ? (long long)d->Cout * d->KH * d->KW : (long long)d->KH * d->KW;
  g->b_bytes = bb;
  return true;
}

}  // namespace b200lic

using namespace b200lic;

extern "C" {

size_t b200lic_conv_packed_weight_bytes(const b200lic_conv_desc* d, int op) {
  PackDst g;
  if (!d || (op != B200LIC_OP_CONV_FWD && op != B200LIC_OP_DECONV_FWD) || d->engine == B200LIC_ENGINE_SIMT) return 0;
  return pack_geometry(d, op, &g) ? 2 * g.b_bytes : 0;
}

int b200lic_conv_pack_weights(const b200lic_conv_desc* d, int op, const float* w, void* packed, size_t packed_bytes,
                              b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(d && w && packed, "conv_pack_weights: null pointer");
  B200_REQUIRE(op == B200LIC_OP_CONV_FWD || op == B200LIC_OP_DECONV_FWD, "conv_pack_weights: op must be a forward op");
  PackDst g;
  if (!pack_geometry(d, op, &g)) {
    set_error("conv_pack_weights: shape has no packed weight operand (folded-tap or ineligible layer)");
    return B200LIC_ERR_UNSUPPORTED;
  }
  B200_REQUIRE(packed_bytes >= 2 * g.b_bytes && (((uintptr_t)packed) & 127) == 0,
               "conv_pack_weights: needs %zu bytes, 128-byte aligned (got %zu)", 2 * g.b_bytes, packed_bytes);
  return tc_pack_weights(w, g.Cout, g.Cin, g.KH, g.KW, g.stride, g.pad, g.transposed, g.CoutPad, g.Cpad, g.Tmax, g.phases,
                         g.s_co, g.s_ci, packed, reinterpret_cast<uint8_t*>(packed) + g.b_bytes, as_stream(stream));
}

int b200lic_quant_pack_weights(const b200lic_conv_desc* d, int op, const float* w, const float* alpha, const float* delta,
                               const float* zero_point, int outer, int ch, int inner, int n_levels, int soft,
                               int integer_mode, void* packed, size_t packed_bytes, float* w_q, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(d && w && delta && zero_point && packed, "quant_pack_weights: null pointer");
  B200_REQUIRE(op == B200LIC_OP_CONV_FWD || op == B200LIC_OP_DECONV_FWD, "quant_pack_weights: op must be a forward op");
  B200_REQUIRE(outer > 0 && ch > 0 && inner > 0 && n_levels >= 2, "quant_pack_weights: bad channel view");
  B200_REQUIRE((size_t)outer * ch * inner == (size_t)d->Cout * d->Cin * d->KH * d->KW,
               "quant_pack_weights: channel view does not cover the weight");
  B200_REQUIRE(!integer_mode || n_levels <= 256, "quant_pack_weights: integer mode needs n_levels <= 256 (bf16-exact)");
  PackDst g;
  if (!pack_geometry(d, op, &g)) {
    set_error("quant_pack_weights: shape has no packed weight operand (folded-tap or ineligible layer)");
    return B200LIC_ERR_UNSUPPORTED;
  }
  B200_REQUIRE(packed_bytes >= 2 * g.b_bytes && (((uintptr_t)packed) & 127) == 0,
               "quant_pack_weights: needs %zu bytes, 128-byte aligned (got %zu)", 2 * g.b_bytes, packed_bytes);
  return launch_quant_pack(g, w, alpha, delta, zero_point, ch, inner, n_levels, soft, integer_mode, packed, w_q,
                           as_stream(stream));
}

int b200lic_conv_fwd_packed(const b200lic_conv_desc* d, int op, const float* x, const void* packed_w, const float* w_scale,
                            const float* bias, const float* gdn_x, float* norm_out, float* y, void* workspace,
                            size_t workspace_bytes, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(op == B200LIC_OP_CONV_FWD || op == B200LIC_OP_DECONV_FWD, "conv_fwd_packed: op must be a forward op");
  const bool tr = op == B200LIC_OP_DECONV_FWD;
  int rc = conv_check_desc(d, "conv_fwd_packed", tr);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(packed_w && y, "conv_fwd_packed: null pointer");
  B200_REQUIRE(!d->gdn_mode || (gdn_x && !tr), "conv_fwd_packed: gdn_mode needs gdn_x and a plain convolution");
  if (d->engine == B200LIC_ENGINE_SIMT || fwd_is_folded(d, op)) {
    set_error("conv_fwd_packed: prepared operands belong to the generic tcgen05 path");
    return B200LIC_ERR_UNSUPPORTED;
  }
  if (tr)
    return tc2_launch_ex(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 1,
                         (long long)d->KH * d->KW, (long long)d->Cout * d->KH * d->KW, d->act, d->act_slope, 0, 0,
                         d->fixed_point, x, nullptr, packed_w, w_scale, bias, nullptr, nullptr, y, workspace,
                         workspace_bytes, as_stream(stream), "deconv_fwd_packed(tc)");
  if (d->k_taps > 0 && d->k_taps < d->KH * d->KW && !d->gdn_mode) tc2_limit_taps_once(d->k_taps);
  return tc2_launch_ex(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->pad, 0,
                       (long long)d->Cin * d->KH * d->KW, (long long)d->KH * d->KW, d->act, d->act_slope, d->in_square,
                       d->gdn_mode, d->fixed_point, x, nullptr, packed_w, w_scale, bias, gdn_x, norm_out, y, workspace,
                       workspace_bytes, as_stream(stream), "conv_fwd_packed(tc)");
}

int b200lic_conv_x_slot(const b200lic_conv_desc* d, int op, void* workspace, size_t workspace_bytes, void** x_hi,
                        void** x_lo, int* cpad) {
  B200_REQUIRE(d && x_hi && x_lo && cpad, "conv_x_slot: null pointer");
  B200_REQUIRE(op == B200LIC_OP_CONV_FWD || op == B200LIC_OP_DECONV_FWD, "conv_x_slot: op must be a forward op");
  *x_hi = *x_lo = nullptr;
  *cpad = 0;
  if (d->engine == B200LIC_ENGINE_SIMT || fwd_is_folded(d, op)) return B200LIC_OK;
  tc2_x_slot(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, op == B200LIC_OP_DECONV_FWD,
             workspace, workspace_bytes, x_hi, x_lo, cpad);
  return B200LIC_OK;
}

int b200lic_conv_dy_slot(const b200lic_conv_desc* d, int op, void* workspace, size_t workspace_bytes, void** dy_hi,
                         void** dy_lo, int* cpad) {
  B200_REQUIRE(d && dy_hi && dy_lo && cpad, "conv_dy_slot: null pointer");
  B200_REQUIRE(op == B200LIC_OP_CONV_WGRAD || op == B200LIC_OP_DECONV_WGRAD, "conv_dy_slot: op must be a wgrad op");
  *dy_hi = *dy_lo = nullptr;
  *cpad = 0;
  if (d->engine == B200LIC_ENGINE_SIMT) return B200LIC_OK;
  tc_wgrad_dy_slot(d, op == B200LIC_OP_DECONV_WGRAD, workspace, workspace_bytes, dy_hi, dy_lo, cpad);
  return B200LIC_OK;
}

int b200lic_stage_mix_sched(const float* q, const float* fp, const long long* idx_table, int table_rows, int rows, int C,
                            int HW, float prob, unsigned long long seed_base, int units, int unit,
                            const b200lic_calib_sched* sched, int square, void* x_hi, void* x_lo, int cpad, float* out,
                            b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE(q && fp && x_hi && x_lo, "stage_mix_sched: null pointer");
  B200_REQUIRE(rows >= 1 && rows <= 65535 && C >= 1 && HW >= 1 && cpad >= C && (cpad % 32) == 0,
               "stage_mix_sched: bad shape (rows=%d C=%d HW=%d cpad=%d)", rows, C, HW, cpad);
  B200_REQUIRE(!sched || (units >= 1 && unit >= 0 && unit < units), "stage_mix_sched: unit %d outside [0,%d)", unit, units);
  B200_REQUIRE(!idx_table || (sched && table_rows >= 1), "stage_mix_sched: an index table needs the schedule");
  B200_REQUIRE(((((uintptr_t)q) | ((uintptr_t)fp) | ((uintptr_t)out)) & 7) == 0, "stage_mix_sched: 8-byte alignment");
  PickSched ps{idx_table, table_rows, units, unit, sched};
  dim3 grid((unsigned)(((cpad + 63) / 64) * ((HW + 63) / 64)), 1, (unsigned)rows);
  launch_pdl(stage_mix_kernel, grid, dim3(256), 0, as_stream(stream), q, fp, ps, C, HW, cpad, prob, seed_base & 0xFFFFFFFFFFFFull, square,
                                                         reinterpret_cast<__nv_bfloat16*>(x_hi),
                                                         reinterpret_cast<__nv_bfloat16*>(x_lo), out);
  B200_LAUNCH_CHECK("stage_mix_kernel");
  return B200LIC_OK;
}

int b200lic_lp_loss_stage_sched(const float* pred, const float* tgt_cache, const long long* idx_table, int table_rows,
                                int rows, int C, int HW, int units, int unit, const b200lic_calib_sched* sched, float p,
                                float scale, float grad_scale, int act, float act_slope, const float* gdn_x,
                                const float* gdn_norm, int gdn_inverse, float* loss, void* dy_hi, void* dy_lo, int cpad,
                                float* d_pred, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  B200_REQUIRE((pred || gdn_x) && tgt_cache && dy_hi && dy_lo, "lp_loss_stage_sched: null pointer");
  B200_REQUIRE(p >= 1.f, "lp_loss_stage_sched: p=%f < 1", p);
  B200_REQUIRE(rows >= 1 && rows <= 65535 && C >= 1 && HW >= 1 && cpad >= C && (cpad % 32) == 0,
               "lp_loss_stage_sched: bad shape (rows=%d C=%d HW=%d cpad=%d)", rows, C, HW, cpad);
  B200_REQUIRE(!idx_table || (sched && table_rows >= 1 && units >= 1 && unit >= 0 && unit < units),
               "lp_loss_stage_sched: bad pick arguments");
  B200_REQUIRE(((((uintptr_t)pred) | ((uintptr_t)tgt_cache) | ((uintptr_t)d_pred) | ((uintptr_t)gdn_x) |
                 ((uintptr_t)gdn_norm)) & 7) == 0, "lp_loss_stage_sched: 8-byte alignment");
  B200_REQUIRE((gdn_x == nullptr) == (gdn_norm == nullptr), "lp_loss_stage_sched: gdn_x and gdn_norm go together");
  PickSched ps{idx_table, table_rows, units, unit, sched};
  dim3 grid((unsigned)(((cpad + 63) / 64) * ((HW + 63) / 64)), 1, (unsigned)rows);
  if (gdn_x != nullptr)
    launch_pdl(loss_stage_kernel<true>, grid, dim3(256), 0, as_stream(stream), pred, tgt_cache, ps, C, HW, cpad, p, scale, grad_scale,
                                                                 act, act_slope, gdn_x, gdn_norm, gdn_inverse, loss,
                                                                 reinterpret_cast<__nv_bfloat16*>(dy_hi),
                                                                 reinterpret_cast<__nv_bfloat16*>(dy_lo), d_pred);
  else
    launch_pdl(loss_stage_kernel<false>, grid, dim3(256), 0, as_stream(stream), pred, tgt_cache, ps, C, HW, cpad, p, scale, grad_scale,
                                                                  act, act_slope, nullptr, nullptr, 0, loss,
                                                                  reinterpret_cast<__nv_bfloat16*>(dy_hi),
                                                                  reinterpret_cast<__nv_bfloat16*>(dy_lo), d_pred);
  B200_LAUNCH_CHECK("loss_stage_kernel");
  return B200LIC_OK;
}

int b200lic_conv_wgrad_adam_sched(const b200lic_conv_desc* d, int transposed, const void* x_hi, const void* x_lo,
                                  int x_cpad, const float* dy, void* workspace, size_t workspace_bytes, const float* w, float* alpha,
                                  const float* delta, const float* zero_point, float* exp_avg, float* exp_avg_sq,
                                  int outer, int ch, int inner, int n_levels, const b200lic_calib_sched* sched, float beta1,
                                  float beta2, float eps, float grad_scale, float reg_weight, float* reg_loss,
                                  float* dw_out, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "conv_wgrad_adam_sched", transposed != 0);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x_hi && x_lo && w && alpha && delta && zero_point && exp_avg && exp_avg_sq && sched,
               "conv_wgrad_adam_sched: null pointer");
  B200_REQUIRE(d->engine != B200LIC_ENGINE_SIMT, "conv_wgrad_adam_sched: staged operands belong to the tensor-core engine");
  B200_REQUIRE((size_t)outer * ch * inner == (size_t)d->Cout * d->Cin * d->KH * d->KW,
               "conv_wgrad_adam_sched: channel view does not cover the weight");
  WgTail t{w, alpha, delta, zero_point, exp_avg, exp_avg_sq, outer, ch, inner, (float)(n_levels - 1), sched,
           beta1, beta2, eps, grad_scale, reg_weight, reg_loss, dw_out};
  return tc_wgrad_prepared(d, transposed, x_hi, x_lo, x_cpad, dy, nullptr, workspace, workspace_bytes, &t,
                           as_stream(stream));
}

/* b200lic_conv_wgrad_staged with dy == NULL: dy was staged by b200lic_lp_loss_stage_sched into the slot
 * b200lic_conv_dy_slot reports. */
int b200lic_conv_wgrad_prepared(const b200lic_conv_desc* d, int transposed, const void* x_hi, const void* x_lo,
                                int x_cpad, const float* dy, float* dw, void* workspace, size_t workspace_bytes,
                                b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "conv_wgrad_prepared", transposed != 0);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x_hi && x_lo && dw, "conv_wgrad_prepared: null pointer");
  B200_REQUIRE(d->engine != B200LIC_ENGINE_SIMT, "conv_wgrad_prepared: staged operands belong to the tensor-core engine");
  return tc_wgrad_prepared(d, transposed, x_hi, x_lo, x_cpad, dy, dw, workspace, workspace_bytes, nullptr,
                           as_stream(stream));
}

}  // extern "C"
