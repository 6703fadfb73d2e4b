// extern "C" convolution entry points (include/b200lic.h) and engine selection.
#include "common.cuh"

namespace b200lic {
int conv_check_desc(const b200lic_conv_desc* d, const char* name, bool transposed);
int simt_conv_fwd(const b200lic_conv_desc*, const float*, const float*, const float*, const float*, float*, float*,
                  cudaStream_t);
int simt_deconv_fwd(const b200lic_conv_desc*, const float*, const float*, const float*, float*, cudaStream_t);
int simt_conv_dgrad(const b200lic_conv_desc*, const float*, const float*, float*, cudaStream_t);
int simt_deconv_dgrad(const b200lic_conv_desc*, const float*, const float*, float*, cudaStream_t);
int simt_conv_wgrad(const b200lic_conv_desc*, const float*, const float*, float*, cudaStream_t);
int simt_deconv_wgrad(const b200lic_conv_desc*, const float*, const float*, float*, cudaStream_t);
// tensor-core engine (conv_tc.cu): B200LIC_ERR_UNSUPPORTED when the shape / workspace does not qualify
size_t tc_workspace_bytes(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride,
                          int transposed);
int tc_conv_fwd(const b200lic_conv_desc*, const float*, const float*, const float*, const float*, float*, float*, void*,
                size_t, cudaStream_t);
int tc_deconv_fwd(const b200lic_conv_desc*, const float*, const float*, const float*, float*, void*, size_t,
                  cudaStream_t);
int tc_conv_dgrad(const b200lic_conv_desc*, const float*, const float*, float*, void*, size_t, cudaStream_t);
int tc_deconv_dgrad(const b200lic_conv_desc*, const float*, const float*, float*, void*, size_t, cudaStream_t);
int tc_conv_wgrad(const b200lic_conv_desc*, const float*, const float*, float*, void*, size_t, cudaStream_t);
int tc_deconv_wgrad(const b200lic_conv_desc*, const float*, const float*, float*, void*, size_t, cudaStream_t);
bool tc_staged_view(const b200lic_conv_desc* d, int transposed, void* fwd_ws, size_t ws_bytes, void** hi, void** lo);
int tc_conv_wgrad_pre(const b200lic_conv_desc*, const void*, const void*, const float*, float*, void*, size_t, cudaStream_t);
int tc_deconv_wgrad_pre(const b200lic_conv_desc*, const void*, const void*, const float*, float*, void*, size_t,
                        cudaStream_t);
int tc_conv_fwd_wq(const b200lic_conv_desc*, const float*, const float*, const float*, const float*, float*, void*, size_t,
                   cudaStream_t);
int tc_deconv_fwd_wq(const b200lic_conv_desc*, const float*, const float*, const float*, const float*, float*, void*,
                     size_t, cudaStream_t);
size_t tc_conv_fwd_ws(const b200lic_conv_desc*);
size_t tc_deconv_fwd_ws(const b200lic_conv_desc*);
size_t tc_conv_wgrad_ws(const b200lic_conv_desc*);
size_t tc_deconv_wgrad_ws(const b200lic_conv_desc*);
}  // namespace b200lic

using namespace b200lic;

// AUTO: tensor cores when eligible, else SIMT.  TC: tensor cores or an error.  SIMT: exact-fp32 engine.
// The AUTO -> SIMT fallback is counted (b200lic_simt_fallback_count) and the first one of a process is reported on
// stderr with the reason the tensor-core engine gave: a layer that silently runs 20x slower is a performance bug.
namespace b200lic {
void note_simt_fallback(const char* op);
}
#define DISPATCH(tc_call, simt_call)                                     \
  do {                                                                   \
    if (d->engine != B200LIC_ENGINE_SIMT) {                              \
      int _rc = (tc_call);                                               \
      if (_rc != B200LIC_ERR_UNSUPPORTED || d->engine == B200LIC_ENGINE_TC) return _rc; \
      ::b200lic::note_simt_fallback(#tc_call);                           \
    }                                                                    \
    return (simt_call);                                                  \
  } while (0)

extern "C" {

size_t b200lic_conv_workspace_bytes(const b200lic_conv_desc* d, int op) {
  if (!d) return 0;
  switch (op) {
    case B200LIC_OP_CONV_FWD:
      return tc_conv_fwd_ws(d);
    case B200LIC_OP_DECONV_FWD:
      return tc_deconv_fwd_ws(d);
    case B200LIC_OP_CONV_DGRAD:
      return tc_workspace_bytes(d->N, d->Cout, d->Ho, d->Wo, d->Cin, d->H, d->W, d->KH, d->KW, d->stride, 1);
    case B200LIC_OP_DECONV_DGRAD:
      return tc_workspace_bytes(d->N, d->Cout, d->Ho, d->Wo, d->Cin, d->H, d->W, d->KH, d->KW, d->stride, 0);
    case B200LIC_OP_CONV_WGRAD:
      return tc_conv_wgrad_ws(d);
    case B200LIC_OP_DECONV_WGRAD:
      return tc_deconv_wgrad_ws(d);
    default:
      return 0;
  }
}

int b200lic_conv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, const float* gdn_x,
                     float* norm_out, float* y, void* workspace, size_t workspace_bytes, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "conv_fwd", false);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x && w && y, "conv_fwd: null pointer");
  B200_REQUIRE(!d->gdn_mode || gdn_x, "conv_fwd: gdn_mode needs gdn_x");
  DISPATCH(tc_conv_fwd(d, x, w, bias, gdn_x, norm_out, y, workspace, workspace_bytes, as_stream(stream)),
           simt_conv_fwd(d, x, w, bias, gdn_x, norm_out, y, as_stream(stream)));
}

int b200lic_deconv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                       void* workspace, size_t workspace_bytes, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "deconv_fwd", true);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x && w && y, "deconv_fwd: null pointer");
  B200_REQUIRE(!d->gdn_mode && !d->in_square, "deconv_fwd: GDN flags are conv-only");
  DISPATCH(tc_deconv_fwd(d, x, w, bias, y, workspace, workspace_bytes, as_stream(stream)),
           simt_deconv_fwd(d, x, w, bias, y, as_stream(stream)));
}

int b200lic_conv_fwd_wq(const b200lic_conv_desc* d, const float* x, const float* w_int, const float* w_scale,
                        const float* bias, float* y, void* workspace, size_t workspace_bytes, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "conv_fwd_wq", false);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x && w_int && w_scale && y, "conv_fwd_wq: null pointer");
  B200_REQUIRE(!d->gdn_mode && !d->in_square, "conv_fwd_wq: GDN flags are not supported");
  if (d->engine == B200LIC_ENGINE_SIMT) {
    set_error("conv_fwd_wq: tensor-core engine only");
    return B200LIC_ERR_UNSUPPORTED;
  }
  return tc_conv_fwd_wq(d, x, w_int, w_scale, bias, y, workspace, workspace_bytes, as_stream(stream));
}

int b200lic_deconv_fwd_wq(const b200lic_conv_desc* d, const float* x, const float* w_int, const float* w_scale,
                          const float* bias, float* y, void* workspace, size_t workspace_bytes,
                          b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "deconv_fwd_wq", true);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x && w_int && w_scale && y, "deconv_fwd_wq: null pointer");
  B200_REQUIRE(!d->gdn_mode && !d->in_square, "deconv_fwd_wq: GDN flags are conv-only");
  if (d->engine == B200LIC_ENGINE_SIMT) {
    set_error("deconv_fwd_wq: tensor-core engine only");
    return B200LIC_ERR_UNSUPPORTED;
  }
  return tc_deconv_fwd_wq(d, x, w_int, w_scale, bias, y, workspace, workspace_bytes, as_stream(stream));
}

int b200lic_conv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw, void* workspace,
                       size_t workspace_bytes, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "conv_wgrad", false);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x && dy && dw, "conv_wgrad: null pointer");
  DISPATCH(tc_conv_wgrad(d, x, dy, dw, workspace, workspace_bytes, as_stream(stream)),
           simt_conv_wgrad(d, x, dy, dw, as_stream(stream)));
}

int b200lic_deconv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw, void* workspace,
                         size_t workspace_bytes, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "deconv_wgrad", true);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x && dy && dw, "deconv_wgrad: null pointer");
  DISPATCH(tc_deconv_wgrad(d, x, dy, dw, workspace, workspace_bytes, as_stream(stream)),
           simt_deconv_wgrad(d, x, dy, dw, as_stream(stream)));
}

int b200lic_conv_staged_view(const b200lic_conv_desc* d, int op, void* fwd_workspace, size_t workspace_bytes,
                             void** x_hi, void** x_lo) {
  B200_REQUIRE(d && x_hi && x_lo, "conv_staged_view: null pointer");
  B200_REQUIRE(op == B200LIC_OP_CONV_FWD || op == B200LIC_OP_DECONV_FWD, "conv_staged_view: op must be a forward op");
  *x_hi = *x_lo = nullptr;
  if (d->engine == B200LIC_ENGINE_SIMT) return B200LIC_OK;
  tc_staged_view(d, op == B200LIC_OP_DECONV_FWD, fwd_workspace, workspace_bytes, x_hi, x_lo);
  return B200LIC_OK;
}

int b200lic_conv_wgrad_staged(const b200lic_conv_desc* d, int transposed, const void* x_hi, const void* x_lo,
                              const float* dy, float* dw, void* workspace, size_t workspace_bytes,
                              b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "conv_wgrad_staged", transposed != 0);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x_hi && x_lo && dy && dw, "conv_wgrad_staged: null pointer");
  B200_REQUIRE(d->engine != B200LIC_ENGINE_SIMT, "conv_wgrad_staged: staged operands belong to the tensor-core engine");
  return transposed ? tc_deconv_wgrad_pre(d, x_hi, x_lo, dy, dw, workspace, workspace_bytes, as_stream(stream))
                    : tc_conv_wgrad_pre(d, x_hi, x_lo, dy, dw, workspace, workspace_bytes, as_stream(stream));
}

int b200lic_conv_dgrad(const b200lic_conv_desc* d, const float* dy, const float* w, float* dx, void* workspace,
                       size_t workspace_bytes, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "conv_dgrad", false);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(dy && w && dx, "conv_dgrad: null pointer");
  DISPATCH(tc_conv_dgrad(d, dy, w, dx, workspace, workspace_bytes, as_stream(stream)),
           simt_conv_dgrad(d, dy, w, dx, as_stream(stream)));
}

int b200lic_deconv_dgrad(const b200lic_conv_desc* d, const float* dy, const float* w, float* dx, void* workspace,
                         size_t workspace_bytes, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "deconv_dgrad", true);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(dy && w && dx, "deconv_dgrad: null pointer");
  DISPATCH(tc_deconv_dgrad(d, dy, w, dx, workspace, workspace_bytes, as_stream(stream)),
           simt_deconv_dgrad(d, dy, w, dx, as_stream(stream)));
}

}  // extern "C"
