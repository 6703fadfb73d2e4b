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
// tensor-core engine (conv_tc.cu): returns B200LIC_ERR_UNSUPPORTED when the shape does not qualify
int tc_conv_fwd(const b200lic_conv_desc*, const float*, const float*, const float*, const float*, float*, float*,
                cudaStream_t);
}  // namespace b200lic

using namespace b200lic;

extern "C" {

int b200lic_conv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, const float* gdn_x,
                     float* norm_out, float* y, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "conv_fwd", false);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x && w && y, "conv_fwd: null pointer");
  B200_REQUIRE(!d->gdn_mode || gdn_x, "conv_fwd: gdn_mode needs gdn_x");
  if (d->engine != B200LIC_ENGINE_SIMT) {
    rc = tc_conv_fwd(d, x, w, bias, gdn_x, norm_out, y, as_stream(stream));
    if (rc != B200LIC_ERR_UNSUPPORTED) return rc;
    if (d->engine == B200LIC_ENGINE_TC) return rc;
  }
  return simt_conv_fwd(d, x, w, bias, gdn_x, norm_out, y, as_stream(stream));
}

int b200lic_deconv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                       b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "deconv_fwd", true);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x && w && y, "deconv_fwd: null pointer");
  B200_REQUIRE(!d->gdn_mode && !d->in_square, "deconv_fwd: GDN flags are conv-only");
  if (d->engine == B200LIC_ENGINE_TC) {
    set_error("deconv_fwd: tensor-core engine not available for this op yet");
    return B200LIC_ERR_UNSUPPORTED;
  }
  return simt_deconv_fwd(d, x, w, bias, y, as_stream(stream));
}

int b200lic_conv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "conv_wgrad", false);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x && dy && dw, "conv_wgrad: null pointer");
  return simt_conv_wgrad(d, x, dy, dw, as_stream(stream));
}

int b200lic_deconv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw,
                         b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "deconv_wgrad", true);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(x && dy && dw, "deconv_wgrad: null pointer");
  return simt_deconv_wgrad(d, x, dy, dw, as_stream(stream));
}

int b200lic_conv_dgrad(const b200lic_conv_desc* d, const float* dy, const float* w, float* dx, b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "conv_dgrad", false);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(dy && w && dx, "conv_dgrad: null pointer");
  return simt_conv_dgrad(d, dy, w, dx, as_stream(stream));
}

int b200lic_deconv_dgrad(const b200lic_conv_desc* d, const float* dy, const float* w, float* dx,
                         b200lic_stream_t stream) {
  B200_ARCH_GATE();
  int rc = conv_check_desc(d, "deconv_dgrad", true);
  if (rc != B200LIC_OK) return rc;
  B200_REQUIRE(dy && w && dx, "deconv_dgrad: null pointer");
  return simt_deconv_dgrad(d, dy, w, dx, as_stream(stream));
}

}  // extern "C"
