// libb200lic: error reporting, device gate and launch accounting (include/b200lic.h).
#include <atomic>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"

namespace b200lic {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

struct DevInfo {
  int major = -1, minor = -1, sms = 0;
};
static DevInfo g_dev[64];

static DevInfo* dev_info() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  DevInfo* d = &g_dev[dev];
  if (d->major < 0) {
    int major = 0, minor = 0, sms = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return nullptr;
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    d->minor = minor;
    d->sms = sms;
    d->major = major;
  }
  return d;
}

int check_arch() {
  DevInfo* d = dev_info();
  if (!d) {
    set_error("no CUDA device available (cudaGetDevice failed): libb200lic has no CPU fallback");
    (void)cudaGetLastError();
    return B200LIC_ERR_ARCH;
  }
  if (d->major != 10) {
    set_error("device is sm_%d%d; libb200lic is built for sm_100a (B200) only", d->major, d->minor);
    return B200LIC_ERR_ARCH;
  }
  return B200LIC_OK;
}

int num_sms() {
  DevInfo* d = dev_info();
  return (d && d->sms > 0) ? d->sms : 148;
}

}  // namespace b200lic

namespace b200lic {
static std::atomic<unsigned long long> g_simt_fallbacks{0};
void note_simt_fallback(const char* op) {
  if (g_simt_fallbacks.fetch_add(1, std::memory_order_relaxed) == 0 && !getenv("B200LIC_QUIET")) {
    char name[48];
    size_t n = 0;
    while (op[n] && op[n] != '(' && n + 1 < sizeof(name)) {
      name[n] = op[n];
      ++n;
    }
    name[n] = 0;
    fprintf(stderr, "libb200lic: %s -> exact-fp32 SIMT engine under B200LIC_ENGINE_AUTO (%s); first occurrence, further ones "
                    "are only counted (b200lic_simt_fallback_count)\n", name, g_err);
  }
}
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("B200LIC_PDL");
    on = (e && atoi(e) == 0) ? 0 : 1;
  }
  return on != 0;
}
int tc2_debug_timeline(unsigned long long* out, int n);
void tc2_set_streamk_mode(int v);
void tc2_set_pair_mode(int v);
void tc2_stats_once(unsigned* keys);
unsigned* tc2_stats_peek();
void tc2_plan_info(int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int KH, int KW, int stride, int transposed,
                   int gdn_mode, int* info);
void gemm1x1_set_mode(int v);
}
extern "C" {
int b200lic_set_option(const char* name, int value) {
  if (name && strcmp(name, "streamk") == 0) {
    b200lic::tc2_set_streamk_mode(value);
    return B200LIC_OK;
  }
  if (name && strcmp(name, "pair") == 0) {
    b200lic::tc2_set_pair_mode(value);
    return B200LIC_OK;
  }
  if (name && strcmp(name, "gemm1x1") == 0) {      // 0: the generic engine also runs the short-K 1x1 layers (A/B, tests)
    b200lic::gemm1x1_set_mode(value);
    return B200LIC_OK;
  }
  b200lic::set_error("set_option: unknown option '%s'", name ? name : "(null)");
  return B200LIC_ERR_ARG;
}
int b200lic_conv_plan_info(const b200lic_conv_desc* d, int op, int* info) {
  if (!d || !info || (op != B200LIC_OP_CONV_FWD && op != B200LIC_OP_DECONV_FWD)) {
    b200lic::set_error("conv_plan_info: needs a descriptor, a forward op and 12 ints");
    return B200LIC_ERR_ARG;
  }
  b200lic::tc2_plan_info(d->N, d->Cin, d->H, d->W, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride,
                         op == B200LIC_OP_DECONV_FWD ? 1 : 0, d->gdn_mode, info);
  return B200LIC_OK;
}
int b200lic_conv_stats_once(float* minmax) {
  b200lic::tc2_stats_once(reinterpret_cast<unsigned*>(minmax));
  return B200LIC_OK;
}
int b200lic_conv_stats_pending(void) {
  const int pending = b200lic::tc2_stats_peek() != nullptr ? 1 : 0;
  b200lic::tc2_stats_once(nullptr);
  return pending;
}
int b200lic_debug_timeline(unsigned long long* out, int n) { return b200lic::tc2_debug_timeline(out, n); }
int b200lic_version(void) { return 200; }
const char* b200lic_last_error_string(void) { return b200lic::g_err; }
int b200lic_device_check(void) { return b200lic::check_arch(); }
unsigned long long b200lic_launch_count(void) { return b200lic::g_launches.load(); }
unsigned long long b200lic_simt_fallback_count(void) { return b200lic::g_simt_fallbacks.load(); }
}
