// Shared declarations of the prepared-operand path (prepared.cu, weight_quant.cu, conv_tc_wgrad.cu, conv_tc2.cu).
#pragma once
#include "common.cuh"

namespace b200lic {

// AdaRound backward + Adam arguments of the fused weight-gradient tail (b200lic_conv_wgrad_adam_sched).
struct WgTail {
  const float* w;
  float* alpha;
  const float* delta;
  const float* zp;
  float* m;
  float* v;
  int outer, ch, inner;
  float top;
  const b200lic_calib_sched* sched;
  float beta1, beta2, eps, grad_scale, reg_weight;
  float* reg_loss;
  float* dw_out;      // may be nullptr: the summed weight gradient is not materialised
};

// Geometry of the packed weight operand (conv_tc.cu's PackGeom with the destination slabs).
struct PackDst {
  int Cout, Cin, KH, KW, stride, pad, transposed;
  int CoutPad, Cpad, Tmax, phases;
  long long s_co, s_ci;
  size_t b_bytes;
};

bool tc2_weight_layout(int Cin, int Cout, int KH, int KW, int stride, int transposed, int* Cpad, int* CoutPad, int* Tmax,
                       int* phases, size_t* b_bytes);

}  // namespace b200lic
