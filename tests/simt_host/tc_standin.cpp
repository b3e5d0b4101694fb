// TEST INFRASTRUCTURE — functional stand-ins for the entry points whose kernels cannot run on the host (tcgen05 / TMA):
// same C ABI and contract as include/nopesac_b200.h, plain C++ arithmetic.  They exist so that the product's Python glue
// (PlaneCameraHead / MatchingHead) can be executed end to end on CPU tensors in the `-m "not gpu"` suite, with the plain-SIMT
// kernels running from their real source (tests/simt_host/cuda_runtime.h).  They are NOT the product and prove nothing about
// the tensor-core kernels themselves (tests/test_gpu_gemm_tc.py, tests/test_gpu_parity.py do that on the device).
//   nsac_split16 / nsac_gemm_split / nsac_conv3x3_split   16-bit hi/lo planes, products of planes summed in double (at least as accurate as the
//                                    TMEM accumulation), epilogue act(out_scale * acc + bias), optional re-split
//   nsac_score_pack* / nsac_score_aggregate_tc   routed to the exact-fp32 nsac_score_aggregate (csrc/score.cu, real source)
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/nopesac_b200.h"

void nsac_set_error(const char* fmt, ...);

namespace {
float plane_to_float(uint16_t v, int fmt) {
  if (fmt == NSAC_SPLIT_F16) {
    _Float16 h;
    memcpy(&h, &v, 2);
    return (float)h;
  }
  uint32_t bits = (uint32_t)v << 16;
  float f;
  memcpy(&f, &bits, 4);
  return f;
}
uint16_t float_to_plane(float x, int fmt) {
  uint16_t v;
  if (fmt == NSAC_SPLIT_F16) {
    _Float16 h = (_Float16)x;
    memcpy(&v, &h, 2);
    return v;
  }
  uint32_t bits;
  memcpy(&bits, &x, 4);
  if ((bits & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((bits >> 16) | 0x40);          // NaN
  bits += 0x7fffu + ((bits >> 16) & 1u);                                                    // round to nearest even
  return (uint16_t)(bits >> 16);
}
void split16(float x, int fmt, uint16_t& hi, uint16_t& lo) {
  hi = float_to_plane(x, fmt);
  lo = float_to_plane(x - plane_to_float(hi, fmt), fmt);
}
}  // namespace

extern "C" int nsac_split16(const float* x, int ldx, int rows, int K, float scale, int fmt, void* hi, void* lo, int ld_split,
                            void*) {
  if (!x || !hi || !lo || rows < 0 || K < 1 || ldx < K || ld_split < K) {
    nsac_set_error("nsac_split16 (stand-in): bad arguments");
    return NSAC_ERR_ARG;
  }
  uint16_t *h = static_cast<uint16_t*>(hi), *l = static_cast<uint16_t*>(lo);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < ld_split; ++c) split16(c < K ? x[(size_t)r * ldx + c] * scale : 0.f, fmt, h[(size_t)r * ld_split + c], l[(size_t)r * ld_split + c]);
  return NSAC_OK;
}

static int gemm_split_standin(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo, int ldw,
                              const float* bias, int bias_group_rows, int M, int N, int K, int act, int passes, int fmt,
                              float out_scale, float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split,
                              const void* res_hi, const void* res_lo, int ld_res, const float* row_bias = nullptr) {
  // a_lo == NULL: A has no lo plane (exact in 16 bits) - include/nopesac_b200.h
  if (!a_hi || !w_hi || passes < 1 || passes > 4 || (passes >= 3 && !w_lo) || K % 64 != 0 || lda < K ||
      ldw < K || (!out_f32 && !out_hi) || (out_hi && !out_lo) || (res_hi && (!res_lo || ld_res < N))) {
    nsac_set_error("nsac_gemm_split (stand-in): bad arguments");
    return NSAC_ERR_ARG;
  }
  const uint16_t *ah = static_cast<const uint16_t*>(a_hi), *al = static_cast<const uint16_t*>(a_lo);
  const uint16_t *wh = static_cast<const uint16_t*>(w_hi), *wl = static_cast<const uint16_t*>(w_lo);
  std::vector<double> WH((size_t)N * K), WL((size_t)N * K, 0.0);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      WH[(size_t)n * K + k] = plane_to_float(wh[(size_t)n * ldw + k], fmt);
      if (passes >= 3) WL[(size_t)n * K + k] = plane_to_float(wl[(size_t)n * ldw + k], fmt);
    }
  const float slope = act == NSAC_ACT_RELU ? 0.f : (act == NSAC_ACT_LEAKY ? 0.01f : 1.f);
  std::vector<double> AH(K), AL(K);
#pragma omp parallel for firstprivate(AH, AL)
  for (int m = 0; m < M; ++m) {
    for (int k = 0; k < K; ++k) {
      AH[k] = plane_to_float(ah[(size_t)m * lda + k], fmt);
      AL[k] = (passes >= 2 && al) ? plane_to_float(al[(size_t)m * lda + k], fmt) : 0.0;
    }
    const float* brow = bias ? (bias_group_rows > 0 ? bias + (size_t)(m / bias_group_rows) * N : bias) : nullptr;
    for (int n = 0; n < N; ++n) {
      const double *w0 = &WH[(size_t)n * K], *w1 = &WL[(size_t)n * K];
      double acc = 0.0;
      for (int k = 0; k < K; ++k) acc += (AH[k] + AL[k]) * w0[k] + AH[k] * w1[k];              // hi.hi + lo.hi + hi.lo
      if (passes >= 4)
        for (int k = 0; k < K; ++k) acc += AL[k] * w1[k];
      float v = out_scale * (float)acc + (brow ? brow[n] : 0.f) + (row_bias ? row_bias[m] : 0.f);
      if (res_hi)
        v += plane_to_float(static_cast<const uint16_t*>(res_hi)[(size_t)m * ld_res + n], fmt) +
             plane_to_float(static_cast<const uint16_t*>(res_lo)[(size_t)m * ld_res + n], fmt);
      v = fmaxf(v, slope * v + 0.f);
      if (out_f32) out_f32[(size_t)m * ldo + n] = v;
      if (out_hi) split16(v, fmt, static_cast<uint16_t*>(out_hi)[(size_t)m * ld_split + n], static_cast<uint16_t*>(out_lo)[(size_t)m * ld_split + n]);
    }
  }
  return NSAC_OK;
}

extern "C" int nsac_gemm_split(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo, int ldw,
                               const float* bias, int bias_group_rows, int M, int N, int K, int act, int passes, int fmt,
                               float out_scale, float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split, void*) {
  return gemm_split_standin(a_hi, a_lo, lda, w_hi, w_lo, ldw, bias, bias_group_rows, M, N, K, act, passes, fmt, out_scale, out_f32, ldo,
                            out_hi, out_lo, ld_split, nullptr, nullptr, 0);
}

extern "C" int nsac_gemm_split_rowbias(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo, int ldw,
                                       const float* bias, const float* row_bias, int M, int N, int K, int act, int passes, int fmt,
                                       float out_scale, float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split, void*) {
  return gemm_split_standin(a_hi, a_lo, lda, w_hi, w_lo, ldw, bias, 0, M, N, K, act, passes, fmt, out_scale, out_f32, ldo, out_hi, out_lo,
                            ld_split, nullptr, nullptr, 0, row_bias);
}

extern "C" int nsac_gemm_split_residual(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo, int ldw,
                                        const float* bias, int M, int N, int K, int act, int passes, int fmt, float out_scale,
                                        const void* res_hi, const void* res_lo, int ld_res, float* out_f32, int ldo, void* out_hi,
                                        void* out_lo, int ld_split, void*) {
  if (!res_hi || !res_lo) {
    nsac_set_error("nsac_gemm_split_residual (stand-in): null residual planes");
    return NSAC_ERR_ARG;
  }
  return gemm_split_standin(a_hi, a_lo, lda, w_hi, w_lo, ldw, bias, 0, M, N, K, act, passes, fmt, out_scale, out_f32, ldo, out_hi, out_lo,
                            ld_split, res_hi, res_lo, ld_res);
}

// 3x3 / stride 1 / pad 1 convolution over NHWC planes as the same hi/lo-plane product (weights [Cout, 9*Cin], (ky,kx,cin) order)
static int conv3x3_standin(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, int N,
                           int H, int W, int Cin, int Cout, int stride, int act, int passes, int fmt, float out_scale, float* out_f32,
                           int ldo, void* out_hi, void* out_lo, int ld_split, int taps = 9) {
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  if (!x_hi || !w_hi || passes < 1 || passes > 4 || (passes >= 2 && !x_lo) || (passes >= 3 && !w_lo) || Cin % 64 != 0 ||
      (!out_f32 && !out_hi) || (out_hi && !out_lo)) {
    nsac_set_error("nsac_conv3x3_split (stand-in): bad arguments");
    return NSAC_ERR_ARG;
  }
  const int K = taps * Cin;
  const uint16_t *xh = static_cast<const uint16_t*>(x_hi), *xl = static_cast<const uint16_t*>(x_lo);
  const uint16_t *wh = static_cast<const uint16_t*>(w_hi), *wl = static_cast<const uint16_t*>(w_lo);
  std::vector<double> WH((size_t)Cout * K), WL((size_t)Cout * K, 0.0);
  for (size_t i = 0; i < (size_t)Cout * K; ++i) {
    WH[i] = plane_to_float(wh[i], fmt);
    if (passes >= 3) WL[i] = plane_to_float(wl[i], fmt);
  }
  const float slope = act == NSAC_ACT_RELU ? 0.f : (act == NSAC_ACT_LEAKY ? 0.01f : 1.f);
  std::vector<double> AH(K), AL(K);
#pragma omp parallel for firstprivate(AH, AL)
  for (int m = 0; m < N * Ho * Wo; ++m) {
    const int n = m / (Ho * Wo), y = (m / Wo) % Ho, x = m % Wo;
    for (int t = 0; t < taps; ++t) {
      const int yy = y * stride + (taps == 9 ? t / 3 - 1 : 0), xx = x * stride + (taps == 9 ? t % 3 - 1 : 0);
      const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
      const size_t src = (((size_t)n * H + yy) * W + xx) * Cin;
      for (int c = 0; c < Cin; ++c) {
        AH[t * Cin + c] = in ? plane_to_float(xh[src + c], fmt) : 0.0;
        AL[t * Cin + c] = (in && passes >= 2) ? plane_to_float(xl[src + c], fmt) : 0.0;
      }
    }
    for (int co = 0; co < Cout; ++co) {
      const double *w0 = &WH[(size_t)co * K], *w1 = &WL[(size_t)co * K];
      double acc = 0.0;
      for (int k = 0; k < K; ++k) acc += (AH[k] + AL[k]) * w0[k] + AH[k] * w1[k];
      if (passes >= 4)
        for (int k = 0; k < K; ++k) acc += AL[k] * w1[k];
      float v = out_scale * (float)acc + (bias ? bias[co] : 0.f);
      v = fmaxf(v, slope * v + 0.f);
      if (out_f32) out_f32[(size_t)m * ldo + co] = v;
      if (out_hi) split16(v, fmt, static_cast<uint16_t*>(out_hi)[(size_t)m * ld_split + co], static_cast<uint16_t*>(out_lo)[(size_t)m * ld_split + co]);
    }
  }
  return NSAC_OK;
}

extern "C" int nsac_conv3x3_split(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, int N,
                                  int H, int W, int Cin, int Cout, int act, int passes, int fmt, float out_scale, float* out_f32,
                                  int ldo, void* out_hi, void* out_lo, int ld_split, void*) {
  return conv3x3_standin(x_hi, x_lo, w_hi, w_lo, bias, N, H, W, Cin, Cout, 1, act, passes, fmt, out_scale, out_f32, ldo, out_hi, out_lo, ld_split);
}
extern "C" int nsac_conv3x3_split_strided(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, int N,
                                          int H, int W, int Cin, int Cout, int stride, int act, int passes, int fmt, float out_scale,
                                          float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split, void*) {
  if (stride != 1 && stride != 2) {
    nsac_set_error("nsac_conv3x3_split_strided (stand-in): stride must be 1 or 2");
    return NSAC_ERR_ARG;
  }
  return conv3x3_standin(x_hi, x_lo, w_hi, w_lo, bias, N, H, W, Cin, Cout, stride, act, passes, fmt, out_scale, out_f32, ldo, out_hi, out_lo, ld_split);
}

extern "C" int nsac_conv1x1_split_strided(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, int N,
                                          int H, int W, int Cin, int Cout, int stride, int act, int passes, int fmt, float out_scale,
                                          float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split, void*) {
  if (stride != 1 && stride != 2) {
    nsac_set_error("nsac_conv1x1_split_strided (stand-in): stride must be 1 or 2");
    return NSAC_ERR_ARG;
  }
  return conv3x3_standin(x_hi, x_lo, w_hi, w_lo, bias, N, H, W, Cin, Cout, stride, act, passes, fmt, out_scale, out_f32, ldo, out_hi, out_lo,
                         ld_split, 1);
}

// ---- scoring on the "tensor pipe": the exact CUDA-core twin does the work
extern "C" size_t nsac_score_pack_bytes(int) { return 2 * sizeof(nsac_score_mlp); }
extern "C" int nsac_score_pack(const nsac_score_mlp* rot_mlp, const nsac_score_mlp* tran_mlp, int, void* pack, void*) {
  memcpy(pack, rot_mlp, sizeof(nsac_score_mlp));
  memcpy(static_cast<char*>(pack) + sizeof(nsac_score_mlp), tran_mlp, sizeof(nsac_score_mlp));
  return NSAC_OK;
}
extern "C" size_t nsac_score_tc_workspace_bytes(int B, int NQ) { return nsac_score_workspace_bytes(B, NQ); }
extern "C" int nsac_score_aggregate_tc(const float* geo_local, const float* q_h, const float* t_h, const float* q0, const float* t0,
                                       const float* feat_rot, const float* feat_tran, const float* feat_rot0,
                                       const float* feat_tran0, const int32_t* matched_num, const void* pack, const float* w_rots,
                                       const float* b_rots, const float* w_trans, const float* b_trans, int B, int NQ,
                                       int out_cam_type, float* pose, float* score_rot, float* score_tran, int32_t* sel_idx,
                                       void* workspace, float* const*, int, int, void* stream) {
  const nsac_score_mlp* mlps = static_cast<const nsac_score_mlp*>(pack);
  return nsac_score_aggregate(geo_local, q_h, t_h, q0, t0, feat_rot, feat_tran, feat_rot0, feat_tran0, matched_num, &mlps[0], &mlps[1],
                              w_rots, b_rots, w_trans, b_trans, B, NQ, out_cam_type, pose, score_rot, score_tran, sel_idx, nullptr,
                              workspace, stream);
}

// host-mirror variant: the stand-in has no vectors to mirror (offset 0 = "nothing to copy", the wrapper then passes NULL)
extern "C" size_t nsac_score_pack_vecs_offset(int) { return 0; }
extern "C" int nsac_score_aggregate_tc_cv(const float* geo_local, const float* q_h, const float* t_h, const float* q0, const float* t0,
                                          const float* feat_rot, const float* feat_tran, const float* feat_rot0,
                                          const float* feat_tran0, const int32_t* matched_num, const void* pack, const float*,
                                          const float* w_rots, const float* b_rots, const float* w_trans, const float* b_trans, int B,
                                          int NQ, int out_cam_type, float* pose, float* score_rot, float* score_tran,
                                          int32_t* sel_idx, void* workspace, float* const* peers, int np, int ro, void* stream) {
  return nsac_score_aggregate_tc(geo_local, q_h, t_h, q0, t0, feat_rot, feat_tran, feat_rot0, feat_tran0, matched_num, pack, w_rots,
                                 b_rots, w_trans, b_trans, B, NQ, out_cam_type, pose, score_rot, score_tran, sel_idx, workspace, peers,
                                 np, ro, stream);
}
