// Whole-stage entry points: host-side orchestration of this library's own kernels behind one C call per stage (no kernels in
// this file).  Mirrors what nopesac_b200/camera_head.py does in Python, launch for launch, so both give identical bits.
#include "common.cuh"

namespace {

struct Planes {
  void* hi; void* lo; int ld;
  Planes cols(int c0) const { return {static_cast<uint16_t*>(hi) + c0, static_cast<uint16_t*>(lo) + c0, ld}; }
};

// bump allocator over the caller's workspace (256-byte granules); base == nullptr only counts
struct Arena {
  uint8_t* base; size_t off;
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  }
  Planes planes(size_t rows, int ld) {
    Planes p;
    p.hi = take(rows * ld * 2); p.lo = take(rows * ld * 2); p.ld = ld;
    return p;
  }
};

struct RefineScratch {
  float *geo8, *x32, *gb_rot, *gb_tran, *fused_rot, *fused_tran;
  Planes p0, p1, p2, cat;
  void* score_ws; size_t score_ws_bytes;
};

size_t carve_refine(Arena& a, int B, int NQ, RefineScratch& s) {
  const size_t rows = (size_t)B * NQ;
  s.geo8 = static_cast<float*>(a.take(rows * 8 * 4));
  s.x32 = static_cast<float*>(a.take(rows * 1024 * 4));            // geo_encoder layer 0 output (fp32, re-split)
  s.p0 = a.planes(rows, 1024); s.p1 = a.planes(rows, 1024); s.p2 = a.planes(rows, 1024);
  s.cat = a.planes(rows, 1280);                                     // cat[s1 (1024), rot (256)] of camera_head.py:961
  s.gb_rot = static_cast<float*>(a.take((size_t)B * 512 * 4));     // per-pair bias rows of decoder_*2.layers.0
  s.gb_tran = static_cast<float*>(a.take((size_t)B * 512 * 4));
  s.fused_rot = static_cast<float*>(a.take(rows * 256 * 4));
  s.fused_tran = static_cast<float*>(a.take(rows * 256 * 4));
  const size_t tc = nsac_score_tc_workspace_bytes(B, NQ), f32 = nsac_score_workspace_bytes(B, NQ);
  s.score_ws_bytes = tc > f32 ? tc : f32;
  s.score_ws = a.take(s.score_ws_bytes);
  return a.off;
}

#define NSAC_TRY(call)            \
  do {                            \
    int st__ = (call);            \
    if (st__ != NSAC_OK) return st__; \
  } while (0)

// MLP on planes: ReLU between the layers, `final_act` after the last.  Intermediates ping-pong between t0 / t1 (neither may
// alias `in`); the last layer writes `out` planes (hi may be NULL) and / or out_f32.  first_bias / first_group: per-group bias
// rows of layer 0 (cat[init_feat, geo_feat] W^T = geo_feat W_geo^T + (init_feat W_init^T + b)).
int run_chain(const nsac_tc_layer* L, int n, Planes in, int M, Planes t0, Planes t1, const float* first_bias, int first_group,
              int final_act, float* out_f32, int ldo, Planes out, int fmt, int passes, void* stream, int& launches) {
  for (int i = 0; i < n; ++i) {
    const bool last = i == n - 1;
    const Planes dst = last ? out : ((i & 1) ? t1 : t0);
    const float* bias = (i == 0 && first_bias) ? first_bias : L[i].bias;
    NSAC_TRY(nsac_gemm_split(in.hi, in.lo, in.ld, L[i].w_hi, L[i].w_lo, L[i].ldw, bias, i == 0 ? first_group : 0, M, L[i].N, L[i].K,
                             last ? final_act : NSAC_ACT_RELU, passes, fmt, 1.0f / L[i].w_scale, last ? out_f32 : nullptr,
                             last ? ldo : 0, dst.hi, dst.lo, dst.hi ? dst.ld : 0, stream));
    ++launches;
    in = dst;
  }
  return NSAC_OK;
}

}  // namespace

extern "C" size_t nsac_refine_workspace_bytes(int B, int NQ) {
  if (B <= 0 || NQ <= 0) return 0;
  Arena a{nullptr, 0};
  RefineScratch s;
  return carve_refine(a, B, NQ, s);
}

extern "C" int nsac_refine_forward(const nsac_refine_weights* w, const float* planes1, const float* planes2, const float* assign,
                                   const int32_t* hyp_pairs, int H, const float* t0, const float* q0, const float* rot_feat0,
                                   const float* trans_feat0, int B, int n1, int n2, int NQ, int out_cam_type, float* pose,
                                   float* assign_pruned, float* geo_local, float* geo_global, float* sig, int32_t* matched_num,
                                   int32_t* pair_idx, float* q_h, float* t_h, float* score_rot, float* score_tran,
                                   int32_t* sel_idx, void* workspace, size_t workspace_bytes, float* const* peer_rows,
                                   int num_peers, int row_offset, int* launches_out, void* stream) {
  NSAC_REQUIRE(w && planes1 && planes2 && t0 && q0 && rot_feat0 && trans_feat0, "nsac_refine_forward: null input");
  NSAC_REQUIRE(assign || hyp_pairs, "nsac_refine_forward: need the assignment matrix or a hypothesis list");
  NSAC_REQUIRE(!assign_pruned || assign, "nsac_refine_forward: assign_pruned needs assign");
  NSAC_REQUIRE(pose && geo_local && geo_global && sig && matched_num && pair_idx && q_h && t_h && sel_idx,
               "nsac_refine_forward: null output");
  NSAC_REQUIRE(B > 0 && n1 > 0 && n2 > 0 && NQ > 0, "nsac_refine_forward: bad sizes B=%d n1=%d n2=%d NQ=%d", B, n1, n2, NQ);
  NSAC_REQUIRE(w->geo_encoder[0].K == 1024 && w->geo_proj_s2[0].K == 1280 && w->decoder_rot2[0].K == 256 &&
               w->decoder_rot[5].N == 256 && w->decoder_tran[5].N == 256 && w->decoder_rot2[2].N == 256,
               "nsac_refine_forward: weight struct does not have the PlaneCameraHead layer shapes");
  Arena a{static_cast<uint8_t*>(workspace), 0};
  RefineScratch s;
  const size_t need = carve_refine(a, B, NQ, s);
  NSAC_REQUIRE(workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "nsac_refine_forward: workspace of %zu bytes (256-byte aligned) needed, got %zu", need, workspace_bytes);
  const int rows = B * NQ, fmt = w->fmt, P = w->passes;
  int n = 0;
  // K6: geo sequences in torch.nonzero order, sig, the 8-vector geo encoding (camera_head.py:513-569, 937-957)
  NSAC_TRY(nsac_geo_sequence(planes1, planes2, hyp_pairs ? nullptr : assign, hyp_pairs, hyp_pairs ? H : 0, t0, q0, B, n1, n2, NQ,
                             geo_local, geo_global, sig, s.geo8, matched_num, pair_idx, stream));
  ++n;
  // K7 (:957-986).  geo_encoder layer 0 has K = 8: exact fp32 on the CUDA cores, then split into planes
  NSAC_TRY(nsac_linear(s.geo8, 8, w->geo0_w, w->geo0_b, 0, s.x32, 1024, rows, 1024, 8, NSAC_ACT_RELU, stream));
  NSAC_TRY(nsac_split16(s.x32, 1024, rows, 1024, 1.0f, fmt, s.p0.hi, s.p0.lo, s.p0.ld, stream));
  n += 2;
  const Planes none{nullptr, nullptr, 0};
  NSAC_TRY(run_chain(w->geo_encoder, 5, s.p0, rows, s.p1, s.p2, nullptr, 0, NSAC_ACT_NONE, nullptr, 0, s.p0, fmt, P, stream, n));
  NSAC_TRY(run_chain(w->geo_proj_s1, 3, s.p0, rows, s.p1, s.p2, nullptr, 0, NSAC_ACT_NONE, nullptr, 0, s.cat, fmt, P, stream, n));
  NSAC_TRY(run_chain(w->decoder_rot, 6, s.cat, rows, s.p1, s.p2, nullptr, 0, NSAC_ACT_NONE, nullptr, 0, s.cat.cols(1024), fmt, P,
                     stream, n));
  NSAC_TRY(run_chain(w->geo_proj_s2, 3, s.cat, rows, s.p1, s.p2, nullptr, 0, NSAC_ACT_NONE, nullptr, 0, s.p0, fmt, P, stream, n));
  NSAC_TRY(run_chain(w->decoder_tran, 6, s.p0, rows, s.p1, s.p2, nullptr, 0, NSAC_ACT_NONE, nullptr, 0, s.p0, fmt, P, stream, n));
  // decoder_*2 on cat[init_feat (per pair), geo_feat (per hypothesis)] (:983-986): the init half is a per-pair bias row
  NSAC_TRY(nsac_linear(rot_feat0, 256, w->rot2_w_init, w->rot2_b0, 0, s.gb_rot, 512, B, 512, 256, NSAC_ACT_NONE, stream));
  ++n;
  NSAC_TRY(run_chain(w->decoder_rot2, 3, s.cat.cols(1024), rows, s.p1, s.p2, s.gb_rot, NQ, NSAC_ACT_RELU, s.fused_rot, 256, none, fmt,
                     P, stream, n));
  NSAC_TRY(nsac_linear(trans_feat0, 256, w->tran2_w_init, w->tran2_b0, 0, s.gb_tran, 512, B, 512, 256, NSAC_ACT_NONE, stream));
  ++n;
  NSAC_TRY(run_chain(w->decoder_tran2, 3, s.p0, rows, s.p1, s.p2, s.gb_tran, NQ, NSAC_ACT_RELU, s.fused_tran, 256, none, fmt, P,
                     stream, n));
  // one pose hypothesis per matched plane pair (:990, 1018)
  NSAC_TRY(nsac_pose_heads(s.fused_rot, s.fused_tran, w->rots_w, w->rots_b, w->trans_w, w->trans_b, rows, 256, q_h, t_h, stream));
  ++n;
  // K8 + K9: scoring, softmax over the hypotheses, selection.  'max-score' is a discrete decision on the scores themselves: exact fp32
  if (out_cam_type == NSAC_CAM_MAX_SCORE) {
    NSAC_REQUIRE(w->rot_mlp && w->tran_mlp, "nsac_refine_forward: NSAC_CAM_MAX_SCORE needs the raw fp32 score MLPs (rot_mlp / tran_mlp)");
    NSAC_REQUIRE(peer_rows == nullptr, "nsac_refine_forward: the fused result exchange is not available with NSAC_CAM_MAX_SCORE");
    NSAC_TRY(nsac_score_aggregate(geo_local, q_h, t_h, q0, t0, s.fused_rot, s.fused_tran, rot_feat0, trans_feat0, matched_num,
                                  w->rot_mlp, w->tran_mlp, w->rots_w, w->rots_b, w->trans_w, w->trans_b, B, NQ, out_cam_type, pose,
                                  score_rot, score_tran, sel_idx, nullptr, s.score_ws, stream));
  } else {
    NSAC_REQUIRE(w->score_pack, "nsac_refine_forward: score_pack missing");
    NSAC_TRY(nsac_score_aggregate_tc_cv(geo_local, q_h, t_h, q0, t0, s.fused_rot, s.fused_tran, rot_feat0, trans_feat0, matched_num,
                                        w->score_pack, w->score_vecs_host, w->rots_w, w->rots_b, w->trans_w, w->trans_b, B, NQ,
                                        out_cam_type, pose, score_rot, score_tran, sel_idx, s.score_ws, peer_rows, num_peers,
                                        row_offset, stream));
  }
  n += 3;
  // K10: assignment pruning with the refined pose (:605-629)
  if (assign && assign_pruned) {
    NSAC_TRY(nsac_prune_assignment(assign, planes1, planes2, pose, 16, B, n1, n2, assign_pruned, stream));
    ++n;
  }
  if (launches_out) *launches_out = n;
  return NSAC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// MatchingHead forward (nopesac_b200/matching_head.py, launch for launch)
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct MatchScratch {
  float *X, *qkv, *msg, *desc;
  Planes Xp, app, msgp, hp;
  int32_t* count12;
};

size_t carve_match(Arena& a, int B, int n1, int n2, MatchScratch& s) {
  const size_t R = (size_t)B * (n1 + n2);
  s.X = static_cast<float*>(a.take(R * 512 * 4));       // cols 0:256 token, 256:512 message slot: cat[x, message] (gnn.py:93) is free
  s.Xp = a.planes(R, 512);
  s.app = a.planes(R, 256);
  s.qkv = static_cast<float*>(a.take(R * 768 * 4));
  s.msgp = a.planes(R, 256);
  s.msg = static_cast<float*>(a.take(R * 256 * 4));
  s.hp = a.planes(R, 512);
  s.desc = static_cast<float*>(a.take(R * 256 * 4));
  s.count12 = static_cast<int32_t*>(a.take((size_t)2 * B * 4));
  return a.off;
}

inline Planes prow(Planes p, size_t row) {
  return {static_cast<uint16_t*>(p.hi) + row * p.ld, static_cast<uint16_t*>(p.lo) + row * p.ld, p.ld};
}

int tc(const nsac_tc_layer& L, Planes a, int M, int act, float* out_f32, int ldo, Planes out, int fmt, int passes, void* stream,
       int& launches) {
  ++launches;
  return nsac_gemm_split(a.hi, a.lo, a.ld, L.w_hi, L.w_lo, L.ldw, L.bias, 0, M, L.N, L.K, act, passes, fmt, 1.0f / L.w_scale, out_f32, ldo,
                         out.hi, out.lo, out.hi ? out.ld : 0, stream);
}

// one TransformerEncoderLayer (gnn.py:73-97) for the query rows [x0, x1) against the source rows starting at s0
int gnn_layer(const nsac_gnn_layer& w, MatchScratch& s, size_t x0, size_t x1, size_t s0, size_t s1, int B, int L, int S,
              const int32_t* kv_count, int fmt, int P, void* stream, int& n) {
  const int rows = (int)(x1 - x0);
  float* x = s.X + x0 * 512;
  const Planes xp = prow(s.Xp, x0), none{nullptr, nullptr, 0};
  const float *q, *k, *v;
  int ldq, ldkv;
  if (w.self_attn) {
    NSAC_TRY(tc(w.qkv, xp, rows, NSAC_ACT_NONE, s.qkv, 768, none, fmt, P, stream, n));
    q = s.qkv; k = s.qkv + 256; v = s.qkv + 512; ldq = ldkv = 768;
  } else {
    float* kv = s.qkv + (size_t)rows * 256;
    NSAC_TRY(tc(w.q, xp, rows, NSAC_ACT_NONE, s.qkv, 256, none, fmt, P, stream, n));
    NSAC_TRY(tc(w.kv, prow(s.Xp, s0), (int)(s1 - s0), NSAC_ACT_NONE, kv, 512, none, fmt, P, stream, n));
    q = s.qkv; k = kv; v = kv + 256; ldq = 256; ldkv = 512;
  }
  if (kv_count) {
    NSAC_TRY(nsac_attention_ragged(q, ldq, k, v, ldkv, nullptr, 0, s.msgp.hi, s.msgp.lo, s.msgp.ld, B, L, S, 8, 32, kv_count, stream));
  } else {
    NSAC_TRY(nsac_attention(q, ldq, k, v, ldkv, nullptr, 0, s.msgp.hi, s.msgp.lo, s.msgp.ld, B, L, S, 8, 32, stream));
  }
  ++n;
  NSAC_TRY(tc(w.merge, s.msgp, rows, NSAC_ACT_NONE, s.msg, 256, none, fmt, P, stream, n));
  NSAC_TRY(nsac_layernorm(s.msg, 256, w.n1w, w.n1b, nullptr, 0, x + 256, 512, static_cast<uint16_t*>(xp.hi) + 256,
                          static_cast<uint16_t*>(xp.lo) + 256, 512, rows, 256, stream));                       // message slot
  ++n;
  NSAC_TRY(tc(w.mlp0, xp, rows, NSAC_ACT_RELU, nullptr, 0, s.hp, fmt, P, stream, n));
  NSAC_TRY(tc(w.mlp2, s.hp, rows, NSAC_ACT_NONE, s.msg, 256, none, fmt, P, stream, n));
  NSAC_TRY(nsac_layernorm(s.msg, 256, w.n2w, w.n2b, x, 512, x, 512, xp.hi, xp.lo, 512, rows, 256, stream));     // x + norm2(.)
  ++n;
  return NSAC_OK;
}

}  // namespace

extern "C" size_t nsac_match_workspace_bytes(int B, int n1, int n2) {
  if (B <= 0 || n1 <= 0 || n2 <= 0) return 0;
  Arena a{nullptr, 0};
  MatchScratch s;
  return carve_match(a, B, n1, n2, s);
}

extern "C" int nsac_match_forward(const nsac_match_weights* w, const float* app1, const float* app2, const float* planes1,
                                  const float* planes2, const float* cam, const int32_t* count1, const int32_t* count2,
                                  float match_threshold, int B, int n1, int n2, float* log_scores_padded, float* assign,
                                  void* workspace, size_t workspace_bytes, int* launches_out, void* stream) {
  NSAC_REQUIRE(w && w->layers && w->num_layers > 0 && w->bin_score, "nsac_match_forward: null weights");
  NSAC_REQUIRE(app1 && app2 && planes1 && planes2 && cam && log_scores_padded && assign, "nsac_match_forward: null argument");
  NSAC_REQUIRE((count1 == nullptr) == (count2 == nullptr), "nsac_match_forward: count1 and count2 go together");
  NSAC_REQUIRE(B > 0 && n1 > 0 && n2 > 0, "nsac_match_forward: bad sizes B=%d n1=%d n2=%d", B, n1, n2);
  Arena a{static_cast<uint8_t*>(workspace), 0};
  MatchScratch s;
  const size_t need = carve_match(a, B, n1, n2, s);
  NSAC_REQUIRE(workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "nsac_match_forward: workspace of %zu bytes (256-byte aligned) needed, got %zu", need, workspace_bytes);
  const size_t R0 = (size_t)B * n1, R1 = (size_t)B * n2, R = R0 + R1;
  const int fmt = w->fmt, P = w->passes;
  const Planes none{nullptr, nullptr, 0};
  int n = 0;
  // planeApp_proj on the rows of both views (view 1 after view 0), straight into the token half of X and its planes
  NSAC_TRY(nsac_split16(app1, 256, (int)R0, 256, 1.0f, fmt, s.app.hi, s.app.lo, 256, stream));
  NSAC_TRY(nsac_split16(app2, 256, (int)R1, 256, 1.0f, fmt, prow(s.app, R0).hi, prow(s.app, R0).lo, 256, stream));
  n += 2;
  NSAC_TRY(tc(w->app_proj, s.app, (int)R, NSAC_ACT_NONE, s.X, 512, s.Xp, fmt, P, stream, n));
  const int32_t* count12 = nullptr;
  if (count1) {
    NSAC_CUDA(cudaMemcpyAsync(s.count12, count1, (size_t)B * 4, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    NSAC_CUDA(cudaMemcpyAsync(s.count12 + B, count2, (size_t)B * 4, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    count12 = s.count12;
  }
  for (int i = 0; i < w->num_layers; ++i) {
    const nsac_gnn_layer& l = w->layers[i];
    if (l.self_attn && n1 == n2) {
      NSAC_TRY(gnn_layer(l, s, 0, R, 0, R, 2 * B, n1, n1, count12, fmt, P, stream, n));
    } else if (l.self_attn) {
      NSAC_TRY(gnn_layer(l, s, 0, R0, 0, R0, B, n1, n1, count1, fmt, P, stream, n));
      NSAC_TRY(gnn_layer(l, s, R0, R, R0, R, B, n2, n2, count2, fmt, P, stream, n));
    } else {
      NSAC_TRY(gnn_layer(l, s, 0, R0, R0, R, B, n1, n2, count2, fmt, P, stream, n));
      NSAC_TRY(gnn_layer(l, s, R0, R, 0, R0, B, n2, n1, count1, fmt, P, stream, n));      // sees the UPDATED view-0 tokens (gnn.py:133-134)
    }
  }
  NSAC_TRY(tc(w->desc_proj, s.Xp, (int)R, NSAC_ACT_NONE, s.desc, 256, none, fmt, P, stream, n));
  if (count1) {
    NSAC_TRY(nsac_match_sinkhorn_assign_ragged(s.desc, s.desc + R0 * 256, planes1, planes2, cam, w->bin_score, w->offset_multiplier,
                                               w->normal_multiplier, w->sinkhorn_iterations, match_threshold, B, n1, n2, 256, count1,
                                               count2, log_scores_padded, assign, stream));
  } else {
    NSAC_TRY(nsac_match_sinkhorn_assign(s.desc, s.desc + R0 * 256, planes1, planes2, cam, w->bin_score, w->offset_multiplier,
                                        w->normal_multiplier, w->sinkhorn_iterations, match_threshold, B, n1, n2, 256,
                                        log_scores_padded, assign, stream));
  }
  ++n;
  if (launches_out) *launches_out = n;
  return NSAC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// K1 + K2: pixel pose network + AIM (nopesac_b200/camera_head.py _forward_pixel_camera_head / _forward_rec_heads)
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct PixelScratch {
  float *t5, *y5, *t4, *y4, *t3, *f, *tc0, *ysl, *ya, *yb, *feat_t, *feat_r, *x32, *teps;
  Planes u4, u3, y3, xa, xb, aff, cols, e0, e1, e2;
  void* gn_ws; size_t gn_ws_bytes;
};

size_t carve_pixel(Arena& a, int B, int H3, int W3, PixelScratch& s) {
  const size_t N = 2 * (size_t)B, r3 = N * H3 * W3, r4 = N * (H3 / 2) * (W3 / 2), r5 = N * (H3 / 4) * (W3 / 4);
  const size_t hw = (size_t)(H3 / 4) * (W3 / 4), rb = (size_t)B * hw;
  auto f32 = [&](size_t n) { return static_cast<float*>(a.take(n * 4)); };
  s.t5 = f32(r5 * 128); s.y5 = f32(r5 * 128);
  s.t4 = f32(r4 * 128); s.y4 = f32(r4 * 128); s.u4 = a.planes(r4, 128);
  s.t3 = f32(r3 * 128); s.u3 = a.planes(r3, 128); s.y3 = a.planes(r3, 128);
  s.xa = a.planes(r3, 256); s.xb = a.planes(r3, 256); s.f = f32(r3 * 256);
  const int Cp = (int)((hw + 63) / 64 * 64);
  s.aff = a.planes(rb, Cp);
  s.tc0 = f32(rb * 256); s.ysl = f32(rb * 128); s.ya = f32(rb * 128); s.yb = f32(rb * 128);
  s.cols = a.planes(rb, 1152);
  s.feat_t = f32((size_t)B * 256); s.feat_r = f32((size_t)B * 256);
  s.x32 = f32((size_t)B * 256); s.teps = f32((size_t)B * 3);
  s.e0 = a.planes(B, 256); s.e1 = a.planes(B, 256); s.e2 = a.planes(B, 256);
  s.gn_ws_bytes = nsac_groupnorm_ws_bytes((int)N, H3, W3, 32);
  if (s.gn_ws_bytes < 16) s.gn_ws_bytes = 16;
  s.gn_ws = a.take(s.gn_ws_bytes);
  return a.off;
}

int conv3(const nsac_tc_layer& L, Planes x, int N, int H, int W, int Cin, int act, float* out_f32, Planes out, int fmt, int P,
          void* stream, int& n) {
  ++n;
  return nsac_conv3x3_split(x.hi, x.lo, L.w_hi, L.w_lo, L.bias, N, H, W, Cin, L.N, act, P, fmt, 1.0f / L.w_scale, out_f32,
                            out_f32 ? L.N : 0, out.hi, out.lo, out.hi ? out.ld : 0, stream);
}

}  // namespace

extern "C" size_t nsac_pixel_workspace_bytes(int B, int H3, int W3) {
  if (B <= 0 || H3 < 4 || W3 < 4) return 0;
  Arena a{nullptr, 0};
  PixelScratch s;
  return carve_pixel(a, B, H3, W3, s);
}

extern "C" int nsac_pixel_forward(const nsac_pixel_weights* w, const void* res3_hi, const void* res3_lo, const void* res4_hi,
                                  const void* res4_lo, const void* res5_hi, const void* res5_lo, int B, int H3, int W3,
                                  float* init_tran, float* init_rot, float* pix_tran_feat, float* pix_rot_feat, float* t0,
                                  float* q0, float* rot_feat0, float* trans_feat0, void* workspace, size_t workspace_bytes,
                                  int* launches_out, void* stream) {
  NSAC_REQUIRE(w && init_tran && init_rot && t0 && q0 && rot_feat0 && trans_feat0, "nsac_pixel_forward: null argument");
  NSAC_REQUIRE(B > 0, "nsac_pixel_forward: bad batch size %d", B);
  const bool run_k1 = res5_hi != nullptr;
  if (run_k1) {
    NSAC_REQUIRE(res3_hi && res3_lo && res4_hi && res4_lo && res5_lo, "nsac_pixel_forward: null feature planes");
    NSAC_REQUIRE(H3 % 4 == 0 && W3 % 4 == 0 && H3 >= 4 && W3 >= 4, "nsac_pixel_forward: res3 size %dx%d must be a multiple of 4", H3, W3);
  } else {
    H3 = W3 = 4;     // scratch of the skipped network is still carved (tiny)
  }
  Arena a{static_cast<uint8_t*>(workspace), 0};
  PixelScratch s;
  const size_t need = carve_pixel(a, B, H3, W3, s);
  NSAC_REQUIRE(workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "nsac_pixel_forward: workspace of %zu bytes (256-byte aligned) needed, got %zu", need, workspace_bytes);
  const int fmt = w->fmt, P = w->passes, N = 2 * B, G = w->gn_groups;
  const Planes none{nullptr, nullptr, 0};
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  int n = 0;
  if (run_k1) {
    const int H4 = H3 / 2, W4 = W3 / 2, H5 = H3 / 4, W5 = W3 / 4;
    auto gn = [&](const float* x, int H, int W, int idx, int relu, const float* skip, float* o32, Planes op) {
      ++n;
      return nsac_groupnorm_nhwc(x, N, H, W, 128, G, w->gn_w[idx], w->gn_b[idx], w->gn_eps, relu, skip, fmt, o32, op.hi, op.lo, s.gn_ws,
                                 s.gn_ws_bytes, stream);
    };
    // BasePixelDecoder.forward_features (camera_modules.py:335-348): res5 -> res4 -> res3, top-down
    const Planes p5{const_cast<void*>(res5_hi), const_cast<void*>(res5_lo), 2048}, p4{const_cast<void*>(res4_hi), const_cast<void*>(res4_lo), 1024},
        p3{const_cast<void*>(res3_hi), const_cast<void*>(res3_lo), 512};
    NSAC_TRY(conv3(w->pd_layer_3, p5, N, H5, W5, 2048, NSAC_ACT_NONE, s.t5, none, fmt, P, stream, n));
    NSAC_TRY(gn(s.t5, H5, W5, 0, 1, nullptr, s.y5, none));
    NSAC_TRY(tc(w->pd_adapter_2, p4, N * H4 * W4, NSAC_ACT_NONE, s.t4, 128, none, fmt, P, stream, n));
    NSAC_TRY(gn(s.t4, H4, W4, 1, 0, s.y5, nullptr, s.u4));
    NSAC_TRY(conv3(w->pd_layer_2, s.u4, N, H4, W4, 128, NSAC_ACT_NONE, s.t4, none, fmt, P, stream, n));
    NSAC_TRY(gn(s.t4, H4, W4, 2, 1, nullptr, s.y4, none));
    NSAC_TRY(tc(w->pd_adapter_1, p3, N * H3 * W3, NSAC_ACT_NONE, s.t3, 128, none, fmt, P, stream, n));
    NSAC_TRY(gn(s.t3, H3, W3, 3, 0, s.y4, nullptr, s.u3));
    NSAC_TRY(conv3(w->pd_layer_1, s.u3, N, H3, W3, 128, NSAC_ACT_NONE, s.t3, none, fmt, P, stream, n));
    NSAC_TRY(gn(s.t3, H3, W3, 4, 1, nullptr, nullptr, s.y3));
    NSAC_TRY(conv3(w->pd_mask_features, s.y3, N, H3, W3, 128, NSAC_ACT_NONE, nullptr, s.xa, fmt, P, stream, n));
    // convs_backbone (camera_head.py:78-91): conv-BN-LeakyReLU x2, pool, x2, pool, x2
    int H = H3, W = W3;
    Planes x = s.xa, y = s.xb;
    for (int i = 0; i < 3; ++i) {
      NSAC_TRY(conv3(w->cb[2 * i], x, N, H, W, 256, NSAC_ACT_LEAKY, nullptr, y, fmt, P, stream, n));
      NSAC_TRY(conv3(w->cb[2 * i + 1], y, N, H, W, 256, NSAC_ACT_LEAKY, s.f, none, fmt, P, stream, n));
      if (i < 2) {
        NSAC_TRY(nsac_maxpool2_planes(s.f, N, H, W, 256, fmt, x.hi, x.lo, stream));
        ++n;
        H /= 2; W /= 2;
      }
    }
    // correlation volume + softmax (:652, :1117-1133), then the two regression branches (:655-662)
    const int HW = H * W, Cp = (HW + 63) / 64 * 64;
    NSAC_TRY(nsac_corr_softmax(s.f, s.f + (size_t)B * HW * 256, B, H, W, 256, Cp, fmt, s.aff.hi, s.aff.lo, stream));
    ++n;
    NSAC_REQUIRE(w->ct0.K == 9 * Cp, "nsac_pixel_forward: correlation width %d does not match the packed ct0 weights (K = %d)", Cp, w->ct0.K);
    NSAC_TRY(conv3(w->ct0, s.aff, B, H, W, Cp, NSAC_ACT_LEAKY, s.tc0, none, fmt, P, stream, n));          // [B*HW, 256] = trans | rots
    for (int br = 0; br < 2; ++br) {
      const nsac_tc_layer* L = br == 0 ? w->convs_trans : w->convs_rots;
      NSAC_CUDA(cudaMemcpy2DAsync(s.ysl, 128 * 4, s.tc0 + br * 128, 256 * 4, 128 * 4, (size_t)B * HW, cudaMemcpyDeviceToDevice, cs));
      const float* yin = s.ysl;
      int h = H, ww = W;
      for (int i = 0; i < 5; ++i) {
        const int stride = (i % 2 == 0) ? 2 : 1;          // layers 1..5 of the branch: strides 2,1,2,1,2
        const int ho = (h - 1) / stride + 1, wo = (ww - 1) / stride + 1;
        NSAC_TRY(nsac_im2col3x3_planes(yin, B, h, ww, 128, stride, 1152, fmt, s.cols.hi, s.cols.lo, stream));
        ++n;
        float* yout = (i & 1) ? s.yb : s.ya;
        NSAC_TRY(tc(L[i], s.cols, B * ho * wo, NSAC_ACT_LEAKY, yout, 128, none, fmt, P, stream, n));
        yin = yout; h = ho; ww = wo;
      }
      NSAC_REQUIRE(h * ww * 128 == 768, "nsac_pixel_forward: regression branch ends at %dx%d, fc expects 2x3", h, ww);
      float* feat = br == 0 ? (pix_tran_feat ? pix_tran_feat : s.feat_t) : (pix_rot_feat ? pix_rot_feat : s.feat_r);
      NSAC_TRY(nsac_linear(yin, 768, br == 0 ? w->fc_trans_w : w->fc_rots_w, br == 0 ? w->fc_trans_b : w->fc_rots_b, 0, feat, 256, B, 256,
                           768, NSAC_ACT_RELU, stream));
      ++n;
    }
    float* ft = pix_tran_feat ? pix_tran_feat : s.feat_t;
    float* fr = pix_rot_feat ? pix_rot_feat : s.feat_r;
    NSAC_TRY(nsac_pose_heads(fr, ft, w->rots_w, w->rots_b, w->trans_w, w->trans_b, B, 256, init_rot, init_tran, stream));
    ++n;
  }
  // w >= 0 per pair (the reference flips the whole batch by sample 0, :436-437); t + 1e-10 for the AIM embedding (:718)
  NSAC_TRY(nsac_pose_canon(init_rot, init_tran, B, init_rot, s.teps, stream));
  ++n;
  // K2: AIM (:685-735) - layer 0 (K = 4 / 3) on the CUDA cores, the five 256-wide layers on the tensor-core engine
  for (int br = 0; br < 2; ++br) {
    const int K0 = br == 0 ? 4 : 3;
    NSAC_TRY(nsac_linear(br == 0 ? init_rot : s.teps, K0, br == 0 ? w->rot_emb0_w : w->trans_emb0_w, br == 0 ? w->rot_emb0_b : w->trans_emb0_b,
                         0, s.x32, 256, B, 256, K0, NSAC_ACT_RELU, stream));
    NSAC_TRY(nsac_split16(s.x32, 256, B, 256, 1.0f, fmt, s.e0.hi, s.e0.lo, 256, stream));
    n += 2;
    NSAC_TRY(run_chain(br == 0 ? w->rot_emb : w->trans_emb, 5, s.e0, B, s.e1, s.e2, nullptr, 0, NSAC_ACT_RELU,
                       br == 0 ? rot_feat0 : trans_feat0, 256, none, fmt, P, stream, n));
  }
  NSAC_TRY(nsac_pose_heads(rot_feat0, trans_feat0, w->rots_w, w->rots_b, w->trans_w, w->trans_b, B, 256, q0, t0, stream));
  ++n;
  if (launches_out) *launches_out = n;
  return NSAC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// The whole head: K1 + K2 -> matcher -> refinement (PlaneCameraHead.inference_Joint)
// ---------------------------------------------------------------------------------------------------------------------
namespace {
size_t head_stage_bytes(int B, int H3, int W3, int n1, int n2, int NQ) {
  size_t m = nsac_pixel_workspace_bytes(B, H3 >= 4 ? H3 : 4, W3 >= 4 ? W3 : 4);
  const size_t a = nsac_match_workspace_bytes(B, n1, n2), r = nsac_refine_workspace_bytes(B, NQ);
  if (a > m) m = a;
  if (r > m) m = r;
  return m;
}
}  // namespace

extern "C" size_t nsac_head_workspace_bytes(int B, int H3, int W3, int n1, int n2, int NQ) {
  if (B <= 0 || n1 <= 0 || n2 <= 0 || NQ <= 0) return 0;
  return (((size_t)B * 7 * 4 + 255) & ~size_t(255)) + head_stage_bytes(B, H3, W3, n1, n2, NQ);      // cam rows [B,7] first
}

extern "C" int nsac_head_forward(const nsac_head_weights* w, const void* res3_hi, const void* res3_lo, const void* res4_hi,
                                 const void* res4_lo, const void* res5_hi, const void* res5_lo, int B, int H3, int W3,
                                 const float* planes1, const float* planes2, const float* app1, const float* app2,
                                 const int32_t* count1, const int32_t* count2, int n1, int n2, const int32_t* hyp_pairs, int H,
                                 int NQ, float match_threshold, int out_cam_type, float* init_tran, float* init_rot, float* t0,
                                 float* q0, float* rot_feat0, float* trans_feat0, float* log_scores_padded, float* assign,
                                 float* pose, float* assign_pruned, float* geo_local, float* geo_global, float* sig,
                                 int32_t* matched_num, int32_t* pair_idx, float* q_h, float* t_h, float* score_rot,
                                 float* score_tran, int32_t* sel_idx, void* workspace, size_t workspace_bytes,
                                 float* const* peer_rows, int num_peers, int row_offset, int* launches_out, void* stream) {
  NSAC_REQUIRE(w && w->pixel && w->match && w->refine, "nsac_head_forward: null weights");
  NSAC_REQUIRE(B > 0 && n1 > 0 && n2 > 0 && NQ > 0, "nsac_head_forward: bad sizes B=%d n1=%d n2=%d NQ=%d", B, n1, n2, NQ);
  const size_t cam_bytes = ((size_t)B * 7 * 4 + 255) & ~size_t(255);
  const size_t stage = head_stage_bytes(B, H3, W3, n1, n2, NQ);
  NSAC_REQUIRE(workspace && workspace_bytes >= cam_bytes + stage && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "nsac_head_forward: workspace of %zu bytes (256-byte aligned) needed, got %zu", cam_bytes + stage, workspace_bytes);
  float* cam = static_cast<float*>(workspace);
  void* ws = static_cast<uint8_t*>(workspace) + cam_bytes;
  int n = 0, k = 0;
  NSAC_TRY(nsac_pixel_forward(w->pixel, res3_hi, res3_lo, res4_hi, res4_lo, res5_hi, res5_lo, B, H3, W3, init_tran, init_rot, nullptr,
                              nullptr, t0, q0, rot_feat0, trans_feat0, ws, stage, &k, stream));
  n += k;
  NSAC_TRY(nsac_cam_rows(t0, q0, B, cam, stream));                 // matcher pose = the AIM-refined initial pose (:493)
  ++n;
  NSAC_TRY(nsac_match_forward(w->match, app1, app2, planes1, planes2, cam, count1, count2, match_threshold, B, n1, n2, log_scores_padded,
                              assign, ws, stage, &k, stream));
  n += k;
  NSAC_TRY(nsac_refine_forward(w->refine, planes1, planes2, assign, hyp_pairs, H, t0, q0, rot_feat0, trans_feat0, B, n1, n2, NQ,
                               out_cam_type, pose, assign_pruned, geo_local, geo_global, sig, matched_num, pair_idx, q_h, t_h, score_rot,
                               score_tran, sel_idx, ws, stage, peer_rows, num_peers, row_offset, &k, stream));
  n += k;
  if (launches_out) *launches_out = n;
  return NSAC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// ResNet-50 backbone from uint8 images (nopesac_b200/backbone.py, launch for launch)
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct BackboneScratch {
  void* cols; float* stem_out;
  Planes x[2], y1, y2, sc;       // block input / output ping-pong (256..2048 wide at the level's size), conv1 / conv2 / shortcut outputs
};

inline int half_up(int v) { return (v - 1) / 2 + 1; }

size_t carve_backbone(Arena& a, int N, int H, int W, BackboneScratch& s) {
  const size_t Ho = half_up(H), Wo = half_up(W), r1 = (size_t)N * Ho * Wo;           // stem output
  const size_t H2 = half_up((int)Ho), W2 = half_up((int)Wo), r2 = (size_t)N * H2 * W2;    // res2 level
  // The stem's im2col matrix (r1 x 192 fp16) and its fp32 output (r1 x 64) are dead once the max-pool has run: the block
  // ping-pong buffers (widest at the res2 level: r2 x 256, two planes) live in the same memory.  x[0] (the max-pool's output)
  // overlays the im2col matrix, which the stem GEMM has consumed by then; x[1] overlays the stem output, first written by
  // res2.0's last convolution.
  const size_t xb = r2 * 256 * 2 * 2, cols_b = r1 * 192 * 2, so_b = r1 * 64 * 4;
  uint8_t* reg0 = static_cast<uint8_t*>(a.take(cols_b > xb ? cols_b : xb));
  uint8_t* reg1 = static_cast<uint8_t*>(a.take(so_b > xb ? so_b : xb));
  s.cols = reg0;
  s.stem_out = reinterpret_cast<float*>(reg1);
  s.x[0] = Planes{reg0, reg0 ? reg0 + xb / 2 : nullptr, 256};
  s.x[1] = Planes{reg1, reg1 ? reg1 + xb / 2 : nullptr, 256};
  s.y1 = a.planes(r2, 128);     // conv1 output: 64 @ res2 .. 512 @ res5; res3.0.conv1 runs at the res2 level with 128 channels
  s.y2 = a.planes(r2, 64);      // conv2 output (after the stride)
  s.sc = a.planes(r2, 256);     // projection shortcut
  return a.off;
}

}  // namespace

extern "C" size_t nsac_backbone_workspace_bytes(int N, int H, int W) {
  if (N <= 0 || H < 9 || W < 9) return 0;
  Arena a{nullptr, 0};
  BackboneScratch s;
  return carve_backbone(a, N, H, W, s);
}

extern "C" int nsac_backbone_forward(const nsac_backbone_weights* w, const uint8_t* images, int N, int H, int W, void* res2_hi,
                                     void* res2_lo, void* res3_hi, void* res3_lo, void* res4_hi, void* res4_lo, void* res5_hi,
                                     void* res5_lo, void* workspace, size_t workspace_bytes, int* launches_out, void* stream) {
  NSAC_REQUIRE(w && w->blocks && images, "nsac_backbone_forward: null argument");
  NSAC_REQUIRE(N > 0 && H >= 9 && W >= 9, "nsac_backbone_forward: bad shape N=%d H=%d W=%d (H, W >= 9)", N, H, W);
  NSAC_REQUIRE(w->num_blocks == w->stage_blocks[0] + w->stage_blocks[1] + w->stage_blocks[2] + w->stage_blocks[3],
               "nsac_backbone_forward: stage_blocks do not add up to num_blocks");
  NSAC_REQUIRE((res2_hi == nullptr) == (res2_lo == nullptr) && (res3_hi == nullptr) == (res3_lo == nullptr) &&
               (res4_hi == nullptr) == (res4_lo == nullptr) && (res5_hi == nullptr) == (res5_lo == nullptr),
               "nsac_backbone_forward: hi / lo output planes go together");
  Arena a{static_cast<uint8_t*>(workspace), 0};
  BackboneScratch s;
  const size_t need = carve_backbone(a, N, H, W, s);
  NSAC_REQUIRE(workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "nsac_backbone_forward: workspace of %zu bytes (256-byte aligned) needed, got %zu", need, workspace_bytes);
  const int fmt = w->fmt, P = w->passes;
  const Planes none{nullptr, nullptr, 0};
  int n = 0;
  // stem: raw pixels are exact in fp16 (one plane), normalisation folded into the weights, zero padding of the NORMALISED image
  // through the border-class columns; ReLU in the GEMM epilogue; max-pool 3x3 / 2 writes planes
  int h = half_up(H), ww = half_up(W);
  NSAC_TRY(nsac_stem_im2col_u8_cls(images, N, H, W, s.cols, stream));
  ++n;
  NSAC_TRY(nsac_gemm_split(s.cols, nullptr, 192, w->stem.w_hi, w->stem.w_lo, w->stem.ldw, w->stem.bias, 0, N * h * ww, 64, 192, NSAC_ACT_RELU,
                           P, fmt, 1.0f / w->stem.w_scale, s.stem_out, 64, nullptr, nullptr, 0, stream));
  ++n;
  Planes x = s.x[0];
  x.ld = 64;
  NSAC_TRY(nsac_maxpool3x3s2_nhwc(s.stem_out, N, h, ww, 64, fmt, nullptr, x.hi, x.lo, stream));
  ++n;
  h = half_up(h); ww = half_up(ww);
  void* outs[4][2] = {{res2_hi, res2_lo}, {res3_hi, res3_lo}, {res4_hi, res4_lo}, {res5_hi, res5_lo}};
  int cur = 0, bi = 0;
  for (int st = 0; st < 4; ++st) {
    for (int b = 0; b < w->stage_blocks[st]; ++b, ++bi) {
      const nsac_bottleneck& blk = w->blocks[bi];
      const int cin = blk.conv1.K, mid = blk.conv1.N, cout = blk.conv3.N, stride = blk.stride;
      NSAC_REQUIRE(x.ld == cin && blk.conv2.K == 9 * mid && blk.conv3.K == mid && (stride == 1 || stride == 2),
                   "nsac_backbone_forward: block %d does not chain (input %d channels, conv1 expects %d)", bi, x.ld, cin);
      const int rows_in = N * h * ww, ho = stride == 1 ? h : half_up(h), wo = stride == 1 ? ww : half_up(ww), rows_out = N * ho * wo;
      Planes y1 = s.y1, y2 = s.y2, sc = s.sc;
      y1.ld = mid; y2.ld = mid; sc.ld = cout;
      // the last block of a stage writes the caller's output planes directly
      const bool last = b == w->stage_blocks[st] - 1;
      Planes out = (last && outs[st][0]) ? Planes{outs[st][0], outs[st][1], cout} : s.x[cur ^ 1];
      out.ld = cout;
      NSAC_TRY(tc(blk.conv1, x, rows_in, NSAC_ACT_RELU, nullptr, 0, y1, fmt, P, stream, n));
      NSAC_TRY(nsac_conv3x3_split_strided(y1.hi, y1.lo, blk.conv2.w_hi, blk.conv2.w_lo, blk.conv2.bias, N, h, ww, mid, mid, stride,
                                          NSAC_ACT_RELU, P, fmt, 1.0f / blk.conv2.w_scale, nullptr, 0, y2.hi, y2.lo, y2.ld, stream));
      ++n;
      Planes res = x;
      if (blk.has_shortcut) {
        if (stride == 1) {
          NSAC_TRY(tc(blk.shortcut, x, rows_in, NSAC_ACT_NONE, nullptr, 0, sc, fmt, P, stream, n));
        } else {       // strided projection: the TMA gather skips the pixels a stride-2 1x1 convolution never reads
          NSAC_TRY(nsac_conv1x1_split_strided(x.hi, x.lo, blk.shortcut.w_hi, blk.shortcut.w_lo, blk.shortcut.bias, N, h, ww, cin, cout,
                                              stride, NSAC_ACT_NONE, P, fmt, 1.0f / blk.shortcut.w_scale, nullptr, 0, sc.hi, sc.lo, sc.ld,
                                              stream));
          ++n;
        }
        res = sc;
      } else {
        NSAC_REQUIRE(cin == cout && stride == 1, "nsac_backbone_forward: block %d has no shortcut but changes shape", bi);
      }
      ++n;      // conv3 + shortcut + ReLU in its epilogue
      NSAC_TRY(nsac_gemm_split_residual(y2.hi, y2.lo, y2.ld, blk.conv3.w_hi, blk.conv3.w_lo, blk.conv3.ldw, blk.conv3.bias, rows_out, cout, mid,
                                        NSAC_ACT_RELU, P, fmt, 1.0f / blk.conv3.w_scale, res.hi, res.lo, res.ld, nullptr, 0, out.hi, out.lo,
                                        out.ld, stream));
      x = out;
      if (!(last && outs[st][0])) cur ^= 1;
      h = ho; ww = wo;
    }
  }
  if (launches_out) *launches_out = n;
  return NSAC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Stage set S5 in one call: backbone on both views' uint8 images -> head
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct ModelScratch { Planes r3, r4, r5; void* stage; size_t stage_bytes; };
size_t carve_model(Arena& a, int B, int H, int W, int n1, int n2, int NQ, ModelScratch& s) {
  const size_t N = 2 * (size_t)B, h3 = H / 8, w3 = W / 8;
  s.r3 = a.planes(N * h3 * w3, 512);
  s.r4 = a.planes(N * (h3 / 2) * (w3 / 2), 1024);
  s.r5 = a.planes(N * (h3 / 4) * (w3 / 4), 2048);
  const size_t bb = nsac_backbone_workspace_bytes((int)N, H, W), hd = nsac_head_workspace_bytes(B, (int)h3, (int)w3, n1, n2, NQ);
  s.stage_bytes = bb > hd ? bb : hd;
  s.stage = a.take(s.stage_bytes);
  return a.off;
}
}  // namespace

extern "C" size_t nsac_model_workspace_bytes(int B, int H, int W, int n1, int n2, int NQ) {
  if (B <= 0 || H < 32 || W < 32 || H % 32 || W % 32 || n1 <= 0 || n2 <= 0 || NQ <= 0) return 0;
  Arena a{nullptr, 0};
  ModelScratch s;
  return carve_model(a, B, H, W, n1, n2, NQ, s);
}

extern "C" int nsac_model_forward(const nsac_backbone_weights* bw, const nsac_head_weights* hw, const uint8_t* images, int B, int H,
                                  int W, const float* planes1, const float* planes2, const float* app1, const float* app2,
                                  const int32_t* count1, const int32_t* count2, int n1, int n2, const int32_t* hyp_pairs, int Hn,
                                  int NQ, float match_threshold, int out_cam_type, float* init_tran, float* init_rot, float* t0,
                                  float* q0, float* rot_feat0, float* trans_feat0, float* log_scores_padded, float* assign,
                                  float* pose, float* assign_pruned, float* geo_local, float* geo_global, float* sig,
                                  int32_t* matched_num, int32_t* pair_idx, float* q_h, float* t_h, float* score_rot,
                                  float* score_tran, int32_t* sel_idx, void* workspace, size_t workspace_bytes,
                                  float* const* peer_rows, int num_peers, int row_offset, int* launches_out, void* stream) {
  NSAC_REQUIRE(bw && hw && images, "nsac_model_forward: null argument");
  NSAC_REQUIRE(B > 0 && H >= 32 && W >= 32 && H % 32 == 0 && W % 32 == 0, "nsac_model_forward: image size %dx%d must be a multiple of 32", H, W);
  Arena a{static_cast<uint8_t*>(workspace), 0};
  ModelScratch s;
  const size_t need = carve_model(a, B, H, W, n1, n2, NQ, s);
  NSAC_REQUIRE(workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "nsac_model_forward: workspace of %zu bytes (256-byte aligned) needed, got %zu", need, workspace_bytes);
  int n = 0, k = 0;
  NSAC_TRY(nsac_backbone_forward(bw, images, 2 * B, H, W, nullptr, nullptr, s.r3.hi, s.r3.lo, s.r4.hi, s.r4.lo, s.r5.hi, s.r5.lo, s.stage,
                                 s.stage_bytes, &k, stream));
  n += k;
  NSAC_TRY(nsac_head_forward(hw, s.r3.hi, s.r3.lo, s.r4.hi, s.r4.lo, s.r5.hi, s.r5.lo, B, H / 8, W / 8, planes1, planes2, app1, app2, count1,
                             count2, n1, n2, hyp_pairs, Hn, NQ, match_threshold, out_cam_type, init_tran, init_rot, t0, q0, rot_feat0,
                             trans_feat0, log_scores_padded, assign, pose, assign_pruned, geo_local, geo_global, sig, matched_num, pair_idx,
                             q_h, t_h, score_rot, score_tran, sel_idx, s.stage, s.stage_bytes, peer_rows, num_peers, row_offset, &k, stream));
  n += k;
  if (launches_out) *launches_out = n;
  return NSAC_OK;
}
