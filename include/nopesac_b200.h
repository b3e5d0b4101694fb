/*
 * nopesac_b200 — C ABI of the B200-native (sm_100a) NopeSAC one-plane RANSAC pose path.
 *
 * The reference (IceTTTb/NopeSAC @ 53c69c8) is pure Python/PyTorch and has no FFI of its own; each
 * entry point below replaces the eager-PyTorch block of the reference named in its comment
 * (paths relative to NopeSAC_Net/modeling/).  INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add inside PlaneCameraHead / MatchingHead.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to caller-owned memory (fp32 / int32, row-major, contiguous
 *     unless a leading dimension is given); the library never allocates, frees or retains it;
 *   - `stream` is a cudaStream_t passed as void*; everything is enqueued on it, nothing synchronises;
 *   - return value 0 = success, negative nsac_status otherwise; nsac_last_error() gives the message
 *     (thread-local); no exceptions, no exit();
 *   - quaternions are (w,x,y,z); poses are (t[3], q[4]).
 */
#ifndef NOPESAC_B200_H
#define NOPESAC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSAC_VERSION 100

typedef enum {
  NSAC_OK = 0,
  NSAC_ERR_ARG = -1,      /* bad shape / null pointer / misalignment */
  NSAC_ERR_LAUNCH = -2,   /* CUDA launch / runtime error */
  NSAC_ERR_UNSUPPORTED = -3
} nsac_status;

/* activation codes for nsac_linear / nsac_conv epilogues */
#define NSAC_ACT_NONE 0
#define NSAC_ACT_RELU 1
#define NSAC_ACT_LEAKY 2  /* LeakyReLU(0.01), camera_modules.py:47 */

/* INFERENCE_OUT_CAM_TYPE (camera_head.py:930) */
#define NSAC_CAM_SOFT 0
#define NSAC_CAM_AVG_ALL 1
#define NSAC_CAM_MIN_COST 2
#define NSAC_CAM_MAX_SCORE 3

int nsac_version(void);
const char* nsac_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Dense layer: out[M,N] = act(x[M,K] . w[N,K]^T + bias).  Replaces every nn.Linear / MLP layer on
 * the path (camera_modules.py:226-244; gnn.py:56-67).  `bias` may be NULL.  If bias_group_rows > 0
 * the bias is a matrix [ceil(M/bias_group_rows), N] and row r uses bias row r / bias_group_rows
 * (folds the broadcast initial-pose half of cat[init_feat, geo_feat], camera_head.py:980-986).
 * ldx / ldo are row strides in floats (>= K / >= N).
 * ---------------------------------------------------------------------------------------------- */
int nsac_linear(const float* x, int ldx, const float* w, const float* bias, int bias_group_rows,
                float* out, int ldo, int M, int N, int K, int act, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core dense layer (tcgen05 / TMEM / TMA, sm_100a): same contract as nsac_linear but the operands are
 * 16-bit hi/lo planes of the fp32 matrices (x ~ x_hi + x_lo; row strides lda / ldw in elements, multiples of
 * 8; K % 64 == 0 with zero padding) and the product is accumulated in fp32 as
 * x_hi.w_hi (+ x_lo.w_hi (+ x_hi.w_lo (+ x_lo.w_lo))) for passes = 1 (2 (3 (4))).
 *   fmt NSAC_SPLIT_F16 : fp16 planes, ~2^-22 relative per operand (3 passes ~ fp32); requires |x| <= 65504
 *                        (overflow -> inf -> NaN outputs, never silently wrong)
 *   fmt NSAC_SPLIT_BF16: bf16 planes, fp32 exponent range, ~2^-17 relative per operand
 * out = act(out_scale * acc + bias): out_scale undoes a power-of-two pre-scaling of the weight planes.
 * Outputs: fp32 rows (out_f32, may be NULL) and / or re-split planes for the next layer (out_hi/out_lo, may
 * be NULL).  nsac_split16 produces planes of x*scale from an fp32 matrix (zero-padded to ld_split columns).
 * Replaces the large nn.Linear layers of camera_head.py:957-962, 980-986 and gnn.py:85-93.
 * ---------------------------------------------------------------------------------------------- */
#define NSAC_SPLIT_F16 0
#define NSAC_SPLIT_BF16 1
int nsac_gemm_split(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo, int ldw,
                    const float* bias, int bias_group_rows, int M, int N, int K, int act, int passes, int fmt,
                    float out_scale, float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split,
                    void* stream);
int nsac_split16(const float* x, int ldx, int rows, int K, float scale, int fmt, void* hi, void* lo,
                 int ld_split, void* stream);
/* Sticky fp16-plane overflow flag: *flag_out (HOST int) = 1 if, since the last clearing call, a finite value with |x| > 65504
 * had to be written into an fp16 plane (nsac_split16, nsac_nchw_to_planes, GEMM epilogues) — it became inf there and every
 * result computed from it is invalid.  Synchronises `stream`; clear != 0 resets the flag.  Validation aid, not on the hot path. */
int nsac_plane_overflow(int* flag_out, int clear, void* stream);
/* nsac_gemm_split with a residual input given as hi/lo planes [M, ld_res] (same fmt): out = act(out_scale * acc + bias +
 * (res_hi + res_lo)) — the `out += shortcut; relu(out)` of a bottleneck block (detectron2 BottleneckBlock.forward, the backbone
 * Base.yaml:2-12 builds) inside the producing GEMM's epilogue.  In both entry points a_lo == NULL means "A has no lo plane"
 * (exactly representable in 16 bits): the lo.hi pass and its loads are skipped. */
int nsac_gemm_split_residual(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo, int ldw,
                             const float* bias, int M, int N, int K, int act, int passes, int fmt, float out_scale,
                             const void* res_hi, const void* res_lo, int ld_res, float* out_f32, int ldo, void* out_hi,
                             void* out_lo, int ld_split, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Pixel pose-regression network K1 (camera_head.py:642-683; camera_modules.py:36-48, 246-348) on the tensor-core
 * engine.  Activations are NHWC ([N*H*W, C] rows), carried as fp32 and / or 16-bit hi/lo planes.
 *   nsac_conv3x3_split   3x3 / stride 1 / pad 1 convolution as an IMPLICIT GEMM: the A tiles are gathered from the
 *                        NHWC planes by 4-D TMA (the zero padding is TMA's out-of-bounds fill), weights are planes
 *                        [Cout, 9*Cin] in (ky, kx, cin) order; BatchNorm(eval) is folded into weights + bias by
 *                        the caller; epilogue as in nsac_gemm_split.  Replaces F.conv2d / cuDNN.
 *   nsac_nchw_to_planes  backbone feature maps [N,C,HW] fp32 -> NHWC planes
 *   nsac_groupnorm_nhwc  GroupNorm(G) (+ReLU) (+ nearest-2x-upsampled skip [N,H/2,W/2,C], camera_modules.py:344-347).
 *                        stats_ws: nsac_groupnorm_ws_bytes(N,H,W,G) bytes of scratch for the coalesced 4-channels-per-group
 *                        path (statistics kernel + apply kernel); NULL selects the one-kernel path
 *   nsac_maxpool2_planes MaxPool2d(2,2) of an fp32 NHWC map -> planes
 *   nsac_corr_softmax    compute_corr_softmax (camera_head.py:1117-1133): f1,f2 [B,HW,C] -> planes [B*HW, Cp]
 *                        (channel c2 = w2*H + h2, zero padded to Cp)
 *   nsac_im2col3x3_planes explicit im2col for the small strided convolutions (K order (ky,kx,c), padded to Kp)
 * ---------------------------------------------------------------------------------------------- */
int nsac_conv3x3_split(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                       int N, int H, int W, int Cin, int Cout, int act, int passes, int fmt, float out_scale,
                       float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split, void* stream);
int nsac_nchw_to_planes(const float* x, int N, int C, int HW, int fmt, void* hi, void* lo, void* stream);
size_t nsac_groupnorm_ws_bytes(int N, int H, int W, int G);
int nsac_groupnorm_nhwc(const float* x, int N, int H, int W, int C, int G, const float* gamma, const float* beta,
                        float eps, int relu, const float* skip_half_res, int fmt, float* out_f32, void* out_hi,
                        void* out_lo, void* stats_ws, size_t stats_ws_bytes, void* stream);
int nsac_maxpool2_planes(const float* x, int N, int H, int W, int C, int fmt, void* hi, void* lo, void* stream);
int nsac_corr_softmax(const float* f1, const float* f2, int B, int H, int W, int C, int Cp, int fmt, void* hi,
                      void* lo, void* stream);
int nsac_im2col3x3_planes(const float* x, int N, int H, int W, int C, int stride, int Kp, int fmt, void* hi,
                          void* lo, void* stream);

/* LayerNorm over the last dim C (eps 1e-5) with optional residual: out = (res ? res : 0) + LN(x); optionally also
 * as fp16 hi/lo operand planes (out_hi/out_lo, row stride ld_split) for the tensor-core engine.  gnn.py:90,94-96. */
int nsac_layernorm(const float* x, int ldx, const float* gamma, const float* beta, const float* res,
                   int ldres, float* out, int ldo, void* out_hi, void* out_lo, int ld_split, int rows, int C,
                   void* stream);

/* Multi-head full attention (gnn.py:19-44): q [B,L,H*D], k,v [B,S,H*D] with row strides ldq/ldkv,
 * out [B,L,H*D] fp32 and / or fp16 hi/lo planes; softmax(QK^T / sqrt(D)) V, no masks (inference). D must be 32. */
int nsac_attention(const float* q, int ldq, const float* k, const float* v, int ldkv, float* out,
                   int ldo, void* out_hi, void* out_lo, int ld_split, int B, int L, int S, int H, int D,
                   void* stream);
/* Ragged batches (batch elements with different token counts, padded to S rows each): only the first kv_count[b]
 * (int32 [B], device; clamped to [1,S]; NULL = S) keys / values of batch element b take part in the softmax — the
 * all-valid case of the reference's masks (gnn.py:31-34 `kv_mask`).  Query rows beyond a count produce values nobody reads. */
int nsac_attention_ragged(const float* q, int ldq, const float* k, const float* v, int ldkv, float* out,
                          int ldo, void* out_hi, void* out_lo, int ld_split, int B, int L, int S, int H, int D,
                          const int32_t* kv_count, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Matching tail (matching_head.py:75-99, 113-128, 228-234, 259-306 + camera_modules.py:15-34):
 * geometry penalties from the matcher pose cam[B,7]=(t,q), descriptor similarity desc1.desc2/16,
 * minus offset/offset_mult and angle/normal_mult, dustbin padding with bin_score, `iters` log-domain
 * Sinkhorn iterations, then the mutual-NN + threshold assignment.
 *   desc1 [B,n1,C], desc2 [B,n2,C], planes1 [B,n1,3], planes2 [B,n2,3]
 *   -> log_scores_padded [B,n1+1,n2+1], assign [B,n1,n2] (0/1 floats)
 * ---------------------------------------------------------------------------------------------- */
int nsac_match_sinkhorn_assign(const float* desc1, const float* desc2, const float* planes1,
                               const float* planes2, const float* cam, const float* bin_score,
                               float offset_mult, float normal_mult, int iters, float threshold,
                               int B, int n1, int n2, int C, float* log_scores_padded, float* assign,
                               void* stream);
/* Ragged batches: pair b has count1[b] x count2[b] planes (int32 [B], device; clamped to [1,n]; NULL = n1 / n2) stored
 * in the first rows of the padded [B,n1,..] / [B,n2,..] inputs.  Each pair solves ITS OWN (count1+1) x (count2+1) transport
 * problem (dustbins, marginals and normalisation of the un-padded pair — what the reference computes one pair at a time, and
 * what its masked variant `log_optimal_transport_withMask`, matching_head.py:259-306, yields on the valid block); the result
 * sits in the top-left corner of log_scores_padded[b] ([n1+1, n2+1], the rest -inf) and of assign[b] ([n1, n2], the rest 0). */
int nsac_match_sinkhorn_assign_ragged(const float* desc1, const float* desc2, const float* planes1,
                                      const float* planes2, const float* cam, const float* bin_score,
                                      float offset_mult, float normal_mult, int iters, float threshold,
                                      int B, int n1, int n2, int C, const int32_t* count1,
                                      const int32_t* count2, float* log_scores_padded, float* assign,
                                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * Geo sequences (camera_head.py:1352-1425 called three times at :513-517, :555-569) + the 8-vector of
 * :937-957.  The hypothesis list of pair b is the row-major nonzeros of assign[b] (torch.nonzero
 * order), or — when hyp_pairs != NULL — the explicit list hyp_pairs[H,2] (int32, shared by all pairs).
 *   planes1 [B,n1,3], planes2 [B,n2,3], t0 [B,3], q0 [B,4]
 *   -> geo_local [B,NQ,6], geo_global [B,NQ,6], sig [B,NQ] (+-1), geo8 [B,NQ,8],
 *      matched_num [B] int32, pair_idx [B,NQ,2] int32 (-1 padded)
 * No host synchronisation (the reference syncs on torch.nonzero).
 * ---------------------------------------------------------------------------------------------- */
int nsac_geo_sequence(const float* planes1, const float* planes2, const float* assign,
                      const int32_t* hyp_pairs, int H, const float* t0, const float* q0, int B, int n1,
                      int n2, int NQ, float* geo_local, float* geo_global, float* sig, float* geo8,
                      int32_t* matched_num, int32_t* pair_idx, void* stream);

/* Pose heads on per-hypothesis features (camera_head.py:990, 1018): q = normalize(Wr f + br),
 * t = Wt f + bt for rows [rows,256]; either branch may be NULL. */
int nsac_pose_heads(const float* feat_rot, const float* feat_tran, const float* w_rots,
                    const float* b_rots, const float* w_trans, const float* b_trans, int rows, int C,
                    float* q_out, float* t_out, void* stream);

/* Score-MLP weights (camera_head.py:134-138): MLP(NQ,128,64,3) + Linear(64,1), rot and trans twins. */
typedef struct {
  const float* w1; const float* b1;   /* [128,NQ], [128] */
  const float* w2; const float* b2;   /* [128,128], [128] */
  const float* w3; const float* b3;   /* [64,128], [64] */
  const float* w4; const float* b4;   /* [1,64], [1] */
} nsac_score_mlp;

/* ------------------------------------------------------------------------------------------------
 * Hypothesis scoring + pose selection (camera_head.py:964-1115): for pair b with m = matched_num[b]
 * one-plane hypotheses (index 0 = initial pose q0/t0, 1..m = q_h/t_h), residuals of every hypothesis
 * against the m matched plane pairs, the two score MLPs, softmax over the m+1 hypotheses, and the
 * avg / soft / min-cost / max-score selection.  Per-sample semantics for m == 0 / m == 1 (:964, :1068).
 *   geo_local [B,NQ,6]; q_h [B,NQ,4]; t_h [B,NQ,3]; q0 [B,4]; t0 [B,3];
 *   feat_rot/feat_tran [B,NQ,256] (fused one-plane features); feat_rot0/feat_tran0 [B,256];
 *   -> pose [B,16] = (t[3], q[4], t_avg[3], q_avg[4], matched_num, 0)
 *      optional: score_rot/score_tran [B,NQ+1]; sel_idx [B,2] int32 (-1 unless min-cost/max-score);
 *      diag [3,B,NQ+1,NQ] = (l2_dist, normal_dist deg, offset_dist)
 *   workspace: nsac_score_workspace_bytes(B, NQ) bytes.
 * ---------------------------------------------------------------------------------------------- */
size_t nsac_score_workspace_bytes(int B, int NQ);
int nsac_score_aggregate(const float* geo_local, const float* q_h, const float* t_h, const float* q0,
                         const float* t0, const float* feat_rot, const float* feat_tran,
                         const float* feat_rot0, const float* feat_tran0, const int32_t* matched_num,
                         const nsac_score_mlp* rot_mlp, const nsac_score_mlp* tran_mlp,
                         const float* w_rots, const float* b_rots, const float* w_trans,
                         const float* b_trans, int B, int NQ, int out_cam_type, float* pose,
                         float* score_rot, float* score_tran, int32_t* sel_idx, float* diag,
                         void* workspace, void* stream);

/* Same contract on the tensor pipe (tcgen05 / TMEM / TMA): residuals on the CUDA cores feed the score MLPs as
 * fp16 tcgen05.mma operands (single pass, fp32 accumulation; hypothesis 0 runs through the same tiles, all
 * selections stay exact fp32), softmax + feature aggregation are streamed flash-style.  No diagnostic outputs
 * on this path.
 *   nsac_score_pack builds the device-side weight pack once per weight version (fp16 padded copies of the two
 *   score MLPs, folded last layers) into `pack` (nsac_score_pack_bytes(NQ) bytes);
 *   workspace: nsac_score_tc_workspace_bytes(B, NQ) bytes.  Both buffers 256-byte aligned.
 *   Fused result exchange (multi-GPU): if peer_rows != NULL it is a DEVICE array of num_peers pointers to the
 *   [world*B, 16] result buffers of every rank (NVLink peer mappings, e.g. torch symmetric memory); the kernel
 *   that finishes pair b also stores its row into row (row_offset + b) of each of them, replacing the all-gather
 *   of mp3d_evaluation.py:317-318.  The caller issues one cross-rank barrier before reading. */
size_t nsac_score_pack_bytes(int NQ);
int nsac_score_pack(const nsac_score_mlp* rot_mlp, const nsac_score_mlp* tran_mlp, int NQ, void* pack,
                    void* stream);
size_t nsac_score_tc_workspace_bytes(int B, int NQ);
int nsac_score_aggregate_tc(const float* geo_local, const float* q_h, const float* t_h, const float* q0,
                            const float* t0, const float* feat_rot, const float* feat_tran,
                            const float* feat_rot0, const float* feat_tran0, const int32_t* matched_num,
                            const void* pack, const float* w_rots, const float* b_rots, const float* w_trans,
                            const float* b_trans, int B, int NQ, int out_cam_type, float* pose,
                            float* score_rot, float* score_tran, int32_t* sel_idx, void* workspace,
                            float* const* peer_rows, int num_peers, int row_offset, void* stream);
/* Same call with a HOST mirror of the pack's vectors (`vecs_host`: the 6 x 128 floats at byte offset
 * nsac_score_pack_vecs_offset(NQ) of the pack, copied to the host once after nsac_score_pack): they then travel as kernel
 * parameters (constant bank) instead of shared-memory broadcast loads in the MLP epilogue.  NULL = nsac_score_aggregate_tc. */
size_t nsac_score_pack_vecs_offset(int NQ);
int nsac_score_aggregate_tc_cv(const float* geo_local, const float* q_h, const float* t_h, const float* q0, const float* t0,
                               const float* feat_rot, const float* feat_tran, const float* feat_rot0, const float* feat_tran0,
                               const int32_t* matched_num, const void* pack, const float* vecs_host, const float* w_rots,
                               const float* b_rots, const float* w_trans, const float* b_trans, int B, int NQ, int out_cam_type,
                               float* pose, float* score_rot, float* score_tran, int32_t* sel_idx, void* workspace,
                               float* const* peer_rows, int num_peers, int row_offset, void* stream);
/* Profiling aid: while `buf` (8 x 256 uint64 device words, zeroed by the caller) is set, CTA `cta` of the scoring
 * kernel records %globaltimer at every role hand-off (rows: residual, mma, epilogue, gather; rows 4 / 5: start / end of every CTA,
 * rows 6 / 7: feature producer / consumer per chunk).  NULL = off. */
int nsac_debug_score_trace(void* buf, int cta);
/* Same for the GEMM engine: CTA 0 of every nsac_gemm_split / nsac_conv3x3_split launch writes 9 timestamps into `buf`
 * (16 uint64 device words): entry, setup done, first TMA issued, first operands landed, first chunk committed, epilogue
 * start, epilogue end, after the final barrier, after TMEM dealloc.  NULL = off. */
int nsac_debug_gemm_trace(void* buf);

/* Assignment pruning with the refined pose (camera_head.py:605-629): keep matches whose warped normal
 * angle < 45 deg and offset distance < 1 m.  pose rows are (t[3], q[4], ...) with stride ldpose. */
int nsac_prune_assignment(const float* assign, const float* planes1, const float* planes2,
                          const float* pose, int ldpose, int B, int n1, int n2, float* assign_out,
                          void* stream);

/* Camera-pose evaluation (evaluation/mp3d_evaluation.py:382-425 `_eval_camera_reg`, :463-465
 * `angle_error_vec`): per-pair translation error |t - t_gt|_2 and rotation error 2 acos(clip(|q.q_gt|,-1,1)) 180/pi of the
 * result rows (t[3], q[4], ... with stride ldpose), plus stats[8] = {T mean, R mean, #T<1.0, #T<0.5, #T<0.2, #R<30, #R<15,
 * #R<10}.  Medians need the sorted errors: nopesac_b200/evaluation.py. */
int nsac_camera_errors(const float* pose, int ldpose, const float* gt_tran, const float* gt_rot, int B, float* err_t,
                       float* err_r, float* stats, void* stream);

/* Plane-list extraction from the PlaneTRHead outputs, the step before the path (SURVEY.md row f1:
 * meta_arch/siamese_planeTR.py:625-803 `_postprocess_planeHeadMask`, batched, no host round trip).
 *   pred_logits [B,NQ,2], pred_params [B,NQ,3], mask_logits [B,NQ,h,w], query_feat [B,NQ,C]   (fp32, contiguous)
 *   H x W = output size, exactly 2x or 4x the mask resolution (the reference: 120x160 -> 480x640), NQ <= 127.
 *   thresholds: cfg.TEST.PLANE_SCORE_THRESHOLD / MASK_PROB_THRESHOLD (compared in fp32 like torch does) and OVERLAP_THRESHOLD
 *   (the reference compares a Python float ratio, hence double).
 * Outputs, padded to NQ rows per image, kept planes first in query order:
 *   count [B]; flags [B] (bit 0: no query passed the plane test -> best p0 forced (:657-661); bit 1: no plane passed the
 *   overlap rule -> the un-thresholded region of the max-overlap plane (:741-790); bit 2: forced plane had an empty mask ->
 *   pixel (0,0) set (:699-702)); ori_idx [B,NQ] (-1 = padding) = pred_plane_oriIdxs; planes [B,NQ,3] = pred_plane;
 *   feats [B,NQ,C] = pred_plane_feats; scores [B,NQ]; centers [B,NQ,2] = pred_plane_ins_center; bboxes [B,NQ,4] = (x,y,w,h)
 *   of pycocotools toBbox; areas [B,NQ]; seg [B,H,W] uint8 label map, 255 = no plane, pred_plane_masks[j] == (seg == j).
 * workspace: nsac_plane_post_workspace_bytes(B,NQ,H,W) bytes; workspace and seg 16-byte aligned. */
/* nsac_gemm_split + a per-ROW scalar: out[m, n] = act(out_scale * acc + bias[n] + row_bias[m]) (row_bias [M] fp32, may be NULL). */
int nsac_gemm_split_rowbias(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo, int ldw,
                            const float* bias, const float* row_bias, int M, int N, int K, int act, int passes, int fmt,
                            float out_scale, float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split,
                            void* stream);
/* nsac_conv3x3_split with a convolution stride of 1 or 2 (H, W = INPUT map; output (H-1)/stride+1 x (W-1)/stride+1, rows in that
 * order): the A tiles are gathered by TMA with a traversal stride, no im2col matrix.  The strided 3x3 of res3.0 / res4.0 /
 * res5.0 of the backbone (detectron2 BottleneckBlock with STRIDE_IN_1X1 = False, Base.yaml:2-12). */
int nsac_conv3x3_split_strided(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                               int N, int H, int W, int Cin, int Cout, int stride, int act, int passes, int fmt,
                               float out_scale, float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split,
                               void* stream);
/* 1x1 convolution with a stride of 1 or 2, no padding (weights [Cout, Cin]): the projection shortcut of res3.0 / res4.0 / res5.0
 * reads every second pixel of every second row through the tensor map's traversal stride - no subsampled copy of the input. */
int nsac_conv1x1_split_strided(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                               int N, int H, int W, int Cin, int Cout, int stride, int act, int passes, int fmt,
                               float out_scale, float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split,
                               void* stream);
size_t nsac_plane_post_workspace_bytes(int B, int NQ, int H, int W);
int nsac_plane_postprocess(const float* pred_logits, const float* pred_params, const float* mask_logits,
                           const float* query_feat, int B, int NQ, int C, int h, int w, int H, int W,
                           float plane_score_thr, float mask_prob_thr, double overlap_thr, int32_t* count,
                           int32_t* flags, int32_t* ori_idx, float* planes, float* feats, float* scores,
                           float* centers, float* bboxes, int32_t* areas, uint8_t* seg, void* workspace,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * ResNet-50 backbone glue (detectron2 build_resnet_backbone, Base.yaml:2-12 — plain ResNet, stride in the
 * 3x3, FrozenBN).  The convolutions themselves run on the tensor-core engine (nsac_gemm_split / nsac_conv3x3_split /
 * nsac_im2col3x3_planes); activations are NHWC rows [N*H*W, C].
 *   nsac_stem_im2col_planes  image [N,3,H,W] fp32 NCHW, (x - mean) / std per channel (PIXEL_MEAN / PIXEL_STD; HOST pointers to
 *                            3 floats) fused with the im2col of the 7x7 / stride 2 / pad 3 stem convolution -> planes
 *                            [N*Ho*Wo, 192], Ho = (H-1)/2+1: columns (ky,kx,c), 147 used, the rest zero
 *   nsac_maxpool3x3s2_nhwc   MaxPool2d(3, stride 2, pad 1) of an fp32 NHWC map -> fp32 and / or planes [N*Ho*Wo, C]
 *   nsac_subsample2_planes   every second pixel (even y, even x) of NHWC planes: the input of a stride-2 1x1 convolution
 *   nsac_add_relu_nhwc       relu(a + b) over `count` fp32 elements (count % 4 == 0) -> fp32 and / or planes (bottleneck output)
 * ---------------------------------------------------------------------------------------------- */
int nsac_stem_im2col_planes(const float* img, int N, int H, int W, const float* mean3_host, const float* std3_host,
                            int fmt, void* hi, void* lo, void* stream);
int nsac_maxpool3x3s2_nhwc(const float* x, int N, int H, int W, int C, int fmt, float* out_f32, void* hi, void* lo,
                           void* stream);
int nsac_subsample2_planes(const void* hi, const void* lo, int N, int H, int W, int C, void* out_hi, void* out_lo,
                           void* stream);
int nsac_add_relu_nhwc(const float* a, const float* b, size_t count, int fmt, float* out_f32, void* hi, void* lo,
                       void* stream);
/*   nsac_stem_im2col_u8        image [N,3,H,W] UINT8 (what preprocess_image receives, siamese_planeTR.py:534-542) -> ONE fp16 plane
 *                              [N*Ho*Wo, 192] of raw pixel values (exact; no lo plane), out-of-image taps = 0.  The normalisation is
 *                              folded into the stem weights by the caller (w / std, bias - sum w mean / std).
 *   nsac_stem_border_fix       exact fp32 recomputation (zero padding of the NORMALISED image, + bias, ReLU) of the stem outputs
 *                              [N*Ho*Wo, 64] whose 7x7 window leaves the image; w_folded [64,147] (ky,kx,c) for normalised input
 *   nsac_im2col3x3_from_planes 3x3 / pad 1 / stride 1|2 im2col NHWC planes -> planes [N*Ho*Wo, 9*C] (16-byte copies) */
int nsac_stem_im2col_u8(const uint8_t* img, int N, int H, int W, void* out_hi, void* stream);
/* Same plus 24 one-hot border-class columns (k = 147 .. 170; class = (row class, column class) of the output pixel, each of
 * {first, second, interior, second-to-last, last}, index = 5 * row class + column class, skipping interior x interior): with the
 * per-class padding correction sum_{taps outside the image} w * mean / std in rows 147 + idx of the weight matrix, the stem GEMM
 * gives the zero-padded convolution of the NORMALISED image for every pixel and nsac_stem_border_fix is not needed. H, W >= 9. */
int nsac_stem_im2col_u8_cls(const uint8_t* img, int N, int H, int W, void* out_hi, void* stream);
int nsac_stem_border_fix(const uint8_t* img, const float* w_folded, const float* bias, int N, int H, int W,
                         const float* mean3_host, const float* std3_host, float* out, void* stream);
int nsac_im2col3x3_from_planes(const void* hi, const void* lo, int N, int H, int W, int C, int stride, void* out_hi,
                               void* out_lo, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PlaneTRHead glue (first half: planeTR_net/planeTR_head.py:116-192, transformer/transformer.py) — what is
 * not a GEMM; the linear layers / 1x1 convolutions run on nsac_gemm_split*.
 *   nsac_row_op              per row of [rows, C] fp32 (C % 32 == 0, <= 1024): s = x (+ y);  t = do_ln ? LayerNorm(s; gamma, beta,
 *                            eps) : s;  optional outputs s (sum_out), t (t_out fp32, t_hi/t_lo planes) and t + pos[row % T]
 *                            (p_hi/p_lo planes): the residual adds, LayerNorms and `with_pos_embed` adds of
 *                            transformer.py:170-185 (post-norm encoder) / :284-311 (pre-norm decoder)
 *   nsac_attention_tiled     multi-head attention softmax(q k^T / sqrt(32)) v, head dim 32, S <= 320 keys, any L: q [B*L, H*32]
 *                            (row stride ldq), k / v [B*S, H*32] (row stride ldkv) -> out [B*L, H*32] fp32 and / or planes;
 *                            replaces nn.MultiheadAttention's attention core (transformer.py:153, 238-239)
 *   nsac_upsample2x_relu_add out = relu(bilinear_2x(a)) + b on NHWC fp32 maps (a [N,h,w,C], b / out [N,2h,2w,C],
 *                            align_corners = False): the top-down path of planeTR_head.py:240-252 with the 1x1 convolution +
 *                            BatchNorm applied before the upsampling
 * ---------------------------------------------------------------------------------------------- */
int nsac_row_op(const float* x, int ldx, const float* y, int ldy, const float* gamma, const float* beta, float eps, int do_ln,
                const float* pos, int T, float* sum_out, int ld_sum, float* t_out, int ld_t, void* t_hi, void* t_lo, int ld_tp,
                void* p_hi, void* p_lo, int ld_pp, int rows, int C, void* stream);
int nsac_attention_tiled(const float* q, int ldq, const float* k, const float* v, int ldkv, float* out, int ldo, void* out_hi,
                         void* out_lo, int ld_split, int B, int L, int S, int H, int D, void* stream);
int nsac_upsample2x_relu_add(const float* a, const float* b, int N, int h, int w, int C, float* out, void* out_hi, void* out_lo,
                             void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole-stage entry points (csrc/forward.cu).  The host-side orchestration nopesac_b200/camera_head.py does in Python, behind
 * ONE call per stage, so that a host in another language (or the reference, through a single ctypes stub) does not have to
 * replicate it.  They only enqueue this library's kernels on `stream`: no allocation, no synchronisation, no host round trip;
 * scratch comes from one caller-provided workspace.
 *
 * nsac_refine_forward — the one-plane RANSAC refinement itself, camera_head.py:513-629 + 925-1115:
 *   geo sequences + sig (K6) -> 8-vector geo encoding -> ~28-layer hypothesis MLP chain on the tensor-core engine (K7) -> one
 *   pose hypothesis per matched plane pair -> residual scoring, softmax, soft / avg / min-cost / max-score selection (K8, K9) ->
 *   assignment pruning with the refined pose (K10).
 *   Inputs : planes1 [B,n1,3], planes2 [B,n2,3]; the matcher's assignment [B,n1,n2] (or an explicit hypothesis list hyp_pairs
 *            int32 [H,2], shared by all pairs, H <= NQ — the BASELINE "P planes x H hypotheses" workloads; assign is then only
 *            pruned and may be NULL together with assign_pruned); the initial pose t0 [B,3], q0 [B,4] (w >= 0) and its
 *            256-d features rot_feat0 / trans_feat0 [B,256] (AIM, camera_head.py:685-735).
 *   Outputs: pose [B,16] = t(3) q(4) t_avg(3) q_avg(4) m 0 (the row that is all-gathered; see nsac_score_aggregate_tc for the
 *            fused multi-GPU exchange arguments), assign_pruned [B,n1,n2], geo_local / geo_global [B,NQ,6], sig [B,NQ],
 *            matched_num [B], pair_idx [B,NQ,2], the per-hypothesis poses q_h [B*NQ,4] / t_h [B*NQ,3], softmax scores
 *            score_rot / score_tran [B,NQ+1] (NULL = skip), sel_idx [B,2].
 *   weights: borrowed pointers, built once per weight version by the caller (planes via nsac_split16 of w * w_scale).
 *   *launches_out (may be NULL) = kernels enqueued.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* w_hi; const void* w_lo;   /* weight planes [N, ldw]: nsac_split16 of (w * w_scale), zero padded to K columns */
  const float* bias;                    /* [N] or NULL */
  int N, K, ldw;                        /* K = in-features padded to a multiple of 64 */
  float w_scale;                        /* power of two (1 for bf16 planes) */
} nsac_tc_layer;

typedef struct {
  const float* geo0_w; const float* geo0_b;   /* geo_encoder.layers.0 [1024,8], [1024]: K = 8 runs on the CUDA cores (nsac_linear) */
  nsac_tc_layer geo_encoder[5];               /* geo_encoder.layers.1-5            (camera_head.py:123) */
  nsac_tc_layer geo_proj_s1[3];               /* (:124) */
  nsac_tc_layer decoder_rot[6];               /* (:126) */
  nsac_tc_layer geo_proj_s2[3];               /* (:128)  K = 1280 = cat[s1, rot] */
  nsac_tc_layer decoder_tran[6];              /* (:129) */
  nsac_tc_layer decoder_rot2[3];              /* (:131)  layer 0 = the geo half W[:, 256:512], bias NULL (see rot2_*) */
  nsac_tc_layer decoder_tran2[3];             /* (:132)  same */
  const float* rot2_w_init; const float* rot2_b0;     /* decoder_rot2.layers.0: init-feature half W[:, 0:256] [512,256] + bias [512] */
  const float* tran2_w_init; const float* tran2_b0;   /* decoder_tran2.layers.0, same */
  const float* rots_w; const float* rots_b;           /* shared pose heads (:64-65): [4,256],[4] / [3,256],[3] */
  const float* trans_w; const float* trans_b;
  const void* score_pack;                     /* nsac_score_pack output for this NQ */
  const float* score_vecs_host;               /* HOST copy of the pack's vectors (nsac_score_pack_vecs_offset) or NULL */
  const nsac_score_mlp* rot_mlp;              /* raw fp32 score MLPs: only NSAC_CAM_MAX_SCORE needs them (exact-fp32 scoring) */
  const nsac_score_mlp* tran_mlp;
  int fmt, passes;                            /* NSAC_SPLIT_F16 / _BF16; 3 = fp32-grade */
} nsac_refine_weights;

size_t nsac_refine_workspace_bytes(int B, int NQ);
int nsac_refine_forward(const nsac_refine_weights* w, const float* planes1, const float* planes2, const float* assign,
                        const int32_t* hyp_pairs, int H, const float* t0, const float* q0, const float* rot_feat0,
                        const float* trans_feat0, int B, int n1, int n2, int NQ, int out_cam_type, float* pose,
                        float* assign_pruned, float* geo_local, float* geo_global, float* sig, int32_t* matched_num,
                        int32_t* pair_idx, float* q_h, float* t_h, float* score_rot, float* score_tran, int32_t* sel_idx,
                        void* workspace, size_t workspace_bytes, float* const* peer_rows, int num_peers, int row_offset,
                        int* launches_out, void* stream);

/* nsac_match_forward — the whole MatchingHead forward (matching_head.py:43-133 with transformer/gnn.py:73-138 and the
 * assignment of camera_modules.py:15-34): appearance projection -> 18 attentional GNN layers (self /
 * cross alternating, six linears each on the tensor-core engine, attention + LayerNorm kernels emitting operand planes) ->
 * descriptor projection -> geometry penalties + similarity + Sinkhorn + mutual-NN assignment (nsac_match_sinkhorn_assign).
 *   app1 [B,n1,256], app2 [B,n2,256] plane appearance embeddings; planes1 [B,n1,3], planes2 [B,n2,3]; cam [B,7] = (t, q) of the
 *   matcher pose; count1 / count2 (int32 [B], device) = planes per pair in a ragged padded batch, or both NULL.
 *   -> log_scores_padded [B,n1+1,n2+1], assign [B,n1,n2].
 * 'self' layers apply the same weights to both views: with n1 == n2 they run as ONE batch of 2B elements. */
typedef struct {
  nsac_tc_layer qkv;      /* cat[q_proj, k_proj, v_proj] [768,256], no bias  (self layers) */
  nsac_tc_layer q, kv;    /* q_proj [256,256]; cat[k_proj, v_proj] [512,256]  (cross layers) */
  nsac_tc_layer merge, mlp0, mlp2;      /* [256,256], [512,512], [256,512], no bias */
  const float* n1w; const float* n1b; const float* n2w; const float* n2b;   /* LayerNorm(256) x 2 */
  int self_attn;          /* 1 = 'self', 0 = 'cross' (gnn.py:128-134) */
} nsac_gnn_layer;

typedef struct {
  nsac_tc_layer app_proj, desc_proj;    /* planeApp_proj / planeDesc_proj (Conv1d 256->256, k = 1) with bias */
  const nsac_gnn_layer* layers; int num_layers;
  const float* bin_score;               /* device scalar */
  float offset_multiplier, normal_multiplier;
  int sinkhorn_iterations;              /* 200 */
  int fmt, passes;
} nsac_match_weights;

size_t nsac_match_workspace_bytes(int B, int n1, int n2);
int nsac_match_forward(const nsac_match_weights* w, const float* app1, const float* app2, const float* planes1,
                       const float* planes2, const float* cam, const int32_t* count1, const int32_t* count2,
                       float match_threshold, int B, int n1, int n2, float* log_scores_padded, float* assign, void* workspace,
                       size_t workspace_bytes, int* launches_out, void* stream);

/* Initial-pose hand-off between K1 / K2 and the matcher: q_out = q_in with w >= 0 per pair (camera_head.py:436-437; the reference
 * flips the whole batch by sample 0), t_eps_out = t_in + 1e-10 (:718); cam [B,7] = [t, q] rows (:493).  Either half of
 * nsac_pose_canon may be NULL. */
int nsac_pose_canon(const float* q_in, const float* t_in, int B, float* q_out, float* t_eps_out, void* stream);
int nsac_cam_rows(const float* t, const float* q, int B, float* cam, void* stream);

/* nsac_pixel_forward — K1 + K2: the pixel pose network (camera_head.py:642-683 with camera_modules.py:246-348: top-down pixel
 * decoder res5 -> res3, six conv-BN-LeakyReLU blocks with two max-pools, correlation volume + softmax, two strided regression
 * branches, fc, shared pose heads) on the stacked views, then the w >= 0 canonicalisation and the AIM embedding MLPs
 * (:685-735).  Inputs are the backbone's res3 / res4 / res5 maps of N = 2B images (first views, then second views) as NHWC
 * hi/lo planes [N*H*W, C] (nsac_nchw_to_planes, or the backbone's own output): C = 512 / 1024 / 2048, H4 = H3/2, H5 = H3/4.
 * If res5_hi == NULL the network is skipped and init_tran / init_rot are INPUTS (an externally supplied initial pose).
 *   -> init_tran [B,3], init_rot [B,4] (camera_init, w >= 0), pix_tran_feat / pix_rot_feat [B,256] (NULL = drop),
 *      t0 [B,3], q0 [B,4] (camera_initRec), rot_feat0 / trans_feat0 [B,256]. */
typedef struct {
  nsac_tc_layer pd_layer_3, pd_layer_2, pd_layer_1, pd_mask_features;   /* 3x3 convolutions as [Cout, 9*Cin] planes, (ky,kx,cin) */
  nsac_tc_layer pd_adapter_2, pd_adapter_1;                             /* 1x1 convolutions */
  const float* gn_w[5]; const float* gn_b[5];                           /* GroupNorm of layer_3, adapter_2, layer_2, adapter_1, layer_1 */
  int gn_groups; float gn_eps;
  nsac_tc_layer cb[6];                    /* convs_backbone.{0,1,3,4,6,7}: BatchNorm(eval) folded, bias = folded shift */
  nsac_tc_layer ct0;                      /* first conv of convs_trans | convs_rots on the shared correlation input [256, 9*320] */
  nsac_tc_layer convs_trans[5]; nsac_tc_layer convs_rots[5];            /* layers 1-5 as im2col GEMMs [128, 1152] */
  const float* fc_trans_w; const float* fc_trans_b;                     /* fc weights re-ordered for NHWC flattening [256,768] */
  const float* fc_rots_w; const float* fc_rots_b;
  const float* rot_emb0_w; const float* rot_emb0_b;                     /* rot_emb_proj.layers.0 [256,4]  (CUDA cores) */
  const float* trans_emb0_w; const float* trans_emb0_b;                 /* trans_emb_proj.layers.0 [256,3] */
  nsac_tc_layer rot_emb[5]; nsac_tc_layer trans_emb[5];                 /* layers 1-5 */
  const float* rots_w; const float* rots_b; const float* trans_w; const float* trans_b;
  int fmt, passes;
} nsac_pixel_weights;

size_t nsac_pixel_workspace_bytes(int B, int H3, int W3);
int nsac_pixel_forward(const nsac_pixel_weights* w, const void* res3_hi, const void* res3_lo, const void* res4_hi,
                       const void* res4_lo, const void* res5_hi, const void* res5_lo, int B, int H3, int W3, float* init_tran,
                       float* init_rot, float* pix_tran_feat, float* pix_rot_feat, float* t0, float* q0, float* rot_feat0,
                       float* trans_feat0, void* workspace, size_t workspace_bytes, int* launches_out, void* stream);

/* nsac_head_forward — PlaneCameraHead.inference_Joint (camera_head.py:400-640) with its MatchingHead in ONE call: the three
 * stage entries above chained on `stream` (nsac_pixel_forward -> cam = [t0, q0] -> nsac_match_forward -> nsac_refine_forward).
 * Everything the reference returns from inference_Joint is an output buffer of this call:
 *   camera_init = (init_tran, init_rot), camera_initRec = (t0, q0), log_scores_padded, pred_assignment_beforeRef0 = assign,
 *   pred_assignment(_afterRef0) = assign_pruned, camera_softRef0 / camera = pose[:,0:7], camera_avgRef0 = pose[:,7:14],
 *   camera_onePP = (t_h, q_h) with matched_num, sig_seq = sig, score_soft_rot / _offset = score_rot / score_tran.
 * Feature planes / plane lists / counts / hyp_pairs as in the stage entries; rot_feat0 / trans_feat0 [B,256] are scratch the
 * caller provides (the AIM features, consumed by the refinement). */
typedef struct {
  const nsac_pixel_weights* pixel;
  const nsac_match_weights* match;
  const nsac_refine_weights* refine;
} nsac_head_weights;

size_t nsac_head_workspace_bytes(int B, int H3, int W3, int n1, int n2, int NQ);
int nsac_head_forward(const nsac_head_weights* w, const void* res3_hi, const void* res3_lo, const void* res4_hi,
                      const void* res4_lo, const void* res5_hi, const void* res5_lo, int B, int H3, int W3,
                      const float* planes1, const float* planes2, const float* app1, const float* app2, const int32_t* count1,
                      const int32_t* count2, int n1, int n2, const int32_t* hyp_pairs, int H, int NQ, float match_threshold,
                      int out_cam_type, float* init_tran, float* init_rot, float* t0, float* q0, float* rot_feat0,
                      float* trans_feat0, float* log_scores_padded, float* assign, float* pose, float* assign_pruned,
                      float* geo_local, float* geo_global, float* sig, int32_t* matched_num, int32_t* pair_idx, float* q_h,
                      float* t_h, float* score_rot, float* score_tran, int32_t* sel_idx, void* workspace,
                      size_t workspace_bytes, float* const* peer_rows, int num_peers, int row_offset, int* launches_out,
                      void* stream);

/* nsac_backbone_forward — the ResNet-50 backbone (detectron2 build_resnet_backbone as configured by
 * Base.yaml:2-12) from UINT8 images in one call: stem (7x7/2 patches of the raw pixels + border-class columns -> GEMM ->
 * max-pool 3x3/2) and the 16 bottleneck blocks (conv1 1x1 -> conv2 3x3, stride 1 or 2 through the TMA gather -> [projection
 * shortcut] -> conv3 1x1 with the shortcut + ReLU in its epilogue), activations as NHWC hi/lo planes end to end.
 *   images [N,3,H,W] uint8 (cfg.INPUT.FORMAT order, NOT normalised: PIXEL_MEAN / PIXEL_STD are folded into `stem`), H, W >= 9.
 *   -> res2 / res3 / res4 / res5 planes [N*h*w, C] (C = 256 / 512 / 1024 / 2048; h = H/4 .. H/32 rounded up), any pair NULL = that
 *      level is not kept (scratch is used).  Weights: FrozenBN folded into planes + bias; `stem` is the [64, 171] matrix of
 *      nsac_stem_im2col_u8_cls for THIS image size (w / std, then the 24 border-class corrections). */
typedef struct {
  nsac_tc_layer conv1, conv2, conv3, shortcut;    /* conv2: [mid, 9*mid] in (ky,kx,cin) order; shortcut unused if !has_shortcut */
  int has_shortcut, stride;                       /* stride of conv2 / shortcut (STRIDE_IN_1X1 = False) */
} nsac_bottleneck;

typedef struct {
  nsac_tc_layer stem;
  const nsac_bottleneck* blocks; int num_blocks;  /* 3 + 4 + 6 + 3 */
  int stage_blocks[4];                            /* blocks per stage res2 .. res5 */
  int fmt, passes;
} nsac_backbone_weights;

size_t nsac_backbone_workspace_bytes(int N, int H, int W);
int nsac_backbone_forward(const nsac_backbone_weights* w, const uint8_t* images, int N, int H, int W, void* res2_hi, void* res2_lo,
                          void* res3_hi, void* res3_lo, void* res4_hi, void* res4_lo, void* res5_hi, void* res5_lo,
                          void* workspace, size_t workspace_bytes, int* launches_out, void* stream);

/* nsac_model_forward — stage set S5 in ONE call: nsac_backbone_forward on the 2B uint8 images of B pairs (first views, then
 * second views; H, W multiples of 32) -> res3 / res4 / res5 planes in the workspace -> nsac_head_forward.  Arguments and outputs as
 * in those two entries; plane lists (planes / appearance embeddings / counts) come from the caller. */
size_t nsac_model_workspace_bytes(int B, int H, int W, int n1, int n2, int NQ);
int nsac_model_forward(const nsac_backbone_weights* bw, const nsac_head_weights* hw, const uint8_t* images, int B, int H, int W,
                       const float* planes1, const float* planes2, const float* app1, const float* app2, const int32_t* count1,
                       const int32_t* count2, int n1, int n2, const int32_t* hyp_pairs, int Hn, int NQ, float match_threshold,
                       int out_cam_type, float* init_tran, float* init_rot, float* t0, float* q0, float* rot_feat0,
                       float* trans_feat0, float* log_scores_padded, float* assign, float* pose, float* assign_pruned,
                       float* geo_local, float* geo_global, float* sig, int32_t* matched_num, int32_t* pair_idx, float* q_h,
                       float* t_h, float* score_rot, float* score_tran, int32_t* sel_idx, void* workspace, size_t workspace_bytes,
                       float* const* peer_rows, int num_peers, int row_offset, int* launches_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NOPESAC_B200_H */
