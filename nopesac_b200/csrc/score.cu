// K8 + K9 — hypothesis scoring and pose selection (camera_head.py:964-1115), exact-fp32 CUDA-core
// version.  Two kernels per call:
//   score_tile_kernel    CTA = 32 hypotheses of one pair: plane-alignment residuals of each hypothesis
//                        against the m matched plane pairs (both branches share R_h p_j), exp(-d), then
//                        MLP(NQ,128,64,3)+Linear(64,1) as register-tiled mini-GEMMs -> one logit per
//                        hypothesis and branch.  The [B,NQ+1,NQ,3] temporaries of the reference are never
//                        materialised.
//   score_select_kernel  CTA = one pair: softmax over the m+1 hypotheses, avg / soft / min-cost /
//                        max-score selection, streaming weighted sum of the [m+1,256] one-plane features,
//                        shared `rots` / `trans` heads, quaternion normalisation -> pose[b, 0:16].
// Variable m per pair is handled in-kernel (m == 0, m == 1 and m > 1 paths of :964, :1052, :1068).
#include "common.cuh"

namespace {

constexpr int TH = 32;          // hypotheses per CTA
constexpr int KC = 32;          // residual columns (matched pairs) per chunk
constexpr int LDX = 36;         // [k][h] tiles, padded for float4 reads
constexpr int LDW = 132;        // [k][o] weight tiles
constexpr int HID = 128;
constexpr int TILE_THREADS = 256;
constexpr float kPi = 3.14159265358979323846f;

// Fold Linear(128,64) (last MLP layer, no activation) with Linear(64,1): w34[k] = sum_o w4[o] W3[o][k],
// c34 = w4.b3 + b4.  fold[0:128] rot, fold[128] c_rot, fold[129:257] tran, fold[257] c_tran.
__global__ void score_fold_kernel(nsac_score_mlp r, nsac_score_mlp t, float* __restrict__ fold) {
  const nsac_score_mlp& p = blockIdx.x == 0 ? r : t;
  float* out = fold + blockIdx.x * (HID + 1);
  const int k = threadIdx.x;
  if (k < HID) {
    float s = 0.f;
    for (int o = 0; o < 64; ++o) s = fmaf(p.w4[o], p.w3[o * HID + k], s);
    out[k] = s;
  } else if (k == HID) {
    float s = p.b4[0];
    for (int o = 0; o < 64; ++o) s = fmaf(p.w4[o], p.b3[o], s);
    out[HID] = s;
  }
}

struct TileSmem {
  float Xr[KC][LDX];
  float Xt[KC][LDX];
  float Wr[KC][LDW];
  float Wt[KC][LDW];
  float H1[HID][LDX];
  float cj[KC][12];
  float rs[2][8][TH];
};

// acc[4][4] += X[k][ty*4..] (x) W[k][tx*4..] over one 32-deep chunk
__device__ __forceinline__ void chunk_fma(float (&acc)[4][4], const float (*X)[LDX], const float (*W)[LDW],
                                          int ty, int tx) {
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(&X[k][ty * 4]);
    const float4 w = *reinterpret_cast<const float4*>(&W[k][tx * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
  }
}

// W[o][k0 + kk] (row-major [128, ldw]) -> Ws[kk][o]
__device__ __forceinline__ void load_w_chunk(float (*Ws)[LDW], const float* __restrict__ W, int ldw, int k0,
                                             int kmax, int warp, int lane) {
#pragma unroll
  for (int i = 0; i < HID / 8; ++i) {
    const int o = warp + 8 * i, k = k0 + lane;
    Ws[lane][o] = (k < kmax) ? W[(size_t)o * ldw + k] : 0.f;
  }
}

// layers 2+3(+reg) of one branch for this CTA's 32 hypotheses; acc = layer-1 pre-activations.
__device__ __forceinline__ void mlp_tail(TileSmem& s, float (&acc)[4][4], const nsac_score_mlp& p,
                                         const float* __restrict__ fold, float* __restrict__ logit_out, int h0,
                                         int m, int warp, int lane) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int o = lane * 4 + j;
    const float bo = p.b1[o];
#pragma unroll
    for (int i = 0; i < 4; ++i) s.H1[o][warp * 4 + i] = fmaxf(acc[i][j] + bo, 0.f);
  }
  float acc2[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[i][j] = 0.f;
  for (int k0 = 0; k0 < HID; k0 += KC) {
    __syncthreads();  // H1 complete / previous chunk consumed
    load_w_chunk(s.Wr, p.w2, HID, k0, HID, warp, lane);
    __syncthreads();
    chunk_fma(acc2, reinterpret_cast<const float(*)[LDX]>(&s.H1[k0][0]), s.Wr, warp, lane);
  }
  float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int o = lane * 4 + j;
    const float bo = p.b2[o], wf = fold[o];
#pragma unroll
    for (int i = 0; i < 4; ++i) part[i] = fmaf(fmaxf(acc2[i][j] + bo, 0.f), wf, part[i]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float v = warp_sum(part[i]) + fold[HID];
    const int h = h0 + warp * 4 + i;
    if (lane == 0 && h <= m) logit_out[h] = v;
  }
  __syncthreads();  // H1 / Wr free for the next branch
}

template <bool DIAG>
__global__ void __launch_bounds__(TILE_THREADS)
score_tile_kernel(const float* __restrict__ geo_local, const float* __restrict__ q_h,
                  const float* __restrict__ t_h, const float* __restrict__ q0, const float* __restrict__ t0,
                  const int32_t* __restrict__ matched_num, nsac_score_mlp rot, nsac_score_mlp tran,
                  const float* __restrict__ fold, int B, int NQ, float* __restrict__ logits,
                  float* __restrict__ sums, float* __restrict__ diag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem& s = *reinterpret_cast<TileSmem*>(smem_raw);
  const int b = blockIdx.y, h0 = blockIdx.x * TH;
  const int m = matched_num[b];
  if (m == 0 || h0 > m) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H1n = NQ + 1;

  // this lane's hypothesis (index 0 = initial pose, camera_head.py:991, 1019)
  const int h = h0 + lane;
  const bool hv = h <= m;
  float qw = 1.f, qx = 0.f, qy = 0.f, qz = 0.f, tx_ = 0.f, ty_ = 0.f, tz_ = 0.f;
  if (hv) {
    const float* q = (h == 0) ? q0 + (size_t)b * 4 : q_h + ((size_t)b * NQ + (h - 1)) * 4;
    const float* t = (h == 0) ? t0 + (size_t)b * 3 : t_h + ((size_t)b * NQ + (h - 1)) * 3;
    qw = q[0]; qx = q[1]; qy = q[2]; qz = q[3];
    tx_ = t[0]; ty_ = t[1]; tz_ = t[2];
  }
  const Mat3 R = quat_to_rot(qw, qx, qy, qz);

  float accR[4][4], accT[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) accR[i][j] = accT[i][j] = 0.f;
  float sumR = 0.f, sumT = 0.f;

  const float* gl = geo_local + (size_t)b * NQ * 6;
  for (int k0 = 0; k0 < m; k0 += KC) {
    if (tid < KC) {
      const int j = k0 + tid;
      float c[12] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (j < m) {
        c[0] = gl[j * 6 + 0]; c[1] = gl[j * 6 + 1]; c[2] = gl[j * 6 + 2];
        c[3] = gl[j * 6 + 3]; c[4] = -gl[j * 6 + 4]; c[5] = -gl[j * 6 + 5];   // view-2 flip (:994-995)
        float nx = c[3], ny = c[4], nz = c[5];
        c[9] = normalize3(nx, ny, nz);
        c[6] = nx; c[7] = ny; c[8] = nz;
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) s.cj[tid][i] = c[i];
    }
    load_w_chunk(s.Wr, rot.w1, NQ, k0, NQ, warp, lane);
    load_w_chunk(s.Wt, tran.w1, NQ, k0, NQ, warp, lane);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int jl = warp * 4 + i, j = k0 + jl;
      float xr = 0.f, xt = 0.f;
      if (hv && j < m) {
        const float* c = s.cj[jl];
        // rot branch (:997-1006): warp with t = 0, unit normals, L2 distance
        float ax, ay, az;
        warp_plane(R, 0.f, 0.f, 0.f, c[0], c[1], c[2], ax, ay, az);
        normalize3(ax, ay, az);
        const float dx = ax - c[6], dy = ay - c[7], dz = az - c[8];
        const float dL2 = sqrtf(dx * dx + dy * dy + dz * dz);
        xr = expf(-dL2);
        sumR += dL2;
        // trans branch (:1021-1035): warp with (q_h, t_h), L2 distance of plane params
        float px, py, pz;
        warp_plane(R, tx_, ty_, tz_, c[0], c[1], c[2], px, py, pz);
        const float ex = px - c[3], ey = py - c[4], ez = pz - c[5];
        const float dl2 = sqrtf(ex * ex + ey * ey + ez * ez);
        xt = expf(-dl2);
        sumT += dl2;
        if (DIAG) {
          const size_t plane = (size_t)B * H1n * NQ;
          const size_t at = ((size_t)b * H1n + h) * NQ + j;
          const float cosang = fminf(fmaxf(ax * c[6] + ay * c[7] + az * c[8], -1.f), 1.f);
          float ux = px, uy = py, uz = pz;
          const float off0 = normalize3(ux, uy, uz);
          const float nTn = ux * c[6] + uy * c[7] + uz * c[8];
          diag[at] = dl2;
          diag[plane + at] = acosf(cosang) / kPi * 180.f;
          diag[2 * plane + at] = nTn < 0.f ? fabsf(off0 + c[9]) : fabsf(off0 - c[9]);
        }
      }
      s.Xr[jl][lane] = xr;
      s.Xt[jl][lane] = xt;
    }
    __syncthreads();
    chunk_fma(accR, s.Xr, s.Wr, warp, lane);
    chunk_fma(accT, s.Xt, s.Wt, warp, lane);
    __syncthreads();
  }
  // deterministic row sums over the 8 warps (feeds argmin in 'min-cost', :1090-1093)
  s.rs[0][warp][lane] = sumR;
  s.rs[1][warp][lane] = sumT;
  __syncthreads();
  if (tid < TH && h0 + tid <= m) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += s.rs[0][w][tid]; c += s.rs[1][w][tid]; }
    sums[(size_t)b * H1n + h0 + tid] = a;
    sums[(size_t)B * H1n + (size_t)b * H1n + h0 + tid] = c;
  }
  mlp_tail(s, accR, rot, fold, logits + (size_t)b * H1n, h0, m, warp, lane);
  mlp_tail(s, accT, tran, fold + HID + 1, logits + (size_t)B * H1n + (size_t)b * H1n, h0, m, warp, lane);
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < nw; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// first index of the extreme value over [0, n) (torch argmin/argmax tie order on a 1-D CPU tensor)
__device__ int block_arg_extreme(const float* vals, int n, bool want_max, float* redv, int* redi) {
  float best = want_max ? -INFINITY : INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = vals[i];
    if ((want_max ? v > best : v < best)) { best = v; bi = i; }
  }
  __syncthreads();
  redv[threadIdx.x] = best;
  redi[threadIdx.x] = bi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 1; t < blockDim.x; ++t) {
      const float v = redv[t];
      const int i = redi[t];
      if (i == 0x7fffffff) continue;
      if ((want_max ? v > best : v < best) || (v == best && i < bi)) { best = v; bi = i; }
    }
    redi[0] = bi == 0x7fffffff ? 0 : bi;      // no finite candidate (all NaN): hypothesis 0, never an out-of-range index
  }
  __syncthreads();
  const int r = redi[0];
  __syncthreads();
  return r;
}

constexpr int SEL_THREADS = 256;

__global__ void __launch_bounds__(SEL_THREADS)
score_select_kernel(const float* __restrict__ q_h, const float* __restrict__ t_h, const float* __restrict__ q0,
                    const float* __restrict__ t0, const float* __restrict__ feat_rot,
                    const float* __restrict__ feat_tran, const float* __restrict__ feat_rot0,
                    const float* __restrict__ feat_tran0, const int32_t* __restrict__ matched_num,
                    const float* __restrict__ w_rots, const float* __restrict__ b_rots,
                    const float* __restrict__ w_trans, const float* __restrict__ b_trans,
                    const float* __restrict__ logits, const float* __restrict__ sums, int B, int NQ,
                    int out_cam_type, float* __restrict__ pose, float* __restrict__ score_rot,
                    float* __restrict__ score_tran, int32_t* __restrict__ sel_idx) {
  extern __shared__ float sm[];
  const int C = 256;
  const int b = blockIdx.x, tid = threadIdx.x, H1n = NQ + 1;
  float* sr = sm;              // [NQ+1]
  float* st = sr + H1n;        // [NQ+1]
  float* fe = st + H1n;        // [4][256]: avg_rot, avg_tran, soft_rot, soft_tran
  float* red = fe + 4 * C;     // [SEL_THREADS]
  int* redi = reinterpret_cast<int*>(red + SEL_THREADS);
  float* outv = reinterpret_cast<float*>(redi + SEL_THREADS);  // [14]

  const int m = matched_num[b];
  float* P = pose + (size_t)b * 16;
  if (score_rot) for (int h = tid; h < H1n; h += blockDim.x) score_rot[(size_t)b * H1n + h] = 0.f;
  if (score_tran) for (int h = tid; h < H1n; h += blockDim.x) score_tran[(size_t)b * H1n + h] = 0.f;
  if (sel_idx && tid < 2) sel_idx[b * 2 + tid] = -1;
  if (m == 0) {  // :964-969 — nothing matched: the (re-embedded) initial pose is the answer
    if (tid < 3) { P[tid] = t0[b * 3 + tid]; P[7 + tid] = t0[b * 3 + tid]; }
    if (tid < 4) { P[3 + tid] = q0[b * 4 + tid]; P[10 + tid] = q0[b * 4 + tid]; }
    if (tid == 0) { P[14] = 0.f; P[15] = 0.f; }
    return;
  }
  // softmax over hypotheses 0..m (:1010-1014, :1039-1043)
  const float* lr = logits + (size_t)b * H1n;
  const float* lt = logits + (size_t)B * H1n + (size_t)b * H1n;
  float mr = -INFINITY, mt = -INFINITY;
  for (int h = tid; h <= m; h += blockDim.x) { mr = fmaxf(mr, lr[h]); mt = fmaxf(mt, lt[h]); }
  mr = block_reduce(mr, red, true);
  mt = block_reduce(mt, red, true);
  float er = 0.f, et = 0.f;
  for (int h = tid; h <= m; h += blockDim.x) {
    const float a = expf(lr[h] - mr), c = expf(lt[h] - mt);
    sr[h] = a; st[h] = c;
    er += a; et += c;
  }
  er = block_reduce(er, red, false);
  et = block_reduce(et, red, false);
  for (int h = tid; h <= m; h += blockDim.x) {
    sr[h] = sr[h] / er;
    st[h] = st[h] / et;
    if (score_rot) score_rot[(size_t)b * H1n + h] = sr[h];
    if (score_tran) score_tran[(size_t)b * H1n + h] = st[h];
  }
  __syncthreads();

  int sel_r = -1, sel_t = -1;
  if (m > 1 && out_cam_type == NSAC_CAM_MIN_COST) {
    sel_r = block_arg_extreme(sums + (size_t)b * H1n, m + 1, false, red, redi);
    sel_t = block_arg_extreme(sums + (size_t)B * H1n + (size_t)b * H1n, m + 1, false, red, redi);
  } else if (m > 1 && out_cam_type == NSAC_CAM_MAX_SCORE) {
    sel_r = block_arg_extreme(sr, m + 1, true, red, redi);
    sel_t = block_arg_extreme(st, m + 1, true, red, redi);
  }
  if (sel_idx && tid == 0) { sel_idx[b * 2] = sel_r; sel_idx[b * 2 + 1] = sel_t; }

  // streaming weighted sums of the one-plane features, one channel per thread (:1047-1087)
  {
    const int c = tid;
    const float* fr = feat_rot + (size_t)b * NQ * C + c;
    const float* ft = feat_tran + (size_t)b * NQ * C + c;
    const float f0r = feat_rot0[(size_t)b * C + c], f0t = feat_tran0[(size_t)b * C + c];
    float ar = 0.f, at = 0.f, wr = sr[0] * f0r, wt = st[0] * f0t;
    int hh = 0;
    for (; hh + 4 <= m; hh += 4) {
      float vr[4], vt[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { vr[u] = fr[(size_t)(hh + u) * C]; vt[u] = ft[(size_t)(hh + u) * C]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ar += vr[u]; at += vt[u];
        wr = fmaf(sr[hh + 1 + u], vr[u], wr);
        wt = fmaf(st[hh + 1 + u], vt[u], wt);
      }
    }
    for (; hh < m; ++hh) {
      const float vr = fr[(size_t)hh * C], vt = ft[(size_t)hh * C];
      ar += vr; at += vt;
      wr = fmaf(sr[hh + 1], vr, wr);
      wt = fmaf(st[hh + 1], vt, wt);
    }
    if (m > 1) {  // the initial pose joins the average only when m > 1 (:1052-1063)
      const float w = 1.f / (float)(m + 1);
      ar = (ar + f0r) * w;
      at = (at + f0t) * w;
    }
    fe[c] = ar; fe[C + c] = at; fe[2 * C + c] = wr; fe[3 * C + c] = wt;
  }
  __syncthreads();
  // shared pose heads on the 4 aggregated features: 14 dot products of length 256
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int o = warp; o < 14; o += (blockDim.x >> 5)) {
      const float* f; const float* w; float bias;
      if (o < 4)       { f = fe;         w = w_rots + o * C;         bias = b_rots[o]; }
      else if (o < 7)  { f = fe + C;     w = w_trans + (o - 4) * C;  bias = b_trans[o - 4]; }
      else if (o < 11) { f = fe + 2 * C; w = w_rots + (o - 7) * C;   bias = b_rots[o - 7]; }
      else             { f = fe + 3 * C; w = w_trans + (o - 11) * C; bias = b_trans[o - 11]; }
      float a = 0.f;
      for (int c = lane; c < C; c += 32) a = fmaf(f[c], w[c], a);
      a = warp_sum(a);
      if (lane == 0) outv[o] = a + bias;
    }
  }
  __syncthreads();
  if (tid == 0) {
    float qa[4] = {outv[0], outv[1], outv[2], outv[3]};
    float na = fmaxf(sqrtf(qa[0] * qa[0] + qa[1] * qa[1] + qa[2] * qa[2] + qa[3] * qa[3]), 1e-12f);
    for (int i = 0; i < 4; ++i) qa[i] /= na;
    const float ta[3] = {outv[4], outv[5], outv[6]};
    float qf[4], tf[3];
    if (m <= 1 || out_cam_type == NSAC_CAM_AVG_ALL) {  // :1068-1075, :1077-1079
      for (int i = 0; i < 4; ++i) qf[i] = qa[i];
      for (int i = 0; i < 3; ++i) tf[i] = ta[i];
    } else if (out_cam_type == NSAC_CAM_SOFT) {
      float qs[4] = {outv[7], outv[8], outv[9], outv[10]};
      const float ns = fmaxf(sqrtf(qs[0] * qs[0] + qs[1] * qs[1] + qs[2] * qs[2] + qs[3] * qs[3]), 1e-12f);
      for (int i = 0; i < 4; ++i) qf[i] = qs[i] / ns;
      for (int i = 0; i < 3; ++i) tf[i] = outv[11 + i];
    } else {  // min-cost / max-score: pick hypothesis sel_* (index 0 = initial pose)
      const float* q = sel_r == 0 ? q0 + (size_t)b * 4 : q_h + ((size_t)b * NQ + sel_r - 1) * 4;
      const float* t = sel_t == 0 ? t0 + (size_t)b * 3 : t_h + ((size_t)b * NQ + sel_t - 1) * 3;
      for (int i = 0; i < 4; ++i) qf[i] = q[i];
      for (int i = 0; i < 3; ++i) tf[i] = t[i];
    }
    for (int i = 0; i < 3; ++i) { P[i] = tf[i]; P[7 + i] = ta[i]; }
    for (int i = 0; i < 4; ++i) { P[3 + i] = qf[i]; P[10 + i] = qa[i]; }
    P[14] = (float)m;
    P[15] = 0.f;
  }
}
}  // namespace

extern "C" size_t nsac_score_workspace_bytes(int B, int NQ) {
  if (B < 0 || NQ < 1) return 0;
  const size_t per = (size_t)B * (NQ + 1);
  return sizeof(float) * (4 * per + 2 * (HID + 1) + 8);
}

extern "C" int nsac_score_aggregate(const float* geo_local, const float* q_h, const float* t_h, const float* q0,
                                    const float* t0, const float* feat_rot, const float* feat_tran,
                                    const float* feat_rot0, const float* feat_tran0, const int32_t* matched_num,
                                    const nsac_score_mlp* rot_mlp, const nsac_score_mlp* tran_mlp,
                                    const float* w_rots, const float* b_rots, const float* w_trans,
                                    const float* b_trans, int B, int NQ, int out_cam_type, float* pose,
                                    float* score_rot, float* score_tran, int32_t* sel_idx, float* diag,
                                    void* workspace, void* stream) {
  NSAC_REQUIRE(geo_local && q_h && t_h && q0 && t0 && feat_rot && feat_tran && feat_rot0 && feat_tran0 &&
                   matched_num && rot_mlp && tran_mlp && w_rots && b_rots && w_trans && b_trans && pose && workspace,
               "nsac_score_aggregate: null pointer");
  NSAC_REQUIRE(B >= 0 && NQ >= 1, "nsac_score_aggregate: bad shape B=%d NQ=%d", B, NQ);
  NSAC_REQUIRE(out_cam_type >= 0 && out_cam_type <= 3, "nsac_score_aggregate: bad out_cam_type %d", out_cam_type);
  const nsac_score_mlp* mm[2] = {rot_mlp, tran_mlp};
  for (int i = 0; i < 2; ++i)
    NSAC_REQUIRE(mm[i]->w1 && mm[i]->b1 && mm[i]->w2 && mm[i]->b2 && mm[i]->w3 && mm[i]->b3 && mm[i]->w4 && mm[i]->b4,
                 "nsac_score_aggregate: incomplete score MLP weights");
  if (B == 0) return NSAC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t per = (size_t)B * (NQ + 1);
  float* logits = static_cast<float*>(workspace);
  float* sums = logits + 2 * per;
  float* fold = sums + 2 * per;

  score_fold_kernel<<<2, 160, 0, s>>>(*rot_mlp, *tran_mlp, fold);
  NSAC_CHECK_LAUNCH("score_fold_kernel");

  const size_t tile_smem = sizeof(TileSmem);
  dim3 grid(nsac_cdiv(NQ + 1, TH), B);
  if (diag) {
    NSAC_CUDA(cudaFuncSetAttribute(score_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
    score_tile_kernel<true><<<grid, TILE_THREADS, tile_smem, s>>>(geo_local, q_h, t_h, q0, t0, matched_num, *rot_mlp,
                                                                   *tran_mlp, fold, B, NQ, logits, sums, diag);
  } else {
    NSAC_CUDA(cudaFuncSetAttribute(score_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
    score_tile_kernel<false><<<grid, TILE_THREADS, tile_smem, s>>>(geo_local, q_h, t_h, q0, t0, matched_num, *rot_mlp,
                                                                    *tran_mlp, fold, B, NQ, logits, sums, nullptr);
  }
  NSAC_CHECK_LAUNCH("score_tile_kernel");

  const size_t sel_smem = sizeof(float) * (2 * (size_t)(NQ + 1) + 4 * 256 + SEL_THREADS + 16) + sizeof(int) * SEL_THREADS;
  NSAC_REQUIRE(sel_smem <= 200 * 1024, "nsac_score_aggregate: NQ=%d too large", NQ);
  if (sel_smem > 48 * 1024)
    NSAC_CUDA(cudaFuncSetAttribute(score_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
  score_select_kernel<<<B, SEL_THREADS, sel_smem, s>>>(q_h, t_h, q0, t0, feat_rot, feat_tran, feat_rot0, feat_tran0,
                                                       matched_num, w_rots, b_rots, w_trans, b_trans, logits, sums, B,
                                                       NQ, out_cam_type, pose, score_rot, score_tran, sel_idx);
  NSAC_CHECK_LAUNCH("score_select_kernel");
  return NSAC_OK;
}
