// Plane-list extraction from the PlaneTRHead outputs — the step right before the hot path (SURVEY.md §8 row f1:
// PlaneTR_NopeSAC._postprocess_planeHeadMask, meta_arch/siamese_planeTR.py:625-803).  The reference walks the queries of ONE
// image in Python (`.cpu().numpy()`, `.item()`, RLE encode per plane); here a batch of images goes through four launches and
// nothing returns to the host:
//   plane_select_kernel    softmax over the 2 class logits, plane test (label 0 and score > PLANE_SCORE_THRESHOLD), ordered
//                          compaction of the surviving queries, `zero_flag` fallback to the best p0 (:652-661)
//   plane_argmax_kernel    the S-fold bilinear upsampling of sigmoid(mask logits) (align_corners = False, :646-647) is never
//                          materialised: a thread owns one low-resolution CELL (the S x S output pixels that share four
//                          corner samples), walks the surviving queries once (4 loads + 4 sigmoids per query), keeps the
//                          running arg-max of score * prob per pixel (:667, :674, first maximum wins) and counts
//                          `prob >= MASK_PROB_THRESHOLD` (:689).  Per plane it then accumulates area, exact fixed-point
//                          centre sums and the bounding box of its thresholded mask (:685) and of its un-thresholded region
//                          (the `len(instances) == 0` branch, :741-790): per-thread runs -> shared-memory atomics -> one
//                          global atomic per touched slot and CTA.
//   plane_finalize_kernel  the overlap rule (:691-698), the fallback plane, zero-flag pixel patch (:699-702), centres
//                          (:726-739), boxes (pycocotools rleToBbox = tight box), gathers of params / query features.
//   plane_seg_kernel       label map [B,H,W] uint8: index into the kept list or 255.  The reference's masks are disjoint by
//                          construction (arg-max), so `pred_plane_masks[j] == (seg == j)`.
// HBM-bound byte work: algorithmic bytes per image = 4 NQ h w (mask logits, read once; the second corner row / column hits
// L1) + 2 H W (raw + final label map) + H W (re-read); no tensor cores.
#include "common.cuh"

namespace {
constexpr int PL_MAXQ = 127;      // raw label map: 7 bits of valid-list index + 1 bit "below the mask threshold"
constexpr int PL_SLOTS = 16;
enum PlaneSlot {
  S_AREA = 0, S_ORIG, S_XS, S_YS, S_XMIN, S_XMAX, S_YMIN, S_YMAX,        // thresholded mask
  S_AREA_A, S_XS_A, S_YS_A, S_XMIN_A, S_XMAX_A, S_YMIN_A, S_YMAX_A,      // un-thresholded arg-max region
  S_PAD
};
__host__ __device__ __forceinline__ bool is_min_slot(int s) { return s == S_XMIN || s == S_YMIN || s == S_XMIN_A || s == S_YMIN_A; }
__host__ __device__ __forceinline__ bool is_max_slot(int s) { return s == S_XMAX || s == S_YMAX || s == S_XMAX_A || s == S_YMAX_A; }
typedef unsigned long long u64;
constexpr u64 U64_MAX = ~0ull;
constexpr double FIX_SCALE = 1099511627776.0;   // 2^40: float32(k / n) for n <= 65536 is an integer multiple of 2^-40

constexpr int FLAG_ZERO = 1, FLAG_FALLBACK = 2, FLAG_PATCH00 = 4;

struct PlaneWorkspace {
  int32_t* nvalid;      // [B]
  int32_t* valid_q;     // [B,NQ]
  float* valid_score;   // [B,NQ]
  u64* stats;           // [B,NQ,16]
  uint8_t* remap;       // [B,128]
  uint8_t* raw;         // [B,H,W]
  u64* xfix;            // [W]  float32(x / W) * 2^40  (exact: the fixed-point centre sums of :808-809)
  u64* yfix;            // [H]
};

__host__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ inline size_t carve(PlaneWorkspace& ws, void* base, int B, int NQ, int H, int W) {
  size_t off = 0;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align256(bytes); return r; };
  ws.nvalid = reinterpret_cast<int32_t*>(take((size_t)B * 4));
  ws.valid_q = reinterpret_cast<int32_t*>(take((size_t)B * NQ * 4));
  ws.valid_score = reinterpret_cast<float*>(take((size_t)B * NQ * 4));
  ws.stats = reinterpret_cast<u64*>(take((size_t)B * NQ * PL_SLOTS * 8));
  ws.remap = reinterpret_cast<uint8_t*>(take((size_t)B * 128));
  ws.raw = reinterpret_cast<uint8_t*>(take((size_t)B * H * W));
  ws.xfix = reinterpret_cast<u64*>(take((size_t)W * 8));
  ws.yfix = reinterpret_cast<u64*>(take((size_t)H * 8));
  return off;
}

// ------------------------------------------------------------------------------------------------------------------
// 1. query selection (one CTA of 128 threads per image)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
plane_select_kernel(const float* __restrict__ logits, int NQ, float score_thr, PlaneWorkspace ws, int32_t* __restrict__ flags, int H,
                    int W) {
  __shared__ float s_p0[128];
  __shared__ int s_warp_cnt[4];
  const int b = blockIdx.x, q = threadIdx.x, lane = q & 31, wid = q >> 5;
  // fixed-point coordinate tables float32(x / W), float32(y / H) as integer multiples of 2^-40 (:808-809): once per call, so
  // the arg-max kernel's threads do no fp64 division (8 per thread before: the kernel was instruction bound)
  for (int i = b * 128 + q; i < W + H; i += gridDim.x * 128) {
    if (i < W) ws.xfix[i] = (u64)((double)__double2float_rn((double)i / (double)W) * FIX_SCALE);
    else ws.yfix[i - W] = (u64)((double)__double2float_rn((double)(i - W) / (double)H) * FIX_SCALE);
  }
  bool pass = false;
  float score = 0.f, p0 = -1.f;
  if (q < NQ) {
    const float l0 = logits[((size_t)b * NQ + q) * 2 + 0], l1 = logits[((size_t)b * NQ + q) * 2 + 1];
    const float m = fmaxf(l0, l1);
    const float e0 = expf(l0 - m), e1 = expf(l1 - m), s = e0 + e1;
    p0 = e0 / s;
    const float p1 = e1 / s;
    score = fmaxf(p0, p1);
    pass = (p0 >= p1) && (score > score_thr);          // labels == 0 (first maximum) & score > threshold, :653-654
  }
  s_p0[q] = p0;
  const unsigned bal = __ballot_sync(NSAC_FULL_MASK, pass);
  if (lane == 0) s_warp_cnt[wid] = __popc(bal);
  for (int i = q; i < NQ * PL_SLOTS; i += 128)          // stats of every possible list slot of this image
    ws.stats[(size_t)b * NQ * PL_SLOTS + i] = is_min_slot(i & (PL_SLOTS - 1)) ? U64_MAX : 0ull;
  __syncthreads();
  int before = 0, total = 0;
  for (int i = 0; i < 4; ++i) {
    before += (i < wid) ? s_warp_cnt[i] : 0;
    total += s_warp_cnt[i];
  }
  if (pass) {
    const int k = before + __popc(bal & ((1u << lane) - 1u));
    ws.valid_q[(size_t)b * NQ + k] = q;
    ws.valid_score[(size_t)b * NQ + k] = score;
  }
  if (q == 0) {
    int flag = 0;
    if (total == 0) {                                   // :657-661: the query with the highest p0 (first maximum)
      int best = 0;
      for (int i = 1; i < NQ; ++i)
        if (s_p0[i] > s_p0[best]) best = i;
      ws.valid_q[(size_t)b * NQ] = best;
      ws.valid_score[(size_t)b * NQ] = s_p0[best];
      total = 1;
      flag = FLAG_ZERO;
    }
    ws.nvalid[b] = total;
    flags[b] = flag;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// 2. fused sigmoid + bilinear upsampling + weighted arg-max + per-plane statistics
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.f / (1.f + expf(-x)); }

// at::native area_pixel_compute_source_index + guard_index_and_lambda for align_corners = False and scale = 1 / S.
template <int S>
__device__ __forceinline__ void source_lambda(int dst, float& l0, float& l1) {
  float src = (1.0f / S) * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  const int i0 = (int)src;
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

struct RunAcc {
  unsigned n, na;
  u64 xs, ys, xsa, ysa;
  int xmin, xmax, ymin, ymax, xmina, xmaxa, ymina, ymaxa;
  __device__ __forceinline__ void reset() {
    n = na = 0; xs = ys = xsa = ysa = 0;
    xmin = ymin = xmina = ymina = 0x7fffffff;
    xmax = ymax = xmaxa = ymaxa = -1;
  }
};

// 64-bit sum over the warp from three 20-bit limbs (each limb sum < 2^25 fits redux.sync's 32-bit adds; values < 2^60)
__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
  const unsigned s0 = __reduce_add_sync(NSAC_FULL_MASK, (unsigned)(v & 0xFFFFFu));
  const unsigned s1 = __reduce_add_sync(NSAC_FULL_MASK, (unsigned)((v >> 20) & 0xFFFFFu));
  const unsigned s2 = __reduce_add_sync(NSAC_FULL_MASK, (unsigned)(v >> 40));
  return (u64)s0 + ((u64)s1 << 20) + ((u64)s2 << 40);
}

// All lanes of the warp hold a run of the SAME plane (or nothing): reduce with redux.sync / shuffles and let lane 0 issue the 14
// shared-memory atomics once.  (r2k capture: 256 threads x 14 contended 64-bit shared atomics per CTA - the min / max ones are CAS
// loops - made this kernel 1.8 ms per 128 images, 10x its byte work.)
__device__ __forceinline__ void flush_run_warp(u64* st, const RunAcc& a, int lane) {
  RunAcc r;
  r.na = __reduce_add_sync(NSAC_FULL_MASK, a.na);
  r.n = __reduce_add_sync(NSAC_FULL_MASK, a.n);
  r.xsa = warp_sum_u64(a.xsa); r.ysa = warp_sum_u64(a.ysa);
  r.xs = warp_sum_u64(a.xs); r.ys = warp_sum_u64(a.ys);
  r.xmina = __reduce_min_sync(NSAC_FULL_MASK, a.xmina); r.xmaxa = __reduce_max_sync(NSAC_FULL_MASK, a.xmaxa);
  r.ymina = __reduce_min_sync(NSAC_FULL_MASK, a.ymina); r.ymaxa = __reduce_max_sync(NSAC_FULL_MASK, a.ymaxa);
  r.xmin = __reduce_min_sync(NSAC_FULL_MASK, a.xmin); r.xmax = __reduce_max_sync(NSAC_FULL_MASK, a.xmax);
  r.ymin = __reduce_min_sync(NSAC_FULL_MASK, a.ymin); r.ymax = __reduce_max_sync(NSAC_FULL_MASK, a.ymax);
  if (lane != 0 || r.na == 0) return;
  atomicAdd(&st[S_AREA_A], (u64)r.na);
  atomicAdd(&st[S_XS_A], r.xsa);
  atomicAdd(&st[S_YS_A], r.ysa);
  atomicMin(&st[S_XMIN_A], (u64)r.xmina);
  atomicMax(&st[S_XMAX_A], (u64)r.xmaxa);
  atomicMin(&st[S_YMIN_A], (u64)r.ymina);
  atomicMax(&st[S_YMAX_A], (u64)r.ymaxa);
  if (r.n == 0) return;
  atomicAdd(&st[S_AREA], (u64)r.n);
  atomicAdd(&st[S_XS], r.xs);
  atomicAdd(&st[S_YS], r.ys);
  atomicMin(&st[S_XMIN], (u64)r.xmin);
  atomicMax(&st[S_XMAX], (u64)r.xmax);
  atomicMin(&st[S_YMIN], (u64)r.ymin);
  atomicMax(&st[S_YMAX], (u64)r.ymax);
}

__device__ __forceinline__ void flush_run(u64* st, const RunAcc& a) {
  if (a.na == 0) return;
  atomicAdd(&st[S_AREA_A], (u64)a.na);
  atomicAdd(&st[S_XS_A], a.xsa);
  atomicAdd(&st[S_YS_A], a.ysa);
  atomicMin(&st[S_XMIN_A], (u64)a.xmina);
  atomicMax(&st[S_XMAX_A], (u64)a.xmaxa);
  atomicMin(&st[S_YMIN_A], (u64)a.ymina);
  atomicMax(&st[S_YMAX_A], (u64)a.ymaxa);
  if (a.n == 0) return;
  atomicAdd(&st[S_AREA], (u64)a.n);
  atomicAdd(&st[S_XS], a.xs);
  atomicAdd(&st[S_YS], a.ys);
  atomicMin(&st[S_XMIN], (u64)a.xmin);
  atomicMax(&st[S_XMAX], (u64)a.xmax);
  atomicMin(&st[S_YMIN], (u64)a.ymin);
  atomicMax(&st[S_YMAX], (u64)a.ymax);
}

template <int S>
__global__ void __launch_bounds__(256)
plane_argmax_kernel(const float* __restrict__ mask_logits, int NQ, int h, int w, float mask_thr, PlaneWorkspace ws) {
  __shared__ u64 s_st[PL_MAXQ * PL_SLOTS];
  __shared__ int s_q[PL_MAXQ];
  __shared__ float s_sc[PL_MAXQ];
  const int b = blockIdx.z;
  const int nv = ws.nvalid[b];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int i = tid; i < nv * PL_SLOTS; i += 256) s_st[i] = is_min_slot(i & (PL_SLOTS - 1)) ? U64_MAX : 0ull;
  for (int i = tid; i < nv; i += 256) {
    s_q[i] = ws.valid_q[(size_t)b * NQ + i];
    s_sc[i] = ws.valid_score[(size_t)b * NQ + i];
  }
  __syncthreads();

  const int H = h * S, W = w * S;
  const int cj = blockIdx.x * 32 + threadIdx.x - 1;       // cell column in [-1, w-1]
  const int ci = blockIdx.y * 8 + threadIdx.y - 1;        // cell row    in [-1, h-1]
  const bool cell_ok = (cj <= w - 1) && (ci <= h - 1);
  const int r0 = max(ci, 0), r1 = min(ci + 1, h - 1), c0 = max(cj, 0), c1 = min(cj + 1, w - 1);
  const int y0 = S * ci + S / 2, x0 = S * cj + S / 2;     // first output pixel of the cell (may be < 0 / beyond the image)

  float lh0[S], lh1[S], lw0[S], lw1[S];
  bool oky[S], okx[S];
#pragma unroll
  for (int a = 0; a < S; ++a) {
    const int y = y0 + a, x = x0 + a;
    oky[a] = cell_ok && y >= 0 && y < H;
    okx[a] = cell_ok && x >= 0 && x < W;
    source_lambda<S>(max(y, 0), lh0[a], lh1[a]);
    source_lambda<S>(max(x, 0), lw0[a], lw1[a]);
  }

  float best[S][S];
  unsigned char bid[S][S];
#pragma unroll
  for (int a = 0; a < S; ++a)
#pragma unroll
    for (int c = 0; c < S; ++c) {
      best[a][c] = -INFINITY;
      bid[a][c] = 0;
    }

  const size_t plane_stride = (size_t)h * w;
  const float* img = mask_logits + (size_t)b * NQ * plane_stride;
  for (int k = 0; k < nv; ++k) {
    const float sc = s_sc[k];
    unsigned cnt = 0;
    if (cell_ok) {
      const float* L = img + (size_t)s_q[k] * plane_stride;
      const float v00 = sigmoidf_ref(__ldg(L + r0 * w + c0)), v01 = sigmoidf_ref(__ldg(L + r0 * w + c1));
      const float v10 = sigmoidf_ref(__ldg(L + r1 * w + c0)), v11 = sigmoidf_ref(__ldg(L + r1 * w + c1));
      float t0[S], t1[S];
#pragma unroll
      for (int c = 0; c < S; ++c) {
        t0[c] = __fmaf_rn(lw0[c], v00, __fmul_rn(lw1[c], v01));      // association of ATen's CPU kernel (bit-identical to
        t1[c] = __fmaf_rn(lw0[c], v10, __fmul_rn(lw1[c], v11));      // F.interpolate on the host for equal corner values)
      }
#pragma unroll
      for (int a = 0; a < S; ++a)
#pragma unroll
        for (int c = 0; c < S; ++c) {
          const float val = __fmaf_rn(lh0[a], t0[c], __fmul_rn(lh1[a], t1[c]));
          const bool ok = oky[a] && okx[c];
          cnt += (ok && val >= mask_thr) ? 1u : 0u;
          const float wv = __fmul_rn(sc, val);
          if (ok && wv > best[a][c]) {
            best[a][c] = wv;
            bid[a][c] = (unsigned char)k;
          }
        }
    }
    cnt = __reduce_add_sync(NSAC_FULL_MASK, cnt);
    if (threadIdx.x == 0 && cnt) atomicAdd(&s_st[k * PL_SLOTS + S_ORIG], (u64)cnt);
  }

  RunAcc acc;
  acc.reset();
  int cur = -1;
  if (cell_ok) {
    u64 xfix[S], yfix[S];
#pragma unroll
    for (int a = 0; a < S; ++a) {
      xfix[a] = ws.xfix[min(max(x0 + a, 0), W - 1)];   // float32(x / W) * 2^40, :808 (table written by plane_select_kernel)
      yfix[a] = ws.yfix[min(max(y0 + a, 0), H - 1)];   // float32(y / H) * 2^40, :809
    }
    uint8_t* raw = ws.raw + (size_t)b * H * W;
#pragma unroll
    for (int a = 0; a < S; ++a)
#pragma unroll
      for (int c = 0; c < S; ++c) {
        if (oky[a] && okx[c]) {
          const int y = y0 + a, x = x0 + c;
          const int id = bid[a][c];
          const bool in = best[a][c] > mask_thr;                         // :685
          raw[(size_t)y * W + x] = (uint8_t)(id | (in ? 0 : 0x80));
          if (id != cur) {
            if (cur >= 0) flush_run(&s_st[cur * PL_SLOTS], acc);
            acc.reset();
            cur = id;
          }
          acc.na += 1; acc.xsa += xfix[c]; acc.ysa += yfix[a];
          acc.xmina = min(acc.xmina, x); acc.xmaxa = max(acc.xmaxa, x);
          acc.ymina = min(acc.ymina, y); acc.ymaxa = max(acc.ymaxa, y);
          if (in) {
            acc.n += 1; acc.xs += xfix[c]; acc.ys += yfix[a];
            acc.xmin = min(acc.xmin, x); acc.xmax = max(acc.xmax, x);
            acc.ymin = min(acc.ymin, y); acc.ymax = max(acc.ymax, y);
          }
        }
      }
  }
  // the thread's last run (in planar regions its ONLY run): warp-aggregated when the whole warp sits on one plane
  {
    const int lead = __reduce_max_sync(NSAC_FULL_MASK, cur);
    if (lead >= 0) {
      if (__all_sync(NSAC_FULL_MASK, cur < 0 || cur == lead)) flush_run_warp(&s_st[lead * PL_SLOTS], acc, threadIdx.x);
      else if (cur >= 0) flush_run(&s_st[cur * PL_SLOTS], acc);
    }
  }
  __syncthreads();
  u64* gst = ws.stats + (size_t)b * NQ * PL_SLOTS;
  for (int i = tid; i < nv * PL_SLOTS; i += 256) {
    const int slot = i & (PL_SLOTS - 1);
    const u64 v = s_st[i];
    if (is_min_slot(slot)) {
      if (v != U64_MAX) atomicMin(&gst[i], v);
    } else if (is_max_slot(slot)) {
      if (v != 0ull) atomicMax(&gst[i], v);
    } else if (v != 0ull) {
      atomicAdd(&gst[i], v);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// 3. keep / drop decisions and the per-plane outputs (one CTA of 128 threads per image)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
plane_finalize_kernel(const float* __restrict__ params, const float* __restrict__ query_feat, int NQ, int C, double overlap_thr,
                      PlaneWorkspace ws, int32_t* __restrict__ count, int32_t* __restrict__ flags, int32_t* __restrict__ ori_idx,
                      float* __restrict__ planes, float* __restrict__ feats, float* __restrict__ scores, float* __restrict__ centers,
                      float* __restrict__ bboxes, int32_t* __restrict__ areas) {
  __shared__ int s_keep[PL_MAXQ];
  __shared__ int s_n, s_flag;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int nv = ws.nvalid[b];
  const u64* st = ws.stats + (size_t)b * NQ * PL_SLOTS;
  if (tid == 0) {
    int flag = flags[b], n = 0;
    if (flag & FLAG_ZERO) {                               // single forced plane: always kept (:699-702)
      s_keep[n++] = 0;
      if (st[S_AREA] == 0ull) flag |= FLAG_PATCH00;
    } else {
      int max_id = 0;
      double max_ov = 0.0;
      for (int pi = 0; pi < nv; ++pi) {                   // :684-698
        const u64 area = st[pi * PL_SLOTS + S_AREA], orig = st[pi * PL_SLOTS + S_ORIG];
        if (area < 1ull || orig < 1ull) continue;
        const double ov = (double)area / (double)orig;
        if (ov > max_ov) {
          max_ov = ov;
          max_id = pi;
        }
        if (ov < overlap_thr) continue;
        s_keep[n++] = pi;
      }
      if (n == 0) {                                       // :741
        s_keep[n++] = max_id;
        flag |= FLAG_FALLBACK;
      }
    }
    s_n = n;
    s_flag = flag;
    count[b] = n;
    flags[b] = flag;
  }
  __syncthreads();
  const int n = s_n, flag = s_flag;
  if (tid < 128) ws.remap[(size_t)b * 128 + tid] = 255;
  __syncthreads();
  for (int j = tid; j < NQ; j += 128) {
    const size_t o = (size_t)b * NQ + j;
    if (j < n) {
      const int pi = s_keep[j];
      const int q = ws.valid_q[(size_t)b * NQ + pi];
      const u64* s = st + pi * PL_SLOTS;
      ws.remap[(size_t)b * 128 + pi] = (uint8_t)j;
      ori_idx[o] = q;
      scores[o] = ws.valid_score[(size_t)b * NQ + pi];
      for (int d = 0; d < 3; ++d) planes[o * 3 + d] = params[((size_t)b * NQ + q) * 3 + d];
      u64 area, xs, ys, xmin, xmax, ymin, ymax;
      double eps = 1e-10;                                  // :734-735
      if (flag & FLAG_FALLBACK) {                          // mask = (ids == pi), centre without eps (:743, :781-782)
        area = s[S_AREA_A]; xs = s[S_XS_A]; ys = s[S_YS_A];
        xmin = s[S_XMIN_A]; xmax = s[S_XMAX_A]; ymin = s[S_YMIN_A]; ymax = s[S_YMAX_A];
        eps = 0.0;
      } else if (flag & FLAG_PATCH00) {                    // empty mask of the forced plane: pixel (0,0) is set
        area = 1; xs = ys = 0; xmin = xmax = ymin = ymax = 0;
      } else {
        area = s[S_AREA]; xs = s[S_XS]; ys = s[S_YS];
        xmin = s[S_XMIN]; xmax = s[S_XMAX]; ymin = s[S_YMIN]; ymax = s[S_YMAX];
      }
      areas[o] = (int32_t)area;
      centers[o * 2 + 0] = (float)(((double)xs / FIX_SCALE) / ((double)area + eps));
      centers[o * 2 + 1] = (float)(((double)ys / FIX_SCALE) / ((double)area + eps));
      if (area == 0ull) {
        for (int d = 0; d < 4; ++d) bboxes[o * 4 + d] = 0.f;           // rleToBbox of an empty mask
      } else {
        bboxes[o * 4 + 0] = (float)xmin;
        bboxes[o * 4 + 1] = (float)ymin;
        bboxes[o * 4 + 2] = (float)(xmax - xmin + 1);
        bboxes[o * 4 + 3] = (float)(ymax - ymin + 1);
      }
    } else {
      ori_idx[o] = -1;
      scores[o] = 0.f;
      areas[o] = 0;
      for (int d = 0; d < 3; ++d) planes[o * 3 + d] = 0.f;
      for (int d = 0; d < 2; ++d) centers[o * 2 + d] = 0.f;
      for (int d = 0; d < 4; ++d) bboxes[o * 4 + d] = 0.f;
    }
  }
  for (int i = tid; i < NQ * C; i += 128) {
    const int j = i / C, c = i - j * C;
    float v = 0.f;
    if (j < n) v = query_feat[((size_t)b * NQ + ws.valid_q[(size_t)b * NQ + s_keep[j]]) * C + c];
    feats[(size_t)b * NQ * C + i] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// 4. final label map
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
plane_seg_kernel(PlaneWorkspace ws, const int32_t* __restrict__ flags, size_t HW, uint8_t* __restrict__ seg) {
  __shared__ uint8_t s_map[256];
  const int b = blockIdx.y;
  const int flag = flags[b];
  // fallback: the un-thresholded region of the one kept plane; otherwise thresholded pixels of kept planes
  const int tid = threadIdx.x;
  {
    const int k = tid & 127;
    uint8_t m = ws.remap[(size_t)b * 128 + k];
    if (tid >= 128 && !(flag & FLAG_FALLBACK)) m = 255;   // entries 128..255: "below the mask threshold"
    s_map[tid] = m;
  }
  __syncthreads();
  const uint8_t* raw = ws.raw + (size_t)b * HW;
  uint8_t* out = seg + (size_t)b * HW;
  const bool vec = (HW % 16 == 0);
  if (vec) {
    const size_t nvec = HW / 16;
    for (size_t i = (size_t)blockIdx.x * 256 + tid; i < nvec; i += (size_t)gridDim.x * 256) {
      uint4 v = reinterpret_cast<const uint4*>(raw)[i];
      unsigned* wds = reinterpret_cast<unsigned*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned x = wds[j];
        wds[j] = (unsigned)s_map[x & 255] | ((unsigned)s_map[(x >> 8) & 255] << 8) | ((unsigned)s_map[(x >> 16) & 255] << 16) |
                 ((unsigned)s_map[x >> 24] << 24);
      }
      if (i == 0 && (flag & FLAG_PATCH00)) wds[0] = (wds[0] & ~0xffu);   // pixel (0,0) -> kept plane 0
      reinterpret_cast<uint4*>(out)[i] = v;
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * 256 + tid; i < HW; i += (size_t)gridDim.x * 256) {
      uint8_t m = s_map[raw[i]];
      if (i == 0 && (flag & FLAG_PATCH00)) m = 0;
      out[i] = m;
    }
  }
}
}  // namespace

extern "C" size_t nsac_plane_post_workspace_bytes(int B, int NQ, int H, int W) {
  if (B < 1 || NQ < 1 || H < 1 || W < 1) return 0;
  PlaneWorkspace ws;
  return carve(ws, nullptr, B, NQ, H, W);
}

extern "C" int nsac_plane_postprocess(const float* pred_logits, const float* pred_params, const float* mask_logits,
                                      const float* query_feat, int B, int NQ, int C, int h, int w, int H, int W,
                                      float plane_score_thr, float mask_prob_thr, double overlap_thr, int32_t* count,
                                      int32_t* flags, int32_t* ori_idx, float* planes, float* feats, float* scores, float* centers,
                                      float* bboxes, int32_t* areas, uint8_t* seg, void* workspace, void* stream) {
  NSAC_REQUIRE(pred_logits && pred_params && mask_logits && query_feat && count && flags && ori_idx && planes && feats && scores &&
                   centers && bboxes && areas && seg && workspace,
               "nsac_plane_postprocess: null pointer");
  NSAC_REQUIRE(B >= 1 && B <= 65535 && NQ >= 1 && NQ <= PL_MAXQ && C >= 1 && h >= 1 && w >= 1,
               "nsac_plane_postprocess: bad shape B=%d NQ=%d (<= %d) C=%d h=%d w=%d", B, NQ, PL_MAXQ, C, h, w);
  NSAC_REQUIRE((H == 4 * h && W == 4 * w) || (H == 2 * h && W == 2 * w),
               "nsac_plane_postprocess: output %dx%d must be 2x or 4x the mask resolution %dx%d", H, W, h, w);
  NSAC_REQUIRE(H <= 65536 && W <= 65536 && (size_t)H * W <= ((size_t)1 << 23), "nsac_plane_postprocess: image %dx%d too large", H, W);
  NSAC_REQUIRE(((uintptr_t)workspace & 15) == 0 && ((uintptr_t)seg & 15) == 0, "nsac_plane_postprocess: workspace / seg must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PlaneWorkspace ws;
  carve(ws, workspace, B, NQ, H, W);
  plane_select_kernel<<<B, 128, 0, st>>>(pred_logits, NQ, plane_score_thr, ws, flags, H, W);
  NSAC_CHECK_LAUNCH("nsac_plane_postprocess(select)");
  const dim3 grid(nsac_cdiv(w + 1, 32), nsac_cdiv(h + 1, 8), B), block(32, 8);
  if (H == 4 * h)
    plane_argmax_kernel<4><<<grid, block, 0, st>>>(mask_logits, NQ, h, w, mask_prob_thr, ws);
  else
    plane_argmax_kernel<2><<<grid, block, 0, st>>>(mask_logits, NQ, h, w, mask_prob_thr, ws);
  NSAC_CHECK_LAUNCH("nsac_plane_postprocess(argmax)");
  plane_finalize_kernel<<<B, 128, 0, st>>>(pred_params, query_feat, NQ, C, overlap_thr, ws, count, flags, ori_idx, planes, feats,
                                          scores, centers, bboxes, areas);
  NSAC_CHECK_LAUNCH("nsac_plane_postprocess(finalize)");
  const size_t HW = (size_t)H * W;
  const int gx = (int)((HW / 16 + 255) / 256 > 0 ? (HW / 16 + 255) / 256 : 1);
  plane_seg_kernel<<<dim3(gx < 64 ? gx : 64, B), 256, 0, st>>>(ws, flags, HW, seg);
  NSAC_CHECK_LAUNCH("nsac_plane_postprocess(seg)");
  return NSAC_OK;
}
