// Matching tail of the path — one CTA per image pair, the whole (n1+1)x(n2+1) transport problem in
// shared memory, no host round-trips:
//   K3  geometry penalties              matching_net/matching_head.py:75-99
//   K5  similarity, dustbins, Sinkhorn  matching_head.py:113-128, 228-234, 259-306
//       mutual-NN + threshold           camera_net/camera_modules.py:15-34
//   K10 assignment pruning              camera_net/camera_head.py:605-629
#include "common.cuh"

namespace {

constexpr float kPi = 3.14159265358979323846f;

// view-2 planes: flip, offset, unit normal  (matching_head.py:76-79)
// view-1 planes: warp by (q,0) -> unit normal ; warp by (q,t) -> offset + unit normal (:81-93)
struct PlaneGeo {
  float* n2;    // [n2][3]
  float* off2;  // [n2]
  float* n1r;   // [n1][3]
  float* n1rt;  // [n1][3]
  float* off1;  // [n1]
};

__device__ void plane_geometry(const float* __restrict__ planes1, const float* __restrict__ planes2,
                               const float* __restrict__ t, const float* __restrict__ q, int n1, int n2,
                               PlaneGeo g) {
  const Mat3 R = quat_to_rot(q[0], q[1], q[2], q[3]);
  for (int j = threadIdx.x; j < n2; j += blockDim.x) {
    float x = planes2[j * 3 + 0], y = -planes2[j * 3 + 1], z = -planes2[j * 3 + 2];
    g.off2[j] = normalize3(x, y, z);
    g.n2[j * 3 + 0] = x; g.n2[j * 3 + 1] = y; g.n2[j * 3 + 2] = z;
  }
  for (int i = threadIdx.x; i < n1; i += blockDim.x) {
    const float px = planes1[i * 3 + 0], py = planes1[i * 3 + 1], pz = planes1[i * 3 + 2];
    float x, y, z;
    warp_plane(R, 0.f, 0.f, 0.f, px, py, pz, x, y, z);
    normalize3(x, y, z);
    g.n1r[i * 3 + 0] = x; g.n1r[i * 3 + 1] = y; g.n1r[i * 3 + 2] = z;
    warp_plane(R, t[0], t[1], t[2], px, py, pz, x, y, z);
    g.off1[i] = normalize3(x, y, z);
    g.n1rt[i * 3 + 0] = x; g.n1rt[i * 3 + 1] = y; g.n1rt[i * 3 + 2] = z;
  }
}

// normal angle in degrees and (unclamped) offset distance of plane pair (i,j)
__device__ __forceinline__ void pair_penalty(const PlaneGeo& g, int i, int j, float& angle_deg, float& off_dist) {
  const float cr = g.n1r[i * 3] * g.n2[j * 3] + g.n1r[i * 3 + 1] * g.n2[j * 3 + 1] + g.n1r[i * 3 + 2] * g.n2[j * 3 + 2];
  angle_deg = acosf(fminf(fmaxf(cr, -1.f), 1.f)) / kPi * 180.f;
  const float crt = g.n1rt[i * 3] * g.n2[j * 3] + g.n1rt[i * 3 + 1] * g.n2[j * 3 + 1] + g.n1rt[i * 3 + 2] * g.n2[j * 3 + 2];
  off_dist = crt < 0.f ? fabsf(g.off1[i] + g.off2[j]) : fabsf(g.off1[i] - g.off2[j]);
}

__device__ __forceinline__ PlaneGeo carve_geo(float*& p, int n1, int n2) {
  PlaneGeo g;
  g.n2 = p; p += n2 * 3;
  g.off2 = p; p += n2;
  g.n1r = p; p += n1 * 3;
  g.n1rt = p; p += n1 * 3;
  g.off1 = p; p += n1;
  return g;
}

__host__ __device__ inline int geo_floats(int n1, int n2) { return n2 * 4 + n1 * 7; }

__global__ void match_sinkhorn_assign_kernel(
    const float* __restrict__ desc1, const float* __restrict__ desc2, const float* __restrict__ planes1,
    const float* __restrict__ planes2, const float* __restrict__ cam, const float* __restrict__ bin_score,
    float offset_mult, float normal_mult, int iters, float threshold, int n1_pad, int n2_pad, int C,
    const int32_t* __restrict__ count1, const int32_t* __restrict__ count2,
    float* __restrict__ lsp_out, float* __restrict__ assign_out) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  // ragged batches: pair b has count1[b] x count2[b] planes in the top-left corner of the padded [n1_pad, n2_pad] layout;
  // its transport problem, dustbins and normalisation are those of the un-padded pair (the reference runs one pair at a time)
  const int n1 = count1 ? min(max(count1[b], 1), n1_pad) : n1_pad;
  const int n2 = count2 ? min(max(count2[b], 1), n2_pad) : n2_pad;
  const int R = n1 + 1, Cc = n2 + 1;
  const int ld = (Cc & 1) ? Cc : Cc + 1;  // odd stride: conflict-free row- and column-walks
  float* p = sm;
  float* Z = p; p += R * ld;
  float* u = p; p += R;
  float* v = p; p += Cc;
  int* idx0 = reinterpret_cast<int*>(p); p += n1;
  int* idx1 = reinterpret_cast<int*>(p); p += n2;
  float* max0 = p; p += n1;
  PlaneGeo g = carve_geo(p, n1, n2);

  desc1 += (size_t)b * n1_pad * C;
  desc2 += (size_t)b * n2_pad * C;
  planes1 += (size_t)b * n1_pad * 3;
  planes2 += (size_t)b * n2_pad * 3;
  cam += (size_t)b * 7;

  plane_geometry(planes1, planes2, cam, cam + 3, n1, n2, g);
  __syncthreads();

  // similarity minus penalties (matching_head.py:113-119): one warp per (i,j)
  for (int e = warp; e < n1 * n2; e += nwarps) {
    const int i = e / n2, j = e - i * n2;
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) dot = fmaf(desc1[i * C + c], desc2[j * C + c], dot);
    dot = warp_sum(dot);
    if (lane == 0) {
      float ang, off;
      pair_penalty(g, i, j, ang, off);
      off = fminf(fmaxf(off, 1e-10f), 5.f);
      float s = dot / 16.f;  // 256 ** .5
      s = s - off / offset_mult;
      s = s - ang / normal_mult;
      Z[i * ld + j] = s;
    }
  }
  const float alpha = bin_score[0];
  for (int i = tid; i < n1; i += blockDim.x) Z[i * ld + n2] = alpha;
  for (int j = tid; j < Cc; j += blockDim.x) Z[n1 * ld + j] = alpha;
  for (int i = tid; i < R; i += blockDim.x) u[i] = 0.f;
  for (int j = tid; j < Cc; j += blockDim.x) v[j] = 0.f;
  __syncthreads();

  // log-domain Sinkhorn (matching_head.py:228-234), thread r owns row r / column r
  const float norm = -logf((float)n1 + (float)n2);
  const float log_mu_last = logf((float)n2) + norm, log_nu_last = logf((float)n1) + norm;
  for (int it = 0; it < iters; ++it) {
    if (tid < R) {
      const float* zr = Z + tid * ld;
      float mx = -INFINITY;
      for (int j = 0; j < Cc; ++j) mx = fmaxf(mx, zr[j] + v[j]);
      float s = 0.f;
      for (int j = 0; j < Cc; ++j) s += expf(zr[j] + v[j] - mx);
      u[tid] = (tid < n1 ? norm : log_mu_last) - (logf(s) + mx);
    }
    __syncthreads();
    if (tid < Cc) {
      float mx = -INFINITY;
      for (int i = 0; i < R; ++i) mx = fmaxf(mx, Z[i * ld + tid] + u[i]);
      float s = 0.f;
      for (int i = 0; i < R; ++i) s += expf(Z[i * ld + tid] + u[i] - mx);
      v[tid] = (tid < n2 ? norm : log_nu_last) - (logf(s) + mx);
    }
    __syncthreads();
  }
  // Z + u + v - norm  (:234, :304)
  const int Rp = n1_pad + 1, Cp = n2_pad + 1;
  float* lsp = lsp_out + (size_t)b * Rp * Cp;
  for (int e = tid; e < Rp * Cp; e += blockDim.x) {
    const int i = e / Cp, j = e - i * Cp;
    float val = -INFINITY;                       // outside the pair's (n1+1) x (n2+1) block: probability 0
    if (i < R && j < Cc) {
      val = Z[i * ld + j] + u[i] + v[j] - norm;
      Z[i * ld + j] = val;
    }
    lsp[e] = val;
  }
  __syncthreads();

  // mutual nearest neighbour + threshold (camera_modules.py:15-32); ties -> lowest index
  for (int i = tid; i < n1; i += blockDim.x) {
    float best = Z[i * ld];
    int bj = 0;
    for (int j = 1; j < n2; ++j) {
      const float z = Z[i * ld + j];
      if (z > best) { best = z; bj = j; }
    }
    idx0[i] = bj;
    max0[i] = best;
  }
  for (int j = tid; j < n2; j += blockDim.x) {
    float best = Z[j];
    int bi = 0;
    for (int i = 1; i < n1; ++i) {
      const float z = Z[i * ld + j];
      if (z > best) { best = z; bi = i; }
    }
    idx1[j] = bi;
  }
  __syncthreads();
  float* A = assign_out + (size_t)b * n1_pad * n2_pad;
  for (int e = tid; e < n1_pad * n2_pad; e += blockDim.x) {
    const int i = e / n2_pad, j = e - i * n2_pad;
    bool one = false;
    if (i < n1 && j < n2) {
      const bool mutual = idx1[idx0[i]] == i;
      const bool valid = mutual && (expf(max0[i]) > threshold);
      one = valid && idx0[i] == j;
    }
    A[e] = one ? 1.f : 0.f;
  }
}

__global__ void prune_assignment_kernel(const float* __restrict__ assign, const float* __restrict__ planes1,
                                        const float* __restrict__ planes2, const float* __restrict__ pose,
                                        int ldpose, int n1, int n2, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int b = blockIdx.x;
  float* p = sm;
  PlaneGeo g = carve_geo(p, n1, n2);
  const float* ps = pose + (size_t)b * ldpose;
  plane_geometry(planes1 + (size_t)b * n1 * 3, planes2 + (size_t)b * n2 * 3, ps, ps + 3, n1, n2, g);
  __syncthreads();
  for (int e = threadIdx.x; e < n1 * n2; e += blockDim.x) {
    const int i = e / n2, j = e - i * n2;
    float ang, off;
    pair_penalty(g, i, j, ang, off);
    off = fminf(fmaxf(off, 1e-4f), 10.f);
    const bool keep = (ang < 45.f) && (off < 1.f);
    out[(size_t)b * n1 * n2 + e] = keep ? assign[(size_t)b * n1 * n2 + e] : 0.f;
  }
}
}  // namespace

extern "C" int nsac_match_sinkhorn_assign_ragged(const float* desc1, const float* desc2, const float* planes1,
                                                 const float* planes2, const float* cam, const float* bin_score,
                                                 float offset_mult, float normal_mult, int iters, float threshold,
                                                 int B, int n1, int n2, int C, const int32_t* count1,
                                                 const int32_t* count2, float* log_scores_padded, float* assign,
                                                 void* stream) {
  NSAC_REQUIRE(desc1 && desc2 && planes1 && planes2 && cam && bin_score && log_scores_padded && assign,
               "nsac_match_sinkhorn_assign: null pointer");
  NSAC_REQUIRE(B >= 0 && n1 >= 1 && n2 >= 1 && C >= 1 && iters >= 0, "nsac_match_sinkhorn_assign: bad shape");
  NSAC_REQUIRE(n1 < 1024 && n2 < 1024, "nsac_match_sinkhorn_assign: at most 1023 planes per view");
  if (B == 0) return NSAC_OK;
  const int R = n1 + 1, Cc = n2 + 1, ld = (Cc & 1) ? Cc : Cc + 1;
  const size_t smem = sizeof(float) * ((size_t)R * ld + R + Cc + n1 + n2 + n1 + geo_floats(n1, n2));
  NSAC_REQUIRE(smem <= 220 * 1024, "nsac_match_sinkhorn_assign: %d x %d planes exceed shared memory", n1, n2);
  if (smem > 48 * 1024)
    NSAC_CUDA(cudaFuncSetAttribute(match_sinkhorn_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int threads = ((R > Cc ? R : Cc) + 31) / 32 * 32;
  if (threads < 128) threads = 128;
  match_sinkhorn_assign_kernel<<<B, threads, smem, static_cast<cudaStream_t>(stream)>>>(
      desc1, desc2, planes1, planes2, cam, bin_score, offset_mult, normal_mult, iters, threshold, n1, n2, C,
      count1, count2, log_scores_padded, assign);
  NSAC_CHECK_LAUNCH("nsac_match_sinkhorn_assign");
  return NSAC_OK;
}

extern "C" int nsac_match_sinkhorn_assign(const float* desc1, const float* desc2, const float* planes1,
                                          const float* planes2, const float* cam, const float* bin_score,
                                          float offset_mult, float normal_mult, int iters, float threshold,
                                          int B, int n1, int n2, int C, float* log_scores_padded,
                                          float* assign, void* stream) {
  return nsac_match_sinkhorn_assign_ragged(desc1, desc2, planes1, planes2, cam, bin_score, offset_mult, normal_mult, iters,
                                           threshold, B, n1, n2, C, nullptr, nullptr, log_scores_padded, assign, stream);
}

extern "C" int nsac_prune_assignment(const float* assign, const float* planes1, const float* planes2,
                                     const float* pose, int ldpose, int B, int n1, int n2, float* assign_out,
                                     void* stream) {
  NSAC_REQUIRE(assign && planes1 && planes2 && pose && assign_out, "nsac_prune_assignment: null pointer");
  NSAC_REQUIRE(B >= 0 && n1 >= 1 && n2 >= 1 && ldpose >= 7, "nsac_prune_assignment: bad shape");
  if (B == 0) return NSAC_OK;
  const size_t smem = sizeof(float) * geo_floats(n1, n2);
  NSAC_REQUIRE(smem <= 48 * 1024, "nsac_prune_assignment: too many planes");
  prune_assignment_kernel<<<B, 128, smem, static_cast<cudaStream_t>(stream)>>>(assign, planes1, planes2, pose,
                                                                                ldpose, n1, n2, assign_out);
  NSAC_CHECK_LAUNCH("nsac_prune_assignment");
  return NSAC_OK;
}
