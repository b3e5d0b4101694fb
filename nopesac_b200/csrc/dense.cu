// Dense building blocks of the path, exact-fp32 CUDA-core versions:
//   nsac_linear      every nn.Linear / MLP layer (camera_modules.py:226-244, gnn.py:56-67)
//   nsac_layernorm   gnn.py:90,94-96 (+ residual)
//   nsac_attention   gnn.py:19-44
//   nsac_pose_heads  camera_head.py:990,1018 (shared `rots` / `trans` Linear + quaternion normalise)
// The split-bf16 tcgen05 engine (gemm_tc.cu) takes over the large contractions; this file stays the
// exact-fp32 path for small / odd shapes (K = 3, 4, 8, N = 3, 4) and the bring-up reference.
#include <stdarg.h>
#include <cuda_fp16.h>
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing (one definition for the whole library)
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void nsac_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* nsac_last_error(void) { return g_err; }
extern "C" int nsac_version(void) { return NSAC_VERSION; }

// ------------------------------------------------------------------------------------------------
// SIMT fp32 GEMM  C[M,N] = act(A[M,K] W[N,K]^T + bias)      (both operands K-contiguous)
// 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile, register-prefetched global loads.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, GEMM_THREADS = 256;

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == NSAC_ACT_RELU) return fmaxf(v, 0.f);
  if (act == NSAC_ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
  return v;
}

template <bool VEC>
__device__ __forceinline__ void load_tile(const float* __restrict__ src, int ld, int rows, int K,
                                          int row0, int k0, float (&reg)[2][4]) {
  // 128 rows x 16 k = 512 float4; 256 threads -> 2 float4 each.
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = tid + i * GEMM_THREADS;  // 0..511
    const int r = idx >> 2, kq = (idx & 3) * 4;
    const int gr = row0 + r, gk = k0 + kq;
    if (VEC) {
      if (gr < rows && gk < K) {  // K % 4 == 0 in VEC mode -> whole float4 in range
        const float4 v = *reinterpret_cast<const float4*>(src + (size_t)gr * ld + gk);
        reg[i][0] = v.x; reg[i][1] = v.y; reg[i][2] = v.z; reg[i][3] = v.w;
      } else {
        reg[i][0] = reg[i][1] = reg[i][2] = reg[i][3] = 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        reg[i][j] = (gr < rows && gk + j < K) ? src[(size_t)gr * ld + gk + j] : 0.f;
    }
  }
}

__device__ __forceinline__ void store_tile(float (*dst)[BM + 4], const float (&reg)[2][4]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = tid + i * GEMM_THREADS;
    const int r = idx >> 2, kq = (idx & 3) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[kq + j][r] = reg[i][j];
  }
}

template <bool VEC>
__global__ void __launch_bounds__(GEMM_THREADS)
linear_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                   const float* __restrict__ bias, int bias_group_rows, float* __restrict__ C, int ldc,
                   int M, int N, int K, int act) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Ws[2][BK][BN + 4];
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[2][4], rw[2][4];
  load_tile<VEC>(A, lda, M, K, m0, 0, ra);
  load_tile<VEC>(W, K, N, K, n0, 0, rw);
  store_tile(As[0], ra);
  store_tile(Ws[0], rw);
  __syncthreads();
  const int nk = (K + BK - 1) / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_tile<VEC>(A, lda, M, K, m0, (kt + 1) * BK, ra);
      load_tile<VEC>(W, K, N, K, n0, (kt + 1) * BK, rw);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ws[cur][k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(As[cur ^ 1], ra);
      store_tile(Ws[cur ^ 1], rw);
    }
    __syncthreads();
  }
  // epilogue: rows ty*4+{0..3}, 64+ty*4+{0..3}; cols tx*4+{0..3}, 64+tx*4+{0..3}
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
    const float* brow = bias ? (bias_group_rows > 0 ? bias + (size_t)(r / bias_group_rows) * N : bias) : nullptr;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int c = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cc = c + j;
        float t = acc[i][jh * 4 + j];
        if (cc < N) {
          if (brow) t += brow[cc];
          t = apply_act(t, act);
        }
        v[j] = t;
      }
      float* dst = C + (size_t)r * ldc + c;
      if (c + 3 < N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < N) dst[j] = v[j];
      }
    }
  }
}

}  // namespace

extern "C" int nsac_linear(const float* x, int ldx, const float* w, const float* bias,
                           int bias_group_rows, float* out, int ldo, int M, int N, int K, int act,
                           void* stream) {
  NSAC_REQUIRE(x && w && out, "nsac_linear: null pointer");
  NSAC_REQUIRE(M >= 0 && N > 0 && K > 0, "nsac_linear: bad shape M=%d N=%d K=%d", M, N, K);
  NSAC_REQUIRE(ldx >= K && ldo >= N, "nsac_linear: ldx=%d < K=%d or ldo=%d < N=%d", ldx, K, ldo, N);
  NSAC_REQUIRE(act >= 0 && act <= 2, "nsac_linear: bad act %d", act);
  if (M == 0) return NSAC_OK;
  dim3 grid(nsac_cdiv(M, BM), nsac_cdiv(N, BN));
  const bool vec = (K % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (vec)
    linear_simt_kernel<true><<<grid, GEMM_THREADS, 0, s>>>(x, ldx, w, bias, bias_group_rows, out, ldo, M, N, K, act);
  else
    linear_simt_kernel<false><<<grid, GEMM_THREADS, 0, s>>>(x, ldx, w, bias, bias_group_rows, out, ldo, M, N, K, act);
  NSAC_CHECK_LAUNCH("nsac_linear");
  return NSAC_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (+ residual), one warp per row
// ------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ void split_f16(float x, uint16_t& hi, uint16_t& lo) {
  const __half h = __float2half_rn(x);
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(__float2half_rn(x - __half2float(h)));
}

__global__ void layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const float* res, int ldres,
                                 float* out, int ldo, uint16_t* out_hi, uint16_t* out_lo, int ld_split, int rows,
                                 int C) {  // res may alias out (in-place residual)
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * ldx;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum(s) / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)C + 1e-5f);
  for (int c = lane; c < C; c += 32) {
    float y = (xr[c] - mean) * rstd * gamma[c] + beta[c];
    if (res) y += res[(size_t)row * ldres + c];
    out[(size_t)row * ldo + c] = y;
    if (out_hi) {       // fp16 hi/lo operand planes for the tensor-core engine
      uint16_t h, l;
      split_f16(y, h, l);
      out_hi[(size_t)row * ld_split + c] = h;
      out_lo[(size_t)row * ld_split + c] = l;
    }
  }
}
}  // namespace

extern "C" int nsac_layernorm(const float* x, int ldx, const float* gamma, const float* beta,
                              const float* res, int ldres, float* out, int ldo, void* out_hi, void* out_lo,
                              int ld_split, int rows, int C, void* stream) {
  NSAC_REQUIRE(x && gamma && beta && out, "nsac_layernorm: null pointer");
  NSAC_REQUIRE(rows >= 0 && C > 0 && ldx >= C && ldo >= C, "nsac_layernorm: bad shape");
  if (rows == 0) return NSAC_OK;
  const int wpb = 8;
  layernorm_kernel<<<nsac_cdiv(rows, wpb), wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, gamma, beta, res, ldres, out, ldo, static_cast<uint16_t*>(out_hi), static_cast<uint16_t*>(out_lo), ld_split, rows, C);
  NSAC_CHECK_LAUNCH("nsac_layernorm");
  return NSAC_OK;
}

// ------------------------------------------------------------------------------------------------
// Multi-head attention over <= a few dozen plane tokens: one CTA per (pair, slice of the queries), one warp per
// head, D = 32 = one lane per channel.  K/V of the head staged in shared memory (every query slice re-stages them:
// 64 CTAs of 16 serial queries took 26 us per call; slicing the queries over gridDim.y fills the SMs).
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void attention_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                 const float* __restrict__ v, int ldkv, float* __restrict__ out, int ldo,
                                 uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo, int ld_split,
                                 int L, int S_pad, int H, const int32_t* __restrict__ kv_count) {
  extern __shared__ float sm[];  // per warp: K [S][33], V [S][33], p [S]
  const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (h >= H) return;
  float* Ks = sm + (size_t)h * (S_pad * 66 + S_pad);
  float* Vs = Ks + S_pad * 33;
  float* ps = Vs + S_pad * 33;
  // ragged batches: only the first kv_count[b] source tokens of this batch element exist (row stride stays S_pad)
  const int S = kv_count ? min(max(kv_count[b], 1), S_pad) : S_pad;
#pragma unroll 4
  for (int s = 0; s < S; ++s) {
    Ks[s * 33 + lane] = k[((size_t)b * S_pad + s) * ldkv + h * 32 + lane];
    Vs[s * 33 + lane] = v[((size_t)b * S_pad + s) * ldkv + h * 32 + lane];
  }
  __syncwarp();
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)
  const int lq = (L + gridDim.y - 1) / gridDim.y, l_end = min(L, (int)(blockIdx.y + 1) * lq);
  for (int l = blockIdx.y * lq; l < l_end; ++l) {
    const float qd = q[((size_t)b * L + l) * ldq + h * 32 + lane];
    float mx = -INFINITY;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        const float qv = __shfl_sync(NSAC_FULL_MASK, qd, d);
        if (s < S) dot = fmaf(qv, Ks[s * 33 + d], dot);
      }
      if (s < S) {
        dot *= scale;
        ps[s] = dot;
        mx = fmaxf(mx, dot);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < S; s += 32) {
      const float e = expf(ps[s] - mx);
      ps[s] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc = fmaf(ps[s], Vs[s * 33 + lane], acc);
    const float o = acc / sum;
    if (out) out[((size_t)b * L + l) * ldo + h * 32 + lane] = o;
    if (out_hi) {
      uint16_t hh, ll;
      split_f16(o, hh, ll);
      out_hi[((size_t)b * L + l) * ld_split + h * 32 + lane] = hh;
      out_lo[((size_t)b * L + l) * ld_split + h * 32 + lane] = ll;
    }
    __syncwarp();
  }
}
}  // namespace

extern "C" int nsac_attention_ragged(const float* q, int ldq, const float* k, const float* v, int ldkv,
                                     float* out, int ldo, void* out_hi, void* out_lo, int ld_split, int B, int L, int S,
                                     int H, int D, const int32_t* kv_count, void* stream) {
  NSAC_REQUIRE(q && k && v && (out || (out_hi && out_lo)), "nsac_attention: null pointer");
  NSAC_REQUIRE(D == 32, "nsac_attention: head dim must be 32 (got %d)", D);
  NSAC_REQUIRE(H >= 1 && H <= 32 && L >= 0 && S >= 1, "nsac_attention: bad shape");
  if (B == 0 || L == 0) return NSAC_OK;
  const size_t smem = (size_t)H * (S * 66 + S) * sizeof(float);
  NSAC_REQUIRE(smem <= 200 * 1024, "nsac_attention: S=%d too large for the shared-memory staging", S);
  if (smem > 48 * 1024)
    NSAC_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int slices = (4 * 148 + B - 1) / B;        // ~4 CTAs per SM
  if (slices > L) slices = L;
  if (slices < 1) slices = 1;
  attention_kernel<<<dim3(B, slices), H * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      q, ldq, k, v, ldkv, out, ldo, static_cast<uint16_t*>(out_hi), static_cast<uint16_t*>(out_lo), ld_split, L, S, H,
      kv_count);
  NSAC_CHECK_LAUNCH("nsac_attention");
  return NSAC_OK;
}

extern "C" int nsac_attention(const float* q, int ldq, const float* k, const float* v, int ldkv,
                              float* out, int ldo, void* out_hi, void* out_lo, int ld_split, int B, int L, int S,
                              int H, int D, void* stream) {
  return nsac_attention_ragged(q, ldq, k, v, ldkv, out, ldo, out_hi, out_lo, ld_split, B, L, S, H, D, nullptr, stream);
}

// ------------------------------------------------------------------------------------------------
// Pose heads: q = normalize(Wr f + br) (F.normalize eps 1e-12), t = Wt f + bt.  One warp per row.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void pose_heads_kernel(const float* __restrict__ fr, const float* __restrict__ ft,
                                  const float* __restrict__ wr, const float* __restrict__ br,
                                  const float* __restrict__ wt, const float* __restrict__ bt, int rows,
                                  int C, float* __restrict__ qo, float* __restrict__ to) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  if (fr && qo) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = lane; c < C; c += 32) {
      const float f = fr[(size_t)row * C + c];
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] = fmaf(f, wr[j * C + c], a[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = warp_sum(a[j]) + br[j];
    const float n = fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]), 1e-12f);
    if (lane < 4) qo[(size_t)row * 4 + lane] = a[lane] / n;
  }
  if (ft && to) {
    float a[3] = {0.f, 0.f, 0.f};
    for (int c = lane; c < C; c += 32) {
      const float f = ft[(size_t)row * C + c];
#pragma unroll
      for (int j = 0; j < 3; ++j) a[j] = fmaf(f, wt[j * C + c], a[j]);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) a[j] = warp_sum(a[j]) + bt[j];
    if (lane < 3) to[(size_t)row * 3 + lane] = a[lane];
  }
}
}  // namespace

extern "C" int nsac_pose_heads(const float* feat_rot, const float* feat_tran, const float* w_rots,
                               const float* b_rots, const float* w_trans, const float* b_trans, int rows,
                               int C, float* q_out, float* t_out, void* stream) {
  NSAC_REQUIRE((feat_rot && q_out && w_rots && b_rots) || (feat_tran && t_out && w_trans && b_trans),
               "nsac_pose_heads: need at least one complete branch");
  NSAC_REQUIRE(rows >= 0 && C > 0, "nsac_pose_heads: bad shape");
  if (rows == 0) return NSAC_OK;
  const int wpb = 8;
  pose_heads_kernel<<<nsac_cdiv(rows, wpb), wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      feat_rot, feat_tran, w_rots, b_rots, w_trans, b_trans, rows, C, q_out, t_out);
  NSAC_CHECK_LAUNCH("nsac_pose_heads");
  return NSAC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Initial-pose hand-off (camera_head.py:436-437, 718): quaternion with w >= 0 (per pair), t + 1e-10 for the AIM translation
// embedding, and the matcher pose cam = [t, q] rows (:493) — the three elementwise torch ops between K1 / K2 and the matcher.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
__global__ void pose_canon_kernel(const float* q_in, const float* t_in, int B, float* q_out, float* t_eps) {   // q_out may alias q_in
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (q_in && q_out) {
    const float w = q_in[b * 4];
    const float s = w < 0.f ? -1.f : 1.f;
    for (int j = 0; j < 4; ++j) q_out[b * 4 + j] = s < 0.f ? -q_in[b * 4 + j] : q_in[b * 4 + j];
  }
  if (t_in && t_eps)
    for (int j = 0; j < 3; ++j) t_eps[b * 3 + j] = t_in[b * 3 + j] + 1e-10f;
}
__global__ void cam_rows_kernel(const float* __restrict__ t, const float* __restrict__ q, int B, float* __restrict__ cam) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int j = 0; j < 3; ++j) cam[b * 7 + j] = t[b * 3 + j];
  for (int j = 0; j < 4; ++j) cam[b * 7 + 3 + j] = q[b * 4 + j];
}
}  // namespace

extern "C" int nsac_pose_canon(const float* q_in, const float* t_in, int B, float* q_out, float* t_eps_out, void* stream) {
  NSAC_REQUIRE((q_in && q_out) || (t_in && t_eps_out), "nsac_pose_canon: nothing to do");
  NSAC_REQUIRE(B >= 0, "nsac_pose_canon: bad batch size");
  if (B == 0) return NSAC_OK;
  pose_canon_kernel<<<nsac_cdiv(B, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(q_in, t_in, B, q_out, t_eps_out);
  NSAC_CHECK_LAUNCH("nsac_pose_canon");
  return NSAC_OK;
}

extern "C" int nsac_cam_rows(const float* t, const float* q, int B, float* cam, void* stream) {
  NSAC_REQUIRE(t && q && cam && B >= 0, "nsac_cam_rows: bad arguments");
  if (B == 0) return NSAC_OK;
  cam_rows_kernel<<<nsac_cdiv(B, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(t, q, B, cam);
  NSAC_CHECK_LAUNCH("nsac_cam_rows");
  return NSAC_OK;
}
