// Glue kernels of the plane detector PlaneTRHead (first half; reference
// modeling/planeTR_net/planeTR_head.py:116-192 + modeling/transformer/transformer.py): everything that is not a GEMM.  The
// linear layers / 1x1 convolutions run on the tensor-core engine (nsac_gemm_split*), these kernels connect them:
//   row_op_kernel            s = x (+ y);  t = LayerNorm(s) or s;  optional outputs: s (fp32), t (fp32 / hi-lo planes) and
//                            t + pos[row % T] (planes) - the residual adds, the post-/pre-norm LayerNorms and the
//                            `with_pos_embed` adds of the DETR layers (transformer.py:170-185, 284-311) in one pass per row
//   attention_tiled_kernel   softmax(Q K^T / sqrt(32)) V for sequences that do not fit one warp's shared-memory slice
//                            (300 context tokens): one CTA per (image, head), K^T / V of the head staged once in shared
//                            memory, 64-query blocks, register-tiled fp32 (8 queries x 5 keys per lane, two warps per query
//                            group), exact expf softmax
//   upsample2x_relu_add_kernel  out = relu(bilinear_2x(a)) + b (align_corners = False) on NHWC maps: the top-down path of
//                            planeTR_head.py:240-252 with the 1x1 convolution + BatchNorm moved BEFORE the upsampling (both
//                            are linear / affine per pixel, so conv(up(x)) == up(conv(x)): 4x fewer GEMM rows)
// Plain SIMT (no TMA / tcgen05): also compiled for the host by tests/simt_host.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ void pt_split16(float x, uint16_t& hi, uint16_t& lo) {
  const __half h = __float2half_rn(x);
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(__float2half_rn(x - __half2float(h)));
}

// ------------------------------------------------------------------------------------------------ row op
// one warp per row, C <= 1024, C % 32 == 0
constexpr int ROW_MAX_PER_LANE = 32;
__global__ void __launch_bounds__(256)
row_op_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ y, int ldy, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, int do_ln, const float* __restrict__ pos, int T, float* __restrict__ sum_out,
              int ld_sum, float* __restrict__ t_out, int ld_t, uint16_t* __restrict__ t_hi, uint16_t* __restrict__ t_lo, int ld_tp,
              uint16_t* __restrict__ p_hi, uint16_t* __restrict__ p_lo, int ld_pp, int rows, int C) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int per = C >> 5;
  float v[ROW_MAX_PER_LANE];
  float mean = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_MAX_PER_LANE; ++i) {
    if (i < per) {
      const int c = lane + 32 * i;
      float s = x[(size_t)row * ldx + c];
      if (y) s += y[(size_t)row * ldy + c];
      v[i] = s;
      mean += s;
      if (sum_out) sum_out[(size_t)row * ld_sum + c] = s;
    }
  }
  if (do_ln) {     // two-pass variance like torch.layer_norm: mean, then sum of squared deviations
    mean = warp_sum(mean) / (float)C;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < ROW_MAX_PER_LANE; ++i)
      if (i < per) { const float d = v[i] - mean; var = fmaf(d, d, var); }
    var = warp_sum(var) / (float)C;
    const float rstd = rsqrtf(var + eps);
#pragma unroll
    for (int i = 0; i < ROW_MAX_PER_LANE; ++i)
      if (i < per) { const int c = lane + 32 * i; v[i] = (v[i] - mean) * rstd * gamma[c] + beta[c]; }
  }
  const float* prow = pos ? pos + (size_t)(row % T) * C : nullptr;
#pragma unroll
  for (int i = 0; i < ROW_MAX_PER_LANE; ++i) {
    if (i < per) {
      const int c = lane + 32 * i;
      if (t_out) t_out[(size_t)row * ld_t + c] = v[i];
      if (t_hi) { uint16_t h, l; pt_split16(v[i], h, l); t_hi[(size_t)row * ld_tp + c] = h; t_lo[(size_t)row * ld_tp + c] = l; }
      if (p_hi) { uint16_t h, l; pt_split16(v[i] + prow[c], h, l); p_hi[(size_t)row * ld_pp + c] = h; p_lo[(size_t)row * ld_pp + c] = l; }
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention
// One CTA per (image, head): K^T / V of the head staged once in shared memory, then 64-query blocks.  16 warps: warp = (query
// group g of 8 queries, key half kh) - the two warps of a group split the keys (5 of the 10 key slots per lane each), exchange
// their row maxima / sums through shared memory, and split the P V sum by key range as well.  (One warp per group - 8 warps
// per SM at this shared-memory footprint - left the FMA pipe half idle.)
constexpr int AT_D = 32, AT_QB = 64, AT_THREADS = 512, AT_KPL = 10, AT_KH = AT_KPL / 2;        // keys per lane: S <= 320
constexpr int AT_GROUPS = AT_QB / 8;
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tiled_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, const float* __restrict__ v, int ldkv,
                       float* __restrict__ out, int ldo, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo, int ld_split,
                       int L, int S, int SP, int H) {
  extern __shared__ float sm[];
  const int SK = SP + 1;                // odd row stride: the transposing stores below hit 32 different banks
  float* Kt = sm;                       // [32][SK]   K transposed (zero beyond S)
  float* Vs = Kt + AT_D * SP + AT_D;    // [S][32]    (16-byte aligned: 32 * (SP + 1) floats before it)
  float* Qt = Vs + (size_t)SP * AT_D;   // [32][64]   this block's queries, scaled, transposed
  float* Pw = Qt + AT_D * AT_QB;        // [8 groups][SP][8]  softmax numerators of the group's 8 queries
  float* Xm = Pw + (size_t)AT_GROUPS * SP * 8;      // [2 key halves][8 groups][8]  row maxima of each half
  float* Xs = Xm + 2 * AT_QB;                       // [2][8][8]                   row sums of each half
  float* Ox = Xs + 2 * AT_QB;                       // [8 groups][8][32]           P V partial of the second key half
  const int b = blockIdx.x / H, h = blockIdx.x % H, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = warp & (AT_GROUPS - 1), kh = warp / AT_GROUPS;
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)
  for (int idx = tid; idx < AT_D * SP; idx += AT_THREADS) {       // idx = s * 32 + d: coalesced 128-byte rows
    const int s = idx >> 5, d = idx & 31;
    float kv = 0.f, vv = 0.f;
    if (s < S) {
      kv = k[((size_t)b * S + s) * ldkv + h * AT_D + d];
      vv = v[((size_t)b * S + s) * ldkv + h * AT_D + d];
    }
    Kt[d * SK + s] = kv;
    Vs[idx] = vv;
  }
  float* Pmine = Pw + (size_t)g * SP * 8;
  const int nj = SP >> 5, j0 = kh * AT_KH;                      // this warp's key slots: j0 .. j0 + 4 (keys lane + 32 j)
  const int s_mid = S < 32 * AT_KH ? S : 32 * AT_KH;            // P V: keys [0, s_mid) on kh = 0, [s_mid, S) on kh = 1
  for (int q0 = 0; q0 < L; q0 += AT_QB) {
    __syncthreads();                    // K / V staged (first block); every warp is done with the previous Qt / Xm / Xs / Ox
    for (int idx = tid; idx < AT_QB * AT_D; idx += AT_THREADS) {
      const int qq = idx >> 5, d = idx & 31;
      Qt[d * AT_QB + qq] = q0 + qq < L ? q[((size_t)b * L + q0 + qq) * ldq + h * AT_D + d] * scale : 0.f;
    }
    __syncthreads();
    const bool active = q0 + g * 8 < L;   // warp-uniform; inactive warps only keep the barriers below company
    // ---- scores of the group's 8 queries against this warp's keys
    float acc[8][AT_KH];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < AT_KH; ++j) acc[i][j] = 0.f;
    if (active) {
#pragma unroll 4
      for (int d = 0; d < AT_D; ++d) {
        const float4 qa = *reinterpret_cast<const float4*>(Qt + d * AT_QB + g * 8);
        const float4 qb = *reinterpret_cast<const float4*>(Qt + d * AT_QB + g * 8 + 4);
        const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
        for (int j = 0; j < AT_KH; ++j) {
          if (j0 + j < nj) {
            const float kk = Kt[d * SK + lane + 32 * (j0 + j)];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i][j] = fmaf(qv[i], kk, acc[i][j]);
          }
        }
      }
      // row maxima over this warp's keys
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < AT_KH; ++j)
          if (j0 + j < nj && lane + 32 * (j0 + j) < S) mx = fmaxf(mx, acc[i][j]);
        mx = warp_max(mx);
        if (lane == 0) Xm[(kh * AT_GROUPS + g) * 8 + i] = mx;
      }
    }
    __syncthreads();
    float inv_sum[8];
    if (active) {
      // ---- softmax numerators (exact expf, subtracted maximum = maximum over ALL keys) into the group's P buffer
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float mx = fmaxf(Xm[g * 8 + i], Xm[(AT_GROUPS + g) * 8 + i]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < AT_KH; ++j) {
          if (j0 + j < nj) {
            const float e = lane + 32 * (j0 + j) < S ? expf(acc[i][j] - mx) : 0.f;
            acc[i][j] = e;
            sum += e;
          }
        }
        sum = warp_sum(sum);
        if (lane == 0) Xs[(kh * AT_GROUPS + g) * 8 + i] = sum;
      }
#pragma unroll
      for (int j = 0; j < AT_KH; ++j) {
        if (j0 + j < nj) {
          float* dst = Pmine + (size_t)(lane + 32 * (j0 + j)) * 8;
          *reinterpret_cast<float4*>(dst) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
        }
      }
    }
    __syncthreads();
    // ---- O = P V: lane = channel d, all 8 queries of the group per lane; this warp sums over its half of the keys.  Per key:
    // one conflict-free row read of V (32 lanes x 4 B) + two broadcast float4 of P = 3 shared-memory wavefronts for 8 FMAs
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) inv_sum[i] = 1.f / (Xs[g * 8 + i] + Xs[(AT_GROUPS + g) * 8 + i]);
      const int s_begin = kh == 0 ? 0 : s_mid, s_end = kh == 0 ? s_mid : S;
#pragma unroll 4
      for (int s = s_begin; s < s_end; ++s) {
        const float vv = Vs[s * AT_D + lane];
        const float4 pa = *reinterpret_cast<const float4*>(Pmine + s * 8), pb = *reinterpret_cast<const float4*>(Pmine + s * 8 + 4);
        o[0] = fmaf(pa.x, vv, o[0]); o[1] = fmaf(pa.y, vv, o[1]); o[2] = fmaf(pa.z, vv, o[2]); o[3] = fmaf(pa.w, vv, o[3]);
        o[4] = fmaf(pb.x, vv, o[4]); o[5] = fmaf(pb.y, vv, o[5]); o[6] = fmaf(pb.z, vv, o[6]); o[7] = fmaf(pb.w, vv, o[7]);
      }
      if (kh == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) Ox[(g * 8 + i) * AT_D + lane] = o[i];
      }
    }
    __syncthreads();
    if (active && kh == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int qrow = q0 + g * 8 + i;
        if (qrow < L) {
          const size_t base = ((size_t)b * L + qrow);
          const float val = (o[i] + Ox[(g * 8 + i) * AT_D + lane]) * inv_sum[i];
          if (out) out[base * ldo + h * AT_D + lane] = val;
          if (out_hi) {
            uint16_t hh, ll;
            pt_split16(val, hh, ll);
            out_hi[base * ld_split + h * AT_D + lane] = hh;
            out_lo[base * ld_split + h * AT_D + lane] = ll;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ top-down upsampling
// out[n, y, x, :] = relu(bilinear_2x(a)[n, y, x, :]) + b[n, y, x, :]; a [N, h, w, C] fp32, b / out [N, 2h, 2w, C];
// F.interpolate(scale_factor=2, mode="bilinear", align_corners=False): src = max(0, (dst + 0.5) / 2 - 0.5).
// One thread = 4 channels of one output pixel.
__global__ void __launch_bounds__(256)
upsample2x_relu_add_kernel(const float* __restrict__ a, const float* __restrict__ b, int N, int h, int w, int C, float* __restrict__ out,
                           uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
  const int C4 = C >> 2, H = 2 * h, W = 2 * w;
  const size_t total = (size_t)N * H * W * C4;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % C4);
    const size_t pix = idx / C4;
    const int x = (int)(pix % W), yy = (int)((pix / W) % H), n = (int)(pix / ((size_t)W * H));
    const float sy = fmaxf(0.f, (yy + 0.5f) * 0.5f - 0.5f), sx = fmaxf(0.f, (x + 0.5f) * 0.5f - 0.5f);
    const int y0 = (int)sy, x0 = (int)sx, y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* base = a + (size_t)n * h * w * C + c4 * 4;
    const float4 v00 = __ldg(reinterpret_cast<const float4*>(base + ((size_t)y0 * w + x0) * C));
    const float4 v01 = __ldg(reinterpret_cast<const float4*>(base + ((size_t)y0 * w + x1) * C));
    const float4 v10 = __ldg(reinterpret_cast<const float4*>(base + ((size_t)y1 * w + x0) * C));
    const float4 v11 = __ldg(reinterpret_cast<const float4*>(base + ((size_t)y1 * w + x1) * C));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b + pix * C + c4 * 4));
    float r[4];
    r[0] = fmaxf(hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x), 0.f) + bb.x;
    r[1] = fmaxf(hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y), 0.f) + bb.y;
    r[2] = fmaxf(hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z), 0.f) + bb.z;
    r[3] = fmaxf(hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w), 0.f) + bb.w;
    const size_t o = pix * C + c4 * 4;
    if (out) *reinterpret_cast<float4*>(out + o) = make_float4(r[0], r[1], r[2], r[3]);
    if (out_hi) {
      uint16_t hh[4], ll[4];
      for (int i = 0; i < 4; ++i) pt_split16(r[i], hh[i], ll[i]);
      *reinterpret_cast<uint2*>(out_hi + o) = make_uint2((uint32_t)hh[0] | ((uint32_t)hh[1] << 16), (uint32_t)hh[2] | ((uint32_t)hh[3] << 16));
      *reinterpret_cast<uint2*>(out_lo + o) = make_uint2((uint32_t)ll[0] | ((uint32_t)ll[1] << 16), (uint32_t)ll[2] | ((uint32_t)ll[3] << 16));
    }
  }
}
}  // namespace

extern "C" int nsac_row_op(const float* x, int ldx, const float* y, int ldy, const float* gamma, const float* beta, float eps,
                           int do_ln, const float* pos, int T, float* sum_out, int ld_sum, float* t_out, int ld_t, void* t_hi,
                           void* t_lo, int ld_tp, void* p_hi, void* p_lo, int ld_pp, int rows, int C, void* stream) {
  NSAC_REQUIRE(x && (sum_out || t_out || t_hi || p_hi), "nsac_row_op: null input / no output requested");
  NSAC_REQUIRE(rows >= 0 && C >= 32 && C % 32 == 0 && C <= 32 * ROW_MAX_PER_LANE, "nsac_row_op: C must be a multiple of 32, <= 1024 (got %d)", C);
  NSAC_REQUIRE(!do_ln || (gamma && beta), "nsac_row_op: LayerNorm needs gamma and beta");
  NSAC_REQUIRE((!t_hi || t_lo) && (!p_hi || (p_lo && pos && T >= 1)), "nsac_row_op: plane outputs need both planes (and pos / T)");
  if (rows == 0) return NSAC_OK;
  row_op_kernel<<<nsac_cdiv(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, y, ldy, gamma, beta, eps, do_ln, pos, T, sum_out, ld_sum, t_out, ld_t, static_cast<uint16_t*>(t_hi),
      static_cast<uint16_t*>(t_lo), ld_tp, static_cast<uint16_t*>(p_hi), static_cast<uint16_t*>(p_lo), ld_pp, rows, C);
  NSAC_CHECK_LAUNCH("nsac_row_op");
  return NSAC_OK;
}

extern "C" int nsac_attention_tiled(const float* q, int ldq, const float* k, const float* v, int ldkv, float* out, int ldo,
                                    void* out_hi, void* out_lo, int ld_split, int B, int L, int S, int H, int D, void* stream) {
  NSAC_REQUIRE(q && k && v && (out || (out_hi && out_lo)), "nsac_attention_tiled: null pointer");
  NSAC_REQUIRE(D == AT_D, "nsac_attention_tiled: head dim must be 32 (got %d)", D);
  NSAC_REQUIRE(H >= 1 && L >= 0 && S >= 1 && S <= 32 * AT_KPL, "nsac_attention_tiled: S must be in [1, %d] (got %d)", 32 * AT_KPL, S);
  NSAC_REQUIRE(ldq % 4 == 0 && ldkv >= H * D && (!out || ldo % 4 == 0), "nsac_attention_tiled: row strides must be multiples of 4");
  if (B == 0 || L == 0) return NSAC_OK;
  const int SP = (S + 31) / 32 * 32;
  const size_t smem = ((size_t)AT_D * SP * 2 + AT_D + AT_D * AT_QB + (size_t)AT_GROUPS * SP * 8 + 4 * AT_QB + AT_QB * AT_D) * sizeof(float);
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    NSAC_CUDA(cudaFuncSetAttribute(attention_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  attention_tiled_kernel<<<B * H, AT_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      q, ldq, k, v, ldkv, out, ldo, static_cast<uint16_t*>(out_hi), static_cast<uint16_t*>(out_lo), ld_split, L, S, SP, H);
  NSAC_CHECK_LAUNCH("nsac_attention_tiled");
  return NSAC_OK;
}

extern "C" int nsac_upsample2x_relu_add(const float* a, const float* b, int N, int h, int w, int C, float* out, void* out_hi,
                                        void* out_lo, void* stream) {
  NSAC_REQUIRE(a && b && (out || (out_hi && out_lo)), "nsac_upsample2x_relu_add: null pointer");
  NSAC_REQUIRE(N >= 0 && h >= 1 && w >= 1 && C >= 4 && C % 4 == 0, "nsac_upsample2x_relu_add: bad shape (C %% 4 == 0)");
  if (N == 0) return NSAC_OK;
  const size_t total = (size_t)N * 4 * h * w * (C / 4);
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)148 * 16) blocks = (size_t)148 * 16;
  upsample2x_relu_add_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, b, N, h, w, C, out, static_cast<uint16_t*>(out_hi), static_cast<uint16_t*>(out_lo));
  NSAC_CHECK_LAUNCH("nsac_upsample2x_relu_add");
  return NSAC_OK;
}
