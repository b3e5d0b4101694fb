// K1 — the pixel pose-regression network (camera_head.py:642-683, camera_modules.py:246-348) around the tensor-core
// engine.  Activations live in NHWC ([N*H*W, C] rows) so that every convolution is a GEMM whose output rows are
// already the next layer's input rows; the 16-bit hi/lo operand planes are produced by whichever kernel writes an
// activation.  This file holds the non-GEMM pieces:
//   nsac_nchw_to_planes     backbone feature maps [N,C,H,W] fp32 -> NHWC planes (tiled transpose + split)
//   nsac_groupnorm_nhwc     GroupNorm(32) (+ReLU) (+ nearest-2x-upsampled skip add) -> fp32 and/or planes
//   nsac_maxpool2_planes    2x2 max-pool of an fp32 NHWC map -> planes
//   nsac_corr_softmax       300x300 feature correlation + softmax over the view-2 positions -> planes
//   nsac_im2col_planes      explicit 3x3 im2col for the few small / strided convolutions
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

// Sticky overflow flag of this translation unit (see nsac_plane_overflow in gemm_tc.cu): a finite value that does not fit an
// fp16 plane (|x| > 65504) is recorded instead of silently becoming inf.
__device__ unsigned int g_pixel_plane_overflow = 0u;

__device__ __forceinline__ void split16(float x, int fmt, uint16_t& hi, uint16_t& lo) {
  if (fmt == NSAC_SPLIT_F16) {
    if (fabsf(x) > 65504.f && fabsf(x) < INFINITY) atomicOr(&g_pixel_plane_overflow, 1u);
    const __half h = __float2half_rn(x);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn(x - __half2float(h)));
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(h)));
  }
}

// ------------------------------------------------------------------------------------------------ NCHW -> NHWC planes
// x [N, C, HW] -> hi/lo [N*HW, C]; 32x32 tiles through shared memory, both sides coalesced
__global__ void nchw_to_planes_kernel(const float* __restrict__ x, int C, int HW, int fmt, uint16_t* __restrict__ hi,
                                      uint16_t* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
  const float* src = x + (size_t)n * C * HW;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, p = p0 + tx;
    tile[ty + i][tx] = (c < C && p < HW) ? src[(size_t)c * HW + p] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int p = p0 + ty + i, c = c0 + tx;
    if (p < HW && c < C) {
      uint16_t h, l;
      split16(tile[tx][ty + i], fmt, h, l);
      const size_t o = ((size_t)n * HW + p) * C + c;
      hi[o] = h;
      lo[o] = l;
    }
  }
}

// 64 channels x 64 pixels per CTA, 16-byte loads along the pixels (needs HW % 4 == 0 and a 16-byte aligned base): four
// times the bytes in flight per thread of the 32x32 version; a warp writes 64 contiguous bytes per plane and instruction.
__global__ void __launch_bounds__(256)
nchw_to_planes64_kernel(const float* __restrict__ x, int C, int HW, int fmt, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  __shared__ float tile[64][65];
  const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64, tid = threadIdx.x;
  const float* src = x + (size_t)n * C * HW;
  {
    const int px = (tid & 15) * 4, cy = tid >> 4;       // 16 float4 per channel row, 16 channel rows per pass
#pragma unroll
    for (int i = 0; i < 64; i += 16) {
      const int c = c0 + cy + i, p = p0 + px;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C && p < HW) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)c * HW + p));
      tile[cy + i][px] = v.x; tile[cy + i][px + 1] = v.y; tile[cy + i][px + 2] = v.z; tile[cy + i][px + 3] = v.w;
    }
  }
  __syncthreads();
  const int lane = tid & 31, w = tid >> 5;
#pragma unroll
  for (int i = 0; i < 64; i += 8) {
    const int pp = w + i, p = p0 + pp;
    if (p < HW) {
      const size_t o = ((size_t)n * HW + p) * C + c0;
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        const int c = lane + 32 * hlf;
        if (c0 + c < C) {
          uint16_t h, l;
          split16(tile[c][pp], fmt, h, l);
          hi[o + c] = h;
          lo[o + c] = l;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ GroupNorm (NHWC)
// one CTA per (image, 32-channel slab = 8 groups of 4 channels when C = 128): exact two-sweep statistics
// (mean, then centred sum of squares), then normalise (+ReLU) (+ skip[n, y/2, x/2, c]) and write fp32 / planes.
__global__ void __launch_bounds__(256)
groupnorm_nhwc_kernel(const float* __restrict__ x, int H, int W, int C, int G, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, int relu, const float* __restrict__ skip, int fmt,
                      float* __restrict__ out_f32, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
  const int n = blockIdx.y, g = blockIdx.x, cpg = C / G;          // one CTA per (group, image)
  const int HW = H * W, tid = threadIdx.x;
  const float* xi = x + (size_t)n * HW * C + g * cpg;
  __shared__ float red[32];
  __shared__ float s_mean, s_rstd;
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += red[w];
    return r;
  };
  const int total = HW * cpg;
  float s = 0.f;
  for (int i = tid; i < total; i += blockDim.x) s += xi[(size_t)(i / cpg) * C + (i % cpg)];
  const float mean = block_sum(s) / (float)total;
  float v = 0.f;
  for (int i = tid; i < total; i += blockDim.x) {
    const float d = xi[(size_t)(i / cpg) * C + (i % cpg)] - mean;
    v = fmaf(d, d, v);
  }
  const float var = block_sum(v) / (float)total;
  if (tid == 0) { s_mean = mean; s_rstd = rsqrtf(var + eps); }
  __syncthreads();
  const float rstd = s_rstd;
  const int H2 = H >> 1, W2 = W >> 1;
  for (int i = tid; i < total; i += blockDim.x) {
    const int pix = i / cpg, c = g * cpg + (i % cpg);
    const size_t o = ((size_t)n * HW + pix) * C + c;
    float y = (x[o] - mean) * rstd * gamma[c] + beta[c];
    if (relu) y = fmaxf(y, 0.f);
    if (skip) {   // F.interpolate(mode="nearest") from the (H/2, W/2) level: src = floor(dst / 2)
      const int yy = pix / W, xx = pix % W;
      y += skip[(((size_t)n * H2 + (yy >> 1)) * W2 + (xx >> 1)) * C + c];
    }
    if (out_f32) out_f32[o] = y;
    if (out_hi) {
      uint16_t h, l;
      split16(y, fmt, h, l);
      out_hi[o] = h;
      out_lo[o] = l;
    }
  }
}

// ---- GroupNorm with 4 channels per group (the pixel decoder: GroupNorm(32, 128)), coalesced ----------------------------------
// The kernel above reads 16 useful bytes per 32-byte sector three times over (1.6 ms of a 17 ms bench step).  Here a
// thread owns one float4 = one group of one pixel, so a warp reads 512 contiguous bytes; two launches:
//   gn4_stats_kernel   grid (pixel chunks, images): shifted sums  S1 = sum(x - K), S2 = sum((x - K)^2)  per (image, chunk,
//                      group), K = the group's first value of the image (the shift keeps S2 - S1^2/n well conditioned)
//   gn4_apply_kernel   same grid: reduces the image's chunk partials in a fixed order (deterministic), then normalises
//                      (+ReLU) (+ skip[n, y/2, x/2, c]) its pixel chunk and writes fp32 / planes with 16-byte accesses
constexpr int GN4_PIX = 256;        // pixels per CTA
__global__ void __launch_bounds__(256)
gn4_stats_kernel(const float* __restrict__ x, int HW, int C, float* __restrict__ part) {
  const int n = blockIdx.y, chunk = blockIdx.x, G = C >> 2, tid = threadIdx.x;
  const int lanes = 256 / G;                       // pixel lanes per CTA (C = 128: 8)
  const int g = tid % G, pl = tid / G;
  extern __shared__ float sm[];                    // [lanes][G][2]
  const float* xi = x + (size_t)n * HW * C;
  const float K = xi[g * 4];
  const int p0 = chunk * GN4_PIX, p1 = min(HW, p0 + GN4_PIX);
  float s1 = 0.f, s2 = 0.f;
  if (pl < lanes) {
    for (int p = p0 + pl; p < p1; p += lanes) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xi + (size_t)p * C + g * 4));
      const float a = v.x - K, b = v.y - K, c = v.z - K, d = v.w - K;
      s1 += (a + b) + (c + d);
      s2 = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(d, d, s2))));
    }
    sm[(pl * G + g) * 2] = s1;
    sm[(pl * G + g) * 2 + 1] = s2;
  }
  __syncthreads();
  if (tid < G) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) { a += sm[(l * G + tid) * 2]; b += sm[(l * G + tid) * 2 + 1]; }
    float* o = part + (((size_t)n * gridDim.x + chunk) * G + tid) * 2;
    o[0] = a; o[1] = b;
  }
}

__global__ void __launch_bounds__(256)
gn4_apply_kernel(const float* __restrict__ x, int H, int W, int C, const float* __restrict__ part, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, int relu, const float* __restrict__ skip, int fmt,
                 float* __restrict__ out_f32, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
  const int n = blockIdx.y, chunk = blockIdx.x, G = C >> 2, tid = threadIdx.x, HW = H * W;
  const int lanes = 256 / G, g = tid % G, pl = tid / G;
  extern __shared__ float sm[];                    // [G][2]: mean, rstd
  const float* xi = x + (size_t)n * HW * C;
  if (tid < G) {
    float a = 0.f, b = 0.f;
    for (int c = 0; c < (int)gridDim.x; ++c) {     // fixed order: bit-reproducible
      const float* q = part + (((size_t)n * gridDim.x + c) * G + tid) * 2;
      a += q[0]; b += q[1];
    }
    const float cnt = (float)HW * 4.f, m1 = a / cnt;
    const float var = fmaxf(b / cnt - m1 * m1, 0.f);
    sm[tid * 2] = xi[tid * 4] + m1;
    sm[tid * 2 + 1] = rsqrtf(var + eps);
  }
  __syncthreads();
  if (pl >= lanes) return;
  const float mean = sm[g * 2], rstd = sm[g * 2 + 1];
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + g * 4)), be = __ldg(reinterpret_cast<const float4*>(beta + g * 4));
  const int H2 = H >> 1, W2 = W >> 1;
  const int p0 = chunk * GN4_PIX, p1 = min(HW, p0 + GN4_PIX);
  for (int p = p0 + pl; p < p1; p += lanes) {
    const size_t o = ((size_t)n * HW + p) * C + g * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + o));
    float y[4] = {(v.x - mean) * rstd * ga.x + be.x, (v.y - mean) * rstd * ga.y + be.y, (v.z - mean) * rstd * ga.z + be.z,
                  (v.w - mean) * rstd * ga.w + be.w};
    if (relu) {
#pragma unroll
      for (int k = 0; k < 4; ++k) y[k] = fmaxf(y[k], 0.f);
    }
    if (skip) {   // F.interpolate(mode="nearest") from the (H/2, W/2) level: src = floor(dst / 2)
      const int yy = p / W, xx = p % W;
      const float4 sk = __ldg(reinterpret_cast<const float4*>(skip + (((size_t)n * H2 + (yy >> 1)) * W2 + (xx >> 1)) * C + g * 4));
      y[0] += sk.x; y[1] += sk.y; y[2] += sk.z; y[3] += sk.w;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(y[0], y[1], y[2], y[3]);
    if (out_hi) {
      uint16_t h[4], l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) split16(y[k], fmt, h[k], l[k]);
      *reinterpret_cast<uint2*>(out_hi + o) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
      *reinterpret_cast<uint2*>(out_lo + o) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
  }
}

// ------------------------------------------------------------------------------------------------ 2x2 max-pool
// VEC = 4: one thread = 4 channels of one output pixel (16-byte loads, 8-byte plane stores); VEC = 1: any C / alignment
template <int VEC>
__global__ void maxpool2_planes_kernel(const float* __restrict__ x, int N, int H, int W, int C, int fmt,
                                       uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  const int Ho = H >> 1, Wo = W >> 1, Cv = C / VEC;
  const size_t total = (size_t)N * Ho * Wo * Cv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cv) * VEC;
    size_t t = i / Cv;
    const int xo = (int)(t % Wo); t /= Wo;
    const int yo = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const float* p = x + (((size_t)n * H + 2 * yo) * W + 2 * xo) * C + c;
    const size_t o = ((((size_t)n * Ho + yo) * Wo + xo)) * C + c;
    if (VEC == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + C));
      const float4 d = __ldg(reinterpret_cast<const float4*>(p + (size_t)W * C)), e = __ldg(reinterpret_cast<const float4*>(p + (size_t)W * C + C));
      const float m[4] = {fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y)),
                          fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w))};
      uint16_t h[4], l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) split16(m[k], fmt, h[k], l[k]);
      *reinterpret_cast<uint2*>(hi + o) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
      *reinterpret_cast<uint2*>(lo + o) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    } else {
      const float m = fmaxf(fmaxf(p[0], p[C]), fmaxf(p[(size_t)W * C], p[(size_t)W * C + C]));
      uint16_t h, l;
      split16(m, fmt, h, l);
      hi[o] = h;
      lo[o] = l;
    }
  }
}

// ------------------------------------------------------------------------------------------------ correlation + softmax
// compute_corr_softmax (camera_head.py:1117-1133): corr[b, c2, p1] = f2[b, :, c2] . f1[b, :, p1] with the view-2
// positions enumerated w-major (c2 = w2 * H + h2), softmax over c2.  NHWC in, NHWC planes out:
// out[(b*HW + p1), c2], channels padded to Cp (multiple of 64) with zeros.  One CTA per (pair, 8 view-1 pixels).
constexpr int CORR_PIX = 8;
__global__ void __launch_bounds__(512)
corr_softmax_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int H, int W, int C, int Cp, int fmt,
                    uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  extern __shared__ float sm[];
  const int HW = H * W;
  float* q = sm;                      // [CORR_PIX][C]
  float* sc = q + CORR_PIX * C;       // [CORR_PIX][HW]
  const int b = blockIdx.y, p0 = blockIdx.x * CORR_PIX, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* F1 = f1 + (size_t)b * HW * C;
  const float* F2 = f2 + (size_t)b * HW * C;
  for (int i = tid; i < CORR_PIX * C; i += blockDim.x) {
    const int pp = p0 + i / C;
    q[i] = pp < HW ? F1[(size_t)pp * C + (i % C)] : 0.f;
  }
  __syncthreads();
  // every thread owns view-2 positions p2 = tid, tid + 256, ...; dot with the 8 query pixels
  for (int p2 = tid; p2 < HW; p2 += blockDim.x) {
    float acc[CORR_PIX];
#pragma unroll
    for (int k = 0; k < CORR_PIX; ++k) acc[k] = 0.f;
    const float4* r = reinterpret_cast<const float4*>(F2 + (size_t)p2 * C);
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 v = __ldg(r + c4);
#pragma unroll
      for (int k = 0; k < CORR_PIX; ++k) {
        const float4 qq = *reinterpret_cast<const float4*>(q + k * C + c4 * 4);
        acc[k] = fmaf(v.x, qq.x, fmaf(v.y, qq.y, fmaf(v.z, qq.z, fmaf(v.w, qq.w, acc[k]))));
      }
    }
    const int h2 = p2 / W, w2 = p2 % W;
    const int c2 = w2 * H + h2;       // im_feature2.transpose(2, 3) -> w-major channel order
#pragma unroll
    for (int k = 0; k < CORR_PIX; ++k) sc[k * HW + c2] = acc[k];
  }
  __syncthreads();
  // softmax over the HW channels of each query pixel: warp w <-> pixel w
  if (warp < CORR_PIX && p0 + warp < HW) {
    float* s = sc + warp * HW;
    float mx = -INFINITY;
    for (int c = lane; c < HW; c += 32) mx = fmaxf(mx, s[c]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < HW; c += 32) {
      const float e = expf(s[c] - mx);
      s[c] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const size_t o = ((size_t)b * HW + p0 + warp) * Cp;
    for (int c = lane; c < Cp; c += 32) {
      uint16_t h, l;
      split16(c < HW ? s[c] / sum : 0.f, fmt, h, l);
      hi[o + c] = h;
      lo[o + c] = l;
    }
  }
}

// ------------------------------------------------------------------------------------------------ explicit im2col
// x fp32 NHWC [N,H,W,C] -> planes [N*Ho*Wo, Kp], K order (ky, kx, c), 3x3, pad 1, given stride; zero padded to Kp
__global__ void im2col3x3_planes_kernel(const float* __restrict__ x, int N, int H, int W, int C, int stride, int Ho, int Wo,
                                        int Kp, int fmt, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  const size_t total = (size_t)N * Ho * Wo * Kp;
  const int K = 9 * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    size_t t = i / Kp;
    const int xo = (int)(t % Wo); t /= Wo;
    const int yo = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float v = 0.f;
    if (k < K) {
      const int tap = k / C, c = k % C;
      const int yi = yo * stride + tap / 3 - 1, xi = xo * stride + tap % 3 - 1;
      if (yi >= 0 && yi < H && xi >= 0 && xi < W) v = x[(((size_t)n * H + yi) * W + xi) * C + c];
    }
    uint16_t h, l;
    split16(v, fmt, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}
}  // namespace

extern "C" int nsac_nchw_to_planes(const float* x, int N, int C, int HW, int fmt, void* hi, void* lo, void* stream) {
  NSAC_REQUIRE(x && hi && lo && N >= 0 && C >= 1 && HW >= 1, "nsac_nchw_to_planes: bad arguments");
  if (N == 0) return NSAC_OK;
  if (HW % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    dim3 grid64(nsac_cdiv(HW, 64), nsac_cdiv(C, 64), N);
    nchw_to_planes64_kernel<<<grid64, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, C, HW, fmt, static_cast<uint16_t*>(hi),
                                                                                    static_cast<uint16_t*>(lo));
    NSAC_CHECK_LAUNCH("nsac_nchw_to_planes");
    return NSAC_OK;
  }
  dim3 grid(nsac_cdiv(HW, 32), nsac_cdiv(C, 32), N), block(32, 8);
  nchw_to_planes_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(x, C, HW, fmt, static_cast<uint16_t*>(hi),
                                                                                 static_cast<uint16_t*>(lo));
  NSAC_CHECK_LAUNCH("nsac_nchw_to_planes");
  return NSAC_OK;
}

extern "C" size_t nsac_groupnorm_ws_bytes(int N, int H, int W, int G) {
  if (N < 0 || H < 1 || W < 1 || G < 1) return 0;
  return (size_t)N * nsac_cdiv(H * W, GN4_PIX) * G * 2 * sizeof(float);
}

extern "C" int nsac_groupnorm_nhwc(const float* x, int N, int H, int W, int C, int G, const float* gamma, const float* beta,
                                   float eps, int relu, const float* skip_half_res, int fmt, float* out_f32, void* out_hi,
                                   void* out_lo, void* stats_ws, size_t stats_ws_bytes, void* stream) {
  NSAC_REQUIRE(x && gamma && beta && (out_f32 || (out_hi && out_lo)), "nsac_groupnorm_nhwc: null pointer");
  NSAC_REQUIRE(N >= 0 && H >= 1 && W >= 1 && C >= 1 && G >= 1 && C % G == 0, "nsac_groupnorm_nhwc: bad shape");
  NSAC_REQUIRE(!skip_half_res || (H % 2 == 0 && W % 2 == 0), "nsac_groupnorm_nhwc: skip add needs even H, W");
  if (N == 0) return NSAC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                         reinterpret_cast<uintptr_t>(skip_half_res) | reinterpret_cast<uintptr_t>(out_f32)) & 15) == 0 &&
                       ((reinterpret_cast<uintptr_t>(out_hi) | reinterpret_cast<uintptr_t>(out_lo)) & 7) == 0;
  if (C == 4 * G && G <= 256 && 256 % G == 0 && aligned && stats_ws) {
    const int chunks = nsac_cdiv(H * W, GN4_PIX);
    NSAC_REQUIRE(stats_ws_bytes >= (size_t)N * chunks * G * 2 * sizeof(float), "nsac_groupnorm_nhwc: statistics workspace too small");
    dim3 grid4(chunks, N);
    gn4_stats_kernel<<<grid4, 256, (256 / G) * G * 2 * sizeof(float), st>>>(x, H * W, C, static_cast<float*>(stats_ws));
    NSAC_CHECK_LAUNCH("gn4_stats_kernel");
    gn4_apply_kernel<<<grid4, 256, G * 2 * sizeof(float), st>>>(x, H, W, C, static_cast<const float*>(stats_ws), gamma, beta, eps, relu,
                                                              skip_half_res, fmt, out_f32, static_cast<uint16_t*>(out_hi),
                                                              static_cast<uint16_t*>(out_lo));
    NSAC_CHECK_LAUNCH("gn4_apply_kernel");
    return NSAC_OK;
  }
  dim3 grid(G, N);
  groupnorm_nhwc_kernel<<<grid, 256, 0, st>>>(
      x, H, W, C, G, gamma, beta, eps, relu, skip_half_res, fmt, out_f32, static_cast<uint16_t*>(out_hi), static_cast<uint16_t*>(out_lo));
  NSAC_CHECK_LAUNCH("nsac_groupnorm_nhwc");
  return NSAC_OK;
}

extern "C" int nsac_maxpool2_planes(const float* x, int N, int H, int W, int C, int fmt, void* hi, void* lo, void* stream) {
  NSAC_REQUIRE(x && hi && lo && N >= 0 && H >= 2 && W >= 2 && C >= 1, "nsac_maxpool2_planes: bad arguments");
  if (N == 0) return NSAC_OK;
  const bool vec = C % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 7) == 0;
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (vec ? C / 4 : C);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (vec)
    maxpool2_planes_kernel<4><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, N, H, W, C, fmt, static_cast<uint16_t*>(hi),
                                                                                      static_cast<uint16_t*>(lo));
  else
    maxpool2_planes_kernel<1><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, N, H, W, C, fmt, static_cast<uint16_t*>(hi),
                                                                                      static_cast<uint16_t*>(lo));
  NSAC_CHECK_LAUNCH("nsac_maxpool2_planes");
  return NSAC_OK;
}

extern "C" int nsac_corr_softmax(const float* f1, const float* f2, int B, int H, int W, int C, int Cp, int fmt, void* hi, void* lo,
                                 void* stream) {
  NSAC_REQUIRE(f1 && f2 && hi && lo && B >= 0 && H >= 1 && W >= 1 && C % 4 == 0 && Cp >= H * W, "nsac_corr_softmax: bad arguments");
  if (B == 0) return NSAC_OK;
  const size_t smem = sizeof(float) * CORR_PIX * ((size_t)C + (size_t)H * W);
  NSAC_REQUIRE(smem <= 200 * 1024, "nsac_corr_softmax: feature map too large");
  if (smem > 48 * 1024) NSAC_CUDA(cudaFuncSetAttribute(corr_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(nsac_cdiv(H * W, CORR_PIX), B);
  // one thread per view-2 position where that fits (15 x 20 maps: 320 threads, a single pass instead of 256 threads + a ragged second one)
  int threads = (H * W + 31) / 32 * 32;
  threads = threads < 256 ? 256 : (threads > 512 ? 512 : threads);
  corr_softmax_kernel<<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(f1, f2, H, W, C, Cp, fmt, static_cast<uint16_t*>(hi),
                                                                                   static_cast<uint16_t*>(lo));
  NSAC_CHECK_LAUNCH("nsac_corr_softmax");
  return NSAC_OK;
}

extern "C" int nsac_im2col3x3_planes(const float* x, int N, int H, int W, int C, int stride, int Kp, int fmt, void* hi, void* lo,
                                     void* stream) {
  NSAC_REQUIRE(x && hi && lo && N >= 0 && H >= 1 && W >= 1 && C >= 1 && (stride == 1 || stride == 2) && Kp >= 9 * C,
               "nsac_im2col3x3_planes: bad arguments");
  if (N == 0) return NSAC_OK;
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const size_t total = (size_t)N * Ho * Wo * Kp;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  im2col3x3_planes_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, N, H, W, C, stride, Ho, Wo, Kp, fmt,
                                                                                  static_cast<uint16_t*>(hi), static_cast<uint16_t*>(lo));
  NSAC_CHECK_LAUNCH("nsac_im2col3x3_planes");
  return NSAC_OK;
}

// host access to this translation unit's overflow flag (used by nsac_plane_overflow)
unsigned int* nsac_pixel_overflow_ptr() {
  unsigned int* p = nullptr;
  cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_pixel_plane_overflow);
  return p;
}
