// K6 — geo sequences without the host round-trip of torch.nonzero:
//   get_pred_geo_sequence x3 (camera_head.py:1352-1425 via :513-517, :555-567), sig_seq (:568-569) and the
//   (n0*sig, d0*sig, n1, d1) 8-vector that feeds the geo encoder (:937-957).
// One CTA per pair: warp 0 compacts the assignment matrix in row-major (torch.nonzero) order with
// ballot/popc, then one thread per matched plane pair gathers, warps and writes all five outputs.
#include "common.cuh"

namespace {

__global__ void geo_sequence_kernel(const float* __restrict__ planes1, const float* __restrict__ planes2,
                                    const float* __restrict__ assign, const int32_t* __restrict__ hyp_pairs,
                                    int H, const float* __restrict__ t0, const float* __restrict__ q0, int n1,
                                    int n2, int NQ, float* __restrict__ geo_local, float* __restrict__ geo_global,
                                    float* __restrict__ sig_out, float* __restrict__ geo8,
                                    int32_t* __restrict__ matched_num, int32_t* __restrict__ pair_idx) {
  extern __shared__ int32_t pairs[];  // [NQ][2]
  __shared__ int s_m;
  const int b = blockIdx.x, tid = threadIdx.x;
  if (hyp_pairs) {
    for (int k = tid; k < H; k += blockDim.x) {
      pairs[2 * k] = min(max(hyp_pairs[2 * k], 0), n1 - 1);          // caller-supplied indices: clamped, never out of bounds
      pairs[2 * k + 1] = min(max(hyp_pairs[2 * k + 1], 0), n2 - 1);
    }
    if (tid == 0) s_m = H;
  } else if (tid < 32) {
    const float* A = assign + (size_t)b * n1 * n2;
    int base = 0;
    const int total = n1 * n2;
    for (int e0 = 0; e0 < total; e0 += 32) {
      const int e = e0 + tid;
      const bool nz = (e < total) && (A[e] != 0.f);
      const unsigned bal = __ballot_sync(NSAC_FULL_MASK, nz);
      const int pos = base + __popc(bal & ((1u << tid) - 1u));
      if (nz && pos < NQ) {
        pairs[2 * pos] = e / n2;
        pairs[2 * pos + 1] = e % n2;
      }
      base += __popc(bal);
    }
    if (tid == 0) s_m = base < NQ ? base : NQ;
  }
  __syncthreads();
  const int m = s_m;
  if (tid == 0) matched_num[b] = m;
  const float* P1 = planes1 + (size_t)b * n1 * 3;
  const float* P2 = planes2 + (size_t)b * n2 * 3;
  const float* t = t0 + (size_t)b * 3;
  const float* q = q0 + (size_t)b * 4;
  const Mat3 R = quat_to_rot(q[0], q[1], q[2], q[3]);
  for (int k = tid; k < NQ; k += blockDim.x) {
    const size_t row = (size_t)b * NQ + k;
    float l[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gl[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float sg = 1.f;  // padded rows: (0 * 0 >= 0) -> +1
    int pi = -1, pj = -1;
    if (k < m) {
      pi = pairs[2 * k];
      pj = pairs[2 * k + 1];
      l[0] = P1[pi * 3]; l[1] = P1[pi * 3 + 1]; l[2] = P1[pi * 3 + 2];
      l[3] = P2[pj * 3]; l[4] = P2[pj * 3 + 1]; l[5] = P2[pj * 3 + 2];
      warp_plane(R, t[0], t[1], t[2], l[0], l[1], l[2], gl[0], gl[1], gl[2]);
      gl[3] = l[3]; gl[4] = -l[4]; gl[5] = -l[5];
      float ax, ay, az;
      warp_plane(R, 0.f, 0.f, 0.f, l[0], l[1], l[2], ax, ay, az);
      sg = (gl[0] * ax >= 0.f) ? 1.f : -1.f;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      geo_local[row * 6 + c] = l[c];
      geo_global[row * 6 + c] = gl[c];
    }
    sig_out[row] = sg;
    pair_idx[row * 2] = pi;
    pair_idx[row * 2 + 1] = pj;
    // (n0*sig, d0*sig, n1, d1), camera_head.py:937-957
    const float d0 = sqrtf(gl[0] * gl[0] + gl[1] * gl[1] + gl[2] * gl[2]);
    const float d1 = sqrtf(gl[3] * gl[3] + gl[4] * gl[4] + gl[5] * gl[5]);
    float* o = geo8 + row * 8;
    o[0] = gl[0] / (d0 + 1e-10f) * sg;
    o[1] = gl[1] / (d0 + 1e-10f) * sg;
    o[2] = gl[2] / (d0 + 1e-10f) * sg;
    o[3] = d0 * sg;
    o[4] = gl[3] / (d1 + 1e-10f);
    o[5] = gl[4] / (d1 + 1e-10f);
    o[6] = gl[5] / (d1 + 1e-10f);
    o[7] = d1;
  }
}
}  // namespace

extern "C" int nsac_geo_sequence(const float* planes1, const float* planes2, const float* assign,
                                 const int32_t* hyp_pairs, int H, const float* t0, const float* q0, int B,
                                 int n1, int n2, int NQ, float* geo_local, float* geo_global, float* sig,
                                 float* geo8, int32_t* matched_num, int32_t* pair_idx, void* stream) {
  NSAC_REQUIRE(planes1 && planes2 && t0 && q0 && geo_local && geo_global && sig && geo8 && matched_num && pair_idx,
               "nsac_geo_sequence: null pointer");
  NSAC_REQUIRE(assign || hyp_pairs, "nsac_geo_sequence: need an assignment matrix or an explicit pair list");
  NSAC_REQUIRE(B >= 0 && n1 >= 1 && n2 >= 1 && NQ >= 1, "nsac_geo_sequence: bad shape");
  NSAC_REQUIRE(!hyp_pairs || (H >= 0 && H <= NQ), "nsac_geo_sequence: H=%d exceeds NUM_OBJECT_QUERIES=%d", H, NQ);
  if (B == 0) return NSAC_OK;
  const size_t smem = sizeof(int32_t) * 2 * (size_t)NQ;
  NSAC_REQUIRE(smem <= 48 * 1024, "nsac_geo_sequence: NQ=%d too large", NQ);
  geo_sequence_kernel<<<B, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      planes1, planes2, assign, hyp_pairs, H, t0, q0, n1, n2, NQ, geo_local, geo_global, sig, geo8, matched_num,
      pair_idx);
  NSAC_CHECK_LAUNCH("nsac_geo_sequence");
  return NSAC_OK;
}
