// Camera-pose evaluation on the device (evaluation/mp3d_evaluation.py:382-425, angle_error_vec :463-465).
//   per pair:   T err = |t - t_gt|_2 ,  R err = 2 acos(clip(|q . q_gt|, -1, 1)) * 180 / pi
//   reduction:  counts below the reference's thresholds (1.0 / 0.5 / 0.2 m, 30 / 15 / 10 deg) and the two error sums.
// One CTA, fixed summation order (thread-strided partials, then a tree): bit-reproducible.  The median needs the sorted
// errors and stays with the caller (nopesac_b200/evaluation.py sorts the per-pair errors on the device).
#include "common.cuh"

namespace {
constexpr int EVAL_THREADS = 256;

__global__ void __launch_bounds__(EVAL_THREADS)
camera_errors_kernel(const float* __restrict__ pose, int ldpose, const float* __restrict__ gt_tran, const float* __restrict__ gt_rot,
                     int B, float* __restrict__ err_t, float* __restrict__ err_r, float* __restrict__ stats) {
  __shared__ double red[EVAL_THREADS];
  __shared__ int redi[EVAL_THREADS];
  const int tid = threadIdx.x;
  double sum_t = 0.0, sum_r = 0.0;
  int cnt[6] = {0, 0, 0, 0, 0, 0};
  for (int b = tid; b < B; b += EVAL_THREADS) {
    const float* p = pose + (size_t)b * ldpose;          // (t[3], q[4], ...)
    const float dx = gt_tran[b * 3 + 0] - p[0], dy = gt_tran[b * 3 + 1] - p[1], dz = gt_tran[b * 3 + 2] - p[2];
    const float et = sqrtf(dx * dx + dy * dy + dz * dz);
    float dot = p[3] * gt_rot[b * 4 + 0] + p[4] * gt_rot[b * 4 + 1] + p[5] * gt_rot[b * 4 + 2] + p[6] * gt_rot[b * 4 + 3];
    dot = fminf(fmaxf(fabsf(dot), -1.f), 1.f);
    const float er = 2.f * acosf(dot) * 180.f / 3.14159265358979323846f;
    err_t[b] = et;
    err_r[b] = er;
    sum_t += (double)et;
    sum_r += (double)er;
    cnt[0] += et < 1.0f; cnt[1] += et < 0.5f; cnt[2] += et < 0.2f;
    cnt[3] += er < 30.f; cnt[4] += er < 15.f; cnt[5] += er < 10.f;
  }
  auto block_sum_d = [&](double v) {
    __syncthreads();
    red[tid] = v;
    __syncthreads();
    for (int s = EVAL_THREADS / 2; s > 0; s >>= 1) {
      if (tid < s) red[tid] += red[tid + s];
      __syncthreads();
    }
    return red[0];
  };
  auto block_sum_i = [&](int v) {
    __syncthreads();
    redi[tid] = v;
    __syncthreads();
    for (int s = EVAL_THREADS / 2; s > 0; s >>= 1) {
      if (tid < s) redi[tid] += redi[tid + s];
      __syncthreads();
    }
    return redi[0];
  };
  const double st = block_sum_d(sum_t), sr = block_sum_d(sum_r);
  int c[6];
  for (int i = 0; i < 6; ++i) c[i] = block_sum_i(cnt[i]);
  if (tid == 0) {
    stats[0] = (float)(st / (double)B);        // T mean err
    stats[1] = (float)(sr / (double)B);        // R mean err
    for (int i = 0; i < 6; ++i) stats[2 + i] = (float)c[i];   // counts: T<1, T<0.5, T<0.2, R<30, R<15, R<10
  }
}
}  // namespace

extern "C" int nsac_camera_errors(const float* pose, int ldpose, const float* gt_tran, const float* gt_rot, int B, float* err_t,
                                  float* err_r, float* stats, void* stream) {
  NSAC_REQUIRE(pose && gt_tran && gt_rot && err_t && err_r && stats, "nsac_camera_errors: null pointer");
  NSAC_REQUIRE(B >= 1 && ldpose >= 7, "nsac_camera_errors: bad shape B=%d ldpose=%d", B, ldpose);
  camera_errors_kernel<<<1, EVAL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(pose, ldpose, gt_tran, gt_rot, B, err_t, err_r, stats);
  NSAC_CHECK_LAUNCH("nsac_camera_errors");
  return NSAC_OK;
}
