// Device-side WOSAC post-processing: the step right after the rollout loop.
// Reference: data_modules/wosac_post_processing.py:31-64 (`_filter_futures`: keep the n_keep joint futures with the
// fewest role-weighted collision / road-edge violations) and :66-75 (`forward`: scene-centric -> global frame,
// transform_utils.py:160-171 torch_pos2global, :215-225 torch_rad2global).
#include "common.cuh"

namespace {

// score[sc, k] = sum_a role[sc, a] * any_{t >= t0} col[(sc K + k), a, t] + w * (same for road edge)
// (wosac_post_processing.py:48-58). One CTA per joint future, one warp per agent at a time.
__global__ void __launch_bounds__(256)
future_score_kernel(const uint8_t* __restrict__ col, const uint8_t* __restrict__ edge, const uint8_t* __restrict__ role,
                    int K, int A, int T, int t0, float w_edge, float* __restrict__ score) {
  const int bk = blockIdx.x, sc = bk / K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ int s_nc[8], s_ne[8];
  int nc = 0, ne = 0;
  for (int a = warp; a < A; a += 8) {
    if (!role[(size_t)sc * A + a]) continue;  // warp-uniform
    const uint8_t* c = col + ((size_t)bk * A + a) * T;
    const uint8_t* e = edge + ((size_t)bk * A + a) * T;
    bool hc = false, he = false;
    for (int t = t0 + lane; t < T; t += 32) { hc |= c[t] != 0; he |= e[t] != 0; }
    nc += __any_sync(TB_FULL_MASK, hc) ? 1 : 0;
    ne += __any_sync(TB_FULL_MASK, he) ? 1 : 0;
  }
  if (lane == 0) { s_nc[warp] = nc; s_ne[warp] = ne; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tc = 0, te = 0;
    for (int i = 0; i < 8; ++i) { tc += s_nc[i]; te += s_ne[i]; }
    // the reference sums 0/1 floats (exact integers) and forms collided + run_road_edge * w (:54-58)
    score[bk] = __fadd_rn((float)tc, __fmul_rn((float)te, w_edge));
  }
}

// n_keep smallest of K scores per scene, ties by lower index (torch.topk(sorted=False) leaves the order and the
// tie-break unspecified: wosac_post_processing.py:61); output in ascending (score, index) order. K <= 1024.
__global__ void __launch_bounds__(128)
future_select_kernel(const float* __restrict__ score, int K, int n_keep, int32_t* __restrict__ sel) {
  const int sc = blockIdx.x;
  extern __shared__ float s_sc[];
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_sc[k] = score[(size_t)sc * K + k];
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float v = s_sc[k];
    int rank = 0;
    for (int j = 0; j < K; ++j) rank += (s_sc[j] < v) || (s_sc[j] == v && j < k);
    if (rank < n_keep) sel[(size_t)sc * n_keep + rank] = k;
  }
}

// out_pos[sc, i, a, t, :] = R(yaw_sc) pose_xy[(sc K + sel[sc,i]), a, t0 + t] + center_sc ; out_yaw = wrap(yaw + yaw_sc)
// with wrap(x) = (x + pi) mod 2 pi - pi (python modulo: result in [-pi, pi), transform_utils.py:9-11).
__global__ void __launch_bounds__(256)
traj_global_kernel(const float* __restrict__ pose, const int32_t* __restrict__ sel, const float* __restrict__ center,
                   const float* __restrict__ yaw0, int K, int n_keep, int A, int T, int t0, size_t total,
                   float* __restrict__ out_pos, float* __restrict__ out_yaw) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Tf = T - t0;
  const int t = (int)(i % Tf);
  const int a = (int)((i / Tf) % A);
  const int f = (int)((i / ((size_t)Tf * A)) % n_keep);
  const int sc = (int)(i / ((size_t)Tf * A * n_keep));
  const int k = sel ? sel[(size_t)sc * n_keep + f] : f;
  const float* p = pose + (((size_t)(sc * K + k) * A + a) * T + t0 + t) * 3;
  const float th = yaw0[sc];
  const float c = cosf(th), s = sinf(th);
  // torch.matmul(pos, rot^T) with rot = [[c,-s],[s,c]]: x' = x c - y s, y' = x s + y c  (then + center)
  out_pos[i * 2 + 0] = __fadd_rn(__fadd_rn(__fmul_rn(p[0], c), __fmul_rn(p[1], -s)), center[sc * 2 + 0]);
  out_pos[i * 2 + 1] = __fadd_rn(__fadd_rn(__fmul_rn(p[0], s), __fmul_rn(p[1], c)), center[sc * 2 + 1]);
  const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
  float y = __fadd_rn(__fadd_rn(p[2], th), pi);
  float m = fmodf(y, two_pi);
  if (m < 0.f) m = __fadd_rn(m, two_pi);  // python-style modulo (torch.remainder)
  out_yaw[i] = __fadd_rn(m, -pi);
}

}  // namespace

extern "C" int tb_future_filter(const uint8_t* collided, const uint8_t* run_road_edge, const uint8_t* role_any, int n_sc,
                                int K, int A, int T, int t0, float w_road_edge, int n_keep, float* score,
                                int32_t* sel, void* stream) {
  if (!collided || !run_road_edge || !role_any || !score || !sel) return TB_ERR_NULL;
  if (n_sc <= 0 || K <= 0 || A <= 0 || T <= 0 || t0 < 0 || t0 >= T || n_keep <= 0 || n_keep > K) return TB_ERR_BAD_SHAPE;
  if (K > 1024) return TB_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  future_score_kernel<<<n_sc * K, 256, 0, st>>>(collided, run_road_edge, role_any, K, A, T, t0, w_road_edge, score);
  TB_CHECK_LAUNCH();
  future_select_kernel<<<n_sc, 128, K * sizeof(float), st>>>(score, K, n_keep, sel);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_traj_global(const float* pose, const int32_t* sel, const float* center, const float* yaw, int n_sc,
                              int K, int n_keep, int A, int T, int t0, float* out_pos, float* out_yaw, void* stream) {
  if (!pose || !center || !yaw || !out_pos || !out_yaw) return TB_ERR_NULL;
  if (n_sc <= 0 || K <= 0 || A <= 0 || T <= 0 || t0 < 0 || t0 >= T || n_keep <= 0 || n_keep > K) return TB_ERR_BAD_SHAPE;
  if (!sel && n_keep != K) return TB_ERR_NULL;
  const size_t total = (size_t)n_sc * n_keep * A * (T - t0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  traj_global_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pose, sel, center, yaw, K, n_keep, A, T, t0, total,
                                                                    out_pos, out_yaw);
  TB_CHECK_LAUNCH();
  return TB_OK;
}
