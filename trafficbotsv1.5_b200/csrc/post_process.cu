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

// ---- WOMD post-processing (data_modules/womd_post_processing.py:36-106, the configured path of
// configs/model/sim_agent.yaml:170-177): per (scene, agent) softmax of the joint-future log-probs (:48-53), top-k_pred
// (`traj_topk` :170-190, here in descending score order, ties by lower index), type-dependent ADE / FDE NMS (`mpa_nms`
// :75-106: in descending score order, a mode within the threshold of a currently higher-scored one drops to 1e-3),
// renormalisation, optional temperature (:66-67) and the 2 Hz gather (:69). One warp per (scene, agent).
constexpr int WOMD_MAXK = 8;     // k_pred
constexpr int WOMD_MAXF = 128;   // joint futures
__global__ void __launch_bounds__(128)
womd_post_kernel(const float* __restrict__ trajs, const float* __restrict__ scores, const uint8_t* __restrict__ ag_type,
                 int n_pairs, int K, int A, int T, int k_pred, int use_ade, int nms_on, float thr_veh, float thr_ped,
                 float thr_cyc, float temperature, int t_first, int t_stride, int n_out, float* __restrict__ out_trajs,
                 float* __restrict__ out_scores, int32_t* __restrict__ out_mode) {
  __shared__ float s_p[4][WOMD_MAXK];
  __shared__ int s_mode[4][WOMD_MAXK];
  __shared__ float s_dist[4][WOMD_MAXK][WOMD_MAXK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 4 + warp;
  if (pair >= n_pairs) return;  // warp-uniform
  const int sc = pair / A, a = pair - sc * A;
  const int ke = K < k_pred ? K : k_pred;

  // softmax over the K futures (lane l owns futures l, l + 32, ...)
  float v[WOMD_MAXF / 32];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < WOMD_MAXF / 32; ++i) {
    const int k = lane + 32 * i;
    v[i] = k < K ? (scores ? scores[((size_t)sc * K + k) * A + a] : 0.f) : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(TB_FULL_MASK, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < WOMD_MAXF / 32; ++i) {
    v[i] = lane + 32 * i < K ? expf(v[i] - mx) : -1.f;  // -1: never selected
    if (v[i] > 0.f) sum += v[i];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(TB_FULL_MASK, sum, o);
#pragma unroll
  for (int i = 0; i < WOMD_MAXF / 32; ++i)
    if (v[i] >= 0.f) v[i] = v[i] / sum;

  if (K > k_pred) {  // traj_topk: repeated arg-max, ties to the lower index, then renormalise over the kept modes
    float kept = 0.f;
    for (int m = 0; m < ke; ++m) {
      float bv = -1.f;
      int bi = 0x7fffffff;
#pragma unroll
      for (int i = 0; i < WOMD_MAXF / 32; ++i)
        if (v[i] > bv) { bv = v[i]; bi = lane + 32 * i; }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(TB_FULL_MASK, bv, o);
        const int oi = __shfl_xor_sync(TB_FULL_MASK, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
#pragma unroll
      for (int i = 0; i < WOMD_MAXF / 32; ++i)
        if (lane + 32 * i == bi) v[i] = -1.f;
      if (lane == 0) { s_p[warp][m] = bv; s_mode[warp][m] = bi; }
      kept += bv;
    }
    __syncwarp();
    if (lane < ke) s_p[warp][lane] = s_p[warp][lane] / kept;
  } else {
#pragma unroll
    for (int i = 0; i < WOMD_MAXF / 32; ++i) {
      const int k = lane + 32 * i;
      if (k < K) { s_p[warp][k] = v[i]; s_mode[warp][k] = k; }
    }
  }
  __syncwarp();

  if (nms_on) {
    // pairwise ADE (mean over the T steps of the xy distance) or FDE between the kept modes
    float acc[WOMD_MAXK * (WOMD_MAXK - 1) / 2];
#pragma unroll
    for (int i = 0; i < WOMD_MAXK * (WOMD_MAXK - 1) / 2; ++i) acc[i] = 0.f;
    const float* base[WOMD_MAXK];
#pragma unroll
    for (int m = 0; m < WOMD_MAXK; ++m)
      base[m] = trajs + (((size_t)sc * K + s_mode[warp][m < ke ? m : 0]) * A + a) * T * 3;
    for (int t = use_ade ? lane : T - 1 + lane * T; t < T; t += 32) {  // FDE: lane 0 visits the last step only
      float x[WOMD_MAXK], y[WOMD_MAXK];
#pragma unroll
      for (int m = 0; m < WOMD_MAXK; ++m) {
        x[m] = m < ke ? base[m][t * 3] : 0.f;
        y[m] = m < ke ? base[m][t * 3 + 1] : 0.f;
      }
      int q = 0;
#pragma unroll
      for (int m = 0; m < WOMD_MAXK; ++m)
#pragma unroll
        for (int n = m + 1; n < WOMD_MAXK; ++n, ++q) {
          const float dx = x[m] - x[n], dy = y[m] - y[n];
          acc[q] += sqrtf(dx * dx + dy * dy);
        }
    }
    int q = 0;
#pragma unroll
    for (int m = 0; m < WOMD_MAXK; ++m)
#pragma unroll
      for (int n = m + 1; n < WOMD_MAXK; ++n, ++q) {
        float d = acc[q];
#pragma unroll
        for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(TB_FULL_MASK, d, o);
        if (use_ade) d = d / (float)T;
        if (lane == 0) { s_dist[warp][m][n] = d; s_dist[warp][n][m] = d; }
      }
    __syncwarp();
    if (lane == 0) {
      const uint8_t* ty = ag_type + ((size_t)sc * A + a) * 3;
      const float thresh = (ty[0] ? thr_veh : 0.f) + (ty[1] ? thr_ped : 0.f) + (ty[2] ? thr_cyc : 0.f);
      float* p = s_p[warp];
      // visit the modes in descending (initial) score order; comparisons use the current scores (:100-105)
      int order[WOMD_MAXK];
      for (int m = 0; m < ke; ++m) order[m] = m;
      for (int i = 1; i < ke; ++i)  // insertion sort, stable: ties keep the lower index first
        for (int j = i; j > 0 && p[order[j]] > p[order[j - 1]]; --j) {
          const int tmp = order[j]; order[j] = order[j - 1]; order[j - 1] = tmp;
        }
      for (int i = 0; i < ke; ++i) {
        const int k = order[i];
        bool hit = false;
        for (int n = 0; n < ke; ++n) hit |= n != k && s_dist[warp][k][n] < thresh && p[n] > p[k];
        if (hit) p[k] = 1e-3f;
      }
      float tot = 0.f;
      for (int m = 0; m < ke; ++m) tot += p[m];
      for (int m = 0; m < ke; ++m) p[m] = p[m] / tot;
    }
    __syncwarp();
  }
  if (temperature > 0.f && lane == 0) {  // softmax(log(p) / temperature)
    float* p = s_p[warp];
    float l[WOMD_MAXK], lm = -INFINITY, tot = 0.f;
    for (int m = 0; m < ke; ++m) { l[m] = logf(p[m]) / temperature; lm = fmaxf(lm, l[m]); }
    for (int m = 0; m < ke; ++m) { l[m] = expf(l[m] - lm); tot += l[m]; }
    for (int m = 0; m < ke; ++m) p[m] = l[m] / tot;
  }
  __syncwarp();

  if (lane < ke) {
    out_scores[(size_t)pair * ke + lane] = s_p[warp][lane];
    if (out_mode) out_mode[(size_t)pair * ke + lane] = s_mode[warp][lane];
  }
  for (int e = lane; e < ke * n_out * 3; e += 32) {
    const int m = e / (n_out * 3), r = e - m * n_out * 3, j = r / 3, c = r - j * 3;
    out_trajs[(size_t)pair * ke * n_out * 3 + e] =
        trajs[((((size_t)sc * K + s_mode[warp][m]) * A + a) * T + (t_first + j * t_stride)) * 3 + c];
  }
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

extern "C" int tb_womd_post(const float* trajs, const float* scores, const uint8_t* ag_type, int n_sc, int K, int A, int T,
                            int k_pred, int use_ade, int nms_on, float thr_veh, float thr_ped, float thr_cyc,
                            float score_temperature, int t_first, int t_stride, int t_end, float* out_trajs,
                            float* out_scores, int32_t* out_mode, void* stream) {
  if (!trajs || !ag_type || !out_trajs || !out_scores) return TB_ERR_NULL;
  if (n_sc <= 0 || K <= 0 || A <= 0 || T <= 0 || k_pred <= 0 || t_first < 0 || t_stride <= 0) return TB_ERR_BAD_SHAPE;
  if (K > WOMD_MAXF || k_pred > WOMD_MAXK) return TB_ERR_UNSUPPORTED;
  const int end = t_end < T ? t_end : T;
  const int n_out = end > t_first ? (end - t_first + t_stride - 1) / t_stride : 0;
  if (n_out <= 0) return TB_ERR_BAD_SHAPE;
  const int n_pairs = n_sc * A;
  womd_post_kernel<<<(n_pairs + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      trajs, scores, ag_type, n_pairs, K, A, T, k_pred, use_ade, nms_on, thr_veh, thr_ped, thr_cyc, score_temperature,
      t_first, t_stride, n_out, out_trajs, out_scores, out_mode);
  TB_CHECK_LAUNCH();
  return TB_OK;
}
