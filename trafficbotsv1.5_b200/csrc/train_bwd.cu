// Backward kernels of the training path (BASELINE config 4): the gradients of the hot-path
// layers that the forward kernels of this library compute, and the differentiable part of the closed loop (unicycle
// dynamics + imitation loss, traffic-light state NLL) as forward / reverse scans over the recorded rollout.
//   reference: pl_modules/waymo_motion.py:313-385 (training_step), utils/dynamics.py:66-141,237-274,
//   utils/rewards.py:35-85, models/metrics/loss.py:9-37, models/metrics/training.py:76-160.
// The policy's inputs are detached in training (waymo_motion.py:158-161, `training_detach_model_input`), so the only
// path through time is the state recurrence pose/speed(t+1) = f(pose/speed(t), action(t)): one thread per agent walks
// it backwards and emits dL/d(action head output) for every step; the network backward of every step then runs on
// tb_linear (dX), tb_linear_wgrad (dW, db), tb_layernorm_bwd, tb_knarpe_attn_bwd and tb_pointnet_pool_bwd.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[n,k] += sum_m dY[m,n] X[m,k],  db[n] += sum_m dY[m,n].  64x64 output tile per CTA, the M range split over
// gridDim.z (split-K with fp32 atomics into the caller-zeroed accumulators: the same buffer collects every call that
// shares the weight). 256 threads, 4x4 outputs each, 32-row stages in shared memory.
constexpr int WG_T = 64, WG_R = 32;

template <bool VEC>
__global__ void __launch_bounds__(256)
linear_wgrad_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ X, int ldx, int M, int N, int K,
                    float* __restrict__ dW, int lddw, float* __restrict__ db, int rows_per_split) {
  __shared__ __align__(16) float sA[WG_R][WG_T];  // dY tile: [m][n]
  __shared__ __align__(16) float sB[WG_R][WG_T];  // X tile:  [m][k]
  const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T;
  const int m_begin = blockIdx.z * rows_per_split, m_end = min(M, m_begin + rows_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;  // loader: rows lr, lr + 16; columns lc .. lc + 3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bs[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_bias = db != nullptr && blockIdx.y == 0 && tx == 0;
  for (int m0 = m_begin; m0 < m_end; m0 += WG_R) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + 16 * h, m = m0 + r;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (m < m_end) {
        const float* ap = dY + (size_t)m * lddy + n0 + lc;
        const float* bp = X + (size_t)m * ldx + k0 + lc;
        if (VEC && n0 + lc + 3 < N) a = *reinterpret_cast<const float4*>(ap);
        else {
          if (n0 + lc < N) a.x = ap[0];
          if (n0 + lc + 1 < N) a.y = ap[1];
          if (n0 + lc + 2 < N) a.z = ap[2];
          if (n0 + lc + 3 < N) a.w = ap[3];
        }
        if (VEC && k0 + lc + 3 < K) b = *reinterpret_cast<const float4*>(bp);
        else {
          if (k0 + lc < K) b.x = bp[0];
          if (k0 + lc + 1 < K) b.y = bp[1];
          if (k0 + lc + 2 < K) b.z = bp[2];
          if (k0 + lc + 3 < K) b.w = bp[3];
        }
      }
      *reinterpret_cast<float4*>(&sA[r][lc]) = a;
      *reinterpret_cast<float4*>(&sB[r][lc]) = b;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < WG_R; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[r][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sB[r][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (do_bias) {
#pragma unroll
        for (int i = 0; i < 4; ++i) bs[i] += av[i];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(dW + (size_t)n * lddw + k, acc[i][j]);
    }
    if (do_bias) atomicAdd(db + n, bs[i]);
  }
}

// ------------------------------------------------------------------------------------------------ epilogue backward
// out[m,n] = keep ? dY[m,n] : 0 with keep = !(mask_a[m] | mask_b[m]) && (Y == nullptr || Y[m,n] > 0): the gradient
// w.r.t. the pre-activation of a tb_linear epilogue (ReLU and row masks; Y is the epilogue's own output).
// blockDim (32, 8): one row per threadIdx.y, 16-byte pieces along the row (VEC) or single floats.
template <bool VEC>
__global__ void __launch_bounds__(256)
grad_mask_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ Y, int ldy,
                 const uint8_t* __restrict__ mask_a, const uint8_t* __restrict__ mask_b, int M, int N,
                 float* __restrict__ out, int ldo) {
  const int m = blockIdx.x * 8 + threadIdx.y;
  if (m >= M) return;
  const bool row_keep = !((mask_a && mask_a[m]) || (mask_b && mask_b[m]));
  const float* dr = dY + (size_t)m * lddy;
  const float* yr = Y ? Y + (size_t)m * ldy : nullptr;
  float* orow = out + (size_t)m * ldo;
  if (VEC) {
    for (int c = threadIdx.x * 4; c < N; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_keep) {
        v = *reinterpret_cast<const float4*>(dr + c);
        if (yr) {
          const float4 y = *reinterpret_cast<const float4*>(yr + c);
          if (!(y.x > 0.f)) v.x = 0.f;
          if (!(y.y > 0.f)) v.y = 0.f;
          if (!(y.z > 0.f)) v.z = 0.f;
          if (!(y.w > 0.f)) v.w = 0.f;
        }
      }
      *reinterpret_cast<float4*>(orow + c) = v;
    }
  } else {
    for (int c = threadIdx.x; c < N; c += 32) {
      bool keep = row_keep;
      if (keep && yr) keep = yr[c] > 0.f;
      orow[c] = keep ? dr[c] : 0.f;
    }
  }
}

// out[g,n] = sum over the L rows of group g of X[g*L + l, n] (gradient of a grouped bias row).
__global__ void group_sum_kernel(const float* __restrict__ X, int ldx, int G, int L, int N, float* __restrict__ out,
                                 int ldo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)G * N) return;
  const int g = (int)(i / N), n = (int)(i - (size_t)g * N);
  float s = 0.f;
  for (int l = 0; l < L; ++l) s += X[((size_t)g * L + l) * ldx + n];
  out[(size_t)g * ldo + n] = s;
}

// ------------------------------------------------------------------------------------------------ PointNet pooling
// Backward of tb_pointnet_pool modes 1 / 2 (max over the valid rows of a group): the gradient of a group's maximum
// goes to the first valid row that attains it (ties only occur at ReLU zeros, where the ReLU backward discards it).
// mode 2 pooled [m | m]: both halves of dOut are summed. Every element of dX is written.
__global__ void pointnet_pool_bwd_kernel(const float* __restrict__ X, int ldx, const uint8_t* __restrict__ invalid,
                                         int G, int L, int C, int mode, const float* __restrict__ dOut, int ldo,
                                         float* __restrict__ dX, int lddx) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)G * C) return;
  const int g = (int)(i / C), c = (int)(i - (size_t)g * C);
  int best = -1;
  float bv = -INFINITY;
  for (int l = 0; l < L; ++l) {
    const size_t row = (size_t)g * L + l;
    if (invalid[row]) continue;
    const float v = X[row * ldx + c];
    if (best < 0 || v > bv) { bv = v; best = l; }
  }
  float d = dOut[(size_t)g * ldo + c];
  if (mode == 2) d += dOut[(size_t)g * ldo + C + c];
  for (int l = 0; l < L; ++l) dX[((size_t)g * L + l) * lddx + c] = (l == best) ? d : 0.f;
}

// ------------------------------------------------------------------------------------------------ closed loop
struct IlArgs {
  const float* act;            // [T, M, 6] action-head outputs of every step (pre-tanh, (veh, ped, cyc) x (acc, yaw rate))
  const uint8_t* ag_type;      // [M, 3]
  float max_acc[3], max_yaw[3], dt;
  const uint8_t* pred_valid;   // [M, T] validity at the start of step s (RolloutBuffer.pred_valid, index s - 1)
  const float* pose0;          // [M, 3] state at time 0
  const float* motion0;        // [M, 3]
  const uint8_t* gt_valid;     // [n_sc, A, n_gt]
  const float* gt_pose;        // [n_sc, A, n_gt, 3]
  const float* gt_motion;      // [n_sc, A, n_gt, 3]
  const uint8_t* tf_mask;      // [n_sc, A, n_gt] teacher-forcing / spawn override of step s
  const uint8_t* loss_mask;    // optional [M]: agents that take part in the loss (relevant-agent mask); nullptr = all
  int n_gt, sc_div, A, M, T, step_start;
  float w_pos, w_rot, w_spd;
  float4* state_in;            // [T, M] (x, y, yaw, speed) before step s
};

__device__ __forceinline__ float sl1(float d) { const float a = fabsf(d); return a < 1.f ? 0.5f * d * d : a - 0.5f; }
__device__ __forceinline__ float dsl1(float d) { return d > 1.f ? 1.f : (d < -1.f ? -1.f : d); }

__device__ __forceinline__ int agent_type(const uint8_t* ag_type, int i) {
  int type = -1;
#pragma unroll
  for (int t = 0; t < 3; ++t) if (ag_type[(size_t)i * 3 + t]) type = t;
  return type;
}

// Forward replay of the state recurrence (same operation order as dyn_step_kernel) + imitation loss
// (rewards.py:62-76: SmoothL1 position / speed, 0.5 (1 - cos) heading, weighted; training.py:93-130 masks).
// out[0] += sum of the weighted errors, out[1] += number of counted (agent, step) entries.
__global__ void __launch_bounds__(128) il_loss_fwd_kernel(IlArgs p, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float sum = 0.f, cnt = 0.f;
  if (i < p.M) {
    const int b = i / p.A, a = i - b * p.A, sc = b / p.sc_div;
    const int type = agent_type(p.ag_type, i);
    const bool counted = !p.loss_mask || p.loss_mask[i];
    float x = p.pose0[(size_t)i * 3], y = p.pose0[(size_t)i * 3 + 1], w = p.pose0[(size_t)i * 3 + 2];
    float spd = p.motion0[(size_t)i * 3];
    const float hdt = 0.5f * p.dt;
    for (int s = 1; s <= p.T; ++s) {
      const bool v_old = p.pred_valid[(size_t)i * p.T + (s - 1)] != 0;
      p.state_in[(size_t)(s - 1) * p.M + i] = make_float4(x, y, w, spd);
      float nx = 0.f, ny = 0.f, nw = 0.f, nspd = 0.f;
      if (v_old && type >= 0) {
        const float* ar = p.act + ((size_t)(s - 1) * p.M + i) * 6 + type * 2;
        const float acc = tanhf(ar[0]) * p.max_acc[type], yr = tanhf(ar[1]) * p.max_yaw[type];
        const float v_t = __fadd_rn(spd, __fmul_rn(hdt, acc)), th_t = __fadd_rn(w, __fmul_rn(hdt, yr));
        float sn, cs;
        sincosf(th_t, &sn, &cs);
        nx = __fadd_rn(x, __fmul_rn(p.dt, __fmul_rn(v_t, cs)));
        ny = __fadd_rn(y, __fmul_rn(p.dt, __fmul_rn(v_t, sn)));
        nw = __fadd_rn(w, __fmul_rn(p.dt, yr));
        nspd = __fadd_rn(spd, __fmul_rn(p.dt, acc));
      }
      const bool in_loss = counted && v_old && (s - 1) >= p.step_start;
      bool forced = false;
      if (s < p.n_gt) {
        const size_t g = ((size_t)sc * p.A + a) * p.n_gt + s;
        if (in_loss && p.gt_valid[g]) {
          const float ex = sl1(p.gt_pose[g * 3] - nx) + sl1(p.gt_pose[g * 3 + 1] - ny);
          const float er = 0.5f * (1.f - cosf(p.gt_pose[g * 3 + 2] - nw));
          const float es = sl1(p.gt_motion[g * 3] - nspd);
          sum += p.w_pos * ex + p.w_rot * er + p.w_spd * es;
          cnt += 1.f;
        }
        forced = p.tf_mask[g] != 0;
        if (forced) { nx = p.gt_pose[g * 3]; ny = p.gt_pose[g * 3 + 1]; nw = p.gt_pose[g * 3 + 2]; nspd = p.gt_motion[g * 3]; }
      } else if (in_loss) {
        cnt += 1.f;  // no ground truth: zero reward, but the entry is counted (rewards.py:46-53, training.py:124-138)
      }
      x = nx; y = ny; w = nw; spd = nspd;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    sum += __shfl_xor_sync(TB_FULL_MASK, sum, o);
    cnt += __shfl_xor_sync(TB_FULL_MASK, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out, sum); atomicAdd(out + 1, cnt); }
}

// Reverse scan: adjoint of (x, y, yaw, speed) carried backwards through the recurrence; d_act [T, M, 6] fully written.
// g_out: upstream gradient of out[0] (device scalar).
__global__ void __launch_bounds__(128) il_loss_bwd_kernel(IlArgs p, const float* __restrict__ g_out,
                                                          float* __restrict__ d_act) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.M) return;
  const int b = i / p.A, a = i - b * p.A, sc = b / p.sc_div;
  const int type = agent_type(p.ag_type, i);
  const bool counted = !p.loss_mask || p.loss_mask[i];
  const float go = *g_out, hdt = 0.5f * p.dt;
  float ax = 0.f, ay = 0.f, aw = 0.f, as = 0.f;  // adjoint of the state after step s
  for (int s = p.T; s >= 1; --s) {
    float* dr = d_act + ((size_t)(s - 1) * p.M + i) * 6;
    const bool v_old = p.pred_valid[(size_t)i * p.T + (s - 1)] != 0;
    float gx = ax, gy = ay, gw = aw, gs = as;  // adjoint of the prediction of step s
    size_t g = 0;
    if (s < p.n_gt) {
      g = ((size_t)sc * p.A + a) * p.n_gt + s;
      if (p.tf_mask[g]) gx = gy = gw = gs = 0.f;  // state replaced by the ground truth: the chain is cut
    }
    float d0 = 0.f, d1 = 0.f;
    ax = ay = aw = as = 0.f;
    if (v_old && type >= 0) {
      const float4 st = p.state_in[(size_t)(s - 1) * p.M + i];
      const float* ar = p.act + ((size_t)(s - 1) * p.M + i) * 6 + type * 2;
      const float t0 = tanhf(ar[0]), t1 = tanhf(ar[1]);
      const float acc = t0 * p.max_acc[type], yr = t1 * p.max_yaw[type];
      const float v_t = __fadd_rn(st.w, __fmul_rn(hdt, acc)), th_t = __fadd_rn(st.z, __fmul_rn(hdt, yr));
      float sn, cs;
      sincosf(th_t, &sn, &cs);
      if (s < p.n_gt && counted && (s - 1) >= p.step_start && p.gt_valid[g]) {
        const float nx = __fadd_rn(st.x, __fmul_rn(p.dt, __fmul_rn(v_t, cs)));
        const float ny = __fadd_rn(st.y, __fmul_rn(p.dt, __fmul_rn(v_t, sn)));
        const float nw = __fadd_rn(st.z, __fmul_rn(p.dt, yr));
        const float nspd = __fadd_rn(st.w, __fmul_rn(p.dt, acc));
        gx -= go * p.w_pos * dsl1(p.gt_pose[g * 3] - nx);
        gy -= go * p.w_pos * dsl1(p.gt_pose[g * 3 + 1] - ny);
        gw -= go * p.w_rot * 0.5f * sinf(p.gt_pose[g * 3 + 2] - nw);
        gs -= go * p.w_spd * dsl1(p.gt_motion[g * 3] - nspd);
      }
      const float d_vt = p.dt * (gx * cs + gy * sn);
      const float d_th = p.dt * v_t * (gy * cs - gx * sn);
      ax = gx; ay = gy; aw = gw + d_th; as = gs + d_vt;
      const float d_acc = hdt * d_vt + p.dt * gs, d_yr = hdt * d_th + p.dt * gw;
      d0 = d_acc * p.max_acc[type] * (1.f - t0 * t0);
      d1 = d_yr * p.max_yaw[type] * (1.f - t1 * t1);
    }
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      dr[t * 2] = t == type ? d0 : 0.f;
      dr[t * 2 + 1] = t == type ? d1 : 0.f;
    }
  }
}

// Traffic-light state NLL over all steps (waymo_motion.py:270-277, traffic_light.py:284-286, training.py:155-160):
// -log_softmax(clamp(logits, +-3))[argmax gt state] for valid lights and steps with ground truth.
// logits [T, n, 5] pre-clamp (rows of invalid lights are zero-filled by the predictor's mask). out[0] += sum, out[1] += count.
__global__ void tl_nll_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ tl_invalid,
                              const uint8_t* __restrict__ gt_tl, int n_gt, int n, int T, float* __restrict__ out,
                              const float* __restrict__ g_out, float* __restrict__ d_logits) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // (s - 1) * n + i
  float nll = 0.f, cnt = 0.f;
  if (j < (size_t)T * n) {
    const int s = (int)(j / n) + 1, i = (int)(j - (size_t)(s - 1) * n);
    float d[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (s < n_gt && !tl_invalid[i]) {
      float v[5], mx = -INFINITY;
      bool pass[5];
      int gt = 0;
      uint8_t best = 0;
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        const float l = logits[j * 5 + c];
        pass[c] = l >= -3.f && l <= 3.f;
        v[c] = fminf(fmaxf(l, -3.f), 3.f);
        mx = fmaxf(mx, v[c]);
        const uint8_t gv = gt_tl[((size_t)i * n_gt + s) * 5 + c];
        if (gv > best) { best = gv; gt = c; }
      }
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < 5; ++c) se += expf(v[c] - mx);
      nll = -(v[gt] - mx - logf(se));
      cnt = 1.f;
      if (d_logits) {
        const float go = *g_out;
#pragma unroll
        for (int c = 0; c < 5; ++c) d[c] = pass[c] ? go * (expf(v[c] - mx) / se - (c == gt ? 1.f : 0.f)) : 0.f;
      }
    }
    if (d_logits) {
#pragma unroll
      for (int c = 0; c < 5; ++c) d_logits[j * 5 + c] = d[c];
    }
  }
  if (out) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      nll += __shfl_xor_sync(TB_FULL_MASK, nll, o);
      cnt += __shfl_xor_sync(TB_FULL_MASK, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0.f) { atomicAdd(out, nll); atomicAdd(out + 1, cnt); }
  }
}

// Categorical NLL over a row of logits with -inf entries (destination classifier, navigation.py:265-278 +
// training.py:146-153): nll[r] = logsumexp(logits[r]) - logits[r, target[r]] for rows with row_valid, else 0.
// One warp per row. out[0] += sum of nll, out[1] += number of valid rows; with d_logits: g_out[0] * (softmax - onehot)
// on valid rows, 0 elsewhere (every element written).
__global__ void __launch_bounds__(256)
softmax_nll_kernel(const float* __restrict__ logits, int ld, const int64_t* __restrict__ target,
                   const uint8_t* __restrict__ row_valid, int R, int Cn, float* __restrict__ out,
                   const float* __restrict__ g_out, float* __restrict__ d_logits, int ldd) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* row = logits + (size_t)r * ld;
  const bool ok = row_valid[r] != 0;
  float mx = -INFINITY;
  for (int c = lane; c < Cn; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(TB_FULL_MASK, mx, o));
  float se = 0.f;
  for (int c = lane; c < Cn; c += 32) se += expf(row[c] - mx);
#pragma unroll
  for (int o = 16; o; o >>= 1) se += __shfl_xor_sync(TB_FULL_MASK, se, o);
  const int t = (int)target[r];
  if (out && ok && lane == 0) {
    atomicAdd(out, mx + logf(se) - row[t]);
    atomicAdd(out + 1, 1.f);
  }
  if (d_logits) {
    const float go = ok ? *g_out : 0.f, inv = 1.f / se;
    float* dr = d_logits + (size_t)r * ldd;
    for (int c = lane; c < Cn; c += 32) dr[c] = ok ? go * (expf(row[c] - mx) * inv - (c == t ? 1.f : 0.f)) : 0.f;
  }
}

IlArgs make_il_args(const float* act, const uint8_t* ag_type, const float* max_acc, const float* max_yaw, float dt,
                    const uint8_t* pred_valid, const float* pose0, const float* motion0, const uint8_t* gt_valid,
                    const float* gt_pose, const float* gt_motion, const uint8_t* tf_mask, const uint8_t* loss_mask,
                    int n_gt, int sc_div, int B, int A, int T, int step_start, float w_pos, float w_rot, float w_spd,
                    float* state_in) {
  IlArgs p;
  p.act = act; p.ag_type = ag_type; p.dt = dt;
  for (int t = 0; t < 3; ++t) { p.max_acc[t] = max_acc[t]; p.max_yaw[t] = max_yaw[t]; }
  p.pred_valid = pred_valid; p.pose0 = pose0; p.motion0 = motion0; p.gt_valid = gt_valid; p.gt_pose = gt_pose;
  p.gt_motion = gt_motion; p.tf_mask = tf_mask; p.loss_mask = loss_mask; p.n_gt = n_gt; p.sc_div = sc_div; p.A = A;
  p.M = B * A; p.T = T; p.step_start = step_start; p.w_pos = w_pos; p.w_rot = w_rot; p.w_spd = w_spd;
  p.state_in = reinterpret_cast<float4*>(state_in);
  return p;
}

}  // namespace

// fp32 FFMA weight gradient (tb_linear_wgrad precision 0 and the shapes the tcgen05 kernel of wgrad_tc.cu cannot take)
int tb_linear_wgrad_f32(const float* dY, int lddy, const float* X, int ldx, int M, int N, int K, float* dW, int lddw,
                        float* db, cudaStream_t st) {
  const int tn = (N + WG_T - 1) / WG_T, tk = (K + WG_T - 1) / WG_T;
  int splits = (148 * 4 + tn * tk - 1) / (tn * tk);
  const int max_splits = (M + 2 * WG_R - 1) / (2 * WG_R);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int rows = (M + splits - 1) / splits;
  rows = (rows + WG_R - 1) / WG_R * WG_R;
  splits = (M + rows - 1) / rows;
  const dim3 grid(tn, tk, splits);
  const bool vec = ((lddy | ldx) & 3) == 0 && tb_aligned16(dY) && tb_aligned16(X);
  if (vec) linear_wgrad_kernel<true><<<grid, 256, 0, st>>>(dY, lddy, X, ldx, M, N, K, dW, lddw, db, rows);
  else linear_wgrad_kernel<false><<<grid, 256, 0, st>>>(dY, lddy, X, ldx, M, N, K, dW, lddw, db, rows);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_grad_mask(const float* dY, int lddy, const float* Y, int ldy, const uint8_t* mask_a,
                            const uint8_t* mask_b, int M, int N, float* out, int ldo, void* stream) {
  if (!dY || !out) return TB_ERR_NULL;
  if (M <= 0 || N <= 0 || lddy < N || ldo < N || (Y && ldy < N)) return TB_ERR_BAD_SHAPE;
  const bool vec = (N & 3) == 0 && ((lddy | ldo | (Y ? ldy : 0)) & 3) == 0 && tb_aligned16(dY) && tb_aligned16(out) &&
                   (!Y || tb_aligned16(Y));
  const dim3 block(32, 8), grid((M + 7) / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec) grad_mask_kernel<true><<<grid, block, 0, st>>>(dY, lddy, Y, ldy, mask_a, mask_b, M, N, out, ldo);
  else grad_mask_kernel<false><<<grid, block, 0, st>>>(dY, lddy, Y, ldy, mask_a, mask_b, M, N, out, ldo);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_group_sum(const float* X, int ldx, int G, int L, int N, float* out, int ldo, void* stream) {
  if (!X || !out) return TB_ERR_NULL;
  if (G <= 0 || L <= 0 || N <= 0 || ldx < N || ldo < N) return TB_ERR_BAD_SHAPE;
  const size_t n = (size_t)G * N;
  group_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(X, ldx, G, L, N, out, ldo);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_pointnet_pool_bwd(const float* X, int ldx, const uint8_t* invalid, int G, int L, int C, int mode,
                                    const float* dOut, int ldo, float* dX, int lddx, void* stream) {
  if (!X || !invalid || !dOut || !dX) return TB_ERR_NULL;
  if (G <= 0 || L <= 0 || C <= 0 || ldx < C || lddx < C || (mode != 1 && mode != 2) || ldo < C * mode)
    return TB_ERR_BAD_SHAPE;
  const size_t n = (size_t)G * C;
  pointnet_pool_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      X, ldx, invalid, G, L, C, mode, dOut, ldo, dX, lddx);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_il_loss_fwd(const float* act, const uint8_t* ag_type, const float* max_acc, const float* max_yaw_rate,
                              float dt, const uint8_t* pred_valid, const float* pose0, const float* motion0,
                              const uint8_t* gt_valid, const float* gt_pose, const float* gt_motion,
                              const uint8_t* tf_mask, const uint8_t* loss_mask, int n_gt, int sc_div, int B, int A, int T,
                              int step_start, float w_pos, float w_rot, float w_spd, float* state_in, float* out,
                              void* stream) {
  if (!act || !ag_type || !max_acc || !max_yaw_rate || !pred_valid || !pose0 || !motion0 || !gt_valid || !gt_pose ||
      !gt_motion || !tf_mask || !state_in || !out)
    return TB_ERR_NULL;
  if (B <= 0 || A <= 0 || T <= 0 || n_gt <= 0 || sc_div <= 0) return TB_ERR_BAD_SHAPE;
  if (!tb_aligned16(state_in)) return TB_ERR_MISALIGNED;
  const IlArgs p = make_il_args(act, ag_type, max_acc, max_yaw_rate, dt, pred_valid, pose0, motion0, gt_valid, gt_pose,
                                gt_motion, tf_mask, loss_mask, n_gt, sc_div, B, A, T, step_start, w_pos, w_rot, w_spd,
                                state_in);
  il_loss_fwd_kernel<<<(p.M + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(p, out);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_il_loss_bwd(const float* act, const uint8_t* ag_type, const float* max_acc, const float* max_yaw_rate,
                              float dt, const uint8_t* pred_valid, const uint8_t* gt_valid, const float* gt_pose,
                              const float* gt_motion, const uint8_t* tf_mask, const uint8_t* loss_mask, int n_gt,
                              int sc_div, int B, int A, int T, int step_start, float w_pos, float w_rot, float w_spd,
                              const float* state_in, const float* g_out, float* d_act, void* stream) {
  if (!act || !ag_type || !max_acc || !max_yaw_rate || !pred_valid || !gt_valid || !gt_pose || !gt_motion || !tf_mask ||
      !state_in || !g_out || !d_act)
    return TB_ERR_NULL;
  if (B <= 0 || A <= 0 || T <= 0 || n_gt <= 0 || sc_div <= 0) return TB_ERR_BAD_SHAPE;
  if (!tb_aligned16(state_in)) return TB_ERR_MISALIGNED;
  const IlArgs p = make_il_args(act, ag_type, max_acc, max_yaw_rate, dt, pred_valid, nullptr, nullptr, gt_valid, gt_pose,
                                gt_motion, tf_mask, loss_mask, n_gt, sc_div, B, A, T, step_start, w_pos, w_rot, w_spd,
                                const_cast<float*>(state_in));
  il_loss_bwd_kernel<<<(p.M + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(p, g_out, d_act);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_tl_nll(const float* logits, const uint8_t* tl_invalid, const uint8_t* gt_tl, int n_gt, int n, int T,
                         float* out, const float* g_out, float* d_logits, void* stream) {
  if (!logits || !tl_invalid || !gt_tl || (!out && !d_logits) || (d_logits && !g_out)) return TB_ERR_NULL;
  if (n <= 0 || T <= 0 || n_gt <= 0) return TB_ERR_BAD_SHAPE;
  const size_t tot = (size_t)T * n;
  tl_nll_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, tl_invalid, gt_tl, n_gt, n, T, out, g_out, d_logits);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_softmax_nll(const float* logits, int ld, const int64_t* target, const uint8_t* row_valid, int R, int C,
                              float* out, const float* g_out, float* d_logits, int ldd, void* stream) {
  if (!logits || !target || !row_valid || (!out && !d_logits) || (d_logits && !g_out)) return TB_ERR_NULL;
  if (R <= 0 || C <= 0 || ld < C || (d_logits && ldd < C)) return TB_ERR_BAD_SHAPE;
  softmax_nll_kernel<<<(R + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, ld, target, row_valid, R, C, out,
                                                                               g_out, d_logits, ldd);
  TB_CHECK_LAUNCH();
  return TB_OK;
}
