// Backward of the KNARPE attention core (tb_knarpe_attn_bwd) - SURVEY.md 8(f) rank 2, first piece: the gradient of
// the one operator of the path that has no library equivalent. Reference: autograd through
// modules/attention_rpe.py:137-190 (RPE branch); in the re-associated form of DESIGN.md 3 the core op is
//   l_hj = q_h.k_hj + u_h.e_j   (q, u pre-scaled by log2(e)/sqrt(d_head)),  p_h = softmax2_j(l_hj) over unmasked j,
//   ov_h = sum_j p_hj v_hj,     z_h = sum_j p_hj e_j
// so with g_hj = d_ov_h.v_hj + d_z_h.e_j and dl_hj = ln2 * p_hj (g_hj - sum_i p_hi g_hi):
//   d_q_h = sum_j dl_hj k_hj,  d_u_h = sum_j dl_hj e_j,  d_k_hj += dl_hj q_h,  d_v_hj += p_hj d_ov_h   (scatter-add).
// The projections around it (q|u, k|v, out) are plain GEMMs whose gradients are GEMMs. e_j depends on poses only: no
// gradient is propagated into the relative pose (the reference detaches the states that feed it during the rollout
// loss, waymo_motion.py:313-385 with teacher forcing).
// fp32 only (parity path), d_model = d_rpe = 128, 4 heads. One warp per token, two passes over its neighbour list
// (A: logits and g together, reduced with a halving butterfly; softmax statistics from shared memory; B: gradients),
// the embedding is re-evaluated in registers in each pass, logits and g live in 4 KB of shared memory per warp. K / V gradients are scattered with vector atomics
// (red.global.add.v4.f32): rows shared by many tokens serialise in L2 - this is the straightforward version.
#include "common.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int D = 128;
constexpr int H = 4;
constexpr int KMAX = 128;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(TB_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWarps * 32)
knarpe_attn_bwd_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ u, int ldu,
                       const float* __restrict__ kv0, int ldkv0, int T0, int div0, int K0,
                       const float* __restrict__ kv1, int ldkv1, int T1, int div1, int K1,
                       const int32_t* __restrict__ idx, const uint8_t* __restrict__ invalid,
                       const float* __restrict__ rel, const float* __restrict__ pe_freq_xy, int n_tok, int S,
                       const float* __restrict__ d_ov, const float* __restrict__ d_z, int ldo,
                       float* __restrict__ d_q, float* __restrict__ d_u, int ldg,
                       float* __restrict__ d_kv0, float* __restrict__ d_kv1) {
  __shared__ float s_l[kWarps][H][KMAX];  // logits, then probabilities
  __shared__ float s_g[kWarps][H][KMAX];  // g_hj
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * kWarps + warp;
  if (tok >= n_tok) return;
  const int b = tok / S;
  const int Ktot = K0 + K1;
  const int hh = lane >> 3;  // head of this lane's q/k/v channels [4 lane, 4 lane + 4)
  const size_t prow = (size_t)tok * Ktot;
  const size_t off0 = (size_t)(b / div0) * T0 * ldkv0;
  const size_t off1 = (K1 > 0) ? (size_t)(b / div1) * T1 * ldkv1 : 0;

  // embedding components of this lane: c = lane + 32 k, k = 0..3 -> x (cos: lane < 16, else sin), y, cos yaw, sin yaw
  const float fxy = __ldg(pe_freq_xy + (lane & 15));
  const float ph = (lane < 16) ? 1.57079632679489662f : 0.f;
  const float mh = (float)(lane + 1);
  auto emb = [&](int j, float (&e)[4]) {
    const float x = rel[(prow + j) * 3 + 0], y = rel[(prow + j) * 3 + 1], w = rel[(prow + j) * 3 + 2];
    e[0] = __sinf(tb_reduce_2pi(x * fxy) + ph);
    e[1] = __sinf(tb_reduce_2pi(y * fxy) + ph);
    const float aw = tb_reduce_2pi(w * mh);
    e[2] = __cosf(aw);
    e[3] = __sinf(aw);
  };
  auto row_off = [&](int j) -> size_t {
    const int id = idx[prow + j];
    return (j < K0) ? off0 + (size_t)id * ldkv0 : off1 + (size_t)id * ldkv1;
  };

  const float4 q4 = ldg4(q + (size_t)tok * ldq + lane * 4);
  const float4 go4 = ldg4(d_ov + (size_t)tok * ldo + lane * 4);
  float uu[H][4], gz[H][4];
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uu[h][k] = __ldg(u + (size_t)tok * ldu + h * D + lane + 32 * k);
      gz[h][k] = __ldg(d_z + (size_t)tok * ldo + h * D + lane + 32 * k);
    }

  // ---- pass A: logits l_hj AND g_hj = d_ov_h.v_hj + d_z_h.e_j of the unmasked neighbours in one sweep (one embedding
  // evaluation, K and V rows gathered together). The 8 per-lane partials (4 heads x {l, g}) are reduced over the warp
  // with a halving butterfly: 4 + 2 + 1 exchanges fold 8 -> 1 value per lane, 2 more finish the sum = 9 shuffles per
  // neighbour instead of 8 x 5. Lane 4 i (i = 0..7) ends up with value i and writes it to shared memory.
  const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
  for (int j = 0; j < Ktot; ++j) {
    if (invalid[prow + j]) continue;  // warp-uniform
    float e[4];
    emb(j, e);
    const float* kp = ((j < K0) ? kv0 : kv1) + row_off(j);
    const float4 k4 = ldg4(kp + lane * 4);
    const float4 v4 = ldg4(kp + D + lane * 4);
    const float qk = q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
    const float gv = go4.x * v4.x + go4.y * v4.y + go4.z * v4.z + go4.w * v4.w;
    float val[8];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      val[h] = uu[h][0] * e[0] + uu[h][1] * e[1] + uu[h][2] * e[2] + uu[h][3] * e[3] + (h == hh ? qk : 0.f);
      val[4 + h] = gz[h][0] * e[0] + gz[h][1] * e[1] + gz[h][2] * e[2] + gz[h][3] * e[3] + (h == hh ? gv : 0.f);
    }
    float r4[4], r2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      r4[i] = (b16 ? val[i + 4] : val[i]) + __shfl_xor_sync(TB_FULL_MASK, b16 ? val[i] : val[i + 4], 16);
#pragma unroll
    for (int i = 0; i < 2; ++i)
      r2[i] = (b8 ? r4[i + 2] : r4[i]) + __shfl_xor_sync(TB_FULL_MASK, b8 ? r4[i] : r4[i + 2], 8);
    float r1 = (b4 ? r2[1] : r2[0]) + __shfl_xor_sync(TB_FULL_MASK, b4 ? r2[0] : r2[1], 4);
    r1 += __shfl_xor_sync(TB_FULL_MASK, r1, 2);
    r1 += __shfl_xor_sync(TB_FULL_MASK, r1, 1);
    if ((lane & 3) == 0) {
      const int vi = lane >> 2;  // 4 b16 + 2 b8 + b4: values 0-3 are the logits of heads 0-3, 4-7 their g
      if (vi < 4) s_l[warp][vi][j] = r1; else s_g[warp][vi - 4][j] = r1;
    }
  }
  __syncwarp();
  // ---- softmax statistics and dot_h = sum_j p_hj g_hj from shared memory (lanes over neighbours)
  float mx[H] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  for (int j = lane; j < Ktot; j += 32) {
    if (invalid[prow + j]) continue;
#pragma unroll
    for (int h = 0; h < H; ++h) mx[h] = fmaxf(mx[h], s_l[warp][h][j]);
  }
#pragma unroll
  for (int h = 0; h < H; ++h) {
#pragma unroll
    for (int o = 16; o; o >>= 1) mx[h] = fmaxf(mx[h], __shfl_xor_sync(TB_FULL_MASK, mx[h], o));
  }
  float sm[H] = {0.f, 0.f, 0.f, 0.f}, dot[H] = {0.f, 0.f, 0.f, 0.f};
  for (int j = lane; j < Ktot; j += 32) {
    const bool ok = !invalid[prow + j];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float p = ok ? ex2(s_l[warp][h][j] - mx[h]) : 0.f;
      s_l[warp][h][j] = p;
      sm[h] += p;
      dot[h] += ok ? p * s_g[warp][h][j] : 0.f;
    }
  }
  float inv[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    sm[h] = warp_sum(sm[h]);
    dot[h] = warp_sum(dot[h]);
    inv[h] = sm[h] > 0.f ? 1.f / sm[h] : 0.f;
    dot[h] *= inv[h];
  }
  __syncwarp();

  // ---- pass B: dl_hj = ln2 p_hj (g_hj - dot_h); accumulate d_q, d_u; scatter d_k, d_v
  float dq[4] = {0.f, 0.f, 0.f, 0.f};
  float du[H][4];
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int k = 0; k < 4; ++k) du[h][k] = 0.f;
  for (int j = 0; j < Ktot; ++j) {
    if (invalid[prow + j]) continue;
    float e[4];
    emb(j, e);
    const size_t ro = row_off(j);
    const float* kp = ((j < K0) ? kv0 : kv1) + ro;
    float* gp = ((j < K0) ? d_kv0 : d_kv1) + ro;
    const float4 k4 = ldg4(kp + lane * 4);
    float dl[H], p_own = 0.f;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float p = s_l[warp][h][j] * inv[h];
      dl[h] = 0.69314718055994531f * p * (s_g[warp][h][j] - dot[h]);
      if (h == hh) p_own = p;
#pragma unroll
      for (int k = 0; k < 4; ++k) du[h][k] = fmaf(dl[h], e[k], du[h][k]);
    }
    const float dlo = hh == 0 ? dl[0] : (hh == 1 ? dl[1] : (hh == 2 ? dl[2] : dl[3]));
    dq[0] = fmaf(dlo, k4.x, dq[0]); dq[1] = fmaf(dlo, k4.y, dq[1]);
    dq[2] = fmaf(dlo, k4.z, dq[2]); dq[3] = fmaf(dlo, k4.w, dq[3]);
    red_add4(gp + lane * 4, dlo * q4.x, dlo * q4.y, dlo * q4.z, dlo * q4.w);                    // d_k
    red_add4(gp + D + lane * 4, p_own * go4.x, p_own * go4.y, p_own * go4.z, p_own * go4.w);    // d_v
  }
  *reinterpret_cast<float4*>(d_q + (size_t)tok * ldg + lane * 4) = make_float4(dq[0], dq[1], dq[2], dq[3]);
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int k = 0; k < 4; ++k) d_u[(size_t)tok * ldg + D + h * D + lane + 32 * k] = du[h][k];
}

}  // namespace

extern "C" int tb_knarpe_attn_bwd(const float* q, int ldq, const float* u, int ldu, const float* kv0, int ldkv0, int T0,
                                  int div0, int K0, const float* kv1, int ldkv1, int T1, int div1, int K1,
                                  const int32_t* idx, const uint8_t* invalid, const float* rel, const float* pe_freq_xy,
                                  int B, int S, int D_, int Hh, const float* d_ov, const float* d_z, int ldo,
                                  float* d_qu, int ldg, float* d_kv0, float* d_kv1, void* stream) {
  if (!q || !u || !kv0 || !idx || !invalid || !rel || !pe_freq_xy || !d_ov || !d_z || !d_qu || !d_kv0) return TB_ERR_NULL;
  if (B <= 0 || S <= 0 || K0 <= 0 || K1 < 0 || T0 <= 0 || div0 <= 0 || (K1 > 0 && (!kv1 || !d_kv1 || T1 <= 0 || div1 <= 0)))
    return TB_ERR_BAD_SHAPE;
  if (D_ != D || Hh != H || K0 + K1 > KMAX) return TB_ERR_UNSUPPORTED;
  if (ldq < D || ldu < H * D || ldkv0 < 2 * D || (K1 > 0 && ldkv1 < 2 * D) || ldo < D || ldg < D + H * D)
    return TB_ERR_BAD_SHAPE;
  if ((ldq | ldu | ldkv0 | ldo | ldg | (K1 > 0 ? ldkv1 : 0)) & 3) return TB_ERR_MISALIGNED;
  if (!tb_aligned16(q) || !tb_aligned16(kv0) || (K1 > 0 && !tb_aligned16(kv1)) || !tb_aligned16(d_ov) ||
      !tb_aligned16(d_qu) || !tb_aligned16(d_kv0) || (K1 > 0 && !tb_aligned16(d_kv1)))
    return TB_ERR_MISALIGNED;
  const int n_tok = B * S;
  knarpe_attn_bwd_kernel<<<(n_tok + kWarps - 1) / kWarps, kWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      q, ldq, u, ldu, kv0, ldkv0, T0, div0, K0, kv1, ldkv1, T1, div1, K1, idx, invalid, rel, pe_freq_xy, n_tok, S, d_ov, d_z,
      ldo, d_qu, d_qu, ldg, d_kv0, d_kv1);
  TB_CHECK_LAUNCH();
  return TB_OK;
}
