// KNARPE attention core (tb_knarpe_attn): neighbour gather + relative-pose bias + masked softmax + weighted sums.
// Reference behaviour: modules/attention_rpe.py:137-190 (RPE branch) between the projections.
//
// Re-associated form (exact in real arithmetic):
//   q_h.(k_hj + W_rk,h e_j + b_rk,h) = q_h.k_hj + (W_rk,h^T q_h).e_j + const_h   (const_h cancels in the softmax)
//   sum_j a_hj (v_hj + W_rv,h e_j + b_rv,h) = sum_j a_hj v_hj + W_rv,h (sum_j a_hj e_j) + b_rv,h
// so the kernel needs no weights: it consumes q, u_h = W_rk,h^T q_h (both pre-scaled by log2(e)/sqrt(d_head) by the
// caller's projection) and emits ov = sum a v and z_h = sum a e; the dense parts stay in the projections.
//
// Mapping (v3; v1/v2 and their ncu instruction mixes are summarised in profiles/r1_notes.md): one warp per token.
//   * q/k/v: lane l owns feature columns [l*D/32, (l+1)*D/32) => head l>>3; a K or V row is one fully coalesced
//     512 B / 1 KiB request; the rows of the G neighbours of a group are in flight before first use.
//   * relative-pose embedding e_j (D sin/cos components, utils/pose_emb.py:50-55): lane l evaluates components
//     {l + 32k} in registers from the 12-byte relative pose (2-term Cody-Waite reduction + SFU) and multiplies them
//     with its slice of u for all 4 heads (packed FFMA2). Heads are held in a lane-permuted order (slot i = head
//     i ^ (l>>3)), which turns the 4-value cross-lane reduction into a select-free halving butterfly (3 shuffles)
//     that ends with each 8-lane group holding ITS head's partial, so it shares the 3-shuffle reduction of q.k.
//   * softmax: one evaluation per head and group of G neighbours (lanes of a head group compute their own head
//     only, probabilities are exchanged with 3 xor-shuffles), one accumulator rescale per group.
//   * everything inside a group is branch-free (masked neighbours get logit -inf), so the G dependency chains
//     interleave (the v2 kernel was latency-bound: 33% stall_wait, 26% short-scoreboard at 14 warps/SM).
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"

namespace {

#ifndef TB_ATTN_WARPS
#define TB_ATTN_WARPS 8
#endif
#ifndef TB_ATTN_MINB
#define TB_ATTN_MINB 2
#endif
#ifndef TB_ATTN_G
#define TB_ATTN_G 4
#endif
constexpr int kWarps = TB_ATTN_WARPS;
constexpr int H = 4;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// OUT_H: [ov|z] written as fp16; IN_H: q / u rows are fp16 (tensor-core mode intermediates); KV_H: the K|V tables are
// fp16 rows (ldkv in halves) - the 16-bit variant for shapes the mma.sync kernel does not cover (d_model 256,
// BASELINE config 2; lists longer than 128): half the gather bytes, fp32 arithmetic.
template <int D, bool FROM_EMB, bool FAST_TRIG, bool OUT_H = false, bool IN_H = false, bool KV_H = false>
__global__ void __launch_bounds__(kWarps * 32, TB_ATTN_MINB)
knarpe_attn_kernel(const void* __restrict__ q_, int ldq, const void* __restrict__ u_, int ldu,
                   const void* __restrict__ kv0_, int ldkv0, int T0, int div0, int K0,
                   const void* __restrict__ kv1_, int ldkv1, int T1, int div1, int K1,
                   const int32_t* __restrict__ idx, const uint8_t* __restrict__ invalid,
                   const float* __restrict__ rel, const float* __restrict__ emb,
                   const float* __restrict__ pe_freq_xy, int n_tok, int S,
                   void* __restrict__ out_ov_, void* __restrict__ out_z_, int ldo,
                   uint8_t* __restrict__ out_none_valid) {
  constexpr int NV = D / 32;            // q/k/v floats per lane
  constexpr int NC = D / 32;            // embedding components per lane (l + 32k)
  constexpr int G = (D == 128) ? TB_ATTN_G : 2; // neighbours per group
  constexpr int NF = D / 8;             // number of xy frequencies
  using KvT = typename std::conditional<KV_H, __half, float>::type;
  const KvT* kv0 = static_cast<const KvT*>(kv0_);
  const KvT* kv1 = static_cast<const KvT*>(kv1_);
  __shared__ const KvT* s_ptr[kWarps][32];    // compacted valid neighbours of the current chunk: K/V row pointer,
  __shared__ float s_rel[kWarps][32][3];      // relative pose,
  __shared__ int s_j[kWarps][32];             // and original neighbour slot (materialised-embedding mode)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * kWarps + warp;
  if (tok >= n_tok) return;  // warp-uniform; only __syncwarp below
  const int b = tok / S;
  const int Ktot = K0 + K1;
  const int hh = lane >> 3;  // own head; slot i of the per-head arrays holds head (i ^ hh)
  const unsigned lt_mask = (1u << lane) - 1u;

  float fxy, ph;
  if (D == 128) {  // one x and one y component per lane: frequency lane&15, cos for lane<16 else sin
    fxy = __ldg(pe_freq_xy + (lane & (NF - 1)));
    ph = (lane < NF) ? 1.57079632679489662f : 0.f;
  } else {
    fxy = __ldg(pe_freq_xy + lane);
    ph = 0.f;
  }
  const float m1 = (float)(lane + 1), m2 = (float)(lane + 33);

  // u slices: component k (= lane + 32k) of head slots (0,1) and (2,3) packed for head-pair FFMA2
  float qr[NV];
  float2 u01[NC], u23[NC], z[H][NC / 2];
  {
    if (IN_H) {
      const __half* qp = static_cast<const __half*>(q_) + (size_t)tok * ldq + lane * NV;
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        const float2 t = __half22float2(__ldg(reinterpret_cast<const __half2*>(qp + i)));
        qr[i] = t.x; qr[i + 1] = t.y;
      }
      const __half* up = static_cast<const __half*>(u_) + (size_t)tok * ldu + lane;
      auto ld = [&](int h, int k) { return __half2float(__ldg(up + h * D + 32 * k)); };
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        u01[k] = make_float2(ld(0 ^ hh, k), ld(1 ^ hh, k));
        u23[k] = make_float2(ld(2 ^ hh, k), ld(3 ^ hh, k));
      }
    } else {
      const float* qp = static_cast<const float*>(q_) + (size_t)tok * ldq + lane * NV;
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        float4 t = ldg4(qp + i);
        qr[i] = t.x; qr[i + 1] = t.y; qr[i + 2] = t.z; qr[i + 3] = t.w;
      }
      const float* up = static_cast<const float*>(u_) + (size_t)tok * ldu + lane;
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        u01[k] = make_float2(__ldg(up + (0 ^ hh) * D + 32 * k), __ldg(up + (1 ^ hh) * D + 32 * k));
        u23[k] = make_float2(__ldg(up + (2 ^ hh) * D + 32 * k), __ldg(up + (3 ^ hh) * D + 32 * k));
      }
    }
#pragma unroll
    for (int i = 0; i < H; ++i)
#pragma unroll
      for (int k = 0; k < NC / 2; ++k) z[i][k] = make_float2(0.f, 0.f);
  }
  float ov[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ov[i] = 0.f;
  float mx = -INFINITY, sm = 0.f;

  const size_t prow = (size_t)tok * Ktot;
  const KvT* kb0 = kv0 + (size_t)(b / div0) * T0 * ldkv0;
  const KvT* kb1 = (K1 > 0) ? kv1 + (size_t)(b / div1) * T1 * ldkv1 : kb0;

  for (int c0 = 0; c0 < Ktot; c0 += 32) {
    // ---- stage the chunk: compact the unmasked neighbours (ballot), pad the last group with weight-0 dummies
    const int j = c0 + lane;
    bool valid = false;
    const KvT* rptr = kb0;
    float rx = 0.f, ry = 0.f, rw = 0.f;
    if (j < Ktot) {
      const size_t p = prow + j;
      valid = invalid[p] == 0;
      const int id = idx[p];
      rptr = (j < K0) ? kb0 + (size_t)id * ldkv0 : kb1 + (size_t)id * ldkv1;
      if (!FROM_EMB) { rx = rel[p * 3 + 0]; ry = rel[p * 3 + 1]; rw = rel[p * 3 + 2]; }
    }
    const unsigned vb = __ballot_sync(TB_FULL_MASK, valid);
    const int nvalid = __popc(vb);
    if (nvalid == 0) continue;  // warp-uniform
    __syncwarp();
    const int pos = valid ? __popc(vb & lt_mask) : nvalid + __popc(~vb & lt_mask);  // invalid lanes fill the tail
    s_ptr[warp][pos] = valid ? rptr : kb0;
    s_j[warp][pos] = valid ? j : 0;
    if (!FROM_EMB) {
      s_rel[warp][pos][0] = valid ? rx : 0.f;
      s_rel[warp][pos][1] = valid ? ry : 0.f;
      s_rel[warp][pos][2] = valid ? rw : 0.f;
    }
    __syncwarp();

    for (int g0 = 0; g0 < nvalid; g0 += G) {  // 32 % G == 0: slots g0..g0+G-1 are always staged
      // ---- L1 prefetch of the NEXT group's K|V rows (G rows x 2*D*4 B = 32 lines for D=128: one line per lane), so
      // that its gathers hit L1 instead of stalling on L2 (ncu v4: 21 % long-scoreboard stalls)
      if (D == 128 && !KV_H && g0 + G < nvalid) {
        const char* pf = reinterpret_cast<const char*>(s_ptr[warp][g0 + G + (lane >> 3)]) + (lane & 7) * 128;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
      }
      // ---- gather: K and V rows of the group in flight before first use
      float kr[G][NV], vr[G][NV];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (KV_H) {  // NV halves of K and of V per lane: one 8-byte (D = 128) or 16-byte (D = 256) load each
          const __half* rp = reinterpret_cast<const __half*>(s_ptr[warp][g0 + g]) + lane * NV;
#pragma unroll
          for (int i = 0; i < NV; i += 4) {
            const uint2 a = __ldg(reinterpret_cast<const uint2*>(rp + i));
            const uint2 c = __ldg(reinterpret_cast<const uint2*>(rp + D + i));
            const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x));
            const float2 a1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
            const float2 c0_ = __half22float2(*reinterpret_cast<const __half2*>(&c.x));
            const float2 c1_ = __half22float2(*reinterpret_cast<const __half2*>(&c.y));
            kr[g][i] = a0.x; kr[g][i + 1] = a0.y; kr[g][i + 2] = a1.x; kr[g][i + 3] = a1.y;
            vr[g][i] = c0_.x; vr[g][i + 1] = c0_.y; vr[g][i + 2] = c1_.x; vr[g][i + 3] = c1_.y;
          }
        } else {
          const float* rp = reinterpret_cast<const float*>(s_ptr[warp][g0 + g]) + lane * NV;
#pragma unroll
          for (int i = 0; i < NV; i += 4) {
            float4 t = ldg4(rp + i);
            kr[g][i] = t.x; kr[g][i + 1] = t.y; kr[g][i + 2] = t.z; kr[g][i + 3] = t.w;
            float4 w = ldg4(rp + D + i);
            vr[g][i] = w.x; vr[g][i + 1] = w.y; vr[g][i + 2] = w.z; vr[g][i + 3] = w.w;
          }
        }
      }

      // ---- embedding, per-head partial logits, reductions: G independent chains
      float2 e[G][NC / 2];
      float lg[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (FROM_EMB) {
          const float* ep = emb + (prow + s_j[warp][g0 + g]) * D + lane;
#pragma unroll
          for (int k = 0; k < NC / 2; ++k) e[g][k] = make_float2(__ldg(ep + 64 * k), __ldg(ep + 64 * k + 32));
        } else {
          const float x = s_rel[warp][g0 + g][0], y = s_rel[warp][g0 + g][1], w = s_rel[warp][g0 + g][2];
          if (D == 128) {
            if (FAST_TRIG) {  // SFU range reduction only (abs error grows ~|arg| * 2^-22: <= ~2e-5 at 150 rad)
              const float aw = w * m1;
              e[g][0] = make_float2(__sinf(fmaf(x, fxy, ph)), __sinf(fmaf(y, fxy, ph)));
              e[g][1 % (NC / 2)] = make_float2(__cosf(aw), __sinf(aw));
            } else {
              const float aw = tb_reduce_2pi(w * m1);
              e[g][0] = make_float2(__sinf(tb_reduce_2pi(x * fxy) + ph), __sinf(tb_reduce_2pi(y * fxy) + ph));
              e[g][1 % (NC / 2)] = make_float2(__cosf(aw), __sinf(aw));
            }
          } else {
            const float ax = FAST_TRIG ? x * fxy : tb_reduce_2pi(x * fxy), ay = FAST_TRIG ? y * fxy : tb_reduce_2pi(y * fxy);
            const float a1 = FAST_TRIG ? w * m1 : tb_reduce_2pi(w * m1), a2 = FAST_TRIG ? w * m2 : tb_reduce_2pi(w * m2);
            e[g][0] = make_float2(__cosf(ax), __sinf(ax));
            e[g][1 % (NC / 2)] = make_float2(__cosf(ay), __sinf(ay));
            e[g][2 % (NC / 2)] = make_float2(__cosf(a1), __cosf(a2));
            e[g][3 % (NC / 2)] = make_float2(__sinf(a1), __sinf(a2));
          }
        }
        // head-pair FFMA2 with the component as broadcast scalar: slots (0,1) and (2,3)
        float2 p01 = make_float2(0.f, 0.f), p23 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < NC / 2; ++k) {
          const float2 ex = make_float2(e[g][k].x, e[g][k].x), ey = make_float2(e[g][k].y, e[g][k].y);
          p01 = __ffma2_rn(ex, u01[2 * k], p01);
          p23 = __ffma2_rn(ex, u23[2 * k], p23);
          p01 = __ffma2_rn(ey, u01[2 * k + 1], p01);
          p23 = __ffma2_rn(ey, u23[2 * k + 1], p23);
        }
        // select-free halving butterfly over the permuted head slots (slot i = head i ^ hh)
        p01.x += __shfl_xor_sync(TB_FULL_MASK, p23.x, 16);
        p01.y += __shfl_xor_sync(TB_FULL_MASK, p23.y, 16);
        float t = p01.x + __shfl_xor_sync(TB_FULL_MASK, p01.y, 8);
#pragma unroll
        for (int i = 0; i < NV; ++i) t = fmaf(qr[i], kr[g][i], t);
        t += __shfl_xor_sync(TB_FULL_MASK, t, 4);
        t += __shfl_xor_sync(TB_FULL_MASK, t, 2);
        t += __shfl_xor_sync(TB_FULL_MASK, t, 1);
        lg[g] = (g0 + g < nvalid) ? t : -INFINITY;
      }

      // ---- softmax of the group for the lane's own head; one rescale per group
      float gm = lg[0];
#pragma unroll
      for (int g = 1; g < G; ++g) gm = fmaxf(gm, lg[g]);
      const float mn = fmaxf(mx, gm);  // finite: slot g0 is a valid neighbour
      float pg[G], ps = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) { pg[g] = ex2(lg[g] - mn); ps += pg[g]; }
      if (__any_sync(TB_FULL_MASK, mn > mx)) {  // lazy rescale: only when some head's running max grew (warp-uniform)
        const float corr = ex2(mx - mn);        // 1 for the heads whose max did not move, 0 on the first group
        mx = mn;
        sm *= corr;
        // rescale factors of the other heads: slot i lives in lane ^ (8 i)
        float cs[H];
        cs[0] = corr;
#pragma unroll
        for (int i = 1; i < H; ++i) cs[i] = __shfl_xor_sync(TB_FULL_MASK, corr, 8 * i);
#pragma unroll
        for (int i = 0; i < H; ++i) {
          const float2 c2 = make_float2(cs[i], cs[i]);
#pragma unroll
          for (int k = 0; k < NC / 2; ++k) z[i][k] = __fmul2_rn(z[i][k], c2);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) ov[i] *= corr;
      }
      sm += ps;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float pi[H];
        pi[0] = pg[g];
#pragma unroll
        for (int i = 1; i < H; ++i) pi[i] = __shfl_xor_sync(TB_FULL_MASK, pg[g], 8 * i);
#pragma unroll
        for (int i = 0; i < H; ++i) {
          const float2 p2 = make_float2(pi[i], pi[i]);
#pragma unroll
          for (int k = 0; k < NC / 2; ++k) z[i][k] = __ffma2_rn(p2, e[g][k], z[i][k]);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) ov[i] = fmaf(pg[g], vr[g][i], ov[i]);
      }
    }
  }

  // ---- normalise + store (all-masked row: zeros, attention_rpe.py:188-190)
  const float inv_o = sm > 0.f ? 1.f / sm : 0.f;
  if (OUT_H) {
    __half* op = static_cast<__half*>(out_ov_) + (size_t)tok * ldo + lane * NV;
#pragma unroll
    for (int i = 0; i < NV; i += 2) *reinterpret_cast<__half2*>(op + i) = __floats2half2_rn(ov[i] * inv_o, ov[i + 1] * inv_o);
    __half* zp = static_cast<__half*>(out_z_) + (size_t)tok * ldo + lane;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float inv_i = i == 0 ? inv_o : __shfl_xor_sync(TB_FULL_MASK, inv_o, 8 * i);
#pragma unroll
      for (int k = 0; k < NC / 2; ++k) {
        zp[(i ^ hh) * D + 64 * k] = __float2half_rn(z[i][k].x * inv_i);
        zp[(i ^ hh) * D + 64 * k + 32] = __float2half_rn(z[i][k].y * inv_i);
      }
    }
  } else {
    float* op = static_cast<float*>(out_ov_) + (size_t)tok * ldo + lane * NV;
#pragma unroll
    for (int i = 0; i < NV; i += 4)
      *reinterpret_cast<float4*>(op + i) =
          make_float4(ov[i] * inv_o, ov[i + 1] * inv_o, ov[i + 2] * inv_o, ov[i + 3] * inv_o);
    float* zp = static_cast<float*>(out_z_) + (size_t)tok * ldo + lane;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float inv_i = i == 0 ? inv_o : __shfl_xor_sync(TB_FULL_MASK, inv_o, 8 * i);
#pragma unroll
      for (int k = 0; k < NC / 2; ++k) {
        zp[(i ^ hh) * D + 64 * k] = z[i][k].x * inv_i;
        zp[(i ^ hh) * D + 64 * k + 32] = z[i][k].y * inv_i;
      }
    }
  }
  if (lane == 0 && out_none_valid) out_none_valid[tok] = sm > 0.f ? 0 : 1;
}

template <int D, bool FROM_EMB, bool FAST_TRIG, bool OUT_H = false, bool IN_H = false, bool KV_H = false>
int launch(const void* q, int ldq, const void* u, int ldu, const void* kv0, int ldkv0, int T0, int div0, int K0,
           const void* kv1, int ldkv1, int T1, int div1, int K1, const int32_t* idx, const uint8_t* invalid,
           const float* rel, const float* emb, const float* pe_freq_xy, int B, int S, void* out_ov, void* out_z,
           int ldo, uint8_t* out_none_valid, cudaStream_t st) {
  const int n_tok = B * S;
  const int grid = (n_tok + kWarps - 1) / kWarps;
  knarpe_attn_kernel<D, FROM_EMB, FAST_TRIG, OUT_H, IN_H, KV_H><<<grid, kWarps * 32, 0, st>>>(q, ldq, u, ldu, kv0, ldkv0, T0, div0, K0, kv1, ldkv1,
                                                              T1, div1, K1, idx, invalid, rel, emb, pe_freq_xy, n_tok,
                                                              S, out_ov, out_z, ldo, out_none_valid);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

}  // namespace

// tensor-core variant (knarpe_attn_mma.cu)
int tb_knarpe_attn_mma_launch(const void* q, int ldq, const void* u, int ldu, int in_f16, const void* kv0, int ldkv0, int T0,
                              int div0, int K0, const void* kv1, int ldkv1, int T1, int div1, int K1,
                              const int32_t* idx, const uint8_t* invalid, const float* rel, const float* pe_freq_xy,
                              int B, int S, void* out_ov, void* out_z, int ldo, int out_f16, uint8_t* out_none_valid,
                              int interleaved, cudaStream_t st);
bool tb_knarpe_attn_mma_supported(int D, int Hh, int Ktot);

extern "C" int tb_knarpe_attn(const void* q, int ldq, const void* u, int ldu, const void* kv0_, int ldkv0, int T0,
                              int div0, int K0, const void* kv1_, int ldkv1, int T1, int div1, int K1,
                              const int32_t* idx, const uint8_t* invalid, const float* rel, const float* emb,
                              const float* pe_freq_xy, int B, int S, int D, int Hh, void* out_ov, void* out_z,
                              int ldo, uint8_t* out_none_valid, int flags, void* stream) {
  const void* kv0 = kv0_;
  const void* kv1 = kv1_;
  if (!q || !u || !kv0 || !idx || !invalid || !out_ov || !out_z || !pe_freq_xy) return TB_ERR_NULL;
  if ((rel == nullptr) == (emb == nullptr)) return TB_ERR_NULL;  // exactly one
  if (B <= 0 || S <= 0 || K0 <= 0 || K1 < 0 || T0 <= 0 || div0 <= 0 || (K1 > 0 && (!kv1 || T1 <= 0 || div1 <= 0)))
    return TB_ERR_BAD_SHAPE;
  if (Hh != 4 || (D != 128 && D != 256)) return TB_ERR_UNSUPPORTED;
  if (ldq < D || ldu < Hh * D || ldkv0 < 2 * D || (K1 > 0 && ldkv1 < 2 * D) || ldo < D) return TB_ERR_BAD_SHAPE;
  if ((ldq | ldu | ldkv0 | ldo | (K1 > 0 ? ldkv1 : 0)) & 3) return TB_ERR_MISALIGNED;
  if (!tb_aligned16(q) || !tb_aligned16(u) || !tb_aligned16(kv0) || (K1 > 0 && !tb_aligned16(kv1)) ||
      !tb_aligned16(out_ov) || !tb_aligned16(out_z) || (emb && !tb_aligned16(emb)))
    return TB_ERR_MISALIGNED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define TB_ATT_ARGS q, ldq, u, ldu, kv0, ldkv0, T0, div0, K0, kv1, ldkv1, T1, div1, K1, idx, invalid, rel, emb, \
                    pe_freq_xy, B, S, out_ov, out_z, ldo, out_none_valid, st
  const bool fast = (flags & 1) != 0;  // bit 0: SFU-only range reduction of the embedding angles (tensor-core mode)
  const int out_h = (flags & 4) != 0;  // bit 2: out_ov / out_z are fp16 rows (ldo in halves)
  const int in_h = (flags & 8) != 0;   // bit 3: q / u are fp16 rows (ldq, ldu in halves)
  if ((out_h && (ldo & 7)) || (in_h && ((ldq | ldu) & 7))) return TB_ERR_MISALIGNED;
  if (flags & 2) {  // bit 1: fp16 K|V tables: all contractions on mma.sync (knarpe_attn_mma.cu) where that kernel applies
    if (!rel) return TB_ERR_UNSUPPORTED;
    if ((ldkv0 | (K1 > 0 ? ldkv1 : 0)) & 7) return TB_ERR_MISALIGNED;
    if (!tb_knarpe_attn_mma_supported(D, Hh, K0 + K1) || !in_h) {
      // d_model 256 (BASELINE config 2), lists longer than 128 or fp32 q / u rows: the SIMT kernel on fp16 tables
      // (fp32 arithmetic)
      if ((flags & 16) || !fast || out_h != in_h) return TB_ERR_UNSUPPORTED;
      if (D == 128)
        return out_h ? launch<128, false, true, true, true, true>(TB_ATT_ARGS)
                     : launch<128, false, true, false, false, true>(TB_ATT_ARGS);
      return out_h ? launch<256, false, true, true, true, true>(TB_ATT_ARGS)
                   : launch<256, false, true, false, false, true>(TB_ATT_ARGS);
    }
    const int il = (flags & 16) != 0;  // bit 4: head-interleaved q / K / V channels, rows read 32 bytes per lane
    if (il) {
      if (!in_h) return TB_ERR_UNSUPPORTED;
      if (((ldkv0 | (K1 > 0 ? ldkv1 : 0)) & 15) || (reinterpret_cast<uintptr_t>(kv0_) & 31) ||
          (K1 > 0 && (reinterpret_cast<uintptr_t>(kv1_) & 31)))
        return TB_ERR_MISALIGNED;
    }
    return tb_knarpe_attn_mma_launch(q, ldq, u, ldu, in_h, kv0_, ldkv0, T0, div0, K0, kv1_, ldkv1, T1, div1, K1, idx,
                                     invalid, rel, pe_freq_xy, B, S, out_ov, out_z, ldo, out_h, out_none_valid, il, st);
  }
  if (flags & 16) return TB_ERR_UNSUPPORTED;
  if (out_h || in_h) {  // fp16 intermediates with the SIMT kernel: the tensor-core mode's short neighbour lists
    if (D != 128 || !rel || !fast || !out_h) return TB_ERR_UNSUPPORTED;
    return in_h ? launch<128, false, true, true, true>(TB_ATT_ARGS) : launch<128, false, true, true, false>(TB_ATT_ARGS);
  }
  if (D == 128) {
    if (!rel) return launch<128, true, false>(TB_ATT_ARGS);
    return fast ? launch<128, false, true>(TB_ATT_ARGS) : launch<128, false, false>(TB_ATT_ARGS);
  }
  if (!rel) return launch<256, true, false>(TB_ATT_ARGS);
  return fast ? launch<256, false, true>(TB_ATT_ARGS) : launch<256, false, false>(TB_ATT_ARGS);
#undef TB_ATT_ARGS
}
