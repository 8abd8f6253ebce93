// KNARPE attention core (tb_knarpe_attn): neighbour gather + relative-pose bias + masked softmax + weighted sums.
// Reference behaviour: modules/attention_rpe.py:137-190 (RPE branch) between the projections.
//
// Re-associated form (exact in real arithmetic):
//   q_h.(k_hj + W_rk,h e_j + b_rk,h) = q_h.k_hj + (W_rk,h^T q_h).e_j + const_h   (const_h cancels in the softmax)
//   sum_j a_hj (v_hj + W_rv,h e_j + b_rv,h) = sum_j a_hj v_hj + W_rv,h (sum_j a_hj e_j) + b_rv,h
// so the kernel needs no weights: it consumes q, u_h = W_rk,h^T q_h (both pre-scaled by log2(e)/sqrt(d_head) by the
// caller's projection) and emits ov = sum a v and z_h = sum a e; the dense parts stay in the projections.
//
// Mapping: one warp per source token; lane l owns feature columns [l*D/32, (l+1)*D/32) of q/k/v (=> head l/8, a
// K or V row is one fully coalesced 512 B / 1 KiB request) and embedding components {l + 32 r}. The 128/256-d
// sin/cos relative-pose embedding is evaluated in registers from the 12-byte relative pose (2-term Cody-Waite
// reduction + SFU sin/cos), never materialised. Neighbours are processed in groups of G with all K/V rows of a
// group in flight before use, and one online-softmax rescale per group.
#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int H = 4;

template <int D, bool FROM_EMB>
__global__ void __launch_bounds__(kWarps * 32)
knarpe_attn_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ u, int ldu,
                   const float* __restrict__ kv0, int ldkv0, int T0, int div0, int K0,
                   const float* __restrict__ kv1, int ldkv1, int T1, int div1, int K1,
                   const int32_t* __restrict__ idx, const uint8_t* __restrict__ invalid,
                   const float* __restrict__ rel, const float* __restrict__ emb,
                   const float* __restrict__ pe_freq_xy, int n_tok, int S,
                   float* __restrict__ out_ov, float* __restrict__ out_z, int ldo,
                   uint8_t* __restrict__ out_none_valid) {
  constexpr int NV = D / 32;          // q/k/v floats per lane
  constexpr int NR = D / 32;          // embedding components per lane
  constexpr int G = (D == 128) ? 4 : 2;
  constexpr int NF = D / 8;           // xy frequencies
  __shared__ int s_idx[kWarps][32];
  __shared__ float s_rel[kWarps][32][3];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * kWarps + warp;
  if (tok >= n_tok) return;  // warp-uniform; no block-level sync below
  const int b = tok / S;
  const int Ktot = K0 + K1;
  const int hh = lane >> 3;  // own head

  // per-lane constants of the embedding
  float fxy, ph;  // D=128: one x and one y component per lane: freq index lane&15, cos for lane<16 else sin
  if (D == 128) {
    fxy = __ldg(pe_freq_xy + (lane & (NF - 1)));
    ph = (lane < NF) ? 1.57079632679489662f : 0.f;
  } else {
    fxy = __ldg(pe_freq_xy + lane);
    ph = 0.f;
  }
  const float m1 = (float)(lane + 1), m2 = (float)(lane + 33);

  float qr[NV], ur[H][NR];
  {
    const float* qp = q + (size_t)tok * ldq + lane * NV;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 t = ldg4(qp + i);
      qr[i] = t.x; qr[i + 1] = t.y; qr[i + 2] = t.z; qr[i + 3] = t.w;
    }
    const float* up = u + (size_t)tok * ldu;
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
      for (int r = 0; r < NR; ++r) ur[h][r] = __ldg(up + h * D + lane + 32 * r);
  }

  float ov[NV], z[H][NR], mx[H], sm[H];
#pragma unroll
  for (int i = 0; i < NV; ++i) ov[i] = 0.f;
#pragma unroll
  for (int h = 0; h < H; ++h) {
    mx[h] = -INFINITY; sm[h] = 0.f;
#pragma unroll
    for (int r = 0; r < NR; ++r) z[h][r] = 0.f;
  }

  const size_t prow = (size_t)tok * Ktot;
  const float* kb0 = kv0 + (size_t)(b / div0) * T0 * ldkv0 + lane * NV;
  const float* kb1 = (K1 > 0) ? kv1 + (size_t)(b / div1) * T1 * ldkv1 + lane * NV : kb0;

  for (int c0 = 0; c0 < Ktot; c0 += 32) {
    const int cnt = min(32, Ktot - c0);
    __syncwarp();
    if (lane < cnt) {
      const size_t p = prow + c0 + lane;
      s_idx[warp][lane] = invalid[p] ? -1 : idx[p];
      if (!FROM_EMB) {
        s_rel[warp][lane][0] = rel[p * 3 + 0];
        s_rel[warp][lane][1] = rel[p * 3 + 1];
        s_rel[warp][lane][2] = rel[p * 3 + 2];
      }
    }
    __syncwarp();

    for (int g0 = 0; g0 < cnt; g0 += G) {
      int id[G];
      bool any = false;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        id[g] = (g0 + g < cnt) ? s_idx[warp][g0 + g] : -1;
        any |= id[g] >= 0;
      }
      if (!any) continue;  // warp-uniform

      // ---- gather: all K and V rows of the group in flight before first use
      float kr[G][NV], vr[G][NV];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (id[g] >= 0) {
          const int j = c0 + g0 + g;
          const float* rp = (j < K0) ? kb0 + (size_t)id[g] * ldkv0 : kb1 + (size_t)id[g] * ldkv1;
#pragma unroll
          for (int i = 0; i < NV; i += 4) {
            float4 t = ldg4(rp + i);
            kr[g][i] = t.x; kr[g][i + 1] = t.y; kr[g][i + 2] = t.z; kr[g][i + 3] = t.w;
            float4 w = ldg4(rp + D + i);
            vr[g][i] = w.x; vr[g][i + 1] = w.y; vr[g][i + 2] = w.z; vr[g][i + 3] = w.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < NV; ++i) { kr[g][i] = 0.f; vr[g][i] = 0.f; }
        }
      }

      // ---- embedding + logits
      float e[G][NR], lg[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (id[g] >= 0) {
          if (FROM_EMB) {
            const float* ep = emb + (prow + c0 + g0 + g) * D + lane;
#pragma unroll
            for (int r = 0; r < NR; ++r) e[g][r] = __ldg(ep + 32 * r);
          } else {
            const float x = s_rel[warp][g0 + g][0], y = s_rel[warp][g0 + g][1], w = s_rel[warp][g0 + g][2];
            if (D == 128) {
              e[g][0] = __sinf(tb_reduce_2pi(x * fxy) + ph);
              e[g][1] = __sinf(tb_reduce_2pi(y * fxy) + ph);
              const float rw = tb_reduce_2pi(w * m1);
              e[g][2] = __cosf(rw);
              e[g][3] = __sinf(rw);
            } else {
              const float rx = tb_reduce_2pi(x * fxy), ry = tb_reduce_2pi(y * fxy);
              const float r1 = tb_reduce_2pi(w * m1), r2 = tb_reduce_2pi(w * m2);
              e[g][0] = __cosf(rx); e[g][1] = __sinf(rx);
              e[g][2] = __cosf(ry); e[g][3] = __sinf(ry);
              e[g][4 % NR] = __cosf(r1); e[g][5 % NR] = __cosf(r2);
              e[g][6 % NR] = __sinf(r1); e[g][7 % NR] = __sinf(r2);
            }
          }
          float p[H];
#pragma unroll
          for (int h = 0; h < H; ++h) {
            float a = 0.f;
#pragma unroll
            for (int r = 0; r < NR; ++r) a = fmaf(ur[h][r], e[g][r], a);
            p[h] = a;
          }
          float qk = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) qk = fmaf(qr[i], kr[g][i], qk);
          // 4 partial sums over 32 lanes -> lane keeps head (lane>>3): halving butterfly, 6 shuffles
          const bool b4 = lane & 16, b3 = lane & 8;
          float a0 = b4 ? p[2] : p[0], a1 = b4 ? p[3] : p[1];
          const float s0 = b4 ? p[0] : p[2], s1 = b4 ? p[1] : p[3];
          a0 += __shfl_xor_sync(TB_FULL_MASK, s0, 16);
          a1 += __shfl_xor_sync(TB_FULL_MASK, s1, 16);
          float t = (b3 ? a1 : a0) + __shfl_xor_sync(TB_FULL_MASK, b3 ? a0 : a1, 8);
          t += qk;
          t += __shfl_xor_sync(TB_FULL_MASK, t, 4);
          t += __shfl_xor_sync(TB_FULL_MASK, t, 2);
          t += __shfl_xor_sync(TB_FULL_MASK, t, 1);
          lg[g] = t;
        } else {
          lg[g] = -INFINITY;
#pragma unroll
          for (int r = 0; r < NR; ++r) e[g][r] = 0.f;
        }
      }

      // ---- online softmax: one rescale per group and head
      float corr_o = 1.f, p_o[G];
#pragma unroll
      for (int g = 0; g < G; ++g) p_o[g] = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float l[G], gm = -INFINITY;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          l[g] = __shfl_sync(TB_FULL_MASK, lg[g], h * 8);
          gm = fmaxf(gm, l[g]);
        }
        const float mn = fmaxf(mx[h], gm);  // finite: the group has >= 1 valid neighbour
        const float corr = exp2f(mx[h] - mn);
        mx[h] = mn;
        float ps = 0.f, pg[G];
#pragma unroll
        for (int g = 0; g < G; ++g) { pg[g] = exp2f(l[g] - mn); ps += pg[g]; }
        sm[h] = fmaf(sm[h], corr, ps);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          float a = z[h][r] * corr;
#pragma unroll
          for (int g = 0; g < G; ++g) a = fmaf(pg[g], e[g][r], a);
          z[h][r] = a;
        }
        if (hh == h) {
          corr_o = corr;
#pragma unroll
          for (int g = 0; g < G; ++g) p_o[g] = pg[g];
        }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float a = ov[i] * corr_o;
#pragma unroll
        for (int g = 0; g < G; ++g) a = fmaf(p_o[g], vr[g][i], a);
        ov[i] = a;
      }
    }
  }

  // ---- normalise + store (all-masked row: zeros, attention_rpe.py:188-190)
  float inv_s[H];
#pragma unroll
  for (int h = 0; h < H; ++h) inv_s[h] = sm[h] > 0.f ? 1.f / sm[h] : 0.f;
  const float inv_o = hh == 0 ? inv_s[0] : hh == 1 ? inv_s[1] : hh == 2 ? inv_s[2] : inv_s[3];
  float* op = out_ov + (size_t)tok * ldo + lane * NV;
#pragma unroll
  for (int i = 0; i < NV; i += 4)
    *reinterpret_cast<float4*>(op + i) =
        make_float4(ov[i] * inv_o, ov[i + 1] * inv_o, ov[i + 2] * inv_o, ov[i + 3] * inv_o);
  float* zp = out_z + (size_t)tok * ldo;
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int r = 0; r < NR; ++r) zp[h * D + lane + 32 * r] = z[h][r] * inv_s[h];
  if (lane == 0 && out_none_valid) out_none_valid[tok] = sm[0] > 0.f ? 0 : 1;
}

template <int D, bool FROM_EMB>
int launch(const float* q, int ldq, const float* u, int ldu, const float* kv0, int ldkv0, int T0, int div0, int K0,
           const float* kv1, int ldkv1, int T1, int div1, int K1, const int32_t* idx, const uint8_t* invalid,
           const float* rel, const float* emb, const float* pe_freq_xy, int B, int S, float* out_ov, float* out_z,
           int ldo, uint8_t* out_none_valid, cudaStream_t st) {
  const int n_tok = B * S;
  const int grid = (n_tok + kWarps - 1) / kWarps;
  knarpe_attn_kernel<D, FROM_EMB><<<grid, kWarps * 32, 0, st>>>(q, ldq, u, ldu, kv0, ldkv0, T0, div0, K0, kv1, ldkv1,
                                                              T1, div1, K1, idx, invalid, rel, emb, pe_freq_xy, n_tok,
                                                              S, out_ov, out_z, ldo, out_none_valid);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

}  // namespace

extern "C" int tb_knarpe_attn(const float* q, int ldq, const float* u, int ldu, const float* kv0, int ldkv0, int T0,
                              int div0, int K0, const float* kv1, int ldkv1, int T1, int div1, int K1,
                              const int32_t* idx, const uint8_t* invalid, const float* rel, const float* emb,
                              const float* pe_freq_xy, int B, int S, int D, int Hh, float* out_ov, float* out_z,
                              int ldo, uint8_t* out_none_valid, void* stream) {
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
  if (D == 128) return rel ? launch<128, false>(TB_ATT_ARGS) : launch<128, true>(TB_ATT_ARGS);
  return rel ? launch<256, false>(TB_ATT_ARGS) : launch<256, true>(TB_ATT_ARGS);
#undef TB_ATT_ARGS
}
