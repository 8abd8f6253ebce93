// KNARPE attention core (tb_knarpe_attn): neighbour gather + relative-pose bias + masked softmax + weighted sums.
// Reference behaviour: modules/attention_rpe.py:137-190 (RPE branch) between the projections.
//
// Re-associated form (exact in real arithmetic):
//   q_h.(k_hj + W_rk,h e_j + b_rk,h) = q_h.k_hj + (W_rk,h^T q_h).e_j + const_h   (const_h cancels in the softmax)
//   sum_j a_hj (v_hj + W_rv,h e_j + b_rv,h) = sum_j a_hj v_hj + W_rv,h (sum_j a_hj e_j) + b_rv,h
// so the kernel needs no weights: it consumes q, u_h = W_rk,h^T q_h (both pre-scaled by log2(e)/sqrt(d_head) by the
// caller's projection) and emits ov = sum a v and z_h = sum a e; the dense parts stay in the projections.
//
// Mapping (v2, from the ncu instruction mix of v1 — profiles/r1_notes.md): one warp per source token.
//   * q/k/v: lane l owns feature columns [l*D/32, (l+1)*D/32) => head l>>3; a K or V row is one fully coalesced
//     512 B / 1 KiB request; the rows of G neighbours are in flight before first use.
//   * relative-pose embedding e_j (D sin/cos components, utils/pose_emb.py:50-55): evaluated cooperatively, lane l
//     computes components {l + 32k} in registers from the 12-byte relative pose (2-term Cody-Waite reduction + SFU),
//     then exchanged through a bank-conflict-free shared-memory transpose so that lane (h = l>>3, s = l&7) holds the
//     D/8 components {s + 8r} it needs for ITS head only. The RPE logit term and z_h therefore share the 8-lane
//     reduction of q.k (3 shuffles per neighbour), the softmax is evaluated once per head (not once per lane x head),
//     and the fp32 multiply-adds are issued as packed FFMA2.
//   * online softmax with lazy rescale (accumulators are rescaled only when a head's running max grows).
#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int H = 4;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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
  constexpr int NV = D / 32;            // q/k/v floats per lane
  constexpr int NC = D / 32;            // embedding components a lane COMPUTES (l + 32k)
  constexpr int NO = D / 8;             // embedding components a lane OWNS for its head (s + 8r)
  constexpr int G = (D == 128) ? 4 : 2; // neighbours with K/V rows in flight
  constexpr int NF = D / 8;             // number of xy frequencies
  constexpr int ES = NO + 4;            // padded stride of the transpose (conflict-free, see header)
  __shared__ int s_idx[kWarps][32];
  __shared__ float s_rel[kWarps][32][3];
  __shared__ __align__(16) float s_e[kWarps][G][8 * ES];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * kWarps + warp;
  if (tok >= n_tok) return;  // warp-uniform; only __syncwarp below
  const int b = tok / S;
  const int Ktot = K0 + K1;
  const int hh = lane >> 3, sub = lane & 7;

  // per-lane constants of the components this lane computes
  float fxy, ph;
  if (D == 128) {  // one x and one y component per lane: frequency lane&15, cos for lane<16 else sin
    fxy = __ldg(pe_freq_xy + (lane & (NF - 1)));
    ph = (lane < NF) ? 1.57079632679489662f : 0.f;
  } else {
    fxy = __ldg(pe_freq_xy + lane);
    ph = 0.f;
  }
  const float m1 = (float)(lane + 1), m2 = (float)(lane + 33);
  const int wpos = sub * ES + hh;  // transpose: component c -> (c & 7) * ES + (c >> 3); c = lane + 32k -> wpos + 4k

  float qr[NV];
  float2 uo[NO / 2], zo[NO / 2];
  {
    const float* qp = q + (size_t)tok * ldq + lane * NV;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 t = ldg4(qp + i);
      qr[i] = t.x; qr[i + 1] = t.y; qr[i + 2] = t.z; qr[i + 3] = t.w;
    }
    const float* up = u + (size_t)tok * ldu + hh * D + sub;
#pragma unroll
    for (int r = 0; r < NO / 2; ++r) {
      uo[r] = make_float2(__ldg(up + 16 * r), __ldg(up + 16 * r + 8));
      zo[r] = make_float2(0.f, 0.f);
    }
  }
  float ov[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ov[i] = 0.f;
  float mx = -INFINITY, sm = 0.f;

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
        }
      }

      // ---- embedding components {lane + 32k} of the group's neighbours -> transposed into shared memory
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (id[g] < 0) continue;  // warp-uniform
        float ec[NC];
        if (FROM_EMB) {
          const float* ep = emb + (prow + c0 + g0 + g) * D + lane;
#pragma unroll
          for (int k = 0; k < NC; ++k) ec[k] = __ldg(ep + 32 * k);
        } else {
          const float x = s_rel[warp][g0 + g][0], y = s_rel[warp][g0 + g][1], w = s_rel[warp][g0 + g][2];
          if (D == 128) {
            ec[0] = __sinf(tb_reduce_2pi(x * fxy) + ph);
            ec[1] = __sinf(tb_reduce_2pi(y * fxy) + ph);
            const float rw = tb_reduce_2pi(w * m1);
            ec[2] = __cosf(rw);
            ec[3] = __sinf(rw);
          } else {
            const float rx = tb_reduce_2pi(x * fxy), ry = tb_reduce_2pi(y * fxy);
            const float r1 = tb_reduce_2pi(w * m1), r2 = tb_reduce_2pi(w * m2);
            ec[0] = __cosf(rx); ec[1] = __sinf(rx);
            ec[2] = __cosf(ry); ec[3] = __sinf(ry);
            ec[4 % NC] = __cosf(r1); ec[5 % NC] = __cosf(r2);
            ec[6 % NC] = __sinf(r1); ec[7 % NC] = __sinf(r2);
          }
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) s_e[warp][g][wpos + 4 * k] = ec[k];
      }
      __syncwarp();

      // ---- per neighbour: own-head logit (q.k + u_h.e_j), 8-lane reduction, online softmax, accumulate
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (id[g] < 0) continue;  // warp-uniform
        float2 eo[NO / 2];
        const float4* sp = reinterpret_cast<const float4*>(&s_e[warp][g][sub * ES]);
#pragma unroll
        for (int r = 0; r < NO / 4; ++r) {
          const float4 t = sp[r];  // components sub + 8*(4r .. 4r+3)
          eo[2 * r] = make_float2(t.x, t.y);
          eo[2 * r + 1] = make_float2(t.z, t.w);
        }
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < NO / 2; ++r) acc = __ffma2_rn(uo[r], eo[r], acc);
        float t = acc.x + acc.y;
#pragma unroll
        for (int i = 0; i < NV; ++i) t = fmaf(qr[i], kr[g][i], t);
        t += __shfl_xor_sync(TB_FULL_MASK, t, 4);
        t += __shfl_xor_sync(TB_FULL_MASK, t, 2);
        t += __shfl_xor_sync(TB_FULL_MASK, t, 1);
        if (__any_sync(TB_FULL_MASK, t > mx)) {  // lazy rescale: some head's running max grows
          const float mn = fmaxf(mx, t);
          const float corr = ex2(mx - mn);  // mx = -inf -> 0
          mx = mn;
          sm *= corr;
          const float2 c2 = make_float2(corr, corr);
#pragma unroll
          for (int r = 0; r < NO / 2; ++r) zo[r] = __fmul2_rn(zo[r], c2);
#pragma unroll
          for (int i = 0; i < NV; ++i) ov[i] *= corr;
        }
        const float p = ex2(t - mx);
        sm += p;
        const float2 p2 = make_float2(p, p);
#pragma unroll
        for (int r = 0; r < NO / 2; ++r) zo[r] = __ffma2_rn(p2, eo[r], zo[r]);
#pragma unroll
        for (int i = 0; i < NV; ++i) ov[i] = fmaf(p, vr[g][i], ov[i]);
      }
      __syncwarp();  // s_e is rewritten by the next group
    }
  }

  // ---- normalise + store (all-masked row: zeros, attention_rpe.py:188-190)
  const float inv_s = sm > 0.f ? 1.f / sm : 0.f;
  float* op = out_ov + (size_t)tok * ldo + lane * NV;
#pragma unroll
  for (int i = 0; i < NV; i += 4)
    *reinterpret_cast<float4*>(op + i) =
        make_float4(ov[i] * inv_s, ov[i + 1] * inv_s, ov[i + 2] * inv_s, ov[i + 3] * inv_s);
  // z_h[c], c = sub + 8r, r = 2*rr (+1): eo pairs hold r = 4k+{0,1} / 4k+{2,3}  => zo[r2] = (c = sub + 8*(2*r2), +8)
  float* zp = out_z + (size_t)tok * ldo + hh * D + sub;
#pragma unroll
  for (int r = 0; r < NO / 2; ++r) {
    zp[16 * r] = zo[r].x * inv_s;
    zp[16 * r + 8] = zo[r].y * inv_s;
  }
  if (lane == 0 && out_none_valid) out_none_valid[tok] = sm > 0.f ? 0 : 1;
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
