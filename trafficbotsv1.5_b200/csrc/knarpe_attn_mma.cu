// KNARPE attention core, tensor-core variant (tb_knarpe_attn flags bit 1; d_model = d_rpe = 128, 4 heads, fp16 K|V
// tables written by tb_linear's fp16 output). Same contract as knarpe_attn.cu (reference:
// modules/attention_rpe.py:137-190 between the projections); used by the tf32 mode of the engine, where every
// projection already rounds its operands to 11 significant bits.
//
// Why: ncu of the SIMT kernel (profiles/r1/ncu_attn_v5.summary.txt) shows ~97 warp instructions per (token, neighbour)
// pair spread over shuffles (L1 data pipe 72 %), SFU (34 %) and FFMA2 (47 %), with 1 KiB of fp32 K|V gathered per
// pair. All four contractions of a token are small GEMMs over its neighbour list:
//   logits[h, j] = sum_c q[c in head h] k_j[c] + sum_c u[h, c] e_j[c]
//   z[h, c] = sum_j p[h, j] e_j[c]          ov[c in head h] = sum_j p[h, j] v_j[c]
// Here they run on mma.sync.m16n8k16 (f16 operands, f32 accumulate) with every operand built IN FRAGMENT LAYOUT in
// registers - nothing is staged through shared memory:
//   * a group = 16 compacted neighbours = two 8-wide MMA tiles. Lane (g = lane>>2, t = lane&3) owns neighbours g and
//     g+8: it loads 16-byte pieces of their fp16 K and V rows that are exactly its B-fragment registers, and evaluates
//     the cos/sin of 16 embedding angles per neighbour - the B-fragment elements (k = channel slot, n = neighbour) of
//     the u.e MMA. The yaw harmonics of a lane form arithmetic progressions: 6 SFU evaluations + 7 plane rotations
//     replace 32 SFU evaluations; the geometric x/y frequencies are evaluated directly.
//   * logits^T [16 x 16 nbr] = A(q, block-diagonal over heads) B(k) + A(u) B(e): A rows 0-3 hold fp16(operand), rows
//     4-7 its fp16 residual (rows 8-15 zero), so q and u keep ~22 significant bits; one shuffle adds the row blocks.
//   * the accumulator layout of the logits IS the B-fragment layout of z^T [128 x 8] = A(e^T) B(p^T) and of
//     ov^T [32 x 8] = sum_heads A(v_head^T) B(p^T masked to that head): columns 0-3 take fp16(p_h), columns 4-7 the
//     residual; the transposed A fragments come from movmatrix (register-only 8x8 transposes).
// Channel "slots" inside a 16-wide MMA chunk are permutations of the reference orders (embedding:
// utils/pose_emb.py:50-55 [cos x|sin x|cos y|sin y|cos yaw|sin yaw]); q/u are read and ov/z written through the same
// permutations, so the C ABI layouts are unchanged.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

#ifndef TB_MMA_WARPS
#define TB_MMA_WARPS 4
#endif
#ifndef TB_MMA_MINB
#define TB_MMA_MINB 3
#endif
constexpr int kWarps = TB_MMA_WARPS;
constexpr int D = 128;
constexpr int H = 4;
constexpr int KMAX = 128;  // compacted neighbour slots per token (K0 + K1 rounded up to 16)
constexpr int kWarpSmem = KMAX * 8 + KMAX * 12;  // row pointers + relative poses; reused as 640-float output staging
static_assert(kWarpSmem >= (D + H * D) * 4, "output staging must fit");

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// fp16 head (lo == false) or fp16 residual (lo == true) of a pair of floats
__device__ __forceinline__ uint32_t split_h2(float a, float b, bool lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 r = __floats2half2_rn(a - f.x, b - f.y);
  const __half2 s = lo ? r : h;
  return *reinterpret_cast<const uint32_t*>(&s);
}
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movm_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ uint4 ldg128(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
// reference embedding index of slot 0 (cos half) / slot 8 (sin half) of chunk c; slots are consecutive from there
__host__ __device__ constexpr int cos_base(int c) { return c < 2 ? 8 * c : (c < 4 ? 32 + 8 * (c - 2) : 64 + 8 * (c - 4)); }
__host__ __device__ constexpr int sin_base(int c) { return cos_base(c) + (c < 4 ? 16 : 32); }

__global__ void __launch_bounds__(kWarps * 32, TB_MMA_MINB)
knarpe_attn_mma_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ u, int ldu,
                       const __half* __restrict__ kv0, int ldkv0, int T0, int div0, int K0,
                       const __half* __restrict__ kv1, int ldkv1, int T1, int div1, int K1,
                       const int32_t* __restrict__ idx, const uint8_t* __restrict__ invalid,
                       const float* __restrict__ rel, const float* __restrict__ pe_freq_xy, int n_tok, int S,
                       float* __restrict__ out_ov, float* __restrict__ out_z, int ldo,
                       uint8_t* __restrict__ out_none_valid) {
  __shared__ __align__(16) unsigned char s_raw[kWarps][kWarpSmem];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * kWarps + warp;
  if (tok >= n_tok) return;  // warp-uniform; only warp-level synchronisation below
  const __half** s_ptr = reinterpret_cast<const __half**>(s_raw[warp]);           // [KMAX] K|V row pointers
  float (*s_rel)[3] = reinterpret_cast<float (*)[3]>(s_raw[warp] + KMAX * 8);      // [KMAX] relative poses
  float* s_out = reinterpret_cast<float*>(s_raw[warp]);                            // epilogue: [ov(128) | z(512)]

  const int b = tok / S;
  const int Ktot = K0 + K1;
  const int g = lane >> 2, t = lane & 3;
  const int hA = g & 3;               // head of this lane's accumulator row / column and softmax state
  const bool lo_part = (g & 4) != 0;  // rows / columns 4-7: fp16 residual operands
  const unsigned lt_mask = (1u << lane) - 1u;

  // ---- per-token operands in A-fragment layout
  float fq[2][2];  // x/y frequencies of this lane's slots: chunk c (0/1), slot pair j
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    fq[c][0] = __ldg(pe_freq_xy + 8 * c + 2 * t);
    fq[c][1] = __ldg(pe_freq_xy + 8 * c + 2 * t + 1);
  }
  uint32_t uA[8][2];  // u: registers a0, a2 of chunk c (a1 = a3 = 0)
  {
    const float* up = u + (size_t)tok * ldu + hA * D + 2 * t;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float2 a = __ldg(reinterpret_cast<const float2*>(up + cos_base(c)));
      const float2 s = __ldg(reinterpret_cast<const float2*>(up + sin_base(c)));
      uA[c][0] = split_h2(a.x, a.y, lo_part);
      uA[c][1] = split_h2(s.x, s.y, lo_part);
    }
  }
  uint32_t qA[2][2];  // q of head hA, channels 32 hA + 8 t + [0, 8): chunk (2 hA + e) registers a0, a2
  {
    const float* qp = q + (size_t)tok * ldq + 32 * hA + 8 * t;
    const float4 a = ldg4(qp), c = ldg4(qp + 4);
    qA[0][0] = split_h2(a.x, a.y, lo_part);
    qA[0][1] = split_h2(a.z, a.w, lo_part);
    qA[1][0] = split_h2(c.x, c.y, lo_part);
    qA[1][1] = split_h2(c.z, c.w, lo_part);
  }
  float zacc[8][4], oacc[2][4];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) zacc[c][r] = 0.f;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int r = 0; r < 4; ++r) oacc[m][r] = 0.f;
  float mx = -INFINITY, sm = 0.f;

  // ---- compact the unmasked neighbours of the whole row; pad to a multiple of 16 with weight-0 dummies
  const size_t prow = (size_t)tok * Ktot;
  const __half* kb0 = kv0 + (size_t)(b / div0) * T0 * ldkv0;
  const __half* kb1 = (K1 > 0) ? kv1 + (size_t)(b / div1) * T1 * ldkv1 : kb0;
  int nvalid = 0;
  for (int c0 = 0; c0 < Ktot; c0 += 32) {
    const int j = c0 + lane;
    bool valid = false;
    if (j < Ktot) valid = invalid[prow + j] == 0;
    const unsigned vb = __ballot_sync(TB_FULL_MASK, valid);
    if (valid) {
      const size_t p = prow + j;
      const int id = idx[p];
      const int pos = nvalid + __popc(vb & lt_mask);
      s_ptr[pos] = (j < K0) ? kb0 + (size_t)id * ldkv0 : kb1 + (size_t)id * ldkv1;
      s_rel[pos][0] = rel[p * 3 + 0];
      s_rel[pos][1] = rel[p * 3 + 1];
      s_rel[pos][2] = rel[p * 3 + 2];
    }
    nvalid += __popc(vb);
  }
  const int npad = (nvalid + 15) & ~15;
  if (lane < npad - nvalid) {
    s_ptr[nvalid + lane] = kb0;
    s_rel[nvalid + lane][0] = 0.f;
    s_rel[nvalid + lane][1] = 0.f;
    s_rel[nvalid + lane][2] = 0.f;
  }
  __syncwarp();

  for (int g0 = 0; g0 < npad; g0 += 16) {
    const int nb = nvalid - g0;  // valid neighbours in this group (>= 1; may exceed 16)
    // ---- K fragments: row of neighbour g0 + 8 tile + g, 16 B at halves [32 i + 8 t, +8) = B registers of the chunks
    // 2i (x, y) and 2i+1 (z, w) of the block-diagonal q.k MMA
    const __half* rowp[2];
    uint4 kf[2][4];
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      rowp[tile] = s_ptr[g0 + tile * 8 + g];
#pragma unroll
      for (int i = 0; i < 4; ++i) kf[tile][i] = ldg128(rowp[tile] + 32 * i + 8 * t);
    }

    // ---- relative-pose embedding in B-fragment layout: eB[chunk][tile] = {slots 2t,2t+1 | slots 2t+8,2t+9} of
    // neighbour g0 + 8 tile + g. chunks 0-1: x, 2-3: y, 4-7: yaw; slots 0-7 cos, 8-15 sin of the same 8 angles
    uint32_t eB[8][2][2];
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      const float* rp = s_rel[g0 + tile * 8 + g];
      const float x = rp[0], y = rp[1], w = rp[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float s0, c0, s1, c1;
        __sincosf(x * fq[c][0], &s0, &c0);
        __sincosf(x * fq[c][1], &s1, &c1);
        eB[c][tile][0] = pack_h2(c0, c1);
        eB[c][tile][1] = pack_h2(s0, s1);
        __sincosf(y * fq[c][0], &s0, &c0);
        __sincosf(y * fq[c][1], &s1, &c1);
        eB[2 + c][tile][0] = pack_h2(c0, c1);
        eB[2 + c][tile][1] = pack_h2(s0, s1);
      }
      // yaw harmonics 8 cc + 2t + 1 (+1): bases by SFU, steps of 8 by plane rotation (pose_emb.py:52, integer freqs)
      float cA, sA, c1, s1, c8, s8;
      __sincosf(w * (float)(2 * t + 1), &sA, &cA);
      __sincosf(w, &s1, &c1);
      __sincosf(w * 8.f, &s8, &c8);
      float cB = fmaf(cA, c1, -sA * s1), sB = fmaf(sA, c1, cA * s1);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        eB[4 + cc][tile][0] = pack_h2(cA, cB);
        eB[4 + cc][tile][1] = pack_h2(sA, sB);
        if (cc < 3) {
          const float nA = fmaf(cA, c8, -sA * s8), nB = fmaf(cB, c8, -sB * s8);
          sA = fmaf(sA, c8, cA * s8);
          sB = fmaf(sB, c8, cB * s8);
          cA = nA;
          cB = nB;
        }
      }
    }

    // ---- logits^T[row = head (+4: residual of q / u)][col = neighbour]
    float sc[2][4];
#pragma unroll
    for (int tile = 0; tile < 2; ++tile)
#pragma unroll
      for (int r = 0; r < 4; ++r) sc[tile][r] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // q.k: chunks 2i, 2i+1 carry head i only
      const bool mine = hA == i;
      const uint32_t a00 = mine ? qA[0][0] : 0u, a01 = mine ? qA[0][1] : 0u;
      const uint32_t a10 = mine ? qA[1][0] : 0u, a11 = mine ? qA[1][1] : 0u;
#pragma unroll
      for (int tile = 0; tile < 2; ++tile) {
        mma16816(sc[tile], a00, 0u, a01, 0u, kf[tile][i].x, kf[tile][i].y);
        mma16816(sc[tile], a10, 0u, a11, 0u, kf[tile][i].z, kf[tile][i].w);
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int tile = 0; tile < 2; ++tile) mma16816(sc[tile], uA[c][0], 0u, uA[c][1], 0u, eB[c][tile][0], eB[c][tile][1]);

    // ---- V fragments (same addressing as K, second half of the row): in flight across the softmax
    uint4 vf[2][4];
#pragma unroll
    for (int tile = 0; tile < 2; ++tile)
#pragma unroll
      for (int i = 0; i < 4; ++i) vf[tile][i] = ldg128(rowp[tile] + D + 32 * i + 8 * t);

    // ---- softmax for head hA over this lane's 4 columns {2t, 2t+1, 8+2t, 9+2t}; lanes g and g^4 run in lockstep
    float lg[2][2];
#pragma unroll
    for (int tile = 0; tile < 2; ++tile)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v = sc[tile][j];
        v += __shfl_xor_sync(TB_FULL_MASK, v, 16);
        lg[tile][j] = (tile * 8 + 2 * t + j < nb) ? v : -INFINITY;
      }
    float gm = fmaxf(fmaxf(lg[0][0], lg[0][1]), fmaxf(lg[1][0], lg[1][1]));
    gm = fmaxf(gm, __shfl_xor_sync(TB_FULL_MASK, gm, 1));
    gm = fmaxf(gm, __shfl_xor_sync(TB_FULL_MASK, gm, 2));
    const float mn = fmaxf(mx, gm);  // finite: slot g0 is a valid neighbour
    float p[2][2], ps = 0.f;
#pragma unroll
    for (int tile = 0; tile < 2; ++tile)
#pragma unroll
      for (int j = 0; j < 2; ++j) { p[tile][j] = ex2(lg[tile][j] - mn); ps += p[tile][j]; }
    ps += __shfl_xor_sync(TB_FULL_MASK, ps, 1);
    ps += __shfl_xor_sync(TB_FULL_MASK, ps, 2);
    if (__any_sync(TB_FULL_MASK, mn > mx)) {  // lazy rescale (warp-uniform)
      const float corr = ex2(mx - mn);        // 1 where the running max did not move, 0 on the first group
      mx = mn;
      sm *= corr;
      // accumulator columns 2t, 2t+1 belong to heads (2t)&3, (2t+1)&3; their state lives in lanes with g == head
      const float ca = __shfl_sync(TB_FULL_MASK, corr, ((2 * t) & 3) << 2);
      const float cb = __shfl_sync(TB_FULL_MASK, corr, ((2 * t + 1) & 3) << 2);
#pragma unroll
      for (int c = 0; c < 8; ++c) { zacc[c][0] *= ca; zacc[c][1] *= cb; zacc[c][2] *= ca; zacc[c][3] *= cb; }
#pragma unroll
      for (int m = 0; m < 2; ++m) { oacc[m][0] *= ca; oacc[m][1] *= cb; oacc[m][2] *= ca; oacc[m][3] *= cb; }
    }
    sm += ps;
    const uint32_t pB0 = split_h2(p[0][0], p[0][1], lo_part), pB1 = split_h2(p[1][0], p[1][1], lo_part);

    // ---- z^T[row = channel slot][col = head (+4: residual of p)] += e^T p^T; e^T fragments by register transpose
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t a0 = movm_trans(eB[c][0][0]), a1 = movm_trans(eB[c][0][1]);
      const uint32_t a2 = movm_trans(eB[c][1][0]), a3 = movm_trans(eB[c][1][1]);
      mma16816(zacc[c], a0, a1, a2, a3, pB0, pB1);
    }

    // ---- ov^T[row = channel within head][col = head (+4)] += sum over heads i of v_i^T (p^T masked to head i).
    // Register r of piece i holds channels 32 i + 8 t + 2 r + {0,1}; after the transpose row g' of the tile is
    // channel 32 i + 8 (g'>>1) + 2 r + (g'&1): m-tile m stacks r = 2m (rows 0-7) and r = 2m+1 (rows 8-15).
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool mine = hA == i;
      const uint32_t b0 = mine ? pB0 : 0u, b1 = mine ? pB1 : 0u;
      {
        const uint32_t a0 = movm_trans(vf[0][i].x), a1 = movm_trans(vf[0][i].y);
        const uint32_t a2 = movm_trans(vf[1][i].x), a3 = movm_trans(vf[1][i].y);
        mma16816(oacc[0], a0, a1, a2, a3, b0, b1);
      }
      {
        const uint32_t a0 = movm_trans(vf[0][i].z), a1 = movm_trans(vf[0][i].w);
        const uint32_t a2 = movm_trans(vf[1][i].z), a3 = movm_trans(vf[1][i].w);
        mma16816(oacc[1], a0, a1, a2, a3, b0, b1);
      }
    }
  }

  // ---- normalise, un-permute through shared memory, store coalesced (all-masked row: zeros,
  // attention_rpe.py:188-190)
  const float inv = sm > 0.f ? 1.f / sm : 0.f;  // head hA
  const float ia = __shfl_sync(TB_FULL_MASK, inv, ((2 * t) & 3) << 2);
  const float ib = __shfl_sync(TB_FULL_MASK, inv, ((2 * t + 1) & 3) << 2);
  __syncwarp();  // every lane is done with s_ptr / s_rel
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    float v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) v[r] = oacc[m][r] + __shfl_xor_sync(TB_FULL_MASK, oacc[m][r], 2);  // + residual cols
    if (t < 2) {  // columns 2t, 2t+1 = heads; rows g (register 2m of the piece) and g+8 (register 2m+1)
      const int cp = 8 * (g >> 1) + 4 * m + (g & 1);
      s_out[32 * (2 * t) + cp] = v[0] * ia;
      s_out[32 * (2 * t + 1) + cp] = v[1] * ib;
      s_out[32 * (2 * t) + cp + 2] = v[2] * ia;
      s_out[32 * (2 * t + 1) + cp + 2] = v[3] * ib;
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) v[r] = zacc[c][r] + __shfl_xor_sync(TB_FULL_MASK, zacc[c][r], 2);
    if (t < 2) {  // rows g (cos slot) and g+8 (sin slot)
      float* zs = s_out + D + g;
      zs[(2 * t) * D + cos_base(c)] = v[0] * ia;
      zs[(2 * t + 1) * D + cos_base(c)] = v[1] * ib;
      zs[(2 * t) * D + sin_base(c)] = v[2] * ia;
      zs[(2 * t + 1) * D + sin_base(c)] = v[3] * ib;
    }
  }
  __syncwarp();
  *reinterpret_cast<float4*>(out_ov + (size_t)tok * ldo + lane * 4) = *reinterpret_cast<const float4*>(s_out + lane * 4);
  float* zp = out_z + (size_t)tok * ldo + lane * 4;
#pragma unroll
  for (int k = 0; k < H; ++k)
    *reinterpret_cast<float4*>(zp + k * D) = *reinterpret_cast<const float4*>(s_out + D + k * D + lane * 4);
  if (lane == 0 && out_none_valid) out_none_valid[tok] = nvalid > 0 ? 0 : 1;
}

}  // namespace

// Called by tb_knarpe_attn (knarpe_attn.cu) after argument validation when flags bit 1 is set. kv tables are fp16
// with leading dimensions in halves.
int tb_knarpe_attn_mma_launch(const float* q, int ldq, const float* u, int ldu, const void* kv0, int ldkv0, int T0,
                              int div0, int K0, const void* kv1, int ldkv1, int T1, int div1, int K1,
                              const int32_t* idx, const uint8_t* invalid, const float* rel, const float* pe_freq_xy,
                              int B, int S, float* out_ov, float* out_z, int ldo, uint8_t* out_none_valid,
                              cudaStream_t st) {
  const int n_tok = B * S;
  const int grid = (n_tok + kWarps - 1) / kWarps;
  knarpe_attn_mma_kernel<<<grid, kWarps * 32, 0, st>>>(
      q, ldq, u, ldu, static_cast<const __half*>(kv0), ldkv0, T0, div0, K0, static_cast<const __half*>(kv1), ldkv1, T1,
      div1, K1, idx, invalid, rel, pe_freq_xy, n_tok, S, out_ov, out_z, ldo, out_none_valid);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

bool tb_knarpe_attn_mma_supported(int D_, int Hh, int Ktot) { return D_ == D && Hh == H && ((Ktot + 15) & ~15) <= KMAX; }
