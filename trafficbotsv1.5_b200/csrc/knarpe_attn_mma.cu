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
// registers - nothing is staged through shared memory. Two tokens share a warp (knarpe_attn_mma_pair_kernel below):
//   * a group = 8 compacted neighbours of token A (MMA rows 0-7) + 8 of token B (rows 8-15). Lane (g = lane>>2,
//     t = lane&3) owns row g of both: it loads 32-byte pieces of their fp16 K and V rows that are exactly its
//     A-fragment registers, and evaluates the cos/sin of 16 embedding angles per neighbour - the A-fragment elements
//     (m = neighbour, k = channel slot) of the e.u MMA. The yaw harmonics of a lane form arithmetic progressions:
//     3 SFU sincos + packed plane rotations replace 16; the geometric x/y frequencies are evaluated directly.
//   * logits [16 x 8] = A(k) B(q, block-diagonal over heads) + A(e) B(u): B columns 0-3 hold the fp16 q / u of token
//     A's heads, columns 4-7 of token B's (q / u arrive as fp16 rows from the in-projection's epilogue).
//   * z^T [128 x 8] = A(e^T) B(p^T) and ov^T [32 x 8] = sum_heads A(v_head^T) B(p^T masked to that head); the
//     transposed fragments (e^T, v^T, p^T) come from movmatrix (register-only 8x8 transposes).
// (Round 1 also had a one-token-per-warp kernel with split-fp16 q / u for fp32 callers; it was never on the engine's
// path and was removed in round 2: fp32 q / u rows on fp16 tables run on the SIMT kernel of knarpe_attn.cu.)
// Staging the gathered rows in shared memory instead (cp.async + ldmatrix: 835 us, cp.async.bulk per row + ldmatrix:
// 634 us on the agent cross-attention launch) lost against these direct fragment loads (493 us): the per-lane
// cp.async clogs the LSU queue, and bulk copies need uniform-register operands, i.e. a serialised 16-trip loop per
// tile (profiles/r1_notes.md).
// Channel "slots" inside a 16-wide MMA chunk are permutations of the reference orders (embedding:
// utils/pose_emb.py:50-55 [cos x|sin x|cos y|sin y|cos yaw|sin yaw]); q/u are read and ov/z written through the same
// permutations, so the C ABI layouts are unchanged.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

// One warp per CTA, 12 CTAs per SM (168 registers): warps are independent, so the smallest CTA hands its register
// slot back as soon as its own token pair is done instead of waiting for the slowest of 4 (measured on the agent
// cross-attention launch: 393 us with 4 warps per CTA, 389 with 2, 380 with 1).
#ifndef TB_MMA_WARPS
#define TB_MMA_WARPS 1
#endif
#ifndef TB_MMA_MINB
#define TB_MMA_MINB 12
#endif
#ifndef TB_MMA_PREFETCH
#define TB_MMA_PREFETCH 1
#endif
constexpr int kWarps = TB_MMA_WARPS;
constexpr int D = 128;
constexpr int H = 4;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
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

// ---------------------------------------------------------------------------------------------------------------
// Two tokens per warp for SHORT neighbour lists (agent self-attention, K = 25): MMA rows 0-7 are 8 neighbours of token
// A, rows 8-15 are 8 neighbours of token B; B-operand columns 0-3 carry A's heads, columns 4-7 carry B's (fp16 q / u:
// no residual columns needed). Every MMA then serves both tokens: logits rows 0-7 x cols 0-3 and rows 8-15 x cols 4-7
// are the useful blocks, p^T is block structured (zero cross blocks), so z^T / ov^T columns 0-3 accumulate A's and
// columns 4-7 B's outputs. Lanes t < 2 own token A (heads 2t, 2t+1), lanes t >= 2 token B (heads 2(t-2), +1).
// Versus one token per warp: 8-neighbour granularity (24 instead of 32 slots for 21 valid neighbours) and the
// per-token set-up / epilogue shared by two tokens. Needs fp16 q / u rows and K0 + K1 <= 128.
constexpr int KPAIR = 128;  // compacted neighbour slots per token
// per warp: the two neighbour lists (row pointer + float4 relative pose per slot), reused as the 2 x 640-float output
// staging of the epilogue
constexpr int kPairSmem = 2 * (KPAIR * 8 + KPAIR * 16);
// epilogue staging of one token: [ov heads 0,1 | 4 pad | ov heads 2,3 | z head 0..3], z head rows kZHead floats apart
// and token blocks kOutTok apart so that the four lanes of a row group (t = 0..3: heads 0 / 2 of token A / B) hit banks
// 0 / 8 / 16 / 24 + g instead of one bank (the unpadded layout was a 4-way conflict on every staging store: 120 extra
// wavefronts per pair)
constexpr int kZHead = D + 4, kZBase = D + 4, kOutTok = kZBase + H * kZHead + 28;
static_assert(kOutTok % 32 == 16 && (2 * kZHead) % 32 == 8 && kOutTok % 4 == 0, "bank spread / float4 alignment");
static_assert(kPairSmem >= 2 * kOutTok * 4, "output staging must fit");
// prologue staging of the two u rows (bytes): head rows 272 apart, token blocks 4 * 272 apart
constexpr int kUHead = 2 * D + 16, kUTok = H * kUHead;
static_assert(2 * kUTok <= 2 * KPAIR * 16, "u staging must fit in the s_rel area");

// One 128-channel fp16 row into 16 registers per lane. Plain layout: four 128-bit pieces, piece i = head i, lane t
// takes halves [32 i + 8 t, +8). Head-interleaved layout (IL): two 256-bit pieces, lane t takes halves
// [64 i + 16 t, +16) whose first 8 belong to head 2 i and last 8 to head 2 i + 1 - half as many L1 wavefronts per row.
// Either way registers 4 h .. 4 h + 3 hold head h, and the position inside the head is 8 t + (0..7).
template <bool IL>
__device__ __forceinline__ void load_row(uint32_t (&r)[16], const __half* row, int t) {
  if (IL) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[8 * i]), "=r"(r[8 * i + 1]), "=r"(r[8 * i + 2]), "=r"(r[8 * i + 3]), "=r"(r[8 * i + 4]),
                     "=r"(r[8 * i + 5]), "=r"(r[8 * i + 6]), "=r"(r[8 * i + 7])
                   : "l"(row + 64 * i + 16 * t));
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 v = ldg128(row + 32 * i + 8 * t);
      r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
  }
}

template <bool OUT_H, bool IL>
__global__ void __launch_bounds__(kWarps * 32, TB_MMA_MINB)
knarpe_attn_mma_pair_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ u, int ldu,
                            const __half* __restrict__ kv0, int ldkv0, int T0, int div0, int K0,
                            const __half* __restrict__ kv1, int ldkv1, int T1, int div1, int K1,
                            const int32_t* __restrict__ idx, const uint8_t* __restrict__ invalid,
                            const float* __restrict__ rel, const float* __restrict__ pe_freq_xy, int n_tok, int S,
                            void* __restrict__ out_ov_, void* __restrict__ out_z_, int ldo,
                            uint8_t* __restrict__ out_none_valid) {
  __shared__ __align__(16) unsigned char s_raw[kWarps][kPairSmem];
  const int Ktot = K0 + K1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok0 = (blockIdx.x * kWarps + warp) * 2;
  if (tok0 >= n_tok) return;  // warp-uniform
  const bool has_b = tok0 + 1 < n_tok;
  const __half** s_ptr = reinterpret_cast<const __half**>(s_raw[warp]);               // [2][KPAIR]
  float4* s_rel = reinterpret_cast<float4*>(s_raw[warp] + 2 * KPAIR * 8);              // [2][KPAIR] (x, y, yaw, -)
  float* s_out = reinterpret_cast<float*>(s_raw[warp]);                                // epilogue: 2 x [ov | z]

  const int g = lane >> 2, t = lane & 3;
  const int hA = g & 3;          // operand column g: head hA of token (g >> 2)
  const int my = t >> 1;         // token whose softmax state / accumulator columns this lane owns
  const int h0 = 2 * (t & 1);    // its heads h0, h0 + 1
  const unsigned lt_mask = (1u << lane) - 1u;

  // ---- neighbour lists of both tokens: lanes 0-31 cover K0 <= 64 candidates of a token in two chunks
  int nvalid[2] = {0, 0};
  float fq[2][2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    fq[c][0] = __ldg(pe_freq_xy + 8 * c + 2 * t);
    fq[c][1] = __ldg(pe_freq_xy + 8 * c + 2 * t + 1);
  }
  uint8_t n_inv[2][4];
  int n_id[2][4];
  float n_rel[2][4][3];
#pragma unroll
  for (int tk = 0; tk < 2; ++tk)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = c * 32 + lane;
      n_inv[tk][c] = 1; n_id[tk][c] = 0; n_rel[tk][c][0] = n_rel[tk][c][1] = n_rel[tk][c][2] = 0.f;
      if ((tk == 0 || has_b) && c * 32 < Ktot && j < Ktot) {
        const size_t p = (size_t)(tok0 + tk) * Ktot + j;
        n_inv[tk][c] = __ldg(invalid + p);
        n_id[tk][c] = __ldg(idx + p);
        n_rel[tk][c][0] = __ldg(rel + p * 3 + 0);
        n_rel[tk][c][1] = __ldg(rel + p * 3 + 1);
        n_rel[tk][c][2] = __ldg(rel + p * 3 + 2);
      }
    }
  // B fragments: column g takes head hA of token g >> 2 (token B absent: zeros). The u fragments are 4-byte pieces of
  // 8 different (token, head) rows per load - 8 L1 tag look-ups for 128 useful bytes (ncu: "L1 Tag Requests" 8.0 on
  // each of the 16 loads). Instead the two u rows are read coalesced (4 x 128-bit loads per lane), staged in the still
  // unused list area with the head rows 272 bytes apart (bank = 16 token + 4 head + t: conflict-free) and picked up
  // with 16 one-wavefront LDS: 16 tag look-ups per pair instead of 128 (time-neutral on its own, profiles/r2_notes.md 10).
  uint32_t uB[8][2], qB[2][2];
  {
    const int tkc = g >> 2;
    const bool live = tkc == 0 || has_b;
    unsigned char* s_u = s_raw[warp] + 2 * KPAIR * 8;  // (the s_rel area: the lists are built after this block)
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // piece i: token i >> 1, bytes [512 (i & 1) + 16 lane, +16) of its u row
      const int tk = i >> 1, piece = (i & 1) * 32 + lane;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (tk == 0 || has_b) v = ldg128(u + (size_t)(tok0 + tk) * ldu + piece * 8);
      *reinterpret_cast<uint4*>(s_u + tk * kUTok + (piece >> 4) * kUHead + (piece & 15) * 16) = v;
    }
    const __half* qp = q + (size_t)(tok0 + (live ? tkc : 0)) * ldq +
                       (IL ? 64 * (hA >> 1) + 16 * t + 8 * (hA & 1) : 32 * hA + 8 * t);
    const uint4 qq = live ? __ldg(reinterpret_cast<const uint4*>(qp)) : make_uint4(0u, 0u, 0u, 0u);
    qB[0][0] = qq.x; qB[0][1] = qq.y; qB[1][0] = qq.z; qB[1][1] = qq.w;
    __syncwarp();
    const unsigned char* up = s_u + tkc * kUTok + hA * kUHead + 4 * t;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uB[c][0] = *reinterpret_cast<const uint32_t*>(up + 2 * cos_base(c));
      uB[c][1] = *reinterpret_cast<const uint32_t*>(up + 2 * sin_base(c));
    }
    __syncwarp();  // the list build below overwrites the staging
  }
  // table bases: the two tokens are consecutive, so the scene index needs one division per warp and the second
  // token almost always shares the first one's tables (the divisions were ~100 instructions per pair)
  const __half* kb[2];
  const __half* kbx[2];
  {
    const unsigned b0 = (unsigned)tok0 / (unsigned)S;
    kb[0] = kv0 + (size_t)(b0 / (unsigned)div0) * T0 * ldkv0;
    kbx[0] = (K1 > 0) ? kv1 + (size_t)(b0 / (unsigned)div1) * T1 * ldkv1 : kb[0];
    kb[1] = kb[0]; kbx[1] = kbx[0];
    if (has_b && (unsigned)(tok0 + 1) - b0 * (unsigned)S >= (unsigned)S) {  // token B opens the next scene
      kb[1] = kv0 + (size_t)((b0 + 1) / (unsigned)div0) * T0 * ldkv0;
      kbx[1] = (K1 > 0) ? kv1 + (size_t)((b0 + 1) / (unsigned)div1) * T1 * ldkv1 : kb[1];
    }
  }
#pragma unroll
  for (int tk = 0; tk < 2; ++tk) {
    const __half* kb1 = kbx[tk];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c * 32 < Ktot) {  // warp-uniform
        const bool valid = n_inv[tk][c] == 0;
        const unsigned vb = __ballot_sync(TB_FULL_MASK, valid);
        if (valid) {
          const int pos = tk * KPAIR + nvalid[tk] + __popc(vb & lt_mask);
          s_ptr[pos] = (c * 32 + lane < K0) ? kb[tk] + (size_t)n_id[tk][c] * ldkv0
                                             : kb1 + (size_t)n_id[tk][c] * ldkv1;
          s_rel[pos] = make_float4(n_rel[tk][c][0], n_rel[tk][c][1], n_rel[tk][c][2], 0.f);
        }
        nvalid[tk] += __popc(vb);
      }
    }
    const int npad_t = (nvalid[tk] + 7) & ~7;
    if (lane < npad_t - nvalid[tk]) {  // weight-0 dummies up to a multiple of 8
      s_ptr[tk * KPAIR + nvalid[tk] + lane] = kb[tk];
      s_rel[tk * KPAIR + nvalid[tk] + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const int npad = max((nvalid[0] + 7) & ~7, (nvalid[1] + 7) & ~7);
  // a token whose list is shorter than the other's keeps reading (masked) dummies: fill the rest of its list
  for (int tk = 0; tk < 2; ++tk)
    for (int j = ((nvalid[tk] + 7) & ~7) + lane; j < npad; j += 32) {
      s_ptr[tk * KPAIR + j] = kb[tk];
      s_rel[tk * KPAIR + j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  __syncwarp();

  float zacc[8][4], oacc[2][4];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) zacc[c][r] = 0.f;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int r = 0; r < 4; ++r) oacc[m][r] = 0.f;
  float mx[2] = {-INFINITY, -INFINITY}, sm[2] = {0.f, 0.f};  // heads h0, h0+1 of token `my`
  const int my_nvalid = my ? nvalid[1] : nvalid[0];

  for (int g0 = 0; g0 < npad; g0 += 8) {
    const __half* rowp[2];
    uint32_t kf[2][16];  // IL: two 256-bit pieces per row; else four 128-bit pieces (one per head)
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      rowp[tile] = s_ptr[tile * KPAIR + g0 + g];
      load_row<IL>(kf[tile], rowp[tile], t);
    }
#if TB_MMA_PREFETCH
    // ptxas sinks the V loads below the logits MMAs (no registers to hold them earlier), so their L2 round trip used to
    // be exposed: the first V movmatrix carried 12 % of all stall samples, the K consumers another 11 %. L1 prefetches
    // need no registers: lanes t < 2 pull the V half (lines 2, 3 of the 512-byte row) of THIS group's rows now, a whole
    // trig + logits phase ahead of the loads; lanes t >= 2 pull the K half (lines 0, 1) of the NEXT group's rows.
    {
      const bool nxt = t >= 2;
      if (!nxt || g0 + 8 < npad) {
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const __half* r = nxt ? s_ptr[tile * KPAIR + g0 + 8 + g] + 64 * (t - 2) : rowp[tile] + D + 64 * t;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(r));
        }
      }
    }
#endif
    uint32_t eA[8][4];
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      const float4 rp = s_rel[tile * KPAIR + g0 + g];
      const float x = rp.x, y = rp.y, w = rp.z;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float s0, c0, s1, c1;
        __sincosf(x * fq[c][0], &s0, &c0);
        __sincosf(x * fq[c][1], &s1, &c1);
        eA[c][tile] = pack_h2(c0, c1);
        eA[c][2 + tile] = pack_h2(s0, s1);
        __sincosf(y * fq[c][0], &s0, &c0);
        __sincosf(y * fq[c][1], &s1, &c1);
        eA[2 + c][tile] = pack_h2(c0, c1);
        eA[2 + c][2 + tile] = pack_h2(s0, s1);
      }
      float cA, sA, c1, s1, c8, s8;
      __sincosf(w * (float)(2 * t + 1), &sA, &cA);
      __sincosf(w, &s1, &c1);
      __sincosf(w * 8.f, &s8, &c8);
      float2 cv = make_float2(cA, fmaf(cA, c1, -sA * s1)), sv = make_float2(sA, fmaf(sA, c1, cA * s1));
      const float2 c8v = make_float2(c8, c8), s8v = make_float2(s8, s8), ns8v = make_float2(-s8, -s8);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        eA[4 + cc][tile] = pack_h2(cv.x, cv.y);
        eA[4 + cc][2 + tile] = pack_h2(sv.x, sv.y);
        if (cc < 3) {
          const float2 nc = __ffma2_rn(cv, c8v, __fmul2_rn(sv, ns8v));
          sv = __ffma2_rn(sv, c8v, __fmul2_rn(cv, s8v));
          cv = nc;
        }
      }
    }
    // ---- logits: rows 0-7 x cols 0-3 (token A) and rows 8-15 x cols 4-7 (token B) are the useful blocks
    float sc[4] = {0.f, 0.f, 0.f, 0.f}, sc1[4] = {0.f, 0.f, 0.f, 0.f}, sc2[4] = {0.f, 0.f, 0.f, 0.f},
          sc3[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 8; ++c) mma16816((c & 1) ? sc1 : sc, eA[c][0], eA[c][1], eA[c][2], eA[c][3], uB[c][0], uB[c][1]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // either layout: registers 4i..4i+3 of a row belong to head i
      const bool mine = hA == i;
      mma16816(sc2, kf[0][4 * i], kf[1][4 * i], kf[0][4 * i + 1], kf[1][4 * i + 1], mine ? qB[0][0] : 0u,
               mine ? qB[0][1] : 0u);
      mma16816(sc3, kf[0][4 * i + 2], kf[1][4 * i + 2], kf[0][4 * i + 3], kf[1][4 * i + 3], mine ? qB[1][0] : 0u,
               mine ? qB[1][1] : 0u);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) sc[r] = (sc[r] + sc1[r]) + (sc2[r] + sc3[r]);

    uint32_t vf[2][16];
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) load_row<IL>(vf[tile], rowp[tile] + D, t);

    // ---- softmax of this lane's token (row block `my`), heads h0 / h0+1, over the 8 rows (lanes with the same t).
    // The reference maximum is lazy: it is re-established (3 shuffles per head + accumulator rescale) only for the
    // first group and when some logit exceeds it by more than 2^kSlack — otherwise p = 2^(logit - mx) <= 2^kSlack
    // is used as it is (exact in real arithmetic; fp32 accumulators, fp16 p keeps its 11 bits). The denominator is a
    // per-lane partial sum reduced once in the epilogue. Saves 12 of the ~200 LSU wavefronts of a group.
    constexpr float kSlack = 8.f;
    const bool row_ok = g0 + g < my_nvalid;
    float lg[2], p[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) lg[j] = row_ok ? (my ? sc[2 + j] : sc[j]) : -INFINITY;
    if (g0 == 0 || __any_sync(TB_FULL_MASK, lg[0] > mx[0] + kSlack || lg[1] > mx[1] + kSlack)) {
      float mn[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float gm = lg[j];
        gm = fmaxf(gm, __shfl_xor_sync(TB_FULL_MASK, gm, 4));
        gm = fmaxf(gm, __shfl_xor_sync(TB_FULL_MASK, gm, 8));
        gm = fmaxf(gm, __shfl_xor_sync(TB_FULL_MASK, gm, 16));
        mn[j] = fmaxf(mx[j], gm);
      }
      if (g0 != 0) {  // (first group: the accumulators are still zero, nothing to rescale)
        const float ca = mn[0] > mx[0] ? ex2(mx[0] - mn[0]) : 1.f, cb = mn[1] > mx[1] ? ex2(mx[1] - mn[1]) : 1.f;
        sm[0] *= ca; sm[1] *= cb;
#pragma unroll
        for (int c = 0; c < 8; ++c) { zacc[c][0] *= ca; zacc[c][1] *= cb; zacc[c][2] *= ca; zacc[c][3] *= cb; }
#pragma unroll
        for (int m = 0; m < 2; ++m) { oacc[m][0] *= ca; oacc[m][1] *= cb; oacc[m][2] *= ca; oacc[m][3] *= cb; }
      }
      mx[0] = mn[0]; mx[1] = mn[1];
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      p[j] = row_ok ? ex2(lg[j] - mx[j]) : 0.f;  // mx is finite whenever row_ok
      sm[j] += p[j];                              // this lane's rows only
    }
    // p^T B fragment: tile 0 = rows 0-7 (token A's neighbours) has values only in columns 0-3 (lanes t < 2), tile 1
    // only in columns 4-7 (lanes t >= 2)
    const uint32_t pk = pack_h2(p[0], p[1]);
    const uint32_t pB0 = movm_trans(my == 0 ? pk : 0u), pB1 = movm_trans(my == 1 ? pk : 0u);

#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t a0 = movm_trans(eA[c][0]), a1 = movm_trans(eA[c][2]);
      const uint32_t a2 = movm_trans(eA[c][1]), a3 = movm_trans(eA[c][3]);
      mma16816(zacc[c], a0, a1, a2, a3, pB0, pB1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool mine = hA == i;
      const uint32_t b0 = mine ? pB0 : 0u, b1 = mine ? pB1 : 0u;
      {
        const uint32_t a0 = movm_trans(vf[0][4 * i]), a1 = movm_trans(vf[0][4 * i + 1]);
        const uint32_t a2 = movm_trans(vf[1][4 * i]), a3 = movm_trans(vf[1][4 * i + 1]);
        mma16816(oacc[0], a0, a1, a2, a3, b0, b1);
      }
      {
        const uint32_t a0 = movm_trans(vf[0][4 * i + 2]), a1 = movm_trans(vf[0][4 * i + 3]);
        const uint32_t a2 = movm_trans(vf[1][4 * i + 2]), a3 = movm_trans(vf[1][4 * i + 3]);
        mma16816(oacc[1], a0, a1, a2, a3, b0, b1);
      }
    }
  }

  // ---- epilogue: lanes t < 2 hold token A's columns, t >= 2 token B's; staging block per token
#pragma unroll
  for (int j = 0; j < 2; ++j) {  // denominators: the 8 lanes with the same t hold the partial sums of a column pair
    sm[j] += __shfl_xor_sync(TB_FULL_MASK, sm[j], 4);
    sm[j] += __shfl_xor_sync(TB_FULL_MASK, sm[j], 8);
    sm[j] += __shfl_xor_sync(TB_FULL_MASK, sm[j], 16);
  }
  const float ia = sm[0] > 0.f ? 1.f / sm[0] : 0.f, ib = sm[1] > 0.f ? 1.f / sm[1] : 0.f;
  __syncwarp();
  {
    float* so = s_out + my * kOutTok;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const int cp = 8 * (g >> 1) + 4 * m + (g & 1);
      so[32 * h0 + 2 * h0 + cp] = oacc[m][0] * ia;  // (+ 2 h0: the 4-float pad in front of heads 2, 3)
      so[32 * (h0 + 1) + 2 * h0 + cp] = oacc[m][1] * ib;
      so[32 * h0 + 2 * h0 + cp + 2] = oacc[m][2] * ia;
      so[32 * (h0 + 1) + 2 * h0 + cp + 2] = oacc[m][3] * ib;
    }
    float* zs = so + kZBase + g;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      zs[h0 * kZHead + cos_base(c)] = zacc[c][0] * ia;
      zs[(h0 + 1) * kZHead + cos_base(c)] = zacc[c][1] * ib;
      zs[h0 * kZHead + sin_base(c)] = zacc[c][2] * ia;
      zs[(h0 + 1) * kZHead + sin_base(c)] = zacc[c][3] * ib;
    }
  }
  __syncwarp();
#pragma unroll
  for (int tk = 0; tk < 2; ++tk) {
    if (tk == 1 && !has_b) break;  // warp-uniform
    const int tok = tok0 + tk;
    const float* so = s_out + tk * kOutTok;
    if (OUT_H) {
      auto to_h4 = [](const float* pp) {
        const float4 v = *reinterpret_cast<const float4*>(pp);
        return make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
      };
      __half* ov_h = static_cast<__half*>(out_ov_) + (size_t)tok * ldo + lane * 4;
      __half* z_h = static_cast<__half*>(out_z_) + (size_t)tok * ldo + lane * 4;
      *reinterpret_cast<uint2*>(ov_h) = to_h4(so + lane * 4 + 4 * (lane >> 4));
#pragma unroll
      for (int k = 0; k < H; ++k) *reinterpret_cast<uint2*>(z_h + k * D) = to_h4(so + kZBase + k * kZHead + lane * 4);
    } else {
      float* ovp = static_cast<float*>(out_ov_) + (size_t)tok * ldo + lane * 4;
      float* zp = static_cast<float*>(out_z_) + (size_t)tok * ldo + lane * 4;
      *reinterpret_cast<float4*>(ovp) = *reinterpret_cast<const float4*>(so + lane * 4 + 4 * (lane >> 4));
#pragma unroll
      for (int k = 0; k < H; ++k)
        *reinterpret_cast<float4*>(zp + k * D) = *reinterpret_cast<const float4*>(so + kZBase + k * kZHead + lane * 4);
    }
    if (lane == 0 && out_none_valid) out_none_valid[tok] = nvalid[tk] > 0 ? 0 : 1;
  }
}

}  // namespace

// Called by tb_knarpe_attn (knarpe_attn.cu) after argument validation when flags bit 1 is set. kv tables are fp16
// with leading dimensions in halves.
int tb_knarpe_attn_mma_launch(const void* q, int ldq, const void* u, int ldu, int in_f16, const void* kv0, int ldkv0, int T0,
                              int div0, int K0, const void* kv1, int ldkv1, int T1, int div1, int K1,
                              const int32_t* idx, const uint8_t* invalid, const float* rel, const float* pe_freq_xy,
                              int B, int S, void* out_ov, void* out_z, int ldo, int out_f16, uint8_t* out_none_valid,
                              int interleaved, cudaStream_t st) {
  if (!in_f16) return TB_ERR_UNSUPPORTED;  // fp32 q / u rows go to the SIMT kernel (knarpe_attn.cu)
  const int n_tok = B * S;
  const int grid2 = (n_tok + 2 * kWarps - 1) / (2 * kWarps);  // two tokens per warp
#define TB_PAIR_LAUNCH(OH, IL)                                                                                         \
  knarpe_attn_mma_pair_kernel<OH, IL><<<grid2, kWarps * 32, 0, st>>>(                                                  \
      static_cast<const __half*>(q), ldq, static_cast<const __half*>(u), ldu, static_cast<const __half*>(kv0), ldkv0,  \
      T0, div0, K0, static_cast<const __half*>(kv1), ldkv1, T1, div1, K1, idx, invalid, rel, pe_freq_xy, n_tok, S,     \
      out_ov, out_z, ldo, out_none_valid)
  if (out_f16 && interleaved) TB_PAIR_LAUNCH(true, true);
  else if (out_f16) TB_PAIR_LAUNCH(true, false);
  else if (interleaved) TB_PAIR_LAUNCH(false, true);
  else TB_PAIR_LAUNCH(false, false);
#undef TB_PAIR_LAUNCH
  TB_CHECK_LAUNCH();
  return TB_OK;
}

bool tb_knarpe_attn_mma_supported(int D_, int Hh, int Ktot) { return D_ == D && Hh == H && Ktot <= KPAIR; }
