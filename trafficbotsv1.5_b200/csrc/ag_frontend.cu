// Fused agent history encoder of the tensor-core mode (tb_ag_frontend): everything AgentEncoder._forward_hptr does
// before the transformer layers (agent_encoder.py:130-162) in ONE kernel, one warp per agent:
//   last-valid token pose (pooling.py:24-29) -> history rows in the token frame -> [attr6 | motion3 | one-hot W]
//   -> input MLP 20 -> 64 -> 64 -> 64 (input_encoder.py:41-61, ReLU between) ++ PoseEmb64 (pose_emb.py:50-55)
//   -> PointNet 3 x (Linear 128 -> 64, ReLU, max over the valid steps, concat) -> max_valid pool
//   (polyline_encoder.py:50-53, pooling.py:18-19,38) -> token [max h | max h].
// The unfused path (tb_ag_featurize + 3 + 6 projections + 3 pooling launches) moves ~2.5 GB of activations per step
// through HBM for the 720,896 history rows of config 3 (1.0 ms = 17 % of a policy iteration, profiles/r1_notes.md);
// here the W <= 16 history rows of an agent are the 16 rows of an mma.sync.m16n8k16 tile, every activation stays in
// registers (accumulator fragments are re-packed as the next layer's A fragments), the fp16 weights sit in shared
// memory (rows padded by 16 B: conflict-free B-fragment loads), and only the 512-byte token leaves the SM.
// PointNet's cat([h, max]) W^T needs no weight split: the max half of the A operand is the lane's own column maxima
// replicated over the rows. Operands are fp16 (10-bit mantissa like the tf32 projections they replace), fp32 accumulate.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int AW = 8;                                   // warps (agents in flight) per CTA
constexpr int S1 = 32 + 8, S2 = 64 + 8, SP = 128 + 8;   // padded weight row strides (halves)
constexpr int OFF_W1 = 0, OFF_W2 = OFF_W1 + 64 * S1, OFF_W3 = OFF_W2 + 64 * S2, OFF_P0 = OFF_W3 + 64 * S2,
              OFF_P1 = OFF_P0 + 64 * SP, OFF_P2 = OFF_P1 + 64 * SP, W_HALVES = OFF_P2 + 64 * SP;
constexpr int N_BIAS = 6 * 64;
constexpr size_t SMEM = (size_t)W_HALVES * 2 + (N_BIAS + 8) * 4;

__device__ __forceinline__ uint32_t pack_h2(float a, float b) { return tb_pack_h2_sat(a, b); }
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// c[nt] (8 n-tiles of 8 outputs) = A (KC k-chunks of 16) x W^T + bias; W row-major [64][stride] fp16 in smem
template <int KC>
__device__ __forceinline__ void dense64(const uint32_t (&a)[KC][4], const __half* __restrict__ Ws, int stride,
                                        const float* __restrict__ bias, int g, int t, float (&c)[8][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const float2 bb = *reinterpret_cast<const float2*>(bias + nt * 8 + 2 * t);
    c[nt][0] = bb.x; c[nt][1] = bb.y; c[nt][2] = bb.x; c[nt][3] = bb.y;
    const __half* wr = Ws + (nt * 8 + g) * stride + 2 * t;
#pragma unroll
    for (int kc = 0; kc < KC; ++kc)
      mma16816(c[nt], a[kc], *reinterpret_cast<const uint32_t*>(wr + kc * 16),
               *reinterpret_cast<const uint32_t*>(wr + kc * 16 + 8));
  }
}
// accumulator fragments -> A fragments of the next layer (k-chunk kc = n-tiles 2kc, 2kc+1), optional ReLU
// (hm: running |max| of every converted activation, the fp16 range guard of common.cuh)
template <bool RELU>
__device__ __forceinline__ void c_to_a(const float (&c)[8][4], uint32_t (*a)[4], uint32_t& hm) {
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    auto r = [](float v) { return RELU ? fmaxf(v, 0.f) : v; };
    a[kc][0] = pack_h2(r(c[2 * kc][0]), r(c[2 * kc][1]));
    a[kc][1] = pack_h2(r(c[2 * kc][2]), r(c[2 * kc][3]));
    a[kc][2] = pack_h2(r(c[2 * kc + 1][0]), r(c[2 * kc + 1][1]));
    a[kc][3] = pack_h2(r(c[2 * kc + 1][2]), r(c[2 * kc + 1][3]));
#pragma unroll
    for (int j = 0; j < 4; ++j) tb_track_h2(hm, a[kc][j]);
  }
}
// ReLU in place + column maxima over the valid rows (rows g: regs 0,1; rows g+8: regs 2,3), replicated over g
__device__ __forceinline__ void relu_colmax(float (&c)[8][4], bool v_lo, bool v_hi, float (&m)[8][2]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int r = 0; r < 4; ++r) c[nt][r] = fmaxf(c[nt][r], 0.f);
    float a = fmaxf(v_lo ? c[nt][0] : -INFINITY, v_hi ? c[nt][2] : -INFINITY);
    float b = fmaxf(v_lo ? c[nt][1] : -INFINITY, v_hi ? c[nt][3] : -INFINITY);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      a = fmaxf(a, __shfl_xor_sync(TB_FULL_MASK, a, o));
      b = fmaxf(b, __shfl_xor_sync(TB_FULL_MASK, b, o));
    }
    m[nt][0] = a; m[nt][1] = b;
  }
}
// PoseEmb pe_dim = 64 component c (agent_encoder.py:50): [cos(x f0..7)|sin(x f)|cos(y f)|sin(y f)|cos(w 1..16)|sin(..)]
// with the SFU's own range reduction (tensor-core mode, cf. tb_knarpe_attn flags bit 0)
__device__ __forceinline__ float pe64(int c, float x, float y, float w, const float* __restrict__ f) {
  float a;
  bool is_sin;
  if (c < 16) { a = x * f[c & 7]; is_sin = c >= 8; }
  else if (c < 32) { a = y * f[c & 7]; is_sin = c >= 24; }
  else { a = w * (float)(((c - 32) & 15) + 1); is_sin = c >= 48; }
  return is_sin ? __sinf(a) : __cosf(a);
}

__global__ void __launch_bounds__(AW * 32, 2)
ag_frontend_kernel(const uint8_t* __restrict__ hist_valid, const float* __restrict__ hist_pose,
                   const float* __restrict__ hist_motion, const float* __restrict__ ag_attr,
                   const int* __restrict__ d_step, int step_stride, int A, const float* __restrict__ freq_xy,
                   int n_ag_tot, int W, const __half* __restrict__ wblob, const float* __restrict__ bias, float* __restrict__ tok_out,
                   int ldo, float* __restrict__ tok_pose, uint8_t* __restrict__ tok_invalid,
                   const float* __restrict__ ln_g, const float* __restrict__ ln_b, __half* __restrict__ ln_out,
                   int ld_ln, unsigned int* __restrict__ sat_flag) {
  extern __shared__ __align__(16) unsigned char smem[];
  __half* sW = reinterpret_cast<__half*>(smem);
  float* sB = reinterpret_cast<float*>(smem + (size_t)W_HALVES * 2);
  for (int i = threadIdx.x; i < W_HALVES / 8; i += blockDim.x)
    reinterpret_cast<uint4*>(sW)[i] = __ldg(reinterpret_cast<const uint4*>(wblob) + i);
  for (int i = threadIdx.x; i < N_BIAS; i += blockDim.x) sB[i] = __ldg(bias + i);
  if (threadIdx.x < 8) sB[N_BIAS + threadIdx.x] = __ldg(freq_xy + threadIdx.x);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int s_all = *d_step;
  float* fr = sB + N_BIAS;  // the 8 xy frequencies (shared: indexed by a lane-dependent component number)
  uint32_t hm = 0u;         // fp16 range guard: |max| of the activations this lane converted

  for (int ba = blockIdx.x * AW + warp; ba < n_ag_tot; ba += gridDim.x * AW) {  // warp-uniform
    const size_t hb = (size_t)ba * W;
    const int s = step_stride ? d_step[(size_t)(ba / A) * step_stride] : s_all;  // one loop counter per batch row
    const int n_step = min(s, W);
    // window position wp in [W-n_step, W) <-> time s - W + wp, ring slot (s - W + wp) % W; wp < W - n_step: absent
    int last_wp = -1;
    for (int wp = W - n_step; wp < W; ++wp)
      if (hist_valid[hb + ((s - W + wp) % W)]) last_wp = wp;
    float px = 0.f, py = 0.f, pw = 0.f;
    if (last_wp >= 0) {
      const int slot = (s - W + last_wp) % W;
      px = hist_pose[(hb + slot) * 3];
      py = hist_pose[(hb + slot) * 3 + 1];
      pw = hist_pose[(hb + slot) * 3 + 2];
    }
    if (lane == 0) {
      tok_pose[(size_t)ba * 3] = px; tok_pose[(size_t)ba * 3 + 1] = py; tok_pose[(size_t)ba * 3 + 2] = pw;
      tok_invalid[ba] = last_wp < 0;
    }
    float sn, cs;
    sincosf(pw, &sn, &cs);

    // this lane's two history rows (MMA rows g and g+8 = window positions)
    bool present[2], valid[2];
    float lx[2], ly[2], lw[2], mo[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int wp = g + 8 * h;
      present[h] = wp < W && wp >= W - n_step;
      const int slot = present[h] ? (s - W + wp) % W : 0;
      valid[h] = present[h] && hist_valid[hb + slot];
      lx[h] = ly[h] = lw[h] = 0.f;
      mo[h][0] = mo[h][1] = mo[h][2] = 0.f;
      if (present[h]) {
        const float* hp = hist_pose + (hb + slot) * 3;
        const float dx = hp[0] - px, dy = hp[1] - py;
        lx[h] = fmaf(dx, cs, dy * sn);   // agent_encoder.py:147-148 (token frame, yaw not wrapped)
        ly[h] = fmaf(dy, cs, -dx * sn);
        lw[h] = hp[2] - pw;
        const float* hm = hist_motion + (hb + slot) * 3;
        mo[h][0] = hm[0]; mo[h][1] = hm[1]; mo[h][2] = hm[2];
      }
    }
    // ---- input rows [attr6 | motion3 | one-hot W | 0 pad to 32] as A fragments (2 k-chunks)
    auto attr_el = [&](int h, int col) -> float {
      if (!present[h]) return 0.f;
      if (col < 6) return __ldg(ag_attr + (size_t)ba * 6 + col);
      if (col < 9) return col == 6 ? mo[h][0] : (col == 7 ? mo[h][1] : mo[h][2]);
      if (col < 9 + W) return (col - 9) == (g + 8 * h) ? 1.f : 0.f;
      return 0.f;
    };
    uint32_t a1[2][4];
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      const int c0 = 16 * kc + 2 * t;
      a1[kc][0] = pack_h2(attr_el(0, c0), attr_el(0, c0 + 1));
      a1[kc][1] = pack_h2(attr_el(1, c0), attr_el(1, c0 + 1));
      a1[kc][2] = pack_h2(attr_el(0, c0 + 8), attr_el(0, c0 + 9));
      a1[kc][3] = pack_h2(attr_el(1, c0 + 8), attr_el(1, c0 + 9));
    }
    // ---- input MLP (input_encoder.py:41-61): Linear+ReLU, Linear+ReLU, Linear
    float c[8][4];
    uint32_t x[8][4];  // PointNet input rows [mlp(64) | pe(64)] as 8 k-chunks
    dense64<2>(a1, sW + OFF_W1, S1, sB, g, t, c);
    c_to_a<true>(c, x, hm);
    {
      uint32_t a2[4][4];
#pragma unroll
      for (int kc = 0; kc < 4; ++kc)
#pragma unroll
        for (int r = 0; r < 4; ++r) a2[kc][r] = x[kc][r];
      dense64<4>(a2, sW + OFF_W2, S2, sB + 64, g, t, c);
      c_to_a<true>(c, a2, hm);
      dense64<4>(a2, sW + OFF_W3, S2, sB + 128, g, t, c);
      c_to_a<false>(c, x, hm);
    }
    // ---- PoseEmb64 of the history pose in the token frame (agent_encoder.py:159) -> k-chunks 4..7
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      const int c0 = 16 * kc + 2 * t;
      x[4 + kc][0] = pack_h2(pe64(c0, lx[0], ly[0], lw[0], fr), pe64(c0 + 1, lx[0], ly[0], lw[0], fr));
      x[4 + kc][1] = pack_h2(pe64(c0, lx[1], ly[1], lw[1], fr), pe64(c0 + 1, lx[1], ly[1], lw[1], fr));
      x[4 + kc][2] = pack_h2(pe64(c0 + 8, lx[0], ly[0], lw[0], fr), pe64(c0 + 9, lx[0], ly[0], lw[0], fr));
      x[4 + kc][3] = pack_h2(pe64(c0 + 8, lx[1], ly[1], lw[1], fr), pe64(c0 + 9, lx[1], ly[1], lw[1], fr));
    }
    // ---- PointNet (polyline_encoder.py:50-53): h = ReLU(Linear([h | max])), max over the valid steps
    float m[8][2];
    dense64<8>(x, sW + OFF_P0, SP, sB + 192, g, t, c);
    relu_colmax(c, valid[0], valid[1], m);
#pragma unroll
    for (int layer = 1; layer < 3; ++layer) {
      c_to_a<false>(c, x, hm);  // h (already ReLU'd) -> k-chunks 0..3
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {  // the group max replicated over all rows -> k-chunks 4..7
        x[4 + kc][0] = x[4 + kc][1] = pack_h2(m[2 * kc][0], m[2 * kc][1]);
        x[4 + kc][2] = x[4 + kc][3] = pack_h2(m[2 * kc + 1][0], m[2 * kc + 1][1]);
      }
      dense64<8>(x, sW + (layer == 1 ? OFF_P1 : OFF_P2), SP, sB + 192 + 64 * layer, g, t, c);
      relu_colmax(c, valid[0], valid[1], m);
    }
    // ---- token = max_valid([h | max]) = [max h | max h] (pooling.py:38); all steps invalid -> 0
    float2 o = make_float2(0.f, 0.f);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      if (nt == g) o = make_float2(m[nt][0], m[nt][1]);  // m is replicated over g: lane (g, t) writes n-tile g
    if (last_wp < 0) o = make_float2(0.f, 0.f);
    float* op = tok_out + (size_t)ba * ldo + 8 * g + 2 * t;
    *reinterpret_cast<float2*>(op) = o;
    *reinterpret_cast<float2*>(op + 64) = o;
    if (ln_out) {  // LayerNorm of the 128-wide token = of its 64 distinct values (each lane holds two of them)
      float sum = o.x + o.y;
#pragma unroll
      for (int of = 16; of; of >>= 1) sum += __shfl_xor_sync(TB_FULL_MASK, sum, of);
      const float mean = sum * (1.f / 64.f);
      const float dx = o.x - mean, dy = o.y - mean;
      float sq = fmaf(dx, dx, dy * dy);
#pragma unroll
      for (int of = 16; of; of >>= 1) sq += __shfl_xor_sync(TB_FULL_MASK, sq, of);
      const float rstd = 1.f / sqrtf(sq * (1.f / 64.f) + 1e-5f);
      const int c0 = 8 * g + 2 * t;
      __half* lp = ln_out + (size_t)ba * ld_ln + c0;
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        const float2 gg = *reinterpret_cast<const float2*>(ln_g + c0 + 64 * hlf);
        const float2 bb = *reinterpret_cast<const float2*>(ln_b + c0 + 64 * hlf);
        *reinterpret_cast<__half2*>(lp + 64 * hlf) = __floats2half2_rn(dx * rstd * gg.x + bb.x, dy * rstd * gg.y + bb.y);
      }
    }
  }
  tb_flag_if_sat(hm, sat_flag);
}

}  // namespace

extern "C" int tb_ag_frontend_blob_halves(void) { return W_HALVES; }

extern "C" int tb_ag_frontend_ex(const uint8_t* hist_valid, const float* hist_pose, const float* hist_motion,
                                 const float* ag_attr, const int* d_step, int step_stride, const float* freq_xy, int B,
                                 int A, int W, const void* wblob, const float* bias, float* tok_out, int ldo,
                                 float* tok_pose, uint8_t* tok_invalid, const float* ln_gamma, const float* ln_beta,
                                 void* ln_out, int ld_ln, void* stream) {
  if (!hist_valid || !hist_pose || !hist_motion || !ag_attr || !d_step || !freq_xy || !wblob || !bias || !tok_out ||
      !tok_pose || !tok_invalid)
    return TB_ERR_NULL;
  if (B <= 0 || A <= 0 || W <= 0 || ldo < 128 || step_stride < 0) return TB_ERR_BAD_SHAPE;
  if (W > 16 || 9 + W > 32) return TB_ERR_UNSUPPORTED;
  if ((ldo & 1) || !tb_aligned16(wblob) || (reinterpret_cast<uintptr_t>(tok_out) & 7)) return TB_ERR_MISALIGNED;
  if (ln_out) {
    if (!ln_gamma || !ln_beta) return TB_ERR_NULL;
    if (ld_ln < 128) return TB_ERR_BAD_SHAPE;
    if ((ld_ln & 1) || (reinterpret_cast<uintptr_t>(ln_out) & 3) || (reinterpret_cast<uintptr_t>(ln_gamma) & 7) ||
        (reinterpret_cast<uintptr_t>(ln_beta) & 7))
      return TB_ERR_MISALIGNED;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(ag_frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM) != cudaSuccess)
      return TB_ERR_CUDA;
    attr_set = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0)
      num_sms = 148;
  }
  const int n = B * A;
  const int want = (n + AW - 1) / AW;
  const int grid = want < 2 * num_sms ? want : 2 * num_sms;  // persistent: 2 CTAs per SM, warps loop over agents
  ag_frontend_kernel<<<grid, AW * 32, SMEM, static_cast<cudaStream_t>(stream)>>>(
      hist_valid, hist_pose, hist_motion, ag_attr, d_step, step_stride, A, freq_xy, n, W,
      static_cast<const __half*>(wblob), bias, tok_out, ldo, tok_pose, tok_invalid, ln_gamma, ln_beta,
      static_cast<__half*>(ln_out), ld_ln, tb_fp16_flag_ptr);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_ag_frontend(const uint8_t* hist_valid, const float* hist_pose, const float* hist_motion,
                              const float* ag_attr, const int* d_step, const float* freq_xy, int B, int A, int W,
                              const void* wblob, const float* bias, float* tok_out, int ldo, float* tok_pose,
                              uint8_t* tok_invalid, const float* ln_gamma, const float* ln_beta, void* ln_out, int ld_ln,
                              void* stream) {
  return tb_ag_frontend_ex(hist_valid, hist_pose, hist_motion, ag_attr, d_step, 0, freq_xy, B, A, W, wblob, bias, tok_out,
                           ldo, tok_pose, tok_invalid, ln_gamma, ln_beta, ln_out, ld_ln, stream);
}
