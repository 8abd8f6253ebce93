// Row-wise helpers around the projections: LayerNorm, PointNet pooling, pose embedding, row gather.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

// ---- LayerNorm (transformer_rpe.py:156-171): one warp per LN_ROWS rows, D/32 floats per lane and row, two-pass in
// registers. The rows of a warp are independent load -> shuffle-reduce -> store chains that the scheduler interleaves
// (one row per warp left the kernel latency-bound at ~55 % of the HBM rate on the 65,536 x 128 agent rows).
#ifndef TB_LN_ROWS
#define TB_LN_ROWS 4
#endif
constexpr int LN_ROWS = TB_LN_ROWS;
template <int D, bool OUT_H>  // OUT_H: fp16 rows (operands of a tb_linear precision-2 projection)
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, void* __restrict__ Y_, int ldy, int M, int relu) {
  constexpr int NV = D / 32;
  const int lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * LN_ROWS;
  if (row0 >= M) return;
  float v[LN_ROWS][NV];
#pragma unroll
  for (int r = 0; r < LN_ROWS; ++r) {
    const float* xp = X + (size_t)min(row0 + r, M - 1) * ldx + lane * NV;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 t = *reinterpret_cast<const float4*>(xp + i);
      v[r][i] = t.x; v[r][i + 1] = t.y; v[r][i + 2] = t.z; v[r][i + 3] = t.w;
    }
  }
  float mean[LN_ROWS], rstd[LN_ROWS];
#pragma unroll
  for (int r = 0; r < LN_ROWS; ++r) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += v[r][i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(TB_FULL_MASK, s, o);
    mean[r] = s * (1.f / D);
  }
#pragma unroll
  for (int r = 0; r < LN_ROWS; ++r) {
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) { const float d = v[r][i] - mean[r]; q = fmaf(d, d, q); }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(TB_FULL_MASK, q, o);
    rstd[r] = 1.f / sqrtf(q * (1.f / D) + 1e-5f);
  }
#pragma unroll
  for (int i = 0; i < NV; i += 4) {
    const float4 g = ldg4(gamma + lane * NV + i), bb = ldg4(beta + lane * NV + i);
#pragma unroll
    for (int r = 0; r < LN_ROWS; ++r) {
      if (row0 + r >= M) break;
      float4 o;
      o.x = (v[r][i] - mean[r]) * rstd[r] * g.x + bb.x;
      o.y = (v[r][i + 1] - mean[r]) * rstd[r] * g.y + bb.y;
      o.z = (v[r][i + 2] - mean[r]) * rstd[r] * g.z + bb.z;
      o.w = (v[r][i + 3] - mean[r]) * rstd[r] * g.w + bb.w;
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      if (OUT_H) {
        __half* yh = static_cast<__half*>(Y_) + (size_t)(row0 + r) * ldy + lane * NV;
        const __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
        *reinterpret_cast<uint2*>(yh + i) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      } else {
        float* yp = static_cast<float*>(Y_) + (size_t)(row0 + r) * ldy + lane * NV;
        *reinterpret_cast<float4*>(yp + i) = o;
      }
    }
  }
}

// ---- LayerNorm backward (training path): dx = rstd (g - mean(g) - xhat mean(g xhat)) with
// g = dy gamma, xhat = (x - mean) rstd; dgamma += sum_rows dy xhat, dbeta += sum_rows dy. One warp per row (statistics
// recomputed from x: cheaper than storing them), the parameter gradients reduced per CTA in shared memory and added to
// the (caller-zeroed) global accumulators with one atomic per CTA and column.
template <int D>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ gamma, const float* __restrict__ dY,
                     int lddy, float* __restrict__ dX, int lddx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     int M, int rows_per_cta) {
  constexpr int NV = D / 32;
  __shared__ float s_g[D], s_b[D];
  for (int i = threadIdx.x; i < D; i += blockDim.x) { s_g[i] = 0.f; s_b[i] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ag[NV], ab[NV], gm[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { ag[i] = 0.f; ab[i] = 0.f; gm[i] = __ldg(gamma + lane * NV + i); }
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, M);
  for (int row = r0 + warp; row < r1; row += 8) {
    float x[NV], dy[NV];
    const float* xp = X + (size_t)row * ldx + lane * NV;
    const float* dp = dY + (size_t)row * lddy + lane * NV;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      const float4 a = *reinterpret_cast<const float4*>(xp + i), b = *reinterpret_cast<const float4*>(dp + i);
      x[i] = a.x; x[i + 1] = a.y; x[i + 2] = a.z; x[i + 3] = a.w;
      dy[i] = b.x; dy[i + 1] = b.y; dy[i + 2] = b.z; dy[i + 3] = b.w;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += x[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(TB_FULL_MASK, s, o);
    const float mean = s * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) { x[i] -= mean; q = fmaf(x[i], x[i], q); }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(TB_FULL_MASK, q, o);
    const float rstd = 1.f / sqrtf(q * (1.f / D) + 1e-5f);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      x[i] *= rstd;  // xhat
      ag[i] = fmaf(dy[i], x[i], ag[i]);
      ab[i] += dy[i];
      dy[i] *= gm[i];  // g
      sg += dy[i];
      sgx = fmaf(dy[i], x[i], sgx);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      sg += __shfl_xor_sync(TB_FULL_MASK, sg, o);
      sgx += __shfl_xor_sync(TB_FULL_MASK, sgx, o);
    }
    sg *= 1.f / D; sgx *= 1.f / D;
    float* op = dX + (size_t)row * lddx + lane * NV;
#pragma unroll
    for (int i = 0; i < NV; i += 4)
      *reinterpret_cast<float4*>(op + i) =
          make_float4(rstd * (dy[i] - sg - x[i] * sgx), rstd * (dy[i + 1] - sg - x[i + 1] * sgx),
                      rstd * (dy[i + 2] - sg - x[i + 2] * sgx), rstd * (dy[i + 3] - sg - x[i + 3] * sgx));
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) { atomicAdd(&s_g[lane * NV + i], ag[i]); atomicAdd(&s_b[lane * NV + i], ab[i]); }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) { atomicAdd(dgamma + i, s_g[i]); atomicAdd(dbeta + i, s_b[i]); }
}

// ---- PointNet pooling (polyline_encoder.py:50-53, pooling.py:18-19,38): one warp per group of L rows.
// mode 0: right half <- max over valid rows of left half (broadcast to valid rows); invalid rows <- 0.
// mode 1: out[g] <- max over valid rows of all C2 columns (0 if none valid); mode 2: same, written twice
// ([m | m], the final token of the de-duplicated PointNet: max_valid([h | max]) = [max h | max h]).
__global__ void __launch_bounds__(256)
pointnet_pool_kernel(float* __restrict__ X, int ldx, const uint8_t* __restrict__ invalid, int G, int L, int C2,
                     int mode, float* __restrict__ out, int ldo) {
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= G) return;
  const int C = C2 / 2;
  const uint8_t* inv = invalid + (size_t)g * L;
  float* xg = X + (size_t)g * L * ldx;
  if (mode == 0) {
    for (int c = lane; c < C; c += 32) {
      float m = -INFINITY;
      for (int r = 0; r < L; ++r)
        if (!inv[r]) m = fmaxf(m, xg[(size_t)r * ldx + c]);
      for (int r = 0; r < L; ++r) {
        if (inv[r]) { xg[(size_t)r * ldx + c] = 0.f; xg[(size_t)r * ldx + C + c] = 0.f; }
        else xg[(size_t)r * ldx + C + c] = m;
      }
    }
  } else if (C2 == 64 && L <= 16 && (ldx & 1) == 0 && (ldo & 1) == 0) {
    // hot shape (PointNet rows of d/2 = 64 channels, 11 history steps): one float2 per lane and row, all rows of the
    // group in flight before the max
    float2 v[16];
    unsigned vm = 0u;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (r < L && !inv[r]) {
        v[r] = *reinterpret_cast<const float2*>(xg + (size_t)r * ldx + 2 * lane);
        vm |= 1u << r;
      }
    }
    float2 m = make_float2(-INFINITY, -INFINITY);
#pragma unroll
    for (int r = 0; r < 16; ++r)
      if (vm & (1u << r)) { m.x = fmaxf(m.x, v[r].x); m.y = fmaxf(m.y, v[r].y); }
    if (!vm) m = make_float2(0.f, 0.f);
    *reinterpret_cast<float2*>(out + (size_t)g * ldo + 2 * lane) = m;
    if (mode == 2) *reinterpret_cast<float2*>(out + (size_t)g * ldo + C2 + 2 * lane) = m;
  } else {
    for (int c = lane; c < C2; c += 32) {
      float m = -INFINITY;
      bool any = false;
      for (int r = 0; r < L; ++r)
        if (!inv[r]) { m = fmaxf(m, xg[(size_t)r * ldx + c]); any = true; }
      out[(size_t)g * ldo + c] = any ? m : 0.f;
      if (mode == 2) out[(size_t)g * ldo + C2 + c] = any ? m : 0.f;
    }
  }
}

// ---- one component of PoseEmb("pe_xy_yaw") (pose_emb.py:50-55, positional_emb.py:24-25,40-41)
__device__ __forceinline__ float pe_component(int c, int pe_dim, float x, float y, float w,
                                              const float* __restrict__ freq_xy) {
  const int n = pe_dim >> 3, q4 = pe_dim >> 2;
  float a;
  bool is_sin;
  if (c < q4) { a = x * __ldg(freq_xy + (c % n)); is_sin = c >= n; }
  else if (c < 2 * q4) { const int cc = c - q4; a = y * __ldg(freq_xy + (cc % n)); is_sin = cc >= n; }
  else { const int cc = c - 2 * q4; a = w * (float)((cc % q4) + 1); is_sin = cc >= q4; }
  const float r = tb_reduce_2pi(a);
  return is_sin ? __sinf(r) : __cosf(r);
}

__global__ void __launch_bounds__(256)
pose_emb_kernel(const float* __restrict__ pose, const float* __restrict__ frame, int frame_div,
                const float* __restrict__ freq_xy, int M, int pe_dim, float* __restrict__ out, int ldo) {
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  float x = pose[(size_t)m * 3], y = pose[(size_t)m * 3 + 1], w = pose[(size_t)m * 3 + 2];
  if (frame) {
    const float* f = frame + (size_t)(m / frame_div) * 3;
    float sn, cs;
    sincosf(f[2], &sn, &cs);
    const float dx = x - f[0], dy = y - f[1];
    x = fmaf(dx, cs, dy * sn);
    y = fmaf(dy, cs, -dx * sn);
    w = w - f[2];
  }
  for (int c = lane; c < pe_dim; c += 32) out[(size_t)m * ldo + c] = pe_component(c, pe_dim, x, y, w, freq_xy);
}

__global__ void gather_rows_kernel(const float* __restrict__ table, int ldt, int T, const int32_t* __restrict__ idx,
                                   int M, int rows_per_batch, int div, int C, float* __restrict__ out, int ldo) {
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  const int bt = (m / rows_per_batch) / div;
  const float* src = table + ((size_t)bt * T + idx[m]) * ldt;
  for (int c = lane; c < C; c += 32) out[(size_t)m * ldo + c] = src[c];
}

}  // namespace

// ---- 3xTF32 operand split (fp32-accurate projections on the tf32 tensor cores): out row = [x | x - trunc_tf32(x) | x].
// With W3 = [W | W | W - trunc_tf32(W)] one kind::tf32 GEMM over K' = 3K evaluates x_hi W_hi + x_lo W_hi + x_hi W_lo
// (the tensor core truncates every operand to its 10 explicit mantissa bits; the dropped x_lo W_lo term is 2^-20 relative).
__global__ void tf32_split3_kernel(const float* __restrict__ X, int ldx, int M, int K, float* __restrict__ out, int ldo) {
  const int m = blockIdx.x * 8 + threadIdx.y;
  if (m >= M) return;
  const float* xr = X + (size_t)m * ldx;
  float* orow = out + (size_t)m * ldo;
  for (int c = threadIdx.x * 4; c < K; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    float4 lo;
    lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    *reinterpret_cast<float4*>(orow + c) = v;
    *reinterpret_cast<float4*>(orow + K + c) = lo;
    *reinterpret_cast<float4*>(orow + 2 * K + c) = v;
  }
}

extern "C" int tb_tf32_split3(const float* X, int ldx, int M, int K, float* out, int ldo, void* stream) {
  if (!X || !out) return TB_ERR_NULL;
  if (M <= 0 || K <= 0 || ldx < K || ldo < 3 * K) return TB_ERR_BAD_SHAPE;
  if ((K & 3) || (ldx & 3) || (ldo & 3) || !tb_aligned16(X) || !tb_aligned16(out)) return TB_ERR_MISALIGNED;
  tf32_split3_kernel<<<(M + 7) / 8, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(X, ldx, M, K, out, ldo);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_layernorm(const float* X, int ldx, const float* gamma, const float* beta, void* Y, int ldy, int M,
                            int D, int flags, void* stream) {
  const int relu = flags & 1, out_h = (flags & 2) != 0;
  if (!X || !gamma || !beta || !Y) return TB_ERR_NULL;
  if (M <= 0 || ldx < D || ldy < D) return TB_ERR_BAD_SHAPE;
  if (D != 128 && D != 256) return TB_ERR_UNSUPPORTED;
  if ((ldx | ldy) & 3 || !tb_aligned16(X) || !tb_aligned16(Y) || !tb_aligned16(gamma) || !tb_aligned16(beta))
    return TB_ERR_MISALIGNED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (M + 8 * LN_ROWS - 1) / (8 * LN_ROWS);
  if (out_h && (ldy & 7)) return TB_ERR_MISALIGNED;
  if (D == 128 && out_h) layernorm_kernel<128, true><<<grid, 256, 0, st>>>(X, ldx, gamma, beta, Y, ldy, M, relu);
  else if (D == 128) layernorm_kernel<128, false><<<grid, 256, 0, st>>>(X, ldx, gamma, beta, Y, ldy, M, relu);
  else if (out_h) layernorm_kernel<256, true><<<grid, 256, 0, st>>>(X, ldx, gamma, beta, Y, ldy, M, relu);
  else layernorm_kernel<256, false><<<grid, 256, 0, st>>>(X, ldx, gamma, beta, Y, ldy, M, relu);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_layernorm_bwd(const float* X, int ldx, const float* gamma, const float* dY, int lddy, float* dX, int lddx,
                                float* dgamma, float* dbeta, int M, int D, void* stream) {
  if (!X || !gamma || !dY || !dX || !dgamma || !dbeta) return TB_ERR_NULL;
  if (M <= 0 || ldx < D || lddy < D || lddx < D) return TB_ERR_BAD_SHAPE;
  if (D != 128 && D != 256) return TB_ERR_UNSUPPORTED;
  if ((ldx | lddy | lddx) & 3 || !tb_aligned16(X) || !tb_aligned16(dY) || !tb_aligned16(dX)) return TB_ERR_MISALIGNED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows_per_cta = 256;  // 32 rows per warp: amortises the per-CTA parameter-gradient reduction
  const int grid = (M + rows_per_cta - 1) / rows_per_cta;
  if (D == 128) layernorm_bwd_kernel<128><<<grid, 256, 0, st>>>(X, ldx, gamma, dY, lddy, dX, lddx, dgamma, dbeta, M, rows_per_cta);
  else layernorm_bwd_kernel<256><<<grid, 256, 0, st>>>(X, ldx, gamma, dY, lddy, dX, lddx, dgamma, dbeta, M, rows_per_cta);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_pointnet_pool(float* X, int ldx, const uint8_t* invalid, int G, int L, int C2, int mode, float* out,
                                int ldo, void* stream) {
  if (!X || !invalid || (mode != 0 && !out)) return TB_ERR_NULL;
  if (G <= 0 || L <= 0 || C2 <= 0 || (C2 & 1) || ldx < C2 || mode < 0 || mode > 2) return TB_ERR_BAD_SHAPE;
  pointnet_pool_kernel<<<(G + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(X, ldx, invalid, G, L, C2, mode, out,
                                                                                 ldo);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_pose_emb(const float* pose, const float* frame, int frame_div, const float* freq_xy, int M,
                           int pe_dim, float* out, int ldo, void* stream) {
  if (!pose || !freq_xy || !out) return TB_ERR_NULL;
  if (M <= 0 || ldo < pe_dim || (frame && frame_div <= 0)) return TB_ERR_BAD_SHAPE;
  if (pe_dim != 64 && pe_dim != 128 && pe_dim != 256) return TB_ERR_UNSUPPORTED;
  pose_emb_kernel<<<(M + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(pose, frame, frame_div, freq_xy, M, pe_dim,
                                                                            out, ldo);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

extern "C" int tb_gather_rows(const float* table, int ldt, int T, const int32_t* idx, int M, int rows_per_batch,
                              int div, int C, float* out, int ldo, void* stream) {
  if (!table || !idx || !out) return TB_ERR_NULL;
  if (M <= 0 || T <= 0 || rows_per_batch <= 0 || div <= 0 || C <= 0 || ldt < C || ldo < C) return TB_ERR_BAD_SHAPE;
  gather_rows_kernel<<<(M + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(table, ldt, T, idx, M, rows_per_batch,
                                                                               div, C, out, ldo);
  TB_CHECK_LAUNCH();
  return TB_OK;
}
