// fp32 FFMA projection kernel (tb_linear precision 0): Y = epilogue(X W^T + bias). This is the bit-faithful parity
// path for the dense projections (F.linear call sites: attention_rpe.py:96-97,147,186; transformer_rpe.py:237-238;
// modules/mlp.py:69); the tensor-core path lives in linear_tc.cu.
// 128 x BN x 16 tiles, 256 threads, 8 x (BN/16) register tile, register-prefetch double buffering, fused epilogue.
#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 16, NT = 256;

struct Epi {
  const float* bias;
  int bgroup;  // 0: bias[N]; g > 0: bias[(row / g) * N + n] (one bias row per group of g consecutive rows)
  int relu;
  const uint8_t* mask_pre;
  const float* res;
  int ldr;
  const uint8_t* mask_post;
};

template <int BN, bool VEC>
__global__ void __launch_bounds__(NT)
linear_f32_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, float* __restrict__ Y, int ldy,
                  int M, int N, int K, Epi ep) {
  constexpr int TN = BN / 16;  // 8 or 4 columns per thread
  constexpr int TM = 8;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 thread grid
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // global->smem staging: A tile 128 rows x 16 k = 512 float4 (2 per thread); B tile BN rows x 16 k
  constexpr int A_LD = 2, B_LD = BN / 64;
  float4 ra[A_LD], rb[B_LD];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      const int f = tid + i * NT, r = f >> 2, kk = (f & 3) * 4;
      const int gm = m0 + r, gk = k0 + kk;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gm < M) {
        const float* p = X + (size_t)gm * ldx + gk;
        if (VEC && gk + 3 < K) v = *reinterpret_cast<const float4*>(p);
        else {
          if (gk < K) v.x = p[0];
          if (gk + 1 < K) v.y = p[1];
          if (gk + 2 < K) v.z = p[2];
          if (gk + 3 < K) v.w = p[3];
        }
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int f = tid + i * NT, r = f >> 2, kk = (f & 3) * 4;
      const int gn = n0 + r, gk = k0 + kk;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gn < N) {
        const float* p = W + (size_t)gn * K + gk;
        if (VEC && gk + 3 < K) v = __ldg(reinterpret_cast<const float4*>(p));
        else {
          if (gk < K) v.x = __ldg(p);
          if (gk + 1 < K) v.y = __ldg(p + 1);
          if (gk + 2 < K) v.z = __ldg(p + 2);
          if (gk + 3 < K) v.w = __ldg(p + 3);
        }
      }
      rb[i] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      const int f = tid + i * NT, r = f >> 2, kk = (f & 3) * 4;
      As[buf][kk + 0][r] = ra[i].x; As[buf][kk + 1][r] = ra[i].y;
      As[buf][kk + 2][r] = ra[i].z; As[buf][kk + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int f = tid + i * NT, r = f >> 2, kk = (f & 3) * 4;
      Bs[buf][kk + 0][r] = rb[i].x; Bs[buf][kk + 1][r] = rb[i].y;
      Bs[buf][kk + 2][r] = rb[i].z; Bs[buf][kk + 3][r] = rb[i].w;
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], bb[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][(j / 4) * 64 + tx * 4]);
        bb[j] = b0.x; bb[j + 1] = b0.y; bb[j + 2] = b0.z; bb[j + 3] = b0.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: rows {ty*4 + i, 64 + ty*4 + i}, cols {(j/4)*64 + tx*4 + j%4}
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
    const bool zpre = ep.mask_pre && ep.mask_pre[gm];
    const bool zpost = ep.mask_post && ep.mask_post[gm];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + (j / 4) * 64 + tx * 4 + (j & 3);
      if (gn >= N) continue;
      float v = acc[i][j];
      if (ep.bias) v += __ldg(ep.bias + (ep.bgroup ? (size_t)(gm / ep.bgroup) * N : 0) + gn);
      if (ep.relu) v = fmaxf(v, 0.f);
      if (zpre) v = 0.f;
      if (ep.res) v += ep.res[(size_t)gm * ep.ldr + gn];
      if (zpost) v = 0.f;
      Y[(size_t)gm * ldy + gn] = v;
    }
  }
}

}  // namespace

int tb_linear_f32(const float* X, int ldx, const float* W, const float* bias, int bias_group, float* Y, int ldy, int M,
                  int N, int K,
                  int relu, const uint8_t* mask_pre, const float* res, int ldr, const uint8_t* mask_post,
                  cudaStream_t st) {
  Epi ep{bias, bias_group, relu, mask_pre, res, ldr, mask_post};
  const bool vec = (K % 4 == 0) && (ldx % 4 == 0) && tb_aligned16(X) && tb_aligned16(W);
  const bool wide = N > 64;
  dim3 grid((M + BM - 1) / BM, wide ? (N + 127) / 128 : 1);
  if (wide) {
    if (vec) linear_f32_kernel<128, true><<<grid, NT, 0, st>>>(X, ldx, W, Y, ldy, M, N, K, ep);
    else linear_f32_kernel<128, false><<<grid, NT, 0, st>>>(X, ldx, W, Y, ldy, M, N, K, ep);
  } else {
    if (vec) linear_f32_kernel<64, true><<<grid, NT, 0, st>>>(X, ldx, W, Y, ldy, M, N, K, ep);
    else linear_f32_kernel<64, false><<<grid, NT, 0, st>>>(X, ldx, W, Y, ldy, M, N, K, ep);
  }
  TB_CHECK_LAUNCH();
  return TB_OK;
}
