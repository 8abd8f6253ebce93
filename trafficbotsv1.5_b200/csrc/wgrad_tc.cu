// Weight gradient on the 5th-gen tensor cores (training path):
//   dW[n, k] += sum_m dY[m, n] X[m, k]      (autograd of modules/mlp.py:69 / attention_rpe.py:96-97,186 projections)
// The contraction runs over the ROW index m of two row-major activations, i.e. both tcgen05 operands are "MN-major":
// their M / N dimension (n of dY, k of X) is the contiguous one. No transposed copy is made: TMA boxes of
// {32 columns (128 B), 32 rows} land in shared memory exactly as the canonical MN-major SWIZZLE_128B_BASE32B layout
//   ((4, 8, chunks), (4, k-groups)) : ((1, 4, LBO), (32, SBO))   [elements; cute/atom/mma_traits_sm100.hpp]
// with LBO = 4096 B between 32-column chunks (one box each) and SBO = 512 B between groups of 4 rows; one
// tcgen05.mma kind::tf32 (UMMA_K = 8) consumes two 4-row groups, so advancing K is +1024 B on the start address.
// One CTA = one 128 x 128 tile of dW over a slice of M (split-K, the grid fills the 148 SMs twice), fp32 accumulator in
// 128 TMEM columns, 4-stage mbarrier ring (32 KB per stage), warp-specialised (4 epilogue warps, TMA producer, MMA
// issuer). The epilogue adds the tile into the caller-zeroed dW with vector atomics. N / K tails and the M tail are
// TMA zero fill. The bias gradient (column sums of dY) is tb_colsum's job.
// Shapes outside the TMA constraints (leading dimensions not multiples of 4 floats, misaligned pointers) fall back
// to the FFMA kernel of train_bwd.cu.
#include <cuda.h>
#include "common.cuh"

int tb_linear_wgrad_f32(const float* dY, int lddy, const float* X, int ldx, int M, int N, int K, float* dW, int lddw,
                        float* db, cudaStream_t st);

namespace {

constexpr int TILE = 128;                    // dW tile: 128 (n) x 128 (k)
constexpr int BR = 32;                       // rows of m per stage = 4 MMAs of K = 8
constexpr int CHUNK_BYTES = BR * 128;        // one TMA box: 32 rows x 32 floats
constexpr int OP_BYTES = 4 * CHUNK_BYTES;    // 128 columns of one operand
constexpr int STAGES = 4;
constexpr int EPI_WARPS = 4;
constexpr int NUM_THREADS = (EPI_WARPS + 2) * 32;
constexpr int TMEM_COLS = 128;
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * 2 * OP_BYTES + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WG_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WG_DONE;\n\t"
      "bra WG_WAIT;\n\t"
      "WG_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// MN-major tf32 operand: the only shared-memory layout the tensor core accepts for 32-bit MN-major operands is
// SWIZZLE_128B_BASE32B (layout type 1; cutlass sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout") = Swizzle<2,5,2>: rows of 128 B, atoms of 4 rows, 32-byte pieces XOR-ed with the row index -
// what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B. LBO = byte distance between 32-element column chunks,
// SBO = between 4-row k-groups; descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(CHUNK_BYTES >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// Instruction descriptor: D fp32 (bit 4), A / B format tf32 (2 at bits 7 / 10), A and B MN-major (bits 15 / 16),
// N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                            ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                float* __restrict__ dW, int lddw, int M, int N, int K, int rows_per_split, int vec_ok) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * OP_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * 2 * OP_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TILE, k0 = blockIdx.y * TILE;
  const int m_begin = blockIdx.z * rows_per_split, m_end = min(M, m_begin + rows_per_split);
  const int n_it = (m_end - m_begin + BR - 1) / BR;
  const int a_chunks = min(4, (N - n0 + 31) / 32), b_chunks = min(4, (K - k0 + 31) / 32);  // chunks that hold data

  if (warp == EPI_WARPS && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // chunks past N / K are never loaded: zero them once so that the unused accumulator rows / columns stay finite
  for (int i = threadIdx.x; i < STAGES * 2 * OP_BYTES / 16; i += NUM_THREADS) {
    const int off = i * 16, op = off / OP_BYTES, chunk = (off % OP_BYTES) / CHUNK_BYTES;
    const bool is_b = op >= STAGES;
    if (chunk >= (is_b ? b_chunks : a_chunks)) *reinterpret_cast<uint4*>(smem + off) = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == EPI_WARPS) {
    if (lane == 0) {  // ===== TMA producer =====
      for (int it = 0; it < n_it; ++it) {
        const int s = it % STAGES, m0 = m_begin + it * BR;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[s], (uint32_t)(a_chunks + b_chunks) * CHUNK_BYTES);
        for (int c = 0; c < a_chunks; ++c) tma_load_2d(&mapA, &full[s], sA + s * OP_BYTES + c * CHUNK_BYTES, n0 + 32 * c, m0);
        for (int c = 0; c < b_chunks; ++c) tma_load_2d(&mapB, &full[s], sB + s * OP_BYTES + c * CHUNK_BYTES, k0 + 32 * c, m0);
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      for (int it = 0; it < n_it; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t adesc = make_desc_mn(smem_u32(sA + s * OP_BYTES));
        const uint64_t bdesc = make_desc_mn(smem_u32(sB + s * OP_BYTES));
#pragma unroll
        for (int kk = 0; kk < BR / 8; ++kk)  // 8 rows (two 4-row k-groups) per MMA: +1024 B = +64 in 16-byte units
          umma_tf32(tmem_base, adesc + 64 * kk, bdesc + 64 * kk, kIdesc, (it | kk) != 0);
        umma_commit(&empty[s]);
      }
      umma_commit(tfull);
    }
  } else {
    // ===== epilogue: warp q owns TMEM lanes [32 q, +32) = rows n0 + 32 q + lane of the tile =====
    mbar_wait(tfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int n = n0 + warp * 32 + lane;
    float* wrow = dW + (size_t)n * lddw + k0;
#pragma unroll 1
    for (int c0 = 0; c0 < TILE; c0 += 32) {
      if (k0 + c0 >= K) break;  // warp-uniform
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (n < N) {
        if (vec_ok && k0 + c0 + 32 <= K) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wrow + c0 + j), "f"(__uint_as_float(r[j])),
                         "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (k0 + c0 + j < K) atomicAdd(wrow + c0 + j, __uint_as_float(r[j]));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == EPI_WARPS + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// out[n] += sum over rows of X[m, n] (bias gradient): one thread per column, rows split over gridDim.y.
__global__ void colsum_kernel(const float* __restrict__ X, int ldx, int M, int N, float* __restrict__ out,
                              int rows_per_split) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int m0 = blockIdx.y * rows_per_split, m1 = min(M, m0 + rows_per_split);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int m = m0;
  for (; m + 3 < m1; m += 4) {
    s0 += X[(size_t)m * ldx + n]; s1 += X[(size_t)(m + 1) * ldx + n];
    s2 += X[(size_t)(m + 2) * ldx + n]; s3 += X[(size_t)(m + 3) * ldx + n];
  }
  for (; m < m1; ++m) s0 += X[(size_t)m * ldx + n];
  atomicAdd(out + n, (s0 + s1) + (s2 + s3));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// fp32 [rows, cols] with row stride ld: boxes of 32 columns (128 B) x BR rows, 128-byte swizzle with 32-byte atoms,
// OOB zero fill
bool make_map(CUtensorMap* map, const float* ptr, int rows, int cols, int ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)BR};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

extern "C" int tb_colsum(const float* X, int ldx, int M, int N, float* out, void* stream) {
  if (!X || !out) return TB_ERR_NULL;
  if (M <= 0 || N <= 0 || ldx < N) return TB_ERR_BAD_SHAPE;
  const int bx = (N + 127) / 128;
  int splits = (148 * 8 + bx - 1) / bx;
  if (splits > (M + 63) / 64) splits = (M + 63) / 64;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  const int rows = (M + splits - 1) / splits;
  colsum_kernel<<<dim3(bx, (M + rows - 1) / rows), 128, 0, static_cast<cudaStream_t>(stream)>>>(X, ldx, M, N, out, rows);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

// precision 0: fp32 FFMA (parity); 1: tf32 tcgen05 (operands read as tf32, fp32 accumulate).
extern "C" int tb_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, int M, int N, int K, float* dW,
                               int lddw, float* db, int precision, void* stream) {
  if (!dY || !X || !dW) return TB_ERR_NULL;
  if (M <= 0 || N <= 0 || K <= 0 || lddy < N || ldx < K || lddw < K) return TB_ERR_BAD_SHAPE;
  if (precision != 0 && precision != 1) return TB_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tma_ok = ((lddy | ldx) & 3) == 0 && tb_aligned16(dY) && tb_aligned16(X);
  if (precision == 0 || !tma_ok || M < 4 * BR) return tb_linear_wgrad_f32(dY, lddy, X, ldx, M, N, K, dW, lddw, db, st);
  if (db) {
    const int rc = tb_colsum(dY, lddy, M, N, db, stream);
    if (rc != TB_OK) return rc;
  }
  CUtensorMap mapA, mapB;
  if (!make_map(&mapA, dY, M, N, lddy) || !make_map(&mapB, X, M, K, ldx)) return TB_ERR_CUDA;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess)
      return TB_ERR_CUDA;
    attr_set = true;
  }
  const int tn = (N + TILE - 1) / TILE, tk = (K + TILE - 1) / TILE;
  int splits = (148 * 2 + tn * tk - 1) / (tn * tk);
  const int max_splits = (M + 8 * BR - 1) / (8 * BR);  // at least 8 stages per CTA
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int rows = (M + splits - 1) / splits;
  rows = (rows + BR - 1) / BR * BR;
  splits = (M + rows - 1) / rows;
  const int vec_ok = ((lddw & 3) == 0 && tb_aligned16(dW)) ? 1 : 0;
  wgrad_tc_kernel<<<dim3(tn, tk, splits), NUM_THREADS, SMEM_BYTES, st>>>(mapA, mapB, dW, lddw, M, N, K, rows, vec_ok);
  TB_CHECK_LAUNCH();
  return TB_OK;
}
