// Fused pairwise relative pose + top-K nearest-target selection (tb_knn_select).
// Reference behaviour: utils/rpe.py:9-37 (get_rel_pose) + :62-90 (get_tgt_knn_idx).
//
// One warp per source token. The scene's targets (x, y, yaw, invalid) are staged once per CTA in shared
// memory and shared by all of the CTA's source rows; each lane keeps T/32 candidate distances in registers
// (never materialising the [S,T,3] rel-pose / [S,T] distance tensors of the reference). The K-th smallest
// distance is found by bisection on the fp32 bit pattern (non-negative floats order like unsigned ints) with
// one REDUX per probe; winners are compacted with ballots in ascending target-index order and their
// relative pose is recomputed from shared memory.
#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kRowsPerWarp = 8;  // source rows per warp => 64 rows per CTA amortise the target staging

template <int TPL>
__global__ void __launch_bounds__(kWarps * 32)
knn_select_kernel(const float* __restrict__ src_pose, const uint8_t* __restrict__ src_invalid,
                  const float* __restrict__ tgt_pose, const uint8_t* __restrict__ tgt_invalid, int S, int T, int div,
                  int K, float dist_limit, int32_t* __restrict__ out_idx, uint8_t* __restrict__ out_invalid,
                  float* __restrict__ out_rel, int ldk, int koff) {
  extern __shared__ float smem[];
  float* tx = smem;
  float* ty = tx + T;
  float* tyaw = ty + T;
  uint8_t* tinv = reinterpret_cast<uint8_t*>(tyaw + T);

  const int b = blockIdx.y;
  const int bt = b / div;
  const float* tp = tgt_pose + (size_t)bt * T * 3;
  const uint8_t* ti = tgt_invalid + (size_t)bt * T;
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    tx[i] = tp[i * 3 + 0];
    ty[i] = tp[i * 3 + 1];
    tyaw[i] = tp[i * 3 + 2];
    tinv[i] = ti[i];
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int row0 = (blockIdx.x * kWarps + warp) * kRowsPerWarp;

  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    const int s = row0 + rr;
    if (s >= S) break;  // warp-uniform
    const size_t row = (size_t)b * S + s;
    const float sx = src_pose[row * 3 + 0], sy = src_pose[row * 3 + 1], syaw = src_pose[row * 3 + 2];
    const bool sinv = src_invalid[row] != 0;
    float sn, cs;
    sincosf(syaw, &sn, &cs);

    // distances of this lane's candidates t = i*32 + lane
    uint32_t key[TPL];
#pragma unroll
    for (int i = 0; i < TPL; ++i) {
      const int t = i * 32 + lane;
      uint32_t k = 0xffffffffu;  // beyond T: never selected (K < T)
      if (t < T) {
        float d = __int_as_float(0x7f800000);
        if (!sinv && !tinv[t]) {
          const float dx = tx[t] - sx, dy = ty[t] - sy;
          const float lx = fmaf(dx, cs, dy * sn);
          const float ly = fmaf(dy, cs, -dx * sn);
          d = sqrtf(fmaf(lx, lx, ly * ly));
        }
        k = __float_as_uint(d);
      }
      key[i] = k;
    }

    // smallest tau with count(key <= tau) >= K
    uint32_t lo = 0u, hi = 0x7f800000u;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      int c = 0;
#pragma unroll
      for (int i = 0; i < TPL; ++i) c += (key[i] <= mid);
      c = __reduce_add_sync(TB_FULL_MASK, c);
      if (c >= K) hi = mid; else lo = mid + 1u;
    }
    const uint32_t tau = lo;
    int c_lt = 0;
#pragma unroll
    for (int i = 0; i < TPL; ++i) c_lt += (key[i] < tau);
    c_lt = __reduce_add_sync(TB_FULL_MASK, c_lt);
    const int need_ties = K - c_lt;  // >= 1

    // compaction in ascending target index (i-major, then lane)
    int n_out = 0, n_tie = 0;
    const size_t obase = row * (size_t)ldk + koff;
#pragma unroll
    for (int i = 0; i < TPL; ++i) {
      const bool is_tie = key[i] == tau;
      const unsigned tie_b = __ballot_sync(TB_FULL_MASK, is_tie);
      const int tie_rank = n_tie + __popc(tie_b & lt_mask);
      const bool sel = (key[i] < tau) || (is_tie && tie_rank < need_ties);
      const unsigned sel_b = __ballot_sync(TB_FULL_MASK, sel);
      if (sel) {
        const int pos = n_out + __popc(sel_b & lt_mask);
        const int t = i * 32 + lane;
        const float d = __uint_as_float(key[i]);
        const float dx = tx[t] - sx, dy = ty[t] - sy;
        out_idx[obase + pos] = t;
        out_invalid[obase + pos] = (uint8_t)((tinv[t] != 0) | (d > dist_limit));
        float* r = out_rel + (obase + pos) * 3;
        r[0] = fmaf(dx, cs, dy * sn);
        r[1] = fmaf(dy, cs, -dx * sn);
        r[2] = tyaw[t] - syaw;
      }
      n_out += __popc(sel_b);
      n_tie += __popc(tie_b);
    }
  }
}

template <int TPL>
int launch(const float* src_pose, const uint8_t* src_invalid, const float* tgt_pose, const uint8_t* tgt_invalid,
           int B, int S, int T, int div, int K, float dist_limit, int32_t* out_idx, uint8_t* out_invalid,
           float* out_rel, int ldk, int koff, cudaStream_t st) {
  dim3 grid((S + kWarps * kRowsPerWarp - 1) / (kWarps * kRowsPerWarp), B);
  size_t smem = (size_t)T * 13 + 16;
  knn_select_kernel<TPL><<<grid, kWarps * 32, smem, st>>>(src_pose, src_invalid, tgt_pose, tgt_invalid, S, T, div, K,
                                                         dist_limit, out_idx, out_invalid, out_rel, ldk, koff);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

}  // namespace

extern "C" int tb_knn_select(const float* src_pose, const uint8_t* src_invalid, const float* tgt_pose,
                             const uint8_t* tgt_invalid, int B, int S, int T, int tgt_batch_div, int K,
                             float dist_limit, int32_t* out_idx, uint8_t* out_invalid, float* out_rel, int out_ldk,
                             int out_koff, void* stream) {
  if (!src_pose || !src_invalid || !tgt_pose || !tgt_invalid || !out_idx || !out_invalid || !out_rel)
    return TB_ERR_NULL;
  if (B <= 0 || S <= 0 || T <= 0 || tgt_batch_div <= 0 || out_koff < 0 || out_ldk < out_koff + K)
    return TB_ERR_BAD_SHAPE;
  if (!(0 < K && K < T)) return TB_ERR_KNN_RANGE;
  if (B > 65535) return TB_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define TB_KNN_ARGS src_pose, src_invalid, tgt_pose, tgt_invalid, B, S, T, tgt_batch_div, K, dist_limit, out_idx, \
                    out_invalid, out_rel, out_ldk, out_koff, st
  if (T <= 64) return launch<2>(TB_KNN_ARGS);
  if (T <= 128) return launch<4>(TB_KNN_ARGS);
  if (T <= 256) return launch<8>(TB_KNN_ARGS);
  if (T <= 512) return launch<16>(TB_KNN_ARGS);
  if (T <= 1024) return launch<32>(TB_KNN_ARGS);
  if (T <= 2048) return launch<64>(TB_KNN_ARGS);
#undef TB_KNN_ARGS
  return TB_ERR_UNSUPPORTED;
}
