// Fused pairwise relative pose + top-K nearest-target selection (tb_knn_select).
// Reference behaviour: utils/rpe.py:9-37 (get_rel_pose) + :62-90 (get_tgt_knn_idx).
//
// One warp per source token. The scene's targets (x, y, yaw, invalid) are staged once per CTA in shared
// memory and shared by all of the CTA's source rows; each lane keeps T/32 candidate keys (squared distances) in
// registers (never materialising the [S,T,3] rel-pose / [S,T] distance tensors of the reference).
// Selection: (1) prune — the per-lane two smallest keys give 64 candidates whose maximum U bounds the K-th smallest
// from above, so only keys <= U (typically ~15 % of T) survive and are ballot-compacted into a per-warp list;
// (2) the K-th smallest of the survivors is found by bisection on the fp32 bit pattern (non-negative floats order
// like unsigned ints) with one REDUX per probe; (3) winners are compacted in ascending target-index order (ties: lower
// index) and their rotated relative pose / exact distance are evaluated from shared memory. Rows whose survivor list
// would overflow (many invalid targets, K > 64, T <= 128) take the general path: bisection over all T keys.
#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kRowsPerWarp = 8;  // source rows per warp => 64 rows per CTA amortise the target staging
constexpr int kCap = 12;         // pruned candidate list: up to kCap*32 survivors per row

// K-th smallest of the warp's keys (N per lane) by bisection on the bit pattern; returns tau and #keys < tau.
template <int N>
__device__ __forceinline__ uint32_t kth_smallest(const uint32_t (&key)[N], int K, uint32_t hi, int* c_lt_out,
                                                 uint32_t lo = 0u) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    int c = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) c += (key[i] <= mid);
    c = __reduce_add_sync(TB_FULL_MASK, c);
    if (c >= K) hi = mid; else lo = mid + 1u;
  }
  int c_lt = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) c_lt += (key[i] < lo);
  *c_lt_out = __reduce_add_sync(TB_FULL_MASK, c_lt);
  return lo;
}

struct RowCtx {
  float sx, sy, syaw, sn, cs, dist_limit;
  const float *tx, *ty, *tyaw;
  const uint8_t* tinv;
  bool sinv;
  const int32_t* tmap;  // staged target position -> caller's target index (NULL: identity)
  int32_t* out_idx; uint8_t* out_invalid; float* out_rel;
  size_t obase;
};

// winner t -> outputs (relative pose in the source frame, utils/rpe.py:26-36; invalid flag :85-86)
__device__ __forceinline__ void emit(const RowCtx& c, int pos, int t) {
  const float dx = c.tx[t] - c.sx, dy = c.ty[t] - c.sy;
  const float lx = fmaf(dx, c.cs, dy * c.sn), ly = fmaf(dy, c.cs, -dx * c.sn);
  const bool dead = c.sinv || c.tinv[t];
  const float d = dead ? __int_as_float(0x7f800000) : sqrtf(fmaf(lx, lx, ly * ly));
  c.out_idx[c.obase + pos] = c.tmap ? c.tmap[t] : t;
  c.out_invalid[c.obase + pos] = (uint8_t)((c.tinv[t] != 0) | (d > c.dist_limit));
  float* r = c.out_rel + (c.obase + pos) * 3;
  r[0] = lx; r[1] = ly; r[2] = c.tyaw[t] - c.syaw;
}

// K smallest of a compacted candidate list (keys + staged target positions, ascending position order): bisection on
// [lo, hi], winners emitted in list order (ties: earlier entries). Returns the K-th smallest key.
template <int NC>
__device__ __forceinline__ uint32_t select_from_list(const RowCtx& c, const uint32_t* ckey, const uint16_t* cidx, int cnt,
                                                     int K, uint32_t lo, uint32_t hi, int lane, unsigned lt_mask) {
  uint32_t ck[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) ck[j] = (j * 32 + lane < cnt) ? ckey[j * 32 + lane] : 0xffffffffu;
  int c_lt;
  const uint32_t tau = kth_smallest<NC>(ck, K, hi, &c_lt, lo);
  const int need_ties = K - c_lt;
  int n_out = 0, n_tie = 0;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    if (j * 32 >= cnt) break;  // warp-uniform
    const bool is_tie = ck[j] == tau;
    const unsigned tie_b = __ballot_sync(TB_FULL_MASK, is_tie);
    const bool sel = (ck[j] < tau) || (is_tie && n_tie + __popc(tie_b & lt_mask) < need_ties);
    const unsigned sel_b = __ballot_sync(TB_FULL_MASK, sel);
    if (sel) emit(c, n_out + __popc(sel_b & lt_mask), cidx[j * 32 + lane]);
    n_out += __popc(sel_b);
    n_tie += __popc(tie_b);
  }
  return tau;
}

template <int TPL>
__global__ void __launch_bounds__(kWarps * 32)
knn_select_kernel(const float* __restrict__ src_pose, const uint8_t* __restrict__ src_invalid,
                  const float* __restrict__ tgt_pose, const uint8_t* __restrict__ tgt_invalid, int S, int T, int div,
                  int K, float dist_limit, int32_t* __restrict__ out_idx, uint8_t* __restrict__ out_invalid,
                  float* __restrict__ out_rel, int ldk, int koff, const int32_t* __restrict__ tgt_index_map,
                  float* __restrict__ row_state, int sorted_by_x) {
  extern __shared__ float smem[];
  float* tx = smem;
  float* ty = tx + T;
  float* tyaw = ty + T;
  uint8_t* tinv = reinterpret_cast<uint8_t*>(tyaw + T);
  // pruned candidate lists (fast path), one per warp: keys then target indices
  uint32_t* ckey_all = reinterpret_cast<uint32_t*>(smem + 3 * T + ((T + 15) / 16) * 4);
  uint16_t* cidx_all = reinterpret_cast<uint16_t*>(ckey_all + kWarps * kCap * 32);

  const int b = blockIdx.y;
  const int bt = b / div;
  const float* tp = tgt_pose + (size_t)bt * T * 3;
  const uint8_t* ti = tgt_invalid + (size_t)bt * T;
  const int32_t* tmap = tgt_index_map ? tgt_index_map + (size_t)bt * T : nullptr;
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
  uint32_t* ckey = ckey_all + warp * kCap * 32;
  uint16_t* cidx = cidx_all + warp * kCap * 32;

  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    const int s = row0 + rr;
    if (s >= S) break;  // warp-uniform
    const size_t row = (size_t)b * S + s;
    RowCtx c;
    c.sx = src_pose[row * 3 + 0]; c.sy = src_pose[row * 3 + 1]; c.syaw = src_pose[row * 3 + 2];
    c.sinv = src_invalid[row] != 0;
    sincosf(c.syaw, &c.sn, &c.cs);
    c.dist_limit = dist_limit; c.tx = tx; c.ty = ty; c.tyaw = tyaw; c.tinv = tinv;
    c.out_idx = out_idx; c.out_invalid = out_invalid; c.out_rel = out_rel; c.tmap = tmap;
    c.obase = row * (size_t)ldk + koff;
    float* rs = row_state ? row_state + row * 3 : nullptr;  // (x, y, K-th smallest squared distance) of the last call

    if (c.sinv) {  // invalid source: every distance is +inf, the (always masked) fillers are targets 0..K-1
      for (int pos = lane; pos < K; pos += 32) emit(c, pos, pos);
      if (rs && lane == 0) rs[2] = __int_as_float(0x7f800000);
      continue;
    }

    // ---- temporal-coherence fast path (static targets sorted by x, e.g. agent -> map): at least K valid targets lay
    // within sqrt(kth_prev) of the previous source position, so they lie within r = sqrt(kth_prev) + |displacement|
    // of the current one (triangle inequality): the K nearest are among the targets with |x - sx| <= r and d <= r.
    // Only that x-slab is scanned (binary search in the staged, x-sorted coordinates); everything else is unchanged.
    if (rs && sorted_by_x && TPL >= 8) {
      const float kprev = rs[2];
      if (kprev < 3.0e38f) {
        const float ddx = c.sx - rs[0], ddy = c.sy - rs[1];
        const float r = (sqrtf(kprev) + sqrtf(fmaf(ddx, ddx, ddy * ddy))) * 1.00001f + 1e-3f;
        const float r2 = r * r;
        int lo = 0, hi = T;  // lo = first t with tx[t] >= sx - r, hi = first t with tx[t] > sx + r
        {
          int a = 0, b2 = T;
          const float xl = c.sx - r, xh = c.sx + r;
          while (a < b2) { const int m = (a + b2) >> 1; if (tx[m] < xl) a = m + 1; else b2 = m; }
          lo = a;
          b2 = T;
          while (a < b2) { const int m = (a + b2) >> 1; if (tx[m] <= xh) a = m + 1; else b2 = m; }
          hi = a;
        }
        __syncwarp();
        int cnt = 0;
        bool overflow = false;
        for (int base = lo; base < hi; base += 32) {  // warp-uniform bounds
          const int t = base + lane;
          uint32_t k = 0xffffffffu;
          if (t < hi) {
            const float dx = tx[t] - c.sx, dy = ty[t] - c.sy;
            k = tinv[t] ? 0x7f800000u : __float_as_uint(fmaf(dx, dx, dy * dy));
          }
          const bool sel = k <= __float_as_uint(r2);
          const unsigned sb = __ballot_sync(TB_FULL_MASK, sel);
          if (cnt + __popc(sb) > kCap * 32) { overflow = true; break; }
          if (sel) {
            const int p = cnt + __popc(sb & lt_mask);
            ckey[p] = k;
            cidx[p] = (uint16_t)t;
          }
          cnt += __popc(sb);
        }
        __syncwarp();
        if (!overflow && cnt >= K) {
          // the K-th distance cannot have shrunk by more than the displacement either: bracket for the bisection
          const float dl = (sqrtf(kprev) - sqrtf(fmaf(ddx, ddx, ddy * ddy))) * 0.9999f - 1e-3f;
          const uint32_t lo_bits = dl > 0.f ? __float_as_uint(dl * dl) : 0u;
          uint32_t tau;
          if (cnt <= 4 * 32) {  // the usual case: K + a few survivors
            tau = select_from_list<4>(c, ckey, cidx, cnt, K, lo_bits, __float_as_uint(r2), lane, lt_mask);
          } else {
            tau = select_from_list<kCap>(c, ckey, cidx, cnt, K, lo_bits, __float_as_uint(r2), lane, lt_mask);
          }
          if (lane == 0) { rs[0] = c.sx; rs[1] = c.sy; rs[2] = __uint_as_float(tau); }
          continue;
        }
      }
    }

    // selection keys of this lane's candidates t = i*32 + lane: squared distance (the rotation into the source frame
    // preserves it; the winners' rotated pose and exact distance are evaluated in emit()), +inf if either end invalid
    uint32_t key[TPL];
#pragma unroll
    for (int i = 0; i < TPL; ++i) {
      const int t = i * 32 + lane;
      uint32_t k = 0xffffffffu;  // beyond T: never selected (K < T)
      if (t < T) {
        const float dx = tx[t] - c.sx, dy = ty[t] - c.sy;
        k = (c.sinv || tinv[t]) ? 0x7f800000u : __float_as_uint(fmaf(dx, dx, dy * dy));
      }
      key[i] = k;
    }

    bool done = false;
    if (TPL >= 8 && K <= 64) {
      // ---- prune: the 64 (32) per-lane two-smallest (smallest) keys bound the K-th smallest from above
      uint32_t m1 = 0xffffffffu, m2 = 0xffffffffu;
#pragma unroll
      for (int i = 0; i < TPL; ++i) {
        m2 = min(m2, max(m1, key[i]));
        m1 = min(m1, key[i]);
      }
      const uint32_t U = __reduce_max_sync(TB_FULL_MASK, K <= 32 ? m1 : m2);
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < TPL; ++i) cnt += (key[i] <= U);
      cnt = __reduce_add_sync(TB_FULL_MASK, cnt);
      if (cnt <= kCap * 32) {  // warp-uniform
        __syncwarp();
        int base = 0;
#pragma unroll
        for (int i = 0; i < TPL; ++i) {  // compaction keeps ascending target-index order
          const bool sel = key[i] <= U;
          const unsigned sb = __ballot_sync(TB_FULL_MASK, sel);
          if (sel) {
            const int p = base + __popc(sb & lt_mask);
            ckey[p] = key[i];
            cidx[p] = (uint16_t)(i * 32 + lane);
          }
          base += __popc(sb);
        }
        __syncwarp();
        uint32_t ck[kCap];
#pragma unroll
        for (int j = 0; j < kCap; ++j) ck[j] = (j * 32 + lane < cnt) ? ckey[j * 32 + lane] : 0xffffffffu;
        int c_lt;
        const uint32_t tau = kth_smallest<kCap>(ck, K, U, &c_lt);
        const int need_ties = K - c_lt;
        int n_out = 0, n_tie = 0;
#pragma unroll
        for (int j = 0; j < kCap; ++j) {
          if (j * 32 >= cnt) break;  // warp-uniform
          const bool is_tie = ck[j] == tau;
          const unsigned tie_b = __ballot_sync(TB_FULL_MASK, is_tie);
          const bool sel = (ck[j] < tau) || (is_tie && n_tie + __popc(tie_b & lt_mask) < need_ties);
          const unsigned sel_b = __ballot_sync(TB_FULL_MASK, sel);
          if (sel) emit(c, n_out + __popc(sel_b & lt_mask), cidx[j * 32 + lane]);
          n_out += __popc(sel_b);
          n_tie += __popc(tie_b);
        }
        if (rs && lane == 0) { rs[0] = c.sx; rs[1] = c.sy; rs[2] = __uint_as_float(tau); }
        done = true;
      }
    }
    if (done) continue;

    // ---- general path: bisection over all T keys (bracketed by the previous call's K-th distance when the caller
    // keeps a row state: static targets, so |sqrt(kth) - sqrt(kth_prev)| <= |displacement|)
    uint32_t b_lo = 0u, b_hi = 0x7f800000u;
    if (rs && rs[2] < 3.0e38f) {
      const float ddx = c.sx - rs[0], ddy = c.sy - rs[1];
      const float disp = sqrtf(fmaf(ddx, ddx, ddy * ddy)), dk = sqrtf(rs[2]);
      const float dh = (dk + disp) * 1.00001f + 1e-3f, dl = (dk - disp) * 0.9999f - 1e-3f;
      b_hi = __float_as_uint(dh * dh);
      b_lo = dl > 0.f ? __float_as_uint(dl * dl) : 0u;
    }
    int c_lt;
    const uint32_t tau = kth_smallest<TPL>(key, K, b_hi, &c_lt, b_lo);
    const int need_ties = K - c_lt;  // >= 1
    int n_out = 0, n_tie = 0;
#pragma unroll
    for (int i = 0; i < TPL; ++i) {  // compaction in ascending target index (i-major, then lane)
      const bool is_tie = key[i] == tau;
      const unsigned tie_b = __ballot_sync(TB_FULL_MASK, is_tie);
      const bool sel = (key[i] < tau) || (is_tie && n_tie + __popc(tie_b & lt_mask) < need_ties);
      const unsigned sel_b = __ballot_sync(TB_FULL_MASK, sel);
      if (sel) emit(c, n_out + __popc(sel_b & lt_mask), i * 32 + lane);
      n_out += __popc(sel_b);
      n_tie += __popc(tie_b);
    }
    if (rs && lane == 0) { rs[0] = c.sx; rs[1] = c.sy; rs[2] = __uint_as_float(tau); }
  }
}

template <int TPL>
int launch(const float* src_pose, const uint8_t* src_invalid, const float* tgt_pose, const uint8_t* tgt_invalid,
           int B, int S, int T, int div, int K, float dist_limit, int32_t* out_idx, uint8_t* out_invalid,
           float* out_rel, int ldk, int koff, const int32_t* tgt_index_map, float* row_state, int sorted_by_x,
           cudaStream_t st) {
  dim3 grid((S + kWarps * kRowsPerWarp - 1) / (kWarps * kRowsPerWarp), B);
  size_t smem = (size_t)T * 12 + ((T + 15) / 16) * 16 + (size_t)kWarps * kCap * 32 * 6;
  knn_select_kernel<TPL><<<grid, kWarps * 32, smem, st>>>(src_pose, src_invalid, tgt_pose, tgt_invalid, S, T, div, K,
                                                         dist_limit, out_idx, out_invalid, out_rel, ldk, koff,
                                                         tgt_index_map, row_state, sorted_by_x);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

}  // namespace

extern "C" int tb_knn_select(const float* src_pose, const uint8_t* src_invalid, const float* tgt_pose,
                             const uint8_t* tgt_invalid, int B, int S, int T, int tgt_batch_div, int K,
                             float dist_limit, int32_t* out_idx, uint8_t* out_invalid, float* out_rel, int out_ldk,
                             int out_koff, const int32_t* tgt_index_map, float* row_state, int sorted_by_x,
                             void* stream) {
  if (!src_pose || !src_invalid || !tgt_pose || !tgt_invalid || !out_idx || !out_invalid || !out_rel)
    return TB_ERR_NULL;
  if (B <= 0 || S <= 0 || T <= 0 || tgt_batch_div <= 0 || out_koff < 0 || out_ldk < out_koff + K)
    return TB_ERR_BAD_SHAPE;
  if (!(0 < K && K < T)) return TB_ERR_KNN_RANGE;
  if (B > 65535) return TB_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define TB_KNN_ARGS src_pose, src_invalid, tgt_pose, tgt_invalid, B, S, T, tgt_batch_div, K, dist_limit, out_idx, \
                    out_invalid, out_rel, out_ldk, out_koff, tgt_index_map, row_state, sorted_by_x, st
  if (T <= 64) return launch<2>(TB_KNN_ARGS);
  if (T <= 128) return launch<4>(TB_KNN_ARGS);
  if (T <= 256) return launch<8>(TB_KNN_ARGS);
  if (T <= 512) return launch<16>(TB_KNN_ARGS);
  if (T <= 1024) return launch<32>(TB_KNN_ARGS);
  if (T <= 2048) return launch<64>(TB_KNN_ARGS);
#undef TB_KNN_ARGS
  return TB_ERR_UNSUPPORTED;
}
