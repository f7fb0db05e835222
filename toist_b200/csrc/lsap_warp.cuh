// Warp-parallel rectangular linear-sum-assignment shared by the matcher (matcher.cu) and the soft-KD loss (distill.cu).
#pragma once
#include <stdint.h>

namespace toist {

// ------------------------------------------------------------------------------------------------ batched LSAP
// One warp per problem.  Same shortest-augmenting-path algorithm, scan order and tie rule as lsap.h (scipy's
// linear_sum_assignment), with the inner scan over the remaining columns spread over the 32 lanes:
//   sequential rule: walk `remaining` in order; take j when shortest[j] < lowest, or == lowest and j is unassigned.
//   => the chosen position is the LAST unassigned column among the minima if one exists, else the FIRST minimum.
// Arithmetic is float64 on the same operands in the same order per column, so reduced costs are bit identical.
constexpr int kLsapMax = 128;

struct LsapShared {
  double u[kLsapMax], v[kLsapMax], shortest[kLsapMax];
  int path[kLsapMax], col4row[kLsapMax], row4col[kLsapMax], remaining[kLsapMax], pos[kLsapMax];
  uint8_t SR[kLsapMax], SC[kLsapMax];
};

__device__ __forceinline__ double warp_min_f64(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
__device__ __forceinline__ int warp_max_i32(int x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
__device__ __forceinline__ int warp_min_i32(int x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

// cost(i, j) = c[i * si + j * sj]; nr <= nc <= kLsapMax.  Returns 0, or -1 when infeasible (warp uniform).
static __device__ int lsap_warp(const float* __restrict__ c, long long si, long long sj, int nr, int nc, LsapShared& w) {
  const int lane = threadIdx.x & 31;
  const double kInf = 1.0 / 0.0;
  for (int i = lane; i < nr; i += 32) {
    w.u[i] = 0.0;
    w.col4row[i] = -1;
  }
  for (int j = lane; j < nc; j += 32) {
    w.v[j] = 0.0;
    w.row4col[j] = -1;
  }
  __syncwarp();
  for (int cur = 0; cur < nr; ++cur) {
    double min_val = 0.0;
    int i = cur;
    int num_remaining = nc;
    for (int it = lane; it < nc; it += 32) {
      w.remaining[it] = nc - it - 1;
      w.pos[nc - it - 1] = it;
      w.shortest[it] = kInf;
      w.SC[it] = 0;
      w.path[it] = -1;
    }
    for (int r = lane; r < nr; r += 32) w.SR[r] = 0;
    __syncwarp();
    int sink = -1;
    while (sink == -1) {
      if (lane == 0) w.SR[i] = 1;
      const double ui = w.u[i];
      double lowest = kInf;
      for (int j = lane; j < nc; j += 32) {
        if (w.SC[j]) continue;
        const double r = min_val + (double)c[i * si + j * sj] - ui - w.v[j];
        if (r < w.shortest[j]) {
          w.path[j] = i;
          w.shortest[j] = r;
        }
        lowest = fmin(lowest, w.shortest[j]);
      }
      lowest = warp_min_f64(lowest);
      if (lowest == kInf) return -1;
      int last_free = -1, first_any = 0x7fffffff;
      for (int j = lane; j < nc; j += 32) {
        if (w.SC[j] || w.shortest[j] != lowest) continue;
        const int p = w.pos[j];
        first_any = min(first_any, p);
        if (w.row4col[j] == -1) last_free = max(last_free, p);
      }
      last_free = warp_max_i32(last_free);
      first_any = warp_min_i32(first_any);
      const int index = last_free >= 0 ? last_free : first_any;
      min_val = lowest;
      const int j = w.remaining[index];
      const int rj = w.row4col[j];
      __syncwarp();
      if (lane == 0) {
        w.SC[j] = 1;
        const int moved = w.remaining[--num_remaining];
        w.remaining[index] = moved;
        w.pos[moved] = index;
      } else {
        --num_remaining;
      }
      if (rj == -1)
        sink = j;
      else
        i = rj;
      __syncwarp();
    }
    if (lane == 0) w.u[cur] += min_val;
    for (int r = lane; r < nr; r += 32)
      if (w.SR[r] && r != cur) w.u[r] += min_val - w.shortest[w.col4row[r]];
    for (int j = lane; j < nc; j += 32)
      if (w.SC[j]) w.v[j] -= min_val - w.shortest[j];
    __syncwarp();
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int r = w.path[j];
        w.row4col[j] = r;
        const int prev = w.col4row[r];
        w.col4row[r] = j;
        j = prev;
        if (r == cur) break;
      }
    }
    __syncwarp();
  }
  return 0;
}


}  // namespace toist
