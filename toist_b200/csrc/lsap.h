// Rectangular linear-sum-assignment shared by the host entry point (toist_lsap_f64) and the device kernel
// (one thread per problem).  Shortest-augmenting-path algorithm of Crouse (2016) as documented for
// scipy.optimize.linear_sum_assignment, which the reference calls at models/matcher.py:85 and
// models/mdetr.py:100,539 (scipy is a compiled third-party dependency, requirements.txt:67).  The scan order, the
// tie-breaking rule (prefer an unassigned column on equal reduced cost) and float64 arithmetic follow that
// documentation so that assignments are identical, ties included.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define TOIST_HD __host__ __device__
#else
#define TOIST_HD
#endif

namespace toist {

struct LsapWork {
  double* u;         // [nr]
  double* v;         // [nc]
  double* shortest;  // [nc]
  int* path;         // [nc]
  int* col4row;      // [nr]
  int* row4col;      // [nc]
  int* remaining;    // [nc]
  uint8_t* SR;       // [nr]
  uint8_t* SC;       // [nc]
};

// cost(i, j) accessor is a functor so callers can present a transposed / strided view.
// Returns 0 on success, -1 when the problem is infeasible.  nr <= nc is required.
template <class Cost>
TOIST_HD inline int lsap_solve(int nr, int nc, Cost cost, LsapWork w) {
  const double kInf = 1.0 / 0.0;
  for (int i = 0; i < nr; ++i) {
    w.u[i] = 0.0;
    w.col4row[i] = -1;
  }
  for (int j = 0; j < nc; ++j) {
    w.v[j] = 0.0;
    w.row4col[j] = -1;
  }
  for (int cur = 0; cur < nr; ++cur) {
    double min_val = 0.0;
    int i = cur;
    int num_remaining = nc;
    for (int it = 0; it < nc; ++it) {
      w.remaining[it] = nc - it - 1;
      w.shortest[it] = kInf;
      w.SC[it] = 0;
      w.path[it] = -1;
    }
    for (int r = 0; r < nr; ++r) w.SR[r] = 0;
    int sink = -1;
    while (sink == -1) {
      int index = -1;
      double lowest = kInf;
      w.SR[i] = 1;
      const double ui = w.u[i];
      for (int it = 0; it < num_remaining; ++it) {
        const int j = w.remaining[it];
        const double r = min_val + cost(i, j) - ui - w.v[j];
        if (r < w.shortest[j]) {
          w.path[j] = i;
          w.shortest[j] = r;
        }
        if (w.shortest[j] < lowest || (w.shortest[j] == lowest && w.row4col[j] == -1)) {
          lowest = w.shortest[j];
          index = it;
        }
      }
      min_val = lowest;
      if (min_val == kInf) return -1;
      const int j = w.remaining[index];
      if (w.row4col[j] == -1)
        sink = j;
      else
        i = w.row4col[j];
      w.SC[j] = 1;
      w.remaining[index] = w.remaining[--num_remaining];
    }
    w.u[cur] += min_val;
    for (int r = 0; r < nr; ++r)
      if (w.SR[r] && r != cur) w.u[r] += min_val - w.shortest[w.col4row[r]];
    for (int j = 0; j < nc; ++j)
      if (w.SC[j]) w.v[j] -= min_val - w.shortest[j];
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
  return 0;
}

}  // namespace toist
