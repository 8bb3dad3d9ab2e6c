// mt_aggregate.cuh -- microscopic -> macroscopic aggregation.
//
// The reference turns SUMO emission records (time, global_position, speed) into
// the macroscopic snapshots the models are fitted on in a notebook loop
// (experiments/plotting_results.ipynb cell 8): vehicle r falls into time row
// floor(time/DT) and section floor(position/DX); per bin the speeds are summed
// and the vehicles counted; density = count/DX, speed = sum/count, and empty
// sections get the free-flow speed (30).  Records whose indices fall outside
// the arrays are dropped (IndexError); indices in [-n, -1] wrap like Python's.
//
// Determinism: the per-bin sum is taken in record order, exactly like the
// reference's sequential loop, so the result is bit-equal to it -- no float
// atomics.  Records must be grouped by time row in non-decreasing order (the
// emission files are; the Python wrapper stable-sorts otherwise); one CTA owns
// a time row, finds its record range by binary search, and each thread walks
// that range for its sections.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mt {

struct AggArgs {
  const double* time;
  const double* pos;
  const double* speed;
  int* t_idx;        // [M] scratch
  int* x_idx;        // [M] scratch
  long long M;
  double dt, dx, empty_speed;
  int n_rows, n_sections;
  double* speed_out;    // [n_rows, n_sections]
  double* density_out;  // [n_rows, n_sections]
  int* unsorted;        // set to 1 if the time rows are not non-decreasing
};

__device__ __forceinline__ int agg_index(double value, double step, int n) {
  const double q = floor(__ddiv_rn(value, step));   // math.floor(float / DT)
  if (!(q >= -(double)n && q < (double)n)) return -1;   // IndexError (or NaN): dropped
  int i = (int)q;
  if (i < 0) i += n;                                    // Python negative index
  return i;
}

__global__ void agg_index_kernel(const AggArgs a) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.M) return;
  const int ti = agg_index(a.time[r], a.dt, a.n_rows);
  const int xi = agg_index(a.pos[r], a.dx, a.n_sections);
  // a record dropped for its time takes the sentinel row n_rows (sorts last);
  // one dropped for its position keeps its row but matches no section
  a.t_idx[r] = (ti < 0) ? a.n_rows : ti;
  a.x_idx[r] = (ti < 0) ? -1 : xi;
}

// rows must be non-decreasing among records with a valid row
__global__ void agg_check_sorted_kernel(const AggArgs a) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r + 1 >= a.M) return;
  if (a.t_idx[r] > a.t_idx[r + 1]) *a.unsorted = 1;
}

__device__ __forceinline__ long long agg_lower_bound(const int* t, long long n, int key) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (t[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void agg_reduce_kernel(const AggArgs a) {
  const int row = blockIdx.x;
  __shared__ long long s_lo, s_hi;
  if (threadIdx.x == 0) {
    s_lo = agg_lower_bound(a.t_idx, a.M, row);
    s_hi = agg_lower_bound(a.t_idx, a.M, row + 1);
  }
  __syncthreads();
  const long long lo = s_lo, hi = s_hi;
  for (int x = threadIdx.x; x < a.n_sections; x += blockDim.x) {
    double sum = 0.0;
    int count = 0;
    for (long long r = lo; r < hi; ++r) {
      if (a.x_idx[r] == x) { sum = __dadd_rn(sum, a.speed[r]); ++count; }
    }
    const size_t o = (size_t)row * a.n_sections + x;
    a.density_out[o] = __ddiv_rn((double)count, a.dx);
    a.speed_out[o] = count == 0 ? a.empty_speed : __ddiv_rn(sum, (double)count);
  }
}

}  // namespace mt
