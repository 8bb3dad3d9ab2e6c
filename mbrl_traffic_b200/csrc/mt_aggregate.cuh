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
// emission files are; the Python wrapper stable-sorts otherwise).  One WARP owns
// a time row: it finds the row's record range by binary search and walks it 32
// records at a time; lanes whose records fall into the same section are found
// with match.any, and the lowest of them adds the group's speeds to the
// section's running sum in lane order -- every record is read once, every
// section's sum is formed in record order (agg_reduce_kernel).  Rows with more
// sections than fit the warp's shared-memory bins take the older kernel, in
// which every thread walks the whole range for its sections
// (agg_reduce_scan_kernel).
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

// One warp per time row; dynamic shared memory: per warp n_sections doubles
// (sums), n_sections ints (counts), 32 doubles (the chunk's speeds).
__global__ void agg_reduce_kernel(const AggArgs a) {
  extern __shared__ __align__(8) unsigned char agg_smem[];
  const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * wpb + w;
  const size_t per_warp = (size_t)a.n_sections * 12 + 32 * 8;
  unsigned char* base = agg_smem + (size_t)w * ((per_warp + 7) & ~(size_t)7);
  double* sum = reinterpret_cast<double*>(base);
  double* chunk = sum + a.n_sections;
  int* cnt = reinterpret_cast<int*>(chunk + 32);
  if (row >= a.n_rows) return;
  for (int x = lane; x < a.n_sections; x += 32) { sum[x] = 0.0; cnt[x] = 0; }
  long long lo = 0, hi = 0;
  if (lane == 0) {
    lo = agg_lower_bound(a.t_idx, a.M, row);
    hi = agg_lower_bound(a.t_idx, a.M, row + 1);
  }
  lo = __shfl_sync(0xffffffffu, lo, 0);
  hi = __shfl_sync(0xffffffffu, hi, 0);
  __syncwarp();
  for (long long b = lo; b < hi; b += 32) {
    const long long r = b + lane;
    const int xi = r < hi ? a.x_idx[r] : -1;
    chunk[lane] = r < hi ? a.speed[r] : 0.0;
    const unsigned peers = __match_any_sync(0xffffffffu, xi);
    __syncwarp();
    if (xi >= 0 && lane == __ffs(peers) - 1) {      // lowest lane of the section's group
      double s = sum[xi];
      int c = cnt[xi];
      for (unsigned m = peers; m != 0; m &= m - 1) {  // its records, in record order
        s = __dadd_rn(s, chunk[__ffs(m) - 1]);
        ++c;
      }
      sum[xi] = s;
      cnt[xi] = c;
    }
    __syncwarp();
  }
  for (int x = lane; x < a.n_sections; x += 32) {
    const size_t o = (size_t)row * a.n_sections + x;
    const int c = cnt[x];
    a.density_out[o] = __ddiv_rn((double)c, a.dx);
    a.speed_out[o] = c == 0 ? a.empty_speed : __ddiv_rn(sum[x], (double)c);
  }
}

__global__ void agg_reduce_scan_kernel(const AggArgs a) {
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
