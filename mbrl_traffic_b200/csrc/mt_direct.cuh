// mt_direct.cuh -- the direct step kernels: one CTA per road tile, global
// loads straight into registers.  They serve
//   * rows longer than one pipeline tile (a road split over several CTAs, each
//     re-deriving its two halo cells from global memory),
//   * any N that is not a multiple of the vector width (scalar kernels, one
//     thread per cell), which double as an independent cross-check of the
//     vector arithmetic in the tests,
//   * the A/B baseline of the TMA-pipelined kernel (MT_FLAG_NO_PIPELINE).
// All of them evaluate exactly the per-cell functions of mt_cell.cuh, so every
// kernel produces bit-identical results.
#pragma once
#include "mtraffic.h"
#include "mt_cell.cuh"

namespace mt {

constexpr int MAX_BLOCK = 512;
constexpr int DIRECT_XCH = MAX_BLOCK + MAX_BLOCK / 2 + 8;

template <typename T> struct StepArgs {
  const T* in;
  T* out;
  const T* table;
  const int* flags;
  const T* action;    // [B] or nullptr
  T* reward;          // [B] or nullptr
  T* partial;         // [B, tiles, 2] scratch when tiles > 1
  long long B;
  int N;
  int nvec;           // N / V (vector kernels)
  int tpe;            // threads per env (padded) when tiles == 1
  int epc;            // envs per CTA when tiles == 1
  int tiles;          // CTAs per env
  BcArgs<T> bc;
  T step;
};

// thread -> (env, vector) mapping of the vector kernels
struct Where {
  long long env;
  int vec;
  int tile;
  int xi;           // exchange slot
  bool active;
  bool tile_first;  // first thread of a tile of a multi-tile row
  bool tile_last;
};

template <typename T>
__device__ __forceinline__ Where locate(const StepArgs<T>& a) {
  Where w;
  const int t = threadIdx.x;
  if (a.tiles == 1) {
    const int el = t / a.tpe;
    w.vec = t - el * a.tpe;
    w.env = (long long)blockIdx.x * a.epc + el;
    w.tile = 0;
    w.xi = t + el;
    w.active = (el < a.epc) && (w.env < a.B) && (w.vec < a.nvec);
    w.tile_first = false;
    w.tile_last = false;
  } else {
    w.env = blockIdx.x / a.tiles;
    w.tile = blockIdx.x - (int)(w.env * a.tiles);
    w.vec = w.tile * blockDim.x + t;
    w.xi = t;
    w.active = w.vec < a.nvec;
    w.tile_first = (t == 0);
    w.tile_last = (t == (int)blockDim.x - 1);
  }
  return w;
}

// Per-road reward sum(rho v) / sum(rho), deterministic order.
// tpe <= 32: segmented butterfly inside the warp.  Otherwise warp butterfly,
// one partial per warp in shared memory, then the road's first thread adds
// its warps in order.  Multi-tile rows write one partial per tile.
template <typename T>
__device__ __forceinline__ void reduce_reward(const StepArgs<T>& a, const Where& w,
                                              T num, T den, T* s_num, T* s_den) {
  const unsigned full = 0xffffffffu;
  const int t = threadIdx.x;
  if (a.tiles == 1 && a.tpe <= 32) {
    for (int off = a.tpe >> 1; off > 0; off >>= 1) {
      num += __shfl_xor_sync(full, num, off);
      den += __shfl_xor_sync(full, den, off);
    }
    if (w.active && w.vec == 0) a.reward[w.env] = num / den;
    return;
  }
  for (int off = 16; off > 0; off >>= 1) {
    num += __shfl_xor_sync(full, num, off);
    den += __shfl_xor_sync(full, den, off);
  }
  if ((t & 31) == 0) { s_num[t >> 5] = num; s_den[t >> 5] = den; }
  __syncthreads();
  if (a.tiles == 1) {
    if (w.active && w.vec == 0) {
      const int w0 = t >> 5, nw = a.tpe >> 5;
      T n = s_num[w0], d = s_den[w0];
      for (int i = 1; i < nw; ++i) { n += s_num[w0 + i]; d += s_den[w0 + i]; }
      a.reward[w.env] = n / d;
    }
  } else if (t == 0) {
    const int nw = blockDim.x >> 5;
    T n = s_num[0], d = s_den[0];
    for (int i = 1; i < nw; ++i) { n += s_num[i]; d += s_den[i]; }
    T* p = a.partial + ((long long)w.env * a.tiles + w.tile) * 2;
    p[0] = n;
    p[1] = d;
  }
}

// one warp per env: fixed-order sum of the per-tile partials
template <typename T>
__global__ void reward_finalize_kernel(const T* __restrict__ partial, T* __restrict__ reward,
                                       long long B, int tiles) {
  const long long env = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= B) return;
  const int lane = threadIdx.x & 31;
  T n = 0, d = 0;
  for (int i = lane; i < tiles; i += 32) {
    n += partial[(env * tiles + i) * 2];
    d += partial[(env * tiles + i) * 2 + 1];
  }
  for (int off = 16; off > 0; off >>= 1) {
    n += __shfl_xor_sync(0xffffffffu, n, off);
    d += __shfl_xor_sync(0xffffffffu, d, off);
  }
  if (lane == 0) reward[env] = n / d;
}

// ---------------------------------------------------------------------------
// vector kernel: MODEL 0 = LWR (lwr.py:192-266), 1 = ARZ (arz.py:205-221)
// ---------------------------------------------------------------------------
template <typename T, int MODEL, bool FAST>
__global__ void __launch_bounds__(MAX_BLOCK, 2) step_vec_kernel(const StepArgs<T> a) {
  constexpr int V = Vec<T>::N;
  constexpr bool ARZ = (MODEL == 1);
  constexpr int LAMK = LAM_RUNTIME;
  using VT = typename Vec<T>::type;
  using A = Ar<T, FAST>;
  __shared__ T xS[DIRECT_XCH], xD[DIRECT_XCH], xR[DIRECT_XCH];
  __shared__ T s_num[MAX_BLOCK / 32], s_den[MAX_BLOCK / 32];

  const int t = threadIdx.x;
  const Where w = locate(a);
  const int N = a.N;
  const bool want_reward = a.reward != nullptr;

  // pad slots feed the (discarded) left-face flux of each road's first cell
  for (int i = t; i < DIRECT_XCH; i += blockDim.x) { xD[i] = T(0); xR[i] = T(0); }
  __syncthreads();

  EnvC<T> e;
  T rho[V], v[V], y[V], r[V], D[V], S[V];
  const T* row = nullptr;
  bool row_first = false, row_last = false;
  if (w.active) {
    row_first = (w.vec == 0);
    row_last = (w.vec == a.nvec - 1);
    load_env(e, a.table, a.flags, w.env, a.action, a.step);
    row = a.in + w.env * (2LL * N);
    unpack(__ldg(reinterpret_cast<const VT*>(row) + w.vec), rho);
    if (ARZ) {
      e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
      unpack(__ldg(reinterpret_cast<const VT*>(row + N) + w.vec), v);
    }
    with_lam<LAMK>(e.flags, [&](auto kind) {
      constexpr int K = decltype(kind)::value;
      if (ARZ) arz_pass1<T, FAST, K, V>(e, rho, v, y, r, D, S);
      else lwr_pass1<T, FAST, K, V>(e, rho, D, S);
    });
    if (ARZ) xR[w.xi + 1] = r[V - 1];
    xS[w.xi] = S[0];
    xD[w.xi + 1] = D[V - 1];
    if (row_last) xS[w.xi + 1] = S[V - 1];
  }
  __syncthreads();

  T num = 0, den = 0;
  if (w.active) {
    T S_right = xS[w.xi + 1], D_left = xD[w.xi], r_left = xR[w.xi];
    T loop_a = 0, loop_b = 0;
    if (row_first && a.bc.bc_l == MT_BC_LOOP) {
      if (ARZ) { loop_a = __ldg(row + N - 1); loop_b = __ldg(row + 2 * N - 1); }
      else { loop_a = __ldg(row + N - 2); loop_b = __ldg(row + N - 1); }
    }
    T rn[V], vn[V];
    with_lam<LAMK>(e.flags, [&](auto kind) {
      constexpr int K = decltype(kind)::value;
      // tile edges of a multi-tile row: re-derive the halo cell from global memory
      if (w.tile_last && !row_last) {
        const int j = (w.vec + 1) * V;
        T yy, rr, dd;
        if (ARZ) arz_cell<T, FAST, K>(e, __ldg(row + j), __ldg(row + N + j), yy, rr, dd, S_right);
        else lwr_cell<T, FAST, K>(e, __ldg(row + j), dd, S_right);
      }
      if (w.tile_first && !row_first) {
        const int j = w.vec * V - 1;
        T yy, ss;
        if (ARZ) arz_cell<T, FAST, K>(e, __ldg(row + j), __ldg(row + N + j), yy, r_left, D_left, ss);
        else lwr_cell<T, FAST, K>(e, __ldg(row + j), D_left, ss);
      }
      if (ARZ)
        arz_pass2<T, FAST, K, V>(e, a.bc, row_first, row_last, rho, v, y, r, D, S, S_right,
                                 D_left, r_left, loop_a, loop_b, rn, vn);
      else
        lwr_pass2<T, FAST, K, V>(e, a.bc, row_first, row_last, rho, D, S, S_right, D_left,
                                 loop_a, loop_b, rn, vn);
    });
    if (want_reward) {
#pragma unroll
      for (int k = 0; k < V; ++k) { num += rn[k] * vn[k]; den += rn[k]; }
    }
    T* orow = a.out + w.env * (2LL * N);
    reinterpret_cast<VT*>(orow)[w.vec] = pack(rn);
    reinterpret_cast<VT*>(orow + N)[w.vec] = pack(vn);
  }
  if (want_reward) reduce_reward(a, w, num, den, s_num, s_den);
}

// ---------------------------------------------------------------------------
// scalar kernels: any N >= 3, one thread per cell, neighbours re-derived from
// global memory.
// ---------------------------------------------------------------------------
template <typename T, bool FAST>
__device__ __forceinline__ void arz_interior(const EnvC<T>& e, const T* __restrict__ row, int N,
                                             int j, T& rho_n, T& y_n) {
  using A = Ar<T, FAST>;
  constexpr int LAMK = LAM_RUNTIME;
  T yL, rL, DL, SL, yC, rC, DC, SC, yR, rR, DR, SR;
  arz_cell<T, FAST, LAMK>(e, __ldg(row + j - 1), __ldg(row + N + j - 1), yL, rL, DL, SL);
  arz_cell<T, FAST, LAMK>(e, __ldg(row + j), __ldg(row + N + j), yC, rC, DC, SC);
  arz_cell<T, FAST, LAMK>(e, __ldg(row + j + 1), __ldg(row + N + j + 1), yR, rR, DR, SR);
  const T Q = npmin(DC, SR), Qm = npmin(DL, SC);
  const T P = A::mul(Q, rC), Pm = A::mul(Qm, rL);
  arz_advance<T, FAST>(e, __ldg(row + j), yC, Qm, Q, Pm, P, rho_n, y_n);
}

template <typename T, bool FAST>
__global__ void arz_step_scalar_kernel(const StepArgs<T> a) {
  using A = Ar<T, FAST>;
  constexpr int LAMK = LAM_RUNTIME;
  const int N = a.N;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= a.B * N) return;
  const long long env = gid / N;
  const int i = (int)(gid - env * N);
  EnvC<T> e;
  load_env(e, a.table, a.flags, env, a.action, a.step);
  e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
  const T* row = a.in + env * (2LL * N);
  T* orow = a.out + env * (2LL * N);
  T rn, yn, vn;
  const int bc = (i == 0) ? a.bc.bc_l : a.bc.bc_r;
  if (i == 0 || i == N - 1) {
    if (bc == MT_BC_HALO) {  // ghost cell: passed through
      orow[i] = row[i];
      orow[N + i] = row[N + i];
      return;
    }
    if (bc == MT_BC_LOOP) {
      const int src = (i == 0) ? N - 1 : N - 2;   // OLD state (arz.py:317-318)
      rn = row[src];
      yn = arz_y<T, FAST, LAMK>(e, rn, row[N + src]);
    } else if (bc == MT_BC_EXTEND) {
      arz_interior<T, FAST>(e, row, N, (i == 0) ? 1 : N - 2, rn, yn);
    } else {
      rn = (i == 0) ? a.bc.bl0 : a.bc.br0;
      yn = (i == 0) ? a.bc.bl1 : a.bc.br1;
    }
  } else {
    arz_interior<T, FAST>(e, row, N, i, rn, yn);
  }
  arz_u<T, FAST, LAMK>(e, rn, yn, vn);
  orow[i] = rn;
  orow[N + i] = vn;
}

// updated LWR density of cell j under the scheme alone (no boundary rule)
template <typename T, bool FAST>
__device__ __forceinline__ T lwr_scheme(const EnvC<T>& e, const T* __restrict__ row, int N, int j) {
  constexpr int LAMK = LAM_RUNTIME;
  T DC, SC, DL = 0, SL, DR, SR;
  const T rc = __ldg(row + j);
  lwr_cell<T, FAST, LAMK>(e, rc, DC, SC);
  if (j + 1 < N) lwr_cell<T, FAST, LAMK>(e, __ldg(row + j + 1), DR, SR); else SR = SC;
  const T f = npmin(DC, SR);
  T fm = f;
  if (j > 0) { lwr_cell<T, FAST, LAMK>(e, __ldg(row + j - 1), DL, SL); fm = npmin(DL, SC); }
  return lwr_advance<T, FAST>(e, rc, f, fm);
}

template <typename T, bool FAST>
__global__ void lwr_step_scalar_kernel(const StepArgs<T> a) {
  const int N = a.N;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= a.B * N) return;
  const long long env = gid / N;
  const int i = (int)(gid - env * N);
  EnvC<T> e;
  load_env(e, a.table, a.flags, env, a.action, a.step);
  const T* row = a.in + env * (2LL * N);
  T rn;
  if (i == 0) {
    if (a.bc.bc_l == MT_BC_LOOP) rn = lwr_scheme<T, FAST>(e, row, N, N - 1);
    else if (a.bc.bc_l == MT_BC_EXTEND) rn = lwr_scheme<T, FAST>(e, row, N, 0);
    else if (a.bc.bc_l == MT_BC_CONSTANT) rn = a.bc.bl0;
    else rn = row[0];
  } else if (i == N - 1) {
    if (a.bc.bc_r == MT_BC_LOOP) rn = lwr_scheme<T, FAST>(e, row, N, N - 2);
    else if (a.bc.bc_r == MT_BC_EXTEND) rn = lwr_scheme<T, FAST>(e, row, N, N - 1);
    else if (a.bc.bc_r == MT_BC_CONSTANT) rn = a.bc.br0;
    else rn = row[N - 1];
  } else {
    rn = lwr_scheme<T, FAST>(e, row, N, i);
  }
  T* orow = a.out + env * (2LL * N);
  orow[i] = rn;
  orow[N + i] = v_eq<T, FAST, LAM_RUNTIME>(rn, e);
}

// reward for the scalar path: one warp per env over the finished output
template <typename T>
__global__ void reward_rows_kernel(const T* __restrict__ obs, T* __restrict__ reward,
                                   long long B, int N) {
  const long long env = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= B) return;
  const int lane = threadIdx.x & 31;
  const T* row = obs + env * (2LL * N);
  T n = 0, d = 0;
  for (int i = lane; i < N; i += 32) { n += row[i] * row[N + i]; d += row[i]; }
  for (int off = 16; off > 0; off >>= 1) {
    n += __shfl_xor_sync(0xffffffffu, n, off);
    d += __shfl_xor_sync(0xffffffffu, d, off);
  }
  if (lane == 0) reward[env] = n / d;
}

}  // namespace mt
