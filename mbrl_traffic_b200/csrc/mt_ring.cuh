// mt_ring.cuh -- persistent TMA-fed step kernel without a producer warp.
//
// Same tiling as mt_pipe.cuh (a tile = consecutive roads, <= 16 KB, fetched by
// ONE cp.async.bulk into a ring of shared-memory stages guarded by mbarriers),
// but leaner around the compute:
//   * no producer warp: after the one CTA barrier of a tile (edge exchange,
//     which also proves that every thread has read the stage) thread 0 re-arms
//     that stage with the load of tile i+STAGES -- nothing ever polls;
//   * results go straight from registers to global memory with 16-byte stores
//     (a warp writes 1 KB of contiguous densities and 1 KB of contiguous
//     speeds), so there is no output staging, no proxy fence and no
//     store hand-off; the shared memory saved buys two more input stages
//     (5 tiles = 80 KB in flight per CTA, 160 KB per SM).
#pragma once
#include <stdint.h>
#include "mtraffic.h"
#include "mt_math.cuh"
#include "mt_cell.cuh"
#include "mt_pipe.cuh"

namespace mt {

template <typename T, int U> struct RingCfg {
  static constexpr int NC = PIPE_BLOCK / U;                 // threads per CTA
  static constexpr int XCH = NC + PIPE_MAX_EPC + 8;         // exchange slots (one pad per road)
  static constexpr int XCH_BYTES = 2 * 3 * XCH * (int)sizeof(T);
  // two CTAs per SM: 2 x (smem + 1 KB reserved) <= 228 KB
  static constexpr int STAGES = (113 * 1024 - XCH_BYTES - 1024) / PIPE_TILE_BYTES;
};

template <typename T, int U> struct RingSmem {
  static constexpr int STAGES = RingCfg<T, U>::STAGES;
  static constexpr int XCH = RingCfg<T, U>::XCH;
  alignas(128) unsigned char in[STAGES][PIPE_TILE_BYTES];
  T xS[2][XCH], xD[2][XCH], xR[2][XCH];   // edge exchange, double-buffered by tile parity
  T s_num[PIPE_BLOCK / 32], s_den[PIPE_BLOCK / 32];
  alignas(8) uint64_t full[STAGES];
};
static_assert(sizeof(RingSmem<double, 1>) <= 113 * 1024, "f64 U=1 smem");
static_assert(sizeof(RingSmem<double, 2>) <= 113 * 1024, "f64 U=2 smem");
static_assert(sizeof(RingSmem<float, 1>) <= 113 * 1024, "f32 U=1 smem");
static_assert(sizeof(RingSmem<float, 2>) <= 113 * 1024, "f32 U=2 smem");
static_assert(RingCfg<double, 2>::STAGES >= 5, "expected at least 5 stages for f64 U=2");

__device__ __forceinline__ void stg16(double* p, const double (&a)[2]) {
  *reinterpret_cast<double2*>(p) = make_double2(a[0], a[1]);
}
__device__ __forceinline__ void stg16(float* p, const float (&a)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
}

// Same template parameters and argument struct as step_pipe_kernel.
template <typename T, int MODEL, bool FAST, int LAMK, bool UNIFORM, int U>
__global__ void __launch_bounds__(PIPE_BLOCK / U, 2) step_ring_kernel(const PipeArgs<T> a) {
  constexpr int V = Vec<T>::N;
  constexpr int CPT = U * V;
  constexpr bool ARZ = (MODEL == 1);
  constexpr int STAGES = RingCfg<T, U>::STAGES;
  constexpr int XCH = RingCfg<T, U>::XCH;
  using A = Ar<T, FAST>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RingSmem<T, U>& sm = *reinterpret_cast<RingSmem<T, U>*>(smem_raw);

  const int t = threadIdx.x;
  const int N = a.N;
  const uint32_t row_bytes = 2u * (uint32_t)N * (uint32_t)sizeof(T);
  const uint32_t tile_bytes = (uint32_t)a.epc * row_bytes;
  const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int last_tile = a.n_tiles - 1;
  const int last_envs = a.B - last_tile * a.epc;  // roads in the (ragged) last tile
  const uint32_t in0 = smem_u32(sm.in[0]);
  const uint32_t full0 = smem_u32(&sm.full[0]);
  const unsigned char* gin = reinterpret_cast<const unsigned char*>(a.in);

  auto issue_load = [&](int i) {  // thread 0 only
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int s = i % STAGES;
    const int envs = tile == last_tile ? last_envs : a.epc;
    if (ARZ) {
      const uint32_t bytes = (uint32_t)envs * row_bytes;
      mbar_expect_tx(full0 + 8 * s, bytes);
      bulk_load(in0 + (uint32_t)s * PIPE_TILE_BYTES, gin + (size_t)tile * tile_bytes, bytes,
                full0 + 8 * s);
    } else {
      // LWR ignores the input velocity (lwr.py:197): density half of each road only
      const uint32_t half = row_bytes >> 1;
      mbar_expect_tx(full0 + 8 * s, (uint32_t)envs * half);
      for (int r = 0; r < envs; ++r)
        bulk_load(in0 + (uint32_t)s * PIPE_TILE_BYTES + (uint32_t)r * row_bytes,
                  gin + (size_t)tile * tile_bytes + (size_t)r * row_bytes, half, full0 + 8 * s);
    }
  };
  auto prefetch = [&](int i) {    // thread 0 only: warm L2 beyond the ring
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int envs = tile == last_tile ? last_envs : a.epc;
    if (ARZ) {
      bulk_prefetch_l2(gin + (size_t)tile * tile_bytes, (uint32_t)envs * row_bytes);
    } else {
      for (int r = 0; r < envs; ++r)
        bulk_prefetch_l2(gin + (size_t)tile * tile_bytes + (size_t)r * row_bytes, row_bytes >> 1);
    }
  };

  pdl_launch_dependents();
  // pad slots of the exchange arrays stay zero for the kernel's life: they
  // feed the (discarded) left-face flux of each road's first cell
  for (int i = t; i < 2 * XCH; i += blockDim.x) { (&sm.xD[0][0])[i] = T(0); (&sm.xR[0][0])[i] = T(0); }
  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&sm.full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  // towards L2 before waiting for the previous launch (see step_pipe_kernel:
  // a prefetch returns no data, so this is safe for dependent launches too)
  if (t == 0)
    for (int i = 0; i < a.early_l2 && i < n_my; ++i) prefetch(i);
  pdl_wait_prior_grid();  // everything above overlapped the previous kernel's tail

  const int ahead = a.l2_ahead;
  if (t == 0) {
    const int pre = n_my < STAGES ? n_my : STAGES;
    for (int i = 0; i < pre; ++i) issue_load(i);
    for (int i = STAGES; i < STAGES + ahead && i < n_my; ++i) prefetch(i);
  }

  // thread -> (road in tile, position in road); fixed for the kernel's life
  const int el = t / a.tpe;
  const int pos = t - el * a.tpe;
  const bool row_first = (pos == 0), row_last = (pos == a.nthr - 1);
  const bool owns_cells = (pos < a.nthr) && (el < a.epc);
  const int xi = t + el;  // exchange slot (one pad slot per road)
  const bool want_reward = a.reward != nullptr;
  const uint32_t my_off = (uint32_t)el * row_bytes + (uint32_t)pos * (U * 16);
  const uint32_t v_off = (uint32_t)N * (uint32_t)sizeof(T);
  const uint32_t row0_off = (uint32_t)el * row_bytes;

  EnvC<T> e;
  if (UNIFORM) {
    load_uniform(e, a.u, 0, (const T*)nullptr, a.step);
    if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
  }

  int s = 0;
  uint32_t ph = 0;
  for (int i = 0; i < n_my; ++i) {
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int envs_here = tile == last_tile ? last_envs : a.epc;
    const int env = tile * a.epc + el;
    const bool active = owns_cells && (el < envs_here);

    if (active && (!UNIFORM || a.action != nullptr)) {
      if (UNIFORM) load_uniform(e, a.u, env, a.action, a.step);
      else load_env(e, a.table, a.flags, env, a.action, a.step);
      if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
    }

    mbar_wait(full0 + 8 * s, ph);

    T rho[CPT], v[CPT], y[CPT], r[CPT], D[CPT], S[CPT];
    T loop_a = 0, loop_b = 0;  // old boundary-source cells for the LOOP rule
    T* const xS = sm.xS[i & 1];
    T* const xD = sm.xD[i & 1];
    T* const xR = sm.xR[i & 1];
    if (active) {
      const uint32_t src = in0 + (uint32_t)s * PIPE_TILE_BYTES + my_off;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T tmp[V];
        lds16(src + 16 * u, tmp);
#pragma unroll
        for (int k = 0; k < V; ++k) rho[u * V + k] = tmp[k];
        if (ARZ) {
          lds16(src + v_off + 16 * u, tmp);
#pragma unroll
          for (int k = 0; k < V; ++k) v[u * V + k] = tmp[k];
        }
      }
      if (row_first && a.bc.bc_l == MT_BC_LOOP) {
        const T* row = reinterpret_cast<const T*>(sm.in[s] + row0_off);
        if (ARZ) { loop_a = row[N - 1]; loop_b = row[2 * N - 1]; }
        else { loop_a = row[N - 2]; loop_b = row[N - 1]; }
      }
      with_lam<LAMK>(e.flags, [&](auto kind) {
        constexpr int K = decltype(kind)::value;
        if (ARZ) arz_pass1<T, FAST, K, CPT>(e, rho, v, y, r, D, S);
        else lwr_pass1<T, FAST, K, CPT>(e, rho, D, S);
      });
      if (ARZ) xR[xi + 1] = r[CPT - 1];
      xS[xi] = S[0];
      xD[xi + 1] = D[CPT - 1];
      if (row_last) xS[xi + 1] = S[CPT - 1];  // arz.py:405 / lwr.py:294
    }
    __syncthreads();  // exchange visible; every thread has read stage s
    if (t == 0) {
      if (i + STAGES < n_my) issue_load(i + STAGES);
      if (ahead > 0 && i + STAGES + ahead < n_my) prefetch(i + STAGES + ahead);
    }

    T num = 0, den = 0;
    if (active) {
      T rn[CPT], vn[CPT];
      const T S_right = xS[xi + 1], D_left = xD[xi], r_left = ARZ ? xR[xi] : T(0);
      with_lam<LAMK>(e.flags, [&](auto kind) {
        constexpr int K = decltype(kind)::value;
        if (ARZ)
          arz_pass2<T, FAST, K, CPT>(e, a.bc, row_first, row_last, rho, v, y, r, D, S, S_right,
                                     D_left, r_left, loop_a, loop_b, rn, vn);
        else
          lwr_pass2<T, FAST, K, CPT>(e, a.bc, row_first, row_last, rho, D, S, S_right, D_left,
                                     loop_a, loop_b, rn, vn);
      });
      if (want_reward) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) { num += rn[k] * vn[k]; den += rn[k]; }
      }
      T* orow = a.out + (size_t)env * 2 * N + (size_t)pos * CPT;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T tr[V], tv[V];
#pragma unroll
        for (int k = 0; k < V; ++k) { tr[k] = rn[u * V + k]; tv[k] = vn[u * V + k]; }
        stg16(orow + u * V, tr);
        stg16(orow + N + u * V, tv);
      }
    }

    if (want_reward) {
      // per-road sum(rho v)/sum(rho) in a fixed order
      const unsigned full = 0xffffffffu;
      if (a.tpe <= 32) {
        for (int off = a.tpe >> 1; off > 0; off >>= 1) {
          num += __shfl_xor_sync(full, num, off);
          den += __shfl_xor_sync(full, den, off);
        }
        if (active && row_first) a.reward[env] = num / den;
      } else {
        for (int off = 16; off > 0; off >>= 1) {
          num += __shfl_xor_sync(full, num, off);
          den += __shfl_xor_sync(full, den, off);
        }
        if ((t & 31) == 0) { sm.s_num[t >> 5] = num; sm.s_den[t >> 5] = den; }
        __syncthreads();
        if (active && row_first) {
          const int w0 = t >> 5, nw = a.tpe >> 5;
          T n = sm.s_num[w0], d = sm.s_den[w0];
          for (int j = 1; j < nw; ++j) { n += sm.s_num[w0 + j]; d += sm.s_den[w0 + j]; }
          a.reward[env] = n / d;
        }
      }
    }
    if (++s == STAGES) { s = 0; ph ^= 1; }
  }
}

}  // namespace mt
