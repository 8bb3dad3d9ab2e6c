// mt_seg.cuh -- the pipelined step kernel for roads LONGER than one tile
// (config 4: a 10 M-cell highway, whole or as one rank's block of it).
//
// Same producer/consumer skeleton as mt_pipe.cuh.  A tile is a SEGMENT of one
// road: `nc * CPT` consecutive cells.  The producer fetches the segment's
// densities and speeds with two bulk copies (the two halves of a road row are
// N cells apart in HBM) into the stage as [rho segment | v segment], plus the
// 16-byte chunks that hold the one neighbouring cell on each side (the halo),
// all arriving on the stage's mbarrier.  The first / last thread of a segment
// derive the flux ingredients of their halo cell themselves; true road ends
// get the boundary rule.  The result leaves through two bulk stores.
#pragma once
#include <stdint.h>
#include "mtraffic.h"
#include "mt_math.cuh"
#include "mt_cell.cuh"
#include "mt_pipe.cuh"

namespace mt {

template <typename T, int U> struct SegSmem {
  static constexpr int STAGES = (sizeof(T) == 8 && U == 1) ? 3 : 4;
  static constexpr int XCH = PIPE_BLOCK / U + 8;
  alignas(128) unsigned char in[STAGES][PIPE_TILE_BYTES];
  alignas(128) unsigned char out[PIPE_OUT][PIPE_TILE_BYTES];
  alignas(16) unsigned char halo[STAGES][4][16];   // left rho, left v, right rho, right v chunks
  T xS[2][XCH], xD[2][XCH], xR[2][XCH];
  T s_num[PIPE_BLOCK / 32], s_den[PIPE_BLOCK / 32];
  alignas(8) uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t out_full[PIPE_OUT];
  uint64_t out_free[PIPE_OUT];
};
static_assert(sizeof(SegSmem<double, 2>) <= 113 * 1024, "f64 U=2 smem");
static_assert(sizeof(SegSmem<double, 1>) <= 113 * 1024, "f64 U=1 smem");
static_assert(sizeof(SegSmem<float, 1>) <= 113 * 1024, "f32 U=1 smem");

template <typename T> struct SegArgs {
  const T* in;
  T* out;
  const T* table;
  const int* flags;
  const T* action;
  T* partial;    // [B, tiles_per_row, 2] reward partials, or nullptr
  int B;
  int N;
  int tiles_per_row;
  int n_tiles;   // B * tiles_per_row
  int nc;        // consumer threads; a full segment has nc * CPT cells
  BcArgs<T> bc;
  T step;
  UniformC<T> u;  // used by the UNIFORM instantiations
};

template <typename T, int MODEL, bool FAST, int LAMK, bool UNIFORM, int U>
__global__ void __launch_bounds__(PIPE_BLOCK / U + 32, 2) step_seg_kernel(const SegArgs<T> a) {
  constexpr int V = Vec<T>::N;
  constexpr int CPT = U * V;
  constexpr bool ARZ = (MODEL == 1);
  constexpr int STAGES = SegSmem<T, U>::STAGES;
  constexpr int XCH = SegSmem<T, U>::XCH;
  constexpr uint32_t HALF = PIPE_TILE_BYTES / 2;   // v segment offset inside a stage
  using A = Ar<T, FAST>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SegSmem<T, U>& sm = *reinterpret_cast<SegSmem<T, U>*>(smem_raw);

  const int t = threadIdx.x;
  const int NC = a.nc;
  const int N = a.N;
  const int seg_cells = NC * CPT;
  const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t in0 = smem_u32(sm.in[0]), out0 = smem_u32(sm.out[0]);
  const uint32_t halo0 = smem_u32(sm.halo[0][0]);
  const uint32_t full0 = smem_u32(&sm.full[0]), empty0 = smem_u32(&sm.empty[0]);
  const uint32_t ofull0 = smem_u32(&sm.out_full[0]), ofree0 = smem_u32(&sm.out_free[0]);

  pdl_launch_dependents();
  for (int i = t; i < 2 * XCH; i += blockDim.x) { (&sm.xD[0][0])[i] = T(0); (&sm.xR[0][0])[i] = T(0); }
  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], (uint32_t)(NC >> 5));
    }
    for (int o = 0; o < PIPE_OUT; ++o) {
      mbar_init(&sm.out_full[o], (uint32_t)(NC >> 5));
      mbar_init(&sm.out_free[o], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait_prior_grid();

  // tile -> (road, first cell, cells)
  auto locate_tile = [&](int tile, int& env, int& c0, int& cells) {
    env = tile / a.tiles_per_row;
    c0 = (tile - env * a.tiles_per_row) * seg_cells;
    cells = (N - c0 < seg_cells) ? (N - c0) : seg_cells;
  };

  // ======================= producer warp ===================================
  if (t >= NC) {
    if (t == NC) {
      auto issue_load = [&](int i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int s = i % STAGES;
        int env, c0, cells;
        locate_tile(tile, env, c0, cells);
        const T* row = a.in + (size_t)env * 2 * N;
        const uint32_t seg_bytes = (uint32_t)cells * (uint32_t)sizeof(T);
        const bool left = c0 > 0, right = c0 + cells < N;
        const uint32_t bar = full0 + 8 * s;
        uint32_t total = seg_bytes * (ARZ ? 2u : 1u);
        total += (left ? 16u : 0u) * (ARZ ? 2u : 1u) + (right ? 16u : 0u) * (ARZ ? 2u : 1u);
        mbar_expect_tx(bar, total);
        const uint32_t dst = in0 + (uint32_t)s * PIPE_TILE_BYTES;
        const uint32_t hdst = halo0 + (uint32_t)s * 64;
        bulk_load(dst, row + c0, seg_bytes, bar);
        if (ARZ) bulk_load(dst + HALF, row + N + c0, seg_bytes, bar);
        if (left) {
          bulk_load(hdst, row + c0 - V, 16, bar);
          if (ARZ) bulk_load(hdst + 16, row + N + c0 - V, 16, bar);
        }
        if (right) {
          bulk_load(hdst + 32, row + c0 + cells, 16, bar);
          if (ARZ) bulk_load(hdst + 48, row + N + c0 + cells, 16, bar);
        }
      };
      const int pre = n_my < STAGES ? n_my : STAGES;
      for (int i = 0; i < pre; ++i) issue_load(i);
      for (int i = 0; i < n_my; ++i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int s = i % STAGES, o = i % PIPE_OUT;
        if (i + STAGES < n_my) {
          mbar_wait_relaxed(empty0 + 8 * s, (uint32_t)((i / STAGES) & 1), PRODUCER_NAP_NS);
          issue_load(i + STAGES);
        }
        mbar_wait_relaxed(ofull0 + 8 * o, (uint32_t)((i / PIPE_OUT) & 1), PRODUCER_NAP_NS);
        int env, c0, cells;
        locate_tile(tile, env, c0, cells);
        T* orow = a.out + (size_t)env * 2 * N;
        const uint32_t seg_bytes = (uint32_t)cells * (uint32_t)sizeof(T);
        bulk_store(orow + c0, out0 + (uint32_t)o * PIPE_TILE_BYTES, seg_bytes);
        bulk_store(orow + N + c0, out0 + (uint32_t)o * PIPE_TILE_BYTES + HALF, seg_bytes);
        bulk_commit();
        bulk_wait_read<PIPE_OUT - 1>();
        if (i >= PIPE_OUT - 1) mbar_arrive(ofree0 + 8 * ((i - (PIPE_OUT - 1)) % PIPE_OUT));
      }
      bulk_wait_read<0>();
    }
    return;
  }

  // ======================= consumers =======================================
  const bool lane0 = (t & 31) == 0;
  const bool want_reward = a.partial != nullptr;
  const uint32_t my_off = (uint32_t)t * (U * 16);
  const int xi = t;

  EnvC<T> e;
  if (UNIFORM) {
    load_uniform(e, a.u, 0, (const T*)nullptr, a.step);
    if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
  }

  int s = 0, o = 0;
  uint32_t ph_in = 0, ph_out = 0;
  for (int i = 0; i < n_my; ++i) {
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    int env, c0, cells;
    locate_tile(tile, env, c0, cells);
    const int nthr = cells / CPT;
    const bool active = t < nthr;
    const bool seg_first = (t == 0), seg_last = (t == nthr - 1);
    const bool row_first = seg_first && c0 == 0;
    const bool row_last = seg_last && (c0 + cells == N);

    if (active && (!UNIFORM || a.action != nullptr)) {
      if (UNIFORM) load_uniform(e, a.u, env, a.action, a.step);
      else load_env(e, a.table, a.flags, env, a.action, a.step);
      if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
    }

    mbar_wait(full0 + 8 * s, ph_in);

    T rho[CPT], v[CPT], y[CPT], r[CPT], D[CPT], S[CPT];
    T loop_a = 0, loop_b = 0;
    T hl_rho = 0, hl_v = 0, hr_rho = 0, hr_v = 0;   // halo cells (segment-edge threads only)
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
          lds16(src + HALF + 16 * u, tmp);
#pragma unroll
          for (int k = 0; k < V; ++k) v[u * V + k] = tmp[k];
        }
      }
      const T* hal = reinterpret_cast<const T*>(sm.halo[s][0]);
      constexpr int HV = 16 / (int)sizeof(T);   // values per halo chunk
      if (seg_first && !row_first) { hl_rho = hal[HV - 1]; if (ARZ) hl_v = hal[2 * HV - 1]; }
      if (seg_last && !row_last) { hr_rho = hal[2 * HV]; if (ARZ) hr_v = hal[3 * HV]; }
      if (row_first && a.bc.bc_l == MT_BC_LOOP) {   // old state of the road's far end
        const T* row = a.in + (size_t)env * 2 * N;
        if (ARZ) { loop_a = __ldg(row + N - 1); loop_b = __ldg(row + 2 * N - 1); }
        else { loop_a = __ldg(row + N - 2); loop_b = __ldg(row + N - 1); }
      }
      with_lam<LAMK>(e.flags, [&](auto kind) {
        constexpr int K = decltype(kind)::value;
        if (ARZ) arz_pass1<T, FAST, K, CPT>(e, rho, v, y, r, D, S);
        else lwr_pass1<T, FAST, K, CPT>(e, rho, D, S);
      });
      if (ARZ) xR[xi + 1] = r[CPT - 1];
      xS[xi] = S[0];
      xD[xi + 1] = D[CPT - 1];
      if (row_last) xS[xi + 1] = S[CPT - 1];
    }
    T rn[CPT], vn[CPT], Q[CPT], P[CPT];
    if (active) {  // interior cells need nothing from other threads: finish them first
      with_lam<LAMK>(e.flags, [&](auto kind) {
        cell_mid<T, ARZ, FAST, decltype(kind)::value, CPT>(e, rho, y, r, D, S, Q, P, rn, vn);
      });
    }
    consumer_sync(NC);
    if (lane0) mbar_arrive(empty0 + 8 * s);
    mbar_wait(ofree0 + 8 * o, ph_out ^ 1);

    T num = 0, den = 0;
    if (active) {
      T S_right = xS[xi + 1], D_left = xD[xi], r_left = xR[xi];
      with_lam<LAMK>(e.flags, [&](auto kind) {
        constexpr int K = decltype(kind)::value;
        if (seg_last && !row_last) {          // supply of the next segment's first cell
          T yy, dd;
          if (ARZ) arz_cell_nodiv<T, FAST, K>(e, hr_rho, hr_v, yy, dd, S_right);
          else lwr_cell<T, FAST, K>(e, hr_rho, dd, S_right);
        }
        if (seg_first && !row_first) {        // demand and y/rho of the previous segment's last cell
          T yy, ss;
          if (ARZ) arz_cell<T, FAST, K>(e, hl_rho, hl_v, yy, r_left, D_left, ss);
          else lwr_cell<T, FAST, K>(e, hl_rho, D_left, ss);
        }
        cell_edges<T, ARZ, FAST, K, CPT>(e, a.bc, row_first, row_last, rho, v, y, r, D, S, Q, P,
                                         S_right, D_left, r_left, loop_a, loop_b, rn, vn);
      });
      if (want_reward) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) { num += rn[k] * vn[k]; den += rn[k]; }
      }
      const uint32_t dst = out0 + (uint32_t)o * PIPE_TILE_BYTES + my_off;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T tr[V], tv[V];
#pragma unroll
        for (int k = 0; k < V; ++k) { tr[k] = rn[u * V + k]; tv[k] = vn[u * V + k]; }
        sts16(dst + 16 * u, tr);
        sts16(dst + HALF + 16 * u, tv);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane0) mbar_arrive(ofull0 + 8 * o);

    if (want_reward) {   // one (num, den) partial per segment; summed by reward_finalize_kernel
      for (int off = 16; off > 0; off >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, off);
        den += __shfl_xor_sync(0xffffffffu, den, off);
      }
      if (lane0) { sm.s_num[t >> 5] = num; sm.s_den[t >> 5] = den; }
      consumer_sync(NC);
      if (t == 0) {
        const int nw = NC >> 5;
        T n = sm.s_num[0], d = sm.s_den[0];
        for (int j = 1; j < nw; ++j) { n += sm.s_num[j]; d += sm.s_den[j]; }
        T* p = a.partial + (size_t)tile * 2;
        p[0] = n;
        p[1] = d;
      }
    }
    if (++s == STAGES) { s = 0; ph_in ^= 1; }
    if (++o == PIPE_OUT) { o = 0; ph_out ^= 1; }
  }
}

}  // namespace mt
