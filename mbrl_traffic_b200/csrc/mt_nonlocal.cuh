// mt_nonlocal.cuh -- step kernels of the non-local look-ahead LWR model.
//
// The reference never implemented this model; the
// scheme is the Godunov-type scheme of Friedrich, Kolb, Goettlich, NHM 13
// (2018) -- the paper cited at lwr.py:209-211 -- with g(rho) = rho:
//
//     W_j       = sum_{k=0}^{n_eta-1} gamma_k V(rho_{j+k+1})     (k ascending)
//     F_{j+1/2} = rho_j W_j
//     rho'_j    = rho_j - (dt/dx) (F_{j+1/2} - F_{j-1/2}),   v'_j = V(rho'_j)
//
// exactly as oracle/np_oracle.py nonlocal_step evaluates it: same operation
// order; the look-ahead sum accumulates with ONE fused multiply-add per term
// (w = fma(gamma_k, V, w)) -- no reference code exists for this model, so the
// rounding of that sum is ours to define, and the fused form is both the more
// accurate one and half the FP64 instructions of the kernel's inner loop (the
// oracle emulates it exactly, np_oracle.fma).  Everything else is separately
// rounded as in the LWR / ARZ kernels.
//
// Second scheme (LXF = true; mt_params.scheme = MT_SCHEME_LXF): the model of
// Blandin & Goatin, Numer. Math. 132 (2016), whose speed depends on a mean of
// the downstream DENSITIES, with a Lax-Friedrichs flux:
//
//     R_j       = sum_{k=0}^{n_eta-1} gamma_k rho_{j+k}          (k ascending)
//     g_j       = rho_j V(R_j)
//     F_{j+1/2} = 0.5 (g_j + g_{j+1}) + (alpha/2) (rho_j - rho_{j+1})
//     rho'_j    = rho_j - (dt/dx) (F_{j+1/2} - F_{j-1/2})
//
// as oracle/np_oracle.py nonlocal_lxf_step evaluates it.  Same kernel: the
// shared array then holds the densities themselves, shifted by one slot (slot
// 0 = the upstream ghost), and a thread accumulates R for its cells and BOTH
// neighbours (CPT + 2 windows instead of CPT + 1).  Ghost cells: one upstream, n_eta downstream;
// LOOP = periodic ring, EXTEND = copy of the edge cell, CONSTANT = given value.
//
// Pipelined kernel: same producer/consumer skeleton as mt_pipe.cuh, but only
// the DENSITY half of each road is fetched (one bulk copy per road, all
// arriving on the stage's mbarrier), since the input velocity is ignored.
// Consumers first publish V(rho) of their cells (and of the downstream ghosts)
// in shared memory, then every thread slides the look-ahead window over that
// array for its cells plus the one cell to its left (re-derived instead of a
// second barrier).  The V(rho) array is stored de-interleaved -- cell c of a
// road lives at plane (c mod CPT), column (c / CPT) -- so that the 32 lanes of
// a warp, whose windows are CPT cells apart, read consecutive words (no bank
// conflicts).
#pragma once
#include <stdint.h>
#include "mtraffic.h"
#include "mt_math.cuh"
#include "mt_cell.cuh"
#include "mt_pipe.cuh"

namespace mt {

// CTAs per SM.  3: 72 registers per thread (no spills), two input stages, two
// V(rho) buffers, tiles published and summed in turn (no AHEAD): 27 warps per
// SM instead of 18 for a kernel bound by the latency of its FP64 chains.
#ifndef MT_NL_CTAS
#define MT_NL_CTAS 2
#endif
constexpr int NL_CTAS = MT_NL_CTAS;
constexpr int NL_STAGES = NL_CTAS >= 3 ? 2 : 4;
constexpr int NL_VE_BUFS = NL_CTAS >= 3 ? 2 : 3;
constexpr int NL_IN_BYTES = PIPE_TILE_BYTES / 2;   // densities of one tile (8 KB)
constexpr int NL_MAX_ETA = 256;                    // look-ahead cells
constexpr int NL_GHOSTS = 320;                     // extra V(rho) slots per tile for ghosts

template <typename T> struct NlSmem {
  static constexpr int VE = NL_IN_BYTES / (int)sizeof(T) + NL_GHOSTS + 64;
  alignas(128) unsigned char in[NL_STAGES][NL_IN_BYTES];
  alignas(128) unsigned char out[PIPE_OUT][PIPE_TILE_BYTES];
  alignas(16) T ve[NL_VE_BUFS][VE];        // V(rho) of the tile incl. ghosts: two buffers in turn, or
                                    // three when tiles are published one ahead (AHEAD)
  T gamma[NL_MAX_ETA];
  T s_num[PIPE_BLOCK / 32], s_den[PIPE_BLOCK / 32];
  alignas(8) uint64_t full[NL_STAGES];
  uint64_t empty[NL_STAGES];
  uint64_t out_full[PIPE_OUT];
  uint64_t out_free[PIPE_OUT];
  uint64_t pub[3];                  // AHEAD: every consumer warp has published buffer b
};
static_assert(sizeof(NlSmem<double>) <= (NL_CTAS >= 3 ? 75 : 113) * 1024,
              "NL_CTAS CTAs per SM must fit in 227 KB (1 KB reserved per CTA)");

template <typename T> struct NlArgs {
  const T* in;
  T* out;
  const T* table;
  const int* flags;
  const T* action;
  T* reward;
  const T* gamma;
  int n_eta;
  int B, N;
  int nthr, tpe, epc, n_tiles, nc;
  int ve_cols;     // columns per plane of a road's V(rho) block: nthr + ghost columns
  BcArgs<T> bc;
  T step;
  T half_alpha;    // LXF: alpha / 2, or < 0: half of each road's v_max
  UniformC<T> u;
};

// ghost density upstream of cell 0 / downstream of cell N-1
template <typename T>
__device__ __forceinline__ T nl_left_ghost(const BcArgs<T>& b, const T* rho_row, int N) {
  if (b.bc_l == MT_BC_LOOP) return rho_row[N - 1];
  if (b.bc_l == MT_BC_CONSTANT) return b.bl0;
  return rho_row[0];
}

template <typename T, bool FAST, int LAMK, bool UNIFORM, int U, bool LXF, bool AHEAD = false>
__global__ void __launch_bounds__(PIPE_BLOCK / U + 32, NL_CTAS) nonlocal_pipe_kernel(const NlArgs<T> a) {
  static_assert(!AHEAD || UNIFORM, "publishing a tile ahead needs scalar parameters");
  if constexpr (AHEAD && NL_VE_BUFS < 3) return;   // never launched in that build
  constexpr int V = Vec<T>::N;
  constexpr int CPT = U * V;
  constexpr int ACC = LXF ? CPT + 2 : CPT + 1;     // look-ahead sums per thread
  using A = Ar<T, FAST>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  NlSmem<T>& sm = *reinterpret_cast<NlSmem<T>*>(smem_raw);

  const int t = threadIdx.x;
  const int NC = a.nc;
  const int N = a.N;
  const int n_eta = a.n_eta;
  const uint32_t half_bytes = (uint32_t)N * (uint32_t)sizeof(T);   // densities of one road
  const uint32_t row_bytes = 2u * half_bytes;
  const uint32_t tile_out_bytes = (uint32_t)a.epc * row_bytes;
  const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int last_tile = a.n_tiles - 1;
  const int last_envs = a.B - last_tile * a.epc;

  const uint32_t in0 = smem_u32(sm.in[0]), out0 = smem_u32(sm.out[0]);
  const uint32_t full0 = smem_u32(&sm.full[0]), empty0 = smem_u32(&sm.empty[0]);
  const uint32_t ofull0 = smem_u32(&sm.out_full[0]), ofree0 = smem_u32(&sm.out_free[0]);

  for (int i = t; i < n_eta; i += blockDim.x) sm.gamma[i] = a.gamma[i];
  if (t == 0) {
    for (int s = 0; s < NL_STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], (uint32_t)(NC >> 5));
    }
    for (int o = 0; o < PIPE_OUT; ++o) {
      mbar_init(&sm.out_full[o], (uint32_t)(NC >> 5));
      mbar_init(&sm.out_free[o], 1);
    }
    for (int b = 0; b < 3; ++b) mbar_init(&sm.pub[b], (uint32_t)(NC >> 5));
    fence_mbar_init();
  }
  __syncthreads();

  // ======================= producer warp ===================================
  if (t >= NC) {
    if (t == NC) {
      const unsigned char* gin = reinterpret_cast<const unsigned char*>(a.in);
      unsigned char* gout = reinterpret_cast<unsigned char*>(a.out);
      auto envs_of = [&](int tile) -> int { return tile == last_tile ? last_envs : a.epc; };
      auto issue_load = [&](int i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int s = i % NL_STAGES;
        const int envs = envs_of(tile);
        mbar_expect_tx(full0 + 8 * s, (uint32_t)envs * half_bytes);
        const unsigned char* src = gin + (size_t)tile * a.epc * row_bytes;
        for (int el = 0; el < envs; ++el)   // density half of each road
          bulk_load(in0 + (uint32_t)s * NL_IN_BYTES + (uint32_t)el * half_bytes,
                    src + (size_t)el * row_bytes, half_bytes, full0 + 8 * s);
      };
      const int pre = n_my < NL_STAGES ? n_my : NL_STAGES;
      for (int i = 0; i < pre; ++i) issue_load(i);
      for (int i = 0; i < n_my; ++i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int s = i % NL_STAGES, o = i % PIPE_OUT;
        if (i + NL_STAGES < n_my) {
          mbar_wait(empty0 + 8 * s, (uint32_t)((i / NL_STAGES) & 1));
          issue_load(i + NL_STAGES);
        }
        mbar_wait(ofull0 + 8 * o, (uint32_t)((i / PIPE_OUT) & 1));
        bulk_store(gout + (size_t)tile * tile_out_bytes, out0 + (uint32_t)o * PIPE_TILE_BYTES,
                   (uint32_t)envs_of(tile) * row_bytes);
        bulk_commit();
        bulk_wait_read<PIPE_OUT - 1>();
        if (i >= PIPE_OUT - 1) mbar_arrive(ofree0 + 8 * ((i - (PIPE_OUT - 1)) % PIPE_OUT));
      }
      bulk_wait_read<0>();
    }
    return;
  }

  // ======================= consumers =======================================
  const int el = t / a.tpe;
  const int pos = t - el * a.tpe;
  const bool row_first = (pos == 0), row_last = (pos == a.nthr - 1);
  const bool owns_cells = (pos < a.nthr) && (el < a.epc);
  const bool lane0 = (t & 31) == 0;
  const bool want_reward = a.reward != nullptr;
  const int j0 = pos * CPT;                        // my first cell
  const uint32_t in_off = (uint32_t)el * half_bytes + (uint32_t)j0 * (uint32_t)sizeof(T);
  const uint32_t out_off = (uint32_t)el * row_bytes + (uint32_t)j0 * (uint32_t)sizeof(T);
  const int VSTR = a.ve_cols;                      // plane stride
  const int ve_row = el * (CPT * VSTR) + pos;      // my column in plane 0

  EnvC<T> e;
  if (UNIFORM) load_uniform(e, a.u, 0, (const T*)nullptr, a.step);

  auto tile_of = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
  auto active_in = [&](int tile) {
    const int envs_here = tile == last_tile ? last_envs : a.epc;
    return owns_cells && (el < envs_here);
  };

  // ---- first half of a tile: my cells out of input stage s, V(rho) (LXF: rho)
  // of my cells and of the ghosts into `ve` ------------------------------------
  auto publish_tile = [&](int s, bool active, T* ve, T (&rho)[CPT], T& rho_left) {
    rho_left = T(0);
    if (!active) return;
    const uint32_t src = in0 + (uint32_t)s * NL_IN_BYTES + in_off;
    const T* rho_row = reinterpret_cast<const T*>(sm.in[s]) + (size_t)el * N;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      T tmp[V];
      lds16(src + 16 * u, tmp);
#pragma unroll
      for (int k = 0; k < V; ++k) rho[u * V + k] = tmp[k];
    }
    if constexpr (!LXF) {
      T mine[CPT];
      T ghost_v = T(0);   // V(rho_R) for a CONSTANT downstream end
      with_lam<LAMK>(e.flags, [&](auto kind) {
        constexpr int K = decltype(kind)::value;
#pragma unroll
        for (int k = 0; k < CPT; ++k) mine[k] = v_eq<T, FAST, K>(rho[k], e);
        if (row_last && a.bc.bc_r == MT_BC_CONSTANT) ghost_v = v_eq<T, FAST, K>(a.bc.br0, e);
      });
#pragma unroll
      for (int k = 0; k < CPT; ++k) ve[ve_row + k * VSTR] = mine[k];
      // downstream ghosts V(rho_{N+c}), c < n_eta: cell N+c sits in plane
      // (c mod CPT), column nthr + c / CPT
      if (a.bc.bc_r == MT_BC_LOOP) {
#pragma unroll
        for (int k = 0; k < CPT; ++k)
          if (j0 + k < n_eta) ve[ve_row + k * VSTR + a.nthr] = mine[k];
      } else if (row_last) {
        const T g = (a.bc.bc_r == MT_BC_CONSTANT) ? ghost_v : mine[CPT - 1];
        T* blk = ve + el * (CPT * VSTR);
        for (int c = 0; c < n_eta; ++c) blk[(c % CPT) * VSTR + a.nthr + c / CPT] = g;
      }
      rho_left = row_first ? nl_left_ghost(a.bc, rho_row, N) : rho_row[j0 - 1];
    } else {
      // LXF: the densities themselves, cell c in slot c + 1 (slot 0 = upstream
      // ghost): my cell j0 + k sits in plane (k + 1) mod CPT, column pos (+ 1
      // for my last cell)
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        const int idx = ve_row + (k < CPT - 1 ? (k + 1) * VSTR : 1);
        ve[idx] = rho[k];
        // ring: cell c < n_eta is also the downstream ghost N + c
        if (a.bc.bc_r == MT_BC_LOOP && j0 + k < n_eta) ve[idx + a.nthr] = rho[k];
      }
      if (row_first) ve[el * (CPT * VSTR)] = nl_left_ghost(a.bc, rho_row, N);
      if (row_last && a.bc.bc_r != MT_BC_LOOP) {
        const T g = (a.bc.bc_r == MT_BC_CONSTANT) ? a.bc.br0 : rho[CPT - 1];
        T* blk = ve + el * (CPT * VSTR);
        for (int c = 1; c <= n_eta; ++c) blk[(c % CPT) * VSTR + a.nthr + c / CPT] = g;
      }
    }
  };

  // ---- second half: look-ahead sums, fluxes, new state into output stage o ----
  auto finish_tile = [&](int o, bool active, int env, const T* ve, const T (&rho)[CPT],
                         T rho_left) {
    T num = 0, den = 0;
    if (active) {
      // Godunov: W for cells j0-1 .. j0+CPT-1: acc[m] covers cell j0-1+m and
      // needs ve[j0+m+k], k < n_eta.  LXF: R for cells j0-1 .. j0+CPT: acc[m]
      // covers cell j0-1+m and needs slot j0+m+k.  Either way thread-relative
      // index m + k: a register window slides by one per k.
      T acc[ACC], win[ACC];
      const T* vcol = ve + ve_row;                   // index x: plane x % CPT, column + x / CPT
#pragma unroll
      for (int m = 0; m < ACC; ++m) {
        acc[m] = T(0);
        win[m] = vcol[(m % CPT) * VSTR + m / CPT];
      }
      // LXF: the window's first and last entry are the two neighbour cells
      const T cell_l = win[0], cell_r = win[ACC - 1];
      for (int k = 0; k < n_eta; ++k) {
        const T g = sm.gamma[k];
        const int x = k + ACC;
        const T nxt = vcol[(x % CPT) * VSTR + x / CPT];
#pragma unroll
        for (int m = 0; m < ACC; ++m) acc[m] = A::fma(g, win[m], acc[m]);  // fma(gamma_k, V, w)
#pragma unroll
        for (int m = 0; m < ACC - 1; ++m) win[m] = win[m + 1];
        win[ACC - 1] = nxt;
      }
      T rn[CPT], vn[CPT];
      if constexpr (!LXF) {
        T Fm = A::mul(rho_left, acc[0]);               // F_{j0-1/2}
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const T F = A::mul(rho[k], acc[k + 1]);
          rn[k] = A::sub(rho[k], A::mul(e.step, A::sub(F, Fm)));
          Fm = F;
        }
        (void)cell_l; (void)cell_r;
      } else {
        T c[CPT + 2], gq[CPT + 2];
        c[0] = cell_l;
        c[CPT + 1] = cell_r;
#pragma unroll
        for (int k = 0; k < CPT; ++k) c[k + 1] = rho[k];
        with_lam<LAMK>(e.flags, [&](auto kind) {
          constexpr int K = decltype(kind)::value;
#pragma unroll
          for (int m = 0; m < CPT + 2; ++m) gq[m] = A::mul(c[m], v_eq<T, FAST, K>(acc[m], e));
        });
        const T ha = a.half_alpha < T(0) ? A::mul(T(0.5), e.v_max) : a.half_alpha;
        T Fm = A::add(A::mul(T(0.5), A::add(gq[0], gq[1])), A::mul(ha, A::sub(c[0], c[1])));
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const T F = A::add(A::mul(T(0.5), A::add(gq[k + 1], gq[k + 2])),
                             A::mul(ha, A::sub(c[k + 1], c[k + 2])));
          rn[k] = A::sub(rho[k], A::mul(e.step, A::sub(F, Fm)));
          Fm = F;
        }
      }
      with_lam<LAMK>(e.flags, [&](auto kind) {
        constexpr int K = decltype(kind)::value;
#pragma unroll
        for (int k = 0; k < CPT; ++k) vn[k] = v_eq<T, FAST, K>(rn[k], e);
      });
      if (want_reward) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) { num += rn[k] * vn[k]; den += rn[k]; }
      }
      const uint32_t dst = out0 + (uint32_t)o * PIPE_TILE_BYTES + out_off;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T tr[V], tv[V];
#pragma unroll
        for (int k = 0; k < V; ++k) { tr[k] = rn[u * V + k]; tv[k] = vn[u * V + k]; }
        sts16(dst + 16 * u, tr);
        sts16(dst + half_bytes + 16 * u, tv);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane0) mbar_arrive(ofull0 + 8 * o);

    if (want_reward) {
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
        if (lane0) { sm.s_num[t >> 5] = num; sm.s_den[t >> 5] = den; }
        consumer_sync(NC);
        if (active && row_first) {
          const int w0 = t >> 5, nw = a.tpe >> 5;
          T n = sm.s_num[w0], d = sm.s_den[w0];
          for (int j = 1; j < nw; ++j) { n += sm.s_num[w0 + j]; d += sm.s_den[w0 + j]; }
          a.reward[env] = n / d;
        }
      }
    }
  };

  if constexpr (!AHEAD) {
    int s = 0, o = 0;
    uint32_t ph_in = 0, ph_out = 0;
    for (int i = 0; i < n_my; ++i) {
      const int tile = tile_of(i);
      const int env = tile * a.epc + el;
      const bool active = active_in(tile);
      if (active && (!UNIFORM || a.action != nullptr)) {
        if (UNIFORM) load_uniform(e, a.u, env, a.action, a.step);
        else load_env(e, a.table, a.flags, env, a.action, a.step);
      }
      T* const ve = sm.ve[i & 1];
      mbar_wait(full0 + 8 * s, ph_in);
      T rho[CPT];
      T rho_left;
      publish_tile(s, active, ve, rho, rho_left);
      consumer_sync(NC);                               // V(rho) published; stage s read
      if (lane0) mbar_arrive(empty0 + 8 * s);
      mbar_wait(ofree0 + 8 * o, ph_out ^ 1);
      finish_tile(o, active, env, ve, rho, rho_left);
      if (++s == NL_STAGES) { s = 0; ph_in ^= 1; }
      if (++o == PIPE_OUT) { o = 0; ph_out ^= 1; }
    }
  } else {
    // Tile i + 1 is published BEFORE the look-ahead sums of tile i, into one of
    // three V(rho) buffers; the CTA barrier becomes an mbarrier per buffer that
    // is long complete by the time it is waited for (the sums take most of a
    // tile's time), and a warp releases an input stage on its own.  A warp may
    // write buffer (i + 1) % 3 once every warp has finished the sums of tile
    // i - 2: that is implied by "every warp has published tile i" (a warp
    // publishes tile i right after its sums of tile i - 2), which this warp
    // waits for first -- it needs tile i's buffer anyway.  Scalar parameters
    // only (two tiles' per-road constants would have to be live at once).
    const uint32_t pub0 = smem_u32(&sm.pub[0]);
    T rho_c[CPT], rho_n[CPT];
    T left_c = 0, left_n = 0;
    bool act_c = active_in(tile_of(0)), act_n = false;
    auto fetch = [&](int i, bool active, T (&rho)[CPT], T& rho_left) {
      const int s = i % NL_STAGES, b = i % 3 % NL_VE_BUFS;   // (AHEAD needs three buffers)
      mbar_wait(full0 + 8 * s, (uint32_t)((i / NL_STAGES) & 1));
      publish_tile(s, active, sm.ve[b], rho, rho_left);
      __syncwarp();
      if (lane0) {
        mbar_arrive(pub0 + 8 * b);       // my warp's part of tile i is published
        mbar_arrive(empty0 + 8 * s);     // ... and its part of stage s is read
      }
    };
    fetch(0, act_c, rho_c, left_c);
    int o = 0;
    uint32_t ph_out = 0;
    for (int i = 0; i < n_my; ++i) {
      const int tile = tile_of(i);
      const int env = tile * a.epc + el;
      mbar_wait(pub0 + 8 * (i % 3), (uint32_t)((i / 3) & 1));
      if (i + 1 < n_my) {
        act_n = active_in(tile_of(i + 1));
        fetch(i + 1, act_n, rho_n, left_n);
      }
      mbar_wait(ofree0 + 8 * o, ph_out ^ 1);
      finish_tile(o, act_c, env, sm.ve[i % 3 % NL_VE_BUFS], rho_c, left_c);
#pragma unroll
      for (int k = 0; k < CPT; ++k) rho_c[k] = rho_n[k];
      left_c = left_n;
      act_c = act_n;
      if (++o == PIPE_OUT) { o = 0; ph_out ^= 1; }
    }
  }
}

// ---------------------------------------------------------------------------
// scalar kernel: any N, one thread per cell, everything re-derived from
// global memory.  Fallback for shapes the pipelined kernel does not take and
// the independent cross-check in the tests.
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T nl_rho_at(const BcArgs<T>& b, const T* __restrict__ row, int N, int j) {
  if (j < 0) return nl_left_ghost(b, row, N);
  if (j < N) return row[j];
  if (b.bc_r == MT_BC_LOOP) return row[j - N];
  if (b.bc_r == MT_BC_CONSTANT) return b.br0;
  return row[N - 1];
}

template <typename T> struct NlScalarArgs {
  const T* in;
  T* out;
  const T* table;
  const int* flags;
  const T* action;
  const T* gamma;
  int n_eta;
  long long B;
  int N;
  BcArgs<T> bc;
  T step;
  T half_alpha;
};

// LXF scheme, one thread per cell (see the header comment)
template <typename T, bool FAST>
__global__ void nonlocal_lxf_scalar_kernel(const NlScalarArgs<T> a) {
  using A = Ar<T, FAST>;
  constexpr int LAMK = LAM_RUNTIME;
  const int N = a.N;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= a.B * N) return;
  const long long env = gid / N;
  const int j = (int)(gid - env * N);
  EnvC<T> e;
  load_env(e, a.table, a.flags, env, a.action, a.step);
  const T* row = a.in + env * (2LL * N);
  T r[3] = {T(0), T(0), T(0)};       // R of cells j-1, j, j+1
  for (int k = 0; k < a.n_eta; ++k) {
    const T g = a.gamma[k];
#pragma unroll
    for (int m = 0; m < 3; ++m) r[m] = A::fma(g, nl_rho_at(a.bc, row, N, j - 1 + m + k), r[m]);
  }
  T c[3], gq[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    c[m] = nl_rho_at(a.bc, row, N, j - 1 + m);
    gq[m] = A::mul(c[m], v_eq<T, FAST, LAMK>(r[m], e));
  }
  const T ha = a.half_alpha < T(0) ? A::mul(T(0.5), e.v_max) : a.half_alpha;
  const T Fm = A::add(A::mul(T(0.5), A::add(gq[0], gq[1])), A::mul(ha, A::sub(c[0], c[1])));
  const T F = A::add(A::mul(T(0.5), A::add(gq[1], gq[2])), A::mul(ha, A::sub(c[1], c[2])));
  const T rn = A::sub(c[1], A::mul(e.step, A::sub(F, Fm)));
  T* orow = a.out + env * (2LL * N);
  orow[j] = rn;
  orow[N + j] = v_eq<T, FAST, LAMK>(rn, e);
}

template <typename T, bool FAST>
__global__ void nonlocal_scalar_kernel(const NlScalarArgs<T> a) {
  using A = Ar<T, FAST>;
  constexpr int LAMK = LAM_RUNTIME;
  const int N = a.N;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= a.B * N) return;
  const long long env = gid / N;
  const int j = (int)(gid - env * N);
  EnvC<T> e;
  load_env(e, a.table, a.flags, env, a.action, a.step);
  const T* row = a.in + env * (2LL * N);
  T w = T(0), wm = T(0);
  for (int k = 0; k < a.n_eta; ++k) {
    const T g = a.gamma[k];
    w = A::fma(g, v_eq<T, FAST, LAMK>(nl_rho_at(a.bc, row, N, j + k + 1), e), w);
    wm = A::fma(g, v_eq<T, FAST, LAMK>(nl_rho_at(a.bc, row, N, j + k), e), wm);
  }
  const T rho = row[j];
  const T F = A::mul(rho, w);
  const T Fm = A::mul(nl_rho_at(a.bc, row, N, j - 1), wm);
  const T rn = A::sub(rho, A::mul(e.step, A::sub(F, Fm)));
  T* orow = a.out + env * (2LL * N);
  orow[j] = rn;
  orow[N + j] = v_eq<T, FAST, LAMK>(rn, e);
}

}  // namespace mt
