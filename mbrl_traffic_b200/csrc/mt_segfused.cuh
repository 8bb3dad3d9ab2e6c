// mt_segfused.cuh -- several model steps per launch for roads LONGER than one
// tile (config 4: a 10 M-cell highway, whole or one rank's block of it), with
// the ghost-cell exchange of a domain-decomposed road folded into the launch.
//
// A tile is a WINDOW of nc * CPT consecutive cells of one road.  Neighbouring
// windows overlap by 2 * hh cells, hh = n_inner rounded up to a multiple of 4:
// a window edge that is not a true road end is treated like a rank's ghost
// edge (MT_BC_HALO: the outermost cell is passed through), so after j steps
// the j cells next to it are stale -- the stencil has radius one per step
// (lwr.py:234, arz.py:456) -- and after n_inner steps the cells at least hh
// away from such an edge are exactly what step-by-step stepping of the whole
// road gives.  Only those cells are stored.  The steps themselves are the
// loop of step_fused_kernel: state in registers, three edge values per thread
// through shared memory, one named barrier per step; HBM sees every cell once
// per launch (plus the 2 * hh / window overlap) instead of once per step.
//
// Ghost-cell exchange (SegHalo, mt_halo.cuh mailboxes).  With a pull epoch the
// producer of the window that holds a rank's ghost cells waits for the
// neighbour's flag and fetches those cells from this rank's MAILBOX instead of
// the block; with a push epoch the producer of the window that holds the
// rank's first / last `halo` own cells stores them -- besides the block --
// straight into the neighbour's mailbox (peer memory, NVLink), waits for the
// bulk stores to complete and publishes the epoch (st.release.sys).  One
// exchange period is then ONE launch per rank: no exchange kernels, no host
// round trip, and the windows in the middle of the block never wait.
#pragma once
#include <stdint.h>
#include "mtraffic.h"
#include "mt_math.cuh"
#include "mt_cell.cuh"
#include "mt_pipe.cuh"
#include "mt_halo.cuh"

namespace mt {

struct SegHalo {
  void* mine = nullptr;    // this rank's mailbox
  void* left = nullptr;    // neighbours' mailboxes (peer mappings) or nullptr
  void* right = nullptr;
  int halo = 0;            // ghost cells per internal edge
  unsigned pull_epoch = 0; // != 0: ghost cells come from the mailbox (after waiting for this epoch)
  unsigned push_epoch = 0; // != 0: edge cells also go to the neighbours, published as this epoch
  unsigned long long timeout_ns = 0;
};

template <typename T> struct SegFusedArgs {
  const T* in;
  T* out;
  const T* table;
  const int* flags;
  const T* action;          // [n_inner, action_stride] speed limits or nullptr
  long long action_stride;
  int B;
  int N;
  int nc;                   // consumer threads; a window has nc * CPT cells
  int n_inner;              // model steps per launch
  int hh;                   // window overlap on each side (>= n_inner, multiple of 4)
  int tiles_per_row;
  int n_tiles;              // B * tiles_per_row
  BcArgs<T> bc;             // rules of the true road ends
  T step;
  UniformC<T> u;
  SegHalo halo;
};

// windows per road: the first covers [0, W), the others start S = W - 2 hh
// apart, the last is right-aligned to the road's end
inline int segfused_tiles_per_row(int N, int window, int hh) {
  const int S = window - 2 * hh;
  return N <= window ? 1 : 1 + (N - window + S - 1) / S;
}

template <typename T, int MODEL, bool FAST, int LAMK, bool UNIFORM, int U>
__global__ void __launch_bounds__(PIPE_BLOCK / U + 32, PIPE_CTAS)
step_segfused_kernel(const SegFusedArgs<T> a) {
  constexpr int V = Vec<T>::N;
  constexpr int CPT = U * V;
  constexpr bool ARZ = (MODEL == 1);
  constexpr int STAGES = PipeCfg<T, U>::STAGES;
  constexpr int XCH = PipeCfg<T, U>::XCH;
  constexpr uint32_t HALF = PIPE_TILE_BYTES / 2;   // v window inside a stage
  constexpr uint32_t ES = (uint32_t)sizeof(T);
  using A = Ar<T, FAST>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PipeSmem<T, U>& sm = *reinterpret_cast<PipeSmem<T, U>*>(smem_raw);

  const int t = threadIdx.x;
  const int NC = a.nc;
  const int N = a.N;
  const int W = NC * CPT;                // cells per window
  const int S = W - 2 * a.hh;            // distance between window starts
  const int TPR = a.tiles_per_row;
  const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t in0 = smem_u32(sm.in[0]), out0 = smem_u32(sm.out[0]);
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

  // tile -> (road, window start, first valid cell, one past the last valid cell)
  auto locate = [&](int tile, int& env, int& j, int& w, int& vs, int& ve) {
    env = tile / TPR;
    j = tile - env * TPR;
    // With a ghost-cell exchange folded in, the road's LAST window trades places
    // with its second: both edge windows are then in the launch's first wave, so
    // a rank's pushes are out long before its neighbours start the next launch
    // (whose edge windows wait for exactly those flags) instead of at its very
    // end.  Windows of one launch are independent: any order gives the same cells.
    if (a.halo.push_epoch != 0u && TPR > 2) {
      if (j == 1) j = TPR - 1;
      else if (j == TPR - 1) j = 1;
    }
    const bool last = (j == TPR - 1);
    w = last ? N - W : j * S;
    vs = (j == 0) ? 0 : a.hh + j * S;
    ve = last ? N : a.hh + (j + 1) * S;
  };

  // ======================= producer warp ===================================
  if (t >= NC) {
    if (t == NC) {
      const SegHalo& hx = a.halo;
      const int h = hx.halo;
      const uint32_t hbytes = (uint32_t)h * ES;
      // wait until the neighbour on `side` has published hx.pull_epoch in my mailbox
      auto wait_flag = [&](int side) {
        HaloBox* box = reinterpret_cast<HaloBox*>(hx.mine);
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(ld_acquire_sys(&box->flag[side]) - hx.pull_epoch) < 0) {
          __nanosleep(100);
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (t1 - t0 > hx.timeout_ns) { box->error = 1u; break; }
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");  // acquired data -> bulk engine
      };
      auto issue_load = [&](int i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int s = i % STAGES;
        int env, j, w, vs, ve;
        locate(tile, env, j, w, vs, ve);
        const T* row = a.in + (size_t)env * 2 * N;
        const uint32_t bar = full0 + 8 * s;
        const uint32_t dst = in0 + (uint32_t)s * PIPE_TILE_BYTES;
        const uint32_t fields = ARZ ? 2u : 1u;
        // ghost cells of this rank's block inside the window, taken from the mailbox
        const bool pull_l = hx.pull_epoch != 0 && hx.left != nullptr && j == 0;
        const bool pull_r = hx.pull_epoch != 0 && hx.right != nullptr && j == TPR - 1;
        mbar_expect_tx(bar, (uint32_t)W * ES * fields);
        const int c0 = pull_l ? h : 0;                 // window cells [c0, c1) come from the block
        const int c1 = pull_r ? W - h : W;
        bulk_load(dst + (uint32_t)c0 * ES, row + w + c0, (uint32_t)(c1 - c0) * ES, bar);
        if (ARZ) bulk_load(dst + HALF + (uint32_t)c0 * ES, row + N + w + c0, (uint32_t)(c1 - c0) * ES, bar);
        const unsigned par = hx.pull_epoch & 1u;
        if (pull_l) {
          wait_flag(HALO_FROM_LEFT);
          const T* src = halo_slot<T>(hx.mine, h, par, HALO_FROM_LEFT);
          bulk_load(dst, src, hbytes, bar);
          if (ARZ) bulk_load(dst + HALF, src + h, hbytes, bar);
        }
        if (pull_r) {
          wait_flag(HALO_FROM_RIGHT);
          const T* src = halo_slot<T>(hx.mine, h, par, HALO_FROM_RIGHT);
          bulk_load(dst + (uint32_t)(W - h) * ES, src, hbytes, bar);
          if (ARZ) bulk_load(dst + HALF + (uint32_t)(W - h) * ES, src + h, hbytes, bar);
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
        int env, j, w, vs, ve;
        locate(tile, env, j, w, vs, ve);
        T* orow = a.out + (size_t)env * 2 * N;
        const uint32_t src = out0 + (uint32_t)o * PIPE_TILE_BYTES;
        const uint32_t off = (uint32_t)(vs - w) * ES, bytes = (uint32_t)(ve - vs) * ES;
        bulk_store(orow + vs, src + off, bytes);
        bulk_store(orow + N + vs, src + HALF + off, bytes);
        const bool push_l = hx.push_epoch != 0 && hx.left != nullptr && j == 0;
        const bool push_r = hx.push_epoch != 0 && hx.right != nullptr && j == TPR - 1;
        if (push_l | push_r) {
          const unsigned par = hx.push_epoch & 1u;
          if (push_l) {   // my first own cells [h, 2h): I am my left neighbour's RIGHT neighbour
            T* dstp = halo_slot<T>(hx.left, h, par, HALO_FROM_RIGHT);
            bulk_store(dstp, src + (uint32_t)(h - w) * ES, hbytes);
            bulk_store(dstp + h, src + HALF + (uint32_t)(h - w) * ES, hbytes);
          }
          if (push_r) {   // my last own cells [N - 2h, N - h)
            T* dstp = halo_slot<T>(hx.right, h, par, HALO_FROM_LEFT);
            bulk_store(dstp, src + (uint32_t)(N - 2 * h - w) * ES, hbytes);
            bulk_store(dstp + h, src + HALF + (uint32_t)(N - 2 * h - w) * ES, hbytes);
          }
          bulk_commit();
          bulk_wait_all<0>();                                   // the peer stores have landed
          asm volatile("fence.proxy.async.global;" ::: "memory");
          __threadfence_system();
          if (push_l) st_release_sys(&reinterpret_cast<HaloBox*>(hx.left)->flag[HALO_FROM_RIGHT], hx.push_epoch);
          if (push_r) st_release_sys(&reinterpret_cast<HaloBox*>(hx.right)->flag[HALO_FROM_LEFT], hx.push_epoch);
        } else {
          bulk_commit();
        }
        bulk_wait_read<PIPE_OUT - 1>();
        if (i >= PIPE_OUT - 1) mbar_arrive(ofree0 + 8 * ((i - (PIPE_OUT - 1)) % PIPE_OUT));
      }
      bulk_wait_read<0>();
    }
    return;
  }

  // ======================= consumers =======================================
  const bool lane0 = (t & 31) == 0;
  const uint32_t my_off = (uint32_t)t * (U * 16);
  const int xi = t;                             // exchange slot (slot 0 of xD / xR is the pad)
  const bool seg_first = (t == 0), seg_last = (t == NC - 1);

  EnvC<T> e;
  if (UNIFORM) {
    load_uniform(e, a.u, 0, (const T*)nullptr, a.step);
    if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
  }

  int s = 0, o = 0, xpar = 0;
  uint32_t ph_in = 0, ph_out = 0;
  const int n_inner = a.n_inner;
  for (int i = 0; i < n_my; ++i) {
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    int env, j, w, vs, ve;
    locate(tile, env, j, w, vs, ve);
    // a window edge inside the road behaves like a ghost edge
    BcArgs<T> b = a.bc;
    if (j != 0) b.bc_l = MT_BC_HALO;
    if (j != TPR - 1) b.bc_r = MT_BC_HALO;

    if (!UNIFORM) {
      load_env(e, a.table, a.flags, env, (const T*)nullptr, a.step);
      if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
    }

    mbar_wait(full0 + 8 * s, ph_in);

    T rho[CPT], v[CPT], y[CPT], r[CPT], D[CPT], Sx[CPT];
    {
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
        } else {
#pragma unroll
          for (int k = 0; k < V; ++k) v[u * V + k] = T(0);
        }
      }
    }
    T act_next = T(0);
    if (a.action != nullptr) act_next = __ldg(a.action + env);
    for (int jj = 0; jj < n_inner; ++jj) {
      const bool last = (jj == n_inner - 1);
      T* const xS = sm.xS[xpar];
      T* const xD = sm.xD[xpar];
      T* const xR = sm.xR[xpar];
      xpar ^= 1;
      if (a.action != nullptr) {   // this step's speed limit for my road
        e.v_max = act_next;
        if (jj + 1 < n_inner) act_next = __ldg(a.action + (size_t)(jj + 1) * a.action_stride + env);
        if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
      }
      with_lam<LAMK>(e.flags, [&](auto kind) {
        constexpr int K = decltype(kind)::value;
        if (ARZ) arz_pass1<T, FAST, K, CPT>(e, rho, v, y, r, D, Sx);
        else lwr_pass1<T, FAST, K, CPT>(e, rho, D, Sx);
      });
      if (ARZ) xR[xi + 1] = r[CPT - 1];
      xS[xi] = Sx[0];
      xD[xi + 1] = D[CPT - 1];
      if (seg_last) xS[xi + 1] = Sx[CPT - 1];   // arz.py:405 / lwr.py:294 (true road end; unused at a ghost edge)
      T rn[CPT], vn[CPT], Q[CPT], P[CPT];
      with_lam<LAMK>(e.flags, [&](auto kind) {
        cell_mid<T, ARZ, FAST, decltype(kind)::value, CPT>(e, rho, y, r, D, Sx, Q, P, rn, vn);
      });
      consumer_sync(NC);                                 // exchange visible
      if (jj == 0 && lane0) mbar_arrive(empty0 + 8 * s);   // stage s fully read: release it
      if (last) mbar_wait(ofree0 + 8 * o, ph_out ^ 1);     // output stage drained by its last store
      {
        const T S_right = xS[xi + 1], D_left = xD[xi], r_left = xR[xi];
        with_lam<LAMK>(e.flags, [&](auto kind) {
          cell_edges<T, ARZ, FAST, decltype(kind)::value, CPT>(
              e, b, seg_first, seg_last, rho, v, y, r, D, Sx, Q, P, S_right, D_left, r_left,
              T(0), T(0), rn, vn);
        });
      }
      if (!last) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) { rho[k] = rn[k]; v[k] = vn[k]; }
      } else {
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
    }
    fence_proxy_async_smem();  // my generic-proxy writes -> visible to the bulk engine
    __syncwarp();
    if (lane0) mbar_arrive(ofull0 + 8 * o);          // non-blocking hand-off to the producer
    if (++s == STAGES) { s = 0; ph_in ^= 1; }
    if (++o == PIPE_OUT) { o = 0; ph_out ^= 1; }
  }
}

}  // namespace mt
