// mt_fused.cuh -- multi-step rollouts in ONE launch.
//
// Same producer/consumer pipeline as mt_pipe.cuh, but every tile visit
// advances its roads by a.n_inner model steps: the state stays in registers
// between steps, only the three edge values (and, for the LOOP rule, the old
// boundary-source cells, which travel in the road's pad slots of the exchange
// arrays) cross threads, at one named barrier per step.  HBM sees each road
// once per rollout instead of once per step, so the loop of
// experiments/evaluate_model.py:398-403 (and any rollout without a policy in
// the loop) runs at the arithmetic rate of the chip instead of the bandwidth
// rate.  Per-step speed limits are read as they are needed
// (action[j * stride + road]); a reward row is produced for the final step or
// for every step (mt_rollout with rewards: the K-shoot planner).
// Kept apart from step_pipe_kernel so that the single-step kernel's code and
// register allocation stay exactly as tuned.
//
// REDUCER (a reward row per step, roads wider than a warp, f64): the CTA gets
// one more warp.  Each step the consumers leave their (sum rho v, sum rho) in
// shared memory and arrive -- without waiting -- on a named barrier the reducer
// warp sleeps on; it adds the road's pairs up in a fixed order and stores the
// reward, then frees the (double-buffered) pair slots through an mbarrier.
// The consumers are spared the per-step butterfly and the road's first warp
// the per-step finish that used to sit on every step's critical path.
#pragma once
#include <stdint.h>
#include "mtraffic.h"
#include "mt_math.cuh"
#include "mt_cell.cuh"
#include "mt_pipe.cuh"

namespace mt {

template <typename T, int MODEL, bool FAST, int LAMK, bool UNIFORM, int U, bool REDUCER = false>
__global__ void __launch_bounds__(PIPE_BLOCK / U + (REDUCER ? 64 : 32), PIPE_CTAS)
step_fused_kernel(const PipeArgs<T> a) {
  constexpr int V = Vec<T>::N;
  constexpr int CPT = U * V;
  constexpr bool ARZ = (MODEL == 1);
  constexpr int PIPE_STAGES_ALL = PipeCfg<T, U>::STAGES;
  // the reducer's pair slots (and its mbarriers) take the last input stage
  constexpr int PIPE_STAGES = REDUCER ? PIPE_STAGES_ALL - 1 : PIPE_STAGES_ALL;
  constexpr int PIPE_XCH = PipeCfg<T, U>::XCH;
  using A = Ar<T, FAST>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PipeSmem<T, U>& sm = *reinterpret_cast<PipeSmem<T, U>*>(smem_raw);

  const int t = threadIdx.x;
  const int NC = a.nc;
  const int N = a.N;
  const uint32_t row_bytes = 2u * (uint32_t)N * (uint32_t)sizeof(T);
  const uint32_t tile_bytes = (uint32_t)a.epc * row_bytes;
  const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int last_tile = a.n_tiles - 1;
  const int last_envs = a.B - last_tile * a.epc;  // roads in the (ragged) last tile

  const uint32_t in0 = smem_u32(sm.in[0]), out0 = smem_u32(sm.out[0]);
  const uint32_t full0 = smem_u32(&sm.full[0]), empty0 = smem_u32(&sm.empty[0]);
  const uint32_t ofull0 = smem_u32(&sm.out_full[0]), ofree0 = smem_u32(&sm.out_free[0]);

  // REDUCER: [2][NC][2] sums, then two mbarriers "pairs of buffer b summed"
  T* const pairbuf = reinterpret_cast<T*>(sm.in[PIPE_STAGES_ALL - 1]);
  uint64_t* const rfree = reinterpret_cast<uint64_t*>(sm.in[PIPE_STAGES_ALL - 1] + PIPE_TILE_BYTES / 2);
  const uint32_t rfree0 = smem_u32(rfree);

  pdl_launch_dependents();
  // pad slots of the exchange arrays stay zero for the kernel's life: they
  // feed the (discarded) left-face flux of each road's first cell
  for (int i = t; i < 2 * PIPE_XCH; i += blockDim.x) { (&sm.xD[0][0])[i] = T(0); (&sm.xR[0][0])[i] = T(0); }
  if (t == 0) {
    for (int s = 0; s < PIPE_STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], (uint32_t)(NC >> 5));
    }
    for (int o = 0; o < PIPE_OUT; ++o) {
      mbar_init(&sm.out_full[o], (uint32_t)(NC >> 5));
      mbar_init(&sm.out_free[o], 1);
    }
    if constexpr (REDUCER) { mbar_init(&rfree[0], 1); mbar_init(&rfree[1], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait_prior_grid();  // everything above overlapped the previous kernel's tail

  // ======================= producer warp ===================================
  if (t >= NC) {
    if (t == NC) {
      const unsigned char* gin = reinterpret_cast<const unsigned char*>(a.in);
      unsigned char* gout = reinterpret_cast<unsigned char*>(a.out);
      auto bytes_of = [&](int tile) -> uint32_t {
        return tile == last_tile ? (uint32_t)last_envs * row_bytes : tile_bytes;
      };
      auto issue_load = [&](int i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int s = i % PIPE_STAGES;
        if (ARZ) {
          const uint32_t bytes = bytes_of(tile);
          mbar_expect_tx(full0 + 8 * s, bytes);
          bulk_load(in0 + (uint32_t)s * PIPE_TILE_BYTES, gin + (size_t)tile * tile_bytes, bytes,
                    full0 + 8 * s);
        } else {
          // LWR ignores the input velocity (lwr.py:197): fetch the density half
          // of each road only, at its usual place in the stage
          const int envs = tile == last_tile ? last_envs : a.epc;
          const uint32_t half = row_bytes >> 1;
          mbar_expect_tx(full0 + 8 * s, (uint32_t)envs * half);
          for (int r = 0; r < envs; ++r)
            bulk_load(in0 + (uint32_t)s * PIPE_TILE_BYTES + (uint32_t)r * row_bytes,
                      gin + (size_t)tile * tile_bytes + (size_t)r * row_bytes, half,
                      full0 + 8 * s);
        }
      };
      auto prefetch = [&](int i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        if (ARZ) {
          bulk_prefetch_l2(gin + (size_t)tile * tile_bytes, bytes_of(tile));
        } else {
          const int envs = tile == last_tile ? last_envs : a.epc;
          for (int r = 0; r < envs; ++r)
            bulk_prefetch_l2(gin + (size_t)tile * tile_bytes + (size_t)r * row_bytes,
                             row_bytes >> 1);
        }
      };
      const int pre = n_my < PIPE_STAGES ? n_my : PIPE_STAGES;
      for (int i = 0; i < pre; ++i) issue_load(i);
      const int ahead = a.l2_ahead;
      for (int i = PIPE_STAGES; i < PIPE_STAGES + ahead && i < n_my; ++i) prefetch(i);
      for (int i = 0; i < n_my; ++i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int s = i % PIPE_STAGES, o = i % PIPE_OUT;
        if (ahead > 0 && i + PIPE_STAGES + ahead < n_my) prefetch(i + PIPE_STAGES + ahead);
        if (i + PIPE_STAGES < n_my) {  // refill stage s once the consumers have read tile i
          mbar_wait_relaxed(empty0 + 8 * s, (uint32_t)((i / PIPE_STAGES) & 1), PRODUCER_NAP_NS);
          issue_load(i + PIPE_STAGES);
        }
        mbar_wait_relaxed(ofull0 + 8 * o, (uint32_t)((i / PIPE_OUT) & 1), PRODUCER_NAP_NS);
        bulk_store(gout + (size_t)tile * tile_bytes, out0 + (uint32_t)o * PIPE_TILE_BYTES,
                   bytes_of(tile));
        bulk_commit();
        // all but the newest PIPE_OUT-1 stores have finished reading shared memory
        bulk_wait_read<PIPE_OUT - 1>();
        if (i >= PIPE_OUT - 1) mbar_arrive(ofree0 + 8 * ((i - (PIPE_OUT - 1)) % PIPE_OUT));
      }
      bulk_wait_read<0>();  // shared memory must outlive the last bulk store
    }
    if constexpr (REDUCER) {
      if (t >= NC + 32) {
        // ===================== reducer warp ================================
        const int lane = t & 31;
        const unsigned full = 0xffffffffu;
        unsigned g = 0;   // steps seen so far (all tiles): buffer g & 1
        for (int i = 0; i < n_my; ++i) {
          const int tile = (int)blockIdx.x + i * (int)gridDim.x;
          const int envs = tile == last_tile ? last_envs : a.epc;
          for (int j = 0; j < a.n_inner; ++j, ++g) {
            const unsigned b = g & 1u;
            // sleep until all NC consumer threads have left this step's pairs
            asm volatile("bar.sync %0, %1;" ::"r"(2u + b), "r"(NC + 32) : "memory");
            for (int el = 0; el < envs; ++el) {
              const T* pr = pairbuf + 2 * ((int)b * NC + el * a.tpe);
              T n = 0, d = 0;
              for (int k = lane; k < a.tpe; k += 32) {
                n += pr[2 * k];
                d += pr[2 * k + 1];
              }
              const bool odd = (lane & 1) != 0;
              T acc = (odd ? d : n) + __shfl_xor_sync(full, odd ? n : d, 1);
              for (int off = 2; off <= 16; off <<= 1) acc += __shfl_xor_sync(full, acc, off);
              const T dsum = __shfl_sync(full, acc, 1);
              if (lane == 0)
                a.reward[(size_t)j * (size_t)a.reward_stride + (size_t)tile * a.epc + el] = acc / dsum;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(rfree0 + 8 * b);   // buffer b may be overwritten
          }
        }
      }
    }
    return;
  }

  // ======================= consumers =======================================
  // thread -> (road in tile, position in road); fixed for the kernel's life
  const int el = t / a.tpe;
  const int pos = t - el * a.tpe;
  const bool row_first = (pos == 0), row_last = (pos == a.nthr - 1);
  const bool owns_cells = (pos < a.nthr) && (el < a.epc);
  const int xi = t + el;  // exchange slot (one pad slot per road)
  const bool lane0 = (t & 31) == 0;
  const bool want_reward = a.reward != nullptr;
  // byte offsets of my first rho vector inside a stage, and of the v block
  const uint32_t my_off = (uint32_t)el * row_bytes + (uint32_t)pos * (U * 16);
  const uint32_t v_off = (uint32_t)N * (uint32_t)sizeof(T);

  EnvC<T> e;
  if (UNIFORM) {
    load_uniform(e, a.u, 0, (const T*)nullptr, a.step);
    if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
  }

  int s = 0, o = 0, xpar = 0;
  uint32_t ph_in = 0, ph_out = 0;
  const int n_inner = a.n_inner;
  unsigned gstep = 0;   // REDUCER: steps done so far (all tiles), as the reducer counts them
  for (int i = 0; i < n_my; ++i) {
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int envs_here = tile == last_tile ? last_envs : a.epc;
    const int env = tile * a.epc + el;
    const bool active = owns_cells && (el < envs_here);

    if (active && !UNIFORM) {
      // per-road constants: issued before the wait so the loads overlap it
      load_env(e, a.table, a.flags, env, (const T*)nullptr, a.step);
      if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
    }

    mbar_wait(full0 + 8 * s, ph_in);

    T rho[CPT], v[CPT], y[CPT], r[CPT], D[CPT], S[CPT];
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
    }

    T num = 0, den = 0;
    // a reward row per inner step (mt_rollout with rewards): each warp leaves its
    // sums in r_num/r_den[j & 1]; the road's first thread finishes step j after
    // the exchange barrier of step j + 1 (or after the loop), so the per-step
    // reduction costs no barrier of its own
    const bool every = want_reward && a.reward_stride > 0;
    // Runs in the warp holding the road's first thread, on all of its lanes
    // (the sums are shared-memory broadcasts): straight-line code that the
    // scheduler can interleave with the step's own arithmetic, instead of one
    // lane walking through a divergent branch while the CTA waits for its warp.
    auto reward_finish = [&](int j) {
      const int w0 = t >> 5, nw = a.tpe >> 5;   // (roads wider than a warp only)
      T n = sm.r_num[j & 1][w0], d = sm.r_den[j & 1][w0];
      for (int w = 1; w < nw; ++w) { n += sm.r_num[j & 1][w0 + w]; d += sm.r_den[j & 1][w0 + w]; }
      using AX = Ar<T, false>;
      T q = AX::div_unchecked(n, d);
      if (!AX::div_quick(n, d)) q = n / d;      // operands outside the inline expansion's range
      if (active && row_first) a.reward[(size_t)j * (size_t)a.reward_stride + env] = q;
    };
    // this step's speed limit is fetched one step ahead: its L2 latency then
    // overlaps a whole step instead of stalling the start of pass 1
    T act_next = T(0);
    if (active && a.action != nullptr) act_next = __ldg(a.action + env);
    // n_inner model steps on this tile; between them the state lives in
    // registers and only the edge values cross threads
    for (int j = 0; j < n_inner; ++j) {
      const bool last = (j == n_inner - 1);
      T* const xS = sm.xS[xpar];
      T* const xD = sm.xD[xpar];
      T* const xR = sm.xR[xpar];
      xpar ^= 1;
      if (active) {
        if (a.action != nullptr) {   // this step's speed limit for my road
          e.v_max = act_next;
          if (j + 1 < n_inner) act_next = __ldg(a.action + (size_t)(j + 1) * a.action_stride + env);
          if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
        }
        with_lam<LAMK>(e.flags, [&](auto kind) {
          constexpr int K = decltype(kind)::value;
          if (ARZ) arz_pass1<T, FAST, K, CPT>(e, rho, v, y, r, D, S);
          else lwr_pass1<T, FAST, K, CPT>(e, rho, D, S);
        });
        if (ARZ) xR[xi + 1] = r[CPT - 1];
        xS[xi] = S[0];
        xD[xi + 1] = D[CPT - 1];
        if (row_last) {
          xS[xi + 1] = S[CPT - 1];  // arz.py:405 / lwr.py:294
          // OLD boundary-source cells for the LOOP rule of the road's first
          // thread travel in the road's pad slots (whose flux is discarded)
          xD[xi - pos] = ARZ ? rho[CPT - 1] : rho[CPT - 2];
          xR[xi - pos] = ARZ ? v[CPT - 1] : rho[CPT - 1];
        }
      }
      T rn[CPT], vn[CPT], Q[CPT], P[CPT];
      if (active) {  // interior cells need nothing from other threads: finish them first
        with_lam<LAMK>(e.flags, [&](auto kind) {
          cell_mid<T, ARZ, FAST, decltype(kind)::value, CPT>(e, rho, y, r, D, S, Q, P, rn, vn);
        });
      }
      consumer_sync(NC);                               // exchange visible
      if (j == 0 && lane0) mbar_arrive(empty0 + 8 * s);  // stage s fully read: release it
      if (last) mbar_wait(ofree0 + 8 * o, ph_out ^ 1);   // output stage drained by its last store
      if constexpr (!REDUCER) {
        if (every && a.tpe > 32 && j > 0 && pos < 32) reward_finish(j - 1);
      }

      if (active) {
        const T S_right = xS[xi + 1], D_left = xD[xi], r_left = xR[xi];
        with_lam<LAMK>(e.flags, [&](auto kind) {
          cell_edges<T, ARZ, FAST, decltype(kind)::value, CPT>(
              e, a.bc, row_first, row_last, rho, v, y, r, D, S, Q, P, S_right, D_left, r_left,
              D_left, r_left, rn, vn);
        });
        if (!last) {
#pragma unroll
          for (int k = 0; k < CPT; ++k) { rho[k] = rn[k]; v[k] = vn[k]; }
        } else {
          if (want_reward && !every) {
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
            sts16(dst + v_off + 16 * u, tv);
          }
        }
      }
      if (every) {
        // this step's sum(rho v) and sum(rho) over my cells, then over the warp
        // (or over the road's lanes when a road is narrower than a warp): the
        // two sums share one butterfly, see step_pipe_kernel
        T pn = 0, pd = 0;
        if (active) {
#pragma unroll
          for (int k = 0; k < CPT; ++k) { pn += rn[k] * vn[k]; pd += rn[k]; }
        }
        const unsigned full = 0xffffffffu;
        const int width = a.tpe <= 32 ? a.tpe : 32;
        if constexpr (REDUCER) {
          // hand my sums to the reducer warp: wait (normally not at all) until it
          // has summed this buffer's previous contents, store, arrive without waiting
          const unsigned b = gstep & 1u;
          if (gstep >= 2) mbar_wait(rfree0 + 8 * b, ((gstep >> 1) - 1) & 1u);
          T* pr = pairbuf + 2 * ((int)b * NC + t);
          pr[0] = pn;
          pr[1] = pd;
          asm volatile("bar.arrive %0, %1;" ::"r"(2u + b), "r"(NC + 32) : "memory");
          ++gstep;
        } else if (width >= 2) {
          const bool odd = (t & 1) != 0;
          T acc = (odd ? pd : pn) + __shfl_xor_sync(full, odd ? pn : pd, 1);
          for (int off = 2; off < width; off <<= 1) acc += __shfl_xor_sync(full, acc, off);
          if (a.tpe <= 32) {
            // lanes 0 / 1 of the road's group hold its two sums: hand the
            // denominator to the first lane, which stores the reward itself
            const T dsum = __shfl_down_sync(full, acc, 1);
            if (active && row_first)
              a.reward[(size_t)j * (size_t)a.reward_stride + env] = acc / dsum;
          } else {
            if ((t & 31) == 0) sm.r_num[j & 1][t >> 5] = acc;
            if ((t & 31) == 1) sm.r_den[j & 1][t >> 5] = acc;
          }
        } else if (active) {
          a.reward[(size_t)j * (size_t)a.reward_stride + env] = pn / pd;
        }
      }
    }
    fence_proxy_async_smem();  // my generic-proxy writes -> visible to the bulk engine
    __syncwarp();
    if (lane0) mbar_arrive(ofull0 + 8 * o);          // non-blocking hand-off to the producer

    if constexpr (!REDUCER) {
      if (every && a.tpe > 32) {                       // the last step's reward row
        consumer_sync(NC);
        if (pos < 32) reward_finish(n_inner - 1);
      }
    }
    if (want_reward && !every) {
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
    if (++s == PIPE_STAGES) { s = 0; ph_in ^= 1; }
    if (++o == PIPE_OUT) { o = 0; ph_out ^= 1; }
  }
}

}  // namespace mt
