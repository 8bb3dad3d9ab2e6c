// mt_pipe.cuh -- persistent, TMA-pipelined step kernel for batched roads.
//
// One CTA per SM slot (grid = min(#tiles, SMs x CTAs/SM)) walks over tiles of
// consecutive roads.  A tile (<= 16 KB of [rho | v] rows, contiguous in HBM)
// is fetched with ONE bulk asynchronous copy (cp.async.bulk, the 1-D TMA path)
// into a ring of STAGES shared-memory buffers, each guarded by an mbarrier;
// the finished tile is staged in shared memory and written back with one bulk
// store.  Loads for the next STAGES-1 tiles and the stores of the previous
// tiles are in flight while the CTA computes, so HBM latency is decoupled
// from the FP64 work instead of being exposed once per CTA.
//
// Per tile: wait(full[s]) -> every thread reads its 16-byte vectors of rho and
// v from shared memory and derives the per-cell flux ingredients -> edge
// values go through a small exchange array (__syncthreads A) -> update,
// boundary rule, (rho,y)->(rho,v), results to the output stage ->
// fence.proxy.async + __syncthreads B -> thread 0 issues the bulk store and
// re-arms stage s with the load of tile i+STAGES.
#pragma once
#include <stdint.h>
#include "mtraffic.h"
#include "mt_math.cuh"

namespace mt {

constexpr int PIPE_BLOCK = 512;
constexpr int PIPE_STAGES = 4;
constexpr int PIPE_OUT = 2;
constexpr int PIPE_TILE_BYTES = PIPE_BLOCK * 32;  // 16 KB: one 16 B vector of rho and of v per thread

// ---- PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                          uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global, tracked by the per-thread bulk async-group
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ---- shared-memory layout ---------------------------------------------------
template <typename T> struct PipeSmem {
  alignas(128) unsigned char in[PIPE_STAGES][PIPE_TILE_BYTES];
  alignas(128) unsigned char out[PIPE_OUT][PIPE_TILE_BYTES];
  T xS[PIPE_BLOCK], xD[PIPE_BLOCK], xR[PIPE_BLOCK];  // edge exchange
  T s_num[PIPE_BLOCK / 32], s_den[PIPE_BLOCK / 32];
  alignas(8) uint64_t full[PIPE_STAGES];
};

template <typename T> struct PipeArgs {
  const T* in;
  T* out;
  const T* table;
  const int* flags;
  const T* action;
  T* reward;   // nullptr: no reward reduction
  long long B;
  int N;
  int nvec;    // N / V
  int tpe;     // threads per road (padded)
  int epc;     // roads per tile
  int n_tiles;
  int bc_l, bc_r;
  T step;
  T bl0, bl1, br0, br1;
  UniformC<T> u;  // used by the UNIFORM instantiations
};

// MODEL: 0 = LWR, 1 = ARZ.  LAMK: exponent kind fixed at compile time, or
// LAM_RUNTIME.  UNIFORM: every road shares the scalar parameters in a.u.
template <typename T, int MODEL, bool FAST, int LAMK, bool UNIFORM>
__global__ void __launch_bounds__(PIPE_BLOCK, 2) step_pipe_kernel(const PipeArgs<T> a) {
  constexpr int V = Vec<T>::N;
  constexpr bool ARZ = (MODEL == 1);
  using VT = typename Vec<T>::type;
  using A = Ar<T, FAST>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PipeSmem<T>& sm = *reinterpret_cast<PipeSmem<T>*>(smem_raw);

  const int t = threadIdx.x;
  const int N = a.N;
  const uint32_t row_bytes = 2u * (uint32_t)N * (uint32_t)sizeof(T);
  const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool want_reward = a.reward != nullptr;

  auto tile_envs = [&](int tile) -> int {
    const long long left = a.B - (long long)tile * a.epc;
    return left < a.epc ? (int)left : a.epc;
  };
  auto issue_load = [&](int i) {  // thread 0 only
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int s = i % PIPE_STAGES;
    const uint32_t bytes = (uint32_t)tile_envs(tile) * row_bytes;
    mbar_expect_tx(&sm.full[s], bytes);
    bulk_load(sm.in[s], reinterpret_cast<const unsigned char*>(a.in) +
                            (size_t)tile * a.epc * row_bytes, bytes, &sm.full[s]);
  };

  if (t == 0) {
    for (int s = 0; s < PIPE_STAGES; ++s) mbar_init(&sm.full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (t == 0) {
    const int pre = n_my < PIPE_STAGES ? n_my : PIPE_STAGES;
    for (int i = 0; i < pre; ++i) issue_load(i);
  }

  // thread -> (road in tile, vector in road); fixed for the kernel's life
  const int el = t / a.tpe;
  const int vec = t - el * a.tpe;
  const bool row_first = (vec == 0), row_last = (vec == a.nvec - 1);
  const size_t row_off = (size_t)el * 2 * N;

  EnvC<T> e;
  if (UNIFORM) {
    load_uniform(e, a.u, 0, (const T*)nullptr, a.step);
    if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
  }

  for (int i = 0; i < n_my; ++i) {
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int s = i % PIPE_STAGES;
    const int o = i % PIPE_OUT;
    const int envs_here = tile_envs(tile);
    const long long env = (long long)tile * a.epc + el;
    const bool active = (el < envs_here) && (vec < a.nvec);

    if (active && (!UNIFORM || a.action != nullptr)) {
      // per-road constants (and/or this step's speed limit): issued before the
      // wait so the loads overlap it
      if (UNIFORM) load_uniform(e, a.u, env, a.action, a.step);
      else load_env(e, a.table, a.flags, env, a.action, a.step);
      if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
    }

    mbar_wait(&sm.full[s], (uint32_t)((i / PIPE_STAGES) & 1));

    T rho[V], v[V], y[V], r[V], D[V], S[V];
    T loop_a = 0, loop_b = 0;  // old boundary-source cells for the LOOP rule
    if (active) {
      const T* row = reinterpret_cast<const T*>(sm.in[s]) + row_off;
      unpack(reinterpret_cast<const VT*>(row)[vec], rho);
      if (ARZ) {
        unpack(reinterpret_cast<const VT*>(row + N)[vec], v);
        bool ok = true;
#pragma unroll
        for (int k = 0; k < V; ++k) {
          arz_cell_nodiv<T, FAST, LAMK>(e, rho[k], v[k], y[k], D[k], S[k]);
          r[k] = A::div_unchecked(y[k], rho[k]);           // y / rho (arz.py:411)
          ok &= A::div_safe(y[k], rho[k]);
        }
        if (!ok) {
#pragma unroll
          for (int k = 0; k < V; ++k) r[k] = A::div(y[k], rho[k]);
        }
        sm.xR[t] = r[V - 1];
        if (row_first && a.bc_l == MT_BC_LOOP) { loop_a = row[N - 1]; loop_b = row[2 * N - 1]; }
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) lwr_cell<T, FAST, LAMK>(e, rho[k], D[k], S[k]);
        if (row_first && a.bc_l == MT_BC_LOOP) { loop_a = row[N - 2]; loop_b = row[N - 1]; }
      }
      sm.xS[t] = S[0];
      sm.xD[t] = D[V - 1];
    }
    // the output stage about to be overwritten must have been drained by the
    // bulk store issued PIPE_OUT tiles ago
    if (t == 0) bulk_wait_read<PIPE_OUT - 1>();
    __syncthreads();  // A

    T num = 0, den = 0;
    if (active) {
      const T S_right = row_last ? S[V - 1] : sm.xS[t + 1];  // arz.py:405 / lwr.py:294
      T Q[V];
#pragma unroll
      for (int k = 0; k < V - 1; ++k) Q[k] = npmin(D[k], S[k + 1]);  // arz.py:408 / lwr.py:297
      Q[V - 1] = npmin(D[V - 1], S_right);
      T Qm = row_first ? Q[0] : npmin(sm.xD[t - 1], S[0]);           // arz.py:414 / lwr.py:231
      T rn[V], vn[V];
      if (ARZ) {
        T P[V], yn[V];
#pragma unroll
        for (int k = 0; k < V; ++k) P[k] = A::mul(Q[k], r[k]);       // arz.py:411
        T Pm = row_first ? P[0] : A::mul(Qm, sm.xR[t - 1]);          // arz.py:415
#pragma unroll
        for (int k = 0; k < V; ++k) {
          arz_advance<T, FAST>(e, rho[k], y[k], Qm, Q[k], Pm, P[k], rn[k], yn[k]);
          Qm = Q[k];
          Pm = P[k];
        }
        // boundary cells (arz.py:316-360)
        if (row_first) {
          if (a.bc_l == MT_BC_LOOP) { rn[0] = loop_a; yn[0] = arz_y<T, FAST, LAMK>(e, loop_a, loop_b); }
          else if (a.bc_l == MT_BC_EXTEND) { rn[0] = rn[1]; yn[0] = yn[1]; }
          else if (a.bc_l == MT_BC_CONSTANT) { rn[0] = a.bl0; yn[0] = a.bl1; }
          else { rn[0] = rho[0]; yn[0] = y[0]; }
        }
        if (row_last) {
          if (a.bc_r == MT_BC_LOOP) { rn[V - 1] = rho[V - 2]; yn[V - 1] = y[V - 2]; }
          else if (a.bc_r == MT_BC_EXTEND) { rn[V - 1] = rn[V - 2]; yn[V - 1] = yn[V - 2]; }
          else if (a.bc_r == MT_BC_CONSTANT) { rn[V - 1] = a.br0; yn[V - 1] = a.br1; }
          else { rn[V - 1] = rho[V - 1]; yn[V - 1] = y[V - 1]; }
        }
        // u(): zero patch, y/rho + V(rho)  (arz.py:258-277), divisions batched
        T qn[V];
        bool ok = true;
#pragma unroll
        for (int k = 0; k < V; ++k) {
          if (rn[k] == T(0)) rn[k] = T(0.0000000001);
          qn[k] = A::div_unchecked(yn[k], rn[k]);
          ok &= A::div_safe(yn[k], rn[k]);
        }
        if (!ok) {
#pragma unroll
          for (int k = 0; k < V; ++k) qn[k] = A::div(yn[k], rn[k]);
        }
#pragma unroll
        for (int k = 0; k < V; ++k) vn[k] = A::add(qn[k], v_eq<T, FAST, LAMK>(rn[k], e));
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          rn[k] = lwr_advance<T, FAST>(e, rho[k], Q[k], Qm);
          Qm = Q[k];
        }
        // boundary cells (lwr.py:237-260); EXTEND keeps the scheme's values
        const T second_last = rn[V - 2];
        if (row_first) {
          if (a.bc_l == MT_BC_LOOP) {  // UPDATED value of the last cell
            T D2, S2, D1, S1;
            lwr_cell<T, FAST, LAMK>(e, loop_a, D2, S2);
            lwr_cell<T, FAST, LAMK>(e, loop_b, D1, S1);
            rn[0] = lwr_advance<T, FAST>(e, loop_b, npmin(D1, S1), npmin(D2, S1));
          } else if (a.bc_l == MT_BC_CONSTANT) { rn[0] = a.bl0; }
          else if (a.bc_l == MT_BC_HALO) { rn[0] = rho[0]; }
        }
        if (row_last) {
          if (a.bc_r == MT_BC_LOOP) rn[V - 1] = second_last;
          else if (a.bc_r == MT_BC_CONSTANT) rn[V - 1] = a.br0;
          else if (a.bc_r == MT_BC_HALO) rn[V - 1] = rho[V - 1];
        }
#pragma unroll
        for (int k = 0; k < V; ++k) vn[k] = v_eq<T, FAST, LAMK>(rn[k], e);  // lwr.py:201
      }
      if (want_reward) {
#pragma unroll
        for (int k = 0; k < V; ++k) { num += rn[k] * vn[k]; den += rn[k]; }
      }
      T* orow = reinterpret_cast<T*>(sm.out[o]) + row_off;
      reinterpret_cast<VT*>(orow)[vec] = pack(rn);
      reinterpret_cast<VT*>(orow + N)[vec] = pack(vn);
    }
    fence_proxy_async_smem();  // my generic-proxy writes -> visible to the bulk engine
    __syncthreads();           // B
    if (t == 0) {
      bulk_store(reinterpret_cast<unsigned char*>(a.out) + (size_t)tile * a.epc * row_bytes,
                 sm.out[o], (uint32_t)envs_here * row_bytes);
      bulk_commit();
      if (i + PIPE_STAGES < n_my) issue_load(i + PIPE_STAGES);
    }
    if (want_reward) {
      // per-road sum(rho v)/sum(rho); same deterministic order as the direct kernel
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
  }
  if (t == 0) bulk_wait_read<0>();  // shared memory must outlive the last bulk store
}

}  // namespace mt
