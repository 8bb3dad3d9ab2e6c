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
#include "mt_cell.cuh"

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
constexpr int PIPE_MAX_EPC = 64;                       // roads per tile (cap)
constexpr int PIPE_XCH = PIPE_BLOCK + PIPE_MAX_EPC + 8;  // one pad slot per road in the tile

template <typename T> struct PipeSmem {
  alignas(128) unsigned char in[PIPE_STAGES][PIPE_TILE_BYTES];
  alignas(128) unsigned char out[PIPE_OUT][PIPE_TILE_BYTES];
  T xS[PIPE_XCH], xD[PIPE_XCH], xR[PIPE_XCH];  // edge exchange
  T s_num[PIPE_BLOCK / 32], s_den[PIPE_BLOCK / 32];
  alignas(8) uint64_t full[PIPE_STAGES];
};

static_assert(sizeof(PipeSmem<double>) <= 113 * 1024, "two CTAs per SM must fit in 227 KB");

template <typename T> struct PipeArgs {
  const T* in;
  T* out;
  const T* table;
  const int* flags;
  const T* action;
  T* reward;   // nullptr: no reward reduction
  long long B;
  int N;
  int nthr;    // threads that own cells in one road: N / CPT
  int tpe;     // threads per road (padded)
  int epc;     // roads per tile
  int n_tiles;
  int block;   // threads per CTA
  BcArgs<T> bc;
  T step;
  UniformC<T> u;  // used by the UNIFORM instantiations
};

// MODEL: 0 = LWR, 1 = ARZ.  LAMK: exponent kind fixed at compile time, or
// LAM_RUNTIME.  UNIFORM: every road shares the scalar parameters in a.u.
// U: 16-byte vectors (of each field) per thread; a thread owns U*V cells.
template <typename T, int MODEL, bool FAST, int LAMK, bool UNIFORM, int U>
__global__ void __launch_bounds__(PIPE_BLOCK / U, 2) step_pipe_kernel(const PipeArgs<T> a) {
  constexpr int V = Vec<T>::N;
  constexpr int CPT = U * V;
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

  // pad slots of the exchange arrays stay zero for the kernel's life: they
  // feed the (discarded) left-face flux of each road's first cell
  for (int i = t; i < PIPE_XCH; i += blockDim.x) { sm.xD[i] = T(0); sm.xR[i] = T(0); }
  if (t == 0) {
    for (int s = 0; s < PIPE_STAGES; ++s) mbar_init(&sm.full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (t == 0) {
    const int pre = n_my < PIPE_STAGES ? n_my : PIPE_STAGES;
    for (int i = 0; i < pre; ++i) issue_load(i);
  }

  // thread -> (road in tile, position in road); fixed for the kernel's life
  const int el = t / a.tpe;
  const int pos = t - el * a.tpe;
  const bool row_first = (pos == 0), row_last = (pos == a.nthr - 1);
  const size_t row_off = (size_t)el * 2 * N;
  const int xi = t + el;  // exchange slot (one pad slot per road)

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
    const bool active = (el < envs_here) && (pos < a.nthr);

    if (active && (!UNIFORM || a.action != nullptr)) {
      // per-road constants (and/or this step's speed limit): issued before the
      // wait so the loads overlap it
      if (UNIFORM) load_uniform(e, a.u, env, a.action, a.step);
      else load_env(e, a.table, a.flags, env, a.action, a.step);
      if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
    }

    mbar_wait(&sm.full[s], (uint32_t)((i / PIPE_STAGES) & 1));

    T rho[CPT], v[CPT], y[CPT], r[CPT], D[CPT], S[CPT];
    T loop_a = 0, loop_b = 0;  // old boundary-source cells for the LOOP rule
    if (active) {
      const T* row = reinterpret_cast<const T*>(sm.in[s]) + row_off;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T tmp[V];
        unpack(reinterpret_cast<const VT*>(row)[pos * U + u], tmp);
#pragma unroll
        for (int k = 0; k < V; ++k) rho[u * V + k] = tmp[k];
        if (ARZ) {
          unpack(reinterpret_cast<const VT*>(row + N)[pos * U + u], tmp);
#pragma unroll
          for (int k = 0; k < V; ++k) v[u * V + k] = tmp[k];
        }
      }
      if (ARZ) {
        arz_pass1<T, FAST, LAMK, CPT>(e, rho, v, y, r, D, S);
        sm.xR[xi + 1] = r[CPT - 1];
        if (row_first) { loop_a = row[N - 1]; loop_b = row[2 * N - 1]; }
      } else {
        lwr_pass1<T, FAST, LAMK, CPT>(e, rho, D, S);
        if (row_first) { loop_a = row[N - 2]; loop_b = row[N - 1]; }
      }
      sm.xS[xi] = S[0];
      sm.xD[xi + 1] = D[CPT - 1];
      if (row_last) sm.xS[xi + 1] = S[CPT - 1];  // arz.py:405 / lwr.py:294
    }
    // the output stage about to be overwritten must have been drained by the
    // bulk store issued PIPE_OUT tiles ago
    if (t == 0) bulk_wait_read<PIPE_OUT - 1>();
    __syncthreads();  // A

    T num = 0, den = 0;
    if (active) {
      T rn[CPT], vn[CPT];
      if (ARZ)
        arz_pass2<T, FAST, LAMK, CPT>(e, a.bc, row_first, row_last, rho, v, y, r, D, S,
                                      sm.xS[xi + 1], sm.xD[xi], sm.xR[xi], loop_a, loop_b, rn, vn);
      else
        lwr_pass2<T, FAST, LAMK, CPT>(e, a.bc, row_first, row_last, rho, D, S, sm.xS[xi + 1],
                                      sm.xD[xi], loop_a, loop_b, rn, vn);
      if (want_reward) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) { num += rn[k] * vn[k]; den += rn[k]; }
      }
      T* orow = reinterpret_cast<T*>(sm.out[o]) + row_off;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T tr[V], tv[V];
#pragma unroll
        for (int k = 0; k < V; ++k) { tr[k] = rn[u * V + k]; tv[k] = vn[u * V + k]; }
        reinterpret_cast<VT*>(orow)[pos * U + u] = pack(tr);
        reinterpret_cast<VT*>(orow + N)[pos * U + u] = pack(tv);
      }
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
  }
  if (t == 0) bulk_wait_read<0>();  // shared memory must outlive the last bulk store
}

}  // namespace mt
