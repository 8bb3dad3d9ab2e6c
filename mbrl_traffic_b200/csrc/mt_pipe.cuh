// mt_pipe.cuh -- persistent, warp-specialised, TMA-pipelined step kernel for
// batched roads (the hot kernel of the engine).
//
// Grid: min(#tiles, SMs x 2) CTAs, each walking over tiles of consecutive
// roads.  A tile (<= 16 KB of [rho | v] rows, contiguous in HBM) is fetched
// with ONE bulk asynchronous copy (cp.async.bulk, the 1-D TMA path) into a
// ring of PIPE_STAGES shared-memory buffers; the finished tile is staged in one
// of PIPE_OUT shared-memory buffers and written back with one bulk store.
//
// Roles inside a CTA (NC consumer threads + one producer warp):
//   producer (1 thread): re-arms input stages (cp.async.bulk + expect_tx) as
//     soon as the consumers release them, issues the bulk store of a finished
//     tile, and hands output stages back once the store has read them.
//   consumers: wait(full[s]) -> read their cells from shared memory, derive
//     the per-cell flux ingredients, swap three edge values through a small
//     exchange array (one named barrier among the consumers), release the
//     input stage, wait(out_free[o]), finish the update and write the result
//     to the output stage, then arrive on out_full[o] WITHOUT blocking and
//     move on to the next tile.
//     (The interior cells of a thread are finished BEFORE the barrier: they
//     need nothing from other threads -- cell_mid / cell_edges in mt_cell.cuh.)
// Loads for the next PIPE_STAGES-1 tiles and the stores of previous tiles are
// in flight while the consumers compute, so HBM latency is decoupled from the
// FP64 work; all hand-offs are mbarriers (no CTA-wide __syncthreads in the
// steady state).
//
// Tile schedule: static round-robin, or (HBM-bound instantiations, large
// batches) tiles drawn from a per-engine counter -- see PipeDynamic.
// Launch boundary: programmatic dependent launch; before waiting for the
// previous launch the producer prefetches its first tiles towards L2, and with
// MT_FLAG_INDEPENDENT_INPUT it issues the first ring-full of loads outright.
#pragma once
#include <stdint.h>
#include <cuda_bf16.h>
#include "mtraffic.h"
#include "mt_math.cuh"
#include "mt_cell.cuh"

namespace mt {

// four values -> four bfloat16 (round to nearest even, through f32 like a
// torch .to(bfloat16)) packed in 8 bytes
template <typename T>
__device__ __forceinline__ uint2 pack_bf16x4(T a, T b, T c, T d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn((float)a, (float)b);
  const __nv_bfloat162 hi = __floats2bfloat162_rn((float)c, (float)d);
  uint2 out;
  out.x = *reinterpret_cast<const uint32_t*>(&lo);
  out.y = *reinterpret_cast<const uint32_t*>(&hi);
  return out;
}

#ifndef MT_PIPE_CTAS
#define MT_PIPE_CTAS 2                            // resident CTAs per SM the kernel is sized for
#endif
constexpr int PIPE_CTAS = MT_PIPE_CTAS;
constexpr int PIPE_BLOCK = 512;                   // consumer threads at U = 1
#ifdef MT_PIPE_OUT_STAGES
constexpr int PIPE_OUT = MT_PIPE_OUT_STAGES;      // (experiments)
#else
constexpr int PIPE_OUT = (PIPE_CTAS >= 3) ? 1 : 2;
#endif
constexpr int PIPE_TILE_BYTES = PIPE_BLOCK * 32;  // 16 KB
constexpr int PIPE_MAX_EPC = 64;                  // roads per tile (cap)
// input stages: as many as fit next to the exchange arrays with two CTAs per SM
template <typename T, int U> struct PipeCfg {
  static constexpr int NC = PIPE_BLOCK / U;                     // consumer threads
  static constexpr int XCH = NC + PIPE_MAX_EPC + 8;             // one pad slot per road
#ifdef MT_PIPE_IN_STAGES
  static constexpr int STAGES = MT_PIPE_IN_STAGES;              // (experiments)
#else
  static constexpr int STAGES = (PIPE_CTAS >= 3) ? 2 : (sizeof(T) == 8 && U == 1) ? 3 : 4;
#endif
};

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
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait parks the warp in hardware for up to this long before it returns
// false, instead of spinning through issue slots other warps could use
#ifndef MT_MBAR_HINT_NS
#define MT_MBAR_HINT_NS 2000
#endif
constexpr uint32_t MBAR_SUSPEND_HINT_NS = MT_MBAR_HINT_NS;
#ifndef MT_PRODUCER_NAP_NS
#define MT_PRODUCER_NAP_NS 100
#endif
constexpr unsigned PRODUCER_NAP_NS = MT_PRODUCER_NAP_NS;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity), "r"(MBAR_SUSPEND_HINT_NS)
      : "memory");
}
// The producer thread sits alone in its warp: between polls it sleeps, so that
// its spinning does not take issue slots from the consumer warps that share
// its scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, unsigned ns) {
#ifndef MT_PRODUCER_POLLS
  // try_wait with a suspend-time hint compiles to NANOSLEEP.SYNCS: the thread
  // sleeps until the barrier flips (or the hint expires) and issues nothing
  // meanwhile.  A __nanosleep() poll loop returned within ~15 ns per poll on
  // B200 and cost 15% of all issued warp instructions (profiles/, round 1).
  (void)ns;
  mbar_wait(bar, parity);
  return;
#endif
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(ns);
  }
}
// global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src_gmem, uint32_t bytes,
                                          uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst_smem),
      "l"(src_gmem), "r"(bytes), "r"(bar)
      : "memory");
}
// the same load with an L2 eviction-priority hint (createpolicy operand)
__device__ __forceinline__ void bulk_load_hint(uint32_t dst_smem, const void* src_gmem,
                                               uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src_gmem), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// shared -> global, tracked by the per-thread bulk async-group
__device__ __forceinline__ void bulk_store(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// global -> L2 only: warms the line for a later bulk_load without holding
// shared memory, so HBM latency is covered beyond what the stage ring can hold
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// completion (not just the shared-memory read) of all but the newest N bulk stores
template <int N> __device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// Programmatic dependent launch: let the next kernel in the stream start its
// prologue while this one drains, and block until the previous kernel's
// writes are visible.  Both are no-ops for an ordinary launch.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_prior_grid() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
// barrier among the consumer threads only (id 1; id 0 is __syncthreads)
__device__ __forceinline__ void consumer_sync(int nc) {
  asm volatile("bar.sync 1, %0;" ::"r"(nc) : "memory");
}
// 16-byte shared-memory accesses on 32-bit shared addresses
__device__ __forceinline__ void lds16(uint32_t addr, double (&a)[2]) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a[0]), "=d"(a[1]) : "r"(addr));
}
__device__ __forceinline__ void lds16(uint32_t addr, float (&a)[4]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3])
               : "r"(addr));
}
__device__ __forceinline__ void sts16(uint32_t addr, const double (&a)[2]) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a[0]), "d"(a[1]) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, const float (&a)[4]) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a[0]), "f"(a[1]),
               "f"(a[2]), "f"(a[3])
               : "memory");
}

// ---- shared-memory layout ---------------------------------------------------
constexpr int PIPE_QMAX = 4;  // bound of the L2 look-ahead (power of two)
template <typename T, int U> struct PipeSmem {
  static constexpr int STAGES = PipeCfg<T, U>::STAGES;
  static constexpr int XCH = PipeCfg<T, U>::XCH;
  alignas(128) unsigned char in[STAGES][PIPE_TILE_BYTES];
  alignas(128) unsigned char out[PIPE_OUT][PIPE_TILE_BYTES];
  // edge exchange, double-buffered by tile parity: the consumers meet at one
  // barrier per tile, so a fast warp may already write tile i+1's edges while
  // a slow one still reads tile i's
  T xS[2][XCH], xD[2][XCH], xR[2][XCH];
  T s_num[PIPE_BLOCK / 32], s_den[PIPE_BLOCK / 32];
  // step_fused_kernel, per-step rewards: per-warp sums of the previous and the current step
  T r_num[2][PIPE_BLOCK / 32], r_den[2][PIPE_BLOCK / 32];
  alignas(8) uint64_t full[STAGES];         // producer -> consumers: tile landed
  uint64_t empty[STAGES];                   // consumers -> producer: stage read
  uint64_t out_full[PIPE_OUT];              // consumers -> producer: result staged
  uint64_t out_free[PIPE_OUT];              // producer -> consumers: store drained
  int tile_of[STAGES];                      // dynamic schedule: tile in each stage; -1: no more
  int drawn[PIPE_QMAX];                     // producer: tiles drawn + L2-prefetched, not yet loaded
};
// two CTAs per SM: 2 x (smem + 1 KB reserved) <= 228 KB
static_assert(sizeof(PipeSmem<double, 2>) <= (228 / PIPE_CTAS - 1) * 1024, "f64 U=2 smem");
static_assert(sizeof(PipeSmem<float, 2>) <= (228 / PIPE_CTAS - 1) * 1024, "f32 U=2 smem");
static_assert(PIPE_CTAS >= 3 || sizeof(PipeSmem<double, 1>) <= 113 * 1024, "f64 U=1 smem");
static_assert(PIPE_CTAS >= 3 || sizeof(PipeSmem<float, 1>) <= 113 * 1024, "f32 U=1 smem");
static_assert(PIPE_OUT <= 2, "r_num / r_den hold one set of warp sums per output stage");

#ifdef MT_TRACE
// experimental build only (tools/trace_pipe.py): per-CTA %globaltimer stamps
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}
#define MT_STAMP(slot) do { if (a.trace) a.trace[(size_t)blockIdx.x * 8 + (slot)] = gtime(); } while (0)
#else
#define MT_STAMP(slot) do {} while (0)
#endif

template <typename T> struct PipeArgs {
#ifdef MT_TRACE
  unsigned long long* trace = nullptr;
#endif
  const T* in;
  T* out;
  const T* table;
  const int* flags;
  const T* action;
  T* reward;   // nullptr: no reward reduction
  int B;
  int N;
  int nthr;    // threads that own cells in one road: N / CPT
  int tpe;     // threads per road (padded)
  int epc;     // roads per tile
  int n_tiles;
  int nc;      // consumer threads per CTA (the CTA has nc + 32 threads)
  int l2_ahead;  // tiles prefetched into L2 beyond the shared-memory ring
  int n_inner;   // model steps per tile visit (step_fused_kernel, mt_fused.cuh)
  long long action_stride;  // elements between consecutive steps' actions
  long long reward_stride;  // step_fused_kernel: > 0: a reward row per inner step, this far apart;
                            // 0: the last step's rewards only
  BcArgs<T> bc;
  T step;
  UniformC<T> u;  // used by the UNIFORM instantiations
  // Dynamic tile schedule of step_pipe_kernel: sched[0] hands out tiles beyond
  // each CTA's first ring-full, sched[1] counts finished CTAs (the last one
  // re-arms both).  nullptr: static round-robin schedule.
  unsigned int* sched = nullptr;
  // The input batch was not written by the launch this one programmatically
  // depends on: the first ring-full of loads is issued BEFORE waiting for that
  // launch to finish (only the stores wait), so this launch's pipeline fills
  // while the previous one drains.  See MT_FLAG_INDEPENDENT_INPUT.
  int early_loads = 0;
  int early_l2 = 2;  // tiles per CTA pulled towards L2 before the grid-dependency wait
  int l2_hint = 0;   // != 0: tile loads carry L2::evict_first (a tile is read once per step)
  // Per-road parameter tables (the non-UNIFORM instantiations): byte offset
  // inside an input stage at which the producer places the tile's table rows
  // (they fit behind the roads: the host picks epc accordingly); 0: the
  // consumers read their rows from global memory instead.
  int tbl_off = 0;
  // mt_step_obs16: a bfloat16 copy of the new state, (B, 2N) row-major, for a
  // policy network that reads the observation next (nullptr: none)
  uint16_t* obs16 = nullptr;
  // Rewards of roads wider than a warp, static schedule: the CTA is launched
  // with one more warp (288 threads are allocated as 10 warps anyway) that adds
  // up the per-thread sums the consumers leave in shared memory; the input ring
  // gives up its last stage for them.
  int reducer = 0;
};

// Which instantiations of step_pipe_kernel draw their tiles from a counter.
// Measured on B200: the HBM-bound ones gain (LWR f64 83% ->
// 99% of the measured peak, ARZ f32 fast 89% -> 93%); ARZ f64 is issue-bound,
// does not gain, and loses to the extra live state, so it keeps the static
// round-robin schedule.
template <typename T, int MODEL, bool FAST> struct PipeDynamic {
#ifdef MT_DYN_ALL
  static constexpr bool value = true;   // (experiments)
#else
  static constexpr bool value = !(MODEL == 1 && sizeof(T) == 8);
#endif
};

// MODEL: 0 = LWR, 1 = ARZ.  LAMK: exponent kind fixed at compile time, or
// LAM_RUNTIME.  UNIFORM: every road shares the scalar parameters in a.u.
// U: 16-byte vectors (of each field) per thread; a thread owns U*V cells.
// (288 threads are allocated as 10 warps: 96 registers per thread is what two
// CTAs per SM leave; a build with __maxnreg__(112) ran one CTA per SM, 35 us)
// REDUCER: the instantiation launched with the extra reducer warp (rewards of
// roads wider than a warp under the static schedule).  A separate instantiation
// so that the plain step keeps its compile-time stage count and its code: with
// the choice made at run time the headline step went from 23.0 to 26.2 us.
template <typename T, int MODEL, bool FAST, int LAMK, bool UNIFORM, int U, bool REDUCER = false>
__global__ void __launch_bounds__(PIPE_BLOCK / U + (REDUCER ? 64 : 32), PIPE_CTAS)
step_pipe_kernel(const PipeArgs<T> a) {
  constexpr int V = Vec<T>::N;
  constexpr int CPT = U * V;
  constexpr bool ARZ = (MODEL == 1);
  constexpr bool DYN = PipeDynamic<T, MODEL, FAST>::value;
  constexpr int PIPE_STAGES = PipeCfg<T, U>::STAGES;
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

  // reducer warp on: the last input stage holds the consumers' (sum rho v, sum rho) pairs
  static_assert(!REDUCER || !DYN, "the reducer warp follows the static tile schedule");
  constexpr bool reducer_on = REDUCER;   // (the host launches it only with rewards and tpe > 32)
  constexpr int n_stages = REDUCER ? PIPE_STAGES - 1 : PIPE_STAGES;
  T* const pairbuf = reinterpret_cast<T*>(sm.in[PIPE_STAGES - 1]);   // [PIPE_OUT][NC][2]

  if (t == 0) MT_STAMP(0);
#ifdef MT_TRACE
  if (t == 0 && a.trace) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    a.trace[(size_t)blockIdx.x * 8 + 7] = smid;
  }
#endif
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
      mbar_init(&sm.out_free[o], reducer_on ? 2u : 1u);   // the store drained (+ the pairs summed)
    }
    fence_mbar_init();
  }
  __syncthreads();
  // everything above overlapped the previous kernel's tail
  const bool early = a.early_loads != 0;
#ifndef MT_NO_EARLY_L2
  if (!early && t == NC) {
    // The loads themselves must wait for the previous launch (its output may be
    // this launch's input), but an L2 prefetch is not a read: it returns no
    // data, and L2 is the point of coherence, so lines the previous launch
    // still writes are simply updated in place.  The CTA's first tiles are
    // therefore pulled towards L2 while the previous launch drains (two tiles
    // measured best: more competes with that launch's own tail).
    const unsigned char* gin0 = reinterpret_cast<const unsigned char*>(a.in);
    for (int i = 0; i < a.early_l2; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      if (tile >= a.n_tiles) break;
      const int envs = tile == last_tile ? last_envs : a.epc;
      if (ARZ) {
        bulk_prefetch_l2(gin0 + (size_t)tile * tile_bytes, (uint32_t)envs * row_bytes);
      } else {
        for (int r = 0; r < envs; ++r)
          bulk_prefetch_l2(gin0 + (size_t)tile * tile_bytes + (size_t)r * row_bytes, row_bytes >> 1);
      }
    }
  }
#endif
  if (!early) pdl_wait_prior_grid();

  // ======================= producer warp ===================================
  if (t >= NC) {
    if (t == NC) {
      const unsigned char* gin = reinterpret_cast<const unsigned char*>(a.in);
      unsigned char* gout = reinterpret_cast<unsigned char*>(a.out);
      auto bytes_of = [&](int tile) -> uint32_t {
        return tile == last_tile ? (uint32_t)last_envs * row_bytes : tile_bytes;
      };
      const bool hint_ld = a.l2_hint != 0;
      const uint64_t pol_ld = hint_ld ? l2_policy_evict_first() : 0;
      auto tile_load = [&](uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
        if (hint_ld) bulk_load_hint(dst, src, bytes, bar, pol_ld);
        else bulk_load(dst, src, bytes, bar);
      };
      // the tile's rows of the per-road parameter table travel with it (no
      // eviction hint: the table is read again by every step)
      constexpr uint32_t TBL_ROW_BYTES = TBL_STRIDE * (uint32_t)sizeof(T);
      auto tbl_bytes_of = [&](int envs) -> uint32_t {
        return (!UNIFORM && a.tbl_off != 0) ? (uint32_t)envs * TBL_ROW_BYTES : 0u;
      };
      auto table_load = [&](int s, int tile, uint32_t bytes) {
        if (bytes != 0)
          bulk_load(in0 + (uint32_t)s * PIPE_TILE_BYTES + (uint32_t)a.tbl_off,
                    reinterpret_cast<const unsigned char*>(a.table) +
                        (size_t)tile * a.epc * TBL_ROW_BYTES,
                    bytes, full0 + 8 * s);
      };
      // Rewards of roads wider than a warp: the consumer warps leave their
      // partial sums next to the finished tile (r_num / r_den[o]); this thread
      // adds them up in warp order and stores the quotient, so the consumers
      // need no second barrier per tile.
      auto finish_rewards = [&](int o, int tile) {
        if (a.reward == nullptr || a.tpe <= 32 || reducer_on) return;
        const int envs = tile == last_tile ? last_envs : a.epc;
        const int nw = a.tpe >> 5;
        for (int el = 0; el < envs; ++el) {
          T n = sm.r_num[o][el * nw], d = sm.r_den[o][el * nw];
          for (int j = 1; j < nw; ++j) { n += sm.r_num[o][el * nw + j]; d += sm.r_den[o][el * nw + j]; }
          a.reward[(size_t)tile * a.epc + el] = n / d;
        }
      };
      if constexpr (!DYN) {
        auto issue_load = [&](int i) {
          const int tile = (int)blockIdx.x + i * (int)gridDim.x;
          const int s = i % n_stages;
          const int envs = tile == last_tile ? last_envs : a.epc;
          const uint32_t tbl = tbl_bytes_of(envs);
          if (ARZ) {
            const uint32_t bytes = bytes_of(tile);
            mbar_expect_tx(full0 + 8 * s, bytes + tbl);
            tile_load(in0 + (uint32_t)s * PIPE_TILE_BYTES, gin + (size_t)tile * tile_bytes, bytes,
                      full0 + 8 * s);
          } else {
            // LWR ignores the input velocity (lwr.py:197): fetch the density half
            // of each road only, at its usual place in the stage
            const uint32_t half = row_bytes >> 1;
            mbar_expect_tx(full0 + 8 * s, (uint32_t)envs * half + tbl);
            for (int r = 0; r < envs; ++r)
              tile_load(in0 + (uint32_t)s * PIPE_TILE_BYTES + (uint32_t)r * row_bytes,
                        gin + (size_t)tile * tile_bytes + (size_t)r * row_bytes, half,
                        full0 + 8 * s);
          }
          table_load(s, tile, tbl);
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
        const int pre = n_my < n_stages ? n_my : n_stages;
        for (int i = 0; i < pre; ++i) issue_load(i);
        const int ahead = a.l2_ahead;
        for (int i = n_stages; i < n_stages + ahead && i < n_my; ++i) prefetch(i);
        if (early) pdl_wait_prior_grid();   // before the first store
        for (int i = 0; i < n_my; ++i) {
          const int tile = (int)blockIdx.x + i * (int)gridDim.x;
          const int s = i % n_stages, o = i % PIPE_OUT;
          if (ahead > 0 && i + n_stages + ahead < n_my) prefetch(i + n_stages + ahead);
          if (i + n_stages < n_my) {  // refill stage s once the consumers have read tile i
            mbar_wait_relaxed(empty0 + 8 * s, (uint32_t)((i / n_stages) & 1), PRODUCER_NAP_NS);
            issue_load(i + n_stages);
          }
          mbar_wait_relaxed(ofull0 + 8 * o, (uint32_t)((i / PIPE_OUT) & 1), PRODUCER_NAP_NS);
          bulk_store(gout + (size_t)tile * tile_bytes, out0 + (uint32_t)o * PIPE_TILE_BYTES,
                     bytes_of(tile));
          bulk_commit();
          finish_rewards(o, tile);
          // all but the newest PIPE_OUT-1 stores have finished reading shared memory
          bulk_wait_read<PIPE_OUT - 1>();
          if (i >= PIPE_OUT - 1) mbar_arrive(ofree0 + 8 * ((i - (PIPE_OUT - 1)) % PIPE_OUT));
        }
      } else {
        // Tile schedule.  The first ring-full of every CTA is static (tile =
        // block + g * grid) so nothing waits on an atomic at start-up; after
        // that the CTAs draw tiles from one counter, which evens out their
        // finishing times (SMs do not progress at the same rate, and the next
        // step cannot start before the slowest CTA is done).
        int g = 0;  // tiles drawn so far
        auto grab = [&]() -> int {
          int tl;
          if (a.sched == nullptr || g < PIPE_STAGES)
            tl = (int)blockIdx.x + g * (int)gridDim.x;
          else
            tl = (int)atomicAdd(a.sched, 1u) + PIPE_STAGES * (int)gridDim.x;
          ++g;
          return tl < a.n_tiles ? tl : -1;
        };
        // fill stage s with `tile`, or mark it as the end of this CTA's work
        auto issue_load = [&](int s, int tile) {
          sm.tile_of[s] = tile;
          if (tile < 0) {
            mbar_arrive(full0 + 8 * s);
            return;
          }
          const int envs = tile == last_tile ? last_envs : a.epc;
          const uint32_t tbl = tbl_bytes_of(envs);
          if (ARZ) {
            const uint32_t bytes = bytes_of(tile);
            mbar_expect_tx(full0 + 8 * s, bytes + tbl);
            tile_load(in0 + (uint32_t)s * PIPE_TILE_BYTES, gin + (size_t)tile * tile_bytes, bytes,
                      full0 + 8 * s);
          } else {
            // LWR ignores the input velocity (lwr.py:197): fetch the density half
            // of each road only, at its usual place in the stage
            const uint32_t half = row_bytes >> 1;
            mbar_expect_tx(full0 + 8 * s, (uint32_t)envs * half + tbl);
            for (int r = 0; r < envs; ++r)
              tile_load(in0 + (uint32_t)s * PIPE_TILE_BYTES + (uint32_t)r * row_bytes,
                        gin + (size_t)tile * tile_bytes + (size_t)r * row_bytes, half,
                        full0 + 8 * s);
          }
          table_load(s, tile, tbl);
        };
        auto prefetch = [&](int tile) {
          if (tile < 0) return;
          if (ARZ) {
            bulk_prefetch_l2(gin + (size_t)tile * tile_bytes, bytes_of(tile));
          } else {
            const int envs = tile == last_tile ? last_envs : a.epc;
            for (int r = 0; r < envs; ++r)
              bulk_prefetch_l2(gin + (size_t)tile * tile_bytes + (size_t)r * row_bytes,
                               row_bytes >> 1);
          }
        };
        MT_STAMP(1);
        for (int s = 0; s < PIPE_STAGES; ++s) issue_load(s, grab());
        if (early) pdl_wait_prior_grid();   // before the first store and the first counter draw
        // tiles drawn and prefetched into L2 but not yet loaded: a ring in
        // shared memory (this thread only), oldest at index i % PIPE_QMAX
        const int ahead = a.l2_ahead < PIPE_QMAX ? a.l2_ahead : PIPE_QMAX;
        for (int k = 0; k < ahead; ++k) {
          const int drawn = grab();
          prefetch(drawn);
          sm.drawn[k] = drawn;
        }
        for (int i = 0;; ++i) {
          const int s = i % PIPE_STAGES, o = i % PIPE_OUT;
          const int tile = sm.tile_of[s];
          if (tile < 0) break;
          const int next = ahead > 0 ? sm.drawn[i & (PIPE_QMAX - 1)] : grab();
          // refill stage s once the consumers have read tile i
          mbar_wait_relaxed(empty0 + 8 * s, (uint32_t)((i / PIPE_STAGES) & 1), PRODUCER_NAP_NS);
          issue_load(s, next);
          if (next < 0) MT_STAMP(3);
          mbar_wait_relaxed(ofull0 + 8 * o, (uint32_t)((i / PIPE_OUT) & 1), PRODUCER_NAP_NS);
          bulk_store(gout + (size_t)tile * tile_bytes, out0 + (uint32_t)o * PIPE_TILE_BYTES,
                     bytes_of(tile));
          bulk_commit();
          finish_rewards(o, tile);
          // all but the newest PIPE_OUT-1 stores have finished reading shared memory
          bulk_wait_read<PIPE_OUT - 1>();
          if (i >= PIPE_OUT - 1) mbar_arrive(ofree0 + 8 * ((i - (PIPE_OUT - 1)) % PIPE_OUT));
          if (ahead > 0) {
            // draw one more tile last: the atomic's round trip then overlaps the
            // wait for the consumers instead of delaying the load and the store
            const int drawn = grab();
            prefetch(drawn);
            sm.drawn[(i + ahead) & (PIPE_QMAX - 1)] = drawn;
          }
        }
        if (a.sched != nullptr) {
          // the last CTA to finish drawing re-arms the schedule for the next launch
          __threadfence();
          if (atomicAdd(a.sched + 1, 1u) == gridDim.x - 1) {
            a.sched[0] = 0u;
            a.sched[1] = 0u;
            __threadfence();
          }
        }
      }
      bulk_wait_read<0>();  // shared memory must outlive the last bulk store
      MT_STAMP(6);
    }
    if constexpr (REDUCER) {
      if (t >= NC + 32) {
        // ===================== reducer warp ================================
        // Per tile and road: each lane adds the pairs of the road's threads
        // lane, lane + 32, ... in that order, a butterfly adds the lanes, lane
        // 0 stores sum(rho v) / sum(rho).  A fixed order; the consumers are
        // spared the butterfly (10 shuffles and their selects per thread).
        const int lane = t & 31;
        const unsigned full = 0xffffffffu;
        if (early) pdl_wait_prior_grid();   // before the first reward is stored
        for (int i = 0; i < n_my; ++i) {
          const int tile = (int)blockIdx.x + i * (int)gridDim.x;
          const int o = i % PIPE_OUT;
          const int envs = tile == last_tile ? last_envs : a.epc;
          mbar_wait(ofull0 + 8 * o, (uint32_t)((i / PIPE_OUT) & 1));
          for (int el = 0; el < envs; ++el) {
            const T* pr = pairbuf + 2 * (o * NC + el * a.tpe);
            T n = 0, d = 0;
            for (int k = lane; k < a.tpe; k += 32) {
              n += pr[2 * k];
              d += pr[2 * k + 1];
            }
            const bool odd = (lane & 1) != 0;
            T acc = (odd ? d : n) + __shfl_xor_sync(full, odd ? n : d, 1);
            for (int off = 2; off <= 16; off <<= 1) acc += __shfl_xor_sync(full, acc, off);
            const T dsum = __shfl_sync(full, acc, 1);
            if (lane == 0) a.reward[(size_t)tile * a.epc + el] = acc / dsum;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(ofree0 + 8 * o);   // the pairs of stage o are summed
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
  const uint32_t row0_off = (uint32_t)el * row_bytes;  // my road's first byte

  EnvC<T> e;
  if (UNIFORM) {
    load_uniform(e, a.u, 0, (const T*)nullptr, a.step);
    if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
  }

  int s = 0, o = 0;
  uint32_t ph_in = 0, ph_out = 0;
  for (int i = 0; DYN || i < n_my; ++i) {
    int tile;
    if constexpr (DYN) {
      // the producer names the tile of each stage; -1 ends this CTA's work
      mbar_wait(full0 + 8 * s, ph_in);
      tile = sm.tile_of[s];
      if (tile < 0) break;
    } else {
      tile = (int)blockIdx.x + i * (int)gridDim.x;
    }
    const int envs_here = tile == last_tile ? last_envs : a.epc;
    const int env = tile * a.epc + el;
    const bool active = owns_cells && (el < envs_here);

    const bool staged_table = !UNIFORM && a.tbl_off != 0;
    // Static schedule: the next tile's speed limit of my road is pulled into L1
    // now (the tiles are ready before the consumers are, so the load below
    // would otherwise be waited for in full: +10 % per step); a prefetch keeps
    // no register alive across the tile.
    const T* act_ptr = a.action;
    if constexpr (!DYN) {
      const int tile_n = tile + (int)gridDim.x;
      if (a.action != nullptr && owns_cells && tile_n < a.n_tiles) {
        const long long env_n = (long long)tile_n * a.epc + el;
        if (env_n < a.B) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.action + env_n));
      }
    }
    if (active && !staged_table && (!UNIFORM || a.action != nullptr)) {
      // per-road constants (and/or this step's speed limit); with the static
      // schedule they are issued before the wait so that the loads overlap it
      if (UNIFORM) load_uniform(e, a.u, env, act_ptr, a.step);
      else load_env(e, a.table, a.flags, env, act_ptr, a.step);
      if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
    }

    if constexpr (!DYN) mbar_wait(full0 + 8 * s, ph_in);
    if constexpr (!UNIFORM) {
      if (active && staged_table) {
        // my road's row of the parameter table arrived with the tile
        const T* trow = reinterpret_cast<const T*>(sm.in[s] + a.tbl_off) + el * TBL_STRIDE;
        load_env_staged(e, trow, env, act_ptr, a.step);
        if (ARZ) e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
      }
    }

    if (t == 0) { if (i == 0) MT_STAMP(2); MT_STAMP(4); }
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
    T num = 0, den = 0;
    T rn[CPT], vn[CPT];
    // finish my interior cells -- they need nothing from other threads --
    // before the barrier: less state lives across it, less work follows it
    T Q[CPT], P[CPT];  // my fluxes (LWR: Q only)
    if (active) {
      with_lam<LAMK>(e.flags, [&](auto kind) {
        cell_mid<T, ARZ, FAST, decltype(kind)::value, CPT>(e, rho, y, r, D, S, Q, P, rn, vn);
      });
    }
    consumer_sync(NC);                               // exchange visible; stage s fully read
    if (lane0) mbar_arrive(empty0 + 8 * s);          // release the input stage
    mbar_wait(ofree0 + 8 * o, ph_out ^ 1);           // output stage drained by its last store
    if (active) {
      const T S_right = xS[xi + 1], D_left = xD[xi], r_left = ARZ ? xR[xi] : T(0);
      with_lam<LAMK>(e.flags, [&](auto kind) {
        cell_edges<T, ARZ, FAST, decltype(kind)::value, CPT>(
            e, a.bc, row_first, row_last, rho, v, y, r, D, S, Q, P, S_right, D_left, r_left,
            loop_a, loop_b, rn, vn);
      });
    }
    if (active) {
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
        sts16(dst + v_off + 16 * u, tv);
      }
      if (a.obs16 != nullptr) {
        // (8-byte stores per four cells of a field, coalesced over the warp)
        uint16_t* o16 = a.obs16 + (size_t)env * (2 * (size_t)N) + (size_t)pos * CPT;
        if constexpr (CPT % 4 == 0) {
#pragma unroll
          for (int q = 0; q < CPT / 4; ++q) {
            *reinterpret_cast<uint2*>(o16 + 4 * q) =
                pack_bf16x4(rn[4 * q], rn[4 * q + 1], rn[4 * q + 2], rn[4 * q + 3]);
            *reinterpret_cast<uint2*>(o16 + N + 4 * q) =
                pack_bf16x4(vn[4 * q], vn[4 * q + 1], vn[4 * q + 2], vn[4 * q + 3]);
          }
        } else {   // two cells per thread (f64, one vector)
          *reinterpret_cast<uint32_t*>(o16) = pack_bf16x4(rn[0], rn[1], rn[0], rn[1]).x;
          *reinterpret_cast<uint32_t*>(o16 + N) = pack_bf16x4(vn[0], vn[1], vn[0], vn[1]).x;
        }
      }
    }
    if constexpr (REDUCER) {
      // my sums go to the reducer warp (all consumer threads store: idle ones zeros)
      T* pr = pairbuf + 2 * (o * NC + t);
      pr[0] = num;
      pr[1] = den;
    } else if (want_reward && a.tpe > 32) {
      // Per-road sum(rho v) / sum(rho), fixed order.  Shuffles are the cost of
      // this reduction (one warp-wide SHFL per clock and SM), so the two sums
      // share one butterfly: in the first round even lanes keep the numerator
      // and odd lanes the denominator of the pair, the other four rounds then
      // move ONE value per lane.  Lane 0 ends up with the warp's numerator,
      // lane 1 with its denominator; they go next to the finished tile and
      // the producer thread adds the warps up (finish_rewards).
      const unsigned full = 0xffffffffu;
      const bool odd = (t & 1) != 0;
      T acc = (odd ? den : num) + __shfl_xor_sync(full, odd ? num : den, 1);
      for (int off = 2; off <= 16; off <<= 1) acc += __shfl_xor_sync(full, acc, off);
      if ((t & 31) == 0) sm.r_num[o][t >> 5] = acc;
      if ((t & 31) == 1) sm.r_den[o][t >> 5] = acc;
    }
    fence_proxy_async_smem();  // my generic-proxy writes -> visible to the bulk engine
    __syncwarp();
    if (lane0) mbar_arrive(ofull0 + 8 * o);          // non-blocking hand-off to the producer
    if (t == 0) MT_STAMP(5);

    if (want_reward && a.tpe <= 32) {
      // a road fits one warp: its lanes reduce among themselves
      const unsigned full = 0xffffffffu;
      for (int off = a.tpe >> 1; off > 0; off >>= 1) {
        num += __shfl_xor_sync(full, num, off);
        den += __shfl_xor_sync(full, den, off);
      }
      if (active && row_first) a.reward[env] = num / den;
    }
    if (++s == n_stages) { s = 0; ph_in ^= 1; }
    if (++o == PIPE_OUT) { o = 0; ph_out ^= 1; }
  }
}

}  // namespace mt
