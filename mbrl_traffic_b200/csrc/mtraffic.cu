// mtraffic.cu -- kernels and C ABI of libmtraffic (sm_100a).
//
// One fused kernel per model step (mt_pipe.cuh: persistent, TMA-pipelined;
// mt_direct.cuh: one CTA per road tile).  Each thread owns a few consecutive
// cells of one road, derives the per-cell flux ingredients, swaps three edge
// values with its neighbours through shared memory, applies the finite-volume
// update, the boundary rule and the (rho, y) -> (rho, v) conversion; rewards
// are reduced per road with warp shuffles in the same kernel.
//
// Reference semantics are cited function by function in mt_math.cuh and
// mt_cell.cuh; the C ABI is documented in include/mtraffic.h.  This file holds
// the parameter-table kernel, the launch logic and the ABI itself.
#include "mtraffic.h"
#include "mt_math.cuh"
#include "mt_pipe.cuh"
#include "mt_ring.cuh"
#include "mt_fused.cuh"
#include "mt_seg.cuh"
#include "mt_halo.cuh"
#include "mt_segfused.cuh"
#include "mt_direct.cuh"
#include "mt_nonlocal.cuh"
#include "mt_aggregate.cuh"
#include "mt_abort.cuh"

// defined in mt_pipe_arz64.cu
namespace mt {
#define MT_PIPE_ARZ64(K, UNI, U, R) \
  extern template __global__ void step_pipe_kernel<double, 1, false, K, UNI, U, R>(const PipeArgs<double>);
#include "mt_pipe_arz64.inc"
#undef MT_PIPE_ARZ64
}  // namespace mt

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

namespace {

using namespace mt;
#ifdef MT_TRACE
unsigned long long* g_trace = nullptr;
unsigned g_trace_seq = 0;
#endif

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess)                                                    \
      return fail(MT_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));  \
  } while (0)

// ---------------------------------------------------------------------------
// parameter table  (one thread per env)
// ---------------------------------------------------------------------------
template <typename T> struct ParamArgs {
  T* table;
  int* flags;
  long long B;
  int model;
  int pow_fast;  // MT_FLAG_FAST_MATH
  double rho_max, v_max, tau, lam, rho_crit, dt;
  const T *rho_max_env, *v_max_env, *tau_env, *lam_env, *rho_crit_env;
  unsigned* summary;   // [0] |= 1 << exponent kind of every road, [1] |= (rho_max != 1)
};

template <typename T>
__global__ void set_params_kernel(const ParamArgs<T> p) {
  const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= p.B) return;
  using A = Ar<T, false>;
  const T rho_max = p.rho_max_env ? p.rho_max_env[env] : (T)p.rho_max;
  const T v_max = p.v_max_env ? p.v_max_env[env] : (T)p.v_max;
  const T lam = p.lam_env ? p.lam_env[env] : (T)p.lam;
  T rho_crit;
  if (p.model == MT_MODEL_ARZ && p.rho_crit_env) rho_crit = p.rho_crit_env[env];
  else if (p.model == MT_MODEL_ARZ && p.rho_crit >= 0) rho_crit = (T)p.rho_crit;
  else rho_crit = A::div(rho_max, T(2));          // lwr.py:282, arz.py:199
  // 1 + dt/tau  (arz.py:461, 492): double arithmetic for a scalar tau, as in
  // the reference where both are Python floats.
  T c;
  if (p.tau_env) c = A::add(T(1), A::div((T)p.dt, p.tau_env[env]));
  else c = (T)(1.0 + p.dt / p.tau);
  int kind = LAM_GENERAL;
  if (lam == T(1)) kind = LAM_ONE;
  else if (lam == T(2)) kind = LAM_TWO;
  else if (lam == T(0.5)) kind = LAM_HALF;
  if (kind == LAM_GENERAL && p.pow_fast) kind = LAM_GENERAL_FAST;
  int fl = kind;
  // (1 - rho_crit/rho_max) ** lam  for q_max (arz.py:396)
  const T base = A::sub(T(1), A::div(rho_crit, rho_max));
  T g;
  if (kind == LAM_ONE) g = base;
  else if (kind == LAM_TWO) g = A::mul(base, base);
  else if (kind == LAM_HALF) g = mt_sqrt(base);
  else g = p.pow_fast ? mt_pow_fast(base, lam) : mt_pow(base, lam);
  T* row = p.table + env * TBL_STRIDE;
  row[TBL_RHO_MAX] = rho_max;
  row[TBL_INV_RHO_MAX] = A::div(T(1), rho_max);
  row[TBL_V_MAX] = v_max;
  row[TBL_LAM] = lam;
  row[TBL_RHO_CRIT] = rho_crit;
  row[TBL_G_CRIT] = g;
  row[TBL_C] = c;
  row[TBL_INV_C] = A::div(T(1), c);
  row[TBL_FLAGS] = (T)fl;
  for (int k = TBL_FLAGS + 1; k < TBL_STRIDE; ++k) row[k] = T(0);
  p.flags[env] = fl;
  // batch-wide summary (one atomic per warp and word)
  const unsigned kinds = __reduce_or_sync(__activemask(), 1u << fl);
  const unsigned not_one = __reduce_or_sync(__activemask(), rho_max != T(1) ? 1u : 0u);
  if ((threadIdx.x & 31) == 0) {
    atomicOr(p.summary, kinds);
    if (not_one) atomicOr(p.summary + 1, 1u);
  }
}

// ---------------------------------------------------------------------------
// compute_loss support: per-env squared density error (lwr.py:338-343)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void density_sqerr_kernel(const T* __restrict__ pred, const T* __restrict__ target,
                                     T* __restrict__ out, long long B, int N) {
  const long long env = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= B) return;
  const int lane = threadIdx.x & 31;
  const T* p = pred + env * (2LL * N);
  const T* q = target + env * (2LL * N);
  T s = 0;
  for (int i = lane; i < N; i += 32) { const T d = p[i] - q[i]; s += d * d; }
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[env] = s;
}

// ---------------------------------------------------------------------------
// failure detection: per-road "state is not finite" flag (a diverged road --
// an erratic controller can drive the reference scheme itself unstable --
// shows up as inf/NaN cells; callers reset or drop such roads)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void nonfinite_kernel(const T* __restrict__ obs, int* __restrict__ out, long long B,
                                 int N) {
  const long long env = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= B) return;
  const int lane = threadIdx.x & 31;
  const T* p = obs + env * (2LL * N);
  int bad = 0;
  for (int i = lane; i < 2 * N; i += 32) {
    const T x = p[i];
    bad |= !(x - x == T(0));   // inf - inf and NaN - NaN are NaN
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) out[env] = bad ? 1 : 0;
}

// bfloat16 copy of a batch (paths of mt_step_obs16 that do not go through the
// pipelined kernel, which writes the copy itself): 4 values per thread
template <typename T>
__global__ void obs_to_bf16_kernel(const T* __restrict__ obs, uint16_t* __restrict__ out,
                                   long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    *reinterpret_cast<uint2*>(out + i) = pack_bf16x4(obs[i], obs[i + 1], obs[i + 2], obs[i + 3]);
  } else {
    for (long long k = i; k < n; ++k) {
      const __nv_bfloat16 h = __float2bfloat16_rn((float)obs[k]);
      out[k] = *reinterpret_cast<const uint16_t*>(&h);
    }
  }
}

// Last layer of the policy MLP fused with its squashing: one warp per road,
//   action = scale * (tanh(h . w + b) + 1),  h: (B, H) bfloat16,
// written in the engine dtype straight into the buffer the next mt_step reads
// its speed limits from (FeedForwardActor, policies/sac.py:303-322 shape).
template <typename T>
__global__ void actor_head_kernel(const __nv_bfloat16* __restrict__ h,
                                  const __nv_bfloat16* __restrict__ w,
                                  const __nv_bfloat16* __restrict__ bias, T* __restrict__ action,
                                  long long B, int H, float scale,
                                  const T* __restrict__ reward_prev, T* __restrict__ return_sum) {
  // the step kernel enqueued next (a programmatic dependent launch) may set up
  // its pipeline while this kernel runs; it reads the actions only after its
  // own grid-dependency wait, i.e. after this kernel has completed
  pdl_launch_dependents();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  const __nv_bfloat16* hr = h + row * H;
  float acc = 0.f;
  for (int k = lane; k < H; k += 32) acc = fmaf(__bfloat162float(hr[k]), __bfloat162float(w[k]), acc);
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    // the reference layer rounds its output to the activation dtype before tanh
    const float z = __bfloat162float(__float2bfloat16_rn(acc + __bfloat162float(bias[0])));
    action[row] = (T)(scale * (tanhf(z) + 1.0f));
    // the rollout's return accumulates here: the previous step's reward (already
    // final: that step kernel completed before this kernel started) joins the sum
    if (return_sum != nullptr) return_sum[row] += reward_prev[row];
  }
}

}  // namespace

// ---------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------
struct mt_engine {
  mt_config cfg;
  size_t elt;             // sizeof dtype
  void* table = nullptr;  // [B, TBL_STRIDE]
  int* flags = nullptr;   // [B]
  void* partial = nullptr;
  // tile-schedule counters of the pipelined kernel: one {next, done} pair per
  // stream the engine launches on (slot 0: the caller's stream; 1..: hs[])
  unsigned int* sched = nullptr;
  int max_tiles = 1;
  int sm_count = 148;
  bool have_params = false;
  mt_params params;       // last set
  // all-scalar parameter sets with lam in {1, 2, 0.5}: constants for the
  // UNIFORM kernel instantiations (host arithmetic == device arithmetic)
  bool uniform = false;
  int lamk = LAM_RUNTIME;
  // per-road tables: the exponent kind ALL roads share (LAM_ONE_RM1 when they
  // also all have rho_max == 1), or LAM_RUNTIME; read back by mt_set_params
  int table_lamk = LAM_RUNTIME;
  unsigned* summary = nullptr;   // device, 2 words
  UniformC<double> u64;
  UniformC<float> u32;
  int64_t launches = 0;
  // mt_step_obs16: where the step being issued leaves a bfloat16 copy of the
  // new state (null for every other call); consumed by step_typed
  uint16_t* obs16 = nullptr;
  bool obs16_written = false;   // the step kernel wrote it (else step_range converts)
  // table readiness: recorded by mt_set_params, awaited once by the next
  // step issued on a different stream
  cudaEvent_t params_ready = nullptr;
  cudaStream_t params_stream = nullptr;
  bool params_pending = false;
  // host staging: chunks of roads are pipelined over these streams so that the
  // upload of one chunk overlaps the compute and the download of others
  static constexpr int HOST_STREAMS = 3;
  cudaStream_t hs[HOST_STREAMS] = {nullptr, nullptr, nullptr};
  cudaEvent_t action_ready = nullptr;
  // per-chunk hand-offs: upload stream -> compute stream -> download stream
  static constexpr int MAX_CHUNKS = 80;
  cudaEvent_t ev_up[MAX_CHUNKS] = {}, ev_done[MAX_CHUNKS] = {};
  cudaStream_t stream = nullptr;
  void *d_a = nullptr, *d_b = nullptr, *d_action = nullptr, *d_reward = nullptr;
};

struct mt_halo {
  int device = 0, dtype = MT_F64, halo = 0;
  size_t bytes = 0;
  void* box = nullptr;                     // my mailbox (device memory of this rank)
  void *left = nullptr, *right = nullptr;  // the neighbours' mailboxes as mapped here
  bool left_ipc = false, right_ipc = false;
  unsigned pushes = 0, pulls = 0;          // exchange numbers issued so far
};

namespace {

// ---------------------------------------------------------------------------
// What this library enqueued LAST on a stream (process-wide: engines share
// streams).  A pipelined step may start loading before the launch in front of
// it has drained only if that launch is known and provably does not write what
// this one reads; every other entry point that enqueues work forgets the
// record, so "unknown" always means "wait".  (Work of other libraries between
// two launches needs no record: an ordinary kernel or copy starts after the
// launch before it has completed and is itself complete before a programmatic
// dependent launch behind it starts.)
// ---------------------------------------------------------------------------
struct StreamNote {
  int device = -1;
  cudaStream_t stream = nullptr;
  bool full_pipe = false;          // a single-step pipelined launch that filled every CTA slot
  const char* out_lo = nullptr;    // bytes it writes: [out_lo, out_hi)
  const char* out_hi = nullptr;
};
constexpr int MAX_NOTES = 32;
std::mutex g_note_mu;
StreamNote g_notes[MAX_NOTES];
int g_note_next = 0;

// returns the record of (device, stream) and forgets it
StreamNote note_take(int device, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_note_mu);
  for (int i = 0; i < MAX_NOTES; ++i) {
    StreamNote& n = g_notes[i];
    if (n.device == device && n.stream == stream) {
      const StreamNote out = n;
      n = StreamNote();
      return out;
    }
  }
  return StreamNote();
}
void note_put(const StreamNote& note) {
  std::lock_guard<std::mutex> lock(g_note_mu);
  for (int i = 0; i < MAX_NOTES; ++i)
    if (g_notes[i].device < 0) { g_notes[i] = note; return; }
  g_notes[g_note_next] = note;     // table full: overwrite round-robin (forgetting is safe)
  g_note_next = (g_note_next + 1) % MAX_NOTES;
}
inline void note_forget(int device, void* stream) { (void)note_take(device, (cudaStream_t)stream); }

struct Geometry {
  bool vector_ok;
  int nvec, tpe, epc, tiles, block;
  long long grid;
};

int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

Geometry plan(long long B, int N, int V, bool aligned) {
  Geometry g;
  g.vector_ok = aligned && (N % V == 0) && (N >= 2 * V);
  g.nvec = N / V;
  if (g.nvec <= MAX_BLOCK) {
    g.tiles = 1;
    g.tpe = g.nvec <= 32 ? next_pow2(g.nvec) : ((g.nvec + 31) / 32) * 32;
    g.block = g.tpe <= 256 ? 256 : g.tpe;
    g.epc = g.block / g.tpe;
    g.grid = (B + g.epc - 1) / g.epc;
  } else {
    g.block = 256;
    g.tpe = g.block;
    g.epc = 1;
    g.tiles = (g.nvec + g.block - 1) / g.block;
    g.grid = B * (long long)g.tiles;
  }
  return g;
}

// programmatic dependent launch of back-to-back step kernels (MT_PDL=0 disables)
bool pipe_use_pdl() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("MT_PDL");
    v = (s && s[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// the automatic independent-input check (MT_AUTO_INDEPENDENT=0 disables: A/B knob)
bool auto_independent() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("MT_AUTO_INDEPENDENT");
    v = (s && s[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// Which persistent kernel: the producer-warp one (mt_pipe.cuh) measured faster
// in f64 and f32 FAST, the producer-less ring (mt_ring.cuh) in f32 EXACT --
// except LWR batches large enough for the dynamic tile schedule, which only
// the producer-warp kernel has (B200, 16384x1000 f32 EXACT: ARZ 82 % ring vs
// 74 % pipe, LWR 79 % ring vs 91 % pipe of HBM peak).
// MT_KERNEL=pipe|ring overrides for A/B runs.
// (ARZ f64 with a per-road parameter table needs ~16 more live registers than
// the 96 the producer-warp kernel is held to and spills 80-100 B; the ring kernel
// has 128 and does not spill, yet measures slower in a build without split
// compilation -- 4096 x 1000, lam = 1: pipe with staged table rows 38.7 us, pipe
// reading the table from global memory 41.0 us, ring 43.4 us -- so such batches
// stay with the pipelined kernel; `uniform` is kept for A/B builds.)
bool pipe_use_ring(bool is_f32, bool fast, bool arz, bool dynamic, bool uniform) {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("MT_KERNEL");
    v = !s ? 2 : (s[0] == 'r') ? 1 : 0;
  }
  (void)uniform;
  if (v == 2) return is_f32 && !fast && (arz || !dynamic);
  return v != 0;
}

// MT_REWARD_REDUCER=0: rewards by the consumers' butterfly and the producer thread (A/B knob)
bool pipe_reward_reducer() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("MT_REWARD_REDUCER");
    v = (s && s[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// MT_TABLE_STAGE=0: the pipelined kernel's consumers read their table rows from
// global memory instead of a copy staged with the tile (A/B knob)
bool pipe_stage_table() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("MT_TABLE_STAGE");
    v = (s && s[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <typename K, typename Args>
int launch_pdl(K kernel, const Args& args, int grid, int block, size_t smem, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pipe_use_pdl() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, args));
  return MT_OK;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per function AND per
// device: every instantiation keeps one "configured" flag per device.
constexpr int MAX_DEVICES = 64;
inline bool* configured_flag(bool (&flags)[MAX_DEVICES]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return nullptr;
  return &flags[dev];
}

template <typename T, int MODEL, bool FAST, int LAMK, bool UNIFORM, int U>
int launch_pipe(const PipeArgs<T>& pa, int grid, cudaStream_t stream) {
  static bool configured_dev[MAX_DEVICES] = {};  // per instantiation and device
  bool* cfg_flag = configured_flag(configured_dev);
  bool unused_flag = false;
  bool& configured = cfg_flag ? *cfg_flag : unused_flag;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(step_pipe_kernel<T, MODEL, FAST, LAMK, UNIFORM, U>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(PipeSmem<T, U>)));
    CUDA_TRY(cudaFuncSetAttribute(step_fused_kernel<T, MODEL, FAST, LAMK, UNIFORM, U>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(PipeSmem<T, U>)));
    CUDA_TRY(cudaFuncSetAttribute(step_ring_kernel<T, MODEL, FAST, LAMK, UNIFORM, U>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(RingSmem<T, U>)));
    configured = true;
  }
  if (pa.n_inner == 1 && pipe_use_ring(sizeof(T) == 4, FAST, MODEL == 1, pa.sched != nullptr, UNIFORM))
    return launch_pdl(step_ring_kernel<T, MODEL, FAST, LAMK, UNIFORM, U>, pa, grid, pa.nc,
                      sizeof(RingSmem<T, U>), stream);
  if (pa.n_inner > 1) {
    if constexpr (sizeof(T) == 8 && U == 2 && !FAST) {
      if (pa.reducer) {   // a reward row per step: one more warp adds the sums up (mt_fused.cuh)
        static bool red_dev[MAX_DEVICES] = {};
        bool* rf = configured_flag(red_dev);
        if (rf == nullptr || !*rf) {
          CUDA_TRY(cudaFuncSetAttribute(step_fused_kernel<T, MODEL, FAST, LAMK, UNIFORM, U, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(PipeSmem<T, U>)));
          if (rf) *rf = true;
        }
        return launch_pdl(step_fused_kernel<T, MODEL, FAST, LAMK, UNIFORM, U, true>, pa, grid,
                          pa.nc + 64, sizeof(PipeSmem<T, U>), stream);
      }
    }
    return launch_pdl(step_fused_kernel<T, MODEL, FAST, LAMK, UNIFORM, U>, pa, grid,
                      pa.nc + 32, sizeof(PipeSmem<T, U>), stream);
  }
  if constexpr (MODEL == 1 && sizeof(T) == 8 && !FAST) {
    if (pa.reducer) {   // one more warp adds up the rewards (mt_pipe.cuh)
      static bool red_dev[MAX_DEVICES] = {};
      bool* rf = configured_flag(red_dev);
      if (rf == nullptr || !*rf) {
        CUDA_TRY(cudaFuncSetAttribute(step_pipe_kernel<T, MODEL, FAST, LAMK, UNIFORM, U, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(PipeSmem<T, U>)));
        if (rf) *rf = true;
      }
      return launch_pdl(step_pipe_kernel<T, MODEL, FAST, LAMK, UNIFORM, U, true>, pa, grid,
                        pa.nc + 64, sizeof(PipeSmem<T, U>), stream);
    }
  }
  return launch_pdl(step_pipe_kernel<T, MODEL, FAST, LAMK, UNIFORM, U>, pa, grid,
                    pa.nc + 32, sizeof(PipeSmem<T, U>), stream);
}

template <typename T, int MODEL, bool FAST, int U>
int launch_pipe_lam(const PipeArgs<T>& pa, int grid, bool uniform, int lamk, cudaStream_t s,
                    int table_lamk = LAM_RUNTIME) {
  if (uniform && lamk == LAM_ONE && pa.u.rho_max == T(1))
    return launch_pipe<T, MODEL, FAST, LAM_ONE_RM1, true, U>(pa, grid, s);
  if (uniform && lamk == LAM_ONE) return launch_pipe<T, MODEL, FAST, LAM_ONE, true, U>(pa, grid, s);
  if (uniform && lamk == LAM_TWO) return launch_pipe<T, MODEL, FAST, LAM_TWO, true, U>(pa, grid, s);
  if (uniform && lamk == LAM_HALF) return launch_pipe<T, MODEL, FAST, LAM_HALF, true, U>(pa, grid, s);
  // per-road tables whose roads all share an exponent kind (f64; the ARZ
  // instantiations live in mt_pipe_arz64.cu)
  if constexpr (sizeof(T) == 8 && !FAST) {
    if (!uniform && table_lamk == LAM_ONE_RM1)
      return launch_pipe<T, MODEL, FAST, LAM_ONE_RM1, false, U>(pa, grid, s);
    if (!uniform && table_lamk == LAM_ONE) return launch_pipe<T, MODEL, FAST, LAM_ONE, false, U>(pa, grid, s);
    if (!uniform && table_lamk == LAM_TWO) return launch_pipe<T, MODEL, FAST, LAM_TWO, false, U>(pa, grid, s);
    if (!uniform && table_lamk == LAM_HALF) return launch_pipe<T, MODEL, FAST, LAM_HALF, false, U>(pa, grid, s);
  }
  return launch_pipe<T, MODEL, FAST, LAM_RUNTIME, false, U>(pa, grid, s);
}

template <typename T> struct FastAllowed { static constexpr bool value = false; };
template <> struct FastAllowed<float> { static constexpr bool value = true; };

template <typename T, int MODEL, int U>
int launch_pipe_model(const PipeArgs<T>& pa, int grid, bool fast, bool uniform, int lamk,
                      cudaStream_t s, int table_lamk = LAM_RUNTIME) {
  if constexpr (FastAllowed<T>::value) {
    if (fast) return launch_pipe_lam<T, MODEL, true, U>(pa, grid, uniform, lamk, s);
  }
  return launch_pipe_lam<T, MODEL, false, U>(pa, grid, uniform, lamk, s, table_lamk);
}

template <typename T> const UniformC<T>& uniform_of(const mt_engine* e);
template <> const UniformC<double>& uniform_of<double>(const mt_engine* e) { return e->u64; }
template <> const UniformC<float>& uniform_of<float>(const mt_engine* e) { return e->u32; }

// Host restatement of set_params_kernel for an all-scalar parameter set with
// lam in {1, 2, 0.5}: +, -, *, / and sqrt are IEEE on both sides, so these
// constants equal the device table bit for bit.
template <typename T>
void make_uniform(UniformC<T>& u, int& lamk, int model, const mt_params& p) {
  const T rho_max = (T)p.rho_max, v_max = (T)p.v_max, lam = (T)p.lam;
  T rho_crit;
  if (model == MT_MODEL_ARZ && p.rho_crit >= 0) rho_crit = (T)p.rho_crit;
  else rho_crit = rho_max / T(2);
  const T c = (T)(1.0 + p.dt / p.tau);
  lamk = lam == T(1) ? LAM_ONE : lam == T(2) ? LAM_TWO : lam == T(0.5) ? LAM_HALF : LAM_GENERAL;
  const T base = T(1) - rho_crit / rho_max;
  T g = base;
  if (lamk == LAM_TWO) g = base * base;
  else if (lamk == LAM_HALF) g = std::sqrt(base);
  u.rho_max = rho_max;
  u.inv_rho_max = T(1) / rho_max;
  u.v_max = v_max;
  u.lam = lam;
  u.rho_crit = rho_crit;
  u.g_crit = g;
  u.c = c;
  u.inv_c = T(1) / c;
  u.flags = lamk;
}

// Tiles per CTA from which the pipelined kernel draws its tiles from a counter
// instead of round-robin.  A CTA draws PIPE_STAGES + l2_ahead tiles ahead of
// the one it computes, so with few tiles per CTA everything is drawn before
// the CTAs have drifted apart and only the counter's cost remains (measured:
// DESIGN.md 4.1).  MT_SCHED=static disables it, MT_SCHED=<n> sets the bound.
int pipe_dynamic_min_tiles() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("MT_SCHED");
    if (s && strcmp(s, "static") == 0) v = 0x7fffffff;
    else v = (s && atoi(s) > 0) ? atoi(s) : 16;
  }
  return v;
}

int pipe_l2_ahead() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("MT_PIPE_L2_AHEAD");
    v = s ? atoi(s) : 0;  // measured best on B200
    if (v < 0) v = 0;
    if (v > 64) v = 64;
  }
  return v;
}

// cells per thread of the pipelined kernel: U 16-byte vectors of each field
int pipe_vectors_per_thread() {
  static int u = -1;
  if (u < 0) {
    const char* s = getenv("MT_PIPE_U");
    u = (s && s[0] == '1') ? 1 : (s && s[0] == '2') ? 2 : 0;  // 0 = default
  }
  return u;
}


// ---- non-local model launchers ------------------------------------------------
template <typename T, bool FAST, int LAMK, bool UNIFORM, int U, bool LXF, bool AHEAD>
int launch_nl_scheme(const NlArgs<T>& na, int grid, cudaStream_t stream) {
  static bool configured_dev[MAX_DEVICES] = {};
  bool* cfg_flag = configured_flag(configured_dev);
  bool unused_flag = false;
  bool& configured = cfg_flag ? *cfg_flag : unused_flag;
  const size_t smem = sizeof(NlSmem<T>);
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(nonlocal_pipe_kernel<T, FAST, LAMK, UNIFORM, U, LXF, AHEAD>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  nonlocal_pipe_kernel<T, FAST, LAMK, UNIFORM, U, LXF, AHEAD><<<grid, na.nc + 32, smem, stream>>>(na);
  return MT_OK;
}

// MT_NL_AHEAD=0: the non-local kernel publishes a tile and sums it in turn (A/B knob)
bool nl_publish_ahead() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("MT_NL_AHEAD");
    v = (s && s[0] == '0') || NL_VE_BUFS < 3 ? 0 : 1;
  }
  return v != 0;
}

template <typename T, bool FAST, int LAMK, bool UNIFORM, int U>
int launch_nl(const NlArgs<T>& na, int grid, cudaStream_t stream, bool lxf = false) {
  if constexpr (UNIFORM) {
    // scalar parameters, no speed limits: tile i + 1 is published before tile i is summed
    if (na.action == nullptr && nl_publish_ahead()) {
      if (lxf) return launch_nl_scheme<T, FAST, LAMK, UNIFORM, U, true, true>(na, grid, stream);
      return launch_nl_scheme<T, FAST, LAMK, UNIFORM, U, false, true>(na, grid, stream);
    }
  }
  if (lxf) return launch_nl_scheme<T, FAST, LAMK, UNIFORM, U, true, false>(na, grid, stream);
  return launch_nl_scheme<T, FAST, LAMK, UNIFORM, U, false, false>(na, grid, stream);
}

template <typename T, bool FAST, int U>
int launch_nl_lam(const NlArgs<T>& na, int grid, bool uniform, int lamk, cudaStream_t s,
                  bool lxf) {
  if (uniform && lamk == LAM_ONE && na.u.rho_max == T(1))
    return launch_nl<T, FAST, LAM_ONE_RM1, true, U>(na, grid, s, lxf);
  if (uniform && lamk == LAM_ONE) return launch_nl<T, FAST, LAM_ONE, true, U>(na, grid, s, lxf);
  if (uniform && lamk == LAM_TWO) return launch_nl<T, FAST, LAM_TWO, true, U>(na, grid, s, lxf);
  if (uniform && lamk == LAM_HALF) return launch_nl<T, FAST, LAM_HALF, true, U>(na, grid, s, lxf);
  return launch_nl<T, FAST, LAM_RUNTIME, false, U>(na, grid, s, lxf);
}

template <typename T, int U>
int launch_nl_fast(const NlArgs<T>& na, int grid, bool fast, bool uniform, int lamk,
                   cudaStream_t s, bool lxf) {
  if constexpr (FastAllowed<T>::value) {
    if (fast) return launch_nl_lam<T, true, U>(na, grid, uniform, lamk, s, lxf);
  }
  return launch_nl_lam<T, false, U>(na, grid, uniform, lamk, s, lxf);
}

template <typename T>
int step_nonlocal(mt_engine* e, const T* obs_in, const T* action, T* obs_out, T* reward_out,
                  const T* table, const int* flags, long long B, bool aligned,
                  const BcArgs<T>& bc, T step, cudaStream_t stream) {
  constexpr int V = Vec<T>::N;
  const mt_params& p = e->params;
  const int N = (int)e->cfg.n_cells;
  const int n_eta = p.n_eta;
  const bool fast = FastAllowed<T>::value && (p.flags & MT_FLAG_FAST_MATH) != 0;
  const bool lxf = p.scheme == MT_SCHEME_LXF;
  const T half_alpha = p.alpha < 0 ? T(-1) : (T)(0.5 * p.alpha);
  if (n_eta > N) return fail(MT_EINVAL, "non-local: n_eta (%d) exceeds n_cells (%d)", n_eta, N);
  int U = pipe_vectors_per_thread();
  // two 16-byte pieces per thread in both types: fewer window reads and look-ahead
  // sums per cell (float32, 16384 x 1000, n_eta 16: 0.54 -> 0.64 of the copy roofline)
  if (U == 0) U = 2;
  if (N % (U * V) != 0 || N < 2 * U * V) U = 1;
  const int CPT = U * V;
  bool pipe_ok = aligned && (N % CPT == 0) && (N >= 2 * CPT) && !(p.flags & MT_FLAG_NO_PIPELINE) &&
                 n_eta <= NL_MAX_ETA && B <= 0x7fffffffLL;
  NlArgs<T> na;
  if (pipe_ok) {
    na.nthr = N / CPT;
    na.nc = PIPE_BLOCK / U;
    pipe_ok = na.nthr <= na.nc;
  }
  if (pipe_ok) {
    na.tpe = na.nthr <= 32 ? next_pow2(na.nthr) : ((na.nthr + 31) / 32) * 32;
    na.ve_cols = na.nthr + (n_eta + 1 + CPT - 1) / CPT + 1;
    na.epc = na.nc / na.tpe;
    if (na.epc > PIPE_MAX_EPC) na.epc = PIPE_MAX_EPC;
    const int cap = NlSmem<T>::VE / (CPT * na.ve_cols);   // roads whose V(rho) block fits
    if (na.epc > cap) na.epc = cap;
    pipe_ok = na.epc >= 1;
  }
  if (pipe_ok) {
    na.in = obs_in; na.out = obs_out; na.table = table; na.flags = flags;
    na.action = action; na.reward = reward_out;
    na.gamma = (const T*)p.gamma; na.n_eta = n_eta;
    na.B = (int)B; na.N = N;
    na.n_tiles = (int)((B + na.epc - 1) / na.epc);
    na.bc = bc; na.step = step; na.u = uniform_of<T>(e);
    na.half_alpha = half_alpha;
    const long long slots = (long long)NL_CTAS * e->sm_count;
    const int grid = (int)(na.n_tiles < slots ? na.n_tiles : slots);
    const int rc = U == 2 ? launch_nl_fast<T, 2>(na, grid, fast, e->uniform, e->lamk, stream, lxf)
                          : launch_nl_fast<T, 1>(na, grid, fast, e->uniform, e->lamk, stream, lxf);
    if (rc != MT_OK) return rc;
    e->launches++;
  } else {
    NlScalarArgs<T> sa;
    sa.in = obs_in; sa.out = obs_out; sa.table = table; sa.flags = flags; sa.action = action;
    sa.gamma = (const T*)p.gamma; sa.n_eta = n_eta; sa.B = B; sa.N = N; sa.bc = bc; sa.step = step;
    const long long cells = B * (long long)N;
    const int block = 256;
    const long long grid = (cells + block - 1) / block;
    if (grid > 0x7fffffffLL) return fail(MT_EINVAL, "batch too large for one launch");
    sa.half_alpha = half_alpha;
    if (lxf) {
      if (fast) nonlocal_lxf_scalar_kernel<T, FastAllowed<T>::value><<<(unsigned)grid, block, 0, stream>>>(sa);
      else nonlocal_lxf_scalar_kernel<T, false><<<(unsigned)grid, block, 0, stream>>>(sa);
    } else {
      if (fast) nonlocal_scalar_kernel<T, FastAllowed<T>::value><<<(unsigned)grid, block, 0, stream>>>(sa);
      else nonlocal_scalar_kernel<T, false><<<(unsigned)grid, block, 0, stream>>>(sa);
    }
    e->launches++;
    if (reward_out) {
      const int wpb = 8;
      reward_rows_kernel<T><<<(unsigned)((B + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
          (const T*)obs_out, reward_out, B, N);
      e->launches++;
    }
  }
  CUDA_TRY(cudaGetLastError());
  return MT_OK;
}

// ---- long roads: segment-pipelined kernel --------------------------------------
template <typename T, int MODEL, bool FAST, int LAMK, bool UNIFORM, int U>
int launch_seg(const SegArgs<T>& sa, int grid, cudaStream_t stream) {
  static bool configured_dev[MAX_DEVICES] = {};
  bool* cfg_flag = configured_flag(configured_dev);
  bool unused_flag = false;
  bool& configured = cfg_flag ? *cfg_flag : unused_flag;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(step_seg_kernel<T, MODEL, FAST, LAMK, UNIFORM, U>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(SegSmem<T, U>)));
    configured = true;
  }
  return launch_pdl(step_seg_kernel<T, MODEL, FAST, LAMK, UNIFORM, U>, sa, grid, sa.nc + 32,
                    sizeof(SegSmem<T, U>), stream);
}

// the reference defaults (scalar parameters, lam == 1, rho_max == 1) get the
// compile-time specialisation; everything else reads the per-road table
template <typename T, int MODEL, bool FAST, int U>
int launch_seg_spec(const SegArgs<T>& sa, int grid, bool uniform, int lamk, cudaStream_t s) {
  if (uniform && lamk == LAM_ONE && sa.u.rho_max == T(1))
    return launch_seg<T, MODEL, FAST, LAM_ONE_RM1, true, U>(sa, grid, s);
  return launch_seg<T, MODEL, FAST, LAM_RUNTIME, false, U>(sa, grid, s);
}

template <typename T, int MODEL, int U>
int launch_seg_fast(const SegArgs<T>& sa, int grid, bool fast, bool uniform, int lamk,
                    cudaStream_t s) {
  if constexpr (FastAllowed<T>::value) {
    if (fast) return launch_seg_spec<T, MODEL, true, U>(sa, grid, uniform, lamk, s);
  }
  return launch_seg_spec<T, MODEL, false, U>(sa, grid, uniform, lamk, s);
}

// ---- long roads, several steps per launch (mt_segfused.cuh) -----------------
template <typename T, int MODEL, bool FAST, int LAMK, bool UNIFORM, int U>
int launch_segfused(const SegFusedArgs<T>& sa, int grid, cudaStream_t stream) {
  static bool configured_dev[MAX_DEVICES] = {};
  bool* cfg_flag = configured_flag(configured_dev);
  bool unused_flag = false;
  bool& configured = cfg_flag ? *cfg_flag : unused_flag;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(step_segfused_kernel<T, MODEL, FAST, LAMK, UNIFORM, U>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(PipeSmem<T, U>)));
    configured = true;
  }
  return launch_pdl(step_segfused_kernel<T, MODEL, FAST, LAMK, UNIFORM, U>, sa, grid, sa.nc + 32,
                    sizeof(PipeSmem<T, U>), stream);
}

template <typename T, int MODEL, bool FAST, int U>
int launch_segfused_spec(const SegFusedArgs<T>& sa, int grid, bool uniform, int lamk,
                         cudaStream_t s) {
  if (uniform && lamk == LAM_ONE && sa.u.rho_max == T(1))
    return launch_segfused<T, MODEL, FAST, LAM_ONE_RM1, true, U>(sa, grid, s);
  return launch_segfused<T, MODEL, FAST, LAM_RUNTIME, false, U>(sa, grid, s);
}

template <typename T, int MODEL, int U>
int launch_segfused_fast(const SegFusedArgs<T>& sa, int grid, bool fast, bool uniform, int lamk,
                         cudaStream_t s) {
  if constexpr (FastAllowed<T>::value) {
    if (fast) return launch_segfused_spec<T, MODEL, true, U>(sa, grid, uniform, lamk, s);
  }
  return launch_segfused_spec<T, MODEL, false, U>(sa, grid, uniform, lamk, s);
}

// Largest number of steps one launch of the window kernel may fuse: the
// overlap 2 * roundup4(k) must leave at least half of a window.
constexpr int SEGFUSED_MAX_STEPS = 128;

// Window geometry of the several-steps-per-launch kernel for this engine, or
// false when it does not apply (short roads take step_fused_kernel, the LOOP
// rule needs the road's far end, the non-local model has its own kernel).
template <typename T>
bool segfused_geometry(const mt_engine* e, const void* obs_in, const void* obs_out, int& U,
                       int& nc, int& window) {
  constexpr int V = Vec<T>::N;
  const mt_params& p = e->params;
  const int N = (int)e->cfg.n_cells;
  if (e->cfg.model == MT_MODEL_NONLOCAL || (p.flags & MT_FLAG_NO_PIPELINE)) return false;
  if (e->cfg.model == MT_MODEL_ARZ && (p.flags & MT_FLAG_FSOLVE_ABORT)) return false;
  if (p.bc_left == MT_BC_LOOP || p.bc_right == MT_BC_LOOP) return false;
  const bool aligned = (((uintptr_t)obs_in | (uintptr_t)obs_out) & 15u) == 0;
  const Geometry g = plan(e->cfg.n_envs, N, V, aligned);
  if (!g.vector_ok || g.tiles <= 1) return false;
  if (N % 4 != 0) return false;                 // window starts must stay 16-byte aligned
  U = pipe_vectors_per_thread();
  if (U == 0) U = (sizeof(T) == 8) ? 2 : 1;
  nc = PIPE_BLOCK / U;
  window = nc * U * V;
  if (N <= window) return false;
  const char* off = getenv("MT_NO_FUSE");
  return !(off && off[0] == '1');
}

// n_inner steps of the whole batch in one launch of the window kernel.
template <typename T>
int segfused_typed(mt_engine* e, const void* obs_in_, const void* action_, long long action_stride,
                   void* obs_out_, int n_inner, const SegHalo* hx, cudaStream_t stream) {
  int U = 0, nc = 0, window = 0;
  if (!segfused_geometry<T>(e, obs_in_, obs_out_, U, nc, window))
    return fail(MT_EINVAL, "the several-steps-per-launch kernel does not apply to this engine");
  if (n_inner < 1 || n_inner > SEGFUSED_MAX_STEPS)
    return fail(MT_EINVAL, "steps per launch must be in [1, %d]", SEGFUSED_MAX_STEPS);
  const mt_params& p = e->params;
  const int N = (int)e->cfg.n_cells;
  const long long B = e->cfg.n_envs;
  SegFusedArgs<T> sa;
  sa.in = (const T*)obs_in_; sa.out = (T*)obs_out_;
  sa.table = (const T*)e->table; sa.flags = e->flags;
  sa.action = (const T*)action_; sa.action_stride = action_stride;
  sa.B = (int)B; sa.N = N; sa.nc = nc; sa.n_inner = n_inner;
  sa.hh = (n_inner + 3) & ~3;
  sa.tiles_per_row = segfused_tiles_per_row(N, window, sa.hh);
  const long long n_tiles = B * (long long)sa.tiles_per_row;
  if (n_tiles > 0x7fffffffLL) return fail(MT_EINVAL, "batch too large for one launch");
  sa.n_tiles = (int)n_tiles;
  sa.bc.bc_l = p.bc_left; sa.bc.bc_r = p.bc_right;
  sa.bc.bl0 = (T)p.bc_left_val[0]; sa.bc.bl1 = (T)p.bc_left_val[1];
  sa.bc.br0 = (T)p.bc_right_val[0]; sa.bc.br1 = (T)p.bc_right_val[1];
  sa.step = (T)(p.dt / p.dx);
  sa.u = uniform_of<T>(e);
  if (hx) {
    sa.halo = *hx;
    const int h = hx->halo;
    if (h * (int)sizeof(T) % 16 != 0 || 2 * h + sa.hh > window || n_inner > h)
      return fail(MT_EINVAL, "halo %d does not fit the window kernel (steps %d, window %d)", h,
                  n_inner, window);
  }
  const bool fast = FastAllowed<T>::value && (p.flags & MT_FLAG_FAST_MATH) != 0;
  const long long slots = (long long)PIPE_CTAS * e->sm_count;
  const int grid = (int)(n_tiles < slots ? n_tiles : slots);
  int rc;
  if (e->cfg.model == MT_MODEL_ARZ)
    rc = U == 2 ? launch_segfused_fast<T, 1, 2>(sa, grid, fast, e->uniform, e->lamk, stream)
                : launch_segfused_fast<T, 1, 1>(sa, grid, fast, e->uniform, e->lamk, stream);
  else
    rc = U == 2 ? launch_segfused_fast<T, 0, 2>(sa, grid, fast, e->uniform, e->lamk, stream)
                : launch_segfused_fast<T, 0, 1>(sa, grid, fast, e->uniform, e->lamk, stream);
  if (rc != MT_OK) return rc;
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MT_OK;
}

int segfused_any(mt_engine* e, const void* obs_in, const void* action, long long action_stride,
                 void* obs_out, int n_inner, const SegHalo* hx, cudaStream_t stream) {
  note_forget(e->cfg.device, stream);
  if (e->cfg.dtype == MT_F64)
    return segfused_typed<double>(e, obs_in, action, action_stride, obs_out, n_inner, hx, stream);
  return segfused_typed<float>(e, obs_in, action, action_stride, obs_out, n_inner, hx, stream);
}

bool can_segfuse(const mt_engine* e, const void* obs_in, const void* obs_out) {
  int U, nc, window;
  if (e->cfg.dtype == MT_F64) return segfused_geometry<double>(e, obs_in, obs_out, U, nc, window);
  return segfused_geometry<float>(e, obs_in, obs_out, U, nc, window);
}

// steps fused per launch by mt_rollout on long roads (MT_SEGFUSE_STEPS: A/B knob)
int segfuse_chunk() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("MT_SEGFUSE_STEPS");
    v = s ? atoi(s) : 16;
    if (v < 1) v = 1;
    if (v > SEGFUSED_MAX_STEPS) v = SEGFUSED_MAX_STEPS;
  }
  return v;
}

// Steps roads [env0, env0 + n_env) of the batch; all pointers address the
// whole batch (road 0) and are offset here.
template <typename T>
int step_typed(mt_engine* e, const void* obs_in_, const void* action_, void* obs_out_,
               void* reward_out_, long long env0, long long n_env, cudaStream_t stream,
               int n_inner = 1, long long action_stride = -1, long long reward_stride = 0) {
  // (see the early-load bookkeeping in the pipelined branch)
  const StreamNote prev = note_take(e->cfg.device, stream);
  uint16_t* const obs16 = e->obs16;
  constexpr int V = Vec<T>::N;
  const mt_params& p = e->params;
  const long long B = n_env;
  const int N = (int)e->cfg.n_cells;
  const T* obs_in = (const T*)obs_in_ + env0 * 2 * N;
  T* obs_out = (T*)obs_out_ + env0 * 2 * N;
  const T* action = action_ ? (const T*)action_ + env0 : nullptr;
  T* reward_out = reward_out_ ? (T*)reward_out_ + env0 : nullptr;
  const T* table = (const T*)e->table + env0 * TBL_STRIDE;
  const int* flags = e->flags + env0;
  T* partial = e->partial ? (T*)e->partial + env0 * e->max_tiles * 2 : nullptr;
  const bool aligned = (((uintptr_t)obs_in | (uintptr_t)obs_out) & 15u) == 0;
  const Geometry g = plan(B, N, V, aligned);
  const bool fast = FastAllowed<T>::value && (p.flags & MT_FLAG_FAST_MATH) != 0;
  const bool rew = reward_out != nullptr;

  BcArgs<T> bc;
  bc.bc_l = p.bc_left;
  bc.bc_r = p.bc_right;
  bc.bl0 = (T)p.bc_left_val[0];
  bc.bl1 = (T)p.bc_left_val[1];
  bc.br0 = (T)p.bc_right_val[0];
  bc.br1 = (T)p.bc_right_val[1];
  const T step = (T)(p.dt / p.dx);

  if (g.grid > 0x7fffffffLL) return fail(MT_EINVAL, "batch too large for one launch");
  if (e->cfg.model == MT_MODEL_NONLOCAL) {
    if (n_inner != 1) return fail(MT_EINVAL, "fused multi-step needs the pipelined kernel");
    return step_nonlocal<T>(e, obs_in, action, obs_out, reward_out, table, flags, B, aligned, bc,
                            step, stream);
  }
  if (g.vector_ok && g.tiles == 1 && !(p.flags & MT_FLAG_NO_PIPELINE)) {
    // persistent TMA-pipelined kernel: tiles of consecutive roads
    int U = pipe_vectors_per_thread();
    if (U == 0) U = (sizeof(T) == 8) ? 2 : 1;
    if (N % (U * V) != 0 || N < 2 * U * V) U = 1;
    PipeArgs<T> pa;
    pa.in = obs_in;
    pa.out = obs_out;
    pa.table = table;
    pa.flags = flags;
    pa.action = action;
    pa.reward = reward_out;
    if (B > 0x7fffffffLL) return fail(MT_EINVAL, "batch too large for one launch");
    pa.B = (int)B;
    pa.N = N;
    pa.nthr = N / (U * V);
    pa.nc = PIPE_BLOCK / U;
    pa.l2_ahead = pipe_l2_ahead();
    {
      static int early_l2 = -1;   // MT_PIPE_EARLY_L2: A/B knob; 2 measured best (0..14 tried)
      if (early_l2 < 0) { const char* s = getenv("MT_PIPE_EARLY_L2"); early_l2 = s ? atoi(s) : 2; }
      pa.early_l2 = early_l2;
    }
    {
      // MT_L2_HINT=0: plain tile loads (A/B knob).  Default: a tile is read once per
      // step, so its lines are marked evict_first and leave L2 before the lines
      // the stores have just written (measured: +1.3 % streaming, +4 % chained).
      static int l2_hint = -1;
      if (l2_hint < 0) { const char* s = getenv("MT_L2_HINT"); l2_hint = s ? atoi(s) : 1; }
      pa.l2_hint = l2_hint;
    }
    pa.n_inner = n_inner;
    pa.action_stride = action_stride >= 0 ? action_stride : e->cfg.n_envs;
    pa.reward_stride = reward_stride;
    pa.tpe = pa.nthr <= 32 ? next_pow2(pa.nthr) : ((pa.nthr + 31) / 32) * 32;
    pa.epc = pa.nc / pa.tpe;
    if (pa.epc > PIPE_MAX_EPC) pa.epc = PIPE_MAX_EPC;
    // per-road parameter table: the tile's table rows ride in the input stage
    // behind its roads when they fit (a 16 KB stage holds 16000 B of a
    // 1000-cell f64 road); otherwise the consumers read them from global memory
    int tbl_off = 0;
    if (!e->uniform && n_inner == 1 && pipe_stage_table()) {
      const long long row_bytes = 2LL * N * (long long)sizeof(T);
      const int tbl_row = TBL_STRIDE * (int)sizeof(T);
      const int fit = (int)(PIPE_TILE_BYTES / (row_bytes + tbl_row));
      if (fit >= 1) {
        if (pa.epc > fit) pa.epc = fit;
        tbl_off = PIPE_TILE_BYTES - pa.epc * tbl_row;
      }
    }
    const long long n_tiles = (B + pa.epc - 1) / pa.epc;
    if (n_tiles > 0x7fffffffLL) return fail(MT_EINVAL, "batch too large for one launch");
    pa.n_tiles = (int)n_tiles;
    pa.bc = bc;
    pa.step = step;
    pa.u = uniform_of<T>(e);
#ifdef MT_TRACE
    pa.trace = g_trace ? g_trace + (size_t)((g_trace_seq++) % 8) * 8 * 1024 : nullptr;
#endif
    const long long slots = (long long)PIPE_CTAS * e->sm_count;
    const int grid = (int)(n_tiles < slots ? n_tiles : slots);
    if (n_tiles >= (long long)pipe_dynamic_min_tiles() * grid) {
      int slot = 0;
      for (int j = 0; j < mt_engine::HOST_STREAMS; ++j)
        if (e->hs[j] != nullptr && stream == e->hs[j]) slot = j + 1;
      pa.sched = e->sched + 2 * slot;
    }
    // Early loads: the producer issues its first ring-full of loads BEFORE it
    // waits for the launch in front of it, so this launch's pipeline fills
    // while that one drains (stores still wait).  Allowed when that launch is
    // on record (note_take above), filled every CTA slot like this one -- a
    // CTA of this launch then only finds room once a CTA of the previous launch
    // has exited, i.e. after the previous launch has itself waited for
    // everything before it -- and writes nothing this launch reads: checked
    // here on the byte ranges, or vouched for by the caller
    // (MT_FLAG_INDEPENDENT_INPUT).  MT_FLAG_DEPENDENT_INPUT turns the check
    // off: every launch waits.  Consumers must not touch global memory before
    // the wait: scalar parameters, no speed limits; rewards only when the
    // producer thread stores them (roads wider than a warp), after its wait.
    const bool full_grid = (long long)grid == slots;
    const char* in_lo = (const char*)obs_in;
    const char* in_hi = in_lo + (size_t)B * 2 * N * sizeof(T);
    const bool disjoint = prev.out_hi <= in_lo || in_hi <= prev.out_lo;
    const bool independent = (p.flags & MT_FLAG_INDEPENDENT_INPUT) ||
                             (!(p.flags & MT_FLAG_DEPENDENT_INPUT) && auto_independent() && disjoint);
    const bool is_ring = pipe_use_ring(sizeof(T) == 4, fast, e->cfg.model == MT_MODEL_ARZ,
                                       pa.sched != nullptr, e->uniform);
    pa.tbl_off = is_ring ? 0 : tbl_off;
    // a reducer warp for the rewards of roads wider than a warp (static schedule)
    // (ARZ f64 is the instantiation with the static schedule: PipeDynamic in mt_pipe.cuh)
    pa.reducer = (reward_out != nullptr && pa.tpe > 32 && e->cfg.model == MT_MODEL_ARZ &&
                  sizeof(T) == 8 && !is_ring && n_inner == 1 && pipe_reward_reducer()) ? 1 : 0;
    // ... and for a reward row per step of a fused rollout (f64, 4 cells per thread)
    if (n_inner > 1 && reward_out != nullptr && reward_stride > 0 && pa.tpe > 32 &&
        sizeof(T) == 8 && U == 2 && pipe_reward_reducer())
      pa.reducer = 1;
    if (obs16 != nullptr && !is_ring && n_inner == 1) {
      pa.obs16 = obs16 + env0 * 2 * N;
      e->obs16_written = true;
    }
    pa.early_loads = (independent && e->uniform && action == nullptr && pa.obs16 == nullptr &&
                      (reward_out == nullptr || pa.tpe > 32) && n_inner == 1 && pipe_use_pdl() && full_grid && !is_ring &&
                      prev.full_pipe) ? 1 : 0;
    StreamNote mine;
    mine.device = e->cfg.device;
    mine.stream = stream;
    mine.full_pipe = full_grid && n_inner == 1 && !is_ring && pipe_use_pdl();
    mine.out_lo = (const char*)obs_out;
    mine.out_hi = mine.out_lo + (size_t)B * 2 * N * sizeof(T);
    int rc;
    if (e->cfg.model == MT_MODEL_ARZ)
      rc = U == 2 ? launch_pipe_model<T, 1, 2>(pa, grid, fast, e->uniform, e->lamk, stream, e->table_lamk)
                  : launch_pipe_model<T, 1, 1>(pa, grid, fast, e->uniform, e->lamk, stream, e->table_lamk);
    else
      rc = U == 2 ? launch_pipe_model<T, 0, 2>(pa, grid, fast, e->uniform, e->lamk, stream, e->table_lamk)
                  : launch_pipe_model<T, 0, 1>(pa, grid, fast, e->uniform, e->lamk, stream, e->table_lamk);
    if (rc != MT_OK) return rc;
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    note_put(mine);
    return MT_OK;
  }

  if (n_inner != 1) return fail(MT_EINVAL, "fused multi-step needs the pipelined kernel");
  if (g.vector_ok && g.tiles > 1 && !(p.flags & MT_FLAG_NO_PIPELINE) && B <= 0x7fffffffLL) {
    // a road longer than one tile: segments of it flow through the pipeline
    int U = pipe_vectors_per_thread();
    if (U == 0) U = (sizeof(T) == 8) ? 2 : 1;
    if (N % (U * V) != 0) U = 1;
    SegArgs<T> sa;
    sa.in = obs_in; sa.out = obs_out; sa.table = table; sa.flags = flags; sa.action = action;
    sa.partial = rew ? partial : nullptr;
    sa.B = (int)B; sa.N = N;
    sa.nc = PIPE_BLOCK / U;
    const int seg_cells = sa.nc * U * V;
    sa.tiles_per_row = (N + seg_cells - 1) / seg_cells;
    const long long n_tiles = B * (long long)sa.tiles_per_row;
    if (n_tiles <= 0x7fffffffLL && sa.tiles_per_row <= e->max_tiles) {
      sa.n_tiles = (int)n_tiles;
      sa.bc = bc; sa.step = step; sa.u = uniform_of<T>(e);
      const long long slots = 2LL * e->sm_count;
      const int grid = (int)(n_tiles < slots ? n_tiles : slots);
      int rc;
      if (e->cfg.model == MT_MODEL_ARZ)
        rc = U == 2 ? launch_seg_fast<T, 1, 2>(sa, grid, fast, e->uniform, e->lamk, stream)
                    : launch_seg_fast<T, 1, 1>(sa, grid, fast, e->uniform, e->lamk, stream);
      else
        rc = U == 2 ? launch_seg_fast<T, 0, 2>(sa, grid, fast, e->uniform, e->lamk, stream)
                    : launch_seg_fast<T, 0, 1>(sa, grid, fast, e->uniform, e->lamk, stream);
      if (rc != MT_OK) return rc;
      e->launches++;
      if (rew) {
        const int wpb = 8;
        reward_finalize_kernel<T><<<(unsigned)((B + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
            (const T*)partial, reward_out, B, sa.tiles_per_row);
        e->launches++;
      }
      CUDA_TRY(cudaGetLastError());
      return MT_OK;
    }
  }

  StepArgs<T> a;
  a.in = obs_in;
  a.out = obs_out;
  a.table = table;
  a.flags = flags;
  a.action = action;
  a.reward = reward_out;
  a.partial = partial;
  a.B = B;
  a.N = N;
  a.nvec = g.nvec;
  a.tpe = g.tpe;
  a.epc = g.epc;
  a.tiles = g.tiles;
  a.bc = bc;
  a.step = step;
  const int wpb = 8;  // warps per block of the per-env helper kernels
  if (g.vector_ok) {
    const dim3 grid((unsigned)g.grid), block(g.block);
    if (e->cfg.model == MT_MODEL_ARZ) {
      if constexpr (FastAllowed<T>::value) {
        if (fast) step_vec_kernel<T, 1, true><<<grid, block, 0, stream>>>(a);
        else step_vec_kernel<T, 1, false><<<grid, block, 0, stream>>>(a);
      } else {
        step_vec_kernel<T, 1, false><<<grid, block, 0, stream>>>(a);
      }
    } else {
      if constexpr (FastAllowed<T>::value) {
        if (fast) step_vec_kernel<T, 0, true><<<grid, block, 0, stream>>>(a);
        else step_vec_kernel<T, 0, false><<<grid, block, 0, stream>>>(a);
      } else {
        step_vec_kernel<T, 0, false><<<grid, block, 0, stream>>>(a);
      }
    }
    e->launches++;
    if (rew && g.tiles > 1) {
      reward_finalize_kernel<T><<<(unsigned)((B + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
          (const T*)partial, reward_out, B, g.tiles);
      e->launches++;
    }
  } else {
    const long long cells = B * (long long)N;
    const int block = 256;
    const long long grid = (cells + block - 1) / block;
    if (grid > 0x7fffffffLL) return fail(MT_EINVAL, "batch too large for one launch");
    if (e->cfg.model == MT_MODEL_ARZ) {
      if (fast) arz_step_scalar_kernel<T, FastAllowed<T>::value><<<(unsigned)grid, block, 0, stream>>>(a);
      else arz_step_scalar_kernel<T, false><<<(unsigned)grid, block, 0, stream>>>(a);
    } else {
      if (fast) lwr_step_scalar_kernel<T, FastAllowed<T>::value><<<(unsigned)grid, block, 0, stream>>>(a);
      else lwr_step_scalar_kernel<T, false><<<(unsigned)grid, block, 0, stream>>>(a);
    }
    e->launches++;
    if (rew) {
      reward_rows_kernel<T><<<(unsigned)((B + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
          (const T*)obs_out, reward_out, B, N);
      e->launches++;
    }
  }
  CUDA_TRY(cudaGetLastError());
  return MT_OK;
}

// True when n_steps chained steps of this batch can run as ONE launch of the
// pipelined kernel with the state kept in registers between steps: LWR / ARZ,
// roads that fit one tile, 16-byte aligned buffers.
bool can_fuse_steps(const mt_engine* e, const void* obs_in, const void* obs_out) {
  if (e->cfg.model == MT_MODEL_NONLOCAL || (e->params.flags & MT_FLAG_NO_PIPELINE)) return false;
  // the reference's abort rule needs a look at every step's input (mt_abort.cuh)
  if (e->cfg.model == MT_MODEL_ARZ && (e->params.flags & MT_FLAG_FSOLVE_ABORT)) return false;
  const int V = (int)(16 / e->elt);
  const bool aligned = (((uintptr_t)obs_in | (uintptr_t)obs_out) & 15u) == 0;
  const Geometry g = plan(e->cfg.n_envs, (int)e->cfg.n_cells, V, aligned);
  const char* off = getenv("MT_NO_FUSE");
  return g.vector_ok && g.tiles == 1 && !(off && off[0] == '1');
}

// MT_FLAG_FSOLVE_ABORT (mt_abort.cuh): after the step, roads whose relaxation
// system the reference could not solve get the reference's result
template <typename T>
int abort_fixup(mt_engine* e, const void* obs_in, const void* action, void* obs_out,
                void* reward_out, uint16_t* obs16, long long env0, long long n_env,
                cudaStream_t stream) {
  const mt_params& p = e->params;
  const long long N = e->cfg.n_cells;
  AbortArgs<T> a;
  a.in = (const T*)obs_in + env0 * 2 * N;
  a.out = (T*)obs_out + env0 * 2 * N;
  a.table = (const T*)e->table + env0 * TBL_STRIDE;
  a.flags = e->flags + env0;
  a.action = action ? (const T*)action + env0 : nullptr;
  a.reward = reward_out ? (T*)reward_out + env0 : nullptr;
  a.obs16 = obs16 ? obs16 + env0 * 2 * N : nullptr;
  a.aborted = nullptr;
  a.B = n_env;
  a.N = (int)N;
  a.bc.bc_l = p.bc_left; a.bc.bc_r = p.bc_right;
  a.bc.bl0 = (T)p.bc_left_val[0]; a.bc.bl1 = (T)p.bc_left_val[1];
  a.bc.br0 = (T)p.bc_right_val[0]; a.bc.br1 = (T)p.bc_right_val[1];
  a.step = (T)(p.dt / p.dx);
  a.dt = (T)p.dt;
  a.tau_scalar = (T)p.tau;
  a.tau_env = p.tau_env ? (const T*)p.tau_env + env0 : nullptr;
  const int wpb = 8;
  const long long grid = (n_env + wpb - 1) / wpb;
  if (grid > 0x7fffffffLL) return fail(MT_EINVAL, "batch too large for one launch");
  arz_abort_fixup_kernel<T><<<(unsigned)grid, wpb * 32, 0, stream>>>(a);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MT_OK;
}

int step_range(mt_engine* e, const void* obs_in, const void* action, void* obs_out,
               void* reward_out, long long env0, long long n_env, cudaStream_t stream,
               int n_inner = 1, long long action_stride = -1, long long reward_stride = 0) {
  e->obs16_written = false;
  const int rc = e->cfg.dtype == MT_F64
      ? step_typed<double>(e, obs_in, action, obs_out, reward_out, env0, n_env, stream, n_inner,
                           action_stride, reward_stride)
      : step_typed<float>(e, obs_in, action, obs_out, reward_out, env0, n_env, stream, n_inner,
                          action_stride, reward_stride);
  uint16_t* const obs16 = e->obs16;
  e->obs16 = nullptr;
  if (rc != MT_OK) return rc;
  const long long row = 2LL * e->cfg.n_cells;
  if (obs16 != nullptr && !e->obs16_written) {
    // this batch did not take the kernel that writes the copy itself
    const long long n = n_env * row;
    const int block = 256;
    const long long grid = (n / 4 + block) / block;
    if (grid > 0x7fffffffLL) return fail(MT_EINVAL, "batch too large for one launch");
    if (e->cfg.dtype == MT_F64)
      obs_to_bf16_kernel<double><<<(unsigned)grid, block, 0, stream>>>(
          (const double*)obs_out + env0 * row, obs16 + env0 * row, n);
    else
      obs_to_bf16_kernel<float><<<(unsigned)grid, block, 0, stream>>>(
          (const float*)obs_out + env0 * row, obs16 + env0 * row, n);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  if ((e->params.flags & MT_FLAG_FSOLVE_ABORT) && e->cfg.model == MT_MODEL_ARZ && n_inner == 1) {
    const int rc2 = e->cfg.dtype == MT_F64
        ? abort_fixup<double>(e, obs_in, action, obs_out, reward_out, obs16, env0, n_env, stream)
        : abort_fixup<float>(e, obs_in, action, obs_out, reward_out, obs16, env0, n_env, stream);
    if (rc2 != MT_OK) return rc2;
  }
  return MT_OK;
}

int step_any(mt_engine* e, const void* obs_in, const void* action, void* obs_out,
             void* reward_out, cudaStream_t stream) {
  if (e->params_pending) {
    if (stream != e->params_stream) CUDA_TRY(cudaStreamWaitEvent(stream, e->params_ready, 0));
    e->params_pending = false;
  }
  return step_range(e, obs_in, action, obs_out, reward_out, 0, e->cfg.n_envs, stream);
}

bool valid_bc(int bc) { return bc >= MT_BC_LOOP && bc <= MT_BC_HALO; }

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int mt_version(void) { return MT_VERSION; }

const char* mt_last_error(void) { return g_err; }

int mt_create(const mt_config* cfg, mt_engine** out) {
  if (!cfg || !out) return fail(MT_EINVAL, "mt_create: NULL argument");
  *out = nullptr;
  if (cfg->model < MT_MODEL_LWR || cfg->model > MT_MODEL_NONLOCAL)
    return fail(MT_EINVAL, "mt_create: unknown model %d", cfg->model);
  if (cfg->dtype != MT_F32 && cfg->dtype != MT_F64)
    return fail(MT_EINVAL, "mt_create: unknown dtype %d", cfg->dtype);
  if (cfg->n_envs < 1) return fail(MT_EINVAL, "mt_create: n_envs must be >= 1");
  if (cfg->n_cells < 3 || cfg->n_cells > 0x3fffffff)
    return fail(MT_EINVAL, "mt_create: n_cells must be in [3, 2^30)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(MT_ECUDA, "mt_create: no CUDA device (this engine has no CPU fallback)");
  }
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(MT_EINVAL, "mt_create: device %d out of range (%d devices)", cfg->device, ndev);
  CUDA_TRY(cudaSetDevice(cfg->device));
  mt_engine* e = new (std::nothrow) mt_engine();
  if (!e) return fail(MT_ENOMEM, "mt_create: out of host memory");
  e->cfg = *cfg;
  cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
  e->elt = cfg->dtype == MT_F64 ? 8 : 4;
  const size_t B = (size_t)cfg->n_envs, N = (size_t)cfg->n_cells;
  const int V = (int)(16 / e->elt);
  const Geometry g = plan((long long)B, (int)N, V, true);
  e->max_tiles = g.tiles;
  cudaError_t err = cudaMalloc(&e->table, B * TBL_STRIDE * e->elt);
  if (err == cudaSuccess) err = cudaMalloc((void**)&e->flags, B * sizeof(int));
  if (err == cudaSuccess) err = cudaMalloc((void**)&e->summary, 2 * sizeof(unsigned));
  if (err == cudaSuccess && g.tiles > 1) err = cudaMalloc(&e->partial, B * g.tiles * 2 * e->elt);
  const size_t sched_bytes = 2 * (1 + mt_engine::HOST_STREAMS) * sizeof(unsigned int);
  if (err == cudaSuccess) err = cudaMalloc((void**)&e->sched, sched_bytes);
  if (err == cudaSuccess) err = cudaMemset(e->sched, 0, sched_bytes);
  if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->params_ready, cudaEventDisableTiming);
  if (err == cudaSuccess && cfg->host_staging) {
    for (int i = 0; i < mt_engine::HOST_STREAMS && err == cudaSuccess; ++i)
      err = cudaStreamCreateWithFlags(&e->hs[i], cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->action_ready, cudaEventDisableTiming);
    for (int i = 0; i < mt_engine::MAX_CHUNKS && err == cudaSuccess; ++i) {
      err = cudaEventCreateWithFlags(&e->ev_up[i], cudaEventDisableTiming);
      if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming);
    }
    if (err == cudaSuccess) err = cudaMalloc(&e->d_a, B * 2 * N * e->elt);
    if (err == cudaSuccess) err = cudaMalloc(&e->d_b, B * 2 * N * e->elt);
    if (err == cudaSuccess) err = cudaMalloc(&e->d_action, B * e->elt);
    if (err == cudaSuccess) err = cudaMalloc(&e->d_reward, B * e->elt);
  }
  if (err != cudaSuccess) {
    const int code = err == cudaErrorMemoryAllocation ? MT_ENOMEM : MT_ECUDA;
    fail(code, "mt_create: %s", cudaGetErrorString(err));
    mt_destroy(e);
    return code;
  }
  *out = e;
  return MT_OK;
}

void mt_destroy(mt_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaFree(e->table);
  cudaFree(e->flags);
  cudaFree(e->summary);
  cudaFree(e->partial);
  cudaFree(e->sched);
  cudaFree(e->d_a);
  cudaFree(e->d_b);
  cudaFree(e->d_action);
  cudaFree(e->d_reward);
  if (e->stream) cudaStreamDestroy(e->stream);
  for (int i = 0; i < mt_engine::HOST_STREAMS; ++i)
    if (e->hs[i]) cudaStreamDestroy(e->hs[i]);
  if (e->action_ready) cudaEventDestroy(e->action_ready);
  for (int i = 0; i < mt_engine::MAX_CHUNKS; ++i) {
    if (e->ev_up[i]) cudaEventDestroy(e->ev_up[i]);
    if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
  }
  if (e->params_ready) cudaEventDestroy(e->params_ready);
  delete e;
}

int mt_set_params(mt_engine* e, const mt_params* p, void* stream) {
  if (!e || !p) return fail(MT_EINVAL, "mt_set_params: NULL argument");
  if (!valid_bc(p->bc_left) || !valid_bc(p->bc_right))
    return fail(MT_EINVAL, "Unknown boundary conditions: (%d, %d)", p->bc_left, p->bc_right);
  if (!(p->dx > 0) || !(p->dt > 0)) return fail(MT_EINVAL, "mt_set_params: dx and dt must be > 0");
  if (p->scheme != MT_SCHEME_GODUNOV && p->scheme != MT_SCHEME_LXF)
    return fail(MT_EINVAL, "mt_set_params: unknown scheme %d", p->scheme);
  if (p->scheme == MT_SCHEME_LXF && e->cfg.model != MT_MODEL_NONLOCAL)
    return fail(MT_EINVAL, "mt_set_params: MT_SCHEME_LXF belongs to the non-local model");
  if (e->cfg.model == MT_MODEL_NONLOCAL) {
    if (p->n_eta < 1 || p->n_eta > e->cfg.max_n_eta || !p->gamma)
      return fail(MT_EINVAL, "mt_set_params: n_eta=%d outside [1, %d] or gamma NULL",
                  p->n_eta, e->cfg.max_n_eta);
  }
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const long long B = e->cfg.n_envs;
  const int block = 128;
  const unsigned grid = (unsigned)((B + block - 1) / block);
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(e->summary, 0, 2 * sizeof(unsigned), s));
  if (e->cfg.dtype == MT_F64) {
    ParamArgs<double> a{(double*)e->table, e->flags, B, e->cfg.model,
                        (p->flags & MT_FLAG_FAST_MATH) ? 1 : 0, p->rho_max, p->v_max,
                        p->tau, p->lam, p->rho_crit, p->dt,
                        (const double*)p->rho_max_env, (const double*)p->v_max_env,
                        (const double*)p->tau_env, (const double*)p->lam_env,
                        (const double*)p->rho_crit_env, e->summary};
    set_params_kernel<double><<<grid, block, 0, s>>>(a);
  } else {
    ParamArgs<float> a{(float*)e->table, e->flags, B, e->cfg.model,
                       (p->flags & MT_FLAG_FAST_MATH) ? 1 : 0, p->rho_max, p->v_max,
                       p->tau, p->lam, p->rho_crit, p->dt,
                       (const float*)p->rho_max_env, (const float*)p->v_max_env,
                       (const float*)p->tau_env, (const float*)p->lam_env,
                       (const float*)p->rho_crit_env, e->summary};
    set_params_kernel<float><<<grid, block, 0, s>>>(a);
  }
  note_forget(e->cfg.device, s);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(e->params_ready, s));
  e->params_stream = s;
  e->params_pending = true;
  e->params = *p;
  e->have_params = true;
  e->uniform = !p->rho_max_env && !p->v_max_env && !p->tau_env && !p->lam_env && !p->rho_crit_env;
  if (e->uniform) {
    int lamk = LAM_GENERAL;
    if (e->cfg.dtype == MT_F64) make_uniform<double>(e->u64, lamk, e->cfg.model, *p);
    else make_uniform<float>(e->u32, lamk, e->cfg.model, *p);
    e->lamk = lamk;
    if (lamk == LAM_GENERAL) e->uniform = false;  // pow(): use the device table
  }
  // Per-road arrays: which exponent kind do ALL roads share?  One 8-byte
  // readback (and a wait for the kernel above) buys a step kernel without the
  // run-time dispatch -- and, when every rho_max is 1, without the division
  // rho / rho_max.  Not during stream capture (no waiting there): such engines
  // keep the general kernel.
  e->table_lamk = LAM_RUNTIME;
  const bool arrays = p->rho_max_env || p->v_max_env || p->tau_env || p->lam_env || p->rho_crit_env;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (arrays && cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) {
    unsigned host[2] = {0u, 0u};
    CUDA_TRY(cudaMemcpyAsync(host, e->summary, sizeof(host), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (host[0] == (1u << LAM_ONE)) e->table_lamk = host[1] ? LAM_ONE : LAM_ONE_RM1;
    else if (host[0] == (1u << LAM_TWO)) e->table_lamk = LAM_TWO;
    else if (host[0] == (1u << LAM_HALF)) e->table_lamk = LAM_HALF;
  }
  return MT_OK;
}

int mt_step(mt_engine* e, const void* obs_in, const void* action, void* obs_out,
            void* reward_out, void* stream) {
  if (!e || !obs_in || !obs_out) return fail(MT_EINVAL, "mt_step: NULL argument");
  if (!e->have_params) return fail(MT_ESTATE, "mt_step: call mt_set_params first");
  if (obs_in == obs_out) return fail(MT_EINVAL, "mt_step: obs_in and obs_out must differ");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  return step_any(e, obs_in, action, obs_out, reward_out, (cudaStream_t)stream);
}

int mt_step_obs16(mt_engine* e, const void* obs_in, const void* action, void* obs_out,
                  void* reward_out, void* obs_out_bf16, void* stream) {
  if (!e || !obs_in || !obs_out) return fail(MT_EINVAL, "mt_step_obs16: NULL argument");
  if (!e->have_params) return fail(MT_ESTATE, "mt_step_obs16: call mt_set_params first");
  if (obs_in == obs_out) return fail(MT_EINVAL, "mt_step_obs16: obs_in and obs_out must differ");
  if (((uintptr_t)obs_out_bf16 & 7u) != 0)
    return fail(MT_EINVAL, "mt_step_obs16: the bfloat16 batch must be 8-byte aligned");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  e->obs16 = (uint16_t*)obs_out_bf16;
  const int rc = step_any(e, obs_in, action, obs_out, reward_out, (cudaStream_t)stream);
  e->obs16 = nullptr;
  return rc;
}

int mt_actor_head(int32_t device, int32_t dtype, const void* hidden_bf16, const void* weight_bf16,
                  const void* bias_bf16, void* action_out, int64_t n_envs, int32_t n_hidden,
                  double scale, const void* reward_prev, void* return_sum, void* stream) {
  if (!hidden_bf16 || !weight_bf16 || !bias_bf16 || !action_out)
    return fail(MT_EINVAL, "mt_actor_head: NULL argument");
  if (n_envs < 0 || n_hidden < 1) return fail(MT_EINVAL, "mt_actor_head: bad sizes");
  if (dtype != MT_F32 && dtype != MT_F64) return fail(MT_EINVAL, "mt_actor_head: unknown dtype");
  if ((reward_prev == nullptr) != (return_sum == nullptr))
    return fail(MT_EINVAL, "mt_actor_head: reward_prev and return_sum go together");
  CUDA_TRY(cudaSetDevice(device));
  if (n_envs == 0) return MT_OK;
  const int wpb = 8;
  const long long grid = (n_envs + wpb - 1) / wpb;
  if (grid > 0x7fffffffLL) return fail(MT_EINVAL, "batch too large for one launch");
  note_forget(device, stream);
  if (dtype == MT_F64)
    actor_head_kernel<double><<<(unsigned)grid, wpb * 32, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)hidden_bf16, (const __nv_bfloat16*)weight_bf16,
        (const __nv_bfloat16*)bias_bf16, (double*)action_out, n_envs, n_hidden, (float)scale,
        (const double*)reward_prev, (double*)return_sum);
  else
    actor_head_kernel<float><<<(unsigned)grid, wpb * 32, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)hidden_bf16, (const __nv_bfloat16*)weight_bf16,
        (const __nv_bfloat16*)bias_bf16, (float*)action_out, n_envs, n_hidden, (float)scale,
        (const float*)reward_prev, (float*)return_sum);
  CUDA_TRY(cudaGetLastError());
  return MT_OK;
}

int mt_rollout(mt_engine* e, void* obs_a, void* obs_b, const void* actions, void* rewards,
               int32_t n_steps, void* stream, int32_t* result_in_b) {
  if (!e || !obs_a || !obs_b) return fail(MT_EINVAL, "mt_rollout: NULL argument");
  if (!e->have_params) return fail(MT_ESTATE, "mt_rollout: call mt_set_params first");
  if (obs_a == obs_b) return fail(MT_EINVAL, "mt_rollout: obs_a and obs_b must differ");
  if (n_steps < 0) return fail(MT_EINVAL, "mt_rollout: n_steps must be >= 0");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const size_t stride = (size_t)e->cfg.n_envs * e->elt;
  if (n_steps >= 2 && can_fuse_steps(e, obs_a, obs_b)) {
    // one launch: every tile is read once, stepped n_steps times on chip
    // (per-step actions are read as they are needed) and written once
    if (e->params_pending) {
      if ((cudaStream_t)stream != e->params_stream)
        CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, e->params_ready, 0));
      e->params_pending = false;
    }
    // (a reward row per step, when asked for, is reduced in the same launch)
    const int rc = step_range(e, obs_a, actions, obs_b, rewards, 0, e->cfg.n_envs,
                              (cudaStream_t)stream, n_steps, -1, rewards ? e->cfg.n_envs : 0);
    if (rc != MT_OK) return rc;
    if (result_in_b) *result_in_b = 1;
    return MT_OK;
  }
  void* cur = obs_a;
  void* nxt = obs_b;
  if (n_steps >= 2 && !rewards && can_segfuse(e, obs_a, obs_b)) {
    // a road longer than one tile: overlapping windows of it advance by up to
    // segfuse_chunk() steps per launch (mt_segfused.cuh)
    if (e->params_pending) {
      if ((cudaStream_t)stream != e->params_stream)
        CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, e->params_ready, 0));
      e->params_pending = false;
    }
    for (int done = 0; done < n_steps;) {
      const int k = n_steps - done < segfuse_chunk() ? n_steps - done : segfuse_chunk();
      const void* act = actions ? (const char*)actions + (size_t)done * stride : nullptr;
      const int rc = segfused_any(e, cur, act, e->cfg.n_envs, nxt, k, nullptr, (cudaStream_t)stream);
      if (rc != MT_OK) return rc;
      void* tmp = cur; cur = nxt; nxt = tmp;
      done += k;
    }
    if (result_in_b) *result_in_b = (cur == obs_b) ? 1 : 0;
    return MT_OK;
  }
  for (int s = 0; s < n_steps; ++s) {
    const void* act = actions ? (const char*)actions + (size_t)s * stride : nullptr;
    void* rew = rewards ? (char*)rewards + (size_t)s * stride : nullptr;
    const int rc = step_any(e, cur, act, nxt, rew, (cudaStream_t)stream);
    if (rc != MT_OK) return rc;
    void* tmp = cur; cur = nxt; nxt = tmp;
  }
  if (result_in_b) *result_in_b = (cur == obs_b) ? 1 : 0;
  return MT_OK;
}

int mt_step_host(mt_engine* e, const void* obs_in_host, const void* action_host,
                 void* obs_out_host, void* reward_out_host, int32_t n_steps) {
  if (!e || !obs_in_host || !obs_out_host) return fail(MT_EINVAL, "mt_step_host: NULL argument");
  if (!e->have_params) return fail(MT_ESTATE, "mt_step_host: call mt_set_params first");
  if (!e->d_a) return fail(MT_ESTATE, "mt_step_host: engine created without host_staging");
  if (n_steps < 1) return fail(MT_EINVAL, "mt_step_host: n_steps must be >= 1");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const long long B = e->cfg.n_envs;
  const size_t N = (size_t)e->cfg.n_cells;
  const size_t row_bytes = 2 * N * e->elt;
  constexpr int NS = mt_engine::HOST_STREAMS;
  // Roads are independent, so the batch is cut into chunks that flow through
  // upload -> n_steps kernels -> download, each stage on its own stream
  // (hs[0], hs[1], hs[2]) with one event per chunk between stages: the two
  // copy engines never wait for each other, PCIe carries both directions at
  // once and the kernels hide behind the copies.
  static long long chunk_bytes = 0;
  if (chunk_bytes == 0) {
    const char* s = getenv("MT_HOST_CHUNK_MB");
    chunk_bytes = (long long)((s ? atof(s) : 8.0) * (1 << 20));
    if (chunk_bytes < (1 << 16)) chunk_bytes = 1 << 16;
  }
  long long n_chunks = (B * (long long)row_bytes) / chunk_bytes;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > 64) n_chunks = 64;
  if (n_chunks > B) n_chunks = B;
  const long long per = (B + n_chunks - 1) / n_chunks;
  // Chunk plan.  Nothing comes back before the first chunk is up, and nothing
  // goes up while the last one comes back: the first and last chunks are small
  // (per/8, per/4, per/2) so that this fill and drain cost little, the middle
  // ones are full-size so that the per-copy overheads stay amortised.
  long long plan_n[mt_engine::MAX_CHUNKS];
  int n_plan = 0;
  {
    long long left = B;
    long long ramp[3] = {per >> 3, per >> 2, per >> 1};
    static const bool want_graded = !(getenv("MT_HOST_GRADED") && atoi(getenv("MT_HOST_GRADED")) == 0);
    const bool graded = want_graded && n_chunks >= 4 && ramp[0] >= 1;
    long long tail[3] = {0, 0, 0};
    if (graded) {
      for (int i = 0; i < 3; ++i) { plan_n[n_plan++] = ramp[i]; left -= ramp[i]; }
      for (int i = 0; i < 3; ++i) { tail[i] = ramp[2 - i]; left -= tail[i]; }
    }
    while (left > 0) {
      const long long n = left < per ? left : per;
      plan_n[n_plan++] = n;
      left -= n;
    }
    if (graded)
      for (int i = 0; i < 3; ++i) plan_n[n_plan++] = tail[i];
  }
  for (int i = 0; i < NS; ++i) CUDA_TRY(cudaStreamWaitEvent(e->hs[i], e->params_ready, 0));
  e->params_pending = false;
  if (action_host) {
    CUDA_TRY(cudaMemcpyAsync(e->d_action, action_host, (size_t)B * e->elt,
                             cudaMemcpyHostToDevice, e->hs[0]));
    CUDA_TRY(cudaEventRecord(e->action_ready, e->hs[0]));
    for (int i = 1; i < NS; ++i) CUDA_TRY(cudaStreamWaitEvent(e->hs[i], e->action_ready, 0));
  }
  // several steps on one upload: fused into one launch per chunk when possible
  // (the same speed limits apply to every step: action stride 0)
  const bool fuse = n_steps >= 2 && can_fuse_steps(e, e->d_a, e->d_b);
  const bool odd = (n_steps & 1) != 0;
  char* result = (char*)((odd || fuse) ? e->d_b : e->d_a);
  cudaStream_t s_up = e->hs[0], s = e->hs[1], s_down = e->hs[2];
  long long env0 = 0;
  for (int c = 0; c < n_plan; env0 += plan_n[c], ++c) {
    const long long n_env = plan_n[c];
    const size_t off = (size_t)env0 * row_bytes, bytes = (size_t)n_env * row_bytes;
    CUDA_TRY(cudaMemcpyAsync((char*)e->d_a + off, (const char*)obs_in_host + off, bytes,
                             cudaMemcpyHostToDevice, s_up));
    CUDA_TRY(cudaEventRecord(e->ev_up[c], s_up));
    CUDA_TRY(cudaStreamWaitEvent(s, e->ev_up[c], 0));
    void* cur = e->d_a;
    void* nxt = e->d_b;
    if (fuse) {
      const int rc = step_range(e, cur, action_host ? e->d_action : nullptr, nxt,
                                reward_out_host ? e->d_reward : nullptr, env0, n_env, s,
                                n_steps, 0);
      if (rc != MT_OK) return rc;
    } else {
      for (int i = 0; i < n_steps; ++i) {
        void* rew = (reward_out_host && i == n_steps - 1) ? e->d_reward : nullptr;
        const int rc = step_range(e, cur, action_host ? e->d_action : nullptr, nxt, rew, env0,
                                  n_env, s);
        if (rc != MT_OK) return rc;
        void* tmp = cur; cur = nxt; nxt = tmp;
      }
    }
    CUDA_TRY(cudaEventRecord(e->ev_done[c], s));
    CUDA_TRY(cudaStreamWaitEvent(s_down, e->ev_done[c], 0));
    CUDA_TRY(cudaMemcpyAsync((char*)obs_out_host + off, result + off, bytes,
                             cudaMemcpyDeviceToHost, s_down));
    if (reward_out_host)
      CUDA_TRY(cudaMemcpyAsync((char*)reward_out_host + (size_t)env0 * e->elt,
                               (char*)e->d_reward + (size_t)env0 * e->elt,
                               (size_t)n_env * e->elt, cudaMemcpyDeviceToHost, s_down));
  }
  for (int i = 0; i < NS; ++i) CUDA_TRY(cudaStreamSynchronize(e->hs[i]));
  return MT_OK;
}

int mt_density_sqerr(mt_engine* e, const void* obs_pred, const void* obs_target,
                     void* out_per_env, void* stream) {
  if (!e || !obs_pred || !obs_target || !out_per_env)
    return fail(MT_EINVAL, "mt_density_sqerr: NULL argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const long long B = e->cfg.n_envs;
  const int N = (int)e->cfg.n_cells, wpb = 8;
  const unsigned grid = (unsigned)((B + wpb - 1) / wpb);
  if (e->cfg.dtype == MT_F64)
    density_sqerr_kernel<double><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(
        (const double*)obs_pred, (const double*)obs_target, (double*)out_per_env, B, N);
  else
    density_sqerr_kernel<float><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(
        (const float*)obs_pred, (const float*)obs_target, (float*)out_per_env, B, N);
  note_forget(e->cfg.device, stream);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MT_OK;
}

int mt_nonfinite(mt_engine* e, const void* obs, int32_t* flags_out, void* stream) {
  if (!e || !obs || !flags_out) return fail(MT_EINVAL, "mt_nonfinite: NULL argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const long long B = e->cfg.n_envs;
  const int N = (int)e->cfg.n_cells, wpb = 8;
  const unsigned grid = (unsigned)((B + wpb - 1) / wpb);
  if (e->cfg.dtype == MT_F64)
    nonfinite_kernel<double><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>((const double*)obs,
                                                                          flags_out, B, N);
  else
    nonfinite_kernel<float><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>((const float*)obs,
                                                                         flags_out, B, N);
  note_forget(e->cfg.device, stream);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MT_OK;
}

int mt_aggregate(const void* time, const void* position, const void* speed, int64_t n_records,
                 double dt, double dx, int32_t n_rows, int32_t n_sections, double empty_speed,
                 void* scratch_idx, void* speed_out, void* density_out, int32_t device,
                 void* stream) {
  if (!time || !position || !speed || !scratch_idx || !speed_out || !density_out)
    return fail(MT_EINVAL, "mt_aggregate: NULL argument");
  if (n_records < 0 || n_rows < 1 || n_sections < 1 || !(dt > 0) || !(dx > 0))
    return fail(MT_EINVAL, "mt_aggregate: bad shape or step");
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  // the "rows out of order" flag is the scratch buffer's last word (no allocation here)
  int* flag = (int*)scratch_idx + 2 * n_records;
  cudaError_t err = cudaMemsetAsync(flag, 0, sizeof(int), s);
  AggArgs a;
  a.time = (const double*)time; a.pos = (const double*)position; a.speed = (const double*)speed;
  a.t_idx = (int*)scratch_idx; a.x_idx = (int*)scratch_idx + n_records;
  a.M = n_records; a.dt = dt; a.dx = dx; a.empty_speed = empty_speed;
  a.n_rows = n_rows; a.n_sections = n_sections;
  a.speed_out = (double*)speed_out; a.density_out = (double*)density_out;
  a.unsorted = flag;
  int unsorted = 0;
  if (err == cudaSuccess && n_records > 0) {
    const int block = 256;
    const unsigned grid = (unsigned)((n_records + block - 1) / block);
    agg_index_kernel<<<grid, block, 0, s>>>(a);
    agg_check_sorted_kernel<<<grid, block, 0, s>>>(a);
  }
  // (the error contract needs the flag on the host: one 4-byte copy and a wait)
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(&unsorted, flag, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (err == cudaSuccess) err = cudaStreamSynchronize(s);
  if (err == cudaSuccess && !unsorted) {
    const size_t per_warp = (((size_t)n_sections * 12 + 32 * 8) + 7) & ~(size_t)7;
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb > 48 * 1024) wpb >>= 1;
    if (per_warp * wpb <= 48 * 1024)
      agg_reduce_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, per_warp * wpb, s>>>(a);
    else
      agg_reduce_scan_kernel<<<(unsigned)n_rows, 128, 0, s>>>(a);
    err = cudaGetLastError();
  }
  if (err != cudaSuccess) return fail(MT_ECUDA, "mt_aggregate: %s", cudaGetErrorString(err));
  if (unsorted) return fail(MT_EINVAL, "mt_aggregate: records must be grouped by non-decreasing time row");
  return MT_OK;
}

#ifdef MT_TRACE
// experimental build only: 8 launches x 1024 CTAs x 8 stamps, round-robin by launch
int mt_debug_trace(void* dev_buf) { g_trace = (unsigned long long*)dev_buf; g_trace_seq = 0; return MT_OK; }
#endif

int mt_launch_count(const mt_engine* e, int64_t* out) {
  if (!e || !out) return fail(MT_EINVAL, "mt_launch_count: NULL argument");
  *out = e->launches;
  return MT_OK;
}

// ---------------------------------------------------------------------------
// peer-memory halo exchange (mt_halo.cuh)
// ---------------------------------------------------------------------------
int mt_halo_create(int32_t device, int32_t dtype, int32_t halo, mt_halo** out) {
  if (!out) return fail(MT_EINVAL, "mt_halo_create: NULL argument");
  if (halo < 1 || halo > (1 << 20)) return fail(MT_EINVAL, "mt_halo_create: halo out of range");
  if (dtype != MT_F32 && dtype != MT_F64) return fail(MT_EINVAL, "mt_halo_create: bad dtype");
  CUDA_TRY(cudaSetDevice(device));
  mt_halo* h = new (std::nothrow) mt_halo();
  if (!h) return fail(MT_ENOMEM, "mt_halo_create: out of host memory");
  h->device = device;
  h->dtype = dtype;
  h->halo = halo;
  h->bytes = sizeof(HaloBox) + (size_t)8 * halo * (dtype == MT_F64 ? 8 : 4);
  // a dedicated cudaMalloc allocation: exportable through CUDA IPC as it is
  cudaError_t err = cudaMalloc(&h->box, h->bytes);
  if (err == cudaSuccess) err = cudaMemset(h->box, 0, h->bytes);
  if (err != cudaSuccess) {
    fail(MT_ECUDA, "mt_halo_create: %s", cudaGetErrorString(err));
    delete h;
    return MT_ECUDA;
  }
  *out = h;
  return MT_OK;
}

void mt_halo_destroy(mt_halo* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (h->left_ipc) cudaIpcCloseMemHandle(h->left);
  if (h->right_ipc) cudaIpcCloseMemHandle(h->right);
  cudaFree(h->box);
  delete h;
}

int mt_halo_handle(const mt_halo* h, unsigned char handle_out[MT_IPC_HANDLE_BYTES]) {
  if (!h || !handle_out) return fail(MT_EINVAL, "mt_halo_handle: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == MT_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t ipc;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaIpcGetMemHandle(&ipc, h->box));
  memcpy(handle_out, &ipc, sizeof(ipc));
  return MT_OK;
}

int mt_halo_connect(mt_halo* h, const unsigned char* left_handle,
                    const unsigned char* right_handle) {
  if (!h) return fail(MT_EINVAL, "mt_halo_connect: NULL argument");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaIpcMemHandle_t ipc;
  if (left_handle) {
    memcpy(&ipc, left_handle, sizeof(ipc));
    CUDA_TRY(cudaIpcOpenMemHandle(&h->left, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->left_ipc = true;
  }
  if (right_handle) {
    memcpy(&ipc, right_handle, sizeof(ipc));
    CUDA_TRY(cudaIpcOpenMemHandle(&h->right, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->right_ipc = true;
  }
  return MT_OK;
}

int mt_halo_connect_local(mt_halo* h, mt_halo* left, mt_halo* right) {
  if (!h) return fail(MT_EINVAL, "mt_halo_connect_local: NULL argument");
  if ((left && (left->halo != h->halo || left->dtype != h->dtype)) ||
      (right && (right->halo != h->halo || right->dtype != h->dtype)))
    return fail(MT_EINVAL, "mt_halo_connect_local: neighbours differ in halo width or dtype");
  h->left = left ? left->box : nullptr;
  h->right = right ? right->box : nullptr;
  return MT_OK;
}

extern "C++" {
namespace {
template <typename T>
HaloArgs<T> halo_args(mt_halo* h, void* local, int64_t n_local, unsigned epoch) {
  HaloArgs<T> a;
  a.local = (T*)local;
  a.n_local = n_local;
  a.halo = h->halo;
  a.has_left = h->left != nullptr;
  a.has_right = h->right != nullptr;
  a.mine = h->box;
  a.left = h->left;
  a.right = h->right;
  a.epoch = epoch;
  a.timeout_ns = 10ull * 1000ull * 1000ull * 1000ull;
  return a;
}
int halo_check(mt_halo* h, void* local, int64_t n_local, const char* who) {
  if (!h || !local) return fail(MT_EINVAL, "%s: NULL argument", who);
  const int sides = (h->left != nullptr) + (h->right != nullptr);
  if (n_local < (int64_t)h->halo * (sides + 1))
    return fail(MT_EINVAL, "%s: block of %lld cells is too small for halo %d", who,
                (long long)n_local, h->halo);
  return MT_OK;
}
}  // namespace
}  // extern "C++"

int mt_halo_push(mt_halo* h, void* local, int64_t n_local, void* stream) {
  const int rc = halo_check(h, local, n_local, "mt_halo_push");
  if (rc != MT_OK) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  const unsigned epoch = ++h->pushes;
  if (h->dtype == MT_F64)
    halo_push_kernel<double><<<1, 256, 0, (cudaStream_t)stream>>>(halo_args<double>(h, local, n_local, epoch));
  else
    halo_push_kernel<float><<<1, 256, 0, (cudaStream_t)stream>>>(halo_args<float>(h, local, n_local, epoch));
  CUDA_TRY(cudaGetLastError());
  return MT_OK;
}

int mt_halo_pull(mt_halo* h, void* local, int64_t n_local, void* stream) {
  const int rc = halo_check(h, local, n_local, "mt_halo_pull");
  if (rc != MT_OK) return rc;
  if (h->pulls >= h->pushes) return fail(MT_ESTATE, "mt_halo_pull: no push to pair with");
  CUDA_TRY(cudaSetDevice(h->device));
  const unsigned epoch = ++h->pulls;
  if (h->dtype == MT_F64)
    halo_pull_kernel<double><<<1, 256, 0, (cudaStream_t)stream>>>(halo_args<double>(h, local, n_local, epoch));
  else
    halo_pull_kernel<float><<<1, 256, 0, (cudaStream_t)stream>>>(halo_args<float>(h, local, n_local, epoch));
  CUDA_TRY(cudaGetLastError());
  return MT_OK;
}

int mt_halo_advance(mt_engine* e, mt_halo* h, void* obs_a, void* obs_b, int32_t n_steps,
                    int32_t ghosts_fresh, void* stream, int32_t* result_in_b) {
  if (!e || !h || !obs_a || !obs_b) return fail(MT_EINVAL, "mt_halo_advance: NULL argument");
  if (!e->have_params) return fail(MT_ESTATE, "mt_halo_advance: call mt_set_params first");
  if (obs_a == obs_b) return fail(MT_EINVAL, "mt_halo_advance: obs_a and obs_b must differ");
  if (n_steps < 0) return fail(MT_EINVAL, "mt_halo_advance: n_steps must be >= 0");
  if (e->cfg.n_envs != 1) return fail(MT_EINVAL, "mt_halo_advance: the engine must hold one road block");
  if ((e->cfg.dtype == MT_F64) != (h->dtype == MT_F64) || e->cfg.device != h->device)
    return fail(MT_EINVAL, "mt_halo_advance: engine and mailbox differ in dtype or device");
  const int rc0 = halo_check(h, obs_a, e->cfg.n_cells, "mt_halo_advance");
  if (rc0 != MT_OK) return rc0;
  if (!can_segfuse(e, obs_a, obs_b))
    return fail(MT_EINVAL, "mt_halo_advance: the block does not take the window kernel "
                           "(too short, LOOP ends, or the direct-kernel flag)");
  if ((h->left && e->params.bc_left != MT_BC_HALO) || (h->right && e->params.bc_right != MT_BC_HALO))
    return fail(MT_EINVAL, "mt_halo_advance: an edge with a neighbour must have the MT_BC_HALO rule");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (e->params_pending) {
    if (s != e->params_stream) CUDA_TRY(cudaStreamWaitEvent(s, e->params_ready, 0));
    e->params_pending = false;
  }
  void* cur = obs_a;
  void* nxt = obs_b;
  bool fresh = ghosts_fresh != 0;
  for (int done = 0; done < n_steps;) {
    const int k = n_steps - done < h->halo ? n_steps - done : h->halo;
    SegHalo hx;
    hx.mine = h->box; hx.left = h->left; hx.right = h->right; hx.halo = h->halo;
    hx.timeout_ns = 10ull * 1000ull * 1000ull * 1000ull;
    if (!fresh) {
      if (h->pulls >= h->pushes) return fail(MT_ESTATE, "mt_halo_advance: stale ghost cells and no push to pair with");
      hx.pull_epoch = ++h->pulls;
    }
    hx.push_epoch = ++h->pushes;
    const int rc = segfused_any(e, cur, nullptr, 0, nxt, k, &hx, s);
    if (rc != MT_OK) return rc;
    void* tmp = cur; cur = nxt; nxt = tmp;
    done += k;
    fresh = false;
  }
  if (result_in_b) *result_in_b = (cur == obs_b) ? 1 : 0;
  return MT_OK;
}

int mt_halo_status(mt_halo* h, int32_t* timed_out) {
  if (!h || !timed_out) return fail(MT_EINVAL, "mt_halo_status: NULL argument");
  CUDA_TRY(cudaSetDevice(h->device));
  HaloBox box;
  CUDA_TRY(cudaMemcpy(&box, h->box, sizeof(box), cudaMemcpyDeviceToHost));
  *timed_out = box.error ? 1 : 0;
  return MT_OK;
}

int mt_alloc_pinned(uint64_t bytes, void** out) {
  if (!out) return fail(MT_EINVAL, "mt_alloc_pinned: NULL argument");
  CUDA_TRY(cudaMallocHost(out, (size_t)bytes));
  return MT_OK;
}

int mt_free_pinned(void* p) {
  CUDA_TRY(cudaFreeHost(p));
  return MT_OK;
}

}  // extern "C"
