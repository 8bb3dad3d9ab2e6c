// mt_abort.cuh -- MT_FLAG_FSOLVE_ABORT: the reference's behaviour on roads
// whose relaxation system cannot be solved.
//
// ARZModel.arz_update_points (arz.py:419-467) hands
//     rhs = y + step (Pm - P) + (dt / tau) rho' V(rho')            (arz.py:460-461)
// to scipy.optimize.fsolve with the initial guess x0 = y (arz.py:462-463).
// With an exact-zero density P = Q (y / rho) is NaN (arz.py:411), every
// residual MINPACK sees is NaN, every trial step is rejected and fsolve
// returns x0: y is FROZEN for the whole road that step (the system is solved
// jointly), while the density still advances.  The step kernels evaluate the
// closed-form root of the relaxation instead, which is
// NaN at the affected cells only.  With the flag set this kernel runs after
// the step: one warp per road re-derives rhs for all N cells, and where one of
// them is not finite rewrites the road's new speeds from the old relative
// flow, v'_i = y_i / rho'_i + V(rho'_i), for the interior cells and the
// boundary cells that copy them; densities are untouched (they never
// depended on P).  Pinned by tests/golden/reference_zero_density.npz
// (outputs of the unmodified reference).
#pragma once
#include <stdint.h>
#include <cuda_bf16.h>
#include "mtraffic.h"
#include "mt_math.cuh"
#include "mt_cell.cuh"

namespace mt {

template <typename T> struct AbortArgs {
  const T* in;        // (B, 2N) state before the step
  T* out;             // (B, 2N) state after the step (speeds of aborted roads rewritten)
  const T* table;
  const int* flags;
  const T* action;    // per-road speed limit of this step, or nullptr
  T* reward;          // rewards of this step to refresh for aborted roads, or nullptr
  uint16_t* obs16;    // bfloat16 copy of `out` to refresh, or nullptr
  int* aborted;       // per-road 0/1 record (device), or nullptr
  long long B;
  int N;
  BcArgs<T> bc;
  T step;
  T dt;               // rhs needs dt / tau itself (the table only holds 1 + dt / tau)
  T tau_scalar;       // used when tau_env is null
  const T* tau_env;
};

template <typename T> __device__ __forceinline__ bool mt_finite(T x) { return x - x == T(0); }

template <typename T>
__global__ void arz_abort_fixup_kernel(const AbortArgs<T> a) {
  using A = Ar<T, false>;
  constexpr int LAMK = LAM_RUNTIME;
  const long long env = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= a.B) return;
  const int lane = threadIdx.x & 31;
  const int N = a.N;
  EnvC<T> e;
  load_env(e, a.table, a.flags, env, a.action, a.step);
  e.qc = A::mul(e.rho_crit, A::mul(e.v_max, e.g_crit));
  const T* rho = a.in + env * (2LL * N);
  const T* v = rho + N;
  T* orho = a.out + env * (2LL * N);
  T* ov = orho + N;
  // dt / tau as the reference forms it (a Python float division for scalars)
  const T dt_tau = A::div(a.dt, a.tau_env ? a.tau_env[env] : a.tau_scalar);

  auto cell = [&](int i, T& y, T& r, T& D, T& S) {   // flux ingredients of cell i
    arz_cell<T, false, LAMK>(e, rho[i], v[i], y, r, D, S);
  };
  bool bad = false;
  for (int i = lane; i < N; i += 32) {
    T y, r, D, S, yl, rl, Dl, Sl, yr, rr, Dr, Sr;
    cell(i, y, r, D, S);
    const int il = i > 0 ? i - 1 : 0, ir = i < N - 1 ? i + 1 : N - 1;
    cell(il, yl, rl, Dl, Sl);
    cell(ir, yr, rr, Dr, Sr);
    const T Q = npmin(D, i < N - 1 ? Sr : S);                 // arz.py:405-408
    const T P = A::mul(Q, r);                                 // arz.py:411
    T Qm, Pm;                                                 // arz.py:414-415
    if (i == 0) { Qm = Q; Pm = P; }
    else { Qm = npmin(Dl, S); Pm = A::mul(Qm, rl); }
    const T rho_n = A::add(rho[i], A::mul(e.step, A::sub(Qm, Q)));            // arz.py:456
    const T rhs = A::add(A::add(y, A::mul(e.step, A::sub(Pm, P))),
                         A::mul(dt_tau, A::mul(rho_n, v_eq<T, false, LAMK>(rho_n, e))));
    bad |= !mt_finite(rhs);
  }
  bad = __any_sync(0xffffffffu, bad);
  if (a.aborted != nullptr && lane == 0) a.aborted[env] = bad ? 1 : 0;
  if (!bad) return;

  // y frozen: interior speeds from the OLD relative flow and the NEW density
  // (already zero-patched by the step kernel, arz.py:274-275)
  for (int i = 1 + lane; i < N - 1; i += 32) {
    const T y = arz_y<T, false, LAMK>(e, rho[i], v[i]);
    T rn = orho[i], vn;
    arz_u<T, false, LAMK>(e, rn, y, vn);
    ov[i] = vn;
  }
  __syncwarp();
  // boundary cells that copy the interior (arz.py:332-344); LOOP copies the old
  // state and CONSTANT the given values: both are what the step kernel wrote
  if (lane == 0) {
    if (a.bc.bc_l == MT_BC_EXTEND) ov[0] = ov[1];
    if (a.bc.bc_r == MT_BC_EXTEND) ov[N - 1] = ov[N - 2];
  }
  __syncwarp();
  if (a.reward != nullptr) {   // sum(rho v) / sum(rho) of the rewritten road
    T num = 0, den = 0;
    for (int i = lane; i < N; i += 32) { num += orho[i] * ov[i]; den += orho[i]; }
    for (int off = 16; off > 0; off >>= 1) {
      num += __shfl_xor_sync(0xffffffffu, num, off);
      den += __shfl_xor_sync(0xffffffffu, den, off);
    }
    if (lane == 0) a.reward[env] = num / den;
  }
  if (a.obs16 != nullptr) {
    uint16_t* o16 = a.obs16 + env * (2LL * N) + N;
    for (int i = lane; i < N; i += 32) {
      const __nv_bfloat16 h = __float2bfloat16_rn((float)ov[i]);
      o16[i] = *reinterpret_cast<const uint16_t*>(&h);
    }
  }
}

}  // namespace mt
