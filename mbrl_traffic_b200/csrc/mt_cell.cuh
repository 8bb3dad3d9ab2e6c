// mt_cell.cuh -- the per-thread finite-volume update, shared by the direct and
// the TMA-pipelined kernels.
//
// A thread owns CPT consecutive cells of one road.  pass1 derives the per-cell
// flux ingredients; the kernel then swaps three edge values with the two
// neighbouring threads (supply of my first cell to the left neighbour, demand
// and y/rho of my last cell to the right neighbour); pass2 forms the fluxes,
// advances the cells, applies the boundary rule on the road's first/last
// thread and converts back to (rho, v).
#pragma once
#include "mtraffic.h"
#include "mt_math.cuh"

namespace mt {

template <typename T> struct BcArgs {
  int bc_l, bc_r;
  T bl0, bl1, br0, br1;  // CONSTANT boundary values
};

// ===========================================================================
// ARZ  (arz.py:205-467)
// ===========================================================================
template <typename T, bool FAST, int LAMK, int CPT>
__device__ __forceinline__ void arz_pass1(const EnvC<T>& e, const T (&rho)[CPT], const T (&v)[CPT],
                                          T (&y)[CPT], T (&r)[CPT], T (&D)[CPT], T (&S)[CPT]) {
  using A = Ar<T, FAST>;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    arz_cell_nodiv<T, FAST, LAMK>(e, rho[k], v[k], y[k], D[k], S[k]);
    r[k] = A::div_unchecked(y[k], rho[k]);  // y / rho (arz.py:411)
    ok &= A::div_quick(y[k], rho[k]);
  }
  if (!ok) {  // rare: an exact zero, or operands outside the inline expansion's range
    bool ok2 = true;
#pragma unroll
    for (int k = 0; k < CPT; ++k) ok2 &= A::div_safe(y[k], rho[k]);
    if (!ok2) {
#pragma unroll
      for (int k = 0; k < CPT; ++k) r[k] = A::div(y[k], rho[k]);
    }
  }
}

// S_right: supply of the cell right of my last cell (my own last supply when
// I own the road's last cell, arz.py:405).  D_left / r_left: demand and y/rho
// of the cell left of my first cell (ignored on the road's first thread: all
// boundary rules overwrite cell 0, arz.py:316-360).  loop_rho / loop_v: OLD
// state of the road's last cell, needed by the LOOP rule on the first thread.
template <typename T, bool FAST, int LAMK, int CPT>
__device__ __forceinline__ void arz_pass2(const EnvC<T>& e, const BcArgs<T>& b, bool row_first,
                                          bool row_last, const T (&rho)[CPT], const T (&v)[CPT],
                                          const T (&y)[CPT], const T (&r)[CPT],
                                          const T (&D)[CPT], const T (&S)[CPT], T S_right,
                                          T D_left, T r_left, T loop_rho, T loop_v,
                                          T (&rn)[CPT], T (&vn)[CPT]) {
  using A = Ar<T, FAST>;
  T Q[CPT], P[CPT], yn[CPT];
#pragma unroll
  for (int k = 0; k < CPT - 1; ++k) Q[k] = npmin(D[k], S[k + 1]);  // arz.py:408
  Q[CPT - 1] = npmin(D[CPT - 1], S_right);
#pragma unroll
  for (int k = 0; k < CPT; ++k) P[k] = A::mul(Q[k], r[k]);         // arz.py:411
  T Qm = npmin(D_left, S[0]);                                      // arz.py:414
  T Pm = A::mul(Qm, r_left);                                       // arz.py:415
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    arz_advance<T, FAST>(e, rho[k], y[k], Qm, Q[k], Pm, P[k], rn[k], yn[k]);
    Qm = Q[k];
    Pm = P[k];
  }
  // u(): y/rho + V(rho) (arz.py:258-277).  The zero-density patch lives in the
  // fallback: div_safe() already rejects rho == 0.
  bool ok = true;
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    vn[k] = A::div_unchecked(yn[k], rn[k]);
    ok &= A::div_quick(yn[k], rn[k]);
  }
  if (!ok) {  // rare: re-test admitting exact zeros
    ok = true;
#pragma unroll
    for (int k = 0; k < CPT; ++k) ok &= A::div_safe(yn[k], rn[k]);
  }
  if (ok) {
#pragma unroll
    for (int k = 0; k < CPT; ++k) vn[k] = A::add(vn[k], v_eq<T, FAST, LAMK>(rn[k], e));
  } else {
#pragma unroll
    for (int k = 0; k < CPT; ++k) arz_u<T, FAST, LAMK>(e, rn[k], yn[k], vn[k]);
  }
  // boundary cells (arz.py:316-360); only two threads per road come here
  if (row_first | row_last) {
    if (row_first) {
      if (b.bc_l == MT_BC_LOOP) {            // OLD state of the last cell
        rn[0] = loop_rho;
        arz_u<T, FAST, LAMK>(e, rn[0], arz_y<T, FAST, LAMK>(e, loop_rho, loop_v), vn[0]);
      } else if (b.bc_l == MT_BC_EXTEND) {   // NEW state of cell 1
        rn[0] = rn[1];
        vn[0] = vn[1];
      } else if (b.bc_l == MT_BC_CONSTANT) {
        rn[0] = b.bl0;
        arz_u<T, FAST, LAMK>(e, rn[0], b.bl1, vn[0]);
      } else {                               // HALO: ghost cell, passed through
        rn[0] = rho[0];
        vn[0] = v[0];
      }
    }
    if (row_last) {
      if (b.bc_r == MT_BC_LOOP) {            // OLD state of cell N-2
        rn[CPT - 1] = rho[CPT - 2];
        arz_u<T, FAST, LAMK>(e, rn[CPT - 1], y[CPT - 2], vn[CPT - 1]);
      } else if (b.bc_r == MT_BC_EXTEND) {   // NEW state of cell N-2
        rn[CPT - 1] = rn[CPT - 2];
        vn[CPT - 1] = vn[CPT - 2];
      } else if (b.bc_r == MT_BC_CONSTANT) {
        rn[CPT - 1] = b.br0;
        arz_u<T, FAST, LAMK>(e, rn[CPT - 1], b.br1, vn[CPT - 1]);
      } else {
        rn[CPT - 1] = rho[CPT - 1];
        vn[CPT - 1] = v[CPT - 1];
      }
    }
  }
}

// ===========================================================================
// LWR  (lwr.py:192-316)
// ===========================================================================
template <typename T, bool FAST, int LAMK, int CPT>
__device__ __forceinline__ void lwr_pass1(const EnvC<T>& e, const T (&rho)[CPT], T (&D)[CPT],
                                          T (&S)[CPT]) {
#pragma unroll
  for (int k = 0; k < CPT; ++k) lwr_cell<T, FAST, LAMK>(e, rho[k], D[k], S[k]);
}

// loop_a / loop_b: OLD densities of cells N-2 and N-1 (LOOP rule, first thread).
template <typename T, bool FAST, int LAMK, int CPT>
__device__ __forceinline__ void lwr_pass2(const EnvC<T>& e, const BcArgs<T>& b, bool row_first,
                                          bool row_last, const T (&rho)[CPT], const T (&D)[CPT],
                                          const T (&S)[CPT], T S_right, T D_left, T loop_a,
                                          T loop_b, T (&rn)[CPT], T (&vn)[CPT]) {
  T f[CPT];
#pragma unroll
  for (int k = 0; k < CPT - 1; ++k) f[k] = npmin(D[k], S[k + 1]);  // lwr.py:297
  f[CPT - 1] = npmin(D[CPT - 1], S_right);                         // lwr.py:294
  T fm = npmin(D_left, S[0]);
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    rn[k] = lwr_advance<T, FAST>(e, rho[k], f[k], fm);             // lwr.py:234
    fm = f[k];
  }
  // boundary cells (lwr.py:237-260)
  if (row_first | row_last) {
    const T second_last = rn[CPT - 2];
    if (row_first) {
      if (b.bc_l == MT_BC_LOOP) {            // UPDATED value of the last cell
        T D2, S2, D1, S1;
        lwr_cell<T, FAST, LAMK>(e, loop_a, D2, S2);
        lwr_cell<T, FAST, LAMK>(e, loop_b, D1, S1);
        rn[0] = lwr_advance<T, FAST>(e, loop_b, npmin(D1, S1), npmin(D2, S1));
      } else if (b.bc_l == MT_BC_EXTEND) {   // fm = f_0 (lwr.py:231): cell 0 frozen
        rn[0] = lwr_advance<T, FAST>(e, rho[0], f[0], f[0]);
      } else if (b.bc_l == MT_BC_CONSTANT) {
        rn[0] = b.bl0;
      } else {
        rn[0] = rho[0];
      }
    }
    if (row_last) {
      if (b.bc_r == MT_BC_LOOP) rn[CPT - 1] = second_last;  // UPDATED cell N-2
      else if (b.bc_r == MT_BC_CONSTANT) rn[CPT - 1] = b.br0;
      else if (b.bc_r == MT_BC_HALO) rn[CPT - 1] = rho[CPT - 1];
    }
  }
#pragma unroll
  for (int k = 0; k < CPT; ++k) vn[k] = v_eq<T, FAST, LAMK>(rn[k], e);  // lwr.py:201
}

}  // namespace mt
