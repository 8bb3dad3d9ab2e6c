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
#ifdef MT_DIV_SHORTCUT
  // (round 2 candidate, off by default: y / rho from t = v - V(rho) and one
  // residual instead of a division -- DESIGN.md section 9, item 1)
  if constexpr (sizeof(T) == 8 && !FAST) {
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      T t;
      arz_cell_nodiv_t<T, FAST, LAMK>(e, rho[k], v[k], t, y[k], D[k], S[k]);
      r[k] = quot_of_product(t, rho[k], y[k], ok);
    }
    if (!ok) {  // rare: a power-of-two t, a zero density, operands out of range
#pragma unroll
      for (int k = 0; k < CPT; ++k) r[k] = A::div(y[k], rho[k]);
    }
    return;
  }
#endif
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

// ---------------------------------------------------------------------------
// Split form of arz_pass2 for the pipelined kernels: arz_mid() needs nothing
// from the neighbouring threads and runs BEFORE the exchange barrier (fewer
// values live across it, less work after it); arz_edges() finishes the first
// and the last cell of the thread once the neighbours' values are visible.
// Same operations in the same per-cell order as arz_pass2: results are equal.
// ---------------------------------------------------------------------------
template <typename T, bool FAST, int LAMK, int M>
__device__ __forceinline__ void arz_u_batch(const EnvC<T>& e, T (&rn)[M], const T (&yn)[M],
                                            T (&vn)[M]) {
  using A = Ar<T, FAST>;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < M; ++k) {
    vn[k] = A::div_unchecked(yn[k], rn[k]);
    ok &= A::div_quick(yn[k], rn[k]);
  }
  if (!ok) {  // rare: re-test admitting exact zeros
    ok = true;
#pragma unroll
    for (int k = 0; k < M; ++k) ok &= A::div_safe(yn[k], rn[k]);
  }
  if (ok) {
#pragma unroll
    for (int k = 0; k < M; ++k) vn[k] = A::add(vn[k], v_eq<T, FAST, LAMK>(rn[k], e));
  } else {
#pragma unroll
    for (int k = 0; k < M; ++k) arz_u<T, FAST, LAMK>(e, rn[k], yn[k], vn[k]);
  }
}

// Q/P: fluxes of my cells 0..CPT-2 (kept for arz_edges); rn/vn[1..CPT-2] final.
template <typename T, bool FAST, int LAMK, int CPT>
__device__ __forceinline__ void arz_mid(const EnvC<T>& e, const T (&rho)[CPT], const T (&y)[CPT],
                                        const T (&r)[CPT], const T (&D)[CPT], const T (&S)[CPT],
                                        T (&Q)[CPT], T (&P)[CPT], T (&rn)[CPT], T (&vn)[CPT]) {
  using A = Ar<T, FAST>;
#pragma unroll
  for (int k = 0; k < CPT - 1; ++k) {
    Q[k] = npmin(D[k], S[k + 1]);  // arz.py:408
    P[k] = A::mul(Q[k], r[k]);     // arz.py:411
  }
  if constexpr (CPT > 2) {
    constexpr int M = CPT - 2;
    T rm[M], ym[M], vm[M];
#pragma unroll
    for (int k = 0; k < M; ++k)
      arz_advance<T, FAST>(e, rho[k + 1], y[k + 1], Q[k], Q[k + 1], P[k], P[k + 1], rm[k], ym[k]);
    arz_u_batch<T, FAST, LAMK, M>(e, rm, ym, vm);
#pragma unroll
    for (int k = 0; k < M; ++k) { rn[k + 1] = rm[k]; vn[k + 1] = vm[k]; }
  }
}

template <typename T, bool FAST, int LAMK, int CPT>
__device__ __forceinline__ void arz_edges(const EnvC<T>& e, const BcArgs<T>& b, bool row_first,
                                          bool row_last, const T (&rho)[CPT], const T (&v)[CPT],
                                          const T (&y)[CPT], const T (&r)[CPT],
                                          const T (&D)[CPT], const T (&S)[CPT], T (&Q)[CPT],
                                          T (&P)[CPT], T S_right, T D_left, T r_left, T loop_rho,
                                          T loop_v, T (&rn)[CPT], T (&vn)[CPT]) {
  using A = Ar<T, FAST>;
  Q[CPT - 1] = npmin(D[CPT - 1], S_right);
  P[CPT - 1] = A::mul(Q[CPT - 1], r[CPT - 1]);
  const T Qm = npmin(D_left, S[0]);   // arz.py:414
  const T Pm = A::mul(Qm, r_left);    // arz.py:415
  T re[2], ye[2], ve[2];
  arz_advance<T, FAST>(e, rho[0], y[0], Qm, Q[0], Pm, P[0], re[0], ye[0]);
  arz_advance<T, FAST>(e, rho[CPT - 1], y[CPT - 1], Q[CPT - 2], Q[CPT - 1], P[CPT - 2],
                       P[CPT - 1], re[1], ye[1]);
  arz_u_batch<T, FAST, LAMK, 2>(e, re, ye, ve);
  rn[0] = re[0]; vn[0] = ve[0];
  rn[CPT - 1] = re[1]; vn[CPT - 1] = ve[1];
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

// Split form of lwr_pass2 (see arz_mid / arz_edges): f = fluxes of my cells
// 0..CPT-2, rn/vn[1..CPT-2] final after lwr_mid().
template <typename T, bool FAST, int LAMK, int CPT>
__device__ __forceinline__ void lwr_mid(const EnvC<T>& e, const T (&rho)[CPT], const T (&D)[CPT],
                                        const T (&S)[CPT], T (&f)[CPT], T (&rn)[CPT],
                                        T (&vn)[CPT]) {
#pragma unroll
  for (int k = 0; k < CPT - 1; ++k) f[k] = npmin(D[k], S[k + 1]);  // lwr.py:297
#pragma unroll
  for (int k = 1; k < CPT - 1; ++k) {
    rn[k] = lwr_advance<T, FAST>(e, rho[k], f[k], f[k - 1]);       // lwr.py:234
    vn[k] = v_eq<T, FAST, LAMK>(rn[k], e);                         // lwr.py:201
  }
}

template <typename T, bool FAST, int LAMK, int CPT>
__device__ __forceinline__ void lwr_edges(const EnvC<T>& e, const BcArgs<T>& b, bool row_first,
                                          bool row_last, const T (&rho)[CPT], const T (&D)[CPT],
                                          const T (&S)[CPT], T (&f)[CPT], T S_right, T D_left,
                                          T loop_a, T loop_b, T (&rn)[CPT], T (&vn)[CPT]) {
  f[CPT - 1] = npmin(D[CPT - 1], S_right);                         // lwr.py:294
  const T fm = npmin(D_left, S[0]);
  rn[0] = lwr_advance<T, FAST>(e, rho[0], f[0], fm);
  rn[CPT - 1] = lwr_advance<T, FAST>(e, rho[CPT - 1], f[CPT - 1], f[CPT - 2]);
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
  vn[0] = v_eq<T, FAST, LAMK>(rn[0], e);
  vn[CPT - 1] = v_eq<T, FAST, LAMK>(rn[CPT - 1], e);
}

// ---------------------------------------------------------------------------
// What the tiled kernels call around their exchange barrier.  cell_mid() runs
// before it, cell_edges() after it; with fewer than three cells per thread
// there is no interior and cell_edges() is the whole of pass2.
// ---------------------------------------------------------------------------
template <typename T, bool ARZ, bool FAST, int LAMK, int CPT>
__device__ __forceinline__ void cell_mid(const EnvC<T>& e, const T (&rho)[CPT], const T (&y)[CPT],
                                         const T (&r)[CPT], const T (&D)[CPT], const T (&S)[CPT],
                                         T (&Q)[CPT], T (&P)[CPT], T (&rn)[CPT], T (&vn)[CPT]) {
  if constexpr (CPT > 2) {
    if constexpr (ARZ) arz_mid<T, FAST, LAMK, CPT>(e, rho, y, r, D, S, Q, P, rn, vn);
    else lwr_mid<T, FAST, LAMK, CPT>(e, rho, D, S, Q, rn, vn);
  }
}

template <typename T, bool ARZ, bool FAST, int LAMK, int CPT>
__device__ __forceinline__ void cell_edges(const EnvC<T>& e, const BcArgs<T>& b, bool row_first,
                                           bool row_last, const T (&rho)[CPT], const T (&v)[CPT],
                                           const T (&y)[CPT], const T (&r)[CPT],
                                           const T (&D)[CPT], const T (&S)[CPT], T (&Q)[CPT],
                                           T (&P)[CPT], T S_right, T D_left, T r_left, T loop_a,
                                           T loop_b, T (&rn)[CPT], T (&vn)[CPT]) {
  if constexpr (CPT > 2) {
    if constexpr (ARZ)
      arz_edges<T, FAST, LAMK, CPT>(e, b, row_first, row_last, rho, v, y, r, D, S, Q, P, S_right,
                                    D_left, r_left, loop_a, loop_b, rn, vn);
    else
      lwr_edges<T, FAST, LAMK, CPT>(e, b, row_first, row_last, rho, D, S, Q, S_right, D_left,
                                    loop_a, loop_b, rn, vn);
  } else {
    if constexpr (ARZ)
      arz_pass2<T, FAST, LAMK, CPT>(e, b, row_first, row_last, rho, v, y, r, D, S, S_right,
                                    D_left, r_left, loop_a, loop_b, rn, vn);
    else
      lwr_pass2<T, FAST, LAMK, CPT>(e, b, row_first, row_last, rho, D, S, S_right, D_left,
                                    loop_a, loop_b, rn, vn);
  }
}

}  // namespace mt
