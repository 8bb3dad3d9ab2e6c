// mt_halo.cuh -- ghost-cell exchange of a domain-decomposed road over peer
// memory (NVLink / NVSwitch), without NCCL on the data path.
//
// Each rank owns a MAILBOX in its own HBM, mapped into its two neighbours'
// address spaces (CUDA IPC).  An exchange is two tiny kernels on the step
// stream:
//   push: store my first / last `halo` own cells ([rho(h) | v(h)]) straight
//         into the neighbours' mailboxes (peer stores), fence, then publish
//         the exchange number in the neighbours' flag words (st.release.sys);
//   pull: wait until both my flag words carry this exchange's number
//         (ld.acquire.sys), then copy my mailbox into my ghost cells.
// A push never waits, so two ranks cannot block each other.  Mailbox slots are
// double-buffered by the parity of the exchange number: a rank can only be one
// exchange ahead of its neighbour (its pull of exchange e needs the
// neighbour's push of e, which the neighbour issues after its own pull of
// e-1), so slot e&1 is free again when exchange e+2 writes it.
#pragma once
#include <cstdint>

namespace mt {

enum { HALO_FROM_LEFT = 0, HALO_FROM_RIGHT = 1 };

// [flags: from_left, from_right, error, pad] then data[2 parities][2 sides][2*halo]
struct HaloBox {
  unsigned int flag[2];
  unsigned int error;   // set by a pull that timed out
  unsigned int pad;
};

template <typename T>
__device__ __forceinline__ T* halo_slot(void* box, int halo, unsigned parity, int side) {
  T* data = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(box) + sizeof(HaloBox));
  return data + ((size_t)(parity * 2 + side)) * (2 * (size_t)halo);
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T> struct HaloArgs {
  T* local;          // this rank's block [rho(n_local) | v(n_local)], ghosts included
  long long n_local;
  int halo;
  int has_left, has_right;
  void* mine;        // my mailbox
  void* left;        // left neighbour's mailbox (peer mapping) or nullptr
  void* right;
  unsigned int epoch;          // number of this exchange (1, 2, ...)
  unsigned long long timeout_ns;
};

// one CTA; a rank's edge cells are a few KB
template <typename T>
__global__ void halo_push_kernel(const HaloArgs<T> a) {
  const int h = a.halo;
  const long long nl = a.n_local;
  const long long first = a.has_left ? h : 0;             // first own cell
  const long long end = nl - (a.has_right ? h : 0);       // one past the last own cell
  const unsigned par = a.epoch & 1u;
  if (a.has_left) {   // I am my left neighbour's RIGHT neighbour
    T* dst = halo_slot<T>(a.left, h, par, HALO_FROM_RIGHT);
    for (int i = threadIdx.x; i < h; i += blockDim.x) {
      dst[i] = a.local[first + i];
      dst[h + i] = a.local[nl + first + i];
    }
  }
  if (a.has_right) {
    T* dst = halo_slot<T>(a.right, h, par, HALO_FROM_LEFT);
    for (int i = threadIdx.x; i < h; i += blockDim.x) {
      dst[i] = a.local[end - h + i];
      dst[h + i] = a.local[nl + end - h + i];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (a.has_left) st_release_sys(&reinterpret_cast<HaloBox*>(a.left)->flag[HALO_FROM_RIGHT], a.epoch);
    if (a.has_right) st_release_sys(&reinterpret_cast<HaloBox*>(a.right)->flag[HALO_FROM_LEFT], a.epoch);
  }
}

template <typename T>
__global__ void halo_pull_kernel(const HaloArgs<T> a) {
  __shared__ int ok;
  HaloBox* box = reinterpret_cast<HaloBox*>(a.mine);
  if (threadIdx.x == 0) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int good = 1;
    for (int side = 0; side < 2 && good; ++side) {
      if (!(side == HALO_FROM_LEFT ? a.has_left : a.has_right)) continue;
      // exchange numbers only grow: ">= epoch" also covers a neighbour that is ahead
      while ((int)(ld_acquire_sys(&box->flag[side]) - a.epoch) < 0) {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > a.timeout_ns) { good = 0; break; }
      }
    }
    if (!good) box->error = 1u;
    ok = good;
  }
  __syncthreads();
  if (!ok) return;   // leave the ghosts alone; mt_halo_status reports the timeout
  const int h = a.halo;
  const long long nl = a.n_local;
  const long long end = nl - (a.has_right ? h : 0);
  const unsigned par = a.epoch & 1u;
  if (a.has_left) {
    const T* src = halo_slot<T>(a.mine, h, par, HALO_FROM_LEFT);
    for (int i = threadIdx.x; i < h; i += blockDim.x) {
      a.local[i] = src[i];
      a.local[nl + i] = src[h + i];
    }
  }
  if (a.has_right) {
    const T* src = halo_slot<T>(a.mine, h, par, HALO_FROM_RIGHT);
    for (int i = threadIdx.x; i < h; i += blockDim.x) {
      a.local[end + i] = src[i];
      a.local[nl + end + i] = src[h + i];
    }
  }
}

}  // namespace mt
