// FP64 issue rate of one SM, measured: independent DFMA chains, clock64() around them.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/fp64_rate tools/microbench/fp64_rate.cu
//   ./tools/microbench/fp64_rate
//
// Prints lanes (thread-level FMAs) per clock and SM for DFMA, DADD+DMUL and FFMA at
// several occupancies.  DESIGN.md section 5b uses the DFMA figure as the second
// roofline of the float64 kernels (their FP64 instruction counts come from ncu).
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int MODE>
__global__ void chains(T* out, long long* cycles, int iters, T a0, T b0) {
  // per-thread operands: all three sources in vector registers, as in the step kernels
  const T a = a0 + (T)threadIdx.x * (T)1e-12, b = b0 + (T)threadIdx.x * (T)1e-15;
  T acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = (T)(threadIdx.x + j);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) acc[j] = fma(acc[j], a, b);                    // one FMA
      else acc[j] = __dadd_rn(__dmul_rn((double)acc[j], (double)a), (double)b);   // MUL + ADD
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  T s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// the look-ahead loop of mt_nonlocal.cuh without its shared-memory reads: five sums per
// thread, one weight per term, the window sliding by one register per term -- three
// different vector registers per DFMA
__global__ void window(double* out, long long* cycles, int iters, double a0, double b0) {
  double acc[5], win[5], g[4];
#pragma unroll
  for (int j = 0; j < 5; ++j) { acc[j] = 0.0; win[j] = a0 + (threadIdx.x + j) * 1e-9; }
#pragma unroll
  for (int j = 0; j < 4; ++j) g[j] = b0 + (threadIdx.x + j) * 1e-12;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double nxt = win[0];
#pragma unroll
      for (int m = 0; m < 5; ++m) acc[m] = fma(g[u], win[m], acc[m]);
#pragma unroll
      for (int m = 0; m < 4; ++m) win[m] = win[m + 1];
      win[4] = nxt;
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  double s = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

void run_window(int sms, int blocks_per_sm, int threads) {
  const int iters = 4096;
  const int grid = sms * blocks_per_sm;
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * grid * threads);
  cudaMalloc(&cyc, sizeof(long long) * grid);
  window<<<grid, threads>>>(out, cyc, 16, 1.0, 1e-3);
  window<<<grid, threads>>>(out, cyc, iters, 1.0, 1e-3);
  cudaDeviceSynchronize();
  long long* h = new long long[grid];
  cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += (double)h[i];
  mean /= grid;
  const double ops_per_sm = (double)blocks_per_sm * threads * 5.0 * iters;
  printf("{\"op\": \"window loop, 5 DFMA per term\", \"blocks_per_sm\": %d, \"threads\": %d, "
         "\"warps_per_sm\": %d, \"instr_lanes_per_clk_per_sm\": %.2f, "
         "\"cycles_per_warp_instr_per_scheduler\": %.3f}\n",
         blocks_per_sm, threads, blocks_per_sm * threads / 32, ops_per_sm / mean,
         mean / (ops_per_sm / 32.0 / 4.0));
  delete[] h;
  cudaFree(out);
  cudaFree(cyc);
}

template <typename T, int MODE>
void run(const char* name, int sms, int blocks_per_sm, int threads, int ops_per_iter) {
  const int iters = 4096;
  const int grid = sms * blocks_per_sm;
  T* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(T) * grid * threads);
  cudaMalloc(&cyc, sizeof(long long) * grid);
  chains<T, MODE><<<grid, threads>>>(out, cyc, 16, (T)1.0000001, (T)1e-9);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  chains<T, MODE><<<grid, threads>>>(out, cyc, iters, (T)1.0000001, (T)1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long* h = new long long[grid];
  cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += (double)h[i];
  mean /= grid;
  const double ops_per_sm = (double)blocks_per_sm * threads * 8.0 * iters * ops_per_iter;
  printf("{\"op\": \"%s\", \"blocks_per_sm\": %d, \"threads\": %d, \"warps_per_sm\": %d, "
         "\"instr_lanes_per_clk_per_sm\": %.2f, \"cycles_per_warp_instr_per_scheduler\": %.3f, "
         "\"ms\": %.4f, \"T_instr_per_s\": %.3f}\n",
         name, blocks_per_sm, threads, blocks_per_sm * threads / 32, ops_per_sm / mean,
         mean / (ops_per_sm / 32.0 / 4.0), ms, ops_per_sm * sms / (ms * 1e-3) / 1e12);
  delete[] h;
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, sms);
  const int shapes[][2] = {{1, 128}, {1, 256}, {1, 512}, {1, 1024}, {2, 1024}};
  for (auto& s : shapes) run<double, 0>("DFMA", sms, s[0], s[1], 1);
  for (auto& s : shapes) run<double, 1>("DMUL+DADD", sms, s[0], s[1], 2);
  run_window(sms, 1, 128);
  run_window(sms, 1, 256);
  run_window(sms, 1, 512);
  run_window(sms, 2, 288);
  run_window(sms, 1, 1024);
  run<float, 0>("FFMA", sms, 1, 1024, 1);
  run<float, 0>("FFMA", sms, 2, 1024, 1);
  return 0;
}
