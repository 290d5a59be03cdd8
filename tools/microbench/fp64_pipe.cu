// FP64 pipe microbenchmark for B200 (sm_100a): what one SM sub-partition sustains for the FP64
// instruction kinds the trace kernel issues, with register (not constant-bank) operands.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipe fp64_pipe.cu && ./fp64_pipe
// Prints warp-instructions per clock per SM sub-partition (peak of a 16-lane pipe = 0.5).
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
template <int KIND>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b, double c) {
  double x[CHAINS], y[CHAINS];
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) { x[j] = a + threadIdx.x + j; y[j] = b + j; }
  double p = b, q = c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) {
      if (KIND == 0) x[j] = fma(x[j], y[j], y[(j + 1) % CHAINS]);          // DFMA, 3 register operands
      if (KIND == 1) x[j] = x[j] * y[j];                                    // DMUL
      if (KIND == 2) x[j] = x[j] + y[j];                                    // DADD
      if (KIND == 3) x[j] = fma(x[j], p, q);                                // DFMA, 2 shared operands
      if (KIND == 4) { x[j] = fma(x[j], y[j], y[(j + 1) % CHAINS]); y[j] = x[j] > p ? y[j] : q; }  // DFMA + DSETP + 2 FSEL
      if (KIND == 5) { x[j] = fma(x[j], y[j], y[(j + 1) % CHAINS]); x[j] = __dadd_rn(__dmul_rn(x[j], y[j]), q); }  // mix (no contraction)
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) s += x[j] + y[j];
  if (s == 123.456) out[0] = s;
}

template <int KIND>
void run(const char* name, int per_iter_fp64, int threads, int blocks_per_sm) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND><<<sms * blocks_per_sm, threads>>>(out, 100, 1.0, 0.999999, 1e-9);
  cudaEventRecord(e0);
  k<KIND><<<sms * blocks_per_sm, threads>>>(out, iters, 1.0, 0.999999, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_inst = (double)sms * blocks_per_sm * (threads / 32) * (double)iters * CHAINS * per_iter_fp64;
  const double clocks = ms * 1e-3 * khz * 1e3;
  printf("%-34s %2d warps/SM: %.3f FP64 warp-inst/clk/SMSP (%.1f%% of 0.5)\n", name, threads / 32 * blocks_per_sm,
         warp_inst / clocks / sms / 4, 100 * warp_inst / clocks / sms / 4 / 0.5);
  cudaFree(out);
}

// issue-slot sharing: NI independent integer multiply-adds per DFMA (does an FP64 instruction hold the issue
// port for its second pipe cycle, or can another pipe's instruction go in between?)
template <int NI>
__global__ void __launch_bounds__(256) mixk(double* out, int iters, double a, double b, double c, int m) {
  double x[CHAINS];
  int k[CHAINS][3];
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) { x[j] = a + threadIdx.x + j; k[j][0] = threadIdx.x + j; k[j][1] = j * 3; k[j][2] = j * 7 + 1; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) {
      x[j] = fma(x[j], b, c);
#pragma unroll
      for (int q = 0; q < NI; ++q) k[j][q] = k[j][q] * m + i;
    }
  }
  double s = 0;
  int t = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) { s += x[j]; t += k[j][0] + k[j][1] + k[j][2]; }
  if (s == 123.456 || t == 12345) out[0] = s + t;
}
template <int NI>
void run_mix() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 20000, threads = 192, bps = 2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  mixk<NI><<<sms * bps, threads>>>(out, 100, 1.0, 0.999999, 1e-9, 3);
  cudaEventRecord(e0);
  mixk<NI><<<sms * bps, threads>>>(out, iters, 1.0, 0.999999, 1e-9, 3);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double clocks = ms * 1e-3 * khz * 1e3;
  const double groups = (double)sms * bps * (threads / 32) * (double)iters * CHAINS;  // (1 DFMA + NI IMAD) groups
  printf("1 DFMA + %d IMAD: %.2f clk per group per SMSP (pipe alone: 2.00, issue slots: %d)\n", NI,
         clocks * sms * 4 / groups, 1 + NI);
  cudaFree(out);
}

// issue-port test with FULL-RATE companions: per DFMA, NA independent FFMA (FP32 pipe, 32 lanes/clk per sub-partition:
// one issue cycle each).  If an FP64 instruction held the issue port for one cycle only, a group of 1 DFMA + NA FFMA
// would take max(2, 1 + NA) clocks; if it holds it for both cycles of its 16-lane pipe, 2 + NA.
template <int NA>
__global__ void __launch_bounds__(256) mixalu(double* out, int iters, double a, double b, double c, float fb, float fc) {
  double x[CHAINS];
  float f[CHAINS][4];
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) {
    x[j] = a + threadIdx.x + j;
#pragma unroll
    for (int q = 0; q < 4; ++q) f[j][q] = (float)(threadIdx.x * (q + 1) + j);
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) {
      x[j] = fma(x[j], b, c);
#pragma unroll
      for (int q = 0; q < NA; ++q) f[j][q] = fmaf(f[j][q], fb, fc);
    }
  }
  double s = 0;
  float t = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) { s += x[j]; t += f[j][0] + f[j][1] + f[j][2] + f[j][3]; }
  if (s == 123.456 || t == 12345.f) out[0] = s + t;
}
template <int NA>
void run_mixalu() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 20000, threads = 256, bps = 2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  mixalu<NA><<<sms * bps, threads>>>(out, 100, 1.0, 0.999999, 1e-9, 0.999f, 1e-3f);
  cudaEventRecord(e0);
  mixalu<NA><<<sms * bps, threads>>>(out, iters, 1.0, 0.999999, 1e-9, 0.999f, 1e-3f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double clocks = ms * 1e-3 * khz * 1e3;
  const double groups = (double)sms * bps * (threads / 32) * (double)iters * CHAINS;
  printf("1 DFMA + %d FFMA: %.2f clk per group per SMSP (one-cycle issue port: %d, two-cycle: %d)\n", NA,
         clocks * sms * 4 / groups, 1 + NA > 2 ? 1 + NA : 2, 2 + NA);
  cudaFree(out);
}

// dependent-issue latency: NCH independent DFMA chains per thread, W warps per SM sub-partition
template <int NCH>
__global__ void __launch_bounds__(1024) lat(double* out, int iters, double a, double b, double c) {
  double x[NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) x[j] = a + threadIdx.x + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) x[j] = fma(x[j], b, c);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < NCH; ++j) s += x[j];
  if (s == 123.456) out[0] = s;
}
template <int NCH>
void run_lat(int warps_per_smsp) {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 200000, threads = 128 * warps_per_smsp;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  lat<NCH><<<sms, threads>>>(out, 100, 1.0, 0.999999, 1e-9);
  cudaEventRecord(e0);
  lat<NCH><<<sms, threads>>>(out, iters, 1.0, 0.999999, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double clocks = ms * 1e-3 * khz * 1e3;
  printf("chains %d, warps/SMSP %d: %.2f clk per loop trip -> %.3f DFMA/clk/SMSP\n", NCH, warps_per_smsp,
         clocks / iters, (double)NCH * warps_per_smsp * iters / clocks);
  cudaFree(out);
}

int main() {
  run_mix<0>(); run_mix<1>(); run_mix<2>(); run_mix<3>();
  run_mixalu<0>(); run_mixalu<1>(); run_mixalu<2>(); run_mixalu<3>(); run_mixalu<4>();
  run_lat<1>(1); run_lat<2>(1); run_lat<4>(1); run_lat<8>(1);
  run_lat<1>(3); run_lat<2>(3); run_lat<4>(3);
  run_lat<1>(4); run_lat<2>(4); run_lat<2>(6); run_lat<2>(8);
  for (int occ = 0; occ < 2; ++occ) {
    const int threads = occ ? 256 : 192, bps = occ ? 4 : 2;
    run<3>("DFMA x*p+q (shared operands)", 1, threads, bps);
    run<0>("DFMA 3 register operands", 1, threads, bps);
    run<1>("DMUL", 1, threads, bps);
    run<2>("DADD", 1, threads, bps);
    run<4>("DFMA + DSETP + 2 FSEL", 2, threads, bps);
    run<5>("DFMA, DMUL, DADD dependent", 3, threads, bps);
  }
  return 0;
}
