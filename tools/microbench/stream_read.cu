// How fast can ANY kernel read the detector kernel's inputs?  8 FP64 columns + 1 byte column of n rays
// (65 B/ray), two rays per thread with 128-bit loads, nothing but a checksum computed.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_read stream_read.cu && ./stream_read
#include <cstdio>
#include <cuda_runtime.h>

template <int UNROLL>
__global__ void __launch_bounds__(256) read_cols(const double* __restrict__ base, const unsigned char* __restrict__ flags,
                                                 long long n, long long cap, double* out) {
  double acc = 0.0;
  const long long npairs = n / 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += stride * UNROLL) {
    double2 v[UNROLL][8];
    uchar2 f[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long q = p + u * stride;
      if (q < npairs) {
        f[u] = *reinterpret_cast<const uchar2*>(flags + 2 * q);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[u][c] = *reinterpret_cast<const double2*>(base + c * cap + 2 * q);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long q = p + u * stride;
      if (q < npairs) {
#pragma unroll
        for (int c = 0; c < 8; ++c) acc += v[u][c].x + v[u][c].y;
        acc += f[u].x + f[u].y;
      }
    }
  }
  if (acc == 123.456) out[0] = acc;
}

int main() {
  const long long n = 10000000, cap = (n + 1023) & ~1023LL;
  double* base;
  unsigned char* flags;
  double* out;
  cudaMalloc(&base, sizeof(double) * 8 * cap);
  cudaMalloc(&flags, cap);
  cudaMalloc(&out, 8);
  cudaMemset(base, 0, sizeof(double) * 8 * cap);
  cudaMemset(flags, 1, cap);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  // a buffer larger than L2 written between runs
  double* flush;
  cudaMalloc(&flush, 512 << 20);
  for (int blocks_per_sm = 2; blocks_per_sm <= 8; blocks_per_sm *= 2) {
    for (int unroll = 1; unroll <= 2; ++unroll) {
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaMemsetAsync(flush, 0, 512 << 20);
        cudaEventRecord(e0);
        if (unroll == 1) read_cols<1><<<sms * blocks_per_sm, 256>>>(base, flags, n, cap, out);
        else read_cols<2><<<sms * blocks_per_sm, 256>>>(base, flags, n, cap, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      printf("blocks/SM %d, pairs in flight per thread %d: %.1f us, %.2f TB/s\n", blocks_per_sm, unroll, best * 1e3,
             65.0 * n / (best * 1e-3) / 1e12);
    }
  }
  return 0;
}
