// Throughput probe: scalar FFMA vs packed FFMA2 / FADD2 (sm_100) per SM and clock.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
  const float2 m = make_float2(s, s * 0.5f), c = make_float2(0.25f, 0.125f);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }          // 2 scalar FFMA
        else if (MODE == 1) a[i] = __ffma2_rn(a[i], m, c);                                             // 1 FFMA2
        else if (MODE == 2) a[i] = __fadd2_rn(a[i], c);                                                // 1 FADD2
        else { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }                                         // 2 scalar FADD
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) r += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, float* out, int sms) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int occ = 1; occ <= 4; occ *= 2) {
    k<MODE><<<sms * occ, 256>>>(out, iters, 0.999f);
    cudaEventRecord(e0);
    k<MODE><<<sms * occ, 256>>>(out, iters, 0.999f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = (double)sms * occ * 256 * iters * 4 * 8 * 2;  // fp32 element operations
    printf("%-14s %d CTA/SM: %.3f ms, %.1f element-ops per SM per ns\n", name, occ, ms, lane_ops / sms / (ms * 1e6));
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 4 * 256 * 4);
  run<0>("FFMA scalar", out, p.multiProcessorCount);
  run<1>("FFMA2 packed", out, p.multiProcessorCount);
  run<3>("FADD scalar", out, p.multiProcessorCount);
  run<2>("FADD2 packed", out, p.multiProcessorCount);
  printf("clock %d kHz\n", p.clockRate);
  return 0;
}
