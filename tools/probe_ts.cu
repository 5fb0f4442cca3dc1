// Hardware probes for the two mechanisms the streaming kernels rely on (run on a B200: nvcc -arch=sm_100a, ./probe_ts):
//   1. tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st from registers) and B in shared memory in
//      the MN-major SWIZZLE_NONE layout whose N chunks are the same 8-channel plane advanced by one pixel.
//   2. cp.async.bulk.tensor (TMA) tile loads of fp32 planes with negative / out-of-range start coordinates (zero fill).
// Development tool only: not part of the library.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#include "../autoencoder-fft_b200/csrc/umma.cuh"
#include "../autoencoder-fft_b200/csrc/tma.cuh"

using namespace aefft::umma;
using namespace aefft::tma;

constexpr int N = 48, K = 16, NS = 6;

__global__ void ts_probe(const float* __restrict__ Ain /*[128][16]*/, const float* __restrict__ Sin /*[K+NS][8]*/,
                         float* __restrict__ out /*[128][N]*/) {
  __shared__ __align__(128) __nv_bfloat16 S[(K + NS + 2) * 8];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  if (tid == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = tid; i < (K + NS + 2) * 8; i += 128) S[i] = __float2bfloat16_rn(i < (K + NS) * 8 ? Sin[i] : 0.f);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_slot;
  // A row = lane = tid: 16 bf16 -> 8 packed registers (even k in the low half)
  uint32_t r[8];
  for (int c = 0; c < 8; c++)
    r[c] = pack2(__float2bfloat16_rn(Ain[tid * 16 + 2 * c]), __float2bfloat16_rn(Ain[tid * 16 + 2 * c + 1]));
  tmem_st8(tb + ((uint32_t)(warp * 32) << 16) + 0, r);
  tmem_wait_st();
  fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 1);
    const uint64_t bdesc = make_desc(smem_u32(S), 128, 16);
    mma_bf16_ts(tb + 64, tb + 0, bdesc, idesc, false);
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + 64 + c0, v);
    for (int e = 0; e < 16; e++) out[tid * N + c0 + e] = v[e];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 128);
}

constexpr int BX = 16, BY = 2, BP = 3;
__device__ __forceinline__ bool mbar_wait_timeout(uint64_t* bar, uint32_t parity, long long max_cycles) {
  const long long t0 = clock64();
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done && clock64() - t0 > max_cycles) return false;
  }
  return true;
}
__global__ void tma_probe(const __grid_constant__ CUtensorMap tm, int cj, int ci, int cp, float* out, int* status) {
  __shared__ __align__(128) float buf[BP * BY * BX];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < BP * BY * BX; i += blockDim.x) buf[i] = -7.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, sizeof(buf));
    tma_load_3d(buf, &tm, cj, ci, cp, &bar);
  }
  const bool ok = mbar_wait_timeout(&bar, 0, 200000000LL);
  if (threadIdx.x == 0) *status = ok ? 1 : 0;
  __syncthreads();
  for (int i = threadIdx.x; i < BP * BY * BX; i += blockDim.x) out[i] = buf[i];
}

// 128-byte swizzle: box {32 px, 1 row, 16 planes}
__global__ void tma_swz_probe(const __grid_constant__ CUtensorMap tm, int cj, int ci, int cp, float* out, int* status) {
  __shared__ __align__(1024) float buf[16 * 32];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, sizeof(buf));
    tma_load_3d(buf, &tm, cj, ci, cp, &bar);
  }
  const bool ok = mbar_wait_timeout(&bar, 0, 200000000LL);
  if (threadIdx.x == 0) *status = ok ? 1 : 0;
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) out[i] = buf[i];
}

// Probe 3: MN-major B operand in SWIZZLE_32B layout, 16 channels per pixel (32 B per K row), N atoms one PIXEL apart
// (LBO = 32 B): does N = 16 x NS cover NS pixel shifts of a 16-channel plane?  mode 0: LBO=32,SBO=256; mode 1: swapped.
constexpr int NS3 = 6;
template <int CH>
__global__ void ts_probe_sw(const float* __restrict__ Ain /*[128][16]*/, const float* __restrict__ Sin /*[K+NS3][CH]*/,
                            float* __restrict__ out /*[128][CH*NS3]*/, int mode) {
  constexpr int N3 = CH * NS3, PXB = CH * 2;  // bytes per pixel row
  __shared__ __align__(1024) unsigned char S[2048];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = tid; i < 2048; i += 128) S[i] = 0;
  __syncthreads();
  for (int i = tid; i < (K + NS3) * CH; i += 128) {
    const int px = i / CH, ch = i % CH;
    uint32_t off = px * PXB + (ch / 8) * 16 + (ch % 8) * 2;
    if (CH == 16) off ^= ((off >> 7) & 1) << 4;   // 32-byte swizzle: address bit 4 ^= bit 7
    else off ^= ((off >> 7) & 3) << 4;            // 64-byte swizzle: bits [4,6) ^= bits [7,9)
    *reinterpret_cast<__nv_bfloat16*>(S + off) = __float2bfloat16_rn(Sin[i]);
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_slot;
  uint32_t r[8];
  for (int c = 0; c < 8; c++)
    r[c] = pack2(__float2bfloat16_rn(Ain[tid * 16 + 2 * c]), __float2bfloat16_rn(Ain[tid * 16 + 2 * c + 1]));
  tmem_st8(tb + ((uint32_t)(warp * 32) << 16) + 0, r);
  tmem_wait_st();
  fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, N3, 0, 1);
    uint64_t bdesc = make_desc(smem_u32(S), PXB, 8 * PXB);  // LBO = one pixel, SBO = 8 pixels
    bdesc |= (uint64_t)(CH == 16 ? 6 : 4) << 61;               // SWIZZLE_32B / SWIZZLE_64B
    (void)mode;
    mma_bf16_ts(tb + 64, tb + 0, bdesc, idesc, false);
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < N3; c0 += 16) {
    float v[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + 64 + c0, v);
    for (int e = 0; e < 16; e++) out[tid * N3 + c0 + e] = v[e];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main(int argc, char** argv) {
  // ---------------- probe 1
  float hA[128 * 16], hS[(K + NS) * 8], hO[128 * N];
  srand(1);
  for (auto& v : hA) v = bf((rand() % 2001 - 1000) / 500.f);
  for (auto& v : hS) v = bf((rand() % 2001 - 1000) / 500.f);
  float *dA, *dS, *dO;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dS, sizeof(hS)); cudaMalloc(&dO, sizeof(hO));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dS, hS, sizeof(hS), cudaMemcpyHostToDevice);
  ts_probe<<<1, 128>>>(dA, dS, dO);
  cudaError_t e = cudaDeviceSynchronize();
  printf("ts_probe: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(hO, dO, sizeof(hO), cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 128; m++)
    for (int s = 0; s < NS; s++)
      for (int x = 0; x < 8; x++) {
        double ref = 0;
        for (int k = 0; k < K; k++) ref += (double)hA[m * 16 + k] * hS[(k + s) * 8 + x];
        maxerr = fmax(maxerr, fabs(ref - hO[m * N + s * 8 + x]));
      }
  printf("ts_probe max abs err %.3g  (%s)\n", maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
  // ---------------- probe 3
  if (argc == 2) {
    const int CH = atoi(argv[1]);  // 16 or 32
    const int N3 = CH * NS3;
    std::vector<float> hS3((K + NS3) * CH), hO3(128 * N3);
    for (auto& v : hS3) v = bf((rand() % 2001 - 1000) / 500.f);
    float *dS3, *dO3;
    cudaMalloc(&dS3, hS3.size() * 4); cudaMalloc(&dO3, hO3.size() * 4);
    cudaMemcpy(dS3, hS3.data(), hS3.size() * 4, cudaMemcpyHostToDevice);
    if (CH == 16) ts_probe_sw<16><<<1, 128>>>(dA, dS3, dO3, 0);
    else ts_probe_sw<32><<<1, 128>>>(dA, dS3, dO3, 0);
    e = cudaDeviceSynchronize();
    printf("ts_probe_sw CH=%d: %s\n", CH, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(hO3.data(), dO3, hO3.size() * 4, cudaMemcpyDeviceToHost);
    double me = 0;
    int nbad = 0;
    for (int m = 0; m < 128; m++)
      for (int s = 0; s < NS3; s++)
        for (int x = 0; x < CH; x++) {
          double ref = 0;
          for (int k = 0; k < K; k++) ref += (double)hA[m * 16 + k] * hS3[(k + s) * CH + x];
          const double d = fabs(ref - hO3[m * N3 + s * CH + x]);
          me = fmax(me, d);
          if (d > 1e-3) nbad++;
        }
    printf("ts_probe_sw CH=%d max abs err %.3g, %d of %d wrong (%s)\n", CH, me, nbad, 128 * N3, me < 1e-3 ? "OK" : "MISMATCH");
    return 0;
  }
  // ---------------- probe 2
  const int P = 6, NX = 10, NY = 24;
  float hT[P * NX * NY];
  for (int i = 0; i < P * NX * NY; i++) hT[i] = (float)i + 1.f;
  float *dT, *dB;
  cudaMalloc(&dT, sizeof(hT)); cudaMalloc(&dB, BP * BY * BX * 4);
  cudaMemcpy(dT, hT, sizeof(hT), cudaMemcpyHostToDevice);
  CUtensorMap tm;
  int rc = make_tmap_3d_f32(&tm, dT, NY, NX, P, BX, BY, BP);
  printf("tensor map rc=%d\n", rc);
  if (rc) return 1;
  int bad = 0;
  int cases[3][3] = {{4, 3, 0}, {12, 9, 4}, {-4, -1, 2}};
  if (argc == 4) for (int q = 0; q < 3; q++) cases[2][q] = atoi(argv[1 + q]);
  for (auto& c : cases) {
    int* dSt; cudaMalloc(&dSt, 4);
    tma_probe<<<1, 64>>>(tm, c[0], c[1], c[2], dB, dSt);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tma_probe: %s\n", cudaGetErrorString(e)); return 1; }
    int st = -1; cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost);
    printf("tma case (%d,%d,%d): barrier %s\n", c[0], c[1], c[2], st == 1 ? "completed" : "TIMED OUT");
    float hB[BP * BY * BX];
    cudaMemcpy(hB, dB, sizeof(hB), cudaMemcpyDeviceToHost);
    for (int p = 0; p < BP; p++)
      for (int i = 0; i < BY; i++)
        for (int j = 0; j < BX; j++) {
          const int gp = c[2] + p, gi = c[1] + i, gj = c[0] + j;
          const float ref = (gp >= 0 && gp < P && gi >= 0 && gi < NX && gj >= 0 && gj < NY) ? hT[(gp * NX + gi) * NY + gj] : 0.f;
          if (hB[(p * BY + i) * BX + j] != ref) bad++;
        }
  }
  printf("tma_probe mismatches %d (%s)\n", bad, bad ? "MISMATCH" : "OK");
  {
    const int P2 = 40, NX2 = 6, NY2 = 72;
    float* hT2 = (float*)malloc(P2 * NX2 * NY2 * 4);
    for (int i = 0; i < P2 * NX2 * NY2; i++) hT2[i] = (float)i + 1.f;
    float *dT2, *dB2; int* dSt;
    cudaMalloc(&dT2, P2 * NX2 * NY2 * 4); cudaMalloc(&dB2, 16 * 32 * 4); cudaMalloc(&dSt, 4);
    cudaMemcpy(dT2, hT2, P2 * NX2 * NY2 * 4, cudaMemcpyHostToDevice);
    CUtensorMap tm2;
    rc = make_tmap_3d_f32(&tm2, dT2, NY2, NX2, P2, 32, 1, 16, true);
    printf("swizzled tensor map rc=%d\n", rc);
    if (rc) return 1;
    const int cj = 60, ci = 2, cp = 16;
    tma_swz_probe<<<1, 64>>>(tm2, cj, ci, cp, dB2, dSt);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tma_swz_probe: %s\n", cudaGetErrorString(e)); return 1; }
    int st = -1; cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost);
    float hB2[16 * 32];
    cudaMemcpy(hB2, dB2, sizeof(hB2), cudaMemcpyDeviceToHost);
    int bad2 = 0;
    for (int p = 0; p < 16; p++)
      for (int j = 0; j < 32; j++) {
        const int gj = cj + j;
        const float ref = gj < NY2 ? hT2[((cp + p) * NX2 + ci) * NY2 + gj] : 0.f;
        const int phys = p * 32 + ((((j >> 2) ^ (p & 7)) << 2) | (j & 3));
        if (hB2[phys] != ref) bad2++;
      }
    printf("tma_swz_probe barrier %s, mismatches %d (%s)\n", st == 1 ? "completed" : "TIMED OUT", bad2, bad2 ? "MISMATCH" : "OK");
    bad += bad2;
  }
  return (maxerr < 1e-3 && !bad) ? 0 : 2;
}
