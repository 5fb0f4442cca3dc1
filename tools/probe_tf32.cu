// Hardware probe for the momentum-space tensor-core contraction (csrc/spec_tc.cu): tcgen05.mma kind::tf32 with BOTH
// operands in shared memory in the 128-byte-swizzled canonical layouts, landed there by TMA (cp.async.bulk.tensor, fp32),
// K-major and MN-major, plus the 3xTF32 split (hi = x & 0xffffe000, lo = x - hi) accuracy against fp64.
// Development tool only: not part of the library.   nvcc -arch=sm_100a ... && ./probe_tf32
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <vector>

#include "../autoencoder-fft_b200/csrc/tma.cuh"
#include "../autoencoder-fft_b200/csrc/umma.cuh"

using namespace aefft::umma;
using namespace aefft::tma;

// SWIZZLE_128B descriptor: bits [61,64) = 2
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t type = 2) {
  return make_desc(addr, lbo, sbo) | ((uint64_t)type << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
  uint32_t p = acc ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(p)
      : "memory");
}

struct Params {
  CUtensorMap amap, bmap;
  int mn_major;  // 0: A [M][K], B [N][K] (K contiguous); 1: A [K][M], B [K][N] (M / N contiguous)
  int M, N, K;   // K multiple of 32 (K-major) or of 8 (MN-major, <= 32 here)
  int split;     // 1: 3xTF32 (hi/lo split in shared memory), 0: single pass on masked values
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b, kstep_bytes, ltype;
  float* out;    // [128][N]
};

__global__ void __launch_bounds__(128) probe(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout: A hi | A lo | B hi | B lo, each 32 KB max
  float* Ahi = (float*)(smem);
  float* Alo = (float*)(smem + 32768);
  float* Bhi = (float*)(smem + 65536);
  float* Blo = (float*)(smem + 98304);
  __shared__ __align__(8) uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid == 32) { mbar_init(&bar_tma, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_slot;
  const int kblocks = p.mn_major ? 1 : p.K / 32;
  uint32_t bytes = 0;
  if (tid == 0) {
    if (!p.mn_major) {
      // K-major: one box {32 floats of K, rows} per 32-float K block; block kb at +kb*rows*128 bytes
      for (int kb = 0; kb < kblocks; kb++) {
        tma_load_3d((uint8_t*)Ahi + kb * 128 * 128, &p.amap, kb * 32, 0, 0, &bar_tma);
        tma_load_3d((uint8_t*)Bhi + kb * p.N * 128, &p.bmap, kb * 32, 0, 0, &bar_tma);
      }
      bytes = kblocks * (128 * 128 + p.N * 128);
    } else {
      // MN-major: one box {32 floats of M/N, K rows} per 32-wide M/N block; block j at +j*K*128 bytes
      for (int j = 0; j < 128 / 32; j++) tma_load_3d((uint8_t*)Ahi + j * p.K * 128, &p.amap, j * 32, 0, 0, &bar_tma);
      for (int j = 0; j < p.N / 32; j++) tma_load_3d((uint8_t*)Bhi + j * p.K * 128, &p.bmap, j * 32, 0, 0, &bar_tma);
      bytes = (128 / 32 + p.N / 32) * p.K * 128;
    }
    mbar_expect_tx(&bar_tma, bytes);
  }
  mbar_wait(&bar_tma, 0);
  // split in place: hi = x & 0xffffe000 (exactly representable in tf32), lo = x - hi (exact in fp32)
  for (int i = tid; i < 8192; i += 128) {
    const uint32_t a = __float_as_uint(Ahi[i]), b = __float_as_uint(Bhi[i]);
    const float ah = __uint_as_float(a & 0xffffe000u), bh = __uint_as_float(b & 0xffffe000u);
    Alo[i] = Ahi[i] - ah; Blo[i] = Bhi[i] - bh;
    Ahi[i] = ah; Bhi[i] = bh;
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    fence_after_sync();
    const uint32_t idesc = idesc_tf32(128, p.N, p.mn_major, p.mn_major);
    bool acc = false;
    const int ksteps = p.mn_major ? p.K / 8 : 4;
    for (int kb = 0; kb < kblocks; kb++)
      for (int ks = 0; ks < ksteps; ks++) {
        const uint32_t ao = (p.mn_major ? 0 : kb * 128 * 128) + ks * p.kstep_bytes;
        const uint32_t bo = (p.mn_major ? 0 : kb * p.N * 128) + ks * p.kstep_bytes;
        const uint64_t ah = desc_sw128(smem_u32(Ahi) + ao, p.lbo_a, p.sbo_a, p.ltype), al = desc_sw128(smem_u32(Alo) + ao, p.lbo_a, p.sbo_a, p.ltype);
        const uint64_t bh = desc_sw128(smem_u32(Bhi) + bo, p.lbo_b, p.sbo_b, p.ltype), bl = desc_sw128(smem_u32(Blo) + bo, p.lbo_b, p.sbo_b, p.ltype);
        mma_tf32(tb, ah, bh, idesc, acc);
        acc = true;
        if (p.split) {
          mma_tf32(tb, ah, bl, idesc, true);
          mma_tf32(tb, al, bh, idesc, true);
        }
      }
    commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    float v[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int e = 0; e < 16; e++) p.out[tid * p.N + c0 + e] = v[e];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 256);
}

static double run(int mn_major, int M, int N, int K, int split, uint32_t lbo, uint32_t sbo, uint32_t kstep, bool masked_ref,
                  int tma_swz = 1, uint32_t ltype = 2) {
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  srand(7);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& v : B) v = ((float)rand() / RAND_MAX * 2 - 1) * 100.f;
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 128 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, 128 * N * 4);
  Params p;
  int rc;
  if (!mn_major) {  // A [M][K], B [N][K]
    rc = make_tmap_3d_f32(&p.amap, dA, K, M, 1, 32, 128, 1, tma_swz);
    rc |= make_tmap_3d_f32(&p.bmap, dB, K, N, 1, 32, N, 1, tma_swz);
  } else {          // A [K][M], B [K][N]
    rc = make_tmap_3d_f32(&p.amap, dA, M, K, 1, 32, K, 1, tma_swz);
    rc |= make_tmap_3d_f32(&p.bmap, dB, N, K, 1, 32, K, 1, tma_swz);
  }
  if (rc) { printf("tensor map failed %d\n", rc); return -1; }
  p.mn_major = mn_major; p.M = M; p.N = N; p.K = K; p.split = split;
  p.lbo_a = p.lbo_b = lbo; p.sbo_a = p.sbo_b = sbo; p.kstep_bytes = kstep; p.out = dO; p.ltype = ltype;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
  probe<<<1, 128, 131072 + 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
  std::vector<float> O((size_t)128 * N);
  cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0, scale = 0;
  auto mask = [&](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; };
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N; n++) {
      double s = 0;
      for (int k = 0; k < K; k++) {
        float a = mn_major ? A[(size_t)k * M + m] : A[(size_t)m * K + k];
        float b = mn_major ? B[(size_t)k * N + n] : B[(size_t)n * K + k];
        if (masked_ref) { a = mask(a); b = mask(b); }
        s += (double)a * b;
      }
      worst = fmax(worst, fabs(s - O[(size_t)m * N + n]));
      scale = fmax(scale, fabs(s));
    }
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
  return worst / scale;
}

int main() {
  // K-major, SW128: SBO = 1024 (8 rows x 128 B), LBO unused (16); K step of 8 floats = +32 bytes inside the atom
  printf("K-major  M128 N64  K32  1 pass (masked ref): rel err %.3e\n", run(0, 128, 64, 32, 0, 16, 1024, 32, true));
  printf("K-major  M128 N128 K64  1 pass (masked ref): rel err %.3e\n", run(0, 128, 128, 64, 0, 16, 1024, 32, true));
  printf("K-major  M128 N256 K32  1 pass (masked ref): rel err %.3e\n", run(0, 128, 256, 32, 0, 16, 1024, 32, true));
  printf("K-major  M128 N128 K64  3xTF32 (fp64 ref)  : rel err %.3e\n", run(0, 128, 128, 64, 1, 16, 1024, 32, false));
  printf("K-major  M96(oob) N64 K32 1 pass           : rel err %.3e\n", run(0, 96, 64, 32, 0, 16, 1024, 32, true));
  // MN-major tf32: the only accepted layout is SWIZZLE_128B_BASE32B (UMMA layout type 1): 32-byte chunks swizzled over 4-row
  // groups, landed by TMA with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Blocks of 32 M/N floats are K*128 bytes apart (LBO),
  // 4-row K groups 512 bytes apart (SBO); one MMA (K = 8) spans two groups, the next K step starts 1024 bytes further.
  for (int K : {8, 16, 32}) {
    printf("MN-major K%-2d N64  base32B (LBO=K*128, SBO=512) : rel err %.3e\n", K, run(1, 128, 64, K, 0, K * 128, 512, 1024, true, 2, 1));
    printf("MN-major K%-2d N64  base32B (LBO=512, SBO=K*128) : rel err %.3e\n", K, run(1, 128, 64, K, 0, 512, K * 128, 1024, true, 2, 1));
  }
  printf("MN-major K32 N128 base32B 3xTF32 (fp64 ref)      : rel err %.3e\n", run(1, 128, 128, 32, 1, 32 * 128, 512, 1024, false, 2, 1));
  printf("MN-major K32 N256 base32B 1 pass                 : rel err %.3e\n", run(1, 128, 256, 32, 0, 32 * 128, 512, 1024, true, 2, 1));
  printf("MN-major K32 N64 plain SW128 (type 2, TMA 128B)  : rel err %.3e\n", run(1, 128, 64, 32, 0, 32 * 128, 1024, 1024, true, 1, 2));
  return 0;
}
