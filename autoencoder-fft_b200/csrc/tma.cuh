// Tensor Memory Accelerator helpers: host-side tensor maps over planar fp32 feature maps, device-side tile loads
// (cp.async.bulk.tensor) that complete on an mbarrier.  Tile mode, no swizzle, out-of-bounds elements are filled with
// zeros - which is exactly the zero padding of the convolutions, so halo handling costs no thread instructions.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "umma.cuh"

namespace aefft {
namespace tma {

// ---- host: tensor map over a [n2][n1][n0] fp32 tensor (n0 contiguous), box [b2][b1][b0] ---------------------------------
// Requirements of the hardware: base 16-byte aligned, n0*4 and n0*n1*4 multiples of 16 bytes, b0*4 a multiple of 16,
// every box extent <= 256.  Returns 0 on success.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// swizzle (inner box extent must then be 32 floats = 128 bytes, destination 1024-byte aligned):
//   1 = SWIZZLE_128B: 16-byte chunk c of inner row r lands at chunk c ^ (r & 7) (r = linear row index b1*i2 + i1 of the box);
//   2 = SWIZZLE_128B_ATOM_32B: 32-byte chunk c of row r lands at chunk c ^ (r & 3) -- the only layout tcgen05.mma accepts
//       for MN-major (transposed) tf32 operands (UMMA layout type SWIZZLE_128B_BASE32B).
inline int make_tmap_3d_f32(CUtensorMap* tm, const float* base, uint64_t n0, uint64_t n1, uint64_t n2, uint32_t b0,
                            uint32_t b1, uint32_t b2, int swizzle = 0) {
  const bool swizzle128 = swizzle != 0;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1;
  if (((uintptr_t)base & 15) || (n0 * 4) % 16 || (b0 * 4) % 16 || b0 > 256 || b1 > 256 || b2 > 256) return -2;
  if (swizzle128 && b0 != 32) return -2;
  cuuint64_t dims[3] = {n0, n1, n2};
  cuuint64_t strides[2] = {n0 * 4, n0 * n1 * 4};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// ---- device ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
// box at element coordinates (c0 fastest, c1, c2) -> dense [b2][b1][b0] at `dst`; completes `bytes` on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          umma::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(umma::smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

}  // namespace tma
}  // namespace aefft
