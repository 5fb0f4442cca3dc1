// Shared-memory operand staging for the tcgen05 kernels: planar fp32 feature maps in HBM -> 8-channel bf16 hi/lo
// planes [plane][linear grid pixel][8] (16 bytes per pixel per plane), the SWIZZLE_NONE canonical layout both the
// K-major (conv) and the MN-major (weight-gradient) descriptors read.
#pragma once
#include "umma.cuh"

namespace aefft {

// Grid pixel h = r*PJ + c maps to image pixel (row0 + r, col0 + c).  Pixels outside [lo, Nx) x [lo, Ny), rows >=
// rows_valid, columns >= cols_valid and channels >= nch are written as zeros.  Value = (s0 - s1) or s0 * scale.
// One work item = 4 consecutive grid pixels x 8 channels: all (up to 64) global loads of an item are issued before the
// first use, so every thread keeps 32-64 independent 4-byte loads in flight (the staging loops are latency bound).
// npx must be a multiple of 4.
template <int NTHREADS>
__device__ __forceinline__ void stage_planes4(unsigned char* hi_base, unsigned char* lo_base, uint32_t plane_bytes, int nplanes,
                                              int npx, int PJ, const float* __restrict__ s0, const float* __restrict__ s1,
                                              float scale, int ch0, int nch, long long plane, int Nx, int Ny, int row0,
                                              int col0, int rows_valid, int cols_valid, int lo, int tid) {
  using namespace umma;
  const int ngroups = npx >> 2;
  for (int idx = tid; idx < nplanes * ngroups; idx += NTHREADS) {
    const int pl = idx / ngroups, h0 = (idx - pl * ngroups) << 2;
    int r = h0 / PJ, c = h0 - r * PJ;
    const int c0 = pl * 8;
    bool inb[4];
    long long pix[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int si = row0 + r, sj = col0 + c;
      inb[q] = r < rows_valid && c < cols_valid && si >= lo && si < Nx && sj >= lo && sj < Ny;
      pix[q] = inb[q] ? (long long)si * Ny + sj : 0;
      if (++c == PJ) { c = 0; r++; }
    }
    float v[4][8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const bool ch_ok = c0 + e < nch;
      const float* src = s0 + (long long)(ch0 + (ch_ok ? c0 + e : 0)) * plane;
#pragma unroll
      for (int q = 0; q < 4; q++) v[q][e] = (ch_ok && inb[q]) ? __ldg(src + pix[q]) : 0.f;
    }
    if (s1) {
      float u[4][8];
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const bool ch_ok = c0 + e < nch;
        const float* src = s1 + (long long)(ch0 + (ch_ok ? c0 + e : 0)) * plane;
#pragma unroll
        for (int q = 0; q < 4; q++) u[q][e] = (ch_ok && inb[q]) ? __ldg(src + pix[q]) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int e = 0; e < 8; e++) v[q][e] -= u[q][e];
    } else if (scale != 1.f) {
#pragma unroll
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int e = 0; e < 8; e++) v[q][e] *= scale;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      __nv_bfloat16 hi[8], lw[8];
#pragma unroll
      for (int e = 0; e < 8; e++) split_bf16(v[q][e], hi[e], lw[e]);
      const size_t off = (size_t)pl * plane_bytes + (size_t)(h0 + q) * 16;
      *reinterpret_cast<uint4*>(hi_base + off) =
          make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
      *reinterpret_cast<uint4*>(lo_base + off) =
          make_uint4(pack2(lw[0], lw[1]), pack2(lw[2], lw[3]), pack2(lw[4], lw[5]), pack2(lw[6], lw[7]));
    }
  }
}

}  // namespace aefft
