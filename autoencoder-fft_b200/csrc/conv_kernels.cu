// Coordinate-space convolution kernels (forward conv and its transpose), fp32 on CUDA cores.
//
// Replaces conv_parallel (reference backproplib.cu:70-111: one thread per output, no reuse) and the hidden-delta
// recomputation buried in gradient_C* (backproplib.cu:186-288, 424-518: delta_h recomputed dD*Nk*Nl times).
// One CTA computes a 16x64 pixel tile for MB output channels of one frame; the input tile (+halo) and the weight
// slice are staged in shared memory per chunk of DC input channels; each thread keeps a 4-pixel x MB-channel
// register block, reads its 4+Nl-1 input pixels with 128-bit LDS and the weights with broadcast 128-bit LDS
// (1 LDS.128 per ~15 FFMA).  The fused variants form e = out - in (src1) or in/dM (pre_div) at tile-load time so
// those tensors never exist in HBM.
#include "common.cuh"

namespace aefft {

struct ConvParams {
  const float* src0;
  const float* src1;
  const float* w;
  const float* bias;
  float* out;
  long long w_so, w_sc;
  float pre_div;
  int C, O, Nx, Ny;
  int ai0, aj0, flip, lo;
  int tiles_j, tiles_i, o_blocks;
  int vec_ok;
};

constexpr int TI = 16, TJ = 64, CONV_THREADS = 256;

template <int NK, int NL, int MB, int DC>
__global__ void __launch_bounds__(CONV_THREADS, 2) conv_tile_kernel(ConvParams p) {
  constexpr int HI = TI + NK - 1;
  constexpr int PJ = ((TJ + NL - 1) + 3) / 4 * 4;
  constexpr int XV = (4 + NL - 1 + 3) / 4;  // float4s per thread row segment
  __shared__ __align__(16) float xs[DC][HI][PJ];
  __shared__ __align__(16) float ws[DC][NK][NL][MB];

  const int tid = threadIdx.x;
  const int tj = tid % (TJ / 4), ti = tid / (TJ / 4);
  const int tile_j = blockIdx.x, tile_i = blockIdx.y;
  const int ob = blockIdx.z % p.o_blocks;
  const long long b = blockIdx.z / p.o_blocks;
  const int i0 = tile_i * TI, j0 = tile_j * TJ, o0 = ob * MB;
  const long long plane = (long long)p.Nx * p.Ny;
  const float* s0 = p.src0 + b * p.C * plane;
  const float* s1 = p.src1 ? p.src1 + b * p.C * plane : nullptr;

  float acc[MB][4];
#pragma unroll
  for (int o = 0; o < MB; o++)
#pragma unroll
    for (int q = 0; q < 4; q++) acc[o][q] = 0.f;

  for (int c0 = 0; c0 < p.C; c0 += DC) {
    // ---- stage input tile (+halo) ----
    for (int idx = tid; idx < DC * HI * PJ; idx += CONV_THREADS) {
      int cc = idx / (HI * PJ);
      int r = (idx / PJ) % HI;
      int col = idx % PJ;
      int si = i0 + p.ai0 + r, sj = j0 + p.aj0 + col;
      float v = 0.f;
      if (c0 + cc < p.C && si >= p.lo && si < p.Nx && sj >= p.lo && sj < p.Ny) {
        long long off = (long long)(c0 + cc) * plane + (long long)si * p.Ny + sj;
        v = __ldg(s0 + off);
        if (s1) v -= __ldg(s1 + off);
        else if (p.pre_div != 0.f) v = __fdiv_rn(v, p.pre_div);
      }
      xs[cc][r][col] = v;
    }
    // ---- stage weights: ws[cc][tk][tl][o] ----
    for (int idx = tid; idx < DC * NK * NL * MB; idx += CONV_THREADS) {
      int o = idx % MB;
      int tl = (idx / MB) % NL;
      int tk = (idx / (MB * NL)) % NK;
      int cc = idx / (MB * NL * NK);
      int k = p.flip ? NK - 1 - tk : tk, l = p.flip ? NL - 1 - tl : tl;
      float v = 0.f;
      if (c0 + cc < p.C && o0 + o < p.O) v = __ldg(p.w + (o0 + o) * p.w_so + (c0 + cc) * p.w_sc + k * NL + l);
      ws[cc][tk][tl][o] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int cc = 0; cc < DC; cc++) {
#pragma unroll
      for (int tk = 0; tk < NK; tk++) {
        float x[4 * XV];
        const float4* xr = reinterpret_cast<const float4*>(&xs[cc][ti + tk][4 * tj]);
#pragma unroll
        for (int v = 0; v < XV; v++) {
          float4 t = xr[v];
          x[4 * v] = t.x; x[4 * v + 1] = t.y; x[4 * v + 2] = t.z; x[4 * v + 3] = t.w;
        }
#pragma unroll
        for (int tl = 0; tl < NL; tl++) {
          const float4* wr = reinterpret_cast<const float4*>(&ws[cc][tk][tl][0]);
#pragma unroll
          for (int o4 = 0; o4 < MB / 4; o4++) {
            float4 wv = wr[o4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
              acc[4 * o4 + 0][q] = fmaf(wv.x, x[q + tl], acc[4 * o4 + 0][q]);
              acc[4 * o4 + 1][q] = fmaf(wv.y, x[q + tl], acc[4 * o4 + 1][q]);
              acc[4 * o4 + 2][q] = fmaf(wv.z, x[q + tl], acc[4 * o4 + 2][q]);
              acc[4 * o4 + 3][q] = fmaf(wv.w, x[q + tl], acc[4 * o4 + 3][q]);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  // ---- epilogue ----
  const int i = i0 + ti, j = j0 + 4 * tj;
  if (i >= p.Nx || j >= p.Ny) return;
  float* ob_ptr = p.out + (b * p.O + o0) * plane + (long long)i * p.Ny + j;
#pragma unroll
  for (int o = 0; o < MB; o++) {
    if (o0 + o >= p.O) break;
    float bv = p.bias ? __ldg(p.bias + o0 + o) : 0.f;
    float* dst = ob_ptr + o * plane;
    if (p.vec_ok && j + 3 < p.Ny) {
      float4 v = make_float4(acc[o][0] + bv, acc[o][1] + bv, acc[o][2] + bv, acc[o][3] + bv);
      *reinterpret_cast<float4*>(dst) = v;
    } else {
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (j + q < p.Ny) dst[q] = acc[o][q] + bv;
    }
  }
}

// Generic fallback for tap shapes without a tiled instantiation (any Nk,Nl): one thread per output.
__global__ void conv_generic_kernel(ConvParams p, int Nk, int Nl, long long total) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= total) return;
  const long long plane = (long long)p.Nx * p.Ny;
  int j = n % p.Ny;
  int i = (n / p.Ny) % p.Nx;
  int o = (n / plane) % p.O;
  long long b = n / (plane * p.O);
  const float* s0 = p.src0 + b * p.C * plane;
  const float* s1 = p.src1 ? p.src1 + b * p.C * plane : nullptr;
  float acc = 0.f;
  for (int c = 0; c < p.C; c++)
    for (int tk = 0; tk < Nk; tk++) {
      int si = i + p.ai0 + tk;
      if (si < p.lo || si >= p.Nx) continue;
      for (int tl = 0; tl < Nl; tl++) {
        int sj = j + p.aj0 + tl;
        if (sj < p.lo || sj >= p.Ny) continue;
        long long off = c * plane + (long long)si * p.Ny + sj;
        float v = s0[off];
        if (s1) v -= s1[off];
        else if (p.pre_div != 0.f) v = __fdiv_rn(v, p.pre_div);
        int k = p.flip ? Nk - 1 - tk : tk, l = p.flip ? Nl - 1 - tl : tl;
        acc = fmaf(p.w[o * p.w_so + c * p.w_sc + k * Nl + l], v, acc);
      }
    }
  p.out[n] = acc + (p.bias ? p.bias[o] : 0.f);
}

template <int NK, int NL>
static int launch_tiled(aefft_ctx* ctx, ConvParams& p, int64_t B) {
  constexpr int MB = 16;
  p.o_blocks = (p.O + MB - 1) / MB;
  dim3 grid(p.tiles_j, p.tiles_i, (unsigned)(B * p.o_blocks));
  if (p.C % 4 == 0)
    conv_tile_kernel<NK, NL, MB, 4><<<grid, CONV_THREADS, 0, ctx->stream>>>(p);
  else if (p.C % 3 == 0)
    conv_tile_kernel<NK, NL, MB, 3><<<grid, CONV_THREADS, 0, ctx->stream>>>(p);
  else if (p.C % 2 == 0)
    conv_tile_kernel<NK, NL, MB, 2><<<grid, CONV_THREADS, 0, ctx->stream>>>(p);
  else
    conv_tile_kernel<NK, NL, MB, 1><<<grid, CONV_THREADS, 0, ctx->stream>>>(p);
  return AEFFT_OK;
}

int launch_conv(aefft_ctx* ctx, const Window& win, int64_t B, int C, int O, int Nx, int Ny, const float* src0,
                const float* src1, float pre_div, const float* w, int64_t w_so, int64_t w_sc, const float* bias,
                float* out) {
  AE_ARG(B > 0 && C > 0 && O > 0 && Nx > 0 && Ny > 0);
  if (ctx->precision != AEFFT_PRECISION_FP32) {
    const int passes = ctx->precision == AEFFT_PRECISION_BF16X3 ? 3 : 1;
    // warp-specialised pipeline when all weights fit in shared memory, else the two-CTA-per-SM kernel, else fp32
    int rc = launch_conv_rs(ctx, win, B, C, O, Nx, Ny, src0, src1, pre_div, w, w_so, w_sc, bias, out, passes);
    if (rc == AEFFT_ERR_UNSUPPORTED)
      rc = launch_conv_tc_ws(ctx, win, B, C, O, Nx, Ny, src0, src1, pre_div, w, w_so, w_sc, bias, out, passes);
    if (rc == AEFFT_ERR_UNSUPPORTED)
      rc = launch_conv_tc(ctx, win, B, C, O, Nx, Ny, src0, src1, pre_div, w, w_so, w_sc, bias, out, passes);
    if (rc != AEFFT_ERR_UNSUPPORTED) return rc;
  }
  ConvParams p;
  p.src0 = src0; p.src1 = src1; p.w = w; p.bias = bias; p.out = out;
  p.w_so = w_so; p.w_sc = w_sc; p.pre_div = pre_div;
  p.C = C; p.O = O; p.Nx = Nx; p.Ny = Ny;
  p.ai0 = win.ai0; p.aj0 = win.aj0; p.flip = win.flip; p.lo = win.lo;
  p.tiles_j = (Ny + TJ - 1) / TJ;
  p.tiles_i = (Nx + TI - 1) / TI;
  p.o_blocks = 1;
  p.vec_ok = (Ny % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  int64_t zdim = B * ((O + 15) / 16);
  bool tiled = zdim <= 65535;
  const double px = (double)B * Nx * Ny;
  ProfScope prof(ctx, win.flip ? "conv_fwd" : "conv_dgrad", 2.0 * px * C * O * win.Nk * win.Nl,
                 4.0 * (px * C * (src1 ? 2 : 1) + px * O + (double)C * O * win.Nk * win.Nl));
  if (tiled && win.Nk == 5 && win.Nl == 5) launch_tiled<5, 5>(ctx, p, B);
  else if (tiled && win.Nk == 3 && win.Nl == 3) launch_tiled<3, 3>(ctx, p, B);
  else if (tiled && win.Nk == 7 && win.Nl == 7) launch_tiled<7, 7>(ctx, p, B);
  else {
    long long total = (long long)B * O * Nx * Ny;
    conv_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(p, win.Nk, win.Nl, total);
  }
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

}  // namespace aefft
