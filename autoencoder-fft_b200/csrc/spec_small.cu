// Momentum-space training iteration for layer pairs with FEW input channels (the 3-channel image side: dD <= 4), fused.
//
// The reference's gradient_k_io (fft_backproplib.cu:395-475) is one thread per (m, d, bin) that recomputes G and H-hat
// for every d; the engine's generic path splits it into contractions that each stream a [frames][dM][bins] tensor
// through HBM (H, G written and read back: ~13 GB per iteration for 3 -> 16 channels at 512^2, 128 frames).  With 3 input
// channels everything a (bin, frame) needs fits in registers, so ONE kernel reads X (and the caller's O on the first
// iteration), forms H-hat, E, G on the fly and accumulates the frame-reduced outer products dC, dF in registers:
// HBM traffic = X + O + the gradient spectra (~1 GB).  A second kernel does the post-update re-forward + mse
// (conv_k twice + calc_mse, :1460-1463) without writing H or O at all.  Later iterations recompute O = F.H from X.
//
// Thread mapping: 4 consecutive lanes share a bin, each owns 4 hidden channels m (its rows of C, columns of F, and
// the matching 4 x dD blocks of dC / dF); sums over m (the decoder output O) are butterfly-reduced inside the lane group.
// Spectra are the reference layout [frame][channel][bins] / [m][d][bins], bins fastest (coalesced 8-byte accesses).
#include <cstdlib>

#include "common.cuh"

namespace aefft {

namespace {

__device__ __forceinline__ void cmac(float2& acc, float2 a, float2 b) {  // acc += a*b
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cmac_conj(float2& acc, float2 a, float2 b) {  // acc += a*conj(b)
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.y, b.x, acc.y); acc.y = fmaf(-a.x, b.y, acc.y);
}

struct SmallParams {
  const float2 *X, *Xt, *O;  // [B][dD][S]; O == nullptr: recompute O = conv(conv(X; C, b); F, p)
  const float2 *C, *F;       // [dM][dD][S], [dD][dM][S]
  const float *bias_b, *bias_p;  // applied at bin 0 (nullptr on devices that do not own the DC column)
  float2 *dC, *dF;           // [dM][dD][S], [dD][dM][S]
  float *db, *dp;            // written by the threads of bin 0 (when bias_b != nullptr, i.e. the DC owner)
  double* part;              // mse kernel: per-block partial sums
  long long S;
  int B, dM, ncols, col0, Ny;
  float norm, gscale, dbscale;
};

template <int DD, int LG, bool HAS_O>
__global__ void __launch_bounds__(128) small_grad_kernel(SmallParams p) {
  // LG = lanes per bin = dM / 4
  const int tid = threadIdx.x;
  const int mq = tid % LG;
  const long long w = (long long)blockIdx.x * (128 / LG) + tid / LG;
  const bool live = w < p.S;
  const long long wc = live ? w : 0;
  const int dM = p.dM;
  const float inv_dM = 1.f / (float)dM, inv_dD = 1.f / (float)DD;
  float2 Cq[4][DD], Fq[DD][4], dCq[4][DD], dFq[DD][4];
  float bb[4], bp[DD];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int m = 4 * mq + a;
#pragma unroll
    for (int d = 0; d < DD; d++) {
      Cq[a][d] = __ldg(p.C + ((long long)m * DD + d) * p.S + wc);
      Fq[d][a] = __ldg(p.F + ((long long)d * dM + m) * p.S + wc);
      dCq[a][d] = make_float2(0.f, 0.f);
      dFq[d][a] = make_float2(0.f, 0.f);
    }
    bb[a] = (wc == 0 && p.bias_b) ? p.bias_b[m] * p.norm : 0.f;
  }
#pragma unroll
  for (int d = 0; d < DD; d++) bp[d] = (wc == 0 && p.bias_p) ? p.bias_p[d] * p.norm : 0.f;
  float gsum[4] = {0.f, 0.f, 0.f, 0.f}, esum[DD];
#pragma unroll
  for (int d = 0; d < DD; d++) esum[d] = 0.f;
  const long long fs = (long long)DD * p.S;
  // Operands of the frame PF steps ahead are pulled into L2 (no registers held): with 161 registers per thread 12 warps fit on
  // an SM and more than half of them waited on DRAM at any time; a register double buffer of the next frame was slower
  // (1.19 vs 1.06 ms at 3 -> 16 channels, 512 x 257 bins, 128 frames), the L2 prefetch gives 0.91 ms.
  constexpr int PF = 3;
  for (int b = 0; b < p.B; b++) {
    float2 x[DD], e[DD], hh[4];
    if (mq == 0 && b + PF < p.B) {
#pragma unroll
      for (int d = 0; d < DD; d++) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.X + (b + PF) * fs + d * p.S + wc));
        if (HAS_O) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.O + (b + PF) * fs + d * p.S + wc));
        if (p.Xt != p.X) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.Xt + (b + PF) * fs + d * p.S + wc));
      }
    }
#pragma unroll
    for (int d = 0; d < DD; d++) x[d] = __ldg(p.X + b * fs + d * p.S + wc);
    // H-hat[m] = sum_d C[m][d] X[d] + b[m] Nx Ny at DC  (no /dM: quirk F1);  H = (H-hat - bias)/dM + bias
    if (!HAS_O) {
#pragma unroll
      for (int d = 0; d < DD; d++) e[d] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
      float2 s = make_float2(0.f, 0.f);
#pragma unroll
      for (int d = 0; d < DD; d++) cmac(s, Cq[a][d], x[d]);
      hh[a] = make_float2(s.x + bb[a], s.y);
      if (!HAS_O) {  // this lane's share of the decoder output O = sum_m F[d][m] H[m] (reduced over the lane group below)
        const float2 h = make_float2(fmaf(s.x, inv_dM, bb[a]), s.y * inv_dM);
#pragma unroll
        for (int d = 0; d < DD; d++) cmac(e[d], Fq[d][a], h);
      }
    }
    if (HAS_O) {
#pragma unroll
      for (int d = 0; d < DD; d++) {
        const float2 o = __ldg(p.O + b * fs + d * p.S + wc), t = __ldg(p.Xt + b * fs + d * p.S + wc);
        e[d] = make_float2(o.x - t.x, o.y - t.y);
      }
    } else {
#pragma unroll
      for (int d = 0; d < DD; d++) {
#pragma unroll
        for (int o = 1; o < LG; o <<= 1) {
          e[d].x += __shfl_xor_sync(0xffffffffu, e[d].x, o);
          e[d].y += __shfl_xor_sync(0xffffffffu, e[d].y, o);
        }
        const float2 t = __ldg(p.Xt + b * fs + d * p.S + wc);
        e[d] = make_float2(fmaf(e[d].x, inv_dD, bp[d]) - t.x, e[d].y * inv_dD - t.y);
      }
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
      float2 g = make_float2(0.f, 0.f);
#pragma unroll
      for (int d = 0; d < DD; d++) cmac_conj(g, e[d], Fq[d][a]);
      gsum[a] += g.x;
#pragma unroll
      for (int d = 0; d < DD; d++) {
        cmac_conj(dCq[a][d], g, x[d]);
        cmac_conj(dFq[d][a], e[d], hh[a]);
      }
    }
#pragma unroll
    for (int d = 0; d < DD; d++) esum[d] += e[d].x;
  }
  if (!live) return;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int m = 4 * mq + a;
#pragma unroll
    for (int d = 0; d < DD; d++) {
      p.dC[((long long)m * DD + d) * p.S + w] = make_float2(dCq[a][d].x * p.gscale, dCq[a][d].y * p.gscale);
      p.dF[((long long)d * dM + m) * p.S + w] = make_float2(dFq[d][a].x * p.gscale, dFq[d][a].y * p.gscale);
    }
  }
  if (w == 0 && p.db) {
#pragma unroll
    for (int a = 0; a < 4; a++) p.db[4 * mq + a] = gsum[a] * p.dbscale;
    if (mq == 0) {
#pragma unroll
      for (int d = 0; d < DD; d++) p.dp[d] = esum[d] * p.dbscale;
    }
  }
}

// re-forward + mse: sum over bins and frames of hw(bin) |conv(conv(X; C, b); F, p) - Xt|^2
template <int DD, int LG>
__global__ void __launch_bounds__(128) small_mse_kernel(SmallParams p) {
  const int tid = threadIdx.x;
  const int mq = tid % LG;
  const long long w = (long long)blockIdx.x * (128 / LG) + tid / LG;
  const bool live = w < p.S;
  const long long wc = live ? w : 0;
  const int dM = p.dM;
  const float inv_dM = 1.f / (float)dM, inv_dD = 1.f / (float)DD;
  float2 Cq[4][DD], Fq[DD][4];
  float bb[4], bp[DD];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int m = 4 * mq + a;
#pragma unroll
    for (int d = 0; d < DD; d++) {
      Cq[a][d] = __ldg(p.C + ((long long)m * DD + d) * p.S + wc);
      Fq[d][a] = __ldg(p.F + ((long long)d * dM + m) * p.S + wc);
    }
    bb[a] = (wc == 0 && p.bias_b) ? p.bias_b[m] * p.norm : 0.f;
  }
#pragma unroll
  for (int d = 0; d < DD; d++) bp[d] = (wc == 0 && p.bias_p) ? p.bias_p[d] * p.norm : 0.f;
  const long long fs = (long long)DD * p.S;
  float acc = 0.f;
  double tot = 0.0;
  for (int b = 0; b < p.B; b++) {
    float2 x[DD], o[DD];
#pragma unroll
    for (int d = 0; d < DD; d++) { x[d] = __ldg(p.X + b * fs + d * p.S + wc); o[d] = make_float2(0.f, 0.f); }
#pragma unroll
    for (int a = 0; a < 4; a++) {
      float2 s = make_float2(0.f, 0.f);
#pragma unroll
      for (int d = 0; d < DD; d++) cmac(s, Cq[a][d], x[d]);
      s.x = fmaf(s.x, inv_dM, bb[a]);
      s.y *= inv_dM;
#pragma unroll
      for (int d = 0; d < DD; d++) cmac(o[d], Fq[d][a], s);
    }
#pragma unroll
    for (int d = 0; d < DD; d++) {
#pragma unroll
      for (int k = 1; k < LG; k <<= 1) {
        o[d].x += __shfl_xor_sync(0xffffffffu, o[d].x, k);
        o[d].y += __shfl_xor_sync(0xffffffffu, o[d].y, k);
      }
      const float2 t = __ldg(p.Xt + b * fs + d * p.S + wc);
      const float ex = fmaf(o[d].x, inv_dD, bp[d]) - t.x, ey = o[d].y * inv_dD - t.y;
      acc = fmaf(ex, ex, fmaf(ey, ey, acc));
    }
    if ((b & 7) == 7) { tot += (double)acc; acc = 0.f; }
  }
  tot += (double)acc;
  const int wy = p.col0 + (int)(wc % p.ncols);
  const double hw = (wy == 0 || wy == p.Ny / 2) ? 1.0 : 2.0;
  if (!live || mq != 0) tot = 0.0;
  tot *= hw;
  __shared__ double red[128];
  red[tid] = tot;
  __syncthreads();
  for (int h = 64; h > 0; h >>= 1) {
    if (tid < h) red[tid] += red[tid + h];
    __syncthreads();
  }
  if (tid == 0) p.part[blockIdx.x] = red[0];
}

__global__ void small_final_kernel(const double* __restrict__ part, long long n, double scale, float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)(red[0] * scale);
}

template <int DD, int LG>
int run_grad(aefft_ctx* ctx, const SmallParams& p) {
  const long long blocks = (p.S + 128 / LG - 1) / (128 / LG);
  if (p.O) small_grad_kernel<DD, LG, true><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p);
  else small_grad_kernel<DD, LG, false><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p);
  return AEFFT_OK;
}
template <int DD, int LG>
int run_mse(aefft_ctx* ctx, const SmallParams& p) {
  const long long blocks = (p.S + 128 / LG - 1) / (128 / LG);
  small_mse_kernel<DD, LG><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p);
  return AEFFT_OK;
}

#define AEFFT_SMALL_DISPATCH(FN, dD, lg, ...)                                                   \
  do {                                                                                          \
    if (dD == 1) { SMALL_LG(FN, 1, lg, __VA_ARGS__); }                                          \
    else if (dD == 2) { SMALL_LG(FN, 2, lg, __VA_ARGS__); }                                     \
    else if (dD == 3) { SMALL_LG(FN, 3, lg, __VA_ARGS__); }                                     \
    else { SMALL_LG(FN, 4, lg, __VA_ARGS__); }                                                  \
  } while (0)
#define SMALL_LG(FN, DD, lg, ...)                                                               \
  switch (lg) {                                                                                 \
    case 1: FN<DD, 1>(__VA_ARGS__); break;                                                      \
    case 2: FN<DD, 2>(__VA_ARGS__); break;                                                      \
    case 4: FN<DD, 4>(__VA_ARGS__); break;                                                      \
    case 8: FN<DD, 8>(__VA_ARGS__); break;                                                      \
    default: FN<DD, 16>(__VA_ARGS__); break;                                                    \
  }

}  // namespace

// ---- forward contraction (conv_k, :162-189) when CI*CO is small: the CO x CI weight block of a bin lives in registers, split
// over PARTS = 4 lanes of a warp (lane = part * 8 + bin): the lanes of a bin share the LARGER channel count -- each takes a
// quarter of the inputs (partial sums combined by two shuffles) when CI >= CO, a quarter of the outputs otherwise.  A thread
// walks over a chunk of frames, two frames per iteration with all loads issued before the arithmetic.
//   out[b][o][w] = in_scale * sum_c W[o][c][w] in[b][c][w]  (+ bias[o]*bias_scale at bin 0)
// One thread per bin with the whole block in registers (round 2's first form: 96 + 32 registers of operands at 16 x 3) ran
// 11-15 warps per SM and waited on its own loads: 2.7 TB/s at 16 -> 3 channels (ncu: issue slots 14 % busy, 27 cycles of
// long-scoreboard stall per issued instruction).
// Fused with the spectral pooling next to it (resize :87-157), MODE != 0: the threads walk over the bins of the SMALL grid
// (Nxm x Nyrm); wb is the same frequency on the BIG grid (rows i < Nxm/2 -> i, Nxm/2 -> Nxb/2, above -> i + Nxb - Nxm; columns
// j < Nyrm-1 -> j, Nyrm-1 -> Nyrb-1).  The kernel spectrum W always lives on the big grid (the conv runs at that resolution).
//   MODE 1 (conv, then pooling by cropping): in at wb, out at the small bin -- the 3/4 of the conv output that the crop
//           discards is never computed;
//   MODE 2 (up-sampling by zero embedding, then conv): in at the small bin, out at wb -- the caller zeroes the output first;
//           the conv of the zero band is zero (the bias lives on the DC bin, which is always kept);
//   MODE 3 (the same, output kept COMPACT on the small grid: the decoder's spectra stay on the support of the innermost level).
struct ConvRegMap {
  int mode, Nxm, Nyrm, Nxb, Nyrb;
};
// element (frame b, channel c, bin w) of a spectrum at b * sb + c * sc + w * sw (in float2):
// bins-fastest [b][c][w]: (C S, S, 1); bin-major [w][b][c]: (C, 1, B C) -- what a tensor-core level next door keeps
struct ConvRegLayout {
  long long sb, sc, sw;
};
namespace {
template <int CI, int CO>
__global__ void __launch_bounds__(128) conv_reg_kernel(const float2* __restrict__ in, const float2* __restrict__ W,
                                                       const float* __restrict__ bias, float2* __restrict__ out, long long S,
                                                       ConvRegLayout in_l, long long S_w, ConvRegLayout out_l, ConvRegMap map,
                                                       int B, int frames_per_block, float in_scale, float bias_scale) {
  constexpr int PARTS = 4, BPW = 32 / PARTS;       // bins per warp
  constexpr bool SPLIT_IN = CI >= CO;
  constexpr int CIP = SPLIT_IN ? (CI + PARTS - 1) / PARTS : CI;  // inputs of this lane
  constexpr int COP = SPLIT_IN ? CO : (CO + PARTS - 1) / PARTS;  // outputs of this lane
  const int lane = threadIdx.x & 31, part = lane / BPW;
  const long long w = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * BPW + (lane % BPW);
  const bool live = w < S;
  const long long wl = live ? w : 0;
  long long w_in = wl, w_w = wl, w_out = wl;
  if (map.mode != 0) {
    const int i = (int)(wl / map.Nyrm), j = (int)(wl - (long long)i * map.Nyrm);
    const int ib = i < map.Nxm / 2 ? i : (i == map.Nxm / 2 ? map.Nxb / 2 : i + map.Nxb - map.Nxm);
    const int jb = j < map.Nyrm - 1 ? j : map.Nyrb - 1;
    const long long wb = (long long)ib * map.Nyrb + jb;
    w_w = wb;
    if (map.mode == 1) w_in = wb; else if (map.mode == 2) w_out = wb;  // mode 3: in and out both on the small grid
  }
  const int c0 = SPLIT_IN ? part * CIP : 0, o0 = SPLIT_IN ? 0 : part * COP;
  float2 Wr[COP][CIP];
  float bo[COP];
#pragma unroll
  for (int o = 0; o < COP; o++) {
#pragma unroll
    for (int c = 0; c < CIP; c++) {
      const bool ok = o0 + o < CO && c0 + c < CI;
      const float2 v = ok ? __ldg(W + ((long long)(o0 + o) * CI + (c0 + c)) * S_w + w_w) : make_float2(0.f, 0.f);
      Wr[o][c] = make_float2(v.x * in_scale, v.y * in_scale);
    }
    bo[o] = (w == 0 && bias && o0 + o < CO && (!SPLIT_IN || part == 0)) ? bias[o0 + o] * bias_scale : 0.f;
  }
  const int b0 = blockIdx.y * frames_per_block, b1 = min(B, b0 + frames_per_block);
  for (int b = b0; b < b1; b += 2) {
    const bool two = b + 1 < b1;
    float2 x[2][CIP];
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
      for (int c = 0; c < CIP; c++)
        x[u][c] = (c0 + c < CI && (u == 0 || two)) ? __ldg(in + (b + u) * in_l.sb + (c0 + c) * in_l.sc + w_in * in_l.sw) : make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 2; u++) {
      float2 acc[COP];
#pragma unroll
      for (int o = 0; o < COP; o++) {
        acc[o] = make_float2(bo[o], 0.f);
#pragma unroll
        for (int c = 0; c < CIP; c++) cmac(acc[o], Wr[o][c], x[u][c]);
      }
      if (SPLIT_IN) {
        // sum over the four parts (fixed order: the xor tree), then part p stores the outputs o = p, p + 4, ...
#pragma unroll
        for (int o = 0; o < COP; o++) {
          acc[o].x += __shfl_xor_sync(0xffffffffu, acc[o].x, BPW);
          acc[o].y += __shfl_xor_sync(0xffffffffu, acc[o].y, BPW);
          acc[o].x += __shfl_xor_sync(0xffffffffu, acc[o].x, 2 * BPW);
          acc[o].y += __shfl_xor_sync(0xffffffffu, acc[o].y, 2 * BPW);
        }
#pragma unroll
        for (int o = 0; o < COP; o++)
          if (live && (o % PARTS) == part && (u == 0 || two)) out[(b + u) * out_l.sb + o * out_l.sc + w_out * out_l.sw] = acc[o];
      } else if (COP % 2 == 0 && CO % COP == 0 && out_l.sc == 1) {
        // channel-contiguous output (bin-major): this lane's COP outputs are one 8 * COP byte run: 16-byte stores
        if (live && (u == 0 || two)) {
          float4* o4 = reinterpret_cast<float4*>(out + (b + u) * out_l.sb + o0 + w_out * out_l.sw);
#pragma unroll
          for (int o = 0; o < COP; o += 2) o4[o / 2] = make_float4(acc[o].x, acc[o].y, acc[o + 1].x, acc[o + 1].y);
        }
      } else {
#pragma unroll
        for (int o = 0; o < COP; o++)
          if (live && o0 + o < CO && (u == 0 || two)) out[(b + u) * out_l.sb + (o0 + o) * out_l.sc + w_out * out_l.sw] = acc[o];
      }
    }
  }
}

int conv_reg_launch(aefft_ctx* ctx, int64_t B, int CI, int CO, int64_t S, ConvRegLayout in_l, int64_t S_w, ConvRegLayout out_l,
                    const ConvRegMap& map, const float2* in, const float2* W, const float* bias, float bias_scale, float in_scale,
                    float2* out, const char* name) {
  const int fpb = 16;
  dim3 grid((unsigned)((S + 31) / 32), (unsigned)((B + fpb - 1) / fpb));  // 128 threads = 4 warps x 8 bins
  if (grid.y > 65535) return AEFFT_ERR_UNSUPPORTED;
#define AEFFT_CONV_REG(ci, co)                                                                                               \
  if (CI == ci && CO == co) {                                                                                                \
    ProfScope prof(ctx, name, 8.0 * B * CI * CO * S, 8.0 * S * ((double)B * (CI + CO) + (double)CI * CO));                   \
    conv_reg_kernel<ci, co><<<grid, 128, 0, ctx->stream>>>(in, W, bias, out, S, in_l, S_w, out_l, map, (int)B, fpb, in_scale,      \
                                                           bias_scale);                                                      \
    ctx->launches++;                                                                                                         \
    AE_CUDA(cudaGetLastError());                                                                                             \
    return AEFFT_OK;                                                                                                         \
  }
  AEFFT_CONV_REG(3, 16) AEFFT_CONV_REG(16, 3) AEFFT_CONV_REG(3, 8) AEFFT_CONV_REG(8, 3) AEFFT_CONV_REG(1, 8) AEFFT_CONV_REG(8, 1)
  AEFFT_CONV_REG(3, 4) AEFFT_CONV_REG(4, 3) AEFFT_CONV_REG(1, 16) AEFFT_CONV_REG(16, 1)
#undef AEFFT_CONV_REG
  return AEFFT_ERR_UNSUPPORTED;
}
}  // namespace

bool spec_conv_reg_supported(int CI, int CO) {
  if (getenv("AEFFT_NO_SPEC_SMALL")) return false;
  const int lo = CI < CO ? CI : CO, hi = CI < CO ? CO : CI;
  return (lo == 3 && (hi == 16 || hi == 8 || hi == 4)) || (lo == 1 && (hi == 16 || hi == 8));
}

int launch_spec_conv_reg(aefft_ctx* ctx, int64_t B, int CI, int CO, int64_t S, const float2* in, const float2* W, const float* bias,
                         float bias_scale, float in_scale, float2* out) {
  if (!spec_conv_reg_supported(CI, CO)) return AEFFT_ERR_UNSUPPORTED;
  return conv_reg_launch(ctx, B, CI, CO, S, ConvRegLayout{CI * S, S, 1}, S, ConvRegLayout{CO * S, S, 1}, ConvRegMap{0, 0, 0, 0, 0},
                         in, W, bias, bias_scale, in_scale, out, "spec_contract_reg");
}

// conv_k at resolution (Nxb, Nyb) followed by the spectral pooling to (Nxm, Nym) [pooled_out], or preceded by the spectral
// up-sampling from (Nxm, Nym) [!pooled_out]; W = kernel spectrum at (Nxb, Nyb).  The big-resolution spectrum is bins-fastest;
// small_bin_major: the small one is bin-major [bin][frame][channel] (the level next door runs on the tensor cores).  The
// up-sampling form zeroes `out` itself.
int launch_spec_conv_reg_resized(aefft_ctx* ctx, int64_t B, int CI, int CO, int Nxb, int Nyb, int Nxm, int Nym, bool pooled_out,
                                 bool small_bin_major, const float2* in, const float2* W, const float* bias, float bias_scale,
                                 float in_scale, float2* out) {
  if (!spec_conv_reg_supported(CI, CO)) return AEFFT_ERR_UNSUPPORTED;
  AE_ARG(Nxm < Nxb && Nym < Nyb && Nxm >= 2 && Nym >= 2);
  const int64_t Sb = (int64_t)Nxb * (Nyb / 2 + 1), Sm = (int64_t)Nxm * (Nym / 2 + 1);
  const ConvRegMap map{pooled_out ? 1 : 2, Nxm, Nym / 2 + 1, Nxb, Nyb / 2 + 1};
  const int Csmall = pooled_out ? CO : CI;
  const ConvRegLayout small = small_bin_major ? ConvRegLayout{Csmall, 1, B * Csmall} : ConvRegLayout{Csmall * Sm, Sm, 1};
  const ConvRegLayout big{(pooled_out ? CI : CO) * Sb, Sb, 1};
  if (!pooled_out) AE_CUDA(cudaMemsetAsync(out, 0, (size_t)B * CO * Sb * sizeof(float2), ctx->stream));
  return conv_reg_launch(ctx, B, CI, CO, Sm, pooled_out ? big : small, Sb, pooled_out ? small : big, map, in, W, bias, bias_scale,
                         in_scale, out, pooled_out ? "spec_contract_reg_pool" : "spec_contract_reg_embed");
}

// conv_k at resolution (Nxb, Nyb) of a spectrum that is non-zero only on the bins an up-sampling from (Nxm, Nym) fills: in and
// out are COMPACT on that (Nxm, Nym) grid (bin-major [bin][frame][ch] or bins-fastest [frame][ch][bins] each)
int launch_spec_conv_reg_support(aefft_ctx* ctx, int64_t B, int CI, int CO, int Nxb, int Nyb, int Nxm, int Nym, bool in_bin_major,
                                 bool out_bin_major, const float2* in, const float2* W, const float* bias, float bias_scale,
                                 float in_scale, float2* out) {
  if (!spec_conv_reg_supported(CI, CO)) return AEFFT_ERR_UNSUPPORTED;
  AE_ARG(Nxm < Nxb && Nym < Nyb && Nxm >= 2 && Nym >= 2);
  const int64_t Sb = (int64_t)Nxb * (Nyb / 2 + 1), Sm = (int64_t)Nxm * (Nym / 2 + 1);
  const ConvRegMap map{3, Nxm, Nym / 2 + 1, Nxb, Nyb / 2 + 1};
  const ConvRegLayout li = in_bin_major ? ConvRegLayout{CI, 1, B * CI} : ConvRegLayout{CI * Sm, Sm, 1};
  const ConvRegLayout lo = out_bin_major ? ConvRegLayout{CO, 1, B * CO} : ConvRegLayout{CO * Sm, Sm, 1};
  return conv_reg_launch(ctx, B, CI, CO, Sm, li, Sb, lo, map, in, W, bias, bias_scale, in_scale, out, "spec_contract_reg_support");
}

bool spec_small_eligible(int dD, int dM) {
  if (getenv("AEFFT_NO_SPEC_SMALL")) return false;
  const int lg = dM / 4;
  return dD >= 1 && dD <= 4 && dM % 4 == 0 && (lg == 1 || lg == 2 || lg == 4 || lg == 8 || lg == 16);
}

// gradient spectra dC [dM][dD][S], dF [dD][dM][S] (scaled by gscale) and the DC-bin bias gradients of one iteration
int launch_small_grad(aefft_ctx* ctx, int64_t B, int dD, int dM, int64_t S, const float2* X, const float2* Xt, const float2* O,
                      const float2* C, const float2* F, const float* bias_b, const float* bias_p, float norm, float gscale,
                      float dbscale, float2* dC, float2* dF, float* db, float* dp) {
  AE_ARG(spec_small_eligible(dD, dM));
  SmallParams p{X, Xt, O, C, F, bias_b, bias_p, dC, dF, bias_b ? db : nullptr, dp, nullptr, S, (int)B, dM, 1, 0, 2, norm, gscale,
                dbscale};
  const double px = (double)B * S;
  ProfScope prof(ctx, "spec_small_grad", 8.0 * px * dM * dD * (O ? 4 : 5), 8.0 * (px * dD * (O ? (Xt == X ? 2 : 3) : (Xt == X ? 1 : 2)) + 4.0 * S * dM * dD));
  AEFFT_SMALL_DISPATCH(run_grad, dD, dM / 4, ctx, p);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// *mse_out = mse_scale * sum hw |conv(conv(X; C, b); F, p) - Xt|^2 over the owned bins and the frames
int launch_small_mse(aefft_ctx* ctx, int64_t B, int dD, int dM, int64_t S, const float2* X, const float2* Xt, const float2* C,
                     const float2* F, const float* bias_b, const float* bias_p, float norm, float* mse_out, double mse_scale,
                     int ncols, int col0, int Ny) {
  AE_ARG(spec_small_eligible(dD, dM));
  const int lg = dM / 4;
  const long long blocks = (S + 128 / lg - 1) / (128 / lg);
  double* part;
  AE_TRY(ctx->getT("small_part", (size_t)blocks, &part));
  SmallParams p{X, Xt, nullptr, C, F, bias_b, bias_p, nullptr, nullptr, nullptr, nullptr, part, S, (int)B, dM,
                ncols > 0 ? ncols : Ny / 2 + 1, ncols > 0 ? col0 : 0, Ny, norm, 0.f, 0.f};
  const double px = (double)B * S;
  {
    ProfScope prof(ctx, "spec_small_mse", 8.0 * px * dM * dD * 2, 8.0 * (px * dD * (Xt == X ? 1 : 2) + 2.0 * S * dM * dD));
    AEFFT_SMALL_DISPATCH(run_mse, dD, lg, ctx, p);
    ctx->launches++;
  }
  small_final_kernel<<<1, 256, 0, ctx->stream>>>(part, blocks, mse_scale, mse_out);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

}  // namespace aefft
