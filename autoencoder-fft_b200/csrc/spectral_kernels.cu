// Momentum-space (per-frequency-bin) kernels: spectral pooling, the per-bin channel contractions of the forward and of
// the gradients, kernel pad/shrink, the kernel-space clipped-momentum update with the multiobjective term, and the
// Hermitian-weighted MSE.  Half spectra are [frame][channel][Nx][Ny/2+1] complex64, bins (w) fastest, so every kernel
// below maps threads to consecutive bins (coalesced 8-byte accesses) and keeps a small channel x frame register tile.
//
// Reference kernels replaced (fft_backproplib.cu): resize :87-157, conv_k :162-189, gradient_k_io :395-475,
// calc_mse + thrust::reduce :480-498/1178-1192, shrink_k :535-565, pad_k :570-600, backprop_d :605-652,
// backprop_double :657-704, gradient_diff :709-753.
#include <cstdlib>

#include "common.cuh"

namespace aefft {

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ void cfma(float2& acc, float2 a, float2 b) {  // acc += a*b
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfma_conj(float2& acc, float2 a, float2 b) {  // acc += a*conj(b)
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.y, b.x, acc.y); acc.y = fmaf(-a.x, b.y, acc.y);
}

// ------------------------------------------------------------------------------------------------ spectral pooling
// resize (:87-157): scale>1 crops the half spectrum around zero frequency, scale<0 embeds it into a zeroed larger
// one; the Nyquist column/row of the source lands on the Nyquist of the target; no amplitude rescale.
// One warp-pair (64 threads) per target row, 4 rows per CTA; the row / plane decomposition is done once per row (no
// per-element 64-bit division) and every thread keeps several independent 8-byte loads in flight.
constexpr int RSZ_ROWS = 4;
__global__ void __launch_bounds__(64 * RSZ_ROWS) spec_resize_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                                    long long rows, int Nx, int Ny, int Nxs, int Nys) {
  const int Nyr = Ny / 2 + 1, Nyrs = Nys / 2 + 1;
  const long long row = (long long)blockIdx.x * RSZ_ROWS + (threadIdx.x >> 6);
  if (row >= rows) return;
  const int lane = threadIdx.x & 63;
  const long long d = row / Nxs;
  const int i = (int)(row - d * Nxs);
  float2* dst = out + row * Nyrs;
  int si = -1;
  if (Nxs <= Nx) {
    si = i < Nxs / 2 ? i : (i == Nxs / 2 ? Nx / 2 : i + Nx - Nxs);
  } else {
    if (i < Nx / 2) si = i;
    else if (i > Nxs - Nx / 2) si = i - Nxs + Nx;
    else if (i == Nxs / 2) si = Nx / 2;
  }
  if (si < 0) {  // a row of the zero band of an embedded spectrum
    for (int j = lane; j < Nyrs; j += 64) dst[j] = make_float2(0.f, 0.f);
    return;
  }
  const float2* src = in + (d * Nx + si) * (long long)Nyr;
  // columns: cropping keeps j < Nyrs-1 and puts the source Nyquist column on the target's; embedding copies j < Nyr-1,
  // zero-fills, and puts the source Nyquist on the target's (the reference tests j<Nyr-1 first; j==Nyrs-1 never
  // satisfies it when upsampling)
  const int ncopy = (Nxs <= Nx ? Nyrs : Nyr) - 1;
#pragma unroll 4
  for (int j = lane; j < Nyrs - 1; j += 64) dst[j] = j < ncopy ? __ldg(src + j) : make_float2(0.f, 0.f);
  if (lane == 0) dst[Nyrs - 1] = __ldg(src + Nyr - 1);
}

int launch_spec_resize(aefft_ctx* ctx, int64_t planes, int Nx, int Ny, int Nxs, int Nys, const float2* in, float2* out) {
  const long long total = (long long)planes * Nxs * (Nys / 2 + 1);
  const long long rows = (long long)planes * Nxs;
  if ((rows + RSZ_ROWS - 1) / RSZ_ROWS > 0x7fffffffLL) return AEFFT_ERR_UNSUPPORTED;
  ProfScope prof(ctx, "spec_resize", 0.0, 8.0 * (total + (double)planes * (Nxs <= Nx ? Nxs : Nx) * (Nys / 2 + 1)));
  spec_resize_kernel<<<(unsigned)((rows + RSZ_ROWS - 1) / RSZ_ROWS), 64 * RSZ_ROWS, 0, ctx->stream>>>(in, out, rows, Nx, Ny, Nxs,
                                                                                                  Nys);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// ------------------------------------------------------------------------------------------------ per-bin contraction
// out[b][o][w] = sum_c Wt(o,c)[w] * (in_scale * (in0[b][c][w] - in1[b][c][w])) + [w==0] bias[o]*bias_scale
//   Wt(o,c) = W[o*w_so + c*w_sc + w], conjugated when conjW.
// conv_k (:162-189):           in_scale = 1/dM, W = C[m][d], bias b*Nx*Ny
// G of gradient_k_io (:410-419): in = O - Xt, W = conj(F[d1][m])
constexpr int SC_OT = 4, SC_FT = 4, SC_THREADS = 128;

struct ContractParams {
  const float2* in0;
  const float2* in1;
  const float2* W;
  const float* bias;
  float2* out;
  long long w_so, w_sc, S;
  int B, C, O, conjW;
  float in_scale, bias_scale;
};

template <int SC_OT, int SC_FT>
__global__ void __launch_bounds__(SC_THREADS) spec_contract_kernel(ContractParams p) {
  const long long w = (long long)blockIdx.x * SC_THREADS + threadIdx.x;
  if (w >= p.S) return;
  const int o0 = blockIdx.y * SC_OT, b0 = blockIdx.z * SC_FT;
  float2 acc[SC_OT][SC_FT];
#pragma unroll
  for (int o = 0; o < SC_OT; o++)
#pragma unroll
    for (int f = 0; f < SC_FT; f++) acc[o][f] = make_float2(0.f, 0.f);
  for (int c = 0; c < p.C; c++) {
    float2 x[SC_FT], wv[SC_OT];
#pragma unroll
    for (int f = 0; f < SC_FT; f++) {
      x[f] = make_float2(0.f, 0.f);
      if (b0 + f < p.B) {
        const long long off = ((long long)(b0 + f) * p.C + c) * p.S + w;
        float2 v = p.in0[off];
        if (p.in1) { float2 u = p.in1[off]; v.x -= u.x; v.y -= u.y; }
        x[f] = make_float2(v.x * p.in_scale, v.y * p.in_scale);
      }
    }
#pragma unroll
    for (int o = 0; o < SC_OT; o++) {
      wv[o] = make_float2(0.f, 0.f);
      if (o0 + o < p.O) {
        wv[o] = __ldg(p.W + (o0 + o) * p.w_so + c * p.w_sc + w);
        if (p.conjW) wv[o].y = -wv[o].y;
      }
    }
#pragma unroll
    for (int o = 0; o < SC_OT; o++)
#pragma unroll
      for (int f = 0; f < SC_FT; f++) cfma(acc[o][f], x[f], wv[o]);
  }
#pragma unroll
  for (int o = 0; o < SC_OT; o++) {
    if (o0 + o >= p.O) continue;
    const float bv = (w == 0 && p.bias) ? p.bias[o0 + o] * p.bias_scale : 0.f;
#pragma unroll
    for (int f = 0; f < SC_FT; f++) {
      if (b0 + f >= p.B) continue;
      float2 v = acc[o][f];
      v.x += bv;
      p.out[((long long)(b0 + f) * p.O + o0 + o) * p.S + w] = v;
    }
  }
}


// ------------------------------------------------------------------------------------------------ shared-memory tiled
// Generic per-bin contraction  Out[i][j][w] = out_scale * sum_r P(i,r)[w] * Q(j,r)[w]  (+ out_bias[i] at w == 0)
//   P(i,r) = p_scale * (P0 - P1)[i*psi + r*psr + w]            (conjugated when conjP)
//   Q(j,r) = q_scale * (Q0 - Q1)[j*qsi + r*qsr + w] + [w==0] q_bias[j]   (conjugated when conjQ)
// used for both the channel contraction (i = output channel, j = frame, r = input channel) and the frame-reduced
// outer products of the gradients (i, j = channels, r = frame).  The 4x4-register-tile kernels above re-read every
// operand O/4 resp. B/4 times through L2 (8.7 GB per launch at 32->64 channels, 128 frames): they are L2-bandwidth
// bound at ~1.1 TB/s algorithmic.  Here a CTA owns 32 consecutive bins x 16 i x 16 j, stages 4 reduction steps of both
// operands in shared memory with coalesced 256-byte rows, and every thread keeps a 4 (i) x 8 (j) complex register
// tile for its bin: operands are re-read 8x less often and the inner loop is 12 shared loads per 128 FMAs.
constexpr int ST_BINS = 32, ST_I = 16, ST_J = 16, ST_R = 4, ST_THREADS = 256;

struct TiledParams {
  const float2 *P0, *P1, *Q0, *Q1;
  const float *q_bias, *out_bias;
  float2* out;
  long long psi, psr, qsi, qsr, osi, osj, S;
  int nI, nJ, nR, conjP, conjQ;
  float p_scale, q_scale, out_scale, out_bias_scale, q_bias_scale;
};

__global__ void __launch_bounds__(ST_THREADS, 2) spec_tiled_kernel(TiledParams p) {
  __shared__ float2 Ps[ST_R][ST_I][ST_BINS];
  __shared__ float2 Qs[ST_R][ST_J][ST_BINS];
  const int tid = threadIdx.x, bin = tid & 31, g = tid >> 5;
  const int ig = g & 3, jg = g >> 2;  // 4 i-groups of 4, 2 j-groups of 8
  // grid: x = (i tile, j tile) fastest, y = bin tile -> the CTAs that share a bin tile run together and its operand rows
  // are fetched from DRAM once (ncu: 3.2x the algorithmic reads with bins fastest)
  const int n_it = (p.nI + ST_I - 1) / ST_I;
  const long long w0 = (long long)blockIdx.y * ST_BINS;
  const int i0 = (blockIdx.x % n_it) * ST_I, j0 = (blockIdx.x / n_it) * ST_J;
  float2 acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 8; b++) acc[a][b] = make_float2(0.f, 0.f);
  for (int r0 = 0; r0 < p.nR; r0 += ST_R) {
    // stage ST_R*(ST_I+ST_J) = 128 rows of 32 bins, one warp per row, 16 rows per warp: all 16 (32 with a second source)
    // global loads of a thread are issued before the first use (the stage is latency bound otherwise)
    {
      const long long w = w0 + bin;
      float2 v[16];
#pragma unroll
      for (int t = 0; t < 16; t++) {
        const int row = g + 8 * t;              // t < 8: P rows, t >= 8: Q rows (ST_R * ST_I == 64)
        const bool isP = t < 8;
        const int rl = isP ? row : row - ST_R * ST_I;
        const int rr = rl / ST_I, k = rl - rr * ST_I;   // ST_I == ST_J
        const int r = r0 + rr;
        const int lim = isP ? p.nI : p.nJ, base = isP ? i0 : j0;
        const bool ok = r < p.nR && base + k < lim && w < p.S;
        const long long off = isP ? (long long)(i0 + k) * p.psi + (long long)r * p.psr + w
                                  : (long long)(j0 + k) * p.qsi + (long long)r * p.qsr + w;
        const float2* s0 = isP ? p.P0 : p.Q0;
        v[t] = ok ? __ldg(s0 + off) : make_float2(0.f, 0.f);
      }
      if (p.P1 || p.Q1) {
#pragma unroll
        for (int t = 0; t < 16; t++) {
          const int row = g + 8 * t;
          const bool isP = t < 8;
          const int rl = isP ? row : row - ST_R * ST_I;
          const int rr = rl / ST_I, k = rl - rr * ST_I;
          const int r = r0 + rr;
          const int lim = isP ? p.nI : p.nJ, base = isP ? i0 : j0;
          const float2* s1 = isP ? p.P1 : p.Q1;
          if (s1 && r < p.nR && base + k < lim && w < p.S) {
            const long long off = isP ? (long long)(i0 + k) * p.psi + (long long)r * p.psr + w
                                      : (long long)(j0 + k) * p.qsi + (long long)r * p.qsr + w;
            const float2 u = __ldg(s1 + off);
            v[t].x -= u.x; v[t].y -= u.y;
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 16; t++) {
        const int row = g + 8 * t;
        const bool isP = t < 8;
        const int rl = isP ? row : row - ST_R * ST_I;
        const int rr = rl / ST_I, k = rl - rr * ST_I;
        float2 x = v[t];
        if (isP) {
          x.x *= p.p_scale; x.y *= p.p_scale;
          if (p.conjP) x.y = -x.y;
          Ps[rr][k][bin] = x;
        } else {
          x.x *= p.q_scale; x.y *= p.q_scale;
          if (w == 0 && p.q_bias && j0 + k < p.nJ && r0 + rr < p.nR) x.x = fmaf(p.q_bias[j0 + k], p.q_bias_scale, x.x);
          if (p.conjQ) x.y = -x.y;
          Qs[rr][k][bin] = x;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < ST_R; rr++) {
      float2 pv[4], qv[8];
#pragma unroll
      for (int a = 0; a < 4; a++) pv[a] = Ps[rr][ig * 4 + a][bin];
#pragma unroll
      for (int b = 0; b < 8; b++) qv[b] = Qs[rr][jg * 8 + b][bin];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) cfma(acc[a][b], pv[a], qv[b]);
    }
    __syncthreads();
  }
  const long long w = w0 + bin;
  if (w >= p.S) return;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int i = i0 + ig * 4 + a;
    if (i >= p.nI) continue;
    const float ob = (w == 0 && p.out_bias) ? p.out_bias[i] * p.out_bias_scale : 0.f;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const int j = j0 + jg * 8 + b;
      if (j >= p.nJ) continue;
      p.out[(long long)i * p.osi + (long long)j * p.osj + w] = make_float2(fmaf(acc[a][b].x, p.out_scale, ob), acc[a][b].y * p.out_scale);
    }
  }
}


// ---- pipelined variant: the operand rows are copied global -> shared with cp.async (16 bytes per thread and copy, no
// registers), double buffered, so the copies of reduction chunk r+1 overlap the FMAs of chunk r.  Raw values are staged;
// the optional second source, conjugation and the DC bias are applied when the operands are read into registers, the
// operand scales are folded into the output scale.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int CONJP, int CONJQ, int HASP1, int HASQ1>
__global__ void __launch_bounds__(ST_THREADS, 2) spec_tiled2_kernel(TiledParams p) {
  extern __shared__ __align__(16) float2 dyn[];
  constexpr int TILE = ST_R * ST_I * ST_BINS;                 // float2 per operand per stage (ST_I == ST_J)
  constexpr int NBUF = 2 + HASP1 + HASQ1;                     // operand arrays per stage
  float2* stage_base[2] = {dyn, dyn + NBUF * TILE};
  const int tid = threadIdx.x, bin = tid & 31, g = tid >> 5;
  const int ig = g & 3, jg = g >> 2;
  const int n_it = (p.nI + ST_I - 1) / ST_I;
  const long long w0 = (long long)blockIdx.y * ST_BINS;
  const int i0 = (blockIdx.x % n_it) * ST_I, j0 = (blockIdx.x / n_it) * ST_J;
  // copy assignment: 64 rows per operand per stage, 16 threads (16 bytes = 2 bins each) per row, 4 rows per thread
  const int crow = tid >> 4, cseg = tid & 15;
  const long long wseg = w0 + 2 * cseg;
  const int seg_bytes = wseg + 2 <= p.S ? 16 : (wseg < p.S ? 8 : 0);
  auto issue = [&](int r0, float2* base) {
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int row = crow + 16 * t;                 // 0..63 = rr*16 + k
      const int rr = row >> 4, k = row & 15;
      const int r = r0 + rr;
      float2* dstrow = base + row * ST_BINS + 2 * cseg;
      {
        const bool ok = r < p.nR && i0 + k < p.nI;
        const long long off = (long long)(i0 + k) * p.psi + (long long)r * p.psr + wseg;
        cp_async16(dstrow, ok ? (const void*)(p.P0 + off) : (const void*)p.P0, ok ? seg_bytes : 0);
        if (HASP1) cp_async16(dstrow + 2 * TILE, ok ? (const void*)(p.P1 + off) : (const void*)p.P0, ok ? seg_bytes : 0);
      }
      {
        const bool ok = r < p.nR && j0 + k < p.nJ;
        const long long off = (long long)(j0 + k) * p.qsi + (long long)r * p.qsr + wseg;
        cp_async16(dstrow + TILE, ok ? (const void*)(p.Q0 + off) : (const void*)p.Q0, ok ? seg_bytes : 0);
        if (HASQ1) cp_async16(dstrow + (2 + HASP1) * TILE, ok ? (const void*)(p.Q1 + off) : (const void*)p.Q0, ok ? seg_bytes : 0);
      }
    }
    cp_async_commit();
  };
  float2 acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 8; b++) acc[a][b] = make_float2(0.f, 0.f);
  float qb[8];
#pragma unroll
  for (int b = 0; b < 8; b++)
    qb[b] = (w0 + bin == 0 && p.q_bias && j0 + jg * 8 + b < p.nJ) ? p.q_bias[j0 + jg * 8 + b] * p.q_bias_scale / p.q_scale : 0.f;
  const int nchunks = (p.nR + ST_R - 1) / ST_R;
  issue(0, stage_base[0]);
  for (int ch = 0; ch < nchunks; ch++) {
    if (ch + 1 < nchunks) {
      issue((ch + 1) * ST_R, stage_base[(ch + 1) & 1]);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float2* Ps = stage_base[ch & 1];
    const float2* Qs = Ps + TILE;
#pragma unroll
    for (int rr = 0; rr < ST_R; rr++) {
      float2 pv[4], qv[8];
      const bool live = ch * ST_R + rr < p.nR;  // (rows beyond nR were zero-filled; the bias must not be added there)
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const int idx = (rr * ST_I + ig * 4 + a) * ST_BINS + bin;
        pv[a] = Ps[idx];
        if (HASP1) { const float2 u = Ps[idx + 2 * TILE]; pv[a].x -= u.x; pv[a].y -= u.y; }
      }
#pragma unroll
      for (int b = 0; b < 8; b++) {
        const int idx = (rr * ST_J + jg * 8 + b) * ST_BINS + bin;
        qv[b] = Qs[idx];
        if (HASQ1) { const float2 u = Ps[idx + (2 + HASP1) * TILE]; qv[b].x -= u.x; qv[b].y -= u.y; }
        if (live) qv[b].x += qb[b];
      }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) {
          if (CONJQ) cfma_conj(acc[a][b], pv[a], qv[b]);
          else if (CONJP) cfma_conj(acc[a][b], qv[b], pv[a]);
          else cfma(acc[a][b], pv[a], qv[b]);
        }
    }
    __syncthreads();
  }
  const long long w = w0 + bin;
  if (w >= p.S) return;
  const float sc = p.out_scale * p.p_scale * p.q_scale;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int i = i0 + ig * 4 + a;
    if (i >= p.nI) continue;
    const float ob = (w == 0 && p.out_bias) ? p.out_bias[i] * p.out_bias_scale : 0.f;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const int j = j0 + jg * 8 + b;
      if (j >= p.nJ) continue;
      p.out[(long long)i * p.osi + (long long)j * p.osj + w] = make_float2(fmaf(acc[a][b].x, sc, ob), acc[a][b].y * sc);
    }
  }
}

template <int CONJP, int CONJQ, int HASP1, int HASQ1>
static int run_spec_tiled2(aefft_ctx* ctx, const TiledParams& p, dim3 grid) {
  const size_t smem = (size_t)2 * (2 + HASP1 + HASQ1) * ST_R * ST_I * ST_BINS * sizeof(float2);
  AE_TRY(ctx->ensure_dyn_smem((const void*)spec_tiled2_kernel<CONJP, CONJQ, HASP1, HASQ1>, smem));
  spec_tiled2_kernel<CONJP, CONJQ, HASP1, HASQ1><<<grid, ST_THREADS, smem, ctx->stream>>>(p);
  return AEFFT_OK;
}

static int launch_spec_tiled(aefft_ctx* ctx, const TiledParams& p) {
  dim3 grid((unsigned)(((p.nI + ST_I - 1) / ST_I) * ((p.nJ + ST_J - 1) / ST_J)), (unsigned)((p.S + ST_BINS - 1) / ST_BINS));
  AE_ARG(grid.y <= 65535);
  // pipelined cp.async variant when every operand row is 16-byte aligned (S even, even strides) and q_scale != 0
  const bool al = p.S % 2 == 0 && p.psi % 2 == 0 && p.psr % 2 == 0 && p.qsi % 2 == 0 && p.qsr % 2 == 0 &&
                  ((((uintptr_t)p.P0 | (uintptr_t)p.P1 | (uintptr_t)p.Q0 | (uintptr_t)p.Q1) & 15) == 0) && p.q_scale != 0.f;
  int rc = AEFFT_ERR_UNSUPPORTED;
  if (al && !getenv("AEFFT_NO_SPEC_PIPE")) {
    if (!p.conjP && !p.conjQ && !p.P1 && !p.Q1) rc = run_spec_tiled2<0, 0, 0, 0>(ctx, p, grid);
    else if (p.conjP && !p.conjQ && !p.P1 && p.Q1) rc = run_spec_tiled2<1, 0, 0, 1>(ctx, p, grid);
    else if (!p.conjP && p.conjQ && !p.P1 && !p.Q1) rc = run_spec_tiled2<0, 1, 0, 0>(ctx, p, grid);
    else if (!p.conjP && p.conjQ && p.P1 && !p.Q1) rc = run_spec_tiled2<0, 1, 1, 0>(ctx, p, grid);
  }
  if (rc == AEFFT_ERR_UNSUPPORTED) spec_tiled_kernel<<<grid, ST_THREADS, 0, ctx->stream>>>(p);
  else if (rc != AEFFT_OK) return rc;
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

int launch_spec_contract(aefft_ctx* ctx, int64_t B, int C, int O, int64_t S, const float2* in0, const float2* in1,
                         const float2* W, int64_t w_so, int64_t w_sc, int conjW, float in_scale, const float* bias,
                         float bias_scale, float2* out) {
  AE_ARG(B > 0 && C > 0 && O > 0 && S > 0);
  if (!in1 && !conjW && w_so == (int64_t)C * S && w_sc == S) {  // plain conv_k with few channels on one side
    const int rc = launch_spec_conv_reg(ctx, B, C, O, S, in0, W, bias, bias_scale, in_scale, out);
    if (rc != AEFFT_ERR_UNSUPPORTED) return rc;
  }
  ContractParams p{in0, in1, W, bias, out, w_so, w_sc, S, (int)B, C, O, conjW, in_scale, bias_scale};
  if (O >= 8 && B >= 8 && !getenv("AEFFT_NO_SPEC_TILED")) {
    // i = output channel (P = W), j = frame (Q = in), r = input channel
    TiledParams t{W, nullptr, in0, in1, nullptr, bias, out, w_so, w_sc, (long long)C * S, S, S, (long long)O * S, S,
                  O, (int)B, C, conjW, 0, 1.f, in_scale, 1.f, bias_scale, 0.f};
    ProfScope prof(ctx, "spec_contract_tiled", 8.0 * B * C * O * S, 8.0 * S * ((double)B * C * (in1 ? 2 : 1) + (double)B * O + (double)C * O));
    return launch_spec_tiled(ctx, t);
  }
  // register tile per thread (one bin): 8 outputs x 8 frames when both extents allow it (16 loads feed 64 complex FMAs
  // and every operand is re-read 2x less often), else 4 x 4
  const bool big = false;  // 8x8 register tiles measured slower (226 registers, occupancy); see spec_tiled_kernel
  const int ot = big ? 8 : SC_OT, ft = big ? 8 : SC_FT;
  dim3 grid((unsigned)((S + SC_THREADS - 1) / SC_THREADS), (O + ot - 1) / ot, (unsigned)((B + ft - 1) / ft));
  AE_ARG(grid.z <= 65535 && grid.y <= 65535);
  ProfScope prof(ctx, "spec_contract", 8.0 * B * C * O * S, 8.0 * S * ((double)B * C * (in1 ? 2 : 1) + (double)B * O + (double)C * O));
  if (big) spec_contract_kernel<8, 8><<<grid, SC_THREADS, 0, ctx->stream>>>(p);
  else spec_contract_kernel<SC_OT, SC_FT><<<grid, SC_THREADS, 0, ctx->stream>>>(p);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// ------------------------------------------------------------------------------------------------ per-bin outer product
// out[a][c][w] = scale * sum_b (A0[b][a][w] - A1[b][a][w]) * conj( bm_alpha * Bm[b][c][w] + [w==0] bm_bias[c]*bm_bias_scale )
//   dC = G . conj(X)           (gradient_k_io :437-445)
//   dF = E . conj(H-hat)       (:447-459; H-hat = dM*H - (dM-1)*b*Nx*Ny at DC: quirk F1, no /dM)
// The frame index is the reduction dimension (this repo's batch extension; B=1 is the reference).
constexpr int SO_AT = 4, SO_CT = 4;

struct OuterParams {
  const float2* A0;
  const float2* A1;
  const float2* Bm;
  const float* bm_bias;
  float2* out;
  long long S;
  int B, nA, nC;
  float bm_alpha, bm_bias_scale, scale;
};

template <int SO_AT, int SO_CT>
__global__ void __launch_bounds__(SC_THREADS) spec_outer_kernel(OuterParams p) {
  const long long w = (long long)blockIdx.x * SC_THREADS + threadIdx.x;
  if (w >= p.S) return;
  const int a0 = blockIdx.y * SO_AT, c0 = blockIdx.z * SO_CT;
  float2 acc[SO_AT][SO_CT];
#pragma unroll
  for (int a = 0; a < SO_AT; a++)
#pragma unroll
    for (int c = 0; c < SO_CT; c++) acc[a][c] = make_float2(0.f, 0.f);
  float cb[SO_CT];
#pragma unroll
  for (int c = 0; c < SO_CT; c++) cb[c] = (w == 0 && p.bm_bias && c0 + c < p.nC) ? p.bm_bias[c0 + c] * p.bm_bias_scale : 0.f;
  for (int b = 0; b < p.B; b++) {
    float2 av[SO_AT], bv[SO_CT];
#pragma unroll
    for (int a = 0; a < SO_AT; a++) {
      av[a] = make_float2(0.f, 0.f);
      if (a0 + a < p.nA) {
        const long long off = ((long long)b * p.nA + a0 + a) * p.S + w;
        av[a] = p.A0[off];
        if (p.A1) { float2 u = p.A1[off]; av[a].x -= u.x; av[a].y -= u.y; }
      }
    }
#pragma unroll
    for (int c = 0; c < SO_CT; c++) {
      bv[c] = make_float2(0.f, 0.f);
      if (c0 + c < p.nC) {
        float2 v = p.Bm[((long long)b * p.nC + c0 + c) * p.S + w];
        bv[c] = make_float2(fmaf(v.x, p.bm_alpha, cb[c]), v.y * p.bm_alpha);
      }
    }
#pragma unroll
    for (int a = 0; a < SO_AT; a++)
#pragma unroll
      for (int c = 0; c < SO_CT; c++) cfma_conj(acc[a][c], av[a], bv[c]);
  }
#pragma unroll
  for (int a = 0; a < SO_AT; a++)
#pragma unroll
    for (int c = 0; c < SO_CT; c++)
      if (a0 + a < p.nA && c0 + c < p.nC)
        p.out[((long long)(a0 + a) * p.nC + c0 + c) * p.S + w] = make_float2(acc[a][c].x * p.scale, acc[a][c].y * p.scale);
}

int launch_spec_outer(aefft_ctx* ctx, int64_t B, int nA, int nC, int64_t S, const float2* A0, const float2* A1,
                      const float2* Bm, float bm_alpha, const float* bm_bias, float bm_bias_scale, float scale, float2* out) {
  AE_ARG(B > 0 && nA > 0 && nC > 0 && S > 0);
  if (nA >= 8 && nC >= 8 && !getenv("AEFFT_NO_SPEC_TILED")) {
    // i = a (P = A), j = c (Q = Bm, conjugated, + bias at DC), r = frame
    TiledParams t{A0, A1, Bm, nullptr, bm_bias, nullptr, out, S, (long long)nA * S, S, (long long)nC * S, (long long)nC * S, S, S,
                  nA, nC, (int)B, 0, 1, 1.f, bm_alpha, scale, 0.f, bm_bias_scale};
    ProfScope prof(ctx, "spec_outer_tiled", 8.0 * B * nA * nC * S, 8.0 * S * ((double)B * nA * (A1 ? 2 : 1) + (double)B * nC + (double)nA * nC));
    return launch_spec_tiled(ctx, t);
  }
  OuterParams p{A0, A1, Bm, bm_bias, out, S, (int)B, nA, nC, bm_alpha, bm_bias_scale, scale};
  const bool big = false;
  const int at = big ? 8 : SO_AT, ct = big ? 8 : SO_CT;
  dim3 grid((unsigned)((S + SC_THREADS - 1) / SC_THREADS), (nA + at - 1) / at, (nC + ct - 1) / ct);
  AE_ARG(grid.z <= 65535 && grid.y <= 65535);
  ProfScope prof(ctx, "spec_outer", 8.0 * B * nA * nC * S, 8.0 * S * ((double)B * nA * (A1 ? 2 : 1) + (double)B * nC + (double)nA * nC));
  if (big) spec_outer_kernel<8, 8><<<grid, SC_THREADS, 0, ctx->stream>>>(p);
  else spec_outer_kernel<SO_AT, SO_CT><<<grid, SC_THREADS, 0, ctx->stream>>>(p);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// db[m] = gscale * sum_b Re G[b][m](0) ; dp[d] = gscale * sum_b Re (O - Xt)[b][d](0)     (:462-473)
__global__ void spec_dc_sums_kernel(const float2* __restrict__ G, const float2* __restrict__ O, const float2* __restrict__ Xt,
                                    float* __restrict__ db, float* __restrict__ dp, int B, int dM, int dD, long long S,
                                    float gscale) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < dM) {
    double s = 0.0;
    for (int b = 0; b < B; b++) s += (double)G[((long long)b * dM + n) * S].x;
    db[n] = (float)(s * (double)gscale);
  } else if (n < dM + dD) {
    const int d = n - dM;
    double s = 0.0;
    for (int b = 0; b < B; b++) s += (double)O[((long long)b * dD + d) * S].x - (double)Xt[((long long)b * dD + d) * S].x;
    dp[d] = (float)(s * (double)gscale);
  }
}

int launch_spec_dc_sums(aefft_ctx* ctx, int64_t B, int dM, int dD, int64_t S, const float2* G, const float2* O,
                        const float2* Xt, float* db, float* dp, float gscale) {
  spec_dc_sums_kernel<<<(dM + dD + 127) / 128, 128, 0, ctx->stream>>>(G, O, Xt, db, dp, (int)B, dM, dD, S, gscale);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// ------------------------------------------------------------------------------------------------ pad / shrink
// pad_k (:570-600) / kernel_pad (:1018-1064): img[(k-Nk/2) mod Nx][(l-Nl/2) mod Ny] = w[k][l], everything else 0.
// One thread per image pixel: writes the whole (zero-filled) image in one pass (the reference memsets first).
__global__ void pad_kernel(const float* __restrict__ taps, float* __restrict__ img, long long n_img, int Nx, int Ny,
                           int Nk, int Nl) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * Nx * Ny) return;
  const int j = idx % Ny;
  const int i = (idx / Ny) % Nx;
  const long long n = idx / ((long long)Nx * Ny);
  // inverse of the wrap: k = i + Nk/2 (i small) or i - Nx + Nk/2 (i near Nx)
  int k = i + Nk / 2;
  if (k >= Nk) k = i - Nx + Nk / 2;
  int l = j + Nl / 2;
  if (l >= Nl) l = j - Ny + Nl / 2;
  float v = 0.f;
  if (k >= 0 && k < Nk && l >= 0 && l < Nl) v = __ldg(taps + (n * Nk + k) * Nl + l);
  img[idx] = v;
}

int launch_pad(aefft_ctx* ctx, int64_t n_img, int Nx, int Ny, int Nk, int Nl, const float* taps, float* img) {
  AE_ARG(Nk <= Nx && Nl <= Ny);
  const long long total = (long long)n_img * Nx * Ny;
  ProfScope prof(ctx, "pad_k", 0.0, 4.0 * total);
  pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(taps, img, n_img, Nx, Ny, Nk, Nl);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// shrink_k (:535-565) / kernel_invpad (:1069-1112): w[k][l] = img[(k-Nk/2) mod Nx][(l-Nl/2) mod Ny]
__global__ void shrink_kernel(const float* __restrict__ img, float* __restrict__ taps, long long n_img, int Nx, int Ny,
                              int Nk, int Nl) {
  long long idk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idk >= n_img * Nk * Nl) return;
  const int l = idk % Nl;
  const int k = (idk / Nl) % Nk;
  const long long n = idk / ((long long)Nk * Nl);
  const int i = k >= Nk / 2 ? k - Nk / 2 : k + Nx - Nk / 2;
  const int j = l >= Nl / 2 ? l - Nl / 2 : l + Ny - Nl / 2;
  taps[idk] = img[(n * Nx + i) * Ny + j];
}

int launch_shrink(aefft_ctx* ctx, int64_t n_img, int Nx, int Ny, int Nk, int Nl, const float* img, float* taps) {
  const long long total = (long long)n_img * Nk * Nl;
  shrink_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(img, taps, n_img, Nx, Ny, Nk, Nl);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}


// ------------------------------------------------------------------------------------------------ pruned kernel DFTs
constexpr int PD_MAXT = 8;  // taps per axis
constexpr int ST_UB = 4;    // spectrum rows in flight per thread (spectrum_to_taps)

// Both kernels are instantiated for the tap counts the reference's parameter file can produce (3, 5, 7 per axis) plus a
// generic 8 x 8 variant; index arithmetic is 32-bit (bins per image < 2^31).

// 128 threads = 128 columns of one image's (slab of the) half spectrum, 16 rows per CTA: grid (col tiles, row tiles, n_img)
template <int NK, int NL>
__global__ void __launch_bounds__(128) kernel_spectrum_direct_kernel(const float* __restrict__ taps, float2* __restrict__ spec,
                                                                      int Nx, int Ny, int Nk, int Nl,
                                                                      const float2* __restrict__ twx,
                                                                      const float2* __restrict__ twy, int col0, int Nyr,
                                                                      int rows_per_cta) {
  // (col0, Nyr): the slab of spectrum columns [col0, col0 + Nyr) this device owns (whole half spectrum: 0, Ny/2+1).
  // One thread per spectrum column: the column factor t[k] = sum_l c[k][l] Ey[l](wy) is formed once, every row then costs
  // NK complex multiply-adds with the (warp-uniform) row factors Ex[k](wx) instead of NK * NL + NK.
  __shared__ float c[PD_MAXT * PD_MAXT];
  const int nk = NK ? NK : Nk, nl = NL ? NL : Nl;
  const unsigned n = blockIdx.z;
  if (threadIdx.x < nk * nl) c[threadIdx.x] = taps[(size_t)n * nk * nl + threadIdx.x];
  __syncthreads();
  const int wl = blockIdx.x * blockDim.x + threadIdx.x;
  if (wl >= Nyr) return;
  const int wy = col0 + wl;
  float2 t[NK ? NK : PD_MAXT];
#pragma unroll
  for (int k = 0; k < (NK ? NK : PD_MAXT); k++) t[k] = make_float2(0.f, 0.f);
#pragma unroll
  for (int l = 0; l < (NL ? NL : PD_MAXT); l++) {
    if (l < nl) {
      const float2 ey = __ldg(twy + (tw_index(wy, l - nl / 2, Ny)));
#pragma unroll
      for (int k = 0; k < (NK ? NK : PD_MAXT); k++) {
        if (k < nk) {
          const float cv = c[k * nl + l];
          t[k].x = fmaf(cv, ey.x, t[k].x);
          t[k].y = fmaf(cv, ey.y, t[k].y);
        }
      }
    }
  }
  const int wx0 = blockIdx.y * rows_per_cta, wx1 = min(Nx, wx0 + rows_per_cta);
  float2* o = spec + ((size_t)n * Nx + wx0) * Nyr + wl;
  for (int wx = wx0; wx < wx1; wx++, o += Nyr) {
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < (NK ? NK : PD_MAXT); k++) {
      if (k < nk) {
        const float2 ex = __ldg(twx + (tw_index(wx, k - nk / 2, Nx)));  // warp-uniform address
        cfma(acc, ex, t[k]);
      }
    }
    *o = acc;
  }
}

int launch_kernel_spectrum_direct(aefft_ctx* ctx, int64_t n_img, int Nx, int Ny, int Nk, int Nl, const float* taps,
                                  float2* spec, int col0, int ncols) {
  AE_ARG(n_img > 0 && n_img <= 65535 && Nx <= 65535 && Nk <= PD_MAXT && Nl <= PD_MAXT && Nk <= Nx && Nl <= Ny);
  const float2 *twx, *twy;
  AE_TRY(get_twiddles(ctx, Nx, &twx));
  AE_TRY(get_twiddles(ctx, Ny, &twy));
  const int Nyr = ncols > 0 ? ncols : Ny / 2 + 1;
  const long long S = (long long)Nx * Nyr;
  ProfScope prof(ctx, "kernel_spectrum", 8.0 * n_img * S * (Nk + Nk * Nl / 4.0), 8.0 * n_img * S);
  const int threads = 128, rows = 16;
  dim3 grid((Nyr + threads - 1) / threads, (Nx + rows - 1) / rows, (unsigned)n_img);
  if (Nk == 5 && Nl == 5) kernel_spectrum_direct_kernel<5, 5><<<grid, threads, 0, ctx->stream>>>(taps, spec, Nx, Ny, Nk, Nl, twx, twy, col0, Nyr, rows);
  else if (Nk == 3 && Nl == 3) kernel_spectrum_direct_kernel<3, 3><<<grid, threads, 0, ctx->stream>>>(taps, spec, Nx, Ny, Nk, Nl, twx, twy, col0, Nyr, rows);
  else if (Nk == 7 && Nl == 7) kernel_spectrum_direct_kernel<7, 7><<<grid, threads, 0, ctx->stream>>>(taps, spec, Nx, Ny, Nk, Nl, twx, twy, col0, Nyr, rows);
  else kernel_spectrum_direct_kernel<0, 0><<<grid, threads, 0, ctx->stream>>>(taps, spec, Nx, Ny, Nk, Nl, twx, twy, col0, Nyr, rows);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// grid (n_split, n_img): every CTA reduces a contiguous range of spectrum ROWS of one image to Nk*Nl partial sums (fixed
// order: per-thread sums over the rows, warp shuffles, shared memory), a second tiny kernel adds the splits --
// deterministic.  A thread owns one column wy (its Ey factors stay in registers) and walks down the rows.
template <int NK, int NL>
__global__ void __launch_bounds__(768) spectrum_to_taps_kernel(const float2* __restrict__ spec, float* __restrict__ part, int Nx,
                                                                int Ny, int Nk, int Nl, const float2* __restrict__ twx,
                                                                const float2* __restrict__ twy, int col0, int Nyr) {
  constexpr int TK = NK ? NK : PD_MAXT, TL = NL ? NL : PD_MAXT;
  const int nk = NK ? NK : Nk, nl = NL ? NL : Nl;
  const unsigned n = blockIdx.y;
  const int nsplit = gridDim.x, sp = blockIdx.x;
  const int r_lo = (int)((long long)Nx * sp / nsplit), r_hi = (int)((long long)Nx * (sp + 1) / nsplit);
  const float2* z = spec + (size_t)n * Nx * Nyr;
  float g[TK * TL];
#pragma unroll
  for (int t = 0; t < TK * TL; t++) g[t] = 0.f;
  // g[k][l] = sum over bins of h * Re( v * conj(Ex[k](wx)) * conj(Ey[l](wy)) ).  Ey depends on the column only, so the
  // row sum b[k] = sum_wx v * conj(Ex[k]) (NK complex MACs per bin) is taken first and the NL column factors are applied
  // once per column -- NK instead of NK * NL + NL multiply-adds per bin.
  for (int wl = threadIdx.x; wl < Nyr; wl += blockDim.x) {
    const int wy = col0 + wl;
    const float h = (wy == 0 || wy == Ny / 2) ? 1.f : 2.f;
    float2 b[TK];
#pragma unroll
    for (int k = 0; k < TK; k++) b[k] = make_float2(0.f, 0.f);
    // ST_UB rows in flight per thread before the first use (one dependent 8-byte load per row left the kernel latency
    // bound at 0.8 TB/s); the sums keep their order
    int wx = r_lo;
    for (; wx + ST_UB <= r_hi; wx += ST_UB) {
      float2 v[ST_UB];
#pragma unroll
      for (int u = 0; u < ST_UB; u++) v[u] = __ldg(z + (size_t)(wx + u) * Nyr + wl);
#pragma unroll
      for (int u = 0; u < ST_UB; u++) {
#pragma unroll
        for (int k = 0; k < TK; k++) {
          if (k < nk) {
            const float2 e = __ldg(twx + (tw_index(wx + u, k - nk / 2, Nx)));  // warp-uniform address
            b[k].x = fmaf(v[u].x, e.x, fmaf(v[u].y, e.y, b[k].x));   // v * conj(e)
            b[k].y = fmaf(v[u].y, e.x, fmaf(-v[u].x, e.y, b[k].y));
          }
        }
      }
    }
    for (; wx < r_hi; wx++) {
      const float2 v = __ldg(z + (size_t)wx * Nyr + wl);
#pragma unroll
      for (int k = 0; k < TK; k++) {
        if (k < nk) {
          const float2 e = __ldg(twx + (tw_index(wx, k - nk / 2, Nx)));
          b[k].x = fmaf(v.x, e.x, fmaf(v.y, e.y, b[k].x));
          b[k].y = fmaf(v.y, e.x, fmaf(-v.x, e.y, b[k].y));
        }
      }
    }
#pragma unroll
    for (int l = 0; l < TL; l++) {
      if (l < nl) {
        const float2 ey = __ldg(twy + (tw_index(wy, l - nl / 2, Ny)));
#pragma unroll
        for (int k = 0; k < TK; k++)
          if (k < nk) g[k * TL + l] = fmaf(h, b[k].x * ey.x + b[k].y * ey.y, g[k * TL + l]);  // h * Re(b * conj(ey))
      }
    }
  }
  __shared__ float red[24][TK * TL];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < TK * TL; t++) {
    float v = g[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][t] = v;
  }
  __syncthreads();
  if (threadIdx.x < nk * nl) {
    const int k = threadIdx.x / nl, l = threadIdx.x - k * nl;
    float v = 0.f;
    for (int wq = 0; wq < (int)(blockDim.x >> 5); wq++) v += red[wq][k * TL + l];
    part[((size_t)n * nsplit + sp) * nk * nl + threadIdx.x] = v;
  }
}
__global__ void spectrum_to_taps_final_kernel(const float* __restrict__ part, float* __restrict__ taps, long long total, int nsplit,
                                              int T, float scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long n = idx / T;
  const int t = (int)(idx - n * T);
  double s = 0.0;
  for (int sp = 0; sp < nsplit; sp++) s += (double)part[(n * nsplit + sp) * T + t];
  taps[idx] = (float)(s * (double)scale);
}

int launch_spectrum_to_taps(aefft_ctx* ctx, int64_t n_img, int Nx, int Ny, int Nk, int Nl, const float2* spec, float* taps,
                            float scale, int col0, int ncols) {
  AE_ARG(n_img > 0 && n_img <= 65535 && Nk <= PD_MAXT && Nl <= PD_MAXT && Nk <= Nx && Nl <= Ny);
  const float2 *twx, *twy;
  AE_TRY(get_twiddles(ctx, Nx, &twx));
  AE_TRY(get_twiddles(ctx, Ny, &twy));
  const int Nyr = ncols > 0 ? ncols : Ny / 2 + 1;
  const long long S = (long long)Nx * Nyr;
  // enough CTAs to fill the machine (row ranges of >= 8 rows)
  int nsplit = (int)((4LL * ctx->sm_count + n_img - 1) / n_img);
  if (nsplit > Nx / 8) nsplit = Nx / 8;
  if (nsplit < 1) nsplit = 1;
  // Half spectra have 2^k + 1 columns and a thread owns a column: size the CTA so that the columns need as few passes as
  // possible (257 columns -> 288 threads, 513 -> 544) instead of a second pass for the Nyquist column alone.
  const int passes = (Nyr + 767) / 768;
  const int threads = (((Nyr + passes - 1) / passes) + 31) / 32 * 32;
  float* part;
  AE_TRY(ctx->getT("s2t_part", (size_t)n_img * nsplit * Nk * Nl, &part));
  {
    ProfScope prof(ctx, "spectrum_to_taps", 2.0 * n_img * S * (4.0 * Nl + 2.0 * Nk * Nl), 8.0 * n_img * S);
    dim3 grid(nsplit, (unsigned)n_img);
    if (Nk == 5 && Nl == 5) spectrum_to_taps_kernel<5, 5><<<grid, threads, 0, ctx->stream>>>(spec, part, Nx, Ny, Nk, Nl, twx, twy, col0, Nyr);
    else if (Nk == 3 && Nl == 3) spectrum_to_taps_kernel<3, 3><<<grid, threads, 0, ctx->stream>>>(spec, part, Nx, Ny, Nk, Nl, twx, twy, col0, Nyr);
    else if (Nk == 7 && Nl == 7) spectrum_to_taps_kernel<7, 7><<<grid, threads, 0, ctx->stream>>>(spec, part, Nx, Ny, Nk, Nl, twx, twy, col0, Nyr);
    else spectrum_to_taps_kernel<0, 0><<<grid, threads, 0, ctx->stream>>>(spec, part, Nx, Ny, Nk, Nl, twx, twy, col0, Nyr);
  }
  const long long total = (long long)n_img * Nk * Nl;
  spectrum_to_taps_final_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(part, taps, total, nsplit, Nk * Nl, scale);
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// ------------------------------------------------------------------------------------------------ update
// gradient_diff (:709-753): cd[m][d][k][l] = sum_{m1!=m, d1!=d} (c[m][d][k][l]-c[m1][d1][k][l]) / |c[m][d]-c[m1][d1]|^2,
// fd likewise on f[d][m]; bd[m] = sum_{m1!=m} 1/(b[m]-b[m1]); pd[d] = sum_{d1!=d} 1/(p[d]-p[d1]).
// One CTA per kernel (m,d): squared distances to every other kernel first (shared memory), then the tap sums.
constexpr int GD_CHUNK = 4096;
__global__ void gradient_diff_kernel(const float* __restrict__ c, const float* __restrict__ f, const float* __restrict__ b,
                                     const float* __restrict__ p, float* __restrict__ cd, float* __restrict__ fd,
                                     float* __restrict__ bd, float* __restrict__ pd, int dM, int dD, int T) {
  // inverse squared distances to the other kernels, GD_CHUNK of them at a time (any dM*dD fits; the o = m1*dD+d1 order of
  // the sums is the reference's loop order)
  __shared__ float inv_c[GD_CHUNK], inv_f[GD_CHUNK];
  const int m = blockIdx.x / dD, d = blockIdx.x % dD;
  const float* cm = c + (m * dD + d) * T;
  const float* fm = f + (d * dM + m) * T;
  float sc = 0.f, sf = 0.f;  // thread t < T owns tap t
  for (int o0 = 0; o0 < dM * dD; o0 += GD_CHUNK) {
    const int o1 = min(o0 + GD_CHUNK, dM * dD);
    for (int o = o0 + threadIdx.x; o < o1; o += blockDim.x) {
      const int m1 = o / dD, d1 = o % dD;
      float dc = 0.f, df = 0.f;
      if (m1 != m && d1 != d) {
        const float* c1 = c + (m1 * dD + d1) * T;
        const float* f1 = f + (d1 * dM + m1) * T;
        for (int t = 0; t < T; t++) {
          float x = cm[t] - c1[t], y = fm[t] - f1[t];
          dc = fmaf(x, x, dc);
          df = fmaf(y, y, df);
        }
        dc = 1.f / dc;
        df = 1.f / df;
      }
      inv_c[o - o0] = dc;
      inv_f[o - o0] = df;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      for (int o = o0; o < o1; o++) {
        const int m1 = o / dD, d1 = o - m1 * dD;
        if (m1 == m || d1 == d) continue;
        sc += (cm[t] - c[(m1 * dD + d1) * T + t]) * inv_c[o - o0];
        sf += (fm[t] - f[(d1 * dM + m1) * T + t]) * inv_f[o - o0];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < T) {
    cd[(m * dD + d) * T + threadIdx.x] = sc;
    fd[(d * dM + m) * T + threadIdx.x] = sf;
  }
  if (threadIdx.x == 0) {
    if (d == 0) {
      float s = 0.f;
      for (int m1 = 0; m1 < dM; m1++)
        if (m1 != m) s += 1.f / (b[m] - b[m1]);
      bd[m] = s;
    }
    if (m == 0) {
      float s = 0.f;
      for (int d1 = 0; d1 < dD; d1++)
        if (d1 != d) s += 1.f / (p[d] - p[d1]);
      pd[d] = s;
    }
  }
}


// Staging copy of c and f for the tiled kernel below: [tensor][npad][TP] with a kernel's T taps followed by its indices
// (b1, b2) and |x|^2 -- TP = T + 3 floats = whole 16-byte units, so a 64-kernel tile is one contiguous block that
// cp.async moves 16 bytes at a time.  Rows beyond n: zero taps, |x|^2 = +inf (distance inf, weight 0).
template <int T>
__global__ void gradient_diff_pack_kernel(const float* __restrict__ c, const float* __restrict__ f, float* __restrict__ xp,
                                          int dM, int dD, int npad) {
  constexpr int TP = T + 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * npad) return;
  const int isf = i >= npad, b = isf ? i - npad : i, n = dM * dD, n2 = isf ? dM : dD;
  float* o = xp + (size_t)i * TP;
  if (b < n) {
    const float* x = (isf ? f : c) + (size_t)b * T;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < T; t++) { const float v = x[t]; o[t] = v; s = fmaf(v, v, s); }
    const int b1 = b / n2;
    o[T] = __int_as_float(b1);
    o[T + 1] = __int_as_float(b - b1 * n2);
    o[T + 2] = s;
  } else {
#pragma unroll
    for (int t = 0; t < T + 2; t++) o[t] = 0.f;
    o[T + 2] = __int_as_float(0x7f800000);
  }
}

// Tiled form of the same sums (the one-CTA-per-kernel version above is O((dM dD)^2) with 25 active threads: 400 ms per call
// at 128 -> 256 channels).  X is [n1][n2][T]; kernel a = (a1, a2) interacts with b = (b1, b2) iff a1 != b1 and a2 != b2:
//   xd[a][t] = sum_b w_ab (x[a][t] - x[b][t]) = x[a][t] * sum_b w_ab - sum_b w_ab x[b][t],  w_ab = 1 / |x[a] - x[b]|^2
// A CTA owns 64 kernels a (one per thread, its taps in registers) and streams all b through shared memory in tiles of
// 64 (cp.async, double buffered, from the packed staging copy); the 4 thread groups of a CTA take every fourth b of a
// tile and are combined at the end.  grid.y selects the tensor (c or f).
// Measured at 128 -> 256 channels (32 768 kernels per tensor, tools/gdiff_probe.py), per launch: 14.5 ms with an integer
// division and a full-precision 1/x per pair in the loop; 8.3 ms with the indices staged next to the taps, prefetched
// tiles and rcp.approx; 7.7 ms with packed FADD2 / FFMA2 (the packed forms halve the issue slots, not the FMA-pipe time:
// tools/probe_ffma2.cu measures 128 fp32 lanes per clock per SM either way); 7.3 ms with dot-product distances; 6.1 ms with
// 16-byte tile copies and four accumulators for the dot product.  Interleaving two pairs by hand, a branch-free loop
// and three CTAs per SM (80 registers, spills: 9.8 ms) gave nothing more: the loop runs ~52 instructions per pair at 0.54
// issues per scheduler-clock with 4 warps per scheduler.  Splitting the 25 taps of a kernel over two lanes (80 registers, three
// CTAs per SM, partial dot products joined by a shuffle) was slower as well: 8.6 ms -- both lanes repeat the reciprocal and
// index work and the loop is bound by issued instructions, not by latency.
template <int T>
__global__ void __launch_bounds__(256) gradient_diff_tiled_kernel(const float* __restrict__ c, const float* __restrict__ f,
                                                                   float* __restrict__ cd, float* __restrict__ fd, int dM, int dD,
                                                                   int tile0, float* __restrict__ part,
                                                                   const float* __restrict__ xp) {
  // xp: the staging copy [tensor][npad][TP] (gradient_diff_pack_kernel)
  // gridDim.z > 1: the streamed kernels b are split into gridDim.z chunks (a bin-sharded device owns few row tiles -- 64 of
  // 512 at 8 devices -- and one CTA per tile would leave most SMs idle while each CTA walks ALL b); every (tile, chunk) CTA
  // then writes its partial (sw, swx[T]) to `part` [tensor][chunk][local a][T+1] and gradient_diff_finish_kernel combines
  // them in chunk order (deterministic).
  // tile0: first 64-kernel tile of this launch (bin-sharded devices split the rows; the hook adds the results)
  // The streamed kernels b sit in shared memory padded to T4 float4 per kernel and are read with 16-byte broadcast loads:
  // with scalar loads the inner loop issued 25 LDS per 77 arithmetic instructions and was bound by the shared-memory
  // pipe (one wavefront per clock) rather than by the FMA pipes.
  constexpr int T4 = (T + 3) / 4, TP = 4 * T4;
  static_assert(TP - T == 3, "the padding of a staged kernel carries its (b1, b2) indices and the valid flag");
  const bool isf = blockIdx.y == 1;
  const float* x = isf ? f : c;
  float* xd = isf ? fd : cd;
  const int n2 = isf ? dM : dD;
  const int n = dM * dD;
  // Two tile buffers: tile k+1 is fetched with cp.async while tile k is consumed (one barrier per tile).  Slots T, T+1, T+2
  // of a staged kernel hold b1, b2 (so the inner loop has no integer division) and 1.0 / 0.0 for rows inside / beyond n.
  __shared__ float4 tb[2][64][T4];
  __shared__ float red[3][64][T + 1];
  const int la = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int a = (tile0 + blockIdx.x) * 64 + la;
  const bool a_ok = a < n;
  const int a1 = a_ok ? a / n2 : -1, a2 = a_ok ? a - a1 * n2 : -1;
  // Taps live in registers as packed pairs (FADD2 / FFMA2, sm_100).  T is odd: TH pairs + the last tap as a scalar (its
  // pair partner in the staged tile is the b1 index).
  // Distances: |a - b|^2 = |a|^2 + |b|^2 - 2 a.b costs one packed FMA per tap pair instead of an add and an FMA (the loop is
  // bound by the FMA pipe); where that form cancels (kernels closer than 10 % of their norm: relative error of d^2 above
  // 6e-8 / 0.01) the differences are summed directly, which is what the reference does (:722-741).
  constexpr int TH = T / 2;
  static_assert(T % 2 == 1 && TH >= 4, "odd tap count, at least four tap pairs");
  float2 xa2[TH], swx2[TH];
  float xa_l, swx_l = 0.f, sw = 0.f, na = 0.f;
#pragma unroll
  for (int t = 0; t < TH; t++) {
    xa2[t] = a_ok ? make_float2(x[(size_t)a * T + 2 * t], x[(size_t)a * T + 2 * t + 1]) : make_float2(0.f, 0.f);
    swx2[t] = make_float2(0.f, 0.f);
    na = fmaf(xa2[t].x, xa2[t].x, fmaf(xa2[t].y, xa2[t].y, na));
  }
  xa_l = a_ok ? x[(size_t)a * T + T - 1] : 0.f;
  na = fmaf(xa_l, xa_l, na);
  const int nbt = (n + 63) / 64;  // 64-kernel tiles of b
  const int bt_lo = (int)((long long)nbt * blockIdx.z / gridDim.z), bt_hi = (int)((long long)nbt * (blockIdx.z + 1) / gridDim.z);
  const float4* xp4 = reinterpret_cast<const float4*>(xp) + (size_t)blockIdx.y * nbt * 64 * T4;
  auto stage = [&](int bt, int buf) {
    const float4* src = xp4 + (size_t)bt * 64 * T4;
    for (int i = threadIdx.x; i < 64 * T4; i += 256) {
      const uint32_t d = (uint32_t)__cvta_generic_to_shared(&tb[buf][0][0] + i);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (bt_lo < bt_hi) stage(bt_lo, 0);
  for (int bt = bt_lo; bt < bt_hi; bt++) {
    const int buf = (bt - bt_lo) & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // tile bt has landed for every thread, and every thread is done with the other buffer
    if (bt + 1 < bt_hi) stage(bt + 1, buf ^ 1);
    // Two streamed kernels per iteration, written stage by stage so that the two dependency chains (loads -> dot product ->
    // reciprocal -> accumulation) interleave in program order: a warp issues in order and with 16 warps per SM one chain
    // per warp left the schedulers without an eligible warp half of the time.  No branch in the loop: a pair whose
    // dot-product distance cancels (d^2 < 1 % of |a|^2 + |b|^2) gets weight 0 here and is noted in `near`; those (rare)
    // pairs are added after the loop with directly summed differences, which is what the reference does (:722-741).
    unsigned near = 0;
#pragma unroll 1
    for (int j = grp; j < 64; j += 8) {
      float2 xb[2][2 * T4];  // xb[u][TH] = (last tap, b1), xb[u][TH + 1] = (b2, |b|^2)
#pragma unroll
      for (int u = 0; u < 2; u++)
#pragma unroll
        for (int q = 0; q < T4; q++) {
          const float4 v = tb[buf][j + 4 * u][q];
          xb[u][2 * q] = make_float2(v.x, v.y);
          xb[u][2 * q + 1] = make_float2(v.z, v.w);
        }
      float2 dq[2][4];
#pragma unroll
      for (int t = 0; t < TH; t++)
#pragma unroll
        for (int u = 0; u < 2; u++)
          dq[u][t & 3] = t < 4 ? __fmul2_rn(xa2[t], xb[u][t]) : __ffma2_rn(xa2[t], xb[u][t], dq[u][t & 3]);
      float w[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const float2 dp = __fadd2_rn(__fadd2_rn(dq[u][0], dq[u][1]), __fadd2_rn(dq[u][2], dq[u][3]));
        const float dot = fmaf(xa_l, xb[u][TH].x, dp.x) + dp.y;
        const float nn = na + xb[u][TH + 1].y;
        const float d2 = fmaf(-2.f, dot, nn);
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d2));
        const bool on = (__float_as_int(xb[u][TH].y) != a1) & (__float_as_int(xb[u][TH + 1].x) != a2);
        const bool cancels = d2 < 0.01f * nn;  // false for rows beyond n (nn = inf)
        if (on && cancels) near |= 1u << ((j >> 2) + u);
        w[u] = (on && !cancels) ? r : 0.f;
        sw += w[u];
      }
#pragma unroll
      for (int t = 0; t < TH; t++)
#pragma unroll
        for (int u = 0; u < 2; u++) swx2[t] = __ffma2_rn(make_float2(w[u], w[u]), xb[u][t], swx2[t]);
#pragma unroll
      for (int u = 0; u < 2; u++) swx_l = fmaf(w[u], xb[u][TH].x, swx_l);
    }
    while (near) {  // near-duplicate kernels: distances from directly summed differences
      const int jj = __ffs(near) - 1;
      near &= near - 1;
      const int j = 4 * jj + grp;
      float2 xb[2 * T4];
#pragma unroll
      for (int q = 0; q < T4; q++) {
        const float4 v = tb[buf][j][q];
        xb[2 * q] = make_float2(v.x, v.y);
        xb[2 * q + 1] = make_float2(v.z, v.w);
      }
      float2 d2p = make_float2(0.f, 0.f);
#pragma unroll
      for (int t = 0; t < TH; t++) {
        const float2 e = __fadd2_rn(xb[t], make_float2(-xa2[t].x, -xa2[t].y));
        d2p = __ffma2_rn(e, e, d2p);
      }
      const float el = xb[TH].x - xa_l;
      const float d2 = fmaf(el, el, d2p.x) + d2p.y;
      float w;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(d2));
      sw += w;
#pragma unroll
      for (int t = 0; t < TH; t++) swx2[t] = __ffma2_rn(make_float2(w, w), xb[t], swx2[t]);
      swx_l = fmaf(w, xb[TH].x, swx_l);
    }
  }
  // back to per-tap scalars for the reduction and the store
  float xa[T], swx[T];
#pragma unroll
  for (int t = 0; t < TH; t++) {
    xa[2 * t] = xa2[t].x; xa[2 * t + 1] = xa2[t].y;
    swx[2 * t] = swx2[t].x; swx[2 * t + 1] = swx2[t].y;
  }
  xa[T - 1] = xa_l;
  swx[T - 1] = swx_l;
  // combine the 4 groups (fixed order)
  if (grp > 0) {
#pragma unroll
    for (int t = 0; t < T; t++) red[grp - 1][la][t] = swx[t];
    red[grp - 1][la][T] = sw;
  }
  __syncthreads();
  if (grp == 0 && a_ok) {
    for (int g = 0; g < 3; g++) {
#pragma unroll
      for (int t = 0; t < T; t++) swx[t] += red[g][la][t];
      sw += red[g][la][T];
    }
    if (gridDim.z == 1) {
#pragma unroll
      for (int t = 0; t < T; t++) xd[(size_t)a * T + t] = xa[t] * sw - swx[t];
    } else {
      float* o = part + ((((size_t)blockIdx.y * gridDim.z + blockIdx.z) * gridDim.x + blockIdx.x) * 64 + la) * (T + 1);
#pragma unroll
      for (int t = 0; t < T; t++) o[t] = swx[t];
      o[T] = sw;
    }
  }
}
template <int T>
__global__ void gradient_diff_finish_kernel(const float* __restrict__ c, const float* __restrict__ f, float* __restrict__ cd,
                                            float* __restrict__ fd, const float* __restrict__ part, int n, int tile0, int ntiles,
                                            int nchunks) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (tensor, local a)
  if (idx >= 2 * ntiles * 64) return;
  const int isf = idx / (ntiles * 64), al = idx - isf * ntiles * 64;
  const int a = tile0 * 64 + al;
  if (a >= n) return;
  const float* x = isf ? f : c;
  float* xd = isf ? fd : cd;
  float sw = 0.f, swx[T];
#pragma unroll
  for (int t = 0; t < T; t++) swx[t] = 0.f;
  for (int z = 0; z < nchunks; z++) {
    const float* o = part + ((((size_t)isf * nchunks + z) * ntiles) * 64 + al) * (T + 1);
#pragma unroll
    for (int t = 0; t < T; t++) swx[t] += o[t];
    sw += o[T];
  }
#pragma unroll
  for (int t = 0; t < T; t++) xd[(size_t)a * T + t] = x[(size_t)a * T + t] * sw - swx[t];
}
__global__ void gradient_diff_bias_kernel(const float* __restrict__ b, const float* __restrict__ p, float* __restrict__ bd,
                                          float* __restrict__ pd, int dM, int dD) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dM) {
    float s = 0.f;
    for (int m1 = 0; m1 < dM; m1++)
      if (m1 != i) s += 1.f / (b[i] - b[m1]);
    bd[i] = s;
  } else if (i < dM + dD) {
    const int d = i - dM;
    float s = 0.f;
    for (int d1 = 0; d1 < dD; d1++)
      if (d1 != d) s += 1.f / (p[d] - p[d1]);
    pd[d] = s;
  }
}

__device__ __forceinline__ float clip10f(float g) { return g / fmaxf(10.f, fabsf(g)); }

// backprop_d (:605-652) / backprop_double (:657-704): v = (1-0.9)*del*clip(g) + 0.9*v ; w -= v, elementwise on the flat
// arrays (c and f share the flat index), g = w0*g_mse - w1*g_div when the multiobjective term is on.
__global__ void fft_update_kernel(float* __restrict__ c, float* __restrict__ f, float* __restrict__ b, float* __restrict__ p,
                                  const float* __restrict__ dck, const float* __restrict__ dfk, const float* __restrict__ db,
                                  const float* __restrict__ dp, float* __restrict__ Dc, float* __restrict__ Df,
                                  float* __restrict__ Db, float* __restrict__ Dp, const float* __restrict__ cd,
                                  const float* __restrict__ fd, const float* __restrict__ bd, const float* __restrict__ pd,
                                  int nC, int dM, int dD, float del, float w0, float w1) {
  const float alpha = 0.9f;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nC) return;
  {
    float g = cd ? w0 * dck[n] - w1 * cd[n] : dck[n];
    float v = (1.f - alpha) * del * clip10f(g) + alpha * Dc[n];
    c[n] -= v;
    Dc[n] = v;
  }
  {
    float g = fd ? w0 * dfk[n] - w1 * fd[n] : dfk[n];
    float v = (1.f - alpha) * del * clip10f(g) + alpha * Df[n];
    f[n] -= v;
    Df[n] = v;
  }
  if (n < dM) {
    float g = bd ? w0 * db[n] - w1 * bd[n] : db[n];
    float v = (1.f - alpha) * del * clip10f(g) + alpha * Db[n];
    b[n] -= v;
    Db[n] = v;
  }
  if (n < dD) {
    float g = pd ? w0 * dp[n] - w1 * pd[n] : dp[n];
    float v = (1.f - alpha) * del * clip10f(g) + alpha * Dp[n];
    p[n] -= v;
    Dp[n] = v;
  }
}

// Multiobjective (kernel-diversity) term [cd | fd | bd | pd] into div (2*nC + dM + dD floats, the layout of the gradient
// block).  rank/world split the kernels over bin-sharded devices: a device computes the rows of its share and leaves zeros
// elsewhere (bias terms on device 0), so that the sum over devices is the whole term.
int launch_gradient_diff(aefft_ctx* ctx, int dM, int dD, int Nk, int Nl, const float* c, const float* f, const float* b,
                         const float* p, float* div, int rank, int world) {
  const int T = Nk * Nl, nC = dM * dD * T, n = dM * dD;
  float *cd = div, *fd = cd + nC, *bd = fd + nC, *pd = bd + dM;
  AE_ARG(T <= 64 && world >= 1);
  ProfScope prof(ctx, "gradient_diff", 2.0 * 75.0 * (double)n * n / world, 8.0 * nC);
  if (world > 1) AE_CUDA(cudaMemsetAsync(div, 0, (size_t)(2 * nC + dM + dD) * sizeof(float), ctx->stream));
  if (n >= 256 && (T == 25 || T == 9)) {
    const int tiles = (n + 63) / 64;
    const int t0 = (int)((long long)tiles * rank / world), t1 = (int)((long long)tiles * (rank + 1) / world);
    if (t1 > t0) {
      const int nt = t1 - t0;
      // enough CTAs for ~4 per SM: split the streamed kernels into chunks when this device owns few row tiles
      int nchunks = (4 * ctx->sm_count + 2 * nt - 1) / (2 * nt);
      if (nchunks > tiles / 8) nchunks = tiles / 8;  // >= 8 b-tiles (512 kernels) per chunk
      if (nchunks < 1) nchunks = 1;
      if (const char* e = getenv("AEFFT_GDIFF_CHUNKS")) {  // tests force the chunked form on small shapes
        nchunks = atoi(e);
        if (nchunks > tiles) nchunks = tiles;
        if (nchunks < 1) nchunks = 1;
      }
      if (T == 25 && n >= 2048 && !getenv("AEFFT_GDIFF_CHUNKS") && !getenv("AEFFT_NO_GDIFF_TC")) {
        // the tensor-core form runs one CTA of 128 rows per SM: pick the chunk count whose last wave is the fullest
        const int atiles = 2 * ((nt + 1) / 2);
        long long best = -1;
        for (int z = 1; z <= 8 && z <= tiles / 8; z++) {
          const long long waves = ((long long)atiles * z + ctx->sm_count - 1) / ctx->sm_count;
          const long long cost = waves * ((tiles + z - 1) / z + 4);  // + the per-CTA prologue / epilogue
          if (best < 0 || cost < best) { best = cost; nchunks = z; }
        }
      }
      float* part = nullptr;
      if (nchunks > 1) AE_TRY(ctx->getT("gdiff_part", (size_t)2 * nchunks * nt * 64 * (T + 1), &part));
      dim3 grid(nt, 2, nchunks);
      // 5 x 5 kernels, many of them: both halves of the all-pairs sum as tcgen05 GEMMs (gdiff_tc.cu)
      int rc_tc = AEFFT_ERR_UNSUPPORTED;
      if (T == 25) rc_tc = launch_gradient_diff_tc(ctx, dM, dD, c, f, cd, fd, t0, nt, nchunks, part);
      if (rc_tc != AEFFT_OK && rc_tc != AEFFT_ERR_UNSUPPORTED) return rc_tc;
      float* xp = nullptr;
      const int npad = tiles * 64;
      if (rc_tc != AEFFT_OK) AE_TRY(ctx->getT("gdiff_pack", (size_t)2 * npad * (T + 3), &xp));
      if (rc_tc == AEFFT_OK) {
        // (launched and counted there)
      } else if (T == 25) {
        gradient_diff_pack_kernel<25><<<(2 * npad + 127) / 128, 128, 0, ctx->stream>>>(c, f, xp, dM, dD, npad);
        gradient_diff_tiled_kernel<25><<<grid, 256, 0, ctx->stream>>>(c, f, cd, fd, dM, dD, t0, part, xp);
        ctx->launches += 2;
      } else {
        gradient_diff_pack_kernel<9><<<(2 * npad + 127) / 128, 128, 0, ctx->stream>>>(c, f, xp, dM, dD, npad);
        gradient_diff_tiled_kernel<9><<<grid, 256, 0, ctx->stream>>>(c, f, cd, fd, dM, dD, t0, part, xp);
        ctx->launches += 2;
      }
      if (nchunks > 1) {
        const int total = 2 * nt * 64;
        if (T == 25) gradient_diff_finish_kernel<25><<<(total + 127) / 128, 128, 0, ctx->stream>>>(c, f, cd, fd, part, n, t0, nt, nchunks);
        else gradient_diff_finish_kernel<9><<<(total + 127) / 128, 128, 0, ctx->stream>>>(c, f, cd, fd, part, n, t0, nt, nchunks);
        ctx->launches++;
      }
    }
    if (rank == 0) {
      gradient_diff_bias_kernel<<<(dM + dD + 127) / 128, 128, 0, ctx->stream>>>(b, p, bd, pd, dM, dD);
      ctx->launches++;
    }
  } else if (rank == 0) {
    gradient_diff_kernel<<<n, 64, 0, ctx->stream>>>(c, f, b, p, cd, fd, bd, pd, dM, dD, T);
    ctx->launches++;
  }
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// y[i] = a*y[i] + b*x[i]
__global__ void axpby_kernel(float* __restrict__ y, const float* __restrict__ x, float a, float b, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * y[i] + b * x[i];
}
int launch_axpby(aefft_ctx* ctx, float* y, const float* x, float a, float b, long long n) {
  axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(y, x, a, b, n);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

int launch_fft_update(aefft_ctx* ctx, int dM, int dD, int Nk, int Nl, float* c, float* f, float* b, float* p,
                      const float* dck, const float* dfk, const float* db, const float* dp, float* Dc, float* Df, float* Db,
                      float* Dp, float del, int maxdiff, float* div_scratch /* 2*nC + dM + dD floats */) {
  const int T = Nk * Nl, nC = dM * dD * T;
  float *cd = nullptr, *fd = nullptr, *bd = nullptr, *pd = nullptr;
  if (maxdiff) {
    cd = div_scratch; fd = cd + nC; bd = fd + nC; pd = bd + dM;
    AE_TRY(launch_gradient_diff(ctx, dM, dD, Nk, Nl, c, f, b, p, div_scratch, 0, 1));
  }
  fft_update_kernel<<<(nC + 127) / 128, 128, 0, ctx->stream>>>(c, f, b, p, dck, dfk, db, dp, Dc, Df, Db, Dp, cd, fd, bd, pd, nC,
                                                                dM, dD, del, 1.f, 10.f);  // w0=1, w1=10 (:1252)
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// ------------------------------------------------------------------------------------------------ column slabs
// Frequency-bin sharding: device r of G owns the spectrum columns [col0, col0+ncols); slab[img][wx][l] = full[img][wx][col0+l]
__global__ void spec_slab_kernel(const float2* __restrict__ full, float2* __restrict__ slab, long long rows, int Nyr, int col0,
                                 int ncols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * ncols) return;
  const long long r = idx / ncols;
  const int l = (int)(idx - r * ncols);
  slab[idx] = full[r * Nyr + col0 + l];
}
int launch_spec_slab(aefft_ctx* ctx, int64_t n_img, int Nx, int Ny, const float2* full, float2* slab, int col0, int ncols) {
  const long long rows = (long long)n_img * Nx, total = rows * ncols;
  spec_slab_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(full, slab, rows, Ny / 2 + 1, col0, ncols);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// ------------------------------------------------------------------------------------------------ mse
// calc_mse + thrust::reduce (:480-498, 1178-1192): sum_w |Xt-O|^2 / n_w, n_w = dD*Nx*Ny halved for 0<j<Nyr-1.
// Deterministic two-stage reduction in double; *out = scale * total.
// sum over the owned bins of w_j |Xt - O|^2 with the Hermitian weights w_j = 1 on the DC / Nyquist columns, 2 elsewhere:
// = 2 * (flat sum over all bins)  -  (sum over the DC / Nyquist columns the device owns).  The flat sum needs no column
// index (the per-element 64-bit modulo made the first version compute bound at half the HBM rate); it reads 16 bytes per
// operand and thread and folds 8 bins in fp32 before every fp64 add.
__global__ void __launch_bounds__(256) spec_mse_kernel(const float2* __restrict__ Xt, const float2* __restrict__ O, long long total,
                                                       int ncols, int col0, int Nyr, double* __restrict__ part) {
  double s = 0.0;
  // two bins per 16-byte access (scalar path for a base that is only 8-byte aligned)
  const bool aligned = ((reinterpret_cast<uintptr_t>(Xt) | reinterpret_cast<uintptr_t>(O)) & 15) == 0;
  const long long pairs = aligned ? total >> 1 : 0;
  const float4* X4 = reinterpret_cast<const float4*>(Xt);
  const float4* O4 = reinterpret_cast<const float4*>(O);
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < pairs; i += 4 * stride) {
    float v = 0.f;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const float4 a = __ldg(X4 + i + u * stride), b = __ldg(O4 + i + u * stride);
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      v += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    s += (double)v;
  }
  for (; i < pairs; i += stride) {
    const float4 a = __ldg(X4 + i), b = __ldg(O4 + i);
    const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
    s += (double)((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
  }
  for (long long e = 2 * pairs + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const float2 a = Xt[e], b = O[e];
    const float dx = a.x - b.x, dy = a.y - b.y;
    s += (double)(dx * dx + dy * dy);
  }
  s *= 2.0;
  // correction: the DC and Nyquist columns count once (only where this device owns them)
  const long long rows = total / ncols;
  const bool own_dc = col0 == 0, own_ny = col0 + ncols == Nyr;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    if (own_dc) {
      const float2 a = Xt[r * ncols], b = O[r * ncols];
      const float dx = a.x - b.x, dy = a.y - b.y;
      s -= (double)(dx * dx + dy * dy);
    }
    if (own_ny && (ncols > 1 || !own_dc)) {
      const float2 a = Xt[r * ncols + ncols - 1], b = O[r * ncols + ncols - 1];
      const float dx = a.x - b.x, dy = a.y - b.y;
      s -= (double)(dx * dx + dy * dy);
    }
  }
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
__global__ void spec_mse_final_kernel(const double* __restrict__ part, int n, double scale, float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)(red[0] * scale);
}

int launch_spec_mse(aefft_ctx* ctx, int64_t B, int dD, int dM, int Nx, int Ny, const float2* Xt, const float2* O, float* out,
                    int col0, int ncols) {
  const int Nyr = Ny / 2 + 1;
  if (ncols <= 0) ncols = Nyr;
  const long long total = (long long)B * dD * Nx * ncols;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4 * ctx->sm_count) blocks = 4 * ctx->sm_count;
  double* part;
  AE_TRY(ctx->getT("mse_part", (size_t)blocks, &part));
  ProfScope prof(ctx, "spec_mse", 0.0, 16.0 * total);
  spec_mse_kernel<<<blocks, 256, 0, ctx->stream>>>(Xt, O, total, ncols, col0, Nyr, part);
  // per frame: [sum |.|^2 / (dD Nx Ny)] / (2 dM Nx Ny); mean over frames
  const double scale = 1.0 / ((double)dD * Nx * Ny) / (2.0 * dM * Nx * Ny) / (double)B;
  spec_mse_final_kernel<<<1, 256, 0, ctx->stream>>>(part, blocks, scale, out);
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// interleaved-float wire format of net_cfreq (copy_out / copy_in, :246-282) is bit-identical to complex64: the
// "conversion" is a plain copy, done with cudaMemcpyAsync by the callers.

}  // namespace aefft
