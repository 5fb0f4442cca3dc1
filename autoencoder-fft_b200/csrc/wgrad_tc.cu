// Weight-gradient ("correlation") contractions of one layer pair on the 5th-generation tensor cores.
//
//   GC[m][d][k][l] = sum_{b,q} dh[b][m](q)       * in[b][d](q + s_t)          s_t = (ai0+tk, aj0+tl) of the forward window
//   GF[d][m][k][l] = sum_{b,q} (out-in)[b][d](q) * hin[b][m](q + s_t)  =  sum_{b,q} hin[b][m](q) * (out-in)[b][d](q - s_t)
//
// Both are the SAME job shape: an unshifted "M-side" operand with dM channels (dh or hin) against a shifted "N-side"
// operand with a chunk of <= 16 of the dD channels (in, shifted by +s_t; or e = out-in, shifted by -s_t, i.e. the
// mirrored window).  GEMM view per tap:  D_t[m][x] += sum_pixels  Mop[pixel][m] * Nop[pixel + shift_t][x]
//   M = 64 (channels m, TMEM rows), K = 16 pixels per MMA, and N = NL taps x 8 channels: the N-side chunk is ONE
//   8-channel plane, and consecutive 8-column groups of N are the same plane advanced by one pixel (SBO = 16 bytes),
//   so one MMA covers a whole window row (tl = 0..NL-1).  NK accumulators of 8*NL columns live in TMEM (200 of 256
//   columns for 5x5) and are accumulated over ALL tiles a CTA processes; warp tk issues the MMAs of window row tk.
// Operands are staged in shared memory as 8-channel planes [plane][linear pixel][8 x bf16] (16 bytes per pixel), the
// SWIZZLE_NONE MN-major canonical layout (K = pixel rows at 16-byte pitch, 8-row core matrices contiguous), so the tap
// shift is again just a start-address offset on the N-side descriptor: 25 taps reuse one staged tile.
// Linear pixel = tile-row * PJ + column with PJ = Ny + NL - 1; the M-side is zero in the NL-1 pad columns, which makes
// the products with wrapped-around N-side pixels vanish.
// fp32 parity: bf16 hi/lo split of both operands, three products per step (BF16X3), fp32 accumulation.
// Per-CTA partial sums go to global memory and are summed in fixed order by reduce_partials (deterministic).
#include "common.cuh"
#include "staging.cuh"
#include "umma.cuh"

namespace aefft {

using namespace umma;

constexpr int WT_THREADS = 256;
constexpr int WT_N = 8;   // N-side channels per job (one plane)

struct WgJob {
  const float* m0;   // M-side operand [B][dM][Nx][Ny]
  const float* n0;   // N-side operand [B][dD][Nx][Ny]
  const float* n1;   // optional: N-side = n0 - n1
  int nch0;          // first N-side channel of this job's chunk
  int oi, oj;        // N-side halo-grid origin relative to the tile origin
  int rev;           // taps enumerated mirrored (GF)
  int is_gf;         // output index order: GF [d][m][k][l], else GC [m][d][k][l]
  long long g_off;   // offset of this gradient inside one partial block
};

struct WgradTcParams {
  WgJob job[8];
  int n_jobs, ctas_per_job;
  float* part;        // [ctas_per_job][n_out]
  long long n_out;
  int dM, dD, Nx, Ny, NK, NL, T;
  int PJ, TI, KQ, HPn, MPl;  // MPl: M-side planes (dM/8)
  int tiles_per_frame;
  long long n_tiles;
  int passes, flip;
  uint32_t tmem_cols;
  uint32_t m_plane, n_plane;  // bytes
};

// Stage `nplanes` 8-channel planes of a planar fp32 tensor (optionally a difference) as bf16 hi/lo into shared memory.
// Grid pixel h = r*PJ + c maps to image pixel (row0 + r, col0 + c); pixels outside the image, rows >= rows_valid and
// channels >= nch are zero.
__device__ __forceinline__ void stage_planes(unsigned char* hi_base, unsigned char* lo_base, uint32_t plane_bytes, int nplanes,
                                             int npx, int PJ, const float* __restrict__ s0, const float* __restrict__ s1,
                                             int ch0, int nch, long long plane, int Nx, int Ny, int row0, int col0,
                                             int rows_valid, int cols_valid) {
  for (int idx = threadIdx.x; idx < nplanes * npx; idx += WT_THREADS) {
    const int pl = idx / npx, h = idx - pl * npx;
    const int r = h / PJ, c = h - r * PJ;
    const int si = row0 + r, sj = col0 + c;
    const bool inb = r < rows_valid && c < cols_valid && si >= 0 && si < Nx && sj >= 0 && sj < Ny;
    const long long pix = inb ? (long long)si * Ny + sj : 0;
    const int c0 = pl * 8;
    float v[8], u[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const int cc = min(c0 + e, nch - 1);
      v[e] = __ldg(s0 + (long long)(ch0 + cc) * plane + pix);
    }
    if (s1) {
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int cc = min(c0 + e, nch - 1);
        u[e] = __ldg(s1 + (long long)(ch0 + cc) * plane + pix);
      }
#pragma unroll
      for (int e = 0; e < 8; e++) v[e] -= u[e];
    }
    __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      if (!inb || c0 + e >= nch) v[e] = 0.f;
      split_bf16(v[e], hi[e], lo[e]);
    }
    *reinterpret_cast<uint4*>(hi_base + (size_t)pl * plane_bytes + (size_t)h * 16) =
        make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
    *reinterpret_cast<uint4*>(lo_base + (size_t)pl * plane_bytes + (size_t)h * 16) =
        make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
  }
}

__global__ void __launch_bounds__(WT_THREADS, 2) wgrad_tc_kernel(WgradTcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  // [M hi: MPl planes][M lo: MPl planes][N hi: 2 planes][N lo: 2 planes] ... the M=64 MMA addresses 8 M-side planes from
  // each base; planes beyond MPl alias whatever follows (their TMEM rows are never read), the host sizes the
  // allocation so that those reads stay inside it.
  unsigned char* Mhi = smem;
  unsigned char* Mlo = Mhi + (size_t)p.MPl * p.m_plane;
  unsigned char* Nhi = Mlo + (size_t)p.MPl * p.m_plane;
  unsigned char* Nlo = Nhi + (size_t)p.n_plane;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int jb = blockIdx.y, cta = blockIdx.x;
  const WgJob& J = p.job[jb];
  if (warp == 0) tmem_alloc(&tmem_slot, p.tmem_cols);
  if (tid == 32) {
    mbar_init(&bar, p.NK);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  uint32_t phase = 0;
  const uint32_t idesc = make_idesc_bf16(64, WT_N * p.NL, 1, 1);
  const long long plane = (long long)p.Nx * p.Ny;
  const int nch = min(WT_N, p.dD - J.nch0);
  bool first = true;

  for (long long tile = cta; tile < p.n_tiles; tile += p.ctas_per_job) {
    const long long b = tile / p.tiles_per_frame;
    const int i0 = (int)(tile % p.tiles_per_frame) * p.TI;
    const int rows_valid = min(p.TI, p.Nx - i0);
    // M-side on the output grid (zero in the pad columns and beyond the tile), N-side on its halo grid
    stage_planes4<WT_THREADS>(Mhi, Mlo, p.m_plane, p.MPl, p.KQ, p.PJ, J.m0 + b * p.dM * plane, nullptr, 1.f, 0, p.dM, plane,
                              p.Nx, p.Ny, i0, 0, rows_valid, p.Ny, 0, tid);
    stage_planes4<WT_THREADS>(Nhi, Nlo, p.n_plane, 1, p.HPn, p.PJ, J.n0 + b * p.dD * plane,
                              J.n1 ? J.n1 + b * p.dD * plane : nullptr, 1.f, J.nch0, nch, plane, p.Nx, p.Ny, i0 + J.oi, J.oj,
                              1 << 30, 1 << 30, 0, tid);
    fence_proxy_async();
    __syncthreads();
    // ---- NK issuing threads (lane 0 of warps 0..NK-1): warp tk owns window row tk and its accumulator ----
    if (warp < p.NK) {  // all lanes run the loop (uniform descriptor arithmetic); lane 0 issues
      fence_after_sync();
      const bool leader = lane == 0;
      const uint64_t m_hi0 = make_desc(smem_u32(Mhi), 128, p.m_plane), m_lo0 = make_desc(smem_u32(Mlo), 128, p.m_plane);
      const uint64_t n_hi0 = make_desc(smem_u32(Nhi), 128, 16), n_lo0 = make_desc(smem_u32(Nlo), 128, 16);
      const int ksteps = p.KQ / 16;
      const uint32_t d = tmem_base + (uint32_t)(warp * WT_N * p.NL);
      const uint64_t shift = (uint64_t)(warp * p.PJ);
      for (int ks = 0; ks < ksteps; ks++) {
        const uint64_t q = (uint64_t)(ks * 16);
        if (leader) {
          mma_bf16(d, m_hi0 + q, n_hi0 + q + shift, idesc, !(first && ks == 0));
          if (p.passes == 3) {
            mma_bf16(d, m_hi0 + q, n_lo0 + q + shift, idesc, true);
            mma_bf16(d, m_lo0 + q, n_hi0 + q + shift, idesc, true);
          }
        }
      }
      if (leader) commit(&bar);
      __syncwarp();
    }
    first = false;
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
  }
  // ---- epilogue: M=64 accumulator layout: row m lives in TMEM lane (m % 16) + 32 * (m / 16), column = x ----
  if (warp < 4) {
    float* part = p.part + (long long)cta * p.n_out + J.g_off;
    const int m = warp * 16 + lane;
    const int TT = p.NK * p.NL;
    const int ncols = p.T * WT_N;  // column = (tk*NL + tl)*8 + x
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      if (lane < 16 && m < p.dM) {
#pragma unroll
        for (int e = 0; e < 16; e++) {
          const int col = c0 + e, t = col >> 3, x = col & 7;
          if (col < ncols && x < nch) {
            int tk = t / p.NL, tl = t - tk * p.NL;
            if (J.rev) { tk = p.NK - 1 - tk; tl = p.NL - 1 - tl; }
            const int k = p.flip ? p.NK - 1 - tk : tk, l = p.flip ? p.NL - 1 - tl : tl;
            const int d = J.nch0 + x;
            const long long gi = J.is_gf ? (((long long)d * p.dM + m) * TT + k * p.NL + l)
                                         : (((long long)m * p.dD + d) * TT + k * p.NL + l);
            part[gi] = v[e];
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, p.tmem_cols);
}

// Per-channel sums of a planar tensor (optionally a difference) and the sum of squares, in double, deterministic:
//   sum[c] = sum_{b,pix} (a0 - a1)[b][c](pix) ;  sumsq = sum (a0 - a1)^2          (bias gradients and the printed mse)
__global__ void channel_sums_kernel(const float* __restrict__ a0, const float* __restrict__ a1, long long B, int ch,
                                    long long plane, int nsplit, double* __restrict__ part_sum, double* __restrict__ part_sq) {
  // block (c, sp): frame b = sp / per_frame, segment seg = sp % per_frame of that frame's channel plane
  const int c = blockIdx.x, sp = blockIdx.y;
  const int per_frame = nsplit / (int)B;
  const long long b = sp / per_frame;
  const int seg = sp - (int)b * per_frame;
  const long long lo = plane * seg / per_frame, hi = plane * (seg + 1) / per_frame;
  const float* p0 = a0 + (b * ch + c) * plane;
  const float* p1 = a1 ? a1 + (b * ch + c) * plane : nullptr;
  float s = 0.f, q = 0.f;  // short per-thread runs in fp32, everything across threads / blocks in double
  double ds = 0.0, dq = 0.0;
  int run = 0;
  for (long long n = lo + threadIdx.x; n < hi; n += blockDim.x) {
    float v = __ldg(p0 + n);
    if (p1) v -= __ldg(p1 + n);
    s += v;
    q = fmaf(v, v, q);
    if (++run == 16) { ds += (double)s; dq += (double)q; s = 0.f; q = 0.f; run = 0; }
  }
  ds += (double)s;
  dq += (double)q;
  __shared__ double rs[256], rq[256];
  rs[threadIdx.x] = ds;
  rq[threadIdx.x] = dq;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) { rs[threadIdx.x] += rs[threadIdx.x + h]; rq[threadIdx.x] += rq[threadIdx.x + h]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part_sum[(long long)c * nsplit + sp] = rs[0];
    part_sq[(long long)c * nsplit + sp] = rq[0];
  }
}
__global__ void channel_sums_final_kernel(const double* __restrict__ part_sum, const double* __restrict__ part_sq, int ch,
                                          int nsplit, float* __restrict__ sum, float* __restrict__ sumsq) {
  const int c = threadIdx.x;
  __shared__ double sq[256];
  double s = 0.0, q = 0.0;
  if (c < ch)
    for (int i = 0; i < nsplit; i++) { s += part_sum[(long long)c * nsplit + i]; q += part_sq[(long long)c * nsplit + i]; }
  if (c < ch && sum) sum[c] = (float)s;
  sq[threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x == 0 && sumsq) {
    double t = 0.0;
    for (int i = 0; i < ch && i < 256; i++) t += sq[i];
    *sumsq = (float)t;
  }
}

int launch_channel_sums(aefft_ctx* ctx, int64_t B, int ch, int Nx, int Ny, const float* a0, const float* a1, float* sum,
                        float* sumsq) {
  AE_ARG(ch > 0 && ch <= 256);
  // enough blocks to fill the machine: every frame's channel plane is cut into per_frame segments
  int per_frame = (int)((4LL * ctx->sm_count + (long long)ch * B - 1) / ((long long)ch * B));
  if (per_frame < 1) per_frame = 1;
  AE_ARG(B * per_frame <= 65535);
  const int nsplit = (int)B * per_frame;
  double* part;
  AE_TRY(ctx->getT("chsum_part", (size_t)2 * ch * nsplit, &part));
  const double px = (double)B * Nx * Ny;
  ProfScope prof(ctx, "channel_sums", 0.0, 4.0 * px * ch * (a1 ? 2 : 1));
  channel_sums_kernel<<<dim3(ch, nsplit), 256, 0, ctx->stream>>>(a0, a1, B, ch, (long long)Nx * Ny, nsplit, part,
                                                                part + (size_t)ch * nsplit);
  channel_sums_final_kernel<<<1, 256, 0, ctx->stream>>>(part, part + (size_t)ch * nsplit, ch, nsplit, sum, sumsq);
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

__global__ void reduce_parts_kernel(const float* __restrict__ part, int n_parts, long long n, float* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  double s = 0.0;
  for (int c = 0; c < n_parts; c++) s += (double)part[(long long)c * n + idx];
  out[idx] = (float)s;
}

// GC (dM*dD*T floats) and GF (dD*dM*T floats) of one pair, raw sums over the B frames, written to G (GC then GF).
// Returns AEFFT_ERR_UNSUPPORTED (no error text) outside the kernel's envelope.
int launch_wgrad_tc(aefft_ctx* ctx, const Window& win, int64_t B, int dD, int dM, int Nx, int Ny, const float* in,
                    const float* out, const float* hin, const float* dh, float* G, int passes) {
  const int T = win.Nk * win.Nl;
  const int ncols = T * WT_N;
  if (dM % 8 != 0 || dM < 8 || dM > 64 || ncols > 512 || win.Nk > 8 || WT_N * win.Nl > 256 || win.lo != 0)
    return AEFFT_ERR_UNSUPPORTED;
  const int n_chunks = (dD + WT_N - 1) / WT_N;
  if (2 * n_chunks > 8) return AEFFT_ERR_UNSUPPORTED;
  const int PJ = Ny + win.Nl - 1;
  const int halo = (win.Nk - 1) * PJ + win.Nl + 8;
  const int MPl = dM / 8;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < ncols + 16) tmem_cols <<= 1;  // +16: the last 16-column epilogue load may overhang
  if (tmem_cols > 512) return AEFFT_ERR_UNSUPPORTED;
  const int ctas = tmem_cols <= 256 ? 2 : 1;           // co-resident CTAs per SM
  const size_t budget = ctas == 2 ? 110 * 1024 : 220 * 1024;
  // largest tile (rows) whose staging fits in shared memory
  int TI = Nx, KQ = 0, HPn = 0;
  size_t smem = 0;
  for (; TI >= 1; TI--) {
    KQ = (TI * PJ + 15) / 16 * 16;
    HPn = (KQ + halo + 7) / 8 * 8;
    const size_t m_plane = (size_t)KQ * 16, n_plane = (size_t)HPn * 16;
    smem = 2 * MPl * m_plane + 2 * n_plane;
    const size_t span = (size_t)(MPl + 8) * m_plane;  // reach of the junk planes of the "lo" descriptor
    if (span > smem) smem = span;
    if (smem <= budget && m_plane < 262000 && n_plane < 262000) break;
  }
  if (TI < 1) return AEFFT_ERR_UNSUPPORTED;
  WgradTcParams p;
  p.dM = dM; p.dD = dD; p.Nx = Nx; p.Ny = Ny; p.NK = win.Nk; p.NL = win.Nl; p.T = T;
  p.PJ = PJ; p.TI = TI; p.KQ = KQ; p.HPn = HPn; p.MPl = MPl;
  p.m_plane = (uint32_t)KQ * 16; p.n_plane = (uint32_t)HPn * 16;
  p.tiles_per_frame = (Nx + TI - 1) / TI;
  p.n_tiles = (long long)B * p.tiles_per_frame;
  p.passes = passes; p.flip = win.flip;
  p.tmem_cols = tmem_cols;
  p.n_jobs = 2 * n_chunks;
  int cpj = ctas * ctx->sm_count / p.n_jobs;
  if (cpj < 1) cpj = 1;
  if (cpj > p.n_tiles) cpj = (int)p.n_tiles;
  p.ctas_per_job = cpj;
  const long long nC = (long long)dM * dD * T;
  p.n_out = 2 * nC;
  for (int c = 0; c < n_chunks; c++) {
    WgJob& gc = p.job[c];
    gc.m0 = dh; gc.n0 = in; gc.n1 = nullptr; gc.nch0 = c * WT_N;
    gc.oi = win.ai0; gc.oj = win.aj0; gc.rev = 0; gc.is_gf = 0; gc.g_off = 0;
    WgJob& gf = p.job[n_chunks + c];
    gf.m0 = hin; gf.n0 = out; gf.n1 = in; gf.nch0 = c * WT_N;
    // e shifted by -s_t: mirrored taps, origin -(ai0 + NK-1), -(aj0 + NL-1)
    gf.oi = -(win.ai0 + win.Nk - 1); gf.oj = -(win.aj0 + win.Nl - 1); gf.rev = 1; gf.is_gf = 1; gf.g_off = nC;
  }
  float* part;
  AE_TRY(ctx->getT("wgtc_part", (size_t)cpj * p.n_out, &part));
  p.part = part;
  AE_TRY(ctx->ensure_dyn_smem((const void*)wgrad_tc_kernel, smem));
  {
    const double px = (double)B * Nx * Ny;
    ProfScope prof(ctx, "wgrad_tc", 2.0 * 2.0 * px * dM * dD * T, 4.0 * px * (3.0 * dD + 2.0 * dM));
    wgrad_tc_kernel<<<dim3(cpj, p.n_jobs), WT_THREADS, smem, ctx->stream>>>(p);
  }
  reduce_parts_kernel<<<(unsigned)((p.n_out + 255) / 256), 256, 0, ctx->stream>>>(part, cpj, p.n_out, G);
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

}  // namespace aefft
