// Elementwise / small kernels of the coordinate path: Pool, Portion, synthetic frames, border sums, the
// bug-compatible dF of backprop_gpu (quirks C3/C4) and the fused clip+momentum weight update of all three modes.
#include <cstring>

#include "common.cuh"

namespace aefft {

// ---------------------------------------------------------------------------------------------- Pool
// netlib.cpp:114-164.  scale>0: out[i/s][j/s] = (int) max(0, window max) (the reference accumulates through
// `int smax=0`, :127-136); scale<0: out[a][b] = in[a/s][b/s].  One thread per OUTPUT element, j fastest.
__global__ void pool_down_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int Nx,
                                 int Ny, int oNx, int oNy, int s) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = planes * oNx * oNy;
  if (n >= total) return;
  int oj = n % oNy;
  int oi = (n / oNy) % oNx;
  long long pl = n / ((long long)oNy * oNx);
  const float* src = in + pl * Nx * Ny;
  int smax = 0;
  int i0 = oi * s, j0 = oj * s;
  if (i0 < Nx && j0 < Ny) {
    for (int k = 0; k < s; k++)
      for (int l = 0; l < s; l++)
        if (i0 + k < Nx && j0 + l < Ny) {
          float v = __ldg(src + (long long)(i0 + k) * Ny + j0 + l);
          if (v > (float)smax) smax = (int)v;
        }
    out[n] = (float)smax;
  } else {
    out[n] = 0.f;  // never written by the reference (caller sizes are divisible)
  }
}

// One thread per 4 consecutive OUTPUT pixels of a row (128-bit store) when the row length allows it.
__global__ void pool_up_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int Nx,
                               int Ny, int oNx, int oNy, int s, int vec) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = vec ? oNy / 4 : oNy;
  long long total = planes * oNx * w;
  if (n >= total) return;
  int oj = (int)(n % w) * (vec ? 4 : 1);
  int oi = (n / w) % oNx;
  long long pl = n / ((long long)w * oNx);
  const float* src = in + pl * Nx * Ny + (long long)min(oi / s, Nx - 1) * Ny;
  float* dst = out + (pl * oNx + oi) * oNy + oj;
  if (vec) {
    float4 v;
    v.x = __ldg(src + min(oj / s, Ny - 1));
    v.y = __ldg(src + min((oj + 1) / s, Ny - 1));
    v.z = __ldg(src + min((oj + 2) / s, Ny - 1));
    v.w = __ldg(src + min((oj + 3) / s, Ny - 1));
    *reinterpret_cast<float4*>(dst) = v;
  } else {
    *dst = __ldg(src + min(oj / s, Ny - 1));
  }
}


// Fast paths for the shipped pooling factor 2 (New_Layer_Param.txt) on frames whose row length is a multiple of 8
// (down) / 4 (up): 128-bit accesses, 32-bit index arithmetic, every input byte read exactly once.
// down: one thread = 4 output pixels of one row <- 2 input rows x 8 pixels
__global__ void pool_down2_kernel(const float4* __restrict__ in, float4* __restrict__ out, unsigned total, int oNx, int w4,
                                  int in_row4) {
  const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= total) return;
  const unsigned q = n % (unsigned)w4, r = n / (unsigned)w4;          // r = plane * oNx + oi
  const unsigned pl = r / (unsigned)oNx, oi = r - pl * (unsigned)oNx;
  const size_t src = ((size_t)pl * oNx * 2 + 2 * oi) * in_row4 + 2 * q;  // float4 units; input plane has 2*oNx rows
  const float4 a0 = __ldg(in + src), a1 = __ldg(in + src + 1);
  const float4 b0 = __ldg(in + src + in_row4), b1 = __ldg(in + src + in_row4 + 1);
  auto red = [](float x, float y, float z, float w) {
    int smax = 0;  // netlib.cpp:127-136: running int maximum, scan order (k,l)
    if (x > (float)smax) smax = (int)x;
    if (y > (float)smax) smax = (int)y;
    if (z > (float)smax) smax = (int)z;
    if (w > (float)smax) smax = (int)w;
    return (float)smax;
  };
  float4 o;
  o.x = red(a0.x, a0.y, b0.x, b0.y);
  o.y = red(a0.z, a0.w, b0.z, b0.w);
  o.z = red(a1.x, a1.y, b1.x, b1.y);
  o.w = red(a1.z, a1.w, b1.z, b1.w);
  out[n] = o;
}
// up: one thread = 2 input pixels of one row -> 2 output rows x 4 pixels: every store instruction of a warp then writes
// 512 contiguous bytes (full 32-byte sectors); with 4 input pixels per thread the two 16-byte halves of a sector came
// from two different store instructions.
__global__ void pool_up2_kernel(const float2* __restrict__ in, float4* __restrict__ out, unsigned total, int Nx, int w2) {
  const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= total) return;
  const unsigned q = n % (unsigned)w2, r = n / (unsigned)w2;  // r = plane * Nx + i
  const float2 v = __ldg(in + n);
  const float4 o = make_float4(v.x, v.x, v.y, v.y);
  const size_t dst = (size_t)r * 2 * w2 + q;  // output row 2*(plane*Nx+i), w2 float4 per output row
  out[dst] = o;
  out[dst + w2] = o;
}

int launch_pool(aefft_ctx* ctx, int64_t B, int D, int Nx, int Ny, int oNx, int oNy, int scale, const float* in,
                float* out) {
  AE_ARG(scale != 0 && B > 0 && D > 0);
  long long planes = (long long)B * D;
  long long total = planes * oNx * oNy;
  unsigned blocks = (unsigned)((total + 255) / 256);
  ProfScope prof(ctx, scale > 0 ? "pool_down" : "pool_up", 0.0, 4.0 * (planes * (double)Nx * Ny + (double)total));
  const bool al16 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (scale == 2 && al16 && Nx == 2 * oNx && Ny == 2 * oNy && Ny % 8 == 0 && total / 4 < 0xffffffffLL) {
    const unsigned n4 = (unsigned)(total / 4);
    pool_down2_kernel<<<(n4 + 255) / 256, 256, 0, ctx->stream>>>(reinterpret_cast<const float4*>(in),
                                                                reinterpret_cast<float4*>(out), n4, oNx, oNy / 4, Ny / 4);
  } else if (scale == -2 && al16 && oNx == 2 * Nx && oNy == 2 * Ny && Ny % 4 == 0 && total / 4 < 0xffffffffLL) {
    const unsigned n2 = (unsigned)(planes * Nx * (Ny / 2));
    pool_up2_kernel<<<(n2 + 255) / 256, 256, 0, ctx->stream>>>(reinterpret_cast<const float2*>(in),
                                                              reinterpret_cast<float4*>(out), n2, Nx, Ny / 2);
  } else if (scale > 0) {
    pool_down_kernel<<<blocks, 256, 0, ctx->stream>>>(in, out, planes, Nx, Ny, oNx, oNy, scale);
  } else {
    const int vec = (oNy % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const long long items = vec ? total / 4 : total;
    pool_up_kernel<<<(unsigned)((items + 255) / 256), 256, 0, ctx->stream>>>(in, out, planes, Nx, Ny, oNx, oNy, -scale, vec);
  }
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// ---------------------------------------------------------------------------------------------- Portion
// netlib.cpp:292-315: centre crop by q.
__global__ void portion_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int Nx,
                               int Ny, int q) {
  int oNx = Nx / q, oNy = Ny / q;
  int dx = (Nx - oNx) / 2, dy = (Ny - oNy) / 2;
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = planes * oNx * oNy;
  if (n >= total) return;
  int j = n % oNy;
  int i = (n / oNy) % oNx;
  long long pl = n / ((long long)oNy * oNx);
  out[n] = __ldg(in + pl * Nx * Ny + (long long)(i + dx) * Ny + j + dy);
}

int launch_portion(aefft_ctx* ctx, int64_t B, int D, int Nx, int Ny, int q, const float* in, float* out) {
  AE_ARG(q >= 1 && B > 0 && D > 0);
  long long planes = (long long)B * D;
  long long total = planes * (Nx / q) * (Ny / q);
  portion_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(in, out, planes, Nx, Ny, q);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// ---------------------------------------------------------------------------------------------- synthetic frames
// pixel = float(splitmix64(seed, linear index of (b,d,i,j)) & 255); raw 0..255 like ImageToSpin_C
// (netlib.cpp:46-48).  Counter-based, so every rank generates its own frames without any transfer.
__global__ void synth_kernel(float* __restrict__ out, unsigned long long seed, unsigned long long first,
                             long long total) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= total) return;
  unsigned long long z = (first + (unsigned long long)n) + seed * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  out[n] = (float)(z & 255ULL);
}

int launch_synth(aefft_ctx* ctx, uint64_t seed, int64_t b0, int64_t B, int D, int Nx, int Ny, float* out) {
  AE_ARG(B > 0 && D > 0 && Nx > 0 && Ny > 0);
  long long per = (long long)D * Nx * Ny;
  long long total = per * B;
  synth_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(out, seed, (unsigned long long)(b0 * per),
                                                                       total);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// ---------------------------------------------------------------------------------------------- border sums
// T[d][k1][l1] = sum_{b,i,j} (out-in)[b][d](i,j) * [lo <= i-(bi+k1) < Nx][lo <= j-(bj+l1) < Ny]
// Needed by quirk C1 (backproplib.cu:220: the bias gradient keeps only d1=dD-1, and its inner tap sum runs over the
// taps whose hidden position is inside the image).  One CTA per (d,k1,l1); deterministic tree reduction.
__global__ void border_sums_kernel(const float* __restrict__ out, const float* __restrict__ in, float* __restrict__ T,
                                   long long B, int D, int Nx, int Ny, int Nk, int Nl, int bi, int bj, int lo) {
  const int o = blockIdx.x;
  const int l1 = o % Nl, k1 = (o / Nl) % Nk, d = o / (Nl * Nk);
  const int si = bi + k1, sj = bj + l1;
  const long long plane = (long long)Nx * Ny;
  double s = 0.0;
  for (long long n = threadIdx.x; n < B * plane; n += blockDim.x) {
    long long b = n / plane;
    int i = (int)((n % plane) / Ny), j = (int)(n % Ny);
    int hi = i - si, hj = j - sj;
    if (hi < lo || hi >= Nx || hj < lo || hj >= Ny) continue;
    long long off = (b * D + d) * plane + (long long)i * Ny + j;
    s += (double)(__ldg(out + off) - __ldg(in + off));
  }
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) T[o] = (float)red[0];
}

int launch_border_sums(aefft_ctx* ctx, int64_t B, int D, int Nx, int Ny, int Nk, int Nl, int bi, int bj, int lo,
                       const float* out, const float* in, float* T) {
  border_sums_kernel<<<D * Nk * Nl, 256, 0, ctx->stream>>>(out, in, T, B, D, Nx, Ny, Nk, Nl, bi, bj, lo);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// ---------------------------------------------------------------------------------------------- quirk dF (C3 + C4)
// backprop_gpu's dF term (gradient_CFBP :224-227, gradient_CF :281-284), bug-compatible, summed over frames:
//   launch t=(m,d,k,l) writes buf(i,j) = e[d](i,j) * hin_flat[m*P + (i-ik)*Ny + (j-jk)]  where (i-ik, j-il) is inside
//   the image; jk = il for (k,l)==(0,0), and jk = ik otherwise when C3 is on (:283 uses (j-ik));
//   buf is NOT re-zeroed between launches (C4), so pixels outside the mask keep the value of the most recent
//   launch that wrote them; gF[d][m][k][l] = sum(buf) after launch t.
// Only square frames are defined for the compiled reference (stride quirk C2); the flat read uses stride Ny.
// Part 1: the in-mask sum per launch (one CTA per (d,m,k,l)).
__global__ void quirk_dF_main_kernel(const float* __restrict__ out, const float* __restrict__ in,
                                     const float* __restrict__ hin, float* __restrict__ gF, long long B, int dD, int dM,
                                     int Nx, int Ny, int Nk, int Nl, int bi, int bj, int c3) {
  const int o = blockIdx.x;  // [d][m][k][l]
  const int l = o % Nl, k = (o / Nl) % Nk, m = (o / (Nl * Nk)) % dM, d = o / (Nl * Nk * dM);
  const int ik = bi + k, il = bj + l;
  const int jk = (c3 && !(k == 0 && l == 0)) ? ik : il;
  const long long plane = (long long)Nx * Ny;
  const long long hin_frame = (long long)dM * plane;
  double s = 0.0;
  for (long long n = threadIdx.x; n < B * plane; n += blockDim.x) {
    long long b = n / plane;
    int i = (int)((n % plane) / Ny), j = (int)(n % Ny);
    if (i - ik < 0 || i - ik >= Nx || j - il < 0 || j - il >= Ny) continue;
    long long flat = (long long)m * plane + (long long)(i - ik) * Ny + (j - jk);
    if (flat < 0 || flat >= hin_frame) continue;  // the reference would read outside its buffer here
    long long off = (b * dD + d) * plane + (long long)i * Ny + j;
    s += (double)(__ldg(out + off) - __ldg(in + off)) * (double)__ldg(hin + b * hin_frame + flat);
  }
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) gF[o] = (float)red[0];
}

// Part 2 (C4): stale border values.  One thread per (frame, border pixel): walks the launches in the reference's
// order (m; d; k; l), carrying the value last written at its pixel, and adds it to every launch whose mask excludes
// the pixel.  Border pixels = those excluded by at least one tap's mask.
__global__ void quirk_dF_stale_kernel(const float* __restrict__ out, const float* __restrict__ in,
                                      const float* __restrict__ hin, double* __restrict__ acc, long long B, int dD,
                                      int dM, int Nx, int Ny, int Nk, int Nl, int bi, int bj, int c3, int n_border,
                                      int ilo, int ihi, int jlo, int jhi) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = n < B * n_border;
  const long long plane = (long long)Nx * Ny;
  const long long hin_frame = (long long)dM * plane;
  long long b = 0;
  int i = 0, j = 0;
  if (active) {
    // border pixel q of the frame, enumerated without a list: full rows i < ilo, then the column bands j < jlo / j >= jhi
    // of the rows [ilo, ihi), then full rows i >= ihi (the host passes ilo == ihi == Nx / jlo == jhi == Ny when every
    // row / column is excluded by some tap)
    b = n / n_border;
    int q = (int)(n % n_border);
    const int n_top = ilo * Ny, per_mid = jlo + (Ny - jhi), n_mid = (ihi - ilo) * per_mid;
    if (q < n_top) {
      i = q / Ny; j = q - i * Ny;
    } else if (q < n_top + n_mid) {
      q -= n_top;
      const int r = q / per_mid, t = q - r * per_mid;
      i = ilo + r; j = t < jlo ? t : jhi + (t - jlo);
    } else {
      q -= n_top + n_mid;
      const int r = q / Ny;
      i = ihi + r; j = q - r * Ny;
    }
  }
  float last = 0.f;
  const int lane = threadIdx.x & 31;
  for (int m = 0; m < dM; m++)
    for (int d = 0; d < dD; d++) {
      float e = 0.f;
      if (active) {
        long long off = (b * dD + d) * plane + (long long)i * Ny + j;
        e = __ldg(out + off) - __ldg(in + off);
      }
      for (int k = 0; k < Nk; k++)
        for (int l = 0; l < Nl; l++) {
          const int ik = bi + k, il = bj + l;
          float contrib = 0.f;
          if (active) {
            bool inmask = !(i - ik < 0 || i - ik >= Nx || j - il < 0 || j - il >= Ny);
            if (inmask) {
              const int jk = (c3 && !(k == 0 && l == 0)) ? ik : il;
              long long flat = (long long)m * plane + (long long)(i - ik) * Ny + (j - jk);
              last = (flat >= 0 && flat < hin_frame) ? e * __ldg(hin + b * hin_frame + flat) : 0.f;
            } else {
              contrib = last;
            }
          }
#pragma unroll
          for (int sft = 16; sft > 0; sft >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, sft);
          // fp64 accumulator: the order in which the warps arrive changes the sum by ~1e-16 relative, far below the fp32
          // rounding of the result (float atomics made the last bit of gF depend on the block schedule)
          if (lane == 0 && contrib != 0.f) atomicAdd(acc + (((long long)d * dM + m) * Nk + k) * Nl + l, (double)contrib);
        }
    }
}

__global__ void quirk_dF_fold_kernel(const double* __restrict__ acc, float* __restrict__ gF, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gF[i] = (float)((double)gF[i] + acc[i]);
}

int launch_quirk_dF(aefft_ctx* ctx, int quirks, int64_t B, int dD, int dM, int Nx, int Ny, int Nk, int Nl,
                    const float* out, const float* in, const float* hin, float* gF) {
  const int bi = tap_base(Nk, AEFFT_CONV_CUDA), bj = tap_base(Nl, AEFFT_CONV_CUDA);
  const int c3 = (quirks & AEFFT_QUIRK_C3) ? 1 : 0;
  quirk_dF_main_kernel<<<dD * dM * Nk * Nl, 256, 0, ctx->stream>>>(out, in, hin, gF, B, dD, dM, Nx, Ny, Nk, Nl, bi, bj,
                                                                  c3);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  if (quirks & AEFFT_QUIRK_C4) {
    // border pixels = those excluded by the mask of at least one tap; the kernel enumerates them from the four bounds
    // (no host-built list: a pinned staging buffer rewritten per pair raced with the previous pair's async copy)
    int ilo = bi + Nk - 1 > 0 ? bi + Nk - 1 : 0;    // i < ilo is excluded by the largest ik
    int ihi = bi < 0 ? Nx + bi : Nx;                // i >= ihi is excluded by the smallest ik
    int jlo = bj + Nl - 1 > 0 ? bj + Nl - 1 : 0;
    int jhi = bj < 0 ? Ny + bj : Ny;
    if (ilo > Nx) ilo = Nx;
    if (jlo > Ny) jlo = Ny;
    if (ihi <= ilo) ilo = ihi = Nx;                 // every row is a border row
    if (jhi <= jlo) jlo = jhi = Ny;                 // every column is a border column
    const long long nb = (long long)ilo * Ny + (long long)(ihi - ilo) * (jlo + Ny - jhi) + (long long)(Nx - ihi) * Ny;
    if (nb > 0) {
      AE_ARG(nb <= 0x7fffffffLL);
      long long total = (long long)B * nb;
      const int nG = dD * dM * Nk * Nl;
      double* acc;
      AE_TRY(ctx->getT("quirk_acc", (size_t)nG, &acc));
      AE_CUDA(cudaMemsetAsync(acc, 0, (size_t)nG * sizeof(double), ctx->stream));
      quirk_dF_stale_kernel<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(out, in, hin, acc, B, dD, dM, Nx,
                                                                                    Ny, Nk, Nl, bi, bj, c3, (int)nb, ilo,
                                                                                    ihi, jlo, jhi);
      quirk_dF_fold_kernel<<<(nG + 255) / 256, 256, 0, ctx->stream>>>(acc, gF, nG);
      ctx->launches += 2;
      AE_CUDA(cudaGetLastError());
    }
  }
  return AEFFT_OK;
}

// ---------------------------------------------------------------------------------------------- weight update
// clip kappa(g) = g / max(10,|g|)  (netlib.cpp:437; backproplib.cu:393)
__device__ __forceinline__ float clip10(float g) { return g / fmaxf(10.f, fabsf(g)); }

// CUDA_REF / CUDA_REF_SYM (backproplib.cu:387-412 / :616-641):  v = (1-alpha)*del*clip(g) + alpha*v; w -= v; dd = g
// (adapt_rate only records dd = g, its rate is overwritten by del = delmax, :34).
// gbuf = [GC dM*dD*T | GF dD*dM*T | GB dM | GP dD | SQ 1] raw sums; g = sum * inv_norm.
// With quirk C1 the GB slot already holds the bug-compatible bias sum (capi.cu).
__global__ void update_cuda_kernel(UpdateArgs a) {
  const int T = a.Nk * a.Nl;
  const int nC = a.dM * a.dD * T;
  const float* GC = a.g;
  const float* GF = a.g + nC;
  const float* GB = a.g + 2 * nC;
  const float* GP = GB + a.dM;
  const float* SQ = GP + a.dD;
  const float lr = (1.f - a.alpha) * a.delmax;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < nC) {
    int t = n % T, d = (n / T) % a.dD, m = n / (T * a.dD);
    int nf = (d * a.dM + m) * T + t;
    if (a.mode == AEFFT_MODE_CUDA_REF_SYM) {
      float g = (GC[n] + GF[nf]) * a.inv_norm;  // :616 dDdC already holds dC+dF (kernels :464-466)
      if (a.ddc) a.ddc[n] = g;
      float v = lr * clip10(g) + a.alpha * a.dc[n];
      a.dc[n] = v;
      float w = a.c[n] - v;
      a.c[n] = w;
      a.f[nf] = w;  // :622 f[d][m][k][l] = c[m][d][k][l]
    } else {
      float g = GC[n] * a.inv_norm;
      if (a.ddc) a.ddc[n] = g;
      float v = lr * clip10(g) + a.alpha * a.dc[n];
      a.dc[n] = v;
      a.c[n] -= v;
      float gf = GF[nf] * a.inv_norm;
      if (a.ddf) a.ddf[nf] = gf;
      float vf = lr * clip10(gf) + a.alpha * a.df[nf];
      a.df[nf] = vf;
      a.f[nf] -= vf;
    }
  } else if (n < nC + a.dM) {
    int m = n - nC;
    float g = GB[m] * a.inv_norm;
    if (a.ddb) a.ddb[m] = g;
    float v = lr * clip10(g) + a.alpha * a.db[m];
    a.db[m] = v;
    a.b[m] -= v;
  } else if (n < nC + a.dM + a.dD) {
    int d = n - nC - a.dM;
    float g = GP[d] * a.inv_norm;
    if (a.ddp) a.ddp[d] = g;
    float v = lr * clip10(g) + a.alpha * a.dp[d];
    a.dp[d] = v;
    a.p[d] -= v;
  } else if (n == nC + a.dM + a.dD) {
    if (a.mse_out) *a.mse_out = SQ[0] * a.mse_scale;
  }
}

// CPU_REF (netlib.cpp:361-451), parallel formulation of the sequential-f update:
// gbuf = [R S*S | BM S | GF dD*dM*T | GP dD | SQ 1], S = dD*T, s = (d1,k1,l1), t = (d,k,l).
//   f_new[s][m] = f[s][m] - del*clip(GF/Norm)                      (all known up front: GF does not depend on f)
//   gC[m][t]    = (1/Norm) sum_s (s < t ? f_new : f_old)[s][m] * R[s][t]
//   gB[m]       = (1/Norm) sum_s f_old[s][m] * BM[s]               (b[m] is updated at the first step of each m)
// Phase 0 computes c,b,p from f_old (f untouched); phase 1 writes f_new.  Two launches keep it race-free.
__global__ void update_cpu_kernel(UpdateArgs a, int phase) {
  const int T = a.Nk * a.Nl;
  const int S = a.dD * T;
  const int nC = a.dM * a.dD * T;
  const float* R = a.g;
  const float* BM = R + (long long)S * S;
  const float* GF = BM + S;
  const float* GP = GF + nC;
  const float* SQ = GP + a.dD;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (phase == 1) {
    if (n < nC) a.f[n] -= a.delmax * clip10(GF[n] * a.inv_norm);
    return;
  }
  if (n < nC) {
    int t = n % S, m = n / S;  // c[m][d][k][l] flat = m*S + t
    double acc = 0.0;
    for (int s = 0; s < S; s++) {
      int d1 = s / T, kl = s % T;
      int nf = (d1 * a.dM + m) * T + kl;
      float fo = a.f[nf];
      float fv = (s < t) ? fo - a.delmax * clip10(GF[nf] * a.inv_norm) : fo;
      acc += (double)fv * (double)R[(long long)s * S + t];
    }
    float g = (float)(acc * (double)a.inv_norm);
    a.c[n] -= a.delmax * clip10(g);
  } else if (n < nC + a.dM) {
    int m = n - nC;
    double acc = 0.0;
    for (int s = 0; s < S; s++) {
      int d1 = s / T, kl = s % T;
      acc += (double)a.f[(d1 * a.dM + m) * T + kl] * (double)BM[s];
    }
    a.b[m] -= a.delmax * clip10((float)(acc * (double)a.inv_norm));
  } else if (n < nC + a.dM + a.dD) {
    int d = n - nC - a.dM;
    a.p[d] -= a.delmax * clip10(GP[d] * a.inv_norm);
  } else if (n == nC + a.dM + a.dD) {
    if (a.mse_out) *a.mse_out = SQ[0] * a.mse_scale;
  }
}

int launch_update(aefft_ctx* ctx, const UpdateArgs& a) {
  const int n = a.dM * a.dD * a.Nk * a.Nl + a.dM + a.dD + 1;
  const unsigned blocks = (n + 127) / 128;
  ProfScope prof(ctx, "update", 0.0, 4.0 * 6.0 * n);
  if (a.mode == AEFFT_MODE_CPU_REF) {
    update_cpu_kernel<<<blocks, 128, 0, ctx->stream>>>(a, 0);
    update_cpu_kernel<<<blocks, 128, 0, ctx->stream>>>(a, 1);
    ctx->launches += 2;
  } else {
    update_cuda_kernel<<<blocks, 128, 0, ctx->stream>>>(a);
    ctx->launches++;
  }
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

}  // namespace aefft
