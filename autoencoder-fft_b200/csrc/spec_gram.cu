// backprop_fft's iteration loop on per-bin Gram matrices (the autoencoder case expout == in).
//
// The reference's loop (fft_backproplib.cu:1443-1464) touches every frame of the pair in every one of its iterations:
// gradient_k_io (:395-475) forms G = E conj(F) and H-hat per (m, d, bin), the re-forward runs conv_k twice and calc_mse
// reads the result.  All of it is linear in the frames for fixed kernels, so the frames enter only through two small
// matrices per bin (associativity; dD = input channels of the pair, B frames):
//     Gx[d][d'] = sum_b X[b][d] conj(X[b][d'])          (iteration independent)
//     M0[d][d'] = sum_b E0[b][d] conj(X[b][d']),  E0 = O - X with the caller's `out` spectrum O (first iteration only)
// With T = F C / (dM dD) (the two convs of the pair composed, dD x dD per bin) and D = T - I the re-forward error is
// E = D X (+ beta at the DC bin, beta[d] = Nx Ny (sum_m F[d][m](0) b[m] / dD + p[d])), hence for every later iteration
//     M = sum_b E conj(X) = D Gx  (+ beta Sx^H at DC, Sx = sum_b X(0)),      sum_b |E|^2 = Re tr(M D^H)  (+ DC terms)
// and for every iteration
//     dC[m][d] = gs sum_k conj(F[k][m]) M[k][d]          (= gs sum_b G[b][m] conj(X[b][d]), :437-445)
//     dF[d][m] = gs (sum_k conj(C[m][k]) M[d][k] + [DC] b[m] Nx Ny Esum[d])    (= gs sum_b E conj(H-hat), :447-459, quirk F1)
//     db[m] = gb Re sum_d conj(F[d][m](0)) Esum[d],  dp[d] = gb Re Esum[d],  Esum = sum_b E(0) = D(0) Sx + B beta.
// ONE pass over the frames per call (the statistics kernels) instead of ~10 per iteration; an iteration costs
// O(bins dM dD^2) instead of O(bins B dM dD).  It pays when B is not small against dD (spec_gram_loop_pays).
//
// Two families, by the layout the pair's spectra live in:
//   bin-major [bin][frame][2 dD] (tensor-core levels, dD in {8, 16, 32, 64}): one CTA per bin;
//   bins-fastest [frame][dD][bins] with dD <= 4 (the image side): statistics one thread per bin, iterations LG = dM / 4
//   lanes per bin (the mapping of spec_small.cu).
#include <cstdlib>

#include "common.cuh"

namespace aefft {

namespace {

__device__ __forceinline__ void cmac(float2& acc, float2 a, float2 b) {  // acc += a*b
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cmac_conjb(float2& acc, float2 a, float2 b) {  // acc += a*conj(b)
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.y, b.x, acc.y); acc.y = fmaf(-a.x, b.y, acc.y);
}
__device__ __forceinline__ void cmac_conja(float2& acc, float2 a, float2 b) {  // acc += conj(a)*b
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(-a.y, b.x, acc.y);
}

// block-wide sum of doubles (256 threads), result valid in thread 0
__device__ __forceinline__ double block_sum_256(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += red[i];
  return s;
}

__global__ void __launch_bounds__(1024) gram_final_kernel(const double* __restrict__ part, long long n, double scale,
                                                          float* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += part[i];  // fixed assignment and order: deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *out = (float)(s * scale);
  }
}

// O given COMPACT on the support grid (sNx x sNyr) of a spectral up-sampling (resize :87-157, embedding): index of dense bin
// (i, j) of the (Nx x Nyr) grid on that grid, or -1 where the up-sampled spectrum is zero
struct Support {
  int Nx, Nyr, sNx, sNyr;  // sNx == 0: dense
};
__device__ __forceinline__ long long support_index(const Support& sp, long long w) {
  if (sp.sNx == 0) return w;
  const int i = (int)(w / sp.Nyr), j = (int)(w - (long long)i * sp.Nyr);
  int si = -1, sj = -1;
  if (i < sp.sNx / 2) si = i;
  else if (i > sp.Nx - sp.sNx / 2) si = i - sp.Nx + sp.sNx;
  else if (i == sp.Nx / 2) si = sp.sNx / 2;
  if (j < sp.sNyr - 1) sj = j;
  else if (j == sp.Nyr - 1) sj = sp.sNyr - 1;
  return (si < 0 || sj < 0) ? -1 : (long long)si * sp.sNyr + sj;
}

// ------------------------------------------------------------------------------------------------ family A: bin-major
// Statistics of one bin: E0 = O - X (sub) or O itself (the caller hands E0), Gx, M0, hw * sum |E0|^2, and at the DC bin
// Sx = sum_b X, Se = sum_b E0.  TR x TR register tiles; with fewer than 256 tiles the frames are split over NG thread
// groups whose partial sums are added in group order (deterministic).
template <int DD>
__global__ void __launch_bounds__(256) gram_stats_bm_kernel(const float* __restrict__ X, const float* __restrict__ O, int sub,
                                                            float2* __restrict__ Gx, float2* __restrict__ M0,
                                                            double* __restrict__ sq_part, float2* __restrict__ dcsum, int B,
                                                            int ncols, int col0, int Ny, Support sup) {
  constexpr int TR = DD >= 32 ? 4 : (DD >= 16 ? 2 : 1), TG = DD / TR, NT = TG * TG, NG = 256 / NT;  // (4 x 4 tiles at 16
  // channels: 16 tiles x 16 frame groups and a 64 KB reduction buffer measured 0.23 ms slower)
  static_assert(NT * NG == 256, "256 threads");
  extern __shared__ __align__(16) float2 gs_sm[];
  __shared__ double red[8];
  float2* Xs = gs_sm;                  // [B][DD]
  float2* Es = Xs + (size_t)B * DD;    // [B][DD]
  const long long w = blockIdx.x;
  const int tid = threadIdx.x;
  float sq = 0.f;
  {
    const long long wo = support_index(sup, w);  // -1: O is zero on this bin
    const float4* x4 = reinterpret_cast<const float4*>(X + w * (long long)B * 2 * DD);
    const float4* o4 = reinterpret_cast<const float4*>(O + (wo < 0 ? 0 : wo) * (long long)B * 2 * DD);
    float4* xs4 = reinterpret_cast<float4*>(Xs);
    float4* es4 = reinterpret_cast<float4*>(Es);
    for (int i = tid; i < B * DD / 2; i += 256) {
      const float4 x = __ldg(x4 + i);
      float4 e = wo < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(o4 + i);
      if (sub) { e.x -= x.x; e.y -= x.y; e.z -= x.z; e.w -= x.w; }
      xs4[i] = x; es4[i] = e;
      sq = fmaf(e.x, e.x, fmaf(e.y, e.y, fmaf(e.z, e.z, fmaf(e.w, e.w, sq))));
    }
  }
  {
    const int wy = col0 + (int)(w % ncols);
    const double hw = (wy == 0 || wy == Ny / 2) ? 1.0 : 2.0;
    const double s = block_sum_256((double)sq, red);  // (includes the barrier that publishes Xs / Es)
    if (tid == 0 && sq_part) sq_part[w] = s * hw;
  }
  if (dcsum && w == 0 && tid < 2 * DD) {  // DC bin: dcsum[0..DD) = Sx, [DD..2DD) = Se
    const float2* src = tid < DD ? Xs : Es;
    const int d = tid < DD ? tid : tid - DD;
    double sr = 0.0, si = 0.0;
    for (int b = 0; b < B; b++) { sr += (double)src[b * DD + d].x; si += (double)src[b * DD + d].y; }
    dcsum[tid] = make_float2((float)sr, (float)si);
  }
  const int tile = tid % NT, grp = tid / NT;
  const int dr = (tile / TG) * TR, dc = (tile % TG) * TR;
  float2 ag[TR][TR], am[TR][TR];
#pragma unroll
  for (int a = 0; a < TR; a++)
#pragma unroll
    for (int c = 0; c < TR; c++) { ag[a][c] = make_float2(0.f, 0.f); am[a][c] = make_float2(0.f, 0.f); }
#pragma unroll 2
  for (int b = grp; b < B; b += NG) {
    float2 xr[TR], er[TR], xc[TR];
#pragma unroll
    for (int a = 0; a < TR; a++) { xr[a] = Xs[b * DD + dr + a]; er[a] = Es[b * DD + dr + a]; xc[a] = Xs[b * DD + dc + a]; }
#pragma unroll
    for (int a = 0; a < TR; a++)
#pragma unroll
      for (int c = 0; c < TR; c++) { cmac_conjb(ag[a][c], xr[a], xc[c]); cmac_conjb(am[a][c], er[a], xc[c]); }
  }
  float2* go = Gx + w * DD * DD;
  float2* mo = M0 + w * DD * DD;
  if constexpr (NG == 1) {
#pragma unroll
    for (int a = 0; a < TR; a++)
#pragma unroll
      for (int c = 0; c < TR; c++) { go[(dr + a) * DD + dc + c] = ag[a][c]; mo[(dr + a) * DD + dc + c] = am[a][c]; }
  } else {
    __syncthreads();  // every group is done with X and E: their space takes the partial sums [group][2][DD][DD]
    float2* Pp = gs_sm + (size_t)grp * 2 * DD * DD;
#pragma unroll
    for (int a = 0; a < TR; a++)
#pragma unroll
      for (int c = 0; c < TR; c++) {
        Pp[(dr + a) * DD + dc + c] = ag[a][c];
        Pp[DD * DD + (dr + a) * DD + dc + c] = am[a][c];
      }
    __syncthreads();
    for (int i = tid; i < 2 * DD * DD; i += 256) {
      float2 sum = gs_sm[i];
#pragma unroll
      for (int g = 1; g < NG; g++) { const float2 v = gs_sm[(size_t)g * 2 * DD * DD + i]; sum.x += v.x; sum.y += v.y; }
      if (i < DD * DD) go[i] = sum; else mo[i - DD * DD] = sum;
    }
  }
}

struct GramIterBm {
  const float2 *Gx, *M0;      // [S][DD][DD]
  const float *Cemb, *Femb;   // embedded kernel spectra [S][2 dM][2 DD], [S][2 DD][2 dM]: row 2r = (Re, -Im) of W[r][:]
  float2 *dCt, *dFt;          // [S][dM][DD], nullptr: only the mse is wanted
  double* sq_part;            // [S], written unless `first`
  const float2* dcsum;        // Sx | Se of the DC bin (bin 0 of a device that owns it), nullptr otherwise
  const float *bias_b, *bias_p;
  float *db, *dp;
  int B, dM, first, ncols, col0, Ny;
  float gs, gb, tscale, norm;  // gradient scale, bias-gradient scale, 1 / (dM dD), Nx Ny
};

// One bin of an iteration: M (from M0 on the first iteration, else D Gx with D = F C / (dM dD) - I, which also gives the
// mse of the current kernels), then both gradient spectra.  Shared memory: Ms, Gs, Ds [DD][DD + 1]; Fs[k][m] = F[k][m],
// Cs[k][m] = C[m][k] (phase 2: NO adjacent m of one d per thread, 16-byte loads), Cm[m][k] = C[m][k] (for T).
template <int DD, int NO>
__global__ void __launch_bounds__(256) gram_iter_bm_kernel(GramIterBm p) {
  constexpr int MP = DD + 1;
  extern __shared__ __align__(16) float2 gi_sm[];
  __shared__ double red[8];
  __shared__ float2 dcv[3 * DD];  // DC bin: Sx, Esum, (beta, 0)
  const int dM = p.dM;
  const int CP = dM + 2;              // pitch of Cs: the transposing stores below hit 2 banks per lane instead of one for all
  float2* Fs = gi_sm;                 // [DD][dM]
  float2* Cs = Fs + DD * dM;          // [DD][CP]
  float2* Cm = Cs + DD * CP;          // [dM][DD]
  float2* Ms = Cm + DD * dM;          // [DD][MP]
  float2* Gs = Ms + DD * MP;
  float2* Ds = Gs + DD * MP;
  const long long w = blockIdx.x;
  const int tid = threadIdx.x;
  const bool dc_bin = p.dcsum && w == 0;
  for (int i = tid; i < dM * DD; i += 256) {
    const int k = i / dM, m = i - k * dM;   // F[k][m]: embedded row 2k, columns 2m, 2m+1
    const float2 u = __ldg(reinterpret_cast<const float2*>(p.Femb + ((w * 2 * DD + 2 * k) * 2 * (long long)dM + 2 * m)));
    Fs[i] = make_float2(u.x, -u.y);
    const int mc = i / DD, kc = i - mc * DD;  // C[mc][kc]: embedded row 2 mc, columns 2 kc, 2 kc + 1
    const float2 v = __ldg(reinterpret_cast<const float2*>(p.Cemb + ((w * 2 * dM + 2 * mc) * 2 * (long long)DD + 2 * kc)));
    Cs[kc * CP + mc] = make_float2(v.x, -v.y);
    Cm[i] = make_float2(v.x, -v.y);
  }
  const float2* src = p.first ? p.M0 : p.Gx;
  for (int i = tid; i < DD * DD; i += 256) {
    const float2 v = __ldg(src + w * DD * DD + i);
    (p.first ? Ms : Gs)[(i / DD) * MP + (i % DD)] = v;
  }
  if (dc_bin && tid < DD) {
    dcv[tid] = p.dcsum[tid];                        // Sx
    if (p.first) dcv[DD + tid] = p.dcsum[DD + tid];  // Esum of the first iteration = Se
  }
  __syncthreads();
  if (!p.first) {
    // D = F C / (dM dD) - I, 2 x 2 register tiles (3 shared-memory loads per 4 complex MACs instead of 8: the kernel is bound
    // by shared-memory wavefronts)
    constexpr int HT = DD / 2;
    for (int t = tid; t < HT * HT; t += 256) {
      const int d = (t / HT) * 2, e = (t % HT) * 2;
      float2 a00 = make_float2(0.f, 0.f), a01 = a00, a10 = a00, a11 = a00;
#pragma unroll 4
      for (int m = 0; m < dM; m++) {
        const float2 f0 = Fs[d * dM + m], f1 = Fs[(d + 1) * dM + m];
        const float4 cv = *reinterpret_cast<const float4*>(Cm + m * DD + e);
        const float2 c0 = make_float2(cv.x, cv.y), c1 = make_float2(cv.z, cv.w);
        cmac(a00, f0, c0); cmac(a01, f0, c1); cmac(a10, f1, c0); cmac(a11, f1, c1);
      }
      Ds[d * MP + e] = make_float2(a00.x * p.tscale - (d == e ? 1.f : 0.f), a00.y * p.tscale);
      Ds[d * MP + e + 1] = make_float2(a01.x * p.tscale, a01.y * p.tscale);
      Ds[(d + 1) * MP + e] = make_float2(a10.x * p.tscale, a10.y * p.tscale);
      Ds[(d + 1) * MP + e + 1] = make_float2(a11.x * p.tscale - (d == e ? 1.f : 0.f), a11.y * p.tscale);
    }
    if (dc_bin && tid < DD) {  // beta[d] = Nx Ny (sum_m Re F[d][m](0) b[m] / dD + p[d])
      float s = 0.f;
      for (int m = 0; m < dM; m++) s = fmaf(Fs[tid * dM + m].x, p.bias_b[m], s);
      dcv[2 * DD + tid] = make_float2(p.norm * (s / (float)DD + p.bias_p[tid]), 0.f);
    }
    __syncthreads();
    // M = D Gx (+ beta Sx^H at DC);  sum_b |E|^2 = Re sum M o conj(D)  (+ 2 beta . Re(D Sx) + B |beta|^2 at DC)
    float sq = 0.f;
    for (int t = tid; t < HT * HT; t += 256) {
      const int d = (t / HT) * 2, l = (t % HT) * 2;
      float2 y[2][2];
      y[0][0] = y[0][1] = y[1][0] = y[1][1] = make_float2(0.f, 0.f);
#pragma unroll 4
      for (int k = 0; k < DD; k++) {
        const float2 d0 = Ds[d * MP + k], d1 = Ds[(d + 1) * MP + k], g0 = Gs[k * MP + l], g1 = Gs[k * MP + l + 1];
        cmac(y[0][0], d0, g0); cmac(y[0][1], d0, g1); cmac(y[1][0], d1, g0); cmac(y[1][1], d1, g1);
      }
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const float2 dd = Ds[(d + a) * MP + l + c];
          sq = fmaf(y[a][c].x, dd.x, fmaf(y[a][c].y, dd.y, sq));
          if (dc_bin) {
            const float beta = dcv[2 * DD + d + a].x;
            y[a][c].x = fmaf(beta, dcv[l + c].x, y[a][c].x);
            y[a][c].y = fmaf(-beta, dcv[l + c].y, y[a][c].y);
          }
          Ms[(d + a) * MP + l + c] = y[a][c];
        }
    }
    if (dc_bin && tid < DD) {
      float2 ds = make_float2(0.f, 0.f);
      for (int k = 0; k < DD; k++) cmac(ds, Ds[tid * MP + k], dcv[k]);
      const float beta = dcv[2 * DD + tid].x;
      dcv[DD + tid] = make_float2(fmaf((float)p.B, beta, ds.x), ds.y);  // Esum = D Sx + B beta
      sq += 2.f * beta * ds.x + (float)p.B * beta * beta;
    }
    const int wy = p.col0 + (int)(w % p.ncols);
    const double hw = (wy == 0 || wy == p.Ny / 2) ? 1.0 : 2.0;
    const double s = block_sum_256((double)sq, red);  // (includes the barrier that publishes Ms and Esum)
    if (tid == 0) p.sq_part[w] = s * hw;
  }
  if (!p.dCt) return;
  if (dc_bin) {  // bias gradients from Esum
    if (tid < dM) {
      float s = 0.f;
      for (int d = 0; d < DD; d++) s = fmaf(Fs[d * dM + tid].x, dcv[DD + d].x, fmaf(Fs[d * dM + tid].y, dcv[DD + d].y, s));
      p.db[tid] = s * p.gb;
    } else if (tid < dM + DD) {
      p.dp[tid - dM] = dcv[DD + tid - dM].x * p.gb;
    }
  }
  const int d = tid % DD, m0 = (tid / DD) * NO;
  if (m0 < dM) {
    float2 aC[NO], aF[NO];
#pragma unroll
    for (int j = 0; j < NO; j++) { aC[j] = make_float2(0.f, 0.f); aF[j] = make_float2(0.f, 0.f); }
#pragma unroll 2
    for (int k = 0; k < DD; k++) {
      const float2 m1 = Ms[k * MP + d], m2 = Ms[d * MP + k];
      float2 f[NO], c[NO];
      if constexpr (NO % 2 == 0) {
#pragma unroll
        for (int j = 0; j < NO; j += 2) {
          const float4 fv = *reinterpret_cast<const float4*>(Fs + k * dM + m0 + j);
          const float4 cv = *reinterpret_cast<const float4*>(Cs + k * CP + m0 + j);
          f[j] = make_float2(fv.x, fv.y); f[j + 1] = make_float2(fv.z, fv.w);
          c[j] = make_float2(cv.x, cv.y); c[j + 1] = make_float2(cv.z, cv.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NO; j++) { f[j] = Fs[k * dM + m0 + j]; c[j] = Cs[k * CP + m0 + j]; }
      }
#pragma unroll
      for (int j = 0; j < NO; j++) { cmac_conja(aC[j], f[j], m1); cmac_conja(aF[j], c[j], m2); }
    }
#pragma unroll
    for (int j = 0; j < NO; j++) {
      if (dc_bin) {  // the bias part of H-hat (quirk F1): + b[m] Nx Ny Esum[d]
        const float bb = p.bias_b[m0 + j] * p.norm;
        aF[j].x = fmaf(bb, dcv[DD + d].x, aF[j].x);
        aF[j].y = fmaf(bb, dcv[DD + d].y, aF[j].y);
      }
      const long long o = (w * dM + m0 + j) * DD + d;
      p.dCt[o] = make_float2(aC[j].x * p.gs, aC[j].y * p.gs);
      p.dFt[o] = make_float2(aF[j].x * p.gs, aF[j].y * p.gs);
    }
  }
}

// ------------------------------------------------------------------------------------------------ family B: bins-fastest, dD <= 4
// Statistics: one thread per bin walks over the frames.  Gx, M0 are written [DD * DD][S] (bins fastest).
template <int DD>
__global__ void __launch_bounds__(128) gram_stats_ff_kernel(const float2* __restrict__ X, const float2* __restrict__ O,
                                                            float2* __restrict__ Gx, float2* __restrict__ M0,
                                                            double* __restrict__ sq_part, float2* __restrict__ dcsum, long long S,
                                                            int B, int ncols, int col0, int Ny, Support sup) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = w < S;
  const long long wc = live ? w : 0;
  const long long wo = support_index(sup, wc);  // -1: O is zero on this bin
  const long long So = sup.sNx ? (long long)sup.sNx * sup.sNyr : S, fso = (long long)DD * So;
  float2 g[DD][DD], m[DD][DD];
  float2 sx[DD], se[DD];
#pragma unroll
  for (int a = 0; a < DD; a++) {
    sx[a] = make_float2(0.f, 0.f); se[a] = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < DD; c++) { g[a][c] = make_float2(0.f, 0.f); m[a][c] = make_float2(0.f, 0.f); }
  }
  float sq = 0.f;
  double tot = 0.0;
  const long long fs = (long long)DD * S;
  constexpr int PF = 4;  // frames ahead pulled into L2
#pragma unroll 2
  for (int b = 0; b < B; b++) {
    if (b + PF < B) {
#pragma unroll
      for (int d = 0; d < DD; d++) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(X + (b + PF) * fs + d * S + wc));
        if (wo >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(O + (b + PF) * fso + d * So + wo));
      }
    }
    float2 x[DD], e[DD];
#pragma unroll
    for (int d = 0; d < DD; d++) {
      x[d] = __ldg(X + b * fs + d * S + wc);
      const float2 o = wo >= 0 ? __ldg(O + b * fso + d * So + wo) : make_float2(0.f, 0.f);
      e[d] = make_float2(o.x - x[d].x, o.y - x[d].y);
      sq = fmaf(e[d].x, e[d].x, fmaf(e[d].y, e[d].y, sq));
    }
    if (wc == 0) {
#pragma unroll
      for (int d = 0; d < DD; d++) { sx[d].x += x[d].x; sx[d].y += x[d].y; se[d].x += e[d].x; se[d].y += e[d].y; }
    }
#pragma unroll
    for (int a = 0; a < DD; a++)
#pragma unroll
      for (int c = 0; c < DD; c++) { cmac_conjb(g[a][c], x[a], x[c]); cmac_conjb(m[a][c], e[a], x[c]); }
    if ((b & 7) == 7) { tot += (double)sq; sq = 0.f; }
  }
  tot += (double)sq;
  if (live) {
#pragma unroll
    for (int a = 0; a < DD; a++)
#pragma unroll
      for (int c = 0; c < DD; c++) {
        Gx[(long long)(a * DD + c) * S + w] = g[a][c];
        M0[(long long)(a * DD + c) * S + w] = m[a][c];
      }
    if (w == 0 && dcsum) {
#pragma unroll
      for (int d = 0; d < DD; d++) { dcsum[d] = sx[d]; dcsum[DD + d] = se[d]; }
    }
  }
  const int wy = col0 + (int)(wc % ncols);
  const double hw = (wy == 0 || wy == Ny / 2) ? 1.0 : 2.0;
  tot = live ? tot * hw : 0.0;
  __shared__ double red[8];
  const double s = block_sum_256(tot, red);
  if (threadIdx.x == 0 && sq_part) sq_part[blockIdx.x] = s;
}

struct GramIterFf {
  const float2 *Gx, *M0;      // [DD * DD][S]
  const float2 *C, *F;        // [dM][DD][S], [DD][dM][S]
  float2 *dC, *dF;            // [dM][DD][S], [DD][dM][S]; nullptr: only the mse is wanted
  double* sq_part;            // per block, written unless `first`
  const float2* dcsum;        // Sx | Se of the DC bin, nullptr on devices that do not own it
  const float *bias_b, *bias_p;
  float *db, *dp;
  long long S;
  int B, dM, first, ncols, col0, Ny;
  float gs, gb, tscale, norm;
};

// LG = dM / 4 lanes per bin, each owns 4 hidden channels (its rows of C, columns of F); T is summed over the lane group.
template <int DD, int LG>
__global__ void __launch_bounds__(128) gram_iter_ff_kernel(GramIterFf p) {
  const int tid = threadIdx.x;
  const int mq = tid % LG;
  const long long w = (long long)blockIdx.x * (128 / LG) + tid / LG;
  const bool live = w < p.S;
  const long long wc = live ? w : 0;
  const int dM = p.dM;
  const bool dc_bin = p.dcsum && wc == 0;
  float2 Cq[4][DD], Fq[DD][4];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int m = 4 * mq + a;
#pragma unroll
    for (int d = 0; d < DD; d++) {
      Cq[a][d] = __ldg(p.C + ((long long)m * DD + d) * p.S + wc);
      Fq[d][a] = __ldg(p.F + ((long long)d * dM + m) * p.S + wc);
    }
  }
  float2 M[DD][DD];
  float2 esum[DD];
#pragma unroll
  for (int d = 0; d < DD; d++) esum[d] = make_float2(0.f, 0.f);
  double tot = 0.0;
  if (p.first) {
#pragma unroll
    for (int a = 0; a < DD; a++)
#pragma unroll
      for (int c = 0; c < DD; c++) M[a][c] = __ldg(p.M0 + (long long)(a * DD + c) * p.S + wc);
    if (dc_bin) {
#pragma unroll
      for (int d = 0; d < DD; d++) esum[d] = p.dcsum[DD + d];
    }
  } else {
    float2 T[DD][DD];
#pragma unroll
    for (int a = 0; a < DD; a++)
#pragma unroll
      for (int c = 0; c < DD; c++) {
        float2 t = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 4; q++) cmac(t, Fq[a][q], Cq[q][c]);
#pragma unroll
        for (int o = 1; o < LG; o <<= 1) {
          t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
          t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
        }
        T[a][c] = make_float2(t.x * p.tscale - (a == c ? 1.f : 0.f), t.y * p.tscale);  // D = T - I
      }
    float2 G[DD][DD];
#pragma unroll
    for (int a = 0; a < DD; a++)
#pragma unroll
      for (int c = 0; c < DD; c++) G[a][c] = __ldg(p.Gx + (long long)(a * DD + c) * p.S + wc);
    float sq = 0.f;
#pragma unroll
    for (int a = 0; a < DD; a++)
#pragma unroll
      for (int c = 0; c < DD; c++) {
        float2 y = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < DD; k++) cmac(y, T[a][k], G[k][c]);
        sq = fmaf(y.x, T[a][c].x, fmaf(y.y, T[a][c].y, sq));
        M[a][c] = y;
      }
    // sum_m Re F[d][m] b[m] over the lane group (every lane takes part in the shuffles; only the DC bin uses the result)
    float fb[DD];
#pragma unroll
    for (int d = 0; d < DD; d++) {
      float s = 0.f;
      if (p.dcsum) {
#pragma unroll
        for (int q = 0; q < 4; q++) s = fmaf(Fq[d][q].x, p.bias_b[4 * mq + q], s);
      }
#pragma unroll
      for (int o = 1; o < LG; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      fb[d] = s;
    }
    if (dc_bin) {
      float2 sx[DD];
      float beta[DD];
#pragma unroll
      for (int d = 0; d < DD; d++) {
        sx[d] = p.dcsum[d];
        beta[d] = p.norm * (fb[d] / (float)DD + p.bias_p[d]);
      }
#pragma unroll
      for (int a = 0; a < DD; a++) {
        float2 ds = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < DD; k++) cmac(ds, T[a][k], sx[k]);
        esum[a] = make_float2(fmaf((float)p.B, beta[a], ds.x), ds.y);
        sq += 2.f * beta[a] * ds.x + (float)p.B * beta[a] * beta[a];
#pragma unroll
        for (int c = 0; c < DD; c++) { M[a][c].x = fmaf(beta[a], sx[c].x, M[a][c].x); M[a][c].y = fmaf(-beta[a], sx[c].y, M[a][c].y); }
      }
    }
    const int wy = p.col0 + (int)(wc % p.ncols);
    const double hw = (wy == 0 || wy == p.Ny / 2) ? 1.0 : 2.0;
    tot = (live && mq == 0) ? (double)sq * hw : 0.0;
  }
  if (!p.first) {
    __shared__ double red[8];
    const double s = block_sum_256(tot, red);
    if (tid == 0) p.sq_part[blockIdx.x] = s;
  }
  if (!p.dC || !live) return;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int m = 4 * mq + a;
    const float bb = dc_bin ? p.bias_b[m] * p.norm : 0.f;
#pragma unroll
    for (int d = 0; d < DD; d++) {
      float2 aC = make_float2(0.f, 0.f), aF = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < DD; k++) { cmac_conja(aC, Fq[k][a], M[k][d]); cmac_conja(aF, Cq[a][k], M[d][k]); }
      aF.x = fmaf(bb, esum[d].x, aF.x);
      aF.y = fmaf(bb, esum[d].y, aF.y);
      p.dC[((long long)m * DD + d) * p.S + w] = make_float2(aC.x * p.gs, aC.y * p.gs);
      p.dF[((long long)d * dM + m) * p.S + w] = make_float2(aF.x * p.gs, aF.y * p.gs);
    }
  }
  if (dc_bin) {
#pragma unroll
    for (int a = 0; a < 4; a++) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DD; d++) s = fmaf(Fq[d][a].x, esum[d].x, fmaf(Fq[d][a].y, esum[d].y, s));
      p.db[4 * mq + a] = s * p.gb;
    }
    if (mq == 0) {
#pragma unroll
      for (int d = 0; d < DD; d++) p.dp[d] = esum[d].x * p.gb;
    }
  }
}

size_t gram_iter_bm_smem(int dD, int dM) { return ((size_t)3 * dD * dM + 2 * dD + (size_t)3 * dD * (dD + 1)) * sizeof(float2); }
size_t gram_stats_bm_smem(int B, int dD) {
  const int tr = dD >= 32 ? 4 : (dD >= 16 ? 2 : 1), ng = 256 / ((dD / tr) * (dD / tr));
  const size_t ex = 2 * (size_t)B * dD, pp = ng > 1 ? (size_t)ng * 2 * dD * dD : 0;
  return (ex > pp ? ex : pp) * sizeof(float2);
}
bool bm_shape_ok(int dD, int dM) {
  if (dD != 8 && dD != 16 && dD != 32 && dD != 64) return false;
  if (dM * dD >= 256) {
    const int no = dM * dD / 256;
    if (dM * dD % 256 != 0 || (no != 1 && no != 2 && no != 4 && no != 8 && no != 16)) return false;
  } else if (256 % dD != 0) {
    return false;
  }
  return true;
}

}  // namespace

// Does the Gram loop pay?  Per bin and iteration it costs ~8 dD^2 (dM + dD) FMA-equivalents on the CUDA cores, the per-frame
// forms stream B (dD + dM)-sized operands several times (tensor-core path: ~10 B (dD + dM) + 16 dD dM floats).
bool spec_gram_loop_pays(int B, int dD, int dM, bool bin_major) {
  if (getenv("AEFFT_NO_GRAM_LOOP")) return false;
  if (bin_major) {
    if (!bm_shape_ok(dD, dM)) return false;
    if (gram_iter_bm_smem(dD, dM) > 200 * 1024 || gram_stats_bm_smem(B, dD) > 200 * 1024) return false;
  } else {
    const int lg = dM / 4;
    if (dD < 1 || dD > 4 || dM % 4 != 0 || !(lg == 1 || lg == 2 || lg == 4 || lg == 8 || lg == 16)) return false;
  }
  if (getenv("AEFFT_FORCE_GRAM_LOOP")) return true;
  if (!bin_major) return true;  // dD <= 4: always far cheaper than walking over the frames
  const double clk_new = 8.0 * dD * dD * ((double)dM + dD) / 128.0;
  const double clk_old = (10.0 * B * ((double)dD + dM) + 16.0 * dD * dM) * 4.0 / 15.6;
  return clk_new < clk_old;
}

int launch_gram_final(aefft_ctx* ctx, const double* part, long long n, double scale, float* out) {
  gram_final_kernel<<<1, 1024, 0, ctx->stream>>>(part, n, scale, out);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// statistics of a bin-major pair: X, O [S][B][2 dD] (sub: E0 = O - X, else O is E0 already) -> Gx, M0 [S][dD][dD][2],
// *mse_out = mse_scale * sum_bins hw |E0|^2 (when mse_out), dcsum = Sx | Se of bin 0 (when the device owns the DC column)
int launch_gram_stats_bm(aefft_ctx* ctx, long long S, int B, int dD, const float* X, const float* O, int sub, float* Gx, float* M0,
                         float* mse_out, double mse_scale, float* dcsum, int ncols, int col0, int Ny, int Nx, int sNx, int sNy) {
  if (ncols <= 0) { ncols = Ny / 2 + 1; col0 = 0; }
  AE_ARG(sNx == 0 || (ncols == Ny / 2 + 1 && Nx > 0 && S == (long long)Nx * ncols && sNy > 0));
  const Support sup{Nx, Ny / 2 + 1, sNx, sNx ? sNy / 2 + 1 : 0};
  const size_t smem = gram_stats_bm_smem(B, dD);
  double* part;
  AE_TRY(ctx->getT("gram_sq_part", (size_t)S, &part));
  {
    ProfScope prof(ctx, "spec_gram_stats", 16.0 * S * B * dD * dD, 16.0 * S * B * dD);
#define AEFFT_GS(dd)                                                                                                       \
  if (dD == dd) {                                                                                                          \
    AE_TRY(ctx->ensure_dyn_smem((const void*)gram_stats_bm_kernel<dd>, smem));                                             \
    gram_stats_bm_kernel<dd><<<(unsigned)S, 256, smem, ctx->stream>>>(X, O, sub, (float2*)Gx, (float2*)M0, part,           \
                                                                     (float2*)dcsum, B, ncols, col0, Ny, sup);             \
  }
    AEFFT_GS(8) AEFFT_GS(16) AEFFT_GS(32) AEFFT_GS(64)
#undef AEFFT_GS
    ctx->launches++;
    AE_CUDA(cudaGetLastError());
  }
  if (mse_out) AE_TRY(launch_gram_final(ctx, part, S, mse_scale, mse_out));
  return AEFFT_OK;
}

// one iteration on the Gram matrices of a bin-major pair.  first: M = M0 (no mse); else M = D Gx and *mse_out gets the mse of
// the current kernels.  dCt == nullptr: mse only.  dcsum / bias_* / db / dp: the device that owns the DC column.
int launch_gram_iter_bm(aefft_ctx* ctx, long long S, int B, int dD, int dM, const float* Gx, const float* M0, const float* Cemb,
                        const float* Femb, int first, float gs, float gb, float norm, const float* dcsum, const float* bias_b,
                        const float* bias_p, float* dCt, float* dFt, float* db, float* dp, float* mse_out, double mse_scale, int ncols,
                        int col0, int Ny) {
  if (ncols <= 0) { ncols = Ny / 2 + 1; col0 = 0; }
  AE_ARG(bm_shape_ok(dD, dM) && (first || mse_out) && (dCt || !first));
  double* part;
  AE_TRY(ctx->getT("gram_sq_part", (size_t)S, &part));
  GramIterBm p{(const float2*)Gx, (const float2*)M0, Cemb, Femb, (float2*)dCt, (float2*)dFt, part, (const float2*)dcsum, bias_b, bias_p,
               db, dp, B, dM, first, ncols, col0, Ny, gs, gb, 1.f / ((float)dM * (float)dD), norm};
  const size_t smem = gram_iter_bm_smem(dD, dM);
  const int no = dM * dD >= 256 ? dM * dD / 256 : 1;
  bool done = false;
  {
    ProfScope prof(ctx, "spec_gram_iter", 8.0 * S * dD * dD * ((first ? 0.0 : (double)dM + dD) + (dCt ? 2.0 * dM : 0.0)),
                   8.0 * S * ((double)dD * dD + 2.0 * dM * dD + (dCt ? 2.0 * dM * dD : 0.0)));
#define AEFFT_GI(dd, n)                                                                           \
  if (!done && dD == dd && no == n) {                                                             \
    AE_TRY(ctx->ensure_dyn_smem((const void*)gram_iter_bm_kernel<dd, n>, smem));                  \
    gram_iter_bm_kernel<dd, n><<<(unsigned)S, 256, smem, ctx->stream>>>(p);                       \
    done = true;                                                                                  \
  }
    AEFFT_GI(8, 1) AEFFT_GI(8, 2) AEFFT_GI(8, 4) AEFFT_GI(8, 8)
    AEFFT_GI(16, 1) AEFFT_GI(16, 2) AEFFT_GI(16, 4) AEFFT_GI(16, 8) AEFFT_GI(16, 16)
    AEFFT_GI(32, 1) AEFFT_GI(32, 2) AEFFT_GI(32, 4) AEFFT_GI(32, 8) AEFFT_GI(32, 16)
    AEFFT_GI(64, 1) AEFFT_GI(64, 2) AEFFT_GI(64, 4) AEFFT_GI(64, 8) AEFFT_GI(64, 16)
#undef AEFFT_GI
    if (!done) return AEFFT_ERR_UNSUPPORTED;
    ctx->launches++;
    AE_CUDA(cudaGetLastError());
  }
  if (!first) AE_TRY(launch_gram_final(ctx, part, S, mse_scale, mse_out));
  return AEFFT_OK;
}

// ---- bins-fastest, dD <= 4
int launch_gram_stats_ff(aefft_ctx* ctx, long long S, int B, int dD, const float2* X, const float2* O, float2* Gx, float2* M0,
                         float* mse_out, double mse_scale, float* dcsum, int ncols, int col0, int Ny, int Nx, int sNx, int sNy) {
  if (ncols <= 0) { ncols = Ny / 2 + 1; col0 = 0; }
  AE_ARG(sNx == 0 || (ncols == Ny / 2 + 1 && Nx > 0 && S == (long long)Nx * ncols && sNy > 0));
  const Support sup{Nx, Ny / 2 + 1, sNx, sNx ? sNy / 2 + 1 : 0};
  const long long blocks = (S + 127) / 128;
  double* part;
  AE_TRY(ctx->getT("gram_sq_part", (size_t)(S > blocks ? S : blocks), &part));
  {
    ProfScope prof(ctx, "spec_gram_stats", 16.0 * S * B * dD * dD, 16.0 * S * B * dD);
    switch (dD) {
      case 1: gram_stats_ff_kernel<1><<<(unsigned)blocks, 128, 0, ctx->stream>>>(X, O, Gx, M0, part, (float2*)dcsum, S, B, ncols, col0, Ny, sup); break;
      case 2: gram_stats_ff_kernel<2><<<(unsigned)blocks, 128, 0, ctx->stream>>>(X, O, Gx, M0, part, (float2*)dcsum, S, B, ncols, col0, Ny, sup); break;
      case 3: gram_stats_ff_kernel<3><<<(unsigned)blocks, 128, 0, ctx->stream>>>(X, O, Gx, M0, part, (float2*)dcsum, S, B, ncols, col0, Ny, sup); break;
      case 4: gram_stats_ff_kernel<4><<<(unsigned)blocks, 128, 0, ctx->stream>>>(X, O, Gx, M0, part, (float2*)dcsum, S, B, ncols, col0, Ny, sup); break;
      default: return AEFFT_ERR_UNSUPPORTED;
    }
    ctx->launches++;
    AE_CUDA(cudaGetLastError());
  }
  if (mse_out) AE_TRY(launch_gram_final(ctx, part, blocks, mse_scale, mse_out));
  return AEFFT_OK;
}

namespace {
template <int DD>
int gram_iter_ff_lg(aefft_ctx* ctx, const GramIterFf& p, int lg, long long blocks) {
  switch (lg) {
    case 1: gram_iter_ff_kernel<DD, 1><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p); break;
    case 2: gram_iter_ff_kernel<DD, 2><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p); break;
    case 4: gram_iter_ff_kernel<DD, 4><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p); break;
    case 8: gram_iter_ff_kernel<DD, 8><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p); break;
    case 16: gram_iter_ff_kernel<DD, 16><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p); break;
    default: return AEFFT_ERR_UNSUPPORTED;
  }
  return AEFFT_OK;
}
}  // namespace

int launch_gram_iter_ff(aefft_ctx* ctx, long long S, int B, int dD, int dM, const float2* Gx, const float2* M0, const float2* C,
                        const float2* F, int first, float gs, float gb, float norm, const float* dcsum, const float* bias_b,
                        const float* bias_p, float2* dC, float2* dF, float* db, float* dp, float* mse_out, double mse_scale, int ncols,
                        int col0, int Ny) {
  if (ncols <= 0) { ncols = Ny / 2 + 1; col0 = 0; }
  AE_ARG((first || mse_out) && (dC || !first));
  const int lg = dM / 4;
  const long long blocks = (S + 128 / lg - 1) / (128 / lg);
  double* part;
  AE_TRY(ctx->getT("gram_sq_part", (size_t)(S > blocks ? S : blocks), &part));
  GramIterFf p{Gx, M0, C, F, dC, dF, part, (const float2*)dcsum, bias_b, bias_p, db, dp, S, B, dM, first, ncols, col0, Ny, gs, gb,
               1.f / ((float)dM * (float)dD), norm};
  {
    ProfScope prof(ctx, "spec_gram_iter", 8.0 * S * dD * dD * ((first ? 0.0 : (double)dM + dD) + (dC ? 2.0 * dM : 0.0)),
                   8.0 * S * ((double)dD * dD + 2.0 * dM * dD + (dC ? 2.0 * dM * dD : 0.0)));
    int rc = AEFFT_ERR_UNSUPPORTED;
    switch (dD) {
      case 1: rc = gram_iter_ff_lg<1>(ctx, p, lg, blocks); break;
      case 2: rc = gram_iter_ff_lg<2>(ctx, p, lg, blocks); break;
      case 3: rc = gram_iter_ff_lg<3>(ctx, p, lg, blocks); break;
      case 4: rc = gram_iter_ff_lg<4>(ctx, p, lg, blocks); break;
    }
    if (rc != AEFFT_OK) return rc;
    ctx->launches++;
    AE_CUDA(cudaGetLastError());
  }
  if (!first) AE_TRY(launch_gram_final(ctx, part, blocks, mse_scale, mse_out));
  return AEFFT_OK;
}

}  // namespace aefft
