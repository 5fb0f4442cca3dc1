// Weight-gradient ("correlation") contraction, fp32 on CUDA cores:
//     G[a][x][k][l] = sum_{b,i,j} A[b][a](i,j) * X[b][x](i + ai0 + tk, j + aj0 + tl)
// summed over every frame and pixel, plus sumA[a] and sum A^2 as by-products (bias gradients and the printed mse).
//
// Replaces the reference's one-launch-per-weight-element loop (backproplib.cu:363-417 / :594-643: dM*dD*Nk*Nl
// launches, each re-deriving the hidden delta and followed by 2-4 thrust::reduce calls) with ONE persistent
// split-K kernel per operand pair: a CTA owns an (AB x XB) channel block and walks a contiguous range of
// 8x64-pixel tiles; thread = (8-channel A group, X channel, tap row tk, tile row g) keeps 8 x Nl accumulators in
// registers across ALL its tiles (1 LDS.128 per 16 FFMA), the 8 row-groups are tree-reduced through shared memory
// once per CTA, and the per-CTA partials are summed in a fixed order by reduce_partials (deterministic).
#include "common.cuh"

namespace aefft {

constexpr int WTI = 8, WTJ = 64;  // pixel tile; WTI == number of row groups

struct WgradParams {
  const float* a0;
  const float* a1;
  const float* X;
  float* part;      // [chunks][nA][nX][Nk][Nl]
  float* part_sum;  // [chunks][nA]
  float* part_sq;   // [chunks]
  int amode, nA, src_ch, nX, Nx, Ny;
  int ai0, aj0, flip, lo;
  int ei0, ej0, eNk, eNl, out_lo;
  int tiles_i, tiles_j;
  long long n_tiles;  // B * tiles_i * tiles_j
  int n_chunks, nXB;
};

__device__ __forceinline__ float load_A(const WgradParams& p, long long b, int a, int i, int j) {
  const long long plane = (long long)p.Nx * p.Ny;
  if (p.amode == A_SHIFT) {
    int tl1 = a % p.eNl, tk1 = (a / p.eNl) % p.eNk, d1 = a / (p.eNl * p.eNk);
    if (i < p.out_lo || j < p.out_lo) return 0.f;
    int si = i + p.ei0 + tk1, sj = j + p.ej0 + tl1;
    if (si < 0 || si >= p.Nx || sj < 0 || sj >= p.Ny) return 0.f;
    long long off = (b * p.src_ch + d1) * plane + (long long)si * p.Ny + sj;
    return __ldg(p.a0 + off) - __ldg(p.a1 + off);
  }
  long long off = (b * p.src_ch + a) * plane + (long long)i * p.Ny + j;
  float v = __ldg(p.a0 + off);
  if (p.amode == A_DIFF) v -= __ldg(p.a1 + off);
  return v;
}

template <int NK, int NL, int AB, int XB>
__global__ void __launch_bounds__((AB / 8) * XB * NK * WTI) wgrad_kernel(WgradParams p) {
  constexpr int NT = (AB / 8) * XB * NK;  // tasks
  constexpr int THREADS = NT * WTI;
  constexpr int HI = WTI + NK - 1;
  constexpr int PJ = ((WTJ + NL - 1) + 3) / 4 * 4;
  constexpr int XCP0 = HI * PJ;
  constexpr int XCP = XCP0 + ((8 - XCP0 % 32) + 32) % 32;  // channel pitch == 8 (mod 32): spreads (x,tk) over banks
  constexpr int XV = (4 + NL - 1 + 3) / 4;
  constexpr int A_FLOATS = AB * WTI * WTJ;
  extern __shared__ __align__(16) float smem[];
  float* as = smem;             // [AB][WTI][WTJ]
  float* xs = smem + A_FLOATS;  // [XB][XCP] rows of PJ (XB * XCP floats)

  const int tid = threadIdx.x;
  const int g = tid / NT, task = tid % NT;
  const int ag = task % (AB / 8);
  const int xl = (task / (AB / 8)) % XB;
  const int tk = task / ((AB / 8) * XB);
  const int chunk = blockIdx.x;
  const int ab = blockIdx.y / p.nXB, xb = blockIdx.y % p.nXB;
  const int a_base = ab * AB, x_base = xb * XB;
  const bool stats = (xb == 0) && (xl == 0) && (tk == 0);
  const long long plane = (long long)p.Nx * p.Ny;

  float acc[8][NL];
  float sA[8];
  float sq = 0.f;
#pragma unroll
  for (int a = 0; a < 8; a++) {
    sA[a] = 0.f;
#pragma unroll
    for (int l = 0; l < NL; l++) acc[a][l] = 0.f;
  }

  const long long t0 = p.n_tiles * chunk / p.n_chunks, t1 = p.n_tiles * (chunk + 1) / p.n_chunks;
  for (long long t = t0; t < t1; t++) {
    const int tile_j = (int)(t % p.tiles_j);
    const int tile_i = (int)((t / p.tiles_j) % p.tiles_i);
    const long long b = t / ((long long)p.tiles_j * p.tiles_i);
    const int i0 = tile_i * WTI, j0 = tile_j * WTJ;
    // ---- stage A tile ----
    for (int idx = tid; idx < A_FLOATS; idx += THREADS) {
      int col = idx % WTJ, r = (idx / WTJ) % WTI, a = idx / (WTJ * WTI);
      int i = i0 + r, j = j0 + col;
      float v = 0.f;
      if (a_base + a < p.nA && i < p.Nx && j < p.Ny) v = load_A(p, b, a_base + a, i, j);
      as[idx] = v;
    }
    // ---- stage X tile (+halo) ----
    for (int idx = tid; idx < XB * HI * PJ; idx += THREADS) {
      int col = idx % PJ, r = (idx / PJ) % HI, x = idx / (PJ * HI);
      int si = i0 + p.ai0 + r, sj = j0 + p.aj0 + col;
      float v = 0.f;
      if (x_base + x < p.nX && si >= p.lo && si < p.Nx && sj >= p.lo && sj < p.Ny)
        v = __ldg(p.X + (b * p.nX + x_base + x) * plane + (long long)si * p.Ny + sj);
      xs[x * XCP + r * PJ + col] = v;
    }
    __syncthreads();
    const float4* arow = reinterpret_cast<const float4*>(as + (ag * 8) * (WTI * WTJ) + g * WTJ);
    const float4* xrow = reinterpret_cast<const float4*>(xs + xl * XCP + (g + tk) * PJ);
#pragma unroll 2
    for (int jj = 0; jj < WTJ / 4; jj++) {
      float x[4 * XV];
#pragma unroll
      for (int v = 0; v < XV; v++) {
        float4 tv = xrow[jj + v];
        x[4 * v] = tv.x; x[4 * v + 1] = tv.y; x[4 * v + 2] = tv.z; x[4 * v + 3] = tv.w;
      }
#pragma unroll
      for (int a = 0; a < 8; a++) {
        float4 av = arow[a * (WTI * WTJ / 4) + jj];
#pragma unroll
        for (int l = 0; l < NL; l++) {
          acc[a][l] = fmaf(av.x, x[l], acc[a][l]);
          acc[a][l] = fmaf(av.y, x[l + 1], acc[a][l]);
          acc[a][l] = fmaf(av.z, x[l + 2], acc[a][l]);
          acc[a][l] = fmaf(av.w, x[l + 3], acc[a][l]);
        }
        if (stats) {
          sA[a] += (av.x + av.y) + (av.z + av.w);
          sq = fmaf(av.x, av.x, sq); sq = fmaf(av.y, av.y, sq);
          sq = fmaf(av.z, av.z, sq); sq = fmaf(av.w, av.w, sq);
        }
      }
    }
    __syncthreads();
  }

  // ---- tree-reduce the WTI row groups through shared memory (reusing the tile buffers) ----
  constexpr int PER = 8 * NL + 9;  // acc + sA + sq
  float* red = smem;               // needs (WTI/2)*NT*PER floats
#pragma unroll
  for (int half = WTI / 2; half >= 1; half >>= 1) {
    if (g >= half && g < 2 * half) {
      float* dst = red + ((g - half) * NT + task) * PER;
#pragma unroll
      for (int a = 0; a < 8; a++) {
#pragma unroll
        for (int l = 0; l < NL; l++) dst[a * NL + l] = acc[a][l];
        dst[8 * NL + a] = sA[a];
      }
      dst[8 * NL + 8] = sq;
    }
    __syncthreads();
    if (g < half) {
      const float* src = red + (g * NT + task) * PER;
#pragma unroll
      for (int a = 0; a < 8; a++) {
#pragma unroll
        for (int l = 0; l < NL; l++) acc[a][l] += src[a * NL + l];
        sA[a] += src[8 * NL + a];
      }
      sq += src[8 * NL + 8];
    }
    __syncthreads();
  }
  if (g == 0) {
    const int x = x_base + xl;
    const int k = p.flip ? NK - 1 - tk : tk;
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int ga = a_base + ag * 8 + a;
      if (ga < p.nA && x < p.nX) {
        float* dst = p.part + (((long long)chunk * p.nA + ga) * p.nX + x) * (NK * NL) + k * NL;
#pragma unroll
        for (int tl = 0; tl < NL; tl++) dst[p.flip ? NL - 1 - tl : tl] = acc[a][tl];
      }
      if (stats && ga < p.nA) p.part_sum[(long long)chunk * p.nA + ga] = sA[a];
    }
    if (stats) {
      // one value per (chunk, ab, ag): summed by reduce_partials
      p.part_sq[((long long)chunk * gridDim.y / p.nXB + ab) * (AB / 8) + ag] = sq;
    }
  }
}

// Generic fallback (any tap shape): one CTA per output element, block reduction over all frames/pixels.
__global__ void wgrad_generic_kernel(WgradParams p, int Nk, int Nl, long long B, float* G, float* sumA, float* sumsq) {
  const int n_out = p.nA * p.nX * Nk * Nl;
  const int o = blockIdx.x;
  __shared__ double red[256];
  double s = 0.0;
  const long long plane = (long long)p.Nx * p.Ny;
  if (o < n_out) {
    int l = o % Nl, k = (o / Nl) % Nk, x = (o / (Nl * Nk)) % p.nX, a = o / (Nl * Nk * p.nX);
    int tk = p.flip ? Nk - 1 - k : k, tl = p.flip ? Nl - 1 - l : l;
    for (long long n = threadIdx.x; n < B * plane; n += blockDim.x) {
      long long b = n / plane;
      int i = (n % plane) / p.Ny, j = n % p.Ny;
      int si = i + p.ai0 + tk, sj = j + p.aj0 + tl;
      if (si < p.lo || si >= p.Nx || sj < p.lo || sj >= p.Ny) continue;
      s += (double)load_A(p, b, a, i, j) * (double)p.X[(b * p.nX + x) * plane + (long long)si * p.Ny + sj];
    }
  } else if (o < n_out + p.nA) {
    int a = o - n_out;
    for (long long n = threadIdx.x; n < B * plane; n += blockDim.x)
      s += (double)load_A(p, n / plane, a, (n % plane) / p.Ny, n % p.Ny);
  } else {
    for (int a = 0; a < p.nA; a++)
      for (long long n = threadIdx.x; n < B * plane; n += blockDim.x) {
        double v = load_A(p, n / plane, a, (n % plane) / p.Ny, n % p.Ny);
        s += v * v;
      }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (o < n_out) G[o] = (float)red[0];
    else if (o < n_out + p.nA) { if (sumA) sumA[o - n_out] = (float)red[0]; }
    else if (sumsq) *sumsq = (float)red[0];
  }
}

// G[idx] = sum_chunk part[chunk][idx]  (fixed order, double accumulation -> deterministic)
__global__ void reduce_partials_kernel(const float* __restrict__ part, int n_chunks, long long n, float* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  double s = 0.0;
  for (int c = 0; c < n_chunks; c++) s += (double)part[(long long)c * n + idx];
  out[idx] = (float)s;
}
__global__ void reduce_scalar_kernel(const float* __restrict__ part, long long n, float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += (double)part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)red[0];
}

template <int NK, int NL, int AB, int XB>
static int run_wgrad(aefft_ctx* ctx, WgradParams& p, float* G, float* sumA, float* sumsq) {
  constexpr int NT = (AB / 8) * XB * NK;
  constexpr int THREADS = NT * WTI;
  constexpr int HI = WTI + NK - 1;
  constexpr int PJ = ((WTJ + NL - 1) + 3) / 4 * 4;
  constexpr int XCP0 = HI * PJ;
  constexpr int XCP = XCP0 + ((8 - XCP0 % 32) + 32) % 32;
  constexpr int PER = 8 * NL + 9;
  size_t tile_bytes = (size_t)(AB * WTI * WTJ + XB * XCP) * sizeof(float);
  size_t red_bytes = (size_t)(WTI / 2) * NT * PER * sizeof(float);
  size_t smem = tile_bytes > red_bytes ? tile_bytes : red_bytes;
  const int nAB = (p.nA + AB - 1) / AB;
  p.nXB = (p.nX + XB - 1) / XB;
  // ~4 CTAs per SM in flight, but never more chunks than tiles
  long long want = (4LL * ctx->sm_count + (long long)nAB * p.nXB - 1) / ((long long)nAB * p.nXB);
  if (want > p.n_tiles) want = p.n_tiles;
  if (want < 1) want = 1;
  p.n_chunks = (int)want;
  const long long n_out = (long long)p.nA * p.nX * NK * NL;
  const long long n_sq = (long long)p.n_chunks * nAB * (AB / 8);
  float *part, *part_sum, *part_sq;
  AE_TRY(ctx->getT("wgrad_part", (size_t)p.n_chunks * n_out, &part));
  AE_TRY(ctx->getT("wgrad_part_sum", (size_t)p.n_chunks * p.nA, &part_sum));
  AE_TRY(ctx->getT("wgrad_part_sq", (size_t)n_sq, &part_sq));
  p.part = part; p.part_sum = part_sum; p.part_sq = part_sq;
  auto kern = wgrad_kernel<NK, NL, AB, XB>;
  AE_TRY(ctx->ensure_dyn_smem((const void*)kern, smem));
  dim3 grid(p.n_chunks, nAB * p.nXB);
  {
    const double px = (double)p.n_tiles / ((double)p.tiles_i * p.tiles_j) * p.Nx * p.Ny;  // B * Nx * Ny
    ProfScope prof(ctx, "wgrad", 2.0 * px * p.nA * p.nX * NK * NL,
                   4.0 * px * ((double)p.src_ch * (p.amode == A_PLAIN ? 1 : 2) + p.nX));
    kern<<<grid, THREADS, smem, ctx->stream>>>(p);
  }
  AE_CUDA(cudaGetLastError());
  reduce_partials_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, ctx->stream>>>(part, p.n_chunks, n_out, G);
  ctx->launches += 2;
  if (sumA) {
    reduce_partials_kernel<<<(p.nA + 255) / 256, 256, 0, ctx->stream>>>(part_sum, p.n_chunks, p.nA, sumA);
    ctx->launches++;
  }
  if (sumsq) {
    reduce_scalar_kernel<<<1, 256, 0, ctx->stream>>>(part_sq, n_sq, sumsq);
    ctx->launches++;
  }
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

template <int NK, int NL>
static int dispatch_block(aefft_ctx* ctx, WgradParams& p, float* G, float* sumA, float* sumsq) {
  if (p.nX == 1) return run_wgrad<NK, NL, 32, 1>(ctx, p, G, sumA, sumsq);
  if (p.nX % 4 != 0 && p.nX % 3 == 0) return run_wgrad<NK, NL, 16, 3>(ctx, p, G, sumA, sumsq);
  return run_wgrad<NK, NL, 16, 4>(ctx, p, G, sumA, sumsq);
}

int launch_wgrad(aefft_ctx* ctx, const Window& win, int64_t B, int Nx, int Ny, const AOperand& A, const float* X,
                 int nX, float* G, float* sumA, float* sumsq) {
  AE_ARG(B > 0 && Nx > 0 && Ny > 0 && A.nA > 0 && nX > 0);
  WgradParams p;
  p.a0 = A.a0; p.a1 = A.a1; p.X = X;
  p.part = nullptr; p.part_sum = nullptr; p.part_sq = nullptr;
  p.amode = A.mode; p.nA = A.nA; p.src_ch = A.mode == A_SHIFT ? A.src_ch : A.nA; p.nX = nX; p.Nx = Nx; p.Ny = Ny;
  p.ai0 = win.ai0; p.aj0 = win.aj0; p.flip = win.flip; p.lo = win.lo;
  p.ei0 = A.ei0; p.ej0 = A.ej0; p.eNk = A.eNk; p.eNl = A.eNl; p.out_lo = A.out_lo;
  p.tiles_i = (Nx + WTI - 1) / WTI;
  p.tiles_j = (Ny + WTJ - 1) / WTJ;
  p.n_tiles = (long long)B * p.tiles_i * p.tiles_j;
  p.n_chunks = 1; p.nXB = 1;
  if (win.Nk == 5 && win.Nl == 5) return dispatch_block<5, 5>(ctx, p, G, sumA, sumsq);
  if (win.Nk == 3 && win.Nl == 3) return dispatch_block<3, 3>(ctx, p, G, sumA, sumsq);
  if (win.Nk == 7 && win.Nl == 7) return dispatch_block<7, 7>(ctx, p, G, sumA, sumsq);
  const int n_out = A.nA * nX * win.Nk * win.Nl;
  wgrad_generic_kernel<<<n_out + A.nA + 1, 256, 0, ctx->stream>>>(p, win.Nk, win.Nl, B, G, sumA, sumsq);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

}  // namespace aefft
