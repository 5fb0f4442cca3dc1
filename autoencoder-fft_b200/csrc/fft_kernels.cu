// Hand-written batched 2-D real-to-complex / complex-to-real FFT (power-of-two Nx, Ny), fp32, unnormalised both ways.
// Replaces the cufftPlanMany + cufftExecR2C / cufftExecC2R call sites of the reference
// (fft_backproplib.cu:779/796, 821/829, 885/910, 937/946, 1208-1209/1219-1220/1281-1282), which also re-create and
// destroy their plans on every call.
//
// Layout: real images [batch][Nx][Ny] (j fastest), half spectra [batch][Nx][Ny/2+1] complex64 -- cuFFT's n={Nx,Ny}.
// A 2-D transform is two HBM passes (read once + write once each):
//   rows : one CTA stages RP row PAIRS in shared memory; rows a,b are transformed together as z = a + i b
//          ("two for one"), then split with X_a[k] = (Z[k] + conj Z[N-k])/2, X_b[k] = (Z[k] - conj Z[N-k])/(2i);
//   cols : one CTA stages an [Nx][CT] tile of CT adjacent columns (CT*8 B contiguous per row), transforms every
//          column in shared memory and writes the tile back in place.
// The in-shared-memory transform is an in-place mixed-radix (8/4/2) decimation-in-time FFT: the digit-reversal
// permutation is applied while staging, each pass keeps its r points in registers, and twiddles come from a
// per-length table computed in double precision on the host (L1-resident via __ldg).
#include <cmath>

#include "common.cuh"

namespace aefft {

struct FftPlan {
  int N, log2N;
  int npass;
  int radix[6];
};

static FftPlan make_plan(int N) {
  FftPlan p;
  p.N = N;
  p.log2N = 0;
  while ((1 << p.log2N) < N) p.log2N++;
  p.npass = 0;
  int rem = p.log2N;
  // small radix first (it runs with the smallest stride, where bank conflicts are worst), radix 8 for the rest
  if (rem % 3 == 1) { p.radix[p.npass++] = 2; rem -= 1; }
  else if (rem % 3 == 2) { p.radix[p.npass++] = 4; rem -= 2; }
  while (rem > 0) { p.radix[p.npass++] = 8; rem -= 3; }
  return p;
}

// position of input sample n in the staged (digit-reversed) order: the LAST pass splits n by its radix first
__device__ __forceinline__ int digit_reverse(const FftPlan& p, int n) {
  int pos = 0, M = p.N;
#pragma unroll 1
  for (int i = p.npass - 1; i >= 0; i--) {
    const int r = p.radix[i];
    M /= r;
    pos += (n & (r - 1)) * M;
    n /= r;
  }
  return pos;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (DIR=-1, forward) or +i (DIR=+1, inverse)
template <int DIR>
__device__ __forceinline__ float2 mul_mi(float2 a) { return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x); }

template <int DIR>
__device__ __forceinline__ void dft2(float2* a) {
  float2 t = a[0];
  a[0] = cadd(t, a[1]);
  a[1] = csub(t, a[1]);
}
template <int DIR>
__device__ __forceinline__ void dft4(float2* a) {
  float2 s0 = cadd(a[0], a[2]), d0 = csub(a[0], a[2]);
  float2 s1 = cadd(a[1], a[3]), d1 = mul_mi<DIR>(csub(a[1], a[3]));
  a[0] = cadd(s0, s1);
  a[2] = csub(s0, s1);
  a[1] = cadd(d0, d1);
  a[3] = csub(d0, d1);
}
template <int DIR>
__device__ __forceinline__ void dft8(float2* a) {
  // two 4-point DFTs on even / odd samples, then the radix-2 combine with W8^j
  float2 e[4] = {a[0], a[2], a[4], a[6]};
  float2 o[4] = {a[1], a[3], a[5], a[7]};
  dft4<DIR>(e);
  dft4<DIR>(o);
  const float h = 0.70710678118654752440f;
  // W8^1 = (h, -h) fwd / (h, +h) inv ; W8^2 = -i / +i ; W8^3 = (-h, -h) fwd / (-h, +h) inv
  float2 w1 = DIR < 0 ? make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x))
                      : make_float2(h * (o[1].x - o[1].y), h * (o[1].y + o[1].x));
  float2 w2 = mul_mi<DIR>(o[2]);
  float2 w3 = DIR < 0 ? make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y))
                      : make_float2(-h * (o[3].x + o[3].y), h * (o[3].x - o[3].y));
  a[0] = cadd(e[0], o[0]); a[4] = csub(e[0], o[0]);
  a[1] = cadd(e[1], w1);   a[5] = csub(e[1], w1);
  a[2] = cadd(e[2], w2);   a[6] = csub(e[2], w2);
  a[3] = cadd(e[3], w3);   a[7] = csub(e[3], w3);
}

// One radix-R pass over `nseq` sequences held in shared memory: element n of sequence c lives at s[n*sn + c*sc].
// seq_fast: consecutive threads take consecutive sequences (column tiles) or consecutive butterflies (row tiles).
template <int DIR, int R>
__device__ __forceinline__ void fft_pass(float2* s, int N, int M, int nseq, int sn, int sc, bool seq_fast,
                                         const float2* __restrict__ tw) {
  const int L = M * R;
  const int nbf = N / R;
  const int tstep = N / L;  // twiddle W_L^x = W_N^(x * N/L)
  for (int item = threadIdx.x; item < nbf * nseq; item += blockDim.x) {
    int c, t;
    if (seq_fast) { c = item % nseq; t = item / nseq; }
    else { t = item % nbf; c = item / nbf; }
    const int k = t % M, blk = t / M;
    float2* base = s + (size_t)(blk * L + k) * sn + (size_t)c * sc;
    float2 a[R];
#pragma unroll
    for (int q = 0; q < R; q++) a[q] = base[(size_t)q * M * sn];
    if (M > 1) {
#pragma unroll
      for (int q = 1; q < R; q++) {
        float2 w = __ldg(tw + (q * k * tstep));
        if (DIR > 0) w.y = -w.y;
        a[q] = cmul(a[q], w);
      }
    }
    if (R == 2) dft2<DIR>(a);
    else if (R == 4) dft4<DIR>(a);
    else dft8<DIR>(a);
#pragma unroll
    for (int j = 0; j < R; j++) base[(size_t)j * M * sn] = a[j];
  }
}

template <int DIR>
__device__ __forceinline__ void fft_smem(float2* s, const FftPlan& p, int nseq, int sn, int sc, bool seq_fast,
                                         const float2* __restrict__ tw) {
  int M = 1;
#pragma unroll 1
  for (int i = 0; i < p.npass; i++) {
    __syncthreads();
    const int r = p.radix[i];
    if (r == 2) fft_pass<DIR, 2>(s, p.N, M, nseq, sn, sc, seq_fast, tw);
    else if (r == 4) fft_pass<DIR, 4>(s, p.N, M, nseq, sn, sc, seq_fast, tw);
    else fft_pass<DIR, 8>(s, p.N, M, nseq, sn, sc, seq_fast, tw);
    M *= r;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ row passes
// R2C rows: real [img][Nx][Ny] -> complex [img][Nx][Nyr]
__global__ void __launch_bounds__(256) fft_rows_r2c_kernel(const float* __restrict__ in, float2* __restrict__ out, int Nx,
                                                           int Ny, int RP, FftPlan plan, const float2* __restrict__ tw) {
  extern __shared__ __align__(16) float2 sm[];
  const int Nyr = Ny / 2 + 1;
  const int SP = Ny + 1;  // row pitch in complex elements (odd -> rows start in different banks)
  const long long img = blockIdx.y;
  const int rp0 = blockIdx.x * RP;
  const float* src = in + img * (long long)Nx * Ny;
  for (int idx = threadIdx.x; idx < RP * Ny; idx += blockDim.x) {
    const int r = idx / Ny, n = idx % Ny;
    const int row = 2 * (rp0 + r);
    float2 v = make_float2(0.f, 0.f);
    if (row < Nx) {
      v.x = __ldg(src + (long long)row * Ny + n);
      v.y = __ldg(src + (long long)(row + 1) * Ny + n);
    }
    sm[r * SP + digit_reverse(plan, n)] = v;
  }
  fft_smem<-1>(sm, plan, RP, 1, SP, false, tw);
  float2* dst = out + img * (long long)Nx * Nyr;
  for (int idx = threadIdx.x; idx < RP * Nyr; idx += blockDim.x) {
    const int r = idx / Nyr, k = idx % Nyr;
    const int row = 2 * (rp0 + r);
    if (row >= Nx) continue;
    const float2 z1 = sm[r * SP + k];
    const float2 z2 = sm[r * SP + ((Ny - k) & (Ny - 1))];
    // Xa = (Z[k] + conj Z[N-k]) / 2 ;  Xb = (Z[k] - conj Z[N-k]) / (2i)
    dst[(long long)row * Nyr + k] = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
    dst[(long long)(row + 1) * Nyr + k] = make_float2(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
  }
}

// C2R rows: complex [img][Nx][Nyr] -> real [img][Nx][Ny], times `scale`.  Imaginary parts of the DC and Nyquist bins
// of each row are ignored, as a Hermitian C2R does.
__global__ void __launch_bounds__(256) fft_rows_c2r_kernel(const float2* __restrict__ in, float* __restrict__ out, int Nx,
                                                           int Ny, int RP, FftPlan plan, const float2* __restrict__ tw,
                                                           float scale) {
  extern __shared__ __align__(16) float2 sm[];
  const int Nyr = Ny / 2 + 1;
  const int SP = Ny + 1;
  const long long img = blockIdx.y;
  const int rp0 = blockIdx.x * RP;
  const float2* src = in + img * (long long)Nx * Nyr;
  for (int idx = threadIdx.x; idx < RP * Nyr; idx += blockDim.x) {
    const int r = idx / Nyr, k = idx % Nyr;
    const int row = 2 * (rp0 + r);
    float2 a = make_float2(0.f, 0.f), b = a;
    if (row < Nx) {
      a = __ldg(src + (long long)row * Nyr + k);
      b = __ldg(src + (long long)(row + 1) * Nyr + k);
    }
    if (k == 0 || k == Ny / 2) { a.y = 0.f; b.y = 0.f; }
    // Z[k] = A + iB ; Z[N-k] = conj A + i conj B
    sm[r * SP + digit_reverse(plan, k)] = make_float2(a.x - b.y, a.y + b.x);
    if (k > 0 && k < Ny / 2) sm[r * SP + digit_reverse(plan, Ny - k)] = make_float2(a.x + b.y, b.x - a.y);
  }
  fft_smem<+1>(sm, plan, RP, 1, SP, false, tw);
  float* dst = out + img * (long long)Nx * Ny;
  for (int idx = threadIdx.x; idx < RP * Ny; idx += blockDim.x) {
    const int r = idx / Ny, n = idx % Ny;
    const int row = 2 * (rp0 + r);
    if (row >= Nx) continue;
    const float2 z = sm[r * SP + n];
    dst[(long long)row * Ny + n] = z.x * scale;
    dst[(long long)(row + 1) * Ny + n] = z.y * scale;
  }
}

// ------------------------------------------------------------------------------------------------ column pass
// complex [img][Nx][W] -> complex [img][Nx][W] (in place allowed), FFT along Nx for every column.
template <int DIR>
__global__ void __launch_bounds__(256) fft_cols_kernel(const float2* __restrict__ in, float2* __restrict__ out, int Nx,
                                                       int W, int CT, FftPlan plan, const float2* __restrict__ tw) {
  extern __shared__ __align__(16) float2 sm[];
  const long long img = blockIdx.y;
  const int c0 = blockIdx.x * CT;
  const float2* src = in + img * (long long)Nx * W;
  const int SC = CT | 1;  // odd pitch between samples: consecutive n land in different banks
  for (int idx = threadIdx.x; idx < Nx * CT; idx += blockDim.x) {
    const int n = idx / CT, c = idx % CT;
    float2 v = make_float2(0.f, 0.f);
    if (c0 + c < W) v = src[(long long)n * W + c0 + c];
    sm[digit_reverse(plan, n) * SC + c] = v;
  }
  fft_smem<DIR>(sm, plan, CT, SC, 1, true, tw);
  float2* dst = out + img * (long long)Nx * W;
  for (int idx = threadIdx.x; idx < Nx * CT; idx += blockDim.x) {
    const int n = idx / CT, c = idx % CT;
    if (c0 + c < W) dst[(long long)n * W + c0 + c] = sm[n * SC + c];
  }
}

// ------------------------------------------------------------------------------------------------ host side
static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

static int get_twiddles(aefft_ctx* ctx, int N, const float2** out) {
  char name[32];
  snprintf(name, sizeof(name), "fft_tw_%d", N);
  auto it = ctx->scratch.find(name);
  if (it != ctx->scratch.end() && it->second.p) {
    *out = (const float2*)it->second.p;
    return AEFFT_OK;
  }
  float2* dev;
  AE_TRY(ctx->getT(name, (size_t)N, &dev));
  std::vector<float2> h(N);
  for (int t = 0; t < N; t++) {
    double a = -2.0 * M_PI * (double)t / (double)N;
    h[t] = make_float2((float)cos(a), (float)sin(a));
  }
  AE_CUDA(cudaMemcpyAsync(dev, h.data(), (size_t)N * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
  AE_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = dev;
  return AEFFT_OK;
}

static int col_tile(int Nx) {
  int ct = 16384 / Nx;  // <= 128 KB (+ padding) of shared memory per CTA
  if (ct > 16) ct = 16;
  if (ct < 1) ct = 1;
  return ct;
}
static int row_pairs(int Nx, int Ny) {
  int rp = 2048 / Ny;
  if (rp < 1) rp = 1;
  if (rp > Nx / 2) rp = Nx / 2;
  return rp;
}

template <class K>
static int set_smem(K kern, size_t bytes) {
  AE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return AEFFT_OK;
}

int launch_fft_r2c(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, const float* in, float2* spec) {
  AE_ARG(batch > 0 && is_pow2(Nx) && is_pow2(Ny) && Nx >= 2 && Ny >= 2 && Nx <= 8192 && Ny <= 8192);
  AE_ARG(batch <= 65535);
  const int Nyr = Ny / 2 + 1;
  const float2 *twx, *twy;
  AE_TRY(get_twiddles(ctx, Nx, &twx));
  AE_TRY(get_twiddles(ctx, Ny, &twy));
  const double px = (double)batch * Nx * Ny, sp = (double)batch * Nx * Nyr;
  {
    const int RP = row_pairs(Nx, Ny);
    const size_t smem = (size_t)RP * (Ny + 1) * sizeof(float2);
    AE_TRY(set_smem(fft_rows_r2c_kernel, smem));
    dim3 grid((Nx / 2 + RP - 1) / RP, (unsigned)batch);
    ProfScope prof(ctx, "fft_rows_r2c", 2.5 * px * log2((double)Ny), 4.0 * px + 8.0 * sp);
    fft_rows_r2c_kernel<<<grid, 256, smem, ctx->stream>>>(in, spec, Nx, Ny, RP, make_plan(Ny), twy);
  }
  {
    const int CT = col_tile(Nx);
    const size_t smem = (size_t)Nx * (CT | 1) * sizeof(float2);
    AE_TRY(set_smem(fft_cols_kernel<-1>, smem));
    dim3 grid((Nyr + CT - 1) / CT, (unsigned)batch);
    ProfScope prof(ctx, "fft_cols", 5.0 * sp * log2((double)Nx), 16.0 * sp);
    fft_cols_kernel<-1><<<grid, 256, smem, ctx->stream>>>(spec, spec, Nx, Nyr, CT, make_plan(Nx), twx);
  }
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// spec is NOT modified: the column pass writes into `work` (batch*Nx*Nyr complex).
int launch_fft_c2r(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, const float2* spec, float2* work, float* out,
                   float scale) {
  AE_ARG(batch > 0 && is_pow2(Nx) && is_pow2(Ny) && Nx >= 2 && Ny >= 2 && Nx <= 8192 && Ny <= 8192);
  AE_ARG(batch <= 65535);
  const int Nyr = Ny / 2 + 1;
  const float2 *twx, *twy;
  AE_TRY(get_twiddles(ctx, Nx, &twx));
  AE_TRY(get_twiddles(ctx, Ny, &twy));
  const double px = (double)batch * Nx * Ny, sp = (double)batch * Nx * Nyr;
  {
    const int CT = col_tile(Nx);
    const size_t smem = (size_t)Nx * (CT | 1) * sizeof(float2);
    AE_TRY(set_smem(fft_cols_kernel<+1>, smem));
    dim3 grid((Nyr + CT - 1) / CT, (unsigned)batch);
    ProfScope prof(ctx, "fft_cols", 5.0 * sp * log2((double)Nx), 16.0 * sp);
    fft_cols_kernel<+1><<<grid, 256, smem, ctx->stream>>>(spec, work, Nx, Nyr, CT, make_plan(Nx), twx);
  }
  {
    const int RP = row_pairs(Nx, Ny);
    const size_t smem = (size_t)RP * (Ny + 1) * sizeof(float2);
    AE_TRY(set_smem(fft_rows_c2r_kernel, smem));
    dim3 grid((Nx / 2 + RP - 1) / RP, (unsigned)batch);
    ProfScope prof(ctx, "fft_rows_c2r", 2.5 * px * log2((double)Ny), 4.0 * px + 8.0 * sp);
    fft_rows_c2r_kernel<<<grid, 256, smem, ctx->stream>>>(work, out, Nx, Ny, RP, make_plan(Ny), twy, scale);
  }
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

}  // namespace aefft
