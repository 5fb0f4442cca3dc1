// Hand-written batched 2-D real-to-complex / complex-to-real FFT, fp32, unnormalised both ways.  Lengths: even numbers of the
// form 2^a 3^b 5^c up to 8192 (the camera's 640 x 480 and its pooled levels in momentum space); powers of two take
// the compile-time kernels further down, every other length the run-time mixed-radix kernels (radix 8 / 5 / 4 / 3 / 2).
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
#include <cstdlib>

#include "common.cuh"

namespace aefft {

struct FftPlan {
  int N;
  int npass;
  int radix[12];
};

// N = 2^a 3^b 5^c: the odd radices first, then the power-of-two part as one small radix and radix 8 for the rest (the small
// radices run with the smallest strides, where bank conflicts are worst)
static FftPlan make_plan(int N) {
  FftPlan p;
  p.N = N;
  p.npass = 0;
  int rem = N;
  while (rem % 5 == 0) { p.radix[p.npass++] = 5; rem /= 5; }
  while (rem % 3 == 0) { p.radix[p.npass++] = 3; rem /= 3; }
  int l2 = 0;
  while (rem > 1 && rem % 2 == 0) { l2++; rem /= 2; }
  if (l2 % 3 == 1) { p.radix[p.npass++] = 2; l2 -= 1; }
  else if (l2 % 3 == 2) { p.radix[p.npass++] = 4; l2 -= 2; }
  while (l2 > 0) { p.radix[p.npass++] = 8; l2 -= 3; }
  return p;
}
// even, only factors 2, 3, 5, at most 8192
static bool fft_len_ok(int N) {
  if (N < 2 || N > 8192 || (N & 1)) return false;
  int r = N;
  while (r % 2 == 0) r /= 2;
  while (r % 3 == 0) r /= 3;
  while (r % 5 == 0) r /= 5;
  return r == 1;
}

// x / d for 0 <= x < 2^20 through the float reciprocal inv = 1 / d: (x + 0.5) * inv stays at least 0.5 / d away from an integer
// while its rounding error is below x / d * 2^-22, so the truncation is exact.  The run-time (mixed-radix) kernels index
// through these instead of ~20-instruction integer divisions.
__device__ __forceinline__ int idiv_f(int x, float inv) { return (int)(((float)x + 0.5f) * inv); }
// position of input sample n in the staged (digit-reversed) order: the LAST pass splits n by its radix first
__device__ __forceinline__ int digit_reverse(const FftPlan& p, int n) {
  int pos = 0, M = p.N;
#pragma unroll 1
  for (int i = p.npass - 1; i >= 0; i--) {
    const int r = p.radix[i];
    const float ir = 1.f / (float)r;
    M = idiv_f(M, ir);
    const int q = idiv_f(n, ir);
    pos += (n - q * r) * M;
    n = q;
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
// W3 = exp(-+ 2 pi i / 3) = (-1/2, -+ sqrt(3)/2)
template <int DIR>
__device__ __forceinline__ void dft3(float2* a) {
  const float s3 = 0.86602540378443864676f;
  const float2 t = cadd(a[1], a[2]), d = mul_mi<DIR>(csub(a[1], a[2]));  // d = -+ i (a1 - a2)
  const float2 m = make_float2(a[0].x - 0.5f * t.x, a[0].y - 0.5f * t.y);
  a[0] = cadd(a[0], t);
  a[1] = make_float2(m.x + s3 * d.x, m.y + s3 * d.y);
  a[2] = make_float2(m.x - s3 * d.x, m.y - s3 * d.y);
}
// 5-point DFT (Rader / Winograd-style pairing): X[k] = a0 + sum_j a_j W5^(jk)
template <int DIR>
__device__ __forceinline__ void dft5(float2* a) {
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;  // cos(2 pi / 5), cos(4 pi / 5)
  const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;   // sin(2 pi / 5), sin(4 pi / 5)
  const float2 p1 = cadd(a[1], a[4]), p2 = cadd(a[2], a[3]);
  const float2 q1 = mul_mi<DIR>(csub(a[1], a[4])), q2 = mul_mi<DIR>(csub(a[2], a[3]));  // -+ i (a1 - a4), -+ i (a2 - a3)
  const float2 a0 = a[0];
  const float2 u1 = make_float2(a0.x + c1 * p1.x + c2 * p2.x, a0.y + c1 * p1.y + c2 * p2.y);
  const float2 u2 = make_float2(a0.x + c2 * p1.x + c1 * p2.x, a0.y + c2 * p1.y + c1 * p2.y);
  const float2 v1 = make_float2(s1 * q1.x + s2 * q2.x, s1 * q1.y + s2 * q2.y);
  const float2 v2 = make_float2(s2 * q1.x - s1 * q2.x, s2 * q1.y - s1 * q2.y);
  a[0] = make_float2(a0.x + p1.x + p2.x, a0.y + p1.y + p2.y);
  a[1] = cadd(u1, v1);
  a[4] = csub(u1, v1);
  a[2] = cadd(u2, v2);
  a[3] = csub(u2, v2);
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

// a * (wr + i wi) with (wr, wi) the FORWARD twiddle; the inverse transform takes the conjugate
template <int DIR>
__device__ __forceinline__ float2 mul_w(float2 a, float wr, float wi) {
  const float y = DIR < 0 ? wi : -wi;
  return make_float2(a.x * wr - a.y * y, a.x * y + a.y * wr);
}
// 16-point DFT in registers as 4 x 4 (n = 4 n1 + n2, k = k1 + 4 k2), natural order in and out
template <int DIR>
__device__ __forceinline__ void dft16(float2* a) {
  const float c = 0.92387953251128674f, s = 0.38268343236508977f, h = 0.70710678118654752440f;
  float2 t[4][4];
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) {
    t[n2][0] = a[n2]; t[n2][1] = a[4 + n2]; t[n2][2] = a[8 + n2]; t[n2][3] = a[12 + n2];
    dft4<DIR>(t[n2]);
  }
  // W16^(n2 k1), forward values (cos, -sin)
  t[1][1] = mul_w<DIR>(t[1][1], c, -s);  t[1][2] = mul_w<DIR>(t[1][2], h, -h);  t[1][3] = mul_w<DIR>(t[1][3], s, -c);
  t[2][1] = mul_w<DIR>(t[2][1], h, -h);  t[2][2] = mul_mi<DIR>(t[2][2]);        t[2][3] = mul_w<DIR>(t[2][3], -h, -h);
  t[3][1] = mul_w<DIR>(t[3][1], s, -c);  t[3][2] = mul_w<DIR>(t[3][2], -h, -h); t[3][3] = mul_w<DIR>(t[3][3], -c, s);
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) {
    float2 y[4] = {t[0][k1], t[1][k1], t[2][k1], t[3][k1]};
    dft4<DIR>(y);
    a[k1] = y[0]; a[k1 + 4] = y[1]; a[k1 + 8] = y[2]; a[k1 + 12] = y[3];
  }
}

// One radix-R pass over `nseq` sequences held in shared memory: element n of sequence c lives at s[n*sn + c*sc].
// seq_fast: consecutive threads take consecutive sequences (column tiles) or consecutive butterflies (row tiles).
template <int DIR, int R>
__device__ __forceinline__ void fft_pass(float2* s, int N, int M, int nseq, int sn, int sc, bool seq_fast,
                                         const float2* __restrict__ tw) {
  const int L = M * R;
  const int nbf = N / R;
  const int tstep = N / L;  // twiddle W_L^x = W_N^(x * N/L)
  const float inv_seq = 1.f / (float)nseq, inv_nbf = 1.f / (float)nbf, inv_M = 1.f / (float)M;
  for (int item = threadIdx.x; item < nbf * nseq; item += blockDim.x) {
    int c, t;
    if (seq_fast) { t = idiv_f(item, inv_seq); c = item - t * nseq; }
    else { c = idiv_f(item, inv_nbf); t = item - c * nbf; }
    const int blk = idiv_f(t, inv_M), k = t - blk * M;
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
    else if (R == 3) dft3<DIR>(a);
    else if (R == 4) dft4<DIR>(a);
    else if (R == 5) dft5<DIR>(a);
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
    else if (r == 3) fft_pass<DIR, 3>(s, p.N, M, nseq, sn, sc, seq_fast, tw);
    else if (r == 4) fft_pass<DIR, 4>(s, p.N, M, nseq, sn, sc, seq_fast, tw);
    else if (r == 5) fft_pass<DIR, 5>(s, p.N, M, nseq, sn, sc, seq_fast, tw);
    else fft_pass<DIR, 8>(s, p.N, M, nseq, sn, sc, seq_fast, tw);
    M *= r;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ row passes
// R2C rows: real [img][Nx][Ny] -> complex [img][Nx][Nyr]
__global__ void __launch_bounds__(256) fft_rows_r2c_kernel(const float* __restrict__ in, float2* __restrict__ out, int Nx,
                                                           int Ny, int RP, FftPlan plan, const float2* __restrict__ tw, int ch,
                                                           long long fstride) {
  extern __shared__ __align__(16) float2 sm[];
  const int Nyr = Ny / 2 + 1;
  const int SP = Ny + 1;  // row pitch in complex elements (odd -> rows start in different banks)
  const long long img = blockIdx.y;
  const int rp0 = blockIdx.x * RP;
  const float* src = in + (img / ch) * fstride + (img % ch) * (long long)Nx * Ny;
  const float inv_Ny = 1.f / (float)Ny, inv_Nyr = 1.f / (float)Nyr;
  for (int idx = threadIdx.x; idx < RP * Ny; idx += blockDim.x) {
    const int r = idiv_f(idx, inv_Ny), n = idx - r * Ny;
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
    const int r = idiv_f(idx, inv_Nyr), k = idx - r * Nyr;
    const int row = 2 * (rp0 + r);
    if (row >= Nx) continue;
    const float2 z1 = sm[r * SP + k];
    const float2 z2 = sm[r * SP + (k == 0 ? 0 : Ny - k)];
    // Xa = (Z[k] + conj Z[N-k]) / 2 ;  Xb = (Z[k] - conj Z[N-k]) / (2i)
    dst[(long long)row * Nyr + k] = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
    dst[(long long)(row + 1) * Nyr + k] = make_float2(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
  }
}

// C2R rows: complex [img][Nx][Nyr] -> real [img][Nx][Ny], times `scale`.  Imaginary parts of the DC and Nyquist bins
// of each row are ignored, as a Hermitian C2R does.
__global__ void __launch_bounds__(256) fft_rows_c2r_kernel(const float2* __restrict__ in, float* __restrict__ out, int Nx,
                                                           int Ny, int RP, FftPlan plan, const float2* __restrict__ tw,
                                                           float scale, int ch, long long fstride) {
  extern __shared__ __align__(16) float2 sm[];
  const int Nyr = Ny / 2 + 1;
  const int SP = Ny + 1;
  const long long img = blockIdx.y;
  const int rp0 = blockIdx.x * RP;
  const float2* src = in + img * (long long)Nx * Nyr;
  const float inv_Ny = 1.f / (float)Ny, inv_Nyr = 1.f / (float)Nyr;
  for (int idx = threadIdx.x; idx < RP * Nyr; idx += blockDim.x) {
    const int r = idiv_f(idx, inv_Nyr), k = idx - r * Nyr;
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
  float* dst = out + (img / ch) * fstride + (img % ch) * (long long)Nx * Ny;
  for (int idx = threadIdx.x; idx < RP * Ny; idx += blockDim.x) {
    const int r = idiv_f(idx, inv_Ny), n = idx - r * Ny;
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
  const float inv_CT = 1.f / (float)CT;
  for (int idx = threadIdx.x; idx < Nx * CT; idx += blockDim.x) {
    const int n = idiv_f(idx, inv_CT), c = idx - n * CT;
    float2 v = make_float2(0.f, 0.f);
    if (c0 + c < W) v = src[(long long)n * W + c0 + c];
    sm[digit_reverse(plan, n) * SC + c] = v;
  }
  fft_smem<DIR>(sm, plan, CT, SC, 1, true, tw);
  float2* dst = out + img * (long long)Nx * W;
  for (int idx = threadIdx.x; idx < Nx * CT; idx += blockDim.x) {
    const int n = idiv_f(idx, inv_CT), c = idx - n * CT;
    if (c0 + c < W) dst[(long long)n * W + c0 + c] = sm[n * SC + c];
  }
}


// ================================================================================================ compile-time sizes
// The kernels above take the transform length at run time (digit reversal and butterfly indexing through integer
// divisions): they are instruction bound (~0.7-0.9 TB/s).  For N = 8..4096 the same algorithm is instantiated with the
// length, the radix schedule and every stride as compile-time constants (shifts and masks only), smaller tiles (more
// CTAs per SM so that the load, butterfly and store phases of different CTAs overlap) and division-free staging loops.
template <int LOG2N>
struct Sched {
  static constexpr int N = 1 << LOG2N;
  static constexpr int first = (LOG2N % 3 == 1) ? 1 : ((LOG2N % 3 == 2) ? 2 : 3);  // log2 of the first radix
  static constexpr int npass = 1 + (LOG2N - first) / 3;
  __host__ __device__ static constexpr int lr(int i) { return i == 0 ? first : 3; }
};

template <int LOG2N>
__device__ __forceinline__ int digit_reverse_t(int n) {
  using S = Sched<LOG2N>;
  int pos = 0;
  int shift = LOG2N;
#pragma unroll
  for (int i = S::npass - 1; i >= 0; i--) {
    const int lr = S::lr(i);
    shift -= lr;
    pos += (n & ((1 << lr) - 1)) << shift;
    n >>= lr;
  }
  return pos;
}

// element n of sequence c: rows s[c*SP + n], columns s[n*SC + c]
template <int DIR, int LR, int LOG2N, int LOG2M, bool COLS, int LOG2SEQ, int PITCH>
__device__ __forceinline__ void fft_pass_t(float2* s, const float2* __restrict__ tw) {
  constexpr int R = 1 << LR, N = 1 << LOG2N, M = 1 << LOG2M;
  constexpr int LOG2NBF = LOG2N - LR;
  constexpr int tstep = N >> (LOG2M + LR);
  constexpr int items = 1 << (LOG2NBF + LOG2SEQ);
  for (int item = threadIdx.x; item < items; item += blockDim.x) {
    int c, t;
    if (COLS) { c = item & ((1 << LOG2SEQ) - 1); t = item >> LOG2SEQ; }
    else { t = item & ((1 << LOG2NBF) - 1); c = item >> LOG2NBF; }
    const int k = t & (M - 1), blk = t >> LOG2M;
    const int n0 = (blk << (LOG2M + LR)) + k;
    float2* base = COLS ? s + n0 * PITCH + c : s + c * PITCH + n0;
    constexpr int step = COLS ? M * PITCH : M;
    float2 a[R];
#pragma unroll
    for (int q = 0; q < R; q++) a[q] = base[q * step];
    if (LOG2M > 0) {
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
    for (int j = 0; j < R; j++) base[j * step] = a[j];
  }
}

template <int DIR, int LOG2N, bool COLS, int LOG2SEQ, int PITCH, int I = 0, int LOG2M = 0>
__device__ __forceinline__ void fft_smem_t(float2* s, const float2* __restrict__ tw) {
  __syncthreads();
  if constexpr (LOG2M < LOG2N) {
    constexpr int LR = Sched<LOG2N>::lr(I);
    fft_pass_t<DIR, LR, LOG2N, LOG2M, COLS, LOG2SEQ, PITCH>(s, tw);
    fft_smem_t<DIR, LOG2N, COLS, LOG2SEQ, PITCH, I + 1, LOG2M + LR>(s, tw);
  }
}

// row pairs per CTA: ~32 KB of shared memory
template <int LOG2N>
struct RowCfg {
  static constexpr int N = 1 << LOG2N;
  static constexpr int LOG2RP = LOG2N >= 12 ? 0 : (12 - LOG2N > 5 ? 5 : 12 - LOG2N);
  static constexpr int RP = 1 << LOG2RP;
  static constexpr int SP = N + 1;
  static constexpr size_t smem = (size_t)RP * SP * sizeof(float2);
};

template <int LOG2N>
__global__ void __launch_bounds__(256) fft_rows_r2c_t(const float* __restrict__ in, float2* __restrict__ out, int Nx,
                                                      const float2* __restrict__ tw) {
  using C = RowCfg<LOG2N>;
  constexpr int Ny = C::N, Nyr = Ny / 2 + 1, SP = C::SP, RP = C::RP;
  extern __shared__ __align__(16) float2 sm[];
  const long long img = blockIdx.y;
  const int rp0 = blockIdx.x * RP;
  const float* src = in + img * (long long)Nx * Ny;
  for (int idx = threadIdx.x; idx < RP * Ny; idx += blockDim.x) {
    const int r = idx >> LOG2N, n = idx & (Ny - 1);
    const int row = 2 * (rp0 + r);
    float2 v = make_float2(0.f, 0.f);
    if (row < Nx) {
      v.x = __ldg(src + (long long)row * Ny + n);
      v.y = __ldg(src + (long long)(row + 1) * Ny + n);
    }
    sm[r * SP + digit_reverse_t<LOG2N>(n)] = v;
  }
  fft_smem_t<-1, LOG2N, false, C::LOG2RP, SP>(sm, tw);
  float2* dst = out + img * (long long)Nx * Nyr;
  for (int r = 0; r < RP; r++) {
    const int row = 2 * (rp0 + r);
    if (row >= Nx) break;
    for (int k = threadIdx.x; k < Nyr; k += blockDim.x) {
      const float2 z1 = sm[r * SP + k];
      const float2 z2 = sm[r * SP + ((Ny - k) & (Ny - 1))];
      dst[(long long)row * Nyr + k] = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
      dst[(long long)(row + 1) * Nyr + k] = make_float2(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
    }
  }
}

template <int LOG2N>
__global__ void __launch_bounds__(256) fft_rows_c2r_t(const float2* __restrict__ in, float* __restrict__ out, int Nx,
                                                      const float2* __restrict__ tw, float scale) {
  using C = RowCfg<LOG2N>;
  constexpr int Ny = C::N, Nyr = Ny / 2 + 1, SP = C::SP, RP = C::RP;
  extern __shared__ __align__(16) float2 sm[];
  const long long img = blockIdx.y;
  const int rp0 = blockIdx.x * RP;
  const float2* src = in + img * (long long)Nx * Nyr;
  for (int r = 0; r < RP; r++) {
    const int row = 2 * (rp0 + r);
    for (int k = threadIdx.x; k < Nyr; k += blockDim.x) {
      float2 a = make_float2(0.f, 0.f), b = a;
      if (row < Nx) {
        a = __ldg(src + (long long)row * Nyr + k);
        b = __ldg(src + (long long)(row + 1) * Nyr + k);
      }
      if (k == 0 || k == Ny / 2) { a.y = 0.f; b.y = 0.f; }
      sm[r * SP + digit_reverse_t<LOG2N>(k)] = make_float2(a.x - b.y, a.y + b.x);
      if (k > 0 && k < Ny / 2) sm[r * SP + digit_reverse_t<LOG2N>(Ny - k)] = make_float2(a.x + b.y, b.x - a.y);
    }
  }
  fft_smem_t<+1, LOG2N, false, C::LOG2RP, SP>(sm, tw);
  float* dst = out + img * (long long)Nx * Ny;
  for (int idx = threadIdx.x; idx < RP * Ny; idx += blockDim.x) {
    const int r = idx >> LOG2N, n = idx & (Ny - 1);
    const int row = 2 * (rp0 + r);
    if (row >= Nx) continue;
    const float2 z = sm[r * SP + n];
    dst[(long long)row * Ny + n] = z.x * scale;
    dst[(long long)(row + 1) * Ny + n] = z.y * scale;
  }
}

// columns per CTA: 8 (64-byte row segments) while the tile stays <= ~74 KB, fewer for the longest transforms
template <int LOG2N>
struct ColCfg {
  static constexpr int N = 1 << LOG2N;
  static constexpr int LOG2CT = LOG2N <= 10 ? 3 : (LOG2N == 11 ? 2 : 1);
  static constexpr int CT = 1 << LOG2CT;
  static constexpr int SC = CT + 1;
  static constexpr size_t smem = (size_t)N * SC * sizeof(float2);
};

template <int DIR, int LOG2N>
__global__ void __launch_bounds__(256) fft_cols_t(const float2* __restrict__ in, float2* __restrict__ out, int W,
                                                  const float2* __restrict__ tw) {
  using C = ColCfg<LOG2N>;
  constexpr int Nx = C::N, CT = C::CT, SC = C::SC;
  extern __shared__ __align__(16) float2 sm[];
  const long long img = blockIdx.y;
  const int c0 = blockIdx.x * CT;
  const float2* src = in + img * (long long)Nx * W;
  for (int idx = threadIdx.x; idx < Nx * CT; idx += blockDim.x) {
    const int n = idx >> C::LOG2CT, c = idx & (CT - 1);
    float2 v = make_float2(0.f, 0.f);
    if (c0 + c < W) v = __ldg(src + (long long)n * W + c0 + c);
    sm[digit_reverse_t<LOG2N>(n) * SC + c] = v;
  }
  fft_smem_t<DIR, LOG2N, true, C::LOG2CT, SC>(sm, tw);
  float2* dst = out + img * (long long)Nx * W;
  for (int idx = threadIdx.x; idx < Nx * CT; idx += blockDim.x) {
    const int n = idx >> C::LOG2CT, c = idx & (CT - 1);
    if (c0 + c < W) dst[(long long)n * W + c0 + c] = sm[n * SC + c];
  }
}


// ================================================================================================ Stockham kernels
// Same transforms with the autosort (Stockham) data flow: a radix-R pass reads element j + q*N/R (consecutive threads ->
// consecutive addresses), multiplies by W_{Ns*R}^{(j mod Ns) q}, and writes the butterfly outputs to
// expand(j) + q*Ns, ping-ponging between two shared-memory buffers; the result is in natural order, so there is no
// digit-reversal scatter (which serialised the staging stores 8-16x on bank conflicts).  The FIRST pass loads straight
// from global memory and -- for the column and the C2R row kernels -- the LAST pass stores straight to global memory:
// a 512-point transform makes 2 shared-memory exchanges instead of 4 staging/pass/readout round trips.
// Buffer index padding pad(i) = i + (i >> 3) makes the stride-R stores of the first passes conflict free.
// (8-byte elements: 16 lanes fill the 32 banks.  One pad element per 16 keeps a half-warp's run of consecutive elements
// inside one wavefront -- a pad per 8 elements, the first choice, split every such run over two -- and still spreads the
// stride-8 and stride-16 stores of the first pass: element 8 j sits at word 16 j + 2 (j >> 1), element 16 j at word 34 j.)
#ifndef AEFFT_FFT_PADSHIFT
#define AEFFT_FFT_PADSHIFT 4
#endif
__device__ __forceinline__ int padi(int i) { return i + (i >> AEFFT_FFT_PADSHIFT); }
// Radix schedule of the Stockham kernels: radix 8 passes after a first pass of radix 2 / 4 / 8 -- or 16 for N = 1024:
// three shared-memory-free register DFTs of 16 * 8 * 8 instead of four passes 2 * 8 * 8 * 8 (the extra pass cost
// a full shared-memory round trip and two barriers: the 1024-point row kernels ran at 0.6 of the 512-point ones per byte).
template <int LOG2N>
struct SSched {
  static constexpr int first = (LOG2N % 3 == 1) ? (LOG2N == 10 ? 4 : 1) : ((LOG2N % 3 == 2) ? 2 : 3);
  static constexpr int npass = 1 + (LOG2N - first) / 3;
  __host__ __device__ static constexpr int lr(int i) { return i == 0 ? first : 3; }
};


// one butterfly of pass I: loads + twiddles + DFT into a[], returns the output base index j0 (outputs go to j0 + q*NS)
template <int DIR, int LOG2N, int I, int LOG2NS, class Load>
__device__ __forceinline__ int stockham_load(int j, Load ld, const float2* __restrict__ tw, float2* a) {
  constexpr int LR = SSched<LOG2N>::lr(I), R = 1 << LR, NS = 1 << LOG2NS;
  const int k = j & (NS - 1);
#pragma unroll
  for (int q = 0; q < R; q++) a[q] = ld(j + (q << (LOG2N - LR)));
  if (LOG2NS > 0) {
#pragma unroll
    for (int q = 1; q < R; q++) {
      float2 w = __ldg(tw + ((k * q) << (LOG2N - LOG2NS - LR)));
      if (DIR > 0) w.y = -w.y;
      a[q] = cmul(a[q], w);
    }
  }
  if (R == 2) dft2<DIR>(a);
  else if (R == 4) dft4<DIR>(a);
  else if (R == 8) dft8<DIR>(a);
  else dft16<DIR>(a);
  return ((j - k) << LR) + k;
}
template <int DIR, int LOG2N, int I, int LOG2NS, class Load, class Store>
__device__ __forceinline__ void stockham_item(int j, Load ld, Store st, const float2* __restrict__ tw) {
  constexpr int LR = SSched<LOG2N>::lr(I), R = 1 << LR;
  float2 a[R];
  const int j0 = stockham_load<DIR, LOG2N, I, LOG2NS>(j, ld, tw, a);
#pragma unroll
  for (int q = 0; q < R; q++) st(j0 + (q << LOG2NS), a[q]);
}

template <int LOG2N>
struct SRowCfg {
  static constexpr int N = 1 << LOG2N;
  static constexpr int LOG2RP0 = LOG2N >= 11 ? 0 : (11 - LOG2N > 5 ? 5 : 11 - LOG2N);  // ~18 KB
  static constexpr int LOG2RPMIN = 8 + SSched<LOG2N>::first - LOG2N;  // >= 256 first-pass butterflies per CTA
  static constexpr int LOG2RP = LOG2RP0 >= LOG2RPMIN ? LOG2RP0 : (LOG2RPMIN > 5 ? 5 : LOG2RPMIN);
  static constexpr int RP = 1 << LOG2RP;
  static constexpr int SP = N + (N >> 3) + 2;
  static constexpr size_t smem = (size_t)RP * SP * sizeof(float2);
  static constexpr int npass = SSched<LOG2N>::npass;
};

// middle passes I = 1 .. npass-1 (shared -> shared) IN PLACE: every thread first loads and transforms all its butterflies
// (registers), the CTA synchronises, then the outputs overwrite the buffer -- one buffer instead of a ping-pong pair
// doubles the CTAs per SM.  256 threads per CTA.
template <int DIR, int LOG2N, int LOG2SEQ, int I, int LOG2NS, bool LAST_TO_CALLER, class Seq>
__device__ __forceinline__ void stockham_middle(float2* buf, const float2* __restrict__ tw, Seq sq) {
  constexpr int npass = SSched<LOG2N>::npass;
  if constexpr (I < npass - (LAST_TO_CALLER ? 1 : 0)) {
    constexpr int LR = SSched<LOG2N>::lr(I), R = 1 << LR;
    constexpr int TOTAL = 1 << (LOG2N - LR + LOG2SEQ);
    constexpr int ITEMS = TOTAL >= 256 ? TOTAL / 256 : 1;
    float2 a[ITEMS][R];
    int j0[ITEMS];
    __syncthreads();
#pragma unroll
    for (int t = 0; t < ITEMS; t++) {
      const int item = threadIdx.x + t * 256;
      if (item < TOTAL) {
        const int c = sq.seq(item, LOG2N - LR), j = sq.idx(item, LOG2N - LR);
        j0[t] = stockham_load<DIR, LOG2N, I, LOG2NS>(j, [&](int n) { return buf[sq.addr(c, n)]; }, tw, a[t]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < ITEMS; t++) {
      const int item = threadIdx.x + t * 256;
      if (item < TOTAL) {
        const int c = sq.seq(item, LOG2N - LR);
#pragma unroll
        for (int q = 0; q < R; q++) buf[sq.addr(c, j0[t] + (q << LOG2NS))] = a[t][q];
      }
    }
    stockham_middle<DIR, LOG2N, LOG2SEQ, I + 1, LOG2NS + LR, LAST_TO_CALLER>(buf, tw, sq);
  }
}

// addressing of `nseq` sequences inside a shared buffer
struct RowSeq {   // rows: sequence c contiguous with pitch SP; items: j fastest
  int SP;
  __device__ __forceinline__ int seq(int item, int log2nbf) const { return item >> log2nbf; }
  __device__ __forceinline__ int idx(int item, int log2nbf) const { return item & ((1 << log2nbf) - 1); }
  __device__ __forceinline__ int addr(int c, int n) const { return c * SP + padi(n); }
};
template <int LOG2CT>
struct ColSeq {   // columns: element n of column c at padi(n)*CT + c; items: c fastest
  __device__ __forceinline__ int seq(int item, int) const { return item & ((1 << LOG2CT) - 1); }
  __device__ __forceinline__ int idx(int item, int) const { return item >> LOG2CT; }
  __device__ __forceinline__ int addr(int c, int n) const { return (padi(n) << LOG2CT) + c; }
};

// PRUNE: the caller pools the spectrum right away (spectral pooling keeps the columns k < keep-1 and puts the Nyquist
// column Ny/2 on column keep-1, resize fft_backproplib.cu:87-157): only those `keep` columns are written, pitch `keep`.
template <int LOG2N, bool PRUNE = false>
__global__ void __launch_bounds__(256) fft_rows_r2c_s(const float* __restrict__ in, float2* __restrict__ out, int Nx,
                                                      const float2* __restrict__ tw, int ch, long long fstride, int keep = 0) {
  using C = SRowCfg<LOG2N>;
  constexpr int Ny = C::N, Nyr = Ny / 2 + 1, SP = C::SP, RP = C::RP, LR0 = SSched<LOG2N>::lr(0);
  extern __shared__ __align__(16) float2 sm[];
  float2* src = sm;
  const long long img = blockIdx.y;
  const int rp0 = blockIdx.x * RP;
  // real image `img` = (frame, channel): frames may be `fstride` floats apart (layers kept per frame by the caller)
  const float* base = in + (img / ch) * fstride + (img % ch) * (long long)Nx * Ny;
  RowSeq rs{SP};
  // pass 0: global (two real rows = one complex sequence) -> shared
  for (int item = threadIdx.x; item < (RP << (LOG2N - LR0)); item += blockDim.x) {
    const int r = item >> (LOG2N - LR0), j = item & ((1 << (LOG2N - LR0)) - 1);
    const int row = 2 * (rp0 + r);
    const bool ok = row < Nx;
    const float* ra = base + (long long)row * Ny;
    float2* d = src;
    stockham_item<-1, LOG2N, 0, 0>(
        j, [&](int n) { return ok ? make_float2(__ldg(ra + n), __ldg(ra + Ny + n)) : make_float2(0.f, 0.f); },
        [&](int i, float2 v) { d[rs.addr(r, i)] = v; }, tw);
  }
  stockham_middle<-1, LOG2N, C::LOG2RP, 1, LR0, false>(src, tw, rs);
  __syncthreads();
  // split Z = A + i B into the half spectra of the two rows.  The RP row pairs of the CTA are ONE index space over the
  // power-of-two part of the half spectrum (columns 0 .. ncol-1: shifts, no division -- the flat index over 2^k + 1 columns
  // cost ~25 integer instructions per element, more than the loads, the arithmetic and the stores together); the Nyquist
  // column Ny/2, which would be an extra one-thread trip per row, is left to the first RP threads.
  const int pitch = PRUNE ? keep : Nyr;
  const int ncol = PRUNE ? keep - 1 : Ny / 2;  // columns 0 .. ncol-1, then the Nyquist column at output column ncol
  float2* o = out + img * (long long)Nx * pitch;
  auto split = [&](const float2* sr, float2* o0, int k, int kk) {
    const float2 z1 = sr[padi(k)];
    const float2 z2 = sr[padi((Ny - k) & (Ny - 1))];
    o0[kk] = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
    o0[pitch + kk] = make_float2(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
  };
  {
    const bool p2 = (ncol & (ncol - 1)) == 0;
    const int sh = PRUNE ? 31 - __clz(ncol) : LOG2N - 1;
    for (int idx = threadIdx.x; idx < RP * ncol; idx += blockDim.x) {
      const int r = (!PRUNE || p2) ? idx >> sh : idx / ncol, k = idx - r * ncol;
      const int row = 2 * (rp0 + r);
      if (row >= Nx) break;
      split(src + r * SP, o + (long long)row * pitch, k, k);
    }
  }
  if (threadIdx.x < RP && 2 * (rp0 + (int)threadIdx.x) < Nx)
    split(src + threadIdx.x * SP, o + (long long)(2 * (rp0 + threadIdx.x)) * pitch, Ny / 2, ncol);
}

// EMBED: the input is a spectrum that was zero-embedded from a smaller one (spectral up-sampling): only `keep` columns
// exist (pitch `keep`): k < keep-1, and the Nyquist column Ny/2 at column keep-1; every other column reads as zero.
template <int LOG2N, bool EMBED = false>
__global__ void __launch_bounds__(256) fft_rows_c2r_s(const float2* __restrict__ in, float* __restrict__ out, int Nx,
                                                      const float2* __restrict__ tw, float scale, int ch, long long fstride,
                                                      int keep = 0) {
  using C = SRowCfg<LOG2N>;
  constexpr int Ny = C::N, Nyr = Ny / 2 + 1, SP = C::SP, RP = C::RP, LR0 = SSched<LOG2N>::lr(0), NP = C::npass;
  constexpr int LRL = SSched<LOG2N>::lr(NP - 1);
  extern __shared__ __align__(16) float2 sm[];
  float2* src = sm;
  const long long img = blockIdx.y;
  const int rp0 = blockIdx.x * RP;
  const int pitch = EMBED ? keep : Nyr;
  const float2* base = in + img * (long long)Nx * pitch;
  float* obase = out + (img / ch) * fstride + (img % ch) * (long long)Nx * Ny;
  RowSeq rs{SP};
  // Z[n] = A[n] + i B[n] for n <= N/2, conj(A[N-n]) + i conj(B[N-n]) above; imaginary parts of DC / Nyquist ignored
  auto load_z = [&](const float2* ra, bool ok, int n) {
    if (!ok) return make_float2(0.f, 0.f);
    const int m = n <= Ny / 2 ? n : Ny - n;
    int mc = m;
    if constexpr (EMBED) {
      if (m == Ny / 2) mc = keep - 1;
      else if (m >= keep - 1) return make_float2(0.f, 0.f);
    }
    float2 a = __ldg(ra + mc), b = __ldg(ra + pitch + mc);
    if (m == 0 || m == Ny / 2) { a.y = 0.f; b.y = 0.f; }
    return n <= Ny / 2 ? make_float2(a.x - b.y, a.y + b.x) : make_float2(a.x + b.y, b.x - a.y);
  };
  auto store_out = [&](int row, int i, float2 v) {
    obase[(long long)row * Ny + i] = v.x * scale;
    obase[(long long)(row + 1) * Ny + i] = v.y * scale;
  };
  if constexpr (NP == 1) {
    for (int item = threadIdx.x; item < (RP << (LOG2N - LR0)); item += blockDim.x) {
      const int r = item >> (LOG2N - LR0), j = item & ((1 << (LOG2N - LR0)) - 1);
      const int row = 2 * (rp0 + r);
      const bool ok = row < Nx;
      const float2* ra = base + (long long)row * pitch;
      stockham_item<+1, LOG2N, 0, 0>(j, [&](int n) { return load_z(ra, ok, n); },
                                     [&](int i, float2 v) { if (ok) store_out(row, i, v); }, tw);
    }
  } else {
    for (int item = threadIdx.x; item < (RP << (LOG2N - LR0)); item += blockDim.x) {
      const int r = item >> (LOG2N - LR0), j = item & ((1 << (LOG2N - LR0)) - 1);
      const int row = 2 * (rp0 + r);
      const bool ok = row < Nx;
      const float2* ra = base + (long long)row * pitch;
      float2* d = src;
      stockham_item<+1, LOG2N, 0, 0>(j, [&](int n) { return load_z(ra, ok, n); }, [&](int i, float2 v) { d[rs.addr(r, i)] = v; },
                                     tw);
    }
    stockham_middle<+1, LOG2N, C::LOG2RP, 1, LR0, true>(src, tw, rs);
    __syncthreads();
    // last pass: shared -> global (natural order: consecutive threads write consecutive pixels)
    for (int item = threadIdx.x; item < (RP << (LOG2N - LRL)); item += blockDim.x) {
      const int r = item >> (LOG2N - LRL), j = item & ((1 << (LOG2N - LRL)) - 1);
      const int row = 2 * (rp0 + r);
      const bool ok = row < Nx;
      const float2* s = src;
      stockham_item<+1, LOG2N, NP - 1, LOG2N - LRL>(j, [&](int n) { return s[rs.addr(r, n)]; },
                                                    [&](int i, float2 v) { if (ok) store_out(row, i, v); }, tw);
    }
  }
}

template <int LOG2N>
struct SColCfg {
  static constexpr int N = 1 << LOG2N;
  static constexpr int LOG2CT = LOG2N <= 9 ? 3 : (LOG2N == 10 ? 2 : (LOG2N == 11 ? 1 : 0));
  static constexpr int CT = 1 << LOG2CT;
  static constexpr size_t smem = (size_t)(N + (N >> 3) + 2) * CT * sizeof(float2);
};

// MODE 1 (forward + spectral pooling): only the Nxo rows the crop keeps are stored (i < Nxo/2 -> i, Nx/2 -> Nxo/2,
// i > Nx - Nxo/2 -> i - (Nx - Nxo)), into an [Nxo][W] image.  MODE 2 (inverse of a zero-embedded spectrum): the input image
// has Nxo rows, the rows in between read as zero.  MODE 0: plain.
template <int DIR, int LOG2N, int MODE = 0>
__global__ void __launch_bounds__(256) fft_cols_s(const float2* __restrict__ in, float2* __restrict__ out, int W,
                                                  const float2* __restrict__ tw, int Nxo = 0) {
  using C = SColCfg<LOG2N>;
  constexpr int Nx = C::N, CT = C::CT, LOG2CT = C::LOG2CT, LR0 = SSched<LOG2N>::lr(0), NP = SSched<LOG2N>::npass;
  constexpr int LRL = SSched<LOG2N>::lr(NP - 1);
  extern __shared__ __align__(16) float2 sm[];
  float2* src = sm;
  const long long img = blockIdx.y;
  const int c0 = blockIdx.x * CT;
  const float2* gin = in + img * (long long)(MODE == 2 ? Nxo : Nx) * W;
  float2* gout = out + img * (long long)(MODE == 1 ? Nxo : Nx) * W;
  // row maps of the spectral pooling (resize :87-157); -1 = dropped / zero
  auto out_row = [&](int i) {
    if constexpr (MODE != 1) return i;
    else return i < Nxo / 2 ? i : (i == Nx / 2 ? Nxo / 2 : (i > Nx - Nxo / 2 ? i - (Nx - Nxo) : -1));
  };
  auto in_row = [&](int n) {
    if constexpr (MODE != 2) return n;
    else return n < Nxo / 2 ? n : (n == Nx / 2 ? Nxo / 2 : (n > Nx - Nxo / 2 ? n - (Nx - Nxo) : -1));
  };
  auto gload = [&](int n, int c, bool ok) {
    const int r = in_row(n);
    return (ok && r >= 0) ? __ldg(gin + (long long)r * W + c0 + c) : make_float2(0.f, 0.f);
  };
  auto gstore = [&](int i, int c, bool ok, float2 v) {
    const int r = out_row(i);
    if (ok && r >= 0) gout[(long long)r * W + c0 + c] = v;
  };
  ColSeq<LOG2CT> cs;
  if constexpr (NP == 1) {
    for (int item = threadIdx.x; item < (1 << (LOG2N - LR0 + LOG2CT)); item += blockDim.x) {
      const int c = item & (CT - 1), j = item >> LOG2CT;
      const bool ok = c0 + c < W;
      stockham_item<DIR, LOG2N, 0, 0>(
          j, [&](int n) { return gload(n, c, ok); }, [&](int i, float2 v) { gstore(i, c, ok, v); }, tw);
    }
  } else {
    for (int item = threadIdx.x; item < (1 << (LOG2N - LR0 + LOG2CT)); item += blockDim.x) {
      const int c = item & (CT - 1), j = item >> LOG2CT;
      const bool ok = c0 + c < W;
      float2* d = src;
      stockham_item<DIR, LOG2N, 0, 0>(
          j, [&](int n) { return gload(n, c, ok); }, [&](int i, float2 v) { d[cs.addr(c, i)] = v; }, tw);
    }
    stockham_middle<DIR, LOG2N, LOG2CT, 1, LR0, true>(src, tw, cs);
    __syncthreads();
    for (int item = threadIdx.x; item < (1 << (LOG2N - LRL + LOG2CT)); item += blockDim.x) {
      const int c = item & (CT - 1), j = item >> LOG2CT;
      const bool ok = c0 + c < W;
      const float2* s = src;
      stockham_item<DIR, LOG2N, NP - 1, LOG2N - LRL>(j, [&](int n) { return s[cs.addr(c, n)]; },
                                                     [&](int i, float2 v) { gstore(i, c, ok, v); }, tw);
    }
  }
}

template <int LOG2N>
static int run_rows_r2c(aefft_ctx* ctx, int64_t batch, int Nx, const float* in, float2* out, const float2* tw, int ch,
                        long long fstride) {
  if (!getenv("AEFFT_FFT_V1")) {
    using S = SRowCfg<LOG2N>;
    AE_TRY(ctx->ensure_dyn_smem((const void*)fft_rows_r2c_s<LOG2N>, S::smem));
    dim3 grid_s((Nx / 2 + S::RP - 1) / S::RP, (unsigned)batch);
    fft_rows_r2c_s<LOG2N><<<grid_s, 256, S::smem, ctx->stream>>>(in, out, Nx, tw, ch, fstride);
    return AEFFT_OK;
  }
  if (fstride != (long long)ch * Nx * (1 << LOG2N)) return AEFFT_ERR_UNSUPPORTED;
  using C = RowCfg<LOG2N>;
  AE_TRY(ctx->ensure_dyn_smem((const void*)fft_rows_r2c_t<LOG2N>, C::smem));
  dim3 grid((Nx / 2 + C::RP - 1) / C::RP, (unsigned)batch);
  fft_rows_r2c_t<LOG2N><<<grid, 256, C::smem, ctx->stream>>>(in, out, Nx, tw);
  return AEFFT_OK;
}
template <int LOG2N>
static int run_rows_c2r(aefft_ctx* ctx, int64_t batch, int Nx, const float2* in, float* out, const float2* tw, float scale,
                        int ch, long long fstride) {
  if (!getenv("AEFFT_FFT_V1")) {
    using S = SRowCfg<LOG2N>;
    AE_TRY(ctx->ensure_dyn_smem((const void*)fft_rows_c2r_s<LOG2N>, S::smem));
    dim3 grid_s((Nx / 2 + S::RP - 1) / S::RP, (unsigned)batch);
    fft_rows_c2r_s<LOG2N><<<grid_s, 256, S::smem, ctx->stream>>>(in, out, Nx, tw, scale, ch, fstride);
    return AEFFT_OK;
  }
  if (fstride != (long long)ch * Nx * (1 << LOG2N)) return AEFFT_ERR_UNSUPPORTED;
  using C = RowCfg<LOG2N>;
  AE_TRY(ctx->ensure_dyn_smem((const void*)fft_rows_c2r_t<LOG2N>, C::smem));
  dim3 grid((Nx / 2 + C::RP - 1) / C::RP, (unsigned)batch);
  fft_rows_c2r_t<LOG2N><<<grid, 256, C::smem, ctx->stream>>>(in, out, Nx, tw, scale);
  return AEFFT_OK;
}
template <int DIR, int LOG2N>
static int run_cols(aefft_ctx* ctx, int64_t batch, int W, const float2* in, float2* out, const float2* tw) {
  if (!getenv("AEFFT_FFT_V1")) {
    // in place is fine: a CTA owns its CT columns, reads all of them in the first pass and writes them in the last
    using S = SColCfg<LOG2N>;
    AE_TRY(ctx->ensure_dyn_smem((const void*)fft_cols_s<DIR, LOG2N>, S::smem));
    dim3 grid_s((W + S::CT - 1) / S::CT, (unsigned)batch);
    fft_cols_s<DIR, LOG2N><<<grid_s, 256, S::smem, ctx->stream>>>(in, out, W, tw);
    return AEFFT_OK;
  }
  using C = ColCfg<LOG2N>;
  AE_TRY(ctx->ensure_dyn_smem((const void*)fft_cols_t<DIR, LOG2N>, C::smem));
  dim3 grid((W + C::CT - 1) / C::CT, (unsigned)batch);
  fft_cols_t<DIR, LOG2N><<<grid, 256, C::smem, ctx->stream>>>(in, out, W, tw);
  return AEFFT_OK;
}
#define AEFFT_FOR_LOG2N(X) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12)
// return AEFFT_ERR_UNSUPPORTED for lengths without a compile-time instantiation
static int rows_r2c_fast(aefft_ctx* ctx, int log2n, int64_t batch, int Nx, const float* in, float2* out, const float2* tw,
                         int ch, long long fstride) {
  switch (log2n) {
#define X(L) case L: return run_rows_r2c<L>(ctx, batch, Nx, in, out, tw, ch, fstride);
    AEFFT_FOR_LOG2N(X)
#undef X
  }
  return AEFFT_ERR_UNSUPPORTED;
}
static int rows_c2r_fast(aefft_ctx* ctx, int log2n, int64_t batch, int Nx, const float2* in, float* out, const float2* tw,
                         float scale, int ch, long long fstride) {
  switch (log2n) {
#define X(L) case L: return run_rows_c2r<L>(ctx, batch, Nx, in, out, tw, scale, ch, fstride);
    AEFFT_FOR_LOG2N(X)
#undef X
  }
  return AEFFT_ERR_UNSUPPORTED;
}
template <int DIR>
static int cols_fast(aefft_ctx* ctx, int log2n, int64_t batch, int W, const float2* in, float2* out, const float2* tw) {
  switch (log2n) {
#define X(L) case L: return run_cols<DIR, L>(ctx, batch, W, in, out, tw);
    AEFFT_FOR_LOG2N(X)
#undef X
  }
  return AEFFT_ERR_UNSUPPORTED;
}
static int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

// ------------------------------------------------------------------------------------------------ host side
static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
bool fft_len_supported(int N) { return fft_len_ok(N); }

int get_twiddles(aefft_ctx* ctx, int N, const float2** out) {
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
  return launch_fft_r2c_strided(ctx, batch, Nx, Ny, in, 1, (long long)Nx * Ny, spec);
}

// real images grouped per frame: image (frame f, channel c) at in + f*fstride + c*Nx*Ny, batch = frames*ch.
// AEFFT_ERR_UNSUPPORTED when the layout is not contiguous and the length has no strided-capable kernel.
int launch_fft_r2c_strided(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, const float* in, int ch, long long fstride,
                           float2* spec) {
  AE_ARG(batch > 0 && fft_len_ok(Nx) && fft_len_ok(Ny));
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
    const int rc = is_pow2(Ny) ? rows_r2c_fast(ctx, ilog2(Ny), batch, Nx, in, spec, twy, ch, fstride) : AEFFT_ERR_UNSUPPORTED;
    if (rc == AEFFT_ERR_UNSUPPORTED)
      fft_rows_r2c_kernel<<<grid, 256, smem, ctx->stream>>>(in, spec, Nx, Ny, RP, make_plan(Ny), twy, ch, fstride);
    else if (rc != AEFFT_OK) return rc;
  }
  {
    const int CT = col_tile(Nx);
    const size_t smem = (size_t)Nx * (CT | 1) * sizeof(float2);
    AE_TRY(set_smem(fft_cols_kernel<-1>, smem));
    dim3 grid((Nyr + CT - 1) / CT, (unsigned)batch);
    ProfScope prof(ctx, "fft_cols", 5.0 * sp * log2((double)Nx), 16.0 * sp);
    const int rc = is_pow2(Nx) ? cols_fast<-1>(ctx, ilog2(Nx), batch, Nyr, spec, spec, twx) : AEFFT_ERR_UNSUPPORTED;
    if (rc == AEFFT_ERR_UNSUPPORTED) fft_cols_kernel<-1><<<grid, 256, smem, ctx->stream>>>(spec, spec, Nx, Nyr, CT, make_plan(Nx), twx);
    else if (rc != AEFFT_OK) return rc;
  }
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// ---- transforms fused with the spectral pooling next to them (resize, fft_backproplib.cu:87-157) --------------------------
// R2C of [batch][Nx][Ny] real images straight into the POOLED half spectrum [batch][Nxs][Nys/2+1] (Nxs <= Nx, Nys <= Ny):
// the row pass writes only the Nys/2+1 columns the pooling keeps (tmp: batch*Nx*(Nys/2+1) complex), the column pass runs
// on those columns only and stores only the Nxs kept rows.  Half of the row-pass writes, half of the column pass and the
// separate resize kernel disappear; the full-resolution spectrum (which nothing else reads) is never formed.
template <int LY>
static int rows_r2c_pruned(aefft_ctx* ctx, int64_t batch, int Nx, const float* in, float2* tmp, const float2* tw, int keep) {
  using S = SRowCfg<LY>;
  AE_TRY(ctx->ensure_dyn_smem((const void*)fft_rows_r2c_s<LY, true>, S::smem));
  dim3 grid((Nx / 2 + S::RP - 1) / S::RP, (unsigned)batch);
  fft_rows_r2c_s<LY, true><<<grid, 256, S::smem, ctx->stream>>>(in, tmp, Nx, tw, 1, (long long)Nx * (1 << LY), keep);
  return AEFFT_OK;
}
template <int DIR, int MODE, int LX>
static int cols_mode(aefft_ctx* ctx, int64_t batch, int W, const float2* in, float2* out, const float2* tw, int Nxo) {
  using S = SColCfg<LX>;
  AE_TRY(ctx->ensure_dyn_smem((const void*)fft_cols_s<DIR, LX, MODE>, S::smem));
  dim3 grid((W + S::CT - 1) / S::CT, (unsigned)batch);
  fft_cols_s<DIR, LX, MODE><<<grid, 256, S::smem, ctx->stream>>>(in, out, W, tw, Nxo);
  return AEFFT_OK;
}
template <int LY>
static int rows_c2r_embed(aefft_ctx* ctx, int64_t batch, int Nx, const float2* tmp, float* out, const float2* tw, float scale,
                          int keep) {
  using S = SRowCfg<LY>;
  AE_TRY(ctx->ensure_dyn_smem((const void*)fft_rows_c2r_s<LY, true>, S::smem));
  dim3 grid((Nx / 2 + S::RP - 1) / S::RP, (unsigned)batch);
  fft_rows_c2r_s<LY, true><<<grid, 256, S::smem, ctx->stream>>>(tmp, out, Nx, tw, scale, 1, (long long)Nx * (1 << LY), keep);
  return AEFFT_OK;
}

int launch_fft_r2c_pooled(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, int Nxs, int Nys, const float* in, float2* tmp,
                          float2* out) {
  AE_ARG(batch > 0 && batch <= 65535);
  if (!(is_pow2(Nx) && is_pow2(Ny) && is_pow2(Nxs) && is_pow2(Nys))) return AEFFT_ERR_UNSUPPORTED;
  if (getenv("AEFFT_FFT_V1") || getenv("AEFFT_NO_FFT_POOL") || Nxs >= Nx || Nys >= Ny || Nxs < 4 || Nys < 4) return AEFFT_ERR_UNSUPPORTED;
  const int ly = ilog2(Ny), lx = ilog2(Nx), keep = Nys / 2 + 1;
  if (ly < 3 || ly > 12 || lx < 3 || lx > 12) return AEFFT_ERR_UNSUPPORTED;
  const float2 *twx, *twy;
  AE_TRY(get_twiddles(ctx, Nx, &twx));
  AE_TRY(get_twiddles(ctx, Ny, &twy));
  const double px = (double)batch * Nx * Ny, mid = (double)batch * Nx * keep, sp = (double)batch * Nxs * keep;
  {
    ProfScope prof(ctx, "fft_rows_r2c", 2.5 * px * log2((double)Ny), 4.0 * px + 8.0 * mid);
    int rc = AEFFT_ERR_UNSUPPORTED;
    switch (ly) {
#define X(L) case L: rc = rows_r2c_pruned<L>(ctx, batch, Nx, in, tmp, twy, keep); break;
      AEFFT_FOR_LOG2N(X)
#undef X
    }
    AE_TRY(rc);
  }
  {
    ProfScope prof(ctx, "fft_cols", 5.0 * mid * log2((double)Nx), 8.0 * (mid + sp));
    int rc = AEFFT_ERR_UNSUPPORTED;
    switch (lx) {
#define X(L) case L: rc = cols_mode<-1, 1, L>(ctx, batch, keep, tmp, out, twx, Nxs); break;
      AEFFT_FOR_LOG2N(X)
#undef X
    }
    AE_TRY(rc);
  }
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// C2R of the zero-EMBEDDED spectrum: spec is the small half spectrum [batch][Nxs][Nys/2+1]; the result is the inverse
// transform of its embedding into Nx x Ny (spectral up-sampling, no amplitude rescale), scaled by `scale`.
// tmp: batch*Nx*(Nys/2+1) complex.
int launch_fft_c2r_embedded(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, int Nxs, int Nys, const float2* spec, float2* tmp,
                            float* out, float scale) {
  AE_ARG(batch > 0 && batch <= 65535);
  if (!(is_pow2(Nx) && is_pow2(Ny) && is_pow2(Nxs) && is_pow2(Nys))) return AEFFT_ERR_UNSUPPORTED;
  if (getenv("AEFFT_FFT_V1") || getenv("AEFFT_NO_FFT_POOL") || Nxs >= Nx || Nys >= Ny || Nxs < 4 || Nys < 4) return AEFFT_ERR_UNSUPPORTED;
  const int ly = ilog2(Ny), lx = ilog2(Nx), keep = Nys / 2 + 1;
  if (ly < 3 || ly > 12 || lx < 3 || lx > 12) return AEFFT_ERR_UNSUPPORTED;
  const float2 *twx, *twy;
  AE_TRY(get_twiddles(ctx, Nx, &twx));
  AE_TRY(get_twiddles(ctx, Ny, &twy));
  const double px = (double)batch * Nx * Ny, mid = (double)batch * Nx * keep, sp = (double)batch * Nxs * keep;
  {
    ProfScope prof(ctx, "fft_cols", 5.0 * mid * log2((double)Nx), 8.0 * (mid + sp));
    int rc = AEFFT_ERR_UNSUPPORTED;
    switch (lx) {
#define X(L) case L: rc = cols_mode<+1, 2, L>(ctx, batch, keep, spec, tmp, twx, Nxs); break;
      AEFFT_FOR_LOG2N(X)
#undef X
    }
    AE_TRY(rc);
  }
  {
    ProfScope prof(ctx, "fft_rows_c2r", 2.5 * px * log2((double)Ny), 4.0 * px + 8.0 * mid);
    int rc = AEFFT_ERR_UNSUPPORTED;
    switch (ly) {
#define X(L) case L: rc = rows_c2r_embed<L>(ctx, batch, Nx, tmp, out, twy, scale, keep); break;
      AEFFT_FOR_LOG2N(X)
#undef X
    }
    AE_TRY(rc);
  }
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// spec is NOT modified: the column pass writes into `work` (batch*Nx*Nyr complex).
int launch_fft_c2r(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, const float2* spec, float2* work, float* out,
                   float scale) {
  return launch_fft_c2r_strided(ctx, batch, Nx, Ny, spec, work, out, 1, (long long)Nx * Ny, scale);
}

int launch_fft_c2r_strided(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, const float2* spec, float2* work, float* out, int ch,
                           long long fstride, float scale) {
  AE_ARG(batch > 0 && fft_len_ok(Nx) && fft_len_ok(Ny));
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
    const int rc = is_pow2(Nx) ? cols_fast<+1>(ctx, ilog2(Nx), batch, Nyr, spec, work, twx) : AEFFT_ERR_UNSUPPORTED;
    if (rc == AEFFT_ERR_UNSUPPORTED) fft_cols_kernel<+1><<<grid, 256, smem, ctx->stream>>>(spec, work, Nx, Nyr, CT, make_plan(Nx), twx);
    else if (rc != AEFFT_OK) return rc;
  }
  {
    const int RP = row_pairs(Nx, Ny);
    const size_t smem = (size_t)RP * (Ny + 1) * sizeof(float2);
    AE_TRY(set_smem(fft_rows_c2r_kernel, smem));
    dim3 grid((Nx / 2 + RP - 1) / RP, (unsigned)batch);
    ProfScope prof(ctx, "fft_rows_c2r", 2.5 * px * log2((double)Ny), 4.0 * px + 8.0 * sp);
    const int rc = is_pow2(Ny) ? rows_c2r_fast(ctx, ilog2(Ny), batch, Nx, work, out, twy, scale, ch, fstride) : AEFFT_ERR_UNSUPPORTED;
    if (rc == AEFFT_ERR_UNSUPPORTED)
      fft_rows_c2r_kernel<<<grid, 256, smem, ctx->stream>>>(work, out, Nx, Ny, RP, make_plan(Ny), twy, scale, ch, fstride);
    else if (rc != AEFFT_OK) return rc;
  }
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

}  // namespace aefft
