// Momentum-space per-bin channel contractions on the tensor cores (tcgen05, kind::tf32, 3xTF32 split).
//
// Replaces conv_k (fft_backproplib.cu:162-189) and the contractions inside gradient_k_io (:395-475) for layer pairs
// with >= 8 channels on both sides.  Per frequency bin w the reference computes small COMPLEX matrix products over the
// channels (forward, G) or over the frames (dC, dF -- this engine's batch extension); here every bin is one REAL GEMM
//     D_w [M x N] = A_w [M x K] * B_w [N x K]^T
// on interleaved (re, im) data: a complex row (x_0, x_1, ...) is the real row (Re x_0, Im x_0, Re x_1, ...), and a
// complex weight matrix W [r][c] is stored "embedded" as the real 2r x 2c matrix [[Wr, -Wi], [Wi, Wr]] so that
// interleaved in -> interleaved out with exactly 4 real multiplies per complex multiply (no wasted tensor work).
//
// Data layout: everything the tensor path touches is BIN-MAJOR, [bin][row][col] fp32 with `col` contiguous:
//   X~, E~ [S][B][2 dD]   H~, G~ [S][B][2 dM]   Cemb [S][2 dM][2 dD]   Femb [S][2 dD][2 dM]   dC, dF^T [S][dM][dD][2]
// so one bin's operand is one contiguous block and TMA lands it directly in the UMMA canonical layout:
//   K-major operand  (rows = M/N index, K contiguous): box {32 floats, rows}, SWIZZLE_128B   -> UMMA SWIZZLE_128B
//   MN-major operand (rows = K index, M/N contiguous): boxes {32 floats, 32 rows}, SWIZZLE_128B_ATOM_32B
//                                                      -> UMMA SWIZZLE_128B_BASE32B (the only transposed tf32 layout)
// (descriptor strides measured with tools/probe_tf32.cu).  The same Femb block serves O = H F^T as a K-major operand and
// G = E conj(F) as an MN-major one; the frame-reduced outer products read G~, H~, X~, E~ as MN-major operands (frames = K).
//
// fp32-grade results from tf32 products: x = hi + lo with hi = x & 0xffffe000 (exactly a tf32 number), lo = x - hi (exact in
// fp32); D += hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator: 1e-6 relative against fp64 (probe), the same budget as
// the BF16X3 split of the coordinate-space kernels.  The split is elementwise and position preserving: the splitter warps
// rewrite the TMA-landed tile in place (hi) and into a twin buffer (lo) without caring about the swizzle.
//
// Kernel: persistent, one CTA per SM, warp specialised: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue
// (TMEM -> registers -> global), warps 6-13 splitters.  Unit of the smem ring = one bin x one 32-wide K block
// (A 16 KB + B NT*128 B, twice for hi/lo); TMEM holds two accumulators so the epilogue of bin i overlaps the MMAs of i+1.
// Bound: HBM (per bin a 128 x 128 x 64 product is ~0.8 us of tensor time but 128+ KB of traffic).
// Measured (tools/tc_sweep.py, c3 shapes): 3.9-4.3 TB/s of algorithmic bytes for the forward / adjoint forms and the
// 64 -> 32 outer product, 2.6 TB/s for the small 32 -> 16 outer product.  Knock-outs (AEFFT_TC_KNOCK): without the global
// stores the store-heavy forms run at 7.9 TB/s-equivalent (the read stream alone is at the HBM read limit), without the
// split or the MMAs nothing changes (<10 %); a deeper ring (3 -> 5 stages), an 8-deep accumulator ring with two epilogue
// groups, 128-byte staged stores and 64-row K units for the outer products were all measured and gave nothing or were
// slower -- what remains is the mixed read/write stream itself (64 KB in, 64 KB out per bin and SM).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "pipe.cuh"
#include "tma.cuh"

namespace aefft {

namespace {

using namespace umma;

constexpr int TC_THREADS = 448;  // warp 0 TMA, 1 MMA, 2-5 epilogue, 6-13 splitters
constexpr int TC_SPLIT_WARPS = 8;
constexpr int TC_MAX_STAGES = 5;
enum { EPI_STORE = 0, EPI_OUTER = 1 };

struct TcParams {
  CUtensorMap amap, bmap;
  int a_mn, b_mn;        // 0: K-major, 1: MN-major
  int Mtot, Ntot;        // extents of D per bin
  int n_mt, n_nt, n_kb;  // tiles: M 128, N NT, K 32
  int NT;
  long long n_items;     // bins * n_mt * n_nt
  int stages;
  uint32_t stage_bytes;
  int epi;
  float scale;
  const float* bias;     // EPI_STORE: + bias[n / 2] * bias_scale on even columns n (real parts) of bin 0
  float bias_scale;
  const float* sub;      // EPI_STORE: - sub[bin][row][n]
  float* out;
  int conj_out;          // EPI_OUTER: negate the imaginary parts
  double* sq_part;       // EPI_STORE: per-warp sums of hw(bin) * out^2 ([grid][4]) or nullptr
  int ncols, col0, Ny;   // Hermitian weight of bin w: column col0 + w % ncols in {0, Ny/2} -> 1, else 2
  int knock;             // development knock-outs (AEFFT_TC_KNOCK bit mask): 1 no global stores, 2 no hi/lo split, 4 no MMAs
};

__device__ __forceinline__ uint64_t desc_k(uint32_t addr) {  // K-major, SWIZZLE_128B: SBO = 8 rows x 128 B
  return make_desc(addr, 16, 1024) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr) {  // MN-major, SWIZZLE_128B_BASE32B: 32-wide blocks 4096 B apart
  return make_desc(addr, 4096, 512) | ((uint64_t)1 << 61);    // (32 K rows x 128 B), 4-row K groups 512 B apart
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1) spec_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET from the shared array (a round trip through an integer makes the pointer generic and the
  // splitters' loads / stores become LD.E / ST.E: ncu showed a third of all stall samples on those loads)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[TC_MAX_STAGES], ready[TC_MAX_STAGES], empty[TC_MAX_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  uint32_t tcols = 32;
  while ((int)tcols < 2 * p.NT) tcols <<= 1;
  if (warp == 0) tmem_alloc(&tmem_slot, tcols);
  if (tid == 32) {
    for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&ready[s], TC_SPLIT_WARPS); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
    fence_mbar_init();
  }
  if (tid == 0) { tma::tma_prefetch_desc(&p.amap); tma::tma_prefetch_desc(&p.bmap); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_slot;
  const int per_bin = p.n_mt * p.n_nt;
  const uint32_t a_lo = 16384, b_hi = 32768, b_lo = 32768 + (uint32_t)p.NT * 128;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      Ring r(p.stages);
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const long long w = item / per_bin;
        const int rem = (int)(item - w * per_bin), mt = rem / p.n_nt, nt = rem - mt * p.n_nt;
        int a_blocks = (p.Mtot - mt * 128 + 31) / 32;
        if (a_blocks > 4) a_blocks = 4;
        const uint32_t bytes = (p.a_mn ? (uint32_t)a_blocks * 4096u : 16384u) + (uint32_t)p.NT * 128u;
        for (int kb = 0; kb < p.n_kb; kb++) {
          mbar_wait_role<false>(&empty[r.slot], r.phase ^ 1);
          uint8_t* st = smem + (size_t)r.slot * p.stage_bytes;
          tma::mbar_expect_tx(&full[r.slot], bytes);
          if (!p.a_mn) {
            tma::tma_load_3d(st, &p.amap, kb * 32, mt * 128, (int)w, &full[r.slot]);
          } else {
            for (int j = 0; j < a_blocks; j++) tma::tma_load_3d(st + j * 4096, &p.amap, mt * 128 + j * 32, kb * 32, (int)w, &full[r.slot]);
          }
          if (!p.b_mn) {
            tma::tma_load_3d(st + b_hi, &p.bmap, kb * 32, nt * p.NT, (int)w, &full[r.slot]);
          } else {
            for (int j = 0; j < p.NT / 32; j++)
              tma::tma_load_3d(st + b_hi + j * 4096, &p.bmap, nt * p.NT + j * 32, kb * 32, (int)w, &full[r.slot]);
          }
          r.next();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      Ring r(p.stages);
      const uint32_t idesc = idesc_tf32(128, p.NT, p.a_mn, p.b_mn);
      const uint32_t a_step = p.a_mn ? 1024u : 32u, b_step = p.b_mn ? 1024u : 32u;
      int it = 0;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x, it++) {
        const int a = it & 1;
        mbar_wait_role<false>(&acc_empty[a], ((it >> 1) & 1) ^ 1);
        fence_after_sync();
        const uint32_t d = tb + (uint32_t)(a * p.NT);
        uint32_t acc = 0;
        for (int kb = 0; kb < p.n_kb; kb++) {
          mbar_wait_role<false>(&ready[r.slot], r.phase);
          fence_after_sync();
          const uint32_t st = smem_u32(smem + (size_t)r.slot * p.stage_bytes);
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            const uint32_t ao = st + ks * a_step, bo = st + b_hi + ks * b_step;
            const uint64_t ah = p.a_mn ? desc_mn(ao) : desc_k(ao), al = p.a_mn ? desc_mn(ao + a_lo) : desc_k(ao + a_lo);
            const uint64_t bh = p.b_mn ? desc_mn(bo) : desc_k(bo);
            const uint64_t bl = p.b_mn ? desc_mn(bo + (b_lo - b_hi)) : desc_k(bo + (b_lo - b_hi));
            if (!(p.knock & 4)) {
              mma_tf32(d, ah, bh, idesc, acc);
              mma_tf32(d, ah, bl, idesc, 1);
              mma_tf32(d, al, bh, idesc, 1);
            }
            acc = 1;
          }
          commit(&empty[r.slot]);  // the stage may be refilled once these MMAs have read it
          r.next();
        }
        commit(&acc_full[a]);
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lanes 32q .. 32q+31 belong to this warp
    double sq = 0.0;
    int it = 0;
    for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x, it++) {
      const long long w = item / per_bin;
      const int rem = (int)(item - w * per_bin), mt = rem / p.n_nt, nt = rem - mt * p.n_nt;
      const int a = it & 1;
      mbar_wait_role<false>(&acc_full[a], (it >> 1) & 1);
      fence_after_sync();
      const int row = mt * 128 + q * 32 + lane;
      const uint32_t taddr = tb + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * p.NT);
      if (p.epi == EPI_STORE) {
        const bool row_ok = row < p.Mtot;
        const long long base = (w * p.Mtot + row) * (long long)p.Ntot + (long long)nt * p.NT;
        float hw = 2.f;
        if (p.sq_part) {
          const int wy = p.col0 + (int)(w % p.ncols);
          if (wy == 0 || wy == p.Ny / 2) hw = 1.f;
        }
        float part = 0.f;
        for (int c0 = 0; c0 < p.NT; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          const int n0 = nt * p.NT + c0;
          if (!row_ok || n0 >= p.Ntot) continue;
#pragma unroll
          for (int e = 0; e < 16; e++) v[e] *= p.scale;
          if (w == 0 && p.bias) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) v[e] = fmaf(__ldg(p.bias + (n0 + e) / 2), p.bias_scale, v[e]);
          }
          if (p.sub) {
            const float4* s4 = reinterpret_cast<const float4*>(p.sub + base + c0);
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const float4 s = __ldg(s4 + e);
              v[4 * e] -= s.x; v[4 * e + 1] -= s.y; v[4 * e + 2] -= s.z; v[4 * e + 3] -= s.w;
            }
          }
          if (p.sq_part) {
#pragma unroll
            for (int e = 0; e < 16; e++) part = fmaf(v[e], v[e], part);
          }
          if ((p.knock & 1) || !p.out) continue;  // out == nullptr: the caller wants the sum of squares only
          float4* o4 = reinterpret_cast<float4*>(p.out + base + c0);
#pragma unroll
          for (int e = 0; e < 4; e++) o4[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        }
        sq += (double)(part * hw);
      } else {
        // frame-reduced outer product of two interleaved complex operands: lane pair (2m, 2m+1) holds the four real sums
        // of row m; re = D[mr][dr] + D[mi][di], im = D[mi][dr] - D[mr][di]  (conj_out flips im)
        const int Mh = p.Mtot >> 1, Nh = p.Ntot >> 1;
        const int m = row >> 1;
        const bool odd = lane & 1;
        for (int c0 = 0; c0 < p.NT; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          const int n0 = nt * p.NT + c0;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const float pa = v[2 * i], pb = __shfl_xor_sync(0xffffffffu, v[2 * i + 1], 1);
            float r = odd ? pa - pb : pa + pb;
            if (odd && p.conj_out) r = -r;
            const int d = (n0 >> 1) + i;
            if (m < Mh && d < Nh && !(p.knock & 1)) p.out[((w * Mh + m) * (long long)Nh + d) * 2 + (odd ? 1 : 0)] = r * p.scale;
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(&acc_empty[a]);
    }
    if (p.sq_part) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_down_sync(0xffffffffu, sq, o);
      if (lane == 0) p.sq_part[(size_t)blockIdx.x * 4 + q] = sq;
    }
  } else {
    // ------------------------------------------------------------------------------------------ splitters
    const int t = tid - 6 * 32;  // 0..255
    Ring r(p.stages);
    const int nb4 = p.NT * 8;    // float4 per B panel
    for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int na4 = 1024;
      if (p.a_mn) {
        const long long w = item / per_bin;
        const int mt = (int)(item - w * per_bin) / p.n_nt;
        int a_blocks = (p.Mtot - mt * 128 + 31) / 32;
        na4 = (a_blocks > 4 ? 4 : a_blocks) * 256;
      }
      for (int kb = 0; kb < p.n_kb; kb++) {
        mbar_wait_role<false>(&full[r.slot], r.phase);
        uint8_t* st = smem + (size_t)r.slot * p.stage_bytes;
        uint4* ah = reinterpret_cast<uint4*>(st);
        float4* al = reinterpret_cast<float4*>(st + a_lo);
        uint4* bh = reinterpret_cast<uint4*>(st + b_hi);
        float4* bl = reinterpret_cast<float4*>(st + b_lo);
        auto split = [](uint4* hi, float4* lo, int i) {
          uint4 u = hi[i];
          const float4 x = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
          u.x &= 0xffffe000u; u.y &= 0xffffe000u; u.z &= 0xffffe000u; u.w &= 0xffffe000u;
          hi[i] = u;
          lo[i] = make_float4(x.x - __uint_as_float(u.x), x.y - __uint_as_float(u.y), x.z - __uint_as_float(u.z),
                              x.w - __uint_as_float(u.w));
        };
        if (!(p.knock & 2)) {
#pragma unroll 4
          for (int i = t; i < na4; i += 32 * TC_SPLIT_WARPS) split(ah, al, i);
#pragma unroll 4
          for (int i = t; i < nb4; i += 32 * TC_SPLIT_WARPS) split(bh, bl, i);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(&ready[r.slot]);
        r.next();
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, tcols);
}

__global__ void sq_final_kernel(const double* __restrict__ part, int n, double scale, float* __restrict__ out) {
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

// ---- [R][S] complex (bins fastest) -> [S][R] complex (bin-major), optionally minus a second source: 32 x 32 tiles
__global__ void __launch_bounds__(256) to_binmajor_kernel(const float2* __restrict__ in0, const float2* __restrict__ in1,
                                                          float2* __restrict__ out, long long R, long long S) {
  __shared__ float2 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long s0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const long long r = r0 + ty + 8 * k, s = s0 + tx;
    float2 v = make_float2(0.f, 0.f);
    if (r < R && s < S) {
      v = __ldg(in0 + r * S + s);
      if (in1) { const float2 u = __ldg(in1 + r * S + s); v.x -= u.x; v.y -= u.y; }
    }
    tile[ty + 8 * k][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const long long s = s0 + ty + 8 * k, r = r0 + tx;
    if (r < R && s < S) out[s * R + r] = tile[tx][ty + 8 * k];
  }
}

// ---- embedded kernel spectra, bin-major: emb[w][2r+a][2c+b] from the taps of the R x C kernels (pruned 25-tap DFT)
constexpr int KE_BINS = 8, KE_MAXT = 64;
__global__ void __launch_bounds__(256) kernel_spectrum_emb_kernel(const float* __restrict__ taps, float* __restrict__ emb, int R,
                                                                  int C, int Nk, int Nl, int Nx, int Ny, int ncols, int col0,
                                                                  long long S, const float2* __restrict__ twx,
                                                                  const float2* __restrict__ twy) {
  __shared__ float2 ph[KE_BINS][KE_MAXT];
  const int T = Nk * Nl;
  const long long w0 = (long long)blockIdx.x * KE_BINS;
  for (int i = threadIdx.x; i < KE_BINS * T; i += blockDim.x) {
    const int bi = i / T, t = i - bi * T, k = t / Nl, l = t - k * Nl;
    const long long w = w0 + bi;
    float2 v = make_float2(0.f, 0.f);
    if (w < S) {
      const int wx = (int)(w / ncols), wy = col0 + (int)(w - (long long)wx * ncols);
      const float2 ex = twx[tw_index(wx, k - Nk / 2, Nx)];
      const float2 ey = twy[tw_index(wy, l - Nl / 2, Ny)];
      v = make_float2(ex.x * ey.x - ex.y * ey.y, ex.x * ey.y + ex.y * ey.x);
    }
    ph[bi][t] = v;
  }
  __syncthreads();
  const int e = blockIdx.y * blockDim.x + threadIdx.x;
  if (e >= R * C) return;
  const int r = e / C, c = e - r * C;
  float tp[KE_MAXT];
#pragma unroll 1
  for (int t = 0; t < T; t++) tp[t] = taps[(size_t)e * T + t];
  for (int bi = 0; bi < KE_BINS; bi++) {
    const long long w = w0 + bi;
    if (w >= S) break;
    float vr = 0.f, vi = 0.f;
    for (int t = 0; t < T; t++) { vr = fmaf(tp[t], ph[bi][t].x, vr); vi = fmaf(tp[t], ph[bi][t].y, vi); }
    float* o = emb + ((w * 2 * R + 2 * r) * 2 * (long long)C + 2 * c);
    *reinterpret_cast<float2*>(o) = make_float2(vr, -vi);
    *reinterpret_cast<float2*>(o + 2 * C) = make_float2(vi, vr);
  }
}
// fixed-size tap registers (the generic kernel above spills for T = 25: tp[] is indexed in a runtime loop)
template <int T>
__global__ void __launch_bounds__(256) kernel_spectrum_emb_kernel_t(const float* __restrict__ taps, float* __restrict__ emb, int R,
                                                                    int C, int Nk, int Nl, int Nx, int Ny, int ncols, int col0,
                                                                    long long S, const float2* __restrict__ twx,
                                                                    const float2* __restrict__ twy) {
  __shared__ float2 ph[KE_BINS][T];
  const long long w0 = (long long)blockIdx.x * KE_BINS;
  for (int i = threadIdx.x; i < KE_BINS * T; i += blockDim.x) {
    const int bi = i / T, t = i - bi * T, k = t / Nl, l = t - k * Nl;
    const long long w = w0 + bi;
    float2 v = make_float2(0.f, 0.f);
    if (w < S) {
      const int wx = (int)(w / ncols), wy = col0 + (int)(w - (long long)wx * ncols);
      const float2 ex = twx[tw_index(wx, k - Nk / 2, Nx)];
      const float2 ey = twy[tw_index(wy, l - Nl / 2, Ny)];
      v = make_float2(ex.x * ey.x - ex.y * ey.y, ex.x * ey.y + ex.y * ey.x);
    }
    ph[bi][t] = v;
  }
  __syncthreads();
  const int e = blockIdx.y * blockDim.x + threadIdx.x;
  if (e >= R * C) return;
  const int r = e / C, c = e - r * C;
  float tp[T];
#pragma unroll
  for (int t = 0; t < T; t++) tp[t] = __ldg(taps + (size_t)e * T + t);
#pragma unroll 2
  for (int bi = 0; bi < KE_BINS; bi++) {
    const long long w = w0 + bi;
    if (w >= S) break;
    float vr = 0.f, vi = 0.f;
#pragma unroll
    for (int t = 0; t < T; t++) { vr = fmaf(tp[t], ph[bi][t].x, vr); vi = fmaf(tp[t], ph[bi][t].y, vi); }
    float* o = emb + ((w * 2 * R + 2 * r) * 2 * (long long)C + 2 * c);
    *reinterpret_cast<float2*>(o) = make_float2(vr, -vi);
    *reinterpret_cast<float2*>(o + 2 * C) = make_float2(vi, vr);
  }
}

// Separable form: spec(wx, wy) = sum_k Ex_k(wx) u_k(wy), u_k(wy) = sum_l taps[k][l] Ey_l(wy).  A thread owns one kernel
// (r, c), forms the NK column factors once per spectrum column and pays NK complex multiply-adds per bin (20 FMA for
// 5 taps instead of 50); a CTA covers KS_COLS columns x KS_ROWS rows.
constexpr int KS_COLS = 4, KS_ROWS = 32;
template <int NK, int NL>
// Nxm > 0: the spectrum is evaluated on the bins that a spectral pooling to Nxm rows x ncols columns keeps (resize :87-157:
// rows i < Nxm/2 -> i, Nxm/2 -> Nx/2, above -> i + Nx - Nxm; columns j < ncols-1 -> j, ncols-1 -> Ny/2), written on that grid.
__global__ void __launch_bounds__(256) kernel_spectrum_emb_sep_kernel(const float* __restrict__ taps, float* __restrict__ emb, int R,
                                                                      int C, int Nx, int Ny, int ncols, int col0,
                                                                      const float2* __restrict__ twx,
                                                                      const float2* __restrict__ twy, int Nxm) {
  __shared__ float2 ex[KS_ROWS][NK], ey[KS_COLS][NL];
  const int wl0 = blockIdx.x * KS_COLS, wx0 = blockIdx.z * KS_ROWS;
  const int rows = Nxm > 0 ? Nxm : Nx;
  for (int i = threadIdx.x; i < KS_ROWS * NK; i += blockDim.x) {
    const int rr = i / NK, k = i - rr * NK, wx = wx0 + rr;
    const int wxb = Nxm > 0 ? (wx < Nxm / 2 ? wx : (wx == Nxm / 2 ? Nx / 2 : wx + Nx - Nxm)) : wx;
    ex[rr][k] = wx < rows ? twx[tw_index(wxb, k - NK / 2, Nx)] : make_float2(0.f, 0.f);
  }
  for (int i = threadIdx.x; i < KS_COLS * NL; i += blockDim.x) {
    const int cc = i / NL, l = i - cc * NL, wy = col0 + wl0 + cc;
    const int wyb = Nxm > 0 ? (wy < ncols - 1 ? wy : Ny / 2) : wy;
    ey[cc][l] = wl0 + cc < ncols ? twy[tw_index(wyb, l - NL / 2, Ny)] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const int e = blockIdx.y * blockDim.x + threadIdx.x;
  if (e >= R * C) return;
  const int r = e / C, c = e - r * C;
  float tp[NK * NL];
#pragma unroll
  for (int t = 0; t < NK * NL; t++) tp[t] = __ldg(taps + (size_t)e * (NK * NL) + t);
  const int nrows = min(KS_ROWS, rows - wx0);
  for (int cc = 0; cc < KS_COLS && wl0 + cc < ncols; cc++) {
    float2 u[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) {
      float ur = 0.f, ui = 0.f;
#pragma unroll
      for (int l = 0; l < NL; l++) { ur = fmaf(tp[k * NL + l], ey[cc][l].x, ur); ui = fmaf(tp[k * NL + l], ey[cc][l].y, ui); }
      u[k] = make_float2(ur, ui);
    }
    for (int rr = 0; rr < nrows; rr++) {
      float vr = 0.f, vi = 0.f;
#pragma unroll
      for (int k = 0; k < NK; k++) {
        const float2 a = ex[rr][k];
        vr = fmaf(a.x, u[k].x, fmaf(-a.y, u[k].y, vr));
        vi = fmaf(a.x, u[k].y, fmaf(a.y, u[k].x, vi));
      }
      const long long w = (long long)(wx0 + rr) * ncols + wl0 + cc;
      float* o = emb + ((w * 2 * R + 2 * r) * 2 * (long long)C + 2 * c);
      *reinterpret_cast<float2*>(o) = make_float2(vr, -vi);
      *reinterpret_cast<float2*>(o + 2 * C) = make_float2(vi, vr);
    }
  }
}

// ---- kernel-space gradients from bin-major gradient spectra: part[(n*nsplit + sp)*T + t] = sum over the bins of split sp of
//      h(wy) Re( z[w][e] conj(Ex[k](wx) Ey[l](wy)) ),  n = e or its transpose (the dF^T block holds dF[d][m] at [m][d])
constexpr int BT_BINS = 16;
template <int T>
__global__ void __launch_bounds__(128) binmajor_to_taps_kernel(const float2* __restrict__ z, float* __restrict__ part, int E, int R,
                                                               int C, int transpose, int Nk, int Nl, int Nx, int Ny, int ncols,
                                                               int col0, long long S, const float2* __restrict__ twx,
                                                               const float2* __restrict__ twy) {
  __shared__ float2 ph[BT_BINS][T];
  const int nsplit = gridDim.y, sp = blockIdx.y;
  const long long w_lo = S * sp / nsplit, w_hi = S * (sp + 1) / nsplit;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  float g[T];
#pragma unroll
  for (int t = 0; t < T; t++) g[t] = 0.f;
  for (long long wb = w_lo; wb < w_hi; wb += BT_BINS) {
    __syncthreads();
    for (int i = threadIdx.x; i < BT_BINS * T; i += blockDim.x) {
      const int bi = i / T, t = i - bi * T, k = t / Nl, l = t - k * Nl;
      const long long w = wb + bi;
      float2 v = make_float2(0.f, 0.f);
      if (w < w_hi) {
        const int wx = (int)(w / ncols), wy = col0 + (int)(w - (long long)wx * ncols);
        const float h = (wy == 0 || wy == Ny / 2) ? 1.f : 2.f;
        const float2 ex = twx[tw_index(wx, k - Nk / 2, Nx)];
        const float2 ey = twy[tw_index(wy, l - Nl / 2, Ny)];
        v = make_float2(h * (ex.x * ey.x - ex.y * ey.y), h * (ex.x * ey.y + ex.y * ey.x));
      }
      ph[bi][t] = v;
    }
    __syncthreads();
    if (e < E) {
#pragma unroll 2
      for (int bi = 0; bi < BT_BINS; bi++) {
        const long long w = wb + bi;
        if (w >= w_hi) break;
        const float2 v = __ldg(z + w * E + e);
#pragma unroll
        for (int t = 0; t < T; t++) g[t] = fmaf(v.x, ph[bi][t].x, fmaf(v.y, ph[bi][t].y, g[t]));
      }
    }
  }
  if (e < E) {
    int n = e;
    if (transpose) { const int r = e / C, c = e - r * C; n = c * R + r; }
#pragma unroll
    for (int t = 0; t < T; t++) part[((size_t)n * nsplit + sp) * T + t] = g[t];
  }
}
// Separable form of the same reduction: Ey_l depends on the column only, so a thread first folds a whole spectrum ROW into
// NL complex partial sums t[l] = sum_wy h(wy) z conj(Ey_l(wy)) (NL complex multiply-adds per bin instead of NK*NL real
// pairs) and applies the NK row factors conj(Ex_k(wx)) once per row.  The h-weighted column factors of the slab sit in
// shared memory.  Splits are over rows.
constexpr int BT_ST = 16;  // bins per staging slot of the separable form (two slots per thread)
__device__ __forceinline__ void bt_cp_async8(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
// The spectrum values of a thread's kernel e are 8 bytes every E * 8 bytes: plain loads left the kernel bound by the memory
// latency (the compiler sinks them next to their uses: one or two in flight per thread, 1.5-2.1 TB/s).  Every thread therefore
// copies ITS next BT_ST bins into its own column of a shared-memory slot with cp.async while it works on the previous slot --
// 16 to 32 loads in flight per thread, no barrier (a thread only reads what it copied itself).  The sums keep their order.
template <int NK, int NL>
__global__ void __launch_bounds__(128) binmajor_to_taps_sep_kernel(const float2* __restrict__ z, float* __restrict__ part, int E,
                                                                   int R, int C, int transpose, int Nx, int Ny, int ncols, int col0,
                                                                   const float2* __restrict__ twx,
                                                                   const float2* __restrict__ twy) {
  extern __shared__ __align__(16) float2 phy[];  // [ncols][NL], then the staging slots [2][BT_ST][128]
  float2* ring = phy + (((size_t)ncols * NL + 1) & ~(size_t)1);
  for (int i = threadIdx.x; i < ncols * NL; i += blockDim.x) {
    const int wl = i / NL, l = i - wl * NL, wy = col0 + wl;
    const float h = (wy == 0 || wy == Ny / 2) ? 1.f : 2.f;
    const float2 ey = twy[tw_index(wy, l - NL / 2, Ny)];
    phy[i] = make_float2(h * ey.x, h * ey.y);
  }
  __syncthreads();
  const int nsplit = gridDim.y, sp = blockIdx.y;
  const int r_lo = (int)((long long)Nx * sp / nsplit), r_hi = (int)((long long)Nx * (sp + 1) / nsplit);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float g[NK * NL];
#pragma unroll
  for (int t = 0; t < NK * NL; t++) g[t] = 0.f;
  const long long nq = (long long)(r_hi - r_lo) * ncols;      // bins of this CTA, row after row
  const float2* zb = z + (long long)r_lo * ncols * E + e;
  float2* mine = ring + threadIdx.x;
  auto issue = [&](int slot, long long q0) {
#pragma unroll
    for (int u = 0; u < BT_ST; u++)
      if (q0 + u < nq) bt_cp_async8(mine + (slot * BT_ST + u) * 128, zb + (q0 + u) * E);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float2 t[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) t[l] = make_float2(0.f, 0.f);
  int wl = 0, wx = r_lo, slot = 0;
  if (nq > 0) issue(0, 0);
  for (long long q0 = 0; q0 < nq; q0 += BT_ST, slot ^= 1) {
    issue(slot ^ 1, q0 + BT_ST);  // (an empty group past the end keeps the group count uniform)
    asm volatile("cp.async.wait_group 1;" ::: "memory");
#pragma unroll 4
    for (int u = 0; u < BT_ST; u++) {
      if (q0 + u >= nq) break;
      const float2 v = mine[(slot * BT_ST + u) * 128];
#pragma unroll
      for (int l = 0; l < NL; l++) {  // t[l] += v * conj(phy)
        const float2 ph = phy[wl * NL + l];
        t[l].x = fmaf(v.x, ph.x, fmaf(v.y, ph.y, t[l].x));
        t[l].y = fmaf(v.y, ph.x, fmaf(-v.x, ph.y, t[l].y));
      }
      if (++wl == ncols) {  // end of a spectrum row: the NK row factors conj(Ex_k(wx)), once per row
#pragma unroll
        for (int k = 0; k < NK; k++) {
          const float2 ex = __ldg(twx + (tw_index(wx, k - NK / 2, Nx)));
#pragma unroll
          for (int l = 0; l < NL; l++) g[k * NL + l] = fmaf(t[l].x, ex.x, fmaf(t[l].y, ex.y, g[k * NL + l]));  // Re(t conj(ex))
        }
#pragma unroll
        for (int l = 0; l < NL; l++) t[l] = make_float2(0.f, 0.f);
        wl = 0; wx++;
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  int n = e;
  if (transpose) { const int r = e / C, c = e - r * C; n = c * R + r; }
#pragma unroll
  for (int t2 = 0; t2 < NK * NL; t2++) part[((size_t)n * nsplit + sp) * (NK * NL) + t2] = g[t2];
}

__global__ void taps_final_kernel(const float* __restrict__ part, float* __restrict__ taps, long long total, int nsplit, int T,
                                  float scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long n = idx / T;
  const int t = (int)(idx - n * T);
  double s = 0.0;
  for (int sp = 0; sp < nsplit; sp++) s += (double)part[(n * nsplit + sp) * T + t];
  taps[idx] = (float)(s * (double)scale);
}

// ---- DC-bin terms of gradient_k_io (:447-473) on bin-major data (bin 0 = the first block):
//   db[m] = gs * sum_b Re G[b][m](0);  dp[d] = gs * sum_b Re E[b][d](0);
//   dF^T[0][m][d] += fs * corr[m] * sum_b E[b][d](0)   (H-hat's bias correction, quirk F1; corr real)
__global__ void __launch_bounds__(512) dc_terms_kernel(const float* __restrict__ G, const float* __restrict__ E,
                                                       const float* __restrict__ bias_b, float* __restrict__ dFt, float* __restrict__ db,
                                                       float* __restrict__ dp, int B, int dM, int dD, float gs, float fs,
                                                       float corr_scale) {
  // ONE CTA, two phases: the frame sums (a thread per channel), then the dM x dD corrections of dF^T spread over all threads
  // (a thread per d walking the dM read-modify-writes one after the other took 90 us at 128 -> 256 channels)
  extern __shared__ double dct_es[];  // [dD][2]: sum_b E[b][d](0)
  for (int n = threadIdx.x; n < dM + dD; n += blockDim.x) {
    if (n < dM) {
      double s = 0.0;
      for (int b = 0; b < B; b++) s += (double)G[(size_t)b * 2 * dM + 2 * n];
      db[n] = (float)(s * (double)gs);
    } else {
      const int d = n - dM;
      double sr = 0.0, si = 0.0;
      for (int b = 0; b < B; b++) { sr += (double)E[(size_t)b * 2 * dD + 2 * d]; si += (double)E[(size_t)b * 2 * dD + 2 * d + 1]; }
      dp[d] = (float)(sr * (double)gs);
      dct_es[2 * d] = sr; dct_es[2 * d + 1] = si;
    }
  }
  if (!bias_b) return;
  __syncthreads();
  for (int i = threadIdx.x; i < dM * dD; i += blockDim.x) {
    const int m = i / dD, d = i - m * dD;
    const double corr = (double)bias_b[m] * (double)corr_scale * (double)fs;
    dFt[(size_t)i * 2] += (float)(corr * dct_es[2 * d]);
    dFt[(size_t)i * 2 + 1] += (float)(corr * dct_es[2 * d + 1]);
  }
}

// ---- Both gradient spectra of an iteration from ONE frame-reduced outer product (associativity):
//   Mg[d][d'] = sum_b E[b][d] conj(X[b][d'])                                   (dD x dD per bin)
//   dC[m][d]  = gs * sum_k conj(F[k][m]) Mg[k][d]             = gs * sum_b G[b][m] conj(X[b][d]),  G = E conj(F)  (:437-445)
//   dF[d][m]  = gs * sum_k conj(C[m][k]) Mg[d][k]             = gs * sum_b E[b][d] conj(H-hat[b][m]) away from DC  (:447-459)
// so neither G nor the hidden spectrum is read (or written) for the gradients: the per-iteration traffic of a pair drops from
// adjoint + two outer products over [frames][dM] operands to ONE pass over the [frames][dD] operands E and X plus
// kernel-spectrum-sized data.  One CTA per bin, CUDA cores: phase 1 stages E and X of the bin in shared memory and forms Mg
// with TR x TR register tiles; phase 2 re-uses that shared memory for C and F (even rows of the embedded blocks: row 2r =
// (Re, -Im) of W[r][:], stored [k][m]) and gives every thread NO adjacent m of one d.  Outputs [bin][m][d] like the
// outer-product epilogue writes them.  (A first version ran Mg on the tensor cores -- 128-row MMA tiles with 32 or 64 live
// rows and one pipeline round trip per bin: 1.97 TB/s -- and phase 2 without register tiles: 0.9 + 0.7 ms at config 3.)
template <int DD, int NO>
__global__ void __launch_bounds__(256) gram_grad_kernel(const float* __restrict__ E, const float* __restrict__ X,
                                                        const float* __restrict__ Cemb, const float* __restrict__ Femb,
                                                        float2* __restrict__ dCt, float2* __restrict__ dFt, int B, int dM, float gs) {
  // phase 1 tiling: (DD / TR)^2 register tiles of TR x TR outputs; with fewer than 256 tiles the frames are split over
  // NG = 256 / tiles thread groups whose partial sums are added in group order (deterministic)
  constexpr int TR = DD >= 32 ? 4 : (DD >= 16 ? 2 : 1), TG = DD / TR, NT = TG * TG, NG = 256 / NT, MP = DD + 1;
  static_assert(NT * NG == 256, "256 threads");
  extern __shared__ __align__(16) float2 gg_sm[];
  float2* Ms = gg_sm;                // [DD][DD + 1]
  float2* Es = Ms + DD * MP;         // [B][DD]   (DD is even: DD * MP float2 keep the 16-byte alignment)
  float2* Xs = Es + (size_t)B * DD;  // [B][DD]
  float2* Fs = Es;                   // phase 2: [DD][dM]
  float2* Cs = Es + (size_t)DD * dM; // phase 2: [DD][dM]  (C transposed: Cs[k][m] = C[m][k])
  const long long w = blockIdx.x;
  const int tid = threadIdx.x;
  {
    const float4* e4 = reinterpret_cast<const float4*>(E + w * (long long)B * 2 * DD);
    const float4* x4 = reinterpret_cast<const float4*>(X + w * (long long)B * 2 * DD);
    float4* es4 = reinterpret_cast<float4*>(Es);
    float4* xs4 = reinterpret_cast<float4*>(Xs);
    for (int i = tid; i < B * DD / 2; i += 256) { es4[i] = __ldg(e4 + i); xs4[i] = __ldg(x4 + i); }
  }
  __syncthreads();
  {
    const int tile = tid % NT, grp = tid / NT;
    const int dr = (tile / TG) * TR, dc = (tile % TG) * TR;
    float2 acc[TR][TR];
#pragma unroll
    for (int a = 0; a < TR; a++)
#pragma unroll
      for (int c = 0; c < TR; c++) acc[a][c] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int b = grp; b < B; b += NG) {
      float2 e[TR], x[TR];
#pragma unroll
      for (int a = 0; a < TR; a++) { e[a] = Es[b * DD + dr + a]; x[a] = Xs[b * DD + dc + a]; }
#pragma unroll
      for (int a = 0; a < TR; a++)
#pragma unroll
        for (int c = 0; c < TR; c++) {  // e * conj(x)
          acc[a][c].x = fmaf(e[a].x, x[c].x, acc[a][c].x); acc[a][c].x = fmaf(e[a].y, x[c].y, acc[a][c].x);
          acc[a][c].y = fmaf(e[a].y, x[c].x, acc[a][c].y); acc[a][c].y = fmaf(-e[a].x, x[c].y, acc[a][c].y);
        }
    }
    if constexpr (NG == 1) {
#pragma unroll
      for (int a = 0; a < TR; a++)
#pragma unroll
        for (int c = 0; c < TR; c++) Ms[(dr + a) * MP + dc + c] = acc[a][c];
    } else {
      __syncthreads();  // every group is done with E and X: their space takes the partial sums [group][DD][MP]
      float2* Pp = Es + (size_t)grp * DD * MP;
#pragma unroll
      for (int a = 0; a < TR; a++)
#pragma unroll
        for (int c = 0; c < TR; c++) Pp[(dr + a) * MP + dc + c] = acc[a][c];
      __syncthreads();
      for (int i = tid; i < DD * MP; i += 256) {
        float2 sum = Es[i];
#pragma unroll
        for (int g = 1; g < NG; g++) { const float2 v = Es[(size_t)g * DD * MP + i]; sum.x += v.x; sum.y += v.y; }
        Ms[i] = sum;
      }
    }
  }
  __syncthreads();  // Mg complete; E and X are dead: their space takes F and C
  for (int i = tid; i < dM * DD; i += 256) {
    const int k = i / dM, m = i - k * dM;   // F[k][m]: embedded row 2k, columns 2m, 2m+1
    const float2 u = __ldg(reinterpret_cast<const float2*>(Femb + ((w * 2 * DD + 2 * k) * 2 * (long long)dM + 2 * m)));
    Fs[i] = make_float2(u.x, -u.y);
    const int mc = i / DD, kc = i - mc * DD;  // C[mc][kc]: embedded row 2 mc, columns 2 kc, 2 kc + 1
    const float2 v = __ldg(reinterpret_cast<const float2*>(Cemb + ((w * 2 * dM + 2 * mc) * 2 * (long long)DD + 2 * kc)));
    Cs[kc * dM + mc] = make_float2(v.x, -v.y);
  }
  __syncthreads();
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
          const float4 cv = *reinterpret_cast<const float4*>(Cs + k * dM + m0 + j);
          f[j] = make_float2(fv.x, fv.y); f[j + 1] = make_float2(fv.z, fv.w);
          c[j] = make_float2(cv.x, cv.y); c[j + 1] = make_float2(cv.z, cv.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NO; j++) { f[j] = Fs[k * dM + m0 + j]; c[j] = Cs[k * dM + m0 + j]; }
      }
#pragma unroll
      for (int j = 0; j < NO; j++) {  // conj(F[k][m]) * Mg[k][d]  and  conj(C[m][k]) * Mg[d][k]
        aC[j].x = fmaf(f[j].x, m1.x, aC[j].x); aC[j].x = fmaf(f[j].y, m1.y, aC[j].x);
        aC[j].y = fmaf(f[j].x, m1.y, aC[j].y); aC[j].y = fmaf(-f[j].y, m1.x, aC[j].y);
        aF[j].x = fmaf(c[j].x, m2.x, aF[j].x); aF[j].x = fmaf(c[j].y, m2.y, aF[j].x);
        aF[j].y = fmaf(c[j].x, m2.y, aF[j].y); aF[j].y = fmaf(-c[j].y, m2.x, aF[j].y);
      }
    }
#pragma unroll
    for (int j = 0; j < NO; j++) {
      const long long o = (w * dM + m0 + j) * DD + d;
      dCt[o] = make_float2(aC[j].x * gs, aC[j].y * gs);
      dFt[o] = make_float2(aF[j].x * gs, aF[j].y * gs);
    }
  }
}

// DC-bin terms of gradient_k_io (:447-473) for the Gram form (bin 0 = the first block of the bin-major arrays), from
// Esum[d] = sum_b E[b][d](0):   dp[d] = gs Re Esum[d];   db[m] = gs Re sum_d conj(F[d][m](0)) Esum[d]  (= gs sum_b Re G[b][m](0));
//   dF^T[0][m][d] += fs * bias_b[m] * norm * Esum[d]   (the bias part of H-hat = sum_d' C X + b Nx Ny at DC, quirk F1)
__global__ void dc_terms_gram_kernel(const float* __restrict__ E, const float* __restrict__ Femb, const float* __restrict__ bias_b,
                                     float* __restrict__ dFt, float* __restrict__ db, float* __restrict__ dp, int B, int dM, int dD,
                                     float gs, float fs, float norm) {
  extern __shared__ double dcg_sm[];  // Esum re / im [dD]
  for (int d = threadIdx.x; d < dD; d += blockDim.x) {
    double sr = 0.0, si = 0.0;
    for (int b = 0; b < B; b++) { sr += (double)E[(size_t)b * 2 * dD + 2 * d]; si += (double)E[(size_t)b * 2 * dD + 2 * d + 1]; }
    dcg_sm[2 * d] = sr; dcg_sm[2 * d + 1] = si;
    dp[d] = (float)(sr * (double)gs);
  }
  __syncthreads();
  for (int m = threadIdx.x; m < dM; m += blockDim.x) {
    double s = 0.0;
    for (int d = 0; d < dD; d++) {
      // F[d][m](0) = (emb[2d][2m], -emb[2d][2m+1]);  Re(conj(F) Esum) = Fr Er + Fi Ei
      const double fr = (double)Femb[((size_t)2 * d) * 2 * dM + 2 * m], fi = -(double)Femb[((size_t)2 * d) * 2 * dM + 2 * m + 1];
      s += fr * dcg_sm[2 * d] + fi * dcg_sm[2 * d + 1];
      if (bias_b) {
        const double corr = (double)bias_b[m] * (double)norm * (double)fs;
        dFt[((size_t)m * dD + d) * 2] += (float)(corr * dcg_sm[2 * d]);
        dFt[((size_t)m * dD + d) * 2 + 1] += (float)(corr * dcg_sm[2 * d + 1]);
      }
    }
    db[m] = (float)(s * (double)gs);
  }
}

struct TcOperand {
  const float* base;
  int rows, cols;  // memory [S][rows][cols]
  int mn;          // 0: rows = M/N index, cols = K (K-major); 1: rows = K, cols = M/N index (MN-major)
};

int launch_bgemm(aefft_ctx* ctx, const char* name, long long S, const TcOperand& A, const TcOperand& B, int Mtot, int Ntot,
                 int Ktot, int epi, float scale, const float* bias, float bias_scale, const float* sub, float* out, int conj_out,
                 float* sq_out, double sq_scale, int ncols, int col0, int Ny, double alg_bytes) {
  AE_ARG(S > 0 && S < (1LL << 31) && Mtot > 0 && Ntot > 0 && Ktot > 0 && Ntot % 16 == 0 && (epi != EPI_OUTER || Mtot % 2 == 0));
  AE_ARG((A.mn ? A.cols : A.rows) == Mtot && (A.mn ? A.rows : A.cols) == Ktot);
  AE_ARG((B.mn ? B.cols : B.rows) == Ntot && (B.mn ? B.rows : B.cols) == Ktot);
  AE_ARG(A.cols % 4 == 0 && B.cols % 4 == 0);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.a_mn = A.mn; p.b_mn = B.mn; p.Mtot = Mtot; p.Ntot = Ntot;
  int NT = Ntot <= 256 ? Ntot : 256;
  if (B.mn) NT = (NT + 31) / 32 * 32;
  p.NT = NT;
  p.n_mt = (Mtot + 127) / 128; p.n_nt = (Ntot + NT - 1) / NT; p.n_kb = (Ktot + 31) / 32;
  p.n_items = S * p.n_mt * p.n_nt;
  p.stage_bytes = 32768u + 2u * (uint32_t)NT * 128u;
  p.stages = (int)((227u * 1024u - 2048u) / p.stage_bytes);
  if (p.stages > TC_MAX_STAGES) p.stages = TC_MAX_STAGES;
  if (const char* e = getenv("AEFFT_TC_STAGES")) {  // development knobs: cap the ring depth, knock out a role's work
    const int cap = atoi(e);
    if (cap >= 2 && cap < p.stages) p.stages = cap;
  }
  if (const char* e = getenv("AEFFT_TC_KNOCK")) p.knock = atoi(e);
  AE_ARG(p.stages >= 2);
  p.epi = epi; p.scale = scale; p.bias = bias; p.bias_scale = bias_scale; p.sub = sub; p.out = out; p.conj_out = conj_out;
  p.ncols = ncols > 0 ? ncols : 1; p.col0 = col0; p.Ny = Ny;
  int rc = A.mn ? tma::make_tmap_3d_f32(&p.amap, A.base, A.cols, A.rows, S, 32, 32, 1, 2)
                : tma::make_tmap_3d_f32(&p.amap, A.base, A.cols, A.rows, S, 32, 128, 1, 1);
  if (rc == 0)
    rc = B.mn ? tma::make_tmap_3d_f32(&p.bmap, B.base, B.cols, B.rows, S, 32, 32, 1, 2)
              : tma::make_tmap_3d_f32(&p.bmap, B.base, B.cols, B.rows, S, 32, NT, 1, 1);
  if (rc != 0) { set_error("spec_tc: cuTensorMapEncodeTiled failed (%d)", rc); return AEFFT_ERR_CUDA; }
  const long long grid = p.n_items < ctx->sm_count ? p.n_items : ctx->sm_count;
  double* part = nullptr;
  if (sq_out) {
    AE_TRY(ctx->getT("tc_sq_part", (size_t)grid * 4, &part));
    p.sq_part = part;
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  AE_TRY(ctx->ensure_dyn_smem((const void*)spec_tc_kernel, smem));
  {
    ProfScope prof(ctx, name, 2.0 * (double)S * Mtot * Ntot * Ktot, alg_bytes);
    spec_tc_kernel<<<(unsigned)grid, TC_THREADS, smem, ctx->stream>>>(p);
    ctx->launches++;
  }
  AE_CUDA(cudaGetLastError());
  if (sq_out) {
    sq_final_kernel<<<1, 256, 0, ctx->stream>>>(part, (int)grid * 4, sq_scale, sq_out);
    ctx->launches++;
    AE_CUDA(cudaGetLastError());
  }
  return AEFFT_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ public (engine-internal)

bool spec_tc_eligible(int dD, int dM, int Nk, int Nl) {
  return dD % 8 == 0 && dM % 8 == 0 && dD >= 8 && dM >= 8 && Nk * Nl <= KE_MAXT && !getenv("AEFFT_NO_SPEC_TC");
}

int launch_to_binmajor(aefft_ctx* ctx, long long R, long long S, const float2* in0, const float2* in1, float2* out) {
  dim3 grid((unsigned)((S + 31) / 32), (unsigned)((R + 31) / 32));
  AE_ARG(grid.y <= 65535);
  ProfScope prof(ctx, "spec_to_binmajor", 0.0, 8.0 * R * S * (in1 ? 3 : 2));
  to_binmajor_kernel<<<grid, 256, 0, ctx->stream>>>(in0, in1, out, R, S);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// the embedded spectra at resolution (Nx, Ny) on the bins a spectral pooling to (Nxm, Nym) keeps: [Nxm * (Nym/2+1)][2R][2C]
int launch_kernel_spectrum_emb_pooled(aefft_ctx* ctx, int R, int C, int Nk, int Nl, int Nx, int Ny, int Nxm, int Nym,
                                      const float* taps, float* emb) {
  const float2 *twx, *twy;
  AE_TRY(get_twiddles(ctx, Nx, &twx));
  AE_TRY(get_twiddles(ctx, Ny, &twy));
  if (!(Nk == Nl && (Nk == 5 || Nk == 3 || Nk == 7)) || Nxm >= Nx || Nym >= Ny || Nxm < 2 || Nym < 2) return AEFFT_ERR_UNSUPPORTED;
  const int ncols = Nym / 2 + 1;
  const long long S = (long long)Nxm * ncols;
  ProfScope prof(ctx, "kernel_spectrum_emb", 4.0 * S * R * C * 2.0 * Nk, 16.0 * S * R * C);
  dim3 g2((unsigned)((ncols + KS_COLS - 1) / KS_COLS), (unsigned)((R * C + 255) / 256), (unsigned)((Nxm + KS_ROWS - 1) / KS_ROWS));
  if (Nk == 5) kernel_spectrum_emb_sep_kernel<5, 5><<<g2, 256, 0, ctx->stream>>>(taps, emb, R, C, Nx, Ny, ncols, 0, twx, twy, Nxm);
  else if (Nk == 3) kernel_spectrum_emb_sep_kernel<3, 3><<<g2, 256, 0, ctx->stream>>>(taps, emb, R, C, Nx, Ny, ncols, 0, twx, twy, Nxm);
  else kernel_spectrum_emb_sep_kernel<7, 7><<<g2, 256, 0, ctx->stream>>>(taps, emb, R, C, Nx, Ny, ncols, 0, twx, twy, Nxm);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

int launch_kernel_spectrum_emb(aefft_ctx* ctx, int R, int C, int Nk, int Nl, int Nx, int Ny, int col0, int ncols, const float* taps,
                               float* emb) {
  const float2 *twx, *twy;
  AE_TRY(get_twiddles(ctx, Nx, &twx));
  AE_TRY(get_twiddles(ctx, Ny, &twy));
  if (ncols <= 0) { ncols = Ny / 2 + 1; col0 = 0; }
  const long long S = (long long)Nx * ncols;
  const int T = Nk * Nl;
  AE_ARG(T <= KE_MAXT);
  dim3 grid((unsigned)((S + KE_BINS - 1) / KE_BINS), (unsigned)((R * C + 255) / 256));
  const bool sep = Nk == Nl && (Nk == 5 || Nk == 3 || Nk == 7) && !getenv("AEFFT_EMB_NOSEP");
  ProfScope prof(ctx, "kernel_spectrum_emb", 4.0 * S * R * C * (sep ? 2.0 * Nk : (double)T), 16.0 * S * R * C);
  if (sep) {
    dim3 g2((unsigned)((ncols + KS_COLS - 1) / KS_COLS), (unsigned)((R * C + 255) / 256), (unsigned)((Nx + KS_ROWS - 1) / KS_ROWS));
    if (Nk == 5) kernel_spectrum_emb_sep_kernel<5, 5><<<g2, 256, 0, ctx->stream>>>(taps, emb, R, C, Nx, Ny, ncols, col0, twx, twy, 0);
    else if (Nk == 3) kernel_spectrum_emb_sep_kernel<3, 3><<<g2, 256, 0, ctx->stream>>>(taps, emb, R, C, Nx, Ny, ncols, col0, twx, twy, 0);
    else kernel_spectrum_emb_sep_kernel<7, 7><<<g2, 256, 0, ctx->stream>>>(taps, emb, R, C, Nx, Ny, ncols, col0, twx, twy, 0);
  } else if (T == 25) kernel_spectrum_emb_kernel_t<25><<<grid, 256, 0, ctx->stream>>>(taps, emb, R, C, Nk, Nl, Nx, Ny, ncols, col0, S, twx, twy);
  else if (T == 9) kernel_spectrum_emb_kernel_t<9><<<grid, 256, 0, ctx->stream>>>(taps, emb, R, C, Nk, Nl, Nx, Ny, ncols, col0, S, twx, twy);
  else if (T == 49) kernel_spectrum_emb_kernel_t<49><<<grid, 256, 0, ctx->stream>>>(taps, emb, R, C, Nk, Nl, Nx, Ny, ncols, col0, S, twx, twy);
  else kernel_spectrum_emb_kernel<<<grid, 256, 0, ctx->stream>>>(taps, emb, R, C, Nk, Nl, Nx, Ny, ncols, col0, S, twx, twy);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// taps[n][k][l] = scale * (pruned inverse DFT of the bin-major spectra z [S][R*C] complex), n = (r, c) or (c, r) when transposed
int launch_binmajor_to_taps(aefft_ctx* ctx, int R, int C, int transpose, int Nk, int Nl, int Nx, int Ny, int col0, int ncols,
                            const float2* z, float* taps, float scale) {
  const float2 *twx, *twy;
  AE_TRY(get_twiddles(ctx, Nx, &twx));
  AE_TRY(get_twiddles(ctx, Ny, &twy));
  if (ncols <= 0) { ncols = Ny / 2 + 1; col0 = 0; }
  const long long S = (long long)Nx * ncols;
  const int T = Nk * Nl, E = R * C;
  AE_ARG(T == 25 || T == 9 || T == 49);
  const int etiles = (E + 127) / 128;
  // CTAs per SM of the reduction grid (AEFFT_TAPS_SPLIT: development knob)
  const int per_sm = getenv("AEFFT_TAPS_SPLIT") ? atoi(getenv("AEFFT_TAPS_SPLIT")) : 4;
  long long nsplit = ((long long)(per_sm > 0 ? per_sm : 4) * ctx->sm_count + etiles - 1) / etiles;
  const bool sep = (Nk == Nl) && (Nk == 5 || Nk == 3 || Nk == 7) && !getenv("AEFFT_TAPS_NOSEP");
  const long long max_split = sep ? Nx : S / BT_BINS;
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > 65535) nsplit = 65535;
  float* part;
  AE_TRY(ctx->getT("tc_taps_part", (size_t)E * nsplit * T, &part));
  dim3 grid(etiles, (unsigned)nsplit);
  {
    ProfScope prof(ctx, "binmajor_to_taps", 4.0 * S * E * (sep ? 2.0 * Nl : (double)T), 8.0 * S * E);
    if (sep) {
      const size_t smem = ((((size_t)ncols * Nl + 1) & ~(size_t)1) + (size_t)2 * BT_ST * 128) * sizeof(float2);
      if (Nk == 5) {
        AE_TRY(ctx->ensure_dyn_smem((const void*)binmajor_to_taps_sep_kernel<5, 5>, smem));
        binmajor_to_taps_sep_kernel<5, 5><<<grid, 128, smem, ctx->stream>>>(z, part, E, R, C, transpose, Nx, Ny, ncols, col0, twx, twy);
      } else if (Nk == 3) {
        AE_TRY(ctx->ensure_dyn_smem((const void*)binmajor_to_taps_sep_kernel<3, 3>, smem));
        binmajor_to_taps_sep_kernel<3, 3><<<grid, 128, smem, ctx->stream>>>(z, part, E, R, C, transpose, Nx, Ny, ncols, col0, twx, twy);
      } else {
        AE_TRY(ctx->ensure_dyn_smem((const void*)binmajor_to_taps_sep_kernel<7, 7>, smem));
        binmajor_to_taps_sep_kernel<7, 7><<<grid, 128, smem, ctx->stream>>>(z, part, E, R, C, transpose, Nx, Ny, ncols, col0, twx, twy);
      }
    } else if (T == 25) binmajor_to_taps_kernel<25><<<grid, 128, 0, ctx->stream>>>(z, part, E, R, C, transpose, Nk, Nl, Nx, Ny, ncols, col0, S, twx, twy);
    else if (T == 9) binmajor_to_taps_kernel<9><<<grid, 128, 0, ctx->stream>>>(z, part, E, R, C, transpose, Nk, Nl, Nx, Ny, ncols, col0, S, twx, twy);
    else binmajor_to_taps_kernel<49><<<grid, 128, 0, ctx->stream>>>(z, part, E, R, C, transpose, Nk, Nl, Nx, Ny, ncols, col0, S, twx, twy);
  }
  const long long total = (long long)E * T;
  taps_final_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(part, taps, total, (int)nsplit, T, scale);
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// forward contraction (conv_k): out~[S][B][2 O] = scale * in~[S][B][2 C] . Wemb[S][2 O][2 C]^T  (+ bias[o] * bias_scale at bin 0)
//   minus `sub` when given (E = O - Xt), and then *mse_out = mse_scale * sum_bins hw * |out|^2 (Hermitian weights);
//   out == nullptr (with mse_out): only the sum is wanted, nothing is stored
int launch_tc_forward(aefft_ctx* ctx, long long S, int B, int C, int O, const float* in, const float* Wemb, float scale,
                      const float* bias, float bias_scale, const float* sub, float* out, float* mse_out, double mse_scale, int ncols,
                      int col0, int Ny) {
  TcOperand A{in, B, 2 * C, 0}, W{Wemb, 2 * O, 2 * C, 0};
  const double bytes = 4.0 * S * (2.0 * B * C + 2.0 * B * O * ((sub ? 1 : 0) + (out ? 1 : 0)) + 4.0 * O * C);
  return launch_bgemm(ctx, "spec_contract_tc", S, A, W, B, 2 * O, 2 * C, EPI_STORE, scale, bias, bias_scale, sub, out, 0, mse_out,
                      mse_scale, ncols, col0, Ny, bytes);
}
// G~[S][B][2 dM] = E~[S][B][2 dD] . conj(F): the Femb[S][2 dD][2 dM] block of O = H F^T read as an MN-major operand
int launch_tc_adjoint(aefft_ctx* ctx, long long S, int B, int dD, int dM, const float* E, const float* Femb, float* G) {
  TcOperand A{E, B, 2 * dD, 0}, W{Femb, 2 * dD, 2 * dM, 1};
  const double bytes = 4.0 * S * (2.0 * B * dD + 2.0 * B * dM + 4.0 * dD * dM);
  return launch_bgemm(ctx, "spec_contract_tc", S, A, W, B, 2 * dM, 2 * dD, EPI_STORE, 1.f, nullptr, 0.f, nullptr, G, 0, nullptr, 0.0,
                      0, 0, 0, bytes);
}
// out[S][nP][nQ][2] = scale * sum_b P~[b][p] conj(Q~[b][q])   (conj_out: the conjugate of that), P~ [S][B][2 nP], Q~ [S][B][2 nQ]
int launch_tc_outer(aefft_ctx* ctx, long long S, int B, int nP, int nQ, const float* P, const float* Q, float scale, int conj_out,
                    float* out) {
  TcOperand A{P, B, 2 * nP, 1}, Bq{Q, B, 2 * nQ, 1};
  const double bytes = 4.0 * S * (2.0 * B * nP + 2.0 * B * nQ + 2.0 * nP * nQ);
  return launch_bgemm(ctx, "spec_outer_tc", S, A, Bq, 2 * nP, 2 * nQ, B, EPI_OUTER, scale, nullptr, 0.f, nullptr, out, conj_out,
                      nullptr, 0.0, 0, 0, 0, bytes);
}

// ---- E = O - X on bin-major spectra, with the Hermitian-weighted sum of |E|^2 (calc_mse, fft_backproplib.cu:480-498)
__global__ void __launch_bounds__(256) bm_sub_mse_kernel(const float4* __restrict__ O, const float4* __restrict__ X, float4* __restrict__ E,
                                                         long long S, long long row4, int ncols, int col0, int Ny,
                                                         double* __restrict__ part) {
  double s = 0.0;
  for (long long w = blockIdx.x; w < S; w += gridDim.x) {
    const int wy = col0 + (int)(w % ncols);
    const float hw = (wy == 0 || wy == Ny / 2) ? 1.f : 2.f;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < row4; i += blockDim.x) {
      const float4 o = __ldg(O + w * row4 + i), x = __ldg(X + w * row4 + i);
      const float4 e = make_float4(o.x - x.x, o.y - x.y, o.z - x.z, o.w - x.w);
      E[w * row4 + i] = e;
      acc += (e.x * e.x + e.y * e.y) + (e.z * e.z + e.w * e.w);
    }
    s += (double)(acc * hw);
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

int launch_bm_sub_mse(aefft_ctx* ctx, long long S, long long rowlen, const float* O, const float* X, float* E, float* mse_out,
                      double mse_scale, int ncols, int col0, int Ny) {
  AE_ARG(rowlen % 4 == 0 && S > 0);
  if (ncols <= 0) { ncols = Ny / 2 + 1; col0 = 0; }
  const int grid = (int)(S < 8LL * ctx->sm_count ? S : 8LL * ctx->sm_count);
  double* part;
  AE_TRY(ctx->getT("tc_sq_part", (size_t)(grid > 4 * ctx->sm_count ? grid : 4 * ctx->sm_count), &part));
  {
    ProfScope prof(ctx, "spec_bm_sub_mse", 0.0, 12.0 * S * rowlen);
    bm_sub_mse_kernel<<<grid, 256, 0, ctx->stream>>>((const float4*)O, (const float4*)X, (float4*)E, S, rowlen / 4, ncols, col0, Ny, part);
    ctx->launches++;
  }
  if (mse_out) {
    sq_final_kernel<<<1, 256, 0, ctx->stream>>>(part, grid, mse_scale, mse_out);
    ctx->launches++;
  }
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// ---- spectral pooling on bin-major data: target bin (i, j) <- source bin (si, sj) or zero; same index map as spec_resize_kernel
__global__ void __launch_bounds__(256) bm_resize_kernel(const float4* __restrict__ in, float4* __restrict__ out, long long row4, int Nx,
                                                        int Ny, int Nxs, int Nys) {
  const int Nyr = Ny / 2 + 1, Nyrs = Nys / 2 + 1;
  const long long w = blockIdx.x;
  const int i = (int)(w / Nyrs), j = (int)(w - (long long)i * Nyrs);
  int si = -1, sj = -1;
  if (Nxs <= Nx) {
    si = i < Nxs / 2 ? i : (i == Nxs / 2 ? Nx / 2 : i + Nx - Nxs);
    sj = j < Nyrs - 1 ? j : Nyr - 1;
  } else {
    if (i < Nx / 2) si = i;
    else if (i > Nxs - Nx / 2) si = i - Nxs + Nx;
    else if (i == Nxs / 2) si = Nx / 2;
    if (j < Nyr - 1) sj = j;
    else if (j == Nyrs - 1) sj = Nyr - 1;
  }
  float4* o = out + w * row4;
  if (si < 0 || sj < 0) {
    for (long long t = threadIdx.x; t < row4; t += blockDim.x) o[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float4* s = in + ((long long)si * Nyr + sj) * row4;
  for (long long t = threadIdx.x; t < row4; t += blockDim.x) o[t] = __ldg(s + t);
}

int launch_bm_resize(aefft_ctx* ctx, long long rowlen, int Nx, int Ny, int Nxs, int Nys, const float* in, float* out) {
  AE_ARG(rowlen % 4 == 0);
  const long long bins = (long long)Nxs * (Nys / 2 + 1);
  const long long src_bins = (long long)(Nxs <= Nx ? Nxs : Nx) * ((Nxs <= Nx ? Nys : Ny) / 2 + 1);
  ProfScope prof(ctx, "spec_resize_bm", 0.0, 4.0 * rowlen * (bins + src_bins));
  bm_resize_kernel<<<(unsigned)bins, 256, 0, ctx->stream>>>((const float4*)in, (float4*)out, rowlen / 4, Nx, Ny, Nxs, Nys);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

int launch_tc_dc_terms(aefft_ctx* ctx, int B, int dM, int dD, const float* G, const float* E, const float* bias_b, float* dFt,
                       float* db, float* dp, float gs, float fs, float corr_scale) {
  dc_terms_kernel<<<1, 512, 2 * dD * sizeof(double), ctx->stream>>>(G, E, bias_b, dFt, db, dp, B, dM, dD, gs, fs, corr_scale);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// Whether the Gram form of the gradients pays: gram_grad_kernel runs on the CUDA cores (4 B dD^2 + 8 dM dD^2 FMA per bin), the
// form it replaces streams 6 B (dD + dM) + 8 dD dM floats per bin through HBM instead of 4 B dD + 8 dD dM.
static size_t gram_smem(int B, int dD, int dM) {
  const int tr = dD >= 32 ? 4 : (dD >= 16 ? 2 : 1), ng = 256 / ((dD / tr) * (dD / tr));
  const size_t ex = 2 * (size_t)B * dD, cf = 2 * (size_t)dD * dM, pp = ng > 1 ? (size_t)ng * dD * (dD + 1) : 0;
  size_t m = ex > cf ? ex : cf;
  if (pp > m) m = pp;
  return ((size_t)dD * (dD + 1) + m) * sizeof(float2);
}
bool spec_tc_gram_pays(int B, int dD, int dM) {
  if (getenv("AEFFT_NO_GRAM")) return false;
  if (dD != 8 && dD != 16 && dD != 32 && dD != 64) return false;
  const int no = dM * dD / 256;
  if (dM * dD >= 256 ? (dM * dD % 256 != 0 || (no != 1 && no != 2 && no != 4 && no != 8 && no != 16)) : false) return false;
  if (gram_smem(B, dD, dM) > 200 * 1024) return false;
  if (getenv("AEFFT_FORCE_GRAM")) return true;
  const double clk_new = (4.0 * B * dD * dD + 8.0 * dM * dD * dD) / 128.0;  // 128 FMA lanes per clock (measured at config 3:
                                                                             // 0.53 vs 1.2 ms at 16 -> 32, 0.51 vs 0.68 ms at 32 -> 64)
  const double clk_old = (6.0 * B * (dD + dM) - 4.0 * B * dD) * 4.0 / 15.6;  // 4.4 TB/s over 148 SMs at 1.9 GHz
  return clk_new < clk_old;
}

int launch_tc_gram_grad(aefft_ctx* ctx, long long S, int B, int dM, int dD, const float* E, const float* X, const float* Cemb,
                        const float* Femb, float gs, float* dCt, float* dFt) {
  const size_t smem = gram_smem(B, dD, dM);
  const int no = dM * dD >= 256 ? dM * dD / 256 : 1;
  ProfScope prof(ctx, "spec_gram_grad", 8.0 * S * ((double)B * dD * dD + 2.0 * dM * dD * dD),
                 4.0 * S * (4.0 * B * dD + 8.0 * dM * dD));
#define AEFFT_GRAM(dd, n)                                                                                                    \
  if (dD == dd && no == n) {                                                                                                 \
    AE_TRY(ctx->ensure_dyn_smem((const void*)gram_grad_kernel<dd, n>, smem));                                                \
    gram_grad_kernel<dd, n><<<(unsigned)S, 256, smem, ctx->stream>>>(E, X, Cemb, Femb, (float2*)dCt, (float2*)dFt, B, dM, gs); \
    ctx->launches++;                                                                                                         \
    AE_CUDA(cudaGetLastError());                                                                                             \
    return AEFFT_OK;                                                                                                         \
  }
  AEFFT_GRAM(8, 1) AEFFT_GRAM(8, 2) AEFFT_GRAM(8, 4) AEFFT_GRAM(8, 8)
  AEFFT_GRAM(16, 1) AEFFT_GRAM(16, 2) AEFFT_GRAM(16, 4) AEFFT_GRAM(16, 8) AEFFT_GRAM(16, 16)
  AEFFT_GRAM(32, 1) AEFFT_GRAM(32, 2) AEFFT_GRAM(32, 4) AEFFT_GRAM(32, 8) AEFFT_GRAM(32, 16)
  AEFFT_GRAM(64, 2) AEFFT_GRAM(64, 4) AEFFT_GRAM(64, 8) AEFFT_GRAM(64, 16)
#undef AEFFT_GRAM
  return AEFFT_ERR_UNSUPPORTED;
}

int launch_tc_dc_terms_gram(aefft_ctx* ctx, int B, int dM, int dD, const float* E, const float* Femb, const float* bias_b, float* dFt,
                            float* db, float* dp, float gs, float fs, float norm) {
  dc_terms_gram_kernel<<<1, 128, 2 * dD * sizeof(double), ctx->stream>>>(E, Femb, bias_b, dFt, db, dp, B, dM, dD, gs, fs, norm);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

}  // namespace aefft

// Diagnostic entry (tests/test_spec_tc_gpu.py): one batched-over-bins real GEMM of the tensor path on device buffers.
extern "C" int aefft_spec_bin_gemm(aefft_ctx* ctx, int64_t S, const float* a, int a_rows, int a_cols, int a_mn, const float* b,
                                   int b_rows, int b_cols, int b_mn, int M, int N, int K, int outer, int conj_out, float scale,
                                   float* out) {
  using namespace aefft;
  AE_ARG(ctx && a && b && out);
  AE_CUDA(cudaSetDevice(ctx->device));
  TcOperand A{a, a_rows, a_cols, a_mn}, B{b, b_rows, b_cols, b_mn};
  return launch_bgemm(ctx, outer ? "spec_outer_tc" : "spec_contract_tc", S, A, B, M, N, K, outer ? EPI_OUTER : EPI_STORE, scale, nullptr,
                      0.f, nullptr, out, conj_out, nullptr, 0.0, 0, 0, 0, 0.0);
}
