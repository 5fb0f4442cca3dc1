// The multiobjective (kernel-diversity) term for 5 x 5 kernels on the tensor cores (gradient_diff, fft_backproplib.cu:709-753):
//   xd[a] = x[a] * sum_b w_ab - sum_b w_ab x[b],   w_ab = 1 / |x[a] - x[b]|^2 for kernels that share neither index.
// An all-pairs sum over n = dM dD kernels of 25 taps: 2 n^2 (25 + 25) flops per tensor.  The CUDA-core kernel
// (spectral_kernels.cu) runs it at ~35 TFLOP/s, half of the fp32 peak, bound by issued instructions.  Here both halves are
// GEMMs, flash-attention shaped, with only the weights computed by CUDA cores:
//   MMA 1:  S[128 a x 64 b] = A Bt          (K = 32: 25 taps, zero padding)            -> dot products in tensor memory
//   epilogue (8 warps, one thread per row a): d^2 = |a|^2 + |b|^2 - 2 S, w = 1 / d^2 (0 for excluded pairs), written to
//           shared memory as the K-major A operand of
//   MMA 2:  V[128 a x 32] += W[128 x 64] Xb [64 x 32]   (column 25 of Xb is 1: V[a][25] = sum_b w_ab)  -> accumulates in tensor
//           memory over all b tiles of the CTA.
// kind::tf32 with the 3xTF32 split (hi = x & 0xffffe000, lo = x - hi; hi hi + hi lo + lo hi) on both products: fp32-grade dot
// products and sums.  Pairs whose dot-product distance cancels (d^2 < 1 % of |a|^2 + |b|^2: near-duplicate kernels) get their
// distance from directly summed differences in the epilogue, as the reference computes it (:722-741).
// Measured at 128 -> 256 channels (32 768 kernels per tensor): 3.75 ms per call against 6.1 ms of the CUDA-core kernel.  Knock-out
// runs of an instrumented build: without MMA 2 -1.1 ms, without MMA 1 -0.1 ms, without the epilogue arithmetic -1.0 ms, with all
// three removed 2.0 ms remain -- the per-tile hand-offs (TMA -> split -> MMA 1 -> epilogue -> MMA 2 -> epilogue of the next tile, a
// serial loop through the single W buffer) are half of the kernel.  A second W buffer would overlap them; it does not fit next to
// the A tile and three B stages (231 KB), it would with the A operand in tensor memory.  Two issuing threads (one per MMA) were
// measured slower (4.0 ms), a first version with the near-duplicate branch inside the element loop 7.9 ms.
// Roles (704 threads, one CTA per SM): warp 0 TMA producer, warp 1 MMA issuer, warps 2-17 epilogue (four warps per TMEM lane
// quarter, 16 columns of S each: the epilogue of tile k and the second MMA of tile k are serial through the single W buffer, so
// its latency is what the kernel runs at), warps 18-21 split the landed tiles into hi / lo.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "pipe.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace aefft {

namespace {

using namespace umma;

constexpr int GT_EPI_WARPS = 16, GT_THREADS = 32 * (2 + GT_EPI_WARPS + 4), GT_NST = 3, GT_T = 25;
constexpr uint32_t GT_A_LO = 16384, GT_W_HI = 32768, GT_W_LO = 65536, GT_ST0 = 98304, GT_STAGE = 34816;
constexpr uint32_t GT_BK_LO = 8192, GT_BM_HI = 16384, GT_BM_LO = 24576, GT_AUX = 32768;
constexpr size_t GT_SMEM = GT_ST0 + GT_NST * GT_STAGE + 1024;

struct GtParams {
  CUtensorMap a_map, bk_map, bm_map, aux_map;
  const float *c, *f;        // the kernels [n][25]
  const float4* aux;         // [2][npad]: (|x|^2 or +inf beyond n, i1, i2, 0)
  float *cd, *fd, *part;
  int dM, dD, n, tile0, nt64, nbt;  // first owned 64-kernel tile, owned 64-kernel tiles, b tiles of 64 in total
  long long npad;
};

__device__ __forceinline__ uint64_t gdesc_k(uint32_t addr) { return make_desc(addr, 16, 1024) | ((uint64_t)2 << 61); }
__device__ __forceinline__ uint64_t gdesc_mn(uint32_t addr) { return make_desc(addr, 4096, 512) | ((uint64_t)1 << 61); }
__host__ __device__ constexpr uint32_t gidesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void gmma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void split4(uint4* hi, float4* lo, int i) {
  uint4 u = hi[i];
  const float4 x = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
  u.x &= 0xffffe000u; u.y &= 0xffffe000u; u.z &= 0xffffe000u; u.w &= 0xffffe000u;
  hi[i] = u;
  lo[i] = make_float4(x.x - __uint_as_float(u.x), x.y - __uint_as_float(u.y), x.z - __uint_as_float(u.z), x.w - __uint_as_float(u.w));
}

// staging copies: xp32 [2][npad][32] = (25 taps, 1, 0 ...; zero rows beyond n), aux [2][npad] = (|x|^2 | +inf, i1, i2, 0)
__global__ void gdiff_tc_pack_kernel(const float* __restrict__ c, const float* __restrict__ f, float* __restrict__ xp32,
                                     float4* __restrict__ aux, int dM, int dD, long long npad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * npad) return;
  const int isf = i >= npad;
  const long long b = isf ? i - npad : i;
  const int n = dM * dD, n2 = isf ? dM : dD;
  float* o = xp32 + i * 32;
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 32; t++) {
    float v = 0.f;
    if (b < n) {
      if (t < GT_T) { v = (isf ? f : c)[b * GT_T + t]; s = fmaf(v, v, s); }
      else if (t == GT_T) v = 1.f;
    }
    o[t] = v;
  }
  const int b1 = b < n ? (int)(b / n2) : -2, b2 = b < n ? (int)(b - (long long)b1 * n2) : -2;
  aux[i] = make_float4(b < n ? s : __int_as_float(0x7f800000), __int_as_float(b1), __int_as_float(b2), 0.f);
}

__global__ void __launch_bounds__(GT_THREADS, 1) gdiff_tc_kernel(const __grid_constant__ GtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t a_full, a_ready, b_full[GT_NST], b_ready[GT_NST], b_empty[GT_NST], s_full[2], s_empty[2], w_full,
      w_empty, v_full;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int isf = blockIdx.y;
  const int bt_lo = (int)((long long)p.nbt * blockIdx.z / gridDim.z), bt_hi = (int)((long long)p.nbt * (blockIdx.z + 1) / gridDim.z);
  const int ntile = bt_hi - bt_lo;
  const long long a_row0 = ((long long)p.tile0 + 2 * blockIdx.x) * 64;  // first kernel of this CTA's 128-row tile
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid == 32) {
    mbar_init(&a_full, 1); mbar_init(&a_ready, 4);
    for (int s = 0; s < GT_NST; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_ready[s], 4); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], GT_EPI_WARPS); }
    mbar_init(&w_full, GT_EPI_WARPS); mbar_init(&w_empty, 1); mbar_init(&v_full, 1);
    fence_mbar_init();
  }
  if (tid == 0) {
    tma::tma_prefetch_desc(&p.a_map); tma::tma_prefetch_desc(&p.bk_map); tma::tma_prefetch_desc(&p.bm_map);
    tma::tma_prefetch_desc(&p.aux_map);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma::mbar_expect_tx(&a_full, 16384);
      tma::tma_load_3d(smem, &p.a_map, 0, (int)a_row0, isf, &a_full);
      Ring r(GT_NST);
      for (int k = 0; k < ntile; k++) {
        mbar_wait_role<false>(&b_empty[r.slot], r.phase ^ 1);
        uint8_t* st = smem + GT_ST0 + (size_t)r.slot * GT_STAGE;
        const int b0 = (bt_lo + k) * 64;
        tma::mbar_expect_tx(&b_full[r.slot], 8192 + 8192 + 1024);
        tma::tma_load_3d(st, &p.bk_map, 0, b0, isf, &b_full[r.slot]);
        tma::tma_load_3d(st + GT_BM_HI, &p.bm_map, 0, b0, isf, &b_full[r.slot]);
        tma::tma_load_3d(st + GT_BM_HI + 4096, &p.bm_map, 0, b0 + 32, isf, &b_full[r.slot]);
        tma::tma_load_3d(st + GT_AUX, &p.aux_map, 0, b0, isf, &b_full[r.slot]);
        r.next();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0 && ntile > 0) {
      const uint32_t id1 = gidesc_tf32(128, 64, 0, 0), id2 = gidesc_tf32(128, 32, 0, 1);
      const uint32_t sA = smem_u32(smem), sW = sA + GT_W_HI;
      auto mma2 = [&](int j) {  // V += W(j) Xb(j)
        const uint32_t st = sA + GT_ST0 + (uint32_t)(j % GT_NST) * GT_STAGE;
        mbar_wait_role<false>(&w_full, (uint32_t)(j & 1));
        fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          const uint32_t ao = sW + (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u, bo = st + GT_BM_HI + (uint32_t)ks * 1024u;
          const uint64_t ah = gdesc_k(ao), al = gdesc_k(ao + (GT_W_LO - GT_W_HI));
          const uint64_t bh = gdesc_mn(bo), bl = gdesc_mn(bo + (GT_BM_LO - GT_BM_HI));
          gmma_tf32(tb + 128, ah, bh, id2, (j > 0 || ks > 0) ? 1u : 0u);
          gmma_tf32(tb + 128, ah, bl, id2, 1);
          gmma_tf32(tb + 128, al, bh, id2, 1);
        }
        commit(&b_empty[j % GT_NST]);
        commit(&w_empty);
      };
      mbar_wait_role<false>(&a_ready, 0);
      fence_after_sync();
      Ring r(GT_NST);
      for (int k = 0; k < ntile; k++) {
        mbar_wait_role<false>(&b_ready[r.slot], r.phase);
        mbar_wait_role<false>(&s_empty[k & 1], (uint32_t)(((k >> 1) & 1) ^ 1));
        fence_after_sync();
        const uint32_t st = sA + GT_ST0 + (uint32_t)r.slot * GT_STAGE, d = tb + (uint32_t)(k & 1) * 64u;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
          const uint64_t ah = gdesc_k(sA + ks * 32u), al = gdesc_k(sA + GT_A_LO + ks * 32u);
          const uint64_t bh = gdesc_k(st + ks * 32u), bl = gdesc_k(st + GT_BK_LO + ks * 32u);
          gmma_tf32(d, ah, bh, id1, ks > 0 ? 1u : 0u);
          gmma_tf32(d, ah, bl, id1, 1);
          gmma_tf32(d, al, bh, id1, 1);
        }
        commit(&s_full[k & 1]);
        if (k > 0) mma2(k - 1);
        r.next();
      }
      mma2(ntile - 1);
      commit(&v_full);
    }
  } else if (warp < 2 + GT_EPI_WARPS) {
    // ------------------------------------------------------------------------------------------ epilogue (weights)
    const int q = warp & 3, cg = (warp - 2) >> 2;  // TMEM lane quarter, 16-column group of the S tile
    const int r = 32 * q + lane;
    const long long a = a_row0 + r;
    const bool a_ok = a < p.n && a < ((long long)p.tile0 + p.nt64) * 64;
    const float* x = isf ? p.f : p.c;
    const float4 au_a = a < p.npad ? __ldg(p.aux + (long long)isf * p.npad + a) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float na = a_ok ? au_a.x : 0.f;
    const int a1 = a_ok ? __float_as_int(au_a.y) : -1, a2 = a_ok ? __float_as_int(au_a.z) : -1;
    const uint32_t lane_base = tb + ((uint32_t)(32 * q) << 16);
    uint8_t* w_hi = smem + GT_W_HI + (size_t)(cg >> 1) * 16384 + (size_t)r * 128;
    const int ch0 = 4 * (cg & 1);  // first 16-byte chunk of this warp's columns inside the 32-column atom
    for (int k = 0; k < ntile; k++) {
      const uint8_t* st = smem + GT_ST0 + (size_t)(k % GT_NST) * GT_STAGE;
      mbar_wait_role<false>(&s_full[k & 1], (uint32_t)((k >> 1) & 1));
      mbar_wait_role<false>(&w_empty, (uint32_t)((k & 1) ^ 1));
      fence_after_sync();
      const float4* aux_s = reinterpret_cast<const float4*>(st + GT_AUX) + cg * 16;
      unsigned near = 0;  // columns whose dot-product distance cancels: patched below (a branch per element cost the loop its ILP)
      float v[16];
      tmem_ld16(lane_base + (uint32_t)(k & 1) * 64u + (uint32_t)(cg * 16), v);
#pragma unroll
      for (int e = 0; e < 16; e++) {
        const float4 au = aux_s[e];
        const float nn = na + au.x;
        const float d2 = fmaf(-2.f, v[e], nn);
        const bool on = (__float_as_int(au.y) != a1) & (__float_as_int(au.z) != a2);
        const bool cancels = d2 < 0.01f * nn;
        if (on && cancels) near |= 1u << e;
        float w;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(d2));
        v[e] = (on && !cancels) ? w : 0.f;
      }
      // hi / lo split, 16-byte chunks of the K-major SWIZZLE_128B tile (chunk c of row r at c ^ (r & 7))
#pragma unroll
      for (int i = 0; i < 4; i++) {
        uint4 h = make_uint4(__float_as_uint(v[4 * i]) & 0xffffe000u, __float_as_uint(v[4 * i + 1]) & 0xffffe000u,
                             __float_as_uint(v[4 * i + 2]) & 0xffffe000u, __float_as_uint(v[4 * i + 3]) & 0xffffe000u);
        const float4 l = make_float4(v[4 * i] - __uint_as_float(h.x), v[4 * i + 1] - __uint_as_float(h.y),
                                     v[4 * i + 2] - __uint_as_float(h.z), v[4 * i + 3] - __uint_as_float(h.w));
        const uint32_t off = (uint32_t)(((ch0 + i) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(w_hi + off) = h;
        *reinterpret_cast<float4*>(w_hi + (GT_W_LO - GT_W_HI) + off) = l;
      }
      while (near) {  // near-duplicate kernels (rare): directly summed differences, x[b] = hi + lo from the K-major tile
        const int cc = __ffs(near) - 1;
        near &= near - 1;
        const int j = cg * 16 + cc;
        const uint8_t* row = st + (size_t)j * 128;
        float s2 = 0.f;
#pragma unroll
        for (int t4 = 0; t4 < 7; t4++) {
          const uint32_t off = (uint32_t)((t4 ^ (j & 7)) << 4);
          const float4 h = *reinterpret_cast<const float4*>(row + off), l = *reinterpret_cast<const float4*>(row + GT_BK_LO + off);
          const float xb[4] = {h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w};
#pragma unroll
          for (int u = 0; u < 4; u++)
            if (4 * t4 + u < GT_T) { const float dlt = (a_ok ? __ldg(x + a * GT_T + 4 * t4 + u) : 0.f) - xb[u]; s2 = fmaf(dlt, dlt, s2); }
        }
        float w;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(s2));
        const uint32_t hb = __float_as_uint(w) & 0xffffe000u;
        const uint32_t off = (uint32_t)(((ch0 + (cc >> 2)) ^ (r & 7)) << 4) + (uint32_t)(cc & 3) * 4u;
        *reinterpret_cast<uint32_t*>(w_hi + off) = hb;
        *reinterpret_cast<float*>(w_hi + (GT_W_LO - GT_W_HI) + off) = w - __uint_as_float(hb);
      }
      fence_proxy_async();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) { tma::mbar_arrive(&w_full); tma::mbar_arrive(&s_empty[k & 1]); }
    }
    if (ntile > 0) {
      mbar_wait_role<false>(&v_full, 0);
      fence_after_sync();
      if (cg == 0) {
        float v[32];
        tmem_ld16(lane_base + 128u, v);
        tmem_ld16(lane_base + 128u + 16u, v + 16);
        if (a_ok) {
          const float sw = v[GT_T];
          if (gridDim.z == 1) {
            float* xd = (isf ? p.fd : p.cd) + a * GT_T;
#pragma unroll
            for (int t = 0; t < GT_T; t++) xd[t] = __ldg(x + a * GT_T + t) * sw - v[t];
          } else {
            const long long local = a - (long long)p.tile0 * 64;
            float* o = p.part + ((((long long)isf * gridDim.z + blockIdx.z) * p.nt64) * 64 + local) * (GT_T + 1);
#pragma unroll
            for (int t = 0; t < GT_T; t++) o[t] = v[t];
            o[GT_T] = sw;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ splitters
    const int t = tid - 32 * (2 + GT_EPI_WARPS);  // 0..127
    {
      mbar_wait_role<false>(&a_full, 0);
      uint4* ah = reinterpret_cast<uint4*>(smem);
      float4* al = reinterpret_cast<float4*>(smem + GT_A_LO);
#pragma unroll 4
      for (int i = t; i < 1024; i += 128) split4(ah, al, i);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // the 1 of column 25 belongs to MMA 2 only: the dot products must not see it (row t: chunk 6, element 1)
      const uint32_t off = (uint32_t)t * 128u + (uint32_t)((6 ^ (t & 7)) << 4) + 4u;
      *reinterpret_cast<float*>(smem + off) = 0.f;
      *reinterpret_cast<float*>(smem + GT_A_LO + off) = 0.f;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(&a_ready);
    }
    Ring r(GT_NST);
    for (int k = 0; k < ntile; k++) {
      mbar_wait_role<false>(&b_full[r.slot], r.phase);
      uint8_t* st = smem + GT_ST0 + (size_t)r.slot * GT_STAGE;
      uint4* kh = reinterpret_cast<uint4*>(st);
      float4* kl = reinterpret_cast<float4*>(st + GT_BK_LO);
      uint4* mh = reinterpret_cast<uint4*>(st + GT_BM_HI);
      float4* ml = reinterpret_cast<float4*>(st + GT_BM_LO);
#pragma unroll
      for (int i = t; i < 512; i += 128) { split4(kh, kl, i); split4(mh, ml, i); }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(&b_ready[r.slot]);
      r.next();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 256);
}

}  // namespace

// returns AEFFT_ERR_UNSUPPORTED when the shape is not taken (the caller falls back to the CUDA-core kernel)
int launch_gradient_diff_tc(aefft_ctx* ctx, int dM, int dD, const float* c, const float* f, float* cd, float* fd, int t0, int nt,
                            int nchunks, float* part) {
  if (getenv("AEFFT_NO_GDIFF_TC")) return AEFFT_ERR_UNSUPPORTED;
  const int n = dM * dD;
  int min_n = 2048;  // small kernel counts: the CUDA-core kernel's launch is shorter (AEFFT_GDIFF_TC_MIN: the parity tests)
  if (const char* e = getenv("AEFFT_GDIFF_TC_MIN")) min_n = atoi(e);
  if (n < min_n || n < 128 || nt < 1) return AEFFT_ERR_UNSUPPORTED;
  const long long npad = ((long long)n + 127) / 128 * 128;
  float *xp32, *auxf;
  AE_TRY(ctx->getT("gdtc_xp", (size_t)2 * npad * 32, &xp32));
  AE_TRY(ctx->getT("gdtc_aux", (size_t)2 * npad * 4, &auxf));
  GtParams p;
  memset(&p, 0, sizeof(p));
  int rc = tma::make_tmap_3d_f32(&p.a_map, xp32, 32, (uint64_t)npad, 2, 32, 128, 1, 1);
  rc |= tma::make_tmap_3d_f32(&p.bk_map, xp32, 32, (uint64_t)npad, 2, 32, 64, 1, 1);
  rc |= tma::make_tmap_3d_f32(&p.bm_map, xp32, 32, (uint64_t)npad, 2, 32, 32, 1, 2);
  rc |= tma::make_tmap_3d_f32(&p.aux_map, auxf, 4, (uint64_t)npad, 2, 4, 64, 1, 0);
  if (rc != 0) return AEFFT_ERR_UNSUPPORTED;
  p.c = c; p.f = f; p.aux = (const float4*)auxf; p.cd = cd; p.fd = fd; p.part = part;
  p.dM = dM; p.dD = dD; p.n = n; p.tile0 = t0; p.nt64 = nt; p.nbt = (n + 63) / 64; p.npad = npad;
  AE_TRY(ctx->ensure_dyn_smem((const void*)gdiff_tc_kernel, GT_SMEM));
  gdiff_tc_pack_kernel<<<(unsigned)((2 * npad + 127) / 128), 128, 0, ctx->stream>>>(c, f, xp32, (float4*)auxf, dM, dD, npad);
  dim3 grid((unsigned)((nt + 1) / 2), 2, (unsigned)nchunks);
  gdiff_tc_kernel<<<grid, GT_THREADS, GT_SMEM, ctx->stream>>>(p);
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

}  // namespace aefft
