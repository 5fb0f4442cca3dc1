// Coordinate-space convolution (forward conv and its transpose) as an implicit GEMM on the 5th-generation tensor
// cores: tcgen05.mma, operands in shared memory, fp32 accumulators in TMEM, tcgen05.ld epilogue.
//
//   out[b][o](i,j) = bias[o] + sum_{c,tk,tl} W(o,c,tk,tl) * S[b][c](i+ai0+tk, j+aj0+tl)        (same contract as launch_conv)
//
// GEMM view, per tap (tk,tl):  D[pixel][o] += A_tap[pixel][c] * W_tap[o][c]   (M = 128 pixels, N = out channels, K = 16
// input channels per MMA).  The input tile (+halo) is staged ONCE in shared memory as channel-interleaved planes
//   A[pass][kchunk][h][8 x bf16]      h = halo-row * PJ + halo-col (linear halo pixel), 16 bytes per pixel per plane,
// which is exactly the SWIZZLE_NONE K-major canonical layout (8-pixel core matrices of 128 contiguous bytes), so the
// operand of tap (tk,tl) is the SAME buffer with the descriptor start address advanced by (tk*PJ + tl)*16 bytes:
// no im2col copy, 25 taps reuse one staged tile.  An M-block is 128 consecutive linear halo pixels; the NL-1 halo
// columns at the end of each row produce garbage rows of D that are simply not stored (<= 6 % of the MMA work).
//
// fp32 parity through bf16 tensor cores: every fp32 operand is split x = hi + lo (two bf16), and
// D = A_hi W_hi + A_hi W_lo + A_lo W_hi accumulates in fp32 (dropped terms ~2^-16 relative): "BF16X3" mode,
// ~1e-5 relative L2 against the fp32 oracle.  AEFFT_PRECISION_BF16 issues only the first product (documented looser
// tolerance).  Weights are pre-split and pre-laid-out by weight_prep_kernel (tiny) so staging them is a plain copy.
#include "common.cuh"
#include "umma.cuh"

namespace aefft {

using namespace umma;

constexpr int TC_THREADS = 256;
constexpr int KC = 16;  // input channels per K stage (= one bf16 MMA K)

struct ConvTcParams {
  const float* src0;
  const float* src1;
  const uint4* wprep;  // [KS][2 pass][T][2 kchunk][N][8 bf16]
  const float* bias;
  float* out;
  float pre_div;
  int C, O, N, Nx, Ny;
  int NK, NL, T;
  int ai0, aj0, lo;
  int PJ, TI, MB, HP, KS;
  int tiles_per_frame;
  long long n_tiles;
  int passes;  // 3 = BF16X3, 1 = BF16
  int kpack;   // C <= 8: one channel plane, the two 16-byte K chunks of an MMA are taps (tl, tl+1) (LBO = one pixel)
  int TE;      // MMAs per (M-block, pass): NK*NL, or NK*ceil(NL/2) when kpack
  uint32_t tmem_cols;
};

// w(o,c,k,l) fp32 -> bf16 hi/lo in the shared-memory image of each K stage:
//   wprep[ks][pass][t][kchunk][n][e]   c = ks*16 + kchunk*8 + e,  t = tk*NL + tl (window position),  zero padded
// kpack (C <= 8): t = tk*ceil(NL/2) + pair, kchunk = tap within the pair (tl = 2*pair + kchunk), c = e.
__global__ void weight_prep_kernel(const float* __restrict__ w, long long w_so, long long w_sc, int C, int O, int N, int NK,
                                   int NL, int flip, int KS, int kpack, __nv_bfloat16* __restrict__ wprep) {
  const int NLP = kpack ? (NL + 1) / 2 : NL;
  const int T = NK * NLP;
  const long long per_pass = (long long)T * 2 * N * 8;
  const long long total = (long long)KS * per_pass;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int e = idx % 8;
    const int n = (idx / 8) % N;
    const int kchunk = (idx / (8 * N)) % 2;
    const int t = (idx / (16 * N)) % T;
    const int ks = idx / per_pass;
    const int c = kpack ? e : ks * KC + kchunk * 8 + e;
    const int tk = t / NLP, tl = kpack ? 2 * (t % NLP) + kchunk : t % NLP;
    const int k = flip ? NK - 1 - tk : tk, l = flip ? NL - 1 - tl : tl;
    float v = 0.f;
    if (c < C && n < O && tl < NL) v = w[n * w_so + c * w_sc + k * NL + l];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    const long long base = (long long)ks * 2 * per_pass + (idx % per_pass);
    wprep[base] = hi;
    wprep[base + per_pass] = lo;
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(ConvTcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int w_bytes = 64 * p.TE * p.N;  // both passes of one K stage
  const int a_plane = p.HP * 16;        // one 8-channel plane
  const int nkc = p.kpack ? 1 : 2;      // planes per pass
  unsigned char* Wsm = smem;
  unsigned char* Asm = smem + w_bytes;  // [pass][plane][HP][16]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + w_bytes + 2 * nkc * a_plane);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(tmem_slot, p.tmem_cols);
  if (tid == 32) {
    mbar_init(bar, 4);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  uint32_t phase = 0;
  const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
  const long long plane = (long long)p.Nx * p.Ny;
  const uint32_t Wsm_addr = smem_u32(Wsm), Asm_addr = smem_u32(Asm);
  const int HI = p.TI + p.NK - 1;
  bool w_resident = false;

  for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const long long b = tile / p.tiles_per_frame;
    const int i0 = (int)(tile % p.tiles_per_frame) * p.TI;
    const float* s0 = p.src0 + b * p.C * plane;
    const float* s1 = p.src1 ? p.src1 + b * p.C * plane : nullptr;
    for (int ks = 0; ks < p.KS; ks++) {
      // ---- stage the weights of this K stage (kept across tiles when there is a single stage) ----
      if (!(p.KS == 1 && w_resident)) {
        const uint4* src = p.wprep + (size_t)ks * (w_bytes / 16);
        uint4* dst = reinterpret_cast<uint4*>(Wsm);
        for (int i = tid; i < w_bytes / 16; i += TC_THREADS) dst[i] = __ldg(src + i);
        w_resident = true;
      }
      // ---- stage the input halo tile: one thread = one halo pixel x 8 channels -> one 16-byte store per pass.
      //      All global loads of an item are issued before any is used (addresses clamped, values masked afterwards).
      for (int idx = tid; idx < nkc * p.HP; idx += TC_THREADS) {
        const int kchunk = idx / p.HP, h = idx - kchunk * p.HP;
        const int r = h / p.PJ, col = h - r * p.PJ;
        const int si = i0 + p.ai0 + r, sj = p.aj0 + col;
        const bool inb = r < HI && si >= p.lo && si < p.Nx && sj >= p.lo && sj < p.Ny;
        const int c0 = ks * KC + kchunk * 8;
        const long long pix = inb ? (long long)si * p.Ny + sj : 0;
        float v[8], u[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const int c = min(c0 + e, p.C - 1);
          v[e] = __ldg(s0 + (long long)c * plane + pix);
        }
        if (s1) {
#pragma unroll
          for (int e = 0; e < 8; e++) {
            const int c = min(c0 + e, p.C - 1);
            u[e] = __ldg(s1 + (long long)c * plane + pix);
          }
#pragma unroll
          for (int e = 0; e < 8; e++) v[e] -= u[e];
        } else if (p.pre_div != 0.f) {
          const float rinv = 1.f / p.pre_div;  // exact for the power-of-two channel counts; 1 ulp otherwise
#pragma unroll
          for (int e = 0; e < 8; e++) v[e] *= rinv;
        }
        __nv_bfloat16 hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
          if (!inb || c0 + e >= p.C) v[e] = 0.f;
          split_bf16(v[e], hi[e], lo[e]);
        }
        uint4 qh = make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
        uint4 ql = make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
        *reinterpret_cast<uint4*>(Asm + (size_t)kchunk * a_plane + (size_t)h * 16) = qh;
        *reinterpret_cast<uint4*>(Asm + (size_t)(nkc + kchunk) * a_plane + (size_t)h * 16) = ql;
      }
      fence_proxy_async();
      __syncthreads();
      // ---- issue every MMA of this stage, then commit to the mbarrier ----
      if (warp < 4 && lane == 0) {  // four issuing threads, M-blocks interleaved between them
        fence_after_sync();
        const uint32_t w_pass = (uint32_t)w_bytes / 2, w_tap = 32u * p.N;
        const uint32_t a_lbo = p.kpack ? 16u : (uint32_t)a_plane;
        const int NLP = p.kpack ? (p.NL + 1) / 2 : p.NL, tstep = p.kpack ? 2 : 1;
        // descriptors differ only in their 14-bit start-address field: build once, then add (bytes >> 4)
        const uint64_t a_hi0 = make_desc(Asm_addr, a_lbo, 128), a_lo0 = make_desc(Asm_addr + nkc * a_plane, a_lbo, 128);
        const uint64_t b_hi0 = make_desc(Wsm_addr, 16u * p.N, 128), b_lo0 = make_desc(Wsm_addr + w_pass, 16u * p.N, 128);
        for (int mb = warp; mb < p.MB; mb += 4) {
          const uint32_t d = tmem_base + (uint32_t)(mb * p.N);
          uint32_t t = 0;
          for (int tk = 0; tk < p.NK; tk++) {
            const uint32_t a_row = (uint32_t)(mb * 128 + tk * p.PJ);
            for (int tp = 0; tp < NLP; tp++, t++) {
              const uint64_t a_add = (uint64_t)(a_row + tp * tstep), b_add = (uint64_t)(t * (w_tap >> 4));
              mma_bf16(d, a_hi0 + a_add, b_hi0 + b_add, idesc, !(ks == 0 && t == 0));
              if (p.passes == 3) {
                mma_bf16(d, a_hi0 + a_add, b_lo0 + b_add, idesc, true);
                mma_bf16(d, a_lo0 + a_add, b_hi0 + b_add, idesc, true);
              }
            }
          }
        }
        commit(bar);
      }
      // ---- everyone waits for the MMAs: shared memory may be restaged, TMEM may be read ----
      mbar_wait(bar, phase);
      phase ^= 1;
      fence_after_sync();
    }
    // ---- epilogue: TMEM lane = pixel, column = out channel; warps w and w+4 share lanes 32*(w%4).. and split columns ----
    {
      const int lane_grp = warp & 3, col_half = warp >> 2;
      const int rows_valid = min(p.TI, p.Nx - i0);
      float* ob = p.out + b * p.O * plane;
      for (int mb = 0; mb < p.MB; mb++) {
        const int q = mb * 128 + lane_grp * 32 + lane;
        const int r = q / p.PJ, col = q - r * p.PJ;
        const bool valid = r < rows_valid && col < p.Ny;
        float* dst = ob + (long long)(i0 + r) * p.Ny + col;
        for (int n0 = col_half * 16; n0 < p.N; n0 += 32) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(mb * p.N + n0), v);
          if (valid) {
#pragma unroll
            for (int e = 0; e < 16; e++) {
              const int o = n0 + e;
              if (o < p.O) dst[(long long)o * plane] = v[e] + (p.bias ? __ldg(p.bias + o) : 0.f);
            }
          }
        }
      }
    }
    fence_before_sync();
    __syncthreads();  // TMEM fully read before the next tile's first MMA overwrites it
    fence_after_sync();
  }
  if (warp == 0) tmem_dealloc(tmem_base, p.tmem_cols);
}

// host wrapper shared with conv_tc_ws.cu
void conv_weight_prep(aefft_ctx* ctx, const float* w, long long w_so, long long w_sc, int C, int O, int N, int NK, int NL,
                      int flip, int KS, int kpack, void* wprep, long long total) {
  weight_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(w, w_so, w_sc, C, O, N, NK, NL, flip, KS, kpack,
                                                                            reinterpret_cast<__nv_bfloat16*>(wprep));
  ctx->launches++;
}

static uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

// Returns AEFFT_ERR_UNSUPPORTED (without setting an error) when the shape is outside the tensor-core kernel's envelope.
int launch_conv_tc(aefft_ctx* ctx, const Window& win, int64_t B, int C, int O, int Nx, int Ny, const float* src0,
                   const float* src1, float pre_div, const float* w, int64_t w_so, int64_t w_sc, const float* bias,
                   float* out, int passes) {
  const int N = (O + 15) / 16 * 16;
  const int kpack = C <= 8 ? 1 : 0;
  const int TE = kpack ? win.Nk * ((win.Nl + 1) / 2) : win.Nk * win.Nl;
  const int nkc = kpack ? 1 : 2;
  const int PJ = Ny + win.Nl - 1;
  if (N > 256 || TE > 64 || PJ > 4096) return AEFFT_ERR_UNSUPPORTED;
  const int w_bytes = 64 * TE * N;
  const int halo = (win.Nk - 1) * PJ + win.Nl + 8;
  // two co-resident CTAs per SM (one stages / stores while the other's MMAs run) when the tile still holds >= 2
  // M-blocks; otherwise one CTA with the whole SM
  int ctas = 2, TI = 0, mb_max = 0;
  for (; ctas >= 1; ctas--) {
    const int smem_budget = ctas == 2 ? 112 * 1024 : 224 * 1024;
    mb_max = (ctas == 2 ? 256 : 512) / N;
    const int mb_smem = ((smem_budget - w_bytes - 64) / (32 * nkc) - halo) / 128;
    if (mb_smem < mb_max) mb_max = mb_smem;
    TI = mb_max >= 1 ? mb_max * 128 / PJ : 0;
    if (TI >= 1 && (ctas == 1 || mb_max >= 2)) break;
  }
  if (ctas < 1 || TI < 1) return AEFFT_ERR_UNSUPPORTED;
  if (TI > Nx) TI = Nx;
  // enough tiles to occupy every SM
  while (TI > 1 && (long long)B * ((Nx + TI - 1) / TI) < (long long)ctas * ctx->sm_count) TI = (TI + 1) / 2;
  ConvTcParams p;
  p.src0 = src0; p.src1 = src1; p.bias = bias; p.out = out; p.pre_div = pre_div;
  p.C = C; p.O = O; p.N = N; p.Nx = Nx; p.Ny = Ny;
  p.NK = win.Nk; p.NL = win.Nl; p.T = win.Nk * win.Nl;
  p.ai0 = win.ai0; p.aj0 = win.aj0; p.lo = win.lo;
  p.PJ = PJ; p.TI = TI;
  p.MB = (TI * PJ + 127) / 128;
  p.HP = (p.MB * 128 + halo + 7) / 8 * 8;
  p.KS = kpack ? 1 : (C + KC - 1) / KC;
  p.kpack = kpack; p.TE = TE;
  p.tiles_per_frame = (Nx + TI - 1) / TI;
  p.n_tiles = (long long)B * p.tiles_per_frame;
  p.passes = passes;
  p.tmem_cols = pow2_cols(p.MB * N);
  if (p.tmem_cols > (ctas == 2 ? 256u : 512u) || (size_t)p.HP * 16 > 262000) return AEFFT_ERR_UNSUPPORTED;
  const size_t smem = (size_t)w_bytes + 32 * (size_t)nkc * p.HP + 64;
  if (smem > (size_t)(ctas == 2 ? 113 : 227) * 1024) return AEFFT_ERR_UNSUPPORTED;
  // weights -> bf16 hi/lo shared-memory images
  __nv_bfloat16* wprep;
  AE_TRY(ctx->getT(win.flip ? "tc_wprep_f" : "tc_wprep_t", (size_t)p.KS * 2 * TE * 2 * N * 8, &wprep));
  {
    const long long total = (long long)p.KS * TE * 2 * N * 8;
    weight_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(w, w_so, w_sc, C, O, N, win.Nk, win.Nl, win.flip,
                                                                              p.KS, kpack, wprep);
    ctx->launches++;
  }
  p.wprep = reinterpret_cast<const uint4*>(wprep);
  AE_TRY(ctx->ensure_dyn_smem((const void*)conv_tc_kernel, 227 * 1024));
  const double px = (double)B * Nx * Ny;
  ProfScope prof(ctx, win.flip ? "conv_fwd_tc" : "conv_dgrad_tc", 2.0 * px * C * O * p.T,
                 4.0 * (px * C * (src1 ? 2 : 1) + px * O + (double)C * O * p.T));
  const long long slots = (long long)ctas * ctx->sm_count;
  const unsigned grid = (unsigned)(p.n_tiles < slots ? p.n_tiles : slots);
  conv_tc_kernel<<<grid, TC_THREADS, smem, ctx->stream>>>(p);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

}  // namespace aefft
