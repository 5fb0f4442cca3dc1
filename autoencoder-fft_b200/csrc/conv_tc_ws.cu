// Warp-specialised, software-pipelined variant of the tcgen05 implicit-GEMM convolution (conv_tc.cu has the operand
// layouts and the numerics; this file only changes the schedule).  One persistent CTA per SM, 16 warps:
//   warps 0-7   PRODUCERS : stage the input tile (+halo) of (tile, K-stage) s+1 as bf16 hi/lo planes into a 2-deep ring
//   warps 8-11  MMA       : lane 0 of each issues the tcgen05.mma of its M-blocks, commits to the ring's `empty` barrier
//                           (and to `acc_full` after the tile's last K stage)
//   warps 12-15 EPILOGUE  : tcgen05.ld of the finished accumulator buffer, bias, coalesced stores, `acc_empty`
// so global loads/conversion, tensor-core math and the output stores of three consecutive tiles overlap.  TMEM holds
// two accumulator buffers (2 x MB x N columns); all weights of all K stages stay resident in shared memory (this
// variant is selected only when they fit).  Tiles are TI rows x TJ columns of one frame; an M-block is 128
// consecutive linear pixels of the (TJ + NL - 1)-wide halo grid.
#include <cstdlib>

#include "common.cuh"
#include "staging.cuh"
#include "umma.cuh"

namespace aefft {

using namespace umma;

constexpr int WS_PROD_WARPS = 8, WS_MMA_WARPS = 4, WS_EPI_WARPS = 4;
constexpr int WS_THREADS = 32 * (WS_PROD_WARPS + WS_MMA_WARPS + WS_EPI_WARPS);
constexpr int WS_KC = 16;

struct ConvWsParams {
  const float* src0;
  const float* src1;
  const uint4* wprep;  // [KS][2 pass][TE][2][N][8 bf16]  (weight_prep_kernel of conv_tc.cu)
  const float* bias;
  float* out;
  float pre_div;
  int C, O, N, Nx, Ny;
  int NK, NL;
  int ai0, aj0, lo;
  int PJ, TI, TJ, MB, HP, KS;
  int tiles_i, tiles_j;
  long long n_tiles;
  int passes, kpack, TE;
  uint32_t tmem_cols;
  int dbg;  // AEFFT_DEBUG_SKIP bitmask (profiling experiments only): 1 skip staging, 2 skip MMAs, 4 skip stores
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(WS_THREADS, 1) conv_tc_ws_kernel(ConvWsParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nkc = p.kpack ? 1 : 2;
  const int w_stage = 64 * p.TE * p.N;          // bytes of one K stage of weights (both passes)
  const int a_plane = p.HP * 16;
  const int a_stage = 2 * nkc * a_plane;        // bytes of one staged input tile (both passes)
  unsigned char* Wsm = smem;
  unsigned char* Asm = smem + (size_t)p.KS * w_stage;  // ring of 2
  __shared__ __align__(8) uint64_t full[2], empty[2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_slot, p.tmem_cols);
  if (tid == 32) {
    for (int i = 0; i < 2; i++) {
      mbar_init(&full[i], WS_PROD_WARPS);
      mbar_init(&empty[i], WS_MMA_WARPS);
      mbar_init(&acc_full[i], WS_MMA_WARPS);
      mbar_init(&acc_empty[i], WS_EPI_WARPS);
    }
    fence_mbar_init();
  }
  // all weights, once
  {
    const uint4* src = p.wprep;
    uint4* dst = reinterpret_cast<uint4*>(Wsm);
    const int n16 = p.KS * w_stage / 16;
    for (int i = tid; i < n16; i += WS_THREADS) dst[i] = __ldg(src + i);
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  const long long plane = (long long)p.Nx * p.Ny;
  const int tiles_per_frame = p.tiles_i * p.tiles_j;
  const int acc_cols = p.MB * p.N;

  if (warp < WS_PROD_WARPS) {
    // ================================================================== PRODUCERS
    const int ptid = tid;  // 0..255
    const int HI = p.TI + p.NK - 1;
    uint32_t stage = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const long long b = tile / tiles_per_frame;
      const int tr = (int)(tile % tiles_per_frame);
      const int i0 = (tr / p.tiles_j) * p.TI, j0 = (tr % p.tiles_j) * p.TJ;
      const float* s0 = p.src0 + b * p.C * plane;
      const float* s1 = p.src1 ? p.src1 + b * p.C * plane : nullptr;
      for (int ks = 0; ks < p.KS; ks++, stage++) {
        const int slot = stage & 1;
        mbar_wait(&empty[slot], ((stage >> 1) & 1) ^ 1);  // first use of each slot passes immediately
        unsigned char* A = Asm + (size_t)slot * a_stage;
        if (!(p.dbg & 1))
        stage_planes4<32 * WS_PROD_WARPS>(A, A + (size_t)nkc * a_plane, (uint32_t)a_plane, nkc, p.HP, p.PJ, s0, s1,
                                          (!s1 && p.pre_div != 0.f) ? 1.f / p.pre_div : 1.f, ks * WS_KC, p.C - ks * WS_KC, plane,
                                          p.Nx, p.Ny, i0 + p.ai0, j0 + p.aj0, HI, p.PJ, p.lo, ptid);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[slot]);
      }
    }
  } else if (warp < WS_PROD_WARPS + WS_MMA_WARPS) {
    // ================================================================== MMA issuers
    // Every lane runs the loops so that the descriptor arithmetic stays warp-uniform (uniform datapath, no per-thread
    // register -> uniform-register moves in front of each UTCHMMA); only the tcgen05 instructions are predicated.
    const int mw = warp - WS_PROD_WARPS;
    const bool leader = lane == 0;
    const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
    const uint32_t w_pass = (uint32_t)w_stage / 2, w_tap16 = 2u * p.N;  // 16-byte units
    const uint32_t a_lbo = p.kpack ? 16u : (uint32_t)a_plane;
    const int NLP = p.kpack ? (p.NL + 1) / 2 : p.NL, tstep = p.kpack ? 2 : 1;
    uint32_t stage = 0, tcount = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, tcount++) {
      const int abuf = tcount & 1;
      mbar_wait(&acc_empty[abuf], ((tcount >> 1) & 1) ^ 1);  // epilogue has drained this accumulator buffer
      fence_after_sync();
      for (int ks = 0; ks < p.KS; ks++, stage++) {
        const int slot = stage & 1;
        mbar_wait(&full[slot], (stage >> 1) & 1);
        fence_after_sync();
        const uint32_t A_addr = smem_u32(Asm + (size_t)slot * a_stage), W_addr = smem_u32(Wsm + (size_t)ks * w_stage);
        const uint64_t a_hi0 = make_desc(A_addr, a_lbo, 128), a_lo0 = make_desc(A_addr + nkc * a_plane, a_lbo, 128);
        const uint64_t b_hi0 = make_desc(W_addr, 16u * p.N, 128), b_lo0 = make_desc(W_addr + w_pass, 16u * p.N, 128);
        if (!(p.dbg & 2)) {
          for (int mb = mw; mb < p.MB; mb += WS_MMA_WARPS) {
            const uint32_t d = tmem_base + (uint32_t)(abuf * acc_cols + mb * p.N);
            uint32_t t = 0;
            for (int tk = 0; tk < p.NK; tk++) {
              const uint32_t a_row = (uint32_t)(mb * 128 + tk * p.PJ);
              for (int tp = 0; tp < NLP; tp++, t++) {
                const uint64_t a_add = (uint64_t)(a_row + tp * tstep), b_add = (uint64_t)(t * w_tap16);
                if (leader) {
                  mma_bf16(d, a_hi0 + a_add, b_hi0 + b_add, idesc, !(ks == 0 && t == 0));
                  if (p.passes == 3) {
                    mma_bf16(d, a_hi0 + a_add, b_lo0 + b_add, idesc, true);
                    mma_bf16(d, a_lo0 + a_add, b_hi0 + b_add, idesc, true);
                  }
                }
              }
            }
          }
        }
        if (leader) {
          commit(&empty[slot]);                         // ring slot reusable once these MMAs have read it
          if (ks == p.KS - 1) commit(&acc_full[abuf]);  // accumulators of this tile complete
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================== EPILOGUE
    const int lane_grp = warp & 3;  // TMEM lanes 32*lane_grp .. +31
    uint32_t tcount = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, tcount++) {
      const int abuf = tcount & 1;
      const long long b = tile / tiles_per_frame;
      const int tr = (int)(tile % tiles_per_frame);
      const int i0 = (tr / p.tiles_j) * p.TI, j0 = (tr % p.tiles_j) * p.TJ;
      const int rows_valid = min(p.TI, p.Nx - i0), cols_valid = min(p.TJ, p.Ny - j0);
      float* ob = p.out + b * p.O * plane;
      mbar_wait(&acc_full[abuf], (tcount >> 1) & 1);
      fence_after_sync();
      for (int mb = 0; mb < p.MB; mb++) {
        const int q = mb * 128 + lane_grp * 32 + lane;
        const int r = q / p.PJ, col = q - r * p.PJ;
        const bool valid = r < rows_valid && col < cols_valid;
        float* dst = ob + (long long)(i0 + r) * p.Ny + j0 + col;
        for (int n0 = 0; n0 < p.N; n0 += 16) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(abuf * acc_cols + mb * p.N + n0), v);
          if (valid && !(p.dbg & 4)) {
#pragma unroll
            for (int e = 0; e < 16; e++) {
              const int o = n0 + e;
              if (o < p.O) dst[(long long)o * plane] = v[e] + (p.bias ? __ldg(p.bias + o) : 0.f);
            }
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[abuf]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, p.tmem_cols);
}

void conv_weight_prep(aefft_ctx* ctx, const float* w, long long w_so, long long w_sc, int C, int O, int N, int NK, int NL,
                      int flip, int KS, int kpack, void* wprep, long long total);

// Returns AEFFT_ERR_UNSUPPORTED when the weights of all K stages do not fit next to a 2-deep input ring.
int launch_conv_tc_ws(aefft_ctx* ctx, const Window& win, int64_t B, int C, int O, int Nx, int Ny, const float* src0,
                      const float* src1, float pre_div, const float* w, int64_t w_so, int64_t w_sc, const float* bias,
                      float* out, int passes) {
  const int N = (O + 15) / 16 * 16;
  const int kpack = C <= 8 ? 1 : 0;
  const int TE = kpack ? win.Nk * ((win.Nl + 1) / 2) : win.Nk * win.Nl;
  const int nkc = kpack ? 1 : 2;
  const int KS = kpack ? 1 : (C + WS_KC - 1) / WS_KC;
  if (N > 128 || TE > 64) return AEFFT_ERR_UNSUPPORTED;
  const size_t w_total = (size_t)KS * 64 * TE * N;
  const size_t budget = 222 * 1024;
  if (w_total + 2 * 32 * 1024 > budget) return AEFFT_ERR_UNSUPPORTED;
  const int mb_tmem = 256 / N;  // two accumulator buffers in 512 columns
  // choose the tile: TJ columns (whole rows when they are short), TI rows, maximising useful pixels per staged pixel
  int best_TI = 0, best_TJ = 0, best_MB = 0, best_HP = 0;
  double best_eff = 0.0;
  for (int TJ = Ny < 32 ? Ny : 32; TJ <= Ny; TJ = (TJ == Ny ? Ny + 1 : (TJ + 32 > Ny ? Ny : TJ + 32))) {
    const int PJ = TJ + win.Nl - 1;
    const int halo = (win.Nk - 1) * PJ + win.Nl + 8;
    for (int MB = mb_tmem; MB >= 1; MB--) {
      const int HP = (MB * 128 + halo + 7) / 8 * 8;
      const size_t a_stage = (size_t)2 * nkc * HP * 16;
      if (w_total + 2 * a_stage > budget || (size_t)HP * 16 > 262000) continue;
      int TI = MB * 128 / PJ;
      if (TI < 1) break;
      if (TI > Nx) TI = Nx;
      const int tj_n = (Ny + TJ - 1) / TJ, ti_n = (Nx + TI - 1) / TI;
      // useful output pixels / (staged pixels + MMA pixels), tail tiles included
      const double eff = (double)Nx * Ny / ((double)tj_n * ti_n * (HP + MB * 128));
      if (eff > best_eff) { best_eff = eff; best_TI = TI; best_TJ = TJ; best_MB = (TI * PJ + 127) / 128; best_HP = ((TI * PJ + 127) / 128 * 128 + halo + 7) / 8 * 8; }
      break;  // largest MB that fits for this TJ
    }
  }
  if (best_TI < 1) return AEFFT_ERR_UNSUPPORTED;
  ConvWsParams p;
  p.src0 = src0; p.src1 = src1; p.bias = bias; p.out = out; p.pre_div = pre_div;
  p.C = C; p.O = O; p.N = N; p.Nx = Nx; p.Ny = Ny; p.NK = win.Nk; p.NL = win.Nl;
  p.ai0 = win.ai0; p.aj0 = win.aj0; p.lo = win.lo;
  p.TI = best_TI; p.TJ = best_TJ; p.PJ = best_TJ + win.Nl - 1; p.MB = best_MB; p.HP = best_HP; p.KS = KS;
  p.tiles_i = (Nx + p.TI - 1) / p.TI; p.tiles_j = (Ny + p.TJ - 1) / p.TJ;
  p.n_tiles = (long long)B * p.tiles_i * p.tiles_j;
  p.passes = passes; p.kpack = kpack; p.TE = TE;
  {
    const char* e = getenv("AEFFT_DEBUG_SKIP");
    p.dbg = e ? atoi(e) : 0;
  }
  p.tmem_cols = 32;
  while ((int)p.tmem_cols < 2 * p.MB * N) p.tmem_cols <<= 1;
  if (p.tmem_cols > 512) return AEFFT_ERR_UNSUPPORTED;
  const size_t smem = w_total + 2 * (size_t)2 * nkc * p.HP * 16;
  if (smem > budget) return AEFFT_ERR_UNSUPPORTED;
  void* wprep;
  const long long total = (long long)KS * TE * 2 * N * 8;
  AE_TRY(ctx->get(win.flip ? "tc_wprep_f" : "tc_wprep_t", (size_t)total * 2 * sizeof(__nv_bfloat16), &wprep));
  conv_weight_prep(ctx, w, w_so, w_sc, C, O, N, win.Nk, win.Nl, win.flip, KS, kpack, wprep, total);
  p.wprep = reinterpret_cast<const uint4*>(wprep);
  AE_TRY(ctx->ensure_dyn_smem((const void*)conv_tc_ws_kernel, smem));
  const double px = (double)B * Nx * Ny;
  ProfScope prof(ctx, win.flip ? "conv_fwd_tc" : "conv_dgrad_tc", 2.0 * px * C * O * win.Nk * win.Nl,
                 4.0 * (px * C * (src1 ? 2 : 1) + px * O + (double)C * O * win.Nk * win.Nl));
  const unsigned grid = (unsigned)(p.n_tiles < ctx->sm_count ? p.n_tiles : ctx->sm_count);
  conv_tc_ws_kernel<<<grid, WS_THREADS, smem, ctx->stream>>>(p);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

}  // namespace aefft
