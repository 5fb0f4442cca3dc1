// Row-streaming tcgen05 convolution (forward conv, its transpose = data gradient with e = out - in fused at load).
//
//   out[b][o](i,j) = bias[o] + sum_{c,tk,tl} W(o,c,tk,tl) * S[b][c](i+ai0+tk, j+aj0+tl)        (contract of launch_conv)
//
// conv_tc.cu / conv_tc_ws.cu issue one M=128 x N=O x K=16 MMA per tap; with O = 16..64 such an MMA is bound by the
// 128 B/clk shared-memory operand path (4 KB of A per 8..32 cycles of math), measured 3-4.5x below the tensor pipe.
// Here the window ROWS are stacked along N instead:
//   * an M-block is one input row of a 128-pixel strip (lanes = columns; two 64-pixel strips of different frames share a
//     block when the image is narrow), K = 16 input channels (or two taps x 8 channels when C <= 8);
//   * for window column tl the A operand is that row advanced by tl pixels (descriptor start + tl*16 B, the [pixel][8 ch]
//     SWIZZLE_NONE K-major planes of conv_tc.cu);
//   * the B operand holds the weights of ALL NK window rows for that column, [tk descending][o], so N = NK*O: input row k
//     adds its contribution to the NK output rows k-NK+1..k AT ONCE, because their accumulators are adjacent column
//     ranges of a TMEM ring (output row rho lives at columns (rho mod Rr)*O).  One A read now feeds NK x more math.
//   * rows stream: TMA -> fp32 row ring -> converter warps (bf16 hi/lo split, e = out - in, 1/dM scale) -> bf16 row
//     ring -> MMA issuer -> epilogue warps drain finished output rows (tcgen05.ld, bias, coalesced stores) and hand the
//     zeroed accumulator slot back.  All MMAs accumulate; the halo is paid once per band.
// fp32 parity: BF16X3 (A_hi W_hi + A_hi W_lo + A_lo W_hi, fp32 accumulation) as in conv_tc.cu.
// Roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-7 converters, warps 8-11 epilogue.
// Per layer shape launch_conv_rs also picks (each choice is documented where it is made, and forced in the tests):
//   * 1, 2 or 3 CTAs per SM (512 / 256 / 128 TMEM columns, a share of the shared memory each);
//   * the K packing: 16 channels, two window columns x 8 channels (C <= 8), or (window column, channel) when a whole
//     window row fits in K = 16 (the 3-channel input layer: one MMA triple per input row);
//   * the N packing: output channels padded to 16 per window row, or (window column, output) for few-output layers (the
//     3-channel reconstruction layer), whose column groups the epilogue adds across lanes through shared memory;
//   * one or two input rows per pipeline unit (barrier round trip of producer, converters and issuer).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "pipe.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace aefft {

using namespace umma;
using namespace tma;

constexpr int RS_THREADS = 384;
constexpr int RS_NSF = 2;      // fp32 staging slots
constexpr int RS_MAXSB = 8;    // bf16 row ring slots (max)
constexpr int RS_MAXACC = 32;  // accumulator ring slots (max)
constexpr int RS_CONV_WARPS = 6;

struct ConvRsParams {
  CUtensorMap x0_map, x1_map;  // [B*C][Nx][Ny] fp32, box {PJs+4, 1, C}
  const uint4* wprep;          // [job][KS][part][NLg][kchunk][Ntot][8 bf16]
  const float* bias;
  float* out;
  float scale;
  int has_x1;
  int C, O, Nx, Ny, NK, NL, ai0, aj0;
  int Oj, O_pad, n_jobs;
  int PJs, G, TJ, strips, bands, BR;
  int n_sub, items, cpj;
  int KS, NP, kpack, NLg, Ntot, Rr, NSB, NSF, passes;
  int RU, NSBR;    // input rows per pipeline unit (one barrier round trip moves RU rows), rows of the bf16 ring = NSB * RU
  int tmem_cols;   // 512 / 256 / 128 columns for 1 / 2 / 3 CTAs per SM
  int npack;       // few outputs (O * NL <= 16): N = (window row, window column, output), the epilogue sums the columns
  uint32_t off_xch;
  int stack2, CW;  // stack2: B rows = [W_hi | W_lo] per window row (2 MMAs instead of 3); CW = accumulator columns per output row
  uint32_t w_bytes, seg_bytes, src_bytes, x_slot_bytes, sb_pitch;
  uint32_t off_w, off_x, off_sb;
  long long* dbg;
  int skip;  // instrumented build only (AEFFT_RS_DEBUG=1 AEFFT_RS_SKIP=mask): knock out 1 converter work, 2 MMAs, 4 epilogue stores
};

__global__ void conv_rs_weight_prep_kernel(const float* __restrict__ w, long long w_so, long long w_sc, int C, int O, int Oj,
                                           int O_pad, int NK, int NL, int NLg, int flip, int KS, int kpack, int n_jobs,
                                           int stack2, int npack, __nv_bfloat16* __restrict__ wprep) {
  const int Ntot = NK * O_pad;
  const long long per_part = (long long)NLg * 2 * Ntot * 8;
  const long long per_job = (long long)KS * 2 * per_part;
  const long long total = (long long)n_jobs * KS * per_part;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx;
    const int e = (int)(r % 8); r /= 8;
    const int n = (int)(r % Ntot); r /= Ntot;
    const int kchunk = (int)(r % 2); r /= 2;
    const int tlg = (int)(r % NLg); r /= NLg;
    const int ks = (int)(r % KS); r /= KS;
    const int job = (int)r;
    const int tk = NK - 1 - n / O_pad;
    // N index inside the window row: output channel, or (window column, output) packed when npack (the A operand is then
    // not advanced per window column; the epilogue adds the NL column groups with their pixel shifts)
    const int q = n % O_pad;
    const int o = npack ? (q < O * NL ? q % O : O_pad) : q;
    // K index of this element: channel block (kpack 0), tap pair x 8 channels (kpack 1), or (tap, channel) packed (kpack 2)
    const int kk = kchunk * 8 + e;
    const int tl = npack ? (q < O * NL ? q / O : NL) : kpack == 2 ? (kk < C * NL ? kk / C : NL) : kpack ? 2 * tlg + kchunk : tlg;
    const int c = kpack == 2 ? kk % C : kpack ? e : ks * 16 + kchunk * 8 + e;
    const int k = flip ? NK - 1 - tk : tk, l = flip ? NL - 1 - tl : tl;
    const int og = job * Oj + o;
    float v = 0.f;
    if (c < C && o < Oj && og < O && tl < NL) v = w[og * w_so + c * w_sc + k * NL + l];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    if (stack2) {
      // rows [tk descending][hi | lo][o]: one B operand yields A*W_hi and A*W_lo in adjacent accumulator columns
      const int tkd = n / O_pad;
      const long long row_hi = (long long)(tkd * 2) * O_pad + o, rows = 2LL * Ntot;
      const long long base = (long long)job * per_job + (long long)ks * 2 * per_part + (((long long)tlg * 2 + kchunk) * rows) * 8 + e;
      wprep[base + row_hi * 8] = hi;
      wprep[base + (row_hi + O_pad) * 8] = lo;
    } else {
      const long long within = (((long long)tlg * 2 + kchunk) * Ntot + n) * 8 + e;
      const long long base = (long long)job * per_job + (long long)ks * 2 * per_part + within;
      wprep[base] = hi;
      wprep[base + per_part] = lo;
    }
  }
}

template <bool DBG>
__global__ void __launch_bounds__(RS_THREADS, 3) conv_rs_kernel(const __grid_constant__ ConvRsParams p) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_full[RS_NSF], s_empty[RS_NSF], xb_full[RS_MAXSB], xb_empty[RS_MAXSB],
      acc_full[RS_MAXACC], acc_empty[RS_MAXACC];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[256];
  __shared__ int voff_s[16];  // kpack 2: offset of K index (tl, c) inside the fp32 row box, -1 = padding  // bias of this job's outputs (0 beyond O or without bias)

  // warp index via a broadcast shuffle: the compiler then knows it is warp-uniform and keeps the role loops (MMA
  // descriptors, ring positions) in uniform registers instead of moving them there lane by lane
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int job = blockIdx.y, cta = blockIdx.x;
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* w_sm = smem + p.off_w;
  unsigned char* x_ring = smem + p.off_x;
  unsigned char* sb_ring = smem + p.off_sb;

  if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  if (tid == 32) {
    for (int i = 0; i < p.NSF; i++) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], RS_CONV_WARPS); }
    for (int i = 0; i < p.NSB; i++) { mbar_init(&xb_full[i], RS_CONV_WARPS); mbar_init(&xb_empty[i], 1); }
    for (int i = 0; i < p.Rr; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_mbar_init();
  }
  // weights of this job (all K stages, both parts) and a zeroed bf16 ring (pad pixels are read and must be finite)
  {
    const uint4* src = p.wprep + (size_t)job * (p.w_bytes / 16);
    uint4* dst = reinterpret_cast<uint4*>(w_sm);
    for (uint32_t i = tid; i < p.w_bytes / 16; i += RS_THREADS) dst[i] = __ldg(src + i);
    const uint32_t n16 = (uint32_t)(2 * p.NP * p.NSBR) * p.sb_pitch / 16;
    uint4* z = reinterpret_cast<uint4*>(sb_ring);
    for (uint32_t i = tid; i < n16; i += RS_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    if (tid < 16) voff_s[tid] = tid < p.C * p.NL ? (tid % p.C) * (p.PJs + 4) + tid / p.C : -1;
    if (tid < 256) bias_s[tid] = (p.bias && tid < p.Oj && job * p.Oj + tid < p.O) ? __ldg(p.bias + job * p.Oj + tid) : 0.f;
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_slot;
  if (warp >= 8) {
    uint32_t z[16];
#pragma unroll
    for (int e = 0; e < 16; e++) z[e] = 0u;
    for (int c0 = 0; c0 < p.Rr * p.CW; c0 += 16) tmem_st16(tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, z);
    tmem_wait_st();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  long long wA = 0, wB = 0, wC = 0;
  const long long t_start = DBG ? clock64() : 0;

  const int n_items = p.items, cpj = p.cpj;
  const int G = p.G, PJs = p.PJs, NK = p.NK;
  const int SP = PJs + 4;
  const int cs_off = p.aj0 & ~3, d_off = p.aj0 - (p.aj0 & ~3);

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&p.x0_map);
      if (p.has_x1) tma_prefetch_desc(&p.x1_map);
      Ring ss(p.NSF);
      const uint32_t seg_tx = (uint32_t)p.C * SP * 4;
      for (int item = cta; item < n_items; item += cpj) {
        const int ug = item / p.bands, band = item - ug * p.bands;
        const int i0 = band * p.BR;
        const int nrows = min(p.BR, p.Nx - i0);
        const int n_in = nrows + NK - 1;
        const int nseg = min(G, p.n_sub - ug * G);
        // per item: the (column, plane) coordinates of the one or two strips of the M-block; per row only the wait, the
        // transaction count and the TMA issues remain on this single thread's path (it is the slowest role of the
        // pipeline skeleton: 780 cycles per row with the divisions inside the row loop)
        int cj[2], cb[2];
        for (int g = 0; g < 2; g++) {
          const int u = ug * G + (g < nseg ? g : 0);
          const int b = u / p.strips;
          cj[g] = (u - b * p.strips) * p.TJ + cs_off;
          cb[g] = b * p.C;
        }
        const uint32_t tx = seg_tx * nseg * (p.has_x1 ? 2 : 1);
        const int row0 = i0 + p.ai0;
        for (int k0 = 0; k0 < n_in; k0 += p.RU) {
          const int nr = min(p.RU, n_in - k0);  // rows of this unit
          wait_t<DBG, true>(&s_empty[ss.slot], ss.phase ^ 1, wA);
          mbar_expect_tx(&s_full[ss.slot], tx * nr);
          for (int r = 0; r < nr; r++) {
            unsigned char* dst = x_ring + (size_t)(ss.slot * p.RU + r) * p.x_slot_bytes;
            const int k = k0 + r;
            tma_load_3d(dst, &p.x0_map, cj[0], row0 + k, cb[0], &s_full[ss.slot]);
            if (p.has_x1) tma_load_3d(dst + p.src_bytes, &p.x1_map, cj[0], row0 + k, cb[0], &s_full[ss.slot]);
            if (nseg > 1) {
              tma_load_3d(dst + p.seg_bytes, &p.x0_map, cj[1], row0 + k, cb[1], &s_full[ss.slot]);
              if (p.has_x1)
                tma_load_3d(dst + p.src_bytes + p.seg_bytes, &p.x1_map, cj[1], row0 + k, cb[1], &s_full[ss.slot]);
            }
          }
          ss.next();
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (whole warp runs the loops, one lane issues)
    const uint32_t sb_base = smem_u32(sb_ring), w_base = smem_u32(w_sm);
    const uint32_t pitch = p.sb_pitch;
    const int NSB = p.NSB, NSBR = p.NSBR, NP = p.NP, KS = p.KS, NLg = p.NLg, Rr = p.Rr, CW = p.CW;
    const bool stack2 = p.stack2 != 0;
    const int Ntot = stack2 ? 2 * p.Ntot : p.Ntot;  // B rows per K chunk
    const uint32_t a_lbo = p.kpack == 1 ? 16u : (uint32_t)NSBR * pitch;
    const uint64_t a_desc0 = make_desc(0, a_lbo, 128), b_desc0 = make_desc(0, (uint32_t)Ntot * 16, 128);
    const uint32_t part_off16 = ((uint32_t)(NP * NSBR) * pitch) >> 4;               // A: hi -> lo part
    const uint32_t wpart_off16 = (uint32_t)(NLg * 2 * Ntot);                       // W: hi -> lo part (16-byte units; !stack2)
    const bool three = p.passes == 3;
    const int maxchunks = 256 / CW;
    const uint32_t a_ks_step16 = p.kpack == 1 ? 0u : (((uint32_t)(2 * NSBR) * pitch) >> 4);
    const uint32_t a_tl_step16 = p.kpack == 1 ? 2u : 1u;
    const uint32_t w_ks_step16 = (uint32_t)((stack2 ? 1 : 2) * NLg * 2 * Ntot), w_tl_step16 = (uint32_t)(2 * Ntot);
    Ring rx(NSB);
    Ring rn(Rr);  // accumulator ring position of the newest output row (rho = k)
    int gro = 0;  // accumulator slot of output row 0 of the current item
    for (int item = cta; item < n_items; item += cpj) {
      const int band = item % p.bands;
      const int i0 = band * p.BR;
      const int nrows = min(p.BR, p.Nx - i0);
      const int n_in = nrows + NK - 1;
      int fslot = gro;
      for (int k0 = 0; k0 < n_in; k0 += p.RU) {
       // one barrier round trip per unit of RU input rows
       wait_t<DBG, true>(&xb_full[rx.slot], rx.phase, wA);
       for (int r = 0; r < p.RU && k0 + r < n_in; r++) {
        const int k = k0 + r;
        if (k < nrows) {
          // the accumulator slot of the newest output row (rho = k) must have been drained and zeroed
          wait_t<DBG, true>(&acc_empty[rn.slot], rn.phase ^ 1, wB);
          rn.next();
        }
        fence_after_sync();
        const long long t_m0 = DBG ? clock64() : 0;
        // The stack = output rows rho_lo..rho_hi (ascending) = window rows tk_hi..tk_lo, one MMA group per contiguous piece
        // of accumulator columns (split at the ring wrap and at N = 256).  fslot = accumulator slot of the oldest row of the
        // stack, advanced as rows complete.  One elected lane walks pieces and descriptors with constant increments; every
        // accumulator column still receives its MMAs in (K stage, window column, pass) order.
        const int tk_lo = max(0, k - nrows + 1), tk_hi = min(NK - 1, k);
        const uint32_t a_row16 = (sb_base + (uint32_t)(rx.slot * p.RU + r) * pitch) >> 4;
        if (elect_one() && !(DBG && (p.skip & 2))) {
          int rem = tk_hi - tk_lo + 1, slot = fslot;
          uint32_t n0 = (uint32_t)((NK - 1 - tk_hi) * CW);
#pragma unroll 1
          while (rem > 0) {
            const int len = min(rem, min(Rr - slot, maxchunks));
            const uint32_t idesc = make_idesc_bf16(128, len * CW, 0, 0);
            const uint32_t d = tb + (uint32_t)(slot * CW);
            uint64_t a_ks = a_desc0 + (uint64_t)a_row16;
            uint64_t b_ks = b_desc0 + (uint64_t)((w_base >> 4) + n0);
#pragma unroll 1
            for (int ks = 0; ks < KS; ks++, a_ks += a_ks_step16, b_ks += w_ks_step16) {
              uint64_t a_hi = a_ks, b_hi = b_ks;
#pragma unroll 1
              for (int tlg = 0; tlg < NLg; tlg++, a_hi += a_tl_step16, b_hi += w_tl_step16) {
                mma_bf16(d, a_hi, b_hi, idesc, true);
                if (stack2) {
                  mma_bf16(d, a_hi + (uint64_t)part_off16, b_hi, idesc, true);
                } else if (three) {
                  mma_bf16(d, a_hi, b_hi + (uint64_t)wpart_off16, idesc, true);
                  mma_bf16(d, a_hi + (uint64_t)part_off16, b_hi, idesc, true);
                }
              }
            }
            rem -= len;
            n0 += (uint32_t)(len * CW);
            slot += len;
            if (slot >= Rr) slot -= Rr;
          }
        }
        __syncwarp();
        if (DBG) wC += clock64() - t_m0;
        const bool row_done = k >= NK - 1;  // this input row completes output row k - NK + 1, the oldest of the stack
        if (row_done && elect_one()) commit(&acc_full[fslot]);
        if (row_done && ++fslot == Rr) fslot = 0;
       }
       if (elect_one()) commit(&xb_empty[rx.slot]);
       rx.next();
      }
      gro = rn.slot;
    }
  } else if (warp < 8) {
    // ============================================================ converters (6 warps): fp32 rows -> bf16 hi/lo planes
    const int t = tid - 64;
    constexpr int NT = RS_CONV_WARPS * 32;
    Ring ss(p.NSF), sb(p.NSB);
    const int n_it = p.NP * 128;
    for (int item = cta; item < n_items; item += cpj) {
      const int ug = item / p.bands, band = item - ug * p.bands;
      const int i0 = band * p.BR;
      const int nrows = min(p.BR, p.Nx - i0);
      const int n_in = nrows + NK - 1;
      const int nseg = min(G, p.n_sub - ug * G);
      for (int k0 = 0; k0 < n_in; k0 += p.RU) {
        const int nr = min(p.RU, n_in - k0);  // rows of this unit: their (row, plane, pixel) items share the threads
        wait_t<DBG, true>(&s_full[ss.slot], ss.phase, wA);
        wait_t<DBG, true>(&xb_empty[sb.slot], sb.phase ^ 1, wB);
        for (int it = t; it < ((DBG && (p.skip & 1)) ? 0 : nr * n_it); it += NT) {
          const int r = it >= n_it ? 1 : 0, idx = it - r * n_it;  // RU <= 2
          const unsigned char* xs = x_ring + (size_t)(ss.slot * p.RU + r) * p.x_slot_bytes;
          const int rslot = sb.slot * p.RU + r;  // row of the bf16 ring
          const int pl = idx >> 7, px = idx & 127;
          const int seg = px / PJs, c = px - seg * PJs;
          const float* s0 = reinterpret_cast<const float*>(xs + (size_t)seg * p.seg_bytes) + c + d_off;
          const float* s1 = reinterpret_cast<const float*>(xs + p.src_bytes + (size_t)seg * p.seg_bytes) + c + d_off;
          float v[8];
          if (p.kpack == 2) {
            // K = (window column tl, channel c): element kk of pixel px is x[c][px + tl]
#pragma unroll
            for (int e = 0; e < 8; e++) {
              const int off = voff_s[pl * 8 + e];
              float x = 0.f;
              if (off >= 0 && seg < nseg && c < p.TJ) {
                x = s0[off];
                if (p.has_x1) x -= s1[off];
                else x *= p.scale;
              }
              v[e] = x;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; e++) {
              const int ch = pl * 8 + e;
              float x = 0.f;
              if (ch < p.C && seg < nseg) {
                x = s0[ch * SP];
                if (p.has_x1) x -= s1[ch * SP];
                else x *= p.scale;
              }
              v[e] = x;
            }
          }
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; e++) split2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
          unsigned char* dst = sb_ring + ((size_t)(pl * p.NSBR) + rslot) * p.sb_pitch + (size_t)px * 16;
          *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(dst + (size_t)(p.NP * p.NSBR) * p.sb_pitch) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&xb_full[sb.slot]);
          mbar_arrive(&s_empty[ss.slot]);
        }
        ss.next();
        sb.next();
      }
    }
  } else {
    // ============================================================ epilogue (warps 8-11): drain finished output rows
    const int quarter = warp & 3;
    const int lg = quarter * 32 + lane;
    const int seg = lg / PJs, c = lg - seg * PJs;
    const uint32_t t_lane = tb + ((uint32_t)(quarter * 32) << 16);
    const long long plane = (long long)p.Nx * p.Ny;
    const size_t plane_u = (size_t)plane;
    const int o0 = job * p.Oj;
    const int n_o = min(p.Oj, p.O - o0);
    uint32_t z[16];
#pragma unroll
    for (int e = 0; e < 16; e++) z[e] = 0u;
    int slot = 0;
    uint32_t phase = 0;
    for (int item = cta; item < n_items; item += cpj) {
      const int ug = item / p.bands, band = item - ug * p.bands;
      const int i0 = band * p.BR;
      const int nrows = min(p.BR, p.Nx - i0);
      const int u = ug * G + seg;
      const int b = u / p.strips, j0 = (u - b * p.strips) * p.TJ;
      const bool lane_ok = u < p.n_sub && c < p.TJ && j0 + c < p.Ny;
      float* obase = p.out + ((long long)b * p.O + o0) * plane + (long long)i0 * p.Ny + j0 + c;
      for (int rho = 0; rho < nrows; rho++) {
        wait_t<DBG, true>(&acc_full[slot], phase, wA);
        fence_after_sync();
        float* orow = obase + (long long)rho * p.Ny;
        if (p.npack) {
          // the 16 accumulator columns of this lane are (window column tl, output o) partial sums of INPUT pixel `lane`:
          // out[o](j) = bias + sum_tl column[tl * O + o] of lane j + tl.  Exchange through shared memory (double
          // buffered by row parity, one named barrier of the four epilogue warps per row).
          float v[16];
          tmem_ld16(t_lane + (uint32_t)(slot * p.CW), v);
          tmem_st16(t_lane + (uint32_t)(slot * p.CW), z);
          float* xch = reinterpret_cast<float*>(smem + p.off_xch) + (rho & 1) * (16 * 132);
#pragma unroll
          for (int e = 0; e < 16; e++) xch[e * 132 + lg] = v[e];
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (lane_ok && !(DBG && (p.skip & 4))) {
            for (int o = 0; o < n_o; o++) {
              float acc = bias_s[o];
              for (int tl = 0; tl < p.NL; tl++) acc += xch[(tl * p.O + o) * 132 + lg + tl];
              orow[(size_t)o * plane_u] = acc;
            }
          }
        }
        for (int c0 = 0; c0 < (p.npack ? 0 : p.O_pad); c0 += 16) {
          float v[16];
          tmem_ld16(t_lane + (uint32_t)(slot * p.CW + c0), v);
          tmem_st16(t_lane + (uint32_t)(slot * p.CW + c0), z);
          if (p.stack2) {  // (A_hi + A_lo) W_hi  +  (A_hi + A_lo) W_lo
            float u[16];
            tmem_ld16(t_lane + (uint32_t)(slot * p.CW + p.O_pad + c0), u);
            tmem_st16(t_lane + (uint32_t)(slot * p.CW + p.O_pad + c0), z);
#pragma unroll
            for (int e = 0; e < 16; e++) v[e] += u[e];
          }
          if (lane_ok && !(DBG && (p.skip & 4))) {
            // one coalesced 512-byte store per output channel; the channel stride is warp-uniform
            const int nv = n_o - c0;  // live channels of this chunk (warp-uniform)
            float* q = orow + (size_t)c0 * plane_u;
            const float4* bq = reinterpret_cast<const float4*>(bias_s + c0);
            float bb[16];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const float4 t4 = bq[e];
              bb[4 * e] = t4.x; bb[4 * e + 1] = t4.y; bb[4 * e + 2] = t4.z; bb[4 * e + 3] = t4.w;
            }
            if (nv >= 16) {
#pragma unroll
              for (int e = 0; e < 16; e++) q[(size_t)e * plane_u] = v[e] + bb[e];
            } else {
#pragma unroll
              for (int e = 0; e < 16; e++)
                if (e < nv) q[(size_t)e * plane_u] = v[e] + bb[e];
            }
          }
        }
        tmem_wait_st();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);
        if (++slot == p.Rr) { slot = 0; phase ^= 1; }
      }
    }
  }
  if (DBG && p.dbg && lane == 0) {
    long long* d = p.dbg + ((long long)(blockIdx.y * gridDim.x + blockIdx.x) * 12 + warp) * 4;
    d[0] = wA; d[1] = wB; d[2] = clock64() - t_start; d[3] = wC;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, (uint32_t)p.tmem_cols);
}

// Returns AEFFT_ERR_UNSUPPORTED outside the envelope (the caller falls back to conv_tc_ws / conv_tc / fp32).
int launch_conv_rs(aefft_ctx* ctx, const Window& win, int64_t B, int C, int O, int Nx, int Ny, const float* src0,
                   const float* src1, float pre_div, const float* w, int64_t w_so, int64_t w_sc, const float* bias,
                   float* out, int passes) {
  if (getenv("AEFFT_NO_CONV_RS")) return AEFFT_ERR_UNSUPPORTED;
  if (win.lo != 0 || Ny % 4 != 0 || win.Nk > 8 || win.Nl > 8 || C > 128) return AEFFT_ERR_UNSUPPORTED;
  if ((((uintptr_t)src0 | (uintptr_t)src1) & 15) != 0) return AEFFT_ERR_UNSUPPORTED;
  if (B * (int64_t)C > 0x7fffffffLL) return AEFFT_ERR_UNSUPPORTED;
  ConvRsParams p;
  p.C = C; p.O = O; p.Nx = Nx; p.Ny = Ny; p.NK = win.Nk; p.NL = win.Nl; p.ai0 = win.ai0; p.aj0 = win.aj0;
  p.has_x1 = src1 ? 1 : 0;
  p.scale = (!src1 && pre_div != 0.f) ? 1.f / pre_div : 1.f;
  p.bias = bias; p.out = out; p.passes = passes;
  // K packing: 16 channels per MMA (0); two window columns x 8 channels for C <= 8 (1); or, when a whole window row fits
  // (C * NL <= 16, the 3-channel image layer), K = (window column, channel): ONE MMA triple per input row instead of NLg
  p.kpack = (C * win.Nl <= 16 && !getenv("AEFFT_RS_NO_TAPPACK")) ? 2 : C <= 8 ? 1 : 0;
  p.KS = p.kpack ? 1 : (C + 15) / 16;
  p.NP = p.kpack == 1 ? 1 : 2 * p.KS;
  p.NLg = p.kpack == 2 ? 1 : p.kpack ? (win.Nl + 1) / 2 : win.Nl;
  // Few outputs (the 3-channel reconstruction layer): N = (window row, window column, output) -- NL * O <= 16 columns per
  // output row instead of 16 padded output channels -- so ONE MMA triple per K stage serves all window columns; the
  // epilogue adds the NL column groups of neighbouring lanes (pixel shifts) through shared memory.
  p.npack = (p.kpack == 0 && O * win.Nl <= 16 && win.lo == 0 && !getenv("AEFFT_RS_NO_NPACK")) ? 1 : 0;
  if (p.npack) p.NLg = 1;
  const size_t xch_bytes = p.npack ? (size_t)2 * 16 * 132 * sizeof(float) : 0;
  const int halo = win.Nl - 1;
  p.PJs = (Ny + halo <= 64) ? 64 : 128;
  p.G = 128 / p.PJs;
  p.TJ = (p.PJs - halo) & ~3;
  p.strips = (Ny + p.TJ - 1) / p.TJ;
  if (B * (int64_t)p.strips > 0x3fffffff) return AEFFT_ERR_UNSUPPORTED;
  p.n_sub = (int)(B * p.strips);
  p.sb_pitch = (128 + 8) * 16;
  p.seg_bytes = ((uint32_t)C * (p.PJs + 4) * 4 + 127) & ~127u;
  p.src_bytes = (uint32_t)p.G * p.seg_bytes;
  p.x_slot_bytes = (p.has_x1 ? 2 : 1) * p.src_bytes;
  // outputs per job: as many as fit next to the rings (weights of all K stages stay resident).
  // Two or three CTAs per SM when a configuration fits in that share of the shared memory and of the tensor memory: the
  // kernel is bound by the latency of its single MMA-issuing thread (tensor pipe 17-28 % active, ncu), and every further
  // CTA on the SM brings another issuer, converter set and epilogue (AEFFT_RS_ONE=1 keeps one CTA per SM).
  size_t x_bytes = 0;
  int found = 0;
  p.tmem_cols = 512;
  int per_sm = 1;
  // rows per pipeline unit: the bare pipeline (one barrier round trip per row through four roles) is ~40 % of the kernel
  const int ru_max = (getenv("AEFFT_RS_RU") && atoi(getenv("AEFFT_RS_RU")) == 1) ? 1 : 2;
  const int max_per_sm = getenv("AEFFT_RS_ONE") ? 1 : getenv("AEFFT_RS_TWO") ? 2 : 3;
  for (int ncta = max_per_sm; ncta >= 1 && !found; ncta--) {
    const bool two = ncta > 1;
    const size_t budget = ncta == 3 ? 74 * 1024 : ncta == 2 ? 112 * 1024 : 225 * 1024 - 1024;
    const int cols = 512 >> (ncta == 3 ? 2 : ncta - 1);
    // preference order: double-buffered fp32 staging, few output jobs, deep bf16 ring (one row per unit; see below)
    const int RU = 1;
    for (int NSF = RS_NSF; NSF >= (two ? RS_NSF : 1) && !found; NSF--) {
      x_bytes = (size_t)NSF * RU * p.x_slot_bytes;
      for (int split = 1; split <= (two ? 1 : 8) && !found; split++) {
        const int Oj = ((O + split - 1) / split + 15) / 16 * 16;
        const int O_pad = Oj;
        // [W_hi | W_lo] stacking (2 MMAs of 2N instead of 3 of N) measured no faster on B200 (the N = 80 MMAs are not purely
        // A-read bound): opt-in for experiments only
        const int stack2 = (getenv("AEFFT_RS_STACK2") && !p.npack && passes == 3 && 2 * win.Nk * O_pad <= 256) ? 1 : 0;
        const int CW = stack2 ? 2 * O_pad : O_pad;
        const int Rr = cols / CW > RS_MAXACC ? RS_MAXACC : cols / CW;
        if (Rr < win.Nk + (two ? 3 : 1)) continue;
        const size_t w_bytes = (size_t)p.KS * 2 * p.NLg * 2 * (win.Nk * O_pad) * 16;
        for (int NSB = 4; NSB >= (two ? 3 : 2) && !found; NSB--) {
          const size_t sb_bytes = (size_t)2 * p.NP * NSB * RU * p.sb_pitch;
          if (w_bytes + x_bytes + sb_bytes + xch_bytes + 3 * 1024 <= budget) {
            p.Oj = Oj; p.O_pad = O_pad; p.Rr = Rr; p.NSB = NSB; p.NSF = NSF; p.w_bytes = (uint32_t)w_bytes;
            p.RU = RU; p.NSBR = NSB * RU;
            p.stack2 = stack2; p.CW = CW;
            p.n_jobs = (O + Oj - 1) / Oj;
            p.tmem_cols = cols;
            per_sm = ncta;
            found = 1;
          }
        }
      }
    }
  }
  if (!found || p.n_jobs > 16) return AEFFT_ERR_UNSUPPORTED;
  // Two input rows per pipeline unit when the SAME configuration (CTAs per SM, output jobs) still fits with double
  // buffered staging and a bf16 ring at least 3 units deep: it halves the barrier round trips of producer, converters
  // and issuer (3->16 data gradient 0.177 -> 0.155 ms).  Never at the price of more output jobs or fewer CTAs per SM.
  if (ru_max >= 2 && p.NSF == RS_NSF) {
    const size_t budget = per_sm == 3 ? 74 * 1024 : per_sm == 2 ? 112 * 1024 : 225 * 1024 - 1024;
    for (int NSB = p.NSB; NSB >= 3 && p.RU == 1; NSB--) {
      const size_t need = p.w_bytes + (size_t)RS_NSF * 2 * p.x_slot_bytes + (size_t)2 * p.NP * NSB * 2 * p.sb_pitch + xch_bytes + 3 * 1024;
      if (need <= budget) { p.RU = 2; p.NSB = NSB; p.NSBR = 2 * NSB; x_bytes = (size_t)RS_NSF * 2 * p.x_slot_bytes; }
    }
  }
  p.Ntot = win.Nk * p.O_pad;
  p.off_w = 0;
  p.off_x = (p.w_bytes + 1023) & ~1023u;
  p.off_sb = (uint32_t)((p.off_x + x_bytes + 1023) & ~(size_t)1023);
  p.off_xch = (uint32_t)((p.off_sb + (size_t)2 * p.NP * p.NSBR * p.sb_pitch + 1023) & ~(size_t)1023);
  const size_t smem = p.off_xch + xch_bytes + 1024;
  // work split
  int cpj = per_sm * ctx->sm_count / p.n_jobs;
  if (cpj < 1) cpj = 1;
  const int n_ug = (p.n_sub + p.G - 1) / p.G;
  {
    long long best = -1;
    int best_BR = Nx;
    for (int bands = 1; bands <= 64 && bands <= Nx; bands++) {
      const int BR = (Nx + bands - 1) / bands;
      const int nb = (Nx + BR - 1) / BR;
      const long long items = (long long)n_ug * nb;
      const long long rounds = (items + cpj - 1) / cpj;
      const long long cost = rounds * (BR + win.Nk - 1);
      if (best < 0 || cost < best) { best = cost; best_BR = BR; }
    }
    p.BR = best_BR;
    p.bands = (Nx + p.BR - 1) / p.BR;
  }
  const long long items = (long long)n_ug * p.bands;
  if (items > 0x7fffffff) return AEFFT_ERR_UNSUPPORTED;
  p.items = (int)items;
  if (p.items < cpj) cpj = p.items;
  p.cpj = cpj;
  if (make_tmap_3d_f32(&p.x0_map, src0, Ny, Nx, (uint64_t)B * C, p.PJs + 4, 1, C) != 0) return AEFFT_ERR_UNSUPPORTED;
  p.x1_map = p.x0_map;
  if (src1 && make_tmap_3d_f32(&p.x1_map, src1, Ny, Nx, (uint64_t)B * C, p.PJs + 4, 1, C) != 0) return AEFFT_ERR_UNSUPPORTED;
  // weights
  void* wprep;
  const long long w_elems = (long long)p.n_jobs * p.w_bytes / 2;
  AE_TRY(ctx->get(win.flip ? "rs_wprep_f" : "rs_wprep_t", (size_t)w_elems * 2, &wprep));
  {
    const long long total = w_elems / 2;
    const unsigned blocks = (unsigned)((total + 255) / 256 > 1024 ? 1024 : (total + 255) / 256);
    conv_rs_weight_prep_kernel<<<blocks, 256, 0, ctx->stream>>>(w, w_so, w_sc, C, O, p.Oj, p.O_pad, win.Nk, win.Nl, p.NLg,
                                                               win.flip, p.KS, p.kpack, p.n_jobs, p.stack2, p.npack,
                                                               reinterpret_cast<__nv_bfloat16*>(wprep));
    ctx->launches++;
  }
  p.wprep = reinterpret_cast<const uint4*>(wprep);
  const bool debug = getenv("AEFFT_RS_DEBUG") != nullptr;
  p.dbg = nullptr;
  p.skip = (debug && getenv("AEFFT_RS_SKIP")) ? atoi(getenv("AEFFT_RS_SKIP")) : 0;
  const size_t n_dbg = (size_t)cpj * p.n_jobs * 12 * 4;
  if (debug) {
    AE_TRY(ctx->getT("rs_dbg", n_dbg, &p.dbg));
    AE_CUDA(cudaMemsetAsync(p.dbg, 0, n_dbg * sizeof(long long), ctx->stream));
  }
  AE_TRY(ctx->ensure_dyn_smem((const void*)conv_rs_kernel<false>, smem));
  AE_TRY(ctx->ensure_dyn_smem((const void*)conv_rs_kernel<true>, smem));
  {
    const double px = (double)B * Nx * Ny;
    ProfScope prof(ctx, win.flip ? "conv_fwd_rs" : "conv_dgrad_rs", 2.0 * px * C * O * win.Nk * win.Nl,
                   4.0 * (px * C * (src1 ? 2 : 1) + px * O + (double)C * O * win.Nk * win.Nl));
    if (debug) conv_rs_kernel<true><<<dim3(cpj, p.n_jobs), RS_THREADS, smem, ctx->stream>>>(p);
    else conv_rs_kernel<false><<<dim3(cpj, p.n_jobs), RS_THREADS, smem, ctx->stream>>>(p);
  }
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  if (debug) {
    std::vector<long long> h(n_dbg);
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
    AE_CUDA(cudaMemcpy(h.data(), p.dbg, n_dbg * sizeof(long long), cudaMemcpyDeviceToHost));
    const char* role[4] = {"producer (s_empty, -)", "issuer   (xb_full, acc_empty)", "convert  (s_full, xb_empty)",
                           "epilogue (acc_full, -)"};
    double acc[4][4] = {};
    int cnt[4] = {};
    for (int c = 0; c < cpj * p.n_jobs; c++)
      for (int wv = 0; wv < 12; wv++) {
        const int r = wv == 0 ? 0 : wv == 1 ? 1 : wv < 8 ? 2 : 3;
        for (int q = 0; q < 4; q++) acc[r][q] += (double)h[((size_t)c * 12 + wv) * 4 + q];
        cnt[r]++;
      }
    fprintf(stderr, "[conv_rs] C=%d O=%d %dx%d B=%lld PJs=%d G=%d Oj=%d jobs=%d cpj=%d bands=%d BR=%d Rr=%d RU=%d NSB=%d KS=%d kpack=%d npack=%d smem=%zu tmem=%d\n",
            C, O, Nx, Ny, (long long)B, p.PJs, p.G, p.Oj, p.n_jobs, cpj, p.bands, p.BR, p.Rr, p.RU, p.NSB, p.KS, p.kpack, p.npack, smem, p.tmem_cols);
    for (int r = 0; r < 4; r++)
      fprintf(stderr, "[conv_rs]   %-30s waitA %9.0f  waitB %9.0f  total %9.0f  mma-issue %9.0f cycles\n", role[r],
              acc[r][0] / cnt[r], acc[r][1] / cnt[r], acc[r][2] / cnt[r], acc[r][3] / cnt[r]);
  }
  return AEFFT_OK;
}

}  // namespace aefft
