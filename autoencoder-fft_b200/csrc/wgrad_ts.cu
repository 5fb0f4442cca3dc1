// Streaming weight-gradient kernel: both correlation gradients of a layer pair, the bias gradients and sum e^2 in ONE
// pass over the activations, on tcgen05 with the A operand in TENSOR MEMORY.
//
//   GC[m][d][k][l] = sum_{b,q} dh[b][m](q)  * in[b][d](q + s_t)               (U = dh,  S = in,        window origin ( ai0,  aj0))
//   GF[d][m][k][l] = sum_{b,q} hin[b][m](q) * (out-in)[b][d](q - s_t)         (U = hin, S = out - in,  mirrored window)
//   GB[m] = sum dh[m],  GP[d] = sum (out-in)[d],  SQ = sum (out-in)^2
//
// Why this shape.  With both operands in shared memory the M=64/128 x N=40 MMAs of wgrad_tc.cu are bound by the
// 128 B/clk shared-memory operand bandwidth (A = 2-4 KB per MMA against 20 cycles of math) and most of the M x N tile
// is padding.  Here:
//   * A (TMEM) = the UNSHIFTED operand U, 128 lanes = dM channels x RS row-replicas: lane (rho, m) holds U[m] delayed by
//     rho image rows, so one MMA covers RS window rows at once and every TMEM lane is useful (RS = 128 / dM).  TMEM
//     operands cost no shared-memory bandwidth; the replicas are free because the converter warps write TMEM themselves
//     (tcgen05.st) from an fp32 row ring.
//   * B (smem) = one 8-channel bf16 plane of the SHIFTED operand S in the SWIZZLE_NONE MN-major layout [pixel][8 ch]:
//     consecutive 8-column groups of N are the plane advanced by one pixel (SBO = 16 B), so N = 48 covers a window row
//     (NL <= 6 taps) and a window-row offset is a different ring slot.  1.5 KB per MMA: the MMA is math bound.
//   * K = 16 consecutive pixels of one image row of the strip (pitch PJ = 64 or 128 incl. the NL-1 halo columns); rows
//     STREAM through rings (TMA -> fp32 rows -> converters -> TMEM / bf16 planes), so the halo is paid once per band.
//   * fp32 parity: bf16 hi/lo split of both operands, products hi*hi + hi*lo + lo*hi (BF16X3), fp32 TMEM accumulators
//     kept for ALL items of a CTA; per-CTA partials are reduced in fixed order (deterministic).
//   * Two homes for A, chosen per layer shape (template parameter ASMEM): tensor memory as above, or a DESCENDING
//     shared-memory ring of converted rows [hi|lo][8-px chunk][position][channel][8 px] whose RS consecutive positions are
//     the M = 128 block of an SS MMA (row replicas = descriptor arithmetic; no replication stage, no TMEM A ring).  Two
//     forms of B: separate hi / lo planes (3 MMAs of N = 48 per plane) or one stacked swizzled plane [S_hi | S_lo]
//     (2 MMAs of N = 16 * np * NL).  launch_wgrad_ts documents which shape takes which and the measurements behind it.
// Roles (512 threads): warps 0 / 1 TMA producers of the S / U rows, warps 2-3 S converters (+ sum e, e^2), warps 4-11 U converters (+ sum U),
// warps 12-15 MMA issuers, warps 4-7 epilogue.
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

constexpr int TS_THREADS = 512;
constexpr int TS_NI = 4;      // MMA issuer warps
constexpr int TS_CHUNK = 64;  // pixels per TMEM A chunk = 4 MMA K-steps
constexpr int TS_MAX_JOBS = 8;
constexpr int TS_NSF = 2;     // fp32 S staging slots
constexpr int TS_MAXRING = 24;
constexpr int TS_MAXUB = 20;   // positions of the shared-memory U ring (asmem)

struct TsJob {
  CUtensorMap u_map;   // U [B*dM][Nx][Ny], box {32, 1, dM}, 128-byte swizzle
  CUtensorMap s0_map;  // S source 0 [B*dD][Nx][Ny], box {PJ+4, 1, nch}
  CUtensorMap s1_map;  // GF: S = s0 - s1
  int has_s1;
  int ch0, nch;        // S channels [ch0, ch0+nch) of this job
  int oi, oj;          // S grid origin relative to the item origin
  int rev, is_gf, want_usum;
  long long g_off;
};

struct WgradTsParams {
  TsJob job[TS_MAX_JOBS];
  int n_jobs, cpj;
  float* part;  // [cpj][n_tot]
  long long n_tot, n_main;
  int dM, dD, Nx, Ny, NK, NL;
  int PJ, TJ, CPR, RS, NR, np, Ncol, Acol0, NA, NU, NSB;
  int strips, bands, BR;
  long long items;
  int passes, flip, by_chunk;
  int stack;   // B operand = [S_hi | S_lo] of 8*np channels per pixel in one swizzled plane (2 MMAs of 2N instead of 3*np of N)
  int nacc;    // accumulators per job: NR (stack) or NR*np
  uint32_t u_slot_bytes, s_slot_bytes, s_src_bytes, sb_pitch;
  uint32_t off_u, off_s, off_sb, off_ub, ub_pitch;
  // asmem: A operand read from SHARED memory (SS MMA) -- the bf16 U rows sit in a descending ring [part][8-px chunk]
  // [position][channel][8 px], so the RS row-replicas of an M = 128 block are RS consecutive ring positions
  int asmem, NUB, NUBT;      // ring positions, positions incl. the mirror of the first RS-1 (NUBT = NUB + RS - 1)
  uint32_t a_lbo, a_part;    // bytes between 8-pixel chunks (K core matrices), bytes between the hi and lo parts
  long long* dbg;  // AEFFT_TS_DEBUG: [cta][warp][8] cycles (wait A, wait B, total, work, work 2)
  int skip;        // instrumented build only (AEFFT_TS_SKIP=mask): knock out 1 the S conversion, 2 the MMAs, 4 the U conversion
};

// S converter, one row: fp32 [channel][PJ + 4] (TMA box, halo origin at d_off) -> bf16 hi / lo rows [pixel][8 channels] of
// plane pl (channels 8 pl .. 8 pl + 7 of the job; channels >= nch read as zero).  64 threads; thread t owns the pixels
// t, t + 64 (, ...) of every plane, so plane and pitch are compile-time and the loads take immediate offsets.  For the
// decoder-gradient jobs the row also feeds the bias-gradient sums fs[] and the squared error fq (own pixels only).
template <int PJ>
__device__ __forceinline__ void s_convert_row(const float* __restrict__ s0, const float* __restrict__ s1, int t, int np, int nch,
                                              bool has_s1, bool own_row, int oj, int TJ, int d_off, float (&fs)[16], float& fq,
                                              unsigned char* sb_ring, int slot, int NSB, uint32_t sb_pitch, size_t lo_part,
                                              bool stack) {
  constexpr int SP = PJ + 4, UPP = PJ / 64;
#pragma unroll
  for (int pl = 0; pl < 2; pl++) {
    if (pl >= np) break;
    const int nl = nch - 8 * pl;  // live channels of this plane (may be <= 0)
    float v[UPP][8];
#pragma unroll
    for (int u = 0; u < UPP; u++) {
      const float* b0 = s0 + pl * 8 * SP + t + 64 * u + d_off;
#pragma unroll
      for (int e = 0; e < 8; e++) v[u][e] = e < nl ? b0[e * SP] : 0.f;
    }
    if (has_s1) {
#pragma unroll
      for (int u = 0; u < UPP; u++) {
        const float* b1 = s1 + pl * 8 * SP + t + 64 * u + d_off;
#pragma unroll
        for (int e = 0; e < 8; e++)
          if (e < nl) v[u][e] -= b1[e * SP];
      }
    }
#pragma unroll
    for (int u = 0; u < UPP; u++) {
      const int px = t + 64 * u;
      if (own_row && px + oj >= 0 && px + oj < TJ) {
#pragma unroll
        for (int e = 0; e < 8; e++) {
          fs[pl * 8 + e] += v[u][e];
          fq = fmaf(v[u][e], v[u][e], fq);
        }
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; e++) split2(v[u][2 * e], v[u][2 * e + 1], hi[e], lo[e]);
      if (stack) {
        // pixel row = [hi of 8*np channels | lo of 8*np channels], 16-byte chunks XOR-swizzled on address bits 7.. (32 B /
        // 64 B swizzle; the ring base is 1024-byte aligned, so offset bits = address bits)
        const uint32_t row = (uint32_t)slot * sb_pitch + (uint32_t)px * (32u * np);
        uint32_t o_hi = row + (uint32_t)pl * 16, o_lo = row + (uint32_t)(np + pl) * 16;
        const uint32_t m = np == 1 ? 1u : 3u;
        o_hi ^= ((o_hi >> 7) & m) << 4;
        o_lo ^= ((o_lo >> 7) & m) << 4;
        *reinterpret_cast<uint4*>(sb_ring + o_hi) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(sb_ring + o_lo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      } else {
        unsigned char* dst = sb_ring + ((size_t)(pl * NSB) + slot) * sb_pitch + (size_t)px * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst + lo_part) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
}

template <bool DBG, bool ASMEM>
__global__ void __launch_bounds__(TS_THREADS, 1) wgrad_ts_kernel(const __grid_constant__ WgradTsParams p) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t u_full[TS_MAXRING], u_empty[TS_MAXRING], s_full[TS_NSF], s_empty[TS_NSF],
      sb_full[TS_MAXRING], sb_empty[TS_MAXRING], a_full[8], a_empty[8], ub_full[TS_MAXUB], ub_empty[TS_MAXUB], done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ double usum[128], esum[16], esq;

  // warp index via a broadcast shuffle: the compiler then knows it is warp-uniform and keeps the role loops (MMA
  // descriptors, ring positions) in uniform registers instead of moving them there lane by lane
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int jb = blockIdx.y, cta = blockIdx.x;
  const TsJob& J = p.job[jb];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* u_ring = smem + p.off_u;
  unsigned char* s_ring = smem + p.off_s;
  unsigned char* sb_ring = smem + p.off_sb;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) {
    for (int i = 0; i < p.NU; i++) { mbar_init(&u_full[i], 1); mbar_init(&u_empty[i], 8); }
    for (int i = 0; i < TS_NSF; i++) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 2); }
    for (int i = 0; i < p.NSB; i++) { mbar_init(&sb_full[i], 2); mbar_init(&sb_empty[i], TS_NI); }
    for (int i = 0; i < p.NA; i++) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], TS_NI); }
    if (ASMEM)
      for (int i = 0; i < p.NUB; i++) { mbar_init(&ub_full[i], 8); mbar_init(&ub_empty[i], TS_NI); }
    mbar_init(&done_bar, TS_NI);
    fence_mbar_init();
  }
  if (tid < 128) usum[tid] = 0.0;
  if (tid < 16) esum[tid] = 0.0;
  if (tid == 0) esq = 0.0;
  // zero the bf16 S ring once: the 8 pad pixels behind every row slot are read (times zero A lanes) and must be finite
  {
    const uint32_t n16 = (uint32_t)((p.stack ? 1 : 2 * p.np) * p.NSB) * p.sb_pitch / 16;
    uint4* z = reinterpret_cast<uint4*>(sb_ring);
    for (uint32_t i = tid; i < n16; i += TS_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    if (ASMEM) {
      // the U ring starts as zero rows: the RS-1 rows "before" the first band
      uint4* zu = reinterpret_cast<uint4*>(smem + p.off_ub);
      for (uint32_t i = tid; i < 2 * p.a_part / 16; i += TS_THREADS) zu[i] = make_uint4(0, 0, 0, 0);
    }
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_slot;
  if (warp >= 4 && warp < 8) {
    // accumulators start at zero: every MMA accumulates (the issue order across issuer warps is not fixed)
    uint32_t z[16];
#pragma unroll
    for (int e = 0; e < 16; e++) z[e] = 0u;
    for (int c0 = 0; c0 < p.Acol0; c0 += 16) tmem_st16(tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, z);
    tmem_wait_st();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  long long wA = 0, wB = 0, wC = 0, wD = 0;
  const long long t_start = DBG ? clock64() : 0;

  const int items_per_frame = p.strips * p.bands;
  const int Rmax = (p.NR - 1) * p.RS;
  const int CU = p.dM;
  const int n_items = (int)p.items, cpj = p.cpj;

  if (warp == 0) {
    // ============================================================ TMA producer of the S rows (one lane).  The U rows have
    // their own producer (warp 1): one thread issuing both streams needed ~880 cycles per K-row (1-2 S boxes + PJ/32 U
    // boxes + two waits) and was the slowest role of the bare pipeline; two threads also decouple the two streams.
    if (lane == 0) {
      tma_prefetch_desc(&J.s0_map);
      if (J.has_s1) tma_prefetch_desc(&J.s1_map);
      const int cs_off = J.oj & ~3;  // aligned start column offset (the hardware needs 16-byte aligned inner coordinates)
      const uint32_t s_bytes = (uint32_t)(J.has_s1 ? 2 : 1) * J.nch * (p.PJ + 4) * 4;
      Ring ss(TS_NSF);
      for (int item = cta; item < n_items; item += cpj) {
        const int b = item / items_per_frame;
        const int rem = item - b * items_per_frame;
        const int strip = rem / p.bands;
        const int j0 = strip * p.TJ, i0 = (rem - strip * p.bands) * p.BR;
        const int nrows = min(p.BR, p.Nx - i0);
        const int n_srows = nrows + p.RS - 1 + Rmax;
        const int col = j0 + cs_off, row0 = i0 + J.oi, plane = b * p.dD + J.ch0;
        for (int k = 0; k < n_srows; k++) {
          wait_t<DBG>(&s_empty[ss.slot], ss.phase ^ 1, wA);
          unsigned char* dst = s_ring + (size_t)ss.slot * p.s_slot_bytes;
          mbar_expect_tx(&s_full[ss.slot], s_bytes);
          tma_load_3d(dst, &J.s0_map, col, row0 + k, plane, &s_full[ss.slot]);
          if (J.has_s1) tma_load_3d(dst + p.s_src_bytes, &J.s1_map, col, row0 + k, plane, &s_full[ss.slot]);
          ss.next();
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ TMA producer of the U rows (one lane)
    if (lane == 0) {
      tma_prefetch_desc(&J.u_map);
      Ring su(p.NU);
      const int n_sub = p.PJ / 32;
      for (int item = cta; item < n_items; item += cpj) {
        const int b = item / items_per_frame;
        const int rem = item - b * items_per_frame;
        const int strip = rem / p.bands;
        const int j0 = strip * p.TJ, i0 = (rem - strip * p.bands) * p.BR;
        const int nrows = min(p.BR, p.Nx - i0);
        const int plane = b * CU;
        for (int ru = 0; ru < nrows; ru++) {
          wait_t<DBG>(&u_empty[su.slot], su.phase ^ 1, wB);
          unsigned char* dst = u_ring + (size_t)su.slot * p.u_slot_bytes;
          mbar_expect_tx(&u_full[su.slot], p.u_slot_bytes);
          for (int sub = 0; sub < n_sub; sub++)
            tma_load_3d(dst + (size_t)sub * CU * 128, &J.u_map, j0 + sub * 32, i0 + ru, plane, &u_full[su.slot]);
          su.next();
        }
      }
    }
  } else if (warp >= 12) {
    // ============================================================ MMA issuers (TS_NI warps).  Each warp runs the
    // (compact, not unrolled) loops with warp-uniform operands; one elected lane issues.  A single issuing thread
    // sustains only ~1 MMA group per 200 cycles (descriptor arithmetic + R2UR latency), so the work is split:
    //   by_chunk (one window-row group, NR == 1): issuer q owns the A chunks cc = q (mod TS_NI) and its OWN copy of the
    //             accumulators (summed in fixed order by the epilogue);
    //   by_acc   : issuer q owns the accumulators a = q (mod TS_NI) for every chunk.
    // Either way each accumulator (copy) receives its MMAs from one thread in program order: results are deterministic.
    const int q = warp - 12;
    const uint32_t idesc = make_idesc_bf16(128, p.Ncol, 0, 1);
    const uint32_t sb_base16 = smem_u32(sb_ring) >> 4, pitch16 = p.sb_pitch >> 4;
    const uint32_t lo_off16 = (uint32_t)(p.np * p.NSB) * pitch16;
    const bool stack = p.stack != 0;
    const uint32_t pxb16 = stack ? (uint32_t)p.np * 2 : 1;  // 16-byte units per pixel row of the B operand
    // stacked: SWIZZLE_32B / 64B MN-major plane, N atoms one pixel apart (LBO), 8-pixel K groups (SBO)
    const uint64_t desc_const = stack ? (make_desc(0, pxb16 * 16, pxb16 * 128) | ((uint64_t)(p.np == 1 ? 6 : 4) << 61))
                                      : make_desc(0, 128, 16);
    const int NR = p.NR, NSB = p.NSB, np = stack ? 1 : p.np, RS = p.RS, CPR = p.CPR;
    const int nacc = p.nacc;
    const uint32_t Ncol = (uint32_t)p.Ncol;
    const bool three = p.passes == 3, by_chunk = p.by_chunk != 0;
    const uint32_t d_base = tb + (by_chunk ? (uint32_t)(q * nacc) * Ncol : 0u);
    Ring ra(p.NA);         // TMEM A chunk ring (advanced for every chunk, owned or not)
    Ring rw(NSB);          // ring position of S row `sb_row` of the current item
    int sb_row = 0;        // next S row of the current item this warp has not waited for
    int s0slot = 0;        // ring slot of the S row of the current K-row (window row offset 0)
    int cc4 = 0;           // chunk counter mod TS_NI
    // asmem: ring position / phase of the U row of the current K-row (positions DEscend so that older rows follow in
    // memory), rows seen so far, descriptor pieces in 16-byte units
    int upos = p.NUB - 1;
    uint32_t uph = 0;
    const uint32_t ub_base16 = smem_u32(smem + p.off_ub) >> 4, a_lbo16 = p.a_lbo >> 4, a_part16 = p.a_part >> 4;
    const uint32_t a_pos16 = (uint32_t)p.dM;  // one ring position = dM rows of 16 bytes
    const uint64_t a_desc0 = make_desc(0, p.a_lbo, 128);
    for (int item = cta; item < n_items; item += cpj) {
      const int rem = item % items_per_frame;
      const int i0 = (rem % p.bands) * p.BR;
      const int nrows = min(p.BR, p.Nx - i0);
      const int n_krows = nrows + RS - 1, n_srows = n_krows + Rmax;
      for (int r = 0; r < n_krows; r++) {
        // S rows [r, r+Rmax].  EVERY issuer waits for EVERY row in order and commits every row's release (count TS_NI),
        // also for K-rows whose chunks belong to other issuers: a parity wait is only sound for a waiter that is neither
        // ahead of the previous phase nor lapped, which in-order waiting + all-issuer release guarantee.
        for (; sb_row < r + Rmax + 1; sb_row++) {
          wait_t<DBG>(&sb_full[rw.slot], rw.phase, wA);
          rw.next();
        }
        if (ASMEM) {
          // the new U row (ring position upos) has been converted; every issuer waits for every row in order
          wait_t<DBG>(&ub_full[upos], uph, wB);
          fence_after_sync();
        }
        for (int h = 0; h < CPR; h++, cc4 = (cc4 + 1) & (TS_NI - 1)) {
          const bool mine = !by_chunk || cc4 == q;
          // TMEM A: every issuer waits for every A chunk in order and takes part in its release (count TS_NI), owner or
          // not: a parity wait must neither skip phases nor be lapped (the ring length need not be a multiple of TS_NI)
          if (!ASMEM) wait_t<DBG>(&a_full[ra.slot], ra.phase, wB);
          if (mine) {
            fence_after_sync();
            const long long t_m0 = DBG ? clock64() : 0;
            if (elect_one() && !(DBG && (p.skip & 2))) {
              // one lane walks the (K-step, window-row group, plane) nest; operands advance by constant increments
              const uint32_t pl_step16 = (uint32_t)NSB * pitch16;
              // A operand: TMEM chunk columns, or the shared-memory descriptor of rows upos .. upos+RS-1, 8-px chunk 8h
              const uint32_t a_tm = tb + (uint32_t)(p.Acol0 + ra.slot * 64);
              uint64_t a_sm = a_desc0 + (uint64_t)(ub_base16 + (uint32_t)(8 * h) * a_lbo16 + (uint32_t)upos * a_pos16);
#pragma unroll 1
              for (int ks = 0; ks < 4; ks++, a_sm += 2 * a_lbo16) {
                const uint32_t a_hi = a_tm + ks * 8, a_lo = a_hi + 32;
                const uint64_t as_hi = a_sm, as_lo = a_sm + (uint64_t)a_part16;
                const uint32_t px16 = sb_base16 + (uint32_t)(h * TS_CHUNK + ks * 16) * pxb16;
                uint32_t d = d_base;
                int slot = s0slot, a = 0;
#pragma unroll 1
                for (int Ri = 0; Ri < NR; Ri++) {
                  uint32_t lo32 = px16 + (uint32_t)slot * pitch16;
#pragma unroll 1
                  for (int pl = 0; pl < np; pl++, d += Ncol, a++, lo32 += pl_step16) {
                    if (by_chunk || (a & (TS_NI - 1)) == q) {
                      const uint64_t b_hi = desc_const + (uint64_t)lo32, b_lo = b_hi + (uint64_t)lo_off16;
                      if (ASMEM) {
                        mma_bf16(d, as_hi, b_hi, idesc, true);
                        if (stack) {
                          if (three) mma_bf16(d, as_lo, b_hi, idesc, true);  // (A_hi + A_lo) [S_hi | S_lo]
                        } else if (three) {
                          mma_bf16(d, as_hi, b_lo, idesc, true);
                          mma_bf16(d, as_lo, b_hi, idesc, true);
                        }
                      } else {
                        mma_bf16_ts(d, a_hi, b_hi, idesc, true);
                        if (stack) {
                          if (three) mma_bf16_ts(d, a_lo, b_hi, idesc, true);  // (A_hi + A_lo) [S_hi | S_lo]
                        } else if (three) {
                          mma_bf16_ts(d, a_hi, b_lo, idesc, true);
                          mma_bf16_ts(d, a_lo, b_hi, idesc, true);
                        }
                      }
                    }
                  }
                  slot += RS;
                  if (slot >= NSB) slot -= NSB;
                }
              }
            }
            __syncwarp();
            if (DBG) wC += clock64() - t_m0;
          }
          if (!ASMEM) {
            if (elect_one()) commit(&a_empty[ra.slot]);
            ra.next();
          }
        }
        if (ASMEM) {
          // the row RS-1 behind the newest one has been multiplied for the last time.  This also holds for the RS-1 virtual
          // zero rows before the very first row: they live in the mirror of positions RS-2 .. 0, which the converters may
          // only overwrite after this release (see the converter's wait).
          {
            int old = upos + RS - 1;
            if (old >= p.NUB) old -= p.NUB;
            if (elect_one()) commit(&ub_empty[old]);
          }
          if (--upos < 0) { upos = p.NUB - 1; uph ^= 1; }
        }
        if (elect_one()) commit(&sb_empty[s0slot]);
        if (++s0slot == NSB) s0slot = 0;
      }
      // S rows that were only ever read at a window-row offset (by_acc only: Rmax == 0 in by_chunk mode)
      for (int k = n_krows; k < n_srows; k++) {
        for (; sb_row < k + 1; sb_row++) {
          wait_t<DBG>(&sb_full[rw.slot], rw.phase, wA);
          rw.next();
        }
        if (elect_one()) commit(&sb_empty[s0slot]);
        if (++s0slot == NSB) s0slot = 0;
      }
      sb_row = 0;
    }
    if (elect_one()) commit(&done_bar);
  } else if (warp < 4) {
    // ============================================================ S converters (64 threads)
    const int t = tid - 64;
    const int d_off = J.oj - (J.oj & ~3);  // sub-offset of the halo origin inside the aligned box
    const int PJ = p.PJ, TJ = p.TJ;
    const int nch = J.nch, oj = J.oj, oi = J.oi;
    const bool has_s1 = J.has_s1 != 0, is_gf = J.is_gf != 0;
    const size_t lo_part = (size_t)(p.np * p.NSB) * p.sb_pitch;
    Ring ss(TS_NSF), sb(p.NSB);
    float fs[16], fq = 0.f;
#pragma unroll
    for (int e = 0; e < 16; e++) fs[e] = 0.f;
    for (int item = cta; item < n_items; item += cpj) {
      const int rem = item % items_per_frame;
      const int i0 = (rem % p.bands) * p.BR;
      const int nrows = min(p.BR, p.Nx - i0);
      const int n_srows = nrows + p.RS - 1 + Rmax;
      for (int k = 0; k < n_srows; k++) {
        wait_t<DBG>(&s_full[ss.slot], ss.phase, wA);
        wait_t<DBG>(&sb_empty[sb.slot], sb.phase ^ 1, wB);
        const float* s0 = reinterpret_cast<const float*>(s_ring + (size_t)ss.slot * p.s_slot_bytes);
        const float* s1 = reinterpret_cast<const float*>(s_ring + (size_t)ss.slot * p.s_slot_bytes + p.s_src_bytes);
        const bool own_row = is_gf && (k + oi >= 0) && (k + oi < nrows);
        const long long t_c0 = DBG ? clock64() : 0;
        if (DBG && (p.skip & 1)) {
        } else if (PJ == 128)
          s_convert_row<128>(s0, s1, t, p.np, nch, has_s1, own_row, oj, TJ, d_off, fs, fq, sb_ring, sb.slot, p.NSB, p.sb_pitch,
                             lo_part, p.stack != 0);
        else
          s_convert_row<64>(s0, s1, t, p.np, nch, has_s1, own_row, oj, TJ, d_off, fs, fq, sb_ring, sb.slot, p.NSB, p.sb_pitch,
                            lo_part, p.stack != 0);
        if (DBG) wC += clock64() - t_c0;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&sb_full[sb.slot]);
          mbar_arrive(&s_empty[ss.slot]);
        }
        ss.next();
        sb.next();
      }
      if (is_gf) {
        // per-item flush of the fp32 running sums into double accumulators (few hundred values per thread and item)
#pragma unroll
        for (int e = 0; e < 16; e++) {
          if (fs[e] != 0.f) atomicAdd(&esum[e], (double)fs[e]);
          fs[e] = 0.f;
        }
        if (fq != 0.f) atomicAdd(&esq, (double)fq);
        fq = 0.f;
      }
    }
  } else {
    // ============================================================ U converters (8 warps = 256 threads), two stages per
    // K-row r:  A) the NEW image row r is converted ONCE: fp32 (TMA ring, 128B-swizzled) -> packed bf16 hi / lo pixel pairs
    //              in a shared ring of RS rows [slot][hi|lo][channel][pixel], halo columns zeroed, channel sums taken;
    //           B) every TMEM lane (rho, m) copies its row r - rho of that ring into the A chunks (16-byte loads, no
    //              arithmetic): the RS row-replicas cost loads only, not RS conversions.
    // The two stages are separated by a named barrier of the 256 converter threads.
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const int L = quarter * 32 + lane;
    const int rho = L / CU, m = L - rho * CU;
    const int t256 = tid - 128;
    const uint32_t t_lane = tb + ((uint32_t)(quarter * 32) << 16) + (uint32_t)p.Acol0;
    const int NA = p.NA, RS = p.RS, PJ = p.PJ, TJ = p.TJ;
    const uint32_t ub_pitch = p.ub_pitch, ub_part = (uint32_t)CU * ub_pitch;  // bytes: channel row, hi (or lo) block
    unsigned char* ub = smem + p.off_ub;
    const int units_per_m = PJ >> 3, n_units = CU * units_per_m;  // 8-pixel units of one image row
    const int upm_shift = PJ == 128 ? 4 : 3;                     // units_per_m is 16 or 8
    if (ASMEM) {
      // ---- A operand in shared memory: every image row is converted ONCE into ring position upos of the layout
      // [hi|lo][8-px chunk][position][channel][8 px] (the SWIZZLE_NONE K-major core matrices of an M x 16 block whose M rows
      // are `position * dM + channel`); positions descend with time, so the block of K-row r = positions upos .. upos+RS-1
      // = rows r, r-1, .. r-RS+1: the row replicas are plain address arithmetic.  Positions < RS-1 are mirrored behind the
      // ring so that a block never wraps.  Every band ends with RS-1 zero rows, which are also the rows "before" the next band.
      Ring ru_ring2(p.NU);
      int upos = p.NUB - 1;
      uint32_t uph = 0;
      const int kcs = PJ >> 3;                          // 8-pixel chunks per row
      const int n_units = CU * kcs;                     // (chunk, channel) units of a row; channel varies fastest
      const int cu_shift = 31 - __clz(CU);              // CU is a power of two (128 % dM == 0)
      unsigned char* a_sm = smem + p.off_ub;
      for (int item = cta; item < n_items; item += cpj) {
        const int rem = item % items_per_frame;
        const int i0 = (rem % p.bands) * p.BR;
        const int nrows = min(p.BR, p.Nx - i0);
        const int n_krows = nrows + RS - 1;
        float us = 0.f;  // this thread's channel sum over the band (its channel is fixed: 256 % dM == 0)
        for (int r = 0; r < n_krows; r++) {
          const bool real = r < nrows;
          if (real) wait_t<DBG>(&u_full[ru_ring2.slot], ru_ring2.phase, wA);
          // Release of this position by the issuers.  Positions >= RS-1 are free on the first lap (parity trick); positions
          // < RS-1 are NOT: their mirror holds the virtual zero rows the first RS-1 K-rows read, and the issuers release
          // them like real rows -- so these positions have seen one more phase and wait with the un-flipped parity.
          const bool mirror = upos < RS - 1;
          wait_t<DBG>(&ub_empty[upos], mirror ? uph : uph ^ 1, wB);
          const long long t_c0 = DBG ? clock64() : 0;
          const unsigned char* urow = u_ring + (size_t)ru_ring2.slot * p.u_slot_bytes;
          for (int unit = t256; unit < ((DBG && (p.skip & 4)) ? 0 : n_units); unit += 256) {
            const int kc = unit >> cu_shift, mm = unit & (CU - 1);
            uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
            if (real) {
              const int px0 = kc * 8, sub = px0 >> 5, g0 = (px0 & 31) >> 2;
              const unsigned char* src = urow + (size_t)sub * CU * 128 + (size_t)mm * 128;
              float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
              if (px0 < TJ) q0 = *reinterpret_cast<const float4*>(src + ((g0 ^ (mm & 7)) << 4));
              if (px0 + 4 < TJ) q1 = *reinterpret_cast<const float4*>(src + (((g0 + 1) ^ (mm & 7)) << 4));
              us += (q0.x + q0.y) + (q0.z + q0.w) + (q1.x + q1.y) + (q1.z + q1.w);
              split2(q0.x, q0.y, hi[0], lo[0]);
              split2(q0.z, q0.w, hi[1], lo[1]);
              split2(q1.x, q1.y, hi[2], lo[2]);
              split2(q1.z, q1.w, hi[3], lo[3]);
            }
            unsigned char* d = a_sm + (size_t)kc * p.a_lbo + ((size_t)upos * CU + mm) * 16;
            *reinterpret_cast<uint4*>(d) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(d + p.a_part) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            if (mirror) {
              unsigned char* d2 = d + (size_t)p.NUB * CU * 16;
              *reinterpret_cast<uint4*>(d2) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(d2 + p.a_part) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
          if (DBG) wC += clock64() - t_c0;
          fence_proxy_async();  // generic-proxy writes -> the MMAs' async-proxy reads
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&ub_full[upos]);
            if (real) mbar_arrive(&u_empty[ru_ring2.slot]);
          }
          if (real) ru_ring2.next();
          if (--upos < 0) { upos = p.NUB - 1; uph ^= 1; }
        }
        if (J.want_usum && us != 0.f) atomicAdd(&usum[t256 & (CU - 1)], (double)us);
      }
    } else {
      Ring ru_ring(p.NU);            // fp32 ring position of the next image row to convert
      Ring rc(NA);                   // A ring position of the current chunk (advanced for EVERY chunk, see below)
      int parity = 0;                // (chunk index & 1) of the next chunk in program order
      auto conv_barrier = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
      for (int item = cta; item < n_items; item += cpj) {
        const int rem = item % items_per_frame;
        const int i0 = (rem % p.bands) * p.BR;
        const int nrows = min(p.BR, p.Nx - i0);
        const int n_krows = nrows + RS - 1;
        for (int r = 0; r < n_krows; r++) {
          // ---- stage A: convert image row r (if any) into ring slot r % RS
          if (r < nrows) {
            wait_t<DBG>(&u_full[ru_ring.slot], ru_ring.phase, wA);
            const long long t_c0 = DBG ? clock64() : 0;
            const unsigned char* urow = u_ring + (size_t)ru_ring.slot * p.u_slot_bytes;
            unsigned char* dst_hi = ub + (size_t)(r & (RS - 1)) * 2 * ub_part;
            for (int unit = t256; unit < ((DBG && (p.skip & 4)) ? 0 : n_units); unit += 256) {
              const int mm = unit >> upm_shift, u8 = unit & (units_per_m - 1);
              const int px0 = u8 * 8, sub = px0 >> 5, g0 = (px0 & 31) >> 2;
              const unsigned char* src = urow + (size_t)sub * CU * 128 + (size_t)mm * 128;
              float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
              if (px0 < TJ) q0 = *reinterpret_cast<const float4*>(src + ((g0 ^ (mm & 7)) << 4));
              if (px0 + 4 < TJ) q1 = *reinterpret_cast<const float4*>(src + (((g0 + 1) ^ (mm & 7)) << 4));
              if (J.want_usum) {
                // 16 (PJ = 128) or 8 (PJ = 64) consecutive lanes share a channel: reduce them, one atomic per channel and row
                float sv = (q0.x + q0.y) + (q0.z + q0.w) + (q1.x + q1.y) + (q1.z + q1.w);
                for (int o = units_per_m >> 1; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
                if ((lane & (units_per_m - 1)) == 0) atomicAdd(&usum[mm], (double)sv);
              }
              uint32_t hi[4], lo[4];
              split2(q0.x, q0.y, hi[0], lo[0]);
              split2(q0.z, q0.w, hi[1], lo[1]);
              split2(q1.x, q1.y, hi[2], lo[2]);
              split2(q1.z, q1.w, hi[3], lo[3]);
              unsigned char* d = dst_hi + (size_t)mm * ub_pitch + (size_t)u8 * 16;
              *reinterpret_cast<uint4*>(d) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(d + ub_part) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            if (DBG) wC += clock64() - t_c0;
          }
          conv_barrier();  // row r is in the bf16 ring; the fp32 slot is free again
          if (r < nrows) {
            if (lane == 0) mbar_arrive(&u_empty[ru_ring.slot]);
            ru_ring.next();
          }
          // ---- stage B: TMEM lane (rho, m) <- row r - rho
          const int rrow = r - rho;
          const bool valid = rrow >= 0 && rrow < nrows;
          const unsigned char* lrow = ub + (size_t)(rrow & (RS - 1)) * 2 * ub_part + (size_t)m * ub_pitch;
          for (int h = 0; h < p.CPR; h++, parity ^= 1, rc.next()) {
            // Every converter warp waits for the release of EVERY chunk slot in order, also for the chunks the other half
            // writes: with an odd ring length both halves alternate on the same barriers, and a parity wait that skips a
            // phase can pass one lap early (the slot would be overwritten while the MMAs still read it).
            wait_t<DBG>(&a_empty[rc.slot], rc.phase ^ 1, wB);
            if (parity != half) continue;
            const int ca = rc.slot;
            const long long t_c1 = DBG ? clock64() : 0;
            fence_after_sync();
  #pragma unroll
            for (int part = 0; part < 2; part++) {
              uint32_t v[32];
  #pragma unroll
              for (int g = 0; g < 8; g++) {
                uint4 q = make_uint4(0u, 0u, 0u, 0u);
                if (valid) q = *reinterpret_cast<const uint4*>(lrow + (size_t)part * ub_part + (size_t)h * 128 + g * 16);
                v[4 * g] = q.x; v[4 * g + 1] = q.y; v[4 * g + 2] = q.z; v[4 * g + 3] = q.w;
              }
              tmem_st16(t_lane + (uint32_t)(ca * 64 + part * 32), v);
              tmem_st16(t_lane + (uint32_t)(ca * 64 + part * 32 + 16), v + 16);
            }
            tmem_wait_st();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[ca]);
            if (DBG) wD += clock64() - t_c1;
          }
          conv_barrier();  // every lane has read its rows before stage A overwrites the oldest slot
        }
      }
    }
    // ============================================================ epilogue (warps 4-7)
    if (half == 0) {
      mbar_wait(&done_bar, 0);
      fence_after_sync();
      const bool had_items = cta < n_items;
      float* part = p.part + (long long)cta * p.n_tot + J.g_off;
      const int TT = p.NK * p.NL;
      for (int acc = 0; acc < p.nacc; acc++) {
        const int Ri = p.stack ? acc : acc / p.np, pl0 = p.stack ? 0 : acc - Ri * p.np;
        const int tk0 = Ri * p.RS + rho;
        const bool pair_cols = p.stack && p.np == 2;  // stacked, two planes: [hi 16 | lo 16] per shift, summed here
        for (int c0 = 0; c0 < p.Ncol; c0 += pair_cols ? 32 : 16) {
          float v[16];
          // accumulator columns c0..c0+15, issuer copies added in fixed order
          auto load_cols = [&](int col, float (&dst)[16]) {
            tmem_ld16(tb + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.Ncol + col), dst);
            if (p.by_chunk) {
              for (int cp = 1; cp < TS_NI; cp++) {
                float w[16];
                tmem_ld16(tb + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((cp * p.nacc + acc) * p.Ncol + col), w);
#pragma unroll
                for (int e = 0; e < 16; e++) dst[e] += w[e];
              }
            }
          };
          load_cols(c0, v);
          if (pair_cols) {
            float w2[16];
            load_cols(c0 + 16, w2);
#pragma unroll
            for (int e = 0; e < 16; e++) v[e] += w2[e];
          }
          // which (shift, channel) do these columns hold?
          int s, chb, nval;
          if (!p.stack) { s = c0 >> 3; chb = pl0 * 8; nval = 16; }           // two shifts x 8 channels
          else if (p.np == 1) { s = c0 >> 4; chb = 0; nval = 8; }             // [hi 8 | lo 8] of one shift
          else { s = c0 >> 5; chb = 0; nval = 16; }                           // [hi 16 | lo 16] of one shift, already summed
          if (p.stack && p.np == 1) {
#pragma unroll
            for (int e = 0; e < 8; e++) v[e] += v[8 + e];
          }
          if (tk0 < p.NK) {
#pragma unroll
            for (int e = 0; e < 16; e++) {
              if (e >= nval) continue;
              const int sh = p.stack ? s : s + (e >> 3);
              const int ch = p.stack ? chb + e : chb + (e & 7);
              if (sh < p.NL && ch < J.nch) {
                int tk = tk0, tl = sh;
                if (J.rev) { tk = p.NK - 1 - tk; tl = p.NL - 1 - tl; }
                const int k = p.flip ? p.NK - 1 - tk : tk, l = p.flip ? p.NL - 1 - tl : tl;
                const int d = J.ch0 + ch;
                const long long gi = J.is_gf ? (((long long)d * p.dM + m) * TT + k * p.NL + l)
                                             : (((long long)m * p.dD + d) * TT + k * p.NL + l);
                const float val = had_items ? v[e] : 0.f;
                part[gi] = val;
              }
            }
          }
        }
      }
    }
  }
  if (DBG && p.dbg && lane == 0) {
    long long* d = p.dbg + ((long long)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + warp) * 8;
    d[0] = wA; d[1] = wB; d[2] = clock64() - t_start; d[3] = wC; d[4] = wD;
  }
  fence_before_sync();
  __syncthreads();
  // bias-gradient sums and sum e^2 of this CTA
  {
    float* prow = p.part + (long long)cta * p.n_tot;
    const long long nC2 = 2LL * p.dM * p.dD * p.NK * p.NL;
    if (J.want_usum && tid < CU) prow[nC2 + tid] = (float)usum[tid];
    if (J.is_gf && tid < J.nch) prow[nC2 + p.dM + J.ch0 + tid] = (float)esum[tid];
    if (tid == 0) prow[p.n_main + jb] = J.is_gf ? (float)esq : 0.f;
  }
  if (warp == 0) tmem_dealloc(tb, 512);
}

// out[0 .. n_main) = sum over CTAs of the partial blocks; *sq = sum over CTAs and jobs of the per-job sum e^2 slots
__global__ void wgrad_ts_reduce_kernel(const float* __restrict__ part, int n_parts, long long n_tot, long long n_main,
                                       int n_jobs, long long nC2, int dM, int dD, float* __restrict__ G, float* __restrict__ GB,
                                       float* __restrict__ GP, float* __restrict__ SQ) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx > n_main) return;
  double s = 0.0;
  if (idx < n_main) {
    for (int c = 0; c < n_parts; c++) s += (double)part[(long long)c * n_tot + idx];
    if (idx < nC2) G[idx] = (float)s;
    else if (idx < nC2 + dM) GB[idx - nC2] = (float)s;
    else GP[idx - nC2 - dM] = (float)s;
  } else {
    for (int c = 0; c < n_parts; c++)
      for (int j = 0; j < n_jobs; j++) s += (double)part[(long long)c * n_tot + n_main + j];
    if (SQ) *SQ = (float)s;
  }
}

// GC | GF (raw sums over the B frames) into G, GB[dM], GP[dD], SQ[1].  AEFFT_ERR_UNSUPPORTED outside the envelope
// (the caller then uses wgrad_tc / the fp32 kernels + channel sums).
int launch_wgrad_ts(aefft_ctx* ctx, const Window& win, int64_t B, int dD, int dM, int Nx, int Ny, const float* in,
                    const float* out, const float* hin, const float* dh, float* G, float* GB, float* GP, float* SQ,
                    int passes) {
  if (getenv("AEFFT_NO_WGRAD_TS")) return AEFFT_ERR_UNSUPPORTED;
  if (win.lo != 0 || win.Nl > 6 || win.Nk > 16 || Ny % 4 != 0 || dM < 8 || dM > 128 || (128 % dM) != 0 || dD < 1)
    return AEFFT_ERR_UNSUPPORTED;
  if ((((uintptr_t)in | (uintptr_t)out | (uintptr_t)hin | (uintptr_t)dh) & 15) != 0) return AEFFT_ERR_UNSUPPORTED;
  if (B * (int64_t)dM > 0x7fffffffLL || B * (int64_t)dD > 0x7fffffffLL) return AEFFT_ERR_UNSUPPORTED;
  WgradTsParams p;
  p.dM = dM; p.dD = dD; p.Nx = Nx; p.Ny = Ny; p.NK = win.Nk; p.NL = win.Nl;
  p.passes = passes; p.flip = win.flip;
  p.RS = 128 / dM;
  p.NR = (win.Nk + p.RS - 1) / p.RS;
  // Stacked B operand: one swizzled plane [pixel][hi of 8*np channels | lo of 8*np channels], N = 16*np*NL columns per
  // MMA, 2 MMAs per (K-step, window-row group) instead of 3*np of N = 48 (validated by tools/probe_ts.cu: MN-major
  // SWIZZLE_32B/64B with a one-pixel atom stride).  A TS MMA costs ~64 cycles up to N = 128 and ~110 at N = 192, so the
  // stack pays where the kernel is MMA bound AND the stacked form keeps two channel planes per job (config 2 pair 1:
  // 0.54 -> 0.46 ms).  With one plane (pair 0: converter bound, 0.667 vs 0.662 ms) or when stacking would halve the
  // channels per job and double the jobs (pair 2: 0.63 vs 0.44 ms) the separate hi / lo planes stay.
  // AEFFT_TS_STACK=0 / 1 forces the choice.
  p.stack = 0;
  bool stack_needs_asmem = false;
  {
    const char* force = getenv("AEFFT_TS_STACK");
    int np_plain = dD > 8 ? 2 : 1;
    while (np_plain > 1 && p.NR * np_plain * 48 > 512 - 2 * 64) np_plain--;
    const bool allow = force ? force[0] != '0' : true;
    for (int np = dD > 8 ? 2 : 1; allow && win.Nl <= 6 && np >= 1 && !p.stack; np--) {
      if (!force && (np != np_plain || np != 2)) continue;
      const int ncol = 16 * np * win.Nl;
      const int copies = (p.NR == 1 && TS_NI * ncol + 2 * 64 <= 512) ? TS_NI : 1;
      const int jobs = 2 * ((dD + 8 * np - 1) / (8 * np));
      if (copies * p.NR * ncol + 2 * 64 <= 512 && jobs <= TS_MAX_JOBS) {
        p.stack = 1; p.np = np; p.Ncol = ncol; p.nacc = p.NR;
      }
    }
    // With the A operand in shared memory the tensor memory holds accumulators only, so the two-plane stack also fits
    // when NR * ncol <= 512 (config 2 pair 2: 3 groups x 160 columns; 24 N160 MMAs per K-row instead of 72 N48).
    const char* as_env = getenv("AEFFT_TS_ASMEM");
    if (!p.stack && allow && win.Nl <= 6 && np_plain == 2 && !(as_env && as_env[0] == '0')) {
      const int ncol = 32 * win.Nl, jobs = 2 * ((dD + 15) / 16);
      const int halo = win.Nl - 1, PJ = (Ny + halo <= 64) ? 64 : 128;
      // shared memory of that configuration at the natural strip pitch (same formulas as the layout search below)
      const size_t sbp = ((size_t)(PJ + 8) * 32 * 2 + 1023) & ~(size_t)1023;
      const size_t need = (size_t)2 * dM * PJ * 4 + 2 * ((size_t)(PJ / 8) * (2 * p.RS) * dM * 16) +
                          (size_t)TS_NSF * 2 * (((size_t)16 * (PJ + 4) * 4 + 127) & ~(size_t)127) +
                          (size_t)((p.NR - 1) * p.RS + 3) * sbp + 4096;
      if (p.NR * ncol <= 512 && jobs <= TS_MAX_JOBS && need <= 225 * 1024 - 1024) {
        p.stack = 1; p.np = 2; p.Ncol = ncol; p.nacc = p.NR;
        stack_needs_asmem = true;
      }
    }
  }
  if (!p.stack) {
    p.Ncol = 48;
    p.np = dD > 8 ? 2 : 1;
    while (p.np > 1 && p.NR * p.np * p.Ncol > 512 - 2 * 64) p.np--;
    if (p.NR * p.np * p.Ncol > 512 - 2 * 64) return AEFFT_ERR_UNSUPPORTED;
    p.nacc = p.NR * p.np;
  }
  const int jobs_per_grad = (dD + 8 * p.np - 1) / (8 * p.np);
  p.n_jobs = 2 * jobs_per_grad;
  if (p.n_jobs > TS_MAX_JOBS) return AEFFT_ERR_UNSUPPORTED;
  // A operand from shared memory (asmem, default) or from tensor memory (AEFFT_TS_ASMEM=0, and the fallback when the
  // shared-memory ring does not fit).  Strip geometry: pitch 64 when one 64-wide strip covers the row, else 128; the
  // shared-memory A ring is large ((2 RS + slack) rows of hi + lo), so 64-pixel strips are also taken when 128 does not fit.
  const int halo = win.Nl - 1;
  const int Rmax = (p.NR - 1) * p.RS;
  p.by_chunk = (p.NR == 1 && TS_NI * p.nacc * p.Ncol + 2 * 64 <= 512) ? 1 : 0;
  p.Acol0 = (p.by_chunk ? TS_NI : 1) * p.nacc * p.Ncol;
  p.NA = (512 - p.Acol0) / 64;  // TMEM A chunks (unused with the shared-memory A form)
  if (p.NA > 4) p.NA = 4;
  if (p.NA < 0) p.NA = 0;
  p.NSB = Rmax + 3;
  if (p.NSB > TS_MAXRING) return AEFFT_ERR_UNSUPPORTED;
  const size_t budget = 225 * 1024 - 1024;
  size_t sb_bytes = 0, s_bytes = 0, ub_bytes = 0;
  p.asmem = -1;
  // Default: shared-memory A when one window-row group covers the window (NR == 1, e.g. dM = 16): there the kernel is
  // bound by the converter warps and the TMEM replication stage disappears (config 2 pair 0: 0.665 -> 0.60 ms).  With
  // several groups the kernel is MMA bound and an SS MMA (4 KB of A per MMA over the 128 B/clk operand path, ~43 cycles)
  // is no faster than the TS MMA (~50-64 cycles), while the larger ring forces 64-pixel strips (pair 1: 0.71 vs 0.47 ms).
  const char* as_env = getenv("AEFFT_TS_ASMEM");
  const bool want_asmem = stack_needs_asmem || (as_env ? as_env[0] != '0' : p.NR == 1);
  const bool narrow_ok = as_env && as_env[0] == '2';  // also try 64-pixel strips for the shared-memory ring (experiments)
  for (int as = want_asmem ? 1 : 0; as >= 0 && p.asmem < 0; as--) {
    const int PJ_nat = (Ny + halo <= 64) ? 64 : 128;
    for (int PJ = PJ_nat; PJ >= 64 && p.asmem < 0; PJ -= 64) {
      if (PJ != PJ_nat && !(as && narrow_ok)) break;  // keep the natural pitch
      p.PJ = PJ;
      p.TJ = (p.PJ - halo) & ~3;
      p.CPR = p.PJ / TS_CHUNK;
      p.strips = (Ny + p.TJ - 1) / p.TJ;
      p.sb_pitch = p.stack ? (((uint32_t)(p.PJ + 8) * 32 * p.np + 1023) & ~1023u) : (uint32_t)(p.PJ + 8) * 16;
      p.u_slot_bytes = (uint32_t)dM * p.PJ * 4;
      p.s_src_bytes = (uint32_t)(8 * p.np) * (p.PJ + 4) * 4;
      p.s_src_bytes = (p.s_src_bytes + 127) & ~127u;
      p.s_slot_bytes = 2 * p.s_src_bytes;
      sb_bytes = (size_t)(p.stack ? 1 : 2 * p.np) * p.NSB * p.sb_pitch;
      s_bytes = (size_t)TS_NSF * p.s_slot_bytes;
      const size_t fixed = s_bytes + sb_bytes + 4096;
      if (as) {
        // ring of NUB >= RS + 1 positions + the mirror of the first RS - 1; prefer two rows of slack and a 3-deep fp32 ring
        for (int extra = 2; extra >= 1 && p.asmem < 0; extra--)
          for (int NU = 3; NU >= 2 && p.asmem < 0; NU--) {
            const int NUB = p.RS + extra, NUBT = NUB + p.RS - 1;
            const size_t part = (size_t)(p.PJ / 8) * NUBT * dM * 16;
            if (NUB <= TS_MAXUB && (size_t)NU * p.u_slot_bytes + 2 * part + fixed <= budget) {
              p.asmem = 1; p.NUB = NUB; p.NUBT = NUBT; p.NU = NU;
              p.a_lbo = (uint32_t)NUBT * dM * 16; p.a_part = (uint32_t)part;
              ub_bytes = 2 * part;
            }
          }
      } else {
        // TMEM A: a short fp32 TMA ring + the bf16 hi / lo ring of the RS most recent rows the TMEM lanes replicate from;
        // pitch + 16 B so that consecutive channels shift one bank group
        p.ub_pitch = (uint32_t)p.PJ * 2 + 16;
        ub_bytes = (size_t)p.RS * 2 * dM * p.ub_pitch;
        for (int NU = 4; NU >= 2 && p.asmem < 0; NU--)
          if ((size_t)NU * p.u_slot_bytes + ub_bytes + fixed <= budget) {
            p.asmem = 0; p.NU = NU; p.NUB = 0; p.NUBT = 0; p.a_lbo = 0; p.a_part = 0;
          }
      }
    }
  }
  if (p.asmem < 0 || (stack_needs_asmem && p.asmem != 1)) return AEFFT_ERR_UNSUPPORTED;
  if (p.asmem) p.ub_pitch = 0;
  p.off_u = 0;
  p.off_s = (uint32_t)(((size_t)p.NU * p.u_slot_bytes + 1023) & ~(size_t)1023);
  p.off_sb = (uint32_t)((p.off_s + s_bytes + 1023) & ~(size_t)1023);
  p.off_ub = (uint32_t)((p.off_sb + sb_bytes + 1023) & ~(size_t)1023);
  const size_t smem = p.off_ub + ub_bytes + 1024;
  // work split: cpj CTAs per job, every frame cut into `bands` row bands; minimise rounds x rows per band
  int cpj = ctx->sm_count / p.n_jobs;
  if (cpj < 1) cpj = 1;
  {
    const int over = p.RS - 1 + Rmax;
    long long best = -1;
    int best_bands = 1;
    for (int bands = 1; bands <= 32 && bands <= Nx; bands++) {
      const int BR = (Nx + bands - 1) / bands;
      const int nb = (Nx + BR - 1) / BR;
      const long long items = (long long)B * p.strips * nb;
      const long long rounds = (items + cpj - 1) / cpj;
      const long long cost = rounds * (BR + over);
      if (best < 0 || cost < best) { best = cost; best_bands = nb; p.BR = BR; }
    }
    p.bands = best_bands;
    p.BR = (Nx + p.bands - 1) / p.bands;
    p.bands = (Nx + p.BR - 1) / p.BR;
  }
  p.items = (long long)B * p.strips * p.bands;
  if (p.items < cpj) cpj = (int)p.items;
  p.cpj = cpj;
  const int T = win.Nk * win.Nl;
  const long long nC = (long long)dM * dD * T;
  p.n_main = 2 * nC + dM + dD;
  p.n_tot = p.n_main + p.n_jobs;
  float* part;
  AE_TRY(ctx->getT("wgts_part", (size_t)cpj * p.n_tot, &part));
  p.part = part;
  for (int c = 0; c < jobs_per_grad; c++) {
    const int ch0 = c * 8 * p.np, nch = (dD - ch0 < 8 * p.np) ? dD - ch0 : 8 * p.np;
    TsJob& gc = p.job[c];
    TsJob& gf = p.job[jobs_per_grad + c];
    gc.has_s1 = 0; gc.ch0 = ch0; gc.nch = nch; gc.oi = win.ai0; gc.oj = win.aj0; gc.rev = 0; gc.is_gf = 0;
    gc.want_usum = (c == 0); gc.g_off = 0;
    gf.has_s1 = 1; gf.ch0 = ch0; gf.nch = nch;
    gf.oi = -(win.ai0 + win.Nk - 1); gf.oj = -(win.aj0 + win.Nl - 1); gf.rev = 1; gf.is_gf = 1; gf.want_usum = 0; gf.g_off = nC;
    int rc = make_tmap_3d_f32(&gc.u_map, dh, Ny, Nx, (uint64_t)B * dM, 32, 1, dM, true);
    rc |= make_tmap_3d_f32(&gf.u_map, hin, Ny, Nx, (uint64_t)B * dM, 32, 1, dM, true);
    rc |= make_tmap_3d_f32(&gc.s0_map, in, Ny, Nx, (uint64_t)B * dD, p.PJ + 4, 1, nch);
    gc.s1_map = gc.s0_map;
    rc |= make_tmap_3d_f32(&gf.s0_map, out, Ny, Nx, (uint64_t)B * dD, p.PJ + 4, 1, nch);
    rc |= make_tmap_3d_f32(&gf.s1_map, in, Ny, Nx, (uint64_t)B * dD, p.PJ + 4, 1, nch);
    if (rc != 0) return AEFFT_ERR_UNSUPPORTED;
  }
  const bool debug = getenv("AEFFT_TS_DEBUG") != nullptr;
  p.dbg = nullptr;
  p.skip = (debug && getenv("AEFFT_TS_SKIP")) ? atoi(getenv("AEFFT_TS_SKIP")) : 0;
  const size_t n_dbg = (size_t)cpj * p.n_jobs * 16 * 8;
  if (debug) {
    AE_TRY(ctx->getT("wgts_dbg", n_dbg, &p.dbg));
    AE_CUDA(cudaMemsetAsync(p.dbg, 0, n_dbg * sizeof(long long), ctx->stream));
  }
  AE_TRY(ctx->ensure_dyn_smem((const void*)wgrad_ts_kernel<false, false>, smem));
  AE_TRY(ctx->ensure_dyn_smem((const void*)wgrad_ts_kernel<true, false>, smem));
  AE_TRY(ctx->ensure_dyn_smem((const void*)wgrad_ts_kernel<false, true>, smem));
  AE_TRY(ctx->ensure_dyn_smem((const void*)wgrad_ts_kernel<true, true>, smem));
  {
    const double px = (double)B * Nx * Ny;
    ProfScope prof(ctx, "wgrad_ts", 2.0 * 2.0 * px * dM * dD * T, 4.0 * px * (3.0 * dD + 2.0 * dM));
    const dim3 grid(cpj, p.n_jobs);
    if (p.asmem) {
      if (debug) wgrad_ts_kernel<true, true><<<grid, TS_THREADS, smem, ctx->stream>>>(p);
      else wgrad_ts_kernel<false, true><<<grid, TS_THREADS, smem, ctx->stream>>>(p);
    } else {
      if (debug) wgrad_ts_kernel<true, false><<<grid, TS_THREADS, smem, ctx->stream>>>(p);
      else wgrad_ts_kernel<false, false><<<grid, TS_THREADS, smem, ctx->stream>>>(p);
    }
  }
  wgrad_ts_reduce_kernel<<<(unsigned)((p.n_main + 1 + 255) / 256), 256, 0, ctx->stream>>>(part, cpj, p.n_tot, p.n_main, p.n_jobs,
                                                                                          2 * nC, dM, dD, G, GB, GP, SQ);
  ctx->launches += 2;
  AE_CUDA(cudaGetLastError());
  if (debug) {
    // development aid: average cycles each role spent waiting on its two barrier classes
    std::vector<long long> h(n_dbg);
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
    AE_CUDA(cudaMemcpy(h.data(), p.dbg, n_dbg * sizeof(long long), cudaMemcpyDeviceToHost));
    const char* role[4] = {"producer (s_empty, u_empty)", "issuer   (sb_full, a_full)", "S conv   (s_full, sb_empty)",
                           "U conv   (u_full, a_empty)"};
    double acc[4][5] = {};
    int cnt[4] = {};
    for (int c = 0; c < cpj * p.n_jobs; c++)
      for (int w = 0; w < 16; w++) {
        if (w == 1) continue;
        const int r = w == 0 ? 0 : w >= 12 ? 1 : w < 4 ? 2 : 3;
        for (int q = 0; q < 5; q++) acc[r][q] += (double)h[((size_t)c * 16 + w) * 8 + q];
        cnt[r]++;
      }
    fprintf(stderr, "[wgrad_ts] dM=%d dD=%d %dx%d B=%lld asmem=%d PJ=%d RS=%d NR=%d np=%d jobs=%d cpj=%d bands=%d BR=%d NU=%d NSB=%d NA=%d smem=%zu\n",
            dM, dD, Nx, Ny, (long long)B, p.asmem, p.PJ, p.RS, p.NR, p.np, p.n_jobs, cpj, p.bands, p.BR, p.NU, p.NSB, p.NA, smem);
    for (int r = 0; r < 4; r++)
      fprintf(stderr, "[wgrad_ts]   %-30s waitA %9.0f  waitB %9.0f  total %9.0f  work %9.0f + %9.0f cycles\n", role[r],
              acc[r][0] / cnt[r], acc[r][1] / cnt[r], acc[r][2] / cnt[r], acc[r][3] / cnt[r], acc[r][4] / cnt[r]);
    // per job (GC jobs first, then GF): mean / max of the CTAs' total cycles (first issuer warp) -- an imbalance between the
    // two gradients shows up here
    for (int j = 0; j < p.n_jobs; j++) {
      double sum = 0, mx = 0;
      for (int c = 0; c < cpj; c++) {
        const double t = (double)h[(((size_t)j * cpj + c) * 16 + 12) * 8 + 2];
        sum += t; if (t > mx) mx = t;
      }
      fprintf(stderr, "[wgrad_ts]   job %d (%s): CTA total mean %9.0f  max %9.0f cycles\n", j, p.job[j].is_gf ? "GF" : "GC", sum / cpj, mx);
    }
  }
  return AEFFT_OK;
}

}  // namespace aefft
