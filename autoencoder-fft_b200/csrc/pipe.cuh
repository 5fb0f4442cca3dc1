// Small device helpers shared by the streaming (TMA -> converter -> tcgen05) kernels: bf16 hi/lo splitting, ring
// positions for mbarrier-guarded slot rings, single-lane election, timed barrier waits.
#pragma once
#include "umma.cuh"

namespace aefft {

using namespace umma;

__device__ __forceinline__ uint32_t cvt_pack_bf16(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
// (a, b) -> packed hi parts and packed lo parts (x ~= hi + lo)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = cvt_pack_bf16(a, b);
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
  lo = cvt_pack_bf16(ra, rb);
}

// one lane of a converged warp (warp-uniform predicate)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// position in a ring of n barrier-guarded slots: slot index + phase parity of the current lap
struct Ring {
  int slot = 0, n;
  uint32_t phase = 0;
  __device__ __forceinline__ explicit Ring(int n_) : n(n_) {}
  __device__ __forceinline__ void next() {
    if (++slot == n) { slot = 0; phase ^= 1; }
  }
  __device__ __forceinline__ void skip(int k) {
    slot += k;
    while (slot >= n) { slot -= n; phase ^= 1; }
  }
};

// Barrier waits of a pipeline role, two flavours (both trap instead of hanging when an arrival is lost):
//  * HINT = false: probe, short nanosleep, probe ... -- lowest wake-up latency; right when the SM has spare issue slots
//    (wgrad_ts: one CTA per SM, measured 3 % faster than the hinted form);
//  * HINT = true: the probe carries a suspend-time hint, the warp is parked by the hardware until the phase completes (or
//    the hint expires) instead of re-issuing try_wait.  With a dozen waiting warps per CTA and three CTAs per SM the
//    re-issued probes were 40 % of all issued instructions of conv_rs (ncu) on an SM whose issue slots were 71 % busy.
#ifndef AEFFT_WAIT_SLEEP_NS
#define AEFFT_WAIT_SLEEP_NS 20
#endif
#ifndef AEFFT_WAIT_HINT_NS
#define AEFFT_WAIT_HINT_NS 2000
#endif
template <bool HINT>
__device__ __forceinline__ void mbar_wait_role(uint64_t* bar, uint32_t parity) {
  uint32_t done, spins = 0;
  for (;;) {
    if (HINT) {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}\n"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)AEFFT_WAIT_HINT_NS)
          : "memory");
    } else {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}\n"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(parity)
          : "memory");
    }
    if (done) break;
    if (++spins > (1u << 22)) __trap();  // a lost arrival must fail loudly, never hang the GPU
    if (!HINT) __nanosleep(AEFFT_WAIT_SLEEP_NS);
  }
}

// with DBG the cycles spent waiting are accumulated (role-stall attribution, AEFFT_*_DEBUG=1)
template <bool DBG, bool HINT = false>
__device__ __forceinline__ void wait_t(uint64_t* bar, uint32_t parity, long long& acc) {
  if (DBG) {
    const long long t0 = clock64();
    mbar_wait_role<HINT>(bar, parity);
    acc += clock64() - t0;
  } else {
    mbar_wait_role<HINT>(bar, parity);
  }
}

}  // namespace aefft
