// Device-resident network state shared by net.cu (coordinate space) and net_fft.cu (momentum space).
#pragma once
#include <vector>

#include "common.cuh"

namespace aefft {

struct ConvL {
  int dM, dD, Nk, Nl, scale;
  float *c = nullptr, *b = nullptr;  // device
};
struct LayerL {
  int D, Nx, Ny;
  float* p = nullptr;  // [B][D][Nx][Ny]
};
// per-pair momentum / last-gradient state.  The reference shares ONE set (dc,db,df,dp,ddc,..) between all pairs and
// zeroes it whenever the active pair changes (autoencoder.cpp:288-292, 412-417, 447-452); here each pair owns its set
// (aefft_net_reset_momentum reproduces the zeroing).
struct PairState {
  float *dc = nullptr, *db = nullptr, *df = nullptr, *dp = nullptr;
  float *ddc = nullptr, *ddb = nullptr, *ddf = nullptr, *ddp = nullptr;
  float* gbuf = nullptr;  // view into aefft_net::gall (the fused gradient block of all pairs), valid for gbuf_mode
  int64_t gbuf_len = 0;
  int gbuf_mode = -1;
};

// Momentum-space state of one layer (net_fft.cu): its half spectrum for all B frames, either bins-fastest
// [B][D][Nx][Ny/2+1] complex (the reference's layout) or bin-major [bin][B][2 D] (tensor-core pairs, spec_tc.cu).
struct SpecL {
  float* p = nullptr;
  int bin_major = 0;
  bool skipped = false;  // the last forward (fft_l <= 0) fused this layer away: its spectrum was not produced
  int sNx = 0, sNy = 0;  // > 0: the spectrum is stored COMPACT on the (sNx, sNy) grid of the innermost level (decoder layers of
                         // a forward with fft_l <= 0: everything an up-sampling adds is zero); 0: dense
};

}  // namespace aefft

struct aefft_net {
  aefft_ctx* ctx;
  int64_t B;
  std::vector<aefft::LayerL> layers;   // 2*convs+1
  std::vector<aefft::ConvL> convs;     // encoder convs 0..P-1, decoder convs P..2P-1 (pair n: convs n and N-1-n)
  std::vector<aefft::PairState> pairs; // index = pair
  float* mse_dev = nullptr;     // [64]
  // Raw gradient blocks of ALL pairs in one contiguous buffer [pair 0 | pair 1 | ...] (layout of `gall_mode`): a
  // data-parallel step all-reduces it ONCE (the pairs are independent given the forward's activations).
  float* gall = nullptr;
  int64_t gall_len = 0, gall_cap = 0;
  int gall_mode = -1;
  // momentum space (net_fft.cu): per-layer spectra, planned lazily for the current topology
  std::vector<aefft::SpecL> spec;
  std::vector<char> emb_valid;  // per conv: the last forward left the full-resolution embedded kernel spectra in "nf_emb_<n>"
  float* fft_trace = nullptr;  // device [pairs][n_iter+1] mse values of the last fft step
  int64_t fft_trace_cap = 0;
};

namespace aefft {
void net_fft_release(aefft_net* net);
}
using aefft::ConvL;
using aefft::LayerL;
using aefft::PairState;

