// Momentum (FFT) space on the device-resident net: forward of the whole stack and training of every pair with the
// activations kept as SPECTRA in HBM.
//
// Reference flow (autoencoder.cpp:131-133, 190-196): autoenc_fft transforms the frame, keeps the activations in frequency
// space across the stack, and -- with fft_l = 1, which training presupposes -- inverse-transforms
// EVERY layer; backprop_fft then forward-transforms the pair's in / out layers again (fft_backproplib.cu:1430-1432) and
// uploads the kernel spectra from the host cache net_cfreq (51-136 MB per layer per frame at config 3, :1434-1435).
// Here the forward keeps every layer's spectrum, the pair training consumes those spectra directly (a C2R followed by an
// R2C is the identity up to fp32 rounding), real-space layers are materialised only where the caller wants to look at
// them (fft_l), and net_cfreq is a lazily computed VIEW of the device-resident kernels (aefft_net_get_cfreq) instead of a
// host cache that has to be kept in sync.
//
// Layout per resolution level: pairs whose channel counts allow the tensor-core contraction (spec_tc.cu) keep their four
// layers bin-major [bin][frame][2 ch]; the others (3-channel image side) keep the reference's [frame][ch][Nx][Nyr].
// Level changes (spectral pooling, resize :87-157) convert where the two sides differ.
#include <cstdio>
#include <vector>

#include "net.cuh"

using namespace aefft;

namespace aefft {

void net_fft_release(aefft_net* net) {
  if (!net) return;
  if (!net->spec.empty() || net->fft_trace) {
    cudaSetDevice(net->ctx->device);
    cudaStreamSynchronize(net->ctx->stream);
  }
  for (auto& s : net->spec)
    if (s.p) cudaFree(s.p);
  net->spec.clear();
  if (net->fft_trace) cudaFree(net->fft_trace);
  net->fft_trace = nullptr;
  net->fft_trace_cap = 0;
}

}  // namespace aefft

namespace {

bool pow2i(int n) { return fft_len_supported(n); }  // even 2^a 3^b 5^c (the name is historical)

// resolution level of layer l (-1: frame resolution): level n holds layers 2n+1, 2n+2, 2N-2-2n, 2N-1-2n
int level_of(int l, int N) {
  if (l == 0 || l == 2 * N) return -1;
  const int m = l <= N ? l : 2 * N - l;  // N = 2P: layers 1..2P on the way down, mirrored on the way up
  return (m - 1) / 2;
}

bool pair_tc(const aefft_net* net, int n) {
  const ConvL& e = net->convs[n];
  // few frames: weight traffic dominates (fft_capi.cu); bin sharding exchanges bins-fastest column slabs
  return net->B >= 16 && net->ctx->shard_world == 1 && spec_tc_eligible(e.dD, e.dM, e.Nk, e.Nl);
}

size_t spec_floats(const aefft_net* net, int l) {
  const LayerL& L = net->layers[l];
  return (size_t)net->B * L.D * L.Nx * (L.Ny / 2 + 1) * 2;
}

int plan(aefft_net* net) {
  const int nl = (int)net->layers.size(), N = (int)net->convs.size();
  if ((int)net->spec.size() == nl) return AEFFT_OK;
  net_fft_release(net);
  AE_ARG(N >= 2);
  for (int l = 0; l < nl; l++) AE_ARG(pow2i(net->layers[l].Nx) && pow2i(net->layers[l].Ny));
  net->spec.resize(nl);
  for (int l = 0; l < nl; l++) {
    const int lev = level_of(l, N);
    net->spec[l].bin_major = lev >= 0 && pair_tc(net, lev);
    AE_CUDA(cudaMalloc((void**)&net->spec[l].p, spec_floats(net, l) * sizeof(float)));
  }
  return AEFFT_OK;
}

// spectrum of layer ls -> spectrum of layer ld across a level change (spectral pooling by `scale`, possibly a layout change)
int move_spec(aefft_net* net, int ls, int ld) {
  aefft_ctx* ctx = net->ctx;
  const LayerL &A = net->layers[ls], &Z = net->layers[ld];
  AE_ARG(A.D == Z.D);
  const SpecL &sa = net->spec[ls], &sz = net->spec[ld];
  const long long B = net->B, R = B * A.D;
  const long long Sa = (long long)A.Nx * (A.Ny / 2 + 1), Sz = (long long)Z.Nx * (Z.Ny / 2 + 1);
  const bool same_res = A.Nx == Z.Nx && A.Ny == Z.Ny;
  if (sa.bin_major == sz.bin_major) {
    if (same_res) {
      AE_CUDA(cudaMemcpyAsync(sz.p, sa.p, spec_floats(net, ls) * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
      return AEFFT_OK;
    }
    if (sa.bin_major) return launch_bm_resize(ctx, 2 * R, A.Nx, A.Ny, Z.Nx, Z.Ny, sa.p, sz.p);
    return launch_spec_resize(ctx, R, A.Nx, A.Ny, Z.Nx, Z.Ny, (const float2*)sa.p, (float2*)sz.p);
  }
  float* tmp;
  if (!sa.bin_major) {  // bins-fastest -> bin-major: pool in the source layout, then transpose
    if (same_res) return launch_to_binmajor(ctx, R, Sa, (const float2*)sa.p, nullptr, (float2*)sz.p);
    AE_TRY(ctx->getT("nf_tmp", (size_t)R * Sz * 2, &tmp));
    AE_TRY(launch_spec_resize(ctx, R, A.Nx, A.Ny, Z.Nx, Z.Ny, (const float2*)sa.p, (float2*)tmp));
    return launch_to_binmajor(ctx, R, Sz, (const float2*)tmp, nullptr, (float2*)sz.p);
  }
  // bin-major -> bins-fastest: transpose ([S][R] -> [R][S] is the same kernel with the roles swapped), then pool
  if (same_res) return launch_to_binmajor(ctx, Sa, R, (const float2*)sa.p, nullptr, (float2*)sz.p);
  AE_TRY(ctx->getT("nf_tmp", (size_t)R * Sa * 2, &tmp));
  AE_TRY(launch_to_binmajor(ctx, Sa, R, (const float2*)sa.p, nullptr, (float2*)tmp));
  return launch_spec_resize(ctx, R, A.Nx, A.Ny, Z.Nx, Z.Ny, (const float2*)tmp, (float2*)sz.p);
}

// conv_k (:162-189) of conv n: spectrum of layer li -> spectrum of layer lo (same level, same layout)
int conv_spec(aefft_net* net, int n, int li, int lo) {
  aefft_ctx* ctx = net->ctx;
  const ConvL& c = net->convs[n];
  const LayerL &A = net->layers[li], &Z = net->layers[lo];
  AE_ARG(A.D == c.dD && Z.D == c.dM && A.Nx == Z.Nx && A.Ny == Z.Ny && net->spec[li].bin_major == net->spec[lo].bin_major);
  AE_ARG(c.Nk <= A.Nx && c.Nl <= A.Ny);
  const long long S = (long long)A.Nx * (A.Ny / 2 + 1);
  const float norm = (float)A.Nx * (float)A.Ny;
  if (net->spec[li].bin_major) {
    // one embedded-spectrum buffer per conv: the training of the pair right after the forward (aefft_net_fft_step) takes
    // them over instead of generating the same spectra again
    float* emb;
    char name[32];
    snprintf(name, sizeof(name), "nf_emb_%d", n);
    AE_TRY(ctx->getT(name, (size_t)4 * c.dM * c.dD * S, &emb));
    AE_TRY(launch_kernel_spectrum_emb(ctx, c.dM, c.dD, c.Nk, c.Nl, A.Nx, A.Ny, 0, 0, c.c, emb));
    net->emb_valid[n] = 1;
    return launch_tc_forward(ctx, S, (int)net->B, c.dD, c.dM, net->spec[li].p, emb, 1.f / (float)c.dM, c.b, norm, nullptr,
                             net->spec[lo].p, nullptr, 0.0, 0, 0, 0);
  }
  float2* kspec;
  float* kimg;
  AE_TRY(ctx->getT("nf_kspec", (size_t)c.dM * c.dD * S, &kspec));
  AE_TRY(ctx->getT("nf_kimg", (size_t)c.dM * c.dD * A.Nx * A.Ny, &kimg));
  AE_TRY(kernel_spectrum_dev(ctx, (int64_t)c.dM * c.dD, c.Nk, c.Nl, A.Nx, A.Ny, c.c, kimg, kspec));
  return launch_spec_contract(ctx, net->B, c.dD, c.dM, S, (const float2*)net->spec[li].p, nullptr, kspec, (int64_t)c.dD * S, S, 0,
                              1.f / (float)c.dM, c.b, norm, (float2*)net->spec[lo].p);
}

// Level 0-style convs (few channels on the image side, bins-fastest spectra) next to a level change, when nobody looks at the
// full-resolution spectrum in between (fft_l <= 0):
//   encoder: conv n (layer li -> li+1) + spectral pooling (li+1 -> li+2): only the bins the crop keeps are computed and the
//            pooled spectrum is written directly -- layer li+1's spectrum is NOT produced;
//   decoder: spectral up-sampling (ls -> ls+1) + conv n (ls+1 -> ls+2): the conv reads the small spectrum and writes the kept
//            bins of its output (the rest is the conv of zeros = zero) -- layer ls+1's spectrum is NOT produced.
// Training never reads the skipped spectra: a pair trains on layers 2n+1 and 2N-1-2n, and only tensor-core pairs reuse the
// forward's hidden spectrum (train_pair_spectra).
bool fusable_conv(const aefft_net* net, int n, int l_conv_in) {
  const ConvL& c = net->convs[n];
  return !net->spec[l_conv_in].bin_major && spec_conv_reg_supported(c.dD, c.dM) && !getenv("AEFFT_NO_FWD_FUSE");
}
int conv_kspec(aefft_net* net, int n, int Nx, int Ny, float2** kspec) {
  aefft_ctx* ctx = net->ctx;
  const ConvL& c = net->convs[n];
  const long long S = (long long)Nx * (Ny / 2 + 1);
  float* kimg;
  AE_TRY(ctx->getT("nf_kspec", (size_t)c.dM * c.dD * S, kspec));
  AE_TRY(ctx->getT("nf_kimg", (size_t)c.dM * c.dD * Nx * Ny, &kimg));
  return kernel_spectrum_dev(ctx, (int64_t)c.dM * c.dD, c.Nk, c.Nl, Nx, Ny, c.c, kimg, *kspec);
}
// encoder: spec[li] --conv n--> (pooled) spec[li + 2]
int conv_then_pool(aefft_net* net, int n, int li) {
  aefft_ctx* ctx = net->ctx;
  const ConvL& c = net->convs[n];
  const LayerL &A = net->layers[li], &Z = net->layers[li + 2];
  AE_ARG(A.D == c.dD && Z.D == c.dM && Z.Nx < A.Nx && Z.Ny < A.Ny);
  float2* kspec;
  AE_TRY(conv_kspec(net, n, A.Nx, A.Ny, &kspec));
  // the pooled spectrum is written in the layout of the next level (bin-major when that level runs on the tensor cores)
  return launch_spec_conv_reg_resized(ctx, net->B, c.dD, c.dM, A.Nx, A.Ny, Z.Nx, Z.Ny, true, net->spec[li + 2].bin_major,
                                      (const float2*)net->spec[li].p, kspec, c.b, (float)A.Nx * (float)A.Ny, 1.f / (float)c.dM,
                                      (float2*)net->spec[li + 2].p);
}
// the same on a tensor-core level: the input spectrum is cropped first (bin-major rows), the kernel spectra are evaluated on
// the kept bins, and the contraction runs on a quarter of the bins.  Only when the pair's training does not want the hidden
// spectrum at full resolution (the Gram loop, spec_gram.cu, reads the pair's in / out spectra only).
int conv_then_pool_tc(aefft_net* net, int n, int li) {
  aefft_ctx* ctx = net->ctx;
  const ConvL& c = net->convs[n];
  const LayerL &A = net->layers[li], &Z = net->layers[li + 2];
  AE_ARG(A.D == c.dD && Z.D == c.dM && Z.Nx < A.Nx && Z.Ny < A.Ny && net->spec[li].bin_major && net->spec[li + 2].bin_major);
  const long long Sz = (long long)Z.Nx * (Z.Ny / 2 + 1);
  float *xs, *emb;
  AE_TRY(ctx->getT("nf_tmp", (size_t)net->B * 2 * c.dD * Sz, &xs));
  AE_TRY(ctx->getT("nf_emb_pool", (size_t)4 * c.dM * c.dD * Sz, &emb));
  const int rc = launch_kernel_spectrum_emb_pooled(ctx, c.dM, c.dD, c.Nk, c.Nl, A.Nx, A.Ny, Z.Nx, Z.Ny, c.c, emb);
  if (rc != AEFFT_OK) return rc;
  AE_TRY(launch_bm_resize(ctx, 2 * net->B * c.dD, A.Nx, A.Ny, Z.Nx, Z.Ny, net->spec[li].p, xs));
  return launch_tc_forward(ctx, Sz, (int)net->B, c.dD, c.dM, xs, emb, 1.f / (float)c.dM, c.b, (float)A.Nx * (float)A.Ny, nullptr,
                           net->spec[li + 2].p, nullptr, 0.0, 0, 0, 0);
}
// decoder: (small) spec[ls] --up-sampling, conv n--> spec[ls + 2]
int unpool_then_conv(aefft_net* net, int n, int ls) {
  aefft_ctx* ctx = net->ctx;
  const ConvL& c = net->convs[n];
  const LayerL &A = net->layers[ls], &Z = net->layers[ls + 2];
  AE_ARG(A.D == c.dD && Z.D == c.dM && A.Nx < Z.Nx && A.Ny < Z.Ny && !net->spec[ls + 2].bin_major);
  float2* kspec;
  AE_TRY(conv_kspec(net, n, Z.Nx, Z.Ny, &kspec));
  return launch_spec_conv_reg_resized(ctx, net->B, c.dD, c.dM, Z.Nx, Z.Ny, A.Nx, A.Ny, false, net->spec[ls].bin_major,
                                      (const float2*)net->spec[ls].p, kspec, c.b, (float)Z.Nx * (float)Z.Ny, 1.f / (float)c.dM,
                                      (float2*)net->spec[ls + 2].p);
}

// ---- The decoder on the support of the innermost level (fft_l <= 0).  A spectral up-sampling only adds zeros, and conv_k is
// per bin, so every decoder spectrum is non-zero only on the bins that came up from the innermost resolution (sNx x sNy): the
// decoder convs run on that grid alone (kernel spectra evaluated at the level's frequencies of those bins), their outputs are
// stored COMPACT (SpecL::sNx / sNy), the up-sampled layers are never produced, the reconstruction's inverse transform embeds
// from the support grid, and the statistics pass of the Gram loop reads the compact `out` spectrum through the same map.
static int ilog2i(int n) { int l = 0; while ((1 << l) < n) l++; return l; }
bool decoder_support_capable(const aefft_net* net, int fft_l) {
  if (fft_l > 0 || getenv("AEFFT_NO_FWD_FUSE") || getenv("AEFFT_NO_SPARSE_DECODER") || net->ctx->shard_world != 1) return false;
  const int N = (int)net->convs.size(), P = N / 2;
  if (N - 1 == P) return false;
  const LayerL& S0 = net->layers[2 * P + 1];
  if (S0.Nx < 4 || S0.Ny < 4) return false;
  for (int n = P + 1; n < N; n++) {
    const LayerL &A = net->layers[2 * n - 1], &Z = net->layers[2 * n + 1];
    const ConvL& c = net->convs[n];
    if (!(A.Nx < Z.Nx && A.Ny < Z.Ny)) return false;
    if (net->spec[2 * n + 1].bin_major) {
      if (!net->spec[2 * n - 1].bin_major || c.Nk != c.Nl || !(c.Nk == 3 || c.Nk == 5 || c.Nk == 7)) return false;
    } else if (!spec_conv_reg_supported(c.dD, c.dM)) {
      return false;
    }
  }
  if (fft_l == 0) {
    const LayerL& Zf = net->layers[2 * N];
    if (!(S0.Nx < Zf.Nx && S0.Ny < Zf.Ny) || getenv("AEFFT_FFT_V1") || getenv("AEFFT_NO_FFT_POOL")) return false;
    auto p2 = [](int n) { return n > 0 && (n & (n - 1)) == 0; };  // the embedding inverse transform has power-of-two kernels only
    if (!p2(Zf.Nx) || !p2(Zf.Ny) || !p2(S0.Nx) || !p2(S0.Ny)) return false;
    const int lx = ilog2i(Zf.Nx), ly = ilog2i(Zf.Ny);
    if (lx < 3 || lx > 12 || ly < 3 || ly > 12 || net->B * Zf.D > 65535) return false;
  }
  return true;
}
int decoder_on_support(aefft_net* net, int fft_l) {
  aefft_ctx* ctx = net->ctx;
  const int N = (int)net->convs.size(), P = N / 2;
  AE_TRY(conv_spec(net, P, 2 * P, 2 * P + 1));  // the innermost decoder conv: dense at its own resolution
  const int sNx = net->layers[2 * P + 1].Nx, sNy = net->layers[2 * P + 1].Ny;
  const long long Ss = (long long)sNx * (sNy / 2 + 1);
  for (int n = P + 1; n < N; n++) {
    const ConvL& c = net->convs[n];
    const LayerL& Z = net->layers[2 * n + 1];
    const float norm = (float)Z.Nx * (float)Z.Ny;
    if (net->spec[2 * n + 1].bin_major) {
      float* emb;
      AE_TRY(ctx->getT("nf_emb_pool", (size_t)4 * c.dM * c.dD * Ss, &emb));
      AE_TRY(launch_kernel_spectrum_emb_pooled(ctx, c.dM, c.dD, c.Nk, c.Nl, Z.Nx, Z.Ny, sNx, sNy, c.c, emb));
      AE_TRY(launch_tc_forward(ctx, Ss, (int)net->B, c.dD, c.dM, net->spec[2 * n - 1].p, emb, 1.f / (float)c.dM, c.b, norm, nullptr,
                               net->spec[2 * n + 1].p, nullptr, 0.0, 0, 0, 0));
    } else {
      float2* kspec;
      AE_TRY(conv_kspec(net, n, Z.Nx, Z.Ny, &kspec));
      AE_TRY(launch_spec_conv_reg_support(ctx, net->B, c.dD, c.dM, Z.Nx, Z.Ny, sNx, sNy, net->spec[2 * n - 1].bin_major, false,
                                          (const float2*)net->spec[2 * n - 1].p, kspec, c.b, norm, 1.f / (float)c.dM,
                                          (float2*)net->spec[2 * n + 1].p));
    }
    net->spec[2 * n].skipped = true;
    net->spec[2 * n + 1].sNx = sNx;
    net->spec[2 * n + 1].sNy = sNy;
  }
  if (fft_l < 0) return AEFFT_OK;
  // reconstruction: fft_inv (:1373) of the last layer's spectrum, embedded from the support grid on the fly
  const LayerL &A = net->layers[2 * N - 1], &Zf = net->layers[2 * N];
  const size_t R = (size_t)net->B * A.D;
  float2 *tmp, *src = (float2*)net->spec[2 * N - 1].p;
  AE_TRY(ctx->getT("nf_fft_tmp", R * Zf.Nx * (sNy / 2 + 1), &tmp));
  if (net->spec[2 * N - 1].bin_major) {
    AE_TRY(ctx->getT("nf_ff", R * Ss, &src));
    AE_TRY(launch_to_binmajor(ctx, Ss, (long long)R, (const float2*)net->spec[2 * N - 1].p, nullptr, src));
  }
  return launch_fft_c2r_embedded(ctx, (int64_t)R, Zf.Nx, Zf.Ny, sNx, sNy, src, tmp, Zf.p, 1.f / ((float)Zf.Nx * (float)Zf.Ny));
}
// dense copy of a compact decoder spectrum (the consumers that do not understand the support map)
int densify(aefft_net* net, int l, float** dense) {
  aefft_ctx* ctx = net->ctx;
  const LayerL& L = net->layers[l];
  const SpecL& sp = net->spec[l];
  const long long R = net->B * L.D;
  AE_TRY(ctx->getT("nf_dense", spec_floats(net, l), dense));
  if (sp.bin_major) return launch_bm_resize(ctx, 2 * R, sp.sNx, sp.sNy, L.Nx, L.Ny, sp.p, *dense);
  return launch_spec_resize(ctx, R, sp.sNx, sp.sNy, L.Nx, L.Ny, (const float2*)sp.p, (float2*)*dense);
}

// fft_inv (:806-864): spectrum of layer l -> real layer l, scaled by 1/(Nx Ny)
int materialise(aefft_net* net, int l) {
  aefft_ctx* ctx = net->ctx;
  const LayerL& L = net->layers[l];
  const long long R = net->B * L.D, S = (long long)L.Nx * (L.Ny / 2 + 1);
  const float2* spec = (const float2*)net->spec[l].p;
  float2* work;
  AE_TRY(ctx->getT("nf_work", (size_t)R * S, &work));
  if (net->spec[l].bin_major) {
    float2* ff;
    AE_TRY(ctx->getT("nf_ff", (size_t)R * S, &ff));
    AE_TRY(launch_to_binmajor(ctx, S, R, spec, nullptr, ff));
    spec = ff;
  }
  return launch_fft_c2r(ctx, R, L.Nx, L.Ny, spec, work, L.p, 1.f / ((float)L.Nx * (float)L.Ny));
}

int forward(aefft_net* net, int loc, const float* frames, int fft_l) {
  aefft_ctx* ctx = net->ctx;
  AE_CUDA(cudaSetDevice(ctx->device));
  AE_TRY(plan(net));
  const int N = (int)net->convs.size();
  LayerL& L0 = net->layers[0];
  const size_t n0 = (size_t)net->B * L0.D * L0.Nx * L0.Ny;
  if (frames && frames != L0.p) {
    AE_CUDA(cudaMemcpyAsync(L0.p, frames, n0 * sizeof(float),
                            loc == AEFFT_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
    if (loc == AEFFT_HOST) AE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  // fft() :764-801 + the first pool_fft :1346.  When the first level pools (crop), the transform writes the pooled spectrum
  // directly: the full-resolution spectrum of the frames is read by nothing else (fft_kernels.cu: launch_fft_r2c_pooled).
  bool pooled0 = false;
  {
    const LayerL& L1 = net->layers[1];
    if (L1.Nx < L0.Nx && L1.Ny < L0.Ny) {
      const size_t R = (size_t)net->B * L0.D, S1 = (size_t)L1.Nx * (L1.Ny / 2 + 1);
      float2 *tmp, *dst = (float2*)net->spec[1].p;
      AE_TRY(ctx->getT("nf_fft_tmp", R * L0.Nx * (L1.Ny / 2 + 1), &tmp));
      if (net->spec[1].bin_major) AE_TRY(ctx->getT("nf_ff", R * S1, &dst));
      const int rc = launch_fft_r2c_pooled(ctx, (int64_t)R, L0.Nx, L0.Ny, L1.Nx, L1.Ny, L0.p, tmp, dst);
      if (rc == AEFFT_OK) {
        pooled0 = true;
        if (net->spec[1].bin_major) AE_TRY(launch_to_binmajor(ctx, (long long)R, (long long)S1, dst, nullptr, (float2*)net->spec[1].p));
      } else if (rc != AEFFT_ERR_UNSUPPORTED) {
        return rc;
      }
    }
  }
  if (!pooled0) AE_TRY(launch_fft_r2c(ctx, net->B * L0.D, L0.Nx, L0.Ny, L0.p, (float2*)net->spec[0].p));
  for (auto& sp : net->spec) { sp.skipped = false; sp.sNx = sp.sNy = 0; }
  net->emb_valid.assign(net->convs.size(), 0);
  const bool sparse_decoder = decoder_support_capable(net, fft_l);
  bool next_in_done = false;  // the previous iteration already produced this conv's input (encoder) / output (decoder)
  for (int n = 0; n < N; n++) {
    if (n == N / 2 && sparse_decoder) return decoder_on_support(net, fft_l);
    if (n < N / 2) {
      if (!(n == 0 && pooled0) && !next_in_done) AE_TRY(move_spec(net, 2 * n, 2 * n + 1));        // pool_fft :1346
      next_in_done = false;
      if (fft_l > 0) AE_TRY(materialise(net, 2 * n + 1));
      const LayerL &Lc = net->layers[2 * n + 2], &Ln = net->layers[2 * n + 3];
      if (fft_l <= 0 && n + 1 < N / 2 && Ln.Nx < Lc.Nx && Ln.Ny < Lc.Ny && Ln.Nx >= 2 && Ln.Ny >= 2) {
        if (fusable_conv(net, n, 2 * n + 1)) {
          AE_TRY(conv_then_pool(net, n, 2 * n + 1));  // conv_fft :1356 + the next pool_fft :1346, on the kept bins only
          net->spec[2 * n + 2].skipped = true;
          next_in_done = true;
          continue;
        }
        const ConvL& cc = net->convs[n];
        if (net->spec[2 * n + 1].bin_major && net->spec[2 * n + 3].bin_major && !getenv("AEFFT_NO_FWD_FUSE") &&
            spec_gram_loop_pays((int)net->B, cc.dD, cc.dM, true)) {
          const int rc = conv_then_pool_tc(net, n, 2 * n + 1);
          if (rc == AEFFT_OK) {
            net->spec[2 * n + 2].skipped = true;
            next_in_done = true;
            continue;
          }
          if (rc != AEFFT_ERR_UNSUPPORTED) return rc;
        }
      }
      AE_TRY(conv_spec(net, n, 2 * n + 1, 2 * n + 2));  // conv_fft :1356
      if (fft_l > 0) AE_TRY(materialise(net, 2 * n + 2));
    } else {
      if (!next_in_done) AE_TRY(conv_spec(net, n, 2 * n, 2 * n + 1));
      next_in_done = false;
      if (fft_l > 0) AE_TRY(materialise(net, 2 * n + 1));
      if (n == N - 1) {
        // last up-sampling (pool_fft :1360) + fft_inv (:1373): the reconstruction is the only reader of the embedded
        // full-resolution spectrum, so the inverse transform takes the small spectrum and embeds on the fly; with
        // fft_l < 0 nobody wants the reconstruction and neither is done
        if (fft_l < 0) break;
        const LayerL &A = net->layers[2 * n + 1], &Z = net->layers[2 * n + 2];
        if (A.Nx < Z.Nx && A.Ny < Z.Ny) {
          const size_t R = (size_t)net->B * A.D, Sa = (size_t)A.Nx * (A.Ny / 2 + 1);
          float2 *tmp, *src = (float2*)net->spec[2 * n + 1].p;
          AE_TRY(ctx->getT("nf_fft_tmp", R * Z.Nx * (A.Ny / 2 + 1), &tmp));
          if (net->spec[2 * n + 1].bin_major) {
            AE_TRY(ctx->getT("nf_ff", R * Sa, &src));
            AE_TRY(launch_to_binmajor(ctx, (long long)Sa, (long long)R, (const float2*)net->spec[2 * n + 1].p, nullptr, src));
          }
          const int rc = launch_fft_c2r_embedded(ctx, (int64_t)R, Z.Nx, Z.Ny, A.Nx, A.Ny, src, tmp, Z.p,
                                                 1.f / ((float)Z.Nx * (float)Z.Ny));
          if (rc == AEFFT_OK) break;
          if (rc != AEFFT_ERR_UNSUPPORTED) return rc;
        }
        AE_TRY(move_spec(net, 2 * n + 1, 2 * n + 2));
        AE_TRY(materialise(net, 2 * n + 2));
        break;
      }
      {
        const LayerL &As = net->layers[2 * n + 1], &Zb = net->layers[2 * n + 2];
        if (fft_l <= 0 && As.Nx < Zb.Nx && As.Ny < Zb.Ny && As.Nx >= 2 && As.Ny >= 2 && !net->spec[2 * n + 3].bin_major &&
            fusable_conv(net, n + 1, 2 * n + 2)) {
          AE_TRY(unpool_then_conv(net, n + 1, 2 * n + 1));  // pool_fft :1360 + the next conv_fft on the embedded bins only
          next_in_done = true;
          continue;
        }
      }
      AE_TRY(move_spec(net, 2 * n + 1, 2 * n + 2));     // pool_fft :1360
      if (fft_l > 0) AE_TRY(materialise(net, 2 * n + 2));
    }
  }
  return AEFFT_OK;
}

}  // namespace

extern "C" {

int aefft_net_fft_forward(aefft_net* net, int loc, const float* frames, int fft_l) {
  AE_ARG(net && net->convs.size() >= 2);
  return forward(net, loc, frames, fft_l);
}

// backprop_fft (:1381-1511) of pair n_l on the spectra the last aefft_net_fft_forward / _step left in HBM
// (autoencoder.cpp:190-196: in = layers[2n+1], out = layers[size-2-2n], c = net_c[n], f = net_c[N-1-n]).
// fresh_forward: the hidden-layer spectrum of the last forward was computed with the pair's CURRENT kernels (true inside
// aefft_net_fft_step, where every pair is trained once right after the forward; a caller of aefft_net_fft_train_pair may
// train the same pair repeatedly on one forward, so it is recomputed there)
static int train_pair_spectra(aefft_net* net, int n, float del0, int maxdiff, int n_iter, float* trace_dev, float* trace_host,
                              bool fresh_forward) {
  aefft_ctx* ctx = net->ctx;
  const int N = (int)net->convs.size();
  AE_ARG((int)net->spec.size() == (int)net->layers.size());  // a forward has planned and filled the spectra
  const ConvL &e = net->convs[n], &d = net->convs[N - 1 - n];
  const int li = 2 * n + 1, lo = 2 * N - 1 - 2 * n;
  const LayerL& L = net->layers[li];
  FftTrainInputs inp;
  inp.resident = true;
  inp.trace_dev = trace_dev;
  const int W = ctx->shard_world;
  if (W > 1) {
    // Frequency-bin sharding (BASELINE config 4).  The forward ran data parallel: this rank holds the FULL spectra of its own
    // net->B frames.  Training wants the opposite split -- this rank's column slab of ALL W*B frames -- so the pair's in /
    // out spectra are cut into W column slabs and exchanged (one all-to-all over NVSwitch per spectrum: the transpose step of
    // a slab-decomposed transform); the received blocks, ordered by source rank, ARE the frame-major slab
    // [W*B][dD][Nx][ncols] that backprop_fft's sharded form consumes.  Kernels / biases stay replicated: the partial
    // kernel-space gradient blocks are all-reduced inside backprop_fft_run.
    AE_ARG(ctx->comm_world == W && ctx->comm_rank == ctx->shard_rank && !net->spec[li].bin_major);
    const int Nyr = L.Ny / 2 + 1, me = ctx->shard_rank;
    const int64_t img = net->B * e.dD;  // images per rank
    std::vector<int64_t> scount(W), soff(W), rcount(W), roff(W);
    int64_t so = 0;
    const int my_c0 = (int)((long long)me * Nyr / W), my_nc = (int)((long long)(me + 1) * Nyr / W) - my_c0;
    for (int r = 0; r < W; r++) {
      const int c0 = (int)((long long)r * Nyr / W), nc = (int)((long long)(r + 1) * Nyr / W) - c0;
      scount[r] = 2 * img * L.Nx * nc;  // floats
      soff[r] = so;
      so += scount[r];
      rcount[r] = 2 * img * L.Nx * my_nc;
      roff[r] = (int64_t)r * rcount[r];
    }
    float *send, *rx, *ro;
    AE_TRY(ctx->getT("nf_a2a_send", (size_t)so, &send));
    AE_TRY(ctx->getT("nf_a2a_X", (size_t)W * rcount[0], &rx));
    AE_TRY(ctx->getT("nf_a2a_O", (size_t)W * rcount[0], &ro));
    for (int which = 0; which < 2; which++) {
      const float2* full = (const float2*)net->spec[which ? lo : li].p;
      for (int r = 0; r < W; r++) {
        const int c0 = (int)((long long)r * Nyr / W), nc = (int)((long long)(r + 1) * Nyr / W) - c0;
        AE_TRY(launch_spec_slab(ctx, img, L.Nx, L.Ny, full, (float2*)(send + soff[r]), c0, nc));
      }
      AE_TRY(comm_alltoallv(ctx, send, scount.data(), soff.data(), which ? ro : rx, rcount.data(), roff.data()));
    }
    inp.Xs = (const float2*)rx;
    inp.Os = (const float2*)ro;
    return backprop_fft_run(ctx, AEFFT_DEVICE, net->B * W, e.dD, e.dM, L.Nx, L.Ny, e.Nk, e.Nl, inp, nullptr, e.c, nullptr, d.c, e.b,
                            d.b, del0, maxdiff, n_iter, trace_host);
  }
  // the pair's `out` spectrum may be compact on the decoder's support grid: the Gram loop reads it through the map, every
  // other path gets a dense copy
  const float* Op = net->spec[lo].p;
  if (net->spec[lo].sNx > 0) {
    const bool bm = net->spec[li].bin_major;
    const bool gram = bm ? spec_gram_loop_pays((int)net->B, e.dD, e.dM, true)
                         : (!(spec_tc_eligible(e.dD, e.dM, e.Nk, e.Nl) && net->B >= 16) && spec_small_eligible(e.dD, e.dM) &&
                            spec_gram_loop_pays((int)net->B, e.dD, e.dM, false));
    if (gram) { inp.o_sNx = net->spec[lo].sNx; inp.o_sNy = net->spec[lo].sNy; }
    else { float* dense; AE_TRY(densify(net, lo, &dense)); Op = dense; }
  }
  if (net->spec[li].bin_major) {
    inp.Xbm = net->spec[li].p; inp.Obm = Op;
    if (fresh_forward) {
      // what the forward just computed with the kernels as they are now: the hidden spectrum hin = conv_k(in; c, b) and the
      // embedded spectra of c and f (the encoder side only when that conv ran at full resolution)
      char name[32];
      const size_t ne = (size_t)4 * e.dM * e.dD * L.Nx * (L.Ny / 2 + 1);
      float* buf = nullptr;
      if (!net->spec[li + 1].skipped) inp.Hbm = net->spec[li + 1].p;
      if ((int)net->emb_valid.size() == N && net->emb_valid[n]) {
        snprintf(name, sizeof(name), "nf_emb_%d", n);
        AE_TRY(ctx->getT(name, ne, &buf));
        inp.Cemb = buf;
      }
      if ((int)net->emb_valid.size() == N && net->emb_valid[N - 1 - n]) {
        snprintf(name, sizeof(name), "nf_emb_%d", N - 1 - n);
        AE_TRY(ctx->getT(name, ne, &buf));
        inp.Femb = buf;
      }
    }
  } else { inp.Xs = (const float2*)net->spec[li].p; inp.Os = (const float2*)Op; }
  return backprop_fft_run(ctx, AEFFT_DEVICE, net->B, e.dD, e.dM, L.Nx, L.Ny, e.Nk, e.Nl, inp, nullptr, e.c, nullptr, d.c, e.b, d.b,
                          del0, maxdiff, n_iter, trace_host);
}

int aefft_net_fft_train_pair(aefft_net* net, int n_l, float del0, int maxdiff, int n_iter, float* mse_trace) {
  AE_ARG(net && n_l >= 0 && n_l < (int)net->pairs.size() && n_iter >= 1);
  AE_CUDA(cudaSetDevice(net->ctx->device));
  return train_pair_spectra(net, n_l, del0, maxdiff, n_iter, nullptr, mse_trace, false);
}

int aefft_net_fft_step(aefft_net* net, int loc, const float* frames, float del0, int maxdiff, int n_iter, int fft_l,
                       float* mse) {
  AE_ARG(net && net->convs.size() >= 2 && n_iter >= 1);
  aefft_ctx* ctx = net->ctx;
  AE_TRY(forward(net, loc, frames, fft_l));
  const int N = (int)net->convs.size(), P = N / 2;
  const int64_t tlen = (int64_t)P * (n_iter + 1);
  if (net->fft_trace_cap < tlen) {
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
    if (net->fft_trace) cudaFree(net->fft_trace);
    net->fft_trace = nullptr;
    AE_CUDA(cudaMalloc((void**)&net->fft_trace, (size_t)tlen * sizeof(float)));
    net->fft_trace_cap = tlen;
  }
  for (int n = 0; n < P; n++)
    AE_TRY(train_pair_spectra(net, n, del0, maxdiff, n_iter, net->fft_trace + (size_t)n * (n_iter + 1), nullptr, true));
  if (mse) {
    AE_CUDA(cudaMemcpyAsync(mse, net->fft_trace, (size_t)tlen * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return AEFFT_OK;
}

// net_cfreq[n] as a view of the device-resident kernels: R2C(kernel_pad(net_c[n])) at the resolution conv n runs at, in the
// reference's interleaved wire format (store_cfreq :1117-1127).  n_floats must equal 2*dM*dD*Nx*(Ny/2+1).
int aefft_net_get_cfreq(aefft_net* net, int n, float* cfreq, int64_t n_floats) {
  AE_ARG(net && cfreq && n >= 0 && n < (int)net->convs.size());
  aefft_ctx* ctx = net->ctx;
  AE_CUDA(cudaSetDevice(ctx->device));
  const int N = (int)net->convs.size();
  const ConvL& c = net->convs[n];
  const LayerL& L = net->layers[n < N / 2 ? 2 * n + 1 : 2 * n];
  const size_t S = (size_t)L.Nx * (L.Ny / 2 + 1), want = 2 * (size_t)c.dM * c.dD * S;
  AE_ARG((size_t)n_floats == want && pow2i(L.Nx) && pow2i(L.Ny));
  float2* kspec;
  float* kimg;
  AE_TRY(ctx->getT("nf_kspec", (size_t)c.dM * c.dD * S, &kspec));
  AE_TRY(ctx->getT("nf_kimg", (size_t)c.dM * c.dD * L.Nx * L.Ny, &kimg));
  AE_TRY(kernel_spectrum_dev(ctx, (int64_t)c.dM * c.dD, c.Nk, c.Nl, L.Nx, L.Ny, c.c, kimg, kspec));
  AE_CUDA(cudaMemcpyAsync(cfreq, kspec, want * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  AE_CUDA(cudaStreamSynchronize(ctx->stream));
  return AEFFT_OK;
}

}  // extern "C"
