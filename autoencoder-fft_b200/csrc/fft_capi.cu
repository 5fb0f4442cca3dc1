// C ABI, momentum (FFT) space: batched transforms, kernel spectra, autoenc_fft and backprop_fft.
// Host orchestration only; the kernels live in fft_kernels.cu and spectral_kernels.cu.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

using namespace aefft;

namespace aefft {

// transform lengths: even 2^a 3^b 5^c (fft_kernels.cu); the name is historical
static bool pow2(int n) { return aefft::fft_len_supported(n); }

// C = R2C(pad(c)) for n_img kernels (StoreLoad_cfreq first-time branch, fft_backproplib.cu:1148-1157; backprop :1274-1282)
// Evaluated directly from the Nk x Nl taps (pruned DFT): mathematically the same spectrum, no padded image, one write.
// Batches above the grid limit and taps beyond the pruned kernel's envelope go through pad_k + R2C.
int kernel_spectrum_dev(aefft_ctx* ctx, int64_t n_img, int Nk, int Nl, int Nx, int Ny, const float* taps, float* img, float2* spec,
                        int col0, int ncols) {
  const bool slab = ncols > 0 && ncols != Ny / 2 + 1;
  if (Nk <= 8 && Nl <= 8 && (slab || !getenv("AEFFT_NO_PRUNED_DFT"))) {
    const int64_t S = (int64_t)Nx * (ncols > 0 ? ncols : Ny / 2 + 1);
    for (int64_t n0 = 0; n0 < n_img; n0 += 65535) {
      const int64_t cnt = n_img - n0 < 65535 ? n_img - n0 : 65535;
      AE_TRY(launch_kernel_spectrum_direct(ctx, cnt, Nx, Ny, Nk, Nl, taps + n0 * Nk * Nl, spec + n0 * S, col0, ncols));
    }
    return AEFFT_OK;
  }
  if (slab) { set_error("bin-sharded kernel spectra need Nk, Nl <= 8"); return AEFFT_ERR_UNSUPPORTED; }
  AE_TRY(launch_pad(ctx, n_img, Nx, Ny, Nk, Nl, taps, img));
  return launch_fft_r2c(ctx, n_img, Nx, Ny, img, spec);
}
// taps = scale * shrink_k(C2R(spec)) for n_img spectra
static int spectrum_taps_dev(aefft_ctx* ctx, int64_t n_img, int Nk, int Nl, int Nx, int Ny, const float2* spec, float2* work,
                             float* img, float* taps, float scale, int col0 = 0, int ncols = 0) {
  const bool slab = ncols > 0 && ncols != Ny / 2 + 1;
  if (Nk <= 8 && Nl <= 8 && (slab || !getenv("AEFFT_NO_PRUNED_DFT"))) {
    const int64_t S = (int64_t)Nx * (ncols > 0 ? ncols : Ny / 2 + 1);
    for (int64_t n0 = 0; n0 < n_img; n0 += 65535) {
      const int64_t cnt = n_img - n0 < 65535 ? n_img - n0 : 65535;
      AE_TRY(launch_spectrum_to_taps(ctx, cnt, Nx, Ny, Nk, Nl, spec + n0 * S, taps + n0 * Nk * Nl, scale, col0, ncols));
    }
    return AEFFT_OK;
  }
  if (slab) { set_error("bin-sharded kernel gradients need Nk, Nl <= 8"); return AEFFT_ERR_UNSUPPORTED; }
  AE_TRY(launch_fft_c2r(ctx, n_img, Nx, Ny, spec, work, img, scale));
  return launch_shrink(ctx, n_img, Nx, Ny, Nk, Nl, img, taps);
}

struct FftPairBufs {
  float2 *X, *Xt, *O, *H, *G, *C, *F, *dCF, *work;
  float *img, *taps, *db, *dp, *Dc, *Df, *Db, *Dp, *div, *mse;
};

}  // namespace aefft

extern "C" {

int aefft_fft_r2c(aefft_ctx* ctx, int loc, int64_t batch, int Nx, int Ny, const float* in, float* spec) {
  AE_ARG(ctx && in && spec && batch > 0 && pow2(Nx) && pow2(Ny));
  AE_CUDA(cudaSetDevice(ctx->device));
  const size_t nin = (size_t)batch * Nx * Ny, nsp = (size_t)batch * Nx * (Ny / 2 + 1) * 2;
  const float* din = in;
  float* dsp = spec;
  if (loc == AEFFT_HOST) {
    float* t;
    AE_TRY(ctx->getT("fft_in", nin, &t));
    AE_CUDA(cudaMemcpyAsync(t, in, nin * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    din = t;
    AE_TRY(ctx->getT("fft_spec", nsp, &dsp));
  }
  AE_TRY(launch_fft_r2c(ctx, batch, Nx, Ny, din, (float2*)dsp));
  if (loc == AEFFT_HOST) {
    AE_CUDA(cudaMemcpyAsync(spec, dsp, nsp * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return AEFFT_OK;
}

int aefft_fft_c2r(aefft_ctx* ctx, int loc, int64_t batch, int Nx, int Ny, const float* spec, float* out) {
  AE_ARG(ctx && spec && out && batch > 0 && pow2(Nx) && pow2(Ny));
  AE_CUDA(cudaSetDevice(ctx->device));
  const size_t nout = (size_t)batch * Nx * Ny, nsp = (size_t)batch * Nx * (Ny / 2 + 1) * 2;
  const float* dsp = spec;
  float* dout = out;
  float* work;
  AE_TRY(ctx->getT("fft_work", nsp, &work));
  if (loc == AEFFT_HOST) {
    float* t;
    AE_TRY(ctx->getT("fft_spec", nsp, &t));
    AE_CUDA(cudaMemcpyAsync(t, spec, nsp * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    dsp = t;
    AE_TRY(ctx->getT("fft_in", nout, &dout));
  }
  AE_TRY(launch_fft_c2r(ctx, batch, Nx, Ny, (const float2*)dsp, (float2*)work, dout, 1.f));
  if (loc == AEFFT_HOST) {
    AE_CUDA(cudaMemcpyAsync(out, dout, nout * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return AEFFT_OK;
}

int aefft_kernel_pad(aefft_ctx* ctx, int loc, int dM, int dD, int Nk, int Nl, int Nx, int Ny, const float* c,
                     float* c_pad) {
  AE_ARG(ctx && c && c_pad && dM > 0 && dD > 0 && Nk > 0 && Nl > 0 && Nk <= Nx && Nl <= Ny);
  AE_CUDA(cudaSetDevice(ctx->device));
  const size_t nC = (size_t)dM * dD * Nk * Nl, nI = (size_t)dM * dD * Nx * Ny;
  const float* dc = c;
  float* di = c_pad;
  if (loc == AEFFT_HOST) {
    float* t;
    AE_TRY(ctx->getT("kp_c", nC, &t));
    AE_CUDA(cudaMemcpyAsync(t, c, nC * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    dc = t;
    AE_TRY(ctx->getT("kp_img", nI, &di));
  }
  AE_TRY(launch_pad(ctx, (int64_t)dM * dD, Nx, Ny, Nk, Nl, dc, di));
  if (loc == AEFFT_HOST) {
    AE_CUDA(cudaMemcpyAsync(c_pad, di, nI * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return AEFFT_OK;
}

int aefft_kernel_spectrum(aefft_ctx* ctx, int loc, int dM, int dD, int Nk, int Nl, int Nx, int Ny, const float* c,
                          float* cfreq) {
  AE_ARG(ctx && c && cfreq && dM > 0 && dD > 0 && Nk > 0 && Nl > 0 && Nk <= Nx && Nl <= Ny && pow2(Nx) && pow2(Ny));
  AE_CUDA(cudaSetDevice(ctx->device));
  const size_t nC = (size_t)dM * dD * Nk * Nl, nI = (size_t)dM * dD * Nx * Ny, nS = (size_t)dM * dD * Nx * (Ny / 2 + 1) * 2;
  const float* dc = c;
  float *img, *ds = cfreq;
  AE_TRY(ctx->getT("kp_img", nI, &img));
  if (loc == AEFFT_HOST) {
    float* t;
    AE_TRY(ctx->getT("kp_c", nC, &t));
    AE_CUDA(cudaMemcpyAsync(t, c, nC * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    dc = t;
    AE_TRY(ctx->getT("kp_spec", nS, &ds));
  }
  AE_TRY(kernel_spectrum_dev(ctx, (int64_t)dM * dD, Nk, Nl, Nx, Ny, dc, img, (float2*)ds));
  if (loc == AEFFT_HOST) {
    AE_CUDA(cudaMemcpyAsync(cfreq, ds, nS * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return AEFFT_OK;
}

// autoenc_fft (fft_backproplib.cu:1331-1376).  Activations stay in frequency space across the whole stack; layers are
// inverse-transformed only where the caller asks (fft_l) -- and, unlike the reference, the inverse transform never
// clobbers the spectrum it reads (cuFFT's multi-dimensional C2R overwrites its input, so with fft_l=1 the reference
// continues from garbage; see DESIGN.md "reference defects").
int aefft_autoenc_fft(aefft_ctx* ctx, int loc, int64_t B, int n_conv, const int* dims, const float* c_all,
                      const int64_t* coff, const float* b_all, const int64_t* boff, const int* scale, int n_layers,
                      const int* ldims, float* layers_all, const int64_t* loff, int64_t lstride, int cfreq_valid,
                      float* cfreq_all, const int64_t* cfoff, int fft_l) {
  AE_ARG(ctx && dims && c_all && coff && b_all && boff && scale && ldims && layers_all && loff);
  AE_ARG(B > 0 && n_conv >= 2 && n_conv % 2 == 0 && n_layers == 2 * n_conv + 1);
  AE_ARG(!cfreq_valid || (cfreq_all && cfoff));
  AE_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // ---- plan: resolutions per conv, buffer sizes
  int D = ldims[0], Nx = ldims[1], Ny = ldims[2];
  AE_ARG(pow2(Nx) && pow2(Ny));
  size_t max_spec = (size_t)B * D * Nx * (Ny / 2 + 1), max_real = (size_t)B * D * Nx * Ny, max_kimg = 0, max_kspec = 0;
  size_t tot_c = 0, tot_b = 0;
  {
    int d = D, nx = Nx, ny = Ny;
    for (int n = 0; n < n_conv; n++) {
      const int dM = dims[4 * n], dD = dims[4 * n + 1], Nk = dims[4 * n + 2], Nl = dims[4 * n + 3];
      AE_ARG(dD == d && scale[n] != 0);
      auto rs = [&](int s) {
        if (s == 1) return;
        float l = s > 0 ? (float)s : -1.f / (float)s;
        nx = (int)(nx / l);
        ny = (int)(ny / l);
      };
      if (n < n_conv / 2) rs(scale[n]);
      AE_ARG(pow2(nx) && pow2(ny) && Nk <= nx && Nl <= ny);
      size_t s1 = (size_t)B * (dM > dD ? dM : dD) * nx * (ny / 2 + 1), r1 = (size_t)B * (dM > dD ? dM : dD) * nx * ny;
      if (s1 > max_spec) max_spec = s1;
      if (r1 > max_real) max_real = r1;
      size_t ki = (size_t)dM * dD * nx * ny, ks = (size_t)dM * dD * nx * (ny / 2 + 1);
      if (ki > max_kimg) max_kimg = ki;
      if (ks > max_kspec) max_kspec = ks;
      if (n >= n_conv / 2) {
        rs(scale[n]);
        size_t s2 = (size_t)B * dM * nx * (ny / 2 + 1), r2 = (size_t)B * dM * nx * ny;
        if (s2 > max_spec) max_spec = s2;
        if (r2 > max_real) max_real = r2;
      }
      d = dM;
      tot_c = (size_t)coff[n] + (size_t)dM * dD * Nk * Nl > tot_c ? (size_t)coff[n] + (size_t)dM * dD * Nk * Nl : tot_c;
      tot_b = (size_t)boff[n] + dM > tot_b ? (size_t)boff[n] + dM : tot_b;
    }
  }
  float2 *fa, *fb, *work, *kspec;
  float *real, *kimg, *dc_all, *db_all;
  AE_TRY(ctx->getT("aef_fa", max_spec, &fa));
  AE_TRY(ctx->getT("aef_fb", max_spec, &fb));
  AE_TRY(ctx->getT("aef_work", max_spec, &work));
  AE_TRY(ctx->getT("aef_real", max_real, &real));
  AE_TRY(ctx->getT("aef_kimg", max_kimg, &kimg));
  AE_TRY(ctx->getT("aef_kspec", max_kspec, &kspec));
  const float* cw = c_all;
  const float* bw = b_all;
  if (loc == AEFFT_HOST) {
    AE_TRY(ctx->getT("aef_c", tot_c, &dc_all));
    AE_TRY(ctx->getT("aef_b", tot_b, &db_all));
    AE_CUDA(cudaMemcpyAsync(dc_all, c_all, tot_c * sizeof(float), cudaMemcpyHostToDevice, st));
    AE_CUDA(cudaMemcpyAsync(db_all, b_all, tot_b * sizeof(float), cudaMemcpyHostToDevice, st));
    cw = dc_all;
    bw = db_all;
  }
  const cudaMemcpyKind k_in = loc == AEFFT_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const cudaMemcpyKind k_out = loc == AEFFT_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  // layer 0 of every frame: transformed in place of the caller's per-frame block when it is device memory, else
  // gathered into a contiguous [B][D][Nx][Ny] block first
  {
    int rc = AEFFT_ERR_UNSUPPORTED;
    if (loc == AEFFT_DEVICE) rc = launch_fft_r2c_strided(ctx, B * D, Nx, Ny, layers_all + loff[0], D, (long long)lstride, fa);
    if (rc == AEFFT_ERR_UNSUPPORTED) {
      const size_t w = (size_t)D * Nx * Ny * sizeof(float);
      AE_CUDA(cudaMemcpy2DAsync(real, w, layers_all + loff[0], (size_t)lstride * sizeof(float), w, (size_t)B, k_in, st));
      AE_TRY(launch_fft_r2c(ctx, B * D, Nx, Ny, real, fa));
    } else if (rc != AEFFT_OK) {
      return rc;
    }
  }
  float2 *freq = fa, *other = fb;
  int l = 1;
  auto emit_layer = [&](const float2* spec, int ch, int nx, int ny) -> int {
    AE_ARG(l < n_layers && ldims[3 * l] == ch && ldims[3 * l + 1] == nx && ldims[3 * l + 2] == ny);
    const float inv = 1.f / ((float)nx * (float)ny);  // fft_inv :831
    if (loc == AEFFT_DEVICE) {
      // inverse transform straight into the caller's per-frame layer block (no gather / scatter copy)
      const int rc = launch_fft_c2r_strided(ctx, B * ch, nx, ny, spec, work, layers_all + loff[l], ch, (long long)lstride, inv);
      if (rc != AEFFT_ERR_UNSUPPORTED) return rc;
    }
    AE_TRY(launch_fft_c2r(ctx, B * ch, nx, ny, spec, work, real, inv));
    const size_t w = (size_t)ch * nx * ny * sizeof(float);
    AE_CUDA(cudaMemcpy2DAsync(layers_all + loff[l], (size_t)lstride * sizeof(float), real, w, w, (size_t)B, k_out, st));
    return AEFFT_OK;
  };
  auto do_resize = [&](int ch, int s) -> int {
    if (s == 1) return AEFFT_OK;
    float lf = s > 0 ? (float)s : -1.f / (float)s;
    const int nxs = (int)(Nx / lf), nys = (int)(Ny / lf);
    AE_TRY(launch_spec_resize(ctx, B * ch, Nx, Ny, nxs, nys, freq, other));
    std::swap(freq, other);
    Nx = nxs;
    Ny = nys;
    return AEFFT_OK;
  };
  for (int n = 0; n < n_conv; n++) {
    const int dM = dims[4 * n], dD = dims[4 * n + 1], Nk = dims[4 * n + 2], Nl = dims[4 * n + 3];
    if (n < n_conv / 2) {
      AE_TRY(do_resize(dD, scale[n]));
      if (fft_l) { AE_TRY(emit_layer(freq, dD, Nx, Ny)); l++; }
    }
    const int64_t S = (int64_t)Nx * (Ny / 2 + 1);
    const size_t nks = (size_t)dM * dD * S;
    if (cfreq_valid) {
      AE_CUDA(cudaMemcpyAsync(kspec, cfreq_all + cfoff[n], nks * sizeof(float2), k_in, st));  // load_cfreq :1131-1141
    } else {
      AE_TRY(kernel_spectrum_dev(ctx, (int64_t)dM * dD, Nk, Nl, Nx, Ny, cw + coff[n], kimg, kspec));
      if (cfreq_all && cfoff)
        AE_CUDA(cudaMemcpyAsync(cfreq_all + cfoff[n], kspec, nks * sizeof(float2), k_out, st));  // store_cfreq :1117-1127
    }
    // conv_k (:162-189): out[m] = sum_d (in[d]/dM) c[m][d] + b[m] Nx Ny at DC
    AE_TRY(launch_spec_contract(ctx, B, dD, dM, S, freq, nullptr, kspec, (int64_t)dD * S, S, 0, 1.f / (float)dM, bw + boff[n],
                                (float)Nx * (float)Ny, other));
    std::swap(freq, other);
    if (fft_l) { AE_TRY(emit_layer(freq, dM, Nx, Ny)); l++; }
    if (n >= n_conv / 2) {
      AE_TRY(do_resize(dM, scale[n]));
      if (fft_l) { AE_TRY(emit_layer(freq, dM, Nx, Ny)); l++; }
    }
    D = dM;
  }
  if (!fft_l) {
    l = n_layers - 1;
    AE_TRY(emit_layer(freq, D, Nx, Ny));
  }
  AE_CUDA(cudaStreamSynchronize(st));
  return AEFFT_OK;
}

}  // extern "C"

// backprop_fft (fft_backproplib.cu:1381-1511) on real-space frames (transformed here, as the reference does) or on spectra
// that already live in HBM (the device-resident net keeps every layer's spectrum: net_fft.cu).
namespace aefft {
int backprop_fft_run(aefft_ctx* ctx, int loc, int64_t B, int dD, int dM, int Nx, int Ny, int Nk, int Nl, const FftTrainInputs& inp,
                     float* cfreq, float* c, float* ffreq, float* f, float* b, float* p, float del0, int maxdiff, int n_iter,
                     float* mse_trace) {
  const float *in = inp.in, *expout = inp.expout, *out = inp.out;
  const int64_t in_fstride = inp.fstride;
  const bool have_real = in != nullptr, have_ff = inp.Xs != nullptr, have_bm = inp.Xbm != nullptr;
  AE_ARG(ctx && c && f && b && p && (int)have_real + (int)have_ff + (int)have_bm == 1);
  AE_ARG(!have_real || (expout && out));
  AE_ARG(!have_ff || inp.Os);
  AE_ARG(!have_bm || (inp.Obm && spec_tc_eligible(dD, dM, Nk, Nl) && !cfreq && !ffreq && n_iter > 0));
  AE_ARG(have_real || loc == AEFFT_DEVICE);
  AE_ARG(B > 0 && dD > 0 && dM > 0 && pow2(Nx) && pow2(Ny) && Nk <= Nx && Nl <= Ny && n_iter >= 0);
  AE_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // frequency-bin sharding (aefft_set_bin_shard): this device owns the spectrum columns [col0, col0+ncols) of every
  // image; every per-bin kernel below then runs on the slab only, the kernel-space gradient block and the mse are
  // partial sums that the gradient hook adds over the devices.  Unsharded: the whole half spectrum.
  const int Nyr = Ny / 2 + 1;
  const int world = ctx->shard_world, srank = ctx->shard_rank;
  const int col0 = (int)((long long)srank * Nyr / world), ncols = (int)((long long)(srank + 1) * Nyr / world) - col0;
  const bool sharded = world > 1;
  if (sharded) {
    AE_ARG(ncols > 0 && !cfreq && !ffreq && loc == AEFFT_DEVICE && !have_bm);
    if (!ctx->grad_hook && ctx->comm_world <= 1) {
      set_error("bin-sharded aefft_backprop_fft needs a communicator (aefft_comm_init) or a gradient hook (sum over devices)");
      return AEFFT_ERR_ARG;
    }
  }
  const bool own_dc = col0 == 0;  // the DC bin (biases, db/dp) lives on the device that owns column 0
  const int64_t S = (int64_t)Nx * ncols;
  const size_t P = (size_t)Nx * Ny, nC = (size_t)dM * dD * Nk * Nl, nKS = (size_t)dM * dD * S;
  const size_t nXs = (size_t)B * dD * S, nHs = (size_t)B * dM * S;
  const cudaMemcpyKind k_in = loc == AEFFT_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const cudaMemcpyKind k_out = loc == AEFFT_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  FftPairBufs q;
  float *real, *wts;  // wts = [c | f | b | p]
  // (with a handful of frames the per-bin weight blocks dominate the traffic and the embedded form doubles them: the
  // bins-fastest kernels are faster there; a bin-major caller has made that choice already)
  const bool use_tc = spec_tc_eligible(dD, dM, Nk, Nl) && !cfreq && !ffreq && n_iter > 0 && (B >= 16 || have_bm);
  q.X = q.Xt = q.O = q.H = q.G = q.C = q.F = q.dCF = q.work = nullptr;
  q.img = real = nullptr;
  if (have_real) {
    AE_TRY(ctx->getT("bpf_X", nXs, &q.X));
    AE_TRY(ctx->getT("bpf_Xt", nXs, &q.Xt));
    AE_TRY(ctx->getT("bpf_O", nXs, &q.O));
    if (loc == AEFFT_HOST) AE_TRY(ctx->getT("bpf_real", (size_t)B * dD * P, &real));
  }
  // <= 4 input channels (the image side): one fused kernel per iteration, H / G / E never touch HBM (spec_small.cu)
  const bool use_small = !use_tc && spec_small_eligible(dD, dM) && n_iter > 0;
  // The autoencoder case expout == in: the whole iteration loop runs on per-bin Gram matrices of the frames, formed in ONE
  // pass (spec_gram.cu); an iteration then touches kernel-spectrum-sized data only.
  const bool same_target = !have_real || expout == in;
  const bool gram_tc = use_tc && same_target && spec_gram_loop_pays((int)B, dD, dM, true);
  const bool gram_ff = use_small && same_target && spec_gram_loop_pays((int)B, dD, dM, false);
  // a compact `out` spectrum (support of an up-sampling) is understood by the statistics pass of the Gram loop only
  AE_ARG(inp.o_sNx == 0 || ((have_bm && gram_tc) || (have_ff && gram_ff)));
  if (!use_tc) {  // the bins-fastest CUDA-core path keeps the gradient spectra, and H, G unless fused
    if (!use_small) {
      AE_TRY(ctx->getT("bpf_H", nHs, &q.H));
      AE_TRY(ctx->getT("bpf_G", nHs, &q.G));
    }
    AE_TRY(ctx->getT("bpf_dCF", 2 * nKS, &q.dCF));
  }
  if (!use_tc || !inp.resident) {
    AE_TRY(ctx->getT("bpf_C", nKS, &q.C));
    AE_TRY(ctx->getT("bpf_F", nKS, &q.F));
    AE_TRY(ctx->getT("bpf_work", 2 * nKS > nXs ? 2 * nKS : nXs, &q.work));
    AE_TRY(ctx->getT("bpf_img", 2 * (size_t)dM * dD * P, &q.img));
  }
  AE_TRY(ctx->getT("bpf_taps", 2 * nC + dM + dD, &q.taps));  // raw gradient block [dck | dfk | db | dp]
  AE_TRY(ctx->getT("bpf_wts", 2 * nC + dM + dD, &wts));
  float* small;
  AE_TRY(ctx->getT("bpf_small", 2 * (size_t)(dM + dD) + 2 * nC + 2 * nC + dM + dD + (size_t)n_iter + 1, &small));
  q.db = q.taps + 2 * nC; q.dp = q.db + dM; q.Db = small + dM + dD; q.Dp = q.Db + dM;
  q.Dc = q.Dp + dD; q.Df = q.Dc + nC; q.div = q.Df + nC; q.mse = q.div + 2 * nC + dM + dD;
  float *dc_w = wts, *df_w = wts + nC, *db_w = wts + 2 * nC, *dp_w = db_w + dM;
  // momentum buffers are zeroed at the start of every call (:1420-1423)
  AE_CUDA(cudaMemsetAsync(q.Db, 0, (2 * (size_t)(dM + dD) - dM - dD + 2 * nC) * sizeof(float), st));
  AE_CUDA(cudaMemcpyAsync(dc_w, c, nC * sizeof(float), k_in, st));  // flatten_kernel :1436-1437
  AE_CUDA(cudaMemcpyAsync(df_w, f, nC * sizeof(float), k_in, st));
  AE_CUDA(cudaMemcpyAsync(db_w, b, dM * sizeof(float), k_in, st));
  AE_CUDA(cudaMemcpyAsync(dp_w, p, dD * sizeof(float), k_in, st));
  // fft(in), fft(expout), fft(out) (:1430-1432)
  auto load_fft = [&](const float* src, float2* dst) -> int {
    const float* d = src;
    if (loc == AEFFT_HOST) {
      AE_CUDA(cudaMemcpyAsync(real, src, (size_t)B * dD * P * sizeof(float), cudaMemcpyHostToDevice, st));
      d = real;
    }
    float2* tgt = dst;
    if (sharded) AE_TRY(ctx->getT("bpf_full", (size_t)B * dD * Nx * Nyr, &tgt));  // full spectrum, then keep the slab
    if (in_fstride) AE_TRY(launch_fft_r2c_strided(ctx, B * dD, Nx, Ny, d, dD, (long long)in_fstride, tgt));
    else AE_TRY(launch_fft_r2c(ctx, B * dD, Nx, Ny, d, tgt));
    if (sharded) AE_TRY(launch_spec_slab(ctx, B * dD, Nx, Ny, tgt, dst, col0, ncols));
    return AEFFT_OK;
  };
  const float2* Xt = nullptr;
  if (have_real) {
    AE_TRY(load_fft(in, q.X));
    Xt = q.X;
    if (expout != in) {
      AE_TRY(load_fft(expout, q.Xt));
      Xt = q.Xt;
    }
    AE_TRY(load_fft(out, q.O));
  } else if (have_ff) {
    q.X = const_cast<float2*>(inp.Xs);  // read only below (sharded: the caller exchanged the slabs already)
    q.O = const_cast<float2*>(inp.Os);
    if (!use_tc && !use_small && n_iter > 0) {  // the generic loop rewrites O in place: work on a copy of the layer spectrum
      AE_TRY(ctx->getT("bpf_O", nXs, &q.O));
      AE_CUDA(cudaMemcpyAsync(q.O, inp.Os, nXs * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    }
    Xt = q.X;
  } else {
    AE_ARG(!sharded);
  }
  // kernel spectra: the caller's cache (load_cfreq :1434-1435) or derived from c,f
  if (!use_tc) {
    if (cfreq) AE_CUDA(cudaMemcpyAsync(q.C, cfreq, nKS * sizeof(float2), k_in, st));
    else AE_TRY(kernel_spectrum_dev(ctx, (int64_t)dM * dD, Nk, Nl, Nx, Ny, dc_w, q.img, q.C, col0, ncols));
    if (ffreq) AE_CUDA(cudaMemcpyAsync(q.F, ffreq, nKS * sizeof(float2), k_in, st));
    else AE_TRY(kernel_spectrum_dev(ctx, (int64_t)dM * dD, Nk, Nl, Nx, Ny, df_w, q.img, q.F, col0, ncols));
  }
  const float norm = (float)Nx * (float)Ny;
  // Reduction of the raw kernel-space block over the devices, before the non-linear clip: the engine's own NCCL
  // all-reduce when the ctx has a communicator (average for data-parallel frames, sum for bin-sharded partial blocks),
  // else the caller's hook.  The mse values feed nothing inside the loop, so their partials stay local until the end of
  // the call and are reduced ONCE as a whole trace -- one collective per iteration instead of three.
  const bool use_comm = ctx->comm_world > 1;
  auto reduce_over_devices = [&](float* block, int64_t n) -> int {
    if (use_comm) return comm_allreduce(ctx, block, n, sharded ? 0 : 1);
    if (ctx->grad_hook && ctx->grad_hook(ctx->grad_hook_user, block, n) != 0) {
      set_error("aefft_backprop_fft: gradient hook failed");
      return AEFFT_ERR_ARG;
    }
    return AEFFT_OK;
  };
  const float* bias_b = own_dc ? db_w : nullptr;
  const float* bias_p = own_dc ? dp_w : nullptr;
  const double mse_scale = 1.0 / ((double)dD * Nx * Ny) / (2.0 * dM * Nx * Ny) / (double)B;
  if (!have_bm && !gram_ff) {
    AE_TRY(launch_spec_mse(ctx, B, dD, dM, Nx, Ny, Xt, q.O, q.mse, col0, ncols));  // "mse fft:" (:1440)
    if (sharded && !use_comm) AE_TRY(reduce_over_devices(q.mse, 1));
  }
  const float del = 0.1f * del0;                                        // :1445
  const double Norm = (double)norm * 2.0 * dM * dD * (double)Nx * Ny;   // :399
  const float gscale = (float)(1.0 / (Norm * (double)B));
  // Tensor-core path (spec_tc.cu): pairs with >= 8 channels on both sides run the five per-bin contractions of an
  // iteration as real GEMMs on tcgen05 over bin-major operands; the frame spectra are transposed once per call, the
  // kernel spectra are generated directly in the embedded bin-major form, the kernel-space gradients are reduced from
  // the bin-major gradient spectra.  Callers that pass spectra caches (cfreq / ffreq) keep the bins-fastest path, whose
  // first iteration consumes those caches as they are.
  if (use_tc) {
    float *Xb = nullptr, *Xtb, *Eb = nullptr, *Hb = nullptr, *Gb = nullptr, *Cemb, *Femb, *dCt, *dFt;
    // Gram form (spec_tc.cu: gram_grad_kernel): both gradient spectra from Mg = sum_b E conj(X); no G, and the hidden
    // spectrum is needed by the re-forward only
    const bool gram = spec_tc_gram_pays((int)B, dD, dM);
    if (!have_bm) AE_TRY(ctx->getT("tc_Xb", 2 * nXs, &Xb));
    if (!gram_tc || !have_bm) AE_TRY(ctx->getT("tc_Eb", 2 * nXs, &Eb));
    if (!gram_tc) AE_TRY(ctx->getT("tc_Hb", 2 * nHs, &Hb));
    if (!gram && !gram_tc) AE_TRY(ctx->getT("tc_Gb", 2 * nHs, &Gb));
    AE_TRY(ctx->getT("tc_Cemb", 4 * nKS, &Cemb));
    AE_TRY(ctx->getT("tc_Femb", 4 * nKS, &Femb));
    AE_TRY(ctx->getT("tc_dCt", 2 * nKS, &dCt));
    AE_TRY(ctx->getT("tc_dFt", 2 * nKS, &dFt));
    if (have_bm) {
      // the layer spectra are bin-major already: E = O - X and the "mse fft:" value in one pass
      Xb = const_cast<float*>(inp.Xbm);
      Xtb = Xb;
      if (!gram_tc) AE_TRY(launch_bm_sub_mse(ctx, S, (long long)B * 2 * dD, inp.Obm, inp.Xbm, Eb, q.mse, mse_scale, ncols, col0, Ny));
    } else {
      AE_TRY(launch_to_binmajor(ctx, (long long)B * dD, S, q.X, nullptr, (float2*)Xb));
      Xtb = Xb;
      if (Xt != q.X) {
        AE_TRY(ctx->getT("tc_Xtb", 2 * nXs, &Xtb));
        AE_TRY(launch_to_binmajor(ctx, (long long)B * dD, S, Xt, nullptr, (float2*)Xtb));
      }
      AE_TRY(launch_to_binmajor(ctx, (long long)B * dD, S, q.O, Xt, (float2*)Eb));  // E = O - Xt of the caller's `out`
    }
    // spectra of the current kernels: generated here unless the caller (the resident net's forward) just did
    const float *Cemb0 = inp.Cemb, *Femb0 = inp.Femb;
    if (!Femb0) {
      AE_TRY(launch_kernel_spectrum_emb(ctx, dD, dM, Nk, Nl, Nx, Ny, col0, ncols, df_w, Femb));
      Femb0 = Femb;
    }
    if (gram_tc) {
      float *Gx, *M0, *dcs;
      AE_TRY(ctx->getT("gr_Gx", (size_t)2 * S * dD * dD, &Gx));
      AE_TRY(ctx->getT("gr_M0", (size_t)2 * S * dD * dD, &M0));
      AE_TRY(ctx->getT("gr_dc", (size_t)4 * dD, &dcs));
      // the one pass over the frames: Gx, M0 (and "mse fft:" when the caller's spectra are bin-major)
      if (have_bm)
        AE_TRY(launch_gram_stats_bm(ctx, S, (int)B, dD, inp.Xbm, inp.Obm, 1, Gx, M0, q.mse, mse_scale, own_dc ? dcs : nullptr, ncols,
                                    col0, Ny, Nx, inp.o_sNx, inp.o_sNy));
      else
        AE_TRY(launch_gram_stats_bm(ctx, S, (int)B, dD, Xb, Eb, 0, Gx, M0, nullptr, 0.0, own_dc ? dcs : nullptr, ncols, col0, Ny));
      if (!Cemb0) {
        AE_TRY(launch_kernel_spectrum_emb(ctx, dM, dD, Nk, Nl, Nx, Ny, col0, ncols, dc_w, Cemb));
        Cemb0 = Cemb;
      }
      const float gb = (float)((double)norm / (Norm * (double)B));
      for (int n = 0; n <= n_iter; n++) {
        // iteration n: mse of the current kernels (n >= 1: what the reference prints after update n, :1463) and the gradients
        const bool last = n == n_iter;
        AE_TRY(launch_gram_iter_bm(ctx, S, (int)B, dD, dM, Gx, M0, n ? Cemb : Cemb0, n ? Femb : Femb0, n == 0, gscale, gb, norm,
                                   own_dc ? dcs : nullptr, bias_b,
                                   bias_p, last ? nullptr : dCt, last ? nullptr : dFt, q.db, q.dp, n ? q.mse + n : nullptr, mse_scale,
                                   ncols, col0, Ny));
        if (n && sharded && !use_comm) AE_TRY(reduce_over_devices(q.mse + n, 1));
        if (last) break;
        if (!own_dc) AE_CUDA(cudaMemsetAsync(q.db, 0, (size_t)(dM + dD) * sizeof(float), st));
        AE_TRY(launch_binmajor_to_taps(ctx, dM, dD, 0, Nk, Nl, Nx, Ny, col0, ncols, (const float2*)dCt, q.taps, 1.f));
        AE_TRY(launch_binmajor_to_taps(ctx, dM, dD, 1, Nk, Nl, Nx, Ny, col0, ncols, (const float2*)dFt, q.taps + nC, 1.f));
        const bool fold_div = sharded && maxdiff;
        if (fold_div) {
          AE_TRY(launch_gradient_diff(ctx, dM, dD, Nk, Nl, dc_w, df_w, db_w, dp_w, q.div, srank, world));
          AE_TRY(launch_axpby(ctx, q.taps, q.div, 1.f, -10.f, (long long)(2 * nC + dM + dD)));
        }
        AE_TRY(reduce_over_devices(q.taps, (int64_t)(2 * nC + dM + dD)));
        AE_TRY(launch_fft_update(ctx, dM, dD, Nk, Nl, dc_w, df_w, db_w, dp_w, q.taps, q.taps + nC, q.db, q.dp, q.Dc, q.Df, q.Db,
                                 q.Dp, del, fold_div ? 0 : maxdiff, q.div));
        AE_TRY(launch_kernel_spectrum_emb(ctx, dM, dD, Nk, Nl, Nx, Ny, col0, ncols, dc_w, Cemb));
        AE_TRY(launch_kernel_spectrum_emb(ctx, dD, dM, Nk, Nl, Nx, Ny, col0, ncols, df_w, Femb));
      }
    } else {
    if (inp.Femb) AE_TRY(launch_kernel_spectrum_emb(ctx, dD, dM, Nk, Nl, Nx, Ny, col0, ncols, df_w, Femb));  // this loop rewrites its own
    // H of the current kernels (the reference recomputes it inside gradient_k_io as H-hat, without the /dM: quirk F1);
    // the resident net hands over the hidden layer its forward just computed with these very kernels
    const float* Hcur = inp.Hbm;
    if (gram) {
      AE_TRY(launch_kernel_spectrum_emb(ctx, dM, dD, Nk, Nl, Nx, Ny, col0, ncols, dc_w, Cemb));
    } else if (!Hcur) {
      AE_TRY(launch_kernel_spectrum_emb(ctx, dM, dD, Nk, Nl, Nx, Ny, col0, ncols, dc_w, Cemb));
      AE_TRY(launch_tc_forward(ctx, S, (int)B, dD, dM, Xb, Cemb, 1.f / (float)dM, bias_b, norm, nullptr, Hb, nullptr, 0.0, 0, 0, 0));
      Hcur = Hb;
    }
    for (int n = 0; n < n_iter; n++) {
      if (gram) {
        AE_TRY(launch_tc_gram_grad(ctx, S, (int)B, dM, dD, Eb, Xb, Cemb, Femb, gscale, dCt, dFt));
        if (own_dc)
          AE_TRY(launch_tc_dc_terms_gram(ctx, (int)B, dM, dD, Eb, Femb, bias_b, dFt, q.db, q.dp,
                                         (float)((double)norm / (Norm * (double)B)), gscale, norm));
        else AE_CUDA(cudaMemsetAsync(q.db, 0, (size_t)(dM + dD) * sizeof(float), st));
      } else {
      AE_TRY(launch_tc_adjoint(ctx, S, (int)B, dD, dM, Eb, Femb, Gb));                               // G = E conj(F)
      AE_TRY(launch_tc_outer(ctx, S, (int)B, dM, dD, Gb, Xb, gscale, 0, dCt));                       // dC[m][d] = G conj(X)
      AE_TRY(launch_tc_outer(ctx, S, (int)B, dM, dD, Hcur, Eb, gscale * (float)dM, 1, dFt));         // dF[d][m] = E conj(dM H) at [m][d]
      Hcur = Hb;  // from now on the re-forward below provides H
      if (own_dc)
        AE_TRY(launch_tc_dc_terms(ctx, (int)B, dM, dD, Gb, Eb, bias_b, dFt, q.db, q.dp, (float)((double)norm / (Norm * (double)B)),
                                  gscale, -(float)(dM - 1) * norm));
      else AE_CUDA(cudaMemsetAsync(q.db, 0, (size_t)(dM + dD) * sizeof(float), st));
      }
      AE_TRY(launch_binmajor_to_taps(ctx, dM, dD, 0, Nk, Nl, Nx, Ny, col0, ncols, (const float2*)dCt, q.taps, 1.f));
      AE_TRY(launch_binmajor_to_taps(ctx, dM, dD, 1, Nk, Nl, Nx, Ny, col0, ncols, (const float2*)dFt, q.taps + nC, 1.f));
      const bool fold_div = sharded && maxdiff;
      if (fold_div) {
        AE_TRY(launch_gradient_diff(ctx, dM, dD, Nk, Nl, dc_w, df_w, db_w, dp_w, q.div, srank, world));
        AE_TRY(launch_axpby(ctx, q.taps, q.div, 1.f, -10.f, (long long)(2 * nC + dM + dD)));
      }
      AE_TRY(reduce_over_devices(q.taps, (int64_t)(2 * nC + dM + dD)));
      AE_TRY(launch_fft_update(ctx, dM, dD, Nk, Nl, dc_w, df_w, db_w, dp_w, q.taps, q.taps + nC, q.db, q.dp, q.Dc, q.Df, q.Db,
                               q.Dp, del, fold_div ? 0 : maxdiff, q.div));
      AE_TRY(launch_kernel_spectrum_emb(ctx, dM, dD, Nk, Nl, Nx, Ny, col0, ncols, dc_w, Cemb));
      AE_TRY(launch_kernel_spectrum_emb(ctx, dD, dM, Nk, Nl, Nx, Ny, col0, ncols, df_w, Femb));
      AE_TRY(launch_tc_forward(ctx, S, (int)B, dD, dM, Xb, Cemb, 1.f / (float)dM, bias_b, norm, nullptr, Hb, nullptr, 0.0, 0, 0, 0));
      // re-forward O = conv(H; F, p), kept as E = O - Xt, and its mse (:1460-1463)
      // (after the last iteration nothing reads E any more: only its mse is formed)
      AE_TRY(launch_tc_forward(ctx, S, (int)B, dM, dD, Hb, Femb, 1.f / (float)dD, bias_p, norm, Xtb, n + 1 < n_iter ? Eb : nullptr,
                               q.mse + n + 1, mse_scale, ncols, col0, Ny));
      if (sharded && !use_comm) AE_TRY(reduce_over_devices(q.mse + n + 1, 1));
    }
    }  // !gram_tc
    if (!sharded && !inp.resident) {  // the bins-fastest spectra of the trained kernels, for the export below
      AE_TRY(kernel_spectrum_dev(ctx, (int64_t)dM * dD, Nk, Nl, Nx, Ny, dc_w, q.img, q.C, col0, ncols));
      AE_TRY(kernel_spectrum_dev(ctx, (int64_t)dM * dD, Nk, Nl, Nx, Ny, df_w, q.img, q.F, col0, ncols));
    }
  } else if (gram_ff) {
    float2 *Gx, *M0;
    float* dcs;
    AE_TRY(ctx->getT("gr_Gx", (size_t)S * dD * dD, &Gx));
    AE_TRY(ctx->getT("gr_M0", (size_t)S * dD * dD, &M0));
    AE_TRY(ctx->getT("gr_dc", (size_t)4 * dD, &dcs));
    // the one pass over the frames: Gx, M0 and "mse fft:" (:1440)
    AE_TRY(launch_gram_stats_ff(ctx, S, (int)B, dD, q.X, q.O, Gx, M0, q.mse, mse_scale, own_dc ? dcs : nullptr, ncols, col0, Ny, Nx,
                                inp.o_sNx, inp.o_sNy));
    if (sharded && !use_comm) AE_TRY(reduce_over_devices(q.mse, 1));
    const float gb = (float)((double)norm / (Norm * (double)B));
    for (int n = 0; n <= n_iter; n++) {
      const bool last = n == n_iter;
      AE_TRY(launch_gram_iter_ff(ctx, S, (int)B, dD, dM, Gx, M0, q.C, q.F, n == 0, gscale, gb, norm, own_dc ? dcs : nullptr, bias_b,
                                 bias_p, last ? nullptr : q.dCF, last ? nullptr : q.dCF + nKS, q.db, q.dp, n ? q.mse + n : nullptr,
                                 mse_scale, ncols, col0, Ny));
      if (n && sharded && !use_comm) AE_TRY(reduce_over_devices(q.mse + n, 1));
      if (last) break;
      if (!own_dc) AE_CUDA(cudaMemsetAsync(q.db, 0, (size_t)(dM + dD) * sizeof(float), st));
      AE_TRY(spectrum_taps_dev(ctx, 2 * (int64_t)dM * dD, Nk, Nl, Nx, Ny, q.dCF, q.work, q.img, q.taps, 1.f, col0, ncols));
      const bool fold_div = sharded && maxdiff;
      if (fold_div) {
        AE_TRY(launch_gradient_diff(ctx, dM, dD, Nk, Nl, dc_w, df_w, db_w, dp_w, q.div, srank, world));
        AE_TRY(launch_axpby(ctx, q.taps, q.div, 1.f, -10.f, (long long)(2 * nC + dM + dD)));
      }
      AE_TRY(reduce_over_devices(q.taps, (int64_t)(2 * nC + dM + dD)));
      AE_TRY(launch_fft_update(ctx, dM, dD, Nk, Nl, dc_w, df_w, db_w, dp_w, q.taps, q.taps + nC, q.db, q.dp, q.Dc, q.Df, q.Db,
                               q.Dp, del, fold_div ? 0 : maxdiff, q.div));
      AE_TRY(kernel_spectrum_dev(ctx, (int64_t)dM * dD, Nk, Nl, Nx, Ny, dc_w, q.img, q.C, col0, ncols));
      AE_TRY(kernel_spectrum_dev(ctx, (int64_t)dM * dD, Nk, Nl, Nx, Ny, df_w, q.img, q.F, col0, ncols));
    }
  } else {
  // H of the current kernels (the reference recomputes it inside gradient_k_io as H-hat, without the /dM: quirk F1)
  if (!use_small)
    AE_TRY(launch_spec_contract(ctx, B, dD, dM, S, q.X, nullptr, q.C, (int64_t)dD * S, S, 0, 1.f / (float)dM, bias_b, norm, q.H));
  for (int n = 0; n < n_iter; n++) {
    if (use_small) {
      // fused: H-hat, E, G formed per (bin, frame) in registers; the first iteration takes the caller's O, later ones
      // recompute O = conv(conv(X; C, b); F, p) from the updated kernels (what :1460-1461 left in freq_out)
      if (!own_dc) AE_CUDA(cudaMemsetAsync(q.db, 0, (size_t)(dM + dD) * sizeof(float), st));
      AE_TRY(launch_small_grad(ctx, B, dD, dM, S, q.X, Xt, n == 0 ? q.O : nullptr, q.C, q.F, bias_b, bias_p, norm, gscale,
                               (float)((double)norm / (Norm * (double)B)), q.dCF, q.dCF + nKS, q.db, q.dp));
    } else {
    // G[m] = sum_d1 (O - Xt)[d1] conj(F[d1][m])
    AE_TRY(launch_spec_contract(ctx, B, dD, dM, S, q.O, Xt, q.F, S, (int64_t)dM * S, 1, 1.f, nullptr, 0.f, q.G));
    // dC[m][d] = G[m] conj(X[d]) / Norm ; dF[d][m] = E[d] conj(H-hat[m]) / Norm, averaged over frames
    AE_TRY(launch_spec_outer(ctx, B, dM, dD, S, q.G, nullptr, q.X, 1.f, nullptr, 0.f, gscale, q.dCF));
    AE_TRY(launch_spec_outer(ctx, B, dD, dM, S, q.O, Xt, q.H, (float)dM, bias_b, -(float)(dM - 1) * norm, gscale, q.dCF + nKS));
    if (own_dc) AE_TRY(launch_spec_dc_sums(ctx, B, dM, dD, S, q.G, q.O, Xt, q.db, q.dp, (float)((double)norm / (Norm * (double)B))));
    else AE_CUDA(cudaMemsetAsync(q.db, 0, (size_t)(dM + dD) * sizeof(float), st));
    }
    // kernel-space gradients: C2R (unnormalised) + shrink_k (:1219-1226)
    AE_TRY(spectrum_taps_dev(ctx, 2 * (int64_t)dM * dD, Nk, Nl, Nx, Ny, q.dCF, q.work, q.img, q.taps, 1.f, col0, ncols));
    // bin-sharded devices also split the kernel-space multiobjective term: each folds its share into its partial block,
    // g = 1*g_mse - 10*g_div (:1252), so that the sum over devices is the whole combined gradient
    const bool fold_div = sharded && maxdiff;
    if (fold_div) {
      AE_TRY(launch_gradient_diff(ctx, dM, dD, Nk, Nl, dc_w, df_w, db_w, dp_w, q.div, srank, world));
      AE_TRY(launch_axpby(ctx, q.taps, q.div, 1.f, -10.f, (long long)(2 * nC + dM + dD)));
    }
    // data-parallel ranks average (bin-sharded devices: add) the raw gradient block here, before the non-linear clip
    AE_TRY(reduce_over_devices(q.taps, (int64_t)(2 * nC + dM + dD)));
    // clipped-momentum update in kernel space (+ multiobjective term)
    AE_TRY(launch_fft_update(ctx, dM, dD, Nk, Nl, dc_w, df_w, db_w, dp_w, q.taps, q.taps + nC, q.db, q.dp, q.Dc, q.Df, q.Db,
                             q.Dp, del, fold_div ? 0 : maxdiff, q.div));
    // new kernel spectra: pad_k + R2C (:1274-1282); c and f are adjacent in wts -> one batched transform
    AE_TRY(kernel_spectrum_dev(ctx, (int64_t)dM * dD, Nk, Nl, Nx, Ny, dc_w, q.img, q.C, col0, ncols));
    AE_TRY(kernel_spectrum_dev(ctx, (int64_t)dM * dD, Nk, Nl, Nx, Ny, df_w, q.img, q.F, col0, ncols));
    // re-forward (:1460-1461) and mse (:1463)
    if (use_small) {
      AE_TRY(launch_small_mse(ctx, B, dD, dM, S, q.X, Xt, q.C, q.F, bias_b, bias_p, norm, q.mse + n + 1, mse_scale, ncols, col0, Ny));
    } else {
    AE_TRY(launch_spec_contract(ctx, B, dD, dM, S, q.X, nullptr, q.C, (int64_t)dD * S, S, 0, 1.f / (float)dM, bias_b, norm, q.H));
    AE_TRY(launch_spec_contract(ctx, B, dM, dD, S, q.H, nullptr, q.F, (int64_t)dM * S, S, 0, 1.f / (float)dD, bias_p, norm, q.O));
    AE_TRY(launch_spec_mse(ctx, B, dD, dM, Nx, Ny, Xt, q.O, q.mse + n + 1, col0, ncols));
    }
    if (sharded && !use_comm) AE_TRY(reduce_over_devices(q.mse + n + 1, 1));
  }
  }  // !use_tc
  if (use_comm) AE_TRY(comm_allreduce(ctx, q.mse, (int64_t)n_iter + 1, sharded ? 0 : 1));
  // store_cfreq (:1484-1485) and export_cfreq (:1487-1488: c,f re-derived from the spectra: C2R/(NxNy) + kernel_invpad)
  if (cfreq) AE_CUDA(cudaMemcpyAsync(cfreq, q.C, nKS * sizeof(float2), k_out, st));
  if (ffreq) AE_CUDA(cudaMemcpyAsync(ffreq, q.F, nKS * sizeof(float2), k_out, st));
  if (sharded || inp.resident) {
    // the kernels in tap space are the master copy (identical on every device / resident in the net); no round trip
    // through the spectra (which is the identity up to fp32 rounding)
    AE_CUDA(cudaMemcpyAsync(q.taps, wts, 2 * nC * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    AE_TRY(spectrum_taps_dev(ctx, (int64_t)dM * dD, Nk, Nl, Nx, Ny, q.C, q.work, q.img, q.taps, 1.f / norm));
    AE_TRY(spectrum_taps_dev(ctx, (int64_t)dM * dD, Nk, Nl, Nx, Ny, q.F, q.work, q.img + (size_t)dM * dD * P, q.taps + nC,
                             1.f / norm));
  }
  AE_CUDA(cudaMemcpyAsync(c, q.taps, nC * sizeof(float), k_out, st));
  AE_CUDA(cudaMemcpyAsync(f, q.taps + nC, nC * sizeof(float), k_out, st));
  AE_CUDA(cudaMemcpyAsync(b, db_w, dM * sizeof(float), k_out, st));
  AE_CUDA(cudaMemcpyAsync(p, dp_w, dD * sizeof(float), k_out, st));
  if (inp.trace_dev)
    AE_CUDA(cudaMemcpyAsync(inp.trace_dev, q.mse, ((size_t)n_iter + 1) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (mse_trace)
    AE_CUDA(cudaMemcpyAsync(mse_trace, q.mse, ((size_t)n_iter + 1) * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (mse_trace || !inp.resident) AE_CUDA(cudaStreamSynchronize(st));
  return AEFFT_OK;
}
}  // namespace aefft

extern "C" {

int aefft_backprop_fft(aefft_ctx* ctx, int loc, int64_t B, int dD, int dM, int Nx, int Ny, int Nk, int Nl,
                       const float* in, const float* expout, const float* out, float* cfreq, float* c, float* ffreq,
                       float* f, float* b, float* p, float del0, int maxdiff, int n_iter, float* mse_trace) {
  AE_ARG(in && expout && out);
  FftTrainInputs inp;
  inp.in = in; inp.expout = expout; inp.out = out;
  return backprop_fft_run(ctx, loc, B, dD, dM, Nx, Ny, Nk, Nl, inp, cfreq, c, ffreq, f, b, p, del0, maxdiff, n_iter, mse_trace);
}

}  // extern "C"

extern "C" {

// backprop_fft on layers that live inside a per-frame strided layer block (the layout aefft_autoenc_fft writes):
// frame b's `in` / `out` start at in + b*frame_stride / out + b*frame_stride (device pointers).  expout = in, as the
// reference's app calls it (autoencoder.cpp:194).  The frames are gathered once into contiguous scratch.
int aefft_backprop_fft_strided(aefft_ctx* ctx, int64_t B, int dD, int dM, int Nx, int Ny, int Nk, int Nl, const float* in,
                               const float* out, int64_t frame_stride, float* c, float* f, float* b, float* p, float del0,
                               int maxdiff, int n_iter, float* mse_trace) {
  AE_ARG(ctx && in && out && B > 0 && frame_stride >= (int64_t)dD * Nx * Ny);
  AE_CUDA(cudaSetDevice(ctx->device));
  // transform the frames where they lie when the row kernels can address the per-frame blocks, else gather
  if (Ny >= 8 && Ny <= 4096 && !getenv("AEFFT_FFT_V1"))
  {
    FftTrainInputs inp;
    inp.in = in; inp.expout = in; inp.out = out; inp.fstride = frame_stride;
    return backprop_fft_run(ctx, AEFFT_DEVICE, B, dD, dM, Nx, Ny, Nk, Nl, inp, nullptr, c, nullptr, f, b, p, del0, maxdiff, n_iter,
                            mse_trace);
  }
  const size_t w = (size_t)dD * Nx * Ny * sizeof(float);
  float *gin, *gout;
  AE_TRY(ctx->getT("bpfs_in", (size_t)B * dD * Nx * Ny, &gin));
  AE_TRY(ctx->getT("bpfs_out", (size_t)B * dD * Nx * Ny, &gout));
  AE_CUDA(cudaMemcpy2DAsync(gin, w, in, (size_t)frame_stride * sizeof(float), w, (size_t)B, cudaMemcpyDeviceToDevice, ctx->stream));
  AE_CUDA(cudaMemcpy2DAsync(gout, w, out, (size_t)frame_stride * sizeof(float), w, (size_t)B, cudaMemcpyDeviceToDevice, ctx->stream));
  return aefft_backprop_fft(ctx, AEFFT_DEVICE, B, dD, dM, Nx, Ny, Nk, Nl, gin, gin, gout, nullptr, c, nullptr, f, b, p, del0, maxdiff,
                            n_iter, mse_trace);
}

// Strided 2-D copy (rows of `width` bytes, `height` rows) ordered on the ctx stream; kind as in aefft_memcpy.
int aefft_memcpy2d(aefft_ctx* ctx, void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width, int64_t height,
                   int kind) {
  AE_ARG(ctx && dst && src && width >= 0 && height >= 0 && dpitch >= width && spitch >= width && kind >= 0 && kind <= 2);
  AE_CUDA(cudaSetDevice(ctx->device));
  const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  AE_CUDA(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)height, k, ctx->stream));
  return AEFFT_OK;
}

}  // extern "C"
