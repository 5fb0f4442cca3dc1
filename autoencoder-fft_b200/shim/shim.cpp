// Implementation of aefft_shim.h: flattens the reference's nested vectors, calls the C ABI with host pointers
// (AEFFT_HOST: copies happen inside the call, like the reference's own per-call H2D/D2H) and writes results back.
#include "aefft_shim.h"

#include <cstdio>
#include <iostream>
#include <stdexcept>
#include <string>

#include "aefft.h"

namespace {

aefft_ctx* g_ctx = nullptr;
int g_device = 0;
int g_quirks = AEFFT_QUIRKS_ALL;

void check(int rc, const char* what) {
  if (rc != AEFFT_OK) throw std::runtime_error(std::string(what) + ": " + aefft_last_error());
}
aefft_ctx* ctx() {
  if (!g_ctx) check(aefft_create(&g_ctx, g_device), "aefft_create");
  return g_ctx;
}

std::vector<float> flat3(const aefft_v3& v) {
  std::vector<float> o;
  if (v.empty()) return o;
  o.reserve(v.size() * v[0].size() * v[0][0].size());
  for (const auto& a : v)
    for (const auto& r : a) o.insert(o.end(), r.begin(), r.end());
  return o;
}
void unflat3(const std::vector<float>& s, aefft_v3& v) {  // v is pre-sized by the caller (backproplib.cu:172-181)
  size_t n = 0;
  for (auto& a : v)
    for (auto& r : a)
      for (auto& x : r) x = s[n++];
}
std::vector<float> flat4(const aefft_v4& v) {
  std::vector<float> o;
  for (const auto& a : v)
    for (const auto& b : a)
      for (const auto& r : b) o.insert(o.end(), r.begin(), r.end());
  return o;
}
void unflat4(const std::vector<float>& s, aefft_v4& v) {
  size_t n = 0;
  for (auto& a : v)
    for (auto& b : a)
      for (auto& r : b)
        for (auto& x : r) x = s[n++];
}
void resize4(aefft_v4& v, int a, int b, int c, int d) { v.assign(a, aefft_v3(b, aefft_v2(c, aefft_v1(d, 0.f)))); }

struct Dims { int dM, dD, Nk, Nl, Nx, Ny; };
Dims dims_of(const aefft_v3& in, const aefft_v4& c) {
  return Dims{(int)c.size(), (int)c[0].size(), (int)c[0][0].size(), (int)c[0][0][0].size(), (int)in[0].size(),
              (int)in[0][0].size()};
}

void conv_any(aefft_v3& in, aefft_v3& out, aefft_v4& c, aefft_v1& b, int convention) {
  Dims d = dims_of(in, c);
  std::vector<float> fi = flat3(in), fc = flat4(c), fo((size_t)d.dM * d.Nx * d.Ny);
  check(aefft_conv_fwd(ctx(), AEFFT_HOST, convention, 1, d.dD, d.dM, d.Nx, d.Ny, d.Nk, d.Nl, fi.data(), fc.data(), b.data(),
                       fo.data()), "aefft_conv_fwd");
  unflat3(fo, out);
}

void backprop_any(int mode, aefft_v3& in, aefft_v3& out, aefft_v3& hin, aefft_v4& c, aefft_v1& b, aefft_v4& f, aefft_v1& p,
                  aefft_v4* dc, aefft_v1* db, aefft_v4* df, aefft_v1* dp, aefft_v4* ddc, aefft_v1* ddb, aefft_v4* ddf,
                  aefft_v1* ddp, float delmax, float alpha, int active) {
  Dims d = dims_of(in, c);
  std::vector<float> fi = flat3(in), fo = flat3(out), fh = flat3(hin), fc = flat4(c), ff = flat4(f);
  std::vector<float> fdc, fdf, fddc, fddf;
  if (dc) { fdc = flat4(*dc); fdf = flat4(*df); fddc = flat4(*ddc); fddf = flat4(*ddf); }
  float mse = 0.f;
  check(aefft_backprop_coord(ctx(), AEFFT_HOST, mode, g_quirks, 1, d.dD, d.dM, d.Nx, d.Ny, d.Nk, d.Nl, fi.data(), fo.data(),
                             fh.data(), fc.data(), b.data(), ff.data(), p.data(), dc ? fdc.data() : nullptr,
                             db ? db->data() : nullptr, df ? fdf.data() : nullptr, dp ? dp->data() : nullptr,
                             ddc ? fddc.data() : nullptr, ddb ? ddb->data() : nullptr, ddf ? fddf.data() : nullptr,
                             ddp ? ddp->data() : nullptr, delmax, alpha, active, &mse), "aefft_backprop_coord");
  std::cout << "mse: " << mse << std::endl;  // netlib.cpp:385, backproplib.cu:357,588
  unflat4(fc, c);
  unflat4(ff, f);
  if (dc) { unflat4(fdc, *dc); unflat4(fdf, *df); unflat4(fddc, *ddc); unflat4(fddf, *ddf); }
}

}  // namespace

void aefft_shim_set_quirks(int quirks) { g_quirks = quirks; }
void aefft_shim_set_device(int device) { g_device = device; }

void Pool(aefft_v3& in, aefft_v3& out, int scale) {
  const int D = (int)in.size(), Nx = (int)in[0].size(), Ny = (int)in[0][0].size();
  const int oNx = (int)out[0].size(), oNy = (int)out[0][0].size();
  std::vector<float> fi = flat3(in), fo((size_t)D * oNx * oNy);
  check(aefft_pool(ctx(), AEFFT_HOST, 1, D, Nx, Ny, oNx, oNy, scale, fi.data(), fo.data()), "aefft_pool");
  unflat3(fo, out);
}

void Init_conv(aefft_v4& c, aefft_v1& b, int mS, int dS, int kS, int lS, float max) {
  resize4(c, mS, dS, kS, lS);
  b.assign(mS, 0.f);
  std::vector<float> fc((size_t)mS * dS * kS * lS);
  check(aefft_init_conv(fc.data(), b.data(), mS, dS, kS, lS, max), "aefft_init_conv");
  unflat4(fc, c);
}

void SaveLoad_conv(aefft_v4& c, aefft_v1& b, int scale, int L, int io, int write) {
  const int dM = (int)c.size(), dD = (int)c[0].size(), Nk = (int)c[0][0].size(), Nl = (int)c[0][0][0].size();
  std::vector<float> fc = flat4(c);
  check(aefft_saveload_conv("./weights", fc.data(), b.data(), dM, dD, Nk, Nl, scale, L, io, write), "aefft_saveload_conv");
  if (write != 1) unflat4(fc, c);
}

void LoadParam(int& dM, int& Lk, int& Ll, int& scal, float& rmax) {
  check(aefft_load_param("New_Layer_Param.txt", &dM, &Lk, &Ll, &scal, &rmax), "aefft_load_param");
}

void Portion(aefft_v3& in, aefft_v3& hin, aefft_v3& out, aefft_v3& in_s, aefft_v3& hin_s, aefft_v3& out_s, int q) {
  aefft_v3* src[3] = {&in, &hin, &out};
  aefft_v3* dst[3] = {&in_s, &hin_s, &out_s};
  for (int t = 0; t < 3; t++) {
    const int D = (int)src[t]->size(), Nx = (int)(*src[t])[0].size(), Ny = (int)(*src[t])[0][0].size();
    std::vector<float> fi = flat3(*src[t]), fo((size_t)D * (Nx / q) * (Ny / q));
    check(aefft_portion(ctx(), AEFFT_HOST, 1, D, Nx, Ny, q, fi.data(), fo.data()), "aefft_portion");
    unflat3(fo, *dst[t]);
  }
}

void Conv(aefft_v3& in, aefft_v3& out, aefft_v4& c, aefft_v1& b) { conv_any(in, out, c, b, AEFFT_CONV_CPU); }
void Conv_gpu(aefft_v3& in, aefft_v3& out, aefft_v4& c, aefft_v1& b) { conv_any(in, out, c, b, AEFFT_CONV_CUDA); }

void backprop(aefft_v3& in, aefft_v3& out, aefft_v3& hin, aefft_v4& c, aefft_v1& b, aefft_v4& f, aefft_v1& p, float del) {
  backprop_any(AEFFT_MODE_CPU_REF, in, out, hin, c, b, f, p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
               nullptr, del, 0.f, 0);
}
void backprop_gpu(aefft_v3& in, aefft_v3& out, aefft_v3& hin, aefft_v4& c, aefft_v1& b, aefft_v4& f, aefft_v1& p,
                  aefft_v4& dc, aefft_v1& db, aefft_v4& df, aefft_v1& dp, aefft_v4& ddc, aefft_v1& ddb, aefft_v4& ddf,
                  aefft_v1& ddp, float delmax, float alpha, int active) {
  backprop_any(AEFFT_MODE_CUDA_REF, in, out, hin, c, b, f, p, &dc, &db, &df, &dp, &ddc, &ddb, &ddf, &ddp, delmax, alpha,
               active);
}
void backprop_gpu_cc(aefft_v3& in, aefft_v3& out, aefft_v3& hin, aefft_v4& c, aefft_v1& b, aefft_v4& f, aefft_v1& p,
                     aefft_v4& dc, aefft_v1& db, aefft_v4& df, aefft_v1& dp, aefft_v4& ddc, aefft_v1& ddb, aefft_v4& ddf,
                     aefft_v1& ddp, float delmax, float alpha, int active) {
  backprop_any(AEFFT_MODE_CUDA_REF_SYM, in, out, hin, c, b, f, p, &dc, &db, &df, &dp, &ddc, &ddb, &ddf, &ddp, delmax, alpha,
               active);
}

float act(float x) { return x; }    // identity: the leaky-ReLU bodies are commented out in the reference
float act1(float) { return 1.f; }

void kernel_pad(aefft_v4& c, aefft_v4& c_pad, int Nx, int Ny) {
  const int dM = (int)c.size(), dD = (int)c[0].size(), Nk = (int)c[0][0].size(), Nl = (int)c[0][0][0].size();
  std::vector<float> fc = flat4(c), fo((size_t)dM * dD * Nx * Ny);
  check(aefft_kernel_pad(ctx(), AEFFT_HOST, dM, dD, Nk, Nl, Nx, Ny, fc.data(), fo.data()), "aefft_kernel_pad");
  resize4(c_pad, dM, dD, Nx, Ny);
  unflat4(fo, c_pad);
}

void autoenc_fft(aefft_v4& layers, aefft_v5& net_c, aefft_v2& net_cfreq, aefft_v2& net_b, std::vector<int>& scale, int fft_l) {
  const int n_conv = (int)net_c.size(), n_layers = (int)layers.size();
  std::vector<int> dims, ldims;
  std::vector<int64_t> coff, boff, loff, cfoff;
  std::vector<float> c_all, b_all, layers_all;
  for (int n = 0; n < n_conv; n++) {
    coff.push_back((int64_t)c_all.size());
    boff.push_back((int64_t)b_all.size());
    std::vector<float> fc = flat4(net_c[n]);
    c_all.insert(c_all.end(), fc.begin(), fc.end());
    b_all.insert(b_all.end(), net_b[n].begin(), net_b[n].end());
    dims.push_back((int)net_c[n].size()); dims.push_back((int)net_c[n][0].size());
    dims.push_back((int)net_c[n][0][0].size()); dims.push_back((int)net_c[n][0][0][0].size());
  }
  for (int l = 0; l < n_layers; l++) {
    loff.push_back((int64_t)layers_all.size());
    std::vector<float> fl = flat3(layers[l]);
    layers_all.insert(layers_all.end(), fl.begin(), fl.end());
    ldims.push_back((int)layers[l].size()); ldims.push_back((int)layers[l][0].size()); ldims.push_back((int)layers[l][0][0].size());
  }
  // spectra of conv n live at the resolution of its input (encoder: after pooling; decoder: before unpooling)
  std::vector<float> cf_all;
  std::vector<size_t> cflen;
  for (int n = 0; n < n_conv; n++) {
    const int l = n < n_conv / 2 ? 2 * n + 1 : 2 * n;
    const size_t len = (size_t)dims[4 * n] * dims[4 * n + 1] * ldims[3 * l + 1] * (ldims[3 * l + 2] / 2 + 1) * 2;
    cfoff.push_back((int64_t)cf_all.size());
    cflen.push_back(len);
    cf_all.resize(cf_all.size() + len, 0.f);
  }
  // StoreLoad_cfreq (:1146-1161): the cache is used only when it holds every conv, otherwise it is rebuilt
  const int valid = (int)net_cfreq.size() >= n_conv ? 1 : 0;
  if (valid)
    for (int n = 0; n < n_conv; n++) std::copy(net_cfreq[n].begin(), net_cfreq[n].end(), cf_all.begin() + cfoff[n]);
  check(aefft_autoenc_fft(ctx(), AEFFT_HOST, 1, n_conv, dims.data(), c_all.data(), coff.data(), b_all.data(), boff.data(),
                          scale.data(), n_layers, ldims.data(), layers_all.data(), loff.data(), (int64_t)layers_all.size(),
                          valid, cf_all.data(), cfoff.data(), fft_l), "aefft_autoenc_fft");
  for (int l = 1; l < n_layers; l++) {
    if (!fft_l && l != n_layers - 1) continue;
    std::vector<float> fl(layers_all.begin() + loff[l], layers_all.begin() + (l + 1 < n_layers ? loff[l + 1] : (int64_t)layers_all.size()));
    unflat3(fl, layers[l]);
  }
  if (!valid) {
    net_cfreq.clear();
    for (int n = 0; n < n_conv; n++) net_cfreq.push_back(aefft_v1(cf_all.begin() + cfoff[n], cf_all.begin() + cfoff[n] + cflen[n]));
  }
}

void backprop_fft(aefft_v3& in, aefft_v3& expout, aefft_v3& out, aefft_v1& cfreq, aefft_v4& c, aefft_v1& ffreq, aefft_v4& f,
                  aefft_v1& b, aefft_v1& p, int dM, float del0, int maxdiff) {
  Dims d = dims_of(in, c);
  (void)dM;
  std::vector<float> fi = flat3(in), fe = flat3(expout), fo = flat3(out), fc = flat4(c), ff = flat4(f);
  const int n_iter = 100;  // hard-coded in the reference (:1446)
  std::vector<float> trace(n_iter + 1);
  check(aefft_backprop_fft(ctx(), AEFFT_HOST, 1, d.dD, d.dM, d.Nx, d.Ny, d.Nk, d.Nl, fi.data(), &in == &expout ? fi.data() : fe.data(),
                           fo.data(), cfreq.data(), fc.data(), ffreq.data(), ff.data(), b.data(), p.data(), del0, maxdiff, n_iter,
                           trace.data()), "aefft_backprop_fft");
  std::cout << "mse fft: " << trace[0] << std::endl;                                            // :1441
  for (int n = 0; n < n_iter; n++) std::cout << "n: " << n << " mse: " << trace[n + 1] << std::endl;  // :1464
  unflat4(fc, c);
  unflat4(ff, f);
}
