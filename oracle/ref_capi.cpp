// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or shipped with the product path.
//
// Flat-pointer extern "C" doors onto the UNMODIFIED reference library functions, so that Python tests
// (ctypes) and bench.py's `--impl reference` / `cpu_baseline` leg can run the reference's own code.
// The reference sources are compiled where they lie (/root/reference/source/*.{cpp,cu}) by
// oracle/Makefile into oracle/_ref/libref.so; this file only converts between flat arrays and the
// nested std::vector types the reference API uses (netlib.h:4-24, backproplib.h:5-16,
// fft_backproplib.h:5-11).  Layout convention everywhere: feature maps [ch][Nx][Ny] (j fastest),
// kernels c[dM][dD][Nk][Nl], f[dD][dM][Nk][Nl]  (SURVEY App. A.1).
#include <opencv2/opencv.hpp>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "netlib.h"
#include "backproplib.h"
#include "fft_backproplib.h"

typedef std::vector<float> V1;
typedef std::vector<V1> V2;
typedef std::vector<V2> V3;
typedef std::vector<V3> V4;
typedef std::vector<V4> V5;

static V3 to3(const float* p, int a, int b, int c) {
  V3 v(a, V2(b, V1(c)));
  for (int i = 0; i < a; i++)
    for (int j = 0; j < b; j++)
      for (int k = 0; k < c; k++) v[i][j][k] = p[((size_t)i * b + j) * c + k];
  return v;
}
static void from3(const V3& v, float* p) {
  size_t n = 0;
  for (size_t i = 0; i < v.size(); i++)
    for (size_t j = 0; j < v[i].size(); j++)
      for (size_t k = 0; k < v[i][j].size(); k++) p[n++] = v[i][j][k];
}
static V4 to4(const float* p, int a, int b, int c, int d) {
  V4 v(a);
  for (int i = 0; i < a; i++) v[i] = to3(p + (size_t)i * b * c * d, b, c, d);
  return v;
}
static void from4(const V4& v, float* p) {
  size_t n = 0;
  for (size_t i = 0; i < v.size(); i++)
    for (size_t j = 0; j < v[i].size(); j++)
      for (size_t k = 0; k < v[i][j].size(); k++)
        for (size_t l = 0; l < v[i][j][k].size(); l++) p[n++] = v[i][j][k][l];
}
static V1 to1(const float* p, int a) { return V1(p, p + a); }
static void from1(const V1& v, float* p) {
  if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(float));
}

extern "C" {

void ref_srand(unsigned seed) { srand(seed); }

// netlib.cpp:167  Init_conv
void ref_init_conv(float* c, float* b, int mS, int dD, int kS, int lS, float max) {
  V4 cv;
  V1 bv;
  Init_conv(cv, bv, mS, dD, kS, lS, max);
  from4(cv, c);
  from1(bv, b);
}

// netlib.cpp:114  Pool.  (oNx,oNy) = caller-sized output, as in autoencoder.cpp:70-74,398-405.
void ref_pool(const float* in, float* out, int D, int Nx, int Ny, int oNx, int oNy, int scale) {
  V3 iv = to3(in, D, Nx, Ny), ov = to3(out, D, oNx, oNy);
  Pool(iv, ov, scale);
  from3(ov, out);
}

// netlib.cpp:318  Conv (CPU)
void ref_conv_cpu(const float* in, float* out, const float* c, const float* b, int dD, int dM, int Nx,
                  int Ny, int Nk, int Nl) {
  V3 iv = to3(in, dD, Nx, Ny), ov(dM, V2(Nx, V1(Ny)));
  V4 cv = to4(c, dM, dD, Nk, Nl);
  V1 bv = to1(b, dM);
  Conv(iv, ov, cv, bv);
  from3(ov, out);
}

// netlib.cpp:361  backprop (CPU).  c,b,f,p updated in place.
void ref_backprop_cpu(const float* in, const float* out, const float* hin, float* c, float* b, float* f,
                      float* p, float del, int dD, int dM, int Nx, int Ny, int Nk, int Nl) {
  V3 iv = to3(in, dD, Nx, Ny), ov = to3(out, dD, Nx, Ny), hv = to3(hin, dM, Nx, Ny);
  V4 cv = to4(c, dM, dD, Nk, Nl), fv = to4(f, dD, dM, Nk, Nl);
  V1 bv = to1(b, dM), pv = to1(p, dD);
  backprop(iv, ov, hv, cv, bv, fv, pv, del);
  from4(cv, c);
  from4(fv, f);
  from1(bv, b);
  from1(pv, p);
}

// netlib.cpp:292  Portion
void ref_portion(const float* in, const float* hin, const float* out, float* in_s, float* hin_s,
                 float* out_s, int D, int M, int Nx, int Ny, int q) {
  V3 iv = to3(in, D, Nx, Ny), hv = to3(hin, M, Nx, Ny), ov = to3(out, D, Nx, Ny);
  V3 is(D, V2(Nx / q, V1(Ny / q))), hs(M, V2(Nx / q, V1(Ny / q))), os(D, V2(Nx / q, V1(Ny / q)));
  Portion(iv, hv, ov, is, hs, os, q);
  from3(is, in_s);
  from3(hs, hin_s);
  from3(os, out_s);
}

// netlib.cpp:220  SaveLoad_conv (cwd-relative ./weights/...)
void ref_saveload_conv(float* c, float* b, int dM, int dD, int Nk, int Nl, int scale, int L, int io,
                       int write) {
  V4 cv = to4(c, dM, dD, Nk, Nl);
  V1 bv = to1(b, dM);
  SaveLoad_conv(cv, bv, scale, L, io, write);
  from4(cv, c);
  from1(bv, b);
}

// netlib.cpp:274  LoadParam (reads ./New_Layer_Param.txt)
void ref_loadparam(int* dM, int* Lk, int* Ll, int* scal, float* rmax) {
  LoadParam(*dM, *Lk, *Ll, *scal, *rmax);
}

// backproplib.cu:114  Conv_gpu   (needs a GPU)
void ref_conv_gpu(const float* in, float* out, const float* c, const float* b, int dD, int dM, int Nx,
                  int Ny, int Nk, int Nl) {
  V3 iv = to3(in, dD, Nx, Ny), ov(dM, V2(Nx, V1(Ny)));
  V4 cv = to4(c, dM, dD, Nk, Nl);
  V1 bv = to1(b, dM);
  Conv_gpu(iv, ov, cv, bv);
  from3(ov, out);
}

// backproplib.cu:291 backprop_gpu (sym=0) / :521 backprop_gpu_cc (sym=1).  (needs a GPU)
// All of c,b,f,p, dc,db,df,dp, ddc,ddb,ddf,ddp are updated in place.
void ref_backprop_gpu(int sym, const float* in, const float* out, const float* hin, float* c, float* b,
                      float* f, float* p, float* dc, float* db, float* df, float* dp, float* ddc,
                      float* ddb, float* ddf, float* ddp, float delmax, float alpha, int active, int dD,
                      int dM, int Nx, int Ny, int Nk, int Nl) {
  V3 iv = to3(in, dD, Nx, Ny), ov = to3(out, dD, Nx, Ny), hv = to3(hin, dM, Nx, Ny);
  V4 cv = to4(c, dM, dD, Nk, Nl), fv = to4(f, dD, dM, Nk, Nl);
  V4 dcv = to4(dc, dM, dD, Nk, Nl), dfv = to4(df, dD, dM, Nk, Nl);
  V4 ddcv = to4(ddc, dM, dD, Nk, Nl), ddfv = to4(ddf, dD, dM, Nk, Nl);
  V1 bv = to1(b, dM), pv = to1(p, dD), dbv = to1(db, dM), dpv = to1(dp, dD), ddbv = to1(ddb, dM),
     ddpv = to1(ddp, dD);
  if (sym)
    backprop_gpu_cc(iv, ov, hv, cv, bv, fv, pv, dcv, dbv, dfv, dpv, ddcv, ddbv, ddfv, ddpv, delmax, alpha,
                    active);
  else
    backprop_gpu(iv, ov, hv, cv, bv, fv, pv, dcv, dbv, dfv, dpv, ddcv, ddbv, ddfv, ddpv, delmax, alpha,
                 active);
  from4(cv, c);
  from4(fv, f);
  from4(dcv, dc);
  from4(dfv, df);
  from4(ddcv, ddc);
  from4(ddfv, ddf);
  from1(bv, b);
  from1(pv, p);
  from1(dbv, db);
  from1(dpv, dp);
  from1(ddbv, ddb);
  from1(ddpv, ddp);
}

// fft_backproplib.cu:1018  kernel_pad
void ref_kernel_pad(const float* c, float* c_pad, int dM, int dD, int Nk, int Nl, int Nx, int Ny) {
  V4 cv = to4(c, dM, dD, Nk, Nl), pv;
  kernel_pad(cv, pv, Nx, Ny);
  from4(pv, c_pad);
}

// fft_backproplib.cu:1331  autoenc_fft   (needs a GPU)
//   n_conv convs; conv n has shape dims[4n..4n+3] = (dM,dD,Nk,Nl), weights at c_all+coff[n], bias at
//   b_all+boff[n]; layer l has shape ldims[3l..3l+2]=(D,Nx,Ny) at layers_all+loff[l] (2*n_conv+1 layers).
//   cfreq_state: 0 = net_cfreq empty on entry (reference computes+stores it), 1 = supplied in cfreq_all
//   at cfoff[n] (length cflen[n]).  On return cfreq_all holds the (possibly newly built) spectra.
void ref_autoenc_fft(int n_conv, const int* dims, const float* c_all, const long* coff,
                     const float* b_all, const long* boff, const int* scale, int n_layers,
                     const int* ldims, float* layers_all, const long* loff, int cfreq_state,
                     float* cfreq_all, const long* cfoff, const long* cflen, int fft_l) {
  V5 net_c(n_conv);
  V2 net_b(n_conv), net_cfreq;
  for (int n = 0; n < n_conv; n++) {
    net_c[n] = to4(c_all + coff[n], dims[4 * n], dims[4 * n + 1], dims[4 * n + 2], dims[4 * n + 3]);
    net_b[n] = to1(b_all + boff[n], dims[4 * n]);
  }
  if (cfreq_state)
    for (int n = 0; n < n_conv; n++) net_cfreq.push_back(to1(cfreq_all + cfoff[n], (int)cflen[n]));
  std::vector<int> sc(scale, scale + n_conv);
  V4 layers(n_layers);
  for (int l = 0; l < n_layers; l++)
    layers[l] = to3(layers_all + loff[l], ldims[3 * l], ldims[3 * l + 1], ldims[3 * l + 2]);
  autoenc_fft(layers, net_c, net_cfreq, net_b, sc, fft_l);
  for (int l = 0; l < n_layers; l++) from3(layers[l], layers_all + loff[l]);
  for (int n = 0; n < n_conv && n < (int)net_cfreq.size(); n++) from1(net_cfreq[n], cfreq_all + cfoff[n]);
}

// fft_backproplib.cu:1381  backprop_fft   (needs a GPU).  cfreq,ffreq,c,f,b,p updated in place.
void ref_backprop_fft(const float* in, const float* expout, const float* out, float* cfreq, float* c,
                      float* ffreq, float* f, float* b, float* p, int dD, int dM, int Nx, int Ny, int Nk,
                      int Nl, float del0, int maxdiff) {
  V3 iv = to3(in, dD, Nx, Ny), ev = to3(expout, dD, Nx, Ny), ov = to3(out, dD, Nx, Ny);
  V4 cv = to4(c, dM, dD, Nk, Nl), fv = to4(f, dD, dM, Nk, Nl);
  V1 bv = to1(b, dM), pv = to1(p, dD);
  int nf = dM * dD * Nx * (Ny / 2 + 1) * 2;
  V1 cfv = to1(cfreq, nf), ffv = to1(ffreq, nf);
  backprop_fft(iv, ev, ov, cfv, cv, ffv, fv, bv, pv, dM, del0, maxdiff);
  from1(cfv, cfreq);
  from1(ffv, ffreq);
  from4(cv, c);
  from4(fv, f);
  from1(bv, b);
  from1(pv, p);
}

}  // extern "C"
