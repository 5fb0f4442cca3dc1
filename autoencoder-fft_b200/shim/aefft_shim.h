// C++ mirror of the reference's library boundary, implemented on top of the C ABI (include/aefft.h) only.
//
// The reference has no FFI layer: source/autoencoder.cpp calls free functions declared in source/netlib.h:4-24,
// source/backproplib.h:5-16 and source/fft_backproplib.h:5-11 (nested std::vector<float> by non-const reference,
// C++ linkage, no namespace).  This header declares the hot-path subset with the SAME names, parameter order and
// meaning, so that autoencoder.cpp links against libaefft_shim.so instead of netlib.o/backproplib.o/fft_backproplib.o
// (INTEGRATION.md).  The OpenCV image helpers (ImageToSpin_C, SpinToImage_*) are not part of the hot path and stay in
// the reference's netlib.cpp.
//
// Differences in error behaviour, on purpose: the reference returns void and checks nothing; these functions throw
// std::runtime_error (message from aefft_last_error()) when the GPU path fails -- there is no CPU fallback.
#ifndef AEFFT_SHIM_H
#define AEFFT_SHIM_H
#include <vector>

typedef std::vector<float> aefft_v1;
typedef std::vector<aefft_v1> aefft_v2;
typedef std::vector<aefft_v2> aefft_v3;  // feature maps [ch][Nx][Ny]
typedef std::vector<aefft_v3> aefft_v4;  // kernels [m][d][k][l] / layer lists
typedef std::vector<aefft_v4> aefft_v5;  // net_c

// ---- netlib.h
void Pool(aefft_v3& in, aefft_v3& out, int scale);                                     // netlib.cpp:114
void Init_conv(aefft_v4& c, aefft_v1& b, int mS, int dS, int kS, int lS, float max);   // netlib.cpp:167
void SaveLoad_conv(aefft_v4& c, aefft_v1& b, int scale, int L, int io, int write);     // netlib.cpp:220
void LoadParam(int& dM, int& Lk, int& Ll, int& scal, float& rmax);                     // netlib.cpp:274
void Portion(aefft_v3& in, aefft_v3& hin, aefft_v3& out, aefft_v3& in_s, aefft_v3& hin_s, aefft_v3& out_s, int q);  // :292
void Conv(aefft_v3& in, aefft_v3& out, aefft_v4& c, aefft_v1& b);                      // netlib.cpp:318
void backprop(aefft_v3& in, aefft_v3& out, aefft_v3& hin, aefft_v4& c, aefft_v1& b, aefft_v4& f, aefft_v1& p,
              float del);                                                              // netlib.cpp:361

// ---- backproplib.h
void Conv_gpu(aefft_v3& in, aefft_v3& out, aefft_v4& c, aefft_v1& b);                  // backproplib.cu:114
void backprop_gpu(aefft_v3& in, aefft_v3& out, aefft_v3& hin, aefft_v4& c, aefft_v1& b, aefft_v4& f, aefft_v1& p,
                  aefft_v4& dc, aefft_v1& db, aefft_v4& df, aefft_v1& dp, aefft_v4& ddc, aefft_v1& ddb, aefft_v4& ddf,
                  aefft_v1& ddp, float delmax, float alpha, int active);               // backproplib.cu:291
void backprop_gpu_cc(aefft_v3& in, aefft_v3& out, aefft_v3& hin, aefft_v4& c, aefft_v1& b, aefft_v4& f, aefft_v1& p,
                     aefft_v4& dc, aefft_v1& db, aefft_v4& df, aefft_v1& dp, aefft_v4& ddc, aefft_v1& ddb, aefft_v4& ddf,
                     aefft_v1& ddp, float delmax, float alpha, int active);            // backproplib.cu:521
float act(float x);                                                                    // backproplib.cu:38
float act1(float x);                                                                   // backproplib.cu:45

// ---- fft_backproplib.h
void autoenc_fft(aefft_v4& layers, aefft_v5& net_c, aefft_v2& net_cfreq, aefft_v2& net_b, std::vector<int>& scale,
                 int fft_l);                                                           // fft_backproplib.cu:1331
void kernel_pad(aefft_v4& c, aefft_v4& c_pad, int Nx, int Ny);                         // fft_backproplib.cu:1018
void backprop_fft(aefft_v3& in, aefft_v3& expout, aefft_v3& out, aefft_v1& cfreq, aefft_v4& c, aefft_v1& ffreq,
                  aefft_v4& f, aefft_v1& b, aefft_v1& p, int dM, float del0, int maxdiff);  // fft_backproplib.cu:1381

// ---- knobs of the shim (not in the reference)
// Bug-compat switches applied by backprop_gpu (AEFFT_QUIRK_*; default AEFFT_QUIRKS_ALL) and the GPU the shim uses.
void aefft_shim_set_quirks(int quirks);
void aefft_shim_set_device(int device);
#endif
