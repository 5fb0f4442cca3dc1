// Headless replay of the reference app's call sequence (autoencoder.cpp:98-120 init, :135-150 forward, :158-201
// training dispatch) written against the REFERENCE's header names -- it compiles unchanged against the reference's own
// netlib.h/backproplib.h/fft_backproplib.h or against autoencoder-fft_b200/shim/.  Dumps every tensor as raw float32
// so tests/test_shim_gpu.py can compare the shim (libaefft_shim.so -> C ABI -> CUDA) with the oracle.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
using namespace std;
#include "netlib.h"
#include "backproplib.h"
#include "fft_backproplib.h"

typedef vector<float> V1;
typedef vector<V1> V2;
typedef vector<V2> V3;
typedef vector<V3> V4;

static void dump3(FILE* f, const V3& v) { for (auto& a : v) for (auto& r : a) fwrite(r.data(), 4, r.size(), f); }
static void dump4(FILE* f, const V4& v) { for (auto& a : v) dump3(f, a); }
static void dump1(FILE* f, const V1& v) { fwrite(v.data(), 4, v.size(), f); }

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const string mode = argv[1];
  FILE* out = fopen(argv[2], "wb");
  int M = 4, Lk = 1, Ll = 1, s = 2;
  float rmax = 0.3f;
  LoadParam(M, Lk, Ll, s, rmax);  // New_Layer_Param.txt in the working directory
  const int Nk = 2 * (Lk + 1) + 1, Nl = 2 * (Ll + 1) + 1, D = 3, Nx = 32, Ny = 32;
  V3 in(D, V2(Nx, V1(Ny))), Pin(D, V2(Nx / s, V1(Ny / s))), hC(M, V2(Nx / s, V1(Ny / s))), PhC(D, V2(Nx / s, V1(Ny / s))),
      outl(D, V2(Nx, V1(Ny)));
  unsigned z = 12345u;
  for (int d = 0; d < D; d++) for (int i = 0; i < Nx; i++) for (int j = 0; j < Ny; j++) { z = z * 1664525u + 1013904223u; in[d][i][j] = (float)((z >> 16) & 255); }
  V4 c, f, dc, df, ddc, ddf;
  V1 b, p, db, dp, ddb, ddp;
  srand(1234);
  Init_conv(c, b, M, D, Nk, Nl, rmax);
  Init_conv(f, p, D, M, Nk, Nl, rmax);
  Init_conv(dc, db, M, D, Nk, Nl, 0); Init_conv(df, dp, D, M, Nk, Nl, 0);
  Init_conv(ddc, ddb, M, D, Nk, Nl, 0); Init_conv(ddf, ddp, D, M, Nk, Nl, 0);
  dump3(out, in); dump4(out, c); dump1(out, b); dump4(out, f); dump1(out, p);
  if (mode == "coord") {
    Pool(in, Pin, s);
    Conv_gpu(Pin, hC, c, b);
    Conv_gpu(hC, PhC, f, p);
    Pool(PhC, outl, -s);
    V3 in_s(D, V2(Nx / s, V1(Ny / s))), out_s = in_s, hC_s(M, V2(Nx / s, V1(Ny / s)));
    Portion(Pin, hC, PhC, in_s, hC_s, out_s, 1);
    backprop_gpu_cc(in_s, out_s, hC_s, c, b, f, p, dc, db, df, dp, ddc, ddb, ddf, ddp, 0.2f, 0.9f, 1);
    dump3(out, Pin); dump3(out, hC); dump3(out, PhC); dump3(out, outl);
    dump4(out, c); dump1(out, b); dump4(out, f); dump1(out, p);
    SaveLoad_conv(c, b, s, 0, 0, 1);
  } else {  // fft
    V4 layers; layers.push_back(in); layers.push_back(Pin); layers.push_back(hC); layers.push_back(PhC); layers.push_back(outl);
    vector<V4> net_c; net_c.push_back(c); net_c.push_back(f);
    V2 net_b; net_b.push_back(b); net_b.push_back(p);
    V2 net_cfreq;
    vector<int> scale; scale.push_back(s); scale.push_back(-s);
    autoenc_fft(layers, net_c, net_cfreq, net_b, scale, 1);
    for (auto& L : layers) dump3(out, L);
    backprop_fft(layers[1], layers[1], layers[3], net_cfreq[0], net_c[0], net_cfreq[1], net_c[1], net_b[0], net_b[1], M, 0.005f, 0);
    dump4(out, net_c[0]); dump1(out, net_b[0]); dump4(out, net_c[1]); dump1(out, net_b[1]);
  }
  fclose(out);
  return 0;
}
