// Device-resident replay of autoencoder.cpp's state model (:69-120 state, :135-150 forward, :158-201 training
// dispatch, :384-457 add/delete layer, :343-356 symmetric copy).  The reference keeps layers[], net_c[], net_b[],
// scale[] as nested host vectors and crosses the PCIe bus twice per conv; here they live in HBM for B frames at once.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "common.cuh"

using namespace aefft;

#include "net.cuh"

namespace {

int dev_alloc(float** p, size_t n) {
  AE_CUDA(cudaMalloc((void**)p, (n ? n : 1) * sizeof(float)));
  return AEFFT_OK;
}
int dev_zero(aefft_ctx* ctx, float* p, size_t n) {
  AE_CUDA(cudaMemsetAsync(p, 0, n * sizeof(float), ctx->stream));
  return AEFFT_OK;
}
void dev_free(float*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

int alloc_layer(aefft_net* net, LayerL& L) {
  AE_TRY(dev_alloc(&L.p, (size_t)net->B * L.D * L.Nx * L.Ny));
  return dev_zero(net->ctx, L.p, (size_t)net->B * L.D * L.Nx * L.Ny);
}

int alloc_pair_state(aefft_net* net, PairState& s, int dM, int dD, int Nk, int Nl) {
  const size_t nC = (size_t)dM * dD * Nk * Nl;
  float** w4[4] = {&s.dc, &s.df, &s.ddc, &s.ddf};
  for (auto p : w4) { AE_TRY(dev_alloc(p, nC)); AE_TRY(dev_zero(net->ctx, *p, nC)); }
  float** bm[2] = {&s.db, &s.ddb};
  for (auto p : bm) { AE_TRY(dev_alloc(p, dM)); AE_TRY(dev_zero(net->ctx, *p, dM)); }
  float** bd[2] = {&s.dp, &s.ddp};
  for (auto p : bd) { AE_TRY(dev_alloc(p, dD)); AE_TRY(dev_zero(net->ctx, *p, dD)); }
  return AEFFT_OK;
}
void free_pair_state(PairState& s) {
  dev_free(s.dc); dev_free(s.db); dev_free(s.df); dev_free(s.dp);
  dev_free(s.ddc); dev_free(s.ddb); dev_free(s.ddf); dev_free(s.ddp);
  s.gbuf = nullptr;  // a view, owned by the net
  s.gbuf_len = 0; s.gbuf_mode = -1;
}

// (re)build the fused gradient block for `mode`: every pair's gbuf becomes a view at its offset
int ensure_fused(aefft_net* net, int mode) {
  if (net->gall_mode == mode) return AEFFT_OK;
  int64_t total = 0;
  const int P = (int)net->pairs.size();
  for (int n = 0; n < P; n++) {
    const ConvL& e = net->convs[n];
    total += aefft::gbuf_len(mode, e.dD, e.dM, e.Nk, e.Nl);
  }
  if (total > net->gall_cap) {
    AE_CUDA(cudaStreamSynchronize(net->ctx->stream));
    dev_free(net->gall);
    AE_TRY(dev_alloc(&net->gall, (size_t)total));
    net->gall_cap = total;
  }
  int64_t off = 0;
  for (int n = 0; n < P; n++) {
    const ConvL& e = net->convs[n];
    PairState& s = net->pairs[n];
    s.gbuf = net->gall + off;
    s.gbuf_len = aefft::gbuf_len(mode, e.dD, e.dM, e.Nk, e.Nl);
    s.gbuf_mode = -1;  // nothing computed yet in this layout
    off += s.gbuf_len;
  }
  net->gall_len = total;
  net->gall_mode = mode;
  return AEFFT_OK;
}

int upload(aefft_ctx* ctx, float* dst, const float* src, size_t n) {
  AE_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  AE_CUDA(cudaStreamSynchronize(ctx->stream));
  return AEFFT_OK;
}

}  // namespace

extern "C" {

int aefft_net_create(aefft_ctx* ctx, aefft_net** out, int D, int Nx, int Ny, int64_t B) {
  AE_ARG(ctx && out && D > 0 && Nx > 0 && Ny > 0 && B > 0);
  AE_CUDA(cudaSetDevice(ctx->device));
  aefft_net* net = new aefft_net();
  net->ctx = ctx;
  net->B = B;
  LayerL in{D, Nx, Ny, nullptr};
  int r = alloc_layer(net, in);
  if (r == AEFFT_OK) r = dev_alloc(&net->mse_dev, 64);
  if (r != AEFFT_OK) { delete net; return r; }
  net->layers.push_back(in);
  *out = net;
  return AEFFT_OK;
}

// The reference zero-initialises its momentum / last-gradient vectors THROUGH Init_conv(..., 0) (autoencoder.cpp:103-107 at
// start-up, and in the 'n', 'z', 'x', 'd' handlers :288-292, :424-428, :447-451), which still draws one rand() per element.
// Burning the same number of draws keeps every later Init_conv on the reference's rand() stream for a given srand seed.
static void burn_zero_init_draws(int dM, int dD, int Nk, int Nl) {
  const long long nC = (long long)dM * dD * Nk * Nl;
  for (long long i = 0; i < 2 * (nC + dM) + 2 * (nC + dD); i++) (void)rand();
}

int aefft_net_destroy(aefft_net* net) {
  if (!net) return AEFFT_OK;
  cudaSetDevice(net->ctx->device);
  cudaStreamSynchronize(net->ctx->stream);
  for (auto& L : net->layers) dev_free(L.p);
  for (auto& c : net->convs) { dev_free(c.c); dev_free(c.b); }
  for (auto& s : net->pairs) free_pair_state(s);
  dev_free(net->mse_dev);
  dev_free(net->gall);
  net_fft_release(net);
  delete net;
  return AEFFT_OK;
}

// 'n' key (autoencoder.cpp:384-431).  The very first pair is what main() builds at start-up (:69-120) from the same
// five parameters.  New pair = innermost: input = current innermost hidden layer (or the frame), channels dD -> dM,
// resolution /scal.  Weights: Init_conv(c,b,dM,dD) then Init_conv(f,p,dD,dM) from libc rand() (:100-101, :412-413).
int aefft_net_add_layer(aefft_net* net, int dM, int Lk, int Ll, int scal, float rmax) {
  AE_ARG(net && dM > 0 && Lk >= 0 && Ll >= 0 && scal >= 1);
  aefft_ctx* ctx = net->ctx;
  AE_CUDA(cudaSetDevice(ctx->device));
  const int Nk = 2 * (Lk + 1) + 1, Nl = 2 * (Ll + 1) + 1;
  const int n = ((int)net->layers.size() - 1) / 2;  // centre layer (:392)
  const LayerL centre = net->layers[n];
  const int dD = centre.D, dNx = centre.Nx, dNy = centre.Ny;
  AE_ARG(dNx / scal > 0 && dNy / scal > 0);
  LayerL Pin{dD, dNx / scal, dNy / scal, nullptr}, hC{dM, dNx / scal, dNy / scal, nullptr},
      PhC{dD, dNx / scal, dNy / scal, nullptr}, outn{dD, dNx, dNy, nullptr};
  const size_t nC = (size_t)dM * dD * Nk * Nl;
  // draw the weights first (host only), then allocate EVERYTHING, and only then commit to the net's vectors: a failed
  // allocation frees what this call allocated and leaves layers / convs / pairs exactly as they were
  std::vector<float> c(nC), b(dM), f(nC), p(dD);
  AE_TRY(aefft_init_conv(c.data(), b.data(), dM, dD, Nk, Nl, rmax));
  AE_TRY(aefft_init_conv(f.data(), p.data(), dD, dM, Nk, Nl, rmax));
  burn_zero_init_draws(dM, dD, Nk, Nl);  // Init_conv(dc/df/ddc/ddf, ..., 0) (:103-107, :424-428)
  ConvL enc{dM, dD, Nk, Nl, scal}, dec{dD, dM, Nk, Nl, -scal};
  PairState st;
  auto build = [&]() -> int {
    AE_TRY(alloc_layer(net, Pin)); AE_TRY(alloc_layer(net, hC)); AE_TRY(alloc_layer(net, PhC)); AE_TRY(alloc_layer(net, outn));
    AE_TRY(dev_alloc(&enc.c, nC)); AE_TRY(dev_alloc(&enc.b, dM));
    AE_TRY(dev_alloc(&dec.c, nC)); AE_TRY(dev_alloc(&dec.b, dD));
    AE_TRY(upload(ctx, enc.c, c.data(), nC)); AE_TRY(upload(ctx, enc.b, b.data(), dM));
    AE_TRY(upload(ctx, dec.c, f.data(), nC)); AE_TRY(upload(ctx, dec.b, p.data(), dD));
    AE_TRY(alloc_pair_state(net, st, dM, dD, Nk, Nl));
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
    return AEFFT_OK;
  };
  const int rc = build();
  if (rc != AEFFT_OK) {
    cudaStreamSynchronize(ctx->stream);
    dev_free(Pin.p); dev_free(hC.p); dev_free(PhC.p); dev_free(outn.p);
    dev_free(enc.c); dev_free(enc.b); dev_free(dec.c); dev_free(dec.b);
    free_pair_state(st);
    return rc;
  }
  if (net->layers.size() == 1) {
    // first pair: layers = in, Pin, hC, PhC, out (:108-112)
    net->layers.push_back(Pin); net->layers.push_back(hC); net->layers.push_back(PhC); net->layers.push_back(outn);
  } else {
    net->layers.insert(net->layers.begin() + n + 1, {Pin, hC, PhC, outn});
  }
  const int mid = (int)net->convs.size() / 2;
  net->convs.insert(net->convs.begin() + mid, {enc, dec});
  net->pairs.push_back(st);  // innermost pair has the highest index
  net->gall_mode = -1;       // the fused gradient block is laid out again on the next gradient call
  net_fft_release(net);      // layer spectra are re-planned for the new topology (the reference clears net_cfreq, :429)
  return AEFFT_OK;
}

// 'd' key (:432-457): remove the innermost pair, never the last remaining one.
int aefft_net_delete_layer(aefft_net* net) {
  AE_ARG(net);
  if (net->convs.size() <= 2) { set_error("aefft_net_delete_layer: only one pair left"); return AEFFT_ERR_ARG; }
  AE_CUDA(cudaSetDevice(net->ctx->device));
  AE_CUDA(cudaStreamSynchronize(net->ctx->stream));
  int n = (int)net->convs.size() / 2;
  for (int i = n - 1; i <= n; i++) { dev_free(net->convs[i].c); dev_free(net->convs[i].b); }
  net->convs.erase(net->convs.begin() + n - 1, net->convs.begin() + n + 1);
  n = ((int)net->layers.size() - 1) / 2;
  for (int i = n - 1; i < n + 3; i++) dev_free(net->layers[i].p);
  net->layers.erase(net->layers.begin() + n - 1, net->layers.begin() + n + 3);
  free_pair_state(net->pairs.back());
  net->pairs.pop_back();
  net->gall_mode = -1;
  net_fft_release(net);  // (:454)
  return AEFFT_OK;
}

int aefft_net_num_pairs(const aefft_net* net) { return net ? (int)net->pairs.size() : -1; }

int aefft_net_conv_dims(const aefft_net* net, int n, int* dM, int* dD, int* Nk, int* Nl, int* scale) {
  AE_ARG(net && n >= 0 && n < (int)net->convs.size());
  const ConvL& c = net->convs[n];
  if (dM) *dM = c.dM;
  if (dD) *dD = c.dD;
  if (Nk) *Nk = c.Nk;
  if (Nl) *Nl = c.Nl;
  if (scale) *scale = c.scale;
  return AEFFT_OK;
}

int aefft_net_get_conv(aefft_net* net, int n, float* c, float* b) {
  AE_ARG(net && n >= 0 && n < (int)net->convs.size());
  const ConvL& L = net->convs[n];
  cudaStream_t s = net->ctx->stream;
  AE_CUDA(cudaSetDevice(net->ctx->device));
  if (c) AE_CUDA(cudaMemcpyAsync(c, L.c, (size_t)L.dM * L.dD * L.Nk * L.Nl * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (b) AE_CUDA(cudaMemcpyAsync(b, L.b, (size_t)L.dM * sizeof(float), cudaMemcpyDeviceToHost, s));
  AE_CUDA(cudaStreamSynchronize(s));
  return AEFFT_OK;
}

int aefft_net_set_conv(aefft_net* net, int n, const float* c, const float* b) {
  AE_ARG(net && n >= 0 && n < (int)net->convs.size());
  const ConvL& L = net->convs[n];
  AE_CUDA(cudaSetDevice(net->ctx->device));
  if (c) AE_TRY(upload(net->ctx, L.c, c, (size_t)L.dM * L.dD * L.Nk * L.Nl));
  if (b) AE_TRY(upload(net->ctx, L.b, b, (size_t)L.dM));
  return AEFFT_OK;
}

__global__ void sym_copy_kernel(const float* __restrict__ c, float* __restrict__ f, int dM, int dD, int T) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= dM * dD * T) return;
  int t = n % T, d = (n / T) % dD, m = n / (T * dD);
  f[(d * dM + m) * T + t] = c[n];
}

// 'p' key (:343-356): net_c[N-n_l][d][m][k][l] = net_c[n_l][m][d][k][l]
int aefft_net_set_symmetric(aefft_net* net, int n_l) {
  AE_ARG(net && n_l >= 0 && n_l < (int)net->pairs.size());
  const int N = (int)net->convs.size() - 1;
  const ConvL& e = net->convs[n_l];
  const ConvL& d = net->convs[N - n_l];
  AE_CUDA(cudaSetDevice(net->ctx->device));
  const int total = e.dM * e.dD * e.Nk * e.Nl;
  sym_copy_kernel<<<(total + 255) / 256, 256, 0, net->ctx->stream>>>(e.c, d.c, e.dM, e.dD, e.Nk * e.Nl);
  net->ctx->launches++;
  AE_CUDA(cudaGetLastError());
  return AEFFT_OK;
}

// 'z'/'x' keys (:281-292): the shared momentum / last-gradient buffers are re-initialised to zero.
int aefft_net_reset_momentum(aefft_net* net, int n_l) {
  AE_ARG(net && n_l >= 0 && n_l < (int)net->pairs.size());
  const ConvL& e = net->convs[n_l];
  PairState& s = net->pairs[n_l];
  const size_t nC = (size_t)e.dM * e.dD * e.Nk * e.Nl;
  AE_CUDA(cudaSetDevice(net->ctx->device));
  burn_zero_init_draws(e.dM, e.dD, e.Nk, e.Nl);  // the reference zeroes through Init_conv(..., 0) (:288-292, :447-451)
  AE_TRY(dev_zero(net->ctx, s.dc, nC)); AE_TRY(dev_zero(net->ctx, s.df, nC));
  AE_TRY(dev_zero(net->ctx, s.ddc, nC)); AE_TRY(dev_zero(net->ctx, s.ddf, nC));
  AE_TRY(dev_zero(net->ctx, s.db, e.dM)); AE_TRY(dev_zero(net->ctx, s.ddb, e.dM));
  AE_TRY(dev_zero(net->ctx, s.dp, e.dD)); AE_TRY(dev_zero(net->ctx, s.ddp, e.dD));
  return AEFFT_OK;
}

int aefft_net_layer(aefft_net* net, int l, int* D, int* Nx, int* Ny, float** dev_ptr) {
  AE_ARG(net && l >= 0 && l < (int)net->layers.size());
  const LayerL& L = net->layers[l];
  if (D) *D = L.D;
  if (Nx) *Nx = L.Nx;
  if (Ny) *Ny = L.Ny;
  if (dev_ptr) *dev_ptr = L.p;
  return AEFFT_OK;
}

int aefft_net_num_layers(const aefft_net* net) { return net ? (int)net->layers.size() : -1; }


// ImageToSpin_C: img[b][j][i][d] (bytes) -> spin[b][d][i][j] (float).  32 x 32 (i, j) tiles through shared memory so that
// both the byte reads (contiguous in i,d) and the float writes (contiguous in j) are coalesced.
}  // extern "C"

__global__ void image_to_spin_kernel(const unsigned char* __restrict__ img, float* __restrict__ spin, int D, int Nx, int Ny) {
  __shared__ unsigned char tile[32][32 * 4 + 4];  // [j][i*D + d], D <= 4
  const long long b = blockIdx.z;
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const unsigned char* src = img + b * (long long)Ny * Nx * D;
  const int wbytes = 32 * D;
  for (int idx = threadIdx.x; idx < 32 * wbytes; idx += blockDim.x) {
    const int j = idx / wbytes, x = idx - j * wbytes;
    const long long col = (long long)i0 * D + x;
    tile[j][x] = (j0 + j < Ny && col < (long long)Nx * D) ? src[(long long)(j0 + j) * Nx * D + col] : 0;
  }
  __syncthreads();
  float* dst = spin + b * (long long)D * Nx * Ny;
  for (int idx = threadIdx.x; idx < D * 32 * 32; idx += blockDim.x) {
    const int j = idx & 31, i = (idx >> 5) & 31, d = idx >> 10;
    if (i0 + i < Nx && j0 + j < Ny) dst[((long long)d * Nx + i0 + i) * Ny + j0 + j] = (float)tile[j][i * D + d];
  }
}

// SpinToImage_C (netlib.cpp:52-76): clamp(round(v), 0, 255) per channel, interleaved [rows = Ny][cols = Nx][D];
// mode 1 = SpinToImage_V (:79-92): (int)v truncated, stored as uchar (wraps modulo 256)
__global__ void spin_to_image_kernel(const float* __restrict__ spin, unsigned char* __restrict__ img, int D, int Nx, int Ny,
                                     int mode) {
  __shared__ unsigned char tile[32][32 * 4 + 4];  // [j][i*D + d], D <= 4
  const long long b = blockIdx.z;
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const float* src = spin + b * (long long)D * Nx * Ny;
  for (int idx = threadIdx.x; idx < D * 32 * 32; idx += blockDim.x) {
    const int j = idx & 31, i = (idx >> 5) & 31, d = idx >> 10;
    unsigned char o = 0;
    if (i0 + i < Nx && j0 + j < Ny) {
      const float v = src[((long long)d * Nx + i0 + i) * Ny + j0 + j];
      if (mode == 0) {
        int val = (int)roundf(v);
        val = val <= 255 ? val : 255;
        val = val >= 0 ? val : 0;
        o = (unsigned char)val;
      } else {
        o = (unsigned char)(int)v;
      }
    }
    tile[j][i * D + d] = o;
  }
  __syncthreads();
  unsigned char* dst = img + b * (long long)Ny * Nx * D;
  const int wbytes = 32 * D;
  for (int idx = threadIdx.x; idx < 32 * wbytes; idx += blockDim.x) {
    const int j = idx / wbytes, x = idx - j * wbytes;
    const long long col = (long long)i0 * D + x;
    if (j0 + j < Ny && col < (long long)Nx * D) dst[(long long)(j0 + j) * Nx * D + col] = tile[j][x];
  }
}

extern "C" {

int aefft_net_get_layer_u8(aefft_net* net, int layer, int mode, int loc, unsigned char* images) {
  AE_ARG(net && images && layer >= 0 && layer < (int)net->layers.size() && (mode == 0 || mode == 1));
  aefft_ctx* ctx = net->ctx;
  AE_CUDA(cudaSetDevice(ctx->device));
  const LayerL& L = net->layers[layer];
  AE_ARG(L.D <= 4 && net->B <= 65535);
  const size_t nbytes = (size_t)net->B * L.D * L.Nx * L.Ny;
  unsigned char* dev = images;
  if (loc == AEFFT_HOST) {
    void* stage;
    AE_TRY(ctx->get("net_layer_u8", nbytes, &stage));
    dev = (unsigned char*)stage;
  }
  dim3 grid((L.Nx + 31) / 32, (L.Ny + 31) / 32, (unsigned)net->B);
  spin_to_image_kernel<<<grid, 256, 0, ctx->stream>>>(L.p, dev, L.D, L.Nx, L.Ny, mode);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  if (loc == AEFFT_HOST) {
    AE_CUDA(cudaMemcpyAsync(images, dev, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return AEFFT_OK;
}

int aefft_net_set_frames_u8(aefft_net* net, int loc, const unsigned char* images) {
  AE_ARG(net && images);
  aefft_ctx* ctx = net->ctx;
  AE_CUDA(cudaSetDevice(ctx->device));
  LayerL& L0 = net->layers[0];
  AE_ARG(L0.D <= 4 && net->B <= 65535);
  const size_t nbytes = (size_t)net->B * L0.D * L0.Nx * L0.Ny;
  const unsigned char* dev = images;
  if (loc == AEFFT_HOST) {
    void* stage;
    AE_TRY(ctx->get("net_frames_u8", nbytes, &stage));
    AE_CUDA(cudaMemcpyAsync(stage, images, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    dev = (const unsigned char*)stage;
  }
  dim3 grid((L0.Nx + 31) / 32, (L0.Ny + 31) / 32, (unsigned)net->B);
  image_to_spin_kernel<<<grid, 256, 0, ctx->stream>>>(dev, L0.p, L0.D, L0.Nx, L0.Ny);
  ctx->launches++;
  AE_CUDA(cudaGetLastError());
  if (loc == AEFFT_HOST) AE_CUDA(cudaStreamSynchronize(ctx->stream));  // the caller may reuse `images` on return
  return AEFFT_OK;
}

// forward, coordinate space (autoencoder.cpp:135-150)
int aefft_net_forward(aefft_net* net, int loc, const float* frames) {
  AE_ARG(net && net->convs.size() >= 2);
  aefft_ctx* ctx = net->ctx;
  AE_CUDA(cudaSetDevice(ctx->device));
  LayerL& L0 = net->layers[0];
  const size_t n0 = (size_t)net->B * L0.D * L0.Nx * L0.Ny;
  if (frames && frames != L0.p) {
    AE_CUDA(cudaMemcpyAsync(L0.p, frames, n0 * sizeof(float),
                            loc == AEFFT_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
    // host frames: return only once they have been consumed (a pinned buffer would otherwise still be in flight and the
    // caller's refill for the next step would corrupt this one); device frames stay asynchronous on the ctx stream
    if (loc == AEFFT_HOST) AE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  const int N = (int)net->convs.size();
  for (int n = 0; n < N; n++) {
    const ConvL& c = net->convs[n];
    const int nl = 2 * n;
    LayerL &a = net->layers[nl], &m = net->layers[nl + 1], &z = net->layers[nl + 2];
    if (n < N / 2) {
      AE_TRY(launch_pool(ctx, net->B, a.D, a.Nx, a.Ny, m.Nx, m.Ny, c.scale, a.p, m.p));
      AE_TRY(launch_conv(ctx, fwd_window(c.Nk, c.Nl, AEFFT_CONV_CUDA), net->B, c.dD, c.dM, m.Nx, m.Ny, m.p, nullptr,
                         (float)c.dM, c.c, (int64_t)c.dD * c.Nk * c.Nl, (int64_t)c.Nk * c.Nl, c.b, z.p));
    } else {
      AE_TRY(launch_conv(ctx, fwd_window(c.Nk, c.Nl, AEFFT_CONV_CUDA), net->B, c.dD, c.dM, a.Nx, a.Ny, a.p, nullptr,
                         (float)c.dM, c.c, (int64_t)c.dD * c.Nk * c.Nl, (int64_t)c.Nk * c.Nl, c.b, m.p));
      AE_TRY(launch_pool(ctx, net->B, m.D, m.Nx, m.Ny, z.Nx, z.Ny, c.scale, m.p, z.p));
    }
  }
  return AEFFT_OK;
}

static int pair_views(aefft_net* net, int n_l, const ConvL** enc, const ConvL** dec, const LayerL** in,
                      const LayerL** hin, const LayerL** out) {
  AE_ARG(net && n_l >= 0 && n_l < (int)net->pairs.size());
  const int N = (int)net->convs.size();
  *enc = &net->convs[n_l];
  *dec = &net->convs[N - 1 - n_l];
  // in = layers[2n+1], hin = layers[2n+2], out = layers[size-2-2n] (autoencoder.cpp:161-169)
  *in = &net->layers[2 * n_l + 1];
  *hin = &net->layers[2 * n_l + 2];
  *out = &net->layers[net->layers.size() - 2 - 2 * n_l];
  return AEFFT_OK;
}

int aefft_net_pair_gradients(aefft_net* net, int n_l, int mode, int quirks, float** gbuf_dev, int64_t* gbuf_n) {
  const ConvL *enc, *dec;
  const LayerL *in, *hin, *out;
  AE_TRY(pair_views(net, n_l, &enc, &dec, &in, &hin, &out));
  aefft_ctx* ctx = net->ctx;
  AE_CUDA(cudaSetDevice(ctx->device));
  AE_TRY(ensure_fused(net, mode));
  PairState& s = net->pairs[n_l];
  const int64_t len = s.gbuf_len;
  s.gbuf_mode = mode;
  AE_TRY(coord_gradients_dev(ctx, mode, quirks, net->B, enc->dD, enc->dM, in->Nx, in->Ny, enc->Nk, enc->Nl, in->p, out->p,
                             hin->p, dec->c, s.gbuf));
  if (gbuf_dev) *gbuf_dev = s.gbuf;
  if (gbuf_n) *gbuf_n = len;
  return AEFFT_OK;
}

int aefft_net_pair_update(aefft_net* net, int n_l, int mode, int64_t B_global, float delmax, float alpha, float* mse) {
  const ConvL *enc, *dec;
  const LayerL *in, *hin, *out;
  AE_TRY(pair_views(net, n_l, &enc, &dec, &in, &hin, &out));
  aefft_ctx* ctx = net->ctx;
  AE_CUDA(cudaSetDevice(ctx->device));
  PairState& s = net->pairs[n_l];
  AE_ARG(s.gbuf && s.gbuf_mode == mode);
  float* mse_dev = net->mse_dev + (n_l % 64);
  AE_TRY(coord_update_dev(ctx, mode, B_global, enc->dD, enc->dM, in->Nx, in->Ny, enc->Nk, enc->Nl, s.gbuf, enc->c, enc->b,
                          dec->c, dec->b, s.dc, s.db, s.df, s.dp, s.ddc, s.ddb, s.ddf, s.ddp, delmax, alpha, mse_dev));
  if (mse) {
    AE_CUDA(cudaMemcpyAsync(mse, mse_dev, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return AEFFT_OK;
}

int aefft_net_train_pair(aefft_net* net, int n_l, int mode, int quirks, float delmax, float alpha, float* mse) {
  float* g = nullptr;
  int64_t len = 0;
  AE_TRY(aefft_net_pair_gradients(net, n_l, mode, quirks, &g, &len));
  AE_TRY(comm_allreduce(net->ctx, g, len, 0));  // data-parallel ranks: sum of the raw blocks (no-op for world == 1)
  return aefft_net_pair_update(net, n_l, mode, net->B * net->ctx->comm_world, delmax, alpha, mse);
}

// One whole training step: forward + train every pair once on that forward's activations (order 0..pairs-1).
// The reference trains one pair per frame (greedy, N4); a sweep over all pairs on the same forward is this repo's
// definition of a multi-pair "step" (DESIGN.md).  Pairs are independent given the activations.
int aefft_net_step(aefft_net* net, int loc, const float* frames, int mode, int quirks, float delmax, float alpha,
                   float* mse) {
  AE_ARG(net);
  aefft_ctx* ctx = net->ctx;
  AE_TRY(aefft_net_forward(net, loc, frames));
  const int P = (int)net->pairs.size();
  if (ctx->comm_world > 1) {
    // data-parallel frames (aefft_comm_init): raw gradient blocks of every pair, ONE all-reduce(sum) of the fused block
    // on the ctx stream, then the identical clipped-momentum update everywhere.  The sum precedes the non-linear clip.
    for (int n = 0; n < P; n++) AE_TRY(aefft_net_pair_gradients(net, n, mode, quirks, nullptr, nullptr));
    AE_TRY(comm_allreduce(ctx, net->gall, net->gall_len, 0));
    for (int n = 0; n < P; n++) AE_TRY(aefft_net_pair_update(net, n, mode, net->B * ctx->comm_world, delmax, alpha, nullptr));
  } else {
    for (int n = 0; n < P; n++) {
      AE_TRY(aefft_net_pair_gradients(net, n, mode, quirks, nullptr, nullptr));
      AE_TRY(aefft_net_pair_update(net, n, mode, net->B, delmax, alpha, nullptr));
    }
  }
  if (mse) {
    AE_CUDA(cudaMemcpyAsync(mse, net->mse_dev, sizeof(float) * (P < 64 ? P : 64), cudaMemcpyDeviceToHost, ctx->stream));
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return AEFFT_OK;
}

// Momentum sidecar: the reference's 's' / 'l' keys save the kernels and biases only -- the inertia state
// dc/db/df/dp and the last gradients ddc/ddb/ddf/ddp (autoencoder.cpp:75-83) are lost, so a reloaded net restarts its
// momentum from zero.  The sidecar keeps them next to the weight files:
//   <dir>/C_momentum_{L}_D={dD}_M={dM}_Lk={Lk}_Ll={Ll}.mom = raw little-endian float32
//   [dc dM*dD*T | db dM | df dD*dM*T | dp dD | ddc | ddb | ddf | ddp]   (T = Nk*Nl; same naming scheme as SaveLoad_conv)
int aefft_net_saveload_momentum(aefft_net* net, const char* dir, int n_l, int write) {
  AE_ARG(net && dir && n_l >= 0 && n_l < (int)net->pairs.size());
  aefft_ctx* ctx = net->ctx;
  AE_CUDA(cudaSetDevice(ctx->device));
  const ConvL& e = net->convs[n_l];
  PairState& st = net->pairs[n_l];
  const size_t nC = (size_t)e.dM * e.dD * e.Nk * e.Nl;
  float* parts[8] = {st.dc, st.db, st.df, st.dp, st.ddc, st.ddb, st.ddf, st.ddp};
  const size_t lens[8] = {nC, (size_t)e.dM, nC, (size_t)e.dD, nC, (size_t)e.dM, nC, (size_t)e.dD};
  size_t total = 0;
  for (size_t l : lens) total += l;
  std::vector<float> host(total + 1);
  const std::string path = std::string(dir) + "/C_momentum_" + std::to_string(n_l) + "_D=" + std::to_string(e.dD) + "_M=" +
                           std::to_string(e.dM) + "_Lk=" + std::to_string((e.Nk - 1) / 2 - 1) + "_Ll=" +
                           std::to_string((e.Nl - 1) / 2 - 1) + ".mom";
  if (write) {
    size_t off = 0;
    for (int i = 0; i < 8; i++) {
      AE_CUDA(cudaMemcpyAsync(host.data() + off, parts[i], lens[i] * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
      off += lens[i];
    }
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
    FILE* fh = fopen(path.c_str(), "wb");
    if (!fh) { set_error("cannot open %s for writing", path.c_str()); return AEFFT_ERR_IO; }
    const size_t n = fwrite(host.data(), sizeof(float), total, fh);
    fclose(fh);
    if (n != total) { set_error("short write to %s", path.c_str()); return AEFFT_ERR_IO; }
    return AEFFT_OK;
  }
  FILE* fh = fopen(path.c_str(), "rb");
  if (!fh) { set_error("cannot open %s", path.c_str()); return AEFFT_ERR_IO; }
  const size_t n = fread(host.data(), sizeof(float), total + 1, fh);  // one past: a longer file is a mismatch too
  fclose(fh);
  if (n != total) { set_error("%s: %zu floats, expected %zu", path.c_str(), n, total); return AEFFT_ERR_IO; }
  size_t off = 0;
  for (int i = 0; i < 8; i++) {
    AE_CUDA(cudaMemcpyAsync(parts[i], host.data() + off, lens[i] * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    off += lens[i];
  }
  AE_CUDA(cudaStreamSynchronize(ctx->stream));
  return AEFFT_OK;
}

// Layout of the fused gradient block for `mode`: offsets[n] = first float of pair n, *total = length (floats).
int aefft_net_fused_layout(aefft_net* net, int mode, int64_t* offsets, int64_t* total) {
  AE_ARG(net);
  AE_CUDA(cudaSetDevice(net->ctx->device));
  AE_TRY(ensure_fused(net, mode));
  if (offsets)
    for (size_t n = 0; n < net->pairs.size(); n++) offsets[n] = net->pairs[n].gbuf - net->gall;
  if (total) *total = net->gall_len;
  return AEFFT_OK;
}

}  // extern "C"
