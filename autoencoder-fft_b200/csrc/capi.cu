// C ABI (include/aefft.h): context, coordinate-space entry points and the netlib glue.
// Every entry point names the reference function it replaces in aefft.h; this file is the host-side orchestration.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

#include "common.cuh"

namespace aefft {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace aefft

using namespace aefft;

int aefft_ctx::ensure_dyn_smem(const void* func, size_t bytes) {
  size_t& have = dyn_smem[func];
  if (bytes > have) {
    AE_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
  return AEFFT_OK;
}

int aefft_ctx::get(const char* name, size_t bytes, void** out) {
  Scratch& s = scratch[name];
  if (s.cap < bytes) {
    if (s.p) {
      // buffers may still be in use by queued work on the stream
      AE_CUDA(cudaStreamSynchronize(stream));
      AE_CUDA(cudaFree(s.p));
      s.p = nullptr;
      s.cap = 0;
    }
    size_t cap = bytes + bytes / 8 + 256;
    AE_CUDA(cudaMalloc(&s.p, cap));
    s.cap = cap;
  }
  *out = s.p;
  return AEFFT_OK;
}

cudaEvent_t aefft_ctx::get_event() {
  if (!event_pool.empty()) {
    cudaEvent_t e = event_pool.back();
    event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

int aefft_ctx::get_pinned(const char* name, size_t bytes, void** out) {
  Scratch& s = pinned[name];
  if (s.cap < bytes) {
    if (s.p) {
      AE_CUDA(cudaStreamSynchronize(stream));
      AE_CUDA(cudaFreeHost(s.p));
      s.p = nullptr;
      s.cap = 0;
    }
    size_t cap = bytes + bytes / 8 + 256;
    AE_CUDA(cudaMallocHost(&s.p, cap));
    s.cap = cap;
  }
  *out = s.p;
  return AEFFT_OK;
}

namespace aefft {

// Host<->device staging for loc == AEFFT_HOST (the reference's per-call H2D/D2H, backproplib.cu:150-151,171).
struct Stage {
  aefft_ctx* ctx;
  int loc;
  int n = 0;
  struct Out { void* host; void* dev; size_t bytes; } outs[24];
  int n_out = 0;
  Stage(aefft_ctx* c, int l) : ctx(c), loc(l) {}
  // read-only input
  int in(const char* name, const float* p, size_t count, const float** dev) {
    if (loc == AEFFT_DEVICE || p == nullptr) { *dev = p; return AEFFT_OK; }
    float* d;
    AE_TRY(ctx->getT(name, count, &d));
    AE_CUDA(cudaMemcpyAsync(d, p, count * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    *dev = d;
    return AEFFT_OK;
  }
  // read-write (upload now, download at finish) or write-only (upload=false)
  int inout(const char* name, float* p, size_t count, float** dev, bool upload = true) {
    if (loc == AEFFT_DEVICE || p == nullptr) { *dev = p; return AEFFT_OK; }
    float* d;
    AE_TRY(ctx->getT(name, count, &d));
    if (upload) AE_CUDA(cudaMemcpyAsync(d, p, count * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    AE_ARG(n_out < (int)(sizeof(outs) / sizeof(outs[0])));
    outs[n_out++] = {p, d, count * sizeof(float)};
    *dev = d;
    return AEFFT_OK;
  }
  int finish() {
    if (loc == AEFFT_DEVICE) return AEFFT_OK;
    for (int i = 0; i < n_out; i++)
      AE_CUDA(cudaMemcpyAsync(outs[i].host, outs[i].dev, outs[i].bytes, cudaMemcpyDeviceToHost, ctx->stream));
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
    return AEFFT_OK;
  }
};

__global__ void c1_bias_kernel(const float* __restrict__ f, const float* __restrict__ T, float* __restrict__ GB, int dD,
                               int dM, int NkNl) {
  // quirk C1 (backproplib.cu:220): gB[m] = sum_{k1,l1} f[dD-1][m][k1][l1] * T[dD-1][k1][l1]
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= dM) return;
  double s = 0.0;
  for (int t = 0; t < NkNl; t++) s += (double)f[((dD - 1) * dM + m) * NkNl + t] * (double)T[(dD - 1) * NkNl + t];
  GB[m] = (float)s;
}

int64_t gbuf_len(int mode, int dD, int dM, int Nk, int Nl) {
  const int64_t T = (int64_t)Nk * Nl, nC = (int64_t)dM * dD * T;
  if (mode == AEFFT_MODE_CPU_REF) {
    const int64_t S = dD * T;
    return S * S + S + nC + dD + 1;
  }
  return 2 * nC + dM + dD + 1;
}

// Raw (un-normalised, summed over the B local frames) gradient block of one layer pair into gbuf (device).
int coord_gradients_dev(aefft_ctx* ctx, int mode, int quirks, int64_t B, int dD, int dM, int Nx, int Ny, int Nk, int Nl,
                        const float* in, const float* out, const float* hin, const float* f, float* gbuf) {
  AE_ARG(mode == AEFFT_MODE_CPU_REF || mode == AEFFT_MODE_CUDA_REF || mode == AEFFT_MODE_CUDA_REF_SYM);
  AE_ARG(B > 0 && dD > 0 && dM > 0 && Nx > 0 && Ny > 0 && Nk > 0 && Nl > 0);
  const int64_t T = (int64_t)Nk * Nl, nC = (int64_t)dM * dD * T;
  if (mode == AEFFT_MODE_CPU_REF) {
    const int64_t S = dD * T;
    float* R = gbuf;
    float* BM = R + S * S;
    float* GF = BM + S;
    float* GP = GF + nC;
    float* SQ = GP + dD;
    Window wc = fwd_window(Nk, Nl, AEFFT_CONV_CPU);
    AOperand A;
    A.mode = A_SHIFT; A.a0 = out; A.a1 = in; A.nA = (int)S; A.src_ch = dD;
    A.ei0 = tap_base(Nk, AEFFT_CONV_CPU); A.ej0 = tap_base(Nl, AEFFT_CONV_CPU); A.eNk = Nk; A.eNl = Nl; A.out_lo = 1;
    AE_TRY(launch_wgrad(ctx, wc, B, Nx, Ny, A, in, dD, R, BM, nullptr));
    AOperand E;
    E.mode = A_DIFF; E.a0 = out; E.a1 = in; E.nA = dD; E.src_ch = dD;
    AE_TRY(launch_wgrad(ctx, wc, B, Nx, Ny, E, hin, dM, GF, GP, SQ));
    return AEFFT_OK;
  }
  float* GC = gbuf;
  float* GF = GC + nC;
  float* GB = GF + nC;
  float* GP = GB + dM;
  float* SQ = GP + dD;
  // hidden delta dh[m](u,v) = sum f[d1][m][k1][l1] e[d1](u+ik1, v+il1), e = out - in fused at tile load
  float* dh;
  AE_TRY(ctx->getT("coord_dh", (size_t)B * dM * Nx * Ny, &dh));
  AE_TRY(launch_conv(ctx, tr_window(Nk, Nl, AEFFT_CONV_CUDA), B, dD, dM, Nx, Ny, out, in, 0.f, f, T, (int64_t)dM * T,
                     nullptr, dh));
  Window wg = fwd_window(Nk, Nl, AEFFT_CONV_CUDA);
  bool done = false;
  if (ctx->precision != AEFFT_PRECISION_FP32) {
    // tensor-core path: GC and GF in one launch (adjacent in gbuf), bias sums / sum e^2 by a streaming reduction
    const int passes = ctx->precision == AEFFT_PRECISION_BF16X3 ? 3 : 1;
    // streaming TMEM-operand kernel: weight gradients, bias gradients and sum e^2 in one pass
    int rc = launch_wgrad_ts(ctx, wg, B, dD, dM, Nx, Ny, in, out, hin, dh, GC, GB, GP, SQ, passes);
    if (rc == AEFFT_OK) {
      done = true;
    } else if (rc != AEFFT_ERR_UNSUPPORTED) {
      return rc;
    }
    if (!done) rc = launch_wgrad_tc(ctx, wg, B, dD, dM, Nx, Ny, in, out, hin, dh, GC, passes);
    if (done) {
    } else if (rc == AEFFT_OK) {
      AE_TRY(launch_channel_sums(ctx, B, dM, Nx, Ny, dh, nullptr, GB, nullptr));
      AE_TRY(launch_channel_sums(ctx, B, dD, Nx, Ny, out, in, GP, SQ));
      done = true;
    } else if (rc != AEFFT_ERR_UNSUPPORTED) {
      return rc;
    }
  }
  if (!done) {
    AOperand A;
    A.mode = A_PLAIN; A.a0 = dh; A.nA = dM; A.src_ch = dM;
    AE_TRY(launch_wgrad(ctx, wg, B, Nx, Ny, A, in, dD, GC, GB, nullptr));
    AOperand E;
    E.mode = A_DIFF; E.a0 = out; E.a1 = in; E.nA = dD; E.src_ch = dD;
    AE_TRY(launch_wgrad(ctx, wg, B, Nx, Ny, E, hin, dM, GF, GP, SQ));
  }
  if (mode == AEFFT_MODE_CUDA_REF) {
    if (quirks & (AEFFT_QUIRK_C3 | AEFFT_QUIRK_C4)) {
      AE_ARG(Nx == Ny);  // the compiled reference is only defined on square frames (stride quirk C2)
      AE_TRY(launch_quirk_dF(ctx, quirks, B, dD, dM, Nx, Ny, Nk, Nl, out, in, hin, GF));
    }
    if (quirks & AEFFT_QUIRK_C1) {
      float* Tb;
      AE_TRY(ctx->getT("coord_border", (size_t)dD * T, &Tb));
      AE_TRY(launch_border_sums(ctx, B, dD, Nx, Ny, Nk, Nl, tap_base(Nk, AEFFT_CONV_CUDA),
                                tap_base(Nl, AEFFT_CONV_CUDA), 0, out, in, Tb));
      c1_bias_kernel<<<(dM + 127) / 128, 128, 0, ctx->stream>>>(f, Tb, GB, dD, dM, (int)T);
      ctx->launches++;
      AE_CUDA(cudaGetLastError());
    }
  }
  return AEFFT_OK;
}

int coord_update_dev(aefft_ctx* ctx, int mode, int64_t B_global, int dD, int dM, int Nx, int Ny, int Nk, int Nl,
                     const float* gbuf, float* c, float* b, float* f, float* p, float* dc, float* db, float* df,
                     float* dp, float* ddc, float* ddb, float* ddf, float* ddp, float delmax, float alpha,
                     float* mse_dev) {
  AE_ARG(B_global > 0);
  if (mode != AEFFT_MODE_CPU_REF) AE_ARG(dc && db && dp && (mode == AEFFT_MODE_CUDA_REF_SYM || df));
  UpdateArgs a;
  a.mode = mode; a.dD = dD; a.dM = dM; a.Nk = Nk; a.Nl = Nl;
  // Norm is a float product in the reference (backproplib.cu:303,533; netlib.cpp:373)
  float Norm = (float)((double)dD * dM * Nk * Nl * Nx * Ny);
  if (mode == AEFFT_MODE_CUDA_REF_SYM) Norm *= 2.f;
  a.inv_norm = (float)(1.0 / ((double)Norm * (double)B_global));
  a.delmax = delmax; a.alpha = alpha;
  a.g = gbuf;
  a.c = c; a.b = b; a.f = f; a.p = p;
  a.dc = dc; a.db = db; a.df = df; a.dp = dp;
  a.ddc = ddc; a.ddb = ddb; a.ddf = ddf; a.ddp = ddp;
  a.mse_out = mse_dev;
  // printed mse: CUDA sum e^2 / Norm (:356,:587); CPU raw sum e^2 (netlib.cpp:385); averaged over frames
  a.mse_scale = (mode == AEFFT_MODE_CPU_REF) ? (float)(1.0 / (double)B_global) : a.inv_norm;
  return launch_update(ctx, a);
}

}  // namespace aefft

extern "C" {

const char* aefft_last_error(void) { return g_err; }
int aefft_abi_version(void) { return AEFFT_ABI_VERSION; }
int64_t aefft_launch_count(const aefft_ctx* ctx) { return ctx ? ctx->launches : -1; }

int aefft_create(aefft_ctx** out, int device) {
  AE_ARG(out != nullptr);
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("aefft_create: no usable CUDA device (%s); this engine has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return AEFFT_ERR_CUDA;
  }
  AE_ARG(device >= 0 && device < n);
  AE_CUDA(cudaSetDevice(device));
  aefft_ctx* ctx = new aefft_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  AE_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  if (prop.major != 10) {
    set_error("aefft_create: device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
              prop.minor);
    delete ctx;
    return AEFFT_ERR_CUDA;
  }
  AE_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  ctx->precision = AEFFT_PRECISION_BF16X3;  // tensor-core path with fp32-grade accuracy is the default
  *out = ctx;
  return AEFFT_OK;
}

int aefft_destroy(aefft_ctx* ctx) {
  if (!ctx) return AEFFT_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm) aefft_comm_destroy(ctx);
  for (auto& kv : ctx->scratch)
    if (kv.second.p) cudaFree(kv.second.p);
  for (auto& kv : ctx->pinned)
    if (kv.second.p) cudaFreeHost(kv.second.p);
  for (auto& r : ctx->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return AEFFT_OK;
}

int aefft_sync(aefft_ctx* ctx) {
  AE_ARG(ctx);
  AE_CUDA(cudaStreamSynchronize(ctx->stream));
  return AEFFT_OK;
}

void* aefft_stream(aefft_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int aefft_set_gradient_hook(aefft_ctx* ctx, aefft_gradient_hook_fn fn, void* user) {
  AE_ARG(ctx);
  ctx->grad_hook = fn;
  ctx->grad_hook_user = user;
  return AEFFT_OK;
}

int aefft_set_bin_shard(aefft_ctx* ctx, int rank, int world) {
  AE_ARG(ctx && world >= 1 && rank >= 0 && rank < world);
  ctx->shard_rank = rank;
  ctx->shard_world = world;
  return AEFFT_OK;
}

int aefft_set_stream(aefft_ctx* ctx, void* cuda_stream) {
  AE_ARG(ctx);
  AE_CUDA(cudaSetDevice(ctx->device));
  AE_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return AEFFT_OK;
}

int aefft_set_precision(aefft_ctx* ctx, int precision) {
  AE_ARG(ctx && precision >= AEFFT_PRECISION_FP32 && precision <= AEFFT_PRECISION_BF16);
  ctx->precision = precision;
  return AEFFT_OK;
}
int aefft_get_precision(const aefft_ctx* ctx) { return ctx ? ctx->precision : -1; }

int aefft_profile_enable(aefft_ctx* ctx, int on) {
  AE_ARG(ctx);
  ctx->profiling = on != 0;
  return AEFFT_OK;
}

int aefft_profile_read(aefft_ctx* ctx, int max_rows, char* names, float* ms, int64_t* counts, double* flops,
                       double* bytes, int* n_rows) {
  AE_ARG(ctx && n_rows && max_rows >= 0);
  AE_CUDA(cudaSetDevice(ctx->device));
  AE_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<std::string> keys;
  std::vector<double> t, fl, by;
  std::vector<int64_t> cnt;
  for (auto& r : ctx->prof) {
    float dt = 0.f;
    AE_CUDA(cudaEventElapsedTime(&dt, r.e0, r.e1));
    size_t k = 0;
    for (; k < keys.size(); k++)
      if (keys[k] == r.name) break;
    if (k == keys.size()) { keys.push_back(r.name); t.push_back(0); fl.push_back(0); by.push_back(0); cnt.push_back(0); }
    t[k] += dt; fl[k] += r.flops; by[k] += r.bytes; cnt[k]++;
    ctx->event_pool.push_back(r.e0);
    ctx->event_pool.push_back(r.e1);
  }
  ctx->prof.clear();
  int n = (int)keys.size() < max_rows ? (int)keys.size() : max_rows;
  for (int k = 0; k < n; k++) {
    if (names) { strncpy(names + 64 * k, keys[k].c_str(), 63); names[64 * k + 63] = 0; }
    if (ms) ms[k] = (float)t[k];
    if (counts) counts[k] = cnt[k];
    if (flops) flops[k] = fl[k];
    if (bytes) bytes[k] = by[k];
  }
  *n_rows = n;
  return AEFFT_OK;
}

int aefft_malloc(aefft_ctx* ctx, void** dev_ptr, int64_t bytes) {
  AE_ARG(ctx && dev_ptr && bytes >= 0);
  AE_CUDA(cudaSetDevice(ctx->device));
  AE_CUDA(cudaMalloc(dev_ptr, bytes > 0 ? (size_t)bytes : 1));
  return AEFFT_OK;
}

int aefft_free(aefft_ctx* ctx, void* dev_ptr) {
  AE_ARG(ctx);
  AE_CUDA(cudaSetDevice(ctx->device));
  AE_CUDA(cudaStreamSynchronize(ctx->stream));
  AE_CUDA(cudaFree(dev_ptr));
  return AEFFT_OK;
}

int aefft_memcpy(aefft_ctx* ctx, void* dst, const void* src, int64_t bytes, int kind) {
  AE_ARG(ctx && dst && src && bytes >= 0 && kind >= 0 && kind <= 2);
  AE_CUDA(cudaSetDevice(ctx->device));
  const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  AE_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, k, ctx->stream));
  AE_CUDA(cudaStreamSynchronize(ctx->stream));
  return AEFFT_OK;
}

// ------------------------------------------------------------------------------------------ forward

int aefft_conv_fwd(aefft_ctx* ctx, int loc, int convention, int64_t B, int dD, int dM, int Nx, int Ny, int Nk, int Nl,
                   const float* in, const float* c, const float* b, float* out) {
  AE_ARG(ctx && in && c && b && out);
  AE_ARG(convention == AEFFT_CONV_CUDA || convention == AEFFT_CONV_CPU);
  AE_ARG(B > 0 && dD > 0 && dM > 0 && Nx > 0 && Ny > 0 && Nk > 0 && Nl > 0);
  AE_CUDA(cudaSetDevice(ctx->device));
  Stage st(ctx, loc);
  const float *din, *dc, *db;
  float* dout;
  const size_t P = (size_t)Nx * Ny;
  AE_TRY(st.in("st_in", in, (size_t)B * dD * P, &din));
  AE_TRY(st.in("st_c", c, (size_t)dM * dD * Nk * Nl, &dc));
  AE_TRY(st.in("st_b", b, (size_t)dM, &db));
  AE_TRY(st.inout("st_out", out, (size_t)B * dM * P, &dout, false));
  // Conv_gpu divides the input by dM on the host (backproplib.cu:134); Conv does not (netlib.cpp:346)
  const float pre_div = convention == AEFFT_CONV_CUDA ? (float)dM : 0.f;
  AE_TRY(launch_conv(ctx, fwd_window(Nk, Nl, convention), B, dD, dM, Nx, Ny, din, nullptr, pre_div, dc,
                     (int64_t)dD * Nk * Nl, (int64_t)Nk * Nl, db, dout));
  return st.finish();
}

int aefft_pool(aefft_ctx* ctx, int loc, int64_t B, int D, int Nx, int Ny, int oNx, int oNy, int scale, const float* in,
               float* out) {
  AE_ARG(ctx && in && out && scale != 0 && B > 0 && D > 0 && Nx > 0 && Ny > 0 && oNx > 0 && oNy > 0);
  AE_CUDA(cudaSetDevice(ctx->device));
  Stage st(ctx, loc);
  const float* din;
  float* dout;
  AE_TRY(st.in("st_in", in, (size_t)B * D * Nx * Ny, &din));
  AE_TRY(st.inout("st_out", out, (size_t)B * D * oNx * oNy, &dout, false));
  AE_TRY(launch_pool(ctx, B, D, Nx, Ny, oNx, oNy, scale, din, dout));
  return st.finish();
}

int aefft_portion(aefft_ctx* ctx, int loc, int64_t B, int D, int Nx, int Ny, int q, const float* in, float* out) {
  AE_ARG(ctx && in && out && q >= 1 && B > 0 && D > 0 && Nx >= q && Ny >= q);
  AE_CUDA(cudaSetDevice(ctx->device));
  Stage st(ctx, loc);
  const float* din;
  float* dout;
  AE_TRY(st.in("st_in", in, (size_t)B * D * Nx * Ny, &din));
  AE_TRY(st.inout("st_out", out, (size_t)B * D * (Nx / q) * (Ny / q), &dout, false));
  AE_TRY(launch_portion(ctx, B, D, Nx, Ny, q, din, dout));
  return st.finish();
}

// ------------------------------------------------------------------------------------------ training

int64_t aefft_coord_gbuf_len(int mode, int dD, int dM, int Nk, int Nl) { return gbuf_len(mode, dD, dM, Nk, Nl); }

int aefft_coord_gradients(aefft_ctx* ctx, int mode, int quirks, int64_t B, int dD, int dM, int Nx, int Ny, int Nk,
                          int Nl, const float* in, const float* out, const float* hin, const float* c, const float* f,
                          float* gbuf) {
  (void)c;
  AE_ARG(ctx && in && out && hin && f && gbuf);
  AE_CUDA(cudaSetDevice(ctx->device));
  return coord_gradients_dev(ctx, mode, quirks, B, dD, dM, Nx, Ny, Nk, Nl, in, out, hin, f, gbuf);
}

int aefft_coord_update(aefft_ctx* ctx, int mode, int64_t B_global, int dD, int dM, int Nx, int Ny, int Nk, int Nl,
                       const float* gbuf, float* c, float* b, float* f, float* p, float* dc, float* db, float* df,
                       float* dp, float* ddc, float* ddb, float* ddf, float* ddp, float delmax, float alpha,
                       float* mse_dev) {
  AE_ARG(ctx && gbuf && c && b && f && p);
  AE_CUDA(cudaSetDevice(ctx->device));
  return coord_update_dev(ctx, mode, B_global, dD, dM, Nx, Ny, Nk, Nl, gbuf, c, b, f, p, dc, db, df, dp, ddc, ddb, ddf,
                          ddp, delmax, alpha, mse_dev);
}

int aefft_backprop_coord(aefft_ctx* ctx, int loc, int mode, int quirks, int64_t B, int dD, int dM, int Nx, int Ny,
                         int Nk, int Nl, const float* in, const float* out, const float* hin, float* c, float* b,
                         float* f, float* p, float* dc, float* db, float* df, float* dp, float* ddc, float* ddb,
                         float* ddf, float* ddp, float delmax, float alpha, int active, float* mse) {
  (void)active;  // quirk C5: adapt_rate ends with del = delmax (backproplib.cu:34)
  AE_ARG(ctx && in && out && hin && c && b && f && p);
  AE_ARG(B > 0 && dD > 0 && dM > 0 && Nx > 0 && Ny > 0 && Nk > 0 && Nl > 0);
  AE_CUDA(cudaSetDevice(ctx->device));
  Stage st(ctx, loc);
  const size_t P = (size_t)Nx * Ny, nC = (size_t)dM * dD * Nk * Nl;
  const float *din, *dout, *dhin;
  float *d_c, *d_b, *d_f, *d_p, *d_dc, *d_db, *d_df, *d_dp, *d_ddc, *d_ddb, *d_ddf, *d_ddp;
  AE_TRY(st.in("bp_in", in, (size_t)B * dD * P, &din));
  AE_TRY(st.in("bp_out", out, (size_t)B * dD * P, &dout));
  AE_TRY(st.in("bp_hin", hin, (size_t)B * dM * P, &dhin));
  AE_TRY(st.inout("bp_c", c, nC, &d_c));
  AE_TRY(st.inout("bp_b", b, dM, &d_b));
  AE_TRY(st.inout("bp_f", f, nC, &d_f));
  AE_TRY(st.inout("bp_p", p, dD, &d_p));
  // momentum (dc..dp) and last-gradient (ddc..ddp) buffers exist only in the CUDA modes (backproplib.cu:387-412); the CPU
  // path (netlib.cpp:437-444) has neither, so they are neither staged nor written back there
  const bool cuda_mode = mode != AEFFT_MODE_CPU_REF;
  d_dc = d_db = d_df = d_dp = d_ddc = d_ddb = d_ddf = d_ddp = nullptr;
  if (cuda_mode || loc == AEFFT_DEVICE) {
    AE_TRY(st.inout("bp_dc", dc, nC, &d_dc));
    AE_TRY(st.inout("bp_db", db, dM, &d_db));
    AE_TRY(st.inout("bp_df", df, nC, &d_df));
    AE_TRY(st.inout("bp_dp", dp, dD, &d_dp));
    AE_TRY(st.inout("bp_ddc", ddc, nC, &d_ddc, false));
    AE_TRY(st.inout("bp_ddb", ddb, dM, &d_ddb, false));
    AE_TRY(st.inout("bp_ddf", ddf, nC, &d_ddf, mode == AEFFT_MODE_CUDA_REF_SYM));  // untouched in the tied path
    AE_TRY(st.inout("bp_ddp", ddp, dD, &d_ddp, false));
  }
  float* gbuf;
  AE_TRY(ctx->getT("bp_gbuf", (size_t)gbuf_len(mode, dD, dM, Nk, Nl), &gbuf));
  float* mse_dev;
  AE_TRY(ctx->getT("bp_mse", 1, &mse_dev));
  AE_TRY(coord_gradients_dev(ctx, mode, quirks, B, dD, dM, Nx, Ny, Nk, Nl, din, dout, dhin, d_f, gbuf));
  AE_TRY(coord_update_dev(ctx, mode, B, dD, dM, Nx, Ny, Nk, Nl, gbuf, d_c, d_b, d_f, d_p, d_dc, d_db, d_df, d_dp, d_ddc,
                          d_ddb, d_ddf, d_ddp, delmax, alpha, mse_dev));
  if (mse) {
    AE_CUDA(cudaMemcpyAsync(mse, mse_dev, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return st.finish();
}

// ------------------------------------------------------------------------------------------ glue (host only)

int aefft_init_conv(float* c, float* b, int mS, int dD, int kS, int lS, float rmax) {
  // netlib.cpp:167-197: draw order m,d,k,l then b[m]; r = -max + 2*max*(float)rand()/(float)RAND_MAX
  AE_ARG(c && b && mS > 0 && dD > 0 && kS > 0 && lS > 0);
  size_t n = 0;
  for (int m = 0; m < mS; m++) {
    for (int i = 0; i < dD * kS * lS; i++) c[n++] = -rmax + 2 * rmax * (float)rand() / (float)RAND_MAX;
    b[m] = -rmax + 2 * rmax * (float)rand() / (float)RAND_MAX;
  }
  return AEFFT_OK;
}

int aefft_saveload_conv(const char* dir, float* c, float* b, int dM, int dD, int Nk, int Nl, int scale, int L, int io,
                        int write) {
  // netlib.cpp:220-272: raw little-endian float32, c flattened [m][d][k][l] then b[dM], no header
  AE_ARG(dir && c && b && dM > 0 && dD > 0 && Nk > 0 && Nl > 0);
  std::string path = std::string(dir) + "/C_weights_" + std::to_string(L) + (io == 0 ? "_in" : "_out") +
                     "_D=" + std::to_string(dD) + "_M=" + std::to_string(dM) +
                     "_Lk=" + std::to_string((Nk - 1) / 2 - 1) + "_Ll=" + std::to_string((Nl - 1) / 2 - 1) +
                     "_S=" + std::to_string(scale) + ".conv";
  const size_t nC = (size_t)dM * dD * Nk * Nl;
  if (write == 1) {
    std::ofstream file(path, std::ios::out | std::ios::binary);
    if (!file) { set_error("cannot open %s for writing", path.c_str()); return AEFFT_ERR_IO; }
    file.write(reinterpret_cast<const char*>(c), nC * sizeof(float));
    file.write(reinterpret_cast<const char*>(b), (size_t)dM * sizeof(float));
    if (!file) { set_error("short write to %s", path.c_str()); return AEFFT_ERR_IO; }
  } else {
    std::ifstream file(path, std::ios::in | std::ios::binary);
    if (!file) { set_error("cannot open %s", path.c_str()); return AEFFT_ERR_IO; }  // reference: silent zeros (N5)
    file.read(reinterpret_cast<char*>(c), nC * sizeof(float));
    file.read(reinterpret_cast<char*>(b), (size_t)dM * sizeof(float));
    if (!file) { set_error("short read from %s", path.c_str()); return AEFFT_ERR_IO; }
  }
  return AEFFT_OK;
}

int aefft_load_param(const char* path, int* dM, int* Lk, int* Ll, int* scal, float* rmax) {
  // netlib.cpp:274-289: positional `name value` pairs
  AE_ARG(path && dM && Lk && Ll && scal && rmax);
  std::ifstream file(path);
  if (!file) { set_error("cannot open %s", path); return AEFFT_ERR_IO; }
  std::vector<float> values;
  std::string name;
  float v;
  while (file >> name >> v) values.push_back(v);
  if (values.size() < 5) { set_error("%s: expected 5 name/value pairs, found %zu", path, values.size()); return AEFFT_ERR_IO; }
  *dM = (int)values[0]; *Lk = (int)values[1]; *Ll = (int)values[2]; *scal = (int)values[3]; *rmax = values[4];
  return AEFFT_OK;
}

int aefft_synth_frames(aefft_ctx* ctx, int loc, uint64_t seed, int64_t b0, int64_t B, int D, int Nx, int Ny,
                       float* out) {
  AE_ARG(ctx && out && B > 0 && D > 0 && Nx > 0 && Ny > 0);
  AE_CUDA(cudaSetDevice(ctx->device));
  Stage st(ctx, loc);
  float* d;
  AE_TRY(st.inout("st_out", out, (size_t)B * D * Nx * Ny, &d, false));
  AE_TRY(launch_synth(ctx, seed, b0, B, D, Nx, Ny, d));
  return st.finish();
}

}  // extern "C"
