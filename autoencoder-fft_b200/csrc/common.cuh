// Internal definitions shared by the engine's translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>

#include "aefft.h"

namespace aefft {

void set_error(const char* fmt, ...);

#define AE_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      aefft::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return AEFFT_ERR_CUDA;                                                                  \
    }                                                                                         \
  } while (0)

#define AE_TRY(call)            \
  do {                          \
    int r__ = (call);           \
    if (r__ != AEFFT_OK) return r__; \
  } while (0)

#define AE_ARG(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      aefft::set_error("%s:%d bad argument: %s", __FILE__, __LINE__, #cond); \
      return AEFFT_ERR_ARG;                                                 \
    }                                                                       \
  } while (0)

// Grow-only device scratch slots keyed by name: the reference cudaMallocs/frees inside every call
// (backproplib.cu:150-151, fft_backproplib.cu:1394-1427); here workspaces persist in the ctx.
struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
};

// One record per kernel launch while profiling is on (aefft_profile_enable): CUDA events on the launching stream
// plus the ALGORITHMIC work of that launch (flops, bytes), which bench.py turns into the roofline numbers.
struct ProfRec {
  const char* name;
  cudaEvent_t e0, e1;
  double flops, bytes;
};

}  // namespace aefft

struct aefft_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int64_t launches = 0;
  int precision = AEFFT_PRECISION_FP32;  // arithmetic of the coordinate-space contractions (aefft_set_precision)
  bool profiling = false;
  int shard_rank = 0, shard_world = 1;  // frequency-bin sharding of aefft_backprop_fft (aefft_set_bin_shard)
  aefft_gradient_hook_fn grad_hook = nullptr;  // data-parallel momentum-space training (aefft_set_gradient_hook)
  void* grad_hook_user = nullptr;
  void* comm = nullptr;                 // ncclComm_t of this rank (aefft_comm_init); collectives run on `stream`
  int comm_rank = 0, comm_world = 1;
  std::vector<aefft::ProfRec> prof;
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t get_event();
  std::map<std::string, aefft::Scratch> scratch;
  std::map<std::string, aefft::Scratch> pinned;
  // cudaFuncAttributeMaxDynamicSharedMemorySize already raised on THIS ctx's device, per kernel function (the attribute
  // is per device: a process-wide "already set" flag would leave the kernels of a second device at the 48 KB default)
  std::map<const void*, size_t> dyn_smem;
  int ensure_dyn_smem(const void* func, size_t bytes);

  int get(const char* name, size_t bytes, void** out);         // device scratch
  int get_pinned(const char* name, size_t bytes, void** out);  // pinned host staging
  template <class T>
  int getT(const char* name, size_t n, T** out) {
    return get(name, n * sizeof(T), (void**)out);
  }
};

namespace aefft {

// RAII event pair around one kernel launch (no-op unless ctx->profiling).
struct ProfScope {
  aefft_ctx* ctx;
  int idx = -1;
  ProfScope(aefft_ctx* c, const char* name, double flops, double bytes) : ctx(c) {
    if (!c->profiling) return;
    ProfRec r{name, c->get_event(), c->get_event(), flops, bytes};
    cudaEventRecord(r.e0, c->stream);
    idx = (int)c->prof.size();
    c->prof.push_back(r);
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(ctx->prof[idx].e1, ctx->stream);
  }
};

// ---- conv geometry -------------------------------------------------------------------------------
// Every coordinate-space contraction is expressed in "window" coordinates: output pixel (i,j) reads source
// pixels (i + ai0 + tk, j + aj0 + tl), tk in [0,Nk), tl in [0,Nl); the weight tap multiplying window position
// (tk,tl) is (k,l) = flip ? (Nk-1-tk, Nl-1-tl) : (tk,tl).  Source pixels outside [lo, N) contribute zero.
struct Window {
  int Nk, Nl;
  int ai0, aj0;  // window origin relative to the output pixel
  int flip;      // 1: forward conv (in[i - ik(k)]), 0: transposed (e[u + ik(k)])
  int lo;        // lowest valid source index (0: CUDA `>=0`, 1: CPU strict `>0`)
};

// reference tap offset ik(k) = base + k  
inline int tap_base(int N, int convention) {
  int a = (N - 1) / 2 - 1;
  if (convention == AEFFT_CONV_CUDA) a = a / 2;
  return -2 * a - 1;
}
// forward conv window: source index i - (base+k), k=0..N-1  ->  origin i - base - (N-1), flipped
inline Window fwd_window(int Nk, int Nl, int convention) {
  Window w;
  w.Nk = Nk; w.Nl = Nl;
  w.ai0 = -tap_base(Nk, convention) - (Nk - 1);
  w.aj0 = -tap_base(Nl, convention) - (Nl - 1);
  w.flip = 1;
  w.lo = (convention == AEFFT_CONV_CPU) ? 1 : 0;
  return w;
}
// transposed window: source index u + (base+k)
inline Window tr_window(int Nk, int Nl, int convention) {
  Window w;
  w.Nk = Nk; w.Nl = Nl;
  w.ai0 = tap_base(Nk, convention);
  w.aj0 = tap_base(Nl, convention);
  w.flip = 0;
  w.lo = 0;
  return w;
}

// ---- kernel launchers (conv_kernels.cu / wgrad_kernels.cu / misc_kernels.cu) ------------------------

// out[b][o][i][j] = bias[o] + sum_{c,tk,tl} W(o,c,k,l) * S[b][c](i+ai0+tk, j+aj0+tl)
//   S = src0 * (pre_div ? 1/pre_div : 1)           when src1 == nullptr
//   S = src0 - src1                                 otherwise (fused e = out - in)
//   W(o,c,k,l) = w[o*w_so + c*w_sc + k*Nl + l]
int launch_conv(aefft_ctx* ctx, const Window& win, int64_t B, int C, int O, int Nx, int Ny, const float* src0,
                const float* src1, float pre_div, const float* w, int64_t w_so, int64_t w_sc, const float* bias,
                float* out);

// Tensor-core (tcgen05) implementation of the same contract (conv_tc.cu); passes = 3 (BF16X3) or 1 (BF16).
// Returns AEFFT_ERR_UNSUPPORTED for shapes outside its envelope (the caller then uses the fp32 kernel).
int launch_conv_tc(aefft_ctx* ctx, const Window& win, int64_t B, int C, int O, int Nx, int Ny, const float* src0,
                   const float* src1, float pre_div, const float* w, int64_t w_so, int64_t w_sc, const float* bias,
                   float* out, int passes);

// Warp-specialised pipelined variant (conv_tc_ws.cu), same contract; needs all weights resident in shared memory.
int launch_conv_tc_ws(aefft_ctx* ctx, const Window& win, int64_t B, int C, int O, int Nx, int Ny, const float* src0,
                      const float* src1, float pre_div, const float* w, int64_t w_so, int64_t w_sc, const float* bias,
                      float* out, int passes);

// Row-streaming variant (conv_rs.cu): window rows stacked along N, TMA-fed row rings; same contract.
int launch_conv_rs(aefft_ctx* ctx, const Window& win, int64_t B, int C, int O, int Nx, int Ny, const float* src0,
                   const float* src1, float pre_div, const float* w, int64_t w_so, int64_t w_sc, const float* bias,
                   float* out, int passes);

// Correlation (weight-gradient) contraction, summed over all frames and pixels:
//   G[a][x][k][l] = sum_{b,i,j} A[b][a](i,j) * X[b][x](i+ai0+tk, j+aj0+tl)          (k,l) <-> (tk,tl) per win.flip
//   sumA[a] = sum A[b][a](i,j);  sumsq = sum A^2
// A operand modes:
//   A_PLAIN : A = a0                       (nA channels)
//   A_DIFF  : A = a0 - a1                  (e = out - in)
//   A_SHIFT : virtual channels a=(d1,tk1,tl1): A = Vlo(i,j) * (a0-a1)[d1](i+ei0+tk1, j+ej0+tl1)   (CPU_REF R tensor)
enum { A_PLAIN = 0, A_DIFF = 1, A_SHIFT = 2 };
struct AOperand {
  int mode = A_PLAIN;
  const float* a0 = nullptr;
  const float* a1 = nullptr;
  int nA = 0;       // logical channel count (for A_SHIFT: d1 count * Nk * Nl)
  int src_ch = 0;   // physical channels of a0/a1 per frame
  int ei0 = 0, ej0 = 0, eNk = 0, eNl = 0, out_lo = 0;  // A_SHIFT only
};
// Results land in ctx scratch and are finished by reduce_wgrad: G (nA*nX*Nk*Nl floats), sumA (nA), sumsq (1).
int launch_wgrad(aefft_ctx* ctx, const Window& win, int64_t B, int Nx, int Ny, const AOperand& A, const float* X,
                 int nX, float* G, float* sumA, float* sumsq);

// Tensor-core weight gradients of one pair (wgrad_tc.cu): G = [GC dM*dD*T | GF dD*dM*T] raw sums over the frames,
// GC = corr(dh, in), GF = corr(out-in, hin) with the forward window `win`.  AEFFT_ERR_UNSUPPORTED outside its envelope.
int launch_wgrad_tc(aefft_ctx* ctx, const Window& win, int64_t B, int dD, int dM, int Nx, int Ny, const float* in,
                    const float* out, const float* hin, const float* dh, float* G, int passes);
// Streaming TMEM-operand variant (wgrad_ts.cu): GC | GF into G plus GB = sum dh, GP = sum (out-in), SQ = sum (out-in)^2
// in the same pass.  AEFFT_ERR_UNSUPPORTED outside its envelope.
int launch_wgrad_ts(aefft_ctx* ctx, const Window& win, int64_t B, int dD, int dM, int Nx, int Ny, const float* in,
                    const float* out, const float* hin, const float* dh, float* G, float* GB, float* GP, float* SQ,
                    int passes);
// sum[c] = sum_{b,pix} (a0-a1)[b][c], *sumsq = sum (a0-a1)^2 (double accumulation, deterministic); a1/sum/sumsq optional
int launch_channel_sums(aefft_ctx* ctx, int64_t B, int ch, int Nx, int Ny, const float* a0, const float* a1, float* sum,
                        float* sumsq);

int launch_pool(aefft_ctx* ctx, int64_t B, int D, int Nx, int Ny, int oNx, int oNy, int scale, const float* in,
                float* out);
int launch_portion(aefft_ctx* ctx, int64_t B, int D, int Nx, int Ny, int q, const float* in, float* out);
int launch_synth(aefft_ctx* ctx, uint64_t seed, int64_t b0, int64_t B, int D, int Nx, int Ny, float* out);

// T[d][k1][l1] = sum_{b,i,j} (out-in)[b][d](i,j) * V_lo(i - (bi+k1), j - (bj+l1))   (border-aware sums of e)
int launch_border_sums(aefft_ctx* ctx, int64_t B, int D, int Nx, int Ny, int Nk, int Nl, int bi, int bj, int lo,
                       const float* out, const float* in, float* T);

// CUDA_REF quirks C3+C4: bug-compatible dF (see misc_kernels.cu)
int launch_quirk_dF(aefft_ctx* ctx, int quirks, int64_t B, int dD, int dM, int Nx, int Ny, int Nk, int Nl,
                    const float* out, const float* in, const float* hin, float* gF /*[dD][dM][Nk][Nl], summed*/);

struct UpdateArgs {
  int mode, dD, dM, Nk, Nl;
  float inv_norm;  // 1/(Norm * B_global)
  float delmax, alpha;
  const float* g;  // gradient block (layout per mode, see capi.cu)
  float *c, *b, *f, *p, *dc, *db, *df, *dp, *ddc, *ddb, *ddf, *ddp;
  float* mse_out;  // device scalar or nullptr
  float mse_scale;
};
int launch_update(aefft_ctx* ctx, const UpdateArgs& a);

// ---- momentum-space launchers (fft_kernels.cu / spectral_kernels.cu) ---------------------------------
// batched 2-D R2C / C2R, unnormalised, n = {Nx, Ny} powers of two; spectra [batch][Nx][Ny/2+1] complex64
int launch_fft_r2c(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, const float* in, float2* spec);
int launch_fft_c2r(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, const float2* spec, float2* work, float* out,
                   float scale);
// frame-strided real side: image (frame f, channel c) at base + f*fstride + c*Nx*Ny, batch = frames*ch (the per-frame layer
// blocks of the net); AEFFT_ERR_UNSUPPORTED when that layout cannot be addressed directly (caller gathers instead)
int launch_fft_r2c_strided(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, const float* in, int ch, long long fstride,
                           float2* spec);
int launch_fft_c2r_strided(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, const float2* spec, float2* work, float* out, int ch,
                           long long fstride, float scale);
// transforms fused with the adjacent spectral pooling (AEFFT_ERR_UNSUPPORTED outside the compile-time-length envelope):
// real [batch][Nx][Ny] -> pooled half spectrum [batch][Nxs][Nys/2+1];  small half spectrum -> real image of its embedding
int launch_fft_r2c_pooled(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, int Nxs, int Nys, const float* in, float2* tmp,
                          float2* out);
int launch_fft_c2r_embedded(aefft_ctx* ctx, int64_t batch, int Nx, int Ny, int Nxs, int Nys, const float2* spec, float2* tmp,
                            float* out, float scale);
int launch_spec_resize(aefft_ctx* ctx, int64_t planes, int Nx, int Ny, int Nxs, int Nys, const float2* in, float2* out);
int launch_spec_contract(aefft_ctx* ctx, int64_t B, int C, int O, int64_t S, const float2* in0, const float2* in1,
                         const float2* W, int64_t w_so, int64_t w_sc, int conjW, float in_scale, const float* bias,
                         float bias_scale, float2* out);
int launch_spec_outer(aefft_ctx* ctx, int64_t B, int nA, int nC, int64_t S, const float2* A0, const float2* A1,
                      const float2* Bm, float bm_alpha, const float* bm_bias, float bm_bias_scale, float scale, float2* out);
int launch_spec_dc_sums(aefft_ctx* ctx, int64_t B, int dM, int dD, int64_t S, const float2* G, const float2* O,
                        const float2* Xt, float* db, float* dp, float gscale);
// W_N^t = exp(-2 pi i t / N), t = 0..N-1, computed in double on the host once per length (device table, L1 resident)
// transform lengths the FFT kernels take: even, 2^a 3^b 5^c, <= 8192 (fft_kernels.cu)
bool fft_len_supported(int N);
// (x mod N) for the twiddle tables: a mask for powers of two, else the remainder (x >= 0)
__host__ __device__ __forceinline__ int tw_mod(int x, int N) { return (N & (N - 1)) == 0 ? (x & (N - 1)) : (x % N); }
// index of W_N^(w * (k - half)) in the N-entry twiddle table (k - half may be negative)
__host__ __device__ __forceinline__ int tw_index(int w, int k_minus_half, int N) {
  const int s = k_minus_half < 0 ? k_minus_half + N : k_minus_half;
  return tw_mod(w * s, N);
}
int get_twiddles(aefft_ctx* ctx, int N, const float2** out);
// Pruned DFTs of the Nk x Nl-tap kernels (replace pad_k + full-size R2C and full-size C2R + shrink_k, which move
// >= 20 bytes per bin to obtain / consume 25 numbers per image):
//   spectrum : spec[n][wx][wy] = sum_{k,l} taps[n][k][l] W_Nx^(wx*i_k) W_Ny^(wy*j_l),  (i_k, j_l) = ((k-Nk/2) mod Nx, (l-Nl/2) mod Ny)
//   taps     : taps[n][k][l] = scale * sum_{wx,wy} h(wy) Re( spec[n][wx][wy] conj(W_Nx^(wx*i_k)) conj(W_Ny^(wy*j_l)) ),
//              h = 1 for wy in {0, Ny/2}, else 2  (= shrink_k(C2R(spec)) with the C2R's Hermitian convention)
int launch_kernel_spectrum_direct(aefft_ctx* ctx, int64_t n_img, int Nx, int Ny, int Nk, int Nl, const float* taps, float2* spec,
                                  int col0 = 0, int ncols = 0);  // column slab [col0, col0+ncols) (0: whole half spectrum)
int launch_spectrum_to_taps(aefft_ctx* ctx, int64_t n_img, int Nx, int Ny, int Nk, int Nl, const float2* spec, float* taps,
                            float scale, int col0 = 0, int ncols = 0);
int launch_spec_slab(aefft_ctx* ctx, int64_t n_img, int Nx, int Ny, const float2* full, float2* slab, int col0, int ncols);
int launch_pad(aefft_ctx* ctx, int64_t n_img, int Nx, int Ny, int Nk, int Nl, const float* taps, float* img);
int launch_shrink(aefft_ctx* ctx, int64_t n_img, int Nx, int Ny, int Nk, int Nl, const float* img, float* taps);
int launch_fft_update(aefft_ctx* ctx, int dM, int dD, int Nk, int Nl, float* c, float* f, float* b, float* p,
                      const float* dck, const float* dfk, const float* db, const float* dp, float* Dc, float* Df, float* Db,
                      float* Dp, float del, int maxdiff, float* div_scratch);
// tcgen05 form for 5 x 5 kernels (gdiff_tc.cu); AEFFT_ERR_UNSUPPORTED: the caller runs the CUDA-core kernel
int launch_gradient_diff_tc(aefft_ctx* ctx, int dM, int dD, const float* c, const float* f, float* cd, float* fd, int t0, int nt,
                            int nchunks, float* part);
int launch_gradient_diff(aefft_ctx* ctx, int dM, int dD, int Nk, int Nl, const float* c, const float* f, const float* b,
                         const float* p, float* div, int rank, int world);
int launch_axpby(aefft_ctx* ctx, float* y, const float* x, float a, float b, long long n);
int launch_spec_mse(aefft_ctx* ctx, int64_t B, int dD, int dM, int Nx, int Ny, const float2* Xt, const float2* O, float* out,
                    int col0 = 0, int ncols = 0);

// ---- momentum-space contractions on the tensor cores (spec_tc.cu): bin-major [bin][row][col] fp32 operands, interleaved
// complex rows, embedded weight blocks [[Wr, -Wi], [Wi, Wr]]; see the header of spec_tc.cu
bool spec_tc_eligible(int dD, int dM, int Nk, int Nl);
// [R][S] complex (bins fastest) -> [S][R] complex, optionally in0 - in1
int launch_to_binmajor(aefft_ctx* ctx, long long R, long long S, const float2* in0, const float2* in1, float2* out);
// emb[w][2r+a][2c+b] for the R x C kernels `taps` [R][C][Nk][Nl] on the slab [col0, col0+ncols) (ncols <= 0: whole half spectrum)
int launch_kernel_spectrum_emb_pooled(aefft_ctx* ctx, int R, int C, int Nk, int Nl, int Nx, int Ny, int Nxm, int Nym,
                                      const float* taps, float* emb);
int launch_kernel_spectrum_emb(aefft_ctx* ctx, int R, int C, int Nk, int Nl, int Nx, int Ny, int col0, int ncols, const float* taps,
                               float* emb);
int launch_binmajor_to_taps(aefft_ctx* ctx, int R, int C, int transpose, int Nk, int Nl, int Nx, int Ny, int col0, int ncols,
                            const float2* z, float* taps, float scale);
int launch_tc_forward(aefft_ctx* ctx, long long S, int B, int C, int O, const float* in, const float* Wemb, float scale,
                      const float* bias, float bias_scale, const float* sub, float* out, float* mse_out, double mse_scale, int ncols,
                      int col0, int Ny);
int launch_tc_adjoint(aefft_ctx* ctx, long long S, int B, int dD, int dM, const float* E, const float* Femb, float* G);
int launch_tc_outer(aefft_ctx* ctx, long long S, int B, int nP, int nQ, const float* P, const float* Q, float scale, int conj_out,
                    float* out);
// backprop_fft's iteration loop on per-bin Gram matrices (spec_gram.cu; expout == in): ONE pass over the frames per call
bool spec_gram_loop_pays(int B, int dD, int dM, bool bin_major);
// (Nx, sNx, sNy: O is compact on the support grid (sNx, sNy) of an up-sampling, zero on the other bins; sNx = 0: dense)
int launch_gram_stats_bm(aefft_ctx* ctx, long long S, int B, int dD, const float* X, const float* O, int sub, float* Gx, float* M0,
                         float* mse_out, double mse_scale, float* dcsum, int ncols, int col0, int Ny, int Nx = 0, int sNx = 0,
                         int sNy = 0);
int launch_gram_iter_bm(aefft_ctx* ctx, long long S, int B, int dD, int dM, const float* Gx, const float* M0, const float* Cemb,
                        const float* Femb, int first, float gs, float gb, float norm, const float* dcsum, const float* bias_b,
                        const float* bias_p, float* dCt, float* dFt, float* db, float* dp, float* mse_out, double mse_scale, int ncols,
                        int col0, int Ny);
int launch_gram_stats_ff(aefft_ctx* ctx, long long S, int B, int dD, const float2* X, const float2* O, float2* Gx, float2* M0,
                         float* mse_out, double mse_scale, float* dcsum, int ncols, int col0, int Ny, int Nx = 0, int sNx = 0,
                         int sNy = 0);
int launch_gram_iter_ff(aefft_ctx* ctx, long long S, int B, int dD, int dM, const float2* Gx, const float2* M0, const float2* C,
                        const float2* F, int first, float gs, float gb, float norm, const float* dcsum, const float* bias_b,
                        const float* bias_p, float2* dC, float2* dF, float* db, float* dp, float* mse_out, double mse_scale, int ncols,
                        int col0, int Ny);
// Gram form of the gradients (spec_tc.cu): both gradient spectra of a bin from Mg = sum_b E conj(X), E / X bin-major
bool spec_tc_gram_pays(int B, int dD, int dM);
int launch_tc_gram_grad(aefft_ctx* ctx, long long S, int B, int dM, int dD, const float* E, const float* X, const float* Cemb,
                        const float* Femb, float gs, float* dCt, float* dFt);
int launch_tc_dc_terms_gram(aefft_ctx* ctx, int B, int dM, int dD, const float* E, const float* Femb, const float* bias_b, float* dFt,
                            float* db, float* dp, float gs, float fs, float norm);
int launch_tc_dc_terms(aefft_ctx* ctx, int B, int dM, int dD, const float* G, const float* E, const float* bias_b, float* dFt,
                       float* db, float* dp, float gs, float fs, float corr_scale);

// ---- fused iteration for pairs with <= 4 input channels (spec_small.cu), bins-fastest spectra
bool spec_small_eligible(int dD, int dM);
// forward contraction with the CO x CI weight block of a bin in registers; AEFFT_ERR_UNSUPPORTED when (CI, CO) has no
// instantiation (W [CO][CI][S], in [B][CI][S], out [B][CO][S])
int launch_spec_conv_reg(aefft_ctx* ctx, int64_t B, int CI, int CO, int64_t S, const float2* in, const float2* W, const float* bias,
                         float bias_scale, float in_scale, float2* out);
bool spec_conv_reg_supported(int CI, int CO);
// the same contraction at resolution (Nxb, Nyb) fused with the spectral pooling after it (pooled_out: out is the cropped
// (Nxm, Nym) spectrum) or with the spectral up-sampling before it (!pooled_out: in is the small spectrum, out the big one)
// small_bin_major: the small spectrum is bin-major [bin][frame][channel]
int launch_spec_conv_reg_resized(aefft_ctx* ctx, int64_t B, int CI, int CO, int Nxb, int Nyb, int Nxm, int Nym, bool pooled_out,
                                 bool small_bin_major, const float2* in, const float2* W, const float* bias, float bias_scale,
                                 float in_scale, float2* out);
int launch_spec_conv_reg_support(aefft_ctx* ctx, int64_t B, int CI, int CO, int Nxb, int Nyb, int Nxm, int Nym, bool in_bin_major,
                                 bool out_bin_major, const float2* in, const float2* W, const float* bias, float bias_scale,
                                 float in_scale, float2* out);
int launch_small_grad(aefft_ctx* ctx, int64_t B, int dD, int dM, int64_t S, const float2* X, const float2* Xt, const float2* O,
                      const float2* C, const float2* F, const float* bias_b, const float* bias_p, float norm, float gscale,
                      float dbscale, float2* dC, float2* dF, float* db, float* dp);
int launch_small_mse(aefft_ctx* ctx, int64_t B, int dD, int dM, int64_t S, const float2* X, const float2* Xt, const float2* C,
                     const float2* F, const float* bias_b, const float* bias_p, float norm, float* mse_out, double mse_scale,
                     int ncols, int col0, int Ny);

// E~ = O~ - X~ on bin-major spectra [S][rowlen] and *mse_out = mse_scale * sum_bins hw(bin) |E|^2 (mse_out may be null)
int launch_bm_sub_mse(aefft_ctx* ctx, long long S, long long rowlen, const float* O, const float* X, float* E, float* mse_out,
                      double mse_scale, int ncols, int col0, int Ny);
// bin-major spectral pooling (resize, fft_backproplib.cu:87-157, on [bin][rowlen] data): whole rows are gathered
int launch_bm_resize(aefft_ctx* ctx, long long rowlen, int Nx, int Ny, int Nxs, int Nys, const float* in, float* out);

// backprop_fft (fft_backproplib.cu:1381-1511) on one of three input forms (fft_capi.cu)
struct FftTrainInputs {
  const float *in = nullptr, *expout = nullptr, *out = nullptr;  // real-space frames [B][dD][Nx][Ny] (per `loc`) ...
  int64_t fstride = 0;                                           // ... or per-frame blocks `fstride` floats apart (device)
  const float2 *Xs = nullptr, *Os = nullptr;   // device spectra, bins-fastest [B][dD][Nx][Nyr]  (expout = in); under bin
                                               // sharding: this device's column slab [B][dD][Nx][ncols] of ALL frames
  const float *Xbm = nullptr, *Obm = nullptr;  // device spectra, bin-major [bin][B][2 dD]         (expout = in)
  const float* Hbm = nullptr;                  // optional with Xbm: conv_k(X; c, b) of the CURRENT kernels, bin-major
                                               // [bin][B][2 dM] (the forward's hidden layer): saves its recomputation
  const float *Cemb = nullptr, *Femb = nullptr; // optional with Xbm: embedded bin-major spectra of the CURRENT c / f (what the
                                               // forward generated for its two convs of this pair): not generated again
  int o_sNx = 0, o_sNy = 0;   // > 0: Os / Obm is COMPACT on the (o_sNx, o_sNy) grid an up-sampling fills (zero elsewhere);
                              // only the Gram-matrix loop takes it (spec_gram.cu)
  bool resident = false;      // c,f,b,p are the device-resident masters: no export through the spectra, and no stream
                              // synchronisation unless a host trace is requested
  float* trace_dev = nullptr; // device destination of the mse trace (n_iter + 1 floats), optional
};
int backprop_fft_run(aefft_ctx* ctx, int loc, int64_t B, int dD, int dM, int Nx, int Ny, int Nk, int Nl, const FftTrainInputs& inp,
                     float* cfreq, float* c, float* ffreq, float* f, float* b, float* p, float del0, int maxdiff, int n_iter,
                     float* mse_trace);
// kernel spectra [n_img][Nx][Nyr] (bins fastest) from taps, on a column slab when ncols > 0 (fft_capi.cu)
int kernel_spectrum_dev(aefft_ctx* ctx, int64_t n_img, int Nk, int Nl, int Nx, int Ny, const float* taps, float* img, float2* spec,
                        int col0 = 0, int ncols = 0);

// ---- collectives inside the engine (comm.cu): NCCL on the ctx stream; op 0 = sum, 1 = average over the ranks
int comm_allreduce(aefft_ctx* ctx, float* dev, int64_t n_floats, int op);
// slab exchange of a bin-sharded transform: rank r sends send + r'*chunk floats to every r' and receives into
// recv + r'*chunk (ncclSend/ncclRecv group = all-to-all over NVSwitch)
int comm_alltoall(aefft_ctx* ctx, const float* send, float* recv, int64_t chunk_floats);
int comm_alltoallv(aefft_ctx* ctx, const float* send, const int64_t* scount, const int64_t* soff, float* recv,
                   const int64_t* rcount, const int64_t* roff);

// ---- host orchestration shared by capi.cu and net.cu ------------------------------------------------
int64_t gbuf_len(int mode, int dD, int dM, int Nk, int Nl);
int coord_gradients_dev(aefft_ctx* ctx, int mode, int quirks, int64_t B, int dD, int dM, int Nx, int Ny, int Nk, int Nl,
                        const float* in, const float* out, const float* hin, const float* f, float* gbuf);
int coord_update_dev(aefft_ctx* ctx, int mode, int64_t B_global, int dD, int dM, int Nx, int Ny, int Nk, int Nl,
                     const float* gbuf, float* c, float* b, float* f, float* p, float* dc, float* db, float* df,
                     float* dp, float* ddc, float* ddb, float* ddf, float* ddp, float delmax, float alpha,
                     float* mse_dev);

}  // namespace aefft
