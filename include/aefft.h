/* aefft.h -- C ABI of the B200-native autoencoder engine (libaefft.so).
 *
 * Drop-in boundary for ONE hot path of fabrii4/AutoEncoder-FFT: the convolutional autoencoder's
 * forward + backprop training step in coordinate space and in momentum (FFT) space.
 * The reference has no FFI layer; its boundary is the set of C++ free functions in
 *   source/netlib.h:4-24, source/backproplib.h:5-16, source/fft_backproplib.h:5-11
 * called from source/autoencoder.cpp (sites: :132 :140-148 :169-200 :42 :100-107 :361-374).
 * Every entry point below names the reference function it replaces.  The C++ mirror with the exact
 * reference signatures lives in autoencoder-fft_b200/shim/ and calls only this ABI.
 *
 * Conventions
 *   - plain pointers + sizes, no C++/torch types; all functions return 0 on success, !=0 on error
 *     (message via aefft_last_error()).  The reference returns void and checks nothing.
 *   - float32 data.  Feature maps are [B][ch][Nx][Ny] (j fastest; B=1 is the reference's only case),
 *     kernels c[dM][dD][Nk][Nl], f[dD][dM][Nk][Nl], flat in nesting order.
 *   - `loc` says where the DATA pointers of a call live: AEFFT_HOST (copies are made inside the call,
 *     like the reference's per-call H2D/D2H) or AEFFT_DEVICE (device pointers, no copies, async on the
 *     ctx stream).  Scalar outputs (mse) are always host pointers (may be NULL).
 *   - B>1 is this engine's batch extension: raw gradients are averaged over the B frames and ONE
 *     clipped update is applied; B=1 reproduces the reference exactly (DESIGN.md "batch semantics").
 *   - There is NO CPU fallback: every compute entry point fails with AEFFT_ERR_CUDA when no GPU is usable.
 */
#ifndef AEFFT_H
#define AEFFT_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AEFFT_ABI_VERSION 1

enum { AEFFT_OK = 0, AEFFT_ERR_ARG = 1, AEFFT_ERR_CUDA = 2, AEFFT_ERR_IO = 3, AEFFT_ERR_UNSUPPORTED = 4 };
enum { AEFFT_HOST = 0, AEFFT_DEVICE = 1 };

/* Tap-offset / bounds convention of a coordinate-space conv. */
enum {
  AEFFT_CONV_CUDA = 0, /* backproplib.cu:123,89: ak=((Nk-1)/2-1)/2, >=0 bounds, input pre-scaled by 1/dM */
  AEFFT_CONV_CPU = 1   /* netlib.cpp:325,338:    ak=(Nk-1)/2-1, strict >0 bounds, no pre-scale            */
};

/* Training back-end being replaced (autoencoder.cpp:171-201). */
enum {
  AEFFT_MODE_CPU_REF = 0,      /* netlib.cpp:361 backprop: sequential-f, no momentum                     */
  AEFFT_MODE_CUDA_REF = 1,     /* backproplib.cu:291 backprop_gpu                                        */
  AEFFT_MODE_CUDA_REF_SYM = 2  /* backproplib.cu:521 backprop_gpu_cc (tied weights)                      */
};

/* Bug-compat switches for AEFFT_MODE_CUDA_REF.  Default for parity runs: all on. */
enum {
  AEFFT_QUIRK_C1 = 1, /* backproplib.cu:220  bias gradient keeps only d1=dD-1                            */
  AEFFT_QUIRK_C3 = 2, /* backproplib.cu:283  (j-ik) instead of (j-il) in the dF term, taps != (0,0)      */
  AEFFT_QUIRK_C4 = 4, /* backproplib.cu:225,282,335  dDdF buffer keeps stale border values               */
  AEFFT_QUIRKS_ALL = 7
};

/* Arithmetic of the coordinate-space contractions (conv forward / data gradient / weight gradients).
 *   FP32   : fp32 FMA on CUDA cores.
 *   BF16X3 : tcgen05 tensor cores, every fp32 operand split into bf16 hi+lo, three products accumulated in fp32
 *            (fp32-grade: ~1e-5 relative L2 against the fp32 oracle; meets the 1e-4 parity bar).
 *   BF16   : tcgen05 tensor cores, single bf16 product (looser tolerance: ~5e-3 relative L2). */
enum { AEFFT_PRECISION_FP32 = 0, AEFFT_PRECISION_BF16X3 = 1, AEFFT_PRECISION_BF16 = 2 };

typedef struct aefft_ctx aefft_ctx; /* one per process/GPU: device, stream, workspace arena, FFT plans */

const char* aefft_last_error(void);
int aefft_abi_version(void);
/* Number of CUDA kernels this library has launched on this ctx since creation (bench.py "gpu_launches"). */
int64_t aefft_launch_count(const aefft_ctx* ctx);

int aefft_create(aefft_ctx** ctx, int device);
int aefft_destroy(aefft_ctx* ctx);
int aefft_sync(aefft_ctx* ctx);
/* The CUDA stream all work of this ctx is launched on (as an opaque cudaStream_t). */
void* aefft_stream(aefft_ctx* ctx);

int aefft_set_precision(aefft_ctx* ctx, int precision);
int aefft_get_precision(const aefft_ctx* ctx);
/* Launch all work of this ctx on a caller-owned stream (e.g. torch's current stream, so torch.distributed
 * collectives and torch.cuda.Event timing order with it); NULL restores the ctx's own stream. */
int aefft_set_stream(aefft_ctx* ctx, void* cuda_stream);
/* Per-kernel timing: while enabled, every kernel launch is bracketed by a CUDA event pair on the launching stream.
 * aefft_profile_read synchronises, aggregates by kernel name (names: max_rows x 64 chars; ms/counts/flops/bytes summed
 * over launches, flops/bytes being the ALGORITHMIC work of those launches) and clears the records. */
int aefft_profile_enable(aefft_ctx* ctx, int on);
int aefft_profile_read(aefft_ctx* ctx, int max_rows, char* names, float* ms, int64_t* counts, double* flops,
                       double* bytes, int* n_rows);

/* Device memory helpers for callers that keep data resident (loc == AEFFT_DEVICE) without linking cudart. */
int aefft_malloc(aefft_ctx* ctx, void** dev_ptr, int64_t bytes);
int aefft_free(aefft_ctx* ctx, void* dev_ptr);
/* kind: 0 host->device, 1 device->host, 2 device->device; ordered on the ctx stream, synchronous on return. */
int aefft_memcpy(aefft_ctx* ctx, void* dst, const void* src, int64_t bytes, int kind);

/* ------------------------------------------------------------------ multi-GPU (one process / ctx per GPU)
 * The reference has no multi-GPU path.  aefft_comm_init gives a ctx an NCCL communicator (NVLink /
 * NVSwitch); afterwards the engine itself issues the collectives on the ctx stream:
 *   - aefft_net_step: the raw gradient blocks of ALL pairs are all-reduced (sum) as ONE fused buffer, then every rank
 *     applies the identical clipped-momentum update with B_global = B * world (weights stay replicated);
 *   - aefft_backprop_fft (data-parallel frames): one all-reduce(avg) of the kernel-space block [dck|dfk|db|dp] per
 *     iteration, the mse trace is averaged once at the end of the call;
 *   - aefft_backprop_fft under aefft_set_bin_shard: all-reduce(sum) of the partial block per iteration, the partial
 *     mse trace is summed once at the end.
 * The reduction always precedes the non-linear clip g/max(10,|g|).  Bootstrap: rank 0 calls aefft_comm_unique_id and
 * hands the AEFFT_COMM_ID_BYTES bytes to the other ranks by any means (file, socket, MPI, torch.distributed). */
#define AEFFT_COMM_ID_BYTES 128
int aefft_comm_unique_id(void* id_bytes);
int aefft_comm_init(aefft_ctx* ctx, const void* id_bytes, int rank, int world);
int aefft_comm_destroy(aefft_ctx* ctx);
int aefft_comm_rank(const aefft_ctx* ctx);
int aefft_comm_world(const aefft_ctx* ctx);
/* all-reduce of a device buffer on the ctx stream; op 0 = sum, 1 = average.  No-op for world == 1. */
int aefft_comm_allreduce(aefft_ctx* ctx, float* dev, int64_t n_floats, int op);

/* ------------------------------------------------------------------ forward, coordinate space */

/* Conv_gpu (backproplib.cu:114-182) / Conv (netlib.cpp:318-358), per `convention`.
 * in [B][dD][Nx][Ny], c [dM][dD][Nk][Nl], b [dM] -> out [B][dM][Nx][Ny]. */
int aefft_conv_fwd(aefft_ctx* ctx, int loc, int convention, int64_t B, int dD, int dM, int Nx, int Ny, int Nk,
                   int Nl, const float* in, const float* c, const float* b, float* out);

/* Pool (netlib.cpp:114-164).  scale>0: max through an int accumulator floored at 0; scale<0: replicate.
 * in [B][D][Nx][Ny] -> out [B][D][oNx][oNy] (caller-sized, as in autoencoder.cpp:70-74). */
int aefft_pool(aefft_ctx* ctx, int loc, int64_t B, int D, int Nx, int Ny, int oNx, int oNy, int scale,
               const float* in, float* out);

/* Portion (netlib.cpp:292-315): centre crop by q of one [B][D][Nx][Ny] tensor -> [B][D][Nx/q][Ny/q]. */
int aefft_portion(aefft_ctx* ctx, int loc, int64_t B, int D, int Nx, int Ny, int q, const float* in, float* out);

/* ------------------------------------------------------------------ training, coordinate space */

/* One layer pair, one step: backprop (netlib.cpp:361), backprop_gpu (backproplib.cu:291) or
 * backprop_gpu_cc (:521) per `mode`.  in,out [B][dD][Nx][Ny], hin [B][dM][Nx][Ny].
 * c,b,f,p and the caller-owned momentum (dc,db,df,dp) / last-gradient (ddc,ddb,ddf,ddp) buffers are
 * updated in place (the d* / dd* pointers are ignored for CPU_REF and may be NULL).
 * *mse receives the value the reference prints (CUDA: sum e^2/Norm; CPU: raw sum e^2), averaged over B. */
int aefft_backprop_coord(aefft_ctx* ctx, int loc, int mode, int quirks, int64_t B, int dD, int dM, int Nx,
                         int Ny, int Nk, int Nl, const float* in, const float* out, const float* hin, float* c,
                         float* b, float* f, float* p, float* dc, float* db, float* df, float* dp, float* ddc,
                         float* ddb, float* ddf, float* ddp, float delmax, float alpha, int active, float* mse);

/* Data-parallel split of the same step (device pointers only):
 *   aefft_coord_gradients  -> raw, un-clipped, SUMMED-over-local-frames gradient block in gbuf (device)
 *   [all-reduce(sum) of gbuf across ranks -- torch.distributed / NCCL, outside this library]
 *   aefft_coord_update     -> divide by B_global, clip, momentum, update weights (replicated on every rank)
 * gbuf length = aefft_coord_gbuf_len(mode,dD,dM,Nk,Nl) floats. */
int64_t aefft_coord_gbuf_len(int mode, int dD, int dM, int Nk, int Nl);
int aefft_coord_gradients(aefft_ctx* ctx, int mode, int quirks, int64_t B, int dD, int dM, int Nx, int Ny, int Nk,
                          int Nl, const float* in, const float* out, const float* hin, const float* c,
                          const float* f, float* gbuf);
int aefft_coord_update(aefft_ctx* ctx, int mode, int64_t B_global, int dD, int dM, int Nx, int Ny, int Nk, int Nl,
                       const float* gbuf, float* c, float* b, float* f, float* p, float* dc, float* db, float* df,
                       float* dp, float* ddc, float* ddb, float* ddf, float* ddp, float delmax, float alpha,
                       float* mse_dev);

/* ------------------------------------------------------------------ momentum (FFT) space */

/* Batched 2-D real-to-complex / complex-to-real transforms, unnormalised both ways, n={Nx,Ny}; replaces the
 * cufftPlanMany+cufftExecR2C/C2R call sites fft_backproplib.cu:779/796, 821/829, 885/910, 937/946, 1208-1282.
 * spec is interleaved (re,im) [batch][Nx][Ny/2+1][2].  Nx, Ny: even numbers of the form 2^a 3^b 5^c up to 8192 (the same
 * holds for every momentum-space entry point below); powers of two take the fast compile-time kernels, other lengths (the
 * camera's 640 x 480 and its pooled levels) the run-time mixed-radix kernels. */
int aefft_fft_r2c(aefft_ctx* ctx, int loc, int64_t batch, int Nx, int Ny, const float* in, float* spec);
int aefft_fft_c2r(aefft_ctx* ctx, int loc, int64_t batch, int Nx, int Ny, const float* spec, float* out);

/* kernel_pad (fft_backproplib.cu:1018-1064): c [dM][dD][Nk][Nl] -> c_pad [dM][dD][Nx][Ny] (wrapped). */
int aefft_kernel_pad(aefft_ctx* ctx, int loc, int dM, int dD, int Nk, int Nl, int Nx, int Ny, const float* c,
                     float* c_pad);

/* Kernel spectra in the reference's net_cfreq wire format (store_cfreq, fft_backproplib.cu:1117-1127):
 * interleaved (re,im) float32 [dM][dD][Nx][Ny/2+1][2] = R2C(kernel_pad(c)). */
int aefft_kernel_spectrum(aefft_ctx* ctx, int loc, int dM, int dD, int Nk, int Nl, int Nx, int Ny, const float* c,
                          float* cfreq);

/* autoenc_fft (fft_backproplib.cu:1331-1376): full-stack forward in frequency space.
 *   n_conv convs; conv n has dims[4n..4n+3]=(dM,dD,Nk,Nl), weights c_all+coff[n], bias b_all+boff[n],
 *   spectral pooling scale[n].  layer l has ldims[3l..3l+2]=(D,Nx,Ny) at layers_all+loff[l] (per frame; frame
 *   stride lstride floats), n_layers = 2*n_conv+1; layer 0 is the input.  fft_l!=0 writes every layer, else only
 *   the last.  cfreq_all (+cfoff[n]) is the net_cfreq cache: cfreq_valid!=0 -> used as the kernel spectra;
 *   ==0 -> rebuilt from c_all and written back (StoreLoad_cfreq :1146-1161).  May be NULL (always rebuild). */
int aefft_autoenc_fft(aefft_ctx* ctx, int loc, int64_t B, int n_conv, const int* dims, const float* c_all,
                      const int64_t* coff, const float* b_all, const int64_t* boff, const int* scale, int n_layers,
                      const int* ldims, float* layers_all, const int64_t* loff, int64_t lstride, int cfreq_valid,
                      float* cfreq_all, const int64_t* cfoff, int fft_l);

/* Data-parallel training in momentum space: a host callback invoked by aefft_backprop_fft once per iteration, between
 * the kernel-space gradients and the clipped-momentum update, on the contiguous RAW gradient block
 * [dck dM*dD*Nk*Nl | dfk dD*dM*Nk*Nl | db dM | dp dD] (device memory, already averaged over this rank's B frames).
 * The callback must average the block over the ranks IN PLACE with work ordered on the ctx stream (e.g. an NCCL
 * all-reduce on that stream) and return 0.  The reference has no multi-GPU path; with no hook (NULL) the call is the
 * reference's single-device algorithm.  The reduction precedes the non-linear clip g/max(10,|g|). */
typedef int (*aefft_gradient_hook_fn)(void* user, float* dev_block, int64_t n_floats);
int aefft_set_gradient_hook(aefft_ctx* ctx, aefft_gradient_hook_fn fn, void* user);
/* Frequency-bin sharding of aefft_backprop_fft over `world` devices (BASELINE config 4): every device receives ALL frames
 * (device pointers), transforms them once per call and keeps only its slab of spectrum columns
 * [rank*Nyr/world, (rank+1)*Nyr/world); the per-bin contractions, the pruned kernel DFTs and the mse then run on the
 * slab.  The gradient hook (required) is called on the PARTIAL gradient block and on each mse value and must ADD them
 * over the devices (NCCL all-reduce(sum) on the ctx stream); kernels / biases end identical everywhere.  The one-off
 * frame transform is amortised over the n_iter (reference: 100) iterations of a call.  world == 1 restores the default. */
int aefft_set_bin_shard(aefft_ctx* ctx, int rank, int world);
/* On the resident net (aefft_net_fft_step / aefft_net_fft_train_pair) bin sharding needs the ctx's communicator
 * (aefft_comm_init with the same rank / world) and must be set BEFORE the first momentum-space call on the net.  The net
 * then holds this rank's share of the frames (global batch = B * world): the forward runs data parallel on them, each
 * trained pair's in / out spectra are cut into column slabs and exchanged with one all-to-all per spectrum (the transpose
 * of a slab-decomposed transform, over NVLink / NVSwitch), training runs on this rank's slab of ALL frames, and the partial
 * kernel-space gradient blocks are all-reduced (sum) -- nothing is replicated except the (tiny) kernels. */

/* backprop_fft (fft_backproplib.cu:1381-1511): n_iter (reference: 100) iterations of spectral gradients ->
 * kernel-space clipped-momentum update (lr 0.1*del0, alpha 0.9, momentum zeroed per call) -> re-forward.
 * in, expout, out [B][dD][Nx][Ny]; cfreq/ffreq wire-format spectra (in: cache, out: trained; may be NULL ->
 * derived from c,f); c,f,b,p updated in place.  mse_trace (host, n_iter+1 floats, may be NULL) receives the
 * "mse fft:" / "n: .. mse:" values the reference prints (:1441,:1464). */
int aefft_backprop_fft(aefft_ctx* ctx, int loc, int64_t B, int dD, int dM, int Nx, int Ny, int Nk, int Nl,
                       const float* in, const float* expout, const float* out, float* cfreq, float* c, float* ffreq,
                       float* f, float* b, float* p, float del0, int maxdiff, int n_iter, float* mse_trace);

/* Diagnostic (tests only, like aefft_profile_*): ONE batched-over-bins real GEMM of the tensor-core momentum path
 * (csrc/spec_tc.cu, tcgen05 kind::tf32 with the 3xTF32 split): D_w[M][N] = A_w . B_w^T for S bins on device buffers
 * [S][rows][cols]; *_mn = 0: rows index M / N and cols index K, 1: rows index K.  outer != 0 applies the complex-pair
 * epilogue of the frame-reduced outer products (out [S][M/2][N/2][2]).  It is the per-bin "complex batched GEMM" that
 * replaces conv_k (fft_backproplib.cu:162-189) and the contractions of gradient_k_io (:395-475). */
int aefft_spec_bin_gemm(aefft_ctx* ctx, int64_t S, const float* a, int a_rows, int a_cols, int a_mn, const float* b, int b_rows,
                        int b_cols, int b_mn, int M, int N, int K, int outer, int conj_out, float scale, float* out);

/* backprop_fft on layers stored per frame with stride `frame_stride` floats (the layer block aefft_autoenc_fft writes);
 * device pointers, expout = in (autoencoder.cpp:194), spectra derived from c,f. */
int aefft_backprop_fft_strided(aefft_ctx* ctx, int64_t B, int dD, int dM, int Nx, int Ny, int Nk, int Nl, const float* in,
                               const float* out, int64_t frame_stride, float* c, float* f, float* b, float* p, float del0,
                               int maxdiff, int n_iter, float* mse_trace);
/* Strided 2-D copy on the ctx stream (asynchronous): `height` rows of `width` bytes; kind as in aefft_memcpy. */
int aefft_memcpy2d(aefft_ctx* ctx, void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width,
                   int64_t height, int kind);

/* ------------------------------------------------------------------ glue (netlib.cpp:167-289) */

/* Init_conv (netlib.cpp:167-197): consumes libc rand() in the reference's order; call srand() first. */
int aefft_init_conv(float* c, float* b, int mS, int dD, int kS, int lS, float rmax);
/* SaveLoad_conv (netlib.cpp:220-272): byte-exact ./weights/C_weights_{L}{_in,_out}_D=.._M=.._Lk=.._Ll=.._S=...conv
 * under `dir` (reference: "./weights").  Unlike the reference a missing file is an error (AEFFT_ERR_IO). */
int aefft_saveload_conv(const char* dir, float* c, float* b, int dM, int dD, int Nk, int Nl, int scale, int L,
                        int io, int write);
/* LoadParam (netlib.cpp:274-289): positional name/value pairs dM, Lk, Ll, scale, rmax from `path`
 * (reference: "New_Layer_Param.txt"). */
int aefft_load_param(const char* path, int* dM, int* Lk, int* Ll, int* scal, float* rmax);

/* Synthetic frames: pixel = float(splitmix64(seed, linear index of (b0+b,d,i,j)) & 255). */
int aefft_synth_frames(aefft_ctx* ctx, int loc, uint64_t seed, int64_t b0, int64_t B, int D, int Nx, int Ny,
                       float* out);

/* ------------------------------------------------------------------ device-resident network
 * Replays autoencoder.cpp's state model (:69-120, :384-457) with all state in HBM: layers[], net_c[], net_b[],
 * scale[], momentum, cached spectra.  Pair n: encoder conv n, decoder conv N-1-n (N = 2*pairs). */
typedef struct aefft_net aefft_net;
int aefft_net_create(aefft_ctx* ctx, aefft_net** net, int D, int Nx, int Ny, int64_t B);
int aefft_net_destroy(aefft_net* net);
/* 'n' key (autoencoder.cpp:384-431): insert a new innermost pair; weights drawn with Init_conv from libc rand(). */
int aefft_net_add_layer(aefft_net* net, int dM, int Lk, int Ll, int scal, float rmax);
/* 'd' key (:432-457): remove the innermost pair (never the last one). */
int aefft_net_delete_layer(aefft_net* net);
int aefft_net_num_pairs(const aefft_net* net);
/* copy weights of conv n (0..2*pairs-1) to/from host: c [dM][dD][Nk][Nl], b [dM] */
int aefft_net_get_conv(aefft_net* net, int n, float* c, float* b);
int aefft_net_set_conv(aefft_net* net, int n, const float* c, const float* b);
int aefft_net_conv_dims(const aefft_net* net, int n, int* dM, int* dD, int* Nk, int* Nl, int* scale);
/* 'p' key (:343-356): copy encoder weights of pair n_l into its decoder (f[d][m]=c[m][d]). */
int aefft_net_set_symmetric(aefft_net* net, int n_l);
/* 'z'/'x' keys (:281-292): zero the momentum / last-gradient buffers of pair n_l (the reference shares one set
 * between all pairs and re-initialises it whenever the active pair changes). */
int aefft_net_reset_momentum(aefft_net* net, int n_l);
int aefft_net_num_layers(const aefft_net* net);
/* layer l (0..2*convs) dims and device pointer ([B][D][Nx][Ny]) */
int aefft_net_layer(aefft_net* net, int l, int* D, int* Nx, int* Ny, float** dev_ptr);
/* forward, coordinate space (autoencoder.cpp:135-150): frames = layer 0 ([B][D][Nx][Ny], per `loc`).
 * Asynchrony contract of aefft_net_forward / aefft_net_step / aefft_net_set_frames_u8: with loc == AEFFT_HOST the call
 * returns after the caller's buffer has been copied (it may be refilled at once); with loc == AEFFT_DEVICE everything is
 * queued on the ctx stream and the device buffer must stay untouched until that work has run (aefft_sync, or order your
 * own work on aefft_stream()).  The compute itself is asynchronous in both cases unless an mse pointer is passed. */
int aefft_net_forward(aefft_net* net, int loc, const float* frames);
/* ImageToSpin_C (netlib.cpp:37-50) on the device: B interleaved 8-bit images [B][rows = Ny][cols = Nx][D] (a cv::Mat's
 * data for D = 3: B,G,R bytes) become layer 0, spin[d][i][j] = (float)img(row j, col i)[d] -- raw 0..255, not
 * normalised.  Uploading bytes moves 4x less data over PCIe than float frames; follow with aefft_net_forward /
 * aefft_net_step with frames == NULL (layer 0 already set). */
int aefft_net_set_frames_u8(aefft_net* net, int loc, const unsigned char* images);
/* SpinToImage_C (netlib.cpp:52-76; mode 0: clamp(round(v), 0, 255)) / SpinToImage_V (:79-92; mode 1: (uchar)(int)v) of a
 * real-space layer with <= 4 channels on the device: B interleaved 8-bit images [B][rows = Ny][cols = Nx][D] -- the display
 * side of the reference's webcam loop, and together with aefft_net_set_frames_u8 the raw-video front end of aefft_replay
 * (--video / --dump).  The layer must be current (coordinate forward, or a momentum-space forward that materialises it). */
int aefft_net_get_layer_u8(aefft_net* net, int layer, int mode, int loc, unsigned char* images);
/* train pair n_l on the activations of the last forward (autoencoder.cpp:158-201, q=1).  *mse host or NULL.
 * With a communicator (world > 1) the pair's raw gradient block is all-reduced before the update (B_global = B * world). */
int aefft_net_train_pair(aefft_net* net, int n_l, int mode, int quirks, float delmax, float alpha, float* mse);
/* raw gradient block of pair n_l into the net's gradient buffer (device ptr returned), then update: the
 * data-parallel split (all-reduce the buffer in between). */
int aefft_net_pair_gradients(aefft_net* net, int n_l, int mode, int quirks, float** gbuf_dev, int64_t* gbuf_len);
int aefft_net_pair_update(aefft_net* net, int n_l, int mode, int64_t B_global, float delmax, float alpha, float* mse);
/* ---- momentum (FFT) space on the resident net (autoencoder.cpp:131-133 forward with fft == 1, :190-196 training).
 * The forward keeps the layer spectra in HBM; real-space layers (aefft_net_layer) are written per fft_l:
 * 1 = every layer (the reference's display mode, fft_backproplib.cu:1347-1361; every spectrum is produced), 0 = the last
 * layer only (:1373), -1 = none.  With fft_l <= 0 only the spectra the training reads are produced (every pair's in / out
 * layers; a conv followed by a spectral pooling computes the kept bins only, the decoder runs on the bins that came up
 * from the innermost level), so real-space layers other than the last are NOT current after such a forward.  aefft_net_fft_step = that forward + n_iter iterations of backprop_fft (:1381-1511; the reference runs 100,
 * lr 0.1*del0, alpha 0.9, momentum zeroed per call) for EVERY pair, consuming the pair's in / out spectra directly -- the
 * reference inverse-transforms the layers and backprop_fft transforms them again, a round trip that is the identity up to
 * fp32 rounding.  mse: host [pairs][n_iter+1] (the "mse fft:" / "n: .. mse:" values) or NULL (then nothing synchronises).
 * With a communicator (aefft_comm_init) the kernel-space gradient block of every iteration is averaged over the ranks. */
int aefft_net_fft_forward(aefft_net* net, int loc, const float* frames, int fft_l);
int aefft_net_fft_step(aefft_net* net, int loc, const float* frames, float del0, int maxdiff, int n_iter, int fft_l,
                       float* mse);
/* backprop_fft of ONE pair (the active pair n_l of autoencoder.cpp:190-196) on the spectra of the last
 * aefft_net_fft_forward / aefft_net_fft_step; mse_trace: host, n_iter+1 floats, or NULL. */
int aefft_net_fft_train_pair(aefft_net* net, int n_l, float del0, int maxdiff, int n_iter, float* mse_trace);
/* net_cfreq[n] (fft_backproplib.cu:1146-1161) as a lazily computed VIEW of the device-resident kernels of conv n:
 * the interleaved wire-format spectrum [dM][dD][Nx][Ny/2+1][2] at the resolution conv n runs at, n_floats = its length.
 * The reference keeps this as a host cache that is uploaded (51-136 MB per layer) on every frame. */
int aefft_net_get_cfreq(aefft_net* net, int n, float* cfreq, int64_t n_floats);

/* Momentum sidecar (the reference's 's'/'l' keys drop this state, autoencoder.cpp:358-383): the inertia and
 * last-gradient buffers of pair n_l to / from <dir>/C_momentum_{n_l}_D=.._M=.._Lk=.._Ll=...mom, raw float32
 * [dc | db | df | dp | ddc | ddb | ddf | ddp].  write != 0 saves, 0 loads (AEFFT_ERR_IO on a missing / mismatching file). */
int aefft_net_saveload_momentum(aefft_net* net, const char* dir, int n_l, int write);

/* Offsets (floats) of every pair's raw gradient block inside the net's fused gradient buffer for `mode`, and its total
 * length: the buffer a data-parallel step all-reduces once (offsets may be NULL). */
int aefft_net_fused_layout(aefft_net* net, int mode, int64_t* offsets, int64_t* total);
/* one whole training step: forward + train every pair once (order 0..pairs-1). mse[pairs] host or NULL.
 * With a communicator (aefft_comm_init, world > 1): gradients of all pairs -> ONE all-reduce -> updates. */
int aefft_net_step(aefft_net* net, int loc, const float* frames, int mode, int quirks, float delmax, float alpha,
                   float* mse);

#ifdef __cplusplus
}
#endif
#endif /* AEFFT_H */
