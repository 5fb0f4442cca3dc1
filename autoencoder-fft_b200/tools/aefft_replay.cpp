// Headless replay of the reference application's state machine (autoencoder.cpp:98-120 start-up, :279-457 key handlers)
// over the C ABI of libaefft: the "caller" side of the drop-in boundary without the webcam / OpenCV window.
//
//   aefft_replay [--frames B] [--size NxxNy] [--channels D] [--seed S] [--param FILE] [--weights DIR]
//                [--precision fp32|bf16x3|bf16] [--del 0.2] [--alpha 0.9] [--quirks 7] [--fft-iters 100]
//                [--device K --rank R --world W --id-file PATH] [--sidecar]
//                --script "n n t5 z t3 p t2 s d l t1 i f g t1"
// --video FILE: training frames come from a raw 8-bit video instead of the synthetic generator -- consecutive interleaved images
//   [frame][rows = Ny][cols = Nx][D] (what `ffmpeg -i in.mp4 -f rawvideo -pix_fmt bgr24 FILE` writes; the reference's webcam
//   loop feeds cv::Mat frames of exactly this layout through ImageToSpin_C), converted on the device; the file wraps around.
// --dump FILE: after the script, one more forward of the last batch and its reconstruction written as raw 8-bit frames of the
//   same layout (SpinToImage_C: clamp(round(v), 0, 255)).
// --rank/--world/--id-file: data-parallel frames over W processes, one GPU each, WITHOUT any Python: rank 0 writes the
// NCCL unique id (aefft_comm_unique_id) to PATH, the others read it, every rank calls aefft_comm_init; rank r then
// trains on frames [it*W*B + r*B, +B) and the engine all-reduces the raw gradient block before every update, so all
// ranks print the same mse lines as ONE process with --frames W*B.
// --quirks: bit mask of the reference's backprop_gpu defects to reproduce (aefft.h AEFFT_QUIRK_*).  Default: all of them
// on square frames (= what the reference computes), none on non-square frames, where the reference indexes out of bounds
// and the library only offers the intended gradients.
//
// Script tokens (the reference's keys, plus `tK` = K training frames-batches on the active pair):
//   n  add the innermost pair from the parameter file (LoadParam + Init_conv, :384-431)      d  delete it (:432-457)
//   z / x  next / previous active pair; the shared momentum / last-gradient buffers restart (:279-310)
//   e  re-draw the weights of the active pair (:311-325)      p  toggle symmetric weights, copying c into f (:331-356)
//   s / l  save / load the active pair's weight files (:357-382)      i  print the structure (:458-)
//   f  toggle momentum (FFT) space (:274)      g  toggle fft_l, the inverse transform of every layer (:275)
//   m  toggle the multiobjective kernel-diversity term (:278)
//   tK  K iterations of: synthetic frames (SURVEY 8d generator, frame counter advancing) -> forward -> backprop of the
//       active pair (backprop_gpu, or backprop_gpu_cc when symmetric), one "mse" line each like the reference prints.
//       In FFT mode: autoenc_fft forward (:131-132) -> backprop_fft of the active pair (:190-196, --fft-iters iterations,
//       the reference hard-codes 100), printing its "mse fft:" / "n: .. mse:" lines (fft_backproplib.cu:1441, :1464).
// Every event prints one line; the run ends with "replay ok" and exit code 0, or the library's error text and 1.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

#include "aefft.h"

#define CHECK(call)                                                                      \
  do {                                                                                   \
    if ((call) != AEFFT_OK) {                                                            \
      std::fprintf(stderr, "aefft_replay: %s failed: %s\n", #call, aefft_last_error()); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

int main(int argc, char** argv) {
  int64_t B = 1;
  int D = 3, Nx = 640, Ny = 480, precision = AEFFT_PRECISION_BF16X3, quirks = -1;
  unsigned seed = 1234;
  float del = 0.2f, alpha = 0.9f;
  std::string param = "New_Layer_Param.txt", weights = "./weights", script, id_file, video, dump;
  int device = 0, rank = 0, world = 1, fft_iters = 100, sidecar = 0;
  for (int a = 1; a < argc; a++) {
    const std::string k = argv[a];
    const char* v = a + 1 < argc ? argv[a + 1] : "";
    if (k == "--frames") { B = std::atoll(v); a++; }
    else if (k == "--size") { if (std::sscanf(v, "%dx%d", &Nx, &Ny) != 2) { std::fprintf(stderr, "bad --size\n"); return 1; } a++; }
    else if (k == "--channels") { D = std::atoi(v); a++; }
    else if (k == "--seed") { seed = (unsigned)std::strtoul(v, nullptr, 10); a++; }
    else if (k == "--param") { param = v; a++; }
    else if (k == "--weights") { weights = v; a++; }
    else if (k == "--del") { del = (float)std::atof(v); a++; }
    else if (k == "--alpha") { alpha = (float)std::atof(v); a++; }
    else if (k == "--script") { script = v; a++; }
    else if (k == "--quirks") { quirks = std::atoi(v); a++; }
    else if (k == "--device") { device = std::atoi(v); a++; }
    else if (k == "--rank") { rank = std::atoi(v); a++; }
    else if (k == "--world") { world = std::atoi(v); a++; }
    else if (k == "--id-file") { id_file = v; a++; }
    else if (k == "--video") { video = v; a++; }
    else if (k == "--dump") { dump = v; a++; }
    else if (k == "--fft-iters") { fft_iters = std::atoi(v); a++; }
    else if (k == "--sidecar") { sidecar = 1; }
    else if (k == "--precision") {
      precision = !std::strcmp(v, "fp32") ? AEFFT_PRECISION_FP32 : !std::strcmp(v, "bf16") ? AEFFT_PRECISION_BF16 : AEFFT_PRECISION_BF16X3;
      a++;
    } else { std::fprintf(stderr, "unknown option %s\n", k.c_str()); return 1; }
  }
  if (quirks < 0) quirks = Nx == Ny ? AEFFT_QUIRKS_ALL : 0;
  aefft_ctx* ctx = nullptr;
  aefft_net* net = nullptr;
  CHECK(aefft_create(&ctx, device));
  CHECK(aefft_set_precision(ctx, precision));
  if (world > 1) {
    // bootstrap of the engine's NCCL communicator through a file (any channel works: the id is 128 opaque bytes)
    if (id_file.empty()) { std::fprintf(stderr, "aefft_replay: --world needs --id-file\n"); return 1; }
    unsigned char id[AEFFT_COMM_ID_BYTES];
    if (rank == 0) {
      CHECK(aefft_comm_unique_id(id));
      const std::string tmp = id_file + ".tmp";
      FILE* fh = std::fopen(tmp.c_str(), "wb");
      if (!fh || std::fwrite(id, 1, sizeof(id), fh) != sizeof(id)) { std::fprintf(stderr, "aefft_replay: cannot write %s\n", tmp.c_str()); return 1; }
      std::fclose(fh);
      std::rename(tmp.c_str(), id_file.c_str());
    } else {
      FILE* fh = nullptr;
      for (int tries = 0; tries < 600 && !(fh = std::fopen(id_file.c_str(), "rb")); tries++) usleep(100000);
      if (!fh || std::fread(id, 1, sizeof(id), fh) != sizeof(id)) { std::fprintf(stderr, "aefft_replay: cannot read %s\n", id_file.c_str()); return 1; }
      std::fclose(fh);
    }
    CHECK(aefft_comm_init(ctx, id, rank, world));
  }
  CHECK(aefft_net_create(ctx, &net, D, Nx, Ny, B));
  FILE* vfh = nullptr;
  const size_t fbytes = (size_t)Nx * Ny * D;
  int64_t vframes = 0;
  std::vector<unsigned char> vbuf;
  if (!video.empty()) {
    vfh = std::fopen(video.c_str(), "rb");
    if (!vfh) { std::fprintf(stderr, "aefft_replay: cannot open %s\n", video.c_str()); return 1; }
    std::fseek(vfh, 0, SEEK_END);
    vframes = (int64_t)(std::ftell(vfh) / (long)fbytes);
    if (vframes < 1 || D > 4) { std::fprintf(stderr, "aefft_replay: %s holds no %dx%dx%d frame\n", video.c_str(), Nx, Ny, D); return 1; }
    vbuf.resize((size_t)B * fbytes);
  }
  std::srand(seed);  // the reference seeds once and draws every Init_conv from the same stream (autoencoder.cpp:100)
  int n_l = 0, sym = 0, fft = 0, fft_l = 0, maxdiff = 0;
  int64_t frame0 = 0;
  // start-up: the reference builds its first pair from the parameter file before the loop (:98-120)
  auto add_layer = [&]() -> int {
    int dM = 10, Lk = 0, Ll = 0, scal = 2;
    float rmax = 3.f;
    if (aefft_load_param(param.c_str(), &dM, &Lk, &Ll, &scal, &rmax) != AEFFT_OK) return 1;
    if (aefft_net_add_layer(net, dM, Lk, Ll, scal, rmax) != AEFFT_OK) return 1;
    n_l = aefft_net_num_pairs(net) - 1;  // the new pair becomes the active one (:426)
    std::printf("Added new layer L %d\n", aefft_net_num_pairs(net));
    return 0;
  };
  std::vector<std::string> tok;
  for (size_t i = 0; i < script.size();) {
    while (i < script.size() && script[i] == ' ') i++;
    size_t j = i;
    while (j < script.size() && script[j] != ' ') j++;
    if (j > i) tok.push_back(script.substr(i, j - i));
    i = j;
  }
  for (const std::string& t : tok) {
    const int pairs = aefft_net_num_pairs(net);
    const int N = 2 * pairs - 1;
    if (t == "n") {
      if (add_layer()) { std::fprintf(stderr, "aefft_replay: add layer: %s\n", aefft_last_error()); return 1; }
    } else if (t == "d") {
      if (pairs > 1) {
        CHECK(aefft_net_delete_layer(net));
        n_l = 0;
        CHECK(aefft_net_reset_momentum(net, n_l));
        std::printf("Deleted last layer\n");
      }
    } else if (t == "z" || t == "x") {
      if (pairs < 1) continue;
      n_l = t == "z" ? (n_l + 1) % pairs : (n_l - 1 + pairs) % pairs;  // (the reference's (n_l-1)%size goes negative: UB)
      CHECK(aefft_net_reset_momentum(net, n_l));
      std::printf("Active layer %d\n", n_l);
    } else if (t == "e") {
      int dM, dD, Nk, Nl, sc;
      CHECK(aefft_net_conv_dims(net, n_l, &dM, &dD, &Nk, &Nl, &sc));
      int a1, a2, a3, a4;
      float rmax = 3.f;
      CHECK(aefft_load_param(param.c_str(), &a1, &a2, &a3, &a4, &rmax));
      std::vector<float> c((size_t)dM * dD * Nk * Nl), b(dM), f(c.size()), p(dD);
      CHECK(aefft_init_conv(c.data(), b.data(), dM, dD, Nk, Nl, rmax));
      CHECK(aefft_init_conv(f.data(), p.data(), dD, dM, Nk, Nl, rmax));
      CHECK(aefft_net_set_conv(net, n_l, c.data(), b.data()));
      CHECK(aefft_net_set_conv(net, N - n_l, f.data(), p.data()));
      std::printf("Initialize random convolutional weights\n");
    } else if (t == "f") {
      fft = (fft + 1) % 2;
      std::printf("fft %d\n", fft);
    } else if (t == "g") {
      fft_l = (fft_l + 1) % 2;
      std::printf("fft_l %d\n", fft_l);
    } else if (t == "m") {
      maxdiff = (maxdiff + 1) % 2;
      std::printf("multiobjective %d\n", maxdiff);
    } else if (t == "p") {
      sym = (sym + 1) % 2;
      std::printf("Symmetric weights %d\n", sym);
      if (sym) CHECK(aefft_net_set_symmetric(net, n_l));
    } else if (t == "s" || t == "l") {
      const int write = t == "s";
      for (int io = 0; io < 2; io++) {
        const int n = io ? N - n_l : n_l;
        int dM, dD, Nk, Nl, sc;
        CHECK(aefft_net_conv_dims(net, n, &dM, &dD, &Nk, &Nl, &sc));
        std::vector<float> c((size_t)dM * dD * Nk * Nl), b(dM);
        if (write) CHECK(aefft_net_get_conv(net, n, c.data(), b.data()));
        CHECK(aefft_saveload_conv(weights.c_str(), c.data(), b.data(), dM, dD, Nk, Nl, sc, n_l, io, write));
        if (!write) CHECK(aefft_net_set_conv(net, n, c.data(), b.data()));
      }
      // --sidecar: also keep the momentum / last-gradient state the reference drops on save
      if (sidecar) CHECK(aefft_net_saveload_momentum(net, weights.c_str(), n_l, write));
      std::printf(write ? "Saved convolutional weights\n" : "Loaded convolutional weights\n");
    } else if (t == "i") {
      std::printf("Network structure: %d pair(s), %d layers, active %d, symmetric %d\n", pairs, aefft_net_num_layers(net), n_l, sym);
      for (int n = 0; n < 2 * pairs; n++) {
        int dM, dD, Nk, Nl, sc;
        CHECK(aefft_net_conv_dims(net, n, &dM, &dD, &Nk, &Nl, &sc));
        std::printf("  conv %d: %d -> %d, %dx%d taps, scale %d\n", n, dD, dM, Nk, Nl, sc);
      }
    } else if (t.size() > 1 && t[0] == 't') {
      if (pairs < 1) { std::fprintf(stderr, "aefft_replay: no layer to train (use n first)\n"); return 1; }
      const int K = std::atoi(t.c_str() + 1);
      int D0, X0, Y0;
      float* layer0 = nullptr;
      CHECK(aefft_net_layer(net, 0, &D0, &X0, &Y0, &layer0));
      std::vector<float> trace((size_t)fft_iters + 1);
      for (int it = 0; it < K; it++) {
        if (vfh) {
          // this rank's B frames of the batch, wrapping around at the end of the file
          for (int64_t b = 0; b < B; b++) {
            const int64_t fidx = (frame0 + (int64_t)rank * B + b) % vframes;
            if (std::fseek(vfh, (long)(fidx * (int64_t)fbytes), SEEK_SET) != 0 ||
                std::fread(vbuf.data() + (size_t)b * fbytes, 1, fbytes, vfh) != fbytes) {
              std::fprintf(stderr, "aefft_replay: cannot read frame %lld of %s\n", (long long)fidx, video.c_str());
              return 1;
            }
          }
          CHECK(aefft_net_set_frames_u8(net, AEFFT_HOST, vbuf.data()));
        } else {
          CHECK(aefft_synth_frames(ctx, AEFFT_DEVICE, 1234, frame0 + (int64_t)rank * B, B, D0, X0, Y0, layer0));
        }
        frame0 += B * world;
        if (fft) {
          CHECK(aefft_net_fft_forward(net, AEFFT_DEVICE, nullptr, fft_l));
          CHECK(aefft_net_fft_train_pair(net, n_l, del, maxdiff, fft_iters, trace.data()));
          std::printf("mse fft: %.6g\n", (double)trace[0]);
          for (int n = 0; n < fft_iters; n++) std::printf("n: %d mse: %.6g\n", n, (double)trace[n + 1]);
          continue;
        }
        CHECK(aefft_net_forward(net, AEFFT_DEVICE, nullptr));
        float mse = 0.f;
        CHECK(aefft_net_train_pair(net, n_l, sym ? AEFFT_MODE_CUDA_REF_SYM : AEFFT_MODE_CUDA_REF, quirks, del, alpha, &mse));
        std::printf("mse %.9g\n", (double)mse);
      }
    } else {
      std::fprintf(stderr, "aefft_replay: unknown script token '%s'\n", t.c_str());
      return 1;
    }
  }
  if (!dump.empty()) {
    if (aefft_net_num_pairs(net) < 1) { std::fprintf(stderr, "aefft_replay: --dump needs a layer\n"); return 1; }
    const int nl = aefft_net_num_layers(net);
    if (fft) CHECK(aefft_net_fft_forward(net, AEFFT_DEVICE, nullptr, 0));
    else CHECK(aefft_net_forward(net, AEFFT_DEVICE, nullptr));
    std::vector<unsigned char> img((size_t)B * fbytes);
    CHECK(aefft_net_get_layer_u8(net, nl - 1, 0, AEFFT_HOST, img.data()));
    FILE* fh = std::fopen(dump.c_str(), "wb");
    if (!fh || std::fwrite(img.data(), 1, img.size(), fh) != img.size()) { std::fprintf(stderr, "aefft_replay: cannot write %s\n", dump.c_str()); return 1; }
    std::fclose(fh);
    std::printf("reconstruction of %lld frames -> %s\n", (long long)B, dump.c_str());
  }
  if (vfh) std::fclose(vfh);
  CHECK(aefft_sync(ctx));
  CHECK(aefft_net_destroy(net));
  CHECK(aefft_destroy(ctx));
  std::printf("replay ok\n");
  return 0;
}
