// cds_feeder — bitstreams -> saliency maps at box scale: ONE process per GPU, fed by N host producer processes.
//
// The reference's host producer is its HM-16.0 decoder (CABAC parse + the capture hooks of TDecSlice.cpp:333-341 and
// TDecCu.cpp:153-289; the driver loop is TAppDecTop::decode, decmain.cpp:81-93).  Here every video is decoded by its own producer
// PROCESS (parse-only: CDS_PARSE_ONLY=1) that writes the hooks' per-picture syntax in the packed wire format of include/cds.h to
// an inherited pipe (HM_PACKED_FD; record = int32 poc, nctu, ntok; int32 hdr[nctu][4]; uint16 bytes[nctu]; int16 tok[ntok]).  This
// process owns the GPU: one clip context and one reader thread per video.  A reader thread reads each record STRAIGHT INTO the
// pinned staging buffers the library hands out (cds_acquire_frame_buffers), submits it (cds_submit_frame: zero copy, asynchronous)
// and collects the maps of the batch before the one just launched (cds_get_map), so the pipe reads, the GPU work of all videos and
// the map copies overlap.  One such process per GPU (tools/feeder_box.py starts them with CUDA_VISIBLE_DEVICES) feeds a whole box;
// videos are independent, so there is no exchange between the processes.
//
//   cds_feeder --producer <decoder> [--fast | --svm-model <libsvm text model>] [--fps F] [--batch B] [--f32] [--full-decode]
//              [--mean a,b,..] [--std a,b,..] [--out-prefix P] <stream.bin>:<W>x<H> ...
//
// Prints one JSON line: pictures, seconds, pictures_per_s, per-video picture counts and an FNV-1a checksum of every video's maps.
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <thread>
#include <vector>

#include "../../include/cds.h"

namespace {

double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

bool read_full(int fd, void* dst, size_t n) {
  char* p = static_cast<char*>(dst);
  while (n > 0) {
    const ssize_t r = read(fd, p, n);
    if (r == 0) return false;
    if (r < 0) { if (errno == EINTR) continue; return false; }
    p += r; n -= (size_t)r;
  }
  return true;
}

struct Options {
  std::string producer, svm_model, out_prefix;
  bool fast = false, u8 = true, parse_only = true;
  double fps = 50.0, mean[9], stdv[9];
  int batch = 32;
};

struct Video {
  std::string path; int w = 0, h = 0;
  pid_t pid = -1; int fd = -1;
  long pictures = 0; uint64_t checksum = 1469598103934665603ull; std::string error;
  double t_first = 0.0, t_last = 0.0;       // first picture submitted / last map drained (the start-up before it is CUDA, the context and the decoder's own)
};

void parse9(const char* s, double* dst) {
  for (int i = 0; i < 9 && s; i++) { char* end; dst[i] = strtod(s, &end); if (end == s) break; s = (*end == ',') ? end + 1 : nullptr; }
}

void run_video(const Options& o, const cds_svm_model* model, Video* v) {
  cds_params p; memset(&p, 0, sizeof(p));
  p.w = v->w; p.h = v->h; p.fps = o.fps; p.max_batch = o.batch; p.out_u8 = o.u8 ? 1 : 0;
  p.use_rbf = o.fast ? CDS_FUSION_FAST : (model ? CDS_FUSION_RBF : CDS_FUSION_LINEAR);
  p.model = model;
  for (int i = 0; i < 9; i++) { p.mean_vec[i] = o.mean[i]; p.std_vec[i] = o.stdv[i]; }
  cds_ctx* ctx = nullptr;
  if (cds_create(&p, &ctx) != CDS_OK) { v->error = std::string("cds_create: ") + cds_last_error(); return; }
  const int nctu = ((v->w + 63) / 64) * ((v->h + 63) / 64);
  const size_t px = (size_t)v->w * v->h * (o.u8 ? 1 : 4);
  std::vector<unsigned char> map(px);
  FILE* out = nullptr;
  if (!o.out_prefix.empty()) {
    const size_t slash = v->path.find_last_of('/');
    std::string stem = v->path.substr(slash == std::string::npos ? 0 : slash + 1);
    out = fopen((o.out_prefix + stem + (o.u8 ? ".u8" : ".f32")).c_str(), "wb");
  }
  long submitted = 0, drained = 0;
  auto drain = [&](long upto) {
    for (; drained < upto && v->error.empty(); drained++) {
      int cam[2];
      if (cds_get_map(ctx, (int)drained, map.data(), cam) != CDS_OK) { v->error = std::string("cds_get_map: ") + cds_last_error(); return; }
      for (size_t i = 0; i < px; i += 64) { v->checksum ^= map[i]; v->checksum *= 1099511628211ull; }     // every 64th byte: cheap, still layout sensitive
      if (out) fwrite(map.data(), 1, px, out);
    }
  };
  while (v->error.empty()) {
    int32_t head[3];
    if (!read_full(v->fd, head, sizeof(head))) break;                                   // end of stream
    const int poc = head[0], rn = head[1], ntok = head[2];
    uint16_t* b; int32_t* h; int16_t* t; size_t cap = 0;
    if (cds_acquire_frame_buffers(ctx, &b, &h, &t, &cap) != CDS_OK) { v->error = std::string("cds_acquire_frame_buffers: ") + cds_last_error(); break; }
    if (rn != nctu || ntok < 0 || (size_t)ntok > cap) { v->error = "syntax record does not match the picture size"; break; }
    if (!read_full(v->fd, h, (size_t)nctu * 16) || !read_full(v->fd, b, (size_t)nctu * 2) || !read_full(v->fd, t, (size_t)ntok * 2)) { v->error = "truncated syntax record"; break; }
    if (cds_submit_frame(ctx, poc, b, h, t, (size_t)ntok) != CDS_OK) { v->error = std::string("cds_submit_frame: ") + cds_last_error(); break; }
    if (submitted == 0) v->t_first = now();
    submitted = poc + 1;
    if (submitted - drained >= 2L * o.batch) drain(drained + o.batch);                  // the batch just launched keeps running
  }
  drain(submitted);                                                                     // cds_get_map launches the partial batch
  v->t_last = now();
  v->pictures = drained;
  if (out) fclose(out);
  cds_destroy(ctx);
}

}  // namespace

int main(int argc, char** argv) {
  Options o;
  for (int i = 0; i < 9; i++) { o.mean[i] = 0.0; o.stdv[i] = 1.0; }
  std::vector<Video> videos;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--producer") o.producer = next();
    else if (a == "--svm-model") o.svm_model = next();
    else if (a == "--fast") o.fast = true;
    else if (a == "--f32") o.u8 = false;
    else if (a == "--full-decode") o.parse_only = false;
    else if (a == "--fps") o.fps = atof(next());
    else if (a == "--batch") o.batch = atoi(next());
    else if (a == "--mean") parse9(next(), o.mean);
    else if (a == "--std") parse9(next(), o.stdv);
    else if (a == "--out-prefix") o.out_prefix = next();
    else {
      const size_t c = a.rfind(':'), x = a.rfind('x');
      if (c == std::string::npos || x == std::string::npos || x < c) { fprintf(stderr, "cds_feeder: bad stream spec %s (want path:WxH)\n", a.c_str()); return 2; }
      Video v; v.path = a.substr(0, c); v.w = atoi(a.substr(c + 1, x - c - 1).c_str()); v.h = atoi(a.substr(x + 1).c_str());
      videos.push_back(v);
    }
  }
  if (o.producer.empty() || videos.empty()) { fprintf(stderr, "usage: cds_feeder --producer <decoder> [--fast | --svm-model m] [--fps F] [--batch B] stream:WxH ...\n"); return 2; }
  const double t0 = now();
  for (Video& v : videos) {                                 // producers first (before this process touches CUDA): they parse while the contexts are being created
    int pfd[2];
    if (pipe(pfd) != 0) { perror("pipe"); return 1; }
    fcntl(pfd[0], F_SETFD, FD_CLOEXEC);
#ifdef F_SETPIPE_SZ
    fcntl(pfd[1], F_SETPIPE_SZ, 1 << 20);
#endif
    v.pid = fork();
    if (v.pid == 0) {
      char fdtxt[16]; snprintf(fdtxt, sizeof(fdtxt), "%d", pfd[1]);
      setenv("HM_PACKED_FD", fdtxt, 1);
      setenv("CDS_PARSE_ONLY", o.parse_only ? "1" : "0", 1);
      unsetenv("CUDA_VISIBLE_DEVICES");                    // the producer never touches the GPU
      const int devnull = open("/dev/null", O_WRONLY);
      if (devnull >= 0) { dup2(devnull, 1); dup2(devnull, 2); }
      execl(o.producer.c_str(), o.producer.c_str(), "-b", v.path.c_str(), (char*)nullptr);
      _exit(127);
    }
    close(pfd[1]);
    v.fd = pfd[0];
  }
  cds_svm_model* model = nullptr;
  if (!cds_device_ok() || (!o.fast && !o.svm_model.empty() && cds_svm_model_load(o.svm_model.c_str(), &model) != CDS_OK)) {
    fprintf(stderr, "cds_feeder: %s\n", cds_last_error());
    for (Video& v : videos) { close(v.fd); kill(v.pid, SIGTERM); waitpid(v.pid, nullptr, 0); }
    return 1;
  }
  std::vector<std::thread> threads;
  for (Video& v : videos) threads.emplace_back(run_video, std::cref(o), model, &v);
  for (auto& t : threads) t.join();
  const double sec = now() - t0;
  long total = 0; bool bad = false;
  for (Video& v : videos) { close(v.fd); int st = 0; waitpid(v.pid, &st, 0); total += v.pictures; bad |= !v.error.empty(); }
  double first = 0.0, last = 0.0;                            // steady state: from the first picture any video submitted to the last map drained
  for (Video& v : videos) if (v.pictures > 0) { if (first == 0.0 || v.t_first < first) first = v.t_first; if (v.t_last > last) last = v.t_last; }
  const double steady = last > first ? last - first : sec;
  printf("{\"pictures\": %ld, \"seconds\": %.4f, \"pictures_per_s\": %.2f, \"startup_seconds\": %.4f, \"steady_seconds\": %.4f, \"steady_pictures_per_s\": %.2f, \"videos\": [",
         total, sec, total / sec, first > t0 ? first - t0 : 0.0, steady, total / steady);
  for (size_t i = 0; i < videos.size(); i++)
    printf("%s{\"stream\": \"%s\", \"pictures\": %ld, \"checksum\": \"%016llx\", \"error\": \"%s\"}", i ? ", " : "", videos[i].path.c_str(), videos[i].pictures,
           (unsigned long long)videos[i].checksum, videos[i].error.c_str());
  printf("]}\n");
  if (model) cds_svm_model_free(model);
  return bad ? 1 : 0;
}
