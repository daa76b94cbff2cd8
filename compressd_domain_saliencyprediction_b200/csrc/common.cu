// common.cu — error state, device probing, scratch buffers.
#include "cds_common.cuh"
#include <stdarg.h>
#include <atomic>
#include <map>
#include <mutex>

namespace cds {
static thread_local char g_err[512] = "";
static std::atomic<long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int ensure_device() {
  static int state = 0;   // 0 unknown, 1 ok, -1 absent
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (state == 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      set_error("no CUDA device: %s (libcds_b200 has no CPU fallback)", cudaGetErrorString(e));
      cudaGetLastError();
      state = -1;
    } else {
      int dev = 0; cudaGetDevice(&dev);
      cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
      if (pr.major != 10) {
        set_error("device %s is sm_%d%d; libcds_b200 is built for sm_100a only", pr.name, pr.major, pr.minor);
        state = -1;
      } else state = 1;
    }
  }
  if (state != 1) {
    if (!g_err[0]) set_error("no usable sm_100 device (libcds_b200 has no CPU fallback)");
    return CDS_ENODEVICE;
  }
  return CDS_OK;
}

int ensure_dyn_smem_impl(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  int dev = 0;
  CDS_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  auto it = done.find({func, dev});
  if (it == done.end()) {
    // Ask for the largest shared-memory carve-out whenever one of these kernels runs.  The L1 / shared split of an SM can only change
    // while the SM is empty: a persistent kernel that configured "just enough" shared memory for itself would keep every CTA of a
    // co-scheduled kernel (pipeline.cu: SVM beside contrast) off its SMs until it exits.
    CDS_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    it = done.emplace(std::make_pair(func, dev), (size_t)0).first;
  }
  size_t& have = it->second;
  if (bytes > have) {
    if (bytes > 48 * 1024) CDS_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
  return CDS_OK;
}

int device_sm_count() {
  static std::mutex mu;
  static std::map<int, int> sms;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::lock_guard<std::mutex> lk(mu);
  int& n = sms[dev];
  if (!n && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  return n;
}

int DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return CDS_OK;
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu): %s", want, cudaGetErrorString(e)); return CDS_ENOMEM; }
  cap = want;
  return CDS_OK;
}
DevBuf::~DevBuf() { /* process teardown: leave to the driver */ }
}  // namespace cds

extern "C" {
const char* cds_last_error(void) { return cds::g_err; }
int cds_version(void) { return 100; }
int cds_device_ok(void) { return cds::ensure_device() == CDS_OK ? 1 : 0; }
long cds_take_launch_count(void) { return cds::g_launches.exchange(0); }
}
