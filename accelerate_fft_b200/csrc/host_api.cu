// Host-side mirror of the reference's operator interface for this path, above the C ABI.
// The reference's host side is Haskell (FFT.hs, LLVM/PTX.hs, LLVM/PTX/Plans.hs); GHC is not in
// this image, so the same logic is restated in C++ and exported with C linkage for the tests.
// The Haskell shim that a maintainer would actually ship is in haskell/ (see INTEGRATION.md).
//
//   Mode / signOfMode            <- Mode.hs:15-26
//   fft / fft1D / fft2D / fft3D  <- FFT.hs:63-173 (dispatch + Inverse scaling) and PTX.hs:52-70
//   Plans / withPlan             <- PTX/Plans.hs:40-90 (one cache per entry point, keyed by
//                                   (context, shape, type); creation under a lock, exec outside it)
#include <cuda.h>
#include <cuda_runtime.h>

#include <nvtx3/nvToolsExt.h>   // header-only: ranges cost nothing unless a profiler has injected itself

#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

extern char** environ;

#include "../../include/b200fft.h"
#include "generic.h"

namespace {

enum Mode { Forward = 0, Reverse = 1, Inverse = 2 };  // Mode.hs:15-19

// PTX.hs:130-132 fftMode
int fft_direction(int mode) { return mode == Forward ? B200FFT_FORWARD : B200FFT_INVERSE; }

// PTX/Plans.hs:40-44: the reference keys on (context pointer, hash of (shape,type)); we key on the
// full tuple so two shapes can never share a plan through a hash collision.
// The planner's developer switches are environment variables (B200FFT_*) read at plan creation: their fingerprint is part
// of the key, so a plan made under one setting is never handed out under another.
using Key = std::tuple<void*, int, int64_t, int64_t, int64_t, uint64_t>;  // ctx, type, d, h, w, planner-environment fingerprint

struct Plans {
  std::mutex lock;
  std::map<Key, b200fftHandle> plans;
  int (*create)(b200fftHandle*, int64_t, int64_t, int64_t, int);
  const char* name;   // the reference's ForeignAcc name for this cache (PTX.hs:59-70): the NVTX range of every exec
};

uint64_t planner_env_fingerprint() {
  uint64_t h = 1469598103934665603ull;   // FNV-1a over every "B200FFT_*=value" string
  for (char** e = environ; e && *e; e++) {
    if (strncmp(*e, "B200FFT_", 8) != 0) continue;
    for (const char* c = *e; *c; c++) { h ^= (unsigned char)*c; h *= 1099511628211ull; }
    h ^= 0xff; h *= 1099511628211ull;
  }
  return h;
}

int mk1d(b200fftHandle* h, int64_t, int64_t, int64_t n, int t) { return b200fftPlan1d(h, n, t, 1); }                   // PTX.hs:141
int mk2d(b200fftHandle* h, int64_t, int64_t hh, int64_t w, int t) { return b200fftPlan2d(h, hh, w, t); }                // PTX.hs:148
int mk3d(b200fftHandle* h, int64_t d, int64_t hh, int64_t w, int t) { return b200fftPlan3d(h, d, hh, w, t); }           // PTX.hs:155
int mk2many(b200fftHandle* h, int64_t, int64_t hh, int64_t w, int t) { return b200fftPlanMany1d(h, w, hh, t); }         // PTX.hs:162
int mk3many(b200fftHandle* h, int64_t d, int64_t hh, int64_t w, int t) { return b200fftPlanMany1d(h, w, d * hh, t); }   // PTX.hs:169

// PTX.hs:137-170: five global caches
Plans fft1D_plans{{}, {}, mk1d, "cuda.fft1d"}, fft2D_plans{{}, {}, mk2d, "cuda.fft2d"}, fft3D_plans{{}, {}, mk3d, "cuda.fft3d"},
    fft2DMany_plans{{}, {}, mk2many, "cuda.fft2.many"}, fft3DMany_plans{{}, {}, mk3many, "cuda.fft3.many"};
Plans* all_caches[] = {&fft1D_plans, &fft2D_plans, &fft3D_plans, &fft2DMany_plans, &fft3DMany_plans};

// how `run` normalises Mode Inverse; a process-wide developer switch read once per call (atomic: GHC calls from many threads)
std::atomic<int> g_fused_inverse{0};

template <typename F>
F driver_entry(const char* name) {
  // Driver entry points are fetched through the runtime so the library carries no link-time dependency on libcuda.so
  // (it must load, and report NO_DEVICE, on a box without a driver).
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qr) != cudaSuccess) { cudaGetLastError(); fn = nullptr; }
  return (F)fn;
}

// PTX/Plans.hs:78-80: the reference drops a cache entry when its CUDA context dies (a finaliser on the context's Lifetime).
// There is no such hook below the FFI, so every insertion sweeps the cache instead: entries whose context no longer answers
// are destroyed (their device memory went with the context; the handle's host side is what is left to free).
void evict_dead_contexts(Plans& ps, void* current) {
  typedef CUresult (*api_version_t)(CUcontext, unsigned int*);
  static api_version_t api_version = driver_entry<api_version_t>("cuCtxGetApiVersion");
  if (!api_version) return;
  for (auto it = ps.plans.begin(); it != ps.plans.end();) {
    void* ctx = std::get<0>(it->first);
    unsigned int v = 0;
    if (ctx != current && api_version((CUcontext)ctx, &v) != CUDA_SUCCESS) {
      b200fftDestroy(it->second);
      it = ps.plans.erase(it);
    } else {
      ++it;
    }
  }
}

// PTX/Plans.hs:66-86 withPlan: look up / create under the lock, return the handle, run outside it
int with_plan(Plans& ps, int64_t d, int64_t h, int64_t w, int type, b200fftHandle* out) {
  typedef CUresult (*ctx_get_t)(CUcontext*);
  static ctx_get_t ctx_get = driver_entry<ctx_get_t>("cuCtxGetCurrent");
  if (!ctx_get) return B200FFT_NO_DEVICE;
  CUcontext ctx = nullptr;
  if (ctx_get(&ctx) != CUDA_SUCCESS || ctx == nullptr) {
    // runtime-API callers may not have touched the device yet: force the primary context
    if (cudaFree(0) != cudaSuccess) { cudaGetLastError(); return B200FFT_NO_DEVICE; }
    if (ctx_get(&ctx) != CUDA_SUCCESS || ctx == nullptr) return B200FFT_NO_DEVICE;
  }
  const uint64_t env = planner_env_fingerprint();
  std::lock_guard<std::mutex> g(ps.lock);
  Key key{(void*)ctx, type, d, h, w, env};
  auto it = ps.plans.find(key);
  if (it != ps.plans.end()) { *out = it->second; return 0; }
  b200fftHandle hnd = nullptr;
  int e = ps.create(&hnd, d, h, w, type);
  if (e) return e;
  evict_dead_contexts(ps, (void*)ctx);
  ps.plans.emplace(key, hnd);
  *out = hnd;
  return 0;
}

template <typename C, typename T>
__global__ void scale_kernel(C* a, long long n, T s) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    C v = a[i];
    v.x = v.x / s; v.y = v.y / s;   // A.map (/scale), FFT.hs:83
    a[i] = v;
  }
}

// fft' (PTX.hs:77-106) + the Inverse post-scale of FFT.hs
int run(Plans& ps, int mode, int64_t d, int64_t h, int64_t w, int type, double scale, const void* in, void* out,
        b200fftStream stream) {
  if (mode < Forward || mode > Inverse) return B200FFT_INVALID_VALUE;
  if (type != B200FFT_C2C && type != B200FFT_Z2Z) return B200FFT_INVALID_TYPE;
  struct Range { Range(const char* n) { nvtxRangePushA(n); } ~Range() { nvtxRangePop(); } } range(ps.name);
  b200fftHandle hnd = nullptr;
  if (int e = with_plan(ps, d, h, w, type, &hnd)) return e;
  if (mode == Inverse && g_fused_inverse.load(std::memory_order_relaxed))
    return b200fftExecScaled(hnd, in, out, fft_direction(mode), 1.0 / scale, stream);
  if (int e = b200fftExec(hnd, in, out, fft_direction(mode), stream)) return e;
  if (mode == Inverse) {  // case mode of Inverse -> A.map (/scale) (go arr)
    const long long n = (long long)d * h * w;
    long long b = (n + 255) / 256;
    unsigned blocks = (unsigned)(b > 148 * 16 ? 148 * 16 : b);
    if (type == B200FFT_C2C) scale_kernel<float2, float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float2*)out, n, (float)scale);
    else scale_kernel<double2, double><<<blocks, 256, 0, (cudaStream_t)stream>>>((double2*)out, n, scale);
    if (cudaGetLastError() != cudaSuccess) return B200FFT_EXEC_FAILED;
  }
  return 0;
}

// stream sets of accfft_run_host_seq, kept per device between calls
struct StreamSet { int dev; cudaStream_t st[3]; };
std::mutex g_ss_lock;
std::vector<StreamSet> g_ss_free;
bool take_streams(StreamSet* out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return false; }
  {
    std::lock_guard<std::mutex> g(g_ss_lock);
    for (size_t i = 0; i < g_ss_free.size(); i++)
      if (g_ss_free[i].dev == dev) { *out = g_ss_free[i]; g_ss_free.erase(g_ss_free.begin() + i); return true; }
  }
  out->dev = dev;
  for (int s = 0; s < 3; s++)
    if (cudaStreamCreateWithFlags(&out->st[s], cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      for (int q = 0; q < s; q++) cudaStreamDestroy(out->st[q]);
      return false;
    }
  return true;
}
void give_streams(const StreamSet& ss) {
  std::lock_guard<std::mutex> g(g_ss_lock);
  g_ss_free.push_back(ss);
}

}  // namespace

extern "C" {

// FFT.hs:63-84 fft: innermost axis; PTX.hs:52-61 rank dispatch (DIM1 -> fft1D plans, DIM2/DIM3 -> many plans).
// Ranks above 3 are accepted by collapsing the outer extents into the batch; the reference's
// PTX.fft raises internalError there (PTX.hs:61) and falls back to the pure path.
int accfft_fft(int mode, int rank, const int64_t* shape, int type, const void* in, void* out, b200fftStream stream) {
  if (rank < 1 || !shape) return B200FFT_INVALID_VALUE;
  for (int i = 0; i < rank; i++) if (shape[i] < 0) return B200FFT_INVALID_SIZE;
  int64_t w = shape[rank - 1];
  int64_t outer = 1;
  for (int i = 0; i < rank - 1; i++) outer *= shape[i];
  if (w == 0 || outer == 0) return 0;  // empty array: nothing to do
  const double scale = (double)w;      // FFT.hs:69
  if (rank == 1) return run(fft1D_plans, mode, 1, 1, w, type, scale, in, out, stream);
  if (rank == 2) return run(fft2DMany_plans, mode, 1, shape[0], w, type, scale, in, out, stream);
  if (rank == 3) return run(fft3DMany_plans, mode, shape[0], shape[1], w, type, scale, in, out, stream);
  // scale is the innermost length only; the plan key collapses the outer extents
  b200fftHandle hnd = nullptr;
  (void)hnd;
  return run(fft3DMany_plans, mode, 1, outer, w, type, scale, in, out, stream);
}

int accfft_fft1D(int mode, int64_t n, int type, const void* in, void* out, b200fftStream stream) {
  if (n < 0) return B200FFT_INVALID_SIZE;
  if (n == 0) return 0;
  return run(fft1D_plans, mode, 1, 1, n, type, (double)n, in, out, stream);  // FFT.hs:98
}

int accfft_fft2D(int mode, int64_t h, int64_t w, int type, const void* in, void* out, b200fftStream stream) {
  if (h < 0 || w < 0) return B200FFT_INVALID_SIZE;
  if (h == 0 || w == 0) return 0;
  // scale by the whole size, FFT.hs:125; run() takes it separately from the plan key
  b200fftHandle hnd = nullptr;
  (void)hnd;
  return run(fft2D_plans, mode, 1, h, w, type, (double)h * (double)w, in, out, stream);
}

int accfft_fft3D(int mode, int64_t d, int64_t h, int64_t w, int type, const void* in, void* out, b200fftStream stream) {
  if (d < 0 || h < 0 || w < 0) return B200FFT_INVALID_SIZE;
  if (d == 0 || h == 0 || w == 0) return 0;
  return run(fft3D_plans, mode, d, h, w, type, (double)d * (double)h * (double)w, in, out, stream);  // FFT.hs:155
}

// ---- DFT/Centre.hs: the step on either side of the path in image / signal pipelines ----------
static int dims3(int rank, const int64_t* shape, long long* d, long long* h, long long* w) {
  if (rank < 1 || rank > 3 || !shape) return B200FFT_INVALID_VALUE;
  for (int i = 0; i < rank; i++) if (shape[i] < 0) return B200FFT_INVALID_SIZE;
  *w = shape[rank - 1];
  *h = rank >= 2 ? shape[rank - 2] : 1;
  *d = rank >= 3 ? shape[rank - 3] : 1;
  return 0;
}
// centre1D/2D/3D (Centre.hs:36-66): out = (-1)^(sum of indices) * in
int accfft_centre(int rank, const int64_t* shape, int type, const void* in, void* out, b200fftStream stream) {
  long long d, h, w;
  if (int e = dims3(rank, shape, &d, &h, &w)) return e;
  if (type != B200FFT_C2C && type != B200FFT_Z2Z) return B200FFT_INVALID_TYPE;
  if (d * h * w == 0) return 0;
  if (!in || !out) return B200FFT_INVALID_VALUE;
  return b200fft::launch_centre(type == B200FFT_Z2Z, in, out, d, h, w, (cudaStream_t)stream) == cudaSuccess ? 0 : B200FFT_EXEC_FAILED;
}
// shift1D/2D/3D (Centre.hs:70-82,98-114,134-153: roll by n/2 + odd n) and ishift* (Centre.hs:84-96,116-132,155-164: n/2)
int accfft_shift(int rank, const int64_t* shape, int type, int inverse, const void* in, void* out, b200fftStream stream) {
  long long d, h, w;
  if (int e = dims3(rank, shape, &d, &h, &w)) return e;
  if (type != B200FFT_C2C && type != B200FFT_Z2Z) return B200FFT_INVALID_TYPE;
  if (d * h * w == 0) return 0;
  if (!in || !out || in == out) return B200FFT_INVALID_VALUE;
  auto amount = [&](long long n) { return n / 2 + ((inverse || !(n & 1)) ? 0 : 1); };
  return b200fft::launch_shift(type == B200FFT_Z2Z, in, out, d, h, w, amount(d), amount(h), amount(w), (cudaStream_t)stream) == cudaSuccess
             ? 0 : B200FFT_EXEC_FAILED;
}
// shiftND (fftND mode x) in one go: kind 1/2/3 = fft1D/fft2D/fft3D.  Power-of-two extents: the rotation rides on the
// stores of each axis' last butterfly pass (b200fftExecShifted: no extra pass over the array); otherwise the transform
// goes to a pool buffer and the stand-alone shift follows.  Even extents: equals fftND mode (centreND x).
int accfft_fft_centred(int kind, int mode, const int64_t* shape, int type, const void* in, void* out, b200fftStream stream) {
  if (kind < 1 || kind > 3 || !shape) return B200FFT_INVALID_VALUE;
  if (mode < Forward || mode > Inverse) return B200FFT_INVALID_VALUE;
  if (type != B200FFT_C2C && type != B200FFT_Z2Z) return B200FFT_INVALID_TYPE;
  long long d, h, w;
  if (int e = dims3(kind, shape, &d, &h, &w)) return e;
  if (d * h * w == 0) return 0;
  Plans& ps = kind == 1 ? fft1D_plans : kind == 2 ? fft2D_plans : fft3D_plans;
  b200fftHandle hnd = nullptr;
  if (int e = with_plan(ps, d, h, w, type, &hnd)) return e;
  const double scale = mode == Inverse ? 1.0 / ((double)d * (double)h * (double)w) : 1.0;
  int e = b200fftExecShifted(hnd, in, out, fft_direction(mode), scale, stream);
  if (e != B200FFT_NOT_SUPPORTED) return e;
  void* tmp = nullptr;
  const size_t bytes = (size_t)(d * h * w) * (type == B200FFT_Z2Z ? 16 : 8);
  if (b200fft::pool_alloc(&tmp, bytes, (cudaStream_t)stream) != cudaSuccess) { cudaGetLastError(); return B200FFT_ALLOC_FAILED; }
  e = b200fftExecScaled(hnd, in, tmp, fft_direction(mode), scale, stream);
  if (!e) e = accfft_shift(kind, shape, type, 0, tmp, out, stream);
  cudaFreeAsync(tmp, (cudaStream_t)stream);
  return e;
}

// One whole-array transform (or a chain of them) of a device-resident array on `stream`; src is preserved,
// the result of the last transform lands in *result (either bufA or bufB, ping-pong).
static int run_chain(int kind, const int* modes, int nmodes, int rank, const int64_t* shape, int type, const void* src, void* bufA,
                     void* bufB, b200fftStream stream, void** result) {
  const void* cur = src;
  void* dst = bufB;
  for (int m = 0; m < nmodes; m++) {
    int e;
    switch (kind) {
      case 0: e = accfft_fft(modes[m], rank, shape, type, cur, dst, stream); break;
      case 1: e = accfft_fft1D(modes[m], shape[0], type, cur, dst, stream); break;
      case 2: e = accfft_fft2D(modes[m], shape[0], shape[1], type, cur, dst, stream); break;
      case 3: e = accfft_fft3D(modes[m], shape[0], shape[1], shape[2], type, cur, dst, stream); break;
      default: e = B200FFT_INVALID_VALUE;
    }
    if (e) return e;
    cur = dst;
    dst = (dst == bufB) ? bufA : bufB;
  }
  *result = const_cast<void*>(cur);
  return 0;
}

// Host-buffer flavour: what Accelerate's `run` does around the foreign call (copy the `use`d array in, run the
// Aforeign nodes, copy the result out), for a chain of `nmodes` transforms applied one after the other
// (e.g. {Forward, Inverse}) with the array staying on the device in between.
//
// B200-first: for kind 0 (`fft`, innermost axis -- the rows are independent units) the array is cut into chunks of
// whole rows that flow through three streams, so the H2D copy of chunk c+1, the kernels of chunk c and the D2H copy
// of chunk c-1 overlap (PCIe is full duplex and the copy engines are separate from the SMs); the step then costs
// max(H2D, D2H) instead of H2D + kernels + D2H.  Device staging comes from the library's scratch pool.  Whole-array
// transforms (fft1D/2D/3D) cannot be cut and run copy -> transform -> copy.  Synchronous; pass pinned host buffers
// for asynchronous DMA (pageable memory works but serialises inside the driver).
int accfft_run_host_seq(int kind, const int* modes, int nmodes, int rank, const int64_t* shape, int type, const void* h_in,
                        void* h_out) {
  if (rank < 1 || rank > 8 || !shape || !modes || nmodes < 1 || !h_in || !h_out) return B200FFT_INVALID_VALUE;
  if ((kind == 1 && rank != 1) || (kind == 2 && rank != 2) || (kind == 3 && rank != 3) || kind < 0 || kind > 3) return B200FFT_INVALID_VALUE;
  if (type != B200FFT_C2C && type != B200FFT_Z2Z) return B200FFT_INVALID_TYPE;
  for (int m = 0; m < nmodes; m++) if (modes[m] < Forward || modes[m] > Inverse) return B200FFT_INVALID_VALUE;
  long long n = 1;
  for (int i = 0; i < rank; i++) { if (shape[i] < 0) return B200FFT_INVALID_SIZE; n *= shape[i]; }
  if (n == 0) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return B200FFT_NO_DEVICE; }
  const size_t esz = type == B200FFT_Z2Z ? 16 : 8;
  const size_t bytes = (size_t)n * esz;
  const int64_t w = shape[rank - 1];
  const int64_t outer = n / w;

  // chunking: ~64 MiB chunks, at least 4 of them, whole rows each
  int64_t rows_per_chunk = outer;
  constexpr int NSLOT = 3;
  if (kind == 0 && outer >= 2 * NSLOT && bytes >= ((size_t)32 << 20)) {
    int64_t nchunks = (int64_t)(bytes >> 26);
    if (nchunks < 2 * NSLOT) nchunks = 2 * NSLOT;
    if (nchunks > outer) nchunks = outer;
    rows_per_chunk = (outer + nchunks - 1) / nchunks;
  }
  const int64_t nchunks = (outer + rows_per_chunk - 1) / rows_per_chunk;
  const int nslot = nchunks > 1 ? NSLOT : 1;
  const size_t chunk_bytes = (size_t)rows_per_chunk * w * esz;

  // the three copy / compute streams are kept between calls (a call takes a set from the free list and returns it)
  StreamSet ss;
  if (!take_streams(&ss)) return B200FFT_ALLOC_FAILED;
  cudaStream_t* st = ss.st;
  void *a[NSLOT] = {nullptr, nullptr, nullptr}, *b[NSLOT] = {nullptr, nullptr, nullptr};
  int e = 0;
  for (int s = 0; s < nslot && !e; s++) {
    if (b200fft::pool_alloc(&a[s], chunk_bytes, st[s]) != cudaSuccess || b200fft::pool_alloc(&b[s], chunk_bytes, st[s]) != cudaSuccess)
      e = B200FFT_ALLOC_FAILED;
  }
  for (int64_t c = 0; c < nchunks && !e; c++) {
    const int s = (int)(c % nslot);
    const int64_t r0 = c * rows_per_chunk;
    const int64_t rows = (outer - r0 < rows_per_chunk) ? outer - r0 : rows_per_chunk;
    const size_t off = (size_t)r0 * w * esz, cb = (size_t)rows * w * esz;
    if (cudaMemcpyAsync(a[s], (const char*)h_in + off, cb, cudaMemcpyHostToDevice, st[s]) != cudaSuccess) { e = B200FFT_EXEC_FAILED; break; }
    void* res = nullptr;
    if (nchunks == 1) {
      e = run_chain(kind, modes, nmodes, rank, shape, type, a[s], a[s], b[s], st[s], &res);
    } else {
      const int64_t sh2[2] = {rows, w};   // a chunk of rows is a DIM2 array for `fft` (same innermost length, same scale)
      e = run_chain(0, modes, nmodes, 2, sh2, type, a[s], a[s], b[s], st[s], &res);
    }
    if (e) break;
    if (cudaMemcpyAsync((char*)h_out + off, res, cb, cudaMemcpyDeviceToHost, st[s]) != cudaSuccess) e = B200FFT_EXEC_FAILED;
  }
  for (int s = 0; s < nslot; s++) {
    if (a[s]) cudaFreeAsync(a[s], st[s]);
    if (b[s]) cudaFreeAsync(b[s], st[s]);
    if (cudaStreamSynchronize(st[s]) != cudaSuccess && !e) e = B200FFT_EXEC_FAILED;
  }
  give_streams(ss);
  if (e) cudaGetLastError();
  return e;
}

int accfft_run_host(int kind, int mode, int rank, const int64_t* shape, int type, const void* h_in, void* h_out) {
  return accfft_run_host_seq(kind, &mode, 1, rank, shape, type, h_in, h_out);
}

void accfft_set_fused_inverse(int on) { g_fused_inverse = on != 0; }

int accfft_plan_cache_size(void) {
  int n = 0;
  for (Plans* p : all_caches) { std::lock_guard<std::mutex> g(p->lock); n += (int)p->plans.size(); }
  return n;
}

void accfft_plan_cache_clear(void) {
  for (Plans* p : all_caches) {
    std::lock_guard<std::mutex> g(p->lock);
    for (auto& kv : p->plans) b200fftDestroy(kv.second);
    p->plans.clear();
  }
}

}  // extern "C"
