// Slab-decomposed 3D transform across the GPUs of one NVLink box, behind the C ABI (include/b200fft.h, "multi-GPU entry
// points").  New capability: the reference is single-device -- what it has per device is the plan cache
// of /root/reference/src/Data/Array/Accelerate/Math/FFT/LLVM/PTX/Plans.hs:68-73 and the plan3D call of PTX.hs:155 -- so parity
// is defined against the same single-array fft3D (FFT.hs:150-173).
//
// One process per GPU; rank g owns z-planes [g*D/P, (g+1)*D/P) of a dense (D, H, W) array.
//   x pass   rows of the local planes                                         (HBM-bound, local)
//   y pass   columns of the local planes; every ky row is stored straight into the memory of the rank that owns it
//            (b200fftExecScatter over CUDA-IPC peer mappings: the all-to-all IS the pass's store, no pack, no collective
//            library, no unpack) -- NVLink-bound, so it runs as a grid-stride loop on a FRACTION of the SMs
//   z pass   columns along z of the received [H/P][D][W] block                 (HBM-bound, local); for the natural layout
//            it scatters back the same way into the z-slabs.
// Pipelining: the y pass goes column chunk by column chunk (and, inside a chunk, plane chunk by plane chunk).  x of plane
// chunk p+1 runs beside y of plane chunk p during the first column chunk; after a chunk's barrier its z pass runs beside the
// y passes of the next chunk.  The NVLink time is the floor; x and all but the last z chunk hide under it.
// Ranks meet in barriers that are a few flag words in peer memory (one tiny kernel: release-store my epoch into every
// peer's slot, acquire-spin on my own slots) -- stream-ordered, no host synchronisation, no NCCL on the data path.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/b200fft.h"

namespace {

constexpr unsigned kSlabMagic = 0x51AB3Du;
constexpr int kMaxRanks = 16;

struct BarrierArgs {
  unsigned* peer[kMaxRanks];   // peer[r] = rank r's flag block (16 slots of 32 bits), mapped here
  unsigned* mine;              // this rank's flag block
  int rank, n;
  unsigned epoch;
};

// Every kernel launched before this one on the stream has completed (its stores, peer stores included, are performed);
// the flag is written with system scope after a system fence, and a rank leaves only when all ranks have arrived.
__global__ void slab_barrier_kernel(BarrierArgs a) {
  const int r = threadIdx.x;
  if (r < a.n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer[r] + a.rank), "r"(a.epoch) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.mine + r) : "memory");
      if ((int)(v - a.epoch) < 0) __nanosleep(200);
    } while ((int)(v - a.epoch) < 0);
  }
  __syncthreads();
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && atoi(e) > 0 ? atoi(e) : dflt;
}

}  // namespace

struct b200fft_slab_s {
  unsigned magic = kSlabMagic;
  int type = 0, esz = 8, rank = 0, P = 1, natural = 0;
  int64_t d = 0, h = 0, w = 0, dl = 0, hl = 0;
  struct Pipe {                               // one pipelining configuration (and its plans) per output layout
    int cp = 1, ck = 1, y_ctas = 0;           // plane chunks, column chunks, CTAs of the scatter pass (0 = one per tile)
    int64_t dlc = 0, wc = 0;
    b200fftHandle px = nullptr, py = nullptr, pz = nullptr;
  } pipe[2];
  char *tmp = nullptr, *recv = nullptr, *back = nullptr;
  unsigned* flags = nullptr;
  std::vector<void*> recv_p, back_p, flag_p, opened;
  unsigned epoch = 0;
  cudaStream_t sy = nullptr, sz = nullptr, sb = nullptr;
  cudaEvent_t e_start = nullptr, e_b0 = nullptr, e_sy = nullptr, e_sz = nullptr, e_sb = nullptr;
  std::vector<cudaEvent_t> e_x, e_y, e_b;
};

namespace {

void free_plans(b200fft_slab_s::Pipe* s) {
  if (s->px) b200fftDestroy(s->px);
  if (s->py) b200fftDestroy(s->py);
  if (s->pz) b200fftDestroy(s->pz);
  s->px = s->py = s->pz = nullptr;
}

// (re)build the three local plans and the events for a chunking
int build_plans(b200fft_slab_s* s, int layout, int cp, int ck, int y_ctas) {
  b200fft_slab_s::Pipe* q = &s->pipe[layout];
  if (cp < 1) cp = 1;
  if (ck < 1) ck = 1;
  while (cp > 1 && s->dl % cp) cp--;
  while (ck > 1 && (s->w % ck || (s->w / ck) % 16)) ck--;     // whole 128-byte runs per chunk
  free_plans(q);
  q->cp = cp; q->ck = ck; q->dlc = s->dl / cp; q->wc = s->w / ck; q->y_ctas = y_ctas < 0 ? 0 : y_ctas;
  int e = b200fftPlanMany1d(&q->px, s->w, q->dlc * s->h, s->type);
  if (!e) e = b200fftPlanAxisView(&q->py, q->dlc, s->h, q->wc, s->h * s->w, s->w, s->type);
  // recv is [hl][D][W]: the z axis has row stride W (short-stride columns: the pipelined column kernel serves them)
  if (!e) e = b200fftPlanAxisView(&q->pz, s->hl, s->d, q->wc, s->d * s->w, s->w, s->type);
  if (e) { free_plans(q); return e; }
  auto grow = [](std::vector<cudaEvent_t>& v, size_t n) {
    while (v.size() < n) {
      cudaEvent_t ev;
      if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return false;
      v.push_back(ev);
    }
    return true;
  };
  if (!grow(s->e_x, (size_t)cp) || !grow(s->e_y, (size_t)ck) || !grow(s->e_b, (size_t)ck)) return B200FFT_ALLOC_FAILED;
  return B200FFT_SUCCESS;
}

int launch_barrier(b200fft_slab_s* s, cudaStream_t st) {
  BarrierArgs a{};
  for (int r = 0; r < s->P; r++) a.peer[r] = (unsigned*)s->flag_p[r];
  a.mine = s->flags;
  a.rank = s->rank; a.n = s->P;
  a.epoch = ++s->epoch;
  slab_barrier_kernel<<<1, 32, 0, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? B200FFT_SUCCESS : B200FFT_EXEC_FAILED;
}

}  // namespace

extern "C" {

int b200fftPlanSlab3d(b200fftSlabHandle* plan, int64_t d, int64_t h, int64_t w, int type, int rank, int nranks, int flags,
                      b200fftAllgatherFn allgather, void* ctx) {
  if (!plan || !allgather) return B200FFT_INVALID_VALUE;
  if (type != B200FFT_C2C && type != B200FFT_Z2Z) return B200FFT_INVALID_TYPE;
  if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) return B200FFT_INVALID_VALUE;
  if (d < 1 || h < 1 || w < 1 || d % nranks || h % nranks) return B200FFT_INVALID_SIZE;
  auto pow2 = [](int64_t v) { return v > 0 && (v & (v - 1)) == 0; };
  // the two scatter passes are single strided-axis passes (b200fftExecScatter): power-of-two H and D up to 2048
  if (!pow2(h) || !pow2(d) || h > 2048 || d > 2048 || h < 2 || d < 2) return B200FFT_NOT_SUPPORTED;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return B200FFT_NO_DEVICE; }
  b200fft_slab_s* s = new b200fft_slab_s;
  s->type = type; s->esz = type == B200FFT_C2C ? 8 : 16;
  s->rank = rank; s->P = nranks; s->natural = (flags & B200FFT_SLAB_NATURAL) ? 1 : 0;
  s->d = d; s->h = h; s->w = w; s->dl = d / nranks; s->hl = h / nranks;
  auto fail = [&](int e) { b200fftDestroySlab(s); return e; };
  const size_t slab_bytes = (size_t)s->dl * h * w * s->esz;     // == D * hl * W
  // Buffers whose IPC handles go to the peers are whole multiples of 2 MiB: the driver carves smaller cudaMalloc requests out of
  // shared 2 MiB blocks, and an IPC handle names the block, not the piece -- a peer would map the block's base.
  auto ipc_size = [](size_t b) { const size_t g = (size_t)2 << 20; return (b + g - 1) / g * g; };
  if (cudaMalloc(&s->tmp, slab_bytes) != cudaSuccess || cudaMalloc(&s->recv, ipc_size(slab_bytes)) != cudaSuccess ||
      (s->natural && cudaMalloc(&s->back, ipc_size(slab_bytes)) != cudaSuccess) || cudaMalloc(&s->flags, ipc_size(256)) != cudaSuccess) {
    cudaGetLastError();
    return fail(B200FFT_ALLOC_FAILED);
  }
  cudaMemset(s->flags, 0, 256);
  cudaDeviceSynchronize();
  // bootstrap: every rank learns the IPC handles of every other rank's receive buffers and flag block
  unsigned char blob[192];
  memset(blob, 0, sizeof blob);
  int e = b200fftPeerExport(s->recv, blob);
  if (!e && s->natural) e = b200fftPeerExport(s->back, blob + 64);
  if (!e) e = b200fftPeerExport(s->flags, blob + 128);
  if (e) return fail(e);
  std::vector<unsigned char> all((size_t)nranks * sizeof blob);
  if (allgather(ctx, blob, all.data(), sizeof blob) != 0) return fail(B200FFT_EXEC_FAILED);
  s->recv_p.assign(nranks, nullptr); s->back_p.assign(nranks, nullptr); s->flag_p.assign(nranks, nullptr);
  for (int r = 0; r < nranks; r++) {
    if (r == rank) { s->recv_p[r] = s->recv; s->back_p[r] = s->back; s->flag_p[r] = s->flags; continue; }
    const unsigned char* hb = all.data() + (size_t)r * sizeof blob;
    void* q = nullptr;
    if ((e = b200fftPeerOpen(hb, &q))) return fail(e);
    s->recv_p[r] = q; s->opened.push_back(q);
    if (s->natural) {
      if ((e = b200fftPeerOpen(hb + 64, &q))) return fail(e);
      s->back_p[r] = q; s->opened.push_back(q);
    }
    if ((e = b200fftPeerOpen(hb + 128, &q))) return fail(e);
    s->flag_p[r] = q; s->opened.push_back(q);
  }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (cudaStreamCreateWithPriority(&s->sy, cudaStreamNonBlocking, hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&s->sb, cudaStreamNonBlocking, hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&s->sz, cudaStreamNonBlocking, lo) != cudaSuccess)
    return fail(B200FFT_INTERNAL_ERROR);
  for (cudaEvent_t* ev : {&s->e_start, &s->e_b0, &s->e_sy, &s->e_sz, &s->e_sb})
    if (cudaEventCreateWithFlags(ev, cudaEventDisableTiming) != cudaSuccess) return fail(B200FFT_INTERNAL_ERROR);
  // defaults measured on 2 and 8 x B200, 1024^3 c64 (profiles/r02_slab_cabi.txt); b200fftSlabTune changes them: 2 plane chunks x
  // 2 column chunks with the scatter pass on 112 CTAs -- x of the second plane chunk and z of the first column chunk run beside
  // the NVLink-bound y passes on the remaining SMs (8 GPUs: 2.10 -> 1.93 ms transposed-out, 3.11 -> 3.00 ms natural).
  const bool big = nranks >= 2 && w >= 512 && s->dl >= 2;
  if ((e = build_plans(s, B200FFT_SLAB_TRANSPOSED_OUT, big ? 2 : 1, big ? 2 : 1, big ? 112 : 0))) return fail(e);
  if (s->natural && (e = build_plans(s, B200FFT_SLAB_NATURAL_OUT, big ? 2 : 1, big ? 2 : 1, big ? 112 : 0))) return fail(e);
  // nobody returns before every rank has mapped everything (a second exchange doubles as the barrier)
  if (allgather(ctx, blob, all.data(), sizeof blob) != 0) return fail(B200FFT_EXEC_FAILED);
  *plan = s;
  return B200FFT_SUCCESS;
}

int b200fftSlabTune(b200fftSlabHandle s, int layout, int plane_chunks, int col_chunks, int y_ctas) {
  if (!s || s->magic != kSlabMagic) return B200FFT_INVALID_PLAN;
  if (layout != B200FFT_SLAB_TRANSPOSED_OUT && layout != B200FFT_SLAB_NATURAL_OUT) return B200FFT_INVALID_VALUE;
  if (layout == B200FFT_SLAB_NATURAL_OUT && !s->natural) return B200FFT_NOT_SUPPORTED;
  cudaDeviceSynchronize();
  return build_plans(s, layout, plane_chunks, col_chunks, y_ctas);
}

int b200fftSlabNaturalBuffer(b200fftSlabHandle s, void** ptr) {
  if (!s || s->magic != kSlabMagic) return B200FFT_INVALID_PLAN;
  if (!ptr) return B200FFT_INVALID_VALUE;
  if (!s->natural) return B200FFT_NOT_SUPPORTED;
  *ptr = s->back;
  return B200FFT_SUCCESS;
}

int b200fftExecSlab(b200fftSlabHandle s, const void* in, void* out, int direction, double scale, int layout, b200fftStream stream_) {
  if (!s || s->magic != kSlabMagic) return B200FFT_INVALID_PLAN;
  if (!in || !out || in == out) return B200FFT_INVALID_VALUE;
  if (direction != B200FFT_FORWARD && direction != B200FFT_INVERSE) return B200FFT_INVALID_VALUE;
  if (layout != B200FFT_SLAB_TRANSPOSED_OUT && layout != B200FFT_SLAB_NATURAL_OUT) return B200FFT_INVALID_VALUE;
  if (layout == B200FFT_SLAB_NATURAL_OUT && !s->natural) return B200FFT_NOT_SUPPORTED;
  cudaStream_t st = (cudaStream_t)stream_;
  const b200fft_slab_s::Pipe* q = &s->pipe[layout];
  const int64_t esz = s->esz, plane = s->h * s->w * esz, rplane = s->hl * s->w * esz, row = s->w * esz;
  int e = B200FFT_SUCCESS;
#define CU(x) do { if ((x) != cudaSuccess) { cudaGetLastError(); return B200FFT_EXEC_FAILED; } } while (0)
#define OK(x) do { if ((e = (x))) return e; } while (0)
  // the side streams start after whatever the caller has queued before this call
  CU(cudaEventRecord(s->e_start, st));
  CU(cudaStreamWaitEvent(s->sy, s->e_start, 0));
  CU(cudaStreamWaitEvent(s->sz, s->e_start, 0));
  CU(cudaStreamWaitEvent(s->sb, s->e_start, 0));
  // every rank has finished the previous transform: its z pass no longer reads `recv`, nobody reads `back`
  OK(launch_barrier(s, s->sb));
  CU(cudaEventRecord(s->e_b0, s->sb));
  CU(cudaStreamWaitEvent(s->sy, s->e_b0, 0));
  // x: rows of the local planes, plane chunk by plane chunk, on the caller's stream
  for (int p = 0; p < q->cp; p++) {
    const int64_t off = (int64_t)p * q->dlc * plane;
    OK(b200fftExec(q->px, (const char*)in + off, s->tmp + off, direction, st));
    CU(cudaEventRecord(s->e_x[p], st));
  }
  void* targets[kMaxRanks];
  for (int c = 0; c < q->ck; c++) {
    const int64_t coff = (int64_t)c * q->wc * esz;
    // y: columns of the local planes, ky row -> its owner's recv[kyl][rank*dl + z][kx]
    for (int p = 0; p < q->cp; p++) {
      if (c == 0) CU(cudaStreamWaitEvent(s->sy, s->e_x[p], 0));
      const int64_t z0 = (int64_t)p * q->dlc;
      for (int r = 0; r < s->P; r++) targets[r] = (char*)s->recv_p[r] + ((int64_t)s->rank * s->dl + z0) * row + coff;
      OK(b200fftExecScatterOn(q->py, s->tmp + z0 * plane + coff, targets, s->P, s->w, s->d * s->w, direction, 1.0, q->y_ctas, s->sy));
    }
    CU(cudaEventRecord(s->e_y[c], s->sy));
    CU(cudaStreamWaitEvent(s->sb, s->e_y[c], 0));
    OK(launch_barrier(s, s->sb));                     // this column chunk has landed everywhere
    CU(cudaEventRecord(s->e_b[c], s->sb));
    CU(cudaStreamWaitEvent(s->sz, s->e_b[c], 0));
    // z: columns along z of recv[hl][D][W], chunk c of the columns
    if (layout == B200FFT_SLAB_TRANSPOSED_OUT) {
      OK(b200fftExecScaled(q->pz, s->recv + coff, (char*)out + coff, direction, scale, s->sz));
    } else {
      // kz plane -> its owner's back[kzl][rank*hl + kyl][kx]
      for (int r = 0; r < s->P; r++) targets[r] = (char*)s->back_p[r] + (int64_t)s->rank * rplane + coff;
      OK(b200fftExecScatterOn(q->pz, s->recv + coff, targets, s->P, s->w, s->h * s->w, direction, scale, 0, s->sz));
    }
  }
  CU(cudaEventRecord(s->e_sz, s->sz));
  if (layout == B200FFT_SLAB_NATURAL_OUT) {
    CU(cudaStreamWaitEvent(s->sb, s->e_sz, 0));
    OK(launch_barrier(s, s->sb));                     // every kz plane has landed in its z-slab
    if (out != (void*)s->back) CU(cudaMemcpyAsync(out, s->back, (size_t)s->dl * plane, cudaMemcpyDeviceToDevice, s->sb));
  }
  CU(cudaEventRecord(s->e_sb, s->sb));
  CU(cudaEventRecord(s->e_sy, s->sy));
  CU(cudaStreamWaitEvent(st, s->e_sz, 0));
  CU(cudaStreamWaitEvent(st, s->e_sb, 0));
  CU(cudaStreamWaitEvent(st, s->e_sy, 0));
#undef CU
#undef OK
  return B200FFT_SUCCESS;
}

int b200fftDestroySlab(b200fftSlabHandle s) {
  if (!s || s->magic != kSlabMagic) return B200FFT_INVALID_PLAN;
  s->magic = 0;
  cudaDeviceSynchronize();
  free_plans(&s->pipe[0]);
  free_plans(&s->pipe[1]);
  for (void* q : s->opened) cudaIpcCloseMemHandle(q);
  for (void* q : {(void*)s->tmp, (void*)s->recv, (void*)s->back, (void*)s->flags})
    if (q) cudaFree(q);
  for (cudaStream_t q : {s->sy, s->sz, s->sb})
    if (q) cudaStreamDestroy(q);
  for (cudaEvent_t q : {s->e_start, s->e_b0, s->e_sy, s->e_sz, s->e_sb})
    if (q) cudaEventDestroy(q);
  for (auto* v : {&s->e_x, &s->e_y, &s->e_b})
    for (cudaEvent_t q : *v) cudaEventDestroy(q);
  cudaGetLastError();
  delete s;
  return B200FFT_SUCCESS;
}

}  // extern "C"
