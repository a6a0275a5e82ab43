// Planner + executor + C ABI of libb200fft (see include/b200fft.h).
//
// A plan is an immutable list of passes; each pass is one launch of a line kernel (fft_kernel.cuh,
// generic_kernel.cu) = one HBM read + one HBM write of the whole array.  Exec enqueues the passes on
// the caller's stream; scratch is stream-ordered (cudaMallocAsync) so concurrent execs of one plan
// on different streams do not share state (cf. PTX/Plans.hs:86 -- exec runs outside the cache lock).
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200fft.h"
#include "generic.h"
#include "registry.h"
#include "fft_kernel.cuh"
#include "fused_kernel.cuh"
#include "band_kernel.cuh"

namespace b200fft {

// ------------------------------------------------------------------------------------------
// registry
// ------------------------------------------------------------------------------------------
static std::vector<KernelEntry>& reg() {
  static std::vector<KernelEntry> r;
  return r;
}
static void add_entry(const KernelEntry& e_) {
  KernelEntry e = e_;
  e.variant = 0;
  for (const auto& o : reg())
    if (o.is_double == e.is_double && o.N == e.N && o.flavor == e.flavor && o.tw4 == e.tw4) e.variant++;
  reg().push_back(e);
}
static std::vector<FusedEntry>& freg() {
  static std::vector<FusedEntry> r;
  return r;
}
static void add_fused(const FusedEntry& e) { freg().push_back(e); }
static std::vector<BandEntry>& breg() {
  static std::vector<BandEntry> r;
  return r;
}
static void add_band(const BandEntry& e) { breg().push_back(e); }
static std::once_flag g_reg_once;
static void ensure_registry() {
  std::call_once(g_reg_once, [] {
    register_f32_small(add_entry);
    register_f32_large(add_entry);
    register_f32_col(add_entry);
    register_f64_small(add_entry);
    register_f64_large(add_entry);
    register_f64_col(add_entry);
    register_ring(add_entry);
    register_pair(add_entry);
    register_cluster(add_entry);
    register_pipe(add_entry);
    register_ringcol(add_entry);
    register_fused(add_fused);
    register_band(add_band);
  });
}

// Developer override: B200FFT_VARIANTS="r4096d=1,c1024f=2" picks registration-order variant 1 of the
// c128 row kernel for N=4096, etc. (flavour r/c/t/g/p/k/q = row/col/trans/ring/pair/cluster/pipe, N, type f/d).  Default is variant 0.
static int forced_variant(int is_double, int N, int flavor) {
  const char* env = getenv("B200FFT_VARIANTS");
  if (!env) return 0;
  char key[64];
  snprintf(key, sizeof key, "%c%d%c=", flavor == FL_ROW ? 'r' : flavor == FL_COL ? 'c' : flavor == FL_RING ? 'g' : flavor == FL_ROWPAIR ? 'p' : flavor == FL_CLUSTER ? 'k' : flavor == FL_PIPE ? 'q' : 't', N,
           is_double ? 'd' : 'f');
  const char* p = env;
  while ((p = strstr(p, key)) != nullptr) {
    if (p == env || p[-1] == ',') return atoi(p + strlen(key));
    p++;
  }
  return 0;
}

// max_tl > 0: the number of lines available to a tile -- among the variants, the first (registration order) whose
// tile is not wider than that, else the narrowest; a forced variant (B200FFT_VARIANTS) always wins.
const KernelEntry* find_kernel(int is_double, int N, int flavor, int tw4, int max_tl) {
  ensure_registry();
  const int want = forced_variant(is_double, N, flavor);
  const bool forced = getenv("B200FFT_VARIANTS") && want > 0;
  const KernelEntry *first = nullptr, *fit = nullptr, *narrow = nullptr;
  for (const auto& e : reg()) {
    if (e.is_double != is_double || e.N != N || e.flavor != flavor || e.tw4 != tw4) continue;
    if (!first) first = &e;
    if (forced && e.variant == want) return &e;
    if (!fit && max_tl > 0 && e.TL <= max_tl) fit = &e;
    if (!narrow || e.TL < narrow->TL) narrow = &e;
  }
  if (max_tl > 0 && first && first->TL > max_tl) return fit ? fit : narrow;
  return first;
}
const FusedEntry* find_fused(int is_double, int NA, int flavA, int twA, int NB, int flavB) {
  ensure_registry();
  // Opt-in (B200FFT_FUSED=1).  Measured on B200 (profiles/r01_fused_l2_staging.txt): the L2 staging works -- ncu
  // shows exactly one HBM read + one HBM write of the array for both phases together -- but the c64 phases are
  // instruction-issue bound (~55 issued instructions per point and phase, 70 % issue utilisation), not HBM
  // bound, so removing the HBM round trip does not make them faster yet: cfg3 600 us fused vs 549 us unfused.
  const char* on = getenv("B200FFT_FUSED");
  if (!(on && atoi(on))) return nullptr;
  for (const auto& e : freg())
    if (e.a.is_double == is_double && e.a.N == NA && e.a.flavor == flavA && e.a.tw4 == twA && e.b.N == NB && e.b.flavor == flavB && e.b.tw4 == 0)
      return &e;
  return nullptr;
}
const BandEntry* find_band(int is_double, int mode, int outer, int N1, int N2) {
  ensure_registry();
  int want = 0, seen = 0;
  if (const char* v = getenv("B200FFT_BAND_VARIANT")) want = atoi(v);
  const BandEntry* first = nullptr;
  for (const auto& e : breg())
    if (e.is_double == is_double && e.mode == mode && e.outer == outer && e.N1 == N1 && e.N2 == N2) {
      if (!first) first = &e;
      if (seen++ == want) return &e;
    }
  return first;
}
int list_kernels(const KernelEntry** out, int max) {
  ensure_registry();
  int n = 0;
  for (const auto& e : reg()) { if (n < max) out[n] = &e; n++; }
  return n;
}

static std::atomic<long long> g_launches{0};

// Library-owned stream-ordered memory pool for per-exec scratch.  The device's default pool trims
// itself to zero at every synchronisation point, which turns each exec of a multi-pass plan into a
// fresh multi-GiB physical allocation (measured: +13 ms per 2 GiB exec on B200).  Our pool keeps
// what it has been given (release threshold = max) so steady-state execs re-use the same pages.
static std::mutex g_pool_lock;
static std::vector<std::pair<int, cudaMemPool_t>> g_pools;   // one per device
static cudaMemPool_t scratch_pool() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> g(g_pool_lock);
  for (auto& kv : g_pools) if (kv.first == dev) return kv.second;
  cudaMemPoolProps props{};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = dev;
  cudaMemPool_t pool = nullptr;
  if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  unsigned long long thr = ~0ull;
  cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  g_pools.emplace_back(dev, pool);
  return pool;
}
static cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t s) {
  cudaMemPool_t pool = scratch_pool();
  if (!pool) return cudaErrorMemoryAllocation;
  return cudaMallocFromPoolAsync(p, bytes, pool, s);
}
// the host-buffer entry points (host_api.cu) stage their chunks in the same pool
cudaError_t pool_alloc(void** p, size_t bytes, cudaStream_t s) { return scratch_alloc(p, bytes, s); }

// ------------------------------------------------------------------------------------------
// plan representation
// ------------------------------------------------------------------------------------------
enum Buf { BUF_IN = 0, BUF_OUT = 1, BUF_SCRATCH = 2 };

enum PassKind { PK_LINES = 0, PK_GENERIC = 1, PK_BLUESTEIN = 2, PK_COPY = 3, PK_FUSED2 = 4, PK_CLUSTER = 5, PK_BAND = 6 };

struct Pass {
  int kind = PK_LINES;
  const KernelEntry* k = nullptr;  // PK_LINES
  const KernelEntry* ring = nullptr;  // PK_LINES rows: persistent TMA-fed alternative (needs 16-byte aligned buffers)
  int ring_ntl = 0, ring_grid = 0;
  Geom g{};
  GenericPass gp{};                // PK_GENERIC / PK_BLUESTEIN
  bool inplace_ok = false;         // reads and writes the same positions tile by tile
  bool axis_last = false;          // the pass that produces an axis' final (natural-order) output index
  int src = BUF_IN, dst = BUF_OUT;
  void* tws = nullptr;             // device: stage twiddles
  void* tw_lo = nullptr;           // device: four-step twiddle tables
  void* tw_hi = nullptr;
  void* ctw = nullptr;             // PK_CLUSTER: inner twiddles w_N^(k1*r), [CS-1][N1]
  // persistent software-pipelined alternative of a column pass (pipe_kernel.cuh); needs 16-byte aligned buffers
  const KernelEntry* pipe = nullptr;
  void* ptws = nullptr;            // its stage twiddles and inner twiddles
  void* pctw = nullptr;
  int pipe_ntl = 0, pipe_grid = 0;
  long long ntiles = 0;
  const FusedEntry* fz = nullptr;  // PK_FUSED2
  FusedParams fp{};
  void* twsB = nullptr;            // PK_FUSED2: stage twiddles of phase B (tws = phase A)
  bool mid_in_dst = false;         // PK_FUSED2: phase A writes into dst (in place there) instead of scratch slots
  int fused_grid = 0;              // PK_FUSED2: persistent CTAs
  bool first_swap_a = true;
  // PK_BAND: two phases fused through L2 (band_kernel.cuh); tws / twsB = stage twiddles, tw_lo / tw_hi = inner four-step table
  const BandEntry* bz = nullptr;
  BandParams bp{};
  // persistent single-buffer TMA-fed column kernel (ringcol_kernel.cuh) for tiles that fill an SM; rctws = its stage twiddles
  const KernelEntry* ringcol = nullptr;
  void* rctws = nullptr;
  int ringcol_ntl = 0, ringcol_grid = 0;
  const KernelEntry* ringtrans = nullptr;     // ... and its transposing (rows in, line-fastest out) sibling
  void* rttws = nullptr;
  int ringtrans_ntl = 0, ringtrans_grid = 0;
  size_t slot_bytes = 0, counter_bytes = 0, counter_off = 0;   // fused / band passes: their share of the plan's band scratch
  void* otw_lo = nullptr;          // outer four-step table (OUTER kernels)
  void* otw_hi = nullptr;
  unsigned long long tm_dims[4] = {1, 1, 1, 1};     // tensor map of phase A's input: extents (elements of the real type) ...
  unsigned long long tm_strides[3] = {0, 0, 0};     // ... byte strides of dims 1..3 ...
  unsigned tm_box[4] = {1, 1, 1, 1};                // ... and the box of one tile
  int band_grid = 0;
  std::string desc;
};

}  // namespace b200fft

using namespace b200fft;

struct b200fft_plan_s {
  unsigned magic = 0xB200FF7u;
  bool upload_failed = false;   // a twiddle table could not be allocated / copied: the plan must not be handed out
  int is_double = 0;
  int rank = 0;
  long long dims[3] = {1, 1, 1};
  long long batch = 1;
  long long total = 0;  // complex elements in / out
  std::vector<Pass> passes;
  std::vector<void*> dev_allocs;
  size_t scratch_bytes = 0;   // main scratch (same size as the array) if any pass uses BUF_SCRATCH
  size_t extra_bytes = 0;     // bluestein workspace
  size_t band_bytes = 0;      // fused / band passes: the L2-resident slots + every such pass's progress counters
  size_t band_slot_bytes = 0; // ... of which the slots (the counters follow)
  bool no_shift = false;      // built by a whole-transform builder that does not mark the axes' last passes (b200fftExecShifted)
  // Plans holding a PK_BAND pass need 16-byte aligned buffers (TMA) and cannot rotate their stores: the same transform
  // planned without band passes serves misaligned buffers and b200fftExecShifted.
  b200fft_plan_s* fallback = nullptr;
};

namespace b200fft {

static size_t esize(const b200fft_plan_s* p) { return p->is_double ? 16 : 8; }

template <typename T>
static void* upload(b200fft_plan_s* p, const std::vector<T>& h) {
  void* d = nullptr;
  if (h.empty()) return nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) { cudaGetLastError(); p->upload_failed = true; return nullptr; }
  p->dev_allocs.push_back(d);
  if (cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); p->upload_failed = true; }
  return d;
}

// exp(-2 pi i num/den) in long double with the argument reduced exactly
static void unit_root(long long num, long long den, long double* re, long double* im) {
  num %= den;
  if (num < 0) num += den;
  // octant reduction for accuracy
  const long double two_pi = 6.283185307179586476925286766559005768L;
  long double a = two_pi * (long double)num / (long double)den;
  *re = cosl(a);
  *im = -sinl(a);
}

template <typename T>
static void fill_root_table(std::vector<T>& v, size_t off, long long num, long long den) {
  long double re, im;
  unit_root(num, den, &re, &im);
  v[2 * off] = (T)re;
  v[2 * off + 1] = (T)im;
}

// stage twiddles of a line kernel: stage s>=1, entry [(r-1)*Ns + k] = w_{Ns*R}^{r k}
static void* make_stage_twiddles(b200fft_plan_s* p, const KernelEntry* k) {
  if (k->tw_len == 0) return nullptr;
  auto build = [&](auto tag) -> void* {
    using T = decltype(tag);
    std::vector<T> h(2 * (size_t)k->tw_len);
    size_t off = 0;
    int Ns = k->rad[0];
    for (int s = 1; s < k->S; s++) {
      int R = k->rad[s];
      for (int r = 1; r < R; r++)
        for (int kk = 0; kk < Ns; kk++) fill_root_table(h, off + (size_t)(r - 1) * Ns + kk, (long long)r * kk, (long long)Ns * R);
      off += (size_t)(R - 1) * Ns;
      Ns *= R;
    }
    return upload(p, h);
  };
  return p->is_double ? build(double{}) : build(float{});
}

// four-step twiddle w_L^x split as lo[x & (2^lb - 1)] * hi[x >> lb]
static void make_fourstep_tables(b200fft_plan_s* p, long long L, int* lo_bits, void** lo, void** hi) {
  int lg = 0;
  while ((1LL << lg) < L) lg++;
  int lb = (lg + 1) / 2;
  *lo_bits = lb;
  long long nlo = 1LL << lb, nhi = (L + nlo - 1) / nlo;
  auto build = [&](auto tag) {
    using T = decltype(tag);
    std::vector<T> a(2 * (size_t)nlo), b(2 * (size_t)nhi);
    for (long long i = 0; i < nlo; i++) fill_root_table(a, (size_t)i, i, L);
    for (long long i = 0; i < nhi; i++) fill_root_table(b, (size_t)i, i << lb, L);
    *lo = upload(p, a);
    *hi = upload(p, b);
  };
  if (p->is_double) build(double{}); else build(float{});
}

static bool is_pow2(long long n) { return n > 0 && (n & (n - 1)) == 0; }
static int ilog2(long long n) { int l = 0; while ((1LL << l) < n) l++; return l; }

struct Builder {
  b200fft_plan_s* p;
  int err = 0;

  Pass& push(Pass ps) { p->passes.push_back(std::move(ps)); return p->passes.back(); }

  // ---- one launch of a pow2 line kernel --------------------------------------------------
  // array viewed as [nb][no][N-axis...]; see Geom.  Returns false if no kernel exists.
  bool lines_pass(int N, int flavor, bool tw4, Geom g, long long twL, bool inplace_ok, int prefer_tl, const char* what) {
    (void)prefer_tl;
    const KernelEntry* k = find_kernel(p->is_double, N, flavor, tw4 ? 1 : 0, (flavor == FL_COL || flavor == FL_TRANS) ? g.nl : 0);
    if (!k) return false;
    Pass ps;
    ps.kind = PK_LINES;
    ps.k = k;
    g.ntl = (g.nl + k->TL - 1) / k->TL;
    {  // the kernels address a thread's points with 32-bit byte strides
      const long long tpt = k->N / k->E, esz = p->is_double ? 16 : 8;
      if (tpt * g.ins * esz >= (1LL << 32) || tpt * g.ons * esz >= (1LL << 32)) { err = B200FFT_INVALID_SIZE; return false; }
    }
    ps.g = g;
    ps.inplace_ok = inplace_ok;
    ps.tws = make_stage_twiddles(p, k);
    if (tw4) make_fourstep_tables(p, twL, &ps.g.tw_lo_bits, &ps.tw_lo, &ps.tw_hi);
    ps.ntiles = (long long)g.nb * g.no * ps.g.ntl;
    if (ps.ntiles >= (1LL << 31)) { err = B200FFT_INVALID_SIZE; return false; }
    char buf[256];
    snprintf(buf, sizeof buf, "%s: lines N=%d v%d E=%d TL=%d minb=%d %s%s radix=%dx%dx%dx%d threads=%d smem=%zu tiles=%lld", what, k->N,
             k->variant, k->E, k->TL, k->minb, flavor == FL_ROW ? "row" : flavor == FL_COL ? "col" : flavor == FL_ROWPAIR ? "row+radix2" : "trans",
             tw4 ? "+tw" : "", k->rad[0], k->rad[1],
             k->rad[2], k->rad[3], k->threads, k->smem, ps.ntiles);
    ps.desc = buf;
    if (flavor == FL_ROW && !tw4 && g.ils == N && g.ols == N && g.ins == 1 && g.ons == 1 && g.nb == 1 && g.no == 1 &&
        !(getenv("B200FFT_NO_RING") && atoi(getenv("B200FFT_NO_RING")))) {
      const KernelEntry* r = find_kernel(p->is_double, N, FL_RING, 0, 0);
      // short c64 rows: the ring only pays for small batches (the plain kernel streams big ones at ~100 %)
      // (opt-in via B200FFT_RING_C64_MAX_LINES; measured slower than the plain kernel at every batch size,
      //  cfg1: 19.7-20.7 us against 16.7 us -- profiles/r01_pair2d_and_narrow_columns.txt)
      if (r && !p->is_double && N < env_int("B200FFT_RING_C64_MIN_N", 8192) && g.nl > env_int("B200FFT_RING_C64_MAX_LINES", 0)) r = nullptr;
      // worth it only when every SM gets a few tiles to pipeline
      if (r && (g.nl + r->TL - 1) / r->TL >= 2LL * 148) {
        ps.ring = r;
        ps.ring_ntl = (g.nl + r->TL - 1) / r->TL;
        snprintf(buf, sizeof buf, " | ring: G=%d NS=%d threads=%d smem=%zu tiles=%d", r->G, r->NS, r->threads, r->smem, ps.ring_ntl);
        ps.desc += buf;
      }
    }
    if (flavor == FL_COL) attach_pipe(ps, N, tw4);
    // (not beside the pipelined kernel: for the slab transform's NVLink-bound scatter passes it measured no better than the
    //  lock-step loop kernel -- 8 GPUs 1.92-2.00 ms against 1.93, profiles/r02_slab_cabi.txt)
    if (flavor == FL_COL && !ps.pipe) attach_ringcol(ps, N, tw4);
    if (flavor == FL_TRANS && !tw4) attach_ringtrans(ps, N);
    if (flavor == FL_ROW && !tw4) attach_pipe_rows(ps, N);
    push(ps);
    return true;
  }


  // inner four-step twiddles of a cluster kernel: w_N^(k1 * r), r = 1 .. CS-1, k1 < N1
  void* make_cluster_twiddles(const KernelEntry* k) {
    if (k->CS <= 1) return nullptr;
    const long long N = (long long)k->N1 * k->CS;
    auto build = [&](auto tag) -> void* {
      using T = decltype(tag);
      std::vector<T> h(2 * (size_t)(k->CS - 1) * k->N1);
      for (int r = 1; r < k->CS; r++)
        for (int k1 = 0; k1 < k->N1; k1++) fill_root_table(h, (size_t)(r - 1) * k->N1 + k1, (long long)r * k1, N);
      return upload(p, h);
    };
    return p->is_double ? build(double{}) : build(float{});
  }

  // Column tiles that fill an SM's shared memory (one lock-step CTA per SM, load / butterflies / store in turn): the persistent
  // single-buffer TMA-fed kernel fetches the next tile while the last stage and the stores of the current one run.
  void attach_ringcol(Pass& ps, long long N, bool tw4) {
    if (getenv("B200FFT_RINGCOL") && atoi(getenv("B200FFT_RINGCOL")) == 0) return;
    const KernelEntry* q = find_kernel(p->is_double, (int)N, FL_RINGCOL, tw4 ? 1 : 0, 0);
    // instead of a lock-step tile that fills the SM -- or, for the +tw kernels, one half as wide (128 B runs against 256 B)
    if (!q || (ps.k->smem < 100 * 1024 && q->TL <= ps.k->TL)) return;
    const Geom& g = ps.g;
    const long long esz = p->is_double ? 16 : 8;
    if (g.nb != 1 || g.ils != 1 || g.ols != 1 || g.nl % q->TL) return;
    if ((g.ins * esz) % 16 || (g.ios * esz) % 16 || 2LL * g.nl >= (1LL << 32) || g.ins * esz >= (1LL << 40) || g.ios * esz >= (1LL << 40)) return;
    {
      const long long tpt = q->N / q->E;
      if (tpt * g.ons * esz >= (1LL << 32)) return;
    }
    const long long ntl = g.nl / q->TL, ntiles = (long long)g.no * ntl;
    if (ntiles < 148 || ntiles >= (1LL << 31)) return;
    ps.ringcol = q;
    ps.ringcol_ntl = (int)ntl;
    ps.rctws = make_stage_twiddles(p, q);
    char buf[200];
    snprintf(buf, sizeof buf, " | ring cols: persistent, TMA-fed, one buffer, TL=%d threads=%d smem=%zu tiles=%lld", q->TL, q->threads, q->smem, ntiles);
    ps.desc += buf;
  }

  // The transposing last pass of a big four-step through the single-buffer TMA-fed kernel when that makes the store runs longer.
  void attach_ringtrans(Pass& ps, long long N) {
    if (getenv("B200FFT_RINGCOL") && atoi(getenv("B200FFT_RINGCOL")) == 0) return;
    const KernelEntry* q = find_kernel(p->is_double, (int)N, FL_RINGTRANS, 0, 0);
    if (!q || q->TL <= ps.k->TL) return;
    const Geom& g = ps.g;
    const long long esz = p->is_double ? 16 : 8;
    if (g.ins != 1 || g.ols != 1 || g.nl % q->TL) return;
    if ((g.ils * esz) % 16 || (g.ios * esz) % 16 || (g.ibs * esz) % 16) return;
    {
      const long long tpt = q->N / q->E;
      if (tpt * g.ons * esz >= (1LL << 32)) return;
    }
    const long long ntl = g.nl / q->TL, ntiles = (long long)g.nb * g.no * ntl;
    if (ntiles < 148 || ntiles >= (1LL << 31)) return;
    ps.ringtrans = q;
    ps.ringtrans_ntl = (int)ntl;
    ps.rttws = make_stage_twiddles(p, q);
    char buf[200];
    snprintf(buf, sizeof buf, " | ring trans: persistent, TMA-fed, one buffer, TL=%d threads=%d smem=%zu tiles=%lld", q->TL, q->threads, q->smem, ntiles);
    ps.desc += buf;
  }

  // Attach the persistent pipelined kernel to a column pass over [O][N][I] whose Geom is ps.g (ils == ols == 1).
  void attach_pipe(Pass& ps, long long N, bool tw4) {
    // B200FFT_PIPE = 0 never, 1 whenever legal, unset = where it was measured to win (profiles/r01_pipe_and_cluster.txt):
    //  * plain (CS = 1) and CS = 2: 84-91 % of HBM peak against 60-68 % for the lock-step kernel -- as long as the rows
    //    of one tile span <= 256 MB.  Beyond that every row sits in a 2 MB page of its own and the 128-entry TLB
    //    thrashes (57 % at 512 MB, 36 % at 8 GB), which the lock-step kernel tolerates better;
    //  * clusters of 4-16 are bound by shared-memory instruction issue (mio_throttle) and lose to the two-pass
    //    four-step through HBM: opt-in only.
    const char* mode = getenv("B200FFT_PIPE");
    if (mode && atoi(mode) == 0) return;
    const bool force = mode && atoi(mode) > 0;
    const KernelEntry* q = find_kernel(p->is_double, (int)N, FL_PIPE, tw4 ? 1 : 0, 0);
    if (!q) return;
    const Geom& g = ps.g;
    const long long esz = p->is_double ? 16 : 8;
    //  * c128: the 512-point lock-step kernel already fits 2-3 CTAs per SM and runs at 90-98 % (pipelined: 70-75 %);
    //    1024 / 2048 points tie or lose -- c128 stays opt-in.
    //    Exception: c128 N=1024 in ONE CTA (4 columns): 78 % against 54 % for its 128 KB lock-step tile.
    if (!force && ((p->is_double && !(q->CS == 1 && q->N1 == 1024)) || q->CS > 2 || N * g.ins * esz > (256LL << 20))) return;
    if (g.ils != 1 || g.ols != 1 || g.nl % q->TL) return;
    // every tile row must start on a 16-byte boundary (cp.async 16) and the kernels use 32-bit byte offsets per tile
    if ((g.ins * esz) % 16 || (g.ios * esz) % 16 || (g.ibs * esz) % 16) return;
    {  // 32-bit byte strides between a thread's points (as in lines_pass); row offsets of the landing copies are 64-bit
      const long long tpt = q->N1 / q->E;
      if (tpt * g.ons * esz >= (1LL << 32) || (long long)q->CS * g.ins * esz >= (1LL << 32)) return;
    }
    const long long ntl = g.nl / q->TL, ntiles = (long long)g.nb * g.no * ntl;
    if (ntiles >= (1LL << 31) / 16 || ntiles < env_int("B200FFT_PIPE_MIN_TILES", 2 * 148 / q->CS)) return;
    ps.pipe = q;
    ps.pipe_ntl = (int)ntl;
    ps.ptws = make_stage_twiddles(p, q);
    ps.pctw = make_cluster_twiddles(q);
    char buf[200];
    snprintf(buf, sizeof buf, " | pipe: N=%dx%d E=%d TL=%d threads=%d smem=%zu tiles=%lld", q->N1, q->CS, q->E, q->TL, q->threads, q->smem, ntiles);
    ps.desc += buf;
  }

  // Persistent pipelined kernel for contiguous rows of one 64 KB line per tile (batched innermost-axis transforms).
  void attach_pipe_rows(Pass& ps, long long N) {
    const char* mode = getenv("B200FFT_PIPE");
    if (mode && atoi(mode) == 0) return;
    const bool force = mode && atoi(mode) > 0;
    const KernelEntry* q = find_kernel(p->is_double, (int)N, FL_PIPEROW, 0, 0);
    if (!q) return;
    const Geom& g = ps.g;
    if (g.ils != N || g.ols != N || g.ins != 1 || g.ons != 1 || g.nb != 1 || g.no != 1) return;
    if (g.nl < env_int("B200FFT_PIPE_MIN_TILES", 2 * 148)) return;
    if (!force && !pipe_rows_default(N)) return;
    ps.pipe = q;
    ps.pipe_ntl = g.nl;
    ps.ptws = make_stage_twiddles(p, q);
    ps.pctw = nullptr;
    ps.ring = nullptr;
    char buf[200];
    snprintf(buf, sizeof buf, " | pipe rows: E=%d threads=%d smem=%zu tiles=%d", q->E, q->threads, q->smem, g.nl);
    ps.desc += buf;
  }
  bool pipe_rows_default(long long N) const { (void)N; return false; }

  // ---- one launch of a cluster (DSMEM) column kernel: strided axis [O][N][I] in place, N = N1*CS ------
  static int cluster_min_n() { return env_int("B200FFT_CLUSTER_MIN_N", 4096); }
  bool cluster_pass(long long O, long long N, long long I, const char* what) {
    // Opt-in (B200FFT_CLUSTER=1).  Measured on B200 (profiles/r01_pipe_and_cluster.txt): correct and one HBM pass, but
    // at 44 % of HBM peak (8192-point c64 columns: 378 us) it does not beat the two four-step passes at 93-98 % each
    // (359 us): the lock-step load / exchange / store phases of 64 KB CTAs leave HBM idle most of the time.
    if (!(getenv("B200FFT_CLUSTER") && atoi(getenv("B200FFT_CLUSTER")))) return false;
    if (N < cluster_min_n() || N >= (1LL << 30)) return false;
    const KernelEntry* k = find_kernel(p->is_double, (int)N, FL_CLUSTER, 0, 0);
    if (!k) return false;
    const long long esz = p->is_double ? 16 : 8, tpt = k->N1 / k->E;
    if (tpt * k->CS * I * esz >= (1LL << 32)) return false;
    Pass ps;
    ps.kind = PK_CLUSTER;
    ps.k = k;
    Geom g{};
    g.nb = 1; g.no = (int)O; g.nl = (int)I;
    g.ios = N * I; g.ils = 1; g.ins = I;
    g.oos = N * I; g.ols = 1; g.ons = I;
    g.tw_div = 1;
    g.ntl = (g.nl + k->TL - 1) / k->TL;
    ps.g = g;
    ps.inplace_ok = true;
    ps.ntiles = (long long)O * g.ntl;
    if (ps.ntiles * k->CS >= (1LL << 31)) return false;
    ps.tws = make_stage_twiddles(p, k);
    ps.ctw = make_cluster_twiddles(k);
    char buf[256];
    snprintf(buf, sizeof buf, "%s: cluster cols N=%d=%dx%d v%d E=%d TL=%d minb=%d radix=%dx%dx%dx%d | %d threads=%d smem=%zu clusters=%lld", what, k->N,
             k->N1, k->CS, k->variant, k->E, k->TL, k->minb, k->rad[0], k->rad[1], k->rad[2], k->rad[3], k->CS, k->threads, k->smem, ps.ntiles);
    ps.desc = buf;
    attach_pipe(ps, N, false);
    push(ps);
    return true;
  }

  // ---- two phases in one launch with an L2-resident intermediate (fused_kernel.cuh) -----------
  static size_t counters_bytes(int nbands) { return (((size_t)(1 + 2 * nbands) * 4) + 255) / 256 * 256; }
  static long long band_target_bytes() {
    const char* e = getenv("B200FFT_BAND_MB");
    return (long long)(e && atoi(e) > 0 ? atoi(e) : 2) << 20;
  }
  static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e && atoi(e) > 0 ? atoi(e) : dflt; }

  bool fused_pass(const FusedEntry* fz, Geom ga, Geom gb, FusedParams fp, long long twL, bool mid_in_dst, bool inplace_ok, const char* what) {
    Pass ps;
    ps.kind = PK_FUSED2;
    ps.fz = fz;
    ga.ntl = (ga.nl + fz->a.TL - 1) / fz->a.TL;
    gb.ntl = (gb.nl + fz->b.TL - 1) / fz->b.TL;
    {
      const long long esz = p->is_double ? 16 : 8, ta = fz->a.N / fz->a.E, tb = fz->b.N / fz->b.E;
      if (ta * ga.ins * esz >= (1LL << 32) || ta * ga.ons * esz >= (1LL << 32) || tb * gb.ins * esz >= (1LL << 32) ||
          tb * gb.ons * esz >= (1LL << 32))
        return false;
    }
    // (phase A streams its input from HBM with evict-first loads and leaves its output in L2; phase B reads L2
    //  with ld.global.cg and streams its output to HBM with evict-first stores: compile-time in fused_kernel.cuh)
    fp.la = env_int("B200FFT_FUSED_LA", 8);
    if (fp.la > fp.nbands) fp.la = fp.nbands;
    fp.nslots = env_int("B200FFT_FUSED_SLOTS", fp.la + 6);
    if (fp.nslots < fp.la + 1) fp.nslots = fp.la + 1;
    const long long nA = (long long)ga.nb * ga.no * ga.ntl, nB = (long long)gb.nb * gb.no * gb.ntl;
    // tickets cover ~64 KB of tiles
    auto per_ticket = [&](const KernelEntry& k, const char* env) {
      const long long tile_bytes = (long long)k.N * k.TL * (long long)esize(p);
      (void)tile_bytes;   // measured: grouping tiles under one ticket only shrinks the pool of runnable tickets
      long long kk = env_int(env, 1);
      return (int)(kk < 1 ? 1 : (kk > 16 ? 16 : kk));
    };
    fp.ka = per_ticket(fz->a, "B200FFT_FUSED_KA");
    fp.kb = per_ticket(fz->b, "B200FFT_FUSED_KB");
    const long long total = (long long)fp.nbands * ((nA + fp.ka - 1) / fp.ka + (nB + fp.kb - 1) / fp.kb);
    if (nA >= (1LL << 30) || nB >= (1LL << 30) || total >= (1LL << 31)) return false;
    ps.tws = make_stage_twiddles(p, &fz->a);
    ps.twsB = make_stage_twiddles(p, &fz->b);
    if (fz->a.tw4) make_fourstep_tables(p, twL, &ga.tw_lo_bits, &ps.tw_lo, &ps.tw_hi);
    fp.a = ga; fp.b = gb;
    fp.nA = (int)nA; fp.nB = (int)nB;
    ps.fp = fp;
    ps.ntiles = total;
    ps.mid_in_dst = mid_in_dst;
    ps.inplace_ok = inplace_ok;
    ps.counter_bytes = counters_bytes(fp.nbands);
    ps.slot_bytes = mid_in_dst ? 0 : (size_t)fp.nslots * fp.slot_elems * esize(p);
    char buf[384];
    snprintf(buf, sizeof buf,
             "%s: fused A[N=%d %s%s E=%d TL=%d] -> %s -> B[N=%d %s E=%d TL=%d] threads=%d smem=%zu bands=%d tiles/band=%lld+%lld",
             what, fz->a.N, fz->a.flavor == FL_ROW ? "row" : "col", fz->a.tw4 ? "+tw" : "", fz->a.E, fz->a.TL,
             mid_in_dst ? "L2 (in place in dst)" : "L2 band slots", fz->b.N, fz->b.flavor == FL_COL ? "col" : "trans", fz->b.E, fz->b.TL,
             fz->threads, fz->smem, fp.nbands, nA, nB);
    snprintf(buf + strlen(buf), sizeof buf - strlen(buf), " tiles/ticket=%d+%d", fp.ka, fp.kb);
    ps.desc = buf;
    if (!mid_in_dst) {
      snprintf(buf, sizeof buf, " slot=%.1f MiB x%d lookahead=%d", (double)fp.slot_elems * esize(p) / 1048576.0, fp.nslots, fp.la);
      ps.desc += buf;
    }
    push(ps);
    return true;
  }

  // strided axis [O][N][I], N = N1*N2: phase A = FFT over n1 (+ twiddle w_N^(k1 n2)) for a band of Wb columns,
  // phase B = FFT over n2 with the index-reversed store.  In place at band granularity.
  bool try_fused_strided(long long O, long long N, long long I) {
    long long N1, N2;
    if (!split2(N, 1024, &N1, &N2)) return false;
    const FusedEntry* fz = find_fused(p->is_double, (int)N1, FL_COL, 1, (int)N2, FL_COL);
    if (!fz) return false;
    long long Wb = band_target_bytes() / (N * (long long)esize(p));
    while (Wb & (Wb - 1)) Wb &= Wb - 1;          // floor to a power of two
    const long long tl = fz->a.TL > fz->b.TL ? fz->a.TL : fz->b.TL;
    if (Wb < tl) Wb = tl;
    while (Wb > tl && I % Wb) Wb >>= 1;
    if (I % Wb || Wb % fz->a.TL || Wb % fz->b.TL) return false;
    if (N2 * I >= (1LL << 40)) return false;
    Geom ga{}, gb{};
    ga.nb = 1; ga.no = (int)N2; ga.nl = (int)Wb;
    ga.ios = I; ga.ils = 1; ga.ins = N2 * I;
    ga.oos = Wb; ga.ols = 1; ga.ons = N2 * Wb;
    ga.tw_div = 1; ga.tw_from_o = 1;
    gb.nb = 1; gb.no = (int)N1; gb.nl = (int)Wb;
    gb.ios = N2 * Wb; gb.ils = 1; gb.ins = Wb;
    gb.oos = I; gb.ols = 1; gb.ons = N1 * I;
    gb.tw_div = 1;
    FusedParams fp{};
    fp.nbi = (int)(I / Wb);
    if (O * fp.nbi >= (1LL << 24)) return false;
    fp.nbands = (int)(O * fp.nbi);
    fp.a_in_bo = N * I; fp.a_in_bi = Wb;
    fp.b_out_bo = N * I; fp.b_out_bi = Wb;
    fp.slot_elems = N * Wb;
    return fused_pass(fz, ga, gb, fp, N, false, true, "4step-strided");
  }

  // contiguous lines [O][N], N = N1*N2 <= 2^20: a band is a group of whole transforms (<= ~8 MiB)
  bool try_fused_contig2(long long O, long long N) {
    const int lg = ilog2(N);
    const long long N1 = 1LL << (lg / 2), N2 = N / N1;
    const FusedEntry* fz = find_fused(p->is_double, (int)N1, FL_COL, 1, (int)N2, FL_TRANS);
    if (!fz) return false;
    long long nbat = band_target_bytes() / (N * (long long)esize(p));
    if (nbat < 1) nbat = 1;
    if (nbat > O) nbat = O;
    while (O % nbat) nbat--;
    if (O / nbat >= (1LL << 24)) return false;
    Geom ga{}, gb{};
    ga.nb = (int)nbat; ga.no = 1; ga.nl = (int)N2;
    ga.ibs = N; ga.ils = 1; ga.ins = N2;
    ga.obs = N; ga.ols = 1; ga.ons = N2;
    ga.tw_div = 1;
    gb.nb = (int)nbat; gb.no = 1; gb.nl = (int)N1;
    gb.ibs = N; gb.ils = N2; gb.ins = 1;
    gb.obs = N; gb.ols = 1; gb.ons = N1;
    gb.tw_div = 1;
    FusedParams fp{};
    fp.nbi = 1;
    fp.nbands = (int)(O / nbat);
    fp.a_in_bo = nbat * N; fp.b_out_bo = nbat * N;
    fp.slot_elems = nbat * N;
    return fused_pass(fz, ga, gb, fp, N, false, true, "4step");
  }

  // second and third factor of a three-factor contiguous transform [O][N1][M], M = N2*N3, fused over bands of
  // Rb rows k1: phase A = FFT over n2 (+ twiddle w_M^(k2 n3)), phase B = rows n3 with the transposed store
  // X[k1 + N1 k2 + N1 N2 k3] whose contiguous runs are the band's Rb values of k1.
  bool try_fused_contig3_tail(long long O, long long N, long long N1, long long N2, long long N3) {
    const FusedEntry* fz = find_fused(p->is_double, (int)N2, FL_COL, 1, (int)N3, FL_TRANS);
    if (!fz) return false;
    const long long M = N2 * N3;
    long long Rb = fz->b.TL;
    { const char* e = getenv("B200FFT_ROWS_PER_BAND"); if (e && atoi(e) > 0) Rb = atoi(e); }
    if (N1 % Rb) return false;
    Geom ga{}, gb{};
    ga.nb = 1; ga.no = (int)Rb; ga.nl = (int)N3;
    ga.ios = M; ga.ils = 1; ga.ins = N3;
    ga.oos = M; ga.ols = 1; ga.ons = N3;
    ga.tw_div = 1;
    gb.nb = 1; gb.no = (int)N2; gb.nl = (int)Rb;
    gb.ios = N3; gb.ils = M; gb.ins = 1;
    gb.oos = N1; gb.ols = 1; gb.ons = N1 * N2;
    gb.tw_div = 1;
    FusedParams fp{};
    fp.nbi = (int)(N1 / Rb);
    if (O * fp.nbi >= (1LL << 24)) return false;
    fp.nbands = (int)(O * fp.nbi);
    fp.a_in_bo = N; fp.a_in_bi = Rb * M;
    fp.b_out_bo = N; fp.b_out_bi = Rb;
    fp.slot_elems = Rb * M;
    return fused_pass(fz, ga, gb, fp, M, false, false, "6step-BC");
  }

  // x rows + y columns of each z-plane of a [D][H][W] array, the plane staying in L2 between the two
  bool try_fused_plane(long long D, long long H, long long W) {
    struct Mark { b200fft_plan_s* p; size_t n0; ~Mark() { if (p->passes.size() > n0) p->no_shift = true; } } mark{p, p->passes.size()};
    if (!is_pow2(H) || !is_pow2(W) || D >= (1LL << 24)) return false;
    const FusedEntry* fz = find_fused(p->is_double, (int)W, FL_ROW, 0, (int)H, FL_COL);
    if (!fz) return false;
    Geom ga{}, gb{};
    ga.nb = 1; ga.no = 1; ga.nl = (int)H;
    ga.ils = W; ga.ins = 1; ga.ols = W; ga.ons = 1;
    ga.tw_div = 1;
    gb.nb = 1; gb.no = 1; gb.nl = (int)W;
    gb.ils = 1; gb.ins = W; gb.ols = 1; gb.ons = W;
    gb.tw_div = 1;
    FusedParams fp{};
    fp.nbi = 1;
    fp.nbands = (int)D;
    fp.a_in_bo = H * W; fp.mid_bo = H * W; fp.b_out_bo = H * W;
    fp.slot_elems = 0;
    return fused_pass(fz, ga, gb, fp, 0, true, true, "plane-xy");
  }

  // ---- two four-step phases in one persistent TMA-fed launch, intermediate in L2 (band_kernel.cuh) ------------
  bool no_band = false;
  // strided axis [O][N][I], N = N1*N2: bands of Wb adjacent columns, in place at band granularity
  bool try_band_strided(long long O, long long N, long long I, bool outer, long long outer_L, long long outer_col0) {
    if (no_band || (getenv("B200FFT_BAND") && atoi(getenv("B200FFT_BAND")) == 0)) return false;
    const BandEntry* bz = nullptr;
    long long N1 = 0, N2 = 0;
    for (long long n1 : {64LL, 128LL, 256LL, 32LL}) {
      if (N % n1) continue;
      bz = find_band(p->is_double, MODE_STRIDED, outer ? 1 : 0, (int)n1, (int)(N / n1));
      if (bz) { N1 = n1; N2 = N / n1; break; }
    }
    if (!bz) return false;
    const long long esz = (long long)esize(p);
    long long Wb = env_int("B200FFT_BAND_COLS", bz->TLA > bz->TLB ? bz->TLA : bz->TLB);
    if (Wb % bz->TLA || Wb % bz->TLB || I % Wb) return false;
    // TMA: 16-byte aligned strides, extents below 2^32, byte strides below 2^40
    if ((I * esz) % 16 || 2 * I >= (1LL << 32) || N * I * esz >= (1LL << 40) || O >= (1LL << 31)) return false;
    const long long nbi = I / Wb, nbands = O * nbi;
    const long long nA = N2 * (Wb / bz->TLA), nB = N1 * (Wb / bz->TLB);
    if (nbands >= (1LL << 24) || nbands * (nA + nB) >= (1LL << 31)) return false;
    if (nbands * (nA + nB) < 4 * 148) return false;              // too little work for a persistent launch
    if ((N1 * bz->b.N / bz->b.E) * I * esz >= (1LL << 32)) return false;   // 32-bit byte step between a thread's stores
    Pass ps;
    ps.kind = PK_BAND;
    ps.bz = bz;
    BandParams bp{};
    bp.nbands = (int)nbands; bp.nA = (int)nA; bp.nB = (int)nB;
    bp.la = 0;   // (unused by the two-list loader: phase A runs ahead as far as the slots allow)
    // 48 MiB of slots: measured optimum on cfg3 with 2 MiB slots (12: 541 us, 16: 504, 24: 472, 32: 486, 48: 554)
    bp.nslots = env_int("B200FFT_BAND_SLOTS", (int)((48LL << 20) / (N * Wb * esz)));
    if (bp.nslots < 4) bp.nslots = 4;
    if (bp.nslots > bp.nbands) bp.nslots = bp.nbands;
    bp.nbi = (int)nbi;
    bp.a_ncg = (int)(Wb / bz->TLA);
    bp.slot_elems = N * Wb;
    bp.out_bo = N * I; bp.out_bi = Wb; bp.out_ks = I;
    bp.wb = (int)Wb;
    bp.debug = env_int("B200FFT_BAND_DEBUG", 0);
    ps.tws = make_stage_twiddles(p, &bz->a);
    ps.twsB = make_stage_twiddles(p, &bz->b);
    make_fourstep_tables(p, N, &bp.tw_lo_bits, &ps.tw_lo, &ps.tw_hi);
    bp.tw_lo_n = 1 << bp.tw_lo_bits;
    bp.tw_hi_n = (int)((N + bp.tw_lo_n - 1) / bp.tw_lo_n);
    if (bp.tw_lo_n + bp.tw_hi_n > 512) return false;
    if (outer) {
      make_fourstep_tables(p, outer_L, &bp.otw_lo_bits, &ps.otw_lo, &ps.otw_hi);
      bp.otw_col0 = outer_col0;
    }
    ps.bp = bp;
    ps.tm_dims[0] = 2ull * (unsigned long long)I; ps.tm_dims[1] = (unsigned long long)N2; ps.tm_dims[2] = (unsigned long long)N1; ps.tm_dims[3] = (unsigned long long)O;
    ps.tm_strides[0] = (unsigned long long)(I * esz); ps.tm_strides[1] = (unsigned long long)(N2 * I * esz); ps.tm_strides[2] = (unsigned long long)(N * I * esz);
    ps.tm_box[0] = 2u * (unsigned)bz->TLA; ps.tm_box[1] = 1; ps.tm_box[2] = (unsigned)N1; ps.tm_box[3] = 1;
    ps.inplace_ok = true;
    ps.ntiles = nbands * (nA + nB);
    ps.counter_bytes = counters_bytes(bp.nbands);
    ps.slot_bytes = (size_t)bp.nslots * (size_t)bp.slot_elems * esize(p);
    char buf[384];
    snprintf(buf, sizeof buf,
             "4step-strided: band A[N=%d col+tw TL=%d] -> L2 slots -> B[N=%d col%s TL=%d] | persistent TMA-fed, threads=%d smem=%zu bands=%lld x %lld cols "
             "tiles/band=%lld+%lld slot=%.1f MiB x%d",
             bz->N1, bz->TLA, bz->N2, outer ? "+outer tw" : "", bz->TLB, bz->threads, bz->smem, nbands, Wb, nA, nB,
             (double)bp.slot_elems * esize(p) / 1048576.0, bp.nslots);
    ps.desc = buf;
    push(ps);
    return true;
  }

  // contiguous rows [O][R][M], M = N1*N2 (band_kernel.cuh, MODE_ROWS): a band is TLB adjacent rows (one contiguous piece of the
  // array); phase A = N1-point columns inside each row (+ twiddle w_M^(k1 n2)), phase B = N2-point rows of the intermediate with
  // the transposed store X[o*R*M + r + R*(k1 + N1*k2)] -- the LAST pass of a big four-step whose first pass left row r = the
  // outer output index.  Not in place.
  bool try_band_rows(long long O, long long R, long long M) {
    if (no_band || (getenv("B200FFT_BAND") && atoi(getenv("B200FFT_BAND")) == 0)) return false;
    const BandEntry* bz = nullptr;
    long long N1 = 0, N2 = 0;
    for (long long n1 : {128LL, 64LL, 256LL}) {
      if (M % n1) continue;
      bz = find_band(p->is_double, MODE_ROWS, 0, (int)n1, (int)(M / n1));
      if (bz) { N1 = n1; N2 = M / n1; break; }
    }
    if (!bz) return false;
    const long long esz = (long long)esize(p);
    const long long TLB = bz->TLB, TLA = bz->TLA;
    if (R % TLB || N2 % TLA) return false;
    const long long rows = O * R;
    if (2 * N2 >= (1LL << 32) || rows >= (1LL << 31) || rows * M * esz >= (1LL << 40)) return false;
    const long long nbi = R / TLB, nbands = O * nbi;
    const long long a_ncg = N2 / TLA, nA = TLB * a_ncg, nB = N1;
    if (nbands >= (1LL << 24) || nbands * (nA + nB) >= (1LL << 31)) return false;
    if (nbands * (nA + nB) < 4 * 148) return false;
    if ((N1 * bz->b.N / bz->b.E) * R * esz >= (1LL << 32)) return false;   // 32-bit byte step between a thread's stores
    Pass ps;
    ps.kind = PK_BAND;
    ps.bz = bz;
    BandParams bp{};
    bp.nbands = (int)nbands; bp.nA = (int)nA; bp.nB = (int)nB;
    bp.la = 0;
    bp.nslots = env_int("B200FFT_BAND_SLOTS", (int)((48LL << 20) / (TLB * M * esz)));
    if (bp.nslots < 4) bp.nslots = 4;
    if (bp.nslots > bp.nbands) bp.nslots = bp.nbands;
    bp.nbi = (int)nbi;
    bp.a_ncg = (int)a_ncg;
    bp.slot_elems = TLB * M;
    bp.out_bo = R * M; bp.out_bi = TLB; bp.out_ks = R;
    bp.wb = 0;
    bp.debug = env_int("B200FFT_BAND_DEBUG", 0);
    ps.tws = make_stage_twiddles(p, &bz->a);
    ps.twsB = make_stage_twiddles(p, &bz->b);
    make_fourstep_tables(p, M, &bp.tw_lo_bits, &ps.tw_lo, &ps.tw_hi);
    bp.tw_lo_n = 1 << bp.tw_lo_bits;
    bp.tw_hi_n = (int)((M + bp.tw_lo_n - 1) / bp.tw_lo_n);
    if (bp.tw_lo_n + bp.tw_hi_n > 512) return false;
    ps.bp = bp;
    // the rows as a tensor {2 N2, N1, rows, 1}; a phase-A tile is the box {2 TLA, N1, 1, 1}
    ps.tm_dims[0] = 2ull * (unsigned long long)N2; ps.tm_dims[1] = (unsigned long long)N1; ps.tm_dims[2] = (unsigned long long)rows; ps.tm_dims[3] = 1;
    ps.tm_strides[0] = (unsigned long long)(N2 * esz); ps.tm_strides[1] = (unsigned long long)(M * esz); ps.tm_strides[2] = (unsigned long long)(rows * M * esz);
    ps.tm_box[0] = 2u * (unsigned)TLA; ps.tm_box[1] = (unsigned)N1; ps.tm_box[2] = 1; ps.tm_box[3] = 1;
    ps.inplace_ok = false;
    ps.ntiles = nbands * (nA + nB);
    ps.counter_bytes = counters_bytes(bp.nbands);
    ps.slot_bytes = (size_t)bp.nslots * (size_t)bp.slot_elems * esize(p);
    char buf[384];
    snprintf(buf, sizeof buf,
             "4step-rows: band A[N=%d col+tw TL=%d] -> L2 slots -> B[N=%d trans TL=%d] | persistent TMA-fed, threads=%d smem=%zu bands=%lld x %lld rows "
             "tiles/band=%lld+%lld slot=%.1f MiB x%d",
             bz->N1, bz->TLA, bz->N2, bz->TLB, bz->threads, bz->smem, nbands, TLB, nA, nB, (double)bp.slot_elems * esize(p) / 1048576.0, bp.nslots);
    ps.desc = buf;
    push(ps);
    return true;
  }

  // contiguous lines [O][N], N = Nout * M: TWO HBM round trips, each a band pass with its intermediate in L2 --
  //   pass 1: the Nout-point strided axis of [O][Nout][M] + the outer twiddle w_N^(kout * m)  (MODE_STRIDED, OUTER)
  //   pass 2: the M-point rows with the transposed store X[kout + Nout * km]                  (MODE_ROWS)
  // cfg4 (2^28 = 2^14 x 2^14): replaces the three lines passes (3 HBM round trips, strict fraction capped at 0.67).
  bool try_band_contig2(long long O, long long N) {
    if (no_band || (getenv("B200FFT_BAND") && atoi(getenv("B200FFT_BAND")) == 0)) return false;
    // Opt-in (B200FFT_BAND_1D=1).  Measured on B200 (profiles/r02_band_experiments.txt): correct and two HBM round trips
    // (each band pass moves its 2 GiB in ~1.2 ms), but 2.39-2.42 ms against 2.35 ms for the three lines passes below: a band
    // pass costs four SM<->L2 transfers per element, and that, not HBM, is what bounds it.
    if (!(getenv("B200FFT_BAND_1D") && atoi(getenv("B200FFT_BAND_1D")))) return false;
    for (long long M : {16384LL}) {
      if (N % M) continue;
      const long long Nout = N / M;
      if (Nout < 4096 || Nout > 16384) continue;
      const size_t n0 = p->passes.size();
      if (try_band_strided(O, Nout, M, true, N, 0) && try_band_rows(O, Nout, M)) return true;
      p->passes.resize(n0);
    }
    return false;
  }

  // longest strided axis done in one pass (>= 64 B runs)
  int max_col_n() const { return env_int("B200FFT_MAX_COL_N", 2048); }
  int max_row_n() const { return p->is_double ? 8192 : 16384; }  // largest N with a row kernel

  // ---- transform along one axis of an array viewed as [O][N][I] (I = element stride of the axis)
  void axis(long long O, long long N, long long I) {
    const size_t before = p->passes.size();
    axis_impl(O, N, I);
    if (!err && p->passes.size() > before) p->passes.back().axis_last = true;
  }
  // A strided axis seen through a window: element (o, n, i) at o*os + n*ns + i, i < I (I may be a sub-range of the real row).
  // One in-place column pass only (power-of-two N <= max_col_n): the building block of the chunked slab transform.
  void axis_view(long long O, long long N, long long I, long long os, long long ns) {
    if (err) return;
    if (O >= (1LL << 31) || I >= (1LL << 31) || !is_pow2(N) || N < 2 || N > max_col_n()) { err = B200FFT_NOT_SUPPORTED; return; }
    Geom g{};
    g.nb = 1; g.no = (int)O; g.nl = (int)I;
    g.ios = os; g.ils = 1; g.ins = ns;
    g.oos = os; g.ols = 1; g.ons = ns;
    g.tw_div = 1;
    if (!lines_pass((int)N, FL_COL, false, g, 0, true, 0, "cols-view")) { if (!err) err = B200FFT_NOT_SUPPORTED; return; }
    p->passes.back().axis_last = true;
  }
  void axis_impl(long long O, long long N, long long I) {
    if (err || N == 1) return;
    if (O >= (1LL << 31) || I >= (1LL << 31) || N >= (1LL << 31)) { err = B200FFT_INVALID_SIZE; return; }
    if (!is_pow2(N)) { generic_axis(O, N, I); return; }
    if (I == 1) {
      if (N <= max_row_n()) {
        Geom g{};
        g.nb = 1; g.no = 1; g.nl = (int)O;
        g.ils = N; g.ins = 1; g.ols = N; g.ons = 1;
        g.tw_div = 1;
        if (!lines_pass((int)N, FL_ROW, false, g, 0, true, 0, "rows")) err = B200FFT_INTERNAL_ERROR;
        return;
      }
      fourstep_contig(O, N);
      return;
    }
    if (N <= max_col_n()) {
      // prefer a tile whose contiguous run is >= 128 B but keep the CTA's shared memory moderate
      Geom g{};
      g.nb = 1; g.no = (int)O; g.nl = (int)I;
      g.ios = N * I; g.ils = 1; g.ins = I;
      g.oos = N * I; g.ols = 1; g.ons = I;
      g.tw_div = 1;
      if (!lines_pass((int)N, FL_COL, false, g, 0, true, 0, "cols")) err = B200FFT_INTERNAL_ERROR;
      return;
    }
    if (cluster_pass(O, N, I, "cols")) return;
    fourstep_strided(O, N, I);
  }

  // choose N1*N2 = N (both pow2) with N1, N2 <= lim, as balanced as possible
  static bool split2(long long N, long long lim, long long* n1, long long* n2) {
    int lg = ilog2(N);
    int a = lg / 2, b = lg - a;
    if ((1LL << b) > lim) return false;
    *n1 = 1LL << a; *n2 = 1LL << b;
    return true;
  }

  // four-step along a strided axis: [O][N][I], N = N1*N2, n = n1*N2 + n2, k = k1 + N1*k2
  void fourstep_strided(long long O, long long N, long long I) {
    if (try_band_strided(O, N, I, false, 0, 0)) return;
    if (try_fused_strided(O, N, I)) return;
    long long N1, N2;
    // small sub-lengths keep the [N][TL] tiles small; cap at 1024 so TL stays >= 8
    if (!split2(N, 1024, &N1, &N2)) { err = B200FFT_NOT_SUPPORTED; return; }
    if (const char* sp = getenv("B200FFT_SPLIT2")) {   // developer override "n1,n2"
      long long a1 = 0, a2 = 0;
      if (sscanf(sp, "%lld,%lld", &a1, &a2) == 2 && a1 * a2 == N && is_pow2(a1) && is_pow2(a2)) { N1 = a1; N2 = a2; }
    }
    if (N2 * I >= (1LL << 31) || O * N1 >= (1LL << 31)) { err = B200FFT_INVALID_SIZE; return; }
    {  // pass A: FFT over n1, lines = (n2,i), twiddle w_N^(k1*n2), same positions
      Geom g{};
      g.nb = 1; g.no = (int)O; g.nl = (int)(N2 * I);
      g.ios = N * I; g.ils = 1; g.ins = N2 * I;
      g.oos = N * I; g.ols = 1; g.ons = N2 * I;
      g.tw_div = (int)I;
      if (!lines_pass((int)N1, FL_COL, true, g, N, true, 0, "4step-A")) { err = B200FFT_INTERNAL_ERROR; return; }
    }
    {  // pass B: FFT over n2 for each (o,k1); out row index k1 + N1*k2
      Geom g{};
      g.nb = (int)O; g.no = (int)N1; g.nl = (int)I;
      g.ibs = N * I; g.ios = N2 * I; g.ils = 1; g.ins = I;
      g.obs = N * I; g.oos = I; g.ols = 1; g.ons = N1 * I;
      g.tw_div = 1;
      if (!lines_pass((int)N2, FL_COL, false, g, 0, false, 0, "4step-B")) { err = B200FFT_INTERNAL_ERROR; return; }
    }
  }

  // four-step for contiguous lines: [O][N], 2 or 3 factors
  void fourstep_contig(long long O, long long N) {
    const long long lim = 1024;  // column / transposing tiles of <= 1024 points keep >= 64 B contiguous runs
    int lg = ilog2(N);
    if (N <= lim * lim) {
      if (try_fused_contig2(O, N)) return;
      long long N1 = 1LL << (lg / 2), N2 = N / N1;
      {  // A: FFT over n1 (stride N2), lines n2, twiddle w_N^(k1*n2)
        Geom g{};
        g.nb = 1; g.no = (int)O; g.nl = (int)N2;
        g.ios = N; g.ils = 1; g.ins = N2;
        g.oos = N; g.ols = 1; g.ons = N2;
        g.tw_div = 1;
        if (!lines_pass((int)N1, FL_COL, true, g, N, true, 0, "4step-A")) { err = B200FFT_INTERNAL_ERROR; return; }
      }
      {  // B: rows n2 for each k1, transposed store X[k1 + N1*k2]; lines = k1
        Geom g{};
        g.nb = (int)O; g.no = 1; g.nl = (int)N1;
        g.ibs = N; g.ils = N2; g.ins = 1;
        g.obs = N; g.ols = 1; g.ons = N1;
        g.tw_div = 1;
        if (!lines_pass((int)N2, FL_TRANS, false, g, 0, false, 0, "4step-B")) { err = B200FFT_INTERNAL_ERROR; return; }
      }
      return;
    }
    if (N > lim * lim * lim) { err = B200FFT_NOT_SUPPORTED; return; }
    if (try_band_contig2(O, N)) return;
    // three factors: n = n1*M + n2*N3 + n3 (M = N2*N3), k = k1 + N1*k2 + N1*N2*k3
    int a = lg / 3, b = (lg - a) / 2, c = lg - a - b;
    long long N1 = 1LL << c, N2 = 1LL << b, N3 = 1LL << a;  // largest first: column tiles are cheapest to keep big
    bool fuse_tail = false;
    for (int lm = 18; lm >= 16 && !fuse_tail; lm--) {      // prefer a tail that has a fused (L2-staged) kernel pair
      const long long n2 = 1LL << (lm / 2), n3 = 1LL << (lm - lm / 2), n1 = N >> lm;
      if (n1 >= 16 && n1 <= max_col_n() && find_fused(p->is_double, (int)n2, FL_COL, 1, (int)n3, FL_TRANS)) {
        N1 = n1; N2 = n2; N3 = n3; fuse_tail = true;
      }
    }
    if (!fuse_tail && !(getenv("B200FFT_PIPE") && atoi(getenv("B200FFT_PIPE")) == 0)) {
      // the middle pass has a short row stride (N3 elements): give it the length of the persistent pipelined column
      // kernel (~90 % of HBM peak against 65-80 % for the lock-step kernels of 512 / 1024 points)
      for (long long cand : {1024LL, 512LL}) {
        const long long n3 = 1LL << a, n1 = N / (cand * n3);
        if (!find_kernel(p->is_double, (int)cand, FL_PIPE, 1, 0) || n1 < 2 || n1 > 1024 || n1 * cand * n3 != N) continue;
        if (!find_kernel(p->is_double, (int)n1, FL_COL, 1, 0)) continue;
        N1 = n1; N2 = cand; N3 = n3;
        break;
      }
    }
    if (const char* sp = getenv("B200FFT_SPLIT3")) {   // developer override "n1,n2,n3"
      long long a1 = 0, a2 = 0, a3 = 0;
      if (sscanf(sp, "%lld,%lld,%lld", &a1, &a2, &a3) == 3 && a1 * a2 * a3 == N && is_pow2(a1) && is_pow2(a2) && is_pow2(a3)) {
        N1 = a1; N2 = a2; N3 = a3; fuse_tail = false;
      }
    }
    long long M = N2 * N3;
    if (O * N1 >= (1LL << 31)) { err = B200FFT_INVALID_SIZE; return; }
    {  // A: FFT over n1 (stride M), lines m < M, twiddle w_N^(k1*m)
      Geom g{};
      g.nb = 1; g.no = (int)O; g.nl = (int)M;
      g.ios = N; g.ils = 1; g.ins = M;
      g.oos = N; g.ols = 1; g.ons = M;
      g.tw_div = 1;
      if (!lines_pass((int)N1, FL_COL, true, g, N, true, 0, "6step-A")) { err = B200FFT_INTERNAL_ERROR; return; }
    }
    if (fuse_tail && try_fused_contig3_tail(O, N, N1, N2, N3)) return;
    {  // B: for each (o,k1): FFT over n2 (stride N3), lines n3, twiddle w_M^(k2*n3)
      Geom g{};
      g.nb = 1; g.no = (int)(O * N1); g.nl = (int)N3;
      g.ios = M; g.ils = 1; g.ins = N3;
      g.oos = M; g.ols = 1; g.ons = N3;
      g.tw_div = 1;
      if (!lines_pass((int)N2, FL_COL, true, g, M, true, 0, "6step-B")) { err = B200FFT_INTERNAL_ERROR; return; }
    }
    {  // C: rows n3 for each (k1,k2); lines = k1; X[k1 + N1*k2 + N1*N2*k3]
      Geom g{};
      g.nb = (int)O; g.no = (int)N2; g.nl = (int)N1;
      g.ibs = N; g.ios = N3; g.ils = M; g.ins = 1;
      g.obs = N; g.oos = N1; g.ols = 1; g.ons = N1 * N2;
      g.tw_div = 1;
      if (!lines_pass((int)N3, FL_TRANS, false, g, 0, false, 0, "6step-C")) { err = B200FFT_INTERNAL_ERROR; return; }
    }
  }

  // ---- 2D [H][W] with a column length H = 2*M the single-pass column kernel cannot hold: two HBM passes --------
  // The column axis is split H = 2 x M (n = n1*M + n2, k = k1 + 2*k2).  Pass 1 is the row pass over W with the
  // radix-2 butterfly over n1 folded into its load: the CTA for (n2, k1) reads rows n2 and n2 + M, transforms
  // x[n2] + (-1)^k1 x[n2+M] along W, multiplies by w_H^(k1*n2) (one constant per row) and stores row 2*n2 + k1.
  // Pass 2 is the M-point column transform over n2 of the rows of parity k1, in place (k2 lands on row k1 + 2*k2).
  bool try_pair_2d(long long H, long long W) {
    struct Mark { b200fft_plan_s* p; size_t n0; ~Mark() { if (p->passes.size() > n0) p->no_shift = true; } } mark{p, p->passes.size()};
    // Opt-in (B200FFT_PAIR2D=1).  Measured on B200 (profiles/r01_pair2d_and_narrow_columns.txt): correct, but the
    // column pass it needs -- 4096-point columns in tiles only 2-4 columns wide (16-32 B runs) -- reaches just
    // 2.5 TB/s, so cfg3 takes 693 us with it against 562 us for rows + two wide-tile column passes.
    if (!(getenv("B200FFT_PAIR2D") && atoi(getenv("B200FFT_PAIR2D")))) return false;
    if (!is_pow2(H) || !is_pow2(W) || H < 4) return false;
    const long long M = H / 2;
    if (H <= max_col_n() || M > env_int("B200FFT_PAIR_MAX_COL_N", 4096)) return false;
    if (!find_kernel(p->is_double, (int)W, FL_ROWPAIR, 0, 0) || !find_kernel(p->is_double, (int)M, FL_COL, 0, 0)) return false;
    if (H * W >= (1LL << 40) || M >= (1LL << 30)) return false;
    {
      Geom g{};
      g.nb = 1; g.no = (int)M; g.nl = 2;
      g.ios = W; g.ils = 0; g.ins = 1; g.pre2_off = M * W;
      g.oos = 2 * W; g.ols = W; g.ons = 1;
      g.tw_div = 1;
      const size_t before = p->passes.size();
      if (!lines_pass((int)W, FL_ROWPAIR, false, g, 0, false, 0, "pair-rows")) return false;
      Pass& ps = p->passes[before];
      make_fourstep_tables(p, H, &ps.g.tw_lo_bits, &ps.tw_lo, &ps.tw_hi);
    }
    {
      Geom g{};
      g.nb = 1; g.no = 2; g.nl = (int)W;
      g.ios = W; g.ils = 1; g.ins = 2 * W;
      g.oos = W; g.ols = 1; g.ons = 2 * W;
      g.tw_div = 1;
      if (!lines_pass((int)M, FL_COL, false, g, 0, true, 0, "pair-cols")) { err = B200FFT_INTERNAL_ERROR; return true; }
    }
    return true;
  }

  // ---- 2D [H][W], H = CS*M: rows + the first radix-CS stage of the column axis in one cluster pass, then ONE M-point
  //      column pass over consecutive rows (cluster_kernel.cuh: fft_cluster_rows_kernel) ------------------------------
  bool try_cluster_rows_2d(long long H, long long W) {
    struct Mark { b200fft_plan_s* p; size_t n0; ~Mark() { if (p->passes.size() > n0) p->no_shift = true; } } mark{p, p->passes.size()};
    const char* mode = getenv("B200FFT_CLUSTER_ROWS");
    if (!(mode && atoi(mode))) return false;
    if (!is_pow2(H) || !is_pow2(W) || H <= max_col_n()) return false;
    const KernelEntry* k = find_kernel(p->is_double, (int)W, FL_CLUSTERROW, 0, 0);
    if (!k || H % k->CS) return false;
    const long long M = H / k->CS;
    if (M > max_col_n() || M < 2 || !find_kernel(p->is_double, (int)M, FL_COL, 0, (int)W)) return false;
    if (H * W >= (1LL << 40)) return false;
    {
      Pass ps;
      ps.kind = PK_CLUSTER;
      ps.k = k;
      Geom g{};
      g.nb = 1; g.no = 1; g.nl = (int)M; g.ntl = (int)M;
      g.ios = M * W; g.ils = W; g.ins = 1;
      g.oos = M * W; g.ols = W; g.ons = 1;
      g.tw_div = 1;
      ps.g = g;
      ps.inplace_ok = true;
      ps.ntiles = M;
      ps.tws = make_stage_twiddles(p, k);
      auto build = [&](auto tag) -> void* {   // w_H^(k1 * n2), k1 = 1 .. CS-1, n2 < M
        using T = decltype(tag);
        std::vector<T> h(2 * (size_t)(k->CS - 1) * M);
        for (int k1 = 1; k1 < k->CS; k1++)
          for (long long n2 = 0; n2 < M; n2++) fill_root_table(h, (size_t)(k1 - 1) * M + n2, (long long)k1 * n2, H);
        return upload(p, h);
      };
      ps.ctw = p->is_double ? build(double{}) : build(float{});
      char buf[256];
      snprintf(buf, sizeof buf, "rows+radix%d: cluster rows N=%d E=%d radix=%dx%dx%dx%d | %d CTAs x threads=%d smem=%zu clusters=%lld", k->CS,
               k->N, k->E, k->rad[0], k->rad[1], k->rad[2], k->rad[3], k->CS, k->threads, k->smem, (long long)M);
      ps.desc = buf;
      push(ps);
    }
    {  // the M-point transforms over n2 for each k1 (consecutive rows), output row k1 + CS*k2
      Geom g{};
      g.nb = 1; g.no = k->CS; g.nl = (int)W;
      g.ios = M * W; g.ils = 1; g.ins = W;
      g.oos = W; g.ols = 1; g.ons = (long long)k->CS * W;
      g.tw_div = 1;
      if (!lines_pass((int)M, FL_COL, false, g, 0, false, 0, "cols-tail")) { err = B200FFT_INTERNAL_ERROR; return true; }
      p->passes.back().axis_last = true;
    }
    return true;
  }

  // ---- arbitrary (non power-of-two) axis lengths: generic_kernel.cu -----------------------
  void generic_axis(long long O, long long N, long long I) {
    Pass ps;
    int e = plan_generic_axis(p->is_double, O, N, I, &ps.gp, [&](const void* h, size_t bytes) -> void* {
      void* d = nullptr;
      if (cudaMalloc(&d, bytes) != cudaSuccess) { cudaGetLastError(); p->upload_failed = true; return nullptr; }
      p->dev_allocs.push_back(d);
      if (cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); p->upload_failed = true; }
      return d;
    });
    if (e) { err = e; return; }
    ps.kind = ps.gp.bluestein ? PK_BLUESTEIN : PK_GENERIC;
    ps.inplace_ok = true;
    ps.desc = ps.gp.desc;
    if (ps.gp.workspace_bytes > p->extra_bytes) p->extra_bytes = ps.gp.workspace_bytes;
    push(ps);
  }

  // ---- assign buffers: the input is never written, the last moving pass lands in `out` ----
  void route() {
    auto& P = p->passes;
    if (P.empty()) {  // all extents 1: the transform is the identity
      Pass ps; ps.kind = PK_COPY; ps.desc = "copy (identity transform)";
      P.push_back(ps);
    }
    int q = 0;
    for (int i = (int)P.size() - 1; i > 0; i--)
      if (!P[i].inplace_ok) { q = i; break; }
    // passes after q run in place on out; walk backwards alternating out/scratch at moving passes
    int cur = BUF_OUT;
    for (int i = (int)P.size() - 1; i >= 0; i--) {
      P[i].dst = cur;
      if (i == 0) { P[i].src = BUF_IN; break; }
      if (i > q || P[i].inplace_ok) { P[i].src = cur; }
      else { cur = (cur == BUF_OUT) ? BUF_SCRATCH : BUF_OUT; P[i].src = cur; }
    }
    bool uses_scratch = false;
    for (auto& ps : P) uses_scratch |= (ps.src == BUF_SCRATCH || ps.dst == BUF_SCRATCH);
    p->scratch_bytes = uses_scratch ? (size_t)p->total * esize(p) : 0;
    // band scratch: the slots (shared by the passes, which run one after the other) then every pass's own counters, so that
    // all counters can be cleared up front and no memset sits between two kernels of an exec
    size_t slots = 0, counters = 0;
    for (auto& ps : P) {
      if (ps.slot_bytes > slots) slots = ps.slot_bytes;
      ps.counter_off = counters;
      counters += ps.counter_bytes;
    }
    slots = (slots + 255) / 256 * 256;
    p->band_slot_bytes = slots;
    p->band_bytes = counters ? slots + counters : 0;
  }
};

static int set_func_attrs(b200fft_plan_s* p) {
  for (auto& ps : p->passes) {
    if (ps.kind == PK_LINES && ps.k->smem > 48 * 1024) {
      if (cudaFuncSetAttribute(ps.k->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.k->smem) != cudaSuccess)
        return B200FFT_INTERNAL_ERROR;
    }
    if (ps.kind == PK_CLUSTER) {
      if (ps.k->smem > 48 * 1024 &&
          cudaFuncSetAttribute(ps.k->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.k->smem) != cudaSuccess)
        return B200FFT_INTERNAL_ERROR;
      if (ps.k->CS > 8 && cudaFuncSetAttribute(ps.k->func, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
        return B200FFT_INTERNAL_ERROR;
      // (the default shared-memory carveout is kept: asking for the maximum shrinks the L1 that stages the in-flight
      //  global loads -- measured 377 -> 449 us on the 8192-point c64 column pass)
    }
    if (ps.pipe) {
      const KernelEntry* q = ps.pipe;
      int dev = 0, nsm = 0, nclus = 0;
      bool ok = cudaFuncSetAttribute(q->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q->smem) == cudaSuccess &&
                (q->CS <= 8 || cudaFuncSetAttribute(q->func, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) &&
                cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess;
      if (ok) {
        if (q->CS > 1) {
          cudaLaunchConfig_t cfg{};
          cfg.gridDim = dim3((unsigned)(nsm / q->CS * q->CS));
          cfg.blockDim = dim3(q->threads);
          cfg.dynamicSmemBytes = q->smem;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = (unsigned)q->CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          ok = cudaOccupancyMaxActiveClusters(&nclus, q->func, &cfg) == cudaSuccess && nclus > 0;
        } else {
          int occ = 0;
          ok = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, q->func, q->threads, q->smem) == cudaSuccess && occ > 0;
          nclus = nsm * occ;
        }
      }
      if (!ok) { cudaGetLastError(); ps.pipe = nullptr; }   // the plain kernel still serves the pass
      else {
        const long long ntiles = (long long)ps.g.nb * ps.g.no * ps.pipe_ntl;
        ps.pipe_grid = (int)((nclus < ntiles ? nclus : ntiles) * q->CS);
      }
    }
    if (ps.kind == PK_FUSED2) {
      if (ps.fz->smem > 48 * 1024 &&
          cudaFuncSetAttribute(ps.fz->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.fz->smem) != cudaSuccess)
        return B200FFT_INTERNAL_ERROR;
      int dev = 0, nsm = 0, occ = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ps.fz->func, ps.fz->threads, ps.fz->smem) != cudaSuccess || occ < 1)
        return B200FFT_INTERNAL_ERROR;
      long long grid = (long long)nsm * occ;
      ps.fused_grid = (int)(grid < ps.ntiles ? grid : ps.ntiles);
    }
    if (ps.kind == PK_BAND) {
      int dev = 0, nsm = 0, occ = 0;
      if (cudaFuncSetAttribute(ps.bz->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.bz->smem) != cudaSuccess ||
          cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ps.bz->func, ps.bz->threads, ps.bz->smem) != cudaSuccess || occ < 1)
        return B200FFT_INTERNAL_ERROR;
      const long long grid = (long long)nsm * occ;
      ps.band_grid = (int)(grid < ps.ntiles ? grid : ps.ntiles);
    }
    if (ps.kind == PK_LINES && ps.ringcol) {
      int dev = 0, nsm = 0, occ = 0;
      bool ok = cudaFuncSetAttribute(ps.ringcol->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.ringcol->smem) == cudaSuccess &&
                cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ps.ringcol->func, ps.ringcol->threads, ps.ringcol->smem) == cudaSuccess && occ > 0;
      if (!ok) { cudaGetLastError(); ps.ringcol = nullptr; }   // the lock-step kernel still serves the pass
      else {
        const long long ntiles = (long long)ps.g.no * ps.ringcol_ntl, grid = (long long)nsm * occ;
        ps.ringcol_grid = (int)(grid < ntiles ? grid : ntiles);
      }
    }
    if (ps.kind == PK_LINES && ps.ringtrans) {
      int dev = 0, nsm = 0, occ = 0;
      bool ok = cudaFuncSetAttribute(ps.ringtrans->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.ringtrans->smem) == cudaSuccess &&
                cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ps.ringtrans->func, ps.ringtrans->threads, ps.ringtrans->smem) == cudaSuccess && occ > 0;
      if (!ok) { cudaGetLastError(); ps.ringtrans = nullptr; }
      else {
        const long long ntiles = (long long)ps.g.nb * ps.g.no * ps.ringtrans_ntl, grid = (long long)nsm * occ;
        ps.ringtrans_grid = (int)(grid < ntiles ? grid : ntiles);
      }
    }
    if (ps.kind == PK_LINES && ps.ring) {
      int dev = 0, nsm = 0, occ = 0;
      bool ok = cudaFuncSetAttribute(ps.ring->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.ring->smem) == cudaSuccess &&
                cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ps.ring->func, ps.ring->threads, ps.ring->smem) == cudaSuccess && occ > 0;
      if (!ok) { cudaGetLastError(); ps.ring = nullptr; continue; }   // the plain kernel still serves the pass
      long long grid = (long long)nsm * occ;
      ps.ring_grid = (int)(grid < ps.ring_ntl ? grid : ps.ring_ntl);
    }
  }
  return generic_set_attrs();
}

static int finish_plan(b200fft_plan_s* p, Builder& b, b200fftHandle* out) {
  if (!b.err) b.route();
  if (!b.err && p->upload_failed) b.err = B200FFT_ALLOC_FAILED;
  if (!b.err) b.err = set_func_attrs(p);
  if (b.err) {
    for (auto& ps : p->passes) destroy_generic(&ps.gp);
    for (void* d : p->dev_allocs) cudaFree(d);
    int e = b.err;
    delete p;
    cudaGetLastError();
    return e;
  }
  *out = p;
  return B200FFT_SUCCESS;
}

// Create a plan with `init` (fills type / rank / extents) and `build` (pushes the passes); when the result holds a
// band pass, the same transform is planned again without band passes as its fallback (see b200fft_plan_s::fallback).
template <class Init, class Build>
static int create_plan(b200fftHandle* out, Init&& init, Build&& build) {
  auto mk = [&](bool no_band, b200fftHandle* h) {
    auto* p = new b200fft_plan_s;
    init(p);
    Builder b{p};
    b.no_band = no_band;
    build(b);
    return finish_plan(p, b, h);
  };
  b200fftHandle h = nullptr;
  if (int e = mk(false, &h)) return e;
  bool has_band = false;
  for (const auto& ps : h->passes) has_band |= ps.kind == PK_BAND;
  if (has_band) {
    b200fftHandle fb = nullptr;
    if (int e = mk(true, &fb)) { b200fftDestroy(h); return e; }
    h->fallback = fb;
  }
  *out = h;
  return B200FFT_SUCCESS;
}

static int check_type(int type, int* is_double) {
  if (type == B200FFT_C2C) { *is_double = 0; return 0; }
  if (type == B200FFT_Z2Z) { *is_double = 1; return 0; }
  return B200FFT_INVALID_TYPE;
}

static int have_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return B200FFT_NO_DEVICE; }
  return 0;
}

}  // namespace b200fft

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int b200fftPlanMany1d(b200fftHandle* plan, int64_t n, int64_t batch, int type) {
  if (!plan) return B200FFT_INVALID_VALUE;
  int dbl;
  if (check_type(type, &dbl)) return B200FFT_INVALID_TYPE;
  if (n < 1 || batch < 1) return B200FFT_INVALID_SIZE;
  if (int e = have_device()) return e;
  return create_plan(plan, [&](b200fft_plan_s* p) { p->is_double = dbl; p->rank = 1; p->dims[0] = n; p->batch = batch; p->total = n * batch; },
                     [&](Builder& b) { b.axis(batch, n, 1); });
}

int b200fftPlanAxis(b200fftHandle* plan, int64_t outer, int64_t n, int64_t inner, int type) {
  if (!plan) return B200FFT_INVALID_VALUE;
  int dbl;
  if (check_type(type, &dbl)) return B200FFT_INVALID_TYPE;
  if (n < 1 || outer < 1 || inner < 1) return B200FFT_INVALID_SIZE;
  if (int e = have_device()) return e;
  return create_plan(plan, [&](b200fft_plan_s* p) { p->is_double = dbl; p->rank = 1; p->dims[0] = n; p->batch = outer * inner; p->total = n * outer * inner; },
                     [&](Builder& b) { b.axis(outer, n, inner); });
}

int b200fftPlanAxisView(b200fftHandle* plan, int64_t outer, int64_t n, int64_t inner, int64_t outer_stride, int64_t n_stride, int type) {
  if (!plan) return B200FFT_INVALID_VALUE;
  int dbl;
  if (check_type(type, &dbl)) return B200FFT_INVALID_TYPE;
  if (n < 1 || outer < 1 || inner < 1 || n_stride < inner || outer_stride < 0) return B200FFT_INVALID_SIZE;
  if (int e = have_device()) return e;
  return create_plan(plan, [&](b200fft_plan_s* p) { p->is_double = dbl; p->rank = 1; p->dims[0] = n; p->batch = outer * inner; p->total = n * outer * inner; },
                     [&](Builder& b) { b.axis_view(outer, n, inner, outer_stride, n_stride); });
}

int b200fftPlan1d(b200fftHandle* plan, int64_t n, int type, int64_t batch) { return b200fftPlanMany1d(plan, n, batch, type); }

int b200fftPlan2d(b200fftHandle* plan, int64_t h, int64_t w, int type) {
  if (!plan) return B200FFT_INVALID_VALUE;
  int dbl;
  if (check_type(type, &dbl)) return B200FFT_INVALID_TYPE;
  if (h < 1 || w < 1) return B200FFT_INVALID_SIZE;
  if (int e = have_device()) return e;
  return create_plan(plan, [&](b200fft_plan_s* p) { p->is_double = dbl; p->rank = 2; p->dims[0] = h; p->dims[1] = w; p->total = h * w; },
                     [&](Builder& b) {
                       if (!b.try_cluster_rows_2d(h, w) && !b.try_pair_2d(h, w)) {
                         b.axis(h, w, 1);   // rows
                         b.axis(1, h, w);   // columns
                       }
                     });
}

int b200fftPlan3d(b200fftHandle* plan, int64_t d, int64_t h, int64_t w, int type) {
  if (!plan) return B200FFT_INVALID_VALUE;
  int dbl;
  if (check_type(type, &dbl)) return B200FFT_INVALID_TYPE;
  if (d < 1 || h < 1 || w < 1) return B200FFT_INVALID_SIZE;
  if (int e = have_device()) return e;
  return create_plan(plan, [&](b200fft_plan_s* p) { p->is_double = dbl; p->rank = 3; p->dims[0] = d; p->dims[1] = h; p->dims[2] = w; p->total = d * h * w; },
                     [&](Builder& b) {
                       if (!b.try_fused_plane(d, h, w)) {
                         b.axis(d * h, w, 1);   // x
                         b.axis(d, h, w);       // y
                       }
                       b.axis(1, d, h * w);   // z
                     });
}

static int exec_common(b200fftHandle p, const void* in, void* out, int direction, double scale, bool shifted, b200fftStream stream_);

// cuTensorMapEncodeTiled through the runtime's driver entry point query: no link-time dependency on libcuda.so
typedef CUresult (*tensor_map_encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tensor_map_encode_t tensor_map_encoder() {
  static tensor_map_encode_t fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      f = nullptr;
    }
    return (tensor_map_encode_t)f;
  }();
  return fn;
}

// Launch the persistent single-buffer TMA-fed column kernel for a lines pass (ringcol_kernel.cuh); false = not launched (the
// caller falls back to the lock-step kernel).  g carries swap / peer fields already; max_ctas > 0 limits the grid.
static bool launch_ringcol(const b200fft_plan_s* p, const Pass& ps, Geom g, const void* src, void* dst, double sc, int max_ctas,
                           cudaStream_t stream) {
  const KernelEntry* q = ps.ringcol;
  if (!q || !tensor_map_encoder() || ((uintptr_t)src & 15)) return false;
  const unsigned long long esz = p->is_double ? 16 : 8;
  alignas(64) CUtensorMap tm;
  // the input as a tensor {2 I (reals of a row piece), N rows, O}; a tile = boxes of {2 TL, rows per box, 1}
  cuuint64_t dims[3] = {2ull * (unsigned long long)g.nl, (unsigned long long)q->N, (unsigned long long)g.no};
  cuuint64_t strides[2] = {(unsigned long long)g.ins * esz, (unsigned long long)(g.no > 1 ? g.ios : (long long)q->N * g.ins) * esz};
  cuuint32_t box[3] = {2u * (unsigned)q->TL, (unsigned)q->N1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if (tensor_map_encoder()(&tm, p->is_double ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(src), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  g.ntl = ps.ringcol_ntl;
  float scf = (float)sc;
  double scd = sc;
  void* args[] = {(void*)&tm, &g, (void*)&dst, (void*)&ps.rctws, (void*)&ps.tw_lo, (void*)&ps.tw_hi, p->is_double ? (void*)&scd : (void*)&scf};
  static const bool pdl = !(getenv("B200FFT_PDL") && atoi(getenv("B200FFT_PDL")) == 0);
  cudaLaunchConfig_t cfg{};
  const int grid = (max_ctas > 0 && max_ctas < ps.ringcol_grid) ? max_ctas : ps.ringcol_grid;
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(q->threads);
  cfg.dynamicSmemBytes = q->smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  if (cudaLaunchKernelExC(&cfg, q->func, args) != cudaSuccess) { cudaGetLastError(); return false; }
  return true;
}

int b200fftExecScaled(b200fftHandle p, const void* in, void* out, int direction, double scale, b200fftStream stream_) {
  return exec_common(p, in, out, direction, scale, false, stream_);
}

// The transform followed by DFT/Centre.hs's shift1D/2D/3D along every transformed axis (zero frequency in the
// middle), the rotation folded into the stores of each axis' last butterfly pass -- for even extents this is also
// fft(centre(x)) (Centre.hs:17-19).  B200FFT_NOT_SUPPORTED when an axis' last pass is not a power-of-two line kernel
// (odd / non-power-of-two extents): the caller then runs the stand-alone shift (accfft_fft_centred does).
int b200fftExecShifted(b200fftHandle p, const void* in, void* out, int direction, double scale, b200fftStream stream_) {
  if (!p || p->magic != 0xB200FF7u) return B200FFT_INVALID_PLAN;
  if (p->fallback) return b200fftExecShifted(p->fallback, in, out, direction, scale, stream_);
  if (p->no_shift) return B200FFT_NOT_SUPPORTED;
  for (const Pass& ps : p->passes)
    if (ps.axis_last && (ps.kind != PK_LINES || ps.k->N < 2 || (ps.k->N & 1))) return B200FFT_NOT_SUPPORTED;
  return exec_common(p, in, out, direction, scale, true, stream_);
}

static int exec_common(b200fftHandle p, const void* in, void* out, int direction, double scale, bool shifted, b200fftStream stream_) {
  if (!p || p->magic != 0xB200FF7u) return B200FFT_INVALID_PLAN;
  if (!in || !out || in == out) return B200FFT_INVALID_VALUE;
  if (direction != B200FFT_FORWARD && direction != B200FFT_INVERSE) return B200FFT_INVALID_VALUE;
  if (p->fallback && ((((uintptr_t)in | (uintptr_t)out) & 15) != 0 || !tensor_map_encoder()))
    return exec_common(p->fallback, in, out, direction, scale, shifted, stream_);
  cudaStream_t stream = (cudaStream_t)stream_;
  void* scratch = nullptr;
  void* extra = nullptr;
  if (p->scratch_bytes && scratch_alloc(&scratch, p->scratch_bytes, stream) != cudaSuccess) { cudaGetLastError(); return B200FFT_ALLOC_FAILED; }
  if (p->extra_bytes && scratch_alloc(&extra, p->extra_bytes, stream) != cudaSuccess) {
    cudaGetLastError();
    if (scratch) cudaFreeAsync(scratch, stream);
    return B200FFT_ALLOC_FAILED;
  }
  void* band = nullptr;
  if (p->band_bytes && scratch_alloc(&band, p->band_bytes, stream) != cudaSuccess) {
    cudaGetLastError();
    if (scratch) cudaFreeAsync(scratch, stream);
    if (extra) cudaFreeAsync(extra, stream);
    return B200FFT_ALLOC_FAILED;
  }
  const int inverse = direction == B200FFT_INVERSE;
  int status = B200FFT_SUCCESS;
  const size_t np = p->passes.size();
  if (band && cudaMemsetAsync((char*)band + p->band_slot_bytes, 0, p->band_bytes - p->band_slot_bytes, stream) != cudaSuccess)
    status = B200FFT_EXEC_FAILED;
  for (size_t i = 0; i < np && status == B200FFT_SUCCESS; i++) {
    const Pass& ps = p->passes[i];
    const void* src = ps.src == BUF_IN ? in : ps.src == BUF_OUT ? out : scratch;
    void* dst = ps.dst == BUF_OUT ? out : scratch;
    const bool first = i == 0, last = i + 1 == np;
    const double sc = last ? scale : 1.0;
    cudaError_t ce = cudaSuccess;
    const bool rotate = shifted && ps.axis_last;
    bool pipe_done = false;
    if (!rotate && ps.pipe && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
      Geom g = ps.g;
      g.swap_in = inverse && first;
      g.swap_out = inverse && last;
      g.ntl = ps.pipe_ntl;
      float scf = (float)sc;
      double scd = sc;
      void* args[] = {&g, (void*)&src, (void*)&dst, (void*)&ps.ptws, (void*)&ps.tw_lo, (void*)&ps.tw_hi,
                      p->is_double ? (void*)&scd : (void*)&scf, (void*)&ps.pctw};
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)ps.pipe_grid);
      cfg.blockDim = dim3(ps.pipe->threads);
      cfg.dynamicSmemBytes = ps.pipe->smem;
      cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)ps.pipe->CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = ps.pipe->CS > 1 ? 1 : 0;
      if (cudaLaunchKernelExC(&cfg, ps.pipe->func, args) == cudaSuccess) {
        pipe_done = true;
        g_launches.fetch_add(1, std::memory_order_relaxed);
      } else {
        cudaGetLastError();   // e.g. cluster launches refused under a partitioned device: the lock-step kernel serves the pass
      }
    }
    if (!pipe_done && !rotate && ps.kind == PK_LINES && ps.ringcol) {
      Geom g = ps.g;
      g.swap_in = inverse && first;
      g.swap_out = inverse && last;
      if (launch_ringcol(p, ps, g, src, dst, sc, 0, stream)) {
        pipe_done = true;
        g_launches.fetch_add(1, std::memory_order_relaxed);
      }
    }
    if (!pipe_done && !rotate && ps.kind == PK_LINES && ps.ringtrans && ((uintptr_t)src & 15) == 0) {
      Geom g = ps.g;
      g.swap_in = inverse && first;
      g.swap_out = inverse && last;
      g.ntl = ps.ringtrans_ntl;
      float scf = (float)sc;
      double scd = sc;
      void* args[] = {&g, (void*)&src, (void*)&dst, (void*)&ps.rttws, p->is_double ? (void*)&scd : (void*)&scf};
      static const bool pdl = !(getenv("B200FFT_PDL") && atoi(getenv("B200FFT_PDL")) == 0);
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)ps.ringtrans_grid);
      cfg.blockDim = dim3(ps.ringtrans->threads);
      cfg.dynamicSmemBytes = ps.ringtrans->smem;
      cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
      if (cudaLaunchKernelExC(&cfg, ps.ringtrans->func, args) == cudaSuccess) {
        pipe_done = true;
        g_launches.fetch_add(1, std::memory_order_relaxed);
      } else {
        cudaGetLastError();
      }
    }
    if (pipe_done) {
      // the pass is on its way
    } else if (ps.kind == PK_LINES) {
      Geom g = ps.g;
      g.swap_in = inverse && first;
      g.swap_out = inverse && last;
      float scf = (float)sc;
      double scd = sc;
      void* args[] = {&g, (void*)&src, (void*)&dst, (void*)&ps.tws, (void*)&ps.tw_lo, (void*)&ps.tw_hi,
                      p->is_double ? (void*)&scd : (void*)&scf};
      if (rotate) {   // output index k of this pass lands at (k + N/2) mod N: two "peers" = the two halves of the axis
        const size_t esz = p->is_double ? 16 : 8;
        g.npeers = 2;
        g.peer_shift = ilog2(ps.k->N / 2);
        g.peer[0] = (char*)dst + (size_t)(ps.k->N / 2) * (size_t)g.ons * esz;
        g.peer[1] = dst;
      }
      // Programmatic dependent launch: the kernel may be scheduled while its predecessor on the stream drains (its CTAs
      // take the SMs the predecessor's last wave leaves idle and sit in griddepcontrol.wait until it has completed and
      // flushed), so back-to-back transforms lose the launch latency and the scheduling ramp between them.  The lines and
      // ring kernels touch no global memory before that wait, so stream order is what it always was.
      static const bool pdl = !(getenv("B200FFT_PDL") && atoi(getenv("B200FFT_PDL")) == 0);
      cudaLaunchConfig_t cfg{};
      cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
      if (!rotate && ps.ring && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
        g.ntl = ps.ring_ntl;
        cfg.gridDim = dim3((unsigned)ps.ring_grid);
        cfg.blockDim = dim3(ps.ring->threads);
        cfg.dynamicSmemBytes = ps.ring->smem;
        ce = cudaLaunchKernelExC(&cfg, ps.ring->func, args);
      } else {
        cfg.gridDim = dim3((unsigned)ps.ntiles);
        cfg.blockDim = dim3(ps.k->threads);
        cfg.dynamicSmemBytes = ps.k->smem;
        ce = cudaLaunchKernelExC(&cfg, ps.k->func, args);
      }
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else if (ps.kind == PK_CLUSTER) {
      Geom g = ps.g;
      g.swap_in = inverse && first;
      g.swap_out = inverse && last;
      float scf = (float)sc;
      double scd = sc;
      void* args[] = {&g, (void*)&src, (void*)&dst, (void*)&ps.tws, (void*)&ps.tw_lo, (void*)&ps.tw_hi,
                      p->is_double ? (void*)&scd : (void*)&scf, (void*)&ps.ctw};
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(ps.ntiles * ps.k->CS));
      cfg.blockDim = dim3(ps.k->threads);
      cfg.dynamicSmemBytes = ps.k->smem;
      cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)ps.k->CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      ce = cudaLaunchKernelExC(&cfg, ps.k->func, args);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else if (ps.kind == PK_FUSED2) {
      FusedParams fp = ps.fp;
      fp.a.swap_in = inverse && first;
      fp.b.swap_out = inverse && last;
      unsigned* counters = (unsigned*)((char*)band + p->band_slot_bytes + ps.counter_off);
      void* mid = ps.mid_in_dst ? dst : band;
      float scf = (float)sc;
      double scd = sc;
      void* args[] = {&fp, (void*)&src, (void*)&dst, (void*)&mid, (void*)&ps.tws, (void*)&ps.twsB, (void*)&ps.tw_lo, (void*)&ps.tw_hi,
                      p->is_double ? (void*)&scd : (void*)&scf, (void*)&counters};
      if (ce == cudaSuccess)
        ce = cudaLaunchCooperativeKernel(ps.fz->func, dim3((unsigned)ps.fused_grid), dim3(ps.fz->threads), args, ps.fz->smem, stream);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else if (ps.kind == PK_BAND) {
      BandParams bp = ps.bp;
      bp.swap_in = inverse && first;
      bp.swap_out = inverse && last;
      unsigned* counters = (unsigned*)((char*)band + p->band_slot_bytes + ps.counter_off);
      void* slots = band;
      alignas(64) CUtensorMap tm;
      cuuint64_t dims[4] = {ps.tm_dims[0], ps.tm_dims[1], ps.tm_dims[2], ps.tm_dims[3]};
      cuuint64_t strides[3] = {ps.tm_strides[0], ps.tm_strides[1], ps.tm_strides[2]};
      cuuint32_t box[4] = {ps.tm_box[0], ps.tm_box[1], ps.tm_box[2], ps.tm_box[3]};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      const CUresult cr = tensor_map_encoder()(&tm, p->is_double ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                                               const_cast<void*>(src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) { status = B200FFT_EXEC_FAILED; break; }
      float scf = (float)sc;
      double scd = sc;
      void* args[] = {(void*)&tm, (void*)&bp, (void*)&dst, (void*)&slots, (void*)&ps.tws, (void*)&ps.twsB, (void*)&ps.tw_lo, (void*)&ps.tw_hi,
                      (void*)&ps.otw_lo, (void*)&ps.otw_hi, p->is_double ? (void*)&scd : (void*)&scf, (void*)&counters};
      if (ce == cudaSuccess)
        ce = cudaLaunchCooperativeKernel(ps.bz->func, dim3((unsigned)ps.band_grid), dim3(ps.bz->threads), args, ps.bz->smem, stream);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else if (ps.kind == PK_COPY) {
      long long nl = 0;
      ce = launch_copy_scale(p->is_double, src, dst, p->total, sc, stream, &nl);
      g_launches.fetch_add(nl, std::memory_order_relaxed);
    } else {
      long long nl = 0;
      ce = launch_generic(p->is_double, ps.gp, src, dst, extra, (inverse && first ? 1 : 0) | (inverse && last ? 2 : 0), sc, stream, &nl);
      g_launches.fetch_add(nl, std::memory_order_relaxed);
    }
    if (ce != cudaSuccess) status = B200FFT_EXEC_FAILED;
  }
  if (scratch) cudaFreeAsync(scratch, stream);
  if (extra) cudaFreeAsync(extra, stream);
  if (band) cudaFreeAsync(band, stream);
  if (status != B200FFT_SUCCESS) cudaGetLastError();
  return status;
}

int b200fftExec(b200fftHandle plan, const void* in, void* out, int direction, b200fftStream stream) {
  return b200fftExecScaled(plan, in, out, direction, 1.0, stream);
}

// One strided-axis pass whose stores are scattered over peer buffers: output index n of the transformed axis goes
// to outs[n / (N/npeers)] at local index n % (N/npeers); within a buffer the element of (outer o, index nl, inner i)
// sits at o*out_outer_stride + nl*out_n_stride + i.  See include/b200fft.h.
int b200fftExecScatter(b200fftHandle p, const void* in, void* const* outs, int npeers, int64_t out_outer_stride,
                       int64_t out_n_stride, int direction, double scale, b200fftStream stream_) {
  return b200fftExecScatterOn(p, in, outs, npeers, out_outer_stride, out_n_stride, direction, scale, 0, stream_);
}
// max_ctas > 0: the pass runs as a grid-stride loop of at most that many CTAs (it then occupies that many SMs and leaves the
// others to kernels running beside it on other streams); 0 = one CTA per tile.
int b200fftExecScatterOn(b200fftHandle p, const void* in, void* const* outs, int npeers, int64_t out_outer_stride,
                         int64_t out_n_stride, int direction, double scale, int max_ctas, b200fftStream stream_) {
  if (!p || p->magic != 0xB200FF7u) return B200FFT_INVALID_PLAN;
  if (!in || !outs || npeers < 1 || npeers > 16) return B200FFT_INVALID_VALUE;
  if (direction != B200FFT_FORWARD && direction != B200FFT_INVERSE) return B200FFT_INVALID_VALUE;
  if (p->passes.size() != 1) return B200FFT_NOT_SUPPORTED;
  const Pass& ps = p->passes[0];
  if (ps.kind != PK_LINES || ps.k->flavor != FL_COL || ps.k->tw4) return B200FFT_NOT_SUPPORTED;
  const long long N = ps.k->N;
  if (N % npeers || !is_pow2(N / npeers)) return B200FFT_INVALID_SIZE;
  for (int i = 0; i < npeers; i++)
    if (!outs[i] || outs[i] == in) return B200FFT_INVALID_VALUE;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int inverse = direction == B200FFT_INVERSE;
  Geom g = ps.g;
  g.swap_in = inverse;
  g.swap_out = inverse;
  g.obs = 0; g.oos = out_outer_stride; g.ons = out_n_stride;
  g.npeers = npeers;
  g.peer_shift = ilog2(N / npeers);
  for (int i = 0; i < npeers; i++) g.peer[i] = outs[i];
  float scf = (float)scale;
  double scd = scale;
  void* dst = outs[0];
  cudaError_t ce;
  // NVLink wants long runs: the lock-step kernel stores 128 B per row (TL = 16), the pipelined one 64 B -- measured on
  // 2 x B200, 1024^3: 6.24 ms against 8.30 ms per transform -- so the pipelined kernel only serves a single target
  if (launch_ringcol(p, ps, g, in, dst, scale, max_ctas, stream)) {
    ce = cudaSuccess;
  } else if (max_ctas > 0 && ps.k->loop_func) {
    if (ps.k->smem > 48 * 1024) cudaFuncSetAttribute(ps.k->loop_func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.k->smem);
    unsigned nt = (unsigned)ps.ntiles;
    void* args[] = {&g, (void*)&in, (void*)&dst, (void*)&ps.tws, (void*)&ps.tw_lo, (void*)&ps.tw_hi,
                    p->is_double ? (void*)&scd : (void*)&scf, (void*)&nt};
    const long long grid = ps.ntiles < max_ctas ? ps.ntiles : max_ctas;
    ce = cudaLaunchKernel(ps.k->loop_func, dim3((unsigned)grid), dim3(ps.k->threads), args, ps.k->smem, stream);
  } else if (npeers == 1 && ps.pipe && ps.pipe->CS == 1 && ((uintptr_t)in & 15) == 0) {
    g.ntl = ps.pipe_ntl;
    void* args[] = {&g, (void*)&in, (void*)&dst, (void*)&ps.ptws, (void*)&ps.tw_lo, (void*)&ps.tw_hi,
                    p->is_double ? (void*)&scd : (void*)&scf, (void*)&ps.pctw};
    ce = cudaLaunchKernel(ps.pipe->func, dim3((unsigned)ps.pipe_grid), dim3(ps.pipe->threads), args, ps.pipe->smem, stream);
  } else {
    void* args[] = {&g, (void*)&in, (void*)&dst, (void*)&ps.tws, (void*)&ps.tw_lo, (void*)&ps.tw_hi,
                    p->is_double ? (void*)&scd : (void*)&scf};
    ce = cudaLaunchKernel(ps.k->func, dim3((unsigned)ps.ntiles), dim3(ps.k->threads), args, ps.k->smem, stream);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (ce != cudaSuccess) { cudaGetLastError(); return B200FFT_EXEC_FAILED; }
  return B200FFT_SUCCESS;
}

// Device buffers that other processes of the box can map (CUDA IPC): the peer buffers of b200fftExecScatter.
int b200fftPeerAlloc(void** ptr, size_t bytes) {
  if (!ptr || !bytes) return B200FFT_INVALID_VALUE;
  // whole multiples of 2 MiB: smaller cudaMalloc requests are carved out of shared 2 MiB blocks and an IPC handle names the
  // block, not the piece
  bytes = (bytes + ((size_t)2 << 20) - 1) / ((size_t)2 << 20) * ((size_t)2 << 20);
  if (cudaMalloc(ptr, bytes) != cudaSuccess) { cudaGetLastError(); return B200FFT_ALLOC_FAILED; }
  return B200FFT_SUCCESS;
}
int b200fftPeerFree(void* ptr) {
  if (cudaFree(ptr) != cudaSuccess) { cudaGetLastError(); return B200FFT_INVALID_VALUE; }
  return B200FFT_SUCCESS;
}
int b200fftPeerExport(void* ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  if (!ptr || !handle) return B200FFT_INVALID_VALUE;
  if (cudaIpcGetMemHandle(&h, ptr) != cudaSuccess) { cudaGetLastError(); return B200FFT_NOT_SUPPORTED; }
  memcpy(handle, &h, 64);
  return B200FFT_SUCCESS;
}
int b200fftPeerOpen(const unsigned char handle[64], void** ptr) {
  cudaIpcMemHandle_t h;
  if (!ptr || !handle) return B200FFT_INVALID_VALUE;
  memcpy(&h, handle, 64);
  if (cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return B200FFT_NOT_SUPPORTED; }
  return B200FFT_SUCCESS;
}
int b200fftPeerClose(void* ptr) {
  if (cudaIpcCloseMemHandle(ptr) != cudaSuccess) { cudaGetLastError(); return B200FFT_INVALID_VALUE; }
  return B200FFT_SUCCESS;
}

int b200fftDestroy(b200fftHandle p) {
  if (!p || p->magic != 0xB200FF7u) return B200FFT_INVALID_PLAN;
  p->magic = 0;
  if (p->fallback) b200fftDestroy(p->fallback);
  for (auto& ps : p->passes) destroy_generic(&ps.gp);
  // cudaFree is legal from any thread; if the owning context is already gone the error is benign
  for (void* d : p->dev_allocs) cudaFree(d);
  cudaGetLastError();
  delete p;
  return B200FFT_SUCCESS;
}

const char* b200fftErrorString(int s) {
  switch (s) {
    case B200FFT_SUCCESS: return "B200FFT_SUCCESS";
    case B200FFT_INVALID_PLAN: return "B200FFT_INVALID_PLAN";
    case B200FFT_ALLOC_FAILED: return "B200FFT_ALLOC_FAILED";
    case B200FFT_INVALID_TYPE: return "B200FFT_INVALID_TYPE";
    case B200FFT_INVALID_VALUE: return "B200FFT_INVALID_VALUE";
    case B200FFT_INTERNAL_ERROR: return "B200FFT_INTERNAL_ERROR";
    case B200FFT_EXEC_FAILED: return "B200FFT_EXEC_FAILED";
    case B200FFT_INVALID_SIZE: return "B200FFT_INVALID_SIZE";
    case B200FFT_NO_DEVICE: return "B200FFT_NO_DEVICE (no CUDA device/context: there is no CPU fallback)";
    case B200FFT_NOT_SUPPORTED: return "B200FFT_NOT_SUPPORTED";
    default: return "B200FFT_UNKNOWN_ERROR";
  }
}

static int slab_pack_common(int type, bool pack, const void* src, void* dst, int64_t dl, int64_t h, int64_t w, int nranks,
                            b200fftStream stream) {
  int dbl;
  if (check_type(type, &dbl)) return B200FFT_INVALID_TYPE;
  if (!src || !dst || src == dst) return B200FFT_INVALID_VALUE;
  if (dl < 1 || h < 1 || w < 1 || nranks < 1 || h % nranks) return B200FFT_INVALID_SIZE;
  if (launch_slab_pack(dbl, pack, src, dst, dl, h, w, nranks, (cudaStream_t)stream) != cudaSuccess) { cudaGetLastError(); return B200FFT_EXEC_FAILED; }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return B200FFT_SUCCESS;
}
int b200fftSlabPack(int type, const void* src, void* dst, int64_t dl, int64_t h, int64_t w, int nranks, b200fftStream stream) {
  return slab_pack_common(type, true, src, dst, dl, h, w, nranks, stream);
}
int b200fftSlabUnpack(int type, const void* src, void* dst, int64_t dl, int64_t h, int64_t w, int nranks, b200fftStream stream) {
  return slab_pack_common(type, false, src, dst, dl, h, w, nranks, stream);
}

int b200fftTrimScratch(void) {
  std::lock_guard<std::mutex> g(g_pool_lock);
  for (auto& kv : g_pools) cudaMemPoolTrimTo(kv.second, 0);
  cudaGetLastError();
  return B200FFT_SUCCESS;
}

int b200fftHasExperimental(void) {
#ifdef B200FFT_EXPERIMENTAL
  return 1;
#else
  return 0;
#endif
}

size_t b200fftScratchBytes(b200fftHandle p) { return p ? p->scratch_bytes + p->extra_bytes + p->band_bytes : 0; }
int b200fftNumPasses(b200fftHandle p) { return p ? (int)p->passes.size() : 0; }
int64_t b200fftKernelLaunches(void) { return g_launches.load(); }

int b200fftDescribe(b200fftHandle p, char* buf, int buflen) {
  if (!p || !buf || buflen <= 0) return 0;
  std::string s;
  static const char* bn[] = {"in", "out", "scratch"};
  for (const auto& ps : p->passes) {
    s += ps.desc;
    s += " [";
    s += bn[ps.src];
    s += "->";
    s += bn[ps.dst];
    s += "]\n";
  }
  int n = (int)s.size() < buflen - 1 ? (int)s.size() : buflen - 1;
  memcpy(buf, s.data(), n);
  buf[n] = 0;
  return n;
}

}  // extern "C"
