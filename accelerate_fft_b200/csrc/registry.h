// Kernel registry: every instantiated line kernel describes itself here so the planner
// (plan.cu) can pick kernels by (type, N, flavour) without knowing template parameters.
#pragma once
#include <cstddef>

namespace b200fft {

// FL_RING: persistent TMA-fed rows (ring_kernel.cuh); FL_ROWPAIR: rows with the radix-2 pre-butterfly (Geom::pre2_off)
// FL_PIPE: persistent software-pipelined column kernel, CS = 1 or a cluster (pipe_kernel.cuh)
// FL_CLUSTER: strided lines of length N = N1*CS transformed by a cluster of CS CTAs (cluster_kernel.cuh)
// FL_RINGCOL: persistent single-buffer TMA-fed column kernel for tiles that fill an SM (ringcol_kernel.cuh)
enum Flavor { FL_ROW = 0, FL_COL = 1, FL_TRANS = 2, FL_RING = 3, FL_ROWPAIR = 4, FL_CLUSTER = 5, FL_PIPE = 6, FL_PIPEROW = 7, FL_CLUSTERROW = 8, FL_RINGCOL = 9, FL_RINGTRANS = 10 };

struct KernelEntry {
  int is_double;
  int N, E, TL, threads;
  int flavor;      // FL_ROW: rows in / rows out; FL_COL: line-fastest both; FL_TRANS: rows in / line-fastest out
  int tw4;         // multiplies by the four-step twiddle at the store
  int variant;     // registration order among entries with the same (type, N, flavor, tw4); 0 is the default
  int minb;
  size_t smem;
  int S, rad[4];
  int tw_len;      // stage twiddle table length (complex elements)
  int G, NS;       // FL_RING: thread groups per CTA, stage buffers
  int CS, N1;      // FL_CLUSTER: cluster size and per-CTA sub-length (N = N1*CS; S, rad, tw_len describe N1)
  const void* func;
  const void* loop_func;   // same kernel as a grid-stride loop over the tiles (last argument: tile count); null if not instantiated
};


// two line-kernel phases fused into one launch (fused_kernel.cuh)
struct FusedEntry {
  KernelEntry a, b;   // func unused; flavor / tw4 / N identify the phases
  const void* func;
  int threads;
  size_t smem;
};
// two phases fused through L2, persistent + TMA-fed (band_kernel.cuh)
struct BandEntry {
  int is_double;
  int mode;           // BandMode: 0 strided axis, 1 contiguous rows (transposing second phase)
  int outer;          // phase B multiplies by the outer four-step twiddle
  int N1, N2, TLA, TLB;
  KernelEntry a, b;   // stage-twiddle description of the two phases (func unused)
  int threads;
  size_t smem;
  const void* func;
};
const BandEntry* find_band(int is_double, int mode, int outer, int N1, int N2);
void register_band(void (*add)(const BandEntry&));

const FusedEntry* find_fused(int is_double, int NA, int flavA, int twA, int NB, int flavB);
void register_fused(void (*add)(const FusedEntry&));

const KernelEntry* find_kernel(int is_double, int N, int flavor, int tw4, int prefer_tl);
int list_kernels(const KernelEntry** out, int max);

// each kernels_*.cu exports one of these
void register_f32_small(void (*add)(const KernelEntry&));
void register_f32_large(void (*add)(const KernelEntry&));
void register_f32_col(void (*add)(const KernelEntry&));
void register_f64_small(void (*add)(const KernelEntry&));
void register_f64_large(void (*add)(const KernelEntry&));
void register_f64_col(void (*add)(const KernelEntry&));
void register_ring(void (*add)(const KernelEntry&));
void register_pair(void (*add)(const KernelEntry&));
void register_cluster(void (*add)(const KernelEntry&));
void register_pipe(void (*add)(const KernelEntry&));
void register_ringcol(void (*add)(const KernelEntry&));

}  // namespace b200fft
