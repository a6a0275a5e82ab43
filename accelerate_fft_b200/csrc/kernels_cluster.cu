// Cluster (DSMEM) column kernels (cluster_kernel.cuh): strided axes of 4096 .. 16384 points in one HBM pass.
//                      ClusterCfg<Cfg<T, N1, E, TL, minb, R0, R1, R2, R3>, CS>   N = N1 * CS
#include "kernel_inst.cuh"
#include "cluster_kernel.cuh"
namespace b200fft {

template <class CC, bool TW4>
KernelEntry make_cluster_entry() {
  using K = typename CC::K;
  KernelEntry e{};
  e.is_double = sizeof(typename K::real) == 8;
  e.N = CC::N; e.N1 = K::N; e.CS = CC::CS; e.E = K::E; e.TL = K::TL;
  e.S = K::S;
  for (int i = 0; i < 4; i++) e.rad[i] = K::rad[i];
  e.tw_len = K::TW_LEN;
  e.flavor = FL_CLUSTER;
  e.tw4 = TW4;
  e.threads = K::THREADS;
  e.smem = CC::SMEM;
  e.minb = K::MINB;
  e.func = reinterpret_cast<const void*>(&fft_cluster_cols_kernel<CC, TW4>);
  return e;
}

template <class CC>
KernelEntry make_cluster_row_entry() {
  using K = typename CC::K;
  KernelEntry e{};
  e.is_double = sizeof(typename K::real) == 8;
  e.N = K::N; e.N1 = K::N; e.CS = CC::CS; e.E = K::E; e.TL = 1;
  e.S = K::S;
  for (int i = 0; i < 4; i++) e.rad[i] = K::rad[i];
  e.tw_len = K::TW_LEN;
  e.flavor = FL_CLUSTERROW;
  e.threads = K::THREADS;
  e.smem = CC::SMEM;
  e.minb = K::MINB;
  e.func = reinterpret_cast<const void*>(&fft_cluster_rows_kernel<CC>);
  return e;
}
#define REG_CLUSTER_ROWS(CS, ...) add(make_cluster_row_entry<ClusterRowCfg<Cfg<__VA_ARGS__>, CS>>())

#define REG_CLUSTER(CS, ...) add(make_cluster_entry<ClusterCfg<Cfg<__VA_ARGS__>, CS>, false>())

void register_cluster(void (*add)(const KernelEntry&)) {
  // c64 N = 8192 (cfg3's column axis)
  REG_CLUSTER(8, float, 1024, 32, 8, 2, 32, 32);          // v0: 256 thr x 128 regs, 64 B runs, 64 KB, one local exchange
  REG_CLUSTER(8, float, 1024, 16, 8, 2, 16, 16, 4);       // v1: 512 thr x 64 regs, 64 B runs, 64 KB
  REG_CLUSTER(8, float, 1024, 32, 16, 1, 32, 32);         // v2: 512 thr x 128 regs, 128 B runs, 128 KB
  REG_CLUSTER(16, float, 512, 32, 16, 2, 32, 16);         // v3: 256 thr x 128 regs, 128 B runs, 64 KB, cluster of 16
  REG_CLUSTER(4, float, 2048, 32, 8, 1, 32, 16, 4);       // v4: 512 thr x 128 regs, 64 B runs, 128 KB, cluster of 4
  REG_CLUSTER(16, float, 512, 32, 8, 4, 32, 16);          // v5: 128 thr x 128 regs, 64 B runs, 32 KB, cluster of 16, 4 CTA/SM
  REG_CLUSTER(16, float, 512, 16, 8, 4, 16, 16, 2);       // v6: 256 thr x 64 regs, 64 B runs, 32 KB, cluster of 16, 4 CTA/SM
  // c64 N = 4096
  REG_CLUSTER(8, float, 512, 32, 16, 2, 32, 16);          // v0: 256 thr, 128 B runs, 64 KB
  REG_CLUSTER(4, float, 1024, 32, 8, 2, 32, 32);          // v1: 256 thr, 64 B runs, 64 KB
  // c64 N = 16384
  REG_CLUSTER(8, float, 2048, 32, 8, 1, 32, 16, 4);       // v0: 512 thr, 64 B runs, 128 KB
  REG_CLUSTER(16, float, 1024, 32, 8, 2, 32, 32);         // v1: cluster of 16, 64 KB
  // rows of W points + the first radix-8 stage of the column axis (2D transforms with H = 8*M)
  REG_CLUSTER_ROWS(8, float, 8192, 32, 1, 2, 32, 16, 16);   // cfg3
  REG_CLUSTER_ROWS(8, float, 4096, 16, 1, 2, 16, 16, 16);
  REG_CLUSTER_ROWS(8, double, 4096, 16, 1, 2, 16, 16, 16);
  // c128 N = 4096 / 8192
  REG_CLUSTER(8, double, 512, 16, 8, 2, 16, 16, 2);       // 256 thr x 128 regs, 128 B runs, 64 KB
  REG_CLUSTER(8, double, 1024, 16, 4, 2, 16, 16, 4);      // 256 thr, 64 B runs, 64 KB
}
}  // namespace b200fft
