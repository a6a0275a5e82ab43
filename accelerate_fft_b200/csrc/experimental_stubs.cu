// Default build: the kernels that were built, measured and found slower than the plans the library takes by default --
// the DSMEM cluster column / cluster rows kernels (cluster_kernel.cuh, kernels_cluster.cu), the lock-step L2-fused pair
// (fused_kernel.cuh, kernels_fused.cu) and the row-pair 2D kernels (kernels_pair.cu) -- are left out of libb200fft.so; the
// planner simply finds none of them registered and its opt-in switches (B200FFT_CLUSTER, B200FFT_CLUSTER_ROWS,
// B200FFT_FUSED, B200FFT_PAIR2D) have nothing to select.  `make B200FFT_EXPERIMENTAL=1` builds them in (DESIGN.md section 8
// records the measurements that closed them).
#include "registry.h"
namespace b200fft {
void register_pair(void (*)(const KernelEntry&)) {}
void register_cluster(void (*)(const KernelEntry&)) {}
void register_fused(void (*)(const FusedEntry&)) {}
}  // namespace b200fft
