// c64 row kernels, N = 2048 .. 16384
#include "kernel_inst.cuh"
namespace b200fft {
void register_f32_large(void (*add)(const KernelEntry&)) {
  REG_ROW(float, 2048, 16, 1, 16, 16, 8);
  REG_ROW(float, 4096, 16, 1, 16, 16, 16);
  REG_ROW(float, 8192, 32, 1, 32, 16, 16);
  REG_ROW(float, 16384, 32, 1, 32, 32, 16);
}
}  // namespace b200fft
