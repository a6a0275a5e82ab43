// c64 column (line-fastest) and transposing kernels
#include "kernel_inst.cuh"
namespace b200fft {
void register_f32_col(void (*add)(const KernelEntry&)) {
  REG_COL(float, 2, 2, 128, 2);
  REG_COL(float, 4, 4, 128, 4);
  REG_COL(float, 8, 8, 128, 8);
  REG_COL(float, 16, 16, 128, 16);
  REG_COL(float, 32, 8, 32, 8, 4);
  REG_COL(float, 64, 8, 16, 8, 8);
  REG_COL(float, 128, 16, 16, 16, 8);
  REG_COL(float, 256, 16, 16, 16, 16);
  REG_COL(float, 512, 16, 16, 16, 16, 2);
  REG_COL(float, 1024, 16, 8, 16, 16, 4);
  REG_COL(float, 2048, 16, 8, 16, 16, 8);
  REG_COL(float, 4096, 16, 4, 16, 16, 16);
  REG_TRANS(float, 4, 2, 64, 2, 2);
  REG_TRANS(float, 8, 4, 64, 4, 2);
  REG_TRANS(float, 16, 4, 32, 4, 4);
  REG_TRANS(float, 32, 8, 32, 8, 4);
  REG_TRANS(float, 64, 8, 16, 8, 8);
  REG_TRANS(float, 128, 16, 16, 16, 8);
  REG_TRANS(float, 256, 16, 16, 16, 16);
  REG_TRANS(float, 512, 16, 16, 16, 16, 2);
  REG_TRANS(float, 1024, 16, 8, 16, 16, 4);
  REG_TRANS(float, 2048, 16, 8, 16, 16, 8);
  REG_TRANS(float, 4096, 16, 4, 16, 16, 16);
}
}  // namespace b200fft
