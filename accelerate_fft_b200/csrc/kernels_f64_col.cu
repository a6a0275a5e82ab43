// c128 column (line-fastest) and transposing kernels
#include "kernel_inst.cuh"
namespace b200fft {
void register_f64_col(void (*add)(const KernelEntry&)) {
  REG_COL(double, 2, 2, 128, 2);
  REG_COL(double, 4, 4, 128, 4);
  REG_COL(double, 8, 8, 128, 8);
  REG_COL(double, 16, 8, 64, 8, 2);
  REG_COL(double, 32, 8, 32, 8, 4);
  REG_COL(double, 64, 8, 16, 8, 8);
  REG_COL(double, 128, 8, 8, 8, 8, 2);
  REG_COL(double, 256, 8, 8, 8, 8, 4);
  REG_COL(double, 512, 8, 8, 8, 8, 8);
  REG_COL(double, 1024, 8, 4, 8, 8, 8, 2);
  REG_COL(double, 2048, 8, 4, 8, 8, 8, 4);
  REG_TRANS(double, 4, 2, 64, 2, 2);
  REG_TRANS(double, 8, 4, 64, 4, 2);
  REG_TRANS(double, 16, 4, 32, 4, 4);
  REG_TRANS(double, 32, 8, 32, 8, 4);
  REG_TRANS(double, 64, 8, 16, 8, 8);
  REG_TRANS(double, 128, 8, 8, 8, 8, 2);
  REG_TRANS(double, 256, 8, 8, 8, 8, 4);
  REG_TRANS(double, 512, 8, 8, 8, 8, 8);
  REG_TRANS(double, 1024, 8, 4, 8, 8, 8, 2);
  REG_TRANS(double, 2048, 8, 4, 8, 8, 8, 4);
}
}  // namespace b200fft
