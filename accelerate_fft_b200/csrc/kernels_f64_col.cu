// c128 column (line-fastest) and transposing kernels
#include "kernel_inst.cuh"
namespace b200fft {
void register_f64_col(void (*add)(const KernelEntry&)) {
  REG_COL(double, 2, 2, 128, 0, 2);
  REG_COL(double, 4, 4, 128, 0, 4);
  REG_COL(double, 8, 8, 128, 0, 8);
  REG_COL(double, 16, 8, 64, 0, 8, 2);
  REG_COL(double, 32, 8, 32, 0, 8, 4);
  REG_COL(double, 64, 8, 16, 0, 8, 8);
  REG_COL(double, 128, 8, 8, 0, 8, 8, 2);
  REG_COL(double, 256, 8, 8, 0, 8, 8, 4);
  REG_COL(double, 512, 16, 8, 0, 16, 8, 4);           // v0: 256 thr x 128 regs
  REG_COL(double, 512, 8, 8, 0, 8, 8, 8);             // v1
  REG_COL(double, 1024, 16, 8, 0, 16, 8, 8);          // v0: 512 thr x 128 regs, 128 B runs, 1 CTA/SM
  REG_COL(double, 1024, 16, 4, 0, 16, 8, 8);          // v1: 256 thr x 128 regs, 64 B runs
  REG_COL(double, 1024, 8, 4, 0, 8, 8, 8, 2);         // v2
  REG_COL(double, 2048, 16, 4, 0, 16, 16, 8);         // v0
  REG_COL(double, 2048, 8, 4, 0, 8, 8, 8, 4);         // v1
  REG_TRANS(double, 4, 2, 64, 0, 2, 2);
  REG_TRANS(double, 8, 4, 64, 0, 4, 2);
  REG_TRANS(double, 16, 4, 32, 0, 4, 4);
  REG_TRANS(double, 32, 8, 32, 0, 8, 4);
  REG_TRANS(double, 64, 8, 16, 0, 8, 8);
  REG_TRANS(double, 128, 8, 8, 0, 8, 8, 2);
  REG_TRANS(double, 256, 8, 8, 0, 8, 8, 4);
  REG_TRANS(double, 512, 16, 8, 0, 16, 8, 4);         // v0
  REG_TRANS(double, 512, 8, 8, 0, 8, 8, 8);           // v1
  REG_TRANS(double, 1024, 16, 8, 0, 16, 8, 8);        // v0
  REG_TRANS(double, 1024, 16, 4, 0, 16, 8, 8);        // v1
  REG_TRANS(double, 1024, 8, 4, 0, 8, 8, 8, 2);       // v2
  REG_TRANS(double, 2048, 16, 4, 0, 16, 16, 8);       // v0
  REG_TRANS(double, 2048, 8, 4, 0, 8, 8, 8, 4);       // v1
}
}  // namespace b200fft
