// c64 column (line-fastest) and transposing kernels
#include "kernel_inst.cuh"
namespace b200fft {
void register_f32_col(void (*add)(const KernelEntry&)) {
  REG_COL(float, 2, 2, 128, 0, 2);
  REG_COL(float, 4, 4, 128, 0, 4);
  REG_COL(float, 8, 8, 128, 0, 8);
  REG_COL(float, 16, 16, 128, 0, 16);
  REG_COL(float, 32, 8, 32, 0, 8, 4);
  REG_COL(float, 64, 16, 32, 0, 16, 4);               // v0: 32 lines (256 B runs), 16 loads in flight per thread (cfg3: 556 -> 537 us)
  REG_COL(float, 64, 8, 16, 0, 8, 8);                 // v1: 16 lines; latency-bound (ncu: long_scoreboard 12.8 per issue)
  REG_COL(float, 128, 16, 32, 0, 16, 8);               // v0: 256 B runs (99-103 % of the copy peak; 95 % with 128 B runs at 4 MB row stride)
  REG_COL(float, 128, 16, 16, 0, 16, 8);               // v1: 128 B runs -- taken when the axis has fewer than 32 columns
  REG_COL(float, 128, 16, 8, 0, 16, 8);                // v2: 64 B runs
  REG_COL(float, 256, 16, 16, 0, 16, 16);
  REG_COL(float, 512, 32, 16, 0, 32, 16);             // v0: 256 thr x 128 regs, one exchange
  REG_COL(float, 512, 16, 16, 0, 16, 16, 2);          // v1
  REG_COL(float, 1024, 32, 16, 0, 32, 32);            // v0: 512 thr x 128 regs, 128 B runs, one exchange (best)
  REG_COL(float, 1024, 16, 8, 0, 16, 16, 4);          // v1: 512 thr, 64 B runs
  REG_COL(float, 1024, 16, 16, 0, 16, 16, 4);         // v2: 1024 thr, 128 B runs
  REG_COL(float, 2048, 32, 8, 0, 32, 16, 4);          // v0: 512 thr x 128 regs
  REG_COL(float, 2048, 16, 8, 0, 16, 16, 8);          // v1
  REG_COL(float, 4096, 16, 4, 0, 16, 16, 16);         // v0: 1024 thr x 64 regs, 32 B runs, 136 KB: 1 CTA/SM
  REG_COL(float, 4096, 16, 2, 0, 16, 16, 16);         // v1: 512 thr x 64 regs, 16 B runs, 68 KB: 2 CTA/SM
  REG_COL(float, 4096, 32, 2, 0, 32, 16, 8);          // v2: 256 thr x 128 regs, 16 B runs, 68 KB: 2 CTA/SM
  REG_COL(float, 4096, 32, 4, 0, 32, 16, 8);          // v3: 512 thr x 128 regs, 32 B runs, 136 KB: 1 CTA/SM
  REG_TRANS(float, 4, 2, 64, 0, 2, 2);
  REG_TRANS(float, 8, 4, 64, 0, 4, 2);
  REG_TRANS(float, 16, 4, 32, 0, 4, 4);
  REG_TRANS(float, 32, 8, 32, 0, 8, 4);
  REG_TRANS(float, 64, 8, 16, 0, 8, 8);
  REG_TRANS(float, 128, 16, 16, 0, 16, 8);
  REG_TRANS(float, 256, 16, 16, 0, 16, 16);
  REG_TRANS(float, 512, 32, 16, 0, 32, 16);           // v0
  REG_TRANS(float, 512, 16, 16, 0, 16, 16, 2);        // v1
  REG_TRANS(float, 1024, 32, 16, 0, 32, 32);          // v0: 512 thr x 128 regs, 128 B runs
  REG_TRANS(float, 1024, 16, 8, 0, 16, 16, 4);        // v1
  REG_TRANS(float, 2048, 16, 8, 0, 16, 16, 8);
  REG_TRANS(float, 4096, 16, 4, 0, 16, 16, 16);
}
}  // namespace b200fft
