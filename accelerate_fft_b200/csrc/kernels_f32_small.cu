// c64 row kernels, N = 2 .. 1024          Cfg<T, N, E, TL, R0, R1, R2, R3>
#include "kernel_inst.cuh"
namespace b200fft {
void register_f32_small(void (*add)(const KernelEntry&)) {
  REG_ROW(float, 2, 2, 128, 0, 2);
  REG_ROW(float, 4, 4, 128, 0, 4);
  REG_ROW(float, 8, 8, 128, 0, 8);
  REG_ROW(float, 16, 16, 128, 0, 16);
  REG_ROW(float, 32, 32, 64, 0, 32);                   // v0: one line per thread, rows staged through shared memory (68 % -> 106 %)
  REG_ROW(float, 32, 8, 32, 0, 8, 4);                  // v1
  REG_ROW(float, 64, 8, 16, 0, 8, 8);
  REG_ROW(float, 128, 16, 16, 0, 16, 8);
  REG_ROW(float, 256, 16, 8, 0, 16, 16);
  REG_ROW(float, 512, 16, 4, 0, 16, 16, 2);
  // N=1024 (cfg1): one exchange (32 x 32), 64 threads x 128 registers.  Small batches are bound by ramp-up and drain
  // (a 32 MB device copy runs at 4.6 TB/s, not 6.5): this variant takes 15.6 us at batch 4096 against 16.6 us for the
  // three-stage one, and both stream large batches at ~6.9 TB/s (profiles/r01_cfg1_variants.txt)
  REG_ROW(float, 1024, 32, 2, 8, 32, 32);              // v0
  REG_ROW(float, 1024, 16, 2, 0, 16, 16, 4);           // v1: 128 threads x 64 registers, two exchanges
}
}  // namespace b200fft
