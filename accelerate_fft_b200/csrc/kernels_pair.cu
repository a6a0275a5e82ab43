// "Row pair" kernels: contiguous row FFTs whose input is the radix-2 butterfly of two rows M rows apart
// (Geom::pre2_off) -- the first stage of a 2*M-long strided axis folded into the row pass, so that a 2D
// transform whose column length is 2 x (longest single-pass column) needs two HBM passes instead of three.
//                                           Cfg<T, N, E, TL, minb, R0, R1, R2, R3>
#include "kernel_inst.cuh"
namespace b200fft {
void register_pair(void (*add)(const KernelEntry&)) {
  REG_PAIR(float, 1024, 16, 1, 0, 16, 16, 4);
  REG_PAIR(float, 2048, 16, 1, 0, 16, 16, 8);
  REG_PAIR(float, 4096, 16, 1, 0, 16, 16, 16);
  REG_PAIR(float, 8192, 32, 1, 0, 32, 16, 16);         // v0: 256 thr x 128 regs
  REG_PAIR(float, 8192, 16, 1, 2, 16, 16, 8, 4);       // v1: 512 thr x 64 regs
  REG_PAIR(double, 1024, 16, 1, 0, 16, 8, 8);
  REG_PAIR(double, 2048, 16, 1, 0, 16, 16, 8);
  REG_PAIR(double, 4096, 16, 1, 2, 16, 16, 16);
}
}  // namespace b200fft
