// Fused two-phase kernels (fused_kernel.cuh): phase A -> L2-resident band -> phase B in one launch.
#include "fused_kernel.cuh"
#include "kernel_inst.cuh"
namespace b200fft {

template <class KA, bool A_LLF, bool A_SLF, bool A_TW4, class KB, bool B_LLF, bool B_SLF, bool B_TW4>
static FusedEntry make_fused() {
  FusedEntry f{};
  f.a = describe_cfg<KA, A_LLF, A_SLF, A_TW4>();
  f.b = describe_cfg<KB, B_LLF, B_SLF, B_TW4>();
  f.threads = KA::THREADS;
  f.smem = f.a.smem > f.b.smem ? f.a.smem : f.b.smem;
  f.func = reinterpret_cast<const void*>(&fft_fused2_kernel<KA, A_LLF, A_SLF, A_TW4, KB, B_LLF, B_SLF, B_TW4>);
  return f;
}
#define COLTW true, true, true
#define COL true, true, false
#define ROW false, false, false
#define TRANS false, true, false

void register_fused(void (*add)(const FusedEntry&)) {
  // ---- c64 ------------------------------------------------------------------------------------
  using F64 = Cfg<float, 64, 16, 32, 0, 16, 4>;               // 128 threads, 32 lines (256 B runs), 16 KB tiles
  using F128 = Cfg<float, 128, 16, 16, 0, 16, 8>;             // 128 threads
  using F256 = Cfg<float, 256, 16, 16, 0, 16, 16>;            // 256 threads
  using F512 = Cfg<float, 512, 32, 16, 0, 32, 16>;            // 256 threads
  using F1024c = Cfg<float, 1024, 32, 16, 0, 32, 32>;         // 512 threads
  using F1024r = Cfg<float, 1024, 16, 8, 1, 16, 16, 4>;       // 512 threads, 8 rows
  // strided axis N = NA*NB (cfg3's column axis 8192 = 64 x 128)
  add(make_fused<F64, COLTW, F64, COL>());
  add(make_fused<F64, COLTW, F128, COL>());
  add(make_fused<F128, COLTW, F128, COL>());
  // contiguous four-step N = NA*NB with the transposing second phase (cfg4's rows of 2^18 = 512 x 512)
  add(make_fused<F256, COLTW, F256, TRANS>());
  add(make_fused<F256, COLTW, F512, TRANS>());
  add(make_fused<F512, COLTW, F512, TRANS>());
  // x rows + y columns of one plane (cfg5)
  add(make_fused<F1024r, ROW, F1024c, COL>());
  // ---- c128 -----------------------------------------------------------------------------------
  using D64 = Cfg<double, 64, 8, 16, 0, 8, 8>;                // 128 threads
  using D128 = Cfg<double, 128, 8, 8, 0, 8, 8, 2>;            // 128 threads, 8 lines
  add(make_fused<D64, COLTW, D64, COL>());
}
}  // namespace b200fft
