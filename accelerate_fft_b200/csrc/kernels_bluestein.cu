// One-launch Bluestein kernels (bluestein_kernel.cuh): the M-point power-of-two engine they run on, per element type.
#include "bluestein_kernel.cuh"

#include <vector>

namespace b200fft {

template <class K>
static BluesteinEntry make_bluestein() {
  BluesteinEntry e{};
  e.is_double = sizeof(typename K::real) == 8;
  e.M = K::N; e.TL = K::TL; e.threads = K::THREADS; e.S = K::S;
  for (int i = 0; i < 4; i++) e.rad[i] = K::rad[i];
  e.tw_len = K::TW_LEN;
  e.smem = K::template smem_bytes<false>();
  e.func = reinterpret_cast<const void*>(&bluestein_rows_kernel<K>);
  return e;
}

static const std::vector<BluesteinEntry>& entries() {
  static const std::vector<BluesteinEntry> v = {
      // (minb chosen so that ptxas may use 128 registers: the kernel holds two copies of the stage code and spills at 64)
      //                       T      M     E   TL minb radices
      make_bluestein<Cfg<float, 128, 16, 16, 4, 16, 8>>(),
      make_bluestein<Cfg<float, 256, 16, 8, 4, 16, 16>>(),
      make_bluestein<Cfg<float, 512, 16, 4, 4, 16, 16, 2>>(),
      make_bluestein<Cfg<float, 1024, 16, 2, 4, 16, 16, 4>>(),
      make_bluestein<Cfg<float, 2048, 16, 1, 4, 16, 16, 8>>(),
      make_bluestein<Cfg<float, 4096, 16, 1, 2, 16, 16, 16>>(),
      make_bluestein<Cfg<double, 128, 8, 8, 4, 8, 8, 2>>(),
      make_bluestein<Cfg<double, 256, 8, 4, 4, 8, 8, 4>>(),
      make_bluestein<Cfg<double, 512, 8, 2, 4, 8, 8, 8>>(),
      make_bluestein<Cfg<double, 1024, 8, 2, 2, 8, 8, 8, 2>>(),
      make_bluestein<Cfg<double, 2048, 8, 1, 2, 8, 8, 8, 4>>(),
      make_bluestein<Cfg<double, 4096, 8, 1, 1, 8, 8, 8, 8>>(),
  };
  return v;
}

const BluesteinEntry* find_bluestein(int is_double, long long M) {
  for (const auto& e : entries())
    if (e.is_double == is_double && e.M == M) return &e;
  return nullptr;
}

}  // namespace b200fft
