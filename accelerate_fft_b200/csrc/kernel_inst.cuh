// Helper included by the kernels_*.cu instantiation units.
#pragma once
#include "fft_kernel.cuh"
#include "registry.h"

namespace b200fft {

// everything but the entry point
template <class K, bool LLF, bool SLF, bool TW4, bool PRE2 = false>
KernelEntry describe_cfg() {
  KernelEntry e{};
  e.is_double = sizeof(typename K::real) == 8;
  e.N = K::N; e.E = K::E; e.TL = K::TL; e.threads = K::THREADS;
  e.flavor = PRE2 ? FL_ROWPAIR : (LLF && SLF) ? FL_COL : (!LLF && SLF) ? FL_TRANS : FL_ROW;
  e.tw4 = TW4;
  e.minb = K::MINB;
  e.smem = K::template smem_bytes<(LLF && SLF)>();
  e.S = K::S;
  for (int i = 0; i < 4; i++) e.rad[i] = K::rad[i];
  e.tw_len = K::TW_LEN;
  return e;
}

template <class K, bool LLF, bool SLF, bool TW4, bool PRE2 = false>
KernelEntry make_entry() {
  KernelEntry e = describe_cfg<K, LLF, SLF, TW4, PRE2>();
  e.func = reinterpret_cast<const void*>(&fft_lines_kernel<K, LLF, SLF, TW4, PRE2>);
  // strided-axis kernels of 128+ points (the ones a scatter pass can use) also come as a grid-stride loop
  if constexpr (LLF && SLF && !TW4 && !PRE2 && K::N >= 128 && K::S >= 2)
    e.loop_func = reinterpret_cast<const void*>(&fft_lines_loop_kernel<K, LLF, SLF, TW4>);
  return e;
}

template <class R>
KernelEntry make_ring_entry();

#define REG_ROW(...)   add(make_entry<Cfg<__VA_ARGS__>, false, false, false>())
#define REG_COL(...)   add(make_entry<Cfg<__VA_ARGS__>, true, true, false>()); add(make_entry<Cfg<__VA_ARGS__>, true, true, true>())
#define REG_PAIR(...)  add(make_entry<Cfg<__VA_ARGS__>, false, false, false, true>())
#define REG_TRANS(...) add(make_entry<Cfg<__VA_ARGS__>, false, true, false>())

}  // namespace b200fft
