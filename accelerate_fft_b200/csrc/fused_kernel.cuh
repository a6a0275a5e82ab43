// Two line-kernel phases in ONE launch, with the intermediate kept in L2 instead of HBM.
//
// Problem: an axis whose lines are strided AND too long for one CTA's shared memory (cfg3's column axis:
// 8192 points x 16 columns = 1 MiB; cfg4's 2^18-point rows) needs two Stockham passes (four-step), and two
// launches mean the intermediate goes out to HBM and back: 3 HBM passes for a 2D FFT instead of 2.
// Thread-block clusters cannot hold it either (DSMEM moves ~20 B/clk/SM, below the HBM share of an SM).
// The 126 MB L2 can: the work is cut into BANDS small enough (8-16 MiB) that phase A's output for a band is
// still in L2 when phase B reads it.
//
// Scheduling: persistent, co-resident CTAs take TICKETS round-robin and decode each into (phase, band, tile) in the order
//     A(0) .. A(la-1), A(la), B(0), A(la+1), B(1), ... , A(nb-1), B(nb-1-la), B(nb-la) .. B(nb-1)
// so phase A runs `la` bands ahead of phase B.  A B-tile spins until all A-tiles of its band have signalled
// (release/acquire on a per-band counter); an A-tile of band b first waits for the B-tiles of band
// b - nslots that still read its scratch slot.  With nslots >= la + 2 both waits refer to work whose tickets
// were handed out at least a whole band earlier, so in steady state nobody actually spins (measured: with
// two slots every band boundary drained the machine, ~8 us each).
// Scratch reads use ld.global.cg (L2 only: another SM wrote the data during this launch); the HBM-side
// traffic uses evict-first hints so the scratch slots stay resident.
//
// Replaces, for its share of a plan, cufftExecC2C / cufftExecZ2Z behind PTX.hs:112-124.
#pragma once
#include "fft_kernel.cuh"

namespace b200fft {

struct FusedParams {
  Geom a, b;                    // per-phase tile geometry, tile indices local to a band
  int nbands, nA, nB;           // tiles per band in each phase
  int ka, kb;                   // tiles per ticket in each phase
  int nbi;                      // band index = bo * nbi + bi
  long long a_in_bo, a_in_bi;   // element offsets of a band in phase A's input
  long long b_out_bo, b_out_bi; // ... and in phase B's output
  int la, nslots;               // phase A runs `la` bands ahead of phase B; scratch slots (>= la + 2)
  long long slot_elems;         // > 0: A writes / B reads scratch slot (band % nslots) of this many elements
  long long mid_bo, mid_bi;     // slot_elems == 0: A writes / B reads `mid` at these band offsets (in place in out)
};

// counters: [0] ticket, [1 .. nbands] A-tiles done per band, [1 + nbands .. 2 nbands] B-tiles done per band
template <class KA, bool A_LLF, bool A_SLF, bool A_TW4, class KB, bool B_LLF, bool B_SLF, bool B_TW4>
__global__ void __launch_bounds__(KA::THREADS, (KA::MINB < KB::MINB ? KA::MINB : KB::MINB))
fft_fused2_kernel(const FusedParams P, const cpx_t<typename KA::real>* __restrict__ in, cpx_t<typename KA::real>* __restrict__ out,
                  cpx_t<typename KA::real>* __restrict__ mid, const cpx_t<typename KA::real>* __restrict__ twsA,
                  const cpx_t<typename KA::real>* __restrict__ twsB, const cpx_t<typename KA::real>* __restrict__ tw_lo,
                  const cpx_t<typename KA::real>* __restrict__ tw_hi, typename KA::real scale, unsigned* __restrict__ counters) {
  using T = typename KA::real;
  using C = cpx_t<T>;
  static_assert(KA::THREADS == KB::THREADS, "both phases run in the same CTA shape");
  static_assert(sizeof(typename KA::real) == sizeof(typename KB::real), "one element type");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* sm = reinterpret_cast<C*>(smem_raw);

  // tickets hand out GROUPS of ka (kb) consecutive tiles, so that the flag check, the fence and the signal
  // are paid once per 32-64 KB of work rather than once per 8-16 KB tile
  const unsigned nA = (P.nA + P.ka - 1) / P.ka, nB = (P.nB + P.kb - 1) / P.kb, seg = nA + nB;
  const unsigned total = (unsigned)P.nbands * seg;
  const unsigned head = (unsigned)P.la * nA;

  // decode: A(0..la-1) | [A(j) B(j-la)] for j = la..nb-1 | B(nb-la..nb-1)     (all by value: no stack traffic)
  struct Tk { int phase, band; unsigned grp; };
  auto decode = [=](unsigned ticket) -> Tk {
    if (ticket < head) return Tk{0, (int)(ticket / nA), ticket % nA};
    const unsigned r0 = ticket - head;
    const int j = (int)(r0 / seg) + P.la;
    if (j < P.nbands) {
      const unsigned r = r0 % seg;
      return r < nA ? Tk{0, j, r} : Tk{1, j - P.la, r - nA};
    }
    const unsigned r1 = r0 - (unsigned)(P.nbands - P.la) * seg;
    return Tk{1, P.nbands - P.la + (int)(r1 / nB), r1 % nB};
  };
  // what a ticket has to wait for: producers of what it reads (B) / consumers of the slot it overwrites (A)
  struct Dep { const unsigned* flag; unsigned target; };
  auto dependency = [=](int phase, int band) -> Dep {
    if (phase == 1) return Dep{counters + 1 + band, nA};
    if (P.slot_elems > 0 && band >= P.nslots) return Dep{counters + 1 + P.nbands + (band - P.nslots), nB};
    return Dep{nullptr, 0u};
  };
  auto peek = [](const unsigned* flag) -> unsigned {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return v;
  };

  // Persistent CTAs, launched cooperatively (all co-resident), take tickets round-robin: CTA c processes
  // tickets c, c + gridDim, c + 2 gridDim, ...  (A shared atomic ticket counter was measured first: ~210
  // tickets/us on one L2 address is at the limit of same-address atomic throughput, and its latency had to be
  // hidden with a two-deep software pipeline.)  A ticket only ever waits on lower tickets, and with every CTA
  // resident the CTA holding the lowest unfinished ticket can always run: no deadlock.
  // Thread 0 looks at the NEXT ticket's dependency flag while the current group is processed and publishes the
  // outcome with the single end-of-group barrier, so in steady state (dependency long satisfied) the other
  // threads go straight into the next group's loads while thread 0 alone fences and signals.
  __shared__ unsigned s_ready[2];
  auto ready_now = [=](unsigned ticket) -> unsigned {
    if (ticket >= total) return 1u;
    const Tk k = decode(ticket);
    const Dep d = dependency(k.phase, k.band);
    return (!d.flag || peek(d.flag) >= d.target) ? 1u : 0u;
  };
  if (threadIdx.x == 0) s_ready[0] = ready_now(blockIdx.x);
  __syncthreads();
  int it = 0;
  for (unsigned ticket = blockIdx.x; ticket < total; ticket += gridDim.x, it++) {
    const Tk cur = decode(ticket);
    const int phase = cur.phase, band = cur.band;
    const unsigned grp = cur.grp;
    if (!s_ready[it & 1]) {   // rare: the producers of this group have not finished yet
      if (threadIdx.x == 0) {
        const Dep d = dependency(phase, band);
        while (peek(d.flag) < d.target) __nanosleep(64);
      }
      __syncthreads();
    }
    const unsigned nxt = ticket + gridDim.x;
    unsigned seen = 0xffffffffu, tgt = 0;
    if (threadIdx.x == 0 && nxt < total) {   // the round trip is only consumed after this group's work
      const Tk k = decode(nxt);
      const Dep d = dependency(k.phase, k.band);
      if (d.flag) { tgt = d.target; seen = peek(d.flag); }
    }

    const int bo = band / P.nbi, bi = band % P.nbi;
    C* midp = (P.slot_elems > 0) ? mid + (long long)(band % P.nslots) * P.slot_elems : mid + (long long)bo * P.mid_bo + (long long)bi * P.mid_bi;
    if (phase == 0) {
      const C* ip = in + (long long)bo * P.a_in_bo + (long long)bi * P.a_in_bi;
      for (int i = 0; i < P.ka; i++) {
        const unsigned tile = grp * P.ka + i;
        if (tile >= (unsigned)P.nA) break;
        if (i) __syncthreads();
        fft_lines_tile<KA, A_LLF, A_SLF, A_TW4, false, 1>(P.a, tile, ip, midp, twsA, tw_lo, tw_hi, (T)1, sm);
      }
    } else {
      C* op = out + (long long)bo * P.b_out_bo + (long long)bi * P.b_out_bi;
      for (int i = 0; i < P.kb; i++) {
        const unsigned tile = grp * P.kb + i;
        if (tile >= (unsigned)P.nB) break;
        if (i) __syncthreads();
        fft_lines_tile<KB, B_LLF, B_SLF, B_TW4, true, 2>(P.b, tile, midp, op, twsB, tw_lo, tw_hi, scale, sm);
      }
    }

    if (threadIdx.x == 0) s_ready[(it + 1) & 1] = seen >= tgt ? 1u : 0u;
    __syncthreads();   // group done by all threads; readiness of the next one published; exchange buffer free
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&counters[1 + (phase ? P.nbands : 0) + band], 1u);
    }
  }
}

}  // namespace b200fft
