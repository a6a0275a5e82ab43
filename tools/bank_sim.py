#!/usr/bin/env python
"""Offline shared-memory bank-conflict model for the Stockham exchange layouts.

Model: 32 banks x 4 B.  An access of `esz` bytes per lane is served in groups of
128/esz lanes (half-warps for 8 B, quarter-warps for 16 B); a group costs as many
wavefronts as the maximum number of distinct 128-byte-row words that fall into one bank.
Prints the average wavefronts per group (1.0 == conflict free) for the write and read side
of every exchange of a radix plan, for the row ("t fastest") thread mapping.
"""
import sys
from collections import defaultdict


def wavefronts(addrs_bytes, esz):
    grp = 128 // esz
    tot = 0
    ngrp = 0
    for g in range(0, len(addrs_bytes), grp):
        banks = defaultdict(set)
        for a in addrs_bytes[g:g + grp]:
            if a is None:
                continue
            for w in range(esz // 4):
                word = a // 4 + w
                banks[word % 32].add(word)
        if banks:
            tot += max(len(v) for v in banks.values())
            ngrp += 1
    return tot, ngrp


def padidx(i, logq):
    return i + (i >> logq) if logq >= 0 else i


def sim_row(N, E, radices, esz, logq, lpc=1, pitch=None):
    TPT = N // E
    if pitch is None:
        pitch = padidx(N - 1, logq) + 1
    nthreads = TPT * lpc
    res = []
    Ns = 1
    for s, R in enumerate(radices[:-1]):
        B = E // R
        wt = wg = rt = rg = 0
        for w0 in range(0, nthreads, 32):
            lanes = range(w0, min(w0 + 32, nthreads))
            for b in range(B):
                for q in range(R):
                    addrs = []
                    for tid in lanes:
                        line, t = divmod(tid, TPT)
                        j = t + b * TPT
                        idx = (j // Ns) * Ns * R + (j % Ns) + q * Ns
                        addrs.append((line * pitch + padidx(idx, logq)) * esz)
                    a, g = wavefronts(addrs, esz)
                    wt += a; wg += g
            for e in range(E):
                addrs = []
                for tid in lanes:
                    line, t = divmod(tid, TPT)
                    idx = t + e * TPT
                    addrs.append((line * pitch + padidx(idx, logq)) * esz)
                a, g = wavefronts(addrs, esz)
                rt += a; rg += g
        res.append((s, R, Ns, wt / wg, rt / rg))
        Ns *= R
    return res


if __name__ == "__main__":
    cases = [
        (1024, 16, [16, 16, 4], 8), (1024, 16, [4, 16, 16], 8), (1024, 16, [16, 4, 16], 8),
        (4096, 16, [16, 16, 16], 8), (4096, 16, [16, 16, 16], 16), (4096, 8, [8, 8, 8, 8], 16),
        (8192, 16, [16, 16, 8, 4], 8), (8192, 32, [32, 16, 16], 8), (8192, 32, [16, 16, 32], 8),
        (256, 16, [16, 16], 8), (64, 8, [8, 8], 8), (64, 16, [16, 4], 8), (128, 16, [16, 8], 8),
        (512, 16, [16, 16, 2], 8), (2048, 16, [16, 16, 8], 8), (16384, 32, [32, 32, 16], 8),
        (2048, 16, [16, 16, 8], 16), (1024, 16, [16, 16, 4], 16), (256, 16, [16, 16], 16),
    ]
    for N, E, rad, esz in cases:
        TPT = N // E
        lpc = max(1, 64 // TPT)
        for logq in ([-1, 4, 5] if esz == 8 else [-1, 3, 4]):
            r = sim_row(N, E, rad, esz, logq, lpc)
            print(f"N={N} E={E} rad={rad} esz={esz} lpc={lpc} logq={logq}: " +
                  "  ".join(f"[s{s} R{R} Ns{Ns} W{w:.2f} R{rd:.2f}]" for s, R, Ns, w, rd in r))
