#!/usr/bin/env python
"""Which path a row length takes (a Python restatement of the planner's rule, csrc/generic.cu plan_generic_axis + factor_small,
checked against the real planner by tests/test_parity_gpu.py::test_prime_radices_17_to_61_one_pass): counts over the range the
reference's suite draws from, n in [1, 1024] (test/Test/Base.hs:44-45).

    python tools/length_classes.py [--c128] [lo hi]
"""
import sys

RAD_SMALL_C64 = {2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 21, 24, 25, 27, 28, 30, 32}
RAD_SMALL_C128 = set(range(2, 17))
PRIMES_C64 = {17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61}
PRIMES_C128 = {17, 19, 23}


def min_stages(n, rad):
    """fewest radices from `rad` whose product is n (None if n does not factor over them)"""
    best = {1: 0}
    for m in range(2, n + 1):
        if n % m:
            continue
        c = [best[m // r] + 1 for r in rad if m % r == 0 and (m // r) in best]
        if c:
            best[m] = min(c)
    return best.get(n)


def classify(n, c128=False):
    if n & (n - 1) == 0:
        return "pow2"
    if n <= 32:
        return "tiny"
    small = RAD_SMALL_C128 if c128 else RAD_SMALL_C64
    primes = PRIMES_C128 if c128 else PRIMES_C64
    if min_stages(n, small) is not None:
        return "mixed"                       # prime factors <= 13: any number of stages
    st = min_stages(n, small | primes)
    if st is not None and st <= 2:
        return "mixed-prime"                 # a prime radix of 17 ... 61 in a one- or two-stage plan
    return "bluestein"


def main():
    c128 = "--c128" in sys.argv
    a = [int(x) for x in sys.argv[1:] if not x.startswith("-")]
    lo, hi = (a + [1, 1024])[:2] if len(a) >= 2 else (1, 1024)
    count = {}
    for n in range(lo, hi + 1):
        k = classify(n, c128)
        count[k] = count.get(k, 0) + 1
    print("%s, n in [%d, %d]:" % ("c128" if c128 else "c64", lo, hi), ", ".join("%s %d" % kv for kv in sorted(count.items())))


if __name__ == "__main__":
    main()
