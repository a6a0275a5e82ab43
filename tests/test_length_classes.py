"""The planner's length classes over the reference suite's range n in [1, 1024] (test/Test/Base.hs:44-45), from the Python
restatement of the rule in tools/length_classes.py -- the counts DESIGN.md / README.md quote.  The restatement itself is checked
against the real planner on the GPU (tests/test_parity_gpu.py::test_prime_radices_17_to_61_one_pass)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_counts_quoted_in_the_docs():
    import length_classes as lc
    c64 = [lc.classify(n) for n in range(33, 1025)]
    assert c64.count("pow2") == 5 and c64.count("mixed") == 214
    assert c64.count("mixed-prime") == 278 and c64.count("bluestein") == 495      # 278 of the 773 lengths with a prime factor above 13
    c128 = [lc.classify(n, c128=True) for n in range(33, 1025)]
    assert c128.count("mixed-prime") == 51 and c128.count("bluestein") == 722


def test_examples():
    import length_classes as lc
    assert lc.classify(1024) == "pow2" and lc.classify(31) == "tiny" and lc.classify(1000) == "mixed"
    for n in (34, 323, 961, 37, 61, 888, 1007):
        assert lc.classify(n) == "mixed-prime", n
    for n in (986, 1023, 969, 67, 521, 2 * 67):         # three stages, or a prime factor above 61
        assert lc.classify(n) == "bluestein", n
    assert lc.classify(34, c128=True) == "mixed-prime" and lc.classify(58, c128=True) == "bluestein"   # c128: 17, 19, 23 only
