import sys; sys.argv=["x","none"]
exec(open("tools/cluster_check.py").read())
for lgI in (9, 10, 11, 12, 13):
    I = 1 << lgI; O = (1 << 26) // (8192 * I)
    for env in ({}, {"B200FFT_NO_PIPE": "1"}, {"B200FFT_NO_CLUSTER": "1"}):
        timing("[%d][8192][%d] stride %d KB %s" % (O, I, I * 8 // 1024, "cluster-simple" if "B200FFT_NO_PIPE" in env else "four-step" if env else "pipe"), "axis", (O, 8192, I), af.C2C, iters=5, env=env)
