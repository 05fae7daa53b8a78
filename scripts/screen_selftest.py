"""Device self-test of the float32 screen of phase 2 (mcall_device.cuh): accepted samples must equal the literal FP64
call; prints the rejection rate per q.  usage: python scripts/screen_selftest.py [nseeds]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcftools_b200 import abi, mcall

bad = 0
with mcall.MCaller(abi.CallParams(4)) as mc:
    for seed in range(1, 64, 3):                                    # pair sites: q1 = 10^(-seed/8), all 256^3 triples
        mism, rej = mc.selftest_div(20, seed=seed), mc.selftest_div(21, seed=seed)
        print("pair   q1=1e-%.2f  mismatches %d  rejected %.2e" % (seed / 8, mism, rej / 2 ** 24), flush=True)
        bad += mism
    for s1 in (1, 9, 20, 33, 47):
        for s2 in (2, 14, 30, 44, 63):
            seed = s1 | s2 << 6
            n = 50_000_000
            mism, rej = mc.selftest_div(22, n=n, seed=seed), mc.selftest_div(23, n=n, seed=seed)
            print("triple q1=1e-%.2f q2=1e-%.2f  mismatches %d  rejected %.2e" % (s1 / 8, s2 / 8, mism, rej / n), flush=True)
            bad += mism
print("TOTAL MISMATCHES", bad)
sys.exit(1 if bad else 0)
