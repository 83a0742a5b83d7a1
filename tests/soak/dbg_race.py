import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from kwage_b200 import capi
from oracle import oracle_py as O
import synth_cases as S

def run(k, lc, n_reads, seed=13, reps=2, read_len=150):
    bases, offsets = S.uniform_reads(seed, 0, n_reads, read_len)
    ob = O.Builder(k, 1, lc, 24); ob.add_reads(bases, offsets); exp = ob.num_valid(); ob.close()
    got = []
    for _ in range(reps):
        with capi.BloomBuilder(k, min_kmer_count=1, log2_count_len=lc, log2_max_len=24) as b:
            b.add_reads(bases, offsets)
            got.append(b.num_valid())
    print("k=%d lc=%d reads=%d tiles=%d exp=%d got=%s %s" % (k, lc, n_reads, (n_reads*read_len+2047)//2048, exp, got, "OK" if all(g == exp for g in got) else "MISMATCH"), flush=True)

run(32, 22, 5000, reps=1)
run(31, 23, 20000)
run(31, 21, 20000)
run(31, 20, 30000)
