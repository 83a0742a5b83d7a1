import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from kwage_b200 import capi
from oracle import oracle_py as O
import synth_cases as S
case = dict(kind="coverage", seed=884924425, genome=200000, n_reads=4000, read_len=150, num_bp=-1)
bases, offsets = S.make_bloom_reads(case)
k = 31
for lc in (21, 22, 20):
    out = []
    with capi.BloomBuilder(k, min_kmer_count=1, log2_count_len=lc, log2_max_len=24) as b:
        for n in range(200, 4001, 100):
            ob = O.Builder(k, 1, lc, 24); ob.add_reads(bases, offsets[: n + 1]); e = ob.num_valid(); ob.close()
            b.reset(); b.add_reads(bases, offsets[: n + 1]); g = b.num_valid()
            out.append((n, -(-n * 150 // 2048), g - e))
    print("lc", lc, [(n, t, d) for n, t, d in out if d != 0] or "all ok")
