"""One accession through the counting construction (device-resident input), for ncu captures:
python profiles/run/construct_once.py [reads] [reps] [min_kmer_count]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from kwage_b200 import capi, hostapi as H

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = int(sys.argv[3]) if len(sys.argv) > 3 else 1
K, RL = 31, 150
n_bases = n_reads * RL
lc = H.counting_filter_log2_len(n_bases)
d_bases = torch.empty(n_bases + 16, dtype=torch.uint8, device="cuda")
d_off = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
capi.synth_reads_dev(12345, 0, n_reads, RL, d_bases.data_ptr(), d_off.data_ptr(), device=0)
d_out = torch.empty((1 << 32) // 8, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
b = capi.BloomBuilder(K, device=0, min_kmer_count=c, log2_count_len=lc, log2_max_len=32)
for r in range(reps):
    b.reset()
    b.add_reads_dev(d_bases.data_ptr(), d_off.data_ptr(), n_reads, n_bases)
    nv = b.num_valid()
    L, h = H.optimal_bloom_param(K, nv, 0.25, 18, 32)
    b.finalize_dev(L, h, d_out.data_ptr())
    b.sync()
    print(r, nv, L, h)
b.close()
