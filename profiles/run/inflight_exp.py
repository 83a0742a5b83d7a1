"""Several accessions in flight on one GPU (one handle, one stream and one host thread each, inputs resident in HBM):
ms per accession for a list of knob settings (read by the library when a handle is created).
python profiles/run/inflight_exp.py [reads] [per_worker]"""
import os
import sys
import threading
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from kwage_b200 import capi, hostapi as H

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
per_worker = int(sys.argv[2]) if len(sys.argv) > 2 else 4
K, RL, LMAX = 31, 150, 32
n_bases = n_reads * RL
kmers = n_reads * (RL - K + 1)
lc = H.counting_filter_log2_len(n_bases)
POOL = 4
d_bases = [torch.empty(n_bases + 16, dtype=torch.uint8, device="cuda") for _ in range(POOL)]
d_off = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
for i in range(POOL):
    capi.synth_reads_dev(12345 + i, 0, n_reads, RL, d_bases[i].data_ptr(), d_off.data_ptr(), device=0)
torch.cuda.synchronize()

KNOBS = ("KWG_HASH_BPS", "KWG_INSERT_BPS", "KWG_INSERT_PRIO", "KWG_APPEND_REGS")
CONFIGS = [
    ({}, 1), ({}, 2), ({}, 3),
    ({"KWG_INSERT_BPS": "4"}, 1), ({"KWG_INSERT_BPS": "2"}, 1), ({"KWG_HASH_BPS": "4"}, 1), ({"KWG_HASH_BPS": "6"}, 1),
    ({"KWG_INSERT_BPS": "4", "KWG_INSERT_PRIO": "1"}, 2), ({"KWG_INSERT_BPS": "4", "KWG_INSERT_PRIO": "1"}, 3),
    ({"KWG_INSERT_BPS": "4", "KWG_HASH_BPS": "4"}, 2), ({"KWG_INSERT_BPS": "4", "KWG_HASH_BPS": "4"}, 3),
    ({"KWG_INSERT_BPS": "4", "KWG_HASH_BPS": "4", "KWG_INSERT_PRIO": "1"}, 3),
    ({"KWG_INSERT_BPS": "2", "KWG_HASH_BPS": "6", "KWG_INSERT_PRIO": "1"}, 3),
    ({"KWG_INSERT_BPS": "4", "KWG_HASH_BPS": "4", "KWG_INSERT_PRIO": "1"}, 4),
]
crcs = {}
for env, nfl in CONFIGS:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    hs = [capi.BloomBuilder(K, device=0, min_kmer_count=1, log2_count_len=lc, log2_max_len=LMAX) for _ in range(nfl)]
    outs = [torch.empty((1 << LMAX) // 8, dtype=torch.uint8, device="cuda") for _ in range(nfl)]
    res = {}

    def step(w, i):
        b = hs[w]
        b.reset()
        b.add_reads_dev(d_bases[i % POOL].data_ptr(), d_off.data_ptr(), n_reads, n_bases)
        nv = b.num_valid()
        L, h = H.optimal_bloom_param(K, nv, 0.25, 18, LMAX)
        b.finalize_dev(L, h, outs[w].data_ptr())
        b.sync()
        res[(w, i)] = (nv, L, h)

    def run(n, base):
        def loop(w):
            torch.cuda.set_device(0)
            for i in range(n):
                step(w, w + (base + i) * nfl)
        ts = [threading.Thread(target=loop, args=(w,)) for w in range(nfl)]
        [t.start() for t in ts]
        [t.join() for t in ts]

    run(2, 0)
    torch.cuda.synchronize()
    streams = [torch.cuda.ExternalStream(b.stream()) for b in hs]
    e0 = [torch.cuda.Event(enable_timing=True) for _ in hs]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in hs]
    for e, st in zip(e0, streams):
        e.record(st)
    run(per_worker, 2)
    for e, st in zip(e1, streams):
        e.record(st)
    torch.cuda.synchronize()
    ms = max(a.elapsed_time(z) for a in e0 for z in e1) / (per_worker * nfl)
    # the filter of accession 0 (worker 0 built it last in its first warm-up step... rebuild it now for the checksum)
    step(0, 0)
    nv, L, h = res[(0, 0)]
    crc = zlib.crc32(outs[0][: (1 << L) // 8].cpu().numpy().tobytes())
    crcs.setdefault((nv, L, h, crc), []).append(str(env))
    print("%-80s in_flight %d: %.3f ms per accession = %.3e k-mer inserts/s (valid %d, L %d, h %d, crc %08x)" %
          (env, nfl, ms, kmers / ms * 1e3, nv, L, h, crc), flush=True)
    for b in hs:
        b.close()
    del outs
    torch.cuda.empty_cache()
print("distinct results:", len(crcs))
assert len(crcs) == 1, crcs
