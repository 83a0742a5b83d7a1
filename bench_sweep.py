#!/usr/bin/env python3
"""BASELINE.json configs[4]: parameter sweep of the construction and search kernels on one B200.

  python bench_sweep.py [--reads N] [--out profiles/rNN_sweep.json]

Construction (raw mode, the reference's ground-truth rig bloom_test.cpp:268-275): k in {21,25,31,32}, 1..8 hashes,
filter sizes 2^20 .. 2^32 bits -- the filter is L2-resident up to 2^29 bits (64 MiB of the 126 MB L2) and HBM-resident
above, which is the regime change the sweep is there to show.  Counting mode: log2 counting-filter length 18..32 at
min_kmer_count 1.  Search: 1/3/5 hashes against slabs of 1024..32768 filter columns.

Every point is timed with CUDA events on the handle's stream after a warm-up pass, inputs resident in HBM and larger
than L2 (reads) or randomly gathered (slab).  k > 32 and more than 8 hashes are beyond the reference (word.h:10,
hash.cpp:243) and are not built.  Prints one JSON document; rows also go to --out.
"""
import argparse
import json
import sys
import time

READ_LEN = 150


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2000000, help="reads per point (150 bp)")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()

    import torch
    from kwage_b200 import capi, hostapi as H
    if capi.device_count() < 1:
        raise SystemExit("needs a CUDA device")
    torch.cuda.set_device(0)
    n_reads, n_bases = args.reads, args.reads * READ_LEN
    d_bases = torch.empty(n_bases + 16, dtype=torch.uint8, device="cuda")
    d_offsets = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
    capi.synth_reads_dev(777, 0, n_reads, READ_LEN, d_bases.data_ptr(), d_offsets.data_ptr(), device=0)
    torch.cuda.synchronize()

    def time_on(stream_ptr, fn, reps):
        st = torch.cuda.ExternalStream(stream_ptr)
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            fn()
        e1.record(st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3 / reps

    rows = {"raw_construction": [], "counting_construction": [], "search": []}
    ks = [21, 25, 31, 32] if not args.quick else [31]
    hs = [1, 2, 3, 4, 5, 7, 8] if not args.quick else [3]
    Ls = [20, 24, 26, 28, 29, 30, 31, 32] if not args.quick else [26, 32]
    for k in ks:
        kmers = n_reads * (READ_LEN - k + 1)
        for h in hs:
            for L in Ls:
                if k != 31 and (h not in (3, 5) or L not in (26, 29, 32)):
                    continue          # the full (h, L) grid at k = 31; the other k at six corner points
                b = capi.BloomBuilder(k, raw_num_hash=h, raw_log2_len=L)

                def step():
                    b.reset()
                    b.add_reads_dev(d_bases.data_ptr(), d_offsets.data_ptr(), n_reads, n_bases)
                sec = time_on(b.stream(), step, args.reps)
                assert b.num_valid() == kmers
                b.close()
                rows["raw_construction"].append({
                    "k": k, "num_hash": h, "log2_len": L, "filter_MiB": (1 << L) / 8 / 2**20,
                    "regime": "L2-resident" if (1 << L) // 8 <= 64 << 20 else "HBM-resident",
                    "kmer_inserts_per_s": kmers / sec, "bit_sets_per_s": kmers * h / sec, "ms": sec * 1e3,
                    # algorithmic HBM bytes: the bases once; HBM-resident filters add one 32-byte sector read-modify-write per bit
                    "hbm_GBps_algorithmic": (n_bases + (kmers * h * 64 if (1 << L) // 8 > 64 << 20 else 0)) / sec / 1e9})
                print(json.dumps(rows["raw_construction"][-1]), file=sys.stderr)

    for lc in ([18, 20, 22, 23, 24, 26, 28, 30, 32] if not args.quick else [23, 30]):
        k = 31
        kmers = n_reads * (READ_LEN - k + 1)
        b = capi.BloomBuilder(k, min_kmer_count=1, log2_count_len=lc, log2_max_len=32)

        def step():
            b.reset()
            b.add_reads_dev(d_bases.data_ptr(), d_offsets.data_ptr(), n_reads, n_bases)
        sec = time_on(b.stream(), step, args.reps)
        n_valid = b.num_valid()
        b.close()
        g = "single level" if lc + 1 - 15 <= 9 else "two levels"
        rows["counting_construction"].append({"k": k, "log2_count_len": lc, "partition": g, "kmer_occurrences_per_s": kmers / sec,
                                              "ms": sec * 1e3, "valid_kmers": n_valid})
        print(json.dumps(rows["counting_construction"][-1]), file=sys.stderr)

    # search: slab of F columns x 2^L rows generated on the device
    nq, qlen, L = 2000, 1000, 24
    for F in ([1024, 8192, 32768] if not args.quick else [8192]):
        row_pitch = (F // 8 + 15) // 16 * 16
        slab = torch.empty((1 << L) * row_pitch, dtype=torch.uint8, device="cuda")
        capi.synth_filter_bits_dev(999, 0, 1, slab.numel(), slab.numel(), slab.data_ptr(), device=0)
        q = torch.empty(nq * qlen + 16, dtype=torch.uint8, device="cuda")
        qo = torch.empty(nq + 1, dtype=torch.int64, device="cuda")
        capi.synth_reads_dev(4242, 0, nq, qlen, q.data_ptr(), qo.data_ptr(), device=0)
        count_pitch = (F + 3) // 4 * 4
        counts = torch.empty(nq * count_pitch, dtype=torch.int32, device="cuda")
        nk = torch.empty(nq, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        for h in ([1, 3, 5] if not args.quick else [3]):
            db = capi.Database.attach_dev(slab.data_ptr(), row_pitch, 31, h, L, F, device=0)

            def step():
                db.search_counts_dev(q.data_ptr(), qo.data_ptr(), nq, nq * qlen, nk.data_ptr(), counts.data_ptr(), count_pitch)
            sec = time_on(db.stream(), step, args.reps)
            n_k = int(nk.sum().item())
            db.close()
            rows["search"].append({"filters": F, "log2_len": L, "num_hash": h, "queries": nq, "query_kmers": n_k,
                                   "filter_kmer_tests_per_s": n_k * F / sec, "ms": sec * 1e3,
                                   "hbm_GBps_algorithmic": (n_k * h * (F // 8) + nq * F * 4) / sec / 1e9})
            print(json.dumps(rows["search"][-1]), file=sys.stderr)
        del slab, counts

    doc = {"what": "configs[4] parameter sweep, 1 x B200, inputs resident in HBM, CUDA-event timed", "reads_per_point": n_reads,
           "read_len": READ_LEN, "time": time.strftime("%Y-%m-%d %H:%M:%S"), **rows}
    s = json.dumps(doc, indent=1)
    if args.out:
        open(args.out, "w").write(s + "\n")
    print(json.dumps(doc))
    return 0


if __name__ == "__main__":
    sys.exit(main())
