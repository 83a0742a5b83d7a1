#!/usr/bin/env python3
"""BASELINE.json configs[4]: parameter sweep of the construction and search kernels on one B200.

  python bench_sweep.py [--reads N] [--out profiles/rNN_sweep.json]

Construction (raw mode, the reference's ground-truth rig bloom_test.cpp:268-275): k in {21,25,31,32}, 1..8 hashes,
filter sizes 2^20 .. 2^32 bits -- the filter is L2-resident up to 2^29 bits (64 MiB of the 126 MB L2) and HBM-resident
above, which is the regime change the sweep is there to show.  Counting mode: log2 counting-filter length 18..32 at
min_kmer_count 1.  Search: 1/3/5 hashes against slabs of 1024..32768 filter columns.

Every point is timed with CUDA events on the handle's stream after a warm-up pass, inputs resident in HBM and larger
than L2 (reads) or randomly gathered (slab).  k in 33..63 (raw mode, 128-bit words) is beyond the reference (word.h:10): parity
unpinned there; more than 8 hashes (hash.cpp:243) is not built.  Prints one JSON document; rows also go to --out.
"""
import argparse
import json
import sys
import time

READ_LEN = 150


GRIDS = {
    # full: every (h, L) at k = 31, six corner points at the other k
    "full": dict(ks=[21, 25, 31, 32, 47, 63], hs=[1, 2, 3, 4, 5, 7, 8], Ls=[20, 24, 26, 28, 29, 30, 31, 32], lcs=[18, 20, 22, 23, 24, 26, 28, 30, 32],
                 Fs=[1024, 8192, 32768], search_hs=[1, 3, 5]),
    # bounded: what `bench.py --stages sweep` runs (a few seconds): 1-7 hashes, filters on both sides of the L2 -> HBM crossover
    "bounded": dict(ks=[21, 25, 31, 32, 47, 63], hs=[1, 2, 3, 5, 7], Ls=[20, 26, 29, 30, 32], lcs=[24, 28, 30], Fs=[2048, 8192], search_hs=[1, 3, 5]),
    "quick": dict(ks=[31], hs=[3], Ls=[26, 32], lcs=[23, 30], Fs=[8192], search_hs=[3]),
}


def run(reads=2000000, reps=3, grid="full", device=0, log=sys.stderr):
    """-> the sweep document (dict).  One GPU; inputs resident in HBM; CUDA events on the handles' streams."""
    import torch
    from kwage_b200 import capi
    G = GRIDS[grid]

    class A:
        pass
    args = A()
    args.reads, args.reps, args.quick = reads, reps, False
    torch.cuda.set_device(device)
    n_reads, n_bases = args.reads, args.reads * READ_LEN
    d_bases = torch.empty(n_bases + 16, dtype=torch.uint8, device="cuda")
    d_offsets = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
    capi.synth_reads_dev(777, 0, n_reads, READ_LEN, d_bases.data_ptr(), d_offsets.data_ptr(), device=device)
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
    ks, hs, Ls = G["ks"], G["hs"], G["Ls"]
    for k in ks:
        kmers = n_reads * (READ_LEN - k + 1)
        for h in hs:
            for L in Ls:
                if k != 31 and (h not in (3, 5) or L not in (26, 29, 32)):
                    continue          # the full (h, L) grid at k = 31; the other k at six corner points
                b = capi.BloomBuilder(k, device=device, raw_num_hash=h, raw_log2_len=L)

                def step():
                    b.reset()
                    b.add_reads_dev(d_bases.data_ptr(), d_offsets.data_ptr(), n_reads, n_bases)
                sec = time_on(b.stream(), step, args.reps)
                assert b.num_valid() == kmers
                b.close()
                rows["raw_construction"].append({
                    "k": k, "parity": "pinned (reference range)" if k <= 32 else "unpinned (beyond word.h:10)", "num_hash": h, "log2_len": L, "filter_MiB": (1 << L) / 8 / 2**20,
                    "regime": "L2-resident" if (1 << L) // 8 <= 64 << 20 else "HBM-resident",
                    "kmer_inserts_per_s": kmers / sec, "bit_sets_per_s": kmers * h / sec, "ms": sec * 1e3,
                    # algorithmic HBM bytes: the bases once; HBM-resident filters add one 32-byte sector read-modify-write per bit
                    "hbm_GBps_algorithmic": (n_bases + (kmers * h * 64 if (1 << L) // 8 > 64 << 20 else 0)) / sec / 1e9})
                print(json.dumps(rows["raw_construction"][-1]), file=log)

    for lc in G["lcs"]:
        k = 31
        kmers = n_reads * (READ_LEN - k + 1)
        b = capi.BloomBuilder(k, device=device, min_kmer_count=1, log2_count_len=lc, log2_max_len=32)

        def step():
            b.reset()
            b.add_reads_dev(d_bases.data_ptr(), d_offsets.data_ptr(), n_reads, n_bases)
        sec = time_on(b.stream(), step, args.reps)
        n_valid = b.num_valid()
        b.close()
        g = "first touch, one level (bloom_first.cuh)" if lc <= 30 else "radix, two levels (bloom_count.cuh)"
        rows["counting_construction"].append({"k": k, "log2_count_len": lc, "partition": g, "kmer_occurrences_per_s": kmers / sec,
                                              "ms": sec * 1e3, "valid_kmers": n_valid})
        print(json.dumps(rows["counting_construction"][-1]), file=log)

    # search: slab of F columns x 2^L rows generated on the device
    nq, qlen, L = 2000, 1000, 24
    for F in G["Fs"]:
        row_pitch = (F // 8 + 15) // 16 * 16
        slab = torch.empty((1 << L) * row_pitch, dtype=torch.uint8, device="cuda")
        capi.synth_filter_bits_dev(999, 0, 1, slab.numel(), slab.numel(), slab.data_ptr(), device=device)
        q = torch.empty(nq * qlen + 16, dtype=torch.uint8, device="cuda")
        qo = torch.empty(nq + 1, dtype=torch.int64, device="cuda")
        capi.synth_reads_dev(4242, 0, nq, qlen, q.data_ptr(), qo.data_ptr(), device=device)
        count_pitch = (F + 3) // 4 * 4
        counts = torch.empty(nq * count_pitch, dtype=torch.int32, device="cuda")
        nk = torch.empty(nq, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        for h in G["search_hs"]:
            db = capi.Database.attach_dev(slab.data_ptr(), row_pitch, 31, h, L, F, device=device)

            def step():
                db.search_counts_dev(q.data_ptr(), qo.data_ptr(), nq, nq * qlen, nk.data_ptr(), counts.data_ptr(), count_pitch)
            sec = time_on(db.stream(), step, args.reps)
            n_k = int(nk.sum().item())
            db.close()
            rows["search"].append({"filters": F, "log2_len": L, "num_hash": h, "queries": nq, "query_kmers": n_k,
                                   "filter_kmer_tests_per_s": n_k * F / sec, "ms": sec * 1e3,
                                   "hbm_GBps_algorithmic": (n_k * h * (F // 8) + nq * F * 4) / sec / 1e9})
            print(json.dumps(rows["search"][-1]), file=log)
        del slab, counts

    del d_bases
    torch.cuda.empty_cache()
    # where the regime changes: the last L2-resident and the first HBM-resident filter length at k = 31, 3 hashes
    pts = sorted((r["log2_len"], r["bit_sets_per_s"], r["regime"]) for r in rows["raw_construction"] if r["k"] == 31 and r["num_hash"] == 3)
    cross = None
    for (l0, v0, g0), (l1, v1, g1) in zip(pts[:-1], pts[1:]):
        if g0 != g1:
            cross = {"last_L2_resident_log2_len": l0, "bit_sets_per_s": v0, "first_HBM_resident_log2_len": l1, "bit_sets_per_s_hbm": v1,
                     "ratio": v0 / v1 if v1 else None}
    return {"what": "configs[4] parameter sweep (%s grid), 1 x B200, inputs resident in HBM, CUDA-event timed; k > 32 and more than 8 hashes "
                    "are beyond the reference (word.h:10, hash.cpp:243): k in 33..63 runs in raw mode on 128-bit words with unpinned parity, more than "
                    "8 hashes is not built" % grid,
            "reads_per_point": n_reads, "read_len": READ_LEN, "time": time.strftime("%Y-%m-%d %H:%M:%S"), "l2_to_hbm_crossover": cross, **rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2000000, help="reads per point (150 bp)")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--grid", default="full", choices=sorted(GRIDS))
    args = ap.parse_args()
    from kwage_b200 import capi
    if capi.device_count() < 1:
        raise SystemExit("needs a CUDA device")
    doc = run(args.reads, args.reps, "quick" if args.quick else args.grid)
    s = json.dumps(doc, indent=1)
    if args.out:
        open(args.out, "w").write(s + "\n")
    print(json.dumps(doc))
    return 0


if __name__ == "__main__":
    sys.exit(main())
