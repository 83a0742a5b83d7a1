#!/usr/bin/env python
"""bench.py -- KWAGE hot path on B200: Bloom construction (headline), transposition, bit-sliced search.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line
from rank 0.  A "step" is one pass of the hot path over one batch of synthetic input:

  construct (headline, BASELINE.json configs[1]): one accession = 1e6 synthetic 150 bp reads
      (1.2e8 k-mer occurrences), k=31, counting mode with min_kmer_count=1, adaptive parameters
      (p=0.25, L in [18,32]) -> the reference picks L=29, 3 hashes.  Accessions shard across GPUs
      (weak scaling, no collective).
  transpose (configs[2]): 4096 filters x 2^26 bits per GPU column slab.
  search    (configs[3]): 10k x 1 kb queries against the per-GPU slab of the 65,536-accession DB
      (8192 columns x 2^26 rows = 64 GiB in HBM); hit lists are gathered to rank 0 over NCCL.

`value` is measured with inputs resident in HBM; `e2e` goes through the host-buffer C-ABI calls with
pinned host buffers (H2D/D2H inside the timed region).  `--impl reference` times the UNMODIFIED
reference (oracle/_ref, compiled from /root/reference) on the host cores for the same metric.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 31
READ_LEN = 150
P_FALSE = 0.25
LMIN, LMAX = 18, 32


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels, from the `ncu --set full` captures of the
# default workload.  profiles/ncu_traffic.json is written by `profiles/extract_ncu.py traffic <reports>` and carries the
# commit the captures were taken at; roofline.traffic is reported only when the workload matches the captured one.
def ncu_traffic():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return d.get("kernels", {}), "profiles/ncu_traffic.json (captured at commit %s: %s)" % (d.get("commit", "?"), ", ".join(d.get("reports", [])))
    except Exception:
        return {}, None


NCU_TRAFFIC, NCU_TRAFFIC_SOURCE = ncu_traffic()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, windows):
        """windows: list of (t0, t1) wall-clock intervals of the timed regions"""
        sm, smax, reasons = [], 0.0, set()
        for ts, line in self.lines:
            if windows and not any(a <= ts <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = max(smax, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- distributed plumbing
class Dist:
    def __init__(self, n_gpus):
        import torch
        self.torch = torch
        self.world = env_int("WORLD_SIZE", 1)
        self.rank = env_int("RANK", 0)
        self.local_rank = env_int("LOCAL_RANK", 0)
        self.enabled = self.world > 1
        if self.enabled:
            import torch.distributed as dist
            self.dist = dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", self.local_rank))
        else:
            torch.cuda.set_device(0)
        self.device = self.local_rank if self.enabled else 0

    def barrier(self):
        if self.enabled:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if not self.enabled:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if not self.enabled:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def finish(self):
        if self.enabled:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed(D, stream_ptr, fn, steps, warmup, windows):
    """W untimed + exactly K timed calls of fn(i); CUDA events on the stream the library launches on,
    barrier + synchronize on both sides, max over ranks.  Returns seconds for the K steps."""
    torch = D.torch
    stream = torch.cuda.ExternalStream(stream_ptr) if stream_ptr else torch.cuda.current_stream()
    for i in range(warmup):
        fn(i)
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record(stream)
    for i in range(steps):
        fn(warmup + i)
    e1.record(stream)
    D.barrier()
    windows.append((w0, time.time()))
    return D.max_over_ranks(e0.elapsed_time(e1) / 1e3)


def run_concurrent(D, dev, handles, step, per_worker, warm, windows):
    """step(w, i) on one host thread per handle (its own stream), `warm` untimed then `per_worker` timed calls each.
    Returns the device time from the earliest start to the latest end over the handles' streams, max over ranks."""
    torch = D.torch

    def run(n, base):
        def loop(w):
            torch.cuda.set_device(dev)
            for i in range(n):
                step(w, base + i)
        ts = [threading.Thread(target=loop, args=(w,)) for w in range(len(handles))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    if warm:
        run(warm, 0)
    D.barrier()
    streams = [torch.cuda.ExternalStream(hd.stream()) for hd in handles]
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in handles]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in handles]
    w0 = time.time()
    for e, st in zip(ev0, streams):
        e.record(st)
    run(per_worker, warm)
    for e, st in zip(ev1, streams):
        e.record(st)
    D.barrier()
    windows.append((w0, time.time()))
    return D.max_over_ranks(max(a.elapsed_time(z) for a in ev0 for z in ev1) / 1e3)


def pack_2na_torch(torch, d_ascii, n_bases):
    """bench input preparation only: ASCII bases on the device -> pinned host 2na bytes (kwg_bloom_add_packed's format)"""
    x = d_ascii[:n_bases].to(torch.int32)
    c = ((x >> 1) & 3)
    c = c ^ (c >> 1)                                   # A:0 C:1 G:2 T:3
    pad = (-n_bases) % 4
    if pad:
        c = torch.cat([c, torch.zeros(pad, dtype=torch.int32, device=c.device)])
    c = c.view(-1, 4)
    packed = ((c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]).to(torch.uint8)
    h = torch.empty(packed.numel(), dtype=torch.uint8).pin_memory()
    h.copy_(packed)
    return h


# ---------------------------------------------------------------------------------------------- construction
def stage_construct(D, args, windows):
    import numpy as np
    torch = D.torch
    from kwage_b200 import capi, hostapi as H
    n_reads = args.reads
    n_bases = n_reads * READ_LEN
    kmers = n_reads * (READ_LEN - K + 1)
    lc = H.counting_filter_log2_len(n_bases)
    dev = D.device
    pool = min(4, args.steps + args.warmup)
    d_bases = [torch.empty(n_bases + 16, dtype=torch.uint8, device="cuda") for _ in range(pool)]
    d_offsets = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
    for i in range(pool):      # accession seeds follow SURVEY 8d: 12345 + accession index (distinct per rank)
        capi.synth_reads_dev(12345 + D.rank * 100000 + i, 0, n_reads, READ_LEN, d_bases[i].data_ptr(), d_offsets.data_ptr(), device=dev)
        if args.coverage > 0:
            # reads sampled from a random genome so that k-mers recur (needed for min_kmer_count > 1); torch only
            # prepares the synthetic input here
            g = torch.Generator(device="cuda")
            g.manual_seed(12345 + D.rank * 100000 + i)
            glen = max(READ_LEN + 1, int(n_bases / args.coverage))
            genome = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")[torch.randint(0, 4, (glen,), generator=g, device="cuda")]
            starts = torch.randint(0, glen - READ_LEN, (n_reads,), generator=g, device="cuda")
            for a in range(0, n_reads, 100000):
                z = min(n_reads, a + 100000)
                idx = starts[a:z, None] + torch.arange(READ_LEN, device="cuda")[None, :]
                d_bases[i][a * READ_LEN: z * READ_LEN] = genome[idx].reshape(-1)
            del genome, starts, idx
    torch.cuda.synchronize()
    d_out = torch.empty((1 << LMAX) // 8, dtype=torch.uint8, device="cuda")
    b = capi.BloomBuilder(K, device=dev, min_kmer_count=args.min_kmer_count, log2_count_len=lc, log2_max_len=LMAX)
    state = {}

    def step_dev(i):
        b.reset()
        b.add_reads_dev(d_bases[i % pool].data_ptr(), d_offsets.data_ptr(), n_reads, n_bases)
        n_valid = b.num_valid()
        L, h = H.optimal_bloom_param(K, n_valid, P_FALSE, LMIN, LMAX)     # host-side parameter choice, as in the reference
        b.finalize_dev(L, h, d_out.data_ptr())
        state.update(n_valid=n_valid, L=L, h=h)

    # per-kernel times and the latency of one accession: one handle, one stream, inputs resident in HBM
    for i in range(args.warmup):
        step_dev(i)
    b.sync()
    b.set_timing(True)
    b.get_timing()
    launches0 = capi.launch_count()
    sec_single = timed(D, b.stream(), step_dev, args.steps, 0, windows)
    launches = capi.launch_count() - launches0
    ms, nl = b.get_timing()
    b.set_timing(False)

    # value: whole-job throughput with inputs resident in HBM.  A job is many accessions (the reference runs one MPI
    # worker per accession, many per node): IN_FLIGHT host threads, each with its own handle and stream, build whole
    # accessions one after the other, so that the kernels of one accession that leave SM resources unused (the
    # issue-bound hash pass, the L2-atomic-bound insert) run beside another accession's.
    in_flight = max(1, args.in_flight)
    vb = [b] + [capi.BloomBuilder(K, device=dev, min_kmer_count=args.min_kmer_count, log2_count_len=lc, log2_max_len=LMAX)
                for _ in range(in_flight - 1)]
    d_outs = [d_out] + [torch.empty((1 << LMAX) // 8, dtype=torch.uint8, device="cuda") for _ in range(in_flight - 1)]

    def step_dev_w(w, i):
        bw = vb[w]
        bw.reset()
        bw.add_reads_dev(d_bases[i % pool].data_ptr(), d_offsets.data_ptr(), n_reads, n_bases)
        n_valid = bw.num_valid()
        L, h = H.optimal_bloom_param(K, n_valid, P_FALSE, LMIN, LMAX)
        bw.finalize_dev(L, h, d_outs[w].data_ptr())
        bw.sync()

    value_steps_per_worker = max(2, (args.steps + in_flight - 1) // in_flight)
    sec = run_concurrent(D, dev, vb, lambda w, i: step_dev_w(w, w + i * in_flight), value_steps_per_worker, 2, windows)
    value_steps = value_steps_per_worker * in_flight
    for bw in vb[1:]:
        bw.close()
    del d_outs

    # e2e: pinned host buffers through the host-pointer C-ABI calls.  E2E_WORKERS host threads per GPU, each with
    # its own handle, stream and buffers and each building whole accessions (the reference's model: one MPI worker
    # per accession, many workers per node), so that one accession's H2D / D2H overlaps the other's kernels.
    n_workers = max(1, args.e2e_workers)
    use_ft = args.min_kmer_count == 1 and lc <= 30
    h_offsets = torch.empty(n_reads + 1, dtype=torch.int64).pin_memory()
    h_offsets.copy_(d_offsets)
    builders = [b] + [capi.BloomBuilder(K, device=dev, min_kmer_count=args.min_kmer_count, log2_count_len=lc, log2_max_len=LMAX)
                      for _ in range(n_workers - 1)]
    h_bases, h_out = [], []
    for w in range(n_workers):
        hb = torch.empty(n_bases, dtype=torch.uint8).pin_memory()
        hb.copy_(d_bases[w % pool][:n_bases])
        h_bases.append(hb)
        h_out.append(torch.empty((1 << state["L"]) // 8, dtype=torch.uint8).pin_memory())
    torch.cuda.synchronize()

    def step_host(w, i):
        bw = builders[w]
        bw.reset()
        bw.add_reads_ptr(h_bases[w].data_ptr(), h_offsets.data_ptr(), n_reads)
        n_valid = bw.num_valid()
        L, h = H.optimal_bloom_param(K, n_valid, P_FALSE, LMIN, LMAX)
        bw.finalize_crc_ptr(L, h, h_out[w].data_ptr())      # bits + their crc32, what make_bloom_filter writes into the .bloom file

    per_worker = max(1, (args.steps + n_workers - 1) // n_workers)
    sec_e2e = run_concurrent(D, dev, builders, step_host, per_worker, 2, windows)
    e2e_steps = per_worker * n_workers
    import zlib
    crc_ascii = zlib.crc32(h_out[0].numpy().tobytes()) & 0xFFFFFFFF

    # the same through kwg_bloom_add_packed: the host holds the reads 2-bit packed (NCBI 2na, what the SRA stores), a
    # quarter of the bytes cross PCIe
    h_packed = [pack_2na_torch(torch, d_bases[w % pool], n_bases) for w in range(n_workers)]
    torch.cuda.synchronize()

    def step_host_packed(w, i):
        bw = builders[w]
        bw.reset()
        bw.add_packed_ptr(h_packed[w].data_ptr(), 0, h_offsets.data_ptr(), n_reads)
        n_valid = bw.num_valid()
        L, h = H.optimal_bloom_param(K, n_valid, P_FALSE, LMIN, LMAX)
        bw.finalize_crc_ptr(L, h, h_out[w].data_ptr())
        state["n_valid_packed"] = n_valid

    sec_e2e_packed = run_concurrent(D, dev, builders, step_host_packed, per_worker, 2, windows)
    if (zlib.crc32(h_out[0].numpy().tobytes()) & 0xFFFFFFFF) != crc_ascii:
        raise SystemExit("bench: the packed-input arm built a different filter than the ASCII arm")
    crc = crc_ascii
    for bw in builders:
        bw.close()
    del d_bases, d_out
    torch.cuda.empty_cache()

    n = D.world
    peak, peak_src = measured_peaks()
    # Algorithmic HBM bytes per k-mer occurrence of each kernel of the counting pipeline.
    ascii_b = READ_LEN / (READ_LEN - K + 1) * (1 + 1 / 8)            # bases + read-start bitmap
    touched_b = (2 << lc) / 8 / kmers                                # touched bitmap written once per batch
    if use_ft:
        # first-touch path (bloom_first.cuh): the four counting hashes of an occurrence are written once (16 B) and read
        # once, its four 6-byte touch records are written once and read once, its canonical word (8 B) is written once
        # and read once by the insert
        alg = {
            "hash": ("ft_count_kernel + ft_scan_kernel + ft_hash_kernel (encode, canonical k-mer, 4 hashes by dense ordinal)", capi.T_SCAN_A,
                     2 * ascii_b + 16.0 + 8.0 + 4.0),        # + seed 2's hash, kept behind the words of the list
            "append": ("ft_append_kernel (one-level partition: ring-buffered append into page chains)", capi.T_REGROUP, 16.0 + 24.0),
            "resolve": ("ft_resolve_kernel (stream-ordered first touch against a 1-bit-per-slot tile in shared memory)", capi.T_RESOLVE,
                        24.0 + touched_b + 0.5),
            "finish": ("ft_finish_kernel (invalid bitmap + valid count)", capi.T_SCAN_B, 0.5),
            # seeds (0,1) and (2,3) of the filter are the fold of the counting tables' touched bitmap (2^lc bits read per table);
            # only a last odd seed is set per occurrence: from its kept hash (three hashes: 4 B + 1 validity bit) or from the word (8 B + 1 bit)
            "insert_words": ("fold_touched_kernel (seed pairs out of the touched bitmap) + insert_words_kernel (a last odd seed, L2-resident red.or)",
                             capi.T_INSERT,
                             ((state["h"] // 2) * (1 << lc) / 8 + (1 << state["L"]) / 8) / kmers + ((4.125 if state["h"] == 3 else 8.125) if state["h"] & 1 else 0.0)
                             if state["L"] <= lc and state["h"] >= 2 else 8.125 + (1 << state["L"]) / 8 / kmers),
        }
    else:
        # two-level radix partition (bloom_count.cuh): a touch record is 8 bytes, four per k-mer: the partition scan writes
        # them once, regroup reads and re-writes them once, resolve reads them once; everything else is small next to that.
        alg = {
            "partition_scan": ("partition_scan_kernel (encode + 4 hashes + tile-local counting sort)", capi.T_SCAN_A, ascii_b + 32.0),
            "regroup": ("regroup_kernel (bulk-copy gather + level-2 counting sort)", capi.T_REGROUP, 64.0 + 2 * 257 * 2 / 8192 * 32),
            "resolve": ("resolve_kernel (first-touch atomicMin in shared memory)", capi.T_RESOLVE, 32.0 + touched_b + 0.5) if args.min_kmer_count == 1 else
                       # one launch per counter level: records read once per level (+ the dense copy written by level 0), the 4-bit
                       # counters (2^lc bytes) read and written once per level
                       ("resolve_kernel<levels> + resolve_dense_kernel x%d levels" % args.min_kmer_count, capi.T_RESOLVE,
                        32.0 * (args.min_kmer_count + 1) + 2.0 * args.min_kmer_count * (1 << lc) / kmers),
            "scan_pass_b": ("kmer_scan_kernel<PASS_B> (valid-word list)", capi.T_SCAN_B, ascii_b + 0.5 + 8.0),
            "insert_words": ("insert_words_kernel (final filter, L2-resident red.or)", capi.T_INSERT, 8.0 + (1 << state["L"]) / 8 / kmers),
        }
    single_ms = sec_single / args.steps * 1e3
    step_ms = sec / value_steps * 1e3
    per_kernel = {}
    for name, (desc, idx, bpk) in alg.items():
        t = ms[idx] / max(int(nl[idx]), 1) / 1e3 if nl[idx] else 0.0
        t_step = float(ms[idx]) / args.steps
        per_kernel[name] = {"kernel": desc, "ms_per_step": round(t_step, 4), "launches_per_step": int(nl[idx]) // max(args.steps, 1),
                            "algorithmic_bytes_per_kmer": round(bpk, 3),
                            "achieved_gbs": round(kmers * bpk / (t_step / 1e3) / 1e9, 1) if t_step > 0 else 0.0}
    per_kernel["mark_read_starts"] = {"kernel": "mark_read_starts_kernel", "ms_per_step": round(float(ms[capi.T_AUX]) / args.steps, 4)}
    dom = max(alg, key=lambda k: per_kernel[k]["ms_per_step"])
    achieved = per_kernel[dom]["achieved_gbs"]
    kernels = {k: v["ms_per_step"] for k, v in per_kernel.items()}
    return {
        "value": n * kmers * value_steps / sec,
        "ms_per_step": step_ms,
        "accessions_in_flight": in_flight,
        "single_stream": {"value": n * kmers * args.steps / sec_single, "ms_per_step": single_ms,
                          "note": "one handle, one stream, one accession at a time (the latency of an accession; the per-kernel times below are taken here)"},
        # headline e2e: the reads cross PCIe 2-bit packed (kwg_bloom_add_packed; SURVEY.md 8d, config 2: "generated on device or
        # streamed 2-bit-packed"; NCBI 2na is how the SRA stores them, and what the host layer's parser thread produces); the
        # ASCII call of the drop-in shim (kwg_bloom_add_reads, 4x the H2D bytes) is reported beside it
        "e2e": {"value": n * kmers * e2e_steps / sec_e2e_packed, "unit": "kmer_inserts/s", "h2d_bytes_per_step": (n_bases + 3) // 4 + 8 * (n_reads + 1),
                "d2h_bytes_per_step": (1 << state["L"]) // 8 + 8, "ms_per_step": sec_e2e_packed / e2e_steps * 1e3, "steps": e2e_steps,
                "workers_per_gpu": n_workers,
                "ascii_input": {"value": n * kmers * e2e_steps / sec_e2e, "ms_per_step": sec_e2e / e2e_steps * 1e3,
                                "h2d_bytes_per_step": n_bases + 8 * (n_reads + 1),
                                "how": "the same through kwg_bloom_add_reads (one byte per base, what NGS hands the reference)"},
                "how": "%d host threads per GPU, one accession at a time each through kwg_bloom_add_packed/num_valid/finalize_crc with pinned "
                       "host buffers (2-bit NCBI 2na reads + 64-bit offsets in, filter bits + crc32 out); device time from the first start "
                       "to the last end over their streams" % n_workers},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": per_kernel[dom]["kernel"], "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC.get({"hash": "ft_hash_kernel", "append": "ft_append_kernel", "resolve": "ft_resolve_kernel"}.get(dom, dom + "_kernel"))
                     if (n_reads == 1000000 and args.min_kmer_count == 1 and use_ft) else None,
                     "traffic_source": NCU_TRAFFIC_SOURCE,
                     "algorithmic_bytes": kmers * per_kernel[dom]["algorithmic_bytes_per_kmer"], "peak_source": peak_src,
                     "algorithmic_bytes_per_kmer": per_kernel[dom]["algorithmic_bytes_per_kmer"],
                     "kernel_ms": per_kernel[dom]["ms_per_step"], "share_of_step": per_kernel[dom]["ms_per_step"] / single_ms,
                     "pipeline_algorithmic_bytes_per_kmer": round(sum(v[2] for v in alg.values()), 2),
                     "pipeline_achieved_gbs": round(kmers * sum(v[2] for v in alg.values()) / (step_ms / 1e3) / 1e9, 1),
                     # SURVEY.md 8d's per-unit figure for production mode: 4 x 64 B counting-table sector RMW + 5 x 64 B valid-bit
                     # sector RMW + 1.25 B of ASCII per k-mer occurrence (tables HBM-resident).  The partition design removes
                     # that traffic, so this figure over the measured time can exceed the peak; the design-bytes figures above
                     # are the stricter ones.
                     "survey_8d_model": {"bytes_per_kmer": 576.0 + READ_LEN / (READ_LEN - K + 1),
                                         "achieved": round(kmers * (576.0 + READ_LEN / (READ_LEN - K + 1)) / (step_ms / 1e3) / 1e9, 1),
                                         "frac": kmers * (576.0 + READ_LEN / (READ_LEN - K + 1)) / (step_ms / 1e3) / 1e9 / peak},
                     "note": "achieved = algorithmic bytes of this kernel / its mean duration (CUDA events on the handle's stream, single-stream pass); "
                             "the first-touch pipeline moves ~104 B per k-mer occurrence where round 1's radix pipeline moved 155 B and SURVEY 8d's "
                             "sector read-modify-write model 577 B, so less time reads as a lower fraction: compare ms_per_step; none of its "
                             "kernels is HBM-bound (instruction issue, shared-memory pipe, L2 atomics: profiles/r3final_ncu_full_construct.md)"},
        "kernel_ms_per_step": kernels,
        "kernels": per_kernel,
        "result": {"num_valid_kmers": state["n_valid"], "log2_filter_len": state["L"], "num_hash": state["h"], "log2_counting_filter_len": lc,
                   "filter_crc32": crc},
        "kmers_per_step": kmers,
    }


# ---------------------------------------------------------------------------------------------- transposition
def stage_transpose(D, args, windows):
    torch = D.torch
    from kwage_b200 import capi
    n_filters, L = args.tr_filters, args.tr_log2
    fbytes = (1 << L) // 8
    dev = D.device
    d_in = torch.empty(n_filters * fbytes, dtype=torch.uint8, device="cuda")
    row_pitch = (n_filters + 7) // 8
    row_pitch = (row_pitch + 15) // 16 * 16
    d_out = torch.empty((1 << L) * row_pitch, dtype=torch.uint8, device="cuda")
    capi.synth_filter_bits_dev(999, D.rank * n_filters, n_filters, fbytes, fbytes, d_in.data_ptr(), device=dev)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()

    def step_dev(i):
        capi.transpose_dev(d_in.data_ptr(), fbytes, n_filters, 1 << L, d_out.data_ptr(), row_pitch, device=dev, stream=stream.cuda_stream)

    steps, warmup = max(2, min(args.steps, 5)), max(3, min(args.warmup, 3))
    launches0 = capi.launch_count()
    sec = timed(D, 0, step_dev, steps, warmup, windows)
    bits = n_filters * (1 << L)
    del d_in, d_out
    torch.cuda.empty_cache()

    # e2e: one reference-sized chunk (2048 filters x 4 Mi slices, build_db.cpp:243) through kwg_transpose
    import ctypes as C
    nf, cbits = 2048, 1 << 22
    h_in = torch.empty(nf * cbits // 8, dtype=torch.uint8).pin_memory()
    h_in.random_(0, 256)
    h_out = torch.empty(cbits * nf // 8, dtype=torch.uint8).pin_memory()
    ptrs = (C.c_void_p * nf)(*[h_in.data_ptr() + j * (cbits // 8) for j in range(nf)])

    def step_host(i):
        capi.check(capi.lib().kwg_transpose(dev, ptrs, nf, cbits, C.c_void_p(h_out.data_ptr())))

    D.barrier()
    step_host(0)
    D.barrier()
    t0 = time.time()
    for i in range(2):
        step_host(i)
    D.barrier()
    sec_e2e = D.max_over_ranks(time.time() - t0)
    windows.append((t0, time.time()))
    n = D.world
    peak, peak_src = measured_peaks()
    achieved = 2 * bits / 8 * steps / sec / 1e9
    return {"metric": "transposed bits/s", "value": n * bits * steps / sec, "unit": "bits/s", "ms_per_step": sec / steps * 1e3,
            "config": {"n_filters_per_gpu": n_filters, "log2_filter_len": L, "hbm_bytes_per_step": 2 * bits // 8},
            "e2e": {"value": n * nf * cbits * 2 / sec_e2e, "unit": "bits/s", "h2d_bytes_per_step": nf * cbits // 8, "d2h_bytes_per_step": nf * cbits // 8,
                    "workload": "kwg_transpose, 2048 filters x 2^22 slices per call (the reference's chunk)"},
            "roofline": {"bound": "hbm", "kernel": "transpose_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC.get("transpose_kernel") if (n_filters, L) == (4096, 26) else None, "traffic_source": NCU_TRAFFIC_SOURCE,
                         "algorithmic_bytes": 2 * bits // 8, "peak_source": peak_src},
            "gpu_launches": int(capi.launch_count() - launches0)}


# ---------------------------------------------------------------------------------------------- raw construction
def stage_construct_raw(D, args, windows):
    """Raw mode (north_star: canonical k-mers + multi-hash + atomicOr into an L2-resident filter; the reference's
    ground-truth rig bloom_test.cpp:268-275): one accession of --reads reads, k=31, 3 hashes, 2^29 bits."""
    torch = D.torch
    from kwage_b200 import capi
    n_reads, L, h = args.reads, 29, 3
    n_bases = n_reads * READ_LEN
    kmers = n_reads * (READ_LEN - K + 1)
    dev = D.device
    pool = 4
    d_bases = [torch.empty(n_bases + 16, dtype=torch.uint8, device="cuda") for _ in range(pool)]
    d_offsets = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
    for i in range(pool):
        capi.synth_reads_dev(22345 + D.rank * 100000 + i, 0, n_reads, READ_LEN, d_bases[i].data_ptr(), d_offsets.data_ptr(), device=dev)
    d_out = torch.empty((1 << L) // 8, dtype=torch.uint8, device="cuda")
    b = capi.BloomBuilder(K, device=dev, raw_num_hash=h, raw_log2_len=L)

    def step_dev(i):
        b.reset()
        b.add_reads_dev(d_bases[i % pool].data_ptr(), d_offsets.data_ptr(), n_reads, n_bases)
        b.finalize_dev(L, h, d_out.data_ptr())

    for i in range(3):
        step_dev(i)
    b.sync()
    b.set_timing(True)
    b.get_timing()
    launches0 = capi.launch_count()
    steps = max(args.steps, 5)
    sec = timed(D, b.stream(), step_dev, steps, 0, windows)
    launches = capi.launch_count() - launches0
    ms, nl = b.get_timing()
    b.set_timing(False)
    t_scan = float(ms[capi.T_SCAN_A]) / steps / 1e3

    # e2e: three host threads per GPU, each with its own handle: one accession's D2H rides beside another's H2D and a
    # third one's scan (two workers fall into lockstep: both scan, then both copy out)
    n_workers = 3
    rb = [b] + [capi.BloomBuilder(K, device=dev, raw_num_hash=h, raw_log2_len=L) for _ in range(n_workers - 1)]
    h_offsets = torch.empty(n_reads + 1, dtype=torch.int64).pin_memory()
    h_offsets.copy_(d_offsets)
    h_bases, h_packed, h_out = [], [], []
    for w in range(n_workers):
        hb = torch.empty(n_bases, dtype=torch.uint8).pin_memory()
        hb.copy_(d_bases[w % pool][:n_bases])
        h_bases.append(hb)
        h_packed.append(pack_2na_torch(torch, d_bases[w % pool], n_bases))
        h_out.append(torch.empty((1 << L) // 8, dtype=torch.uint8).pin_memory())
    torch.cuda.synchronize()

    def step_host(w, i):
        rb[w].reset()
        rb[w].add_reads_ptr(h_bases[w].data_ptr(), h_offsets.data_ptr(), n_reads)
        rb[w].finalize_crc_ptr(L, h, h_out[w].data_ptr())

    def step_host_packed(w, i):
        rb[w].reset()
        rb[w].add_packed_ptr(h_packed[w].data_ptr(), 0, h_offsets.data_ptr(), n_reads)
        rb[w].finalize_crc_ptr(L, h, h_out[w].data_ptr())

    per_worker = max(3, (steps + n_workers - 1) // n_workers)
    e2e_steps = per_worker * n_workers
    sec_e2e = run_concurrent(D, dev, rb, step_host, per_worker, 1, windows)
    sec_e2e_packed = run_concurrent(D, dev, rb, step_host_packed, per_worker, 1, windows)
    for x in rb[1:]:
        x.close()
    b.close()
    del d_bases, d_out
    torch.cuda.empty_cache()
    peak, peak_src = measured_peaks()
    n = D.world
    # algorithmic HBM bytes: the bases and the read-start bitmap in, the filter cleared once and copied out once; the
    # 3 bit sets per k-mer are L2 atomics on a 64 MiB filter that never leaves the 126 MB L2
    alg = n_bases * (1 + 1 / 8) + 2 * (1 << L) / 8
    return {"metric": "Bloom k-mer inserts/s (raw mode)", "value": n * kmers * steps / sec, "unit": "kmer_inserts/s", "ms_per_step": sec / steps * 1e3,
            "config": {"reads_per_accession": n_reads, "kmer_len": K, "num_hash": h, "log2_filter_len": L},
            "e2e": {"value": n * kmers * e2e_steps / sec_e2e_packed, "unit": "kmer_inserts/s", "h2d_bytes_per_step": (n_bases + 3) // 4 + 8 * (n_reads + 1),
                    "d2h_bytes_per_step": (1 << L) // 8 + 4, "ms_per_step": sec_e2e_packed / e2e_steps * 1e3, "workers_per_gpu": n_workers,
                    "how": "kwg_bloom_add_packed (2-bit reads) + kwg_bloom_finalize_crc, pinned host buffers",
                    "ascii_input": {"value": n * kmers * e2e_steps / sec_e2e, "ms_per_step": sec_e2e / e2e_steps * 1e3,
                                    "h2d_bytes_per_step": n_bases + 8 * (n_reads + 1)}},
            "gpu_launches": int(launches),
            "l2_atomics_per_s": kmers * h / t_scan if t_scan > 0 else None,
            "roofline": {"bound": "hbm", "kernel": "kmer_scan_kernel<RAW,3> (bound by L2 atomics and integer issue, not by HBM: see l2_atomics_per_s)",
                         "achieved": alg / t_scan / 1e9 if t_scan > 0 else None, "peak": peak, "unit": "GB/s",
                         "frac": alg / t_scan / 1e9 / peak if t_scan > 0 else None, "traffic": None, "algorithmic_bytes": alg,
                         "peak_source": peak_src, "kernel_ms": t_scan * 1e3},
            "kernel_ms_per_step": {"kmer_scan_raw": t_scan * 1e3}}


# ---------------------------------------------------------------------------------------------- checksums
def stage_crc32(D, args, windows):
    """zlib-compatible crc32 on the device (crc32.cu): the checksum of a 64 MiB filter (.bloom header) and of a 4 GiB
    slice region (.db header)."""
    torch = D.torch
    from kwage_b200 import capi
    peak, peak_src = measured_peaks()
    res = {}
    launches0 = capi.launch_count()
    for name, n_bytes in (("filter_64MiB", 64 << 20), ("slices_4GiB", 4 << 30)):
        buf = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
        capi.synth_filter_bits_dev(31 + D.rank, 0, 1, n_bytes, n_bytes, buf.data_ptr(), device=D.device)
        st = torch.cuda.Stream()
        capi.crc32_dev(buf.data_ptr(), 1, n_bytes, n_bytes, device=D.device, stream=st.cuda_stream)
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record(st)
        for _ in range(reps):
            crc = capi.crc32_dev(buf.data_ptr(), 1, n_bytes, n_bytes, device=D.device, stream=st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        windows.append((w0, time.time()))
        sec = e0.elapsed_time(e1) / 1e3 / reps
        res[name] = {"bytes": n_bytes, "ms": sec * 1e3, "GBps": n_bytes / sec / 1e9, "frac_of_hbm_peak": n_bytes / sec / 1e9 / peak, "crc32": crc}
        del buf
    big = res["slices_4GiB"]
    return {"gpu_launches": int(capi.launch_count() - launches0), "metric": "crc32 bytes/s", "value": D.world * big["bytes"] / (big["ms"] / 1e3), "unit": "bytes/s", "ms_per_step": big["ms"],
            "e2e": {"value": D.world * big["bytes"] / (big["ms"] / 1e3), "unit": "bytes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                    "note": "the message is produced on the device (filter / slices); only the 4-byte checksum leaves"},
            "roofline": {"bound": "hbm", "kernel": "crc_tile_kernel", "achieved": big["GBps"], "peak": peak, "unit": "GB/s",
                         "frac": big["GBps"] / peak, "traffic": None, "algorithmic_bytes": big["bytes"], "peak_source": peak_src},
            "points": res}


# ---------------------------------------------------------------------------------------------- search
def stage_search(D, args, windows):
    import numpy as np
    torch = D.torch
    from kwage_b200 import capi
    F, L, h = args.se_filters, args.se_log2, 3
    nq, qlen = args.se_queries, 1000
    dev = D.device
    row_pitch = (F // 8 + 15) // 16 * 16
    slab = torch.empty((1 << L) * row_pitch, dtype=torch.uint8, device="cuda")
    capi.synth_filter_bits_dev(777 + D.rank, 0, 1, slab.numel(), slab.numel(), slab.data_ptr(), device=dev)
    d_q = torch.empty(nq * qlen + 16, dtype=torch.uint8, device="cuda")
    d_qo = torch.empty(nq + 1, dtype=torch.int64, device="cuda")
    capi.synth_reads_dev(4242, 0, nq, qlen, d_q.data_ptr(), d_qo.data_ptr(), device=dev)     # same queries on every rank
    # known positives: every 10th query carries PLANT bp whose k-mers are inserted into three columns of the slab
    PLANT, plant_cols = 600, sorted({0, 4097 % F, F - 1})
    n_planted = (nq + 9) // 10
    torch.cuda.synchronize()
    for c in plant_cols:
        capi.synth_plant_dev(slab.data_ptr(), row_pitch, K, h, L, c, d_q.data_ptr(), d_qo.data_ptr(), 0, 10, n_planted, PLANT, device=dev)
    count_pitch = (F + 3) // 4 * 4
    d_counts = torch.empty(nq * count_pitch, dtype=torch.int32, device="cuda")
    d_nk = torch.empty(nq, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    db = capi.Database.attach_dev(slab.data_ptr(), row_pitch, K, h, L, F, device=dev)

    def step_dev(i):
        db.search_counts_dev(d_q.data_ptr(), d_qo.data_ptr(), nq, nq * qlen, d_nk.data_ptr(), d_counts.data_ptr(), count_pitch)

    steps, warmup = max(2, min(args.steps, 5)), 3
    for i in range(warmup):
        step_dev(i)
    db.sync()
    db.set_timing(True)
    db.get_timing()
    launches0 = capi.launch_count()
    sec = timed(D, db.stream(), step_dev, steps, 0, windows)
    launches = capi.launch_count() - launches0
    ms, nl = db.get_timing()
    db.set_timing(False)
    n_kmers = int(d_nk.to(torch.int64).sum().item())
    tests = n_kmers * F

    # e2e: host queries in, thresholded hits out, hit lists gathered on rank 0 over NCCL
    h_q = torch.empty(nq * qlen, dtype=torch.uint8).pin_memory()
    h_q.copy_(d_q[: nq * qlen])
    h_qo = torch.empty(nq + 1, dtype=torch.int64).pin_memory()
    h_qo.copy_(d_qo)
    torch.cuda.synchronize()
    qb, qo = h_q.numpy(), h_qo.numpy().view(np.uint64)
    n_hits = [0]

    # the one collective of the path, inside the product: kwg_search_gather = per-slab search + one NCCL exchange of the
    # compacted hit lists (HBM to HBM over NVLink) + merge on rank 0 with global column indices
    comm = None
    if D.enabled:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if D.rank == 0:
            uid = torch.tensor(list(capi.Comm.unique_id()), dtype=torch.uint8, device="cuda")
        D.dist.broadcast(uid, src=0)
        comm = capi.Comm.create(dev, D.world, D.rank, bytes(uid.cpu().numpy().tobytes()))
    h_nk = torch.empty(nq, dtype=torch.int32).pin_memory()

    def step_host(i):
        if comm is not None:
            n = db.search_gather_ptr(comm, h_q.data_ptr(), h_qo.data_ptr(), nq, 0.5, D.rank * F, h_nk.data_ptr(), root=0)
            if D.rank == 0:
                n_hits[0] = n
        else:
            hits, nk = db.search_flat(qb, qo, 0.5)
            n_hits[0] = len(hits)

    # (three untimed steps: NCCL sets up its connections lazily, per kind of operation, over the first calls -- at 8 GPUs the
    # first gather takes 3 s and the second is still slow)
    sec_e2e = timed(D, db.stream(), step_host, steps, 3, windows)
    if comm is not None:
        comm.close()
    db.close()
    del slab, d_counts
    torch.cuda.empty_cache()
    n = D.world
    peak, peak_src = measured_peaks()
    t_k = ms[capi.T_SEARCH] / max(int(nl[capi.T_SEARCH]), 1) / 1e3
    alg_bytes = n_kmers * h * (F // 8) + nq * F * 4
    achieved = alg_bytes / t_k / 1e9 if t_k > 0 else 0.0
    return {"metric": "filter-kmer tests/s", "value": n * tests * steps / sec, "unit": "tests/s", "ms_per_step": sec / steps * 1e3,
            "config": {"filters_per_gpu": F, "log2_filter_len": L, "num_hash": h, "queries": nq, "query_len": qlen, "slab_bytes": (1 << L) * row_pitch,
                       "unique_query_kmers": n_kmers},
            "e2e": {"value": n * tests * steps / sec_e2e, "unit": "tests/s", "h2d_bytes_per_step": nq * qlen + 8 * (nq + 1), "d2h_bytes_per_step": 12 * n_hits[0] + 4 * nq,
                    "threshold": 0.5, "hits": n_hits[0], "hits_expected_at_least": n * n_planted * len(plant_cols),
                    "planted": "%d of %d queries carry %d bp whose k-mers are inserted into columns %s of every slab" % (n_planted, nq, PLANT, plant_cols),
                    "early_exit": "off (KWG_SEARCH_NO_EXIT)" if os.environ.get("KWG_SEARCH_NO_EXIT") else
                                  "kwg_search stops reading the rows of a (query, 4096-column chunk) once no column of it can reach the threshold "
                                  "(the reference's early exit, kwage.cpp:397,459-482); tests/s counts the nominal filter x k-mer pairs; "
                                  "`value` (kwg_search_counts_dev) reads every row"},
            "roofline": {"bound": "hbm", "kernel": "search_count_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC.get("search_count_kernel") if (F, L, nq, qlen) == (8192, 26, 10000, 1000) else None,
                         "traffic_source": NCU_TRAFFIC_SOURCE, "peak_source": peak_src, "kernel_ms": t_k * 1e3, "share_of_step": t_k * 1e3 / (sec / steps * 1e3),
                         "algorithmic_bytes": alg_bytes},
            "kernel_ms_per_step": {"query_kmers": float(ms[capi.T_AUX]) / steps, "search_count": float(ms[capi.T_SEARCH]) / steps},
            "gpu_launches": int(launches)}


# ---------------------------------------------------------------------------------------------- host ingestion
def stage_ingest(D, args, windows):
    """FASTQ.gz -> .bloom wall clock through the C++ host layer (kwage_b200/host: gz parser thread + 2-bit packing +
    kwg_bloom_add_packed + finalize + file write), the path a maestro worker runs per accession (make_bloom.cpp:76-504).
    Host-bound by design of the format: zlib inflates a few hundred MB/s per thread."""
    import gzip
    import numpy as np
    torch = D.torch
    from kwage_b200 import capi, hostapi as H
    from kwage_b200.host import build as hbuild
    hbuild.build()
    n_reads = args.ingest_reads
    n_bases = n_reads * READ_LEN
    kmers = n_reads * (READ_LEN - K + 1)
    d_b = torch.empty(n_bases + 16, dtype=torch.uint8, device="cuda")
    capi.synth_reads_dev(555 + D.rank, 0, n_reads, READ_LEN, d_b.data_ptr(), 0, device=D.device)
    seq = d_b[:n_bases].cpu().numpy().reshape(n_reads, READ_LEN)
    del d_b
    # fixed-width FASTQ records: "@r0000000\n" + bases + "\n+\n" + qualities + "\n"
    rec = np.empty((n_reads, 10 + READ_LEN + 3 + READ_LEN + 1), dtype=np.uint8)
    ids = np.char.zfill(np.arange(n_reads).astype(str), 7)
    rec[:, 0] = ord("@")
    rec[:, 1] = ord("r")
    rec[:, 2:9] = np.frombuffer("".join(ids).encode(), dtype=np.uint8).reshape(n_reads, 7)
    rec[:, 9] = ord("\n")
    rec[:, 10:10 + READ_LEN] = seq
    rec[:, 10 + READ_LEN: 13 + READ_LEN] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, 13 + READ_LEN: 13 + 2 * READ_LEN] = ord("I")
    rec[:, -1] = ord("\n")
    tmp = tempfile.mkdtemp(prefix="kwage_ingest_")
    out = {}
    try:
        fq = os.path.join(tmp, "SRR0000001.fastq.gz")
        with open(fq, "wb") as f:
            f.write(gzip.compress(rec.tobytes(), compresslevel=1))
        reads = os.path.join(tmp, "SRR0000002.reads")
        np.concatenate([seq, np.full((n_reads, 1), 10, dtype=np.uint8)], axis=1).tofile(reads)
        for name, acc, path in (("fastq_gz", "SRR0000001", fq), ("plain_reads", "SRR0000002", reads)):
            H.make_bloom_file(acc, path, n_bases, tmp, k=K, min_kmer_count=1, p=P_FALSE, min_log2=LMIN, max_log2=LMAX, device=D.device)   # warm-up (page cache, allocations)
            t0 = time.time()
            r = H.make_bloom_file(acc, path, n_bases, tmp, k=K, min_kmer_count=1, p=P_FALSE, min_log2=LMIN, max_log2=LMAX, device=D.device)
            sec = time.time() - t0
            windows.append((t0, time.time()))
            if r["status"] != H.STATUS_BLOOM_SUCCESS:
                raise SystemExit("bench ingest: make_bloom_filter failed: %r" % (r,))
            out[name] = {"seconds": sec, "kmer_inserts_per_s": kmers / sec, "input_bytes": os.path.getsize(path), "input_MBps": os.path.getsize(path) / sec / 1e6,
                         "num_kmer": r["num_kmer"], "log2_filter_len": r["log2_len"], "num_hash": r["num_hash"]}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    sec = out["fastq_gz"]["seconds"]
    return {"metric": "FASTQ.gz -> .bloom wall clock", "value": D.world * kmers / sec, "unit": "kmer_inserts/s", "ms_per_step": sec * 1e3,
            "config": {"reads": n_reads, "read_len": READ_LEN, "host_threads": "1 parser/packer + 1 feeder per accession"},
            "e2e": {"value": D.world * kmers / sec, "unit": "kmer_inserts/s", "h2d_bytes_per_step": (n_bases + 3) // 4 + 8 * (n_reads + 1),
                    "d2h_bytes_per_step": (1 << out["fastq_gz"]["log2_filter_len"]) // 8 + 8},
            "points": out,
            "note": "wall clock of make_bloom_filter() incl. gunzip, parsing, 2-bit packing, H2D, kernels, D2H, crc32 and the .bloom file write; "
                    "bound by zlib inflate on the parser thread, not by the device"}


def stage_db_load(D, args, windows):
    """.db file -> HBM slab through SubjectDatabase (two page-locked buffers, four reader threads per piece, flat copies):
    GB/s of the load, beside the rate of one flat pinned H2D copy of the same bytes."""
    import numpy as np
    import struct
    torch = D.torch
    from kwage_b200 import capi, hostapi as H
    from kwage_b200.host import build as hbuild
    hbuild.build()
    F, L = 2048, args.load_log2
    row = F // 8
    n_bytes = (1 << L) * row
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 2 * n_bytes + (1 << 30) else None
    tmp = tempfile.mkdtemp(prefix="kwage_dbload_", dir=base)
    try:
        d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
        capi.synth_filter_bits_dev(4711 + D.rank, 0, 1, n_bytes, n_bytes, d.data_ptr(), device=D.device)
        h_pin = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
        h_pin.copy_(d)
        torch.cuda.synchronize()
        path = os.path.join(tmp, "bench.db")
        # DBFileHeader (kwage.h:30-72, 44 bytes), the slices, an (unused here) metadata location table
        hdr = struct.pack("<IIIIIIIiIQ", 0x20191025, 2, 0, K, 3, L, F, 0, 0, 44 + n_bytes)
        with open(path, "wb") as f:
            f.write(hdr)
            f.write(memoryview(h_pin.numpy()))
            f.write(bytes(8 * F))
        # yardstick: one flat pinned copy of the same bytes
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d.copy_(h_pin, non_blocking=True)
        torch.cuda.synchronize()
        e0.record()
        d.copy_(h_pin, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        pinned_gbs = n_bytes / (e0.elapsed_time(e1) / 1e3) / 1e9
        del d, h_pin
        torch.cuda.empty_cache()
        # yardstick 2: the file read alone (one thread, readinto a page-locked buffer), no device involved
        h_buf = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        mv = memoryview(h_buf.numpy())
        t_r = time.time()
        with open(path, "rb", buffering=0) as f:
            while f.readinto(mv):
                pass
        read_gbs = (n_bytes + 44 + 8 * F) / (time.time() - t_r) / 1e9
        del h_buf, mv
        H.db_load_seconds([path], device=D.device)           # warm-up (page cache)
        D.barrier()
        t0 = time.time()
        sec, _ = H.db_load_seconds([path], device=D.device)
        windows.append((t0, time.time()))
        sec = D.max_over_ranks(sec)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    gbs = n_bytes / sec / 1e9
    return {"metric": ".db -> HBM load bytes/s", "value": D.world * n_bytes / sec, "unit": "bytes/s", "ms_per_step": sec * 1e3,
            "config": {"filters": F, "log2_filter_len": L, "slice_bytes": n_bytes, "file_on": "tmpfs" if base else "tmp"},
            "e2e": {"value": D.world * n_bytes / sec, "unit": "bytes/s", "h2d_bytes_per_step": n_bytes, "d2h_bytes_per_step": 0},
            "GBps": gbs, "pinned_h2d_copy_GBps": pinned_gbs, "frac_of_pinned_copy": gbs / pinned_gbs, "file_read_alone_GBps": read_gbs,
            "note": "wall clock of SubjectDatabase(file): pread by four threads into two alternating page-locked buffers, each piece uploaded with "
                    "kwg_db_upload_rows_async while the next is read"}


def stage_sweep(D, args, windows):
    """BASELINE.json configs[4], bounded grid (bench_sweep.GRIDS['bounded']): k 21..32, 1-7 hashes, filters of 2^20..2^32 bits."""
    import bench_sweep
    t0 = time.time()
    doc = bench_sweep.run(reads=args.sweep_reads, reps=2, grid="bounded", device=D.device, log=open(os.devnull, "w"))
    windows.append((t0, time.time()))
    best = max(doc["raw_construction"], key=lambda r: r["kmer_inserts_per_s"])
    return {"metric": "configs[4] sweep: raw-mode k-mer inserts/s (best point)", "value": best["kmer_inserts_per_s"], "unit": "kmer_inserts/s",
            "ms_per_step": best["ms"], "seconds": time.time() - t0,
            "e2e": {"value": best["kmer_inserts_per_s"], "unit": "kmer_inserts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "device-resident sweep: the host-buffer numbers are in the construct / construct_raw / search stages"},
            **doc}


# ---------------------------------------------------------------------------------------------- CPU reference
def reference_construct(n_proc, reads_per_proc, steps, warmup):
    """The UNMODIFIED reference make_bloom_filter (oracle/_ref/ref_driver), one accession per process,
    n_proc processes at once (= the reference's MPI worker model without MPI).  Returns
    (k-mer occurrences / s aggregate, description, kind)."""
    import numpy as np
    from oracle import oracle_py as O
    kmers = reads_per_proc * (READ_LEN - K + 1)
    tmp = tempfile.mkdtemp(prefix="kwage_ref_")
    try:
        nl = np.full((reads_per_proc, 1), ord("\n"), dtype=np.uint8)
        for p in range(n_proc):
            d = os.path.join(tmp, "p%d" % p)
            os.makedirs(d)
            bases = O.gen_reads(12345 + p, 0, reads_per_proc, READ_LEN).reshape(reads_per_proc, READ_LEN)
            np.concatenate([bases, nl], axis=1).tofile(os.path.join(d, "SRR%06d.reads" % (p + 1)))
        if not O.have_ref():
            # the reference could not travel: time the C restatement instead (single thread)
            t0 = time.time()
            bases = O.gen_reads(12345, 0, reads_per_proc, READ_LEN)
            offsets = np.arange(reads_per_proc + 1, dtype=np.uint64) * np.uint64(READ_LEN)
            O.make_bloom(bases, offsets, K, 1, P_FALSE, LMIN, LMAX, reads_per_proc * READ_LEN)
            return kmers / (time.time() - t0), "oracle port, 1 accession of %d reads" % reads_per_proc, "port", 1, time.time() - t0

        def one_step():
            procs = []
            for p in range(n_proc):
                d = os.path.join(tmp, "p%d" % p)
                env = dict(os.environ, KWAGE_READS_DIR=d)
                procs.append(subprocess.Popen([os.path.join(O.REF_DIR, "ref_driver"), "make_bloom", "SRR%06d" % (p + 1), d, str(K), "1", str(P_FALSE),
                                               str(LMIN), str(LMAX), str(reads_per_proc * READ_LEN)], stdout=subprocess.PIPE, text=True, env=env))
            t0 = time.time()
            outs = [json.loads(p.communicate()[0]) for p in procs]
            wall = time.time() - t0
            assert all(o["status"] == 14 for o in outs), outs
            return wall

        for _ in range(warmup):
            one_step()
        total = sum(one_step() for _ in range(steps))
        rate = n_proc * kmers * steps / total
        desc = "make_bloom_filter (reference, unmodified), %d processes x 1 accession of %d reads x %d bp, k=%d, min count 1, L in [%d,%d]" % (
            n_proc, reads_per_proc, READ_LEN, K, LMIN, LMAX)
        return rate, desc, "reference", n_proc, total / steps
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def ref_procs():
    """as many worker processes as cores, bounded by memory (each allocates ~2.6 GiB + counting table)"""
    cores = cpu_cores()
    try:
        with open("/proc/meminfo") as f:
            avail = [int(l.split()[1]) for l in f if l.startswith("MemAvailable")][0] / 1e6
        cores = max(1, min(cores, int(avail // 4)))
    except Exception:
        pass
    return cores


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--stages", default="construct,construct_c5,construct_raw,ingest,crc32,transpose,db_load,search,sweep")
    ap.add_argument("--ingest-reads", type=int, default=200000, help="reads of the FASTQ.gz -> .bloom wall-clock stage")
    ap.add_argument("--load-log2", type=int, default=23, help="db_load stage: 2048 filters x 2^this slices (23: 2 GiB)")
    ap.add_argument("--sweep-reads", type=int, default=1000000)
    ap.add_argument("--e2e-workers", type=int, default=3, help="host threads (handles) per GPU in the construct e2e arm")
    ap.add_argument("--in-flight", type=int, default=1, help="accessions in flight per GPU (handles, streams) in the device-resident construct arm")
    ap.add_argument("--reads", type=int, default=1000000, help="reads per accession (step)")
    ap.add_argument("--min-kmer-count", type=int, default=1, help="counting-filter threshold (reference default 5; needs --coverage)")
    ap.add_argument("--coverage", type=float, default=0.0, help=">0: reads are sampled from a random genome at this coverage")
    ap.add_argument("--tr-filters", type=int, default=4096)
    ap.add_argument("--tr-log2", type=int, default=26)
    ap.add_argument("--se-filters", type=int, default=8192)
    ap.add_argument("--se-log2", type=int, default=26)
    ap.add_argument("--se-queries", type=int, default=10000)
    ap.add_argument("--cpu-baseline-reads", type=int, default=100000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    stages = [s for s in args.stages.split(",") if s]
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    config = {"workload": "configs[1] Bloom construction: one accession per step = %d synthetic %d bp reads (%.3g k-mer occurrences), "
                          "k=%d, counting filter with min_kmer_count=%d, p=%.2f, L in [%d,%d]; accessions sharded across GPUs%s" % (
                              args.reads, READ_LEN, args.reads * (READ_LEN - K + 1), K, args.min_kmer_count, P_FALSE, LMIN, LMAX,
                              "; reads sampled from a random genome at coverage %g" % args.coverage if args.coverage > 0 else ""),
              "reads_per_accession": args.reads, "read_len": READ_LEN, "kmer_len": K, "min_kmer_count": args.min_kmer_count,
              "parallelism": "accession-per-GPU x%d" % world,
              "l2_policy": "inputs (150 MB reads, 3.8 GB of touch records per accession, 64 GiB slabs) are larger than the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        # bounded sample: the whole run should end within a few minutes
        n_proc = ref_procs()
        budget_s = 150.0 / max(args.steps + args.warmup, 1)
        reads = int(min(args.reads, max(20000, budget_s * 0.9e6 / (READ_LEN - K + 1))))
        rate, desc, kind, cores, step_s = reference_construct(n_proc, reads, args.steps, args.warmup)
        # the workload this arm actually ran: a bounded sample of the GPU arm's accession
        config = dict(config)
        config["workload"] = ("bounded sample of configs[1]: each step = %d processes x 1 accession of %d synthetic %d bp reads (%.3g k-mer occurrences "
                              "each; the GPU arm's accession has %d reads), k=%d, counting filter with min_kmer_count=%d, p=%.2f, L in [%d,%d]; the "
                              "reference's make_bloom_filter is single-threaded, one process per accession on every host core" % (
                                  cores, reads, READ_LEN, reads * (READ_LEN - K + 1), args.reads, K, args.min_kmer_count, P_FALSE, LMIN, LMAX))
        config["reads_per_accession"] = reads
        config["full_workload_reads_per_accession"] = args.reads
        config["parallelism"] = "accession-per-process x%d host processes" % cores
        config["l2_policy"] = "n/a (CPU)"
        line = {"impl": "reference", "metric": "Bloom k-mer inserts/s", "value": rate, "unit": "kmer_inserts/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rate, "unit": "kmer_inserts/s", "cores": cores, "kind": kind, "sample": desc},
                "e2e": {"value": rate, "unit": "kmer_inserts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    D = Dist(args.gpus)
    from kwage_b200 import capi
    if capi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libkwage_cuda has no CPU fallback")
    sampler = ClockSampler(D.device)
    if D.rank == 0:
        sampler.start()
    windows = []
    out = {}
    if "construct" in stages:
        out["construct"] = stage_construct(D, args, windows)
    if "construct_c5" in stages:
        # the reference's DEFAULT threshold (--min-kmer-count 5, options.h) on reads that cover a random genome 30x
        import copy
        a5 = copy.copy(args)
        a5.min_kmer_count, a5.coverage = 5, 30.0
        out["construct_c5"] = stage_construct(D, a5, windows)
        out["construct_c5"]["config"] = {"min_kmer_count": 5, "coverage": 30.0, "reads_per_accession": args.reads}
    if "construct_raw" in stages:
        out["construct_raw"] = stage_construct_raw(D, args, windows)
    if "ingest" in stages:
        out["ingest"] = stage_ingest(D, args, windows)
    if "crc32" in stages:
        out["crc32"] = stage_crc32(D, args, windows)
    if "transpose" in stages:
        out["transpose"] = stage_transpose(D, args, windows)
    if "db_load" in stages:
        out["db_load"] = stage_db_load(D, args, windows)
    if "search" in stages:
        out["search"] = stage_search(D, args, windows)
    if "sweep" in stages and world == 1:
        out["sweep"] = stage_sweep(D, args, windows)          # (one GPU: the sweep is about kernel regimes, not scaling)
    if D.rank == 0:
        sampler.stop()
    cpu = None
    if D.rank == 0 and world == 1 and not args.no_cpu_baseline and "construct" in stages:
        n_proc = ref_procs()
        rate, desc, kind, cores, _ = reference_construct(n_proc, args.cpu_baseline_reads, 1, 0)
        cpu = {"value": rate, "unit": "kmer_inserts/s", "cores": cores, "kind": kind, "sample": desc}
    D.finish()
    if D.rank != 0:
        return 0
    head = out.get("construct") or next(iter(out.values()))
    line = {"metric": "Bloom k-mer inserts/s", "value": head["value"], "unit": "kmer_inserts/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": config, "e2e": head["e2e"], "gpu_launches": head.get("gpu_launches"),
            "roofline": head.get("roofline"), "cpu_baseline": cpu, "clocks": sampler.summary(windows),
            "stages": {k: v for k, v in out.items()}}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
