"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (column slabs, accession shards,
hit gather).  The per-slab hit lists come from the ORACLE here (no GPU): the point is the plumbing."""
import os
import socket

import numpy as np
import pytest

from kwage_b200 import sharding
from kwage_b200.capi import HIT_DTYPE


def test_column_slabs_cover_and_align():
    for n, w in [(65536, 8), (4096, 4), (257, 2), (2048, 8), (100, 3), (1, 2), (1000, 8), (129, 2)]:
        slabs = sharding.column_slabs(n, w)
        assert len(slabs) == w and slabs[0][0] == 0 and slabs[-1][1] == n
        for (a, b), (c, d) in zip(slabs[:-1], slabs[1:]):
            assert b == c and a <= b
        for a, b in slabs:
            if b > a:                      # ranks beyond the last aligned unit own nothing
                assert a % 128 == 0
    assert sharding.column_slabs(65536, 8)[3] == (3 * 8192, 4 * 8192)


def test_accession_shard_round_robin():
    seen = sorted(a for r in range(3) for a in sharding.accession_shard(10, r, 3))
    assert seen == list(range(10))
    assert sharding.accession_shard(10, 1, 3) == [1, 4, 7]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import oracle_py as O
    import synth_cases as S
    import util
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = S.SEARCH_CASES["random_n257"]
        dbd = util.search_case_db("random_n257")
        queries = [s for _, s in S.search_queries(case)]
        slabs = sharding.column_slabs(dbd["n"], world)
        a, b = slabs[rank]
        t = 0.01
        # what the GPU of this rank would return for its slab: hits with slab-local filter indices
        local = []
        for qi, seq in enumerate(queries):
            hf, hm, _ = O.search_matches(dbd["slices"], dbd["n"], dbd["L"], dbd["h"], dbd["k"], seq, t)
            local += [(qi, int(f) - a, int(m)) for f, m in zip(hf, hm) if a <= f < b]
        hits = np.array(local, dtype=HIT_DTYPE) if local else np.zeros(0, dtype=HIT_DTYPE)
        merged = sharding.gather_hits(hits, a, dist, dst=0)
        if rank == 0:
            exp = []
            for qi, seq in enumerate(queries):
                hf, hm, _ = O.search_matches(dbd["slices"], dbd["n"], dbd["L"], dbd["h"], dbd["k"], seq, t)
                exp += [(qi, int(f), int(m)) for f, m in zip(hf, hm)]
            got = [(int(x["query"]), int(x["filter"]), int(x["num_match"])) for x in merged]
            q.put(("ok", got == exp, len(exp)))
        else:
            q.put(("other", merged is None, 0))
    finally:
        dist.destroy_process_group()


def test_gather_hits_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for tag, ok, n in res:
        assert ok, tag
    assert max(n for _, _, n in res) > 100


def test_gather_hits_single_process():
    hits = np.array([(0, 2, 4), (0, 7, 3), (1, 5, 9)], dtype=HIT_DTYPE)         # kwg_search's order: (query, filter)
    out = sharding.gather_hits(hits, 128)
    assert [(int(x["query"]), int(x["filter"])) for x in out] == [(0, 130), (0, 135), (1, 133)]


def test_merge_hits_is_the_order_of_one_search_over_all_columns():
    # kwg_merge_hits (the root's step of kwg_search_gather): per-slab lists, each in (query, filter) order, slabs in column
    # order -> the list one search over the whole database returns (reference: kwage.cpp:154-177 merges thread-local maps)
    from kwage_b200 import capi
    rng = np.random.default_rng(5)
    n_q, slabs = 40, [(0, 128), (128, 384), (384, 400), (400, 1000)]
    whole = sorted({(int(rng.integers(0, n_q)), int(rng.integers(0, 1000))) for _ in range(3000)})
    whole = np.array([(q, f, (q * 31 + f) % 97) for q, f in whole], dtype=HIT_DTYPE)
    lists = [whole[(whole["filter"] >= a) & (whole["filter"] < b)] for a, b in slabs]
    assert all(len(x) for x in lists)
    merged = capi.merge_hits(lists, n_q)
    assert np.array_equal(merged, whole)
    assert len(capi.merge_hits([np.zeros(0, HIT_DTYPE), np.zeros(0, HIT_DTYPE)], 5)) == 0
