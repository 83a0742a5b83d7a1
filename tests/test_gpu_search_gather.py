"""kwg_search batching / device-side hit lists on one GPU, and the NCCL gather (kwg_search_gather) when the box has two."""
import threading

import numpy as np
import pytest

import synth_cases as S
import util
from kwage_b200 import capi, sharding
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


def _case():
    dbd = util.search_case_db("random_n257")
    seqs = [s for _, s in S.search_queries(S.SEARCH_CASES["random_n257"])]
    return dbd, seqs


def _oracle_hits(dbd, seqs, t):
    exp = []
    for qi, seq in enumerate(seqs):
        hf, hm, _ = O.search_matches(dbd["slices"], dbd["n"], dbd["L"], dbd["h"], dbd["k"], seq, t)
        exp += [(qi, int(f), int(m)) for f, m in zip(hf, hm)]
    return exp


@pytest.mark.parametrize("budget", [4096, 1 << 14, 1 << 30])
def test_search_in_query_batches_equals_one_batch(budget):
    # kwg_search holds at most `budget` bytes of per-(query, filter) counts at a time
    dbd, seqs = _case()
    seqs = seqs * 3
    with capi.Database.load(dbd["slices"], dbd["k"], dbd["h"], dbd["L"], dbd["n"]) as db:
        db.set_count_budget(budget)
        for t in (0.01, 1.0):
            hits, nk = db.search(seqs, t)
            got = [(int(x["query"]), int(x["filter"]), int(x["num_match"])) for x in hits]
            assert got == _oracle_hits(dbd, seqs, t)
            assert list(nk) == [O.search_counts(dbd["slices"], dbd["n"], dbd["L"], dbd["h"], dbd["k"], s)[1] for s in seqs]


def test_search_hits_dev_keeps_the_list_in_hbm():
    import torch
    dbd, seqs = _case()
    bases, offsets = capi.flatten(seqs)
    with capi.Database.load(dbd["slices"], dbd["k"], dbd["h"], dbd["L"], dbd["n"], col_begin=64, col_end=200) as db:
        ptr, n, nk = db.search_hits_dev(bases, offsets, 0.01, filter0=64)
        exp = [h for h in _oracle_hits(dbd, seqs, 0.01) if 64 <= h[1] < 200]
        assert n == len(exp) and ptr
        buf = torch.empty(n * 3, dtype=torch.int32, device="cuda")
        import ctypes as C
        cudart = C.CDLL("libcudart.so")
        assert cudart.cudaMemcpy(C.c_void_p(buf.data_ptr()), C.c_void_p(ptr), C.c_size_t(n * 12), 3) == 0
        got = buf.cpu().numpy().view(np.uint32).reshape(n, 3)
        assert [tuple(int(v) for v in r) for r in got] == exp


def test_search_gather_two_slabs_on_two_devices():
    if capi.device_count() < 2:
        pytest.skip("needs two devices (run with gpurun --gpus 2)")
    dbd, seqs = _case()
    bases, offsets = capi.flatten(seqs)
    slabs = sharding.column_slabs(dbd["n"], 2, align=8)
    comms = capi.Comm.create_all([0, 1])
    out = [None, None]
    err = []

    def run(r):
        try:
            a, z = slabs[r]
            with capi.Database.load(dbd["slices"], dbd["k"], dbd["h"], dbd["L"], dbd["n"], device=r, col_begin=a, col_end=z) as db:
                for t in (0.01, 1.0):
                    out[r] = db.search_gather(comms[r], bases, offsets, t, a, root=0)
                    if r == 0:
                        got = [(int(x["query"]), int(x["filter"]), int(x["num_match"])) for x in out[0][0]]
                        assert got == _oracle_hits(dbd, seqs, t), t
                    else:
                        assert len(out[1][0]) == 0
        except BaseException as e:       # noqa: BLE001 - reported by the main thread
            err.append(e)

    ts = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for c in comms:
        c.close()
    assert not err, err
