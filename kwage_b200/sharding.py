"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed).

The path shards without any data-path collective except one: construction by accession, transposition
and search by filter-column slab; search gathers the per-slab hit lists on one rank.
On the GPUs that exchange is the library's own kwg_search_gather (NCCL inside the C ABI, capi.Database.search_gather);
gather_hits below is the same exchange over any torch.distributed backend (gloo in the CPU tests) with the library's
merge step (kwg_merge_hits)."""
import numpy as np

from .capi import HIT_DTYPE


def accession_shard(n_accessions, rank, world):
    """Accessions are independent filters (reference: one accession per MPI worker): round-robin."""
    return list(range(rank, n_accessions, world))


def column_slabs(n_filters, world, align=128):
    """Split [0, n_filters) into `world` contiguous column ranges whose boundaries are multiples of
    `align` columns (16-byte aligned slab rows; at least a multiple of 8 is required by kwg_db_load).
    Ranks at the end may own an empty range when there are fewer than world*align columns."""
    if align % 8:
        raise ValueError("align must be a multiple of 8")
    units = -(-n_filters // align)
    base, extra = divmod(units, world)
    out, begin = [], 0
    for r in range(world):
        u = base + (1 if r < extra else 0)
        end = min(n_filters, begin + u * align)
        out.append((begin, end))
        begin = end
    return out


def gather_hits(hits, col_begin, dist=None, dst=0, device="cpu"):
    """Every rank passes the kwg_hit_t array of ITS slab (filter indices local to the slab) and the
    slab's first column.  Returns, on rank `dst`, the merged hit list with global filter indices
    ordered by (query, filter); None elsewhere.  One gather of sizes + one gather of payload."""
    import torch
    local = np.zeros((len(hits), 3), dtype=np.int64)
    if len(hits):
        local[:, 0] = hits["query"]
        local[:, 1] = hits["filter"].astype(np.int64) + col_begin
        local[:, 2] = hits["num_match"]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        merged = local
    else:
        world, rank = dist.get_world_size(), dist.get_rank()
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
        sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(sizes, n)
        cap = max(1, max(int(s.item()) for s in sizes))
        pad = torch.zeros((cap, 3), dtype=torch.int64, device=device)
        if local.shape[0]:
            pad[: local.shape[0]] = torch.from_numpy(local).to(device)
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst)
        if rank != dst:
            return None
        merged = np.concatenate([b[: int(s.item())].cpu().numpy() for b, s in zip(bufs, sizes)], axis=0)
    out = np.zeros(merged.shape[0], dtype=HIT_DTYPE)
    out["query"], out["filter"], out["num_match"] = merged[:, 0], merged[:, 1], merged[:, 2]
    # rank-major concatenation of (query, filter)-ordered slab lists -> (query, filter) order: the library's merge
    from .capi import merge_hits
    n_queries = int(out["query"].max()) + 1 if len(out) else 0
    return merge_hits([out], n_queries)
