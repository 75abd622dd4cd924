"""Sharding of the graph build across the GPUs of one box (torch.distributed: NCCL over NVLink
on the GPU box, gloo in the CPU tests).  There is no collective inside a distance sweep.

One-sided sweeps (epsilon count / fill, queries, small or long-list kNN): every output row
depends on that row and the whole table only, so rank g owns the query rows of `row_range`
against all N columns -- no cross-rank merge, tie semantics untouched -- and the
result shards are all-gathered.

Symmetric sweeps (graph.hamming_knn_graph / hamming_eps_graph on large tables): the triangle of
unordered pairs is cut into bands of stream rows holding equal numbers of pair evaluations
(pg_knn_sym_band); rank g sweeps every row block against its band.  kNN: every rank then holds
candidate lists for ALL rows, and the exchange step is an all-to-all -- rank g receives every
rank's lists of ITS rows (`exchange_lists`), merges them (pg_knn_lists_merge) and the merged
8-byte keys are all-gathered.  Epsilon: the ranks' edge-key buffers are all-gathered and sorted.

`agree` makes rank-local failures collective: a rank that cannot go on (unpackable shard, edge
budget exceeded) must not leave its peers waiting in the next collective.
"""
import torch
import torch.distributed as dist


def rank_world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


ROW_ALIGN = 512   # packed tables are tiled in 512-row units (pg_packed_rows)


def rows_per_rank(n, world):
    """Rows per rank; blocks of large tables start on packed-tile boundaries so that the
    packed shards can be all-gathered straight into one table."""
    per = -(-n // world)
    aligned = -(-per // ROW_ALIGN) * ROW_ALIGN
    if (world - 1) * aligned < n:        # every rank keeps at least one row
        return aligned
    return per


def is_tile_aligned(n, world):
    return rows_per_rank(n, world) % ROW_ALIGN == 0


def row_range(n, rank, world):
    """(first row, number of rows) of this rank's block.  Fewer rows than ranks: every
    rank builds everything (nothing worth sharding).  Large tables: tile-aligned blocks of
    rows_per_rank rows; small ones: a balanced split."""
    if world <= 1 or n < world:
        return 0, n
    per = rows_per_rank(n, world)
    if per % ROW_ALIGN == 0:
        r0 = min(n, rank * per)
        return r0, min(per, n - r0)
    r0, r1 = rank * n // world, (rank + 1) * n // world
    return r0, r1 - r0


def _sharded(n, world):
    return world > 1 and n >= world


def _all_gather_padded(t, cap, world, group):
    """t: (rows <= cap, ...) from every rank -> (world, cap, ...) with zero padding."""
    shape = (cap,) + tuple(t.shape[1:])
    mine = torch.zeros(shape, dtype=t.dtype, device=t.device)
    mine[: t.shape[0]] = t
    out = torch.empty((world * cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.reshape((world, cap) + tuple(t.shape[1:]))


def _gather_blocks(t, n, world, group):
    """Row blocks (one per rank, sizes from row_range) -> the full (n, ...) array."""
    sizes = [row_range(n, r, world)[1] for r in range(world)]
    allb = _all_gather_padded(t.contiguous(), max(sizes), world, group)
    if all(s == sizes[0] for s in sizes):
        return allb.reshape((world * sizes[0],) + tuple(t.shape[1:]))
    return torch.cat([allb[r, : sizes[r]] for r in range(world)])


def gather_rows(part, n, rank, world, group, eng=None):
    """Fixed-width per-row results (kNN idx / weights): all-gather the row blocks."""
    if not _sharded(n, world):
        return part
    return tuple(_gather_blocks(t, n, world, group) for t in part)


def all_gather_stack(t, world, group):
    """Same-shaped tensor from every rank -> (world, *t.shape)."""
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out.reshape((world,) + tuple(t.shape))


def exchange_lists(lists, n, rank, world, group):
    """Candidate lists of ALL n rows on every rank, (n, k1) -> (world, my rows, k1): list s of my
    row r as found by rank s (one all-to-all; rank g receives only its row block's lists)."""
    sizes = [row_range(n, r, world)[1] for r in range(world)]
    mine = sizes[rank]
    out = torch.empty((world * mine,) + tuple(lists.shape[1:]), dtype=lists.dtype, device=lists.device)
    dist.all_to_all_single(out, lists.contiguous(), output_split_sizes=[mine] * world, input_split_sizes=sizes,
                           group=group)
    return out.reshape((world, mine) + tuple(lists.shape[1:]))


OK, UNPACKABLE, UNSUPPORTED, NO_MEMORY = 0, 1, 2, 3


def agree(code, world, group, device):
    """MAX over the ranks of a small status code: every rank learns whether any rank failed a
    rank-local step and all of them take the same branch before the next collective."""
    if world <= 1 or not (dist.is_available() and dist.is_initialized()):
        return int(code)
    t = torch.tensor([int(code)], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def gather_edge_keys(keys, edges, world, group):
    """Per-rank edge-key buffers of the symmetric epsilon sweep (ragged) -> one buffer holding all
    of them (padding = the sentinel -1, which sorts behind every edge) and the total edge count."""
    meta = torch.tensor([keys.numel(), edges], dtype=torch.int64, device=keys.device)
    allmeta = torch.empty((world * 2,), dtype=torch.int64, device=keys.device)
    dist.all_gather_into_tensor(allmeta, meta, group=group)
    allmeta = allmeta.reshape(world, 2).cpu()
    cap = max(int(allmeta[:, 0].max()), 1)
    mine = torch.full((cap,), -1, dtype=keys.dtype, device=keys.device)
    mine[: keys.numel()] = keys
    out = torch.empty((world * cap,), dtype=keys.dtype, device=keys.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out, int(allmeta[:, 1].sum())


def gather_csr(part, n, rank, world, group, eng=None):
    """Ragged per-row results (epsilon graph): exchange the row counts, then the padded
    index / weight shards, and rebuild one global CSR on every rank."""
    if not _sharded(n, world):
        return part
    indptr, idx, w = part
    counts = _gather_blocks(indptr[1:] - indptr[:-1], n, world, group)
    nnz_local = torch.tensor([idx.numel()], dtype=torch.int64, device=idx.device)
    nnz_all = torch.empty(world, dtype=torch.int64, device=idx.device)
    dist.all_gather_into_tensor(nnz_all, nnz_local, group=group)
    nnz_host = [int(x) for x in nnz_all.cpu()]
    cap = max(max(nnz_host), 1)
    idx_all = _all_gather_padded(idx.contiguous(), cap, world, group)
    w_all = _all_gather_padded(w.contiguous(), cap, world, group)
    idx_out = torch.cat([idx_all[r, : nnz_host[r]] for r in range(world)])
    w_out = torch.cat([w_all[r, : nnz_host[r]] for r in range(world)])
    indptr_out = torch.zeros(n + 1, dtype=torch.int64, device=counts.device)
    torch.cumsum(counts, 0, out=indptr_out[1:])
    return indptr_out, idx_out, w_out
