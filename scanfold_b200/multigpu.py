"""Window-range sharding of one record over the ranks of a torch.distributed job (one process per GPU).

The scan itself needs no communication: windows are independent, the record is replicated, and the device
shuffles are keyed by absolute window index, so any sharding folds the same sequences.  The ScanFold-Fold
accumulators overlap only at shard boundaries: the last W - step nucleotides a shard touches are also covered
by the first windows of the next shard.  Each rank therefore sends those halo rows (integer counts / exact split
sums, so the merge is bit exact) to its right neighbour with one NCCL send/recv, compacts the nucleotides it
owns, and rank 0 gathers the compact per-nucleotide partner lists for the (cheap) argmin / competition / writers.
"""
import numpy as np

from . import foldstep


def shard_windows(total_windows, world, rank):
    """contiguous window range [w0, w1) of `rank`"""
    return total_windows * rank // world, total_windows * (rank + 1) // world


def exchange_halo(acc, W, step, rank, world, dist):
    """Send the rows this shard shares with its right neighbour, merge the rows the left neighbour shares with us.
    `acc` has export_tensors(row0, n_rows) -> (count, first_seen, sums) and merge_tensors(row0, n_rows, ...), on
    whatever device the process group moves (CUDA for NCCL, CPU for gloo).  Returns the number of owned rows."""
    halo = max(W - step, 0)
    own = acc.n_nt - halo if rank < world - 1 else acc.n_nt
    if world == 1 or halo == 0:
        return own
    if rank < world - 1 and own < 0:
        raise ValueError("shard smaller than the window overlap: fewer GPUs or a longer record needed")
    ops, recv = [], None
    if rank < world - 1:
        send = acc.export_tensors(own, halo)
        ops += [dist.P2POp(dist.isend, t, rank + 1) for t in send]
    if rank > 0:
        if acc.n_nt < halo:
            raise ValueError("shard smaller than the window overlap: fewer GPUs or a longer record needed")
        recv = acc.empty_tensors(halo)
        ops += [dist.P2POp(dist.irecv, t, rank - 1) for t in recv]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    if recv is not None:
        acc.merge_tensors(0, halo, *recv)
    return own


def gather_tables(table, rank, world, dist):
    """per-rank PartnerTable over consecutive nucleotide ranges -> the whole table on rank 0 (None elsewhere)"""
    if world == 1:
        return table
    payload = (table.nt_ptr, table.partner, table.count, table.first_seen, table.sums)
    out = [None] * world if rank == 0 else None
    dist.gather_object(payload, out, dst=0)
    if rank != 0:
        return None
    return foldstep.concat_tables([foldstep.PartnerTable(*p) for p in out])


def partner_table_distributed(acc, W, step, rank, world, dist):
    """halo exchange + compaction of the owned nucleotides + gather on rank 0"""
    own = exchange_halo(acc, W, step, rank, world, dist)
    table = foldstep.table_from_compact(*acc.compact(0, own))
    return gather_tables(table, rank, world, dist)


def gather_window_tables(tables, rank, world, dist):
    """per-rank scan.WindowTable shards -> concatenated arrays on rank 0"""
    from .scan import WindowTable
    if world == 1:
        return tables
    t = tables
    payload = {k: getattr(t, k) for k in ("start1", "end1", "mfe_dcal", "mfe", "z", "p", "ed", "pair_tbl", "centroid_tbl",
                                          "native_unconstrained_dcal", "shuffle_dcal", "final")}
    payload.update(W=t.W, step=t.step, r=t.r, first_window=t.first_window)
    out = [None] * world if rank == 0 else None
    dist.gather_object(payload, out, dst=0)
    if rank != 0:
        return None
    res = WindowTable()
    res.W, res.step, res.r, res.first_window = out[0]["W"], out[0]["step"], out[0]["r"], out[0]["first_window"]
    for k in ("start1", "end1", "mfe_dcal", "mfe", "z", "p", "ed", "pair_tbl", "centroid_tbl",
              "native_unconstrained_dcal", "shuffle_dcal"):
        setattr(res, k, np.concatenate([o[k] for o in out]))
    res.final = out[-1]["final"]
    return res
