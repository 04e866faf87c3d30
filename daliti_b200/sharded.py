"""Multi-GPU glue for the spatially sharded map (BASELINE config C4).

Each rank owns the search-cell tiles `tile_owner(tile) == rank` (plus a halo so that every in-range
neighbour of an owned query is local) and evaluates the measurement model only on the query points it
owns; the 158 doubles of partial normal equations (H^T H 144, H^T r 12, effective count, residual sum)
are summed with ONE all-reduce per IEKF iteration -- NCCL over NVLink on the GPU box, gloo in the CPU
tests.  torch.distributed is plumbing only: the partials come out of dlt_measure_dev.
"""
from __future__ import annotations

N_EQ = 158


def allreduce_measure(handle, pose24, do_match: bool, device: str = "cuda", buf=None):
    """dlt_measure on this rank's shard + all-reduce(SUM) of the partial normal equations.

    device="cuda": the partials stay in device memory (dlt_measure_dev writes into a torch tensor on the
    handle's stream, NCCL reduces it in place).  device="cpu": host tensor + gloo (tests)."""
    import torch
    import torch.distributed as dist

    if device == "cpu":
        m = handle.measure(pose24, do_match)
        t = torch.zeros(N_EQ, dtype=torch.float64)
        t[:144] = torch.from_numpy(m.HtH.reshape(-1).copy())
        t[144:156] = torch.from_numpy(m.Htr.copy())
        t[156] = float(m.effct_feat_num)
        t[157] = m.total_residual
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        r = t.numpy()
    else:
        if buf is None:
            buf = torch.zeros(256, dtype=torch.float64, device=device)
        # the handle launches on its own non-blocking stream unless told otherwise: put k_residual on the stream NCCL and the
        # read-back below are ordered on, or the all-reduce could read the buffer before the partial sums are written
        cur = torch.cuda.current_stream().cuda_stream
        if cur:
            handle.set_stream(cur)
        handle.measure_dev(pose24, do_match, buf.data_ptr())
        if not cur:  # torch's legacy default stream has no handle to adopt: wait for the handle's own stream instead
            handle.sync()
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(buf[:N_EQ], op=dist.ReduceOp.SUM)
        r = buf[:N_EQ].cpu().numpy()
    return dict(HtH=r[:144].reshape(12, 12).copy(), Htr=r[144:156].copy(), effct_feat_num=int(round(r[156])), total_residual=float(r[157]))


def attach_peers(obj):
    """Collective: exchange the peer-mailbox blobs of all ranks over torch.distributed (any backend -- it is 128 bytes per
    rank, once) and attach them, so that from here on the sums over the ranks run inside the kernels over NVLink peer
    memory (dlt_peer_attach) and torch.distributed is no longer on the data path.  obj: ScanToMap or LaserMapping of this
    rank (shard_rank == dist rank, shard_count == world size).  Returns True when attached on EVERY rank; False (nothing
    attached anywhere) when some rank could not map a peer -- the caller then stays on the all-reduce callback."""
    import torch.distributed as dist

    from .binding import DltError

    world = dist.get_world_size()
    blobs = [None] * world
    dist.all_gather_object(blobs, obj.peer_export())
    # every rank first checks that it CAN map its peers before anybody switches over: a half-attached job would wait forever
    ok = True
    try:
        obj.peer_attach(blobs)
    except DltError:
        ok = False
    oks = [None] * world
    dist.all_gather_object(oks, ok)
    if all(oks):
        return True
    if ok:
        obj.peer_detach()
    return False


def attach_native_nccl(lm, device: int):
    """Collective: rank 0 draws an NCCL unique id, torch.distributed carries its 128 bytes to the other ranks (once; any
    backend), every rank creates the library's own communicator and binds it to `lm` (dlt_nccl_attach).  From then on the
    sums over the ranks are plain ncclAllReduce calls issued from C on the handle's stream."""
    import torch.distributed as dist

    box = [lm.native_nccl_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    lm.attach_native_nccl(box[0], dist.get_rank(), dist.get_world_size(), device)
