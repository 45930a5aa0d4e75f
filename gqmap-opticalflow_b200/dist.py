"""Multi-process plumbing (one process per GPU, torch.distributed for the rendezvous): how the two natural splits of the
path are laid over ranks.  torch.distributed is plumbing only: the data path of a row-band run is
libqgmap's own peer-memory publish kernel (qgmap_band_p2p_connect) or NCCL send/recv + all-reduce issued by libqgmap.so
itself (qgmap_band_connect); independent frame pairs need no collective.
"""
import ctypes as C

import numpy as np

from ._lib import check, lib


def shard_pairs(n_pairs, rank, world):
    """Indices of the frame pairs rank `rank` owns (contiguous blocks, sizes differ by at most one)."""
    lo = n_pairs * rank // world
    hi = n_pairs * (rank + 1) // world
    return list(range(lo, hi))


def band_rows(M, rank, world):
    """Rows [row_begin,row_end) of an M-row belief grid owned by `rank` (same rule as qgmap_group_create)."""
    if M // world < 2:
        raise ValueError("%d bands over %d rows: need at least 2 rows per band" % (world, M))
    return M * rank // world, M * (rank + 1) // world


def new_unique_id():
    """128-byte NCCL unique id (rank 0 creates it; works without a GPU)."""
    buf = C.create_string_buffer(128)
    check(lib.qgmap_band_unique_id(buf))
    return bytes(buf.raw)


def broadcast_unique_id(dist, device=None):
    """Rank 0 creates the NCCL id, every rank receives it through torch.distributed (gloo or nccl backend)."""
    import torch
    t = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if dist.get_rank() == 0:
        t.copy_(torch.frombuffer(bytearray(new_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def connect_band(solver, dist, device=None, transport=None):
    """Attach a Solver created with options.row_begin/row_end = band_rows(M, rank, world) to its neighbours.
    transport 'p2p' (default; env QGMAP_BAND_TRANSPORT): libqgmap's own publish kernel stores the boundary rows and partial sums
    into the peers' memory over NVLink (CUDA IPC mappings); 'nccl': ncclSend/ncclRecv + ncclAllReduce on libqgmap's communicator.
    torch.distributed only carries the rendezvous data (IPC handles / NCCL id)."""
    import os
    transport = transport or os.environ.get("QGMAP_BAND_TRANSPORT", "p2p")
    rank, world = dist.get_rank(), dist.get_world_size()
    if transport == "nccl":
        uid = broadcast_unique_id(dist, device)
        buf = C.create_string_buffer(uid, 128)
        check(lib.qgmap_band_connect(solver._h, rank, world, buf), solver._h)
    elif transport == "p2p":
        import torch
        from ._lib import QGMAP_P2P_BLOB_BYTES as NB
        blob = C.create_string_buffer(NB)
        check(lib.qgmap_band_p2p_export(solver._h, blob), solver._h)
        dev = device if device is not None else "cpu"
        mine = torch.frombuffer(bytearray(blob.raw), dtype=torch.uint8).to(dev)
        parts = [torch.zeros(NB, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(parts, mine)
        allb = C.create_string_buffer(b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts), NB * world)
        check(lib.qgmap_band_p2p_connect(solver._h, rank, world, allb), solver._h)
    else:
        raise ValueError("transport must be 'p2p' or 'nccl'")
    dist.barrier()
    return transport


def assemble_bands(dist, state, M):
    """Every rank's get_state() fills only the rows it owns; gather the owned row blocks into full arrays on all ranks."""
    rank, world = dist.get_rank(), dist.get_world_size()
    rb, re = band_rows(M, rank, world)
    mine = {k: np.ascontiguousarray(v[rb:re]) for k, v in state.items() if isinstance(v, np.ndarray) and v.ndim >= 3}
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    out = dict(state)
    for k in mine:
        out[k] = np.asfortranarray(np.concatenate([p[k] for p in parts], axis=0))
    return out


def gqmap_gpu_mixture_bands(options, I1, I2, dist, device=None, variant="full", transport=None):
    """[mu,sigma,alpha,AEPE,Energy,logP] = gqmap_gpu_mixture(options,I1,I2) (gqmap_gpu_mixture.m:1) over the ranks of a
    torch.distributed job, one row band of the frame pair per rank -- the multi-process twin of options.devices.
    Every rank passes the same options and frames (host arrays) and calls this collectively.  mu and sigma come back with the rows
    this rank owns filled in (assemble_bands gathers them); alpha, AEPE, Energy, logP are complete on every rank.
    The loop is the reference's: iterations up to and including the next monitored one (:52: it == 1 or mod(it,300) == 0), then
    the monitoring block -- each rank extracts the MAP of its own rows on its GPU and contributes its share of logP and of the AEPE
    sum; only those two scalars cross ranks (one small all-reduce per monitored iteration)."""
    import torch
    from . import host
    rank, world = dist.get_rank(), dist.get_world_size()
    its = int(host._opt(options, "its", required=True))
    every = int(host._opt(options, "log_every", 300))
    sup = variant == "super"
    Mo, No = np.asarray(I1).shape
    M = Mo // 4 if sup else Mo
    opts = dict(options) if isinstance(options, dict) else {k: getattr(options, k) for k in dir(options) if not k.startswith("_")}
    rb, re = band_rows(M, rank, world)
    opts.update(row_begin=rb, row_end=re)
    s = host.Solver(opts, I1, I2, variant=variant)
    try:
        if world > 1:
            connect_band(s, dist, device=device, transport=transport)
        init = host._opt(options, "init")
        if init is not None:
            s.set_state(init)
        else:
            s.init_state(int(host._opt(options, "seed", 0)))
        tflow = host._opt(options, "trueFlow")
        if tflow is not None:
            s.set_truth(tflow, host._opt(options, "unknownIdx"))
        AEPE = np.full(its, np.nan)
        Energy = np.zeros(its)
        logP = np.full(its, np.nan)
        b = 4 if sup else 1
        red = torch.zeros(2, dtype=torch.float64, device=device if device is not None else "cpu")
        it, stopped = 1, False
        while not stopped and it <= its:
            nxt = 1 if it == 1 else -(-it // every) * every
            n = min(nxt, its) - it + 1
            r = s.step(n, its=its)
            Energy[it - 1:it - 1 + r["n_done"]] = r["Energy"]
            it += r["n_done"]
            stopped = r["stopped"]
            last = it - 1
            if r["n_done"] > 0 and (last == 1 or last % every == 0):
                lp, ae = s.monitor_partial(aepe=tflow is not None)
                red[0], red[1] = lp, ae
                if world > 1:
                    dist.all_reduce(red)
                logP[last - 1] = float(red[0])
                if tflow is not None:
                    AEPE[last - 1] = float(red[1]) / ((Mo - 2 * b) * (No - 2 * b))
            if r["n_done"] < n:
                break
        st = s.get_state()
    finally:
        s.close()
    L = s.L
    mu = np.stack([st["muu"], st["muv"]], axis=3)
    sigma = np.stack([st["sigmau"], st["sigmav"]], axis=3)
    return (np.asfortranarray(mu), np.asfortranarray(sigma), st["alpha"].reshape(1, 1, L), AEPE.reshape(its, 1), Energy.reshape(its, 1),
            logP.reshape(its, 1))
