"""Slab decomposition of one grid across the GPUs of a node (BASELINE config 4).

One process per GPU (torchrun): physical fields are split by rows, spectral fields by columns in the
conjugate-symmetric ownership of csrc/common.cuh (struct Grid), and every 2-D transform is
local pass -> NCCL all-to-all over NVLink -> local pass, inside the C-ABI library.  torch.distributed is only
the plumbing that shares the ncclUniqueId and gathers results for the user.

    import torch.distributed as dist
    from niwqg_b200 import slab, CoupledModel
    dist.init_process_group("nccl")                       # or "gloo": only used to broadcast 128 bytes
    m = slab.make_model(CoupledModel.Model, nx=8192, ...)  # every rank, same kwargs
    m.set_q(q_global); m.set_phi(phi_global)               # each rank keeps its rows
    m.run()
    q = slab.gather_rows(m, m.q)                           # whole-grid array on every rank
"""
import numpy as np

from . import _native as nat


def rows_of(rank, nranks, nx):
    """Row range [lo, hi) of the physical fields owned by `rank`."""
    n = nx // nranks
    return rank * n, (rank + 1) * n


def share_unique_id(dist, src=0):
    """Rank `src` creates the ncclUniqueId; everybody gets the 128 bytes through torch.distributed."""
    import torch
    rank = dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if rank == src:
        t = torch.tensor(list(nat.nccl_unique_id()), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def make_model(model_cls, dist=None, device=None, **kw):
    """Construct `model_cls(**kw)` as this process's slab of the grid (all ranks call it with the same kwargs)."""
    if dist is None:
        import torch.distributed as dist
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = torch.cuda.current_device()
    if world == 1:
        return model_cls(device=device, **kw)
    m = model_cls(device=device, rank=rank, nranks=world, nccl_id=share_unique_id(dist), **kw)
    import os
    if os.environ.get("NIWQG_SLAB_NCCL", "0") != "1":
        enable_peer_exchange(m, dist)
    return m


def enable_peer_exchange(model, dist):
    """All-gather the CUDA-IPC handles of every rank's receive buffers and map them: the transposes of the distributed
    FFT are then stores into peer memory over NVLink, fused into the FFT pass kernels (no ncclSend/ncclRecv)."""
    import torch
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    ok = 1
    try:
        mine = torch.tensor(list(model._h.ipc_export()), dtype=torch.uint8, device=dev)
    except RuntimeError:
        ok, mine = 0, torch.zeros(nat.Handle.IPC_BYTES, dtype=torch.uint8, device=dev)
    parts = [torch.empty_like(mine) for _ in range(model.nranks)]
    dist.all_gather(parts, mine)
    if ok:
        try:
            model._h.ipc_import(b"".join(bytes(p.cpu().tolist()) for p in parts))
        except RuntimeError:
            ok = 0
    # every rank must use the same exchange: if anybody could not map its peers, all fall back to NCCL all-to-all
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        import logging
        logging.getLogger(__name__).warning("CUDA IPC peer mapping unavailable: slab transposes use NCCL all-to-all")
        model._h.ipc_disable()
    dist.barrier()


def gather_rows(model, local, dist=None):
    """All-gather a row-split physical field (the rank's (nx/P, nx) block) into the (nx, nx) array."""
    if model.nranks == 1:
        return local
    if dist is None:
        import torch.distributed as dist
    import torch
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    loc = np.ascontiguousarray(local)
    t = torch.from_numpy(loc.view(np.float64).reshape(-1)).to(dev)
    parts = [torch.empty_like(t) for _ in range(model.nranks)]
    dist.all_gather(parts, t)
    out = np.concatenate([p.cpu().numpy().view(loc.dtype).reshape(loc.shape) for p in parts], axis=0)
    return out


def gather_columns(model, local, dist=None):
    """All-gather a column-split spectral field ((nx, nx/P) slabs in the symmetric ownership) into natural order."""
    if model.nranks == 1:
        return local
    if dist is None:
        import torch.distributed as dist
    import torch
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    loc = np.ascontiguousarray(local)
    t = torch.from_numpy(loc.view(np.float64).reshape(-1)).to(dev)
    parts = [torch.empty_like(t) for _ in range(model.nranks)]
    dist.all_gather(parts, t)
    N = model.nx
    out = np.empty((N, N), loc.dtype)
    for r, p in enumerate(parts):
        out[:, nat.slab_kx(N, model.nranks, r)] = p.cpu().numpy().view(loc.dtype).reshape(loc.shape)
    return out
