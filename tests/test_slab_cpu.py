"""Host-side logic of the slab decomposition on CPU: the conjugate-symmetric column map (mirror of struct Grid in
csrc/common.cuh) and the data flow of one distributed 2-D transform (local row pass -> all-to-all in the exchange
layout -> local column pass), replayed with numpy over a world_size-2 gloo process group."""
import os
import socket

import numpy as np
import pytest

from niwqg_b200 import _native as nat


@pytest.mark.parametrize("N,P", [(32, 1), (32, 2), (64, 4), (256, 8), (8192, 8)])
def test_column_map_is_a_symmetric_partition(N, P):
    seen = np.zeros(N, int)
    for r in range(P):
        kx = nat.slab_kx(N, P, r)
        assert len(kx) == N // P
        seen[kx] += 1
        if P > 1:
            h = N // (2 * P)
            for lc, k in enumerate(kx):
                assert nat.slab_owner(N, P, int(k)) == (r, lc)
                # the conjugate partner N-kx is on the same rank, at the mirrored slot
                lcp = lc if (r == 0 and lc in (0, h)) else (lc + h if lc < h else lc - h)
                assert kx[lcp] == (N - k) % N
    assert np.all(seen == 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(5)
        x = rng.randn(N, N) + 1j * rng.randn(N, N)          # same field on every rank
        nyl = ncl = N // world
        lo, hi = rank * nyl, (rank + 1) * nyl
        # forward: row pass on my rows, stored in the exchange layout [owner(kx)][yl][lc(kx)]
        rows = np.fft.fft(x[lo:hi], axis=1)
        send = np.empty((world, nyl, ncl), complex)
        for kx in range(N):
            r, lc = nat.slab_owner(N, world, kx)
            send[r, :, lc] = rows[:, kx]
        recv = np.empty_like(send)
        outs = list(torch.from_numpy(recv.view(np.float64)).unbind(0))
        ins = list(torch.from_numpy(send.view(np.float64)).unbind(0))
        _gloo_all_to_all(dist, outs, ins, rank, world)
        recv = torch.stack(outs).numpy().view(np.complex128).reshape(world, nyl, ncl)
        # chunk r holds rows [r nyl, (r+1) nyl): the receive buffer IS the (N, ncl) column slab
        slab = recv.reshape(N, ncl)
        spec = np.fft.fft(slab, axis=0)
        ref = np.fft.fft2(x)[:, nat.slab_kx(N, world, rank)]
        err_f = np.abs(spec - ref).max() / np.abs(ref).max()
        # inverse: column pass, chunks by destination rows are contiguous, row pass gathers through the exchange layout
        cols = np.fft.ifft(spec, axis=0)
        send = np.ascontiguousarray(cols.reshape(world, nyl, ncl))
        outs = list(torch.from_numpy(np.empty_like(send).view(np.float64)).unbind(0))
        ins = list(torch.from_numpy(send.view(np.float64)).unbind(0))
        _gloo_all_to_all(dist, outs, ins, rank, world)
        recv = torch.stack(outs).numpy().view(np.complex128).reshape(world, nyl, ncl)
        line = np.empty((nyl, N), complex)
        for kx in range(N):
            r, lc = nat.slab_owner(N, world, kx)
            line[:, kx] = recv[r, :, lc]
        back = np.fft.ifft(line, axis=1)
        err_i = np.abs(back - x[lo:hi]).max()
        # budget sums: every rank reduces its slab, all-reduce gives the grid-wide mean
        t = torch.tensor([float((np.abs(x[lo:hi]) ** 2).sum())], dtype=torch.float64)
        dist.all_reduce(t)
        err_s = abs(t.item() - (np.abs(x) ** 2).sum()) / (np.abs(x) ** 2).sum()
        q.put((rank, err_f, err_i, err_s))
    finally:
        dist.destroy_process_group()


def _gloo_all_to_all(dist, outs, ins, rank, world):
    """gloo has no all_to_all: the same exchange as ncclSend/ncclRecv pairs, written with isend/irecv."""
    reqs = []
    for r in range(world):
        if r == rank:
            outs[r].copy_(ins[r])
        else:
            reqs.append(dist.isend(ins[r].contiguous(), r))
            reqs.append(dist.irecv(outs[r], r))
    for rq in reqs:
        rq.wait()


def test_distributed_fft_data_flow_world2_gloo():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 32, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ef, ei, es in res:
        assert ef < 1e-13 and ei < 1e-13 and es < 1e-13, (rank, ef, ei, es)
