"""world_size-2 gloo test of the one-process-per-GPU plumbing (dist.py): handle broadcast, barriers, max-over-ranks
timing, and the create/open/reset/render/read call order of SharedRender — against a recording fake scene, since
the CUDA IPC calls themselves need GPUs."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeScene:
    def __init__(self, rank):
        self.rank, self.calls, self.handle = rank, [], None

    def _c(self, name, *a):
        self.calls.append(name)
        if name == "shared_create":
            buf = a[2]
            for i in range(len(buf)):
                buf[i] = (i * 7 + 3) & 255
        if name == "shared_open":
            self.handle = bytes(a[2])
        if name == "render_shared_epoch":
            self.epochs = getattr(self, "epochs", []) + [a[2]]
            st = a[3]._obj
            st.ms_render = 10.0 + self.rank
            st.rays = 100 * (self.rank + 1)
        return 0


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import rtw_pkg
    m = rtw_pkg.load()
    from rtw_b200 import dist
    comm = dist.Comm(backend="gloo", device="cpu")
    assert (comm.rank, comm.world) == (rank, world)
    got = comm.broadcast_bytes(b"\x01\x02\x03\x04" if rank == 0 else None, 4)
    assert got == b"\x01\x02\x03\x04"
    assert comm.reduce_max(5.0 + rank) == 5.0 + world - 1
    assert comm.reduce_sum(1.0 + rank) == sum(1.0 + r for r in range(world))
    sc = FakeScene(rank)
    sr = dist.SharedRender(comm, sc, 64, 32)
    want = bytes((i * 7 + 3) & 255 for i in range(160))
    assert sr.handle == want
    if rank != 0:
        assert sc.handle == want
    cam, prm = m.api.Camera(), m.make_params(64, 32, 4)
    st = sr.step(cam, prm)
    sr.step(cam, prm)                            # a second epoch: the halves alternate, no reset call in between
    if rank == 0:
        sr.read()
    ms = comm.reduce_max(st["ms_render"])
    rays = comm.reduce_sum(st["rays"])
    sr.close()
    comm.close()
    q.put((rank, sc.calls, ms, rays, sc.epochs))


def test_two_rank_gloo_plumbing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == ["shared_create", "render_shared_epoch", "render_shared_epoch", "shared_read_epoch", "shared_close"]
    assert res[1][1] == ["shared_open", "render_shared_epoch", "render_shared_epoch", "shared_close"]
    assert res[0][4] == res[1][4] == [0, 1]       # every rank renders the same epoch (= the same half of the shared buffers)
    assert res[0][2] == res[1][2] == 11.0          # max over ranks of the device time
    assert res[0][3] == res[1][3] == 300.0


def test_single_rank_needs_no_process_group():
    sys.path.insert(0, ROOT)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        os.environ.pop(k, None)
    import rtw_pkg
    rtw_pkg.load()
    from rtw_b200 import dist
    comm = dist.Comm(backend="gloo", device="cpu")
    assert comm.world == 1 and comm.reduce_max(3.0) == 3.0 and comm.broadcast_bytes(b"ab", 2) == b"ab"
    comm.barrier()
