"""One-process-per-GPU plumbing (torch.distributed is used for rendezvous/barriers only — never for pixels).

Pixels shard naturally, so there is no data-path collective: rank 0 owns the framebuffer and the atomic
work-unit counter in its HBM, exports them with CUDA IPC, every other rank maps them, and all ranks'
megakernels pull units from the same counter and add finished tiles straight into rank 0's framebuffer over
NVLink (peer atomics).  The only messages on the torch.distributed side are the 160-byte IPC handle, barriers
and the max-over-ranks of the device time.
"""
import os

import numpy as np


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class Comm:
    """Thin wrapper so the same code runs over NCCL (GPU box) and gloo (CPU tests)."""

    def __init__(self, backend=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.local_rank, self.world = env_rank()
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        self.backend = backend
        if device is None:
            device = f"cuda:{self.local_rank}" if backend == "nccl" else "cpu"
        self.device = torch.device(device)
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        self.owns = False
        if self.world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            kw = {}
            if backend == "nccl":
                kw["device_id"] = self.device
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world, **kw)
            self.owns = True

    def barrier(self):
        if self.world > 1:
            if self.backend == "nccl":
                self.dist.barrier(device_ids=[self.device.index])
            else:
                self.dist.barrier()
        if self.device.type == "cuda":
            self.torch.cuda.synchronize(self.device)

    def broadcast_bytes(self, payload, nbytes, src=0):
        """rank `src` passes `payload` (bytes-like of length nbytes); everyone gets the bytes back."""
        t = self.torch.zeros(nbytes, dtype=self.torch.uint8)
        if self.rank == src:
            t.copy_(self.torch.frombuffer(bytearray(payload), dtype=self.torch.uint8))
        if self.world > 1:
            t = t.to(self.device)
            self.dist.broadcast(t, src=src)
            t = t.cpu()
        return bytes(t.numpy().tobytes())

    def reduce_max(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.owns and self.dist.is_initialized():
            self.dist.destroy_process_group()


class SharedRender:
    """Drives rtw_shared_* across ranks: create/open once, then step() = render epoch e -> barrier.  Epochs alternate between
    the two halves of the shared allocation (rank 0 zeroes the idle half during the render), so a step needs no reset call
    and no barrier before the launch: ONE barrier per step (it also is the contract's end-of-step synchronisation)."""

    def __init__(self, comm, scene, width, height, handle_bytes=160):
        import ctypes as C
        self.comm, self.scene, self.w, self.h = comm, scene, width, height
        self.C = C
        buf = (C.c_uint8 * handle_bytes)()
        if comm.rank == 0:
            scene._c("shared_create", width, height, buf)
        raw = comm.broadcast_bytes(bytes(buf), handle_bytes, src=0)
        if comm.rank != 0:
            buf2 = (C.c_uint8 * handle_bytes).from_buffer_copy(raw)
            scene._c("shared_open", width, height, buf2)
        self.handle = raw
        self.epoch = 0
        comm.barrier()

    def step(self, cam, params):
        """One whole-image render over all ranks.  Returns this rank's stats dict (ms_render = its device time)."""
        from .api import Stats
        C = self.C
        st = Stats()
        self.scene._c("render_shared_epoch", C.byref(cam), C.byref(params), self.epoch, C.byref(st))
        self.comm.barrier()                 # every rank is done with this epoch (and rank 0's side stream has zeroed the other half)
        self.epoch += 1
        return st.as_dict()

    def read(self, out=None):
        assert self.comm.rank == 0
        if out is None:
            out = np.zeros((self.h, self.w, 3), np.float32)
        self.scene._c("shared_read_epoch", max(self.epoch - 1, 0), out.ctypes.data_as(self.C.POINTER(self.C.c_float)))
        return out

    def close(self):
        self.comm.barrier()
        if self.comm.rank != 0:
            self.scene._c("shared_close")
        self.comm.barrier()
        if self.comm.rank == 0:
            self.scene._c("shared_close")
