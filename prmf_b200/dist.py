"""Sample-sharded data parallelism: plumbing only (torch.distributed rendezvous, id exchange, gathers).

Rows of X and U are split into contiguous blocks over ranks; V, the pathway tables,
gamma/delta, the candidate bookkeeping and the host RNG are replicated -- every rank runs the same seeded
host loop and sees bitwise identical all-reduced buffers, so no broadcast of decisions is needed.
The only data-path collective is the per-step all-reduce inside `prmf_step` (NCCL, on the device).
"""
import os

import numpy as np


def row_block(m, world, rank):
    """Contiguous row block [lo, hi) of rank `rank`: ceil(m/world) rows each, the last ones may be short
    or empty."""
    per = -(-m // world)
    lo = min(m, rank * per)
    return lo, min(m, lo + per)


class DistContext:
    """Thin view of a torch.distributed process group (or of a single process when world == 1)."""

    def __init__(self, rank=0, world=1, local_rank=0, backend=None):
        self.rank, self.world, self.local_rank, self.backend = rank, world, local_rank, backend

    @classmethod
    def from_env(cls, backend=None, init=True):
        """Read RANK / WORLD_SIZE / LOCAL_RANK (torchrun) and join the default process group."""
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if world == 1:
            return cls(0, 1, local_rank, None)
        import torch
        import torch.distributed as dist
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl" and torch.cuda.device_count() < world:
            backend = "gloo"             # NCCL cannot put two ranks on one device; the process group is plumbing only
        if init and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(local_rank)
                kw["device_id"] = torch.device("cuda", local_rank)
            elif torch.cuda.is_available():
                torch.cuda.set_device(local_rank % torch.cuda.device_count())
            dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
        return cls(rank, world, local_rank, backend)

    @classmethod
    def current(cls):
        """The already-initialised default group, or a single-process context."""
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                return cls(dist.get_rank(), dist.get_world_size(),
                           int(os.environ.get("LOCAL_RANK", "0")), dist.get_backend())
        except ImportError:
            pass
        return cls()

    # -- small host-side collectives (setup / teardown only) -----------------------------------------
    def broadcast_bytes(self, payload, src=0):
        if self.world == 1:
            return payload
        import torch.distributed as dist
        box = [payload if self.rank == src else None]
        dist.broadcast_object_list(box, src=src)
        return box[0]

    def all_gather_bytes(self, payload):
        """Every rank's `payload` (bytes), in rank order."""
        if self.world == 1:
            return [payload]
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, payload)
        return parts

    def all_gather_rows(self, local, m_global):
        """Concatenate the ranks' row blocks (U at return, :778-792) in rank order."""
        if self.world == 1:
            return local
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, np.ascontiguousarray(local))
        out = np.concatenate([p for p in parts if p.shape[0] > 0], axis=0)
        assert out.shape[0] == m_global
        return out

    def all_reduce_sum(self, arr):
        """Host-array all-reduce (used by CPU stand-in engines in the gloo tests)."""
        if self.world == 1:
            return arr
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if self.backend == "nccl":
            t = t.cuda()
        dist.all_reduce(t)
        return t.cpu().numpy()

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
