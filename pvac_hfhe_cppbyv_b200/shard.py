"""Host-side sharding of a batch over the GPUs of one box.

The reference has no distribution at all (single-threaded header-only library); what shards is the batch itself:
ciphertexts are independent, so global item i goes to rank floor(i * world / n) as contiguous ranges, every item keeps the
RNG tape stream of its GLOBAL index (so an N-rank run is byte-identical to the 1-rank run), the keys (pk.H 16 MiB + 6 KB
of tables, sk 544 B: one flat device blob) are replicated ONCE with a broadcast over NVLink, and there is no collective
in the steady state. Small per-item results (16-byte decrypts, digests) can be gathered to rank 0 at the end.

One process per GPU, torch.distributed for the plumbing only (NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np

_M1, _M2, _GOLD, _ITEM = 0xBF58476D1CE4E5B9, 0x94D049BB133111EB, 0x9E3779B97F4A7C15, 0xD1342543DE82EF95


def mix64(z):
    z = np.asarray(z, np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        return z ^ (z >> np.uint64(31))


def item_tape_states(batch_seed, first, count):
    """initial tape state of global items [first, first+count) -- the derivation include/pvacb.h documents"""
    with np.errstate(over="ignore"):
        i = np.arange(first, first + count, dtype=np.uint64)
        return mix64(np.uint64(batch_seed & (2**64 - 1)) + np.uint64(_ITEM) * (i + np.uint64(1)))


def partition(n_items, world):
    """contiguous ranges: rank r owns [bounds[r], bounds[r+1]); item i lives on rank floor(i*world/n)"""
    bounds = [(r * n_items + world - 1) // world for r in range(world + 1)]
    return [(bounds[r], bounds[r + 1] - bounds[r]) for r in range(world)]


def owner(i, n_items, world):
    return (i * world) // n_items


def tiles(first, count, tile):
    """split a rank's range into HBM-sized tiles (2^18..2^24-item configs must be streamed)"""
    out = []
    while count > 0:
        c = min(tile, count)
        out.append((first, c))
        first += c
        count -= c
    return out


class Shard:
    """This rank's view of a sharded job."""

    def __init__(self, rank=0, world=1, group=None):
        self.rank, self.world, self.group = rank, world, group

    @classmethod
    def from_env(cls):
        import os
        return cls(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))

    def my_range(self, n_items):
        return partition(n_items, self.world)[self.rank]

    def replicate_keys(self, engine, blob_tensor):
        """rank 0 holds the keys; everybody ends up with them. blob_tensor: uint8[KEY_BLOB_BYTES] on this rank's device.
        One broadcast (NCCL over NVLink / NVSwitch on the GPU box); never called again in the steady state."""
        import torch
        import torch.distributed as dist
        if self.rank == 0:
            engine.copy_key_blob_to(blob_tensor.data_ptr())
        if self.world > 1:
            if blob_tensor.is_cuda:
                torch.cuda.synchronize()
            dist.broadcast(blob_tensor, 0, group=self.group)
            if blob_tensor.is_cuda:
                torch.cuda.synchronize()
            if self.rank != 0:
                engine.adopt_key_blob_from(blob_tensor.data_ptr())

    def gather_to_root(self, local, n_items):
        """concatenate per-item results (numpy array, first axis = this rank's items) on rank 0 in global order"""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return local
        parts = partition(n_items, self.world)
        width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
        cap = max(c for _, c in parts)
        buf = torch.zeros(cap * width, dtype=torch.int64)
        flat = np.ascontiguousarray(local).view(np.int64).reshape(-1) if local.dtype.itemsize == 8 else np.ascontiguousarray(local, np.int64).reshape(-1)
        buf[: flat.size] = torch.from_numpy(flat.copy())
        backend = dist.get_backend(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        buf = buf.to(dev)
        outs = [torch.zeros_like(buf) for _ in range(self.world)] if self.rank == 0 else None
        if backend == "nccl":                       # NCCL has no gather-to-one primitive in every torch build: all_gather
            outs = [torch.zeros_like(buf) for _ in range(self.world)]
            dist.all_gather(outs, buf, group=self.group)
        else:
            dist.gather(buf, outs, dst=0, group=self.group)
        if self.rank != 0:
            return None
        cols = [o.cpu().numpy()[: c * width] for o, (_, c) in zip(outs, parts)]
        res = np.concatenate(cols)
        if local.dtype.itemsize == 8:
            res = res.view(local.dtype)
        else:
            res = res.astype(local.dtype)
        return res.reshape((n_items,) + tuple(local.shape[1:]))
