"""Multi-GPU host logic on CPU: world_size-2 gloo processes exercise pvac_hfhe_cppbyv_b200/shard.py (partition by batch
index, per-item tape streams by GLOBAL index, one key-blob broadcast, result gather). The compute stand-in inside the
ranks is the oracle (this is test code): the point is that a 2-rank run is byte-identical to the 1-rank run."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

from pvac_hfhe_cppbyv_b200 import shard


def test_partition_properties():
    for n in (0, 1, 2, 7, 8, 1000, 2**20, 2**24 + 5):
        for world in (1, 2, 3, 4, 8):
            parts = shard.partition(n, world)
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for r in range(world - 1):
                assert parts[r][0] + parts[r][1] == parts[r + 1][0]
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
            for i in ([0, n // 3, n - 1] if n else []):
                r = shard.owner(i, n, world)
                assert parts[r][0] <= i < parts[r][0] + parts[r][1]
    assert shard.tiles(5, 10, 4) == [(5, 4), (9, 4), (13, 2)]


def test_tape_states_match_engine_definition(port):
    st = shard.item_tape_states(1000, 5, 4)
    assert [int(x) for x in st] == [port.item_stream_state(1000, i) for i in range(5, 9)]
    # independent of how the range is cut
    assert np.array_equal(np.concatenate([shard.item_tape_states(7, 0, 3), shard.item_tape_states(7, 3, 5)]), shard.item_tape_states(7, 0, 8))


class _FakeEngine:
    """stands in for api.Engine's key-blob calls on a CPU tensor"""

    def __init__(self, blob=None):
        self.blob = blob
        self.adopted = None

    def copy_key_blob_to(self, ptr):
        import ctypes
        ctypes.memmove(ptr, self.blob.ctypes.data, self.blob.nbytes)

    def adopt_key_blob_from(self, ptr):
        import ctypes
        out = np.zeros(self._n, np.uint8)
        ctypes.memmove(out.ctypes.data, ptr, self._n)
        self.adopted = out


def _worker(rank, world, port_no, n_items, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle import port
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = shard.Shard.from_env()
    # ---- keys: only rank 0 generates; one broadcast replicates the blob
    nbytes = 4096 + 8 * 674
    blob = torch.zeros(nbytes, dtype=torch.uint8)
    eng = _FakeEngine()
    eng._n = nbytes
    if rank == 0:
        K0 = port.Keys.keygen(1)
        e = K0.export(with_H=False)
        raw = np.zeros(nbytes, np.uint8)
        raw[:8] = np.frombuffer(np.uint64(e["canon_tag"]).tobytes(), np.uint8)
        raw[8:40] = e["H_digest"]
        raw[40:72] = e["prf_k"].view(np.uint8)
        raw[72:584] = e["lpn_s"].view(np.uint8)
        raw[4096:] = np.ascontiguousarray(e["powg"]).view(np.uint8).reshape(-1)
        eng.blob = raw
    sh.replicate_keys(eng, blob)
    got = blob.numpy() if rank == 0 else eng.adopted
    key_digest = hashlib.sha256(got.tobytes()).hexdigest()
    # ---- the job: enc_value of items [first, first+count) with the tape stream of the GLOBAL index, then decrypt
    K = port.Keys.keygen(1)          # H is too big for this toy blob; the oracle regenerates the same keys from tape state 1
    K.set_lpn_t(127)
    first, count = sh.my_range(n_items)
    vals = (np.arange(n_items, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ np.uint64(0xABCDEF)
    states = shard.item_tape_states(4242, first, count)
    digests = np.zeros((count, 4), np.uint64)
    decs = np.zeros((count, 2), np.uint64)
    for j in range(count):
        c = K.enc_value(int(states[j]), int(vals[first + j]))
        d = port.ct_export(c)
        h = hashlib.sha256()
        for k in ("rule", "ztag", "nlo", "nhi", "lid", "idx", "ch", "w", "sigma"):
            h.update(np.ascontiguousarray(d[k]).tobytes())
        digests[j] = np.frombuffer(h.digest(), np.uint64)
        decs[j] = K.dec_value(c)
    all_dig = sh.gather_to_root(digests, n_items)
    all_dec = sh.gather_to_root(decs, n_items)
    keys = [None] * world
    dist.all_gather_object(keys, key_digest)
    if rank == 0:
        np.savez(os.path.join(out_dir, f"w{world}.npz"), dig=all_dig, dec=all_dec, keys=np.array(keys))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(600)
def test_two_ranks_equal_one_rank(tmp_path):
    import torch.multiprocessing as mp
    n_items = 7        # odd on purpose: ragged shards (4 + 3)
    for world in (1, 2):
        mp.spawn(_worker, args=(world, _free_port(), n_items, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "w1.npz"), np.load(tmp_path / "w2.npz")
    assert len(set(b["keys"].tolist())) == 1 and b["keys"][0] == a["keys"][0]      # every rank holds rank 0's key bytes
    assert np.array_equal(a["dig"], b["dig"])                                      # ciphertext bytes identical item by item
    assert np.array_equal(a["dec"], b["dec"])
    vals = (np.arange(n_items, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ np.uint64(0xABCDEF)
    assert np.array_equal(b["dec"][:, 0], vals) and not b["dec"][:, 1].any()       # and they decrypt to the plaintexts
