"""BASELINE.json config 5 on N GPUs: 2^24 mixed enc / ct_mul / dec items sharded by batch index, keys replicated by one NCCL
broadcast, no steady-state collective. Every product is verified against a*b mod p on the rank that made it.

  python profiles/mixed_pipeline.py --items 16777216                                   (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/mixed_pipeline.py --items 16777216
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from pvac_hfhe_cppbyv_b200 import api, shard
from pvac_hfhe_cppbyv_b200.pipeline import run_mixed_pipeline

ap = argparse.ArgumentParser()
ap.add_argument("--items", type=int, default=1 << 20)
ap.add_argument("--tile", type=int, default=1 << 12, help="pairs per tile")
args = ap.parse_args()
sh = shard.Shard.from_env()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if sh.world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = api.Engine(device=local, prf_mode=api.PRF_LIVE, tape=api.TAPE_SPLITMIX)
blob = torch.empty(api.KEY_BLOB_BYTES, dtype=torch.uint8, device=f"cuda:{local}")
if sh.rank == 0:
    eng.keygen(1)
sh.replicate_keys(eng, blob)
del blob
pairs = args.items // 2
first, count = shard.partition(pairs, sh.world)[sh.rank]
torch.cuda.synchronize()
if sh.world > 1:
    dist.barrier()
t0 = time.perf_counter()
checked, bad = run_mixed_pipeline(eng, 2 * first, 2 * count, args.tile)
torch.cuda.synchronize()
secs = time.perf_counter() - t0
res = torch.tensor([checked, bad, secs], dtype=torch.float64, device=f"cuda:{local}")
if sh.world > 1:
    allr = [torch.zeros_like(res) for _ in range(sh.world)]
    dist.all_gather(allr, res)
else:
    allr = [res]
if sh.rank == 0:
    tot_checked = int(sum(r[0].item() for r in allr)); tot_bad = int(sum(r[1].item() for r in allr)); tmax = max(r[2].item() for r in allr)
    print(json.dumps({"workload": "mixed enc/ct_mul/dec (config 5)", "items": args.items, "pairs_checked": tot_checked, "mismatches": tot_bad, "n_gpus": sh.world,
                      "seconds_max_over_ranks": tmax, "items_per_s": args.items / tmax, "per_rank_seconds": [r[2].item() for r in allr], "prf_mode": "live"}))
    assert tot_checked == pairs and tot_bad == 0
if sh.world > 1:
    dist.destroy_process_group()
eng.close()
