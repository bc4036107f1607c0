#!/usr/bin/env python3
"""PCIe ceiling of the box: pinned H2D alone, D2H alone, both at once (torch, two streams).
The e2e leg of bench.py moves 118 MB each way per 8K frame; this is the number it is bound by.

Single process: one GPU.  Under torchrun (`python -m torch.distributed.run --nproc-per-node N
--master-addr 127.0.0.1 tools/pcie_probe.py`) every rank drives its own GPU at the SAME time (a
barrier separates the phases), so the output is the concurrent ceiling of the N links - what the
e2e legs of an N-GPU bench.py run share."""
import json
import os
import time

import torch

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = 1 << 29
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=8):
    fn()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    d_a.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_b, non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


gb = n / 1e9
res = {"h2d_alone": gb / timed(h2d), "d2h_alone": gb / timed(d2h), "both_each_way": gb / timed(both)}
if dist:
    t = torch.tensor([res["h2d_alone"], res["d2h_alone"], res["both_each_way"]], device="cuda",
                     dtype=torch.float64)
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    if rank == 0:
        rows = [[round(float(v), 1) for v in r.tolist()] for r in allr]
        tot = [round(sum(r[k] for r in rows), 1) for k in range(3)]
        print(json.dumps({"gpus_concurrent": world, "unit": "GB/s",
                          "per_rank[h2d_alone,d2h_alone,both_each_way]": rows, "total": tot,
                          "8k_frames_per_s_at_both_total": round(tot[2] / 0.11796, 1)}))
    dist.destroy_process_group()
else:
    print(json.dumps({"gpus_concurrent": 1, "unit": "GB/s", **{k: round(v, 1) for k, v in res.items()},
                      "8k_frames_per_s_at_both": round(res["both_each_way"] / 0.11796, 1)}))
