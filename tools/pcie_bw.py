"""Per-rank pinned H2D / D2H bandwidth with all ranks copying at once (torchrun) -- what the host gives
each GPU when several share it."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True)),
                 ("both", None)):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    if fn is None:
        s2 = torch.cuda.Stream()
        h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(5):
            d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    else:
        for _ in range(5):
            fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("rank %d %s %.1f GB/s%s" % (rank, name, 5 * n / dt / 1e9, " each way" if fn is None else ""), flush=True)
