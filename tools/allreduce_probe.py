"""Latency of the one collective of the view-sharded step (11.5 MB fp32 all-reduce at 131 072 Gaussians) under a few
NCCL settings; launched with torchrun, environment variables select the variant (NCCL reads them at init)."""
import os, sys, time
import torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n = 131072 * 22
x = torch.randn(n, device=dev)
work = torch.empty(64 << 20, dtype=torch.float32, device=dev)      # ~a step's worth of unrelated kernel time before the collective
def timed(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        work.mul_(1.0001)                      # keeps the ranks' streams busy so that launch latency is hidden, like in the bench
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
res = {"all_reduce": timed(lambda: dist.all_reduce(x))}
half = n // 2
res["two_halves"] = timed(lambda: (dist.all_reduce(x[:half]), dist.all_reduce(x[half:])))
out = torch.empty(n // world, device=dev)
res["reduce_scatter+all_gather"] = timed(lambda: (dist.reduce_scatter_tensor(out, x), dist.all_gather_into_tensor(x, out)))
if rank == 0:
    print({k: round(v, 1) for k, v in res.items()}, "us;", {k: os.environ[k] for k in os.environ if k.startswith("NCCL_")}, flush=True)
dist.destroy_process_group()
