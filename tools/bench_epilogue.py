"""Time the fused render_img epilogue against the torch ops it replaces (fwd + bwd, one 512x512 view)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lara_b200.epilogue import render_img_epilogue
from oracle.torch_restatements import render_img_epilogue_torch
from test_epilogue import _inputs

dev = torch.device("cuda:0")
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 512
color, allmap, rays, vm = _inputs(H, W, 1, dev)
allmap[1].clamp_(min=0.05); 
w = {k: torch.randn(s, device=dev) for k, s in (("image", (H, W, 3)), ("depth", (H, W, 1)), ("acc_map", (H, W)),
                                                ("rend_normal", (H, W, 3)), ("depth_normal", (H, W, 3)), ("rend_dist", (H, W)))}
for name, fn in (("torch ops (reference)", render_img_epilogue_torch), ("fused kernel", render_img_epilogue)):
    def step():
        c = color.clone().requires_grad_(True); a = allmap.clone().requires_grad_(True)
        out = fn(c, a, rays, vm, 0.0)
        torch.autograd.backward([out[k] for k in w], [w[k] for k in w])
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    N = 100
    for _ in range(N):
        step()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:24s} {H}x{W}: {e0.elapsed_time(e1) / N * 1e3:8.1f} us/view GPU, {(time.perf_counter() - t0) / N * 1e6:8.1f} us/view wall")
