"""Host-side overhead of one fwd+bwd view through the drop-in autograd API (tiny scene, GPU work negligible)."""
import cProfile, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lara_b200 import scene as S
import diff_surfel_rasterization as DSR
from oracle import ref as REF

dev = torch.device("cuda:0")
sc = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in S.scene(2000, 0).items()}
cam = S.cameras(1, 64, 64, 0)[0]
gc, ga = [t.to(dev) for t in S.upstream_grads(64, 64, 0)]


def make(mod):
    st = S.settings_for(cam, torch.ones(3), 1, dev, mod.GaussianRasterizationSettings)
    leaves = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)

    def step():
        rast = mod.GaussianRasterizer(raster_settings=st)
        c, rd, am = rast(means3D=leaves["means3D"], means2D=m2d, shs=leaves["shs"], opacities=leaves["opacities"],
                         scales=leaves["scales"], rotations=leaves["rotations"])
        torch.autograd.backward((c, am), (gc, ga))
    return step


for name, mod in (("mine", DSR), ("ref", REF.load() if REF.available() else None)):
    if mod is None:
        continue
    step = make(mod)
    for _ in range(50):
        step()
    torch.cuda.synchronize()
    t = time.perf_counter()
    N = 500
    for _ in range(N):
        step()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t) / N * 1e6:.1f} us per fwd+bwd view (host-bound)")
    if name == "mine" and "--profile" in sys.argv:
        pr = cProfile.Profile(); pr.enable()
        for _ in range(300):
            step()
        torch.cuda.synchronize(); pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
