import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lara_b200 import scene as S, rasterizer as R
from lara_b200.debug import unpack_state
from oracle import ref as REF
dev = torch.device("cuda:0")
ref = REF.load()
P,H,W=32768,512,512
sc = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in S.scene(P, 0).items()}
cam = S.cameras(3, H, W, 0)[0]
bg = torch.ones(3)
ms = S.settings_for(cam, bg, 1, dev, R.GaussianRasterizationSettings)
rs = S.settings_for(cam, bg, 1, dev, ref.GaussianRasterizationSettings)
color, allmap, radii, st = R.forward_raw(sc["means3D"], sc["shs"], None, sc["opacities"], sc["scales"], sc["rotations"], None, ms)
mine = unpack_state(st, P, H, W)
r = REF.forward_raw(ref, sc, rs)
torch.cuda.synchronize()
bad = (mine["n_contrib"][1] != r["n_contrib"][1]).nonzero()
print(bad.shape)
for y,x in bad[:12].tolist():
    print(y,x,"mine",mine["n_contrib"][1,y,x].item(),"ref",r["n_contrib"][1,y,x].item(),"last",mine["n_contrib"][0,y,x].item(), "T", mine["accum"][0,y,x].item(), "meddepth", allmap[5,y,x].item(), r["allmap"][5,y,x].item(), "range", mine["ranges"][(y//16)*32+x//16].tolist())
