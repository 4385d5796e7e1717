import ctypes, os, torch
here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, "act_probe.so"))
n = 1 << 20
g = torch.Generator().manual_seed(0)
q = (torch.randn(n, 4, generator=g) * 1.3).cuda()
tn = torch.linalg.vector_norm(q, dim=1)
out = torch.empty(n, 12, device="cuda"); qd = torch.empty(n, 4, device="cuda")
P = ctypes.c_void_p
lib.probe.argtypes = [ctypes.c_int] + [P] * 4
torch.cuda.synchronize()
assert lib.probe(n, q.data_ptr(), tn.data_ptr(), out.data_ptr(), qd.data_ptr()) == 0
def same(a, b): return int((a.contiguous().view(torch.int32) != b.contiguous().view(torch.int32)).sum())
names = ["rn seq", "rn pw01_23", "rn pw02_13", "rn fma", "approx seq", "approx pw01_23", "approx pw02_13", "approx fma",
         "double sqrt", "rn(float(double sum))", "approx.ftz pw", "approx(float(double sum))"]
for j, nm in enumerate(names):
    print(f"{nm:28s} mismatches {same(out[:, j], tn)}")
print("division given torch's norm:", same(qd, torch.nn.functional.normalize(q)))
print("torch sqrt((q*q).sum(1)) vs vector_norm:", same(torch.sqrt((q * q).sum(1)), tn))
print("normalize == q / tn.clamp_min:", same(q / tn.clamp_min(1e-12)[:, None], torch.nn.functional.normalize(q)))
