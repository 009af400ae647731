import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
import __graft_entry__ as ge
ge.build()
import balf_b200._capi as c
dev = torch.device("cuda:0")
torch.manual_seed(0)
d1 = torch.nn.functional.normalize(torch.randn(2048, 128, device=dev), dim=1)
d2 = torch.nn.functional.normalize(d1 + 0.05 * torch.randn(2048, 128, device=dev), dim=1)
for impl in (1, 0):
    c.debug_set(3, impl)
    dist, ids, dm = c.match_smnn(d1, d2, 0.99, want_dm=True)
    ref = torch.cdist(d1.double(), d2.double())
    print("impl", impl, "matches", ids.shape[0], "max |dm - cdist64|", float((dm.double() - ref).abs().max()))
    c.profile_enable(True); c.profile_report(reset=True)
    for _ in range(5): c.match_smnn(d1, d2, 0.99)
    torch.cuda.synchronize()
    print({k: (v[0], round(v[1] / v[0] * 1e3, 1)) for k, v in c.profile_report().items() if k.startswith("match")})
    c.profile_enable(False)
    if impl == 1: ids_ffma = ids.clone()
print("same match set:", ids.shape == ids_ffma.shape and bool((ids == ids_ffma).all()))
