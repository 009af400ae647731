"""Development helper (GPU): per-kernel times of the fp32 (FFMA) detector path and of the 'hybrid' precision, batch B."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import balf_b200._capi as c
from balf_b200.model import get_model
from balf_b200.utils import test_utils
from balf_b200.configs import config

dev = torch.device("cuda:0")
cfg = test_utils.get_cfg_from_yaml_file(config.DEFAULT_CFG)
torch.manual_seed(0)
det = get_model.load_model(cfg["model"]).eval().to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
precs = sys.argv[2:] or ["fp32"]
g = torch.Generator().manual_seed(1234)
u8 = torch.randint(0, 256, (B, 480, 640, 1), dtype=torch.uint8, generator=g).to(dev)
x, _ = c.preprocess_u8(u8)
ref = None
for prec in precs:
    det.precision = prec
    with torch.inference_mode():
        p = det(x)["prob"]
        torch.cuda.synchronize()
        c.profile_enable(True)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            p = det(x)["prob"]
        e1.record()
        torch.cuda.synchronize()
        rep = c.profile_report(True)
        c.profile_enable(False)
    ms = e0.elapsed_time(e1) / 3
    if ref is None:
        ref = p.double()
    rel = (p.double() - ref).abs() / ref
    print("%s: %.3f ms / %d images = %.1f img/s; vs %s: max rel %.3e mean %.3e" % (prec, ms, B, B / ms * 1e3, precs[0], rel.max().item(), rel.mean().item()))
    for name, (n, tot) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:24]:
        print("    %-26s %3d launches  %8.3f ms each  %8.3f ms per pass" % (name, n, tot / max(n, 1), tot / 3))
