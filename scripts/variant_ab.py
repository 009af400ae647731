"""Development helper (GPU): A/B of the branch-kernel variants selected by balf_debug_set(4, mask)
(2 bits per stage, see BranchSel in csrc/detector_tc.cuh): score-map difference against variant 0 and per-kernel times.
    python scripts/variant_ab.py [B] [mask ...]"""
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
det.precision = "tf32"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
masks = [int(a, 0) for a in sys.argv[2:]] or [0, 0x01, 0x02, 0x40, 0x41]
g = torch.Generator().manual_seed(1234)
u8 = torch.randint(0, 256, (B, 480, 640, 1), dtype=torch.uint8, generator=g).to(dev)
x, _ = c.preprocess_u8(u8)
ref = None
import copy
d32 = copy.deepcopy(det); d32.precision = "fp32"
with torch.inference_mode():
    p32 = d32(x[:2])["prob"].double()
for m in masks:
    c.debug_set(4, m)
    with torch.inference_mode():
        for _ in range(2):
            p = det(x)["prob"]
        torch.cuda.synchronize()
        c.profile_enable(True)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5):
            p = det(x)["prob"]
        e1.record()
        torch.cuda.synchronize()
        rep = c.profile_report(True)
        c.profile_enable(False)
    ms = e0.elapsed_time(e1) / 5
    if ref is None:
        ref = p.clone()
    rel = ((p - ref).abs() / ref).max().item()
    bad = int((~torch.isfinite(p)).sum().item())
    r32 = ((p[:2].double() - p32).abs() / p32)
    print("          vs fp32 path (2 images): max rel %.3e  mean %.3e" % (r32.max().item(), r32.mean().item()))
    print("mask 0x%02x: %.3f ms / %d images = %.1f img/s   max rel vs mask 0: %.3e  nonfinite %d" % (m, ms, B, B / ms * 1e3, rel, bad))
    rows = sorted(rep.items(), key=lambda kv: -kv[1][1]) if isinstance(rep, dict) else []
    for name, (n, tot) in rows[:16]:
        print("    %-26s %3d launches  %8.3f ms each" % (name, n, tot / max(n, 1)))
c.debug_set(4, 0)
