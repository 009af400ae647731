"""Development helper: per-stage error of the tensor-core detector path against the fp32 path."""
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
shapes = [(1, 64, 64), (1, 128, 192), (3, 64, 128), (9, 128, 128)] if len(sys.argv) < 2 else [tuple(map(int, sys.argv[1].split("x")))]
for (B, H, W) in shapes:
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(5)).to(dev)
    det.precision = "fp32"
    with torch.inference_mode():
        ref = det(x)
    det.precision = "tf32"
    for mask in (1, 2, 4, 8, 16, 31):
        c.debug_set(0, mask)
        try:
            with torch.inference_mode():
                out = det(x)
            torch.cuda.synchronize()
        except Exception as e:
            print("B%d %dx%d mask %2d: ERROR %s" % (B, H, W, mask, str(e)[:200]))
            sys.exit(1)
        rel = ((out["prob"] - ref["prob"]).abs() / ref["prob"]).max().item()
        la = (out["logits"] - ref["logits"]).abs().max().item()
        print("B%d %dx%d mask %2d: prob max rel %.3e  logits max abs %.3e  nan %d" % (B, H, W, mask, rel, la, int(torch.isnan(out["prob"]).sum())), flush=True)
c.debug_set(0, 31)
