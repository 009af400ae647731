"""Development helper (GPU): which Down stages' operand rounding costs greedy keypoint agreement?  Runs the single-rounded
tensor path with some stages sent back to the fp32 kernels (balf_debug_set(0, mask): bit l = stage l+1 on tensor cores,
bit 4 = head) and reports score-map error and greedy agreement with the all-fp32 path over the six parity seeds.
    python scripts/mixed_precision_probe.py [mask ...]"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
import balf_b200._capi as capi
from balf_b200.model import get_model
from balf_b200.utils import test_utils
from balf_b200.configs import config
from balf_b200.demo import demo_match

dev = torch.device("cuda:0")
cfg = test_utils.get_cfg_from_yaml_file(config.DEFAULT_CFG)
torch.manual_seed(0)
det = get_model.load_model(cfg["model"]).eval().to(dev)
d32 = copy.deepcopy(det); d32.precision = "fp32"
dtc = copy.deepcopy(det); dtc.precision = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].isdigit() else "tf32"
masks = [int(a) for a in sys.argv[1:] if a.isdigit()] or [31, 3, 28, 1, 2, 4, 8, 16, 30, 29, 27, 23, 15]
args = config.default_test_args(sub_pixel=False)
ims, refs, maps = [], [], []
for seed in (1234, 1, 2, 3, 4, 5):
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (1, 480, 640), generator=g, dtype=torch.uint8)
    im = u8.permute(1, 2, 0).expand(480, 640, 3).contiguous().numpy()
    ims.append(im)
    refs.append(set(map(tuple, demo_match.detect(args, im, d32, dev)[:, :2])))
    x, _ = capi.preprocess_u8(torch.from_numpy(im[None]).to(dev))
    with torch.inference_mode():
        maps.append((x, d32(x)["prob"].double()))
for m in masks:
    capi.debug_set(0, m)
    agree, worst, tot_i, tot_n, mx, mean = [], 1.0, 0, 0, 0.0, 0.0
    for im, ref, (x, p32) in zip(ims, refs, maps):
        got = set(map(tuple, demo_match.detect(args, im, dtc, dev)[:, :2]))
        tot_i += len(got & ref); tot_n += len(ref)
        worst = min(worst, len(got & ref) / len(ref))
        with torch.inference_mode():
            rel = ((dtc(x)["prob"].double() - p32) / p32).abs()
        mx = max(mx, rel.max().item()); mean += rel.mean().item() / len(ims)
    print("mask %2d (tensor stages %s)  max rel %.2e  mean %.2e  greedy agreement %.4f  worst image %.4f" %
          (m, "".join(str(i + 1) if m >> i & 1 else "-" for i in range(5)), mx, mean, tot_i / tot_n, worst), flush=True)
capi.debug_set(0, 31)
