"""Development helper (GPU): the split-precision tensor path ('f16x3') against the fp32 FFMA path and the tf32-class path:
score-map / logit error on several shapes, per-kernel times at batch 64, greedy keypoint agreement on 480x640 images."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
import balf_b200._capi as c
from balf_b200.model import get_model
from balf_b200.utils import test_utils
from balf_b200.configs import config
from balf_b200.demo import demo_match
from conftest import synth_u8

dev = torch.device("cuda:0")
cfg = test_utils.get_cfg_from_yaml_file(config.DEFAULT_CFG)
torch.manual_seed(0)
det = get_model.load_model(cfg["model"]).eval().to(dev)
for (B, H, W) in ((2, 128, 192), (1, 64, 64), (3, 512, 640), (2, 960, 1216)):
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(7)).to(dev)
    with torch.inference_mode():
        o32 = det(x, precision="fp32")
        for prec in ("tf32", "f16x3"):
            o = det(x, precision=prec)
            rel = ((o["prob"].double() - o32["prob"].double()).abs() / o32["prob"].double())
            dl = (o["logits"].double() - o32["logits"].double()).abs().max().item()
            print("%dx%dx%d %-6s: prob max rel %.3e mean %.3e  logits max abs %.3e  finite %s" % (
                B, H, W, prec, rel.max().item(), rel.mean().item(), dl, bool(torch.isfinite(o["prob"]).all())))
if "--agree" in sys.argv:
    args = config.default_test_args(sub_pixel=False)
    import copy
    for seed in (1234, 1, 2, 3, 4, 5):
        im = synth_u8(480, 640, seed)
        res = {}
        for prec in ("fp32", "tf32", "f16x3"):
            d = copy.deepcopy(det); d.precision = prec
            res[prec] = set(map(tuple, demo_match.detect(args, im, d, "cuda:0")[:, :2]))
        print("seed %d: %d kps; agreement with the fp32 path: tf32 %.4f  f16x3 %.4f" % (
            seed, len(res["fp32"]), len(res["tf32"] & res["fp32"]) / len(res["fp32"]), len(res["f16x3"] & res["fp32"]) / len(res["fp32"])))
if "--time" in sys.argv:
    g = torch.Generator().manual_seed(1234)
    u8 = torch.randint(0, 256, (64, 480, 640, 1), dtype=torch.uint8, generator=g).to(dev)
    x, _ = c.preprocess_u8(u8)
    for prec in ("tf32", "f16x3"):
        with torch.inference_mode():
            for _ in range(2):
                det(x, precision=prec)
            torch.cuda.synchronize()
            c.profile_enable(True)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(5):
                det(x, precision=prec)
            e1.record()
            torch.cuda.synchronize()
            rep = c.profile_report(True)
            c.profile_enable(False)
        ms = e0.elapsed_time(e1) / 5
        print("%s: %.3f ms / 64 images = %.1f img/s" % (prec, ms, 64 / ms * 1e3))
        for name, (n, tot) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:16]:
            print("    %-26s %8.3f ms per pass" % (name, tot / 5))
