"""Development helper (GPU): one small tf32 forward + windowed / greedy extraction + SMNN, meant to run under
compute-sanitizer:   compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import balf_b200._capi as c
from balf_b200.model import get_model
from balf_b200.utils import test_utils
from balf_b200.configs import config
from balf_b200.demo import demo_match
dev = torch.device("cuda:0")
cfg = test_utils.get_cfg_from_yaml_file(config.DEFAULT_CFG)
torch.manual_seed(0)
det = get_model.load_model(cfg["model"]).eval().to(dev)
det.precision = "tf32"            # the tensor-core kernels (fp16 / tf32 operands); the default 'auto' would pick fp32 for greedy
args = config.default_test_args(sub_pixel=False, num_features=256)
u8 = torch.randint(0, 256, (3, 120, 186, 1), dtype=torch.uint8, device=dev)
for nms in ("windowed", "greedy"):
    xy, sc, _, cnt = demo_match.detect_batch_device(args, u8, det, nms)
    print(nms, cnt.tolist())
det.precision = "f16x3"           # the same kernels in split precision (fp16 hi + lo operand pairs)
xy, sc, _, cnt = demo_match.detect_batch_device(args, u8, det, "greedy")
print("greedy f16x3", cnt.tolist())
det.precision = "tf32"
xy, sc, lv, cnt = demo_match.detect_multiscale_batch_device(args, u8, det, scale=0.7, levels=2)
print("multiscale", cnt.tolist())
d1 = torch.nn.functional.normalize(torch.randn(300, 128, device=dev), dim=1)
d2 = torch.nn.functional.normalize(d1[:257] + 0.05 * torch.randn(257, 128, device=dev), dim=1)
print("smnn", c.match_smnn(d1, d2, 0.99)[1].shape)
# round 2: HardNet fp16 path, box_nms, repeatability metrics, RGB -> L
from balf_b200.third_party.hardnet.hardnet_pytorch import HardNet
from balf_b200.benchmark_test import repeatability_tools as rt, geometry_tools as gt
import numpy as np
torch.manual_seed(0)
hn = HardNet().eval().to(dev)
with torch.inference_mode():
    print("hardnet", hn(torch.rand(70, 1, 32, 32, device=dev)).shape)
print("box_nms", int((rt.box_nms(torch.rand(1, 60, 80, device=dev) * 0.05) != 0).sum()))
rng = np.random.default_rng(0)
a = np.stack([rng.uniform(0, 300, 200), rng.uniform(0, 200, 200), rng.uniform(0.5, 3, 200), rng.random(200)], 1)
b = a[:150] + np.concatenate([rng.normal(0, 1.5, (150, 2)), np.zeros((150, 2))], 1)
print("rep", rt.compute_repeatability(a, b)["num_points_single_scale"])
Hm = np.array([[1.02, 0.03, 5.0], [-0.02, 0.98, -3.0], [1e-5, -2e-5, 1.0]])
print("masks", [int(m.sum()) for m in gt.create_common_region_masks(Hm, (200, 300, 3), (210, 290, 3))])
print("gray", c.rgb_to_gray(torch.randint(0, 256, (40, 50, 3), dtype=torch.uint8, device=dev)).shape)
torch.cuda.synchronize()
print("done")
