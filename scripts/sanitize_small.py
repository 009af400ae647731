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
args = config.default_test_args(sub_pixel=False, num_features=256)
u8 = torch.randint(0, 256, (3, 120, 186, 1), dtype=torch.uint8, device=dev)
for nms in ("windowed", "greedy"):
    xy, sc, _, cnt = demo_match.detect_batch_device(args, u8, det, nms)
    print(nms, cnt.tolist())
xy, sc, lv, cnt = demo_match.detect_multiscale_batch_device(args, u8, det, scale=0.7, levels=2)
print("multiscale", cnt.tolist())
d1 = torch.nn.functional.normalize(torch.randn(300, 128, device=dev), dim=1)
d2 = torch.nn.functional.normalize(d1[:257] + 0.05 * torch.randn(257, 128, device=dev), dim=1)
print("smnn", c.match_smnn(d1, d2, 0.99)[1].shape)
torch.cuda.synchronize()
print("done")
