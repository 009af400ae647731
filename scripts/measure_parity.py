"""Development helper (GPU): measured errors behind the tolerances of tests/test_gpu_describe_match.py and
tests/test_gpu_full_size_properties.py -- HardNet descriptor error against the reference's 2048-patch golden, match
agreement of the full pair pipeline against the CPU oracle at 200x264 and 900x1200.    python scripts/measure_parity.py"""
import os, sys, time, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from conftest import synth_u8, load_golden
from balf_b200.configs import config
from balf_b200.demo import demo_match
from balf_b200.model import get_model
from balf_b200.utils import test_utils
from balf_b200.third_party.hardnet.hardnet_pytorch import HardNet
from oracle import pipeline, weights

dev = torch.device("cuda:0")
cfg = test_utils.get_cfg_from_yaml_file(config.DEFAULT_CFG)
torch.manual_seed(0)
det = get_model.load_model(cfg["model"]).eval().to(dev)
sd = weights.detector_state_dict(0)
hn_sd = weights.hardnet_state_dict(0)
g = load_golden("r2_hardnet2048.npz")["out"]
x = torch.rand(2048, 1, 32, 32, generator=torch.Generator().manual_seed(4321)).to(dev)
for prec in ("fp16", "tf32", "fp32"):
    torch.manual_seed(0)
    hn = HardNet().eval().to(dev); hn.precision = prec
    with torch.inference_mode():
        d = hn(x).cpu().numpy()
    e = np.abs(d - g)
    print("hardnet %s vs reference golden (2048 patches): max |err| %.3e  mean %.3e  p99.9 %.3e  max angle err %.3e" %
          (prec, e.max(), e.mean(), np.quantile(e, 0.999), np.arccos(np.clip((d * g).sum(1), -1, 1)).max()))
args = config.default_test_args()
for (h, w) in ((200, 264), (900, 1200)):
    rgb1 = synth_u8(h, w, 21 if h == 200 else 1234)
    noise = np.random.default_rng(5).integers(-2, 3, rgb1.shape[:2])[..., None]
    rgb2 = np.clip(rgb1.astype(np.int64) + noise, 0, 255).astype(np.uint8)
    g1, g2 = rgb1[..., 0].copy(), rgb2[..., 0].copy()
    t0 = time.time()
    w1, w2 = pipeline.extract_matches(args, sd, hn_sd, rgb1, g1, rgb2, g2)
    print("oracle pair %dx%d: %d matches, %.1f s" % (h, w, len(w1), time.time() - t0))
    want = set(map(tuple, np.round(np.concatenate([w1, w2], 1), 2)))
    for dprec in ("auto", "tf32"):
        for hprec in ("tf32", "fp32"):
            d2 = copy.deepcopy(det); d2.precision = dprec
            torch.manual_seed(0)
            hn = HardNet().eval().to(dev); hn.precision = hprec
            p1, p2 = demo_match.extract_matches(args, rgb1, g1, rgb2, g2, d2, hn, dev)
            got = set(map(tuple, np.round(np.concatenate([p1, p2], 1), 2)))
            print("  detector %-5s hardnet %-5s: %d matches, recall %.4f precision %.4f" %
                  (dprec, hprec, len(got), len(got & want) / len(want), len(got & want) / max(len(got), 1)))
