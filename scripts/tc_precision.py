"""Development helper (GPU): precision of the tf32 detector path against the library's own fp32 path (rel ~2e-6 vs the
reference): max / mean / SIGNED mean relative error of the score map, and keypoint agreement of the greedy and the
windowed extraction, on the seeds the parity tests use.     python scripts/tc_precision.py"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from balf_b200.model import get_model
from balf_b200.utils import test_utils
from balf_b200.configs import config
from balf_b200.demo import demo_match

dev = torch.device("cuda:0")
cfg = test_utils.get_cfg_from_yaml_file(config.DEFAULT_CFG)
torch.manual_seed(0)
det = get_model.load_model(cfg["model"]).eval().to(dev)
d32 = copy.deepcopy(det); d32.precision = "fp32"
dtc = copy.deepcopy(det); dtc.precision = "tf32"
args = config.default_test_args(sub_pixel=False)
tot_i = tot_n = wi = wn = 0
for seed in (1234, 1, 2, 3, 4, 5):
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (1, 480, 640), generator=g, dtype=torch.uint8)
    im = u8.permute(1, 2, 0).expand(480, 640, 3).contiguous().numpy()
    x, _ = __import__("balf_b200._capi", fromlist=["x"]).preprocess_u8(torch.from_numpy(im[None]).to(dev))
    with torch.inference_mode():
        p32 = d32(x)["prob"].double().cpu().numpy()
        ptc = dtc(x)["prob"].double().cpu().numpy()
    rel = (ptc - p32) / p32
    a = demo_match.detect(args, im, d32, dev); b = demo_match.detect(args, im, dtc, dev)
    sa = set(map(tuple, a[:, :2])); sb = set(map(tuple, b[:, :2]))
    tot_i += len(sa & sb); tot_n += len(sa)
    u8d = torch.from_numpy(im[None, :, :, :1].copy()).to(dev)
    xa, _, _, ca = demo_match.detect_batch_device(args, u8d, d32, nms="windowed")
    xb_, _, _, cb = demo_match.detect_batch_device(args, u8d, dtc, nms="windowed")
    wa = set(map(tuple, xa[0, :int(ca[0])].cpu().numpy().tolist())); wb = set(map(tuple, xb_[0, :int(cb[0])].cpu().numpy().tolist()))
    wi += len(wa & wb); wn += len(wa)
    print("seed %4d  max rel %.3e  mean |rel| %.3e  signed mean %+.3e   greedy agreement %d/%d = %.4f" %
          (seed, np.abs(rel).max(), np.abs(rel).mean(), rel.mean(), len(sa & sb), len(sa), len(sa & sb) / len(sa)))
print("total greedy agreement %d/%d = %.4f   windowed top-%d agreement %d/%d = %.4f" % (tot_i, tot_n, tot_i / tot_n, args.num_features, wi, wn, wi / wn))
