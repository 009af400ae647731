"""Development helper (GPU): where a configs[2] pair (demo_match.extract_matches, 1200x900, host buffers) spends its time --
wall clock per call, the library's per-kernel CUDA-event times, and a synchronised stage-by-stage host timeline.
    python scripts/pair_profile.py [precision]"""
import os, sys, time
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
from balf_b200.third_party.hardnet.hardnet_pytorch import HardNet

dev = torch.device("cuda:0")
cfg = test_utils.get_cfg_from_yaml_file(config.DEFAULT_CFG)
torch.manual_seed(0)
det = get_model.load_model(cfg["model"]).eval().to(dev)
det.precision = sys.argv[1] if len(sys.argv) > 1 else "auto"
hn = HardNet().eval().to(dev)
args = config.default_test_args()
rng = np.random.default_rng(0)
pairs = []
for i in range(5):
    g = torch.Generator().manual_seed(1234 + i)
    a_ = torch.randint(0, 256, (900, 1200, 1), generator=g, dtype=torch.uint8).expand(900, 1200, 3).contiguous().numpy()
    b_ = np.clip(a_.astype(np.int64) + rng.integers(-2, 3, (900, 1200, 1)), 0, 255).astype(np.uint8)
    pairs.append((a_, a_[..., 0].copy(), b_, b_[..., 0].copy()))
for _ in range(2):
    demo_match.extract_matches(args, *pairs[0], det, hn, dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for pr in pairs[1:]:
    demo_match.extract_matches(args, *pr, det, hn, dev)
torch.cuda.synchronize()
print("wall: %.3f ms per pair" % ((time.perf_counter() - t0) / 4 * 1e3))
capi.profile_enable(True); capi.profile_report(reset=True)
for pr in pairs[1:]:
    demo_match.extract_matches(args, *pr, det, hn, dev)
torch.cuda.synchronize()
rep = capi.profile_report(reset=True); capi.profile_enable(False)
tot = 0.0
for name, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print("  %-28s %5.1f launches/pair %8.3f ms/pair" % (name, n / 4, ms / 4)); tot += ms / 4
print("kernel sum: %.3f ms per pair" % tot)
# synchronised host timeline of one pair (each stage followed by a device synchronise)
def T(label, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print("  %-34s %7.3f ms" % (label, (time.perf_counter() - t) * 1e3)); return r
a, ag, b, bg = pairs[2]
st = T("np.stack", lambda: np.ascontiguousarray(np.stack((a, b))))
u8 = T("H2D rgb (pageable)", lambda: torch.from_numpy(st).to(dev, non_blocking=True))
out = T("detect_batch_device", lambda: demo_match.detect_batch_device(args, u8, det, "greedy"))
gd = T("H2D gray x2", lambda: [torch.from_numpy(np.ascontiguousarray(g)).to(dev, non_blocking=True) for g in (ag, bg)])
xy, _, dxdy, cnt = out
counts = T("cnt.cpu().tolist()", lambda: cnt.cpu().tolist())
pt = T("patches x2", lambda: [capi.extract_patches(gd[i], xy[i, :counts[i]].float() + dxdy[i, :counts[i]], float(args.s_mult), 32) for i in range(2)])
with torch.inference_mode():
    ds = T("hardnet", lambda: hn(torch.cat(pt)))
ids = T("smnn", lambda: capi.match_smnn(ds[:counts[0]], ds[counts[0]:], 0.99)[1].long())
T("gather + D2H", lambda: [(xy[i][ids[:, i]].cpu().numpy().astype(np.float64) + dxdy[i][ids[:, i]].cpu().numpy().astype(np.float64)) for i in range(2)])
