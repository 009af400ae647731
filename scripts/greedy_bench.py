"""Development helper (GPU): per-kernel times of the greedy NMS stage on the bench batch, for several counts of
multi-CTA rounds (balf_debug_set key 6).     python scripts/greedy_bench.py [B]"""
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
g = torch.Generator().manual_seed(1234)
u8 = torch.randint(0, 256, (B, 480, 640, 1), dtype=torch.uint8, generator=g).to(dev)
x, (top, left) = c.preprocess_u8(u8)
with torch.inference_mode():
    prob = det(x, precision="tf32")["prob"]
for rounds in (int(a) for a in (sys.argv[2:] or ["2", "4", "6", "8", "12"])):
    c.debug_set(6, rounds)
    for _ in range(3):
        c.greedy_nms_topk(prob, 2048, border=15, thr=0.001, radius=15, subpixel_ps=0, crop=(top, left, 480, 640))
    torch.cuda.synchronize()
    c.profile_enable(True); c.profile_report(True)
    n = 10
    for _ in range(n):
        out = c.greedy_nms_topk(prob, 2048, border=15, thr=0.001, radius=15, subpixel_ps=0, crop=(top, left, 480, 640))
    rep = c.profile_report(True); c.profile_enable(False)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        c.greedy_nms_topk(prob, 2048, border=15, thr=0.001, radius=15, subpixel_ps=0, crop=(top, left, 480, 640))
    e1.record()
    torch.cuda.synchronize()
    print("rounds %2d: %.1f us per call without per-kernel events" % (rounds, e0.elapsed_time(e1) / 20 * 1e3))
    tot = sum(v[1] for v in rep.values()) / n
    print("rounds %2d: stage %.1f us   " % (rounds, tot * 1e3) + "  ".join("%s %.1f (x%d)" % (k[4:], v[1] / n * 1e3, v[0] // n) for k, v in sorted(rep.items())), " kept/img %.0f" % out[3].float().mean().item())
c.debug_set(6, 3)
