"""Development helper: CUDA-event timings of the main entry points (not the bench of record)."""
import sys, os, time
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


u8 = torch.randint(0, 256, (B, 480, 640, 1), dtype=torch.uint8, device=dev)
x, (top, left) = c.preprocess_u8(u8)
with torch.inference_mode():
    t = timeit(lambda: det(x))
    print("detector B=%d 512x640: %.3f ms  (%.1f img/s, %.2f TFLOP/s)" % (B, t, B / t * 1e3, B * 39.157e9 / t / 1e9))
    prob = det(x)["prob"]
t = timeit(lambda: c.windowed_nms_topk(prob, 2048, crop=(top, left, 480, 640)))
print("windowed nms+topk B=%d: %.3f ms (%.1f GB/s algorithmic)" % (B, t, B * (480 * 640 * 4 + 2048 * 16) / t / 1e6))
t = timeit(lambda: c.greedy_nms_topk(prob, 2048, crop=(top, left, 480, 640), subpixel_ps=4))
print("greedy nms+topk B=%d: %.3f ms" % (B, t))
