"""Development helper: per-phase SM-clock timeline of CTA 0 of one tensor-core branch kernel.
    python scripts/tc_trace.py [level 0..3] [batch]
Stamps (csrc/detector_tc.cuh, TC_TRACE): 0 tile start; then per phase k: 3k+1 epilogue/loads done (before the MMA
sync), 3k+2 MMA issued + committed (thread 0) / passed the sync (epilogue thread), 3k+3 accumulator ready."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import balf_b200._capi as c
from balf_b200.model import get_model
from balf_b200.utils import test_utils
from balf_b200.configs import config

level = int(sys.argv[1]) if len(sys.argv) > 1 else 0
which = sys.argv[3] if len(sys.argv) > 3 else "branch"   # "branch": last writer = block branch; "merge": needs BALF trace of merge (runs last)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
cfg = test_utils.get_cfg_from_yaml_file(config.DEFAULT_CFG)
torch.manual_seed(0)
det = get_model.load_model(cfg["model"]).eval().to(dev)
det.precision = "tf32"
x = torch.rand(B, 3, 512, 640, device=dev)
with torch.inference_mode():
    det(x); det(x)
    torch.cuda.synchronize()
    buf = torch.zeros(8192, dtype=torch.int64, device=dev)
    # only the chosen level on the tensor-core path, so that its branch kernel is the last writer of the buffer
    c.debug_set(0, 1 << level)
    c.debug_set(2, 1 if which == "merge" else 0)
    c.debug_set_trace(buf.data_ptr())
    det(x[:B])
    torch.cuda.synchronize()
    c.debug_set_trace(0)
    c.debug_set(0, 0x1F)
t = buf.cpu()[:512].view(16, 16, 2).numpy()
names = ["start"] + [s for k in range(5) for s in ("epi%d" % k, "issued%d" % k, "acc%d" % k)]
for it in range(2, 8):
    row0, row1 = t[it, :, 0], t[it, :, 1]
    if row0[0] == 0:
        break
    print("tile %d  (thread 0 | last thread)  cycles since tile start, delta" % it)
    for p in range(16):
        d0 = row0[p] - row0[0]; d1 = row1[p] - row0[0]
        print("   %-8s %7d (+%5d) | %7d (+%5d)" % (names[p], d0, row0[p] - row0[p - 1] if p else 0, d1, row1[p] - row1[p - 1] if p else 0))
    if it + 1 < 16 and t[it + 1, 0, 0]:
        print("   tile total %d cycles" % (t[it + 1, 0, 0] - row0[0]))
