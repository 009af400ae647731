"""Development helper: time the windowed NMS + top-k entry point alone on a resident batch of score maps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import balf_b200._capi as c

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mode = sys.argv[2] if len(sys.argv) > 2 else "windowed"
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(1)
prob = torch.rand(B, 512, 640, device=dev, generator=g) * 0.0125 + 0.01
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fn = (lambda: c.windowed_nms_topk(prob, 2048, crop=(16, 0, 480, 640))) if mode == "windowed" else \
     (lambda: c.greedy_nms_topk(prob, 2048, crop=(16, 0, 480, 640)))
for _ in range(3):
    fn()
c.profile_enable(True); c.profile_report(reset=True)
n = 10
for _ in range(n):
    flush.zero_()
    fn()
torch.cuda.synchronize()
rep = c.profile_report(reset=True)
c.profile_enable(False)
alg = B * (480 * 640 * 4)
for k, (cnt, ms) in rep.items():
    print("%-22s %8.1f us/launch  %8.1f GB/s (score-map bytes / time)" % (k, ms / cnt * 1e3, alg / (ms / cnt * 1e-3) / 1e9))
if mode == "windowed":        # A/B of the TMA tile load (balf_debug_set key 7)
    for tma in (0, 1):
        c.debug_set(7, tma)
        c.profile_enable(True); c.profile_report(reset=True)
        for _ in range(n):
            flush.zero_()
            fn()
        torch.cuda.synchronize()
        rep = c.profile_report(reset=True)
        c.profile_enable(False)
        print("tma=%d " % tma + "  ".join("%s %.1f us" % (k, ms / cnt * 1e3) for k, (cnt, ms) in rep.items()))
    c.debug_set(7, 1)
