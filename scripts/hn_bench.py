"""Development helper: HardNet forward timing (patches resident in HBM), per-kernel event times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import balf_b200._capi as c
from balf_b200.third_party.hardnet.hardnet_pytorch import HardNet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = HardNet().eval().to(dev)
for a in sys.argv[2:]:
    setattr(net, "precision", a)
x = torch.rand(n, 1, 32, 32, device=dev)
with torch.inference_mode():
    for _ in range(3):
        net(x)
    c.profile_enable(True); c.profile_report(reset=True)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        d = net(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    rep = c.profile_report(reset=True); c.profile_enable(False)
print("hardnet %d patches: %.3f ms  %.2f Mpatch/s  %.1f TFLOP/s" % (n, ms, n / ms / 1e3, n * 78.184e6 / ms / 1e9))
for k, (cnt, t) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print("   %-16s %8.3f ms/launch" % (k, t / cnt))
