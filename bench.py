#!/usr/bin/env python
"""Benchmark of record for the BALF inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch 64] [--nms windowed]

One "step" = one pass of the hot path over one batch of synthetic input on every rank:
uint8 images -> /255 + pad (D0) -> detector forward (D1-D3) -> border mask + NMS + top-2048
(P1-P8) -> keypoint records (and, for N > 1, one NCCL all-gather of the records).
Workload = BASELINE.json configs[1]: batch 64 synthetic grayscale 640x480 per GPU.

Prints ONE JSON line (rank 0).  ``value`` = whole-job images/s with the uint8 batch already in
HBM; ``e2e`` = the same metric through the public host-buffer API (``demo_match.DetectPipeline``, the streaming ``detect_batch``:
pinned host uint8 in, keypoint records back on the host, copies inside the timed region).
``--impl reference`` times the reference's CPU algorithm (the oracle restatement; the reference
itself cannot travel to the GPU box) on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

H, W, K_FEATURES = 480, 640, 2048
METRIC = "images/sec detect+NMS+top-k @640x480"
FLOP_PER_PADDED_PIXEL = 119496          # SURVEY.md 8d (Linear / matmul MACs x 2 only)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    p.add_argument("--nms", default="windowed", choices=["windowed", "greedy"])
    p.add_argument("--precision", default=None, help="detector precision (default: the module's)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--chunk", type=int, default=0, help="development: images per internal detector pass (library default if 0)")
    return p.parse_args()


def workload_name(a):
    return ("configs[1]: batch %d synthetic grayscale 640x480 per GPU, detector + %s NMS + top-%d"
            % (a.batch, a.nms, K_FEATURES))


def synth_batch(batch, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, H, W, 1), generator=g, dtype=torch.uint8)


def model_cfg():
    from balf_b200.configs import config
    from balf_b200.utils import test_utils
    return test_utils.get_cfg_from_yaml_file(config.DEFAULT_CFG)["model"]


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_pipeline(a):
    """-> callable(image index) running the reference algorithm (oracle port) on one image."""
    from oracle import pipeline, postproc, detector as odet          # the CPU arm only
    from balf_b200.model import get_model                             # weights only (same init as the reference)
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in get_model.load_model(model_cfg()).state_dict().items()}
    args = pipeline.default_args(sub_pixel=False)
    imgs = synth_batch(8, 1234).expand(8, H, W, 3).contiguous().numpy()

    def one(i):
        im = imgs[i % len(imgs)]
        score = pipeline.score_map(sd, im)
        if a.nms == "windowed":
            return postproc.windowed_detect(score, args.border_size, args.nms_size, K_FEATURES)
        return pipeline.detect_from_score_map(args, score)[0]
    return one


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    one = cpu_pipeline(a)
    per_step = 2
    for i in range(a.warmup):
        one(i)
    t0 = time.perf_counter()
    for s in range(a.steps):
        for j in range(per_step):
            one(s * per_step + j)
    dt = time.perf_counter() - t0
    v = a.steps * per_step / dt
    sample = "%d images per step, batch 1 each (the reference's own batching), %d steps" % (per_step, a.steps)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": workload_name(a), "nms": a.nms, "k": K_FEATURES},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def cpu_baseline(a):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    one = cpu_pipeline(a)
    one(0)
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 32):
        one(n)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d images of the same workload, batch 1 each, %.1f s" % (n, dt)}


# ------------------------------------------------------------------------------------------ roofline
def level_dims():
    hp, wp = 512, 640
    return [(3, 32, hp * wp), (32, 64, hp * wp // 4), (64, 128, hp * wp // 16), (128, 256, hp * wp // 64)]


def kernel_work(name, images):
    """Algorithmic work of ONE step's launches of kernel `name`: (flops, bytes).
    Detector kernels: Linear / token-mixing MACs x 2 counted once (recomputation is not credited) and the
    fp32 tensors each kernel must read and write once (x, u, v, r, q: DESIGN.md section 3).  NMS: 4*H*W bytes per
    image (+ 16*K for the select/sort kernel)."""
    lv = {32: 0, 64: 1, 128: 2, 256: 3}
    if name.startswith("det_"):
        tail = name.rsplit("_c", 1)
        c = int(tail[1]) if len(tail) == 2 and tail[1].isdigit() else None
        if c in lv:
            cin, ch, px = level_dims()[lv[c]]
            n = images * px
            if "branch" in name:                       # half of dense1, branch dense1, token mix, dense2; x in, u' out
                return n * (2 * cin * ch + 2 * ch * ch + 4 * ch * ch + 128 * ch + 2 * ch * ch), n * 4 * (cin + ch)
            if "merge" in name:                        # conv.0, dense2 (2C->C), conv1, conv2; x, u', v' in, r, q out
                return n * (2 * cin * ch + 4 * ch * ch + 4 * ch * ch), n * 4 * (cin + 4 * ch)
        if name == "det_head":                         # r, q in; prob out (64 values per 8x8 cell)
            n = images * level_dims()[3][2]
            return n * (2 * 256 * 256 + 2 * 256 * 65), n * 4 * (2 * 256 + 64)
        if name == "det_pool":                         # r, q in, pooled out, stages 1-3
            return 0, sum(images * px * 4 * (2 * ch + ch // 4) for _, ch, px in level_dims()[:3])
        return 0, 0
    if name.startswith("nms_windowed"):
        return 0, images * (4 * H * W)
    if name == "nms_select_sort":
        return 0, images * 16 * K_FEATURES
    if name == "preprocess_u8":
        return 0, images * (H * W + 12 * 512 * 640)
    return 0, 0


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm": p["hbm_gbs"], "tensor": p["bf16_tflops_sustained"], "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tensor": 1400.0, "src": "fallback (B200_PROFILING.md)"}


def ncu_traffic(name):
    """dram__bytes_read + dram__bytes_write per launch of kernel `name` from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by scripts/summarize_profiles.py), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get(name)
    return None


def roofline_of(name, launches, total_ms, steps, images, pk):
    """The bound is the kernel's own: arithmetic intensity (algorithmic flops / algorithmic bytes) against the machine
    balance for tf32 operands (half the bf16 tensor peak / HBM peak) decides between "tensor" and "hbm"."""
    flops, nbytes = kernel_work(name, images)
    per_launch_s = total_ms / max(launches, 1) * 1e-3
    per = steps / max(launches, 1)
    balance = (pk["tensor"] / 2 * 1e12) / (pk["hbm"] * 1e9)
    if flops == 0 and nbytes == 0:
        return {"kernel": name, "bound": "latency", "launch_ms": per_launch_s * 1e3, "launches_per_step": launches / steps}
    kind = "tensor" if nbytes == 0 or (flops and flops / nbytes > balance) else "hbm"
    out = {"kernel": name, "bound": kind, "launch_ms": per_launch_s * 1e3, "launches_per_step": launches / steps,
           "peak_source": pk["src"], "traffic": ncu_traffic(name),
           "algorithmic": {"flops_per_launch": flops * per, "bytes_per_launch": nbytes * per,
                           "intensity_flop_per_byte": (flops / nbytes) if nbytes else None, "tf32_balance": balance}}
    if kind == "tensor":
        ach = flops * per / per_launch_s / 1e12
        out.update({"achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
                    "frac_of_tf32_peak": ach / (pk["tensor"] / 2)})
    else:
        ach = nbytes * per / per_launch_s / 1e9
        out.update({"achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]})
        if flops:
            out["tensor_tflops"] = flops * per / per_launch_s / 1e12
    return out


def extras(dev, pk):
    """Secondary rows of the hot path (BASELINE.json configs[2] shapes), not part of `value`:
    HardNet on 4096 patches and SMNN on 2048 x 2048 descriptors, device-resident, CUDA-event timed."""
    import balf_b200._capi as capi
    from balf_b200.third_party.hardnet.hardnet_pytorch import HardNet
    torch.manual_seed(0)
    hn = HardNet().eval().to(dev)
    x = torch.rand(4096, 1, 32, 32, device=dev)

    def timed(fn, n=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    from balf_b200.configs import config
    from balf_b200.demo import demo_match
    from balf_b200.model import get_model
    torch.manual_seed(0)
    det = get_model.load_model(model_cfg()).eval().to(dev)       # outside inference_mode: its weight cache reads ._version
    out = {}
    with torch.inference_mode():
        ms = timed(lambda: hn(x))
        tf = 4096 * 78.184e6 / (ms * 1e-3) / 1e12
        out["hardnet"] = {"patches": 4096, "ms": ms, "patches_per_s": 4096 / (ms * 1e-3), "tflops": tf, "dtype": hn.precision,
                          "frac_of_tf32_peak": tf / (pk["tensor"] / 2), "peak": "half of the bf16 figure (%s)" % pk["src"]}
        d1 = torch.nn.functional.normalize(torch.randn(2048, 128, device=dev), dim=1)
        d2 = torch.nn.functional.normalize(d1 + 0.05 * torch.randn(2048, 128, device=dev), dim=1)
        ms = timed(lambda: capi.match_smnn(d1, d2, 0.99))
        out["smnn"] = {"n1": 2048, "n2": 2048, "ms": ms, "pairs_per_s": 2048 * 2048 / (ms * 1e-3),
                       "note": "includes the device->host read of the match count"}
        # F1: 2048 patches of one 900 x 1200 image (level-1 pyramid build + bilinear gather); algorithmic bytes =
        # source read + level write/read + patches written
        gray = torch.randint(0, 256, (900, 1200), dtype=torch.uint8, device=dev)
        kp = torch.stack([torch.rand(2048, device=dev) * 1100 + 50, torch.rand(2048, device=dev) * 800 + 50], 1)
        ms = timed(lambda: capi.extract_patches(gray, kp, 60.0, 32))
        nbytes = 900 * 1200 + 2 * 4 * 450 * 600 + 4096 * 2048
        out["patches"] = {"keypoints": 2048, "image": "900x1200", "ms": ms, "bound": "hbm", "achieved_GBs": nbytes / (ms * 1e-3) / 1e9,
                          "frac": nbytes / (ms * 1e-3) / 1e9 / pk["hbm"]}
        # BASELINE.json configs[4] in miniature: 3-level pyramid (0.7x), 8192 keypoints per image, 8 x 1024 x 1024
        u8 = torch.randint(0, 256, (8, 1024, 1024, 1), dtype=torch.uint8, device=dev)
        margs = config.default_test_args(sub_pixel=False, num_features=8192)
        ms = timed(lambda: demo_match.detect_multiscale_batch_device(margs, u8, det, scale=0.7, levels=3), n=3)
        out["multiscale"] = {"workload": "8 x 1024x1024, 3 levels x 0.7, windowed NMS, top-8192 merged", "ms": ms,
                             "images_per_s": 8 / (ms * 1e-3)}
    return out


# ------------------------------------------------------------------------------------------ main arm
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    import balf_b200._capi as capi
    from balf_b200.configs import config
    from balf_b200.demo import demo_match
    from balf_b200.model import get_model
    from balf_b200.sharding import gather_keypoints

    if a.chunk:
        capi.debug_set(1, a.chunk)
    torch.manual_seed(0)
    det = get_model.load_model(model_cfg()).eval().to(dev)
    if a.precision:
        det.precision = a.precision
    args = config.default_test_args(sub_pixel=False, num_features=K_FEATURES)
    host = synth_batch(a.batch, 1234 + rank).pin_memory()
    u8 = host.to(dev)

    def step_device():
        xy, sc, _, cnt = demo_match.detect_batch_device(args, u8, det, a.nms)
        if world > 1:
            return gather_keypoints(xy, sc, cnt)
        return xy, sc, cnt

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    capi.profile_enable(True)
    capi.profile_report(reset=True)
    l0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        out = step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = capi.launch_count() - l0
    prof = capi.profile_report(reset=True)
    capi.profile_enable(False)
    clocks = sampler.stop() if sampler else None

    # end to end through the host-buffer API: every step's uint8 batch goes pinned host -> device and its keypoint
    # records come back device -> host inside the timed region.  DetectPipeline is the streaming form of
    # demo_match.detect_batch: the copy of step i+1 overlaps the kernels of step i.
    pipe = demo_match.DetectPipeline(args, det, dev, a.nms)
    prev = None
    for _ in range(max(a.warmup, 3) + pipe.depth):      # warm up in the pipelined pattern itself: the staging ring and the
        cur = pipe.submit(host)                          # second in-flight device batch are allocated here, not in the timed passes
        if prev is not None:
            res = pipe.result(prev)
        prev = cur
    res = pipe.result(prev)
    e2e_passes = []
    for _ in range(2):                  # two passes of K steps; the faster one is reported, both are listed
        barrier()
        t0 = time.perf_counter()
        prev = None
        for _ in range(a.steps):
            cur = pipe.submit(host)
            if prev is not None:
                res = pipe.result(prev)
            prev = cur
        res = pipe.result(prev)
        torch.cuda.synchronize()
        e2e_passes.append((time.perf_counter() - t0) * 1e3)
    t_e2e = min(e2e_passes)
    d2h = sum(int(r.nbytes) for r in res)

    t = torch.tensor([ms, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, t_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    images = a.batch * world
    pk = peaks()
    total_kernel_ms = sum(v[1] for v in prof.values())
    kernels = {}
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        r = roofline_of(k, v[0], v[1], a.steps, a.batch, pk)
        kernels[k] = {"launches_per_step": v[0] / a.steps, "ms_per_step": v[1] / a.steps, "share": v[1] / total_kernel_ms,
                      "bound": r["bound"], "achieved": r.get("achieved"), "unit": r.get("unit"), "frac": r.get("frac")}
    top = max(prof.items(), key=lambda kv: kv[1][1])
    roof = roofline_of(top[0], top[1][0], top[1][1], a.steps, a.batch, pk)
    nms_name = "nms_windowed" if a.nms == "windowed" else "nms_greedy_rounds"
    roof_nms = roofline_of(nms_name, prof[nms_name][0], prof[nms_name][1], a.steps, a.batch, pk) if nms_name in prof else None
    det_ms = sum(v[1] for k, v in prof.items() if k.startswith("det_")) / a.steps
    line = {
        "metric": METRIC, "value": images * a.steps / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": det.precision, "data": "synthetic",
        "config": {"workload": workload_name(a), "batch_per_gpu": a.batch, "nms": a.nms, "k": K_FEATURES,
                   "padded": "512x640", "weights": "random-init torch.manual_seed(0)",
                   "l2": "no explicit flush: each step streams >2 GB of activations per GPU, far above the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": images * a.steps / (t_e2e * 1e-3), "unit": "images/s", "passes_ms": [round(x, 3) for x in e2e_passes],
                "h2d_bytes_per_step": int(host.numel()), "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": roof,
        "roofline_nms": roof_nms,
        "detector": {"ms_per_step": det_ms, "tflops": a.batch * 512 * 640 * FLOP_PER_PADDED_PIXEL / (det_ms * 1e-3) / 1e12
                     if det_ms else None},
        "kernels": kernels,
    }
    if world == 1:
        line["extras"] = extras(dev, pk)
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(a)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
