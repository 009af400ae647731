"""Turn gpurun_out/{launches.csv, prof.ncu-rep, bench.json} into committed summaries under profiles/.

    python scripts/summarize_profiles.py <tag>      # e.g. r01a -> profiles/r01a_*.{md,csv,json}
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ("gpu__time_duration.sum", "sm__inst_executed_pipe_tensor", "smsp__issue_active.avg.pct", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg")


def launches(tag):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [l for l in open(path) if l.startswith('"')]
    agg = OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(rows))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        full = r["Kernel Name"]
        if "<" in full and "balf::" in full:
            name = full.split("(")[0]
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") == "us":
            ns *= 1e3
        elif r.get("Metric Unit") == "ms":
            ns *= 1e6
        a = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, tag + "_launches.md"), "w") as f:
        f.write("# %s -- ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n" % tag)
        f.write("| kernel | launches | total us | share | grid | block |\n|---|---:|---:|---:|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% | %s | %s |\n" % (name[:110], a[0], a[1] / 1e3, 100 * a[1] / total, a[2], a[3]))
    print("wrote", tag + "_launches.md")


def raw_page(report):
    """raw page of a report: the CSV exported on the GPU box (gpu_check.sh) or, if the report itself is here, ncu -i"""
    rep = os.path.join(OUT, report)
    pre = rep.replace(".ncu-rep", "_raw.csv")
    if os.path.exists(pre):
        return open(pre).read()
    if os.path.exists(rep):
        return subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return None


def full(tag, report="prof.ncu-rep", suffix="_ncu_full.md", what="the dominant detector kernels"):
    raw = raw_page(report)
    if raw is None:
        return
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print("no rows in report")
        return
    head, units = rows[0], rows[1]
    with open(os.path.join(PROF, tag + suffix), "w") as f:
        f.write("# %s -- `ncu --set full --clock-control none` of %s (one row per captured launch)\n\n" % (tag, what))
        for r in rows[2:]:
            d = dict(zip(head, r))
            f.write("## %s  grid %s block %s\n\n| metric | value | unit |\n|---|---:|---|\n" % (
                d.get("Kernel Name", "?")[:120], d.get("Grid Size", "?"), d.get("Block Size", "?")))
            for i, h in enumerate(head):
                if any(h.startswith(k) for k in KEYS):
                    f.write("| %s | %s | %s |\n" % (h, r[i], units[i]))
            f.write("\n")
    print("wrote", tag + suffix)


def traffic():
    """profiles/ncu_traffic.json: bench.py kernel name -> dram bytes (read + write) per launch, from the full captures."""
    import re
    out = {}
    for report in ("prof.ncu-rep", "prof_nms.ncu-rep"):
        raw = raw_page(report)
        if raw is None:
            continue
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        head, units = rows[0], rows[1]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            d = dict(zip(head, r))
            kn = d.get("Kernel Name", "")
            name = None
            m = re.search(r"tc_branch_kernel<\(?(?:int\))?(\d+), \(?(?:int\))?(\d+), \(?(?:int\))?(\d+)(?:, \(?(?:int\))?\d+)*>", kn)
            if m:
                name = "det_branch_%s_c%s" % ("grid" if m.group(3) == "0" else "block", m.group(2))
            m = re.search(r"tc_merge(?:_bulk)?_kernel<\(?(?:int\))?(\d+), \(?(?:int\))?(\d+)(?:, \(?(?:int\))?\d+)*>", kn)
            if m:
                name = "det_merge_c%s" % m.group(2)
            m = re.search(r"pool_kernel<\(?(?:int\))?(\d+)(?:, \(?(?:int\))?\d+)*>", kn)
            if m:
                name = "det_pool_c%s" % m.group(1)
            if "tc_head_kernel" in kn:
                name = "det_head"
            if "nms15_kernel" in kn or "nms15_tma_kernel" in kn:
                name = "nms_windowed"
            if "select_sort_kernel" in kn:
                name = "nms_select_sort"
            if not name or name in out:
                continue
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = head.index(key)
                tot += float(r[i].replace(",", "")) * mult.get(units[i], 1)
            out[name] = tot
    if out:
        path = os.path.join(PROF, "ncu_traffic.json")
        old = json.load(open(path)) if os.path.exists(path) else {}
        old.update(out)
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)
        print("wrote ncu_traffic.json", out)


def bench(tag):
    for name in ("bench.json", "bench_ref.json"):
        p = os.path.join(OUT, name)
        if os.path.exists(p):
            txt = [l for l in open(p).read().splitlines() if l.startswith("{")]
            if txt:
                json.dump(json.loads(txt[-1]), open(os.path.join(PROF, tag + "_" + name), "w"), indent=1)
                print("wrote", tag + "_" + name)


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    full(tag)
    full(tag, "prof_nms.ncu-rep", "_ncu_nms.md", "the windowed NMS + select/sort kernels (scripts/nms_bench.py, 64 maps)")
    full(tag, "prof_hn.ncu-rep", "_ncu_hardnet.md", "the HardNet tensor-core kernels (scripts/hn_bench.py, 4096 patches)")
    full(tag, "prof_greedy.ncu-rep", "_ncu_greedy.md", "the cell-based greedy NMS kernels (bench.py --nms greedy, 64 maps)")
    full(tag, "prof_smnn.ncu-rep", "_ncu_smnn.md", "the SMNN kernels (scripts/smnn_bench.py, 2048 x 2048)")
    full(tag, "prof_x3.ncu-rep", "_ncu_f16x3.md", "the detector kernels in split precision (bench.py --nms greedy: precision f16x3, 64 images)")
    traffic()
    bench(tag)
