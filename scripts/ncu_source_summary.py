"""Aggregate `ncu --page source --csv` of a report: per kernel, instruction mix by opcode and
warp-stall samples by reason.   python scripts/ncu_source_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import csv, io, subprocess, sys, collections, re

rep = sys.argv[1]
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in txt.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
blk = blocks[want]
print(blk[0][:200])
rows = list(csv.DictReader(io.StringIO("\n".join(blk[1:]))))
ops = collections.Counter(); samp = collections.Counter(); stalls = collections.Counter()
tot = 0
for r in rows:
    src = r["Source"].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDTM", "STTM", "LDS", "STS", "LDG", "STG", "MUFU", "BAR", "SYNCS", "UTC")) else op.split(".")[0]
    n = int(r["Instructions Executed"] or 0)
    ops[op] += n; tot += n
    samp[op] += int(r["# Samples"] or 0)
    for k, v in r.items():
        if k.startswith("stall_") and "Not Issued" not in k and v and v != "0":
            stalls[k] += int(v)
print("total warp instructions", tot)
ts = sum(samp.values())
print("%-14s %12s %7s %9s" % ("opcode", "warp-instr", "share", "samples%"))
for op, n in ops.most_common(28):
    print("%-14s %12d %6.1f%% %8.1f%%" % (op, n, 100.0 * n / tot, 100.0 * samp[op] / max(ts, 1)))
print("stall reasons (all samples):")
s = sum(stalls.values())
for k, v in stalls.most_common(12):
    print("  %-28s %6.1f%%" % (k, 100.0 * v / s))
