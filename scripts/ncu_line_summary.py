"""Per CUDA source line: stall samples and executed warp instructions (needs -lineinfo + --import-source on).
   python scripts/ncu_line_summary.py report.ncu-rep [kernel-substring] [top-n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
# split into (file, function) sections
secs, i = [], 0
while i < len(lines):
    if lines[i].startswith('"File Path"'):
        f = lines[i]; fn = lines[i + 1]; j = i + 2
        while j < len(lines) and not lines[j].startswith('"File Path"'):
            j += 1
        secs.append((f, fn, lines[i + 2:j])); i = j
    else:
        i += 1
seen = None
agg = collections.OrderedDict()
for f, fn, body in secs:
    if want not in fn:
        continue
    if seen is None:
        seen = fn
    if fn != seen:
        continue
    rd = csv.reader(io.StringIO("\n".join(body)))
    hdr = next(rd)
    ci = {h: k for k, h in enumerate(hdr) if h not in ("Source",)}
    for r in rd:
        if len(r) < len(hdr) or r[2] != "-":     # keep the per-line summary rows (Address == "-")
            continue
        key = (f.split(",")[1].strip('"').split("/")[-1], int(r[0]))
        a = agg.setdefault(key, [r[1][:90], 0, 0])
        a[1] += int(r[ci["# Samples"]] or 0)
        a[2] += int(r[ci["Instructions Executed"]] or 0)
print(seen)
ts = sum(a[1] for a in agg.values()); ti = sum(a[2] for a in agg.values())
print("samples %d, warp instructions %d" % (ts, ti))
for (fl, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * a[1] / max(ts, 1), 100.0 * a[2] / max(ti, 1), fl, ln, a[0]))
