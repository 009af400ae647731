"""Per CUDA source line: excessive shared-memory wavefronts (bank conflicts), global sectors and stall samples from an exported
source page (ncu -i rep --page source --print-source sass,cuda --csv > file.csv).
   python scripts/ncu_bank_summary.py file.csv [kernel-substring] [top-n]"""
import csv, io, sys, collections
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
lines = open(path).read().splitlines()
secs, i = [], 0
while i < len(lines):
    if lines[i].startswith('"File Path"'):
        f = lines[i]; fn = lines[i + 1]; j = i + 2
        while j < len(lines) and not lines[j].startswith('"File Path"'):
            j += 1
        secs.append((f, fn, lines[i + 2:j])); i = j
    else:
        i += 1
done = set()
for f, fn, body in secs:
    if want not in fn or fn in done:
        continue
    done.add(fn)
    rd = csv.reader(io.StringIO("\n".join(body)))
    hdr = next(rd)
    ci = {h: k for k, h in enumerate(hdr) if h != "Source"}
    agg = collections.OrderedDict()
    for r in rd:
        if len(r) < len(hdr) or r[2] != "-":
            continue
        def g(name):
            v = r[ci[name]] if name in ci else ""
            try: return int(float(v or 0))
            except ValueError: return 0
        key = (f.split(",")[1].strip('"').split("/")[-1], int(r[0]))
        a = agg.setdefault(key, [r[1][:100], 0, 0, 0, 0, 0])
        a[1] += g("L1 Wavefronts Shared Excessive"); a[2] += g("L1 Wavefronts Shared"); a[3] += g("# Samples")
        a[4] += g("Instructions Executed"); a[5] += g("L2 Theoretical Sectors Global Excessive")
    print(fn[:160])
    te = sum(a[1] for a in agg.values()); tw = sum(a[2] for a in agg.values()); ts = sum(a[3] for a in agg.values())
    print("  shared wavefronts %d, excessive %d (%.1f%%); samples %d" % (tw, te, 100.0 * te / max(tw, 1), ts))
    for (fl, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
        if a[1] == 0:
            break
        print("  %9d excess / %9d wf  %5.1f%% smp  %s:%d  %s" % (a[1], a[2], 100.0 * a[3] / max(ts, 1), fl, ln, a[0].strip()))
