"""Per CUDA source line of an exported source page (ncu -i rep --page source --print-source sass,cuda --csv): share of the stall
samples and of the executed warp instructions, with the two main stall reasons; kernel-wide stall breakdown on top.
   python scripts/ncu_stall_summary.py file.csv [section-index] [top-n]"""
import csv, io, sys, collections
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 16
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
names = []
for f, fn, body in secs:
    if fn not in names:
        names.append(fn)
fn0 = names[which]
agg = collections.OrderedDict(); tot = collections.Counter()
for f, fn, body in secs:
    if fn != fn0:
        continue
    rd = csv.reader(io.StringIO("\n".join(body))); hdr = next(rd)
    ci = {h: k for k, h in enumerate(hdr) if h != "Source"}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in rd:
        if len(r) < len(hdr) or r[2] != "-":
            continue
        key = (f.split(",")[1].strip('"').split("/")[-1], int(r[0]))
        a = agg.setdefault(key, [r[1][:95], 0, 0, collections.Counter()])
        a[1] += int(r[ci["# Samples"]] or 0); a[2] += int(r[ci["Instructions Executed"]] or 0)
        for h in stalls:
            v = int(r[ci[h]] or 0); a[3][h] += v; tot[h] += v
    break
ts = sum(a[1] for a in agg.values()); ti = sum(a[2] for a in agg.values())
print(fn0[17:140]); print("  samples", ts, "warp instructions", ti, "stalls:", [(k[6:], round(100 * v / max(ts, 1), 1)) for k, v in tot.most_common(8)])
for (fl, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print("  %5.1f%% smp %5.1f%% ins  %s:%d  %s | %s" % (100 * a[1] / max(ts, 1), 100 * a[2] / max(ti, 1), fl, ln, a[0].strip()[:70], [(k[6:], v) for k, v in a[3].most_common(2)]))
