"""`ncu --metrics gpu__time_duration.sum --csv` log -> markdown table of kernels (launches, us, share).
    python tools/ncu_launch_list.py launches.csv "title line" [last_n_launches] >> profiles/rNN_launch_list.md"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
body = rows[1:]
if len(sys.argv) > 3:
    body = body[-int(sys.argv[3]):]   # the last N launches = the steady-state scan(s)
agg = collections.OrderedDict()
for r in body:
    name = r[ik].split("(")[0].replace("void ", "").replace("gm::", "")
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print(f"\n# {sys.argv[2] if len(sys.argv) > 2 else ''}\nlaunches {sum(v[0] for v in agg.values())}  total {tot:.1f} us\n")
print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
