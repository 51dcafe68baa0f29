"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (per-kernel launches, total,
average, share).  usage: python tools/launch_summary.py gpurun_out/r02_launches_bench.csv "title" "command" > profiles/x.md"""
import collections, csv, re, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
n = 0
for r in rows[hdr + 1:]:
    name = r[ki]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"void |bg::|\(anonymous namespace\)::|<unnamed>::", "", name).strip()
    if name.startswith("at::") or "elementwise_kernel" in name or "at::native" in r[ki]:
        name = "torch glue (" + re.sub(r".*native::", "", name)[:40] + ")"
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)   # -> us
    a = agg.setdefault(name[:70], [0, 0.0])
    a[0] += 1
    a[1] += v
    n += 1
tot = sum(v for _, v in agg.values())
print(f"# {sys.argv[2]}\ncommand: `{sys.argv[3]}`\n(serialised, cold-cache per-launch times: compare shares, not absolutes; raw CSV next to this file)\n")
print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
for name, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    if v / tot < 0.0005:
        continue
    print(f"| `{name}` | {c} | {v / 1e3:.2f} | {v / c:.1f} | {100 * v / tot:.1f}% |")
print(f"\ntotal {tot / 1e3:.1f} ms over {n} launches")
