"""Summarise an `ncu --set full --import-source on` report of the attention kernel: pipe utilisation (raw page) and the
warp-state samples inside the softmax loop, split at barriers / TMEM loads (source page).
usage: python tools/ncu_stalls.py gpurun_out/prof_attn_b256.ncu-rep > profiles/r01_attn_stall_breakdown.md"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
page = lambda p: list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", p, "--csv"], capture_output=True, text=True).stdout)))

raw = page("raw")
hdr, units, row = raw[0], raw[1], raw[2]
get = lambda k: (row[hdr.index(k)], units[hdr.index(k)])
print(f"# Attention kernel: pipes and warp states (`{rep.split('/')[-1]}`, {row[hdr.index('Kernel Name')][:60]}…)\n")
print("| metric | value |\n|---|---|")
for k in ("gpu__time_duration.sum", "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio"):
    if k in hdr:
        v, u = get(k)
        print(f"| `{k}` | {v} {u} |")

src = page("source")
h = src[1]
ix = {n: i for i, n in enumerate(h)}
data = src[2:]
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
op = lambda s: (re.sub(r"^@!?U?P\d+\s+", "", s.strip()).split() or [""])[0]
n = lambda r, k: int(r[ix[k]] or 0)
tot = sum(n(r, "# Samples") for r in data)
ops = [op(r[ix["Source"]]) for r in data]
first = next(i for i, o in enumerate(ops) if o.startswith("LDTM")) - 40
last = max(i for i, o in enumerate(ops) if o.startswith(("MEMBAR", "STTM"))) + 12     # end of the key-block loop body
loop = data[first:last]
ls = sum(n(r, "# Samples") for r in loop)
print(f"\nWarp-state samples: {tot} in the kernel, {ls} ({100 * ls / tot:.1f} %) inside the softmax key-block loop "
      f"(the rest: mbarrier spin loops of the producer / MMA warps, idle warps at the final barrier).\n")
agg = collections.Counter()
for r in loop:
    for s in stalls:
        agg[s] += n(r, s)
print("| state (softmax loop) | share |\n|---|---|")
for k, v in agg.most_common():
    if v:
        print(f"| {k.replace('stall_', '')}{' (= issuing)' if k == 'stall_selected' else ''} | {100 * v / ls:.1f} % |")

print("\n| segment starts at | instructions | samples | top states | instruction mix |\n|---|---|---|---|---|")
seg, cur, name = [], [], "loop top"
for r in loop:
    cur.append(r)
    o = op(r[ix["Source"]])
    if o.startswith(("BAR", "SYNCS", "LDTM", "MEMBAR", "WARPSYNC")):
        seg.append((name, cur))
        cur, name = [], f"{o} @{r[ix['Address']][-4:]}"
seg.append((name, cur))
for name, c in seg:
    s = sum(n(r, "# Samples") for r in c)
    if s < 0.02 * ls:
        continue
    a = collections.Counter()
    for r in c:
        for st in stalls:
            a[st] += n(r, st)
    mix = collections.Counter(op(r[ix["Source"]]) for r in c)
    print(f"| `{name}` | {len(c)} | {100 * s / ls:.1f} % | " + ", ".join(f"{k.replace('stall_', '')} {100 * v / s:.0f} %" for k, v in a.most_common(3))
          + " | " + ", ".join(f"{k} {v}" for k, v in mix.most_common(4)) + " |")
