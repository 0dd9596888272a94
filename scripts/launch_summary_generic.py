"""Per-kernel totals of the LAST step in an ncu launch list (--metrics gpu__time_duration.sum --csv): from the last launch of
the kernel named by argv[2] (default patchify_kernel) to the end of the list / the optimizer."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = [(r[ki], float(r[vi].replace(",", "")) * (1e-3 if r[ui].startswith("n") else 1.0)) for r in rows]
mark = sys.argv[2] if len(sys.argv) > 2 else "patchify"
starts = [i for i, s in enumerate(seq) if mark in s[0]]
step = seq[starts[-1]:]
tot, t = collections.OrderedDict(), 0.0
for name, us in step:
    short = name.split("(")[0].replace("void ", "").replace("countr::<unnamed>::", "")[:60]
    if "gemm_kernel" in name:
        short = "gemm_kernel<pair>" if "(bool)1" in name or "<1>" in name else "gemm_kernel<single>"
    d = tot.setdefault(short, [0, 0.0]); d[0] += 1; d[1] += us; t += us
print(f"step: {len(step)} launches, {t:.0f} us serialised")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"  {k:60s} {n:4d} {us:9.1f} us  {100 * us / t:5.1f} %")
