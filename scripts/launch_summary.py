"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of `bench.py --steps 1 --no-graph`:
the last fine-tune step — from its weight_refresh launch (the first kernel of a step: the decoder weight re-cast and the
exemplar CNN run on the side stream BEFORE the encoder's patchify) to the optimizer update — per-kernel totals."""
import csv, sys, collections

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ki, vi, gi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Metric Unit")
seq = [(r[ki], r[gi], float(r[vi].replace(",", "")) * (1e-3 if r[ui].startswith("n") else 1.0)) for r in rows]
starts = [i for i, s in enumerate(seq) if "weight_refresh" in s[0]] or [i for i, s in enumerate(seq) if "patchify" in s[0]]
step = seq[starts[-1]:]
# cut the per-kernel timing loop / input generation that follows the step
end = len(step)
for i, s in enumerate(step):
    if "multi_tensor_apply" in s[0] or "adam_finish" in s[0]:
        end = i + 1
step = step[:end]
verbose = len(sys.argv) > 2
tot = collections.OrderedDict()
t = 0.0
for i, (name, grid, us) in enumerate(step):
    short = name.split("(")[0].replace("void ", "").replace("countr::<unnamed>::", "")[:48]
    if "gemm_kernel" in name:
        short = "gemm_kernel<pair>" if "(bool)1" in name or "<1>" in name else "gemm_kernel<single>"
    d = tot.setdefault(short, [0, 0.0])
    d[0] += 1
    d[1] += us
    t += us
    if verbose:
        print(f"{i:4d} {short:50s} {grid:>14s} {us:8.1f} {t:9.0f}")
print(f"step: {len(step)} launches, {t:.0f} us serialised")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:50s} {n:4d} {us:9.1f} us  {100 * us / t:5.1f} %")
