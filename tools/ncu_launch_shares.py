"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name.

    python tools/ncu_launch_shares.py gpurun_out/launches.csv > profiles/r1_train_launch_shares.txt
"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter()
cnt = collections.Counter()
for r in rd:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::|void ", "", name)[:90]
    t = float(r[ix["Metric Value"]].replace(",", ""))
    if r[ix["Metric Unit"]] in ("us", "usecond"):
        t *= 1e3
    elif r[ix["Metric Unit"]] in ("ms", "msecond"):
        t *= 1e6
    tot[name] += t
    cnt[name] += 1
total = sum(tot.values())
ours = sum(v for k, v in tot.items() if any(s in k for s in ("gemm", "gemm2", "dsp_", "attn_", "ln_", "bn_", "glu_", "colsum",
                                                               "sum_partials", "accumulate_partials", "adamw", "sumsq",
                                                               "dropout", "gelu_dropout", "ce_", "dwconv", "se_scale",
                                                               "group_mean", "nct_to_rows", "zero_invalid")))
print(f"# {len(lines) - 1} launches, total kernel time {total / 1e6:.3f} ms (serialised, cold cache: compare shares, not absolutes)")
print(f"# libeegx kernels: {100 * ours / total:.1f} % of the kernel time")
print(f"{'share %':>8} {'ms':>9} {'calls':>6}  kernel")
for k, v in tot.most_common(45):
    print(f"{100 * v / total:8.2f} {v / 1e6:9.3f} {cnt[k]:6d}  {k}")
