"""Summarise `ncu --page source --csv` output: instruction mix, stall reasons, and
per-region totals (regions split at BAR.SYNC instructions)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in body)
tot_samp = sum(f(r, "# Samples") for r in body)
print("total warp-instr", tot_inst, "samples", tot_samp)
op = collections.Counter(); ops = collections.Counter()
for r in body:
    src = r[ix["Source"]].split()
    name = next((t for t in src if not t.startswith("@")), "?").split(".")[0]
    op[name] += f(r, "Instructions Executed"); ops[name] += f(r, "# Samples")
print("opcode: %instr  %samples")
for k, v in op.most_common(22): print(f"  {k:10s} {100*v/tot_inst:6.2f} {100*ops[k]/tot_samp:6.2f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = {s: sum(f(r, s) for r in body) for s in stalls}
tot = sum(st.values())
print("stall reasons (all samples):", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
# regions split at barriers
reg = []; cur = dict(inst=0, samp=0, start=0, conflicts=0, wave=0)
for n, r in enumerate(body):
    cur["inst"] += f(r, "Instructions Executed"); cur["samp"] += f(r, "# Samples")
    cur["conflicts"] += f(r, "L1 Wavefronts Shared Excessive"); cur["wave"] += f(r, "L1 Wavefronts Shared")
    if "BAR.SYNC" in r[ix["Source"]] or n == len(body) - 1:
        cur["end"] = n; reg.append(cur); cur = dict(inst=0, samp=0, start=n + 1, conflicts=0, wave=0)
print("regions between barriers: [sass lines] %instr %samples  smem wavefronts (excess)")
for c in reg:
    print(f"  [{c['start']:5d},{c['end']:5d}] {100*c['inst']/tot_inst:6.2f} {100*c['samp']/tot_samp:6.2f}   {c['wave']:.0f} ({c['conflicts']:.0f})")
# top stalled instructions
top = sorted(body, key=lambda r: -f(r, "# Samples"))[:14]
print("top sampled instructions:")
for r in top:
    reasons = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"  {100*f(r,'# Samples')/tot_samp:5.2f}%  {r[ix['Source']][:70]:70s} {reasons}")
