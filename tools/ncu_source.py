#!/usr/bin/env python
"""Top SASS lines of an .ncu-rep by stall samples / shared-memory excess wavefronts.
usage: ncu_source.py rep kernel_id(:::N) [n]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kid, "--print-source", "sass"],
                     capture_output=True, text=True).stdout.splitlines()
name = out[0]
rows = list(csv.DictReader(out[1:]))
def f(r, k):
    try: return float(r[k])
    except Exception: return 0.0
print(name)
tot = sum(f(r, "# Samples") for r in rows)
tot_inst = sum(f(r, "Instructions Executed") for r in rows)
print("total samples", tot, "warp instr", tot_inst, "sass lines", len(rows))
ex = sum(f(r, "L1 Wavefronts Shared Excessive") for r in rows); wf = sum(f(r, "L1 Wavefronts Shared") for r in rows)
print("smem wavefronts", wf, "excessive", ex)
print("--- top by samples")
for r in sorted(rows, key=lambda r: -f(r, "# Samples"))[:n]:
    st = {k: f(r, k) for k in r if k.startswith("stall_") and "Not Issued" not in k and f(r, k) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f'{r["Address"][-5:]} {r["Source"][:70]:70s} smp {f(r,"# Samples"):7.0f} ex {f(r,"Instructions Executed"):9.0f} ', top)
print("--- top by excessive smem wavefronts")
for r in sorted(rows, key=lambda r: -f(r, "L1 Wavefronts Shared Excessive"))[:n]:
    if f(r, "L1 Wavefronts Shared Excessive") <= 0: break
    print(f'{r["Address"][-5:]} {r["Source"][:70]:70s} wf {f(r,"L1 Wavefronts Shared"):9.0f} ideal {f(r,"L1 Wavefronts Shared Ideal"):9.0f} excess {f(r,"L1 Wavefronts Shared Excessive"):9.0f} ex {f(r,"Instructions Executed"):9.0f}')
# opcode histogram
import collections
h = collections.Counter()
for r in rows:
    op = r["Source"].split()[0] if r["Source"].split() else "?"
    if op.startswith("@"): op = r["Source"].split()[1]
    h[op.split(".")[0]] += f(r, "Instructions Executed")
print("--- executed warp-instructions by opcode")
for op, c in h.most_common(18): print(f"  {op:10s} {c:12.0f} {100*c/tot_inst:5.1f}%")
