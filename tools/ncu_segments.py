#!/usr/bin/env python
"""Split the SASS of one kernel of an .ncu-rep at its BAR.SYNC instructions and print, per segment, the executed
warp-instructions by opcode class, stall samples and shared-memory wavefronts (which stage costs what).
usage: ncu_segments.py rep kernel-name-regex [min_share]"""
import collections, csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout.splitlines()
# several kernels may match: take the first block
blocks, cur = [], None
for ln in out:
    if ln.startswith('"Kernel Name"'):
        cur = [ln]; blocks.append(cur)
    elif cur is not None:
        cur.append(ln)
import re
blk = [b for b in blocks if re.search(pat, b[0])][0]
print(blk[0])
rows = list(csv.DictReader(blk[1:]))
def f(r, k):
    try: return float(r[k])
    except Exception: return 0.0
def opclass(src):
    t = src.split()
    if not t: return "?"
    op = t[1] if t[0].startswith("@") else t[0]
    return op.split(".")[0]
tot = sum(f(r, "Instructions Executed") for r in rows)
tot_s = sum(f(r, "# Samples") for r in rows)
print(f"total warp-instr {tot:.0f}  samples {tot_s:.0f}  lines {len(rows)}")
segs, cur = [], []
for r in rows:
    cur.append(r)
    if "BAR.SYNC" in r["Source"]:
        segs.append(cur); cur = []
if cur: segs.append(cur)
for i, s in enumerate(segs):
    ex = sum(f(r, "Instructions Executed") for r in s)
    if ex < min_share * tot: continue
    smp = sum(f(r, "# Samples") for r in s)
    wf = sum(f(r, "L1 Wavefronts Shared") for r in s); wfx = sum(f(r, "L1 Wavefronts Shared Excessive") for r in s)
    h = collections.Counter()
    for r in s: h[opclass(r["Source"])] += f(r, "Instructions Executed")
    st = collections.Counter()
    for r in s:
        for k in r:
            if k.startswith("stall_") and "Not Issued" not in k: st[k] += f(r, k)
    a0, a1 = s[0]["Address"][-5:], s[-1]["Address"][-5:]
    print(f"\n== segment {i} [{a0}..{a1}] lines {len(s)}  instr {ex:.3g} ({100*ex/tot:.1f}%)  samples {smp:.0f} ({100*smp/tot_s:.1f}%)  "
          f"smem wavefronts {wf:.3g} excess {wfx:.3g}")
    print("   ops: " + "  ".join(f"{k} {100*v/ex:.1f}%" for k, v in h.most_common(14)))
    print("   stalls: " + "  ".join(f"{k[6:]} {100*v/max(smp,1):.0f}%" for k, v in st.most_common(6)))
if len(sys.argv) > 4:
    seg = segs[int(sys.argv[4])]
    print("--- memory instructions of segment", sys.argv[4])
    for r in seg:
        if f(r, "L1 Wavefronts Shared") > 0:
            print(f'{r["Address"][-5:]} {r["Source"][:60]:60s} ex {f(r,"Instructions Executed"):9.0f} wf {f(r,"L1 Wavefronts Shared"):10.0f} ideal {f(r,"L1 Wavefronts Shared Ideal"):10.0f}')
