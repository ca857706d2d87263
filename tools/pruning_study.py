#!/usr/bin/env python
"""Exactness-preserving pruning study: how much of the Laplacian / energy work could be
skipped without changing a single output bit?  Runs on the CPU with the oracle (test infrastructure).

For every pyramid level of a stack it measures
  * skip_warp  -- the fraction of (frame, 8-row x 64-column warp region) pairs whose energy stage the product kernel
                  skips (csrc/ss_kernels.cuh, ss_skip_factor): S * (1 + 2^-18) * max(window gray^2) < min(running best);
  * skip_pixel -- the same bound applied per pixel (upper limit of what any finer-grained skip could reach);
  * contenders -- per tolerance eps: the fraction of (frame, pixel) pairs with E_f >= (1 - eps) * max_g E_g, i.e. the
                  pairs an approximate pass with relative error eps/2 could NOT rule out and would have to re-evaluate
                  exactly (eps = 0: exact ties only).
usage: pruning_study.py [out.json]   (uses /root/reference/examples when present, plus the bench generator)"""
import glob
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyramid_oracle as po  # noqa: E402
from shinestacker_b200 import synthetic  # noqa: E402

EPS = (0.0, 4e-6, 4e-4, 1e-2)


def study(frames, name):
    po.build()
    k = po.make_kernel(0.4)
    factor = float(np.float64(np.sum(k.astype(np.float64))) * (1 + 2.0 ** -18))
    h, w = frames[0].shape[:2]
    L = po.num_levels(h, w)
    energies = [[] for _ in range(L)]
    squares = [[] for _ in range(L)]
    for f in frames:
        laps = po.laplacian_pyramid(f, L, k)
        for l in range(min(L, len(laps) - 1)):
            lap = laps[l]
            gray = (lap[2] * np.float32(.299) + (lap[0] * np.float32(.114) + lap[1] * np.float32(.587))).astype(np.float32)
            sq = gray * gray
            squares[l].append(sq)
            energies[l].append(po.filter5x5(sq, k))
    out = {"stack": name, "frames": len(frames), "size": [h, w], "levels": []}
    for l in range(L):
        if not energies[l]:
            continue
        E, S = np.array(energies[l]), np.array(squares[l])
        n, hh, ww = E.shape
        best = E[0].copy()
        warp_tot = warp_skip = 0
        px_skip = 0
        for f in range(1, n):
            # per-pixel bound: largest square in the 5x5 window
            pad = np.pad(S[f], 2, mode="reflect")
            wmax = np.zeros_like(S[f])
            for dy in range(5):
                for dx in range(5):
                    np.maximum(wmax, pad[dy:dy + hh, dx:dx + ww], out=wmax)
            px_skip += int((wmax * factor < best).sum())
            for y0 in range(0, hh - 7, 8):
                for x0 in range(0, ww - 63, 64):
                    m = S[f][max(0, y0 - 2):y0 + 10, max(0, x0 - 2):x0 + 66].max() * factor
                    warp_tot += 1
                    warp_skip += bool(m < best[y0:y0 + 8, x0:x0 + 64].min())
            best = np.maximum(best, E[f])
        emax = E.max(axis=0)
        cont = {}
        for eps in EPS:
            cont[str(eps)] = float((E >= emax * (1 - eps)).mean())
        out["levels"].append({"level": l, "size": [hh, ww],
                              "skip_warp": warp_skip / warp_tot if warp_tot else None,
                              "skip_pixel": px_skip / ((n - 1) * hh * ww) if n > 1 else None,
                              "contenders": cont})
    return out


def load_dir(d):
    import cv2
    files = sorted(glob.glob(os.path.join(d, "*")))
    return [cv2.imread(f, cv2.IMREAD_UNCHANGED if f.lower().endswith((".tif", ".tiff", ".png")) else cv2.IMREAD_COLOR)
            for f in files if f.lower().endswith((".jpg", ".jpeg", ".tif", ".tiff", ".png"))]


def main():
    res = []
    res.append(study(synthetic.make_stack(16, 768, 1024, np.uint8, seed=2), "bench generator, 16 x 1024x768 (64-px bands, as config 2)"))
    res.append(study(synthetic.make_noise_stack(8, 512, 768, np.uint8, seed=3), "iid noise, 8 x 768x512 (worst case)"))
    ex = "/root/reference/examples/input"
    for sub in ("img-jpg", "img-tif", "amber-flies"):
        d = os.path.join(ex, sub)
        if os.path.isdir(d):
            fr = [f for f in load_dir(d) if f is not None and f.ndim == 3]
            if fr:
                res.append(study(fr, f"reference example {sub}"))
    txt = json.dumps(res, indent=1)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(txt)
    for r in res:
        print(r["stack"], r["frames"], r["size"])
        for lv in r["levels"]:
            print("   level", lv["level"], lv["size"], "skip_warp %.3f skip_pixel %.3f" % (lv["skip_warp"] or 0, lv["skip_pixel"] or 0),
                  "contenders", {k: round(v, 4) for k, v in lv["contenders"].items()})


if __name__ == "__main__":
    main()
