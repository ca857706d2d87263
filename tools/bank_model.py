#!/usr/bin/env python
"""Shared-memory bank model of the fused level kernel's fast path (validated against ncu's per-instruction wavefront
counts: 32-bit = whole warp, 64-bit = two half warps, 128-bit = four quarter warps; wavefronts of a group = the largest
number of distinct words that fall into one bank).  Prints wavefronts per tile and frame for a layout."""
import sys


def wf(addrs, width):
    """addrs: 32 entries (word address or None)."""
    g = {1: 32, 2: 16, 4: 8}[width]
    tot = 0
    for g0 in range(0, 32, g):
        banks = {}
        for a in addrs[g0:g0 + g]:
            if a is None:
                continue
            for k in range(width):
                banks.setdefault((a + k) % 32, set()).add(a + k)
        if banks:
            tot += max(len(v) for v in banks.values())
    return tot


def per_warps(fn, width, nthreads=256):
    tot = 0
    for w in range(8):
        addrs = [fn(w * 32 + l) if w * 32 + l < nthreads else None for l in range(32)]
        if any(a is not None for a in addrs):
            tot += wf(addrs, width)
    return tot


def model(P0S=84, P0PAD=24, P1S=38, P1PAD=4, P1O=1, SQS=72, s1map="cb,ch,rb", verbose=False):
    P0P = 76 * P0S + P0PAD
    P1P = 36 * P1S + P1PAD
    res = {}

    def s1_ids(tid):
        if s1map == "cb,ch,rb":
            return tid % 6, (tid // 6) % 3, tid // 18
        if s1map == "ch,cb,rb":
            return (tid // 3) % 6, tid % 3, tid // 18
        if s1map == "cb,rb,ch":
            return tid % 6, tid // 72, (tid // 6) % 12
        raise ValueError

    # ---- S1 loads (9 rows; the row only shifts all lanes by the same amount, but P0S matters mod 32 -> sum rows)
    t = 0
    for r in range(9):
        base = lambda tid: (lambda cb, ch, rb: ch * P0P + (6 * rb + r) * P0S + 12 * cb)(*s1_ids(tid))
        t += per_warps(lambda tid: base(tid) + 2, 2, 216)
        for k in (4, 8, 12):
            t += per_warps(lambda tid: base(tid) + k, 4, 216)
        t += per_warps(lambda tid: base(tid) + 16, 1, 216)
    res["S1 loads"] = t
    # ---- S1 stores: per output row o: word 6cb + c + P1O, c = 0..5: the compiler pairs aligned neighbours
    t = 0
    for o in range(3):
        base = lambda tid: (lambda cb, ch, rb: ch * P1P + (3 * rb + o) * P1S + 6 * cb + P1O)(*s1_ids(tid))
        c = 0
        while c < 6:
            if (P1O + c) % 2 == 0 and c + 1 < 6:
                t += per_warps(lambda tid: base(tid) + c, 2, 216); c += 2
            else:
                t += per_warps(lambda tid: base(tid) + c, 1, 216); c += 1
    res["S1 stores"] = t
    # ---- S2 main
    t = 0
    for ch in range(3):
        for a in range(4):
            for k in (0, 2):
                t += per_warps(lambda tid: ch * P1P + (2 * (tid >> 4) + 1 + a) * P1S + 2 * (tid & 15) + 1 + P1O + k, 2)
    res["S2 window"] = t
    t = 0
    for ch in range(3):
        for r in range(4):
            t += 2 * per_warps(lambda tid: ch * P0P + (4 * (tid >> 4) + 6 + r) * P0S + 4 * (tid & 15) + 8, 4)
    res["S2 g_l + park"] = t
    t = 0
    for r in range(4):
        for k in (0, 2):
            t += per_warps(lambda tid: (4 * (tid >> 4) + r + 2) * SQS + 4 * (tid & 15) + 2 + k, 2)
    res["S2 SQ store"] = t

    # ---- ring (132 quads)
    def quad(tid):
        if tid < 34: return 0, 2 * tid
        if tid < 68: return 66, 2 * (tid - 34)
        if tid < 100: return 2 + 2 * (tid - 68), 0
        return 2 + 2 * (tid - 100), 66
    t = 0
    for ch in range(3):
        for a in range(3):
            for b in range(3):
                t += per_warps(lambda tid: (lambda qq, pp: ch * P1P + (qq // 2 + a) * P1S + pp // 2 + P1O + b)(*quad(tid)), 1, 132)
        for r in range(2):
            t += per_warps(lambda tid: (lambda qq, pp: ch * P0P + (qq + 4 + r) * P0S + pp + 6)(*quad(tid)), 2, 132)
    for r in range(2):
        t += per_warps(lambda tid: (lambda qq, pp: (qq + r) * SQS + pp)(*quad(tid)), 2, 132)
    res["S2 ring"] = t
    # ---- S3
    t = 0
    for r in range(8):
        for k in (0, 4):
            t += per_warps(lambda tid: (4 * (tid >> 4) + r) * SQS + 4 * (tid & 15) + k, 4)
    res["S3 loads"] = t
    # ---- S0 owner (uint8)
    t = 0
    def chunks(tid):
        tx, ty = tid & 15, tid >> 4
        out = [(4 * ty + 6 + r, tx + 2) for r in range(4)]
        def second(k): return (6 + (k >> 2), (k & 3) if (k & 3) < 2 else 16 + (k & 3))
        if tid < 240:
            rr = tid // 20
            out.append((rr if rr < 6 else 64 + rr, tid % 20))
            out.append(second(tid + 16))
        else:
            out.append(second(tid - 240)); out.append(None)
        return out
    for i in range(6):
        for j in range(3):
            t += per_warps(lambda tid: (lambda c: None if c is None else c[0] * 60 + 3 * c[1] + j)(chunks(tid)[i]), 1)
        for ch in range(3):
            t += per_warps(lambda tid: (lambda c: None if c is None else ch * P0P + c[0] * P0S + 4 * c[1])(chunks(tid)[i]), 4)
    res["S0 owner"] = t
    res["total"] = sum(res.values())
    return res


if __name__ == "__main__":
    kw = {}
    for a in sys.argv[1:]:
        k, v = a.split("=")
        kw[k] = v if k == "s1map" else int(v)
    for k, v in model(**kw).items():
        print(f"{k:16s} {v}")
