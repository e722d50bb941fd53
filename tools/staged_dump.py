#!/usr/bin/env python
"""Check the phase A tables of the TMA-staged kernel (development aid).

    BEVIPM_ST_DUMP=/tmp/d.bin python tools/staged_dump.py <variant> V C Hf Wf Hb Wb B [seed]

Launches the kernel in dump mode (phase A only: every tile's shared-memory tables go to the file), decodes the
stage lists and verifies them against the oracle's sample positions: every stage's row copies are disjoint, fit the
slot and add up to the bytes the barrier expects; every (row, view, cell) that samples the view has exactly one
stage and its tap offsets land on the texels (x0, y0) .. (x0+1, y0+1) some copy delivers."""
import os
import struct
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "vision-based-spatio-temporal-analysis_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np

CELLS, ROWS_G, ROWS_R = 8, 32, 12
WIDTHS = [2, 4, 6, 8, 10, 12, 14, 16, 20, 24, 32, 48, 64]


def layout(V, R):
    up = lambda x, a: (x + a - 1) // a * a
    L, o = {}, 0
    L["nst_max"] = V * R
    L["gt"] = max(ROWS_G, R * 16)
    L["max_ops"] = V * L["gt"]
    L["wts"] = o; o += R * V * CELLS * 16
    L["ent"] = o; o += R * V * CELLS * 8
    L["wst"] = o; o += up(R * L["nst_max"] * 4, 16)
    L["sdesc"] = o; o += L["nst_max"] * 16
    L["sched"] = o; o += 256 * 8
    L["ops"] = o; o += L["max_ops"] * 8
    L["bars"] = o; o += 2 * 64 * 8
    L["misc"] = o; o += 128
    L["sH"] = o; o += V * 48
    L["ring"] = up(o, 128)
    return L


def check(path, x0, y0, Hf, Wf, B_frames=1, verbose=False):
    raw = open(path, "rb").read()
    V, NW, ring_bytes, cap, ring, gx, gz, tiles_x, tiles_y, fpc, Hb, Wb, Hf2, Wf2, C, es = struct.unpack("16i", raw[:64])
    L = layout(V, NW)
    assert L["ring"] == ring, (L["ring"], ring)
    body = np.frombuffer(raw[64:], np.uint8).reshape(gz, gx, ring)
    errs = 0
    kinds = {"multi-row": 0, "one-row": 0, "blocks": 0}
    staged_texels = 0
    nst_hist = []
    for cta in range(gx):
        sm = body[0, cta]
        ty, tx = divmod(cta, tiles_x)
        i0, j0 = ty * NW, tx * CELLS
        i32 = lambda off, n: sm[off:off + 4 * n].view(np.int32)
        nst = int(i32(L["misc"], 1)[0])
        nst_hist.append(nst)
        sdesc = i32(L["sdesc"], 4 * L["nst_max"]).reshape(-1, 4)[:nst]
        ops = i32(L["ops"], 2 * L["max_ops"]).reshape(-1, 2)
        wst = sm[L["wst"]:L["wst"] + 4 * NW * L["nst_max"]].view(np.uint32).reshape(NW, L["nst_max"])
        ent = i32(L["ent"], 2 * NW * V * CELLS).reshape(NW, V, CELLS, 2)
        stage_of = {}
        texmaps, regions = [], []
        last_v = -1
        for s, (nbytes, yw, opw, vw) in enumerate(sdesc):
            v, rowmask = vw & 0xff, (vw >> 8) & 0xffff
            o0, n = opw & 0xffff, opw >> 16
            if v < last_v:
                print(f"cta {cta}: stage {s} view {v} after view {last_v}"); errs += 1
            last_v = v
            if nbytes > cap:
                print(f"cta {cta} stage {s}: {nbytes} bytes (cap {cap})"); errs += 1
            tm = {}
            tot = 0
            isblk = False
            for q in range(n):
                a, bb = int(ops[o0 + q, 0]), int(ops[o0 + q, 1])
                x = ((a & 0xffff) ^ 0x8000) - 0x8000
                y = a >> 16
                o16 = (bb & 0xffff) * 16
                mp = bb >> 16
                if mp == 13:
                    w, h = 2, 2
                    isblk = True
                else:
                    w, h = WIDTHS[mp], 1
                tot += w * h * 512
                for yy in range(h):
                    for xx in range(w):
                        o = o16 + (yy * w + xx) * 512
                        if o in tm or o + 512 > max(nbytes, 0):
                            print(f"cta {cta} stage {s}: copy {q} overlaps / leaves the stage at {o}"); errs += 1
                        tm[o] = (x + xx, y + yy)
            if tot != nbytes:
                print(f"cta {cta} stage {s}: barrier expects {nbytes} bytes, copies deliver {tot}"); errs += 1
            staged_texels += tot // 512
            kinds["blocks" if isblk else ("multi-row" if bin(rowmask).count("1") > 1 else "one-row")] += 1
            texmaps.append(tm)
            for r in range(NW):
                if (rowmask >> r) & 1:
                    if (r, v) in stage_of:
                        print(f"cta {cta}: (row {r}, view {v}) in two stages"); errs += 1
                    stage_of[(r, v)] = s
        # the copy schedule: places inside the ring, nothing among the `back - 1` stages before an entry (cyclically) may
        # overlap it, predecessors in non-decreasing order
        chunks = -(-(C * es) // 512)
        n_items = min(fpc, B_frames) * chunks
        if nst:
            PL = max(1, min(n_items, 256 // nst)) * nst
            sched = i32(L["sched"], 2 * 256).reshape(-1, 2)[:PL]
            prev_pred = None
            for q in range(PL):
                off, back, nb = (int(sched[q, 0]) & 0xffff) << 7, int(sched[q, 0]) >> 16, int(sched[q, 1])
                if nb != int(sdesc[q % nst][0]) or off + nb > ring_bytes or back < 1:
                    print(f"cta {cta} schedule {q}: bytes {nb} (stage {int(sdesc[q % nst][0])}) at {off}, back {back}"); errs += 1
                for k in range(1, min(back, PL)):
                    q2 = (q - k) % PL
                    qo, qb = (int(sched[q2, 0]) & 0xffff) << 7, int(sched[q2, 1])
                    if nb > 0 and qb > 0 and qo < off + nb and off < qo + qb:
                        print(f"cta {cta} schedule {q}: overlaps entry {q2}, only {k} back (table says {back})"); errs += 1
                        break
                if prev_pred is not None and q - back < prev_pred:
                    print(f"cta {cta} schedule {q}: predecessor {q - back} older than the one before ({prev_pred})"); errs += 1
                prev_pred = q - back
        for r in range(NW):
            i = i0 + r
            for v in range(V):
                seen_want, rl_want = 0, 0
                prev = None
                for c in range(CELLS):
                    j = j0 + c
                    if i < Hb and j < Wb and -1 <= x0[v, i, j] <= Wf - 1 and -1 <= y0[v, i, j] <= Hf - 1:
                        seen_want |= 1 << c
                        cur = (int(x0[v, i, j]), int(y0[v, i, j]))
                        if cur != prev:
                            rl_want |= 1 << c
                        prev = cur
                    else:
                        prev = None
                words = [(s, int(wst[r, s])) for s in range(nst) if int(wst[r, s]) and (int(wst[r, s]) >> 24) == v]
                if not seen_want:
                    if words:
                        print(f"cta {cta} row {r} view {v}: stage words {words} but nothing seen"); errs += 1
                    continue
                if len(words) != 1 or (r, v) not in stage_of or words[0][0] != stage_of[(r, v)]:
                    print(f"cta {cta} row {r} view {v}: stage words {words}, stage list says {stage_of.get((r, v))}"); errs += 1
                    continue
                w = words[0][1]
                if (w & 0xff) != seen_want or ((w >> 8) & 0xff) != rl_want:
                    print(f"cta {cta} row {r} view {v}: seen/reload {w & 0xff:08b}/{(w >> 8) & 0xff:08b} want {seen_want:08b}/{rl_want:08b}"); errs += 1
                    continue
                tm = texmaps[stage_of[(r, v)]]
                for c in range(CELLS):
                    if not (seen_want >> c) & 1:
                        continue
                    X, Y = int(x0[v, i0 + r, j0 + c]), int(y0[v, i0 + r, j0 + c])
                    top, bot = int(ent[r, v, c, 0]), int(ent[r, v, c, 1])
                    got = [tm.get(top), tm.get(top + 512), tm.get(bot), tm.get(bot + 512)]
                    want = [(X, Y), (X + 1, Y), (X, Y + 1), (X + 1, Y + 1)]
                    if got != want:
                        print(f"cta {cta} (tile {ty},{tx}) row {r} view {v} cell {c}: taps {got} want {want} (ent {top},{bot}, stage {stage_of[(r, v)]})")
                        errs += 1
        if verbose:
            print(f"cta {cta}: {nst} stages", [tuple(int(z) for z in s) for s in sdesc])
    print(f"stages per tile: mean {np.mean(nst_hist):.2f} max {max(nst_hist)}; ring {ring_bytes} cap {cap}")
    return errs, kinds, staged_texels


def main():
    import torch
    from oracle import ipm_oracle as orc
    from test_gpu_parity import _rig_case, _run
    variant = int(sys.argv[1])
    V, C, Hf, Wf, Hb, Wb, B = [int(x) for x in sys.argv[2:9]]
    seed = int(sys.argv[9]) if len(sys.argv) > 9 else 5
    path = os.environ.setdefault("BEVIPM_ST_DUMP", "/tmp/staged_dump.bin")
    feats, K, Rt, xs, ys, img = _rig_case(B, V, C, (Hf, Wf), (Hb, Wb), seed=seed)
    _run(feats, K, Rt, xs, ys, img, "mean", True, variant=variant)
    ix, iy = orc.coords(K[:1], Rt[:1], xs, ys, (Hf, Wf), img)
    ix = ix.reshape(V, Hb, Wb); iy = iy.reshape(V, Hb, Wb)
    fin = np.isfinite(ix) & np.isfinite(iy)
    x0 = np.where(fin, np.floor(np.where(fin, ix, 0)), -2).clip(-2, Wf).astype(np.int64)
    y0 = np.where(fin, np.floor(np.where(fin, iy, 0)), -2).clip(-2, Hf).astype(np.int64)
    errs, kinds, tex = check(path, x0, y0, Hf, Wf, B_frames=B, verbose="-v" in sys.argv)
    seen = ((x0 >= -1) & (x0 <= Wf - 1) & (y0 >= -1) & (y0 <= Hf - 1)).sum()
    print(f"variant {variant} shape {(V, C, Hf, Wf, Hb, Wb, B)}: {errs} errors; stages {kinds}; staged texels per item {tex} "
          f"= {tex / max(seen, 1):.3f} per cell-view")


if __name__ == "__main__":
    main()
