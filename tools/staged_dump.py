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

CELLS, ROWS_T, ROWS_R = 8, 32, 12
WIDTHS = [2, 4, 6, 8, 10, 12, 14, 16, 20, 24, 32, 48, 64]


def layout(V, R, S, D):
    up = lambda x, a: (x + a - 1) // a * a
    L, o = {}, 0
    L["wts"] = o; o += R * V * CELLS * 16
    L["ent"] = o; o += R * V * CELLS * 8
    L["mask"] = o; o += up(R * V * 4, 16)
    L["stg"] = o; o += V * R * 16
    max_ops = V * max(ROWS_T, R * ROWS_R)
    L["ops"] = o; o += max_ops * 8
    L["misc"] = o; o += 128
    L["sH"] = o; o += V * 48
    L["bars"] = o; o += up(2 * D * 8, 128)
    L["ring"] = up(o, 128)
    return L


def check(path, x0, y0, Hf, Wf, verbose=False):
    raw = open(path, "rb").read()
    V, NW, S, D, ring, gx, gz, tiles_x, tiles_y, fpc, Hb, Wb, Hf2, Wf2, C, es = struct.unpack("16i", raw[:64])
    L = layout(V, NW, S, D)
    assert L["ring"] == ring, (L["ring"], ring)
    body = np.frombuffer(raw[64:], np.uint8).reshape(gz, gx, ring)
    errs = 0
    kinds = {"tile": 0, "rows": 0, "blocks": 0}
    staged_texels = 0
    for cta in range(gx):
        sm = body[0, cta]
        ty, tx = divmod(cta, tiles_x)
        i0, j0 = ty * NW, tx * CELLS
        i32 = lambda off, n: sm[off:off + 4 * n].view(np.int32)
        nst = int(i32(L["misc"], 1)[0])
        stg = i32(L["stg"], 4 * V * NW).reshape(-1, 4)[:nst]
        ops = i32(L["ops"], 2 * V * max(ROWS_T, NW * ROWS_R)).reshape(-1, 2)
        mask = sm[L["mask"]:L["mask"] + 4 * NW * V].view(np.uint32).reshape(NW, V)
        ent = i32(L["ent"], 2 * NW * V * CELLS).reshape(NW, V, CELLS, 2)
        stage_of = {}
        texmaps = []
        last_v = -1
        for s, (v, rowmask, nbytes, opw) in enumerate(stg):
            o0, n = opw & 0xffff, opw >> 16
            if v < last_v:
                print(f"cta {cta}: stage {s} view {v} after view {last_v}"); errs += 1
            last_v = v
            tm = {}
            tot = 0
            for q in range(n):
                a, b = int(ops[o0 + q, 0]), int(ops[o0 + q, 1])
                x = ((a & 0xffff) ^ 0x8000) - 0x8000
                y = a >> 16
                off = (b & 0xffff) * 16
                mp = b >> 16
                if mp == 13:
                    w, h = 2, 2
                    kinds["blocks"] += 1
                else:
                    w, h = WIDTHS[mp], 1
                tot += w * h * 512
                for yy in range(h):
                    for xx in range(w):
                        o = off + (yy * w + xx) * 512
                        if o in tm or o + 512 > S:
                            print(f"cta {cta} stage {s}: copy {q} overlaps / leaves the slot at {o}"); errs += 1
                        tm[o] = (x + xx, y + yy)
            if tot != nbytes:
                print(f"cta {cta} stage {s}: barrier expects {nbytes} bytes, copies deliver {tot}"); errs += 1
            staged_texels += tot // 512
            if bin(rowmask).count("1") > 1 or NW == 1:
                kinds["tile"] += 1
            else:
                kinds["rows"] += 1
            texmaps.append(tm)
            for r in range(NW):
                if (rowmask >> r) & 1:
                    if (r, v) in stage_of:
                        print(f"cta {cta}: (row {r}, view {v}) in two stages"); errs += 1
                    stage_of[(r, v)] = s
        for r in range(NW):
            i = i0 + r
            for v in range(V):
                seen_want = 0
                for c in range(CELLS):
                    j = j0 + c
                    if i < Hb and j < Wb and -1 <= x0[v, i, j] <= Wf - 1 and -1 <= y0[v, i, j] <= Hf - 1:
                        seen_want |= 1 << c
                seen = int(mask[r, v]) & 0xffff
                if seen != seen_want:
                    print(f"cta {cta} row {r} view {v}: seen mask {seen:08b} want {seen_want:08b}"); errs += 1
                    continue
                if seen and (r, v) not in stage_of:
                    print(f"cta {cta} row {r} view {v}: seen {seen:08b} but no stage"); errs += 1
                    continue
                if not seen:
                    continue
                tm = texmaps[stage_of[(r, v)]]
                for c in range(CELLS):
                    if not (seen >> c) & 1:
                        continue
                    X, Y = int(x0[v, i0 + r, j0 + c]), int(y0[v, i0 + r, j0 + c])
                    top, bot = int(ent[r, v, c, 0]), int(ent[r, v, c, 1])
                    got = [tm.get(top), tm.get(top + 512), tm.get(bot), tm.get(bot + 512)]
                    want = [(X, Y), (X + 1, Y), (X, Y + 1), (X + 1, Y + 1)]
                    if got != want:
                        print(f"cta {cta} (tile {ty},{tx}) row {r} view {v} cell {c}: taps {got} want {want} (ent {top},{bot}, stage {stage_of[(r, v)]})")
                        errs += 1
        if verbose:
            print(f"cta {cta}: {nst} stages", [tuple(int(z) for z in s) for s in stg])
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
    errs, kinds, tex = check(path, x0, y0, Hf, Wf, verbose="-v" in sys.argv)
    seen = ((x0 >= -1) & (x0 <= Wf - 1) & (y0 >= -1) & (y0 <= Hf - 1)).sum()
    print(f"variant {variant} shape {(V, C, Hf, Wf, Hb, Wb, B)}: {errs} errors; stages {kinds}; staged texels per item {tex} "
          f"= {tex / max(seen, 1):.3f} per cell-view")


if __name__ == "__main__":
    main()
