"""CPU-only analysis: how much of each igemm problem's MEASURED time (profiles/r01_igemm_tile_sweep.json,
B200, CUDA events) is tile quantisation — partially filled 128-row tiles and a partially filled last
wave of the persistent grid — according to the tile plan the launcher picks (`ctrlv_igemm_plan`, no CUDA
calls).  Ranks the problems a k-split / stream-K scheduler would help.

    python scripts/quantization_loss.py            # table on stdout (kept: profiles/r01_tile_quantisation.txt)

Model (stated, not measured): a launch takes ceil(units / slots) waves of equal length, units = m-tiles x
n-tiles (CTA pairs: ceil(m-tiles / 2) x n-tiles on 74 slots); the recoverable share is
1 - (units / (waves x slots)) x (rows / (m-tiles x 128)).  It is an upper bound: a real stream-K pays a
partial-sum exchange, and short launches are partly fill/drain latency rather than waves.
"""
import ctypes
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlv_b200 import _lib  # noqa: E402

SMS = 148
PEAK = 1398.3e12  # MEASURED_PEAKS.json bf16_tflops_sustained


def problem(op, key):
    """-> list of (X, Y, Z, K, N) GEMMs of one call (the up-conv is four phase convs)."""
    if op == "linear":
        M, K, N = key[:3]
        return [(M, 1, 1, K, N)]
    if op == "conv3x3":
        frames, H, W, stride, C, SC, N = key
        return [(W // stride, H // stride, frames, 9 * C + SC, N)]
    if op == "conv_t3":
        B, T, HW, C, N = key
        return [(HW, T, B, 3 * C, N)]
    if op == "upconv3x3":
        frames, H, W, C, N = key
        return [(W, H, frames, 4 * C, N)] * 4
    raise ValueError(op)


def main():
    lib = _lib.load(build_if_missing=False)
    sweep = json.load(open(os.path.join(ROOT, "profiles", "r01_igemm_tile_sweep.json")))
    rows = []
    for e in sweep:
        gemms = problem(e["op"], e["key"])
        X, Y, Z, K, N = gemms[0]
        d = _lib.IgemmDesc()
        d.nsrc, d.nseg = 1, 1
        d.X, d.Y, d.Z, d.N, d.K = X, Y, Z, N, K
        d.seg[0].nchunk = (K + 63) // 64
        box = (ctypes.c_int32 * 3)(); bn = ctypes.c_int32(); cg = ctypes.c_int32()
        if lib.ctrlv_igemm_plan(ctypes.byref(d), SMS, box, ctypes.byref(bn), ctypes.byref(cg)) != 0:
            raise RuntimeError(lib.ctrlv_last_error().decode())
        m_tiles = math.ceil(X / box[0]) * math.ceil(Y / box[1]) * math.ceil(Z / box[2])
        n_tiles = math.ceil(N / bn.value)
        units = math.ceil(m_tiles / cg.value) * n_tiles
        slots = SMS // cg.value
        waves = math.ceil(units / slots)
        eff_wave = units / (waves * slots)
        eff_rows = X * Y * Z / (m_tiles * 128)
        eff_cols = N / (n_tiles * bn.value)
        flops = 2.0 * X * Y * Z * K * N * len(gemms)
        us = e["auto_us"]
        frac = flops / PEAK / (us * 1e-6)
        loss = us * (1.0 - eff_wave * eff_rows * eff_cols)
        rows.append(dict(op=e["op"], key=e["key"], n=e["count"], us=us, bn=bn.value, cg=cg.value, m_tiles=m_tiles,
                         n_tiles=n_tiles, waves=units / slots, eff=eff_wave * eff_rows * eff_cols, frac=frac,
                         loss_total_us=loss * e["count"]))
    rows.sort(key=lambda r: -r["loss_total_us"])
    tot = sum(r["us"] * r["n"] for r in rows)
    lost = sum(r["loss_total_us"] for r in rows)
    print(f"igemm problems of one step: {len(rows)} shapes, {sum(r['n'] for r in rows)} calls, {tot / 1e3:.2f} ms measured in isolation")
    print(f"upper bound of the tile-quantisation share (model above): {lost / 1e3:.2f} ms = {100 * lost / tot:.1f} %")
    print(f"{'calls':>5} {'us/call':>8} {'of peak':>7} {'bn':>4} {'cg':>2} {'m x n tiles':>12} {'waves':>6} {'fill':>5} {'lost us/step':>12}  problem")
    for r in rows[:28]:
        print(f"{r['n']:5d} {r['us']:8.1f} {r['frac']:7.2f} {r['bn']:4d} {r['cg']:2d} {r['m_tiles']:6d} x{r['n_tiles']:4d} "
              f"{r['waves']:6.2f} {r['eff']:5.2f} {r['loss_total_us']:12.1f}  {r['op']} {tuple(r['key'])}")


if __name__ == "__main__":
    main()
