#!/usr/bin/env python
"""Bench lines for BASELINE.json configs 1-4 (config 5 is bench.py): one stage or stage group over a ring of synthetic
pictures larger than the L2, device-resident, CUDA events on the context's stream, next to the reference C for the same
pictures on one host thread and on all host threads (one picture per thread), with a parity check of one picture.

  1  ALF luma classify + filter, chroma filter, CC-ALF on 1920x1080 10-bit
  2  in-loop chain deblock V + H -> SAO -> ALF on 1920x1080 10-bit
  3  inverse transform batch (dequant, LFNST, DCT-II / DST-VII / DCT-VIII 4x4..64x64, add_residual) on 3840x2160
  4  inter prediction batch (8/4-tap MC, bi, weighted, GPM, DMVR, BDOF, PROF) on 3840x2160
Usage: bench_configs.py [--configs 1,2,3,4] [--steps K] [--warmup W]; prints one JSON line per config."""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffvvc_b200 import abi, device, lib  # noqa: E402
import bench  # noqa: E402

ALGO_BYTES = {1: 6.0, 2: 18.0, 3: 12.0, 4: 9.0}     # per luma pixel, SURVEY.md 8(d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    peak, peak_src = bench.load_peaks()
    kind, fns = bench.cpu_lib()
    cores = len(os.sched_getaffinity(0))
    ctx = lib.Context(0)
    torch.cuda.set_stream(ctx.torch_stream())
    for cfg in [int(c) for c in args.configs.split(",")]:
        w, h = (1920, 1080) if cfg <= 2 else (3840, 2160)
        frames = 64 if cfg <= 2 else 16
        inp = bench.Inputs(w, h, distinct=2)
        g1, gr = inp.g1, abi.FrameGeom(w, h, batch=frames)
        keep = []

        def up(a):
            t, p = device.to_device(a)
            keep.append(t)
            return p

        reps = frames // inp.distinct + 1
        ring = [np.ascontiguousarray(np.concatenate([p] * reps)[:frames]) for p in inp.ref_planes]
        src, dst = device.DeviceFrames(gr, planes=ring), device.DeviceFrames(gr)
        cat = lambda parts: np.concatenate([parts[k % inp.distinct] for k in range(frames)])
        if cfg <= 2:
            md = abi.VVCCudaDeblockMaps()
            for d in range(2):
                for c in range(3):
                    rows, pitch = abi.deblock_map_shape(g1, d, c)
                    md.edge[d][c] = up(np.concatenate([inp.maps[k % inp.distinct][d][c] for k in range(frames)]))
                    md.pitch[d][c], md.rows[d][c], md.size[d][c] = pitch, rows, rows * pitch
            p_sao, p_alf, p_sets = up(cat(inp.sao)), up(cat(inp.alf)), up(inp.sets)
            desc = abi.inloop_desc(md, p_sao, p_alf, p_sets)
            run = (lambda: ctx.alf_frame(dst.desc, src.desc, p_alf, p_sets, 0)) if cfg == 1 else (lambda: ctx.inloop_frame(dst.desc, src.desc, desc))
            launches = 1 if cfg == 1 else 4
        elif cfg == 3:
            tb_parts, co_parts, off = [], [], 0
            for k in range(frames):
                t = inp.tbs[k % inp.distinct].copy()
                t["pic"] = k
                t["coeff_offset"] += off
                off += len(inp.coeffs[k % inp.distinct])
                tb_parts.append(t)
                co_parts.append(inp.coeffs[k % inp.distinct])
            tbs, coeffs = np.concatenate(tb_parts), np.concatenate(co_parts)
            cd = abi.coeffs_desc(up(coeffs), len(coeffs), abi.COEFF_DENSE32, up(cat(inp.quant)), up(inp.scaling))
            p_tbs = up(tbs)
            run = lambda: ctx.itx_frame_q(src.desc, cd, p_tbs, len(tbs), 15)       # residuals accumulate in place: timing only
            launches = 4
        else:
            pbs = np.concatenate([inp.records(k, frames, k) for k in range(frames)])
            p_pbs, p_wp, p_prof = up(pbs), up(inp.wp), up(inp.prof)
            run = lambda: ctx.inter_frame(dst.desc, src.desc, p_pbs, len(pbs), p_wp, p_prof, None)
            launches = 7
        for _ in range(max(args.warmup, 3)):
            run()
        ctx.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        a.record()
        for _ in range(args.steps):
            run()
        b.record()
        ctx.sync()
        ms = a.elapsed_time(b) / args.steps
        n_launch = (ctx.launches - l0) // args.steps
        # parity of picture 0 against the reference C (config 3 accumulates in place over the timed passes: checked on a fresh ring)
        scratch = [abi.alloc_planes(g1) for _ in range(3)]
        refs2 = [p[:inp.distinct] for p in ring]

        def cpu_picture(i, sc):
            cur, x, y = sc
            if cfg == 1:
                fns["alf_frame"](abi.frame_from_numpy(g1, y), abi.frame_from_numpy(g1, [p[i:i + 1] for p in inp.ref_planes]), inp.alf[i].ctypes.data, inp.sets.ctypes.data, 0)
                return y
            if cfg == 2:
                m = abi.deblock_maps_desc(g1, inp.maps[i])
                s0 = abi.frame_from_numpy(g1, [p[i:i + 1] for p in inp.ref_planes])
                fns["deblock_frame"](abi.frame_from_numpy(g1, x), s0, C.byref(m), 1)
                fns["deblock_frame"](abi.frame_from_numpy(g1, y), abi.frame_from_numpy(g1, x), C.byref(m), 0)
                fns["sao_frame"](abi.frame_from_numpy(g1, x), abi.frame_from_numpy(g1, y), inp.sao[i].ctypes.data)
                fns["alf_frame"](abi.frame_from_numpy(g1, y), abi.frame_from_numpy(g1, x), inp.alf[i].ctypes.data, inp.sets.ctypes.data, 0)
                return y
            if cfg == 3:
                for c in range(3):
                    cur[c][:] = inp.ref_planes[c][i:i + 1]
                co = abi.coeffs_desc(inp.coeffs[i].ctypes.data, inp.coeffs[i].size, abi.COEFF_DENSE32, inp.quant[i].ctypes.data, inp.scaling.ctypes.data)
                fns["itx_frame_q"](abi.frame_from_numpy(g1, cur), C.byref(co), inp.tbs[i].ctypes.data, len(inp.tbs[i]), 15)
                return cur
            gd = abi.FrameGeom(w, h, batch=inp.distinct)
            fns["inter_frame"](abi.frame_from_numpy(g1, cur), abi.frame_from_numpy(gd, refs2), inp.pbs[i].ctypes.data, len(inp.pbs[i]),
                               inp.wp.ctypes.data, inp.prof.ctypes.data, None)
            return cur

        want = cpu_picture(0, scratch)
        if cfg == 3:
            fresh = device.DeviceFrames(gr, planes=ring)
            ctx.itx_frame_q(fresh.desc, cd, p_tbs, len(tbs), 15)
            ctx.sync()
            got = fresh.to_numpy()
        elif cfg == 4:
            # picture 0 of the ring reads ring slots (0, 1) = the two distinct contents
            got = dst.to_numpy()
        else:
            got = dst.to_numpy()
        equal = all(np.array_equal(got[c][0, :, :g1.plane_wh(c)[0]], want[c][0, :, :g1.plane_wh(c)[0]]) for c in range(3))
        # the reference C: one thread, then one picture per host thread
        t0 = time.perf_counter()
        cpu_picture(1, scratch)
        one = time.perf_counter() - t0
        scr = [[abi.alloc_planes(g1) for _ in range(3)] for _ in range(cores)]
        ts = [threading.Thread(target=cpu_picture, args=(t % inp.distinct, scr[t])) for t in range(cores)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        allc = time.perf_counter() - t0
        px = w * h * frames
        achieved = ALGO_BYTES[cfg] * px / (ms * 1e-3) / 1e9
        print(json.dumps({
            "config": cfg, "workload": {1: "ALF luma classify + 7x7, chroma 5x5, CC-ALF", 2: "deblock V + H -> SAO -> ALF / CC-ALF",
                                        3: "dequant + LFNST + inverse transforms 2x2..64x64 + add_residual", 4: "inter prediction: MC, bi, weighted, GPM, DMVR, BDOF, PROF"}[cfg],
            "width": w, "height": h, "bit_depth": 10, "pictures_per_launch": frames, "ring_mb": round(src.nbytes / 1e6),
            "metric": "mpix_per_s", "value": px / (ms * 1e-3) / 1e6, "ms_per_launch": ms, "us_per_picture": ms * 1e3 / frames,
            "gpu_launches_per_step": n_launch, "steps": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_luma_px": ALGO_BYTES[cfg]},
            "cpu_baseline": {"kind": kind, "one_thread_mpix_per_s": w * h / one / 1e6, "all_threads_mpix_per_s": w * h * cores / allc / 1e6, "cores": cores,
                             "cpu_model": bench.cpu_model(), "sample": "one picture on one thread; one picture per thread on all threads"},
            "parity": {"pictures": 1, "equal": bool(equal), "against": kind}}), flush=True)
        del src, dst, keep
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
