#!/usr/bin/env python
"""Experiment: the whole reconstruction chain of two picture groups issued on two contexts (two streams), so that kernels
of different stages overlap on the SMs, against the same pictures on one stream.
Usage: exp_two_ctx.py [pictures per context] [contexts] [passes]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffvvc_b200 import abi, device, lib  # noqa: E402
from bench import Inputs  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 10
inp = Inputs(3840, 2160, distinct=2)
g1 = inp.g1
gring = abi.FrameGeom(3840, 2160, batch=frames)
keep = []


def build(ctx):
    torch.cuda.set_stream(ctx.torch_stream())

    def up(a):
        t, p = device.to_device(a)
        keep.append(t)
        return p

    reps = frames // inp.distinct + 1
    refs = device.DeviceFrames(gring, planes=[np.ascontiguousarray(np.concatenate([p] * reps)[:frames]) for p in inp.ref_planes])
    cur, out = device.DeviceFrames(gring), device.DeviceFrames(gring)
    pbs = np.concatenate([inp.records(k, frames, k) for k in range(frames)])
    tb_parts, co_parts, off = [], [], 0
    for k in range(frames):
        t = inp.tbs[k % inp.distinct].copy()
        t["pic"] = k
        t["coeff_offset"] += off
        off += len(inp.coeffs[k % inp.distinct])
        tb_parts.append(t)
        co_parts.append(inp.coeffs[k % inp.distinct])
    tbs, coeffs = np.concatenate(tb_parts), np.concatenate(co_parts)
    md = abi.VVCCudaDeblockMaps()
    for d in range(2):
        for c in range(3):
            rows, pitch = abi.deblock_map_shape(g1, d, c)
            md.edge[d][c] = up(np.concatenate([inp.maps[k % inp.distinct][d][c] for k in range(frames)]))
            md.pitch[d][c], md.rows[d][c], md.size[d][c] = pitch, rows, rows * pitch
    desc = abi.VVCCudaReconDesc()
    desc.pbs, desc.n_pbs, desc.wp, desc.n_wp, desc.prof, desc.n_prof = up(pbs), len(pbs), up(inp.wp), len(inp.wp), up(inp.prof), len(inp.prof)
    desc.log2_transform_range = 15
    desc.coeffs, desc.n_coeffs, desc.tbs, desc.n_tbs = up(coeffs), len(coeffs), up(tbs), len(tbs)
    desc.coeff_format = abi.COEFF_DENSE32
    desc.quant = up(np.concatenate([inp.quant[k % inp.distinct] for k in range(frames)]))
    desc.scaling = up(inp.scaling)
    desc.lmcs_inv_lut = up(inp.inv_lut)
    desc.inloop.deblock = C.pointer(md)
    desc.inloop.sao = up(np.concatenate([inp.sao[k % inp.distinct] for k in range(frames)]))
    desc.inloop.alf = up(np.concatenate([inp.alf[k % inp.distinct] for k in range(frames)]))
    desc.inloop.alf_sets = up(inp.sets)
    keep.extend([md, refs, cur, out])
    return refs, cur, out, desc


ctxs = [lib.Context(0) for _ in range(n_ctx)]
sets = [build(c) for c in ctxs]
for c in ctxs:
    c.sync()


def run(n):
    for _ in range(n):
        for c, (refs, cur, out, desc) in zip(ctxs, sets):
            c.recon_frame(out.desc, cur.desc, refs.desc, desc)
    for c in ctxs:
        c.sync()


run(3)
torch.cuda.synchronize()
t0 = time.perf_counter()
run(passes)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
px = 3840 * 2160 * frames * n_ctx * passes
print("pictures/context %d contexts %d: %.3f ms per pass, %.1f Mpix/s" % (frames, n_ctx, dt / passes * 1e3, px / dt / 1e6))
