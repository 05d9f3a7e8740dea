#!/usr/bin/env python
"""All-intra reconstruction by wavefronts (vvc_cuda_intra_recon_frame): time per picture ring on the GPU and the CPU oracle's time
for the same blocks in decoding order; parity of every picture.  Usage: bench_intra.py [width height pictures]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffvvc_b200 import abi, device, lib, synth  # noqa: E402

w = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
h = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8
geom = abi.FrameGeom(w, h, batch=batch)
t0 = time.perf_counter()
case = synth.intra_picture(geom, seed=11)
t_synth = time.perf_counter() - t0
planes = abi.alloc_planes(geom, fill=512)
oracle = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
FP = C.POINTER(abi.VVCCudaFrame)
oracle.vvco_intra_recon_frame.argtypes = [FP, C.c_void_p, C.c_void_p, C.POINTER(abi.VVCCudaCoeffs), C.c_void_p, C.c_void_p, C.c_int, C.c_int]
oracle.vvco_intra_recon_frame.restype = None
want = [p.copy() for p in planes]
co = case["coeffs"].copy()
cd = abi.coeffs_desc(co.ctypes.data, co.size)
t0 = time.perf_counter()
oracle.vvco_intra_recon_frame(abi.frame_from_numpy(geom, want), case["dec_blks"].ctypes.data, case["dec_blk_end"].ctypes.data, C.byref(cd),
                              case["dec_tbs"].ctypes.data, case["dec_tb_end"].ctypes.data, len(case["dec_blk_end"]), 15)
t_cpu = time.perf_counter() - t0
ctx = lib.Context(0)
torch.cuda.set_stream(ctx.torch_stream())
keep = [device.to_device(a) for a in (case["blks"], case["coeffs"], case["tbs"])]
dcd = abi.coeffs_desc(keep[1][1], case["coeffs"].size)
fr = device.DeviceFrames(geom, planes=planes)
ctx.intra_recon_frame(fr.desc, keep[0][1], case["blk_end"], dcd, keep[2][1], case["tb_end"], 15)
ctx.sync()
got = fr.to_numpy()
equal = all(np.array_equal(got[c][:, :, :geom.plane_wh(c)[0]], want[c][:, :, :geom.plane_wh(c)[0]]) for c in range(3))
times = []
for _ in range(5):
    fr2 = device.DeviceFrames(geom, planes=planes)
    ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ctx.intra_recon_frame(fr2.desc, keep[0][1], case["blk_end"], dcd, keep[2][1], case["tb_end"], 15)
    b.record()
    ctx.sync()
    times.append(a.elapsed_time(b))
ms = float(np.median(times))
# the same pictures in ONE launch, dependencies resolved on the device; the steps in three legal orders
one = {}
for how in ("decode", "ctu_wavefront", "block_wave"):
    arrs = synth.intra_step_order(case, how)
    keep2 = [device.to_device(a) for a in arrs]
    n_steps = len(arrs[1])
    args = (keep2[0][1], keep2[1][1], dcd, keep2[2][1], keep2[3][1], n_steps, len(arrs[0]), len(arrs[2]), 15)
    fr3 = device.DeviceFrames(geom, planes=planes)
    ctx.intra_recon_frame_ordered(fr3.desc, *args)
    got3 = fr3.to_numpy()
    equal3 = all(np.array_equal(got3[c][:, :, :geom.plane_wh(c)[0]], want[c][:, :, :geom.plane_wh(c)[0]]) for c in range(3))
    times3 = []
    for _ in range(5):
        fr4 = device.DeviceFrames(geom, planes=planes)
        ctx.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ctx.intra_recon_frame_ordered(fr4.desc, *args)
        b.record()
        ctx.sync()
        times3.append(a.elapsed_time(b))
    ms3 = float(np.median(times3))
    one[how] = {"steps": n_steps, "ms_per_ring": ms3, "ms_per_picture": ms3 / batch, "mpix_per_s": w * h * batch / (ms3 * 1e-3) / 1e6,
                "parity_equal": bool(equal3)}
print(json.dumps({"workload": "all-intra reconstruction (intra_pred / MIP / CCLM + residual)", "width": w, "height": h, "pictures": batch,
                  "blocks": int(len(case["blks"])), "transform_blocks": int(len(case["tbs"])), "waves": int(case["n_waves"]),
                  "launch_per_wave": {"ms_per_ring": ms, "ms_per_picture": ms / batch, "mpix_per_s": w * h * batch / (ms * 1e-3) / 1e6,
                                      "us_per_wave": ms * 1e3 / case["n_waves"], "parity_equal": bool(equal)},
                  "one_launch": one,
                  "cpu_oracle_one_thread_mpix_per_s": w * h * batch / t_cpu / 1e6, "against": "oracle", "synth_s": t_synth}))
