#!/usr/bin/env python
"""Short in-loop run for ncu captures: a ring of 4K pictures, a few passes (see profiles/README)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffvvc_b200 import abi, device, lib  # noqa: E402
from bench import Inputs  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 6
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
inp = Inputs(3840, 2160, frames)
geom = inp.geom
ctx = lib.Context(0)
torch.cuda.set_stream(ctx.torch_stream())
src, dst = device.DeviceFrames(geom, planes=inp.planes), device.DeviceFrames(geom)
keep, ptrs = [], {}
for d in range(2):
    for c in range(3):
        t, ptrs[id(inp.maps[d][c])] = device.to_device(inp.maps[d][c])
        keep.append(t)
md = abi.deblock_maps_desc(geom, inp.maps, ptr_of=lambda a: ptrs[id(a)])
t1, p1 = device.to_device(inp.sao)
t2, p2 = device.to_device(inp.alf)
t3, p3 = device.to_device(inp.sets)
desc = abi.inloop_desc(md, p1, p2, p3)
for _ in range(passes):
    ctx.inloop_frame(dst.desc, src.desc, desc)
ctx.sync()
print("ok", ctx.launches, "launches")
