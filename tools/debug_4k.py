#!/usr/bin/env python
"""GPU-side triage at the bench's full size: the bench inputs for a seed, stage by stage, CUDA (one picture and a
ring of pictures per launch) against the CPU oracle.  Usage: debug_4k.py [seed] [pictures]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffvvc_b200 import abi, device, lib  # noqa: E402
from bench import Inputs  # noqa: E402
from tests import util  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 12346
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
olib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
olib.vvco_lfnst_tr_set.argtypes = [C.c_int]
inp = Inputs(3840, 2160, seed=seed, distinct=2, lfnst_set_of=olib.vvco_lfnst_tr_set)
g1 = inp.g1
gring = abi.FrameGeom(3840, 2160, batch=frames)
ctx = lib.Context(0)
torch.cuda.set_stream(ctx.torch_stream())
reps = frames // inp.distinct + 1
ring_planes = [np.ascontiguousarray(np.concatenate([p] * reps)[:frames]) for p in inp.ref_planes]
refs = device.DeviceFrames(gring, planes=ring_planes)
keep = []


def up(a):
    t, p = device.to_device(a)
    keep.append(t)
    return p


def diff(name, a, b, geom):
    bad = 0
    for c in range(3):
        wv = geom.plane_wh(c)[0]
        x, y = a[c][:, :, :wv], b[c][:, :, :wv]
        if not np.array_equal(x, y):
            idx = np.argwhere(x != y)
            bad += len(idx)
            print("  %s plane %d: %d mismatches, first (pic %d, y %d, x %d) %d vs %d" % (name, c, len(idx), *idx[0], x[tuple(idx[0])], y[tuple(idx[0])]))
    print("%s: %s" % (name, "OK" if not bad else "%d mismatches" % bad))
    return bad


# ---- inter: ring launch vs per-picture launches vs oracle (picture 0 and 1) ----
pbs_ring = np.concatenate([inp.records(k, frames, k) for k in range(frames)])
cur = device.DeviceFrames(gring, planes=abi.alloc_planes(gring, fill=7))
ctx.inter_frame(cur.desc, refs.desc, up(pbs_ring), len(pbs_ring), up(inp.wp), up(inp.prof), None)
ctx.sync()
ring_out = cur.to_numpy()
single = abi.alloc_planes(gring, fill=7)
for k in range(frames):
    one = device.DeviceFrames(g1, planes=abi.alloc_planes(g1, fill=7))
    pb = inp.records(k, frames, 0)
    ctx.inter_frame(one.desc, refs.desc, up(pb), len(pb), up(inp.wp), up(inp.prof), None)
    ctx.sync()
    o = one.to_numpy()
    for c in range(3):
        single[c][k] = o[c][0]
diff("inter ring vs per-picture", ring_out, single, gring)
ora = abi.alloc_planes(gring, fill=7)
util.oracle().vvco_inter_frame(abi.frame_from_numpy(gring, ora), abi.frame_from_numpy(gring, ring_planes), pbs_ring.ctypes.data,
                               len(pbs_ring), inp.wp.ctypes.data, inp.prof.ctypes.data, None)
if diff("inter ring vs oracle", ring_out, ora, gring):
    # which records
    import collections
    from tools.debug_inter import kind
    bad_kinds, shown = collections.Counter(), 0
    for i, r in enumerate(pbs_ring):
        for c in range(3):
            if not (r["planes"] & (2 if c else 1)):
                continue
            sh = 1 if c else 0
            x0, y0, bw, bh = r["x0"] >> sh, r["y0"] >> sh, r["w"] >> sh, r["h"] >> sh
            a = ring_out[c][r["pic"], y0:y0 + bh, x0:x0 + bw]
            b = ora[c][r["pic"], y0:y0 + bh, x0:x0 + bw]
            if not np.array_equal(a, b):
                bad_kinds[kind(r) + " plane %d" % c] += 1
                if shown < 4:
                    shown += 1
                    print("   rec", i, r, "plane", c)
                    print(a.astype(int)); print(b.astype(int))
    for k_, v_ in sorted(bad_kinds.items()):
        print("   BAD", k_, v_)
# ---- residual on top of the oracle's prediction ----
tb_parts, co_parts, off = [], [], 0
for k in range(frames):
    t = inp.tbs[k % inp.distinct].copy()
    t["pic"] = k
    t["coeff_offset"] += off
    off += len(inp.coeffs[k % inp.distinct])
    tb_parts.append(t)
    co_parts.append(inp.coeffs[k % inp.distinct])
tbs, coeffs = np.concatenate(tb_parts), np.concatenate(co_parts)
pic = device.DeviceFrames(gring, planes=ora)
ctx.itx_frame(pic.desc, up(coeffs), up(tbs), len(tbs), 15)
ctx.sync()
got = pic.to_numpy()
ref_planes = [p.copy() for p in ora]
co = coeffs.copy()
util.oracle().vvco_itx_frame(abi.frame_from_numpy(gring, ref_planes), co.ctypes.data, tbs.ctypes.data, len(tbs), 15)
diff("residual ring vs oracle", got, ref_planes, gring)

# ---- whole chain: device ring entry and host entry vs the oracle's chain, picture by picture ----
def chain_oracle(k):
    o = util.oracle()
    gk = abi.FrameGeom(3840, 2160)
    pb = inp.records(k, frames, 0)
    cur_ = abi.alloc_planes(gk, fill=0)
    o.vvco_inter_frame(abi.frame_from_numpy(gk, cur_), abi.frame_from_numpy(gring, ring_planes), pb.ctypes.data, len(pb),
                       inp.wp.ctypes.data, inp.prof.ctypes.data, None)
    stages = {"inter": [p.copy() for p in cur_]}
    co_ = inp.coeffs[k % inp.distinct].copy()
    tb_ = inp.tbs[k % inp.distinct]
    o.vvco_itx_frame(abi.frame_from_numpy(gk, cur_), co_.ctypes.data, tb_.ctypes.data, len(tb_), 15)
    stages["residual"] = [p.copy() for p in cur_]
    o.vvco_lmcs_frame(abi.frame_from_numpy(gk, cur_), inp.inv_lut.ctypes.data, None)
    stages["lmcs"] = [p.copy() for p in cur_]
    md = abi.deblock_maps_desc(gk, inp.maps[k % inp.distinct])
    a_, b_ = abi.alloc_planes(gk), abi.alloc_planes(gk)
    o.vvco_deblock_frame(abi.frame_from_numpy(gk, a_), abi.frame_from_numpy(gk, cur_), C.byref(md), 1)
    o.vvco_deblock_frame(abi.frame_from_numpy(gk, b_), abi.frame_from_numpy(gk, a_), C.byref(md), 0)
    stages["deblock"] = [p.copy() for p in b_]
    o.vvco_sao_frame(abi.frame_from_numpy(gk, a_), abi.frame_from_numpy(gk, b_), inp.sao[k % inp.distinct].ctypes.data)
    stages["sao"] = [p.copy() for p in a_]
    o.vvco_alf_frame(abi.frame_from_numpy(gk, b_), abi.frame_from_numpy(gk, a_), inp.alf[k % inp.distinct].ctypes.data, inp.sets.ctypes.data, 0)
    stages["alf"] = [p.copy() for p in b_]
    return stages


want = chain_oracle(0)
# device stage by stage on picture 0
gk = abi.FrameGeom(3840, 2160)
d_cur = device.DeviceFrames(gk, planes=abi.alloc_planes(gk, fill=0))
pb0 = inp.records(0, frames, 0)
ctx.inter_frame(d_cur.desc, refs.desc, up(pb0), len(pb0), up(inp.wp), up(inp.prof), None); ctx.sync()
diff("chain/inter", d_cur.to_numpy(), want["inter"], gk)
ctx.itx_frame(d_cur.desc, up(inp.coeffs[0]), up(inp.tbs[0]), len(inp.tbs[0]), 15); ctx.sync()
diff("chain/residual", d_cur.to_numpy(), want["residual"], gk)
ctx.lmcs_frame(d_cur.desc, up(inp.inv_lut), None); ctx.sync()
diff("chain/lmcs", d_cur.to_numpy(), want["lmcs"], gk)
md = abi.VVCCudaDeblockMaps()
for d in range(2):
    for c in range(3):
        rows, pitch = abi.deblock_map_shape(gk, d, c)
        md.edge[d][c] = up(inp.maps[0][d][c])
        md.pitch[d][c], md.rows[d][c], md.size[d][c] = pitch, rows, rows * pitch
ta, tb2 = device.DeviceFrames(gk), device.DeviceFrames(gk)
ctx.deblock_frame(ta.desc, d_cur.desc, md, 1); ctx.deblock_frame(tb2.desc, ta.desc, md, 0); ctx.sync()
diff("chain/deblock", tb2.to_numpy(), want["deblock"], gk)
ctx.sao_frame(ta.desc, tb2.desc, up(inp.sao[0])); ctx.sync()
diff("chain/sao", ta.to_numpy(), want["sao"], gk)
ctx.alf_frame(tb2.desc, ta.desc, up(inp.alf[0]), up(inp.sets), 0); ctx.sync()
diff("chain/alf", tb2.to_numpy(), want["alf"], gk)

# ---- the two composed entries ----
want1 = chain_oracle(1)
d_out, d_cur2 = device.DeviceFrames(gring), device.DeviceFrames(gring)
mdr = abi.VVCCudaDeblockMaps()
for d in range(2):
    for c in range(3):
        rows, pitch = abi.deblock_map_shape(gk, d, c)
        mdr.edge[d][c] = up(np.concatenate([inp.maps[k % inp.distinct][d][c] for k in range(frames)]))
        mdr.pitch[d][c], mdr.rows[d][c], mdr.size[d][c] = pitch, rows, rows * pitch
desc = abi.VVCCudaReconDesc()
desc.pbs, desc.n_pbs, desc.wp, desc.n_wp, desc.prof, desc.n_prof = up(pbs_ring), len(pbs_ring), up(inp.wp), len(inp.wp), up(inp.prof), len(inp.prof)
desc.log2_transform_range = 15
desc.coeffs, desc.n_coeffs, desc.tbs, desc.n_tbs = up(coeffs), len(coeffs), up(tbs), len(tbs)
desc.lmcs_inv_lut = up(inp.inv_lut)
desc.inloop.deblock = C.pointer(mdr)
desc.inloop.sao = up(np.concatenate([inp.sao[k % inp.distinct] for k in range(frames)]))
desc.inloop.alf = up(np.concatenate([inp.alf[k % inp.distinct] for k in range(frames)]))
desc.inloop.alf_sets = up(inp.sets)
ctx.recon_frame(d_out.desc, d_cur2.desc, refs.desc, desc)
ctx.sync()
o = d_out.to_numpy()
diff("recon_frame ring pic0", [p[0:1] for p in o], want["alf"], gk)
diff("recon_frame ring pic1", [p[1:2] for p in o], want1["alf"], gk)
# host entry
h_out = abi.alloc_planes(gring)
descs = (abi.VVCCudaReconDesc * frames)()
keep_h = []
for k in range(frames):
    i = k % inp.distinct
    hmd = abi.deblock_maps_desc(gk, inp.maps[i])
    keep_h.append(hmd)
    pb = inp.records(k, frames, 0)
    keep_h.append(pb)
    dsc = descs[k]
    dsc.pbs, dsc.n_pbs, dsc.wp, dsc.n_wp, dsc.prof, dsc.n_prof = pb.ctypes.data, len(pb), inp.wp.ctypes.data, len(inp.wp), inp.prof.ctypes.data, len(inp.prof)
    dsc.log2_transform_range = 15
    dsc.coeffs, dsc.n_coeffs, dsc.tbs, dsc.n_tbs = inp.coeffs[i].ctypes.data, len(inp.coeffs[i]), inp.tbs[i].ctypes.data, len(inp.tbs[i])
    dsc.lmcs_inv_lut = inp.inv_lut.ctypes.data
    dsc.inloop.deblock = C.pointer(hmd)
    dsc.inloop.sao, dsc.inloop.alf, dsc.inloop.alf_sets = inp.sao[i].ctypes.data, inp.alf[i].ctypes.data, inp.sets.ctypes.data
ctx.recon_frame_host(abi.frame_from_numpy(gring, h_out), abi.frame_from_numpy(gring, ring_planes), descs)
diff("recon_frame_host pic0", [p[0:1] for p in h_out], want["alf"], gk)
diff("recon_frame_host pic1", [p[1:2] for p in h_out], want1["alf"], gk)
diff("recon_frame ring vs recon_frame_host (all pictures)", o, h_out, gring)
# coverage: samples the inter stage never writes (sentinel survives) per plane, on the device and in the oracle
SENT = 1001
cov = device.DeviceFrames(gring, planes=abi.alloc_planes(gring, fill=SENT))
ctx.inter_frame(cov.desc, refs.desc, up(pbs_ring), len(pbs_ring), up(inp.wp), up(inp.prof), None); ctx.sync()
cv = cov.to_numpy()
oc = abi.alloc_planes(gring, fill=SENT)
util.oracle().vvco_inter_frame(abi.frame_from_numpy(gring, oc), abi.frame_from_numpy(gring, ring_planes), pbs_ring.ctypes.data,
                               len(pbs_ring), inp.wp.ctypes.data, inp.prof.ctypes.data, None)
for c in range(3):
    wv = gring.plane_wh(c)[0]
    print("plane", c, "sentinel left: device", int((cv[c][:, :, :wv] == SENT).sum()), "oracle", int((oc[c][:, :, :wv] == SENT).sum()))
# two passes of the residual stage in place must differ from one pass only where inter does not overwrite
# repeated passes in place must give the same pictures
for rep in range(3):
    ctx.recon_frame(d_out.desc, d_cur2.desc, refs.desc, desc)
ctx.sync()
diff("recon_frame ring, 4th pass vs 1st pass", d_out.to_numpy(), o, gring)
# the bench's own sequence of stage calls
tmp_a, tmp_b = device.DeviceFrames(gring), device.DeviceFrames(gring)
out2 = device.DeviceFrames(gring)
for rep in range(2):
    ctx.inter_frame(d_cur2.desc, refs.desc, desc.pbs, desc.n_pbs, desc.wp, desc.prof, None)
    ctx.itx_frame(d_cur2.desc, desc.coeffs, desc.tbs, desc.n_tbs, 15)
    ctx.lmcs_frame(d_cur2.desc, desc.lmcs_inv_lut, None)
    ctx.deblock_frame(tmp_a.desc, d_cur2.desc, mdr, 1)
    ctx.deblock_frame(tmp_b.desc, tmp_a.desc, mdr, 0)
    ctx.sao_frame(tmp_a.desc, tmp_b.desc, desc.inloop.sao)
    ctx.alf_frame(out2.desc, tmp_a.desc, desc.inloop.alf, desc.inloop.alf_sets, 0)
ctx.sync()
diff("stage calls vs recon_frame", out2.to_numpy(), o, gring)
print("oracle Cr(0,0) pic0 per stage:", {k_: int(v_[2][0, 0, 0]) for k_, v_ in want.items()}, "device recon_frame:", int(o[2][0, 0, 0]), "host:", int(h_out[2][0, 0, 0]))
