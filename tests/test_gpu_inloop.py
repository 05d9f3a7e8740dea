"""GPU parity for the whole in-loop chain (BASELINE.json config 2): deblock -> SAO -> ALF."""
import ctypes as C

import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    from ffvvc_b200 import lib
    c = lib.Context(0)                          # library-owned stream
    with torch.cuda.stream(c.torch_stream()):   # torch uploads/downloads are ordered on the same stream
        yield c
    c.close()


def oracle_chain(geom, planes, maps, sao, alf, sets):
    o = util.oracle()
    md = abi.deblock_maps_desc(geom, maps)
    a, b = abi.alloc_planes(geom), abi.alloc_planes(geom)
    fa, fb = abi.frame_from_numpy(geom, a), abi.frame_from_numpy(geom, b)
    o.vvco_deblock_frame(fa, abi.frame_from_numpy(geom, planes), C.byref(md), 1)
    o.vvco_deblock_frame(fb, fa, C.byref(md), 0)
    o.vvco_sao_frame(fa, fb, sao.ctypes.data)
    o.vvco_alf_frame(fb, fa, alf.ctypes.data, sets.ctypes.data, 0)
    return b


def make_inputs(geom, seed):
    planes = synth.struct_planes(geom, seed=seed)
    maps = synth.deblock_maps(geom, seed=seed + 1, qp_base=27, qp_span=16)
    sao = synth.sao_params(geom, seed=seed + 2)
    alf, sets = synth.alf_params(geom, seed=seed + 3)
    return planes, maps, sao, alf, sets


@pytest.mark.parametrize("w,h,batch", [(416, 240, 1), (1920, 1080, 1), (384, 256, 3)])
def test_inloop_chain_device(ctx, w, h, batch):
    from ffvvc_b200 import device
    from tests.test_gpu_lf_sao import upload_maps
    geom = abi.FrameGeom(w, h, batch=batch)
    planes, maps, sao, alf, sets = make_inputs(geom, seed=w)
    src, dst = device.DeviceFrames(geom, planes=planes), device.DeviceFrames(geom)
    md, keep = upload_maps(geom, maps)
    t1, p1 = device.to_device(sao)
    t2, p2 = device.to_device(alf)
    t3, p3 = device.to_device(sets)
    ctx.inloop_frame(dst.desc, src.desc, abi.inloop_desc(md, p1, p2, p3))
    ctx.sync()
    util.assert_planes_equal(geom, dst.to_numpy(), oracle_chain(geom, planes, maps, sao, alf, sets), "in-loop chain")


def test_inloop_chain_host_entry_pipelined(ctx):
    geom = abi.FrameGeom(416, 240, batch=4)
    planes, maps, sao, alf, sets = make_inputs(geom, seed=77)
    out = abi.alloc_planes(geom)
    md = abi.deblock_maps_desc(geom, maps)
    desc = abi.inloop_desc(md, sao.ctypes.data, alf.ctypes.data, sets.ctypes.data)
    ctx.inloop_frame_host(abi.frame_from_numpy(geom, out), abi.frame_from_numpy(geom, planes), desc)
    util.assert_planes_equal(geom, out, oracle_chain(geom, planes, maps, sao, alf, sets), "host in-loop chain")
