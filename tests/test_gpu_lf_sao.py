"""GPU parity: deblocking (both passes) and SAO through the C ABI vs the CPU oracle, bit exact."""
import ctypes as C

import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util

pytestmark = pytest.mark.gpu

GEOMS = [(256, 128, 7), (200, 136, 7), (416, 240, 7), (176, 144, 6), (128, 64, 5), (1920, 1080, 7)]


@pytest.fixture(scope="module")
def ctx():
    import torch
    from ffvvc_b200 import lib
    assert torch.cuda.is_available()
    c = lib.Context(0)                          # library-owned stream
    with torch.cuda.stream(c.torch_stream()):   # torch uploads/downloads are ordered on the same stream
        yield c
    c.close()


def upload_maps(geom, maps):
    from ffvvc_b200 import device
    keep, ptrs = [], {}
    for d in range(2):
        for c in range(3):
            t, ptr = device.to_device(maps[d][c])
            keep.append(t)
            ptrs[id(maps[d][c])] = ptr
    desc = abi.deblock_maps_desc(geom, maps, ptr_of=lambda a: ptrs[id(a)])
    return desc, keep


def cuda_deblock(ctx, geom, planes, maps):
    from ffvvc_b200 import device
    src = device.DeviceFrames(geom, planes=planes)
    tmp = device.DeviceFrames(geom)
    out = device.DeviceFrames(geom)
    desc, keep = upload_maps(geom, maps)
    ctx.deblock_frame(tmp.desc, src.desc, desc, 1)
    ctx.deblock_frame(out.desc, tmp.desc, desc, 0)
    ctx.sync()
    return tmp.to_numpy(), out.to_numpy()


def oracle_deblock(geom, planes, maps):
    desc = abi.deblock_maps_desc(geom, maps)
    tmp, out = abi.alloc_planes(geom), abi.alloc_planes(geom)
    fn = util.oracle().vvco_deblock_frame
    fn(abi.frame_from_numpy(geom, tmp), abi.frame_from_numpy(geom, planes), C.byref(desc), 1)
    fn(abi.frame_from_numpy(geom, out), abi.frame_from_numpy(geom, tmp), C.byref(desc), 0)
    return tmp, out


@pytest.mark.parametrize("w,h,ctb_log2", GEOMS)
def test_deblock_bit_exact(ctx, w, h, ctb_log2):
    geom = abi.FrameGeom(w, h, ctb_log2=ctb_log2)
    planes = synth.struct_planes(geom, seed=w + h)
    maps = synth.deblock_maps(geom, seed=w * 3 + h, qp_base=27, qp_span=16)
    gv, gh = cuda_deblock(ctx, geom, planes, maps)
    ov, oh = oracle_deblock(geom, planes, maps)
    util.assert_planes_equal(geom, gv, ov, "vertical pass cuda vs oracle")
    util.assert_planes_equal(geom, gh, oh, "horizontal pass cuda vs oracle")


def test_deblock_noise_12bit_and_ring(ctx):
    geom = abi.FrameGeom(320, 192, bit_depth=12, batch=2)
    planes = synth.uniform_planes(geom, seed=12)
    maps = synth.deblock_maps(geom, seed=12, qp_base=40, qp_span=24)
    gv, gh = cuda_deblock(ctx, geom, planes, maps)
    ov, oh = oracle_deblock(geom, planes, maps)
    util.assert_planes_equal(geom, gv, ov, "vertical")
    util.assert_planes_equal(geom, gh, oh, "horizontal")


def test_deblock_host_entry(ctx):
    geom = abi.FrameGeom(416, 240)
    planes = synth.struct_planes(geom, seed=3)
    maps = synth.deblock_maps(geom, seed=4, qp_base=30, qp_span=10)
    out = abi.alloc_planes(geom)
    desc = abi.deblock_maps_desc(geom, maps)
    ctx.deblock_frame_host(abi.frame_from_numpy(geom, out), abi.frame_from_numpy(geom, planes), desc)
    util.assert_planes_equal(geom, out, oracle_deblock(geom, planes, maps)[1], "host entry vs oracle")


def test_deblock_4k_all_off_is_identity(ctx):
    geom = abi.FrameGeom(3840, 2160)
    planes = synth.uniform_planes(geom, seed=1)
    maps = synth.deblock_maps(geom, seed=2)
    for d in range(2):
        for c in range(3):
            maps[d][c]["tc"] = 0
    gv, gh = cuda_deblock(ctx, geom, planes, maps)
    util.assert_planes_equal(geom, gh, planes, "tc == 0 everywhere must be the identity")


def cuda_sao(ctx, geom, planes, params):
    from ffvvc_b200 import device
    src = device.DeviceFrames(geom, planes=planes)
    dst = device.DeviceFrames(geom)
    t, ptr = device.to_device(params)
    ctx.sao_frame(dst.desc, src.desc, ptr)
    ctx.sync()
    return dst.to_numpy()


def oracle_sao(geom, planes, params):
    out = abi.alloc_planes(geom)
    util.oracle().vvco_sao_frame(abi.frame_from_numpy(geom, out), abi.frame_from_numpy(geom, planes), params.ctypes.data)
    return out


@pytest.mark.parametrize("w,h,ctb_log2", GEOMS)
@pytest.mark.parametrize("restore", [False, True])
def test_sao_bit_exact(ctx, w, h, ctb_log2, restore):
    geom = abi.FrameGeom(w, h, ctb_log2=ctb_log2)
    planes = synth.uniform_planes(geom, seed=w ^ h) if (w & 16) else synth.struct_planes(geom, seed=w ^ h)
    params = synth.sao_params(geom, seed=h, with_restore=restore)
    util.assert_planes_equal(geom, cuda_sao(ctx, geom, planes, params), oracle_sao(geom, planes, params), "SAO cuda vs oracle")


def test_sao_ring_and_host_entry(ctx):
    geom = abi.FrameGeom(384, 256, batch=3)
    planes = synth.uniform_planes(geom, seed=8)
    params = synth.sao_params(geom, seed=9, with_restore=True)
    want = oracle_sao(geom, planes, params)
    util.assert_planes_equal(geom, cuda_sao(ctx, geom, planes, params), want, "ring")
    out = abi.alloc_planes(geom)
    ctx.sao_frame_host(abi.frame_from_numpy(geom, out), abi.frame_from_numpy(geom, planes), params.ctypes.data)
    util.assert_planes_equal(geom, out, want, "host entry")


def test_sao_4k_band_only_is_pointwise(ctx):
    """4K property: with every CTB in band mode the result is a pure per-sample LUT (checked with numpy)."""
    geom = abi.FrameGeom(3840, 2160)
    planes = synth.uniform_planes(geom, seed=5)
    params = synth.sao_params(geom, seed=6)
    params["type_idx"][:] = 1
    params["band_position"][:] = 7
    params["offset_val"][:, :, 1:] = np.array([3, -2, 5, -7])
    got = cuda_sao(ctx, geom, planes, params)
    for c in range(3):
        v = planes[c].astype(np.int64)
        band = ((v >> 5) - 7) & 31
        off = np.select([band == 0, band == 1, band == 2, band == 3], [3, -2, 5, -7], 0)
        want = np.clip(v + off, 0, 1023).astype(np.uint16)
        w = geom.plane_wh(c)[0]
        assert np.array_equal(got[c][:, :, :w], want[:, :, :w])
