"""GPU parity: ALF through the C ABI vs the CPU oracle, bit exact (integer path)."""
import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    from ffvvc_b200 import lib
    assert torch.cuda.is_available()
    c = lib.Context(0)                          # library-owned stream
    with torch.cuda.stream(c.torch_stream()):   # torch uploads/downloads are ordered on the same stream
        yield c
    c.close()


def cuda_alf(ctx, geom, planes, ctbs, sets):
    from ffvvc_b200 import device
    src = device.DeviceFrames(geom, planes=planes)
    dst = device.DeviceFrames(geom)
    dst.t[0].fill_(-1)
    t1, p1 = device.to_device(ctbs)
    t2, p2 = device.to_device(sets)
    ctx.alf_frame(dst.desc, src.desc, p1, p2)
    ctx.sync()
    return dst.to_numpy()


def oracle_alf(geom, planes, ctbs, sets):
    out = abi.alloc_planes(geom)
    util.oracle().vvco_alf_frame(abi.frame_from_numpy(geom, out), abi.frame_from_numpy(geom, planes),
                                 ctbs.ctypes.data, sets.ctypes.data, 0)
    return out


@pytest.mark.parametrize("w,h,ctb_log2", [(256, 128, 7), (200, 136, 7), (416, 240, 7), (176, 144, 6), (128, 64, 5),
                                          (1920, 1080, 7)])
@pytest.mark.parametrize("dist", ["uniform", "struct"])
def test_alf_frame_bit_exact(ctx, w, h, ctb_log2, dist):
    geom = abi.FrameGeom(w, h, ctb_log2=ctb_log2)
    planes = (synth.uniform_planes if dist == "uniform" else synth.struct_planes)(geom, seed=w * 7 + h)
    ctbs, sets = synth.alf_params(geom, seed=h + ctb_log2, all_on=(dist == "uniform"))
    got = cuda_alf(ctx, geom, planes, ctbs, sets)
    want = oracle_alf(geom, planes, ctbs, sets)
    util.assert_planes_equal(geom, got, want, "cuda vs oracle")


@pytest.mark.parametrize("coeffs", ["full", "max"])
@pytest.mark.parametrize("w,h,ctb_log2,bd", [(256, 128, 7, 10), (176, 144, 6, 10), (128, 64, 5, 10), (256, 128, 7, 12),
                                             (1920, 1080, 7, 10)])
def test_alf_full_coefficient_range(ctx, w, h, ctb_log2, bd, coeffs):
    """Coefficients over the whole legal range -128..+128 (cbs_h266_syntax_template.c:2285,2314; +128 does not fit the
    signed byte of the packed path) in every APS slot, AlfCtbFiltSetIdxY 0..23."""
    geom = abi.FrameGeom(w, h, ctb_log2=ctb_log2, bit_depth=bd)
    planes = (synth.uniform_planes if w < 1000 else synth.struct_planes)(geom, seed=w + h)
    ctbs, sets = synth.alf_params(geom, seed=h + ctb_log2, coeffs=coeffs)
    assert sets["luma_coeff"].max() == 128 and sets["chroma_coeff"].max() == 128
    got = cuda_alf(ctx, geom, planes, ctbs, sets)
    want = oracle_alf(geom, planes, ctbs, sets)
    util.assert_planes_equal(geom, got, want, "cuda vs oracle")


@pytest.mark.parametrize("w,h,ctb_log2,bd", [(128, 64, 5, 10), (160, 96, 5, 12), (256, 192, 5, 10), (176, 144, 6, 12),
                                             (384, 256, 7, 10), (384, 256, 7, 12)])
def test_alf_packed_vs_wide_multiply(ctx, w, h, ctb_log2, bd):
    """The packed 16x2 / IDP.2A arithmetic against the 32-bit-multiply path of the same kernel (and the oracle), over
    seeds: every clip index, the virtual-boundary rows, 10 and 12 bit, 16-wide chroma tiles (32x32 CTBs) included."""
    geom = abi.FrameGeom(w, h, ctb_log2=ctb_log2, bit_depth=bd, batch=2)
    for seed in range(6):
        planes = (synth.uniform_planes if seed & 1 else synth.struct_planes)(geom, seed=100 + seed)
        ctbs, sets = synth.alf_params(geom, seed=200 + seed, all_on=seed < 4)
        ctbs["edges"] = synth.LCG(seed).below(len(ctbs), 16) if seed >= 2 else 0
        packed = cuda_alf(ctx, geom, planes, ctbs, sets)
        ctx.set_option(abi.OPT_ALF_WIDE_MULTIPLY, 1)
        try:
            wide = cuda_alf(ctx, geom, planes, ctbs, sets)
        finally:
            ctx.set_option(abi.OPT_ALF_WIDE_MULTIPLY, 0)
        util.assert_planes_equal(geom, packed, wide, "packed vs wide multiply (seed %d)" % seed)
        util.assert_planes_equal(geom, packed, oracle_alf(geom, planes, ctbs, sets), "cuda vs oracle (seed %d)" % seed)


def test_alf_interior_edges_batch_ring(ctx):
    geom = abi.FrameGeom(384, 256, batch=3)
    planes = synth.struct_planes(geom, seed=99)
    ctbs, sets = synth.alf_params(geom, seed=5)
    ctbs["edges"] = synth.LCG(3).below(len(ctbs), 16)
    got = cuda_alf(ctx, geom, planes, ctbs, sets)
    want = oracle_alf(geom, planes, ctbs, sets)
    util.assert_planes_equal(geom, got, want, "cuda vs oracle")


def test_alf_host_entry_matches_device_entry(ctx):
    """The *_host entry (host buffers, copies inside) gives the same picture."""
    geom = abi.FrameGeom(416, 240)
    planes = synth.uniform_planes(geom, seed=4)
    ctbs, sets = synth.alf_params(geom, seed=6)
    out = abi.alloc_planes(geom)
    ctx.alf_frame_host(abi.frame_from_numpy(geom, out), abi.frame_from_numpy(geom, planes),
                       ctbs.ctypes.data, sets.ctypes.data)
    util.assert_planes_equal(geom, out, oracle_alf(geom, planes, ctbs, sets), "host entry vs oracle")


def test_alf_4k_properties(ctx):
    """Full-size 4K: flags off -> identity; result independent of the ring slot (batch invariance)."""
    geom = abi.FrameGeom(3840, 2160, batch=2)
    one = abi.FrameGeom(3840, 2160)
    p1 = synth.struct_planes(one, seed=21)
    planes = [np.concatenate([p, p]) for p in p1]
    ctbs1, sets = synth.alf_params(one, seed=8)
    ctbs = np.concatenate([ctbs1, ctbs1])
    got = cuda_alf(ctx, geom, planes, ctbs, sets)
    for c in range(3):
        assert np.array_equal(got[c][0], got[c][1])
    off = ctbs.copy()
    off["ctb_flag"][:] = 0
    off["cc_idc"][:] = 0
    same = cuda_alf(ctx, geom, planes, off, sets)
    util.assert_planes_equal(geom, same, planes, "all flags off must be the identity")
