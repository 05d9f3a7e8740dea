"""GPU parity: deblocking parameters derived on the device (vvc_cuda_deblock_params_frame) vs the oracle - the edge maps
entry by entry, and the pictures the deblocking stage makes of them."""
import ctypes as C

import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util
from tests.test_oracle_vs_ref_deblock_params import derive_and_filter, empty_maps, oracle_params, smooth_planes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    from ffvvc_b200 import lib
    c = lib.Context(0)
    with torch.cuda.stream(c.torch_stream()):
        yield c
    c.close()


@pytest.mark.parametrize("w,h,batch,seed,bd,ctb_log2,ladf,smooth", [(416, 240, 2, 1, 10, 7, True, False), (256, 192, 2, 2, 10, 6, False, True),
                                                                    (832, 480, 1, 3, 10, 7, True, True), (200, 136, 2, 4, 12, 5, True, False),
                                                                    (1920, 1080, 1, 5, 10, 7, True, False), (3840, 2160, 1, 7, 10, 7, True, True)])
def test_parameters_and_filtered_pictures_bit_exact(ctx, w, h, batch, seed, bd, ctb_log2, ladf, smooth):
    from ffvvc_b200 import device
    geom = abi.FrameGeom(w, h, batch=batch, bit_depth=bd, ctb_log2=ctb_log2)
    planes = smooth_planes(geom, seed + 30) if smooth else synth.struct_planes(geom, seed=seed + 30)
    tus, mvfs, ctbs, prm = synth.deblock_side_info(geom, seed=seed, ladf=ladf)
    # the oracle's two derivations (V from the picture, H from the V pass's output) and its pictures
    o = util.oracle()
    want_maps = empty_maps(geom)
    md = abi.deblock_maps_desc(geom, want_maps)
    va, want = abi.alloc_planes(geom), abi.alloc_planes(geom)
    oracle_params(abi.frame_from_numpy(geom, planes), tus, mvfs, ctbs, prm, md, 1)
    o.vvco_deblock_frame(abi.frame_from_numpy(geom, va), abi.frame_from_numpy(geom, planes), C.byref(md), 1)
    oracle_params(abi.frame_from_numpy(geom, va), tus, mvfs, ctbs, prm, md, 0)
    o.vvco_deblock_frame(abi.frame_from_numpy(geom, want), abi.frame_from_numpy(geom, va), C.byref(md), 0)
    # the device: same sequence
    src, mid, out = device.DeviceFrames(geom, planes=planes), device.DeviceFrames(geom), device.DeviceFrames(geom)
    keep = [device.to_device(a) for a in (tus, mvfs, ctbs)]
    got_maps = empty_maps(geom)
    dev_maps = [[device.to_device(got_maps[d][c]) for c in range(3)] for d in range(2)]
    dmd = abi.deblock_maps_desc(geom, got_maps, ptr_of=lambda a: next(dev_maps[d][c][1] for d in range(2) for c in range(3) if got_maps[d][c] is a))
    ctx.deblock_params_frame(src.desc, keep[0][1], len(tus), keep[1][1], len(mvfs), keep[2][1], prm, dmd, 1)
    ctx.deblock_frame(mid.desc, src.desc, dmd, 1)
    ctx.deblock_params_frame(mid.desc, keep[0][1], len(tus), keep[1][1], len(mvfs), keep[2][1], prm, dmd, 0)
    ctx.deblock_frame(out.desc, mid.desc, dmd, 0)
    ctx.sync()
    for d in range(2):
        for c in range(3):
            got = dev_maps[d][c][0].cpu().numpy().view(abi.DBK_EDGE_DTYPE).reshape(want_maps[d][c].shape)
            bad = np.argwhere(got != want_maps[d][c])
            assert not len(bad), "map dir %d plane %d differs at %s: cuda %s oracle %s" % (d, c, bad[0], got[tuple(bad[0])], want_maps[d][c][tuple(bad[0])])
    util.assert_planes_equal(geom, out.to_numpy(), want, "derived on the device + deblocked vs oracle")


def test_inloop_chain_with_parameters_derived_on_the_device(ctx):
    """vvc_cuda_inloop_frame with VVCCudaInloopDesc.dbk_side instead of precomputed maps: deblock V / H with the parameters
    derived before each pass, then SAO and ALF - against the oracle's chain"""
    from ffvvc_b200 import device
    geom = abi.FrameGeom(416, 240, batch=2)
    planes = smooth_planes(geom, 77)
    tus, mvfs, ctbs, prm = synth.deblock_side_info(geom, seed=9, ladf=True)
    sao = synth.sao_params(geom, seed=3)
    alf, sets = synth.alf_params(geom, seed=4)
    o = util.oracle()
    deblocked, _ = derive_and_filter(oracle_params, geom, planes, tus, mvfs, ctbs, prm)
    a, want = abi.alloc_planes(geom), abi.alloc_planes(geom)
    o.vvco_sao_frame(abi.frame_from_numpy(geom, a), abi.frame_from_numpy(geom, deblocked), sao.ctypes.data)
    o.vvco_alf_frame(abi.frame_from_numpy(geom, want), abi.frame_from_numpy(geom, a), alf.ctypes.data, sets.ctypes.data, 0)
    src, dst = device.DeviceFrames(geom, planes=planes), device.DeviceFrames(geom)
    keep = [device.to_device(x) for x in (tus, mvfs, ctbs, sao, alf, sets)]
    side = abi.VVCCudaDbkSide()
    side.tus, side.mvfs, side.ctbs, side.params, side.n_tus, side.n_mvfs = keep[0][1], keep[1][1], keep[2][1], C.pointer(prm), len(tus), len(mvfs)
    d = abi.VVCCudaInloopDesc()
    d.sao, d.alf, d.alf_sets, d.dbk_side = keep[3][1], keep[4][1], keep[5][1], C.pointer(side)
    ctx.inloop_frame(dst.desc, src.desc, d)
    ctx.sync()
    util.assert_planes_equal(geom, dst.to_numpy(), want, "in-loop chain, parameters derived on the device")
