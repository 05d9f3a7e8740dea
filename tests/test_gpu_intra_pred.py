"""GPU parity: full intra prediction (reference lines prepared on the device, wide-angle mapping, every predictor, PDPC,
MIP), CCLM, and the all-intra reconstruction by wavefronts - vs the oracle, through the C ABI."""
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
    c = lib.Context(0)
    with torch.cuda.stream(c.torch_stream()):
        yield c
    c.close()


@pytest.mark.parametrize("w,h,batch,bd,ctb_log2,seed", [(2048, 1536, 6, 10, 7, 1), (1920, 1080, 8, 10, 7, 2), (1024, 768, 16, 12, 6, 3),
                                                        (1280, 720, 16, 10, 5, 4), (3840, 2160, 2, 10, 7, 5)])
def test_intra_pred_and_cclm_bit_exact(ctx, w, h, batch, bd, ctb_log2, seed):
    from ffvvc_b200 import device
    geom = abi.FrameGeom(w, h, batch=batch, bit_depth=bd, ctb_log2=ctb_log2)
    planes = synth.uniform_planes(geom, seed=seed + 20)
    blks = synth.intra_blk_list(geom, seed=seed)
    want = [a.copy() for a in planes]
    util.oracle().vvco_intra_pred_frame(abi.frame_from_numpy(geom, want), blks.ctypes.data, len(blks))
    fr = device.DeviceFrames(geom, planes=planes)
    t, p = device.to_device(blks)
    ctx.intra_pred_frame(fr.desc, p, len(blks))
    ctx.sync()
    got = fr.to_numpy()
    for c in range(3):
        bad = np.argwhere(util.visible(geom, got)[c] != util.visible(geom, want)[c]) if hasattr(util, "visible") else []
        if len(bad):
            k, y, x = bad[0]
            hit = blks[(blks["pic"] == k) & ((blks["c_idx"] == c) | ((blks["kind"] == 2) & (c > 0))) &
                       (blks["x0"] <= x) & (x < blks["x0"] + blks["w"].astype(int)) & (blks["y0"] <= y) & (y < blks["y0"] + blks["h"].astype(int))]
            raise AssertionError("plane %d differs at pic %d (%d, %d): cuda %d oracle %d, record %s" % (c, k, x, y, got[c][k, y, x], want[c][k, y, x], hit))
    util.assert_planes_equal(geom, got, want, "cuda vs oracle")
    if w == 1920:
        host = [a.copy() for a in planes]
        ctx.intra_pred_frame_host(abi.frame_from_numpy(geom, host), blks.ctypes.data, len(blks))
        util.assert_planes_equal(geom, host, want, "host entry vs oracle")


@pytest.mark.parametrize("w,h,batch,seed,mode", [(416, 240, 2, 3, "dense"), (256, 192, 3, 4, "window_q"), (832, 480, 1, 5, "dense"), (1920, 1080, 1, 11, "dense")])
def test_all_intra_pictures_by_wavefronts_bit_exact(ctx, w, h, batch, seed, mode):
    """prediction and residual alternating wavefront by wavefront on the GPU vs the oracle walking the same blocks in
    decoding order (luma step, chroma step per coding unit)"""
    from ffvvc_b200 import device
    geom = abi.FrameGeom(w, h, batch=batch)
    case = synth.intra_picture(geom, seed=seed)
    planes = abi.alloc_planes(geom, fill=512)
    fmt, coeffs = abi.COEFF_DENSE32, case["coeffs"]
    quant = dq = sl = None
    tbs, dtbs = case["tbs"], case["dec_tbs"]
    if mode == "window_q":
        # one packing for both orders: pack in decoding order, then carry the new offsets over to the wave order
        dtbs, coeffs = abi.pack_window16(synth.tb_for_window(dtbs), coeffs)
        dq, sl = synth.tb_quant(dtbs, seed=seed + 7, scaling=True)
        key = lambda t: (t["pic"].astype(np.int64) << 40) | (t["c_idx"].astype(np.int64) << 36) | (t["y0"].astype(np.int64) << 16) | t["x0"]
        pos = {int(k): i for i, k in enumerate(key(dtbs))}
        idx = np.array([pos[int(k)] for k in key(tbs)])
        tbs, quant = dtbs[idx], dq[idx]
        fmt = abi.COEFF_WINDOW16
    want = [a.copy() for a in planes]
    co = coeffs.copy()
    cd = abi.coeffs_desc(co.ctypes.data, co.size, fmt, dq.ctypes.data if dq is not None else None, sl.ctypes.data if sl is not None else None)
    util.oracle().vvco_intra_recon_frame(abi.frame_from_numpy(geom, want), case["dec_blks"].ctypes.data, case["dec_blk_end"].ctypes.data,
                                         C.byref(cd), dtbs.ctypes.data, case["dec_tb_end"].ctypes.data, len(case["dec_blk_end"]), 15)
    fr = device.DeviceFrames(geom, planes=planes)
    keep = [device.to_device(a) for a in (case["blks"], coeffs, tbs)]
    pq = psl = None
    if quant is not None:
        keep += [device.to_device(quant), device.to_device(sl)]
        pq, psl = keep[-2][1], keep[-1][1]
    before = ctx.launches
    ctx.intra_recon_frame(fr.desc, keep[0][1], case["blk_end"], abi.coeffs_desc(keep[1][1], coeffs.size, fmt, pq, psl), keep[2][1], case["tb_end"], 15)
    ctx.sync()
    assert ctx.launches - before <= 2 * case["n_waves"]
    util.assert_planes_equal(geom, fr.to_numpy(), want, "wavefronts on the GPU vs decoding order on the CPU")
    assert case["n_waves"] > 20 and (case["blks"]["kind"] == 2).any() and (case["blks"]["kind"] == 1).any()


@pytest.mark.parametrize("w,h,batch,seed,mode", [(416, 240, 2, 3, "dense"), (256, 192, 3, 4, "window_q"), (832, 480, 1, 5, "dense"), (1920, 1080, 2, 11, "dense")])
def test_all_intra_pictures_one_launch_dependencies_on_the_device(ctx, w, h, batch, seed, mode):
    """vvc_cuda_intra_recon_frame_ordered: the decoder's steps in decoding order, ONE launch, every block waiting on the
    progress map for exactly the samples its availability counts name - vs the oracle walking the same steps"""
    from ffvvc_b200 import device
    geom = abi.FrameGeom(w, h, batch=batch)
    case = synth.intra_picture(geom, seed=seed)
    planes = abi.alloc_planes(geom, fill=512)
    fmt, coeffs, dtbs = abi.COEFF_DENSE32, case["coeffs"], case["dec_tbs"]
    dq = sl = None
    if mode == "window_q":
        dtbs, coeffs = abi.pack_window16(synth.tb_for_window(dtbs), coeffs)
        dq, sl = synth.tb_quant(dtbs, seed=seed + 7, scaling=True)
        fmt = abi.COEFF_WINDOW16
    want = [a.copy() for a in planes]
    co = coeffs.copy()
    cd = abi.coeffs_desc(co.ctypes.data, co.size, fmt, dq.ctypes.data if dq is not None else None, sl.ctypes.data if sl is not None else None)
    util.oracle().vvco_intra_recon_frame(abi.frame_from_numpy(geom, want), case["dec_blks"].ctypes.data, case["dec_blk_end"].ctypes.data,
                                         C.byref(cd), dtbs.ctypes.data, case["dec_tb_end"].ctypes.data, len(case["dec_blk_end"]), 15)
    fr = device.DeviceFrames(geom, planes=planes)
    keep = [device.to_device(a) for a in (case["dec_blks"], case["dec_blk_end"], coeffs, dtbs, case["dec_tb_end"])]
    pq = psl = None
    if dq is not None:
        keep += [device.to_device(dq), device.to_device(sl)]
        pq, psl = keep[-2][1], keep[-1][1]
    before = ctx.launches
    ctx.intra_recon_frame_ordered(fr.desc, keep[0][1], keep[1][1], abi.coeffs_desc(keep[2][1], coeffs.size, fmt, pq, psl), keep[3][1], keep[4][1],
                                  len(case["dec_blk_end"]), len(case["dec_blks"]), len(dtbs), 15)
    ctx.sync()
    assert ctx.launches - before == 1
    util.assert_planes_equal(geom, fr.to_numpy(), want, "one launch, dependencies on the device vs decoding order on the CPU")


@pytest.mark.parametrize("how", ["ctu_wavefront", "block_wave"])
@pytest.mark.parametrize("w,h,batch,seed", [(416, 240, 3, 21), (832, 480, 2, 22)])
def test_one_launch_other_legal_step_orders(ctx, w, h, batch, seed, how):
    """the same steps by CTU anti-diagonal (the wavefront-parallel order, pictures interleaved) and by block wave: other legal
    orders give the same pictures as the decoding order on the CPU"""
    from ffvvc_b200 import device
    geom = abi.FrameGeom(w, h, batch=batch)
    case = synth.intra_picture(geom, seed=seed)
    planes = abi.alloc_planes(geom, fill=512)
    want = [a.copy() for a in planes]
    co = case["coeffs"].copy()
    cd = abi.coeffs_desc(co.ctypes.data, co.size)
    util.oracle().vvco_intra_recon_frame(abi.frame_from_numpy(geom, want), case["dec_blks"].ctypes.data, case["dec_blk_end"].ctypes.data,
                                         C.byref(cd), case["dec_tbs"].ctypes.data, case["dec_tb_end"].ctypes.data, len(case["dec_blk_end"]), 15)
    blks, blk_end, tbs, tb_end = synth.intra_step_order(case, how)
    assert not np.array_equal(blks, case["dec_blks"])
    fr = device.DeviceFrames(geom, planes=planes)
    keep = [device.to_device(a) for a in (blks, blk_end, case["coeffs"], tbs, tb_end)]
    ctx.intra_recon_frame_ordered(fr.desc, keep[0][1], keep[1][1], abi.coeffs_desc(keep[2][1], case["coeffs"].size), keep[3][1], keep[4][1],
                                  len(blk_end), len(blks), len(tbs), 15)
    ctx.sync()
    util.assert_planes_equal(geom, fr.to_numpy(), want, "one launch, steps in %s order vs decoding order on the CPU" % how)


def test_ordered_entry_reports_unsatisfiable_dependencies(ctx):
    """a block that calls samples available which no step reconstructs can never be served: the kernel's watchdog ends the
    launch and the entry returns an error instead of hanging"""
    from ffvvc_b200 import device, lib
    geom = abi.FrameGeom(64, 64)
    blks = np.zeros(2, dtype=abi.INTRA_BLK_DTYPE)
    blks["x0"], blks["y0"], blks["w"], blks["h"], blks["pred_mode"] = (0, 32), (0, 0), 8, 8, 1
    blks["cb_w"], blks["cb_h"] = 8, 8
    blks["avail_left"] = (8, 0)              # block 0 at x = 0 claims a left neighbour; units left of x = 0 do not exist -> clamped away
    blks[0]["x0"], blks[0]["avail_left"] = 16, 8        # block 0 at x = 16 waits for x = 12..15, which no step reconstructs
    blk_end, tb_end = np.array([1, 2], np.int32), np.array([0, 0], np.int32)
    c = lib.Context(0)
    try:
        fr = device.DeviceFrames(geom, planes=abi.alloc_planes(geom, fill=512))
        keep = [device.to_device(a) for a in (blks, blk_end, tb_end)]
        with pytest.raises(lib.VVCCudaError):
            c.intra_recon_frame_ordered(fr.desc, keep[0][1], keep[1][1], abi.coeffs_desc(None, 0), None, keep[2][1], 2, 2, 0, 15)
    finally:
        c.close()
