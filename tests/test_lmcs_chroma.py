"""LMCS chroma residual scaling (intra.lmcs_scale_chroma, libavcodec/vvc/vvc_intra_template.c:377-448): the scaling
arithmetic inside the residual stage and the per-VPDU derivation of the scale - oracle vs the compiled reference
(CPU) and CUDA vs oracle (GPU)."""
import ctypes as C

import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util


def case_for(w, h, batch, seed, bd=10, ctb_log2=7, literal=False):
    geom = abi.FrameGeom(w, h, batch=batch, bit_depth=bd, ctb_log2=ctb_log2)
    tbs, coeffs = synth.tb_list(geom, seed=seed, lfnst_set_of=util.oracle().vvco_lfnst_tr_set, extras=False)
    tbs, order, vpdus, params, n_luma = synth.lmcs_chroma(geom, tbs, seed=seed + 1, literal=literal)
    pred = synth.uniform_planes(geom, seed=seed + 10)
    return geom, tbs, coeffs, vpdus, params, n_luma, pred


@pytest.mark.parametrize("w,h,batch,seed,bd", [(416, 240, 1, 1, 10), (256, 128, 2, 2, 10), (832, 480, 1, 3, 12)])
def test_scaled_residuals_match_reference(w, h, batch, seed, bd):
    """itransform()'s order with chroma_scale (vvc_intra.c:449-475, joint blocks :166-186): the reference's own
    lmcs_scale_chroma between its itx and add_residual entries, literal scales per block."""
    geom, tbs, coeffs, _, _, _, pred = case_for(w, h, batch, seed, bd, literal=True)
    assert (tbs["chroma_scale"] > 0).any() and ((tbs["flags"] & abi.TB_JOINT) != 0)[tbs["chroma_scale"] > 0].any()
    res = []
    for lib, fn in ((util.oracle(), "vvco_itx_frame"), (util.ref(), "vvcref_itx_frame")):
        planes = [p.copy() for p in pred]
        co = coeffs.copy()
        getattr(lib, fn)(abi.frame_from_numpy(geom, planes), co.ctypes.data, tbs.ctypes.data, len(tbs), 15)
        res.append(planes)
    util.assert_planes_equal(geom, res[0], res[1], "oracle vs reference, scaled chroma residuals")
    plain = tbs.copy()
    plain["chroma_scale"] = 0
    planes = [p.copy() for p in pred]
    co = coeffs.copy()
    util.oracle().vvco_itx_frame(abi.frame_from_numpy(geom, planes), co.ctypes.data, plain.ctypes.data, len(plain), 15)
    assert not np.array_equal(planes[1], res[0][1]), "the scale changed nothing"
    assert np.array_equal(planes[0], res[0][0])


@pytest.mark.parametrize("w,h,batch,bd,ctb_log2", [(416, 240, 2, 10, 7), (200, 136, 1, 10, 6), (256, 128, 1, 12, 5), (1920, 1080, 1, 10, 7)])
def test_derivation_matches_reference(w, h, batch, bd, ctb_log2):
    """lmcs_derive_chroma_scale through the reference's table entry and its own availability functions
    (ff_vvc_get_left_available / _top_available), VPDUs at picture borders, CTB borders and inside CTBs."""
    geom = abi.FrameGeom(w, h, batch=batch, bit_depth=bd, ctb_log2=ctb_log2)
    planes = synth.struct_planes(geom, seed=w + bd)
    _, _, vpdus, params, _ = synth.lmcs_chroma(geom, np.zeros(0, dtype=abi.TB_DTYPE), seed=h)
    a, b = np.zeros(len(vpdus), np.uint16), np.zeros(len(vpdus), np.uint16)
    f = abi.frame_from_numpy(geom, planes)
    util.oracle().vvco_lmcs_chroma_scale(f, vpdus.ctypes.data, len(vpdus), params.ctypes.data, a.ctypes.data)
    util.ref().vvcref_lmcs_chroma_scale(f, vpdus.ctypes.data, len(vpdus), params.ctypes.data, b.ctypes.data)
    assert np.array_equal(a, b), np.nonzero(a != b)[0][:8]
    assert len(np.unique(a)) > 2 and (vpdus["avail_l"] == 0).any() and (vpdus["avail_t"] != 0).any()


def oracle_stage(geom, tbs, coeffs, vpdus, params, n_luma, pred, fmt=abi.COEFF_DENSE32, quant=None, sl=None):
    """luma residuals -> scales -> chroma residuals, the order of the whole-picture entry"""
    o = util.oracle()
    planes = [p.copy() for p in pred]
    f = abi.frame_from_numpy(geom, planes)
    co = coeffs.copy()
    q = lambda lo: (quant[lo:].ctypes.data if quant is not None else None)
    cd = abi.coeffs_desc(co.ctypes.data, co.size, fmt, q(0), sl.ctypes.data if sl is not None else None)
    o.vvco_itx_frame_q(f, C.byref(cd), tbs.ctypes.data, n_luma, 15)
    scales = np.zeros(len(vpdus), np.uint16)
    o.vvco_lmcs_chroma_scale(f, vpdus.ctypes.data, len(vpdus), params.ctypes.data, scales.ctypes.data)
    cd = abi.coeffs_desc(co.ctypes.data, co.size, fmt, q(n_luma), sl.ctypes.data if sl is not None else None, scales.ctypes.data)
    o.vvco_itx_frame_q(f, C.byref(cd), tbs[n_luma:].ctypes.data, len(tbs) - n_luma, 15)
    return planes, scales


@pytest.fixture(scope="module")
def ctx():
    import torch
    from ffvvc_b200 import lib
    c = lib.Context(0)
    with torch.cuda.stream(c.torch_stream()):
        yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,batch,seed,bd", [(416, 240, 1, 1, 10), (256, 128, 2, 2, 10), (832, 480, 1, 3, 12), (1920, 1080, 1, 4, 10)])
def test_cuda_scaled_residuals_bit_exact(ctx, w, h, batch, seed, bd):
    """literal scales; 10-bit goes through the warp / tiny / leftover kernels, 12-bit and the generic option through itx_kernel"""
    from ffvvc_b200 import device
    geom, tbs, coeffs, _, _, _, pred = case_for(w, h, batch, seed, bd, literal=True)
    want = [p.copy() for p in pred]
    co = coeffs.copy()
    util.oracle().vvco_itx_frame(abi.frame_from_numpy(geom, want), co.ctypes.data, tbs.ctypes.data, len(tbs), 15)
    for generic in (0, 1):
        ctx.set_option(abi.OPT_GENERIC_KERNELS, generic)
        fr = device.DeviceFrames(geom, planes=pred)
        t1, p1 = device.to_device(coeffs)
        t2, p2 = device.to_device(tbs)
        ctx.itx_frame(fr.desc, p1, p2, len(tbs), 15)
        ctx.sync()
        util.assert_planes_equal(geom, fr.to_numpy(), want, "cuda vs oracle, generic=%d" % generic)
    ctx.set_option(abi.OPT_GENERIC_KERNELS, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,batch,seed,bd,ctb_log2,mode", [(416, 240, 2, 5, 10, 7, "dense"), (200, 136, 1, 6, 10, 6, "window_q"),
                                                             (256, 128, 1, 7, 12, 5, "dense"), (1920, 1080, 1, 8, 10, 7, "window_q")])
def test_cuda_derivation_and_indexed_scales_bit_exact(ctx, w, h, batch, seed, bd, ctb_log2, mode):
    """the three calls of the whole-picture entry: luma blocks, vvc_cuda_lmcs_chroma_scale, chroma blocks reading the
    per-VPDU scales by index"""
    from ffvvc_b200 import device
    geom, tbs, coeffs, vpdus, params, n_luma, pred = case_for(w, h, batch, seed, bd, ctb_log2)
    fmt, quant, sl = abi.COEFF_DENSE32, None, None
    if mode == "window_q":
        tbs, coeffs = abi.pack_window16(synth.tb_for_window(tbs), coeffs)
        quant, sl = synth.tb_quant(tbs, seed=seed + 7, scaling=True)
        fmt = abi.COEFF_WINDOW16
    want, want_scales = oracle_stage(geom, tbs, coeffs, vpdus, params, n_luma, pred, fmt, quant, sl)
    fr = device.DeviceFrames(geom, planes=pred)
    keep = [device.to_device(a) for a in (coeffs, tbs, vpdus, params, np.zeros(len(vpdus), np.uint16))]
    (_, pc), (_, pt), (_, pv), (_, pp), (ts, ps) = keep
    pq = psl = None
    if quant is not None:
        keep += [device.to_device(quant), device.to_device(sl)]
        pq, psl = keep[-2][1], keep[-1][1]
    esz = abi.TB_DTYPE.itemsize
    ctx.itx_frame_q(fr.desc, abi.coeffs_desc(pc, coeffs.size, fmt, pq, psl), pt, n_luma, 15)
    ctx.lmcs_chroma_scale(fr.desc, pv, len(vpdus), pp, ps)
    ctx.itx_frame_q(fr.desc, abi.coeffs_desc(pc, coeffs.size, fmt, pq + 4 * n_luma if pq else None, psl, ps), pt + esz * n_luma, len(tbs) - n_luma, 15)
    ctx.sync()
    assert np.array_equal(ts.cpu().numpy().view(np.uint16)[:len(vpdus)], want_scales)
    util.assert_planes_equal(geom, fr.to_numpy(), want, "cuda vs oracle")
