"""GPU parity: residual stage (LFNST + inverse transforms + BDPCM + add_residual / joint CbCr) vs the oracle."""
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


def cuda_itx(ctx, geom, tbs, coeffs, pred):
    from ffvvc_b200 import device
    fr = device.DeviceFrames(geom, planes=pred)
    t1, p1 = device.to_device(coeffs)
    t2, p2 = device.to_device(tbs)
    ctx.itx_frame(fr.desc, p1, p2, len(tbs), 15)
    ctx.sync()
    return fr.to_numpy(), t1.cpu().numpy().view(np.int32)


def oracle_itx(geom, tbs, coeffs, pred):
    planes = [p.copy() for p in pred]
    co = coeffs.copy()
    util.oracle().vvco_itx_frame(abi.frame_from_numpy(geom, planes), co.ctypes.data, tbs.ctypes.data, len(tbs), 15)
    return planes, co


@pytest.mark.parametrize("w,h,batch,seed", [(416, 240, 1, 1), (256, 128, 2, 2), (832, 480, 1, 3), (1920, 1080, 1, 4)])
def test_residual_stage_bit_exact(ctx, w, h, batch, seed):
    geom = abi.FrameGeom(w, h, batch=batch)
    tbs, coeffs = synth.tb_list(geom, seed=seed, lfnst_set_of=util.oracle().vvco_lfnst_tr_set)
    pred = synth.uniform_planes(geom, seed=seed + 10)
    gp, gc = cuda_itx(ctx, geom, tbs, coeffs, pred)
    op, oc = oracle_itx(geom, tbs, coeffs, pred)
    util.assert_planes_equal(geom, gp, op, "cuda vs oracle")
    assert np.array_equal(gc, oc), "stored residuals differ"


def test_every_table_cell(ctx):
    from tests.test_oracle_vs_ref_itx import test_every_table_cell_like_checkasm as _unused  # noqa: F401
    rng = synth.LCG(5)
    recs = []
    sizes = [1, 2, 4, 8, 16, 32, 64]
    for trh in range(3):
        for trv in range(3):
            for lw, w in enumerate(sizes):
                for lh, h in enumerate(sizes):
                    if (w == 1 and h == 1) or ((w == 1 or h == 1) and max(w, h) < 16):
                        continue
                    if (trh and not 4 <= w <= 32) or (trv and not 4 <= h <= 32):
                        continue
                    for rep in range(4):
                        r = np.zeros(1, dtype=abi.TB_DTYPE)
                        r["log2_w"], r["log2_h"], r["trh"], r["trv"] = lw, lh, trh, trv
                        r["nzw"] = int(rng.below(1, min(32 if trh == 0 else 16, w))[0]) + 1
                        r["nzh"] = int(rng.below(1, min(32 if trv == 0 else 16, h))[0]) + 1
                        r["flags"] = abi.TB_STORE_RESIDUAL
                        recs.append(r)
    tbs = np.concatenate(recs)
    area = (1 << tbs["log2_w"].astype(np.int64)) * (1 << tbs["log2_h"].astype(np.int64))
    off = np.concatenate([[0], np.cumsum(area)])
    tbs["coeff_offset"] = off[:-1]
    raw = rng.take(int(off[-1])).astype(np.int64)
    coeffs = np.clip(((raw << 9) & 0xFFFFFFFF) - (1 << 31), -32768, 32767).astype(np.int32)
    geom = abi.FrameGeom(64, 64)
    pred = abi.alloc_planes(geom)
    gp, gc = cuda_itx(ctx, geom, tbs, coeffs, pred)
    op, oc = oracle_itx(geom, tbs, coeffs, pred)
    bad = np.nonzero(gc != oc)[0]
    if len(bad):
        i = np.searchsorted(tbs["coeff_offset"], bad[0], side="right") - 1
        raise AssertionError("%d residual mismatches, first in TB %s" % (len(bad), tbs[i]))


def test_host_entry_and_4k_linearity(ctx):
    geom = abi.FrameGeom(416, 240)
    tbs, coeffs = synth.tb_list(geom, seed=9, lfnst_set_of=util.oracle().vvco_lfnst_tr_set)
    pred = synth.uniform_planes(geom, seed=19)
    planes = [p.copy() for p in pred]
    co = coeffs.copy()
    ctx.itx_frame_host(abi.frame_from_numpy(geom, planes), co.ctypes.data, co.size, tbs.ctypes.data, len(tbs), 15)
    op, oc = oracle_itx(geom, tbs, coeffs, pred)
    util.assert_planes_equal(geom, planes, op, "host entry vs oracle")
    assert np.array_equal(co, oc)
    # 4K property: zero coefficients leave the prediction untouched; DC-only blocks add a constant
    g4 = abi.FrameGeom(3840, 2160)
    tbs4, co4 = synth.tb_list(g4, seed=11, extras=False)
    pred4 = synth.uniform_planes(g4, seed=12)
    gp, _ = cuda_itx(ctx, g4, tbs4, np.zeros_like(co4), pred4)
    util.assert_planes_equal(g4, gp, pred4, "zero residual must be the identity")


# ---- compact coefficient layout + dequantisation on the device (vvc_cuda_itx_frame_q) ----
def _q_case(geom, seed, scaling=True):
    tbs, levels = synth.tb_list(geom, seed=seed, lfnst_set_of=util.oracle().vvco_lfnst_tr_set, extras=False)
    tbs = synth.tb_for_window(tbs)
    quant, sl = synth.tb_quant(tbs, seed=seed + 50, scaling=scaling)
    wt, win = abi.pack_window16(tbs, levels)
    return tbs, levels, wt, win, quant, sl


def _oracle_q(geom, pred, data, fmt, tbs, quant, sl):
    import ctypes as C
    planes = [p.copy() for p in pred]
    d = data.copy()
    co = abi.coeffs_desc(d.ctypes.data, d.size, fmt, quant.ctypes.data if quant is not None else None,
                         sl.ctypes.data if sl is not None else None)
    util.oracle().vvco_itx_frame_q(abi.frame_from_numpy(geom, planes), C.byref(co), tbs.ctypes.data, len(tbs), 15)
    return planes


def _cuda_q(ctx, geom, pred, data, fmt, tbs, quant, sl):
    from ffvvc_b200 import device
    fr = device.DeviceFrames(geom, planes=pred)
    keep = [device.to_device(data), device.to_device(tbs)]
    qp = sp = None
    if quant is not None:
        keep.append(device.to_device(quant))
        qp = keep[-1][1]
    if sl is not None:
        keep.append(device.to_device(sl))
        sp = keep[-1][1]
    co = abi.coeffs_desc(keep[0][1], data.size, fmt, qp, sp)
    ctx.itx_frame_q(fr.desc, co, keep[1][1], len(tbs), 15)
    ctx.sync()
    return fr.to_numpy()


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("layout", ["dense", "window"])
@pytest.mark.parametrize("with_quant,scaling", [(False, False), (True, False), (True, True)])
def test_compact_layout_and_device_dequant_bit_exact(ctx, layout, with_quant, scaling, generic):
    geom = abi.FrameGeom(832, 480, batch=2)
    tbs, levels, wt, win, quant, sl = _q_case(geom, seed=21 + generic, scaling=scaling)
    pred = synth.uniform_planes(geom, seed=33)
    data, fmt, t = (levels, abi.COEFF_DENSE32, tbs) if layout == "dense" else (win, abi.COEFF_WINDOW16, wt)
    q = quant if with_quant else None
    s = sl if (with_quant and scaling) else None
    ctx.set_option(1, generic)
    try:
        got = _cuda_q(ctx, geom, pred, data, fmt, t, q, s)
    finally:
        ctx.set_option(1, 0)
    util.assert_planes_equal(geom, got, _oracle_q(geom, pred, data, fmt, t, q, s), "cuda vs oracle (%s)" % layout)


def test_quantised_host_entry_and_4k_layout_equivalence(ctx):
    import ctypes as C
    geom = abi.FrameGeom(416, 240)
    tbs, levels, wt, win, quant, sl = _q_case(geom, seed=41)
    pred = synth.uniform_planes(geom, seed=42)
    planes = [p.copy() for p in pred]
    co = abi.coeffs_desc(win.ctypes.data, win.size, abi.COEFF_WINDOW16, quant.ctypes.data, sl.ctypes.data)
    ctx.itx_frame_q_host(abi.frame_from_numpy(geom, planes), co, wt.ctypes.data, len(wt), 15)
    util.assert_planes_equal(geom, planes, _oracle_q(geom, pred, win, abi.COEFF_WINDOW16, wt, quant, sl), "host entry vs oracle")
    # a STORE_RESIDUAL block has nowhere to go in the window layout: refused (errors are sticky, so on a throwaway context)
    bad = wt.copy()
    bad["flags"][0] |= abi.TB_STORE_RESIDUAL
    from ffvvc_b200 import lib
    c2 = lib.Context(0)
    with pytest.raises(Exception):
        c2.itx_frame_q_host(abi.frame_from_numpy(geom, [p.copy() for p in pred]), co, bad.ctypes.data, len(bad), 15)
    c2.close()
    # 4K property: both layouts of the same levels reconstruct the same picture (size-independent check)
    g4 = abi.FrameGeom(3840, 2160)
    tbs4, lv4, wt4, win4, q4, sl4 = _q_case(g4, seed=43)
    pred4 = synth.uniform_planes(g4, seed=44)
    a = _cuda_q(ctx, g4, pred4, lv4, abi.COEFF_DENSE32, tbs4, q4, sl4)
    b = _cuda_q(ctx, g4, pred4, win4, abi.COEFF_WINDOW16, wt4, q4, sl4)
    util.assert_planes_equal(g4, a, b, "4K: window layout vs dense layout")
    assert win4.nbytes * 3 < lv4.nbytes


def test_quantised_entry_argument_errors():
    """Unsupported combinations are refused with VVC_CUDA_ERR_ARG (sticky, so one context per case), an empty list is a no-op."""
    from ffvvc_b200 import device, lib
    geom = abi.FrameGeom(64, 64)
    fr = device.DeviceFrames(geom, planes=synth.uniform_planes(geom, seed=1))
    buf = device.to_device(np.zeros(64, dtype=np.int16))
    tb = device.to_device(np.zeros(1, dtype=abi.TB_DTYPE))
    c = lib.Context(0)
    c.itx_frame_q(fr.desc, abi.coeffs_desc(buf[1], 64, abi.COEFF_WINDOW16), tb[1], 0, 15)      # n_tbs == 0
    c.sync()
    c.close()
    # host entry: a scaling matrix id without a scaling list (or beyond Table 38) is refused before anything is launched
    q = np.zeros(1, dtype=abi.TB_QUANT_DTYPE)
    q["sl_id"] = 3
    planes = synth.uniform_planes(geom, seed=2)
    hb, ht = np.zeros(64, dtype=np.int16), np.zeros(1, dtype=abi.TB_DTYPE)
    ht["log2_w"], ht["log2_h"], ht["nzw"], ht["nzh"] = 2, 2, 4, 4
    c = lib.Context(0)
    with pytest.raises(lib.VVCCudaError):
        c.itx_frame_q_host(abi.frame_from_numpy(geom, planes), abi.coeffs_desc(hb.ctypes.data, 64, abi.COEFF_WINDOW16, q.ctypes.data, None),
                           ht.ctypes.data, 1, 15)
    c.close()
    for fmt, rng in ((abi.COEFF_WINDOW16, 16), (7, 15)):          # 16-bit window needs range 15; unknown layout
        c = lib.Context(0)
        with pytest.raises(lib.VVCCudaError):
            c.itx_frame_q(fr.desc, abi.coeffs_desc(buf[1], 64, fmt), tb[1], 1, rng)
        c.close()
