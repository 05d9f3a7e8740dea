"""Oracle pinning of the fused dequantisation + compact coefficient layout:
the oracle's dequant restatement against the reference's own static dequant() (vvc_intra.c:397-417, reached
through oracle/refbuild/ref_glue_dequant.c), per block over the parameter space and at stage level."""
import ctypes as C

import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util


def _levels(rng, n, kind):
    raw = rng.take(n).astype(np.int64)
    if kind == 0:                                        # realistic: small levels
        return ((raw >> 4) % 257 - 128).astype(np.int32)
    wide = ((raw << 8) ^ (raw >> 3)) & 0xFFFFFFFF        # full 16-bit range: exercises the clip and int wrap
    wide = np.where(wide >= 1 << 31, wide - (1 << 32), wide)
    return np.clip(wide, -32768, 32767).astype(np.int32)


@pytest.mark.parametrize("bit_depth", [10, 12])
def test_dequant_block_vs_reference(bit_depth):
    ref, ora = util.ref(), util.oracle()
    rng = synth.LCG(2024 + bit_depth)
    sl = np.zeros(1, dtype=abi.SCALING_LIST_DTYPE)
    sl["matrix_rec"][0] = (1 + rng.below(28 * 64, 255)).reshape(28, 64)
    sl["dc_rec"][0] = 1 + rng.below(14, 255)
    qp_bd = 6 * (bit_depth - 8)
    n_checked = 0
    for l2w in range(0, 7):
        for l2h in range(0, 7):
            if l2w == 0 and l2h == 0:
                continue
            w, h = 1 << l2w, 1 << l2h
            for rep in range(12):
                c_idx = int(rng.below(1, 3)[0])
                ts = int(rng.below(1, 4)[0] == 0) if max(w, h) <= 32 else 0
                dep = int(rng.below(1, 2)[0])
                explicit = int(rng.below(1, 3)[0] > 0)
                is_intra = int(rng.below(1, 2)[0])
                apply_lfnst = int(rng.below(1, 4)[0] == 0)
                lfnst_dis = int(rng.below(1, 2)[0])
                # qp + dep_quant must stay inside the reference's rem6/div6 tables (76 entries)
                tb_qp = int(rng.below(1, 75 - 4)[0]) + 4
                cu_qp = tb_qp - qp_bd if c_idx == 0 else tb_qp
                if cu_qp < -128 or tb_qp > 63 + qp_bd:
                    continue
                max_x, max_y = int(rng.below(1, w)[0]), int(rng.below(1, h)[0])
                lev = _levels(rng, w * h, rep % 2).reshape(h, w)
                lev[max_y + 1:, :] = 0
                lev[:, max_x + 1:] = 0
                a = np.ascontiguousarray(lev.copy())
                got_qp = ref.vvcref_dequant_tb(a.ctypes.data, l2w, l2h, c_idx, ts, 0, 0, max_x, max_y,
                                               cu_qp, is_intra, 0, apply_lfnst, 0, 0, bit_depth, 15, 0,
                                               dep, explicit, lfnst_dis, sl.ctypes.data)
                assert got_qp == tb_qp
                # the record the host would build: Table 38 id unless derive_scale_m falls back to the flat matrix
                q = np.zeros(1, dtype=abi.TB_QUANT_DTYPE)
                q["qp"], q["dep_quant"] = got_qp, dep
                flat = (not explicit) or ts or (lfnst_dis and apply_lfnst)
                size_idx = max(l2w, l2h) - 1
                q["sl_id"] = 0 if flat else synth.SL_IDS[0 if is_intra else 1, c_idx, size_idx] + 1
                b = np.ascontiguousarray(lev.copy())
                ora.vvco_dequant_tb(b.ctypes.data, l2w, l2h, ts, q.ctypes.data, sl.ctypes.data, 15, bit_depth)
                assert np.array_equal(a, b), (l2w, l2h, c_idx, ts, dep, explicit, is_intra, tb_qp, int(q["sl_id"][0]))
                n_checked += 1
    assert n_checked > 400


@pytest.mark.parametrize("w,h,seed,scaling", [(416, 240, 1, True), (256, 128, 2, False), (832, 480, 3, True)])
def test_quantised_stage_vs_reference(w, h, seed, scaling):
    """Whole residual stage from quantised levels: oracle (window layout) vs the reference (dense layout through
    transform_bdpcm -> dequant() -> itx -> add_residual), and the two layouts against each other."""
    geom = abi.FrameGeom(w, h)
    tbs, levels = synth.tb_list(geom, seed=seed, lfnst_set_of=util.oracle().vvco_lfnst_tr_set, extras=False)
    tbs = synth.tb_for_window(tbs)
    quant, sl = synth.tb_quant(tbs, seed=seed + 50, scaling=scaling)
    wt, win = abi.pack_window16(tbs, levels)
    pred = synth.uniform_planes(geom, seed=seed + 10)
    slp = sl.ctypes.data if sl is not None else None

    def run(fn, layout):
        planes = [p.copy() for p in pred]
        if layout == "dense":
            lv = levels.copy()
            co = abi.coeffs_desc(lv.ctypes.data, lv.size, abi.COEFF_DENSE32, quant.ctypes.data, slp)
            fn(abi.frame_from_numpy(geom, planes), C.byref(co), tbs.ctypes.data, len(tbs), 15)
        else:
            co = abi.coeffs_desc(win.ctypes.data, win.size, abi.COEFF_WINDOW16, quant.ctypes.data, slp)
            fn(abi.frame_from_numpy(geom, planes), C.byref(co), wt.ctypes.data, len(wt), 15)
        return planes

    ref = util.ref()
    ref.vvcref_itx_frame_q.argtypes = util.oracle().vvco_itx_frame_q.argtypes
    ref.vvcref_itx_frame_q.restype = None
    want = run(ref.vvcref_itx_frame_q, "dense")
    util.assert_planes_equal(geom, run(util.oracle().vvco_itx_frame_q, "dense"), want, "oracle dense vs reference")
    util.assert_planes_equal(geom, run(util.oracle().vvco_itx_frame_q, "window"), want, "oracle window vs reference")
    util.assert_planes_equal(geom, run(ref.vvcref_itx_frame_q, "window"), want, "reference glue window vs dense")


def test_window_layout_is_lossless_for_dequantised_input():
    """Without quant records the window layout must give what the dense int32 entry gives."""
    geom = abi.FrameGeom(416, 240)
    tbs, coeffs = synth.tb_list(geom, seed=7, lfnst_set_of=util.oracle().vvco_lfnst_tr_set, extras=False)
    tbs = synth.tb_for_window(tbs)
    wt, win = abi.pack_window16(tbs, coeffs)
    assert win.nbytes * 3 < coeffs.nbytes
    pred = synth.uniform_planes(geom, seed=17)
    a = [p.copy() for p in pred]
    co = coeffs.copy()
    util.oracle().vvco_itx_frame(abi.frame_from_numpy(geom, a), co.ctypes.data, tbs.ctypes.data, len(tbs), 15)
    b = [p.copy() for p in pred]
    cd = abi.coeffs_desc(win.ctypes.data, win.size, abi.COEFF_WINDOW16)
    util.oracle().vvco_itx_frame_q(abi.frame_from_numpy(geom, b), C.byref(cd), wt.ctypes.data, len(wt), 15)
    util.assert_planes_equal(geom, a, b, "window vs dense")
