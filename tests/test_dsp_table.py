"""The drop-in table override ff_vvc_dsp_init_cuda (include/vvcdsp_table.h).

CPU part: the hook exists, installs exactly the documented entries for 10-bit and nothing for other depths,
and the table layout the library was compiled with has the size of the ctypes mirror (and of the compiled
reference's table when oracle/_ref is built).  GPU part: the installed entries, called through the
reference's own signatures with host pointers, equal the oracle / numpy restatement and - when the compiled
reference is present - the reference's C table on the same inputs."""
import ctypes as C

import numpy as np
import pytest

from ffvvc_b200 import abi, lib, synth
from ffvvc_b200.dsp_tables import VVCDSPContext
from tests import util


def ptr(fn):
    return C.cast(fn, C.c_void_p).value


def fresh_table(bit_depth=10):
    t = VVCDSPContext()
    lib.load().ff_vvc_dsp_init_cuda(C.byref(t), bit_depth)
    return t


def valid_cell(trh, trv, lw, lh):
    """The cells the reference installs (libavcodec/vvc/vvcdsp.c:140-195, vvcdsp_template.c:142-159)."""
    if lw == 0 and lh == 0:
        return False
    if lw == 0 or lh == 0:
        tr, other, l = (trh, trv, lw) if lw else (trv, trh, lh)
        return other == 0 and l >= 4 and (tr == 0 or l <= 5)
    return (trh == 0 or 2 <= lw <= 5) and (trv == 0 or 2 <= lh <= 5)


def test_hook_installs_the_documented_entries():
    handle = lib.load()
    assert handle.ff_vvc_dsp_cuda_sizeof_table() == C.sizeof(VVCDSPContext)
    t = fresh_table(10)
    for trh in range(3):
        for trv in range(3):
            for lw in range(7):
                for lh in range(7):
                    assert bool(ptr(t.itx.itx[trh][trv][lw][lh])) == valid_cell(trh, trv, lw, lh), (trh, trv, lw, lh)
    for depth in (10, 12):
        t = fresh_table(depth)
        assert ptr(t.itx.transform_bdpcm) and ptr(t.itx.add_residual) and ptr(t.lmcs.filter)
        assert ptr(t.itx.add_residual_joint) and ptr(t.itx.pred_residual_joint)
        for ch in range(2):
            for i in range(7):
                for a in range(2):
                    for b in range(2):
                        assert ptr(t.inter.put[ch][i][a][b]) and ptr(t.inter.put_uni[ch][i][a][b]) and ptr(t.inter.put_uni_w[ch][i][a][b])
        for name in ("avg", "w_avg", "put_ciip", "put_gpm", "fetch_samples", "bdof_fetch_samples", "prof_grad_filter", "apply_prof",
                     "apply_prof_uni", "apply_prof_uni_w", "apply_bdof", "sad"):
            assert ptr(getattr(t.inter, name)), name
        assert all(ptr(t.inter.dmvr[a][b]) for a in range(2) for b in range(2))
        assert all(ptr(t.sao.band_filter[i]) and ptr(t.sao.edge_filter[i]) for i in range(9))
        assert ptr(t.alf.filter[0]) and ptr(t.alf.filter[1]) and ptr(t.alf.filter_cc) and ptr(t.alf.classify) and ptr(t.alf.recon_coeff_and_clip)
        assert all(ptr(t.lf.filter_luma[d]) and ptr(t.lf.filter_chroma[d]) and ptr(t.lf.ladf_level[d]) for d in range(2))
        # entries that take the decoder's VVCLocalContext, and the SAO restore pass (takes SAOParams), keep what the caller installed
        assert not ptr(t.intra.intra_pred) and not ptr(t.intra.intra_cclm_pred) and not ptr(t.intra.lmcs_scale_chroma)   # these take VVCLocalContext*
        assert not ptr(t.sao.edge_restore[0])
    t = fresh_table(8)                      # 8-bit pictures (pixel = uint8_t): nothing is installed
    assert not ptr(t.lmcs.filter) and not ptr(t.itx.itx[0][0][2][2]) and not ptr(t.inter.avg)


def test_reference_cells_are_the_cells_we_install():
    if not util.have_ref():
        pytest.skip("compiled reference not built")
    ref = util.ref_dsp(10)
    for trh in range(3):
        for trv in range(3):
            for lw in range(7):
                for lh in range(7):
                    assert bool(ptr(ref.itx.itx[trh][trv][lw][lh])) == valid_cell(trh, trv, lw, lh), (trh, trv, lw, lh)


def _coeffs(rng, n):
    raw = rng.take(n).astype(np.int64)
    return np.clip(((raw << 9) & 0xFFFFFFFF) - (1 << 31), -32768, 32767).astype(np.int32)


@pytest.mark.gpu
def test_itx_cells_through_the_table():
    t = fresh_table(10)
    ref = util.ref_dsp(10) if util.have_ref() else None
    rng = synth.LCG(77)
    ip = C.POINTER(C.c_int)
    checked = 0
    for trh in range(3):
        for trv in range(3):
            for lw in range(7):
                for lh in range(7):
                    if not valid_cell(trh, trv, lw, lh) or (trh * 3 + trv + lw + lh) % 2:
                        continue
                    w, h = 1 << lw, 1 << lh
                    nzw = int(rng.below(1, min(32 if trh == 0 else 16, w))[0]) + 1
                    nzh = int(rng.below(1, min(32 if trv == 0 else 16, h))[0]) + 1
                    src = _coeffs(rng, w * h)
                    got = src.copy()
                    t.itx.itx[trh][trv][lw][lh](got.ctypes.data_as(ip), nzw, nzh, 15, 10)
                    tb = np.zeros(1, dtype=abi.TB_DTYPE)
                    tb["log2_w"], tb["log2_h"], tb["trh"], tb["trv"], tb["nzw"], tb["nzh"] = lw, lh, trh, trv, nzw, nzh
                    tb["flags"] = abi.TB_STORE_RESIDUAL
                    want = src.copy()
                    g = abi.FrameGeom(64, 64)
                    util.oracle().vvco_itx_frame(abi.frame_from_numpy(g, abi.alloc_planes(g)), want.ctypes.data, tb.ctypes.data, 1, 15)
                    assert np.array_equal(got, want), ("oracle", trh, trv, w, h, nzw, nzh)
                    if ref is not None:
                        rr = src.copy()
                        ref.itx.itx[trh][trv][lw][lh](rr.ctypes.data_as(ip), nzw, nzh, 15, 10)
                        assert np.array_equal(got, rr), ("reference", trh, trv, w, h, nzw, nzh)
                    checked += 1
    assert checked > 100
    assert lib.load().ff_vvc_dsp_cuda_last_error() == 0, lib.load().ff_vvc_dsp_cuda_error_string()


@pytest.mark.gpu
def test_bdpcm_add_residual_and_lmcs_through_the_table():
    t = fresh_table(10)
    ref = util.ref_dsp(10) if util.have_ref() else None
    rng = synth.LCG(78)
    ip = C.POINTER(C.c_int)
    for (w, h, vertical) in [(4, 4, 0), (8, 16, 1), (32, 32, 0), (16, 4, 1)]:
        src = _coeffs(rng, w * h)
        got = src.copy()
        t.itx.transform_bdpcm(got.ctypes.data_as(ip), w, h, vertical, 15)
        want = src.astype(np.int64).reshape(h, w).copy()
        if vertical:
            for y in range(1, h):
                want[y] = np.clip(want[y] + want[y - 1], -32768, 32767)
        else:
            for x in range(1, w):
                want[:, x] = np.clip(want[:, x] + want[:, x - 1], -32768, 32767)
        assert np.array_equal(got.reshape(h, w), want), ("bdpcm", w, h, vertical)
        if ref is not None:
            rr = src.copy()
            ref.itx.transform_bdpcm(rr.ctypes.data_as(ip), w, h, vertical, 15)
            assert np.array_equal(got, rr)
    for (w, h) in [(2, 2), (4, 8), (16, 16), (64, 32), (8, 64)]:
        pitch = 80
        pic = (rng.take(pitch * h) & 1023).astype(np.uint16).reshape(h, pitch)
        res = (rng.take(w * h).astype(np.int64) % 4097 - 2048).astype(np.int32)
        got = pic.copy()
        t.itx.add_residual(got.ctypes.data, res.ctypes.data_as(ip), w, h, pitch * 2)
        want = pic.copy()
        want[:, :w] = np.clip(pic[:, :w].astype(np.int64) + res.reshape(h, w), 0, 1023)
        assert np.array_equal(got, want), ("add_residual", w, h)
        if ref is not None:
            rr = pic.copy()
            ref.itx.add_residual(rr.ctypes.data, res.ctypes.data_as(ip), w, h, pitch * 2)
            assert np.array_equal(got, rr)
    fwd, inv = synth.lmcs_luts(10, seed=3)
    for (w, h) in [(4, 4), (8, 16), (16, 8), (128, 128), (12, 20)]:
        pitch = 136
        pic = (rng.take(pitch * h) & 1023).astype(np.uint16).reshape(h, pitch)
        got = pic.copy()
        t.lmcs.filter(got.ctypes.data, pitch * 2, w, h, inv.ctypes.data)
        want = pic.copy()
        want[:, :w] = inv[pic[:, :w]]
        assert np.array_equal(got, want), ("lmcs", w, h)
        if ref is not None:
            rr = pic.copy()
            ref.lmcs.filter(rr.ctypes.data, pitch * 2, w, h, inv.ctypes.data)
            assert np.array_equal(got, rr)
    assert lib.load().ff_vvc_dsp_cuda_last_error() == 0, lib.load().ff_vvc_dsp_cuda_error_string()
