"""Pin the oracle's residual stage (LFNST, DCT-II/DST-VII/DCT-VIII, BDPCM, add_residual, joint CbCr)
against the compiled reference (CPU only)."""
import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util


def both(geom, tbs, coeffs, pred, rng_range=15):
    res = []
    for lib, fn in ((util.oracle(), "vvco_itx_frame"), (util.ref(), "vvcref_itx_frame")):
        planes = [p.copy() for p in pred]
        co = coeffs.copy()
        getattr(lib, fn)(abi.frame_from_numpy(geom, planes), co.ctypes.data, tbs.ctypes.data, len(tbs), rng_range)
        res.append((planes, co))
    return res


@pytest.mark.parametrize("w,h,batch,seed", [(416, 240, 1, 1), (256, 128, 2, 2), (832, 480, 1, 3), (1920, 1080, 1, 4)])
def test_residual_stage_matches_reference(w, h, batch, seed):
    geom = abi.FrameGeom(w, h, batch=batch)
    tbs, coeffs = synth.tb_list(geom, seed=seed, lfnst_set_of=util.oracle().vvco_lfnst_tr_set)
    pred = synth.uniform_planes(geom, seed=seed + 10)
    (po, co), (pr, cr) = both(geom, tbs, coeffs, pred)
    util.assert_planes_equal(geom, po, pr, "oracle vs reference")
    assert np.array_equal(co, cr), "stored residuals differ"
    # every feature was exercised
    assert (tbs["lfnst"] != 0).any() and (tbs["flags"] & abi.TB_JOINT).any() and (tbs["flags"] & abi.TB_BDPCM).any()
    assert set(np.unique(tbs["trh"])) == {0, 1, 2} and set(np.unique(tbs["log2_w"])) >= {1, 2, 3, 4, 5, 6}


def test_every_table_cell_like_checkasm():
    """All (trh, trv, w, h) cells of itx.itx[][][][] the reference fills (tests/checkasm/vvc_itx.c:41-98),
    several nz windows each, with stale data outside the nz window as checkasm leaves it."""
    rng = synth.LCG(5)
    recs = []
    sizes = [1, 2, 4, 8, 16, 32, 64]
    for trh in range(3):
        for trv in range(3):
            for lw, w in enumerate(sizes):
                for lh, h in enumerate(sizes):
                    if w == 1 and h == 1:
                        continue
                    if (w == 1 or h == 1) and max(w, h) < 16:
                        continue
                    if trh and not (4 <= w <= 32 or (w == 1 and False)):
                        continue
                    if trv and not (4 <= h <= 32):
                        continue
                    if (w == 1 and trh) or (h == 1 and trv):
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
    coeffs = np.clip(((raw << 9) & 0xFFFFFFFF) - (1 << 31), -32768, 32767).astype(np.int32)   # nothing zeroed: stale tail
    geom = abi.FrameGeom(64, 64)
    pred = abi.alloc_planes(geom)
    (po, co), (pr, cr) = both(geom, tbs, coeffs, pred)
    assert np.array_equal(co, cr)
    assert len(tbs) > 600
