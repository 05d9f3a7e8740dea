"""Pin the oracle's full intra prediction (reference-line preparation, wide-angle mapping, every predictor, PDPC) and
its CCLM against the compiled reference's own intra.intra_pred / intra.intra_cclm_pred (CPU only)."""
import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util


def both(geom, blks, planes):
    out = []
    for lib, fn in ((util.oracle(), "vvco_intra_pred_frame"), (util.ref(), "vvcref_intra_pred_frame")):
        p = [a.copy() for a in planes]
        getattr(lib, fn)(abi.frame_from_numpy(geom, p), blks.ctypes.data, len(blks))
        out.append(p)
    return out


@pytest.mark.parametrize("w,h,batch,bd,ctb_log2,seed", [(2048, 1536, 6, 10, 7, 1), (1920, 1080, 8, 10, 7, 2), (1024, 768, 16, 12, 6, 3),
                                                        (1280, 720, 16, 10, 5, 4), (2048, 2048, 6, 10, 7, 5)])
def test_intra_pred_and_cclm_match_reference(w, h, batch, bd, ctb_log2, seed):
    geom = abi.FrameGeom(w, h, batch=batch, bit_depth=bd, ctb_log2=ctb_log2)
    planes = synth.uniform_planes(geom, seed=seed + 20)
    blks = synth.intra_blk_list(geom, seed=seed)
    o, r = both(geom, blks, planes)
    for c in range(3):
        bad = np.argwhere(o[c] != r[c])
        if len(bad):
            k, y, x = bad[0]
            sh = 1 if c else 0
            hit = blks[(blks["pic"] == k) & ((blks["c_idx"] == c) | ((blks["kind"] == 2) & (c > 0))) &
                       (blks["x0"] <= x) & (x < blks["x0"] + blks["w"].astype(int)) & (blks["y0"] <= y) & (y < blks["y0"] + blks["h"].astype(int))]
            raise AssertionError("plane %d differs at pic %d (%d, %d): oracle %d reference %d, record %s" % (c, k, x, y, o[c][k, y, x], r[c][k, y, x], hit))
    assert not np.array_equal(o[0], planes[0]) and not np.array_equal(o[1], planes[1])
    # coverage of the record space
    kinds = set(np.unique(blks["kind"]))
    assert kinds == {0, 1, 2}, kinds
    pm = blks["pred_mode"][blks["kind"] == 0]
    assert len(np.unique(pm)) > 45 and (blks["flags"] & abi.INTRA_F_ISP).any() and (blks["ref_idx"] > 0).any()
    assert (blks["avail_left"] == 0).any() and (blks["avail_top"] == 0).any() and ((blks["avail_top"] > 0) & (blks["avail_top"] < blks["w"])).any()
