"""Host-side logic of the all-intra reconstruction (CPU only): the wave assignment of synth.intra_picture must make the
picture independent of the order of the blocks inside a wave, and equal to decoding order - checked on the oracle."""
import ctypes as C

import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util


def run(geom, case, blks, blk_end, tbs, tb_end):
    planes = abi.alloc_planes(geom, fill=512)
    co = case["coeffs"].copy()
    cd = abi.coeffs_desc(co.ctypes.data, co.size)
    util.oracle().vvco_intra_recon_frame(abi.frame_from_numpy(geom, planes), blks.ctypes.data, blk_end.ctypes.data, C.byref(cd),
                                         tbs.ctypes.data, tb_end.ctypes.data, len(blk_end), 15)
    return planes


@pytest.mark.parametrize("w,h,batch,seed", [(416, 240, 2, 3), (832, 480, 1, 5), (1280, 720, 1, 11)])
def test_wave_order_equals_decoding_order_and_is_order_free_inside_a_wave(w, h, batch, seed):
    geom = abi.FrameGeom(w, h, batch=batch)
    case = synth.intra_picture(geom, seed=seed)
    dec = run(geom, case, case["dec_blks"], case["dec_blk_end"], case["dec_tbs"], case["dec_tb_end"])
    wav = run(geom, case, case["blks"], case["blk_end"], case["tbs"], case["tb_end"])
    util.assert_planes_equal(geom, wav, dec, "waves vs decoding order")
    blks, tbs, s, t = case["blks"].copy(), case["tbs"].copy(), 0, 0
    for e, f in zip(case["blk_end"], case["tb_end"]):
        blks[s:e] = blks[s:e][::-1]
        tbs[t:f] = tbs[t:f][::-1]
        s, t = e, f
    rev = run(geom, case, blks, case["blk_end"], tbs, case["tb_end"])
    util.assert_planes_equal(geom, rev, dec, "blocks of every wave in reverse order vs decoding order")
    assert case["n_waves"] > 30 and (case["blks"]["kind"] == 2).any() and (case["blks"]["kind"] == 1).any() and (case["blks"]["ref_idx"] > 0).any()


@pytest.mark.parametrize("how", ["ctu_wavefront", "block_wave"])
def test_other_step_orders_are_legal(how):
    """synth.intra_step_order: the steps by CTU anti-diagonal / by block wave are a permutation of the decoding order that
    keeps every step after the steps it reads - the oracle, walking them one after another, gives the same pictures"""
    geom = abi.FrameGeom(416, 240, batch=2)
    case = synth.intra_picture(geom, seed=8)
    dec = run(geom, case, case["dec_blks"], case["dec_blk_end"], case["dec_tbs"], case["dec_tb_end"])
    blks, blk_end, tbs, tb_end = synth.intra_step_order(case, how)
    assert len(blks) == len(case["dec_blks"]) and len(tbs) == len(case["dec_tbs"]) and not np.array_equal(blks, case["dec_blks"])
    assert sorted(map(bytes, blks)) == sorted(map(bytes, case["dec_blks"]))
    util.assert_planes_equal(geom, run(geom, case, blks, blk_end, tbs, tb_end), dec, how + " order vs decoding order")
    same = synth.intra_step_order(case, "decode")
    assert all(np.array_equal(a, b) for a, b in zip(same, (case["dec_blks"], case["dec_blk_end"], case["dec_tbs"], case["dec_tb_end"])))
