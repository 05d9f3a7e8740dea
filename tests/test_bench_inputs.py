"""CPU-only: the synthetic workload bench.py feeds to every arm is self-consistent - the 16-bit window layout it
uploads in the end-to-end number holds the same quantised levels as the dense int32 layout the device-resident number
and the reference arm read, and its quantisation records are legal inputs of dequant()."""
import ctypes as C

import numpy as np

import bench
from ffvvc_b200 import abi, synth
from tests import util


def test_bench_inputs_layouts_agree_and_records_are_legal():
    inp = bench.Inputs(416, 240, seed=12345, distinct=2, lfnst_set_of=util.oracle().vvco_lfnst_tr_set)
    g = inp.g1
    for i in range(inp.distinct):
        tbs, lv, wt, win, q = inp.tbs[i], inp.coeffs[i], inp.win_tbs[i], inp.win[i], inp.quant[i]
        assert len(tbs) == len(wt) == len(q)
        assert np.abs(lv).max() < 1 << 15 and win.dtype == np.int16
        # records: qp inside the reference's rem6 / div6 tables even with the dependent-quantisation offset, matrix ids
        # from Table 38 for the block's size class, none on transform-skip blocks
        assert int(q["qp"].max()) + 1 <= 75 and int(q["sl_id"].max()) <= 28
        ts = (tbs["flags"] & abi.TB_TS) != 0
        assert not q["sl_id"][ts].any()
        size_idx = np.maximum(tbs["log2_w"], tbs["log2_h"]).astype(np.int64) - 1
        used = q["sl_id"] != 0
        legal = (synth.SL_IDS[:, tbs["c_idx"].astype(np.int64)[used], size_idx[used]] + 1 == q["sl_id"][used]).any(axis=0)
        assert legal.all()
        # the two layouts reconstruct the same picture through the oracle
        pred = synth.uniform_planes(g, seed=3 + i)
        a, b = [p.copy() for p in pred], [p.copy() for p in pred]
        dense = abi.coeffs_desc(lv.ctypes.data, lv.size, abi.COEFF_DENSE32, q.ctypes.data, inp.scaling.ctypes.data)
        window = abi.coeffs_desc(win.ctypes.data, win.size, abi.COEFF_WINDOW16, q.ctypes.data, inp.scaling.ctypes.data)
        util.oracle().vvco_itx_frame_q(abi.frame_from_numpy(g, a), C.byref(dense), tbs.ctypes.data, len(tbs), 15)
        util.oracle().vvco_itx_frame_q(abi.frame_from_numpy(g, b), C.byref(window), wt.ctypes.data, len(wt), 15)
        util.assert_planes_equal(g, a, b, "bench inputs: dense vs window layout")
        assert any(not np.array_equal(x, y) for x, y in zip(a, pred)), "the residual must change the picture"
