"""CPU: the oracle's intra leaf predictors and CIIP blend vs the compiled, unmodified reference table entries
(intra.pred_planar / pred_dc / pred_v / pred_h / pred_angular_v / pred_angular_h / pred_mip, inter.put_ciip)."""
import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util


def run_intra(fn, geom, pbs, edges, fill=5):
    planes = abi.alloc_planes(geom, fill=fill)
    fn(abi.frame_from_numpy(geom, planes), pbs.ctypes.data, len(pbs), edges.ctypes.data)
    return planes


def run_ciip(fn, geom, intra, inter, blocks):
    planes = [p.copy() for p in intra]
    fn(abi.frame_from_numpy(geom, planes), abi.frame_from_numpy(geom, inter), blocks.ctypes.data, len(blocks))
    return planes


@pytest.mark.parametrize("w,h,seed,bd", [(416, 240, 1, 10), (256, 128, 2, 10), (832, 480, 3, 10), (192, 128, 4, 12)])
def test_intra_leaf_oracle_equals_reference(w, h, seed, bd):
    geom = abi.FrameGeom(w, h, bit_depth=bd)
    pbs, edges = synth.intra_list(geom, seed=seed)
    assert len(set(pbs["kind"].tolist())) == 7
    a = run_intra(util.oracle().vvco_intra_leaf_frame, geom, pbs, edges)
    b = run_intra(util.ref().vvcref_intra_leaf_frame, geom, pbs, edges)
    util.assert_planes_equal(geom, a, b, "oracle vs reference")


def test_inverse_angle_is_an_integer_expression():
    """The CUDA kernel replaces the reference's float (32 * 512.0 / angle, rounded) by (32768 + a) / (2a)."""
    for a in synth.INTRA_ANGLES[1:]:
        f = np.float32(32 * 512.0 / int(a))
        assert int(float(f) + 0.5) == (32768 + int(a)) // (2 * int(a))


@pytest.mark.parametrize("w,h,seed", [(416, 240, 5), (256, 128, 6)])
def test_ciip_oracle_equals_reference(w, h, seed):
    geom = abi.FrameGeom(w, h)
    intra, inter = synth.uniform_planes(geom, seed=seed), synth.uniform_planes(geom, seed=seed + 50)
    blocks = synth.ciip_list(geom, seed=seed)
    a = run_ciip(util.oracle().vvco_ciip_frame, geom, intra, inter, blocks)
    b = run_ciip(util.ref().vvcref_ciip_frame, geom, intra, inter, blocks)
    util.assert_planes_equal(geom, a, b, "oracle vs reference")
