"""GPU parity: intra leaf predictors (planar, DC, V, H, angular with reference-line index / filter / PDPC / wide
angles, MIP) and the CIIP blend vs the oracle, through the C ABI."""
import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util
from tests.test_oracle_vs_ref_intra import run_ciip, run_intra

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    from ffvvc_b200 import lib
    c = lib.Context(0)
    with torch.cuda.stream(c.torch_stream()):
        yield c
    c.close()


@pytest.mark.parametrize("w,h,seed,bd,batch", [(416, 240, 1, 10, 1), (256, 128, 2, 10, 2), (832, 480, 3, 10, 1), (192, 128, 4, 12, 1),
                                               (1920, 1080, 5, 10, 1)])
def test_intra_leaf_bit_exact(ctx, w, h, seed, bd, batch):
    from ffvvc_b200 import device
    geom = abi.FrameGeom(w, h, bit_depth=bd, batch=batch)
    pbs, edges = synth.intra_list(geom, seed=seed)
    fr = device.DeviceFrames(geom, planes=abi.alloc_planes(geom, fill=5))
    t1, p1 = device.to_device(pbs)
    t2, p2 = device.to_device(edges)
    ctx.intra_leaf_frame(fr.desc, p1, len(pbs), p2)
    ctx.sync()
    want = run_intra(util.oracle().vvco_intra_leaf_frame, geom, pbs, edges)
    util.assert_planes_equal(geom, fr.to_numpy(), want, "cuda vs oracle")


def test_intra_host_entry_and_ciip(ctx):
    from ffvvc_b200 import device
    geom = abi.FrameGeom(416, 240)
    pbs, edges = synth.intra_list(geom, seed=11)
    host = abi.alloc_planes(geom, fill=5)
    ctx.intra_leaf_frame_host(abi.frame_from_numpy(geom, host), pbs.ctypes.data, len(pbs), edges.ctypes.data, len(edges))
    want = run_intra(util.oracle().vvco_intra_leaf_frame, geom, pbs, edges)
    util.assert_planes_equal(geom, host, want, "host entry vs oracle")
    inter = synth.uniform_planes(geom, seed=12)
    blocks = synth.ciip_list(geom, seed=13)
    d = device.DeviceFrames(geom, planes=want)
    s = device.DeviceFrames(geom, planes=inter)
    t, p = device.to_device(blocks)
    ctx.ciip_frame(d.desc, s.desc, p, len(blocks))
    ctx.sync()
    blended = run_ciip(util.oracle().vvco_ciip_frame, geom, want, inter, blocks)
    util.assert_planes_equal(geom, d.to_numpy(), blended, "ciip cuda vs oracle")
    h2 = [x.copy() for x in want]
    ctx.ciip_frame_host(abi.frame_from_numpy(geom, h2), abi.frame_from_numpy(geom, inter), blocks.ctypes.data, len(blocks))
    util.assert_planes_equal(geom, h2, blended, "ciip host entry vs oracle")
    # property: equal intra and inter pictures are a fixed point of the blend for every weight
    same = device.DeviceFrames(geom, planes=inter)
    ctx.ciip_frame(same.desc, s.desc, p, len(blocks))
    ctx.sync()
    util.assert_planes_equal(geom, same.to_numpy(), inter, "ciip fixed point")
