"""Golden vectors (tests/golden/stages.json, produced by the compiled reference, tools/gen_golden.py):
the oracle must reproduce them on the CPU, the CUDA library through its *_host C-ABI entries on the GPU."""
import json
import os

import numpy as np
import pytest

from tests import golden_cases as gc, util

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "stages.json")))["cases"]


def check(name, res):
    want = GOLDEN[name]
    for stage, arrs in res.items():
        assert gc.crop(arrs) == want[stage]["crop"], "%s/%s: raw crop differs from the reference's" % (name, stage)
        assert gc.digest(arrs) == want[stage]["sha256"], "%s/%s: digest differs from the reference's" % (name, stage)


@pytest.mark.parametrize("name", sorted(gc.CASES))
def test_oracle_reproduces_reference_vectors(name):
    case = gc.build_case(name, util.oracle().vvco_lfnst_tr_set)
    check(name, gc.run_case(case, gc.HostBackend(util.oracle(), "vvco_")))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(gc.CASES))
def test_cuda_reproduces_reference_vectors(name):
    from ffvvc_b200 import lib
    ctx = lib.Context(0)
    try:
        case = gc.build_case(name, util.oracle().vvco_lfnst_tr_set)
        check(name, gc.run_case(case, gc.CudaHostBackend(ctx)))
    finally:
        ctx.close()


def test_golden_file_covers_every_case_and_stage():
    assert set(GOLDEN) == set(gc.CASES)
    for c in GOLDEN.values():
        assert set(c) == {"inter", "intra", "ciip", "residual", "residual_q", "lmcs", "deblock", "sao", "alf"}
