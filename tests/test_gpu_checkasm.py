"""The reference's OWN checkasm harness against the CUDA-backed tables (north_star: "bit-exact against the reference C
templates and the checkasm harness").

oracle/_ref/checkasm is tests/checkasm/{checkasm,vvc_alf,vvc_itx,vvc_mc,vvc_sao}.c and libavcodec/vvc/vvcdsp.c of the
reference, compiled UNMODIFIED where they lie by `make -C oracle checkasm` (needs /root/reference, so it is built in the
development container by __graft_entry__.build() and travels to the GPU box like the other prebuilt files).  Its arch hook
ff_vvc_dsp_init_x86() is oracle/refbuild/chk_glue.c, which installs ff_vvc_dsp_init_cuda()'s entries for the pass checkasm
labels "avx2"; checkasm then compares every such entry with the C entry on its own random inputs (bit depths 10 and 12)."""
import os
import re
import subprocess

import pytest

from tests import util

pytestmark = pytest.mark.gpu
CHECKASM = os.path.join(util.ROOT, "oracle", "_ref", "checkasm")


@pytest.mark.parametrize("name,at_least", [("vvc_alf", 3000), ("vvc_sao", 30), ("vvc_mc", 400), ("vvc_itx", 300)])
def test_reference_checkasm_passes_on_the_cuda_tables(name, at_least):
    if not os.path.exists(CHECKASM):
        pytest.skip("oracle/_ref/checkasm not built (needs /root/reference: make -C oracle checkasm)")
    for seed in ("1", "20261019"):
        p = subprocess.run([CHECKASM, "--test=" + name, seed], capture_output=True, text=True, timeout=1500)
        out = p.stdout + p.stderr
        assert p.returncode == 0, out[-3000:]
        assert "CUDA table error" not in out, out[-3000:]
        m = re.search(r"all (\d+) tests passed", out)
        assert m, out[-2000:]
        assert int(m.group(1)) >= at_least, "only %s functions were checked" % m.group(1)
