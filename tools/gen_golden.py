#!/usr/bin/env python
"""Generate tests/golden/stages.json from the COMPILED REFERENCE (oracle/_ref/libvvcref.so).

The reference's own tests store no function-level vectors for this path (checkasm compares C against
asm on random input), so the vectors are produced by running the unmodified reference C -
its table entries driven in the reference drivers' order by oracle/refbuild/ref_glue_*.c - over the
seeded cases of tests/golden_cases.py.  Needs /root/reference (make -C oracle ref).  Run:
    python tools/gen_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import golden_cases as gc, util  # noqa: E402


def main():
    ref = util.ref()
    be = gc.HostBackend(ref, "vvcref_")
    out = {}
    for name in gc.CASES:
        case = gc.build_case(name, util.oracle().vvco_lfnst_tr_set)
        res = gc.run_case(case, be)
        out[name] = {st: {"sha256": gc.digest(arrs), "crop": gc.crop(arrs)} for st, arrs in res.items()}
        print(name, {k: v["sha256"][:12] for k, v in out[name].items()})
    path = os.path.join(ROOT, "tests", "golden", "stages.json")
    with open(path, "w") as f:
        json.dump({"generator": "tools/gen_golden.py", "source": "oracle/_ref/libvvcref.so (unmodified reference C)", "cases": out}, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
