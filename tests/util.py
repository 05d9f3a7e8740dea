"""Test-side loaders: the CPU oracle (oracle/liboracle.so), the compiled reference
(oracle/_ref/libvvcref.so) and the product library.  Only tests import this module."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from ffvvc_b200 import abi
from ffvvc_b200.dsp_tables import VVCDSPContext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libvvcref.so")

FP = C.POINTER(abi.VVCCudaFrame)

_oracle = None
_ref = None


def _newer(src_dir, target):
    t = os.path.getmtime(target)
    for dp, _, fs in os.walk(src_dir):
        for f in fs:
            if f.endswith((".c", ".h", ".inc")) and os.path.getmtime(os.path.join(dp, f)) > t:
                return True
    return False


def oracle():
    """liboracle.so, (re)built on demand with the committed Makefile (gcc only)."""
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO) or _newer(os.path.join(ROOT, "oracle", "src"), ORACLE_SO) \
                or _newer(os.path.join(ROOT, "include"), ORACLE_SO):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
        lib = C.CDLL(ORACLE_SO)
        lib.vvco_alf_frame.argtypes = [FP, FP, C.c_void_p, C.c_void_p, C.c_int]
        lib.vvco_alf_frame.restype = None
        MP = C.POINTER(abi.VVCCudaDeblockMaps)
        lib.vvco_deblock_frame.argtypes = [FP, FP, MP, C.c_int]
        lib.vvco_deblock_frame.restype = None
        lib.vvco_sao_frame.argtypes = [FP, FP, C.c_void_p]
        lib.vvco_sao_frame.restype = None
        lib.vvco_deblock_params_frame.argtypes = [FP, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(abi.VVCCudaDbkParams), MP, C.c_int]
        lib.vvco_deblock_params_frame.restype = None
        lib.vvco_itx_frame.argtypes = [FP, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.vvco_itx_frame.restype = None
        lib.vvco_itx_frame_q.argtypes = [FP, C.POINTER(abi.VVCCudaCoeffs), C.c_void_p, C.c_int, C.c_int]
        lib.vvco_itx_frame_q.restype = None
        lib.vvco_dequant_tb.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.vvco_dequant_tb.restype = None
        lib.vvco_lfnst_tr_set.argtypes = [C.c_int]
        lib.vvco_lmcs_frame.argtypes = [FP, C.c_void_p, C.c_void_p]
        lib.vvco_lmcs_frame.restype = None
        lib.vvco_lmcs_rects.argtypes = [FP, C.c_void_p, C.c_void_p, C.c_int]
        lib.vvco_lmcs_rects.restype = None
        lib.vvco_lmcs_chroma_scale.argtypes = [FP, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.vvco_lmcs_chroma_scale.restype = None
        lib.vvco_inter_frame.argtypes = [FP, FP, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.vvco_inter_frame.restype = None
        lib.vvco_intra_leaf_frame.argtypes = [FP, C.c_void_p, C.c_int, C.c_void_p]
        lib.vvco_intra_leaf_frame.restype = None
        lib.vvco_ciip_frame.argtypes = [FP, FP, C.c_void_p, C.c_int]
        lib.vvco_ciip_frame.restype = None
        lib.vvco_intra_pred_frame.argtypes = [FP, C.c_void_p, C.c_int]
        lib.vvco_intra_pred_frame.restype = None
        lib.vvco_intra_recon_frame.argtypes = [FP, C.c_void_p, C.c_void_p, C.POINTER(abi.VVCCudaCoeffs), C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.vvco_intra_recon_frame.restype = None
        _oracle = lib
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    """The compiled, unmodified reference (tests skip when it was not built)."""
    global _ref
    if _ref is None:
        if not have_ref():
            pytest.skip("oracle/_ref/libvvcref.so not built (needs /root/reference; make -C oracle ref)")
        lib = C.CDLL(REF_SO)
        lib.vvcref_dsp.restype = C.POINTER(VVCDSPContext)
        lib.vvcref_dsp.argtypes = [C.c_int]
        lib.vvcref_alf_frame.argtypes = [FP, FP, C.c_void_p, C.c_void_p, C.c_int]
        lib.vvcref_alf_frame.restype = None
        MP = C.POINTER(abi.VVCCudaDeblockMaps)
        lib.vvcref_deblock_frame.argtypes = [FP, FP, MP, C.c_int]
        lib.vvcref_deblock_frame.restype = None
        lib.vvcref_sao_frame.argtypes = [FP, FP, C.c_void_p]
        lib.vvcref_sao_frame.restype = None
        lib.vvcref_deblock_params_filter.argtypes = [FP, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(abi.VVCCudaDbkParams)]
        lib.vvcref_deblock_params_filter.restype = None
        lib.vvcref_itx_frame.argtypes = [FP, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.vvcref_itx_frame.restype = None
        lib.vvcref_dequant_tb.argtypes = [C.c_void_p] + [C.c_int] * 20 + [C.c_void_p]
        lib.vvcref_dequant_tb.restype = C.c_int
        lib.vvcref_lmcs_frame.argtypes = [FP, C.c_void_p, C.c_void_p]
        lib.vvcref_lmcs_frame.restype = None
        lib.vvcref_lmcs_rects.argtypes = [FP, C.c_void_p, C.c_void_p, C.c_int]
        lib.vvcref_lmcs_rects.restype = None
        lib.vvcref_lmcs_chroma_scale.argtypes = [FP, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.vvcref_lmcs_chroma_scale.restype = None
        lib.vvcref_inter_frame.argtypes = [FP, FP, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.vvcref_inter_frame.restype = None
        lib.vvcref_intra_leaf_frame.argtypes = [FP, C.c_void_p, C.c_int, C.c_void_p]
        lib.vvcref_intra_leaf_frame.restype = None
        lib.vvcref_ciip_frame.argtypes = [FP, FP, C.c_void_p, C.c_int]
        lib.vvcref_ciip_frame.restype = None
        lib.vvcref_intra_pred_frame.argtypes = [FP, C.c_void_p, C.c_int]
        lib.vvcref_intra_pred_frame.restype = None
        _ref = lib
    return _ref


def ref_dsp(bit_depth=10):
    return ref().vvcref_dsp(bit_depth).contents


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def visible(geom, planes):
    """Crop planes to their visible width (drops pitch padding) for comparisons."""
    return [p[:, :, :geom.plane_wh(c)[0]] for c, p in enumerate(planes)]


def assert_planes_equal(geom, a, b, what=""):
    for c, (x, y) in enumerate(zip(visible(geom, a), visible(geom, b))):
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)
            k, r, col = bad[0]
            raise AssertionError("%s plane %d: %d mismatches, first at frame %d (x=%d,y=%d): %d vs %d" % (
                what, c, len(bad), k, col, r, x[k, r, col], y[k, r, col]))
