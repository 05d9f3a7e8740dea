"""CPU-only: the C-ABI library loads and exports every symbol include/vvcdsp_cuda.h declares."""
import ctypes as C

from ffvvc_b200 import abi, lib


def test_library_exports_every_declared_symbol():
    handle = lib.load()
    names = lib.declared_symbols()
    assert "vvc_cuda_ctx_create" in names and "vvc_cuda_alf_frame" in names
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing


def test_descriptor_sizes_match_the_header():
    # sizes the CUDA translation units were compiled with (exported for this check)
    handle = lib.load()
    handle.vvc_cuda_abi_sizeof.restype = C.c_size_t
    handle.vvc_cuda_abi_sizeof.argtypes = [C.c_int]
    want = {0: C.sizeof(abi.VVCCudaFrame), 1: C.sizeof(abi.VVCCudaALFCtb), 2: C.sizeof(abi.VVCCudaALFSets),
            3: C.sizeof(abi.VVCCudaDbkEdge), 4: C.sizeof(abi.VVCCudaDeblockMaps), 5: C.sizeof(abi.VVCCudaSAOCtb), 6: C.sizeof(abi.VVCCudaInloopDesc), 7: C.sizeof(abi.VVCCudaTB),
            8: abi.PB_DTYPE.itemsize, 9: abi.WP_DTYPE.itemsize, 10: abi.PROF_DTYPE.itemsize, 11: abi.DMVR_OUT_DTYPE.itemsize,
            12: C.sizeof(abi.VVCCudaRect), 13: C.sizeof(abi.VVCCudaReconDesc),
            14: abi.INTRA_PB_DTYPE.itemsize, 15: abi.CIIP_DTYPE.itemsize,
            16: abi.TB_QUANT_DTYPE.itemsize, 17: abi.SCALING_LIST_DTYPE.itemsize, 18: C.sizeof(abi.VVCCudaCoeffs),
            19: abi.LMCS_VPDU_DTYPE.itemsize, 20: abi.LMCS_PARAMS_DTYPE.itemsize, 21: abi.INTRA_BLK_DTYPE.itemsize, 22: abi.DBK_TU_DTYPE.itemsize,
            23: abi.DBK_MVF_DTYPE.itemsize, 24: abi.DBK_CTB_DTYPE.itemsize, 25: C.sizeof(abi.VVCCudaDbkParams), 26: C.sizeof(abi.VVCCudaDbkSide)}
    for which, size in want.items():
        assert handle.vvc_cuda_abi_sizeof(which) == size, which


def test_ctx_create_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(lib.VVCCudaError):
        lib.Context(0)


def test_window_layout_packing_round_trip():
    """abi.pack_window16 (what a host would do while writing levels): every value inside a TB's window survives,
    offsets are dense, nothing outside the windows is stored."""
    import numpy as np
    from ffvvc_b200 import synth
    geom = abi.FrameGeom(256, 128)
    tbs, coeffs = synth.tb_list(geom, seed=3, extras=False)
    tbs = synth.tb_for_window(tbs)
    wt, win = abi.pack_window16(tbs, coeffs)
    sizes = wt["nzw"].astype(np.int64) * wt["nzh"].astype(np.int64)
    assert np.array_equal(wt["coeff_offset"], np.concatenate([[0], np.cumsum(sizes)[:-1]]))
    for i in (0, 1, len(tbs) // 2, len(tbs) - 1):
        t, u = tbs[i], wt[i]
        w = 1 << int(t["log2_w"])
        block = coeffs[int(t["coeff_offset"]):int(t["coeff_offset"]) + w * (1 << int(t["log2_h"]))].reshape(-1, w)
        got = win[int(u["coeff_offset"]):int(u["coeff_offset"]) + int(u["nzw"]) * int(u["nzh"])].reshape(int(u["nzh"]), int(u["nzw"]))
        assert np.array_equal(got, block[:int(u["nzh"]), :int(u["nzw"])])
        assert not block[int(u["nzh"]):, :].any() and not block[:, int(u["nzw"]):].any()
