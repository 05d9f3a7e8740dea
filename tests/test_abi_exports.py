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
            14: abi.INTRA_PB_DTYPE.itemsize, 15: abi.CIIP_DTYPE.itemsize}
    for which, size in want.items():
        assert handle.vvc_cuda_abi_sizeof(which) == size, which


def test_ctx_create_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(lib.VVCCudaError):
        lib.Context(0)
