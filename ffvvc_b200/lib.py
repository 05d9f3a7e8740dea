"""Loader and thin Python face of libvvcdsp_cuda.so (the product's C ABI, include/vvcdsp_cuda.h).

There is deliberately no fallback: if the CUDA library is missing or no device is usable,
importing callers get an exception.  PyTorch is used only to own device memory and streams.
"""
import ctypes as C

import numpy as np
import os
import re

from . import abi
from .dsp_tables import VVCDSPContext

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvvcdsp_cuda.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "vvcdsp_cuda.h")
TABLE_HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "vvcdsp_table.h")

FP = C.POINTER(abi.VVCCudaFrame)
CTX = C.c_void_p
NOTIFY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int)

_lib = None


class VVCCudaError(RuntimeError):
    pass


def declared_symbols():
    """Every function name include/vvcdsp_cuda.h and include/vvcdsp_table.h declare (used by the export test)."""
    names = set()
    for path in (HEADER_PATH, TABLE_HEADER_PATH):
        src = open(path).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b((?:vvc_cuda|ff_vvc)_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VVCCudaError("%s not built - run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.vvc_cuda_version.restype = C.c_char_p
    lib.vvc_cuda_ctx_create.argtypes = [C.POINTER(CTX), C.c_int, C.c_void_p]
    lib.vvc_cuda_ctx_destroy.argtypes = [CTX]
    lib.vvc_cuda_ctx_destroy.restype = None
    lib.vvc_cuda_sync.argtypes = [CTX]
    lib.vvc_cuda_last_error.argtypes = [CTX]
    lib.vvc_cuda_error_string.argtypes = [CTX]
    lib.vvc_cuda_error_string.restype = C.c_char_p
    lib.vvc_cuda_stream.argtypes = [CTX]
    lib.vvc_cuda_stream.restype = C.c_void_p
    lib.vvc_cuda_notify.argtypes = [CTX, NOTIFY_FN, C.c_void_p]
    lib.vvc_cuda_launch_count.argtypes = [CTX]
    lib.vvc_cuda_launch_count.restype = C.c_uint64
    for name in ("vvc_cuda_alf_frame", "vvc_cuda_alf_frame_host"):
        fn = getattr(lib, name)
        fn.argtypes = [CTX, FP, FP, C.c_void_p, C.c_void_p, C.c_int]
    MP = C.POINTER(abi.VVCCudaDeblockMaps)
    lib.vvc_cuda_deblock_frame.argtypes = [CTX, FP, FP, MP, C.c_int]
    lib.vvc_cuda_deblock_frame_host.argtypes = [CTX, FP, FP, MP]
    lib.vvc_cuda_deblock_params_frame.argtypes = [CTX, FP, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                  C.POINTER(abi.VVCCudaDbkParams), MP, C.c_int]
    lib.vvc_cuda_sao_frame.argtypes = [CTX, FP, FP, C.c_void_p]
    lib.vvc_cuda_sao_frame_host.argtypes = [CTX, FP, FP, C.c_void_p]
    IP = C.POINTER(abi.VVCCudaInloopDesc)
    lib.vvc_cuda_inloop_frame.argtypes = [CTX, FP, FP, IP]
    lib.vvc_cuda_inloop_frame_host.argtypes = [CTX, FP, FP, IP]
    lib.vvc_cuda_itx_frame.argtypes = [CTX, FP, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.vvc_cuda_itx_frame_host.argtypes = [CTX, FP, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
    CP = C.POINTER(abi.VVCCudaCoeffs)
    lib.vvc_cuda_itx_frame_q.argtypes = [CTX, FP, CP, C.c_void_p, C.c_int, C.c_int]
    lib.vvc_cuda_itx_frame_q_host.argtypes = [CTX, FP, CP, C.c_void_p, C.c_int, C.c_int]
    lib.vvc_cuda_lmcs_frame.argtypes = [CTX, FP, C.c_void_p, C.c_void_p]
    lib.vvc_cuda_lmcs_frame_host.argtypes = [CTX, FP, C.c_void_p, C.c_void_p]
    lib.vvc_cuda_lmcs_rects.argtypes = [CTX, FP, C.c_void_p, C.c_void_p, C.c_int]
    lib.vvc_cuda_lmcs_chroma_scale.argtypes = [CTX, FP, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.vvc_cuda_inter_frame.argtypes = [CTX, FP, FP, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vvc_cuda_inter_frame_host.argtypes = [CTX, FP, FP, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    RP = C.POINTER(abi.VVCCudaReconDesc)
    lib.vvc_cuda_recon_frame.argtypes = [CTX, FP, FP, FP, RP]
    lib.vvc_cuda_recon_frame_host.argtypes = [CTX, FP, FP, RP]
    lib.vvc_cuda_recon_frame_host_async.argtypes = [CTX, FP, FP, RP]
    lib.vvc_cuda_recon_arena_size.argtypes = [FP, RP]
    lib.vvc_cuda_recon_arena_size.restype = C.c_size_t
    lib.vvc_cuda_recon_arena_bind.argtypes = [FP, RP, C.POINTER(abi.VVCCudaDeblockMaps), C.c_void_p]
    lib.vvc_cuda_ctx_set_option.argtypes = [CTX, C.c_int, C.c_int]
    lib.vvc_cuda_pad_frame.argtypes = [CTX, FP, C.c_int]
    lib.vvc_cuda_intra_leaf_frame.argtypes = [CTX, FP, C.c_void_p, C.c_int, C.c_void_p]
    lib.vvc_cuda_intra_leaf_frame_host.argtypes = [CTX, FP, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    lib.vvc_cuda_intra_pred_frame.argtypes = [CTX, FP, C.c_void_p, C.c_int]
    lib.vvc_cuda_intra_pred_frame_host.argtypes = [CTX, FP, C.c_void_p, C.c_int]
    lib.vvc_cuda_intra_recon_frame.argtypes = [CTX, FP, C.c_void_p, C.c_void_p, CP, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.vvc_cuda_intra_recon_frame_ordered.argtypes = [CTX, FP, C.c_void_p, C.c_void_p, CP, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.vvc_cuda_ciip_frame.argtypes = [CTX, FP, FP, C.c_void_p, C.c_int]
    lib.vvc_cuda_ciip_frame_host.argtypes = [CTX, FP, FP, C.c_void_p, C.c_int]
    lib.vvc_cuda_abi_sizeof.argtypes = [C.c_int]
    lib.vvc_cuda_abi_sizeof.restype = C.c_size_t
    lib.ff_vvc_dsp_init_cuda.argtypes = [C.POINTER(VVCDSPContext), C.c_int]
    lib.ff_vvc_dsp_init_cuda.restype = None
    lib.ff_vvc_dsp_cuda_last_error.restype = C.c_int
    lib.ff_vvc_dsp_cuda_error_string.restype = C.c_char_p
    lib.ff_vvc_dsp_cuda_reset_error.restype = None
    lib.ff_vvc_dsp_cuda_sizeof_table.restype = C.c_size_t
    _lib = lib
    return lib


class Context:
    """One decoder stream's CUDA context (vvc_cuda_ctx_create / _destroy)."""

    def __init__(self, device=0, stream=None):
        """stream: a cudaStream_t handle (int) all work is queued on.  None/0 lets the library
        create its own non-blocking stream - callers that also queue torch work (uploads, events)
        must then do so on `torch_stream()`, the legacy default stream does not order with it."""
        self.lib = load()
        self.handle = CTX()
        self._callbacks = []
        self.device = device
        rc = self.lib.vvc_cuda_ctx_create(C.byref(self.handle), device, stream)
        if rc != 0:
            raise VVCCudaError("vvc_cuda_ctx_create(device=%d) failed with %d: no usable CUDA device" % (device, rc))

    def close(self):
        if self.handle:
            self.lib.vvc_cuda_ctx_destroy(self.handle)
            self.handle = CTX()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise VVCCudaError("libvvcdsp_cuda error %d: %s" % (rc, self.lib.vvc_cuda_error_string(self.handle).decode()))

    def sync(self):
        self.check(self.lib.vvc_cuda_sync(self.handle))

    def notify(self, fn):
        """fn(status) is called from a driver thread once everything submitted so far has finished on the GPU
        (vvc_cuda_notify: the decoder's progress report for a submitted stage).  fn must not touch CUDA."""
        cb = NOTIFY_FN(lambda _opaque, status: fn(status))
        self._callbacks.append(cb)               # ctypes thunks must outlive the call
        self.check(self.lib.vvc_cuda_notify(self.handle, cb, None))

    @property
    def stream_ptr(self):
        return int(self.lib.vvc_cuda_stream(self.handle) or 0)

    def torch_stream(self):
        """The context's stream as a torch stream object (for events / `with torch.cuda.stream`)."""
        import torch
        return torch.cuda.ExternalStream(self.stream_ptr, device="cuda:%d" % self.device)

    def set_option(self, option, value):
        self.check(self.lib.vvc_cuda_ctx_set_option(self.handle, option, value))

    @property
    def launches(self):
        return int(self.lib.vvc_cuda_launch_count(self.handle))

    # ---- stages -------------------------------------------------------------------------
    def alf_frame(self, dst, src, ctbs_ptr, sets_ptr, sets_per_frame=0):
        self.check(self.lib.vvc_cuda_alf_frame(self.handle, C.byref(dst), C.byref(src), ctbs_ptr, sets_ptr, sets_per_frame))

    def alf_frame_host(self, dst, src, ctbs_ptr, sets_ptr, sets_per_frame=0):
        self.check(self.lib.vvc_cuda_alf_frame_host(self.handle, C.byref(dst), C.byref(src), ctbs_ptr, sets_ptr, sets_per_frame))

    def deblock_frame(self, dst, src, maps, direction):
        """direction 1 = vertical edges (DEBLOCK_V), 0 = horizontal edges (DEBLOCK_H)."""
        self.check(self.lib.vvc_cuda_deblock_frame(self.handle, C.byref(dst), C.byref(src), C.byref(maps), direction))

    def deblock_frame_host(self, dst, src, maps):
        self.check(self.lib.vvc_cuda_deblock_frame_host(self.handle, C.byref(dst), C.byref(src), C.byref(maps)))

    def deblock_params_frame(self, frame, tus_ptr, n_tus, mvfs_ptr, n_mvfs, ctbs_ptr, params, maps, direction):
        """Boundary strengths, filter lengths, beta / tc (LADF from `frame`) of one direction into the maps' device arrays."""
        self.check(self.lib.vvc_cuda_deblock_params_frame(self.handle, C.byref(frame), tus_ptr, n_tus, mvfs_ptr, n_mvfs, ctbs_ptr,
                                                          C.byref(params), C.byref(maps), direction))

    def sao_frame(self, dst, src, ctbs_ptr):
        self.check(self.lib.vvc_cuda_sao_frame(self.handle, C.byref(dst), C.byref(src), ctbs_ptr))

    def sao_frame_host(self, dst, src, ctbs_ptr):
        self.check(self.lib.vvc_cuda_sao_frame_host(self.handle, C.byref(dst), C.byref(src), ctbs_ptr))

    def inloop_frame(self, dst, src, desc):
        """DEBLOCK_V -> DEBLOCK_H -> SAO -> ALF on device-resident pictures and descriptors."""
        self.check(self.lib.vvc_cuda_inloop_frame(self.handle, C.byref(dst), C.byref(src), C.byref(desc)))

    def inloop_frame_host(self, dst, src, desc):
        self.check(self.lib.vvc_cuda_inloop_frame_host(self.handle, C.byref(dst), C.byref(src), C.byref(desc)))

    def itx_frame(self, frame, coeffs_ptr, tbs_ptr, n_tbs, log2_transform_range=15):
        """Inverse LFNST/transform + add_residual for a TB list; frame holds the prediction (in place)."""
        self.check(self.lib.vvc_cuda_itx_frame(self.handle, C.byref(frame), coeffs_ptr, tbs_ptr, n_tbs, log2_transform_range))

    def itx_frame_host(self, frame, coeffs_ptr, n_coeffs, tbs_ptr, n_tbs, log2_transform_range=15):
        self.check(self.lib.vvc_cuda_itx_frame_host(self.handle, C.byref(frame), coeffs_ptr, n_coeffs, tbs_ptr, n_tbs, log2_transform_range))

    def itx_frame_q(self, frame, coeffs_desc, tbs_ptr, n_tbs, log2_transform_range=15):
        """Same stage on a VVCCudaCoeffs (dense int32 or 16-bit window layout, optional dequant on the device)."""
        self.check(self.lib.vvc_cuda_itx_frame_q(self.handle, C.byref(frame), C.byref(coeffs_desc), tbs_ptr, n_tbs, log2_transform_range))

    def itx_frame_q_host(self, frame, coeffs_desc, tbs_ptr, n_tbs, log2_transform_range=15):
        self.check(self.lib.vvc_cuda_itx_frame_q_host(self.handle, C.byref(frame), C.byref(coeffs_desc), tbs_ptr, n_tbs, log2_transform_range))

    def lmcs_frame(self, frame, lut_ptr, ctb_enable_ptr=None):
        self.check(self.lib.vvc_cuda_lmcs_frame(self.handle, C.byref(frame), lut_ptr, ctb_enable_ptr))

    def lmcs_frame_host(self, frame, lut_ptr, ctb_enable_ptr=None):
        self.check(self.lib.vvc_cuda_lmcs_frame_host(self.handle, C.byref(frame), lut_ptr, ctb_enable_ptr))

    def pad_frame(self, frame, pad):
        """Replicate the border samples of every plane into its margin (the pre-padded DPB format)."""
        self.check(self.lib.vvc_cuda_pad_frame(self.handle, C.byref(frame), pad))

    def lmcs_chroma_scale(self, frame, vpdus_ptr, n, params_ptr, scales_ptr):
        """Per-VPDU chroma residual scales from the reconstructed luma (lmcs_derive_chroma_scale)."""
        self.check(self.lib.vvc_cuda_lmcs_chroma_scale(self.handle, C.byref(frame), vpdus_ptr, n, params_ptr, scales_ptr))

    def inter_frame(self, dst, refs, pbs_ptr, n_pbs, wp_ptr, prof_ptr, dmvr_out_ptr=None):
        """Motion compensation (+ DMVR / BDOF / PROF / GPM / weighted prediction) of a record list."""
        self.check(self.lib.vvc_cuda_inter_frame(self.handle, C.byref(dst), C.byref(refs), pbs_ptr, n_pbs, wp_ptr, prof_ptr, dmvr_out_ptr))

    def inter_frame_host(self, dst, refs, pbs_ptr, n_pbs, wp_ptr, n_wp, prof_ptr, n_prof, dmvr_out_ptr=None):
        self.check(self.lib.vvc_cuda_inter_frame_host(self.handle, C.byref(dst), C.byref(refs), pbs_ptr, n_pbs,
                                                      wp_ptr, n_wp, prof_ptr, n_prof, dmvr_out_ptr))

    def recon_frame(self, out, cur, refs, desc):
        """INTER -> residual -> LMCS -> deblock V/H -> SAO -> ALF on device-resident pictures and descriptors."""
        self.check(self.lib.vvc_cuda_recon_frame(self.handle, C.byref(out), C.byref(cur), C.byref(refs), C.byref(desc)))

    def recon_frame_host(self, out, refs, descs):
        """descs: ctypes array of VVCCudaReconDesc, one per picture of `out` (host pointers everywhere)."""
        self.check(self.lib.vvc_cuda_recon_frame_host(self.handle, C.byref(out), C.byref(refs), descs))

    def recon_frame_host_async(self, out, refs, descs):
        """The same, returning once everything is queued: every host buffer stays valid and untouched until sync()."""
        self.check(self.lib.vvc_cuda_recon_frame_host_async(self.handle, C.byref(out), C.byref(refs), descs))

    def intra_leaf_frame(self, frame, pbs_ptr, n_pbs, edges_ptr):
        """Intra leaf predictors (planar / DC / V / H / angular / MIP) of a list of independent blocks."""
        self.check(self.lib.vvc_cuda_intra_leaf_frame(self.handle, C.byref(frame), pbs_ptr, n_pbs, edges_ptr))

    def intra_leaf_frame_host(self, frame, pbs_ptr, n_pbs, edges_ptr, n_edges):
        self.check(self.lib.vvc_cuda_intra_leaf_frame_host(self.handle, C.byref(frame), pbs_ptr, n_pbs, edges_ptr, n_edges))

    def intra_pred_frame(self, frame, blks_ptr, n_blks):
        """intra_pred (reference lines prepared on the device) / intra_cclm_pred of one wavefront of blocks."""
        self.check(self.lib.vvc_cuda_intra_pred_frame(self.handle, C.byref(frame), blks_ptr, n_blks))

    def intra_pred_frame_host(self, frame, blks_ptr, n_blks):
        self.check(self.lib.vvc_cuda_intra_pred_frame_host(self.handle, C.byref(frame), blks_ptr, n_blks))

    def intra_recon_frame(self, frame, blks_ptr, blk_end, coeffs_desc, tbs_ptr, tb_end, log2_transform_range=15):
        """All-intra reconstruction, prediction and residual alternating wavefront by wavefront (blk_end / tb_end: host
        int32 arrays of running totals)."""
        assert len(blk_end) == len(tb_end) and blk_end.dtype == tb_end.dtype == np.int32
        self.check(self.lib.vvc_cuda_intra_recon_frame(self.handle, C.byref(frame), blks_ptr, blk_end.ctypes.data, C.byref(coeffs_desc),
                                                       tbs_ptr, tb_end.ctypes.data, len(blk_end), log2_transform_range))

    def intra_recon_frame_ordered(self, frame, blks_ptr, blk_end_ptr, coeffs_desc, tbs_ptr, tb_end_ptr, n_steps, n_blks, n_tbs, log2_transform_range=15):
        """All-intra reconstruction in one launch: steps in decoding order, dependencies resolved on the device."""
        self.check(self.lib.vvc_cuda_intra_recon_frame_ordered(self.handle, C.byref(frame), blks_ptr, blk_end_ptr, C.byref(coeffs_desc), tbs_ptr,
                                                               tb_end_ptr, n_steps, n_blks, n_tbs, log2_transform_range))

    def ciip_frame(self, dst, inter, blocks_ptr, n_blocks):
        """CIIP blend of the intra prediction in dst with the inter prediction picture."""
        self.check(self.lib.vvc_cuda_ciip_frame(self.handle, C.byref(dst), C.byref(inter), blocks_ptr, n_blocks))

    def ciip_frame_host(self, dst, inter, blocks_ptr, n_blocks):
        self.check(self.lib.vvc_cuda_ciip_frame_host(self.handle, C.byref(dst), C.byref(inter), blocks_ptr, n_blocks))

    def lmcs_rects(self, frame, lut_ptr, rects_ptr, n):
        self.check(self.lib.vvc_cuda_lmcs_rects(self.handle, C.byref(frame), lut_ptr, rects_ptr, n))
