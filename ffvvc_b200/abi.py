"""ctypes mirror of include/vvcdsp_cuda.h (the C ABI of libvvcdsp_cuda.so).

The same plain-old-data descriptors are consumed by the CPU oracle (tests only), so the
structures live here once.  Nothing in this module touches the GPU.
"""
import ctypes as C

import numpy as np

EDGE_LEFT, EDGE_TOP, EDGE_RIGHT, EDGE_BOTTOM = 1, 2, 4, 8


OPT_GENERIC_KERNELS = 1      # vvc_cuda_ctx_set_option()
OPT_ALF_WIDE_MULTIPLY = 2
OPT_INTER_TMA = 3
OPT_REF_PAD = 4
INTER_TMA_DEFAULT = 0       # what vvc_cuda_ctx_create() starts with


class VVCCudaFrame(C.Structure):
    _fields_ = [
        ("data", C.c_void_p * 3),
        ("stride", C.c_ssize_t * 3),
        ("batch_stride", C.c_ssize_t * 3),
        ("width", C.c_int32), ("height", C.c_int32),
        ("hshift", C.c_int32), ("vshift", C.c_int32),
        ("bit_depth", C.c_int32), ("ctb_log2", C.c_int32),
        ("batch", C.c_int32), ("chroma_format_idc", C.c_int32),
    ]


class VVCCudaALFCtb(C.Structure):
    _fields_ = [
        ("ctb_flag", C.c_uint8 * 3),
        ("filt_set_idx_y", C.c_uint8),
        ("chroma_alt_idx", C.c_uint8 * 2),
        ("cc_idc", C.c_uint8 * 2),
        ("edges", C.c_uint8),
        ("reserved", C.c_uint8 * 3),
    ]


class VVCCudaALFSets(C.Structure):
    _fields_ = [
        ("luma_coeff", C.c_int16 * 12 * 25 * 8),
        ("luma_clip_idx", C.c_uint8 * 12 * 25 * 8),
        ("chroma_coeff", C.c_int16 * 6 * 8),
        ("chroma_clip_idx", C.c_uint8 * 6 * 8),
        ("cc_coeff", C.c_int16 * 7 * 5 * 2),
    ]


ALF_CTB_DTYPE = np.dtype([
    ("ctb_flag", np.uint8, (3,)), ("filt_set_idx_y", np.uint8), ("chroma_alt_idx", np.uint8, (2,)),
    ("cc_idc", np.uint8, (2,)), ("edges", np.uint8), ("reserved", np.uint8, (3,))])
ALF_SETS_DTYPE = np.dtype([
    ("luma_coeff", np.int16, (8, 25, 12)), ("luma_clip_idx", np.uint8, (8, 25, 12)),
    ("chroma_coeff", np.int16, (8, 6)), ("chroma_clip_idx", np.uint8, (8, 6)),
    ("cc_coeff", np.int16, (2, 5, 7))])
assert ALF_CTB_DTYPE.itemsize == C.sizeof(VVCCudaALFCtb)
assert ALF_SETS_DTYPE.itemsize == C.sizeof(VVCCudaALFSets)


class FrameGeom:
    """Geometry of a 4:2:0 (default) high-bit-depth picture ring."""

    def __init__(self, width, height, bit_depth=10, ctb_log2=7, hshift=1, vshift=1, batch=1,
                 chroma_format_idc=1, pitch_align=128):
        self.width, self.height, self.bit_depth, self.ctb_log2 = width, height, bit_depth, ctb_log2
        self.hshift, self.vshift, self.batch = hshift, vshift, batch
        self.chroma_format_idc = chroma_format_idc
        self.pitch_align = pitch_align

    @property
    def ctb_size(self):
        return 1 << self.ctb_log2

    @property
    def ctb_cols(self):
        return (self.width + self.ctb_size - 1) >> self.ctb_log2

    @property
    def ctb_rows(self):
        return (self.height + self.ctb_size - 1) >> self.ctb_log2

    @property
    def ctb_count(self):
        return self.ctb_cols * self.ctb_rows

    def plane_wh(self, c):
        if c == 0:
            return self.width, self.height
        return self.width >> self.hshift, self.height >> self.vshift

    def plane_pitch(self, c):
        """Row pitch in samples: a multiple of pitch_align (128 samples = 256 bytes)."""
        w, _ = self.plane_wh(c)
        a = self.pitch_align
        return (w + a - 1) // a * a

    def plane_shape(self, c):
        return (self.batch, self.plane_wh(c)[1], self.plane_pitch(c))

    @property
    def luma_pixels(self):
        return self.width * self.height * self.batch


def frame_desc(geom, ptrs, pitches_bytes, batch_strides_bytes):
    """Build a VVCCudaFrame from raw plane base addresses (host or device)."""
    f = VVCCudaFrame()
    for c in range(3):
        f.data[c] = ptrs[c]
        f.stride[c] = pitches_bytes[c]
        f.batch_stride[c] = batch_strides_bytes[c]
    f.width, f.height = geom.width, geom.height
    f.hshift, f.vshift = geom.hshift, geom.vshift
    f.bit_depth, f.ctb_log2 = geom.bit_depth, geom.ctb_log2
    f.batch, f.chroma_format_idc = geom.batch, geom.chroma_format_idc
    return f


def frame_from_numpy(geom, planes):
    """planes: three C-contiguous uint16 arrays of shape geom.plane_shape(c)."""
    ptrs, pitches, bstr = [], [], []
    for c, p in enumerate(planes):
        assert p.dtype == np.uint16 and p.flags.c_contiguous and p.shape == geom.plane_shape(c), (c, p.shape)
        ptrs.append(p.ctypes.data)
        pitches.append(p.strides[1])
        bstr.append(p.strides[0])
    return frame_desc(geom, ptrs, pitches, bstr)


def alloc_planes(geom, fill=None):
    planes = [np.zeros(geom.plane_shape(c), dtype=np.uint16) for c in range(3)]
    if fill is not None:
        for p in planes:
            p[...] = fill
    return planes


# ---- deblocking / SAO descriptors (include/vvcdsp_cuda.h) ----------------------------------
class VVCCudaDbkEdge(C.Structure):
    _fields_ = [("tc", C.c_uint16), ("beta", C.c_uint8), ("max_len", C.c_uint8)]


class VVCCudaDeblockMaps(C.Structure):
    _fields_ = [
        ("edge", C.c_void_p * 3 * 2),
        ("pitch", C.c_int32 * 3 * 2),
        ("rows", C.c_int32 * 3 * 2),
        ("size", C.c_int64 * 3 * 2),
    ]


class VVCCudaSAOCtb(C.Structure):
    _fields_ = [
        ("type_idx", C.c_uint8 * 3), ("band_position", C.c_uint8 * 3), ("eo_class", C.c_uint8 * 3),
        ("restore", C.c_uint8), ("no_filter", C.c_uint8), ("reserved", C.c_uint8),
        ("offset_val", C.c_int16 * 5 * 3),
    ]


DBK_EDGE_DTYPE = np.dtype([("tc", np.uint16), ("beta", np.uint8), ("max_len", np.uint8)])
SAO_CTB_DTYPE = np.dtype([
    ("type_idx", np.uint8, (3,)), ("band_position", np.uint8, (3,)), ("eo_class", np.uint8, (3,)),
    ("restore", np.uint8), ("no_filter", np.uint8), ("reserved", np.uint8), ("offset_val", np.int16, (3, 5))])
assert DBK_EDGE_DTYPE.itemsize == C.sizeof(VVCCudaDbkEdge) == 4
assert SAO_CTB_DTYPE.itemsize == C.sizeof(VVCCudaSAOCtb) == 42


# deblocking parameters on the device: transform-unit / motion / CTB records (include/vvcdsp_cuda.h)
DBK_TU_LUMA, DBK_TU_CHROMA = 1, 2
DBK_CBF_Y, DBK_CBF_CB, DBK_CBF_CR, DBK_JOINT, DBK_BDPCM_Y, DBK_BDPCM_C = 1, 2, 4, 8, 16, 32
DBK_CU_SUBBLOCK = 1
DBK_TU_DTYPE = np.dtype([("x0", np.uint16), ("y0", np.uint16), ("log2_w", np.uint8), ("log2_h", np.uint8), ("planes", np.uint8),
                         ("flags", np.uint8), ("qp", np.int8, (3,)), ("cu_flags", np.uint8), ("cu_dx", np.uint8), ("cu_dy", np.uint8),
                         ("cb_log2_w", np.uint8), ("cb_log2_h", np.uint8), ("pic", np.uint8), ("reserved", np.uint8, (3,))])
DBK_MVF_DTYPE = np.dtype([("x0", np.uint16), ("y0", np.uint16), ("w4", np.uint8), ("h4", np.uint8), ("pred_flag", np.uint8),
                          ("ciip_flag", np.uint8), ("ref_pic", np.int16, (2,)), ("mv", np.int32, (2, 2)), ("pic", np.uint8),
                          ("reserved", np.uint8, (3,))])
DBK_CTB_DTYPE = np.dtype([("beta_offset", np.int8, (3,)), ("tc_offset", np.int8, (3,)), ("no_left", np.uint8), ("no_top", np.uint8)])
assert DBK_TU_DTYPE.itemsize == 20 and DBK_MVF_DTYPE.itemsize == 32 and DBK_CTB_DTYPE.itemsize == 8


class VVCCudaDbkParams(C.Structure):
    _fields_ = [("qp_bd_offset", C.c_int32), ("ladf_enabled", C.c_int32), ("num_ladf_intervals", C.c_int32),
                ("ladf_lowest_interval_qp_offset", C.c_int32), ("ladf_qp_offset", C.c_int32 * 4),
                ("ladf_interval_lower_bound", C.c_int32 * 5)]


def deblock_map_shape(geom, direction, c):
    """(rows, pitch) of the edge map for plane c; direction 1 = vertical edges, 0 = horizontal."""
    w, h = geom.plane_wh(c)
    grid = 8 if c else 4
    seg = (4 >> (geom.vshift if direction else geom.hshift)) if c else 4
    if direction:
        return (h + seg - 1) // seg, (w + grid - 1) // grid
    return (h + grid - 1) // grid, (w + seg - 1) // seg


def deblock_maps_desc(geom, arrays, ptr_of=lambda a: a.ctypes.data):
    """arrays[dir][c]: DBK_EDGE_DTYPE arrays of shape (batch, rows, pitch) -> VVCCudaDeblockMaps."""
    m = VVCCudaDeblockMaps()
    for d in range(2):
        for c in range(3):
            a = arrays[d][c]
            rows, pitch = deblock_map_shape(geom, d, c)
            assert a.shape == (geom.batch, rows, pitch), (a.shape, rows, pitch)
            m.edge[d][c] = ptr_of(a)
            m.pitch[d][c] = pitch
            m.rows[d][c] = rows
            m.size[d][c] = rows * pitch
    return m


class VVCCudaDbkSide(C.Structure):
    _fields_ = [("tus", C.c_void_p), ("mvfs", C.c_void_p), ("ctbs", C.c_void_p), ("params", C.POINTER(VVCCudaDbkParams)),
                ("n_tus", C.c_int32), ("n_mvfs", C.c_int32)]


class VVCCudaInloopDesc(C.Structure):
    _fields_ = [
        ("deblock", C.POINTER(VVCCudaDeblockMaps)),
        ("sao", C.c_void_p), ("alf", C.c_void_p), ("alf_sets", C.c_void_p),
        ("alf_sets_per_frame", C.c_int32), ("reserved", C.c_int32),
        ("dbk_side", C.POINTER(VVCCudaDbkSide)),
    ]


def inloop_desc(maps_desc, sao_ptr, alf_ptr, sets_ptr, sets_per_frame=0):
    d = VVCCudaInloopDesc()
    d.deblock = C.pointer(maps_desc)
    d.sao, d.alf, d.alf_sets = sao_ptr, alf_ptr, sets_ptr
    d.alf_sets_per_frame = sets_per_frame
    d._keep = maps_desc
    return d


# ---- residual stage (include/vvcdsp_cuda.h) ---------------------------------------------------
TB_TS, TB_BDPCM, TB_BDPCM_VERT, TB_JOINT, TB_STORE_RESIDUAL = 1, 2, 4, 8, 16


class VVCCudaTB(C.Structure):
    _fields_ = [
        ("coeff_offset", C.c_uint32), ("x0", C.c_uint16), ("y0", C.c_uint16),
        ("log2_w", C.c_uint8), ("log2_h", C.c_uint8), ("c_idx", C.c_uint8),
        ("trh", C.c_uint8), ("trv", C.c_uint8), ("nzw", C.c_uint8), ("nzh", C.c_uint8),
        ("flags", C.c_uint8), ("lfnst", C.c_uint8), ("joint_sign", C.c_int8),
        ("joint_shift", C.c_uint8), ("joint_c_idx", C.c_uint8), ("pic", C.c_uint8),
        ("reserved", C.c_uint8), ("chroma_scale", C.c_uint16),
    ]


TB_DTYPE = np.dtype([
    ("coeff_offset", np.uint32), ("x0", np.uint16), ("y0", np.uint16),
    ("log2_w", np.uint8), ("log2_h", np.uint8), ("c_idx", np.uint8), ("trh", np.uint8), ("trv", np.uint8),
    ("nzw", np.uint8), ("nzh", np.uint8), ("flags", np.uint8), ("lfnst", np.uint8), ("joint_sign", np.int8),
    ("joint_shift", np.uint8), ("joint_c_idx", np.uint8), ("pic", np.uint8),
    ("reserved", np.uint8), ("chroma_scale", np.uint16)], align=True)
assert TB_DTYPE.itemsize == C.sizeof(VVCCudaTB) == 24, (TB_DTYPE.itemsize, C.sizeof(VVCCudaTB))


# compact coefficient layouts + fused dequantisation
COEFF_DENSE32, COEFF_WINDOW16 = 0, 1
TB_QUANT_DTYPE = np.dtype([("qp", np.uint8), ("dep_quant", np.uint8), ("sl_id", np.uint8), ("reserved", np.uint8)])
SCALING_LIST_DTYPE = np.dtype([("matrix_rec", np.uint8, (28, 64)), ("dc_rec", np.uint8, (14,)), ("reserved", np.uint8, (2,))])
assert TB_QUANT_DTYPE.itemsize == 4 and SCALING_LIST_DTYPE.itemsize == 28 * 64 + 16


class VVCCudaCoeffs(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_size_t), ("format", C.c_int32), ("reserved", C.c_int32),
                ("quant", C.c_void_p), ("scaling", C.c_void_p), ("lmcs_scales", C.c_void_p)]


def coeffs_desc(data_ptr, n, fmt=COEFF_DENSE32, quant_ptr=None, scaling_ptr=None, lmcs_scales_ptr=None):
    d = VVCCudaCoeffs()
    d.data, d.n, d.format, d.quant, d.scaling, d.lmcs_scales = data_ptr, n, fmt, quant_ptr, scaling_ptr, lmcs_scales_ptr
    return d


# LMCS chroma residual scaling: VPDU records and the VVCLMCS fields the derivation reads
LMCS_VPDU_DTYPE = np.dtype([("x", np.uint16), ("y", np.uint16), ("avail_l", np.uint8), ("avail_t", np.uint8),
                            ("pic", np.uint8), ("reserved", np.uint8)])
LMCS_PARAMS_DTYPE = np.dtype([("pivot", np.uint16, (17,)), ("chroma_scale_coeff", np.uint16, (16,)),
                              ("min_bin_idx", np.uint8), ("max_bin_idx", np.uint8)])
assert LMCS_VPDU_DTYPE.itemsize == 8 and LMCS_PARAMS_DTYPE.itemsize == 68


def pack_window16(tbs, coeffs):
    """DENSE32 -> WINDOW16: (tbs with coeff_offset rewritten, int16 array).  Values outside each TB's
    nzh x nzw window must be zero (they do not exist in the compact layout)."""
    w = 1 << tbs["log2_w"].astype(np.int64)
    nzw = np.minimum(tbs["nzw"].astype(np.int64), w)
    nzh = np.minimum(tbs["nzh"].astype(np.int64), 1 << tbs["log2_h"].astype(np.int64))
    sizes = nzw * nzh
    offs = np.concatenate([[0], np.cumsum(sizes)])
    out = np.zeros(int(offs[-1]) + 8, dtype=np.int16)
    t2 = tbs.copy()
    t2["coeff_offset"] = offs[:-1]
    # group by (w, nzw, nzh) so each group is one fancy-indexed gather
    key = (w << 16) | (nzw << 8) | nzh
    for k in np.unique(key):
        idx = np.nonzero(key == k)[0]
        ww, a, b = int(k >> 16), int((k >> 8) & 255), int(k & 255)
        yy, xx = np.meshgrid(np.arange(b), np.arange(a), indexing="ij")
        src = tbs["coeff_offset"][idx].astype(np.int64)[:, None] + (yy * ww + xx).reshape(1, -1)
        dst = offs[idx][:, None] + (yy * a + xx).reshape(1, -1)
        out[dst] = coeffs[src].astype(np.int16)
    return t2, out


class VVCCudaRect(C.Structure):
    _fields_ = [("x", C.c_uint16), ("y", C.c_uint16), ("w", C.c_uint16), ("h", C.c_uint16),
                ("pic", C.c_uint16), ("reserved", C.c_uint16)]


RECT_DTYPE = np.dtype([("x", np.uint16), ("y", np.uint16), ("w", np.uint16), ("h", np.uint16),
                       ("pic", np.uint16), ("reserved", np.uint16)])
assert RECT_DTYPE.itemsize == C.sizeof(VVCCudaRect) == 12


# ---- inter prediction stage (include/vvcdsp_cuda.h) -------------------------------------------
PB_LUMA, PB_CHROMA = 1, 2
PB_DMVR, PB_BDOF, PB_PROF0, PB_PROF1, PB_GPM, PB_WEIGHTED = 1, 2, 4, 8, 16, 32
PF_L0, PF_L1, PF_BI = 1, 2, 3

PB_DTYPE = np.dtype([
    ("x0", np.uint16), ("y0", np.uint16), ("w", np.uint8), ("h", np.uint8), ("planes", np.uint8),
    ("pred_flag", np.uint8), ("ref", np.uint8, (2,)), ("pic", np.uint8), ("flags", np.uint8),
    ("mv", np.int32, (2, 2)), ("filt", np.uint8), ("bcw_idx", np.uint8), ("wp", np.uint16),
    ("prof", np.uint16), ("gpm_step_x", np.int16), ("gpm_step_y", np.int16), ("reserved", np.uint16),
    ("gpm_weights", np.int32)], align=True)
WP_DTYPE = np.dtype([("weight", np.int16, (2, 3)), ("offset", np.int16, (2, 3)),
                     ("log2_denom", np.uint8, (2,)), ("reserved", np.uint8, (2,))], align=True)
PROF_DTYPE = np.dtype([("diff_mv_x", np.int16, (2, 16)), ("diff_mv_y", np.int16, (2, 16))])
DMVR_OUT_DTYPE = np.dtype([("mv", np.int32, (2, 2)), ("min_sad", np.int32), ("bdof_applied", np.int32)])
assert PB_DTYPE.itemsize == 44 and WP_DTYPE.itemsize == 28 and PROF_DTYPE.itemsize == 128 and DMVR_OUT_DTYPE.itemsize == 24


class VVCCudaReconDesc(C.Structure):
    _fields_ = [
        ("pbs", C.c_void_p), ("wp", C.c_void_p), ("prof", C.c_void_p), ("dmvr_out", C.c_void_p),
        ("n_pbs", C.c_int32), ("n_wp", C.c_int32), ("n_prof", C.c_int32), ("log2_transform_range", C.c_int32),
        ("lmcs_fwd_lut", C.c_void_p), ("lmcs_rects", C.c_void_p),
        ("n_lmcs_rects", C.c_int32), ("n_tbs", C.c_int32),
        ("coeffs", C.c_void_p), ("n_coeffs", C.c_size_t), ("tbs", C.c_void_p),
        ("coeff_format", C.c_int32), ("ref_slots", C.c_uint32), ("quant", C.c_void_p), ("scaling", C.c_void_p),
        ("lmcs_inv_lut", C.c_void_p), ("lmcs_ctb_enable", C.c_void_p),
        ("inloop", VVCCudaInloopDesc),
        ("arena", C.c_void_p), ("arena_bytes", C.c_size_t),
        ("lmcs_vpdus", C.c_void_p), ("lmcs_params", C.c_void_p), ("n_lmcs_vpdus", C.c_int32), ("n_luma_tbs", C.c_int32),
    ]


# ---- intra leaf predictors / CIIP (include/vvcdsp_cuda.h) ----------------------------------------------
INTRA_PLANAR, INTRA_DC, INTRA_VERT, INTRA_HORZ, INTRA_ANGULAR_V, INTRA_ANGULAR_H, INTRA_MIP = range(7)
INTRA_PDPC, INTRA_MIP_TRANSPOSED = 1, 2
INTRA_PB_DTYPE = np.dtype([("x0", np.uint16), ("y0", np.uint16), ("w", np.uint8), ("h", np.uint8), ("c_idx", np.uint8),
                           ("pic", np.uint8), ("kind", np.uint8), ("mode", np.int8), ("ref_idx", np.uint8),
                           ("filter_flag", np.uint8), ("flags", np.uint8), ("reserved", np.uint8, 3),
                           ("top", np.uint32), ("left", np.uint32)])
CIIP_DTYPE = np.dtype([("x0", np.uint16), ("y0", np.uint16), ("w", np.uint8), ("h", np.uint8), ("c_idx", np.uint8),
                       ("pic", np.uint8), ("intra_weight", np.uint8), ("reserved", np.uint8, 3)])
assert INTRA_PB_DTYPE.itemsize == 24 and CIIP_DTYPE.itemsize == 12

# full intra prediction (edge preparation on the device) and CCLM
INTRA_KIND_PRED, INTRA_KIND_MIP, INTRA_KIND_CCLM = 0, 1, 2
INTRA_F_ISP, INTRA_F_BDPCM, INTRA_F_MIP_TRANSP, INTRA_F_UP_LEFT = 1, 2, 4, 8
INTRA_F_LUMA_AVAIL_T, INTRA_F_LUMA_AVAIL_L, INTRA_F_COLLOCATED = 16, 32, 64
INTRA_BLK_DTYPE = np.dtype([("x0", np.uint16), ("y0", np.uint16), ("w", np.uint8), ("h", np.uint8), ("c_idx", np.uint8),
                            ("pic", np.uint8), ("kind", np.uint8), ("pred_mode", np.uint8), ("ref_idx", np.uint8),
                            ("flags", np.uint8), ("cb_w", np.uint8), ("cb_h", np.uint8), ("avail_left", np.uint8),
                            ("avail_top", np.uint8)])
assert INTRA_BLK_DTYPE.itemsize == 16


def recon_arena(handle, geom1, alloc, *, pbs, wp, prof, tbs, coeffs, coeff_format, quant, scaling, inv_lut, maps, sao, alf, sets,
                ref_slots=0, log2_transform_range=15):
    """One picture's descriptors in ONE arena (vvc_cuda_recon_arena_size / _bind, include/vvcdsp_cuda.h): returns
    (VVCCudaReconDesc, VVCCudaDeblockMaps, arena).  handle: the loaded library; geom1: FrameGeom of one picture;
    alloc(nbytes) -> (object to keep alive, address) of 256-byte aligned (ideally pinned) host memory; maps[dir][c]:
    (rows, pitch) DBK_EDGE arrays of this picture.  What the parser would write in place is copied here once."""
    f = frame_desc(geom1, [0, 0, 0], [0, 0, 0], [0, 0, 0])
    d, m = VVCCudaReconDesc(), VVCCudaDeblockMaps()
    d.n_pbs, d.n_wp, d.n_prof, d.n_tbs, d.n_coeffs = len(pbs), len(wp), len(prof), len(tbs), len(coeffs)
    d.coeff_format, d.log2_transform_range, d.ref_slots = coeff_format, log2_transform_range, ref_slots
    for dr in range(2):
        for c in range(3):
            rows, pitch = deblock_map_shape(geom1, dr, c)
            m.pitch[dr][c], m.rows[dr][c], m.size[dr][c] = pitch, rows, rows * pitch
    d.inloop.deblock = C.pointer(m)
    size = handle.vvc_cuda_recon_arena_size(C.byref(f), C.byref(d))
    assert size > 0
    keep, addr = alloc(size)
    assert addr % 256 == 0
    rc = handle.vvc_cuda_recon_arena_bind(C.byref(f), C.byref(d), C.byref(m), addr)
    assert rc == 0, rc

    def put(dst, a):
        a = np.ascontiguousarray(a)
        assert addr <= dst and dst + a.nbytes <= addr + size
        C.memmove(dst, a.ctypes.data, a.nbytes)

    put(d.pbs, pbs); put(d.wp, wp); put(d.prof, prof); put(d.tbs, tbs); put(d.coeffs, coeffs)
    if quant is not None:
        put(d.quant, quant); put(d.scaling, scaling)
    else:
        d.quant, d.scaling = None, None
    put(d.lmcs_inv_lut, inv_lut)
    d.lmcs_fwd_lut, d.lmcs_rects, d.lmcs_ctb_enable = None, None, None
    d.lmcs_vpdus, d.lmcs_params = None, None
    for dr in range(2):
        for c in range(3):
            put(m.edge[dr][c], maps[dr][c])
    put(d.inloop.sao, sao); put(d.inloop.alf, alf); put(d.inloop.alf_sets, sets)
    return d, m, keep
