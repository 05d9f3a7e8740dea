"""ctypes mirror of the reference's DSP function-pointer tables.

Layout follows ``struct VVCDSPContext`` and its seven sub-tables
(reference: libavcodec/vvc/vvcdsp.h:48-168).  The same mirror is used for

* the table filled by our drop-in ``ff_vvc_dsp_init_cuda`` (include/vvcdsp_cuda.h), and
* the table filled by the compiled reference (oracle/_ref, tests only),

so parity tests call both sides through identical signatures.
"""
import ctypes as C

u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)
i16p = C.POINTER(C.c_int16)
i32p = C.POINTER(C.c_int32)
intp = C.POINTER(C.c_int)
ptrdiff = C.c_ssize_t
vp = C.c_void_p

# --- inter (vvcdsp.h:48-93) -------------------------------------------------
PUT_FN = C.CFUNCTYPE(None, i16p, vp, ptrdiff, C.c_int, i8p, i8p, C.c_int)
PUT_UNI_FN = C.CFUNCTYPE(None, vp, ptrdiff, vp, ptrdiff, C.c_int, i8p, i8p, C.c_int)
PUT_UNI_W_FN = C.CFUNCTYPE(None, vp, ptrdiff, vp, ptrdiff, C.c_int, C.c_int, C.c_int, C.c_int, i8p, i8p, C.c_int)
AVG_FN = C.CFUNCTYPE(None, vp, ptrdiff, i16p, i16p, C.c_int, C.c_int)
W_AVG_FN = C.CFUNCTYPE(None, vp, ptrdiff, i16p, i16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
PUT_CIIP_FN = C.CFUNCTYPE(None, vp, ptrdiff, C.c_int, C.c_int, vp, ptrdiff, C.c_int)
PUT_GPM_FN = C.CFUNCTYPE(None, vp, ptrdiff, C.c_int, C.c_int, i16p, i16p, vp, C.c_int, C.c_int)
FETCH_FN = C.CFUNCTYPE(None, vp, vp, ptrdiff, C.c_int, C.c_int)
BDOF_FETCH_FN = C.CFUNCTYPE(None, vp, vp, ptrdiff, C.c_int, C.c_int, C.c_int, C.c_int)
PROF_GRAD_FN = C.CFUNCTYPE(None, vp, vp, ptrdiff, vp, ptrdiff, C.c_int, C.c_int, C.c_int)
APPLY_PROF_FN = C.CFUNCTYPE(None, vp, vp, i16p, i16p)
APPLY_PROF_UNI_FN = C.CFUNCTYPE(None, vp, ptrdiff, vp, i16p, i16p)
APPLY_PROF_UNI_W_FN = C.CFUNCTYPE(None, vp, ptrdiff, vp, i16p, i16p, C.c_int, C.c_int, C.c_int)
APPLY_BDOF_FN = C.CFUNCTYPE(None, vp, ptrdiff, vp, vp, C.c_int, C.c_int)
SAD_FN = C.CFUNCTYPE(C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int)
DMVR_FN = C.CFUNCTYPE(None, vp, vp, ptrdiff, C.c_int, C.c_ssize_t, C.c_ssize_t, C.c_int)


class VVCInterDSPContext(C.Structure):
    _fields_ = [
        ("put", PUT_FN * 2 * 2 * 7 * 2),          # [2][7][2][2]
        ("put_uni", PUT_UNI_FN * 2 * 2 * 7 * 2),
        ("put_uni_w", PUT_UNI_W_FN * 2 * 2 * 7 * 2),
        ("avg", AVG_FN),
        ("w_avg", W_AVG_FN),
        ("put_ciip", PUT_CIIP_FN),
        ("put_gpm", PUT_GPM_FN),
        ("fetch_samples", FETCH_FN),
        ("bdof_fetch_samples", BDOF_FETCH_FN),
        ("prof_grad_filter", PROF_GRAD_FN),
        ("apply_prof", APPLY_PROF_FN),
        ("apply_prof_uni", APPLY_PROF_UNI_FN),
        ("apply_prof_uni_w", APPLY_PROF_UNI_W_FN),
        ("apply_bdof", APPLY_BDOF_FN),
        ("sad", SAD_FN),
        ("dmvr", DMVR_FN * 2 * 2),
    ]


# --- intra (vvcdsp.h:97-111) ------------------------------------------------
LC = vp  # struct VVCLocalContext *
INTRA_CCLM_FN = C.CFUNCTYPE(None, LC, C.c_int, C.c_int, C.c_int, C.c_int)
LMCS_SCALE_CHROMA_FN = C.CFUNCTYPE(None, LC, intp, intp, C.c_int, C.c_int, C.c_int, C.c_int)
INTRA_PRED_FN = C.CFUNCTYPE(None, LC, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
PRED_PLANAR_FN = C.CFUNCTYPE(None, vp, vp, vp, C.c_int, C.c_int, ptrdiff)
PRED_MIP_FN = C.CFUNCTYPE(None, vp, vp, vp, C.c_int, C.c_int, ptrdiff, C.c_int, C.c_int)
PRED_DC_FN = PRED_PLANAR_FN
PRED_V_FN = C.CFUNCTYPE(None, vp, vp, C.c_int, C.c_int, ptrdiff)
PRED_H_FN = PRED_V_FN
PRED_ANGULAR_FN = C.CFUNCTYPE(None, vp, vp, vp, C.c_int, C.c_int, ptrdiff,
                              C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)


class VVCIntraDSPContext(C.Structure):
    _fields_ = [
        ("intra_cclm_pred", INTRA_CCLM_FN),
        ("lmcs_scale_chroma", LMCS_SCALE_CHROMA_FN),
        ("intra_pred", INTRA_PRED_FN),
        ("pred_planar", PRED_PLANAR_FN),
        ("pred_mip", PRED_MIP_FN),
        ("pred_dc", PRED_DC_FN),
        ("pred_v", PRED_V_FN),
        ("pred_h", PRED_H_FN),
        ("pred_angular_v", PRED_ANGULAR_FN),
        ("pred_angular_h", PRED_ANGULAR_FN),
    ]


# --- itx (vvcdsp.h:113-121) -------------------------------------------------
ADD_RES_FN = C.CFUNCTYPE(None, vp, intp, C.c_int, C.c_int, ptrdiff)
ADD_RES_JOINT_FN = C.CFUNCTYPE(None, vp, intp, C.c_int, C.c_int, ptrdiff, C.c_int, C.c_int)
PRED_RES_JOINT_FN = C.CFUNCTYPE(None, intp, C.c_int, C.c_int, C.c_int, C.c_int)
ITX_FN = C.CFUNCTYPE(None, intp, C.c_size_t, C.c_size_t, C.c_ssize_t, C.c_ssize_t)
BDPCM_FN = C.CFUNCTYPE(None, intp, C.c_int, C.c_int, C.c_int, C.c_int)

DCT2, DST7, DCT8 = 0, 1, 2
N_TX_TYPE = 3
N_TX_SIZE = 7


class VVCItxDSPContext(C.Structure):
    _fields_ = [
        ("add_residual", ADD_RES_FN),
        ("add_residual_joint", ADD_RES_JOINT_FN),
        ("pred_residual_joint", PRED_RES_JOINT_FN),
        ("itx", ITX_FN * N_TX_SIZE * N_TX_SIZE * N_TX_TYPE * N_TX_TYPE),  # [trh][trv][log2w][log2h]
        ("transform_bdpcm", BDPCM_FN),
    ]


# --- lmcs / lf / sao / alf (vvcdsp.h:123-158) --------------------------------
LMCS_FILTER_FN = C.CFUNCTYPE(None, vp, ptrdiff, C.c_int, C.c_int, vp)


class VVCLMCSDSPContext(C.Structure):
    _fields_ = [("filter", LMCS_FILTER_FN)]


LADF_FN = C.CFUNCTYPE(C.c_int, vp, ptrdiff)
LF_LUMA_FN = C.CFUNCTYPE(None, vp, ptrdiff, i32p, i32p, u8p, u8p, u8p, u8p, C.c_int)
LF_CHROMA_FN = LF_LUMA_FN


class VVCLFDSPContext(C.Structure):
    _fields_ = [
        ("ladf_level", LADF_FN * 2),
        ("filter_luma", LF_LUMA_FN * 2),
        ("filter_chroma", LF_CHROMA_FN * 2),
    ]


SAO_BAND_FN = C.CFUNCTYPE(None, vp, vp, ptrdiff, ptrdiff, i16p, C.c_int, C.c_int, C.c_int)
SAO_EDGE_FN = C.CFUNCTYPE(None, vp, vp, ptrdiff, i16p, C.c_int, C.c_int, C.c_int)
SAO_RESTORE_FN = C.CFUNCTYPE(None, vp, vp, ptrdiff, ptrdiff, vp, intp, C.c_int, C.c_int, C.c_int, u8p, u8p, u8p)


class VVCSAODSPContext(C.Structure):
    _fields_ = [
        ("band_filter", SAO_BAND_FN * 9),
        ("edge_filter", SAO_EDGE_FN * 9),
        ("edge_restore", SAO_RESTORE_FN * 2),
    ]


ALF_FILTER_FN = C.CFUNCTYPE(None, vp, ptrdiff, vp, ptrdiff, C.c_int, C.c_int, i16p, i16p, C.c_int)
ALF_CC_FN = C.CFUNCTYPE(None, vp, ptrdiff, vp, ptrdiff, C.c_int, C.c_int, C.c_int, C.c_int, i16p, C.c_int)
ALF_CLASSIFY_FN = C.CFUNCTYPE(None, intp, intp, vp, ptrdiff, C.c_int, C.c_int, C.c_int, intp)
ALF_RECON_FN = C.CFUNCTYPE(None, i16p, i16p, intp, intp, C.c_int, i16p, u8p, u8p)


class VVCALFDSPContext(C.Structure):
    _fields_ = [
        ("filter", ALF_FILTER_FN * 2),
        ("filter_cc", ALF_CC_FN),
        ("classify", ALF_CLASSIFY_FN),
        ("recon_coeff_and_clip", ALF_RECON_FN),
    ]


class VVCDSPContext(C.Structure):
    _fields_ = [
        ("inter", VVCInterDSPContext),
        ("intra", VVCIntraDSPContext),
        ("itx", VVCItxDSPContext),
        ("lmcs", VVCLMCSDSPContext),
        ("lf", VVCLFDSPContext),
        ("sao", VVCSAODSPContext),
        ("alf", VVCALFDSPContext),
    ]
