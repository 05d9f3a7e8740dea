// Coefficient input of the residual stage: the two layouts of VVCCudaCoeffs and dequant() on the device.
//
// Reference: dequant (libavcodec/vvc/vvc_intra.c:397-417) = derive_scale_m (:341-384), derive_qp's shift
// (:277-308), derive_scale (:311-338), scale_coeff (:387-395).  The host passes tb->qp after derive_qp's
// offsets and clip (syntax-level values: cu->qp, act, joint CbCr) and the scaling matrix id of Table 38;
// shift, rounding offset, level scale, the matrix entry of every position and the clip are computed here.
#pragma once
#include "common.cuh"

struct CoefSrc {
    const int32_t            *dense;     // VVC_CUDA_COEFF_DENSE32 (NULL otherwise)
    const int16_t            *window;    // VVC_CUDA_COEFF_WINDOW16 (NULL otherwise)
    const VVCCudaTBQuant     *quant;     // NULL: already dequantised
    const VVCCudaScalingList *scaling;
    const uint16_t           *lmcs_scales; // VVCCudaCoeffs.lmcs_scales: per-VPDU chroma residual scales or NULL
    int                       range, bd;
};

// VVCCudaTB.chroma_scale -> the scale lmcs_scale_chroma multiplies with (0: no scaling)
__device__ __forceinline__ int tb_chroma_scale(const CoefSrc &s, int field)
{
    return !field ? 0 : s.lmcs_scales ? (int)__ldg(s.lmcs_scales + field - 1) : field;
}

// per-TB view
struct TbCoef {
    const int32_t *d32;
    const int16_t *w16;
    int            pitch;                // elements per row of the layout
    int            nzw, nzh;             // window bounds (WINDOW16: nothing exists outside)
    // dequant
    bool           dq;
    int            scale, shift, add, range;
    const uint8_t *m;                    // scaling matrix of this TB's id or NULL (flat 16)
    int            lms, l2w, l2h, dc;    // log2 matrix size, log2 TB size, DC entry (-1: none)
};

// MODE: bit 0 = WINDOW16 layout, bit 1 = dequantise (compile-time so the dense, dequantised path costs nothing)
template <int MODE>
__device__ __forceinline__ TbCoef tb_coef(const CoefSrc &s, int ti, uint32_t coeff_offset, int l2w, int l2h, int nzw, int nzh, bool ts)
{
    TbCoef t;
    t.d32 = (MODE & 1) ? nullptr : s.dense + coeff_offset;
    t.w16 = (MODE & 1) ? s.window + coeff_offset : nullptr;
    t.pitch = (MODE & 1) ? nzw : (1 << l2w);
    t.nzw = nzw; t.nzh = nzh;
    t.dq = (MODE & 2) != 0;
    t.m = nullptr; t.dc = -1; t.lms = 0; t.l2w = l2w; t.l2h = l2h; t.range = s.range;
    t.scale = t.shift = t.add = 0;
    if (t.dq) {
        const uint32_t q = __ldg(reinterpret_cast<const uint32_t *>(s.quant) + ti);
        const int qp0 = q & 0xff, dep = (q >> 8) & 1, sl = (q >> 16) & 0xff;
        const int log_sum = l2w + l2h, rect = ts ? 0 : (log_sum & 1);
        t.shift = ts ? 10 : s.bd + rect + (log_sum >> 1) + 10 - s.range + dep;
        t.add = (1 << t.shift) >> 1;
        const int qp = qp0 + (dep && !ts);
        const int rem = qp % 6, div = qp / 6;
        // level_scale[2][6] (vvc_intra.c:331-334)
        const int ls = rect ? (rem == 0 ? 57 : rem == 1 ? 64 : rem == 2 ? 72 : rem == 3 ? 80 : rem == 4 ? 90 : 102)
                            : (rem == 0 ? 40 : rem == 1 ? 45 : rem == 2 ? 51 : rem == 3 ? 57 : rem == 4 ? 64 : 72);
        t.scale = ls << div;
        if (sl && s.scaling) {
            const int id = sl - 1;
            t.lms = id < 2 ? 1 : id < 8 ? 2 : 3;
            t.m = s.scaling->matrix_rec[id];
            if (id >= 14)
                t.dc = s.scaling->dc_rec[id - 14];
        }
    }
    return t;
}

// raw value at (y, x) of the TB: zero outside the window for WINDOW16
template <int MODE>
__device__ __forceinline__ int coef_raw(const TbCoef &t, int y, int x)
{
    if (!(MODE & 1))
        return __ldg(t.d32 + y * t.pitch + x);
    return (y < t.nzh && x < t.nzw) ? (int)__ldg(t.w16 + y * t.pitch + x) : 0;
}

// scale_coeff with the matrix entry of position (y, x); int arithmetic wraps like the reference's (-fwrapv)
template <int MODE>
__device__ __forceinline__ int coef_dequant(const TbCoef &t, int v, int y, int x)
{
    if (!(MODE & 2))
        return v;           // (a zero level stays zero: the rounding offset is below 1 << shift)
    int m = 16;
    if (t.m) {
        m = __ldg(t.m + (((y << t.lms) >> t.l2h) << t.lms) + ((x << t.lms) >> t.l2w));
        if (t.dc >= 0 && !(x | y))
            m = t.dc;
    }
    const int r = (int)((uint32_t)v * (uint32_t)t.scale * (uint32_t)m + (uint32_t)t.add) >> t.shift;
    return d_clip_sbits(r, t.range);
}

template <int MODE>
__device__ __forceinline__ int coef_load(const TbCoef &t, int y, int x) { return coef_dequant<MODE>(t, coef_raw<MODE>(t, y, x), y, x); }

static inline int coef_mode(const VVCCudaCoeffs *c) { return (c->format == VVC_CUDA_COEFF_WINDOW16 ? 1 : 0) | (c->quant ? 2 : 0); }
