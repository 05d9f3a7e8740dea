/*
 * Oracle: inverse LFNST, inverse transforms and residual addition for a list of transform blocks
 * (TEST INFRASTRUCTURE, see vvc_oracle.h).
 *
 * Restates the reference with plain matrix products:
 *   2-D / 1-D drivers   itx_2d / itx_1d, scale_clip, scale      libavcodec/vvc/vvcdsp.c:67-138
 *   DCT-II              ff_vvc_inv_dct2_{2..64} (partial butterflies, zero-out guards G2..G16,
 *                       libavcodec/vvc/vvc_itx_1d.c:64-655) - here the equivalent matrix product with
 *                       the same "inputs are ignored in power-of-two groups" rule
 *   DST-VII / DCT-VIII  matrix_mul                              libavcodec/vvc/vvc_itx_1d.c:657-706
 *   LFNST               ff_vvc_inv_lfnst_1d :708-721, gather/scatter of ilfnst_transform
 *                       libavcodec/vvc/vvc_intra.c:65-127
 *   BDPCM               transform_bdpcm                         libavcodec/vvc/vvcdsp_template.c:76-95
 *   add_residual(_joint)                                       libavcodec/vvc/vvcdsp_template.c:32-63
 * The DCT-II matrices were recovered from the reference's own 1-D functions (tools/gen_tables.py),
 * so product == butterfly exactly (both are exact integer arithmetic mod 2^32).
 */
#include "vvc_oracle.h"
#include "vvc_tables_c.h"

static const int8_t *tx_matrix(int type, int n)
{
    if (type == 0)
        switch (n) {
        case 2:  return &vvct_dct2_2[0][0];   case 4:  return &vvct_dct2_4[0][0];
        case 8:  return &vvct_dct2_8[0][0];   case 16: return &vvct_dct2_16[0][0];
        case 32: return &vvct_dct2_32[0][0];  default: return &vvct_dct2_64[0][0];
        }
    if (type == 1)
        switch (n) {
        case 4:  return &vvct_dst7_4[0][0];   case 8:  return &vvct_dst7_8[0][0];
        case 16: return &vvct_dst7_16[0][0];  default: return &vvct_dst7_32[0][0];
        }
    switch (n) {
    case 4:  return &vvct_dct8_4[0][0];   case 8:  return &vvct_dct8_8[0][0];
    case 16: return &vvct_dct8_16[0][0];  default: return &vvct_dct8_32[0][0];
    }
}

/* how many inputs the reference's 1-D transform actually reads for a declared nz */
static int inputs_read(int type, int n, int nz)
{
    int r;
    if (type != 0)
        return nz;                       /* matrix_mul reads exactly nz (<= 16) */
    r = nz <= 2 ? 2 : nz <= 4 ? 4 : nz <= 8 ? 8 : nz <= 16 ? 16 : 32;
    return o_min(r, o_min(n, 32));       /* DCT2-64 never reads inputs 32..63 (:498) */
}

/* out[i] = sum_j in[j] * M[j][i], in/out strided, in place */
static void inv_1d(int *v, ptrdiff_t stride, int type, int n, int nz)
{
    const int8_t *m = tx_matrix(type, n);
    const int rd = inputs_read(type, n, nz);
    int in[32], out[64];
    if (n == 1)
        return;
    for (int j = 0; j < rd; j++)
        in[j] = v[j * stride];
    for (int i = 0; i < n; i++) {
        int acc = 0;
        for (int j = 0; j < rd; j++)
            acc += in[j] * m[j * n + i];
        out[i] = acc;
    }
    for (int i = 0; i < n; i++)
        v[i * stride] = out[i];
}

static void inverse_transform(int *c, int w, int h, int trh, int trv, int nzw, int nzh, int range, int bd)
{
    /* DC-only shortcut of the DCT2 x DCT2 cells (vvcdsp.c:101-108, :125-131): inputs other than c[0] are not read */
    if (trh == 0 && trv == 0 && nzw == 1 && nzh == 1 && (w == h || w == 1 || h == 1)) {
        int dc;
        if (w > 1 && h > 1) {
            const int s2 = 5 + range - bd;
            const int t = (c[0] * 64 + 64) >> 7;
            dc = (t * 64 + (1 << (s2 - 1))) >> s2;
        } else {
            const int s = 6 + range - bd;
            dc = (c[0] * 64 + (1 << (s - 1))) >> s;
        }
        for (int i = 0; i < w * h; i++)
            c[i] = dc;
        return;
    }
    if (w > 1 && h > 1) {
        const int s2 = 5 + range - bd;
        for (int x = 0; x < nzw; x++)
            inv_1d(c + x, w, trv, h, nzh);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                c[y * w + x] = x < nzw ? o_clip_sbits((c[y * w + x] + 64) >> 7, range) : 0;
        for (int y = 0; y < h; y++)
            inv_1d(c + y * w, 1, trh, w, nzw);
        for (int i = 0; i < w * h; i++)
            c[i] = (c[i] + (1 << (s2 - 1))) >> s2;
    } else {
        const int s = 6 + range - bd;
        if (w > 1) inv_1d(c, 1, trh, w, nzw);
        else       inv_1d(c, 1, trv, h, nzh);
        for (int i = 0; i < w * h; i++)
            c[i] = (c[i] + (1 << (s - 1))) >> s;
    }
}

/* 4x4 up-right diagonal scan (ff_vvc_diag_scan_x/y[2][2]): position k -> (x, y) */
static const uint8_t k_diag4_x[16] = { 0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 1, 2, 3, 2, 3, 3 };
static const uint8_t k_diag4_y[16] = { 0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 3, 2, 1, 3, 2, 3 };

static void inverse_lfnst(int *c, int w, int h, int lfnst, int range, int *nzw, int *nzh)
{
    const int idx = lfnst & 3, set = (lfnst >> 2) & 3, transpose = (lfnst >> 4) & 1;
    const int n_in = (lfnst >> 5) & 1 ? 8 : 16;
    const int big = w >= 8 && h >= 8, n_out = big ? 48 : 16, side = big ? 8 : 4;
    const int8_t *m = big ? &vvct_lfnst_8x8[set][idx - 1][0][0] : &vvct_lfnst_4x4[set][idx - 1][0][0];
    int u[16], v[48];
    for (int k = 0; k < n_in; k++)
        u[k] = c[w * k_diag4_y[k] + k_diag4_x[k]];
    for (int j = 0; j < n_out; j++) {
        int t = 0;
        for (int i = 0; i < n_in; i++)
            t += u[i] * m[i * n_out + j];
        v[j] = o_clip_sbits((t + 64) >> 7, range);
    }
    /* v is a raster of rows 0..3 at full width `side`, then (8x8 only) rows 4..7 of width 4 */
    for (int k = 0; k < n_out; k++) {
        const int r = k < 4 * side ? k / side : 4 + (k - 4 * side) / 4;
        const int q = k < 4 * side ? k % side : (k - 4 * side) % 4;
        if (transpose) c[q * w + r] = v[k];
        else           c[r * w + q] = v[k];
    }
    *nzw = *nzh = side;
}

static void bdpcm(int *c, int w, int h, int vertical, int range)
{
    if (vertical) {
        for (int y = 1; y < h; y++)
            for (int x = 0; x < w; x++)
                c[y * w + x] = o_clip_sbits(c[y * w + x] + c[(y - 1) * w + x], range);
    } else {
        for (int y = 0; y < h; y++)
            for (int x = 1; x < w; x++)
                c[y * w + x] = o_clip_sbits(c[y * w + x] + c[y * w + x - 1], range);
    }
}

int vvco_lfnst_tr_set(int pred_mode_intra)
{
    return pred_mode_intra < 0 ? 1 : vvct_lfnst_tr_set_index[pred_mode_intra];
}

/*
 * dequant() of one TB, in place on the dense block (libavcodec/vvc/vvc_intra.c:397-417):
 *   shift / rounding offset of derive_qp (:294-308), level scale of derive_scale (:311-338), the
 *   matrix entry of derive_scale_m (:341-384; flat 16 = ff_vvc_default_scale_m), scale_coeff (:387-395).
 * q->qp is tb->qp after derive_qp's offsets and clip.  Zero levels are skipped like the reference does, so
 * running over the whole block equals running over [min_scan, max_scan].
 */
static void dequant_tb(int *c, int log2_w, int log2_h, int ts, const VVCCudaTBQuant *q, const VVCCudaScalingList *sl,
                       int range, int bd)
{
    static const int level_scale[2][6] = { { 40, 45, 51, 57, 64, 72 }, { 57, 64, 72, 80, 90, 102 } };
    const int w = 1 << log2_w, h = 1 << log2_h, log_sum = log2_w + log2_h;
    const int rect = ts ? 0 : (log_sum & 1);
    const int shift = ts ? 10 : bd + rect + log_sum / 2 + 10 - range + q->dep_quant;
    const int add = (1 << shift) >> 1;
    const int qp = q->qp + (q->dep_quant && !ts);
    const int scale = level_scale[rect][qp % 6] << (qp / 6);
    const int id = q->sl_id - 1;
    const int lms = id < 2 ? 1 : id < 8 ? 2 : 3;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int m = 16;
            if (!c[y * w + x])
                continue;
            if (q->sl_id && sl) {
                m = sl->matrix_rec[id][(((y << lms) >> log2_h) << lms) + ((x << lms) >> log2_w)];
                if (id >= 14 && !x && !y)
                    m = sl->dc_rec[id - 14];
            }
            c[y * w + x] = o_clip_sbits((c[y * w + x] * scale * m + add) >> shift, range);   /* wraps: -fwrapv */
        }
}

/* one residual through lmcs_scale_chroma (libavcodec/vvc/vvc_intra_template.c:431-448) */
static inline int lmcs_scale_one(int res, int scale, int bd)
{
    const int c = o_clip3(res, -(1 << bd), (1 << bd) - 1);
    return c > 0 ? (c * scale + (1 << 10)) >> 11 : -((-c * scale + (1 << 10)) >> 11);
}

/* Residual stage on either coefficient layout, with optional dequantisation; order of itransform()
 * (vvc_intra.c:453-470): transform_bdpcm -> dequant -> LFNST -> inverse transform -> add_residual(_joint). */
void vvco_itx_frame_q(const VVCCudaFrame *f, const VVCCudaCoeffs *co, const VVCCudaTB *tbs, int n_tbs, int range)
{
    const int bd = f->bit_depth;
    int tmp[64 * 64];
    for (int i = 0; i < n_tbs; i++) {
        const VVCCudaTB *tb = &tbs[i];
        const int w = 1 << tb->log2_w, h = 1 << tb->log2_h;
        const int ts = !!(tb->flags & VVC_CUDA_TB_TS);
        int nzw = tb->nzw, nzh = tb->nzh;
        int *c = tmp;
        if (co->format == VVC_CUDA_COEFF_WINDOW16) {
            const int16_t *src = (const int16_t *)co->data + tb->coeff_offset;
            memset(c, 0, sizeof(int) * w * h);
            for (int y = 0; y < nzh && y < h; y++)
                for (int x = 0; x < nzw && x < w; x++)
                    c[y * w + x] = src[y * nzw + x];
        } else {
            int *src = (int32_t *)co->data + tb->coeff_offset;
            if (tb->flags & VVC_CUDA_TB_STORE_RESIDUAL)
                c = src;
            else
                memcpy(c, src, sizeof(int) * w * h);
        }
        if (tb->flags & (VVC_CUDA_TB_BDPCM | VVC_CUDA_TB_BDPCM_VERT))
            bdpcm(c, w, h, !!(tb->flags & VVC_CUDA_TB_BDPCM_VERT), range);
        if (co->quant)
            dequant_tb(c, tb->log2_w, tb->log2_h, ts, &co->quant[i], co->scaling, range, bd);
        if (!ts) {
            if (tb->lfnst)
                inverse_lfnst(c, w, h, tb->lfnst, range, &nzw, &nzh);
            inverse_transform(c, w, h, tb->trh, tb->trv, nzw, nzh, range, bd);
        }
        if (tb->flags & VVC_CUDA_TB_STORE_RESIDUAL)
            continue;
        {
            /* LMCS chroma residual scaling between the transform and add_residual (itransform, vvc_intra.c:449-472) */
            const int scale = !tb->chroma_scale ? 0 : co->lmcs_scales ? co->lmcs_scales[tb->chroma_scale - 1] : tb->chroma_scale;
            const OPlane pl = o_plane(f, tb->c_idx, tb->pic);
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) {
                    pel *d = &pl.p[(tb->y0 + y) * pl.pitch + tb->x0 + x];
                    *d = (pel)o_clip_pel(*d + (scale ? lmcs_scale_one(c[y * w + x], scale, bd) : c[y * w + x]), bd);
                }
            if (tb->flags & VVC_CUDA_TB_JOINT) {
                /* add_residual_for_joint_coding_chroma (:166-186): derive the second plane's residual, then scale it */
                const OPlane pj = o_plane(f, tb->joint_c_idx, tb->pic);
                for (int y = 0; y < h; y++)
                    for (int x = 0; x < w; x++) {
                        pel *d = &pj.p[(tb->y0 + y) * pj.pitch + tb->x0 + x];
                        const int r = (c[y * w + x] * tb->joint_sign) >> tb->joint_shift;
                        *d = (pel)o_clip_pel(*d + (scale ? lmcs_scale_one(r, scale, bd) : r), bd);
                    }
            }
        }
    }
}

void vvco_itx_frame(const VVCCudaFrame *f, int32_t *coeffs, const VVCCudaTB *tbs, int n_tbs, int range)
{
    VVCCudaCoeffs co;
    memset(&co, 0, sizeof(co));
    co.data = coeffs;
    co.format = VVC_CUDA_COEFF_DENSE32;
    vvco_itx_frame_q(f, &co, tbs, n_tbs, range);
}

/* dequant() alone on a dense block (pinned against the reference's static dequant() through
 * oracle/refbuild/ref_glue_dequant.c) */
void vvco_dequant_tb(int32_t *c, int log2_w, int log2_h, int ts, const VVCCudaTBQuant *q, const VVCCudaScalingList *sl,
                     int range, int bd)
{
    dequant_tb(c, log2_w, log2_h, ts, q, sl, range, bd);
}
