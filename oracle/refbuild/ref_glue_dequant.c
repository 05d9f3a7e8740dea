/*
 * TEST INFRASTRUCTURE - the reference's own dequant() (libavcodec/vvc/vvc_intra.c:397-417, with
 * derive_qp :277-308, derive_scale :311-338, derive_scale_m :341-384, scale_coeff :387-395).
 * Those functions are static, so this glue compiles the UNMODIFIED vvc_intra.c a second time
 * (textually included where it lies, its exported names prefixed so they do not collide with the
 * first copy in libvvcref.so) and calls dequant() with the smallest decoder contexts that carry
 * the fields it reads.
 */
#define ff_vvc_reconstruct              vvcdq_ff_vvc_reconstruct
#define ff_vvc_get_mip_size_id          vvcdq_ff_vvc_get_mip_size_id
#define ff_vvc_nscale_derive            vvcdq_ff_vvc_nscale_derive
#define ff_vvc_need_pdpc                vvcdq_ff_vvc_need_pdpc
#define ff_vvc_get_top_available        vvcdq_ff_vvc_get_top_available
#define ff_vvc_get_left_available       vvcdq_ff_vvc_get_left_available
#define ff_vvc_ref_filter_flag_derive   vvcdq_ff_vvc_ref_filter_flag_derive
#define ff_vvc_intra_pred_angle_derive  vvcdq_ff_vvc_intra_pred_angle_derive
#define ff_vvc_intra_inv_angle_derive   vvcdq_ff_vvc_intra_inv_angle_derive
#define ff_vvc_wide_angle_mode_mapping  vvcdq_ff_vvc_wide_angle_mode_mapping
#include "libavcodec/vvc/vvc_intra.c"

#include <stdlib.h>
#include <string.h>

/*
 * One TB through dequant().  coeffs: dense int[h][w] of quantised levels, dequantised in place.
 * Returns tb->qp as derive_qp left it (what VVCCudaTBQuant.qp carries) and the scaling matrix id
 * the reference picked through *sl_id (0 = flat, else 1 + id), recomputed with the same Table 38 walk.
 */
int vvcref_dequant_tb(int *coeffs, int log2_w, int log2_h, int c_idx, int ts,
                      int min_x, int min_y, int max_x, int max_y,
                      int cu_qp, int is_intra, int act_enabled, int apply_lfnst, int jcbcr, int jcbcr_qp,
                      int bit_depth, int log2_transform_range, int min_qp_prime_ts,
                      int dep_quant, int explicit_sl, int sl_lfnst_disabled, const uint8_t *sl_bytes)
{
    static _Thread_local VVCFrameContext *fc;
    static _Thread_local VVCLocalContext *lc;
    static _Thread_local SliceContext *sc;
    static _Thread_local CodingUnit *cu;
    static _Thread_local VVCSPS *sps;
    static _Thread_local H266RawSPS *rsps;
    static _Thread_local H266RawSliceHeader *rsh;
    static _Thread_local VVCScalingList *sl;
    TransformUnit tu;
    TransformBlock *tb = &tu.tbs[0];

    if (!fc) {
        fc = calloc(1, sizeof(*fc));   lc = calloc(1, sizeof(*lc));   sc = calloc(1, sizeof(*sc));
        cu = calloc(1, sizeof(*cu));   sps = calloc(1, sizeof(*sps)); rsps = calloc(1, sizeof(*rsps));
        rsh = calloc(1, sizeof(*rsh)); sl = calloc(1, sizeof(*sl));
        for (int i = 0; i < 64 * 64; i++)
            ff_vvc_default_scale_m[i] = 16;             /* what vvc_ps.c's init does once per process */
    }
    memset(&tu, 0, sizeof(tu));
    sps->r = rsps;
    sps->bit_depth = bit_depth;
    sps->qp_bd_offset = 6 * (bit_depth - 8);
    sps->log2_transform_range = log2_transform_range;
    rsps->sps_min_qp_prime_ts = min_qp_prime_ts;
    rsps->sps_scaling_matrix_for_lfnst_disabled_flag = sl_lfnst_disabled;
    rsps->sps_scaling_matrix_for_alternative_colour_space_disabled_flag = 0;
    rsh->sh_dep_quant_used_flag = dep_quant;
    rsh->sh_explicit_scaling_list_used_flag = explicit_sl;
    sc->sh.r = rsh;
    fc->ps.sps = sps;
    if (sl_bytes) {
        memcpy(sl->scaling_matrix_rec, sl_bytes, sizeof(sl->scaling_matrix_rec));
        memcpy(sl->scaling_matrix_dc_rec, sl_bytes + sizeof(sl->scaling_matrix_rec), sizeof(sl->scaling_matrix_dc_rec));
        fc->ps.sl = sl;
    } else {
        fc->ps.sl = NULL;
    }
    lc->fc = fc; lc->sc = sc; lc->cu = cu;
    memset(cu, 0, sizeof(*cu));
    cu->pred_mode = is_intra ? MODE_INTRA : MODE_INTER;
    cu->act_enabled_flag = act_enabled;
    cu->apply_lfnst_flag[c_idx] = apply_lfnst;
    cu->qp[LUMA] = cu->qp[CB] = cu->qp[CR] = cu_qp;
    cu->qp[JCBCR] = jcbcr_qp;
    tu.joint_cbcr_residual_flag = jcbcr;
    tu.coded_flag[CB] = tu.coded_flag[CR] = 1;
    tb->has_coeffs = 1; tb->c_idx = c_idx; tb->ts = ts;
    tb->tb_width = 1 << log2_w; tb->tb_height = 1 << log2_h;
    tb->log2_tb_width = log2_w; tb->log2_tb_height = log2_h;
    tb->min_scan_x = min_x; tb->min_scan_y = min_y; tb->max_scan_x = max_x; tb->max_scan_y = max_y;
    tb->coeffs = coeffs;
    dequant(lc, &tu, tb);
    return tb->qp;
}

/*
 * Residual stage from quantised levels in either layout, every arithmetic step by the reference:
 * itx.transform_bdpcm -> dequant() -> vvcref_itx_frame (LFNST, itx.itx[][][][], add_residual*), the order
 * of itransform() (vvc_intra.c:453-470).  VVCCudaTBQuant carries derive_qp's result and the Table 38 id;
 * the syntax-level inputs dequant() wants are chosen so that the reference derives exactly those
 * (cu->qp = qp - qp_bd_offset for luma; pred_mode from which row of Table 38 holds the id) - if a record is
 * inconsistent with the table the reference picks another matrix and the comparison fails.
 */
#include "vvcdsp_cuda.h"

const VVCDSPContext *vvcref_dsp(int bit_depth);
void vvcref_itx_frame(const VVCCudaFrame *f, int32_t *coeffs, const VVCCudaTB *tbs, int n_tbs, int range);

void vvcref_itx_frame_q(const VVCCudaFrame *f, const VVCCudaCoeffs *co, const VVCCudaTB *tbs, int n_tbs, int range)
{
    static const int intra_ids[3][6] = { { 0, 2, 8, 14, 20, 26 }, { 0, 3, 9, 15, 21, 21 }, { 0, 4, 10, 16, 22, 22 } };
    const VVCDSPContext *dsp = vvcref_dsp(f->bit_depth);
    static _Thread_local int dense[64 * 64];
    for (int i = 0; i < n_tbs; i++) {
        VVCCudaTB tb = tbs[i];
        const int w = 1 << tb.log2_w, h = 1 << tb.log2_h, ts = !!(tb.flags & VVC_CUDA_TB_TS);
        int max_x = tb.nzw - 1, max_y = tb.nzh - 1;
        int *c = dense;
        if (co->format == VVC_CUDA_COEFF_WINDOW16) {
            const int16_t *src = (const int16_t *)co->data + tb.coeff_offset;
            memset(dense, 0, sizeof(int) * w * h);
            for (int y = 0; y < tb.nzh && y < h; y++)
                for (int x = 0; x < tb.nzw && x < w; x++)
                    dense[y * w + x] = src[y * tb.nzw + x];
        } else if (tb.flags & VVC_CUDA_TB_STORE_RESIDUAL) {
            c = (int32_t *)co->data + tb.coeff_offset;
        } else {
            memcpy(dense, (int32_t *)co->data + tb.coeff_offset, sizeof(int) * w * h);
        }
        if (tb.flags & (VVC_CUDA_TB_BDPCM | VVC_CUDA_TB_BDPCM_VERT)) {
            const int vertical = !!(tb.flags & VVC_CUDA_TB_BDPCM_VERT);
            dsp->itx.transform_bdpcm(c, w, h, vertical, range);
            if (vertical) max_y = h - 1; else max_x = w - 1;           /* transform_bdpcm(), vvc_intra.c:419-430 */
        }
        if (co->quant) {
            const VVCCudaTBQuant *q = &co->quant[i];
            const int size_idx = (tb.log2_w > tb.log2_h ? tb.log2_w : tb.log2_h) - 1;
            const int is_intra = q->sl_id && size_idx >= 0 && intra_ids[tb.c_idx][size_idx] == q->sl_id - 1;
            const int qp_bd_offset = 6 * (f->bit_depth - 8);
            const int cu_qp = tb.c_idx == 0 ? q->qp - qp_bd_offset : q->qp;
            vvcref_dequant_tb(c, tb.log2_w, tb.log2_h, tb.c_idx, ts, 0, 0, max_x, max_y,
                              cu_qp, is_intra, 0, 0, 0, 0, f->bit_depth, range, 0,
                              q->dep_quant, q->sl_id != 0, 0, co->scaling ? (const uint8_t *)co->scaling : NULL);
        }
        tb.flags &= ~(VVC_CUDA_TB_BDPCM | VVC_CUDA_TB_BDPCM_VERT);
        if (c == dense) {
            tb.coeff_offset = 0;
            vvcref_itx_frame(f, dense, &tb, 1, range);
        } else {
            vvcref_itx_frame(f, (int32_t *)co->data, &tb, 1, range);
        }
    }
}
