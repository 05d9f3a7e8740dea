/*
 * TEST INFRASTRUCTURE - frame-level ALF through the UNMODIFIED reference table entries.
 *
 * Walks the picture CTB by CTB exactly like ff_vvc_alf_filter
 * (libavcodec/vvc/vvc_filter.c:1254-1319): build the padded per-CTB source with the halo rule
 * of alf_prepare_buffer (:1105-1137), then call the reference's own
 *   alf.classify -> alf.recon_coeff_and_clip -> alf.filter[LUMA]      (:1142-1186)
 *   alf.filter[CHROMA]                                               (:1196-1211)
 *   alf.filter_cc on the padded pre-ALF luma                         (:1213-1229, :1311-1315)
 * All arithmetic is the reference's; only the walk and the halo copy are ours.
 * Out of place: src is the pre-ALF picture (what the reference keeps in its saved border
 * lines), dst receives what the reference would leave in the frame.
 */
#include <stdint.h>
#include <string.h>
#include "libavcodec/vvc/vvcdsp.h"
#include "libavcodec/vvc/vvc_data.h"
#include "vvcdsp_cuda.h"

const VVCDSPContext *vvcref_dsp(int bit_depth);

#define PAD_STRIDE (128 + 32)          /* EDGE_EMU_BUFFER_STRIDE, vvc_ctu.h:49 */
#define PAD_ORIGIN (PAD_STRIDE * 8 + 8) /* ALF_PADDING_SIZE rows + cols, vvc_filter.c:1262-1263 */

typedef uint16_t pixel;

static int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

static void fill_padded(pixel *pad, const pixel *plane, ptrdiff_t pitch, int pw, int ph,
                        int x0, int y0, int w, int h, int halo, unsigned edges)
{
    const int xlo = (edges & VVC_CUDA_EDGE_LEFT)   ? x0         : 0;
    const int xhi = (edges & VVC_CUDA_EDGE_RIGHT)  ? x0 + w - 1 : pw - 1;
    const int ylo = (edges & VVC_CUDA_EDGE_TOP)    ? y0         : 0;
    const int yhi = (edges & VVC_CUDA_EDGE_BOTTOM) ? y0 + h - 1 : ph - 1;
    for (int y = -halo; y < h + halo; y++)
        for (int x = -halo; x < w + halo; x++)
            pad[PAD_ORIGIN + y * PAD_STRIDE + x] =
                plane[clampi(y0 + y, ylo, yhi) * pitch + clampi(x0 + x, xlo, xhi)];
}

void vvcref_alf_frame(const VVCCudaFrame *dstf, const VVCCudaFrame *srcf,
                      const VVCCudaALFCtb *ctbs, const VVCCudaALFSets *sets, int sets_per_frame)
{
    const VVCDSPContext *dsp = vvcref_dsp(srcf->bit_depth);
    const int ctb = 1 << srcf->ctb_log2;
    const int cols = (srcf->width + ctb - 1) >> srcf->ctb_log2;
    const int rows = (srcf->height + ctb - 1) >> srcf->ctb_log2;
    const int planes = srcf->chroma_format_idc ? 3 : 1;
    static _Thread_local pixel pad_luma[PAD_STRIDE * (128 + 16)], pad_chroma[PAD_STRIDE * (128 + 16)];
    static _Thread_local int16_t coeff[1024 * 12], clip[1024 * 12];
    static _Thread_local int class_idx[1024], transpose_idx[1024];
    static _Thread_local int gradient_tmp[66 * 66 * 4];
    const int clip_shift[4] = { 0, 3, 5, 7 };

    for (int k = 0; k < srcf->batch; k++) {
        const VVCCudaALFSets *fs = sets + (sets_per_frame ? k : 0);
        for (int cy = 0; cy < rows; cy++)
            for (int cx = 0; cx < cols; cx++) {
                const VVCCudaALFCtb *a = &ctbs[(size_t)k * cols * rows + cy * cols + cx];
                unsigned edges = a->edges | (cx == 0 ? VVC_CUDA_EDGE_LEFT : 0) | (cy == 0 ? VVC_CUDA_EDGE_TOP : 0)
                               | (cx == cols - 1 ? VVC_CUDA_EDGE_RIGHT : 0) | (cy == rows - 1 ? VVC_CUDA_EDGE_BOTTOM : 0);
                for (int c = 0; c < planes; c++) {
                    const int hs = c ? srcf->hshift : 0, vs = c ? srcf->vshift : 0;
                    const int pw = srcf->width >> hs, ph = srcf->height >> vs;
                    const int x = (cx * ctb) >> hs, y = (cy * ctb) >> vs;
                    const int w = (ctb >> hs) < pw - x ? (ctb >> hs) : pw - x;
                    const int h = (ctb >> vs) < ph - y ? (ctb >> vs) : ph - y;
                    const ptrdiff_t sp = srcf->stride[c] / 2, dp = dstf->stride[c] / 2;
                    const pixel *s = (const pixel *)((const uint8_t *)srcf->data[c] + k * srcf->batch_stride[c]);
                    pixel *d = (pixel *)((uint8_t *)dstf->data[c] + k * dstf->batch_stride[c]);
                    pixel *pad = c ? pad_chroma : pad_luma;
                    uint8_t *dst8 = (uint8_t *)(d + y * dp + x);

                    /* the reference leaves the frame untouched when the flag is off */
                    for (int r = 0; r < h; r++)
                        memcpy(d + (y + r) * dp + x, s + (y + r) * sp + x, w * sizeof(pixel));

                    if (a->ctb_flag[c] || (!c && (a->cc_idc[0] || a->cc_idc[1])))
                        fill_padded(pad, s, sp, pw, ph, x, y, w, h, c ? 2 : 3, edges);

                    if (a->ctb_flag[c]) {
                        if (!c) {
                            const int vb_pos = ctb - 4;
                            const int size = w * h / 16;
                            const int16_t *coeff_set;
                            const uint8_t *clip_idx_set, *class_to_filt;
                            uint8_t fixed_clip_set[25][12] = { 0 };
                            if (a->filt_set_idx_y < 16) {
                                coeff_set     = &ff_vvc_alf_fix_filt_coeff[0][0];
                                clip_idx_set  = &fixed_clip_set[0][0];
                                class_to_filt = ff_vvc_alf_class_to_filt_map[a->filt_set_idx_y];
                            } else {
                                coeff_set     = &fs->luma_coeff[a->filt_set_idx_y - 16][0][0];
                                clip_idx_set  = &fs->luma_clip_idx[a->filt_set_idx_y - 16][0][0];
                                class_to_filt = ff_vvc_alf_aps_class_to_filt_map;
                            }
                            dsp->alf.classify(class_idx, transpose_idx, (uint8_t *)(pad + PAD_ORIGIN), PAD_STRIDE * 2,
                                              w, h, vb_pos, gradient_tmp);
                            dsp->alf.recon_coeff_and_clip(coeff, clip, class_idx, transpose_idx, size,
                                                          coeff_set, clip_idx_set, class_to_filt);
                            dsp->alf.filter[0](dst8, dstf->stride[c], (uint8_t *)(pad + PAD_ORIGIN), PAD_STRIDE * 2,
                                               w, h, coeff, clip, vb_pos);
                        } else {
                            const int alt = a->chroma_alt_idx[c - 1];
                            int16_t cclip[6];
                            for (int i = 0; i < 6; i++)
                                cclip[i] = 1 << (srcf->bit_depth - clip_shift[fs->chroma_clip_idx[alt][i]]);
                            dsp->alf.filter[1](dst8, dstf->stride[c], (uint8_t *)(pad + PAD_ORIGIN), PAD_STRIDE * 2,
                                               w, h, fs->chroma_coeff[alt], cclip, (ctb >> vs) - 2);
                        }
                    }
                    if (c && a->cc_idc[c - 1])
                        dsp->alf.filter_cc(dst8, dstf->stride[c], (uint8_t *)(pad_luma + PAD_ORIGIN), PAD_STRIDE * 2,
                                           w, h, hs, vs, fs->cc_coeff[c - 1][a->cc_idc[c - 1] - 1], ctb - 4);
                }
            }
    }
}
