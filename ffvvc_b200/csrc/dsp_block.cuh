// Launchers of the per-call block kernels (dsp_block.cu) used by the table shims (dsp_table.cu).  Device pointers only.
#pragma once
#include "common.cuh"

namespace vvcblk {

struct McArgs {                 // put / put_uni / put_uni_w
    const pel *src; int sstride;        // src: the block's first sample inside the staged window (pitch in samples)
    int w, h, taps, hfrac, vfrac;       // taps 8 (luma) / 4 (chroma); which passes filter
    int mode;                           // 0 put (int16 tile, pitch 128), 1 put_uni, 2 put_uni_w
    int8_t hf[8], vf[8];
    int16_t *dst16; pel *dst; int dstride;
    int denom, wx, ox;
};
struct BlendArgs {              // mode 0 avg, 1 w_avg, 2 put_gpm, 3 put_ciip
    pel *dst; int dstride; const int16_t *src0, *src1; int w, h, mode;
    int denom, w0, w1, o0, o1;          // w_avg; put_ciip: w0 = intra weight
    const uint8_t *weights; int step_x, step_y;
    const pel *inter; int istride;
};
struct ProfArgs {               // mode 0 apply_prof, 1 apply_prof_uni, 2 apply_prof_uni_w
    const int16_t *src, *dx, *dy; int mode; int16_t *dst16; pel *dst; int dstride; int denom, wx, ox;
};
struct SaoArgs { pel *dst; int dstride; const pel *src; int sstride; int w, h, edge, eo, left_class; int16_t offset_val[5]; };
struct AlfArgs { pel *dst; int dstride; const pel *src; int sstride; int w, h, chroma, vb_pos; const int16_t *filter, *clip; };
struct LfArgs {                 // one call = 8 lines along the edge
    pel *pix; long long xs, ys;         // step across the edge / along it, in samples
    int beta[4], tc[4]; uint8_t no_p[4], no_q[4], max_len_p[4], max_len_q[4];
    int param;                          // luma: hor_ctu_edge; chroma: shift
};

void mc(cudaStream_t st, int bd, const McArgs &a);
void blend(cudaStream_t st, int bd, const BlendArgs &a);
void dmvr(cudaStream_t st, int bd, int16_t *dst, const pel *src, int ss, int h, int w, int mx, int my);
void sad(cudaStream_t st, int *out, const int16_t *s0, const int16_t *s1, int dx, int dy, int bw, int bh);
void fetch(cudaStream_t st, int bd, int16_t *dst, const pel *src, int ss, int w, int h);
void prof_grad(cudaStream_t st, int16_t *gh, int16_t *gv, int gs, const int16_t *src, int ss, int w, int h, int pad);
void prof(cudaStream_t st, int bd, const ProfArgs &a);
void bdof(cudaStream_t st, int bd, pel *dst, int dstride, int16_t *s0, int16_t *s1, int w, int h);
void sao(cudaStream_t st, int bd, const SaoArgs &a);
void alf_filter(cudaStream_t st, int bd, const AlfArgs &a);
void alf_cc(cudaStream_t st, int bd, pel *dst, int dstride, const pel *luma, int ls, int w, int h, int hs, int vs, const int16_t *f, int vb_pos);
void alf_classify(cudaStream_t st, int bd, int *cls, int *tr, const pel *src, int ss, int w, int h, int vb);
void alf_recon(cudaStream_t st, int bd, int16_t *coeff, int16_t *clip, const int *cls, const int *tr, int size,
               const int16_t *coeff_set, const uint8_t *clip_idx_set, const uint8_t *class_to_filt);
void lf_luma(cudaStream_t st, int bd, const LfArgs &a);
void lf_chroma(cudaStream_t st, int bd, const LfArgs &a);
void ladf(cudaStream_t st, int *out, const pel *four);
void residual_joint(cudaStream_t st, int bd, pel *dst, int dstride, int *res, int w, int h, int c_sign, int shift, int to_buffer);

}  // namespace vvcblk
