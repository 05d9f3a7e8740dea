// Residual stage for sm_100a: inverse LFNST + inverse transform + add_residual over a TB list.
//
// Replaces the per-TB tail of itransform() (libavcodec/vvc/vvc_intra.c:432-478):
//   itx.transform_bdpcm                     libavcodec/vvc/vvcdsp_template.c:76-95
//   ilfnst_transform / ff_vvc_inv_lfnst_1d  libavcodec/vvc/vvc_intra.c:65-127, vvc_itx_1d.c:708-721
//   itx.itx[trh][trv][log2w][log2h]         libavcodec/vvc/vvcdsp.c:67-195 (itx_2d / itx_1d / scale_clip),
//                                           1-D kernels libavcodec/vvc/vvc_itx_1d.c:70-706
//   itx.add_residual / add_residual_joint   libavcodec/vvc/vvcdsp_template.c:32-63
//
// B200 design (CUDA-core integer path; the int8 tcgen05 variant is only worth keeping if ncu shows it
// beating this one): a CTA takes TBS_PER_CTA consecutive descriptors.  Blocks of up
// to 256 samples are transformed by single warps (one TB per warp at a time, warp-level sync only),
// larger ones by the whole CTA.  Only the non-zero window of the coefficients is read from HBM
// (rows the reference's butterflies ignore are never touched), both 1-D passes run out of shared
// memory as exact int32 matrix products with 4-output register strips (one packed int8x4 matrix
// load per 4 MACs), and the epilogue adds the residual straight into the picture - the residual
// never goes back to HBM (8 B/sample -> <= 4*nz/area + 4 B/sample).
#include "common.cuh"
#include "coeff_src.cuh"
#include "tables.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int TBS_PER_CTA = 8;
constexpr int SMALL_AREA = 256;
constexpr int SMALL_BUF = 16 * 17 + 16;      // ints per small buffer (pitch w + 1)
constexpr int LARGE_BUF = 64 * 65;

struct ItxK {
    pel       *plane[3];
    int        pitch[3];
    long long  bstride[3];
    CoefSrc    src;
    int32_t   *store;                    // DENSE32 buffer for VVC_CUDA_TB_STORE_RESIDUAL blocks (NULL otherwise)
    const VVCCudaTB *tbs;
    int        n_tbs, range, bd;
    const uint32_t *list, *list_count;   // optional: process tbs[list[0 .. *list_count)] instead of tbs[0 .. n_tbs)
};

__device__ __forceinline__ const int8_t *tx_matrix(int type, int n)
{
    if (type == 0) {
        switch (n) {
        case 2:  return &vvct_dct2_2[0][0];   case 4:  return &vvct_dct2_4[0][0];
        case 8:  return &vvct_dct2_8[0][0];   case 16: return &vvct_dct2_16[0][0];
        case 32: return &vvct_dct2_32[0][0];  default: return &vvct_dct2_64[0][0];
        }
    }
    if (type == 1) {
        switch (n) {
        case 4:  return &vvct_dst7_4[0][0];   case 8:  return &vvct_dst7_8[0][0];
        case 16: return &vvct_dst7_16[0][0];  default: return &vvct_dst7_32[0][0];
        }
    }
    switch (n) {
    case 4:  return &vvct_dct8_4[0][0];   case 8:  return &vvct_dct8_8[0][0];
    case 16: return &vvct_dct8_16[0][0];  default: return &vvct_dct8_32[0][0];
    }
}

// inputs the reference's 1-D transform reads for a declared nz (zero-out guards G2..G16, vvc_itx_1d.c:64-67)
__device__ __forceinline__ int inputs_read(int type, int n, int nz)
{
    if (type != 0)
        return nz;
    const int r = nz <= 2 ? 2 : nz <= 4 ? 4 : nz <= 8 ? 8 : nz <= 16 ? 16 : 32;
    return min(r, min(n, 32));
}

template <int NT> __device__ __forceinline__ void group_sync()
{
    if (NT == 32) __syncwarp();
    else          __syncthreads();
}

// One 1-D pass over `lines` lines.  VERT: line = column, transform along rows.  out[i] = sum_j in[j] * M[j][i].
// post_shift < 0: mid-stage rounding (x + 64) >> 7 with clip to the transform range; else (x + rnd) >> post_shift.
template <int NT, bool VERT>
__device__ __forceinline__ void run_pass(const int *src, int *dst, int pitch, int type, int n, int nz, int lines,
                                         int post_shift, int range, int t)
{
    const int8_t *M = tx_matrix(type, n);
    const int rd = inputs_read(type, n, nz);
    if (n >= 4) {
        const int strips = n >> 2;
        for (int item = t; item < lines * strips; item += NT) {
            const int line = item % lines, o0 = (item / lines) << 2;
            int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            const int *in = VERT ? src + line : src + line * pitch;
            const int istep = VERT ? pitch : 1;
            for (int j = 0; j < rd; j++) {
                const int v = in[j * istep];
                const int m4 = __ldg(reinterpret_cast<const int *>(M + j * n + o0));
                a0 += v * (int)(int8_t)(m4);
                a1 += v * (int)(int8_t)(m4 >> 8);
                a2 += v * (int)(int8_t)(m4 >> 16);
                a3 += v * (m4 >> 24);
            }
            int r[4] = { a0, a1, a2, a3 };
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int v = post_shift < 0 ? d_clip_sbits((r[q] + 64) >> 7, range)
                                             : (r[q] + (1 << (post_shift - 1))) >> post_shift;
                if (VERT) dst[(o0 + q) * pitch + line] = v;
                else      dst[line * pitch + o0 + q] = v;
            }
        }
    } else {        // n == 2
        for (int item = t; item < lines * n; item += NT) {
            const int line = item % lines, o = item / lines;
            const int *in = VERT ? src + line : src + line * pitch;
            const int istep = VERT ? pitch : 1;
            int acc = 0;
            for (int j = 0; j < rd; j++)
                acc += in[j * istep] * (int)M[j * n + o];
            const int v = post_shift < 0 ? d_clip_sbits((acc + 64) >> 7, range)
                                         : (acc + (1 << (post_shift - 1))) >> post_shift;
            if (VERT) dst[o * pitch + line] = v;
            else      dst[line * pitch + o] = v;
        }
    }
}

__constant__ uint8_t c_diag4_x[16] = { 0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 1, 2, 3, 2, 3, 3 };
__constant__ uint8_t c_diag4_y[16] = { 0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 3, 2, 1, 3, 2, 3 };

template <int NT, int MODE>
__device__ void process_tb(const ItxK &p, const VVCCudaTB &tb, int ti, int *sC, int *sM, int t)
{
    const int w = 1 << tb.log2_w, h = 1 << tb.log2_h, pitch = w + 1;
    const int flags = tb.flags;
    const bool ts = flags & VVC_CUDA_TB_TS;
    const bool pcm = flags & (VVC_CUDA_TB_BDPCM | VVC_CUDA_TB_BDPCM_VERT);
    int nzw = tb.nzw, nzh = tb.nzh;
    const int lf_side = (w >= 8 && h >= 8) ? 8 : 4;
    const TbCoef tc = tb_coef<MODE>(p.src, ti, tb.coeff_offset, tb.log2_w, tb.log2_h, nzw, nzh, ts);

    // ---- load the window of coefficients that will be read ----
    int LR, LC;
    if (ts || pcm)            { LR = h; LC = w; }
    else if (tb.lfnst)        { LR = LC = lf_side; }
    else if (w > 1 && h > 1)  { LR = inputs_read(tb.trv, h, nzh); LC = nzw; }
    else if (w > 1)           { LR = 1; LC = inputs_read(tb.trh, w, nzw); }
    else                      { LR = inputs_read(tb.trv, h, nzh); LC = 1; }
    for (int i = t; i < LR * LC; i += NT) {
        const int y = i / LC, x = i - y * LC;
        // BDPCM accumulates quantised levels; they are dequantised afterwards (vvc_intra.c:453-455)
        sC[y * pitch + x] = pcm ? coef_raw<MODE>(tc, y, x) : coef_load<MODE>(tc, y, x);
    }
    group_sync<NT>();

    // ---- BDPCM accumulate with clipping (sequential along the accumulation direction) ----
    if (pcm) {
        if (flags & VVC_CUDA_TB_BDPCM_VERT) {
            for (int x = t; x < w; x += NT)
                for (int y = 1; y < h; y++)
                    sC[y * pitch + x] = d_clip_sbits(sC[y * pitch + x] + sC[(y - 1) * pitch + x], p.range);
        } else {
            for (int y = t; y < h; y += NT)
                for (int x = 1; x < w; x++)
                    sC[y * pitch + x] = d_clip_sbits(sC[y * pitch + x] + sC[y * pitch + x - 1], p.range);
        }
        group_sync<NT>();
        if (MODE & 2) {
            for (int i = t; i < w * h; i += NT) {
                const int y = i / w, x = i - y * w;
                sC[y * pitch + x] = coef_dequant<MODE>(tc, sC[y * pitch + x], y, x);
            }
            group_sync<NT>();
        }
    }

    if (!ts) {
        // ---- inverse LFNST: 8/16 inputs in 4x4 diagonal order -> 16/48 outputs ----
        if (tb.lfnst) {
            const int idx = tb.lfnst & 3, set = (tb.lfnst >> 2) & 3;
            const bool transpose = (tb.lfnst >> 4) & 1;
            const int n_in = ((tb.lfnst >> 5) & 1) ? 8 : 16;
            const int n_out = lf_side == 8 ? 48 : 16;
            const int8_t *M = lf_side == 8 ? &vvct_lfnst_8x8[set][idx - 1][0][0] : &vvct_lfnst_4x4[set][idx - 1][0][0];
            int v[2];
            int cnt = 0;
            for (int j = t; j < n_out; j += NT, cnt++) {
                int acc = 0;
                for (int i = 0; i < n_in; i++)
                    acc += sC[c_diag4_y[i] * pitch + c_diag4_x[i]] * (int)M[i * n_out + j];
                v[cnt] = d_clip_sbits((acc + 64) >> 7, p.range);
            }
            group_sync<NT>();
            cnt = 0;
            for (int j = t; j < n_out; j += NT, cnt++) {
                const int r = j < 4 * lf_side ? j / lf_side : 4 + ((j - 4 * lf_side) >> 2);
                const int q = j < 4 * lf_side ? j % lf_side : (j - 4 * lf_side) & 3;
                if (transpose) sC[q * pitch + r] = v[cnt];
                else           sC[r * pitch + q] = v[cnt];
            }
            nzw = nzh = lf_side;
            group_sync<NT>();
        }
        // ---- inverse transform ----
        if (tb.trh == 0 && tb.trv == 0 && nzw == 1 && nzh == 1 && (w == h || w == 1 || h == 1)) {
            // DC-only shortcut of the DCT2 x DCT2 cells (vvcdsp.c:101-108, :125-131)
            const int c0 = sC[0];
            int dc;
            if (w > 1 && h > 1) {
                const int s2 = 5 + p.range - p.bd;
                dc = ((((c0 * 64 + 64) >> 7) * 64) + (1 << (s2 - 1))) >> s2;
            } else {
                const int s = 6 + p.range - p.bd;
                dc = (c0 * 64 + (1 << (s - 1))) >> s;
            }
            group_sync<NT>();
            for (int i = t; i < w * h; i += NT) {
                const int y = i / w, x = i - y * w;
                sC[y * pitch + x] = dc;
            }
        } else if (w > 1 && h > 1) {
            run_pass<NT, true>(sC, sM, pitch, tb.trv, h, nzh, nzw, -1, p.range, t);
            const int rd_h = inputs_read(tb.trh, w, nzw);
            for (int i = t; i < h * (rd_h - nzw); i += NT) {     // columns >= nzw are zero (scale_clip memset)
                const int y = i / (rd_h - nzw), x = nzw + i - y * (rd_h - nzw);
                sM[y * pitch + x] = 0;
            }
            group_sync<NT>();
            run_pass<NT, false>(sM, sC, pitch, tb.trh, w, nzw, h, 5 + p.range - p.bd, p.range, t);
        } else {
            if (w > 1) run_pass<NT, false>(sC, sM, pitch, tb.trh, w, nzw, 1, 6 + p.range - p.bd, p.range, t);
            else       run_pass<NT, true>(sC, sM, pitch, tb.trv, h, nzh, 1, 6 + p.range - p.bd, p.range, t);
            group_sync<NT>();
            for (int i = t; i < w * h; i += NT) {
                const int y = i / w, x = i - y * w;
                sC[y * pitch + x] = sM[y * pitch + x];
            }
        }
        group_sync<NT>();
    }

    // ---- epilogue ----
    if (flags & VVC_CUDA_TB_STORE_RESIDUAL) {
        int32_t *out = p.store + tb.coeff_offset;
        for (int i = t; i < w * h; i += NT) {
            const int y = i / w, x = i - y * w;
            out[i] = sC[y * pitch + x];
        }
    } else {
        const int planes = (flags & VVC_CUDA_TB_JOINT) ? 2 : 1;
        const int cscale = tb_chroma_scale(p.src, tb.chroma_scale);
        // LMCS chroma residual scaling sits between the transform and add_residual (itransform, vvc_intra.c:468-475; the
        // second plane of a joint block is derived first and scaled afterwards, :179-183)
        auto res = [&](int r, int sign, int shift) -> int {
            const int v = (r * sign) >> shift;
            return cscale ? d_lmcs_scale(v, cscale, p.bd) : v;
        };
        for (int pl = 0; pl < planes; pl++) {
            const int c = pl ? tb.joint_c_idx : tb.c_idx;
            pel *base = p.plane[c] + tb.pic * p.bstride[c] + (long long)tb.y0 * p.pitch[c] + tb.x0;
            const int sign = pl ? tb.joint_sign : 1, shift = pl ? tb.joint_shift : 0;
            if (w >= 4 && !(tb.x0 & 3)) {
                const int q4 = w >> 2;
                for (int i = t; i < h * q4; i += NT) {
                    const int y = i / q4, x = (i - y * q4) << 2;
                    uint2 *d = reinterpret_cast<uint2 *>(base + (long long)y * p.pitch[c] + x);
                    const uint2 cur = *d;
                    const int *r = &sC[y * pitch + x];
                    const int o0 = d_clip_pel((int)(cur.x & 0xffff) + res(r[0], sign, shift), p.bd);
                    const int o1 = d_clip_pel((int)(cur.x >> 16)    + res(r[1], sign, shift), p.bd);
                    const int o2 = d_clip_pel((int)(cur.y & 0xffff) + res(r[2], sign, shift), p.bd);
                    const int o3 = d_clip_pel((int)(cur.y >> 16)    + res(r[3], sign, shift), p.bd);
                    *d = make_uint2(o0 | (o1 << 16), o2 | (o3 << 16));
                }
            } else {
                for (int i = t; i < h * w; i += NT) {
                    const int y = i / w, x = i - y * w;
                    pel *d = base + (long long)y * p.pitch[c] + x;
                    *d = (pel)d_clip_pel(*d + res(sC[y * pitch + x], sign, shift), p.bd);
                }
            }
        }
    }
    group_sync<NT>();
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) itx_kernel(const ItxK p)
{
    __shared__ alignas(16) int s_buf[2 * LARGE_BUF];
    const int total = p.list ? (int)*p.list_count : p.n_tbs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int first = blockIdx.x * TBS_PER_CTA; first < total; first += gridDim.x * TBS_PER_CTA) {
        const int count = min(TBS_PER_CTA, total - first);
        __syncthreads();
        // phase A: small blocks, one warp each
        {
            int *sC = s_buf + warp * 2 * SMALL_BUF, *sM = sC + SMALL_BUF;
            for (int i = warp; i < count; i += kThreads / 32) {
                const int ti = p.list ? p.list[first + i] : first + i;
                const VVCCudaTB tb = p.tbs[ti];
                if ((1 << (tb.log2_w + tb.log2_h)) <= SMALL_AREA && tb.log2_w <= 4 && tb.log2_h <= 4)
                    process_tb<32, MODE>(p, tb, ti, sC, sM, lane);
            }
        }
        __syncthreads();
        // phase B: large blocks, whole CTA
        for (int i = 0; i < count; i++) {
            const int ti = p.list ? p.list[first + i] : first + i;
            const VVCCudaTB tb = p.tbs[ti];
            if (!((1 << (tb.log2_w + tb.log2_h)) <= SMALL_AREA && tb.log2_w <= 4 && tb.log2_h <= 4))
                process_tb<kThreads, MODE>(p, tb, ti, s_buf, s_buf + LARGE_BUF, threadIdx.x);
        }
    }
}

}  // namespace

static void launch_generic(const ItxK &p, int mode, int grid, cudaStream_t st)
{
    switch (mode) {
    case 0:  itx_kernel<0><<<grid, kThreads, 0, st>>>(p); break;
    case 1:  itx_kernel<1><<<grid, kThreads, 0, st>>>(p); break;
    case 2:  itx_kernel<2><<<grid, kThreads, 0, st>>>(p); break;
    default: itx_kernel<3><<<grid, kThreads, 0, st>>>(p); break;
    }
}

extern "C" int vvc_cuda_itx_frame_q(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaCoeffs *co,
                                    const VVCCudaTB *tbs, int n_tbs, int log2_transform_range)
{
    if (ctx->err)
        return ctx->err;
    if (!frame || !co || !co->data || !tbs || n_tbs < 0)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: null argument");
    if (frame->bit_depth != 10 && frame->bit_depth != 12)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: bit depth %d not accelerated", frame->bit_depth);
    if (log2_transform_range < 15 || log2_transform_range > 20)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: log2_transform_range %d out of range", log2_transform_range);
    if (co->format != VVC_CUDA_COEFF_DENSE32 && co->format != VVC_CUDA_COEFF_WINDOW16)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: unknown coefficient layout %d", co->format);
    if (co->format == VVC_CUDA_COEFF_WINDOW16 && log2_transform_range != 15)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: the 16-bit window layout needs log2_transform_range 15");
    if (!frame_vec_ok(frame))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: planes and strides must be 16-byte aligned");
    if (frame->batch > 256)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: a ring of %d pictures cannot be addressed by the 8-bit picture field of the records", frame->batch);
    if (!n_tbs)
        return VVC_CUDA_OK;
    const int mode = coef_mode(co);
    ItxK p;
    for (int c = 0; c < 3; c++) {
        p.plane[c] = (pel *)frame->data[c];
        p.pitch[c] = (int)(frame->stride[c] / 2);
        p.bstride[c] = frame->batch_stride[c] / 2;
    }
    p.src.dense = (mode & 1) ? nullptr : (const int32_t *)co->data;
    p.src.window = (mode & 1) ? (const int16_t *)co->data : nullptr;
    p.src.quant = co->quant; p.src.scaling = co->scaling; p.src.lmcs_scales = co->lmcs_scales;
    p.src.range = log2_transform_range; p.src.bd = frame->bit_depth;
    p.store = (mode & 1) ? nullptr : (int32_t *)co->data;
    p.tbs = tbs; p.n_tbs = n_tbs; p.range = log2_transform_range; p.bd = frame->bit_depth;
    p.list = p.list_count = NULL;
    if (p.bd == 10 && p.range == 15 && !ctx->force_generic) {
        // common kinds: warp-per-TB kernel (itx_warp.cu); it lists what it leaves (transform skip, BDPCM, 1-D blocks)
        uint32_t *scratch = (uint32_t *)vvc_ctx_scratch(ctx, 2, (16 + 6 * (size_t)n_tbs) * sizeof(uint32_t));
        if (!scratch)
            return ctx->err;
        if (vvc_itx_launch_warp(ctx, frame, co, tbs, n_tbs, scratch, &p.list, &p.list_count))
            return ctx->err;
        const int ctas = ceil_div(n_tbs, TBS_PER_CTA);
        launch_generic(p, mode, ctas < 148 * 6 ? ctas : 148 * 6, ctx->side[0]);      // beside the warp kernel: disjoint blocks
        VVC_LAUNCHED(ctx);
        return vvc_ctx_join(ctx, 1) ? ctx->err : VVC_CUDA_OK;
    }
    launch_generic(p, mode, ceil_div(n_tbs, TBS_PER_CTA), ctx->stream);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

// One launch of the generic kernel over a short list: the per-wavefront residual step of vvc_cuda_intra_recon_frame,
// where the sort + three-kernel structure above would cost more launches than the blocks are worth.
int vvc_itx_launch_short(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaCoeffs *co, const VVCCudaTB *tbs, int n_tbs, int range)
{
    if (!frame || !co || !co->data || !tbs || n_tbs <= 0)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: null argument");
    if ((frame->bit_depth != 10 && frame->bit_depth != 12) || range < 15 || range > 20 ||
        (co->format != VVC_CUDA_COEFF_DENSE32 && co->format != VVC_CUDA_COEFF_WINDOW16) || (co->format == VVC_CUDA_COEFF_WINDOW16 && range != 15) ||
        !frame_vec_ok(frame))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: unsupported picture or coefficient format");
    const int mode = coef_mode(co);
    ItxK p;
    for (int c = 0; c < 3; c++) {
        p.plane[c] = (pel *)frame->data[c];
        p.pitch[c] = (int)(frame->stride[c] / 2);
        p.bstride[c] = frame->batch_stride[c] / 2;
    }
    p.src.dense = (mode & 1) ? nullptr : (const int32_t *)co->data;
    p.src.window = (mode & 1) ? (const int16_t *)co->data : nullptr;
    p.src.quant = co->quant; p.src.scaling = co->scaling; p.src.lmcs_scales = co->lmcs_scales;
    p.src.range = range; p.src.bd = frame->bit_depth;
    p.store = (mode & 1) ? nullptr : (int32_t *)co->data;
    p.tbs = tbs; p.n_tbs = n_tbs; p.range = range; p.bd = frame->bit_depth;
    p.list = p.list_count = NULL;
    launch_generic(p, mode, ceil_div(n_tbs, TBS_PER_CTA), ctx->stream);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_itx_frame(VVCCudaCtx *ctx, const VVCCudaFrame *frame, int32_t *coeffs,
                                  const VVCCudaTB *tbs, int n_tbs, int log2_transform_range)
{
    if (ctx->err)
        return ctx->err;
    if (!coeffs)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: null argument");
    VVCCudaCoeffs co;
    memset(&co, 0, sizeof(co));
    co.data = coeffs; co.format = VVC_CUDA_COEFF_DENSE32;
    return vvc_cuda_itx_frame_q(ctx, frame, &co, tbs, n_tbs, log2_transform_range);
}

extern "C" int vvc_cuda_itx_frame_q_host(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaCoeffs *co,
                                         const VVCCudaTB *tbs, int n_tbs, int log2_transform_range)
{
    if (ctx->err)
        return ctx->err;
    if (!frame || !co || !co->data || !tbs || n_tbs < 0)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx_host: null argument");
    const bool win = co->format == VVC_CUDA_COEFF_WINDOW16;
    if (win)
        for (int i = 0; i < n_tbs; i++)
            if (tbs[i].flags & VVC_CUDA_TB_STORE_RESIDUAL)
                return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx_host: VVC_CUDA_TB_STORE_RESIDUAL needs the dense int32 layout");
    if (co->quant)                  // the lists are in host memory here: refuse ids the device would index out of range with
        for (int i = 0; i < n_tbs; i++)
            if (co->quant[i].sl_id > 28 || (co->quant[i].sl_id && !co->scaling))
                return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx_host: TB %d names scaling matrix %d (%s)", i, co->quant[i].sl_id - 1,
                                    co->scaling ? "ids are 0..27" : "no scaling list given");
    const size_t esz = win ? sizeof(int16_t) : sizeof(int32_t);
    const size_t fsz = align_up(vvc_stage_frame_size(frame), 256);
    const size_t csz = align_up(co->n * esz, 256);
    const size_t tsz = align_up((size_t)n_tbs * sizeof(VVCCudaTB), 256);
    const size_t qsz = align_up((size_t)n_tbs * sizeof(VVCCudaTBQuant), 256);
    const size_t ssz = align_up(sizeof(VVCCudaScalingList), 256);
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, fsz + csz + tsz + qsz + ssz);
    if (!base)
        return ctx->err;
    VVCCudaFrame df;
    vvc_stage_frame_layout(frame, base, &df);
    VVCCudaCoeffs dco = *co;
    dco.data = base + fsz;
    VVCCudaTB *dtb = (VVCCudaTB *)(base + fsz + csz);
    if (vvc_stage_frame_h2d(ctx, &df, frame))
        return ctx->err;
    VVC_TRY(ctx, cudaMemcpyAsync(dco.data, co->data, co->n * esz, cudaMemcpyHostToDevice, ctx->stream));
    VVC_TRY(ctx, cudaMemcpyAsync(dtb, tbs, (size_t)n_tbs * sizeof(VVCCudaTB), cudaMemcpyHostToDevice, ctx->stream));
    if (co->quant) {
        dco.quant = (const VVCCudaTBQuant *)(base + fsz + csz + tsz);
        VVC_TRY(ctx, cudaMemcpyAsync((void *)dco.quant, co->quant, (size_t)n_tbs * sizeof(VVCCudaTBQuant), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (co->scaling) {
        dco.scaling = (const VVCCudaScalingList *)(base + fsz + csz + tsz + qsz);
        VVC_TRY(ctx, cudaMemcpyAsync((void *)dco.scaling, co->scaling, sizeof(VVCCudaScalingList), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (vvc_cuda_itx_frame_q(ctx, &df, &dco, dtb, n_tbs, log2_transform_range))
        return ctx->err;
    if (vvc_stage_frame_d2h(ctx, frame, &df))
        return ctx->err;
    // blocks flagged STORE_RESIDUAL return their residual in place, like the reference's itx entries
    if (!win)
        VVC_TRY(ctx, cudaMemcpyAsync(co->data, dco.data, co->n * esz, cudaMemcpyDeviceToHost, ctx->stream));
    return vvc_cuda_sync(ctx);
}

extern "C" int vvc_cuda_itx_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *frame, int32_t *coeffs, size_t n_coeffs,
                                       const VVCCudaTB *tbs, int n_tbs, int log2_transform_range)
{
    if (ctx->err)
        return ctx->err;
    if (!coeffs)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx_host: null argument");
    VVCCudaCoeffs co;
    memset(&co, 0, sizeof(co));
    co.data = coeffs; co.n = n_coeffs; co.format = VVC_CUDA_COEFF_DENSE32;
    return vvc_cuda_itx_frame_q_host(ctx, frame, &co, tbs, n_tbs, log2_transform_range);
}
