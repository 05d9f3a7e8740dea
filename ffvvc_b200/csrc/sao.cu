// SAO stage for sm_100a: one launch per picture ring, all planes.
//
// Replaces ff_vvc_sao_filter (libavcodec/vvc/vvc_filter.c:154-298) and the table entries
//   sao_band_filter      libavcodec/h26x/h2656_sao_template.c:24-46
//   sao_edge_filter      libavcodec/h26x/h2656_sao_template.c:50-79
//   sao_edge_restore_0/1 libavcodec/h26x/h2656_sao_template.c:81-215
//
// B200 design: purely streaming.  A thread owns 8 consecutive samples (one 128-bit load/store) of
// RPT consecutive rows and keeps a rolling 3-row window in registers; the left/right neighbour
// samples of its 8-sample group come from the adjacent lanes by warp shuffle (only the two lanes
// at the ends of a warp touch memory for them).  The reference reads its neighbours from a per-CTB
// copy with a halo of saved pre-SAO lines; here the pre-SAO picture is simply read-only, which is
// the same thing.  The per-CTB restore rules are evaluated only on CTB border samples.
#include "common.cuh"

namespace {

struct SaoK {
    const pel *src[3];
    pel       *dst[3];
    int        sp[3], dp[3];
    long long  sb[3], db[3];
    int        pw[3], ph[3], hs[3], vs[3];
    int        bd, ctb_log2, ctb_cols, ctb_rows, planes;
    const VVCCudaSAOCtb *ctbs;
};

constexpr int RPT = 4;          // rows per thread
constexpr int WARPS = 8;        // warps per CTA, each on its own row group

struct Row10 { int v[10]; };    // samples x-1 .. x+8 of one row

__device__ __forceinline__ Row10 load_row(const pel *plane, int pitch, int pw, int ph, int x, int y, int lane)
{
    Row10 r;
    y = min(max(y, 0), ph - 1);
    const pel *row = plane + (long long)y * pitch;
    unsigned w[4];
    if (x + 7 < pw) {
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(row + x));
        w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const unsigned lo = x + 2 * e     < pw ? __ldg(row + x + 2 * e)     : 0;
            const unsigned hi = x + 2 * e + 1 < pw ? __ldg(row + x + 2 * e + 1) : 0;
            w[e] = lo | (hi << 16);
        }
    }
#pragma unroll
    for (int e = 0; e < 4; e++) {
        r.v[1 + 2 * e] = w[e] & 0xffff;
        r.v[2 + 2 * e] = w[e] >> 16;
    }
    // neighbours across the 8-sample group: from the adjacent lanes, memory only at the warp ends
    int left  = __shfl_up_sync(0xffffffffu, r.v[8], 1);
    int right = __shfl_down_sync(0xffffffffu, r.v[1], 1);
    if (lane == 0)
        left = x > 0 ? __ldg(row + x - 1) : 0;
    if (lane == 31)
        right = x + 8 < pw ? __ldg(row + x + 8) : 0;
    r.v[0] = left;
    r.v[9] = right;
    return r;
}

// byte (sel & 7) of {lo, hi} sign-extended to 32 bits: PTX prmt replicates the selected byte's sign where bit 3 of a
// selector nibble is set (__byte_perm masks that bit)
__device__ __forceinline__ int prmt_s8(unsigned lo, unsigned hi, unsigned sel)
{
    int d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
    return d;
}

__global__ void __launch_bounds__(32 * WARPS, 3) sao_kernel(const SaoK p)
{
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int c = blockIdx.z % p.planes, k = blockIdx.z / p.planes;
    const int pw = p.pw[c], ph = p.ph[c];
    const int x = (blockIdx.x * 32 + lane) * 8;
    const int y0 = (blockIdx.y * WARPS + warp) * RPT;
    if (y0 >= ph || (int)blockIdx.x * 256 >= pw)
        return;                                    // whole warp leaves together
    const bool live = x < pw;                      // dead lanes still take part in the shuffles
    const pel *src = p.src[c] + k * p.sb[c];
    pel *dst = p.dst[c] + k * p.db[c];
    const int pitch = p.sp[c];
    const int bd = p.bd;

    // per-CTB parameters of this 8-sample group (a group never straddles CTBs: CTB widths are multiples of 8)
    const int ctb_w = (1 << p.ctb_log2) >> p.hs[c], ctb_h = (1 << p.ctb_log2) >> p.vs[c];
    const int xs = live ? x : 0;
    const int cx = xs / ctb_w;
    const int bx0 = cx * ctb_w, bw = min(ctb_w, pw - bx0);

    // The RPT rows of a thread start at a multiple of RPT, so they lie in one CTB: its parameters are read once.  The five
    // edge offsets / four band offsets (6-bit magnitudes) become an 8-byte table indexed with PRMT (byte `idx`,
    // sign-extended), instead of a chain of compares per sample.
    const int cy = y0 / ctb_h;
    const int by0 = cy * ctb_h, bh = min(ctb_h, ph - by0);
    const VVCCudaSAOCtb *sp = p.ctbs + ((long long)k * p.ctb_rows + cy) * p.ctb_cols + cx;
    const int type = sp->type_idx[c];
    int off[5];
#pragma unroll
    for (int i = 0; i < 5; i++) off[i] = sp->offset_val[c][i];
    const int bp = sp->band_position[c], eo = sp->eo_class[c];
    // edge: category 0..4 -> offset index {1,2,0,3,4} (edge_idx, :53); band: band 0..3 -> offsets 1..4, band >= 4 -> 0
    const unsigned tab_lo = type == 1 ? ((off[1] & 0xff) | ((off[2] & 0xff) << 8) | ((off[3] & 0xff) << 16) | ((unsigned)(off[4] & 0xff) << 24))
                                      : ((off[1] & 0xff) | ((off[2] & 0xff) << 8) | ((off[0] & 0xff) << 16) | ((unsigned)(off[3] & 0xff) << 24));
    const unsigned tab_hi = type == 1 ? 0u : (unsigned)(off[4] & 0xff);
#define SAO_LOOKUP(idx) prmt_s8(tab_lo, tab_hi, (unsigned)(idx) * 0x1111u + 0x8880u)

    Row10 above, cur, below;
    cur   = load_row(src, pitch, pw, ph, xs, y0 - 1, lane);
    below = load_row(src, pitch, pw, ph, xs, y0, lane);
    for (int r = 0; r < RPT; r++) {
        const int y = y0 + r;
        if (y >= ph)
            break;                                 // warp-uniform
        above = cur; cur = below;
        below = load_row(src, pitch, pw, ph, xs, y + 1, lane);
        if (!live)
            continue;

        int out[8];
        if (type == 1) {
            const int sh = bd - 5;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int v = cur.v[1 + i];
                const int band = ((v >> sh) - bp) & 31;
                out[i] = d_clip_pel(v + SAO_LOOKUP(min(band, 4)), bd);
            }
        } else if (type == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int v = cur.v[1 + i];
                int a, b;
                if (eo == 0)      { a = cur.v[i];       b = cur.v[i + 2]; }
                else if (eo == 1) { a = above.v[i + 1]; b = below.v[i + 1]; }
                else if (eo == 2) { a = above.v[i];     b = below.v[i + 2]; }
                else              { a = above.v[i + 2]; b = below.v[i]; }
                const int cat = 2 + min(max(v - a, -1), 1) + min(max(v - b, -1), 1);
                out[i] = d_clip_pel(v + SAO_LOOKUP(cat), bd);
            }
            // ---- CTB border rules (:81-215), only for groups that touch the CTB border ----
            const int ry = y - by0;
            const int rx = x - bx0;
            if (ry == 0 || ry >= bh - 2 || rx == 0 || rx + 8 >= bw) {
                const bool bl = cx == 0, bt = cy == 0, br = cx == p.ctb_cols - 1, bb = cy == p.ctb_rows - 1;
                const bool not_v = eo != 1, not_h = eo != 0;
                const int nf = sp->no_filter;
                const bool restore = sp->restore;
                const int init_x = not_v && bl, w1 = bw - (not_v && br);
                const int init_y = not_h && bt, h1 = bh - (not_h && bb);
                const bool dg0 = nf & 16, dg1 = nf & 32, dg2 = nf & 64, dg3 = nf & 128;
                const int keep_ul = !dg0 && eo == 2 && !bl && !bt, keep_ur = !dg1 && eo == 3 && !bt && !br;
                const int keep_lr = !dg2 && eo == 2 && !br && !bb, keep_ll = !dg3 && eo == 3 && !bl && !bb;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int px = rx + i;
                    if (px >= bw)
                        break;
                    const int v = cur.v[1 + i];
                    if ((not_v && ((bl && px == 0) || (br && px == bw - 1))) ||
                        (not_h && ((bt && ry == 0) || (bb && ry == bh - 1))))
                        out[i] = d_clip_pel(v + off[0], bd);
                    if (restore) {
                        bool keep = false;
                        if ((nf & 1) && not_v && px == 0      && ry >= init_y + keep_ul && ry < h1 - keep_ll) keep = true;
                        if ((nf & 2) && not_v && px == w1 - 1 && ry >= init_y + keep_ur && ry < h1 - keep_lr) keep = true;
                        if ((nf & 4) && not_h && ry == 0      && px >= init_x + keep_ul && px < w1 - keep_ur) keep = true;
                        if ((nf & 8) && not_h && ry == h1 - 1 && px >= init_x + keep_ll && px < w1 - keep_lr) keep = true;
                        if (dg0 && eo == 2 && px == 0      && ry == 0)      keep = true;
                        if (dg1 && eo == 3 && px == w1 - 1 && ry == 0)      keep = true;
                        if (dg2 && eo == 2 && px == w1 - 1 && ry == h1 - 1) keep = true;
                        if (dg3 && eo == 3 && px == 0      && ry == h1 - 1) keep = true;
                        if (keep)
                            out[i] = v;
                    }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = cur.v[1 + i];
        }
        pel *drow = dst + (long long)y * p.dp[c] + x;
        if (x + 7 < pw) {
            uint4 o;
            o.x = out[0] | (out[1] << 16); o.y = out[2] | (out[3] << 16);
            o.z = out[4] | (out[5] << 16); o.w = out[6] | (out[7] << 16);
            *reinterpret_cast<uint4 *>(drow) = o;
        } else {
            for (int i = 0; x + i < pw; i++)
                drow[i] = (pel)out[i];
        }
    }
#undef SAO_LOOKUP
}

}  // namespace

extern "C" int vvc_cuda_sao_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                                  const VVCCudaSAOCtb *ctbs)
{
    if (ctx->err)
        return ctx->err;
    if (!dst || !src || !ctbs)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "sao: null argument");
    if ((src->bit_depth != 10 && src->bit_depth != 12) || src->ctb_log2 < 5 || src->ctb_log2 > 7 ||
        (src->width & 7) || (src->height & 7) || src->batch < 1 ||
        (src->chroma_format_idc && (src->hshift != 1 || src->vshift != 1)))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "sao: unsupported picture format");
    if (dst->width != src->width || dst->height != src->height || dst->batch != src->batch || dst->data[0] == src->data[0])
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "sao: dst must match src and not alias it");
    if (!frame_vec_ok(dst) || !frame_vec_ok(src))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "sao: planes and strides must be 16-byte aligned");
    SaoK p;
    p.planes = src->chroma_format_idc ? 3 : 1;
    for (int c = 0; c < 3; c++) {
        p.src[c] = (const pel *)src->data[c];  p.dst[c] = (pel *)dst->data[c];
        p.sp[c] = (int)(src->stride[c] / 2);   p.dp[c] = (int)(dst->stride[c] / 2);
        p.sb[c] = src->batch_stride[c] / 2;    p.db[c] = dst->batch_stride[c] / 2;
        p.hs[c] = c ? src->hshift : 0;         p.vs[c] = c ? src->vshift : 0;
        p.pw[c] = src->width >> p.hs[c];       p.ph[c] = src->height >> p.vs[c];
    }
    p.bd = src->bit_depth; p.ctb_log2 = src->ctb_log2;
    p.ctb_cols = ceil_div(src->width, 1 << src->ctb_log2);
    p.ctb_rows = ceil_div(src->height, 1 << src->ctb_log2);
    p.ctbs = ctbs;
    // the grid is sized for luma; chroma planes use its upper-left part and idle CTAs exit at once
    dim3 grid(ceil_div(src->width, 256), ceil_div(src->height, WARPS * RPT), src->batch * p.planes);
    sao_kernel<<<grid, dim3(32, WARPS), 0, ctx->stream>>>(p);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_sao_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                                       const VVCCudaSAOCtb *ctbs)
{
    if (ctx->err)
        return ctx->err;
    if (!dst || !src || !ctbs)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "sao_host: null argument");
    const int n_ctb = ceil_div(src->width, 1 << src->ctb_log2) * ceil_div(src->height, 1 << src->ctb_log2) * src->batch;
    const size_t fsz = align_up(vvc_stage_frame_size(src), 256);
    const size_t csz = align_up((size_t)n_ctb * sizeof(VVCCudaSAOCtb), 256);
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, 2 * fsz + csz);
    if (!base)
        return ctx->err;
    VVCCudaFrame a, b;
    vvc_stage_frame_layout(src, base, &a);
    vvc_stage_frame_layout(src, base + fsz, &b);
    VVCCudaSAOCtb *dctb = (VVCCudaSAOCtb *)(base + 2 * fsz);
    if (vvc_stage_frame_h2d(ctx, &a, src))
        return ctx->err;
    VVC_TRY(ctx, cudaMemcpyAsync(dctb, ctbs, (size_t)n_ctb * sizeof(VVCCudaSAOCtb), cudaMemcpyHostToDevice, ctx->stream));
    if (vvc_cuda_sao_frame(ctx, &b, &a, dctb))
        return ctx->err;
    if (vvc_stage_frame_d2h(ctx, dst, &b))
        return ctx->err;
    return vvc_cuda_sync(ctx);
}
