// SAO stage for sm_100a: one launch per picture ring, all planes.
//
// Replaces ff_vvc_sao_filter (libavcodec/vvc/vvc_filter.c:154-298) and the table entries
//   sao_band_filter      libavcodec/h26x/h2656_sao_template.c:24-46
//   sao_edge_filter      libavcodec/h26x/h2656_sao_template.c:50-79
//   sao_edge_restore_0/1 libavcodec/h26x/h2656_sao_template.c:81-215
//
// B200 design: purely streaming, 16x2 packed arithmetic, CTB-uniform warps.
//  * A warp works inside ONE CTB of one plane: its lanes tile the CTB's width in 8-sample groups (one 128-bit load /
//    store each) and the remaining lane bits select row groups of RPT rows, so type / class / offsets are warp-uniform
//    and no lane idles because its neighbour's CTB has another SAO type (a 256-sample-wide warp straddles two to four
//    CTBs and ran with 18 of 32 lanes active).
//  * Samples stay packed two to a register.  Edge: 1 + clamp(n - v, -1, 1) is one VIADDMNMX.RELU per neighbour, the
//    category is their sum, the offset comes out of an 8-byte table with two PRMTs (sign-extending), and
//    clip(v + offset) is one more VIADDMNMX.RELU: 8 instructions per sample pair.  Band: shift / mask / add / mask /
//    min, then the same lookup and clip.
//  * Left / right neighbours of an 8-sample group come from the adjacent lanes by shuffle; only the lanes at the CTB's
//    sides touch memory for them.  The reference reads its neighbours from a per-CTB copy with a halo of saved pre-SAO
//    lines; here the pre-SAO picture is simply read-only, which is the same thing.
//  * The picture-border and restore rules (:81-215) are per-lane 8-bit masks, computed only in CTBs that have any.
#include "common.cuh"

namespace {

struct SaoK {
    const pel *src[3];
    pel       *dst[3];
    int        sp[3], dp[3];
    long long  sb[3], db[3];
    int        pw[3], ph[3], hs[3], vs[3];
    int        cta_start[4];        // first CTA of each plane inside a picture's share of the grid; [3] = CTAs per picture
    int        bd, ctb_log2, ctb_cols, ctb_rows, planes;
    const VVCCudaSAOCtb *ctbs;
};

#ifndef SAO_RPT
#define SAO_RPT 4
#endif
constexpr int RPT = SAO_RPT;    // rows per lane
constexpr int WARPS = 8;        // warps per CTA, each on its own rows

struct Row { uint32_t w[4]; uint32_t left, right; };     // 8 samples as 4 pairs; left = (., x-1) in the high half, right = (x+8, .) in the low half

__device__ __forceinline__ void load_row8(uint32_t w[4], const pel *row, int x, int pw)
{
    if (x + 7 < pw) {
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(row + x));
        w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
    } else {                                       // the last group of a plane whose width is not a multiple of 8
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const unsigned lo = x + 2 * e     < pw ? __ldg(row + x + 2 * e)     : 0;
            const unsigned hi = x + 2 * e + 1 < pw ? __ldg(row + x + 2 * e + 1) : 0;
            w[e] = lo | (hi << 16);
        }
    }
}

__device__ __forceinline__ void store_row8(pel *row, int x, int pw, const uint32_t o[4])
{
    if (x + 7 < pw) {
        *reinterpret_cast<uint4 *>(row + x) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (x + i < pw)
                row[x + i] = (pel)(o[i >> 1] >> ((i & 1) * 16));
    }
}

// a row with its two outer neighbours: from the adjacent lanes of the row group, from memory at the CTB's sides
__device__ __forceinline__ Row load_row_nb(const pel *plane, int pitch, int pw, int ph, int x, int y, bool first, bool last)
{
    Row r;
    y = min(max(y, 0), ph - 1);
    const pel *row = plane + (long long)y * pitch;
    load_row8(r.w, row, x, pw);
    r.left  = __shfl_up_sync(0xffffffffu, r.w[3], 1);
    r.right = __shfl_down_sync(0xffffffffu, r.w[0], 1);
    if (first)
        r.left = x > 0 ? (uint32_t)__ldg(row + x - 1) << 16 : 0;
    if (last)
        r.right = x + 8 < pw ? __ldg(row + x + 8) : 0;
    return r;
}

// offsets of both halves of a pair of table indices (0..7 each) as sign-extended 16-bit values: PTX prmt replicates
// the selected byte's sign where bit 3 of a selector nibble is set
__device__ __forceinline__ uint32_t lookup2(uint32_t tab_lo, uint32_t tab_hi, uint32_t idx2)
{
    const uint32_t n = idx2 * 0x11u + 0x00800080u;      // per half: (idx, idx | 8) in the low byte
    uint32_t sel, d;
    asm("prmt.b32 %0, %1, %1, 0x0020;" : "=r"(sel) : "r"(n));
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(tab_lo), "r"(tab_hi), "r"(sel));
    return d;
}

__device__ __forceinline__ uint32_t pack_bytes(int b0, int b1, int b2, int b3)
{
    return (uint32_t)(b0 & 0xff) | ((uint32_t)(b1 & 0xff) << 8) | ((uint32_t)(b2 & 0xff) << 16) | ((uint32_t)(b3 & 0xff) << 24);
}

// 8-bit mask -> the pairs' 16-bit lane masks
__device__ __forceinline__ uint32_t pair_mask(unsigned m8, int i)
{
    return ((m8 >> (2 * i)) & 1 ? 0x0000ffffu : 0u) | ((m8 >> (2 * i + 1)) & 1 ? 0xffff0000u : 0u);
}

// samples x-1 / x+1 of pair i of a row
__device__ __forceinline__ uint32_t left_of(const Row &r, int i)  { return __funnelshift_r(i ? r.w[i - 1] : r.left, r.w[i], 16); }
__device__ __forceinline__ uint32_t right_of(const Row &r, int i) { return __funnelshift_r(r.w[i], i < 3 ? r.w[i + 1] : r.right, 16); }

// One row of edge offset, class EO: neighbours a / b = left / right, above / below, above-left / below-right,
// above-right / below-left (:54-59).
template <int EO>
__device__ __forceinline__ void edge_row(uint32_t o[4], const Row &above, const Row &cur, const Row &below,
                                         uint32_t tab_lo, uint32_t tab_hi, uint32_t maxv2)
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t a = EO == 0 ? left_of(cur, i) : EO == 1 ? above.w[i] : EO == 2 ? left_of(above, i) : right_of(above, i);
        const uint32_t b = EO == 0 ? right_of(cur, i) : EO == 1 ? below.w[i] : EO == 2 ? right_of(below, i) : left_of(below, i);
        const uint32_t v = cur.w[i], nv1 = __vsub2(0x00010001u, v);
        const uint32_t cat = __viaddmin_s16x2_relu(a, nv1, 0x00020002u) + __viaddmin_s16x2_relu(b, nv1, 0x00020002u);
        o[i] = __viaddmin_s16x2_relu(v, lookup2(tab_lo, tab_hi, cat), maxv2);
    }
}

#ifndef SAO_MIN_CTAS
#define SAO_MIN_CTAS 8               // resident CTAs (of 8 warps) per SM the kernel is compiled for: full occupancy at 32 registers (tools/sweep_sao_occ.sh: 3 / 4 / 5 / 6 / 8 -> 0.32 / 0.257 / 0.236 / 0.244 / 0.216 ms per 16 pictures)
#endif
__global__ void __launch_bounds__(32 * WARPS, SAO_MIN_CTAS) sao_kernel(const SaoK p)
{
    const int lane = threadIdx.x, warp = threadIdx.y;
    // grid: x = CTB column, y = the row chunks of plane 0, then 1, then 2, z = picture (no index arithmetic to undo)
    const int k = blockIdx.z, cx = blockIdx.x;
    int r = blockIdx.y;
    const int c = r >= p.cta_start[2] ? 2 : r >= p.cta_start[1] ? 1 : 0;
    r -= p.cta_start[c];
    const int chunk = r * WARPS + warp;

    const int pw = p.pw[c], ph = p.ph[c];
    const int ctb_w = (1 << p.ctb_log2) >> p.hs[c], ctb_h = (1 << p.ctb_log2) >> p.vs[c];
    const int lxl = p.ctb_log2 - p.hs[c] - 3;                  // log2 of the 8-sample groups across a CTB
    const int groups = min(32 >> lxl, ctb_h / RPT);            // row groups of a warp (all inside one CTB row)
    const int lx = lane & ((1 << lxl) - 1), g = lane >> lxl;
    const int bx0 = cx * ctb_w, bw = min(ctb_w, pw - bx0);
    const int x = bx0 + 8 * lx;
    const int wy0 = chunk * RPT * groups;                      // first row of the warp
    if (wy0 >= ph)
        return;
    const int y0 = wy0 + g * RPT;
    const bool live = g < groups && 8 * lx < bw && y0 < ph;    // dead lanes still take part in the shuffles
    const int xs = live ? x : bx0, ys = live ? y0 : wy0;
    const int cy = wy0 >> (p.ctb_log2 - p.vs[c]);
    const int by0 = cy * ctb_h, bh = min(ctb_h, ph - by0);

    const pel *src = p.src[c] + k * p.sb[c];
    pel *dst = p.dst[c] + k * p.db[c];
    const int pitch = p.sp[c], dpitch = p.dp[c];
    const int bd = p.bd;
    const uint32_t maxv2 = (uint32_t)((1 << bd) - 1) * 0x00010001u;

    const VVCCudaSAOCtb *sp = p.ctbs + ((long long)k * p.ctb_rows + cy) * p.ctb_cols + cx;
    const int type = sp->type_idx[c];
    int off[5];
#pragma unroll
    for (int i = 0; i < 5; i++) off[i] = sp->offset_val[c][i];

    if (type == 1) {
        // band (:24-46): band k = 0..3 from band_position on takes offset_val[k + 1], every other band 0
        const uint32_t tab_lo = pack_bytes(off[1], off[2], off[3], off[4]), tab_hi = 0;
        const uint32_t sub = (32u - sp->band_position[c]) * 0x00010001u;
        const int sh = bd - 5;
        if (live) {
            for (int rr = 0; rr < RPT && ys + rr < ph; rr++) {
                uint32_t w[4], o[4];
                load_row8(w, src + (long long)(ys + rr) * pitch, xs, pw);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t band = (((w[i] >> sh) & 0x001f001fu) + sub) & 0x001f001fu;
                    o[i] = __viaddmin_s16x2_relu(w[i], lookup2(tab_lo, tab_hi, __vminu2(band, 0x00040004u)), maxv2);
                }
                store_row8(dst + (long long)(ys + rr) * dpitch, xs, pw, o);
            }
        }
        return;
    }
    if (type != 2) {
        if (live) {
            for (int rr = 0; rr < RPT && ys + rr < ph; rr++) {
                uint32_t w[4];
                load_row8(w, src + (long long)(ys + rr) * pitch, xs, pw);
                store_row8(dst + (long long)(ys + rr) * dpitch, xs, pw, w);
            }
        }
        return;
    }

    // ---- edge (:50-79).  s = 1 + clamp(n - v, -1, 1) per neighbour; s_a + s_b = 4 - (2 + sign(v - a) + sign(v - b)), so
    // the table is edge_idx[] = {1, 2, 0, 3, 4} read backwards ----
    const int eo = sp->eo_class[c];
    const uint32_t tab_lo = pack_bytes(off[4], off[3], off[0], off[2]), tab_hi = pack_bytes(off[1], 0, 0, 0);
    const uint32_t off0 = (uint32_t)(off[0] & 0xffff) * 0x00010001u;
    const bool first = lx == 0, last = 8 * lx + 8 >= bw;

    // CTB border rules (:81-215): which of them can fire in this CTB at all (warp-uniform)
    const bool bl = cx == 0, bt = cy == 0, br = cx == p.ctb_cols - 1, bb = cy == p.ctb_rows - 1;
    const bool not_v = eo != 1, not_h = eo != 0;
    const int nf = sp->restore ? sp->no_filter : 0;
    const bool rules = nf || (not_v && (bl || br)) || (not_h && (bt || bb));
    const int init_x = not_v && bl, w1 = bw - (not_v && br);
    const int init_y = not_h && bt, h1 = bh - (not_h && bb);
    const bool dg0 = nf & 16, dg1 = nf & 32, dg2 = nf & 64, dg3 = nf & 128;
    const int keep_ul = !dg0 && eo == 2 && !bl && !bt, keep_ur = !dg1 && eo == 3 && !bt && !br;
    const int keep_lr = !dg2 && eo == 2 && !br && !bb, keep_ll = !dg3 && eo == 3 && !bl && !bb;
    const int rx = xs - bx0;

    Row above, cur, below;
    cur   = load_row_nb(src, pitch, pw, ph, xs, ys - 1, first, last);
    below = load_row_nb(src, pitch, pw, ph, xs, ys, first, last);
    for (int rr = 0; rr < RPT; rr++) {
        const int y = ys + rr;
        if (wy0 + rr >= ph)
            break;                                 // warp-uniform: row groups only add multiples of RPT, and ph is one too
        above = cur; cur = below;
        below = load_row_nb(src, pitch, pw, ph, xs, y + 1, first, last);

        uint32_t o[4];
        switch (eo) {
        case 0:  edge_row<0>(o, above, cur, below, tab_lo, tab_hi, maxv2); break;
        case 1:  edge_row<1>(o, above, cur, below, tab_lo, tab_hi, maxv2); break;
        case 2:  edge_row<2>(o, above, cur, below, tab_lo, tab_hi, maxv2); break;
        default: edge_row<3>(o, above, cur, below, tab_lo, tab_hi, maxv2); break;
        }
        if (rules) {
            const int ry = y - by0;
            unsigned zero = 0, keep = 0;
            const bool zrow = not_h && ((bt && ry == 0) || (bb && ry == bh - 1));
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int px = rx + i;
                if (zrow || (not_v && ((bl && px == 0) || (br && px == bw - 1))))
                    zero |= 1u << i;
                bool kp = false;
                if ((nf & 1) && not_v && px == 0      && ry >= init_y + keep_ul && ry < h1 - keep_ll) kp = true;
                if ((nf & 2) && not_v && px == w1 - 1 && ry >= init_y + keep_ur && ry < h1 - keep_lr) kp = true;
                if ((nf & 4) && not_h && ry == 0      && px >= init_x + keep_ul && px < w1 - keep_ur) kp = true;
                if ((nf & 8) && not_h && ry == h1 - 1 && px >= init_x + keep_ll && px < w1 - keep_lr) kp = true;
                if (dg0 && eo == 2 && px == 0      && ry == 0)      kp = true;
                if (dg1 && eo == 3 && px == w1 - 1 && ry == 0)      kp = true;
                if (dg2 && eo == 2 && px == w1 - 1 && ry == h1 - 1) kp = true;
                if (dg3 && eo == 3 && px == 0      && ry == h1 - 1) kp = true;
                if (kp)
                    keep |= 1u << i;
            }
            if (zero | keep) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t v = cur.w[i];
                    const uint32_t mz = pair_mask(zero, i), mk = pair_mask(keep, i);
                    o[i] = (o[i] & ~mz) | (__viaddmin_s16x2_relu(v, off0, maxv2) & mz);
                    o[i] = (o[i] & ~mk) | (v & mk);
                }
            }
        }
        if (live && y < ph)
            store_row8(dst + (long long)y * dpitch, xs, pw, o);
    }
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
    // grid.y: the row chunks of plane 0, then 1, then 2; a CTA is one CTB column x WARPS warps of rows
    int total = 0;
    for (int c = 0; c < p.planes; c++) {
        const int ctb_h = (1 << p.ctb_log2) >> p.vs[c], lxl = p.ctb_log2 - p.hs[c] - 3;
        const int groups = (32 >> lxl) < ctb_h / RPT ? (32 >> lxl) : ctb_h / RPT;
        p.cta_start[c] = total;
        total += ceil_div(ceil_div(p.ph[c], RPT * groups), WARPS);
    }
    for (int c = p.planes; c < 4; c++)
        p.cta_start[c] = total;
    p.cta_start[3] = total;
    if (total > 65535 || src->batch > 65535)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "sao: picture ring exceeds the grid");
    sao_kernel<<<dim3(p.ctb_cols, total, src->batch), dim3(32, WARPS), 0, ctx->stream>>>(p);
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
