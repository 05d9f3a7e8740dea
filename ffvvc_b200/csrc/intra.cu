// Intra leaf predictors and the CIIP blend for sm_100a: one launch predicts every listed block.
//
// Replaces the table entries intra.pred_planar :686, pred_dc :826-864, pred_v :866, pred_h :877,
// pred_angular_v :894, pred_angular_h :950, pred_mip :773 (with mip_downsampling :708, mip_reduced_pred :728,
// mip_upsampling_1d :749) of libavcodec/vvc/vvc_intra_template.c, the helpers they call
// (ff_vvc_intra_pred_angle_derive / inv_angle_derive / nscale_derive / get_mip_size_id,
// libavcodec/vvc/vvc_intra.c:529-690) and inter.put_ciip (libavcodec/vvc/vvc_inter_template.c:60-76).
//
// B200 design: intra blocks of one call are independent (the host submits one wavefront of the RECON stage
// at a time), so the kernel is a flat map: a CTA walks records, a thread owns output samples, each sample is
// computed from the block's two reference lines (at most 2 x 133 samples, L1 resident) as a closed formula
// of its position -- the reference's row-by-row running state (pos += dp, inv_angle_sum += inv_angle) is just
// a linear function of the row.  MIP goes through shared memory: reduced boundary, 16/64-sample matrix
// product, then the two separable up-sampling passes.  The float in the reference's inverse-angle derivation
// is an exact integer expression for the 30 angles that occur ((32768 + a) / (2 a), checked by the tests).
#include "common.cuh"
#include "tables.cuh"

namespace {

constexpr int kThreads = 128;

struct IntraK {
    pel       *plane[3];
    int        pitch[3];
    long long  bstride[3];
    int        bd;
    const VVCCudaIntraPB *pbs;
    int        n;
    const uint16_t *edges;
};

__constant__ short c_angles[31] = { 0, 1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 23, 26, 29,
                                    32, 35, 39, 45, 51, 57, 64, 73, 86, 102, 128, 171, 256, 341, 512 };

__device__ __forceinline__ int pred_angle(int mode)
{
    int idx = mode > 34 ? mode - 50 : mode > 0 ? 18 - mode : 16 - mode;
    const int sign = idx < 0 ? -1 : 1;
    return sign * c_angles[idx < 0 ? -idx : idx];
}

#define SEL3(a, c) ((c) == 0 ? (a)[0] : (c) == 1 ? (a)[1] : (a)[2])

__global__ void __launch_bounds__(kThreads) intra_leaf_kernel(const IntraK p)
{
    __shared__ int s_red[16];
    __shared__ int s_small[8][8];
    __shared__ int s_dc;
    const int tid = threadIdx.x, bd = p.bd, maxv = (1 << bd) - 1;
    for (int ri = blockIdx.x; ri < p.n; ri += gridDim.x) {
        const VVCCudaIntraPB b = p.pbs[ri];
        const int w = b.w, h = b.h, lw = 31 - __clz(w), lh = 31 - __clz(h);
        const int pitch = SEL3(p.pitch, b.c_idx);
        pel *dst = SEL3(p.plane, b.c_idx) + b.pic * SEL3(p.bstride, b.c_idx) + (long long)b.y0 * pitch + b.x0;
        const uint16_t *top = p.edges + b.top, *left = p.edges + b.left;
        __syncthreads();                                   // previous record is done with shared memory

        if (b.kind == VVC_CUDA_INTRA_MIP) {
            const bool transposed = b.flags & VVC_CUDA_INTRA_MIP_TRANSPOSED;
            const int size_id = (w == 4 && h == 4) ? 0 : ((w == 4 || h == 4) || (w == 8 && h == 8)) ? 1 : 2;
            const int bsz = size_id ? 4 : 2, psz = size_id == 2 ? 8 : 4, in_size = 2 * bsz - (size_id == 2);
            const uint8_t *matrix = size_id == 0 ? &vvct_mip_matrix_0[b.mode][0][0] : size_id == 1 ? &vvct_mip_matrix_1[b.mode][0][0]
                                                                                                     : &vvct_mip_matrix_2[b.mode][0][0];
            const int up_hor = w / psz, up_ver = h / psz;
            if (tid < 2 * bsz) {                           // boundary down-sampling (:708-726)
                const int side = tid >= bsz, i = tid - side * bsz;
                const uint16_t *ref = side ? left : top;
                const int n = side ? h : w, dwn = n / bsz, lg = 31 - __clz(dwn);
                int r = 0;
                for (int j = 0; j < dwn; j++)
                    r += ref[i * dwn + j];
                r = dwn == 1 ? r : (r + (1 << (lg - 1))) >> lg;
                s_red[(side != transposed ? bsz : 0) + i] = r;     // top first, unless transposed
            }
            __syncthreads();
            int temp0 = 0, ow = 0;
            if (tid == 0) {
                temp0 = s_red[0];
                const int off = size_id != 2 ? 0 : 1;
                ow = size_id != 2 ? (1 << (bd - 1)) - temp0 : s_red[1] - temp0;
                s_red[0] = ow;
                for (int i = 1; i < in_size; i++) {
                    s_red[i] = s_red[i + off] - temp0;
                    ow += s_red[i];
                }
                s_small[0][0] = 32 - 32 * ow;              // parked for the broadcast below
                s_dc = temp0;
            }
            __syncthreads();
            ow = s_small[0][0]; temp0 = s_dc;
            __syncthreads();
            if (tid < psz * psz) {                         // mip_reduced_pred (:728-747)
                const int y = tid / psz, x = tid - y * psz;
                int pred = 0;
                for (int i = 0; i < in_size; i++)
                    pred += s_red[i] * (int)matrix[tid * in_size + i];
                pred = d_clip3(((pred + ow) >> 6) + temp0, 0, maxv);
                if (transposed) s_small[x][y] = pred; else s_small[y][x] = pred;
            }
            __syncthreads();
            // rows that hold reduced samples: horizontal up-sampling between the left boundary and the samples (:749-771)
            for (int idx = tid; idx < psz * w; idx += kThreads) {
                const int j = idx / w, x = idx - j * w;
                const int i = x / up_hor, k = x - i * up_hor + 1;          // k = 1 .. up_hor, k == up_hor is the sample itself
                const int after = s_small[j][i];
                const int before = i ? s_small[j][i - 1] : (int)left[(j + 1) * up_ver - 1];
                const int v = k == up_hor ? after : ((up_hor - k) * before + k * after + up_hor / 2) / up_hor;
                dst[(long long)((j + 1) * up_ver - 1) * pitch + x] = (pel)v;
            }
            __syncthreads();
            if (up_ver > 1) {                              // vertical up-sampling between the top boundary and those rows
                for (int idx = tid; idx < w * h; idx += kThreads) {
                    const int y = idx / w, x = idx - y * w;
                    const int j = y / up_ver, k = y - j * up_ver + 1;
                    if (k == up_ver)
                        continue;
                    const int after = dst[(long long)((j + 1) * up_ver - 1) * pitch + x];
                    const int before = j ? (int)dst[(long long)(j * up_ver - 1) * pitch + x] : (int)top[x];
                    dst[(long long)y * pitch + x] = (pel)(((up_ver - k) * before + k * after + up_ver / 2) / up_ver);
                }
            }
            continue;
        }

        if (b.kind == VVC_CUDA_INTRA_DC) {                 // pred_dc_val (:826-845), one warp sums
            if (tid < 32) {
                int sum = 0;
                if (w >= h) for (int k = tid; k < w; k += 32) sum += top[k];
                if (w <= h) for (int k = tid; k < h; k += 32) sum += left[k];
#pragma unroll
                for (int o = 16; o; o >>= 1)
                    sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const unsigned offset = w == h ? (unsigned)w << 1 : (unsigned)max(w, h);
                if (tid == 0)
                    s_dc = (sum + (int)(offset >> 1)) >> (31 - __clz(offset));
            }
            __syncthreads();
        }
        const bool vertical = b.kind == VVC_CUDA_INTRA_ANGULAR_V;
        const bool angular = vertical || b.kind == VVC_CUDA_INTRA_ANGULAR_H;
        int angle = 0, inv_angle = 0, nscale = 0;
        const bool pdpc = angular && (b.flags & VVC_CUDA_INTRA_PDPC);
        if (angular) {
            angle = pred_angle(b.mode);
            if (pdpc) {
                inv_angle = (32768 + angle) / (2 * angle);                 // ff_vvc_intra_inv_angle_derive for angle > 0
                const int side = b.mode >= 50 ? h : w;
                nscale = min(2, (31 - __clz(side)) - (31 - __clz(3 * inv_angle - 2)) + 8);
            }
        }
        const int ref_idx = b.ref_idx, is_luma = !b.c_idx;
        const uint16_t *mainr = (vertical ? top : left) - (1 + ref_idx), *side_ref = vertical ? left : top;
        for (int idx = tid; idx < w * h; idx += kThreads) {
            const int y = idx >> lw, x = idx & (w - 1);
            int pred;
            if (b.kind == VVC_CUDA_INTRA_PLANAR) {
                const int pv = ((h - 1 - y) * top[x] + (y + 1) * left[h]) << lw;
                const int ph = ((w - 1 - x) * left[y] + (x + 1) * top[w]) << lh;
                pred = (pv + ph + w * h) >> (lw + lh + 1);
            } else if (b.kind == VVC_CUDA_INTRA_DC) {
                pred = s_dc;
            } else if (b.kind == VVC_CUDA_INTRA_VERT) {
                pred = top[x];
            } else if (b.kind == VVC_CUDA_INTRA_HORZ) {
                pred = left[y];
            } else {
                const int along = vertical ? x : y, across = vertical ? y : x;
                const int pos = (across + 1 + ref_idx) * angle, ix = (pos >> 5) + ref_idx, fact = pos & 31;
                const uint16_t *q = mainr + along + ix;
                if (!fact && (!is_luma || !b.filter_flag)) {
                    pred = q[1];
                } else if (is_luma) {
                    const int8_t *f = vvct_intra_luma_filter[b.filter_flag][fact];
                    pred = d_clip3((q[0] * f[0] + q[1] * f[1] + q[2] * f[2] + q[3] * f[3] + 32) >> 6, 0, maxv);
                } else {
                    pred = ((32 - fact) * q[1] + fact * q[2] + 16) >> 5;
                }
                if (pdpc) {
                    if (vertical) {
                        if (x < min(w, 3 << nscale)) {
                            const int l = side_ref[y + ((256 + (x + 1) * inv_angle) >> 9)], wl = 32 >> ((x << 1) >> nscale);
                            pred = d_clip3(pred + (((l - pred) * wl + 32) >> 6), 0, maxv);
                        }
                    } else if (y < (3 << nscale)) {
                        const int t = side_ref[x + ((256 + (y + 1) * inv_angle) >> 9)], wt = 32 >> min(31, (y * 2) >> nscale);
                        pred = d_clip3(pred + (((t - pred) * wt + 32) >> 6), 0, maxv);
                    }
                }
            }
            dst[(long long)y * pitch + x] = (pel)pred;
        }
    }
}

struct CiipK {
    pel       *dst[3];
    const pel *src[3];
    int        dp[3], sp[3];
    long long  db[3], sb[3];
    const VVCCudaCiip *blocks;
    int        n;
};

__global__ void __launch_bounds__(kThreads) ciip_kernel(const CiipK p)
{
    for (int ri = blockIdx.x; ri < p.n; ri += gridDim.x) {
        const VVCCudaCiip b = p.blocks[ri];
        const int c = b.c_idx, wi = b.intra_weight;
        pel *d = SEL3(p.dst, c) + b.pic * SEL3(p.db, c) + (long long)b.y0 * SEL3(p.dp, c) + b.x0;
        const pel *s = SEL3(p.src, c) + b.pic * SEL3(p.sb, c) + (long long)b.y0 * SEL3(p.sp, c) + b.x0;
        const int dpitch = SEL3(p.dp, c), spitch = SEL3(p.sp, c);
        for (int idx = threadIdx.x; idx < b.w * b.h; idx += kThreads) {
            const int y = idx / b.w, x = idx - y * b.w;
            pel *q = d + (long long)y * dpitch + x;
            *q = (pel)((*q * wi + s[(long long)y * spitch + x] * (4 - wi) + 2) >> 2);
        }
    }
}

}  // namespace

extern "C" int vvc_cuda_intra_leaf_frame(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraPB *pbs, int n_pbs,
                                         const uint16_t *edges)
{
    if (ctx->err)
        return ctx->err;
    if (!frame || (n_pbs > 0 && (!pbs || !edges)))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "intra: null argument");
    if (frame->bit_depth != 10 && frame->bit_depth != 12)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "intra: bit depth %d not accelerated (10/12 only)", frame->bit_depth);
    if (n_pbs <= 0)
        return VVC_CUDA_OK;
    IntraK p;
    for (int c = 0; c < 3; c++) {
        p.plane[c] = (pel *)frame->data[c];
        p.pitch[c] = (int)(frame->stride[c] / 2);
        p.bstride[c] = frame->batch_stride[c] / 2;
    }
    p.bd = frame->bit_depth; p.pbs = pbs; p.n = n_pbs; p.edges = edges;
    intra_leaf_kernel<<<n_pbs < 148 * 8 ? n_pbs : 148 * 8, kThreads, 0, ctx->stream>>>(p);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_intra_leaf_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraPB *pbs, int n_pbs,
                                              const uint16_t *edges, size_t n_edges)
{
    if (ctx->err)
        return ctx->err;
    if (!frame || (n_pbs > 0 && (!pbs || !edges)))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "intra_host: null argument");
    const size_t fsz = align_up(vvc_stage_frame_size(frame), 256), psz = align_up((size_t)n_pbs * sizeof(VVCCudaIntraPB), 256);
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, fsz + psz + align_up(n_edges * sizeof(uint16_t), 256));
    if (!base)
        return ctx->err;
    VVCCudaFrame d;
    vvc_stage_frame_layout(frame, base, &d);
    VVCCudaIntraPB *dpb = (VVCCudaIntraPB *)(base + fsz);
    uint16_t *ded = (uint16_t *)(base + fsz + psz);
    // the picture travels both ways: samples outside the listed blocks keep their content
    if (vvc_stage_frame_h2d(ctx, &d, frame))
        return ctx->err;
    if (n_pbs > 0) {
        VVC_TRY(ctx, cudaMemcpyAsync(dpb, pbs, (size_t)n_pbs * sizeof(VVCCudaIntraPB), cudaMemcpyHostToDevice, ctx->stream));
        VVC_TRY(ctx, cudaMemcpyAsync(ded, edges, n_edges * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (vvc_cuda_intra_leaf_frame(ctx, &d, dpb, n_pbs, ded) || vvc_stage_frame_d2h(ctx, frame, &d))
        return ctx->err;
    return vvc_cuda_sync(ctx);
}

extern "C" int vvc_cuda_ciip_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *inter,
                                   const VVCCudaCiip *blocks, int n_blocks)
{
    if (ctx->err)
        return ctx->err;
    if (!dst || !inter || (n_blocks > 0 && !blocks))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "ciip: null argument");
    if (dst->width != inter->width || dst->height != inter->height || dst->bit_depth != inter->bit_depth)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "ciip: dst / inter geometry differs");
    if (n_blocks <= 0)
        return VVC_CUDA_OK;
    CiipK p;
    for (int c = 0; c < 3; c++) {
        p.dst[c] = (pel *)dst->data[c];          p.src[c] = (const pel *)inter->data[c];
        p.dp[c] = (int)(dst->stride[c] / 2);     p.sp[c] = (int)(inter->stride[c] / 2);
        p.db[c] = dst->batch_stride[c] / 2;      p.sb[c] = inter->batch_stride[c] / 2;
    }
    p.blocks = blocks; p.n = n_blocks;
    ciip_kernel<<<n_blocks < 148 * 8 ? n_blocks : 148 * 8, kThreads, 0, ctx->stream>>>(p);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_ciip_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *inter,
                                        const VVCCudaCiip *blocks, int n_blocks)
{
    if (ctx->err)
        return ctx->err;
    if (!dst || !inter || (n_blocks > 0 && !blocks))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "ciip_host: null argument");
    const size_t fsz = align_up(vvc_stage_frame_size(dst), 256), bsz = align_up((size_t)n_blocks * sizeof(VVCCudaCiip), 256);
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, 2 * fsz + bsz);
    if (!base)
        return ctx->err;
    VVCCudaFrame dd, di;
    vvc_stage_frame_layout(dst, base, &dd);
    vvc_stage_frame_layout(inter, base + fsz, &di);
    VVCCudaCiip *db = (VVCCudaCiip *)(base + 2 * fsz);
    if (vvc_stage_frame_h2d(ctx, &dd, dst) || vvc_stage_frame_h2d(ctx, &di, inter))
        return ctx->err;
    if (n_blocks > 0)
        VVC_TRY(ctx, cudaMemcpyAsync(db, blocks, (size_t)n_blocks * sizeof(VVCCudaCiip), cudaMemcpyHostToDevice, ctx->stream));
    if (vvc_cuda_ciip_frame(ctx, &dd, &di, db, n_blocks) || vvc_stage_frame_d2h(ctx, dst, &dd))
        return ctx->err;
    return vvc_cuda_sync(ctx);
}
