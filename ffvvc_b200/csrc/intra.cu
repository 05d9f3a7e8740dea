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
#include <stdlib.h>
#include "common.cuh"
#include "coeff_src.cuh"
#include "tables.cuh"

namespace {

#ifndef INTRA_THREADS
#define INTRA_THREADS 128
#endif
constexpr int kThreads = INTRA_THREADS;

#include "itx_generic.cuh"        // process_tb: the residual step of the dependency-driven kernel below

// A picture sample another CTA may have written during this launch (the dependency-driven kernel): read through L2, the
// SM's own L1 may hold the line from before the sample was final.  The wave-by-wave kernels read plainly.
template <bool CG> __device__ __forceinline__ int ld_pel(const pel *q) { return CG ? (int)__ldcg(q) : (int)*q; }

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

struct LeafScratch {
    int red[16];
    int small[8][8];
    int dc;
};

// what the leaf predictors need to know about a block
struct Leaf {
    int kind, mode, ref_idx, filter_flag, flags, w, h, c_idx;
    int pdpc4;               // position-dependent filtering of planar / DC / V / H (intra_pred's own tail, :653-681)
};

// One block by the whole CTA.  top / left: the reference lines as IntraEdgeParams carries them (global or shared memory).
__device__ void leaf_predict(const Leaf &b, const uint16_t *top, const uint16_t *left, pel *dst, int pitch, int bd, LeafScratch &sc)
{
    const int tid = threadIdx.x, maxv = (1 << bd) - 1;
    const int w = b.w, h = b.h, lw = 31 - __clz(w), lh = 31 - __clz(h);
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
            sc.red[(side != transposed ? bsz : 0) + i] = r;     // top first, unless transposed
        }
        __syncthreads();
        int temp0 = 0, ow = 0;
        if (tid == 0) {
            temp0 = sc.red[0];
            const int off = size_id != 2 ? 0 : 1;
            ow = size_id != 2 ? (1 << (bd - 1)) - temp0 : sc.red[1] - temp0;
            sc.red[0] = ow;
            for (int i = 1; i < in_size; i++) {
                sc.red[i] = sc.red[i + off] - temp0;
                ow += sc.red[i];
            }
            sc.small[0][0] = 32 - 32 * ow;              // parked for the broadcast below
            sc.dc = temp0;
        }
        __syncthreads();
        ow = sc.small[0][0]; temp0 = sc.dc;
        __syncthreads();
        if (tid < psz * psz) {                         // mip_reduced_pred (:728-747)
            const int y = tid / psz, x = tid - y * psz;
            int pred = 0;
            for (int i = 0; i < in_size; i++)
                pred += sc.red[i] * (int)matrix[tid * in_size + i];
            pred = d_clip3(((pred + ow) >> 6) + temp0, 0, maxv);
            if (transposed) sc.small[x][y] = pred; else sc.small[y][x] = pred;
        }
        __syncthreads();
        // rows that hold reduced samples: horizontal up-sampling between the left boundary and the samples (:749-771)
        for (int idx = tid; idx < psz * w; idx += kThreads) {
            const int j = idx / w, x = idx - j * w;
            const int i = x / up_hor, k = x - i * up_hor + 1;          // k = 1 .. up_hor, k == up_hor is the sample itself
            const int after = sc.small[j][i];
            const int before = i ? sc.small[j][i - 1] : (int)left[(j + 1) * up_ver - 1];
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
        return;
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
                sc.dc = (sum + (int)(offset >> 1)) >> (31 - __clz(offset));
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
    const int ref_idx = b.ref_idx, is_luma = !b.c_idx, scale4 = (lw + lh - 2) >> 2;
    const uint16_t *mainr = (vertical ? top : left) - (1 + ref_idx), *side_ref = vertical ? left : top;
    for (int idx = tid; idx < w * h; idx += kThreads) {
        const int y = idx >> lw, x = idx & (w - 1);
        int pred;
        if (b.kind == VVC_CUDA_INTRA_PLANAR) {
            const int pv = ((h - 1 - y) * top[x] + (y + 1) * left[h]) << lw;
            const int ph = ((w - 1 - x) * left[y] + (x + 1) * top[w]) << lh;
            pred = (pv + ph + w * h) >> (lw + lh + 1);
        } else if (b.kind == VVC_CUDA_INTRA_DC) {
            pred = sc.dc;
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
        if (b.pdpc4) {
            int l, t, wl, wt;
            if (b.kind == VVC_CUDA_INTRA_PLANAR || b.kind == VVC_CUDA_INTRA_DC) {
                l = left[y]; t = top[x];
                wl = 32 >> min((x << 1) >> scale4, 31);
                wt = 32 >> min((y << 1) >> scale4, 31);
            } else {
                l = (int)left[y] - (int)left[-1] + pred; t = (int)top[x] - (int)top[-1] + pred;
                wl = b.kind == VVC_CUDA_INTRA_VERT ? 32 >> min((x << 1) >> scale4, 31) : 0;
                wt = b.kind == VVC_CUDA_INTRA_HORZ ? 32 >> min((y << 1) >> scale4, 31) : 0;
            }
            pred = d_clip3(pred + ((wl * (l - pred) + wt * (t - pred) + 32) >> 6), 0, maxv);
        }
        dst[(long long)y * pitch + x] = (pel)pred;
    }
}

__global__ void __launch_bounds__(kThreads) intra_leaf_kernel(const IntraK p)
{
    __shared__ LeafScratch sc;
    for (int ri = blockIdx.x; ri < p.n; ri += gridDim.x) {
        const VVCCudaIntraPB b = p.pbs[ri];
        const int pitch = SEL3(p.pitch, b.c_idx);
        pel *dst = SEL3(p.plane, b.c_idx) + b.pic * SEL3(p.bstride, b.c_idx) + (long long)b.y0 * pitch + b.x0;
        Leaf lf;
        lf.kind = b.kind; lf.mode = b.mode; lf.ref_idx = b.ref_idx; lf.filter_flag = b.filter_flag; lf.flags = b.flags;
        lf.w = b.w; lf.h = b.h; lf.c_idx = b.c_idx; lf.pdpc4 = 0;
        __syncthreads();                                   // previous record is done with shared memory
        leaf_predict(lf, p.edges + b.top, p.edges + b.left, dst, pitch, p.bd, sc);
    }
}

// ---- intra_pred with the reference lines prepared here, and CCLM --------------------------------------------------
// Replaces intra.intra_pred (vvc_intra_template.c:595-683: prepare_intra_edge_params :467-592, ref_filter :450-465,
// the PDPC of planar / DC / H / V :653-681), ff_vvc_wide_angle_mode_mapping / ff_vvc_need_pdpc /
// ff_vvc_ref_filter_flag_derive (vvc_intra.c:557-573, :655-659, :693-715) and intra.intra_cclm_pred (:29-375).
//
// A CTA takes a block.  Every sample of the two reference lines is a pure function of the record (which picture sample,
// or which substitute) - the reference's running fills (copy, then extend, then patch the corner) collapse into one
// expression per index, so the lines are filled by all threads with one global load each and no ordering; smoothing and
// the projected part of an angular reference are two more barrier-separated passes over shared memory, then the leaf
// predictor above runs from shared memory.  CCLM: eight threads fetch the selected neighbour positions (down-sampled
// luma + both chroma samples), one thread derives a / b / k, all threads down-sample the block's luma on the fly.
constexpr int E_NEG = 80, E_LEN = E_NEG + 192;     // indices -(64 + 3) .. 2 * 64 + 16 * 2 + 1

struct PredK {
    pel       *plane[3];
    int        pitch[3];
    long long  bstride[3];
    int        bd, hs, vs, ctb_log2;
    const VVCCudaIntraBlk *blks;
    int        n;
};

struct CclmShared {
    int sel[3][8];
    int a[2], b[2], k[2];
};

__device__ __forceinline__ bool smoothing_mode(int m)
{
    return m == -14 || m == -12 || m == -10 || m == -6 || m == 0 || m == 2 || m == 34 || m == 66 || m == 72 || m == 76 || m == 78 || m == 80;
}

template <bool CG>
__device__ void cclm_block(const PredK &p, const VVCCudaIntraBlk &b, CclmShared &cs)
{
    const int tid = threadIdx.x, bd = p.bd, hs = p.hs, vs = p.vs, w = b.w, h = b.h, x = b.x0, y = b.y0;
    const int x0 = x << hs, y0 = y << vs;
    const int at = (b.flags & VVC_CUDA_INTRA_F_LUMA_AVAIL_T) != 0, al = (b.flags & VVC_CUDA_INTRA_F_LUMA_AVAIL_L) != 0;
    const bool colloc = b.flags & VVC_CUDA_INTRA_F_COLLOCATED;
    const pel *Y = p.plane[0] + b.pic * p.bstride[0];
    const int yp = p.pitch[0], cp = p.pitch[1];
    pel *C0 = p.plane[1] + b.pic * p.bstride[1], *C1 = p.plane[2] + b.pic * p.bstride[2];
    auto L = [&](int xx, int yy) -> int { return ld_pel<CG>(Y + (long long)yy * yp + xx); };
    if (!at && !al) {
        for (int idx = tid; idx < w * h; idx += kThreads) {
            const int i = idx / w, j = idx - i * w;
            C0[(long long)(y + i) * cp + x + j] = (pel)(1 << (bd - 1));
            C1[(long long)(y + i) * cp + x + j] = (pel)(1 << (bd - 1));
        }
        return;
    }
    // luma at (chroma column j, chroma row i) of the block, at chroma resolution (cclm_get_luma_rec_pixels :282-335)
    auto down = [&](int j, int i) -> int {
        const int lx = x0 + (j << hs), ly = y0 + (i << vs);
        if (!hs && !vs)
            return L(lx, ly);
        const int xl = (j || al) ? lx - 1 : lx;
        if (!vs)
            return (L(xl, ly) + 2 * L(lx, ly) + L(lx + 1, ly) + 2) >> 2;
        if (colloc)
            return (L(xl, ly) + L(lx, (i || at) ? ly - 1 : ly) + 4 * L(lx, ly) + L(lx + 1, ly) + L(lx, ly + 1) + 4) >> 3;
        return (L(xl, ly) + L(xl, ly + 1) + 2 * L(lx, ly) + 2 * L(lx, ly + 1) + L(lx + 1, ly) + L(lx + 1, ly + 1) + 4) >> 3;
    };
    // selected neighbour positions (cclm_get_select_pos :61-88)
    const int mode = b.pred_mode, is4 = !at || !al || mode != 81;
    int n[2];
    if (mode == 81) { n[0] = at ? w : 0; n[1] = al ? h : 0; }
    else {
        n[0] = (at && mode == 83) ? min(w + min(w, h), (int)b.avail_top) : 0;
        n[1] = (al && mode == 82) ? min(h + min(w, h), (int)b.avail_left) : 0;
    }
    const int cnt0 = min(n[0], (1 + is4) << 1), cnt1 = min(n[1], (1 + is4) << 1);
    if (tid < cnt0 + cnt1) {
        const int s = tid >= cnt0, c = tid - (s ? cnt0 : 0);
        const int pos = (n[s] >> (2 + is4)) + c * max(1, n[s] >> (1 + is4));
        int v;
        if (!s) {                                              // above (cclm_select_luma :99-139)
            const int px = pos << hs, lx = x0 + px;
            if (!hs && !vs) {
                v = L(x0 + pos, y0 - at);
            } else {
                const int xl = (px || al) ? lx - 1 : lx;
                const bool ctu_top = !(y0 & ((1 << p.ctb_log2) - 1));
                if (vs && !ctu_top) {
                    const int ly = y0 - 2;
                    v = colloc ? (L(lx, ly - 1) + L(xl, ly) + 4 * L(lx, ly) + L(lx + 1, ly) + L(lx, ly + 1) + 4) >> 3
                               : (L(xl, ly) + L(xl, ly + 1) + 2 * (L(lx, ly) + L(lx, ly + 1)) + L(lx + 1, ly) + L(lx + 1, ly + 1) + 4) >> 3;
                } else {
                    v = (L(xl, y0 - 1) + 2 * L(lx, y0 - 1) + L(lx + 1, y0 - 1) + 2) >> 2;
                }
            }
            cs.sel[1][tid] = ld_pel<CG>(C0 + (long long)(y - 1) * cp + x + pos);
            cs.sel[2][tid] = ld_pel<CG>(C1 + (long long)(y - 1) * cp + x + pos);
        } else {                                               // left (:141-166)
            if (!hs && !vs) {
                v = L(x0 - al, y0 + pos);
            } else {
                const int ly = y0 + (pos << vs), lx = x0 - (1 + hs) * al, xl = lx - al;
                if (!vs)         v = (L(xl, ly) + 2 * L(lx, ly) + L(lx + 1, ly) + 2) >> 2;
                else if (colloc) v = (L(xl, ly) + L(lx, (pos || at) ? ly - 1 : ly) + 4 * L(lx, ly) + L(lx + 1, ly) + L(lx, ly + 1) + 4) >> 3;
                else             v = (L(xl, ly) + L(xl, ly + 1) + 2 * L(lx, ly) + 2 * L(lx, ly + 1) + L(lx + 1, ly) + L(lx + 1, ly + 1) + 4) >> 3;
            }
            cs.sel[1][tid] = ld_pel<CG>(C0 + (long long)(y + pos) * cp + x - 1);
            cs.sel[2][tid] = ld_pel<CG>(C1 + (long long)(y + pos) * cp + x - 1);
        }
        cs.sel[0][tid] = v;
    }
    __syncthreads();
    if (tid == 0) {
        int a[2] = { 0, 0 }, k[2] = { 0, 0 }, bb[2] = { 1 << (bd - 1), 1 << (bd - 1) };
        if (cnt0 + cnt1) {
            if (cnt0 + cnt1 == 2)
                for (int c = 0; c < 3; c++) {
                    cs.sel[c][3] = cs.sel[c][0]; cs.sel[c][2] = cs.sel[c][1]; cs.sel[c][0] = cs.sel[c][1]; cs.sel[c][1] = cs.sel[c][3];
                }
            // the two smallest / two largest luma values, by the reference's compare-exchange network (:203-226)
            int lo0 = 0, lo1 = 2, hi0 = 1, hi1 = 3, t;
            if (cs.sel[0][lo0] > cs.sel[0][lo1]) { t = lo0; lo0 = lo1; lo1 = t; }
            if (cs.sel[0][hi0] > cs.sel[0][hi1]) { t = hi0; hi0 = hi1; hi1 = t; }
            if (cs.sel[0][lo0] > cs.sel[0][hi1]) { t = lo0; lo0 = hi0; hi0 = t; t = lo1; lo1 = hi1; hi1 = t; }
            if (cs.sel[0][lo1] > cs.sel[0][hi0]) { t = lo1; lo1 = hi0; hi0 = t; }
            int mn[3], mx[3];
            for (int c = 0; c < 3; c++) {
                mx[c] = (cs.sel[c][hi0] + cs.sel[c][hi1] + 1) >> 1;
                mn[c] = (cs.sel[c][lo0] + cs.sel[c][lo1] + 1) >> 1;
            }
            const int diff = mx[0] - mn[0];
            for (int c = 0; c < 2; c++) {
                if (!diff) { a[c] = k[c] = 0; bb[c] = mn[c + 1]; continue; }
                const int diffc = mx[c + 1] - mn[c + 1];
                int lx = d_ilog2(diff);
                const int norm = ((diff << 4) >> lx) & 15;
                lx += norm ? 1 : 0;
                const int ly = abs(diffc) > 0 ? d_ilog2(abs(diffc)) + 1 : 0;
                // div_sig_table (:248) as a packed constant: 0 7 6 5 5 4 4 3 3 2 2 1 1 1 1 0
                const int v = (int)((0x0111122334455670ull >> (4 * norm)) & 15) | 8;
                a[c] = (diffc * v + ((1 << ly) >> 1)) >> ly;
                k[c] = max(1, 3 + lx - ly);
                if (3 + lx - ly < 1)
                    a[c] = d_sign(a[c]) * 15;
                bb[c] = mn[c + 1] - ((a[c] * mn[0]) >> k[c]);
            }
        }
        for (int c = 0; c < 2; c++) { cs.a[c] = a[c]; cs.b[c] = bb[c]; cs.k[c] = k[c]; }
    }
    __syncthreads();
    const int a0 = cs.a[0], a1 = cs.a[1], b0 = cs.b[0], b1 = cs.b[1], k0 = cs.k[0], k1 = cs.k[1];
    for (int idx = tid; idx < w * h; idx += kThreads) {
        const int i = idx / w, j = idx - i * w, d = down(j, i);
        C0[(long long)(y + i) * cp + x + j] = (pel)d_clip_pel(((d * a0) >> k0) + b0, bd);
        C1[(long long)(y + i) * cp + x + j] = (pel)d_clip_pel(((d * a1) >> k1) + b1, bd);
    }
}

// One record by the whole CTA (the caller has synchronised: shared memory is free).
template <bool CG>
__device__ void predict_record(const PredK &p, const VVCCudaIntraBlk &b, LeafScratch &sc, CclmShared &cs, uint16_t (*s_line)[E_LEN])
{
    const int tid = threadIdx.x, bd = p.bd;
    if (b.kind == VVC_CUDA_INTRA_KIND_CCLM) {
        cclm_block<CG>(p, b, cs);
        return;
    }
    const int c_idx = b.c_idx, w = b.w, h = b.h, x0 = b.x0, y0 = b.y0;
    const int pitch = SEL3(p.pitch, c_idx);
    const pel *pic = SEL3(p.plane, c_idx) + b.pic * SEL3(p.bstride, c_idx);
    const bool is_mip = b.kind == VVC_CUDA_INTRA_KIND_MIP, isp = (b.flags & VVC_CUDA_INTRA_F_ISP) && !c_idx;
    const int ref_idx = c_idx ? 0 : b.ref_idx, rl = -1 - ref_idx;
    int mode = 0;
    if (!is_mip) {                                     // ff_vvc_wide_angle_mode_mapping
        const int nw = isp ? b.cb_w : w, nh = isp ? b.cb_h : h;
        const int ratio = abs(d_ilog2(nw) - d_ilog2(nh));
        const int hi = ratio > 1 ? 8 + 2 * ratio : 8, lo = ratio > 1 ? 60 - 2 * ratio : 60;
        mode = b.pred_mode;
        if (nw > nh && mode >= 2 && mode < hi)         mode += 65;
        else if (nh > nw && mode <= 66 && mode > lo)   mode -= 67;
    }
    const bool non_angular = mode == 0 || mode == 1 || mode == 18 || mode == 50;
    int angle = 0, inv_angle = 0;
    if (!is_mip && !non_angular) {
        angle = pred_angle(mode);
        inv_angle = angle > 0 ? (32768 + angle) / (2 * angle) : -((32768 - angle) / (-2 * angle));
    }
    bool pdpc = false;                                 // ff_vvc_need_pdpc
    if (!is_mip && w >= 4 && h >= 4 && !ref_idx && !(b.flags & VVC_CUDA_INTRA_F_BDPCM)) {
        if (non_angular)                  pdpc = true;
        else if (mode > 18 && mode < 50)  pdpc = false;
        else                              pdpc = min(2, d_ilog2(mode >= 50 ? h : w) - d_ilog2(3 * inv_angle - 2) + 8) >= 0;
    }
    const bool rff = !is_mip && smoothing_mode(mode);
    const bool smooth = rff && !ref_idx && w * h > 32 && !c_idx && !(b.flags & VVC_CUDA_INTRA_F_ISP);
    int n_left, n_top, refw = 0, refh = 0;
    if (is_mip || mode == 0)  { n_left = h + 1 + smooth; n_top = w + 1 + smooth; }
    else if (mode == 1)       { n_left = h; n_top = w; }
    else if (mode == 50)      { n_left = pdpc ? h : 1; n_top = w; }
    else if (mode == 18)      { n_left = h; n_top = pdpc ? w : 1; }
    else {
        refw = isp ? b.cb_w + w : 2 * w;
        refh = isp ? b.cb_h + h : 2 * h;
        n_top = refw; n_left = refh;
    }
    const int got_l = min(n_left, (int)b.avail_left), got_t = min(n_top, (int)b.avail_top);
    const bool up_left = b.flags & VVC_CUDA_INTRA_F_UP_LEFT;
    auto PIC = [&](int xx, int yy) -> int { return ld_pel<CG>(pic + (long long)(y0 + yy) * pitch + x0 + xx); };
    // what the corner samples (negative indices) hold when the up-left block is not available
    const int corner = got_l ? PIC(rl, 0) : got_t ? PIC(0, rl) : 1 << (bd - 1);
    uint16_t *left = s_line[0] + E_NEG, *top = s_line[1] + E_NEG;
    for (int i = rl + tid; i < n_left; i += kThreads)
        left[i] = (uint16_t)(i < 0 ? (up_left ? PIC(rl, i) : corner)
                                   : i < got_l ? PIC(rl, i) : got_l ? PIC(rl, got_l - 1) : (up_left ? PIC(rl, -1) : corner));
    for (int i = rl + tid; i < n_top; i += kThreads)
        top[i] = (uint16_t)(i < 0 ? (up_left ? PIC(i, rl) : corner)
                                  : i < got_t ? PIC(i, rl) : got_t ? PIC(got_t - 1, rl) : (up_left ? PIC(-1, rl) : corner));
    __syncthreads();
    if (smooth) {                                      // ref_filter: [1 2 1], the angular modes keep their last sample
        const int keep_last = !(is_mip || mode == 0);
        uint16_t *fl = s_line[2] + E_NEG, *ft = s_line[3] + E_NEG;
        for (int i = tid - 1; i < n_left; i += kThreads)
            fl[i] = (uint16_t)(i < 0 ? (left[0] + 2 * left[-1] + top[0] + 2) >> 2
                                     : (keep_last && i == n_left - 1) ? left[i] : (left[i - 1] + 2 * left[i] + left[i + 1] + 2) >> 2);
        for (int i = tid - 1; i < n_top; i += kThreads)
            ft[i] = (uint16_t)(i < 0 ? (left[0] + 2 * left[-1] + top[0] + 2) >> 2
                                     : (keep_last && i == n_top - 1) ? top[i] : (top[i - 1] + 2 * top[i] + top[i + 1] + 2) >> 2);
        __syncthreads();
        left = fl; top = ft;
    }
    Leaf lf;
    lf.w = w; lf.h = h; lf.c_idx = c_idx; lf.ref_idx = ref_idx; lf.filter_flag = 0; lf.flags = 0; lf.pdpc4 = 0;
    if (is_mip) {
        lf.kind = VVC_CUDA_INTRA_MIP; lf.mode = b.pred_mode;
        lf.flags = (b.flags & VVC_CUDA_INTRA_F_MIP_TRANSP) ? VVC_CUDA_INTRA_MIP_TRANSPOSED : 0;
    } else {
        lf.kind = mode == 0 ? VVC_CUDA_INTRA_PLANAR : mode == 1 ? VVC_CUDA_INTRA_DC : mode == 50 ? VVC_CUDA_INTRA_VERT
                : mode == 18 ? VVC_CUDA_INTRA_HORZ : mode >= 34 ? VVC_CUDA_INTRA_ANGULAR_V : VVC_CUDA_INTRA_ANGULAR_H;
        lf.mode = mode;
        lf.flags = (pdpc && !non_angular) ? VVC_CUDA_INTRA_PDPC : 0;
        lf.pdpc4 = pdpc && non_angular;
        if (mode != 0 && mode != 1 && !c_idx && !(rff || ref_idx || (b.flags & VVC_CUDA_INTRA_F_ISP))) {
            const int dist = min(abs(mode - 50), abs(mode - 18)), ntbs = (d_ilog2(w) + d_ilog2(h)) >> 1;
            lf.filter_flag = dist > (ntbs == 2 ? 24 : ntbs == 3 ? 14 : ntbs == 4 ? 2 : 0);
        }
        if (!non_angular) {
            // the main reference beyond the fetched samples: projected from the other line (negative angles) or the
            // last sample repeated (positive angles)
            const bool vertical = mode >= 34;
            uint16_t *mainr = vertical ? top : left;
            const uint16_t *side = vertical ? left : top;
            const int n_main = vertical ? refw : refh, across = vertical ? h : w, along = vertical ? w : h;
            if (angle < 0) {
                for (int k = -across + tid; k < 0; k += kThreads)
                    mainr[k - (ref_idx + 1)] = side[-1 - ref_idx + min((k * inv_angle + 256) >> 9, across)];
            } else {
                const int last = n_main + max(1, along / across) * ref_idx + 1;
                for (int i = n_main + tid; i <= last; i += kThreads)
                    mainr[i] = mainr[n_main - 1];
            }
            __syncthreads();
        }
    }
    pel *dst = SEL3(p.plane, c_idx) + b.pic * SEL3(p.bstride, c_idx) + (long long)y0 * pitch + x0;
    leaf_predict(lf, top, left, dst, pitch, bd, sc);
}

__global__ void __launch_bounds__(kThreads) intra_pred_kernel(const PredK p)
{
    __shared__ LeafScratch sc;
    __shared__ CclmShared cs;
    __shared__ uint16_t s_line[4][E_LEN];                  // raw left, raw top, smoothed left, smoothed top
    for (int ri = blockIdx.x; ri < p.n; ri += gridDim.x) {
        const VVCCudaIntraBlk b = p.blks[ri];
        __syncthreads();                                   // previous record is done with shared memory
        predict_record<false>(p, b, sc, cs, s_line);
    }
}

// ---- all-intra reconstruction driven by dependencies, one launch per picture ring ---------------------------------------
// The host lists the steps of a decoder in decoding order - per coding unit a luma step and a chroma step, each "predict
// these blocks, then add the residual of these transform blocks" (what predict_intra + itransform do per TU,
// vvc_intra.c:233-281, 432-478).  Persistent CTAs draw steps from a counter in that order.  Before predicting a block a
// CTA waits until the 4x4 units holding the samples the block may read - the availability counts of its record, i.e. what
// the reference's reconstructed-area list answers - are marked done in a per-picture progress map; after the residual it
// fences and marks its own units.  Every unit a step waits for belongs to an earlier step, and every drawn step is held by
// a resident CTA (the grid never exceeds what the device keeps resident), so the earliest unfinished step can always
// run: no deadlock, no global barrier, and the critical path is the dependency chain itself (about a thousand steps for
// a 1080p picture) instead of two kernel launches per wavefront.  Samples written by other CTAs are read through L2.
struct DagK {
    PredK          pk;
    ItxK           ik;
    const int32_t *blk_end, *tb_end;    // running totals per step (device memory)
    int            n_steps;
    uint32_t      *counter;             // [0] next step, [1] error flag
    uint8_t       *done[2];             // progress maps per plane type (luma, chroma), [batch][uh][uw] of 4x4 luma units
    int            uw, uh;
};

constexpr unsigned kSpinLimit = 1u << 24;     // polls (with back-off) before a step gives up: malformed availability counts

// wait until the units [ux0, ux1) x [uy0, uy1) of map `m` are done; returns false when the watchdog fires
__device__ __forceinline__ bool wait_units(const DagK &d, const uint8_t *m, int ux0, int uy0, int ux1, int uy1)
{
    ux0 = max(ux0, 0); uy0 = max(uy0, 0); ux1 = min(ux1, d.uw); uy1 = min(uy1, d.uh);
    const int nx = ux1 - ux0, n = nx > 0 && uy1 > uy0 ? nx * (uy1 - uy0) : 0;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const volatile uint8_t *f = m + (size_t)(uy0 + i / nx) * d.uw + ux0 + i % nx;
        unsigned spins = 0;
        while (!*f) {
            __nanosleep(64);
            if (++spins > kSpinLimit || *reinterpret_cast<volatile uint32_t *>(d.counter + 1)) {
                atomicExch(d.counter + 1, 1u);
                return false;
            }
        }
    }
    return true;
}

__device__ bool wait_for(const DagK &d, const VVCCudaIntraBlk &b)
{
    const int hs = d.pk.hs, vs = d.pk.vs;
    const size_t pic = (size_t)b.pic * d.uw * d.uh;
    const int ch = b.c_idx > 0;
    const int sx = ch ? hs : 0, sy = ch ? vs : 0;                    // plane sample -> luma sample
    const uint8_t *m = d.done[ch] + pic;
    // samples of the block's own plane type: left column, top row, corner - as far as the record calls them available
    const int x = b.x0, y = b.y0;
    // ... and no further than any predictor reaches (2 w, or cb_w + w for an ISP part, + 2 for planar with smoothing), so
    // that step orders other than the decoding order stay free of waits on samples the block never reads
    const int nl = min((int)b.avail_left, max(2 * b.h, (b.cb_h >> sy) + b.h) + 2);
    const int nt = min((int)b.avail_top, max(2 * b.w, (b.cb_w >> sx) + b.w) + 2);
    bool ok = true;
    if (nl > 0) ok &= wait_units(d, m, ((x - 1) << sx) >> 2, (y << sy) >> 2, (((x - 1) << sx) >> 2) + 1, (((y + nl - 1) << sy) >> 2) + 1);
    if (nt > 0) ok &= wait_units(d, m, (x << sx) >> 2, ((y - 1) << sy) >> 2, (((x + nt - 1) << sx) >> 2) + 1, (((y - 1) << sy) >> 2) + 1);
    if ((b.flags & VVC_CUDA_INTRA_F_UP_LEFT) && x > 0 && y > 0)
        ok &= wait_units(d, m, ((x - 1) << sx) >> 2, ((y - 1) << sy) >> 2, (((x - 1) << sx) >> 2) + 1, (((y - 1) << sy) >> 2) + 1);
    if (b.kind == VVC_CUDA_INTRA_KIND_CCLM) {
        // the luma it reads: the coding unit's own luma, and the neighbouring luma rows / columns its flags call available
        const uint8_t *ml = d.done[0] + pic;
        const int x0 = x << hs, y0 = y << vs, wl = b.w << hs, hl = b.h << vs;
        const bool at = b.flags & VVC_CUDA_INTRA_F_LUMA_AVAIL_T, al = b.flags & VVC_CUDA_INTRA_F_LUMA_AVAIL_L;
        ok &= wait_units(d, ml, x0 >> 2, y0 >> 2, (x0 + wl + 3) >> 2, (y0 + hl + 3) >> 2);
        if (al) {
            const int rows = max(hl, min(b.h + min((int)b.w, (int)b.h), (int)b.avail_left) << vs);
            ok &= wait_units(d, ml, (x0 - 1) >> 2, y0 >> 2, ((x0 - 1) >> 2) + 1, (y0 + rows + 3) >> 2);
        }
        if (at) {
            const int cols = max(wl, min(b.w + min((int)b.w, (int)b.h), (int)b.avail_top) << hs);
            ok &= wait_units(d, ml, x0 >> 2, (y0 - 1) >> 2, (x0 + cols + 3) >> 2, ((y0 - 1) >> 2) + 1);
        }
        if (al && at)
            ok &= wait_units(d, ml, (x0 - 1) >> 2, (y0 - 1) >> 2, ((x0 - 1) >> 2) + 1, ((y0 - 1) >> 2) + 1);
        // INTRA_LT_CCLM picks its chroma neighbours by the LUMA flags (w above, h to the left: cclm_get_select_pos :68-71)
        if (b.pred_mode == 81) {
            if (al && nl < b.h) ok &= wait_units(d, m, ((x - 1) << sx) >> 2, (y << sy) >> 2, (((x - 1) << sx) >> 2) + 1, (((y + b.h - 1) << sy) >> 2) + 1);
            if (at && nt < b.w) ok &= wait_units(d, m, (x << sx) >> 2, ((y - 1) << sy) >> 2, (((x + b.w - 1) << sx) >> 2) + 1, (((y - 1) << sy) >> 2) + 1);
        }
    }
    return ok;
}

__device__ __forceinline__ void mark_done(const DagK &d, const VVCCudaIntraBlk &b)
{
    const int ch = b.c_idx > 0, sx = ch ? d.pk.hs : 0, sy = ch ? d.pk.vs : 0;
    uint8_t *m = d.done[ch] + (size_t)b.pic * d.uw * d.uh;
    const int ux0 = (b.x0 << sx) >> 2, uy0 = (b.y0 << sy) >> 2;
    const int ux1 = min((((b.x0 + b.w) << sx) + 3) >> 2, d.uw), uy1 = min((((b.y0 + b.h) << sy) + 3) >> 2, d.uh), nx = ux1 - ux0;
    for (int i = threadIdx.x; i < nx * (uy1 - uy0); i += kThreads)
        *reinterpret_cast<volatile uint8_t *>(m + (size_t)(uy0 + i / nx) * d.uw + ux0 + i % nx) = 1;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) intra_dag_kernel(const DagK d)
{
    __shared__ LeafScratch sc;
    __shared__ CclmShared cs;
    __shared__ uint16_t s_line[4][E_LEN];
    __shared__ alignas(16) int s_buf[2 * 64 * 65];
    __shared__ int s_step;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0)
            s_step = (int)atomicAdd(d.counter, 1u);
        __syncthreads();
        const int step = s_step;
        if (step >= d.n_steps)
            return;
        const int b0 = step ? __ldg(d.blk_end + step - 1) : 0, b1 = __ldg(d.blk_end + step);
        const int t0 = step ? __ldg(d.tb_end + step - 1) : 0, t1 = __ldg(d.tb_end + step);
        // the step is on the pictures' critical path from the moment its neighbours are done: bring its coefficients from
        // HBM into L2 while it waits for them
        for (int ti = t0; ti < t1; ti++) {
            const VVCCudaTB tb = d.ik.tbs[ti];
            const char *c = (MODE & 1) ? (const char *)(d.ik.src.window + tb.coeff_offset) : (const char *)(d.ik.src.dense + tb.coeff_offset);
            const int bytes = (MODE & 1) ? tb.nzw * tb.nzh * 2 : (tb.nzh << tb.log2_w) * 4;
            for (int o = threadIdx.x * 128; o < bytes; o += kThreads * 128)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(c + o));
        }
        for (int bi = b0; bi < b1; bi++) {
            const VVCCudaIntraBlk b = d.pk.blks[bi];
            const bool ok = wait_for(d, b);
            __threadfence();                                   // the flags were seen: the samples behind them are next
            if (!__syncthreads_and(ok))
                return;                                        // watchdog: the error flag is set, every CTA leaves at its next poll
            predict_record<true>(d.pk, b, sc, cs, s_line);
            __syncthreads();
        }
        for (int ti = t0; ti < t1; ti++) {
            const VVCCudaTB tb = d.ik.tbs[ti];
            process_tb<kThreads, MODE>(d.ik, tb, ti, s_buf, s_buf + 64 * 65, threadIdx.x);
        }
        __threadfence();                                       // this step's samples before its flags
        __syncthreads();
        for (int bi = b0; bi < b1; bi++)
            mark_done(d, d.pk.blks[bi]);
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

static int intra_pred_check(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraBlk *blks, int n_blks)
{
    if (!frame || (n_blks > 0 && !blks))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "intra_pred: null argument");
    if (frame->bit_depth != 10 && frame->bit_depth != 12)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "intra_pred: bit depth %d not accelerated (10/12 only)", frame->bit_depth);
    if (frame->chroma_format_idc && !((frame->hshift == 1 && frame->vshift <= 1) || (!frame->hshift && !frame->vshift)))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "intra_pred: chroma format (hshift %d, vshift %d) is none of 4:2:0 / 4:2:2 / 4:4:4",
                            frame->hshift, frame->vshift);
    return 0;
}

static int intra_pred_launch(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraBlk *blks, int n_blks)
{
    PredK p;
    for (int c = 0; c < 3; c++) {
        p.plane[c] = (pel *)frame->data[c];
        p.pitch[c] = (int)(frame->stride[c] / 2);
        p.bstride[c] = frame->batch_stride[c] / 2;
    }
    p.bd = frame->bit_depth; p.hs = frame->hshift; p.vs = frame->vshift; p.ctb_log2 = frame->ctb_log2;
    p.blks = blks; p.n = n_blks;
    intra_pred_kernel<<<n_blks < 148 * 8 ? n_blks : 148 * 8, kThreads, 0, ctx->stream>>>(p);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_intra_pred_frame(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraBlk *blks, int n_blks)
{
    if (ctx->err)
        return ctx->err;
    if (intra_pred_check(ctx, frame, blks, n_blks))
        return ctx->err;
    if (n_blks <= 0)
        return VVC_CUDA_OK;
    return intra_pred_launch(ctx, frame, blks, n_blks);
}

extern "C" int vvc_cuda_intra_pred_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraBlk *blks, int n_blks)
{
    if (ctx->err)
        return ctx->err;
    if (intra_pred_check(ctx, frame, blks, n_blks))
        return ctx->err;
    const size_t fsz = align_up(vvc_stage_frame_size(frame), 256);
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, fsz + align_up((size_t)(n_blks > 0 ? n_blks : 1) * sizeof(VVCCudaIntraBlk), 256));
    if (!base)
        return ctx->err;
    VVCCudaFrame d;
    vvc_stage_frame_layout(frame, base, &d);
    VVCCudaIntraBlk *db = (VVCCudaIntraBlk *)(base + fsz);
    if (vvc_stage_frame_h2d(ctx, &d, frame))           // the reference samples come from the picture: it travels both ways
        return ctx->err;
    if (n_blks > 0) {
        VVC_TRY(ctx, cudaMemcpyAsync(db, blks, (size_t)n_blks * sizeof(VVCCudaIntraBlk), cudaMemcpyHostToDevice, ctx->stream));
        if (intra_pred_launch(ctx, &d, db, n_blks))
            return ctx->err;
    }
    if (vvc_stage_frame_d2h(ctx, frame, &d))
        return ctx->err;
    return vvc_cuda_sync(ctx);
}

// itx.cu: the generic residual kernel alone over a short list (one launch)
int vvc_itx_launch_short(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaCoeffs *co, const VVCCudaTB *tbs, int n_tbs, int range);

extern "C" int vvc_cuda_intra_recon_frame(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraBlk *blks, const int32_t *blk_end,
                                          const VVCCudaCoeffs *coeffs, const VVCCudaTB *tbs, const int32_t *tb_end, int n_waves,
                                          int log2_transform_range)
{
    if (ctx->err)
        return ctx->err;
    if (!blk_end || !tb_end || n_waves < 0 || (n_waves > 0 && tb_end[n_waves - 1] > 0 && (!coeffs || !tbs)))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "intra_recon: null argument");
    if (intra_pred_check(ctx, frame, blks, n_waves > 0 ? blk_end[n_waves - 1] : 0))
        return ctx->err;
    int b0 = 0, t0 = 0;
    for (int g = 0; g < n_waves; g++) {
        if (blk_end[g] < b0 || tb_end[g] < t0)
            return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "intra_recon: wave %d ends before it starts", g);
        if (blk_end[g] > b0 && intra_pred_launch(ctx, frame, blks + b0, blk_end[g] - b0))
            return ctx->err;
        if (tb_end[g] > t0) {
            VVCCudaCoeffs c = *coeffs;
            if (c.quant)
                c.quant += t0;
            if (vvc_itx_launch_short(ctx, frame, &c, tbs + t0, tb_end[g] - t0, log2_transform_range))
                return ctx->err;
        }
        b0 = blk_end[g]; t0 = tb_end[g];
    }
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_intra_recon_frame_ordered(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraBlk *blks,
                                                  const int32_t *blk_end, const VVCCudaCoeffs *co, const VVCCudaTB *tbs,
                                                  const int32_t *tb_end, int n_steps, int n_blks, int n_tbs, int range)
{
    if (ctx->err)
        return ctx->err;
    if (!blk_end || !tb_end || n_steps < 0 || n_blks < 0 || n_tbs < 0 || (n_tbs > 0 && (!co || !co->data || !tbs)))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "intra_recon_ordered: null argument");
    if (intra_pred_check(ctx, frame, blks, n_blks))
        return ctx->err;
    if (n_tbs > 0 && (range < 15 || range > 20 || (co->format != VVC_CUDA_COEFF_DENSE32 && co->format != VVC_CUDA_COEFF_WINDOW16) ||
                      (co->format == VVC_CUDA_COEFF_WINDOW16 && range != 15) || !frame_vec_ok(frame)))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "intra_recon_ordered: unsupported picture or coefficient format");
    if ((frame->width & 3) || (frame->height & 3))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "intra_recon_ordered: picture size must be a multiple of 4");
    if (!n_steps)
        return VVC_CUDA_OK;
    DagK d;
    for (int c = 0; c < 3; c++) {
        d.pk.plane[c] = d.ik.plane[c] = (pel *)frame->data[c];
        d.pk.pitch[c] = d.ik.pitch[c] = (int)(frame->stride[c] / 2);
        d.pk.bstride[c] = d.ik.bstride[c] = frame->batch_stride[c] / 2;
    }
    d.pk.bd = frame->bit_depth; d.pk.hs = frame->hshift; d.pk.vs = frame->vshift; d.pk.ctb_log2 = frame->ctb_log2;
    d.pk.blks = blks; d.pk.n = n_blks;
    const int mode = n_tbs > 0 ? coef_mode(co) : 0;
    memset(&d.ik.src, 0, sizeof(d.ik.src));
    if (n_tbs > 0) {
        d.ik.src.dense = (mode & 1) ? nullptr : (const int32_t *)co->data;
        d.ik.src.window = (mode & 1) ? (const int16_t *)co->data : nullptr;
        d.ik.src.quant = co->quant; d.ik.src.scaling = co->scaling; d.ik.src.lmcs_scales = co->lmcs_scales;
    }
    d.ik.src.range = range; d.ik.src.bd = frame->bit_depth;
    d.ik.store = (n_tbs > 0 && !(mode & 1)) ? (int32_t *)co->data : nullptr;
    d.ik.tbs = tbs; d.ik.n_tbs = n_tbs; d.ik.range = range; d.ik.bd = frame->bit_depth;
    d.ik.list = d.ik.list_count = NULL;
    d.blk_end = blk_end; d.tb_end = tb_end; d.n_steps = n_steps;
    d.uw = frame->width >> 2; d.uh = frame->height >> 2;
    const size_t map = (size_t)d.uw * d.uh * frame->batch;
    uint8_t *scratch = (uint8_t *)vvc_ctx_scratch(ctx, 4, 256 + 2 * map);
    if (!scratch)
        return ctx->err;
    d.counter = (uint32_t *)scratch; d.done[0] = scratch + 256; d.done[1] = scratch + 256 + map;
    VVC_TRY(ctx, cudaMemsetAsync(scratch, 0, 256 + 2 * map, ctx->stream));
    // persistent grid: never more CTAs than the device keeps resident (a waiting CTA must not block one that has not started)
    void (*kern)(const DagK) = mode == 0 ? intra_dag_kernel<0> : mode == 1 ? intra_dag_kernel<1> : mode == 2 ? intra_dag_kernel<2> : intra_dag_kernel<3>;
    int per_sm = 0, sms = 0;
    VVC_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0));
    VVC_TRY(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    if (per_sm < 1)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_CUDA, "intra_recon_ordered: the kernel does not fit an SM");
    int grid = n_steps < per_sm * sms ? n_steps : per_sm * sms;
    if (const char *g = getenv("VVC_CUDA_INTRA_GRID"))     // debugging aid: fewer CTAs (1 = strictly sequential in decoding order)
        grid = atoi(g) > 0 && atoi(g) < grid ? atoi(g) : grid;
    kern<<<grid, kThreads, 0, ctx->stream>>>(d);
    VVC_LAUNCHED(ctx);
    // a fired watchdog (availability counts that name samples of later steps) must not pass silently
    uint32_t flag = 0;
    VVC_TRY(ctx, cudaMemcpyAsync(&flag, d.counter + 1, sizeof(flag), cudaMemcpyDeviceToHost, ctx->stream));
    VVC_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "intra_recon_ordered: a step waited for samples that no earlier step reconstructs (availability counts / step order)");
    return VVC_CUDA_OK;
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
