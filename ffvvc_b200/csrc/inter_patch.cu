// Inter prediction for 10-bit pictures, thread-per-patch kernel (sm_100a): every record that needs no
// cooperative tile -- plain uni / bi prediction, BCW and explicit weights, GPM -- which is most of a picture.
//
// Reference semantics: pred_regular_luma / pred_regular_chroma (libavcodec/vvc/vvc_inter.c:545-639),
// pred_gpm_blk :466-521, the edge emulation :33-58, and the table entries put / put_uni / put_uni_w
// (libavcodec/h26x/h2656_inter_template.c:29-577), avg, w_avg, put_gpm (libavcodec/vvc/vvc_inter_template.c
// :25-98).
//
// B200 design: 8 lanes per record, one lane per patch of 4 columns x 8 rows.  A lane runs the whole separable
// filter in registers: per window row six 32-bit loads (12 samples, L1/L2 resident: neighbouring patches
// share them), the horizontal 8-tap sums as IDP.2A on sample pairs, and a streaming vertical pass that keeps
// only the last 7 row-pairs per column.  No shared memory, no synchronisation, no per-record scalar work
// replicated over a warp -- about 25 thread-instructions per predicted sample.
#include "inter_common.cuh"
#include "tables.cuh"

#ifndef INTER_PLAIN_ASM
#define INTER_PLAIN_ASM 1
#endif
#if INTER_PLAIN_ASM
#define __dp2a_lo nv_dp2a_lo
#define __dp2a_hi nv_dp2a_hi
#define __funnelshift_rc nv_frc
#endif

namespace {

constexpr int kThreads = 128;
// window rows requested ahead of the row being filtered (the loads and the funnel shifts are volatile asm, so the
// compiler keeps their program order: without this every row costs one full memory round trip)
#ifndef PATCH_DEPTH
#define PATCH_DEPTH 2
#endif
// resident CTAs per SM the class kernels are compiled for (register budget = 65536 / (128 * n)); measured sweep in profiles/README.md
#ifndef PATCH_MB_LU
#define PATCH_MB_LU 4
#endif
#ifndef PATCH_MB_LB
#define PATCH_MB_LB 4
#endif
#ifndef PATCH_MB_CU
#define PATCH_MB_CU 7
#endif
#ifndef PATCH_MB_CB
#define PATCH_MB_CB 6
#endif
constexpr int patch_ctas(bool luma, bool bi) { return luma ? (bi ? PATCH_MB_LB : PATCH_MB_LU) : (bi ? PATCH_MB_CB : PATCH_MB_CU); }

// 8 samples x .. x + 7 of a row with the column clamped to the picture (the rare path: one out-of-line copy)
__device__ __noinline__ uint4 load8_clamped(const pel *row, int x, int W)
{
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
        v[k] = __ldg(row + d_clip3(x + k, 0, W - 1));
    return make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
}

struct Blend { int w0, w1, off, sh, ox; };

// One list of one patch, 4 columns x 8 rows: the 14-bit intermediate prediction put() produces.  (x, y) =
// integer position of the patch's first sample in the reference plane; rows >= nrows are not computed.
// KEEP: the rows are kept as int16 pairs in keep[] (first list of a bi record).  Otherwise each finished row is
// blended -- with keep[] when BI -- and stored at d: clip(((a * w0 + b * w1 + off) >> sh) + ox); gw != NULL
// gives per-sample GPM weights (w0 = g, w1 = 8 - g; first sample's address, steps gsx / gsy, columns clamped
// to cmax).
template <int TAPS, bool KEEP, bool BI>
__device__ __forceinline__ void mc_patch(const pel *plane, int pitch, int W, int H, int margin, int x, int y, int nrows,
                                         uint32_t hf0, uint32_t hf1, uint32_t vf0, uint32_t vf1, int shh, int shv,
                                         uint32_t keep[16], const Blend &k, const uint8_t *gw, int gsx, int gsy, int cmax,
                                         pel *d, int dpitch, bool two_words)
{
    constexpr int B = TAPS / 2 - 1, NW = TAPS == 8 ? 6 : 4, NR = 8 + TAPS - 1;
    const int e = (x - B) & 1, bx = x - B - e, sh = e << 4, y0 = y - B;
    // pre-padded reference planes (margin > 0) hold the clamped samples themselves: the plain loads may reach into the margin
    const bool inside = bx >= -margin && bx + 2 * NW <= W + margin && y0 >= -margin && y0 + nrows + TAPS - 1 <= H + margin;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(plane + (long long)y0 * pitch + bx);
    const int wpitch = pitch >> 1;
    uint32_t P[NR - 1][4];
    int prev[4] = { 0, 0, 0, 0 };
    constexpr int D = PATCH_DEPTH;
    uint32_t ahead[D][NW];
    auto fetch = [&](int r, uint32_t (&wd)[NW]) {
        if (inside) {
#pragma unroll
            for (int i = 0; i < NW; i++)
                wd[i] = __ldg(src + r * wpitch + i);
        } else {                                       // emulated_edge_mc: coordinates clamped to the picture
            const pel *row = plane + (long long)d_clip3(y0 + r, 0, H - 1) * pitch;
            const uint4 q0 = load8_clamped(row, bx, W);
            wd[0] = q0.x; wd[1] = q0.y; wd[2] = q0.z; wd[3] = q0.w;
            if (NW > 4) {
                const uint4 q1 = load8_clamped(row, bx + 8, W);
                wd[NW - 2] = q1.x; wd[NW - 1] = q1.y;
            }
        }
    };
#pragma unroll
    for (int r = 0; r < D; r++)
        if (r < nrows + TAPS - 1)
            fetch(r, ahead[r]);
#pragma unroll
    for (int r = 0; r < NR; r++) {
        if (r < nrows + TAPS - 1) {
            uint32_t wd[NW];
#pragma unroll
            for (int i = 0; i < NW; i++)
                wd[i] = ahead[r % D][i];
            if (r + D < NR && r + D < nrows + TAPS - 1)
                fetch(r + D, ahead[r % D]);
            uint32_t A[2 * NW - 2];
#pragma unroll
            for (int i = 0; i < NW - 1; i++) {
                A[2 * i] = __funnelshift_rc(wd[i], wd[i + 1], sh);
                A[2 * i + 1] = __funnelshift_rc(wd[i], wd[i + 1], sh + 16);
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int v;
                if (TAPS == 8)
                    v = __dp2a_lo((int)A[c], (int)hf0, __dp2a_hi((int)A[c + 2], (int)hf0,
                        __dp2a_lo((int)A[c + 4], (int)hf1, __dp2a_hi((int)A[c + 6], (int)hf1, 0))));
                else
                    v = __dp2a_lo((int)A[c], (int)hf0, __dp2a_hi((int)A[c + 2], (int)hf0, 0));
                v >>= shh;
                if (r > 0)
                    P[r - 1][c] = __byte_perm(prev[c], v, 0x5410);      // (row r - 1, row r) of column c
                prev[c] = v;
            }
            if (r >= TAPS - 1) {
                const int yo = r - (TAPS - 1);
                int o[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if (TAPS == 8)
                        o[c] = __dp2a_lo((int)P[yo][c], (int)vf0, __dp2a_hi((int)P[yo + 2][c], (int)vf0,
                               __dp2a_lo((int)P[yo + 4][c], (int)vf1, __dp2a_hi((int)P[yo + 6][c], (int)vf1, 0)))) >> shv;
                    else
                        o[c] = __dp2a_lo((int)P[yo][c], (int)vf0, __dp2a_hi((int)P[yo + 2][c], (int)vf0, 0)) >> shv;
                }
                if (KEEP) {
                    keep[2 * yo] = pack16(o[0], o[1]);
                    keep[2 * yo + 1] = pack16(o[2], o[3]);
                } else {
                    const int a[4] = { lo16(keep[2 * yo]), hi16(keep[2 * yo]), lo16(keep[2 * yo + 1]), hi16(keep[2 * yo + 1]) };
                    int out[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int cur = (short)o[c];            // put() stores int16_t
                        int v;
                        if (BI) {
                            int w0 = k.w0, w1 = k.w1;
                            if (gw) {
                                w0 = gw[yo * gsy + min(c, cmax) * gsx];
                                w1 = 8 - w0;
                            }
                            v = a[c] * w0 + cur * w1;
                        } else {
                            v = cur * k.w0;
                        }
                        out[c] = d_clip_pel(((v + k.off) >> k.sh) + k.ox, 10);
                    }
                    uint32_t *q = reinterpret_cast<uint32_t *>(d + (long long)yo * dpitch);
                    q[0] = pack16(out[0], out[1]);
                    if (two_words)
                        q[1] = pack16(out[2], out[3]);
                }
            }
        }
    }
}

// Final rounding of every record kind as ONE formula, clip(((a * w0 + b * w1 + off) >> sh) + ox):
//   put_uni            w0 = 1,  off = 8, sh = 4                                  (h2656_inter_template.c:44-58)
//   put_uni_w          w0 = wx, off = 1 << (sh - 1), sh = denom + 4, ox          (:60-78)
//   avg                w0 = w1 = 1, off = 16, sh = 5                             (vvc_inter_template.c:25-40)
//   w_avg (BCW / WP)   w0, w1, off = ((o0 + o1) * 4 + 1) << (sh - 1), sh = denom + 5   (:42-57)
//   put_gpm            w0 = g, w1 = 8 - g per sample, off = 64, sh = 7           (:78-98)
__device__ __forceinline__ Blend blend_of(const Rec &pb, const VVCCudaWP *wp, bool bi, int lx, int c)
{
    Blend k;
    if (pb.flags & VVC_CUDA_PB_GPM) {
        k.w0 = 0; k.w1 = 0; k.off = 64; k.sh = 7; k.ox = 0;
    } else if (!bi) {
        const UniW u = uni_weights(pb, wp, lx, c);
        k.w0 = u.on ? u.wx : 1; k.w1 = 0; k.sh = u.on ? u.shift : 4; k.off = 1 << (k.sh - 1); k.ox = u.ox;
    } else {
        const Weights v = bi_weights(pb, wp, c);
        k.w0 = v.on ? v.w0 : 1; k.w1 = v.on ? v.w1 : 1; k.sh = (v.on ? v.denom : 0) + 5;
        k.off = v.on ? (((v.o0 + v.o1) << 2) + 1) << (k.sh - 1) : 16; k.ox = 0;
    }
    return k;
}

// ---- work lists -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inter_classify_kernel(const InterK p, const InterLists ls)
{
    // one reservation per list per CTA: block-wide exclusive scan of the five per-record counts
    __shared__ uint32_t warp_tot[8][10], cta_base[10];
    const int ri = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int cls_l = -1, cls_c = -1, n_l = 0, n_c = 0, coop = -1, border = 0;
    if (ri < p.n) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p.pbs + ri);
        const uint32_t r0 = __ldg(q), r1 = __ldg(q + 1), r2 = __ldg(q + 2);
        const int w = r1 & 0xff, h = (r1 >> 8) & 0xff, planes = (r1 >> 16) & 0xff, pred = r1 >> 24, flags = r2 >> 24;
        if (flags & VVC_PB_COOPERATIVE) {
            coop = (flags & (VVC_CUDA_PB_PROF0 | VVC_CUDA_PB_PROF1)) ? 1 : 0;      // 0: DMVR / BDOF, 1: PROF
        } else {
            const int bi = (flags & VVC_CUDA_PB_GPM) || pred == 3;
            if (planes & VVC_CUDA_PB_LUMA) { cls_l = bi; n_l = (w >> 2) * ((h + 7) >> 3); }
            if ((planes & VVC_CUDA_PB_CHROMA) && p.planes == 3) { cls_c = bi; n_c = w > 8 ? 4 : 2; }   // 2 planes x patch columns
            // Records whose windows may leave the picture (a conservative test; mc_patch decides per patch) go to lists
            // of their own: their clamped loads cost ten times a plain patch, and one such lane stalls its whole warp.
            const int x0 = r0 & 0xffff, y0 = r0 >> 16;
            for (int l = 0; l < 2; l++) {
                if (!bi && l != pred - 1)
                    continue;
                const int mx = (int)__ldg(q + 3 + 2 * l), my = (int)__ldg(q + 4 + 2 * l);
                const int xx = x0 + (mx >> 4), yy = y0 + (my >> 4);
                const int m = p.margin, mc = p.margin >> 1;
                border |= xx - 4 < -m || xx + w + 5 > p.w + m || yy - 3 < -m || yy + h + 4 > p.h + m;
                const int xc = (x0 >> 1) + (mx >> 5), yc = (y0 >> 1) + (my >> 5);
                border |= xc - 2 < -mc || xc + (w >> 1) + 6 > (p.w >> 1) + mc || yc - 1 < -mc || yc + (h >> 1) + 2 > (p.h >> 1) + mc;
            }
        }
    }
    const int bl = border ? 6 : 0, bc = border ? 8 : 2;       // slots of this record's luma / chroma class in mine[]
    uint32_t mine[10] = { 0, 0, 0, 0, coop == 0 ? 1u : 0u, coop == 1 ? 1u : 0u, 0, 0, 0, 0 };
#pragma unroll
    for (int c = 0; c < 2; c++) {
        mine[c]     = (!border && cls_l == c) ? (uint32_t)n_l : 0u;
        mine[2 + c] = (!border && cls_c == c) ? (uint32_t)n_c : 0u;
        mine[6 + c] = (border && cls_l == c) ? (uint32_t)n_l : 0u;
        mine[8 + c] = (border && cls_c == c) ? (uint32_t)n_c : 0u;
    }
    uint32_t excl[10];
#pragma unroll
    for (int c = 0; c < 10; c++) {
        uint32_t v = mine[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        excl[c] = v - mine[c];
        if (lane == 31) warp_tot[wid][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        uint32_t tot = 0;
        for (int k = 0; k < 8; k++) { const uint32_t t = warp_tot[k][threadIdx.x]; warp_tot[k][threadIdx.x] = tot; tot += t; }
        // count[] slots: 0..5 as documented, border lists at 10..13
        cta_base[threadIdx.x] = tot ? atomicAdd(ls.count + (threadIdx.x < 6 ? threadIdx.x : threadIdx.x + 4), tot) : 0u;
    }
    __syncthreads();
    auto at = [&](int c) { uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < 10; k++) if (k == c) r = cta_base[k] + warp_tot[wid][k] + excl[k];
        return (int)r; };
    if (coop >= 0) {
        const int pos = at(4 + coop);
        ls.coop[coop ? p.n - 1 - pos : pos] = ri;
    }
    if (cls_l >= 0) {
        const int base = at(bl + cls_l);
        uint32_t *list = border ? ls.luma_b : ls.luma;
        for (int k = 0; k < n_l; k++)
            list[cls_l ? ls.cap_luma - 1 - (base + k) : base + k] = ((uint32_t)ri << 3) | k;
    }
    if (cls_c >= 0) {
        const int base = at(bc + cls_c);
        uint32_t *list = border ? ls.chroma_b : ls.chroma;
        for (int k = 0; k < n_c; k++)
            list[cls_c ? ls.cap_chroma - 1 - (base + k) : base + k] = ((uint32_t)ri << 3) | k;
    }
}

#define MV0(l, c) ((l) ? pb.mv[1][c] : pb.mv[0][c])
#define REF(l)    ((l) ? pb.ref[1] : pb.ref[0])

// ---- luma: patch pi of (w / 4) x ceil(h / 8) ---------------------------------------------------------------
template <bool BI>
__device__ __forceinline__ void luma_task(const InterK &p, const Rec &pb, int pi)
{
    const int w = pb.w, h = pb.h, lw = 31 - __clz(w);
    const bool gpm = BI && (pb.flags & VVC_CUDA_PB_GPM);
    const int lx = pb.pred - 1;
    const uint2 *lumaf = reinterpret_cast<const uint2 *>(&vvct_luma_mc_filters[0][0][0]);
    const uint8_t *wt = &vvct_gpm_weights[0][0] + pb.gw;
    const int ox = (pi & ((w >> 2) - 1)) << 2, oy = (pi >> (lw - 2)) << 3, nrows = min(8, h - oy);
    const Blend k = blend_of(pb, p.wp, BI, lx, 0);
    pel *d = p.dst[0] + pb.pic * p.db[0] + (long long)(pb.y0 + oy) * p.dp[0] + pb.x0 + ox;
    const uint8_t *gw = gpm ? wt + oy * pb.gsy + ox * pb.gsx : nullptr;
    uint32_t keep[16];
#define LUMA_ARGS(l) p.ref[0] + REF(l) * p.rb[0], p.rp[0], p.w, p.h, p.margin, pb.x0 + ox + (MV0(l, 0) >> 4), pb.y0 + oy + (MV0(l, 1) >> 4), nrows, \
        (MV0(l, 0) & 15) ? fh.x : 0x01000000u, (MV0(l, 0) & 15) ? fh.y : 0u, \
        (MV0(l, 1) & 15) ? fv.x : ((MV0(l, 0) & 15) ? 0x01000000u : 0x10000000u), (MV0(l, 1) & 15) ? fv.y : 0u, \
        (MV0(l, 0) & 15) ? 2 : 0, (MV0(l, 1) & 15) ? ((MV0(l, 0) & 15) ? 6 : 2) : 0
    const int filt = gpm ? 0 : pb.filt;
    if (BI) {
        {
            const uint2 fh = lumaf[filt * 16 + (pb.mv[0][0] & 15)], fv = lumaf[filt * 16 + (pb.mv[0][1] & 15)];
            mc_patch<8, true, true>(LUMA_ARGS(0), keep, k, gw, pb.gsx, pb.gsy, 3, d, p.dp[0], true);
        }
        {
            const uint2 fh = lumaf[filt * 16 + (pb.mv[1][0] & 15)], fv = lumaf[filt * 16 + (pb.mv[1][1] & 15)];
            mc_patch<8, false, true>(LUMA_ARGS(1), keep, k, gw, pb.gsx, pb.gsy, 3, d, p.dp[0], true);
        }
    } else {
        const uint2 fh = lumaf[filt * 16 + (MV0(lx, 0) & 15)], fv = lumaf[filt * 16 + (MV0(lx, 1) & 15)];
        mc_patch<8, false, false>(LUMA_ARGS(lx), keep, k, nullptr, 0, 0, 3, d, p.dp[0], true);
    }
#undef LUMA_ARGS
}

// ---- chroma (4:2:0): task pi = (plane, patch column) ----------------------------------------------------------
template <bool BI>
__device__ __forceinline__ void chroma_task(const InterK &p, const Rec &pb, int pi)
{
    const bool gpm = BI && (pb.flags & VVC_CUDA_PB_GPM);
    const int lx = pb.pred - 1;
    const uint32_t *chromaf = reinterpret_cast<const uint32_t *>(&vvct_chroma_mc_filters[0][0][0]);
    const uint8_t *wt = &vvct_gpm_weights[0][0] + pb.gw;
    const int bw = pb.w >> 1, bh = pb.h >> 1, ncx = bw > 4 ? 2 : 1;
    const int pc = pi >> (ncx >> 1), ox = (pi & (ncx - 1)) << 2;
    const int x0 = pb.x0 >> 1, y0 = pb.y0 >> 1;
    const pel *rplane = pc ? p.ref[2] : p.ref[1];
    const long long rb = pc ? p.rb[2] : p.rb[1];
    const int rp = pc ? p.rp[2] : p.rp[1];
    const Blend k = blend_of(pb, p.wp, BI, lx, pc + 1);
    const int dpitch = pc ? p.dp[2] : p.dp[1];
    pel *d = (pc ? p.dst[2] + pb.pic * p.db[2] : p.dst[1] + pb.pic * p.db[1]) + (long long)y0 * dpitch + x0 + ox;
    // GPM weights of chroma: every second luma weight; columns past a 2-wide block reuse its last weight
    const uint8_t *gw = gpm ? wt + ox * 2 * pb.gsx : nullptr;
    const int cmax = bw - 1 - ox > 3 ? 3 : bw - 1 - ox;
    uint32_t keep[16];
#define CHROMA_ARGS(l) rplane + REF(l) * rb, rp, p.w >> 1, p.h >> 1, p.margin >> 1, x0 + ox + (MV0(l, 0) >> 5), y0 + (MV0(l, 1) >> 5), bh, \
        (MV0(l, 0) & 31) ? chromaf[MV0(l, 0) & 31] : 0x00000100u, 0u, \
        (MV0(l, 1) & 31) ? chromaf[MV0(l, 1) & 31] : ((MV0(l, 0) & 31) ? 0x00000100u : 0x00001000u), 0u, \
        (MV0(l, 0) & 31) ? 2 : 0, (MV0(l, 1) & 31) ? ((MV0(l, 0) & 31) ? 6 : 2) : 0
    if (BI) {
        mc_patch<4, true, true>(CHROMA_ARGS(0), keep, k, gw, 2 * pb.gsx, 2 * pb.gsy, cmax, d, dpitch, bw > 2);
        mc_patch<4, false, true>(CHROMA_ARGS(1), keep, k, gw, 2 * pb.gsx, 2 * pb.gsy, cmax, d, dpitch, bw > 2);
    } else {
        mc_patch<4, false, false>(CHROMA_ARGS(lx), keep, k, nullptr, 0, 0, cmax, d, dpitch, bw > 2);
    }
#undef CHROMA_ARGS
}
#undef MV0
#undef REF

// Persistent kernels, one per task class (own register budget each): a grid-stride loop over the class's
// list, so every warp of the launch runs the same specialised code.
template <bool LUMA, bool BI>
__global__ void __launch_bounds__(kThreads, patch_ctas(LUMA, BI)) inter_patch_kernel(const InterK p, const InterLists ls)
{
    // the class's border tasks, then its plain tasks: border tasks are contiguous in the index range, so the warps that
    // run the clamped path are full of such tasks instead of dragging plain patches through it, and they start first
    const int n = (int)ls.count[(LUMA ? 0 : 2) + (BI ? 1 : 0)], nb = (int)ls.count[10 + (LUMA ? 0 : 2) + (BI ? 1 : 0)];
    const uint32_t *list = LUMA ? ls.luma : ls.chroma, *blist = LUMA ? ls.luma_b : ls.chroma_b;
    const int cap = LUMA ? ls.cap_luma : ls.cap_chroma;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n + nb; i += gridDim.x * kThreads) {
        const uint32_t t = i < nb ? __ldg(blist + (BI ? cap - 1 - i : i)) : __ldg(list + (BI ? cap - 1 - (i - nb) : i - nb));
        const Rec pb = load_rec(p.pbs + (t >> 3));
        if (LUMA)
            luma_task<BI>(p, pb, t & 7);
        else
            chroma_task<BI>(p, pb, t & 7);
    }
}

}  // namespace

int vvc_inter_launch_classify(VVCCudaCtx *ctx, const InterK &p, InterLists *ls)
{
    // scratch: [count: 64 bytes][luma: 8 n][chroma: 4 n][coop: n][luma border: 8 n][chroma border: 4 n] words
    const size_t n = (size_t)p.n;
    uint32_t *base = (uint32_t *)vvc_ctx_scratch(ctx, 2, 64 + 25 * n * sizeof(uint32_t) + n * sizeof(VVCCudaDmvrOut));     // + refined vectors of the split DMVR kernels
    if (!base)
        return ctx->err;
    ls->count = base;
    ls->luma = base + 16;            ls->cap_luma = (int)(8 * n);
    ls->chroma = ls->luma + 8 * n;   ls->cap_chroma = (int)(4 * n);
    ls->coop = ls->chroma + 4 * n;
    ls->luma_b = ls->coop + n;
    ls->chroma_b = ls->luma_b + 8 * n;
    VVC_TRY(ctx, cudaMemsetAsync(ls->count, 0, 64, ctx->stream));
    inter_classify_kernel<<<ceil_div(p.n, 256), 256, 0, ctx->stream>>>(p, *ls);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

// spread: luma bi on side stream 0, luma uni on 1, both chroma classes on 2 (the warp kernels take the context stream);
// otherwise everything on the context stream
// spread == 2 (long launches): only the two chroma classes leave the context stream (side stream 2)
int vvc_inter_launch_patch(VVCCudaCtx *ctx, const InterK &p, const InterLists &ls, int spread)
{
    cudaStream_t s0 = spread == 1 ? ctx->side[0] : ctx->stream, s1 = spread == 1 ? ctx->side[1] : ctx->stream, s2 = spread ? ctx->side[2] : ctx->stream;
    inter_patch_kernel<true, true><<<148 * patch_ctas(true, true), kThreads, 0, s0>>>(p, ls);
    VVC_LAUNCHED(ctx);
    inter_patch_kernel<true, false><<<148 * patch_ctas(true, false), kThreads, 0, s1>>>(p, ls);
    VVC_LAUNCHED(ctx);
    inter_patch_kernel<false, true><<<148 * patch_ctas(false, true), kThreads, 0, s2>>>(p, ls);
    VVC_LAUNCHED(ctx);
    inter_patch_kernel<false, false><<<148 * patch_ctas(false, false), kThreads, 0, s2>>>(p, ls);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}
