// Deblocking stage for sm_100a: one launch per direction per picture ring.
//
// Replaces the pixel work of ff_vvc_deblock_vertical / _horizontal (libavcodec/vvc/vvc_filter.c:861-1003):
//   vvc_loop_filter_luma              libavcodec/vvc/vvc_filter_template.c:546-631
//   loop_filter_luma_large            libavcodec/vvc/vvc_filter_template.c:466-544
//   loop_filter_luma_strong / _weak   libavcodec/h26x/h2656_deblock_template.c:25-81
//   vvc_loop_filter_chroma            libavcodec/vvc/vvc_filter_template.c:681-754
//   loop_filter_chroma_strong(_one_side) :633-679, loop_filter_chroma_weak h2656_deblock_template.c:83-99
//
// B200 design: a CTA owns a TWxTH tile of one plane.  It stages the tile plus an 8-sample apron
// across the edge direction into shared memory with 128-bit loads, then one thread per
// (edge, line) evaluates the segment decision and filters its line from the *unfiltered* copy into
// a second shared copy; the tile's own samples are written back with 128-bit stores.  Each sample
// is written exactly once by its owning CTA (edges on both tile borders are evaluated by both
// neighbours), so the pass is deterministic, out of place, and for edge sets that obey the
// reference's filter-length rules identical to the reference's in-place sequential walk
// (same-direction edges never read each other's output).
#include "common.cuh"

namespace {

struct DbkK {
    const pel *src[3];
    pel       *dst[3];
    int        sp[3], dp[3];
    long long  sb[3], db[3];
    int        pw[3], ph[3];              // plane sizes
    int        vs[3], hs[3];
    const VVCCudaDbkEdge *map[3];
    int        mpitch[3];
    long long  msize[3];
    int        bd, ctb_log2, planes;
    int        tiles_x[3], tiles_y[3];                 // tiles per plane
};

constexpr int kThreads = 256;

#define AT(base, k) ((base)[(k) * XS])
#ifndef DBK_LINE_UNROLL
#define DBK_LINE_UNROLL 1
#endif

// One luma segment (4 lines across one edge) by one thread: the 4 + 4 samples next to the edge of all four lines are
// loaded once, the segment's decisions (vvc_loop_filter_luma :546-631 takes them from lines 0 and 3) are made once, and
// the lines are filtered from registers.  in / out: Q0 of line 0 in the unfiltered / output tile; XS / YS: sample strides
// across / along the edge.
template <int XS, int YS>
__device__ __forceinline__ void luma_segment(const pel *in, pel *out, int tc_in, int beta_in,
                                             int lp, int lq, int hor_ctu_edge, int bd)
{
    const int tc = tc_in << (bd - 10), beta = beta_in << (bd - 8);
    if (!tc)
        return;
    // lines 0 and 3 decide for the segment: v0 / v3[k + 4], k = -4 .. 3 (P3 .. P0, Q0 .. Q3)
    int v0[8], v3[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        v0[k] = in[(k - 4) * XS];
        v3[k] = in[3 * YS + (k - 4) * XS];
    }
#define CURV(v, a, b, c) abs(v[(a) + 4] - 2 * v[(b) + 4] + v[(c) + 4])
    const int dp0 = CURV(v0, -3, -2, -1), dq0 = CURV(v0, 2, 1, 0), dp3 = CURV(v3, -3, -2, -1), dq3 = CURV(v3, 2, 1, 0);
#undef CURV
    const int d0 = dp0 + dq0, d3 = dp3 + dq3;
    const int tc25 = (tc * 5 + 1) >> 1;
    const bool big_p = lp > 3 && !hor_ctu_edge, big_q = lq > 3;

    if (big_p || big_q) {
        const pel *l0 = in, *l3 = in + 3 * YS;
        auto curv = [&](const pel *l, int a, int b, int c) { return abs((int)AT(l, a) - 2 * (int)AT(l, b) + (int)AT(l, c)); };
        const int dp0l = big_p ? (dp0 + curv(l0, -6, -5, -4) + 1) >> 1 : dp0;
        const int dq0l = big_q ? (dq0 + curv(l0, 5, 4, 3) + 1) >> 1 : dq0;
        const int dp3l = big_p ? (dp3 + curv(l3, -6, -5, -4) + 1) >> 1 : dp3;
        const int dq3l = big_q ? (dq3 + curv(l3, 5, 4, 3) + 1) >> 1 : dq3;
        const int d0l = dp0l + dq0l, d3l = dp3l + dq3l;
        lp = big_p ? lp : 3;          // sticky for the normal decision below (:591-592)
        lq = big_q ? lq : 3;
        if (d0l + d3l < beta) {
            const int b53 = (beta * 3) >> 5, b4 = beta >> 4;
            int sp0 = abs(AT(l0, -4) - AT(l0, -1)) + (lp == 7 ? abs(AT(l0, -8) - AT(l0, -7) - AT(l0, -6) + AT(l0, -5)) : 0);
            int sq0 = abs(AT(l0, 0) - AT(l0, 3))   + (lq == 7 ? abs(AT(l0, 4) - AT(l0, 5) - AT(l0, 6) + AT(l0, 7)) : 0);
            int sp3 = abs(AT(l3, -4) - AT(l3, -1)) + (lp == 7 ? abs(AT(l3, -8) - AT(l3, -7) - AT(l3, -6) + AT(l3, -5)) : 0);
            int sq3 = abs(AT(l3, 0) - AT(l3, 3))   + (lq == 7 ? abs(AT(l3, 4) - AT(l3, 5) - AT(l3, 6) + AT(l3, 7)) : 0);
            if (big_p) {
                sp0 = (sp0 + abs(AT(l0, -4) - AT(l0, -1 - lp)) + 1) >> 1;
                sp3 = (sp3 + abs(AT(l3, -4) - AT(l3, -1 - lp)) + 1) >> 1;
            }
            if (big_q) {
                sq0 = (sq0 + abs(AT(l0, 3) - AT(l0, lq)) + 1) >> 1;
                sq3 = (sq3 + abs(AT(l3, 3) - AT(l3, lq)) + 1) >> 1;
            }
            if (sp0 + sq0 < b53 && abs(AT(l0, -1) - AT(l0, 0)) < tc25 &&
                sp3 + sq3 < b53 && abs(AT(l3, -1) - AT(l3, 0)) < tc25 &&
                (d0l << 1) < b4 && (d3l << 1) < b4) {
                // ---- long filter (:466-544): bilinear pull towards the middle value m, line by line ----
                const int w0p = lp == 3 ? 53 : lp == 5 ? 58 : 59, wsp = lp == 3 ? 21 : lp == 5 ? 13 : 9, ksp = lp == 3 ? 2 : 1;
                const int w0q = lq == 3 ? 53 : lq == 5 ? 58 : 59, wsq = lq == 3 ? 21 : lq == 5 ? 13 : 9, ksq = lq == 3 ? 2 : 1;
#pragma unroll 1
                for (int line = 0; line < 4; line++) {
                    const pel *li = in + line * YS;
                    pel *lo = out + line * YS;
                    int p[8], q[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) { p[i] = AT(li, -1 - i); q[i] = AT(li, i); }
                    int m;
                    if (lp == 5 && lq == 5)
                        m = (p[4] + p[3] + 2 * (p[2] + p[1] + p[0] + q[0] + q[1] + q[2]) + q[3] + q[4] + 8) >> 4;
                    else if (lp == lq)
                        m = (p[6] + p[5] + p[4] + p[3] + p[2] + p[1] + 2 * (p[0] + q[0]) + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + 8) >> 4;
                    else if (lp + lq == 12)
                        m = (p[5] + p[4] + p[3] + p[2] + 2 * (p[1] + p[0] + q[0] + q[1]) + q[2] + q[3] + q[4] + q[5] + 8) >> 4;
                    else if (lp + lq == 8)
                        m = (p[3] + p[2] + p[1] + p[0] + q[0] + q[1] + q[2] + q[3] + 4) >> 3;
                    else if (lq == 7)
                        m = (2 * (p[2] + p[1] + p[0] + q[0]) + p[0] + p[1] + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + 8) >> 4;
                    else
                        m = (p[6] + p[5] + p[4] + p[3] + p[2] + p[1] + 2 * (q[2] + q[1] + q[0] + p[0]) + q[0] + q[1] + 8) >> 4;
                    // tap weights 3:{53,32,11} 5:{58,45,32,19,6} 7:{59,50,41,32,23,14,5}; tc scale 6,(5),4,(3),2,(1,1)
                    {
                        const int ref = lp == 3 ? (p[3] + p[2] + 1) >> 1 : lp == 5 ? (p[5] + p[4] + 1) >> 1 : (p[7] + p[6] + 1) >> 1;
#pragma unroll
                        for (int i = 0; i < 7; i++)
                            if (i < lp) {
                                const int w = w0p - wsp * i, lim = (tc * max(6 - ksp * i, 1)) >> 1;
                                AT(lo, -1 - i) = (pel)(p[i] + d_clip3(((m * w + ref * (64 - w) + 32) >> 6) - p[i], -lim, lim));
                            }
                    }
                    {
                        const int ref = lq == 3 ? (q[3] + q[2] + 1) >> 1 : lq == 5 ? (q[5] + q[4] + 1) >> 1 : (q[7] + q[6] + 1) >> 1;
#pragma unroll
                        for (int i = 0; i < 7; i++)
                            if (i < lq) {
                                const int w = w0q - wsq * i, lim = (tc * max(6 - ksq * i, 1)) >> 1;
                                AT(lo, i) = (pel)(q[i] + d_clip3(((m * w + ref * (64 - w) + 32) >> 6) - q[i], -lim, lim));
                            }
                    }
                }
                return;
            }
        }
    }
    if (d0 + d3 >= beta)
        return;
    const bool strong = lp > 2 && lq > 2 &&
        abs(v0[0] - v0[3]) + abs(v0[7] - v0[4]) < (beta >> 3) && abs(v0[3] - v0[4]) < tc25 &&
        abs(v3[0] - v3[3]) + abs(v3[7] - v3[4]) < (beta >> 3) && abs(v3[3] - v3[4]) < tc25 &&
        (d0 << 1) < (beta >> 2) && (d3 << 1) < (beta >> 2);
    int np = 1, nq = 1;
    if (!strong && lp > 1 && lq > 1) {
        const int side = (beta + (beta >> 1)) >> 3;
        if (dp0 + dp3 < side) np = 2;
        if (dq0 + dq3 < side) nq = 2;
    }
#pragma unroll 1
    for (int line = 0; line < 4; line++) {
        const pel *li = in + line * YS;
        pel *lo = out + line * YS;
        const int p2 = AT(li, -3), p1 = AT(li, -2), p0 = AT(li, -1), q0 = AT(li, 0), q1 = AT(li, 1), q2 = AT(li, 2);
        if (strong) {
            const int p3 = AT(li, -4), q3 = AT(li, 3);
            AT(lo, -1) = (pel)(p0 + d_clip3(((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3) - p0, -3 * tc, 3 * tc));
            AT(lo, -2) = (pel)(p1 + d_clip3(((p2 + p1 + p0 + q0 + 2) >> 2) - p1, -2 * tc, 2 * tc));
            AT(lo, -3) = (pel)(p2 + d_clip3(((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3) - p2, -tc, tc));
            AT(lo, 0)  = (pel)(q0 + d_clip3(((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3) - q0, -3 * tc, 3 * tc));
            AT(lo, 1)  = (pel)(q1 + d_clip3(((p0 + q0 + q1 + q2 + 2) >> 2) - q1, -2 * tc, 2 * tc));
            AT(lo, 2)  = (pel)(q2 + d_clip3(((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3) - q2, -tc, tc));
        } else {
            int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
            if (abs(delta) < 10 * tc) {
                const int half = tc >> 1;
                delta = d_clip3(delta, -tc, tc);
                AT(lo, -1) = (pel)d_clip_pel(p0 + delta, bd);
                AT(lo, 0)  = (pel)d_clip_pel(q0 - delta, bd);
                if (np > 1)
                    AT(lo, -2) = (pel)d_clip_pel(p1 + d_clip3((((p2 + p0 + 1) >> 1) - p1 + delta) >> 1, -half, half), bd);
                if (nq > 1)
                    AT(lo, 1)  = (pel)d_clip_pel(q1 + d_clip3((((q2 + q0 + 1) >> 1) - q1 - delta) >> 1, -half, half), bd);
            }
        }
    }
}

// One chroma segment of `lines` lines (2 when the edge direction is subsampled, else 4) by one thread.
template <int XS, int YS>
__device__ __forceinline__ void chroma_segment(const pel *in, pel *out, int lines,
                                               int tc_in, int beta_in, int lp, int lq, int bd)
{
    const int tc = tc_in << (bd - 10), beta = beta_in << (bd - 8);
    if (!tc || !lp || !lq)
        return;
    if (lq == 3) {
        const pel *l0 = in, *l1 = in + (lines == 2 ? 1 : 3) * YS;
        auto curv = [&](const pel *l, int a, int b, int c) { return abs((int)AT(l, a) - 2 * (int)AT(l, b) + (int)AT(l, c)); };
        const int tc25 = (tc * 5 + 1) >> 1;
        const bool one = lp == 1;
        const int p0 = AT(l0, -1), p1 = AT(l0, -2), p2 = one ? p1 : AT(l0, -3), p3 = one ? p1 : AT(l0, -4);
        const int n0 = AT(l1, -1), n1 = AT(l1, -2), n2 = one ? n1 : AT(l1, -3), n3 = one ? n1 : AT(l1, -4);
        const int d0 = abs(p2 - 2 * p1 + p0) + curv(l0, 2, 1, 0);
        const int d1 = abs(n2 - 2 * n1 + n0) + curv(l1, 2, 1, 0);
        bool strong = false;
        if (d0 + d1 < beta) {
            const bool ok0 = (d0 << 1) < (beta >> 2) && abs(p3 - p0) + abs(AT(l0, 0) - AT(l0, 3)) < (beta >> 3) && abs(p0 - AT(l0, 0)) < tc25;
            const bool ok1 = (d1 << 1) < (beta >> 2) && abs(n3 - n0) + abs(AT(l1, 0) - AT(l1, 3)) < (beta >> 3) && abs(n0 - AT(l1, 0)) < tc25;
            strong = ok0 && ok1;
        }
        if (!strong)
            lp = lq = 1;
    }
#pragma unroll 1
    for (int line = 0; line < lines; line++) {
        const pel *li = in + line * YS;
        pel *lo = out + line * YS;
        const int p1 = AT(li, -2), p0 = AT(li, -1), q0 = AT(li, 0), q1 = AT(li, 1);
        if (lq == 3) {
            const int q2 = AT(li, 2), q3 = AT(li, 3);
            if (lp == 3) {
                const int p3 = AT(li, -4), p2 = AT(li, -3);
                AT(lo, -1) = (pel)d_clip3((p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3, p0 - tc, p0 + tc);
                AT(lo, -2) = (pel)d_clip3((2 * p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3, p1 - tc, p1 + tc);
                AT(lo, -3) = (pel)d_clip3((3 * p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3, p2 - tc, p2 + tc);
                AT(lo, 0)  = (pel)d_clip3((p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3, q0 - tc, q0 + tc);
            } else {
                AT(lo, -1) = (pel)d_clip3((3 * p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3, p0 - tc, p0 + tc);
                AT(lo, 0)  = (pel)d_clip3((2 * p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3, q0 - tc, q0 + tc);
            }
            AT(lo, 1) = (pel)d_clip3((p1 + p0 + q0 + 2 * q1 + q2 + 2 * q3 + 4) >> 3, q1 - tc, q1 + tc);
            AT(lo, 2) = (pel)d_clip3((p0 + q0 + q1 + 2 * q2 + 3 * q3 + 4) >> 3, q2 - tc, q2 + tc);
        } else {
            const int delta = d_clip3((((q0 - p0) * 4) + p1 - q1 + 4) >> 3, -tc, tc);
            AT(lo, -1) = (pel)d_clip_pel(p0 + delta, bd);
            AT(lo, 0)  = (pel)d_clip_pel(q0 - delta, bd);
        }
    }
}

#undef AT

// VERT: vertical edges (filtering runs along rows).  Tile geometry:
//   VERT : TW = 128, TH = 64, apron of 8 columns left and right, row pitch TW + 16 + 2 (odd word count
//          so the 32 lanes of a warp, which walk consecutive rows of one edge, hit distinct banks)
//   !VERT: TW = 128, TH = 64, apron of 8 rows above and below, lanes walk consecutive columns
// Tile shapes by tools/sweep_dbk_tile.sh (ms per 16 4K pictures, V / H): 64x32 / 128x16 0.364 / 0.323, 128x32 / 128x32
// 0.293 / 0.285, 128x64 / 128x64 0.269 / 0.244, 128x128 0.318 / 0.273, 256x64 V 0.312, 256x32 H 0.263: the apron and the
// per-CTA set-up are amortised over more samples until the shared memory costs resident CTAs.
#ifndef DBK_TW_V
#define DBK_TW_V 128
#endif
#ifndef DBK_TH_H
#define DBK_TH_H 64
#endif
#ifndef DBK_TH_V
#define DBK_TH_V 64
#endif
#ifndef DBK_TW_H
#define DBK_TW_H 128
#endif
template <bool VERT>
struct DbkTile {
    static constexpr int TW = VERT ? DBK_TW_V : DBK_TW_H, TH = VERT ? DBK_TH_V : DBK_TH_H;
    static constexpr int AX = VERT ? 8 : 0, AY = VERT ? 0 : 8;
    static constexpr int PITCH = TW + 2 * AX + (VERT ? 2 : 0);
    static constexpr int ROWS = TH + 2 * AY;
};

#ifndef DBK_STAGE_ASYNC
#define DBK_STAGE_ASYNC 1            // tile staging by cp.async (tools/sweep_alf_stage.sh)
#endif
#ifndef DBK_MIN_CTAS
#define DBK_MIN_CTAS 8               // resident CTAs per SM the kernel is compiled for: full occupancy at 32 registers (tools/sweep_dbk2.sh: 4 / 5 / 6 / 7-8 -> V 0.62 / 0.53 / 0.47 / 0.40 ms per 16 pictures; the pass is latency bound)
#endif
template <bool VERT>
__global__ void
#if DBK_MIN_CTAS > 0
__launch_bounds__(kThreads, DBK_MIN_CTAS)
#else
__launch_bounds__(kThreads)
#endif
deblock_kernel(const DbkK p)
{
    using T = DbkTile<VERT>;
    constexpr int TW = T::TW, TH = T::TH, AX = T::AX, AY = T::AY, PITCH = T::PITCH, ROWS = T::ROWS;
    // ONE tile, filtered in place: no edge reads a sample another edge of the same direction modifies (the filter lengths
    // are derived so - it is what lets the reference filter in place edge after edge, and lets the threads here take the
    // edges concurrently); every line's samples are in registers before its first store
    __shared__ alignas(16) pel s_t[ROWS * PITCH];

    // which plane / tile.  grid: x = tile column (the two chroma planes side by side), y = the tile rows of luma, then
    // those of chroma, z = picture
    int c = 0, tcol = blockIdx.x, trow = blockIdx.y;
    if (trow >= p.tiles_y[0]) {
        trow -= p.tiles_y[0];
        c = 1;
        if (tcol >= p.tiles_x[1]) { tcol -= p.tiles_x[1]; c = 2; }
    }
    if (tcol >= p.tiles_x[c])
        return;
    const int k = blockIdx.z;
    const int tx0 = tcol * TW, ty0 = trow * TH;
    const int pw = p.pw[c], ph = p.ph[c];
    const pel *src = p.src[c] + k * p.sb[c];
    pel *dst = p.dst[c] + k * p.db[c];
    const int tid = threadIdx.x;

    // ---- stage tile + apron (coordinates clamped to the picture) ----
    constexpr int CHUNKS = (TW + 2 * AX) / 8;
    for (int idx = tid; idx < ROWS * CHUNKS; idx += kThreads) {
        const int i = idx / CHUNKS, q = idx - i * CHUNKS;
        const int y = min(max(ty0 - AY + i, 0), ph - 1);
        const int xs0 = tx0 - AX + 8 * q;
        const pel *row = src + (long long)y * p.sp[c];
        unsigned v[4];
        if (xs0 >= 0 && xs0 + 7 < pw) {
            // cp.async where the tile's pitch keeps the chunks 16-byte aligned (the horizontal-edge tile): every chunk of the thread is
            // in flight before the one wait in front of the barrier, H pass 0.245 -> 0.208 ms per 16 4K pictures.  The odd pitch of
            // the vertical-edge tile would need 4-byte requests, which cost more than they save (0.270 -> 0.277): it keeps load + store.
            if (DBK_STAGE_ASYNC && (PITCH * 2) % 16 == 0) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(&s_t[i * PITCH + 8 * q])), "l"(row + xs0) : "memory");
                continue;
            }
            const uint4 u = __ldg(reinterpret_cast<const uint4 *>(row + xs0));
            v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const unsigned lo = __ldg(row + min(max(xs0 + 2 * e, 0), pw - 1));
                const unsigned hi = __ldg(row + min(max(xs0 + 2 * e + 1, 0), pw - 1));
                v[e] = lo | (hi << 16);
            }
        }
        unsigned *o = reinterpret_cast<unsigned *>(&s_t[i * PITCH + 8 * q]);
#pragma unroll
        for (int e = 0; e < 4; e++) o[e] = v[e];
    }
    if (DBK_STAGE_ASYNC && (PITCH * 2) % 16 == 0)
        asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    // ---- filter: one thread per edge segment (4 luma lines; 2 or 4 chroma lines) ----
    const bool chroma = c != 0;
    const int lgrid = chroma ? 3 : 2, grid = 1 << lgrid;
    const int shift = chroma ? (VERT ? p.vs[c] : p.hs[c]) : 0;
    const int lseg = chroma ? 2 - shift : 2, seg = 1 << lseg;   // lines per segment
    constexpr int ALONG = VERT ? TH : TW;                    // tile extent along the edges
    constexpr int ACROSS = VERT ? TW : TH;
    const int n_edges = (ACROSS >> lgrid) + 1;               // both tile borders included
    const int n_segs = ALONG >> lseg;
    const VVCCudaDbkEdge *map = p.map[c] + k * p.msize[c];
    const int ctb_mask = (1 << p.ctb_log2) - 1;
    constexpr int xs = VERT ? 1 : PITCH, ys = VERT ? PITCH : 1;

    // VERT: consecutive lanes take consecutive edges of one row of segments (their samples are 4 or 8 columns apart: the
    // 32 lanes of a load touch 32 different banks); !VERT: consecutive lanes take consecutive segments of one edge
    for (int it = tid; it < n_edges * n_segs; it += kThreads) {
        // it / n_edges as a multiply by ceil(65536 / n_edges): exact for it < 65536 / n_edges
        constexpr unsigned NE_L = ACROSS / 4 + 1, NE_C = ACROSS / 8 + 1;
        const int qv = (int)(((unsigned)it * (chroma ? (65536u + NE_C - 1) / NE_C : (65536u + NE_L - 1) / NE_L)) >> 16);
        const int e = VERT ? it - qv * n_edges : it >> ((TW == 256 ? 8 : 7) - lseg), sg = VERT ? qv : it - e * n_segs;
        static_assert(VERT || TW == 128 || TW == 256, "n_segs = TW >> lseg as a shift");
        const int pos = (VERT ? tx0 : ty0) + e * grid;       // edge coordinate in the plane
        const int a0 = sg * seg, along = (VERT ? ty0 : tx0) + a0;
        if (pos == 0 || pos >= (VERT ? pw : ph) || along >= (VERT ? ph : pw))
            continue;
        const int sidx = along >> lseg;
        const VVCCudaDbkEdge ed = VERT ? map[(long long)sidx * p.mpitch[c] + (pos >> lgrid)]
                                       : map[(long long)(pos >> lgrid) * p.mpitch[c] + sidx];
        if (!ed.tc)
            continue;
        // Q0 of line 0 of this segment inside the tile
        const int off = VERT ? (a0 + AY) * PITCH + (e * grid + AX) : (e * grid + AY) * PITCH + (a0 + AX);
        if (!chroma) {
            const int ctu_edge = !VERT && !(pos & ctb_mask);
            luma_segment<xs, ys>(s_t + off, s_t + off, ed.tc, ed.beta, ed.max_len & 15, ed.max_len >> 4, ctu_edge, p.bd);
        } else {
            chroma_segment<xs, ys>(s_t + off, s_t + off, seg, ed.tc, ed.beta, ed.max_len & 15, ed.max_len >> 4, p.bd);
        }
    }
    __syncthreads();

    // ---- write the tile's own samples ----
    constexpr int OCH = TW / 8;
    for (int idx = tid; idx < TH * OCH; idx += kThreads) {
        const int i = idx / OCH, q = idx - i * OCH;
        const int y = ty0 + i, x = tx0 + 8 * q;
        if (y >= ph || x >= pw)
            continue;
        const pel *sp = &s_t[(i + AY) * PITCH + AX + 8 * q];
        pel *drow = dst + (long long)y * p.dp[c] + x;
        if (x + 7 < pw) {
            const unsigned *s32 = reinterpret_cast<const unsigned *>(sp);
            *reinterpret_cast<uint4 *>(drow) = make_uint4(s32[0], s32[1], s32[2], s32[3]);
        } else {
            for (int e2 = 0; x + e2 < pw; e2++)
                drow[e2] = sp[e2];
        }
    }
}

int check_pair(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src, const char *who)
{
    if (!dst || !src)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "%s: null frame", who);
    if (src->bit_depth != 10 && src->bit_depth != 12)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "%s: bit depth %d not accelerated (10/12 only)", who, src->bit_depth);
    if (src->ctb_log2 < 5 || src->ctb_log2 > 7 || src->batch < 1 || src->width < 8 || src->height < 8 ||
        (src->width & 7) || (src->height & 7))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "%s: unsupported geometry %dx%d", who, src->width, src->height);
    if (src->chroma_format_idc && (src->hshift != 1 || src->vshift != 1))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "%s: only 4:0:0 and 4:2:0 are accelerated", who);
    if (dst->width != src->width || dst->height != src->height || dst->batch != src->batch ||
        dst->bit_depth != src->bit_depth || dst->chroma_format_idc != src->chroma_format_idc)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "%s: dst/src geometry differs", who);
    if (!frame_vec_ok(dst) || !frame_vec_ok(src))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "%s: planes and strides must be 16-byte aligned", who);
    if (dst->data[0] == src->data[0])
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "%s: dst must not alias src", who);
    return 0;
}

template <bool VERT>
int launch(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src, const VVCCudaDeblockMaps *maps, int dir)
{
    using T = DbkTile<VERT>;
    DbkK p;
    p.planes = src->chroma_format_idc ? 3 : 1;
    for (int c = 0; c < 3; c++) {
        p.src[c] = (const pel *)src->data[c];  p.dst[c] = (pel *)dst->data[c];
        p.sp[c] = (int)(src->stride[c] / 2);   p.dp[c] = (int)(dst->stride[c] / 2);
        p.sb[c] = src->batch_stride[c] / 2;    p.db[c] = dst->batch_stride[c] / 2;
        p.hs[c] = c ? src->hshift : 0;         p.vs[c] = c ? src->vshift : 0;
        p.pw[c] = src->width >> p.hs[c];       p.ph[c] = src->height >> p.vs[c];
        p.map[c] = maps->edge[dir][c];         p.mpitch[c] = maps->pitch[dir][c];  p.msize[c] = maps->size[dir][c];
        p.tiles_x[c] = ceil_div(p.pw[c], T::TW);  p.tiles_y[c] = ceil_div(p.ph[c], T::TH);
    }
    p.bd = src->bit_depth; p.ctb_log2 = src->ctb_log2;
    const dim3 grid(p.planes == 3 ? max(p.tiles_x[0], 2 * p.tiles_x[1]) : p.tiles_x[0],
                    p.tiles_y[0] + (p.planes == 3 ? p.tiles_y[1] : 0), src->batch);
    if (grid.y > 65535 || grid.z > 65535)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "deblock: picture ring exceeds the grid");
    deblock_kernel<VERT><<<grid, kThreads, 0, ctx->stream>>>(p);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

}  // namespace

extern "C" int vvc_cuda_deblock_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                                      const VVCCudaDeblockMaps *maps, int dir)
{
    if (ctx->err)
        return ctx->err;
    if (!maps)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "deblock: null maps");
    if (check_pair(ctx, dst, src, "deblock"))
        return ctx->err;
    return dir ? launch<true>(ctx, dst, src, maps, 1) : launch<false>(ctx, dst, src, maps, 0);
}

extern "C" int vvc_cuda_deblock_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                                           const VVCCudaDeblockMaps *maps)
{
    if (ctx->err)
        return ctx->err;
    if (!dst || !src || !maps)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "deblock_host: null argument");
    const int planes = src->chroma_format_idc ? 3 : 1;
    const size_t fsz = align_up(vvc_stage_frame_size(src), 256);
    size_t msz = 0;
    for (int d = 0; d < 2; d++)
        for (int c = 0; c < planes; c++)
            msz += align_up((size_t)maps->size[d][c] * src->batch * sizeof(VVCCudaDbkEdge), 256);
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, 2 * fsz + msz);
    if (!base)
        return ctx->err;
    VVCCudaFrame a, b;
    vvc_stage_frame_layout(src, base, &a);
    vvc_stage_frame_layout(src, base + fsz, &b);
    VVCCudaDeblockMaps dm = *maps;
    uint8_t *at = base + 2 * fsz;
    for (int d = 0; d < 2; d++)
        for (int c = 0; c < planes; c++) {
            const size_t bytes = (size_t)maps->size[d][c] * src->batch * sizeof(VVCCudaDbkEdge);
            VVC_TRY(ctx, cudaMemcpyAsync(at, maps->edge[d][c], bytes, cudaMemcpyHostToDevice, ctx->stream));
            dm.edge[d][c] = (const VVCCudaDbkEdge *)at;
            at += align_up(bytes, 256);
        }
    if (vvc_stage_frame_h2d(ctx, &a, src))
        return ctx->err;
    if (vvc_cuda_deblock_frame(ctx, &b, &a, &dm, 1) || vvc_cuda_deblock_frame(ctx, &a, &b, &dm, 0))
        return ctx->err;
    if (vvc_stage_frame_d2h(ctx, dst, &a))
        return ctx->err;
    return vvc_cuda_sync(ctx);
}
