// Inter prediction for 10-bit pictures, warp-per-record kernel (sm_100a).
//
// Same contract as inter.cu (the generic kernel): replaces the pixel work of ff_vvc_predict_inter
// (libavcodec/vvc/vvc_inter.c:899-913) -- pred_regular_blk :782-811, dmvr_mv_refine :685-748,
// parametric_mv_refine :642-681, pred_affine_blk :828-873 with luma_prof_uni/bi :368-446, pred_gpm_blk
// :466-521, the edge emulation :33-110 -- and the table entries they call: put / put_uni / put_uni_w
// (libavcodec/h26x/h2656_inter_template.c:29-577), avg, w_avg, put_gpm, bdof_fetch_samples,
// fetch_samples, prof_grad_filter, apply_prof*, apply_bdof, dmvr* (libavcodec/vvc/vvc_inter_template.c
// :25-436), sad and pad_int16 (libavcodec/vvc/vvcdsp.c:29-65).
//
// B200 design
//  * One warp owns one record (<= 16x16 luma + its chroma); warps of a CTA are independent, so the very
//    different record kinds (plain uni 4x4 ... DMVR+BDOF 16x16) never wait on each other and the only
//    synchronisation is __syncwarp().
//  * Reference windows are staged with aligned 32-bit loads (two samples) when they lie inside the
//    picture, otherwise sample by sample with the clamp ff_emulated_edge_mc materialises.  For DMVR
//    records the window of the UNREFINED block is staged once with a 2-sample replicated apron: that is
//    exactly emulated_edge_dmvr's clamp window (vvc_inter.c:60-89), and it serves the bilinear search, the
//    final motion compensation at the refined vector and the BDOF ring.
//  * The separable 8/4-tap filters run on sample PAIRS: IDP.2A (two 16-bit samples x two 8-bit taps, exact
//    32-bit accumulate) does an 8-tap sum in 4 instructions.  A lane filters 8 outputs of one row from 16
//    samples it holds in registers and stores them transposed, so the vertical pass reads its column as
//    two 128-bit shared loads and runs the same 8-outputs-per-lane code.
//  * The bilinear DMVR prediction and the 25 SADs use 16x2 packed arithmetic (VIMNMX.U16x2, IDP.2A).
#include <type_traits>
#include <cuda.h>            // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "inter_common.cuh"
#include "tables.cuh"

namespace {

// ---- TMA staging of DMVR windows (KIND 0, TMA == 1) -----------------------------------------------------------
// One tensor map describes the luma plane of the whole reference ring as a 3-D tensor (x, y, picture).  A warp's
// elected lane asks for the (w + 7) x (h + 7) window of each list as one box (24 x 23 samples, the 16x16 case; smaller
// records use its upper-left part) and the copy engine writes it densely into the warp's staging area while the warp is
// still busy with the previous record; completion is a transaction count on the warp's own mbarrier.  Measured on B200
// (tools/tma_probe2.cu): with 16-bit elements the innermost box coordinate must be a multiple of 8 (a 16-byte aligned
// start), any other origin raises "illegal instruction" - so the box starts at the window's column rounded down to 8 and
// is 32 samples wide (window of at most 23 + offset of at most 7), and the move into the padded window layout shifts
// by the offset.  Out-of-picture reads would be zero-filled where the codec wants clamp-to-edge: records whose windows
// leave the picture keep the clamped staging loop.
constexpr int kBoxW = 32, kBoxH = 23, kBoxBytes = kBoxW * kBoxH * 2, kBoxSlot = 1536;   // slot: 128-byte multiple

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, int parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_box_3d(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

#ifndef INTER_STAGE_UNROLL
#define INTER_STAGE_UNROLL 4             // window rows a lane requests before it stores the first one
#endif
#ifndef INTER_STAGE_ASYNC
#define INTER_STAGE_ASYNC 1              // window rows by cp.async instead of load + store (tools/sweep_stage_async.sh)
#endif
#ifndef INTER_WARP_PREFETCH
#define INTER_WARP_PREFETCH 0            // measured slower (tools/sweep_prefetch2.sh: inter 2.285 -> 2.381 ms per 16 pictures)
#endif
constexpr int kWarps = 4, kThreads = kWarps * 32;
constexpr int PWL = 15, WUL = 27 * PWL + 1;   // luma window: 27 rows x 15 words (30 samples); unit stride 406 words
constexpr int PWC = 9, WUC = 15 * PWC;        // chroma window: 15 rows x 9 words (18 samples)
constexpr int HPL = 24, HUL = 16 * HPL;       // transposed first-pass output: [col][row], luma
constexpr int HPC = 16, HUC = 8 * HPC;        //                                            chroma
constexpr int TP = 36, TU = 18 * TP;          // int16 tile: sample (x, y) at (y + 1) * TP + 8 + x, ring around it.  18-word rows: the
                                              // lane-per-row passes (BDOF gradients) meet 16 different banks (14-way conflicts at 16 words)
constexpr int DP = 26, DU = 20 * DP;          // DMVR bilinear tiles, 13-word rows (odd: lane-per-row stores, row-per-lane SAD loads)
constexpr int GP = 20, GU = 16 * GP;          // BDOF gradients: [y * GP + x], 10-word rows

struct UnitMC {
    int      woff, r0, c0;      // window of the unit: word offset, first row, first sample column
    int      shh, shv;          // shifts after the two passes
    uint32_t hf0, hf1, vf0, vf1;
};

struct __align__(16) WarpSmem {
    union {
        struct {
            uint32_t win[2 * WUL];            // 3248 bytes
            short    hbt[2 * HUL];            // 1536 bytes
        } a;
        struct {                              // BDOF, after the windows are dead
            short grad[4][GU];                // [list * 2 + (0 horizontal, 1 vertical)]
            int   csum[5][16][4];             // per row, per block column: sums over the 6 window columns
        } b;
    };
    union {
        short tile[2][TU];
        short dm[2][DU];
        short ctile[4][64];
    };
    UnitMC um[4];
    int    sad[25];
    int    vxy[16][2];
};

// what the DMVR search alone needs (PHASE 1): both windows, the bilinear tiles, the 25 SADs - 5.4 KB instead of 7.7 KB per
// warp, so that more of its CTAs fit an SM
struct __align__(16) SearchSmem {
    struct { uint32_t win[2 * WUL]; } a;
    short dm[2][DU];
    int   sad[25];
};

// (t * inv_rows(d)) >> 16 == t / d for t < 128; d is a window row count: 23, 15, 11 (luma), 11, 7, 5 (chroma)
__device__ __forceinline__ uint32_t inv_rows(int d) { return d == 23 ? 2850u : d == 15 ? 4370u : d == 11 ? 5958u : d == 7 ? 9363u : 13108u; }

// ---- window staging ----------------------------------------------------------------------------------
// All units of a phase (the two lists of luma, or the (plane, list) pairs of chroma) are staged by ONE call: the
// warp is split into equal lane groups, one per unit, and every lane passes its own unit's plane and window
// origin.  Window = samples [wx0, wx0 + cols) x [wy0, wy0 + rows), picture-clamped, stored from the even
// column wx0 & ~1 at word (row + 2 pad) * pw + pad of the unit's window.  Inside a group a lane is (word of
// the row, row sub-group): no index division.  pad: replicated apron of a DMVR window (2 samples each side,
// 2 rows above and below), which equals clamping the coordinates to the unrefined block's window
// (emulated_edge_dmvr, vvc_inter.c:60-89).  One out-of-line copy serves every caller (the kernel's code size
// matters: warps sit in different phases and share the instruction cache).
__device__ __noinline__ void stage_apron(uint32_t *win, int pw, int e, int cols, int rows, int sub, int group);

__device__ __noinline__ void stage_units(uint32_t *win, int pw, int pad, const pel *plane, int pitch, int W, int H, int margin,   //@region stage_core
                                         int wx0, int wy0, int cols, int rows, int sub, int group)
{
    const int e = wx0 & 1, bx = wx0 - e, nw = (cols + e + 1) >> 1;
    const int lk = nw > 8 ? 4 : 3;
    const int k = sub & ((1 << lk) - 1), rsub = sub >> lk, rstep = group >> lk;
    if (k < nw) {
        const bool inside = bx >= -margin && bx + 2 * nw <= W + margin && wy0 >= -margin && wy0 + rows <= H + margin;   // margin: pre-padded planes
        uint32_t *dst = win + (2 * pad + rsub) * pw + pad + k;
        const int dstep = rstep * pw;
        if (inside) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(plane + (long long)(wy0 + rsub) * pitch + bx) + k;
            const int sstep = rstep * (pitch >> 1);
#if INTER_STAGE_ASYNC
            // cp.async: no destination register and no wait between the rows - every row of the lane is in flight before the
            // single wait below (the register form paid a round trip per INTER_STAGE_UNROLL rows: a quarter of the kernel's stall samples)
            uint32_t sd = smem_u32(dst);
#pragma unroll 4
            for (int r = rsub; r < rows; r += rstep, src += sstep, sd += dstep * 4)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sd), "l"(src) : "memory");
            asm volatile("cp.async.wait_all;" ::: "memory");
#else
            constexpr int kStageUnroll = INTER_STAGE_UNROLL;
#pragma unroll kStageUnroll
            for (int r = rsub; r < rows; r += rstep, src += sstep, dst += dstep)
                *dst = __ldg(src);
#endif
        } else {
            const int xa = d_clip3(bx + 2 * k, 0, W - 1), xb = d_clip3(bx + 2 * k + 1, 0, W - 1);
#pragma unroll 2
            for (int r = rsub; r < rows; r += rstep, dst += dstep) {
                const pel *row = plane + (long long)d_clip3(wy0 + r, 0, H - 1) * pitch;
                *dst = (uint32_t)__ldg(row + xa) | ((uint32_t)__ldg(row + xb) << 16);
            }
        }
    }
    if (!pad)
        return;
    stage_apron(win, pw, e, cols, rows, sub, group);
}

// replicated apron of a DMVR window whose core is in place (2 samples each side, 2 rows above and below)
__device__ __noinline__ void stage_apron(uint32_t *win, int pw, int e, int cols, int rows, int sub, int group)
{
    __syncwarp();
    for (int r = sub; r < rows; r += group) {              // lane = core row
        uint16_t *row = reinterpret_cast<uint16_t *>(win) + (r + 2) * (2 * pw);
        const int first = 2 + e, last = 2 + e + cols - 1;
        const uint16_t a = row[first], b = row[last];
        row[0] = a; row[1] = a;
        if (e)
            row[2] = a;
        row[last + 1] = b; row[last + 2] = b;
    }
    __syncwarp();
    for (int c = sub; c < pw; c += group) {
        const uint32_t top = win[2 * pw + c], bot = win[(rows + 1) * pw + c];
        win[c] = top; win[pw + c] = top;
        win[(rows + 2) * pw + c] = bot; win[(rows + 3) * pw + c] = bot;
    }
}

// ---- separable interpolation -------------------------------------------------------------------------
// First pass (put_*_h and the tmp_array loop of put_*_hv, h2656_inter_template.c:97-150, 342-395): a task is
// (unit, 8-column half, window row).  Output stored transposed: hbt[unit][col][row].
template <int TAPS>
__device__ __forceinline__ void pass_h(WarpSmem &s, int lane, int n_units, int rows, int lhh)   //@region pass_h
{
    constexpr int NW = TAPS == 8 ? 8 : 6, PW = TAPS == 8 ? PWL : PWC, HP = TAPS == 8 ? HPL : HPC, HU = TAPS == 8 ? HUL : HUC;
    const int ntask = (n_units << lhh) * rows;
    const uint32_t inv = inv_rows(rows);
    for (int t = lane; t < ntask; t += 32) {
        const int q = (t * inv) >> 16, r = t - q * rows;
        const int hh = q & ((1 << lhh) - 1), u = q >> lhh;
        const UnitMC m = s.um[u];
        const int c = m.c0 + 8 * hh, sh = (c & 1) << 4;
        const uint32_t *wp = s.a.win + m.woff + (m.r0 + r) * PW + (c >> 1);
        uint32_t wd[NW], A[2 * NW - 2];
#pragma unroll
        for (int k = 0; k < NW; k++)
            wd[k] = wp[k];
#pragma unroll
        for (int i = 0; i < NW - 1; i++) {
            A[2 * i] = frc(wd[i], wd[i + 1], sh);
            A[2 * i + 1] = frc(wd[i], wd[i + 1], sh + 16);
        }
        short *hb = s.a.hbt + u * HU + (8 * hh) * HP + r;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int v;
            if (TAPS == 8)
                v = __dp2a_lo((int)A[j], (int)m.hf0, __dp2a_hi((int)A[j + 2], (int)m.hf0,
                    __dp2a_lo((int)A[j + 4], (int)m.hf1, __dp2a_hi((int)A[j + 6], (int)m.hf1, 0))));
            else
                v = __dp2a_lo((int)A[j], (int)m.hf0, __dp2a_hi((int)A[j + 2], (int)m.hf0, 0));
            hb[j * HP] = (short)(v >> m.shh);
        }
    }
}

// Second pass: a task is (unit, 8-row half, column); writes the int16 tile the reference's put() produces.
template <int TAPS>
__device__ __forceinline__ void pass_v(WarpSmem &s, int lane, int n_units, int lbw, int bh)   //@region pass_v
{
    constexpr int HP = TAPS == 8 ? HPL : HPC, HU = TAPS == 8 ? HUL : HUC;
    const int lvh = bh > 8 ? 1 : 0;
    const int ntask = (n_units << lvh) << lbw;
    for (int t = lane; t < ntask; t += 32) {
        const int x = t & ((1 << lbw) - 1), q = t >> lbw;
        const int hv = q & ((1 << lvh) - 1), u = q >> lvh;
        const UnitMC m = s.um[u];
        const uint4 *cp = reinterpret_cast<const uint4 *>(s.a.hbt + u * HU + x * HP + 8 * hv);
        const uint4 c0 = cp[0], c1 = cp[1];
        const uint32_t wd[8] = { c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w };
        uint32_t A[14];
#pragma unroll
        for (int i = 0; i < 7; i++) {
            A[2 * i] = wd[i];
            A[2 * i + 1] = __funnelshift_r(wd[i], wd[i + 1], 16);
        }
        short *out = TAPS == 8 ? &s.tile[u][(8 * hv + 1) * TP + 8 + x] : &s.ctile[u][x];
        constexpr int OP = TAPS == 8 ? TP : 8;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int v;
            if (TAPS == 8)
                v = __dp2a_lo((int)A[j], (int)m.vf0, __dp2a_hi((int)A[j + 2], (int)m.vf0,
                    __dp2a_lo((int)A[j + 4], (int)m.vf1, __dp2a_hi((int)A[j + 6], (int)m.vf1, 0))));
            else
                v = __dp2a_lo((int)A[j], (int)m.vf0, __dp2a_hi((int)A[j + 2], (int)m.vf0, 0));
            if (8 * hv + j < bh)
                out[j * OP] = (short)(v >> m.shv);
        }
    }
}

__device__ __forceinline__ int vsign(int v) { return v < 0 ? -1 : (v != 0); }

__device__ int parametric(const int *sd, int stride)      // parametric_mv_refine, vvc_inter.c:642-681
{
    const int sm = sd[-stride], sc = sd[0], sp = sd[stride];
    int denom = ((sm + sp) - (sc << 1)) << 3;
    if (!denom) return 0;
    if (sm == sc) return -8;
    if (sp == sc) return 8;
    int num = (sm - sp) * 16, neg = 0, q = 0;
    if (num < 0) { num = -num; neg = 1; }
    for (int i = 0; i < 3; i++) {
        q <<= 1;
        if (num >= denom) { num -= denom; q++; }
        denom >>= 1;
    }
    return neg ? -q : q;
}

// 4 luma samples of tile `ui` at (x, y)
__device__ __forceinline__ void tile4(const WarpSmem &s, int ui, int x, int y, int v[4])
{
    const uint2 q = *reinterpret_cast<const uint2 *>(&s.tile[ui][(y + 1) * TP + 8 + x]);
    v[0] = lo16(q.x); v[1] = hi16(q.x); v[2] = lo16(q.y); v[3] = hi16(q.y);
}

// ring of integer samples around a tile: bdof_fetch_samples / fetch_samples (vvc_inter_template.c:101-133)
__device__ __forceinline__ void fetch_ring(WarpSmem &s, int ui, int bw, int bh, int mx, int my, int lane)   //@region fetch_ring
{
    const UnitMC m = s.um[ui];
    const uint16_t *w16 = reinterpret_cast<const uint16_t *>(s.a.win + m.woff);
    const int xo = m.c0 + (mx >> 3) + 3, yo = m.r0 + (my >> 3) + 3;
    const int per = 2 * (bw + 2) + 2 * bh;
    for (int idx = lane; idx < per; idx += 32) {
        int tx, ty;
        if (idx < bw + 2)                 { tx = idx - 1;            ty = -1; }
        else if (idx < 2 * (bw + 2))      { tx = idx - (bw + 2) - 1; ty = bh; }
        else if (idx < 2 * (bw + 2) + bh) { tx = -1;                 ty = idx - 2 * (bw + 2); }
        else                              { tx = bw;                 ty = idx - 2 * (bw + 2) - bh; }
        s.tile[ui][(ty + 1) * TP + 8 + tx] = (short)(w16[(ty + yo) * (2 * PWL) + tx + xo] << 4);
    }
}

// The reference windows of the NEXT record of a warp (bi-predicted record of the DMVR / BDOF list) asked into L2 while the
// current one is computed: the staging loop's first use of its loads was where a quarter of the main kernel's stall
// samples sat (the reference ring of a 16-picture launch is three times the L2, so every window comes from DRAM).
// One row per lane and list, both ends of the row (a row may straddle two lines); coordinates clamped into the plane.
__device__ __noinline__ void prefetch_windows(const InterK &p, const uint32_t *__restrict__ coop, int ci, int lane)
{
    const uint32_t *q = reinterpret_cast<const uint32_t *>(p.pbs + __ldg(coop + ci));
    const uint32_t r0 = __ldg(q), r1 = __ldg(q + 1), r2 = __ldg(q + 2);
    const int x0 = r0 & 0xffff, y0 = r0 >> 16, w = r1 & 0xff, h = (r1 >> 8) & 0xff;
    const int l = lane >> 4, k = lane & 15;
    const int mx = (int)__ldg(q + 3 + 2 * l), my = (int)__ldg(q + 4 + 2 * l), ref = l ? (r2 >> 8) & 0xff : r2 & 0xff;
    {
        const pel *plane = p.ref[0] + ref * p.rb[0];
        const int xa = d_clip3(x0 + (mx >> 4) - 3, 0, p.w - 1), xb = d_clip3(x0 + (mx >> 4) + w + 3, 0, p.w - 1), ya = y0 + (my >> 4) - 3;
#pragma unroll
        for (int r = k; r < 23; r += 16)
            if (r < h + 7) {
                const pel *row = plane + (long long)d_clip3(ya + r, 0, p.h - 1) * p.rp[0];
                asm volatile("prefetch.global.L2 [%0];" :: "l"(row + xa));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(row + xb));
            }
    }
    if (p.planes == 3) {
        // chroma: lanes 0-7 / 8-15 of a list take Cb / Cr, rows k and k + 8 of the (h / 2 + 3)-row window (one line per row)
        const int pc = (k >> 3) + 1, kk = k & 7, pw = p.w >> 1, ph = p.h >> 1;
        const pel *plane = (pc == 1 ? p.ref[1] + ref * p.rb[1] : p.ref[2] + ref * p.rb[2]);
        const int pitch = pc == 1 ? p.rp[1] : p.rp[2];
        const int xa = d_clip3((x0 >> 1) + (mx >> 5) - 1, 0, pw - 1), ya = (y0 >> 1) + (my >> 5) - 1;
#pragma unroll
        for (int r = kk; r < 11; r += 8)
            if (r < (h >> 1) + 3)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(plane + (long long)d_clip3(ya + r, 0, ph - 1) * pitch + xa));
    }
}

// KIND 0: records with DMVR / BDOF and no PROF (always bi-predicted, never GPM): the uni, GPM and PROF paths are
// compiled out.  KIND 1: the rest of the cooperative records (PROF), listed from the back of coop[].
template <int TMA> struct TmaSmem { alignas(128) uint8_t box[kWarps][2][kBoxSlot]; uint64_t bar[kWarps]; };
template <> struct TmaSmem<0> { };

// Start the window copies of cooperative record `ci` if it is a DMVR record whose two unrefined windows lie inside the
// picture (warp-uniform decision; lane 0 issues).  Returns whether the copies are in flight.
template <int TMA>
__device__ __forceinline__ int tma_prefetch(TmaSmem<TMA> &ts, const InterK &p, const uint32_t *__restrict__ coop, int ci,
                                            const CUtensorMap *map, int warp, int lane)
{
    if constexpr (TMA) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p.pbs + __ldg(coop + ci));
        const uint32_t r0 = __ldg(q), r1 = __ldg(q + 1), r2 = __ldg(q + 2);
        const int x0 = r0 & 0xffff, y0 = r0 >> 16, w = r1 & 0xff, h = (r1 >> 8) & 0xff;
        if (!((r2 >> 24) & VVC_CUDA_PB_DMVR) || !((r1 >> 16) & VVC_CUDA_PB_LUMA))
            return 0;
        int wx[2], wy[2];
        bool inside = true;
#pragma unroll
        for (int l = 0; l < 2; l++) {
            wx[l] = x0 + ((int)__ldg(q + 3 + 2 * l) >> 4) - 3;
            wy[l] = y0 + ((int)__ldg(q + 4 + 2 * l) >> 4) - 3;
            inside = inside && wx[l] >= 0 && wx[l] + w + 7 <= p.w && wy[l] >= 0 && wy[l] + h + 7 <= p.h;
        }
        if (!inside)
            return 0;
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the staging area was just read by the warp
            mbar_expect_tx(&ts.bar[warp], 2 * kBoxBytes);
            tma_box_3d(ts.box[warp][0], map, wx[0] & ~7, wy[0], (int)(r2 & 0xff), &ts.bar[warp]);
            tma_box_3d(ts.box[warp][1], map, wx[1] & ~7, wy[1], (int)((r2 >> 8) & 0xff), &ts.bar[warp]);
        }
        return 1;
    } else {
        return 0;
    }
}

// KIND 0: records with DMVR / BDOF and no PROF (always bi-predicted, never GPM): the uni, GPM and PROF paths are
// compiled out.  KIND 1: the rest of the cooperative records (PROF), listed from the back of coop[].
// TMA 1 (KIND 0 only): the DMVR windows of the NEXT record are fetched by the copy engine while this one is computed.
// PHASE (KIND 0 only): 0 = the whole record in one kernel; 1 = the DMVR search alone (window staging, bilinear, SADs, decision),
// refined vectors and the BDOF switch go to `refined`; 2 = everything after it, reading them back.  Splitting halves the
// code each kernel's warps run through (the one-kernel form spends 38 % of its stall samples waiting for instructions).
template <int KIND, int TMA, int PHASE>
__global__ void __launch_bounds__(kThreads) inter_warp_kernel(const InterK p, const uint32_t *__restrict__ coop, int cap, uint32_t *count,   //@region rec_load
                                                              const __grid_constant__ CUtensorMap tmap, VVCCudaDmvrOut *__restrict__ refined)
{
    using SM = typename std::conditional<PHASE == 1, SearchSmem, WarpSmem>::type;
    __shared__ SM sm[kWarps];
    __shared__ TmaSmem<TMA> ts;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SM &s = sm[warp];
    const uint2 *lumaf = reinterpret_cast<const uint2 *>(&vvct_luma_mc_filters[0][0][0]);
    const uint32_t *chromaf = reinterpret_cast<const uint32_t *>(&vvct_chroma_mc_filters[0][0][0]);
    if constexpr (TMA) {
        if (lane == 0) {
            mbar_init(&ts.bar[warp], 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }

    // Records are drawn from a counter (count[8 + KIND], zeroed with the list counts), two at a time: DMVR's early
    // termination and the per-record BDOF switch make their cost uneven, and with a fixed stride the launch lasted as
    // long as its unluckiest warp.  The draw runs one record ahead of the computation so that the next record's windows
    // can be in flight.
    const int n_coop = (int)count[4 + KIND];
    int q_base = 0, q_k = 2, parity = 0;
#define DRAW_NEXT()                                                                 \
    do {                                                                            \
        if (q_k == 2) {                                                             \
            int f_ = 0;                                                             \
            if (lane == 0)                                                          \
                f_ = (int)atomicAdd(count + (PHASE == 1 ? 14 : 8 + KIND), 2u);      \
            q_base = __shfl_sync(0xffffffffu, f_, 0);                               \
            q_k = 0;                                                                \
        }                                                                           \
        nxt = q_base + q_k++;                                                       \
        if (nxt >= n_coop)                                                          \
            nxt = -1;                                                               \
        nxt_tma = nxt >= 0 ? tma_prefetch<TMA>(ts, p, coop, KIND ? cap - 1 - nxt : nxt, &tmap, warp, lane) : 0;   \
        if (INTER_WARP_PREFETCH && KIND == 0 && !TMA && nxt >= 0)                   \
            prefetch_windows(p, coop, nxt, lane);                                   \
    } while (0)
    int nxt, nxt_tma;
    DRAW_NEXT();
    while (nxt >= 0) {
    {
        const int ci = nxt, use_tma = nxt_tma;
        __syncwarp();
        const int ri = (int)__ldg(coop + (KIND ? cap - 1 - ci : ci));
        const Rec pb = load_rec(p.pbs + ri);
        const int w = pb.w, h = pb.h, lw = 31 - __clz(w);
        const bool gpm = KIND == 0 ? false : (pb.flags & VVC_CUDA_PB_GPM) != 0;
        const bool bi = KIND == 0 ? true : gpm || pb.pred == 3;
        const bool dmvr = (pb.flags & VVC_CUDA_PB_DMVR) != 0;
        const int lx = pb.pred - 1;                                  // the list of a uni record
        int mvr[2][2] = { { pb.mv[0][0], pb.mv[0][1] }, { pb.mv[1][0], pb.mv[1][1] } };   // vectors used for MC
        int bdof = (pb.flags & VVC_CUDA_PB_BDOF) ? 1 : 0;
        const bool dmvr_luma = dmvr && (pb.planes & VVC_CUDA_PB_LUMA);
#define MVR(l, c) ((l) ? mvr[1][c] : mvr[0][c])
#define MV0(l, c) ((l) ? pb.mv[1][c] : pb.mv[0][c])
#define REF(l)    ((l) ? pb.ref[1] : pb.ref[0])

        // ---- DMVR: stage both unrefined windows, bilinear prediction, 25 SADs, refinement -------------
        if constexpr (TMA) if (use_tma) {   //@region dmvr_stage
            // the copy engine has (or will have) put both windows into the staging area: move them into the padded
            // window layout (core at sample column 2, no parity offset) and replicate the apron
            mbar_wait(&ts.bar[warp], parity);
            parity ^= 1;
            const int l = lane >> 4, k = lane & 15, nw = (w + 8) >> 1;
            if (k < nw) {
                const int o = (pb.x0 + (MV0(l, 0) >> 4) - 3) & 7, sh = (o & 1) << 4;      // window column 0 inside the box
                const uint32_t *src = reinterpret_cast<const uint32_t *>(ts.box[warp][l]) + (o >> 1) + k;
                uint32_t *dst = s.a.win + l * WUL + 2 * PWL + 1 + k;
#pragma unroll 4
                for (int r = 0; r < h + 7; r++, src += kBoxW / 2, dst += PWL)
                    *dst = frc(src[0], src[1], sh);
            }
            stage_apron(s.a.win + l * WUL, PWL, 0, w + 7, h + 7, k, 16);
        }
        DRAW_NEXT();            // the staging area is free again: the next record's windows start now
        if (PHASE == 1 && !dmvr_luma)
            continue;
        if (PHASE == 2 && dmvr_luma) {
            // the search ran in its own kernel: windows again (the final motion compensation reads the same clamped windows),
            // refined vectors and BDOF switch from its output
            const int l = lane >> 4;
            stage_units(s.a.win + l * WUL, PWL, 1, p.ref[0] + REF(l) * p.rb[0], p.rp[0], p.w, p.h, p.margin,
                        pb.x0 + (MV0(l, 0) >> 4) - 3, pb.y0 + (MV0(l, 1) >> 4) - 3, w + 7, h + 7, lane & 15, 16);
            const int *r = reinterpret_cast<const int *>(refined + ri);
            mvr[0][0] = r[0]; mvr[0][1] = r[1]; mvr[1][0] = r[2]; mvr[1][1] = r[3];
            bdof = r[5];
            __syncwarp();
        }
        if (PHASE != 2 && dmvr_luma) {
            if (!(TMA && use_tma)) {
                const int l = lane >> 4;                    // lanes 0-15: list 0, lanes 16-31: list 1
                if (PHASE == 1)
                    // the search alone reads rows 1 .. h + 5 of the window (bilinear taps at -2 .. h + 2) and never its apron:
                    // the same placement as below, two rows and the apron passes fewer
                    stage_units(s.a.win + l * WUL + 3 * PWL + 1, PWL, 0, p.ref[0] + REF(l) * p.rb[0], p.rp[0], p.w, p.h, p.margin,
                                pb.x0 + (MV0(l, 0) >> 4) - 3, pb.y0 + (MV0(l, 1) >> 4) - 2, w + 7, h + 5, lane & 15, 16);
                else
                    stage_units(s.a.win + l * WUL, PWL, 1, p.ref[0] + REF(l) * p.rb[0], p.rp[0], p.w, p.h, p.margin,
                                pb.x0 + (MV0(l, 0) >> 4) - 3, pb.y0 + (MV0(l, 1) >> 4) - 3, w + 7, h + 7, lane & 15, 16);
            }
            __syncwarp();
            // bilinear (dmvr / dmvr_h / dmvr_v / dmvr_hv, vvc_inter_template.c:324-409) on sample pairs; lane = row
            const int nwo = (w + 4) >> 1;   //@region dmvr_bilinear
#pragma unroll 1
            for (int l = 0; l < 2; l++) {
                const int mx = MV0(l, 0) & 15, my = MV0(l, 1) & 15;
                const int fx0 = vvct_dmvr_filters[mx][0], fx1 = vvct_dmvr_filters[mx][1];
                const int fy0 = vvct_dmvr_filters[my][0], fy1 = vvct_dmvr_filters[my][1];
                const int e = (TMA && use_tma) ? 0 : (pb.x0 + (MV0(l, 0) >> 4) - 3) & 1, c0 = 3 + e, sh = (c0 & 1) << 4;
                const int y = min(lane, h + 4);
                const uint32_t *wp = s.a.win + l * WUL + (3 + y) * PWL + (c0 >> 1);
                uint32_t hrow[10];
                uint32_t prev = wp[0];
#pragma unroll
                for (int i = 0; i < 10; i++) {
                    hrow[i] = 0;
                    if (i < nwo) {
                        const uint32_t next = wp[i + 1];
                        const uint32_t P = frc(prev, next, sh), Q = frc(prev, next, sh + 16);
                        hrow[i] = ((fx0 * P + fx1 * Q + 0x00080008u) >> 4) & 0x0fff0fffu;
                        prev = next;
                    }
                }
                uint32_t *dst = reinterpret_cast<uint32_t *>(&s.dm[l][y * DP]);
#pragma unroll
                for (int i = 0; i < 10; i++) {
                    const uint32_t below = __shfl_down_sync(0xffffffffu, hrow[i], 1);
                    if (i < nwo && lane < h + 4)
                        dst[i] = ((fy0 * hrow[i] + fy1 * below + 0x00080008u) >> 4) & 0x0fff0fffu;
                }
            }
            __syncwarp();
            // SAD on every other row (vvc_sad, vvcdsp.c:49-65).  A task is (dy, row[, quarter of the row]); the 5 dx of a
            // dy share the loaded rows.  With 8 rows, dy = 0..3 fill the 32 lanes with whole rows and dy = 4 is split
            // into row quarters over all 32 lanes; with 4 rows the 20 (dy, row) tasks fit one pass.
            {
                const int lhr = 31 - __clz(h >> 1), hr = h >> 1, nwr = w >> 1;   //@region dmvr_sad
                const int passes = hr == 8 ? 2 : 1;
                for (int pass = 0; pass < passes; pass++) {
                    int dyi, yy, k0, k1;
                    bool act = true;
                    if (hr == 8 && pass == 1) { dyi = 4; yy = lane >> 2; k0 = (lane & 3) * (nwr >> 2); k1 = k0 + (nwr >> 2); }
                    else { dyi = lane >> lhr; yy = lane & (hr - 1); k0 = 0; k1 = nwr; act = dyi < (hr == 8 ? 4 : 5); }
                    if (!act) { dyi = 0; k1 = 0; }
                    const int y = yy << 1;
                    const uint32_t *ra = reinterpret_cast<const uint32_t *>(&s.dm[0][(dyi + y) * DP]);
                    const uint32_t *rb = reinterpret_cast<const uint32_t *>(&s.dm[1][(4 - dyi + y) * DP]);
                    uint32_t sd[5] = { 0, 0, 0, 0, 0 };         // per dx: two 16-bit partial sums (at most 8 x 1023 each)
                    uint32_t a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
                    if (k0 < k1) {
                        a0 = ra[k0]; a1 = ra[k0 + 1]; b0 = rb[k0]; b1 = rb[k0 + 1];
                    }
#pragma unroll
                    for (int kk = 0; kk < 8; kk++) {
                        const int k = k0 + kk;
                        if (k < k1) {
                            a2 = ra[k + 2]; b2 = rb[k + 2];
                            const uint32_t a01 = __funnelshift_r(a0, a1, 16), a12 = __funnelshift_r(a1, a2, 16);
                            const uint32_t b01 = __funnelshift_r(b0, b1, 16), b12 = __funnelshift_r(b1, b2, 16);
#define SADW(acc, u, v) acc = acc + __vmaxu2(u, v) - __vminu2(u, v)
                            SADW(sd[0], a0, b2);               // dx = -2: a at sample 0, b at sample 4
                            SADW(sd[1], a01, b12);             // dx = -1: a at 1, b at 3
                            SADW(sd[2], a1, b1);               // dx =  0
                            SADW(sd[3], a12, b01);             // dx = +1: a at 3, b at 1
                            SADW(sd[4], a2, b0);               // dx = +2
#undef SADW
                            a0 = a1; a1 = a2; b0 = b1; b1 = b2;
                        }
                    }
                    const int span = (hr == 8 && pass == 1) ? 32 : hr;
#pragma unroll
                    for (int d = 0; d < 5; d++) {
                        int v = (int)((sd[d] & 0xffff) + (sd[d] >> 16));
                        for (int o = 1; o < span; o <<= 1)
                            v += __shfl_xor_sync(0xffffffffu, v, o);
                        if (act && !(lane & (span - 1)))
                            s.sad[dyi * 5 + d] = v;
                    }
                }
            }
            __syncwarp();
            // decision (dmvr_mv_refine, vvc_inter.c:700-747), computed redundantly by every lane
            {
                const int mine = s.sad[lane < 25 ? lane : 0];   //@region dmvr_decision
                const int centre = __shfl_sync(0xffffffffu, mine, 12);
                int min_sad = centre - (centre >> 2);
                if (min_sad >= w * h) {
                    __syncwarp();
                    if (lane == 12)
                        s.sad[12] = min_sad;
                    // first strict minimum in scan order, the centre first (ties keep the earlier candidate)
                    uint32_t key = lane == 12 ? (uint32_t)min_sad << 5 : lane < 25 ? ((uint32_t)mine << 5) | (lane + 1) : 0xffffffffu;
#pragma unroll
                    for (int o = 16; o; o >>= 1)
                        key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
                    const int pos = (key & 31) ? (int)(key & 31) - 1 : 12;
                    min_sad = (int)(key >> 5);
                    const int min_dx = pos % 5, min_dy = pos / 5;
                    int dmv0 = (min_dx - 2) * 16, dmv1 = (min_dy - 2) * 16;
                    __syncwarp();
                    if (min_dx != 0 && min_dx != 4 && min_dy != 0 && min_dy != 4) {
                        dmv0 += parametric(&s.sad[pos], 1);
                        dmv1 += parametric(&s.sad[pos], 5);
                    }
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        mvr[i][0] = d_clip3(mvr[i][0] + (1 - 2 * i) * dmv0, -(1 << 17), (1 << 17) - 1);
                        mvr[i][1] = d_clip3(mvr[i][1] + (1 - 2 * i) * dmv1, -(1 << 17), (1 << 17) - 1);
                    }
                }
                if (min_sad < 2 * w * h)
                    bdof = 0;
                if ((p.dmvr_out || PHASE == 1) && lane == 0) {
                    VVCCudaDmvrOut o;
                    o.mv[0][0] = mvr[0][0]; o.mv[0][1] = mvr[0][1]; o.mv[1][0] = mvr[1][0]; o.mv[1][1] = mvr[1][1];
                    o.min_sad = min_sad; o.bdof_applied = bdof;
                    if (p.dmvr_out)
                        p.dmvr_out[ri] = o;
                    if (PHASE == 1)
                        refined[ri] = o;
                }
            }
            __syncwarp();
        }
        if constexpr (PHASE != 1) {
        const bool do_bdof = bdof && !gpm;   //@region luma_setup
        pel *dstp[3];
#pragma unroll
        for (int c = 0; c < 3; c++)
            dstp[c] = p.dst[c] + pb.pic * p.db[c];

        // ---- luma ------------------------------------------------------------------------------------
        if (pb.planes & VVC_CUDA_PB_LUMA) {
            const int n_units = bi ? 2 : 1;
            {   // unit descriptors; unit i is list i (bi / GPM) or the single list (uni)
                const int u = lane & 1, list = bi ? u : lx;
                const int mvx = MVR(list, 0), mvy = MVR(list, 1), m0x = MV0(list, 0), m0y = MV0(list, 1);
                const int mx = mvx & 15, my = mvy & 15, filt = gpm ? 0 : pb.filt;
                UnitMC m;
                m.woff = u * WUL;
                if (dmvr) {
                    const int e = (TMA && use_tma) ? 0 : (pb.x0 + (m0x >> 4) - 3) & 1;
                    m.c0 = 2 + e + d_clip3((mvx >> 4) - (m0x >> 4), -2, 2);
                    m.r0 = 2 + d_clip3((mvy >> 4) - (m0y >> 4), -2, 2);
                } else {
                    m.c0 = (pb.x0 + (mvx >> 4) - 3) & 1;
                    m.r0 = 0;
                }
                const uint2 fh = lumaf[filt * 16 + mx], fv = lumaf[filt * 16 + my];
                m.hf0 = mx ? fh.x : 0x01000000u;  m.hf1 = mx ? fh.y : 0u;  m.shh = mx ? 2 : 0;
                m.vf0 = my ? fv.x : (mx ? 0x01000000u : 0x10000000u);  m.vf1 = my ? fv.y : 0u;
                m.shv = my ? (mx ? 6 : 2) : 0;
                if (lane < n_units)
                    s.um[lane] = m;
                if (!dmvr_luma) {
                    const int i = bi ? lane >> 4 : 0, li = bi ? i : lx;
                    stage_units(s.a.win + i * WUL, PWL, 0, p.ref[0] + REF(li) * p.rb[0], p.rp[0], p.w, p.h, p.margin,
                                pb.x0 + (MVR(li, 0) >> 4) - 3, pb.y0 + (MVR(li, 1) >> 4) - 3, w + 7, h + 7,
                                bi ? lane & 15 : lane, bi ? 16 : 32);
                }
            }
            __syncwarp();
            pass_h<8>(s, lane, n_units, h + 7, w > 8 ? 1 : 0);   //@region luma_passes
            __syncwarp();
            pass_v<8>(s, lane, n_units, lw, h);
            __syncwarp();
            pel *d = dstp[0] + (long long)pb.y0 * p.dp[0] + pb.x0;   //@region luma_final_uni
            const int ntask = h << (lw - 2);
            if (!bi) {
                const UniW uw = uni_weights(pb, p.wp, lx, 0);
                const bool use_prof = pb.flags & (lx ? VVC_CUDA_PB_PROF1 : VVC_CUDA_PB_PROF0);
                if (use_prof) {                             // luma_prof_uni, vvc_inter.c:368-408 (4x4 blocks)
                    fetch_ring(s, 0, w, h, MVR(lx, 0) & 15, MVR(lx, 1) & 15, lane);
                    __syncwarp();
                    if (lane < 16) {
                        const int x = lane & 3, y = lane >> 2;
                        const VVCCudaProf *pr = p.prof + pb.prof;
                        const short *q = &s.tile[0][(y + 1) * TP + 8 + x];
                        const int gh = (short)((q[1] >> 6) - (q[-1] >> 6)), gv = (short)((q[TP] >> 6) - (q[-TP] >> 6));
                        const int di = gh * pr->diff_mv_x[lx][lane] + gv * pr->diff_mv_y[lx][lane];
                        d[(long long)y * p.dp[0] + x] = (pel)finish_uni(q[0] + d_clip3(di, -8192, 8191), uw);
                    }
                } else {
                    for (int t = lane; t < ntask; t += 32) {
                        const int x = (t & ((w >> 2) - 1)) << 2, y = t >> (lw - 2);
                        int a[4];
                        tile4(s, 0, x, y, a);
                        uint2 o;
                        o.x = pack16(finish_uni(a[0], uw), finish_uni(a[1], uw));
                        o.y = pack16(finish_uni(a[2], uw), finish_uni(a[3], uw));
                        *reinterpret_cast<uint2 *>(d + (long long)y * p.dp[0] + x) = o;
                    }
                }
            } else {
                const int prof_mask = (KIND == 0 || gpm) ? 0 : pb.flags & (VVC_CUDA_PB_PROF0 | VVC_CUDA_PB_PROF1);   //@region bi_ring_prof
                if (do_bdof || prof_mask) {
                    fetch_ring(s, 0, w, h, mvr[0][0] & 15, mvr[0][1] & 15, lane);
                    fetch_ring(s, 1, w, h, mvr[1][0] & 15, mvr[1][1] & 15, lane);
                    __syncwarp();
                }
                if (prof_mask) {                            // luma_prof_bi, vvc_inter.c:410-446 (4x4 blocks)
                    const int ui = lane >> 4, e = lane & 15, x = e & 3, y = e >> 2;
                    const bool act = prof_mask & (ui ? VVC_CUDA_PB_PROF1 : VVC_CUDA_PB_PROF0);
                    int val = 0;
                    if (act) {
                        const VVCCudaProf *pr = p.prof + pb.prof;
                        const short *q = &s.tile[ui][(y + 1) * TP + 8 + x];
                        const int gh = (short)((q[1] >> 6) - (q[-1] >> 6)), gv = (short)((q[TP] >> 6) - (q[-TP] >> 6));
                        const int di = gh * pr->diff_mv_x[ui][e] + gv * pr->diff_mv_y[ui][e];
                        val = q[0] + d_clip3(di, -8192, 8191);
                    }
                    __syncwarp();
                    if (act)
                        s.tile[ui][(y + 1) * TP + 8 + x] = (short)val;
                    __syncwarp();
                }
                if (gpm) {                                  // put_gpm, vvc_inter_template.c:78-98
                    const uint8_t *wt = &vvct_gpm_weights[0][0] + pb.gw;   //@region gpm
                    for (int t = lane; t < ntask; t += 32) {
                        const int x = (t & ((w >> 2) - 1)) << 2, y = t >> (lw - 2);
                        int a[4], b[4], o[4];
                        tile4(s, 0, x, y, a);
                        tile4(s, 1, x, y, b);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int g = wt[y * pb.gsy + (x + i) * pb.gsx];
                            o[i] = d_clip_pel((a[i] * g + b[i] * (8 - g) + 64) >> 7, 10);
                        }
                        *reinterpret_cast<uint2 *>(d + (long long)y * p.dp[0] + x) = make_uint2(pack16(o[0], o[1]), pack16(o[2], o[3]));
                    }
                } else if (do_bdof) {                       // apply_bdof, vvc_inter_template.c:237-317
                    // gradients (prof_grad_filter :135-158) of both tiles; lane = (list, row), samples in pairs.
                    // The padded borders the reference creates with pad_int16 equal clamped coordinates below.
                    uint32_t gh[8], gv[8];   //@region bdof_grad
                    {
                        const int l = lane >= h, y = lane - l * h;
                        const bool act = lane < 2 * h;
                        const int yy = act ? y : 0;
                        const uint32_t *up = reinterpret_cast<const uint32_t *>(&s.tile[l][yy * TP]);
                        const uint32_t *mid = up + TP / 2, *dn = mid + TP / 2;
                        const int nwr = w >> 1;
                        uint32_t prevS = ((mid[3] ^ 0x80008000u) >> 6) & 0x03ff03ffu;
                        uint32_t curS = ((mid[4] ^ 0x80008000u) >> 6) & 0x03ff03ffu;
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            gh[k] = gv[k] = 0;
                            if (k < nwr) {
                                const uint32_t nextS = ((mid[5 + k] ^ 0x80008000u) >> 6) & 0x03ff03ffu;
                                gh[k] = __vsub2(__funnelshift_r(curS, nextS, 16), __funnelshift_r(prevS, curS, 16));
                                const uint32_t su = ((up[4 + k] ^ 0x80008000u) >> 6) & 0x03ff03ffu;
                                const uint32_t sd = ((dn[4 + k] ^ 0x80008000u) >> 6) & 0x03ff03ffu;
                                gv[k] = __vsub2(sd, su);
                                prevS = curS; curS = nextS;
                            }
                        }
                    }
                    __syncwarp();                           // windows and first-pass rows are dead: reuse as gradient storage
                    if (lane < 2 * h) {
                        const int l = lane >= h, y = lane - l * h;
                        uint32_t *g0 = reinterpret_cast<uint32_t *>(&s.b.grad[l * 2][y * GP]);
                        uint32_t *g1 = reinterpret_cast<uint32_t *>(&s.b.grad[l * 2 + 1][y * GP]);
#pragma unroll
                        for (int k = 0; k < 8; k++)
                            if (k < (w >> 1)) { g0[k] = gh[k]; g1[k] = gv[k]; }
                    }
                    __syncwarp();
                    // derive_bdof_vx_vy :237-265: 6x6 window sums per 4x4 block, separable.  Column pass:
                    const int lbx = lw - 2;   //@region bdof_sums
                    for (int t = lane; t < (h << lbx); t += 32) {
                        const int bxi = t & ((1 << lbx) - 1), y = t >> lbx;
                        int sgx2 = 0, sgy2 = 0, sgxgy = 0, sgxdi = 0, sgydi = 0;
#pragma unroll
                        for (int dx = -1; dx < 5; dx++) {
                            const int x = d_clip3(4 * bxi + dx, 0, w - 1);
                            const int ti = (y + 1) * TP + 8 + x, gi = y * GP + x;
                            const int diff = (s.tile[0][ti] >> 4) - (s.tile[1][ti] >> 4);
                            const int th = (s.b.grad[0][gi] + s.b.grad[2][gi]) >> 1;
                            const int tv = (s.b.grad[1][gi] + s.b.grad[3][gi]) >> 1;
                            sgx2 += abs(th); sgy2 += abs(tv);
                            sgxgy += vsign(tv) * th;
                            sgxdi -= vsign(th) * diff;
                            sgydi -= vsign(tv) * diff;
                        }
                        s.b.csum[0][y][bxi] = sgx2; s.b.csum[1][y][bxi] = sgy2; s.b.csum[2][y][bxi] = sgxgy;
                        s.b.csum[3][y][bxi] = sgxdi; s.b.csum[4][y][bxi] = sgydi;
                    }
                    __syncwarp();
                    if (lane < ((h >> 2) << lbx)) {         // row pass + vx, vy; lane = 4x4 block
                        const int bxi = lane & ((1 << lbx) - 1), byi = lane >> lbx;
                        int sum[5] = { 0, 0, 0, 0, 0 };
#pragma unroll
                        for (int dy = -1; dy < 5; dy++) {
                            const int y = d_clip3(4 * byi + dy, 0, h - 1);
#pragma unroll
                            for (int q = 0; q < 5; q++)
                                sum[q] += s.b.csum[q][y][bxi];
                        }
                        const int vx = sum[0] > 0 ? d_clip3((sum[3] * 4) >> d_ilog2(sum[0]), -15, 15) : 0;
                        const int vy = sum[1] > 0 ? d_clip3(((sum[4] * 4) - ((vx * sum[2]) >> 1)) >> d_ilog2(sum[1]), -15, 15) : 0;
                        s.vxy[lane][0] = vx; s.vxy[lane][1] = vy;
                    }
                    __syncwarp();
                    for (int t = lane; t < ntask; t += 32) {   // apply_bdof_min_block :267-286   //@region bdof_apply
                        const int x4 = t & ((w >> 2) - 1), x = x4 << 2, y = t >> (lw - 2);
                        const int vx = s.vxy[((y >> 2) << lbx) + x4][0], vy = s.vxy[((y >> 2) << lbx) + x4][1];
                        int a[4], b[4], o[4];
                        tile4(s, 0, x, y, a);
                        tile4(s, 1, x, y, b);
                        const uint2 g0h = *reinterpret_cast<const uint2 *>(&s.b.grad[0][y * GP + x]);
                        const uint2 g0v = *reinterpret_cast<const uint2 *>(&s.b.grad[1][y * GP + x]);
                        const uint2 g1h = *reinterpret_cast<const uint2 *>(&s.b.grad[2][y * GP + x]);
                        const uint2 g1v = *reinterpret_cast<const uint2 *>(&s.b.grad[3][y * GP + x]);
                        const int dh[4] = { lo16(g0h.x) - lo16(g1h.x), hi16(g0h.x) - hi16(g1h.x), lo16(g0h.y) - lo16(g1h.y), hi16(g0h.y) - hi16(g1h.y) };
                        const int dv[4] = { lo16(g0v.x) - lo16(g1v.x), hi16(g0v.x) - hi16(g1v.x), lo16(g0v.y) - lo16(g1v.y), hi16(g0v.y) - hi16(g1v.y) };
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            o[i] = d_clip_pel((a[i] + 16 + b[i] + vx * dh[i] + vy * dv[i]) >> 5, 10);
                        *reinterpret_cast<uint2 *>(d + (long long)y * p.dp[0] + x) = make_uint2(pack16(o[0], o[1]), pack16(o[2], o[3]));
                    }
                } else {
                    const Weights wt = bi_weights(pb, p.wp, 0);   //@region bi_avg
                    for (int t = lane; t < ntask; t += 32) {
                        const int x = (t & ((w >> 2) - 1)) << 2, y = t >> (lw - 2);
                        int a[4], b[4];
                        tile4(s, 0, x, y, a);
                        tile4(s, 1, x, y, b);
                        uint2 o;
                        o.x = pack16(combine_bi(a[0], b[0], wt), combine_bi(a[1], b[1], wt));
                        o.y = pack16(combine_bi(a[2], b[2], wt), combine_bi(a[3], b[3], wt));
                        *reinterpret_cast<uint2 *>(d + (long long)y * p.dp[0] + x) = o;
                    }
                }
            }
        }

        // ---- chroma (both planes, 4:2:0) -----------------------------------------------------------------
        if ((pb.planes & VVC_CUDA_PB_CHROMA) && p.planes == 3) {
            __syncwarp();
            const int bw = w >> 1, bh = h >> 1, lbw = lw - 1;   //@region chroma_setup_stage
            const int x0 = pb.x0 >> 1, y0 = pb.y0 >> 1, pw = p.w >> 1, ph = p.h >> 1;
            const int n_units = bi ? 4 : 2;
            {   // unit u: bi -> (plane u >> 1, list u & 1); uni -> (plane u, the single list)
                const int u = lane & 3, list = bi ? (u & 1) : lx;
                const int mvx = MVR(list, 0), mvy = MVR(list, 1), m0x = MV0(list, 0), m0y = MV0(list, 1);
                const int mx = mvx & 31, my = mvy & 31;
                UnitMC m;
                m.woff = u * WUC;
                if (dmvr) {
                    const int e = (x0 + (m0x >> 5) - 1) & 1;
                    m.c0 = 2 + e + d_clip3((mvx >> 5) - (m0x >> 5), -2, 2);
                    m.r0 = 2 + d_clip3((mvy >> 5) - (m0y >> 5), -2, 2);
                } else {
                    m.c0 = (x0 + (mvx >> 5) - 1) & 1;
                    m.r0 = 0;
                }
                m.hf0 = mx ? chromaf[mx] : 0x00000100u;  m.hf1 = 0;  m.shh = mx ? 2 : 0;
                m.vf0 = my ? chromaf[my] : (mx ? 0x00000100u : 0x00001000u);  m.vf1 = 0;
                m.shv = my ? (mx ? 6 : 2) : 0;
                if (lane < n_units)
                    s.um[lane] = m;
            }
            {   // bi: 4 units of 8 lanes, uni: 2 units of 16 lanes
                const int u = bi ? lane >> 3 : lane >> 4;
                const int list = bi ? (u & 1) : lx, pc = bi ? (u >> 1) : u;
                const pel *plane = (pc ? p.ref[2] + REF(list) * p.rb[2] : p.ref[1] + REF(list) * p.rb[1]);
                stage_units(s.a.win + u * WUC, PWC, dmvr ? 1 : 0, plane, pc ? p.rp[2] : p.rp[1], pw, ph, p.margin >> 1,
                            x0 + ((dmvr ? MV0(list, 0) : MVR(list, 0)) >> 5) - 1, y0 + ((dmvr ? MV0(list, 1) : MVR(list, 1)) >> 5) - 1,
                            bw + 3, bh + 3, bi ? lane & 7 : lane & 15, bi ? 8 : 16);
            }
            __syncwarp();
            pass_h<4>(s, lane, n_units, bh + 3, 0);   //@region chroma_passes
            __syncwarp();
            pass_v<4>(s, lane, n_units, lbw, bh);
            __syncwarp();
            // lane task = (plane, row, sample pair)
            const int lpw = lbw - 1, ntask = (2 * bh) << lpw;   //@region chroma_final
            const uint8_t *wt = &vvct_gpm_weights[0][0] + pb.gw;
            for (int t = lane; t < ntask; t += 32) {
                const int x = (t & ((1 << lpw) - 1)) << 1, q = t >> lpw;
                const int pc = q >= bh, y = q - pc * bh;
                int o[2];
                if (!bi) {
                    const UniW uw = uni_weights(pb, p.wp, lx, pc + 1);
                    const uint32_t a = *reinterpret_cast<const uint32_t *>(&s.ctile[pc][y * 8 + x]);
                    o[0] = finish_uni(lo16(a), uw); o[1] = finish_uni(hi16(a), uw);
                } else {
                    const uint32_t a = *reinterpret_cast<const uint32_t *>(&s.ctile[pc * 2][y * 8 + x]);
                    const uint32_t b = *reinterpret_cast<const uint32_t *>(&s.ctile[pc * 2 + 1][y * 8 + x]);
                    if (gpm) {
                        const int g0 = wt[y * 2 * pb.gsy + x * 2 * pb.gsx], g1 = wt[y * 2 * pb.gsy + (x + 1) * 2 * pb.gsx];
                        o[0] = d_clip_pel((lo16(a) * g0 + lo16(b) * (8 - g0) + 64) >> 7, 10);
                        o[1] = d_clip_pel((hi16(a) * g1 + hi16(b) * (8 - g1) + 64) >> 7, 10);
                    } else {
                        const Weights bw_ = bi_weights(pb, p.wp, pc + 1);
                        o[0] = combine_bi(lo16(a), lo16(b), bw_); o[1] = combine_bi(hi16(a), hi16(b), bw_);
                    }
                }
                pel *d = (pc ? dstp[2] + (long long)(y0 + y) * p.dp[2] : dstp[1] + (long long)(y0 + y) * p.dp[1]) + x0 + x;
                *reinterpret_cast<uint32_t *>(d) = pack16(o[0], o[1]);
            }
        }
        }       // PHASE != 1
#undef MVR
#undef MV0
#undef REF
    }
    }
#undef DRAW_NEXT
}

}  // namespace

// The luma plane of the reference ring as a 3-D tensor (x, y, picture) of 16-bit samples, box = one DMVR window.
// cuTensorMapEncodeTiled is a host-side encoder in the driver library; it is fetched through the runtime so that the
// library keeps linking against libcudart only.
static int make_window_map(VVCCudaCtx *ctx, const InterK &p, CUtensorMap *map)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
            return vvc_ctx_fail(ctx, VVC_CUDA_ERR_CUDA, "inter: cuTensorMapEncodeTiled is not available in this driver");
        encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[3] = { (cuuint64_t)p.w, (cuuint64_t)p.h, (cuuint64_t)p.nref };
    const cuuint64_t strides[2] = { (cuuint64_t)p.rp[0] * 2, (cuuint64_t)p.rb[0] * 2 };
    const cuuint32_t box[3] = { kBoxW, kBoxH, 1 }, estr[3] = { 1, 1, 1 };
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, (void *)p.ref[0], dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_CUDA, "inter: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return VVC_CUDA_OK;
}

// spread == 2 (long launches): the PROF kernel goes to side stream 2 behind the chroma patch kernels
int vvc_inter_launch_warp(VVCCudaCtx *ctx, const InterK &p, const InterLists &lists, int spread)
{
#ifndef INTER_WARP_CTAS
#define INTER_WARP_CTAS 7                                        // persistent: 7 CTAs fit an SM (shared memory)
#endif
#ifndef INTER_WARP_SPLIT
#define INTER_WARP_SPLIT 1                                       // DMVR search and the rest of a record as two kernels (tools/sweep_split.sh: inter 2.41 -> 2.31 ms per 16 pictures)
#endif
#ifndef INTER_WARP_CTAS_SEARCH
#define INTER_WARP_CTAS_SEARCH 9                                 // 22.8 KB of shared memory per CTA
#endif
#ifndef INTER_WARP_CTAS_TMA
#define INTER_WARP_CTAS_TMA 5                                    // with the staging area: 5
#endif
    const int ctas = ceil_div(p.n, kWarps);
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    VVCCudaDmvrOut *refined = reinterpret_cast<VVCCudaDmvrOut *>(lists.tail);      // behind the lists
    if (ctx->inter_tma) {
        if (make_window_map(ctx, p, &map))
            return ctx->err;
        const int grid = ctas < 148 * INTER_WARP_CTAS_TMA ? ctas : 148 * INTER_WARP_CTAS_TMA;
        inter_warp_kernel<0, 1, 0><<<grid, kThreads, 0, ctx->stream>>>(p, lists.coop, p.n, lists.count, map, refined);
    } else if (INTER_WARP_SPLIT) {
        const int g1 = ctas < 148 * INTER_WARP_CTAS_SEARCH ? ctas : 148 * INTER_WARP_CTAS_SEARCH;
        inter_warp_kernel<0, 0, 1><<<g1, kThreads, 0, ctx->stream>>>(p, lists.coop, p.n, lists.count, map, refined);
        VVC_LAUNCHED(ctx);
        const int grid = ctas < 148 * INTER_WARP_CTAS ? ctas : 148 * INTER_WARP_CTAS;
        inter_warp_kernel<0, 0, 2><<<grid, kThreads, 0, ctx->stream>>>(p, lists.coop, p.n, lists.count, map, refined);
    } else {
        const int grid = ctas < 148 * INTER_WARP_CTAS ? ctas : 148 * INTER_WARP_CTAS;
        inter_warp_kernel<0, 0, 0><<<grid, kThreads, 0, ctx->stream>>>(p, lists.coop, p.n, lists.count, map, refined);
    }
    VVC_LAUNCHED(ctx);
    const int grid = ctas < 148 * INTER_WARP_CTAS ? ctas : 148 * INTER_WARP_CTAS;
    inter_warp_kernel<1, 0, 0><<<grid, kThreads, 0, spread == 2 ? ctx->side[2] : ctx->stream>>>(p, lists.coop, p.n, lists.count, map, refined);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}
