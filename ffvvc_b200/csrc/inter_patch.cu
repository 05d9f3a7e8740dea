// Inter prediction for 10-bit pictures, thread-per-patch kernel (sm_100a): every record that needs no
// cooperative tile -- plain uni / bi prediction, BCW and explicit weights, GPM -- which is most of a picture.
//
// Reference semantics: pred_regular_luma / pred_regular_chroma (libavcodec/vvc/vvc_inter.c:545-639),
// pred_gpm_blk :466-521, the edge emulation :33-58, and the table entries put / put_uni / put_uni_w
// (libavcodec/h26x/h2656_inter_template.c:29-577), avg, w_avg, put_gpm (libavcodec/vvc/vvc_inter_template.c
// :25-98).
//
// B200 design: at most 8 lanes per record, one lane per patch of 4 columns x 8 rows.  A lane runs the whole separable
// filter in registers: per window row 12 samples (six words), the horizontal 8-tap sums as IDP.2A on sample pairs, and a
// streaming vertical pass that keeps only the last 7 row-pairs per column.  No per-record scalar work replicated over a
// warp -- about 25 thread-instructions per predicted sample.
//
// Window staging (PATCH_STAGED): what bounds the register-staged form is not arithmetic but the L1: every lane fetched its
// own 12 x 15 window with 90 scattered 4-byte loads (3.5 L1 wavefronts and 11 sectors per request, ncu).  Now the lanes
// of one row of patches of a record (4 lanes for w = 16, 2 for w = 8) own ONE window region in shared memory, the union
// of their windows, and fill it together by cp.async (LDGSTS, no destination register, no scoreboard wait): one
// 16-byte request per lane and window row for w = 16 (15 requests per window instead of 90), three 8-byte ones for
// w = 8; w = 4 records keep a private slot.  A region is refilled in place: as soon as the group has read row r of the
// current window, row r of the NEXT window (the second list of a bi record, or the first window of the lanes' next
// tasks) is requested into the same row, so every row is in flight for one whole window of arithmetic before it is read.
// Records whose windows leave the picture (or its margin) keep the register-staged form with clamped loads.
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
// resident CTAs per SM the class kernels are compiled for (register budget = 65536 / (128 * n)): profiles/r02_sweep_patch_staged.txt.
// The chroma kernels run on a side stream beside the luma and warp kernels: more of their CTAs only take room from those.
#ifndef PATCH_MB_LU
#define PATCH_MB_LU 3
#endif
#ifndef PATCH_MB_LB
#define PATCH_MB_LB 3
#endif
#ifndef PATCH_MB_CU
#define PATCH_MB_CU 4
#endif
#ifndef PATCH_MB_CB
#define PATCH_MB_CB 4
#endif
#ifndef PATCH_STAGED
#define PATCH_STAGED 1
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

// put_gpm for one row of a patch (vvc_inter_template.c:78-98): per-sample weights g / 8 - g, rounding 64 >> 7.  Out of line on
// purpose: inlined, the compiler predicates these forty instructions into every row of every bi record.
__device__ __noinline__ uint2 gpm_row(uint32_t k0, uint32_t k1, int o0, int o1, int o2, int o3, const uint8_t *g, int gsx, int cmax)
{
    const int a[4] = { (short)(k0 & 0xffff), (int)k0 >> 16, (short)(k1 & 0xffff), (int)k1 >> 16 }, o[4] = { o0, o1, o2, o3 };
    int out[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int w0 = g[min(c, cmax) * gsx];
        out[c] = d_clip_pel((a[c] * w0 + (short)o[c] * (8 - w0) + 64) >> 7, 10);
    }
    return make_uint2((uint32_t)out[0] | ((uint32_t)out[1] << 16), (uint32_t)out[2] | ((uint32_t)out[3] << 16));
}

// ---- staged windows ------------------------------------------------------------------------------------------
// Every warp has its own area of 32 slots.  MODE 0: registers (clamped loads allowed).  MODE 1: a private slot per lane,
// word (row r, i) of lane t at word [(r * NW + i) * 32 + t] of the area (4-byte requests).  MODE 2 / 3: a region per group of 2 / 4 lanes = one row of patches
// of a record 8 / 16 wide, rows of 40 / 64 bytes starting at the 8 / 16-byte aligned sample gx <= x - 3; lane k of the
// group requests chunks k, k + 2, k + 4 (8 bytes each) / chunk k (16 bytes) of every row.
__host__ __device__ constexpr int mode_lanes(int m)  { return m == 3 ? 4 : m == 2 ? 2 : 1; }
__host__ __device__ constexpr int mode_chunk(int m)  { return m == 3 ? 8 : m == 2 ? 4 : 2; }      // samples per request
__host__ __device__ constexpr int mode_pitch(int m)  { return m == 3 ? 64 : 40; }                 // bytes per window row in a region
__host__ __device__ constexpr int mode_region(int m) { return m == 3 ? 992 : 616; }               // bytes per region: 15 rows, and a stride that spreads the groups of a warp over the banks
constexpr int kSlotBytes = 15 * 6 * 4;                                        // per lane, all modes

struct Win { const char *src; int pitch, nr, nch; };    // the lane's first request of row 0, bytes per plane row, rows to stage (0: none), requests per row (MODE 2)

__device__ __forceinline__ void cp_async4(uint32_t saddr, const void *g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(saddr), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void *g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(saddr), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(saddr), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory"); return v; }

// The lane's share of a window whose first patch (column 0 of the row of patches) starts at reference sample (x, y);
// k = the lane's place in its group.
template <int TAPS, int MODE>
__device__ __forceinline__ Win win_of(const pel *plane, int pitch, int x, int y, int nrows, int k)
{
    constexpr int B = TAPS / 2 - 1, CS = mode_chunk(MODE);
    const int gx = (x - B) & ~(CS - 1);
    Win w;
    w.src = reinterpret_cast<const char *>(plane + (long long)(y - B) * pitch + gx + k * CS);
    w.pitch = pitch * 2;
    w.nr = nrows + TAPS - 1;
    w.nch = MODE == 2 ? 3 - k : 1;
    return w;
}

// sm = the lane's staging address: its slot (MODE 1) or region + k * chunk bytes
template <int TAPS, int MODE>
__device__ __forceinline__ void stage_row(uint32_t sm, const Win &w, int r)
{
    constexpr int NW = TAPS == 8 ? 6 : 4;
    if (r < w.nr) {
        const char *g = w.src + (long long)r * w.pitch;
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < NW; i++)
                cp_async4(sm + (r * NW + i) * 128, g + 4 * i);
        } else if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 3; j++)
                if (j < w.nch)
                    cp_async8(sm + r * mode_pitch(2) + j * 16, g + j * 16);
        } else {
            cp_async16(sm + r * mode_pitch(3), g);
        }
    }
}

// a group's first window: every row requested at once, one cp.async group per row, as the rolling refill commits them
template <int TAPS, int MODE>
__device__ __forceinline__ void stage_all(uint32_t sm, const Win &w)
{
    constexpr int NR = 8 + TAPS - 1;
#pragma unroll
    for (int r = 0; r < NR; r++) {
        stage_row<TAPS, MODE>(sm, w, r);
        cp_commit();
    }
}

// One list of one patch, 4 columns x 8 rows: the 14-bit intermediate prediction put() produces.  (x, y) =
// integer position of the patch's first sample in the reference plane; rows >= nrows are not computed.
// KEEP: the rows are kept as int16 pairs in keep[] (first list of a bi record).  Otherwise each finished row is
// blended -- with keep[] when BI -- and stored at d: clip(((a * w0 + b * w1 + off) >> sh) + ox); gw != NULL
// gives per-sample GPM weights (w0 = g, w1 = 8 - g; first sample's address, steps gsx / gsy, columns clamped
// to cmax).
// MODE > 0: the window was requested one window earlier into shared memory (always inside the plane or its margin): the lane
// reads its 12 samples of row r at sm_rd, and as soon as its group has read the row, the row is refilled
// with row r of `next` through sm_st.
//
// Code size: the row loop is ROLLED - two passes (q) over a body of eight rows (j), every index into the rings static inside
// the body: row r = 8 q + j writes the row pair P[(r - 1) & 7] and finishes output row yo = r - (TAPS - 1) = (j + 9 - TAPS) & 7
// from P[yo & 7], P[(yo + 2) & 7] ...; and the two lists of a bi record run through the same body (keep_pass = the first
// list, whose rows are kept).  Fully unrolled, the bi luma kernel was 4300 instructions of straight-line code per window
// pair and a quarter of its stall samples waited for instruction fetch.
template <int TAPS, bool BI, int MODE = 0>
__device__ __forceinline__ void mc_patch(const pel *plane, int pitch, int W, int H, int margin, int x, int y, int nrows,
                                         uint32_t hf0, uint32_t hf1, uint32_t vf0, uint32_t vf1, int shh, int shv,
                                         bool keep_pass, uint32_t keep[16], const Blend &k, const uint8_t *gw, int gsx, int gsy, int cmax,
                                         pel *d, int dpitch, bool two_words, uint32_t sm_rd = 0, uint32_t sm_st = 0, const Win &next = Win())
{
    constexpr bool STAGED = MODE > 0;
    constexpr int B = TAPS / 2 - 1, NW = TAPS == 8 ? 6 : 4, NR = 8 + TAPS - 1;
    const int e = (x - B) & 1, bx = x - B - e, sh = e << 4, y0 = y - B;
    // pre-padded reference planes (margin > 0) hold the clamped samples themselves: the plain loads may reach into the margin
    const bool inside = bx >= -margin && bx + 2 * NW <= W + margin && y0 >= -margin && y0 + nrows + TAPS - 1 <= H + margin;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(plane + (long long)y0 * pitch + bx);
    const int wpitch = pitch >> 1;
    uint32_t P[8][4];
    int prev[4] = { 0, 0, 0, 0 };
    constexpr int D = STAGED ? 1 : PATCH_DEPTH;
    static_assert(8 % D == 0, "the look-ahead ring is indexed by the row's place in the body");
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
    const int nr = nrows + TAPS - 1;                   // window rows of this patch
    if (!STAGED) {
#pragma unroll
        for (int r = 0; r < D; r++)
            if (r < nr)
                fetch(r, ahead[r]);
    }
#pragma unroll 1
    for (int q = 0; q < 2; q++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int r = 8 * q + j;
            if (r >= NR)                        // the second pass is shorter than the body
                break;
            // The row is read and filtered horizontally by every lane, with or without a use for it (rows past a short patch, a
            // lane walking without a task): no branch around it, so the rings need no copies where paths would join again.
            uint32_t wd[NW];
            if (STAGED) {
                // groups are committed in the order they are read: all but the NR - 1 youngest complete = this row has landed
                cp_wait<NR - 1>();
                if (MODE > 1)
                    __syncwarp();               // ... for every lane of the group (the warp walks its list in step)
#pragma unroll
                for (int i = 0; i < NW; i++)
                    wd[i] = lds32(MODE == 1 ? sm_rd + (r * NW + i) * 128 : sm_rd + r * mode_pitch(MODE) + 4 * i);
                if (MODE > 1)
                    __syncwarp();               // the whole group has read the row: it may be overwritten
                stage_row<TAPS, MODE>(sm_st, next, r);
                cp_commit();
            } else {
#pragma unroll
                for (int i = 0; i < NW; i++)
                    wd[i] = r < nr ? ahead[j % D][i] : 0u;
                if (r + D < nr)
                    fetch(r + D, ahead[j % D]);
            }
            {
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
                    P[(j + 7) & 7][c] = __byte_perm(prev[c], v, 0x5410);      // (row r - 1, row r) of column c; row 0's is never read
                    prev[c] = v;
                }
                if (r >= TAPS - 1 && r < nr) {
                    const int yo = (j + 9 - TAPS) & 7;                         // = r - (TAPS - 1): r < nr <= 8 + TAPS - 1
                    int o[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        if (TAPS == 8)
                            o[c] = __dp2a_lo((int)P[yo][c], (int)vf0, __dp2a_hi((int)P[(yo + 2) & 7][c], (int)vf0,
                                   __dp2a_lo((int)P[(yo + 4) & 7][c], (int)vf1, __dp2a_hi((int)P[(yo + 6) & 7][c], (int)vf1, 0)))) >> shv;
                        else
                            o[c] = __dp2a_lo((int)P[yo][c], (int)vf0, __dp2a_hi((int)P[(yo + 2) & 7][c], (int)vf0, 0)) >> shv;
                    }
                    if (BI && keep_pass) {
                        keep[2 * yo] = pack16(o[0], o[1]);
                        keep[2 * yo + 1] = pack16(o[2], o[3]);
                    } else {
                        uint2 res;
                        if (BI && gw) {
                            res = gpm_row(keep[2 * yo], keep[2 * yo + 1], o[0], o[1], o[2], o[3], gw + yo * gsy, gsx, cmax);
                        } else {
                            const int a[4] = { lo16(keep[2 * yo]), hi16(keep[2 * yo]), lo16(keep[2 * yo + 1]), hi16(keep[2 * yo + 1]) };
                            int out[4];
#pragma unroll
                            for (int c = 0; c < 4; c++) {
                                const int cur = (short)o[c];            // put() stores int16_t
                                const int v = BI ? a[c] * k.w0 + cur * k.w1 : cur * k.w0;
                                out[c] = d_clip_pel(((v + k.off) >> k.sh) + k.ox, 10);
                            }
                            res = make_uint2(pack16(out[0], out[1]), pack16(out[2], out[3]));
                        }
                        uint32_t *o32 = reinterpret_cast<uint32_t *>(d + (long long)yo * dpitch);
                        o32[0] = res.x;
                        if (two_words)
                            o32[1] = res.y;
                    }
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
// scan slots of the classifier -> ls.count[] entries: luma uni by width class, luma bi by width class, chroma uni / bi,
// DMVR-BDOF / PROF, luma border uni / bi, chroma border uni / bi
constexpr int kSlots = 14;
__constant__ int kCountOf[kSlots] = { 16, 17, 18, 20, 21, 22, 2, 3, 4, 5, 10, 11, 12, 13 };

__global__ void __launch_bounds__(256) inter_classify_kernel(const InterK p, const InterLists ls)
{
    // one reservation per list per CTA: block-wide exclusive scan of the per-record counts
    __shared__ uint32_t warp_tot[8][kSlots], cta_base[kSlots];
    const int ri = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int cls_l = -1, cls_c = -1, n_l = 0, n_c = 0, coop = -1, border = 0, wc = 0;
    if (ri < p.n) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p.pbs + ri);
        const uint32_t r0 = __ldg(q), r1 = __ldg(q + 1), r2 = __ldg(q + 2);
        const int w = r1 & 0xff, h = (r1 >> 8) & 0xff, planes = (r1 >> 16) & 0xff, pred = r1 >> 24, flags = r2 >> 24;
        wc = w >= 16 ? 2 : w >= 8 ? 1 : 0;
        border = w > 16;                      // not a record of the ABI (4, 8 or 16 wide): the unstaged path takes any patch grid
        if (flags & VVC_PB_COOPERATIVE) {
            coop = (flags & (VVC_CUDA_PB_PROF0 | VVC_CUDA_PB_PROF1)) ? 1 : 0;      // 0: DMVR / BDOF, 1: PROF
        } else {
            const int bi = (flags & VVC_CUDA_PB_GPM) || pred == 3;
            if (planes & VVC_CUDA_PB_LUMA) { cls_l = bi; n_l = (w >> 2) * ((h + 7) >> 3); }
            if ((planes & VVC_CUDA_PB_CHROMA) && p.planes == 3) { cls_c = bi; n_c = w > 8 ? 4 : 2; }   // 2 planes x patch columns
            // Records whose windows may leave the picture (a conservative test; mc_patch decides per patch) go to lists
            // of their own: their clamped loads cost ten times a plain patch, and one such lane stalls its whole warp.
            // Luma: 8 samples more on either side, the reach of the staged windows' aligned 8 / 16-byte requests.
            const int x0 = r0 & 0xffff, y0 = r0 >> 16;
            for (int l = 0; l < 2; l++) {
                if (!bi && l != pred - 1)
                    continue;
                const int mx = (int)__ldg(q + 3 + 2 * l), my = (int)__ldg(q + 4 + 2 * l);
                const int xx = x0 + (mx >> 4), yy = y0 + (my >> 4);
                const int m = p.margin, mc = p.margin >> 1;
                border |= xx - 12 < -m || xx + w + 13 > p.w + m || yy - 3 < -m || yy + h + 4 > p.h + m;
                const int xc = (x0 >> 1) + (mx >> 5), yc = (y0 >> 1) + (my >> 5);
                border |= xc - 2 < -mc || xc + (w >> 1) + 6 > (p.w >> 1) + mc || yc - 1 < -mc || yc + (h >> 1) + 2 > (p.h >> 1) + mc;
            }
        }
    }
    // this record's slots
    const int sl = cls_l < 0 ? -1 : border ? 10 + cls_l : cls_l * 3 + wc;
    const int sc = cls_c < 0 ? -1 : border ? 12 + cls_c : 6 + cls_c;
    const int sk = coop < 0 ? -1 : 8 + coop;
    uint32_t excl[kSlots];
#pragma unroll
    for (int c = 0; c < kSlots; c++) {
        const uint32_t mine = c == sl ? (uint32_t)n_l : c == sc ? (uint32_t)n_c : c == sk ? 1u : 0u;
        uint32_t v = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        excl[c] = v - mine;
        if (lane == 31) warp_tot[wid][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < kSlots) {
        uint32_t tot = 0;
        for (int k = 0; k < 8; k++) { const uint32_t t = warp_tot[k][threadIdx.x]; warp_tot[k][threadIdx.x] = tot; tot += t; }
        cta_base[threadIdx.x] = tot ? atomicAdd(ls.count + kCountOf[threadIdx.x], tot) : 0u;
    }
    __syncthreads();
    auto at = [&](int c) { uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < kSlots; k++) if (k == c) r = cta_base[k] + warp_tot[wid][k] + excl[k];
        return (int)r; };
    if (coop >= 0) {
        const int pos = at(sk);
        ls.coop[coop ? p.n - 1 - pos : pos] = ri;
    }
    if (cls_l >= 0) {
        // every record of a width class adds a multiple of its patches per row, so rows of patches stay aligned groups of lanes
        const int base = at(sl);
        uint32_t *list = border ? ls.luma_b : wc == 2 ? ls.luma[2] : wc == 1 ? ls.luma[1] : ls.luma[0];
        const int cap = border || wc == 2 ? ls.cap_luma[2] : wc == 1 ? ls.cap_luma[1] : ls.cap_luma[0];
        for (int k = 0; k < n_l; k++)
            list[cls_l ? cap - 1 - (base + k) : base + k] = ((uint32_t)ri << 3) | k;
    }
    if (cls_c >= 0) {
        const int base = at(sc);
        uint32_t *list = border ? ls.chroma_b : ls.chroma;
        for (int k = 0; k < n_c; k++)
            list[cls_c ? ls.cap_chroma - 1 - (base + k) : base + k] = ((uint32_t)ri << 3) | k;
    }
}

#define MV0(l, c) ((l) ? pb.mv[1][c] : pb.mv[0][c])
#define REF(l)    ((l) ? pb.ref[1] : pb.ref[0])

// ---- luma: patch pi of (w / 4) x ceil(h / 8) ---------------------------------------------------------------
// where a lane stages and reads: sm_st (its slot, or its chunk column of the group's region), the region itself, its group
struct Lane { uint32_t sm_st, region; int k; };

template <int TAPS, int MODE>
__device__ __forceinline__ Lane lane_of(uint32_t smem)
{
    constexpr int G = mode_lanes(MODE), NW = TAPS == 8 ? 6 : 4, NR = 8 + TAPS - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Lane l;
    l.k = lane & (G - 1);
    if (MODE == 1) {
        l.region = smem + warp * (32 * NR * NW * 4) + lane * 4;
        l.sm_st = l.region;
    } else {
        l.region = smem + warp * (32 * kSlotBytes) + (lane / G) * mode_region(MODE);
        l.sm_st = l.region + l.k * mode_chunk(MODE) * 2;
    }
    return l;
}

// window `l` of a luma task: MODE 1 the patch's own, MODE 2 / 3 the lane's share of the window of its row of patches
template <int MODE>
__device__ __forceinline__ Win luma_win(const InterK &p, const Rec &pb, int pi, int l, int k)
{
    const int w = pb.w, lw = 31 - __clz(w);
    const int ox = MODE == 1 ? (pi & ((w >> 2) - 1)) << 2 : 0, oy = (pi >> (lw - 2)) << 3, nrows = min(8, pb.h - oy);
    return win_of<8, MODE>(p.ref[0] + REF(l) * p.rb[0], p.rp[0], pb.x0 + ox + (MV0(l, 0) >> 4), pb.y0 + oy + (MV0(l, 1) >> 4), nrows, k);
}

// byte offset of the patch's first window word in its group's region
template <int TAPS, int MODE>
__device__ __forceinline__ uint32_t read_offset(int x_first, int x_patch)
{
    constexpr int B = TAPS / 2 - 1;
    return MODE == 1 ? 0u : (uint32_t)((((x_patch - B) & ~1) - ((x_first - B) & ~(mode_chunk(MODE) - 1))) * 2);
}

template <bool BI, int MODE = 0>
__device__ __forceinline__ void luma_task(const InterK &p, const Rec &pb, int pi, const Lane &ln = Lane(), const Win &next = Win(), bool live = true)
{
    const int w = pb.w, h = pb.h, lw = 31 - __clz(w);
    const bool gpm = BI && (pb.flags & VVC_CUDA_PB_GPM);
    const int lx = pb.pred - 1;
    const uint2 *lumaf = reinterpret_cast<const uint2 *>(&vvct_luma_mc_filters[0][0][0]);
    const uint8_t *wt = &vvct_gpm_weights[0][0] + pb.gw;
    // a lane without a task of its own walks a neighbour's with no rows: it only keeps the warp's waits and barriers in step
    const int ox = (pi & ((w >> 2) - 1)) << 2, oy = (pi >> (lw - 2)) << 3, nrows = live ? min(8, h - oy) : -7;
    const Blend k = blend_of(pb, p.wp, BI, lx, 0);
    pel *d = p.dst[0] + pb.pic * p.db[0] + (long long)(pb.y0 + oy) * p.dp[0] + pb.x0 + ox;
    const uint8_t *gw = gpm ? wt + oy * pb.gsy + ox * pb.gsx : nullptr;
    uint32_t keep[16];
#define LUMA_ARGS(l) p.ref[0] + REF(l) * p.rb[0], p.rp[0], p.w, p.h, p.margin, pb.x0 + ox + (MV0(l, 0) >> 4), pb.y0 + oy + (MV0(l, 1) >> 4), nrows, \
        (MV0(l, 0) & 15) ? fh.x : 0x01000000u, (MV0(l, 0) & 15) ? fh.y : 0u, \
        (MV0(l, 1) & 15) ? fv.x : ((MV0(l, 0) & 15) ? 0x01000000u : 0x10000000u), (MV0(l, 1) & 15) ? fv.y : 0u, \
        (MV0(l, 0) & 15) ? 2 : 0, (MV0(l, 1) & 15) ? ((MV0(l, 0) & 15) ? 6 : 2) : 0
#define LUMA_RD(l) ln.region + read_offset<8, MODE>(pb.x0 + (MV0(l, 0) >> 4), pb.x0 + ox + (MV0(l, 0) >> 4))
    const int filt = gpm ? 0 : pb.filt;
    if (BI) {
        Win second = Win();
        if (MODE)
            second = luma_win<MODE ? MODE : 1>(p, pb, pi, 1, ln.k);
        if (!live)
            second.nr = 0;
#pragma unroll 1
        for (int l = 0; l < 2; l++) {           // one body for both lists: the first one's rows are kept, the second one's blended with them
            const uint2 fh = lumaf[filt * 16 + (MV0(l, 0) & 15)], fv = lumaf[filt * 16 + (MV0(l, 1) & 15)];
            mc_patch<8, true, MODE>(LUMA_ARGS(l), l == 0, keep, k, gw, pb.gsx, pb.gsy, 3, d, p.dp[0], true, LUMA_RD(l), ln.sm_st, l == 0 ? second : next);
        }
    } else {
        const uint2 fh = lumaf[filt * 16 + (MV0(lx, 0) & 15)], fv = lumaf[filt * 16 + (MV0(lx, 1) & 15)];
        mc_patch<8, false, MODE>(LUMA_ARGS(lx), false, keep, k, nullptr, 0, 0, 3, d, p.dp[0], true, LUMA_RD(lx), ln.sm_st, next);
    }
#undef LUMA_ARGS
#undef LUMA_RD
}

// ---- chroma (4:2:0): task pi = (plane, patch column) ----------------------------------------------------------
// window `l` of a chroma task (private slots only)
__device__ __forceinline__ Win chroma_win(const InterK &p, const Rec &pb, int pi, int l)
{
    const int bw = pb.w >> 1, bh = pb.h >> 1, ncx = bw > 4 ? 2 : 1;
    const int pc = pi >> (ncx >> 1), ox = (pi & (ncx - 1)) << 2;
    return win_of<4, 1>((pc ? p.ref[2] : p.ref[1]) + REF(l) * (pc ? p.rb[2] : p.rb[1]), pc ? p.rp[2] : p.rp[1],
                        (pb.x0 >> 1) + ox + (MV0(l, 0) >> 5), (pb.y0 >> 1) + (MV0(l, 1) >> 5), bh, 0);
}

template <bool BI, int MODE = 0>
__device__ __forceinline__ void chroma_task(const InterK &p, const Rec &pb, int pi, const Lane &ln = Lane(), const Win &next = Win(), bool live = true)
{
    const bool gpm = BI && (pb.flags & VVC_CUDA_PB_GPM);
    const int lx = pb.pred - 1;
    const uint32_t *chromaf = reinterpret_cast<const uint32_t *>(&vvct_chroma_mc_filters[0][0][0]);
    const uint8_t *wt = &vvct_gpm_weights[0][0] + pb.gw;
    const int bw = pb.w >> 1, bh = live ? pb.h >> 1 : -3, ncx = bw > 4 ? 2 : 1;
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
        Win second = Win();
        if (MODE)
            second = chroma_win(p, pb, pi, 1);
        if (!live)
            second.nr = 0;
#pragma unroll 1
        for (int l = 0; l < 2; l++)
            mc_patch<4, true, MODE>(CHROMA_ARGS(l), l == 0, keep, k, gw, 2 * pb.gsx, 2 * pb.gsy, cmax, d, dpitch, bw > 2, ln.region, ln.sm_st, l == 0 ? second : next);
    } else {
        mc_patch<4, false, MODE>(CHROMA_ARGS(lx), false, keep, k, nullptr, 0, 0, cmax, d, dpitch, bw > 2, ln.region, ln.sm_st, next);
    }
#undef CHROMA_ARGS
}
#undef MV0
#undef REF

// One task list through the staged windows.  The warps walk the list with the launch's stride, all lanes of a warp in step
// (a lane past the end of the list walks its warp's first task without rows); the lanes of a group hold consecutive tasks
// of one row of patches (the lists are built that way).
template <bool LUMA, bool BI, int MODE>
__device__ __forceinline__ void run_list(const InterK &p, const uint32_t *list, int cap, int n, uint32_t smem)
{
    const int stride = gridDim.x * kThreads;
    int i = blockIdx.x * kThreads + threadIdx.x, i0 = i - (threadIdx.x & 31);
    if (i0 >= n)
        return;
    const Lane ln = lane_of<LUMA ? 8 : 4, MODE>(smem);
    auto task_at = [&](int j, int j0) { return __ldg(list + (BI ? cap - 1 - (j < n ? j : j0) : (j < n ? j : j0))); };
    auto first_win = [&](uint32_t t, bool live) {
        const Rec r = load_rec(p.pbs + (t >> 3));
        const int l = BI ? 0 : r.pred - 1;
        Win w = LUMA ? luma_win<MODE>(p, r, t & 7, l, ln.k) : chroma_win(p, r, t & 7, l);
        if (!live)
            w.nr = 0;
        return w;
    };
    uint32_t t = task_at(i, i0), tn = i0 + stride < n ? task_at(i + stride, i0 + stride) : 0u;
    stage_all<LUMA ? 8 : 4, MODE>(ln.sm_st, first_win(t, i < n));
    for (;;) {
        const int in = i + stride, in0 = i0 + stride;
        const uint32_t tnn = in0 + stride < n ? task_at(in + stride, in0 + stride) : 0u;       // the task after the next: its record is read one task from now
        Win next = Win();
        if (in0 < n)
            next = first_win(tn, in < n);
        const Rec pb = load_rec(p.pbs + (t >> 3));
        if (LUMA)
            luma_task<BI, MODE>(p, pb, t & 7, ln, next, i < n);
        else
            chroma_task<BI, MODE>(p, pb, t & 7, ln, next, i < n);
        if (in0 >= n)
            break;
        i = in; i0 = in0; t = tn; tn = tnn;
    }
}

// Persistent kernels, one per task class (own register budget each): a grid-stride loop over the class's
// lists, so every warp of the launch runs the same specialised code.
template <bool LUMA, bool BI>
__global__ void __launch_bounds__(kThreads, patch_ctas(LUMA, BI)) inter_patch_kernel(const InterK p, const InterLists ls)
{
    // the class's border tasks first (register-staged, clamped loads): they are contiguous in the index range, so the warps
    // that run the clamped path are full of such tasks instead of dragging plain patches through it, and they start first
    const int nb = (int)ls.count[10 + (LUMA ? 0 : 2) + (BI ? 1 : 0)];
    const uint32_t *blist = LUMA ? ls.luma_b : ls.chroma_b;
    const int bcap = LUMA ? ls.cap_luma[2] : ls.cap_chroma;
    const int stride = gridDim.x * kThreads;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < nb; i += stride) {
        const uint32_t t = __ldg(blist + (BI ? bcap - 1 - i : i));
        const Rec pb = load_rec(p.pbs + (t >> 3));
        if (LUMA)
            luma_task<BI>(p, pb, t & 7);
        else
            chroma_task<BI>(p, pb, t & 7);
    }
#if PATCH_STAGED
    // then the plain tasks through staged windows: the classifier sends every record with a window outside the plane (or its
    // margin, with room for the aligned requests of the group modes) to the border list, so these need no clamp
    __shared__ __align__(16) uint32_t slots[(LUMA ? kSlotBytes / 4 : 11 * 4) * kThreads];
    const uint32_t smem = (uint32_t)__cvta_generic_to_shared(slots);
    if (LUMA) {
        // the lists lay the warp's area out differently: the whole warp leaves one before any lane enters the next
        run_list<true, BI, 3>(p, ls.luma[2], ls.cap_luma[2], (int)ls.count[(BI ? 20 : 16) + 2], smem);
        __syncwarp();
        run_list<true, BI, 2>(p, ls.luma[1], ls.cap_luma[1], (int)ls.count[(BI ? 20 : 16) + 1], smem);
        __syncwarp();
        run_list<true, BI, 1>(p, ls.luma[0], ls.cap_luma[0], (int)ls.count[(BI ? 20 : 16) + 0], smem);
    } else {
        run_list<false, BI, 1>(p, ls.chroma, ls.cap_chroma, (int)ls.count[2 + (BI ? 1 : 0)], smem);
    }
#else
    for (int c = LUMA ? 2 : 0; c >= 0; c--) {
        const uint32_t *list = LUMA ? ls.luma[c] : ls.chroma;
        const int cap = LUMA ? ls.cap_luma[c] : ls.cap_chroma, n = (int)ls.count[LUMA ? (BI ? 20 : 16) + c : 2 + (BI ? 1 : 0)];
        for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
            const uint32_t t = __ldg(list + (BI ? cap - 1 - i : i));
            const Rec pb = load_rec(p.pbs + (t >> 3));
            if (LUMA)
                luma_task<BI>(p, pb, t & 7);
            else
                chroma_task<BI>(p, pb, t & 7);
        }
    }
#endif
}

}  // namespace

int vvc_inter_launch_classify(VVCCudaCtx *ctx, const InterK &p, InterLists *ls)
{
    // scratch: [count: 128 bytes][luma by width class: 2 n, 4 n, 8 n][chroma: 4 n][coop: n][luma border: 8 n][chroma border: 4 n] words,
    // then the refined vectors of the split DMVR kernels
    const size_t n = (size_t)p.n;
    uint32_t *base = (uint32_t *)vvc_ctx_scratch(ctx, 2, 128 + 31 * n * sizeof(uint32_t) + n * sizeof(VVCCudaDmvrOut));
    if (!base)
        return ctx->err;
    ls->count = base;
    ls->luma[0] = base + 32;             ls->cap_luma[0] = (int)(2 * n);
    ls->luma[1] = ls->luma[0] + 2 * n;   ls->cap_luma[1] = (int)(4 * n);
    ls->luma[2] = ls->luma[1] + 4 * n;   ls->cap_luma[2] = (int)(8 * n);
    ls->chroma = ls->luma[2] + 8 * n;    ls->cap_chroma = (int)(4 * n);
    ls->coop = ls->chroma + 4 * n;
    ls->luma_b = ls->coop + n;
    ls->chroma_b = ls->luma_b + 8 * n;
    ls->tail = ls->chroma_b + 4 * n;
    VVC_TRY(ctx, cudaMemsetAsync(ls->count, 0, 128, ctx->stream));
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
