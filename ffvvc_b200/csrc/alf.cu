// ALF stage for sm_100a: one fused launch per picture ring.
//
// Replaces the reference's per-CTB walk ff_vvc_alf_filter (libavcodec/vvc/vvc_filter.c:1254-1319)
// and the table entries it calls (libavcodec/vvc/vvc_filter_template.c):
//   alf_classify + alf_get_idx :270-381, alf_recon_coeff_and_clip :383-408,
//   alf_filter_luma :43-135, alf_filter_chroma :137-221, alf_filter_cc :223-263.
//
// B200 design: a CTA owns a TWxTH luma tile and the co-located 4:2:0 chroma tiles of both
// chroma planes.  It stages luma (+3 halo) and chroma (+2 halo) once into shared memory with
// 128-bit loads, derives the Laplacian lattice, per-4x4 class/transpose and the 12 (coeff, clip)
// pairs in shared memory, then filters luma 7x7, chroma 5x5 and adds the CC-ALF correction taken
// from the staged pre-ALF luma - each plane is read once and written once (6 B per luma pixel).
// Halo samples across a flagged CTB edge are a coordinate clamp, which is what the reference's
// alf_prepare_buffer (:1105-1137) materialises in its padded scratch copy.
#include "common.cuh"
#include "tables.cuh"

namespace {

struct AlfK {
    const pel *src[3];
    pel       *dst[3];
    int        sp[3], dp[3];          // row pitch in samples
    long long  sb[3], db[3];          // picture-to-picture stride in samples
    int        w, h, bd, ctb_log2, ctb_cols, ctb_rows, planes;
    const VVCCudaALFCtb  *ctbs;
    const VVCCudaALFSets *sets;
    int        sets_per_frame;
    int        wide_multiply;         // VVC_CUDA_OPT_ALF_WIDE_MULTIPLY: every block through the 32-bit-multiply path
};

constexpr int kThreads = 256;

template <int TW, int TH>
struct AlfSmem {
    static constexpr int LP = TW + 16;           // luma row pitch (samples), 8-sample aligned halo
    static constexpr int CP = TW / 2 + 16;
    alignas(16) pel     luma[TH + 6][LP];
    alignas(16) pel     chroma[2][TH / 2 + 4][CP];
    alignas(16) ushort4 cell[TH / 2 + 2][TW / 2 + 2];
    // per 4x4 block: 12 coefficients as IDP.2A words (f, 0, 0, f), 12 clip values in both halves, and one word of class
    // information (bit 0: some f == +128 -> wide path, bit 1: no clip can bind -> linear path, bits 8..: class, bits 16..:
    // transpose, bits 20..31: sum of the coefficients).  25 words
    // per block: an odd pitch, so the blocks a warp reads side by side sit in different banks.
    alignas(16) uint32_t coef[(TH / 4) * (TW / 4)][25];
};

#ifndef ALF_STAGE_ASYNC
#define ALF_STAGE_ASYNC 1           // tile staging by cp.async (tools/sweep_alf_stage.sh)
#endif
struct ClampWin { int xlo, xhi, ylo, yhi; };

__device__ __forceinline__ ClampWin make_win(int x0, int y0, int w, int h, int pw, int ph, unsigned edges)
{
    ClampWin c;
    c.xlo = (edges & VVC_CUDA_EDGE_LEFT)   ? x0         : 0;
    c.xhi = (edges & VVC_CUDA_EDGE_RIGHT)  ? x0 + w - 1 : pw - 1;
    c.ylo = (edges & VVC_CUDA_EDGE_TOP)    ? y0         : 0;
    c.yhi = (edges & VVC_CUDA_EDGE_BOTTOM) ? y0 + h - 1 : ph - 1;
    return c;
}

// Stage rows [ry0, ry0+rows) x columns [rx0, rx0 + 8*chunks) of a plane into shared memory,
// clamping coordinates to the CTB's window.
template <int PITCH>
__device__ __forceinline__ void stage_tile(pel *sm, const pel *__restrict__ plane, int pitch,
                                           int rx0, int ry0, int rows, int chunks, ClampWin cw)
{
    for (int idx = threadIdx.x; idx < rows * chunks; idx += kThreads) {
        const int i = idx / chunks, q = idx - i * chunks;
        const int y = d_clip3(ry0 + i, cw.ylo, cw.yhi);
        const int xs = rx0 + 8 * q;
        const pel *row = plane + (long long)y * pitch;
        pel *out = sm + i * PITCH + 8 * q;
        if (xs >= cw.xlo && xs + 7 <= cw.xhi) {
#if ALF_STAGE_ASYNC
            // cp.async: the chunk goes to shared memory without a destination register, so a thread's chunks of all three planes
            // are in flight together; one wait in front of the tile's barrier (the load -> store form paid a round trip per chunk)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(out)), "l"(row + xs) : "memory");
#else
            *reinterpret_cast<uint4 *>(out) = __ldg(reinterpret_cast<const uint4 *>(row + xs));
#endif
        } else {
#pragma unroll
            for (int e = 0; e < 8; e++)
                out[e] = __ldg(row + d_clip3(xs + e, cw.xlo, cw.xhi));
        }
    }
}

// vertical reach of k rows next to the virtual boundary (t = row - vb_pos), :80-96
__device__ __forceinline__ int vb_reach(int k, int t, int span)
{
    if (t < 0 && t >= -span) return min(k, -t - 1);
    if (t >= 0 && t < span)  return min(k, t);
    return k;
}

__constant__ uint8_t c_perm[4][12] = {
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 },
    { 9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6 },
    { 0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11 },
    { 9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6 },
};
__constant__ uint8_t c_act[16] = { 0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4 };
__constant__ uint8_t c_clip_shift[4] = { 0, 3, 5, 7 };

// Tap geometry of the two diamonds (:102-113, :196-201): vertical reach (index into d[], 0 = same row) and column offset
// of the first sample of each point-symmetric pair.
__constant__ int8_t c_luma_tap[12][2]   = { {3, 0}, {2, 1}, {2, 0}, {2, -1}, {1, 2}, {1, 1}, {1, 0}, {1, -1}, {1, -2}, {0, 3}, {0, 2}, {0, 1} };
__constant__ int8_t c_chroma_tap[6][2]  = { {2, 0}, {1, 1}, {1, 0}, {1, -1}, {0, 2}, {0, 1} };

// One 4x1 strip with 32-bit multiplies: the path of coefficient sets that hold +128, which does not fit the signed byte
// of the IDP.2A form (and the comparison path of VVC_CUDA_OPT_ALF_WIDE_MULTIPLY).  Rare, so written for few registers
// (rolled loops, coefficients fetched per tap) and kept out of line: the packed path's registers and code are unaffected.
// luma: info = the block's class word (sm.coef[..][24]), alt = -1; chroma: alt = the CTB's alternative filter.
__device__ __noinline__ uint2 alf_strip_wide(const pel *p0, int d1, int d2, int d3, bool near_vb, int bd,
                                             const VVCCudaALFSets *sets, int set, uint32_t info, int alt)
{
    const bool chroma = alt >= 0;
    const int cls = (info >> 8) & 0xff, tr = (info >> 16) & 3;
    const int16_t *ff;
    const uint8_t *ci = nullptr;
    if (chroma)        { ff = sets->chroma_coeff[alt]; ci = sets->chroma_clip_idx[alt]; }
    else if (set < 16) { ff = vvct_alf_fix_filt_coeff[vvct_alf_class_to_filt_map[set][cls]]; }
    else               { ff = sets->luma_coeff[set - 16][vvct_alf_aps_class_to_filt_map[cls]]; ci = sets->luma_clip_idx[set - 16][cls]; }
    const int d[4] = { 0, d1, d2, d3 };
    int sum[4] = { 0, 0, 0, 0 };
#pragma unroll 1
    for (int j = 0; j < (chroma ? 6 : 12); j++) {
        const int s = chroma ? j : c_perm[tr][j];
        const int f = ff[s], c = 1 << (bd - (ci ? c_clip_shift[ci[s]] : 0));
        const int dy = d[chroma ? c_chroma_tap[j][0] : c_luma_tap[j][0]], dx = chroma ? c_chroma_tap[j][1] : c_luma_tap[j][1];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int cur = p0[i];
            sum[i] += f * (d_clip3((int)p0[i + dy + dx] - cur, -c, c) + d_clip3((int)p0[i - dy - dx] - cur, -c, c));
        }
    }
    unsigned res[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        res[i] = (unsigned)d_clip_pel(p0[i] + (near_vb ? (sum[i] + 512) >> 10 : (sum[i] + 64) >> 7), bd);
    return make_uint2(res[0] | (res[1] << 16), res[2] | (res[3] << 16));
}

#ifndef ALF_TH
#define ALF_TH 32                    // tile height of the 64-wide tiles
#endif
#ifndef ALF_MIN_CTAS
#define ALF_MIN_CTAS 0               // resident CTAs per SM the kernel is compiled for (0: the compiler's choice; tools/sweep_alf_occ.sh)
#endif
template <int TW, int TH>
__global__ void
#if ALF_MIN_CTAS > 0
__launch_bounds__(kThreads, ALF_MIN_CTAS)
#else
__launch_bounds__(kThreads)
#endif
alf_frame_kernel(const AlfK p)
{
    using SM = AlfSmem<TW, TH>;
    __shared__ SM sm;
    constexpr int LP = SM::LP, CP = SM::CP;
    constexpr int CW = TW / 2, CH = TH / 2;
    constexpr int BX = TW / 4, BY = TH / 4;

    const int tid = threadIdx.x;
    const int k = blockIdx.z;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int ctb = 1 << p.ctb_log2;
    const int cx = tx0 >> p.ctb_log2, cy = ty0 >> p.ctb_log2;
    VVCCudaALFCtb a;                     // 12 bytes of byte fields: three word loads where the table is word aligned
    {
        const VVCCudaALFCtb *ap = p.ctbs + ((long long)k * p.ctb_rows + cy) * p.ctb_cols + cx;
        static_assert(sizeof(VVCCudaALFCtb) == 12, "VVCCudaALFCtb is three words");
        if (((uintptr_t)p.ctbs & 3) == 0) {
            const uint32_t *aw = reinterpret_cast<const uint32_t *>(ap);
            const uint32_t wv[3] = { __ldg(aw), __ldg(aw + 1), __ldg(aw + 2) };
            memcpy(&a, wv, 12);
        } else {
            a = *ap;
        }
    }
    const VVCCudaALFSets *sets = p.sets + (p.sets_per_frame ? k : 0);
    const int bd = p.bd;

    unsigned edges = a.edges;
    if (cx == 0)              edges |= VVC_CUDA_EDGE_LEFT;
    if (cy == 0)              edges |= VVC_CUDA_EDGE_TOP;
    if (cx == p.ctb_cols - 1) edges |= VVC_CUDA_EDGE_RIGHT;
    if (cy == p.ctb_rows - 1) edges |= VVC_CUDA_EDGE_BOTTOM;
    const int x0 = cx << p.ctb_log2, y0 = cy << p.ctb_log2;
    const int lw = min(ctb, p.w - x0), lh = min(ctb, p.h - y0);

    // ---- stage the tiles ------------------------------------------------------------------
    {
        const ClampWin cw = make_win(x0, y0, lw, lh, p.w, p.h, edges);
        stage_tile<LP>(&sm.luma[0][0], p.src[0] + k * p.sb[0], p.sp[0], tx0 - 8, ty0 - 3, TH + 6, LP / 8, cw);
    }
    if (p.planes == 3) {
        const ClampWin cw = make_win(x0 >> 1, y0 >> 1, lw >> 1, lh >> 1, p.w >> 1, p.h >> 1, edges);
#pragma unroll
        for (int c = 0; c < 2; c++)
            stage_tile<CP>(&sm.chroma[c][0][0], p.src[c + 1] + k * p.sb[c + 1], p.sp[c + 1],
                           (tx0 >> 1) - 8, (ty0 >> 1) - 2, CH + 4, CP / 8, cw);
    }
#if ALF_STAGE_ASYNC
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
    __syncthreads();

    const int vb = ctb - 4;               // luma virtual boundary, CTB relative (:1304)
    const int trow = ty0 - y0;            // tile row inside the CTB

    if (a.ctb_flag[0]) {
        // ---- Laplacian lattice, one cell per 2x2 luma positions (:315-344) -----------------
        for (int idx = tid; idx < (TH / 2 + 2) * (TW / 2 + 2); idx += kThreads) {
            const int ci = idx / (TW / 2 + 2), cj = idx - ci * (TW / 2 + 2);
            const int yy = trow + 2 * ci;                 // the reference's loop variable y
            int r0 = 2 * ci, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3;   // shared-memory rows
            if (yy == vb)          r3 = r2;
            else if (yy == vb + 2) r0 = r1;
            const int ca = 2 * cj + 6;                    // column of the first sample point
            const pel *R0 = sm.luma[r0], *R1 = sm.luma[r1], *R2 = sm.luma[r2], *R3 = sm.luma[r3];
            const int c0 = 2 * R1[ca], c1 = 2 * R2[ca + 1];
            ushort4 g;
            g.x = abs(c0 - R0[ca] - R2[ca])         + abs(c1 - R1[ca + 1] - R3[ca + 1]);   // vertical
            g.y = abs(c0 - R1[ca - 1] - R1[ca + 1]) + abs(c1 - R2[ca] - R2[ca + 2]);       // horizontal
            g.z = abs(c0 - R0[ca - 1] - R2[ca + 1]) + abs(c1 - R1[ca] - R3[ca + 2]);       // diagonal 0
            g.w = abs(c0 - R0[ca + 1] - R2[ca - 1]) + abs(c1 - R1[ca + 2] - R3[ca]);       // diagonal 1
            sm.cell[ci][cj] = g;
        }
        __syncthreads();

        // ---- class / transpose per 4x4 block and its 12 (coeff, clip) pairs (:270-297,:346-408)
        for (int b = tid; b < BX * BY; b += kThreads) {
            const int bi = b / BX, bj = b - bi * BX;
            const int by = trow + 4 * bi;
            int first = 0, last = 4, scale = 2;
            if (by + 4 == vb)  { last = 3;  scale = 3; }
            else if (by == vb) { first = 1; scale = 3; }
            int sv = 0, sh = 0, sd0 = 0, sd1 = 0;
            for (int i = first; i < last; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const ushort4 g = sm.cell[2 * bi + i][2 * bj + j];
                    sv += g.x; sh += g.y; sd0 += g.z; sd1 += g.w;
                }
            const int v_le_h = sv <= sh, d0_le_d1 = sd0 <= sd1;
            const int hv_hi = max(sv, sh), hv_lo = min(sv, sh);
            const int d_hi = max(sd0, sd1), d_lo = min(sd0, sd1);
            const int hv_wins = (unsigned long long)d_hi * (unsigned)hv_lo <= (unsigned long long)hv_hi * (unsigned)d_lo;
            const int hi = hv_wins ? hv_hi : d_hi, lo = hv_wins ? hv_lo : d_lo;
            int cls = c_act[d_clip_ubits(((sh + sv) * scale) >> (bd - 1), 4)];
            if (hi * 2 > 9 * lo)  cls += (2 * hv_wins + 2) * 5;
            else if (hi > 2 * lo) cls += (2 * hv_wins + 1) * 5;
            const int tr = d0_le_d1 * 2 + v_le_h;

            // Coefficients are -128..128 by syntax (alf_luma_coeff_abs 0..128, cbs_h266_syntax_template.c:2285, stored as
            // +-abs in int16, vvc_ps.c:803-808).  -128..127 go into signed bytes: byte 0 feeds the first sample of a pair
            // (IDP.2A.LO), byte 3 the second (IDP.2A.HI).  A block whose permuted set holds +128 is flagged and filtered by
            // alf_strip_wide() with 32-bit multiplies.
            uint32_t *out = sm.coef[b];
            const int set = a.filt_set_idx_y;
            int wide = 0, clips = 0, fsum = 0;        // clips: some clip index other than 0 (a clip of 1 << bd never binds)
            if (set < 16) {
                const int16_t *f = vvct_alf_fix_filt_coeff[vvct_alf_class_to_filt_map[set][cls]];
#pragma unroll
                for (int j = 0; j < 12; j++) {
                    const uint32_t fb = (uint32_t)f[c_perm[tr][j]] & 0xff, cv = 1u << bd;
                    fsum += f[c_perm[tr][j]];
                    out[j]      = fb | (fb << 24);
                    out[12 + j] = cv | (cv << 16);
                }
            } else {
                const int16_t *f  = sets->luma_coeff[set - 16][vvct_alf_aps_class_to_filt_map[cls]];
                const uint8_t *ci = sets->luma_clip_idx[set - 16][cls];   // indexed by class (:400)
#pragma unroll
                for (int j = 0; j < 12; j++) {
                    const int s = c_perm[tr][j];
                    const uint32_t fb = (uint32_t)f[s] & 0xff, cv = 1u << (bd - c_clip_shift[ci[s]]);
                    wide |= f[s] > 127;
                    clips |= ci[s];
                    fsum += f[s];
                    out[j]      = fb | (fb << 24);
                    out[12 + j] = cv | (cv << 16);
                }
            }
            out[24] = (uint32_t)(wide | p.wide_multiply) | (clips ? 0u : 2u) | ((uint32_t)cls << 8) | ((uint32_t)tr << 16) | ((uint32_t)fsum << 20);
        }
        __syncthreads();
    }

    // ---- luma 7x7 diamond (:43-135), 4x1 strips ------------------------------------------------
    {
        pel *dplane = p.dst[0] + k * p.db[0];
        bool any_wide = false;
        for (int s = tid; s < BX * TH; s += kThreads) {
            const int r = s / BX, c4 = s - r * BX;
            const int y = ty0 + r, x = tx0 + 4 * c4;
            if (y >= p.h || x >= p.w)
                continue;
            const pel *p0 = &sm.luma[r + 3][8 + 4 * c4];
            uint2 o;
            if (!a.ctb_flag[0]) {
                o = *reinterpret_cast<const uint2 *>(p0);
            } else {
                const int t = (y - y0) - vb;
                const int d1 = vb_reach(1, t, 4) * LP, d2 = vb_reach(2, t, 4) * LP, d3 = vb_reach(3, t, 4) * LP;
                const bool near_vb = (t == -1 || t == 0);
                const uint32_t *cf = sm.coef[(r >> 2) * BX + c4];
                if (cf[24] & 1) {           // +128 in this block's set: left to the wide-multiply loop below
                    any_wide = true;
                    continue;
                }
                // The strip as two sample pairs in 16x2 arithmetic: clip3(-c, c, n - cur) = max(min(n + (-cur), c), -c) is
                // VIADDMNMX + VIMNMX per pair, the two clipped differences add as VIADD.16x2, and IDP.2A multiplies the
                // pair by the tap (exact 32-bit accumulate per sample).  wd[i] = samples (-4 + 2i, -3 + 2i) of a row.
                const uint2 c01 = *reinterpret_cast<const uint2 *>(p0);
                const uint32_t ncur[2] = { __vneg2(c01.x), __vneg2(c01.y) };
                int sum[4] = { 0, 0, 0, 0 };
                const bool linear = cf[24] & 2;
#define ALF_ROW(w, row, lo, hi)                                                        \
                {                                                                      \
                    const pel *rp_ = (row);                                            \
                    const uint2 m_ = *reinterpret_cast<const uint2 *>(rp_);            \
                    (w)[2] = m_.x; (w)[3] = m_.y;                                      \
                    if (lo) { if ((lo) > 1) { const uint2 l_ = *reinterpret_cast<const uint2 *>(rp_ - 4); (w)[0] = l_.x; (w)[1] = l_.y; } \
                              else (w)[1] = *reinterpret_cast<const uint32_t *>(rp_ - 2); }                                            \
                    if (hi) { if ((hi) > 1) { const uint2 h_ = *reinterpret_cast<const uint2 *>(rp_ + 4); (w)[4] = h_.x; (w)[5] = h_.y; } \
                              else (w)[4] = *reinterpret_cast<const uint32_t *>(rp_ + 4); }                                            \
                }
#define ALF_AT(w, o) ((((o) + 4) & 1) ? __funnelshift_r((w)[((o) + 4) >> 1], (w)[(((o) + 4) >> 1) + 1], 16) : (w)[((o) + 4) >> 1])
#define ALF_TAP(k, wp, wm, off) ALF_TAPW(cf[k], cf[12 + (k)], wp, wm, off)
#define ALF_TAPW(fword, cword, wp, wm, off)                                            \
                {                                                                      \
                    const uint32_t fw_ = (fword), c2_ = (cword), nc2_ = (~c2_) + 0x00010001u;                                          \
                    ALF_PAIR(0, wp, wm, off) ALF_PAIR(1, wp, wm, off)                  \
                }
// the linear form: with every clip at 1 << bd no difference is ever clipped, so sum f (n1 - cur + n2 - cur) =
// sum f (n1 + n2) - 2 cur sum f; one packed add and the two IDP.2A per tap and sample pair
#define ALF_TAPL(k, wp, wm, off)                                                       \
                {                                                                      \
                    const uint32_t fw_ = cf[k];                                        \
                    ALF_PAIRL(0, wp, wm, off) ALF_PAIRL(1, wp, wm, off)                \
                }
#define ALF_PAIRL(j_, wp, wm, off)                                                     \
                    {                                                                  \
                        const uint32_t s_ = ALF_AT(wp, 2 * (j_) + (off)) + ALF_AT(wm, 2 * (j_) - (off));                               \
                        sum[2 * (j_)] = __dp2a_lo((int)s_, (int)fw_, sum[2 * (j_)]);  \
                        sum[2 * (j_) + 1] = __dp2a_hi((int)s_, (int)fw_, sum[2 * (j_) + 1]);                                          \
                    }
#define ALF_PAIR(j_, wp, wm, off)                                                      \
                    {                                                                  \
                        const uint32_t a_ = __vmaxs2(__viaddmin_s16x2(ALF_AT(wp, 2 * (j_) + (off)), ncur[j_], c2_), nc2_);             \
                        const uint32_t b_ = __vmaxs2(__viaddmin_s16x2(ALF_AT(wm, 2 * (j_) - (off)), ncur[j_], c2_), nc2_);             \
                        const uint32_t s_ = __vadd2(a_, b_);                           \
                        sum[2 * (j_)] = __dp2a_lo((int)s_, (int)fw_, sum[2 * (j_)]);  \
                        sum[2 * (j_) + 1] = __dp2a_hi((int)s_, (int)fw_, sum[2 * (j_) + 1]);                                          \
                    }
                const int cur[4] = { (int)(c01.x & 0xffff), (int)(c01.x >> 16), (int)(c01.y & 0xffff), (int)(c01.y >> 16) };
                if (linear) {
                    {
                        uint32_t w0[6];
                        ALF_ROW(w0, p0, 2, 2)
                        ALF_TAPL(9, w0, w0, 3) ALF_TAPL(10, w0, w0, 2) ALF_TAPL(11, w0, w0, 1)
                    }
                    {
                        uint32_t wp[6], wm[6];
                        ALF_ROW(wp, p0 + d1, 1, 1) ALF_ROW(wm, p0 - d1, 1, 1)
                        ALF_TAPL(4, wp, wm, 2) ALF_TAPL(5, wp, wm, 1) ALF_TAPL(6, wp, wm, 0) ALF_TAPL(7, wp, wm, -1) ALF_TAPL(8, wp, wm, -2)
                    }
                    {
                        uint32_t wp[6], wm[6];
                        ALF_ROW(wp, p0 + d2, 1, 1) ALF_ROW(wm, p0 - d2, 1, 1)
                        ALF_TAPL(1, wp, wm, 1) ALF_TAPL(2, wp, wm, 0) ALF_TAPL(3, wp, wm, -1)
                    }
                    {
                        uint32_t wp[6], wm[6];
                        ALF_ROW(wp, p0 + d3, 0, 0) ALF_ROW(wm, p0 - d3, 0, 0)
                        ALF_TAPL(0, wp, wm, 0)
                    }
                    const int fsum2 = 2 * ((int)cf[24] >> 20);
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        sum[j] -= cur[j] * fsum2;
                } else {
                {
                    uint32_t w0[6];
                    ALF_ROW(w0, p0, 2, 2)
                    ALF_TAP(9, w0, w0, 3) ALF_TAP(10, w0, w0, 2) ALF_TAP(11, w0, w0, 1)
                }
                {
                    uint32_t wp[6], wm[6];
                    ALF_ROW(wp, p0 + d1, 1, 1) ALF_ROW(wm, p0 - d1, 1, 1)
                    ALF_TAP(4, wp, wm, 2) ALF_TAP(5, wp, wm, 1) ALF_TAP(6, wp, wm, 0) ALF_TAP(7, wp, wm, -1) ALF_TAP(8, wp, wm, -2)
                }
                {
                    uint32_t wp[6], wm[6];
                    ALF_ROW(wp, p0 + d2, 1, 1) ALF_ROW(wm, p0 - d2, 1, 1)
                    ALF_TAP(1, wp, wm, 1) ALF_TAP(2, wp, wm, 0) ALF_TAP(3, wp, wm, -1)
                }
                {
                    uint32_t wp[6], wm[6];
                    ALF_ROW(wp, p0 + d3, 0, 0) ALF_ROW(wm, p0 - d3, 0, 0)
                    ALF_TAP(0, wp, wm, 0)
                }
                }
                unsigned res[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int v = near_vb ? (sum[j] + 512) >> 10 : (sum[j] + 64) >> 7;
                    res[j] = (unsigned)d_clip_pel(cur[j] + v, bd);
                }
                o.x = res[0] | (res[1] << 16);
                o.y = res[2] | (res[3] << 16);
            }
            *reinterpret_cast<uint2 *>(dplane + (long long)y * p.dp[0] + x) = o;
        }
        // strips of blocks whose coefficient set holds +128 (rare; no thread of most tiles ever gets here)
        if (any_wide) {
            for (int s = tid; s < BX * TH; s += kThreads) {
                const int r = s / BX, c4 = s - r * BX;
                const int y = ty0 + r, x = tx0 + 4 * c4;
                const uint32_t info = sm.coef[(r >> 2) * BX + c4][24];
                if (y >= p.h || x >= p.w || !(info & 1))
                    continue;
                const int t = (y - y0) - vb;
                *reinterpret_cast<uint2 *>(dplane + (long long)y * p.dp[0] + x) =
                    alf_strip_wide(&sm.luma[r + 3][8 + 4 * c4], vb_reach(1, t, 4) * LP, vb_reach(2, t, 4) * LP, vb_reach(3, t, 4) * LP,
                                            t == -1 || t == 0, bd, sets, a.filt_set_idx_y, info, -1);
            }
        }
    }

    // ---- chroma 5x5 diamond (:137-221) + CC-ALF from the staged pre-ALF luma (:223-263) --------
    if (p.planes == 3) {
        const int cvb = (ctb >> 1) - 2;
        const int pw = p.w >> 1, ph = p.h >> 1;
        constexpr int SX = CW / 4;                      // strips per chroma row
        for (int s = tid; s < 2 * SX * CH; s += kThreads) {
            const int pc = s / (SX * CH);               // 0 = Cb, 1 = Cr
            const int rem = s - pc * (SX * CH);
            const int r = rem / SX, c4 = rem - r * SX;
            const int y = (ty0 >> 1) + r, x = (tx0 >> 1) + 4 * c4;
            if (y >= ph || x >= pw)
                continue;
            const pel *p0 = &sm.chroma[pc][r + 2][8 + 4 * c4];
            int val[4];
            if (!a.ctb_flag[pc + 1]) {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    val[j] = p0[j];
            } else {
                const int t = (y - (y0 >> 1)) - cvb;
                const int d1 = vb_reach(1, t, 2) * CP, d2 = vb_reach(2, t, 2) * CP;
                const bool near_vb = (t == -1 || t == 0);
                const int alt = a.chroma_alt_idx[pc];
                bool wide = p.wide_multiply;
#pragma unroll
                for (int j = 0; j < 6; j++)
                    wide |= sets->chroma_coeff[alt][j] > 127;
                if (wide) {
                    // 32-bit multiplies: sets holding +128 (alf_chroma_coeff_abs 0..128, cbs_h266_syntax_template.c:2314)
                    const uint2 o = alf_strip_wide(p0, d1, d2, 0, near_vb, bd, sets, 0, 0u, alt);
                    val[0] = o.x & 0xffff; val[1] = o.x >> 16; val[2] = o.y & 0xffff; val[3] = o.y >> 16;
                } else {
                    // same 16x2 scheme as luma: one coefficient / clip word per tap for the whole CTB
                    uint32_t fw[6], cw[6];
    #pragma unroll
                    for (int j = 0; j < 6; j++) {
                        const uint32_t fb = (uint32_t)sets->chroma_coeff[alt][j] & 0xff, cv = 1u << (bd - c_clip_shift[sets->chroma_clip_idx[alt][j]]);
                        fw[j] = fb | (fb << 24);
                        cw[j] = cv | (cv << 16);
                    }
                    const uint2 c01 = *reinterpret_cast<const uint2 *>(p0);
                    const uint32_t ncur[2] = { __vneg2(c01.x), __vneg2(c01.y) };
                    int sum[4] = { 0, 0, 0, 0 };
                    {
                        uint32_t w0[6];
                        ALF_ROW(w0, p0, 1, 1)
                        ALF_TAPW(fw[4], cw[4], w0, w0, 2) ALF_TAPW(fw[5], cw[5], w0, w0, 1)
                    }
                    {
                        uint32_t wp[6], wm[6];
                        ALF_ROW(wp, p0 + d1, 1, 1) ALF_ROW(wm, p0 - d1, 1, 1)
                        ALF_TAPW(fw[1], cw[1], wp, wm, 1) ALF_TAPW(fw[2], cw[2], wp, wm, 0) ALF_TAPW(fw[3], cw[3], wp, wm, -1)
                    }
                    {
                        uint32_t wp[6], wm[6];
                        ALF_ROW(wp, p0 + d2, 0, 0) ALF_ROW(wm, p0 - d2, 0, 0)
                        ALF_TAPW(fw[0], cw[0], wp, wm, 0)
                    }
                    const int cur[4] = { (int)(c01.x & 0xffff), (int)(c01.x >> 16), (int)(c01.y & 0xffff), (int)(c01.y >> 16) };
    #pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int v = near_vb ? (sum[j] + 512) >> 10 : (sum[j] + 64) >> 7;
                        val[j] = d_clip_pel(cur[j] + v, bd);
                    }
                }
            }
            if (a.cc_idc[pc]) {
                const int t = (((y - (y0 >> 1)) << 1)) - vb;     // luma row relative to the VB
                int up = -LP, dn = LP, dn2 = 2 * LP;
                if (t == -2 || t == 1)       dn2 = LP;
                else if (t == -1 || t == 0)  up = dn = dn2 = 0;
                const int16_t *f = sets->cc_coeff[pc][a.cc_idc[pc] - 1];
                const pel *l0 = &sm.luma[2 * r + 3][8 + 8 * c4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const pel *l = l0 + 2 * j;
                    const int cur = l[0];
                    int sum = 0;
                    sum += f[0] * (l[up]      - cur);
                    sum += f[1] * (l[-1]      - cur);
                    sum += f[2] * (l[1]       - cur);
                    sum += f[3] * (l[dn - 1]  - cur);
                    sum += f[4] * (l[dn]      - cur);
                    sum += f[5] * (l[dn + 1]  - cur);
                    sum += f[6] * (l[dn2]     - cur);
                    sum = d_clip3((sum + 64) >> 7, -(1 << (bd - 1)), (1 << (bd - 1)) - 1);
                    val[j] = d_clip_pel(val[j] + sum, bd);
                }
            }
            uint2 o;
            o.x = (unsigned)val[0] | ((unsigned)val[1] << 16);
            o.y = (unsigned)val[2] | ((unsigned)val[3] << 16);
            *reinterpret_cast<uint2 *>(p.dst[pc + 1] + k * p.db[pc + 1] + (long long)y * p.dp[pc + 1] + x) = o;
        }
    }
}

#undef ALF_TAP
#undef ALF_TAPL
#undef ALF_PAIRL
#undef ALF_TAPW
#undef ALF_PAIR
#undef ALF_AT
#undef ALF_ROW

int check_frames(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src)
{
    if (!dst || !src)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "alf: null frame");
    if (src->bit_depth != 10 && src->bit_depth != 12)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "alf: bit depth %d not accelerated (10/12 only)", src->bit_depth);
    if (src->ctb_log2 < 5 || src->ctb_log2 > 7 || src->batch < 1 || src->width < 8 || src->height < 8 ||
        (src->width & 7) || (src->height & 7))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "alf: unsupported geometry %dx%d ctb %d", src->width, src->height, 1 << src->ctb_log2);
    if (src->chroma_format_idc && (src->hshift != 1 || src->vshift != 1))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "alf: only 4:0:0 and 4:2:0 are accelerated");
    if (dst->width != src->width || dst->height != src->height || dst->batch != src->batch ||
        dst->bit_depth != src->bit_depth || dst->chroma_format_idc != src->chroma_format_idc)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "alf: dst/src geometry differs");
    if (!frame_vec_ok(dst) || !frame_vec_ok(src))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "alf: planes and strides must be 16-byte aligned");
    return 0;
}

}  // namespace

extern "C" int vvc_cuda_alf_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                                  const VVCCudaALFCtb *ctbs, const VVCCudaALFSets *sets, int sets_per_frame)
{
    if (ctx->err)
        return ctx->err;
    if (check_frames(ctx, dst, src))
        return ctx->err;
    AlfK p;
    p.planes = src->chroma_format_idc ? 3 : 1;
    for (int c = 0; c < 3; c++) {
        p.src[c] = (const pel *)src->data[c];
        p.dst[c] = (pel *)dst->data[c];
        p.sp[c] = (int)(src->stride[c] / 2);  p.dp[c] = (int)(dst->stride[c] / 2);
        p.sb[c] = src->batch_stride[c] / 2;   p.db[c] = dst->batch_stride[c] / 2;
    }
    p.w = src->width; p.h = src->height; p.bd = src->bit_depth; p.ctb_log2 = src->ctb_log2;
    p.ctb_cols = ceil_div(p.w, 1 << p.ctb_log2);
    p.ctb_rows = ceil_div(p.h, 1 << p.ctb_log2);
    p.ctbs = ctbs; p.sets = sets; p.sets_per_frame = sets_per_frame;
    p.wide_multiply = ctx->alf_wide_multiply;

    if (p.ctb_log2 >= 6) {
        dim3 grid(ceil_div(p.w, 64), ceil_div(p.h, ALF_TH), src->batch);
        alf_frame_kernel<64, ALF_TH><<<grid, kThreads, 0, ctx->stream>>>(p);
    } else {
        dim3 grid(ceil_div(p.w, 32), ceil_div(p.h, 32), src->batch);
        alf_frame_kernel<32, 32><<<grid, kThreads, 0, ctx->stream>>>(p);
    }
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_alf_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                                       const VVCCudaALFCtb *ctbs, const VVCCudaALFSets *sets, int sets_per_frame)
{
    if (ctx->err)
        return ctx->err;
    if (!dst || !src || !ctbs || !sets)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "alf_host: null argument");
    const int n_ctb = ceil_div(src->width, 1 << src->ctb_log2) * ceil_div(src->height, 1 << src->ctb_log2) * src->batch;
    const size_t fsz = align_up(vvc_stage_frame_size(src), 256);
    const size_t csz = align_up((size_t)n_ctb * sizeof(VVCCudaALFCtb), 256);
    const size_t ssz = align_up(sizeof(VVCCudaALFSets) * (sets_per_frame ? src->batch : 1), 256);
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, 2 * fsz + csz + ssz);
    if (!base)
        return ctx->err;
    VVCCudaFrame dsrc, ddst;
    vvc_stage_frame_layout(src, base, &dsrc);
    vvc_stage_frame_layout(dst, base + fsz, &ddst);
    VVCCudaALFCtb  *dctb  = (VVCCudaALFCtb *)(base + 2 * fsz);
    VVCCudaALFSets *dsets = (VVCCudaALFSets *)(base + 2 * fsz + csz);
    if (vvc_stage_frame_h2d(ctx, &dsrc, src))
        return ctx->err;
    VVC_TRY(ctx, cudaMemcpyAsync(dctb, ctbs, (size_t)n_ctb * sizeof(VVCCudaALFCtb), cudaMemcpyHostToDevice, ctx->stream));
    VVC_TRY(ctx, cudaMemcpyAsync(dsets, sets, sizeof(VVCCudaALFSets) * (sets_per_frame ? src->batch : 1), cudaMemcpyHostToDevice, ctx->stream));
    if (vvc_cuda_alf_frame(ctx, &ddst, &dsrc, dctb, dsets, sets_per_frame))
        return ctx->err;
    if (vvc_stage_frame_d2h(ctx, dst, &ddst))
        return ctx->err;
    return vvc_cuda_sync(ctx);
}
