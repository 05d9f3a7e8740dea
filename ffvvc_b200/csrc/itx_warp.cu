// Residual stage, warp-per-TB kernel for 10-bit pictures with 16-bit coefficients (log2_transform_range 15).
//
// Same contract as itx.cu (the generic kernel) for the common kinds of transform block: 2-D blocks 2..64
// per side, DCT-II / DST-VII / DCT-VIII, optional LFNST, add_residual / add_residual_joint or the residual
// written back.  Reference: itx_2d + scale_clip (libavcodec/vvc/vvcdsp.c:67-117), the 1-D kernels
// (libavcodec/vvc/vvc_itx_1d.c:70-706) as exact matrix products, ilfnst_transform (vvc_intra.c:65-127),
// add_residual(_joint) (vvcdsp_template.c:32-63).  Transform skip, BDPCM and 1-D blocks stay in itx.cu.
//
// B200 design: both passes are int16 x int8 products with exact int32 accumulation, i.e. IDP.2A on pairs
// along the reduction dimension (coefficients are clipped to 16 bits by dequant, the mid-stage clip keeps
// the second pass in 16 bits too).
//  * Pass 1 (vertical) is input-stationary: a lane owns one coefficient column of the non-zero window, packed
//    in registers straight from HBM (rows are read coalesced across lanes), and sweeps the outputs with the
//    matrix fetched as 128-bit words of the packed table (4 words = 16 taps of one output).
//  * Pass 2 (horizontal) is matrix-stationary: a lane owns one (two for 64) output columns with their taps in
//    registers, every mid-stage row is one broadcast 128-bit shared load per 8 inputs, and the residual goes
//    straight into the picture -- it never exists in HBM.
// Only the non-zero window of the dense int32 coefficients is read (what the reference's butterflies read).
#include "common.cuh"
#include "coeff_src.cuh"
#include "tables.cuh"

namespace {

// CUDA's __dp2a_* and __ldg are `asm volatile`: the compiler keeps them in program order, so a loop of them runs as one
// dependent chain behind its loads.  These are the same instructions as plain asm (pure functions of their operands;
// the tables and coefficients they read are constant for the launch), which lets independent chains interleave.
__device__ __forceinline__ int dp2a_lo(int a, int b, int c) { int d; asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_hi(int a, int b, int c) { int d; asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint4 ldg_nc(const uint4 *p) { uint4 v; asm("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ uint2 ldg_nc(const uint2 *p) { uint2 v; asm("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ldg_nc(const uint32_t *p) { uint32_t v; asm("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }

constexpr int kWarps = 4, kThreads = kWarps * 32;
#ifndef ITX_MID_PITCH
#define ITX_MID_PITCH 40
#endif
constexpr int P2 = ITX_MID_PITCH;             // pitch (int16) of the mid-stage rows: 32 inputs + pad, 16-byte multiple

struct ItxW {
    pel       *plane[3];
    int        pitch[3];
    long long  bstride[3];
    CoefSrc    src;
    int32_t   *store;               // DENSE32 buffer for VVC_CUDA_TB_STORE_RESIDUAL blocks (NULL otherwise)
    const VVCCudaTB *tbs;
    int        n_tbs;
    uint32_t  *counts;              // [0..5] blocks per class, [8] work counter (zeroed per launch)
    uint32_t  *lists;               // class c: lists[c * n_tbs ..): 0/1/2 = 4096/2048/1024+ samples, 3 = smaller, 4 = left to
                                    // itx_kernel (transform skip, BDPCM, 1-D), 5 = 2x2 .. 4x4 blocks of itx_tiny_kernel
};

// Packed matrices: for output i of (type, n) the 8 words g_wpt[(base + i) * 8 + q] hold taps M[4q .. 4q+3][i]
// (zero past the matrix's rows).  base(type, log2 n): DCT2 sizes 2..64, then DST7 4..32, then DCT8 4..32.
__device__ __align__(16) uint32_t g_wpt[246 * 8];

__host__ __device__ __forceinline__ int wpt_base(int type, int l2)
{
    return type == 0 ? (1 << l2) - 2 : (type == 1 ? 126 : 186) + (1 << l2) - 4;
}

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

__global__ void itx_pack_kernel()
{
    for (int type = 0; type < 3; type++)
        for (int l2 = type ? 2 : 1; l2 <= (type ? 5 : 6); l2++) {
            const int n = 1 << l2, rows = n == 64 ? 32 : n;
            const int8_t *m = tx_matrix(type, n);
            for (int idx = threadIdx.x; idx < n * 8; idx += blockDim.x) {
                const int i = idx >> 3, q = idx & 7;
                uint32_t word = 0;
                for (int k = 0; k < 4; k++)
                    if (4 * q + k < rows)
                        word |= (uint32_t)(uint8_t)m[(4 * q + k) * n + i] << (8 * k);
                g_wpt[(wpt_base(type, l2) + i) * 8 + q] = word;
            }
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

__device__ __forceinline__ int round_rd(int rd) { return rd <= 4 ? 4 : rd <= 8 ? 8 : rd <= 16 ? 16 : 32; }

template <int MODE>
struct __align__(16) WarpSmem {
    short mid[64 * P2];          // pass-1 output: [row][input of pass 2]
    short lf[8 * 8];             // LFNST output window [row][col]
    short win[32 * 32];          // dequantised input window [row][col] (pitch 32) when dequant() runs here (MODE & 2)
};

// Park area of a block with LMCS chroma residual scaling: int32 [h][w], h, w <= 32 (chroma blocks), the last 4 KB of the
// warp's shared memory - mid-stage rows 40.. (pass 2 of such a block reads rows below 32), the LFNST window and the
// dequantised window, all dead once pass 1 is done.
template <int MODE>
__device__ __forceinline__ int32_t *park_of(WarpSmem<MODE> &s)
{
    static_assert(sizeof(WarpSmem<MODE>) >= 4096 + 32 * P2 * sizeof(short), "park area overlaps live mid-stage rows");
    return reinterpret_cast<int32_t *>(reinterpret_cast<char *>(&s) + sizeof(WarpSmem<MODE>) - 4096);
}

__constant__ uint8_t c_diag4_x[16] = { 0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 1, 2, 3, 2, 3, 3 };
__constant__ uint8_t c_diag4_y[16] = { 0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 3, 2, 1, 3, 2, 3 };

// Pass 1: mid[i][x] = clip16((sum_j in[j][x] * M[j][i] + 64) >> 7).  lane = (column x, output group g).
template <int RD, int MODE>
__device__ __forceinline__ void pass1(WarpSmem<MODE> &s, const TbCoef &tc, const short *lf, int lfp, int nzw, int rd, int h,
                                      const uint32_t *wpt, int lane)
{
    int xb = 0;
    while ((1 << xb) < nzw)
        xb++;
    const int x = lane & ((1 << xb) - 1), g = lane >> xb, ng = 32 >> xb;
    if (x >= nzw)
        return;
    uint32_t in[RD / 2];
#pragma unroll
    for (int jp = 0; jp < RD / 2; jp++) {
        int v0 = 0, v1 = 0;
        if (lf) {
            if (2 * jp < rd)     v0 = lf[(2 * jp) * lfp + x];
            if (2 * jp + 1 < rd) v1 = lf[(2 * jp + 1) * lfp + x];
        } else {
            if (2 * jp < rd)     v0 = coef_load<MODE>(tc, 2 * jp, x);
            if (2 * jp + 1 < rd) v1 = coef_load<MODE>(tc, 2 * jp + 1, x);
        }
        in[jp] = (uint32_t)(v0 & 0xffff) | ((uint32_t)v1 << 16);
    }
    // two outputs per iteration: their matrix words are requested together and the two tap chains are independent
    auto taps = [&](const uint4 (&m)[RD >= 16 ? RD / 16 : 1]) -> int {
        int acc = 0;
        if (RD == 4) {
            acc = dp2a_lo((int)in[0], (int)m[0].x, dp2a_hi((int)in[1], (int)m[0].x, 0));
        } else if (RD == 8) {
            acc = dp2a_lo((int)in[0], (int)m[0].x, dp2a_hi((int)in[1], (int)m[0].x, 0)) +
                  dp2a_lo((int)in[2], (int)m[0].y, dp2a_hi((int)in[3], (int)m[0].y, 0));
        } else {
            int a0 = 0, a1 = 0;
#pragma unroll
            for (int q4 = 0; q4 < RD / 16; q4++) {
                a0 = dp2a_lo((int)in[8 * q4 + 0], (int)m[q4].x, dp2a_hi((int)in[8 * q4 + 1], (int)m[q4].x, a0));
                a1 = dp2a_lo((int)in[8 * q4 + 2], (int)m[q4].y, dp2a_hi((int)in[8 * q4 + 3], (int)m[q4].y, a1));
                a0 = dp2a_lo((int)in[8 * q4 + 4], (int)m[q4].z, dp2a_hi((int)in[8 * q4 + 5], (int)m[q4].z, a0));
                a1 = dp2a_lo((int)in[8 * q4 + 6], (int)m[q4].w, dp2a_hi((int)in[8 * q4 + 7], (int)m[q4].w, a1));
            }
            acc = a0 + a1;
        }
        return acc;
    };
    auto fetch = [&](int i, uint4 (&m)[RD >= 16 ? RD / 16 : 1]) {
        const uint4 *mp = reinterpret_cast<const uint4 *>(wpt + i * 8);
        if (RD == 4) {
            m[0].x = ldg_nc(reinterpret_cast<const uint32_t *>(mp));
        } else if (RD == 8) {
            const uint2 v = ldg_nc(reinterpret_cast<const uint2 *>(mp));
            m[0].x = v.x; m[0].y = v.y;
        } else {
#pragma unroll
            for (int q4 = 0; q4 < RD / 16; q4++)
                m[q4] = ldg_nc(mp + q4);
        }
    };
    for (int i = g; i < h; i += 2 * ng) {
        uint4 ma[RD >= 16 ? RD / 16 : 1], mb[RD >= 16 ? RD / 16 : 1];
        const int i2 = i + ng < h ? i + ng : i;             // odd tail: the second chain repeats the first
        fetch(i, ma);
        fetch(i2, mb);
        const int ra = taps(ma), rb = taps(mb);
        s.mid[i * P2 + x] = (short)d_clip_sbits((ra + 64) >> 7, 15);
        s.mid[i2 * P2 + x] = (short)d_clip_sbits((rb + 64) >> 7, 15);
    }
}

struct Epi {
    int32_t *store;              // != NULL: residual written back as int32 [h][w]
    pel     *d0, *d1;            // picture samples of the TB's first row (d1: joint CbCr plane or NULL)
    int      pitch0, pitch1, sign, shift, w;
};

// Pass 2: out[y][i] = (sum_x mid[y][x] * M[x][i] + 512) >> 10.  lane = (4 adjacent output columns, row group):
// the taps of its 4 columns live in registers, a mid-stage row is one broadcast 128-bit shared load per 8
// inputs, and the picture is updated with one 64-bit load + store per row (32-bit for 2-wide blocks); the
// samples of the next row are requested before the current row's products so their latency is hidden.
template <int RD, int MODE>
__device__ __forceinline__ void pass2(const WarpSmem<MODE> &s, int l2w, int h, const uint32_t *wpt, const Epi &e, int lane)
{
    const int qb = max(l2w - 2, 0), cq = lane & ((1 << qb) - 1), g = lane >> qb, ng = 32 >> qb;
    const int c0 = cq << 2;
    const bool narrow = l2w == 1;                          // 2 columns
    uint32_t m[4][RD / 4];
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int q = 0; q < RD / 4; q++)
            m[c][q] = (narrow && c >= 2) ? 0u : ldg_nc(wpt + (c0 + c) * 8 + q);
    const int step0 = ng * e.pitch0, step1 = ng * e.pitch1;
    pel *d0 = e.d0 + g * e.pitch0 + c0, *d1 = e.d1 ? e.d1 + g * e.pitch1 + c0 : nullptr;
    auto load4 = [&](const pel *p) -> uint2 {
        if (narrow)
            return make_uint2(*reinterpret_cast<const uint32_t *>(p), 0u);
        return *reinterpret_cast<const uint2 *>(p);
    };
    uint2 cur0 = make_uint2(0u, 0u), cur1 = make_uint2(0u, 0u);
    if (!e.store && g < h) {
        cur0 = load4(d0);
        if (d1) cur1 = load4(d1);
    }
    for (int y = g; y < h; y += ng, d0 += step0, d1 += step1) {
        uint2 nxt0 = make_uint2(0u, 0u), nxt1 = make_uint2(0u, 0u);
        if (!e.store && y + ng < h) {
            nxt0 = load4(d0 + step0);
            if (d1) nxt1 = load4(d1 + step1);
        }
        uint32_t in[RD / 2];
        if (RD == 4) {
            const uint2 v = *reinterpret_cast<const uint2 *>(&s.mid[y * P2]);
            in[0] = v.x; in[1] = v.y;
        } else {
#pragma unroll
            for (int c = 0; c < RD / 8; c++) {
                const uint4 v = *reinterpret_cast<const uint4 *>(&s.mid[y * P2 + 8 * c]);
                in[4 * c] = v.x; in[4 * c + 1] = v.y; in[4 * c + 2] = v.z; in[4 * c + 3] = v.w;
            }
        }
        int r[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            int a = 0;
#pragma unroll
            for (int q = 0; q < RD / 4; q++)
                a = dp2a_lo((int)in[2 * q], (int)m[c][q], dp2a_hi((int)in[2 * q + 1], (int)m[c][q], a));
            r[c] = (a + 512) >> 10;                         // shift = 5 + log2_transform_range - bit_depth
        }
        if (e.store) {
            int32_t *o = e.store + y * e.w + c0;
            o[0] = r[0]; o[1] = r[1];
            if (!narrow) { o[2] = r[2]; o[3] = r[3]; }
        } else {
            const uint32_t lo = (uint32_t)d_clip_pel((int)(cur0.x & 0xffff) + r[0], 10) | ((uint32_t)d_clip_pel((int)(cur0.x >> 16) + r[1], 10) << 16);
            const uint32_t hi = (uint32_t)d_clip_pel((int)(cur0.y & 0xffff) + r[2], 10) | ((uint32_t)d_clip_pel((int)(cur0.y >> 16) + r[3], 10) << 16);
            if (narrow) *reinterpret_cast<uint32_t *>(d0) = lo;
            else        *reinterpret_cast<uint2 *>(d0) = make_uint2(lo, hi);
            if (d1) {
                const uint32_t jl = (uint32_t)d_clip_pel((int)(cur1.x & 0xffff) + ((r[0] * e.sign) >> e.shift), 10)
                                  | ((uint32_t)d_clip_pel((int)(cur1.x >> 16) + ((r[1] * e.sign) >> e.shift), 10) << 16);
                const uint32_t jh = (uint32_t)d_clip_pel((int)(cur1.y & 0xffff) + ((r[2] * e.sign) >> e.shift), 10)
                                  | ((uint32_t)d_clip_pel((int)(cur1.y >> 16) + ((r[3] * e.sign) >> e.shift), 10) << 16);
                if (narrow) *reinterpret_cast<uint32_t *>(d1) = jl;
                else        *reinterpret_cast<uint2 *>(d1) = make_uint2(jl, jh);
            }
        }
        cur0 = nxt0; cur1 = nxt1;
    }
}

// Epilogue of a block with LMCS chroma residual scaling (lmcs_scale_chroma between the transform and add_residual,
// itransform vvc_intra.c:468-475; the second plane of a joint block is derived first and scaled afterwards, :179-183), from
// the residuals pass 2 parked in shared memory (through its residual-store path: Epi.store pointing at the warp's own
// park area, so the common path carries no extra instruction).  Luma blocks and pictures without
// ph_chroma_residual_scale_flag never come here.
__device__ __noinline__ void scaled_epilogue(const int32_t *park, int w, int h, int scale, pel *d0, int pitch0, pel *d1, int pitch1,
                                             int sign, int shift, int lane)
{
    const int lw = 31 - __clz(w);
    for (int i = lane; i < w * h; i += 32) {
        const int y = i >> lw, x = i & (w - 1), r = park[i];
        pel *p = d0 + (long long)y * pitch0 + x;
        *p = (pel)d_clip_pel(*p + d_lmcs_scale(r, scale, 10), 10);
        if (d1) {
            pel *q = d1 + (long long)y * pitch1 + x;
            *q = (pel)d_clip_pel(*q + d_lmcs_scale((r * sign) >> shift, scale, 10), 10);
        }
    }
}

// 64-bit picture accesses need blocks of 4+ columns to start on a multiple of 4 samples (a 4-wide chroma block
// under an 8-wide luma CU at x = 4 does not: left to the generic kernel)
__device__ __forceinline__ bool eligible(int l2w, int l2h, int flags, int x0)
{
    return l2w >= 1 && l2h >= 1 && !(flags & (VVC_CUDA_TB_TS | VVC_CUDA_TB_BDPCM | VVC_CUDA_TB_BDPCM_VERT)) &&
           (l2w == 1 ? !(x0 & 1) : !(x0 & 3));
}

// Work lists.  A 64x64 block keeps a warp busy a hundred times longer than a 4x4 one, and with a fixed stride the
// launch lasted as long as the warp that happened to draw the most large blocks.  The blocks are binned by size on
// the device; the transform kernel then hands them out through a counter, largest first, small ones in chunks.
constexpr int kSmallChunk = 16;

__global__ void __launch_bounds__(256) itx_sort_kernel(const ItxW p)
{
    const int ti = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31;
    int cls = -1;
    if (ti < p.n_tbs) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p.tbs + ti);
        const uint32_t r1 = __ldg(q + 1), r2 = __ldg(q + 2), r3 = __ldg(q + 3);
        const int l2w = r2 & 0xff, l2h = (r2 >> 8) & 0xff, flags = r3 >> 24, lfnst = __ldg(q + 4) & 0xff;
        if (!eligible(l2w, l2h, flags, r1 & 0xffff) || ((__ldg(q + 5) >> 16) && (l2w > 5 || l2h > 5))) cls = 4;   // (scaled 4:4:4 chroma blocks above 32: no room to park)
        else if (l2w <= 2 && l2h <= 2 && !lfnst && !(flags & VVC_CUDA_TB_STORE_RESIDUAL) && !(__ldg(q + 5) >> 16)) cls = 5;   // (not with LMCS chroma scaling)
        else cls = l2w + l2h >= 12 ? 0 : l2w + l2h == 11 ? 1 : l2w + l2h == 10 ? 2 : 3;
    }
    // one reservation per class and warp, all six requested before the first answer is awaited (one after the other, a warp
    // paid a round trip to the counters per class it holds)
    unsigned m[6];
    uint32_t base[6];
#pragma unroll
    for (int c = 0; c < 6; c++) {
        m[c] = __ballot_sync(0xffffffffu, cls == c);
        base[c] = 0;
        if (m[c] && lane == __ffs(m[c]) - 1)
            base[c] = atomicAdd(p.counts + c, (uint32_t)__popc(m[c]));
    }
#pragma unroll
    for (int c = 0; c < 6; c++) {
        if (!m[c])
            continue;
        const uint32_t b = __shfl_sync(0xffffffffu, base[c], __ffs(m[c]) - 1);
        if (cls == c)
            p.lists[(size_t)c * p.n_tbs + b + __popc(m[c] & ((1u << lane) - 1))] = ti;
    }
}

// Blocks of 2x2 .. 4x4 samples (44 % of the blocks of a picture, 1.5 % of its samples): one THREAD per block.  The warp
// kernel spends ~400 warp-instructions of bookkeeping on each of them with 28 lanes idle; here a block is ~300
// thread-instructions: up to 16 coefficients, two passes of 4-tap IDP.2A pairs from the packed matrices, the residual
// added with one 32- or 64-bit access per row.  Same arithmetic as itx_warp_kernel (same inputs read, same clips).
template <int MODE>
__global__ void __launch_bounds__(128) itx_tiny_kernel(const ItxW p)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= (int)p.counts[5])
        return;
    const int ti = (int)__ldg(p.lists + 5 * (size_t)p.n_tbs + i);
    const uint32_t *q = reinterpret_cast<const uint32_t *>(p.tbs + ti);
    const uint32_t r0 = __ldg(q), r1 = __ldg(q + 1), r2 = __ldg(q + 2), r3 = __ldg(q + 3), r4 = __ldg(q + 4), r5 = __ldg(q + 5);
    const int l2w = r2 & 0xff, l2h = (r2 >> 8) & 0xff, c_idx = (r2 >> 16) & 0xff, flags = r3 >> 24;
    const int x0 = r1 & 0xffff, y0 = r1 >> 16, w = 1 << l2w, h = 1 << l2h;
    const int trh = r2 >> 24, trv = r3 & 0xff, nzw = (r3 >> 8) & 0xff, nzh = (r3 >> 16) & 0xff;
    const int jsign = (int8_t)((r4 >> 8) & 0xff), jshift = (r4 >> 16) & 0xff, jc = r4 >> 24, pic = r5 & 0xff;
    const TbCoef tc = tb_coef<MODE>(p.src, ti, r0, l2w, l2h, nzw, nzh, false);
    const bool dc_only = trh == 0 && trv == 0 && nzw == 1 && nzh == 1 && w == h;
    const int rdv = dc_only ? 1 : inputs_read(trv, h, nzh);
    // pass 1 per column x < nzw: inputs (rows 0..rdv-1) as two pairs, outputs i < h
    uint32_t mid01[4] = { 0, 0, 0, 0 }, mid23[4] = { 0, 0, 0, 0 };       // row y: (col 0, col 1), (col 2, col 3)
    uint32_t wv[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
        wv[k] = k < h ? ldg_nc(g_wpt + (wpt_base(trv, l2h) + k) * 8) : 0u;
#pragma unroll
    for (int x = 0; x < 4; x++) {
        if (x < nzw && x < w) {
            int v[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                v[j] = j < rdv ? coef_load<MODE>(tc, j, x) : 0;
            const uint32_t p01 = (uint32_t)(v[0] & 0xffff) | ((uint32_t)v[1] << 16), p23 = (uint32_t)(v[2] & 0xffff) | ((uint32_t)v[3] << 16);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k < h) {
                    const int acc = dp2a_lo((int)p01, (int)wv[k], dp2a_hi((int)p23, (int)wv[k], 0));
                    const uint32_t m = (uint32_t)d_clip_sbits((acc + 64) >> 7, 15) & 0xffff;
                    if (x < 2) mid01[k] |= m << (16 * (x & 1));
                    else       mid23[k] |= m << (16 * (x & 1));
                }
            }
        }
    }
    // pass 2 per row + epilogue
    uint32_t wh[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
        wh[k] = k < w ? ldg_nc(g_wpt + (wpt_base(trh, l2w) + k) * 8) : 0u;
#define SEL3(a, c) ((c) == 0 ? (a)[0] : (c) == 1 ? (a)[1] : (a)[2])
    const int pitch0 = SEL3(p.pitch, c_idx);
    pel *d0 = SEL3(p.plane, c_idx) + pic * SEL3(p.bstride, c_idx) + (long long)y0 * pitch0 + x0;
    pel *d1 = nullptr;
    int pitch1 = 0;
    if (flags & VVC_CUDA_TB_JOINT) {
        pitch1 = SEL3(p.pitch, jc);
        d1 = SEL3(p.plane, jc) + pic * SEL3(p.bstride, jc) + (long long)y0 * pitch1 + x0;
    }
#undef SEL3
#pragma unroll
    for (int y = 0; y < 4; y++) {
        if (y < h) {
            int r[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
                r[k] = (dp2a_lo((int)mid01[y], (int)wh[k], dp2a_hi((int)mid23[y], (int)wh[k], 0)) + 512) >> 10;
            auto add2 = [&](uint32_t cur, int a, int b) -> uint32_t {
                return (uint32_t)d_clip_pel((int)(cur & 0xffff) + a, 10) | ((uint32_t)d_clip_pel((int)(cur >> 16) + b, 10) << 16);
            };
            if (w == 4) {
                uint2 *t = reinterpret_cast<uint2 *>(d0 + y * pitch0);
                const uint2 cur = *t;
                *t = make_uint2(add2(cur.x, r[0], r[1]), add2(cur.y, r[2], r[3]));
                if (d1) {
                    uint2 *u = reinterpret_cast<uint2 *>(d1 + y * pitch1);
                    const uint2 c1 = *u;
                    *u = make_uint2(add2(c1.x, (r[0] * jsign) >> jshift, (r[1] * jsign) >> jshift),
                                    add2(c1.y, (r[2] * jsign) >> jshift, (r[3] * jsign) >> jshift));
                }
            } else {
                uint32_t *t = reinterpret_cast<uint32_t *>(d0 + y * pitch0);
                *t = add2(*t, r[0], r[1]);
                if (d1) {
                    uint32_t *u = reinterpret_cast<uint32_t *>(d1 + y * pitch1);
                    *u = add2(*u, (r[0] * jsign) >> jshift, (r[1] * jsign) >> jshift);
                }
            }
        }
    }
}

#ifndef ITX_WARP_MB
#define ITX_WARP_MB 7            // resident CTAs per SM the kernel is compiled for (tools/sweep_itx.sh: 1 / 6 / 7 / 8 ->
                                 // 0.80 / 0.65 / 0.63 / 0.65 ms per 8 pictures once the work is balanced)
#endif
template <int MODE>
__global__ void __launch_bounds__(kThreads, ITX_WARP_MB) itx_warp_kernel(const ItxW p)
{
    __shared__ WarpSmem<MODE> sm[kWarps];
    const int lane = threadIdx.x & 31;
    WarpSmem<MODE> &s = sm[threadIdx.x >> 5];
    const uint32_t nA = p.counts[0], nB = p.counts[1], nC = p.counts[2], nS = p.counts[3];
    const uint32_t nbig = nA + nB + nC, total = nbig + (nS + kSmallChunk - 1) / kSmallChunk;
    for (;;) {
        uint32_t v = 0;
        if (lane == 0)
            v = atomicAdd(p.counts + 8, 1u);
        v = __shfl_sync(0xffffffffu, v, 0);
        if (v >= total)
            break;
        const uint32_t *list;
        int cnt = 1;
        if (v < nA)             list = p.lists + v;
        else if (v < nA + nB)   list = p.lists + (size_t)p.n_tbs + (v - nA);
        else if (v < nbig)      list = p.lists + 2 * (size_t)p.n_tbs + (v - nA - nB);
        else {
            const uint32_t c = (v - nbig) * kSmallChunk;
            list = p.lists + 3 * (size_t)p.n_tbs + c;
            cnt = min(kSmallChunk, (int)(nS - c));
        }
    for (int k = 0; k < cnt; k++) {
        const int ti = (int)__ldg(list + k);
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p.tbs + ti);
        const uint32_t r0 = __ldg(q), r1 = __ldg(q + 1), r2 = __ldg(q + 2), r3 = __ldg(q + 3), r4 = __ldg(q + 4), r5 = __ldg(q + 5);
        const int l2w = r2 & 0xff, l2h = (r2 >> 8) & 0xff, c_idx = (r2 >> 16) & 0xff, flags = r3 >> 24;
        const int x0 = r1 & 0xffff, y0 = r1 >> 16, w = 1 << l2w, h = 1 << l2h;
        int trh = r2 >> 24, trv = r3 & 0xff, nzw = (r3 >> 8) & 0xff, nzh = (r3 >> 16) & 0xff;
        const int lfnst = r4 & 0xff, jsign = (int8_t)((r4 >> 8) & 0xff), jshift = (r4 >> 16) & 0xff, jc = r4 >> 24, pic = r5 & 0xff;
        const TbCoef tc = tb_coef<MODE>(p.src, ti, r0, l2w, l2h, nzw, nzh, false);
        // the warp's next block of the chunk: its record is requested now, its coefficient window is prefetched into L2 once the
        // record has arrived (after pass 1), so the next iteration does not start with a DRAM round trip
        const int tn = k + 1 < cnt ? (int)__ldg(list + k + 1) : -1;
        uint32_t nr0 = 0, nr2 = 0, nr3 = 0;
        if (tn >= 0) {
            const uint32_t *qn = reinterpret_cast<const uint32_t *>(p.tbs + tn);
            nr0 = __ldg(qn); nr2 = __ldg(qn + 2); nr3 = __ldg(qn + 3);
        }
        __syncwarp();                                       // previous block is done with shared memory

        // ---- inverse LFNST: 8/16 inputs in 4x4 diagonal order -> 16/48 outputs, top-left 4x4 / 8x8 ----
        const short *lf = nullptr;
        int lfp = 8;
        if (lfnst) {
            const int idx = lfnst & 3, set = (lfnst >> 2) & 3, side = (w >= 8 && h >= 8) ? 8 : 4;
            const bool transpose = (lfnst >> 4) & 1;
            const int n_in = ((lfnst >> 5) & 1) ? 8 : 16, n_out = side == 8 ? 48 : 16;
            const int8_t *M = side == 8 ? &vvct_lfnst_8x8[set][idx - 1][0][0] : &vvct_lfnst_4x4[set][idx - 1][0][0];
            s.lf[lane] = 0; s.lf[lane + 32] = 0;
            const int u = lane < n_in ? coef_load<MODE>(tc, c_diag4_y[lane & 15], c_diag4_x[lane & 15]) : 0;
            __syncwarp();
            for (int base = 0; base < n_out; base += 32) {
                const bool act = base + lane < n_out;
                const int j = act ? base + lane : 0;
                int acc = 0;
                for (int i = 0; i < n_in; i++)
                    acc += __shfl_sync(0xffffffffu, u, i) * (int)M[i * n_out + j];
                const int v = d_clip_sbits((acc + 64) >> 7, 15);
                const int rr = j < 4 * side ? j / side : 4 + ((j - 4 * side) >> 2);
                const int qq = j < 4 * side ? j % side : (j - 4 * side) & 3;
                if (act)
                    s.lf[transpose ? qq * 8 + rr : rr * 8 + qq] = (short)v;
            }
            nzw = nzh = side;
            lf = s.lf;
            __syncwarp();
        }
        // DC-only cells of square DCT2 x DCT2 blocks read nothing but c[0] (vvcdsp.c:101-108)
        const bool dc_only = trh == 0 && trv == 0 && nzw == 1 && nzh == 1 && w == h;
        const int rdv = dc_only ? 1 : inputs_read(trv, h, nzh), rdh = inputs_read(trh, w, nzw);
        const int rdv_e = round_rd(rdv), rdh_e = round_rd(rdh);

        // ---- dequant(): the rows pass 1 reads, all lanes over the window (one compact loop instead of a copy of
        // scale_coeff per unrolled register of pass 1) ----
        if ((MODE & 2) && !lf) {
            int xb = 0;
            while ((1 << xb) < nzw)
                xb++;
#pragma unroll 1
            for (int idx = lane; idx < (rdv << xb); idx += 32) {
                const int y = idx >> xb, x = idx & ((1 << xb) - 1);
                if (x < nzw)
                    s.win[y * 32 + x] = (short)coef_load<MODE>(tc, y, x);
            }
            lf = s.win;
            lfp = 32;
            __syncwarp();
        }
        // ---- pass 1 ----
        {
            const uint32_t *wpt = g_wpt + wpt_base(trv, l2h) * 8;
            switch (rdv_e) {
            case 4:  pass1<4, MODE>(s, tc, lf, lfp, nzw, rdv, h, wpt, lane); break;
            case 8:  pass1<8, MODE>(s, tc, lf, lfp, nzw, rdv, h, wpt, lane); break;
            case 16: pass1<16, MODE>(s, tc, lf, lfp, nzw, rdv, h, wpt, lane); break;
            default: pass1<32, MODE>(s, tc, lf, lfp, nzw, rdv, h, wpt, lane); break;
            }
        }
        if (tn >= 0) {
            const int nl2w = nr2 & 0xff, nl2h = (nr2 >> 8) & 0xff, nnzw = (nr3 >> 8) & 0xff, nnzh = (nr3 >> 16) & 0xff;
            if (MODE & 1) {
                const char *b = reinterpret_cast<const char *>(p.src.window + nr0);
                if (lane * 64 < nnzw * nnzh * 2)
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(b + lane * 64));
            } else {
                const int rows = min(round_rd(max(nnzh, 2)), 1 << nl2h);
                if (lane < rows && nnzw > 0) {
                    const int32_t *row = p.src.dense + nr0 + (lane << nl2w);
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(row));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(row + nnzw - 1));
                }
            }
        }
        // columns nzw .. of the mid stage are zero (scale_clip's memset); pad to the rounded reduction length
        if (rdh_e > nzw) {
            // lane = (column of the pad, row group): h * pad stores take ceil(h * pad' / 32) rounds, pad' = pad rounded up
            // to a power of two (most blocks: a pad of 1 - 3 columns, one or two rounds)
            const int pad = rdh_e - nzw;                    // 1 .. 31
            const int lp = 32 - __clz(pad - 1);
            const int k = lane & ((1 << lp) - 1);
            if (k < pad)
                for (int y = lane >> lp; y < h; y += 32 >> lp)
                    s.mid[y * P2 + nzw + k] = 0;
        }
        __syncwarp();

        // ---- pass 2 + epilogue ----
        Epi e;
        e.w = w;
        e.store = (flags & VVC_CUDA_TB_STORE_RESIDUAL) ? p.store + r0 : nullptr;
#define SEL3(a, c) ((c) == 0 ? (a)[0] : (c) == 1 ? (a)[1] : (a)[2])
        e.pitch0 = SEL3(p.pitch, c_idx);
        e.d0 = SEL3(p.plane, c_idx) + pic * SEL3(p.bstride, c_idx) + (long long)y0 * e.pitch0 + x0;
        e.d1 = nullptr; e.pitch1 = 0; e.sign = jsign; e.shift = jshift;
        const int cscale = e.store ? 0 : tb_chroma_scale(p.src, (int)(r5 >> 16));
        if (cscale)                 // pass 2 "stores" the residual of a scaled block into the warp's park area
            e.store = park_of(s);
        if (flags & VVC_CUDA_TB_JOINT) {
            e.pitch1 = SEL3(p.pitch, jc);
            e.d1 = SEL3(p.plane, jc) + pic * SEL3(p.bstride, jc) + (long long)y0 * e.pitch1 + x0;
#undef SEL3
        }
        {
            const uint32_t *wpt = g_wpt + wpt_base(trh, l2w) * 8;
            switch (rdh_e) {
            case 4:  pass2<4, MODE>(s, l2w, h, wpt, e, lane); break;
            case 8:  pass2<8, MODE>(s, l2w, h, wpt, e, lane); break;
            case 16: pass2<16, MODE>(s, l2w, h, wpt, e, lane); break;
            default: pass2<32, MODE>(s, l2w, h, wpt, e, lane); break;
            }
        }
        if (cscale) {
            __syncwarp();
            scaled_epilogue(park_of(s), w, h, cscale, e.d0, e.pitch0, e.d1, e.pitch1, jsign, jshift, lane);
        }
    }
    }
}

}  // namespace

// Launch over the whole list.  scratch: 16 + 6 * n_tbs words; blocks this kernel does not handle (transform skip,
// BDPCM, 1-D) end up in list 4 (*rest, count in *rest_count) for itx_kernel (itx.cu).
int vvc_itx_launch_warp(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaCoeffs *co, const VVCCudaTB *tbs, int n_tbs,
                        uint32_t *scratch, const uint32_t **rest, const uint32_t **rest_count)
{
    if (!ctx->itx_packed) {
        itx_pack_kernel<<<1, 256, 0, ctx->stream>>>();
        if (vvc_ctx_check(ctx, cudaGetLastError(), "itx_pack_kernel"))
            return ctx->err;
        ctx->itx_packed = true;
    }
    VVC_TRY(ctx, cudaMemsetAsync(scratch, 0, 16 * sizeof(uint32_t), ctx->stream));
    ItxW p;
    for (int c = 0; c < 3; c++) {
        p.plane[c] = (pel *)frame->data[c];
        p.pitch[c] = (int)(frame->stride[c] / 2);
        p.bstride[c] = frame->batch_stride[c] / 2;
    }
    const int mode = coef_mode(co);
    p.src.dense = (mode & 1) ? nullptr : (const int32_t *)co->data;
    p.src.window = (mode & 1) ? (const int16_t *)co->data : nullptr;
    p.src.quant = co->quant; p.src.scaling = co->scaling; p.src.lmcs_scales = co->lmcs_scales; p.src.range = 15; p.src.bd = 10;
    p.store = (mode & 1) ? nullptr : (int32_t *)co->data;
    p.tbs = tbs; p.n_tbs = n_tbs; p.counts = scratch; p.lists = scratch + 16;
    *rest = p.lists + 4 * (size_t)n_tbs; *rest_count = p.counts + 4;
    itx_sort_kernel<<<ceil_div(n_tbs, 256), 256, 0, ctx->stream>>>(p);
    VVC_LAUNCHED(ctx);
    if (vvc_ctx_fork(ctx, 1))           // side stream 0: the tiny blocks here, then the caller's itx_kernel; the caller joins
        return ctx->err;
    {
        const int tg = ceil_div(n_tbs, 128);
        switch (mode) {
        case 0:  itx_tiny_kernel<0><<<tg, 128, 0, ctx->side[0]>>>(p); break;
        case 1:  itx_tiny_kernel<1><<<tg, 128, 0, ctx->side[0]>>>(p); break;
        case 2:  itx_tiny_kernel<2><<<tg, 128, 0, ctx->side[0]>>>(p); break;
        default: itx_tiny_kernel<3><<<tg, 128, 0, ctx->side[0]>>>(p); break;
        }
        VVC_LAUNCHED(ctx);
    }
    const int ctas = ceil_div(n_tbs, kWarps), grid = ctas < 148 * ITX_WARP_MB ? ctas : 148 * ITX_WARP_MB;
    switch (mode) {
    case 0:  itx_warp_kernel<0><<<grid, kThreads, 0, ctx->stream>>>(p); break;
    case 1:  itx_warp_kernel<1><<<grid, kThreads, 0, ctx->stream>>>(p); break;
    case 2:  itx_warp_kernel<2><<<grid, kThreads, 0, ctx->stream>>>(p); break;
    default: itx_warp_kernel<3><<<grid, kThreads, 0, ctx->stream>>>(p); break;
    }
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}
