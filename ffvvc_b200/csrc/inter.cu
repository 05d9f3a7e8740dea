// Inter prediction stage for sm_100a: one launch predicts every motion-compensated block of a
// picture ring from its record list.
//
// Replaces the pixel work of ff_vvc_predict_inter (libavcodec/vvc/vvc_inter.c:899-913):
//   pred_regular_blk :782-811 (derive_sb_mv :764-780, dmvr_mv_refine :685-748 with
//   parametric_mv_refine :642-681, pred_regular_luma :545-578, pred_regular_chroma :580-639),
//   pred_affine_blk :828-873 (luma_prof_uni :368-408, luma_prof_bi :410-446), pred_gpm_blk :466-521,
//   the edge emulation :33-110 (libavcodec/videodsp_template.c:26-105),
// and the table entries they call: put / put_uni / put_uni_w (libavcodec/h26x/h2656_inter_template.c
// :29-577), avg, w_avg, put_gpm, bdof_fetch_samples, fetch_samples, prof_grad_filter, apply_prof*,
// apply_bdof, dmvr* (libavcodec/vvc/vvc_inter_template.c:25-436), sad and pad_int16
// (libavcodec/vvc/vvcdsp.c:29-65).
//
// B200 design: a CTA of 128 threads owns one record at a time (persistent grid-stride loop).  The
// reference windows of both lists are staged once into shared memory with the clamp the
// reference materialises through emulated_edge_mc (picture, or for DMVR blocks the unrefined
// block's window), then the separable filter runs as two shared-memory passes; DMVR (bilinear +
// 25 SADs + refinement), BDOF and PROF work entirely out of shared memory, so DRAM traffic is the
// reference window in and the predicted block out.  The whole DMVR decision stays on the device;
// refined vectors are written to dmvr_out for the host's later motion-vector prediction.
#include "inter_common.cuh"
#include "tables.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int TP = 20;                 // row pitch of the prediction tiles (16 + ring, padded)
constexpr int GP = 18;                 // BDOF_PADDED_SIZE
constexpr int WIN_LUMA = 23 * 24;      // (16 + 7) rows, pitch 24
constexpr int WIN_CHROMA = 11 * 12;


// one reference fetch unit: a list (luma) or a (plane, list) pair (chroma)
struct Unit {
    const pel *plane;
    int        pitch;
    int        xlo, xhi, ylo, yhi;     // clamp window
    int        ox, oy;                 // integer position of the block in the reference plane
    int        mx, my;                 // fractional phase
    int8_t     hf[8], vf[8];
};

struct Smem {
    VVCCudaPB rec;
    Unit      u[4];
    int       mv[2][2];
    int       bdof;
    int       sad[25];
    int       vxy[16][2];
    alignas(16) pel   win[2 * WIN_LUMA];
    alignas(16) short hb[2 * 23 * 16];
    alignas(16) short tile[4][18 * TP];
    alignas(16) short dm[2][20 * 20];
    alignas(16) short grad[2][2][GP * GP];
};

__device__ __forceinline__ int fetch(const Unit &u, int x, int y)
{
    x = d_clip3(x, u.xlo, u.xhi);
    y = d_clip3(y, u.ylo, u.yhi);
    return __ldg(u.plane + (long long)y * u.pitch + x);
}

// Stage, for every unit, the (bw + TAPS - 1) x (bh + TAPS - 1) window whose first sample is
// (ox - B, oy - B); coordinates are clamped like ff_emulated_edge_mc does.
template <int TAPS>
__device__ __forceinline__ void stage_windows(Smem &s, int n_units, int bw, int bh)
{
    constexpr int B = TAPS / 2 - 1, WP = TAPS == 8 ? 24 : 12, WS = TAPS == 8 ? WIN_LUMA : WIN_CHROMA;
    const int cols = bw + TAPS - 1, rows = bh + TAPS - 1;
    for (int idx = threadIdx.x; idx < n_units * rows * WP; idx += kThreads) {
        const int ui = idx / (rows * WP), rem = idx - ui * rows * WP;
        const int r = rem / WP, c = rem - r * WP;
        if (c < cols) {
            const Unit &u = s.u[ui];
            s.win[ui * WS + r * WP + c] = (pel)fetch(u, u.ox - B + c, u.oy - B + r);
        }
    }
}

// First pass: rows of the horizontal filter (or the raw samples when mx == 0), h2656_inter_template.c
// put_*_h / the tmp_array loop of put_*_hv.
template <int TAPS>
__device__ __forceinline__ void pass_h(Smem &s, int n_units, int bw, int bh, int lw, int bd)
{
    constexpr int B = TAPS / 2 - 1, WP = TAPS == 8 ? 24 : 12, WS = TAPS == 8 ? WIN_LUMA : WIN_CHROMA;
    constexpr int HS = TAPS == 8 ? 23 * 16 : 11 * 8, HP = TAPS == 8 ? 16 : 8;
    const int rows = bh + TAPS - 1;
    for (int idx = threadIdx.x; idx < (n_units * rows) << lw; idx += kThreads) {
        const int x = idx & (bw - 1), q = idx >> lw;
        const int ui = q / rows, r = q - ui * rows;
        const Unit &u = s.u[ui];
        const pel *p = &s.win[ui * WS + r * WP + x];
        int v;
        if (u.mx) {
            v = 0;
#pragma unroll
            for (int k = 0; k < TAPS; k++)
                v += u.hf[k] * p[k];
            v >>= bd - 8;
        } else {
            v = p[B];
        }
        s.hb[ui * HS + r * HP + x] = (short)v;
    }
}

// Second pass: the value the reference's put / put_uni computes before its final rounding.
template <int TAPS>
__device__ __forceinline__ int pass_v(const Smem &s, int ui, int x, int y, int bd)
{
    constexpr int B = TAPS / 2 - 1;
    constexpr int HS = TAPS == 8 ? 23 * 16 : 11 * 8, HP = TAPS == 8 ? 16 : 8;
    const Unit &u = s.u[ui];
    const short *p = &s.hb[ui * HS + y * HP + x];
    if (u.my) {
        int v = 0;
#pragma unroll
        for (int k = 0; k < TAPS; k++)
            v += u.vf[k] * (u.mx ? (int)p[k * HP] : (int)(unsigned short)p[k * HP]);
        return v >> (u.mx ? 6 : bd - 8);
    }
    return u.mx ? (int)p[B * HP] : (int)(unsigned short)p[B * HP] << (14 - bd);
}

struct Weights { int on, denom, w0, w1, o0, o1; };

__device__ __forceinline__ Weights bi_weights(const VVCCudaPB &pb, const VVCCudaWP *wp, int c)
{
    Weights w = { 0, 0, 0, 0, 0, 0 };
    const int lut[5] = { 4, 5, 3, 10, -2 };
    if (pb.bcw_idx) {
        w.on = 1; w.denom = 2; w.w1 = lut[pb.bcw_idx]; w.w0 = 8 - w.w1;
    } else if ((pb.flags & VVC_CUDA_PB_WEIGHTED) && !(pb.flags & VVC_CUDA_PB_DMVR)) {
        const VVCCudaWP e = wp[pb.wp];
        w.on = 1; w.denom = e.log2_denom[c > 0];
        w.w0 = e.weight[0][c]; w.w1 = e.weight[1][c]; w.o0 = e.offset[0][c]; w.o1 = e.offset[1][c];
    }
    return w;
}

__device__ __forceinline__ int combine_bi(int a, int b, const Weights &w, int bd)
{
    if (!w.on) {
        const int shift = max(3, 15 - bd);
        return d_clip_pel((a + b + (1 << (shift - 1))) >> shift, bd);
    }
    const int shift = w.denom + max(3, 15 - bd);
    const int offset = (((w.o0 + w.o1) << (bd - 8)) + 1) << (shift - 1);
    return d_clip_pel((a * w.w0 + b * w.w1 + offset) >> shift, bd);
}

struct UniW { int on, shift, wx, ox; };

__device__ __forceinline__ UniW uni_weights(const VVCCudaPB &pb, const VVCCudaWP *wp, int lx, int c, int bd)
{
    UniW w = { 0, 0, 0, 0 };
    if (pb.flags & VVC_CUDA_PB_WEIGHTED) {
        const VVCCudaWP e = wp[pb.wp];
        w.on = 1; w.shift = e.log2_denom[c > 0] + 14 - bd; w.wx = e.weight[lx][c]; w.ox = e.offset[lx][c] * (1 << (bd - 8));
    }
    return w;
}

__device__ __forceinline__ int finish_uni(int val, const UniW &w, int bd)
{
    if (w.on)
        return d_clip_pel(((val * w.wx + (1 << (w.shift - 1))) >> w.shift) + w.ox, bd);
    return d_clip_pel((val + (1 << (13 - bd))) >> (14 - bd), bd);
}

// ring of integer samples around a tile, bdof_fetch_samples / fetch_samples (vvc_inter_template.c:101-133)
__device__ __forceinline__ void fetch_ring(Smem &s, int ui, int bw, int bh, int bd)
{
    const Unit &u = s.u[ui];
    const int xo = (u.mx >> 3) + 3, yo = (u.my >> 3) + 3;      // window-local position of tile sample (0, 0)
    const int per = 2 * (bw + 2) + 2 * bh;
    for (int idx = threadIdx.x; idx < per; idx += kThreads) {
        int tx, ty;
        if (idx < bw + 2)            { tx = idx - 1;              ty = -1; }
        else if (idx < 2 * (bw + 2)) { tx = idx - (bw + 2) - 1;   ty = bh; }
        else if (idx < 2 * (bw + 2) + bh) { tx = -1;              ty = idx - 2 * (bw + 2); }
        else                         { tx = bw;                   ty = idx - 2 * (bw + 2) - bh; }
        s.tile[ui][(ty + 1) * TP + tx + 1] = (short)(s.win[ui * WIN_LUMA + (ty + yo) * 24 + tx + xo] << (14 - bd));
    }
}

// pad_int16 (vvcdsp.c:29-47): columns first, then rows including the corners
__device__ __forceinline__ void pad_cols(short *t, int pitch, int w, int h)
{
    for (int y = threadIdx.x; y < h; y += kThreads) {
        t[y * pitch - 1] = t[y * pitch];
        t[y * pitch + w] = t[y * pitch + w - 1];
    }
}
__device__ __forceinline__ void pad_rows(short *t, int pitch, int w, int h)
{
    for (int x = threadIdx.x; x < w + 2; x += kThreads) {
        t[-pitch + x - 1] = t[x - 1];
        t[h * pitch + x - 1] = t[(h - 1) * pitch + x - 1];
    }
}

__device__ __forceinline__ int vsign(int v) { return v < 0 ? -1 : (v != 0); }

__device__ int parametric(const int *sd, int stride)
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

__global__ void __launch_bounds__(kThreads) inter_kernel(const InterK p)
{
    __shared__ Smem s;
    const int tid = threadIdx.x, bd = p.bd;

    for (int ri = blockIdx.x; ri < p.n; ri += gridDim.x) {
        __syncthreads();                                   // previous record is done with shared memory
        if (tid < (int)(sizeof(VVCCudaPB) / 4))
            reinterpret_cast<unsigned *>(&s.rec)[tid] = __ldg(reinterpret_cast<const unsigned *>(p.pbs + ri) + tid);
        __syncthreads();
        const VVCCudaPB &pb = s.rec;
        const int w = pb.w, h = pb.h;
        const int lw = 31 - __clz(w);
        const bool gpm = pb.flags & VVC_CUDA_PB_GPM;
        const bool bi = gpm || pb.pred_flag == 3;
        const bool dmvr = (pb.flags & VVC_CUDA_PB_DMVR) != 0;
        if (tid < 4)
            s.mv[tid >> 1][tid & 1] = pb.mv[tid >> 1][tid & 1];
        if (tid == 4)
            s.bdof = (pb.flags & VVC_CUDA_PB_BDOF) ? 1 : 0;

        // ---- DMVR (dmvr_mv_refine, vvc_inter.c:685-748) -----------------------------------------
        if (dmvr && (pb.planes & VVC_CUDA_PB_LUMA)) {
            if (tid < 2) {
                Unit &u = s.u[tid];
                u.plane = p.ref[0] + pb.ref[tid] * p.rb[0]; u.pitch = p.rp[0];
                u.xlo = 0; u.xhi = p.w - 1; u.ylo = 0; u.yhi = p.h - 1;
                u.ox = pb.x0 + (pb.mv[tid][0] >> 4) - 2; u.oy = pb.y0 + (pb.mv[tid][1] >> 4) - 2;
                u.mx = pb.mv[tid][0] & 15; u.my = pb.mv[tid][1] & 15;
            }
            __syncthreads();
            const int pw = w + 4, ph = h + 4;
            for (int idx = tid; idx < 2 * (ph + 1) * 24; idx += kThreads) {       // (pw + 1) x (ph + 1) samples
                const int ui = idx / ((ph + 1) * 24), rem = idx - ui * (ph + 1) * 24;
                const int r = rem / 24, c = rem - r * 24;
                if (c <= pw)
                    s.win[ui * WIN_LUMA + r * 24 + c] = (pel)fetch(s.u[ui], s.u[ui].ox + c, s.u[ui].oy + r);
            }
            __syncthreads();
            for (int idx = tid; idx < 2 * ph * 20; idx += kThreads) {
                const int ui = idx / (ph * 20), rem = idx - ui * ph * 20;
                const int y = rem / 20, x = rem - y * 20;
                if (x >= pw)
                    continue;
                const Unit &u = s.u[ui];
                const pel *q = &s.win[ui * WIN_LUMA + y * 24 + x];
                const int8_t *fx = vvct_dmvr_filters[u.mx], *fy = vvct_dmvr_filters[u.my];
                const int sh1 = bd - 6, off1 = 1 << (sh1 - 1);
                int v;
                if (!u.mx && !u.my)
                    v = bd > 10 ? (q[0] + (1 << (bd - 11))) >> (bd - 10) : q[0] << (10 - bd);
                else if (!u.my)
                    v = (fx[0] * q[0] + fx[1] * q[1] + off1) >> sh1;
                else if (!u.mx)
                    v = (fy[0] * q[0] + fy[1] * q[24] + off1) >> sh1;
                else {
                    const int t0 = (short)((fx[0] * q[0] + fx[1] * q[1] + off1) >> sh1);
                    const int t1 = (short)((fx[0] * q[24] + fx[1] * q[25] + off1) >> sh1);
                    v = (fy[0] * t0 + fy[1] * t1 + 8) >> 4;
                }
                s.dm[ui][y * 20 + x] = (short)v;
            }
            __syncthreads();
            // 25 SADs on every other row (vvc_sad, vvcdsp.c:49-65); one warp per search position
            {
                const int warp = tid >> 5, lane = tid & 31;
                const int cnt = (h >> 1) << lw;
                for (int pos = warp; pos < 25; pos += kThreads / 32) {
                    const int dy = pos / 5 - 2, dx = pos - (pos / 5) * 5 - 2;
                    int acc = 0;
                    for (int e = lane; e < cnt; e += 32) {
                        const int x = e & (w - 1), y = (e >> lw) << 1;
                        acc += abs(s.dm[0][(2 + dy + y) * 20 + 2 + dx + x] - s.dm[1][(2 - dy + y) * 20 + 2 - dx + x]);
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1)
                        acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (!lane)
                        s.sad[pos] = acc;
                }
            }
            __syncthreads();
            if (tid == 0) {
                int min_sad = s.sad[12], min_dx = 2, min_dy = 2;
                min_sad -= min_sad >> 2;
                s.sad[12] = min_sad;
                int mv[2][2] = { { pb.mv[0][0], pb.mv[0][1] }, { pb.mv[1][0], pb.mv[1][1] } };
                if (min_sad >= w * h) {
                    for (int pos = 0; pos < 25; pos++)
                        if (pos != 12 && s.sad[pos] < min_sad) {
                            min_sad = s.sad[pos]; min_dx = pos % 5; min_dy = pos / 5;
                        }
                    int dmv0 = (min_dx - 2) * 16, dmv1 = (min_dy - 2) * 16;
                    if (min_dx != 0 && min_dx != 4 && min_dy != 0 && min_dy != 4) {
                        dmv0 += parametric(&s.sad[min_dy * 5 + min_dx], 1);
                        dmv1 += parametric(&s.sad[min_dy * 5 + min_dx], 5);
                    }
                    for (int i = 0; i < 2; i++) {
                        mv[i][0] = d_clip3(mv[i][0] + (1 - 2 * i) * dmv0, -(1 << 17), (1 << 17) - 1);
                        mv[i][1] = d_clip3(mv[i][1] + (1 - 2 * i) * dmv1, -(1 << 17), (1 << 17) - 1);
                    }
                }
                if (min_sad < 2 * w * h)
                    s.bdof = 0;
                for (int i = 0; i < 2; i++) { s.mv[i][0] = mv[i][0]; s.mv[i][1] = mv[i][1]; }
                if (p.dmvr_out) {
                    VVCCudaDmvrOut o;
                    o.mv[0][0] = mv[0][0]; o.mv[0][1] = mv[0][1]; o.mv[1][0] = mv[1][0]; o.mv[1][1] = mv[1][1];
                    o.min_sad = min_sad; o.bdof_applied = s.bdof;
                    p.dmvr_out[ri] = o;
                }
            }
        }
        __syncthreads();
        const bool do_bdof = s.bdof && !gpm;

        // ---- luma --------------------------------------------------------------------------------
        if (pb.planes & VVC_CUDA_PB_LUMA) {
            if (tid < 2 && (gpm || (pb.pred_flag >> tid & 1))) {
                Unit &u = s.u[tid];
                const int filt = gpm ? 0 : pb.filt;
                u.plane = p.ref[0] + pb.ref[tid] * p.rb[0]; u.pitch = p.rp[0];
                u.xlo = 0; u.xhi = p.w - 1; u.ylo = 0; u.yhi = p.h - 1;
                if (dmvr) {            // emulated_edge_dmvr, vvc_inter.c:60-89
                    const int xs = pb.x0 + (pb.mv[tid][0] >> 4), ys = pb.y0 + (pb.mv[tid][1] >> 4);
                    const int sx = min(max(xs - 3, 0), p.w - 1), sy = min(max(ys - 3, 0), p.h - 1);
                    u.xlo = sx; u.xhi = sx + max(min(p.w, xs + w + 4) - sx, 1) - 1;
                    u.ylo = sy; u.yhi = sy + max(min(p.h, ys + h + 4) - sy, 1) - 1;
                }
                u.ox = pb.x0 + (s.mv[tid][0] >> 4); u.oy = pb.y0 + (s.mv[tid][1] >> 4);
                u.mx = s.mv[tid][0] & 15; u.my = s.mv[tid][1] & 15;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    u.hf[k] = vvct_luma_mc_filters[filt][u.mx][k];
                    u.vf[k] = vvct_luma_mc_filters[filt][u.my][k];
                }
            }
            __syncthreads();
            pel *d = p.dst[0] + pb.pic * p.db[0] + (long long)pb.y0 * p.dp[0] + pb.x0;
            if (!bi) {
                const int lx = pb.pred_flag - 1;
                if (lx == 1 && tid == 0)
                    s.u[0] = s.u[1];                       // single unit lives in slot 0
                __syncthreads();
                stage_windows<8>(s, 1, w, h);
                __syncthreads();
                pass_h<8>(s, 1, w, h, lw, bd);
                __syncthreads();
                const UniW uw = uni_weights(pb, p.wp, lx, 0, bd);
                const bool use_prof = pb.flags & (lx ? VVC_CUDA_PB_PROF1 : VVC_CUDA_PB_PROF0);
                if (use_prof) {                             // luma_prof_uni, vvc_inter.c:368-408
                    for (int idx = tid; idx < h << lw; idx += kThreads) {
                        const int x = idx & (w - 1), y = idx >> lw;
                        s.tile[0][(y + 1) * TP + x + 1] = (short)pass_v<8>(s, 0, x, y, bd);
                    }
                    fetch_ring(s, 0, w, h, bd);
                    __syncthreads();
                    const VVCCudaProf *pr = p.prof + pb.prof;
                    for (int idx = tid; idx < h << lw; idx += kThreads) {
                        const int x = idx & (w - 1), y = idx >> lw;
                        const short *q = &s.tile[0][(y + 1) * TP + x + 1];
                        const int gh = (short)((q[1] >> 6) - (q[-1] >> 6)), gv = (short)((q[TP] >> 6) - (q[-TP] >> 6));
                        const int di = gh * pr->diff_mv_x[lx][y * 4 + x] + gv * pr->diff_mv_y[lx][y * 4 + x];
                        const int limit = 1 << max(13, bd + 1);
                        d[(long long)y * p.dp[0] + x] = (pel)finish_uni(q[0] + d_clip3(di, -limit, limit - 1), uw, bd);
                    }
                } else {
                    for (int idx = tid; idx < h << lw; idx += kThreads) {
                        const int x = idx & (w - 1), y = idx >> lw;
                        d[(long long)y * p.dp[0] + x] = (pel)finish_uni(pass_v<8>(s, 0, x, y, bd), uw, bd);
                    }
                }
            } else {
                stage_windows<8>(s, 2, w, h);
                __syncthreads();
                pass_h<8>(s, 2, w, h, lw, bd);
                __syncthreads();
                for (int idx = tid; idx < 2 * h << lw; idx += kThreads) {
                    const int x = idx & (w - 1), q = idx >> lw;
                    const int ui = q >= h, y = q - ui * h;
                    s.tile[ui][(y + 1) * TP + x + 1] = (short)pass_v<8>(s, ui, x, y, bd);
                }
                const int prof_mask = gpm ? 0 : pb.flags & (VVC_CUDA_PB_PROF0 | VVC_CUDA_PB_PROF1);
                if (do_bdof || prof_mask) {
                    fetch_ring(s, 0, w, h, bd);
                    fetch_ring(s, 1, w, h, bd);
                }
                __syncthreads();
                if (prof_mask) {                            // luma_prof_bi, vvc_inter.c:410-446 (4x4 blocks)
                    int val = 0;
                    const int ui = tid >> 4, e = tid & 15, x = e & 3, y = e >> 2;
                    const bool act = tid < 32 && (prof_mask & (ui ? VVC_CUDA_PB_PROF1 : VVC_CUDA_PB_PROF0));
                    if (act) {
                        const VVCCudaProf *pr = p.prof + pb.prof;
                        const short *q = &s.tile[ui][(y + 1) * TP + x + 1];
                        const int gh = (short)((q[1] >> 6) - (q[-1] >> 6)), gv = (short)((q[TP] >> 6) - (q[-TP] >> 6));
                        const int di = gh * pr->diff_mv_x[ui][e] + gv * pr->diff_mv_y[ui][e];
                        const int limit = 1 << max(13, bd + 1);
                        val = q[0] + d_clip3(di, -limit, limit - 1);
                    }
                    __syncthreads();
                    if (act)
                        s.tile[ui][(y + 1) * TP + x + 1] = (short)val;
                    __syncthreads();
                }
                if (gpm) {                                  // put_gpm, vvc_inter_template.c:78-98
                    const int shift = max(5, 17 - bd);
                    const uint8_t *wt = &vvct_gpm_weights[0][0] + pb.gpm_weights;
                    for (int idx = tid; idx < h << lw; idx += kThreads) {
                        const int x = idx & (w - 1), y = idx >> lw;
                        const int a = s.tile[0][(y + 1) * TP + x + 1], b = s.tile[1][(y + 1) * TP + x + 1];
                        const int g = wt[y * pb.gpm_step_y + x * pb.gpm_step_x];
                        d[(long long)y * p.dp[0] + x] = (pel)d_clip_pel((a * g + b * (8 - g) + (1 << (shift - 1))) >> shift, bd);
                    }
                } else if (do_bdof) {                       // apply_bdof, vvc_inter_template.c:288-317
                    for (int idx = tid; idx < 2 * h << lw; idx += kThreads) {
                        const int x = idx & (w - 1), q = idx >> lw;
                        const int ui = q >= h, y = q - ui * h;
                        const short *t = &s.tile[ui][(y + 1) * TP + x + 1];
                        s.grad[ui][0][(y + 1) * GP + x + 1] = (short)((t[1] >> 6) - (t[-1] >> 6));
                        s.grad[ui][1][(y + 1) * GP + x + 1] = (short)((t[TP] >> 6) - (t[-TP] >> 6));
                    }
                    __syncthreads();
                    for (int a = 0; a < 2; a++) {
                        pad_cols(&s.tile[a][TP + 1], TP, w, h);
                        pad_cols(&s.grad[a][0][GP + 1], GP, w, h);
                        pad_cols(&s.grad[a][1][GP + 1], GP, w, h);
                    }
                    __syncthreads();
                    for (int a = 0; a < 2; a++) {
                        pad_rows(&s.tile[a][TP + 1], TP, w, h);
                        pad_rows(&s.grad[a][0][GP + 1], GP, w, h);
                        pad_rows(&s.grad[a][1][GP + 1], GP, w, h);
                    }
                    __syncthreads();
                    {   // derive_bdof_vx_vy: 8 lanes per 4x4 block sum its 6x6 window
                        const int blk = tid >> 3, sub = tid & 7, nbx = w >> 2, nblk = nbx * (h >> 2);
                        int sgx2 = 0, sgy2 = 0, sgxgy = 0, sgxdi = 0, sgydi = 0;
                        if (blk < nblk) {
                            const int bx = (blk % nbx) << 2, by = (blk / nbx) << 2;
                            for (int e = sub; e < 36; e += 8) {
                                const int y = e / 6, x = e - y * 6;
                                const int ti = (by + y) * TP + bx + x, gi = (by + y) * GP + bx + x;
                                const int diff = (s.tile[0][ti] >> 4) - (s.tile[1][ti] >> 4);
                                const int th = (s.grad[0][0][gi] + s.grad[1][0][gi]) >> 1;
                                const int tv = (s.grad[0][1][gi] + s.grad[1][1][gi]) >> 1;
                                sgx2 += abs(th); sgy2 += abs(tv);
                                sgxgy += vsign(tv) * th;
                                sgxdi -= vsign(th) * diff;
                                sgydi -= vsign(tv) * diff;
                            }
                        }
#pragma unroll
                        for (int o = 4; o; o >>= 1) {
                            sgx2  += __shfl_xor_sync(0xffffffffu, sgx2, o);
                            sgy2  += __shfl_xor_sync(0xffffffffu, sgy2, o);
                            sgxgy += __shfl_xor_sync(0xffffffffu, sgxgy, o);
                            sgxdi += __shfl_xor_sync(0xffffffffu, sgxdi, o);
                            sgydi += __shfl_xor_sync(0xffffffffu, sgydi, o);
                        }
                        if (blk < nblk && !sub) {
                            const int vx = sgx2 > 0 ? d_clip3((sgxdi * 4) >> d_ilog2(sgx2), -15, 15) : 0;
                            const int vy = sgy2 > 0 ? d_clip3(((sgydi * 4) - ((vx * sgxgy) >> 1)) >> d_ilog2(sgy2), -15, 15) : 0;
                            s.vxy[blk][0] = vx; s.vxy[blk][1] = vy;
                        }
                    }
                    __syncthreads();
                    const int shift = 15 - bd;
                    for (int idx = tid; idx < h << lw; idx += kThreads) {
                        const int x = idx & (w - 1), y = idx >> lw;
                        const int blk = (y >> 2) * (w >> 2) + (x >> 2);
                        const int ti = (y + 1) * TP + x + 1, gi = (y + 1) * GP + x + 1;
                        const int off = s.vxy[blk][0] * (s.grad[0][0][gi] - s.grad[1][0][gi]) + s.vxy[blk][1] * (s.grad[0][1][gi] - s.grad[1][1][gi]);
                        d[(long long)y * p.dp[0] + x] = (pel)d_clip_pel((s.tile[0][ti] + (1 << (shift - 1)) + s.tile[1][ti] + off) >> shift, bd);
                    }
                } else {
                    const Weights wt = bi_weights(pb, p.wp, 0);
                    for (int idx = tid; idx < h << lw; idx += kThreads) {
                        const int x = idx & (w - 1), y = idx >> lw;
                        d[(long long)y * p.dp[0] + x] = (pel)combine_bi(s.tile[0][(y + 1) * TP + x + 1], s.tile[1][(y + 1) * TP + x + 1], wt, bd);
                    }
                }
            }
        }

        // ---- chroma (both planes, 4:2:0) ------------------------------------------------------------
        if ((pb.planes & VVC_CUDA_PB_CHROMA) && p.planes == 3) {
            __syncthreads();
            const int bw = w >> 1, bh = h >> 1, lbw = lw - 1;
            const int x0 = pb.x0 >> 1, y0 = pb.y0 >> 1, pw = p.w >> 1, ph = p.h >> 1;
            const int lx = pb.pred_flag - 1;
            if (tid < 4) {
                const int list = bi ? (tid & 1) : lx, pc = tid >> 1;
                Unit &u = s.u[tid];
                u.plane = p.ref[pc + 1] + pb.ref[list] * p.rb[pc + 1]; u.pitch = p.rp[pc + 1];
                u.xlo = 0; u.xhi = pw - 1; u.ylo = 0; u.yhi = ph - 1;
                if (dmvr) {
                    const int xs = x0 + (pb.mv[list][0] >> 5), ys = y0 + (pb.mv[list][1] >> 5);
                    const int sx = min(max(xs - 1, 0), pw - 1), sy = min(max(ys - 1, 0), ph - 1);
                    u.xlo = sx; u.xhi = sx + max(min(pw, xs + bw + 2) - sx, 1) - 1;
                    u.ylo = sy; u.yhi = sy + max(min(ph, ys + bh + 2) - sy, 1) - 1;
                }
                u.ox = x0 + (s.mv[list][0] >> 5); u.oy = y0 + (s.mv[list][1] >> 5);
                u.mx = s.mv[list][0] & 31; u.my = s.mv[list][1] & 31;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    u.hf[k] = vvct_chroma_mc_filters[0][u.mx][k];
                    u.vf[k] = vvct_chroma_mc_filters[0][u.my][k];
                }
            }
            __syncthreads();
            // uni: units 0 (Cb) and 2 (Cr) only -> compact to 0, 1
            if (!bi) {
                if (tid == 0)
                    s.u[1] = s.u[2];
                __syncthreads();
            }
            const int n_units = bi ? 4 : 2;
            stage_windows<4>(s, n_units, bw, bh);
            __syncthreads();
            pass_h<4>(s, n_units, bw, bh, lbw, bd);
            __syncthreads();
            if (!bi) {
                for (int idx = tid; idx < 2 * bh << lbw; idx += kThreads) {
                    const int x = idx & (bw - 1), q = idx >> lbw;
                    const int pc = q >= bh, y = q - pc * bh;
                    const UniW uw = uni_weights(pb, p.wp, lx, pc + 1, bd);
                    pel *d = p.dst[pc + 1] + pb.pic * p.db[pc + 1] + (long long)(y0 + y) * p.dp[pc + 1] + x0 + x;
                    *d = (pel)finish_uni(pass_v<4>(s, pc, x, y, bd), uw, bd);
                }
            } else {
                const int shift = max(5, 17 - bd);
                const uint8_t *wt = &vvct_gpm_weights[0][0] + pb.gpm_weights;
                for (int idx = tid; idx < 2 * bh << lbw; idx += kThreads) {
                    const int x = idx & (bw - 1), q = idx >> lbw;
                    const int pc = q >= bh, y = q - pc * bh;
                    const int a = (short)pass_v<4>(s, pc * 2, x, y, bd), b = (short)pass_v<4>(s, pc * 2 + 1, x, y, bd);
                    pel *d = p.dst[pc + 1] + pb.pic * p.db[pc + 1] + (long long)(y0 + y) * p.dp[pc + 1] + x0 + x;
                    if (gpm) {
                        const int g = wt[y * 2 * pb.gpm_step_y + x * 2 * pb.gpm_step_x];
                        *d = (pel)d_clip_pel((a * g + b * (8 - g) + (1 << (shift - 1))) >> shift, bd);
                    } else {
                        const Weights bw_ = bi_weights(pb, p.wp, pc + 1);
                        *d = (pel)combine_bi(a, b, bw_, bd);
                    }
                }
            }
        }
    }
}

int check(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *refs)
{
    if (!dst || !refs)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "inter: null frame");
    if ((dst->bit_depth != 10 && dst->bit_depth != 12) || refs->bit_depth != dst->bit_depth)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "inter: bit depth %d not accelerated (10/12 only)", dst->bit_depth);
    if (dst->width != refs->width || dst->height != refs->height || dst->chroma_format_idc != refs->chroma_format_idc)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "inter: dst/ref geometry differs (reference scaling is not supported)");
    if (dst->chroma_format_idc && (dst->hshift != 1 || dst->vshift != 1))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "inter: only 4:0:0 and 4:2:0 are accelerated");
    // VVCCudaPB.pic and .ref[] are 8-bit: a larger ring cannot be addressed by the records
    if (dst->batch > 256 || refs->batch > 256)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "inter: rings of more than 256 pictures (%d destination, %d reference) cannot be addressed by the 8-bit picture fields of the records",
                            dst->batch, refs->batch);
    return 0;
}

}  // namespace

extern "C" int vvc_cuda_inter_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *refs,
                                    const VVCCudaPB *pbs, int n_pbs, const VVCCudaWP *wp, const VVCCudaProf *prof,
                                    VVCCudaDmvrOut *dmvr_out)
{
    if (ctx->err)
        return ctx->err;
    if (check(ctx, dst, refs))
        return ctx->err;
    if (n_pbs <= 0)
        return VVC_CUDA_OK;
    if (!pbs)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "inter: null record list");
    InterK p;
    p.planes = dst->chroma_format_idc ? 3 : 1;
    for (int c = 0; c < 3; c++) {
        p.ref[c] = (const pel *)refs->data[c];   p.dst[c] = (pel *)dst->data[c];
        p.rp[c] = (int)(refs->stride[c] / 2);    p.dp[c] = (int)(dst->stride[c] / 2);
        p.rb[c] = refs->batch_stride[c] / 2;     p.db[c] = dst->batch_stride[c] / 2;
    }
    p.w = dst->width; p.h = dst->height; p.bd = dst->bit_depth; p.nref = refs->batch;
    p.margin = ctx->ref_pad;
    p.pbs = pbs; p.n = n_pbs; p.wp = wp; p.prof = prof; p.dmvr_out = dmvr_out;
    if (p.bd == 10 && frame_vec_ok(dst) && frame_vec_ok(refs) && !ctx->force_generic) {
        InterLists lists;
        if (vvc_inter_launch_classify(ctx, p, &lists))
            return ctx->err;
        // A launch of one or two pictures does not fill the machine per task class: its six class kernels (disjoint
        // samples) go to parallel streams, the longest first, so that each one's tail is filled by the others (281 ->
        // 211 us for one 4K picture).  Long launches keep one stream: there the kernels only disturb each other's
        // caches (measured 1385 -> 1533 us for eight pictures).
#ifndef INTER_LONG_SPREAD
#define INTER_LONG_SPREAD 1         // long launches move the small kernels (chroma classes, PROF) to a side stream: inter 2.77 -> 2.65 ms per 16 pictures (tools/sweep_long_spread.sh)
#endif
        const bool spread = p.n < 200000 || INTER_LONG_SPREAD == 3;     // 3 (experiment): every class kernel on its own stream at any size
        if (spread) {
            if (vvc_ctx_fork(ctx, 3))
                return ctx->err;
            if (vvc_inter_launch_warp(ctx, p, lists, 1) || vvc_inter_launch_patch(ctx, p, lists, 1))
                return ctx->err;
            return vvc_ctx_join(ctx, 3) ? ctx->err : VVC_CUDA_OK;
        }
        if (INTER_LONG_SPREAD) {
            if (vvc_ctx_fork(ctx, 3))
                return ctx->err;
            if (vvc_inter_launch_patch(ctx, p, lists, 2) || vvc_inter_launch_warp(ctx, p, lists, 2))
                return ctx->err;
            return vvc_ctx_join(ctx, 3) ? ctx->err : VVC_CUDA_OK;
        }
        if (vvc_inter_launch_patch(ctx, p, lists, 0))
            return ctx->err;
        return vvc_inter_launch_warp(ctx, p, lists, 0);
    }
    const int grid = n_pbs < 148 * 12 ? n_pbs : 148 * 12;
    inter_kernel<<<grid, kThreads, 0, ctx->stream>>>(p);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_inter_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *refs,
                                         const VVCCudaPB *pbs, int n_pbs, const VVCCudaWP *wp, int n_wp,
                                         const VVCCudaProf *prof, int n_prof, VVCCudaDmvrOut *dmvr_out)
{
    if (ctx->err)
        return ctx->err;
    if (!dst || !refs || (n_pbs > 0 && !pbs))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "inter_host: null argument");
    const VVCRefPadOff pad_guard(ctx, true);               // staged host references carry no margins
    const size_t dsz = align_up(vvc_stage_frame_size(dst), 256), rsz = align_up(vvc_stage_frame_size(refs), 256);
    const size_t psz = align_up((size_t)n_pbs * sizeof(VVCCudaPB), 256);
    const size_t wsz = align_up((size_t)(n_wp > 0 ? n_wp : 1) * sizeof(VVCCudaWP), 256);
    const size_t fsz = align_up((size_t)(n_prof > 0 ? n_prof : 1) * sizeof(VVCCudaProf), 256);
    const size_t osz = align_up((size_t)n_pbs * sizeof(VVCCudaDmvrOut), 256);
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, dsz + rsz + psz + wsz + fsz + osz);
    if (!base)
        return ctx->err;
    VVCCudaFrame dd, dr;
    vvc_stage_frame_layout(dst, base, &dd);
    vvc_stage_frame_layout(refs, base + dsz, &dr);
    VVCCudaPB *dpb = (VVCCudaPB *)(base + dsz + rsz);
    VVCCudaWP *dwp = (VVCCudaWP *)(base + dsz + rsz + psz);
    VVCCudaProf *dpf = (VVCCudaProf *)(base + dsz + rsz + psz + wsz);
    VVCCudaDmvrOut *dout = (VVCCudaDmvrOut *)(base + dsz + rsz + psz + wsz + fsz);
    // the destination travels too: samples not covered by a record (intra CUs) keep their content
    if (vvc_stage_frame_h2d(ctx, &dd, dst) || vvc_stage_frame_h2d(ctx, &dr, refs))
        return ctx->err;
    if (n_pbs > 0)
        VVC_TRY(ctx, cudaMemcpyAsync(dpb, pbs, (size_t)n_pbs * sizeof(VVCCudaPB), cudaMemcpyHostToDevice, ctx->stream));
    if (wp && n_wp > 0)
        VVC_TRY(ctx, cudaMemcpyAsync(dwp, wp, (size_t)n_wp * sizeof(VVCCudaWP), cudaMemcpyHostToDevice, ctx->stream));
    if (prof && n_prof > 0)
        VVC_TRY(ctx, cudaMemcpyAsync(dpf, prof, (size_t)n_prof * sizeof(VVCCudaProf), cudaMemcpyHostToDevice, ctx->stream));
    if (dmvr_out && n_pbs > 0)
        VVC_TRY(ctx, cudaMemsetAsync(dout, 0, (size_t)n_pbs * sizeof(VVCCudaDmvrOut), ctx->stream));
    if (vvc_cuda_inter_frame(ctx, &dd, &dr, dpb, n_pbs, dwp, dpf, dmvr_out ? dout : NULL))
        return ctx->err;
    if (vvc_stage_frame_d2h(ctx, dst, &dd))
        return ctx->err;
    if (dmvr_out && n_pbs > 0)
        VVC_TRY(ctx, cudaMemcpyAsync(dmvr_out, dout, (size_t)n_pbs * sizeof(VVCCudaDmvrOut), cudaMemcpyDeviceToHost, ctx->stream));
    return vvc_cuda_sync(ctx);
}
