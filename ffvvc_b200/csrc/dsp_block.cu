// Per-call block kernels behind the drop-in VVCDSPContext table (dsp_table.cu): one launch = one call of one table
// entry on operands staged in device memory.  They exist for signature compatibility and for the reference's own
// checkasm harness; the batched per-picture entries are the fast path.  Every kernel restates one table entry:
//   inter   put / put_uni / put_uni_w           libavcodec/h26x/h2656_inter_template.c:29-577
//           avg, w_avg, put_ciip, put_gpm, fetch_samples, bdof_fetch_samples, prof_grad_filter, apply_prof*,
//           apply_bdof, dmvr*                   libavcodec/vvc/vvc_inter_template.c:25-415;  sad, pad_int16  vvcdsp.c:29-65
//   sao     band / edge                         libavcodec/h26x/h2656_sao_template.c:24-79
//   alf     filter[2], filter_cc, classify, recon_coeff_and_clip   libavcodec/vvc/vvc_filter_template.c:43-408
//   lf      filter_luma[2], filter_chroma[2], ladf_level[2]        libavcodec/vvc/vvc_filter_template.c:466-804,
//                                                                  libavcodec/h26x/h2656_deblock_template.c:25-99
//   itx     add_residual_joint, pred_residual_joint                libavcodec/vvc/vvcdsp_template.c:48-74
// One thread per output sample where samples are independent; one thread per 4x4 block / edge segment where the
// reference's own unit of work is one.  BD = bit depth of the table (10 or 12; pixel = uint16_t).
#include "dsp_block.cuh"

namespace vvcblk {

constexpr int PB = 128;     // MAX_PB_SIZE: pitch of the reference's int16 prediction tiles

__device__ __forceinline__ int clip_pel(int v, int bd) { return min(max(v, 0), (1 << bd) - 1); }
__device__ __forceinline__ int clip3(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---- put / put_uni / put_uni_w ----------------------------------------------------------------------------------------
template <int BD>
__global__ void mc_kernel(const McArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.w || y >= a.h)
        return;
    const pel *s = a.src + (long long)y * a.sstride + x;
    const int taps = a.taps, before = taps / 2 - 1, ss = a.sstride;
    int val;
    if (!a.hfrac && !a.vfrac) {
        val = s[0] << (14 - BD);
    } else if (!a.vfrac) {
        int t = 0;
        for (int k = 0; k < taps; k++) t += a.hf[k] * s[k - before];
        val = t >> (BD - 8);
    } else if (!a.hfrac) {
        int t = 0;
        for (int k = 0; k < taps; k++) t += a.vf[k] * s[(k - before) * ss];
        val = t >> (BD - 8);
    } else {
        int acc = 0;
        for (int j = 0; j < taps; j++) {
            int t = 0;
            for (int k = 0; k < taps; k++) t += a.hf[k] * s[(j - before) * ss + k - before];
            acc += a.vf[j] * (int16_t)(t >> (BD - 8));          // the first stage is stored in an int16_t array
        }
        val = acc >> 6;
    }
    if (a.mode == 0) {
        a.dst16[y * PB + x] = (int16_t)val;
    } else if (a.mode == 1) {
        if (!a.hfrac && !a.vfrac)
            a.dst[(long long)y * a.dstride + x] = s[0];             // put_uni_pixels is a copy
        else
            a.dst[(long long)y * a.dstride + x] = (pel)clip_pel((val + (1 << (13 - BD))) >> (14 - BD), BD);
    } else {
        const int shift = a.denom + 14 - BD, ox = a.ox * (1 << (BD - 8));
        a.dst[(long long)y * a.dstride + x] = (pel)clip_pel(((val * a.wx + (1 << (shift - 1))) >> shift) + ox, BD);
    }
}

// ---- avg / w_avg / put_gpm / put_ciip ---------------------------------------------------------------------------------
template <int BD>
__global__ void blend_kernel(const BlendArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.w || y >= a.h)
        return;
    pel *d = a.dst + (long long)y * a.dstride + x;
    if (a.mode == 3) {                                            // put_ciip: dst holds the intra prediction
        *d = (pel)((*d * a.w0 + a.inter[(long long)y * a.istride + x] * (4 - a.w0) + 2) >> 2);
        return;
    }
    const int s0 = a.src0[y * PB + x], s1 = a.src1[y * PB + x];
    if (a.mode == 0) {
        const int shift = 15 - BD > 3 ? 15 - BD : 3;
        *d = (pel)clip_pel((s0 + s1 + (1 << (shift - 1))) >> shift, BD);
    } else if (a.mode == 1) {
        const int shift = a.denom + (15 - BD > 3 ? 15 - BD : 3);
        const int offset = (((a.o0 + a.o1) << (BD - 8)) + 1) << (shift - 1);
        *d = (pel)clip_pel((s0 * a.w0 + s1 * a.w1 + offset) >> shift, BD);
    } else {
        const int shift = 17 - BD > 5 ? 17 - BD : 5;
        const int g = a.weights[y * a.step_y + x * a.step_x];
        *d = (pel)clip_pel((s0 * g + s1 * (8 - g) + (1 << (shift - 1))) >> shift, BD);
    }
}

// ---- dmvr[my != 0][mx != 0] --------------------------------------------------------------------------------------------
template <int BD>
__global__ void dmvr_kernel(int16_t *dst, const pel *src, int ss, int h, int w, int mx, int my)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h)
        return;
    const pel *s = src + (long long)y * ss + x;
    const int fx[2] = { 16 - mx, mx }, fy[2] = { 16 - my, my };      // ff_vvc_inter_luma_dmvr_filters (vvc_data.c)
    const int shift1 = BD - 6, off1 = 1 << (shift1 - 1);
    int v;
    if (!mx && !my)      v = BD > 10 ? (s[0] + (1 << (BD - 11))) >> (BD - 10) : s[0] << (10 - BD);
    else if (!my)        v = (fx[0] * s[0] + fx[1] * s[1] + off1) >> shift1;
    else if (!mx)        v = (fy[0] * s[0] + fy[1] * s[ss] + off1) >> shift1;
    else {
        const int t0 = (int16_t)((fx[0] * s[0]  + fx[1] * s[1]      + off1) >> shift1);
        const int t1 = (int16_t)((fx[0] * s[ss] + fx[1] * s[ss + 1] + off1) >> shift1);
        v = (fy[0] * t0 + fy[1] * t1 + 8) >> 4;
    }
    dst[y * PB + x] = (int16_t)v;
}

// ---- sad ---------------------------------------------------------------------------------------------------------------
__global__ void sad_kernel(int *out, const int16_t *src0, const int16_t *src1, int dx, int dy, int bw, int bh)
{
    dx -= 2; dy -= 2;
    const int16_t *a = src0 + (2 + dy) * PB + 2 + dx, *b = src1 + (2 - dy) * PB + 2 - dx;
    int s = 0;
    for (int i = threadIdx.x; i < (bh / 2) * bw; i += 32) {
        const int y = (i / bw) * 2, x = i % bw;
        s += abs(a[y * PB + x] - b[y * PB + x]);
    }
    for (int o = 16; o; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0)
        *out = s;
}

// ---- fetch_samples / bdof_fetch_samples: the ring of integer samples around a tile ----------------------------------------
template <int BD>
__global__ void fetch_kernel(int16_t *dst, const pel *src, int ss, int w, int h)
{
    // dst: tile origin (the ring lies at x = -1, w and y = -1, h); src: the sample under ring position (-1, -1)
    for (int i = threadIdx.x; i < (w + 2) * (h + 2); i += blockDim.x) {
        const int y = i / (w + 2), x = i % (w + 2);
        if (y == 0 || y == h + 1 || x == 0 || x == w + 1)
            dst[(y - 1) * PB + x - 1] = (int16_t)(src[(long long)y * ss + x] << (14 - BD));
    }
}

// ---- prof_grad_filter (pad: BDOF's replicated border, pad_int16) ---------------------------------------------------------
__global__ void prof_grad_kernel(int16_t *gh, int16_t *gv, int gs, const int16_t *src, int ss, int w, int h, int pad)
{
    int16_t *oh = gh + pad * (1 + gs), *ov = gv + pad * (1 + gs);
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
        const int y = i / w, x = i % w;
        const int16_t *p = src + y * ss + x;
        oh[y * gs + x] = (int16_t)((p[1] >> 6) - (p[-1] >> 6));
        ov[y * gs + x] = (int16_t)((p[ss] >> 6) - (p[-ss] >> 6));
    }
    if (!pad)
        return;
    __syncthreads();
    for (int g = 0; g < 2; g++) {
        int16_t *t = (g ? gv : gh) + 1 + gs;
        for (int y = threadIdx.x; y < h; y += blockDim.x) {
            t[y * gs - 1] = t[y * gs];
            t[y * gs + w] = t[y * gs + w - 1];
        }
        __syncthreads();
        for (int x = threadIdx.x; x < w + 2; x += blockDim.x) {
            t[x - 1 - gs] = t[x - 1];
            t[x - 1 + h * gs] = t[x - 1 + (h - 1) * gs];
        }
        __syncthreads();
    }
}

// ---- apply_prof / apply_prof_uni / apply_prof_uni_w (4x4) ------------------------------------------------------------------
template <int BD>
__global__ void prof_kernel(const ProfArgs a)
{
    const int x = threadIdx.x & 3, y = threadIdx.x >> 2, o = threadIdx.x;
    const int16_t *p = a.src + y * PB + x;
    const int limit = 1 << (BD + 1 > 13 ? BD + 1 : 13);
    const int g_h = (int16_t)((p[1] >> 6) - (p[-1] >> 6)), g_v = (int16_t)((p[PB] >> 6) - (p[-PB] >> 6));
    const int di = g_h * a.dx[o] + g_v * a.dy[o];
    const int val = p[0] + clip3(di, -limit, limit - 1);
    if (a.mode == 0) {
        a.dst16[y * PB + x] = (int16_t)val;
    } else if (a.mode == 1) {
        a.dst[(long long)y * a.dstride + x] = (pel)clip_pel((val + (1 << (13 - BD))) >> (14 - BD), BD);
    } else {
        const int shift = a.denom + (14 - BD > 2 ? 14 - BD : 2), ox = a.ox * (1 << (BD - 8));
        a.dst[(long long)y * a.dstride + x] = (pel)clip_pel(((val * a.wx + (1 << (shift - 1))) >> shift) + ox, BD);
    }
}

// ---- apply_bdof: one CTA per call, thread per 4x4 block for the decision, per sample for gradients and output ----------------
__device__ __forceinline__ int vsign(int v) { return v < 0 ? -1 : (v != 0); }

template <int BD>
__global__ void bdof_kernel(pel *dst, int dstride, int16_t *s0, int16_t *s1, int w, int h)
{
    constexpr int GP = 18;
    __shared__ int16_t gh[2][GP * GP], gv[2][GP * GP];
    __shared__ int vxy[16][2];
    int16_t *s[2] = { s0, s1 };
    const int tid = threadIdx.x;
    for (int i = 0; i < 2; i++) {
        for (int q = tid; q < w * h; q += blockDim.x) {
            const int y = q / w, x = q % w;
            const int16_t *p = s[i] + y * PB + x;
            gh[i][(y + 1) * GP + x + 1] = (int16_t)((p[1] >> 6) - (p[-1] >> 6));
            gv[i][(y + 1) * GP + x + 1] = (int16_t)((p[PB] >> 6) - (p[-PB] >> 6));
        }
    }
    __syncthreads();
    // pad_int16 of the gradients and of the source tiles themselves (:299-302): sides, then top / bottom rows
    for (int i = 0; i < 2; i++) {
        for (int y = tid; y < h; y += blockDim.x) {
            gh[i][(y + 1) * GP] = gh[i][(y + 1) * GP + 1];  gh[i][(y + 1) * GP + w + 1] = gh[i][(y + 1) * GP + w];
            gv[i][(y + 1) * GP] = gv[i][(y + 1) * GP + 1];  gv[i][(y + 1) * GP + w + 1] = gv[i][(y + 1) * GP + w];
            s[i][y * PB - 1] = s[i][y * PB];                s[i][y * PB + w] = s[i][y * PB + w - 1];
        }
    }
    __syncthreads();
    for (int i = 0; i < 2; i++) {
        for (int x = tid; x < w + 2; x += blockDim.x) {
            gh[i][x] = gh[i][GP + x];  gh[i][(h + 1) * GP + x] = gh[i][h * GP + x];
            gv[i][x] = gv[i][GP + x];  gv[i][(h + 1) * GP + x] = gv[i][h * GP + x];
            s[i][x - 1 - PB] = s[i][x - 1];  s[i][x - 1 + h * PB] = s[i][x - 1 + (h - 1) * PB];
        }
    }
    __syncthreads();
    const int nbx = w >> 2;
    if (tid < nbx * (h >> 2)) {
        const int bx = (tid % nbx) * 4, by = (tid / nbx) * 4;
        int sgx2 = 0, sgy2 = 0, sgxgy = 0, sgxdi = 0, sgydi = 0;
        for (int y = 0; y < 6; y++)
            for (int x = 0; x < 6; x++) {
                const int ti = (by + y - 1) * PB + bx + x - 1, gi = (by + y) * GP + bx + x;
                const int diff = (s0[ti] >> 4) - (s1[ti] >> 4);
                const int th = (gh[0][gi] + gh[1][gi]) >> 1, tv = (gv[0][gi] + gv[1][gi]) >> 1;
                sgx2 += abs(th); sgy2 += abs(tv);
                sgxgy += vsign(tv) * th; sgxdi -= vsign(th) * diff; sgydi -= vsign(tv) * diff;
            }
        const int vx = sgx2 > 0 ? clip3((sgxdi * 4) >> (31 - __clz(sgx2)), -15, 15) : 0;
        const int vy = sgy2 > 0 ? clip3(((sgydi * 4) - ((vx * sgxgy) >> 1)) >> (31 - __clz(sgy2)), -15, 15) : 0;
        vxy[tid][0] = vx; vxy[tid][1] = vy;
    }
    __syncthreads();
    for (int q = tid; q < w * h; q += blockDim.x) {
        const int y = q / w, x = q % w, b = (y >> 2) * nbx + (x >> 2);
        const int ti = y * PB + x, gi = (y + 1) * GP + x + 1;
        const int off = vxy[b][0] * (gh[0][gi] - gh[1][gi]) + vxy[b][1] * (gv[0][gi] - gv[1][gi]);
        dst[(long long)y * dstride + x] = (pel)clip_pel((s0[ti] + (1 << (14 - BD)) + s1[ti] + off) >> (15 - BD), BD);
    }
}

// ---- SAO band / edge --------------------------------------------------------------------------------------------------------
template <int BD>
__global__ void sao_kernel(const SaoArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.w || y >= a.h)
        return;
    const pel *s = a.src + (long long)y * a.sstride + x;
    const int v = s[0];
    int off;
    if (a.edge) {
        const int dx[4][2] = { { -1, 1 }, { 0, 0 }, { -1, 1 }, { 1, -1 } }, dy[4][2] = { { 0, 0 }, { -1, 1 }, { -1, 1 }, { -1, 1 } };
        const int idx[5] = { 1, 2, 0, 3, 4 };
        const int na = s[dy[a.eo][0] * a.sstride + dx[a.eo][0]], nb = s[dy[a.eo][1] * a.sstride + dx[a.eo][1]];
        off = a.offset_val[idx[2 + ((v > na) - (v < na)) + ((v > nb) - (v < nb))]];
    } else {
        const int k = ((v >> (BD - 5)) - a.left_class) & 31;
        off = k < 4 ? a.offset_val[k + 1] : 0;
    }
    a.dst[(long long)y * a.dstride + x] = (pel)clip_pel(v + off, BD);
}

// ---- ALF ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int vb_reach(int k, int t, int span)
{
    if (t < 0 && t >= -span) return min(k, -t - 1);
    if (t >= 0 && t < span)  return min(k, t);
    return k;
}

// filter[LUMA / CHROMA]: per 4x4 block its own 12 (6: one set for the whole call) coefficients and clips
template <int BD>
__global__ void alf_filter_kernel(const AlfArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.w || y >= a.h)
        return;
    const int ss = a.sstride;
    const pel *p = a.src + (long long)y * ss + x;
    const int t = y - a.vb_pos, span = a.chroma ? 2 : 4;
    const int d1 = vb_reach(1, t, span) * ss, d2 = vb_reach(2, t, span) * ss, d3 = vb_reach(3, t, span) * ss;
    const int cur = p[0];
    int sum = 0;
#define PAIR(k, u, v) sum += f[k] * (clip3((int)(u) - cur, -c[k], c[k]) + clip3((int)(v) - cur, -c[k], c[k]))
    if (a.chroma) {
        const int16_t *f = a.filter, *c = a.clip;
        PAIR(0, p[d2], p[-d2]);  PAIR(1, p[d1 + 1], p[-d1 - 1]);  PAIR(2, p[d1], p[-d1]);  PAIR(3, p[d1 - 1], p[-d1 + 1]);
        PAIR(4, p[2], p[-2]);    PAIR(5, p[1], p[-1]);
    } else {
        const int blk = (y >> 2) * (a.w >> 2) + (x >> 2);          // the reference advances filter / clip per 4x4 block, row-major
        const int16_t *f = a.filter + blk * 12, *c = a.clip + blk * 12;
        PAIR(0, p[d3], p[-d3]);
        PAIR(1, p[d2 + 1], p[-d2 - 1]);  PAIR(2, p[d2], p[-d2]);  PAIR(3, p[d2 - 1], p[-d2 + 1]);
        PAIR(4, p[d1 + 2], p[-d1 - 2]);  PAIR(5, p[d1 + 1], p[-d1 - 1]);  PAIR(6, p[d1], p[-d1]);
        PAIR(7, p[d1 - 1], p[-d1 + 1]);  PAIR(8, p[d1 - 2], p[-d1 + 2]);
        PAIR(9, p[3], p[-3]);  PAIR(10, p[2], p[-2]);  PAIR(11, p[1], p[-1]);
    }
#undef PAIR
    sum = (t == -1 || t == 0) ? (sum + 512) >> 10 : (sum + 64) >> 7;
    a.dst[(long long)y * a.dstride + x] = (pel)clip_pel(cur + sum, BD);
}

template <int BD>
__global__ void alf_cc_kernel(pel *dst, int dstride, const pel *luma, int ls, int w, int h, int hs, int vs, const int16_t *f, int vb_pos)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h)
        return;
    const int pos = y << vs;
    if (!vs && (pos == vb_pos || pos == vb_pos + 1))
        return;
    const pel *src = luma + (long long)pos * ls + (x << hs);
    int up = -ls, dn = ls, dn2 = 2 * ls;
    if (pos == vb_pos - 2 || pos == vb_pos + 1)   dn2 = ls;
    else if (pos == vb_pos - 1 || pos == vb_pos)  up = dn = dn2 = 0;
    const int cur = src[0];
    int sum = f[0] * (src[up] - cur) + f[1] * (src[-1] - cur) + f[2] * (src[1] - cur) + f[3] * (src[dn - 1] - cur) +
              f[4] * (src[dn] - cur) + f[5] * (src[dn + 1] - cur) + f[6] * (src[dn2] - cur);
    sum = clip3((sum + 64) >> 7, -(1 << (BD - 1)), (1 << (BD - 1)) - 1);
    pel *d = dst + (long long)y * dstride + x;
    *d = (pel)clip_pel(*d + sum, BD);
}

// classify: thread per 4x4 block (:299-381)
template <int BD>
__global__ void alf_classify_kernel(int *class_idx, int *transpose_idx, const pel *src, int ss, int w, int h, int vb)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x, nbx = w >> 2;
    if (b >= nbx * (h >> 2))
        return;
    const int bx = (b % nbx) * 4, by = (b / nbx) * 4;
    const uint8_t act[16] = { 0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4 };
    int first = 0, last = 4, scale = 2;
    if (by + 4 == vb)  { last = 3;  scale = 3; }
    else if (by == vb) { first = 1; scale = 3; }
    int sv = 0, sh = 0, sd0 = 0, sd1 = 0;
#define T(xx, yy) ((int)src[(long long)(yy) * ss + (xx)])
    for (int i = first; i < last; i++) {
        const int yy = by + 2 * i;
        int r0 = yy - 3, r1 = yy - 2, r2 = yy - 1, r3 = yy;
        if (yy == vb)          r3 = r2;
        else if (yy == vb + 2) r0 = r1;
        for (int j = 0; j < 4; j++) {
            const int xa = bx + 2 * j - 2, xb = xa + 1;
            const int c0 = 2 * T(xa, r1), c1 = 2 * T(xb, r2);
            sv  += abs(c0 - T(xa, r0) - T(xa, r2))         + abs(c1 - T(xb, r1) - T(xb, r3));
            sh  += abs(c0 - T(xa - 1, r1) - T(xa + 1, r1)) + abs(c1 - T(xb - 1, r2) - T(xb + 1, r2));
            sd0 += abs(c0 - T(xa - 1, r0) - T(xa + 1, r2)) + abs(c1 - T(xb - 1, r1) - T(xb + 1, r3));
            sd1 += abs(c0 - T(xa + 1, r0) - T(xa - 1, r2)) + abs(c1 - T(xb + 1, r1) - T(xb - 1, r3));
        }
    }
#undef T
    const int v_le_h = sv <= sh, d0_le_d1 = sd0 <= sd1;
    const int hv_hi = max(sv, sh), hv_lo = min(sv, sh), d_hi = max(sd0, sd1), d_lo = min(sd0, sd1);
    const int hv_wins = (unsigned long long)d_hi * (unsigned)hv_lo <= (unsigned long long)hv_hi * (unsigned)d_lo;
    const int hi = hv_wins ? hv_hi : d_hi, lo = hv_wins ? hv_lo : d_lo;
    int c = act[min(max(((sh + sv) * scale) >> (BD - 1), 0), 15)];
    if (hi * 2 > 9 * lo)  c += (2 * hv_wins + 2) * 5;
    else if (hi > 2 * lo) c += (2 * hv_wins + 1) * 5;
    class_idx[b] = c;
    transpose_idx[b] = d0_le_d1 * 2 + v_le_h;
}

template <int BD>
__global__ void alf_recon_kernel(int16_t *coeff, int16_t *clip, const int *class_idx, const int *transpose_idx, int size,
                                 const int16_t *coeff_set, const uint8_t *clip_idx_set, const uint8_t *class_to_filt)
{
    const uint8_t perm[4][12] = { { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 }, { 9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6 },
                                  { 0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11 }, { 9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6 } };
    const uint8_t shift[4] = { 0, 3, 5, 7 };
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size * 12)
        return;
    const int blk = i / 12, j = i % 12, cls = class_idx[blk], idx = perm[transpose_idx[blk]][j];
    coeff[i] = coeff_set[class_to_filt[cls] * 12 + idx];
    clip[i] = (int16_t)(1 << (BD - shift[clip_idx_set[cls * 12 + idx]]));
}

// ---- deblocking: thread per edge segment (4 luma lines; 4 or 2 chroma lines) ---------------------------------------------------
#define AT(base, k) ((base)[(long long)(k) * xs])
__device__ __forceinline__ int curv(const pel *l, long long xs, int a, int b, int c) { return abs((int)AT(l, a) - 2 * (int)AT(l, b) + (int)AT(l, c)); }

// long filters of 3 / 5 / 7 samples (:497-541): interpolation weights towards the far reference sample and clip factors
__device__ const uint8_t k_long_w[3][7] = { { 53, 32, 11 }, { 58, 45, 32, 19, 6 }, { 59, 50, 41, 32, 23, 14, 5 } };
__device__ const uint8_t k_long_k[3][7] = { { 6, 4, 2 }, { 6, 5, 4, 3, 2 }, { 6, 5, 4, 3, 2, 1, 1 } };

__device__ void luma_long(pel *pix, long long xs, long long ys, int tc, int lp, int lq, int no_p, int no_q)
{
    for (int line = 0; line < 4; line++, pix += ys) {
        int p[8], q[8], m;
        for (int i = 0; i < 8; i++) {                       // P(max_len) is the farthest sample the filter reads on a side
            p[i] = i <= lp ? (int)AT(pix, -1 - i) : 0;
            q[i] = i <= lq ? (int)AT(pix, i) : 0;
        }
        if (lp == 5 && lq == 5)   m = (p[4] + p[3] + 2 * (p[2] + p[1] + p[0] + q[0] + q[1] + q[2]) + q[3] + q[4] + 8) >> 4;
        else if (lp == lq)        m = (p[6] + p[5] + p[4] + p[3] + p[2] + p[1] + 2 * (p[0] + q[0]) + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + 8) >> 4;
        else if (lp + lq == 12)   m = (p[5] + p[4] + p[3] + p[2] + 2 * (p[1] + p[0] + q[0] + q[1]) + q[2] + q[3] + q[4] + q[5] + 8) >> 4;
        else if (lp + lq == 8)    m = (p[3] + p[2] + p[1] + p[0] + q[0] + q[1] + q[2] + q[3] + 4) >> 3;
        else if (lq == 7)         m = (2 * (p[2] + p[1] + p[0] + q[0]) + p[0] + p[1] + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + 8) >> 4;
        else                      m = (p[6] + p[5] + p[4] + p[3] + p[2] + p[1] + 2 * (q[2] + q[1] + q[0] + p[0]) + q[0] + q[1] + 8) >> 4;
        if (!no_p) {
            const int ref = (p[lp] + p[lp - 1] + 1) >> 1;
            for (int i = 0; i < lp; i++) {
                const int wt = k_long_w[(lp >> 1) - 1][i], kk = k_long_k[(lp >> 1) - 1][i];
                const int lim = (tc * kk) >> 1;
                AT(pix, -1 - i) = (pel)(p[i] + clip3(((m * wt + ref * (64 - wt) + 32) >> 6) - p[i], -lim, lim));
            }
        }
        if (!no_q) {
            const int ref = (q[lq] + q[lq - 1] + 1) >> 1;
            for (int i = 0; i < lq; i++) {
                const int wt = k_long_w[(lq >> 1) - 1][i], kk = k_long_k[(lq >> 1) - 1][i];
                const int lim = (tc * kk) >> 1;
                AT(pix, i) = (pel)(q[i] + clip3(((m * wt + ref * (64 - wt) + 32) >> 6) - q[i], -lim, lim));
            }
        }
    }
}

template <int BD>
__global__ void lf_luma_kernel(const LfArgs a)
{
    const int seg = threadIdx.x;
    if (seg >= 2)
        return;
    const long long xs = a.xs, ys = a.ys;
    pel *pix = a.pix + seg * 4 * ys;
    const int tc = a.tc[seg] << (BD - 10), beta = a.beta[seg] << (BD - 8);
    const int no_p = a.no_p[seg], no_q = a.no_q[seg];
    int lp = a.max_len_p[seg], lq = a.max_len_q[seg];
    if (!tc)
        return;
    pel *l0 = pix, *l3 = pix + 3 * ys;
    const int dp0 = curv(l0, xs, -3, -2, -1), dq0 = curv(l0, xs, 2, 1, 0);
    const int dp3 = curv(l3, xs, -3, -2, -1), dq3 = curv(l3, xs, 2, 1, 0);
    const int d0 = dp0 + dq0, d3 = dp3 + dq3;
    const int tc25 = (tc * 5 + 1) >> 1;
    const int big_p = lp > 3 && !a.param, big_q = lq > 3;          // param: hor_ctu_edge
#define AD(l, u, v) abs((int)AT(l, u) - (int)AT(l, v))
    if (big_p || big_q) {
        const int dp0l = big_p ? (dp0 + curv(l0, xs, -6, -5, -4) + 1) >> 1 : dp0;
        const int dq0l = big_q ? (dq0 + curv(l0, xs, 5, 4, 3) + 1) >> 1 : dq0;
        const int dp3l = big_p ? (dp3 + curv(l3, xs, -6, -5, -4) + 1) >> 1 : dp3;
        const int dq3l = big_q ? (dq3 + curv(l3, xs, 5, 4, 3) + 1) >> 1 : dq3;
        const int d0l = dp0l + dq0l, d3l = dp3l + dq3l;
        lp = big_p ? lp : 3;
        lq = big_q ? lq : 3;
        if (d0l + d3l < beta) {
            const int b53 = (beta * 3) >> 5, b4 = beta >> 4;
            int sp0 = AD(l0, -4, -1) + (lp == 7 ? abs((int)AT(l0, -8) - AT(l0, -7) - AT(l0, -6) + AT(l0, -5)) : 0);
            int sq0 = AD(l0, 0, 3)   + (lq == 7 ? abs((int)AT(l0, 4) - AT(l0, 5) - AT(l0, 6) + AT(l0, 7)) : 0);
            int sp3 = AD(l3, -4, -1) + (lp == 7 ? abs((int)AT(l3, -8) - AT(l3, -7) - AT(l3, -6) + AT(l3, -5)) : 0);
            int sq3 = AD(l3, 0, 3)   + (lq == 7 ? abs((int)AT(l3, 4) - AT(l3, 5) - AT(l3, 6) + AT(l3, 7)) : 0);
            if (big_p) { sp0 = (sp0 + AD(l0, -4, -1 - lp) + 1) >> 1; sp3 = (sp3 + AD(l3, -4, -1 - lp) + 1) >> 1; }
            if (big_q) { sq0 = (sq0 + AD(l0, 3, lq) + 1) >> 1;       sq3 = (sq3 + AD(l3, 3, lq) + 1) >> 1; }
            if (sp0 + sq0 < b53 && AD(l0, -1, 0) < tc25 && sp3 + sq3 < b53 && AD(l3, -1, 0) < tc25 && (d0l << 1) < b4 && (d3l << 1) < b4) {
                luma_long(pix, xs, ys, tc, lp, lq, no_p, no_q);
                return;
            }
        }
    }
    if (d0 + d3 >= beta)
        return;
    if (lp > 2 && lq > 2 && AD(l0, -4, -1) + AD(l0, 3, 0) < (beta >> 3) && AD(l0, -1, 0) < tc25 &&
        AD(l3, -4, -1) + AD(l3, 3, 0) < (beta >> 3) && AD(l3, -1, 0) < tc25 && (d0 << 1) < (beta >> 2) && (d3 << 1) < (beta >> 2)) {
        for (int line = 0; line < 4; line++, pix += ys) {
            const int p3 = AT(pix, -4), p2 = AT(pix, -3), p1 = AT(pix, -2), p0 = AT(pix, -1);
            const int q0 = AT(pix, 0), q1 = AT(pix, 1), q2 = AT(pix, 2), q3 = AT(pix, 3);
            if (!no_p) {
                AT(pix, -1) = (pel)(p0 + clip3(((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3) - p0, -3 * tc, 3 * tc));
                AT(pix, -2) = (pel)(p1 + clip3(((p2 + p1 + p0 + q0 + 2) >> 2) - p1, -2 * tc, 2 * tc));
                AT(pix, -3) = (pel)(p2 + clip3(((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3) - p2, -tc, tc));
            }
            if (!no_q) {
                AT(pix, 0) = (pel)(q0 + clip3(((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3) - q0, -3 * tc, 3 * tc));
                AT(pix, 1) = (pel)(q1 + clip3(((p0 + q0 + q1 + q2 + 2) >> 2) - q1, -2 * tc, 2 * tc));
                AT(pix, 2) = (pel)(q2 + clip3(((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3) - q2, -tc, tc));
            }
        }
    } else {
        int np = 1, nq = 1;
        if (lp > 1 && lq > 1) {
            const int side = (beta + (beta >> 1)) >> 3;
            if (dp0 + dp3 < side) np = 2;
            if (dq0 + dq3 < side) nq = 2;
        }
        const int half = tc >> 1;
        for (int line = 0; line < 4; line++, pix += ys) {
            const int p2 = AT(pix, -3), p1 = AT(pix, -2), p0 = AT(pix, -1), q0 = AT(pix, 0), q1 = AT(pix, 1), q2 = AT(pix, 2);
            int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
            if (abs(delta) >= 10 * tc)
                continue;
            delta = clip3(delta, -tc, tc);
            if (!no_p) AT(pix, -1) = (pel)clip_pel(p0 + delta, BD);
            if (!no_q) AT(pix, 0)  = (pel)clip_pel(q0 - delta, BD);
            if (!no_p && np > 1) AT(pix, -2) = (pel)clip_pel(p1 + clip3((((p2 + p0 + 1) >> 1) - p1 + delta) >> 1, -half, half), BD);
            if (!no_q && nq > 1) AT(pix, 1)  = (pel)clip_pel(q1 + clip3((((q2 + q0 + 1) >> 1) - q1 - delta) >> 1, -half, half), BD);
        }
    }
}

template <int BD>
__global__ void lf_chroma_kernel(const LfArgs a)
{
    const int lines = a.param ? 2 : 4, seg = threadIdx.x;         // param: shift (the edge direction is subsampled)
    if (seg >= 8 / lines)
        return;
    const long long xs = a.xs, ys = a.ys;
    pel *pix = a.pix + seg * lines * ys;
    const int tc = a.tc[seg] << (BD - 10), beta = a.beta[seg] << (BD - 8);
    const int no_p = a.no_p[seg], no_q = a.no_q[seg];
    int lp = a.max_len_p[seg], lq = a.max_len_q[seg];
    if (!tc || !lp || !lq)
        return;
    if (lq == 3) {
        const pel *l0 = pix, *l1 = pix + (lines == 2 ? 1 : 3) * ys;
        const int tc25 = (tc * 5 + 1) >> 1, one = lp == 1;
        const int p0 = AT(l0, -1), p1 = AT(l0, -2), p2 = one ? p1 : AT(l0, -3), p3 = one ? p1 : AT(l0, -4);
        const int n0 = AT(l1, -1), n1 = AT(l1, -2), n2 = one ? n1 : AT(l1, -3), n3 = one ? n1 : AT(l1, -4);
        const int d0 = abs(p2 - 2 * p1 + p0) + curv(l0, xs, 2, 1, 0), d1 = abs(n2 - 2 * n1 + n0) + curv(l1, xs, 2, 1, 0);
        bool strong = false;
        if (d0 + d1 < beta) {
            const bool ok0 = (d0 << 1) < (beta >> 2) && abs(p3 - p0) + AD(l0, 0, 3) < (beta >> 3) && abs(p0 - (int)AT(l0, 0)) < tc25;
            const bool ok1 = (d1 << 1) < (beta >> 2) && abs(n3 - n0) + AD(l1, 0, 3) < (beta >> 3) && abs(n0 - (int)AT(l1, 0)) < tc25;
            strong = ok0 && ok1;
        }
        if (!strong)
            lp = lq = 1;
    }
    for (int line = 0; line < lines; line++, pix += ys) {
        const int p3 = AT(pix, -4), p2 = AT(pix, -3), p1 = AT(pix, -2), p0 = AT(pix, -1);
        const int q0 = AT(pix, 0), q1 = AT(pix, 1), q2 = AT(pix, 2), q3 = AT(pix, 3);
        if (lp == 3 && lq == 3) {
            if (!no_p) {
                AT(pix, -1) = (pel)clip3((p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3, p0 - tc, p0 + tc);
                AT(pix, -2) = (pel)clip3((2 * p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3, p1 - tc, p1 + tc);
                AT(pix, -3) = (pel)clip3((3 * p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3, p2 - tc, p2 + tc);
            }
            if (!no_q) {
                AT(pix, 0) = (pel)clip3((p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3, q0 - tc, q0 + tc);
                AT(pix, 1) = (pel)clip3((p1 + p0 + q0 + 2 * q1 + q2 + 2 * q3 + 4) >> 3, q1 - tc, q1 + tc);
                AT(pix, 2) = (pel)clip3((p0 + q0 + q1 + 2 * q2 + 3 * q3 + 4) >> 3, q2 - tc, q2 + tc);
            }
        } else if (lq == 3) {
            if (!no_p)
                AT(pix, -1) = (pel)clip3((3 * p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3, p0 - tc, p0 + tc);
            if (!no_q) {
                AT(pix, 0) = (pel)clip3((2 * p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3, q0 - tc, q0 + tc);
                AT(pix, 1) = (pel)clip3((p1 + p0 + q0 + 2 * q1 + q2 + 2 * q3 + 4) >> 3, q1 - tc, q1 + tc);
                AT(pix, 2) = (pel)clip3((p0 + q0 + q1 + 2 * q2 + 3 * q3 + 4) >> 3, q2 - tc, q2 + tc);
            }
        } else {
            const int delta = clip3((((q0 - p0) * 4) + p1 - q1 + 4) >> 3, -tc, tc);
            if (!no_p) AT(pix, -1) = (pel)clip_pel(p0 + delta, BD);
            if (!no_q) AT(pix, 0)  = (pel)clip_pel(q0 - delta, BD);
        }
    }
}
#undef AD
#undef AT

__global__ void ladf_kernel(int *out, const pel *four)      // P0, P0 three lines on, Q0, Q0 three lines on
{
    *out = ((int)four[0] + four[1] + four[2] + four[3]) >> 2;
}

// ---- add_residual_joint / pred_residual_joint (vvcdsp_template.c:48-74) ---------------------------------------------------------
template <int BD>
__global__ void residual_joint_kernel(pel *dst, int dstride, int *res, int w, int h, int c_sign, int shift, int to_buffer)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h)
        return;
    const int r = (res[y * w + x] * c_sign) >> shift;
    if (to_buffer)
        res[y * w + x] = r;
    else
        dst[(long long)y * dstride + x] = (pel)clip_pel(dst[(long long)y * dstride + x] + r, BD);
}

// ---- launchers (all on the given stream; BD selects the instantiation) ------------------------------------------------------------
#define BY_BD(bd, call10, call12) do { if ((bd) == 12) { call12; } else { call10; } } while (0)
static dim3 grid2d(int w, int h) { return dim3((unsigned)((w + 63) / 64), (unsigned)h); }

void mc(cudaStream_t st, int bd, const McArgs &a)        { BY_BD(bd, (mc_kernel<10><<<grid2d(a.w, a.h), 64, 0, st>>>(a)), (mc_kernel<12><<<grid2d(a.w, a.h), 64, 0, st>>>(a))); }
void blend(cudaStream_t st, int bd, const BlendArgs &a)  { BY_BD(bd, (blend_kernel<10><<<grid2d(a.w, a.h), 64, 0, st>>>(a)), (blend_kernel<12><<<grid2d(a.w, a.h), 64, 0, st>>>(a))); }
void dmvr(cudaStream_t st, int bd, int16_t *dst, const pel *src, int ss, int h, int w, int mx, int my)
{
    BY_BD(bd, (dmvr_kernel<10><<<grid2d(w, h), 64, 0, st>>>(dst, src, ss, h, w, mx, my)),
              (dmvr_kernel<12><<<grid2d(w, h), 64, 0, st>>>(dst, src, ss, h, w, mx, my)));
}
void sad(cudaStream_t st, int *out, const int16_t *s0, const int16_t *s1, int dx, int dy, int bw, int bh) { sad_kernel<<<1, 32, 0, st>>>(out, s0, s1, dx, dy, bw, bh); }
void fetch(cudaStream_t st, int bd, int16_t *dst, const pel *src, int ss, int w, int h)
{
    BY_BD(bd, (fetch_kernel<10><<<1, 128, 0, st>>>(dst, src, ss, w, h)), (fetch_kernel<12><<<1, 128, 0, st>>>(dst, src, ss, w, h)));
}
void prof_grad(cudaStream_t st, int16_t *gh, int16_t *gv, int gs, const int16_t *src, int ss, int w, int h, int pad) { prof_grad_kernel<<<1, 256, 0, st>>>(gh, gv, gs, src, ss, w, h, pad); }
void prof(cudaStream_t st, int bd, const ProfArgs &a)    { BY_BD(bd, (prof_kernel<10><<<1, 16, 0, st>>>(a)), (prof_kernel<12><<<1, 16, 0, st>>>(a))); }
void bdof(cudaStream_t st, int bd, pel *dst, int dstride, int16_t *s0, int16_t *s1, int w, int h)
{
    BY_BD(bd, (bdof_kernel<10><<<1, 256, 0, st>>>(dst, dstride, s0, s1, w, h)), (bdof_kernel<12><<<1, 256, 0, st>>>(dst, dstride, s0, s1, w, h)));
}
void sao(cudaStream_t st, int bd, const SaoArgs &a)      { BY_BD(bd, (sao_kernel<10><<<grid2d(a.w, a.h), 64, 0, st>>>(a)), (sao_kernel<12><<<grid2d(a.w, a.h), 64, 0, st>>>(a))); }
void alf_filter(cudaStream_t st, int bd, const AlfArgs &a) { BY_BD(bd, (alf_filter_kernel<10><<<grid2d(a.w, a.h), 64, 0, st>>>(a)), (alf_filter_kernel<12><<<grid2d(a.w, a.h), 64, 0, st>>>(a))); }
void alf_cc(cudaStream_t st, int bd, pel *dst, int dstride, const pel *luma, int ls, int w, int h, int hs, int vs, const int16_t *f, int vb_pos)
{
    BY_BD(bd, (alf_cc_kernel<10><<<grid2d(w, h), 64, 0, st>>>(dst, dstride, luma, ls, w, h, hs, vs, f, vb_pos)),
              (alf_cc_kernel<12><<<grid2d(w, h), 64, 0, st>>>(dst, dstride, luma, ls, w, h, hs, vs, f, vb_pos)));
}
void alf_classify(cudaStream_t st, int bd, int *cls, int *tr, const pel *src, int ss, int w, int h, int vb)
{
    const int n = (w >> 2) * (h >> 2);
    BY_BD(bd, (alf_classify_kernel<10><<<(n + 63) / 64, 64, 0, st>>>(cls, tr, src, ss, w, h, vb)),
              (alf_classify_kernel<12><<<(n + 63) / 64, 64, 0, st>>>(cls, tr, src, ss, w, h, vb)));
}
void alf_recon(cudaStream_t st, int bd, int16_t *coeff, int16_t *clip, const int *cls, const int *tr, int size,
               const int16_t *coeff_set, const uint8_t *clip_idx_set, const uint8_t *class_to_filt)
{
    const int n = size * 12;
    BY_BD(bd, (alf_recon_kernel<10><<<(n + 127) / 128, 128, 0, st>>>(coeff, clip, cls, tr, size, coeff_set, clip_idx_set, class_to_filt)),
              (alf_recon_kernel<12><<<(n + 127) / 128, 128, 0, st>>>(coeff, clip, cls, tr, size, coeff_set, clip_idx_set, class_to_filt)));
}
void lf_luma(cudaStream_t st, int bd, const LfArgs &a)   { BY_BD(bd, (lf_luma_kernel<10><<<1, 32, 0, st>>>(a)), (lf_luma_kernel<12><<<1, 32, 0, st>>>(a))); }
void lf_chroma(cudaStream_t st, int bd, const LfArgs &a) { BY_BD(bd, (lf_chroma_kernel<10><<<1, 32, 0, st>>>(a)), (lf_chroma_kernel<12><<<1, 32, 0, st>>>(a))); }
void ladf(cudaStream_t st, int *out, const pel *four) { ladf_kernel<<<1, 1, 0, st>>>(out, four); }
void residual_joint(cudaStream_t st, int bd, pel *dst, int dstride, int *res, int w, int h, int c_sign, int shift, int to_buffer)
{
    BY_BD(bd, (residual_joint_kernel<10><<<grid2d(w, h), 64, 0, st>>>(dst, dstride, res, w, h, c_sign, shift, to_buffer)),
              (residual_joint_kernel<12><<<grid2d(w, h), 64, 0, st>>>(dst, dstride, res, w, h, c_sign, shift, to_buffer)));
}

}  // namespace vvcblk
