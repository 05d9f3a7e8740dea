// Generic residual block code shared by itx.cu (the stage's leftover / 12-bit kernel) and intra.cu (the residual step of
// the dependency-driven all-intra kernel): one transform block by a group of NT threads (a warp or a whole CTA) out of two
// shared-memory buffers.  Included INSIDE the including file's anonymous namespace, after common.cuh, coeff_src.cuh and
// tables.cuh.  Reference: see itx.cu.
#pragma once

struct ItxK {
    pel       *plane[3];
    int        pitch[3];
    long long  bstride[3];
    CoefSrc    src;
    int32_t   *store;                    // DENSE32 buffer for VVC_CUDA_TB_STORE_RESIDUAL blocks (NULL otherwise)
    const VVCCudaTB *tbs;
    int        n_tbs, range, bd;
    const uint32_t *list, *list_count;   // optional: process tbs[list[0 .. *list_count)] instead of tbs[0 .. n_tbs)
};

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

// inputs the reference's 1-D transform reads for a declared nz (zero-out guards G2..G16, vvc_itx_1d.c:64-67)
__device__ __forceinline__ int inputs_read(int type, int n, int nz)
{
    if (type != 0)
        return nz;
    const int r = nz <= 2 ? 2 : nz <= 4 ? 4 : nz <= 8 ? 8 : nz <= 16 ? 16 : 32;
    return min(r, min(n, 32));
}

template <int NT> __device__ __forceinline__ void group_sync()
{
    if (NT == 32) __syncwarp();
    else          __syncthreads();
}

// One 1-D pass over `lines` lines.  VERT: line = column, transform along rows.  out[i] = sum_j in[j] * M[j][i].
// post_shift < 0: mid-stage rounding (x + 64) >> 7 with clip to the transform range; else (x + rnd) >> post_shift.
template <int NT, bool VERT>
__device__ __forceinline__ void run_pass(const int *src, int *dst, int pitch, int type, int n, int nz, int lines,
                                         int post_shift, int range, int t)
{
    const int8_t *M = tx_matrix(type, n);
    const int rd = inputs_read(type, n, nz);
    if (n >= 4) {
        const int strips = n >> 2;
        for (int item = t; item < lines * strips; item += NT) {
            const int line = item % lines, o0 = (item / lines) << 2;
            int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            const int *in = VERT ? src + line : src + line * pitch;
            const int istep = VERT ? pitch : 1;
            for (int j = 0; j < rd; j++) {
                const int v = in[j * istep];
                const int m4 = __ldg(reinterpret_cast<const int *>(M + j * n + o0));
                a0 += v * (int)(int8_t)(m4);
                a1 += v * (int)(int8_t)(m4 >> 8);
                a2 += v * (int)(int8_t)(m4 >> 16);
                a3 += v * (m4 >> 24);
            }
            int r[4] = { a0, a1, a2, a3 };
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int v = post_shift < 0 ? d_clip_sbits((r[q] + 64) >> 7, range)
                                             : (r[q] + (1 << (post_shift - 1))) >> post_shift;
                if (VERT) dst[(o0 + q) * pitch + line] = v;
                else      dst[line * pitch + o0 + q] = v;
            }
        }
    } else {        // n == 2
        for (int item = t; item < lines * n; item += NT) {
            const int line = item % lines, o = item / lines;
            const int *in = VERT ? src + line : src + line * pitch;
            const int istep = VERT ? pitch : 1;
            int acc = 0;
            for (int j = 0; j < rd; j++)
                acc += in[j * istep] * (int)M[j * n + o];
            const int v = post_shift < 0 ? d_clip_sbits((acc + 64) >> 7, range)
                                         : (acc + (1 << (post_shift - 1))) >> post_shift;
            if (VERT) dst[o * pitch + line] = v;
            else      dst[line * pitch + o] = v;
        }
    }
}

__constant__ uint8_t c_diag4_x[16] = { 0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 1, 2, 3, 2, 3, 3 };
__constant__ uint8_t c_diag4_y[16] = { 0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 3, 2, 1, 3, 2, 3 };

template <int NT, int MODE>
__device__ void process_tb(const ItxK &p, const VVCCudaTB &tb, int ti, int *sC, int *sM, int t)
{
    const int w = 1 << tb.log2_w, h = 1 << tb.log2_h, pitch = w + 1;
    const int flags = tb.flags;
    const bool ts = flags & VVC_CUDA_TB_TS;
    const bool pcm = flags & (VVC_CUDA_TB_BDPCM | VVC_CUDA_TB_BDPCM_VERT);
    int nzw = tb.nzw, nzh = tb.nzh;
    const int lf_side = (w >= 8 && h >= 8) ? 8 : 4;
    const TbCoef tc = tb_coef<MODE>(p.src, ti, tb.coeff_offset, tb.log2_w, tb.log2_h, nzw, nzh, ts);

    // ---- load the window of coefficients that will be read ----
    int LR, LC;
    if (ts || pcm)            { LR = h; LC = w; }
    else if (tb.lfnst)        { LR = LC = lf_side; }
    else if (w > 1 && h > 1)  { LR = inputs_read(tb.trv, h, nzh); LC = nzw; }
    else if (w > 1)           { LR = 1; LC = inputs_read(tb.trh, w, nzw); }
    else                      { LR = inputs_read(tb.trv, h, nzh); LC = 1; }
    for (int i = t; i < LR * LC; i += NT) {
        const int y = i / LC, x = i - y * LC;
        // BDPCM accumulates quantised levels; they are dequantised afterwards (vvc_intra.c:453-455)
        sC[y * pitch + x] = pcm ? coef_raw<MODE>(tc, y, x) : coef_load<MODE>(tc, y, x);
    }
    group_sync<NT>();

    // ---- BDPCM accumulate with clipping (sequential along the accumulation direction) ----
    if (pcm) {
        if (flags & VVC_CUDA_TB_BDPCM_VERT) {
            for (int x = t; x < w; x += NT)
                for (int y = 1; y < h; y++)
                    sC[y * pitch + x] = d_clip_sbits(sC[y * pitch + x] + sC[(y - 1) * pitch + x], p.range);
        } else {
            for (int y = t; y < h; y += NT)
                for (int x = 1; x < w; x++)
                    sC[y * pitch + x] = d_clip_sbits(sC[y * pitch + x] + sC[y * pitch + x - 1], p.range);
        }
        group_sync<NT>();
        if (MODE & 2) {
            for (int i = t; i < w * h; i += NT) {
                const int y = i / w, x = i - y * w;
                sC[y * pitch + x] = coef_dequant<MODE>(tc, sC[y * pitch + x], y, x);
            }
            group_sync<NT>();
        }
    }

    if (!ts) {
        // ---- inverse LFNST: 8/16 inputs in 4x4 diagonal order -> 16/48 outputs ----
        if (tb.lfnst) {
            const int idx = tb.lfnst & 3, set = (tb.lfnst >> 2) & 3;
            const bool transpose = (tb.lfnst >> 4) & 1;
            const int n_in = ((tb.lfnst >> 5) & 1) ? 8 : 16;
            const int n_out = lf_side == 8 ? 48 : 16;
            const int8_t *M = lf_side == 8 ? &vvct_lfnst_8x8[set][idx - 1][0][0] : &vvct_lfnst_4x4[set][idx - 1][0][0];
            int v[2];
            int cnt = 0;
            for (int j = t; j < n_out; j += NT, cnt++) {
                int acc = 0;
                for (int i = 0; i < n_in; i++)
                    acc += sC[c_diag4_y[i] * pitch + c_diag4_x[i]] * (int)M[i * n_out + j];
                v[cnt] = d_clip_sbits((acc + 64) >> 7, p.range);
            }
            group_sync<NT>();
            cnt = 0;
            for (int j = t; j < n_out; j += NT, cnt++) {
                const int r = j < 4 * lf_side ? j / lf_side : 4 + ((j - 4 * lf_side) >> 2);
                const int q = j < 4 * lf_side ? j % lf_side : (j - 4 * lf_side) & 3;
                if (transpose) sC[q * pitch + r] = v[cnt];
                else           sC[r * pitch + q] = v[cnt];
            }
            nzw = nzh = lf_side;
            group_sync<NT>();
        }
        // ---- inverse transform ----
        if (tb.trh == 0 && tb.trv == 0 && nzw == 1 && nzh == 1 && (w == h || w == 1 || h == 1)) {
            // DC-only shortcut of the DCT2 x DCT2 cells (vvcdsp.c:101-108, :125-131)
            const int c0 = sC[0];
            int dc;
            if (w > 1 && h > 1) {
                const int s2 = 5 + p.range - p.bd;
                dc = ((((c0 * 64 + 64) >> 7) * 64) + (1 << (s2 - 1))) >> s2;
            } else {
                const int s = 6 + p.range - p.bd;
                dc = (c0 * 64 + (1 << (s - 1))) >> s;
            }
            group_sync<NT>();
            for (int i = t; i < w * h; i += NT) {
                const int y = i / w, x = i - y * w;
                sC[y * pitch + x] = dc;
            }
        } else if (w > 1 && h > 1) {
            run_pass<NT, true>(sC, sM, pitch, tb.trv, h, nzh, nzw, -1, p.range, t);
            const int rd_h = inputs_read(tb.trh, w, nzw);
            for (int i = t; i < h * (rd_h - nzw); i += NT) {     // columns >= nzw are zero (scale_clip memset)
                const int y = i / (rd_h - nzw), x = nzw + i - y * (rd_h - nzw);
                sM[y * pitch + x] = 0;
            }
            group_sync<NT>();
            run_pass<NT, false>(sM, sC, pitch, tb.trh, w, nzw, h, 5 + p.range - p.bd, p.range, t);
        } else {
            if (w > 1) run_pass<NT, false>(sC, sM, pitch, tb.trh, w, nzw, 1, 6 + p.range - p.bd, p.range, t);
            else       run_pass<NT, true>(sC, sM, pitch, tb.trv, h, nzh, 1, 6 + p.range - p.bd, p.range, t);
            group_sync<NT>();
            for (int i = t; i < w * h; i += NT) {
                const int y = i / w, x = i - y * w;
                sC[y * pitch + x] = sM[y * pitch + x];
            }
        }
        group_sync<NT>();
    }

    // ---- epilogue ----
    if (flags & VVC_CUDA_TB_STORE_RESIDUAL) {
        int32_t *out = p.store + tb.coeff_offset;
        for (int i = t; i < w * h; i += NT) {
            const int y = i / w, x = i - y * w;
            out[i] = sC[y * pitch + x];
        }
    } else {
        const int planes = (flags & VVC_CUDA_TB_JOINT) ? 2 : 1;
        const int cscale = tb_chroma_scale(p.src, tb.chroma_scale);
        // LMCS chroma residual scaling sits between the transform and add_residual (itransform, vvc_intra.c:468-475; the
        // second plane of a joint block is derived first and scaled afterwards, :179-183)
        auto res = [&](int r, int sign, int shift) -> int {
            const int v = (r * sign) >> shift;
            return cscale ? d_lmcs_scale(v, cscale, p.bd) : v;
        };
        for (int pl = 0; pl < planes; pl++) {
            const int c = pl ? tb.joint_c_idx : tb.c_idx;
            pel *base = p.plane[c] + tb.pic * p.bstride[c] + (long long)tb.y0 * p.pitch[c] + tb.x0;
            const int sign = pl ? tb.joint_sign : 1, shift = pl ? tb.joint_shift : 0;
            if (w >= 4 && !(tb.x0 & 3)) {
                const int q4 = w >> 2;
                for (int i = t; i < h * q4; i += NT) {
                    const int y = i / q4, x = (i - y * q4) << 2;
                    uint2 *d = reinterpret_cast<uint2 *>(base + (long long)y * p.pitch[c] + x);
                    const uint2 cur = *d;
                    const int *r = &sC[y * pitch + x];
                    const int o0 = d_clip_pel((int)(cur.x & 0xffff) + res(r[0], sign, shift), p.bd);
                    const int o1 = d_clip_pel((int)(cur.x >> 16)    + res(r[1], sign, shift), p.bd);
                    const int o2 = d_clip_pel((int)(cur.y & 0xffff) + res(r[2], sign, shift), p.bd);
                    const int o3 = d_clip_pel((int)(cur.y >> 16)    + res(r[3], sign, shift), p.bd);
                    *d = make_uint2(o0 | (o1 << 16), o2 | (o3 << 16));
                }
            } else {
                for (int i = t; i < h * w; i += NT) {
                    const int y = i / w, x = i - y * w;
                    pel *d = base + (long long)y * p.pitch[c] + x;
                    *d = (pel)d_clip_pel(*d + res(sC[y * pitch + x], sign, shift), p.bd);
                }
            }
        }
    }
    group_sync<NT>();
}

