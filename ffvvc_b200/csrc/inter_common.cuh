// Shared between the two inter-prediction kernels (inter.cu: generic CTA-per-record kernel for any
// bit depth / alignment; inter_warp.cu: the 10-bit warp-per-record kernel the batched entry uses).
#pragma once
#include "common.cuh"

struct InterK {
    const pel *ref[3];
    pel       *dst[3];
    int        rp[3], dp[3];
    long long  rb[3], db[3];
    int        w, h, bd, planes, nref;      // nref: pictures in the reference ring
    int        margin;                      // replicated luma samples around every reference plane (VVC_CUDA_OPT_REF_PAD; chroma: half)
    const VVCCudaPB   *pbs;
    int                n;
    const VVCCudaWP   *wp;
    const VVCCudaProf *prof;
    VVCCudaDmvrOut    *dmvr_out;
};

// Work lists of the 10-bit path, built on the device by inter_classify_kernel (inter_patch.cu), one set per
// launch in the context's scratch slot 2.  A patch task = record << 3 | patch.  The uni lists grow from the
// front of their array, the bi lists from the back (a record is one or the other, so they never meet).
struct InterLists {
    uint32_t *luma[3]; int cap_luma[3];  // patches of 4 columns x 8 rows, one list per record width 4 / 8 / 16 (at most 2 / 4 / 8 tasks per
                                         // record): the tasks of one row of patches are consecutive and aligned, they share a staged window
    uint32_t *chroma;  int cap_chroma;   // 4 tasks per record at most: (plane, patch column)
    uint32_t *luma_b, *chroma_b;         // the same for records whose reference windows may leave the picture (clamped loads)
    uint32_t *coop;                      // records of the warp-per-record kernels: DMVR / BDOF from the front, PROF from the back
    uint32_t *count;                     // [2] chroma uni, [3] chroma bi, [4] DMVR / BDOF, [5] PROF, [8], [9], [14] work counters of the warp
                                         // kernels, [10..13] border lists (luma uni, bi, chroma uni, bi), [16..18] luma uni by width class,
                                         // [20..22] luma bi by width class
    uint32_t *tail;                      // first word behind the lists (refined vectors of the split DMVR kernels)
};

// 10-bit path (bd == 10, 4:2:0 or 4:0:0, 16-byte aligned planes / pitches): classify + patch kernel
// (inter_patch.cu), then the warp-per-record kernel over the cooperative records (inter_warp.cu)
// classify runs on the context stream; the six task-class kernels write disjoint samples and are spread over the
// context stream and its side streams (vvc_ctx_fork / vvc_ctx_join around the two calls below)
int vvc_inter_launch_classify(VVCCudaCtx *ctx, const InterK &p, InterLists *lists);
int vvc_inter_launch_patch(VVCCudaCtx *ctx, const InterK &p, const InterLists &lists, int spread);
int vvc_inter_launch_warp(VVCCudaCtx *ctx, const InterK &p, const InterLists &lists, int spread);

// Records whose prediction needs a cooperative tile (DMVR search, BDOF windows, PROF gradients) go to the
// warp-per-record kernel; everything else (plain uni / bi, BCW, explicit weights, GPM) to the
// thread-per-patch kernel.  Both kernels are launched over the whole record list and skip the other's share.
#define VVC_PB_COOPERATIVE (VVC_CUDA_PB_DMVR | VVC_CUDA_PB_BDOF | VVC_CUDA_PB_PROF0 | VVC_CUDA_PB_PROF1)

#ifdef __CUDACC__
struct Rec {
    int x0, y0, w, h, planes, pred, ref[2], pic, flags;
    int mv[2][2];
    int filt, bcw, wp, prof, gsx, gsy, gw;
};

__device__ __forceinline__ Rec load_rec(const VVCCudaPB *pb)
{
    const uint32_t *q = reinterpret_cast<const uint32_t *>(pb);
    uint32_t r[11];
#pragma unroll
    for (int i = 0; i < 11; i++)
        r[i] = __ldg(q + i);
    Rec o;
    o.x0 = r[0] & 0xffff;         o.y0 = r[0] >> 16;
    o.w = r[1] & 0xff;            o.h = (r[1] >> 8) & 0xff;   o.planes = (r[1] >> 16) & 0xff;  o.pred = r[1] >> 24;
    o.ref[0] = r[2] & 0xff;       o.ref[1] = (r[2] >> 8) & 0xff; o.pic = (r[2] >> 16) & 0xff;  o.flags = r[2] >> 24;
    o.mv[0][0] = (int)r[3];       o.mv[0][1] = (int)r[4];     o.mv[1][0] = (int)r[5];          o.mv[1][1] = (int)r[6];
    o.filt = r[7] & 0xff;         o.bcw = (r[7] >> 8) & 0xff; o.wp = r[7] >> 16;
    o.prof = r[8] & 0xffff;       o.gsx = (short)(r[8] >> 16);
    o.gsy = (short)(r[9] & 0xffff);
    o.gw = (int)r[10];
    return o;
}

__device__ __forceinline__ uint32_t frc(uint32_t lo, uint32_t hi, int sh) { return __funnelshift_rc(lo, hi, sh); }
// The same instructions as CUDA's __dp2a_lo / __dp2a_hi / __funnelshift_rc, which the headers declare `asm volatile`: pure
// functions of their operands, so the compiler may interleave independent chains, hoist and merge them (INTER_PLAIN_ASM).
__device__ __forceinline__ int nv_dp2a_lo(int a, int b, int c) { int d; asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int nv_dp2a_hi(int a, int b, int c) { int d; asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t nv_frc(uint32_t lo, uint32_t hi, int sh) { uint32_t d; asm("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sh)); return d; }
__device__ __forceinline__ int lo16(uint32_t v) { return (short)(v & 0xffff); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)v >> 16; }
__device__ __forceinline__ uint32_t pack16(int a, int b) { return (uint32_t)(a & 0xffff) | ((uint32_t)b << 16); }
// ---- final roundings (avg / w_avg / put_uni / put_uni_w, vvc_inter_template.c:25-57, h2656_inter_template.c) ----
struct Weights { int on, denom, w0, w1, o0, o1; };

__device__ __forceinline__ Weights bi_weights(const Rec &pb, const VVCCudaWP *wp, int c)
{
    Weights w = { 0, 0, 0, 0, 0, 0 };
    if (pb.bcw) {
        const int w1 = pb.bcw == 1 ? 5 : pb.bcw == 2 ? 3 : pb.bcw == 3 ? 10 : -2;      // {4,5,3,10,-2}[bcw_idx]
        w.on = 1; w.denom = 2; w.w1 = w1; w.w0 = 8 - w1;
    } else if ((pb.flags & VVC_CUDA_PB_WEIGHTED) && !(pb.flags & VVC_CUDA_PB_DMVR)) {
        const VVCCudaWP *e = wp + pb.wp;
        w.on = 1; w.denom = e->log2_denom[c > 0];
        w.w0 = e->weight[0][c]; w.w1 = e->weight[1][c]; w.o0 = e->offset[0][c]; w.o1 = e->offset[1][c];
    }
    return w;
}

__device__ __forceinline__ int combine_bi(int a, int b, const Weights &w)
{
    if (!w.on)
        return d_clip_pel((a + b + 16) >> 5, 10);
    const int shift = w.denom + 5;
    const int offset = (((w.o0 + w.o1) << 2) + 1) << (shift - 1);
    return d_clip_pel((a * w.w0 + b * w.w1 + offset) >> shift, 10);
}

struct UniW { int on, shift, wx, ox; };

__device__ __forceinline__ UniW uni_weights(const Rec &pb, const VVCCudaWP *wp, int lx, int c)
{
    UniW w = { 0, 0, 0, 0 };
    if (pb.flags & VVC_CUDA_PB_WEIGHTED) {
        const VVCCudaWP *e = wp + pb.wp;
        w.on = 1; w.shift = e->log2_denom[c > 0] + 4; w.wx = e->weight[lx][c]; w.ox = e->offset[lx][c] * 4;
    }
    return w;
}

__device__ __forceinline__ int finish_uni(int val, const UniW &w)
{
    if (w.on)
        return d_clip_pel(((val * w.wx + (1 << (w.shift - 1))) >> w.shift) + w.ox, 10);
    return d_clip_pel((val + 8) >> 4, 10);
}

#endif  // __CUDACC__
