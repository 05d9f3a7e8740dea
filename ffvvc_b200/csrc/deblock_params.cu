// Deblocking parameters on the device for sm_100a: boundary strengths, maximum filter lengths and QP -> beta / tc
// (with the luma-adaptive offset) of one direction, written as the edge maps deblock.cu consumes.
//
// Replaces, of libavcodec/vvc/vvc_filter.c: boundary_strength :308-370, derive_max_filter_length_luma :373-397,
// vvc_deblock_subblock_bs_vertical / _horizontal :399-470, vvc_deblock_bs_luma_* :472-634, vvc_deblock_bs_chroma_*
// :636-754, vvc_deblock_bs :756-781, max_filter_length_chroma :793-812, TC_CALC :823-826, get_qp_y / get_qp_c :829-852 and
// the per-edge part of ff_vvc_deblock_vertical / _horizontal :861-1003; and lf.ladf_level (vvc_filter_template.c:788-804).
//
// B200 design: the reference walks transform blocks and, inside each, the 4x4 units of its left / upper edge and of the
// 8x8 sub-block grid.  Here the parser's lists (transform units, motion rectangles) are scattered into two per-4x4 tables
// in HBM (24 + 16 bytes per unit, a warp per record), and then ONE thread per 4x4 unit decides what kind of edge its
// left / upper side is - transform edge, sub-block edge or none: the two kinds never coincide - from its own and its
// neighbour's table entries and writes the finished (tc, beta, lengths) entries of all three planes.  No ordering between
// units is needed, every map entry is written exactly once (zero where there is no edge), and the only picture samples
// read are the four LADF samples per luma segment.
#include "common.cuh"

namespace {

struct __align__(8) DUnit {          // one 4x4 luma unit: what fc->tab.* holds for it
    uint16_t tbx[2], tby[2];         // transform block origin per tree (0xffff: none)
    uint8_t  tbw[2], tbh[2];         // its luma size (log2)
    uint8_t  fl_y;                   // written by luma records: VVC_CUDA_DBK_CBF_Y | BDPCM_Y, bit 6: merge-subblock / affine coding unit
    uint8_t  fl_c;                   // written by chroma records: VVC_CUDA_DBK_CBF_CB | CBF_CR | JOINT | BDPCM_C
    int8_t   qp[3];
    uint8_t  cbw, cbh;               // coding block size (log2)
    uint8_t  pred;                   // PF_*, bit 7: ciip
    uint16_t cbx, cby;
};
static_assert(sizeof(DUnit) == 24, "DUnit layout");

struct __align__(16) DMotion { int32_t mv[2][2]; };
struct DRef { int16_t ref[2]; };

struct ParamsK {
    DUnit   *units;                  // [batch][uh][uw]
    DMotion *motion;
    DRef    *refs;
    int      uw, uh, w, h, batch, planes, hs, vs, ctb_log2, ctb_cols, ctb_rows, bd;
    const pel *luma;  int lpitch;  long long lbstride;
    const VVCCudaDbkCtb *ctbs;
    VVCCudaDbkParams prm;
    VVCCudaDbkEdge *map[3];
    int      mpitch[3];
    long long msize[3];
};

__constant__ uint16_t c_tc[66] = {
      0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,
      0,   0,   3,   4,   4,   4,   4,   5,   5,   5,   5,   7,   7,   8,   9,  10,
     10,  11,  13,  14,  15,  17,  19,  21,  24,  25,  29,  33,  36,  41,  45,  51,
     57,  64,  71,  80,  89, 100, 112, 125, 141, 157, 177, 198, 222, 250, 280, 314,
    352, 395,
};
__constant__ uint8_t c_beta[64] = {
      0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,
      6,   7,   8,   9,  10,  11,  12,  13,  14,  15,  16,  17,  18,  20,  22,  24,
     26,  28,  30,  32,  34,  36,  38,  40,  42,  44,  46,  48,  50,  52,  54,  56,
     58,  60,  62,  64,  66,  68,  70,  72,  74,  76,  78,  80,  82,  84,  86,  88,
};

constexpr int kThreads = 256;
constexpr int kSb = 64;              // DUnit.fl_y: sub-block coding unit

// a warp per transform unit record
__global__ void __launch_bounds__(kThreads) scatter_tu_kernel(const ParamsK p, const VVCCudaDbkTU *tus, int n)
{
    const int lane = threadIdx.x & 31;
    for (int i = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); i < n; i += gridDim.x * (kThreads / 32)) {
        const VVCCudaDbkTU t = tus[i];
        const int uw4 = 1 << (t.log2_w - 2), uh4 = 1 << (t.log2_h - 2);
        DUnit *base = p.units + (size_t)t.pic * p.uw * p.uh;
        for (int j = lane; j < uw4 * uh4; j += 32) {
            const int ux = (t.x0 >> 2) + (j & (uw4 - 1)), uy = (t.y0 >> 2) + (j >> (t.log2_w - 2));
            if (ux >= p.uw || uy >= p.uh)
                continue;
            DUnit *u = base + (size_t)uy * p.uw + ux;
            // luma and chroma trees may come as separate records: each writes its own fields only
            if (t.planes & VVC_CUDA_DBK_TU_LUMA) {
                u->tbx[0] = t.x0; u->tby[0] = t.y0; u->tbw[0] = t.log2_w; u->tbh[0] = t.log2_h;
                u->qp[0] = t.qp[0];
                u->fl_y = (uint8_t)((t.flags & (VVC_CUDA_DBK_CBF_Y | VVC_CUDA_DBK_BDPCM_Y)) | ((t.cu_flags & VVC_CUDA_DBK_CU_SUBBLOCK) ? kSb : 0));
                u->cbw = t.cb_log2_w; u->cbh = t.cb_log2_h;
                u->cbx = (uint16_t)(t.x0 - 4 * t.cu_dx); u->cby = (uint16_t)(t.y0 - 4 * t.cu_dy);
            }
            if (t.planes & VVC_CUDA_DBK_TU_CHROMA) {
                u->tbx[1] = t.x0; u->tby[1] = t.y0; u->tbw[1] = t.log2_w; u->tbh[1] = t.log2_h;
                u->qp[1] = t.qp[1]; u->qp[2] = t.qp[2];
                u->fl_c = (uint8_t)(t.flags & (VVC_CUDA_DBK_CBF_CB | VVC_CUDA_DBK_CBF_CR | VVC_CUDA_DBK_JOINT | VVC_CUDA_DBK_BDPCM_C));
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) scatter_mvf_kernel(const ParamsK p, const VVCCudaDbkMvf *mvfs, int n)
{
    const int lane = threadIdx.x & 31;
    for (int i = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); i < n; i += gridDim.x * (kThreads / 32)) {
        const VVCCudaDbkMvf m = mvfs[i];
        const size_t pic = (size_t)m.pic * p.uw * p.uh;
        DMotion mo;
        mo.mv[0][0] = m.mv[0][0]; mo.mv[0][1] = m.mv[0][1]; mo.mv[1][0] = m.mv[1][0]; mo.mv[1][1] = m.mv[1][1];
        DRef r;
        r.ref[0] = m.ref_pic[0]; r.ref[1] = m.ref_pic[1];
        for (int j = lane; j < m.w4 * m.h4; j += 32) {
            const int ux = (m.x0 >> 2) + j % m.w4, uy = (m.y0 >> 2) + j / m.w4;
            if (ux >= p.uw || uy >= p.uh)
                continue;
            const size_t at = pic + (size_t)uy * p.uw + ux;
            p.units[at].pred = (uint8_t)(m.pred_flag | (m.ciip_flag ? 0x80 : 0));
            p.motion[at] = mo;
            p.refs[at] = r;
        }
    }
}

__device__ __forceinline__ bool far_apart(const int32_t a[2], const int32_t b[2]) { return abs(a[0] - b[0]) >= 8 || abs(a[1] - b[1]) >= 8; }

// boundary_strength: motion discontinuity between two inter units (c: the Q side)
__device__ int motion_bs(int cp, int np, const DMotion &c, const DMotion &n, const DRef &cr, const DRef &nr)
{
    if (cp == 3 && np == 3) {
        if (cr.ref[0] == nr.ref[0] && cr.ref[0] == cr.ref[1] && nr.ref[0] == nr.ref[1])
            return (far_apart(n.mv[0], c.mv[0]) || far_apart(n.mv[1], c.mv[1])) && (far_apart(n.mv[1], c.mv[0]) || far_apart(n.mv[0], c.mv[1]));
        if (nr.ref[0] == cr.ref[0] && nr.ref[1] == cr.ref[1])
            return far_apart(n.mv[0], c.mv[0]) || far_apart(n.mv[1], c.mv[1]);
        if (nr.ref[1] == cr.ref[0] && nr.ref[0] == cr.ref[1])
            return far_apart(n.mv[1], c.mv[0]) || far_apart(n.mv[0], c.mv[1]);
        return 1;
    }
    if (cp != 3 && np != 3) {
        const int lc = (cp & 1) ? 0 : 1, ln = (np & 1) ? 0 : 1;
        return cr.ref[lc] == nr.ref[ln] ? far_apart(c.mv[lc], n.mv[ln]) : 1;
    }
    return 1;
}

template <bool VERT>
__global__ void __launch_bounds__(kThreads) edge_params_kernel(const ParamsK p)
{
    const int ux = blockIdx.x * 32 + (threadIdx.x & 31), uy = blockIdx.y * (kThreads / 32) + (threadIdx.x >> 5), k = blockIdx.z;
    if (ux >= p.uw || uy >= p.uh)
        return;
    const int x = ux << 2, y = uy << 2, pos = VERT ? x : y;
    const size_t pic = (size_t)k * p.uw * p.uh, at = pic + (size_t)uy * p.uw + ux;
    const VVCCudaDbkEdge none = { 0, 0, 0 };
    // map entries owned by this unit (plane units): luma always, chroma when the unit sits on the chroma edge grid
    VVCCudaDbkEdge *out_y = p.map[0] + k * p.msize[0] + (size_t)uy * p.mpitch[0] + ux;
    const int cgrid = 8 << (VERT ? p.hs : p.vs);
    const bool on_cgrid = p.planes == 3 && !(pos & (cgrid - 1));
    VVCCudaDbkEdge *out_c[2] = { nullptr, nullptr };
    if (on_cgrid) {
        const int cx = x >> p.hs, cy = y >> p.vs, seg = 4 >> (VERT ? p.vs : p.hs);
        const size_t e = VERT ? (size_t)(cy / seg) * p.mpitch[1] + cx / 8 : (size_t)(cy / 8) * p.mpitch[1] + cx / seg;
        out_c[0] = p.map[1] + k * p.msize[1] + e;
        out_c[1] = p.map[2] + k * p.msize[2] + e;
    }
    if (!pos) {                                            // picture border: never filtered
        *out_y = none;
        if (on_cgrid) { *out_c[0] = none; *out_c[1] = none; }
        return;
    }
    const size_t atp = VERT ? at - 1 : at - p.uw;
    const DUnit q = p.units[at], pu = p.units[atp];
    const int qpred = q.pred & 3, ppred = pu.pred & 3;
    const bool q_ciip = q.pred & 0x80, p_ciip = pu.pred & 0x80;
    const VVCCudaDbkCtb ct = p.ctbs[((size_t)k * p.ctb_rows + (y >> p.ctb_log2)) * p.ctb_cols + (x >> p.ctb_log2)];
    const bool ctb_blocked = !(pos & ((1 << p.ctb_log2) - 1)) && (VERT ? ct.no_left : ct.no_top);
    const bool strong = qpred == 0 || ppred == 0 || q_ciip || p_ciip;

    // ---- luma ----
    int bs = 0, lp = 0, lq = 0;
    {
        const int tb0 = VERT ? q.tbx[0] : q.tby[0], cb0 = VERT ? q.cbx : q.cby;
        const bool is_intra = qpred == 0;                  // uniform over the coding unit, so the unit's own flag serves
        const bool has_sb = !is_intra && (q.fl_y & kSb) && (1 << (VERT ? q.cbw : q.cbh)) > 8;
        if (q.tbx[0] != 0xffff && pos == tb0) {            // left / upper edge of the transform block
            if (!ctb_blocked) {
                const int off = cb0 - pos;
                if ((pu.fl_y & VVC_CUDA_DBK_BDPCM_Y) && (q.fl_y & VVC_CUDA_DBK_BDPCM_Y))      bs = 0;
                else if (strong)                                                              bs = 2;
                else if ((q.fl_y | pu.fl_y) & VVC_CUDA_DBK_CBF_Y)                             bs = 1;
                else if (off && ((off % 8) || !has_sb))                                       bs = 0;
                else bs = motion_bs(qpred, ppred, p.motion[at], p.motion[atp], p.refs[at], p.refs[atp]);
                const int size_p = 1 << (VERT ? pu.tbw[0] : pu.tbh[0]), size_q = 1 << (VERT ? q.tbw[0] : q.tbh[0]);
                if (size_p <= 4 || size_q <= 4) {
                    lp = lq = 1;
                } else {
                    lp = size_p >= 32 ? 7 : 3;
                    lq = size_q >= 32 ? 7 : 3;
                }
                if (has_sb) lq = min(5, lq);
                if (pu.fl_y & kSb) lp = min(5, lp);
            }
        } else if (q.tbx[0] != 0xffff && !is_intra && (q.fl_y & kSb) && !((pos - cb0) & 7)) {
            // an edge of the coding block's 8x8 sub-block grid inside the transform block
            const int i = pos - tb0, across = 1 << (VERT ? q.tbw[0] : q.tbh[0]);
            bs = motion_bs(qpred, ppred, p.motion[at], p.motion[atp], p.refs[at], p.refs[atp]);
            lp = lq = (i == 4 || i == across - 4) ? 1 : (i == 8 || i == across - 8) ? 2 : 3;
        }
    }
    VVCCudaDbkEdge e = none;
    if (bs) {
        int qp = (pu.qp[0] + q.qp[0] + 1) >> 1;
        if (p.prm.ladf_enabled) {
            const pel *s = p.luma + k * p.lbstride + (long long)y * p.lpitch + x;
            const int xs = VERT ? 1 : p.lpitch, ys = VERT ? p.lpitch : 1;
            const int level = ((int)s[-xs] + s[-xs + 3 * ys] + s[0] + s[3 * ys]) >> 2;
            int off = p.prm.ladf_lowest_interval_qp_offset;
            for (int i = 0; i < p.prm.num_ladf_intervals - 1 && level > p.prm.ladf_interval_lower_bound[i + 1]; i++)
                off = p.prm.ladf_qp_offset[i];
            qp += off;
        }
        e.beta = c_beta[d_clip3(qp + ct.beta_offset[0], 0, 63)];
        e.tc = c_tc[d_clip3(qp + 2 * (bs - 1) + (ct.tc_offset[0] & -2), 0, 65)];
        e.max_len = (uint8_t)(lp | (lq << 4));
    }
    *out_y = e;

    // ---- chroma ----
    if (on_cgrid) {
        const bool tb_edge = q.tbx[1] != 0xffff && pos == (VERT ? q.tbx[1] : q.tby[1]) && !ctb_blocked;
#pragma unroll
        for (int c = 1; c <= 2; c++) {
            VVCCudaDbkEdge ec = none;
            if (tb_edge) {
                int cbs = 0;
                const int cbf = c == 1 ? VVC_CUDA_DBK_CBF_CB : VVC_CUDA_DBK_CBF_CR;
                if ((pu.fl_c & VVC_CUDA_DBK_BDPCM_C) && (q.fl_c & VVC_CUDA_DBK_BDPCM_C))     cbs = 0;
                else if (strong)                                                             cbs = 2;
                else if ((pu.fl_c | q.fl_c) & (cbf | VVC_CUDA_DBK_JOINT))                    cbs = 1;
                if (cbs) {
                    const int qp = (pu.qp[c] + q.qp[c] - 2 * p.prm.qp_bd_offset + 1) >> 1;
                    const int sh = VERT ? p.hs : p.vs;
                    const int size_p = (1 << (VERT ? pu.tbw[1] : pu.tbh[1])) >> sh, size_q = (1 << (VERT ? q.tbw[1] : q.tbh[1])) >> sh;
                    int clp, clq;
                    if (size_p >= 8 && size_q >= 8) {
                        clp = clq = 3;
                        if (!VERT && !(y & ((1 << p.ctb_log2) - 1)))
                            clp = 1;
                    } else {
                        clp = clq = cbs == 2;
                    }
                    ec.beta = c_beta[d_clip3(qp + ct.beta_offset[c], 0, 63)];
                    ec.tc = c_tc[d_clip3(qp + 2 * (cbs - 1) + (ct.tc_offset[c] & -2), 0, 65)];
                    ec.max_len = (uint8_t)(clp | (clq << 4));
                }
            }
            *out_c[c - 1] = ec;
        }
    }
}

}  // namespace

extern "C" int vvc_cuda_deblock_params_frame(VVCCudaCtx *ctx, const VVCCudaFrame *f, const VVCCudaDbkTU *tus, int n_tus,
                                             const VVCCudaDbkMvf *mvfs, int n_mvfs, const VVCCudaDbkCtb *ctbs,
                                             const VVCCudaDbkParams *params, const VVCCudaDeblockMaps *maps, int dir)
{
    if (ctx->err)
        return ctx->err;
    if (!f || !tus || !mvfs || !ctbs || !params || !maps || n_tus < 0 || n_mvfs < 0)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "deblock_params: null argument");
    if ((f->bit_depth != 10 && f->bit_depth != 12) || f->ctb_log2 < 5 || f->ctb_log2 > 7 || (f->width & 7) || (f->height & 7) ||
        (f->chroma_format_idc && (f->hshift != 1 || f->vshift != 1)))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "deblock_params: unsupported picture format");
    if (params->num_ladf_intervals < 0 || params->num_ladf_intervals > 5)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "deblock_params: %d LADF intervals", params->num_ladf_intervals);
    ParamsK p;
    p.uw = f->width >> 2; p.uh = f->height >> 2; p.w = f->width; p.h = f->height; p.batch = f->batch;
    p.planes = f->chroma_format_idc ? 3 : 1; p.hs = f->hshift; p.vs = f->vshift; p.ctb_log2 = f->ctb_log2; p.bd = f->bit_depth;
    p.ctb_cols = ceil_div(f->width, 1 << f->ctb_log2); p.ctb_rows = ceil_div(f->height, 1 << f->ctb_log2);
    const size_t nu = (size_t)p.uw * p.uh * f->batch;
    const size_t usz = align_up(nu * sizeof(DUnit), 256), msz = align_up(nu * sizeof(DMotion), 256);
    uint8_t *scratch = (uint8_t *)vvc_ctx_scratch(ctx, 4, usz + msz + nu * sizeof(DRef));
    if (!scratch)
        return ctx->err;
    p.units = (DUnit *)scratch; p.motion = (DMotion *)(scratch + usz); p.refs = (DRef *)(scratch + usz + msz);
    p.luma = (const pel *)f->data[0]; p.lpitch = (int)(f->stride[0] / 2); p.lbstride = f->batch_stride[0] / 2;
    p.ctbs = ctbs; p.prm = *params;
    for (int c = 0; c < 3; c++) {
        p.map[c] = (VVCCudaDbkEdge *)maps->edge[dir][c]; p.mpitch[c] = maps->pitch[dir][c]; p.msize[c] = maps->size[dir][c];
        if (c < p.planes && !p.map[c])
            return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "deblock_params: no map array for plane %d", c);
    }
    if (p.mpitch[0] != p.uw || (p.planes == 3 && maps->rows[dir][0] < p.uh))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "deblock_params: map geometry does not match the picture");
    // units no record covers: no transform block (0xffff origins), intra, nothing coded
    VVC_TRY(ctx, cudaMemsetAsync(p.units, 0, nu * sizeof(DUnit), ctx->stream));
    VVC_TRY(ctx, cudaMemset2DAsync(p.units, sizeof(DUnit), 0xff, 8, nu, ctx->stream));
    if (n_tus) {
        const int g = ceil_div(n_tus, kThreads / 32);
        scatter_tu_kernel<<<g < 148 * 8 ? g : 148 * 8, kThreads, 0, ctx->stream>>>(p, tus, n_tus);
        VVC_LAUNCHED(ctx);
    }
    if (n_mvfs) {
        const int g = ceil_div(n_mvfs, kThreads / 32);
        scatter_mvf_kernel<<<g < 148 * 8 ? g : 148 * 8, kThreads, 0, ctx->stream>>>(p, mvfs, n_mvfs);
        VVC_LAUNCHED(ctx);
    }
    const dim3 grid(ceil_div(p.uw, 32), ceil_div(p.uh, kThreads / 32), f->batch);
    if (dir) edge_params_kernel<true><<<grid, kThreads, 0, ctx->stream>>>(p);
    else     edge_params_kernel<false><<<grid, kThreads, 0, ctx->stream>>>(p);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}
