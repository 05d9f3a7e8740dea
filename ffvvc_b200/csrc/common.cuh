// Shared declarations of libvvcdsp_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "vvcdsp_cuda.h"

typedef uint16_t pel;   // high-bit-depth sample, as in the reference's 10/12-bit templates

struct VVCCudaCtx {
    int           device;
    cudaStream_t  stream;
    bool          own_stream;
    int           err;                  // sticky VVC_CUDA_ERR_*
    char          msg[256];
    uint64_t      launches;
    int           force_generic;        // vvc_cuda_ctx_set_option(VVC_CUDA_OPT_GENERIC_KERNELS)
    int           alf_wide_multiply;    // vvc_cuda_ctx_set_option(VVC_CUDA_OPT_ALF_WIDE_MULTIPLY)
    int           inter_tma;            // vvc_cuda_ctx_set_option(VVC_CUDA_OPT_INTER_TMA)
    int           ref_pad;              // vvc_cuda_ctx_set_option(VVC_CUDA_OPT_REF_PAD): replicated luma samples around the reference planes
    bool          itx_packed;           // itx_warp.cu's packed transform matrices are built on this device
    // staging for the *_host entries and the per-call table shims
    void         *d_stage;  size_t d_stage_size;
    void         *h_stage;  size_t h_stage_size;    // pinned
    void         *d_scratch[6]; size_t d_scratch_size[6];   // [0],[1] intermediate pictures of chained stages, [2] inter / residual task lists, [3] LMCS chroma scales, [4] deblocking side tables, [5] derived deblocking maps
    cudaStream_t  copy_in, copy_out;                 // lazily created, *_host pipelines
    cudaStream_t  side[3];                           // lazily created: independent kernels of one stage run beside the context stream
    cudaEvent_t   fork_ev, join_ev[3];
    cudaEvent_t   ev[8];
    cudaEvent_t   ev_desc[3][4];                      // reconstruction host entry, per descriptor slot: uploaded, kernels done, refined vectors downloaded
    cudaEvent_t   ev_out[16], ev_refs[2];             // reconstruction host entry: copy-out of an output slot finished, kernels of the call that used a reference area finished
    // vvc_cuda_recon_frame_host(_async): pictures issued so far (slot and event rotation continue across calls), the staging
    // layout of the calls in flight, and whether copies of an earlier call may still be running on the copy streams
    uint64_t      host_seq, host_calls;
    size_t        host_layout[4];
    bool          host_pending, host_owner;
};

int  vvc_ctx_fail(VVCCudaCtx *ctx, int code, const char *fmt, ...);
int  vvc_ctx_check(VVCCudaCtx *ctx, cudaError_t e, const char *what);
// Grow-only staging buffers; contents are not preserved across a grow.
void *vvc_ctx_dev_stage(VVCCudaCtx *ctx, size_t bytes);
void *vvc_ctx_host_stage(VVCCudaCtx *ctx, size_t bytes);
void *vvc_ctx_scratch(VVCCudaCtx *ctx, int slot, size_t bytes);
// Fork / join inside a stage: after vvc_ctx_fork the side streams ctx->side[0 .. n) see everything issued on
// ctx->stream so far; vvc_ctx_join makes ctx->stream wait for what was issued on them since.  The kernels of one stage
// that touch disjoint samples (task classes of the inter stage, the two residual kernels) are issued this way so that
// the tail of one fills with the next instead of draining the machine between launches.
// the copy streams and events of the *_host pipelines, created on first use
int   vvc_ctx_copy_streams(VVCCudaCtx *ctx);
int   vvc_ctx_fork(VVCCudaCtx *ctx, int n);
int   vvc_ctx_join(VVCCudaCtx *ctx, int n);

// The *_host entries stage host reference pictures without margins: VVC_CUDA_OPT_REF_PAD does not apply to them.
struct VVCRefPadOff {
    VVCCudaCtx *ctx; int saved;
    explicit VVCRefPadOff(VVCCudaCtx *c, bool off) : ctx(c), saved(c->ref_pad) { if (off) c->ref_pad = 0; }
    ~VVCRefPadOff() { ctx->ref_pad = saved; }
};

#define VVC_TRY(ctx, call)  do { if (vvc_ctx_check((ctx), (call), #call)) return (ctx)->err; } while (0)
#define VVC_LAUNCHED(ctx)   do { (ctx)->launches++; if (vvc_ctx_check((ctx), cudaGetLastError(), "kernel launch")) return (ctx)->err; } while (0)

// ---- integer helpers with the reference's semantics (libavutil/common.h:174-280) ----
__device__ __forceinline__ int d_clip3(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int d_clip_pel(int v, int bd) { return min(max(v, 0), (1 << bd) - 1); }
__device__ __forceinline__ int d_clip_sbits(int v, int bits) { return min(max(v, -(1 << bits)), (1 << bits) - 1); }
__device__ __forceinline__ int d_clip_ubits(int v, int bits) { return min(max(v, 0), (1 << bits) - 1); }
__device__ __forceinline__ int d_ilog2(unsigned v) { return v ? 31 - __clz(v) : 0; }   // av_log2
__device__ __forceinline__ int d_sign(int v) { return (v > 0) - (v < 0); }
// one residual through lmcs_scale_chroma (libavcodec/vvc/vvc_intra_template.c:431-448): av_clip_intp2(res, bit_depth),
// magnitude * scale rounded at 11 bits, sign restored
__device__ __forceinline__ int d_lmcs_scale(int res, int scale, int bd)
{
    const int c = d_clip_sbits(res, bd), m = (abs(c) * scale + (1 << 10)) >> 11;
    return c > 0 ? m : -m;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Geometry checks shared by the frame-level entries.
static inline bool frame_vec_ok(const VVCCudaFrame *f)
{
    const int planes = f->chroma_format_idc ? 3 : 1;
    for (int c = 0; c < planes; c++)
        if (((uintptr_t)f->data[c] & 15) || (f->stride[c] & 15) || (f->batch_stride[c] & 15))
            return false;
    return true;
}

// ---- host<->device staging of pictures for the *_host entries (host_stage.cu) ----
size_t vvc_stage_frame_size(const VVCCudaFrame *f);
// Lay a device picture ring of the same geometry as `host` out at `dbase` (256-byte pitch).
void   vvc_stage_frame_layout(const VVCCudaFrame *host, void *dbase, VVCCudaFrame *dev);
int    vvc_stage_frame_h2d(VVCCudaCtx *ctx, const VVCCudaFrame *dev, const VVCCudaFrame *host);
int    vvc_stage_frame_d2h(VVCCudaCtx *ctx, const VVCCudaFrame *host, const VVCCudaFrame *dev);
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// itx_warp.cu (10-bit pictures, log2_transform_range 15)
int vvc_itx_launch_warp(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaCoeffs *coeffs, const VVCCudaTB *tbs, int n_tbs,
                        uint32_t *scratch, const uint32_t **rest, const uint32_t **rest_count);
