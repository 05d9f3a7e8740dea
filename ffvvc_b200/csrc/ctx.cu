// Context management of libvvcdsp_cuda.so: stream ownership, sticky error latch, staging.
// Error model follows SURVEY.md 8(b): the reference's DSP entries return void, so CUDA failures
// are latched here and surface from vvc_cuda_sync()/vvc_cuda_last_error().
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

int vvc_ctx_fail(VVCCudaCtx *ctx, int code, const char *fmt, ...)
{
    if (ctx->err == VVC_CUDA_OK) {
        va_list ap;
        ctx->err = code;
        va_start(ap, fmt);
        vsnprintf(ctx->msg, sizeof(ctx->msg), fmt, ap);
        va_end(ap);
    }
    return code;
}

int vvc_ctx_check(VVCCudaCtx *ctx, cudaError_t e, const char *what)
{
    if (e == cudaSuccess)
        return 0;
    vvc_ctx_fail(ctx, VVC_CUDA_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return 1;
}

// the copy streams of the host entries: an asynchronous call may have left copies running on them
static void drain_copy_streams(VVCCudaCtx *ctx)
{
    if (ctx->host_pending) {
        vvc_ctx_check(ctx, cudaStreamSynchronize(ctx->copy_in), "cudaStreamSynchronize(copy_in)");
        vvc_ctx_check(ctx, cudaStreamSynchronize(ctx->copy_out), "cudaStreamSynchronize(copy_out)");
        ctx->host_pending = false;
    }
}

void *vvc_ctx_dev_stage(VVCCudaCtx *ctx, size_t bytes)
{
    // the staging area may still be the source / destination of an asynchronous host entry's copies: any other user waits
    if (ctx->host_pending && !ctx->host_owner) {
        cudaStreamSynchronize(ctx->stream);
        drain_copy_streams(ctx);
    }
    if (bytes > ctx->d_stage_size) {
        if (ctx->d_stage) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(ctx->d_stage);
        }
        ctx->d_stage = NULL; ctx->d_stage_size = 0;
        size_t want = bytes + bytes / 4 + (1 << 20);
        if (vvc_ctx_check(ctx, cudaMalloc(&ctx->d_stage, want), "cudaMalloc(stage)"))
            return NULL;
        ctx->d_stage_size = want;
    }
    return ctx->d_stage;
}

void *vvc_ctx_host_stage(VVCCudaCtx *ctx, size_t bytes)
{
    if (bytes > ctx->h_stage_size) {
        if (ctx->h_stage) {
            cudaStreamSynchronize(ctx->stream);
            cudaFreeHost(ctx->h_stage);
        }
        ctx->h_stage = NULL; ctx->h_stage_size = 0;
        size_t want = bytes + bytes / 4 + (1 << 20);
        if (vvc_ctx_check(ctx, cudaMallocHost(&ctx->h_stage, want), "cudaMallocHost(stage)"))
            return NULL;
        ctx->h_stage_size = want;
    }
    return ctx->h_stage;
}

void *vvc_ctx_scratch(VVCCudaCtx *ctx, int slot, size_t bytes)
{
    if (bytes > ctx->d_scratch_size[slot]) {
        if (ctx->d_scratch[slot]) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(ctx->d_scratch[slot]);
        }
        ctx->d_scratch[slot] = NULL; ctx->d_scratch_size[slot] = 0;
        if (vvc_ctx_check(ctx, cudaMalloc(&ctx->d_scratch[slot], bytes), "cudaMalloc(scratch)"))
            return NULL;
        ctx->d_scratch_size[slot] = bytes;
    }
    return ctx->d_scratch[slot];
}

int vvc_ctx_copy_streams(VVCCudaCtx *ctx)
{
    if (ctx->copy_in)
        return 0;
    VVC_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    VVC_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < 8; i++)
        VVC_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 16; i++)
        VVC_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; i++)
        VVC_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_refs[i], cudaEventDisableTiming));
    for (int i = 0; i < 12; i++)
        VVC_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_desc[i / 4][i % 4], cudaEventDisableTiming));
    return 0;
}

int vvc_ctx_fork(VVCCudaCtx *ctx, int n)
{
    if (!ctx->fork_ev) {
        VVC_TRY(ctx, cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
        for (int i = 0; i < 3; i++) {
            VVC_TRY(ctx, cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking));
            VVC_TRY(ctx, cudaEventCreateWithFlags(&ctx->join_ev[i], cudaEventDisableTiming));
        }
    }
    VVC_TRY(ctx, cudaEventRecord(ctx->fork_ev, ctx->stream));
    for (int i = 0; i < n; i++)
        VVC_TRY(ctx, cudaStreamWaitEvent(ctx->side[i], ctx->fork_ev, 0));
    return 0;
}

int vvc_ctx_join(VVCCudaCtx *ctx, int n)
{
    for (int i = 0; i < n; i++) {
        VVC_TRY(ctx, cudaEventRecord(ctx->join_ev[i], ctx->side[i]));
        VVC_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->join_ev[i], 0));
    }
    return 0;
}

extern "C" {

const char *vvc_cuda_version(void) { return "vvcdsp-b200 0.1 (sm_100a)"; }

int vvc_cuda_ctx_create(VVCCudaCtx **out, int device, void *stream)
{
    if (!out)
        return VVC_CUDA_ERR_ARG;
    *out = NULL;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n)
        return VVC_CUDA_ERR_CUDA;          // no usable device: fail loudly, never fall back
    VVCCudaCtx *ctx = (VVCCudaCtx *)calloc(1, sizeof(*ctx));
    if (!ctx)
        return VVC_CUDA_ERR_NOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { free(ctx); return VVC_CUDA_ERR_CUDA; }
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { free(ctx); return VVC_CUDA_ERR_CUDA; }
        ctx->own_stream = true;
    }
    *out = ctx;
    return VVC_CUDA_OK;
}

void vvc_cuda_ctx_destroy(VVCCudaCtx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    for (int i = 0; i < 6; i++) if (ctx->d_scratch[i]) cudaFree(ctx->d_scratch[i]);
    if (ctx->fork_ev) {
        cudaEventDestroy(ctx->fork_ev);
        for (int i = 0; i < 3; i++) { cudaStreamSynchronize(ctx->side[i]); cudaStreamDestroy(ctx->side[i]); cudaEventDestroy(ctx->join_ev[i]); }
    }
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    for (int i = 0; i < 8; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 16; i++) if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
    for (int i = 0; i < 2; i++) if (ctx->ev_refs[i]) cudaEventDestroy(ctx->ev_refs[i]);
    for (int i = 0; i < 12; i++) if (ctx->ev_desc[i / 4][i % 4]) cudaEventDestroy(ctx->ev_desc[i / 4][i % 4]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    free(ctx);
}

int vvc_cuda_sync(VVCCudaCtx *ctx)
{
    vvc_ctx_check(ctx, cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    drain_copy_streams(ctx);
    return ctx->err;
}

namespace {
struct Notify { vvc_cuda_notify_fn fn; void *opaque; const VVCCudaCtx *ctx; };
void CUDART_CB notify_trampoline(void *p)
{
    Notify *n = (Notify *)p;
    n->fn(n->opaque, n->ctx->err);
    free(n);
}
}  // namespace

int vvc_cuda_notify(VVCCudaCtx *ctx, vvc_cuda_notify_fn fn, void *opaque)
{
    if (!ctx || !fn)
        return VVC_CUDA_ERR_ARG;
    if (ctx->err)
        return ctx->err;
    Notify *n = (Notify *)malloc(sizeof(*n));
    if (!n)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_NOMEM, "notify: out of memory");
    n->fn = fn; n->opaque = opaque; n->ctx = ctx;
    // "everything submitted so far" includes the copy-out of an asynchronous host entry: the report waits for it
    if (ctx->host_pending &&
        (vvc_ctx_check(ctx, cudaEventRecord(ctx->ev[7], ctx->copy_out), "cudaEventRecord") ||
         vvc_ctx_check(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev[7], 0), "cudaStreamWaitEvent"))) {
        free(n);
        return ctx->err;
    }
    if (vvc_ctx_check(ctx, cudaLaunchHostFunc(ctx->stream, notify_trampoline, n), "cudaLaunchHostFunc")) {
        free(n);
        return ctx->err;
    }
    return VVC_CUDA_OK;
}

int         vvc_cuda_last_error(const VVCCudaCtx *ctx)   { return ctx->err; }
const char *vvc_cuda_error_string(const VVCCudaCtx *ctx) { return ctx->err ? ctx->msg : "ok"; }
void       *vvc_cuda_stream(const VVCCudaCtx *ctx)       { return (void *)ctx->stream; }
uint64_t    vvc_cuda_launch_count(const VVCCudaCtx *ctx) { return ctx->launches; }

int vvc_cuda_ctx_set_option(VVCCudaCtx *ctx, int option, int value)
{
    if (!ctx)
        return VVC_CUDA_ERR_ARG;
    switch (option) {
    case VVC_CUDA_OPT_GENERIC_KERNELS: ctx->force_generic = value != 0; return VVC_CUDA_OK;
    case VVC_CUDA_OPT_ALF_WIDE_MULTIPLY: ctx->alf_wide_multiply = value != 0; return VVC_CUDA_OK;
    case VVC_CUDA_OPT_INTER_TMA: ctx->inter_tma = value != 0; return VVC_CUDA_OK;
    case VVC_CUDA_OPT_REF_PAD:
        if (value < 0 || (value & 15) || value > 1024)       // 16: the chroma planes stay 16-byte aligned
            return VVC_CUDA_ERR_ARG;
        ctx->ref_pad = value;
        return VVC_CUDA_OK;
    default: return VVC_CUDA_ERR_ARG;
    }
}

}  // extern "C"

// sizeof() of the descriptor PODs as compiled here, so bindings can verify their mirrors.
extern "C" size_t vvc_cuda_abi_sizeof(int which)
{
    switch (which) {
    case 0: return sizeof(VVCCudaFrame);
    case 1: return sizeof(VVCCudaALFCtb);
    case 2: return sizeof(VVCCudaALFSets);
    case 3: return sizeof(VVCCudaDbkEdge);
    case 4: return sizeof(VVCCudaDeblockMaps);
    case 5: return sizeof(VVCCudaSAOCtb);
    case 6: return sizeof(VVCCudaInloopDesc);
    case 7: return sizeof(VVCCudaTB);
    case 8: return sizeof(VVCCudaPB);
    case 9: return sizeof(VVCCudaWP);
    case 10: return sizeof(VVCCudaProf);
    case 11: return sizeof(VVCCudaDmvrOut);
    case 12: return sizeof(VVCCudaRect);
    case 13: return sizeof(VVCCudaReconDesc);
    case 14: return sizeof(VVCCudaIntraPB);
    case 15: return sizeof(VVCCudaCiip);
    case 16: return sizeof(VVCCudaTBQuant);
    case 17: return sizeof(VVCCudaScalingList);
    case 18: return sizeof(VVCCudaCoeffs);
    case 19: return sizeof(VVCCudaLmcsVpdu);
    case 20: return sizeof(VVCCudaLmcsParams);
    case 21: return sizeof(VVCCudaIntraBlk);
    case 22: return sizeof(VVCCudaDbkTU);
    case 23: return sizeof(VVCCudaDbkMvf);
    case 24: return sizeof(VVCCudaDbkCtb);
    case 25: return sizeof(VVCCudaDbkParams);
    case 26: return sizeof(VVCCudaDbkSide);
    default: return 0;
    }
}
