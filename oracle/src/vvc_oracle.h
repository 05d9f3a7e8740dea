/*
 * vvc_oracle.h - shared helpers of the CPU oracle.
 *
 * TEST INFRASTRUCTURE.  This directory is a plain-C restatement of the reference's
 * pixel-reconstruction arithmetic, written against the same descriptors as the CUDA
 * library (include/vvcdsp_cuda.h) so tests can compare the two bit for bit.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the
 * product library never links it.
 *
 * Parity pinning: every function here is checked against the compiled, unmodified
 * reference (oracle/_ref/libvvcref.so, loops of the reference's own table entries in the
 * reference drivers' order) by tests/test_oracle_vs_ref_*.py and against the committed
 * fixtures in tests/golden/ (generated from that same reference by tools/gen_golden.py).
 *
 * Integer helpers follow libavutil/common.h:174-182 (av_clip), :262-280 (clip_intp2,
 * clip_uintp2) and the >>-on-negatives-is-arithmetic rule of SURVEY.md Appendix A.
 */
#ifndef VVC_ORACLE_H
#define VVC_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "vvcdsp_cuda.h"

typedef uint16_t pel;

static inline int o_clip3(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static inline int o_clip_pel(int v, int bd) { return o_clip3(v, 0, (1 << bd) - 1); }
static inline int o_clip_sbits(int v, int bits) { return o_clip3(v, -(1 << bits), (1 << bits) - 1); }
static inline int o_clip_ubits(int v, int bits) { return o_clip3(v, 0, (1 << bits) - 1); }
static inline int o_abs(int v) { return v < 0 ? -v : v; }
static inline int o_min(int a, int b) { return a < b ? a : b; }
static inline int o_max(int a, int b) { return a > b ? a : b; }
static inline int o_sign(int v) { return (v > 0) - (v < 0); }
static inline int o_ilog2(unsigned v) { int n = 0; while (v >>= 1) n++; return n; } /* av_log2: floor, log2(0)=0 */

/* one plane of picture k of a VVCCudaFrame */
typedef struct OPlane {
    pel      *p;
    ptrdiff_t pitch;     /* in samples */
    int       w, h;
} OPlane;

static inline OPlane o_plane(const VVCCudaFrame *f, int c, int k)
{
    OPlane pl;
    pl.p     = (pel *)((uint8_t *)f->data[c] + (ptrdiff_t)k * f->batch_stride[c]);
    pl.pitch = f->stride[c] / (ptrdiff_t)sizeof(pel);
    pl.w     = c ? f->width  >> f->hshift : f->width;
    pl.h     = c ? f->height >> f->vshift : f->height;
    return pl;
}

static inline int o_ctb_cols(const VVCCudaFrame *f) { return (f->width  + (1 << f->ctb_log2) - 1) >> f->ctb_log2; }
static inline int o_ctb_rows(const VVCCudaFrame *f) { return (f->height + (1 << f->ctb_log2) - 1) >> f->ctb_log2; }

#endif
