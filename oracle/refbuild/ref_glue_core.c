/*
 * TEST INFRASTRUCTURE — never linked into the product library.
 *
 * Glue compiled into oracle/_ref/libvvcref.so together with the UNMODIFIED
 * reference sources (see oracle/Makefile).  It only exposes the reference's
 * own function-pointer tables to the parity tests.
 *
 * Reference entry used: ff_vvc_dsp_init()  libavcodec/vvc/vvcdsp.c:228-257
 */
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "libavcodec/vvc/vvcdsp.h"

/* libavutil/log.c is not part of the hot path; the DSP code only logs on
 * internal asserts.  Route it to stderr so a failing assert is visible. */
void av_log(void *avcl, int level, const char *fmt, ...)
{
    va_list ap;
    (void)avcl;
    if (level > 16)
        return;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}

static VVCDSPContext g_tables[3];
static int g_ready[3];

/* Returns the reference C tables for bit depth 8, 10 or 12. */
const VVCDSPContext *vvcref_dsp(int bit_depth)
{
    const int slot = bit_depth == 12 ? 2 : bit_depth == 10 ? 1 : 0;
    if (!g_ready[slot]) {
        ff_vvc_dsp_init(&g_tables[slot], bit_depth);
        g_ready[slot] = 1;
    }
    return &g_tables[slot];
}

size_t vvcref_dsp_sizeof(void)
{
    return sizeof(VVCDSPContext);
}

/* vvc_intra.c calls ff_log2() without the header that defines it as a macro (libavutil/intmath.h), so the compiler
 * emits a call: the function of that header (floor(log2(v)), v | 1) */
int ff_log2(unsigned v)
{
    return 31 - __builtin_clz(v | 1);
}
