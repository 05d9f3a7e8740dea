/*
 * TEST INFRASTRUCTURE - glue of the out-of-tree build of the reference's checkasm harness (oracle/Makefile, target
 * `checkasm`): tests/checkasm/{checkasm,vvc_alf,vvc_itx,vvc_mc,vvc_sao}.c and libavcodec/vvc/vvcdsp.c are compiled
 * UNMODIFIED, where they lie, with oracle/refbuild/stubcfg_x86/config.h (an x86 build without external assembly).
 *
 * The reference's own arch hook is the seam: ff_vvc_dsp_init() (libavcodec/vvc/vvcdsp.c:228-257) ends with
 *     #if ARCH_X86
 *         ff_vvc_dsp_init_x86(vvcdsp, bit_depth);
 * and this file provides ff_vvc_dsp_init_x86().  When the CPU flags checkasm is currently forcing contain AVX2 it calls
 * ff_vvc_dsp_init_cuda() of libvvcdsp_cuda.so - exactly the place the x86 hook installs its AVX2 entries
 * (libavcodec/x86/vvc/vvcdsp_init.c:294-361) - so checkasm's "AVX2" pass compares every CUDA-backed table entry with the
 * reference C entry on the harness's own random inputs.
 *
 * The rest are the few libavutil symbols the harness needs besides lfg.c (CPU flags, a seed, logging, basename).
 */
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include "libavutil/cpu.h"
#include "libavcodec/vvc/vvcdsp.h"

void ff_vvc_dsp_init_cuda(VVCDSPContext *c, int bit_depth);      /* include/vvcdsp_table.h */
int  ff_vvc_dsp_cuda_last_error(void);
const char *ff_vvc_dsp_cuda_error_string(void);

static int g_flags = -1;                                          /* -1: not forced */
#define HOST_FLAGS (AV_CPU_FLAG_MMX | AV_CPU_FLAG_CMOV | AV_CPU_FLAG_AVX2)

int av_get_cpu_flags(void) { return g_flags == -1 ? HOST_FLAGS : g_flags; }
void av_force_cpu_flags(int flags) { g_flags = flags; }

void ff_vvc_dsp_init_x86(VVCDSPContext *c, const int bit_depth)
{
    if (av_get_cpu_flags() & AV_CPU_FLAG_AVX2) {
        ff_vvc_dsp_init_cuda(c, bit_depth);
        if (ff_vvc_dsp_cuda_last_error()) {
            fprintf(stderr, "checkasm glue: %s\n", ff_vvc_dsp_cuda_error_string());
            fflush(stderr);
        }
    }
}

/* checked by tests/test_gpu_checkasm.py after the run: a latched CUDA error fails the test even if no comparison did */
__attribute__((destructor)) static void report_cuda_error(void)
{
    if (ff_vvc_dsp_cuda_last_error())
        fprintf(stderr, "checkasm glue: CUDA table error at exit: %s\n", ff_vvc_dsp_cuda_error_string());
}

uint32_t av_get_random_seed(void) { return (uint32_t)time(NULL) * 2654435761u; }

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

const char *av_basename(const char *path)
{
    const char *p = path ? strrchr(path, '/') : NULL;
    return p ? p + 1 : (path ? path : ".");
}
