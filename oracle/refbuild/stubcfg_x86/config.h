/*
 * Second hand-written config.h stand-in, for the out-of-tree build of the reference's checkasm harness
 * (oracle/Makefile, target `checkasm`).  TEST INFRASTRUCTURE ONLY.  It differs from ../stubcfg/config.h in
 * declaring an x86 build WITHOUT external assembly: checkasm then iterates over its x86 CPU-flag list and the
 * reference's vvcdsp.c calls its arch hook ff_vvc_dsp_init_x86() - which oracle/refbuild/chk_glue.c provides
 * and forwards to ff_vvc_dsp_init_cuda() of libvvcdsp_cuda.so when the "AVX2" flag is the one being checked.
 */
#ifndef VVCREF_STUB_CONFIG_X86_H
#define VVCREF_STUB_CONFIG_X86_H
#define ARCH_AARCH64 0
#define ARCH_ARM 0
#define ARCH_AVR32 0
#define ARCH_LOONGARCH 0
#define ARCH_MIPS 0
#define ARCH_PPC 0
#define ARCH_RISCV 0
#define ARCH_X86 1
#define ARCH_X86_32 0
#define ARCH_X86_64 1
#define HAVE_X86ASM 0
#define HAVE_BIGENDIAN 0
#define HAVE_FAST_UNALIGNED 1
#define HAVE_FAST_64BIT 1
#define HAVE_FAST_CLZ 1
#define HAVE_LOCAL_ALIGNED 1
#define HAVE_THREADS 1
#define HAVE_PTHREADS 1
#define HAVE_INLINE_ASM 0
#define HAVE_MMX_INLINE 0
#define HAVE_MMX_EXTERNAL 0
#define HAVE_UNISTD_H 1
#define HAVE_ISATTY 1
#define HAVE_IO_H 0
#define HAVE_SETCONSOLETEXTATTRIBUTE 0
#define HAVE_GETSTDHANDLE 0
#define HAVE_ARMV5TE_EXTERNAL 0
#define HAVE_LINUX_PERF 0
#define HAVE_MACOS_KPERF 0
#define HAVE_RDTSC 0
#define HAVE_RV 0
#define CONFIG_LINUX_PERF 0
#define CONFIG_MACOS_KPERF 0
#define CONFIG_SMALL 0
#define CONFIG_SAFE_BITSTREAM_READER 1
#define CONFIG_MEMORY_POISONING 0
#define CONFIG_FTRAPV 0
#define CONFIG_VVC_DECODER 1
#define CONFIG_AVCODEC 1
#define HAVE_ATAN2F 1
#define HAVE_ATANF 1
#define HAVE_CBRT 1
#define HAVE_CBRTF 1
#define HAVE_COPYSIGN 1
#define HAVE_COSF 1
#define HAVE_ERF 1
#define HAVE_EXP2 1
#define HAVE_EXP2F 1
#define HAVE_EXPF 1
#define HAVE_HYPOT 1
#define HAVE_ISFINITE 1
#define HAVE_ISINF 1
#define HAVE_ISNAN 1
#define HAVE_LDEXPF 1
#define HAVE_LLRINT 1
#define HAVE_LLRINTF 1
#define HAVE_LOG10F 1
#define HAVE_LOG2 1
#define HAVE_LOG2F 1
#define HAVE_LRINT 1
#define HAVE_LRINTF 1
#define HAVE_POWF 1
#define HAVE_RINT 1
#define HAVE_ROUND 1
#define HAVE_ROUNDF 1
#define HAVE_SINF 1
#define HAVE_TRUNC 1
#define HAVE_TRUNCF 1
#define HAVE_MIPSFPU 0
#define HAVE_GETHRTIME 0
#define HAVE_LIBC_MSVCRT 0
#define HAVE_MACH_ABSOLUTE_TIME 0
#define HAVE_PRAGMA_DEPRECATED 1
#define CONFIG_SHARED 0
#define av_restrict restrict
#endif
