/*
 * Hand-written stand-in for FFmpeg's configure-generated config.h.
 * TEST INFRASTRUCTURE ONLY: lets oracle/Makefile compile the few reference
 * DSP sources where they lie under /root/reference without running the
 * reference's own build system.  Pure-C path: no arch-specific code enabled.
 */
#ifndef VVCREF_STUB_CONFIG_H
#define VVCREF_STUB_CONFIG_H
#define ARCH_AARCH64 0
#define ARCH_ARM 0
#define ARCH_AVR32 0
#define ARCH_MIPS 0
#define ARCH_PPC 0
#define ARCH_RISCV 0
#define ARCH_X86 0
#define ARCH_X86_32 0
#define ARCH_X86_64 0
#define HAVE_BIGENDIAN 0
#define HAVE_FAST_UNALIGNED 1
#define HAVE_FAST_64BIT 1
#define HAVE_FAST_CLZ 1
#define HAVE_LOCAL_ALIGNED 1
#define HAVE_THREADS 1
#define HAVE_PTHREADS 1
#define HAVE_INLINE_ASM 0
#define HAVE_MMX_INLINE 0
#define CONFIG_SMALL 0
#define CONFIG_SAFE_BITSTREAM_READER 1
#define CONFIG_MEMORY_POISONING 0
#define CONFIG_FTRAPV 0
#define av_restrict restrict
#endif
