// H.266 constant tables in device memory (generated data: tools/gen_tables.py).
// Each translation unit gets its own static copy; unused tables are dropped by the compiler.
#pragma once
#include <stdint.h>
#define VVCT_TABLE(type, name, dims) static __device__ const __align__(16) type name dims
#include "vvc_tables.inc"
#undef VVCT_TABLE
