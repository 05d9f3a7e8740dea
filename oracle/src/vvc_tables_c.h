/* Constant tables for the oracle's C translation units (generated data, see tools/gen_tables.py). */
#ifndef VVC_TABLES_C_H
#define VVC_TABLES_C_H
#include <stdint.h>
#define VVCT_TABLE(type, name, dims) static const type name dims __attribute__((unused))
#include "../../ffvvc_b200/csrc/vvc_tables.inc"
#undef VVCT_TABLE
#endif
