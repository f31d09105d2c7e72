// Wrapper that exposes the reference's own BSWAP_64 macro (cudautil.cuh:118-125) as a C symbol.
// Test infrastructure: built by oracle/Makefile into oracle/_ref/ from the header where it lies
// under /root/reference (never copied); used only by tools/make_golden.py to produce
// tests/golden/bswap64_vectors.json.
#include <stdint.h>
#include "cudautil.cuh"
extern "C" uint64_t ref_bswap_64(uint64_t x) { return BSWAP_64(x); }
