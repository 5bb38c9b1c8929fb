/* Shim for the pool allocator header that is absent from the reference checkout
 * (R/src/main.cpp:14 includes "../src.BestFitMem/bmalloc.h"). Test infrastructure only. */
#pragma once
#include <stdlib.h>
static inline void* bmalloc(size_t n) { return malloc(n); }
static inline void  bfree(void* p) { free(p); }
static inline void  add_pool(void*, size_t) {}
