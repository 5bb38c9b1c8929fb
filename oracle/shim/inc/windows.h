/* Shim for <windows.h> (R/src/core.h:148). Only MessageBox is used on the scene path. */
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline int MessageBoxA(int, const char* a, const char* b, int) { fprintf(stderr, "[ref] %s: %s\n", b, a); return 0; }
#define MessageBox MessageBoxA
static inline unsigned int timeGetTime() { return 0; }
