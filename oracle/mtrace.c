/* oracle/mtrace.c -- TEST INFRASTRUCTURE: an LD_PRELOAD malloc logger for the unmodified reference binary.
 * It records the addresses of the reference's BTree nodes (the size-24 allocations that follow the
 * malloc(2048); malloc(32) pair of huffman(), main.cpp:231-234) in creation order and writes them to
 * $MTRACE_OUT when the process ends.  tests/test_heap_replay.py compares their address order with what
 * bzap_heap_replay (csrc/heap_replay.c) predicts.  SURVEY App. E, probe 2. */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
extern void *__libc_malloc(size_t);
extern void __libc_free(void *);
#define CAP (1 << 20)
static struct { char op; size_t size; void *p; } logv[CAP];
static int nlog = 0;
void *malloc(size_t s)
{
    void *p = __libc_malloc(s);
    if (nlog < CAP) { logv[nlog].op = 'm'; logv[nlog].size = s; logv[nlog].p = p; ++nlog; }
    return p;
}
void free(void *p)
{
    if (p && nlog < CAP) { logv[nlog].op = 'f'; logv[nlog].size = 0; logv[nlog].p = p; ++nlog; }
    __libc_free(p);
}
__attribute__((destructor)) static void dump(void)
{
    const char *path = getenv("MTRACE_OUT");
    if (!path) return;
    FILE *f = fopen(path, "w");
    if (!f) return;
    int start = -1;
    for (int i = 0; i + 1 < nlog; ++i)
        if (logv[i].op == 'm' && logv[i].size == 2048 && logv[i + 1].op == 'm' && logv[i + 1].size == 32) start = i + 2;
    for (int i = start < 0 ? nlog : start; i < nlog; ++i)
        if (logv[i].op == 'm' && logv[i].size == 24) fprintf(f, "%p\n", logv[i].p);
    fclose(f);
}
