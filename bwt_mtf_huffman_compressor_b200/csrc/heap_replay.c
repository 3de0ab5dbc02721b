/* heap_replay.c -- bzap_heap_replay <N> <n_leaves> : the address order of the reference's Huffman nodes.
 *
 * The reference breaks frequency ties by the ADDRESS of its BTree nodes (std::pair<long, BTree*> in the
 * priority queue, main.cpp:232, 240-241, 252-253).  For N >= 64,600 and in the windows listed in SURVEY App. B
 * the addresses follow a closed-form law (huffman_host.cpp); elsewhere they depend on where the small chunks
 * freed while the input was read happen to lie.  Nothing about the DATA enters: up to the last `new BTree`
 * the one-shot COMPRESS process issues a fixed script of malloc / free calls whose sizes depend only on N and
 * on the number of leaves.  This helper replays that script against the allocator of the host it runs on (the
 * same glibc the reference would use) in a fresh process and prints, for every node in creation order (leaves
 * by first appearance, then the merges), its position in ascending address order.
 *
 * The script (traced with an LD_PRELOAD malloc logger, tests/test_heap_replay.py keeps it honest):
 *   libstdc++ start-up pool 73728 | two path strings | FILE 472 + stream buffer 8192                (main, read_bytes)
 *   the input vector grown by push_back: capacities 1, 2, 4, ... >= N, each new buffer before the old is freed
 *   stream buffer and FILE freed                                                                   (io_utilities.h:29-55)
 *   N, N, free(read buffer), free, N, 8N (shift order), 8 ceil(N/2) (stable_sort buffer), free, N, N, free x3   (bwt, main.cpp:77-91)
 *   N, N, 256, N, free x2                                                                           (move_to_front, :93-112)
 *   2048 (frequencies), 32 (vector<bool>), then per leaf: node 24, and when the queue is full its next buffer
 *   16 * {1, 2, 4, ...} before the old one is freed; then one node per merge                        (huffman, :229-254)
 * Plain C, no stdio buffers before the script ends (they would be allocations of their own). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static void *sink[64];
static int nsink;
static void *keep(void *p) { if (nsink < 64) sink[nsink++] = p; return p; }

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    const long n = atol(argv[1]);
    const int leaves = atoi(argv[2]);
    if (n < 1 || n > (1L << 30) || leaves < 1 || leaves > 256) return 2;
    static void *node[511];
    keep(malloc(73728));
    keep(malloc(18));
    keep(malloc(21));
    void *file = malloc(472), *sbuf = malloc(8192);
    /* read_bytes: push_back growth */
    size_t cap = 1;
    void *v = malloc(1);
    while ((long)cap < n) {
        void *nv = malloc(cap * 2);
        free(v);
        v = nv;
        cap *= 2;
    }
    free(sbuf);
    free(file);
    /* bwt */
    void *a = malloc((size_t)n);
    keep(malloc((size_t)n));
    free(v);
    free(a);
    void *c = malloc((size_t)n), *d = malloc(8 * (size_t)n), *e = malloc(8 * (size_t)((n + 1) / 2));
    free(e);
    void *f = malloc((size_t)n);
    keep(malloc((size_t)n));
    free(f);
    free(d);
    free(c);
    /* move_to_front */
    keep(malloc((size_t)n));
    void *i = malloc((size_t)n), *j = malloc(256);
    keep(malloc((size_t)n));
    free(j);
    free(i);
    /* huffman */
    keep(malloc(2048));
    keep(malloc(32));
    size_t qcap = 0, qsize = 0;
    void *q = NULL;
    int k = 0;
    for (; k < leaves; ++k) {
        node[k] = malloc(24);
        if (qsize == qcap) {
            const size_t ncap = qcap ? 2 * qcap : 1;
            void *nq = malloc(16 * ncap);
            free(q);
            q = nq;
            qcap = ncap;
        }
        ++qsize;
    }
    for (int m = 0; m + 1 < leaves; ++m, ++k) node[k] = malloc(24);
    /* rank of every node in ascending address order */
    const int total = 2 * leaves - 1;
    char out[511 * 5 + 8];
    size_t len = 0;
    for (int x = 0; x < total; ++x) {
        int r = 0;
        for (int y = 0; y < total; ++y) r += (uintptr_t)node[y] < (uintptr_t)node[x];
        char tmp[8];
        int t = 0;
        do { tmp[t++] = (char)('0' + r % 10); r /= 10; } while (r);
        while (t) out[len++] = tmp[--t];
        out[len++] = x + 1 < total ? ' ' : '\n';
    }
    return write(1, out, len) == (ssize_t)len ? 0 : 1;
}
