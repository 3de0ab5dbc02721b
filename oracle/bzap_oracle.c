/*
 * bzap_oracle.c -- CPU restatement of the reference hot path (see bzap_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into, loaded by or executed from the product.
 * Parity: pinned against the unmodified reference (oracle/_ref) by tests/test_oracle_vs_ref.py.
 */
#include "bzap_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Forward BWT.  Reference: main.cpp:38-59 (cyclic_index, bwt_cmp_straight), :77-91 (bwt).
 *
 * The reference stable-sorts rotation start indices with a full cyclic lexicographic "<".
 * Rotations that compare equal (periodic input) keep ascending start index.  Restated as
 * cyclic prefix doubling with SPARSE ranks: rank[i] = number of rotations strictly smaller
 * than rotation i under the current prefix length; equal prefixes keep equal rank.  After the
 * last round equal rank <=> equal rotation, the last column is independent of the order
 * inside an equal group (equal rotations end in the same byte), and rotation 0 -- the lowest
 * start index of its group -- sits at row rank[0]  (main.cpp:88).
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t key; uint32_t idx; } kv_t;

static void radix_sort_kv(kv_t *a, kv_t *tmp, size_t n)
{
    /* LSD, four 16-bit digits, constant digits skipped; stable. */
    size_t *hist = (size_t *)calloc(4 * 65536, sizeof(size_t));
    for (size_t i = 0; i < n; ++i) {
        uint64_t k = a[i].key;
        hist[0 * 65536 + (k & 0xffff)]++;
        hist[1 * 65536 + ((k >> 16) & 0xffff)]++;
        hist[2 * 65536 + ((k >> 32) & 0xffff)]++;
        hist[3 * 65536 + ((k >> 48) & 0xffff)]++;
    }
    kv_t *src = a, *dst = tmp;
    for (int d = 0; d < 4; ++d) {
        size_t *h = hist + (size_t)d * 65536;
        int shift = 16 * d;
        if (h[(src[0].key >> shift) & 0xffff] == n) continue; /* every key shares this digit */
        size_t sum = 0;
        for (int b = 0; b < 65536; ++b) { size_t c = h[b]; h[b] = sum; sum += c; }
        for (size_t i = 0; i < n; ++i) dst[h[(src[i].key >> shift) & 0xffff]++] = src[i];
        kv_t *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(kv_t));
    free(hist);
}

int orc_bwt(const uint8_t *in, size_t n, uint8_t *last_col, uint64_t *primary)
{
    if (n == 0) { *primary = 0; return 0; }
    if (n >= 0xffffffffull) return -1;
    uint32_t *rank = (uint32_t *)malloc(n * sizeof(uint32_t));
    kv_t *a = (kv_t *)malloc(n * sizeof(kv_t));
    kv_t *tmp = (kv_t *)malloc(n * sizeof(kv_t));
    if (!rank || !a || !tmp) { free(rank); free(a); free(tmp); return -1; }

    /* round 0: prefix length 1; sparse rank of a byte = number of smaller bytes */
    size_t cnt[257]; memset(cnt, 0, sizeof cnt);
    for (size_t i = 0; i < n; ++i) cnt[in[i] + 1]++;
    for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
    size_t groups = 0;
    for (int b = 0; b < 256; ++b) if (cnt[b + 1] != cnt[b]) ++groups;
    for (size_t i = 0; i < n; ++i) rank[i] = (uint32_t)cnt[in[i]];
    /* a[] in sorted order for prefix 1 (only needed if the loop below never runs) */
    {
        size_t pos[256]; for (int b = 0; b < 256; ++b) pos[b] = cnt[b];
        for (size_t i = 0; i < n; ++i) { a[pos[in[i]]].idx = (uint32_t)i; a[pos[in[i]]].key = 0; pos[in[i]]++; }
    }
    for (size_t k = 1; groups < n && k < n; k *= 2) {
        for (size_t i = 0; i < n; ++i) {
            size_t j = i + k; if (j >= n) j %= n;
            a[i].key = ((uint64_t)rank[i] << 32) | rank[j];
            a[i].idx = (uint32_t)i;
        }
        radix_sort_kv(a, tmp, n);
        groups = 0;
        uint32_t cur = 0;
        for (size_t j = 0; j < n; ++j) {
            if (j == 0 || a[j].key != a[j - 1].key) { cur = (uint32_t)j; ++groups; }
            rank[a[j].idx] = cur;
        }
    }
    for (size_t j = 0; j < n; ++j) {
        size_t s = a[j].idx;
        last_col[j] = in[s == 0 ? n - 1 : s - 1];      /* main.cpp:87 */
    }
    *primary = rank[0];                                  /* main.cpp:88 */
    free(rank); free(a); free(tmp);
    return 0;
}

/* Inverse BWT.  Reference: main.cpp:28-36 (bwt_cmp_reverse), :61-75 (bwt_reverse).
 * stable_sort of positions by byte == counting sort -> l_shift; then the N-step walk. */
int orc_ibwt(const uint8_t *last_col, size_t n, uint64_t primary, uint8_t *out)
{
    if (n == 0) return 0;
    if (primary >= n) return -1;
    size_t *t = (size_t *)malloc(n * sizeof(size_t));
    if (!t) return -1;
    size_t cnt[257]; memset(cnt, 0, sizeof cnt);
    for (size_t i = 0; i < n; ++i) cnt[last_col[i] + 1]++;
    for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
    for (size_t i = 0; i < n; ++i) t[cnt[last_col[i]]++] = i;   /* main.cpp:67 */
    size_t row = (size_t)primary;
    for (size_t i = 0; i < n; ++i) {                              /* main.cpp:70-73 */
        out[i] = last_col[t[row]];
        row = t[row];
    }
    free(t);
    return 0;
}

/* Move-to-front.  Reference: main.cpp:93-112. */
void orc_mtf(const uint8_t *in, size_t n, uint8_t *out)
{
    uint8_t list[256];
    for (int i = 0; i < 256; ++i) list[i] = (uint8_t)i;
    for (size_t i = 0; i < n; ++i) {
        uint8_t c = in[i];
        int p = 0;
        while (list[p] != c) ++p;
        out[i] = (uint8_t)p;
        if (p) { memmove(list + 1, list, (size_t)p); list[0] = c; }
    }
}

/* Inverse move-to-front.  Reference: main.cpp:114-130. */
void orc_imtf(const uint8_t *in, size_t n, uint8_t *out)
{
    uint8_t list[256];
    for (int i = 0; i < 256; ++i) list[i] = (uint8_t)i;
    for (size_t i = 0; i < n; ++i) {
        int p = in[i];
        uint8_t c = list[p];
        out[i] = c;
        if (p) { memmove(list + 1, list, (size_t)p); list[0] = c; }
    }
}

/* ------------------------------------------------------------------------------------------
 * Huffman model.  Reference: main.cpp:229-254.
 *
 * The heap holds std::pair<long, BTree*> = (-freq, pointer) and is a max-heap: top() is the
 * smallest frequency and, among equal frequencies, the LARGEST pointer.  Pointer order in a
 * one-shot process follows SURVEY App. B.2 (glibc 2.39 allocator geometry), expressed on the
 * node creation index k.  addr_rank() returns the position of node k in ascending address
 * order.
 * ------------------------------------------------------------------------------------------ */
static int addr_rank(int k, int n_leaves)
{
    if (n_leaves <= 128) {
        /* 1 < 3 < 4 < ... < 127 < 0 < 2 < 128 < 129 < ... */
        if (k == 1) return 0;
        if (k >= 3 && k <= 127) return k - 2;
        if (k == 0) return 126;
        if (k == 2) return 127;
        return k;
    }
    /* 1 < 3..64 < 129..192 < 65..127 < 0 < 2 < 128 < 193 < ... */
    if (k == 1) return 0;
    if (k >= 3 && k <= 64) return k - 2;           /* 1..62   */
    if (k >= 129 && k <= 192) return 63 + (k - 129); /* 63..126 */
    if (k >= 65 && k <= 127) return 127 + (k - 65);  /* 127..189 */
    if (k == 0) return 190;
    if (k == 2) return 191;
    if (k == 128) return 192;
    return k;
}

int orc_huff_build_from_hist(const uint64_t freq[256], const uint8_t *order, int n_leaves, orc_tree *t)
{
    if (n_leaves <= 0 || n_leaves > 256) return -1;
    memset(t, 0, sizeof *t);
    t->n_leaves = n_leaves;
    uint64_t w[511];
    int alive[511];
    int n = 0;
    for (int k = 0; k < n_leaves; ++k) {            /* main.cpp:238-244 */
        t->left[n] = t->right[n] = -1;
        t->value[n] = order[k];
        w[n] = freq[order[k]];
        alive[n] = 1;
        ++n;
    }
    int live = n_leaves;
    while (live > 1) {                              /* main.cpp:245-254 */
        int pick[2];
        for (int r = 0; r < 2; ++r) {
            int best = -1;
            for (int k = 0; k < n; ++k) {
                if (!alive[k]) continue;
                if (best < 0 || w[k] < w[best] ||
                    (w[k] == w[best] && addr_rank(k, n_leaves) > addr_rank(best, n_leaves)))
                    best = k;
            }
            alive[best] = 0;
            pick[r] = best;
        }
        t->left[n] = pick[0];                        /* first pop  -> left  (main.cpp:246,252) */
        t->right[n] = pick[1];                       /* second pop -> right (main.cpp:248,252) */
        t->value[n] = 0;
        w[n] = w[pick[0]] + w[pick[1]];
        alive[n] = 1;
        ++n;
        --live;
    }
    t->n_nodes = n;
    t->root = n - 1;
    return 0;
}

int orc_huff_build(const uint8_t *mtf, size_t n, orc_tree *t)
{
    if (n == 0) return -1;
    uint64_t freq[256]; memset(freq, 0, sizeof freq);
    uint8_t order[256]; int seen[256]; memset(seen, 0, sizeof seen);
    int n_leaves = 0;
    for (size_t i = 0; i < n; ++i) freq[mtf[i]]++;                       /* main.cpp:235-237 */
    for (size_t i = 0; i < n && n_leaves < 256; ++i)                      /* main.cpp:238-244 */
        if (!seen[mtf[i]]) { seen[mtf[i]] = 1; order[n_leaves++] = mtf[i]; }
    return orc_huff_build_from_hist(freq, order, n_leaves, t);
}

/* Code table.  Reference: main.cpp:132-156. */
typedef struct { uint8_t bits[256][256]; int len[256]; } code_table;

static void codes_rec(const orc_tree *t, int node, uint8_t *path, int depth, code_table *ct)
{
    if (t->left[node] < 0) {                        /* has_value(): main.cpp:19-21,137-140 */
        ct->len[t->value[node]] = depth;
        memcpy(ct->bits[t->value[node]], path, (size_t)depth);
        return;
    }
    path[depth] = 0; codes_rec(t, t->left[node], path, depth + 1, ct);   /* main.cpp:143,145 */
    path[depth] = 1; codes_rec(t, t->right[node], path, depth + 1, ct);  /* main.cpp:144,146 */
}

static void build_codes(const orc_tree *t, code_table *ct)
{
    uint8_t path[512];
    for (int i = 0; i < 256; ++i) ct->len[i] = -1;
    codes_rec(t, t->root, path, 0, ct);
}

void orc_codes(const orc_tree *t, uint64_t code_lo[256], uint64_t code_hi[256], int len[256])
{
    code_table *ct = (code_table *)malloc(sizeof *ct);
    build_codes(t, ct);
    for (int s = 0; s < 256; ++s) {
        len[s] = ct->len[s];
        unsigned __int128 v = 0;
        for (int i = 0; i < ct->len[s] && i < 128; ++i) v = (v << 1) | ct->bits[s][i];
        code_lo[s] = (uint64_t)v; code_hi[s] = (uint64_t)(v >> 64);
    }
    free(ct);
}

/* MSB-first bit appender.  Reference: io_utilities.h:87-94 (append_bit): the buffer starts as
 * one zero byte; a new byte is pushed only when a bit arrives and the current byte is full. */
typedef struct { uint8_t *buf; size_t cap; size_t size; int free_bit; int overflow; } bitw;

static void bw_init(bitw *w, uint8_t *buf, size_t cap)
{
    w->buf = buf; w->cap = cap; w->size = 1; w->free_bit = 7; w->overflow = cap < 1;
    if (cap) buf[0] = 0;
}
static void bw_bit(bitw *w, int bit)
{
    if (w->free_bit < 0) {
        if (w->size >= w->cap) { w->overflow = 1; return; }
        w->buf[w->size++] = 0; w->free_bit = 7;
    }
    if (!w->overflow) w->buf[w->size - 1] |= (uint8_t)(bit << w->free_bit);
    w->free_bit--;
}

/* Tree serialisation.  Reference: main.cpp:174-196, io_utilities.h:96-101 (append_byte). */
static void tree_ser_rec(const orc_tree *t, int node, bitw *w)
{
    if (t->left[node] < 0) {
        bw_bit(w, 0);
        for (int i = 7; i >= 0; --i) bw_bit(w, (t->value[node] >> i) & 1);
        return;
    }
    bw_bit(w, 1);
    tree_ser_rec(t, t->left[node], w);
    tree_ser_rec(t, t->right[node], w);
}

size_t orc_tree_to_bytes(const orc_tree *t, uint8_t *out)
{
    bitw w; bw_init(&w, out, 320);
    tree_ser_rec(t, t->root, &w);
    return w.size;
}

/* Tree parse.  Reference: main.cpp:198-227, io_utilities.h:57-85 (read_bit/read_byte). */
typedef struct { const uint8_t *p; size_t nbits; size_t pos; int err; } bitr;
static int br_bit(bitr *r)
{
    if (r->pos >= r->nbits) { r->err = 1; return 0; }
    int b = (r->p[r->pos >> 3] >> (7 - (r->pos & 7))) & 1;
    r->pos++;
    return b;
}
static int tree_parse_rec(bitr *r, orc_tree *t, int depth)
{
    if (r->err || t->n_nodes >= 511 || depth > 256) { r->err = 1; return -1; }
    int id = t->n_nodes++;
    if (!br_bit(r)) {
        int v = 0;
        for (int i = 0; i < 8; ++i) v = (v << 1) | br_bit(r);
        t->left[id] = t->right[id] = -1; t->value[id] = (uint8_t)v; t->n_leaves++;
        return id;
    }
    t->value[id] = 0;
    int l = tree_parse_rec(r, t, depth + 1);
    int rr = tree_parse_rec(r, t, depth + 1);
    t->left[id] = l; t->right[id] = rr;
    return id;
}
int orc_bytes_to_tree(const uint8_t *bytes, size_t nbytes, orc_tree *t)
{
    memset(t, 0, sizeof *t);
    bitr r = { bytes, nbytes * 8, 0, 0 };
    t->root = tree_parse_rec(&r, t, 0);   /* note: ids here are pre-order, not creation order */
    return r.err ? -1 : 0;
}

/* Encode.  Reference: main.cpp:158-172. */
size_t orc_huff_encode(const uint8_t *mtf, size_t n, const orc_tree *t, uint8_t *out, size_t cap)
{
    code_table *ct = (code_table *)malloc(sizeof *ct);
    build_codes(t, ct);
    bitw w; bw_init(&w, out, cap);
    for (size_t i = 0; i < n; ++i) {
        int s = mtf[i];
        for (int b = 0; b < ct->len[s]; ++b) bw_bit(&w, ct->bits[s][b]);
    }
    free(ct);
    return w.overflow ? 0 : w.size;
}

/* Decode.  Reference: main.cpp:259-281: grow a bit vector until it is a code word; that is a
 * root-to-leaf walk.  A single-leaf tree has the empty code and consumes no bits. */
int orc_huff_decode(const uint8_t *payload, size_t payload_len, const orc_tree *t, size_t n, uint8_t *out)
{
    bitr r = { payload, payload_len * 8, 0, 0 };
    for (size_t i = 0; i < n; ++i) {
        int node = t->root;
        while (t->left[node] >= 0) {
            node = br_bit(&r) ? t->right[node] : t->left[node];
            if (r.err) return -1;
        }
        out[i] = t->value[node];
    }
    return 0;
}

/* Container.  Reference: io_utilities.h:7-27 (write_bytes), :29-55 (read_bytes). */
static void put_u64(uint8_t *p, uint64_t v) { for (int i = 0; i < 8; ++i) p[i] = (uint8_t)(v >> (8 * i)); }
static uint64_t get_u64(const uint8_t *p) { uint64_t v = 0; for (int i = 7; i >= 0; --i) v = (v << 8) | p[i]; return v; }

uint64_t orc_decompressed_size(const uint8_t *file, size_t len) { return len < 24 ? 0 : get_u64(file + 8); }

/* Reference: main.cpp:300-325. */
int orc_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len)
{
    if (n == 0) return -1;     /* the reference segfaults here (main.cpp:245-246); we refuse */
    uint8_t *l = (uint8_t *)malloc(n), *m = (uint8_t *)malloc(n);
    if (!l || !m) { free(l); free(m); return -1; }
    uint64_t primary = 0;
    int rc = orc_bwt(in, n, l, &primary);
    if (rc) { free(l); free(m); return rc; }
    orc_mtf(l, n, m);
    orc_tree t;
    orc_huff_build(m, n, &t);
    uint8_t tree_bytes[320];
    size_t tb = orc_tree_to_bytes(&t, tree_bytes);
    if (cap < 24 + tb + 1) { free(l); free(m); return -2; }
    put_u64(out, primary); put_u64(out + 8, n); put_u64(out + 16, tb);
    memcpy(out + 24, tree_bytes, tb);
    size_t pl = orc_huff_encode(m, n, &t, out + 24 + tb, cap - 24 - tb);
    free(l); free(m);
    if (!pl) return -2;
    *out_len = 24 + tb + pl;
    return 0;
}

/* Reference: main.cpp:327-345. */
int orc_decompress(const uint8_t *in, size_t in_len, uint8_t *out, size_t cap, size_t *out_len)
{
    if (in_len < 24) return -1;
    uint64_t primary = get_u64(in), n = get_u64(in + 8), tb = get_u64(in + 16);
    if (tb > in_len - 24 || n > cap) return -1;
    orc_tree t;
    if (orc_bytes_to_tree(in + 24, (size_t)tb, &t)) return -1;
    if (n == 0) { *out_len = 0; return 0; }
    uint8_t *m = (uint8_t *)malloc(n), *l = (uint8_t *)malloc(n);
    if (!m || !l) { free(m); free(l); return -1; }
    int rc = orc_huff_decode(in + 24 + tb, in_len - 24 - (size_t)tb, &t, (size_t)n, m);
    if (!rc) { orc_imtf(m, (size_t)n, l); rc = orc_ibwt(l, (size_t)n, primary, out); }
    free(m); free(l);
    if (rc) return -1;
    *out_len = (size_t)n;
    return 0;
}
