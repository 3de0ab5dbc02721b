/*
 * bzap_oracle.h -- CPU restatement of komour/bwt-mtf-huffman-compressor's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (libbzap.so, the CLI drivers, the
 * python package) may include, link, load or execute this.  Only tests/, the
 * __graft_entry__.smoke() check and bench.py's cpu_baseline / --impl reference leg use it,
 * and only as the checker / the reported CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function below against
 * the unmodified reference compiled into oracle/_ref/ (one-shot ref_compress / ref_decompress
 * binaries and the in-process stage harness libref_stages.so), on all 14 Calgary files and
 * the periodic KATs of SURVEY.md App. C; tests/golden/ holds the committed vectors.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 */
#ifndef BZAP_ORACLE_H
#define BZAP_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* main.cpp:46-59 (bwt_cmp_straight) + main.cpp:77-91 (bwt): last column of the stably sorted
 * rotation matrix and the row of rotation 0.  Returns 0, or -1 on allocation failure. */
int orc_bwt(const uint8_t *in, size_t n, uint8_t *last_col, uint64_t *primary);
/* main.cpp:61-75 (bwt_reverse). */
int orc_ibwt(const uint8_t *last_col, size_t n, uint64_t primary, uint8_t *out);
/* main.cpp:93-112 (move_to_front) and :114-130 (move_to_front_reverse). */
void orc_mtf(const uint8_t *in, size_t n, uint8_t *out);
void orc_imtf(const uint8_t *in, size_t n, uint8_t *out);

/* Huffman tree in array form.  Node ids are creation indices: leaves
 * 0..n_leaves-1 in order of first appearance, internal nodes n_leaves..2*n_leaves-2. */
typedef struct {
    int n_leaves;
    int n_nodes;
    int root;
    int left[511], right[511]; /* -1 for leaves */
    uint8_t value[511];
} orc_tree;

/* main.cpp:229-254 (huffman: histogram, first-appearance leaf order, merge loop) with the
 * pointer-order tie-break of std::priority_queue<std::pair<long,BTree*>> restated by the
 * allocator address law of SURVEY App. B.2.  Returns 0, -1 if the stream is empty. */
int orc_huff_build(const uint8_t *mtf, size_t n, orc_tree *t);
/* same, from a histogram and the first-appearance order (order[k] = symbol of leaf k). */
int orc_huff_build_from_hist(const uint64_t freq[256], const uint8_t *order, int n_leaves, orc_tree *t);
/* main.cpp:132-156 (traverse/build_hashmap): code = root-to-leaf path, left 0, right 1.
 * code_hi:code_lo hold the path MSB-first in the low `len` bits of a 128-bit value. */
void orc_codes(const orc_tree *t, uint64_t code_lo[256], uint64_t code_hi[256], int len[256]);
/* main.cpp:174-196 (dfs/tree_to_bytes) + io_utilities.h:87-101.  out needs 320 bytes. */
size_t orc_tree_to_bytes(const orc_tree *t, uint8_t *out);
/* main.cpp:198-227 (bytes_to_tree).  Returns 0, -1 on a malformed serialisation. */
int orc_bytes_to_tree(const uint8_t *bytes, size_t nbytes, orc_tree *t);
/* main.cpp:158-172 (encode_with_huffman): returns payload size max(1, ceil(bits/8)). */
size_t orc_huff_encode(const uint8_t *mtf, size_t n, const orc_tree *t, uint8_t *out, size_t cap);
/* main.cpp:259-281 (huffman_reverse): decodes exactly n symbols.  Returns 0 / -1. */
int orc_huff_decode(const uint8_t *payload, size_t payload_len, const orc_tree *t, size_t n, uint8_t *out);

/* main.cpp:300-325 (compress) + io_utilities.h:7-27 (write_bytes), buffer level. */
int orc_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len);
/* main.cpp:327-345 (decompress) + io_utilities.h:29-55 (read_bytes). out needs N bytes. */
int orc_decompress(const uint8_t *in, size_t in_len, uint8_t *out, size_t cap, size_t *out_len);
/* header field 2 (io_utilities.h:46) */
uint64_t orc_decompressed_size(const uint8_t *file, size_t len);

#ifdef __cplusplus
}
#endif
#endif
