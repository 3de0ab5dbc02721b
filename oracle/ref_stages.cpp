// ref_stages.cpp -- in-process stage harness around the UNMODIFIED reference translation unit.
//
// TEST INFRASTRUCTURE ONLY.  This file contains no reference code: it #includes the reference's
// main.cpp where it lies (REF_DIR, default /root/reference) at compile time and exposes its stage
// functions under a C ABI so tests can pin the oracle restatement stage by stage.
// Built by oracle/Makefile into oracle/_ref/libref_stages.so.  Do NOT use huffman() from here for
// tree-byte goldens: its tie-breaking depends on heap history; byte goldens come
// from the one-shot ref_compress binary.
#include <climits>
#include <cstdint>
#include <algorithm>
#include <tuple>
#include <string>
#include <cstring>
#define main ref_main
#include "main.cpp"
#undef main

extern "C" {

// main.cpp:77-91
int ref_bwt(const unsigned char *in, size_t n, unsigned char *out, uint64_t *primary) {
    std::vector<unsigned char> v(in, in + n);
    auto r = bwt(v);
    *primary = r.first;
    std::memcpy(out, r.second.data(), n);
    return 0;
}
// main.cpp:61-75
int ref_ibwt(const unsigned char *in, size_t n, uint64_t primary, unsigned char *out) {
    std::vector<unsigned char> v(in, in + n);
    auto r = bwt_reverse(v, primary);
    std::memcpy(out, r.data(), n);
    return 0;
}
// main.cpp:93-112
int ref_mtf(const unsigned char *in, size_t n, unsigned char *out) {
    std::vector<unsigned char> v(in, in + n);
    auto r = move_to_front(v);
    std::memcpy(out, r.data(), n);
    return 0;
}
// main.cpp:114-130
int ref_imtf(const unsigned char *in, size_t n, unsigned char *out) {
    std::vector<unsigned char> v(in, in + n);
    auto r = move_to_front_reverse(v);
    std::memcpy(out, r.data(), n);
    return 0;
}
// main.cpp:198-227, 132-156, 283-292, 259-281: parse a serialised tree and decode n symbols
int ref_huff_decode(const unsigned char *tree, size_t tree_len, const unsigned char *payload, size_t payload_len,
                    size_t n, unsigned char *out) {
    std::vector<unsigned char> tv(tree, tree + tree_len), pv(payload, payload + payload_len);
    auto root = bytes_to_tree(tv);
    auto cw = build_hashmap(root);
    auto rcw = reverse_map(cw);
    auto r = huffman_reverse(pv, rcw, n);
    std::memcpy(out, r.data(), n);
    return 0;
}
// main.cpp:198-227, 158-172: parse a serialised tree and encode with it (tie-break independent)
size_t ref_huff_encode_with_tree(const unsigned char *tree, size_t tree_len, const unsigned char *in, size_t n,
                                 unsigned char *out, size_t cap) {
    std::vector<unsigned char> tv(tree, tree + tree_len), v(in, in + n);
    auto root = bytes_to_tree(tv);
    auto r = encode_with_huffman(v, root);
    if (r.size() > cap) return 0;
    std::memcpy(out, r.data(), r.size());
    return r.size();
}
// main.cpp:198-227 then :189-196: parse + re-serialise (round-trips the tree format)
size_t ref_tree_roundtrip(const unsigned char *tree, size_t tree_len, unsigned char *out, size_t cap) {
    std::vector<unsigned char> tv(tree, tree + tree_len);
    auto root = bytes_to_tree(tv);
    auto r = tree_to_bytes(root);
    if (r.size() > cap) return 0;
    std::memcpy(out, r.data(), r.size());
    return r.size();
}

}
