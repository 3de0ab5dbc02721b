// huffman_host.cpp -- host half of the Huffman stage: the <=511-node model and its tables.
//
// The tree is tiny (<=256 leaves) and its shape depends on a tie-break that is a property of the
// reference PROCESS, not of the data, so it is built on the host between two GPU kernels:
//   histogram (GPU) -> tree + code table (here) -> bit packing (GPU).
// Reference: huffman() main.cpp:229-254, traverse()/build_hashmap() :132-156, tree_to_bytes()
// :174-196, bytes_to_tree() :198-227.
#include "bzap_internal.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <queue>
#include <string>
#include <utility>
#include <vector>

// ---- tie-break --------------------------------------------------------------------------------------
// std::priority_queue<std::pair<long, BTree*>> (main.cpp:232) pops the smallest frequency and, among
// equal frequencies, the LARGEST node address.  In the reference's one-shot COMPRESS process the
// addresses returned by `new BTree` follow a fixed pattern in the node creation index k
// (leaves 0..n-1 by first appearance, then internal nodes):
//   n <= 128 leaves:  1 < 3 < 4 < ... < 127 < 0 < 2 < 128 < 129 < ...
//   n  > 128 leaves:  1 < 3..64 < 129..192 < 65..127 < 0 < 2 < 128 < 193 < ...
// address_order() returns the position of node k in that ascending-address order.
static int address_order(int k, bool many_leaves)
{
    if (k == 1) return 0;
    if (!many_leaves) {
        if (k >= 3 && k <= 127) return k - 2;
        if (k == 0) return 126;
        if (k == 2) return 127;
        return k;
    }
    if (k >= 3 && k <= 64) return k - 2;
    if (k >= 129 && k <= 192) return k - 129 + 63;
    if (k >= 65 && k <= 127) return k - 65 + 127;
    if (k == 0) return 190;
    if (k == 2) return 191;
    if (k == 128) return 192;
    return k;
}

// Outside the windows where that law is validated (N < 64,600) the node addresses depend on
// where the chunks freed while the input was read lie -- a function of N and the number of leaves only, because
// the reference's allocation script up to the last `new BTree` does not depend on the data.  With
// BZAP_HEAP_REPLAY=1 the order comes from bzap_heap_replay (csrc/heap_replay.c), a helper that replays that
// script against this host's allocator in a fresh process: byte identity with the one-shot reference binary for
// every N (tests/test_heap_replay.py checks it in every failure window).  Opt-in: it costs a process spawn per
// (N, leaves) pair (cached), and the test oracle restates the closed-form law.
static bool replay_order(u64 n, int n_leaves, std::vector<int> *ranks)
{
    static std::mutex mu;
    static std::map<std::pair<u64, int>, std::vector<int>> cache;
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(std::make_pair(n, n_leaves));
    if (it != cache.end()) { *ranks = it->second; return !ranks->empty(); }
    std::vector<int> r;
    Dl_info info;
    if (dladdr((const void *)&huff_total_bits, &info) && info.dli_fname) {
        std::string dir(info.dli_fname);
        const size_t slash = dir.rfind('/');
        dir = slash == std::string::npos ? std::string(".") : dir.substr(0, slash);
        char cmd[4200];
        snprintf(cmd, sizeof cmd, "'%s/bzap_heap_replay' %llu %d", dir.c_str(), (unsigned long long)n, n_leaves);
        if (FILE *p = popen(cmd, "r")) {
            int v;
            while (fscanf(p, "%d", &v) == 1) r.push_back(v);
            if (pclose(p) != 0 || (int)r.size() != 2 * n_leaves - 1) r.clear();
        }
    }
    cache[std::make_pair(n, n_leaves)] = r;          // an empty entry remembers the failure: fall back to the law
    *ranks = r;
    return !r.empty();
}

int huff_build_tree(const u64 freq[256], const u8 *order, int n_leaves, bzap_tree *t)
{
    if (!freq || !order || !t || n_leaves < 1 || n_leaves > 256) return BZAP_ERR_ARG;
    std::memset(t, 0, sizeof *t);
    t->n_leaves = n_leaves;
    const bool many = n_leaves > 128;
    std::vector<int> replayed;
    {
        u64 n = 0;
        for (int k = 0; k < n_leaves; ++k) n += freq[order[k]];
        const char *e = getenv("BZAP_HEAP_REPLAY");
        if (e && e[0] == '1' && n < 64600) replay_order(n, n_leaves, &replayed);
    }
    auto address_of = [&](int k) { return replayed.empty() ? address_order(k, many) : replayed[k]; };
    // min-heap on (weight, -address_order): smallest weight first, highest address first
    typedef std::pair<std::pair<u64, int>, int> Item;   // ((weight, -addr), node)
    std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
    int next = 0;
    for (int k = 0; k < n_leaves; ++k, ++next) {
        t->left[next] = t->right[next] = -1;
        t->value[next] = order[k];
        heap.push(Item(std::make_pair(freq[order[k]], -address_of(next)), next));
    }
    while (heap.size() > 1) {
        Item a = heap.top(); heap.pop();     // first pop  -> left  (main.cpp:246, 252)
        Item b = heap.top(); heap.pop();     // second pop -> right (main.cpp:248, 252)
        t->left[next] = a.second;
        t->right[next] = b.second;
        t->value[next] = 0;
        heap.push(Item(std::make_pair(a.first.first + b.first.first, -address_of(next)), next));
        ++next;
    }
    t->n_nodes = next;
    t->root = next - 1;
    return BZAP_OK;
}

static bool tree_is_sane(const bzap_tree *t)
{
    if (!t || t->n_nodes < 1 || t->n_nodes > 511 || t->root < 0 || t->root >= t->n_nodes) return false;
    for (int i = 0; i < t->n_nodes; ++i) {
        bool leaf = t->left[i] < 0;
        if (leaf != (t->right[i] < 0)) return false;
        if (!leaf && (t->left[i] >= t->n_nodes || t->right[i] >= t->n_nodes)) return false;
    }
    return true;
}

// ---- code table: root-to-leaf paths, left = 0, right = 1 (main.cpp:141-146) -------------------------
int huff_code_table(const bzap_tree *t, CodeTable *ct)
{
    if (!tree_is_sane(t) || !ct) return BZAP_ERR_ARG;
    std::memset(ct, 0, sizeof *ct);
    struct Frame { int node; u64 code; int len; };
    std::vector<Frame> stack;
    stack.push_back({t->root, 0, 0});
    int visited = 0;
    while (!stack.empty()) {
        Frame f = stack.back();
        stack.pop_back();
        if (++visited > 511) return BZAP_ERR_CORRUPT;
        if (t->left[f.node] < 0) {
            if (f.len > 64) return BZAP_ERR_TOO_LARGE;
            ct->code[t->value[f.node]] = f.code;
            ct->len[t->value[f.node]] = (u8)f.len;
            ct->max_len = std::max(ct->max_len, f.len);
            continue;
        }
        u64 c = f.len < 64 ? f.code << 1 : 0;
        stack.push_back({t->right[f.node], c | 1, f.len + 1});
        stack.push_back({t->left[f.node], c, f.len + 1});
    }
    return BZAP_OK;
}

u64 huff_total_bits(const u64 freq[256], const CodeTable *ct)
{
    u64 bits = 0;
    for (int s = 0; s < 256; ++s) bits += freq[s] * ct->len[s];
    return bits;
}

// ---- serialisation: pre-order, leaf = 0 + 8 value bits, internal = 1, MSB first (main.cpp:174-196) ---
int huff_tree_to_bytes(const bzap_tree *t, u8 *out, size_t *len)
{
    if (!tree_is_sane(t) || !out || !len) return BZAP_ERR_ARG;
    std::memset(out, 0, BZAP_MAX_TREE_BYTES);
    size_t bit = 0;
    auto put = [&](int b) {
        if (b) out[bit >> 3] |= (u8)(0x80u >> (bit & 7));
        ++bit;
    };
    std::vector<int> stack(1, t->root);
    int visited = 0;
    while (!stack.empty()) {
        int node = stack.back();
        stack.pop_back();
        if (++visited > 511 || bit + 9 > 8 * BZAP_MAX_TREE_BYTES) return BZAP_ERR_CORRUPT;
        if (t->left[node] < 0) {
            put(0);
            for (int i = 7; i >= 0; --i) put((t->value[node] >> i) & 1);
        } else {
            put(1);
            stack.push_back(t->right[node]);
            stack.push_back(t->left[node]);
        }
    }
    // the reference's bit buffer starts as one zero byte and grows only when a bit needs a new
    // byte (io_utilities.h:87-94): size = max(1, ceil(bits/8))
    *len = std::max<size_t>(1, (bit + 7) / 8);
    return BZAP_OK;
}

int huff_bytes_to_tree(const u8 *bytes, size_t len, bzap_tree *t)
{
    if (!bytes || !t) return BZAP_ERR_ARG;
    std::memset(t, 0, sizeof *t);
    const size_t nbits = len * 8;
    size_t bit = 0;
    bool bad = false;
    auto get = [&]() -> int {
        if (bit >= nbits) { bad = true; return 0; }
        int b = (bytes[bit >> 3] >> (7 - (bit & 7))) & 1;
        ++bit;
        return b;
    };
    // iterative pre-order parse; pending[] holds internal nodes still missing children
    struct Pending { int node; int have; };
    std::vector<Pending> pending;
    int n = 0;
    t->root = 0;
    while (true) {
        if (n >= 511) return BZAP_ERR_CORRUPT;
        int id = n++;
        int internal = get();
        if (bad) return BZAP_ERR_CORRUPT;
        if (!pending.empty()) {
            Pending &p = pending.back();
            if (p.have == 0) t->left[p.node] = id; else t->right[p.node] = id;
            ++p.have;
        }
        if (internal) {
            t->left[id] = t->right[id] = -1;
            pending.push_back({id, 0});
        } else {
            int v = 0;
            for (int i = 0; i < 8; ++i) v = (v << 1) | get();
            if (bad) return BZAP_ERR_CORRUPT;
            t->left[id] = t->right[id] = -1;
            t->value[id] = (u8)v;
            ++t->n_leaves;
            while (!pending.empty() && pending.back().have == 2) pending.pop_back();
        }
        while (!pending.empty() && pending.back().have == 2) pending.pop_back();
        if (pending.empty()) break;
        if ((int)pending.size() > 256) return BZAP_ERR_CORRUPT;
    }
    t->n_nodes = n;
    if (t->n_leaves > 256) return BZAP_ERR_CORRUPT;
    return BZAP_OK;
}

// ---- decode tables --------------------------------------------------------------------------------------
// u16 entry: bit15 = leaf; leaf: bits 0-7 symbol, bits 8-11 = code bits consumed - 1;
// inner (tree deeper than the table): bits 0-14 = index of the 256-entry sub table to continue in
// (all bits of this table's width are consumed).
namespace {
struct DecBuilder {
    const bzap_tree *t;
    std::vector<u16> e;
    int sub_of_node[511];
    void fill(size_t base, int node, int width)
    {
        for (u32 pat = 0; pat < (1u << width); ++pat) {
            int cur = node, used = 0;
            while (t->left[cur] >= 0 && used < width) {
                cur = ((pat >> (width - 1 - used)) & 1) ? t->right[cur] : t->left[cur];
                ++used;
            }
            if (t->left[cur] < 0) {
                e[base + pat] = (u16)(0x8000u | ((u32)(used - 1) << 8) | t->value[cur]);
            } else {
                if (sub_of_node[cur] < 0) {
                    sub_of_node[cur] = (int)((e.size() - (1u << DEC_PRIMARY_BITS)) >> DEC_SUB_BITS);
                    size_t nb = e.size();
                    e.resize(nb + (1u << DEC_SUB_BITS));
                    fill(nb, cur, DEC_SUB_BITS);
                }
                e[base + pat] = (u16)sub_of_node[cur];
            }
        }
    }
};
}   // namespace

int huff_decode_tables(const bzap_tree *t, DecodeTables *dt)
{
    if (!tree_is_sane(t) || !dt) return BZAP_ERR_ARG;
    std::memset(dt, 0, sizeof *dt);
    // depth (may exceed 64 for hostile files: the tables do not care)
    {
        std::vector<std::pair<int, int>> st(1, std::make_pair(t->root, 0));
        int visited = 0;
        while (!st.empty()) {
            std::pair<int, int> f = st.back();
            st.pop_back();
            if (++visited > 511) return BZAP_ERR_CORRUPT;
            if (t->left[f.first] < 0) { dt->max_len = std::max(dt->max_len, f.second); continue; }
            st.push_back(std::make_pair(t->left[f.first], f.second + 1));
            st.push_back(std::make_pair(t->right[f.first], f.second + 1));
        }
    }
    if (t->left[t->root] < 0) {
        dt->single_leaf = 1;
        dt->single_value = t->value[t->root];
        return BZAP_OK;
    }
    DecBuilder b;
    b.t = t;
    for (int i = 0; i < 511; ++i) b.sub_of_node[i] = -1;
    b.e.assign(1u << DEC_PRIMARY_BITS, 0);
    b.fill(0, t->root, DEC_PRIMARY_BITS);
    dt->n_entries = b.e.size();
    dt->entries = (u16 *)std::malloc(dt->n_entries * sizeof(u16));
    if (!dt->entries) return BZAP_ERR_NOMEM;
    std::memcpy(dt->entries, b.e.data(), dt->n_entries * sizeof(u16));
    return BZAP_OK;
}

void huff_free_decode_tables(DecodeTables *dt)
{
    if (dt && dt->entries) { std::free(dt->entries); dt->entries = nullptr; }
}
