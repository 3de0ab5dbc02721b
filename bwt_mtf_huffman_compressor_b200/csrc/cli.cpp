// cli.cpp -- command line drivers with the reference's argv contract (main.cpp:439-456):
//   bzap_compress   <input> <output>      (reference built with -DCOMPRESS)
//   bzap_decompress <input> <output>      (reference built with -DDECOMPRESS)
// Wrong argument count prints the reference's message (no newline) and returns 1
// (main.cpp:440-443).  compress prints the reference's metrics line (main.cpp:321, 402-413);
// decompress prints nothing.  Unlike the reference, failures return a non-zero status and a
// message on stderr instead of crashing.
#include "bzap.h"
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#if defined(BZAP_CLI_COMPRESS) && !defined(BZAP_CLI_FULL)
static void metrics_line(const char *in, const char *out);
#endif
static long file_size(const char *p)
{
    FILE *f = fopen(p, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long s = ftell(f);
    fclose(f);
    return s;
}

#ifdef BZAP_CLI_FULL
// FULL_PIPELINE mode of the reference (main.cpp:416-438): the 14 Calgary files of calgarycorpus/
// are compressed, decompressed and compared; "k/14 <metrics line>success|fail" per file.
static bool same_file(const std::string &a, const std::string &b)
{
    FILE *fa = fopen(a.c_str(), "rb"), *fb = fopen(b.c_str(), "rb");
    bool ok = fa && fb;
    while (ok) {
        unsigned char ba[65536], bb[65536];
        size_t na = fread(ba, 1, sizeof ba, fa), nb = fread(bb, 1, sizeof bb, fb);
        if (na != nb || memcmp(ba, bb, na) != 0) ok = false;
        if (na == 0) break;
    }
    if (fa) fclose(fa);
    if (fb) fclose(fb);
    return ok;
}
static void metrics_line(const char *in, const char *out);
int main(int argc, char *argv[])
{
    std::string dir = argc > 1 ? std::string(argv[1]) + "/" : "calgarycorpus/";
    const char *files[] = {"bib", "book1", "book2", "geo", "news", "obj1", "obj2", "paper1", "paper2",
                           "pic", "progc", "progl", "progp", "trans"};
    int failures = 0;
    for (int k = 0; k < 14; ++k) {
        std::cout << k + 1 << "/" << 14 << ' ';
        std::string in = dir + files[k], enc = in + ".bzap", dec = in + ".decoded";
        int rc = bzap_compress_file(nullptr, in.c_str(), enc.c_str());
        if (rc == BZAP_OK) { metrics_line(in.c_str(), enc.c_str()); rc = bzap_decompress_file(nullptr, enc.c_str(), dec.c_str()); }
        bool ok = rc == BZAP_OK && same_file(in, dec);
        failures += !ok;
        std::cout << (ok ? "success" : "fail") << std::endl;
    }
    return failures ? 2 : 0;
}
#else
int main(int argc, char *argv[])
{
    if (argc != 3) {
        std::cout << "Wrong arguments. Pass only input and output file as parameters";
        return 1;
    }
#ifdef BZAP_CLI_COMPRESS
    int rc = bzap_compress_file(nullptr, argv[1], argv[2]);
    if (rc != BZAP_OK) {
        std::cerr << "bzap_compress: " << bzap_strerror(rc) << ": " << bzap_last_error(nullptr) << std::endl;
        return 2;
    }
    metrics_line(argv[1], argv[2]);
#else
    int rc = bzap_decompress_file(nullptr, argv[1], argv[2]);
    if (rc != BZAP_OK) {
        std::cerr << "bzap_decompress: " << bzap_strerror(rc) << ": " << bzap_last_error(nullptr) << std::endl;
        return 2;
    }
#endif
    return 0;
}
#endif

#if defined(BZAP_CLI_COMPRESS) || defined(BZAP_CLI_FULL)
// metrics line, same arithmetic and formatting as main.cpp:319-323 + print_metrics :402-413
static void metrics_line(const char *in, const char *out)
{
    long initial = file_size(in), encoded = file_size(out);
    FILE *f = fopen(out, "rb");
    unsigned char hdr[24] = {0};
    if (f) { if (fread(hdr, 1, 24, f) != 24) hdr[16] = 0; fclose(f); }
    unsigned long long tree_bytes = 0;
    for (int i = 7; i >= 0; --i) tree_bytes = (tree_bytes << 8) | hdr[16 + i];
    std::cout << "header size: " << double(tree_bytes + 24) << " $$ ";
    std::cout << "file_name: " << out << " $$ initial_data_size: " << initial
              << " $$ encoded_file_size: " << encoded
              << " $$ bits_avg: " << (8 * double(encoded)) / double(initial)
              << " $$ compress_rate = " << double(encoded) / double(initial) << std::endl;
}
#endif
