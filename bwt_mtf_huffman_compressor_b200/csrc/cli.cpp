// cli.cpp -- command line drivers with the reference's argv contract (main.cpp:439-456):
//   bzap_compress   <input> <output>      (reference built with -DCOMPRESS)
//   bzap_decompress <input> <output>      (reference built with -DDECOMPRESS)
// Wrong argument count prints the reference's message (no newline) and returns 1
// (main.cpp:440-443).  compress prints the reference's metrics line (main.cpp:321, 402-413);
// decompress prints nothing.  Unlike the reference, failures return a non-zero status and a
// message on stderr instead of crashing.
#include "bzap.h"
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

static long file_size(const char *p)
{
    FILE *f = fopen(p, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long s = ftell(f);
    fclose(f);
    return s;
}

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
    // metrics line, same arithmetic and formatting as main.cpp:319-323 + print_metrics :402-413
    long initial = file_size(argv[1]), encoded = file_size(argv[2]);
    FILE *f = fopen(argv[2], "rb");
    unsigned char hdr[24] = {0};
    if (f) { if (fread(hdr, 1, 24, f) != 24) hdr[16] = 0; fclose(f); }
    unsigned long long tree_bytes = 0;
    for (int i = 7; i >= 0; --i) tree_bytes = (tree_bytes << 8) | hdr[16 + i];
    std::cout << "header size: " << double(tree_bytes + 24) << " $$ ";
    std::cout << "file_name: " << argv[2] << " $$ initial_data_size: " << initial
              << " $$ encoded_file_size: " << encoded
              << " $$ bits_avg: " << (8 * double(encoded)) / double(initial)
              << " $$ compress_rate = " << double(encoded) / double(initial) << std::endl;
#else
    int rc = bzap_decompress_file(nullptr, argv[1], argv[2]);
    if (rc != BZAP_OK) {
        std::cerr << "bzap_decompress: " << bzap_strerror(rc) << ": " << bzap_last_error(nullptr) << std::endl;
        return 2;
    }
#endif
    return 0;
}
