// Test-only shim: the product's column parsing (vcf-compression_b200/csrc/vcfc_index.cuh, the code k_index_lines runs on
// the device and the indexed query runs on the host) compiled for the HOST, so the CPU suite can compare it with the
// oracle line by line.  Mirrors the body of k_index_lines.  Built by tests/test_abi.py with g++ (-D__host__= -D__device__=); not part of the library.
#include <stddef.h>
#include <stdint.h>

#include "../vcf-compression_b200/csrc/vcfc_index.cuh"

extern "C" int shim_line_index_fields(const uint8_t* in, size_t n, long long* end_out, uint8_t* ref_out) {
    const uint8_t* f[8];
    int fl[8];
    size_t q = 0;
    for (int c = 0; c < 8; c++) {
        size_t e = q;
        while (e < n && in[e] != '\t') e++;
        if (e >= n) return 2;
        f[c] = in + q; fl[c] = (int)(e - q);
        q = e + 1;
    }
    long long pos = 0, endp = 0;
    bool ok = vcfc::idx::parse_ul(f[1], fl[1], &pos);
    if (ok) ok = vcfc::idx::line_end_position(pos, fl[3], f[4], fl[4], f[7], fl[7], &endp);
    *end_out = ok ? endp : 0;
    *ref_out = vcfc::idx::ref_name_index(f[0], fl[0]);
    return ok ? 0 : 1;
}
