// vcfc_index.cu -- binned index (.vcfci): per-line END position and reference index on the device.
//
// Restates, does not copy, the per-line part of create_binned_index4 (/root/reference/src/main.cpp:1284-1637):
// CHROM, POS, REF, ALT and INFO are read from the line's required section, END comes from compute_end_position
// (main.cpp:763-852: INFO END / SVLEN for symbolic alleles, else POS + max(len(REF), longest ALT) - 1), the
// chromosome index from reference_name_map (utils.hpp:90-103).  One thread per compressed line; the line table
// is the decoder's (k_dec_walk ... k_dec_fill).  The bin rule itself (a sequential grow-or-open over the lines,
// main.cpp:1430-1470) runs on the host over 10 bytes per line: see vcfc_create_binned_index_file.
#include "vcfc_common.cuh"
#include "vcfc_internal.h"

namespace vcfc {
namespace idx {

// strtoul(s, &end, 10) with end == s + n, as str_to_uint64 / str_to_long use it (utils.cpp:152-175): leading
// white space, an optional sign, digits; saturates at ULONG_MAX; a '-' negates modulo 2^64.  Fields of 64 bytes
// and more, or with a NUL inside, fail (as in the test oracle).
__device__ bool parse_ul(const uint8_t* p, int n, long long* out) {
    if (n >= 64) return false;
    for (int i = 0; i < n; i++) if (p[i] == 0) return false;
    int i = 0;
    while (i < n && (p[i] == ' ' || (p[i] >= 9 && p[i] <= 13))) i++;
    bool neg = false;
    if (i < n && (p[i] == '+' || p[i] == '-')) { neg = p[i] == '-'; i++; }
    const int d0 = i;
    unsigned long long v = 0;
    bool sat = false;
    while (i < n && p[i] >= '0' && p[i] <= '9') {
        const unsigned long long d = (unsigned long long)(p[i] - '0');
        if (v > (0xFFFFFFFFFFFFFFFFull - d) / 10ull) sat = true; else v = v * 10ull + d;
        i++;
    }
    if (i == d0) return n == 0;                 // no digits: strtoul leaves end at the start of the string
    if (i != n) return false;
    if (sat) v = 0xFFFFFFFFFFFFFFFFull; else if (neg) v = 0ull - v;
    *out = (long long)v;
    return true;
}

__device__ uint8_t ref_name_index(const uint8_t* p, int n) {
    if (n == 1 && p[0] == 'X') return 23;
    if (n == 1 && p[0] == 'Y') return 24;
    if (n == 1 && p[0] == 'M') return 25;
    if (n == 1 && p[0] >= '1' && p[0] <= '9') return (uint8_t)(p[0] - '0');
    if (n == 2 && p[0] >= '1' && p[0] <= '2' && p[1] >= '0' && p[1] <= '9') {
        const int v = 10 * (p[0] - '0') + (p[1] - '0');
        return v <= 22 ? (uint8_t)v : 0;
    }
    return 0;
}

// value of `key` in a ';'-separated key=value list with parse_kvp's rules (main.cpp:737-757; empty terms are dropped
// by split_string, utils.cpp:95): 1 found, 0 absent, -1 malformed ("k=a=b", "=" alone), a later duplicate wins
__device__ int kvp_lookup(const uint8_t* info, int n, const char* key, int klen, const uint8_t** val, int* val_len) {
    int found = 0, i = 0;
    while (i < n) {
        int j = i;
        while (j < n && info[j] != ';') j++;
        if (j > i) {
            const uint8_t* part[2] = {nullptr, nullptr};
            int plen[2] = {0, 0}, np = 0, a = i;
            while (a < j) {
                int b = a;
                while (b < j && info[b] != '=') b++;
                if (b > a) {
                    if (np == 2) return -1;
                    part[np] = info + a; plen[np] = b - a; np++;
                }
                a = b + 1;
            }
            if (np == 0) return -1;
            bool eq = plen[0] == klen;
            for (int t = 0; eq && t < klen; t++) eq = part[0][t] == (uint8_t)key[t];
            if (eq) {
                found = 1;
                if (np == 2) { *val = part[1]; *val_len = plen[1]; } else { *val = part[0]; *val_len = 0; }
            }
        }
        i = j + 1;
    }
    return found;
}

// max over the ','-separated integers of v (empty terms dropped); absolute values when `absolute`
__device__ bool max_of_list(const uint8_t* v, int vl, bool absolute, long long* out) {
    long long m = 0;
    int i = 0;
    while (i < vl) {
        int j = i;
        while (j < vl && v[j] != ',') j++;
        if (j > i) {
            long long e;
            if (!parse_ul(v + i, j - i, &e)) return false;
            if (absolute && e < 0) e = -e;
            if (e > m) m = e;
        }
        i = j + 1;
    }
    *out = m;
    return true;
}

// err: 0 ok, 1 malformed (the reference throws), 2 ran off the end of the block while reading the columns
__global__ void k_index_lines(const uint8_t* __restrict__ in, unsigned long long n, const unsigned long long* __restrict__ line_start,
                              unsigned long long n_lines, long long* __restrict__ end_pos, uint8_t* __restrict__ ref_idx,
                              uint8_t* __restrict__ err) {
    const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines) return;
    unsigned long long q = line_start[k] + 8;
    const uint8_t* f[8];
    int fl[8];
    for (int c = 0; c < 8; c++) {                            // read_to(..., '\t') x 8 (main.cpp:1371-1405)
        unsigned long long e = q;
        while (e < n && in[e] != '\t') e++;
        if (e >= n || e - q > 0x3fffffffull) { end_pos[k] = 0; ref_idx[k] = 0; err[k] = 2; return; }
        f[c] = in + q; fl[c] = (int)(e - q);
        q = e + 1;
    }
    long long pos = 0, endp = 0;
    bool ok = parse_ul(f[1], fl[1], &pos);
    if (ok) {
        bool structural = false;                             // alt_is_structural (main.cpp:759-761)
        for (int i = 0; i < fl[4]; i++) structural |= f[4][i] == '<';
        if (structural) {
            const uint8_t* v = nullptr;
            int vl = 0;
            int r = kvp_lookup(f[7], fl[7], "END", 3, &v, &vl);
            if (r < 0) ok = false;
            else if (r == 1) {
                long long m;
                ok = max_of_list(v, vl, false, &m);
                endp = m < 0 ? -m : m;
            } else {
                r = kvp_lookup(f[7], fl[7], "SVLEN", 5, &v, &vl);
                if (r < 0) ok = false;
                else if (r == 1) {
                    long long m;
                    ok = max_of_list(v, vl, true, &m);
                    endp = pos + m - 1;
                } else {
                    endp = pos;
                }
            }
        } else {
            int max_alt = 0, i = 0;
            while (i < fl[4]) {                              // longest ',' separated ALT allele
                int j = i;
                while (j < fl[4] && f[4][j] != ',') j++;
                if (j - i > max_alt) max_alt = j - i;
                i = j + 1;
            }
            endp = pos + (long long)(fl[3] >= max_alt ? fl[3] : max_alt) - 1;
        }
    }
    end_pos[k] = endp;
    ref_idx[k] = ref_name_index(f[0], fl[0]);
    err[k] = ok ? 0 : 1;
}

}  // namespace idx

int index_line_ends(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, const unsigned long long* d_line_start,
                    unsigned long long n_lines, long long* d_end, uint8_t* d_ref, uint8_t* d_err, cudaStream_t stream) {
    if (n_lines == 0) return VCFC_OK;
    idx::k_index_lines<<<(unsigned)((n_lines + 127) / 128), 128, 0, stream>>>(d_in, (unsigned long long)in_len, d_line_start, n_lines,
                                                                              d_end, d_ref, d_err);
    ctx->launches++;
    VCFC_CUDA(ctx, cudaGetLastError());
    return VCFC_OK;
}

}  // namespace vcfc
