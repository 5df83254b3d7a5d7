// vcfc_index.cuh -- column parsing shared by the index kernel (device) and the indexed query's host walk.
// Restates str_to_uint64 / str_to_long (utils.cpp:152-175), reference_name_map (utils.hpp:90-103, utils.cpp:16-25),
// parse_kvp (main.cpp:737-757) and compute_end_position (main.cpp:763-852) on raw bytes.
#pragma once
#include <stdint.h>

namespace vcfc {
namespace idx {

// strtoul(s, &end, 10) with end == s + n, as str_to_uint64 / str_to_long use it (utils.cpp:152-175): leading
// white space, an optional sign, digits; saturates at ULONG_MAX; a '-' negates modulo 2^64.  Fields of 64 bytes
// and more, or with a NUL inside, fail (as in the test oracle).
__host__ __device__ inline bool parse_ul(const uint8_t* p, int n, long long* out) {
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

__host__ __device__ inline uint8_t ref_name_index(const uint8_t* p, int n) {
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
__host__ __device__ inline int kvp_lookup(const uint8_t* info, int n, const char* key, int klen, const uint8_t** val, int* val_len) {
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
__host__ __device__ inline bool max_of_list(const uint8_t* v, int vl, bool absolute, long long* out) {
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

// compute_end_position (main.cpp:763-852) on raw column bytes; false where the reference throws
__host__ __device__ inline bool line_end_position(long long pos, int ref_len, const uint8_t* alt, int alt_len, const uint8_t* info,
                                                  int info_len, long long* end_out) {
    bool structural = false;                             // alt_is_structural (main.cpp:759-761)
    for (int i = 0; i < alt_len; i++) structural |= alt[i] == '<';
    if (structural) {
        const uint8_t* v = nullptr;
        int vl = 0;
        int r = kvp_lookup(info, info_len, "END", 3, &v, &vl);
        if (r < 0) return false;
        if (r == 1) {
            long long m;
            if (!max_of_list(v, vl, false, &m)) return false;
            *end_out = m < 0 ? -m : m;
            return true;
        }
        r = kvp_lookup(info, info_len, "SVLEN", 5, &v, &vl);
        if (r < 0) return false;
        if (r == 1) {
            long long m;
            if (!max_of_list(v, vl, true, &m)) return false;
            *end_out = pos + m - 1;
            return true;
        }
        *end_out = pos;
        return true;
    }
    int max_alt = 0, i = 0;
    while (i < alt_len) {                                // longest ',' separated ALT allele
        int j = i;
        while (j < alt_len && alt[j] != ',') j++;
        if (j - i > max_alt) max_alt = j - i;
        i = j + 1;
    }
    *end_out = pos + (long long)(ref_len >= max_alt ? ref_len : max_alt) - 1;
    return true;
}

}  // namespace idx
}  // namespace vcfc
