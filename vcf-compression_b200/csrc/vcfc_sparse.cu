// vcfc_sparse.cu -- the reference's sparse-file verbs on the block decoder (host code).  Restates, does not copy:
//   sparsify_file()           src/sparse.cpp:290-580   every compressed line is copied to offset (300,000,000 + POS) * 4 * 4096
//                                                      behind the header of a holey file, with 16 bytes of previous / next
//                                                      distance links in front of it
//   query_sparse_file_fd()    src/main.cpp:235-582     seek to the computed offset (lseek SEEK_DATA over the holes), follow the
//                                                      next links, decode the lines
// The walk is the reference's, system call by system call where the result depends on it (SEEK_DATA granularity); the lines a
// range query walks over are decoded in ONE GPU block call instead of one by one.
#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "vcfc_internal.h"

namespace {

constexpr uint64_t kMaxPosition = 300000000ull;        // SparsificationConfiguration, sparse.hpp:29-32
constexpr uint64_t kSlot = 4ull * 4096ull;             // multiplication_factor * block_size

uint64_t sparse_offset(uint64_t pos) { return (kMaxPosition + pos) * kSlot; }   // compute_sparse_offset, sparse.cpp:18-51 (one reference per file)

void put_be64(uint8_t* d, uint64_t v) { for (int i = 0; i < 8; i++) d[i] = (uint8_t)(v >> (8 * (7 - i))); }   // uint64_to_uint8_array
uint64_t get_be64(const uint8_t* s) { uint64_t v = 0; for (int i = 0; i < 8; i++) v = (v << 8) | s[i]; return v; }

int pwrite_all(int fd, const uint8_t* p, size_t n, off_t off) {
    while (n) {
        ssize_t w = pwrite(fd, p, n, off);
        if (w < 0 && errno == EINTR) continue;
        if (w <= 0) return VCFC_E_IO;
        p += w; n -= (size_t)w; off += w;
    }
    return VCFC_OK;
}
int write_all_fd(int fd, const uint8_t* p, size_t n) {
    while (n) {
        ssize_t w = write(fd, p, n);
        if (w < 0 && errno == EINTR) continue;
        if (w <= 0) return VCFC_E_IO;
        p += w; n -= (size_t)w;
    }
    return VCFC_OK;
}
// bytes [off, off + n) of fd; short reads at EOF return what is there
ssize_t pread_some(int fd, uint8_t* p, size_t n, off_t off) {
    size_t got = 0;
    while (got < n) {
        ssize_t r = pread(fd, p + got, n - got, off + (off_t)got);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) break;
        got += (size_t)r;
    }
    return (ssize_t)got;
}

// the header region of a .vcfc / sparse file read through the descriptor (the sparse file is ~4.9 TB logical: no mapping)
int read_header_region(int fd, std::vector<uint8_t>* head, size_t* hlen, uint64_t* sample_count) {
    for (size_t want = (size_t)1 << 20; want <= ((size_t)1 << 30); want <<= 2) {
        head->resize(want);
        const ssize_t got = pread_some(fd, head->data(), want, 0);
        if (got <= 0) return VCFC_E_HEADER;
        const int rc = vcfc_parse_headers(head->data(), (size_t)got, hlen, sample_count);
        if (rc == VCFC_OK) return rc;
        if ((size_t)got < want) return rc;               // the whole file was looked at
    }
    return VCFC_E_HEADER;
}

// one compressed line at `off` (behind the 16 link bytes): [len4][req4][bytes]; appended to `blk`
int read_line_at(int fd, off_t off, std::vector<uint8_t>* blk, size_t* total) {
    uint8_t h[8];
    if (pread_some(fd, h, 8, off) < 8) return VCFC_E_TRUNC;
    if ((h[0] >> 6) != 3 || (h[4] >> 6) != 3) return VCFC_E_FORMAT;
    const size_t ll = ((size_t)(h[0] & 0x3F) << 24) | ((size_t)h[1] << 16) | ((size_t)h[2] << 8) | h[3];
    if (ll < 4) return VCFC_E_FORMAT;
    const size_t at = blk->size();
    blk->resize(at + 4 + ll);
    memcpy(blk->data() + at, h, 8);
    if (pread_some(fd, blk->data() + at + 8, ll - 4, off + 8) < (ssize_t)(ll - 4)) return VCFC_E_TRUNC;
    *total = 4 + ll;
    return VCFC_OK;
}

int decode_and_print(vcfc_ctx* ctx, const std::vector<uint8_t>& blk, uint64_t sc, int out_fd) {
    if (blk.empty()) return VCFC_OK;
    std::vector<uint8_t> out;
    size_t olen = 0, nl = 0, cap = std::max<size_t>(blk.size() * 24, (size_t)1 << 20);
    uint64_t el = 0;
    int r = VCFC_OK;
    for (int attempt = 0; attempt < 8; attempt++) {
        out.resize(cap);
        r = vcfc_decode_block(ctx, blk.data(), blk.size(), sc, out.data(), cap, &olen, &nl, &el);
        if (r != VCFC_E_CAP) break;
        cap *= 4;
    }
    const int w = write_all_fd(out_fd, out.data(), olen);
    return r != VCFC_OK ? r : w;
}

}  // namespace

extern "C" {

int vcfc_sparsify_file(const char* vcfc_path, const char* sparse_path) {
    if (!vcfc_path || !sparse_path) return VCFC_E_ARG;
    int in_fd = open(vcfc_path, O_RDONLY);
    if (in_fd < 0) return VCFC_E_IO;
    struct stat st;
    if (fstat(in_fd, &st) != 0) { close(in_fd); return VCFC_E_IO; }
    std::vector<uint8_t> f((size_t)st.st_size);
    if (pread_some(in_fd, f.data(), f.size(), 0) < (ssize_t)f.size()) { close(in_fd); return VCFC_E_IO; }
    close(in_fd);
    size_t hlen = 0;
    uint64_t sc = 0;
    int rc = vcfc_parse_headers(f.data(), f.size(), &hlen, &sc);
    if (rc) return rc;
    int fd = open(sparse_path, O_CREAT | O_TRUNC | O_RDWR, S_IRUSR | S_IWUSR);        // utils.hpp:27-28
    if (fd < 0) return VCFC_E_IO;
    uint8_t zeros[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    rc = pwrite_all(fd, f.data(), hlen, 0);                                        // header lines as they are (sparse.cpp:316-318)
    if (!rc) rc = pwrite_all(fd, zeros, 8, (off_t)hlen);                             // the first line's offset goes here (:332-335)
    const uint64_t data_start = (uint64_t)hlen + 8;
    uint64_t previous = data_start;
    bool first = true;
    std::vector<uint8_t> rec;
    size_t pos = hlen;
    while (!rc && f.size() - pos >= 8) {                                            // fewer than 8 bytes left: EOF (compress.cpp:270-330)
        if ((f[pos] >> 6) != 3 || (f[pos + 4] >> 6) != 3) { rc = VCFC_E_FORMAT; break; }
        const size_t ll = ((size_t)(f[pos] & 0x3F) << 24) | ((size_t)f[pos + 1] << 16) | ((size_t)f[pos + 2] << 8) | f[pos + 3];
        if (ll < 4 || ll + 4 > f.size() - pos) { rc = VCFC_E_TRUNC; break; }         // "Unexpectedly reached end of compressed file"
        // CHROM and POS: the first two tab-terminated fields of the line's bytes (sparse.cpp:432-471)
        const uint8_t* lb = f.data() + pos + 8;
        const size_t ln = ll - 4;
        size_t i = 0, r0 = 0;
        while (i < ln && lb[i] != '\t') i++;
        const bool got_ref = i < ln;
        if (got_ref && i == 0) { rc = VCFC_E_FORMAT; break; }                        // "Line did not contain a reference name"
        uint64_t vpos = 0;
        if (got_ref) {
            r0 = ++i;
            while (i < ln && lb[i] != '\t') i++;
            if (i < ln) {
                if (i == r0) { rc = VCFC_E_FORMAT; break; }                          // "Line did not contain a position value"
                std::string ps((const char*)lb + r0, i - r0);
                char* end = nullptr;
                vpos = strtoul(ps.c_str(), &end, 10);
                if (end != ps.c_str() + ps.size()) { rc = VCFC_E_FORMAT; break; }    // "Failed to parse full position value"
            }
        }
        const uint64_t voff = sparse_offset(vpos), file_off = voff + data_start;
        rec.assign(16, 0);
        put_be64(rec.data(), file_off - previous);                                  // distance to the previous record (:476-483)
        rec.insert(rec.end(), f.data() + pos, f.data() + pos + 4 + ll);             // both length headers re-serialised = as stored
        if (first) {
            rc = pwrite_all(fd, reinterpret_cast<const uint8_t*>(&voff), 8, (off_t)(data_start - 8));   // NATIVE byte order (:511)
            first = false;
        } else {
            uint8_t nx[8];
            put_be64(nx, file_off - previous);
            rc = pwrite_all(fd, nx, 8, (off_t)(previous + 8));                       // the previous record's next link (:530-553)
        }
        if (!rc) rc = pwrite_all(fd, rec.data(), rec.size(), (off_t)file_off);
        previous = file_off;
        pos += 4 + ll;
    }
    if (close(fd) != 0 && !rc) rc = VCFC_E_IO;
    return rc;
}

int vcfc_sparse_query_file(vcfc_ctx* ctx, const char* sparse_path, const char* region, int out_fd) {
    if (!ctx || !sparse_path || !region) return VCFC_E_ARG;
    // parse_coordinate_string, main.cpp:3993-4026
    std::string s(region), ref;
    uint64_t q_start = 0, q_end = 0;
    bool has_range = false;
    {
        const size_t colon = s.find(':');
        if (colon == std::string::npos) ref = s;
        else {
            ref = s.substr(0, colon);
            const size_t dash = s.find('-', colon + 1);
            if (dash == std::string::npos) return VCFC_E_QUERY;
            const std::string a = s.substr(colon + 1, dash - (colon + 1)), b = s.substr(dash + 1);
            char* e = nullptr;
            q_start = strtoul(a.c_str(), &e, 10);
            if (e != a.c_str() + a.size()) return VCFC_E_QUERY;
            q_end = strtoul(b.c_str(), &e, 10);
            if (e != b.c_str() + b.size()) return VCFC_E_QUERY;
            has_range = true;
        }
    }
    if (!has_range) return VCFC_E_QUERY;                  // "sparse query with no filter is not yet implemented" (main.cpp:570)
    int fd = open(sparse_path, O_RDONLY);
    if (fd < 0) return VCFC_E_IO;
    struct FdGuard { int fd; ~FdGuard() { close(fd); } } guard{fd};
    std::vector<uint8_t> head;
    size_t hlen = 0;
    uint64_t sc = 0;
    int rc = read_header_region(fd, &head, &hlen, &sc);
    if (rc) return rc;
    head.clear(); head.shrink_to_fit();
    const off_t data_start = (off_t)hlen + 8;
    uint64_t first_line_offset = 0;                       // native byte order, as written
    if (pread_some(fd, reinterpret_cast<uint8_t*>(&first_line_offset), 8, (off_t)hlen) < 8) return VCFC_E_TRUNC;
    std::vector<uint8_t> blk;
    uint8_t d[16];
    if (q_start == q_end) {
        // one position (main.cpp:273-318): a record is there when its first link is non-zero, or it is the file's first line
        const off_t at = data_start + (off_t)sparse_offset(q_start);
        if (pread_some(fd, d, 16, at) == 0) return VCFC_E_TRUNC;          // "Reached end of file unexpectedly"
        uint64_t prev_native;
        memcpy(&prev_native, d, 8);
        if (prev_native == 0 && at != (off_t)(first_line_offset + (uint64_t)data_start)) return VCFC_OK;   // a hole: nothing to print
        size_t total = 0;
        if ((rc = read_line_at(fd, at + 16, &blk, &total))) return rc == VCFC_E_TRUNC ? VCFC_E_TRUNC : rc;
        return decode_and_print(ctx, blk, sc, out_fd);
    }
    // a range (main.cpp:319-560): the first record at or behind the start position ...
    const off_t lookup = data_start + (off_t)sparse_offset(q_start);
    off_t cur = lseek(fd, lookup, SEEK_DATA);
    if (cur < lookup) return VCFC_E_TRUNC;                 // no data behind it: the reference throws
    if (cur != lookup) {
        const off_t m = (cur - data_start) % (off_t)kSlot;
        if (m != 0) cur += (off_t)kSlot - m;               // data extents start on the file system's blocks: on to the next slot
    }
    for (;;) {
        if (pread_some(fd, d, 16, cur) < 16) return VCFC_E_TRUNC;
        if (get_be64(d) == 0 && lookup != (off_t)(first_line_offset + (uint64_t)data_start)) cur += (off_t)kSlot;   // an empty slot inside an extent
        else break;
    }
    // ... then along the next links while the reference name matches and POS <= end
    for (;;) {
        if (pread_some(fd, d, 16, cur) < 16) { rc = VCFC_E_TRUNC; break; }
        const uint64_t d_prev = get_be64(d);
        uint64_t d_next = get_be64(d + 8);
        if (d_prev == 0 && d_next == 0) { rc = VCFC_E_FORMAT; break; }      // "No previous or next distance values"
        const bool end_of_reference = d_next == 0;
        const size_t at = blk.size();
        size_t total = 0;
        if ((rc = read_line_at(fd, cur + 16, &blk, &total))) { blk.resize(at); break; }
        // CHROM and POS of the decoded line = the first two fields of the line's required section
        const uint8_t* lb = blk.data() + at + 8;
        const size_t ln = total - 8;
        size_t i = 0;
        while (i < ln && lb[i] != '\t') i++;
        const std::string r(reinterpret_cast<const char*>(lb), i);
        size_t j = i < ln ? i + 1 : i;
        const size_t p0 = j;
        while (j < ln && lb[j] != '\t') j++;
        const std::string ps(reinterpret_cast<const char*>(lb) + p0, j - p0);
        char* e = nullptr;
        const uint64_t pos = strtoul(ps.c_str(), &e, 10);
        if (e != ps.c_str() + ps.size()) { blk.resize(at); rc = VCFC_E_FORMAT; break; }
        if (!(r == ref && pos <= q_end)) { blk.resize(at); break; }
        if (end_of_reference || pos >= q_end) break;
        cur += (off_t)d_next;                               // (the reference subtracts what it has read and seeks the rest)
        if (blk.size() > ((size_t)64 << 20)) { const int w = decode_and_print(ctx, blk, sc, out_fd); blk.clear(); if (w) return w; }
    }
    const int w = decode_and_print(ctx, blk, sc, out_fd);   // lines found before an error are still printed, as in the reference
    return rc ? rc : w;
}

}  // extern "C"
