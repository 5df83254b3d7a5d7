// vcfc_files.cu -- host-side file drivers with the reference's verb semantics.  Only host code:
// '#' lines pass through on the host (compress.cpp:222-238), data-line regions go through the
// block codecs (GPU).  Restates, does not copy:
//   compress()                          src/compress.cpp:205-257
//   decompress2_fd()                    src/compress.cpp:1214-1257
//   decompress2_metadata_headers_fd()   src/compress.cpp:1108-1211
//   query_compressed_file()             src/main.cpp:3777-3929
//   parse_coordinate_string()           src/main.cpp:3993-4026
//   create_binned_index4()              src/main.cpp:1284-1637 (per-line fields on the GPU: vcfc_index.cu)
//   query_binned_index_binarysearch()   src/main.cpp:2974-3350
#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "vcfc_index.cuh"
#include "vcfc_internal.h"

namespace {

struct HostFile {                 // the file mapped read-only: a range query touches only the pages it walks over
    uint8_t* p = nullptr;
    size_t   n = 0;
    ~HostFile() { if (p && n) munmap(p, n); }
    int load(const char* path) {
        int fd = open(path, O_RDONLY);
        if (fd < 0) return VCFC_E_IO;
        struct stat st;
        if (fstat(fd, &st) != 0) { close(fd); return VCFC_E_IO; }
        n = (size_t)st.st_size;
        if (n) {
            void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { close(fd); p = nullptr; n = 0; return VCFC_E_IO; }
            p = (uint8_t*)m;
        }
        close(fd);
        return VCFC_OK;
    }
};

int write_all(int fd, const uint8_t* p, size_t n) {
    while (n) {
        ssize_t w = write(fd, p, n);
        if (w < 0 && errno == EINTR) continue;
        if (w <= 0) return VCFC_E_IO;
        p += w;
        n -= (size_t)w;
    }
    return VCFC_OK;
}

// str_to_uint64 (utils.cpp:152-165): strtoul must consume the whole string.
bool parse_u64(const std::string& s, uint64_t* v) {
    char* end = nullptr;
    unsigned long x = strtoul(s.c_str(), &end, 10);
    if (end != s.c_str() + s.size()) return false;
    *v = x;
    return true;
}

struct Query {
    std::string ref;
    uint64_t start = 0, end = 0;
    bool has_range = false;
    // VcfCoordinateQuery::matches, main.cpp:75-86
    bool matches(const std::string& r, uint64_t pos) const {
        if (!ref.empty() && ref != r) return false;
        if (has_range && pos < start) return false;
        if (has_range && pos > end) return false;
        return true;
    }
};

// parse_coordinate_string, main.cpp:3993-4026
int parse_query(const char* s0, Query* q) {
    std::string s(s0 ? s0 : "");
    size_t colon = s.find(':');
    if (colon == std::string::npos) { q->ref = s; return VCFC_OK; }
    q->ref = s.substr(0, colon);
    size_t dash = s.find('-', colon + 1);
    if (dash == std::string::npos) return VCFC_E_QUERY;
    if (!parse_u64(s.substr(colon + 1, dash - (colon + 1)), &q->start)) return VCFC_E_QUERY;
    if (!parse_u64(s.substr(dash + 1), &q->end)) return VCFC_E_QUERY;
    q->has_range = true;
    return VCFC_OK;
}

}  // namespace

extern "C" {

int vcfc_parse_headers(const uint8_t* in, size_t in_len, size_t* header_len, uint64_t* sample_count) {
    if (!in && in_len) return VCFC_E_ARG;
    size_t p = 0;
    bool got_meta = false, got_header = false;
    uint64_t sc = 0;
    for (;;) {
        if (p >= in_len) {
            // compress.cpp:1136-1153: at EOF the stale first byte is still '#': the reference throws
            // "missing headers" before both kinds were seen and "row after header" afterwards.
            return VCFC_E_HEADER;
        }
        if (in[p] != '#') {
            if (!got_meta || !got_header) return VCFC_E_HEADER;
            break;
        }
        if (got_header) return VCFC_E_HEADER;
        if (p + 1 >= in_len) return VCFC_E_HEADER;
        if (in[p + 1] == '#') got_meta = true;
        else { if (!got_meta) return VCFC_E_HEADER; got_header = true; }
        size_t q = p + 2, tabs = 0;
        for (;;) {
            if (q >= in_len) return VCFC_E_HEADER;
            uint8_t c = in[q++];
            if (c == '\n') break;
            if (got_header && c == '\t' && ++tabs > 8) sc++;
        }
        p = q;
    }
    if (header_len) *header_len = p;
    if (sample_count) *sample_count = sc;
    return VCFC_OK;
}

// vcfc_compress_file / vcfc_decompress_file: the pinned, threaded file pipeline in vcfc_pipeline.cu.

int vcfc_query_file(vcfc_ctx* ctx, const char* in_path, const char* region, int out_fd) {
    if (!ctx || !in_path) return VCFC_E_ARG;
    Query q;
    int rc = parse_query(region, &q);
    if (rc) return rc;
    HostFile f;
    if ((rc = f.load(in_path))) return rc;
    size_t hlen = 0;
    uint64_t sc = 0;
    if ((rc = vcfc_parse_headers(f.p, f.n, &hlen, &sc))) return rc;     // header lines are not printed (main.cpp:3789)
    const uint8_t* in = f.p;
    size_t pos = hlen;
    std::vector<uint8_t> hits, out;
    auto flush = [&]() -> int {
        if (hits.empty()) return VCFC_OK;
        size_t olen = 0, nl = 0, cap = std::max<size_t>(hits.size() * 12, (size_t)1 << 20);
        uint64_t el = 0;
        int r = VCFC_OK;
        for (int attempt = 0; attempt < 8; attempt++) {
            out.resize(cap);
            r = vcfc_decode_block(ctx, hits.data(), hits.size(), sc, out.data(), cap, &olen, &nl, &el);
            if (r != VCFC_E_CAP) break;
            cap *= 4;
        }
        int w = write_all(out_fd, out.data(), olen);
        hits.clear();
        return r != VCFC_OK ? r : w;
    };
    while (pos < f.n) {                                                 // main.cpp:3799-3924
        if (f.n - pos < 8) { rc = flush(); return rc ? rc : VCFC_E_TRUNC; }   // "Only read %d bytes, expected 4": hits so far were printed
        if ((in[pos] >> 6) != 3) { flush(); return VCFC_E_FORMAT; }
        size_t ll = ((size_t)(in[pos] & 0x3F) << 24) | ((size_t)in[pos + 1] << 16) | ((size_t)in[pos + 2] << 8) | in[pos + 3];
        size_t p = pos + 8;
        std::string ref, ps;
        while (true) { if (p >= f.n) { flush(); return VCFC_E_TRUNC; } uint8_t c = in[p++]; if (c == '\t') break; ref.push_back((char)c); }
        while (true) { if (p >= f.n) { flush(); return VCFC_E_TRUNC; } uint8_t c = in[p++]; if (c == '\t') break; ps.push_back((char)c); }
        uint64_t v = 0;
        if (!parse_u64(ps, &v)) { flush(); return VCFC_E_FORMAT; }
        if (ll + 4 > f.n - pos) {
            rc = flush();
            return rc ? rc : VCFC_E_TRUNC;
        }
        if (q.matches(ref, v)) {
            hits.insert(hits.end(), in + pos, in + pos + 4 + ll);
            if (hits.size() > ((size_t)64 << 20) && (rc = flush())) return rc;
        }
        pos += 4 + ll;
    }
    return flush();
}

}  // extern "C"

// create-binned-index (main.cpp:4097-4115 -> create_binned_index4, main.cpp:1284-1637).  The host walks the line-length
// headers (one header per line, as every reference consumer does), the GPU reads columns 1-8 of every line and
// computes its END position and chromosome index (vcfc_index.cu), the host applies the bin rule (main.cpp:1430-1470:
// a line whose number is a multiple of entries_per_bin opens an entry if its END exceeds the last entry's position,
// any other line can only grow that position) and writes 13 bytes per entry (main.cpp:600-626).
int vcfc_create_binned_index_file(vcfc_ctx* ctx, const char* vcfc_path, const char* index_path, uint64_t entries_per_bin,
                                  uint64_t* n_entries) {
    if (!ctx || !vcfc_path || !index_path || entries_per_bin == 0) return VCFC_E_ARG;
    if (n_entries) *n_entries = 0;
    HostFile f;
    int rc = f.load(vcfc_path);
    if (rc) return rc;
    size_t hlen = 0;
    uint64_t sc = 0;
    if ((rc = vcfc_parse_headers(f.p, f.n, &hlen, &sc))) return rc;
    // line starts (absolute file offsets = the index's byte offsets)
    std::vector<unsigned long long> starts;
    size_t pos = hlen;
    while (f.n - pos >= 8) {                                            // fewer than 8 bytes left: EOF (compress.cpp:270-330)
        if ((f.p[pos] >> 6) != 3 || (f.p[pos + 4] >> 6) != 3) return VCFC_E_FORMAT;
        size_t ll = ((size_t)(f.p[pos] & 0x3F) << 24) | ((size_t)f.p[pos + 1] << 16) | ((size_t)f.p[pos + 2] << 8) | f.p[pos + 3];
        if (ll + 4 > f.n - pos) return VCFC_E_TRUNC;
        starts.push_back((unsigned long long)pos);
        pos += 4 + ll;
    }
    const size_t n_lines = starts.size();
    std::vector<long long> ends(n_lines);
    std::vector<uint8_t> refs(n_lines), errs(n_lines);
    // per-line fields on the device, in pieces of whole lines
    const size_t piece = (size_t)512 << 20;
    cudaStream_t st = ctx->stream;
    for (size_t l0 = 0; l0 < n_lines;) {
        size_t l1 = l0 + 1;
        while (l1 < n_lines && starts[l1] - starts[l0] < piece) l1++;
        const size_t b0 = (size_t)starts[l0], b1 = l1 < n_lines ? (size_t)starts[l1] : f.n, nl = l1 - l0;
        vcfc::DevBuf &d_data = ctx->d_in[0], &d_ls = ctx->ws[3], &d_end = ctx->ws[4], &d_ref = ctx->ws[5], &d_err = ctx->ws[6];
        if ((rc = vcfc::dev_reserve(ctx, &d_data, b1 - b0 + 16))) return rc;
        if ((rc = vcfc::dev_reserve(ctx, &d_ls, nl * 8 + 8))) return rc;
        if ((rc = vcfc::dev_reserve(ctx, &d_end, nl * 8 + 8))) return rc;
        if ((rc = vcfc::dev_reserve(ctx, &d_ref, nl + 8))) return rc;
        if ((rc = vcfc::dev_reserve(ctx, &d_err, nl + 8))) return rc;
        std::vector<unsigned long long> rel(nl);
        for (size_t k = 0; k < nl; k++) rel[k] = starts[l0 + k] - b0;
        VCFC_CUDA(ctx, cudaMemcpyAsync(d_data.p, f.p + b0, b1 - b0, cudaMemcpyHostToDevice, st));
        VCFC_CUDA(ctx, cudaMemcpyAsync(d_ls.p, rel.data(), nl * 8, cudaMemcpyHostToDevice, st));
        if ((rc = vcfc::index_line_ends(ctx, (const uint8_t*)d_data.p, b1 - b0, (const unsigned long long*)d_ls.p, nl, (long long*)d_end.p,
                                        (uint8_t*)d_ref.p, (uint8_t*)d_err.p, st)))
            return rc;
        VCFC_CUDA(ctx, cudaMemcpyAsync(ends.data() + l0, d_end.p, nl * 8, cudaMemcpyDeviceToHost, st));
        VCFC_CUDA(ctx, cudaMemcpyAsync(refs.data() + l0, d_ref.p, nl, cudaMemcpyDeviceToHost, st));
        VCFC_CUDA(ctx, cudaMemcpyAsync(errs.data() + l0, d_err.p, nl, cudaMemcpyDeviceToHost, st));
        VCFC_CUDA(ctx, cudaStreamSynchronize(st));
        l0 = l1;
    }
    // the bin rule, and the entries
    std::vector<uint8_t> out;
    out.reserve(13 * (n_lines / entries_per_bin + 2));
    size_t n_ent = 0;
    uint32_t last_end = 0;
    auto put = [&](uint8_t ref, uint32_t position, uint64_t offset) {
        out.push_back(ref);
        const uint8_t* a = reinterpret_cast<const uint8_t*>(&position);
        out.insert(out.end(), a, a + 4);
        const uint8_t* b = reinterpret_cast<const uint8_t*>(&offset);
        out.insert(out.end(), b, b + 8);
    };
    for (size_t k = 0; k < n_lines; k++) {
        if (errs[k]) return errs[k] == 2 ? VCFC_E_TRUNC : VCFC_E_FORMAT;    // the reference throws at this line
        const unsigned long e = (unsigned long)ends[k];
        if (n_ent == 0) {
            last_end = (uint32_t)e;
            put(refs[k], last_end, starts[k]);
            n_ent = 1;
        } else if (e > (unsigned long)last_end) {
            last_end = (uint32_t)e;
            if (k % entries_per_bin == 0) { put(refs[k], last_end, starts[k]); n_ent++; }
            else memcpy(out.data() + 13 * (n_ent - 1) + 1, &last_end, 4);
        }
    }
    int fd = open(index_path, O_CREAT | O_TRUNC | O_WRONLY, 0644);
    if (fd < 0) return VCFC_E_IO;
    rc = write_all(fd, out.data(), out.size());
    close(fd);
    if (n_entries) *n_entries = n_ent;
    return rc;
}

// query-binned-index (main.cpp:4117-4143 -> query_binned_index_binarysearch, main.cpp:2974-3350): binary search of the
// .vcfci entries for the query's chromosome index and start position, then a walk over the compressed lines from the
// chosen entry's byte offset -- a line is printed when [POS, END] overlaps the query, the walk stops at the first line
// behind it -- and ONE GPU decode of the lines that matched.  The search keeps the reference's exact steps (the entry
// it ends on is the last one READ, not necessarily the one at the final bounds).  Differences: an index with a single
// entry starts at that entry (the reference reads an uninitialised struct there), and the text is written as is
// (the reference passes it to printf as the format string, so a '%' in a line would be mangled).
int vcfc_query_binned_index_file(vcfc_ctx* ctx, const char* vcfc_path, const char* region, int out_fd) {
    if (!ctx || !vcfc_path) return VCFC_E_ARG;
    Query q;
    int rc = parse_query(region, &q);
    if (rc) return rc;
    HostFile f, x;
    if ((rc = f.load(vcfc_path))) return rc;
    if ((rc = x.load((std::string(vcfc_path) + ".vcfci").c_str())) != VCFC_OK) return rc;
    size_t hlen = 0;
    uint64_t sc = 0;
    if ((rc = vcfc_parse_headers(f.p, f.n, &hlen, &sc))) return rc;
    const uint8_t qidx = vcfc::idx::ref_name_index((const uint8_t*)q.ref.data(), (int)q.ref.size());
    if (x.n % 13 != 0) return VCFC_E_FORMAT;                             // "Index size was not a multiple of entry size"
    const long count = (long)(x.n / 13);
    if (count == 0) return VCFC_OK;
    struct Entry { uint8_t ref; uint32_t position; uint64_t offset; };
    auto read_entry = [&](long i) {
        Entry e;
        e.ref = x.p[13 * i];
        memcpy(&e.position, x.p + 13 * i + 1, 4);
        memcpy(&e.offset, x.p + 13 * i + 5, 8);
        return e;
    };
    auto greater = [&](const Entry& e) { return e.ref > qidx || (e.ref == qidx && (uint64_t)e.position > q.start); };
    auto less = [&](const Entry& e) { return e.ref < qidx || (e.ref == qidx && (uint64_t)e.position < q.start); };
    long s_lo = 0, s_hi = count - 1, mid = (s_lo + s_hi) / 2;
    Entry entry = read_entry(0);
    while (s_lo < s_hi) {                                                // main.cpp:3040-3128
        mid = (s_lo + s_hi) / 2;
        entry = read_entry(mid);
        if (entry.ref == qidx && (uint64_t)entry.position == q.start) break;
        if (greater(entry)) {
            if (mid == 0) break;
            s_hi = mid - 1;
        } else if (less(entry)) {
            s_lo = mid + 1;
        }
    }
    if (mid > 0 && greater(entry)) {                                     // main.cpp:3130-3152: one entry back
        mid--;
        entry = read_entry(mid);
    }
    const uint8_t* in = f.p;
    std::vector<uint8_t> hits, out;
    size_t pos = (size_t)entry.offset;
    while (pos <= f.n && f.n - pos >= 8) {                               // main.cpp:3170-3330
        if ((in[pos] >> 6) != 3 || (in[pos + 4] >> 6) != 3) return VCFC_E_FORMAT;
        const size_t ll = ((size_t)(in[pos] & 0x3F) << 24) | ((size_t)in[pos + 1] << 16) | ((size_t)in[pos + 2] << 8) | in[pos + 3];
        const uint8_t* fld[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        int fl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        size_t p = pos + 8;
        int n_fields = 5;
        for (int c = 0; c < n_fields; c++) {                             // read_to(..., '\t'): CHROM POS ID REF ALT [QUAL FILTER INFO]
            size_t e = p;
            while (e < f.n && in[e] != '\t') e++;
            if (e >= f.n) return VCFC_E_TRUNC;
            fld[c] = in + p; fl[c] = (int)(e - p);
            p = e + 1;
            if (c == 4) for (int i = 0; i < fl[4]; i++) if (fld[4][i] == '<') { n_fields = 8; break; }   // alt_is_structural
        }
        long long lpos = 0, lend = 0;
        if (!vcfc::idx::parse_ul(fld[1], fl[1], &lpos)) return VCFC_E_FORMAT;
        if (!vcfc::idx::line_end_position(lpos, fl[3], fld[4], fl[4], fld[7], fl[7], &lend)) return VCFC_E_FORMAT;
        const uint8_t lidx = vcfc::idx::ref_name_index(fld[0], fl[0]);
        // VcfCoordinateQuery::compare_to_range (main.cpp:108-140)
        int cmp;
        if (lidx < qidx || (lidx == qidx && (uint64_t)lend < q.start)) cmp = 1;            // the line lies before the query
        else if (lidx > qidx || (lidx == qidx && (uint64_t)lpos > q.end)) cmp = -1;        // ... behind it
        else cmp = 0;
        if (cmp < 0) break;
        if (cmp == 0) {
            if (ll + 4 > f.n - pos) return VCFC_E_TRUNC;
            hits.insert(hits.end(), in + pos, in + pos + 4 + ll);
        }
        pos += 4 + ll;
    }
    if (hits.empty()) return VCFC_OK;
    size_t olen = 0, nl = 0, cap = std::max<size_t>(hits.size() * 12, (size_t)1 << 20);
    uint64_t el = 0;
    for (int attempt = 0; attempt < 8; attempt++) {
        out.resize(cap);
        rc = vcfc_decode_block(ctx, hits.data(), hits.size(), sc, out.data(), cap, &olen, &nl, &el);
        if (rc != VCFC_E_CAP) break;
        cap *= 4;
    }
    const int w = write_all(out_fd, out.data(), olen);
    return rc != VCFC_OK ? rc : w;
}
