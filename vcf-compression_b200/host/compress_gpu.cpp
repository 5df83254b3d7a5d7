// compress_gpu.cpp -- link-compatible replacement of the reference's src/compress.cpp on top of libvcfc_gpu.so.
//
// Every function the reference declares in src/compress.hpp:17-56 is defined here with its exact C++ signature, so
// the reference's own main.cpp / utils.cpp / sparse.cpp link against this file INSTEAD of compress.cpp and every verb
// of the reference binary (compress, decompress, query, create-binned-index, query-binned-index, sparsify, ...) runs
// its coding work on the GPU.  Only the reference's headers are included (found with -I<reference>/src at build
// time); no reference source is copied.  Recipe: oracle/Makefile, target ref_gpu -> oracle/_ref_gpu/main_gpu.
//
//   compress / decompress2_fd                    -> the file pipeline (vcfc_compress_file / vcfc_decompress_file)
//   compress_data_line / decompress2_data_line   -> a block of ONE line through vcfc_encode_block / vcfc_decode_block
//                                                   (correct, and slow: callers that want speed use the block API)
//   decompress2_metadata_headers(_fd)            -> vcfc_parse_headers
//   read_compressed_line_length_headers(_fd)     -> 8 bytes of host I/O, as in the reference
// Where the reference throws (and its main() lets the exception abort the process), this throws the same type.
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "compress.hpp"      // the reference's header: declarations + VcfCompressionSchema, compressed_line_length_headers
#include "vcfc_gpu.h"

namespace {

vcfc_ctx* gpu() {            // one context per process (the reference is single-threaded, compress.cpp:411)
    static vcfc_ctx* ctx = nullptr;
    if (!ctx) {
        const char* d = getenv("VCFC_DEVICE");
        if (vcfc_gpu_init(d ? atoi(d) : 0, &ctx) != VCFC_OK)
            throw std::runtime_error("vcfc: no usable sm_100 device (libvcfc_gpu has no CPU path)");
    }
    return ctx;
}

[[noreturn]] void fail(int rc) {
    if (rc == VCFC_E_CUDA) throw std::runtime_error(std::string("vcfc: CUDA error: ") + vcfc_last_cuda_error(gpu()));
    throw VcfValidationError(vcfc_strerror(rc));
}

struct Src {                 // a FILE* or an fd, read sequentially from its current position
    FILE* f;
    int   fd;
    size_t get(void* p, size_t n) const {
        if (f) return fread(p, 1, n, f);
        size_t got = 0;
        while (got < n) {
            ssize_t r = read(fd, (char*)p + got, n - got);
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) break;
            got += (size_t)r;
        }
        return got;
    }
    void back(long n) const {
        if (f) fseek(f, -n, SEEK_CUR);
        else lseek(fd, -n, SEEK_CUR);
    }
};

// compress.cpp:270-330 / 342-401: 8 header bytes -> the two lengths; returns 8, 0 at EOF, or the short count
int read_headers(const Src& s, compressed_line_length_headers* h) {
    uint8_t b[8];
    const size_t got = s.get(b, 8);
    if (got == 0) return 0;
    if (got < 8) return (int)got;
    LineLengthHeader a, r;
    a.deserialize(b);            // throws unless the top two bits are 11 (utils.hpp:201-206)
    r.deserialize(b + 4);
    h->line_length = a.length;
    h->required_columns_length = r.length;
    return 8;
}

// compress.cpp:741-986: one compressed line at the stream position -> its text appended to linebuf
int decode_line(const Src& s, const VcfCompressionSchema& schema, std::string& linebuf, size_t* compressed_line_length) {
    std::vector<uint8_t> in(8);
    const size_t got = s.get(in.data(), 8);
    if (got < 8) return 0;                                        // EOF or a short tail (compress.cpp:771-777)
    if ((in[0] >> 6) != 3 || (in[4] >> 6) != 3) fail(VCFC_E_FORMAT);
    const size_t ll = ((size_t)(in[0] & 0x3F) << 24) | ((size_t)in[1] << 16) | ((size_t)in[2] << 8) | in[3];
    if (ll < 4) fail(VCFC_E_FORMAT);
    in.resize(4 + ll);
    if (s.get(in.data() + 8, ll - 4) != ll - 4) fail(VCFC_E_TRUNC);
    // a token byte expands to at most 127 samples of 4 bytes; the required section passes through
    std::vector<uint8_t> out(in.size() * 512 + 64);
    size_t olen = 0, nl = 0;
    uint64_t el = 0;
    const int rc = vcfc_decode_block(gpu(), in.data(), in.size(), schema.sample_count, out.data(), out.size(), &olen, &nl, &el);
    if (rc != VCFC_OK) fail(rc);
    linebuf.append((const char*)out.data(), olen);
    if (compressed_line_length) *compressed_line_length = in.size();
    return 1;
}

// compress.cpp:995-1098 / 1108-1211: '##' lines, then the '#CHROM' line; leaves the stream at the first data line
int read_meta(const Src& s, std::vector<std::string>& lines, VcfCompressionSchema& schema) {
    std::vector<uint8_t> buf;
    size_t hlen = 0;
    uint64_t sc = 0;
    int rc = VCFC_E_HEADER;
    bool eof = false;
    for (size_t want = 1 << 16; !eof; want *= 4) {
        const size_t old = buf.size();
        buf.resize(old + want);
        const size_t got = s.get(buf.data() + old, want);
        buf.resize(old + got);
        eof = got < want;
        rc = vcfc_parse_headers(buf.data(), buf.size(), &hlen, &sc);
        if (rc == VCFC_OK) break;
    }
    if (rc != VCFC_OK) throw VcfValidationError("File was missing headers or metadata");
    s.back((long)(buf.size() - hlen));
    for (size_t p = 0; p < hlen;) {
        const uint8_t* e = (const uint8_t*)memchr(buf.data() + p, '\n', hlen - p);
        const size_t q = e ? (size_t)(e - buf.data()) + 1 : hlen;
        lines.emplace_back((const char*)buf.data() + p, q - p);   // the lines keep their newline (compress.cpp:1186)
        p = q;
    }
    schema.sample_count += sc;
    return 0;
}

}  // namespace

int compress(const std::string& input_filename, const std::string& output_filename) {
    const int rc = vcfc_compress_file(gpu(), input_filename.c_str(), output_filename.c_str());
    if (rc != VCFC_OK) fail(rc);                                  // compress.cpp:9-11, 88-106, 230-233 throw
    return 0;
}

int compress_data_line(const std::string& line, const VcfCompressionSchema&, std::vector<byte_t>& byte_vec, bool add_newline) {
    std::string in = line;
    in.push_back('\n');
    std::vector<uint8_t> out(vcfc_encode_bound(in.size()));
    size_t olen = 0, nl = 0;
    uint64_t el = 0;
    const int rc = vcfc_encode_block(gpu(), (const uint8_t*)in.data(), in.size(), out.data(), out.size(), &olen, nullptr, 0, &nl, &el);
    if (rc != VCFC_OK) fail(rc);
    if (!add_newline && olen) {                                   // compress.cpp:188-199: the length header counts the newline only if it is there
        olen--;
        const size_t ll = olen - 4;
        out[0] = (uint8_t)(0xC0 | (ll >> 24)); out[1] = (uint8_t)(ll >> 16); out[2] = (uint8_t)(ll >> 8); out[3] = (uint8_t)ll;
    }
    byte_vec.insert(byte_vec.end(), out.begin(), out.begin() + olen);
    return 0;
}

int decompress2_fd(const std::string& input_filename, const std::string& output_filename) {
    const int rc = vcfc_decompress_file(gpu(), input_filename.c_str(), output_filename.c_str());
    if (rc != VCFC_OK) fail(rc);                                  // compress.cpp:793, 825, 877, 959, 965, 1136-1170 throw
    return 0;
}

int decompress2_metadata_headers_fd(int input_fd, std::vector<std::string>& output_vector, VcfCompressionSchema& output_schema) {
    return read_meta(Src{nullptr, input_fd}, output_vector, output_schema);
}
int decompress2_metadata_headers(FILE* input_file, std::vector<std::string>& output_vector, VcfCompressionSchema& output_schema) {
    return read_meta(Src{input_file, -1}, output_vector, output_schema);
}

int decompress2_data_line(FILE* input_file, const VcfCompressionSchema& schema, std::string& linebuf, size_t* compressed_line_length) {
    return decode_line(Src{input_file, -1}, schema, linebuf, compressed_line_length);
}
int decompress2_data_line_FILEwrapper(int input_fd, const VcfCompressionSchema& schema, std::string& linebuf, size_t* compressed_line_length) {
    return decode_line(Src{nullptr, input_fd}, schema, linebuf, compressed_line_length);
}

int read_compressed_line_length_headers(FILE* input_file, struct compressed_line_length_headers* length_headers) {
    return read_headers(Src{input_file, -1}, length_headers);
}
int read_compressed_line_length_headers_fd(int input_fd, struct compressed_line_length_headers* length_headers) {
    return read_headers(Src{nullptr, input_fd}, length_headers);
}
