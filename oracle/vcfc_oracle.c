/*
 * vcfc_oracle.c -- CPU restatement of the vcf-compression genotype-column codec.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (libvcfc_gpu.so, the vcfc CLI) never links, loads or
 * calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement byte-for-byte
 * against (a) the known-answer vectors of SURVEY.md 8(c), (b) tests/golden/ fixtures
 * produced by the unmodified reference binary (oracle/make_golden.py), and, when
 * oracle/_ref/main_release is present, (c) the reference binary run live.
 *
 * Each function cites the reference lines (under /root/reference/) that it restates.
 * Nothing here is copied from the reference: the reference works on std::string /
 * std::vector one line at a time; this is a flat byte-buffer restatement in C.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define VCFC_OK            0
#define VCFC_E_TOOFEW     -1  /* <8 tab-separated terms: reference throws (compress.cpp:9-11)      */
#define VCFC_E_EIGHTCOLS  -2  /* exactly 8 terms: reference aborts in resize (compress.cpp:88-106)  */
#define VCFC_E_CAP        -3  /* output buffer too small                                           */
#define VCFC_E_FORMAT     -4  /* decode: malformed compressed line                                 */
#define VCFC_E_TRUNC      -5  /* decode: input ended inside a line                                 */
#define VCFC_E_IO         -6
#define VCFC_E_HEADER     -7  /* decode: metadata / #CHROM validation (compress.cpp:1130-1170)     */

/* utils.hpp:44-55 -- token byte layout */
#define TOK_00   0x00u  /* 0xxxxxxx : 0|0 x count (1..127)  */
#define TOK_11   0x80u  /* 100xxxxx : 1|1 x count (1..31)   */
#define TOK_01   0xA0u  /* 101xxxxx : 0|1                   */
#define TOK_10   0xC0u  /* 110xxxxx : 1|0                   */
#define TOK_LIT  0xE0u  /* 111xxxxx : literal, x = columns  */


/* utils.hpp:177-186 (serialize) / compress.cpp:96-100,194-199: 30-bit BE length, top bits 11 */
static void put_len_header(uint8_t *dst, uint32_t v) {
    dst[0] = (uint8_t)(((v >> 24) & 0xFF) | 0xC0);
    dst[1] = (uint8_t)((v >> 16) & 0xFF);
    dst[2] = (uint8_t)((v >> 8) & 0xFF);
    dst[3] = (uint8_t)(v & 0xFF);
}

/* utils.hpp:188-231 (deserialize): returns -1 when the extension tag is not 3 */
static int64_t get_len_header(const uint8_t *src) {
    if ((src[0] >> 6) != 3) return -1;
    return ((int64_t)(src[0] & 0x3F) << 24) | ((int64_t)src[1] << 16) | ((int64_t)src[2] << 8) | src[3];
}

/* class of one sample term: 0..3 = 0|0,0|1,1|0,1|1 ; 4 = literal (compress.cpp:129,145,171) */
static int gt_class(const uint8_t *p, size_t n) {
    if (n != 3 || p[1] != '|') return 4;
    if ((p[0] != '0' && p[0] != '1') || (p[2] != '0' && p[2] != '1')) return 4;
    return ((p[0] - '0') << 1) | (p[2] - '0');
}

static const uint8_t class_flag[4] = { TOK_00, TOK_01, TOK_10, TOK_11 };
static const unsigned class_max[4] = { 127, 31, 31, 31 };

/*
 * One data line (no trailing '\n' in [line, line+len)) -> compressed line.
 * Restates compress_data_line (compress.cpp:5-203) with add_newline = true and the
 * tokeniser split_string(line, "\t") (utils.cpp:82-116: empty terms are dropped).
 * Returns the number of bytes written, or a negative VCFC_E_* code.
 */
long vcfc_oracle_compress_line(const uint8_t *line, size_t len, uint8_t *out, size_t cap) {
    /* pass 1: count terms (utils.cpp:88-108) */
    size_t nterms = 0, i = 0;
    while (i < len) {
        while (i < len && line[i] == '\t') i++;
        if (i >= len) break;
        nterms++;
        while (i < len && line[i] != '\t') i++;
    }
    if (nterms < 8) return VCFC_E_TOOFEW;          /* compress.cpp:9-11 */
    if (nterms == 8) return VCFC_E_EIGHTCOLS;      /* compress.cpp:88-89,106: size_t underflow -> abort */
    if (cap < 8 + len + nterms + 2) return VCFC_E_CAP;

    size_t o = 8;                                   /* compress.cpp:32-49: two placeholder headers */
    size_t t = 0;
    i = 0;
    /* required columns 1-8 and FORMAT joined by single tabs (compress.cpp:51-86) */
    while (t < 9) {
        while (line[i] == '\t') i++;
        if (t > 0) out[o++] = '\t';
        while (i < len && line[i] != '\t') out[o++] = line[i++];
        t++;
    }
    size_t nsamples = nterms - 9;
    if (nsamples > 0) out[o++] = '\t';              /* compress.cpp:88-93 */
    put_len_header(out + 4, (uint32_t)(o - 8));     /* compress.cpp:96-100 */

    /* sample run-length loop (compress.cpp:124-186) */
    int run_class = -1;
    unsigned run_count = 0;
    size_t s = 0;
    while (s < nsamples) {
        while (line[i] == '\t') i++;
        size_t b = i;
        while (i < len && line[i] != '\t') i++;
        int c = gt_class(line + b, i - b);
        if (run_class >= 0 && (c != run_class || run_count == class_max[run_class])) {
            out[o++] = (uint8_t)(class_flag[run_class] | run_count);
            run_class = -1;
        }
        if (c == 4) {
            out[o++] = TOK_LIT | 1;                 /* compress.cpp:179-181 */
            memcpy(out + o, line + b, i - b);
            o += i - b;
            if (s + 1 < nsamples) out[o++] = '\t';  /* compress.cpp:182-184 */
        } else if (run_class < 0) {
            run_class = c;
            run_count = 1;
        } else {
            run_count++;
        }
        s++;
    }
    if (run_class >= 0) out[o++] = (uint8_t)(class_flag[run_class] | run_count);
    out[o++] = '\n';                                /* compress.cpp:188-190 */
    put_len_header(out, (uint32_t)(o - 4));         /* compress.cpp:194-199 */
    return (long)o;
}

/*
 * Block of data lines (the region of a VCF after the #CHROM line), restating the
 * per-line part of compress() (compress.cpp:218-251): '\n'-separated, empty lines
 * dropped (219-221), a missing final '\n' is tolerated (std::getline).  Lines starting
 * with '#' are NOT special-cased here: the block API carries data lines only.
 * line_offsets (nullable) receives the output offset of every encoded line.
 * On error returns the code and sets *err_line to the 0-based index of the data line.
 */
int vcfc_oracle_compress_block(const uint8_t *in, size_t in_len, uint8_t *out, size_t cap,
                               size_t *out_len, uint64_t *line_offsets, size_t *n_lines,
                               size_t *err_line) {
    size_t pos = 0, o = 0, nl = 0;
    while (pos < in_len) {
        const uint8_t *e = (const uint8_t *)memchr(in + pos, '\n', in_len - pos);
        size_t llen = e ? (size_t)(e - (in + pos)) : in_len - pos;
        if (llen > 0) {
            long r = vcfc_oracle_compress_line(in + pos, llen, out + o, cap - o);
            if (r < 0) { if (err_line) *err_line = nl; *out_len = o; *n_lines = nl; return (int)r; }
            if (line_offsets) line_offsets[nl] = o;
            o += (size_t)r;
            nl++;
        }
        pos += llen + 1;
    }
    *out_len = o;
    *n_lines = nl;
    return VCFC_OK;
}

/*
 * One compressed line -> text.  Restates decompress2_data_line (compress.cpp:741-986).
 * in points at the first length header.  Returns 1 = line decoded, 0 = clean EOF
 * (fewer than 8 header bytes left, compress.cpp:770-777), negative = error.
 * *consumed = compressed bytes used, *produced = text bytes written (incl. '\n').
 */
int vcfc_oracle_decompress_line(const uint8_t *in, size_t in_len, uint64_t sample_count,
                                uint8_t *out, size_t cap, size_t *consumed, size_t *produced) {
    if (in_len < 8) return 0;
    int64_t line_length = get_len_header(in);
    int64_t req = get_len_header(in + 4);
    if (line_length < 0 || req < 0) return VCFC_E_FORMAT;   /* utils.hpp:201-206 */
    size_t p = 8, o = 0;
    if ((size_t)req > in_len - p) return VCFC_E_TRUNC;      /* compress.cpp:792-798 */
    if ((size_t)req == 0) return VCFC_E_FORMAT;             /* fread of 0 bytes -> throw (792) */
    if (cap < (size_t)req) return VCFC_E_CAP;
    /* A NUL byte here would truncate the reference's linebuf.append(buf) (compress.cpp:807);
     * VCF text has none, and this restatement passes it through unchanged. */
    size_t tabs = 0;
    for (size_t k = 0; k < (size_t)req; k++) {
        uint8_t c = in[p + k];
        if (c == '\t') tabs++;
        out[o++] = c;
    }
    p += (size_t)req;
    if (tabs != 9 && !(tabs == 8 && sample_count == 0)) return VCFC_E_FORMAT; /* compress.cpp:820-828 */

    uint64_t ns = 0;
    while (ns < sample_count) {                              /* compress.cpp:832 */
        if (p >= in_len) return VCFC_E_TRUNC;
        uint8_t b = in[p++];
        if ((b & 0x80) == 0) {                               /* 0|0 run, compress.cpp:843-868 */
            unsigned cnt = b & 0x7F;
            if (cap - o < (size_t)cnt * 4) return VCFC_E_CAP;
            for (unsigned k = 0; k < cnt; k++) { memcpy(out + o, "0|0\t", 4); o += 4; }
            ns += cnt;
            if (ns >= sample_count) {
                if (o == 0) return VCFC_E_FORMAT;
                o--;                                         /* pop_back of the last tab */
            }
        } else if ((b & 0xE0) == 0xE0) {                     /* literal, compress.cpp:869-906 */
            unsigned ncols = b & 0x1F, u = 0;
            while (u < ncols) {
                if (p >= in_len) return VCFC_E_TRUNC;
                uint8_t c = in[p++];
                if (c == '\n') {
                    u++; ns++;
                    if (u != ncols) return VCFC_E_FORMAT;
                    p--;                                     /* fseek(-1): newline handled below */
                } else if (c == '\t') {
                    u++; ns++;
                    if (ns < sample_count) { if (o >= cap) return VCFC_E_CAP; out[o++] = c; }
                } else {
                    if (o >= cap) return VCFC_E_CAP;
                    out[o++] = c;
                }
            }
        } else {                                             /* 0|1, 1|0, 1|1, compress.cpp:907-953 */
            const char *gt = (b & 0xE0) == TOK_01 ? "0|1" : (b & 0xE0) == TOK_10 ? "1|0" : "1|1";
            unsigned cnt = b & 0x1F;
            if (cap - o < (size_t)cnt * 4) return VCFC_E_CAP;
            while (cnt--) {
                memcpy(out + o, gt, 3); o += 3;
                ns++;
                if (ns < sample_count) out[o++] = '\t';
            }
        }
    }
    if (p >= in_len) return VCFC_E_TRUNC;                    /* compress.cpp:958-960 */
    if (in[p++] != '\n') return VCFC_E_FORMAT;               /* compress.cpp:961-966 */
    if (o >= cap) return VCFC_E_CAP;
    out[o++] = '\n';
    *consumed = p;
    *produced = o;
    (void)line_length;                                       /* the reference decoder never uses it */
    return 1;
}

/* Block of compressed lines -> text; the per-line loop of decompress2_fd (compress.cpp:1236-1250). */
int vcfc_oracle_decompress_block(const uint8_t *in, size_t in_len, uint64_t sample_count,
                                 uint8_t *out, size_t cap, size_t *out_len, size_t *n_lines,
                                 size_t *err_line) {
    size_t p = 0, o = 0, nl = 0;
    for (;;) {
        size_t c = 0, w = 0;
        int r = vcfc_oracle_decompress_line(in + p, in_len - p, sample_count, out + o, cap - o, &c, &w);
        if (r == 0) break;
        if (r < 0) { if (err_line) *err_line = nl; *out_len = o; *n_lines = nl; return r; }
        p += c; o += w; nl++;
    }
    *out_len = o;
    *n_lines = nl;
    return VCFC_OK;
}

/*
 * Header region of a .vcf / .vcfc (identical in both: compress.cpp:222-238 writes '#'
 * lines verbatim + "\n").  Restates decompress2_metadata_headers_fd (compress.cpp:1108-1211):
 * >=1 "##" line, then exactly one "#" line; sample_count = tabs beyond the 8th in it.
 * Returns the byte length of the header region or a negative code.
 */
long vcfc_oracle_parse_headers(const uint8_t *in, size_t in_len, uint64_t *sample_count) {
    size_t p = 0;
    int got_meta = 0, got_header = 0;
    uint64_t sc = 0;
    for (;;) {
        /* compress.cpp:1136-1153: at EOF the read fails but the stale first byte is still '#', so the
         * reference throws either way ("missing headers" / "row after already reading a header"):
         * a file without data lines does not decode [probed, SURVEY.md 8a]. */
        if (p >= in_len) return VCFC_E_HEADER;
        if (in[p] != '#') { if (!got_meta || !got_header) return VCFC_E_HEADER; break; }
        if (got_header) return VCFC_E_HEADER;
        if (p + 1 >= in_len) return VCFC_E_HEADER;
        if (in[p + 1] == '#') { got_meta = 1; }
        else { if (!got_meta) return VCFC_E_HEADER; got_header = 1; }
        size_t q = p + 2, tabs = 0;
        for (;;) {
            if (q >= in_len) return VCFC_E_HEADER;
            uint8_t c = in[q++];
            if (c == '\n') break;
            if (got_header && c == '\t') { tabs++; if (tabs > 8) sc++; }
        }
        p = q;
    }
    *sample_count = sc;
    return (long)p;
}

/* ---- binned index (.vcfci), the next row of the scope table --------------------------
 * Restates create_binned_index4 (main.cpp:1284-1637) on a whole .vcfc file held in memory.
 * Entry = {u8 reference index, u32 position, u64 byte offset of the line in the file}, 13 bytes,
 * native little-endian (write_index_entry, main.cpp:600-626).  Not used by the product yet. */

/* strtoul-based integer parse of utils.cpp:152-175: whole field must be consumed; leading blanks,
 * '+', '-' accepted (a '-' negates modulo 2^64).  Returns 0 on success. */
static int parse_ul_field(const uint8_t *p, size_t n, long *out) {
    char buf[64];
    char *end = NULL;
    if (n >= sizeof buf) return -1;
    memcpy(buf, p, n);
    buf[n] = 0;
    if (strlen(buf) != n) return -1;                       /* an embedded NUL ends the C string early */
    *out = (long)strtoul(buf, &end, 10);
    return end == buf + n ? 0 : -1;
}

/* reference_name_map (utils.hpp:90-103, utils.cpp:16-25): "1".."22","X","Y","M" -> 1..25, else 0 */
static uint8_t ref_name_index(const uint8_t *p, size_t n) {
    if (n == 1 && p[0] == 'X') return 23;
    if (n == 1 && p[0] == 'Y') return 24;
    if (n == 1 && p[0] == 'M') return 25;
    if (n == 1 && p[0] >= '1' && p[0] <= '9') return (uint8_t)(p[0] - '0');
    if (n == 2 && p[0] >= '1' && p[0] <= '2' && p[1] >= '0' && p[1] <= '9') {
        int v = 10 * (p[0] - '0') + (p[1] - '0');
        return v <= 22 ? (uint8_t)v : 0;
    }
    return 0;
}

/* value of key (exact match) in a ';'-separated key=value list, parse_kvp semantics (main.cpp:737-757):
 * empty pairs are skipped (split_string drops empty terms, utils.cpp:95), "k" alone has the value "",
 * "k=a=b" is an error (-1), a later duplicate wins.  Returns 1 found, 0 absent, -1 malformed list. */
static int kvp_lookup(const uint8_t *info, size_t n, const char *key, const uint8_t **val, size_t *val_len) {
    size_t klen = strlen(key), i = 0;
    int found = 0;
    while (i < n) {
        size_t j = i;
        while (j < n && info[j] != ';') j++;
        if (j > i) {
            /* split the pair on '=' dropping empty parts */
            const uint8_t *part[3];
            size_t plen[3];
            int np = 0;
            size_t a = i;
            while (a < j) {
                size_t b = a;
                while (b < j && info[b] != '=') b++;
                if (b > a) {
                    if (np == 2) return -1;
                    part[np] = info + a; plen[np] = b - a; np++;
                }
                a = b + 1;
            }
            if (np == 0) return -1;                          /* "=" alone: parts.size() == 0 -> throws */
            if (plen[0] == klen && memcmp(part[0], key, klen) == 0) {
                found = 1;
                if (np == 2) { *val = part[1]; *val_len = plen[1]; }
                else { *val = part[0]; *val_len = 0; }
            }
        }
        i = j + 1;
    }
    return found;
}

/* compute_end_position (main.cpp:763-852).  Returns 0, or -1 where the reference throws. */
static int end_position(long pos, size_t ref_len, const uint8_t *alt, size_t alt_len,
                        const uint8_t *info, size_t info_len, long *end_out) {
    if (memchr(alt, '<', alt_len) != NULL) {                /* alt_is_structural, main.cpp:759-761 */
        const uint8_t *v = NULL;
        size_t vl = 0;
        int r = kvp_lookup(info, info_len, "END", &v, &vl);
        if (r < 0) return -1;
        if (r == 1) {
            long max_end = 0;
            size_t i = 0;
            while (i < vl) {                                /* split on ',' dropping empty terms */
                size_t j = i;
                long e;
                while (j < vl && v[j] != ',') j++;
                if (j > i) {
                    if (parse_ul_field(v + i, j - i, &e) != 0) return -1;
                    if (e > max_end) max_end = e;
                }
                i = j + 1;
            }
            *end_out = max_end < 0 ? -max_end : max_end;
            return 0;
        }
        r = kvp_lookup(info, info_len, "SVLEN", &v, &vl);
        if (r < 0) return -1;
        if (r == 1) {
            long max_len = 0;
            size_t i = 0;
            while (i < vl) {
                size_t j = i;
                long e;
                while (j < vl && v[j] != ',') j++;
                if (j > i) {
                    if (parse_ul_field(v + i, j - i, &e) != 0) return -1;
                    if (e < 0) e = -e;
                    if (e > max_len) max_len = e;
                }
                i = j + 1;
            }
            *end_out = pos + max_len - 1;
            return 0;
        }
        *end_out = pos;
        return 0;
    }
    {
        size_t max_alt = 0, i = 0;
        while (i < alt_len) {                               /* longest ',' separated ALT allele */
            size_t j = i;
            while (j < alt_len && alt[j] != ',') j++;
            if (j - i > max_alt) max_alt = j - i;
            i = j + 1;
        }
        *end_out = pos + (long)(ref_len >= max_alt ? ref_len : max_alt) - 1;
    }
    return 0;
}

/* Columns 1-8 of one line (in points at CHROM, i.e. just behind the two length headers; reading may run up to in + n):
 * END position and chromosome index as create_binned_index4 computes them (main.cpp:1371-1423).  0, or a negative code
 * where the reference throws. */
int vcfc_oracle_line_index_fields(const uint8_t *in, size_t n, long *end_out, uint8_t *ref_out) {
    const uint8_t *f[8];
    size_t fl[8], q = 0;
    int k;
    long pos;
    for (k = 0; k < 8; k++) {
        size_t e = q;
        while (e < n && in[e] != '\t') e++;
        if (e >= n) return VCFC_E_TRUNC;
        f[k] = in + q; fl[k] = e - q;
        q = e + 1;
    }
    if (parse_ul_field(f[1], fl[1], &pos) != 0) return VCFC_E_FORMAT;
    if (end_position(pos, fl[3], f[4], fl[4], f[7], fl[7], end_out) != 0) return VCFC_E_FORMAT;
    *ref_out = ref_name_index(f[0], fl[0]);
    return 0;
}

/*
 * Whole .vcfc file -> .vcfci bytes.  Walks the compressed lines by their length headers, reads only
 * columns 1-8 of each required section, and applies the bin rule of main.cpp:1430-1470: a line whose
 * number is a multiple of entries_per_bin opens a new entry if its END exceeds the last entry's position,
 * every other line can only grow the last entry's position (compared as u32, never looking at the
 * chromosome).  Returns the number of entries or a negative code.
 */
long vcfc_oracle_build_binned_index(const uint8_t *in, size_t in_len, uint64_t entries_per_bin,
                                    uint8_t *out, size_t cap, size_t *out_len) {
    uint64_t sample_count = 0, line_number = 0;
    long hdr = vcfc_oracle_parse_headers(in, in_len, &sample_count);
    size_t p, n_entries = 0;
    if (hdr < 0) return hdr;
    if (entries_per_bin == 0) return VCFC_E_FORMAT;         /* the reference divides by it */
    p = (size_t)hdr;
    *out_len = 0;
    while (in_len - p >= 8) {                               /* fewer than 8 bytes left: EOF (compress.cpp:270-330) */
        int64_t ll = get_len_header(in + p), rq = get_len_header(in + p + 4);
        size_t line_end;
        long endp;
        uint8_t idx;
        int frc;
        if (ll < 0 || rq < 0) return VCFC_E_FORMAT;
        line_end = p + 4 + (size_t)ll;
        if (line_end > in_len) return VCFC_E_TRUNC;
        frc = vcfc_oracle_line_index_fields(in + p + 8, in_len - (p + 8), &endp, &idx);   /* read_to x 8 etc. (main.cpp:1371-1423) */
        if (frc != 0) return frc;
        if (n_entries == 0) {
            if (cap < 13) return VCFC_E_CAP;
            out[0] = idx;
            { uint32_t v = (uint32_t)endp; memcpy(out + 1, &v, 4); }
            { uint64_t v = (uint64_t)p; memcpy(out + 5, &v, 8); }
            n_entries = 1;
        } else {
            uint8_t *last = out + 13 * (n_entries - 1);
            uint32_t last_end;
            memcpy(&last_end, last + 1, 4);
            if ((unsigned long)endp > (unsigned long)last_end) {
                if (line_number % entries_per_bin == 0) {
                    uint8_t *e = out + 13 * n_entries;
                    if (cap < 13 * (n_entries + 1)) return VCFC_E_CAP;
                    e[0] = idx;
                    { uint32_t v = (uint32_t)endp; memcpy(e + 1, &v, 4); }
                    { uint64_t v = (uint64_t)p; memcpy(e + 5, &v, 8); }
                    n_entries++;
                } else {
                    uint32_t v = (uint32_t)endp;
                    memcpy(last + 1, &v, 4);
                }
            }
        }
        line_number++;
        p = line_end;
    }
    *out_len = 13 * n_entries;
    return (long)n_entries;
}

/* ---- file-level drivers (compress.cpp:205-257 and 1214-1257), for the CPU baseline ---- */

static uint8_t *read_file(const char *path, size_t *len) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *buf = (uint8_t *)malloc((size_t)n + 1);
    if (!buf) { fclose(f); return NULL; }
    if (n > 0 && fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); fclose(f); return NULL; }
    fclose(f);
    *len = (size_t)n;
    return buf;
}

/* compress(): '#' lines pass through with "\n"; empty lines dropped; the rest encoded. */
int vcfc_oracle_compress_file(const char *in_path, const char *out_path) {
    size_t n = 0;
    uint8_t *in = read_file(in_path, &n);
    if (!in) return VCFC_E_IO;
    FILE *fo = fopen(out_path, "wb");
    if (!fo) { free(in); return VCFC_E_IO; }
    size_t cap = 1 << 20;
    uint8_t *lb = (uint8_t *)malloc(cap);
    size_t pos = 0;
    int rc = VCFC_OK;
    while (pos < n) {
        const uint8_t *e = (const uint8_t *)memchr(in + pos, '\n', n - pos);
        size_t llen = e ? (size_t)(e - (in + pos)) : n - pos;
        if (llen == 0) { pos += 1; continue; }
        if (in[pos] == '#') {                               /* compress.cpp:222-238 */
            fwrite(in + pos, 1, llen, fo);
            fputc('\n', fo);
        } else {
            size_t need = 2 * llen + 64;
            if (need > cap) { cap = need; lb = (uint8_t *)realloc(lb, cap); }
            long r = vcfc_oracle_compress_line(in + pos, llen, lb, cap);
            if (r < 0) { rc = (int)r; break; }
            fwrite(lb, 1, (size_t)r, fo);
        }
        pos += llen + 1;
    }
    free(lb); free(in); fclose(fo);
    return rc;
}

int vcfc_oracle_decompress_file(const char *in_path, const char *out_path) {
    size_t n = 0;
    uint8_t *in = read_file(in_path, &n);
    if (!in) return VCFC_E_IO;
    uint64_t sc = 0;
    long h = vcfc_oracle_parse_headers(in, n, &sc);
    if (h < 0) { free(in); return (int)h; }
    FILE *fo = fopen(out_path, "wb");
    if (!fo) { free(in); return VCFC_E_IO; }
    fwrite(in, 1, (size_t)h, fo);
    size_t cap = 1 << 20;
    uint8_t *lb = (uint8_t *)malloc(cap);
    size_t p = (size_t)h;
    int rc = VCFC_OK;
    for (;;) {
        size_t c = 0, w = 0;
        int r;
        for (;;) {
            r = vcfc_oracle_decompress_line(in + p, n - p, sc, lb, cap, &c, &w);
            if (r != VCFC_E_CAP) break;
            cap *= 2; lb = (uint8_t *)realloc(lb, cap);
        }
        if (r == 0) break;
        if (r < 0) { rc = r; break; }
        fwrite(lb, 1, w, fo);
        p += c;
    }
    free(lb); free(in); fclose(fo);
    return rc;
}
