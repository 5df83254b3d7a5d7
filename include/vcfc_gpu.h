/*
 * vcfc_gpu.h -- C ABI of the B200 (sm_100a) genotype-column codec for the .vcfc format of
 * theferrit32/vcf-compression.
 *
 * The reference has no FFI of its own; its seam for this path is the set of C++ functions
 * in src/compress.hpp:17-56.  A per-line call cannot feed a GPU, so the ABI is
 * block-granular: one call encodes / decodes a block of data lines.  Each entry point
 * names the reference interface it replaces.  Plain pointers and sizes only -- no CUDA,
 * torch or STL types cross this boundary (streams travel as void*).
 *
 * Byte contract: for the same data lines, vcfc_encode_block writes exactly the bytes the
 * reference's compress() (src/compress.cpp:205-257) writes for those lines, and
 * vcfc_decode_block writes exactly the bytes decompress2_fd() (src/compress.cpp:1214-1257)
 * writes.  There is no CPU fallback: every compute entry point returns VCFC_E_CUDA when no
 * sm_100-class device is usable.
 */
#ifndef VCFC_GPU_H
#define VCFC_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vcfc_ctx vcfc_ctx;

/* Status codes (0 = success).  Where the reference throws / aborts, we return a code. */
enum {
    VCFC_OK          = 0,
    VCFC_E_TOOFEW    = 1,  /* data line with <8 terms: VcfValidationError, compress.cpp:9-11        */
    VCFC_E_EIGHTCOLS = 2,  /* data line with exactly 8 terms: reference aborts, compress.cpp:88-106 */
    VCFC_E_CAP       = 3,  /* output buffer too small                                               */
    VCFC_E_FORMAT    = 4,  /* decode: malformed compressed line (compress.cpp:825,877,965; utils.hpp:201) */
    VCFC_E_TRUNC     = 5,  /* decode: input ends inside a line (compress.cpp:793,839,887,959)       */
    VCFC_E_IO        = 6,  /* file drivers: open/read/write failed                                  */
    VCFC_E_HEADER    = 7,  /* decode: '##'/'#CHROM' validation failed (compress.cpp:1136-1170)      */
    VCFC_E_CUDA      = 8,  /* CUDA runtime / driver error, or no usable device                      */
    VCFC_E_ARG       = 9,  /* invalid argument                                                      */
    VCFC_E_LINE2BIG  = 10, /* a compressed line would exceed the 30-bit length header (utils.hpp:151) */
    VCFC_E_QUERY     = 11  /* malformed REF:START-END string (main.cpp:3993-4026)                   */
};

/* Result of one block operation.  err_line = 0-based index of the first offending data line. */
typedef struct vcfc_result {
    int32_t  status;
    int32_t  reserved;
    uint64_t out_len;
    uint64_t n_lines;
    uint64_t err_line;
} vcfc_result;

/* Context: one per (process, GPU).  Owns workspace, pinned staging and two streams.
 * Threading: a context is used by ONE thread at a time (the library keeps no process-global mutable state, so
 * different contexts -- also several on the same device -- may be driven from different threads concurrently). */
int  vcfc_gpu_init(int device, vcfc_ctx **ctx);
void vcfc_gpu_destroy(vcfc_ctx *ctx);
const char *vcfc_strerror(int code);
/* Last CUDA error text seen by this context ("" if none). */
const char *vcfc_last_cuda_error(const vcfc_ctx *ctx);

/* Worst-case encoded size of in_len bytes of data lines (8-byte headers + 1 escape byte per
 * sample; compress.cpp:32-49,179-184). */
size_t vcfc_encode_bound(size_t in_len);

/*
 * Encode a block of data lines.  Replaces the per-line loop of compress()
 * (src/compress.cpp:218-251) calling compress_data_line (src/compress.cpp:5-203).
 *   in        data lines only (no '#' lines), '\n'-separated; empty lines are dropped and a
 *             missing final '\n' is supplied, as std::getline + compress.cpp:219-221,188 do.
 *   out       receives the .vcfc bytes of those lines, in order.
 *   line_out_offsets  nullable; receives the out offset of each encoded line (what
 *             create_binned_index4, main.cpp:1284-1637, derives by walking headers).
 * Host-pointer form: H2D, kernels, D2H and a stream sync happen inside the call.
 */
int vcfc_encode_block(vcfc_ctx *ctx, const uint8_t *in, size_t in_len,
                      uint8_t *out, size_t out_cap, size_t *out_len,
                      uint64_t *line_out_offsets, size_t line_cap, size_t *n_lines,
                      uint64_t *err_line);

/*
 * Device-pointer form: all work is queued on `stream` (a cudaStream_t passed as void*; NULL = the
 * context's stream), input and output stay in device memory, and d_result (device memory,
 * sizeof(vcfc_result)) is written by the last kernel.  The call is NOT fully asynchronous: before
 * it returns it synchronises `stream` once to read the block's status, because a block outside the
 * tile kernels' grammar is rerun on the generic kernels inside the same call, and the first block with
 * sample columns that are not 3 bytes wide is relaunched once on the encoder instantiation that
 * carries the term walkers (the context then stays with it; see vcfc_last_reject_reason).  The decode form
 * synchronises twice more to size its line table and tile map.  The status words travel through mapped
 * pinned memory, not through the copy engine, so bulk copies on other streams do not delay them.  When it returns, d_result is
 * final; read it with vcfc_fetch_result or your own copy.  A context is used by one host thread
 * at a time; use one context per GPU and per concurrent caller.
 */
int vcfc_encode_block_dev(vcfc_ctx *ctx, const uint8_t *d_in, size_t in_len,
                          uint8_t *d_out, size_t out_cap,
                          uint64_t *d_line_out_offsets, size_t line_cap,
                          vcfc_result *d_result, void *stream);

/*
 * Decode a block of compressed data lines.  Replaces the per-line loop of decompress2_fd()
 * (src/compress.cpp:1236-1250) calling decompress2_data_line (src/compress.cpp:741-986).
 *   in            the .vcfc bytes after the header region, starting at a line header.
 *   sample_count  schema.sample_count (compress.cpp:1190-1194).
 *   out           receives the decoded text lines.
 */
int vcfc_decode_block(vcfc_ctx *ctx, const uint8_t *in, size_t in_len, uint64_t sample_count,
                      uint8_t *out, size_t out_cap, size_t *out_len, size_t *n_lines,
                      uint64_t *err_line);
int vcfc_decode_block_dev(vcfc_ctx *ctx, const uint8_t *d_in, size_t in_len, uint64_t sample_count,
                          uint8_t *d_out, size_t out_cap,
                          vcfc_result *d_result, void *stream);
/* Exact decoded size of a block without writing it (device-pointer input, synchronous). */
int vcfc_decode_size_dev(vcfc_ctx *ctx, const uint8_t *d_in, size_t in_len, uint64_t sample_count,
                         vcfc_result *h_result, void *stream);

/* Synchronise `stream` and copy a device-side result to the host. */
int vcfc_fetch_result(vcfc_ctx *ctx, const vcfc_result *d_result, vcfc_result *h_result, void *stream);

/*
 * Header region shared by .vcf and .vcfc files.  Replaces decompress2_metadata_headers_fd
 * (src/compress.cpp:1108-1211): >=1 "##" line then one "#" line; *sample_count = tabs
 * beyond the 8th in it; *header_len = bytes up to the first data line.  Host only.
 */
int vcfc_parse_headers(const uint8_t *buf, size_t len, size_t *header_len, uint64_t *sample_count);

/* File drivers with the reference's verb semantics (main.cpp:4038-4069). */
/* compress(in, out), src/compress.cpp:205-257 */
int vcfc_compress_file(vcfc_ctx *ctx, const char *in_path, const char *out_path);
/* decompress2_fd(in, out), src/compress.cpp:1214-1257 */
int vcfc_decompress_file(vcfc_ctx *ctx, const char *in_path, const char *out_path);
/* The same two verbs over SEVERAL contexts (normally one per GPU of the box; SURVEY.md 8(e), no reference counterpart):
 * the file is cut into newline-aligned (compress) / line-header-aligned (decompress) chunks, every context takes chunks
 * in file order on its own thread, and the host concatenates the chunk outputs by their out_len running sum -- no
 * collective.  The output file is byte-identical to the single-context call.  ctxs[0] also owns the pinned buffers. */
int vcfc_compress_file_multi(vcfc_ctx **ctxs, int n_ctx, const char *in_path, const char *out_path);
int vcfc_decompress_file_multi(vcfc_ctx **ctxs, int n_ctx, const char *in_path, const char *out_path);
/* compress() and create_binned_index4() (main.cpp:1284-1637) in ONE pass: the index's per-line fields (END position,
 * chromosome index) are computed from the encoder's own line offsets while each compressed chunk is still on the device;
 * nothing is re-read.  Writes out_path exactly as vcfc_compress_file does and index_path exactly as
 * vcfc_create_binned_index_file(out_path, ...) would.  If the index cannot be built (the reference's builder throws: bad
 * header region, malformed POS / INFO, '#' line behind a data line) the compressed file stands and the code says why. */
int vcfc_compress_index_file_multi(vcfc_ctx **ctxs, int n_ctx, const char *in_path, const char *out_path,
                                   const char *index_path, uint64_t entries_per_bin, uint64_t *n_entries);
/* query_compressed_file(in, REF:START-END) -> matching lines to out_fd, src/main.cpp:3777-3929 */
int vcfc_query_file(vcfc_ctx *ctx, const char *in_path, const char *region, int out_fd);
/* create_binned_index4(compressed, index, entries_per_bin), src/main.cpp:1284-1637 (CLI verb create-binned-index,
 * main.cpp:4097-4115): writes the .vcfci index -- 13-byte entries {u8 chromosome index, u32 max END position of the
 * bin, u64 byte offset of the bin's first line} -- of a .vcfc file.  Columns 1-8 of every line are read and turned
 * into an END position on the GPU; the sequential bin rule runs on the host.  *n_entries (nullable) = entries written. */
int vcfc_create_binned_index_file(vcfc_ctx *ctx, const char *vcfc_path, const char *index_path, uint64_t entries_per_bin,
                                  uint64_t *n_entries);
/* query_binned_index_binarysearch(compressed, REF:START-END), src/main.cpp:2974-3350 (CLI verb query-binned-index,
 * main.cpp:4117-4143): the index is <vcfc_path>.vcfci.  Lines whose [POS, END] overlaps the query, found from the index
 * entry the reference's binary search ends on, are decoded on the GPU and written to out_fd. */
int vcfc_query_binned_index_file(vcfc_ctx *ctx, const char *vcfc_path, const char *region, int out_fd);

/* Instrumentation: device time (ms, CUDA events on the launch stream) of the kernels of the
 * most recent *_dev call when timing is enabled.  which: 0 = encode kernel, 1 = decode scan
 * kernel, 2 = decode expand kernel. */
int   vcfc_set_timing(vcfc_ctx *ctx, int enabled);
float vcfc_last_kernel_ms(vcfc_ctx *ctx, int which);
/* Number of kernels this context has launched since init (bench.py's gpu_launches). */
uint64_t vcfc_launch_count(const vcfc_ctx *ctx);
/* Which kernels served the most recent block call: 1 = single-pass tile kernels (regular GT-only
 * lines), 2 = generic line-serial kernels (any input the reference accepts).  Both run on the GPU. */
int vcfc_last_path(const vcfc_ctx *ctx);
/* Why the tile kernels last declined a block (0 = never): 1 the input ends inside a line that nothing closes (a trailing
 * tab), 2 cut point not found inside the halo (required section > ~32 KB), 4 line grammar (empty field, < 10 columns),
 * 6 an empty sample column, 7 more than ~3 % of the block walked by the single-lane term walker (literals of many
 * kilobytes), 8 line table overflow, 9 a sample column that is not 3 bytes wide met by the encoder instantiation for regular
 * blocks -- not a hand-over to the generic kernels: the block is relaunched on the instantiation that carries the term
 * walkers, and the context stays with it until eight blocks in a row did not need them. */
int vcfc_last_reject_reason(const vcfc_ctx *ctx);
/* Testing aid: on = 1 routes every block through the generic kernels; on = 2 keeps the tile kernels but makes the
 * decoder use its span-walking expansion kernel even when the fill-and-patch kernel applies; 0 = automatic. */
int vcfc_force_generic(vcfc_ctx *ctx, int on);

/* sparsify_file(in, out), src/sparse.cpp:290-580 (verb sparsify, main.cpp:4073-4085): every compressed line is copied to
 * offset (300,000,000 + POS) * 16384 behind the header of a holey file, with previous / next distance links.  Host only. */
int vcfc_sparsify_file(const char *vcfc_path, const char *sparse_path);
/* query_sparse_file_fd(in, query), src/main.cpp:235-582 (verb sparse-query, main.cpp:4086-4096): "REF:START-END" on a
 * sparsified file; a single position is looked up directly, a range follows the next links from the first record at
 * or behind START.  The lines found are decoded in one GPU block call and written to out_fd. */
int vcfc_sparse_query_file(vcfc_ctx *ctx, const char *sparse_path, const char *region, int out_fd);

#ifdef __cplusplus
}
#endif
#endif /* VCFC_GPU_H */
