// vcfc_internal.h -- internals shared by the translation units of libvcfc_gpu.so.
// Nothing here crosses the C ABI (include/vcfc_gpu.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/vcfc_gpu.h"

namespace vcfc {

// A growable device allocation owned by the context.
struct DevBuf {
    void*  p   = nullptr;
    size_t cap = 0;
};

enum { kTimeEncode = 0, kTimeDecodeScan = 1, kTimeDecodeExpand = 2, kTimeSlots = 3 };

// Which path the last block call took (bench/test introspection; see vcfc_last_path).
enum { kPathNone = 0, kPathFast = 1, kPathGeneric = 2 };

}  // namespace vcfc

struct vcfc_ctx {
    int          device      = -1;
    int          sm_count    = 0;
    cudaStream_t stream      = nullptr;  // default work stream
    cudaStream_t copy_stream[2] = {nullptr, nullptr};
    // workspace (device)
    vcfc::DevBuf ws[12];
    // device staging for the host-pointer API
    vcfc::DevBuf d_in[2], d_out[2];
    vcfc::DevBuf ix[3];                   // per-line index fields of the fused compress + index pass (END, chromosome, error)
    // pinned host buffers of the file pipeline (vcfc_pipeline.cu): a pool that survives across calls
    struct PinBuf { uint8_t* p; size_t cap; };
    std::mutex          pin_mu;
    std::vector<PinBuf> pin_free;
    // launch configuration of the cooperative encoder, per context (= per device)
    int          enc_attr_set = 0;
    int          enc_resident = 0;
    int          enc_odd = 0;         // the encoder instantiation with the odd-width term walkers is in use (vcfc_encode_block_dev)
    int          enc_odd_idle = 0;    // blocks in a row that did not need it
    int          enc_odd_keep = 0;    // VCFC_ENC_FORCE_ODD: never go back to the regular instantiation
    int          dec_attr_set = 0;
    vcfc_result* h_result = nullptr;      // pinned
    vcfc_ctx*    twin = nullptr;          // a second context on the same device, made by the file verbs (two chunks in flight per GPU)
    uint32_t*    h_map = nullptr;         // 256 bytes of mapped pinned memory: small results written by a kernel (fetch_small)
    uint32_t*    d_map = nullptr;         // ... its device address
    vcfc_result* d_result = nullptr;
    // instrumentation
    int          timing      = 0;
    cudaEvent_t  ev[2 * vcfc::kTimeSlots] = {};
    float        last_ms[vcfc::kTimeSlots] = {0, 0, 0};
    int          ev_pending[vcfc::kTimeSlots] = {0, 0, 0};
    uint64_t     launches    = 0;
    int          last_path   = 0;
    int          last_reject = 0;     // why the tile kernels last handed a block to the generic kernels
    int          force_generic = 0;
    char         cuda_err[256] = {0};
};

namespace vcfc {

// Records a CUDA failure in the context; returns VCFC_E_CUDA.
int cuda_fail(vcfc_ctx* ctx, cudaError_t e, const char* what);

#define VCFC_CUDA(ctx, call)                                              \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return ::vcfc::cuda_fail((ctx), e__, #call); \
    } while (0)

// Ensures b has at least `bytes` capacity (contents are NOT preserved).
int dev_reserve(vcfc_ctx* ctx, DevBuf* b, size_t bytes);

// ---- generic (any input; slow, line-serial) path: vcfc_generic.cu ----
// Both synchronise `stream` internally (they size their workspace from device counts).
int encode_generic(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, uint8_t* d_out, size_t out_cap,
                   uint64_t* d_line_out_offsets, size_t line_cap, vcfc_result* d_result,
                   cudaStream_t stream);
int decode_generic(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, uint64_t sample_count,
                   uint8_t* d_out, size_t out_cap, vcfc_result* d_result, bool size_only,
                   cudaStream_t stream);

// Exclusive prefix sum over n uint64 values (in place allowed: out may alias in); writes the
// grand total to d_total (nullable).  `scratch` is grown as needed.
int scan_exclusive_u64(vcfc_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, size_t n,
                       uint64_t* d_total, DevBuf* scratch, cudaStream_t stream);

// ---- fast (regular GT-only lines; single pass, tile-parallel) path ----
// Return VCFC_OK with d_result->status == kStatusIrregular when the input needs the generic path.
constexpr int kStatusIrregular = -100;
// A few words from device memory to the host WITHOUT the copy engine: a one-warp kernel writes them to mapped pinned memory,
// then the stream is synchronised.  (A cudaMemcpyAsync of 40 bytes queues behind whatever bulk copy of the same direction is in
// flight on another stream -- in vcfc_decode_block that is the previous chunk's half gigabyte of text.)  bytes <= 256, multiple of 4.
int fetch_small(vcfc_ctx* ctx, const void* d_src, void* h_dst, size_t bytes, cudaStream_t st);
int encode_fast(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, uint8_t* d_out, size_t out_cap,
                uint64_t* d_line_out_offsets, size_t line_cap, vcfc_result* d_result,
                cudaStream_t stream);
int decode_fast(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, uint64_t sample_count,
                uint8_t* d_out, size_t out_cap, vcfc_result* d_result, bool size_only,
                cudaStream_t stream);

// ---- host-pointer block encode with the per-line fields of the binned index taken from the encoder's own line offsets
//      while the compressed chunk is still on the device (vcfc_api.cu); vcfc_encode_block is this with idx = nullptr ----
struct LineIndexOut {
    std::vector<uint64_t>  offs;     // offset of every encoded line in `out`
    std::vector<long long> ends;     // END position (compute_end_position, main.cpp:763-852)
    std::vector<uint8_t>   refs;     // chromosome index (utils.hpp:90-103)
    std::vector<uint8_t>   errs;     // 0 ok, 1 malformed columns (the reference's index builder throws), 2 truncated
};
int encode_block_host(vcfc_ctx* ctx, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len,
                      uint64_t* line_out_offsets, size_t line_cap, size_t* n_lines, uint64_t* err_line, LineIndexOut* idx);

// ---- per-line fields of the binned index (vcfc_index.cu): END position, chromosome index, error flag for every
//      compressed line whose start offset is listed in d_line_start (offsets into d_in) ----
int index_line_ends(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, const unsigned long long* d_line_start,
                    unsigned long long n_lines, long long* d_end, uint8_t* d_ref, uint8_t* d_err, cudaStream_t stream);

}  // namespace vcfc
