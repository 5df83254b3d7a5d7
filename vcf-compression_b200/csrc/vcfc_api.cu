// vcfc_api.cu -- the C ABI of libvcfc_gpu.so (include/vcfc_gpu.h): context, block codecs in
// host-pointer and device-pointer form, instrumentation.  File drivers live in vcfc_files.cu.
#include <stdlib.h>

#include <algorithm>

#include "vcfc_internal.h"

namespace vcfc {

int cuda_fail(vcfc_ctx* ctx, cudaError_t e, const char* what) {
    if (ctx) snprintf(ctx->cuda_err, sizeof(ctx->cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return VCFC_E_CUDA;
}

int dev_reserve(vcfc_ctx* ctx, DevBuf* b, size_t bytes) {
    if (bytes <= b->cap && b->p) return VCFC_OK;
    if (b->p) {
        VCFC_CUDA(ctx, cudaFree(b->p));
        b->p = nullptr;
        b->cap = 0;
    }
    size_t want = std::max<size_t>(bytes + bytes / 8, 4096);
    want = (want + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&b->p, want);
    if (e != cudaSuccess) {            // retry with the exact size before giving up
        want = (bytes + 255) & ~size_t(255);
        e = cudaMalloc(&b->p, want);
    }
    if (e != cudaSuccess) {
        b->p = nullptr;
        return cuda_fail(ctx, e, "cudaMalloc(workspace)");
    }
    b->cap = want;
    return VCFC_OK;
}

__global__ void k_to_host(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += 32) dst[i] = src[i];
    __threadfence_system();
}

int fetch_small(vcfc_ctx* ctx, const void* d_src, void* h_dst, size_t bytes, cudaStream_t st) {
    if (bytes > 256 || (bytes & 3)) return VCFC_E_ARG;
    k_to_host<<<1, 32, 0, st>>>(ctx->d_map, (const uint32_t*)d_src, (int)(bytes / 4));
    ctx->launches++;
    VCFC_CUDA(ctx, cudaGetLastError());
    VCFC_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(h_dst, ctx->h_map, bytes);
    return VCFC_OK;
}

static size_t env_size(const char* name, size_t dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    char* end = nullptr;
    unsigned long long x = strtoull(v, &end, 10);
    return end && *end == 0 && x > 0 ? (size_t)x : dflt;
}

}  // namespace vcfc

using namespace vcfc;

extern "C" {

int vcfc_gpu_init(int device, vcfc_ctx** out) {
    if (!out) return VCFC_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return VCFC_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VCFC_E_CUDA;
    if (prop.major < 10) return VCFC_E_CUDA;   // kernels are built for sm_100a only; there is no other path
    vcfc_ctx* ctx = new vcfc_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaStreamCreateWithFlags(&ctx->copy_stream[i], cudaStreamNonBlocking);
    for (int i = 0; i < 2 * kTimeSlots && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&ctx->h_result, 4 * sizeof(vcfc_result));
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->h_map, 256, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&ctx->d_map, ctx->h_map, 0);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_result, 4 * sizeof(vcfc_result));
    if (e != cudaSuccess) {
        vcfc_gpu_destroy(ctx);
        return VCFC_E_CUDA;
    }
    ctx->enc_odd_keep = env_size("VCFC_ENC_FORCE_ODD", 0) != 0;   // (tests / tuning: always the encoder instantiation with the term walkers)
    ctx->enc_odd = ctx->enc_odd_keep;
    *out = ctx;
    return VCFC_OK;
}

void vcfc_gpu_destroy(vcfc_ctx* ctx) {
    if (!ctx) return;
    if (ctx->twin) { vcfc_gpu_destroy(ctx->twin); ctx->twin = nullptr; }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& b : ctx->ws) if (b.p) cudaFree(b.p);
    for (auto& b : ctx->ix) if (b.p) cudaFree(b.p);
    for (int i = 0; i < 2; i++) {
        if (ctx->d_in[i].p) cudaFree(ctx->d_in[i].p);
        if (ctx->d_out[i].p) cudaFree(ctx->d_out[i].p);
        if (ctx->copy_stream[i]) cudaStreamDestroy(ctx->copy_stream[i]);
    }
    for (auto& b : ctx->pin_free) if (b.p) cudaFreeHost(b.p);
    ctx->pin_free.clear();
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    if (ctx->h_map) cudaFreeHost(ctx->h_map);
    if (ctx->d_result) cudaFree(ctx->d_result);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* vcfc_strerror(int code) {
    switch (code) {
        case VCFC_OK: return "ok";
        case VCFC_E_TOOFEW: return "data line has fewer than 8 columns";
        case VCFC_E_EIGHTCOLS: return "data line has exactly 8 columns (the reference aborts on it)";
        case VCFC_E_CAP: return "output buffer too small";
        case VCFC_E_FORMAT: return "malformed compressed line";
        case VCFC_E_TRUNC: return "compressed input ends inside a line";
        case VCFC_E_IO: return "file I/O failed";
        case VCFC_E_HEADER: return "metadata / #CHROM header validation failed";
        case VCFC_E_CUDA: return "CUDA error or no sm_100 device";
        case VCFC_E_ARG: return "invalid argument";
        case VCFC_E_LINE2BIG: return "compressed line exceeds the 30-bit length header";
        case VCFC_E_QUERY: return "malformed query, expected <ref> or <ref>:<start>-<end>";
        default: return "unknown error";
    }
}

const char* vcfc_last_cuda_error(const vcfc_ctx* ctx) { return ctx ? ctx->cuda_err : ""; }

size_t vcfc_encode_bound(size_t in_len) {
    // Every line gains 8 header bytes (compress.cpp:32-49) and possibly a '\n'; every literal sample
    // gains its 0xE1 escape (compress.cpp:179-181).  A line is >= 18 bytes, a sample >= 2 bytes.
    return in_len + in_len / 2 + 9 * (in_len / 18 + 1) + 64;
}

int vcfc_set_timing(vcfc_ctx* ctx, int enabled) {
    if (!ctx) return VCFC_E_ARG;
    ctx->timing = enabled;
    return VCFC_OK;
}
float vcfc_last_kernel_ms(vcfc_ctx* ctx, int which) {
    if (!ctx || which < 0 || which >= kTimeSlots) return -1.f;
    if (ctx->ev_pending[which]) {
        float ms = -1.f;
        if (cudaEventSynchronize(ctx->ev[2 * which + 1]) == cudaSuccess &&
            cudaEventElapsedTime(&ms, ctx->ev[2 * which], ctx->ev[2 * which + 1]) == cudaSuccess)
            ctx->last_ms[which] = ms;
        ctx->ev_pending[which] = 0;
    }
    return ctx->last_ms[which];
}
uint64_t vcfc_launch_count(const vcfc_ctx* ctx) { return ctx ? ctx->launches : 0; }
int vcfc_last_path(const vcfc_ctx* ctx) { return ctx ? ctx->last_path : 0; }
int vcfc_last_reject_reason(const vcfc_ctx* ctx) { return ctx ? ctx->last_reject : 0; }
int vcfc_force_generic(vcfc_ctx* ctx, int on) {
    if (!ctx) return VCFC_E_ARG;
    ctx->force_generic = on;
    return VCFC_OK;
}

int vcfc_fetch_result(vcfc_ctx* ctx, const vcfc_result* d_result, vcfc_result* h_result, void* stream) {
    if (!ctx || !d_result || !h_result) return VCFC_E_ARG;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    return fetch_small(ctx, d_result, h_result, sizeof(vcfc_result), st);
}

// ---- device-pointer forms -----------------------------------------------------------------
static int peek_status(vcfc_ctx* ctx, const vcfc_result* d_result, cudaStream_t st, int* status) {
    int rc = fetch_small(ctx, d_result, ctx->h_result + 2, sizeof(vcfc_result), st);
    if (rc) return rc;
    *status = ctx->h_result[2].status;
    if (*status == kStatusIrregular) ctx->last_reject = ctx->h_result[2].reserved;
    return VCFC_OK;
}

int vcfc_encode_block_dev(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, uint8_t* d_out, size_t out_cap,
                          uint64_t* d_line_out_offsets, size_t line_cap, vcfc_result* d_result, void* stream) {
    if (!ctx || !d_result || (in_len && (!d_in || !d_out))) return VCFC_E_ARG;
    VCFC_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    if (ctx->force_generic != 1) {
        // Two instantiations of the tile kernel: the regular one gives a block with odd-width sample columns (10|0, haploid calls,
        // GT:DP:GQ) up at once (reject reason 9) and the one with the term walkers takes it; the context stays with that one until
        // eight blocks in a row did not need it.
        for (;;) {
            int rc = encode_fast(ctx, d_in, in_len, d_out, out_cap, d_line_out_offsets, line_cap, d_result, st);
            if (rc != VCFC_OK) return rc;
            int status = 0;
            if ((rc = peek_status(ctx, d_result, st, &status))) return rc;
            if (status != kStatusIrregular) {
                if (ctx->enc_odd && status == VCFC_OK) {
                    ctx->enc_odd_idle = ctx->h_result[2].reserved ? 0 : ctx->enc_odd_idle + 1;
                    if (ctx->enc_odd_idle >= 8 && !ctx->enc_odd_keep) { ctx->enc_odd = 0; ctx->enc_odd_idle = 0; }
                }
                return VCFC_OK;
            }
            if (ctx->last_reject == 9 && !ctx->enc_odd) { ctx->enc_odd = 1; ctx->enc_odd_idle = 0; continue; }
            break;
        }
    }
    if (ctx->timing) cudaEventRecord(ctx->ev[2 * kTimeEncode], st);       // generic: the whole pipeline is "the kernel"
    int rc = encode_generic(ctx, d_in, in_len, d_out, out_cap, d_line_out_offsets, line_cap, d_result, st);
    if (ctx->timing) { cudaEventRecord(ctx->ev[2 * kTimeEncode + 1], st); ctx->ev_pending[kTimeEncode] = 1; }
    return rc;
}

static int decode_dev_common(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, uint64_t sample_count, uint8_t* d_out,
                             size_t out_cap, vcfc_result* d_result, bool size_only, void* stream) {
    if (!ctx || !d_result || (in_len && !d_in) || (!size_only && in_len && !d_out)) return VCFC_E_ARG;
    VCFC_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    if (ctx->force_generic != 1) {
        int rc = decode_fast(ctx, d_in, in_len, sample_count, d_out, out_cap, d_result, size_only, st);
        if (rc != VCFC_OK) return rc;
        int status = 0;
        if ((rc = peek_status(ctx, d_result, st, &status))) return rc;
        if (status != kStatusIrregular) return VCFC_OK;
    }
    if (ctx->timing) cudaEventRecord(ctx->ev[2 * kTimeDecodeExpand], st);
    int rc = decode_generic(ctx, d_in, in_len, sample_count, d_out, out_cap, d_result, size_only, st);
    if (ctx->timing) { cudaEventRecord(ctx->ev[2 * kTimeDecodeExpand + 1], st); ctx->ev_pending[kTimeDecodeExpand] = 1; }
    return rc;
}

int vcfc_decode_block_dev(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, uint64_t sample_count, uint8_t* d_out,
                          size_t out_cap, vcfc_result* d_result, void* stream) {
    return decode_dev_common(ctx, d_in, in_len, sample_count, d_out, out_cap, d_result, false, stream);
}

int vcfc_decode_size_dev(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, uint64_t sample_count,
                         vcfc_result* h_result, void* stream) {
    if (!ctx || !h_result) return VCFC_E_ARG;
    int rc = decode_dev_common(ctx, d_in, in_len, sample_count, nullptr, 0, ctx->d_result + 1, true, stream);
    if (rc) return rc;
    return vcfc_fetch_result(ctx, ctx->d_result + 1, h_result, stream);
}

// ---- host-pointer forms: newline-aligned chunks, two device buffer sets, two streams ---------
// Chunk i: H2D -> kernels -> (result) on stream i&1; its output D2H is issued after chunk i+1 has
// been queued on the other stream, so copies in both directions overlap the kernels.
static size_t text_chunk_end(const uint8_t* in, size_t pos, size_t in_len, size_t chunk) {
    if (in_len - pos <= chunk) return in_len;
    const void* r = memrchr(in + pos, '\n', chunk);
    if (r) return (size_t)((const uint8_t*)r - in) + 1;
    const void* f = memchr(in + pos + chunk, '\n', in_len - pos - chunk);   // one line longer than a chunk
    return f ? (size_t)((const uint8_t*)f - in) + 1 : in_len;
}

int vcfc_encode_block(vcfc_ctx* ctx, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len,
                      uint64_t* line_out_offsets, size_t line_cap, size_t* n_lines, uint64_t* err_line) {
    return vcfc::encode_block_host(ctx, in, in_len, out, out_cap, out_len, line_out_offsets, line_cap, n_lines, err_line, nullptr);
}

}  // extern "C"

int vcfc::encode_block_host(vcfc_ctx* ctx, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len,
                            uint64_t* line_out_offsets, size_t line_cap, size_t* n_lines, uint64_t* err_line, LineIndexOut* idx) {
    if (!ctx || (in_len && (!in || !out)) || !out_len) return VCFC_E_ARG;
    VCFC_CUDA(ctx, cudaSetDevice(ctx->device));
    // device chunk: 256 MB for large inputs; a sixteenth of the input (at least 32 MB) for smaller ones, so that the first upload --
    // which nothing overlaps -- stays a small part of the call
    size_t chunk = env_size("VCFC_CHUNK_MB", 0) << 20;
    if (!chunk) chunk = std::min<size_t>((size_t)256 << 20, std::max<size_t>((size_t)32 << 20, ((in_len / 16) + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1)));
    size_t pos = 0, o = 0, lines = 0;
    int status = VCFC_OK;
    uint64_t eline = 0;
    struct Pending { bool live; size_t in_len, line_base; } pend[2] = {{false, 0, 0}, {false, 0, 0}};
    DevBuf* d_lo = &ctx->ws[9];   // two halves: per-slot line offsets

    auto finish = [&](int s) -> int {   // collect slot s: result, then output + line offsets to the host
        if (!pend[s].live) return VCFC_OK;
        pend[s].live = false;
        cudaStream_t st = ctx->copy_stream[s];
        VCFC_CUDA(ctx, cudaStreamSynchronize(st));
        vcfc_result r = ctx->h_result[s];
        if (r.status == VCFC_E_CAP) { status = VCFC_E_CAP; return VCFC_OK; }
        if (o + r.out_len > out_cap) { status = VCFC_E_CAP; return VCFC_OK; }
        if (r.out_len) VCFC_CUDA(ctx, cudaMemcpyAsync(out + o, ctx->d_out[s].p, r.out_len, cudaMemcpyDeviceToHost, st));
        if (line_out_offsets && r.n_lines && lines < line_cap) {
            size_t nl = std::min<size_t>(r.n_lines, line_cap - lines);
            uint64_t* dst = line_out_offsets + lines;
            const uint64_t* src = (const uint64_t*)d_lo->p + (size_t)s * (d_lo->cap / 16);
            VCFC_CUDA(ctx, cudaMemcpyAsync(dst, src, nl * 8, cudaMemcpyDeviceToHost, st));
            VCFC_CUDA(ctx, cudaStreamSynchronize(st));
            for (size_t k = 0; k < nl; k++) dst[k] += o;
        }
        if (idx && r.n_lines) {
            // the binned index's per-line fields, from the encoder's own line offsets, while the chunk is on the device
            const size_t nl = (size_t)r.n_lines, per = d_lo->cap / 16, b0 = idx->ends.size();
            const unsigned long long* d_offs = (const unsigned long long*)d_lo->p + (size_t)s * per;
            long long* d_end = (long long*)ctx->ix[0].p + (size_t)s * per;
            uint8_t* d_ref = (uint8_t*)ctx->ix[1].p + (size_t)s * per;
            uint8_t* d_err = (uint8_t*)ctx->ix[2].p + (size_t)s * per;
            int irc = index_line_ends(ctx, (const uint8_t*)ctx->d_out[s].p, (size_t)r.out_len, d_offs, nl, d_end, d_ref, d_err, st);
            if (irc) return irc;
            idx->offs.resize(b0 + nl); idx->ends.resize(b0 + nl); idx->refs.resize(b0 + nl); idx->errs.resize(b0 + nl);
            VCFC_CUDA(ctx, cudaMemcpyAsync(idx->offs.data() + b0, d_offs, nl * 8, cudaMemcpyDeviceToHost, st));
            VCFC_CUDA(ctx, cudaMemcpyAsync(idx->ends.data() + b0, d_end, nl * 8, cudaMemcpyDeviceToHost, st));
            VCFC_CUDA(ctx, cudaMemcpyAsync(idx->refs.data() + b0, d_ref, nl, cudaMemcpyDeviceToHost, st));
            VCFC_CUDA(ctx, cudaMemcpyAsync(idx->errs.data() + b0, d_err, nl, cudaMemcpyDeviceToHost, st));
            VCFC_CUDA(ctx, cudaStreamSynchronize(st));
            for (size_t k = 0; k < nl; k++) idx->offs[b0 + k] += o;
        }
        if (r.status != VCFC_OK) { status = r.status; eline = lines + r.err_line; }
        o += r.out_len;
        lines += r.n_lines;
        return VCFC_OK;
    };

    int rc = VCFC_OK, i = 0;
    const bool want_offs = (line_out_offsets && line_cap) || idx;
    if (want_offs) {
        size_t per = std::min(chunk, std::max<size_t>(in_len, 1)) / 18 + 2;
        if ((rc = dev_reserve(ctx, d_lo, 2 * per * 8 + 32))) return rc;
        if (idx) {
            const size_t per_cap = d_lo->cap / 16;
            if ((rc = dev_reserve(ctx, &ctx->ix[0], 2 * per_cap * 8 + 32))) return rc;
            if ((rc = dev_reserve(ctx, &ctx->ix[1], 2 * per_cap + 32))) return rc;
            if ((rc = dev_reserve(ctx, &ctx->ix[2], 2 * per_cap + 32))) return rc;
        }
    }
    // Chunk i lives in slot i & 1.  Its upload is queued a whole iteration before the encode call that waits for it (that
    // call synchronises its stream to read the block's status), so while the host sits in encode(i) the upload of chunk
    // i + 1 is already running on the other stream: the host -> device direction never idles between chunks.  Per slot and
    // stream the order is H2D(i), kernels(i), D2H(result, output of i), H2D(i + 2).
    struct Slot { size_t len; bool loaded; } sl[2] = {{0, false}, {0, false}};
    auto load = [&](int s) -> int {     // the next chunk goes to slot s: queue its upload
        sl[s].loaded = false;
        if (pos >= in_len) return VCFC_OK;
        const size_t end = text_chunk_end(in, pos, in_len, chunk), len = end - pos;
        int r = dev_reserve(ctx, &ctx->d_in[s], len + 64);
        if (r) return r;
        cudaError_t e = cudaMemcpyAsync(ctx->d_in[s].p, in + pos, len, cudaMemcpyHostToDevice, ctx->copy_stream[s]);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "H2D");
        sl[s].len = len; sl[s].loaded = true;
        pos = end;
        return VCFC_OK;
    };
    if ((rc = load(0)) == VCFC_OK) rc = load(1);
    while (rc == VCFC_OK && status == VCFC_OK) {
        const int s = i & 1;
        if (!sl[s].loaded) break;
        const size_t len = sl[s].len;
        cudaStream_t st = ctx->copy_stream[s];
        size_t bound = std::min(vcfc_encode_bound(len), out_cap);
        if ((rc = dev_reserve(ctx, &ctx->d_out[s], bound + 64))) break;
        uint64_t* d_offs = nullptr;
        size_t cap_s = 0;
        if (want_offs) {
            cap_s = d_lo->cap / 16;
            d_offs = (uint64_t*)d_lo->p + (size_t)s * cap_s;
        }
        if ((rc = vcfc_encode_block_dev(ctx, (const uint8_t*)ctx->d_in[s].p, len, (uint8_t*)ctx->d_out[s].p, bound, d_offs,
                                        cap_s, ctx->d_result + s, st)))
            break;
        VCFC_CUDA(ctx, cudaMemcpyAsync(ctx->h_result + s, ctx->d_result + s, sizeof(vcfc_result), cudaMemcpyDeviceToHost, st));
        pend[s] = {true, len, lines};
        if ((rc = finish(s))) break;                 // (chunks finish in order: the output position is the running sum)
        if ((rc = load(s))) break;
        i++;
    }
    for (int s = 0; s < 2; s++) cudaStreamSynchronize(ctx->copy_stream[s]);
    *out_len = o;
    if (n_lines) *n_lines = lines;
    if (err_line) *err_line = eline;
    return rc != VCFC_OK ? rc : status;
}

extern "C" {

int vcfc_decode_block(vcfc_ctx* ctx, const uint8_t* in, size_t in_len, uint64_t sample_count, uint8_t* out,
                      size_t out_cap, size_t* out_len, size_t* n_lines, uint64_t* err_line) {
    if (!ctx || (in_len && (!in || !out)) || !out_len) return VCFC_E_ARG;
    VCFC_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t chunk = env_size("VCFC_DCHUNK_MB", 32) << 20;
    size_t pos = 0, o = 0, lines = 0;
    int status = VCFC_OK, rc = VCFC_OK, i = 0;
    uint64_t eline = 0;
    // Two slots on two streams: while slot s copies its decoded text to the host, slot s^1 uploads and decodes the next chunk.
    while (status == VCFC_OK) {
        // chunk end = last line boundary within `chunk` bytes: walk the 4-byte line-length headers
        // (compress.cpp:270-330); a broken header ends the walk and the device reports it.
        size_t end = pos;
        bool broken = false;
        while (in_len - end >= 8) {
            if ((in[end] >> 6) != 3) { broken = true; break; }
            size_t ll = ((size_t)(in[end] & 0x3F) << 24) | ((size_t)in[end + 1] << 16) | ((size_t)in[end + 2] << 8) | in[end + 3];
            if (ll + 4 > in_len - end) { broken = true; break; }
            if (end + 4 + ll - pos > chunk && end > pos) break;
            end += 4 + ll;
        }
        if (end == pos) {
            if (in_len - pos < 8) break;       // clean EOF (compress.cpp:770-777)
            end = in_len;                      // let the device classify the damage
        } else if (broken && in_len - end >= 8) {
            end = in_len;
        } else if (in_len - end < 8) {
            end = in_len;
        }
        const int s = i & 1;
        cudaStream_t st = ctx->copy_stream[s];
        VCFC_CUDA(ctx, cudaStreamSynchronize(st));      // slot s: its previous D2H has finished
        size_t len = end - pos;
        if ((rc = dev_reserve(ctx, &ctx->d_in[s], len + 64))) break;
        VCFC_CUDA(ctx, cudaMemcpyAsync(ctx->d_in[s].p, in + pos, len, cudaMemcpyHostToDevice, st));
        size_t want = std::min(out_cap - o, std::max<size_t>(len * 20, (size_t)1 << 20));
        vcfc_result r;
        memset(&r, 0, sizeof(r));
        for (int attempt = 0; attempt < 2; attempt++) {
            if ((rc = dev_reserve(ctx, &ctx->d_out[s], want + 64))) break;
            if ((rc = vcfc_decode_block_dev(ctx, (const uint8_t*)ctx->d_in[s].p, len, sample_count, (uint8_t*)ctx->d_out[s].p,
                                            want, ctx->d_result + s, st)))
                break;
            if ((rc = vcfc_fetch_result(ctx, ctx->d_result + s, &r, st))) break;
            if (r.status != VCFC_E_CAP || r.out_len > out_cap - o) break;
            want = r.out_len;                  // the device told us the exact size: retry once
        }
        if (rc) break;
        if (r.status == VCFC_E_CAP) { status = VCFC_E_CAP; break; }
        if (r.out_len) VCFC_CUDA(ctx, cudaMemcpyAsync(out + o, ctx->d_out[s].p, r.out_len, cudaMemcpyDeviceToHost, st));
        if (r.status != VCFC_OK) { status = r.status; eline = lines + r.err_line; }
        o += r.out_len;
        lines += r.n_lines;
        pos = end;
        i++;
        if (pos >= in_len) break;
    }
    for (int s = 0; s < 2; s++) cudaStreamSynchronize(ctx->copy_stream[s]);
    *out_len = o;
    if (n_lines) *n_lines = lines;
    if (err_line) *err_line = eline;
    return rc != VCFC_OK ? rc : status;
}

}  // extern "C"
