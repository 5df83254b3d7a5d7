// vcfc_generic.cu -- the generic GPU path: correct for ANY input the reference accepts
// (empty fields, GT:DP:GQ samples, haploid calls, CR-LF, lines without samples ...), one
// thread per line.  It is the route taken when the single-pass tile kernels
// (vcfc_encode_fast.cu / vcfc_decode_fast.cu) report an input outside their regular
// grammar.  It runs on the GPU: there is no CPU fallback anywhere in this library.
//
// Reference behaviour restated here (file:line under /root/reference/src):
//   tokeniser      utils.cpp:82-116        split_string drops empty terms
//   line encoder   compress.cpp:5-203      compress_data_line
//   line decoder   compress.cpp:741-986    decompress2_data_line
//   length header  utils.hpp:134-247       30-bit big-endian, tag bits 11
#include "vcfc_common.cuh"
#include "vcfc_internal.h"

namespace vcfc {

// ------------------------------------------------------------------------------------------
// small block-level helpers
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += n;
    }
    return v;
}

// Exclusive scan over the block (blockDim.x <= 1024, multiple of 32).  Returns the exclusive
// prefix of v; *total receives the block sum.  `sh` needs 33 entries.
template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T* sh, T* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    T inc = warp_incl_scan(v);
    if (lane == 31) sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        T s = lane < nw ? sh[lane] : T(0);
        T si = warp_incl_scan(s);
        sh[lane] = si - s;
        if (lane == 31) sh[32] = si;
    }
    __syncthreads();
    T r = sh[w] + inc - v;
    *total = sh[32];
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------------------------------
// device-wide exclusive scan of uint64 (three small kernels; only used by the generic path)
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

__global__ void k_scan_partials(const uint64_t* __restrict__ in, size_t n, uint64_t* __restrict__ partial) {
    __shared__ unsigned long long sh[33];
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < n) s += in[base + k];
    unsigned long long tot;
    block_excl_scan<unsigned long long>(s, sh, &tot);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void k_scan_single(uint64_t* __restrict__ partial, size_t nb, uint64_t* __restrict__ total) {
    __shared__ unsigned long long sh[33];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < nb; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        unsigned long long v = i < nb ? partial[i] : 0ull, tot;
        unsigned long long ex = block_excl_scan<unsigned long long>(v, sh, &tot);
        if (i < nb) partial[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void k_scan_apply(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, size_t n,
                             const uint64_t* __restrict__ partial) {
    __shared__ unsigned long long sh[33];
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    unsigned long long v[kScanItems], s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = base + k < n ? in[base + k] : 0ull;
        s += v[k];
    }
    unsigned long long tot;
    unsigned long long ex = block_excl_scan<unsigned long long>(s, sh, &tot) + partial[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
}

int scan_exclusive_u64(vcfc_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, size_t n, uint64_t* d_total,
                       DevBuf* scratch, cudaStream_t stream) {
    if (n == 0) {
        if (d_total) VCFC_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(uint64_t), stream));
        return VCFC_OK;
    }
    size_t nb = (n + kScanTile - 1) / kScanTile;
    int rc = dev_reserve(ctx, scratch, nb * sizeof(uint64_t));
    if (rc) return rc;
    uint64_t* partial = (uint64_t*)scratch->p;
    k_scan_partials<<<(unsigned)nb, kScanThreads, 0, stream>>>(d_in, n, partial);
    k_scan_single<<<1, 1024, 0, stream>>>(partial, nb, d_total);
    k_scan_apply<<<(unsigned)nb, kScanThreads, 0, stream>>>(d_in, d_out, n, partial);
    ctx->launches += 3;
    VCFC_CUDA(ctx, cudaGetLastError());
    return VCFC_OK;
}

// ------------------------------------------------------------------------------------------
// raw line table of a text block: LS[0] = 0, LS[r+1] = (position of r-th '\n') + 1
// ------------------------------------------------------------------------------------------
constexpr int kNlThreads = 256, kNlBytesPerThread = 32, kNlChunk = kNlThreads * kNlBytesPerThread;

__device__ __forceinline__ unsigned count_nl(const uint8_t* __restrict__ p, size_t lo, size_t hi) {
    unsigned c = 0;
    for (size_t i = lo; i < hi; i++) c += (p[i] == '\n');
    return c;
}

__global__ void k_count_nl(const uint8_t* __restrict__ in, size_t n, uint64_t* __restrict__ counts) {
    __shared__ unsigned sh[33];
    size_t lo = (size_t)blockIdx.x * kNlChunk + (size_t)threadIdx.x * kNlBytesPerThread;
    size_t hi = lo + kNlBytesPerThread;
    if (hi > n) hi = n;
    unsigned c = lo < n ? count_nl(in, lo, hi) : 0u, tot;
    block_excl_scan<unsigned>(c, sh, &tot);
    if (threadIdx.x == 0) counts[blockIdx.x] = tot;
}

__global__ void k_fill_lines(const uint8_t* __restrict__ in, size_t n, const uint64_t* __restrict__ chunk_base,
                             uint64_t* __restrict__ line_start) {
    __shared__ unsigned sh[33];
    size_t lo = (size_t)blockIdx.x * kNlChunk + (size_t)threadIdx.x * kNlBytesPerThread;
    size_t hi = lo + kNlBytesPerThread;
    if (hi > n) hi = n;
    unsigned c = lo < n ? count_nl(in, lo, hi) : 0u, tot;
    unsigned ex = block_excl_scan<unsigned>(c, sh, &tot);
    if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = 0;
    if (c) {
        uint64_t r = chunk_base[blockIdx.x] + ex;
        for (size_t i = lo; i < hi; i++)
            if (in[i] == '\n') line_start[++r] = i + 1;
    }
}

// ------------------------------------------------------------------------------------------
// per-line encoder (one thread walks one line)
// ------------------------------------------------------------------------------------------
struct TermWalk {
    const uint8_t* p;
    size_t len, i;
    // next maximal non-tab run (utils.cpp:88-108); false when the line is exhausted
    __device__ __forceinline__ bool next(size_t& b, size_t& e) {
        while (i < len && p[i] == '\t') i++;
        if (i >= len) return false;
        b = i;
        while (i < len && p[i] != '\t') i++;
        e = i;
        return true;
    }
};

__device__ __forceinline__ int gt_class(const uint8_t* p, size_t n) {
    if (n != 3 || p[1] != '|') return 4;
    unsigned a = p[0] - '0', b = p[2] - '0';
    if (a > 1u || b > 1u) return 4;
    return (int)((a << 1) | b);
}

__device__ __forceinline__ void put_len_header(uint8_t* dst, uint32_t v) {
    dst[0] = (uint8_t)((v >> 24) | 0xC0);
    dst[1] = (uint8_t)(v >> 16);
    dst[2] = (uint8_t)(v >> 8);
    dst[3] = (uint8_t)v;
}

__device__ __forceinline__ uint32_t class_flag(int c) { return c == 0 ? kTok00 : c == 1 ? kTok01 : c == 2 ? kTok10 : kTok11; }
__device__ __forceinline__ uint32_t class_max(int c) { return c == 0 ? 127u : 31u; }

// Returns the encoded size, or -code.  WRITE=false computes the size only.
template <bool WRITE>
__device__ long long encode_line(const uint8_t* __restrict__ line, size_t len, uint8_t* __restrict__ out) {
    TermWalk it{line, len, 0};
    size_t b = 0, e = 0, o = 8;
    int t = 0;
    bool have = it.next(b, e);
    while (have && t < 9) {                       // compress.cpp:51-86
        if (t > 0) { if (WRITE) out[o] = '\t'; o++; }
        if (WRITE) for (size_t k = b; k < e; k++) out[o + (k - b)] = line[k];
        o += e - b;
        t++;
        have = it.next(b, e);
    }
    if (t < 8) return -(long long)kETooFew;       // compress.cpp:9-11
    if (t == 8) return -(long long)kEEightCols;   // compress.cpp:88-106 (size_t underflow -> abort)
    if (have) { if (WRITE) out[o] = '\t'; o++; }  // compress.cpp:88-93
    if (WRITE) put_len_header(out + 4, (uint32_t)(o - 8));
    int run_class = -1;
    uint32_t run_count = 0;
    while (have) {                                // compress.cpp:124-186
        int c = gt_class(line + b, e - b);
        size_t lb = b, le = e;
        have = it.next(b, e);
        if (run_class >= 0 && (c != run_class || run_count == class_max(run_class))) {
            if (WRITE) out[o] = (uint8_t)(class_flag(run_class) | run_count);
            o++;
            run_class = -1;
        }
        if (c == 4) {
            if (WRITE) {
                out[o] = (uint8_t)(kTokLit | 1);
                for (size_t k = lb; k < le; k++) out[o + 1 + (k - lb)] = line[k];
            }
            o += 1 + (le - lb);
            if (have) { if (WRITE) out[o] = '\t'; o++; }
        } else if (run_class < 0) {
            run_class = c;
            run_count = 1;
        } else {
            run_count++;
        }
    }
    if (run_class >= 0) { if (WRITE) out[o] = (uint8_t)(class_flag(run_class) | run_count); o++; }
    if (WRITE) out[o] = '\n';
    o++;
    if (o - 4 > 0x3FFFFFFFull) return -(long long)kELine2Big;
    if (WRITE) put_len_header(out, (uint32_t)(o - 4));
    return (long long)o;
}

// sizes[k] = encoded size of raw line k (0 for empty lines and lines in error);
// flags[k] = 1 for lines that produce output; first error recorded in err[0] (min raw index) / err codes.
__global__ void k_encode_size(const uint8_t* __restrict__ in, size_t in_len, const uint64_t* __restrict__ line_start,
                              size_t n_raw, size_t n_nl, uint64_t* __restrict__ sizes, uint64_t* __restrict__ flags,
                              uint8_t* __restrict__ codes, unsigned long long* __restrict__ err_line) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_raw) return;
    size_t lo = line_start[k];
    size_t hi = k < n_nl ? line_start[k + 1] - 1 : in_len;
    unsigned long long sz = 0, fl = 0;
    uint8_t code = 0;
    if (hi > lo) {
        long long r = encode_line<false>(in + lo, hi - lo, nullptr);
        if (r < 0) {
            code = (uint8_t)(-r);
            atomicMin(err_line, (unsigned long long)k);
        } else {
            sz = (unsigned long long)r;
            fl = 1;
        }
    }
    sizes[k] = sz;
    flags[k] = fl;
    codes[k] = code;
}

__global__ void k_encode_write(const uint8_t* __restrict__ in, size_t in_len, const uint64_t* __restrict__ line_start,
                               size_t n_raw, size_t n_nl, const uint64_t* __restrict__ offs,
                               const uint64_t* __restrict__ ranks, const uint64_t* __restrict__ flags,
                               size_t raw_limit, uint8_t* __restrict__ out, uint64_t* __restrict__ line_out_offsets,
                               size_t line_cap) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_raw || k >= raw_limit || !flags[k]) return;
    size_t lo = line_start[k];
    size_t hi = k < n_nl ? line_start[k + 1] - 1 : in_len;
    encode_line<true>(in + lo, hi - lo, out + offs[k]);
    if (line_out_offsets && ranks[k] < line_cap) line_out_offsets[ranks[k]] = offs[k];
}

struct GenericTotals {
    unsigned long long n_nl, total_out, total_lines, err_raw;
};

int encode_generic(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, uint8_t* d_out, size_t out_cap,
                   uint64_t* d_line_out_offsets, size_t line_cap, vcfc_result* d_result, cudaStream_t stream) {
    ctx->last_path = kPathGeneric;
    vcfc_result res;
    memset(&res, 0, sizeof(res));
    if (in_len == 0) {
        VCFC_CUDA(ctx, cudaMemcpyAsync(d_result, &res, sizeof(res), cudaMemcpyHostToDevice, stream));
        VCFC_CUDA(ctx, cudaStreamSynchronize(stream));
        return VCFC_OK;
    }
    int rc;
    // 1. raw line table
    size_t n_chunks = (in_len + kNlChunk - 1) / kNlChunk;
    DevBuf &b_counts = ctx->ws[0], &b_scr = ctx->ws[1], &b_tot = ctx->ws[2], &b_ls = ctx->ws[3];
    DevBuf &b_sizes = ctx->ws[4], &b_flags = ctx->ws[5], &b_codes = ctx->ws[6], &b_offs = ctx->ws[7], &b_ranks = ctx->ws[8];
    if ((rc = dev_reserve(ctx, &b_counts, n_chunks * 8))) return rc;
    if ((rc = dev_reserve(ctx, &b_tot, 64))) return rc;
    uint64_t* d_tot = (uint64_t*)b_tot.p;   // [0]=n_nl [1]=total_out [2]=total_lines [3]=err_raw
    k_count_nl<<<(unsigned)n_chunks, kNlThreads, 0, stream>>>(d_in, in_len, (uint64_t*)b_counts.p);
    ctx->launches++;
    if ((rc = scan_exclusive_u64(ctx, (uint64_t*)b_counts.p, (uint64_t*)b_counts.p, n_chunks, d_tot, &b_scr, stream))) return rc;
    uint64_t n_nl = 0;
    uint8_t last = 0;
    VCFC_CUDA(ctx, cudaMemcpyAsync(&n_nl, d_tot, 8, cudaMemcpyDeviceToHost, stream));
    VCFC_CUDA(ctx, cudaMemcpyAsync(&last, d_in + in_len - 1, 1, cudaMemcpyDeviceToHost, stream));
    VCFC_CUDA(ctx, cudaStreamSynchronize(stream));
    size_t n_raw = (size_t)n_nl + (last != '\n' ? 1 : 0);
    if ((rc = dev_reserve(ctx, &b_ls, (n_nl + 2) * 8))) return rc;
    k_fill_lines<<<(unsigned)n_chunks, kNlThreads, 0, stream>>>(d_in, in_len, (uint64_t*)b_counts.p, (uint64_t*)b_ls.p);
    ctx->launches++;
    // 2. sizes
    if ((rc = dev_reserve(ctx, &b_sizes, (n_raw + 1) * 8))) return rc;
    if ((rc = dev_reserve(ctx, &b_flags, (n_raw + 1) * 8))) return rc;
    if ((rc = dev_reserve(ctx, &b_offs, (n_raw + 1) * 8))) return rc;
    if ((rc = dev_reserve(ctx, &b_ranks, (n_raw + 1) * 8))) return rc;
    if ((rc = dev_reserve(ctx, &b_codes, n_raw + 1))) return rc;
    unsigned long long no_err = ~0ull;
    VCFC_CUDA(ctx, cudaMemcpyAsync(d_tot + 3, &no_err, 8, cudaMemcpyHostToDevice, stream));
    unsigned nb = (unsigned)((n_raw + 127) / 128);
    if (n_raw) {
        k_encode_size<<<nb, 128, 0, stream>>>(d_in, in_len, (uint64_t*)b_ls.p, n_raw, (size_t)n_nl, (uint64_t*)b_sizes.p,
                                              (uint64_t*)b_flags.p, (uint8_t*)b_codes.p, (unsigned long long*)(d_tot + 3));
        ctx->launches++;
    }
    if ((rc = scan_exclusive_u64(ctx, (uint64_t*)b_sizes.p, (uint64_t*)b_offs.p, n_raw, d_tot + 1, &b_scr, stream))) return rc;
    if ((rc = scan_exclusive_u64(ctx, (uint64_t*)b_flags.p, (uint64_t*)b_ranks.p, n_raw, d_tot + 2, &b_scr, stream))) return rc;
    GenericTotals tot;
    VCFC_CUDA(ctx, cudaMemcpyAsync(&tot, d_tot, sizeof(tot), cudaMemcpyDeviceToHost, stream));
    VCFC_CUDA(ctx, cudaStreamSynchronize(stream));
    size_t raw_limit = n_raw;
    res.out_len = tot.total_out;
    res.n_lines = tot.total_lines;
    if (tot.err_raw != ~0ull) {   // the reference stops at the first bad line; earlier lines stand
        uint8_t code = 0;
        uint64_t off = 0, rank = 0;
        VCFC_CUDA(ctx, cudaMemcpy(&code, (uint8_t*)b_codes.p + tot.err_raw, 1, cudaMemcpyDeviceToHost));
        VCFC_CUDA(ctx, cudaMemcpy(&off, (uint64_t*)b_offs.p + tot.err_raw, 8, cudaMemcpyDeviceToHost));
        VCFC_CUDA(ctx, cudaMemcpy(&rank, (uint64_t*)b_ranks.p + tot.err_raw, 8, cudaMemcpyDeviceToHost));
        res.status = code;
        res.err_line = rank;
        res.n_lines = rank;
        res.out_len = off;
        raw_limit = (size_t)tot.err_raw;
    }
    if (res.out_len > out_cap) {
        res.status = VCFC_E_CAP;
        res.n_lines = 0;
        res.err_line = 0;
        raw_limit = 0;
    }
    if (raw_limit) {
        k_encode_write<<<nb, 128, 0, stream>>>(d_in, in_len, (uint64_t*)b_ls.p, n_raw, (size_t)n_nl, (uint64_t*)b_offs.p,
                                               (uint64_t*)b_ranks.p, (uint64_t*)b_flags.p, raw_limit, d_out,
                                               d_line_out_offsets, line_cap);
        ctx->launches++;
    }
    if (res.status == VCFC_E_CAP) res.out_len = tot.total_out;   // tell the caller what it needs
    VCFC_CUDA(ctx, cudaMemcpyAsync(d_result, &res, sizeof(res), cudaMemcpyHostToDevice, stream));
    VCFC_CUDA(ctx, cudaStreamSynchronize(stream));
    VCFC_CUDA(ctx, cudaGetLastError());
    return VCFC_OK;
}

// ------------------------------------------------------------------------------------------
// per-line decoder
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ long long get_len_header(const uint8_t* s) {
    if ((s[0] >> 6) != 3) return -1;   // utils.hpp:201-206
    return ((long long)(s[0] & 0x3F) << 24) | ((long long)s[1] << 16) | ((long long)s[2] << 8) | s[3];
}

// Walks the line-length headers (what every reference consumer does: compress.cpp:270-330,
// main.cpp:3805-3922).  Single thread: the generic path is the fallback, not the fast path.
// info[0] = number of lines, info[1] = status, info[2] = bytes covered by whole lines.
__global__ void k_walk_chain(const uint8_t* __restrict__ in, size_t n, uint64_t* __restrict__ line_start, size_t cap,
                             unsigned long long* __restrict__ info) {
    size_t p = 0, k = 0;
    int status = kOk;
    while (n - p >= 8) {                                  // compress.cpp:770-777: <8 bytes left = EOF
        long long ll = get_len_header(in + p), rq = get_len_header(in + p + 4);
        if (ll < 0 || rq < 0) { status = kEFormat; break; }
        if ((unsigned long long)ll + 4 > n - p) { status = kETrunc; break; }
        if (ll < 5 || rq + 5 > ll) { status = kEFormat; break; }
        if (line_start && k < cap) line_start[k] = p;
        k++;
        p += 4 + (size_t)ll;
    }
    if (line_start && k < cap) line_start[k] = p;
    info[0] = k;
    info[1] = (unsigned long long)status;
    info[2] = p;
}

// One compressed line [in, in+avail) -> text.  avail = 4 + line_length.  Returns bytes produced
// or -code.  Follows compress.cpp:741-986 token by token.
template <bool WRITE>
__device__ long long decode_line(const uint8_t* __restrict__ in, size_t avail, uint64_t sample_count,
                                 uint8_t* __restrict__ out) {
    long long rq = get_len_header(in + 4);
    size_t p = 8, o = 0;
    if (rq <= 0) return -(long long)kEFormat;              // fread of 0 bytes throws (compress.cpp:792)
    if ((size_t)rq > avail - p) return -(long long)kETrunc;
    unsigned tabs = 0;
    for (size_t k = 0; k < (size_t)rq; k++) {
        uint8_t c = in[p + k];
        tabs += (c == '\t');
        if (WRITE) out[o + k] = c;
    }
    o += (size_t)rq;
    p += (size_t)rq;
    if (tabs != 9 && !(tabs == 8 && sample_count == 0)) return -(long long)kEFormat;   // compress.cpp:820-828
    uint64_t ns = 0;
    while (ns < sample_count) {
        if (p >= avail) return -(long long)kETrunc;
        uint32_t b = in[p++];
        if ((b & 0x80) == 0) {                              // 0|0 run, compress.cpp:843-868
            uint32_t cnt = b & 0x7F;
            if (WRITE)
                for (uint32_t k = 0; k < cnt; k++) {
                    out[o + 4 * k] = '0'; out[o + 4 * k + 1] = '|'; out[o + 4 * k + 2] = '0'; out[o + 4 * k + 3] = '\t';
                }
            o += 4 * (size_t)cnt;
            ns += cnt;
            if (ns >= sample_count) {
                if (o == 0) return -(long long)kEFormat;
                o--;
            }
        } else if ((b & 0xE0) == 0xE0) {                    // literal, compress.cpp:869-906
            uint32_t ncols = b & 0x1F, u = 0;
            while (u < ncols) {
                if (p >= avail) return -(long long)kETrunc;
                uint8_t c = in[p++];
                if (c == '\n') {
                    u++; ns++;
                    if (u != ncols) return -(long long)kEFormat;
                    p--;
                } else if (c == '\t') {
                    u++; ns++;
                    if (ns < sample_count) { if (WRITE) out[o] = c; o++; }
                } else {
                    if (WRITE) out[o] = c;
                    o++;
                }
            }
        } else {                                            // 0|1, 1|0, 1|1, compress.cpp:907-953
            uint32_t f = b & 0xE0, cnt = b & 0x1F;
            uint8_t a = f == kTok01 ? '0' : '1', c2 = f == kTok10 ? '0' : '1';
            for (uint32_t k = 0; k < cnt; k++) {
                if (WRITE) { out[o] = a; out[o + 1] = '|'; out[o + 2] = c2; }
                o += 3;
                ns++;
                if (ns < sample_count) { if (WRITE) out[o] = '\t'; o++; }
            }
        }
    }
    if (p >= avail) return -(long long)kETrunc;             // compress.cpp:958-960
    if (in[p++] != '\n') return -(long long)kEFormat;       // compress.cpp:961-966
    if (p != avail) return -(long long)kEFormat;            // line_length header disagrees with the tokens
    if (WRITE) out[o] = '\n';
    o++;
    return (long long)o;
}

__global__ void k_decode_size(const uint8_t* __restrict__ in, const uint64_t* __restrict__ line_start, size_t n_lines,
                              uint64_t sample_count, uint64_t* __restrict__ sizes, uint8_t* __restrict__ codes,
                              unsigned long long* __restrict__ err_line) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines) return;
    size_t lo = line_start[k], hi = line_start[k + 1];
    long long r = decode_line<false>(in + lo, hi - lo, sample_count, nullptr);
    uint8_t code = 0;
    if (r < 0) {
        code = (uint8_t)(-r);
        atomicMin(err_line, (unsigned long long)k);
        r = 0;
    }
    sizes[k] = (uint64_t)r;
    codes[k] = code;
}

__global__ void k_decode_write(const uint8_t* __restrict__ in, const uint64_t* __restrict__ line_start, size_t n_lines,
                               uint64_t sample_count, const uint64_t* __restrict__ offs, uint8_t* __restrict__ out) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines) return;
    size_t lo = line_start[k], hi = line_start[k + 1];
    decode_line<true>(in + lo, hi - lo, sample_count, out + offs[k]);
}

int decode_generic(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, uint64_t sample_count, uint8_t* d_out,
                   size_t out_cap, vcfc_result* d_result, bool size_only, cudaStream_t stream) {
    ctx->last_path = kPathGeneric;
    vcfc_result res;
    memset(&res, 0, sizeof(res));
    int rc;
    DevBuf &b_scr = ctx->ws[1], &b_tot = ctx->ws[2], &b_ls = ctx->ws[3], &b_sizes = ctx->ws[4], &b_codes = ctx->ws[6],
           &b_offs = ctx->ws[7];
    if ((rc = dev_reserve(ctx, &b_tot, 64))) return rc;
    unsigned long long* d_tot = (unsigned long long*)b_tot.p;
    unsigned long long info[3] = {0, 0, 0};
    if (in_len >= 8) {
        k_walk_chain<<<1, 1, 0, stream>>>(d_in, in_len, nullptr, 0, d_tot);
        ctx->launches++;
        VCFC_CUDA(ctx, cudaMemcpyAsync(info, d_tot, sizeof(info), cudaMemcpyDeviceToHost, stream));
        VCFC_CUDA(ctx, cudaStreamSynchronize(stream));
    }
    size_t n_lines = (size_t)info[0];
    int chain_status = (int)info[1];
    if ((rc = dev_reserve(ctx, &b_ls, (n_lines + 2) * 8))) return rc;
    if ((rc = dev_reserve(ctx, &b_sizes, (n_lines + 1) * 8))) return rc;
    if ((rc = dev_reserve(ctx, &b_offs, (n_lines + 1) * 8))) return rc;
    if ((rc = dev_reserve(ctx, &b_codes, n_lines + 1))) return rc;
    unsigned long long err_raw = ~0ull, total = 0;
    if (n_lines) {
        k_walk_chain<<<1, 1, 0, stream>>>(d_in, in_len, (uint64_t*)b_ls.p, n_lines + 1, d_tot);
        VCFC_CUDA(ctx, cudaMemcpyAsync(d_tot + 3, &err_raw, 8, cudaMemcpyHostToDevice, stream));
        unsigned nb = (unsigned)((n_lines + 127) / 128);
        k_decode_size<<<nb, 128, 0, stream>>>(d_in, (uint64_t*)b_ls.p, n_lines, sample_count, (uint64_t*)b_sizes.p,
                                              (uint8_t*)b_codes.p, d_tot + 3);
        ctx->launches += 2;
        if ((rc = scan_exclusive_u64(ctx, (uint64_t*)b_sizes.p, (uint64_t*)b_offs.p, n_lines, (uint64_t*)(d_tot + 4), &b_scr, stream)))
            return rc;
        VCFC_CUDA(ctx, cudaMemcpyAsync(&err_raw, d_tot + 3, 8, cudaMemcpyDeviceToHost, stream));
        VCFC_CUDA(ctx, cudaMemcpyAsync(&total, d_tot + 4, 8, cudaMemcpyDeviceToHost, stream));
        VCFC_CUDA(ctx, cudaStreamSynchronize(stream));
    }
    size_t n_write = n_lines;
    res.out_len = total;
    res.n_lines = n_lines;
    if (err_raw != ~0ull) {
        uint8_t code = 0;
        uint64_t off = 0;
        VCFC_CUDA(ctx, cudaMemcpy(&code, (uint8_t*)b_codes.p + err_raw, 1, cudaMemcpyDeviceToHost));
        VCFC_CUDA(ctx, cudaMemcpy(&off, (uint64_t*)b_offs.p + err_raw, 8, cudaMemcpyDeviceToHost));
        res.status = code;
        res.err_line = err_raw;
        res.n_lines = err_raw;
        res.out_len = off;
        n_write = (size_t)err_raw;
    } else if (chain_status != kOk) {
        res.status = chain_status;
        res.err_line = n_lines;
    }
    if (!size_only) {
        if (res.out_len > out_cap) {
            res.status = VCFC_E_CAP;
            res.n_lines = 0;
            res.err_line = 0;
            res.out_len = total;
        } else if (n_write) {
            unsigned nb = (unsigned)((n_write + 127) / 128);
            k_decode_write<<<nb, 128, 0, stream>>>(d_in, (uint64_t*)b_ls.p, n_write, sample_count, (uint64_t*)b_offs.p, d_out);
            ctx->launches++;
        }
    }
    VCFC_CUDA(ctx, cudaMemcpyAsync(d_result, &res, sizeof(res), cudaMemcpyHostToDevice, stream));
    VCFC_CUDA(ctx, cudaStreamSynchronize(stream));
    VCFC_CUDA(ctx, cudaGetLastError());
    return VCFC_OK;
}

}  // namespace vcfc
