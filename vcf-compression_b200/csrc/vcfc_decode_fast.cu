// vcfc_decode_fast.cu -- tile-parallel decoder (sm_100a) for well-formed .vcfc data lines:
// every line's token stream holds exactly sample_count samples, literals are single-column
// (0xE1, what the reference encoder writes, compress.cpp:179-181), >= 1 sample per line.
// Anything else sets ctrl->irregular and the caller reruns the block on the generic kernels,
// which reproduce the reference decoder's behaviour on odd input token by token.
//
// Output bytes are those of decompress2_data_line (/root/reference/src/compress.cpp:741-986).
// Pipeline (all on the launch stream; two small host reads size the workspace):
//   D1 k_dec_walk     line table without a serial pointer chase: one warp per 2 KB segment finds the
//                     segment's first plausible line start and follows the 4-byte length headers
//                     (utils.hpp:134-247) to the segment end; k_dec_verify accepts the table only if every
//                     segment's chain ends exactly where the next one began (induction from offset 0).
//   D2 k_dec_sizes    one warp per line: validates the line, sizes its text and writes the chunk table.  The
//                     token/payload state of a byte is the kind of the last "setter" before it (a byte >= 0xE0
//                     enters a payload, a tab/newline leaves it), so 16-byte chunks parse in parallel.
//   D3 scan           line text offsets; k_dec_tilemap: first line (and chunk) of every output tile
//   D4 k_dec_expand_grid  one CTA per 32 KB OUTPUT tile, sample text on the 4-byte grid (the normal case): the
//                     tile image is filled with the default genotype in each line's phase, then one thread per
//                     16 token bytes patches in what differs; aligned 16-byte stores.  HBM traffic = 2 C + N.
//      k_dec_expand   the same for text off the grid: every thread walks the tokens of one 64-byte span.
#include <algorithm>

#include "vcfc_common.cuh"
#include "vcfc_internal.h"

namespace vcfc {
namespace dec {

constexpr int kSeg = 2048;              // D1 segment (compressed bytes): one thread each, so keep the serial walk short
constexpr int kTile = 16384;            // D4 output tile
constexpr int kCmax = 24576;            // longest compressed line the tile kernel parses from smem
constexpr long long kNoCand = -1, kBroken = -2;

constexpr int kMaxFix = 1024;
struct Ctrl {
    int irregular;
    int n_fix;                                    // segments whose first plausible line start is not on the chain
    unsigned long long n_lines, end_pos, total_out;
    int not_grid, pad;                            // some sample text is not 4 bytes wide (or a required section < 16 bytes)
    int fix[kMaxFix];
};

__device__ __forceinline__ long long hdr_len(const uint8_t* p) {     // utils.hpp:188-231; -1 = bad tag
    if ((p[0] >> 6) != 3) return -1;
    return ((long long)(p[0] & 0x3F) << 24) | ((long long)p[1] << 16) | ((long long)p[2] << 8) | p[3];
}

// a position where a line can start: both tags present, lengths consistent, line ends with '\n'
__device__ __forceinline__ bool plausible_line(const uint8_t* __restrict__ in, long long n, long long p, long long* ll_out) {
    if (n - p < 8) return false;
    long long ll = hdr_len(in + p), rq = hdr_len(in + p + 4);
    if (ll < 0 || rq < 1 || rq + 5 > ll || ll + 4 > n - p) return false;
    if (in[p + 4 + ll - 1] != '\n') return false;
    if (in[p + 8 + rq - 1] != '\t') return false;            // the required section of a line with samples ends with a tab
    *ll_out = ll;
    return true;
}

// ---- D1: speculative line table ---------------------------------------------------------------------
__global__ void k_dec_walk(const uint8_t* __restrict__ in, long long n, long long n_seg, long long* __restrict__ cand,
                           long long* __restrict__ endp, unsigned long long* __restrict__ cnt) {
    // one warp per segment: the lanes look for the first plausible line start 128 bytes at a time, then the warp
    // follows the length headers (uniform loads) to the end of the segment
    const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n_seg) return;
    const long long lo = s * kSeg, hi = lo + kSeg < n ? lo + kSeg : n;
    long long c = kNoCand, ll = 0;
    if ((reinterpret_cast<uintptr_t>(in) & 15) == 0) {
        // 512 bytes per round trip: newline bytes by word-parallel tests, then the candidates behind them in order
        if (lo == 0) {
            if (plausible_line(in, n, 0, &ll)) c = 0;            // offset 0 must be a line start
        } else {
            for (long long wb = lo - 16; wb < hi && c == kNoCand; wb += 512) {
                const long long g = wb + 16 * lane;
                unsigned nlm = 0;                                // bit j: byte g + j is a newline
                if (g + 16 <= n) {
                    const uint4 v = *reinterpret_cast<const uint4*>(in + g);
                    nlm = nibble_of(zero_bytes(v.x ^ 0x0A0A0A0Au)) | (nibble_of(zero_bytes(v.y ^ 0x0A0A0A0Au)) << 4) |
                          (nibble_of(zero_bytes(v.z ^ 0x0A0A0A0Au)) << 8) | (nibble_of(zero_bytes(v.w ^ 0x0A0A0A0Au)) << 12);
                } else {
                    for (int j = 0; j < 16; j++) if (g + j < n && in[g + j] == '\n') nlm |= 1u << j;
                }
                unsigned cand = 0;                               // bit j: a line may start at g + j + 1
                for (unsigned t = nlm; t; t &= t - 1) {
                    const int j = __ffs(t) - 1;
                    const long long p = g + j + 1;
                    if (p >= lo && p < hi && n - p >= 8 && (in[p] >> 6) == 3 && (in[p + 4] >> 6) == 3) cand |= 1u << j;   // both length tags
                }
                unsigned any;
                while (c == kNoCand && (any = __ballot_sync(0xffffffffu, cand != 0u)) != 0u) {
                    const int leader = __ffs(any) - 1;
                    const long long p = __shfl_sync(0xffffffffu, g + __ffs(cand), leader);
                    if (lane == leader) cand &= cand - 1;
                    if (plausible_line(in, n, p, &ll)) c = p;
                }
            }
        }
    } else {
        for (long long base = lo; base < hi && c == kNoCand; base += 128) {
            unsigned hit[4];
    #pragma unroll
            for (int u = 0; u < 4; u++) {
                const long long p = base + 32 * u + lane;
                bool h = false;
                if (p < hi && n - p >= 8) {
                    const bool after_nl = p == 0 || in[p - 1] == '\n';
                    h = after_nl && (in[p] >> 6) == 3 && (in[p + 4] >> 6) == 3;      // both length tags present
                }
                hit[u] = __ballot_sync(0xffffffffu, h);
            }
    #pragma unroll
            for (int u = 0; u < 4; u++) {
                unsigned m = hit[u];
                while (m && c == kNoCand) {
                    const long long p = base + 32 * u + (__ffs(m) - 1);
                    m &= m - 1;
                    if (plausible_line(in, n, p, &ll)) c = p;
                }
            }
            if (lo == 0 && c != 0) break;                        // offset 0 must be a line start
        }
        if (lo == 0 && c != 0) c = kNoCand;
    }
    unsigned long long k = 0;
    long long p = c;
    if (c >= 0) {
        while (true) {
            k++;
            p += 4 + ll;
            if (p >= hi) break;
            if (n - p < 8) break;                            // <8 bytes left: EOF for the reference (compress.cpp:770-777)
            if (!plausible_line(in, n, p, &ll)) { p = kBroken; break; }
        }
    }
    if (lane == 0) {
        cand[s] = c;
        endp[s] = p;
        cnt[s] = k;
    }
}

// Chain check, by induction from offset 0: a segment's first line start must be where the previous chain
// ended, and a chain that ends inside a segment must find a line start there (so the last chain reaches
// EOF: fewer than 8 bytes left, compress.cpp:770-777).  A '\n'-valued token followed by header-looking
// bytes can fake a line start; such segments are listed (strict = 0) for k_dec_repair, and the table is
// accepted only if a second, strict pass finds nothing.
__global__ void k_dec_verify(long long n, long long n_seg, const long long* __restrict__ cand, const long long* __restrict__ endp,
                             Ctrl* __restrict__ ctrl, int strict) {
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const long long c = cand[s], e = endp[s];
    if (s == 0) { if (c != 0 || e == kBroken) atomicExch(&ctrl->irregular, 1); return; }
    bool ok = true;
    long long q = s - 1;                                      // previous segment that has a line start
    while (q > 0 && cand[q] < 0) q--;
    const long long pe = endp[q];                             // where its chain ends = where the next line starts
    if (c >= 0) {
        if (pe != c || e == kBroken) ok = false;
    } else if (pe >= 0 && pe / kSeg == s && n - pe >= 8) {
        ok = false;                                           // the chain says a line starts here but none was found
    }
    if (!ok) {
        if (strict) atomicExch(&ctrl->irregular, 1);
        else { int slot = atomicAdd(&ctrl->n_fix, 1); if (slot < kMaxFix) ctrl->fix[slot] = (int)s; }
    }
}

// Re-walks the listed segments from where the chain really arrives (one thread, left to right: few entries).
__global__ void k_dec_repair(const uint8_t* __restrict__ in, long long n, long long n_seg, long long* __restrict__ cand,
                             long long* __restrict__ endp, unsigned long long* __restrict__ cnt, Ctrl* __restrict__ ctrl) {
    const int nf = ctrl->n_fix;
    if (nf == 0 || ctrl->irregular) return;
    if (nf > kMaxFix || n_seg >= (1ll << 31)) { ctrl->irregular = 1; return; }
    for (int i = 1; i < nf; i++) {                            // insertion sort
        int v = ctrl->fix[i], j = i - 1;
        while (j >= 0 && ctrl->fix[j] > v) { ctrl->fix[j + 1] = ctrl->fix[j]; j--; }
        ctrl->fix[j + 1] = v;
    }
    for (int i = 0; i < nf; i++) {
        const long long s = ctrl->fix[i];
        long long q = s - 1;
        while (q > 0 && cand[q] < 0) q--;
        long long p = endp[q], ll = 0;
        const long long hi = (s + 1) * kSeg < n ? (s + 1) * kSeg : n;
        unsigned long long k = 0;
        long long c = kNoCand, e = kNoCand;
        if (p >= 0 && p / kSeg == s && n - p >= 8) {          // the chain arrives inside this segment: walk from there
            c = p;
            if (!plausible_line(in, n, p, &ll)) { ctrl->irregular = 1; return; }
            while (true) {
                k++;
                p += 4 + ll;
                if (p >= hi || n - p < 8) break;
                if (!plausible_line(in, n, p, &ll)) { ctrl->irregular = 1; return; }
            }
            e = p;
        } else if (!(p >= 0 && (p / kSeg > s || n - p < 8))) { // else: the chain passes over this segment -> it is empty
            ctrl->irregular = 1;
            return;
        }
        cand[s] = c;
        endp[s] = e;
        cnt[s] = k;
    }
}

__global__ void k_dec_fill(const uint8_t* __restrict__ in, long long n, long long n_seg, const long long* __restrict__ cand,
                           const unsigned long long* __restrict__ base, unsigned long long* __restrict__ line_start,
                           unsigned* __restrict__ rq_arr, unsigned long long n_lines, const Ctrl* __restrict__ ctrl) {
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg || ctrl->irregular) return;
    long long p = cand[s];
    if (p < 0) return;
    const long long hi = (s + 1) * kSeg < n ? (s + 1) * kSeg : n;
    unsigned long long k = base[s];
    while (true) {
        long long ll = hdr_len(in + p);
        rq_arr[k] = (unsigned)hdr_len(in + p + 4);          // (same sector as the line length: saves k_dec_sizes a dependent load)
        line_start[k++] = (unsigned long long)p;
        p += 4 + ll;
        if (p >= hi || n - p < 8) break;
    }
    if (k == n_lines) line_start[k] = (unsigned long long)p;
}

// ---- token stream parsing -----------------------------------------------------------------------------
// kinds of setter: 1 = a byte >= 0xE0 (enters a literal payload when read as a token, stays in it when read as payload),
// 2 = tab / newline (ends a payload; as a token it is a 0|0 run of 9 / 10 and the state stays "token")

// Walks nb token-region bytes starting in `payload` state; accumulates text bytes and samples.
// is_last: the chunk ends with the line's final '\n'.  err bits: 1 = malformed for this path.
// *tres: bit r set = a literal ended (its terminator included) at chunk-relative text offset == r (mod 4).
__device__ __forceinline__ void chunk_measure(const uint8_t* b, int nb, bool payload, bool is_last, unsigned* out_len,
                                              unsigned* samples, int* err, unsigned* tres) {
    unsigned o = 0, ns = 0, tr = 0;
    int e = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t c = b[i];
        const bool line_end = is_last && i == nb - 1;
        if (i >= nb) {
        } else if (payload) {
            o++;
            if (c == 9u) { payload = false; ns++; tr |= 1u << (o & 3u); }
            else if (c == 10u) { payload = false; ns++; tr |= 1u << (o & 3u); if (!line_end) e = 1; }     // '\n' inside a payload: compress.cpp:875-884
        } else if (line_end) {
            // the newline after a run: it replaces the run's last tab, no text of its own
        } else if (c < 0x80u) {
            if (c == 0) e = 1;
            o += 4u * c; ns += c;
        } else if (c < 0xE0u) {
            const uint32_t n = c & 0x1Fu;
            if (n == 0) e = 1;
            o += 4u * n; ns += n;
        } else {
            if ((c & 0x1Fu) != 1u) e = 1;                    // multi-column literal: generic path
            payload = true;
        }
    }
    *out_len = o; *samples = ns; *err = e; *tres = tr;
}

// ---- D2: per-line validation and text size.  kSzGroup lanes per line (four lines per warp): most token regions of a sparse
//      file are shorter than 128 bytes, so a whole warp per line would leave most lanes idle -------------------------------------
#ifndef VCFC_DEC_SZGROUP
#define VCFC_DEC_SZGROUP 16
#endif
constexpr int kSzGroup = VCFC_DEC_SZGROUP;                       // 8, 16 or 32 lanes per line
constexpr int kSzLines = 32 / kSzGroup;
__global__ void k_dec_sizes(const uint8_t* __restrict__ in, const unsigned long long* __restrict__ line_start,
                            unsigned long long n_lines, unsigned long long sample_count, unsigned long long* __restrict__ sizes,
                            unsigned* __restrict__ ctab, unsigned* __restrict__ rq_arr, uint8_t* __restrict__ lflag, Ctrl* __restrict__ ctrl) {
    const unsigned long long w = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, gl = lane & (kSzGroup - 1), g0 = lane & ~(kSzGroup - 1);
    const unsigned gmask = (kSzGroup == 32 ? 0xFFFFFFFFu : ((1u << kSzGroup) - 1u)) << g0;
    const unsigned long long k = w * kSzLines + (unsigned long long)(lane / kSzGroup);
    const bool live = k < n_lines;
    const unsigned long long ls = live ? line_start[k] : 0ull, le = live ? line_start[k + 1] : 0ull;
    const uint8_t* p = in + ls;
    const long long clen = (long long)(le - ls);
    const long long rq = live ? (long long)(int)rq_arr[k] : 0ll;   // second length header, read by k_dec_fill (-1: bad tag)
    bool bad = rq < 1 || rq + 9 > clen || clen > 0x7fffffffll;
    if (live && gl == 0) rq_arr[k] = (unsigned)(bad ? 0 : rq);
    // required section: exactly 9 tabs (compress.cpp:820-828; the 8-tab form means no samples -> generic path)
    unsigned tabs = 0;
    {
        const long long rqv = (live && !bad) ? rq : 0ll;
        for (long long o16 = 16ll * gl; __any_sync(0xffffffffu, o16 < rqv); o16 += 16 * kSzGroup) {   // sixteen bytes per lane
            if (o16 < rqv) {
                const uintptr_t ga = reinterpret_cast<uintptr_t>(p + 8 + o16);
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(ga & ~uintptr_t(3));
                const int nv = (int)(rqv - o16 < 16 ? rqv - o16 : 16), sh = 8 * (int)(ga & 3);
                const int nwd = (nv + (int)(ga & 3) + 3) >> 2;
                const uint32_t w0 = wp[0], w1 = nwd > 1 ? wp[1] : 0u, w2 = nwd > 2 ? wp[2] : 0u, w3 = nwd > 3 ? wp[3] : 0u,
                               w4 = nwd > 4 ? wp[4] : 0u;
                const uint32_t v[4] = {__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                                       __funnelshift_r(w3, w4, sh)};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int nj = nv - 4 * j;
                    const uint32_t bm = nj >= 4 ? 0xFFFFFFFFu : (nj <= 0 ? 0u : ((1u << (8 * nj)) - 1u));
                    tabs += __popc(zero_bytes((v[j] ^ 0x09090909u) | ~bm));
                }
            }
        }
    }
#pragma unroll
    for (int d = kSzGroup / 2; d; d >>= 1) tabs += __shfl_xor_sync(0xffffffffu, tabs, d);
    if (tabs != 9) bad = true;
    unsigned long long total = 0, ns_total = 0;
    const long long tb = 8 + rq, tn = (live && !bad) ? clen - tb : 0ll;   // token region, its last byte is the line's '\n'
    int carry_kind = 0;                                         // last setter seen so far (0 = none: token state)
    int err_any = 0, off_grid = (live && !bad && rq < 16) ? 1 : 0;
    for (long long base = 0; __any_sync(0xffffffffu, base < tn); base += 16 * kSzGroup) {
        const long long off = base + 16ll * gl;
        int nb = (int)(tn - off < 0 ? 0 : (tn - off > 16 ? 16 : tn - off));
        if (base >= tn) nb = 0;
        uint32_t v[4] = {0u, 0u, 0u, 0u};
        uint32_t lit_any = 0, set_any = 0, set_lit = 0, run_sum = 0, run_zero = 0;
        const bool is_last = nb > 0 && off + nb == tn;
        const bool full = nb == 16 && !is_last;                      // a full chunk inside the line (most are): no byte masks
        const int nbv = nb - (is_last ? 1 : 0);                      // the line's final '\n' carries no text
        if (nb > 0) {       // five aligned 32-bit loads instead of sixteen byte loads; bytes past the line are never used
            const uintptr_t ga = reinterpret_cast<uintptr_t>(p + tb + off);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(ga & ~uintptr_t(3));
            const int sh = 8 * (int)(ga & 3);
            const int nwd = (nb + (int)(ga & 3) + 3) >> 2;          // aligned words that hold the chunk's nb bytes
            uint32_t w0 = wp[0], w1 = nwd > 1 ? wp[1] : 0u, w2 = nwd > 2 ? wp[2] : 0u, w3 = nwd > 3 ? wp[3] : 0u,
                     w4 = nwd > 4 ? wp[4] : 0u;
            v[0] = __funnelshift_r(w0, w1, sh); v[1] = __funnelshift_r(w1, w2, sh); v[2] = __funnelshift_r(w2, w3, sh);
            v[3] = __funnelshift_r(w3, w4, sh);
            // word-parallel view of the chunk: is there any literal marker (>= 0xE0) among its bytes, and what the bytes add up
            // to when every one of them is a run token
            if (full) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t wv = v[j];
                    lit_any |= wv & (wv << 1) & (wv << 2) & 0x80808080u;
                    const uint32_t cv = wv & (0x7F7F7F7Fu ^ (((wv >> 7) & 0x01010101u) * 0x60u));
                    run_sum = __dp4a(cv, 0x01010101u, run_sum);
                    run_zero |= zero_bytes(cv);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int nj = nb - 4 * j, njv = nbv - 4 * j;
                    const uint32_t bm = nj >= 4 ? 0xFFFFFFFFu : (nj <= 0 ? 0u : ((1u << (8 * nj)) - 1u));
                    const uint32_t bmv = njv >= 4 ? 0xFFFFFFFFu : (njv <= 0 ? 0u : ((1u << (8 * njv)) - 1u));
                    const uint32_t wv = v[j];
                    lit_any |= wv & (wv << 1) & (wv << 2) & 0x80808080u & bm;                                // bytes >= 0xE0
                    const uint32_t cv = wv & (0x7F7F7F7Fu ^ (((wv >> 7) & 0x01010101u) * 0x60u)) & bmv;     // run lengths
                    run_sum = __dp4a(cv, 0x01010101u, run_sum);
                    run_zero |= zero_bytes(cv | ~bmv);
                }
            }
        }
        // The chunk's state (token / literal payload) at its first byte.  While no literal marker has been seen in the line -- in
        // this round or before it (carry_kind) -- every chunk starts in token state and the tabs / newlines among its bytes are
        // run tokens (counts 9 / 10): nothing to look for.  (A sparse file has a literal in about one line in a hundred.)
        int kind = 0, k_in = 0;
        unsigned has = 0;
        if (__any_sync(0xffffffffu, lit_any != 0u || carry_kind == 1)) {
            if (nb > 0) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int nj = nb - 4 * j;
                    const uint32_t bm = nj >= 4 ? 0xFFFFFFFFu : (nj <= 0 ? 0u : ((1u << (8 * nj)) - 1u));
                    const uint32_t wv = v[j];
                    const uint32_t lm = wv & (wv << 1) & (wv << 2) & 0x80808080u & bm;
                    const uint32_t am = lm | ((zero_bytes(wv ^ 0x09090909u) | zero_bytes(wv ^ 0x0A0A0A0Au)) & bm);   // ... or tab / newline
                    if (am) { set_any = am; set_lit = lm; }                                             // the highest word that has a setter
                }
            }
            // kind of the chunk's last setter: 1 = a byte >= 0xE0, 2 = tab / newline (kinds of setter above)
            kind = set_any ? (((set_lit >> (31 - __clz(set_any))) & 1u) ? 1 : 2) : 0;
            // state at chunk start: last setter of the nearest lower lane of the group that has one, else the carry
            has = __ballot_sync(0xffffffffu, kind != 0) & gmask;
            const unsigned below = has & ((1u << lane) - 1u);
            const int src = below ? 31 - __clz(below) : lane;
            const int k_src = __shfl_sync(0xffffffffu, kind, src);
            k_in = below ? k_src : carry_kind;
        }
        unsigned o = 0, ns = 0, tres = 0;
        int e = 0;
        if (nb > 0) {
            if (k_in != 1 && !lit_any) { o = 4u * run_sum; ns = run_sum; e = run_zero ? 1 : 0; }   // run tokens only
            else {
                uint8_t b[16];
#pragma unroll
                for (int i = 0; i < 16; i++) b[i] = (uint8_t)(v[i >> 2] >> (8 * (i & 3)));
                chunk_measure(b, nb, k_in == 1, is_last, &o, &ns, &e, &tres);
            }
        }
        err_any |= e;
        // chunk table (text offset << 1 | payload state at the chunk's first byte), consumed by k_dec_expand;
        // line k owns the slots [(ls >> 4) + k, ...): disjoint between lines, <= clen / 16 + 1 of them
        unsigned inc = o, n32 = ns;
#pragma unroll
        for (int d = 1; d < kSzGroup; d <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, inc, d, kSzGroup), u = __shfl_up_sync(0xffffffffu, n32, d, kSzGroup);
            if (gl >= d) { inc += t; n32 += u; }
        }
        if (nb > 0) ctab[(ls >> 4) + k + (unsigned long long)(off >> 4)] = (((unsigned)total + inc - o) << 1) | (k_in == 1 ? 1u : 0u);
        // every literal must end on the 4-byte sample grid for the fill-and-patch kernel (k_dec_expand_grid)
        if (tres & ~(1u << ((4u - ((unsigned)total + inc - o)) & 3u))) off_grid = 1;
        total += __shfl_sync(0xffffffffu, inc, g0 + kSzGroup - 1);
        ns_total += __shfl_sync(0xffffffffu, n32, g0 + kSzGroup - 1);
        if (__any_sync(0xffffffffu, has != 0u)) {
            const int k_last = __shfl_sync(0xffffffffu, kind, has ? 31 - __clz(has) : lane);
            if (has) carry_kind = k_last;
        }
    }
    const unsigned eg = __ballot_sync(0xffffffffu, err_any != 0) & gmask;
    if (live && !bad) {
        if (eg || ns_total != sample_count) bad = true;
        total += (unsigned long long)rq;
    }
    const unsigned og = __ballot_sync(0xffffffffu, off_grid != 0) & gmask;
    if (live && gl == 0) {
        lflag[k] = (og && !bad) ? 1 : 0;           // the line is off the 4-byte grid: its tiles go to the span-walking kernel
        if (og && !bad) atomicExch(&ctrl->not_grid, 1);
        sizes[k] = bad ? 0ull : total;
        if (bad) atomicExch(&ctrl->irregular, 1);
    }
}

// First line of every output tile (line k covers text bytes [off[k], off[k+1])) and, when the tile starts inside
// that line's sample text, the chunk that holds the tile's first byte: a long line is then staged from that chunk
// on instead of from its start (kWholeLine otherwise).
constexpr unsigned kWholeLine = 0xFFFFFFFFu;
constexpr int kReqDirect = 2048;          // required sections longer than this are copied from global memory, not staged
__global__ void k_dec_tilemap(const unsigned long long* __restrict__ off, unsigned long long n_lines, unsigned long long total,
                              const unsigned long long* __restrict__ line_start, const unsigned* __restrict__ rq_arr,
                              const unsigned* __restrict__ gtab, unsigned int* __restrict__ first_line,
                              unsigned int* __restrict__ first_chunk, const unsigned long long tile,
                              const uint8_t* __restrict__ lflag, uint8_t* __restrict__ gflag, const unsigned long long gtile) {
    const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines) return;
    const unsigned long long a = off[k], b = k + 1 < n_lines ? off[k + 1] : total;
    if (b == a) return;
    // a block that holds lines off the 4-byte sample grid: the fill-and-patch tiles (gtile bytes) such a line overlaps are
    // left to the span-walking kernel, all others stay with k_dec_expand_grid
    if (gflag && lflag[k]) for (unsigned long long t = a / gtile; t * gtile < b; t++) gflag[t] = 1;
    const unsigned long long ls = line_start[k];
    const long long rq = rq_arr[k];
    const int nch = (int)(((long long)(line_start[k + 1] - ls) - 8 - rq + 15) >> 4);
    const unsigned* tab = gtab + (ls >> 4) + k;
    for (unsigned long long t = (a + tile - 1) / tile; t * tile < b; t++) {
        first_line[t] = (unsigned int)k;
        const long long xs = (long long)(t * tile - a) - rq;          // text offset of the tile start inside the sample text
        unsigned fc = kWholeLine;
        if (xs < 0 && 8 + rq > kReqDirect) fc = 0u;                   // a long required section is not staged: tokens from chunk 0
        if (xs >= 0) {
            int lo = 0, hi = nch;
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if ((long long)(tab[mid] >> 1) <= xs) lo = mid; else hi = mid; }
            fc = (unsigned)lo;
        }
        first_chunk[t] = fc;
    }
}

// ---- D4: expansion, one CTA per output tile --------------------------------------------------------------------
// Every thread produces one contiguous 64-byte span of the tile image, so the work is balanced whatever the
// token mix.  Per batch of lines that overlap the tile: (1) warp 0 reads the line table entries, (2) the lines'
// compressed bytes (contiguous in the block) are staged in smem with 16-byte copies, (3) one warp per line builds
// a chunk table (text offset + token/payload state at every 16th token byte) with warp scans only, (4) each
// thread binary-searches the table for its span, walks to the token that covers it and generates its bytes.
#ifndef VCFC_DEC_XTHREADS
#define VCFC_DEC_XTHREADS 256
#define VCFC_DEC_XCTAS 4
#endif
constexpr int kXThreads = VCFC_DEC_XTHREADS, kXWarps = kXThreads / 32;   // D4 block size
constexpr int kSpan = kTile / kXThreads;    // 64
constexpr int kChunks = kCmax / 16 + 2 * 32;
constexpr int kMaxL = 31;                   // lines per batch

struct Smem {
    alignas(16) uint8_t stage[kTile];
    alignas(16) uint8_t cbuf[kCmax + 32];
    unsigned ctab[kChunks];                  // (text offset at chunk start) << 1 | payload state
    int l_pos[kMaxL + 1];                    // tile position of the line's first text byte (may be < 0)
    int l_end[kMaxL + 1];                    // tile position one past its last text byte, clipped to the tile
    int l_last[kMaxL + 1];                   // tile position of its final '\n' (may lie outside the tile)
    int l_coff[kMaxL + 2];                   // offset of its compressed bytes in cbuf
    int l_ctab[kMaxL + 1];                   // its first chunk table slot
    int n_batch, more;
    int staged;                              // compressed bytes of the batch held in cbuf
    int c_first, rq0;                        // first line staged from chunk c_first of its token region (-1: from its start)
    unsigned long long c_lo;                 // block offset of cbuf's first staged byte
};

__device__ __forceinline__ uint32_t sample_word(uint32_t tok) {      // token byte -> "x|y\t" little-endian
    uint32_t a = '0', b = '0';
    if (tok & 0x80u) {
        const uint32_t f = tok & 0xE0u;
        a = f == kTok01 ? '0' : '1';
        b = f == kTok10 ? '0' : '1';
    }
    return a | ((uint32_t)'|' << 8) | (b << 16) | ((uint32_t)'\t' << 24);
}

// Sequential writer for one thread's span of the tile image: bytes are collected in a register and stored as
// aligned 32-bit words (a sample is 4 bytes but starts at any byte offset, so byte stores would cost 4x more).
struct Writer {
    uint32_t* wp;          // next word of the image
    uint32_t acc;          // pending bytes, low byte first
    int fill;              // number of pending bytes, 0..3
    __device__ __forceinline__ void byte(uint32_t c) {
        acc |= c << (8 * fill);
        if (++fill == 4) { *wp++ = acc; acc = 0; fill = 0; }
    }
    // n bytes of the periodic stream w[ph], w[ph+1], ... (w = 4-byte sample word)
    __device__ __forceinline__ void run(uint32_t w, int ph, int n) {
        const uint32_t P = __funnelshift_r(w, w, 8 * ph);              // stream bytes 0..3
        const int tot = fill + n, nw = tot >> 2, rem = tot & 3;
        if (nw == 0) { acc |= (P & ((1u << (8 * n)) - 1u)) << (8 * fill); fill = tot; return; }
        *wp++ = acc | (P << (8 * fill));
        const uint32_t Q = __funnelshift_r(P, P, 8 * ((4 - fill) & 3));  // the stream, re-aligned to image words
        int m = nw - 1;
        if (m > 0) {
            while (m > 0 && (reinterpret_cast<uintptr_t>(wp) & 15)) { *wp++ = Q; m--; }
            const uint4 Q4 = make_uint4(Q, Q, Q, Q);
            while (m >= 4) { *reinterpret_cast<uint4*>(wp) = Q4; wp += 4; m -= 4; }
            while (m > 0) { *wp++ = Q; m--; }
        }
        acc = rem ? (Q & ((1u << (8 * rem)) - 1u)) : 0u;
        fill = rem;
    }
    __device__ __forceinline__ void flush_tail() {                     // only the image's last, partial word
        uint8_t* bp = reinterpret_cast<uint8_t*>(wp);
        for (int i = 0; i < fill; i++) bp[i] = (uint8_t)(acc >> (8 * i));
        fill = 0; acc = 0;
    }
};

__global__ void __launch_bounds__(kXThreads, VCFC_DEC_XCTAS)
k_dec_expand(const uint8_t* __restrict__ in, const unsigned long long* __restrict__ line_start,
             const unsigned long long* __restrict__ off, unsigned long long n_lines, unsigned long long total,
             const unsigned int* __restrict__ first_line, const unsigned int* __restrict__ first_chunk,
             const unsigned* __restrict__ rq_arr, const unsigned* __restrict__ gtab, uint8_t* __restrict__ out,
             const Ctrl* __restrict__ ctrl, const uint8_t* __restrict__ gflag, unsigned per_g) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    if (ctrl->irregular) return;
    if (gflag && !gflag[blockIdx.x / per_g]) return;    // (a mixed block: this tile belongs to k_dec_expand_grid)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long T0 = (unsigned long long)blockIdx.x * kTile;
    const int tile_len = (int)(total - T0 < (unsigned long long)kTile ? total - T0 : (unsigned long long)kTile);
    const unsigned long long T1 = T0 + (unsigned long long)tile_len;
    const int s_lo = tid * kSpan, s_hi = min(s_lo + kSpan, tile_len);   // this thread's span of the tile image
    unsigned long long k0 = first_line[blockIdx.x];
    unsigned fc_tile = first_chunk[blockIdx.x];         // applies to the first batch's first line only
    Writer wr;
    wr.wp = reinterpret_cast<uint32_t*>(sm.stage + s_lo);
    wr.acc = 0;
    wr.fill = 0;

    for (;;) {
        // (1) line table entries of the next batch: lines k0 .. k0 + nb - 1 overlap the tile
        if (warp == 0) {
            const unsigned long long kk = k0 + (unsigned long long)lane;
            const unsigned long long o = kk < n_lines ? off[kk] : total;
            const unsigned long long ls = kk <= n_lines ? line_start[kk] : 0ull;
            const unsigned long long ls0 = __shfl_sync(0xffffffffu, ls, 0);
            // a long first line whose text began before the tile is staged from the chunk that holds the tile's first byte
            unsigned long long s0 = ls0;
            int rq0 = 0;
            if (fc_tile != kWholeLine) { rq0 = (int)rq_arr[k0]; s0 = ls0 + 8ull + (unsigned long long)rq0 + 16ull * fc_tile; }
            const bool overlaps = kk < n_lines && o < T1;
            unsigned m = __ballot_sync(0xffffffffu, overlaps) & 0x7fffffffu;         // lane 31 only supplies the end of line 30
            int nb = __popc(m);                                                        // off[] is monotone: a prefix of the lanes
            // later lines are staged whole, so they must fit behind the first; the first may be cut at kCmax
            // (kCmax compressed bytes always expand to at least a tile of text)
            const unsigned fits = __ballot_sync(0xffffffffu, ls >= s0 && ls - s0 <= (unsigned long long)kCmax);
            while (nb > 1 && !((fits >> nb) & 1u)) nb--;
            const unsigned long long o_next = __shfl_down_sync(0xffffffffu, o, 1);
            const unsigned long long ls_next = __shfl_down_sync(0xffffffffu, ls, 1);
            if (lane < nb) {
                sm.l_pos[lane] = (int)((long long)o - (long long)T0);
                sm.l_end[lane] = (int)(o_next < T1 ? o_next - T0 : (unsigned long long)tile_len);
                sm.l_last[lane] = (int)min((long long)o_next - 1 - (long long)T0, (long long)(1 << 30));
                sm.l_coff[lane] = lane == 0 ? 0 : (int)(ls - s0);
                if (lane == nb - 1) sm.l_coff[nb] = (int)min(ls_next - s0, (unsigned long long)0x7fffffff);
            }
            const unsigned long long o_after = __shfl_sync(0xffffffffu, o, nb);       // nb <= 31
            if (lane == 0) {
                sm.n_batch = nb;
                sm.more = (nb > 0 && k0 + (unsigned long long)nb < n_lines && o_after < T1) ? 1 : 0;   // another batch follows
                sm.c_lo = s0;
                sm.c_first = fc_tile != kWholeLine ? (int)fc_tile : -1;
                sm.rq0 = rq0;
            }
        }
        __syncthreads();
        const int nb = sm.n_batch;
        if (nb == 0) break;
        const unsigned long long c_lo = sm.c_lo;
        const int c_first = sm.c_first;
        // (2) stage the batch's compressed bytes; cbuf keeps the source's 16-byte phase so both sides are aligned
        const int phase = (int)(reinterpret_cast<uintptr_t>(in + c_lo) & 15);
        const int staged = min(sm.l_coff[nb], kCmax);
        {
            const uint8_t* src = in + c_lo;
            const int head = min((16 - phase) & 15, staged);
            if (tid < head) sm.cbuf[phase + tid] = src[tid];
            const int n16 = (staged - head) >> 4;
            const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
            uint4* d4 = reinterpret_cast<uint4*>(sm.cbuf + phase + head);
            for (int i = tid; i < n16; i += kXThreads) d4[i] = s4[i];
            const int t0 = head + 16 * n16;
            if (tid < staged - t0) sm.cbuf[phase + t0 + tid] = src[t0 + tid];
        }
        // (3) chunk tables of the batch's lines (built once per line by k_dec_sizes): copy the staged part.
        //     Line li's entries go to slot (l_coff >> 4) + li, the global table's rule relative to the batch.
        for (int li = warp; li < nb; li += kXWarps) {
            const int coff = sm.l_coff[li];
            const int avail = min(sm.l_coff[li + 1], staged) - coff;          // staged bytes of this line
            const int nslots = (avail >> 4) + 1;
            const unsigned long long lsk = li == 0 ? line_start[k0] : c_lo + (unsigned long long)coff;
            const unsigned* src = gtab + (lsk >> 4) + (k0 + (unsigned long long)li) + (li == 0 && c_first >= 0 ? c_first : 0);
            unsigned* tab = sm.ctab + (coff >> 4) + li;
            for (int i = lane; i < nslots; i += 32) tab[i] = src[i];
        }
        __syncthreads();
        // (4) generate: the lines of the batch that intersect this thread's span, strictly left to right
        for (int li = 0; li < nb; li++) {
            const int lpos = sm.l_pos[li], lend = sm.l_end[li];
            if (lend <= s_lo) continue;
            if (lpos >= s_hi) break;
            const int g_lo = max(s_lo, max(lpos, 0)), g_hi = min(s_hi, lend);
            if (g_lo >= g_hi) continue;
            const int coff = sm.l_coff[li];
            const bool cont = li == 0 && c_first >= 0;                // the line is staged from chunk c_first of its tokens
            const uint8_t* lp = sm.cbuf + phase + coff;               // whole-line mode: the line's first byte
            const int rq = cont ? sm.rq0 : (int)hdr_len(lp + 4);
            const int tb = cont ? -16 * c_first : 8 + rq;             // lp[tb + ci] = token byte ci of the line
            const int avail = min(sm.l_coff[li + 1], staged) - coff;  // staged bytes of this line
            const int tpos = lpos + rq;                               // tile position of the first sample's text
            int sp = g_lo;
            // required section passes through (compress.cpp:788-807); one that is not (wholly) staged comes from global memory
            {
                const uint8_t* rqp = lp + 8;
                if (cont || 8 + rq > avail) rqp = in + (li == 0 ? line_start[k0] : c_lo + (unsigned long long)coff) + 8;
                for (; sp < g_hi && sp < tpos; sp++) wr.byte(rqp[sp - lpos]);
            }
            if (sp < g_hi) {
                const int cbase = cont ? c_first : 0;                 // first chunk whose table entry is in smem
                const unsigned* tab = sm.ctab + (coff >> 4) + li - cbase;
                int x = sp - tpos;                                    // text offset inside the sample text
                const int x_end = g_hi - tpos;
                const int x_nl = sm.l_last[li] - tpos;                // the line's last text byte: '\n' instead of the tab
                // chunk whose start offset is the last one <= x, among the staged chunks
                int lo = cbase, hi = cbase + ((avail - (cont ? 0 : 8 + rq) + 15) >> 4);
                if (hi <= lo) hi = lo + 1;
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if ((int)(tab[mid] >> 1) <= x) lo = mid; else hi = mid; }
                int ci = lo << 4, cur = (int)(tab[lo] >> 1);
                bool payload = (tab[lo] & 1u) != 0;
                // skip what lies before x inside the chunk (tight loop), then emit
                for (;;) {
                    const uint32_t c = lp[tb + ci];
                    if (payload) {
                        if (cur == x) break;
                        cur++; ci++;
                        if (c == 9u || c == 10u) payload = false;
                    } else if (c >= 0xE0u) {
                        payload = true; ci++;
                    } else {
                        const int len = 4 * (int)(c < 0x80u ? c : (c & 0x1Fu));
                        if (cur + len > x) break;
                        cur += len; ci++;
                    }
                }
                while (x < x_end) {
                    const uint32_t c = lp[tb + ci];
                    if (payload) {
                        wr.byte(c); x++;
                        ci++;
                        if (c == 9u || c == 10u) payload = false;
                        cur = x;
                    } else if (c >= 0xE0u) {
                        payload = true; ci++;
                    } else {
                        const int len = 4 * (int)(c < 0x80u ? c : (c & 0x1Fu));
                        const int stop = min(cur + len, x_end);
                        const bool nl = stop > x_nl;                   // this run carries the line's final byte
                        wr.run(sample_word(c), (x - cur) & 3, stop - x - (nl ? 1 : 0));
                        if (nl) wr.byte('\n');
                        x = stop;
                        cur += len; ci++;                              // (a run cut by x_end also ends the loop)
                    }
                }
            }
        }
        const int more = sm.more;
        __syncthreads();
        if (!more) break;
        k0 += (unsigned long long)nb;
        fc_tile = kWholeLine;
    }

    if (wr.fill) wr.flush_tail();
    __syncthreads();
    // tile image -> HBM
    {
        uint8_t* dst = out + T0;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            const int n16 = tile_len >> 4;
            const uint4* s4 = reinterpret_cast<const uint4*>(sm.stage);
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            for (int i = tid; i < n16; i += kXThreads) d4[i] = s4[i];
            for (int i = (n16 << 4) + tid; i < tile_len; i += kXThreads) dst[i] = sm.stage[i];
        } else {
            for (int i = tid; i < tile_len; i += kXThreads) dst[i] = sm.stage[i];
        }
    }
}

// ---- D4', sample text on the 4-byte grid (every sample column is 3 bytes wide): fill and patch -------------------------
// Almost all of the text is the default genotype, so the tile image is first FILLED with "0|0\t" in each line's phase
// (16-byte stores, no parsing) and then PATCHED: one thread per 16 token bytes walks its chunk from the chunk table's
// text offset and writes only what differs -- a '1' for the allele bytes of 0|1 / 1|0 / 1|1 runs, literal payloads --
// plus the required sections and the line ends.  Work is proportional to the COMPRESSED size of the tile.
#ifndef VCFC_DEC_ASYNC_TAB
#define VCFC_DEC_ASYNC_TAB 0
#endif
#ifndef VCFC_DEC_RUNFN
#define VCFC_DEC_RUNFN 0
#endif
#ifndef VCFC_DEC_FILLPTR
#define VCFC_DEC_FILLPTR 0
#endif
#ifndef VCFC_DEC_RUNDO
#define VCFC_DEC_RUNDO 0
#endif
#ifndef VCFC_DEC_BULK_LD
#define VCFC_DEC_BULK_LD 1
#endif
#ifndef VCFC_DEC_BULK_ST
#define VCFC_DEC_BULK_ST 1
#endif
#ifndef VCFC_DEC_LINEFILL
#define VCFC_DEC_LINEFILL 8
#endif
#ifndef VCFC_DEC_GTHREADS
#define VCFC_DEC_GTHREADS 256
#define VCFC_DEC_GCTAS 5
#define VCFC_DEC_GCMAX 8192
#define VCFC_DEC_GTILE 32768
#endif
constexpr int kTileG = VCFC_DEC_GTILE;         // output tile of the fill-and-patch kernel
constexpr int kGThreads = VCFC_DEC_GTHREADS, kGWarps = kGThreads / 32;
constexpr int kCmaxG = VCFC_DEC_GCMAX;      // compressed bytes staged per batch; a line that does not fit is continued in the next batch
constexpr int kChunksG = kCmaxG / 16 + 2 * 32;
static_assert(kCmaxG >= 2048 && kCmaxG % 16 == 0, "a batch must at least hold a line start (8 + 960 bytes) and some chunks");

// ---- bulk asynchronous copies (sm_90+ / sm_100a): one thread moves a whole buffer, completion on an mbarrier (loads) or a
//      bulk group (stores); addresses and sizes are multiples of 16 -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// a run of 0|1 / 1|0 / 1|1 samples that crosses the tile's first or last byte: every byte tested
__device__ __noinline__ void patch_run_checked(uint8_t* stage, int pos, int cnt, bool a1, bool b1, int tile_len) {
    for (int q = 0; q < cnt; q++, pos += 4) {
        if (a1 && (unsigned)pos < (unsigned)tile_len) stage[pos] = '1';
        if (b1 && (unsigned)(pos + 2) < (unsigned)tile_len) stage[pos + 2] = '1';
    }
}

struct SmemG {
    alignas(16) uint8_t stage[kTileG];
    alignas(16) uint8_t cbuf[kCmaxG + 32];
    unsigned ctab[kChunksG];
    int l_pos[kMaxL + 1], l_end[kMaxL + 1], l_last[kMaxL + 1], l_coff[kMaxL + 2];
    int l_rq[kMaxL + 1];                     // required length
    int l_c0[kMaxL + 2];                     // first flat chunk index of the line (exclusive prefix of the chunk counts)
    int l_wb[kMaxL + 1];                     // token bytes of the line to walk (the line's final '\n' excluded)
    int n_batch, more, staged, c_first;
    int cut_next;                            // the batch's only line was cut: chunk to continue from (else -1)
    unsigned long long c_lo;
    alignas(8) unsigned long long bar;       // mbarrier of the compressed-byte bulk load
};

__global__ void __launch_bounds__(kGThreads, VCFC_DEC_GCTAS)
k_dec_expand_grid(const uint8_t* __restrict__ in, const unsigned long long* __restrict__ line_start,
                  const unsigned long long* __restrict__ off, unsigned long long n_lines, unsigned long long total,
                  const unsigned int* __restrict__ first_line, const unsigned int* __restrict__ first_chunk,
                  const unsigned* __restrict__ rq_arr, const unsigned* __restrict__ gtab, uint8_t* __restrict__ out,
                  const Ctrl* __restrict__ ctrl, const uint8_t* __restrict__ gflag, unsigned per_g) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SmemG& sm = *reinterpret_cast<SmemG*>(smem_raw);
    if (ctrl->irregular) return;
    // a mixed block: the tile map has one entry per tile of the span-walking kernel (per_g of them in one of these tiles), and
    // the tiles that a line off the 4-byte grid overlaps are that kernel's
    if (gflag && gflag[blockIdx.x]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long T0 = (unsigned long long)blockIdx.x * kTileG;
    const int tile_len = (int)(total - T0 < (unsigned long long)kTileG ? total - T0 : (unsigned long long)kTileG);
    const unsigned long long T1 = T0 + (unsigned long long)tile_len;
    unsigned long long k0 = first_line[blockIdx.x * per_g];
    unsigned fc_tile = first_chunk[blockIdx.x * per_g]; // chunk the batch's first line is staged from (kWholeLine: from its header)
    bool contd = false;                                  // that line was begun by an earlier batch of this tile
    if (VCFC_DEC_BULK_LD && tid == 0) mbar_init(&sm.bar, 1);                 // (its first use is by this same thread; the others wait behind a barrier)
    uint32_t parity = 0;

    for (;;) {
        // (1) line table entries of the next batch: lines k0 .. k0 + nb - 1 overlap the tile
        if (warp == 0) {
            const unsigned long long kk = k0 + (unsigned long long)lane;
            const unsigned long long o = kk < n_lines ? off[kk] : total;
            const unsigned long long ls = kk <= n_lines ? line_start[kk] : 0ull;
            int rq = kk < n_lines ? (int)rq_arr[kk] : 0;
            const unsigned long long ls0 = __shfl_sync(0xffffffffu, ls, 0);
            const int rq0 = __shfl_sync(0xffffffffu, rq, 0);
            // a long first line whose text began before the tile is staged from the chunk that holds the tile's first byte
            unsigned long long s0 = ls0;
            const bool cont = fc_tile != kWholeLine;
            if (cont) s0 = ls0 + 8ull + (unsigned long long)rq0 + 16ull * fc_tile;
            const bool overlaps = kk < n_lines && o < T1;
            unsigned m = __ballot_sync(0xffffffffu, overlaps) & 0x7fffffffu;         // lane 31 only supplies the end of line 30
            int nb = __popc(m);                                                        // off[] is monotone: a prefix of the lanes
            // later lines are staged whole, so they must fit behind the first; a first line that does not fit is cut at a
            // chunk boundary and continued by the next batch
            const unsigned fits = __ballot_sync(0xffffffffu, ls >= s0 && ls - s0 <= (unsigned long long)kCmaxG);
            while (nb > 1 && !((fits >> nb) & 1u)) nb--;
            const unsigned long long o_next = __shfl_down_sync(0xffffffffu, o, 1);
            const unsigned long long ls_next = __shfl_down_sync(0xffffffffu, ls, 1);
            const unsigned long long ls_after = __shfl_sync(0xffffffffu, ls, nb);     // nb <= 31
            const int staged = (int)min(ls_after - s0, (unsigned long long)kCmaxG);
            // token bytes to walk and their 16-byte chunks
            const int coff = lane == 0 ? 0 : (int)(ls - s0);
            const int cend = (int)min(ls_next - s0, (unsigned long long)0x7fffffff);
            const bool whole = cend <= staged;                                         // the line's final '\n' is staged
            const int hdrb = (lane == 0 && cont) ? 0 : 8 + rq;
            int wb = lane < nb ? min(cend, staged) - coff - hdrb - (whole ? 1 : 0) : 0;
            if (wb < 0) wb = 0;
            if (!whole) wb &= ~15;                                                     // a cut line is resumed at a chunk boundary
            int nch = (wb + 15) >> 4, inc = nch;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
            if (lane <= nb) sm.l_c0[lane] = inc - nch;
            if (lane < nb) {
                sm.l_pos[lane] = (int)((long long)o - (long long)T0);
                sm.l_end[lane] = (int)(o_next < T1 ? o_next - T0 : (unsigned long long)tile_len);
                sm.l_last[lane] = (int)min((long long)o_next - 1 - (long long)T0, (long long)(1 << 30));
                sm.l_coff[lane] = coff;
                sm.l_rq[lane] = rq;
                sm.l_wb[lane] = wb;
                if (lane == nb - 1) sm.l_coff[nb] = cend;
            }
            const unsigned long long o_after = __shfl_sync(0xffffffffu, o, nb);
            if (lane == 0) {
                sm.n_batch = nb;
                sm.more = (nb > 0 && k0 + (unsigned long long)nb < n_lines && o_after < T1) ? 1 : 0;   // another batch follows
                sm.c_lo = s0;
                sm.c_first = cont ? (int)fc_tile : -1;
                sm.staged = staged;
                sm.cut_next = (nb == 1 && !whole) ? (cont ? (int)fc_tile : 0) + (wb >> 4) : -1;
                // (2) the batch's compressed bytes: ONE bulk copy, in flight while the tile image is filled.  cbuf keeps the
                //     source's 16-byte phase, so the copy runs from the aligned address below the first byte
                if (VCFC_DEC_BULK_LD && nb > 0) {
                    const uintptr_t ga = reinterpret_cast<uintptr_t>(in + s0);
                    const uint32_t bytes = (uint32_t)(((ga & 15) + (uintptr_t)staged + 15) & ~uintptr_t(15));
                    bulk_load(sm.cbuf, reinterpret_cast<const void*>(ga & ~uintptr_t(15)), bytes, &sm.bar);
                }
            }
        }
        __syncthreads();
        const int nb = sm.n_batch;
        if (nb == 0) break;
        const unsigned long long c_lo = sm.c_lo;
        const int c_first = sm.c_first;
        const int phase = (int)(reinterpret_cast<uintptr_t>(in + c_lo) & 15);
        const int staged = sm.staged;
#if !VCFC_DEC_BULK_LD
        {
            const uint8_t* src = in + c_lo;
            const int head = min((16 - phase) & 15, staged);
            if (tid < head) sm.cbuf[phase + tid] = src[tid];
            const int n16 = (staged - head) >> 4;
            const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
            uint4* d4 = reinterpret_cast<uint4*>(sm.cbuf + phase + head);
            for (int i = tid; i < n16; i += kGThreads) d4[i] = s4[i];
            const int t0 = head + 16 * n16;
            if (tid < staged - t0) sm.cbuf[phase + t0 + tid] = src[t0 + tid];
        }
#endif
        // (3) chunk tables of the batch's lines (built once per line by k_dec_sizes): line li's entries go to slot
        //     (l_coff >> 4) + li, the global table's rule relative to the batch
        for (int li = warp; li < nb; li += kGWarps) {
            const int coff = sm.l_coff[li];
            const int nslots = ((min(sm.l_coff[li + 1], staged) - coff) >> 4) + 1;
            const unsigned long long lsk = li == 0 ? line_start[k0] : c_lo + (unsigned long long)coff;
            const unsigned* src = gtab + (lsk >> 4) + (k0 + (unsigned long long)li) + (li == 0 && c_first >= 0 ? c_first : 0);
            unsigned* tab = sm.ctab + (coff >> 4) + li;
#if VCFC_DEC_ASYNC_TAB
            for (int i = lane; i < nslots; i += 32)      // asynchronous 4-byte copies: nobody waits for them before the fill
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(tab + i)), "l"(src + i) : "memory");
#else
            for (int i = lane; i < nslots; i += 32) tab[i] = src[i];
#endif
        }
        // (4) fill: every 16-byte unit that starts inside the batch's lines gets "0|0\t" in the phase of its line
        {
            const int f_first = contd ? (nb > 1 ? sm.l_pos[1] : sm.l_end[0]) : sm.l_pos[0];   // a continued line is already filled
            const int f_lo = f_first <= 0 ? 0 : (f_first + 15) >> 4, f_hi = (sm.l_end[nb - 1] + 15) >> 4;
            if (nb <= VCFC_DEC_LINEFILL) {
                // few long lines: line by line, one constant pattern each (a unit belongs to the line it starts in)
                for (int li = contd ? 1 : 0; li < nb; li++) {
                    const int lp = sm.l_pos[li];
                    const int lo = lp <= 0 ? 0 : (lp + 15) >> 4;
                    const int hi = li + 1 < nb ? min((sm.l_pos[li + 1] + 15) >> 4, f_hi) : f_hi;
                    const int ph = ((lo << 4) - lp - sm.l_rq[li]) & 3;
                    const uint32_t P = __funnelshift_r(0x09307C30u, 0x09307C30u, 8 * ph);
                    const uint4 P4 = make_uint4(P, P, P, P);
#if VCFC_DEC_FILLPTR
                    uint4* q4 = reinterpret_cast<uint4*>(sm.stage) + lo + tid;
                    uint4* const e4 = reinterpret_cast<uint4*>(sm.stage) + hi;
#pragma unroll 4
                    for (; q4 < e4; q4 += kGThreads) *q4 = P4;
#else
                    for (int u = lo + tid; u < hi; u += kGThreads) *reinterpret_cast<uint4*>(sm.stage + (u << 4)) = P4;
#endif
                }
            } else {
                int li = 0;
                for (int u = f_lo + tid; u < f_hi; u += kGThreads) {
                    const int pos = u << 4;
                    while (li + 1 < nb && sm.l_pos[li + 1] <= pos) li++;          // (a thread's units ascend)
                    const int ph = (pos - sm.l_pos[li] - sm.l_rq[li]) & 3;
                    const uint32_t P = __funnelshift_r(0x09307C30u, 0x09307C30u, 8 * ph);
                    *reinterpret_cast<uint4*>(sm.stage + pos) = make_uint4(P, P, P, P);
                }
            }
        }
#if VCFC_DEC_ASYNC_TAB
        asm volatile("cp.async.wait_all;" ::: "memory");   // this thread's chunk-table copies (the barrier below publishes them)
#endif
#if VCFC_DEC_BULK_LD
        mbar_wait(&sm.bar, parity);                        // the compressed bytes have landed
        parity ^= 1u;
#endif
        __syncthreads();
        // (5) patch: required sections and line ends (one warp per line) ...
        for (int li = warp; li < nb; li += kGWarps) {
            const int lpos = sm.l_pos[li], rq = sm.l_rq[li];
            if (!(li == 0 && contd)) {                                // (a continued line's required section is already there)
                const bool cont0 = li == 0 && c_first >= 0;
                const uint8_t* lp = sm.cbuf + phase + sm.l_coff[li] + 8;
                // not (wholly) staged -- a long required section, or a line staged from its tokens on: from global memory
                if (cont0 || 8 + rq > min(sm.l_coff[li + 1], staged) - sm.l_coff[li])
                    lp = in + (li == 0 ? line_start[k0] : c_lo + (unsigned long long)sm.l_coff[li]) + 8;
                const int i_lo = lpos < 0 ? -lpos : 0, i_hi = min(rq, tile_len - lpos);
                for (int i = i_lo + lane; i < i_hi; i += 32) sm.stage[lpos + i] = lp[i];
            }
            const int last = sm.l_last[li];
            if (lane == 0 && last >= 0 && last < tile_len) sm.stage[last] = '\n';
        }
        // ... and the token chunks: only what is not the default genotype is written
        {
            // two threads per chunk: the second one starts at byte 8, with the text offset (and literal state) the first 8
            // bytes lead to -- word-parallel when they are run tokens only (nearly always)
            const int n_items = 2 * sm.l_c0[nb];
            int li = 0;
            for (int idx = tid; idx < n_items; idx += kGThreads) {
                const int ch = idx >> 1, half = idx & 1;
                while (li + 1 < nb && sm.l_c0[li + 1] <= ch) li++;
                const int c = ch - sm.l_c0[li];
                const int coff = sm.l_coff[li];
                const bool cont = li == 0 && c_first >= 0;
                const int rq = sm.l_rq[li];
                const unsigned* tab = sm.ctab + (coff >> 4) + li;
                const unsigned t0 = tab[c];
                const int tpos = sm.l_pos[li] + rq;                  // tile position of the first sample's text
                int pos = tpos + (int)(t0 >> 1);                     // tile position of the chunk's first text byte
                if (pos >= tile_len) continue;
                const int wb = sm.l_wb[li];
                if (16 * (c + 1) < wb && tpos + (int)(tab[c + 1] >> 1) <= 0) continue;     // ends before the tile
                const int nby = min(16, wb - 16 * c);
                // the chunk's bytes: aligned 32-bit loads, funnel-shifted
                const uint8_t* bp = sm.cbuf + phase + coff + (cont ? 0 : 8 + rq) + 16 * c;
                const uintptr_t ga = reinterpret_cast<uintptr_t>(bp);
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(ga & ~uintptr_t(3));
                const int sh = 8 * (int)(ga & 3);
                const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
                const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh),
                               v3 = __funnelshift_r(w3, w4, sh);
                bool payload = (t0 & 1u) != 0;
                const uint32_t lit_lo = ((v0 & (v0 << 1) & (v0 << 2)) | (v1 & (v1 << 1) & (v1 << 2))) & 0x80808080u;   // a byte >= 0xE0
                int b_lo = 0, b_hi = nby;
                if (nby <= 8) {
                    if (half) continue;                                // a short last chunk: one thread
                } else if (half) {
                    if (!payload && !lit_lo) {                         // run tokens only: their lengths add up word-parallel
                        const uint32_t c0 = v0 & (0x7F7F7F7Fu ^ (((v0 >> 7) & 0x01010101u) * 0x60u));
                        const uint32_t c1 = v1 & (0x7F7F7F7Fu ^ (((v1 >> 7) & 0x01010101u) * 0x60u));
                        pos += 4 * (int)__dp4a(c1, 0x01010101u, __dp4a(c0, 0x01010101u, 0u));
                    } else {                                           // literals: follow the state through the first 8 bytes
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const uint32_t b = ((i < 4 ? v0 : v1) >> (8 * (i & 3))) & 0xFFu;
                            if (payload) { pos++; if (b == 9u || b == 10u) payload = false; }
                            else if (b >= 0xE0u) payload = true;
                            else pos += 4 * (int)(b < 0x80u ? b : (b & 0x1Fu));
                        }
                    }
                    if (pos >= tile_len) continue;
                    b_lo = 8;
                } else {
                    b_hi = 8;
                }
                for (int b0 = b_lo; b0 < b_hi; b0 += 8) {
                    const uint32_t wa = b0 ? v2 : v0, wbb = b0 ? v3 : v1;
                    const int nb8 = b_hi - b0;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t b = ((i < 4 ? wa : wbb) >> (8 * (i & 3))) & 0xFFu;
                        if (i < nb8) {
                            if (payload) {
                                if ((unsigned)pos < (unsigned)tile_len) sm.stage[pos] = (uint8_t)b;
                                pos++;
                                if (b == 9u || b == 10u) payload = false;
                            } else if (b >= 0xE0u) {
                                payload = true;
                            } else if (b < 0x80u) {
                                pos += 4 * (int)b;                        // 0|0: already there
                            } else {
                                const uint32_t f = b & 0xE0u;
                                const bool a1 = f != kTok01, b1 = f != kTok10;
                                const int cnt = (int)(b & 0x1Fu);
#if VCFC_DEC_RUNFN
                                if ((unsigned)pos <= (unsigned)(tile_len - 4 * cnt)) {     // the whole run lies inside the tile (nearly always)
                                    for (int q = 0; q < cnt; q++, pos += 4) {
                                        if (a1) sm.stage[pos] = '1';
                                        if (b1) sm.stage[pos + 2] = '1';
                                    }
                                } else {
                                    patch_run_checked(sm.stage, pos, cnt, a1, b1, tile_len);   // (out of line: tile edges only)
                                    pos += 4 * cnt;
                                }
#elif VCFC_DEC_RUNDO
                                const int pe = pos + 4 * cnt;                 // (cnt >= 1: k_dec_sizes rejects empty runs)
                                do {
                                    if (a1 && (unsigned)pos < (unsigned)tile_len) sm.stage[pos] = '1';
                                    if (b1 && (unsigned)(pos + 2) < (unsigned)tile_len) sm.stage[pos + 2] = '1';
                                    pos += 4;
                                } while (pos != pe);
#else
                                for (int q = 0; q < cnt; q++, pos += 4) {
                                    if (a1 && (unsigned)pos < (unsigned)tile_len) sm.stage[pos] = '1';
                                    if (b1 && (unsigned)(pos + 2) < (unsigned)tile_len) sm.stage[pos + 2] = '1';
                                }
#endif
                            }
                        }
                    }
                }
            }
        }
        // a cut line goes on in the next batch unless what is left of it lies behind the tile
        const int cut_next = sm.cut_next;
        int more = sm.more;
        if (cut_next >= 0) {
            const unsigned* tab = sm.ctab + (c_first >= 0 ? -c_first : 0);       // line 0: slot 0 holds chunk max(c_first, 0)
            more = sm.l_pos[0] + sm.l_rq[0] + (int)(tab[cut_next] >> 1) < tile_len ? 1 : 0;
        }
        __syncthreads();
        if (!more) break;
        if (cut_next >= 0) { fc_tile = (unsigned)cut_next; contd = true; }
        else { k0 += (unsigned long long)nb; fc_tile = kWholeLine; contd = false; }
    }
    if (VCFC_DEC_BULK_ST) fence_async_smem();              // this thread's writes to the tile image, ordered before the bulk store reads them
    __syncthreads();
    // tile image -> HBM: one bulk store of the 16-byte multiple (tile starts are 16-byte aligned when `out` is), the last
    // tile's odd bytes by hand
    {
        uint8_t* dst = out + T0;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#if VCFC_DEC_BULK_ST
            const int nbulk = tile_len & ~15;
            if (tid == 0 && nbulk) bulk_store(dst, sm.stage, (uint32_t)nbulk);
            for (int i = nbulk + tid; i < tile_len; i += kGThreads) dst[i] = sm.stage[i];
            if (tid == 0 && nbulk) bulk_store_wait_read();  // the image must outlive the copy's reads
#else
            const int n16 = tile_len >> 4;
            const uint4* s4 = reinterpret_cast<const uint4*>(sm.stage);
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            for (int i = tid; i < n16; i += kGThreads) d4[i] = s4[i];
            for (int i = (n16 << 4) + tid; i < tile_len; i += kGThreads) dst[i] = sm.stage[i];
#endif
        } else {
            for (int i = tid; i < tile_len; i += kGThreads) dst[i] = sm.stage[i];
        }
    }
}

__global__ void k_dec_result(vcfc_result* r, const Ctrl* ctrl, int status, unsigned long long out_len, unsigned long long n_lines) {
    r->reserved = 0;
    r->err_line = 0;
    if (ctrl && ctrl->irregular) { r->status = kStatusIrregular; r->out_len = 0; r->n_lines = 0; }
    else { r->status = status; r->out_len = out_len; r->n_lines = n_lines; }
}

}  // namespace dec

int decode_fast(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, uint64_t sample_count, uint8_t* d_out, size_t out_cap,
                vcfc_result* d_result, bool size_only, cudaStream_t stream) {
    using namespace dec;
    ctx->last_path = kPathFast;
    if (in_len < 8) {                      // nothing but (at most) a few trailing bytes: clean EOF
        k_dec_result<<<1, 1, 0, stream>>>(d_result, nullptr, VCFC_OK, 0, 0);
        ctx->launches++;
        return VCFC_OK;
    }
    if (sample_count == 0 || in_len >= (1ull << 46)) {
        k_dec_result<<<1, 1, 0, stream>>>(d_result, nullptr, kStatusIrregular, 0, 0);
        ctx->launches++;
        return VCFC_OK;
    }
    if (!ctx->dec_attr_set) {                                  // per context (= per device)
        VCFC_CUDA(ctx, cudaFuncSetAttribute(k_dec_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        VCFC_CUDA(ctx, cudaFuncSetAttribute(k_dec_expand_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemG)));
        ctx->dec_attr_set = 1;
    }
    int rc;
    const long long n = (long long)in_len, n_seg = (n + kSeg - 1) / kSeg;
    DevBuf &b_ctrl = ctx->ws[2], &b_seg = ctx->ws[0], &b_scr = ctx->ws[1], &b_ls = ctx->ws[3], &b_sizes = ctx->ws[4],
           &b_offs = ctx->ws[7], &b_tiles = ctx->ws[5], &b_tab = ctx->ws[8], &b_rq = ctx->ws[6];
    if ((rc = dev_reserve(ctx, &b_ctrl, sizeof(Ctrl) + 64))) return rc;
    if ((rc = dev_reserve(ctx, &b_seg, (size_t)n_seg * 8 * 4 + 64))) return rc;
    Ctrl* ctrl = (Ctrl*)b_ctrl.p;
    long long* cand = (long long*)b_seg.p;
    long long* endp = cand + n_seg;
    unsigned long long* cnt = (unsigned long long*)(endp + n_seg);
    unsigned long long* base = cnt + n_seg;
    VCFC_CUDA(ctx, cudaMemsetAsync(ctrl, 0, sizeof(Ctrl), stream));
    const unsigned gs = (unsigned)((n_seg + 127) / 128);
    if (ctx->timing) cudaEventRecord(ctx->ev[2 * kTimeDecodeScan], stream);
    k_dec_walk<<<(unsigned)((n_seg * 32 + 127) / 128), 128, 0, stream>>>(d_in, n, n_seg, cand, endp, cnt);
    k_dec_verify<<<gs, 128, 0, stream>>>(n, n_seg, cand, endp, ctrl, 0);
    k_dec_repair<<<1, 1, 0, stream>>>(d_in, n, n_seg, cand, endp, cnt, ctrl);
    k_dec_verify<<<gs, 128, 0, stream>>>(n, n_seg, cand, endp, ctrl, 1);
    ctx->launches += 4;
    if ((rc = scan_exclusive_u64(ctx, (const uint64_t*)cnt, (uint64_t*)base, (size_t)n_seg, (uint64_t*)&ctrl->n_lines, &b_scr, stream))) return rc;
    struct { int irregular, n_fix; unsigned long long n_lines, end_pos, total_out; int not_grid, pad; } h;
    if ((rc = fetch_small(ctx, ctrl, &h, sizeof(h), stream))) return rc;
    if (h.irregular || h.n_lines == 0 || h.n_lines >= (1ull << 32)) {
        k_dec_result<<<1, 1, 0, stream>>>(d_result, nullptr, kStatusIrregular, 0, 0);
        ctx->launches++;
        return VCFC_OK;
    }
    const unsigned long long n_lines = h.n_lines;
    if ((rc = dev_reserve(ctx, &b_ls, (n_lines + 2) * 8))) return rc;
    if ((rc = dev_reserve(ctx, &b_sizes, (n_lines + 1) * 8))) return rc;
    if ((rc = dev_reserve(ctx, &b_offs, (n_lines + 1) * 8))) return rc;
    if ((rc = dev_reserve(ctx, &b_tab, ((size_t)in_len / 16 + n_lines + 8) * 4))) return rc;
    if ((rc = dev_reserve(ctx, &b_rq, (n_lines + 2) * 4 + n_lines + 64))) return rc;
    uint8_t* lflag = (uint8_t*)b_rq.p + (n_lines + 2) * 4;       // per line: off the 4-byte sample grid
    unsigned long long* line_start = (unsigned long long*)b_ls.p;
    k_dec_fill<<<gs, 128, 0, stream>>>(d_in, n, n_seg, cand, base, line_start, (unsigned*)b_rq.p, n_lines, ctrl);
    k_dec_sizes<<<(unsigned)(((n_lines + kSzLines - 1) / kSzLines * 32 + 127) / 128), 128, 0, stream>>>(d_in, line_start, n_lines, sample_count,
                                                                             (unsigned long long*)b_sizes.p, (unsigned*)b_tab.p, (unsigned*)b_rq.p, lflag, ctrl);
    ctx->launches += 2;
    if ((rc = scan_exclusive_u64(ctx, (uint64_t*)b_sizes.p, (uint64_t*)b_offs.p, (size_t)n_lines, (uint64_t*)&ctrl->total_out, &b_scr, stream)))
        return rc;
    if (ctx->timing) { cudaEventRecord(ctx->ev[2 * kTimeDecodeScan + 1], stream); ctx->ev_pending[kTimeDecodeScan] = 1; }
    if ((rc = fetch_small(ctx, ctrl, &h, sizeof(h), stream))) return rc;
    if (h.irregular) {
        k_dec_result<<<1, 1, 0, stream>>>(d_result, nullptr, kStatusIrregular, 0, 0);
        ctx->launches++;
        return VCFC_OK;
    }
    const unsigned long long total = h.total_out;
    if (size_only || total > out_cap) {
        k_dec_result<<<1, 1, 0, stream>>>(d_result, nullptr, size_only ? VCFC_OK : VCFC_E_CAP, total, size_only ? n_lines : 0);
        ctx->launches++;
        return VCFC_OK;
    }
    // Lines off the 4-byte sample grid (odd-width literals, required sections < 16 bytes) need the span-walking kernel; in a
    // block that has some, only the fill-and-patch tiles such a line overlaps go there (mixed), the rest stays on k_dec_expand_grid
    static_assert(kTileG % kTile == 0, "a fill-and-patch tile is a whole number of span-walking tiles");
    const bool walk_all = ctx->force_generic == 2, mixed = h.not_grid && !walk_all;
    const unsigned per_g = (walk_all || mixed) ? (unsigned)(kTileG / kTile) : 1u;
    const unsigned long long tile = (walk_all || mixed) ? kTile : kTileG;
    const unsigned long long n_tiles = (total + tile - 1) / tile, n_gt = (total + kTileG - 1) / kTileG;
    if ((rc = dev_reserve(ctx, &b_tiles, n_tiles * 8 + n_gt + 128))) return rc;
    unsigned int* first_chunk = (unsigned int*)b_tiles.p + n_tiles;
    uint8_t* gflag = (uint8_t*)(first_chunk + n_tiles);
    if (mixed) VCFC_CUDA(ctx, cudaMemsetAsync(gflag, 0, n_gt, stream));
    k_dec_tilemap<<<(unsigned)((n_lines + 255) / 256), 256, 0, stream>>>((unsigned long long*)b_offs.p, n_lines, total, line_start,
                                                                          (const unsigned*)b_rq.p, (const unsigned*)b_tab.p,
                                                                          (unsigned int*)b_tiles.p, first_chunk, tile, lflag,
                                                                          mixed ? gflag : nullptr, (unsigned long long)kTileG);
    if (ctx->timing) cudaEventRecord(ctx->ev[2 * kTimeDecodeExpand], stream);
    if (!walk_all)
        k_dec_expand_grid<<<(unsigned)n_gt, kGThreads, sizeof(SmemG), stream>>>(d_in, line_start, (unsigned long long*)b_offs.p, n_lines,
                                                                               total, (unsigned int*)b_tiles.p, first_chunk, (const unsigned*)b_rq.p, (const unsigned*)b_tab.p, d_out, ctrl,
                                                                               mixed ? gflag : nullptr, per_g);
    if (walk_all || mixed) {
        k_dec_expand<<<(unsigned)n_tiles, kXThreads, sizeof(Smem), stream>>>(d_in, line_start, (unsigned long long*)b_offs.p, n_lines,
                                                                            total, (unsigned int*)b_tiles.p, first_chunk, (const unsigned*)b_rq.p, (const unsigned*)b_tab.p, d_out, ctrl,
                                                                            mixed ? gflag : nullptr, per_g);
        if (mixed) ctx->launches++;
    }
    if (ctx->timing) { cudaEventRecord(ctx->ev[2 * kTimeDecodeExpand + 1], stream); ctx->ev_pending[kTimeDecodeExpand] = 1; }
    k_dec_result<<<1, 1, 0, stream>>>(d_result, ctrl, VCFC_OK, total, n_lines);
    ctx->launches += 3;
    VCFC_CUDA(ctx, cudaGetLastError());
    return VCFC_OK;
}

}  // namespace vcfc
