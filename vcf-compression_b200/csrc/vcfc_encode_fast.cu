// vcfc_encode_fast.cu -- single-pass, tile-parallel encoder for data lines (sm_100a).
//
// Grammar of the tile path: '\n'-terminated lines (the last one may end with the input), single tabs, >= 10 columns, no empty
// column, required section (CHROM..FORMAT) of at most kMaxReq bytes (~32 KB: long REF / ALT / INFO columns stay on this path).
// Anything else sets ctrl->irregular and the caller reruns the block on the generic kernels (vcfc_generic.cu).
// Output bytes are those of compress_data_line (/root/reference/src/compress.cpp:5-203) for every line.
//
// k_encode_stream<kOdd>: ONE WARP per tile of 32 ... 256 KB, no CTA barriers; HBM traffic = input read once + output
// written once (+ the tile log, see 6):
//   1. cut points: a tile owns the units (one sample column, or one whole required section) that
//      START in [cut(i*T), cut((i+1)*T)); both neighbours derive the shared cut from the same bytes
//   2. look-back #1 (one word per tile): how long the run that enters the tile already is -- chunks
//      are 127 / 31 samples counted from the run's head (compress.cpp:129-170), which may lie many
//      tiles back; a tile publishes its record before doing anything else
//   3. the warp walks its tile line by line: a line start = 9th tab found 512 bytes per round trip,
//      two length headers + the required section copied through
//   4. samples in steps of 2 KB on the 4-byte GRID (every sample column 3 bytes + separator): every lane loads one 64-byte
//      block (16 phase-aligned sample words) straight into registers and classifies it into 16-bit masks: valid / coded /
//      literal / run head / closing token / allele bits; the line's '\n' ends the step
//   5. closed-form byte count per lane -> warp scan -> tokens and literals into a per-warp staging area
//   6. the tile's bytes are appended to a tile log at a position reserved with ONE atomicAdd (no scan
//      chain); two device scans over the per-tile (bytes, lines) records give the final positions,
//      k_gather_tiles moves the bytes and k_patch_headers fills in the 4-byte line-length headers
//      (they need the NEXT line's offset) and the result block.
// Sample columns that are NOT 3 bytes wide (10|0, haploid calls, GT:DP:GQ ...): <false> gives the block up at once and the host
// relaunches it on <true>, where odd_step sends the 2 KB window that holds such a column to the term walkers
// (parallel_portion: all lanes, 64-bit separator masks; literal-only windows in one pass; serial_portion: terms longer than
// the window) and the grid goes on behind it.  See DESIGN.md, "Odd-width sample columns".
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "vcfc_common.cuh"
#include "vcfc_internal.h"

namespace vcfc {
namespace enc {

constexpr int kHalo = 32768;            // how far around a nominal tile boundary a cut point is searched
constexpr int kMaxReq = kHalo - 64;     // longest required section taken by this path
constexpr int kMaxNl = 62;              // line starts per log segment (every lane keeps the offsets of two lines)

enum { kCutLine = 0, kCutSample = 1, kCutSampleFirst = 2, kCutEnd = 3, kCutBad = 4 };
constexpr int kNone = 7;                // "no open run" class
constexpr int kNoHead = -(1 << 30);

struct Ctrl {                           // one per launch, zeroed by the host
    unsigned int ticket;
    int irregular;
    int cap_exceeded;
    int line2big;
    unsigned long long total_bytes, total_lines;
    unsigned long long line_cap;
    unsigned long long log_cursor, log_cap;   // the tile log: tile outputs in arrival order, gathered by k_gather_tiles
    unsigned long long serial_bytes;          // input bytes walked term by term (odd-width samples); too many: generic kernels
    int odd_used;                             // the term walkers ran (the host keeps launching the instantiation that has them)
};

__device__ __forceinline__ bool is_sep(uint32_t c) { return c == '\t' || c == '\n'; }

// ---- cut point of nominal boundary b (global offset); warp-collective ------------------------
// Returns the global offset of the first unit start >= b and its kind.
__device__ long long cut_find(const uint8_t* __restrict__ win, long long wbase, long long vlo, long long vhi,
                              long long n, long long b, int lane, int* kind) {
    if (b <= 0) { *kind = kCutLine; return 0; }
    if (b >= n) { *kind = kCutEnd; return n; }
    // forward: first p >= b whose previous byte is a separator
    long long p = -1;
    uint32_t sepc = 0;
    long long lim = b - 1 + kMaxReq;
    for (long long base = b - 1; base < lim && base < vhi; base += 32) {
        long long g = base + lane;
        uint32_t c = g < vhi ? win[g - wbase] : 0u;
        unsigned m = __ballot_sync(0xffffffffu, is_sep(c));
        if (m) {
            int j = __ffs(m) - 1;
            p = base + j + 1;
            sepc = __shfl_sync(0xffffffffu, c, j);
            break;
        }
    }
    if (p < 0) {
        if (vhi >= n && b - 1 + kMaxReq >= n) { *kind = kCutEnd; return n; }   // no unit starts between b and the end of the input
        *kind = kCutBad;
        return b;
    }
    if (sepc == '\n') { *kind = p >= n ? kCutEnd : kCutLine; return p; }
    // backward from p-1 (a tab): count tabs until a newline, the buffer start, or 10 tabs
    int k = 0;
    bool decided = false;
    for (long long top = p - 1; !decided; top -= 32) {
        long long g = top - lane;
        bool before = g < 0, oob = g >= 0 && g < vlo;
        uint32_t c = (!before && !oob) ? win[g - wbase] : 0u;
        unsigned stop = __ballot_sync(0xffffffffu, before || oob || c == '\n');
        unsigned tabm = __ballot_sync(0xffffffffu, c == '\t');
        if (stop) {
            int j = __ffs(stop) - 1;
            k += __popc(tabm & ((1u << j) - 1u));
            if (k >= 10) { *kind = kCutSample; return p; }
            unsigned oobm = __ballot_sync(0xffffffffu, oob);
            if (oobm & (1u << j)) { *kind = kCutBad; return b; }   // required section longer than the halo
            decided = true;
        } else {
            k += __popc(tabm);
            if (k >= 10) { *kind = kCutSample; return p; }
        }
    }
    if (k == 9) { *kind = kCutSampleFirst; return p; }
    // inside a required section: the unit started before b; the next unit is this line's first sample
    int need = 9 - k;
    for (long long base = p; base < p + kMaxReq; base += 32) {
        long long g = base + lane;
        uint32_t c = g < vhi ? win[g - wbase] : 0u;
        unsigned tabm = __ballot_sync(0xffffffffu, c == '\t');
        unsigned nlm = __ballot_sync(0xffffffffu, c == '\n' || g >= vhi);
        int cnt = __popc(tabm);
        if (cnt >= need) {
            int j = __fns(tabm, 0, need);
            if (nlm & ((1u << j) - 1u)) break;
            *kind = kCutSampleFirst;
            return base + j + 1;
        }
        if (nlm) break;
        need -= cnt;
    }
    *kind = kCutBad;
    return b;
}

// ---- sample helpers -----------------------------------------------------------------------------------
// class of the 3-byte genotype at p: 0..3 = 0|0 0|1 1|0 1|1, 4 = anything else (compress.cpp:129,145)
__device__ __forceinline__ int gt_class3(const uint8_t* p) {
    uint32_t a = p[0], b = p[1], c = p[2];
    if (b != '|' || (a & 0xFEu) != 0x30u || (c & 0xFEu) != 0x30u) return 4;
    return (int)(((a & 1u) << 1) | (c & 1u));
}
__device__ __forceinline__ uint32_t cls_flag(int c) { return (0x80C0A000u >> (8 * c)) & 0xFFu; }   // utils.hpp:44-55
__device__ __forceinline__ int mod_chunk(int x, bool m127) { return m127 ? x % 127 : x % 31; }

struct Item {                  // one 64-byte block of up to 16 samples of one line; lives in registers
    int base;                  // window offset of sample 0 (block start + the line's phase); sample k sits at base + 4k
    uint32_t V, C, L, Hd, CL;  // 16-bit masks: valid, coded, literal, run head, closing token before the sample
    uint32_t Ap, Bp;           // 17-bit masks: bit k = low bit of the first / second allele of sample k-1
    int kend;                  // index of the sample that ends the line, or -1
    int pcoded;                // the sample before the first valid one is a coded sample of the same line
};

// bytes the item emits (tokens, literals, line end) -- compress.cpp:124-190 in closed form.
// ein = address of the last run head before the item.  *h = sample index (may be negative) at which the chunk that is
// open at the first valid sample began; *cfbit = the sample before which that chunk fills up (127 / 31 samples).
__device__ __forceinline__ int item_count(const Item& it, int ein, uint32_t* cfbit, int* h) {
    *cfbit = 0; *h = 0;
    if (!it.V) return 0;
    const int k0 = __ffs(it.V) - 1;
    *h = k0;
    int n = __popc(it.CL) + 5 * __popc(it.L);
    if (it.pcoded) {
        const bool m127 = (((it.Ap | it.Bp) >> k0) & 1u) == 0;   // 0|0 chunks by 127, the others by 31
        const unsigned dist = (unsigned)(it.base + 4 * k0 - 4 - ein) >> 2;   // samples between the run's head and the previous sample (< 2^24)
        const unsigned dv = m127 ? 127u : 31u;
        const unsigned md = dist - __umulhi(dist, m127 ? 0x02040811u : 0x08421085u) * dv;   // dist % dv: exact below 2^25, one multiply
        const int hh = k0 - 1 - (int)md;
        *h = hh;
        const uint32_t t = (~it.Hd & it.V) >> k0;                // leading samples that continue the entering run
        const int nlead = __ffs(~t) - 1;
        const int kf = hh + (int)dv;                             // a new chunk starts at sample kf if the run reaches it
        if (kf < k0 + nlead) { *cfbit = 1u << kf; n++; }
    }
    if (it.kend >= 0) n += ((it.L >> it.kend) & 1u) ? 0 : 2;      // literal: its tab becomes the '\n'; coded: token + '\n'
    return n;
}

// cls_flag of the genotype whose allele bits are bit 16 / bit 0 of y, as a table lookup (PRMT) in the flag constant
__device__ __forceinline__ uint32_t flag_of(uint32_t y) {
    const uint32_t c2 = y & 0x00010001u;
    return __byte_perm(0x80C0A000u, 0u, (c2 >> 15) | c2);           // selector nibble 0 = 2a + b; the other result bytes are not stored
}

// rare (warp-uniform): some lane of the step has a literal or the line's end
__device__ __forceinline__ void item_emit(const uint8_t* __restrict__ win, const Item& it, uint32_t cfbit, int h,
                                          uint8_t* __restrict__ dst, bool rare) {
    const uint32_t tok = it.V ? (it.CL | cfbit) : 0u;
    const uint32_t kendbit = it.kend >= 0 ? (1u << it.kend) : 0u;
    // run tokens only (the common case): a tight loop -- unless many lanes of the warp have literals, then one loop for all
    const bool many = rare && __popc(__ballot_sync(0xffffffffu, (it.L | kendbit) != 0u)) > 4;
    if (!rare || (!(it.L | kendbit) && !many)) {
        if (!tok) return;
        // a token's count is the distance to the token below it (to h for the lowest): walk from the top, so one FLO per token
        // finds both the next token and this one's count
        const uint32_t AB = (it.Ap << 16) | (it.Bp & 0xFFFFu);    // tokens sit at bits 0..15 here
        const uint32_t m0 = tok & (0u - tok);
        const int k0 = 31 - __clz(m0);
        dst[0] = (uint8_t)(flag_of(AB >> k0) | (uint32_t)(k0 - h));
        uint32_t t = tok;
        uint8_t* d = dst + __popc(tok) - 1;
        int k = 31 - __clz(t);
        while (t != m0) {
            t ^= 1u << k;
            const int kn = 31 - __clz(t);
            *d-- = (uint8_t)(flag_of(AB >> k) | (uint32_t)(k - kn));
            k = kn;
        }
        return;
    }
    int o = 0;
    uint32_t ev = tok | it.L | kendbit;
    while (ev) {
        const int k = __ffs(ev) - 1;
        ev &= ev - 1;
        if ((tok >> k) & 1u) {                                    // token closing the chunk that ends at sample k-1
            const uint32_t c = (((it.Ap >> k) & 1u) << 1) | ((it.Bp >> k) & 1u);
            dst[o++] = (uint8_t)(cls_flag((int)c) | (uint32_t)(k - h));
            h = k;
        }
        if ((it.L | kendbit) >> k & 1u) {
            const uint8_t* p = win + it.base + 4 * k;
            if ((it.L >> k) & 1u) {                               // literal escape (compress.cpp:171-185)
                dst[o] = (uint8_t)(kTokLit | 1u);
                dst[o + 1] = p[0]; dst[o + 2] = p[1]; dst[o + 3] = p[2];
                o += 4;
                if (k != it.kend) dst[o++] = '\t';
                h = k + 1;
            }
            if (k == it.kend) {
                if ((it.C >> k) & 1u) {                           // the open run ends with the line
                    const uint32_t c = (((it.Ap >> (k + 1)) & 1u) << 1) | ((it.Bp >> (k + 1)) & 1u);
                    dst[o++] = (uint8_t)(cls_flag((int)c) | (uint32_t)(k - h + 1));
                }
                dst[o++] = '\n';
            }
        }
    }
}

// Staged bytes -> the log, as one segment of the tile's chain (warp-collective).  Segments are 16-byte aligned: header
// {u32 bytes, u32 line starts, u64 position of the next segment}, the bytes, and the tile-relative u32 offsets of the line
// starts that lie in the segment (at most 62: every lane keeps two).  Reserved with one atomicAdd; the previous segment's header learns where this one lies.
__device__ __noinline__ void flush_segment(const uint8_t* __restrict__ stage, int nbytes, uint8_t* __restrict__ log, Ctrl* __restrict__ ctrl,
                                           unsigned long long log_cap, int lane, unsigned long long* seg_first,
                                           unsigned long long* seg_prev, bool* dead, int nl, int my_off, int my_off2) {
    __syncwarp();
    const int body = (nbytes + 15) & ~15, trailer = (4 * nl + 15) & ~15;
    const unsigned long long need = 16ull + (unsigned long long)body + (unsigned long long)trailer;
    unsigned long long pos = 0ull;
    if (lane == 0 && !*dead) {
        pos = atomicAdd(&ctrl->log_cursor, need);
        if (pos + need > log_cap) { atomicExch(&ctrl->cap_exceeded, 1); pos = ~0ull; }
    }
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (*dead || pos == ~0ull) { *dead = true; __syncwarp(); return; }
    uint8_t* const dst = log + pos;
    if (lane == 0) {
        *reinterpret_cast<uint4*>(dst) = make_uint4((unsigned)nbytes, (unsigned)nl, 0u, 0u);
        if (*seg_prev) *reinterpret_cast<unsigned long long*>(log + (*seg_prev - 1ull) + 8) = pos;
    }
    const uint4* s4 = reinterpret_cast<const uint4*>(stage);
    uint4* d4 = reinterpret_cast<uint4*>(dst + 16);
    for (int k = lane; k < (body >> 4); k += 32) d4[k] = s4[k];
    uint32_t* tr = reinterpret_cast<uint32_t*>(dst + 16 + body);
    if (lane < nl) tr[lane] = (uint32_t)my_off;
    if (lane + 32 < nl) tr[lane + 32] = (uint32_t)my_off2;
    if (!*seg_first) *seg_first = pos + 1ull;             // (+1: position 0 is a valid one)
    *seg_prev = pos + 1ull;
    __syncwarp();
}

// A line start whose required section does not fit the staging area: a segment of its own, copied from the input
// (two length headers + rq bytes; one line start at tile offset o).  Rare, byte-wise.
__device__ __noinline__ void flush_line_start(const uint8_t* __restrict__ src, int rq, int o, uint8_t* __restrict__ log, Ctrl* __restrict__ ctrl,
                                              unsigned long long log_cap, int lane, unsigned long long* seg_first,
                                              unsigned long long* seg_prev, bool* dead) {
    const int nbytes = 8 + rq, body = (nbytes + 15) & ~15;
    const unsigned long long need = 16ull + (unsigned long long)body + 16ull;
    unsigned long long pos = 0ull;
    if (lane == 0 && !*dead) {
        pos = atomicAdd(&ctrl->log_cursor, need);
        if (pos + need > log_cap) { atomicExch(&ctrl->cap_exceeded, 1); pos = ~0ull; }
    }
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (*dead || pos == ~0ull) { *dead = true; return; }
    uint8_t* const dst = log + pos;
    if (lane == 0) {
        *reinterpret_cast<uint4*>(dst) = make_uint4((unsigned)nbytes, 1u, 0u, 0u);
        if (*seg_prev) *reinterpret_cast<unsigned long long*>(log + (*seg_prev - 1ull) + 8) = pos;
        *reinterpret_cast<uint32_t*>(dst + 16 + body) = (uint32_t)o;
    }
    uint8_t* d = dst + 16;
    if (lane < 4) d[lane] = lane == 0 ? 0xC0 : 0;                       // line length: patched by k_patch_headers
    if (lane >= 4 && lane < 8) {
        const unsigned v = (unsigned)rq;
        d[lane] = lane == 4 ? (uint8_t)((v >> 24) | 0xC0) : (uint8_t)(v >> (8 * (7 - lane)));
    }
    for (int k = lane; k < rq; k += 32) d[8 + k] = src[k];
    for (int k = nbytes + lane; k < body + 4; k += 32) if (k < body) d[k] = 0;   // (padding: the gather reads whole words)
    if (!*seg_first) *seg_first = pos + 1ull;
    *seg_prev = pos + 1ull;
    __syncwarp();
}

// Look-back #1, read side (one lane): samples in the open chunk of the run that enters `tile`, 1..M.  Waits for the
// predecessors' records (each is published by its ticket holder before it waits on anything).
__device__ __noinline__ int lookback_count(const unsigned int* s1, int tile, int pc0) {
    int acc = 0;
    for (int j = tile - 1;; j--) {
        unsigned v = *((volatile const unsigned*)&s1[j]);
        while ((v >> 30) == 0) { __nanosleep(64); v = *((volatile const unsigned*)&s1[j]); }
        acc += (int)(v & 0xFFu);
        if ((v >> 30) == 2u || j == 0) break;
    }
    return mod_chunk(acc - 1 + 127 * 31, pc0 == 0) + 1;
}

// ---- the streaming kernel: one warp per tile -------------------------------------------------------------------------
#ifndef VCFC_ENC_PARWALK
#define VCFC_ENC_PARWALK 1
#endif
#ifndef VCFC_ENC_NOSERIAL
#define VCFC_ENC_NOSERIAL 0
#endif
#ifndef VCFC_ENC_PFW
#define VCFC_ENC_PFW 0          // 1 / 2: the next step's words are loaded before this step's byte-count scan / token emission
#endif
#ifndef VCFC_ENC_TICKET
#define VCFC_ENC_TICKET 1
#endif
#ifndef VCFC_ENC_LDHINT
#define VCFC_ENC_LDHINT 0
#endif
#ifndef VCFC_ENC_NOPF1
#define VCFC_ENC_NOPF1 1
#endif
#ifndef VCFC_ENC_STILE
#define VCFC_ENC_STILE 32768
#define VCFC_ENC_SWARPS 4
#define VCFC_ENC_SCTAS 7
#endif
#ifndef VCFC_ENC_SSTAGE
#define VCFC_ENC_SSTAGE 3200
#endif
constexpr int kSTile = VCFC_ENC_STILE;          // smallest nominal input bytes per tile (the host doubles it for large inputs)
constexpr int kSTileMax = 262144;           // (measured at 18 GB: 64 KB 6.68 ms, 128 KB 6.44, 256 KB 6.39, 512 KB 6.65)
constexpr int kSWarps = VCFC_ENC_SWARPS;        // warps per CTA (independent of each other)
constexpr int kSCtas = VCFC_ENC_SCTAS;
constexpr int kSStage = VCFC_ENC_SSTAGE;        // per-warp staging; flushed to the log as a segment whenever the next piece would not fit
constexpr int kStep = 2048;                     // bytes per step: 32 lanes x 64
constexpr int kBackRows = 4;                    // rows of 512 bytes per round trip of the look-back's coarse search
constexpr int kStepMaxOut = 2576;               // most bytes one step can emit: 512 literals of 5 bytes, a chunk token, the line end
static_assert(kSTile % 64 == 0, "tiles start on block boundaries");
static_assert(kSStage % 16 == 0 && kSStage >= kStepMaxOut + 512, "a step must fit the staging area (a long line start goes to the log directly)");
static_assert(kSStage >= 1024 + 1024 + 64, "the term walker splits the staging area into an output part and an input window");

struct SmemS {
    alignas(16) uint8_t stage[kSWarps][kSStage + 16];
};
// (kOdd) Everything the warp knows, parked in shared memory while odd_step runs: a value that is live across that call gets a
// stack slot for its whole life (few registers survive a call), and the hot loop then reads its state from local memory --
// measured +33 % on regular steps.  With the state saved here and read back behind the call, nothing is live across it.
struct OddSave {
    int o, flushed, nl, nl_seg, cur, ein_carry, ein0, bits, irregular;    // bits: in_req | first << 1 | need_lb << 2
    int ce, cs, pc0, tile, nt, nn, cur_ce, cur_fl, irr_seen, fl;
    int odd_a, odd_direct;                       // the step that met an odd-width sample column (direct: see odd_step)
    int par_mode;                                // the last odd-width stretch ended with odd terms: the next line goes to the walkers at once
    int action;                                  // what odd_call decided: 0 go on, 2 give the block up
    int my_off[32], my_off2[32];
};
struct SmemSO : SmemS {
    OddSave odd[kSWarps];
};

// byte / word of the input at tile-relative offset r (win = in + gb); outside [0, n) reads as 0
__device__ __forceinline__ uint32_t ldb(const uint8_t* __restrict__ win, int r, int r_lo, int r_hi) {
    return (r >= r_lo && r < r_hi) ? win[r] : 0u;
}
// the same with the bytes behind the input reading as '\n': a last line without a newline ends at EOF (compress.cpp:218 getline)
__device__ __forceinline__ uint32_t ldb_nl(const uint8_t* __restrict__ win, int r, int r_lo, int r_hi) {
    return r >= r_hi ? (uint32_t)'\n' : (r >= r_lo ? win[r] : 0u);
}
// (the byte-wise form is needed at the first / last bytes of the input only; inlined at every use it was a quarter of the kernel's code)
__device__ __noinline__ uint32_t ldw_edge(const uint8_t* __restrict__ win, int r, int r_lo, int r_hi, uint32_t fill) {
    uint32_t w = 0;
#pragma unroll 1
    for (int b = 0; b < 4; b++) w |= ((r + b >= r_lo && r + b < r_hi) ? (uint32_t)win[r + b] : fill) << (8 * b);
    return w;
}
__device__ __forceinline__ uint32_t ldw(const uint8_t* __restrict__ win, int r, int r_lo, int r_hi) {   // r multiple of 4
    if (r >= r_lo && r + 4 <= r_hi) return *reinterpret_cast<const uint32_t*>(win + r);
    return ldw_edge(win, r, r_lo, r_hi, 0u);
}

// A step that touches the first / last bytes of the input: words outside [r_lo, r_hi) read as 0 (out of line: rare)
__device__ __noinline__ void step_load_edge(const uint8_t* __restrict__ win, int blk, int r_lo, int r_hi, uint32_t* W) {
#pragma unroll 1
    for (int q = 0; q < 18; q++) {
        const int r = blk - 4 + 4 * q;
        W[q] = (r >= r_lo && r + 4 <= r_hi) ? *reinterpret_cast<const uint32_t*>(win + r) : ldw_edge(win, r, r_lo, r_hi, 0u);
    }
}

// End of the required section of the line that starts at relative offset ls, 512 bytes per round trip (16 per lane):
// returns the offset of the first sample (after the 9th tab) or -1 when irregular -- same rules as line_scan.
__device__ int line_scan16(const uint8_t* __restrict__ win, int ls, int r_lo, int r_hi, int lane) {
    if (is_sep(ldb(win, ls, r_lo, r_hi))) return -1;           // empty first column / empty line
    int tabs = 0;
    unsigned carry = 0;                                          // the byte before the round's first byte is a tab
    for (int base16 = ls & ~15; base16 < ls + kMaxReq + 16; base16 += 512) {
        const int g = base16 + 16 * lane;
        uint32_t w[4];
        if (g >= r_lo && g + 16 <= r_hi) {
            const uint4 v = *reinterpret_cast<const uint4*>(win + g);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) w[q] = ldw_edge(win, g + 4 * q, r_lo, r_hi, (uint32_t)'\n');   // past the end: a newline
        }
        unsigned tm = 0, nm = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            tm |= nibble_of(zero_bytes(w[q] ^ 0x09090909u)) << (4 * q);
            nm |= nibble_of(zero_bytes(w[q] ^ 0x0A0A0A0Au)) << (4 * q);
        }
        if (g < ls) { const unsigned keep = ~((1u << (ls - g)) - 1u); tm &= keep; nm &= keep; }   // bytes before the line (g + 16 > ls here)
        if (g + 16 <= ls) { tm = 0; nm = 0; }
        const int cnt = __popc(tm);
        int inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        const unsigned prev15 = __shfl_up_sync(0xffffffffu, tm >> 15, 1);
        const unsigned tin = lane == 0 ? carry : (prev15 & 1u);  // the byte before this lane's first byte is a tab
        const unsigned dbl = tm & ((tm << 1) | tin) & 0xFFFFu;   // a tab right after a tab: empty field
        const unsigned hitm = __ballot_sync(0xffffffffu, tabs + inc >= 9);
        if (hitm) {
            const int j = __ffs(hitm) - 1;
            const int before = __shfl_sync(0xffffffffu, tabs + inc - cnt, j);
            const unsigned tmj = __shfl_sync(0xffffffffu, tm, j);
            // position of the 9th tab inside lane j's 16 bytes: lane l (mod 16) asks whether byte l is a tab and the (9 - before)-th
            // one -- a vote instead of __fns (a loop of ~30 instructions)
            const int l16 = lane & 15;
            const unsigned nth = __ballot_sync(0xffffffffu, ((tmj >> l16) & 1u) != 0u && __popc(tmj & ((2u << l16) - 1u)) == 9 - before);
            const int bit = (__ffs(nth) - 1) & 15;
            const unsigned upto = (2u << bit) - 1u;
            const unsigned badm = lane < j ? (nm | dbl) : (lane == j ? ((nm | dbl) & upto) : 0u);
            if (__any_sync(0xffffffffu, badm != 0u)) return -1;
            const int s0 = base16 + 16 * j + bit + 1;
            if (s0 - ls > kMaxReq) return -1;
            // a FORMAT column spelled like a genotype ("0|1") would make the first sample look like a run continuation
            const uint32_t wl = __funnelshift_r(ldw(win, (s0 - 4) & ~3, r_lo, r_hi), ldw(win, ((s0 - 4) & ~3) + 4, r_lo, r_hi), 8 * ((s0 - 4) & 3));
            return ((wl & 0xFFFEFFFEu) ^ 0x09307C30u) == 0u ? -1 : s0;
        }
        if (__any_sync(0xffffffffu, (nm | dbl) != 0u)) return -1; // fewer than 10 columns / empty field
        tabs += __shfl_sync(0xffffffffu, inc, 31);
        carry = __shfl_sync(0xffffffffu, tm >> 15, 31) & 1u;
    }
    return -1;
}

// A term that is longer than parallel_portion's window (a multi-kilobyte literal), from its start `pos`: compress_data_line's
// sample loop (compress.cpp:124-190) term by term, by lane 0, until that term is out (*ended = 2: all lanes go on behind it),
// the line ends (*ended = 1) or the tile's end cut `ce` (a term start) is reached.  The warp stages the
// next kSerWin input bytes in the upper part of its staging area (lane 0 then walks shared memory, not global memory); output
// goes to the lower part and is flushed to the log when it fills up (both warp-collective, hence the outer loop).
// rc / rn: the run that is open at pos.  Returns the offset where the caller goes on, -1 for an empty term, -2 when the block
// has been given up meanwhile.
constexpr int kSerWin = 1024;                               // input window of the term walker
__device__ __noinline__ int serial_portion(const uint8_t* __restrict__ win, int pos, int ce, int r_hi, int rc, int rn,
                                           uint8_t* __restrict__ stage, int* o_io, int* flushed_io, uint8_t* __restrict__ log,
                                           Ctrl* __restrict__ ctrl, unsigned long long log_cap, int lane, unsigned long long* seg_first,
                                           unsigned long long* seg_prev, bool* dead, int* nl_seg, int my_off, int my_off2, int* ended) {
    constexpr int kOutCap = kSStage - kSerWin - 16;         // output part of the staging area while the walker runs
    uint8_t* const ibuf = stage + (kSStage - kSerWin);      // 16-byte aligned (kSStage and kSerWin are multiples of 16)
    int o = *o_io, flushed = *flushed_io;
    if (o - flushed > kOutCap - 8) {                        // what the grid steps staged does not leave room: flush first
        flush_segment(stage, o - flushed, log, ctrl, log_cap, lane, seg_first, seg_prev, dead, *nl_seg, my_off, my_off2);
        flushed = o; *nl_seg = 0;
    }
    int lit_src = 0, lit_rem = 0, lit_tab = 0;             // a literal that is being copied, a tab owed behind it
    int result = 0;                                         // 0: running / portion done; > 0: line done; -1: error
    bool done = false, line_end = false, stop_after = false, want_all = false;
    int n_odd = 0;                                          // odd-width terms walked by this call
    while (!done) {
        int giveup = 0;                                     // the block is going to the generic kernels anyway (one lane looks: uniform)
        if (lane == 0) giveup = *((volatile int*)&ctrl->irregular);
        if (__shfl_sync(0xffffffffu, giveup, 0)) { result = -2; break; }
        // the window [w0, w0 + kSerWin) around the walker's position; bytes behind the input read as '\n' (EOF ends the line)
        const int at = lit_rem > 0 ? lit_src : pos;
        const int w0 = at & ~15;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kSerWin / 512; q++) {
            const int r = w0 + 512 * q + 16 * lane;
            uint4 v;
            if (r + 16 <= r_hi) v = *reinterpret_cast<const uint4*>(win + r);
            else {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; j++) w[j] = ldw_edge(win, r + 4 * j, 0, r_hi, (uint32_t)'\n');
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
            *reinterpret_cast<uint4*>(ibuf + 512 * q + 16 * lane) = v;
        }
        __syncwarp();
        const int w1 = w0 + kSerWin;
        int fill = o - flushed;
        bool refill = false;
        if (lane == 0) {
            while (!done && !refill && fill <= kOutCap - 8) {
                if (lit_rem > 0) {
                    int n = min(lit_rem, kOutCap - 8 - fill + 1);
                    n = min(n, w1 - lit_src);
                    if (n <= 0) { refill = true; break; }
                    for (int i = 0; i < n; i++) stage[fill + i] = ibuf[lit_src - w0 + i];
                    fill += n; lit_src += n; lit_rem -= n;
                    continue;
                }
                if (lit_tab) { stage[fill++] = '\t'; lit_tab = 0; continue; }
                if (stop_after) { result = pos; done = true; break; }
                if (n_odd >= 1 && rc < 0 && !line_end && pos < ce) { result = pos; want_all = true; done = true; break; }   // the long term is out: all lanes again
                if (line_end) {
                    if (rc >= 0) stage[fill++] = (uint8_t)(cls_flag(rc) | (uint32_t)rn);
                    stage[fill++] = '\n';
                    done = true;
                    break;
                }
                if (pos >= ce) { result = ce; done = true; break; }   // the tile ends inside the line (behind an odd term: no run is open)
                int e = pos;
                uint32_t ch = 0;
                if (pos - w0 < 16) {                         // a term at the window's start may be longer than the window: global reads behind it
                    for (;; e++) { ch = e >= r_hi ? (uint32_t)'\n' : (e < w1 ? (uint32_t)ibuf[e - w0] : (uint32_t)win[e]); if (ch == '\t' || ch == '\n') break; }
                } else {
                    for (; e < w1; e++) { ch = ibuf[e - w0]; if (ch == '\t' || ch == '\n') break; }
                    if (e >= w1) { refill = true; break; }   // the term's end is not in the window: move the window to the term
                }
                const int len = e - pos;
                const bool last = ch == '\n';
                if (len == 0) { result = -1; done = true; break; }
                int c = 4;
                if (len == 3) c = gt_class3(ibuf + (pos - w0));   // (3 bytes from the window: pos + 3 = e < w1 or pos - w0 < 16)
                if (rc >= 0 && (c != rc || rn == (rc == 0 ? 127 : 31))) {
                    stage[fill++] = (uint8_t)(cls_flag(rc) | (uint32_t)rn);
                    rc = -1;
                }
                if (c == 4) {
                    stage[fill++] = (uint8_t)(kTokLit | 1u);
                    lit_src = pos; lit_rem = len; lit_tab = last ? 0 : 1;
                    if (len != 3) n_odd++;
                } else if (rc < 0) {
                    rc = c; rn = 1;
                } else {
                    rn++;
                }
                pos = e + 1;
                if (last) { line_end = true; result = pos; }
                else if (c == 4 && len != 3 && pos < ce && pos + 4 <= w1) {
                    // behind an odd-width term: when the next term is 3 bytes wide the grid takes over again (once the literal is out)
                    const uint8_t* nx = ibuf + (pos - w0);
                    if (!is_sep(nx[0]) && !is_sep(nx[1]) && !is_sep(nx[2]) && is_sep(nx[3])) stop_after = true;
                }
            }
        }
        done = __shfl_sync(0xffffffffu, (int)done, 0) != 0;
        fill = __shfl_sync(0xffffffffu, fill, 0);
        lit_rem = __shfl_sync(0xffffffffu, lit_rem, 0);
        lit_src = __shfl_sync(0xffffffffu, lit_src, 0);
        pos = __shfl_sync(0xffffffffu, pos, 0);
        o = flushed + fill;
        if (!done && fill > kOutCap - 8) {                   // the output part is full: flush and go on
            flush_segment(stage, fill, log, ctrl, log_cap, lane, seg_first, seg_prev, dead, *nl_seg, my_off, my_off2);
            flushed = o; *nl_seg = 0;
        }
    }
    result = __shfl_sync(0xffffffffu, result, 0);
    *ended = __shfl_sync(0xffffffffu, want_all ? 2 : (int)line_end, 0);      // 2: odd terms go on: parallel_portion takes the rest
    *o_io = o; *flushed_io = flushed;
    return result;
}

// ---- many odd-width terms in a row (a GT:DP:GQ file, haploid calls): the rest of the line portion, ALL LANES walking ------------
// A window of 2 KB (32 blocks of 64 bytes, one per lane): every lane turns its block into a 64-bit separator mask (word-parallel
// byte tests) and owns the terms that START in its block; a term's end is the next set bit of the mask (its own, or the first
// one of the lanes behind it).  Every lane walks its terms twice with the sample loop of compress_data_line
// (compress.cpp:124-190): once to size them and to learn the run it begins with / ends with, once to write.  In between, the
// open run is handed from lane to lane (a lane that is one single run hands on what it got, extended), and what the incoming run
// costs a lane is closed form: a closing token when the lane's first term differs, chunk-fill tokens every 127 / 31 samples of
// a leading run that continues it.
struct LaneWalk {                 // one lane's terms, walked without an incoming run
    int lead_c, lead_n;           // the leading run: class (-1: the range begins with a literal or is empty), samples
    bool lead_closed;             // a later term of the range ends it (else the whole range is that one run)
    int rc, rn;                   // the run that is open behind the range when it is not the leading one (-1: none)
    int sz;                       // bytes behind the leading run's tokens
    int nterms, err, n_odd;       // terms, empty terms, terms that are not 3 bytes wide
};

__device__ __forceinline__ unsigned long long bits_from(int k) {     // mask of the bit positions >= k (k may be < 0 or > 63)
    return k <= 0 ? ~0ull : (k >= 64 ? 0ull : (~0ull << k));
}

// T: the block's term starts, S: its separators (bit i = byte B + i); nxt: first separator behind the block.
// WRITE: emits into d (the caller has put the leading run's tokens there already); else sizes only.
template <bool WRITE>
__device__ __forceinline__ LaneWalk lane_walk(const uint8_t* __restrict__ win, int B, unsigned long long T, unsigned long long S, int nxt,
                                              int p_end, bool ended, uint8_t* d) {
    LaneWalk w = {-1, 0, false, -1, 0, 0, 0, 0, 0};
    bool in_lead = true;
    int o = 0;
    int i_next = T ? __ffsll((long long)T) - 1 : -1;
    while (i_next >= 0) {
        const int i = i_next;
        T &= T - 1ull;
        i_next = T ? __ffsll((long long)T) - 1 : -1;
        const int p = B + i;
        int e;
        if (i_next >= 0) e = B + i_next - 1;                   // the byte in front of the block's next term start is this term's separator
        else { const unsigned long long sm = S >> i; e = sm ? p + __ffsll((long long)sm) - 1 : nxt; }
        const int len = e - p;
        if (len == 0) w.err = 1;
        if (len != 3) w.n_odd++;
        const int c = len == 3 ? gt_class3(win + p) : 4;
        const bool is_last = ended && e == p_end;              // the line's last term: its separator is the newline
        bool normal = true;
        if (in_lead) {
            if (c < 4 && (w.lead_c < 0 || c == w.lead_c)) { w.lead_c = c; w.lead_n++; normal = false; }
            else { in_lead = false; w.lead_closed = w.lead_n > 0; }
        }
        if (normal) {
            if (w.rc >= 0 && (c != w.rc || w.rn == (w.rc == 0 ? 127 : 31))) {
                if (WRITE) d[o] = (uint8_t)(cls_flag(w.rc) | (uint32_t)w.rn);
                o++;
                w.rc = -1;
            }
            if (c == 4) {
                if (WRITE) {
                    d[o] = (uint8_t)(kTokLit | 1u);
                    for (int k = 0; k < len; k++) d[o + 1 + k] = win[p + k];
                    if (!is_last) d[o + 1 + len] = '\t';
                }
                o += 1 + len + (is_last ? 0 : 1);
            } else if (w.rc < 0) {
                w.rc = c; w.rn = 1;
            } else {
                w.rn++;
            }
        }
        w.nterms++;
    }
    w.sz = o;
    return w;
}

// Returns the offset where the caller goes on (behind the line's newline: *ended_out = 1, or ce), -1 for an empty term, -2 when
// the term at pos is longer than the window (the term walker takes it).  *odd_out: terms of the portion that are not 3 bytes wide.
__device__ __noinline__ int parallel_portion(const uint8_t* __restrict__ win, int pos, int ce, int r_hi, uint8_t* __restrict__ stage,
                                             int* o_io, int* flushed_io, uint8_t* __restrict__ log, Ctrl* __restrict__ ctrl,
                                             unsigned long long log_cap, int lane, unsigned long long* seg_first,
                                             unsigned long long* seg_prev, bool* dead, int* nl_seg, int my_off, int my_off2, int* ended_out,
                                             int* run_c, int* run_n, int* odd_out) {
    // 1. separator masks of the window's 32 blocks; bytes behind the input read as '\n' (EOF ends the line)
    constexpr int kParWin = 2048;
    const int wstart = pos & ~63, wend = wstart + kParWin, B = wstart + 64 * lane;
    uint32_t s_lo = 0, s_hi = 0, nl_any = 0;
    if (B + 64 <= r_hi) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 v = *reinterpret_cast<const uint4*>(win + B + 16 * q);
            const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
            uint32_t nib = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t zn = zero_bytes(wv[j] ^ 0x0A0A0A0Au);
                nl_any |= zn;
                nib |= nibble_of(zero_bytes(wv[j] ^ 0x09090909u) | zn) << (4 * j);
            }
            if (q < 2) s_lo |= nib << (16 * q); else s_hi |= nib << (16 * (q - 2));
        }
    } else {
#pragma unroll 1
        for (int q = 0; q < 16; q++) {
            const uint32_t wv = ldw_edge(win, B + 4 * q, 0, r_hi, 0x0A0A0A0Au);
            const uint32_t zn = zero_bytes(wv ^ 0x0A0A0A0Au);
            nl_any |= zn;
            const uint32_t nib = nibble_of(zero_bytes(wv ^ 0x09090909u) | zn);
            if (q < 8) s_lo |= nib << (4 * q); else s_hi |= nib << (4 * (q - 8));
        }
    }
    const unsigned long long S = (unsigned long long)s_lo | ((unsigned long long)s_hi << 32);
    const unsigned long long Sp = S & bits_from(pos - B);      // separators at or behind pos
    // 2. the portion's end: the line's newline (or the end of the input) when it lies in front of ce, else ce, else the window's
    //    last separator (the caller calls again with the run that is open there: *run_c, *run_n)
    int my_nl = -1;
    if (nl_any) {
        for (unsigned long long c = Sp; c; c &= c - 1ull) {
            const int g = B + __ffsll((long long)c) - 1;
            if (g >= r_hi || win[g] == '\n') { my_nl = g; break; }
        }
    }
    const unsigned nlm = __ballot_sync(0xffffffffu, my_nl >= 0);
    const int p_nl = nlm ? __shfl_sync(0xffffffffu, my_nl, __ffs(nlm) - 1) : -1;
    bool ended;
    int p_end;                                               // index of the separator behind the portion's last term
    if (p_nl >= 0 && (p_nl < ce || ce >= r_hi)) { ended = true; p_end = p_nl; }
    else if (ce <= wend) {
        if (ce >= r_hi) { ended = true; p_end = r_hi; }      // the input ends without a newline: EOF ends the line
        else { ended = false; p_end = ce - 1; }              // (ce is a term start: the byte in front of it is a tab)
    } else {
        const unsigned hm = __ballot_sync(0xffffffffu, Sp != 0ull);
        if (!hm) return -2;                                  // no separator in the window: one long term
        const int lt = B + 63 - __clzll((long long)Sp);
        ended = false;
        p_end = __shfl_sync(0xffffffffu, lt, 31 - __clz(hm));
    }
    // 3. term starts of the block: the byte behind a separator, and pos itself; the first separator behind the block
    const unsigned prev_hi = __shfl_up_sync(0xffffffffu, s_hi >> 31, 1);
    unsigned long long T = ((S << 1) | (unsigned long long)(lane > 0 ? (prev_hi & 1u) : 0u)) & bits_from(pos - B);
    if (pos >= B && pos < B + 64) T |= 1ull << (pos - B);
    T &= ~bits_from(p_end + 1 - B);
    // where a portion of `total` bytes goes: the staging area when it fits (what is staged is flushed first if need be), else a
    // log segment of its own; nullptr when the log is full (the caller goes on counting)
    int o = *o_io, flushed = *flushed_io;
    auto dest = [&](int total, bool& own) -> uint8_t* {
        if (o - flushed + total > kSStage && (o > flushed || *nl_seg > 0)) {
            flush_segment(stage, o - flushed, log, ctrl, log_cap, lane, seg_first, seg_prev, dead, *nl_seg, my_off, my_off2);
            flushed = o; *nl_seg = 0;
        }
        uint8_t* dst = stage + (o - flushed);
        own = false;
        if (o - flushed + total > kSStage) {                     // (the staging area is empty now)
            own = true;
            const int body = (total + 15) & ~15;
            const unsigned long long need = 16ull + (unsigned long long)body;
            unsigned long long lp = 0ull;
            if (lane == 0 && !*dead) {
                lp = atomicAdd(&ctrl->log_cursor, need);
                if (lp + need > log_cap) { atomicExch(&ctrl->cap_exceeded, 1); lp = ~0ull; }
            }
            lp = __shfl_sync(0xffffffffu, lp, 0);
            if (*dead || lp == ~0ull) { *dead = true; dst = nullptr; }
            else {
                dst = log + lp + 16;
                if (lane == 0) {
                    *reinterpret_cast<uint4*>(log + lp) = make_uint4((unsigned)total, 0u, 0u, 0u);
                    if (*seg_prev) *reinterpret_cast<unsigned long long*>(log + (*seg_prev - 1ull) + 8) = lp;
                }
                if (!*seg_first) *seg_first = lp + 1ull;
                *seg_prev = lp + 1ull;
            }
        }
        return dst;
    };
    // 3b. a window of literals only -- no term is exactly 3 bytes wide (GT:DP:GQ columns, haploid calls): no run logic, one pass.
    //     The output is the input with a literal token in front of every term (compress.cpp:171-185), so every lane copies the
    //     bytes of its BLOCK, shifted by the term starts in front of them.
    {
        unsigned nx_lo = __shfl_down_sync(0xffffffffu, s_lo, 1);
        if (lane == 31) nx_lo = 0u;                          // (a term that starts in the window's last bytes ends in front of p_end: shorter than 3)
        const unsigned long long S1 = (S >> 1) | ((unsigned long long)(nx_lo & 1u) << 63), S2 = (S >> 2) | ((unsigned long long)(nx_lo & 3u) << 62),
                                 S3 = (S >> 3) | ((unsigned long long)(nx_lo & 7u) << 61);
        if (__any_sync(0xffffffffu, (T & S) != 0ull)) return -1;                     // an empty term
        if (!__any_sync(0xffffffffu, (T & S3 & ~S2 & ~S1) != 0ull)) {
            const int nT = __popcll(T);
            int incl = nT;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, dd); if (lane >= dd) incl += t; }
            const int nterms = __shfl_sync(0xffffffffu, incl, 31);
            const bool eof_nl = ended && p_end >= r_hi;      // the line ends with the input: its '\n' is not among the bytes
            const int last = eof_nl ? r_hi - 1 : p_end;
            const int pre = *run_c >= 0 ? 1 : 0;             // the first term closes the run that is open in front of the portion
            const int total = pre + (last - pos + 1) + nterms + (eof_nl ? 1 : 0);
            bool own = false;
            uint8_t* const dst = dest(total, own);
            if (dst) {
                if (lane == 0 && pre) dst[0] = (uint8_t)(cls_flag(*run_c) | (uint32_t)*run_n);
                int g = max(B, pos);
                const int hi = min(B + 63, last);
                int shift = pre + (incl - nT) - pos;         // dst index of byte g = g + shift + (term starts of this block up to g)
                unsigned long long t = T;
                while (g <= hi) {
                    const int nxs = t ? B + __ffsll((long long)t) - 1 : hi + 1;      // the block's next term start
                    const int e = min(nxs, hi + 1);
                    for (; g < e; g++) dst[g + shift] = win[g];     // (four bytes per trip with shared-space stores: measured 13 % slower)
                    if (nxs <= hi) { dst[nxs + shift] = (uint8_t)(kTokLit | 1u); shift++; t &= t - 1ull; }
                }
                if (eof_nl && lane == 31) dst[total - 1] = '\n';
            }
            __syncwarp();
            o += total;
            if (own) flushed = o;
            *o_io = o; *flushed_io = flushed;
            *ended_out = ended ? 1 : 0;
            *run_c = -1; *run_n = 0;
            *odd_out = nterms;
            return p_end + 1;
        }
    }
    // the first separator behind the block (a term that starts in the block may end there)
    int nxt = S ? B + __ffsll((long long)S) - 1 : 0x7fffffff;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) { const int t = __shfl_down_sync(0xffffffffu, nxt, dd); if (lane + dd < 32) nxt = min(nxt, t); }
    nxt = __shfl_down_sync(0xffffffffu, nxt, 1);
    if (lane == 31 || nxt > p_end) nxt = p_end;
    // first walk: sizes and runs
    const LaneWalk w = lane_walk<false>(win, B, T, S, nxt, p_end, ended, nullptr);
    if (__any_sync(0xffffffffu, w.err != 0)) return -1;
    *odd_out = __reduce_add_sync(0xffffffffu, w.n_odd);
    // 4. the open run, from lane to lane
    int in_c = *run_c, in_n = *run_n, my_c = -1, my_n = 0;      // (lane 0: the run that is open in front of the portion)
    if (lane != 0) { in_c = -1; in_n = 0; }
    const bool one_run = !w.lead_closed && w.lead_c >= 0 && w.rc < 0 && w.lead_n == w.nterms;   // the whole range is one run
    if (!__any_sync(0xffffffffu, w.nterms == 0 || one_run)) {   // every lane leaves a run of its own making (or none) behind
        my_c = w.rc; my_n = w.rn;
        const int oc = __shfl_up_sync(0xffffffffu, my_c, 1), on = __shfl_up_sync(0xffffffffu, my_n, 1);
        if (lane > 0) { in_c = oc; in_n = on; }
    } else {
        for (int L = 0; L < 32; L++) {
            if (w.nterms == 0) { my_c = in_c; my_n = in_n; }
            else if (one_run) {
                const int m = w.lead_c == 0 ? 127 : 31;
                my_c = w.lead_c;
                my_n = (((in_c == w.lead_c ? in_n : 0) + w.lead_n - 1) % m) + 1;
            } else { my_c = w.rc; my_n = w.rn; }
            const int oc = __shfl_sync(0xffffffffu, my_c, L), on = __shfl_sync(0xffffffffu, my_n, L);
            if (lane == L + 1) { in_c = oc; in_n = on; }
        }
    }
    const int fin_c = __shfl_sync(0xffffffffu, my_c, 31), fin_n = __shfl_sync(0xffffffffu, my_n, 31);
    // 5. what the incoming run adds in front of the range
    const bool merged = w.nterms > 0 && w.lead_n > 0 && in_c == w.lead_c;
    const int c0 = merged ? in_n : 0;
    const int m_lead = w.lead_c == 0 ? 127 : 31;
    int pre = 0, fills = 0;
    if (w.nterms > 0) {
        if (in_c >= 0 && !merged) pre += 1;                  // my first term ends the incoming run
        if (w.lead_n > 0) { fills = (c0 + w.lead_n - 1) / m_lead; pre += fills + (w.lead_closed ? 1 : 0); }
    }
    const int mine = pre + w.sz;
    int inc = mine;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, dd); if (lane >= dd) inc += t; }
    int total = __shfl_sync(0xffffffffu, inc, 31);
    const int tail = ended ? (fin_c >= 0 ? 2 : 1) : 0;       // the line's last run token and its newline
    total += tail;
    // 6. where it goes
    bool own = false;
    uint8_t* dst = dest(total, own);
    // 7. second walk: write
    if (dst) {
        uint8_t* d = dst + (inc - mine);
        int k = 0;
        if (w.nterms > 0) {
            if (in_c >= 0 && !merged) d[k++] = (uint8_t)(cls_flag(in_c) | (uint32_t)in_n);
            if (w.lead_n > 0) {
                for (int f = 0; f < fills; f++) d[k++] = (uint8_t)(cls_flag(w.lead_c) | (uint32_t)m_lead);
                if (w.lead_closed) d[k++] = (uint8_t)(cls_flag(w.lead_c) | (uint32_t)(((c0 + w.lead_n - 1) % m_lead) + 1));
            }
        }
        lane_walk<true>(win, B, T, S, nxt, p_end, ended, d + k);
        if (ended && lane == 31) {
            uint8_t* t = dst + (total - tail);
            if (fin_c >= 0) *t++ = (uint8_t)(cls_flag(fin_c) | (uint32_t)fin_n);
            *t = '\n';
        }
    }
    __syncwarp();
    o += total;
    if (own) flushed = o;
    *o_io = o; *flushed_io = flushed;
    *ended_out = ended ? 1 : 0;
    *run_c = ended ? -1 : fin_c; *run_n = ended ? 0 : fin_n;
    return p_end + 1;
}

struct OddOut { int cur, o, flushed, nl_seg, ein0, flags, ein_carry; };
enum { kOddLineEnd = 1, kOddEmpty = 2, kOddGiveUp = 4, kOddLookedBack = 16, kOddRunOpen = 32, kOddParMode = 64 };
constexpr int kOddDense = 8;      // odd-width terms per 2 KB window from which on the walkers keep the line (and take the next one at once)
// A step [a, ...) that holds a sample column that is not 3 bytes + separator: all lanes walk the window that begins at a
// (parallel_portion; the samples in front of the odd one included), and the windows behind it while they are dense in odd
// terms; then the grid goes on, in whatever phase that is, with the run that is open there.  A term that is longer than the
// window goes to the term walker (serial_portion, lane 0).
// direct: the previous line ended in a dense window -- this one (a line start: no run is open) comes here without the grid's attempt.
__device__ __noinline__ OddOut odd_step(const uint8_t* __restrict__ win, int a, int ce, int cs, int r_lo, int r_hi, bool first, bool need_lb,
                                        int ein_carry, int pc0, int tile, const unsigned int* s1, uint8_t* __restrict__ stage, int o,
                                        int flushed, int nl_seg, int my_off, int my_off2, uint8_t* __restrict__ log, Ctrl* __restrict__ ctrl,
                                        unsigned long long log_cap, unsigned long long serial_budget, int lane,
                                        unsigned long long* seg_first, unsigned long long* seg_prev, bool* dead, bool direct) {
    OddOut out = {a, o, flushed, nl_seg, 0, 0, kNoHead};
    int rcl = -1, rcn = 0;                               // open run: class, samples in its open chunk (1..M)
    if (!direct && !first) {
        if (need_lb) {                                   // (the tile's first step: the entering run's chunk count)
            int cnt_in = 0;
            if (lane == 0) cnt_in = lookback_count(s1, tile, pc0);
            cnt_in = __shfl_sync(0xffffffffu, cnt_in, 0);
            ein_carry = cs - 4 * cnt_in;
            out.ein0 = ein_carry; out.flags |= kOddLookedBack;
        }
        const int pcl = win[a - 5] == '\t' ? gt_class3(win + a - 4) : 4;
        if (pcl < 4 && ein_carry != kNoHead) {
            rcl = pcl;
            rcn = (int)mod_chunk(((a - ein_carry) >> 2) - 1, pcl == 0) + 1;
        }
    }
    if (lane == 0) ctrl->odd_used = 1;
    int at = a, ended = 0, n_odd = 0;
    while (!ended && at < ce) {
        const int nx = parallel_portion(win, at, ce, r_hi, stage, &out.o, &out.flushed, log, ctrl, log_cap, lane, seg_first, seg_prev, dead,
                                        &out.nl_seg, my_off, my_off2, &ended, &rcl, &rcn, &n_odd);
        if (nx == -1) { out.flags |= kOddEmpty; return out; }
        if (nx == -2) {
            // the term at `at` is longer than the window: the term walker writes it (with the run that is open in front of it)
            const int r = serial_portion(win, at, ce, r_hi, rcl, rcn, stage, &out.o, &out.flushed, log, ctrl, log_cap, lane, seg_first, seg_prev,
                                         dead, &out.nl_seg, my_off, my_off2, &ended);
            if (r < 0) { out.flags |= r == -1 ? kOddEmpty : kOddGiveUp; return out; }   // an empty sample column / the block is given up
            int over = 0;
            if (lane == 0 && atomicAdd(&ctrl->serial_bytes, (unsigned long long)(r - at)) > serial_budget) over = 1;
            if (__shfl_sync(0xffffffffu, over, 0)) { out.flags |= kOddGiveUp; return out; }   // (a safety valve: generic kernels)
            at = r;
            rcl = -1; rcn = 0;                           // (the walker stops behind an odd term or at the line's end: no run is open)
            n_odd = kOddDense;
            ended = ended == 1 ? 1 : 0;
            continue;
        }
        at = nx;
        if (n_odd < kOddDense) break;                    // few odd-width terms: the grid takes over again
    }
    if (n_odd >= kOddDense) out.flags |= kOddParMode;
    if (!ended && at < ce && rcl >= 0) { out.flags |= kOddRunOpen; out.ein_carry = at - 4 * rcn; }   // (a virtual head: rcn samples back)
    out.cur = at;                                        // (at ce with a run open: the next tile's look-back record has it)
    if (ended) out.flags |= kOddLineEnd;
    return out;
}

// The call of odd_step with everything read from / written back to the warp's OddSave (see there).
__device__ __noinline__ void odd_call(OddSave* S, const uint8_t* __restrict__ in, long long n, int tile_sz, const unsigned int* s1,
                                      uint8_t* __restrict__ stage, uint8_t* __restrict__ log, Ctrl* __restrict__ ctrl,
                                      unsigned long long log_cap, unsigned long long serial_budget, int lane,
                                      unsigned long long* seg_first, unsigned long long* seg_prev, bool* dead) {
    const int tile = S->tile;
    const long long gb = (long long)tile * tile_sz - 64;
    const uint8_t* const win = in + gb;
    const int r_lo = gb < 0 ? 64 : 0;
    const int r_hi = (int)(n - gb < (long long)(1 << 24) ? n - gb : (long long)(1 << 24));
    const int a = S->odd_a;
    const int bits = S->bits;
    const OddOut r = odd_step(win, a, S->ce, S->cs, r_lo, r_hi, (bits & 2) != 0, (bits & 4) != 0, S->ein_carry, S->pc0, tile, s1, stage, S->o,
                              S->flushed, S->nl_seg, S->my_off[lane], S->my_off2[lane], log, ctrl, log_cap, serial_budget, lane, seg_first, seg_prev,
                              dead, S->odd_direct != 0);
    __syncwarp();
    if (lane == 0) {
        if (r.flags & (kOddEmpty | kOddGiveUp)) { S->irregular = (r.flags & kOddEmpty) ? 6 : 7; S->action = 2; }
        else {
            int b = bits;
            S->o = r.o; S->flushed = r.flushed; S->nl_seg = r.nl_seg; S->cur = r.cur;
            S->par_mode = (r.flags & kOddParMode) ? 1 : 0;
            if (r.flags & kOddLookedBack) { S->ein0 = r.ein0; b &= ~4; }
            if (r.flags & kOddLineEnd) b |= 1;                            // behind the line's newline
            if (r.flags & kOddRunOpen) { b &= ~2; S->ein_carry = r.ein_carry; }   // all lanes walked a stretch of 3-byte terms: its run is open
            else { b |= 2; S->ein_carry = kNoHead; }                      // (the walkers stop behind an odd term: no run is open)
            S->bits = b;
            S->action = 0;
        }
    }
}

// kOdd = false: the kernel of regular blocks -- a sample column that is not 3 bytes + separator gives the block up with reject
// reason kRejectOddTerms, and the host launches it again with kOdd = true (the term walkers compiled in; that instantiation is
// ~4 % slower on regular blocks, which is why there are two).  The context remembers which one its stream of blocks needs.
constexpr int kRejectOddTerms = 9;
#ifndef VCFC_ENC_SCTAS_ODD
#define VCFC_ENC_SCTAS_ODD VCFC_ENC_SCTAS
#endif
template <bool kOdd>
__global__ void __launch_bounds__(32 * kSWarps, kOdd ? VCFC_ENC_SCTAS_ODD : kSCtas)
k_encode_stream(const uint8_t* __restrict__ in, long long n, uint8_t* __restrict__ log, Ctrl* __restrict__ ctrl,
                unsigned int* __restrict__ s1, unsigned long long* __restrict__ rec_pos, unsigned long long* __restrict__ rec_size,
                unsigned long long* __restrict__ rec_lines, int n_tiles, unsigned long long log_cap, int tile_sz,
                unsigned long long serial_budget) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    using Smem = typename std::conditional<kOdd, SmemSO, SmemS>::type;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* const stage = sm.stage[warp];
#if !VCFC_ENC_TICKET
    const int gw = (int)blockIdx.x * kSWarps + warp, nw = (int)gridDim.x * kSWarps;
#endif
    const size_t lane64 = 64u * (size_t)lane;
    int irr_seen = 0;                       // ctrl->irregular as of one tile ago (the load stays off the critical path)
    if constexpr (kOdd) { if (lane == 0) sm.odd[warp].par_mode = 0; __syncwarp(); }
    // The warp runs one tile AHEAD with the end cut and the look-back #1 record: iteration k publishes the record of
    // this warp's tile k+1 and then encodes tile k, so a record is there a whole tile time before its reader needs it.
    constexpr int kFlSkip = 1 << 9;
    int cur_ce = 0, cur_fl = 0;             // the tile to encode: end cut (relative to its own base), cut kind | class << 4 | uniform << 8

#if VCFC_ENC_TICKET
    // Tiles are handed out in increasing order by one atomic counter: a dense tile costs several times a sparse one, and a
    // fixed round-robin leaves the warps that drew the cheap tiles idle at the end.  Whoever holds ticket t publishes tile t's
    // look-back record before it waits on anything, so a reader of s1[t - 1] never waits on a warp that has not started.
    int tile = -1, nt = 0, nn = 0;
    if (lane == 0) nt = (int)atomicAdd(&ctrl->ticket, 1u);
    nt = __shfl_sync(0xffffffffu, nt, 0);
    for (; tile < n_tiles; tile = nt, nt = __shfl_sync(0xffffffffu, nn, 0)) {
        int n_ce = 0, n_fl = 0;
        nn = n_tiles;                        // the ticket after nt: asked for now, needed when the iteration ends
        if (lane == 0 && nt < n_tiles) nn = (int)atomicAdd(&ctrl->ticket, 1u);
        if (nt < n_tiles) {
#else
    for (int tile = gw - nw; tile < n_tiles; tile += nw) {
        // ---- ahead: end cut of the next tile and its look-back #1 record -- what the run that leaves the tile looks
        //      like.  A sample is a run head unless it and the word before it are the same coded genotype (line_scan16
        //      guarantees that a line's first sample never passes this test), so the last head is found from the bytes
        //      alone, scanning back from the end cut --------------------------------------------------------------------
        int n_ce = 0, n_fl = 0;
        if (tile + nw < n_tiles) {
            const int nt = tile + nw;
#endif
            const int irr_now = irr_seen;
            irr_seen = *((volatile int*)&ctrl->irregular);
            unsigned word = (2u << 30) | ((unsigned)kNone << 8);            // nothing carried out
            if (irr_now) {                  // some tile already gave up: keep the look-back chains alive and move on
                n_fl = kFlSkip;
            } else {
                const long long t0 = (long long)nt * tile_sz, gb = t0 - 64, t1 = t0 + tile_sz;
                const uint8_t* const win = in + gb;
                const int r_lo = gb < 0 ? 64 : 0;
                const int r_hi = (int)(n - gb < (long long)(1 << 24) ? n - gb : (long long)(1 << 24));
                const long long vlo_e = t1 - kHalo > 0 ? t1 - kHalo : 0, vhi_e = t1 + kHalo < n ? t1 + kHalo : n;
                int ke;
                const int ce = (int)(cut_find(in, 0, vlo_e, vhi_e, n, t1, lane, &ke) - gb);
                int lc = kNone, uniform = 0;
                if (ke == kCutSample) {
                    const int lo = 64 > r_lo + 4 ? 64 : r_lo + 4;             // sample starts below t0 belong to the previous tile
                    lc = win[ce - 5] == '\t' ? gt_class3(win + ce - 4) : 4;      // "10|0" ends like "0|0": the term must be 3 bytes
                    word = (2u << 30) | ((unsigned)lc << 8);
                    if (lc < 4) {
                        int found = -1;
                        bool coarse = lc == 0;                                    // a "0|0" run may reach back for kilobytes
                        for (int top = ce - 4; top >= lo && found < 0;) {
                            const int p = top - 4 * lane;
                            bool head = false;
                            if (p >= lo) {
                                const int pa = p & ~3, shq = 8 * (p & 3);
                                const uint32_t wa = ldw(win, pa - 4, r_lo, r_hi), wb = ldw(win, pa, r_lo, r_hi), wc = ldw(win, pa + 4, r_lo, r_hi);
                                const uint32_t w1 = __funnelshift_r(wb, wc, shq), w0 = __funnelshift_r(wa, wb, shq);
                                head = !((((w1 & 0xFFFEFFFEu) ^ 0x09307C30u) == 0u) && w1 == w0);
                            }
                            const unsigned hm = __ballot_sync(0xffffffffu, head);
                            if (hm) { found = top - 4 * (__ffs(hm) - 1); break; }
                            top -= 128;
                            if (coarse && top >= lo) {
                                // 32 samples of "0|0" so far: go on in groups of four samples (16 bytes on the run's own 4-byte grid),
                                // 2 KB per round trip, down to the first group that is not "0|0\t0|0\t0|0\t0|0\t" or reaches below
                                // the tile; the head is in that group or right behind it, and the sample-wise search resumes there
                                coarse = false;
                                const int ph = ce & 3, sh8 = 8 * ph;
                                const int A = (top + 4 - ph) & ~15;               // group i = the four samples from A - 16 (i + 1) + ph
                                int hit = -1;
                                for (int ib = 0; hit < 0; ib += 32 * kBackRows) {
                                    uint4 v4[kBackRows];
                                    uint32_t nx4[kBackRows];
                                    bool val[kBackRows];
#pragma unroll
                                    for (int r = 0; r < kBackRows; r++) {
                                        const int U = A - 16 * (ib + 32 * r + lane + 1);
                                        val[r] = U >= r_lo && U + ph >= lo;
                                        v4[r] = make_uint4(0u, 0u, 0u, 0u);
                                        nx4[r] = 0u;
                                        if (val[r]) {
                                            v4[r] = *reinterpret_cast<const uint4*>(win + U);
                                            nx4[r] = *reinterpret_cast<const uint32_t*>(win + U + 16);
                                        }
                                    }
#pragma unroll
                                    for (int r = 0; r < kBackRows; r++) {
                                        const uint4 v = v4[r];
                                        const uint32_t dd = (__funnelshift_r(v.x, v.y, sh8) ^ 0x09307C30u) | (__funnelshift_r(v.y, v.z, sh8) ^ 0x09307C30u) |
                                                            (__funnelshift_r(v.z, v.w, sh8) ^ 0x09307C30u) | (__funnelshift_r(v.w, nx4[r], sh8) ^ 0x09307C30u);
                                        const unsigned E = __ballot_sync(0xffffffffu, !val[r] || dd != 0u);
                                        if (E) { hit = ib + 32 * r + __ffs(E) - 1; break; }
                                    }
                                }
                                top = min(top, A - 16 * hit + ph);
                            }
                        }
                        if (found >= 0) {
                            if (win[found - 1] != '\t') found += 4;               // the earliest word is the tail of a longer term
                            word |= (unsigned)mod_chunk(((ce - found) >> 2) - 1, lc == 0) + 1u;
                        } else {                                                 // the entering run covers the whole tile
                            uniform = 1;
                            word = (1u << 30) | ((unsigned)lc << 8) | (unsigned)mod_chunk((ce - 64) >> 2, lc == 0);   // relative
                        }
                    }
                }
                n_ce = ce;
                n_fl = ke | (lc << 4) | (uniform << 8);
            }
            if (lane == 0) *((volatile unsigned*)&s1[nt]) = word;
        }
        int ce = cur_ce;
        int fl = cur_fl;
        cur_ce = n_ce; cur_fl = n_fl;
        if (tile < 0 || (fl & kFlSkip)) continue;
        // ---- the tile itself ---------------------------------------------------------------------------------------------
        // (not const: the kOdd instantiation sets them again behind the call of odd_step, see OddSave)
        long long t0 = (long long)tile * tile_sz;
        long long gb = t0 - 64;                                         // tile-relative offsets: r = g - gb (a multiple of 64 apart)
        const uint8_t* win = in + gb;
        int r_lo = gb < 0 ? 64 : 0;                                     // valid relative range [r_lo, r_hi)
        int r_hi = (int)(n - gb < (long long)(1 << 24) ? n - gb : (long long)(1 << 24));
        int irregular = 0;
        int pf_lim = min(r_hi, ce + 127), pf_lim2 = min(r_hi, ce + 4095);   // how far the steps' prefetches may reach
        // the first 8 KB of the tile into L2 now; every step asks for the 2 KB that lie 8 KB ahead of it
        {
            const int pr = 64 + 128 * lane;
            if (pr + 128 <= r_hi) asm volatile("prefetch.global.L2 [%0];" ::"l"(win + pr));
            if (pr + 4096 + 128 <= r_hi) asm volatile("prefetch.global.L2 [%0];" ::"l"(win + pr + 4096));
        }
        int ke = fl & 15, lb_lc = (fl >> 4) & 15, lb_uniform = (fl >> 8) & 1, lb_nsamp = (ce - 64) >> 2;
        if (ke == kCutBad) irregular = 2;
        int ks, cs;
        {
            const long long vlo_s = t0 - kHalo > 0 ? t0 - kHalo : 0, vhi_s = t0 + kHalo < n ? t0 + kHalo : n;
            cs = (int)(cut_find(in, 0, vlo_s, vhi_s, n, t0, lane, &ks) - gb);
            if (ks == kCutBad) irregular = 2;
        }
        // ---- look-back #1, read: chunk count of the run that enters the tile -- resolved as late as possible (just before
        //      the first byte count), so the predecessor's record is normally there already -------------------------------
        int ein0 = kNoHead;
        int pc0 = kNone;
        bool prev_lit = false;                           // the tile starts inside a line, behind a term that is not a coded genotype
        if (ks == kCutSample && tile > 0 && cs < ce && !irregular) {
            pc0 = win[cs - 5] == '\t' ? gt_class3(win + cs - 4) : 4;
            prev_lit = pc0 == 4;
        }
        bool need_lb = pc0 < 4;
        // ---- the tile, line by line.  Output goes to the staging area; when the next piece (a line start: <= 8 + kMaxReq bytes,
        //      a step: <= 2.6 KB) would not fit, what is staged is flushed to the log as a SEGMENT: [u32 bytes, u32 0, u64 position
        //      of the tile's next segment][bytes, padded to 16].  A sparse tile is one segment; a dense one never runs twice ----
        int nl = 0, nl_seg = 0, my_off = 0, my_off2 = 0;  // line starts of the tile / of the segment being staged, their offsets
        int flushed = 0;                                  // tile bytes already in the log (o - flushed = staging fill)
        unsigned long long seg_first = 0ull, seg_prev = 0ull;
        bool dead = false;                                // the log is full: keep counting, write nothing more
        {
            int o = 0, cur = cs, ein_carry = ein0;
            bool in_req = ks == kCutLine, first = ks == kCutSampleFirst || prev_lit;    // first: no run is open before the next sample
            bool odd_pending = false;                        // a step that met an odd-width sample column (rare): see odd_step
            for (;;) {
            while (cur < ce && !irregular) {
                if (in_req) {
                    // ---- a line start: two length headers + the required section (compress.cpp:32-100) ----------------
                    const int ls = cur;
                    const int s0 = line_scan16(win, ls, r_lo, r_hi, lane);
                    if (s0 < 0) { irregular = 4; break; }
                    const int rq = s0 - ls;
                    if ((o - flushed + 8 + rq > kSStage || nl_seg >= kMaxNl) && (o > flushed || nl_seg > 0)) {
                        flush_segment(stage, o - flushed, log, ctrl, log_cap, lane, &seg_first, &seg_prev, &dead, nl_seg, my_off, my_off2);
                        flushed = o; nl_seg = 0;
                    }
                    if (8 + rq > kSStage) {
                        // a required section longer than the staging area (long REF / ALT / INFO): its own segment, input -> log
                        flush_line_start(win + ls, rq, o, log, ctrl, log_cap, lane, &seg_first, &seg_prev, &dead);
                        nl++;
                        o += 8 + rq;
                        flushed = o;
                        cur = s0;
                        in_req = false; first = true; ein_carry = kNoHead;
                        if constexpr (kOdd) {
                            if (sm.odd[warp].par_mode) {
                                if (lane == 0) { sm.odd[warp].odd_a = cur; sm.odd[warp].odd_direct = 1; }
                                odd_pending = true;
                                break;
                            }
                        }
                        continue;
                    }
                    {
                        uint8_t* d = stage + (o - flushed);
                        if (lane < 4) d[lane] = lane == 0 ? 0xC0 : 0;                       // line length: patched by k_patch_headers
                        if (lane >= 4 && lane < 8) {
                            const unsigned v = (unsigned)rq;
                            d[lane] = lane == 4 ? (uint8_t)((v >> 24) | 0xC0) : (uint8_t)(v >> (8 * (7 - lane)));
                        }
                        for (int k = lane; k < rq; k += 32) d[8 + k] = win[ls + k];
                    }
                    if (lane == (nl_seg & 31)) { if (nl_seg < 32) my_off = o; else my_off2 = o; }
                    nl++; nl_seg++;
                    o += 8 + rq;
                    cur = s0;
                    in_req = false; first = true; ein_carry = kNoHead;
                    if constexpr (kOdd) {
                        if (sm.odd[warp].par_mode) {
                                if (lane == 0) { sm.odd[warp].odd_a = cur; sm.odd[warp].odd_direct = 1; }
                                odd_pending = true;
                                break;
                            }
                    }
                    continue;
                }
                // ---- a step of samples: those that start in [cur, bound), up to the line's end ---------------------------
#if VCFC_ENC_PFW
                // (consecutive steps of a line run in a loop of their own: the words loaded ahead are live inside it only, and no
                //  call lies between the load and their use)
                bool w_ready = false;
                uint32_t W[18];                              // the step's words; loaded one step ahead when the next step is known
                for (;;) {
#endif
                const int a = cur, wstart = a & ~63, blk = wstart + 64 * lane, phase = a & 3, base = blk + phase;
                const int bound = min(ce, wstart + kStep);
#if !VCFC_ENC_PFW
                uint32_t W[18];
#endif
                const uint8_t* const pb = win + blk;
#if VCFC_ENC_PFW
                if (w_ready) { w_ready = false; } else
#endif
                if (wstart - 4 >= r_lo && wstart + kStep + 4 <= r_hi) {
                    // the whole step lies inside the input (warp-uniform, all but the first / last step of the block): every lane
                    // loads its block and the two words around it itself -- the neighbours' words are L1 hits, cheaper than two
                    // shuffles plus two divergent edge loads
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        uint4 v;
#if VCFC_ENC_LDHINT == 1
                        asm volatile("ld.global.L1::evict_first.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(pb + 16 * q));
#elif VCFC_ENC_LDHINT == 2
                        asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(pb + 16 * q));
#else
                        v = *reinterpret_cast<const uint4*>(pb + 16 * q);
#endif
                        W[4 * q + 1] = v.x; W[4 * q + 2] = v.y; W[4 * q + 3] = v.z; W[4 * q + 4] = v.w;
                    }
                    W[0] = *reinterpret_cast<const uint32_t*>(pb - 4);
                    W[17] = *reinterpret_cast<const uint32_t*>(pb + 64);
                } else {
                    uint32_t We[18];                       // (in local memory: the address escapes; W itself stays in registers)
                    step_load_edge(win, blk, r_lo, r_hi, We);
#pragma unroll
                    for (int q = 0; q < 18; q++) W[q] = We[q];
                }
                // the next step's block: a prefetch hint costs no registers and turns its loads into L1 hits; 8 KB ahead into L2
#if !VCFC_ENC_NOPF1
                if (blk + kStep + 64 <= pf_lim) asm volatile("prefetch.global.L1 [%0];" ::"l"(pb + kStep));
#endif
                if (lane < 16 && wstart + 8192 + 2048 <= pf_lim2) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + lane64 + 8192));
                const int sh = 8 * phase;
                const uint32_t sp = __funnelshift_r(W[0], W[1], sh);
                const uint32_t pc_all = ((sp & 0xFFFEFFFEu) ^ 0x09307C30u) == 0u ? 1u : 0u;
                // X[k] = sample word k ^ "0|0\t": zero but for the two allele bits (bit 0, bit 16) when the sample is "x|y\t" with
                // x, y in {0,1}.  When every sample of the step is (no literal, no line end, no bytes outside the line) the allele
                // masks are one sum of shifted words; otherwise per-sample tests, and the allele bits masked before they are summed
                uint32_t X[16];
#pragma unroll
                for (int k = 0; k < 16; k++) X[k] = __funnelshift_r(W[k + 1], W[k + 2], sh) ^ 0x09307C30u;
                uint32_t anyx = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) anyx |= X[k];
                uint32_t Craw = 0xFFFFu, acc = 0;              // acc = second-allele bits << 16 | first-allele bits
                if (!__any_sync(0xffffffffu, (anyx & 0xFFFEFFFEu) != 0u)) {
#pragma unroll
                    for (int k = 0; k < 16; k++) acc += X[k] << k;
                } else {
                    Craw = 0;
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        if ((X[k] & 0xFFFEFFFEu) == 0u) Craw |= 1u << k;
                        acc += (X[k] & 0x00010001u) << k;
                    }
                }
                Item it;
                it.base = base; it.Ap = ((acc & 0xFFFFu) << 1) | (sp & 1u); it.Bp = ((acc >> 16) << 1) | ((sp >> 16) & 1u);
                // samples that start in [a, bound)
                const int rel_a = a - base, rel_b = bound - base;
                const int klo = rel_a > 0 ? rel_a >> 2 : 0, khi = rel_b >= 64 ? 16 : (rel_b > 0 ? (rel_b + 3) >> 2 : 0);
                uint32_t V = khi > klo ? (((1u << khi) - 1u) & ~((1u << klo) - 1u)) : 0u;
                // everything valid that is not "x|y\t": literals, and the line's last sample, which carries the '\n' (the first
                // such sample ends the step).  Most steps of a sparse file have neither: one ballot skips all of it.
                int kend = -1, q_end = 0;
                unsigned endm = 0;
                int le = 32;
                uint32_t L = 0;
                const bool rare = __any_sync(0xffffffffu, (V & ~Craw) != 0u);
                if (rare) {
                    // in rounds, one candidate sample per lane and round: behind the line's end every sample looks odd, and those
                    // lanes must not walk all of theirs -- only lanes in front of the best newline found so far keep looking
                    for (uint32_t rem = V & ~Craw;;) {
                        const bool look = rem != 0u && lane < le;
                        if (!__any_sync(0xffffffffu, look)) break;
                        if (look) {
                            const int k = __ffs(rem) - 1;
                            rem &= rem - 1;
                            if (ldb_nl(win, base + 4 * k + 3, r_lo, r_hi) == '\n') { kend = k; rem = 0u; }
                        }
                        endm = __ballot_sync(0xffffffffu, kend >= 0);
                        if (endm) le = __ffs(endm) - 1;
                    }
                    if (lane > le) kend = -1;
                    if (lane > le) V = 0;
                    if (lane == le) V &= (2u << kend) - 1u;
                    q_end = __shfl_sync(0xffffffffu, base + 4 * kend + 3, le & 31);   // the '\n' (valid when endm)
                    bool irr = false;
                    for (uint32_t t = V & ~Craw; t; t &= t - 1) {
                        const int k = __ffs(t) - 1;
                        const int r0 = base + 4 * k;
                        uint32_t b0, b1, b2, b3;
                        if (r0 >= r_lo && r0 + 4 <= r_hi) {
                            const uint8_t* p = win + r0;
                            b0 = p[0]; b1 = p[1]; b2 = p[2]; b3 = p[3];
                        } else {                                  // the input's first / last bytes
                            b0 = ldb_nl(win, r0, r_lo, r_hi); b1 = ldb_nl(win, r0 + 1, r_lo, r_hi); b2 = ldb_nl(win, r0 + 2, r_lo, r_hi);
                            b3 = ldb_nl(win, r0 + 3, r_lo, r_hi);
                        }
                        if (k == kend ? (b3 != '\n') : (b3 != '\t')) irr = true;
                        if (b1 == '|' && (b0 & 0xFEu) == 0x30u && (b2 & 0xFEu) == 0x30u) {
                            Craw |= 1u << k;           // coded sample terminated by the line's newline
                        } else {
                            if (is_sep(b0) || is_sep(b1) || is_sep(b2)) irr = true;
                            L |= 1u << k;
                        }
                    }
                    if (__any_sync(0xffffffffu, irr)) {      // a sample column that is not 3 bytes + separator: handled behind the loop
                        if constexpr (!kOdd) { irregular = VCFC_ENC_NOSERIAL ? 6 : kRejectOddTerms; break; }
                        else {
                            if (lane == 0) { sm.odd[warp].odd_a = a; sm.odd[warp].odd_direct = 0; }
                            odd_pending = true;
                            break;
                        }
                    }
                }
                const uint32_t F = (first && lane == 0) ? (1u << klo) : 0u;
                const uint32_t Cprev = (Craw << 1) | pc_all;
                const uint32_t same = ~(((it.Ap >> 1) ^ it.Ap) | ((it.Bp >> 1) ^ it.Bp));   // same genotype bits as the previous word
                const uint32_t Hd = V & (F | ~(Craw & Cprev & same));
                const uint32_t PC = (Cprev & 0xFFFFu) & ~F;
                it.V = V; it.C = Craw & V; it.L = L; it.Hd = Hd; it.CL = Hd & PC;
                it.kend = (lane == le) ? kend : -1;
                it.pcoded = (int)((PC >> klo) & 1u);
                const int lh = Hd ? base + 4 * (31 - __clz(Hd)) : kNoHead;
                if (need_lb) {
                    int cnt_in = 0;
                    if (lane == 0) {
                        cnt_in = lookback_count(s1, tile, pc0);                      // open chunk count before the tile, 1..M
                        if (lb_uniform)
                            *((volatile unsigned*)&s1[tile]) =
                                (2u << 30) | ((unsigned)lb_lc << 8) | ((unsigned)mod_chunk(cnt_in + lb_nsamp - 1, lb_lc == 0) + 1u);
                    }
                    cnt_in = __shfl_sync(0xffffffffu, cnt_in, 0);
                    ein0 = cs - 4 * cnt_in;
                    ein_carry = ein0;
                    need_lb = false;
                }
                // last run head before each item: nearest lower lane that has one, else what the earlier steps left
                int ein;
                {
                    const unsigned hm = __ballot_sync(0xffffffffu, lh != kNoHead);
                    const unsigned below = hm & ((1u << lane) - 1u);
                    const int g = __shfl_sync(0xffffffffu, lh, below ? 31 - __clz(below) : 0);
                    ein = below ? g : ein_carry;
                    const int last = __shfl_sync(0xffffffffu, lh, hm ? 31 - __clz(hm) : 0);
                    if (hm) ein_carry = last;
                }
                uint32_t cf = 0;
                int h0 = 0;
                const int n0 = item_count(it, ein, &cf, &h0);
#if VCFC_ENC_PFW
#define VCFC_LOAD_NEXT_STEP()                                                                                                   \
                if (!endm) {                                                                                                    \
                    const int a2 = a + 4 * ((bound - a + 3) >> 2), ws2 = a2 & ~63;                                              \
                    if (a2 < ce && ws2 - 4 >= r_lo && ws2 + kStep + 4 <= r_hi) {                                                \
                        const uint8_t* const pb2 = win + ws2 + 64 * lane;                                                       \
                        _Pragma("unroll") for (int q = 0; q < 4; q++) {                                                         \
                            const uint4 v = *reinterpret_cast<const uint4*>(pb2 + 16 * q);                                      \
                            W[4 * q + 1] = v.x; W[4 * q + 2] = v.y; W[4 * q + 3] = v.z; W[4 * q + 4] = v.w;                     \
                        }                                                                                                       \
                        W[0] = *reinterpret_cast<const uint32_t*>(pb2 - 4);                                                     \
                        W[17] = *reinterpret_cast<const uint32_t*>(pb2 + 64);                                                   \
                        w_ready = true;                                                                                         \
                    }                                                                                                           \
                }
#endif
#if VCFC_ENC_PFW == 1
                VCFC_LOAD_NEXT_STEP()
#endif
                int inc = n0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
                const int step_total = __shfl_sync(0xffffffffu, inc, 31);
                if (o - flushed + step_total > kSStage) {
                    flush_segment(stage, o - flushed, log, ctrl, log_cap, lane, &seg_first, &seg_prev, &dead, nl_seg, my_off, my_off2);
                    flushed = o; nl_seg = 0;
                }
#if VCFC_ENC_PFW == 2
                VCFC_LOAD_NEXT_STEP()
#endif
                item_emit(win, it, cf, h0, stage + (o - flushed + inc - n0), rare);
                o += step_total;
                first = false;
                if (endm) { cur = q_end + 1; in_req = true; }
                else cur = a + 4 * ((bound - a + 3) >> 2);
#if VCFC_ENC_PFW
                if (!w_ready) break;
                }
                if (odd_pending) break;
#endif
            }
            if constexpr (!kOdd) break;
            if (!odd_pending) break;
            if constexpr (kOdd) {   // (out of the hot loop; nothing is live across the call: see OddSave)
                odd_pending = false;
                OddSave& S = sm.odd[warp];
                if (lane == 0) {
                    S.o = o; S.flushed = flushed; S.nl = nl; S.nl_seg = nl_seg; S.cur = cur; S.ein_carry = ein_carry; S.ein0 = ein0;
                    S.bits = (in_req ? 1 : 0) | (first ? 2 : 0) | (need_lb ? 4 : 0); S.irregular = 0;
                    S.ce = ce; S.cs = cs; S.pc0 = pc0; S.tile = tile; S.cur_ce = cur_ce; S.cur_fl = cur_fl; S.irr_seen = irr_seen; S.fl = fl;
#if VCFC_ENC_TICKET
                    S.nt = nt; S.nn = nn;
#endif
                }
                S.my_off[lane] = my_off; S.my_off2[lane] = my_off2;
                __syncwarp();
                odd_call(&S, in, n, tile_sz, s1, stage, log, ctrl, log_cap, serial_budget, lane, &seg_first, &seg_prev, &dead);
                __syncwarp();
                o = S.o; flushed = S.flushed; nl = S.nl; nl_seg = S.nl_seg; cur = S.cur; ein_carry = S.ein_carry; ein0 = S.ein0;
                in_req = (S.bits & 1) != 0; first = (S.bits & 2) != 0; need_lb = (S.bits & 4) != 0;
                ce = S.ce; cs = S.cs; pc0 = S.pc0; tile = S.tile; cur_ce = S.cur_ce; cur_fl = S.cur_fl; irr_seen = S.irr_seen; fl = S.fl;
#if VCFC_ENC_TICKET
                nt = S.nt; nn = S.nn;
#endif
                my_off = S.my_off[lane]; my_off2 = S.my_off2[lane];
                t0 = (long long)tile * tile_sz; gb = t0 - 64; win = in + gb;
                r_lo = gb < 0 ? 64 : 0;
                r_hi = (int)(n - gb < (long long)(1 << 24) ? n - gb : (long long)(1 << 24));
                pf_lim = min(r_hi, ce + 127); pf_lim2 = min(r_hi, ce + 4095);
                ke = fl & 15; lb_lc = (fl >> 4) & 15; lb_uniform = (fl >> 8) & 1; lb_nsamp = (ce - 64) >> 2;
                const int action = S.action;
                if (action == 2) { irregular = S.irregular; break; }
            }
            }
            // the input ends inside a line that no newline-less last sample closed (a trailing tab, a cut sample): generic path
            if (!irregular && ke == kCutEnd && cs < ce && !in_req) irregular = 1;
            // ---- the tile's last segment, followed by the u32 line offsets; the final position comes from the scan over the
            //      tile records (k_gather_tiles follows the segment chain) ------------------------------------------------------
            if (irregular) {
                if (lane == 0 && atomicCAS(&ctrl->irregular, 0, irregular) == 0) ctrl->total_lines = (unsigned long long)tile;   // (diagnostics)
                o = 0; nl = 0; flushed = 0;
            }
            if (o > flushed) flush_segment(stage, o - flushed, log, ctrl, log_cap, lane, &seg_first, &seg_prev, &dead, nl_seg, my_off, my_off2);
            if (lane == 0) { rec_pos[tile] = seg_first; rec_size[tile] = (unsigned long long)o; rec_lines[tile] = (unsigned long long)nl; }
        }
    }
}

// Final totals from the tile-record scans; decides capacity before any byte reaches the caller's buffer.
__global__ void k_enc_totals(Ctrl* __restrict__ ctrl, unsigned long long out_cap) {
    if (ctrl->total_bytes > out_cap) ctrl->cap_exceeded = 1;
    if (ctrl->total_lines > ctrl->line_cap && !ctrl->irregular) ctrl->irregular = 8;
}

// One warp per tile: the tile's segment chain in the log -> out[off ..], line offsets rebased from the trailer behind the last segment.
__global__ void k_gather_tiles(const uint8_t* __restrict__ log, uint8_t* __restrict__ out, const Ctrl* __restrict__ ctrl,
                               const unsigned long long* __restrict__ rec_pos, const unsigned long long* __restrict__ rec_size,
                               const unsigned long long* __restrict__ rec_lines, const unsigned long long* __restrict__ off_b,
                               const unsigned long long* __restrict__ off_l, unsigned long long* __restrict__ line_offs, int n_tiles) {
    const int t = (int)(((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (t >= n_tiles || ctrl->irregular || ctrl->cap_exceeded) return;
    const unsigned long long ob = off_b[t], ol = off_l[t];
    int remaining = (int)rec_size[t];
    if (remaining == 0 || rec_pos[t] == 0ull) return;
    unsigned long long pos = rec_pos[t] - 1ull;
    uint8_t* dst = out + ob;
    int lbase = 0;
    for (;;) {
        const uint4 hdr = *reinterpret_cast<const uint4*>(log + pos);
        const int size = (int)hdr.x;
        const uint8_t* src = log + pos + 16;                  // 16-byte aligned
        // bytes until dst is 16-byte aligned, then 16 bytes per lane and round trip: five source words (the segment is 16-byte
        // aligned and padded) funnel-shifted into one 16-byte store, then the last odd bytes
        const int head = min((int)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15), size);
        if (lane < head) dst[lane] = src[lane];
        const int nquads = (size - head) >> 4;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(src + (head & ~3));
        const int sh = 8 * (head & 3);
        uint4* dq = reinterpret_cast<uint4*>(dst + head);
        for (int k = lane; k < nquads; k += 32) {
            const uint32_t* q = sw + 4 * k;
            const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3], w4 = q[4];
            dq[k] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
        }
        const int tail0 = head + 16 * nquads;
        if (lane < size - tail0) dst[tail0 + lane] = src[tail0 + lane];
        const int nls = (int)hdr.y;
        const uint32_t* tr = reinterpret_cast<const uint32_t*>(src + ((size + 15) & ~15));
        for (int l = lane; l < nls; l += 32) line_offs[ol + (unsigned long long)(lbase + l)] = ob + (unsigned long long)tr[l];
        lbase += nls;
        dst += size;
        remaining -= size;
        if (remaining <= 0) break;
        pos = (unsigned long long)hdr.z | ((unsigned long long)hdr.w << 32);
        if (pos == 0ull && size == 0) break;                   // (a broken chain: never in a block that passed the checks)
    }
}

// Line-length headers (compress.cpp:194-199: line_length = bytes after the first header) and the result block.
__global__ void k_patch_headers(uint8_t* __restrict__ out, const unsigned long long* __restrict__ line_offs,
                                Ctrl* __restrict__ ctrl, uint64_t* __restrict__ user_offs, unsigned long long user_cap,
                                vcfc_result* __restrict__ res) {
    const int irregular = ctrl->irregular, cap = ctrl->cap_exceeded;
    const unsigned long long nl = ctrl->total_lines, total = ctrl->total_bytes;
    if (!irregular && !cap) {
        for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < nl;
             k += (unsigned long long)gridDim.x * blockDim.x) {
            unsigned long long o = line_offs[k], nx = k + 1 < nl ? line_offs[k + 1] : total;
            unsigned long long len = nx - o - 4;
            if (len > 0x3FFFFFFFull) atomicExch(&ctrl->line2big, 1);
            out[o] = (uint8_t)((len >> 24) | 0xC0);
            out[o + 1] = (uint8_t)(len >> 16);
            out[o + 2] = (uint8_t)(len >> 8);
            out[o + 3] = (uint8_t)len;
            if (user_offs && k < user_cap) user_offs[k] = o;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        res->reserved = 0;
        res->err_line = 0;
        if (irregular) { res->status = kStatusIrregular; res->reserved = irregular; res->out_len = 0; res->n_lines = 0; res->err_line = nl; }
        else if (cap)  { res->status = VCFC_E_CAP; res->out_len = total; res->n_lines = 0; }
        else           { res->status = VCFC_OK; res->out_len = total; res->n_lines = nl; res->reserved = ctrl->odd_used ? 1 : 0; }   // (reserved: term walkers used)
    }
}

__global__ void k_finish_line2big(Ctrl* ctrl, vcfc_result* res) {
    if (ctrl->line2big && res->status == VCFC_OK) res->status = kStatusIrregular;   // let the generic path report it
}

__global__ void k_set_result(vcfc_result* r, int status) {
    r->status = status; r->reserved = 0; r->out_len = 0; r->n_lines = 0; r->err_line = 0;
}

}  // namespace enc

int encode_fast(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, uint8_t* d_out, size_t out_cap,
                uint64_t* d_line_out_offsets, size_t line_cap, vcfc_result* d_result, cudaStream_t stream) {
    using namespace enc;
    ctx->last_path = kPathFast;
    if (in_len == 0) {
        k_set_result<<<1, 1, 0, stream>>>(d_result, VCFC_OK);
        ctx->launches++;
        return VCFC_OK;
    }
    // host-checkable preconditions of the tile path; anything else goes to the generic kernels
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) || in_len >= (1ull << 46)) {
        k_set_result<<<1, 1, 0, stream>>>(d_result, kStatusIrregular);
        ctx->launches++;
        return VCFC_OK;
    }
    if (!ctx->enc_attr_set) {                                  // per context: one context per (process, device), used by one thread at a time
        VCFC_CUDA(ctx, cudaFuncSetAttribute(k_encode_stream<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemS)));
        VCFC_CUDA(ctx, cudaFuncSetAttribute(k_encode_stream<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSO)));
        ctx->enc_attr_set = 1;
    }
    // tile size (a multiple of 64): large tiles amortise the per-tile work (cuts, look-back record, log reservation), small
    // inputs need enough tiles to occupy every resident warp several times over
    size_t tile_sz = kSTile;
    {
        const size_t warps = (size_t)ctx->sm_count * kSCtas * kSWarps;
        while (tile_sz < (size_t)kSTileMax && in_len / (2 * tile_sz) >= 8 * warps) tile_sz *= 2;
        if (const char* e = getenv("VCFC_ENC_TILE")) { const long v = atol(e); if (v >= 4096 && v <= (1 << 22) && v % 64 == 0) tile_sz = (size_t)v; }
    }
    const size_t n_tiles = (in_len + tile_sz - 1) / tile_sz;
    const size_t lines_cap = in_len / 64 + 1024;
    DevBuf &ws = ctx->ws[10], &b_log = ctx->ws[11], &b_scr = ctx->ws[1];
    const size_t off_s1 = 256, off_rec = off_s1 + ((n_tiles * 4 + 255) & ~size_t(255));
    const size_t off_zero_end = off_rec + 3 * n_tiles * 8;                       // ctrl, s1, rec_*: zeroed every call
    const size_t off_scan = off_zero_end, off_lines = off_scan + 2 * n_tiles * 8, total_ws = off_lines + lines_cap * 8;
    int rc = dev_reserve(ctx, &ws, total_ws);
    if (rc) return rc;
    // the log holds the output once, plus per segment a 16-byte header and padding, plus 4 bytes per line start
    const size_t log_cap = out_cap + 4 * lines_cap + 48 * (n_tiles + out_cap / (kSStage - kStepMaxOut) + lines_cap / kMaxNl + 2 * (in_len / kSStage) + 4) + 64;
    if ((rc = dev_reserve(ctx, &b_log, log_cap + 64))) return rc;
    uint8_t* base = (uint8_t*)ws.p;
    Ctrl* ctrl = (Ctrl*)base;
    unsigned long long* rec_pos = (unsigned long long*)(base + off_rec);
    unsigned long long *rec_size = rec_pos + n_tiles, *rec_lines = rec_size + n_tiles;
    unsigned long long *off_b = (unsigned long long*)(base + off_scan), *off_l = off_b + n_tiles;
    unsigned long long* line_offs = (unsigned long long*)(base + off_lines);
    VCFC_CUDA(ctx, cudaMemsetAsync(base, 0, off_zero_end, stream));
    unsigned long long caps[3] = {lines_cap, 0ull, (unsigned long long)log_cap};  // line_cap, log_cursor, log_cap
    VCFC_CUDA(ctx, cudaMemcpyAsync(&ctrl->line_cap, caps, sizeof(caps), cudaMemcpyHostToDevice, stream));
    if (!ctx->enc_resident) {       // CTAs that are guaranteed to be co-resident (look-back #1 spins on its neighbour)
        int per_sm = 0;
        int per_sm_odd = 0;
        VCFC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_encode_stream<false>, 32 * kSWarps, sizeof(SmemS)));
        VCFC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_odd, k_encode_stream<true>, 32 * kSWarps, sizeof(SmemSO)));
        ctx->enc_resident = std::max(1, std::min(per_sm, per_sm_odd)) * ctx->sm_count;
    }
    const int resident = ctx->enc_resident;
    const unsigned grid = (unsigned)std::min<size_t>((n_tiles + kSWarps - 1) / kSWarps, (size_t)resident);
    if (ctx->timing) cudaEventRecord(ctx->ev[2 * kTimeEncode], stream);
    {   // cooperative launch: every CTA is resident, which look-back #1 (a CTA spins on its neighbour's tile) relies on
        const uint8_t* a_in = d_in;
        long long a_n = (long long)in_len;
        uint8_t* a_log = (uint8_t*)b_log.p;
        unsigned int* a_s1 = (unsigned int*)(base + off_s1);
        int a_tiles = (int)n_tiles, a_tile_sz = (int)tile_sz;
        unsigned long long a_cap = (unsigned long long)log_cap;
        unsigned long long a_budget = (unsigned long long)(in_len / 32 + (1u << 18));  // odd-width terms beyond ~3 % of the block: generic kernels
        void* args[] = {&a_in, &a_n, &a_log, &ctrl, &a_s1, &rec_pos, &rec_size, &rec_lines, &a_tiles, &a_cap, &a_tile_sz, &a_budget};
        const void* kern = ctx->enc_odd ? (const void*)k_encode_stream<true> : (const void*)k_encode_stream<false>;
        VCFC_CUDA(ctx, cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(32 * kSWarps), args, ctx->enc_odd ? sizeof(SmemSO) : sizeof(SmemS), stream));
    }
    if (ctx->timing) { cudaEventRecord(ctx->ev[2 * kTimeEncode + 1], stream); ctx->ev_pending[kTimeEncode] = 1; }
    ctx->launches += 1;
    if ((rc = scan_exclusive_u64(ctx, (const uint64_t*)rec_size, (uint64_t*)off_b, n_tiles, (uint64_t*)&ctrl->total_bytes, &b_scr, stream))) return rc;
    if ((rc = scan_exclusive_u64(ctx, (const uint64_t*)rec_lines, (uint64_t*)off_l, n_tiles, (uint64_t*)&ctrl->total_lines, &b_scr, stream))) return rc;
    k_enc_totals<<<1, 1, 0, stream>>>(ctrl, (unsigned long long)out_cap);
    k_gather_tiles<<<(unsigned)((n_tiles * 32 + 255) / 256), 256, 0, stream>>>((const uint8_t*)b_log.p, d_out, ctrl, rec_pos, rec_size,
                                                                               rec_lines, off_b, off_l, line_offs, (int)n_tiles);
    unsigned pb = (unsigned)std::min<size_t>((lines_cap + 255) / 256, 148 * 8);
    k_patch_headers<<<pb, 256, 0, stream>>>(d_out, line_offs, ctrl, d_line_out_offsets, (unsigned long long)line_cap, d_result);
    k_finish_line2big<<<1, 1, 0, stream>>>(ctrl, d_result);
    ctx->launches += 4;
    VCFC_CUDA(ctx, cudaGetLastError());
    return VCFC_OK;
}

}  // namespace vcfc
