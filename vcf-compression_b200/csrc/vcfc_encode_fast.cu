// vcfc_encode_fast.cu -- single-pass, tile-parallel encoder for REGULAR data lines (sm_100a).
//
// "Regular" = what a GT-only VCF looks like: '\n'-terminated lines, single tabs, >= 10 columns,
// every sample column exactly 3 bytes (a|b, a/b, ./. ...), required section (CHROM..FORMAT) of
// at most kMaxReq bytes.  Anything else sets ctrl->irregular and the caller reruns the block on
// the generic kernels (vcfc_generic.cu).  Output bytes are those of compress_data_line
// (/root/reference/src/compress.cpp:5-203) for every line.
//
// One CTA = one tile of ~14 KB of input, HBM traffic = input read once + output written once:
//   1. bulk async copy (TMA engine, cp.async.bulk + mbarrier) of the tile and 1 KB halos into smem
//   2. cut points: a tile owns the units (a sample column, or a whole required section) that START
//      in [cut(i*T), cut((i+1)*T)); both neighbours compute the shared cut from the same bytes
//   3. newline list -> line segments (required section end = 9th tab, warp ballot/popc)
//   4. items = 32-byte blocks of 8 phase-aligned sample words: classify, run heads
//   5. look-back #1 (tiny): run length carried into the tile  (chunks are 127 / 31 samples from
//      the run's head, compress.cpp:129-170, so a tile must know how long the run already is)
//   6. per-item byte counts, block scan, look-back #2 (decoupled, 16-byte status): output offset
//      and line index of the tile
//   7. emit tokens / literals / required sections into smem staging, 16-byte stores to HBM
// A second tiny kernel patches the 4-byte line-length headers (they need the NEXT line's offset)
// and writes the result block.
#include <algorithm>

#include "vcfc_common.cuh"
#include "vcfc_internal.h"

namespace vcfc {
namespace enc {

constexpr int kTile = 14336;            // nominal input bytes per tile (448 blocks of 32)
constexpr int kHalo = 1024;
constexpr int kPad = 32;                // zeroed bytes in front of / behind the window
constexpr int kMaxReq = kHalo - 64;     // longest required section taken by this path
constexpr int kWin = kTile + 2 * kHalo + 2 * kPad;
constexpr int kStage = 21504 + 32;      // staging bytes (regular data expands at most ~1.45x)
constexpr int kMaxSeg = 126;
constexpr int kMaxItems = 768;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRound = 2 * kThreads;    // items per round (2 per thread)

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
};

struct Seg {                            // a run of sample words of one line inside the tile
    int a, e;                           // sample bytes [a, e), window-relative
    int ls, s0;                         // line start / first sample if the line starts in this tile, else ls = -1
    int item0;                          // first item index
    int flags;                          // bit0: span ends with the line's '\n'; bit1: first sample of span is first of line
    int out0;                           // staging offset of the line start (header) -- valid if ls >= 0
};

struct Smem {
    alignas(128) uint8_t win[kWin];
    alignas(16) uint8_t stage[kStage];
    alignas(8) uint64_t mbar;
    uint32_t itemCls[kMaxItems];        // 8 x 4-bit classes
    uint32_t itemMeta[kMaxItems];       // valid mask (8) | head mask (8) << 8 | prev class << 16 | line-end idx+1 << 20
    int itemOff[kMaxItems];
    Seg seg[kMaxSeg + 2];
    int nlpos[kMaxSeg + 2];
    int nlsorted[kMaxSeg + 2];
    int n_nl, n_seg, n_items, n_lines;
    int tile, irregular;
    int cs, cs_kind, ce, ce_kind;
    int tile_last_head, tile_last_cls;
    int ein_virtual;
    int warp_h[kWarps + 1], warp_s[kWarps + 1];
    int carry_i, carry_head;
    int total_bytes;
    unsigned long long excl_bytes, excl_lines;
    int skip_write;
};

static_assert(sizeof(Smem) <= 53 * 1024, "4 CTAs per SM need <= ~56 KB each");

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
    for (long long base = b - 1; base < lim; base += 32) {
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
    if (p < 0) { *kind = kCutBad; return b; }
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

// ---- end of the required section of the line starting at window offset ls; warp-collective ------
// Returns the window offset of the first sample (after the 9th tab) or -1 when irregular.
__device__ int line_scan(const uint8_t* __restrict__ win, int ls, int vhi_w, int lane) {
    uint32_t c0 = win[ls];
    if (is_sep(c0)) return -1;                       // empty first column / empty line
    int tabs = 0;
    unsigned carry = 0;
    for (int base = ls; base < ls + kMaxReq + 32; base += 32) {
        int g = base + lane;
        uint32_t c = g < vhi_w ? win[g] : (uint32_t)'\n';
        unsigned tabm = __ballot_sync(0xffffffffu, c == '\t');
        unsigned nlm = __ballot_sync(0xffffffffu, c == '\n');
        int cnt = __popc(tabm);
        if (tabs + cnt >= 9) {
            int j = __fns(tabm, 0, 9 - tabs);
            unsigned upto = j == 31 ? 0xffffffffu : ((1u << (j + 1)) - 1u);
            if (nlm & upto) return -1;
            if ((tabm & ((tabm << 1) | carry)) & upto) return -1;   // empty field in the required section
            int s0 = base + j + 1;
            return (s0 - ls) <= kMaxReq ? s0 : -1;
        }
        if (nlm) return -1;                          // fewer than 10 columns
        if (tabm & ((tabm << 1) | carry)) return -1;
        tabs += cnt;
        carry = tabm >> 31;
    }
    return -1;
}

// ---- sample word helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t classify_word(uint32_t s) {     // 0..3 coded, 4 otherwise
    uint32_t m = s & 0xFFFEFFFEu;
    return (m == 0x09307C30u || m == 0x0A307C30u) ? (((s & 1u) << 1) | ((s >> 16) & 1u)) : 4u;
}
__device__ __forceinline__ bool has_tab_low3(uint32_t s) {
    uint32_t v = ((s ^ 0x00090909u) & 0x00FFFFFFu) | 0xFF000000u;
    return ((v - 0x01010101u) & ~v & 0x80808080u) != 0;
}
__device__ __forceinline__ uint32_t cls_flag(uint32_t c) { return c == 0 ? kTok00 : c == 1 ? kTok01 : c == 2 ? kTok10 : kTok11; }
__device__ __forceinline__ int cls_max(uint32_t c) { return c == 0 ? 127 : 31; }

__device__ __forceinline__ int seg_of_item(const Smem& sm, int item) {
    int lo = 0, hi = sm.n_seg - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (sm.seg[mid].item0 <= item) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Walks the (up to 8) samples of an item in order, replaying the reference's run logic
// (compress.cpp:124-186).  cp/cnt = open run class / count before the item.  Returns bytes emitted;
// when WRITE, stores them at dst.  *cp_out/*cnt_out = state after the item.
template <bool WRITE>
__device__ __forceinline__ int item_walk(const uint8_t* __restrict__ win, int blk, int phase, uint32_t cls8, uint32_t meta,
                                         int cp, int cnt, uint8_t* __restrict__ dst, int* cp_out, int* cnt_out) {
    int o = 0;
    const uint32_t valid = meta & 0xFFu;
    const int lineend = (int)((meta >> 20) & 0xFu) - 1;      // index of the sample that ends the line, or -1
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (!((valid >> k) & 1u)) continue;
        int c = (int)((cls8 >> (4 * k)) & 0xFu);
        if (cp < 4 && (c != cp || cnt == cls_max(cp))) {
            if (WRITE) dst[o] = (uint8_t)(cls_flag(cp) | (uint32_t)cnt);
            o++;
            cp = kNone;
        }
        if (c == 4) {
            if (WRITE) {
                const uint8_t* s = win + blk + phase + 4 * k;
                dst[o] = (uint8_t)(kTokLit | 1u);
                dst[o + 1] = s[0]; dst[o + 2] = s[1]; dst[o + 3] = s[2];
                if (k != lineend) dst[o + 4] = '\t';
            }
            o += (k != lineend) ? 5 : 4;
        } else if (cp == kNone) {
            cp = c;
            cnt = 1;
        } else {
            cnt++;
        }
        if (k == lineend) {
            if (cp < 4) { if (WRITE) dst[o] = (uint8_t)(cls_flag(cp) | (uint32_t)cnt); o++; }
            if (WRITE) dst[o] = '\n';
            o++;
            cp = kNone;
        }
    }
    *cp_out = cp;
    *cnt_out = cnt;
    return o;
}

// open-run state before the first valid sample of an item, from the last run head before it
__device__ __forceinline__ void run_state_before(uint32_t meta, int first_addr, int ein, int* cp, int* cnt) {
    int pc = (int)((meta >> 16) & 0xFu);
    *cp = pc < 4 ? pc : kNone;
    *cnt = 0;
    if (pc < 4) {
        int nrun = (first_addr - ein) >> 2;              // samples of the open run so far (>= 1)
        *cnt = ((nrun - 1) % cls_max(pc)) + 1;
    }
}

__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_status(const unsigned long long* p, unsigned long long* a, unsigned long long* b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(*a), "=l"(*b) : "l"(p) : "memory");
}

__global__ void __launch_bounds__(kThreads, 4)
k_encode_tiles(const uint8_t* __restrict__ in, long long n, uint8_t* __restrict__ out, unsigned long long out_cap,
               Ctrl* __restrict__ ctrl, unsigned int* __restrict__ s1, unsigned long long* __restrict__ s2,
               unsigned long long* __restrict__ line_offs, int n_tiles) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- 0. ticket (tiles are started in index order, which look-back relies on) ------------------
    if (tid == 0) {
        sm.tile = (int)atomicAdd(&ctrl->ticket, 1u);
        sm.irregular = *((volatile int*)&ctrl->irregular);
        sm.n_nl = 0;
        sm.skip_write = 0;
        mbar_init(&sm.mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int tile = sm.tile;
    if (tile >= n_tiles) return;
    const long long t0 = (long long)tile * kTile;
    const long long wbase = t0 - kHalo - kPad;                 // global offset of win[0] (multiple of 32)
    const long long vlo = t0 - kHalo > 0 ? t0 - kHalo : 0;      // valid global range held in the window
    const long long vhi = t0 + kTile + kHalo < n ? t0 + kTile + kHalo : n;
    const int vlo_w = (int)(vlo - wbase), vhi_w = (int)(vhi - wbase);

    if (sm.irregular) {          // some tile already gave up: keep the look-back chains alive and leave
        if (tid == 0) {
            s1[tile] = (2u << 30) | ((unsigned)kNone << 8);
            st_status(s2 + 2 * (size_t)tile, kFlagAgg, kFlagAgg);
        }
        return;
    }

    // ---- 1. stage the window: one bulk async copy + a few tail bytes --------------------------------
    const unsigned bulk = (unsigned)((vhi - vlo) & ~15ll);
    if (tid == 0) {
        mbar_expect_tx(&sm.mbar, bulk);
        if (bulk)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sm.win + vlo_w)), "l"(in + vlo), "r"(bulk), "r"(smem_u32(&sm.mbar)) : "memory");
    }
    for (int g = vlo_w + (int)bulk + tid; g < vhi_w; g += kThreads) sm.win[g] = in[wbase + g];
    if (tid < kPad) { sm.win[vlo_w - kPad + tid] = 0; }
    if (tid >= 32 && tid < 32 + kPad + 16) { int g = vhi_w + tid - 32; if (g < kWin) sm.win[g] = 0; }
    mbar_wait(&sm.mbar, 0);
    __syncthreads();
    if (tid == 0 && vhi == n && sm.win[vhi_w - 1] != '\n') sm.irregular = 1;   // no final newline: generic path

    // ---- 2. cut points ---------------------------------------------------------------------------------
    if (warp < 2) {
        int kind;
        long long c = cut_find(sm.win, wbase, vlo, vhi, n, warp == 0 ? t0 : t0 + kTile, lane, &kind);
        if (lane == 0) {
            if (warp == 0) { sm.cs = (int)(c - wbase); sm.cs_kind = kind; }
            else           { sm.ce = (int)(c - wbase); sm.ce_kind = kind; }
            if (kind == kCutBad) sm.irregular = 1;
        }
    }
    __syncthreads();
    const int cs = sm.cs, ce = sm.ce, cs_kind = sm.cs_kind;
    bool bad = sm.irregular != 0;

    // ---- 3. newline list of [cs, ce) ---------------------------------------------------------------------
    if (!bad) {
        for (int c16 = (cs >> 4) + tid; (c16 << 4) < ce; c16 += kThreads) {
            uint4 v = *reinterpret_cast<const uint4*>(sm.win + (c16 << 4));
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t z = zero_bytes(w[j] ^ 0x0A0A0A0Au);
                while (z) {
                    int bpos = (__ffs(z) - 1) >> 3;
                    z &= z - 1;
                    int g = (c16 << 4) + 4 * j + bpos;
                    if (g >= cs && g < ce) {
                        int slot = atomicAdd(&sm.n_nl, 1);
                        if (slot < kMaxSeg) sm.nlpos[slot] = g;
                    }
                }
            }
        }
    }
    __syncthreads();
    int n_nl = sm.n_nl;
    if (n_nl > kMaxSeg - 2) { bad = true; n_nl = 0; }
    if (!bad) {   // rank sort (n_nl is tiny for real data)
        for (int i = tid; i < n_nl; i += kThreads) {
            int v = sm.nlpos[i], r = 0;
            for (int j = 0; j < n_nl; j++) r += sm.nlpos[j] < v;
            sm.nlsorted[r] = v;
        }
    }
    __syncthreads();

    // ---- 4. segments: [partial first line] + one per line start in (cs, ce) --------------------------------
    // line starts: cs itself when cs_kind == kCutLine, and nl+1 for every newline with nl+1 < ce
    const int first_partial = (cs_kind == kCutSample || cs_kind == kCutSampleFirst) && cs < ce ? 1 : 0;
    int n_lines = 0;
    if (!bad) {
        n_lines = (cs_kind == kCutLine && cs < ce ? 1 : 0);
        for (int i = 0; i < n_nl; i++) n_lines += (sm.nlsorted[i] + 1 < ce);     // uniform, tiny
        const int n_seg = first_partial + n_lines;
        if (n_seg > kMaxSeg) bad = true;
        if (!bad) {
            if (first_partial && tid == 0) {
                Seg& s = sm.seg[0];
                s.a = cs; s.ls = -1; s.s0 = -1;
                s.e = n_nl ? sm.nlsorted[0] + 1 : ce;
                s.flags = (n_nl ? 1 : 0) | (cs_kind == kCutSampleFirst ? 2 : 0);
            }
            const int own_first = (cs_kind == kCutLine && cs < ce) ? 1 : 0;
            for (int l = warp; l < n_lines; l += kWarps) {
                // line l starts at cs (if own_first and l == 0) or after newline (l - own_first)
                int nli = l - own_first;                       // index of the newline before this line
                int ls = nli < 0 ? cs : sm.nlsorted[nli] + 1;
                int s0 = line_scan(sm.win, ls, vhi_w, lane);
                if (lane == 0) {
                    Seg& s = sm.seg[first_partial + l];
                    int nxt = nli + 1;                          // the newline that ends this line, if in the tile
                    s.ls = ls; s.s0 = s0; s.a = s0;
                    s.e = nxt < n_nl ? sm.nlsorted[nxt] + 1 : ce;
                    s.flags = (nxt < n_nl ? 1 : 0) | 2;
                    if (s0 < 0 || s0 > s.e) sm.irregular = 1;
                }
            }
            if (tid == 0) { sm.n_seg = n_seg; sm.n_lines = n_lines; }
        }
    }
    __syncthreads();
    bad = bad || sm.irregular != 0;
    if (!bad && tid == 0) {
        int items = 0;
        for (int i = 0; i < sm.n_seg; i++) {
            Seg& s = sm.seg[i];
            if (((s.e - s.a) & 3) != 0) { sm.irregular = 1; break; }
            s.item0 = items;
            // blocks are indexed by where a sample word STARTS; a line start with no sample here still needs an item
            items += s.e > s.a ? ((s.e - 4) >> 5) - (s.a >> 5) + 1 : 1;
        }
        sm.n_items = items;
        if (items > kMaxItems) sm.irregular = 1;
        sm.tile_last_head = kNoHead;
        sm.tile_last_cls = kNone;
    }
    __syncthreads();
    bad = bad || sm.irregular != 0;
    const int n_items = bad ? 0 : sm.n_items;
    const int n_seg = bad ? 0 : sm.n_seg;

    // ---- 5. classify items ---------------------------------------------------------------------------------
    // item -> (segment, 32-byte block); sample k of the block sits at blk + phase + 4k
    int my_last_head = kNoHead;
    for (int base = 0; base < n_items; base += kRound) {
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
            int item = base + warp * 64 + sub * 32 + lane;
            if (item >= n_items) continue;
            int si = seg_of_item(sm, item);
            const Seg sg = sm.seg[si];
            int blk = ((sg.a >> 5) + (item - sg.item0)) << 5;
            int phase = sg.a & 3;
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(sm.win + blk);
            uint32_t W[10];
            W[0] = wp[-1];
            uint4 v0 = *reinterpret_cast<const uint4*>(wp), v1 = *reinterpret_cast<const uint4*>(wp + 4);
            W[1] = v0.x; W[2] = v0.y; W[3] = v0.z; W[4] = v0.w; W[5] = v1.x; W[6] = v1.y; W[7] = v1.z; W[8] = v1.w;
            W[9] = wp[8];
            uint32_t cls8 = 0, valid = 0, heads = 0;
            int lineend = -1, prevc = kNone, lasthead = kNoHead, lastc = kNone;
            bool first = true, irr = false;
            int pc = 4;                                            // class of the previous word (4 = not coded)
            {
                uint32_t sp = __funnelshift_r(W[0], W[1], 8 * phase);
                pc = (int)classify_word(sp);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                int addr = blk + phase + 4 * k;
                uint32_t s = __funnelshift_r(W[k + 1], W[k + 2], 8 * phase);
                int c = (int)classify_word(s);
                bool v = addr >= sg.a && addr < sg.e;
                if (v) {
                    bool ends = (addr + 4 == sg.e) && (sg.flags & 1);
                    uint32_t b3 = s >> 24;
                    if (ends ? (b3 != '\n') : (b3 != '\t')) irr = true;
                    if (c == 4 && has_tab_low3(s)) irr = true;
                    bool fol = (addr == sg.a) && (sg.flags & 2);
                    bool head = fol || c == 4 || c != pc;
                    if (first) { prevc = fol ? kNone : pc; first = false; }
                    valid |= 1u << k;
                    cls8 |= (uint32_t)c << (4 * k);
                    if (head) { heads |= 1u << k; lasthead = addr; }
                    if (ends) lineend = k;
                    lastc = ends ? kNone : c;
                }
                pc = c;
            }
            if (irr) sm.irregular = 1;
            sm.itemCls[item] = cls8;
            sm.itemMeta[item] = valid | (heads << 8) | ((uint32_t)prevc << 16) | ((uint32_t)(lineend + 1) << 20);
            my_last_head = max(my_last_head, lasthead);
            if (item == n_items - 1) sm.tile_last_cls = valid ? lastc : kNone;
        }
    }
    // tile summary for look-back #1: last run head of the tile
    {
        int m = my_last_head;
#pragma unroll
        for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (lane == 0) sm.warp_h[warp] = m;
    }
    __syncthreads();
    bad = bad || sm.irregular != 0;

    // ---- 6. look-back #1: run length carried into the tile -------------------------------------------------------
    if (tid == 0) {
        int tl = kNoHead;
        for (int w = 0; w < kWarps; w++) tl = max(tl, sm.warp_h[w]);
        sm.tile_last_head = tl;
        int ein = kNoHead;
        unsigned my = (2u << 30) | ((unsigned)kNone << 8);        // default: nothing carried out (bad tile / line end)
        if (!bad) {
            const int lc = sm.tile_last_cls;                       // class of the tile's last sample, kNone after '\n'
            const bool need_in = first_partial && cs_kind == kCutSample;   // first sample continues a line
            int pc0 = kNone;
            if (need_in) {
                uint32_t meta0 = sm.itemMeta[0];
                pc0 = (int)((meta0 >> 16) & 0xFu);
            }
            const bool uniform = tl == kNoHead;                   // no run head in the tile: the incoming run covers it
            const int nsamp = (ce - cs) >> 2;                      // only meaningful when uniform (single partial segment)
            if (!uniform) {
                unsigned cnt = lc < 4 ? (unsigned)((((ce - tl) >> 2) - 1) % cls_max(lc)) + 1u : 0u;
                my = (2u << 30) | ((unsigned)lc << 8) | cnt;
                s1[tile] = my;                                     // absolute: publish before waiting
                __threadfence();
            } else if (lc < 4) {
                s1[tile] = (1u << 30) | ((unsigned)lc << 8) | (unsigned)(nsamp % cls_max(lc));   // relative
                __threadfence();
            }
            int cnt_in = 0;
            if (need_in && pc0 < 4 && tile > 0) {
                const int M = cls_max(pc0);
                int acc = 0;
                for (int j = tile - 1; j >= 0; j--) {
                    unsigned v;
                    do { v = *((volatile unsigned*)&s1[j]); } while ((v >> 30) == 0);
                    acc += (int)(v & 0xFFu);
                    if ((v >> 30) == 2u) break;
                }
                cnt_in = ((acc - 1) % M + M) % M + 1;              // open chunk count before the tile, 1..M
                ein = cs - 4 * cnt_in;
            }
            if (uniform) {
                unsigned cnt = lc < 4 ? (unsigned)((cnt_in + nsamp - 1) % cls_max(lc)) + 1u : 0u;
                s1[tile] = (2u << 30) | ((unsigned)lc << 8) | cnt;
                __threadfence();
            }
        } else {
            s1[tile] = my;
            __threadfence();
        }
#ifdef VCFC_DEBUG
        printf("tile %d cs=%d(k%d) ce=%d bad=%d n_items=%d n_seg=%d tl=%d lc=%d s1=%08x ein=%d wbase=%lld\n", tile, cs, cs_kind, ce, (int)bad,
               n_items, n_seg, tl, sm.tile_last_cls, s1[tile], ein, wbase);
#endif
        sm.ein_virtual = ein;
        sm.carry_i = 0;
        sm.carry_head = ein;
    }
    __syncthreads();

    // ---- 7. byte counts: exclusive max-scan of run heads, walk, exclusive sum-scan ----------------------------------
    for (int base = 0; base < n_items; base += kRound) {
        int head_sub[2], cnt_sub[2], ein_sub[2];
        uint32_t cls_sub[2], meta_sub[2];
        int blk_sub[2], ph_sub[2], first_sub[2];
        // (a) last head per item
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
            int item = base + warp * 64 + sub * 32 + lane;
            head_sub[sub] = kNoHead; cls_sub[sub] = 0; meta_sub[sub] = 0; blk_sub[sub] = 0; ph_sub[sub] = 0; first_sub[sub] = 0;
            if (item < n_items) {
                int si = seg_of_item(sm, item);
                const Seg sg = sm.seg[si];
                int blk = ((sg.a >> 5) + (item - sg.item0)) << 5, phase = sg.a & 3;
                uint32_t meta = sm.itemMeta[item];
                uint32_t heads = (meta >> 8) & 0xFFu, valid = meta & 0xFFu;
                if (heads) head_sub[sub] = blk + phase + 4 * (31 - __clz(heads));
                first_sub[sub] = valid ? blk + phase + 4 * (__ffs(valid) - 1) : 0;
                cls_sub[sub] = sm.itemCls[item]; meta_sub[sub] = meta; blk_sub[sub] = blk; ph_sub[sub] = phase;
            }
        }
        // (b) exclusive max-scan over the round's items (item order = warp-major, sub, lane)
        int warp_tot;
        {
            int inc0 = head_sub[0];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc0, d); if (lane >= d) inc0 = max(inc0, t); }
            int tot0 = __shfl_sync(0xffffffffu, inc0, 31);
            int ex0 = __shfl_up_sync(0xffffffffu, inc0, 1);
            if (lane == 0) ex0 = kNoHead;
            int inc1 = head_sub[1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc1, d); if (lane >= d) inc1 = max(inc1, t); }
            int tot1 = __shfl_sync(0xffffffffu, inc1, 31);
            int ex1 = __shfl_up_sync(0xffffffffu, inc1, 1);
            if (lane == 0) ex1 = kNoHead;
            ein_sub[0] = ex0;
            ein_sub[1] = max(ex1, tot0);
            warp_tot = max(tot0, tot1);
        }
        if (lane == 0) sm.warp_h[warp] = warp_tot;
        __syncthreads();
        {
            int pre = sm.carry_head;
            for (int w = 0; w < warp; w++) pre = max(pre, sm.warp_h[w]);
            ein_sub[0] = max(ein_sub[0], pre);
            ein_sub[1] = max(ein_sub[1], pre);
        }
        __syncthreads();
        if (tid == 0) { int c = sm.carry_head; for (int w = 0; w < kWarps; w++) c = max(c, sm.warp_h[w]); sm.carry_head = c; }
        // (c) count bytes per item (+ header and required section on the first item of a line that starts here)
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
            int item = base + warp * 64 + sub * 32 + lane;
            cnt_sub[sub] = 0;
            if (item < n_items) {
                int cp, cnt, cpo, cno;
                run_state_before(meta_sub[sub], first_sub[sub], ein_sub[sub], &cp, &cnt);
                int nb = item_walk<false>(sm.win, blk_sub[sub], ph_sub[sub], cls_sub[sub], meta_sub[sub], cp, cnt, nullptr, &cpo, &cno);
                int si = seg_of_item(sm, item);
                if (sm.seg[si].item0 == item && sm.seg[si].ls >= 0) nb += 8 + (sm.seg[si].s0 - sm.seg[si].ls);
                cnt_sub[sub] = nb;
            }
        }
        // (d) exclusive sum-scan
        {
            int inc0 = cnt_sub[0];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc0, d); if (lane >= d) inc0 += t; }
            int tot0 = __shfl_sync(0xffffffffu, inc0, 31);
            int inc1 = cnt_sub[1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc1, d); if (lane >= d) inc1 += t; }
            int tot1 = __shfl_sync(0xffffffffu, inc1, 31);
            if (lane == 0) sm.warp_s[warp] = tot0 + tot1;
            __syncthreads();
            int pre = sm.carry_i;
            for (int w = 0; w < warp; w++) pre += sm.warp_s[w];
            int item0 = base + warp * 64 + lane, item1 = item0 + 32;
            if (item0 < n_items) sm.itemOff[item0] = pre + inc0 - cnt_sub[0];
            if (item1 < n_items) sm.itemOff[item1] = pre + tot0 + inc1 - cnt_sub[1];
            __syncthreads();
            if (tid == 0) { int c = sm.carry_i; for (int w = 0; w < kWarps; w++) c += sm.warp_s[w]; sm.carry_i = c; }
        }
        __syncthreads();
    }
    __syncthreads();
    int total = bad ? 0 : sm.carry_i;
    if (total > kStage - 32) { bad = true; total = 0; if (tid == 0) sm.irregular = 1; }
    if (bad && tid == 0) atomicExch(&ctrl->irregular, 1);

    // ---- 8. look-back #2: output offset and line index of the tile ------------------------------------------------
    if (warp == 0) {
        const unsigned long long my_b = (unsigned long long)total, my_l = (unsigned long long)(bad ? 0 : sm.n_lines);
        if (lane == 0) st_status(s2 + 2 * (size_t)tile, kFlagAgg | my_b, kFlagAgg | my_l);
        unsigned long long eb = 0, el = 0;
        int j = tile - 1;
        while (j >= 0) {
            int idx = j - lane;
            unsigned long long a = kFlagPrefix, b = kFlagPrefix;
            if (idx >= 0) {
                do { ld_status(s2 + 2 * (size_t)idx, &a, &b); } while ((a >> 62) == 0 || (a >> 62) != (b >> 62));
            }
            unsigned pm = __ballot_sync(0xffffffffu, (a >> 62) == 2ull);
            int firstp = pm ? __ffs(pm) - 1 : 32;
            unsigned long long ca = lane <= firstp ? (a & kValMask) : 0ull, cb = lane <= firstp ? (b & kValMask) : 0ull;
#pragma unroll
            for (int d = 16; d; d >>= 1) { ca += __shfl_xor_sync(0xffffffffu, ca, d); cb += __shfl_xor_sync(0xffffffffu, cb, d); }
            eb += ca; el += cb;
            if (pm) break;
            j -= 32;
        }
        if (lane == 0) {
            st_status(s2 + 2 * (size_t)tile, kFlagPrefix | (eb + my_b), kFlagPrefix | (el + my_l));
            sm.excl_bytes = eb;
            sm.excl_lines = el;
            if (eb + my_b > out_cap) { sm.skip_write = 1; atomicExch(&ctrl->cap_exceeded, 1); }
            if (el + my_l > ctrl->line_cap) { sm.skip_write = 1; atomicExch(&ctrl->irregular, 1); }
            if (tile == n_tiles - 1) { ctrl->total_bytes = eb + my_b; ctrl->total_lines = el + my_l; }
        }
    }
    __syncthreads();
    if (bad || sm.skip_write) return;
    const unsigned long long obase = sm.excl_bytes;
    const int shift = (int)(obase & 15ull);                   // staging is laid out 16-byte congruent with the output
    uint8_t* const stg = sm.stage + shift;

    // ---- 9. emit into staging --------------------------------------------------------------------------------------
    // (a) sample bytes; the max-scan is replayed from itemOff-independent data: ein is recomputed per item
    sm.carry_head = sm.ein_virtual;   // all threads write the same value
    __syncthreads();
    for (int base = 0; base < n_items; base += kRound) {
        int head_sub[2], ein_sub[2];
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
            int item = base + warp * 64 + sub * 32 + lane;
            head_sub[sub] = kNoHead;
            if (item < n_items) {
                int si = seg_of_item(sm, item);
                const Seg sg = sm.seg[si];
                int blk = ((sg.a >> 5) + (item - sg.item0)) << 5, phase = sg.a & 3;
                uint32_t heads = (sm.itemMeta[item] >> 8) & 0xFFu;
                if (heads) head_sub[sub] = blk + phase + 4 * (31 - __clz(heads));
            }
        }
        int warp_tot;
        {
            int inc0 = head_sub[0];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc0, d); if (lane >= d) inc0 = max(inc0, t); }
            int tot0 = __shfl_sync(0xffffffffu, inc0, 31);
            int ex0 = __shfl_up_sync(0xffffffffu, inc0, 1);
            if (lane == 0) ex0 = kNoHead;
            int inc1 = head_sub[1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc1, d); if (lane >= d) inc1 = max(inc1, t); }
            int tot1 = __shfl_sync(0xffffffffu, inc1, 31);
            int ex1 = __shfl_up_sync(0xffffffffu, inc1, 1);
            if (lane == 0) ex1 = kNoHead;
            ein_sub[0] = ex0;
            ein_sub[1] = max(ex1, tot0);
            warp_tot = max(tot0, tot1);
        }
        if (lane == 0) sm.warp_h[warp] = warp_tot;
        __syncthreads();
        {
            int pre = sm.carry_head;
            for (int w = 0; w < warp; w++) pre = max(pre, sm.warp_h[w]);
            ein_sub[0] = max(ein_sub[0], pre);
            ein_sub[1] = max(ein_sub[1], pre);
        }
        __syncthreads();
        if (tid == 0) { int c = sm.carry_head; for (int w = 0; w < kWarps; w++) c = max(c, sm.warp_h[w]); sm.carry_head = c; }
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
            int item = base + warp * 64 + sub * 32 + lane;
            if (item < n_items) {
                int si = seg_of_item(sm, item);
                const Seg sg = sm.seg[si];
                int blk = ((sg.a >> 5) + (item - sg.item0)) << 5, phase = sg.a & 3;
                uint32_t meta = sm.itemMeta[item], cls8 = sm.itemCls[item], valid = meta & 0xFFu;
                int first_addr = valid ? blk + phase + 4 * (__ffs(valid) - 1) : 0;
                int cp, cnt, cpo, cno;
                run_state_before(meta, first_addr, ein_sub[sub], &cp, &cnt);
                int o = sm.itemOff[item];
                if (sg.item0 == item && sg.ls >= 0) {
                    sm.seg[si].out0 = o;
                    o += 8 + (sg.s0 - sg.ls);
                }
                item_walk<true>(sm.win, blk, phase, cls8, meta, cp, cnt, stg + o, &cpo, &cno);
            }
        }
        __syncthreads();
    }
    __syncthreads();
    // (b) line starts: two length headers + required section; record the line's output offset
    {
        const unsigned long long lbase = sm.excl_lines;
        for (int l = warp; l < sm.n_lines; l += kWarps) {
            const Seg sg = sm.seg[first_partial + l];
            const int rq = sg.s0 - sg.ls;
            uint8_t* d = stg + sg.out0;
            if (lane < 4) d[lane] = lane == 0 ? 0xC0 : 0;                       // line length: patched by k_patch_headers
            if (lane >= 4 && lane < 8) {
                unsigned v = (unsigned)rq;
                d[lane] = lane == 4 ? (uint8_t)((v >> 24) | 0xC0) : (uint8_t)(v >> (8 * (7 - lane)));
            }
            for (int k = lane; k < rq; k += 32) d[8 + k] = sm.win[sg.ls + k];
            if (lane == 0) line_offs[lbase + (unsigned long long)l] = obase + (unsigned long long)sg.out0;
        }
    }
    __syncthreads();

    // ---- 10. staging -> HBM, 16-byte stores -------------------------------------------------------------------------
    {
        uint8_t* dst = out + obase;                       // dst + k <-> stg[k]; (dst - shift) is 16-byte aligned
        const int lo = shift, hi = shift + total;         // staging byte range [lo, hi) of sm.stage
        const int body_lo = (lo + 15) & ~15, body_hi = hi & ~15;
        if (body_lo >= body_hi) {
            for (int k = lo + tid; k < hi; k += kThreads) dst[k - shift] = sm.stage[k];
        } else {
            for (int k = lo + tid; k < body_lo; k += kThreads) dst[k - shift] = sm.stage[k];
            for (int k = body_hi + tid; k < hi; k += kThreads) dst[k - shift] = sm.stage[k];
            uint4* d4 = reinterpret_cast<uint4*>(dst - shift);
            const uint4* s4 = reinterpret_cast<const uint4*>(sm.stage);
            for (int k = (body_lo >> 4) + tid; k < (body_hi >> 4); k += kThreads) d4[k] = s4[k];
        }
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
        if (irregular) { res->status = kStatusIrregular; res->out_len = 0; res->n_lines = 0; }
        else if (cap)  { res->status = VCFC_E_CAP; res->out_len = total; res->n_lines = 0; }
        else           { res->status = VCFC_OK; res->out_len = total; res->n_lines = nl; }
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
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15) || in_len >= (1ull << 46)) {
        k_set_result<<<1, 1, 0, stream>>>(d_result, kStatusIrregular);
        ctx->launches++;
        return VCFC_OK;
    }
    static bool attr_set = false;
    if (!attr_set) {
        VCFC_CUDA(ctx, cudaFuncSetAttribute(k_encode_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        attr_set = true;
    }
    const size_t n_tiles = (in_len + kTile - 1) / kTile;
    const size_t lines_cap = in_len / 64 + 1024;
    DevBuf& ws = ctx->ws[10];
    const size_t off_s1 = 256, off_s2 = off_s1 + ((n_tiles * 4 + 255) & ~size_t(255));
    const size_t off_lines = off_s2 + n_tiles * 16, total_ws = off_lines + lines_cap * 8;
    int rc = dev_reserve(ctx, &ws, total_ws);
    if (rc) return rc;
    uint8_t* base = (uint8_t*)ws.p;
    Ctrl* ctrl = (Ctrl*)base;
    VCFC_CUDA(ctx, cudaMemsetAsync(base, 0, off_lines, stream));
    Ctrl h;
    memset(&h, 0, sizeof(h));
    h.line_cap = lines_cap;
    // the last byte must be '\n' (a missing final newline is the generic path's business)
    VCFC_CUDA(ctx, cudaMemcpyAsync(&ctrl->line_cap, &h.line_cap, sizeof(h.line_cap), cudaMemcpyHostToDevice, stream));
    if (ctx->timing) cudaEventRecord(ctx->ev[2 * kTimeEncode], stream);
    k_encode_tiles<<<(unsigned)n_tiles, kThreads, sizeof(Smem), stream>>>(
        d_in, (long long)in_len, d_out, (unsigned long long)out_cap, ctrl, (unsigned int*)(base + off_s1),
        (unsigned long long*)(base + off_s2), (unsigned long long*)(base + off_lines), (int)n_tiles);
    if (ctx->timing) { cudaEventRecord(ctx->ev[2 * kTimeEncode + 1], stream); ctx->ev_pending[kTimeEncode] = 1; }
    unsigned pb = (unsigned)std::min<size_t>((lines_cap + 255) / 256, 148 * 8);
    k_patch_headers<<<pb, 256, 0, stream>>>(d_out, (const unsigned long long*)(base + off_lines), ctrl, d_line_out_offsets,
                                            (unsigned long long)line_cap, d_result);
    k_finish_line2big<<<1, 1, 0, stream>>>(ctrl, d_result);
    ctx->launches += 3;
    VCFC_CUDA(ctx, cudaGetLastError());
    return VCFC_OK;
}

}  // namespace vcfc
