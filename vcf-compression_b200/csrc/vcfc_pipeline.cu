// vcfc_pipeline.cu -- the host side of the file verbs: ingest, hand-off to the GPU(s), egress.
//
// Replaces the reference's one-line-at-a-time file loops:
//   compress()        std::getline + one ofstream.write per byte          src/compress.cpp:218, 248-250
//   decompress2_fd()  dup/fdopen/fread/fclose + one write(2) per line     src/compress.cpp:1226-1248
// with a pipeline (host C++ threads, no PyTorch):
//   readers   pread() newline-aligned chunks of the file into a ring of PINNED buffers (chunk k owns the lines that
//             START in [k*C, (k+1)*C), so readers are independent of each other); the compressed file of the decode
//             direction is split along its 4-byte line-length headers by one sequential splitter
//   workers   one thread per GPU context: chunks are handed out in file order; vcfc_encode_block / vcfc_decode_block
//             move the chunk pinned -> device -> pinned on two streams (copies overlap the kernels)
//   egress    chunk outputs get their file offsets in order (a running sum of the chunks' out_len -- the per-shard
//             offsets of SURVEY.md 8(e)), writer threads pwrite() them concurrently
// '#' lines pass through on the host exactly as compress.cpp:222-238 does, wherever they stand.
// Multi-GPU: the same pipeline with one worker per context; no collective, the host concatenates by offsets.
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "vcfc_internal.h"

namespace vcfc {
namespace pipe {

using PinBuf = vcfc_ctx::PinBuf;

static size_t env_sz(const char* name, size_t dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    char* end = nullptr;
    unsigned long long x = strtoull(v, &end, 10);
    return end && *end == 0 && x > 0 ? (size_t)x : dflt;
}

// ---- pinned buffers: a pool in the lead context, kept across calls ------------------------------------------------
static PinBuf pin_get(vcfc_ctx* lead, size_t min_cap) {
    {
        std::lock_guard<std::mutex> g(lead->pin_mu);
        int best = -1;
        for (int i = 0; i < (int)lead->pin_free.size(); i++)
            if (lead->pin_free[i].cap >= min_cap && (best < 0 || lead->pin_free[i].cap < lead->pin_free[best].cap)) best = i;
        if (best >= 0) {
            PinBuf b = lead->pin_free[best];
            lead->pin_free.erase(lead->pin_free.begin() + best);
            return b;
        }
    }
    PinBuf b{nullptr, 0};
    size_t cap = (min_cap + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    if (cudaHostAlloc((void**)&b.p, cap, cudaHostAllocPortable) != cudaSuccess) { b.p = nullptr; return b; }
    b.cap = cap;
    return b;
}
static void pin_put(vcfc_ctx* lead, PinBuf b) {
    if (!b.p) return;
    std::lock_guard<std::mutex> g(lead->pin_mu);
    lead->pin_free.push_back(b);
}
// keep the pool bounded: free what a call of this size does not need again
static void pin_trim(vcfc_ctx* lead, size_t keep_bytes) {
    std::lock_guard<std::mutex> g(lead->pin_mu);
    size_t total = 0;
    for (auto& b : lead->pin_free) total += b.cap;
    while (total > keep_bytes && !lead->pin_free.empty()) {
        total -= lead->pin_free.back().cap;
        cudaFreeHost(lead->pin_free.back().p);
        lead->pin_free.pop_back();
    }
}

// ---- threads near the GPU: pinned pages are first touched, and copies are driven, from the GPU's NUMA node ---------
static void bind_near_gpu(int device) {
    if (getenv("VCFC_NO_AFFINITY")) return;
    char id[32] = {0};
    if (cudaDeviceGetPCIBusId(id, sizeof(id), device) != cudaSuccess) return;
    for (char* c = id; *c; c++) *c = (char)tolower(*c);
    std::string path = std::string("/sys/bus/pci/devices/") + id + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return;
    char line[1024] = {0};
    if (!fgets(line, sizeof(line), f)) { fclose(f); return; }
    fclose(f);
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (char* p = line; *p && *p != '\n';) {
        char* e = nullptr;
        long a = strtol(p, &e, 10), b = a;
        if (e == p) break;
        if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET((int)c, &set); n++; }
        p = (*e == ',') ? e + 1 : e;
    }
    if (n > 0) sched_setaffinity(0, sizeof(set), &set);     // best effort
}

static int pread_all(int fd, uint8_t* p, size_t n, size_t off, size_t* got_out) {
    size_t got = 0;
    while (got < n) {
        ssize_t r = pread(fd, p + got, n - got, (off_t)(off + got));
        if (r < 0 && errno == EINTR) continue;
        if (r < 0) return VCFC_E_IO;
        if (r == 0) break;
        got += (size_t)r;
    }
    *got_out = got;
    return VCFC_OK;
}
static int pwrite_all(int fd, const uint8_t* p, size_t n, size_t off) {
    while (n) {
        ssize_t w = pwrite(fd, p, n, (off_t)off);
        if (w < 0 && errno == EINTR) continue;
        if (w <= 0) return VCFC_E_IO;
        p += w; n -= (size_t)w; off += (size_t)w;
    }
    return VCFC_OK;
}

// ---- the pipeline -----------------------------------------------------------------------------------------------------
struct Chunk {
    int     state = 0;            // 0 = not read, 1 = read, 2 = transformed, 3 = offset assigned, 4 = written
    PinBuf  in{nullptr, 0}, out{nullptr, 0};
    size_t  begin = 0, end = 0;   // the chunk's bytes inside `in`
    size_t  out_len = 0, out_off = 0;
    int     rc = VCFC_OK;
    bool    last = false;         // decode: the splitter saw the end of the file
    // fused index (encode with P.want_index): per-line fields of this chunk's data lines, offsets relative to the chunk's output
    std::vector<size_t>* hash = nullptr;   // encode: offsets (in `in`) of the lines that start with '#', found by the reader
    LineIndexOut* idx = nullptr;
    size_t  n_hash = 0;           // '#' lines in the chunk
    bool    hash_after_data = false;
};

struct Pipe {
    vcfc_ctx**  ctxs;
    int         n_ctx;
    vcfc_ctx*   lead;
    int         ifd = -1, ofd = -1;
    size_t      file_len = 0;
    bool        encode = true;
    bool        want_index = false;       // encode: also collect the binned index's per-line fields
    uint64_t    sample_count = 0;         // decode
    size_t      data_off = 0;             // decode: first byte after the header region
    size_t      chunk_bytes = 0, slack = 0;
    size_t      out_base = 0;             // file offset of the first chunk's output
    bool        map_writes = false;       // writers copy into a shared mapping of the output file (needs O_RDWR)
    std::mutex  grow_mu;
    size_t      file_size = 0;            // how far the output file has been grown
    std::mutex  mu;
    std::condition_variable cv;
    std::vector<Chunk> chunks;            // encode: sized up front; decode: grows as the splitter advances
    size_t      n_chunks = 0;             // decode: known once the splitter is done
    bool        split_done = false;
    size_t      next_read = 0, next_work = 0, next_seq = 0, next_write = 0;
    size_t      in_held = 0, out_held = 0, in_max = 0, out_max = 0;
    size_t      out_total = 0;            // bytes sequenced so far (file offset = out_base + out_total)
    int         rc = VCFC_OK;             // first error in file order
    size_t      err_chunk = (size_t)-1;   // chunks behind it are dropped
    bool        abort = false;            // I/O failure: everybody stops
    size_t      total_chunks() const { return n_chunks; }
};

// compress.cpp:222-238: a line that starts with '#' passes through; one that is not "##" must have 8 columns
// (split_string drops empty terms, utils.cpp:82-116), else the reference throws "VCF Header did not have enough columns".
static int hash_line_ok(const uint8_t* p, size_t n) {
    if (n >= 2 && p[1] == '#') return VCFC_OK;
    size_t terms = 0;
    bool in_term = false;
    for (size_t i = 0; i < n; i++) {
        if (p[i] == '\t') in_term = false;
        else if (!in_term) { in_term = true; terms++; }
    }
    return terms >= 8 ? VCFC_OK : VCFC_E_HEADER;
}

// Encodes the lines of one chunk into c.out: data lines through the block codec, '#' lines verbatim + "\n".
static void encode_chunk(Pipe& P, vcfc_ctx* ctx, Chunk& c) {
    const uint8_t* p = c.in.p;
    const size_t len = c.end - c.begin;
    size_t cap = std::min(vcfc_encode_bound(len), std::max<size_t>(len / 3, (size_t)1 << 20));
    for (int attempt = 0; attempt < 2; attempt++) {
        if (!c.out.p || c.out.cap < cap) { pin_put(P.lead, c.out); c.out = pin_get(P.lead, cap); }
        if (!c.out.p) { c.rc = VCFC_E_CUDA; return; }
        size_t pos = c.begin, o = 0;
        int rc = VCFC_OK;
        if (P.want_index) {
            if (!c.idx) c.idx = new LineIndexOut();
            c.idx->offs.clear(); c.idx->ends.clear(); c.idx->refs.clear(); c.idx->errs.clear();
            c.n_hash = 0; c.hash_after_data = false;
        }
        while (pos < c.end && rc == VCFC_OK) {
            size_t h = c.end;                                       // next line that starts with '#' (found by the reader)
            if (c.hash) {
                auto it = std::lower_bound(c.hash->begin(), c.hash->end(), pos);
                if (it != c.hash->end()) h = *it;
            }
            if (h > pos) {
                size_t olen = 0, nl = 0;
                uint64_t el = 0;
                const size_t i0 = c.idx ? c.idx->offs.size() : 0;
                rc = encode_block_host(ctx, p + pos, h - pos, c.out.p + o, c.out.cap - o, &olen, nullptr, 0, &nl, &el, c.idx);
                if (rc == VCFC_E_CAP) break;
                if (c.idx) for (size_t k = i0; k < c.idx->offs.size(); k++) c.idx->offs[k] += o;
                o += olen;                                          // lines before a bad one stand
                if (rc != VCFC_OK) break;
            }
            if (h < c.end) {
                const uint8_t* e = (const uint8_t*)memchr(p + h, '\n', c.end - h);
                const size_t ll = e ? (size_t)(e - (p + h)) : c.end - h;
                if ((rc = hash_line_ok(p + h, ll))) break;
                c.n_hash++;
                if (c.idx && !c.idx->offs.empty()) c.hash_after_data = true;
                if (o + ll + 1 > c.out.cap) { rc = VCFC_E_CAP; break; }
                memcpy(c.out.p + o, p + h, ll);
                c.out.p[o + ll] = '\n';
                o += ll + 1;
                pos = h + ll + 1;
            } else {
                pos = c.end;
            }
        }
        if (rc == VCFC_E_CAP && attempt == 0) { cap = vcfc_encode_bound(len) + len / 16 + 4096; continue; }
        c.out_len = o;
        c.rc = rc;
        return;
    }
}

static void decode_chunk(Pipe& P, vcfc_ctx* ctx, Chunk& c) {
    const size_t len = c.end - c.begin;
    size_t cap = std::max<size_t>(len * 24, (size_t)4 << 20);
    for (int attempt = 0; attempt < 6; attempt++) {
        if (!c.out.p || c.out.cap < cap) { pin_put(P.lead, c.out); c.out = pin_get(P.lead, cap); }
        if (!c.out.p) { c.rc = VCFC_E_CUDA; return; }
        size_t olen = 0, nl = 0;
        uint64_t el = 0;
        int rc = vcfc_decode_block(ctx, c.in.p + c.begin, len, P.sample_count, c.out.p, c.out.cap, &olen, &nl, &el);
        if (rc == VCFC_E_CAP) { cap = c.out.cap * 4; continue; }
        c.out_len = olen;
        c.rc = rc;
        return;
    }
    c.rc = VCFC_E_CAP;
}

// ---- producers --------------------------------------------------------------------------------------------------------
// Text: chunk k owns the lines that start in [k*C, (k+1)*C); a line starts at 0 or behind a '\n'.
static void text_reader(Pipe& P) {
    bind_near_gpu(P.lead->device);
    for (;;) {
        size_t k;
        {
            std::unique_lock<std::mutex> lk(P.mu);
            P.cv.wait(lk, [&] { return P.abort || P.next_read >= P.n_chunks || P.in_held < P.in_max; });
            if (P.abort || P.next_read >= P.n_chunks) return;
            k = P.next_read++;
            P.in_held++;
        }
        Chunk c;
        const size_t C = P.chunk_bytes, lo = k * C, hi = std::min(P.file_len, lo + C);
        const size_t rd0 = lo ? lo - 1 : 0;                       // one byte ahead: is lo itself a line start?
        size_t want = (hi - rd0) + P.slack;
        int rc = VCFC_OK;
        for (;;) {
            c.in = pin_get(P.lead, want + 64);
            if (!c.in.p) { rc = VCFC_E_CUDA; break; }
            size_t got = 0;
            if ((rc = pread_all(P.ifd, c.in.p, std::min(want, P.file_len - rd0), rd0, &got))) break;
            // begin: first line start >= lo
            size_t b = 0;
            if (lo) {
                const uint8_t* q = (const uint8_t*)memchr(c.in.p, '\n', got);
                b = q ? (size_t)(q - c.in.p) + 1 : got;
            }
            // end: first line start >= hi (or the end of the file)
            size_t e;
            if (hi >= P.file_len) e = got;
            else {
                const size_t from = hi - 1 - rd0;
                const uint8_t* q = from < got ? (const uint8_t*)memchr(c.in.p + from, '\n', got - from) : nullptr;
                if (!q) {
                    if (rd0 + got >= P.file_len) e = got;            // the file's last line has no newline
                    else { pin_put(P.lead, c.in); c.in = PinBuf{nullptr, 0}; want *= 4; continue; }   // a line longer than the slack
                } else e = (size_t)(q - c.in.p) + 1;
            }
            if (b > e) b = e;
            c.begin = b; c.end = e;
            for (size_t h = b; h < e;) {                             // lines that start with '#': passed through by the worker
                const uint8_t* q = (const uint8_t*)memchr(c.in.p + h, '#', e - h);
                if (!q) break;
                h = (size_t)(q - c.in.p);
                if (h == b || c.in.p[h - 1] == '\n') {
                    if (!c.hash) c.hash = new std::vector<size_t>();
                    c.hash->push_back(h);
                }
                h++;
            }
            break;
        }
        {
            std::lock_guard<std::mutex> g(P.mu);
            c.state = 1;
            c.rc = rc;
            if (rc != VCFC_OK) P.abort = true, P.rc = P.rc ? P.rc : rc;
            P.chunks[k] = c;
        }
        P.cv.notify_all();
    }
}

// Compressed: one splitter walks the 4-byte line-length headers (compress.cpp:270-330) and cuts at line boundaries.
static void vcfc_splitter(Pipe& P) {
    bind_near_gpu(P.lead->device);
    size_t fpos = P.data_off;
    while (true) {
        {
            std::unique_lock<std::mutex> lk(P.mu);
            P.cv.wait(lk, [&] { return P.abort || P.in_held < P.in_max; });
            if (P.abort) break;
            P.in_held++;
        }
        if (P.file_len - fpos < 8) {                               // clean end (fewer than 8 bytes: EOF, compress.cpp:771-777)
            std::lock_guard<std::mutex> g(P.mu);
            P.in_held--;
            break;
        }
        Chunk c;
        size_t want = P.chunk_bytes + P.slack;
        int rc = VCFC_OK;
        size_t end = 0;
        for (;;) {
            c.in = pin_get(P.lead, want + 64);
            if (!c.in.p) { rc = VCFC_E_CUDA; break; }
            size_t got = 0;
            if ((rc = pread_all(P.ifd, c.in.p, std::min(want, P.file_len - fpos), fpos, &got))) break;
            const uint8_t* in = c.in.p;
            end = 0;
            bool broken = false;
            while (got - end >= 8) {
                if ((in[end] >> 6) != 3) { broken = true; break; }
                const size_t ll = ((size_t)(in[end] & 0x3F) << 24) | ((size_t)in[end + 1] << 16) | ((size_t)in[end + 2] << 8) | in[end + 3];
                if (ll + 4 > got - end) {
                    if (fpos + got >= P.file_len) broken = true;     // the line runs past the end of the file
                    break;
                }
                if (end + 4 + ll > P.chunk_bytes && end > 0) break;
                end += 4 + ll;
            }
            if (broken) {                                            // let the decoder classify the damage: it gets the rest of the file
                if (fpos + got < P.file_len) { pin_put(P.lead, c.in); c.in = PinBuf{nullptr, 0}; want = P.file_len - fpos; continue; }
                end = got;
                c.last = true;
            } else if (end == 0) {                                   // one line longer than the buffer
                pin_put(P.lead, c.in); c.in = PinBuf{nullptr, 0}; want *= 4; continue;
            } else if (fpos + end + 8 > P.file_len) {
                end = got = std::min(got, P.file_len - fpos);        // a tail of fewer than 8 bytes travels with the last chunk
                c.last = true;
            }
            break;
        }
        c.begin = 0; c.end = end;
        c.state = 1; c.rc = rc;
        fpos += end;
        bool stop = c.last || rc != VCFC_OK;
        {
            std::lock_guard<std::mutex> g(P.mu);
            if (rc != VCFC_OK) P.abort = true, P.rc = P.rc ? P.rc : rc;
            P.chunks.push_back(c);
        }
        P.cv.notify_all();
        if (stop) break;
    }
    {
        std::lock_guard<std::mutex> g(P.mu);
        P.n_chunks = P.chunks.size();
        P.split_done = true;
    }
    P.cv.notify_all();
}

// ---- workers: one per GPU context ---------------------------------------------------------------------------------------
static void worker(Pipe& P, int g) {
    vcfc_ctx* ctx = P.ctxs[g];
    bind_near_gpu(ctx->device);
    cudaSetDevice(ctx->device);
    for (;;) {
        size_t k;
        Chunk c;
        {
            std::unique_lock<std::mutex> lk(P.mu);
            // take the next chunk in file order together with an output slot
            P.cv.wait(lk, [&] {
                if (P.abort) return true;
                const bool known_end = P.encode || P.split_done;
                if (known_end && P.next_work >= P.n_chunks) return true;
                return P.out_held < P.out_max && P.next_work < P.chunks.size() && P.chunks[P.next_work].state >= 1;
            });
            if (P.abort) return;
            if ((P.encode || P.split_done) && P.next_work >= P.n_chunks) return;
            k = P.next_work++;
            P.out_held++;
            c = P.chunks[k];
        }
        if (k > P.err_chunk) { c.out_len = 0; }                      // behind an error: dropped
        else if (c.end > c.begin) { if (P.encode) encode_chunk(P, ctx, c); else decode_chunk(P, ctx, c); }
        pin_put(P.lead, c.in);
        c.in = PinBuf{nullptr, 0};
        {
            std::lock_guard<std::mutex> g2(P.mu);
            c.state = 2;
            P.chunks[k] = c;
            P.in_held--;
            if (c.rc != VCFC_OK && k < P.err_chunk) P.err_chunk = k;
            // offsets in file order: the per-chunk out_len running sum
            while (P.next_seq < P.chunks.size() && P.chunks[P.next_seq].state == 2) {
                Chunk& s = P.chunks[P.next_seq];
                if (P.next_seq > P.err_chunk) s.out_len = 0;
                s.out_off = P.out_base + P.out_total;
                P.out_total += s.out_len;
                if (s.rc != VCFC_OK && P.rc == VCFC_OK) P.rc = s.rc;
                s.state = 3;
                P.next_seq++;
            }
        }
        P.cv.notify_all();
    }
}

// Buffered pwrite()s to ONE file serialise on its inode lock (one memcpy at a time: ~3 GB/s on tmpfs); copying into a
// shared mapping of the chunk's range does not, so the writers really run side by side.  The file is grown first
// (ftruncate up only; the final size is set when the pipeline ends).  Falls back to pwrite where mmap is refused.
static int write_chunk(Pipe& P, const uint8_t* p, size_t n, size_t off) {
    if (P.map_writes) {
        const size_t end = off + n;
        {
            std::lock_guard<std::mutex> g(P.grow_mu);
            if (end > P.file_size) {
                const size_t want = std::max(end, P.file_size + ((size_t)256 << 20));
                if (ftruncate(P.ofd, (off_t)want) == 0) P.file_size = want;
            }
        }
        if (end <= P.file_size) {
            const size_t pg = (size_t)sysconf(_SC_PAGESIZE), lo = off & ~(pg - 1);
            void* m = mmap(nullptr, end - lo, PROT_READ | PROT_WRITE, MAP_SHARED, P.ofd, (off_t)lo);
            if (m != MAP_FAILED) {
                memcpy((uint8_t*)m + (off - lo), p, n);
                munmap(m, end - lo);
                return VCFC_OK;
            }
        }
    }
    return pwrite_all(P.ofd, p, n, off);
}

static void writer(Pipe& P) {
    for (;;) {
        size_t k;
        Chunk c;
        {
            std::unique_lock<std::mutex> lk(P.mu);
            P.cv.wait(lk, [&] {
                if (P.abort) return true;
                if ((P.encode || P.split_done) && P.next_write >= P.n_chunks) return true;
                return P.next_write < P.chunks.size() && P.chunks[P.next_write].state == 3;
            });
            if (P.abort) return;
            if ((P.encode || P.split_done) && P.next_write >= P.n_chunks) return;
            k = P.next_write++;
            c = P.chunks[k];
        }
        int rc = c.out_len ? write_chunk(P, c.out.p, c.out_len, c.out_off) : VCFC_OK;
        pin_put(P.lead, c.out);
        {
            std::lock_guard<std::mutex> g(P.mu);
            P.chunks[k].out = PinBuf{nullptr, 0};
            P.chunks[k].state = 4;
            P.out_held--;
            if (rc != VCFC_OK) { P.abort = true; if (P.rc == VCFC_OK) P.rc = rc; }
        }
        P.cv.notify_all();
    }
}

static int run(Pipe& P, int n_readers, int n_writers) {
    std::vector<std::thread> th;
    if (P.encode) for (int i = 0; i < n_readers; i++) th.emplace_back(text_reader, std::ref(P));
    else th.emplace_back(vcfc_splitter, std::ref(P));
    for (int g = 0; g < P.n_ctx; g++) th.emplace_back(worker, std::ref(P), g);
    for (int i = 0; i < n_writers; i++) th.emplace_back(writer, std::ref(P));
    for (auto& t : th) t.join();
    for (auto& c : P.chunks) { pin_put(P.lead, c.in); pin_put(P.lead, c.out); }
    return P.rc;
}

// create_binned_index4's entries (main.cpp:1430-1470, 600-626) from the per-line fields the chunks collected while they were
// encoded: a line whose number is a multiple of entries_per_bin opens an entry if its END exceeds the last entry's position,
// any other line can only grow that position.  Byte offsets are positions in the output file.
static int build_index(Pipe& P, uint64_t entries_per_bin, std::vector<uint8_t>& out, uint64_t* n_entries) {
    size_t n_ent = 0, k = 0;
    uint32_t last_end = 0;
    bool seen_data = false;
    for (auto& c : P.chunks) {
        if (c.n_hash && (seen_data || c.hash_after_data)) return VCFC_E_FORMAT;   // a '#' line behind a data line: the reference's walk fails there
        if (!c.idx) continue;
        const LineIndexOut& li = *c.idx;
        for (size_t i = 0; i < li.offs.size(); i++, k++) {
            seen_data = true;
            if (li.errs[i]) return li.errs[i] == 2 ? VCFC_E_TRUNC : VCFC_E_FORMAT;
            const unsigned long e = (unsigned long)li.ends[i];
            const uint64_t off = (uint64_t)(c.out_off + li.offs[i]);
            bool open_entry = false;
            if (n_ent == 0) { last_end = (uint32_t)e; open_entry = true; }
            else if (e > (unsigned long)last_end) {
                last_end = (uint32_t)e;
                if (k % entries_per_bin == 0) open_entry = true;
                else memcpy(out.data() + 13 * (n_ent - 1) + 1, &last_end, 4);
            }
            if (open_entry) {
                const size_t b = out.size();
                out.resize(b + 13);
                out[b] = li.refs[i];
                memcpy(out.data() + b + 1, &last_end, 4);
                memcpy(out.data() + b + 5, &off, 8);
                n_ent++;
            }
        }
    }
    if (n_entries) *n_entries = n_ent;
    return VCFC_OK;
}

static int n_threads_default(const char* env, int dflt) {
    return (int)std::max<size_t>(1, env_sz(env, (size_t)dflt));
}

}  // namespace pipe
}  // namespace vcfc

using namespace vcfc;
using namespace vcfc::pipe;

extern "C" {

// One worker thread drives one context, and a chunk goes through upload, kernels and download one after the other there: with a
// single context per GPU the bus idles while the kernels run and the host waits for the chunk's status.  So every context gets a
// twin on its device (made once, kept in the context; VCFC_FILE_TWINS=0 turns it off) and two chunks are in flight per GPU.
static void with_twins(vcfc_ctx** ctxs, int n_ctx, std::vector<vcfc_ctx*>& all) {
    all.assign(ctxs, ctxs + n_ctx);
    if (const char* e = getenv("VCFC_FILE_TWINS")) if (*e == '0') return;
    for (int i = 0; i < n_ctx; i++) {
        if (!ctxs[i]->twin && vcfc_gpu_init(ctxs[i]->device, &ctxs[i]->twin) != VCFC_OK) { ctxs[i]->twin = nullptr; continue; }
        all.push_back(ctxs[i]->twin);
    }
}

static int compress_file_impl(vcfc_ctx** ctxs_in, int n_ctx_in, const char* in_path, const char* out_path, const char* index_path,
                              uint64_t entries_per_bin, uint64_t* n_entries) {
    if (!ctxs_in || n_ctx_in <= 0 || !ctxs_in[0] || !in_path || !out_path || (index_path && entries_per_bin == 0)) return VCFC_E_ARG;
    for (int i = 0; i < n_ctx_in; i++) if (!ctxs_in[i]) return VCFC_E_ARG;
    if (n_entries) *n_entries = 0;
    std::vector<vcfc_ctx*> all;
    with_twins(ctxs_in, n_ctx_in, all);
    vcfc_ctx** ctxs = all.data();
    const int n_ctx = (int)all.size();
    Pipe P;
    P.ctxs = ctxs; P.n_ctx = n_ctx; P.lead = ctxs[0]; P.encode = true; P.want_index = index_path != nullptr;
    P.ifd = open(in_path, O_RDONLY);
    if (P.ifd < 0) return VCFC_E_IO;
    struct stat st;
    if (fstat(P.ifd, &st) != 0) { close(P.ifd); return VCFC_E_IO; }
    P.file_len = (size_t)st.st_size;
    P.ofd = open(out_path, O_CREAT | O_TRUNC | O_RDWR, 0644);
    if (P.ofd < 0) { close(P.ifd); return VCFC_E_IO; }
    P.map_writes = !getenv("VCFC_NO_MAP_WRITES");
    const int hw = (int)std::max(2u, std::thread::hardware_concurrency());
    const int n_readers = n_threads_default("VCFC_READERS", std::min(12, std::max(2, hw * 3 / 4)));   // (the readers' copies bound the verb: 8 -> 12 threads +10 % on 16 cores)
    const int n_writers = n_threads_default("VCFC_WRITERS", 3);
    // chunk size: VCFC_FILE_CHUNK_MB (default 16: measured best on tmpfs, tools/file_sweep.py), smaller for small files so that every GPU and reader has work
    size_t C = env_sz("VCFC_FILE_CHUNK_MB", 16) << 20;
    const size_t per = P.file_len / (size_t)(4 * n_ctx) + 1;
    C = std::max<size_t>((size_t)1 << 20, std::min(C, (per + 4095) & ~(size_t)4095));
    P.chunk_bytes = C;
    P.slack = std::min<size_t>((size_t)1 << 20, C);
    P.n_chunks = (P.file_len + C - 1) / C;
    P.chunks.resize(P.n_chunks);
    P.in_max = (size_t)(n_readers + n_ctx + 1);
    P.out_max = (size_t)(n_ctx + n_writers + 1);
    cudaSetDevice(P.lead->device);
    int rc = P.n_chunks ? run(P, n_readers, n_writers) : VCFC_OK;
    if (ftruncate(P.ofd, (off_t)(P.out_base + P.out_total)) != 0 && rc == VCFC_OK) rc = VCFC_E_IO;
    if (index_path && rc == VCFC_OK) {
        // what create_binned_index4 checks before its walk: the header region of the compressed file (main.cpp:1304)
        std::vector<uint8_t> head(std::min<size_t>(P.out_total, (size_t)64 << 20));
        size_t got = 0, hlen = 0;
        uint64_t sc = 0;
        rc = pread_all(P.ofd, head.data(), head.size(), 0, &got);
        if (rc == VCFC_OK) rc = vcfc_parse_headers(head.data(), got, &hlen, &sc);
        std::vector<uint8_t> entries;
        if (rc == VCFC_OK) rc = build_index(P, entries_per_bin, entries, n_entries);
        if (rc == VCFC_OK) {
            int xfd = open(index_path, O_CREAT | O_TRUNC | O_WRONLY, 0644);
            if (xfd < 0) rc = VCFC_E_IO;
            else {
                rc = pwrite_all(xfd, entries.data(), entries.size(), 0);
                if (close(xfd) != 0 && rc == VCFC_OK) rc = VCFC_E_IO;
            }
        }
    }
    for (auto& c : P.chunks) { delete c.idx; c.idx = nullptr; delete c.hash; c.hash = nullptr; }
    close(P.ifd);
    if (close(P.ofd) != 0 && rc == VCFC_OK) rc = VCFC_E_IO;
    pin_trim(P.lead, env_sz("VCFC_PIN_KEEP_MB", 4096) << 20);
    return rc;
}

int vcfc_compress_file_multi(vcfc_ctx** ctxs, int n_ctx, const char* in_path, const char* out_path) {
    return compress_file_impl(ctxs, n_ctx, in_path, out_path, nullptr, 0, nullptr);
}

int vcfc_compress_index_file_multi(vcfc_ctx** ctxs, int n_ctx, const char* in_path, const char* out_path, const char* index_path,
                                   uint64_t entries_per_bin, uint64_t* n_entries) {
    if (!index_path) return VCFC_E_ARG;
    return compress_file_impl(ctxs, n_ctx, in_path, out_path, index_path, entries_per_bin, n_entries);
}

int vcfc_decompress_file_multi(vcfc_ctx** ctxs_in, int n_ctx_in, const char* in_path, const char* out_path) {
    if (!ctxs_in || n_ctx_in <= 0 || !ctxs_in[0] || !in_path || !out_path) return VCFC_E_ARG;
    for (int i = 0; i < n_ctx_in; i++) if (!ctxs_in[i]) return VCFC_E_ARG;
    std::vector<vcfc_ctx*> all;
    with_twins(ctxs_in, n_ctx_in, all);
    vcfc_ctx** ctxs = all.data();
    const int n_ctx = (int)all.size();
    Pipe P;
    P.ctxs = ctxs; P.n_ctx = n_ctx; P.lead = ctxs[0]; P.encode = false;
    P.ifd = open(in_path, O_RDONLY);
    if (P.ifd < 0) return VCFC_E_IO;
    struct stat st;
    if (fstat(P.ifd, &st) != 0) { close(P.ifd); return VCFC_E_IO; }
    P.file_len = (size_t)st.st_size;
    P.ofd = open(out_path, O_CREAT | O_TRUNC | O_RDWR, 0644);           // the reference truncates first (compress.cpp:1217)
    if (P.ofd < 0) { close(P.ifd); return VCFC_E_IO; }
    P.map_writes = !getenv("VCFC_NO_MAP_WRITES");
    // header region (compress.cpp:1108-1211): read until it parses or the file ends
    int rc = VCFC_OK;
    {
        std::vector<uint8_t> head;
        size_t want = (size_t)1 << 20, hlen = 0;
        for (;;) {
            want = std::min(want, P.file_len);
            head.resize(want);
            size_t got = 0;
            if ((rc = pread_all(P.ifd, head.data(), want, 0, &got))) break;
            // the header region is complete once a line that does not start with '#' begins inside what was read
            rc = vcfc_parse_headers(head.data(), got, &hlen, &P.sample_count);
            if (rc == VCFC_OK || want >= P.file_len) break;
            want *= 4;
        }
        if (rc == VCFC_OK) rc = pwrite_all(P.ofd, head.data(), hlen, 0);
        if (rc != VCFC_OK) { close(P.ifd); close(P.ofd); return rc; }
        P.data_off = hlen;
        P.out_base = hlen;
        P.file_size = hlen;
    }
    const int n_writers = n_threads_default("VCFC_WRITERS", std::min(12, std::max(2, (int)std::thread::hardware_concurrency() * 3 / 4)));   // the text side is the wide one
    size_t C = env_sz("VCFC_FILE_DCHUNK_MB", 4) << 20;
    const size_t per = (P.file_len - P.data_off) / (size_t)(4 * n_ctx) + 1;
    C = std::max<size_t>((size_t)256 << 10, std::min(C, (per + 4095) & ~(size_t)4095));
    P.chunk_bytes = C;
    P.slack = std::min<size_t>((size_t)1 << 20, C);
    P.in_max = (size_t)(n_ctx + 3);
    P.out_max = (size_t)(n_ctx + n_writers + 1);
    cudaSetDevice(P.lead->device);
    rc = run(P, 1, n_writers);
    if (ftruncate(P.ofd, (off_t)(P.out_base + P.out_total)) != 0 && rc == VCFC_OK) rc = VCFC_E_IO;
    close(P.ifd);
    if (close(P.ofd) != 0 && rc == VCFC_OK) rc = VCFC_E_IO;
    pin_trim(P.lead, env_sz("VCFC_PIN_KEEP_MB", 4096) << 20);
    return rc;
}

int vcfc_compress_file(vcfc_ctx* ctx, const char* in_path, const char* out_path) {
    return vcfc_compress_file_multi(&ctx, 1, in_path, out_path);
}
int vcfc_decompress_file(vcfc_ctx* ctx, const char* in_path, const char* out_path) {
    return vcfc_decompress_file_multi(&ctx, 1, in_path, out_path);
}

}  // extern "C"
