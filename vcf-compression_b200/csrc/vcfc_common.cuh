// vcfc_common.cuh -- shared device helpers (sm_100a): mbarrier / TMA PTX wrappers, the
// 128B-swizzle address map, decoupled look-back status words.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vcfc {

// ---- token byte layout (reference: src/utils.hpp:44-55) ----
constexpr uint32_t kTok00 = 0x00, kTok11 = 0x80, kTok01 = 0xA0, kTok10 = 0xC0, kTokLit = 0xE0;

// ---- error codes mirrored from include/vcfc_gpu.h ----
constexpr int kOk = 0, kETooFew = 1, kEEightCols = 2, kECap = 3, kEFormat = 4, kETrunc = 5,
              kELine2Big = 10;

struct DevResult {          // layout == vcfc_result
    int32_t status;
    int32_t reserved;
    unsigned long long out_len;
    unsigned long long n_lines;
    unsigned long long err_line;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- TMA: 2-D tiled load, 128B swizzle, completion on an mbarrier (SASS: UTMALDG) ----
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// Rows of 128 bytes, CU_TENSOR_MAP_SWIZZLE_128B: inside every 1 KiB block the 16-byte unit
// index is XORed with (row & 7).  `off` is the linear byte offset inside a 1 KiB-aligned buffer.
__device__ __forceinline__ uint32_t swz(uint32_t off) {
    return off ^ (((off >> 7) & 7u) << 4);
}

// ---- decoupled look-back status words: [63:62] flag, [61:0] value ----
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62, kValMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// exact per-byte zero test: 0x80 in every byte of x that is zero
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) {
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
// gather the four 0x80 flags of z into a 4-bit value (bit i <- byte i)
__device__ __forceinline__ uint32_t nibble_of(uint32_t z) {
    return (z * 0x00204081u) >> 28;
}

}  // namespace vcfc
