// vcfc_common.cuh -- shared device definitions (sm_100a): token byte layout, error codes, the result
// block, word-parallel byte tests.
#pragma once
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

// ---- word-parallel byte tests (the kernels scan for tabs, newlines and token markers four bytes at a time) ----
// exact per-byte zero test: 0x80 in every byte of x that is zero
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) {
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
// gather the four 0x80 flags of z into a 4-bit value (bit i <- byte i)
__device__ __forceinline__ uint32_t nibble_of(uint32_t z) {
    return (z * 0x00204081u) >> 28;
}

}  // namespace vcfc
