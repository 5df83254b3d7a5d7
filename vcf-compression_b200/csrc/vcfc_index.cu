// vcfc_index.cu -- binned index (.vcfci): per-line END position and reference index on the device.
//
// Restates, does not copy, the per-line part of create_binned_index4 (/root/reference/src/main.cpp:1284-1637):
// CHROM, POS, REF, ALT and INFO are read from the line's required section, END comes from compute_end_position
// (main.cpp:763-852: INFO END / SVLEN for symbolic alleles, else POS + max(len(REF), longest ALT) - 1), the
// chromosome index from reference_name_map (utils.hpp:90-103).  One thread per compressed line; the line table
// is the decoder's (k_dec_walk ... k_dec_fill).  The bin rule itself (a sequential grow-or-open over the lines,
// main.cpp:1430-1470) runs on the host over 10 bytes per line: see vcfc_create_binned_index_file.
#include "vcfc_common.cuh"
#include "vcfc_index.cuh"
#include "vcfc_internal.h"

namespace vcfc {
namespace idx {

// err: 0 ok, 1 malformed (the reference throws), 2 ran off the end of the block while reading the columns
__global__ void k_index_lines(const uint8_t* __restrict__ in, unsigned long long n, const unsigned long long* __restrict__ line_start,
                              unsigned long long n_lines, long long* __restrict__ end_pos, uint8_t* __restrict__ ref_idx,
                              uint8_t* __restrict__ err) {
    const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines) return;
    unsigned long long q = line_start[k] + 8;
    const uint8_t* f[8];
    int fl[8];
    for (int c = 0; c < 8; c++) {                            // read_to(..., '\t') x 8 (main.cpp:1371-1405)
        unsigned long long e = q;
        while (e < n && in[e] != '\t') e++;
        if (e >= n || e - q > 0x3fffffffull) { end_pos[k] = 0; ref_idx[k] = 0; err[k] = 2; return; }
        f[c] = in + q; fl[c] = (int)(e - q);
        q = e + 1;
    }
    long long pos = 0, endp = 0;
    bool ok = parse_ul(f[1], fl[1], &pos);
    if (ok) ok = line_end_position(pos, fl[3], f[4], fl[4], f[7], fl[7], &endp);
    if (!ok) endp = 0;
    end_pos[k] = endp;
    ref_idx[k] = ref_name_index(f[0], fl[0]);
    err[k] = ok ? 0 : 1;
}

}  // namespace idx

int index_line_ends(vcfc_ctx* ctx, const uint8_t* d_in, size_t in_len, const unsigned long long* d_line_start,
                    unsigned long long n_lines, long long* d_end, uint8_t* d_ref, uint8_t* d_err, cudaStream_t stream) {
    if (n_lines == 0) return VCFC_OK;
    idx::k_index_lines<<<(unsigned)((n_lines + 127) / 128), 128, 0, stream>>>(d_in, (unsigned long long)in_len, d_line_start, n_lines,
                                                                              d_end, d_ref, d_err);
    ctx->launches++;
    VCFC_CUDA(ctx, cudaGetLastError());
    return VCFC_OK;
}

}  // namespace vcfc
