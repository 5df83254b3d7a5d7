// vcfc -- command line with the reference's verb surface for the hot path
// (/root/reference/src/main.cpp:4028-4184):
//     vcfc compress   IN.vcf  OUT.vcfc
//     vcfc decompress IN.vcfc OUT.vcf
//     vcfc query      IN.vcfc REF[:START-END]
//     vcfc create-binned-index BIN_SIZE IN.vcfc          (writes IN.vcfc.vcfci, main.cpp:4097-4115)
//     vcfc query-binned-index  IN.vcfc REF:START-END     (reads IN.vcfc.vcfci, main.cpp:4117-4143)
//     vcfc sparsify IN.vcfc OUT.sparse | sparse-query IN.sparse REF:START-END   (main.cpp:4073-4096)
// Host C++ only; all coding work is done by libvcfc_gpu.so through its C ABI.  Where the
// reference lets an exception escape (abort, exit 134) this prints the reason and exits 1.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "vcfc_gpu.h"

// VCFC_TRACE=1: wall-clock of the process's stages on stderr (context creation dominates a one-shot CLI run)
static double t_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static const double t_start = t_now();
static void trace(const char* what) {
    static const bool on = getenv("VCFC_TRACE") != nullptr;
    if (on) fprintf(stderr, "[vcfc %8.3f s] %s\n", t_now() - t_start, what);
}

static int usage() {
    fprintf(stderr,
            "usage: vcfc compress IN.vcf OUT.vcfc | decompress IN.vcfc OUT.vcf | query IN.vcfc REF[:START-END]\n"
            "       vcfc create-binned-index BIN_SIZE IN.vcfc | query-binned-index IN.vcfc REF:START-END\n"
            "       vcfc sparsify IN.vcfc OUT.sparse | sparse-query IN.sparse REF:START-END\n"
            "       env: VCFC_DEVICE (default 0), VCFC_GPUS (compress / decompress over that many GPUs starting at VCFC_DEVICE,\n"
            "            default 1, \"all\" = every GPU of the box), VCFC_FILE_CHUNK_MB (default 16),\n"
            "            VCFC_INDEX_BIN=N (compress also writes OUT.vcfci with N lines per bin, in the same pass)\n");
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 2) return usage();
    std::string action(argv[1]);
    const char* verbs_elsewhere[] = {"gap-analysis", "create-sparse-index", "query-sparse-index"};
    for (const char* v : verbs_elsewhere)
        if (action == v) {
            fprintf(stderr, "vcfc: verb '%s' is outside the GPU hot path; use the reference binary for it\n", v);
            return 2;
        }
    if (action == "sparsify") {                                  // main.cpp:4073-4085 (host only: no device needed)
        if (argc < 4) return usage();
        if (strcmp(argv[2], argv[3]) == 0) {
            fprintf(stderr, "input and output file are the same\n");
            return 1;
        }
        if (access(argv[2], F_OK) != 0) printf("Input file does not exist: %s\n", argv[2]);
        const int src = vcfc_sparsify_file(argv[2], argv[3]);
        if (src != VCFC_OK) fprintf(stderr, "vcfc sparsify: %s\n", vcfc_strerror(src));
        return src == VCFC_OK ? 0 : 1;
    }
    if (action == "create-binned-index") {
        if (argc != 4) {
            printf("Usage: ./main create-binned-index <bin-size> <compressed-filename>\n");   // main.cpp:4098-4101
            return 1;
        }
        char* endp = nullptr;
        const unsigned long bin = strtoul(argv[2], &endp, 10);                               // str_to_uint64, utils.cpp:152-165
        if (*endp != 0 || bin == 0) {
            printf("bin size must be a positive integer\n");
            return 1;
        }
        int dev = getenv("VCFC_DEVICE") ? atoi(getenv("VCFC_DEVICE")) : 0;
        vcfc_ctx* ictx = nullptr;
        int irc = vcfc_gpu_init(dev, &ictx);
        if (irc != VCFC_OK) {
            fprintf(stderr, "vcfc: cannot use CUDA device %d: %s (there is no CPU path)\n", dev, vcfc_strerror(irc));
            return 1;
        }
        const std::string index_path = std::string(argv[3]) + ".vcfci";                       // VCFC_BINNING_INDEX_EXTENSION, utils.hpp:29
        irc = vcfc_create_binned_index_file(ictx, argv[3], index_path.c_str(), bin, nullptr);
        if (irc != VCFC_OK) fprintf(stderr, "vcfc create-binned-index: %s\n", vcfc_strerror(irc));
        vcfc_gpu_destroy(ictx);
        return irc == VCFC_OK ? 0 : 1;
    }
    if (action != "compress" && action != "decompress" && action != "query" && action != "query-binned-index" && action != "sparse-query") {
        printf("Unknown action name: %s\n", action.c_str());   // main.cpp:4181-4183
        return 0;
    }
    if (argc < 4) return usage();
    const char* in = argv[2];
    if (access(in, F_OK) != 0) printf("Input file does not exist: %s\n", in);   // main.cpp:4040-4042
    int dev = getenv("VCFC_DEVICE") ? atoi(getenv("VCFC_DEVICE")) : 0;
    vcfc_ctx* ctx = nullptr;
    trace("start");
    int rc = vcfc_gpu_init(dev, &ctx);
    if (rc != VCFC_OK) {
        fprintf(stderr, "vcfc: cannot use CUDA device %d: %s (there is no CPU path)\n", dev, vcfc_strerror(rc));
        return 1;
    }
    trace("context ready");
    if (action == "query" || action == "query-binned-index" || action == "sparse-query") {
        rc = action == "query" ? vcfc_query_file(ctx, in, argv[3], STDOUT_FILENO)
             : action == "sparse-query" ? vcfc_sparse_query_file(ctx, in, argv[3], STDOUT_FILENO)     // main.cpp:4086-4096
                                        : vcfc_query_binned_index_file(ctx, in, argv[3], STDOUT_FILENO);
        if (rc == VCFC_E_QUERY) printf("Failed to parse query string: %s\n", argv[3]);   // main.cpp:4064-4067
    } else {
        if (strcmp(in, argv[3]) == 0) {
            fprintf(stderr, "input and output file are the same\n");   // main.cpp:4044-4046
            vcfc_gpu_destroy(ctx);
            return 1;
        }
        // one context per GPU: line-block chunks are handed out in file order, outputs concatenated by offsets
        std::vector<vcfc_ctx*> ctxs{ctx};
        const char* g = getenv("VCFC_GPUS");
        int want = 1;
        if (g && *g) want = strcmp(g, "all") == 0 ? 1 << 20 : atoi(g);
        for (int d = dev + 1; (int)ctxs.size() < want; d++) {
            vcfc_ctx* c = nullptr;
            if (vcfc_gpu_init(d, &c) != VCFC_OK) break;              // fewer GPUs than asked for: use what is there
            ctxs.push_back(c);
        }
        const char* xb = getenv("VCFC_INDEX_BIN");                 // compress + create-binned-index in one pass: OUT.vcfci beside OUT
        if (action == "compress" && xb && atol(xb) > 0) {
            const std::string index_path = std::string(argv[3]) + ".vcfci";
            rc = vcfc_compress_index_file_multi(ctxs.data(), (int)ctxs.size(), in, argv[3], index_path.c_str(), (uint64_t)atol(xb), nullptr);
        } else {
            rc = action == "compress" ? vcfc_compress_file_multi(ctxs.data(), (int)ctxs.size(), in, argv[3])
                                      : vcfc_decompress_file_multi(ctxs.data(), (int)ctxs.size(), in, argv[3]);
        }
        for (size_t k = 1; k < ctxs.size(); k++) vcfc_gpu_destroy(ctxs[k]);
    }
    trace("verb done");
    if (rc != VCFC_OK) {
        fprintf(stderr, "vcfc %s: %s", action.c_str(), vcfc_strerror(rc));
        if (rc == VCFC_E_CUDA) fprintf(stderr, " [%s]", vcfc_last_cuda_error(ctx));
        fprintf(stderr, "\n");
    }
    vcfc_gpu_destroy(ctx);
    trace("context destroyed");
    return rc == VCFC_OK ? 0 : 1;
}
