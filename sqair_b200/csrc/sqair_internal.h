// sqair_internal.h -- declarations shared by the translation units of libsqair_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "sqair_core.h"

namespace sqi {

int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define CUDA_TRY(x)                                       \
    do {                                                  \
        cudaError_t e_ = (x);                             \
        if (e_ != cudaSuccess) return sqi::cuda_fail(e_, #x); \
    } while (0)

// The launch shape the library picked for a configuration: R rows per cluster of C blocks, the frame plan built for
// it and the packing table.  Cached per configuration; see choose_shape in sqair_api.cu.
struct Shape {
    sq::Plan plan;
    std::vector<sq::Piece> pieces;
    int64_t packed_total = 0;
    int R = 0, C = 0;
};

// "" on success, else an error message
std::string choose_shape(const sqair_cfg& c, const std::vector<sq::ParamEntry>& tab, Shape& out);
int env_int(const char* name);

}  // namespace sqi
