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

// Weight-gradient GEMM on the tcgen05 tensor cores (sqair_wgrad_tc.cu): dW[k, n] += sum_m X[m, k] dY[m, n].
// Row m = z * ny + y of an operand starts at p + z * outer + y * inner (floats); ny <= 1: p + m * outer.
struct TcOperand {
    const float* p;
    int64_t outer, inner;
    int ny;
};
// operands TMA can describe (16-byte aligned base and strides), a reduction block that fits the pipeline, a tile worth it
bool wgrad_tc_supported(const TcOperand& x, const TcOperand& dy, int M, int K, int N);
// adds into dw (row stride ldw): zero or initialise it first
int wgrad_tc(const TcOperand& x, const TcOperand& dy, float* dw, int ldw, int M, int K, int N, cudaStream_t st);

}  // namespace sqi
