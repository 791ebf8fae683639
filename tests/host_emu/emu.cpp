// Host emulation of the CUDA kernel's per-block program (TEST INFRASTRUCTURE ONLY).
//
// Compiles sqair_b200/csrc/sqair_device.cuh with SQAIR_HOST_EMU: every thread block becomes one
// sequential "thread" (tid 0 of 1), barriers are no-ops.  This lets the kernel's arithmetic,
// indexing and slot bookkeeping be checked against the oracle on a machine without a GPU.  It is
// never loaded by the sqair_b200 package; the product path is the CUDA library only.
#define SQAIR_HOST_EMU 1
#include <stdlib.h>
#include <algorithm>
#include <vector>
struct float4 { float x, y, z, w; };
#include "../../sqair_b200/csrc/sqair_device.cuh"

using namespace sq;

template <int R>
static void run_blocks(const Plan& plan, const Job& job) {
    const int nblk = (plan.rows + R - 1) / R;
    std::vector<float> smem(plan.sm.total + 64);
    for (int b = 0; b < nblk; ++b) {
        std::fill(smem.begin(), smem.end(), 1e30f);     // poison: catches reads of unwritten scratch
        Ctx c{0, 1, 0, 1, 0, 1, smem.data()};
        Block<R> blk(c, plan, job, b * R);
        blk.run();
    }
}

extern "C" int emu_smem_floats(const sqair_cfg* cfg, int R) {
    auto tab = param_table(*cfg);
    Plan plan;
    if (!build_plan(*cfg, R, plan, tab).empty()) return -1;
    return plan.sm.total;
}

extern "C" int emu_forward(const sqair_cfg* cfg, const float* params, const float* obs,
                           const float* eps_where, const float* eps_what, const float* u_pres,
                           const sqair_outputs* out, int R) {
    std::string e = validate_cfg(*cfg);
    if (!e.empty()) { fprintf(stderr, "emu: %s\n", e.c_str()); return -1; }
    auto tab = param_table(*cfg);
    std::vector<float> packed(packed_floats(tab), 0.f);
    for (const auto& p : tab)
        for (int64_t i = 0; i < p.count; ++i) packed[p.packed_offset + i] = params[p.offset + i];
    Plan plan;
    e = build_plan(*cfg, R, plan, tab);
    if (!e.empty()) { fprintf(stderr, "emu: %s\n", e.c_str()); return -1; }
    Job job{packed.data(), obs, eps_where, eps_what, u_pres, *out};
    switch (R) {
        case 1: run_blocks<1>(plan, job); break;
        case 2: run_blocks<2>(plan, job); break;
        case 3: run_blocks<3>(plan, job); break;
        case 4: run_blocks<4>(plan, job); break;
        case 5: run_blocks<5>(plan, job); break;
        case 8: run_blocks<8>(plan, job); break;
        default: fprintf(stderr, "emu: unsupported R=%d\n", R); return -2;
    }
    return 0;
}
