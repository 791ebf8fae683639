// sqair_backward.h -- the backward pass of the SQAIR hot path (reference: `opt.compute_gradients(target)` in
// sqair/model.py:150-168, i.e. TF autodiff through seq.py:181-276, sqair_modules.py:446-582, core.py:164-359,
// propagate.py:68-184, modules.py:39-644, prior.py:61-102, index.py:132-221; gradient-flow facts: SURVEY Appendix F).
//
// Design.  The forward kernel (sqair_device.cuh) writes every activation the adjoint needs into a row-major global
// "stash" (sqair_core.h: build_stash).  The backward pass walks the frames in reverse; within a frame it follows the
// forward program backwards.  Rows (b, k) are independent, so every step is batched over all rows:
//   * dense layers:  dX = dY . W^T  (`dgrad`, M = rows or rows x slots) on the layer's unpadded "virtual matrix"
//     (LayerB), with the activation derivative applied while loading dY and the input segments scattered /
//     accumulated straight into the gradient buffers of the signals they came from;
//   * everything else (samplers, log-probabilities, GRU gate algebra, resamplers, slot compaction, the f64 count
//     posterior, prior post-processing) in hand-written per-row stages (`bw_stage<STAGE>`: one thread block per row);
//   * every layer's pre-activation gradient dY is kept ([T, rows, entries, N]) so that all weight gradients are
//     computed at the end as large GEMMs  dW = X^T . dY  over M = T x rows x slots (`wgrad`), bias gradients as column
//     sums, and scattered back to the reference's variable layout (`sqair_param_layout`).
// Single source for the CUDA library (sqair_api.cu supplies the kernels) and for the host emulator (tests/host_emu:
// sequential loops) that checks the adjoint arithmetic against torch autograd on the oracle without a GPU.
#pragma once
#include <math.h>
#include "sqair_core.h"

#ifdef __CUDACC__
#define SQB_HD __host__ __device__ __forceinline__
#else
#define SQB_HD inline
#endif

namespace sq {

// element (m, j) of a batched operand: p[(m / ny) * outer + (m % ny) * inner + j]
struct Addr {
    float* p;
    int outer, inner;
};
SQB_HD Addr mk_addr(const float* p, int outer, int inner) {
    Addr a;
    a.p = const_cast<float*>(p); a.outer = outer; a.inner = inner;
    return a;
}

constexpr int BW_MAXSEG = 6;
enum { SEGM_SKIP = 0, SEGM_STORE = 1, SEGM_ACC = 2 };

// dX_seg[m, k - k0] (=|+=) sum_n A[m, n] * W[k, n],  A[m, n] = a[m, n] * act'(y[m, n]);  A is also stored to `dy`
struct DgradArgs {
    int layer;              // LayerId (diagnostics only)
    int M, N, K, ny;
    Addr a, y, dy;          // y.p == nullptr: no activation derivative; dy.p == nullptr: A is not stored
    int act;
    float act_scale, act_add;
    const float* w;         // [K (+1), N] row-major (LayerB virtual matrix)
    int nseg;
    struct Seg {
        int k0, k1, mode;
        Addr d;
    } seg[BW_MAXSEG];
};

// dW[k, n] += sum_m X[m, k] * dY[m, n]
struct WgradArgs {
    int M, K, N, ny;
    Addr x, dy;
    float* dw;
    int ldw;
};

// out[n] += sum_m dY[m, n]
struct ColsumArgs {
    int M, N, ny;
    Addr dy;
    float* out;
};

// activation derivative expressed through the activation's OUTPUT y (what the stash holds)
SQB_HD float act_deriv(int act, float y, float scale, float add) {
    switch (act) {
        case ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
        case ACT_TANH: return 1.f - y * y;
        case ACT_SIGMOID: { const float s = y / scale; return scale * s * (1.f - s); }
        case ACT_SOFTPLUS: return -expm1f(-(y - add));       // sigmoid(x) = 1 - exp(-softplus(x))
        default: return 1.f;
    }
}

// dY buffers beyond one per layer: per-row quantities whose column sums are parameter gradients
enum ExtraId { X_PH0, X_DH0, X_T0, X_P0, X_MEAN, X_RNINIT, X_RNSAMPLE, X_COUNT };
// small parameter gradients accumulated with atomics (a few adds per row and stage)
enum SmallId { SM_CHOL = 0, SM_DSO = 10, SM_PSO = 11, SM_OUTSCALE = 12, SM_COUNT = 16 };

struct BwdCtx {
    sqair_cfg cfg;
    int T, rows, n, nw, nh, g, hs, PX, zw, recw, npri;
    RecF rec;
    POff po;                 // CANONICAL offsets of the non-matrix parameters
    int vimco;               // 1: VIMCO target, 0: -elbo_iwae (model.py:152-156)
    const float* prm;        // canonical flat parameters
    const float* bw;         // backward parameter buffer (virtual matrices)
    int64_t bw_off[L_COUNT];
    int bw_nu[L_COUNT], bw_ku[L_COUNT];
    const float* obs;
    const float* eps_where;
    const float* eps_what;
    const float* stash;
    Sig st[S_COUNT];
    const float* gw;         // [rows] d target / d (sum_t log w)
    const float* gp;         // [rows] d target / d (sum_t discrete log prob)
    float* dy[L_COUNT];      // [T, rows, dy_e, dy_w]
    int dy_e[L_COUNT], dy_w[L_COUNT];
    float* xt[X_COUNT];      // [T, rows, width]
    float* small;            // [SM_COUNT]
    // gradient buffers of one frame, [rows, ...]
    float *gZc, *gTc, *gPc;              // w.r.t. the state LEAVING the frame (what / where / . / plogit; GRU states)
    float *gZo, *gTo, *gPo;              // w.r.t. the state ENTERING the frame
    float *gPropRec, *gDiscRec;          // [n+1, zw]
    float *gTnew, *gPnew;                // [n, nh]
    float *gPH, *gDH;                    // [n+1, nh]
    float *gDIn, *gExp, *gHrn;           // [2nh], [1], [128]
    float *gPri;                         // [n, npri]
    float *gMask, *gGlm;                 // [n, g]
    float *gLoc1;                        // [n, nw]
    float *gHwbmk;                       // [n, 256]
    float *gEnc;                         // [2nw]
    float *gRH;                          // [n, nh]
    float *tA0, *tA1;                    // [n, nh]
};

SQB_HD const float* sgp(const BwdCtx& c, int sig, int t, int row, int e) {
    const Sig g = c.st[sig];
    return c.stash + (size_t)g.off + ((size_t)(t * c.rows + row) * g.entries + e) * g.width;
}
SQB_HD float* dyp(const BwdCtx& c, int l, int t, int row, int e) {
    return c.dy[l] + ((size_t)(t * c.rows + row) * c.dy_e[l] + e) * c.dy_w[l];
}
SQB_HD const float* wmat(const BwdCtx& c, int l) { return c.bw + c.bw_off[l]; }
SQB_HD float bsigmoid(float x) { return 1.f / (1.f + expf(-x)); }
SQB_HD float blin11(int i, int n) {
    const float step = 2.f / (float)(n - 1);
    return (i < n / 2) ? (-1.f + step * (float)i) : (1.f - step * (float)(n - 1 - i));
}

#if defined(__CUDA_ARCH__)
#define SQB_AADD(p, v) atomicAdd((p), (v))
#else
#define SQB_AADD(p, v) (*(p) += (v))
#endif

// bilinear sample with zero padding (tf.contrib.resampler) and its position derivatives
struct BilinG {
    float val, ddx, ddy;
    float w4[4];
    int idx4[4];
};
template <class Fetch>
SQB_HD BilinG bilin_grad(float x, float y, int w, int h, Fetch fetch) {
    BilinG r;
    r.val = r.ddx = r.ddy = 0.f;
    for (int q = 0; q < 4; ++q) { r.w4[q] = 0.f; r.idx4[q] = -1; }
    if (!(x > -1.f && y > -1.f && x < (float)w && y < (float)h)) return r;
    const float fx = floorf(x), fy = floorf(y);
    const float dx = fx + 1.f - x, dy = fy + 1.f - y;
    const int ifx = (int)fx, ify = (int)fy, icx = ifx + 1, icy = ify + 1;
    const bool fx_ok = ifx >= 0 && ifx <= w - 1, cx_ok = icx >= 0 && icx <= w - 1;
    const bool fy_ok = ify >= 0 && ify <= h - 1, cy_ok = icy >= 0 && icy <= h - 1;
    const float v00 = (fx_ok && fy_ok) ? fetch(ifx, ify) : 0.f;
    const float v11 = (cx_ok && cy_ok) ? fetch(icx, icy) : 0.f;
    const float v01 = (fx_ok && cy_ok) ? fetch(ifx, icy) : 0.f;
    const float v10 = (cx_ok && fy_ok) ? fetch(icx, ify) : 0.f;
    r.val = dx * dy * v00 + (1.f - dx) * (1.f - dy) * v11 + dx * (1.f - dy) * v01 + (1.f - dx) * dy * v10;
    r.ddx = dy * (v10 - v00) + (1.f - dy) * (v11 - v01);
    r.ddy = dx * (v01 - v00) + (1.f - dx) * (v11 - v10);
    if (fx_ok && fy_ok) { r.w4[0] = dx * dy; r.idx4[0] = ify * w + ifx; }
    if (cx_ok && cy_ok) { r.w4[1] = (1.f - dx) * (1.f - dy); r.idx4[1] = icy * w + icx; }
    if (fx_ok && cy_ok) { r.w4[2] = dx * (1.f - dy); r.idx4[2] = icy * w + ifx; }
    if (cx_ok && fy_ok) { r.w4[3] = (1.f - dx) * dy; r.idx4[3] = ify * w + icx; }
    return r;
}

// tfd.Normal.log_prob(x; mu, sd) partial derivatives: d/dx = -z/sd, d/dmu = z/sd, d/dsd = (z^2 - 1)/sd
struct NormG {
    float dx, dmu, dsd;
};
SQB_HD NormG normal_lp_grad(float x, float mu, float sd) {
    const float z = (x - mu) / sd;
    NormG g;
    g.dx = -z / sd; g.dmu = z / sd; g.dsd = (z * z - 1.f) / sd;
    return g;
}

// ---------------------------------------------------------------------------------------------
// Row stages.  `EX` gives the threads that cooperate on one row: ex.tid / ex.nt, ex.sync(), ex.sum4() (sum over the
// row's threads, result in every thread) and ex.scratch (floats shared by the row's threads).
// ---------------------------------------------------------------------------------------------
enum StageId {
    BS_CANVAS, BS_COMPACT, BS_DISC_POST, BS_DISC_A, BS_DISC_B, BS_DISC_C, BS_LAT_PRE,
    BS_PROP_A, BS_PROP_B, BS_PROP_C, BS_PROP_D, BS_PROP_E, BS_PROP_F,
    BS_STN1, BS_PRIOR_PRE, BS_PGRU_A, BS_PGRU_B, BS_FINAL_STATES, BS_COUNT
};

// Gradient of a (masked) glimpse w.r.t. the where-logits (modules.py:165-172,204-227: AffineGridWarper + resampler
// warp gradient, to_coords, straight-through scale clip); optionally accumulates the mask gradient dglm * raw glimpse.
template <class EX>
SQB_HD void stn_where_grad(const BwdCtx& c, EX& ex, const float* img, const float (&wl)[4], const float* dglm,
                           const float* mask, float* gmask, float (&out)[4]) {
    const int G = c.cfg.G, W = c.cfg.W, H = c.cfg.H;
    const float s0 = bsigmoid(wl[0]), s1 = bsigmoid(wl[1]);
    const float sx = fmaxf(s0, 1e-4f), sy = fmaxf(s1, 1e-4f);
    const float tx = tanhf(wl[2]), ty = tanhf(wl[3]);
    const float hw = 0.5f * (float)(W - 1), hh = 0.5f * (float)(H - 1);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = ex.tid; i < G * G; i += ex.nt) {
        const int gx = i % G, gy = i / G;
        const float u = blin11(gx, G), v = blin11(gy, G);
        const float x = hw * (sx * u + tx) + hw, y = hh * (sy * v + ty) + hh;
        const BilinG b = bilin_grad(x, y, W, H, [&](int ix, int iy) { return img[iy * W + ix]; });
        float dg = dglm[i];
        if (mask) {
            if (gmask) gmask[i] += dg * b.val;
            dg *= mask[i];
        }
        acc[0] += dg * b.ddx * hw * u; acc[1] += dg * b.ddy * hh * v;
        acc[2] += dg * b.ddx * hw; acc[3] += dg * b.ddy * hh;
    }
    ex.sum4(acc);
    out[0] = acc[0] * s0 * (1.f - s0);          // straight-through clip: the gradient ignores the 1e-4 floor
    out[1] = acc[1] * s1 * (1.f - s1);
    out[2] = acc[2] * (1.f - tx * tx);
    out[3] = acc[3] * (1.f - ty * ty);
}

// L = fill_triangular(cholesky_scale) * scale[:, None] + diag(scale) (modules.py:535-545); TF element order
SQB_HD int chol_idx(int i, int j) {
    const int idx[4][4] = {{4, -1, -1, -1}, {8, 9, -1, -1}, {7, 6, 5, -1}, {3, 2, 1, 0}};
    return idx[i][j];
}

template <class EX>
SQB_HD void bw_canvas(const BwdCtx& c, EX& ex, int t, int row) {
    // AIRDecoder._decode/_add_mean_image + pixel likelihood backward (modules.py:435-467; seq.py:272-273)
    const int n = c.n, g = c.g, G = c.cfg.G, W = c.cfg.W, H = c.cfg.H, nw = c.nw;
    float* sg = ex.scratch;                // glimpses [n][g]
    float* dgl = sg + n * g;               // their gradient accumulators
    float* cc = dgl + n * g;               // coords [n][7]
    float* wsum = cc + n * 7;              // where sums [n][4]
    const float* gl = sgp(c, S_DGL, t, row, 0);
    const float* Z = sgp(c, S_Z, t + 1, row, 0);
    for (int i = ex.tid; i < n * g; i += ex.nt) { sg[i] = gl[i]; dgl[i] = 0.f; }
    for (int s = ex.tid; s < n; s += ex.nt) {
        const float* wl = Z + s * c.zw + nw;
        const float s0 = bsigmoid(wl[0]), s1 = bsigmoid(wl[1]);
        cc[s * 7 + 0] = fmaxf(s0, 1e-4f); cc[s * 7 + 1] = fmaxf(s1, 1e-4f);
        cc[s * 7 + 2] = tanhf(wl[2]); cc[s * 7 + 3] = tanhf(wl[3]);
        cc[s * 7 + 4] = Z[s * c.zw + nw + 4];
        cc[s * 7 + 5] = s0 * (1.f - s0); cc[s * 7 + 6] = s1 * (1.f - s1);
    }
    for (int i = ex.tid; i < n * 4; i += ex.nt) wsum[i] = 0.f;
    ex.sync();
    const float hg = 0.5f * (float)(G - 1);
    const float sf0 = sqrtf(c.cfg.output_std), sb0 = sqrtf(c.cfg.bg_std);
    const float sf = sf0 * sf0, sb = sb0 * sb0;
    const float up = c.gw[row];
    const float* mean_img = c.prm + c.po.mean_img;
    const float* img = c.obs + ((size_t)t * c.cfg.B + row / c.cfg.K) * c.PX;
    float* dmean = c.xt[X_MEAN] + (size_t)(t * c.rows + row) * c.PX;
    for (int px0 = 0; px0 < c.PX; px0 += ex.nt) {            // uniform trip count: the where sums are reduced per warp
        const int px = px0 + ex.tid;
        const bool act = px < c.PX;
        const int iy = act ? px / W : 0, ix = act ? px % W : 0;
        const float u = blin11(ix, W), v = blin11(iy, H);
        float d_cv = 0.f, d_nz = 0.f;
        if (act) {
            float cv = 0.f, nz = 0.f;
            for (int s = 0; s < n; ++s) {
                const float pres = cc[s * 7 + 4];
                if (pres == 0.f) continue;
                const float xg = hg * ((u - cc[s * 7 + 2]) / cc[s * 7 + 0]) + hg;
                const float yg = hg * ((v - cc[s * 7 + 3]) / cc[s * 7 + 1]) + hg;
                const float* gs = sg + s * g;
                cv += pres * bilin_grad(xg, yg, G, G, [&](int gx, int gy) { return gs[gy * G + gx]; }).val;
                nz += pres * bilin_grad(xg, yg, G, G, [&](int, int) { return 1.f; }).val;
            }
            const float mask = bsigmoid(-10.f + nz * 20.f);
            const float mi = mean_img[px];
            cv += mi * mask;
            const float sd = mask * sf + (1.f - mask) * sb;
            const float z = (img[px] - cv) / sd;
            d_cv = up * z / sd;
            const float d_sd = up * (z * z - 1.f) / sd;
            const float d_mask = d_cv * mi + d_sd * (sf - sb);
            d_nz = d_mask * 20.f * mask * (1.f - mask);
            dmean[px] = d_cv * mask;
        }
        for (int s = 0; s < n; ++s) {
            const float pres = cc[s * 7 + 4];
            if (pres == 0.f) continue;                        // (uniform over the row's threads)
            float w4[4] = {0.f, 0.f, 0.f, 0.f};
            if (act) {
                const float sx = cc[s * 7 + 0], sy = cc[s * 7 + 1], tx = cc[s * 7 + 2], ty = cc[s * 7 + 3];
                const float xg = hg * ((u - tx) / sx) + hg, yg = hg * ((v - ty) / sy) + hg;
                const float* gs = sg + s * g;
                const BilinG bg = bilin_grad(xg, yg, G, G, [&](int gx, int gy) { return gs[gy * G + gx]; });
                const BilinG bo = bilin_grad(xg, yg, G, G, [&](int, int) { return 1.f; });
                for (int q = 0; q < 4; ++q)
                    if (bg.idx4[q] >= 0) SQB_AADD(dgl + s * g + bg.idx4[q], d_cv * pres * bg.w4[q]);
                const float d_xg = pres * (d_cv * bg.ddx + d_nz * bo.ddx), d_yg = pres * (d_cv * bg.ddy + d_nz * bo.ddy);
                w4[0] = d_xg * (-hg * (u - tx) / (sx * sx));
                w4[1] = d_yg * (-hg * (v - ty) / (sy * sy));
                w4[2] = d_xg * (-hg / sx);
                w4[3] = d_yg * (-hg / sy);
            }
            ex.warp_add4(wsum + s * 4, w4);
        }
    }
    ex.sync();
    // decoded glimpse = linear output * output_scale (modules.py:147): dY of the last decoder layer, d output_scale
    const float scale = c.prm[c.po.output_scale];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float* dy3 = dyp(c, L_DEC3, t, row, 0);
    for (int i = ex.tid; i < n * g; i += ex.nt) {
        dy3[i] = dgl[i] * scale;
        acc[0] += dgl[i] * (sg[i] / scale);
    }
    ex.sum4(acc);
    if (ex.tid == 0) SQB_AADD(c.small + SM_OUTSCALE, acc[0]);
    float* gz = c.gZc + (size_t)row * n * c.zw;
    for (int i = ex.tid; i < n * 4; i += ex.nt) {
        const int s = i / 4, q = i % 4;
        const float chain = q == 0 ? cc[s * 7 + 5] : q == 1 ? cc[s * 7 + 6] : q == 2 ? (1.f - cc[s * 7 + 2] * cc[s * 7 + 2])
                                                                                     : (1.f - cc[s * 7 + 3] * cc[s * 7 + 3]);
        gz[s * c.zw + nw + q] += wsum[i] * chain;
    }
}

template <class EX>
SQB_HD void bw_compact(const BwdCtx& c, EX& ex, int t, int row) {
    // adjoint of the stable present-first compaction (sqair_modules.py:514-582; index.py:132-165): a gather
    const int n = c.n, nw = c.nw, nh = c.nh, zw = c.zw;
    const float* perm = sgp(c, S_PERM, t, row, 0);
    const float* gz = c.gZc + (size_t)row * n * zw;
    float* gpr = c.gPropRec + (size_t)row * (n + 1) * zw;
    float* gdr = c.gDiscRec + (size_t)row * (n + 1) * zw;
    for (int i = ex.tid; i < n * zw; i += ex.nt) {
        const int j = i / zw, f = i - j * zw;
        if (f == nw + 4) continue;                            // presence: not differentiable
        const int k = (int)perm[j];
        float* dst = k < n ? gpr + (k + 1) * zw : gdr + (k - n + 1) * zw;
        dst[f] += gz[i];
    }
    const float* gt = c.gTc + (size_t)row * n * nh;
    const float* gq = c.gPc + (size_t)row * n * nh;
    float* t0 = c.xt[X_T0] + (size_t)(t * c.rows + row) * nh;
    float* p0 = c.xt[X_P0] + (size_t)(t * c.rows + row) * nh;
    for (int f = ex.tid; f < nh; f += ex.nt) {
        float a0 = 0.f, b0 = 0.f;
        for (int j = 0; j < n; ++j) {
            const int k = (int)perm[j];
            if (k < n) {
                c.gTnew[((size_t)row * n + k) * nh + f] += gt[j * nh + f];
                c.gPnew[((size_t)row * n + k) * nh + f] += gq[j * nh + f];
            } else {                                          // discovered objects start from the trainable initial states
                a0 += gt[j * nh + f]; b0 += gq[j * nh + f];
            }
        }
        t0[f] += a0; p0[f] += b0;
    }
}

template <class EX>
SQB_HD void bw_disc_post(const BwdCtx& c, EX& ex, int t, int row) {
    // discovery priors and the number-of-steps posterior (sqair_modules.py:149-226; modules.py:548-607; prior.py:61-102)
    const int n = c.n, nw = c.nw, zw = c.zw;
    const RecF& F = c.rec;
    const float gw = c.gw[row], gp = c.gp ? c.gp[row] : 0.f;
    float* gdr = c.gDiscRec + (size_t)row * (n + 1) * zw;
    float* sc = ex.scratch;                   // [n][4] dRN2
    if (ex.tid == 0) {
        int num = 0;
        for (int s = 0; s < n; ++s) num += sgp(c, S_DISCREC, t, row, s + 1)[F.pres] != 0.f ? 1 : 0;
        if (c.cfg.disc_prior_type == SQAIR_DISC_PRIOR_CAT) {
            // p(N) = Categorical(elu(bias + (t > 0) tbias + MLP(E[n_prop]))) (sqair_modules.py:208-221), coefficient +gw
            const float* spl = sgp(c, S_SPL, t, row, 0);
            const float* hsp = sgp(c, S_HSP, t, row, 0);
            float lg[MAX_SLOTS + 1], mx = -INFINITY, se = 0.f;
            for (int k = 0; k <= n; ++k) {
                const float v = c.prm[c.po.step_prior_bias + k] + (t == 0 ? 0.f : 1.f) * c.prm[c.po.step_prior_tbias + k] + spl[k];
                lg[k] = v > 0.f ? v : expm1f(v);
                mx = fmaxf(mx, lg[k]);
            }
            for (int k = 0; k <= n; ++k) se += expf(lg[k] - mx);
            float* d2 = dyp(c, L_SP2, t, row, 0);
            for (int k = 0; k <= n; ++k) {
                const float sm = expf(lg[k] - mx) / se;
                d2[k] = gw * ((k == num ? 1.f : 0.f) - sm) * (lg[k] > 0.f ? 1.f : lg[k] + 1.f);
            }
            const float* W2 = wmat(c, L_SP2);              // [10 + 1][n + 1]
            const float* W1 = wmat(c, L_SP1);              // [1 + 1][10]
            float* d1 = dyp(c, L_SP1, t, row, 0);
            float dexp = 0.f;
            for (int j = 0; j < 10; ++j) {
                float a = 0.f;
                for (int k = 0; k <= n; ++k) a += d2[k] * W2[j * (n + 1) + k];
                d1[j] = a * (hsp[j] > 0.f ? 1.f : hsp[j] + 1.f);
                dexp += d1[j] * W1[j];
            }
            c.gExp[row] += dexp;
        }
        // q(N): log(clip_preserve(joint[num])) (prior.py:95-102), coefficient -gw + gp; joint in f64 (prior.py:61-67)
        double pp[MAX_SLOTS], cum = 1.0, tot = 0.0, modn = 0.0;
        for (int k = 0; k < n; ++k) pp[k] = (double)sgp(c, S_DISCREC, t, row, k + 1)[F.prob];
        for (int k = 0; k < n; ++k) {
            const double mk = (1.0 - pp[k]) * cum;
            if (k == num) modn = mk;
            tot += mk;
            cum *= pp[k];
        }
        if (num == n) modn = cum;
        tot += cum;
        const float v = (float)(modn / tot);
        const double coef = (double)(-gw + gp) / (double)fminf(fmaxf(v, 1e-16f), 1.f) / tot;
        for (int j = 0; j < n && j <= num; ++j) {
            double d;                                        // d mod_num / d p_j, zero-safe (no division)
            if (j == num) {
                d = -1.0;
                for (int i = 0; i < num; ++i) d *= pp[i];
            } else {
                d = num < n ? (1.0 - pp[num]) : 1.0;
                for (int i = 0; i < (num < n ? num : n); ++i) if (i != j) d *= pp[i];
            }
            const double pj = pp[j];
            gdr[(j + 1) * zw + nw + 5] += (float)(coef * d * pj * (1.0 - pj));
        }
        // where prior of the discovered objects, coefficient +gw * presence
        for (int s = n - 1; s >= 0; --s) {
            const float* rec = sgp(c, S_DISCREC, t, row, s + 1);
            const float cs = gw * rec[F.pres];
            if (c.cfg.rec_where_prior) {
                const float* rns = sgp(c, S_RNS, t, row, s);
                const float* rno = sgp(c, S_RNO, t, row, s);
                float d3[8];
                for (int i = 0; i < 4; ++i) {
                    const NormG q = normal_lp_grad(rec[F.where + i], rns[i], rns[4 + i]);
                    gdr[(s + 1) * zw + nw + i] += cs * q.dx;
                    d3[i] = cs * q.dmu;
                    d3[4 + i] = cs * q.dsd * -expm1f(-(rns[4 + i] - 1e-2f));
                }
                float* dy3 = dyp(c, L_RN3, t, row, s);
                for (int i = 0; i < 8; ++i) dy3[i] = d3[i];
                const float* W3 = wmat(c, L_RN3);          // [4 + 1][8]
                const float* W2 = wmat(c, L_RN2);          // [4 + 128 + 1][4]
                float* dy2 = dyp(c, L_RN2, t, row, s);
                for (int k = 0; k < 4; ++k) {
                    float a = 0.f;
                    for (int j = 0; j < 8; ++j) a += d3[j] * W3[k * 8 + j];
                    dy2[k] = a * (1.f - rno[k] * rno[k]);
                    sc[s * 4 + k] = dy2[k];
                }
                for (int i = 0; i < 4; ++i) {                // previous sample: init_sample, then where_{s-1}
                    float a = 0.f;
                    for (int k = 0; k < 4; ++k) a += dy2[k] * W2[i * 4 + k];
                    if (s == 0) c.xt[X_RNSAMPLE][(size_t)(t * c.rows + row) * 4 + i] = a;
                    else gdr[s * zw + nw + i] += a;
                }
            } else {
                for (int i = 0; i < 4; ++i)
                    gdr[(s + 1) * zw + nw + i] += cs * normal_lp_grad(rec[F.where + i], c.cfg.where_mean[i], c.cfg.where_std[i]).dx;
            }
        }
    }
    ex.sync();
    if (c.cfg.rec_where_prior) {
        const float* W2 = wmat(c, L_RN2);
        for (int j = ex.tid; j < 128; j += ex.nt) {
            float a = 0.f;
            for (int s = 0; s < n; ++s)
                for (int k = 0; k < 4; ++k) a += sc[s * 4 + k] * W2[(4 + j) * 4 + k];
            c.gHrn[(size_t)row * 128 + j] = a;
        }
    }
    // what prior N(0, 1), coefficient +gw * presence
    for (int i = ex.tid; i < n * nw; i += ex.nt) {
        const int s = i / nw, j = i - s * nw;
        const float* rec = sgp(c, S_DISCREC, t, row, s + 1);
        gdr[(s + 1) * zw + j] += gw * rec[F.pres] * (-rec[F.what + j]);
    }
}

template <class EX>
SQB_HD void bw_disc_a(const BwdCtx& c, EX& ex, int t, int s, int row) {
    // presence logit of a discovery slot (core.py:206-208; modules.py:506-513) and the last steps-predictor layer
    const RecF& F = c.rec;
    const int e = s + 1;
    const float pkm1 = sgp(c, S_DISCREC, t, row, e - 1)[F.pres];
    const float dlg = c.gDiscRec[((size_t)row * (c.n + 1) + e) * c.zw + c.nw + 5] * pkm1;
    if (ex.tid == 0) dyp(c, L_DST2, t, row, s)[0] = dlg;
    const float* W2 = wmat(c, L_DST2);                     // [hs + 1][1]
    const float* hsv = sgp(c, S_DHS, t, row, s);
    float* d1 = dyp(c, L_DST1, t, row, s);
    for (int j = ex.tid; j < c.hs; j += ex.nt) d1[j] = dlg * W2[j] * (hsv[j] > 0.f ? 1.f : hsv[j] + 1.f);
}

template <class EX>
SQB_HD void bw_disc_b(const BwdCtx& c, EX& ex, int t, int s, int row) {
    // what ~ N(loc, scale) of a discovery slot (core.py:216-218) with its posterior term (sqair_modules.py:177-186)
    const RecF& F = c.rec;
    const int e = s + 1, nw = c.nw;
    const float* rec = sgp(c, S_DISCREC, t, row, e);
    const float cq = -c.gw[row] * rec[F.pres];
    const float* gr = c.gDiscRec + ((size_t)row * (c.n + 1) + e) * c.zw;
    const float* eps = c.eps_what + (((size_t)t * c.rows + row) * (2 * c.n) + c.n + s) * nw;
    float* d3 = dyp(c, L_ENC3, t, row, 2 * c.n + s);
    for (int j = ex.tid; j < nw; j += ex.nt) {
        const float ws = rec[F.what_scale + j];
        const NormG q = normal_lp_grad(rec[F.what + j], rec[F.what_loc + j], ws);
        const float gx = gr[j] + cq * q.dx;
        d3[j] = gx + cq * q.dmu;
        d3[nw + j] = (gx * eps[j] + cq * q.dsd) * -expm1f(-(ws - c.cfg.min_std));
    }
}

template <class EX>
SQB_HD void bw_disc_c(const BwdCtx& c, EX& ex, int t, int s, int row) {
    // where ~ N(loc, scale) of a discovery slot (core.py:220-227): glimpse-sampler gradient + posterior term, then the
    // last layer of the transform estimator by hand
    const RecF& F = c.rec;
    const int e = s + 1, nw = c.nw, nh = c.nh;
    const float* rec = sgp(c, S_DISCREC, t, row, e);
    const float* img = c.obs + ((size_t)t * c.cfg.B + row / c.cfg.K) * c.PX;
    float wl[4], stn[4];
    for (int i = 0; i < 4; ++i) wl[i] = rec[F.where + i];
    stn_where_grad(c, ex, img, wl, c.gGlm + (size_t)row * c.n * c.g, nullptr, nullptr, stn);
    float* tp = ex.scratch;                 // [8]
    if (ex.tid == 0) {
        const float cq = -c.gw[row] * rec[F.pres];
        const float* gr = c.gDiscRec + ((size_t)row * (c.n + 1) + e) * c.zw;
        const float* eps = c.eps_where + (((size_t)t * c.rows + row) * (2 * c.n) + c.n + s) * 4;
        float* d3 = dyp(c, L_DT3, t, row, s);
        float dso = 0.f;
        for (int i = 0; i < 4; ++i) {
            const float sd = rec[F.where_scale + i];
            const NormG q = normal_lp_grad(rec[F.where + i], rec[F.where_loc + i], sd);
            const float gx = gr[nw + i] + stn[i] + cq * q.dx;
            d3[i] = tp[i] = gx + cq * q.dmu;
            d3[4 + i] = tp[4 + i] = (gx * eps[i] + cq * q.dsd) * -expm1f(-(sd - 1e-2f));
            dso += d3[4 + i];
        }
        SQB_AADD(c.small + SM_DSO, dso);
    }
    ex.sync();
    const float* W3 = wmat(c, L_DT3);                      // [nh + 1][8]
    const float* a1 = sgp(c, S_DT2, t, row, s);
    float* d2 = dyp(c, L_DT2, t, row, s);
    for (int j = ex.tid; j < nh; j += ex.nt) {
        float a = 0.f;
        for (int i = 0; i < 8; ++i) a += tp[i] * W3[j * 8 + i];
        d2[j] = a * (a1[j] > 0.f ? 1.f : a1[j] + 1.f);
    }
}

template <class EX>
SQB_HD void bw_lat_pre(const BwdCtx& c, EX& ex, int t, int row) {
    // latent summary sum_s pres_s * MLP([what_s, where_s]) (sqair_modules.py:368-385,501) and the expected number of
    // propagated objects (sqair_modules.py:505-507)
    const int n = c.n, nh = c.nh;
    const float* gc = c.gDIn + (size_t)row * 2 * nh + nh;
    for (int i = ex.tid; i < n * nh; i += ex.nt) {
        const int s = i / nh, j = i - s * nh;
        c.tA1[(size_t)row * n * nh + i] = gc[j] * sgp(c, S_PROPREC, t, row, s + 1)[c.rec.pres];
    }
    for (int s = ex.tid; s < n; s += ex.nt) {
        const float sg = bsigmoid(sgp(c, S_PRI, t, row, s)[0]);
        c.gPri[((size_t)row * n + s) * c.npri] += c.gExp[row] * sg * (1.f - sg) / (float)n;
    }
}

template <class EX>
SQB_HD void bw_prop_a(const BwdCtx& c, EX& ex, int t, int s, int row) {
    // presence of a propagated slot: Bernoulli log-probs under q and p (sqair_modules.py:290-320; core.py:141-144)
    const RecF& F = c.rec;
    const int e = s + 1, nw = c.nw;
    const float* rec = sgp(c, S_PROPREC, t, row, e);
    const float ptm1 = sgp(c, S_Z, t, row, s)[nw + 4];
    const float gw = c.gw[row], gp = c.gp ? c.gp[row] : 0.f;
    float dlogit = c.gPropRec[((size_t)row * (c.n + 1) + e) * c.zw + nw + 5];
    dlogit += (-gw + gp) * ptm1 * (rec[F.pres] - rec[F.prob]);
    const float dlg = dlogit * ptm1;
    if (ex.tid == 0) {
        const float pl = sgp(c, S_PRI, t, row, s)[0];
        c.gPri[((size_t)row * c.n + s) * c.npri] += gw * ptm1 * (rec[F.pres] - bsigmoid(pl));
        dyp(c, L_PST2, t, row, s)[0] = dlg;
    }
    const float* W2 = wmat(c, L_PST2);
    const float* hsv = sgp(c, S_PHS, t, row, s);
    float* d1 = dyp(c, L_PST1, t, row, s);
    for (int j = ex.tid; j < c.hs; j += ex.nt) d1[j] = dlg * W2[j] * (hsv[j] > 0.f ? 1.f : hsv[j] + 1.f);
}

template <class EX>
SQB_HD void bw_prop_b(const BwdCtx& c, EX& ex, int t, int s, int row) {
    // what of a propagated slot (core.py:335-359) with its q / p terms (sqair_modules.py:290-317)
    const RecF& F = c.rec;
    const int e = s + 1, nw = c.nw, n = c.n;
    const float* rec = sgp(c, S_PROPREC, t, row, e);
    const float* Z = sgp(c, S_Z, t, row, s);
    const float* pri = sgp(c, S_PRI, t, row, s);
    const float* gt = sgp(c, S_GT, t, row, s);
    const float* tg = sgp(c, S_TG, t, row, s);
    const float* enc = sgp(c, S_ENC, t, row, n + s);
    const float mk = Z[nw + 4] * rec[F.pres];
    const float cq = -c.gw[row] * mk, cp = c.gw[row] * mk;
    const float* gr = c.gPropRec + ((size_t)row * (n + 1) + e) * c.zw;
    const float* eps = c.eps_what + (((size_t)t * c.rows + row) * (2 * n) + s) * nw;
    float* gpri = c.gPri + ((size_t)row * n + s) * c.npri;
    float* gzo = c.gZo + ((size_t)row * n + s) * c.zw;
    float* genc = c.gEnc + (size_t)row * 2 * nw;
    float* dh = dyp(c, L_PHEADS, t, row, s);
    for (int j = ex.tid; j < nw; j += ex.nt) {
        const float ws = rec[F.what_scale + j], x = rec[F.what + j];
        const NormG q = normal_lp_grad(x, rec[F.what_loc + j], ws);
        const NormG p = normal_lp_grad(x, pri[5 + j], pri[9 + nw + j]);
        const float gx = gr[j] + cq * q.dx + cp * p.dx;
        gpri[5 + j] += cp * p.dmu;
        gpri[9 + nw + j] += cp * p.dsd;
        const float dwl = gx + cq * q.dmu, dws = gx * eps[j] + cq * q.dsd;
        const float fg = gt[j], ig = gt[nw + j], tgt = gt[2 * nw + j];
        const float loc2 = enc[j], sc2 = enc[nw + j], loct = tg[j], sct = tg[nw + j];
        gzo[j] += dwl * fg;
        const float dfg = dwl * Z[j], dig = -(dwl * loc2 + dws * sc2), dtg = -(dwl * loct + dws * sct);
        genc[j] = dwl * (1.f - ig);
        genc[nw + j] = dws * (1.f - ig);
        dh[j] = dwl * (1.f - tgt);
        dh[nw + j] = dws * (1.f - tgt) * -expm1f(-(sct - c.cfg.min_std));
        dh[2 * nw + j] = dfg * act_deriv(ACT_SIGMOID, fg, 0.9999f, 0.f);
        dh[3 * nw + j] = dig * act_deriv(ACT_SIGMOID, ig, 0.9999f, 0.f);
        dh[4 * nw + j] = dtg * act_deriv(ACT_SIGMOID, tgt, 0.9999f, 0.f);
    }
}

// snt.GRU (Appendix B): h' = (1 - z) h + z c, c = tanh(x Wh + (r h) Uh + b).  Stage A: through the combination and the
// candidate's tanh; stage B: through r h and the gates' sigmoids.
template <class EX>
SQB_HD void gru_bw_a(EX& ex, int nh, const float* dnew, const float* z, const float* cc, const float* h, float* dh,
                     float* dy_c, float* dy_zr) {
    for (int j = ex.tid; j < nh; j += ex.nt) {
        const float d = dnew[j], zz = z[j], cv = cc[j];
        dh[j] += d * (1.f - zz);
        dy_c[j] = d * zz * (1.f - cv * cv);
        dy_zr[j] = d * (cv - h[j]) * zz * (1.f - zz);
    }
}
template <class EX>
SQB_HD void gru_bw_b(EX& ex, int nh, const float* drh, const float* r, const float* h, float* dh, float* dy_zr) {
    for (int j = ex.tid; j < nh; j += ex.nt) {
        const float d = drh[j], rr = r[j];
        dh[j] += d * rr;
        dy_zr[nh + j] = d * h[j] * rr * (1.f - rr);
    }
}

template <class EX>
SQB_HD void bw_prop_c(const BwdCtx& c, EX& ex, int t, int s, int row) {
    const size_t o = ((size_t)row * c.n + s) * c.nh;
    gru_bw_a(ex, c.nh, c.gTnew + o, sgp(c, S_TGZ, t, row, s), sgp(c, S_TGC, t, row, s), sgp(c, S_TST, t, row, s), c.gTo + o,
             dyp(c, L_TGRU_C, t, row, s), dyp(c, L_TGRU_ZR, t, row, s));
}
template <class EX>
SQB_HD void bw_prop_d(const BwdCtx& c, EX& ex, int t, int s, int row) {
    const size_t o = ((size_t)row * c.n + s) * c.nh;
    gru_bw_b(ex, c.nh, c.gRH + o, sgp(c, S_TGR, t, row, s), sgp(c, S_TST, t, row, s), c.gTo + o, dyp(c, L_TGRU_ZR, t, row, s));
}
template <class EX>
SQB_HD void bw_prop_e(const BwdCtx& c, EX& ex, int t, int s, int row) {
    // (loc2, scale2) of the glimpse encoder at the sampled where (core.py:336-337): softplus derivative of the scale
    const int nw = c.nw;
    const float* enc = sgp(c, S_ENC, t, row, c.n + s);
    const float* genc = c.gEnc + (size_t)row * 2 * nw;
    float* d3 = dyp(c, L_ENC3, t, row, c.n + s);
    for (int j = ex.tid; j < nw; j += ex.nt) {
        d3[j] = genc[j];
        d3[nw + j] = genc[nw + j] * -expm1f(-(enc[nw + j] - c.cfg.min_std));
    }
}

template <class EX>
SQB_HD void bw_prop_f(const BwdCtx& c, EX& ex, int t, int s, int row) {
    // where ~ MVN_TriL(where_tm1 + us * loc, L) of a propagated slot (core.py:321-333; modules.py:535-545): gradient of
    // the second glimpse, q (full covariance) and p (diagonal) terms, then the last transform-estimator layer by hand
    const RecF& F = c.rec;
    const int e = s + 1, nw = c.nw, nh = c.nh, n = c.n;
    const float* rec = sgp(c, S_PROPREC, t, row, e);
    const float* img = c.obs + ((size_t)t * c.cfg.B + row / c.cfg.K) * c.PX;
    const bool masked = c.cfg.masked_glimpse != 0;
    float wl[4], stn[4];
    for (int i = 0; i < 4; ++i) wl[i] = rec[F.where + i];
    stn_where_grad(c, ex, img, wl, c.gGlm + (size_t)row * n * c.g, masked ? sgp(c, S_MASK, t, row, s) : nullptr,
                   masked ? c.gMask + ((size_t)row * n + s) * c.g : nullptr, stn);
    float* tp = ex.scratch;
    if (ex.tid == 0) {
        const float* Z = sgp(c, S_Z, t, row, s);
        const float* pri = sgp(c, S_PRI, t, row, s);
        const float* gr = c.gPropRec + ((size_t)row * (n + 1) + e) * c.zw;
        const float* eps = c.eps_where + (((size_t)t * c.rows + row) * (2 * n) + s) * 4;
        float* gpri = c.gPri + ((size_t)row * n + s) * c.npri;
        float* gzo = c.gZo + ((size_t)row * n + s) * c.zw;
        // every input first (independent loads, one memory round trip), then the arithmetic, then the stores: interleaved
        // read-modify-writes would serialise a dozen L2 round trips in this single-thread section
        float x[4], loc[4], sd[4], L[4][4], y[4], w[4], gx[4], cs[4][4], pmu[4], psd[4], grv[4], ev[4], gpm[4], gps[4], gz[4], chol[10];
        const float mk = Z[nw + 4] * rec[F.pres], up = c.gw[row];
        for (int i = 0; i < 4; ++i) {
            x[i] = rec[F.where + i]; loc[i] = rec[F.where_loc + i]; sd[i] = rec[F.where_scale + i];
            pmu[i] = pri[1 + i]; psd[i] = pri[5 + nw + i]; grv[i] = gr[nw + i]; ev[i] = eps[i];
            gpm[i] = gpri[1 + i]; gps[i] = gpri[5 + nw + i]; gz[i] = gzo[nw + i];
        }
        for (int i = 0; i < 10; ++i) chol[i] = c.prm[c.po.cholesky + i];
        const float cq = -up * mk, cp = up * mk;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                cs[i][j] = j <= i ? chol[chol_idx(i, j)] : 0.f;
                L[i][j] = j <= i ? cs[i][j] * sd[i] + (i == j ? sd[i] : 0.f) : 0.f;
            }
        for (int i = 0; i < 4; ++i) {                       // L y = x - loc
            float a = x[i] - loc[i];
            for (int j = 0; j < i; ++j) a -= L[i][j] * y[j];
            y[i] = a / L[i][i];
        }
        for (int i = 3; i >= 0; --i) {                      // L^T w = y
            float a = y[i];
            for (int j = i + 1; j < 4; ++j) a -= L[j][i] * w[j];
            w[i] = a / L[i][i];
        }
        float d3[8], dso = 0.f, dchol[10];
        for (int i = 0; i < 10; ++i) dchol[i] = 0.f;
        for (int i = 0; i < 4; ++i) {
            const NormG p = normal_lp_grad(x[i], pmu[i], psd[i]);
            gx[i] = grv[i] + stn[i] + cp * p.dx + cq * (-w[i]);
            gpm[i] += cp * p.dmu;
            gps[i] += cp * p.dsd;
        }
        for (int i = 0; i < 4; ++i) {
            const float dloc = gx[i] + cq * w[i];
            gz[i] += dloc;
            d3[i] = c.cfg.where_update_scale * dloc;
            float dsd = 0.f;
            for (int j = 0; j <= i; ++j) {
                const float dL = gx[i] * ev[j] + cq * (w[i] * y[j] - (i == j ? 1.f / L[i][i] : 0.f));
                dsd += dL * (cs[i][j] + (i == j ? 1.f : 0.f));
                dchol[chol_idx(i, j)] += dL * sd[i];
            }
            d3[4 + i] = dsd * -expm1f(-(sd[i] - 1e-2f));
            dso += d3[4 + i];
        }
        for (int i = 0; i < 4; ++i) { gpri[1 + i] = gpm[i]; gpri[5 + nw + i] = gps[i]; gzo[nw + i] = gz[i]; }
        for (int i = 0; i < 10; ++i) SQB_AADD(c.small + SM_CHOL + i, dchol[i]);
        SQB_AADD(c.small + SM_PSO, dso);
        float* dy3 = dyp(c, L_PT3, t, row, s);
        for (int i = 0; i < 8; ++i) dy3[i] = tp[i] = d3[i];
    }
    ex.sync();
    const float* W3 = wmat(c, L_PT3);
    const float* a1 = sgp(c, S_PT2, t, row, s);
    float* d2 = dyp(c, L_PT2, t, row, s);
    for (int j = ex.tid; j < nh; j += ex.nt) {
        float a = 0.f;
        for (int i = 0; i < 8; ++i) a += tp[i] * W3[j * 8 + i];
        d2[j] = a * (a1[j] > 0.f ? 1.f : a1[j] + 1.f);
    }
}

template <class EX>
SQB_HD void bw_stn1(const BwdCtx& c, EX& ex, int t, int row) {
    // first glimpse of every propagated slot at where_tm1 + where_bias (core.py:291-293)
    const int n = c.n, nw = c.nw;
    const float* img = c.obs + ((size_t)t * c.cfg.B + row / c.cfg.K) * c.PX;
    const bool masked = c.cfg.masked_glimpse != 0;
    for (int s = 0; s < n; ++s) {
        const float* Z = sgp(c, S_Z, t, row, s);
        const float* wb = sgp(c, S_WB, t, row, s);
        float wl[4], stn[4];
        for (int i = 0; i < 4; ++i) wl[i] = Z[nw + i] + wb[i];
        stn_where_grad(c, ex, img, wl, c.gGlm + ((size_t)row * n + s) * c.g, masked ? sgp(c, S_MASK, t, row, s) : nullptr,
                       masked ? c.gMask + ((size_t)row * n + s) * c.g : nullptr, stn);
        for (int i = ex.tid; i < 4; i += ex.nt) {
            c.gZo[((size_t)row * n + s) * c.zw + nw + i] += stn[i];
            dyp(c, L_WB2, t, row, s)[i] = 0.1f * stn[i];                  // where_bias = 0.1 * MLP(...) (core.py:291)
        }
    }
}

template <class EX>
SQB_HD void bw_prior_pre(const BwdCtx& c, EX& ex, int t, int row) {
    // post-processing of the propagation prior statistics (propagate.py:84-98,123-158)
    const int n = c.n, nw = c.nw, np = c.npri;
    for (int i = ex.tid; i < n * np; i += ex.nt) {
        const int s = i / np, f = i - s * np;
        const float g = c.gPri[(size_t)row * n * np + i];
        const float* Z = sgp(c, S_Z, t, row, s);
        float* gzo = c.gZo + ((size_t)row * n + s) * c.zw;
        float d;
        if (f == 0) {
            d = g * Z[nw + 4];
            if (c.cfg.prior_type != SQAIR_PRIOR_RNN) { d *= 0.1f; gzo[nw + 5] += g; }
        } else if (f <= 4 + nw) {
            const int zf = (f <= 4) ? (nw + f - 1) : (f - 5);
            if (c.cfg.prior_type == SQAIR_PRIOR_RW) { d = 0.f; gzo[zf] += g; }
            else if (c.cfg.prior_type == SQAIR_PRIOR_GUIDED) { d = 0.1f * g; gzo[zf] += g; }
            else d = g;
        } else {
            d = g * -expm1f(-(sgp(c, S_PRI, t, row, s)[f] - 1e-2f));
        }
        dyp(c, L_PLIN, t, row, s)[f] = d;
    }
}

template <class EX>
SQB_HD void bw_pgru_a(const BwdCtx& c, EX& ex, int t, int row) {
    for (int s = 0; s < c.n; ++s) {
        const size_t o = ((size_t)row * c.n + s) * c.nh;
        gru_bw_a(ex, c.nh, c.gPnew + o, sgp(c, S_PGZ, t, row, s), sgp(c, S_PGC, t, row, s), sgp(c, S_PST, t, row, s), c.gPo + o,
                 dyp(c, L_PGRU_C, t, row, s), dyp(c, L_PGRU_ZR, t, row, s));
    }
}
template <class EX>
SQB_HD void bw_pgru_b(const BwdCtx& c, EX& ex, int t, int row) {
    for (int s = 0; s < c.n; ++s) {
        const size_t o = ((size_t)row * c.n + s) * c.nh;
        gru_bw_b(ex, c.nh, c.gRH + o, sgp(c, S_PGR, t, row, s), sgp(c, S_PST, t, row, s), c.gPo + o, dyp(c, L_PGRU_ZR, t, row, s));
    }
}

template <class EX>
SQB_HD void bw_final_states(const BwdCtx& c, EX& ex, int row) {
    // t = 0: every slot starts from the trainable initial GRU states (seq.py:97-104)
    const int n = c.n, nh = c.nh;
    float* t0 = c.xt[X_T0] + (size_t)row * nh;
    float* p0 = c.xt[X_P0] + (size_t)row * nh;
    for (int f = ex.tid; f < nh; f += ex.nt) {
        float a = 0.f, b = 0.f;
        for (int s = 0; s < n; ++s) { a += c.gTc[((size_t)row * n + s) * nh + f]; b += c.gPc[((size_t)row * n + s) * nh + f]; }
        t0[f] += a; p0[f] += b;
    }
}

template <int STAGE, class EX>
SQB_HD void bw_stage(const BwdCtx& c, EX& ex, int t, int s, int row) {
    if (STAGE == BS_CANVAS) bw_canvas(c, ex, t, row);
    else if (STAGE == BS_COMPACT) bw_compact(c, ex, t, row);
    else if (STAGE == BS_DISC_POST) bw_disc_post(c, ex, t, row);
    else if (STAGE == BS_DISC_A) bw_disc_a(c, ex, t, s, row);
    else if (STAGE == BS_DISC_B) bw_disc_b(c, ex, t, s, row);
    else if (STAGE == BS_DISC_C) bw_disc_c(c, ex, t, s, row);
    else if (STAGE == BS_LAT_PRE) bw_lat_pre(c, ex, t, row);
    else if (STAGE == BS_PROP_A) bw_prop_a(c, ex, t, s, row);
    else if (STAGE == BS_PROP_B) bw_prop_b(c, ex, t, s, row);
    else if (STAGE == BS_PROP_C) bw_prop_c(c, ex, t, s, row);
    else if (STAGE == BS_PROP_D) bw_prop_d(c, ex, t, s, row);
    else if (STAGE == BS_PROP_E) bw_prop_e(c, ex, t, s, row);
    else if (STAGE == BS_PROP_F) bw_prop_f(c, ex, t, s, row);
    else if (STAGE == BS_STN1) bw_stn1(c, ex, t, row);
    else if (STAGE == BS_PRIOR_PRE) bw_prior_pre(c, ex, t, row);
    else if (STAGE == BS_PGRU_A) bw_pgru_a(c, ex, t, row);
    else if (STAGE == BS_PGRU_B) bw_pgru_b(c, ex, t, row);
    else if (STAGE == BS_FINAL_STATES) bw_final_states(c, ex, row);
}

SQ_HD int bw_stage_scratch_floats(const sqair_cfg& c) { return 2 * c.n * c.G * c.G + 11 * c.n + 64; }

// ---------------------------------------------------------------------------------------------
// Workspace layout (host)
// ---------------------------------------------------------------------------------------------
struct BwdLayout {
    int64_t dy_off[L_COUNT];
    int dy_e[L_COUNT], dy_w[L_COUNT];
    int64_t xt_off[X_COUNT];
    int xt_w[X_COUNT];
    int64_t small_off;
    int64_t dyz_begin, dyz_end;       // region zeroed once per backward call (extras + small)
    int64_t frame_begin, frame_end;   // per-frame gradient buffers, zeroed at the start of every frame...
    int64_t carry_off[6];             // ... except the two carried state-gradient sets (ping-pong)
    int64_t gPropRec, gDiscRec, gTnew, gPnew, gPH, gDH, gDIn, gExp, gHrn, gPri, gMask, gGlm, gLoc1, gHwbmk, gEnc, gRH, tA0, tA1;
    int64_t dwv_off;                  // [bw_total] virtual-matrix gradients
    int64_t img_dy_off;               // [T, B, nh] image-encoder dY summed over the particles of a sequence
    int64_t bcast_dy_off;             // [T, rows, nh] dY summed over the slots (weight gradient against an operand shared by the slots)
    int nframe;                       // the arrays of the per-frame region, each [rows, frame_stride[i]] (cleared per row range
    int64_t frame_off[16];            //  by the persistent reverse-program kernel: every cluster clears the rows it owns)
    int frame_stride[16];
    int64_t total;
};

inline BwdLayout build_bwd_layout(const sqair_cfg& c, const Plan& plan) {
    BwdLayout L;
    memset(&L, 0, sizeof(L));
    const int n = c.n, nw = c.n_what, nh = c.n_hidden, g = c.G * c.G, rows = c.B * c.K, T = c.T, zw = nw + 6;
    const int npri = 2 * (4 + nw) + 1;
    int64_t cur = 0;
    auto take = [&](int64_t floats) { int64_t o = cur; cur += (floats + 3) / 4 * 4; return o; };
    for (int l = 0; l < L_COUNT; ++l) {
        if (plan.L[l].nhead == 0) { L.dy_off[l] = -1; continue; }
        int e = n;
        if (l == L_ENC1 || l == L_ENC2 || l == L_ENC3) e = 3 * n;
        if (l == L_IMG1 || l == L_IMG2 || l == L_RN1 || l == L_SP1 || l == L_SP2) e = 1;
        L.dy_e[l] = e;
        L.dy_w[l] = plan.LB[l].NU >= 32 ? (plan.LB[l].NU + 3) / 4 * 4 : plan.LB[l].NU;      // 16-byte row stride: TMA-addressable
        L.dy_off[l] = take((int64_t)T * rows * e * L.dy_w[l]);
    }
    L.dyz_begin = cur;
    const int xw[X_COUNT] = {nh, nh, nh, nh, c.H * c.W, 4, 4};
    for (int i = 0; i < X_COUNT; ++i) { L.xt_w[i] = xw[i]; L.xt_off[i] = take((int64_t)T * rows * xw[i]); }
    L.small_off = take(SM_COUNT);
    L.dyz_end = cur;
    for (int i = 0; i < 6; ++i) L.carry_off[i] = take((int64_t)rows * n * (i % 3 == 0 ? zw : nh));
    L.frame_begin = cur;
    auto frame_take = [&](int stride) { int64_t o = take((int64_t)rows * stride); L.frame_off[L.nframe] = o; L.frame_stride[L.nframe++] = stride; return o; };
    L.gPropRec = frame_take((n + 1) * zw); L.gDiscRec = frame_take((n + 1) * zw);
    L.gTnew = frame_take(n * nh); L.gPnew = frame_take(n * nh);
    L.gPH = frame_take((n + 1) * nh); L.gDH = frame_take((n + 1) * nh);
    L.gDIn = frame_take(2 * nh); L.gExp = frame_take(1); L.gHrn = frame_take(128);
    L.gPri = frame_take(n * npri);
    L.gMask = frame_take(n * g);
    L.frame_end = cur;
    L.gGlm = take((int64_t)rows * n * g);
    L.gLoc1 = take((int64_t)rows * n * nw);
    L.gHwbmk = take((int64_t)rows * n * 256);
    L.gEnc = take((int64_t)rows * 2 * nw);
    L.gRH = take((int64_t)rows * n * nh);
    L.tA0 = take((int64_t)rows * n * nh); L.tA1 = take((int64_t)rows * n * nh);
    L.dwv_off = take(plan.bw_total);
    L.img_dy_off = take((int64_t)T * c.B * nh);
    L.bcast_dy_off = take((int64_t)T * rows * nh);
    L.total = cur;
    return L;
}

// ---------------------------------------------------------------------------------------------
// Driver.  Backend BE supplies: stage<STAGE>(ctx, t, s) (all rows), dgrad(args), wgrad(args), colsum(args),
// zero(ptr, floats), img_reduce(dy [T*rows, nh] -> [T*B, nh]) and finish() hooks.
// ---------------------------------------------------------------------------------------------
struct BwdInputs {
    const float* params;      // canonical flat parameters
    const float* bw;          // backward parameter buffer (sqair_pack_backward)
    const float* obs;
    const float* eps_where;
    const float* eps_what;
    const float* stash;
    const float* d_log_w;     // [rows]
    const float* d_disc_lp;   // [rows]
    float* ws;                // workspace (build_bwd_layout)
    float* d_params;          // canonical flat gradient (overwritten)
    int vimco;
};

inline void make_bwd_ctx(const sqair_cfg& cfg, const Plan& plan, const POff& poc, const BwdLayout& BL, const BwdInputs& in,
                         BwdCtx& c) {
    memset((void*)&c, 0, sizeof(c));
    c.cfg = cfg;
    c.T = cfg.T; c.rows = cfg.B * cfg.K; c.n = cfg.n; c.nw = cfg.n_what; c.nh = cfg.n_hidden; c.g = cfg.G * cfg.G;
    c.hs = cfg.n_hidden / 2; c.PX = cfg.H * cfg.W; c.zw = cfg.n_what + 6; c.recw = plan.rec.size; c.npri = 2 * (4 + cfg.n_what) + 1;
    c.rec = plan.rec; c.po = poc; c.vimco = in.vimco;
    c.prm = in.params; c.bw = in.bw;
    for (int l = 0; l < L_COUNT; ++l) {
        c.bw_off[l] = plan.LB[l].bw_off; c.bw_nu[l] = plan.LB[l].NU; c.bw_ku[l] = plan.LB[l].KU;
        c.dy[l] = BL.dy_off[l] >= 0 ? in.ws + BL.dy_off[l] : nullptr;
        c.dy_e[l] = BL.dy_e[l]; c.dy_w[l] = BL.dy_w[l];
    }
    c.obs = in.obs; c.eps_where = in.eps_where; c.eps_what = in.eps_what; c.stash = in.stash;
    for (int i = 0; i < S_COUNT; ++i) c.st[i] = plan.st[i];
    c.gw = in.d_log_w; c.gp = in.d_disc_lp;
    for (int i = 0; i < X_COUNT; ++i) c.xt[i] = in.ws + BL.xt_off[i];
    c.small = in.ws + BL.small_off;
    c.gPropRec = in.ws + BL.gPropRec; c.gDiscRec = in.ws + BL.gDiscRec; c.gTnew = in.ws + BL.gTnew; c.gPnew = in.ws + BL.gPnew;
    c.gPH = in.ws + BL.gPH; c.gDH = in.ws + BL.gDH; c.gDIn = in.ws + BL.gDIn; c.gExp = in.ws + BL.gExp; c.gHrn = in.ws + BL.gHrn;
    c.gPri = in.ws + BL.gPri; c.gMask = in.ws + BL.gMask; c.gGlm = in.ws + BL.gGlm; c.gLoc1 = in.ws + BL.gLoc1;
    c.gHwbmk = in.ws + BL.gHwbmk; c.gEnc = in.ws + BL.gEnc; c.gRH = in.ws + BL.gRH; c.tA0 = in.ws + BL.tA0; c.tA1 = in.ws + BL.tA1;
}

template <class BE>
struct BwdDriver {
    BE& be;
    const sqair_cfg& cfg;
    const Plan& plan;
    const BwdLayout& BL;
    BwdCtx c;
    float* ws;
    int n, nw, nh, g, rows, zw, recw, npri, T;

    BwdDriver(BE& b, const sqair_cfg& cf, const Plan& pl, const POff& poc, const BwdLayout& bl, const BwdInputs& in)
        : be(b), cfg(cf), plan(pl), BL(bl) {
        make_bwd_ctx(cf, pl, poc, bl, in, c);
        ws = in.ws;
        n = c.n; nw = c.nw; nh = c.nh; g = c.g; rows = c.rows; zw = c.zw; recw = c.recw; npri = c.npri; T = c.T;
    }

    // ---- operand helpers: one slot (M = rows) or all slots (M = rows * n) of a [rows, entries, width] buffer ----
    Addr slot_of(float* base, int entries, int width, int e, int col = 0) const {
        return mk_addr(base + (size_t)e * width + col, entries * width, 0);
    }
    Addr all_of(float* base, int entries, int width, int e0, int col = 0) const {
        return mk_addr(base + (size_t)e0 * width + col, entries * width, width);
    }
    Addr sig_slot(int sig, int t, int e, int col = 0) const {
        const Sig s = c.st[sig];
        return mk_addr(c.stash + (size_t)s.off + ((size_t)t * rows * s.entries + e) * s.width + col, s.entries * s.width, 0);
    }
    Addr sig_all(int sig, int t, int e0, int col = 0) const {
        const Sig s = c.st[sig];
        return mk_addr(c.stash + (size_t)s.off + ((size_t)t * rows * s.entries + e0) * s.width + col, s.entries * s.width, s.width);
    }
    Addr dy_slot(int l, int t, int e) const {
        return mk_addr(c.dy[l] + ((size_t)t * rows * c.dy_e[l] + e) * c.dy_w[l], c.dy_e[l] * c.dy_w[l], 0);
    }
    Addr dy_all(int l, int t, int e0) const {
        return mk_addr(c.dy[l] + ((size_t)t * rows * c.dy_e[l] + e0) * c.dy_w[l], c.dy_e[l] * c.dy_w[l], c.dy_w[l]);
    }
    static Addr none() { return mk_addr(nullptr, 0, 0); }

    struct SegOut {
        int mode;
        Addr d;
        int skip_tail;       // rows at the end of the segment that receive no gradient (the presence input)
    };
    static SegOut seg(int mode, Addr d, int skip_tail = 0) { SegOut s; s.mode = mode; s.d = d; s.skip_tail = skip_tail; return s; }
    static SegOut skip() { return seg(SEGM_SKIP, none()); }

    // dX = (a * act'(y)) . W_l^T for layer l; `outs[i]` says where input segment i's gradient goes
    void dgrad(int l, int M, int ny, Addr a, Addr y, int act, Addr dy, const SegOut* outs, int nouts, float act_scale = 1.f,
               float act_add = 0.f) {
        const Layer& L = plan.L[l];
        const LayerB& LB = plan.LB[l];
        DgradArgs A;
        memset((void*)&A, 0, sizeof(A));
        A.layer = l;
        A.M = M; A.N = LB.NU; A.K = LB.KU; A.ny = ny;
        A.a = a; A.y = y; A.dy = dy; A.act = act; A.act_scale = act_scale; A.act_add = act_add;
        A.w = c.bw + LB.bw_off;
        const int nseg = L.nseg - LB.has_bias;
        A.nseg = 0;
        for (int i = 0; i < nseg && i < nouts; ++i) {
            if (outs[i].mode == SEGM_SKIP) continue;
            DgradArgs::Seg& s = A.seg[A.nseg++];
            s.k0 = LB.u0[i]; s.k1 = LB.u0[i] + L.seg[i].K - outs[i].skip_tail; s.mode = outs[i].mode; s.d = outs[i].d;
        }
        be.dgrad(A);
    }

    void zero(int64_t off0, int64_t off1) { be.zero(ws + off0, off1 - off0); }

    // ------------------------------------------------------------------------------------------
    void frame(int t, int parity) {
        // carried gradients: c.g?c = w.r.t. the state leaving frame t (from frame t + 1), c.g?o = entering (to frame t - 1)
        c.gZc = ws + BL.carry_off[parity * 3 + 0]; c.gTc = ws + BL.carry_off[parity * 3 + 1]; c.gPc = ws + BL.carry_off[parity * 3 + 2];
        c.gZo = ws + BL.carry_off[(1 - parity) * 3 + 0]; c.gTo = ws + BL.carry_off[(1 - parity) * 3 + 1];
        c.gPo = ws + BL.carry_off[(1 - parity) * 3 + 2];
        zero(BL.frame_begin, BL.frame_end);
        for (int i = 0; i < 3; ++i) be.zero(ws + BL.carry_off[(1 - parity) * 3 + i], (int64_t)rows * n * (i == 0 ? zw : nh));
        const bool masked = cfg.masked_glimpse != 0;
        const int M1 = rows, MN = rows * n;

        // ---- canvas + likelihood, decoder (modules.py:131-147,435-467) ----
        be.template stage<BS_CANVAS>(c, t, 0);
        { SegOut o[1] = {seg(SEGM_STORE, all_of(c.tA1, n, nh, 0))};
          dgrad(L_DEC3, MN, n, dy_all(L_DEC3, t, 0), none(), ACT_NONE, none(), o, 1); }
        { SegOut o[1] = {seg(SEGM_STORE, all_of(c.tA0, n, nh, 0))};
          dgrad(L_DEC2, MN, n, all_of(c.tA1, n, nh, 0), sig_all(S_D2, t, 0), ACT_ELU, dy_all(L_DEC2, t, 0), o, 1); }
        { SegOut o[1] = {seg(SEGM_ACC, all_of(c.gZc, n, zw, 0))};
          dgrad(L_DEC1, MN, n, all_of(c.tA0, n, nh, 0), sig_all(S_D1, t, 0), ACT_ELU, dy_all(L_DEC1, t, 0), o, 1); }
        // ---- slot compaction ----
        be.template stage<BS_COMPACT>(c, t, 0);
        // ---- discovery priors, count posterior ----
        be.template stage<BS_DISC_POST>(c, t, 0);
        if (cfg.rec_where_prior) {
            SegOut o[3] = {seg(SEGM_STORE, mk_addr(c.xt[X_RNINIT] + (size_t)t * rows * 4, 4, 0)),
                           seg(SEGM_ACC, mk_addr(c.gDIn + nh, 2 * nh, 0)), seg(SEGM_ACC, mk_addr(c.gExp, 1, 0))};
            dgrad(L_RN1, M1, 1, mk_addr(c.gHrn, 128, 0), sig_slot(S_HRN, t, 0), ACT_ELU, dy_slot(L_RN1, t, 0), o, 3);
        }
        // ---- discovery slots, last to first (core.py:192-227) ----
        for (int s = n - 1; s >= 0; --s) {
            const int e = s + 1;
            be.template stage<BS_DISC_A>(c, t, s);
            { SegOut o[2] = {seg(SEGM_ACC, slot_of(c.gDH, n + 1, nh, e)), seg(SEGM_ACC, slot_of(c.gDiscRec, n + 1, zw, e))};
              dgrad(L_DST1, M1, 1, dy_slot(L_DST1, t, s), none(), ACT_NONE, none(), o, 2); }
            be.template stage<BS_DISC_B>(c, t, s);
            encoder_bwd(t, 2 * n + s, M1, 1, dy_slot(L_ENC3, t, 2 * n + s), L_ENC3, slot_of(c.gGlm, n, g, 0));
            be.template stage<BS_DISC_C>(c, t, s);
            { SegOut o[1] = {seg(SEGM_STORE, slot_of(c.tA0, n, nh, 0))};
              dgrad(L_DT2, M1, 1, dy_slot(L_DT2, t, s), none(), ACT_NONE, none(), o, 1); }
            { SegOut o[1] = {seg(SEGM_ACC, slot_of(c.gDH, n + 1, nh, e))};
              dgrad(L_DT1, M1, 1, slot_of(c.tA0, n, nh, 0), sig_slot(S_DT1, t, s), ACT_ELU, dy_slot(L_DT1, t, s), o, 1); }
            { SegOut o[3] = {seg(SEGM_ACC, mk_addr(c.gDIn, 2 * nh, 0)), seg(SEGM_ACC, slot_of(c.gDiscRec, n + 1, zw, e - 1), 1),
                             s > 0 ? seg(SEGM_ACC, slot_of(c.gDH, n + 1, nh, e - 1))
                                   : seg(SEGM_STORE, mk_addr(c.xt[X_DH0] + (size_t)t * rows * nh, nh, 0))};
              dgrad(L_DRNN, M1, 1, slot_of(c.gDH, n + 1, nh, e), sig_slot(S_DH, t, e), ACT_TANH, dy_slot(L_DRNN, t, s), o, 3); }
        }
        // ---- image encoder (core.py:165) ----
        { SegOut o[1] = {seg(SEGM_STORE, slot_of(c.tA0, n, nh, 0))};
          dgrad(L_IMG2, M1, 1, mk_addr(c.gDIn, 2 * nh, 0), sig_slot(S_DIN, t, 0), ACT_ELU, dy_slot(L_IMG2, t, 0), o, 1); }
        dgrad(L_IMG1, M1, 1, slot_of(c.tA0, n, nh, 0), sig_slot(S_IMG1, t, 0), ACT_ELU, dy_slot(L_IMG1, t, 0), nullptr, 0);
        // ---- latent summary (sqair_modules.py:368-385) ----
        be.template stage<BS_LAT_PRE>(c, t, 0);
        { SegOut o[1] = {seg(SEGM_STORE, all_of(c.tA0, n, nh, 0))};
          dgrad(L_LAT2, MN, n, all_of(c.tA1, n, nh, 0), sig_all(S_L2, t, 0), ACT_ELU, dy_all(L_LAT2, t, 0), o, 1); }
        { SegOut o[1] = {seg(SEGM_ACC, all_of(c.gPropRec, n + 1, zw, 1))};
          dgrad(L_LAT1, MN, n, all_of(c.tA0, n, nh, 0), sig_all(S_L1, t, 0), ACT_ELU, dy_all(L_LAT1, t, 0), o, 1); }
        // ---- propagation slots, last to first (core.py:280-359) ----
        for (int s = n - 1; s >= 0; --s) {
            const int e = s + 1;
            const Addr gph = slot_of(c.gPH, n + 1, nh, e), gto = slot_of(c.gTo, n, nh, s), grec = slot_of(c.gPropRec, n + 1, zw, e);
            be.template stage<BS_PROP_A>(c, t, s);
            { SegOut o[3] = {seg(SEGM_ACC, gph), seg(SEGM_ACC, gto), seg(SEGM_ACC, grec)};
              dgrad(L_PST1, M1, 1, dy_slot(L_PST1, t, s), none(), ACT_NONE, none(), o, 3); }
            be.template stage<BS_PROP_B>(c, t, s);
            { SegOut o[1] = {seg(SEGM_ACC, slot_of(c.gTnew, n, nh, s))};
              dgrad(L_PHEADS, M1, 1, dy_slot(L_PHEADS, t, s), none(), ACT_NONE, none(), o, 1); }
            be.template stage<BS_PROP_C>(c, t, s);
            const Addr gwhere = slot_of(c.gPropRec, n + 1, zw, e, nw), genc = mk_addr(c.gEnc, 2 * nw, 0);
            { SegOut o[4] = {seg(SEGM_ACC, gph), seg(SEGM_ACC, gwhere), seg(SEGM_ACC, genc), seg(SEGM_STORE, slot_of(c.gRH, n, nh, s))};
              dgrad(L_TGRU_C, M1, 1, dy_slot(L_TGRU_C, t, s), none(), ACT_NONE, none(), o, 4); }
            be.template stage<BS_PROP_D>(c, t, s);
            { SegOut o[4] = {seg(SEGM_ACC, gph), seg(SEGM_ACC, gwhere), seg(SEGM_ACC, genc), seg(SEGM_ACC, gto)};
              dgrad(L_TGRU_ZR, M1, 1, dy_slot(L_TGRU_ZR, t, s), none(), ACT_NONE, none(), o, 4); }
            be.template stage<BS_PROP_E>(c, t, s);
            encoder_bwd(t, n + s, M1, 1, dy_slot(L_ENC3, t, n + s), L_ENC3, slot_of(c.gGlm, n, g, 0));
            be.template stage<BS_PROP_F>(c, t, s);
            { SegOut o[1] = {seg(SEGM_STORE, slot_of(c.tA0, n, nh, 0))};
              dgrad(L_PT2, M1, 1, dy_slot(L_PT2, t, s), none(), ACT_NONE, none(), o, 1); }
            { SegOut o[3] = {seg(SEGM_ACC, gph), seg(SEGM_ACC, slot_of(c.gZo, n, zw, s, nw)), seg(SEGM_ACC, gto)};
              dgrad(L_PT1, M1, 1, slot_of(c.tA0, n, nh, 0), sig_slot(S_PT1, t, s), ACT_ELU, dy_slot(L_PT1, t, s), o, 3); }
            { SegOut o[5] = {seg(SEGM_STORE, slot_of(c.gLoc1, n, nw, s)), seg(SEGM_ACC, slot_of(c.gPropRec, n + 1, zw, e - 1), 1),
                             seg(SEGM_ACC, slot_of(c.gZo, n, zw, s), 1), seg(SEGM_ACC, gto),
                             s > 0 ? seg(SEGM_ACC, slot_of(c.gPH, n + 1, nh, e - 1))
                                   : seg(SEGM_STORE, mk_addr(c.xt[X_PH0] + (size_t)t * rows * nh, nh, 0))};
              dgrad(L_PRNN, M1, 1, gph, sig_slot(S_PH, t, e), ACT_TANH, dy_slot(L_PRNN, t, s), o, 5); }
        }
        // ---- parts of propagation off the slot recursion, all slots at once ----
        encoder_bwd(t, 0, MN, n, all_of(c.gLoc1, n, nw, 0), L_ENC3_LOC, all_of(c.gGlm, n, g, 0));
        be.template stage<BS_STN1>(c, t, 0);
        if (masked) {
            SegOut o[1] = {seg(SEGM_STORE, all_of(c.gHwbmk, n, 256, 0, 128))};
            dgrad(L_MK2, MN, n, all_of(c.gMask, n, g, 0), sig_all(S_MASK, t, 0), ACT_SIGMOID, dy_all(L_MK2, t, 0), o, 1);
        }
        { SegOut o[1] = {seg(SEGM_STORE, all_of(c.gHwbmk, n, 256, 0, 0))};
          dgrad(L_WB2, MN, n, dy_all(L_WB2, t, 0), none(), ACT_NONE, none(), o, 1); }
        { SegOut o[1] = {seg(SEGM_ACC, all_of(c.gTo, n, nh, 0))};
          dgrad(L_WBMK1, MN, n, all_of(c.gHwbmk, n, 256, 0), sig_all(S_HWBMK, t, 0), ACT_ELU, dy_all(L_WBMK1, t, 0), o, 1); }
        // ---- propagation prior (propagate.py:68-98) ----
        be.template stage<BS_PRIOR_PRE>(c, t, 0);
        { SegOut o[1] = {seg(SEGM_ACC, all_of(c.gPnew, n, nh, 0))};
          dgrad(L_PLIN, MN, n, dy_all(L_PLIN, t, 0), none(), ACT_NONE, none(), o, 1); }
        be.template stage<BS_PGRU_A>(c, t, 0);
        { SegOut o[2] = {seg(SEGM_ACC, all_of(c.gZo, n, zw, 0)), seg(SEGM_STORE, all_of(c.gRH, n, nh, 0))};
          dgrad(L_PGRU_C, MN, n, dy_all(L_PGRU_C, t, 0), none(), ACT_NONE, none(), o, 2); }
        be.template stage<BS_PGRU_B>(c, t, 0);
        { SegOut o[2] = {seg(SEGM_ACC, all_of(c.gZo, n, zw, 0)), seg(SEGM_ACC, all_of(c.gPo, n, nh, 0))};
          dgrad(L_PGRU_ZR, MN, n, dy_all(L_PGRU_ZR, t, 0), none(), ACT_NONE, none(), o, 2); }
    }

    // glimpse encoder backward for stash entries [entry, entry + ny): last layer `l3` (L_ENC3 / L_ENC3_LOC) given its
    // pre-activation gradient `a3` -> gradient of the (masked) glimpse in `gglm`
    void encoder_bwd(int t, int entry, int M, int ny, Addr a3, int l3, Addr gglm) {
        const bool all = ny > 1;
        const Addr ta1 = all ? all_of(c.tA1, n, nh, 0) : slot_of(c.tA1, n, nh, 0);
        const Addr ta0 = all ? all_of(c.tA0, n, nh, 0) : slot_of(c.tA0, n, nh, 0);
        { SegOut o[1] = {seg(SEGM_STORE, ta1)};
          dgrad(l3, M, ny, a3, none(), ACT_NONE, l3 == L_ENC3_LOC ? dy_all(L_ENC3_LOC, t, 0) : none(), o, 1); }
        { SegOut o[1] = {seg(SEGM_STORE, ta0)};
          dgrad(L_ENC2, M, ny, ta1, all ? sig_all(S_ENCB, t, entry) : sig_slot(S_ENCB, t, entry), ACT_ELU,
                all ? dy_all(L_ENC2, t, entry) : dy_slot(L_ENC2, t, entry), o, 1); }
        { SegOut o[1] = {seg(SEGM_STORE, gglm)};
          dgrad(L_ENC1, M, ny, ta0, all ? sig_all(S_ENCA, t, entry) : sig_slot(S_ENCA, t, entry), ACT_ELU,
                all ? dy_all(L_ENC1, t, entry) : dy_slot(L_ENC1, t, entry), o, 1); }
    }

    // ------------------------------------------------------------------------------------------
    // weight gradients: dWv_l[segment rows] = X_seg^T . dY_l over all frames, rows (and slots)
    // ------------------------------------------------------------------------------------------
    struct XSrc {
        const float* p;
        int outer, inner;
    };
    XSrc xs(int sig, int e0, int col, bool per_slot = true) const {
        const Sig s = c.st[sig];
        XSrc x;
        x.p = c.stash + (size_t)s.off + (size_t)e0 * s.width + col;
        x.outer = s.entries * s.width; x.inner = per_slot ? s.width : 0;
        return x;
    }
    void wg(int l, int segi, XSrc x, int ny, int dy_e0 = 0, int frame0 = 0) {
        const Layer& L = plan.L[l];
        const LayerB& LB = plan.LB[l];
        if (L.nhead == 0) return;
        WgradArgs A;
        A.M = T * rows * ny; A.K = L.seg[segi].K; A.N = LB.NU; A.ny = ny;
        A.x = mk_addr(x.p + (size_t)frame0 * rows * x.outer, x.outer, x.inner);
        A.dy = mk_addr(c.dy[l] + (size_t)dy_e0 * c.dy_w[l], c.dy_e[l] * c.dy_w[l], c.dy_w[l]);
        if (!bcast_used_ && ny > 1 && x.inner == 0 && x.outer > 0 && dy_e0 == 0 && c.dy_e[l] == ny && c.dy_w[l] <= nh && A.K >= 32 && LB.NU >= 32) {
            bcast_used_ = true;                    // one scratch buffer: one user per backward call
            // the operand is shared by the slots of a row (the discovery RNN's image / conditioning input): sum dY over the
            // slots first -- a fifth of the reduction length, and rows TMA can address
            float* red = ws + BL.bcast_dy_off;
            be.img_reduce(c.dy[l], red, T * rows, ny, c.dy_w[l]);
            A.M = T * rows; A.ny = 1;
            A.x = mk_addr(x.p + (size_t)frame0 * rows * x.outer, x.outer, 0);
            A.dy = mk_addr(red, c.dy_w[l], 0);
        }
        A.dw = ws + BL.dwv_off + LB.bw_off + (size_t)LB.u0[segi] * LB.NU;
        A.ldw = LB.NU;
        be.wgrad(A);
    }
    void bias(int l, int ny, int dy_e0 = 0) {
        const LayerB& LB = plan.LB[l];
        if (plan.L[l].nhead == 0 || !LB.has_bias) return;
        ColsumArgs A;
        A.M = T * rows * ny; A.N = LB.NU; A.ny = ny;
        A.dy = mk_addr(c.dy[l] + (size_t)dy_e0 * c.dy_w[l], c.dy_e[l] * c.dy_w[l], c.dy_w[l]);
        A.out = ws + BL.dwv_off + LB.bw_off + (size_t)LB.KU * LB.NU;
        be.colsum(A);
    }
    void colsum_to(const float* src, int M, int N, float* out) {
        ColsumArgs A;
        A.M = M; A.N = N; A.ny = 1;
        A.dy = mk_addr(src, N, 0);
        A.out = out;
        be.colsum(A);
    }

    void weight_grads(float* d_params) {
        const RecF& F = plan.rec;
        const bool masked = cfg.masked_glimpse != 0;
        // propagation prior
        wg(L_PGRU_ZR, 0, xs(S_Z, 0, 0), n); wg(L_PGRU_ZR, 1, xs(S_PST, 0, 0), n); bias(L_PGRU_ZR, n);
        wg(L_PGRU_C, 0, xs(S_Z, 0, 0), n); wg(L_PGRU_C, 1, xs(S_PGRH, 0, 0), n); bias(L_PGRU_C, n);
        wg(L_PLIN, 0, xs(S_PSTNEW, 0, 0), n); bias(L_PLIN, n);
        wg(L_WBMK1, 0, xs(S_TST, 0, 0), n); bias(L_WBMK1, n);
        wg(L_WB2, 0, xs(S_HWBMK, 0, 0), n); bias(L_WB2, n);
        if (masked) { wg(L_MK2, 0, xs(S_HWBMK, 0, 128), n); bias(L_MK2, n); }
        // glimpse encoder: three uses per slot pair
        wg(L_ENC1, 0, xs(S_GLM, 0, 0), 3 * n); bias(L_ENC1, 3 * n);
        wg(L_ENC2, 0, xs(S_ENCA, 0, 0), 3 * n); bias(L_ENC2, 3 * n);
        wg(L_ENC3_LOC, 0, xs(S_ENCB, 0, 0), n); bias(L_ENC3_LOC, n);
        wg(L_ENC3, 0, xs(S_ENCB, n, 0), 2 * n, n); bias(L_ENC3, 2 * n, n);
        // propagation core
        wg(L_PRNN, 0, xs(S_ENC, 0, 0), n); wg(L_PRNN, 1, xs(S_PROPREC, 0, 0), n); wg(L_PRNN, 2, xs(S_Z, 0, 0), n);
        wg(L_PRNN, 3, xs(S_TST, 0, 0), n); wg(L_PRNN, 4, xs(S_PH, 0, 0), n); bias(L_PRNN, n);
        wg(L_PT1, 0, xs(S_PH, 1, 0), n); wg(L_PT1, 1, xs(S_Z, 0, nw), n); wg(L_PT1, 2, xs(S_TST, 0, 0), n); bias(L_PT1, n);
        wg(L_PT2, 0, xs(S_PT1, 0, 0), n); bias(L_PT2, n);
        wg(L_PT3, 0, xs(S_PT2, 0, 0), n); bias(L_PT3, n);
        for (int l : {L_TGRU_ZR, L_TGRU_C}) {
            wg(l, 0, xs(S_PH, 1, 0), n); wg(l, 1, xs(S_PROPREC, 1, F.where), n); wg(l, 2, xs(S_ENC, n, 0), n);
            wg(l, 3, xs(l == L_TGRU_ZR ? S_TST : S_TGRH, 0, 0), n); bias(l, n);
        }
        wg(L_PHEADS, 0, xs(S_TSTNEW, 0, 0), n); bias(L_PHEADS, n);
        wg(L_PST1, 0, xs(S_PH, 1, 0), n); wg(L_PST1, 1, xs(S_TST, 0, 0), n); wg(L_PST1, 2, xs(S_PROPREC, 1, F.what), n); bias(L_PST1, n);
        wg(L_PST2, 0, xs(S_PHS, 0, 0), n); bias(L_PST2, n);
        wg(L_LAT1, 0, xs(S_PROPREC, 1, 0), n); bias(L_LAT1, n);
        wg(L_LAT2, 0, xs(S_L1, 0, 0), n); bias(L_LAT2, n);
        // image encoder: the K particles of a sequence share the frame -> reduce dY over particles first
        {
            float* red = ws + BL.img_dy_off;
            be.img_reduce(c.dy[L_IMG1], red, T * cfg.B, cfg.K, nh);
            const LayerB& LB = plan.LB[L_IMG1];
            WgradArgs A;
            A.M = T * cfg.B; A.K = c.PX; A.N = nh; A.ny = 1;
            A.x = mk_addr(c.obs, c.PX, 0);
            A.dy = mk_addr(red, nh, 0);
            A.dw = ws + BL.dwv_off + LB.bw_off; A.ldw = nh;
            be.wgrad(A);
            colsum_to(red, T * cfg.B, nh, ws + BL.dwv_off + LB.bw_off + (size_t)LB.KU * LB.NU);
        }
        wg(L_IMG2, 0, xs(S_IMG1, 0, 0), 1); bias(L_IMG2, 1);
        // discovery core
        wg(L_DRNN, 0, xs(S_DIN, 0, 0, false), n); wg(L_DRNN, 1, xs(S_DISCREC, 0, 0), n); wg(L_DRNN, 2, xs(S_DH, 0, 0), n); bias(L_DRNN, n);
        wg(L_DT1, 0, xs(S_DH, 1, 0), n); bias(L_DT1, n);
        wg(L_DT2, 0, xs(S_DT1, 0, 0), n); bias(L_DT2, n);
        wg(L_DT3, 0, xs(S_DT2, 0, 0), n); bias(L_DT3, n);
        wg(L_DST1, 0, xs(S_DH, 1, 0), n); wg(L_DST1, 1, xs(S_DISCREC, 1, F.what), n); bias(L_DST1, n);
        wg(L_DST2, 0, xs(S_DHS, 0, 0), n); bias(L_DST2, n);
        if (cfg.rec_where_prior) {
            XSrc init; init.p = c.prm + c.po.rn_init_state; init.outer = 0; init.inner = 0;
            wg(L_RN1, 0, init, 1); wg(L_RN1, 1, xs(S_DIN, 0, nh), 1); wg(L_RN1, 2, xs(S_EXP, 0, 0), 1); bias(L_RN1, 1);
            wg(L_RN2, 0, xs(S_RNPREV, 0, 0), n); wg(L_RN2, 1, xs(S_HRN, 0, 0, false), n); bias(L_RN2, n);
            wg(L_RN3, 0, xs(S_RNO, 0, 0), n); bias(L_RN3, n);
        }
        if (cfg.disc_prior_type == SQAIR_DISC_PRIOR_CAT) {
            wg(L_SP1, 0, xs(S_EXP, 0, 0), 1); bias(L_SP1, 1);
            wg(L_SP2, 0, xs(S_HSP, 0, 0), 1); bias(L_SP2, 1);
        }
        // decoder: its input is the compacted what of the same frame = the latents entering frame t + 1
        wg(L_DEC1, 0, xs(S_Z, 0, 0), n, 0, 1); bias(L_DEC1, n);
        wg(L_DEC2, 0, xs(S_D1, 0, 0), n); bias(L_DEC2, n);
        wg(L_DEC3, 0, xs(S_D2, 0, 0), n); bias(L_DEC3, n);
        // virtual matrices -> the reference's variables; per-row quantities -> vector parameters
        be.unpack(ws + BL.dwv_off, d_params);
        const POff& po = c.po;
        colsum_to(c.xt[X_PH0], T * rows, nh, d_params + po.prop_h0);
        colsum_to(c.xt[X_DH0], T * rows, nh, d_params + po.disc_h0);
        colsum_to(c.xt[X_T0], T * rows, nh, d_params + po.temporal_h0);
        colsum_to(c.xt[X_P0], T * rows, nh, d_params + po.prior_h0);
        colsum_to(c.xt[X_MEAN], T * rows, c.PX, d_params + po.mean_img);
        if (cfg.rec_where_prior) {
            colsum_to(c.xt[X_RNINIT], T * rows, 4, d_params + po.rn_init_state);
            colsum_to(c.xt[X_RNSAMPLE], T * rows, 4, d_params + po.rn_init_sample);
        }
        if (cfg.disc_prior_type == SQAIR_DISC_PRIOR_CAT) {
            colsum_to(c.dy[L_SP2], T * rows, n + 1, d_params + po.step_prior_bias);
            if (T > 1) colsum_to(c.dy[L_SP2] + (size_t)rows * (n + 1), (T - 1) * rows, n + 1, d_params + po.step_prior_tbias);
        }
        be.small_to_params(c.small, d_params, po);
    }

    void run(float* d_params) {
        zero(BL.dyz_begin, BL.dyz_end);
        zero(BL.dwv_off, BL.dwv_off + plan.bw_total);
        be.zero(d_params, plan_param_count());
        // the state leaving the last frame receives no gradient
        for (int i = 0; i < 3; ++i) be.zero(ws + BL.carry_off[i], (int64_t)rows * n * (i == 0 ? zw : nh));
        int parity = 0;
        for (int t = T - 1; t >= 0; --t) { frame(t, parity); parity ^= 1; }
        c.gZc = ws + BL.carry_off[parity * 3 + 0]; c.gTc = ws + BL.carry_off[parity * 3 + 1]; c.gPc = ws + BL.carry_off[parity * 3 + 2];
        be.template stage<BS_FINAL_STATES>(c, 0, 0);
        weight_grads(d_params);
    }
    int64_t param_count_ = 0;
    bool bcast_used_ = false;
    int64_t plan_param_count() const { return param_count_; }
};

}  // namespace sq
