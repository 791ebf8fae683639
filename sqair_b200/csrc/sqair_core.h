// sqair_core.h -- host/device shared description of the per-frame schedule.
//
// Plain C++ (no CUDA headers): included by the CUDA translation unit and by the host-side
// kernel-logic emulator under tests/host_emu (test infrastructure only).
//
// Execution model.  The per-frame SQAIR step (reference: sqair/seq.py:181-269 ->
// sqair/sqair_modules.py:446-582) runs in ONE persistent kernel.  A thread-block CLUSTER of C blocks
// owns R rows (row = b*K + k) for the whole sequence.  Every block of the cluster keeps a full copy
// of the rows' activations in shared memory in FEATURE-MAJOR layout x[feature][row]; a dense layer is
// split by OUTPUT COLUMNS across the C blocks: block c multiplies the activations with its column
// panel of the layer's weight matrix and writes its slice of the result into the shared memory of
// all C blocks (distributed shared memory), so the next layer again finds the full vector locally.
// Weights never sit in shared memory: every weight is used exactly once per block and layer, so each warp
// streams its tensor-core A fragments straight from L2 into registers (LDG.128 on a fragment-ordered copy of
// the panel, several k-steps in flight) and multiplies them with the activations in shared memory by
// mma.sync.m16n8k8 TF32 in three passes (hi*hi + lo*hi + hi*lo, fp32 accumulate: fp32-faithful results).
// Out-features are the MMA M dimension, the block's rows the N dimension (R <= 8 rows ride for free).
//
// The schedule is data: a `Plan` holds one `Layer` descriptor per dense layer (packed weight
// panels, input segments and output heads as shared-memory offsets); the header lives in constant memory,
// the layer table in global memory (staged into shared memory one call ahead).
#pragma once
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/sqair_b200.h"

#ifdef __CUDACC__
#define SQ_HD __host__ __device__ inline
#else
#define SQ_HD inline
#endif

namespace sq {

constexpr int MAXSEG = 6;
constexpr int MAXHEAD = 3;
#ifndef SQAIR_NT
#define SQAIR_NT 384
#endif
constexpr int NT = SQAIR_NT;      // threads per block (12 warps)
constexpr int NT_LAUNCH = NT;
constexpr int NWARP = NT / 32;
constexpr int MAX_KS = NWARP;    // max k-slices of a dense layer
constexpr int MAX_SLOTS = 8;
constexpr int MAXC = 8;          // max cluster size (portable limit)
constexpr int MAXSEQ = 400;      // dense calls per frame
constexpr int MAXR = 8;          // rows per block: the N dimension of one m16n8k8 MMA
constexpr int DESC_WORDS = 104;  // >= sizeof(Layer) / 4, multiple of 4, <= NT (one word per thread when staging)

enum Act { ACT_NONE = 0, ACT_ELU = 1, ACT_SIGMOID = 2, ACT_TANH = 3, ACT_SOFTPLUS = 4 };
enum SegKind { SEG_SMEM = 0, SEG_IMAGE = 1 };

// One input segment: K consecutive rows of the layer's (virtual) weight matrix, multiplied with
// x[k][r] = smem[x_off + slot*x_sstride + k*ld + r]  (or with the frame pixels for SEG_IMAGE).
// In the packed weights every segment is padded with zero rows to a multiple of 8 rows (one MMA k-step never
// straddles two segments); ks0 = index of the segment's first k-step.
struct Seg {
    int x_off, x_sstride, ld, K, kind, ks0;
};

// One output head: columns [col0, col0 + N) of the virtual matrix (col0 is a multiple of 4).
//   out[j*out_ld + r] = (act(acc + b[j] (+ b2[j])) * scale + add) * (scale_p_off >= 0 ? prm[scale_p_off] : 1)
struct Head {
    int col0, N;
    int b_off, b2_off;
    int act;
    float scale, add;
    int scale_p_off;
    int out_off, out_sstride, out_ld;
    // training stash (sqair_forward with a stash buffer): the finished output also goes to global memory at
    //   stash[st_off + ((t * rows + row) * st_entries + entry) * st_width + j];  st_off < 0: not stashed
    int st_off, st_entries, st_width;
};

struct Layer {
    int nseg, nhead;
    Seg seg[MAXSEG];
    Head head[MAXHEAD];
    int Ktot;      // rows of the virtual matrix = sum of segment K (unpadded)
    int ksteps;    // MMA k-steps = sum of ceil(K_seg / 8)
    int Ntot;      // columns of the virtual matrix (heads padded to multiples of 4)
    int split;     // 1: columns split across the cluster; 0: every block computes all columns (tiny layers)
    int Nc;        // columns per panel (multiple of 16)
    int nmt;       // Nc / 16: m-tiles per panel
    int npanel;    // panels with real columns (blocks with rank >= npanel idle in this layer); 1 if !split
    int w_off;     // packed-parameter offset of panel 0 in (hi, lo) fragment order; panel p at w_off + p * panel_floats
    int panel_floats;   // nmt * ksteps * 256
    int w1_off;    // staging copy of the panels in plain fragment order (fp32 sums, written by the pack kernels): panel p at
                   // w1_off + p * panel_floats / 2; split_panels then derives the (hi, lo) copy the kernel reads
    int ksplit;    // k-slices: work units = nmt * ksplit, dealt round-robin to the warps
    int kper;      // k-steps per slice
};

// Fragment order of a panel: [m-tile][k-step][lane][4] floats = the A operand of mma.m16n8k8 (row-major 16x8 tile
// A[m][k] = W[k-step*8 + k][m-tile*16 + m]): lane = (m%8)*4 + k%4 holds a0 = (m, k), a1 = (m+8, k), a2 = (m, k+4),
// a3 = (m+8, k+4).  The copy the kernel reads stores every fragment twice, [m-tile][k-step][hi | lo][lane][4]: the tf32
// split of the weights (split_weight below) is done once at pack time instead of in every block and k-step -- twice
// the L2 traffic for 11 fewer issued instructions per k-step in a loop that is instruction-bound.  Two LDG.128 per
// lane and k-step, 1 KB contiguous per warp.
SQ_HD int frag_off(int ksteps, int mt, int kstep, int m, int k) {
    const int lane = (m & 7) * 4 + (k & 3), q = (m >> 3) + 2 * (k >> 2);
    return ((mt * ksteps + kstep) * 32 + lane) * 4 + q;
}

enum LayerId {
    L_PGRU_ZR, L_PGRU_C, L_PLIN, L_WBMK1, L_WB2, L_MK2, L_ENC1, L_ENC2, L_ENC3_LOC, L_ENC3,
    L_PRNN, L_PT1, L_PT2, L_PT3, L_TGRU_ZR, L_TGRU_C, L_PHEADS, L_PST1, L_PST2,
    L_LAT1, L_LAT2, L_IMG1, L_IMG2, L_DRNN, L_DT1, L_DT2, L_DT3, L_DST1, L_DST2,
    L_RN1, L_RN2, L_RN3, L_SP1, L_SP2, L_DEC1, L_DEC2, L_DEC3, L_COUNT
};

// ------------------------------------------------------------------------------------------
// Training stash: every activation the backward pass needs, written by the forward kernel (row-major per
// (frame, row, entry): the layout the batched backward GEMMs read).  Signal g lives at
//   stash[g.off + ((t * rows + row) * g.entries + e) * g.width + f]
// ------------------------------------------------------------------------------------------
enum SigId {
    S_Z,                                        // [T+1] latents entering frame t (t = T: final): what, where(4), pres, plogit
    S_TST, S_PST,                               // GRU states entering the frame
    S_PGZ, S_PGR, S_PGRH, S_PGC, S_PSTNEW,      // prior GRU: z, r, r*h, candidate, new state
    S_PRI,                                      // post-processed prior statistics
    S_HWBMK,                                    // hidden of the where-bias MLP | hidden of the mask MLP
    S_WB, S_MASK,
    S_GLM, S_ENCA, S_ENCB, S_ENC,               // glimpse encoder, entries = 3n (use * n + slot): prop glimpse 1, prop glimpse 2, discovery
    S_PH,                                       // [n+1] propagation RNN hidden, entry 0 = initial state
    S_PT1, S_PT2, S_TP,
    S_PROPREC,                                  // [n+1] slot records (RecF), entry 0 = initial "previous slot"
    S_TGZ, S_TGR, S_TGRH, S_TGC, S_TSTNEW,      // temporal GRU
    S_TG, S_GT, S_PHS,
    S_L1, S_L2,
    S_IMG1, S_DIN, S_EXP,                       // per row
    S_DH, S_DT1, S_DT2, S_DTP, S_DISCREC, S_DHS,
    S_HRN, S_RNPREV, S_RNO, S_RNS, S_HSP, S_SPL, S_PERM,
    S_D1, S_D2, S_DGL,
    S_COUNT
};

struct Sig {
    int off, entries, width;
};

struct StashLayout {
    Sig s[S_COUNT];
    int frames[S_COUNT];
    int64_t total;         // floats
};

inline StashLayout build_stash(const sqair_cfg& c) {
    StashLayout L;
    memset(&L, 0, sizeof(L));
    const int n = c.n, nw = c.n_what, nh = c.n_hidden, g = c.G * c.G, hs = nh / 2, rows = c.B * c.K;
    const int rec = 3 * nw + 15;
    int64_t cur = 0;
    auto add = [&](int id, int entries, int width, int frames) {
        L.s[id].off = (int)cur; L.s[id].entries = entries; L.s[id].width = width; L.frames[id] = frames;
        cur += (int64_t)frames * rows * entries * width;
        cur = (cur + 3) / 4 * 4;
        if (cur > 0x7fffff00LL) L.total = -1;
    };
    const int T = c.T;
    add(S_Z, n, nw + 6, T + 1);
    add(S_TST, n, nh, T); add(S_PST, n, nh, T);
    add(S_PGZ, n, nh, T); add(S_PGR, n, nh, T); add(S_PGRH, n, nh, T); add(S_PGC, n, nh, T); add(S_PSTNEW, n, nh, T);
    add(S_PRI, n, 2 * (4 + nw) + 1, T);
    add(S_HWBMK, n, 256, T);
    add(S_WB, n, 4, T); add(S_MASK, n, g, T);
    add(S_GLM, 3 * n, g, T); add(S_ENCA, 3 * n, nh, T); add(S_ENCB, 3 * n, nh, T); add(S_ENC, 3 * n, 2 * nw, T);
    add(S_PH, n + 1, nh, T);
    add(S_PT1, n, nh, T); add(S_PT2, n, nh, T); add(S_TP, n, 8, T);
    add(S_PROPREC, n + 1, rec, T);
    add(S_TGZ, n, nh, T); add(S_TGR, n, nh, T); add(S_TGRH, n, nh, T); add(S_TGC, n, nh, T); add(S_TSTNEW, n, nh, T);
    add(S_TG, n, 2 * nw, T); add(S_GT, n, 3 * nw, T); add(S_PHS, n, hs, T);
    add(S_L1, n, nh, T); add(S_L2, n, nh, T);
    add(S_IMG1, 1, nh, T); add(S_DIN, 1, 2 * nh, T); add(S_EXP, 1, 1, T);
    add(S_DH, n + 1, nh, T); add(S_DT1, n, nh, T); add(S_DT2, n, nh, T); add(S_DTP, n, 8, T);
    add(S_DISCREC, n + 1, rec, T); add(S_DHS, n, hs, T);
    add(S_HRN, 1, 128, T); add(S_RNPREV, n, 4, T); add(S_RNO, n, 4, T); add(S_RNS, n, 8, T);
    add(S_HSP, 1, 10, T); add(S_SPL, 1, n + 1, T); add(S_PERM, 1, n, T);
    add(S_D1, n, nh, T); add(S_D2, n, nh, T); add(S_DGL, n, g, T);
    if (L.total == 0) L.total = cur;
    return L;
}

// fp32 -> (hi, lo), both exactly representable in tf32: hi = w truncated to 10 mantissa bits (so w - hi is exact), lo = the
// remainder rounded to nearest.  hi + lo differs from w by <= 2^-22 |w|.
SQ_HD void split_weight(float w, float& hi, float& lo) {
    union { float f; uint32_t u; } a, b;
    a.f = w;
    a.u &= 0xffffe000u;
    hi = a.f;
    b.f = w - hi;
    b.u = (b.u + 0x1000u) & 0xffffe000u;
    lo = b.f;
}

// Feature offsets inside a PropOut / DiscOut entry (a "slot record").
struct RecF {
    int what, where, pres, what_loc, what_scale, where_loc, where_scale, prob, logit, size;
};

// Shared-memory layout (float offsets from the dynamic shared memory base).  [f][S][R] = feature-major.
struct Smem {
    int Ctl;      // [16] ints: reserved (the call counters moved to registers)
    int Desc;     // [2][DESC_WORDS] staged descriptors of the current / next dense call
    int Z;        // [nw+6][NS][R]: what, where(4), pres, plogit     (latents of the previous frame)
    int Ids;      // [NS][R]
    int LastId;   // [R]
    int Tst, Pst; // [nh][NS][R] temporal / prior GRU states
    int PropOut, DiscOut;   // [RecF.size][NS+1][R]; entry 0 = initial "previous slot" record
    int Pri;      // [2(4+nw)+1][NS][R] prior stats: logit, where_loc(4), what_loc, where_scale(4), what_scale
    int Lp;       // [10][NS][R] per-slot log-prob terms
    int RnInit, RnPrev0;    // [4][R], [4][NS][R] (previous-sample inputs of the recurrent where prior)
    int DIn;      // [2nh][R]: image encoding, conditioning
    int Exp;      // [R]
    int Hrnn;     // [2][nh][R] old / new hidden state of the slot RNN
    int Gz, Gr, Gc;         // [nh][R]
    int Hwb, Hmk; // [128][R]
    int Wb;       // [4][R]
    int Mask, Glm;          // [g][R]
    int A0, A1;   // [nh][R] generic hidden buffers
    int Loc1;     // [nw][R]
    int Enc;      // [2nw][R]
    int Tp;       // [8][R]
    int Tg;       // [2nw][R]
    int Gt;       // [3nw][R]
    int Hs;       // [nh/2][R]
    int Lg;       // [R]
    int Hrn;      // [128][R]
    int Rno;      // [4][R]
    int Rns;      // [8][R]
    int Hsp;      // [10][R]
    int Spl;      // [n+1][R]
    int Ones;     // [R] constant 1 (input of the bias row of every dense layer)
    int Coords;   // [4][R]
    int Dgl;      // [g][NS][R] decoded glimpses (aliases the per-slot scratch)
    int Red;      // k-slice partial sums
    int RowAcc;   // [16][R] per-row scalars
    int Perm;     // [2NS][R] compaction order (as floats)
    int Img;      // [img_n][H*W] frames of this block's sequences, staged by TMA at frame start (img_n = 0: read from global)
    int ImgBar;   // mbarrier (8 bytes) of the frame copy
    int img_n;
    int total;    // floats
    int red_floats;
};

// Packed-parameter offsets of everything that is not a dense-layer weight.
struct POff {
    int mean_img, output_scale;
    int disc_h0, prop_h0, temporal_h0, prior_h0;
    int rn_init_state, rn_init_sample;
    int d_scale_offset, p_scale_offset;
    int step_prior_bias, step_prior_tbias;
    int cholesky;
};

// Everything the device program needs except the layer table (which is staged from global memory one call
// ahead): small enough to stay resident in the constant cache.
struct PlanHdr {
    sqair_cfg cfg;
    int R, C, NS, rows, nw, nh, g, PX, LDS;  // LDS = NS*R; C = cluster size; PX = H*W
    int nseq;                                 // dense calls per frame
    int phdr_off;                             // packed-parameter offset of the 4-word layout header {magic, cluster size, n, H*W}
    RecF rec;
    Smem sm;
    POff po;
    Sig st[S_COUNT];                          // training stash signals (build_stash)
    unsigned char seq[MAXSEQ];                // layer ids in program order (one frame)
};

// Host-side view of a layer for the backward pass: the layer's "virtual matrix" without the k-step padding of the
// forward panels.  Rows = concatenated input segments (+ one bias row), columns = concatenated heads.
struct LayerB {
    int u0[MAXSEG];       // first unpadded row of every segment
    int ucol0[MAXHEAD];   // first unpadded column of every head
    int KU, NU;           // rows without the bias row / columns
    int has_bias;
    int64_t bw_off;       // offset of the [KU + 1][NU] row-major matrix in the backward parameter buffer
    int64_t bwt_off;      // offset of its transpose [NU][ldt] (after all row-major matrices); rows padded to 16 bytes
    int ldt;              // round_up(KU + 1, 4)
};

struct Plan : PlanHdr {
    Layer L[L_COUNT];
    LayerB LB[L_COUNT];
    int64_t bw_total;     // floats of the row-major matrices of the backward parameter buffer
    int64_t bwt_total;    // floats of the transposed copies that follow them
    POff poc;             // like PlanHdr::po, but offsets into the CANONICAL flat buffer (backward pass)
};
constexpr uint32_t PACK_MAGIC = 0x53514152u;

// ------------------------------------------------------------------------------------------
// Host-side builders
// ------------------------------------------------------------------------------------------
struct ParamEntry {
    std::string name;
    int ndim;
    int shape[3];
    int64_t offset, packed_offset, count;
};

// One rectangular piece of a layer's virtual weight matrix, copied from a canonical variable.
struct Piece {
    int layer, vrow0, vcol0, K, N;   // vrow0 in the PADDED row space (k-step * 8 + k)
    int64_t src_off;     // canonical offset of element (row0, col0) of the source matrix
    int src_ld;
    int urow0, ucol0;    // the same block in the unpadded virtual matrix of the backward pass (LayerB)
};

inline void add_param(std::vector<ParamEntry>& v, const std::string& name, int d0 = -1, int d1 = -1, int d2 = -1) {
    ParamEntry e;
    e.name = name;
    e.ndim = 0;
    e.shape[0] = e.shape[1] = e.shape[2] = 1;
    int dims[3] = {d0, d1, d2};
    e.count = 1;
    for (int i = 0; i < 3; ++i)
        if (dims[i] >= 0) { e.shape[e.ndim++] = dims[i]; e.count *= dims[i]; }
    e.offset = v.empty() ? 0 : v.back().offset + v.back().count;
    int64_t po = v.empty() ? 0 : v.back().packed_offset + v.back().count;
    e.packed_offset = (po + 3) / 4 * 4;     // 16-byte alignment of every variable
    v.push_back(e);
}

inline void add_linear(std::vector<ParamEntry>& v, const std::string& name, int i, int o) {
    add_param(v, name + "/w", i, o);
    add_param(v, name + "/b", o);
}

// Variable inventory in TF variable order (notebooks/play.ipynb:239-362).
inline std::vector<ParamEntry> param_table(const sqair_cfg& c) {
    std::vector<ParamEntry> v;
    const int n = c.n, nw = c.n_what, nh = c.n_hidden, g = c.G * c.G, P = c.H * c.W, s = nh / 2;
    const std::string RN = "discovery/discover/recurrent_normal_impl/";
    const std::string DC = "discovery/discovery_core/";
    const std::string PC = "propagation/propagation_core/";
    add_param(v, "decoder/air_decoder/Variable", c.H, c.W, 1);
    add_linear(v, "decoder/air_decoder/decoder/mlp/linear", nw, nh);
    add_linear(v, "decoder/air_decoder/decoder/mlp/linear_1", nh, nh);
    add_linear(v, "decoder/air_decoder/decoder/mlp/linear_2", nh, g);
    add_param(v, "decoder/air_decoder/decoder/output_scale");
    add_param(v, "discovery/discover/discovery/vanilla_rnn_initial_state_0/w", 1, nh);
    add_linear(v, "discovery/discover/mlp/linear", 1, 10);
    add_linear(v, "discovery/discover/mlp/linear_1", 10, n + 1);
    if (c.rec_where_prior) {
        add_param(v, RN + "discovery/discover/recurrent_normal_impl/vanilla_rnn_initial_state_0/w", 1, 4);
        add_param(v, RN + "init_sample", 1, 4);
        add_linear(v, RN + "linear", 4, 8);
        add_linear(v, RN + "linear_1", 4 + nh + 1, 128);
        add_linear(v, RN + "vanilla_rnn/hidden_to_hidden", 128, 4);
        add_linear(v, RN + "vanilla_rnn/in_to_hidden", 4, 4);
    }
    add_linear(v, DC + "air_encoder/gaussian_from_param_vec/linear", nh, 2 * nw);
    if (c.masked_glimpse) {
        add_linear(v, DC + "air_encoder/mlp/linear", nh, 128);
        add_linear(v, DC + "air_encoder/mlp/linear_1", 128, g);
    }
    add_linear(v, DC + "encoder/mlp/linear", P, nh);
    add_linear(v, DC + "encoder/mlp/linear_1", nh, nh);
    add_linear(v, DC + "encoder_1/mlp/linear", g, nh);
    add_linear(v, DC + "encoder_1/mlp/linear_1", nh, nh);
    add_linear(v, DC + "steps_predictor/mlp/linear", nh + nw, s);
    add_linear(v, DC + "steps_predictor/mlp/linear_1", s, 1);
    add_linear(v, DC + "stochastic_transform_param/mlp/linear", nh, nh);
    add_linear(v, DC + "stochastic_transform_param/mlp/linear_1", nh, nh);
    add_linear(v, DC + "stochastic_transform_param/mlp/linear_2", nh, 8);
    add_param(v, DC + "stochastic_transform_param/scale_offset");
    add_linear(v, "discovery/vanilla_rnn/hidden_to_hidden", nh, nh);
    add_linear(v, "discovery/vanilla_rnn/in_to_hidden", 2 * nh + nw + 5, nh);
    add_param(v, "model/sequential_air/while/sqair_timestep/discover/step_prior_bias", n + 1);
    add_param(v, "model/sequential_air/while/sqair_timestep/discover/step_prior_timestep_bias", n + 1);
    const char* gates = "zrh";
    for (int which = 0; which < 2; ++which) {
        std::string scope = which == 0 ? "propagation/gru" : "propagation/gru_1";
        int nin = which == 0 ? nh + 4 + 2 * nw : nw + 4;
        for (int gi = 0; gi < 3; ++gi) {
            std::string gname(1, gates[gi]);
            add_param(v, scope + "/w" + gname, nin, nh);
            add_param(v, scope + "/u" + gname, nh, nh);
            add_param(v, scope + "/b" + gname, nh);
        }
    }
    add_linear(v, "propagation/propagate_prior/linear", nh, 2 * (4 + nw) + 1);
    add_param(v, PC + "affine_diag_normal/cholesky_scale", 10);
    add_linear(v, PC + "rnn_inpt/mlp/linear", nh, 128);
    add_linear(v, PC + "rnn_inpt/mlp/linear_1", 128, 4);
    add_linear(v, PC + "steps_predictor/mlp/linear", 2 * nh + nw, s);
    add_linear(v, PC + "steps_predictor/mlp/linear_1", s, 1);
    add_linear(v, PC + "stochastic_transform_param/mlp/linear", 2 * nh + 4, nh);
    add_linear(v, PC + "stochastic_transform_param/mlp/linear_1", nh, nh);
    add_linear(v, PC + "stochastic_transform_param/mlp/linear_2", nh, 8);
    add_param(v, PC + "stochastic_transform_param/scale_offset");
    add_linear(v, PC + "what/gaussian_from_param_vec/linear", nh, 2 * nw);
    add_linear(v, PC + "what/linear", nh, 3 * nw);
    add_param(v, "propagation/sequential_ssm/propagation/vanilla_rnn_initial_state_0/w", 1, nh);
    add_linear(v, "propagation/vanilla_rnn/hidden_to_hidden", nh, nh);
    add_linear(v, "propagation/vanilla_rnn/in_to_hidden", 3 * nw + 10 + nh, nh);
    add_param(v, "sequence/sequential_air/propagation/gru_1_initial_state_0/w", 1, nh);
    add_param(v, "sequence/sequential_air/propagation/gru_initial_state_0/w", 1, nh);
    add_linear(v, "sequence/sequential_air/sqair_timestep/mlp/linear", nw + 4, nh);
    add_linear(v, "sequence/sequential_air/sqair_timestep/mlp/linear_1", nh, nh);
    return v;
}

// floats of the "variables" region of the packed buffer (every variable, 16-byte aligned)
inline int64_t vars_floats(const std::vector<ParamEntry>& v) {
    return (v.back().packed_offset + v.back().count + 31) / 32 * 32;
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct PlanBuilder {
    double unit_cost = 3.0, slice_cost = 0.5;     // k-slicing cost model (in k-steps): start-up of a work unit, one more partial sum
    Plan& p;
    const std::vector<ParamEntry>& tab;
    std::vector<Piece>& pieces;
    std::string err;
    int cursor = 0;          // shared-memory cursor (floats)
    int64_t wcursor = 0;     // packed-parameter cursor (floats)
    int64_t bwcursor = 0;    // backward parameter buffer cursor (floats)
    int64_t bwtcursor = 0;   // cursor of the transposed copies

    PlanBuilder(Plan& plan, const std::vector<ParamEntry>& t, std::vector<Piece>& pc) : p(plan), tab(t), pieces(pc) {
        wcursor = vars_floats(t);
    }

    const ParamEntry* find(const std::string& name) {
        for (const auto& e : tab)
            if (e.name == name) return &e;
        err = "unknown parameter " + name;
        return nullptr;
    }
    int off(const std::string& name) {
        const ParamEntry* e = find(name);
        return e ? (int)e->packed_offset : -1;
    }
    int alloc(int floats, int align = 4) {
        cursor = round_up(cursor, align);
        int o = cursor;
        cursor += round_up(floats, 4);
        return o;
    }
    Layer& layer(int id) {
        Layer& l = p.L[id];
        memset(&l, 0, sizeof(l));
        return l;
    }
    int seg(Layer& l, int x_off, int ld, int K, int x_sstride = 0, int kind = SEG_SMEM) {
        Seg& s = l.seg[l.nseg];
        s.x_off = x_off; s.ld = ld; s.K = K; s.kind = kind; s.x_sstride = x_sstride;
        s.ks0 = l.ksteps;
        p.LB[&l - p.L].u0[l.nseg] = l.Ktot;
        l.Ktot += K;
        l.ksteps += (K + 7) / 8;
        return l.nseg++;
    }
    // canonical offset of a bias vector (name + "/b" style variable given by packed offset lookup is not enough:
    // packing reads the canonical buffer), -1 = none
    int64_t canon(const std::string& name) {
        const ParamEntry* e = find(name);
        return e ? e->offset : -1;
    }
    int head(Layer& l, int N, int b_off, int act, int out_off, int out_ld, int out_sstride = 0) {
        Head& h = l.head[l.nhead];
        h.col0 = l.Ntot; h.N = N; h.b_off = b_off; h.b2_off = -1; h.act = act; h.scale = 1.f; h.add = 0.f;
        h.scale_p_off = -1; h.out_off = out_off; h.out_ld = out_ld; h.out_sstride = out_sstride;
        h.st_off = -1; h.st_entries = 1; h.st_width = 0;
        LayerB& lb = p.LB[&l - p.L];
        lb.ucol0[l.nhead] = lb.NU;
        lb.NU += N;
        l.Ntot += round_up(N, 4);
        return l.nhead++;
    }
    // the head's finished output is also a training-stash signal (columns [col, col + N) of signal `sig`)
    void stash(Layer& l, int h, int sig, int col = 0) {
        l.head[h].st_off = p.st[sig].off + col;
        l.head[h].st_entries = p.st[sig].entries;
        l.head[h].st_width = p.st[sig].width;
    }
    // weights of (segment s, head h) come from rows [row0, row0+K_s) x cols [col0, col0+N_h) of variable `name`
    void w(int id, int s, int h, const std::string& name, int row0, int col0 = 0) {
        const ParamEntry* e = find(name);
        if (!e) return;
        const Layer& l = p.L[id];
        const int vr = l.seg[s].ks0 * 8;          // padded row space: every segment starts on a k-step
        Piece pc;
        pc.layer = id; pc.vrow0 = vr; pc.vcol0 = l.head[h].col0; pc.K = l.seg[s].K; pc.N = l.head[h].N;
        pc.src_ld = e->shape[1];
        pc.src_off = e->offset + (int64_t)row0 * pc.src_ld + col0;
        pc.urow0 = p.LB[id].u0[s]; pc.ucol0 = p.LB[id].ucol0[h];
        if (row0 + pc.K > e->shape[0] || col0 + pc.N > e->shape[1]) err = "piece out of range in " + name;
        pieces.push_back(pc);
    }
    // segments first_seg..last_seg are consecutive row blocks of one matrix (starting at row0), for head h
    void wrows(int id, int h, const std::string& name, int first_seg = 0, int last_seg = -1, int row0 = 0) {
        const Layer& l = p.L[id];
        if (last_seg < 0) last_seg = l.nseg - 1;
        int r = row0;
        for (int s = first_seg; s <= last_seg; ++s) { w(id, s, h, name, r); r += l.seg[s].K; }
    }
    // packed offset of (part of) a variable -> canonical offset
    int64_t packed_to_canonical(int poff) {
        for (const auto& e : tab)
            if (poff >= e.packed_offset && poff < e.packed_offset + e.count) return e.offset + (poff - e.packed_offset);
        err = "bias offset not inside a variable";
        return 0;
    }
    // Finalise a layer: fold the biases into the product (one extra weight row against the constant-1 input),
    // decide the column split and the k-slicing, reserve the packed (fragment-ordered) panels.
    void finish(int id) {
        Layer& l = p.L[id];
        const int C = p.C;
        bool any_bias = false;
        for (int h = 0; h < l.nhead; ++h) any_bias |= (l.head[h].b_off >= 0 || l.head[h].b2_off >= 0);
        LayerB& lb = p.LB[id];
        lb.KU = l.Ktot;
        lb.has_bias = any_bias ? 1 : 0;
        if (any_bias) {
            if (l.nseg >= MAXSEG) { err = "too many segments"; return; }
            const int brow = l.ksteps * 8;
            seg(l, p.sm.Ones, p.R, 1);
            for (int h = 0; h < l.nhead; ++h) {
                int offs[2] = {l.head[h].b_off, l.head[h].b2_off};
                for (int q = 0; q < 2; ++q) {
                    if (offs[q] < 0) continue;
                    Piece pc;
                    pc.layer = id; pc.vrow0 = brow; pc.vcol0 = l.head[h].col0; pc.K = 1; pc.N = l.head[h].N;
                    pc.src_off = packed_to_canonical(offs[q]);
                    pc.src_ld = l.head[h].N;
                    pc.urow0 = lb.KU; pc.ucol0 = lb.ucol0[h];
                    pieces.push_back(pc);
                }
                l.head[h].b_off = l.head[h].b2_off = -1;
            }
        }
        lb.bw_off = bwcursor;
        bwcursor += (int64_t)(lb.KU + 1) * lb.NU;
        bwcursor = (bwcursor + 3) / 4 * 4;
        lb.ldt = round_up(lb.KU + 1, 4);
        lb.bwt_off = bwtcursor;
        bwtcursor += (int64_t)lb.NU * lb.ldt;
        int per = (l.Ntot + C - 1) / C;
        if (C > 1 && per >= 8) {          // at least half an m16 tile of real columns per block
            l.split = 1;
            l.Nc = round_up(per, 16);
            l.npanel = (l.Ntot + l.Nc - 1) / l.Nc;
        } else {
            l.split = 0;
            l.Nc = round_up(l.Ntot, 16);
            l.npanel = 1;
        }
        l.nmt = l.Nc / 16;
        l.panel_floats = l.nmt * l.ksteps * 256;
        l.w1_off = (int)wcursor;
        wcursor += (int64_t)l.npanel * (l.panel_floats / 2);
        wcursor = (wcursor + 31) / 32 * 32;
        l.w_off = (int)wcursor;
        wcursor += (int64_t)l.npanel * l.panel_floats;
        wcursor = (wcursor + 31) / 32 * 32;
        // k-slicing: minimise (units per warp) x (k-steps per unit), with a small charge per extra slice for the
        // cross-warp reduction; slices of fewer than 2 k-steps are not worth it
        int best = 1;
        double best_cost = 1e30;
        for (int ks = 1; ks <= MAX_KS; ++ks) {
            const int kper = (l.ksteps + ks - 1) / ks;
            if (ks > 1 && (kper < 2 || (ks - 1) * kper >= l.ksteps)) continue;
            const int rounds = (l.nmt * ks + NWARP - 1) / NWARP;
            const double cost = (double)rounds * (kper + unit_cost) + slice_cost * ks;
            if (cost < best_cost) { best_cost = cost; best = ks; }
        }
        l.ksplit = best;
        l.kper = (l.ksteps + best - 1) / best;
    }
};

// Dense calls of one frame in program order; must mirror Block::frame() in sqair_device.cuh (the
// emulator asserts it on every call).
inline std::vector<int> frame_sequence(const sqair_cfg& c) {
    std::vector<int> q;
    for (int s = 0; s < c.n; ++s) {
        q.insert(q.end(), {L_PGRU_ZR, L_PGRU_C, L_PLIN, L_WBMK1, L_WB2});
        if (c.masked_glimpse) q.push_back(L_MK2);
        q.insert(q.end(), {L_ENC1, L_ENC2, L_ENC3_LOC, L_PRNN, L_PT1, L_PT2, L_PT3, L_ENC1, L_ENC2, L_ENC3,
                           L_TGRU_ZR, L_TGRU_C, L_PHEADS, L_PST1, L_PST2});
    }
    for (int s = 0; s < c.n; ++s) q.insert(q.end(), {L_LAT1, L_LAT2});
    q.insert(q.end(), {L_IMG1, L_IMG2});
    for (int s = 0; s < c.n; ++s)
        q.insert(q.end(), {L_DRNN, L_DT1, L_DT2, L_DT3, L_ENC1, L_ENC2, L_ENC3, L_DST1, L_DST2});
    if (c.rec_where_prior) {
        q.push_back(L_RN1);
        for (int s = 0; s < c.n; ++s) q.insert(q.end(), {L_RN2, L_RN3});
    }
    if (c.disc_prior_type == SQAIR_DISC_PRIOR_CAT) q.insert(q.end(), {L_SP1, L_SP2});
    for (int s = 0; s < c.n; ++s) q.insert(q.end(), {L_DEC1, L_DEC2, L_DEC3});
    return q;
}

// Builds the plan for R rows per cluster of C blocks.  Returns "" on success, else an error message.
// `pieces` receives the packing table; *packed_total the floats of the packed parameter buffer.
inline std::string build_plan(const sqair_cfg& c, int R, int C, Plan& p, const std::vector<ParamEntry>& tab,
                              std::vector<Piece>& pieces, int64_t* packed_total, bool stage_frame = true,
                              double unit_cost = 3.0, double slice_cost = 0.5) {
    memset((void*)&p, 0, sizeof(p));
    pieces.clear();
    p.cfg = c;
    const int NS = c.n, nw = c.n_what, nh = c.n_hidden, g = c.G * c.G, P = c.H * c.W, s = nh / 2;
    p.R = R; p.C = C; p.NS = NS; p.rows = c.B * c.K; p.nw = nw; p.nh = nh; p.g = g; p.PX = P; p.LDS = NS * R;
    RecF& rf = p.rec;
    rf.what = 0; rf.where = nw; rf.pres = nw + 4; rf.what_loc = nw + 5; rf.what_scale = 2 * nw + 5;
    rf.where_loc = 3 * nw + 5; rf.where_scale = 3 * nw + 9; rf.prob = 3 * nw + 13; rf.logit = 3 * nw + 14;
    rf.size = 3 * nw + 15;

    {
        const StashLayout SL = build_stash(c);
        if (SL.total < 0) return "training stash exceeds 2^31 floats";
        for (int i = 0; i < S_COUNT; ++i) p.st[i] = SL.s[i];
    }
    PlanBuilder B(p, tab, pieces);
    B.unit_cost = unit_cost; B.slice_cost = slice_cost;
    Smem& m = p.sm;
    if (R < 1 || R > MAXR) return "rows per block must be in [1, 8]";
    const int LDS = NS * R, LDE = (NS + 1) * R;
    m.Ctl = B.alloc(16);
    m.Desc = B.alloc(2 * DESC_WORDS);
    m.Z = B.alloc((nw + 6) * LDS);
    m.Ids = B.alloc(LDS);
    m.LastId = B.alloc(R);
    m.Tst = B.alloc(nh * LDS);
    m.Pst = B.alloc(nh * LDS);
    m.PropOut = B.alloc(rf.size * LDE);
    m.DiscOut = B.alloc(rf.size * LDE);
    m.Pri = B.alloc((2 * (4 + nw) + 1) * LDS);
    m.Lp = B.alloc(10 * LDS);
    m.RnInit = B.alloc(4 * R);
    m.RnPrev0 = B.alloc(4 * LDS);
    m.DIn = B.alloc(2 * nh * R);
    m.Exp = B.alloc(R);
    m.Hrnn = B.alloc(2 * nh * R);
    m.A0 = B.alloc(nh * R);
    m.A1 = B.alloc(nh * R);
    m.Loc1 = B.alloc(nw * R);
    m.Enc = B.alloc(2 * nw * R);
    m.Tp = B.alloc(8 * R);
    m.Tg = B.alloc(2 * nw * R);
    m.Gt = B.alloc(3 * nw * R);
    m.Hs = B.alloc(s * R);
    m.Lg = B.alloc(R);
    m.Hrn = B.alloc(128 * R);
    m.Rno = B.alloc(4 * R);
    m.Rns = B.alloc(8 * R);
    m.Hsp = B.alloc(10 * R);
    m.Spl = B.alloc((NS + 1) * R);
    m.Ones = B.alloc(R);
    m.Coords = B.alloc(4 * R);
    m.RowAcc = B.alloc(16 * R);
    m.Perm = B.alloc(2 * NS * R);
    // Frames in shared memory: the rows of a cluster belong to at most img_n consecutive sequences (row = b*K + k).
    // A bulk copy needs 16-byte sizes and addresses, i.e. H*W a multiple of 4.
    m.img_n = 0;
    m.ImgBar = B.alloc(4, 4);
    if (stage_frame && P % 4 == 0) {
        const int rows = c.B * c.K;
        for (int r0 = 0; r0 < rows; r0 += R) {
            const int r1 = (r0 + R - 1 < rows - 1) ? (r0 + R - 1) : (rows - 1);
            const int cnt = r1 / c.K - r0 / c.K + 1;
            if (cnt > m.img_n) m.img_n = cnt;
        }
        m.Img = B.alloc(m.img_n * P, 4);
    }
    m.Wb = B.alloc(4 * R);
    // per-slot scratch; the decoded glimpses alias it (it is dead once the slots are compacted)
    int scratch0 = B.cursor;
    m.Gz = B.alloc(nh * R);
    m.Gr = B.alloc(nh * R);
    m.Gc = B.alloc(nh * R);
    m.Hwb = B.alloc(128 * R);
    m.Hmk = B.alloc(128 * R);
    m.Mask = B.alloc(g * R);
    m.Glm = B.alloc(g * R);
    int scratch1 = B.cursor;
    m.Dgl = scratch0;
    if (g * LDS > scratch1 - scratch0) B.cursor = scratch0 + round_up(g * LDS, 4);

    POff& po = p.po;
    const std::string RN = "discovery/discover/recurrent_normal_impl/";
    const std::string DC = "discovery/discovery_core/";
    const std::string PC = "propagation/propagation_core/";
    const std::string SQ = "sequence/sequential_air/";
    po.mean_img = B.off("decoder/air_decoder/Variable");
    po.output_scale = B.off("decoder/air_decoder/decoder/output_scale");
    po.disc_h0 = B.off("discovery/discover/discovery/vanilla_rnn_initial_state_0/w");
    po.prop_h0 = B.off("propagation/sequential_ssm/propagation/vanilla_rnn_initial_state_0/w");
    po.temporal_h0 = B.off(SQ + "propagation/gru_initial_state_0/w");
    po.prior_h0 = B.off(SQ + "propagation/gru_1_initial_state_0/w");
    po.rn_init_state = c.rec_where_prior ? B.off(RN + "discovery/discover/recurrent_normal_impl/vanilla_rnn_initial_state_0/w") : -1;
    po.rn_init_sample = c.rec_where_prior ? B.off(RN + "init_sample") : -1;
    po.d_scale_offset = B.off(DC + "stochastic_transform_param/scale_offset");
    po.p_scale_offset = B.off(PC + "stochastic_transform_param/scale_offset");
    po.step_prior_bias = B.off("model/sequential_air/while/sqair_timestep/discover/step_prior_bias");
    po.step_prior_tbias = B.off("model/sequential_air/while/sqair_timestep/discover/step_prior_timestep_bias");
    po.cholesky = B.off(PC + "affine_diag_normal/cholesky_scale");
    {
        POff& q = p.poc;
        auto CO = [&](const std::string& n) { return (int)B.canon(n); };
        q.mean_img = CO("decoder/air_decoder/Variable");
        q.output_scale = CO("decoder/air_decoder/decoder/output_scale");
        q.disc_h0 = CO("discovery/discover/discovery/vanilla_rnn_initial_state_0/w");
        q.prop_h0 = CO("propagation/sequential_ssm/propagation/vanilla_rnn_initial_state_0/w");
        q.temporal_h0 = CO(SQ + "propagation/gru_initial_state_0/w");
        q.prior_h0 = CO(SQ + "propagation/gru_1_initial_state_0/w");
        q.rn_init_state = c.rec_where_prior ? CO(RN + "discovery/discover/recurrent_normal_impl/vanilla_rnn_initial_state_0/w") : -1;
        q.rn_init_sample = c.rec_where_prior ? CO(RN + "init_sample") : -1;
        q.d_scale_offset = CO(DC + "stochastic_transform_param/scale_offset");
        q.p_scale_offset = CO(PC + "stochastic_transform_param/scale_offset");
        q.step_prior_bias = CO("model/sequential_air/while/sqair_timestep/discover/step_prior_bias");
        q.step_prior_tbias = CO("model/sequential_air/while/sqair_timestep/discover/step_prior_timestep_bias");
        q.cholesky = CO(PC + "affine_diag_normal/cholesky_scale");
    }

    auto Bi = [&](const std::string& n) { return B.off(n + "/b"); };
    const int zw = m.Z, zwhere = m.Z + nw * LDS;      // Z: what rows 0..nw-1, where nw..nw+3, pres nw+4
    const int Hnew = m.Hrnn + nh * R;                 // new hidden state of the slot RNN
    // a plain MLP layer: one segment, one head, one weight matrix
    auto simple = [&](int id, int x_off, int x_ld, int K, int x_ss, const std::string& lin, int N, int act, int out_off,
                      int out_ld, int out_ss = 0) -> Layer& {
        Layer& l = B.layer(id);
        B.seg(l, x_off, x_ld, K, x_ss);
        B.head(l, N, Bi(lin), act, out_off, out_ld, out_ss);
        B.w(id, 0, 0, lin + "/w", 0);
        return l;
    };

    // ---- propagation prior GRU (propagate.py:68-98): x = [what_tm1, where_tm1] (nw+4), h = Pst
    {
        Layer& l = B.layer(L_PGRU_ZR);
        B.seg(l, zw, LDS, nw + 4, R);
        B.seg(l, m.Pst, LDS, nh, R);
        B.head(l, nh, B.off("propagation/gru_1/bz"), ACT_SIGMOID, m.Gz, R);
        B.head(l, nh, B.off("propagation/gru_1/br"), ACT_SIGMOID, m.Gr, R);
        B.w(L_PGRU_ZR, 0, 0, "propagation/gru_1/wz", 0); B.w(L_PGRU_ZR, 1, 0, "propagation/gru_1/uz", 0);
        B.w(L_PGRU_ZR, 0, 1, "propagation/gru_1/wr", 0); B.w(L_PGRU_ZR, 1, 1, "propagation/gru_1/ur", 0);
        B.stash(l, 0, S_PGZ); B.stash(l, 1, S_PGR);
        B.finish(L_PGRU_ZR);
    }
    {
        Layer& l = B.layer(L_PGRU_C);
        B.seg(l, zw, LDS, nw + 4, R);
        B.seg(l, m.Gr, R, nh);                          // Gr holds r*h
        B.head(l, nh, B.off("propagation/gru_1/bh"), ACT_TANH, m.Gc, R);
        B.w(L_PGRU_C, 0, 0, "propagation/gru_1/wh", 0); B.w(L_PGRU_C, 1, 0, "propagation/gru_1/uh", 0);
        B.stash(l, 0, S_PGC);
        B.finish(L_PGRU_C);
    }
    simple(L_PLIN, m.Pst, LDS, nh, R, "propagation/propagate_prior/linear", 2 * (4 + nw) + 1, ACT_NONE, m.Pri, LDS, R);
    B.finish(L_PLIN);                                   // Pst already updated in place
    // ---- where-bias MLP (core.py:291) and glimpse-mask MLP (modules.py:322-324): same input (temporal state)
    {
        Layer& l = B.layer(L_WBMK1);
        B.seg(l, m.Tst, LDS, nh, R);
        B.head(l, 128, Bi(PC + "rnn_inpt/mlp/linear"), ACT_ELU, m.Hwb, R);
        B.w(L_WBMK1, 0, 0, PC + "rnn_inpt/mlp/linear/w", 0);
        if (c.masked_glimpse) {
            B.head(l, 128, Bi(DC + "air_encoder/mlp/linear"), ACT_ELU, m.Hmk, R);
            B.w(L_WBMK1, 0, 1, DC + "air_encoder/mlp/linear/w", 0);
            B.stash(l, 1, S_HWBMK, 128);
        }
        B.stash(l, 0, S_HWBMK, 0);
        B.finish(L_WBMK1);
    }
    simple(L_WB2, m.Hwb, R, 128, 0, PC + "rnn_inpt/mlp/linear_1", 4, ACT_NONE, m.Wb, R).head[0].scale = 0.1f;
    B.stash(p.L[L_WB2], 0, S_WB);
    B.finish(L_WB2);
    if (c.masked_glimpse) {
        simple(L_MK2, m.Hmk, R, 128, 0, DC + "air_encoder/mlp/linear_1", g, ACT_SIGMOID, m.Mask, R);
        B.stash(p.L[L_MK2], 0, S_MASK);
        B.finish(L_MK2);
    }
    // ---- glimpse encoder (modules.py:100-112,358-364), shared by discovery and propagation
    simple(L_ENC1, m.Glm, R, g, 0, DC + "encoder_1/mlp/linear", nh, ACT_ELU, m.A0, R);
    B.stash(p.L[L_ENC1], 0, S_ENCA);
    B.finish(L_ENC1);
    simple(L_ENC2, m.A0, R, nh, 0, DC + "encoder_1/mlp/linear_1", nh, ACT_ELU, m.A1, R);
    B.stash(p.L[L_ENC2], 0, S_ENCB);
    B.finish(L_ENC2);
    {   // only .loc is consumed at core.py:293: first nw columns of the [nh, 2nw] head
        Layer& l = B.layer(L_ENC3_LOC);
        B.seg(l, m.A1, R, nh);
        B.head(l, nw, Bi(DC + "air_encoder/gaussian_from_param_vec/linear"), ACT_NONE, m.Loc1, R);
        B.w(L_ENC3_LOC, 0, 0, DC + "air_encoder/gaussian_from_param_vec/linear/w", 0);
        B.stash(l, 0, S_ENC, 0);
        B.finish(L_ENC3_LOC);
    }
    auto gauss_heads = [&](int id, Layer& l, const std::string& lin, int out_off) {    // (loc | softplus(scale)+min_std)
        const int b = Bi(lin);
        B.head(l, nw, b, ACT_NONE, out_off, R);
        int h1 = B.head(l, nw, b + nw, ACT_SOFTPLUS, out_off + nw * R, R);
        l.head[h1].add = c.min_std;
        B.w(id, 0, 0, lin + "/w", 0, 0);
        B.w(id, 0, 1, lin + "/w", 0, nw);
    };
    {
        Layer& l = B.layer(L_ENC3);
        B.seg(l, m.A1, R, nh);
        gauss_heads(L_ENC3, l, DC + "air_encoder/gaussian_from_param_vec/linear", m.Enc);
        B.stash(l, 0, S_ENC, 0); B.stash(l, 1, S_ENC, nw);
        B.finish(L_ENC3);
    }
    // ---- propagation RNN (core.py:295-302): [loc1, km1(what,where,pres), tm1(what,where,pres), temporal] + h
    {
        Layer& l = B.layer(L_PRNN);
        B.seg(l, m.Loc1, R, nw);
        B.seg(l, m.PropOut, LDE, nw + 5, R);            // entry s = record of slot s-1
        B.seg(l, zw, LDS, nw + 5, R);
        B.seg(l, m.Tst, LDS, nh, R);
        B.seg(l, m.Hrnn, R, nh);                        // old h in Hrnn[0], new h -> Hrnn[1]
        int h = B.head(l, nh, Bi("propagation/vanilla_rnn/in_to_hidden"), ACT_TANH, Hnew, R);
        l.head[h].b2_off = Bi("propagation/vanilla_rnn/hidden_to_hidden");
        B.wrows(L_PRNN, 0, "propagation/vanilla_rnn/in_to_hidden/w", 0, 3);
        B.w(L_PRNN, 4, 0, "propagation/vanilla_rnn/hidden_to_hidden/w", 0);
        B.stash(l, h, S_PH);
        B.finish(L_PRNN);
    }
    // ---- propagation transform estimator (core.py:321-327): [h, where_tm1, temporal]
    {
        Layer& l = B.layer(L_PT1);
        B.seg(l, Hnew, R, nh);
        B.seg(l, zwhere, LDS, 4, R);
        B.seg(l, m.Tst, LDS, nh, R);
        B.head(l, nh, Bi(PC + "stochastic_transform_param/mlp/linear"), ACT_ELU, m.A0, R);
        B.wrows(L_PT1, 0, PC + "stochastic_transform_param/mlp/linear/w");
        B.stash(l, 0, S_PT1);
        B.finish(L_PT1);
    }
    simple(L_PT2, m.A0, R, nh, 0, PC + "stochastic_transform_param/mlp/linear_1", nh, ACT_ELU, m.A1, R);
    B.stash(p.L[L_PT2], 0, S_PT2);
    B.finish(L_PT2);
    simple(L_PT3, m.A1, R, nh, 0, PC + "stochastic_transform_param/mlp/linear_2", 8, ACT_NONE, m.Tp, R);
    B.stash(p.L[L_PT3], 0, S_TP);
    B.finish(L_PT3);
    // ---- temporal GRU (core.py:339-340): x = [h, where, loc2, scale2], state = Tst
    const int where_cur = m.PropOut + R + rf.where * LDE;          // this slot's where (entry s+1)
    {
        Layer& l = B.layer(L_TGRU_ZR);
        B.seg(l, Hnew, R, nh);
        B.seg(l, where_cur, LDE, 4, R);
        B.seg(l, m.Enc, R, 2 * nw);
        B.seg(l, m.Tst, LDS, nh, R);
        B.head(l, nh, B.off("propagation/gru/bz"), ACT_SIGMOID, m.Gz, R);
        B.head(l, nh, B.off("propagation/gru/br"), ACT_SIGMOID, m.Gr, R);
        B.wrows(L_TGRU_ZR, 0, "propagation/gru/wz", 0, 2); B.w(L_TGRU_ZR, 3, 0, "propagation/gru/uz", 0);
        B.wrows(L_TGRU_ZR, 1, "propagation/gru/wr", 0, 2); B.w(L_TGRU_ZR, 3, 1, "propagation/gru/ur", 0);
        B.stash(l, 0, S_TGZ); B.stash(l, 1, S_TGR);
        B.finish(L_TGRU_ZR);
    }
    {
        Layer& l = B.layer(L_TGRU_C);
        B.seg(l, Hnew, R, nh);
        B.seg(l, where_cur, LDE, 4, R);
        B.seg(l, m.Enc, R, 2 * nw);
        B.seg(l, m.Gr, R, nh);
        B.head(l, nh, B.off("propagation/gru/bh"), ACT_TANH, m.Gc, R);
        B.wrows(L_TGRU_C, 0, "propagation/gru/wh", 0, 2); B.w(L_TGRU_C, 3, 0, "propagation/gru/uh", 0);
        B.stash(l, 0, S_TGC);
        B.finish(L_TGRU_C);
    }
    // ---- what heads on the new temporal state (core.py:343-349); Gc holds the new temporal state
    {
        Layer& l = B.layer(L_PHEADS);
        B.seg(l, m.Gc, R, nh);
        gauss_heads(L_PHEADS, l, PC + "what/gaussian_from_param_vec/linear", m.Tg);
        int h = B.head(l, 3 * nw, Bi(PC + "what/linear"), ACT_SIGMOID, m.Gt, R);
        l.head[h].scale = 0.9999f;
        B.w(L_PHEADS, 0, h, PC + "what/linear/w", 0);
        B.stash(l, 0, S_TG, 0); B.stash(l, 1, S_TG, nw); B.stash(l, h, S_GT);
        B.finish(L_PHEADS);
    }
    // ---- propagation steps predictor (modules.py:506-513): [h, temporal(old), what]
    {
        Layer& l = B.layer(L_PST1);
        B.seg(l, Hnew, R, nh);
        B.seg(l, m.Tst, LDS, nh, R);
        B.seg(l, m.PropOut + R + rf.what * LDE, LDE, nw, R);
        B.head(l, s, Bi(PC + "steps_predictor/mlp/linear"), ACT_ELU, m.Hs, R);
        B.wrows(L_PST1, 0, PC + "steps_predictor/mlp/linear/w");
        B.stash(l, 0, S_PHS);
        B.finish(L_PST1);
    }
    simple(L_PST2, m.Hs, R, s, 0, PC + "steps_predictor/mlp/linear_1", 1, ACT_NONE, m.Lg, R);
    B.finish(L_PST2);
    // ---- latent encoder (sqair_modules.py:368-385): [what, where] of a propagated slot
    simple(L_LAT1, m.PropOut + R, LDE, nw + 4, R, SQ + "sqair_timestep/mlp/linear", nh, ACT_ELU, m.A0, R);
    B.stash(p.L[L_LAT1], 0, S_L1);
    B.finish(L_LAT1);
    simple(L_LAT2, m.A0, R, nh, 0, SQ + "sqair_timestep/mlp/linear_1", nh, ACT_ELU, m.A1, R);
    B.stash(p.L[L_LAT2], 0, S_L2);
    B.finish(L_LAT2);
    // ---- image encoder (core.py:165), once per frame
    {
        Layer& l = B.layer(L_IMG1);
        B.seg(l, 0, 0, P, 0, SEG_IMAGE);
        B.head(l, nh, Bi(DC + "encoder/mlp/linear"), ACT_ELU, m.A0, R);
        B.w(L_IMG1, 0, 0, DC + "encoder/mlp/linear/w", 0);
        B.stash(l, 0, S_IMG1);
        B.finish(L_IMG1);
    }
    simple(L_IMG2, m.A0, R, nh, 0, DC + "encoder/mlp/linear_1", nh, ACT_ELU, m.DIn, R);
    B.stash(p.L[L_IMG2], 0, S_DIN, 0);
    B.finish(L_IMG2);
    // ---- discovery RNN (core.py:164-176,197-198): [img_enc, conditioning, km1(what,where,pres)] + h
    {
        Layer& l = B.layer(L_DRNN);
        B.seg(l, m.DIn, R, 2 * nh);
        B.seg(l, m.DiscOut, LDE, nw + 5, R);
        B.seg(l, m.Hrnn, R, nh);
        int h = B.head(l, nh, Bi("discovery/vanilla_rnn/in_to_hidden"), ACT_TANH, Hnew, R);
        l.head[h].b2_off = Bi("discovery/vanilla_rnn/hidden_to_hidden");
        B.wrows(L_DRNN, 0, "discovery/vanilla_rnn/in_to_hidden/w", 0, 1);
        B.w(L_DRNN, 2, 0, "discovery/vanilla_rnn/hidden_to_hidden/w", 0);
        B.stash(l, h, S_DH);
        B.finish(L_DRNN);
    }
    simple(L_DT1, Hnew, R, nh, 0, DC + "stochastic_transform_param/mlp/linear", nh, ACT_ELU, m.A0, R);
    B.stash(p.L[L_DT1], 0, S_DT1);
    B.finish(L_DT1);
    simple(L_DT2, m.A0, R, nh, 0, DC + "stochastic_transform_param/mlp/linear_1", nh, ACT_ELU, m.A1, R);
    B.stash(p.L[L_DT2], 0, S_DT2);
    B.finish(L_DT2);
    simple(L_DT3, m.A1, R, nh, 0, DC + "stochastic_transform_param/mlp/linear_2", 8, ACT_NONE, m.Tp, R);
    B.stash(p.L[L_DT3], 0, S_DTP);
    B.finish(L_DT3);
    {
        Layer& l = B.layer(L_DST1);
        B.seg(l, Hnew, R, nh);
        B.seg(l, m.DiscOut + R + rf.what * LDE, LDE, nw, R);
        B.head(l, s, Bi(DC + "steps_predictor/mlp/linear"), ACT_ELU, m.Hs, R);
        B.wrows(L_DST1, 0, DC + "steps_predictor/mlp/linear/w");
        B.stash(l, 0, S_DHS);
        B.finish(L_DST1);
    }
    simple(L_DST2, m.Hs, R, s, 0, DC + "steps_predictor/mlp/linear_1", 1, ACT_NONE, m.Lg, R);
    B.finish(L_DST2);
    // ---- recurrent where prior (modules.py:548-607)
    if (c.rec_where_prior) {
        {
            Layer& l = B.layer(L_RN1);
            B.seg(l, m.RnInit, R, 4);
            B.seg(l, m.DIn + nh * R, R, nh);            // conditioning from propagation
            B.seg(l, m.Exp, R, 1);
            B.head(l, 128, Bi(RN + "linear_1"), ACT_ELU, m.Hrn, R);
            B.wrows(L_RN1, 0, RN + "linear_1/w");
            B.stash(l, 0, S_HRN);
            B.finish(L_RN1);
        }
        {
            Layer& l = B.layer(L_RN2);
            B.seg(l, m.RnPrev0, LDS, 4, R);
            B.seg(l, m.Hrn, R, 128);
            int h = B.head(l, 4, Bi(RN + "vanilla_rnn/in_to_hidden"), ACT_TANH, m.Rno, R);
            l.head[h].b2_off = Bi(RN + "vanilla_rnn/hidden_to_hidden");
            B.w(L_RN2, 0, 0, RN + "vanilla_rnn/in_to_hidden/w", 0);
            B.w(L_RN2, 1, 0, RN + "vanilla_rnn/hidden_to_hidden/w", 0);
            B.stash(l, h, S_RNO);
            B.finish(L_RN2);
        }
        {
            Layer& l = B.layer(L_RN3);
            B.seg(l, m.Rno, R, 4);
            const int b = Bi(RN + "linear");
            B.head(l, 4, b, ACT_NONE, m.Rns, R);
            int h1 = B.head(l, 4, b + 4, ACT_SOFTPLUS, m.Rns + 4 * R, R);
            l.head[h1].add = 1e-2f;
            B.w(L_RN3, 0, 0, RN + "linear/w", 0, 0);
            B.w(L_RN3, 0, 1, RN + "linear/w", 0, 4);
            B.stash(l, 0, S_RNS, 0); B.stash(l, h1, S_RNS, 4);
            B.finish(L_RN3);
        }
    }
    // ---- step-count prior MLP (sqair_modules.py:217-218)
    simple(L_SP1, m.Exp, R, 1, 0, "discovery/discover/mlp/linear", 10, ACT_ELU, m.Hsp, R);
    B.stash(p.L[L_SP1], 0, S_HSP);
    B.finish(L_SP1);
    simple(L_SP2, m.Hsp, R, 10, 0, "discovery/discover/mlp/linear_1", NS + 1, ACT_NONE, m.Spl, R);
    B.stash(p.L[L_SP2], 0, S_SPL);
    B.finish(L_SP2);
    // ---- glimpse decoder (modules.py:131-147)
    simple(L_DEC1, zw, LDS, nw, R, "decoder/air_decoder/decoder/mlp/linear", nh, ACT_ELU, m.A0, R);
    B.stash(p.L[L_DEC1], 0, S_D1);
    B.finish(L_DEC1);
    simple(L_DEC2, m.A0, R, nh, 0, "decoder/air_decoder/decoder/mlp/linear_1", nh, ACT_ELU, m.A1, R);
    B.stash(p.L[L_DEC2], 0, S_D2);
    B.finish(L_DEC2);
    simple(L_DEC3, m.A1, R, nh, 0, "decoder/air_decoder/decoder/mlp/linear_2", g, ACT_NONE, m.Dgl, LDS, R)
        .head[0].scale_p_off = po.output_scale;
    B.stash(p.L[L_DEC3], 0, S_DGL);
    B.finish(L_DEC3);

    // reduction scratch: every k-slice parks its partial sums, [ksplit][Nc][R]
    int red = 0;
    for (int i = 0; i < L_COUNT; ++i) {
        if (p.L[i].nhead == 0) continue;
        int r = p.L[i].ksplit * p.L[i].Nc * R;
        if (r > red) red = r;
    }
    if (red < 16 * R) red = 16 * R;                        // also used by the block reduction of the likelihood
    m.red_floats = red;
    m.Red = B.alloc(red);
    B.alloc(8 * (NS + 1) * R + 64);                        // slack: B fragments over-read at most 7 feature rows
    m.total = B.cursor;

    std::vector<int> q = frame_sequence(c);
    if ((int)q.size() > MAXSEQ) return "too many dense calls per frame";
    p.nseq = (int)q.size();
    for (size_t i = 0; i < q.size(); ++i) p.seq[i] = (unsigned char)q[i];
    static_assert(sizeof(Layer) <= DESC_WORDS * 4, "DESC_WORDS too small");
    static_assert(DESC_WORDS <= NT, "descriptor staging uses one thread per word");
    p.bw_total = B.bwcursor;
    p.bwt_total = B.bwtcursor;
    p.phdr_off = (int)((B.wcursor + 31) / 32 * 32);
    if (packed_total) *packed_total = (B.wcursor + 31) / 32 * 32 + 32;
    return B.err;
}

inline std::string validate_cfg(const sqair_cfg& c) {
    char buf[256];
    if (c.T < 1 || c.B < 1 || c.K < 1) return "T, B, K must be >= 1";
    if (c.n < 1 || c.n > MAX_SLOTS) { snprintf(buf, sizeof buf, "n_steps_per_image must be in [1, %d]", MAX_SLOTS); return buf; }
    if (c.H < 2 || c.W < 2 || c.G < 2) return "H, W, glimpse_size must be >= 2";
    if (c.n_what < 1 || c.n_what > 248) return "n_what must be in [1, 248]";
    if (c.n_hidden < 8 || c.n_hidden % 4 != 0) return "n_hidden must be a positive multiple of 4";
    if (c.prior_type < 0 || c.prior_type > 2) return "Invalid prior type. Choose from {rnn, rw, guided}.";
    if (c.disc_prior_type < 0 || c.disc_prior_type > 1) return "Invalid prior type: choose from {cat, geom}";
    if (!(c.output_std > 0.f) || !(c.bg_std > 0.f)) return "output_std and bg_std must be > 0";
    return "";
}

}  // namespace sq
