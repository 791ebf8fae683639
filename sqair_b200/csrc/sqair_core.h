// sqair_core.h -- host/device shared description of the per-frame schedule.
//
// Plain C++ (no CUDA headers): included by the CUDA translation unit and by the host-side
// kernel-logic emulator under tests/host_emu (test infrastructure only).
//
// The per-frame SQAIR step (reference: sqair/seq.py:181-269 -> sqair/sqair_modules.py:446-582) is
// executed by one thread block per group of R rows (row = b*K + k).  All activations of those rows
// live in shared memory in FEATURE-MAJOR layout x[feature][row] so that a dense layer reads one
// weight per (k, column) from L2 and broadcasts the R activations of feature k to every thread.
// The schedule is data: a `Plan` holds one `Layer` descriptor per dense layer (weight offsets into
// the packed parameter buffer, input segments and output buffers as shared-memory offsets) and is
// passed to the kernel as a __grid_constant__ parameter.
#pragma once
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/sqair_b200.h"

namespace sq {

constexpr int MAXSEG = 5;
constexpr int MAXBLK = 2;
constexpr int NT = 256;          // threads per block
constexpr int MAX_KS = 8;        // max k-slices of a dense layer
constexpr int MAX_SLOTS = 8;

enum Act { ACT_NONE = 0, ACT_ELU = 1, ACT_SIGMOID = 2, ACT_TANH = 3, ACT_SOFTPLUS = 4 };
enum SegKind { SEG_SMEM = 0, SEG_IMAGE = 1 };

struct Seg {
    int x_off;          // smem float offset of x[0][0] (slot 0)
    int x_sstride;      // added per slot index
    int ld;             // floats between consecutive features
    int K;              // number of features
    int kind;           // SegKind
    int w_off[MAXBLK];  // packed-parameter offset of the first weight row of this segment, per block
};

struct Blk {
    int N;              // output columns of this block
    int ldw;            // row stride of its weight matrix
    int b_off, b2_off;  // bias offsets (-1 = none)
    int act, split, act_hi;       // activation for col < split / col >= split
    float scale, add, scale_hi, add_hi;   // v = act(v) * scale + add
    int scale_p_off;    // >= 0: additionally multiply by params[scale_p_off]
    int out_off, out_sstride, out_ld;     // out[col*out_ld + row] (+ slot * out_sstride)
};

struct Layer {
    int nseg, nblk;
    Seg seg[MAXSEG];
    Blk blk[MAXBLK];
};

enum LayerId {
    L_PGRU_ZR, L_PGRU_C, L_PLIN, L_WBMK1, L_WB2, L_MK2, L_ENC1, L_ENC2, L_ENC3_LOC, L_ENC3,
    L_PRNN, L_PT1, L_PT2, L_PT3, L_TGRU_ZR, L_TGRU_C, L_PHEADS, L_PST1, L_PST2,
    L_LAT1, L_LAT2, L_IMG1, L_IMG2, L_DRNN, L_DT1, L_DT2, L_DT3, L_DST1, L_DST2,
    L_RN1, L_RN2, L_RN3, L_SP1, L_SP2, L_DEC1, L_DEC2, L_DEC3, L_COUNT
};

// Feature offsets inside a PropOut / DiscOut entry (a "slot record").
struct RecF {
    int what, where, pres, what_loc, what_scale, where_loc, where_scale, prob, logit, size;
};

// Shared-memory layout (float offsets).  [f][S][R] = feature-major, S slots, R rows.
struct Smem {
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
    int Hrnn;     // [2][nh][R] ping-pong hidden state of the slot RNN
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
    int Hs;       // [128][R]
    int Lg;       // [R]
    int Hrn;      // [128][R]
    int Rno;      // [4][R]
    int Rns;      // [8][R]
    int Hsp;      // [10][R]
    int Spl;      // [n+1][R]
    int Coords;   // [4][R]
    int Dgl;      // [g][NS][R] decoded glimpses (aliases the per-slot scratch)
    int Red;      // k-slice reduction scratch
    int RowAcc;   // [16][R] per-row scalars
    int Perm;     // [2NS][R] compaction order (as floats)
    int total;    // floats
    int red_floats;
};

// Packed-parameter offsets of everything that is not a dense-layer weight/bias.
struct POff {
    int mean_img, output_scale;
    int disc_h0, prop_h0, temporal_h0, prior_h0;
    int rn_init_state, rn_init_sample;
    int d_scale_offset, p_scale_offset;
    int step_prior_bias, step_prior_tbias;
    int cholesky;
};

struct Plan {
    sqair_cfg cfg;
    int R, NS, rows, nw, nh, g, P, LDS;   // LDS = NS*R
    RecF rec;
    Smem sm;
    POff po;
    Layer L[L_COUNT];
};

// ------------------------------------------------------------------------------------------
// Host-side builders
// ------------------------------------------------------------------------------------------
struct ParamEntry {
    std::string name;
    int ndim;
    int shape[3];
    int64_t offset, packed_offset, count;
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

inline int64_t packed_floats(const std::vector<ParamEntry>& v) {
    return (v.back().packed_offset + v.back().count + 3) / 4 * 4 + 4;    // slack for vector tail loads
}

struct PlanBuilder {
    Plan& p;
    const std::vector<ParamEntry>& tab;
    std::string err;
    int cursor = 0;

    PlanBuilder(Plan& plan, const std::vector<ParamEntry>& t) : p(plan), tab(t) {}

    int off(const std::string& name) {
        for (const auto& e : tab)
            if (e.name == name) return (int)e.packed_offset;
        err = "unknown parameter " + name;
        return -1;
    }
    int alloc(int floats) {
        int o = cursor;
        cursor += (floats + 3) / 4 * 4;
        return o;
    }
    Layer& layer(int id, int nblk) {
        Layer& l = p.L[id];
        memset(&l, 0, sizeof(l));
        l.nblk = nblk;
        for (int b = 0; b < MAXBLK; ++b) {
            l.blk[b].b_off = l.blk[b].b2_off = l.blk[b].scale_p_off = -1;
            l.blk[b].scale = l.blk[b].scale_hi = 1.f;
        }
        return l;
    }
    // block b: N columns [col0, col0+N) of a matrix with row stride ldw
    void blk(Layer& l, int b, int N, int ldw, int b_off, int act, int out_off, int out_ld, int out_sstride = 0) {
        Blk& k = l.blk[b];
        k.N = N; k.ldw = ldw; k.b_off = b_off; k.act = act; k.split = N; k.act_hi = act;
        k.out_off = out_off; k.out_ld = out_ld; k.out_sstride = out_sstride;
    }
    // input segment; w0/w1: offset of the weight row where this segment starts in block 0/1
    void seg(Layer& l, int x_off, int ld, int K, int w0, int w1 = -1, int x_sstride = 0, int kind = SEG_SMEM) {
        Seg& s = l.seg[l.nseg++];
        s.x_off = x_off; s.ld = ld; s.K = K; s.kind = kind; s.x_sstride = x_sstride;
        s.w_off[0] = w0; s.w_off[1] = w1;
    }
};

inline int red_need(const Layer& l, int R) {
    int G = 0;
    for (int b = 0; b < l.nblk; ++b) G += (l.blk[b].N + 3) / 4;
    if (G == 0) return 0;            // layer not part of this configuration
    int Gp = (G + 7) / 8 * 8;
    int ks = NT / Gp;
    if (ks > MAX_KS) ks = MAX_KS;
    if (ks <= 1) return 0;
    return (ks - 1) * Gp * 4 * R;
}

// Builds the plan for R rows per block.  Returns "" on success, else an error message.
inline std::string build_plan(const sqair_cfg& c, int R, Plan& p, const std::vector<ParamEntry>& tab) {
    memset(&p, 0, sizeof(p));
    p.cfg = c;
    const int NS = c.n, nw = c.n_what, nh = c.n_hidden, g = c.G * c.G, P = c.H * c.W, s = nh / 2;
    p.R = R; p.NS = NS; p.rows = c.B * c.K; p.nw = nw; p.nh = nh; p.g = g; p.P = P; p.LDS = NS * R;
    RecF& rf = p.rec;
    rf.what = 0; rf.where = nw; rf.pres = nw + 4; rf.what_loc = nw + 5; rf.what_scale = 2 * nw + 5;
    rf.where_loc = 3 * nw + 5; rf.where_scale = 3 * nw + 9; rf.prob = 3 * nw + 13; rf.logit = 3 * nw + 14;
    rf.size = 3 * nw + 15;

    PlanBuilder B(p, tab);
    Smem& m = p.sm;
    const int LDS = NS * R, LDE = (NS + 1) * R;
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
    m.Coords = B.alloc(4 * R);
    m.RowAcc = B.alloc(16 * R);
    m.Perm = B.alloc(2 * NS * R);
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
    if (g * LDS > scratch1 - scratch0) B.cursor = scratch0 + (g * LDS + 3) / 4 * 4;

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

    auto W = [&](const std::string& n) { return B.off(n + "/w"); };
    auto Bi = [&](const std::string& n) { return B.off(n + "/b"); };
    const int zw = m.Z, zwhere = m.Z + nw * LDS;      // Z: what rows 0..nw-1, where nw..nw+3, pres nw+4

    // ---- propagation prior GRU (propagate.py:68-98): x = [what_tm1, where_tm1] (nw+4), h = Pst
    {
        Layer& l = B.layer(L_PGRU_ZR, 2);
        int wz = B.off("propagation/gru_1/wz"), wr = B.off("propagation/gru_1/wr");
        int uz = B.off("propagation/gru_1/uz"), ur = B.off("propagation/gru_1/ur");
        B.blk(l, 0, nh, nh, B.off("propagation/gru_1/bz"), ACT_SIGMOID, m.Gz, R);
        B.blk(l, 1, nh, nh, B.off("propagation/gru_1/br"), ACT_SIGMOID, m.Gr, R);
        B.seg(l, zw, LDS, nw + 4, wz, wr, R);
        B.seg(l, m.Pst, LDS, nh, uz, ur, R);
    }
    {
        Layer& l = B.layer(L_PGRU_C, 1);
        B.blk(l, 0, nh, nh, B.off("propagation/gru_1/bh"), ACT_TANH, m.Gc, R);
        B.seg(l, zw, LDS, nw + 4, B.off("propagation/gru_1/wh"), -1, R);
        B.seg(l, m.Gr, R, nh, B.off("propagation/gru_1/uh"));          // Gr holds r*h
    }
    {
        Layer& l = B.layer(L_PLIN, 1);
        int N = 2 * (4 + nw) + 1;
        B.blk(l, 0, N, N, Bi("propagation/propagate_prior/linear"), ACT_NONE, m.Pri, LDS, R);
        B.seg(l, m.Pst, LDS, nh, W("propagation/propagate_prior/linear"), -1, R);   // Pst already updated in place
    }
    // ---- where-bias MLP (core.py:291) and glimpse-mask MLP (modules.py:322-324): same input (temporal state)
    {
        const bool mk = c.masked_glimpse != 0;
        Layer& l = B.layer(L_WBMK1, mk ? 2 : 1);
        B.blk(l, 0, 128, 128, Bi(PC + "rnn_inpt/mlp/linear"), ACT_ELU, m.Hwb, R);
        if (mk) B.blk(l, 1, 128, 128, Bi(DC + "air_encoder/mlp/linear"), ACT_ELU, m.Hmk, R);
        B.seg(l, m.Tst, LDS, nh, W(PC + "rnn_inpt/mlp/linear"), mk ? W(DC + "air_encoder/mlp/linear") : -1, R);
    }
    {
        Layer& l = B.layer(L_WB2, 1);
        B.blk(l, 0, 4, 4, Bi(PC + "rnn_inpt/mlp/linear_1"), ACT_NONE, m.Wb, R);
        l.blk[0].scale = 0.1f; l.blk[0].scale_hi = 0.1f;
        B.seg(l, m.Hwb, R, 128, W(PC + "rnn_inpt/mlp/linear_1"));
    }
    if (c.masked_glimpse) {
        Layer& l = B.layer(L_MK2, 1);
        B.blk(l, 0, g, g, Bi(DC + "air_encoder/mlp/linear_1"), ACT_SIGMOID, m.Mask, R);
        B.seg(l, m.Hmk, R, 128, W(DC + "air_encoder/mlp/linear_1"));
    }
    // ---- glimpse encoder (modules.py:100-112,358-364), shared by discovery and propagation
    {
        Layer& l = B.layer(L_ENC1, 1);
        B.blk(l, 0, nh, nh, Bi(DC + "encoder_1/mlp/linear"), ACT_ELU, m.A0, R);
        B.seg(l, m.Glm, R, g, W(DC + "encoder_1/mlp/linear"));
    }
    {
        Layer& l = B.layer(L_ENC2, 1);
        B.blk(l, 0, nh, nh, Bi(DC + "encoder_1/mlp/linear_1"), ACT_ELU, m.A1, R);
        B.seg(l, m.A0, R, nh, W(DC + "encoder_1/mlp/linear_1"));
    }
    {
        Layer& l = B.layer(L_ENC3_LOC, 1);     // only .loc is consumed at core.py:293
        B.blk(l, 0, nw, 2 * nw, Bi(DC + "air_encoder/gaussian_from_param_vec/linear"), ACT_NONE, m.Loc1, R);
        B.seg(l, m.A1, R, nh, W(DC + "air_encoder/gaussian_from_param_vec/linear"));
    }
    {
        Layer& l = B.layer(L_ENC3, 1);
        B.blk(l, 0, 2 * nw, 2 * nw, Bi(DC + "air_encoder/gaussian_from_param_vec/linear"), ACT_NONE, m.Enc, R);
        l.blk[0].split = nw; l.blk[0].act_hi = ACT_SOFTPLUS; l.blk[0].add_hi = c.min_std;
        B.seg(l, m.A1, R, nh, W(DC + "air_encoder/gaussian_from_param_vec/linear"));
    }
    // ---- propagation RNN (core.py:295-302): [loc1, km1(what,where,pres), tm1(what,where,pres), temporal] + h
    {
        Layer& l = B.layer(L_PRNN, 1);
        int wi = W("propagation/vanilla_rnn/in_to_hidden");
        B.blk(l, 0, nh, nh, Bi("propagation/vanilla_rnn/in_to_hidden"), ACT_TANH, m.Hrnn + nh * R, R);
        l.blk[0].b2_off = Bi("propagation/vanilla_rnn/hidden_to_hidden");
        B.seg(l, m.Loc1, R, nw, wi);
        B.seg(l, m.PropOut, LDE, nw + 5, wi + nw * nh, -1, R);                 // entry s = record of slot s-1
        B.seg(l, zw, LDS, nw + 5, wi + (2 * nw + 5) * nh, -1, R);
        B.seg(l, m.Tst, LDS, nh, wi + (3 * nw + 10) * nh, -1, R);
        B.seg(l, m.Hrnn, R, nh, W("propagation/vanilla_rnn/hidden_to_hidden"));  // old h in Hrnn[0], new h -> Hrnn[1]
    }
    // ---- propagation transform estimator (core.py:321-327): [h, where_tm1, temporal]
    {
        Layer& l = B.layer(L_PT1, 1);
        int w = W(PC + "stochastic_transform_param/mlp/linear");
        B.blk(l, 0, nh, nh, Bi(PC + "stochastic_transform_param/mlp/linear"), ACT_ELU, m.A0, R);
        B.seg(l, m.Hrnn + nh * R, R, nh, w);
        B.seg(l, zwhere, LDS, 4, w + nh * nh, -1, R);
        B.seg(l, m.Tst, LDS, nh, w + (nh + 4) * nh, -1, R);
    }
    {
        Layer& l = B.layer(L_PT2, 1);
        B.blk(l, 0, nh, nh, Bi(PC + "stochastic_transform_param/mlp/linear_1"), ACT_ELU, m.A1, R);
        B.seg(l, m.A0, R, nh, W(PC + "stochastic_transform_param/mlp/linear_1"));
    }
    {
        Layer& l = B.layer(L_PT3, 1);
        B.blk(l, 0, 8, 8, Bi(PC + "stochastic_transform_param/mlp/linear_2"), ACT_NONE, m.Tp, R);
        B.seg(l, m.A1, R, nh, W(PC + "stochastic_transform_param/mlp/linear_2"));
    }
    // ---- temporal GRU (core.py:339-340): x = [h, where, loc2, scale2], state = Tst
    {
        Layer& l = B.layer(L_TGRU_ZR, 2);
        int wz = B.off("propagation/gru/wz"), wr = B.off("propagation/gru/wr");
        B.blk(l, 0, nh, nh, B.off("propagation/gru/bz"), ACT_SIGMOID, m.Gz, R);
        B.blk(l, 1, nh, nh, B.off("propagation/gru/br"), ACT_SIGMOID, m.Gr, R);
        B.seg(l, m.Hrnn + nh * R, R, nh, wz, wr);
        B.seg(l, m.PropOut + R + rf.where * LDE, LDE, 4, wz + nh * nh, wr + nh * nh, R);      // this slot's where
        B.seg(l, m.Enc, R, 2 * nw, wz + (nh + 4) * nh, wr + (nh + 4) * nh);
        B.seg(l, m.Tst, LDS, nh, B.off("propagation/gru/uz"), B.off("propagation/gru/ur"), R);
    }
    {
        Layer& l = B.layer(L_TGRU_C, 1);
        int wh = B.off("propagation/gru/wh");
        B.blk(l, 0, nh, nh, B.off("propagation/gru/bh"), ACT_TANH, m.Gc, R);
        B.seg(l, m.Hrnn + nh * R, R, nh, wh);
        B.seg(l, m.PropOut + R + rf.where * LDE, LDE, 4, wh + nh * nh, -1, R);
        B.seg(l, m.Enc, R, 2 * nw, wh + (nh + 4) * nh);
        B.seg(l, m.Gr, R, nh, B.off("propagation/gru/uh"));
    }
    // ---- what heads on the new temporal state (core.py:343-349); Gc holds the new temporal state
    {
        Layer& l = B.layer(L_PHEADS, 2);
        B.blk(l, 0, 2 * nw, 2 * nw, Bi(PC + "what/gaussian_from_param_vec/linear"), ACT_NONE, m.Tg, R);
        l.blk[0].split = nw; l.blk[0].act_hi = ACT_SOFTPLUS; l.blk[0].add_hi = c.min_std;
        B.blk(l, 1, 3 * nw, 3 * nw, Bi(PC + "what/linear"), ACT_SIGMOID, m.Gt, R);
        l.blk[1].scale = 0.9999f; l.blk[1].scale_hi = 0.9999f;
        B.seg(l, m.Gc, R, nh, W(PC + "what/gaussian_from_param_vec/linear"), W(PC + "what/linear"));
    }
    // ---- propagation steps predictor (modules.py:506-513): [h, temporal(old), what]
    {
        Layer& l = B.layer(L_PST1, 1);
        int w = W(PC + "steps_predictor/mlp/linear");
        B.blk(l, 0, s, s, Bi(PC + "steps_predictor/mlp/linear"), ACT_ELU, m.Hs, R);
        B.seg(l, m.Hrnn + nh * R, R, nh, w);
        B.seg(l, m.Tst, LDS, nh, w + nh * s, -1, R);
        B.seg(l, m.PropOut + R + rf.what * LDE, LDE, nw, w + 2 * nh * s, -1, R);
    }
    {
        Layer& l = B.layer(L_PST2, 1);
        B.blk(l, 0, 1, 1, Bi(PC + "steps_predictor/mlp/linear_1"), ACT_NONE, m.Lg, R);
        B.seg(l, m.Hs, R, s, W(PC + "steps_predictor/mlp/linear_1"));
    }
    // ---- latent encoder (sqair_modules.py:368-385): [what, where] of a propagated slot
    {
        Layer& l = B.layer(L_LAT1, 1);
        B.blk(l, 0, nh, nh, Bi(SQ + "sqair_timestep/mlp/linear"), ACT_ELU, m.A0, R);
        B.seg(l, m.PropOut + R, LDE, nw + 4, W(SQ + "sqair_timestep/mlp/linear"), -1, R);
    }
    {
        Layer& l = B.layer(L_LAT2, 1);
        B.blk(l, 0, nh, nh, Bi(SQ + "sqair_timestep/mlp/linear_1"), ACT_ELU, m.A1, R);
        B.seg(l, m.A0, R, nh, W(SQ + "sqair_timestep/mlp/linear_1"));
    }
    // ---- image encoder (core.py:165), once per frame
    {
        Layer& l = B.layer(L_IMG1, 1);
        B.blk(l, 0, nh, nh, Bi(DC + "encoder/mlp/linear"), ACT_ELU, m.A0, R);
        B.seg(l, 0, 0, P, W(DC + "encoder/mlp/linear"), -1, 0, SEG_IMAGE);
    }
    {
        Layer& l = B.layer(L_IMG2, 1);
        B.blk(l, 0, nh, nh, Bi(DC + "encoder/mlp/linear_1"), ACT_ELU, m.DIn, R);
        B.seg(l, m.A0, R, nh, W(DC + "encoder/mlp/linear_1"));
    }
    // ---- discovery RNN (core.py:164-176,197-198): [img_enc, conditioning, km1(what,where,pres)] + h
    {
        Layer& l = B.layer(L_DRNN, 1);
        int wi = W("discovery/vanilla_rnn/in_to_hidden");
        B.blk(l, 0, nh, nh, Bi("discovery/vanilla_rnn/in_to_hidden"), ACT_TANH, m.Hrnn + nh * R, R);
        l.blk[0].b2_off = Bi("discovery/vanilla_rnn/hidden_to_hidden");
        B.seg(l, m.DIn, R, 2 * nh, wi);
        B.seg(l, m.DiscOut, LDE, nw + 5, wi + 2 * nh * nh, -1, R);
        B.seg(l, m.Hrnn, R, nh, W("discovery/vanilla_rnn/hidden_to_hidden"));
    }
    {
        Layer& l = B.layer(L_DT1, 1);
        B.blk(l, 0, nh, nh, Bi(DC + "stochastic_transform_param/mlp/linear"), ACT_ELU, m.A0, R);
        B.seg(l, m.Hrnn + nh * R, R, nh, W(DC + "stochastic_transform_param/mlp/linear"));
    }
    {
        Layer& l = B.layer(L_DT2, 1);
        B.blk(l, 0, nh, nh, Bi(DC + "stochastic_transform_param/mlp/linear_1"), ACT_ELU, m.A1, R);
        B.seg(l, m.A0, R, nh, W(DC + "stochastic_transform_param/mlp/linear_1"));
    }
    {
        Layer& l = B.layer(L_DT3, 1);
        B.blk(l, 0, 8, 8, Bi(DC + "stochastic_transform_param/mlp/linear_2"), ACT_NONE, m.Tp, R);
        B.seg(l, m.A1, R, nh, W(DC + "stochastic_transform_param/mlp/linear_2"));
    }
    {
        Layer& l = B.layer(L_DST1, 1);
        int w = W(DC + "steps_predictor/mlp/linear");
        B.blk(l, 0, s, s, Bi(DC + "steps_predictor/mlp/linear"), ACT_ELU, m.Hs, R);
        B.seg(l, m.Hrnn + nh * R, R, nh, w);
        B.seg(l, m.DiscOut + R + rf.what * LDE, LDE, nw, w + nh * s, -1, R);
    }
    {
        Layer& l = B.layer(L_DST2, 1);
        B.blk(l, 0, 1, 1, Bi(DC + "steps_predictor/mlp/linear_1"), ACT_NONE, m.Lg, R);
        B.seg(l, m.Hs, R, s, W(DC + "steps_predictor/mlp/linear_1"));
    }
    // ---- recurrent where prior (modules.py:548-607)
    if (c.rec_where_prior) {
        {
            Layer& l = B.layer(L_RN1, 1);
            int w = W(RN + "linear_1");
            B.blk(l, 0, 128, 128, Bi(RN + "linear_1"), ACT_ELU, m.Hrn, R);
            B.seg(l, m.RnInit, R, 4, w);
            B.seg(l, m.DIn + nh * R, R, nh, w + 4 * 128);          // conditioning from propagation
            B.seg(l, m.Exp, R, 1, w + (4 + nh) * 128);
        }
        {
            Layer& l = B.layer(L_RN2, 1);
            B.blk(l, 0, 4, 4, Bi(RN + "vanilla_rnn/in_to_hidden"), ACT_TANH, m.Rno, R);
            l.blk[0].b2_off = Bi(RN + "vanilla_rnn/hidden_to_hidden");
            B.seg(l, m.RnPrev0, LDS, 4, W(RN + "vanilla_rnn/in_to_hidden"), -1, R);
            B.seg(l, m.Hrn, R, 128, W(RN + "vanilla_rnn/hidden_to_hidden"));
        }
        {
            Layer& l = B.layer(L_RN3, 1);
            B.blk(l, 0, 8, 8, Bi(RN + "linear"), ACT_NONE, m.Rns, R);
            l.blk[0].split = 4; l.blk[0].act_hi = ACT_SOFTPLUS; l.blk[0].add_hi = 1e-2f;
            B.seg(l, m.Rno, R, 4, W(RN + "linear"));
        }
    }
    // ---- step-count prior MLP (sqair_modules.py:217-218)
    {
        Layer& l = B.layer(L_SP1, 1);
        B.blk(l, 0, 10, 10, Bi("discovery/discover/mlp/linear"), ACT_ELU, m.Hsp, R);
        B.seg(l, m.Exp, R, 1, W("discovery/discover/mlp/linear"));
    }
    {
        Layer& l = B.layer(L_SP2, 1);
        B.blk(l, 0, NS + 1, NS + 1, Bi("discovery/discover/mlp/linear_1"), ACT_NONE, m.Spl, R);
        B.seg(l, m.Hsp, R, 10, W("discovery/discover/mlp/linear_1"));
    }
    // ---- glimpse decoder (modules.py:131-147)
    {
        Layer& l = B.layer(L_DEC1, 1);
        B.blk(l, 0, nh, nh, Bi("decoder/air_decoder/decoder/mlp/linear"), ACT_ELU, m.A0, R);
        B.seg(l, zw, LDS, nw, W("decoder/air_decoder/decoder/mlp/linear"), -1, R);
    }
    {
        Layer& l = B.layer(L_DEC2, 1);
        B.blk(l, 0, nh, nh, Bi("decoder/air_decoder/decoder/mlp/linear_1"), ACT_ELU, m.A1, R);
        B.seg(l, m.A0, R, nh, W("decoder/air_decoder/decoder/mlp/linear_1"));
    }
    {
        Layer& l = B.layer(L_DEC3, 1);
        B.blk(l, 0, g, g, Bi("decoder/air_decoder/decoder/mlp/linear_2"), ACT_NONE, m.Dgl, LDS, R);
        l.blk[0].scale_p_off = po.output_scale;
        B.seg(l, m.A1, R, nh, W("decoder/air_decoder/decoder/mlp/linear_2"));
    }

    int red = 0;
    for (int i = 0; i < L_COUNT; ++i) {
        int r = red_need(p.L[i], R);
        if (r > red) red = r;
    }
    m.red_floats = red;
    m.Red = B.alloc(red);
    m.total = B.cursor;
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
