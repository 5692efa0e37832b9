// Host-side plan: the static layer graph, arena layouts and workspace map of the hot path.
// Mirrors the construction order of core/architectures.py:30-173 and core/networks.py:37-66,
// 115-137,255-275 (same order as oracle/spec.py; tests assert the layouts agree).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <map>
#include "../../include/cdra.h"

namespace cdra {

struct ArenaTensor {
    std::string name;
    int64_t offset;
    int ndim;
    int dims[4];
    int64_t numel() const { int64_t n = 1; for (int i = 0; i < ndim; ++i) n *= dims[i]; return n; }
};

struct Arena {
    std::vector<ArenaTensor> tensors;
    int64_t size = 0;
    int64_t add(const std::string& name, std::initializer_list<int> dims) {
        ArenaTensor t; t.name = name; t.offset = size; t.ndim = (int)dims.size();
        int i = 0; for (int d : dims) t.dims[i++] = d; for (; i < 4; ++i) t.dims[i] = 1;
        size += t.numel(); tensors.push_back(t); return t.offset;
    }
};

// conv / dense layer followed by BatchNorm: offsets (in floats) into the trainable / state arenas
struct BnConv {
    std::string name;
    int K = 0, N = 0;
    int64_t w = -1, b = -1, g = -1, be = -1;   // trainable arena
    int64_t mm = -1, mv = -1;                  // state arena
    int counter = -1;                          // ticket counter index
    // pointwise convs in bf16 mode: per-step bf16 copies of the weights (workspace byte offsets), see pw_mma.cuh
    bool is_pw = false;
    int split = 0, Kp = 0, Np = 0;
    size_t wt = 0, wn = 0;
};

struct WsTensor {       // activation tensor living in the workspace
    std::string name;
    int Rt = 0;          // rows (pixels x samples) per time slice
    int H = 0, W = 0, C = 0;
    int elem = 4;        // bytes per element
    size_t data = 0, grad = 0;          // byte offsets
    size_t fst = 0, bst = 0, aff = 0, bnp = 0;
    bool tables = false, has_grad = true;
    int bcounter = -1;   // ticket counter of the BN-backward-sums kernel
    size_t bytes() const { return (size_t)4 * Rt * C * elem; }
};

struct Unit {
    std::string name;
    int stride, cin, c, half;
    int Hi, Wi, Ho, Wo, pad_t, pad_l;
    BnConv pw1, dw, pw2, scdw, scpw;
    int t_in, t_r1, t_r2, t_rs, t_out;   // WsTensor indices (t_rs = shortcut depthwise raw, stride 2 only)
};

struct GruSpec {
    std::string name;
    int din, units;
    int64_t k, r, b;                      // trainable arena offsets
    size_t xp, hp, hs, dxp, dhp, dh;      // workspace byte offsets: [4][B][3u], [4][B][3u], [4][B][u], ...
    size_t x_in, dx_in;                   // input sequence [4][B][din] and its gradient
};

struct FeatSpec {
    std::string name;
    int d;
    BnConv d1, d2;
    size_t h1, h2, n1, out, dbuf1, dbuf2; // [4][B][16] fp32 buffers
    size_t st1, st2;                      // per-slice (mean, inv) [4][16] float2
};

struct HeadSpec {       // policy or value head
    Arena params, state;
    int64_t bn1_g, bn1_be, d1_w, d1_b, bn2_g, bn2_be, d2_w, d2_b;
    int64_t out_w[4], out_b[4]; int out_n[4];
    int64_t bn1_mm, bn1_mv, bn2_mm, bn2_mv;
};

// ---- v2 tower (bf16 perf mode): padded plane layout, see v2_common.cuh
struct V2Tensor {
    std::string name;
    int Rt = 0, H = 0, W = 0;
    int cp = 0;                  // slots per row (multiple of 8)
    int n0 = 0, n0p = 0, n1 = 0; // slot map: [part0 n0 valid of n0p | part1 n1 | pad]
    size_t data = 0, grad = 0;   // bf16 [4*Rt][cp]
    size_t fsum = 0, bsum = 0;   // double2 [4][cp] (inside the re-zeroed region)
    size_t aff = 0, bnp = 0;     // float2 [4][cp]
    bool has_bn = true;
    int C() const { return n0 + n1; }
    size_t bytes() const { return (size_t)4 * Rt * cp * 2; }
};
struct V2Pw {                    // one GEMM launch (pw1 / unit tail / head conv)
    int KP = 0, NPall = 0, nplanes = 1, gwp = 0;
    size_t wf = 0, wb = 0, bias = 0;    // workspace byte offsets of the prepared bf16 operands
    size_t wfs = 0, wbs = 0;            // the same operands cut into 64-column blocks of 128-byte-swizzled rows (v4_pwg.cuh)
    int counter = -1, bcounter = -1;
};
struct V2Unit {
    int inA = -1, inB = -1;      // input tensor(s): previous unit's planes, or (inA only) the pool output
    int r1 = -1, r2 = -1, rsA = -1, rsB = -1, outA = -1, outB = -1;
    V2Pw pw1, tail;
    int c_dw = -1, c_scA = -1, c_scB = -1;          // forward ticket counters of the depthwise launches
    int cb_dw = -1, cb_scA = -1, cb_scB = -1;       // backward
};
struct V2Plan {
    bool on = false;
    std::vector<V2Tensor> t;
    std::map<std::string, int> index;
    std::vector<V2Unit> u;
    int p0 = -1, head = -1;      // pool output (plain, already activated), head conv output
    V2Pw head_pw;
    int c_gap = -1;
    size_t desc_off = 0;         // device copies of the GEMM descriptors (uploaded at the start of every pass)
    // tensor-core stem for uint8 frames (v2_stem.cuh): winner positions of the max pool, per-slice fp64 sums of the backward
    bool stem_on = false;
    size_t stem_idx = 0, stem_gacc = 0;
    size_t dr_scratch = 0;       // dR hand-off between the data-gradient and the tcgen05 weight-gradient kernel of one layer
    int stem_fwd_hb = 0, stem_bwd_hb = 0, pool_pb = 0;
    mutable std::vector<char> host_descs_buf;
    void* host_descs = nullptr;
    mutable const char* desc_ws[2] = {nullptr, nullptr};    // workspace whose (forward, backward) descriptor copies match host_descs
};

struct Plan {
    cdra_config cfg;
    V2Plan v2;
    int B, H, W, elem;
    Arena dyn_params, dyn_state;
    HeadSpec policy, value;
    std::vector<WsTensor> tensors;
    std::map<std::string, int> tensor_index;
    BnConv stem, head;
    int Hs, Ws, Hp, Wp, pool_pad_t, pool_pad_l;     // stem / pool output sizes
    int t_stem, t_pool, t_head;
    std::vector<Unit> units;
    std::vector<FeatSpec> feats;
    std::vector<GruSpec> grus;
    int64_t trunk_g, trunk_be, trunk_mm, trunk_mv, trunk_w, trunk_b;
    int n_counters = 0;
    // workspace regions
    size_t zero_bytes = 0;       // [0, zero_bytes): fst/bst tables, re-zeroed at the start of every forward
    size_t counters_off = 0;
    size_t ws_bytes = 0;
    // tail buffers (byte offsets)
    size_t gap, dgap;            // [4][B][768] fp32
    size_t dyn_in, ddyn_in;      // [B][352]
    size_t trunk_n, trunk_stat;  // normalised [B][352]; (mean, inv) [352] float2
    size_t head_buf;             // scratch for the policy/value head (see heads.cuh)
    size_t scratch;              // generic scratch
    std::map<std::string, std::pair<size_t, std::vector<int>>> named;   // extra fp32 buffers for taps
    // side stream (+ fork / join events) for work that is independent of the image tower: the feature MLPs and their
    // GRUs in the forward, the leaf parameter gradients of the GRUs and the feature-MLP backward in the backward.
    // Created on first use (cudaStream_t / cudaEvent_t as void*), destroyed with the plan.
    mutable void* side_stream = nullptr;
    mutable void* ev_fork = nullptr;
    mutable void* ev_join = nullptr;
    mutable void* ev_prep = nullptr;      // GEMM operand preparation (side stream) -> first pointwise kernel of the tower
};

Plan* build_plan(const cdra_config& cfg, std::string& err);

}  // namespace cdra
