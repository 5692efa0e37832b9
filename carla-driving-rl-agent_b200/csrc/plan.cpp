// Plan construction (host only).  See plan.h.
#include "plan.h"
#include <cstdlib>
#include <algorithm>
#include <cstdio>

namespace cdra {

static const int kStageC[3] = {116, 232, 464};      // core/architectures.py:34 (g = 1.0)
static const int kStageBlocks[3] = {4, 8, 4};       // core/architectures.py:165-167
static const int kStemC = 24, kLastC = 768;         // :159 ; core/carla_agent.py:66
static const int kFeatUnits = 16, kTrunkIn = 352, kTrunkUnits = 512, kHeadUnits = 320;

static void same_pad(int n, int k, int s, int& out, int& before) {
    out = (n + s - 1) / s;
    int total = (out - 1) * s + k - n;
    if (total < 0) total = 0;
    before = total / 2;                               // TF puts the extra cell after
}

static BnConv add_bnconv(Plan& p, const std::string& name, std::initializer_list<int> wdims, int K, int N) {
    BnConv l; l.name = name; l.K = K; l.N = N;
    l.w = p.dyn_params.add(name + ".w", wdims);
    l.b = p.dyn_params.add(name + ".b", {N});
    l.g = p.dyn_params.add(name + ".g", {N});
    l.be = p.dyn_params.add(name + ".be", {N});
    l.mm = p.dyn_state.add(name + ".mm", {N});
    l.mv = p.dyn_state.add(name + ".mv", {N});
    l.counter = p.n_counters++;
    return l;
}

static int add_tensor(Plan& p, const std::string& name, int H, int W, int C, bool tables, bool has_grad = true) {
    WsTensor t; t.name = name; t.H = H; t.W = W; t.C = C; t.Rt = p.B * H * W; t.elem = p.elem;
    t.tables = tables; t.has_grad = has_grad;
    if (tables) t.bcounter = p.n_counters++;
    p.tensor_index[name] = (int)p.tensors.size();
    p.tensors.push_back(t);
    return (int)p.tensors.size() - 1;
}

static void build_head(HeadSpec& h, bool policy) {
    h.bn1_g = h.params.add("bn1.g", {kTrunkUnits}); h.bn1_be = h.params.add("bn1.be", {kTrunkUnits});
    h.bn1_mm = h.state.add("bn1.mm", {kTrunkUnits}); h.bn1_mv = h.state.add("bn1.mv", {kTrunkUnits});
    h.d1_w = h.params.add("d1.w", {kTrunkUnits, kHeadUnits}); h.d1_b = h.params.add("d1.b", {kHeadUnits});
    h.bn2_g = h.params.add("bn2.g", {kHeadUnits}); h.bn2_be = h.params.add("bn2.be", {kHeadUnits});
    h.bn2_mm = h.state.add("bn2.mm", {kHeadUnits}); h.bn2_mv = h.state.add("bn2.mv", {kHeadUnits});
    h.d2_w = h.params.add("d2.w", {kHeadUnits, kHeadUnits}); h.d2_b = h.params.add("d2.b", {kHeadUnits});
    const char* pn[4] = {"alpha", "beta", "similarity", "speed"};
    const char* vn[4] = {"base", "exp", "speed", "similarity"};
    const int pnn[4] = {2, 2, 1, 1}, vnn[4] = {1, 1, 1, 1};
    for (int i = 0; i < 4; ++i) {
        std::string n = policy ? pn[i] : vn[i];
        h.out_n[i] = policy ? pnn[i] : vnn[i];
        h.out_w[i] = h.params.add(n + ".w", {kHeadUnits, h.out_n[i]});
        h.out_b[i] = h.params.add(n + ".b", {h.out_n[i]});
    }
}

Plan* build_plan(const cdra_config& cfg, std::string& err) {
    if (cfg.batch < 1 || cfg.height < 19 || cfg.width < 19) { err = "bad batch / image size"; return nullptr; }
    if (cfg.dtype != CDRA_DTYPE_F32 && cfg.dtype != CDRA_DTYPE_BF16) { err = "bad dtype"; return nullptr; }
    Plan* pp = new Plan();
    Plan& p = *pp;
    p.cfg = cfg; p.B = cfg.batch; p.H = cfg.height; p.W = cfg.width;
    p.elem = cfg.dtype == CDRA_DTYPE_BF16 ? 2 : 4;

    // ---- layer graph + arenas (order == oracle/spec.py::dynamics_params)
    p.stem = add_bnconv(p, "tower.stem", {3, 3, 3, kStemC}, 27, kStemC);
    p.Hs = (p.H - 3) / 2 + 1; p.Ws = (p.W - 3) / 2 + 1;
    same_pad(p.Hs, 3, 2, p.Hp, p.pool_pad_t); same_pad(p.Ws, 3, 2, p.Wp, p.pool_pad_l);
    p.t_stem = add_tensor(p, "tower.stem", p.Hs, p.Ws, kStemC, true);
    p.t_pool = add_tensor(p, "tower.pool", p.Hp, p.Wp, kStemC, false);
    int cin = kStemC, h = p.Hp, w = p.Wp, t_prev = p.t_pool;
    for (int s = 0; s < 3; ++s) {
        const int c = kStageC[s];
        for (int u = 0; u < kStageBlocks[s]; ++u) {
            Unit un; char buf[64]; snprintf(buf, sizeof buf, "tower.s%d.u%d", s + 1, u);
            un.name = buf; un.stride = u == 0 ? 2 : 1; un.cin = cin; un.c = c; un.half = c / 2;
            un.Hi = h; un.Wi = w;
            if (un.stride == 2) { same_pad(h, 3, 2, un.Ho, un.pad_t); same_pad(w, 3, 2, un.Wo, un.pad_l); }
            else { un.Ho = h; un.Wo = w; un.pad_t = un.pad_l = 1; }
            const int sc = un.stride == 2 ? cin : cin / 2;       // shortcut_channels, :127
            const int kin = un.stride == 2 ? cin : cin / 2;
            un.pw1 = add_bnconv(p, un.name + ".pw1", {kin, un.half}, kin, un.half);
            un.dw = add_bnconv(p, un.name + ".dw", {3, 3, un.half}, 9, un.half);
            un.pw2 = add_bnconv(p, un.name + ".pw2", {un.half, c - sc}, un.half, c - sc);
            if (un.stride == 2) {
                un.scdw = add_bnconv(p, un.name + ".scdw", {3, 3, sc}, 9, sc);
                un.scpw = add_bnconv(p, un.name + ".scpw", {sc, sc}, sc, sc);
            }
            un.t_in = t_prev;
            un.t_r1 = add_tensor(p, un.name + ".pw1", h, w, un.half, true);
            un.t_r2 = add_tensor(p, un.name + ".dw", un.Ho, un.Wo, un.half, true);
            un.t_rs = un.stride == 2 ? add_tensor(p, un.name + ".scdw", un.Ho, un.Wo, sc, true) : -1;
            un.t_out = add_tensor(p, un.name + ".out", un.Ho, un.Wo, c, true);
            p.units.push_back(un);
            t_prev = un.t_out; cin = c; h = un.Ho; w = un.Wo;
        }
    }
    p.head = add_bnconv(p, "tower.head", {kStageC[2], kLastC}, kStageC[2], kLastC);
    p.t_head = add_tensor(p, "tower.head", h, w, kLastC, true);

    const char* fnames[3] = {"road", "vehicle", "navigation"};
    const int fd[3] = {9, 4, 5};
    for (int m = 0; m < 3; ++m) {
        FeatSpec f; f.name = fnames[m]; f.d = fd[m];
        f.d1 = add_bnconv(p, std::string("feat.") + fnames[m] + ".d1", {fd[m], kFeatUnits}, fd[m], kFeatUnits);
        f.d2 = add_bnconv(p, std::string("feat.") + fnames[m] + ".d2", {kFeatUnits, kFeatUnits}, kFeatUnits, kFeatUnits);
        p.feats.push_back(f);
    }
    const char* gnames[4] = {"image", "road", "vehicle", "navigation"};
    const int gdin[4] = {kLastC, 16, 16, 16}, gun[4] = {256, 32, 32, 32};
    for (int g = 0; g < 4; ++g) {
        GruSpec s; s.name = gnames[g]; s.din = gdin[g]; s.units = gun[g];
        s.k = p.dyn_params.add(std::string("gru.") + gnames[g] + ".k", {gdin[g], 3 * gun[g]});
        s.r = p.dyn_params.add(std::string("gru.") + gnames[g] + ".r", {gun[g], 3 * gun[g]});
        s.b = p.dyn_params.add(std::string("gru.") + gnames[g] + ".b", {2, 3 * gun[g]});
        p.grus.push_back(s);
    }
    p.trunk_g = p.dyn_params.add("trunk.bn.g", {kTrunkIn});
    p.trunk_be = p.dyn_params.add("trunk.bn.be", {kTrunkIn});
    p.trunk_mm = p.dyn_state.add("trunk.bn.mm", {kTrunkIn});
    p.trunk_mv = p.dyn_state.add("trunk.bn.mv", {kTrunkIn});
    p.trunk_w = p.dyn_params.add("trunk.dense.w", {kTrunkIn, kTrunkUnits});
    p.trunk_b = p.dyn_params.add("trunk.dense.b", {kTrunkUnits});
    build_head(p.policy, true);
    build_head(p.value, false);

    // ---- v2 tower tensors (bf16 perf mode): padded planes, see v2_common.cuh
    p.v2.on = (p.elem == 2);
    // (the v2 depthwise kernels band frames that do not fit shared memory -- 180x240 at stage 1 -- so every geometry stays on v2)
    if (p.v2.on) {
        V2Plan& v = p.v2;
        auto r8 = [](int x) { return (x + 7) / 8 * 8; };
        auto r16 = [](int x) { return (x + 15) / 16 * 16; };
        auto addT = [&](const std::string& name, int H, int W, int n0, int n0p, int n1, bool has_bn) {
            V2Tensor t; t.name = name; t.H = H; t.W = W; t.Rt = p.B * H * W; t.n0 = n0; t.n0p = n0p; t.n1 = n1;
            t.cp = r8(n0p + n1); t.has_bn = has_bn;
            v.index[name] = (int)v.t.size(); v.t.push_back(t);
            return (int)v.t.size() - 1;
        };
        auto plain = [&](const std::string& name, int H, int W, int C, bool bn = true) { return addT(name, H, W, C, r8(C), 0, bn); };
        auto like = [&](const std::string& name, int H, int W, int src) {
            const V2Tensor s = v.t[src]; return addT(name, H, W, s.n0, s.n0p, s.n1, true);
        };
        auto pw = [&](V2Pw& g, int ksum, int nplanes, int gwp) {
            g.KP = r16(ksum); g.nplanes = nplanes; g.gwp = gwp; g.NPall = nplanes * gwp;
            g.counter = p.n_counters++; g.bcounter = p.n_counters++;
        };
        v.p0 = plain("tower.pool", p.Hp, p.Wp, kStemC, false);
        int inA = v.p0, inB = -1;
        for (const Unit& un : p.units) {
            V2Unit u;
            u.inA = inA; u.inB = inB;
            u.r1 = plain(un.name + ".pw1", un.Hi, un.Wi, un.half);
            u.r2 = plain(un.name + ".dw", un.Ho, un.Wo, un.half);
            u.c_dw = p.n_counters++; u.cb_dw = p.n_counters++;
            int n0, n0p, n1;
            if (un.stride == 2) {
                u.rsA = like(un.name + ".scdwA", un.Ho, un.Wo, inA);
                u.c_scA = p.n_counters++; u.cb_scA = p.n_counters++;
                if (inB >= 0) { u.rsB = like(un.name + ".scdwB", un.Ho, un.Wo, inB); u.c_scB = p.n_counters++; u.cb_scB = p.n_counters++; }
                n0 = (un.c - un.cin) / 2; n0p = (n0 + 1) / 2 * 2; n1 = un.cin / 2;
                pw(u.pw1, v.t[inA].cp + (inB >= 0 ? v.t[inB].cp : 0), 1, v.t[u.r1].cp);
            } else {
                n0 = un.half / 2; n0p = (n0 + 1) / 2 * 2; n1 = un.half / 2;
                pw(u.pw1, v.t[inB].cp, 1, v.t[u.r1].cp);
            }
            u.outA = addT(un.name + ".outA", un.Ho, un.Wo, n0, n0p, n1, true);
            u.outB = addT(un.name + ".outB", un.Ho, un.Wo, n0, n0p, n1, true);
            if (un.stride == 2) pw(u.tail, v.t[u.r2].cp + v.t[u.rsA].cp + (u.rsB >= 0 ? v.t[u.rsB].cp : 0), 2, v.t[u.outA].cp);
            else pw(u.tail, v.t[u.r2].cp, 2, r8(n0p));
            v.u.push_back(u);
            inA = u.outA; inB = u.outB;
        }
        v.head = plain("tower.head", v.t[inA].H, v.t[inA].W, kLastC);
        pw(v.head_pw, v.t[inA].cp + v.t[inB].cp, 1, v.t[v.head].cp);
        v.c_gap = p.n_counters++;
    }

    // ---- workspace map
    size_t off = 0;
    auto alloc = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t kCopies = 4;       // == cdra::kStatCopies (cdra_common.cuh): replicated fp64 sums
    auto legacy = [&](const WsTensor& t) { return !p.v2.on || t.name == "tower.stem" || t.name == "tower.pool"; };
    for (auto& t : p.tensors) if (t.tables && legacy(t)) { t.fst = alloc(kCopies * 4 * t.C * 16); t.bst = alloc(kCopies * 4 * t.C * 16); }
    for (auto& t : p.v2.t) { t.fsum = alloc((size_t)4 * t.cp * 16); t.bsum = alloc((size_t)4 * t.cp * 16); }
    if (p.v2.on) p.v2.stem_gacc = alloc((size_t)4 * 2048 * 8);      // >= 4 * v2::kGaccN doubles
    p.zero_bytes = off;
    p.counters_off = alloc((size_t)(p.n_counters + 16) * 4);
    for (auto& t : p.tensors) if (t.tables && legacy(t)) { t.aff = alloc((size_t)4 * t.C * 8); t.bnp = alloc((size_t)4 * t.C * 8); }
    for (auto& t : p.v2.t) { t.aff = alloc((size_t)4 * t.cp * 8); t.bnp = alloc((size_t)4 * t.cp * 8); }
    {   // bf16 weight copies for the tensor-core pointwise kernels
        auto pw = [&](BnConv& l, int split) {
            l.is_pw = true; l.split = split;
            l.Kp = (l.K + 31) / 32 * 32; l.Np = (l.N + 31) / 32 * 32;
            if (p.elem == 2 && !p.v2.on) {
                l.wt = alloc((size_t)((l.N + 7) / 8 * 8) * l.Kp * 2);
                l.wn = alloc((size_t)((l.K + 7) / 8 * 8) * l.Np * 2);
            }
        };
        for (auto& u : p.units) { pw(u.pw1, 0); pw(u.pw2, 1); if (u.stride == 2) pw(u.scpw, 1); }
        pw(p.head, 0);
    }
    for (auto& t : p.tensors) if (legacy(t)) { t.data = alloc(t.bytes()); }
    for (auto& t : p.tensors) if (legacy(t)) { t.grad = t.has_grad ? alloc(t.bytes()) : 0; }
    if (p.v2.on) {
        V2Plan& v = p.v2;
        for (auto& t : v.t) {
            if (t.name == "tower.pool") { t.data = p.tensors[p.t_pool].data; t.grad = p.tensors[p.t_pool].grad; continue; }
            t.data = alloc(t.bytes() + 256);          // +256: TMA tiles / vector tails never leave the allocation
        }
        for (auto& t : v.t) if (t.name != "tower.pool") t.grad = alloc(t.bytes() + 256);
        auto pwalloc = [&](V2Pw& g) {
            g.wf = alloc((size_t)g.NPall * g.KP * 2); g.wb = alloc((size_t)g.NPall * g.KP * 2); g.bias = alloc((size_t)g.NPall * 4);
            g.wfs = alloc((size_t)((g.KP + 63) / 64) * g.NPall * 128 + 16384); g.wbs = alloc((size_t)((g.NPall + 63) / 64) * g.KP * 128 + 16384);
        };
        for (auto& u : v.u) { pwalloc(u.pw1); pwalloc(u.tail); }
        pwalloc(v.head_pw);
        v.stem_idx = alloc((size_t)4 * p.B * p.Hp * p.Wp * kStemC + 256);
        // the tensor-core stem needs 16-byte aligned TMA rows of the uint8 frames and 8-bit pixel coordinates
        v.stem_on = cfg.image_u8 != 0 && p.W % 8 == 0 && ((size_t)p.H * p.W * 3) % 16 == 0 && p.Ws < 256 && getenv("CDRA_LEGACY_STEM") == nullptr;
        {   // largest [4*Rt][NPall] bf16 gradient matrix of any GEMM launch
            size_t mx = 0;
            for (auto& u : v.u) {
                mx = std::max(mx, (size_t)4 * v.t[u.r1].Rt * u.pw1.NPall * 2);
                mx = std::max(mx, (size_t)4 * v.t[u.outA].Rt * u.tail.NPall * 2);
            }
            mx = std::max(mx, (size_t)4 * v.t[v.head].Rt * v.head_pw.NPall * 2);
            v.dr_scratch = alloc(mx + 256);
        }
        v.desc_off = alloc(64 * 1024);
        v.host_descs_buf.resize(64 * 1024);
        v.host_descs = v.host_descs_buf.data();
    }
    auto f32 = [&](const std::string& name, std::vector<int> dims) {
        size_t n = 1; for (int d : dims) n *= d;
        size_t o = alloc(n * 4);
        p.named[name] = {o, dims};
        return o;
    };
    const int B = p.B;
    p.gap = f32("tower.gap", {4, B, kLastC});
    p.dgap = f32("d.tower.gap", {4, B, kLastC});
    for (auto& f : p.feats) {
        f.h1 = f32("feat." + f.name + ".h1", {4, B, kFeatUnits});
        f.h2 = f32("feat." + f.name + ".h2", {4, B, kFeatUnits});
        f.n1 = f32("feat." + f.name + ".n1", {4, B, kFeatUnits});
        f.out = f32("feat." + f.name + ".out", {4, B, kFeatUnits});
        f.dbuf1 = f32("d.feat." + f.name + ".out", {4, B, kFeatUnits});
        f.dbuf2 = f32("d.feat." + f.name + ".tmp", {4, B, kFeatUnits});
        f.st1 = f32("feat." + f.name + ".st1", {4, kFeatUnits, 2});
        f.st2 = f32("feat." + f.name + ".st2", {4, kFeatUnits, 2});
    }
    for (size_t g = 0; g < p.grus.size(); ++g) {
        GruSpec& s = p.grus[g];
        const int u3 = 3 * s.units;
        s.xp = f32("gru." + s.name + ".xp", {4, B, u3});
        s.hp = f32("gru." + s.name + ".hp", {4, B, u3});
        s.hs = f32("gru." + s.name + ".hs", {4, B, s.units});
        s.dxp = f32("d.gru." + s.name + ".xp", {4, B, u3});
        s.dhp = f32("d.gru." + s.name + ".hp", {4, B, u3});
        s.dh = f32("d.gru." + s.name + ".h", {2, B, s.units});
        s.x_in = g == 0 ? p.gap : p.feats[g - 1].out;
        s.dx_in = g == 0 ? p.dgap : p.feats[g - 1].dbuf1;
    }
    p.dyn_in = f32("dynamics_in", {B, kTrunkIn});
    p.ddyn_in = f32("d.dynamics_in", {B, kTrunkIn});
    p.trunk_n = f32("trunk.n", {B, kTrunkIn});
    f32("d.trunk.n", {B, kTrunkIn});
    p.trunk_stat = f32("trunk.stat", {kTrunkIn, 2});
    // head scratch: n1[512] pre1 a1 n2 pre2 a2 [320 each] + their gradients + stats
    f32("head.n1", {B, kTrunkUnits}); f32("head.pre1", {B, kHeadUnits}); f32("head.a1", {B, kHeadUnits});
    f32("head.n2", {B, kHeadUnits}); f32("head.pre2", {B, kHeadUnits}); f32("head.a2", {B, kHeadUnits});
    f32("head.st1", {kTrunkUnits, 2}); f32("head.st2", {kHeadUnits, 2});
    f32("d.head.a2", {B, kHeadUnits}); f32("d.head.pre2", {B, kHeadUnits}); f32("d.head.n2", {B, kHeadUnits});
    f32("d.head.a1", {B, kHeadUnits}); f32("d.head.pre1", {B, kHeadUnits}); f32("d.head.n1", {B, kTrunkUnits});
    f32("head.dlogits", {B, 8});
    f32("head.acc", {64});          // fp64 x 32 loss accumulators
    p.scratch = f32("scratch", {4, B < 4 ? 4 : B, kLastC});
    f32("wgrad.partials", {1024, 64, 64});       // per-CTA partial weight-gradient tiles (pw_mma.cuh)
    p.ws_bytes = off;
    return pp;
}

}  // namespace cdra
