// v3: ONE warp-specialised tcgen05 kernel for the whole backward of a pointwise (1x1) convolution layer --
// what tape.gradient computes through Conv2D(1x1) + BatchNormalization + ReLU6 (core/architectures.py:130,134,140;
// core/carla_agent.py:361-365): BatchNorm(+ReLU6) backward of the layer's output gradient, data gradient, weight
// gradient, pass-through gradient of a stride-1 unit, and the BatchNorm-backward sums of every gradient it writes.
//
// Per row tile (R rows, contiguous in HBM):
//   TMA producer (1 thread)   bulk copies of d out / out (both planes), the raw source rows (+ x1, + the partial gradient
//                             of a second consumer) into a multi-stage ring; mbarrier complete_tx
//   transform warps (6)       dR = scale * (dz - S1/n - xhat * S2/n) and act(src) = relu6(scale * raw + shift), written
//                             straight into 128-byte-swizzled tiles (double buffered)
//   MMA issuer (1 thread)     dSrc^T [kk][r] = sum_j Wb[kk][j] dR[r][j]     A = weights (K-major, resident), B = dR tile
//                                                                          (K-major), D in TMEM (double buffered)
//                             dW [kk][j]   += sum_r act(src)[r][kk] dR[r][j]  A = act tile, B = the SAME dR tile (both
//                                                                          MN-major), D resident in TMEM for the CTA's life
//   epilogue warps (8)        tcgen05.ld of dSrc^T: a thread owns ONE channel kk and walks the tile's rows, so the
//                             BatchNorm-backward sums of the source accumulate in registers with no cross-thread
//                             reduction; bf16 rows are staged and leave through one TMA bulk store per source
// No CTA-wide barrier inside the loop: the roles are chained by mbarriers only (ring full/empty, staging full/empty,
// TMEM full/empty).  Replaces pw_dgrad_kernel + the dR hand-off matrix + pw_wgrad_tc_kernel of v2_bwd.cuh for every
// layer whose accumulators fit the 512 TMEM columns (stage 1 and stage 2 except the stage-2 stride-2 tail).
#pragma once
#ifndef CDRA_EMU
#include "v2_bwd.cuh"
#include "v2_stem.cuh"

namespace cdra {
namespace v2 {

CDRA_DEV void bulk_s2g_nc(void* dst, const void* src, uint32_t bytes) {      // bulk store shared -> global, not yet committed
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
CDRA_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
CDRA_DEV void bulk_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// role timeline of block 0 (see v4_pwg.cuh / cdra_debug_timeline)
__device__ unsigned long long g_bf_ts[32];
CDRA_DEV unsigned long long gtimer3() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define BF_TS(i) do { if (a.timeline && blockIdx.x == 0) g_bf_ts[i] = gtimer3(); } while (0)
constexpr int kBfThreads = 512, kBfTransformWarps = 6, kBfEpilogueWarps = 8, kBfMaxStages = 8;
constexpr int kBfTransformThreads = kBfTransformWarps * 32, kBfEpilogueThreads = kBfEpilogueWarps * 32;
constexpr int kBfGroupWarps = 4, kBfGroupThreads = kBfGroupWarps * 32;     // one epilogue GROUP (4 warps = the 4 TMEM lane blocks) per tile; two groups alternate tiles

struct PwBfSmem {
    int maps, mdesc, w, dr, xs, st, st2, ring, total;
    int stage_bytes, dr_bytes, xs_bytes, w_bytes;
    int o_dout, o_out, o_src[kMaxSrc], o_x1, o_ge[kMaxSrc], st_off[kMaxSrc];
    int mbk, np16, nkb, cols_dw, tmem_cols;
    int st_bytes, st2_bytes;
};
// `cps[i]`, `acc[i]`: slots / accumulate flag of source i
inline __host__ __device__ PwBfSmem pw_bf_smem(int R, int nsrc, const int* cps, const int* acc, int NPall, int nplanes, int cpo, int x1cp, int nstage) {
    PwBfSmem s;
    int ksum = 0;
    for (int i = 0; i < nsrc; ++i) ksum += cps[i];
    s.mbk = (ksum + 127) / 128; s.np16 = (NPall + 15) & ~15; s.nkb = (s.np16 + 63) / 64;
    s.cols_dw = s.mbk * s.np16;
    const int cols = s.cols_dw + 2 * s.mbk * R;
    s.tmem_cols = 32; while (s.tmem_cols < cols) s.tmem_cols *= 2;
    // one ring stage: [d out planes][out planes][sources][x1][existing gradients of accumulate sources], R rows each
    int off = 0;
    s.o_dout = off; off += nplanes * R * cpo * 2;
    s.o_out = off; off += nplanes * R * cpo * 2;
    for (int i = 0; i < kMaxSrc; ++i) { s.o_src[i] = off; if (i < nsrc) off += R * cps[i] * 2; }
    s.o_x1 = off; off += R * x1cp * 2;
    for (int i = 0; i < kMaxSrc; ++i) { s.o_ge[i] = off; if (i < nsrc && acc[i]) off += R * cps[i] * 2; }
    s.stage_bytes = (off + 127) & ~127;
    s.w_bytes = s.mbk * s.nkb * 128 * 128;
    s.dr_bytes = s.nkb * R * 128;
    s.xs_bytes = s.mbk * 2 * R * 128;
    off = 1024;                                         // mbarriers + TMEM slot
    s.maps = off; off += 2 * 256 * 4;                   // logical row / column maps of the weight-gradient flush
    s.mdesc = off; off += 96 * 8;                       // the MMA issuer's precomputed shared-memory descriptors
    off = (off + 1023) & ~1023;
    s.w = off; off += s.w_bytes;
    s.dr = off; off += 2 * s.dr_bytes;
    s.xs = off; off += 2 * s.xs_bytes;
    s.st = off;                                         // gradient rows staged for the bulk stores: DOUBLE buffered (the stores of tile i
    { int o = 0; for (int i = 0; i < kMaxSrc; ++i) { s.st_off[i] = o; if (i < nsrc) o += R * cps[i] * 2; } s.st_bytes = (o + 127) & ~127; off += 2 * s.st_bytes; }     // read buffer i & 1 while tile i + 1 fills the other)
    s.st2_bytes = (R * x1cp * 2 + 127) & ~127;
    s.st2 = off; off += 2 * s.st2_bytes;
    off = (off + 1023) & ~1023;
    s.ring = off;
    { int ring = nstage * s.stage_bytes; const int scratch = 128 * 65 * 4; if (ring < scratch) ring = scratch; off += ring; }   // doubles as the dW transposition tile
    s.total = off + 1024;                               // slack for the manual 1024-byte alignment of the base
    return s;
}

template <int R>
__global__ void __launch_bounds__(kBfThreads, 1) pw_bwd_fused_kernel(const PwBwdArgs a) {
    static_assert(R == 32 || R == 64, "row tile");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    if (threadIdx.x == 0) BF_TS(0);
    const PwDesc& d = pw_desc_to_smem(a.d, smem + 520);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    const int NP = d.NPall, gwp = d.cols.gwp, nplanes = d.cols.nplanes, nsrc = d.nsrc, cpo = a.cpo;
    int cps[kMaxSrc], accs[kMaxSrc], ksum = 0;
    for (int i = 0; i < kMaxSrc; ++i) { cps[i] = i < nsrc ? d.src[i].cp : 0; accs[i] = i < nsrc ? d.src[i].accumulate : 0; ksum += cps[i]; }
    const int x1cp = a.x1 ? a.x1cp : 0;
    const int S = a.nbuf;                               // ring depth
    const PwBfSmem L = pw_bf_smem(R, nsrc, cps, accs, NP, nplanes, cpo, x1cp, S);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);               // [kBfMaxStages]
    uint64_t* empty = full + kBfMaxStages;                            // [kBfMaxStages]
    uint64_t* stg_full = empty + kBfMaxStages;                        // [2]
    uint64_t* stg_empty = stg_full + 2;                               // [2]
    uint64_t* tm_full = stg_empty + 2;                                // [2]
    uint64_t* tm_empty = tm_full + 2;                                 // [2]
    uint64_t* all_done = tm_empty + 2;                                // [1]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + 512);
    int* s_rmap = reinterpret_cast<int*>(smem + L.maps);              // GEMM row kk    -> (layer << 24 | k), -1 = padding
    int* s_cmap = s_rmap + 256;                                        // GEMM column j  -> (layer << 24 | n), -1 = padding
    unsigned char* Ws = smem + L.w;
    unsigned char* ring = smem + L.ring;

    const int tps = (a.Rt + R - 1) / R, ntile = kT * tps;
    const int tile_lo = blockIdx.x * a.tiles_per_cta, tile_hi = min(ntile, tile_lo + a.tiles_per_cta);
    const int my_tiles = max(0, tile_hi - tile_lo);

    if (warp == 0) tmem_alloc(s_tmem, (uint32_t)L.tmem_cols);
    if (tid == 32) {
        for (int s = 0; s < kBfMaxStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kBfTransformWarps + kBfGroupWarps); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&stg_full[b], kBfTransformWarps); mbar_init(&stg_empty[b], 1);
            mbar_init(&tm_full[b], 1); mbar_init(&tm_empty[b], kBfGroupWarps);
        }
        mbar_init(all_done, 1);
        mbar_fence_init();
    }
    // weights: Wb [KP][NP] (row kk, column j) -> K-major swizzled A operand blocks [mb][kb] of [128 rows][128 bytes]; the
    // staging tiles start as zeros (K / N padding and the rows of partial tiles contribute nothing)
    for (int i = tid; i < (L.w_bytes + 2 * L.dr_bytes + 2 * L.xs_bytes) / 16; i += kBfThreads) reinterpret_cast<uint4*>(Ws)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // narrow layers (KP <= 64 / 32): the weight rows are replicated 2 / 4 times over the 128 M rows, so that EVERY TMEM lane
    // block holds the data gradient and all four warps of an epilogue group share the tile's rows (a warp can only read its own 32 lanes)
    const int kpad = ksum <= 32 ? 32 : (ksum <= 64 ? 64 : 128);
    const int copies = min(128 / kpad, R / 8);
    for (int i = tid; i < d.KP * (NP >> 3); i += kBfThreads) {
        const int kk = i / (NP >> 3), c = i - kk * (NP >> 3);
        if (kk >= 128) continue;
        const uint4 w = *reinterpret_cast<const uint4*>(d.wb + (size_t)kk * NP + c * 8);
        for (int q = 0; q < copies; ++q)
            if (kk < kpad) *reinterpret_cast<uint4*>(Ws + (size_t)(c >> 3) * 16384 + sw128_offset(kk + q * kpad, (c & 7) * 8, 128)) = w;
    }
    for (int i = tid; i < 512; i += kBfThreads) {
        int l, k, p, sl, nn, v = -1;
        if (i < 256) { if (i < d.KP && pw_row(d, i, l, k)) v = (l << 24) | k; }
        else if (i - 256 < NP && pw_col(d, i - 256, p, sl, l, nn)) v = (l << 24) | nn;
        s_rmap[i] = v;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    if (tid == 0) BF_TS(1);
    pdl_wait();                                         // everything above read only this launch's descriptor / prepared weights
    if (tid == 0) BF_TS(2);

    // tile cursor: (slice, first row, rows) and (ring stage, ring phase) advance incrementally -- no divisions in the loops
    struct Cursor { int t, r0, s, k; };
    const int t_first = tile_lo / tps;
    auto cursor0 = [&]() { Cursor c; c.t = t_first; c.r0 = (tile_lo - t_first * tps) * R; c.s = 0; c.k = 0; return c; };
    auto advance = [&](Cursor& c) { c.r0 += R; if (c.r0 >= a.Rt) { c.r0 = 0; ++c.t; } if (++c.s == S) { c.s = 0; c.k ^= 1; } };

    if (warp == 0) {
        // ================================================================ TMA producer
        // One copy stream per LANE (d out / out of each plane, every source, its partial gradient, x1): a single thread issuing the
        // tile's 7-11 small bulk copies one after the other needs ~2 us per tile, which was the kernel's tile period.
        const bf16* my_base = nullptr; int my_cp = 0, my_off = 0;
        {
            int id = 0;
            auto claim = [&](const bf16* base, int cp, int off) { if (id++ == lane) { my_base = base; my_cp = cp; my_off = off; } };
            for (int p = 0; p < nplanes; ++p) { claim(a.dout[p], cpo, L.o_dout + p * R * cpo * 2); claim(a.out[p], cpo, L.o_out + p * R * cpo * 2); }
            for (int i = 0; i < nsrc; ++i) { claim(d.src[i].data, cps[i], L.o_src[i]); if (accs[i]) claim(d.src[i].grad, cps[i], L.o_ge[i]); }
            if (x1cp) claim(a.x1, x1cp, L.o_x1);
        }
        int row_elems = 2 * nplanes * cpo + ksum + x1cp;
        for (int i = 0; i < nsrc; ++i) if (accs[i]) row_elems += cps[i];
        Cursor cur = cursor0();
        for (int it = 0; it < my_tiles; ++it, advance(cur)) {
            const int s = cur.s, k = cur.k, t = cur.t, r0 = cur.r0, rows = min(R, a.Rt - cur.r0);
            if (lane == 0 && it == 20) BF_TS(16);
            mbar_wait(&empty[s], (k & 1) ^ 1);
            if (lane == 0 && it == 20) BF_TS(17);
            if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)rows * row_elems * 2);
            const size_t row = (size_t)t * a.Rt + r0;
            if (my_base) bulk_g2s(ring + (size_t)s * L.stage_bytes + my_off, my_base + row * my_cp, rows * my_cp * 2, &full[s]);
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        if (lane == 0) {
            const uint32_t idesc_d = umma_idesc(128, R, 0, 0), idesc_w = umma_idesc(128, L.np16, 1, 1);
            const uint32_t wa = smem_u32(Ws);
            // Every operand address is one of a few fixed ones (resident weights, two staging buffers): the descriptors are built
            // ONCE into a private table, so a tile costs 2 shared loads + 1 tcgen05.mma per K step (a lone warp issues dependent
            // instructions ~5 cycles apart: building descriptors in the loop made the issue 0.9 us per tile, next to a 1.1 us transform).
            // Table: [0,16) weights (dgrad A), [16 + 16 b, +16) dR K-major (dgrad B), [48 + 8 b, +8) act(x) MN-major (wgrad A),
            // [64 + 8 b, +8) dR MN-major (wgrad B).  The launcher guarantees one 128-row weight block (mbk == 1).
            uint64_t* md = reinterpret_cast<uint64_t*>(smem + L.mdesc);
            const int nks = L.np16 >> 4;
            for (int ks = 0; ks < nks; ++ks) md[ks] = umma_desc(wa + (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u, 16, 1024);
            for (int b = 0; b < 2; ++b) {
                const uint32_t ra = smem_u32(smem + L.dr + (size_t)b * L.dr_bytes), xa = smem_u32(smem + L.xs + (size_t)b * L.xs_bytes);
                for (int ks = 0; ks < nks; ++ks) md[16 + 16 * b + ks] = umma_desc(ra + (uint32_t)(ks >> 2) * (uint32_t)(R * 128) + (uint32_t)(ks & 3) * 32u, 16, 1024);
                for (int ks = 0; ks < R / 16; ++ks) {
                    md[48 + 8 * b + ks] = umma_desc(xa + (uint32_t)ks * 2048u, R * 128, 1024);
                    md[64 + 8 * b + ks] = umma_desc(ra + (uint32_t)ks * 2048u, R * 128, 1024);
                }
            }
            for (int it = 0; it < my_tiles; ++it) {
                const int b = it & 1, n = it >> 1;
                if (it == 20) BF_TS(18);
                mbar_wait(&stg_full[b], n & 1);
                if (it == 0) BF_TS(4); if (it == 8) BF_TS(10);
                if (it == 20) BF_TS(19);
                mbar_wait(&tm_empty[b], (n & 1) ^ 1);
                if (it == 20) BF_TS(20);
                tc_fence_after();
                const uint64_t* dB = md + 16 + 16 * b;
                const uint32_t tacc = tmem + (uint32_t)(L.cols_dw + b * R);
                for (int ks = 0; ks < nks; ++ks) umma_bf16(tacc, md[ks], dB[ks], idesc_d, ks > 0);
                umma_commit(&tm_full[b]);
#pragma unroll
                for (int ks = 0; ks < R / 16; ++ks) umma_bf16(tmem, md[48 + 8 * b + ks], md[64 + 8 * b + ks], idesc_w, it > 0 || ks > 0);
                umma_commit(&stg_empty[b]);
                if (it == 20) BF_TS(21);
            }
            umma_commit(all_done);
            mbar_wait(all_done, 0);
        }
    } else if (warp < 2 + kBfTransformWarps) {
        // ================================================================ transform warps
        const int ttid = tid - 64;
        const int nqr = NP >> 3, rq = ttid % nqr, rrl = ttid / nqr, rnrl = kBfTransformThreads / nqr;       // dR chunks
        const int nqx = ksum >> 3, xq = ttid % nqx, xrl = ttid / nqx, xnrl = kBfTransformThreads / nqx;    // act(src) chunks
        const int tp = (rq * 8) / gwp, tc = rq * 8 - tp * gwp;
        int xsrc = 0, xch = xq;
        while (xsrc < nsrc - 1 && xch >= (cps[xsrc] >> 3)) { xch -= cps[xsrc] >> 3; ++xsrc; }
        int xcp = 0, o_xsrc = 0;
#pragma unroll
        for (int i = 0; i < kMaxSrc; ++i) if (i == xsrc) { xcp = cps[i]; o_xsrc = L.o_src[i]; }
        const float2* xaff = d.src[xsrc].aff;
        const int o_dout = L.o_dout, o_out = L.o_out, stage_bytes = L.stage_bytes, o_dr = L.dr, o_xs = L.xs, dr_bytes = L.dr_bytes, xs_bytes = L.xs_bytes;
        const bool sclamp = d.src[xsrc].clamp != 0, oclamp = a.out_clamp != 0;
        const double inv_n = 1.0 / (double)a.Rt;
        float4 c8[8];
        float2 x8[8];
        int cur_t = -1;
        Cursor cur = cursor0();
        for (int it = 0; it < my_tiles; ++it, advance(cur)) {
            const int s = cur.s, k = cur.k, b = it & 1, n = it >> 1, t = cur.t, rows = min(R, a.Rt - cur.r0);
            if (t != cur_t) {                           // this thread's chunk constants for the slice (no shared table, no barrier)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    int p, sl, l, nn;
                    c8[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rrl < rnrl && pw_col(d, rq * 8 + q, p, sl, l, nn)) {
                        const size_t idx = (size_t)t * cpo + sl;
                        c8[q] = bnbwd_consts(a.tb[p].aff[idx], a.tb[p].bnp[idx], ld_sum(a.tb[p].bsum + idx), inv_n);
                    }
                    x8[q] = (xrl < xnrl && xaff) ? xaff[(size_t)t * xcp + xch * 8 + q] : make_float2(1.f, 0.f);
                }
                cur_t = t;
            }
            if (ttid == 0 && it == 20) BF_TS(22);
            mbar_wait(&full[s], k & 1);
            if (ttid == 0 && it == 0) BF_TS(3);
            if (ttid == 0 && it == 20) BF_TS(23);
            mbar_wait(&stg_empty[b], (n & 1) ^ 1);
            if (ttid == 0 && it == 20) BF_TS(24);
            const unsigned char* rb = ring + (size_t)s * stage_bytes;
            unsigned char* Dr = smem + o_dr + (size_t)b * dr_bytes;
            unsigned char* Xs = smem + o_xs + (size_t)b * xs_bytes;
            if (rrl < rnrl) {
                const uint4* dv = reinterpret_cast<const uint4*>(rb + o_dout + (size_t)tp * R * cpo * 2) + (tc >> 3);
                const uint4* ov = reinterpret_cast<const uint4*>(rb + o_out + (size_t)tp * R * cpo * 2) + (tc >> 3);
                const int nch = cpo >> 3;
                for (int rbase = rrl; rbase < R; rbase += 3 * rnrl) {
                    uint4 dvv[3], ovv[3];
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int r = rbase + u * rnrl;
                        dvv[u] = make_uint4(0, 0, 0, 0); ovv[u] = dvv[u];
                        if (r < rows) { dvv[u] = dv[r * nch]; ovv[u] = ov[r * nch]; }
                    }
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int r = rbase + u * rnrl;
                        if (r < R) {
                            uint32_t* dw = reinterpret_cast<uint32_t*>(&dvv[u]); const uint32_t* ow = reinterpret_cast<const uint32_t*>(&ovv[u]);
                            if (r < rows) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float2 dd = unpack2(dw[i]), oo = unpack2(ow[i]);
                                    dw[i] = pack2(bnbwd_apply(dd.x, oo.x, c8[2 * i], oclamp), bnbwd_apply(dd.y, oo.y, c8[2 * i + 1], oclamp));
                                }
                            }
                            *reinterpret_cast<uint4*>(Dr + sw128_offset(r, rq * 8, R)) = dvv[u];
                        }
                    }
                }
            }
            if (xrl < xnrl) {
                const uint4* sv = reinterpret_cast<const uint4*>(rb + o_xsrc) + xch;
                const int nch = xcp >> 3;
                for (int rbase = xrl; rbase < R; rbase += 3 * xnrl) {
                    uint4 v[3];
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int r = rbase + u * xnrl;
                        v[u] = make_uint4(0, 0, 0, 0);
                        if (r < rows) v[u] = sv[r * nch];
                    }
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int r = rbase + u * xnrl;
                        if (r < R) *reinterpret_cast<uint4*>(Xs + sw128_offset(r, xq * 8, R)) = r < rows ? affine8(v[u], x8, sclamp) : make_uint4(0, 0, 0, 0);
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&stg_full[b]); mbar_arrive(&empty[s]); }
            if (ttid == 0 && it == 20) BF_TS(25);
            if (ttid == 0 && it == 21) BF_TS(26);
        }
    } else {
        // ================================================================ epilogue warps
        // Two GROUPS of four warps alternate tiles (group g owns tiles it = g, g + 2, ...: TMEM accumulator g, staging buffers g), so
        // the per-tile latency chain of one group (barrier -> tcgen05.ld -> shared-memory gather -> bulk store) overlaps the other's.
        const int ew = warp - (2 + kBfTransformWarps), lg = ew & 3, grp = ew >> 2, etid = tid - (2 + kBfTransformWarps) * 32, gtid = etid & (kBfGroupThreads - 1);
        // this thread's data-gradient channel: TMEM lane 32 * lg + lane = copy q of channel kk; it walks rows
        // [q * hrt, + hrt) of its group's tile.  Everything the tile loop needs is hoisted into registers here.
        const int lane_id = 32 * lg + lane, q_copy = lane_id / kpad, kk = lane_id - q_copy * kpad;
        const int hrt = R / copies, row_first = q_copy < copies ? q_copy * hrt : 0;
        struct Chan {
            int cp, o_src, o_ge, st_off, slot; bool on, clamp, want, acc;
            const float2* aff; const float2* bnp; double2* bsum;
            float s1, s2; float4 sc;
        } ch;
        ch.on = false; ch.cp = 0; ch.o_src = ch.o_ge = ch.st_off = ch.slot = 0; ch.clamp = ch.want = ch.acc = false;
        ch.aff = ch.bnp = nullptr; ch.bsum = nullptr; ch.s1 = ch.s2 = 0.f; ch.sc = make_float4(1.f, 0.f, 0.f, 0.f);
        if (q_copy < copies) {
            int off = 0;
            for (int i = 0; i < nsrc; ++i) {
                const int cp = d.src[i].cp;
                if (kk >= off && kk < off + cp) {
                    const PwSrc& Sx = d.src[i];
                    ch.on = true; ch.cp = cp; ch.slot = kk - off; ch.clamp = Sx.clamp != 0; ch.acc = Sx.accumulate != 0;
                    ch.aff = Sx.aff; ch.bnp = Sx.bnp; ch.bsum = Sx.bsum;
                    ch.want = Sx.bsum != nullptr && ch.slot >= Sx.sum_lo && ch.slot < Sx.sum_hi;
                    ch.o_src = L.o_src[i]; ch.o_ge = L.o_ge[i]; ch.st_off = L.st_off[i];
                }
                off += cp;
            }
        }
        const int cols_dw = L.cols_dw, o_dout = L.o_dout, o_x1 = L.o_x1, stage_bytes = L.stage_bytes;
        const int o_st = L.st + grp * L.st_bytes, o_st2 = L.st2 + grp * L.st2_bytes;       // this group's staging buffers
        // pass-through role: thread <-> x1 slots gtid and gtid + 128
        int x_src[2] = {-1, -1};                        // element offset inside the d out region of a stage, -2: padding (zero), -1: no role
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int slot = gtid + h * kBfGroupThreads;
            if (x1cp && slot < x1cp) {
                const int l = slot_logical(a.x1map, slot);
                x_src[h] = (l >= 0 && (l >> 1) < a.ncopy) ? (l & 1) * R * cpo + a.copy_dst0 + (l >> 1) : -2;
            }
        }
        float xs1[2] = {0.f, 0.f}, xs2[2] = {0.f, 0.f};
        float4 xc[2] = {make_float4(1.f, 0.f, 0.f, 0.f), make_float4(1.f, 0.f, 0.f, 0.f)};
        const bool xclamp = a.x1clamp != 0;
        // bulk-store plan of thread 0 (one store per gradient tensor)
        bf16* g_ptr[kMaxSrc]; int g_cp[kMaxSrc], g_off[kMaxSrc];
#pragma unroll
        for (int i = 0; i < kMaxSrc; ++i) { g_ptr[i] = i < nsrc ? d.src[i].grad : nullptr; g_cp[i] = cps[i]; g_off[i] = L.st_off[i]; }
        auto flush = [&](int t) {
            if (ch.want && (ch.s1 != 0.f || ch.s2 != 0.f)) {
                double2* dst = ch.bsum + (size_t)t * ch.cp + ch.slot;
                atomicAdd(&dst->x, (double)ch.s1); atomicAdd(&dst->y, (double)ch.s2);
            }
            ch.s1 = ch.s2 = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (x_src[h] != -1 && a.x1bsum && (xs1[h] != 0.f || xs2[h] != 0.f)) {
                    double2* dst = a.x1bsum + (size_t)t * x1cp + gtid + h * kBfGroupThreads;
                    atomicAdd(&dst->x, (double)xs1[h]); atomicAdd(&dst->y, (double)xs2[h]);
                }
                xs1[h] = xs2[h] = 0.f;
            }
        };
        auto run_chan = [&](int b, int rows, const unsigned char* rb) {
            const uint32_t taddr = tmem + ((uint32_t)(32 * lg) << 16) + (uint32_t)(cols_dw + b * R + row_first);
            const int cp = ch.cp;
            float s1 = ch.s1, s2 = ch.s2;
            const float4 sc = ch.sc;
            for (int c8 = 0; c8 < hrt; c8 += 8) {       // 8 rows at a time: one tcgen05.ld, the shared-memory loads up front, then the math
                float v[8];
                tmem_ld8(taddr + c8, v);
                if (!ch.on) continue;
                const int row0 = row_first + c8;
                const unsigned short* rawp = reinterpret_cast<const unsigned short*>(rb + ch.o_src) + ch.slot + row0 * cp;
                const unsigned short* gep = reinterpret_cast<const unsigned short*>(rb + ch.o_ge) + ch.slot + row0 * cp;
                unsigned short* stp = reinterpret_cast<unsigned short*>(smem + o_st + ch.st_off) + ch.slot + row0 * cp;
                const int nr = rows - row0;             // rows of this chunk inside the slice (rows past it are staged but never stored)
                uint32_t rawv[8], gev[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { rawv[j] = rawp[j * cp]; gev[j] = ch.acc ? (uint32_t)gep[j * cp] : 0u; }
                uint32_t gbv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float g = v[j];
                    if (ch.acc) g = __bfloat162float(__float2bfloat16_rn(g)) + __uint_as_float(gev[j] << 16);
                    gbv[j] = __bfloat16_as_ushort(__float2bfloat16_rn(g));
                    stp[j * cp] = (unsigned short)gbv[j];
                }
                if (ch.want) {
                    if (nr >= 8) {                      // every tile but a slice's last: no per-row guards (warp-uniform: one copy index per warp)
#pragma unroll
                        for (int j = 0; j < 8; ++j) sum_accum(__uint_as_float(gbv[j] << 16), __uint_as_float(rawv[j] << 16), sc, ch.clamp, s1, s2);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < nr) sum_accum(__uint_as_float(gbv[j] << 16), __uint_as_float(rawv[j] << 16), sc, ch.clamp, s1, s2);
                    }
                }
            }
            ch.s1 = s1; ch.s2 = s2;
        };
        int cur_t = -1;
        Cursor cur = cursor0();
        if (grp) advance(cur);
        const int bar_a = 1 + 2 * grp, bar_b = 2 + 2 * grp;
        for (int it = grp; it < my_tiles; it += 2, advance(cur), advance(cur)) {
            const int s = cur.s, k = cur.k, b = grp, n = it >> 1, t = cur.t, r0 = cur.r0, rows = min(R, a.Rt - cur.r0);
            if (t != cur_t) {
                if (cur_t >= 0) flush(cur_t);
                if (ch.on) ch.sc = sum_consts(ch.aff, ch.bnp, (size_t)t * ch.cp + ch.slot);
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (x_src[h] != -1) xc[h] = sum_consts(a.x1aff, a.x1bnp, (size_t)t * x1cp + gtid + h * kBfGroupThreads);
                cur_t = t;
            }
            if (etid == 0 && it == 20) BF_TS(27);
            if (gtid == 0) bulk_store_wait_read();         // this group's previous stores (same staging buffers) have finished READING them
            named_bar_sync(bar_a, kBfGroupThreads);
            if (etid == 0 && it == 20) BF_TS(28);
            mbar_wait(&full[s], k & 1);                 // (long complete: acquires the TMA writes for this thread)
            mbar_wait(&tm_full[b], n & 1);
            if (etid == 0 && it == 0) BF_TS(5);
            if (etid == 0 && it == 20) BF_TS(29);
            tc_fence_after();
            const unsigned char* rb = ring + (size_t)s * stage_bytes;
            run_chan(b, rows, rb);
            if (etid == 0 && it == 0) BF_TS(6);
            if (etid == 0 && it == 20) BF_TS(30);
            tc_fence_before();
            // pass-through half of a stride-1 unit: d x1[slot(2i + p)] = d out_p[copy_dst0 + i]   (bit-exact gather)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (x_src[h] == -1) continue;
                const int slot = gtid + h * kBfGroupThreads, xo = x_src[h];
                const unsigned short* dreg = reinterpret_cast<const unsigned short*>(rb + o_dout);
                const unsigned short* xr = reinterpret_cast<const unsigned short*>(rb + o_x1) + slot;
                unsigned short* dst = reinterpret_cast<unsigned short*>(smem + o_st2) + slot;
                float a1 = xs1[h], a2 = xs2[h];
#pragma unroll 4
                for (int r = 0; r < rows; ++r) {
                    const unsigned short gb = xo >= 0 ? dreg[xo + r * cpo] : (unsigned short)0;
                    dst[r * x1cp] = gb;
                    sum_accum(__uint_as_float((uint32_t)gb << 16), __uint_as_float((uint32_t)xr[r * x1cp] << 16), xc[h], xclamp, a1, a2);
                }
                xs1[h] = a1; xs2[h] = a2;
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(&tm_empty[b]); mbar_arrive(&empty[s]); }
            fence_proxy_async();
            named_bar_sync(bar_b, kBfGroupThreads);
            if (gtid == 0) {                            // one TMA bulk store per gradient tensor (a row tile is contiguous in HBM)
                const size_t row = (size_t)t * a.Rt + r0;
#pragma unroll
                for (int i = 0; i < kMaxSrc; ++i)
                    if (g_ptr[i]) bulk_s2g_nc(g_ptr[i] + row * g_cp[i], smem + o_st + g_off[i], (uint32_t)rows * g_cp[i] * 2);
                if (x1cp) bulk_s2g_nc(a.dx1 + row * x1cp, smem + o_st2, (uint32_t)rows * x1cp * 2);
                bulk_commit();                          // ONE bulk group per tile
            }
            if (etid == 0 && it == 20) BF_TS(31);
            if (etid == 0 && it == 22) BF_TS(12);
        }
        if (etid == 0) BF_TS(7);
        if (cur_t >= 0) flush(cur_t);
        if (gtid == 0) bulk_store_wait_all();
    }

    // ================================================================ all roles: weight-gradient epilogue
    tc_fence_before();
    __syncthreads();
    if (tid == 0) BF_TS(8);
    tc_fence_after();
    if (my_tiles > 0) {
        float* Sc = reinterpret_cast<float*>(ring);        // [128][65]
        for (int mb = 0; mb < L.mbk; ++mb)
            for (int c0 = 0; c0 < L.np16; c0 += 64) {
                const int ncol = min(64, L.np16 - c0);
                if (warp < 4) {
                    for (int c = 0; c < ncol; c += 8) {
                        float v[8];
                        tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(mb * L.np16 + c0 + c), v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) Sc[(32 * warp + lane) * 65 + c + i] = v[i];
                    }
                }
                __syncthreads();
                {
                    const int col = tid & 63, cm = (col < ncol) ? s_cmap[c0 + col] : -1;
                    if (cm >= 0) {
                        float* wl0 = d.layer[0].dw; float* wl1 = d.layer[1].dw;
                        const int N0 = d.layer[0].N, N1 = d.layer[1].N;
                        for (int rowk = tid >> 6; rowk < 128; rowk += kBfThreads >> 6) {
                            const int rm = s_rmap[mb * 128 + rowk];
                            if (rm >= 0 && (rm >> 24) == (cm >> 24)) {
                                const float val = Sc[rowk * 65 + col];
                                if (val != 0.f) {
                                    if ((rm >> 24) == 0) atomicAdd(wl0 + (size_t)(rm & 0xffffff) * N0 + (cm & 0xffffff), val);
                                    else atomicAdd(wl1 + (size_t)(rm & 0xffffff) * N1 + (cm & 0xffffff), val);
                                }
                            }
                        }
                    }
                }
                __syncthreads();
            }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) BF_TS(9);
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)L.tmem_cols);
    // ---- BatchNorm parameter gradients (one CTA): dgamma = sum_t S2, dbeta = sum_t S1
    if (blockIdx.x == 0) {
        for (int j = tid; j < NP; j += kBfThreads) {
            int p, sl, l, nn;
            if (!pw_col(d, j, p, sl, l, nn)) continue;
            double gs = 0.0, bs = 0.0;
            double2 v[kT];                               // all four loads in flight before the sums (block 0 ends the kernel)
#pragma unroll
            for (int t = 0; t < kT; ++t) v[t] = ld_sum(a.tb[p].bsum + (size_t)t * cpo + sl);
#pragma unroll
            for (int t = 0; t < kT; ++t) { bs += v[t].x; gs += v[t].y; }
            d.layer[l].dg[nn] = (float)gs; d.layer[l].dbe[nn] = (float)bs;
        }
        __syncthreads();
        if (tid == 0) BF_TS(11);
    }
}

}  // namespace v2
}  // namespace cdra
#endif
