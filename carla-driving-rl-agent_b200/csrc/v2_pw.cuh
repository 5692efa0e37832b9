// v2 tower: pointwise (1x1) convolutions as bf16 tensor-core GEMMs over contiguous row tiles
// (core/architectures.py:130,134,140,170).  One kernel covers
//   pw1 / head conv      : 1-2 sources (a plain tensor, or both planes of the previous unit) -> one plain tensor
//   stride-1 unit tail   : pw2 GEMM + channel shuffle + pass-through half (:142-144) -> both output planes
//   stride-2 unit tail   : pw2 and shortcut-pw as ONE block-structured GEMM over [r2 | rs] -> both output planes
// Data path per row tile:  TMA bulk copy of the raw source rows (contiguous bytes) -> shared memory;
// BatchNorm affine (+ReLU6) of the producers applied by a vectorised pass into a padded MMA tile;
// mma.sync m16n8k16 against the layer's bf16 weights (resident in shared memory for the whole CTA);
// epilogue stages the bf16 result rows (with the shuffled pass-through slots merged in), then one pass stores
// 16-byte vectors to HBM and accumulates the per-(slice, channel) BatchNorm sums of the stored values.
// The last CTA turns the sums into the consumer-side tables (scale/shift, mean/inv_std, moving statistics).
#pragma once
#ifndef CDRA_EMU
#include "v2_common.cuh"

namespace cdra {
namespace v2 {

constexpr int kMaxSrc = 3;

struct PwSrc {
    const bf16* data;       // [4*Rt][cp]
    bf16* grad;             // same shape (backward)
    const float2* aff;      // [4][cp] (scale, shift); pad slots hold (0, 0)
    const float2* bnp;      // [4][cp] (mean, inv_std)
    double2* bsum;          // [4][cp] backward sums of this tensor (written by the kernel that finalises its gradient)
    int cp, clamp;
    SlotMap map;
    int kbase, layer;       // logical input channel of logical slot 0; which of the (<=2) layers this source feeds
    int sum_lo, sum_hi;     // slots whose backward sums are needed downstream
    int accumulate;         // backward: add to the existing gradient (second consumer of the tensor)
};
struct PwCols {
    int nplanes, gwp;       // GEMM columns per output plane (multiple of 8); NPall = nplanes * gwp
    int seg0p, seg0n;       // columns [0, seg0p) of a plane group: layer 0 outputs (seg0n valid)
    int seg1n;              // columns [seg0p, seg0p + seg1n): layer 1 outputs
    int interleave;         // logical output = interleave ? 2*c + plane : c
};
struct PwDesc {
    PwSrc src[kMaxSrc]; int nsrc;
    LayerP layer[2];
    PwCols cols;
    int KP, NPall;          // KP: multiple of 16 >= sum(cp);  NPall = nplanes * gwp
    bf16* wf;               // [NPall][KP]   wf[j][k]  (forward B operand, k contiguous)
    bf16* wb;               // [KP][NPall]   wb[k][j]  (data-gradient B operand, j contiguous)
    float* bias;            // [NPall]
    // the same operands as 64-column blocks of 128-byte-swizzled rows (chunk c of row r at c ^ (r & 7)): a contiguous run
    // of rows of one block is a ready-made K-major tcgen05 operand tile, fetched by ONE bulk copy (v4_pwg.cuh)
    bf16* wfs;              // [ceil(KP/64)][NPall][64]    block kb, row j  : wf[j][64 kb ..]
    bf16* wbs;              // [ceil(NPall/64)][KP][64]    block jb, row kk : wb[kk][64 jb ..]
};

// GEMM row kk -> (layer, logical k) ; returns false for padding
CDRA_DEV bool pw_row(const PwDesc& d, int kk, int& layer, int& k) {
    int off = 0;
    for (int i = 0; i < d.nsrc; ++i) {
        if (kk < off + d.src[i].cp) {
            const int l = slot_logical(d.src[i].map, kk - off);
            if (l < 0) return false;
            layer = d.src[i].layer; k = d.src[i].kbase + l;
            return true;
        }
        off += d.src[i].cp;
    }
    return false;
}
// GEMM column j -> (plane, slot, layer, logical n) ; returns false for padding
CDRA_DEV bool pw_col(const PwDesc& d, int j, int& plane, int& slot, int& layer, int& n) {
    plane = j / d.cols.gwp; slot = j - plane * d.cols.gwp;
    int c = slot;
    if (c < d.cols.seg0p) {
        if (c >= d.cols.seg0n) return false;
        layer = 0;
    } else {
        c -= d.cols.seg0p;
        if (c >= d.cols.seg1n) return false;
        layer = 1;
    }
    n = d.cols.interleave ? 2 * c + plane : c;
    return true;
}

// The launch's descriptor copied once into (dynamic) shared memory: every role reads its fields many times, and each
// dependent read of the global copy costs an L2 round trip in the kernels' prologues.  Ends with a CTA barrier.
static_assert(sizeof(PwDesc) <= 504 && sizeof(PwDesc) % 4 == 0, "PwDesc must fit the 504 bytes the kernels reserve at offset 520 of their header");
CDRA_DEV const PwDesc& pw_desc_to_smem(const PwDesc* g, void* s) {
    for (int i = threadIdx.x; i < (int)(sizeof(PwDesc) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(s)[i] = reinterpret_cast<const uint32_t*>(g)[i];
    __syncthreads();
    return *reinterpret_cast<const PwDesc*>(s);
}

// ---- bf16 operand matrices in slot order, rebuilt at the start of every forward (weights change every SGD step)
constexpr int kPrepMax = 36;
struct PrepArgs { PwDesc d[kPrepMax]; int n; };

__global__ void __launch_bounds__(256) pw_prep_kernel(const PwDesc* __restrict__ descs) {
    const PwDesc& d = descs[blockIdx.x];
    const int total = d.NPall * d.KP;
    for (int i = blockIdx.y * 256 + threadIdx.x; i < total; i += gridDim.y * 256) {
        const int j = i / d.KP, kk = i - j * d.KP;
        int lk, k, p, s, lj, n;
        float v = 0.f;
        if (pw_row(d, kk, lk, k) && pw_col(d, j, p, s, lj, n) && lk == lj) v = d.layer[lk].w[(size_t)k * d.layer[lk].N + n];
        const bf16 h = __float2bfloat16_rn(v);
        d.wf[(size_t)j * d.KP + kk] = h;
        d.wb[(size_t)kk * d.NPall + j] = h;
    }
    if (blockIdx.y == 0)
        for (int j = threadIdx.x; j < d.NPall; j += 256) {
            int p, s, l, n;
            d.bias[j] = pw_col(d, j, p, s, l, n) ? d.layer[l].b[n] : 0.f;
        }
    // swizzled blocks (zero beyond KP / NPall inside the last block)
    const int nkb = (d.KP + 63) >> 6, njb = (d.NPall + 63) >> 6;
    for (int i = blockIdx.y * 256 + threadIdx.x; i < nkb * d.NPall * 64; i += gridDim.y * 256) {
        const int kb = i / (d.NPall * 64), rem = i - kb * d.NPall * 64, j = rem >> 6, c = rem & 63, kk = kb * 64 + c;
        int lk, k, p, s, lj, n;
        float v = 0.f;
        if (kk < d.KP && pw_row(d, kk, lk, k) && pw_col(d, j, p, s, lj, n) && lk == lj) v = d.layer[lk].w[(size_t)k * d.layer[lk].N + n];
        d.wfs[((size_t)kb * d.NPall + j) * 64 + ((((c >> 3) ^ (j & 7)) << 3) | (c & 7))] = __float2bfloat16_rn(v);
    }
    for (int i = blockIdx.y * 256 + threadIdx.x; i < njb * d.KP * 64; i += gridDim.y * 256) {
        const int jb = i / (d.KP * 64), rem = i - jb * d.KP * 64, kk = rem >> 6, c = rem & 63, j = jb * 64 + c;
        int lk, k, p, s, lj, n;
        float v = 0.f;
        if (j < d.NPall && pw_row(d, kk, lk, k) && pw_col(d, j, p, s, lj, n) && lk == lj) v = d.layer[lk].w[(size_t)k * d.layer[lk].N + n];
        d.wbs[((size_t)jb * d.KP + kk) * 64 + ((((c >> 3) ^ (kk & 7)) << 3) | (c & 7))] = __float2bfloat16_rn(v);
    }
}

// one M block (<= 128 GEMM columns / rows) of the tcgen05 GEMM family of v4_pwg.cuh
constexpr int kGMaxBlk = 8;
struct GBlock { int base, n, aux0, aux1; };   // forward: first GEMM column, columns, plane, first slot; data gradient: first GEMM row, rows, source, first slot

// ======================================================================================== forward
struct PwFwdArgs {
    const PwDesc* d;            // device copy of the descriptor (workspace)
    int Rt;                     // rows per slice
    bf16* out[2]; int cpo;      // output plane(s), row stride cpo
    Tables tb[2];
    int gwv;                    // columns [0, gwv) of every plane group are stored
    // pass-through half of a stride-1 unit (nullptr otherwise): dst slot n0p_out + i <- x1 logical channel 2i + plane
    const bf16* x1; int x1cp; SlotMap x1map; const float2* x1aff; const float2* x1bnp;
    int ncopy, copy_dst0;
    int training;
    unsigned* counter;
    int colmode;                // 0: every CTA covers all GEMM columns; 1: blockIdx.y = (plane, N tile)
    int ntiles_n;               // colmode 1: N tiles per plane
    int tiles_per_cta;
    int nbuf;                   // tcgen05 kernel: depth of the TMA ring
    GBlock blk[kGMaxBlk]; int nblk;   // v4_pwg.cuh: blockIdx.y = M block
    int timeline;                     // record the role timeline of block (0, 0) (cdra_debug_timeline)
};

template <int R, int WM, int WN, int MT, int NBW>
struct PwFwdCfg {
    static constexpr int NT = WN * NBW * 8;
    static_assert(WM * WN == 8 && WM * MT * 16 == R, "warp layout");
};

// dynamic shared memory carve-up (bytes), shared by host and device
struct PwFwdSmem {
    int aff, bias, stat, w, raw, a, st, total;
    int raw_stride;             // bytes per raw buffer
    int lda, ldw, lds;          // element strides
};
inline __host__ __device__ PwFwdSmem pw_fwd_smem(int R, int NT, int KP, int src_row_bytes, int splanes, int swidth) {
    PwFwdSmem s;
    s.lda = pad_ld(KP); s.ldw = pad_ld(KP); s.lds = pad_ld(swidth);
    int off = 64;                                     // mbarriers
    s.aff = off; off += KP * 8;
    s.bias = off; off += NT * 4;
    s.stat = off; off += splanes * swidth * 8;        // float sum, sq per covered slot
    off = (off + 127) & ~127;
    s.w = off; off += NT * s.ldw * 2;
    off = (off + 127) & ~127;
    s.raw_stride = (R * src_row_bytes + 127) & ~127;
    s.raw = off; off += 2 * s.raw_stride;
    s.a = off; off += R * s.lda * 2;
    off = (off + 127) & ~127;
    { const int st_bytes = splanes * R * s.lds * 2; s.st = off; off += st_bytes > 16384 ? st_bytes : 16384; }   // doubles as the 16 KB flush scratch
    s.total = off;
    return s;
}

template <int R, int WM, int WN, int MT, int NBW>
__global__ void __launch_bounds__(256, 2) pw_fwd_kernel(const PwFwdArgs a) {
    constexpr int NT = WN * NBW * 8;
    extern __shared__ __align__(128) unsigned char smem[];
    const PwDesc& d = *a.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    const int KP = d.KP, gwp = d.cols.gwp;
    // ---- this CTA's column tile
    int jt0, ncols, p0, splanes, scol0, swidth;
    if (a.colmode == 0) { jt0 = 0; ncols = d.NPall; p0 = 0; splanes = d.cols.nplanes; scol0 = 0; swidth = a.cpo; }
    else {
        p0 = blockIdx.y / a.ntiles_n; const int ti = blockIdx.y - p0 * a.ntiles_n;
        scol0 = ti * NT; swidth = min(NT, gwp - scol0); jt0 = p0 * gwp + scol0; ncols = swidth; splanes = 1;
    }
    int src_row_bytes = 0;
    for (int i = 0; i < d.nsrc; ++i) src_row_bytes += d.src[i].cp * 2;
    const int x1_off_rows = src_row_bytes;            // x1 tile sits after the sources in every raw buffer
    if (a.x1) src_row_bytes += a.x1cp * 2;
    const PwFwdSmem L = pw_fwd_smem(R, NT, KP, src_row_bytes, a.colmode == 0 ? d.cols.nplanes : 1, a.colmode == 0 ? a.cpo : NT);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    float2* s_aff = reinterpret_cast<float2*>(smem + L.aff);
    float* s_bias = reinterpret_cast<float*>(smem + L.bias);
    float* s_stat = reinterpret_cast<float*>(smem + L.stat);
    bf16* Ws = reinterpret_cast<bf16*>(smem + L.w);
    unsigned char* raw = smem + L.raw;
    bf16* As = reinterpret_cast<bf16*>(smem + L.a);
    bf16* St = reinterpret_cast<bf16*>(smem + L.st);
    const int lda = L.lda, ldw = L.ldw, lds = L.lds;

    // ---- tile schedule: contiguous range of row tiles
    const int tps = (a.Rt + R - 1) / R, ntile = kT * tps;
    const int tile_lo = blockIdx.x * a.tiles_per_cta, tile_hi = min(ntile, tile_lo + a.tiles_per_cta);

    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
    // weights + bias of this column tile, zero the staging tile (pad slots stay zero forever)
    for (int i = tid; i < NT * (KP / 8); i += 256) {
        const int j = i / (KP / 8), c = i - j * (KP / 8);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (j < ncols) v = *reinterpret_cast<const uint4*>(d.wf + (size_t)(jt0 + j) * KP + c * 8);
        *reinterpret_cast<uint4*>(Ws + (size_t)j * ldw + c * 8) = v;
    }
    for (int j = tid; j < NT; j += 256) s_bias[j] = j < ncols ? d.bias[jt0 + j] : 0.f;
    for (int i = tid; i < splanes * R * lds / 2; i += 256) reinterpret_cast<uint32_t*>(St)[i] = 0u;
    for (int i = tid; i < R * lda / 2; i += 256) reinterpret_cast<uint32_t*>(As)[i] = 0u;      // K padding columns stay zero
    for (int i = tid; i < splanes * swidth * 2; i += 256) s_stat[i] = 0.f;
    __syncthreads();

    auto issue = [&](int tile, int buf) {            // thread 0 only
        const int t = tile / tps, r0 = (tile - t * tps) * R, rows = min(R, a.Rt - r0);
        unsigned char* dst = raw + (size_t)buf * L.raw_stride;
        uint32_t bytes = 0;
        for (int i = 0; i < d.nsrc; ++i) bytes += rows * d.src[i].cp * 2;
        if (a.x1) bytes += rows * a.x1cp * 2;
        mbar_expect_tx(&full[buf], bytes);
        int off = 0;
        for (int i = 0; i < d.nsrc; ++i) {
            bulk_g2s(dst + (size_t)R * off, d.src[i].data + ((size_t)t * a.Rt + r0) * d.src[i].cp, rows * d.src[i].cp * 2, &full[buf]);
            off += d.src[i].cp * 2;
        }
        if (a.x1) bulk_g2s(dst + (size_t)R * x1_off_rows, a.x1 + ((size_t)t * a.Rt + r0) * a.x1cp, rows * a.x1cp * 2, &full[buf]);
    };
    pdl_wait();                                       // everything above touched only this launch's descriptor / prepared weights
    if (tid == 0) {
        if (tile_lo < tile_hi) issue(tile_lo, 0);
        if (tile_lo + 1 < tile_hi) issue(tile_lo + 1, 1);
    }

    // ---- per-thread constant roles (everything that needs a division is computed once, outside the tile loop)
    const int wm = warp % WM, wn = warp / WM;
    const int g = lane >> 2, tg = lane & 3;
    int dsto[NBW];                                    // staging offset of this thread's column pair per n-block, -1 = not stored
    float bias0[NBW], bias1[NBW];
#pragma unroll
    for (int nb = 0; nb < NBW; ++nb) {
        const int jl = wn * NBW * 8 + nb * 8 + 2 * tg;
        dsto[nb] = -1;
        bias0[nb] = s_bias[jl]; bias1[nb] = s_bias[jl + 1];
        if (jl < ncols) {
            const int j = jt0 + jl, p = j / gwp, c = j - p * gwp;
            if (c < a.gwv) dsto[nb] = (p - p0) * R * lds + (c - scol0) + (wm * MT * 16 + g) * lds;
        }
    }
    // transform roles: per source, thread <-> one 8-slot chunk, row lanes stride the rows
    int x_off[kMaxSrc], x_offb[kMaxSrc], x_nch[kMaxSrc], x_ch[kMaxSrc], x_rl[kMaxSrc], x_nrl[kMaxSrc];
    {
        int off = 0, offb = 0;
        for (int i = 0; i < kMaxSrc; ++i) {
            x_off[i] = off; x_offb[i] = offb; x_nch[i] = 1; x_ch[i] = 0; x_rl[i] = 1; x_nrl[i] = 0;
            if (i < d.nsrc) {
                const int cp = d.src[i].cp;
                x_nch[i] = cp >> 3; x_ch[i] = tid % x_nch[i]; x_rl[i] = tid / x_nch[i]; x_nrl[i] = 256 / x_nch[i];
                off += cp; offb += cp * 2;
            }
        }
    }
    // vector pass: thread <-> one 8-slot chunk of one covered plane, row lanes stride the rows
    const int nq = splanes * (swidth >> 3), vq = tid % nq, vrl = tid / nq, vnrl = 256 / nq;
    const int vp = vq / (swidth >> 3), vc = (vq - vp * (swidth >> 3)) * 8;
    float ssum[8], ssq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
    // pass-through copy: thread <-> one destination slot PAIR (plane, 2q, 2q+1), row lanes stride the rows
    int cp_s0 = -1, cp_s1 = -1, cp_dst = 0, cp_rl = 0, cp_nrl = 1;
    if (a.x1) {
        const int npairs = (a.ncopy + 1) >> 1, nitem = 2 * npairs;
        cp_nrl = 256 / nitem; cp_rl = tid / nitem;
        const int it = tid % nitem, p = it / npairs, q = it - p * npairs;
        if (cp_rl < cp_nrl) {
            cp_s0 = logical_slot(a.x1map, 2 * (2 * q) + p);
            cp_s1 = (2 * q + 1 < a.ncopy) ? logical_slot(a.x1map, 2 * (2 * q + 1) + p) : -1;
            cp_dst = p * R * lds + a.copy_dst0 + 2 * q;
        }
    }

    // per-(slice, channel) sums: registers -> scratch rows in the (idle) staging tile -> one thread per column -> fp64 atomics
    auto flush_stats = [&](int t) {                   // CTA-uniform; the staging tile is free here
        __syncthreads();
        float* scr = reinterpret_cast<float*>(St);    // [vnrl][splanes*swidth][2]
        const int ncolt = splanes * swidth;
        if (vrl < vnrl) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                reinterpret_cast<float2*>(scr)[vrl * ncolt + vp * swidth + vc + i] = make_float2(ssum[i], ssq[i]);
                ssum[i] = 0.f; ssq[i] = 0.f;
            }
        }
        __syncthreads();
        for (int i = tid; i < ncolt; i += 256) {
            float sx = 0.f, sq = 0.f;
            for (int l = 0; l < vnrl; ++l) { const float2 v = reinterpret_cast<const float2*>(scr)[l * ncolt + i]; sx += v.x; sq += v.y; }
            const int p = i / swidth, c = i - p * swidth;
            double2* dst = a.tb[p0 + p].fsum + (size_t)t * a.cpo + scol0 + c;
            atomicAdd(&dst->x, (double)sx);
            atomicAdd(&dst->y, (double)sq);
        }
        __syncthreads();
        for (int i = tid; i < splanes * R * lds / 2; i += 256) reinterpret_cast<uint32_t*>(St)[i] = 0u;   // pad slots back to zero
        __syncthreads();
    };

    int cur_t = -1;
    int t = tile_lo / tps, r0 = (tile_lo - t * tps) * R;
    for (int tile = tile_lo, it = 0; tile < tile_hi; ++tile, ++it, r0 += R) {
        const int buf = it & 1;
        if (r0 >= a.Rt) { r0 = 0; ++t; }
        const int rows = min(R, a.Rt - r0);
        if (t != cur_t) {
            if (cur_t >= 0 && a.training) flush_stats(cur_t);
            __syncthreads();
            int off = 0;
            for (int i = 0; i < d.nsrc; ++i) {
                for (int k = tid; k < d.src[i].cp; k += 256)
                    s_aff[tcol(off + k, KP >> 3)] = d.src[i].aff ? d.src[i].aff[(size_t)t * d.src[i].cp + k] : make_float2(1.f, 0.f);
                off += d.src[i].cp;
            }
            cur_t = t;
            __syncthreads();
        }
        mbar_wait(&full[buf], (it >> 1) & 1);
        const unsigned char* rb = raw + (size_t)buf * L.raw_stride;
        // ---- transform: raw rows -> BN affine (+ReLU6) -> padded MMA tile
#pragma unroll
        for (int i = 0; i < kMaxSrc; ++i) {
            if (i < d.nsrc && x_rl[i] < x_nrl[i]) {
                float2 c8[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) c8[q] = s_aff[q * (KP >> 3) + (x_off[i] >> 3) + x_ch[i]];
                const bool clamp = d.src[i].clamp != 0;
                const uint4* srcv = reinterpret_cast<const uint4*>(rb + (size_t)R * x_offb[i]) + x_rl[i] * x_nch[i] + x_ch[i];
                bf16* dstp = As + x_rl[i] * lda + x_off[i] + x_ch[i] * 8;
                const int sstep = x_nrl[i] * x_nch[i], dstep = x_nrl[i] * lda;
                for (int r = x_rl[i]; r < rows; r += x_nrl[i]) {
                    *reinterpret_cast<uint4*>(dstp) = affine8(*srcv, c8, clamp);
                    srcv += sstep; dstp += dstep;
                }
            }
        }
        __syncthreads();
        // ---- MMA: warp (wm, wn) -> rows wm*MT*16.., columns wn*NBW*8..   (accumulators start at the bias)
        float acc[MT][NBW][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) { acc[mt][nb][0] = acc[mt][nb][2] = bias0[nb]; acc[mt][nb][1] = acc[mt][nb][3] = bias1[nb]; }
        {
            const int mi = lane >> 3;
            const bf16* ap = As + (wm * MT * 16 + (lane & 15)) * lda + (lane >> 4) * 8;
            const bf16* bp = Ws + (wn * NBW * 8 + (mi >> 1) * 8 + (lane & 7)) * ldw + (mi & 1) * 8;
            for (int ks = 0; ks < KP; ks += 16) {
                uint32_t af[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) ldsm4(af[mt], ap + mt * 16 * lda + ks);
#pragma unroll
                for (int nb2 = 0; nb2 < NBW / 2; ++nb2) {
                    uint32_t bfr[4];
                    ldsm4(bfr, bp + nb2 * 16 * ldw + ks);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma16816(acc[mt][2 * nb2], af[mt], bfr[0], bfr[1]);
                        mma16816(acc[mt][2 * nb2 + 1], af[mt], bfr[2], bfr[3]);
                    }
                }
            }
        }
        // ---- epilogue: bf16, into the staging rows at the final slot positions
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb) {
            if (dsto[nb] >= 0) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    bf16* sp = St + dsto[nb] + mt * 16 * lds;
                    *reinterpret_cast<uint32_t*>(sp) = pack2(acc[mt][nb][0], acc[mt][nb][1]);
                    *reinterpret_cast<uint32_t*>(sp + 8 * lds) = pack2(acc[mt][nb][2], acc[mt][nb][3]);
                }
            }
        }
        // ---- pass-through half: bit-exact gather of the raw x1 values into the shuffled slots (two slots per store)
        if (cp_s0 >= 0) {
            const unsigned short* xr = reinterpret_cast<const unsigned short*>(rb + (size_t)R * x1_off_rows) + cp_rl * a.x1cp;
            bf16* dp = St + cp_dst + cp_rl * lds;
            const int xstep = cp_nrl * a.x1cp, dstep = cp_nrl * lds;
            for (int r = cp_rl; r < rows; r += cp_nrl) {
                const uint32_t lo = xr[cp_s0], hi = cp_s1 >= 0 ? xr[cp_s1] : 0u;
                *reinterpret_cast<uint32_t*>(dp) = lo | (hi << 16);
                xr += xstep; dp += dstep;
            }
        }
        __syncthreads();
        if (tid == 0 && tile + 2 < tile_hi) issue(tile + 2, buf);
        // ---- store + statistics
        if (vrl < vnrl) {
            bf16* orow = a.out[p0 + vp] + ((size_t)t * a.Rt + r0 + vrl) * a.cpo + scol0 + vc;
            const bf16* srow = St + vp * R * lds + vc + vrl * lds;
            const int ostep = vnrl * a.cpo, sstep = vnrl * lds;
            for (int r = vrl; r < rows; r += vnrl) {
                const uint4 v = *reinterpret_cast<const uint4*>(srow);
                *reinterpret_cast<uint4*>(orow) = v;
                const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = unpack2(w[i]);
                    ssum[2 * i] += f.x; ssq[2 * i] = fmaf(f.x, f.x, ssq[2 * i]);
                    ssum[2 * i + 1] += f.y; ssq[2 * i + 1] = fmaf(f.y, f.y, ssq[2 * i + 1]);
                }
                orow += ostep; srow += sstep;
            }
        }
    }
    if (cur_t >= 0 && a.training) flush_stats(cur_t);

    // ---- last CTA: BatchNorm tables of every output channel (+ the pass-through slots' tables)
    if (a.counter == nullptr) return;
    if (!last_cta(a.counter, gridDim.x * gridDim.y)) return;
    for (int j = tid; j < d.NPall; j += 256) {
        int p, s, l, n;
        if (pw_col(d, j, p, s, l, n)) bn_finalize_channel(a.tb[p], a.cpo, s, d.layer[l], n, (double)a.Rt, a.training);
        else if (s < a.gwv) {
            for (int t = 0; t < kT; ++t) { a.tb[p].aff[(size_t)t * a.cpo + s] = make_float2(0.f, 0.f); a.tb[p].bnp[(size_t)t * a.cpo + s] = make_float2(0.f, 1.f); }
        }
    }
    if (a.x1) {
        for (int i = tid; i < 2 * a.ncopy * kT; i += 256) {
            const int t = i / (2 * a.ncopy), q = i - t * 2 * a.ncopy, p = q / a.ncopy, c = q - p * a.ncopy;
            const int ss = logical_slot(a.x1map, 2 * c + p);
            a.tb[p].aff[(size_t)t * a.cpo + a.copy_dst0 + c] = a.x1aff ? a.x1aff[(size_t)t * a.x1cp + ss] : make_float2(1.f, 0.f);
            a.tb[p].bnp[(size_t)t * a.cpo + a.copy_dst0 + c] = a.x1bnp ? a.x1bnp[(size_t)t * a.x1cp + ss] : make_float2(0.f, 1.f);
        }
    }
}

}  // namespace v2
}  // namespace cdra
#endif
