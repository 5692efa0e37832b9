// v2 tower: pointwise forward on tcgen05 for the plain-output layers (pw1 of every unit: one output tensor, no
// pass-through, no shuffle).  out[r][n] = sum_k act(src)[r][k] W[k][n] + b[n] with both operands K-major: the transform
// pass writes the activated 128-row tile straight into the 128-byte-swizzled layout, the layer's bf16 weights sit in
// the same layout for the whole CTA, one elected thread issues tcgen05.mma (M = 128, N = NPall, K = 16 per
// instruction) into a TMEM accumulator, tcgen05.ld brings it back for the bias / bf16 / BatchNorm-sum epilogue.
// One CTA per SM, 16 warps, deep TMA ring (see pw_wgrad_tc_kernel for the reasoning).
#pragma once
#ifndef CDRA_EMU
#include "v2_pw.cuh"
#include "v2_umma.cuh"

namespace cdra {
namespace v2 {

struct PwFwdTcSmem { int aff, bias, w, a, st, raw, total, raw_stride, nkb, np, lds, tmem_cols; };
inline __host__ __device__ PwFwdTcSmem pw_fwd_tc_smem(int KP, int NPall, int cpo, int src_cp_sum, int nbuf, int nt, int nplanes = 1, int x1cp = 0) {
    PwFwdTcSmem s;
    s.nkb = (KP + 63) / 64; s.np = (NPall + 15) & ~15; s.lds = pad_ld(cpo);
    s.tmem_cols = 32; while (s.tmem_cols < s.np) s.tmem_cols *= 2;
    int off = 384;                                     // mbarriers (8 TMA + 1 MMA) + TMEM slot; [128, 384): the MMA issuer's descriptor table
    s.aff = off; off += s.nkb * 64 * 8;
    s.bias = off; off += s.np * 4;
    off = (off + 1023) & ~1023;
    s.w = off; off += s.nkb * s.np * 128;              // np is a multiple of 16 -> every block is a multiple of 1024 bytes... (np * 128)
    off = (off + 1023) & ~1023;
    s.a = off; off += s.nkb * 128 * 128;
    { int st = nplanes * 128 * s.lds * 2; const int scr = nt * 64; if (st < scr) st = scr; s.st = off; off += (st + 127) & ~127; }   // doubles as the sum-flush scratch
    s.raw_stride = (128 * (src_cp_sum + x1cp) * 2 + 127) & ~127;
    s.raw = off; off += nbuf * s.raw_stride;
    s.total = off + 1024;
    return s;
}

// role timeline of block 0 (cdra_debug_timeline, CDRA_TIMELINE=1): tile 2 of the CTA's schedule phase by phase
__device__ unsigned long long g_tc_ts[16];
CDRA_DEV unsigned long long gtimer_tc() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TC_TS(i) do { if (a.timeline && blockIdx.x == 0 && threadIdx.x == 0) g_tc_ts[i] = gtimer_tc(); } while (0)
template <int NT>
__global__ void __launch_bounds__(NT, 1) pw_fwd_tc_kernel(const PwFwdArgs a) {
    constexpr int R = 128;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    TC_TS(0);
    const PwDesc& d = *a.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    const int KP = d.KP, NP = d.NPall, cpo = a.cpo;
    int src_cp_sum = 0;
    for (int i = 0; i < d.nsrc; ++i) src_cp_sum += d.src[i].cp;
    const int nplanes = d.cols.nplanes, gwp = d.cols.gwp, x1cp = a.x1 ? a.x1cp : 0;
    const PwFwdTcSmem L = pw_fwd_tc_smem(KP, NP, cpo, src_cp_sum, a.nbuf, NT, nplanes, x1cp);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* mma_done = full + 8;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + 96);
    float2* s_aff = reinterpret_cast<float2*>(smem + L.aff);
    float* s_bias = reinterpret_cast<float*>(smem + L.bias);
    unsigned char* Ws = smem + L.w;
    unsigned char* As = smem + L.a;
    bf16* St = reinterpret_cast<bf16*>(smem + L.st);
    unsigned char* raw = smem + L.raw;
    const int lds = L.lds;

    const int tps = (a.Rt + R - 1) / R, ntile = kT * tps;
    const int tile_lo = blockIdx.x * a.tiles_per_cta, tile_hi = min(ntile, tile_lo + a.tiles_per_cta);
    if (warp == 0) tmem_alloc(s_tmem, (uint32_t)L.tmem_cols);
    if (tid == 0) { for (int b = 0; b < 9; ++b) mbar_init(&full[b], 1); mbar_fence_init(); }
    // weights (prepared by pw_prep long before) into the swizzled K-major layout; zero K / N padding
    for (int i = tid; i < L.nkb * L.np * 8; i += NT) reinterpret_cast<uint4*>(Ws)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < L.nkb * 128 * 8; i += NT) reinterpret_cast<uint4*>(As)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < nplanes * R * lds / 2; i += NT) reinterpret_cast<uint32_t*>(St)[i] = 0u;    // pad slots stay zero
    __syncthreads();
    for (int i = tid; i < NP * (KP >> 3); i += NT) {
        const int j = i / (KP >> 3), c = i - j * (KP >> 3);
        *reinterpret_cast<uint4*>(Ws + (size_t)(c >> 3) * L.np * 128 + sw128_offset(j, (c & 7) * 8, L.np)) =
            *reinterpret_cast<const uint4*>(d.wf + (size_t)j * KP + c * 8);
    }
    for (int j = tid; j < L.np; j += NT) s_bias[j] = j < NP ? d.bias[j] : 0.f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t idesc = umma_idesc(128, L.np, 0, 0);

    // the operand addresses never change: descriptors built once (the issuing thread is on every tile's critical path)
    uint64_t* s_md = reinterpret_cast<uint64_t*>(smem + 128);
    if (tid == 0) {
        const uint32_t aa = smem_u32(As), wa = smem_u32(Ws);
        for (int kb = 0; kb < L.nkb; ++kb)
            for (int ks = 0; ks < 4; ++ks) {
                s_md[2 * (4 * kb + ks)] = umma_desc(aa + kb * R * 128 + ks * 32, 16, 1024);
                s_md[2 * (4 * kb + ks) + 1] = umma_desc(wa + kb * L.np * 128 + ks * 32, 16, 1024);
            }
    }
    auto issue = [&](int tile, int buf) {
        const int t = tile / tps, r0 = (tile - t * tps) * R, rows = min(R, a.Rt - r0);
        unsigned char* dst = raw + (size_t)buf * L.raw_stride;
        mbar_expect_tx(&full[buf], (uint32_t)rows * (src_cp_sum + x1cp) * 2);
        int off = 0;
        for (int i = 0; i < d.nsrc; ++i) {
            bulk_g2s(dst + (size_t)R * off, d.src[i].data + ((size_t)t * a.Rt + r0) * d.src[i].cp, rows * d.src[i].cp * 2, &full[buf]);
            off += d.src[i].cp * 2;
        }
        if (x1cp) bulk_g2s(dst + (size_t)R * src_cp_sum * 2, a.x1 + ((size_t)t * a.Rt + r0) * x1cp, rows * x1cp * 2, &full[buf]);
    };
    TC_TS(1);
    pdl_wait();
    TC_TS(2);
    if (tid == 32) for (int b = 0; b < a.nbuf; ++b) if (tile_lo + b < tile_hi) issue(tile_lo + b, b);

    // transform role: thread <-> one 8-slot chunk over the concatenated sources, row lanes stride the rows
    const int nqx = src_cp_sum >> 3, xq = tid % nqx, xrl = tid / nqx, xnrl = NT / nqx;
    int xsrc = 0, xch = xq, xoffb = 0;
    while (xsrc < d.nsrc - 1 && xch >= (d.src[xsrc].cp >> 3)) { xch -= d.src[xsrc].cp >> 3; xoffb += d.src[xsrc].cp * 2; ++xsrc; }
    const int xnch = d.src[xsrc].cp >> 3;
    const bool sclamp = d.src[xsrc].clamp != 0;
    // store role: thread <-> one 8-column chunk of the output, row lanes stride the rows
    const int nqp = cpo >> 3, nq = nplanes * nqp, vq = tid % nq, vrl = tid / nq, vnrl = NT / nq, vp = vq / nqp, vc = (vq - vp * nqp) * 8;
    float ssum[8], ssq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
    float2 c8[8];
    // pass-through copy (stride-1 unit tail): thread <-> one destination slot PAIR (plane, 2q, 2q+1), row lanes stride the rows
    int cp_s0 = -1, cp_s1 = -1, cp_dst = 0, cp_rl = 0, cp_nrl = 1;
    if (a.x1) {
        const int npairs = (a.ncopy + 1) >> 1, nitem = 2 * npairs;
        cp_nrl = NT / nitem; cp_rl = tid / nitem;
        const int it = tid % nitem, p = it / npairs, q = it - p * npairs;
        if (cp_rl < cp_nrl) {
            cp_s0 = logical_slot(a.x1map, 2 * (2 * q) + p);
            cp_s1 = (2 * q + 1 < a.ncopy) ? logical_slot(a.x1map, 2 * (2 * q + 1) + p) : -1;
            cp_dst = p * R * lds + a.copy_dst0 + 2 * q;
        }
    }
    // epilogue role: warp -> TMEM lane group (rows 32 * (warp & 3) + lane), column share warp >> 2 of NT/128
    const int lg = warp & 3, cshare = warp >> 2, ncshare = NT / 128;

    auto flush_stats = [&](int t) {                    // CTA-uniform; the staging tile is free here
        __syncthreads();
        float2* scr = reinterpret_cast<float2*>(St);   // [vnrl][nplanes * cpo]
        const int ncolt = nplanes * cpo;
        if (vrl < vnrl) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { scr[vrl * ncolt + vp * cpo + vc + i] = make_float2(ssum[i], ssq[i]); ssum[i] = 0.f; ssq[i] = 0.f; }
        }
        __syncthreads();
        for (int i = tid; i < ncolt; i += NT) {
            float sx = 0.f, sq = 0.f;
            for (int l = 0; l < vnrl; ++l) { const float2 v = scr[l * ncolt + i]; sx += v.x; sq += v.y; }
            const int p = i / cpo, c = i - p * cpo;
            double2* dst = a.tb[p].fsum + (size_t)t * cpo + c;
            atomicAdd(&dst->x, (double)sx); atomicAdd(&dst->y, (double)sq);
        }
        __syncthreads();
        for (int i = tid; i < nplanes * R * lds / 2; i += NT) reinterpret_cast<uint32_t*>(St)[i] = 0u;   // pad slots back to zero
        __syncthreads();
    };

    int cur_t = -1;
    for (int tile = tile_lo, it = 0; tile < tile_hi; ++tile, ++it) {
        const int buf = it % a.nbuf;
        const int t = tile / tps, r0 = (tile - t * tps) * R, rows = min(R, a.Rt - r0);
        if (t != cur_t) {
            if (cur_t >= 0 && a.training) flush_stats(cur_t);
            __syncthreads();
            int off = 0;
            for (int i = 0; i < d.nsrc; ++i) {
                for (int k = tid; k < d.src[i].cp; k += NT) s_aff[off + k] = d.src[i].aff ? d.src[i].aff[(size_t)t * d.src[i].cp + k] : make_float2(1.f, 0.f);
                off += d.src[i].cp;
            }
            cur_t = t;
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 8; ++q) c8[q] = s_aff[xq * 8 + q];
        }
        if (it == 2) TC_TS(3);
        mbar_wait(&full[buf], (it / a.nbuf) & 1);
        if (it == 2) TC_TS(4);
        const unsigned char* rb = raw + (size_t)buf * L.raw_stride;
        // ---- transform: raw rows -> BN affine (+ReLU6) -> swizzled K-major tile (rows past the slice end are zero)
        if (xrl < xnrl) {
            const uint4* sv = reinterpret_cast<const uint4*>(rb + (size_t)R * xoffb);
#pragma unroll 2
            for (int r = xrl; r < R; r += xnrl) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (r < rows) v = affine8(sv[r * xnch + xch], c8, sclamp);
                *reinterpret_cast<uint4*>(As + sw128_offset(r, xq * 8, R)) = v;
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (it == 2) TC_TS(5);
        if (tid == 0) {
            tc_fence_after();
            for (int q = 0; q < 4 * L.nkb; ++q) umma_bf16(tmem, s_md[2 * q], s_md[2 * q + 1], idesc, q > 0);
            umma_commit(mma_done);
        }
        mbar_wait(mma_done, it & 1);
        if (it == 2) TC_TS(6);
        tc_fence_after();
        // ---- epilogue: TMEM -> registers -> + bias -> bf16 -> staging rows
        {
            const int row = 32 * lg + lane;
            for (int c0 = cshare * 8; c0 < NP; c0 += ncshare * 8) {          // GEMM column j = plane * gwp + slot
                float v[8];
                tmem_ld8(tmem + ((uint32_t)(32 * lg) << 16) + (uint32_t)c0, v);
                const int p = c0 / gwp, sl = c0 - p * gwp;
                if (sl >= a.gwv) continue;
                uint4 o;
                o.x = pack2(v[0] + s_bias[c0], v[1] + s_bias[c0 + 1]); o.y = pack2(v[2] + s_bias[c0 + 2], v[3] + s_bias[c0 + 3]);
                o.z = pack2(v[4] + s_bias[c0 + 4], v[5] + s_bias[c0 + 5]); o.w = pack2(v[6] + s_bias[c0 + 6], v[7] + s_bias[c0 + 7]);
                bf16* sp = St + (size_t)p * R * lds + (size_t)row * lds + sl;
                if (sl + 8 <= a.gwv) *reinterpret_cast<uint4*>(sp) = o;
                else {                                  // the stored columns end inside this chunk (the pass-through slots follow): pairs only
                    const uint32_t* ow = reinterpret_cast<const uint32_t*>(&o);
                    for (int q = 0; sl + 2 * q + 1 < a.gwv; ++q) reinterpret_cast<uint32_t*>(sp)[q] = ow[q];
                }
            }
        }
        // ---- pass-through half: bit-exact gather of the raw x1 values into the shuffled slots (two slots per store)
        if (cp_s0 >= 0) {
            const unsigned short* xr = reinterpret_cast<const unsigned short*>(rb + (size_t)R * src_cp_sum * 2) + cp_rl * x1cp;
            bf16* dp = St + cp_dst + cp_rl * lds;
            const int xstep = cp_nrl * x1cp, dstep = cp_nrl * lds;
            for (int r = cp_rl; r < rows; r += cp_nrl) {
                const uint32_t lo = xr[cp_s0], hi = cp_s1 >= 0 ? xr[cp_s1] : 0u;
                *reinterpret_cast<uint32_t*>(dp) = lo | (hi << 16);
                xr += xstep; dp += dstep;
            }
        }
        tc_fence_before();
        __syncthreads();
        if (it == 2) TC_TS(7);
        if (tid == 32 && tile + a.nbuf < tile_hi) issue(tile + a.nbuf, buf);       // the raw rows (sources and x1) are consumed
        // ---- store + statistics
        if (vrl < vnrl) {
            bf16* orow = a.out[vp] + ((size_t)t * a.Rt + r0 + vrl) * cpo + vc;
            const bf16* srow = St + (size_t)vp * R * lds + vc + vrl * lds;
            const int ostep = vnrl * cpo, sstep = vnrl * lds;
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
        if (it == 2) TC_TS(8);
    }
    TC_TS(9);
    if (cur_t >= 0 && a.training) flush_stats(cur_t);
    tc_fence_before();
    __syncthreads();
    TC_TS(10);
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)L.tmem_cols);

    // ---- last CTA: BatchNorm tables of every output channel
    if (a.counter == nullptr) return;
    if (!last_cta(a.counter, gridDim.x * gridDim.y)) return;
    for (int j = tid; j < NP; j += NT) {
        int p, s, l, n;
        if (pw_col(d, j, p, s, l, n)) bn_finalize_channel(a.tb[p], cpo, s, d.layer[l], n, (double)a.Rt, a.training);
        else if (s < a.gwv) {
            for (int t = 0; t < kT; ++t) { a.tb[p].aff[(size_t)t * cpo + s] = make_float2(0.f, 0.f); a.tb[p].bnp[(size_t)t * cpo + s] = make_float2(0.f, 1.f); }
        }
    }
    if (a.x1) {
        for (int i = tid; i < 2 * a.ncopy * kT; i += NT) {
            const int t = i / (2 * a.ncopy), q = i - t * 2 * a.ncopy, p = q / a.ncopy, c = q - p * a.ncopy;
            const int ss = logical_slot(a.x1map, 2 * c + p);
            a.tb[p].aff[(size_t)t * cpo + a.copy_dst0 + c] = a.x1aff ? a.x1aff[(size_t)t * a.x1cp + ss] : make_float2(1.f, 0.f);
            a.tb[p].bnp[(size_t)t * cpo + a.copy_dst0 + c] = a.x1bnp ? a.x1bnp[(size_t)t * a.x1cp + ss] : make_float2(0.f, 1.f);
        }
    }
}

}  // namespace v2
}  // namespace cdra
#endif
