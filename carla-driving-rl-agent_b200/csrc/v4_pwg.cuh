// v4: warp-specialised tcgen05 GEMM family for the GEMM-dense pointwise layers (stage 3 and the 464 -> 768 head conv,
// core/architectures.py:130,134,140,170): forward and data gradient.  Both are computed TRANSPOSED,
//     forward        out^T  [j ][r] = sum_kk Wf[j ][kk] * act(src)[r][kk]       (+ bias, BatchNorm sums of the stored values)
//     data gradient  dsrc^T [kk][r] = sum_j  Wb[kk][j ] * dR[r][j]              (+ BatchNorm-backward sums of the source)
// so that the accumulator row (TMEM lane) is a CHANNEL and the accumulator column a tile row: an epilogue thread owns one
// channel and walks the rows, which makes every per-channel statistic a plain register accumulation (no cross-thread
// reduction, no per-element unpacking) and the bias a per-thread constant.
//
//   A operand  = weights: pw_prep leaves them in HBM as 64-column blocks of 128-byte-swizzled rows (PwDesc::wfs / wbs), so
//                an M block (<= 128 channels) of one K block is ONE contiguous bulk copy.  Resident for the CTA's life when
//                all K blocks fit next to the ring, else streamed through the ring with the activation blocks.
//   B operand  = activations: producer warps read the raw rows straight from HBM / L2 (16-byte loads, software
//                pipelined one K block ahead), apply the producer layer's BatchNorm affine (+ReLU6) -- or the layer's own
//                BatchNorm(+ReLU6) backward for the data gradient -- and write the 128-byte-swizzled K-major tile.
//   MMA        = one thread, tcgen05.mma M = 128 (channels) x N = 128 (rows) x K = 16, accumulator double buffered in TMEM.
//   epilogue   = 4 warps, tcgen05.ld 32x32b.x16; bf16 values are transposed through a shared-memory tile (2-byte stores,
//                conflict free: a warp's lanes are consecutive channels) and leave as 16-byte row chunks.
// Roles are chained by mbarriers only.  grid = (row-tile groups, M blocks).
#pragma once
#ifndef CDRA_EMU
#include "v3_pw_bwd.cuh"

namespace cdra {
namespace v2 {

constexpr int kGProd = 8, kGEpi = 4, kGThreads = 32 * (2 + kGProd + kGEpi);
constexpr int kGProdThreads = 32 * kGProd, kGEpiThreads = 32 * kGEpi;
constexpr int kGRows = 128;             // rows per tile = the MMA N dimension
constexpr int kGMaxStages = 8;

CDRA_DEV void cp_async16(void* dst, const void* src, bool valid) {     // 16 bytes global -> shared, L2 only; !valid: zero fill
    const uint32_t sz = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
CDRA_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> CDRA_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct PwgSmem { int ck, tab, a, ring, st, total, stage_bytes, nkb, b_off; };
// kdim: reduction length (KP forward, NPall data gradient); tab_bytes: per-slice constant table; extra: bytes per stage
// next to the activation block (the raw output rows of the data gradient)
inline __host__ __device__ PwgSmem pwg_smem(int kdim, int tab_bytes, bool ares, int stw, int nstage, int extra = 0) {
    PwgSmem s;
    s.nkb = (kdim + 63) >> 6;
    int off = 1024;                                     // mbarriers + TMEM slot
    s.ck = off; off += 128 * 4;                         // chunk table (<= 96 eight-column chunks)
    s.tab = off; off += tab_bytes;
    off = (off + 1023) & ~1023;
    s.a = off; off += ares ? s.nkb * 16384 : 0;
    s.b_off = ares ? 0 : 16384;
    s.stage_bytes = s.b_off + 16384 + extra;            // [A block] B block [raw rows]
    s.ring = off; off += nstage * s.stage_bytes;
    s.st = off; off += kGRows * stw * 2;
    s.total = off + 1024;                               // slack for the manual 1024-byte alignment of the base
    return s;
}

struct GCur { int it, kb, t, r0; };
// role timeline of block (0, 0) in %globaltimer nanoseconds, recorded when the launch was made with CDRA_TIMELINE=1 and read
// back by cdra_debug_timeline (profiles/pwg_timeline_probe.py): where a launch's time goes between the roles' hand-offs
__device__ unsigned long long g_pwg_ts[16];
CDRA_DEV unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PWG_TS(i) do { if (a.timeline && blockIdx.x == 0 && blockIdx.y == 0) g_pwg_ts[i] = gtimer(); } while (0)

// ======================================================================================== forward
template <int D>       // D: K blocks in flight per producer thread (ring depth >= D + 1)
__global__ void __launch_bounds__(kGThreads, 1) pwg_fwd_kernel(const PwFwdArgs a, const int ares) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    if (threadIdx.x == 0) PWG_TS(0);
    const PwDesc& d = pw_desc_to_smem(a.d, smem + 520);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    const int KP = d.KP, NPall = d.NPall, nsrc = d.nsrc, S = a.nbuf, Rt = a.Rt;
    const GBlock gb = a.blk[blockIdx.y];
    const int stw = (gb.n + 7) & ~7;
    const PwgSmem L = pwg_smem(KP, ((KP + 63) & ~63) * 8, ares != 0, stw, S);
    const int nkb = L.nkb, stage_bytes = L.stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kGMaxStages;
    uint64_t* tm_full = empty + kGMaxStages;
    uint64_t* tm_empty = tm_full + 2;
    uint64_t* a_full = tm_empty + 2;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + 512);
    int* s_ck = reinterpret_cast<int*>(smem + L.ck);
    float2* s_aff = reinterpret_cast<float2*>(smem + L.tab);
    unsigned char* As = smem + L.a;
    unsigned char* ring = smem + L.ring;
    unsigned short* St = reinterpret_cast<unsigned short*>(smem + L.st);

    const int tps = (Rt + kGRows - 1) / kGRows, ntile = kT * tps;
    const int tile_lo = blockIdx.x * a.tiles_per_cta, tile_hi = min(ntile, tile_lo + a.tiles_per_cta);
    const int my_tiles = max(0, tile_hi - tile_lo);
    const int t_first = tile_lo / tps, r_first = (tile_lo - t_first * tps) * kGRows;

    if (warp == 0) tmem_alloc(s_tmem, 2 * kGRows);
    if (tid == 32) {
        for (int s = 0; s < kGMaxStages; ++s) { mbar_init(&full[s], kGProd + (ares ? 0 : 1)); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tm_full[b], 1); mbar_init(&tm_empty[b], kGEpi); }
        mbar_init(a_full, 1);
        mbar_fence_init();
    }
    for (int ck = tid; ck < nkb * 8; ck += kGThreads) {          // 8-slot chunk of the GEMM K index -> (source, clamp, slot)
        int v = -1, off = 0;
        for (int i = 0; i < nsrc; ++i) {
            const int nch = d.src[i].cp >> 3;
            if (ck >= off && ck < off + nch) v = (i << 24) | ((d.src[i].clamp ? 1 : 0) << 23) | ((ck - off) * 8);
            off += nch;
        }
        s_ck[ck] = v;
    }
    // weight rows beyond the block's channels must read as zero: clear the A tiles once (bulk copies only cover gb.n rows)
    if (ares) { for (int i = tid; i < nkb * 1024; i += kGThreads) reinterpret_cast<uint4*>(As)[i] = make_uint4(0, 0, 0, 0); }
    else { for (int i = tid; i < S * 1024; i += kGThreads) reinterpret_cast<uint4*>(ring + (size_t)(i >> 10) * stage_bytes)[i & 1023] = make_uint4(0, 0, 0, 0); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    if (tid == 0) PWG_TS(1);
    if (ares && tid == 32) {                            // resident weights (prepared long before this launch)
        mbar_expect_tx(a_full, (uint32_t)nkb * gb.n * 128);
        for (int kb = 0; kb < nkb; ++kb) bulk_g2s(As + (size_t)kb * 16384, d.wfs + ((size_t)kb * NPall + gb.base) * 64, gb.n * 128, a_full);
    }
    pdl_wait();
    if (tid == 0) PWG_TS(2);

    auto advance = [&](GCur& c) { if (++c.kb == nkb) { c.kb = 0; ++c.it; c.r0 += kGRows; if (c.r0 >= Rt) { c.r0 = 0; ++c.t; } } };

    if (warp == 0) {
        // ================================================================ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, kGRows, 0, 0);
            if (ares) mbar_wait(a_full, 0);
            PWG_TS(3);
            int s = 0, ph = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int b = it & 1;
                mbar_wait(&tm_empty[b], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[s], ph);
                    if (it == 0 && kb == 0) PWG_TS(4);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(ring + (size_t)s * stage_bytes);
                    const uint32_t aa = ares ? smem_u32(As) + (uint32_t)kb * 16384u : sb, ba = ares ? sb : sb + 16384u;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_bf16(tmem + (uint32_t)(b * kGRows), umma_desc(aa + ks * 32, 16, 1024), umma_desc(ba + ks * 32, 16, 1024), idesc, (kb | ks) != 0);
                    umma_commit(&empty[s]);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
                umma_commit(&tm_full[b]);
            }
        }
    } else if (warp == 1) {
        // ================================================================ weight streamer (only when the weights are not resident)
        if (lane == 0 && !ares) {
            int s = 0, ph = 0;
            for (int it = 0; it < my_tiles; ++it)
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], (uint32_t)gb.n * 128);
                    bulk_g2s(ring + (size_t)s * stage_bytes, d.wfs + ((size_t)kb * NPall + gb.base) * 64, gb.n * 128, &full[s]);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
        }
    } else if (warp < 2 + kGProd) {
        // ================================================================ producers: raw rows -> act -> swizzled K-major tile
        // cp.async lands this thread's 16-byte chunks at their final (swizzled) position D K blocks ahead; the SAME thread
        // later applies the BatchNorm affine (+ReLU6) in place, so no barrier sits between the copy and the transform
        const int pt = tid - 64, c = pt & 7, rl = pt >> 3;       // chunk column of the K block, row lane (rows rl + 32 i)
        const uint32_t swz = (uint32_t)((c ^ (rl & 7)) << 4);   // (rl + 32 i) & 7 == rl & 7
        const bf16* sp0 = d.src[0].data; const bf16* sp1 = nsrc > 1 ? d.src[1].data : nullptr; const bf16* sp2 = nsrc > 2 ? d.src[2].data : nullptr;
        const int cp0 = d.src[0].cp, cp1 = nsrc > 1 ? d.src[1].cp : 0, cp2 = nsrc > 2 ? d.src[2].cp : 0;
        const int nq = my_tiles * nkb;
        const int b_off = L.b_off;
        int is = 0, iph = 0;                            // issue cursor: ring stage / phase
        auto issue = [&](const GCur& cu) {
            mbar_wait(&empty[is], iph ^ 1);
            const int e = s_ck[cu.kb * 8 + c];
            const int si = e >> 24, col = e & 0xffff;
            const bf16* sp = si <= 0 ? sp0 : (si == 1 ? sp1 : sp2);
            const int cp = si <= 0 ? cp0 : (si == 1 ? cp1 : cp2);
            const int rows = min(kGRows, Rt - cu.r0);
            const bf16* base = sp + ((size_t)cu.t * Rt + cu.r0 + rl) * cp + (e >= 0 ? col : 0);
            unsigned char* Bs = ring + (size_t)is * stage_bytes + b_off + swz + rl * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = e >= 0 && rl + 32 * i < rows;
                cp_async16(Bs + i * 32 * 128, ok ? base + (size_t)(32 * i) * cp : sp0, ok);
            }
            if (++is == S) { is = 0; iph ^= 1; }
        };
        GCur lc{0, 0, t_first, r_first}, pc = lc;
        for (int j = 0; j < D; ++j) { if (j < nq) { issue(lc); advance(lc); } cp_async_commit(); }
        int s = 0, cur_t = -1;
        for (int q = 0; q < nq; ++q) {
            if (q + D < nq) { issue(lc); advance(lc); }
            cp_async_commit();
            if (pc.t != cur_t) {                        // per-slice (scale, shift) of every K column, laid out [kb][q2][c] (conflict-free 16-byte reads)
                named_bar_sync(2, kGProdThreads);
                for (int kk = pt; kk < nkb * 64; kk += kGProdThreads) {
                    const int e = s_ck[kk >> 3];
                    float2 v = make_float2(1.f, 0.f);
                    if (e >= 0) {
                        const PwSrc& Sx = d.src[e >> 24];
                        if (Sx.aff) v = Sx.aff[(size_t)pc.t * Sx.cp + (e & 0xffff) + (kk & 7)];
                    }
                    s_aff[(((kk >> 6) * 32 + ((kk & 7) >> 1) * 8 + ((kk >> 3) & 7)) << 1) | (kk & 1)] = v;
                }
                named_bar_sync(2, kGProdThreads);
                cur_t = pc.t;
            }
            const int e = s_ck[pc.kb * 8 + c];
            const int rows = min(kGRows, Rt - pc.r0);
            const float4* ap = reinterpret_cast<const float4*>(s_aff) + pc.kb * 32 + c;
            const float4 a0 = ap[0], a1 = ap[8], a2 = ap[16], a3 = ap[24];
            const float2 c8[8] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w),
                                  make_float2(a2.x, a2.y), make_float2(a2.z, a2.w), make_float2(a3.x, a3.y), make_float2(a3.z, a3.w)};
            const bool clamp = ((e >> 23) & 1) != 0;
            cp_async_wait<D>();                         // this thread's chunks of K block q have landed
            unsigned char* Bs = ring + (size_t)s * stage_bytes + b_off + swz + rl * 128;
            if (e >= 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (rl + 32 * i < rows) {
                        uint4* bp = reinterpret_cast<uint4*>(Bs + i * 32 * 128);
                        *bp = affine8(*bp, c8, clamp);
                    }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);

            if (++s == S) s = 0;
            advance(pc);
        }
    } else {
        // ================================================================ epilogue: one channel per thread
        const int lg = warp & 3, m = 32 * lg + lane, et = tid - 32 * (2 + kGProd);
        int p = 0, sl = 0, l = 0, nn = 0;
        const bool valid = m < gb.n && pw_col(d, gb.base + m, p, sl, l, nn);
        const float bias = valid ? d.bias[gb.base + m] : 0.f;
        const int plane = gb.aux0, slot0 = gb.aux1, cpo = a.cpo;
        const int nst = min(gb.n, a.gwv - slot0);          // stored slots of this block (even)
        const int nch = nst >> 3, ntail = (nst & 7) >> 1;    // full 16-byte chunks, trailing 4-byte pairs
        const int cc = et & 15, crl = et >> 4;               // copy role: chunk, row lane (rows crl + 8 i)
        double2* fsum = a.tb[plane].fsum;
        bf16* outp = a.out[plane];
        float s1 = 0.f, s2 = 0.f;
        const bool in_st = m < stw;
        auto flush = [&](int t) {
            if (valid && a.training && (s1 != 0.f || s2 != 0.f)) {
                double2* dst = fsum + (size_t)t * cpo + sl;
                atomicAdd(&dst->x, (double)s1); atomicAdd(&dst->y, (double)s2);
            }
            s1 = s2 = 0.f;
        };
        int t = t_first, r0 = r_first, cur_t = -1;
        for (int it = 0; it < my_tiles; ++it) {
            const int b = it & 1, rows = min(kGRows, Rt - r0);
            if (t != cur_t) { if (cur_t >= 0) flush(cur_t); cur_t = t; }
            mbar_wait(&tm_full[b], (it >> 1) & 1);
            if (et == 0 && it == 0) PWG_TS(5);
            tc_fence_after();
            named_bar_sync(1, kGEpiThreads);               // the previous tile's rows have left the staging tile
            const uint32_t taddr = tmem + ((uint32_t)(32 * lg) << 16) + (uint32_t)(b * kGRows);
#pragma unroll 2
            for (int c0 = 0; c0 < kGRows; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const unsigned short h = __bfloat16_as_ushort(__float2bfloat16_rn(v[i] + bias));
                    if (in_st) St[(c0 + i) * stw + m] = h;
                    if (c0 + i < rows) { const float x = __uint_as_float((uint32_t)h << 16); s1 += x; s2 = fmaf(x, x, s2); }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tm_empty[b]);
            named_bar_sync(1, kGEpiThreads);
            // staged rows -> HBM: 16-byte chunks (+ 4-byte pairs where the stored columns end inside a chunk)
            bf16* orow = outp + ((size_t)t * Rt + r0) * cpo + slot0;
            if (cc < nch) {
#pragma unroll 4
                for (int r = crl; r < rows; r += 8)
                    *reinterpret_cast<uint4*>(orow + (size_t)r * cpo + cc * 8) = *reinterpret_cast<const uint4*>(St + r * stw + cc * 8);
            } else if (cc == nch && ntail > 0) {
                for (int r = crl; r < rows; r += 8)
                    for (int w = 0; w < ntail; ++w)
                        *reinterpret_cast<uint32_t*>(orow + (size_t)r * cpo + cc * 8 + 2 * w) = *reinterpret_cast<const uint32_t*>(St + r * stw + cc * 8 + 2 * w);
            }
            if (et == 0 && it == 0) PWG_TS(6);
            r0 += kGRows; if (r0 >= Rt) { r0 = 0; ++t; }
        }
        if (et == 0) PWG_TS(7);
        if (cur_t >= 0) flush(cur_t);
        if (et == 0) PWG_TS(8);
    }

    tc_fence_before();
    __syncthreads();
    if (tid == 0) PWG_TS(9);
    if (warp == 0) tmem_dealloc(tmem, 2 * kGRows);
    // ---- last CTA: BatchNorm tables of every output channel (+ the pass-through slots' tables)
    if (a.counter == nullptr) return;
    if (!last_cta(a.counter, gridDim.x * gridDim.y)) return;
    if (tid == 0 && a.timeline) g_pwg_ts[10] = gtimer();
    for (int j = tid; j < NPall; j += kGThreads) {
        int p, s, l, n;
        if (pw_col(d, j, p, s, l, n)) bn_finalize_channel(a.tb[p], a.cpo, s, d.layer[l], n, (double)Rt, a.training);
        else if (s < a.gwv) {
            for (int t = 0; t < kT; ++t) { a.tb[p].aff[(size_t)t * a.cpo + s] = make_float2(0.f, 0.f); a.tb[p].bnp[(size_t)t * a.cpo + s] = make_float2(0.f, 1.f); }
        }
    }
    if (a.x1) {
        for (int i = tid; i < 2 * a.ncopy * kT; i += kGThreads) {
            const int t = i / (2 * a.ncopy), q = i - t * 2 * a.ncopy, p = q / a.ncopy, c = q - p * a.ncopy;
            const int ss = logical_slot(a.x1map, 2 * c + p);
            a.tb[p].aff[(size_t)t * a.cpo + a.copy_dst0 + c] = a.x1aff ? a.x1aff[(size_t)t * a.x1cp + ss] : make_float2(1.f, 0.f);
            a.tb[p].bnp[(size_t)t * a.cpo + a.copy_dst0 + c] = a.x1bnp ? a.x1bnp[(size_t)t * a.x1cp + ss] : make_float2(0.f, 1.f);
        }
    }
    __syncthreads();
    if (tid == 0 && a.timeline) g_pwg_ts[11] = gtimer();
}

// pass-through half of a stride-1 unit (core/architectures.py:142-144): out_p[n0p + i] = x1[logical 2i + p], bit exact
struct PassFwdArgs { const bf16* x1; int x1cp; SlotMap x1map; bf16* out[2]; int cpo, ncopy, copy_dst0; long long rows; };
__global__ void __launch_bounds__(256) pass_fwd_kernel(const PassFwdArgs a) {
    pdl_trigger();
    pdl_wait();
    const long long total = a.rows * a.ncopy;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const long long r = idx / a.ncopy; const int i = (int)(idx - r * a.ncopy);
        const bf16* xr = a.x1 + r * a.x1cp;
        a.out[0][r * a.cpo + a.copy_dst0 + i] = xr[logical_slot(a.x1map, 2 * i)];
        a.out[1][r * a.cpo + a.copy_dst0 + i] = xr[logical_slot(a.x1map, 2 * i + 1)];
    }
}

// ======================================================================================== data gradient
template <int D>
__global__ void __launch_bounds__(kGThreads, 1) pwg_dgrad_kernel(const PwBwdArgs a, const int ares) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const PwDesc& d = pw_desc_to_smem(a.d, smem + 520);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    const int KP = d.KP, NP = d.NPall, gwp = d.cols.gwp, S = a.nbuf, Rt = a.Rt, cpo = a.cpo;
    const GBlock gb = a.blk[blockIdx.y];
    const int stw = (gb.n + 7) & ~7;
    const PwgSmem L = pwg_smem(NP, ((NP + 63) & ~63) * 16, ares != 0, stw, S, 16384);
    const int nkb = L.nkb, stage_bytes = L.stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kGMaxStages;
    uint64_t* tm_full = empty + kGMaxStages;
    uint64_t* tm_empty = tm_full + 2;
    uint64_t* a_full = tm_empty + 2;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + 512);
    int* s_ck = reinterpret_cast<int*>(smem + L.ck);
    float4* s_bc = reinterpret_cast<float4*>(smem + L.tab);
    unsigned char* As = smem + L.a;
    unsigned char* ring = smem + L.ring;
    unsigned short* St = reinterpret_cast<unsigned short*>(smem + L.st);

    const int tps = (Rt + kGRows - 1) / kGRows, ntile = kT * tps;
    const int tile_lo = blockIdx.x * a.tiles_per_cta, tile_hi = min(ntile, tile_lo + a.tiles_per_cta);
    const int my_tiles = max(0, tile_hi - tile_lo);
    const int t_first = tile_lo / tps, r_first = (tile_lo - t_first * tps) * kGRows;

    if (warp == 0) tmem_alloc(s_tmem, 2 * kGRows);
    if (tid == 32) {
        for (int s = 0; s < kGMaxStages; ++s) { mbar_init(&full[s], kGProd + (ares ? 0 : 1)); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tm_full[b], 1); mbar_init(&tm_empty[b], kGEpi); }
        mbar_init(a_full, 1);
        mbar_fence_init();
    }
    for (int ck = tid; ck < nkb * 8; ck += kGThreads) {          // 8-column chunk of the GEMM N index -> (plane, slot)
        const int j0 = ck * 8;
        int v = -1;
        if (j0 < NP) { const int p = j0 / gwp; v = (p << 24) | (j0 - p * gwp); }
        s_ck[ck] = v;
    }
    if (ares) { for (int i = tid; i < nkb * 1024; i += kGThreads) reinterpret_cast<uint4*>(As)[i] = make_uint4(0, 0, 0, 0); }
    else { for (int i = tid; i < S * 1024; i += kGThreads) reinterpret_cast<uint4*>(ring + (size_t)(i >> 10) * stage_bytes)[i & 1023] = make_uint4(0, 0, 0, 0); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    if (ares && tid == 32) {
        mbar_expect_tx(a_full, (uint32_t)nkb * gb.n * 128);
        for (int kb = 0; kb < nkb; ++kb) bulk_g2s(As + (size_t)kb * 16384, d.wbs + ((size_t)kb * KP + gb.base) * 64, gb.n * 128, a_full);
    }
    pdl_wait();

    auto advance = [&](GCur& c) { if (++c.kb == nkb) { c.kb = 0; ++c.it; c.r0 += kGRows; if (c.r0 >= Rt) { c.r0 = 0; ++c.t; } } };

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, kGRows, 0, 0);
            if (ares) mbar_wait(a_full, 0);
            int s = 0, ph = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int b = it & 1;
                mbar_wait(&tm_empty[b], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(ring + (size_t)s * stage_bytes);
                    const uint32_t aa = ares ? smem_u32(As) + (uint32_t)kb * 16384u : sb, ba = ares ? sb : sb + 16384u;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_bf16(tmem + (uint32_t)(b * kGRows), umma_desc(aa + ks * 32, 16, 1024), umma_desc(ba + ks * 32, 16, 1024), idesc, (kb | ks) != 0);
                    umma_commit(&empty[s]);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
                umma_commit(&tm_full[b]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && !ares) {
            int s = 0, ph = 0;
            for (int it = 0; it < my_tiles; ++it)
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], (uint32_t)gb.n * 128);
                    bulk_g2s(ring + (size_t)s * stage_bytes, d.wbs + ((size_t)kb * KP + gb.base) * 64, gb.n * 128, &full[s]);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
        }
    } else if (warp < 2 + kGProd) {
        // ================================================================ producers: (d out, out) -> dR -> swizzled K-major tile (+ hand-off)
        const int pt = tid - 64, c = pt & 7, rl = pt >> 3;
        const uint32_t swz = (uint32_t)((c ^ (rl & 7)) << 4);
        const bf16* do0 = a.dout[0]; const bf16* do1 = a.dout[1]; const bf16* o0 = a.out[0]; const bf16* o1 = a.out[1];
        const int nq = my_tiles * nkb;
        const int b_off = L.b_off;
        const bool oclamp = a.out_clamp != 0;
        const bool handoff = a.dr != nullptr && blockIdx.y == 0;
        const double inv_n = 1.0 / (double)Rt;
        int is = 0, iph = 0;
        auto issue = [&](const GCur& cu) {
            mbar_wait(&empty[is], iph ^ 1);
            const int e = s_ck[cu.kb * 8 + c];
            const int pl = e >> 24, col = e & 0xffff;
            const int rows = min(kGRows, Rt - cu.r0);
            const size_t off = ((size_t)cu.t * Rt + cu.r0 + rl) * cpo + (e >= 0 ? col : 0);
            const bf16* dp = (pl > 0 ? do1 : do0) + off; const bf16* op = (pl > 0 ? o1 : o0) + off;
            unsigned char* Bs = ring + (size_t)is * stage_bytes + b_off + swz + rl * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = e >= 0 && rl + 32 * i < rows;
                cp_async16(Bs + i * 32 * 128, ok ? dp + (size_t)(32 * i) * cpo : do0, ok);
                cp_async16(Bs + 16384 + i * 32 * 128, ok ? op + (size_t)(32 * i) * cpo : o0, ok);
            }
            if (++is == S) { is = 0; iph ^= 1; }
        };
        GCur lc{0, 0, t_first, r_first}, pc = lc;
        for (int j = 0; j < D; ++j) { if (j < nq) { issue(lc); advance(lc); } cp_async_commit(); }
        int s = 0, cur_t = -1;
        for (int q = 0; q < nq; ++q) {
            if (q + D < nq) { issue(lc); advance(lc); }
            cp_async_commit();
            if (pc.t != cur_t) {                        // per-slice BatchNorm-backward constants of every GEMM column, [kb][q][c]
                named_bar_sync(2, kGProdThreads);
                for (int j = pt; j < nkb * 64; j += kGProdThreads) {
                    int p, sl, l, nn;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < NP && pw_col(d, j, p, sl, l, nn)) {
                        const size_t idx = (size_t)pc.t * cpo + sl;
                        v = bnbwd_consts(a.tb[p].aff[idx], a.tb[p].bnp[idx], ld_sum(a.tb[p].bsum + idx), inv_n);
                    }
                    s_bc[(j >> 6) * 64 + (j & 7) * 8 + ((j >> 3) & 7)] = v;
                }
                named_bar_sync(2, kGProdThreads);
                cur_t = pc.t;
            }
            const int e = s_ck[pc.kb * 8 + c];
            const int rows = min(kGRows, Rt - pc.r0);
            float4 c8[8];
            {
                const float4* bp = s_bc + pc.kb * 64 + c;
#pragma unroll
                for (int k = 0; k < 8; ++k) c8[k] = bp[k * 8];
            }
            cp_async_wait<D>();
            unsigned char* Bs = ring + (size_t)s * stage_bytes + b_off + swz + rl * 128;
            bf16* drp = handoff ? a.dr + ((size_t)pc.t * Rt + pc.r0 + rl) * NP + pc.kb * 64 + c * 8 : nullptr;
            if (e >= 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (rl + 32 * i < rows) {
                        uint4 v = *reinterpret_cast<const uint4*>(Bs + i * 32 * 128);
                        const uint4 o = *reinterpret_cast<const uint4*>(Bs + 16384 + i * 32 * 128);
                        uint32_t* dw = reinterpret_cast<uint32_t*>(&v); const uint32_t* ow = reinterpret_cast<const uint32_t*>(&o);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 dd = unpack2(dw[k]), oo = unpack2(ow[k]);
                            dw[k] = pack2(bnbwd_apply(dd.x, oo.x, c8[2 * k], oclamp), bnbwd_apply(dd.y, oo.y, c8[2 * k + 1], oclamp));
                        }
                        *reinterpret_cast<uint4*>(Bs + i * 32 * 128) = v;
                        if (drp) *reinterpret_cast<uint4*>(drp + (size_t)(32 * i) * NP) = v;
                    }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            if (++s == S) s = 0;
            advance(pc);
        }
    } else {
        // ================================================================ epilogue: one source slot per thread
        // (the BatchNorm-backward sums of the finished gradient are taken by bsum_kernel right after this launch)
        const int lg = warp & 3, m = 32 * lg + lane, et = tid - 32 * (2 + kGProd);
        const PwSrc& Sx = d.src[gb.aux0];
        const int cp = Sx.cp, slot0 = gb.aux1;
        const bool acc = Sx.accumulate != 0;
        bf16* gout = Sx.grad;
        const int nch = gb.n >> 3, cc = et & 15, crl = et >> 4;
        const bool in_st = m < stw;
        int t = t_first, r0 = r_first;
        for (int it = 0; it < my_tiles; ++it) {
            const int b = it & 1, rows = min(kGRows, Rt - r0);
            const size_t rowbase = (size_t)t * Rt + r0;
            if (et == 0 && it == 5) PWG_TS(12);
            mbar_wait(&tm_full[b], (it >> 1) & 1);
            if (et == 0 && it == 5) PWG_TS(13);
            tc_fence_after();
            named_bar_sync(1, kGEpiThreads);
            const uint32_t taddr = tmem + ((uint32_t)(32 * lg) << 16) + (uint32_t)(b * kGRows);
#pragma unroll 2
            for (int c0 = 0; c0 < kGRows; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
                if (in_st) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) St[(c0 + i) * stw + m] = __bfloat16_as_ushort(__float2bfloat16_rn(v[i]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tm_empty[b]);
            named_bar_sync(1, kGEpiThreads);
            if (et == 0 && it == 5) PWG_TS(14);
            if (cc < nch) {
                bf16* grow = gout + rowbase * cp + slot0 + cc * 8;
                // four rows per round, every load of the round issued before the first use (the existing-share loads are L2 round trips)
                for (int rb = crl; rb < rows; rb += 32) {
                    uint4 v[4], ex[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = rb + 8 * u;
                        v[u] = make_uint4(0, 0, 0, 0); ex[u] = v[u];
                        if (r < rows) {
                            v[u] = *reinterpret_cast<const uint4*>(St + r * stw + cc * 8);
                            if (acc) ex[u] = ldg_cg16(grow + (size_t)r * cp);     // second consumer of the tensor: add to the share already written
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = rb + 8 * u;
                        if (r >= rows) continue;
                        if (acc) {
                            uint32_t* w = reinterpret_cast<uint32_t*>(&v[u]); const uint32_t* ew = reinterpret_cast<const uint32_t*>(&ex[u]);
#pragma unroll
                            for (int k = 0; k < 4; ++k) { const float2 x = unpack2(w[k]), y = unpack2(ew[k]); w[k] = pack2(x.x + y.x, x.y + y.y); }
                        }
                        *reinterpret_cast<uint4*>(grow + (size_t)r * cp) = v[u];
                    }
                }
            }
            r0 += kGRows; if (r0 >= Rt) { r0 = 0; ++t; }
            if (et == 0 && it == 5) PWG_TS(15);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 2 * kGRows);
}

// BatchNorm-backward sums S1 = sum dz, S2 = sum dz * xhat of a freshly written gradient tensor (slots [lo, hi)):
// grid = (row chunks, slices), one thread per slot, 2-byte loads coalesced along the row
struct BsumArgs { const bf16* x; const bf16* dx; int cp, lo, hi, clamp; const float2* aff; const float2* bnp; double2* bsum; int Rt, rows_per_cta; };
// vectorised: thread <-> (8-slot chunk, row lane), 16-byte loads; row lanes are reduced through shared memory.
// Needs lo % 8 == 0 and (hi - lo) % 8 == 0 (every tower tensor: cp is a multiple of 8 and the sums cover [0, cp)).
__global__ void __launch_bounds__(256) bsum_kernel(const BsumArgs a) {
    __shared__ float2 s_part[8 * 256];                 // [row lane][slot] partial (S1, S2); <= 8 row lanes are used
    pdl_trigger();
    const int nsl = a.hi - a.lo, nch = nsl >> 3, t = blockIdx.y, tid = threadIdx.x;
    const int ch = tid % nch, rl = tid / nch, nrl = min(8, 256 / nch);
    const int r_lo = blockIdx.x * a.rows_per_cta, r_hi = min(a.Rt, r_lo + a.rows_per_cta);
    pdl_wait();
    if (rl < nrl) {
        const int s0 = a.lo + ch * 8;
        float4 sc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) sc[q] = sum_consts(a.aff, a.bnp, (size_t)t * a.cp + s0 + q);
        const bool clamp = a.clamp != 0;
        float s1[8], s2[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { s1[q] = 0.f; s2[q] = 0.f; }
        const size_t base = ((size_t)t * a.Rt) * a.cp + s0;
#pragma unroll 4
        for (int r = r_lo + rl; r < r_hi; r += nrl) {
            const uint4 g = ldg_cg16(a.dx + base + (size_t)r * a.cp), x = ldg_cg16(a.x + base + (size_t)r * a.cp);
            const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g); const uint32_t* xw = reinterpret_cast<const uint32_t*>(&x);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 gg = unpack2(gw[i]), xx = unpack2(xw[i]);
                sum_accum(gg.x, xx.x, sc[2 * i], clamp, s1[2 * i], s2[2 * i]);
                sum_accum(gg.y, xx.y, sc[2 * i + 1], clamp, s1[2 * i + 1], s2[2 * i + 1]);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) s_part[rl * nsl + ch * 8 + q] = make_float2(s1[q], s2[q]);
    }
    __syncthreads();
    for (int i = tid; i < nsl; i += 256) {
        float v1 = 0.f, v2 = 0.f;
        for (int l = 0; l < nrl; ++l) { const float2 v = s_part[l * nsl + i]; v1 += v.x; v2 += v.y; }
        if (v1 != 0.f || v2 != 0.f) {
            double2* dst = a.bsum + (size_t)t * a.cp + a.lo + i;
            atomicAdd(&dst->x, (double)v1); atomicAdd(&dst->y, (double)v2);
        }
    }
}

// ======================================================================================== weight gradient
// dW[kk][j] += sum_r act(src)[r][kk] * dR[r][j] as a split-K tcgen05 GEMM: an output tile is (M block of <= 128 source
// slots) x (N block of <= 256 GEMM columns), blockIdx.y; the rows are split over blockIdx.x.  Both operands are
// MN-major views of row tiles: the same [128 rows][64 channels] swizzled blocks the forward uses as K-major tiles
// (v2_umma.cuh), so the producers are the forward's (cp.async + in-place BatchNorm affine / ReLU6 for act(src)) plus a
// plain copy of the dR hand-off matrix the data-gradient kernel left.  The accumulator stays in TMEM for the CTA's life;
// the epilogue transposes it through shared memory and adds it to the fp32 gradient arena with coalesced atomics.
constexpr int kWgRows = 64;            // rows per ring stage of the weight gradient (4 tcgen05.mma K steps): small stages, so that the
                                       // ring is deep enough for the producers never to wait for an MMA round trip
struct PwgWgSmem { int ck, tab, maps, ring, total, stage_bytes; };
inline __host__ __device__ PwgWgSmem pwg_wg_smem(int KP, int nbk, int nstage) {
    PwgWgSmem s;
    int off = 1024;
    s.ck = off; off += 128 * 4;
    s.tab = off; off += ((KP + 63) & ~63) * 8;
    s.maps = off; off += (128 + 256) * 4;
    off = (off + 1023) & ~1023;
    s.stage_bytes = (2 + nbk) * kWgRows * 128;
    s.ring = off; off += nstage * s.stage_bytes;
    s.total = off + 1024;
    return s;
}
struct PwgWgArgs {
    const PwDesc* d; int Rt; const bf16* dr;             // dR hand-off [4*Rt][NPall]
    GBlock blk[kGMaxBlk]; int nblk;                      // M blocks: first GEMM row, rows, source, first slot
    int nnb;                                             // N blocks of 256 GEMM columns; blockIdx.y = mblock * nnb + nblock
    int tiles_per_cta, nbuf;
    Tables tb[2]; int cpo;                               // for the BatchNorm parameter gradients (dgamma, dbeta)
};

template <int D>
__global__ void __launch_bounds__(kGThreads, 1) pwg_wgrad_kernel(const PwgWgArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const PwDesc& d = pw_desc_to_smem(a.d, smem + 520);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    const int KP = d.KP, NP = d.NPall, nsrc = d.nsrc, S = a.nbuf, Rt = a.Rt;
    const int mbi = blockIdx.y / a.nnb, nbi = blockIdx.y - mbi * a.nnb;
    const GBlock gb = a.blk[mbi];
    const int j_lo = nbi * 256, ncol = min(256, NP - j_lo), nbk = (ncol + 63) >> 6, npad = nbk * 64;
    const PwgWgSmem L = pwg_wg_smem(KP, (min(256, NP) + 63) >> 6, S);
    const int stage_bytes = (2 + ((min(256, NP) + 63) >> 6)) * kWgRows * 128;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kGMaxStages;
    uint64_t* acc_full = empty + kGMaxStages;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + 512);
    int* s_ck = reinterpret_cast<int*>(smem + L.ck);
    float2* s_aff = reinterpret_cast<float2*>(smem + L.tab);
    int* s_rmap = reinterpret_cast<int*>(smem + L.maps);
    int* s_cmap = s_rmap + 128;
    unsigned char* ring = smem + L.ring;

    const int tps = (Rt + kWgRows - 1) / kWgRows, ntile = kT * tps;
    const int tile_lo = blockIdx.x * a.tiles_per_cta, tile_hi = min(ntile, tile_lo + a.tiles_per_cta);
    const int my_tiles = max(0, tile_hi - tile_lo);
    const int t_first = tile_lo / tps, r_first = (tile_lo - t_first * tps) * kWgRows;

    if (warp == 0) tmem_alloc(s_tmem, 256);
    if (tid == 32) {
        for (int s = 0; s < kGMaxStages; ++s) { mbar_init(&full[s], kGProd); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    const int nkc = ((KP + 63) & ~63) >> 3;
    for (int ck = tid; ck < nkc; ck += kGThreads) {
        int v = -1, off = 0;
        for (int i = 0; i < nsrc; ++i) {
            const int nch = d.src[i].cp >> 3;
            if (ck >= off && ck < off + nch) v = (i << 24) | ((d.src[i].clamp ? 1 : 0) << 23) | ((ck - off) * 8);
            off += nch;
        }
        s_ck[ck] = v;
    }
    for (int i = tid; i < 128 + 256; i += kGThreads) {
        int l, k, p, sl, nn, v = -1;
        if (i < 128) { if (i < gb.n && pw_row(d, gb.base + i, l, k)) v = (l << 24) | k; }
        else if (i - 128 < ncol && pw_col(d, j_lo + i - 128, p, sl, l, nn)) v = (l << 24) | nn;
        s_rmap[i] = v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, npad, 1, 1);
            int s = 0, ph = 0;
            for (int it = 0; it < my_tiles; ++it) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes), sb = sa + 2u * kWgRows * 128u;
#pragma unroll
                for (int ks = 0; ks < kWgRows / 16; ++ks)
                    umma_bf16(tmem, umma_desc(sa + ks * 2048, kWgRows * 128, 1024), umma_desc(sb + ks * 2048, kWgRows * 128, 1024), idesc, (it | ks) != 0);
                umma_commit(&empty[s]);
                if (++s == S) { s = 0; ph ^= 1; }
            }
            umma_commit(acc_full);
        }
    } else if (warp >= 2 && warp < 2 + kGProd) {
        const int pt = tid - 64, c = pt & 7, rl = pt >> 3;
        const uint32_t swz = (uint32_t)((c ^ (rl & 7)) << 4);
        const bf16* sp0 = d.src[0].data; const bf16* sp1 = nsrc > 1 ? d.src[1].data : nullptr; const bf16* sp2 = nsrc > 2 ? d.src[2].data : nullptr;
        const int cp0 = d.src[0].cp, cp1 = nsrc > 1 ? d.src[1].cp : 0, cp2 = nsrc > 2 ? d.src[2].cp : 0;
        // this thread's two act chunks (one per 64-channel block of the M block): table entries, validity
        int ea[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) ea[b] = (b * 64 + c * 8 < gb.n) ? s_ck[(gb.base >> 3) + b * 8 + c] : -1;
        int is = 0, iph = 0;
        int lt = t_first, lr0 = r_first;                // issue cursor
        auto issue = [&]() {
            mbar_wait(&empty[is], iph ^ 1);
            const int rows = min(kWgRows, Rt - lr0);
            unsigned char* st = ring + (size_t)is * stage_bytes + swz + rl * 128;
            const size_t row = (size_t)lt * Rt + lr0 + rl;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int e = ea[b], si = e >> 24;
                const bf16* sp = si <= 0 ? sp0 : (si == 1 ? sp1 : sp2);
                const int cp = si <= 0 ? cp0 : (si == 1 ? cp1 : cp2);
                const bf16* base = sp + row * cp + (e >= 0 ? (e & 0xffff) : 0);
#pragma unroll
                for (int i = 0; i < kWgRows / 32; ++i) {
                    const bool ok = e >= 0 && rl + 32 * i < rows;
                    cp_async16(st + b * (kWgRows * 128) + i * 32 * 128, ok ? base + (size_t)(32 * i) * cp : sp0, ok);
                }
            }
            for (int b = 0; b < nbk; ++b) {
                const int j0 = j_lo + b * 64 + c * 8;
                const bf16* base = a.dr + row * NP + j0;
#pragma unroll
                for (int i = 0; i < kWgRows / 32; ++i) {
                    const bool ok = j0 < NP && rl + 32 * i < rows;
                    cp_async16(st + (2 + b) * (kWgRows * 128) + i * 32 * 128, ok ? base + (size_t)(32 * i) * NP : a.dr, ok);
                }
            }
            if (++is == S) { is = 0; iph ^= 1; }
            lr0 += kWgRows; if (lr0 >= Rt) { lr0 = 0; ++lt; }
        };
        for (int j = 0; j < D; ++j) { if (j < my_tiles) issue(); cp_async_commit(); }
        int s = 0, cur_t = -1, t = t_first, r0 = r_first;
        for (int q = 0; q < my_tiles; ++q) {
            if (q + D < my_tiles) issue();
            cp_async_commit();
            if (t != cur_t) {
                named_bar_sync(2, kGProdThreads);
                for (int kk = pt; kk < nkc * 8; kk += kGProdThreads) {
                    const int e = s_ck[kk >> 3];
                    float2 v = make_float2(1.f, 0.f);
                    if (e >= 0) {
                        const PwSrc& Sx = d.src[e >> 24];
                        if (Sx.aff) v = Sx.aff[(size_t)t * Sx.cp + (e & 0xffff) + (kk & 7)];
                    }
                    s_aff[kk] = v;
                }
                named_bar_sync(2, kGProdThreads);
                cur_t = t;
            }
            const int rows = min(kWgRows, Rt - r0);
            cp_async_wait<D>();
            unsigned char* st = ring + (size_t)s * stage_bytes + swz + rl * 128;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int e = ea[b];
                if (e < 0) continue;
                float2 c8[8];
                const float4* ap = reinterpret_cast<const float4*>(s_aff + gb.base + b * 64 + c * 8);
                const float4 a0 = ap[0], a1 = ap[1], a2 = ap[2], a3 = ap[3];
                c8[0] = make_float2(a0.x, a0.y); c8[1] = make_float2(a0.z, a0.w); c8[2] = make_float2(a1.x, a1.y); c8[3] = make_float2(a1.z, a1.w);
                c8[4] = make_float2(a2.x, a2.y); c8[5] = make_float2(a2.z, a2.w); c8[6] = make_float2(a3.x, a3.y); c8[7] = make_float2(a3.z, a3.w);
                const bool clamp = ((e >> 23) & 1) != 0;
#pragma unroll
                for (int i = 0; i < kWgRows / 32; ++i)
                    if (rl + 32 * i < rows) {
                        uint4* bp = reinterpret_cast<uint4*>(st + b * (kWgRows * 128) + i * 32 * 128);
                        *bp = affine8(*bp, c8, clamp);
                    }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            if (++s == S) s = 0;
            r0 += kWgRows; if (r0 >= Rt) { r0 = 0; ++t; }
        }
    }
    // ================================================================ all roles: accumulator -> gradient arena
    __syncthreads();
    if (my_tiles > 0) {
        mbar_wait(acc_full, 0);
        tc_fence_after();
        float* Sc = reinterpret_cast<float*>(ring);        // [128][65]
        float* wl0 = d.layer[0].dw; float* wl1 = d.layer[1].dw;
        const int N0 = d.layer[0].N, N1 = d.layer[1].N;
        for (int c0 = 0; c0 < npad; c0 += 64) {
            if (warp < 4) {
#pragma unroll
                for (int cq = 0; cq < 64; cq += 16) {
                    float v[16];
                    tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(c0 + cq), v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) Sc[(32 * warp + lane) * 65 + cq + i] = v[i];
                }
            }
            __syncthreads();
            {
                const int col = tid & 63, cm = (c0 + col < ncol) ? s_cmap[c0 + col] : -1;
                if (cm >= 0)
                    for (int rowk = tid >> 6; rowk < 128; rowk += kGThreads >> 6) {
                        const int rm = s_rmap[rowk];
                        if (rm >= 0 && (rm >> 24) == (cm >> 24)) {
                            const float val = Sc[rowk * 65 + col];
                            if (val != 0.f) {
                                if ((rm >> 24) == 0) atomicAdd(wl0 + (size_t)(rm & 0xffffff) * N0 + (cm & 0xffffff), val);
                                else atomicAdd(wl1 + (size_t)(rm & 0xffffff) * N1 + (cm & 0xffffff), val);
                            }
                        }
                    }
            }
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
    // ---- BatchNorm parameter gradients (one CTA): dgamma = sum_t S2, dbeta = sum_t S1
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        for (int j = tid; j < NP; j += kGThreads) {
            int p, sl, l, nn;
            if (!pw_col(d, j, p, sl, l, nn)) continue;
            double gs = 0.0, bs = 0.0;
            double2 v[kT];                               // all four loads in flight before the sums (block 0 ends the kernel)
#pragma unroll
            for (int t = 0; t < kT; ++t) v[t] = ld_sum(a.tb[p].bsum + (size_t)t * a.cpo + sl);
#pragma unroll
            for (int t = 0; t < kT; ++t) { bs += v[t].x; gs += v[t].y; }
            d.layer[l].dg[nn] = (float)gs; d.layer[l].dbe[nn] = (float)bs;
        }
    }
}

// pass-through half of a stride-1 unit, backward: d x1[slot(2i + p)] = d out_p[n0p + i] (bit-exact gather) and the
// BatchNorm-backward sums of x1.  grid = (row chunks, slices), one thread per x1 slot.
struct PassBwdArgs {
    const bf16* x1; bf16* dx1; int x1cp; SlotMap x1map; const float2* x1aff; const float2* x1bnp; double2* x1bsum; int x1clamp;
    const bf16* dout[2]; int cpo, ncopy, copy_dst0, Rt, rows_per_cta;
};
__global__ void __launch_bounds__(256) pass_bwd_kernel(const PassBwdArgs a) {
    pdl_trigger();
    const int s = threadIdx.x, t = blockIdx.y;
    const int r_lo = blockIdx.x * a.rows_per_cta, r_hi = min(a.Rt, r_lo + a.rows_per_cta);
    pdl_wait();
    if (s >= a.x1cp) return;
    const int l = slot_logical(a.x1map, s);
    const bf16* src = (l >= 0 && (l >> 1) < a.ncopy) ? a.dout[l & 1] + a.copy_dst0 + (l >> 1) : nullptr;
    const float4 sc = sum_consts(a.x1aff, a.x1bnp, (size_t)t * a.x1cp + s);
    const bool clamp = a.x1clamp != 0;
    float s1 = 0.f, s2 = 0.f;
    // __restrict__ + read-only loads: without them every row's loads wait behind the previous row's store (possible aliasing) and the
    // column walk runs at one memory round trip (0.6 us) per row
    const unsigned short* __restrict__ xr = reinterpret_cast<const unsigned short*>(a.x1) + (size_t)t * a.Rt * a.x1cp + s;
    unsigned short* __restrict__ dx = reinterpret_cast<unsigned short*>(a.dx1) + (size_t)t * a.Rt * a.x1cp + s;
    const unsigned short* __restrict__ sp = src ? reinterpret_cast<const unsigned short*>(src) + (size_t)t * a.Rt * a.cpo : xr;
    const int sps = src ? a.cpo : a.x1cp;               // (no source: read x1 again and discard -> no per-row branch on the pointer)
    const bool has = src != nullptr;
    for (int rb = r_lo; rb < r_hi; rb += 8) {
        unsigned short g[8], x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = min(rb + u, r_hi - 1);
            g[u] = __ldg(sp + (size_t)r * sps); x[u] = __ldg(xr + (size_t)r * a.x1cp);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = rb + u;
            if (r < r_hi) {
                const unsigned short gv = has ? g[u] : (unsigned short)0;
                dx[(size_t)r * a.x1cp] = gv;
                sum_accum(__uint_as_float((uint32_t)gv << 16), __uint_as_float((uint32_t)x[u] << 16), sc, clamp, s1, s2);
            }
        }
    }
    if (a.x1bsum && (s1 != 0.f || s2 != 0.f)) {
        double2* dst = a.x1bsum + (size_t)t * a.x1cp + s;
        atomicAdd(&dst->x, (double)s1); atomicAdd(&dst->y, (double)s2);
    }
}

}  // namespace v2
}  // namespace cdra
#endif
