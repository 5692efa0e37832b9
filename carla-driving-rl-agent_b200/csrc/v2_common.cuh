// "v2" image tower (bf16 perf mode): shared device helpers.
//
// Layout: every tower activation is a bf16 matrix [4*Rt rows][cp slots], rows ordered
// [slice t][sample b][y][x], cp a multiple of 8 so that rows are 16-byte aligned and whole row tiles are
// CONTIGUOUS in HBM (one TMA bulk copy per tile).  A unit output (core/architectures.py:120-145) is stored as
// two such matrices ("planes"): plane g holds the logical channels out[g*C/2 + i] = concat[2i + g]
// (channel_shuffle, :109-118).  Inside a plane the physical slot order is
//     [ part0: branch (pw2) outputs, n0 valid of n0p | part1: shortcut / pass-through, n1 valid | zero pad ]
// while the plane's LOGICAL channel order (what the next layer's weights are indexed by) is [part1 | part0].
// Plain tensors (pw1 / dw outputs, pool output, head conv) are the special case n1 = 0.
#pragma once
#ifndef CDRA_EMU
#include "cdra_common.cuh"

namespace cdra {
namespace v2 {

struct SlotMap { int n0, n0p, n1; };
CDRA_DEV int slot_logical(const SlotMap& m, int s) {        // logical channel of slot s, -1 = padding
    if (s < m.n0p) return s < m.n0 ? m.n1 + s : -1;
    s -= m.n0p;
    return s < m.n1 ? s : -1;
}
CDRA_DEV int logical_slot(const SlotMap& m, int l) { return l < m.n1 ? m.n0p + l : l - m.n1; }

// shared-memory row stride (elements) of a bf16 tile with n used columns: covers the 16-column MMA k steps and is an
// ODD multiple of 16 bytes, so the 8 rows of an ldmatrix / of an accumulator store fall into 8 different bank groups
// (n + 8 is not enough: 120 + 8 = 128 elements = 256 bytes puts every row on the same banks)
inline __host__ __device__ int pad_ld(int n) { return ((n + 15) & ~15) + 8; }

// ---------------------------------------------------------------------------------------------- PTX wrappers
CDRA_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
CDRA_DEV void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
CDRA_DEV void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
CDRA_DEV void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
CDRA_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
CDRA_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
CDRA_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared (contiguous bytes; 16-byte aligned, size multiple of 16); completion on `bar`
CDRA_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// programmatic dependent launch (see CDRA_LAUNCH_PDL): let the next kernel in the stream start its prologue / block until
// the previous kernel's results are complete and visible
CDRA_DEV void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
CDRA_DEV void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

CDRA_DEV uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
CDRA_DEV float2 unpack2(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
CDRA_DEV void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
CDRA_DEV void ldsm4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
CDRA_DEV void ldsm4t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
CDRA_DEV uint4 ldg_cg16(const void* p) {        // L2-only 16-byte load (data written by other CTAs / streams of tiles)
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// relu6(scale * x + shift) on 8 packed bf16: fp32 FMA, one rounding, then the clamp on the packed pair (0 and 6 are
// exact in bf16 and rounding is monotonic, so clamping after the rounding gives the same bits as clamping before)
CDRA_DEV uint32_t relu6_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    v = __hmin2(__hmax2(v, __float2bfloat162_rn(0.f)), __float2bfloat162_rn(6.f));
    return *reinterpret_cast<uint32_t*>(&v);
}
CDRA_DEV uint4 affine8(uint4 v, const float2 (&c)[8], bool clamp) {
    uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = unpack2(w[i]);
        w[i] = pack2(fmaf(f.x, c[2 * i].x, c[2 * i].y), fmaf(f.y, c[2 * i + 1].x, c[2 * i + 1].y));
        if (clamp) w[i] = relu6_bf16x2(w[i]);
    }
    return v;
}

// Per-column constant tables in shared memory are stored "chunk-transposed": column c of a table with nch 8-column
// chunks lives at (c & 7) * nch + (c >> 3), so that the q-th constants of the chunks a warp works on are contiguous
// (one wavefront per load instead of an 8-way bank conflict).
CDRA_DEV int tcol(int c, int nch) { return (c & 7) * nch + (c >> 3); }

// "last CTA done" ticket with a single fencing thread (release: bar.sync orders the CTA's atomics before the fence)
CDRA_DEV bool last_cta(unsigned* counter, unsigned total) {
    __shared__ unsigned s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(counter, 1u);
        s_last = (t == total - 1) ? 1u : 0u;
        if (t == total - 1) *counter = 0;
        __threadfence();
    }
    __syncthreads();
    return s_last != 0;
}
CDRA_DEV double2 ld_sum(const double2* p) {      // sums written by atomics of other CTAs: read through L2
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// BatchNorm tables of one stored tensor (all [kT][cp])
struct Tables {
    double2* fsum;      // forward sums (sum x, sum x^2) over the stored (rounded) raw values
    double2* bsum;      // backward sums (sum dz, sum dz*xhat)
    float2* aff;        // (scale, shift)
    float2* bnp;        // (mean, inv_std)
};

// One BatchNorm-ed conv layer (parameter / state arena offsets resolved to pointers)
struct LayerP {
    const float* w; const float* b; const float* g; const float* be;   // trainable
    float* mm; float* mv;                                               // moving statistics (may be null)
    float* dw; float* db; float* dg; float* dbe;                        // gradients (backward only)
    int K, N;
};

// scale/shift/mean/inv for one (slice, channel) from batch sums or (inference) moving statistics, plus the
// kT sequential Keras moving-average updates (core/architectures.py:44-57 applies the same layer object to every
// time slice; FusedBatchNorm feeds the unbiased variance to the moving average).
struct BnFin { float scale, shift, mean, inv; };
CDRA_DEV BnFin bn_from_moving(float mm, float mv, float g, float b) {
    BnFin r;
    r.inv = 1.0f / sqrtf(mv + kBnEps);
    r.mean = mm;
    r.scale = g * r.inv;
    r.shift = b - mm * r.scale;
    return r;
}
// finalise one output channel: tables for the 4 slices + moving statistics
CDRA_DEV void bn_finalize_channel(const Tables& tb, int cp, int slot, const LayerP& L, int n_logical, double n, int training) {
    const float g = L.g[n_logical], b = L.be[n_logical];
    float mm = L.mm ? L.mm[n_logical] : 0.f, mv = L.mv ? L.mv[n_logical] : 1.f;
    const float mm0 = mm, mv0 = mv;
    // this runs in the LAST CTA while the rest of the GPU idles: all four slices' sums are fetched before any arithmetic, so
    // the L2 round trips overlap instead of alternating with the double-precision math
    const double inv_n = 1.0 / n, unbias = n / (n > 1.0 ? n - 1.0 : 1.0);     // FusedBatchNorm feeds the unbiased variance to the moving average
    double2 s[kT];
#pragma unroll
    for (int t = 0; t < kT; ++t) s[t] = training ? ld_sum(tb.fsum + (size_t)t * cp + slot) : make_double2(0.0, 0.0);
#pragma unroll
    for (int t = 0; t < kT; ++t) {
        BnFin f;
        if (training) {
            const double mean = s[t].x * inv_n;
            double var = s[t].y * inv_n - mean * mean;
            if (var < 0.0) var = 0.0;
            f.inv = (float)rsqrt(var + (double)kBnEps);
            f.mean = (float)mean;
            f.scale = g * f.inv;
            f.shift = b - f.mean * f.scale;
            mm -= (mm - (float)mean) * (1.f - kBnMomentum);
            mv -= (mv - (float)(var * unbias)) * (1.f - kBnMomentum);
        } else {
            f = bn_from_moving(mm0, mv0, g, b);
        }
        tb.aff[(size_t)t * cp + slot] = make_float2(f.scale, f.shift);
        tb.bnp[(size_t)t * cp + slot] = make_float2(f.mean, f.inv);
    }
    if (training && L.mm) { L.mm[n_logical] = mm; L.mv[n_logical] = mv; }
}

}  // namespace v2
}  // namespace cdra
#endif
