// v2 tower: 5th-generation tensor core (tcgen05 / UMMA) primitives for sm_100a, written against the PTX ISA directly.
//
// Operand tiles live in shared memory in the canonical 128-byte-swizzled layout: a tile of `rows` x `cols` bf16 is cut
// into column blocks of 64 elements; block b is a dense [rows][128 bytes] matrix (1024-byte aligned) whose 16-byte
// chunk c of row r sits at chunk position c ^ (r & 7).  The same bytes serve as
//   * a K-major operand    (rows = M or N index, the 64 columns = K):  8-row groups 1024 bytes apart (SBO)
//   * an MN-major operand  (rows = K index, the 64 columns = M or N):  column blocks LBO bytes apart, 8-row K groups
//                                                                        1024 bytes apart (SBO)
// which is what lets one staged gradient tile feed both the data-gradient and the weight-gradient product.
// Accumulators live in tensor memory (TMEM, 128 lanes x 512 32-bit columns per SM): lane = accumulator row, column =
// accumulator column; tcgen05.ld brings them back to registers (warp w of a warpgroup reads lanes 32w .. 32w+31).
#pragma once
#ifndef CDRA_EMU
#include "v2_common.cuh"

namespace cdra {
namespace v2 {

constexpr int kUmmaBlockCols = 64;              // bf16 columns per 128-byte swizzled row

// byte offset of element (r, col) inside a swizzled tile with `rows` rows (col blocks of 64, 16-byte chunk granularity)
CDRA_DEV uint32_t sw128_offset(int r, int col, int rows) {
    const int blk = col >> 6, cc = (col >> 3) & 7;
    return (uint32_t)blk * (uint32_t)rows * 128u + (uint32_t)r * 128u + (uint32_t)((cc ^ (r & 7)) << 4) + (uint32_t)(col & 7) * 2u;
}

// shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor"): start address, leading / stride byte offsets
// (all >> 4), version 1, 128-byte swizzle
CDRA_DEV uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}
// instruction descriptor of tcgen05.mma.kind::f16: bf16 x bf16 -> fp32, M x N tile, operand majors (0 = K, 1 = MN)
CDRA_DEV uint32_t umma_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
CDRA_DEV void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when they have completed (implies fence::before_thread_sync)
CDRA_DEV void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
CDRA_DEV void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
CDRA_DEV void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (tensor core operand fetch)
CDRA_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMEM allocation (one full warp; column count a power of two >= 32); the base address lands in *slot (shared memory)
CDRA_DEV void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
CDRA_DEV void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 8 consecutive accumulator columns of this thread's lane (warp-collective; lane base = 32 * (warp % 4) in bits 16..31)
CDRA_DEV void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive accumulator columns (one tcgen05.wait::ld per 16 values)
CDRA_DEV void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------- self test
// C[Mw][Nw] (fp32) = X^T Y for row-major bf16 X [rows][Mw], Y [rows][Nw]: both operands MN-major, 64-row tiles accumulated
// in TMEM (the weight-gradient shape).  Mw in {128, 256}, Nw % 16 == 0, Nw <= 256, Mw/128 * Nw <= 512, rows % 64 == 0.
struct UmmaTestArgs { const bf16* X; const bf16* Y; float* C; int rows, Mw, Nw; };
__global__ void __launch_bounds__(256) umma_selftest_kernel(const UmmaTestArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t mma_done;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int R = 64;
    const int mblk = a.Mw >> 6, nblk = (a.Nw + 63) >> 6;
    unsigned char* Xs = smem;                                   // mblk blocks of [64][128 B]
    unsigned char* Ys = smem + (size_t)mblk * R * 128;          // nblk blocks
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 0) { mbar_init(&mma_done, 1); mbar_fence_init(); }
    for (int i = tid; i < (mblk + nblk) * R * 128 / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc(128, a.Nw, 1, 1);
    const int ntile = a.rows / R;
    for (int tile = 0; tile < ntile; ++tile) {
        if (tile > 0) { mbar_wait(&mma_done, (tile - 1) & 1); tc_fence_after(); }      // the previous tile's MMAs have read the staging tiles
        for (int i = tid; i < R * (a.Mw >> 3); i += 256) {
            const int r = i / (a.Mw >> 3), c = i - r * (a.Mw >> 3);
            *reinterpret_cast<uint4*>(Xs + sw128_offset(r, c * 8, R)) = *reinterpret_cast<const uint4*>(a.X + ((size_t)(tile * R + r)) * a.Mw + c * 8);
        }
        for (int i = tid; i < R * (a.Nw >> 3); i += 256) {
            const int r = i / (a.Nw >> 3), c = i - r * (a.Nw >> 3);
            *reinterpret_cast<uint4*>(Ys + sw128_offset(r, c * 8, R)) = *reinterpret_cast<const uint4*>(a.Y + ((size_t)(tile * R + r)) * a.Nw + c * 8);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int mb = 0; mb < (a.Mw >> 7); ++mb)
                for (int ks = 0; ks < R / 16; ++ks) {
                    const uint64_t ad = umma_desc(smem_u32(Xs) + mb * 2 * R * 128 + ks * 2048, R * 128, 1024);
                    const uint64_t bd = umma_desc(smem_u32(Ys) + ks * 2048, R * 128, 1024);
                    umma_bf16(tmem + mb * a.Nw, ad, bd, idesc, tile > 0 || ks > 0);
                }
            umma_commit(&mma_done);
        }
    }
    mbar_wait(&mma_done, (ntile - 1) & 1);
    tc_fence_after();
    if (warp < 4) {
        for (int mb = 0; mb < (a.Mw >> 7); ++mb)
            for (int c0 = 0; c0 < a.Nw; c0 += 8) {
                float v[8];
                tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + mb * a.Nw + c0, v);
                float* dst = a.C + (size_t)(mb * 128 + 32 * warp + lane) * a.Nw + c0;
#pragma unroll
                for (int i = 0; i < 8; ++i) dst[i] = v[i];
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// C[Mw][Nw] (fp32) = A B^T for row-major bf16 A [Mw][Kw], B [Nw][Kw]: both operands K-major (the forward / data-gradient
// shape): per 64-column K block a [rows][128 B] swizzled tile; 16-element K steps advance the descriptor start by 32 bytes.
struct UmmaTestKArgs { const bf16* A; const bf16* B; float* C; int Mw, Nw, Kw; };
__global__ void __launch_bounds__(256) umma_selftest_k_kernel(const UmmaTestKArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t mma_done;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = a.Kw >> 6, nmb = a.Mw >> 7;
    unsigned char* As = smem;                                          // [nmb][nkb] blocks of [128][128 B]
    unsigned char* Bs = smem + (size_t)nmb * nkb * 128 * 128;           // [nkb] blocks of [Nw][128 B]
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 0) { mbar_init(&mma_done, 1); mbar_fence_init(); }
    for (int i = tid; i < a.Mw * (a.Kw >> 3); i += 256) {
        const int r = i / (a.Kw >> 3), c = i - r * (a.Kw >> 3), mb = r >> 7, kb = c >> 3;
        *reinterpret_cast<uint4*>(As + (size_t)(mb * nkb + kb) * 128 * 128 + sw128_offset(r & 127, (c & 7) * 8, 128)) =
            *reinterpret_cast<const uint4*>(a.A + (size_t)r * a.Kw + c * 8);
    }
    for (int i = tid; i < a.Nw * (a.Kw >> 3); i += 256) {
        const int r = i / (a.Kw >> 3), c = i - r * (a.Kw >> 3), kb = c >> 3;
        *reinterpret_cast<uint4*>(Bs + (size_t)kb * a.Nw * 128 + sw128_offset(r, (c & 7) * 8, a.Nw)) =
            *reinterpret_cast<const uint4*>(a.B + (size_t)r * a.Kw + c * 8);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc(128, a.Nw, 0, 0);
        for (int mb = 0; mb < nmb; ++mb)
            for (int kb = 0; kb < nkb; ++kb)
                for (int ks = 0; ks < 4; ++ks)
                    umma_bf16(tmem + mb * a.Nw, umma_desc(smem_u32(As) + (mb * nkb + kb) * 128 * 128 + ks * 32, 16, 1024),
                              umma_desc(smem_u32(Bs) + kb * a.Nw * 128 + ks * 32, 16, 1024), idesc, kb > 0 || ks > 0);
        umma_commit(&mma_done);
    }
    mbar_wait(&mma_done, 0);
    tc_fence_after();
    if (warp < 4)
        for (int mb = 0; mb < nmb; ++mb)
            for (int c0 = 0; c0 < a.Nw; c0 += 8) {
                float v[8];
                tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + mb * a.Nw + c0, v);
                float* dst = a.C + (size_t)(mb * 128 + 32 * warp + lane) * a.Nw + c0;
#pragma unroll
                for (int i = 0; i < 8; ++i) dst[i] = v[i];
            }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace v2
}  // namespace cdra
#endif
