// Common device helpers for libcdra (sm_100a).  Compiled by nvcc for the product library; the same
// sources are compiled by g++ with -DCDRA_EMU against tests/emu/cuda_emu.h for CPU logic tests.
#pragma once
#include <cstdint>
#include <cstddef>
#ifdef CDRA_EMU
#include "cuda_emu.h"
#include <cuda_bf16.h>
#define CDRA_KERNEL static void
#define CDRA_DEV static inline
#define CDRA_SHARED static
#define CDRA_LAUNCH_BOUNDS(n)
#define CDRA_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(grid, block, smem, [=]() { kernel(__VA_ARGS__); })
#define CDRA_DYN_SMEM(name) char* name = emu::S().dyn_smem
#define CDRA_RESTRICT
#else
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#define CDRA_KERNEL __global__ void
#define CDRA_DEV __device__ __forceinline__
#define CDRA_SHARED __shared__
#define CDRA_LAUNCH_BOUNDS(n) __launch_bounds__(n)
namespace cdra {
// launch accounting / optional per-kernel CUDA-event timing (cdra_profile_* in include/cdra.h)
void prof_pre(const void* func, cudaStream_t stream);
void prof_post(cudaStream_t stream);
}
#define CDRA_LAUNCH(kernel, grid, block, smem, stream, ...)          \
    do {                                                              \
        cdra::prof_pre((const void*)(kernel), stream);                \
        kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);           \
        cdra::prof_post(stream);                                      \
    } while (0)
// Programmatic dependent launch: the kernel may start while its predecessor in the stream is still draining; it runs
// its private prologue (shared-memory setup, weights that were final long before), then pdl_wait() blocks until the
// predecessor grid has completed and its memory is visible.  Kernels launched this way must not read upstream-produced
// data or write global memory before pdl_wait().  CDRA_NO_PDL=1 falls back to plain stream order.
namespace cdra {
bool pdl_enabled();
template <typename K, typename... Args>
inline void launch_pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}
}
#define CDRA_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                    \
    do {                                                                            \
        cdra::prof_pre((const void*)(kernel), stream);                              \
        cdra::launch_pdl(kernel, grid, block, smem, stream, __VA_ARGS__);           \
        cdra::prof_post(stream);                                                    \
    } while (0)
#define CDRA_DYN_SMEM(name) extern __shared__ __align__(1024) char name[]
#define CDRA_RESTRICT __restrict__
#endif
namespace cdra { void prof_bytes(double algorithmic_bytes); }   // bytes of the next launch (for the roofline)

typedef __nv_bfloat16 bf16;

namespace cdra {

constexpr int kT = 4;                 // time slices (env.time_horizon of the reference, core/carla_env.py:26)
constexpr float kBnEps = 1e-3f;       // Keras BatchNormalization default epsilon
constexpr float kBnMomentum = 0.99f;  // Keras default momentum

// ---------------------------------------------------------------- element access
CDRA_DEV float ldf(const float* p) { return *p; }
CDRA_DEV float ldf(const bf16* p) { return __bfloat162float(*p); }
CDRA_DEV void stf(float* p, float v) { *p = v; }
CDRA_DEV void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
// value after a round trip through the storage type (statistics are taken over what is stored)
CDRA_DEV float rnd(float v, const float*) { return v; }
CDRA_DEV float rnd(float v, const bf16*) { return __bfloat162float(__float2bfloat16_rn(v)); }

// two adjacent channels in one access (element index must be even)
CDRA_DEV float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
CDRA_DEV float2 ld2(const bf16* p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); }
CDRA_DEV void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
CDRA_DEV void st2(bf16* p, float2 v) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y); }

CDRA_DEV float relu6f(float v) { return fminf(fmaxf(v, 0.f), 6.f); }

// A channel sub-range of an NHWC activation tensor whose values are stored *raw* (pre-BatchNorm);
// consumers apply the producer's per-(slice, channel) affine (+ ReLU6) on load.
struct ActView {
    const void* data;     // element type T; points at channel 0 of the underlying tensor
    int ld;               // channels of the underlying tensor (= pixel stride in elements)
    int coff;             // first channel of the view
    const float2* aff;    // [kT][ld] (scale, shift); nullptr = identity
    int clamp;            // 1 = ReLU6 after the affine
};

CDRA_DEV float act_apply(float raw, const float2* aff_tc, int clamp) {
    float v = raw;
    if (aff_tc) { float2 a = *aff_tc; v = fmaf(raw, a.x, a.y); }
    if (clamp) v = relu6f(v);
    return v;
}

// BatchNorm bookkeeping tables that live next to every raw tensor (all indexed [kT][C]).
struct BnTables {
    double2* fst;     // forward sums (sum x, sum x^2)
    float2* aff;      // (scale, shift) = (gamma*inv_std, beta - mean*scale)
    float2* bnp;      // (mean, inv_std)
    double2* bst;     // backward sums (sum dz, sum dz*xhat)
};

// The fp64 sums are the one place where thousands of CTAs hit the same addresses; they are spread over
// kStatCopies replicas (chosen by block index) and folded by the last block, which cuts the same-address
// atomic contention that otherwise puts a ~150 us floor under every kernel.
constexpr int kStatCopies = 4;
CDRA_DEV int stat_copy() { return (int)((blockIdx.x + 5u * blockIdx.z) & (kStatCopies - 1)); }
CDRA_DEV double2* stat_slot(double2* table, int ld, int copy, int t, int c) {
    return table + ((size_t)copy * kT + t) * ld + c;
}
CDRA_DEV double2 stat_fold(const double2* table, int ld, int t, int c) {
    double sx = 0.0, sy = 0.0;
    for (int k = 0; k < kStatCopies; ++k) {
        const volatile double2* p = table + ((size_t)k * kT + t) * ld + c;
        sx += p->x; sy += p->y;
    }
    return make_double2(sx, sy);
}

// Destination-channel map of a pointwise conv's output column j (split = channel-shuffle aware):
//   split=0: weight column j   -> channel j
//   split=1: j <  N/2: weight column 2j        -> channel off + j                (even outputs)
//            j >= N/2: weight column 2(j-N/2)+1 -> channel chalf + off + (j-N/2) (odd outputs)
// which is exactly out[g*C/2+i] = concat[2i+g] of core/architectures.py:109-118 for an output block
// that starts at (even) concat position 2*off.
struct ColMap {
    int n, split, off, chalf;
};
CDRA_DEV int colmap_w(const ColMap& m, int j) {
    if (!m.split) return j;
    int h = m.n >> 1;
    return j < h ? 2 * j : 2 * (j - h) + 1;
}
CDRA_DEV int colmap_c(const ColMap& m, int j) {
    if (!m.split) return j;
    int h = m.n >> 1;
    return j < h ? m.off + j : m.chalf + m.off + (j - h);
}

CDRA_DEV float warp_sum(float v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
CDRA_DEV double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// "last block done" ticket: returns true in exactly one block (the last to arrive); resets the counter.
CDRA_DEV bool last_block_ticket(unsigned* counter, unsigned total) {
    CDRA_SHARED unsigned s_ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        unsigned t = atomicAdd(counter, 1u);
        s_ticket = t;
        if (t == total - 1) *counter = 0;
    }
    __syncthreads();
    __threadfence();
    return s_ticket == total - 1;
}

// BatchNorm finalisation of one layer by the last block of its producer kernel: turns the fp64 sums
// into (scale, shift), (mean, inv_std) for every (slice, channel), and applies the kT sequential
// moving-average updates Keras performs for a time-shared layer (core/architectures.py:44-57).
//   n        rows per slice;   unbiased   1 = FusedBatchNorm (4-D input) moving variance
CDRA_DEV void bn_finalize(const ColMap& cm, const BnTables& tb, int ld, const float* gamma, const float* beta,
                          float* mov_mean, float* mov_var, double n, int unbiased, int training, int nthreads, int tid) {
    for (int j = tid; j < cm.n; j += nthreads) {
        const int c = colmap_c(cm, j), w = colmap_w(cm, j);
        const float g = gamma[w], b = beta[w];
        float mm = mov_mean ? mov_mean[w] : 0.f, mv = mov_var ? mov_var[w] : 1.f;
        for (int t = 0; t < kT; ++t) {
            const double2 sp = stat_fold(tb.fst, ld, t, c);
            const double sx = sp.x, sxx = sp.y;
            double mean = sx / n;
            double var = sxx / n - mean * mean;
            if (var < 0.0) var = 0.0;
            const float inv = (float)(1.0 / sqrt(var + (double)kBnEps));
            const float scale = g * inv;
            tb.aff[(size_t)t * ld + c] = make_float2(scale, b - (float)mean * scale);
            tb.bnp[(size_t)t * ld + c] = make_float2((float)mean, inv);
            if (training) {
                const double vm = unbiased ? var * (n / (n > 1.0 ? n - 1.0 : 1.0)) : var;
                // Keras assign_moving_average: variable -= (variable - value) * (1 - momentum)
                mm -= (mm - (float)mean) * (1.f - kBnMomentum);
                mv -= (mv - (float)vm) * (1.f - kBnMomentum);
            }
        }
        if (training && mov_mean) { mov_mean[w] = mm; mov_var[w] = mv; }
    }
}

}  // namespace cdra
