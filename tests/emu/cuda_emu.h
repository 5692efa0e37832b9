// Minimal SIMT emulator so the library's plain-CUDA kernels (no inline PTX, no tensor-core paths)
// can be *logic-checked* on a CPU-only box.  TEST INFRASTRUCTURE ONLY: the emulated library
// (tests/emu/libcdra_emu.so) is loaded exclusively by tests marked "not gpu"; the product loader
// (cdra/_lib.py) refuses it, and bench.py / smoke() never touch it.
//
// Model: one OS thread; every CUDA thread of a block is a ucontext fiber; blocks run one after the
// other.  __syncthreads / warp collectives are generation barriers that yield to the scheduler.
#pragma once
#include <ucontext.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdio>
#include <vector>
#include <functional>
#include <algorithm>

#include <vector_types.h>
#include <vector_functions.h>
typedef uint3 uint3_emu;
using std::min;
using std::max;

namespace emu {
struct Barrier { int arrived = 0; int gen = 0; };
struct State {
    uint3_emu tid, bid;
    dim3 bdim, gdim;
    int linear_tid = 0;
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<char*> stacks;
    std::vector<char> done;
    std::function<void()> body;
    Barrier block_bar;
    std::vector<Barrier> warp_bar;
    std::vector<uint64_t> warp_buf;   // 32 slots per warp
    char* dyn_smem = nullptr;
    size_t dyn_cap = 0;
    int cur = -1;
};
inline State& S() { static State s; return s; }

inline void yield() {
    State& s = S();
    int me = s.cur;
    swapcontext(&s.ctx[me], &s.sched);
}
inline void barrier_wait(Barrier& b, int count) {
    int g = b.gen;
    if (++b.arrived == count) { b.arrived = 0; b.gen++; }
    else { while (b.gen == g) yield(); }
}
inline void fiber_entry() {
    State& s = S();
    int me = s.cur;
    s.body();
    s.done[me] = 1;
    swapcontext(&s.ctx[me], &s.sched);
}
inline void set_thread(int i) {
    State& s = S();
    s.cur = i; s.linear_tid = i;
    s.tid.x = i % s.bdim.x; s.tid.y = (i / s.bdim.x) % s.bdim.y; s.tid.z = i / (s.bdim.x * s.bdim.y);
}
inline void launch(dim3 grid, dim3 block, size_t shmem, std::function<void()> body) {
    State& s = S();
    const int nt = block.x * block.y * block.z;
    const size_t STK = 256 * 1024;
    s.bdim = block; s.gdim = grid; s.body = body;
    if ((int)s.stacks.size() < nt) {
        size_t old = s.stacks.size();
        s.stacks.resize(nt);
        for (size_t i = old; i < (size_t)nt; ++i) s.stacks[i] = (char*)malloc(STK);
    }
    s.ctx.resize(nt); s.done.assign(nt, 0);
    if (shmem > s.dyn_cap) { free(s.dyn_smem); s.dyn_smem = (char*)aligned_alloc(1024, ((shmem + 1023) / 1024) * 1024); s.dyn_cap = shmem; }
    const int nwarps = (nt + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.bid = {bx, by, bz};
        s.block_bar = Barrier();
        s.warp_bar.assign(nwarps, Barrier());
        s.warp_buf.assign(nwarps * 32, 0);
        std::fill(s.done.begin(), s.done.end(), 0);
        for (int i = 0; i < nt; ++i) {
            getcontext(&s.ctx[i]);
            s.ctx[i].uc_stack.ss_sp = s.stacks[i];
            s.ctx[i].uc_stack.ss_size = STK;
            s.ctx[i].uc_link = &s.sched;
            makecontext(&s.ctx[i], (void (*)())fiber_entry, 0);
        }
        int remaining = nt;
        while (remaining > 0) {
            for (int i = 0; i < nt; ++i) {
                if (s.done[i] == 1) continue;
                if (s.done[i] == 2) continue;
                set_thread(i);
                swapcontext(&s.sched, &s.ctx[i]);
                if (s.done[i] == 1) { s.done[i] = 2; --remaining; }
            }
        }
    }
}
inline int warp_count(int w) {  // active lanes of warp w
    State& s = S();
    int nt = s.bdim.x * s.bdim.y * s.bdim.z;
    return std::min(32, nt - w * 32);
}
template <typename T> inline T shfl(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shfl width");
    State& s = S();
    int w = s.linear_tid / 32, l = s.linear_tid % 32, n = warp_count(w);
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    s.warp_buf[w * 32 + l] = raw;
    barrier_wait(s.warp_bar[w], n);
    uint64_t got = s.warp_buf[w * 32 + ((src_lane % 32 + 32) % 32 < n ? (src_lane % 32 + 32) % 32 : l)];
    barrier_wait(s.warp_bar[w], n);
    T out; memcpy(&out, &got, sizeof(T));
    return out;
}
}  // namespace emu

#define threadIdx (emu::S().tid)
#define blockIdx (emu::S().bid)
#define blockDim (emu::S().bdim)
#define gridDim (emu::S().gdim)

inline void __syncthreads() { emu::State& s = emu::S(); emu::barrier_wait(s.block_bar, s.bdim.x * s.bdim.y * s.bdim.z); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::State& s = emu::S(); int w = s.linear_tid / 32; emu::barrier_wait(s.warp_bar[w], emu::warp_count(w)); }
inline void __threadfence() {}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shfl(v, (emu::S().linear_tid % 32) ^ m); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int d) { int l = emu::S().linear_tid % 32; return emu::shfl(v, l + d < 32 ? l + d : l); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl(v, src); }

template <typename T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned atomicInc(unsigned* p, unsigned lim) { unsigned o = *p; *p = (o >= lim) ? 0 : o + 1; return o; }
inline int atomicMax(int* p, int v) { int o = *p; *p = std::max(o, v); return o; }
inline int atomicMin(int* p, int v) { int o = *p; *p = std::min(o, v); return o; }
inline unsigned atomicExch(unsigned* p, unsigned v) { unsigned o = *p; *p = v; return o; }

inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
inline float __expf(float x) { return expf(x); }
inline float __ldg(const float* p) { return *p; }
inline float fminf_(float a, float b) { return a < b ? a : b; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }


