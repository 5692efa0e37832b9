// On-device image augmentation: the `augment_fn` closure of the reference (core/carla_agent.py:527-579) -- colour jitter
// (rl/augmentations/simclr.py:44-58: brightness, contrast, saturation, hue, clip), random-kernel blur
// (rl/augmentations/augmentations.py:193-207), salt & pepper (:176-191), gaussian noise (:147-157), per-sample min-max
// normalisation (:253-263), cutout (:56-68) and coarse dropout (:81-92) -- as three launches over the frame batch.
//
// Randomness is explicit: the scalars TF would draw once per call (chance gates, jitter factors, the blur kernel, the
// cutout cell, the dropout grid) arrive in `cdra_augment_params` (the host mirror draws them); the per-pixel draws
// (selection masks, noise) come from a counter-based hash of (seed, frame, pixel, stream), which the oracle restates
// bit for bit, so kernel and oracle can be compared on identical random numbers.
// Reference quirks kept: cutout / coarse-dropout masks of the FIRST image are applied to the whole batch (the `[0]` after
// tf.image.resize of the mask batch); gaussian noise only ever adds (the product mask * noise is clipped to [0, 1]).
#pragma once
#ifndef CDRA_EMU
#include "cdra_common.cuh"
#include "../../include/cdra.h"

namespace cdra {
namespace aug {

__host__ __device__ inline uint32_t hash32(uint32_t seed, uint32_t a, uint32_t b) {      // "lowbias32" finaliser over a mixed key
    uint32_t x = seed ^ (a * 0x9E3779B1u) ^ (b * 0x85EBCA77u);
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ inline uint32_t fkey(float v) {          // order-preserving float -> uint key (atomicMin / atomicMax on floats of any sign)
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ inline float fkey_inv(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

struct AugArgs {
    const void* img; int u8; long long frames; int H, W;
    cdra_augment_params p;
    const uint8_t* dropout_mask;
    float* out;
    float* mean;            // [frames][3] channel means of the input (contrast)
    uint32_t* kmin; uint32_t* kmax;     // [groups] order-preserving keys of the per-sample min / max
};

__device__ inline void load_px(const AugArgs& a, long long f, int y, int x, float (&v)[3]) {
    const long long i = ((f * a.H + y) * a.W + x) * 3;
    if (a.u8) { const uint8_t* p = (const uint8_t*)a.img + i; v[0] = p[0] * (1.f / 255.f); v[1] = p[1] * (1.f / 255.f); v[2] = p[2] * (1.f / 255.f); }
    else { const float* p = (const float*)a.img + i; v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; }
}
__device__ inline void rgb_to_hsv(const float (&c)[3], float& h, float& s, float& v) {
    const float mx = fmaxf(c[0], fmaxf(c[1], c[2])), mn = fminf(c[0], fminf(c[1], c[2])), d = mx - mn;
    v = mx; s = mx > 0.f ? d / mx : 0.f;
    if (d <= 0.f) h = 0.f;
    else if (mx == c[0]) { h = (c[1] - c[2]) / d; if (h < 0.f) h += 6.f; }
    else if (mx == c[1]) h = (c[2] - c[0]) / d + 2.f;
    else h = (c[0] - c[1]) / d + 4.f;
    h *= (1.f / 6.f);
}
__device__ inline void hsv_to_rgb(float h, float s, float v, float (&c)[3]) {
    const float dh = h * 6.f;
    const float dr = fminf(fmaxf(fabsf(dh - 3.f) - 1.f, 0.f), 1.f), dg = fminf(fmaxf(2.f - fabsf(dh - 2.f), 0.f), 1.f),
                db = fminf(fmaxf(2.f - fabsf(dh - 4.f), 0.f), 1.f);
    c[0] = ((dr - 1.f) * s + 1.f) * v; c[1] = ((dg - 1.f) * s + 1.f) * v; c[2] = ((db - 1.f) * s + 1.f) * v;
}
// colour jitter of one pixel: brightness -> contrast (around the frame's channel mean) -> saturation -> hue -> clip
__device__ inline void jitter_px(const AugArgs& a, long long f, float (&c)[3]) {
    const cdra_augment_params& p = a.p;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float m = a.mean[f * 3 + k] + p.brightness;
        c[k] = (c[k] + p.brightness - m) * p.contrast + m;
    }
    float h, s, v;
    rgb_to_hsv(c, h, s, v);
    s = fminf(fmaxf(s * p.saturation, 0.f), 1.f);
    hsv_to_rgb(h, s, v, c);
    rgb_to_hsv(c, h, s, v);
    h += p.hue; h -= floorf(h);
    hsv_to_rgb(h, s, v, c);
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = fminf(fmaxf(c[k], 0.f), 1.f);
}

__global__ void __launch_bounds__(256) aug_mean_kernel(const AugArgs a) {
    __shared__ float red[3][8];
    const long long f = blockIdx.x;
    float s[3] = {0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < a.H * a.W; i += 256) {
        float v[3];
        load_px(a, f, i / a.W, i % a.W, v);
        s[0] += v[0]; s[1] += v[1]; s[2] += v[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        a.mean[f * 3 + threadIdx.x] = t / (float)(a.H * a.W);
    }
}

// jitter -> blur -> salt & pepper -> gaussian noise, one thread per pixel; per-sample min / max of the result
__global__ void __launch_bounds__(256) aug_main_kernel(const AugArgs a) {
    const cdra_augment_params& p = a.p;
    const long long f = blockIdx.y;
    const int px = blockIdx.x * 256 + threadIdx.x, HW = a.H * a.W;
    float c[3] = {0.f, 0.f, 0.f};
    const bool on = px < HW;
    if (on) {
        const int y = px / a.W, x = px - y * a.W;
        if (p.blur_size > 0) {
            const int r = p.blur_size / 2;
            for (int i = 0; i < p.blur_size; ++i)
                for (int j = 0; j < p.blur_size; ++j) {
                    const int yy = y + i - r, xx = x + j - r;
                    if (yy < 0 || yy >= a.H || xx < 0 || xx >= a.W) continue;
                    float v[3];
                    load_px(a, f, yy, xx, v);
                    if (p.jitter) jitter_px(a, f, v);
                    const float* kw = p.blur_kernel + (i * p.blur_size + j) * 3;
                    c[0] = fmaf(kw[0], v[0], c[0]); c[1] = fmaf(kw[1], v[1], c[1]); c[2] = fmaf(kw[2], v[2], c[2]);
                }
        } else {
            load_px(a, f, y, x, c);
            if (p.jitter) jitter_px(a, f, c);
        }
        const uint32_t fi = (uint32_t)f, b = (uint32_t)px * 16u;
        if (p.salt_pepper) {
            const uint32_t thr = (uint32_t)(p.sp_amount * 0.1f * 16777216.f);
            if ((hash32(p.seed, fi, b) >> 8) < thr) {
                const float nz = (hash32(p.seed, fi, b + 1) >> 8) < 8388608u ? 1.f : 0.f;
                c[0] = c[1] = c[2] = nz;
            }
        }
        if (p.gauss_noise) {
            const uint32_t thr = (uint32_t)(p.gn_amount * 16777216.f);
            if ((hash32(p.seed, fi, b + 2) >> 8) < thr) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float u1 = (float)((hash32(p.seed, fi, b + 3 + k) >> 8) + 1u) * (1.f / 16777216.f);
                    const float u2 = (float)(hash32(p.seed, fi, b + 6 + k) >> 8) * (1.f / 16777216.f);
                    const float n = sqrtf(-2.f * logf(u1)) * cosf(6.28318530717958647692f * u2) * p.gn_std;
                    c[k] += fminf(fmaxf(n, 0.f), 1.f);
                }
            }
        }
        float* o = a.out + (f * HW + px) * 3;
        o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
    }
    if (p.normalize) {
        float mn = on ? fminf(c[0], fminf(c[1], c[2])) : 3.4e38f, mx = on ? fmaxf(c[0], fmaxf(c[1], c[2])) : -3.4e38f;
        for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
        if ((threadIdx.x & 31) == 0) {
            const long long g = f / p.group;
            atomicMin(a.kmin + g, fkey(mn)); atomicMax(a.kmax + g, fkey(mx));
        }
    }
}

// min-max normalisation per sample, cutout, coarse dropout (masks of the first image, nearest-neighbour resize with
// half-pixel centres: grid cell = floor((i + 0.5) * size / extent))
__global__ void __launch_bounds__(256) aug_finish_kernel(const AugArgs a) {
    const cdra_augment_params& p = a.p;
    const long long f = blockIdx.y;
    const int px = blockIdx.x * 256 + threadIdx.x, HW = a.H * a.W;
    if (px >= HW) return;
    const int y = px / a.W, x = px - y * a.W;
    float* o = a.out + (f * HW + px) * 3;
    float c[3] = {o[0], o[1], o[2]};
    if (p.normalize) {
        const long long g = f / p.group;
        const float mn = fkey_inv(a.kmin[g]), mx = fkey_inv(a.kmax[g]);
        const float den = (mx - mn) + p.eps;
#pragma unroll
        for (int k = 0; k < 3; ++k) c[k] = (c[k] - mn) / den;
    }
    float keep = 1.f;
    if (p.cutout_size > 0) {
        const int cy = min((int)(((float)y + 0.5f) * (float)p.cutout_size / (float)a.H), p.cutout_size - 1);
        const int cx = min((int)(((float)x + 0.5f) * (float)p.cutout_size / (float)a.W), p.cutout_size - 1);
        if (cy * p.cutout_size + cx == p.cutout_cell) keep = 0.f;
    }
    if (p.dropout_size > 0 && a.dropout_mask) {
        const int cy = min((int)(((float)y + 0.5f) * (float)p.dropout_size / (float)a.H), p.dropout_size - 1);
        const int cx = min((int)(((float)x + 0.5f) * (float)p.dropout_size / (float)a.W), p.dropout_size - 1);
        if (a.dropout_mask[cy * p.dropout_size + cx] == 0) keep = 0.f;
    }
    o[0] = c[0] * keep; o[1] = c[1] * keep; o[2] = c[2] * keep;
}

}  // namespace aug
}  // namespace cdra
#endif
