// PPO numerics: output heads + fused losses (core/carla_agent.py:394-428,469-486 on
// core/networks.py:96-137,255-275), GAE / returns (rl/agents/ppo.py:692-727, rl/utils.py:57-84,
// 140-151,344-349), per-tensor clip + Keras Adam (rl/utils.py:120-121, rl/agents/ppo.py:238-275),
// minibatch row gather (rl/utils.py:365-393).
#pragma once
#include "cdra_common.cuh"

namespace cdra {

constexpr int kHU = 320;              // control-branch units
constexpr float kActEps = 1.1920928955078125e-07f;     // utils.EPSILON (rl/utils.py:24-25)

CDRA_DEV double digamma_d(double x) {
    double r = 0.0;
    while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
    const double f = 1.0 / (x * x);
    return r + log(x) - 0.5 / x - f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132)))));
}
CDRA_DEV double trigamma_d(double x) {
    double r = 0.0;
    while (x < 6.0) { r += 1.0 / (x * x); x += 1.0; }
    const double f = 1.0 / (x * x);
    return r + 1.0 / x + 0.5 * f + (1.0 / x) * f * (1.0 / 6 - f * (1.0 / 30 - f * (1.0 / 42 - f * (1.0 / 30 - f * (5.0 / 66)))));
}
CDRA_DEV double softplus_d(double x) { return (x > 0 ? x : 0.0) + log1p(exp(-fabs(x))); }
CDRA_DEV double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

struct HeadLossArgs {
    const float* a2;                  // [B][320] control-branch output
    const float* w[4]; const float* b[4]; int n[4];     // output heads
    // policy inputs
    const float* actions; const float* logp_old; const float* adv;
    const float* actions_jac;         // [B][2][2] (d action / d alpha, d action / d beta) of a reparameterised sample, or null
    // value inputs
    const float* returns_be;
    const float* true_speed; const float* true_sim;
    float clip, ent_coef, grad_scale, exp_scale;
    int B;
    double* acc;                      // [32] loss accumulators (+ ticket counter at acc[31])
    float* scalars;                   // [16]
    float* head_out;                  // policy: [B][8] alpha2 beta2 sim speed logp2 ; value: [B][4]
    float* da2;                       // [B][320]
    float* dw[4]; float* db[4];
};

constexpr int kHeadRows = 8;          // rows (= warps) per block

// policy == true : logits = [alpha0 alpha1 beta0 beta1 similarity speed]
// policy == false: logits = [base exp speed similarity]
template <bool POLICY>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) head_loss_kernel(HeadLossArgs a) {
    constexpr int NL = POLICY ? 6 : 4;
    CDRA_SHARED float sw[kHU * NL];
    CDRA_SHARED float sb[NL];
    CDRA_SHARED float sdl[kHeadRows][NL];
    CDRA_SHARED double sacc[16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {   // stage head weights as one [320][NL] matrix
        int col = 0;
        for (int h = 0; h < 4; ++h) {
            for (int i = tid; i < kHU * a.n[h]; i += 256) { const int k = i / a.n[h], j = i - k * a.n[h]; sw[k * NL + col + j] = a.w[h][i]; }
            if (tid < a.n[h]) sb[col + tid] = a.b[h][tid];
            col += a.n[h];
        }
        if (tid < 16) sacc[tid] = 0.0;
    }
    __syncthreads();
    const int row = blockIdx.x * kHeadRows + warp;
    const bool valid = row < a.B;
    float dots[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) dots[j] = 0.f;
    if (valid) for (int k = lane; k < kHU; k += 32) {
        const float x = a.a2[(size_t)row * kHU + k];
#pragma unroll
        for (int j = 0; j < NL; ++j) dots[j] = fmaf(x, sw[k * NL + j], dots[j]);
    }
#pragma unroll
    for (int j = 0; j < NL; ++j) dots[j] = warp_sum(dots[j]) + sb[j];
    if (lane == 0) {
        double dl[NL];
        for (int j = 0; j < NL; ++j) dl[j] = 0.0;
        if (valid) {
            const double invB = 1.0 / a.B;
            if (POLICY) {
                const double A = a.adv[row], c = a.clip;
                double ratio = 0.0, rat[2], dlogp_da[2], dlogp_db[2], dH_da[2], dH_db[2], sig_a[2], sig_b[2], Hs = 0.0, lps = 0.0;
                for (int i = 0; i < 2; ++i) {
                    const double la = dots[i], lb = dots[2 + i];
                    const double al = softplus_d(la) + 1.01, be = softplus_d(lb) + 1.01;      // core/networks.py:133-134
                    sig_a[i] = sigmoid_d(la); sig_b[i] = sigmoid_d(lb);
                    float xf = a.actions[(size_t)row * 2 + i];
                    xf = fminf(fmaxf(xf, kActEps), 1.f - kActEps);                          // _clip_actions :139-144
                    const double x = xf;
                    const double lbeta = lgamma(al) + lgamma(be) - lgamma(al + be);
                    const double pa = digamma_d(al), pb = digamma_d(be), pab = digamma_d(al + be);
                    const double logp = (al - 1.0) * log(x) + (be - 1.0) * log1p(-x) - lbeta;
                    const double H = lbeta - (al - 1.0) * pa - (be - 1.0) * pb + (al + be - 2.0) * pab;
                    const double tab = trigamma_d(al + be);
                    dlogp_da[i] = log(x) - (pa - pab); dlogp_db[i] = log1p(-x) - (pb - pab);
                    dH_da[i] = -(al - 1.0) * trigamma_d(al) + (al + be - 2.0) * tab;
                    dH_db[i] = -(be - 1.0) * trigamma_d(be) + (al + be - 2.0) * tab;
                    if (a.actions_jac) {
                        // the evaluated action is a reparameterised sample x(alpha, beta) of the new policy (PolicyNetwork.call,
                        // core/networks.py:97-100): d logp / d alpha = partial + d logp / dx * dx / d alpha.  tf.clip_by_value
                        // passes the gradient only where the sample was not clipped.
                        const float xr = a.actions[(size_t)row * 2 + i];
                        if (xr >= kActEps && xr <= 1.f - kActEps) {
                            const double dlogp_dx = (al - 1.0) / x - (be - 1.0) / (1.0 - x);
                            dlogp_da[i] += dlogp_dx * (double)a.actions_jac[(size_t)row * 4 + 2 * i];
                            dlogp_db[i] += dlogp_dx * (double)a.actions_jac[(size_t)row * 4 + 2 * i + 1];
                        }
                    }
                    rat[i] = exp(logp - (double)a.logp_old[(size_t)row * 2 + i]);
                    ratio += 0.5 * rat[i]; Hs += H; lps += logp;
                    a.head_out[(size_t)row * 8 + i] = (float)al; a.head_out[(size_t)row * 8 + 2 + i] = (float)be;
                    a.head_out[(size_t)row * 8 + 6 + i] = (float)logp;
                }
                const double min_adv = A > 0.0 ? (1.0 + c) * A : (1.0 - c) * A;
                const double surr = ratio * A;
                const double s = surr <= min_adv ? surr : min_adv;
                const double ds_dratio = surr <= min_adv ? A : 0.0;
                const double sim = tanh((double)dots[4]), spd = 2.0 * sigmoid_d((double)dots[5]);
                const double es = spd - (double)a.true_speed[row], em = sim - (double)a.true_sim[row];
                a.head_out[(size_t)row * 8 + 4] = (float)sim; a.head_out[(size_t)row * 8 + 5] = (float)spd;
                for (int i = 0; i < 2; ++i) {
                    const double dlp = -invB * ds_dratio * 0.5 * rat[i];       // d L / d logp_i
                    const double dH = -a.ent_coef * invB * 0.5;                // d L / d H_i
                    dl[i] = (dlp * dlogp_da[i] + dH * dH_da[i]) * sig_a[i];
                    dl[2 + i] = (dlp * dlogp_db[i] + dH * dH_db[i]) * sig_b[i];
                }
                dl[4] = invB * em * (1.0 - sim * sim);                         // 0.5 * mean (sim - t)^2
                dl[5] = invB * es * spd * (1.0 - 0.5 * spd);                   // speed = 2 sigmoid
                atomicAdd(&sacc[0], -s * invB);                 // loss_policy
                atomicAdd(&sacc[1], Hs * 0.5 * invB);           // entropy (mean over b, a)
                atomicAdd(&sacc[2], 0.5 * es * es * invB);      // loss_speed
                atomicAdd(&sacc[3], 0.5 * em * em * invB);      // loss_similarity
                atomicAdd(&sacc[4], ratio * invB);
                atomicAdd(&sacc[5], lps * 0.5 * invB);
                atomicAdd(&sacc[6], spd * invB);
                atomicAdd(&sacc[7], sim * invB);
            } else {
                const double base = tanh((double)dots[0]), ex = a.exp_scale * sigmoid_d((double)dots[1]);
                const double spd = 2.0 * sigmoid_d((double)dots[2]), sim = tanh((double)dots[3]);
                const double e0 = base - (double)a.returns_be[(size_t)row * 2], e1 = ex - (double)a.returns_be[(size_t)row * 2 + 1];
                const double es = spd - (double)a.true_speed[row], em = sim - (double)a.true_sim[row];
                const double k = 0.25 * invB, es2 = (double)a.exp_scale * a.exp_scale;
                dl[0] = k * 0.25 * 2.0 * e0 * (1.0 - base * base);
                dl[1] = k * (2.0 * e1 / es2) * ex * (1.0 - ex / a.exp_scale);
                dl[2] = k * 2.0 * es * spd * (1.0 - 0.5 * spd);
                dl[3] = k * 2.0 * em * (1.0 - sim * sim);
                a.head_out[(size_t)row * 4] = (float)base; a.head_out[(size_t)row * 4 + 1] = (float)ex;
                a.head_out[(size_t)row * 4 + 2] = (float)spd; a.head_out[(size_t)row * 4 + 3] = (float)sim;
                atomicAdd(&sacc[0], (0.25 * e0 * e0 + e1 * e1 / es2) * invB);      // loss_v
                atomicAdd(&sacc[1], es * es * invB);
                atomicAdd(&sacc[2], em * em * invB);
                atomicAdd(&sacc[3], spd * invB);
                atomicAdd(&sacc[4], sim * invB);
            }
        }
        for (int j = 0; j < NL; ++j) sdl[warp][j] = (float)(dl[j] * a.grad_scale);
    }
    __syncthreads();
    // d a2[row][k] = sum_j dl[j] W[k][j]
    if (valid) for (int k = lane; k < kHU; k += 32) {
        float g = 0.f;
#pragma unroll
        for (int j = 0; j < NL; ++j) g = fmaf(sdl[warp][j], sw[k * NL + j], g);
        a.da2[(size_t)row * kHU + k] = g;
    }
    // head weight gradients: dW[k][j] += sum_rows a2[row][k] * dl[row][j]
    const int nrows = min(kHeadRows, a.B - (int)blockIdx.x * kHeadRows);
    for (int k = tid; k < kHU; k += 256) {
        float g[NL];
#pragma unroll
        for (int j = 0; j < NL; ++j) g[j] = 0.f;
        for (int r = 0; r < nrows; ++r) {
            const float x = a.a2[((size_t)blockIdx.x * kHeadRows + r) * kHU + k];
#pragma unroll
            for (int j = 0; j < NL; ++j) g[j] = fmaf(x, sdl[r][j], g[j]);
        }
        int col = 0;
        for (int h = 0; h < 4; ++h) { for (int j = 0; j < a.n[h]; ++j) atomicAdd(a.dw[h] + (size_t)k * a.n[h] + j, g[col + j]); col += a.n[h]; }
    }
    if (tid < NL) {
        float g = 0.f;
        for (int r = 0; r < nrows; ++r) g += sdl[r][tid];
        int col = 0;
        for (int h = 0; h < 4; ++h) { if (tid >= col && tid < col + a.n[h]) atomicAdd(a.db[h] + (tid - col), g); col += a.n[h]; }
    }
    if (tid < 8) atomicAdd(&a.acc[tid], sacc[tid]);
    if (last_block_ticket((unsigned*)(a.acc + 31), gridDim.x) && tid == 0) {
        volatile double* v = a.acc;
        if (POLICY) {
            const double lp = v[0], H = v[1], ls = v[2], lm = v[3];
            a.scalars[0] = (float)(lp - a.ent_coef * H + ls + lm);
            a.scalars[1] = (float)lp; a.scalars[2] = (float)(a.ent_coef * H); a.scalars[3] = (float)ls; a.scalars[4] = (float)lm;
            a.scalars[5] = (float)v[4]; a.scalars[6] = (float)v[5]; a.scalars[7] = (float)H; a.scalars[8] = (float)v[6]; a.scalars[9] = (float)v[7];
        } else {
            a.scalars[0] = (float)((v[0] + v[1] + v[2]) * 0.25);
            a.scalars[1] = (float)v[0]; a.scalars[2] = (float)v[1]; a.scalars[3] = (float)v[2]; a.scalars[4] = (float)v[3]; a.scalars[5] = (float)v[4];
        }
    }
}

// --------------------------------------------------------------------------- returns + GAE (one warp per trajectory)
struct GaeArgs {
    const float* rewards;       // [bs][T]
    const float* values_be;     // [bs][T][2]
    const float* last_be;       // [bs][2]
    double gamma, gamma_lambda;
    float gamma_f, scale;
    int bs, T;
    float* returns_be;          // [bs][T][2]
    float* adv;                 // [bs][T]
};

CDRA_KERNEL CDRA_LAUNCH_BOUNDS(32) gae_kernel(GaeArgs a) {
    CDRA_DYN_SMEM(smem_raw);
    const int T = a.T, lane = threadIdx.x, traj = blockIdx.x;
    float* r = (float*)smem_raw;              // [T+1] rewards with bootstrap
    float* v = r + (T + 1);                   // [T+1] decoded values
    float* ret = v + (T + 1);                 // [T+1]
    float* adv = ret + (T + 1);               // [T]
    // value = base * 10^exponent (rl/agents/ppo.py:694,717), fp32 like tf.pow
    for (int i = lane; i <= T; i += 32) {
        const float* be = i < T ? a.values_be + ((size_t)traj * T + i) * 2 : a.last_be + (size_t)traj * 2;
        const float p = (float)pow(10.0, (double)be[1]);
        v[i] = __fmul_rn(be[0], p);
        if (i < T) r[i] = a.rewards[(size_t)traj * T + i];
    }
    __syncwarp();
    if (lane == 0) r[T] = v[T];               // end_trajectory appends v_T to the rewards (:694-696)
    __syncwarp();
    // scipy.signal.lfilter([1],[1,-d], x[::-1])[::-1] in float64: y[n] = x[n] + d*y[n+1], sequential,
    // product and sum rounded separately (direct-form II transposed, no FMA) -> bit-identical scan.
    if (lane == 0) {
        double y = 0.0;
        for (int i = T; i >= 0; --i) { y = __dadd_rn((double)r[i], __dmul_rn(a.gamma, y)); ret[i] = (float)y; }
        double g = 0.0;
        for (int i = T - 1; i >= 0; --i) {
            const float delta = __fsub_rn(__fadd_rn(r[i], __fmul_rn(a.gamma_f, v[i + 1])), v[i]);   // rl/utils.py:66
            g = __dadd_rn((double)delta, __dmul_rn(a.gamma_lambda, g));
            adv[i] = (float)g;
        }
    }
    __syncwarp();
    // decompose_number (rl/utils.py:140-151) in fp32; sign-preserving normalisation (:344-349)
    float mx = -INFINITY, mn = INFINITY;
    for (int i = lane; i < T; i += 32) {
        float num = ret[i], e = 0.f;
        while (fabsf(num) > 1.0f) { num = __fdiv_rn(num, 10.0f); e += 1.f; }
        a.returns_be[((size_t)traj * T + i) * 2] = num;
        a.returns_be[((size_t)traj * T + i) * 2 + 1] = e;
        mx = fmaxf(mx, adv[i]); mn = fminf(mn, adv[i]);
    }
    for (int o = 16; o; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); }
    const float dpos = __fadd_rn(mx, 1e-3f), dneg = -__fsub_rn(mn, 1e-3f);
    for (int i = lane; i < T; i += 32) {
        const float x = adv[i];
        const float pos = x > 0.f ? x : 0.f, neg = x < 0.f ? x : 0.f;
        const float nrm = __fadd_rn(__fdiv_rn(pos, dpos), __fdiv_rn(neg, dneg));
        a.adv[(size_t)traj * T + i] = __fmul_rn(nrm, a.scale);
    }
}

// --------------------------------------------------------------------------- per-tensor clip + Keras Adam
struct AdamArgs {
    float* p; const float* g; float* m; float* v;
    const int64_t* offs; int n_tensors;
    float clip, lr_t, beta1, beta2, eps, grad_scale;
    float* norms;             // [n_tensors] sum of squares of (grad_scale * g) per tensor
    int64_t total;
};
constexpr int kAdamChunk = 2048;

CDRA_DEV int find_tensor(const int64_t* offs, int n, int64_t i) {
    int lo = 0, hi = n - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (offs[mid] <= i) lo = mid; else hi = mid - 1; }
    return lo;
}

CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) sqnorm_kernel(AdamArgs a) {
    CDRA_SHARED float red[8];
    const int64_t c0 = (int64_t)blockIdx.x * kAdamChunk, c1 = min(a.total, c0 + (int64_t)kAdamChunk);
    int ti = find_tensor(a.offs, a.n_tensors, c0);
    int64_t s = c0;
    while (s < c1) {                       // walk the tensors that intersect this chunk
        const int64_t e = min(c1, a.offs[ti + 1]);
        float acc = 0.f;
        for (int64_t i = s + threadIdx.x; i < e; i += 256) { const float g = a.g[i] * a.grad_scale; acc = fmaf(g, g, acc); }
        acc = warp_sum(acc);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += red[w]; atomicAdd(a.norms + ti, t); }
        s = e; ++ti;
    }
}

CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) adam_kernel(AdamArgs a) {
    const int64_t c0 = (int64_t)blockIdx.x * kAdamChunk, c1 = min(a.total, c0 + (int64_t)kAdamChunk);
    int ti = a.clip > 0.f ? find_tensor(a.offs, a.n_tensors, c0) : 0;
    int64_t s = c0;
    while (s < c1) {
        int64_t e = c1;
        float mult = a.grad_scale;
        if (a.clip > 0.f) {                // tf.clip_by_norm: g * clip / max(||g||, clip)
            e = min(c1, a.offs[ti + 1]);
            const float l2sum = a.norms[ti];
            const float l2 = sqrtf(l2sum > 0.f ? l2sum : 1.f);
            mult = a.grad_scale * a.clip / fmaxf(l2, a.clip);
        }
        for (int64_t i = s + threadIdx.x; i < e; i += 256) {
            const float g = a.g[i] * mult;
            const float m = a.beta1 * a.m[i] + (1.f - a.beta1) * g;
            const float v = a.beta2 * a.v[i] + (1.f - a.beta2) * g * g;
            a.m[i] = m; a.v[i] = v;
            a.p[i] -= a.lr_t * m / (sqrtf(v) + a.eps);
        }
        s = e; ++ti;
    }
}

// --------------------------------------------------------------------------- row gather
struct GatherArgs { const char* src; const int64_t* index; int64_t n, row_bytes; char* dst; };
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) gather_rows_kernel(GatherArgs a) {
    const int64_t row = blockIdx.x;
    const char* s = a.src + a.index[row] * a.row_bytes;
    char* d = a.dst + row * a.row_bytes;
    if ((a.row_bytes & 15) == 0 && (((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
        const int64_t n16 = a.row_bytes >> 4;
        for (int64_t i = (int64_t)blockIdx.y * 256 + threadIdx.x; i < n16; i += (int64_t)gridDim.y * 256)
            ((uint4*)d)[i] = ((const uint4*)s)[i];
    } else {
        for (int64_t i = (int64_t)blockIdx.y * 256 + threadIdx.x; i < a.row_bytes; i += (int64_t)gridDim.y * 256) d[i] = s[i];
    }
}

// every tensor of a minibatch in ONE launch: blockIdx.y -> (tensor, column split of its row)
constexpr int kGatherMax = 12;
struct GatherMultiArgs { const char* src[kGatherMax]; char* dst[kGatherMax]; int64_t row_bytes[kGatherMax]; int y0[kGatherMax + 1]; const int64_t* index; int64_t n; int nt; };
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) gather_rows_multi_kernel(GatherMultiArgs a) {
    int k = 0;
    while (k + 1 < a.nt && (int)blockIdx.y >= a.y0[k + 1]) ++k;
    const int64_t row = blockIdx.x, rb = a.row_bytes[k];
    const int part = (int)blockIdx.y - a.y0[k], nparts = a.y0[k + 1] - a.y0[k];
    const char* s = a.src[k] + a.index[row] * rb;
    char* d = a.dst[k] + row * rb;
    if ((rb & 15) == 0 && (((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
        const int64_t n16 = rb >> 4;
        for (int64_t i = (int64_t)part * 256 + threadIdx.x; i < n16; i += (int64_t)nparts * 256) ((uint4*)d)[i] = ((const uint4*)s)[i];
    } else {
        for (int64_t i = (int64_t)part * 256 + threadIdx.x; i < rb; i += (int64_t)nparts * 256) d[i] = s[i];
    }
}

}  // namespace cdra
