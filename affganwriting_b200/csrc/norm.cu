// Normalisation family: instance norm / AdaIN / BatchNorm(1d,2d) statistics, apply, and backward.
//
// One statistics engine serves all of them: a tensor is viewed as [G groups][P positions][C channels] (C contiguous).
//   InstanceNorm2d / AdaIN (reference blocks.py:127-128,188-204, vgg_tro_channel3_modi.py:50): G = N, P = H*W
//   BatchNorm2d in iAFF (blocks.py:250-281):  G = 1, P = N*H*W      BatchNorm1d (modules_tro.py:275,278): G = 1, P = N
//   get_key (blocks.py:218-235): G = N, P = h*w, unbiased variance, eps added before the square root
// Sums are taken about a per-channel pivot K = x[g][0][c] so that E[(x-K)^2] - E[x-K]^2 stays well conditioned in fp32.
// These kernels are HBM-bound: 128-bit accesses, fp32 math, one read for statistics and one read + one write to apply.
#include "common.cuh"

namespace {

template <int VEC> struct StatTile {
    static constexpr int TC = (VEC == 8) ? 8 : 32;  // threads across channels
    static constexpr int TP = 256 / TC;             // threads across positions
    static constexpr int CB = TC * VEC;             // channels per block
};

// ws[(g*C + c)*2 + {0,1}] += sum(x-K), sum((x-K)^2)
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
norm_stats_partial_kernel(const T* __restrict__ x, float* __restrict__ ws, long long P, int C, long long p_per_split) {
    using S = StatTile<VEC>;
    __shared__ float red[2][S::TP][S::CB + 1];
    const int tc = threadIdx.x % S::TC, tp = threadIdx.x / S::TC;
    const int c = blockIdx.x * S::CB + tc * VEC;
    const int g = blockIdx.y;
    const long long pbeg = blockIdx.z * p_per_split, pend = min(P, pbeg + p_per_split);
    const T* xg = x + (long long)g * P * C;
    float s1[VEC], s2[VEC], K[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) s1[i] = s2[i] = 0.f;
    if (c < C) {
        ldv<VEC>(xg + c, K);
        long long p = pbeg + tp;
        for (; p + S::TP < pend; p += 2 * S::TP) {          // two positions in flight per thread
            float v[VEC], w[VEC];
            ldv<VEC>(xg + p * C + c, v);
            ldv<VEC>(xg + (p + S::TP) * C + c, w);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float d = v[i] - K[i], e = w[i] - K[i];
                s1[i] += d;
                s2[i] = fmaf(d, d, s2[i]);
                s1[i] += e;
                s2[i] = fmaf(e, e, s2[i]);
            }
        }
        if (p < pend) {
            float v[VEC];
            ldv<VEC>(xg + p * C + c, v);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float d = v[i] - K[i];
                s1[i] += d;
                s2[i] = fmaf(d, d, s2[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        red[0][tp][tc * VEC + i] = s1[i];
        red[1][tp][tc * VEC + i] = s2[i];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * S::CB; j += 256) {
        const int which = j / S::CB, cc = j % S::CB;
        if (blockIdx.x * S::CB + cc >= C) continue;
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < S::TP; ++r) s += red[which][r][cc];
        atomicAdd(&ws[((long long)g * C + blockIdx.x * S::CB + cc) * 2 + which], s);
    }
}

template <typename T>
__global__ void norm_stats_finalize_kernel(const T* __restrict__ x, const float* __restrict__ ws, float* __restrict__ mean,
                                           float* __restrict__ rstd, float* __restrict__ var_unbiased, int G, long long P,
                                           int C, float eps, int unbiased) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * C) return;
    const int g = i / C, c = i % C;
    const float K = to_f(x[(long long)g * P * C + c]);
    const float inv = 1.f / (float)P;
    const float m1 = ws[2 * i] * inv;
    float var = fmaf(-m1, m1, ws[2 * i + 1] * inv);
    var = fmaxf(var, 0.f);
    const float varu = P > 1 ? var * ((float)P / (float)(P - 1)) : var;
    mean[i] = K + m1;
    rstd[i] = 1.f / sqrtf((unbiased ? varu : var) + eps);
    if (var_unbiased) var_unbiased[i] = varu;
}

// y = act( (x - mean) * rstd * gamma + beta ) + residual        gamma/beta indexed [g_or_0][c]
// Same thread tile as the statistics kernels: a thread keeps the statistics of its VEC channels in registers and walks
// positions (no index arithmetic per element); two positions are in flight per thread and the grid is sized for ~8 blocks per SM
// so that enough bytes are outstanding to cover the HBM latency.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
norm_apply_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                  const float* __restrict__ gamma, const float* __restrict__ beta, const T* __restrict__ residual,
                  T* __restrict__ y, long long P, int C, int act, int affine_per_group, long long p_per_split) {
    using S = StatTile<VEC>;
    const int tc = threadIdx.x % S::TC, tp = threadIdx.x / S::TC;
    const int c = blockIdx.x * S::CB + tc * VEC;
    const int g = blockIdx.y;
    if (c >= C) return;
    const long long pbeg = blockIdx.z * p_per_split, pend = min(P, pbeg + p_per_split);
    const long long base = (long long)g * P * C + c;
    float mu[VEC], rs[VEC], ga[VEC], be[VEC];
    const int sc = g * C + c, ac = (affine_per_group ? g * C : 0) + c;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        mu[i] = mean[sc + i];
        rs[i] = rstd[sc + i];
        ga[i] = gamma ? gamma[ac + i] : 1.f;
        be[i] = gamma ? beta[ac + i] : 0.f;
    }
    const bool affine = gamma != nullptr;
    auto one = [&](float (&v)[VEC], const float (&r)[VEC]) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float t = (v[i] - mu[i]) * rs[i];
            if (affine) t = fmaf(t, ga[i], be[i]);
            t = act_apply(t, act);
            v[i] = residual ? t + r[i] : t;
        }
    };
    long long p = pbeg + tp;
    for (; p + S::TP < pend; p += 2 * S::TP) {
        float v0[VEC], v1[VEC], r0[VEC], r1[VEC];
        const long long o0 = base + p * C, o1 = o0 + (long long)S::TP * C;
        ldv<VEC>(x + o0, v0);
        ldv<VEC>(x + o1, v1);
        if (residual) {
            ldv<VEC>(residual + o0, r0);
            ldv<VEC>(residual + o1, r1);
        }
        one(v0, r0);
        one(v1, r1);
        stv<VEC>(y + o0, v0);
        stv<VEC>(y + o1, v1);
    }
    if (p < pend) {
        float v0[VEC], r0[VEC];
        const long long o0 = base + p * C;
        ldv<VEC>(x + o0, v0);
        if (residual) ldv<VEC>(residual + o0, r0);
        one(v0, r0);
        stv<VEC>(y + o0, v0);
    }
}

// s1[g,c] += sum dy', s2[g,c] += sum dy' * xhat     with dy' = dy * act'(xhat*gamma + beta)
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
norm_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                       const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                       float* __restrict__ s1o, float* __restrict__ s2o, long long P, int C, int act,
                       int affine_per_group, long long p_per_split) {
    using S = StatTile<VEC>;
    __shared__ float red[2][S::TP][S::CB + 1];
    const int tc = threadIdx.x % S::TC, tp = threadIdx.x / S::TC;
    const int c = blockIdx.x * S::CB + tc * VEC;
    const int g = blockIdx.y;
    const long long pbeg = blockIdx.z * p_per_split, pend = min(P, pbeg + p_per_split);
    const long long base = (long long)g * P * C;
    float s1[VEC], s2[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) s1[i] = s2[i] = 0.f;
    if (c < C) {
        float mu[VEC], rs[VEC], ga[VEC], be[VEC];
        const int sc = g * C + c, ac = (affine_per_group ? g * C : 0) + c;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            mu[i] = mean[sc + i];
            rs[i] = rstd[sc + i];
            ga[i] = gamma ? gamma[ac + i] : 1.f;
            be[i] = gamma ? beta[ac + i] : 0.f;
        }
        for (long long p = pbeg + tp; p < pend; p += S::TP) {
            float v[VEC], d[VEC];
            ldv<VEC>(x + base + p * C + c, v);
            ldv<VEC>(dy + base + p * C + c, d);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float xh = (v[i] - mu[i]) * rs[i];
                const float dd = d[i] * act_grad(fmaf(xh, ga[i], be[i]), act);
                s1[i] += dd;
                s2[i] = fmaf(dd, xh, s2[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        red[0][tp][tc * VEC + i] = s1[i];
        red[1][tp][tc * VEC + i] = s2[i];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * S::CB; j += 256) {
        const int which = j / S::CB, cc = j % S::CB;
        if (blockIdx.x * S::CB + cc >= C) continue;
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < S::TP; ++r) s += red[which][r][cc];
        atomicAdd(&(which ? s2o : s1o)[(long long)g * C + blockIdx.x * S::CB + cc], s);
    }
}

// dx = rstd * gamma * ( dy' - [batch_stats] (s1 + xhat * s2) / P )
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
norm_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ s1, const float* __restrict__ s2, T* __restrict__ dx, long long P, int C, int act,
                      int affine_per_group, int batch_stats, int unbiased, long long p_per_split) {
    using S = StatTile<VEC>;
    const int tc = threadIdx.x % S::TC, tp = threadIdx.x / S::TC;
    const int c = blockIdx.x * S::CB + tc * VEC;
    const int g = blockIdx.y;
    if (c >= C) return;
    const long long pbeg = blockIdx.z * p_per_split, pend = min(P, pbeg + p_per_split);
    const long long base = (long long)g * P * C + c;
    const float invP = 1.f / (float)P;
    const float invP2 = (unbiased && P > 1) ? 1.f / (float)(P - 1) : invP;
    float mu[VEC], rs[VEC], ga[VEC], be[VEC], t1[VEC], t2[VEC];
    const int sc = g * C + c, ac = (affine_per_group ? g * C : 0) + c;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        mu[i] = mean[sc + i];
        rs[i] = rstd[sc + i];
        ga[i] = gamma ? gamma[ac + i] : 1.f;
        be[i] = gamma ? beta[ac + i] : 0.f;
        t1[i] = batch_stats ? s1[sc + i] * invP : 0.f;
        t2[i] = batch_stats ? s2[sc + i] * invP2 : 0.f;
    }
    auto one = [&](float (&v)[VEC], const float (&d)[VEC]) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float xh = (v[i] - mu[i]) * rs[i];
            float dd = d[i] * act_grad(fmaf(xh, ga[i], be[i]), act);
            if (batch_stats) dd -= t1[i] + xh * t2[i];
            v[i] = rs[i] * ga[i] * dd;
        }
    };
    long long p = pbeg + tp;
    for (; p + S::TP < pend; p += 2 * S::TP) {
        float v0[VEC], v1[VEC], d0[VEC], d1[VEC];
        const long long o0 = base + p * C, o1 = o0 + (long long)S::TP * C;
        ldv<VEC>(x + o0, v0);
        ldv<VEC>(x + o1, v1);
        ldv<VEC>(dy + o0, d0);
        ldv<VEC>(dy + o1, d1);
        one(v0, d0);
        one(v1, d1);
        stv<VEC>(dx + o0, v0);
        stv<VEC>(dx + o1, v1);
    }
    if (p < pend) {
        float v0[VEC], d0[VEC];
        const long long o0 = base + p * C;
        ldv<VEC>(x + o0, v0);
        ldv<VEC>(dy + o0, d0);
        one(v0, d0);
        stv<VEC>(dx + o0, v0);
    }
}

// BatchNorm running-statistics update (momentum form, unbiased variance) + num_batches_tracked
__global__ void bn_update_running_kernel(float* running_mean, float* running_var, long long* num_batches,
                                         const float* __restrict__ mean, const float* __restrict__ var_unbiased, int C,
                                         float momentum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean[c];
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * var_unbiased[c];
    }
    if (c == 0 && num_batches) *num_batches += 1;
}

// fixed (eval-mode) statistics -> mean / rstd arrays
__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                     float* __restrict__ mean, float* __restrict__ rstd, int C, float eps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        mean[c] = running_mean[c];
        rstd[c] = 1.f / sqrtf(running_var[c] + eps);
    }
}

inline int ew_blocks(long long total) { return (int)min((long long)148 * 8, (total + 255) / 256); }

// split the position range so that the grid has about `target` blocks (reductions: a few waves; streaming kernels: ~8 resident
// blocks per SM so that enough loads are in flight)
inline long long pick_split(int gx, int G, long long P, int& splits, long long target = 4LL * 148) {
    long long s = (target + (long long)gx * G - 1) / ((long long)gx * G);
    const long long maxs = (P + 127) / 128;
    if (s > maxs) s = maxs;
    if (s < 1) s = 1;
    const long long pps = (P + s - 1) / s;
    splits = (int)((P + pps - 1) / pps);
    return pps;
}

}  // namespace

#define DISPATCH_T_VEC(dt, C, CALL)                                  \
    do {                                                             \
        if ((dt) == AFFGW_F32) {                                     \
            if ((C) % 8 == 0) { CALL(float, 8); } else { CALL(float, 1); } \
        } else {                                                     \
            if ((C) % 8 == 0) { CALL(bf16, 8); } else { CALL(bf16, 1); }   \
        }                                                            \
    } while (0)

int norm_stats(const void* x, int dt, float* ws, float* mean, float* rstd, float* var_unbiased, int G, long long P, int C,
               float eps, int unbiased, cudaStream_t st) {
    if (cudaMemsetAsync(ws, 0, sizeof(float) * 2 * G * C, st) != cudaSuccess) {
        affgw_set_error("norm_stats: memset failed");
        return -2;
    }
    const int vec = (C % 8 == 0) ? 8 : 1;
    const int cb = vec == 8 ? 64 : 32;
    int splits;
    const long long pps = pick_split(cdiv(C, cb), G, P, splits);
    dim3 grid(cdiv(C, cb), G, splits);
#define CALL(T, V) norm_stats_partial_kernel<T, V><<<grid, 256, 0, st>>>((const T*)x, ws, P, C, pps)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("norm_stats_partial");
    if (dt == AFFGW_F32)
        norm_stats_finalize_kernel<float><<<cdiv(G * C, 256), 256, 0, st>>>((const float*)x, ws, mean, rstd, var_unbiased, G, P, C, eps, unbiased);
    else
        norm_stats_finalize_kernel<bf16><<<cdiv(G * C, 256), 256, 0, st>>>((const bf16*)x, ws, mean, rstd, var_unbiased, G, P, C, eps, unbiased);
    AFFGW_LAUNCH_CHECK("norm_stats_finalize");
    return 0;
}

int norm_apply(const void* x, int dt, const float* mean, const float* rstd, const float* gamma, const float* beta,
               const void* residual, void* y, int G, long long P, int C, int act, int affine_per_group, cudaStream_t st) {
    const int cb = (C % 8 == 0) ? 64 : 32;
    int splits;
    const long long pps = pick_split(cdiv(C, cb), G, P, splits, 8LL * 148);
    dim3 grid(cdiv(C, cb), G, splits);
#define CALL(T, V) norm_apply_kernel<T, V><<<grid, 256, 0, st>>>((const T*)x, mean, rstd, gamma, beta, (const T*)residual, (T*)y, P, C, act, affine_per_group, pps)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("norm_apply");
    return 0;
}

int norm_bwd(const void* dy, const void* x, int dt, const float* mean, const float* rstd, const float* gamma,
             const float* beta, float* s1, float* s2, void* dx, int G, long long P, int C, int act, int affine_per_group,
             int batch_stats, int unbiased, cudaStream_t st) {
    if (cudaMemsetAsync(s1, 0, sizeof(float) * G * C, st) != cudaSuccess ||
        cudaMemsetAsync(s2, 0, sizeof(float) * G * C, st) != cudaSuccess) {
        affgw_set_error("norm_bwd: memset failed");
        return -2;
    }
    const int vec = (C % 8 == 0) ? 8 : 1;
    const int cb = vec == 8 ? 64 : 32;
    int splits;
    const long long pps = pick_split(cdiv(C, cb), G, P, splits);
    dim3 grid(cdiv(C, cb), G, splits);
#define CALL(T, V) norm_bwd_reduce_kernel<T, V><<<grid, 256, 0, st>>>((const T*)dy, (const T*)x, mean, rstd, gamma, beta, s1, s2, P, C, act, affine_per_group, pps)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("norm_bwd_reduce");
    int asplits;
    const long long apps = pick_split(cdiv(C, cb), G, P, asplits, 8LL * 148);
    dim3 agrid(cdiv(C, cb), G, asplits);
#define CALL(T, V) norm_bwd_apply_kernel<T, V><<<agrid, 256, 0, st>>>((const T*)dy, (const T*)x, mean, rstd, gamma, beta, s1, s2, (T*)dx, P, C, act, affine_per_group, batch_stats, unbiased, apps)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("norm_bwd_apply");
    return 0;
}

int bn_update_running(float* rm, float* rv, long long* nbt, const float* mean, const float* var_unbiased, int C,
                      float momentum, cudaStream_t st) {
    bn_update_running_kernel<<<cdiv(C, 256), 256, 0, st>>>(rm, rv, nbt, mean, var_unbiased, C, momentum);
    AFFGW_LAUNCH_CHECK("bn_update_running");
    return 0;
}

int bn_eval_stats(const float* rm, const float* rv, float* mean, float* rstd, int C, float eps, cudaStream_t st) {
    bn_eval_stats_kernel<<<cdiv(C, 256), 256, 0, st>>>(rm, rv, mean, rstd, C, eps);
    AFFGW_LAUNCH_CHECK("bn_eval_stats");
    return 0;
}
