// HBM-bound pointwise / small-reduction kernels of the generator and discriminator path (NHWC, 128-bit accesses).
//   max-pool 2x2            vgg_tro_channel3_modi.py:45, modules_tro.py:224
//   reflect-pad avg-pool    modules_tro.py:133-134
//   iAFF gate               blocks.py:286-299          (x*w + r*(1-w), w = sigmoid(local + global))
//   global average pool     blocks.py:255-256
//   nearest resize          blocks.py:214
//   text tiling / embedding modules_tro.py:285-317
//   BCE-with-logits / CE    modules_tro.py:152-168,195-201
#include "common.cuh"

namespace {

inline int ew_blocks(long long total) { return (int)max(1LL, min((long long)148 * 8, (total + 255) / 256)); }

#define GRID_STRIDE(idx, total)                                                                  \
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (total);        \
         idx += (long long)gridDim.x * blockDim.x)

// ---------------------------------------------------------------- max pool 2x2 stride 2 (floor)
template <typename T, int VEC>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
    const int Ho = H / 2, Wo = W / 2, cv = C / VEC;
    const long long total = (long long)N * Ho * Wo * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        long long t = idx / cv;
        const int ox = (int)(t % Wo);
        t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const T* p = x + (((long long)n * H + 2 * oy) * W + 2 * ox) * C + c;
        float a[VEC], b[VEC];
        ldv<VEC>(p, a);
        ldv<VEC>(p + C, b);
#pragma unroll
        for (int i = 0; i < VEC; ++i) a[i] = fmaxf(a[i], b[i]);
        ldv<VEC>(p + (long long)W * C, b);
#pragma unroll
        for (int i = 0; i < VEC; ++i) a[i] = fmaxf(a[i], b[i]);
        ldv<VEC>(p + (long long)W * C + C, b);
#pragma unroll
        for (int i = 0; i < VEC; ++i) a[i] = fmaxf(a[i], b[i]);
        stv<VEC>(y + idx * VEC, a);
    }
}

// gather form: the first maximum in window scan order receives the gradient (PyTorch tie rule)
template <typename T, int VEC>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, int N, int H,
                                    int W, int C) {
    const int Ho = H / 2, Wo = W / 2, cv = C / VEC;
    const long long total = (long long)N * H * W * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        long long t = idx / cv;
        const int xx = (int)(t % W);
        t /= W;
        const int yy = (int)(t % H);
        const int n = (int)(t / H);
        float out[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) out[i] = 0.f;
        const int oy = yy >> 1, ox = xx >> 1;
        if (oy < Ho && ox < Wo) {
            const T* p = x + (((long long)n * H + 2 * oy) * W + 2 * ox) * C + c;
            float w4[4][VEC], g[VEC];
            ldv<VEC>(p, w4[0]);
            ldv<VEC>(p + C, w4[1]);
            ldv<VEC>(p + (long long)W * C, w4[2]);
            ldv<VEC>(p + (long long)W * C + C, w4[3]);
            ldv<VEC>(dy + (((long long)n * Ho + oy) * Wo + ox) * C + c, g);
            const int me = (yy & 1) * 2 + (xx & 1);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                int best = 0;
                float bv = w4[0][i];
#pragma unroll
                for (int k = 1; k < 4; ++k)
                    if (w4[k][i] > bv) { bv = w4[k][i]; best = k; }
                out[i] = (best == me) ? g[i] : 0.f;
            }
        }
        stv<VEC>(dx + idx * VEC, out);
    }
}

// ---------------------------------------------------------------- ReflectionPad2d(1) + AvgPool2d(3, stride 2)
__device__ __forceinline__ int refl(int v, int V) { return v < 0 ? -v : (v >= V ? 2 * (V - 1) - v : v); }

template <typename T, int VEC>
__global__ void avgpool3s2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int Ho,
                                      int Wo) {
    const int cv = C / VEC;
    const long long total = (long long)N * Ho * Wo * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        long long t = idx / cv;
        const int ox = (int)(t % Wo);
        t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int sy = refl(2 * oy + ky - 1, H);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int sx = refl(2 * ox + kx - 1, W);
                float v[VEC];
                ldv<VEC>(x + (((long long)n * H + sy) * W + sx) * C + c, v);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] += v[i];
            }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] *= (1.f / 9.f);
        stv<VEC>(y + idx * VEC, acc);
    }
}

template <typename T, int VEC>
__global__ void avgpool3s2_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C, int Ho,
                                      int Wo) {
    const int cv = C / VEC;
    const long long total = (long long)N * H * W * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        long long t = idx / cv;
        const int xx = (int)(t % W);
        t /= W;
        const int yy = (int)(t % H);
        const int n = (int)(t / H);
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        const int oy0 = max(0, (yy - 1) / 2 - 1), oy1 = min(Ho - 1, (yy + 1) / 2 + 1);
        const int ox0 = max(0, (xx - 1) / 2 - 1), ox1 = min(Wo - 1, (xx + 1) / 2 + 1);
        for (int oy = oy0; oy <= oy1; ++oy) {
            int cy = 0;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) cy += (refl(2 * oy + ky - 1, H) == yy);
            if (!cy) continue;
            for (int ox = ox0; ox <= ox1; ++ox) {
                int cx = 0;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) cx += (refl(2 * ox + kx - 1, W) == xx);
                if (!cx) continue;
                float g[VEC];
                ldv<VEC>(dy + (((long long)n * Ho + oy) * Wo + ox) * C + c, g);
                const float wgt = (float)(cy * cx) * (1.f / 9.f);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fmaf(g[i], wgt, acc[i]);
            }
        }
        stv<VEC>(dx + idx * VEC, acc);
    }
}

// Row-per-block form of the kernel above (fp32, 8-channel vectors): the candidate output rows of an input row and their
// multiplicities are block-uniform; no 64-bit index arithmetic per element.
template <int VEC>
__global__ void __launch_bounds__(256)
avgpool3s2_bwd_rows_kernel(const float* __restrict__ dy, float* __restrict__ dx, int H, int W, int C, int Ho, int Wo) {
    const int cv = C / VEC;
    const int row = blockIdx.y, n = row / H, yy = row - n * H;
    const int oyb = max(0, (yy - 1) / 2 - 1);
    int cy[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int oy = oyb + j;
        cy[j] = 0;
        if (oy < Ho) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) cy[j] += (refl(2 * oy + ky - 1, H) == yy);
        }
    }
    const float* const img = dy + (long long)n * Ho * Wo * C;
    const int items = W * cv;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < items; i += gridDim.x * 256) {
        const int xx = i / cv, c = (i - xx * cv) * VEC;
        const int oxb = max(0, (xx - 1) / 2 - 1);
        int cx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ox = oxb + k;
            cx[k] = 0;
            if (ox < Wo) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) cx[k] += (refl(2 * ox + kx - 1, W) == xx);
            }
        }
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!cy[j]) continue;                                   // block-uniform
            const float* const rp = img + (long long)(oyb + j) * Wo * C + c;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!cx[k]) continue;
                float g[VEC];
                ldv<VEC>(rp + (oxb + k) * C, g);
                const float wgt = (float)(cy[j] * cx[k]) * (1.f / 9.f);
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[e] = fmaf(g[e], wgt, acc[e]);
            }
        }
        stv<VEC>(dx + ((long long)row * W + xx) * C + c, acc);
    }
}

// ---------------------------------------------------------------- iAFF gate
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

template <typename T, int VEC>
__global__ void gate_fwd_kernel(const T* __restrict__ x, const T* __restrict__ r, const T* __restrict__ xl,
                                const T* __restrict__ xg, T* __restrict__ y, int N, long long P, int C) {
    const int cv = C / VEC;
    const long long total = (long long)N * P * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        const int n = (int)(idx / cv / P);
        float a[VEC], b[VEC], l[VEC], gl[VEC];
        ldv<VEC>(x + idx * VEC, a);
        ldv<VEC>(r + idx * VEC, b);
        ldv<VEC>(xl + idx * VEC, l);
        ldv<VEC>(xg + (long long)n * C + c, gl);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float w = 1.f / (1.f + expf(-(l[i] + gl[i])));
            a[i] = a[i] * w + b[i] * (1.f - w);
        }
        stv<VEC>(y + idx * VEC, a);
    }
}

// grid (C blocks, N): dx, dr, dxl elementwise; dxg[n,c] = sum_p dz  (block-local reduction, no atomics)
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
gate_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ r, const T* __restrict__ xl,
                const T* __restrict__ xg, T* __restrict__ dx, T* __restrict__ dr, T* __restrict__ dxl, T* __restrict__ dxg,
                long long P, int C) {
    constexpr int TC = (VEC == 8) ? 8 : 32, TP = 256 / TC, CB = TC * VEC;
    __shared__ float red[TP][CB + 1];
    const int tc = threadIdx.x % TC, tp = threadIdx.x / TC;
    const int c = blockIdx.x * CB + tc * VEC, n = blockIdx.y;
    float s[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) s[i] = 0.f;
    if (c < C) {
        float gl[VEC];
        ldv<VEC>(xg + (long long)n * C + c, gl);
        for (long long p = tp; p < P; p += TP) {
            const long long o = ((long long)n * P + p) * C + c;
            float g[VEC], a[VEC], b[VEC], l[VEC], ox[VEC], orr[VEC], ol[VEC];
            ldv<VEC>(dy + o, g);
            ldv<VEC>(x + o, a);
            ldv<VEC>(r + o, b);
            ldv<VEC>(xl + o, l);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float w = 1.f / (1.f + expf(-(l[i] + gl[i])));
                ox[i] = g[i] * w;
                orr[i] = g[i] * (1.f - w);
                ol[i] = g[i] * (a[i] - b[i]) * w * (1.f - w);
                s[i] += ol[i];
            }
            stv<VEC>(dx + o, ox);
            stv<VEC>(dr + o, orr);
            stv<VEC>(dxl + o, ol);
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) red[tp][tc * VEC + i] = s[i];
    __syncthreads();
    if (threadIdx.x < CB && blockIdx.x * CB + threadIdx.x < C) {
        float t = 0.f;
        for (int k = 0; k < TP; ++k) t += red[k][threadIdx.x];
        dxg[(long long)n * C + blockIdx.x * CB + threadIdx.x] = from_f<T>(t);
    }
}

// global average pool: out[n,c] = mean_p x[n,p,c]     grid (C blocks, N)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) gap_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, long long P, int C) {
    constexpr int TC = (VEC == 8) ? 8 : 32, TP = 256 / TC, CB = TC * VEC;
    __shared__ float red[TP][CB + 1];
    const int tc = threadIdx.x % TC, tp = threadIdx.x / TC;
    const int c = blockIdx.x * CB + tc * VEC, n = blockIdx.y;
    float s[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) s[i] = 0.f;
    if (c < C)
        for (long long p = tp; p < P; p += TP) {
            float v[VEC];
            ldv<VEC>(x + ((long long)n * P + p) * C + c, v);
#pragma unroll
            for (int i = 0; i < VEC; ++i) s[i] += v[i];
        }
#pragma unroll
    for (int i = 0; i < VEC; ++i) red[tp][tc * VEC + i] = s[i];
    __syncthreads();
    if (threadIdx.x < CB && blockIdx.x * CB + threadIdx.x < C) {
        float t = 0.f;
        for (int k = 0; k < TP; ++k) t += red[k][threadIdx.x];
        out[(long long)n * C + blockIdx.x * CB + threadIdx.x] = from_f<T>(t / (float)P);
    }
}

// out[n,p,c] = a[n,p,c] (optional) + v[n,c] * scale
template <typename T, int VEC>
__global__ void bcast_add_kernel(const T* __restrict__ a, const T* __restrict__ v, T* __restrict__ out, int N, long long P,
                                 int C, float scale) {
    const int cv = C / VEC;
    const long long total = (long long)N * P * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        const int n = (int)(idx / cv / P);
        float o[VEC], vv[VEC];
        ldv<VEC>(v + (long long)n * C + c, vv);
        if (a) ldv<VEC>(a + idx * VEC, o);
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = (a ? o[i] : 0.f) + vv[i] * scale;
        stv<VEC>(out + idx * VEC, o);
    }
}

template <typename T, int VEC>
__global__ void add2_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long nvec) {
    GRID_STRIDE(idx, nvec) {
        float x[VEC], y[VEC];
        ldv<VEC>(a + idx * VEC, x);
        ldv<VEC>(b + idx * VEC, y);
#pragma unroll
        for (int i = 0; i < VEC; ++i) x[i] += y[i];
        stv<VEC>(out + idx * VEC, x);
    }
}


// ---------------------------------------------------------------- max pool 3x3 pad 1, strides (SY, SX) in {1, 2}
// (2,2): torchvision ResNet stem; (2,1) and (1,1): Resnet18.py:45-46
template <typename T, int VEC>
__global__ void maxpool3s2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int Ho, int Wo,
                                      int SY, int SX) {
    const int cv = C / VEC;
    const long long total = (long long)N * Ho * Wo * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        long long t = idx / cv;
        const int ox = (int)(t % Wo);
        t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float m[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) m[i] = -INFINITY;
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = SY * oy - 1 + ky;
            if (yy < 0 || yy >= H) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = SX * ox - 1 + kx;
                if (xx < 0 || xx >= W) continue;
                float v[VEC];
                ldv<VEC>(x + (((long long)n * H + yy) * W + xx) * C + c, v);
#pragma unroll
                for (int i = 0; i < VEC; ++i) m[i] = fmaxf(m[i], v[i]);
            }
        }
        stv<VEC>(y + idx * VEC, m);
    }
}

// gather form: input pixel (yy, xx) receives dy of every window whose FIRST maximum in scan order it is (PyTorch tie rule)
template <typename T, int VEC>
__global__ void maxpool3s2_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, int N, int H, int W,
                                      int C, int Ho, int Wo, int SY, int SX) {
    const int cv = C / VEC;
    const long long total = (long long)N * H * W * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        long long t = idx / cv;
        const int xx = (int)(t % W);
        t /= W;
        const int yy = (int)(t % H);
        const int n = (int)(t / H);
        float mine[VEC], out[VEC];
        ldv<VEC>(x + idx * VEC, mine);
#pragma unroll
        for (int i = 0; i < VEC; ++i) out[i] = 0.f;
        // windows (oy, ox) with S*o - 1 <= coordinate <= S*o + 1, i.e. ceil((c - 1) / S) <= o <= floor((c + 1) / S)
        const int oy_lo = yy <= 1 ? 0 : (yy - 1 + SY - 1) / SY, oy_hi = min((yy + 1) / SY, Ho - 1);
        const int ox_lo = xx <= 1 ? 0 : (xx - 1 + SX - 1) / SX, ox_hi = min((xx + 1) / SX, Wo - 1);
        for (int oy = oy_lo; oy <= oy_hi; ++oy) {
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                const int my_pos = (yy - (SY * oy - 1)) * 3 + (xx - (SX * ox - 1));
                bool win[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) win[i] = true;
                for (int ky = 0; ky < 3; ++ky) {
                    const int y2 = SY * oy - 1 + ky;
                    if (y2 < 0 || y2 >= H) continue;
                    for (int kx = 0; kx < 3; ++kx) {
                        const int x2 = SX * ox - 1 + kx;
                        if (x2 < 0 || x2 >= W) continue;
                        const int pos = ky * 3 + kx;
                        if (pos == my_pos) continue;
                        float v[VEC];
                        ldv<VEC>(x + (((long long)n * H + y2) * W + x2) * C + c, v);
#pragma unroll
                        for (int i = 0; i < VEC; ++i)      // an earlier element wins ties, a later one must be strictly greater
                            if (pos < my_pos ? v[i] >= mine[i] : v[i] > mine[i]) win[i] = false;
                    }
                }
                float g[VEC];
                ldv<VEC>(dy + (((long long)n * Ho + oy) * Wo + ox) * C + c, g);
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                    if (win[i]) out[i] += g[i];
            }
        }
        stv<VEC>(dx + idx * VEC, out);
    }
}

// ---------------------------------------------------------------- bilinear resize, align_corners = False (F.interpolate)
__device__ __forceinline__ void bilinear_src(int dst, int in, int out, int& i0, int& i1, float& w1) {
    float s = ((float)dst + 0.5f) * ((float)in / (float)out) - 0.5f;
    if (s < 0.f) s = 0.f;
    i0 = min((int)s, in - 1);
    i1 = min(i0 + 1, in - 1);
    w1 = s - (float)i0;
}
template <typename T, int VEC>
__global__ void resize_bilinear_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int Ho, int Wo) {
    const int cv = C / VEC;
    const long long total = (long long)N * Ho * Wo * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        long long t = idx / cv;
        const int ox = (int)(t % Wo);
        t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        int y0, y1, x0, x1;
        float wy, wx;
        bilinear_src(oy, H, Ho, y0, y1, wy);
        bilinear_src(ox, W, Wo, x0, x1, wx);
        const T* base = x + (long long)n * H * W * C + c;
        float a[VEC], b[VEC], d[VEC], e[VEC], o[VEC];
        ldv<VEC>(base + ((long long)y0 * W + x0) * C, a);
        ldv<VEC>(base + ((long long)y0 * W + x1) * C, b);
        ldv<VEC>(base + ((long long)y1 * W + x0) * C, d);
        ldv<VEC>(base + ((long long)y1 * W + x1) * C, e);
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            o[i] = (1.f - wy) * ((1.f - wx) * a[i] + wx * b[i]) + wy * ((1.f - wx) * d[i] + wx * e[i]);
        stv<VEC>(y + idx * VEC, o);
    }
}
// scatter form with fp32 atomics (dx zeroed by the caller; maps are tiny: 2x7 -> 8x27)
template <typename T>
__global__ void resize_bilinear_bwd_kernel(const T* __restrict__ dy, float* __restrict__ dx, int N, int H, int W, int C, int Ho,
                                           int Wo) {
    const long long total = (long long)N * Ho * Wo * C;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % C);
        long long t = idx / C;
        const int ox = (int)(t % Wo);
        t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        int y0, y1, x0, x1;
        float wy, wx;
        bilinear_src(oy, H, Ho, y0, y1, wy);
        bilinear_src(ox, W, Wo, x0, x1, wx);
        const float g = to_f(dy[idx]);
        float* base = dx + (long long)n * H * W * C + c;
        atomicAdd(base + ((long long)y0 * W + x0) * C, g * (1.f - wy) * (1.f - wx));
        atomicAdd(base + ((long long)y0 * W + x1) * C, g * (1.f - wy) * wx);
        atomicAdd(base + ((long long)y1 * W + x0) * C, g * wy * (1.f - wx));
        atomicAdd(base + ((long long)y1 * W + x1) * C, g * wy * wx);
    }
}

// ---------------------------------------------------------------- out = act(a + b)   (BasicBlock / Bottleneck tail)
template <typename T, int VEC>
__global__ void add_act_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long nvec, int act) {
    GRID_STRIDE(idx, nvec) {
        float x[VEC], y[VEC];
        ldv<VEC>(a + idx * VEC, x);
        ldv<VEC>(b + idx * VEC, y);
#pragma unroll
        for (int i = 0; i < VEC; ++i) x[i] = act_apply(x[i] + y[i], act);
        stv<VEC>(out + idx * VEC, x);
    }
}

// ---------------------------------------------------------------- nearest resize (F.interpolate default mode)
__device__ __forceinline__ int nearest_src(int dst, int in, int out) {
    const float scale = (float)in / (float)out;
    return min((int)floorf((float)dst * scale), in - 1);
}

template <typename T, int VEC>
__global__ void resize_nearest_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int Ho,
                                          int Wo) {
    const int cv = C / VEC;
    const long long total = (long long)N * Ho * Wo * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        long long t = idx / cv;
        const int ox = (int)(t % Wo);
        t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float v[VEC];
        ldv<VEC>(x + (((long long)n * H + nearest_src(oy, H, Ho)) * W + nearest_src(ox, W, Wo)) * C + c, v);
        stv<VEC>(y + idx * VEC, v);
    }
}

template <typename T, int VEC>
__global__ void resize_nearest_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C, int Ho,
                                          int Wo) {
    const int cv = C / VEC;
    const long long total = (long long)N * H * W * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        long long t = idx / cv;
        const int xx = (int)(t % W);
        t /= W;
        const int yy = (int)(t % H);
        const int n = (int)(t / H);
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        const int oy0 = max(0, (int)((long long)yy * Ho / H) - 1), oy1 = min(Ho - 1, (int)((long long)(yy + 1) * Ho / H) + 1);
        const int ox0 = max(0, (int)((long long)xx * Wo / W) - 1), ox1 = min(Wo - 1, (int)((long long)(xx + 1) * Wo / W) + 1);
        for (int oy = oy0; oy <= oy1; ++oy) {
            if (nearest_src(oy, H, Ho) != yy) continue;
            for (int ox = ox0; ox <= ox1; ++ox) {
                if (nearest_src(ox, W, Wo) != xx) continue;
                float g[VEC];
                ldv<VEC>(dy + (((long long)n * Ho + oy) * Wo + ox) * C + c, g);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] += g[i];
            }
        }
        stv<VEC>(dx + idx * VEC, acc);
    }
}

// ---------------------------------------------------------------- text: embedding + column tiling
template <typename T>
__global__ void embedding_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ table, T* __restrict__ out,
                                     long long n_ids, int E, int V, int* __restrict__ err) {
    GRID_STRIDE(idx, n_ids * E) {
        const long long id = ids[idx / E];
        if (id < 0 || id >= V) { *err = 1; continue; }
        out[idx] = from_f<T>(table[id * E + idx % E]);
    }
}
template <typename T>
__global__ void embedding_bwd_kernel(const long long* __restrict__ ids, const T* __restrict__ dout, float* __restrict__ dtable,
                                     long long n_ids, int E, int V) {
    GRID_STRIDE(idx, n_ids * E) {
        const long long id = ids[idx / E];
        if (id >= 0 && id < V) atomicAdd(&dtable[id * E + idx % E], to_f(dout[idx]));
    }
}

// column -> token slot (modules_tro.py:295-313): ts tokens each repeated reps times, then PAD slot (index ts)
__device__ __forceinline__ int text_slot(int col, int ts, int reps) { return col < ts * reps ? col / reps : ts; }

template <typename T, int VEC>
__global__ void text_tile_fwd_kernel(const T* __restrict__ chars, T* __restrict__ out, int B, int H, int W, int C, int ts,
                                     int reps) {
    const int cv = C / VEC;
    const long long total = (long long)B * H * W * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        long long t = idx / cv;
        const int col = (int)(t % W);
        const int b = (int)(t / W / H);
        float v[VEC];
        ldv<VEC>(chars + ((long long)b * (ts + 1) + text_slot(col, ts, reps)) * C + c, v);
        stv<VEC>(out + idx * VEC, v);
    }
}
template <typename T, int VEC>
__global__ void text_tile_bwd_kernel(const T* __restrict__ dout, T* __restrict__ dchars, int B, int H, int W, int C, int ts,
                                     int reps) {
    const int cv = C / VEC;
    const long long total = (long long)B * (ts + 1) * cv;
    GRID_STRIDE(idx, total) {
        const int c = (int)(idx % cv) * VEC;
        const int slot = (int)((idx / cv) % (ts + 1));
        const int b = (int)(idx / cv / (ts + 1));
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        const int c0 = slot < ts ? slot * reps : ts * reps, c1 = slot < ts ? min(W, c0 + reps) : W;
        for (int h = 0; h < H; ++h)
            for (int col = c0; col < c1; ++col) {
                float g[VEC];
                ldv<VEC>(dout + (((long long)b * H + h) * W + col) * C + c, g);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] += g[i];
            }
        stv<VEC>(dchars + idx * VEC, acc);
    }
}

// ---------------------------------------------------------------- losses (mean reduction)
template <typename T>
__global__ void __launch_bounds__(256) bce_logits_fwd_kernel(const T* __restrict__ x, float target, float* __restrict__ loss,
                                                             long long n) {
    __shared__ float red[8];
    float s = 0.f;
    GRID_STRIDE(idx, n) {
        const float v = to_f(x[idx]);
        s += fmaxf(v, 0.f) - v * target + log1pf(expf(-fabsf(v)));
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        atomicAdd(loss, t / (float)n);
    }
}
template <typename T>
__global__ void bce_logits_bwd_kernel(const T* __restrict__ x, float target, const float* __restrict__ gout, T* __restrict__ dx,
                                      long long n) {
    const float g = gout[0] / (float)n;
    GRID_STRIDE(idx, n) {
        const float v = to_f(x[idx]);
        dx[idx] = from_f<T>((1.f / (1.f + expf(-v)) - target) * g);
    }
}

// one warp per row
template <typename T>
__global__ void softmax_ce_fwd_kernel(const T* __restrict__ x, const long long* __restrict__ y, float* __restrict__ loss, int B,
                                      int C, int* __restrict__ err) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const T* xr = x + (long long)row * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, to_f(xr[c]));
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(to_f(xr[c]) - mx);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const long long t = y[row];
        if (t < 0 || t >= C) { *err = 1; return; }
        atomicAdd(loss, (logf(s) + mx - to_f(xr[t])) / (float)B);
    }
}
template <typename T>
__global__ void softmax_ce_bwd_kernel(const T* __restrict__ x, const long long* __restrict__ y, const float* __restrict__ gout,
                                      T* __restrict__ dx, int B, int C) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const T* xr = x + (long long)row * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, to_f(xr[c]));
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(to_f(xr[c]) - mx);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float g = gout[0] / (float)B, inv = 1.f / s;
    const long long t = y[row];
    for (int c = lane; c < C; c += 32)
        dx[(long long)row * C + c] = from_f<T>((expf(to_f(xr[c]) - mx) * inv - (c == t ? 1.f : 0.f)) * g);
}

// ---------------------------------------------------------------- uint8 wire format -> normalised image
// load_data.py:147-166 after the resize: img/255. and 1. - img in float64 (numpy promotes uint8 / float), stored into a
// float32 canvas, then (canvas - 0.5) / 0.5 in float32.  256 possible results: a per-block table built with exactly those
// operations, then one table look-up per byte (16 bytes in, 64 bytes out per thread and iteration).
__global__ void u8_to_image_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, long long n) {
    __shared__ float lut[256];
    {
        const double ink = 1.0 - (double)threadIdx.x / 255.0;
        const float canvas = (float)ink;
        lut[threadIdx.x] = (canvas - 0.5f) / 0.5f;
    }
    __syncthreads();
    const long long vecs = n / 16;
    GRID_STRIDE(i, vecs) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 o;
            o.x = lut[w[k] & 255u];
            o.y = lut[(w[k] >> 8) & 255u];
            o.z = lut[(w[k] >> 16) & 255u];
            o.w = lut[w[k] >> 24];
            reinterpret_cast<float4*>(dst)[i * 4 + k] = o;
        }
    }
    if (blockIdx.x == 0)
        for (long long i = vecs * 16 + threadIdx.x; i < n; i += blockDim.x) dst[i] = lut[src[i]];
}

// ---------------------------------------------------------------- layout / dtype conversion
// NCHW fp32 -> NHWC T with the channel dimension zero-padded to cpad (tile-transposed through shared memory)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int C, long long HW, int cpad) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j;
        const long long p = p0 + tx;
        tile[j][tx] = (c < C && p < HW) ? x[((long long)n * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const long long p = p0 + j;
        const int c = c0 + tx;
        if (p < HW && c < cpad) y[((long long)n * HW + p) * cpad + c] = from_f<T>(tile[tx][j]);
    }
}
// NHWC T (pitch cpad) -> NCHW fp32 (first C channels)
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int C, long long HW, int cpad) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const long long p = p0 + j;
        const int c = c0 + tx;
        tile[j][tx] = (p < HW && c < C) ? to_f(x[((long long)n * HW + p) * cpad + c]) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j;
        const long long p = p0 + tx;
        if (c < C && p < HW) y[((long long)n * C + c) * HW + p] = tile[tx][j];
    }
}
// dz = dy * act'(.) evaluated from the activation OUTPUT y (relu / lrelu by sign, tanh by 1 - y^2)
template <typename T, int VEC>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dz, long long nvec, int act) {
    GRID_STRIDE(idx, nvec) {
        float g[VEC], o[VEC];
        ldv<VEC>(dy + idx * VEC, g);
        ldv<VEC>(y + idx * VEC, o);
#pragma unroll
        for (int i = 0; i < VEC; ++i) g[i] *= (act == ACT_TANH) ? (1.f - o[i] * o[i]) : act_grad(o[i], act);
        stv<VEC>(dz + idx * VEC, g);
    }
}
// channel concatenation / split of NHWC tensors (torch.cat(dim=1) at modules_tro.py:256): rows = N*H*W
template <typename T, int VEC>
__global__ void concat2_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long rows, int ca,
                               int cb, int to_out) {
    const int ct = ca + cb, cv = ct / VEC;
    GRID_STRIDE(idx, rows * cv) {
        const int c = (int)(idx % cv) * VEC;
        const long long r = idx / cv;
        float v[VEC];
        T* big = out + r * ct + c;
        T* small_ = (c < ca) ? const_cast<T*>(a) + r * ca + c : const_cast<T*>(b) + r * cb + (c - ca);
        if (to_out) { ldv<VEC>(small_, v); stv<VEC>(big, v); }
        else { ldv<VEC>(big, v); stv<VEC>(small_, v); }
    }
}
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n) {
    GRID_STRIDE(idx, n) y[idx] = from_f<TO>(to_f(x[idx]));
}

}  // namespace

#define DISPATCH_T_VEC(dt, C, CALL)                                        \
    do {                                                                   \
        if ((dt) == AFFGW_F32) {                                           \
            if ((C) % 8 == 0) { CALL(float, 8); } else { CALL(float, 1); } \
        } else {                                                           \
            if ((C) % 8 == 0) { CALL(bf16, 8); } else { CALL(bf16, 1); }   \
        }                                                                  \
    } while (0)
#define VECN(C) ((C) % 8 == 0 ? (C) / 8 : (C))

int maxpool2_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, cudaStream_t st) {
    const long long total = (long long)N * (H / 2) * (W / 2) * VECN(C);
#define CALL(T, V) maxpool2_fwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("maxpool2_fwd");
    return 0;
}
int maxpool2_bwd(const void* dy, const void* x, void* dx, int dt, int N, int H, int W, int C, cudaStream_t st) {
    const long long total = (long long)N * H * W * VECN(C);
#define CALL(T, V) maxpool2_bwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)dy, (const T*)x, (T*)dx, N, H, W, C)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("maxpool2_bwd");
    return 0;
}
int avgpool3s2_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, cudaStream_t st) {
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long total = (long long)N * Ho * Wo * VECN(C);
#define CALL(T, V) avgpool3s2_fwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, Ho, Wo)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("avgpool3s2_fwd");
    return 0;
}
int avgpool3s2_bwd(const void* dy, void* dx, int dt, int N, int H, int W, int C, cudaStream_t st) {
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long total = (long long)N * H * W * VECN(C);
    if (dt == AFFGW_F32 && C % 8 == 0 && (long long)N * H <= 65535) {
        const int items = W * (C / 8);
        dim3 grid((unsigned)min(8, (items + 255) / 256), (unsigned)(N * H));
        avgpool3s2_bwd_rows_kernel<8><<<grid, 256, 0, st>>>((const float*)dy, (float*)dx, H, W, C, Ho, Wo);
        AFFGW_LAUNCH_CHECK("avgpool3s2_bwd");
        return 0;
    }
#define CALL(T, V) avgpool3s2_bwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)dy, (T*)dx, N, H, W, C, Ho, Wo)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("avgpool3s2_bwd");
    return 0;
}
int gate_fwd(const void* x, const void* r, const void* xl, const void* xg, void* y, int dt, int N, long long P, int C,
             cudaStream_t st) {
    const long long total = (long long)N * P * VECN(C);
#define CALL(T, V) gate_fwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)x, (const T*)r, (const T*)xl, (const T*)xg, (T*)y, N, P, C)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("gate_fwd");
    return 0;
}
int gate_bwd(const void* dy, const void* x, const void* r, const void* xl, const void* xg, void* dx, void* dr, void* dxl,
             void* dxg, int dt, int N, long long P, int C, cudaStream_t st) {
    dim3 grid(cdiv(C, C % 8 == 0 ? 64 : 32), N);
#define CALL(T, V) gate_bwd_kernel<T, V><<<grid, 256, 0, st>>>((const T*)dy, (const T*)x, (const T*)r, (const T*)xl, (const T*)xg, (T*)dx, (T*)dr, (T*)dxl, (T*)dxg, P, C)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("gate_bwd");
    return 0;
}
int gap_fwd(const void* x, void* out, int dt, int N, long long P, int C, cudaStream_t st) {
    dim3 grid(cdiv(C, C % 8 == 0 ? 64 : 32), N);
#define CALL(T, V) gap_fwd_kernel<T, V><<<grid, 256, 0, st>>>((const T*)x, (T*)out, P, C)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("gap_fwd");
    return 0;
}
int bcast_add(const void* a, const void* v, void* out, int dt, int N, long long P, int C, float scale, cudaStream_t st) {
    const long long total = (long long)N * P * VECN(C);
#define CALL(T, V) bcast_add_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)a, (const T*)v, (T*)out, N, P, C, scale)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("bcast_add");
    return 0;
}

int maxpool3_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, int SY, int SX, cudaStream_t st) {
    const int Ho = (H - 1) / SY + 1, Wo = (W - 1) / SX + 1;
    const long long total = (long long)N * Ho * Wo * VECN(C);
#define CALL(T, V) maxpool3s2_fwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, Ho, Wo, SY, SX)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("maxpool3s2_fwd");
    return 0;
}
int maxpool3_bwd(const void* dy, const void* x, void* dx, int dt, int N, int H, int W, int C, int SY, int SX, cudaStream_t st) {
    const int Ho = (H - 1) / SY + 1, Wo = (W - 1) / SX + 1;
    const long long total = (long long)N * H * W * VECN(C);
#define CALL(T, V) maxpool3s2_bwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)dy, (const T*)x, (T*)dx, N, H, W, C, Ho, Wo, SY, SX)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("maxpool3s2_bwd");
    return 0;
}
int resize_bilinear_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, int Ho, int Wo, cudaStream_t st) {
    const long long total = (long long)N * Ho * Wo * VECN(C);
#define CALL(T, V) resize_bilinear_fwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, Ho, Wo)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("resize_bilinear_fwd");
    return 0;
}
int resize_bilinear_bwd(const void* dy, float* dx, int dt, int N, int H, int W, int C, int Ho, int Wo, cudaStream_t st) {
    const long long total = (long long)N * Ho * Wo * C;
    if (dt == AFFGW_F32)
        resize_bilinear_bwd_kernel<float><<<ew_blocks(total), 256, 0, st>>>((const float*)dy, dx, N, H, W, C, Ho, Wo);
    else
        resize_bilinear_bwd_kernel<bf16><<<ew_blocks(total), 256, 0, st>>>((const bf16*)dy, dx, N, H, W, C, Ho, Wo);
    AFFGW_LAUNCH_CHECK("resize_bilinear_bwd");
    return 0;
}
int add_act(const void* a, const void* b, void* out, int dt, long long n, int act, cudaStream_t st) {
    const long long nv = n % 8 == 0 ? n / 8 : n;
#define CALL(T, V) add_act_kernel<T, V><<<ew_blocks(nv), 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, nv, act)
    DISPATCH_T_VEC(dt, n, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("add_act");
    return 0;
}
int add2(const void* a, const void* b, void* out, int dt, long long n, cudaStream_t st) {
    const long long nv = n % 8 == 0 ? n / 8 : n;
#define CALL(T, V) add2_kernel<T, V><<<ew_blocks(nv), 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, nv)
    DISPATCH_T_VEC(dt, n, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("add2");
    return 0;
}
int act_bwd(const void* dy, const void* y, void* dz, int dt, long long n, int act, cudaStream_t st) {
    const long long nv = n % 8 == 0 ? n / 8 : n;
#define CALL(T, V) act_bwd_kernel<T, V><<<ew_blocks(nv), 256, 0, st>>>((const T*)dy, (const T*)y, (T*)dz, nv, act)
    DISPATCH_T_VEC(dt, n, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("act_bwd");
    return 0;
}
int resize_nearest_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, int Ho, int Wo, cudaStream_t st) {
    const long long total = (long long)N * Ho * Wo * VECN(C);
#define CALL(T, V) resize_nearest_fwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, Ho, Wo)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("resize_nearest_fwd");
    return 0;
}
int resize_nearest_bwd(const void* dy, void* dx, int dt, int N, int H, int W, int C, int Ho, int Wo, cudaStream_t st) {
    const long long total = (long long)N * H * W * VECN(C);
#define CALL(T, V) resize_nearest_bwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)dy, (T*)dx, N, H, W, C, Ho, Wo)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("resize_nearest_bwd");
    return 0;
}
int embedding_fwd(const long long* ids, const float* table, void* out, int dt, long long n_ids, int E, int V, int* err,
                  cudaStream_t st) {
    if (dt == AFFGW_F32) embedding_fwd_kernel<float><<<ew_blocks(n_ids * E), 256, 0, st>>>(ids, table, (float*)out, n_ids, E, V, err);
    else embedding_fwd_kernel<bf16><<<ew_blocks(n_ids * E), 256, 0, st>>>(ids, table, (bf16*)out, n_ids, E, V, err);
    AFFGW_LAUNCH_CHECK("embedding_fwd");
    return 0;
}
int embedding_bwd(const long long* ids, const void* dout, float* dtable, int dt, long long n_ids, int E, int V,
                  cudaStream_t st) {
    if (dt == AFFGW_F32) embedding_bwd_kernel<float><<<ew_blocks(n_ids * E), 256, 0, st>>>(ids, (const float*)dout, dtable, n_ids, E, V);
    else embedding_bwd_kernel<bf16><<<ew_blocks(n_ids * E), 256, 0, st>>>(ids, (const bf16*)dout, dtable, n_ids, E, V);
    AFFGW_LAUNCH_CHECK("embedding_bwd");
    return 0;
}
int text_tile_fwd(const void* chars, void* out, int dt, int B, int H, int W, int C, int ts, int reps, cudaStream_t st) {
    const long long total = (long long)B * H * W * VECN(C);
#define CALL(T, V) text_tile_fwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)chars, (T*)out, B, H, W, C, ts, reps)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("text_tile_fwd");
    return 0;
}
int text_tile_bwd(const void* dout, void* dchars, int dt, int B, int H, int W, int C, int ts, int reps, cudaStream_t st) {
    const long long total = (long long)B * (ts + 1) * VECN(C);
#define CALL(T, V) text_tile_bwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T*)dout, (T*)dchars, B, H, W, C, ts, reps)
    DISPATCH_T_VEC(dt, C, CALL);
#undef CALL
    AFFGW_LAUNCH_CHECK("text_tile_bwd");
    return 0;
}
int bce_logits_fwd(const void* x, int dt, float target, float* loss, long long n, cudaStream_t st) {
    cudaMemsetAsync(loss, 0, sizeof(float), st);
    const int blocks = (int)max(1LL, min(148LL, (n + 255) / 256));
    if (dt == AFFGW_F32) bce_logits_fwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, target, loss, n);
    else bce_logits_fwd_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)x, target, loss, n);
    AFFGW_LAUNCH_CHECK("bce_logits_fwd");
    return 0;
}
int bce_logits_bwd(const void* x, int dt, float target, const float* gout, void* dx, long long n, cudaStream_t st) {
    if (dt == AFFGW_F32) bce_logits_bwd_kernel<float><<<ew_blocks(n), 256, 0, st>>>((const float*)x, target, gout, (float*)dx, n);
    else bce_logits_bwd_kernel<bf16><<<ew_blocks(n), 256, 0, st>>>((const bf16*)x, target, gout, (bf16*)dx, n);
    AFFGW_LAUNCH_CHECK("bce_logits_bwd");
    return 0;
}
int softmax_ce_fwd(const void* x, int dt, const long long* y, float* loss, int B, int C, int* err, cudaStream_t st) {
    cudaMemsetAsync(loss, 0, sizeof(float), st);
    if (dt == AFFGW_F32) softmax_ce_fwd_kernel<float><<<cdiv(B, 8), 256, 0, st>>>((const float*)x, y, loss, B, C, err);
    else softmax_ce_fwd_kernel<bf16><<<cdiv(B, 8), 256, 0, st>>>((const bf16*)x, y, loss, B, C, err);
    AFFGW_LAUNCH_CHECK("softmax_ce_fwd");
    return 0;
}
int softmax_ce_bwd(const void* x, int dt, const long long* y, const float* gout, void* dx, int B, int C, cudaStream_t st) {
    if (dt == AFFGW_F32) softmax_ce_bwd_kernel<float><<<cdiv(B, 8), 256, 0, st>>>((const float*)x, y, gout, (float*)dx, B, C);
    else softmax_ce_bwd_kernel<bf16><<<cdiv(B, 8), 256, 0, st>>>((const bf16*)x, y, gout, (bf16*)dx, B, C);
    AFFGW_LAUNCH_CHECK("softmax_ce_bwd");
    return 0;
}
int u8_to_image(const unsigned char* src, float* dst, long long n, cudaStream_t st) {
    const long long vecs = (n + 15) / 16;
    u8_to_image_kernel<<<ew_blocks(vecs), 256, 0, st>>>(src, dst, n);
    AFFGW_LAUNCH_CHECK("u8_to_image");
    return 0;
}
int nchw_to_nhwc(const float* x, void* y, int dt, int N, int C, long long HW, int cpad, cudaStream_t st) {
    dim3 grid(cdiv(HW, 32), cdiv(cpad, 32), N);
    if (dt == AFFGW_F32) nchw_to_nhwc_kernel<float><<<grid, 256, 0, st>>>(x, (float*)y, C, HW, cpad);
    else nchw_to_nhwc_kernel<bf16><<<grid, 256, 0, st>>>(x, (bf16*)y, C, HW, cpad);
    AFFGW_LAUNCH_CHECK("nchw_to_nhwc");
    return 0;
}
int nhwc_to_nchw(const void* x, float* y, int dt, int N, int C, long long HW, int cpad, cudaStream_t st) {
    dim3 grid(cdiv(HW, 32), cdiv(C, 32), N);
    if (dt == AFFGW_F32) nhwc_to_nchw_kernel<float><<<grid, 256, 0, st>>>((const float*)x, y, C, HW, cpad);
    else nhwc_to_nchw_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, y, C, HW, cpad);
    AFFGW_LAUNCH_CHECK("nhwc_to_nchw");
    return 0;
}
// to_out = 1: out = cat(a, b) ; to_out = 0: a, b = split(out)
int concat2(void* a, void* b, void* out, int dt, long long rows, int ca, int cb, int to_out, cudaStream_t st) {
    const bool v8 = (ca % 8 == 0) && (cb % 8 == 0);
    const long long total = rows * (v8 ? (ca + cb) / 8 : (ca + cb));
    if (dt == AFFGW_F32) {
        if (v8) concat2_kernel<float, 8><<<ew_blocks(total), 256, 0, st>>>((const float*)a, (const float*)b, (float*)out, rows, ca, cb, to_out);
        else concat2_kernel<float, 1><<<ew_blocks(total), 256, 0, st>>>((const float*)a, (const float*)b, (float*)out, rows, ca, cb, to_out);
    } else {
        if (v8) concat2_kernel<bf16, 8><<<ew_blocks(total), 256, 0, st>>>((const bf16*)a, (const bf16*)b, (bf16*)out, rows, ca, cb, to_out);
        else concat2_kernel<bf16, 1><<<ew_blocks(total), 256, 0, st>>>((const bf16*)a, (const bf16*)b, (bf16*)out, rows, ca, cb, to_out);
    }
    AFFGW_LAUNCH_CHECK("concat2");
    return 0;
}
int cast_dtype(const void* x, int in_dt, void* y, int out_dt, long long n, cudaStream_t st) {
    if (in_dt == AFFGW_F32 && out_dt == AFFGW_BF16) cast_kernel<float, bf16><<<ew_blocks(n), 256, 0, st>>>((const float*)x, (bf16*)y, n);
    else if (in_dt == AFFGW_BF16 && out_dt == AFFGW_F32) cast_kernel<bf16, float><<<ew_blocks(n), 256, 0, st>>>((const bf16*)x, (float*)y, n);
    else if (in_dt == AFFGW_F32) cast_kernel<float, float><<<ew_blocks(n), 256, 0, st>>>((const float*)x, (float*)y, n);
    else cast_kernel<bf16, bf16><<<ew_blocks(n), 256, 0, st>>>((const bf16*)x, (bf16*)y, n);
    AFFGW_LAUNCH_CHECK("cast_dtype");
    return 0;
}
