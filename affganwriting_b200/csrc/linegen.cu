// Line-level generator pieces (reference line_generation/model/pure_gen.py) that are not convolutions /
// GEMMs / instance norms: the depthwise 3x3 binomial blur behind every up-sampling convolution (pure_gen.py:123-136) and
// PixelNorm on the style vector (:306-311).  fp32, NHWC, HBM-bound (channel counts 256 -> 16 at up to 64 x 1024 pixels).
#include "common.cuh"

namespace {

// y[n][h][w][c] = sum_{dy,dx} k[dy] k[dx] x[n][h+dy-1][w+dx-1][c] / 16, k = (1, 2, 1), zero padding.
// The kernel is symmetric, so the same call is its own backward (pure_gen.py:82-118 keeps a flipped copy for that).
// One thread per (pixel, 4-channel vector): 9 coalesced 16-byte loads, the neighbouring threads' loads hit L1/L2.
template <int VEC>
__global__ void __launch_bounds__(256) blur3_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C) {
    const int cv = C / VEC;
    const long long total = (long long)N * H * W * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * VEC;
        long long p = i / cv;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int n = (int)(p / H);
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int hh = h + dy;
            if (hh < 0 || hh >= H) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int ww = w + dx;
                if (ww < 0 || ww >= W) continue;
                const float wgt = (dy == 0 ? 2.f : 1.f) * (dx == 0 ? 2.f : 1.f) * (1.f / 16.f);
                const float* src = x + (((long long)n * H + hh) * W + ww) * C + c;
                if constexpr (VEC == 4) {
                    const float4 v = *reinterpret_cast<const float4*>(src);
                    acc[0] = fmaf(wgt, v.x, acc[0]); acc[1] = fmaf(wgt, v.y, acc[1]);
                    acc[2] = fmaf(wgt, v.z, acc[2]); acc[3] = fmaf(wgt, v.w, acc[3]);
                } else {
                    acc[0] = fmaf(wgt, src[0], acc[0]);
                }
            }
        }
        float* dst = y + (((long long)n * H + h) * W + w) * C + c;
        if constexpr (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else dst[0] = acc[0];
    }
}

// y[r][:] = x[r][:] / sqrt(mean_c x[r][c]^2 + eps)      one warp per row
__global__ void pixelnorm_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int C, float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long long)row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(xr[c], xr[c], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float r = 1.f / sqrtf(s / (float)C + eps);
    for (int c = lane; c < C; c += 32) y[(long long)row * C + c] = xr[c] * r;
}

}  // namespace

int blur3(const float* x, float* y, int N, int H, int W, int C, cudaStream_t st) {
    const bool v4 = (C % 4 == 0) && (((uintptr_t)x | (uintptr_t)y) % 16 == 0);
    const long long total = (long long)N * H * W * (v4 ? C / 4 : C);
    const int blocks = (int)max(1LL, min((long long)148 * 16, (total + 255) / 256));
    if (v4) blur3_kernel<4><<<blocks, 256, 0, st>>>(x, y, N, H, W, C);
    else blur3_kernel<1><<<blocks, 256, 0, st>>>(x, y, N, H, W, C);
    AFFGW_LAUNCH_CHECK("blur3");
    return 0;
}
int pixelnorm(const float* x, float* y, int rows, int C, float eps, cudaStream_t st) {
    pixelnorm_kernel<<<cdiv(rows, 8), 256, 0, st>>>(x, y, rows, C, eps);
    AFFGW_LAUNCH_CHECK("pixelnorm");
    return 0;
}
