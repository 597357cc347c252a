// ViT pieces of the DINOv2 style encoder (BASELINE.json configs[3]; reference wrapper GAN_word/dinomodel.py:127-163 around a
// DINOv2 ViT backbone): LayerNorm, exact GELU, LayerScale + residual, and the fused multi-head self-attention of short token
// sequences (81 tokens for a 64 x 216 image: 5 x 16 patches + cls).  The projections (qkv, proj, fc1, fc2, patch embedding,
// 1x1 reducers) run on the tcgen05 GEMM kernels.  fp32, forward (generation) only.
#include "common.cuh"

namespace {

// y[r][:] = (x[r][:] - mean) * rstd * w + b          one warp per row, two passes over the row held in L1
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                 float* __restrict__ y, long long rows, int D, float eps) {
    const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + row * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += xr[c];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)D;
    float v = 0.f;
    for (int c = lane; c < D; c += 32) { const float d = xr[c] - mean; v = fmaf(d, d, v); }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / (float)D + eps);
    for (int c = lane; c < D; c += 32) y[row * D + c] = (xr[c] - mean) * rstd * w[c] + b[c];
}

__global__ void gelu_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        y[i] = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));        // nn.GELU() (erf form)
    }
}

// y = x + gamma[c] * t      (LayerScale followed by the residual add)
__global__ void scale_residual_kernel(const float* __restrict__ x, const float* __restrict__ t, const float* __restrict__ gamma,
                                      float* __restrict__ y, long long n, int D) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = fmaf(gamma ? gamma[i % D] : 1.f, t[i], x[i]);
}

// qkv [B][N][3][H][hd] (the layout of Linear(D, 3D) followed by reshape(B, N, 3, H, hd)) -> out [B][N][H*hd]
// One block per (batch, head): K and V of the head in shared memory, one warp per query row:
// scores over the N keys (lanes stride the keys), soft-max, then the hd output columns (lanes stride hd).
constexpr int ATT_MAX_N = 128;
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int N, int H, int hd,
                                                        float scale) {
    extern __shared__ float sm[];
    float* ks = sm;                         // [N][hd + 1]
    float* vs = ks + (size_t)N * (hd + 1);  // [N][hd + 1]
    float* ps = vs + (size_t)N * (hd + 1);  // [warps][ATT_MAX_N]
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int D3 = 3 * H * hd;
    const float* base = qkv + (size_t)b * N * D3;
    for (int i = threadIdx.x; i < N * hd; i += blockDim.x) {
        const int n = i / hd, d = i % hd;
        ks[n * (hd + 1) + d] = base[(size_t)n * D3 + (1 * H + h) * hd + d];
        vs[n * (hd + 1) + d] = base[(size_t)n * D3 + (2 * H + h) * hd + d];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float* p = ps + warp * ATT_MAX_N;
    for (int qi = warp; qi < N; qi += nw) {
        const float* q = base + (size_t)qi * D3 + (0 * H + h) * hd;
        float mx = -INFINITY;
        for (int j = lane; j < N; j += 32) {
            float s = 0.f;
            for (int d = 0; d < hd; ++d) s = fmaf(q[d], ks[j * (hd + 1) + d], s);
            s *= scale;
            p[j] = s;
            mx = fmaxf(mx, s);
        }
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < N; j += 32) { const float e = expf(p[j] - mx); p[j] = e; sum += e; }
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __syncwarp();
        const float inv = 1.f / sum;
        for (int d = lane; d < hd; d += 32) {
            float acc = 0.f;
            for (int j = 0; j < N; ++j) acc = fmaf(p[j], vs[j * (hd + 1) + d], acc);
            out[((size_t)b * N + qi) * (H * hd) + h * hd + d] = acc * inv;
        }
        __syncwarp();
    }
}

}  // namespace

static inline int vit_blocks(long long n) { long long b = (n + 255) / 256; return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b)); }

int layernorm_fwd(const float* x, const float* w, const float* b, float* y, long long rows, int D, float eps, cudaStream_t st) {
    layernorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, w, b, y, rows, D, eps);
    AFFGW_LAUNCH_CHECK("layernorm_fwd");
    return 0;
}
int gelu_fwd(const float* x, float* y, long long n, cudaStream_t st) {
    gelu_kernel<<<vit_blocks(n), 256, 0, st>>>(x, y, n);
    AFFGW_LAUNCH_CHECK("gelu_fwd");
    return 0;
}
int scale_residual(const float* x, const float* t, const float* gamma, float* y, long long n, int D, cudaStream_t st) {
    scale_residual_kernel<<<vit_blocks(n), 256, 0, st>>>(x, t, gamma, y, n, D);
    AFFGW_LAUNCH_CHECK("scale_residual");
    return 0;
}
int attention_fwd(const float* qkv, float* out, int B, int N, int H, int hd, float scale, cudaStream_t st) {
    if (N > ATT_MAX_N) { affgw_set_error("attention_fwd: at most %d tokens (got %d)", ATT_MAX_N, N); return -1; }
    const size_t smem = ((size_t)2 * N * (hd + 1) + 8 * ATT_MAX_N) * sizeof(float);
    if (smem > 200 * 1024) { affgw_set_error("attention_fwd: head dimension %d too large", hd); return -1; }
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
            affgw_set_error("attention_fwd: cannot reserve shared memory");
            return -2;
        }
        configured = true;
    }
    attention_kernel<<<B * H, 256, smem, st>>>(qkv, out, N, H, hd, scale);
    AFFGW_LAUNCH_CHECK("attention_fwd");
    return 0;
}
