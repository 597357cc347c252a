// Convolutions with ONE channel on one side, on the CUDA cores in fp32: the 7x7 stems of the discriminator / writer
// classifier (1 -> 16, reference modules_tro.py:125-128,175-178) and the 7x7 output convolution of the decoder
// (64 -> 1 + tanh, modules_tro.py:600-603), forward, input gradient and weight gradient.
//
// A tensor-core MMA is 128 rows x (>= 16) columns: with a single input or output channel 15/16 of it is padding and, worse,
// every 128-row MMA reads 4 KB of shared memory whatever N is, so these layers took ~17 ms of a 155 ms training step for
// 0.4 % of its FLOPs.  They are stencils, not GEMMs: each kernel below stages a halo tile in shared
// memory, keeps a sliding window of it in registers and does 7 - 28 FMAs per shared-memory load.  fp32 throughout, so
// no operand split is needed either.
//
//   conv_1toN   out[p][c]      = bias[c] + sum_tap in[p + tap] * w[tap][c]              (stem forward; output-conv dgrad)
//   conv_Nto1   out[p]         = act(bias + sum_{tap,c} in[p + tap][c] * w[c][tap])     (output-conv forward; stem dgrad)
//   corr_wgrad  dw[c][tap]    += sum_p one[..] * many[..][c]                            (both weight gradients)
// Input gradients are produced on the padded frame (zero-padded correlation with the flipped filter) and folded back by
// conv_fold (reflect padding), like the tensor-core path does.
#include "common.cuh"

namespace {

constexpr int KMAX = 7;

// ------------------------------------------------------------------------------------------------ 1 -> NC channels
// tile = 8 rows x 32 columns of output pixels, one pixel per thread, NC accumulators in registers
template <int NC>
__global__ void __launch_bounds__(256)
conv_1toN_kernel(const float* __restrict__ in, const float* __restrict__ wt /* [K*K][NC] */, const float* __restrict__ bias,
                 float* __restrict__ out, int H, int W, int Ho, int Wo, int K, int pad, int pad_mode, int post_act) {
    __shared__ float tile[8 + KMAX - 1][32 + KMAX - 1 + 2];
    __shared__ __align__(16) float ws[KMAX * KMAX * NC];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8, n = blockIdx.z;
    const int th = 8 + K - 1, tw = 32 + K - 1;
    for (int i = tid; i < th * tw; i += 256) {
        const int yy = i / tw, xx = i - yy * tw;
        const int sy = map_coord(y0 + yy - pad, H, pad_mode, 1, 1), sx = map_coord(x0 + xx - pad, W, pad_mode, 1, 1);
        tile[yy][xx] = (sy >= 0 && sx >= 0) ? __ldg(in + ((size_t)n * H + sy) * W + sx) : 0.f;
    }
    for (int i = tid; i < K * K * NC; i += 256) ws[i] = __ldg(wt + i);
    __syncthreads();
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = bias ? __ldg(bias + c) : 0.f;
    for (int ky = 0; ky < K; ++ky)
        for (int kx = 0; kx < K; ++kx) {
            const float xv = tile[ty + ky][tx + kx];
            const float4* w4 = reinterpret_cast<const float4*>(ws + (ky * K + kx) * NC);
#pragma unroll
            for (int c = 0; c < NC / 4; ++c) {
                const float4 w = w4[c];
                acc[4 * c] = fmaf(xv, w.x, acc[4 * c]);
                acc[4 * c + 1] = fmaf(xv, w.y, acc[4 * c + 1]);
                acc[4 * c + 2] = fmaf(xv, w.z, acc[4 * c + 2]);
                acc[4 * c + 3] = fmaf(xv, w.w, acc[4 * c + 3]);
            }
        }
    const int y = y0 + ty, x = x0 + tx;
    if (y < Ho && x < Wo) {
        float4* o = reinterpret_cast<float4*>(out + (((size_t)n * Ho + y) * Wo + x) * NC);
#pragma unroll
        for (int c = 0; c < NC / 4; ++c)
            o[c] = make_float4(act_apply(acc[4 * c], post_act), act_apply(acc[4 * c + 1], post_act),
                               act_apply(acc[4 * c + 2], post_act), act_apply(acc[4 * c + 3], post_act));
    }
}

// ------------------------------------------------------------------------------------------------ C -> 1 channel
// tile = 16 rows x 128 columns; a thread owns 2 rows x 4 consecutive pixels, so every input row it reads from shared memory
// feeds both output rows, and the K x K filter of the current channel sits in registers (the kernel is bound by shared-memory
// reads, not FMAs: 3 128-bit loads per 28 FMAs with one row per thread).  8 channels of the halo tile in shared memory at a time.
constexpr int N1_CC = 8, N1_TW = 128, N1_ROWS = 16, N1_PITCH = N1_TW + 8;     // pitch: >= 128 + K - 1, multiple of 4 floats
template <int K>
__global__ void __launch_bounds__(256)
conv_Nto1_kernel(const float* __restrict__ in, const float* __restrict__ w /* [C][K][K] */, const float* __restrict__ bias,
                 float* __restrict__ out, int H, int W, int C, int Ho, int Wo, int pad, int pad_mode, int post_act) {
    extern __shared__ __align__(16) float smem[];
    constexpr int TH = N1_ROWS + K - 1;
    float* tile = smem;                                    // [N1_CC][TH][N1_PITCH]
    float* ws = smem + N1_CC * TH * N1_PITCH;              // [N1_CC][K][K]
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int x0 = blockIdx.x * N1_TW, y0 = blockIdx.y * N1_ROWS, n = blockIdx.z;
    constexpr int TWL = N1_TW + K - 1;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int c0 = 0; c0 < C; c0 += N1_CC) {
        __syncthreads();
        for (int p = tid; p < TH * TWL; p += 256) {            // one pixel (8 contiguous channels = 2 x 16 bytes) per thread
            const int yy = p / TWL, xx = p - yy * TWL;
            const int sy = map_coord(y0 + yy - pad, H, pad_mode, 1, 1), sx = map_coord(x0 + xx - pad, W, pad_mode, 1, 1);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (sy >= 0 && sx >= 0) {
                const float4* src = reinterpret_cast<const float4*>(in + (((size_t)n * H + sy) * W + sx) * C + c0);
                a = __ldg(src);
                b = __ldg(src + 1);
            }
            float* d = tile + yy * N1_PITCH + xx;
            d[0 * TH * N1_PITCH] = a.x; d[1 * TH * N1_PITCH] = a.y; d[2 * TH * N1_PITCH] = a.z; d[3 * TH * N1_PITCH] = a.w;
            d[4 * TH * N1_PITCH] = b.x; d[5 * TH * N1_PITCH] = b.y; d[6 * TH * N1_PITCH] = b.z; d[7 * TH * N1_PITCH] = b.w;
        }
        for (int i = tid; i < N1_CC * K * K; i += 256) ws[i] = __ldg(w + (size_t)c0 * K * K + i);
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < N1_CC; ++c) {
            float wr[K][K];
#pragma unroll
            for (int ky = 0; ky < K; ++ky)
#pragma unroll
                for (int kx = 0; kx < K; ++kx) wr[ky][kx] = ws[(c * K + ky) * K + kx];
#pragma unroll
            for (int r = 0; r < K + 1; ++r) {                  // input row 2 ty + r feeds output row j with filter row r - j
                const float* row = tile + (c * TH + 2 * ty + r) * N1_PITCH + 4 * tx;
                float v[K + 3];
#pragma unroll
                for (int j = 0; j < (K + 3 + 3) / 4; ++j) {
                    const float4 f = reinterpret_cast<const float4*>(row)[j];
                    if (4 * j < K + 3) v[4 * j] = f.x;
                    if (4 * j + 1 < K + 3) v[4 * j + 1] = f.y;
                    if (4 * j + 2 < K + 3) v[4 * j + 2] = f.z;
                    if (4 * j + 3 < K + 3) v[4 * j + 3] = f.w;
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int ky = r - j;
                    if (ky < 0 || ky >= K) continue;           // resolved at compile time
#pragma unroll
                    for (int kx = 0; kx < K; ++kx)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(v[i + kx], wr[ky][kx], acc[j][i]);
                }
            }
        }
    }
    const float b = bias ? __ldg(bias) : 0.f;
    const int x = x0 + 4 * tx;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int y = y0 + 2 * ty + j;
        if (y < Ho) {
            float* o = out + ((size_t)n * Ho + y) * Wo + x;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (x + i < Wo) o[i] = act_apply(acc[j][i] + b, post_act);
        }
    }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dw[c][ky][kx] += sum over output pixels (n, y, x) of
//   MANY_SHIFTED:  one[n][y][x]        * many[n][y + ky][x + kx][c]      (one = dY, many = padded input; C -> 1 layer)
//   else        :  one[n][y+ky][x+kx]  * many[n][y][x][c]               (one = padded input, many = dY; 1 -> C layer)
// thread = (channel of an 8-channel chunk, ky, one of 4 column segments); K accumulators and a K-wide sliding window of the
// shifted operand in registers; a block walks units of 8 output rows and flushes its accumulators once with atomics.
constexpr int WG_CC = 8, WG_R = 8, WG_SEG = 4;
template <int K, bool MANY_SHIFTED>
__global__ void __launch_bounds__(256)
corr_wgrad_kernel(const float* __restrict__ many, const float* __restrict__ one, float* __restrict__ dw, int N, int H, int W,
                  int C, int Ho, int Wo, int pad, int pad_mode, int units) {
    extern __shared__ __align__(16) float smem[];
    // shifted operand: (WG_R + K - 1) x (Wo + K - 1) pixels; the other: WG_R x Wo
    const int sw = Wo + K - 1, sh = WG_R + K - 1;
    const int many_w = MANY_SHIFTED ? sw : Wo, many_h = MANY_SHIFTED ? sh : WG_R;
    const int one_w = MANY_SHIFTED ? Wo : sw, one_h = MANY_SHIFTED ? WG_R : sh;
    const int mpitch = many_w * WG_CC + 8;                 // row pitch = 8 mod 32 words: the 4 ky rows of a warp hit 4 bank groups
    float* mt = smem;                                      // [many_h][many_w][WG_CC] (+8 pad per row)
    float* ot = smem + many_h * mpitch;                    // [one_h][one_w]
    const int tid = threadIdx.x;
    const int c = tid & 7, ky = (tid >> 3) & 7, seg = tid >> 6;
    const int c0 = blockIdx.y * WG_CC;
    const int segw = (Wo + WG_SEG - 1) / WG_SEG;
    const int xa = seg * segw, xb = min(Wo, xa + segw);
    const int rows_per_img = (Ho + WG_R - 1) / WG_R;
    float acc[K];
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i] = 0.f;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int n = u / rows_per_img, y0 = (u - n * rows_per_img) * WG_R;
        __syncthreads();
        // many tile (C-channel operand)
        for (int p = tid; p < many_h * many_w; p += 256) {     // one pixel (8 contiguous channels = 2 x 16 bytes) per thread
            const int yy = p / many_w, xx = p - yy * many_w;
            const float4* src = nullptr;
            if (MANY_SHIFTED) {
                const int sy = map_coord(y0 + yy - pad, H, pad_mode, 1, 1), sx = map_coord(xx - pad, W, pad_mode, 1, 1);
                if (sy >= 0 && sx >= 0) src = reinterpret_cast<const float4*>(many + (((size_t)n * H + sy) * W + sx) * C + c0);
            } else if (y0 + yy < Ho) {
                src = reinterpret_cast<const float4*>(many + (((size_t)n * Ho + y0 + yy) * Wo + xx) * C + c0);
            }
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (src) {
                a = __ldg(src);
                b = __ldg(src + 1);
            }
            float4* d = reinterpret_cast<float4*>(mt + yy * mpitch + xx * WG_CC);
            d[0] = a;
            d[1] = b;
        }
        // one tile (single-channel operand)
        for (int i = tid; i < one_h * one_w; i += 256) {
            const int yy = i / one_w, xx = i - yy * one_w;
            float v = 0.f;
            if (MANY_SHIFTED) {
                if (y0 + yy < Ho) v = __ldg(one + ((size_t)n * Ho + y0 + yy) * Wo + xx);
            } else {
                const int sy = map_coord(y0 + yy - pad, H, pad_mode, 1, 1), sx = map_coord(xx - pad, W, pad_mode, 1, 1);
                if (sy >= 0 && sx >= 0) v = __ldg(one + ((size_t)n * H + sy) * W + sx);
            }
            ot[yy * one_w + xx] = v;
        }
        __syncthreads();
        if (ky < K && xa < xb) {
            for (int r = 0; r < WG_R; ++r) {
                float win[K];                              // sliding window of the shifted operand along x
                if (MANY_SHIFTED) {
                    const float* mrow = mt + (r + ky) * mpitch + c;
                    const float* orow = ot + r * one_w;
#pragma unroll
                    for (int i = 0; i < K - 1; ++i) win[i + 1] = mrow[(xa + i) * WG_CC];
                    for (int x = xa; x < xb; ++x) {
#pragma unroll
                        for (int i = 0; i < K - 1; ++i) win[i] = win[i + 1];
                        win[K - 1] = mrow[(x + K - 1) * WG_CC];
                        const float o = orow[x];
#pragma unroll
                        for (int i = 0; i < K; ++i) acc[i] = fmaf(o, win[i], acc[i]);
                    }
                } else {
                    const float* mrow = mt + r * mpitch + c;
                    const float* orow = ot + (r + ky) * one_w;
#pragma unroll
                    for (int i = 0; i < K - 1; ++i) win[i + 1] = orow[xa + i];
                    for (int x = xa; x < xb; ++x) {
#pragma unroll
                        for (int i = 0; i < K - 1; ++i) win[i] = win[i + 1];
                        win[K - 1] = orow[x + K - 1];
                        const float m = mrow[x * WG_CC];
#pragma unroll
                        for (int i = 0; i < K; ++i) acc[i] = fmaf(m, win[i], acc[i]);
                    }
                }
            }
        }
    }
    if (ky < K && c0 + c < C) {
#pragma unroll
        for (int i = 0; i < K; ++i) atomicAdd(dw + ((size_t)(c0 + c) * K + ky) * K + i, acc[i]);
    }
}

// w OIHW with one of O / I equal to 1 -> [K*K][NC] (optionally the spatially flipped filter for the input gradient)
__global__ void thin_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int NC, int K, int flip, int tap_major) {
    const int total = NC * K * K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i / (K * K), t = i - c * K * K;
        const int ky = t / K, kx = t - ky * K;
        const float v = w[c * K * K + (flip ? (K - 1 - ky) * K + (K - 1 - kx) : t)];
        out[tap_major ? t * NC + c : i] = v;
    }
}

}  // namespace

// Is (Cin, Cout, K) one of the single-channel-sided stencils?
int conv_thin_ok(int Cin, int Cout, int K, int stride, int up) {
    if (stride != 1 || up != 1 || (K != 3 && K != 5 && K != 7)) return 0;
    const int many = Cin == 1 ? Cout : Cin;
    if (many != 8 && many != 16 && many != 32 && many != 64) return 0;      // instantiated widths of the 1 -> N kernel
    if (Cin == 1) return 1;
    if (Cout == 1) return 2;
    return 0;
}

static int launch_1toN(const float* in, const float* wt, const float* bias, float* out, int N, int H, int W, int NC, int Ho, int Wo,
                       int K, int pad, int pad_mode, int post_act, cudaStream_t st) {
    dim3 grid((Wo + 31) / 32, (Ho + 7) / 8, N);
    switch (NC) {
        case 8: conv_1toN_kernel<8><<<grid, 256, 0, st>>>(in, wt, bias, out, H, W, Ho, Wo, K, pad, pad_mode, post_act); break;
        case 16: conv_1toN_kernel<16><<<grid, 256, 0, st>>>(in, wt, bias, out, H, W, Ho, Wo, K, pad, pad_mode, post_act); break;
        case 32: conv_1toN_kernel<32><<<grid, 256, 0, st>>>(in, wt, bias, out, H, W, Ho, Wo, K, pad, pad_mode, post_act); break;
        default: conv_1toN_kernel<64><<<grid, 256, 0, st>>>(in, wt, bias, out, H, W, Ho, Wo, K, pad, pad_mode, post_act); break;
    }
    AFFGW_LAUNCH_CHECK("conv_1toN");
    return 0;
}

template <int K>
static int launch_Nto1_k(const float* in, const float* w, const float* bias, float* out, int N, int H, int W, int C, int Ho, int Wo,
                         int pad, int pad_mode, int post_act, cudaStream_t st) {
    const int smem = (N1_CC * (N1_ROWS + K - 1) * N1_PITCH + N1_CC * K * K) * 4;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_Nto1_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    dim3 grid((Wo + N1_TW - 1) / N1_TW, (Ho + N1_ROWS - 1) / N1_ROWS, N);
    conv_Nto1_kernel<K><<<grid, 256, smem, st>>>(in, w, bias, out, H, W, C, Ho, Wo, pad, pad_mode, post_act);
    AFFGW_LAUNCH_CHECK("conv_Nto1");
    return 0;
}
static int launch_Nto1(const float* in, const float* w, const float* bias, float* out, int N, int H, int W, int C, int Ho, int Wo,
                       int K, int pad, int pad_mode, int post_act, cudaStream_t st) {
    if (K == 3) return launch_Nto1_k<3>(in, w, bias, out, N, H, W, C, Ho, Wo, pad, pad_mode, post_act, st);
    if (K == 5) return launch_Nto1_k<5>(in, w, bias, out, N, H, W, C, Ho, Wo, pad, pad_mode, post_act, st);
    return launch_Nto1_k<7>(in, w, bias, out, N, H, W, C, Ho, Wo, pad, pad_mode, post_act, st);
}

template <int K, bool MS>
static int launch_corr_k(const float* many, const float* one, float* dw, int N, int H, int W, int C, int Ho, int Wo, int pad,
                         int pad_mode, cudaStream_t st) {
    const int sw = Wo + K - 1, sh = WG_R + K - 1;
    const int many_w = MS ? sw : Wo, many_h = MS ? sh : WG_R, one_w = MS ? Wo : sw, one_h = MS ? WG_R : sh;
    const int smem = (many_h * (many_w * WG_CC + 8) + one_h * one_w) * 4;
    if (smem > 200 * 1024) {
        affgw_set_error("conv_thin wgrad: a %d-pixel wide map does not fit the shared-memory tile", Wo);
        return -1;
    }
    static int configured = 0;
    if (configured < smem) {
        cudaFuncSetAttribute(corr_wgrad_kernel<K, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured = 200 * 1024;
    }
    const int units = N * ((Ho + WG_R - 1) / WG_R);
    const int chunks = (C + WG_CC - 1) / WG_CC;
    int bx = (148 * 2 + chunks - 1) / chunks;
    if (bx > units) bx = units;
    dim3 grid(bx, chunks);
    corr_wgrad_kernel<K, MS><<<grid, 256, smem, st>>>(many, one, dw, N, H, W, C, Ho, Wo, pad, pad_mode, units);
    AFFGW_LAUNCH_CHECK("conv_thin_wgrad");
    return 0;
}
template <bool MS>
static int launch_corr(const float* many, const float* one, float* dw, int N, int H, int W, int C, int Ho, int Wo, int K, int pad,
                       int pad_mode, cudaStream_t st) {
    if (K == 3) return launch_corr_k<3, MS>(many, one, dw, N, H, W, C, Ho, Wo, pad, pad_mode, st);
    if (K == 5) return launch_corr_k<5, MS>(many, one, dw, N, H, W, C, Ho, Wo, pad, pad_mode, st);
    return launch_corr_k<7, MS>(many, one, dw, N, H, W, C, Ho, Wo, pad, pad_mode, st);
}

// y = act(conv(pad(x)) + bias).  scratch: K*K*max(Cin,Cout) floats (re-laid filter for the 1 -> N kernel)
int conv_thin_fwd(const float* x, const float* w, const float* bias, float* y, float* scratch, const ConvGeom& g, cudaStream_t st) {
    const int kind = conv_thin_ok(g.Cin, g.Cout, g.KH, max(g.stride, g.stride_w), g.up);
    if (!kind || g.KH != g.KW || g.pre_act != ACT_NONE) {
        affgw_set_error("conv_thin: not a single-channel-sided stride-1 convolution");
        return -1;
    }
    if (kind == 1) {
        thin_weight_kernel<<<4, 256, 0, st>>>(w, scratch, g.Cout, g.KH, 0, 1);
        AFFGW_LAUNCH_CHECK("thin_weight");
        return launch_1toN(x, scratch, bias, y, g.N, g.H, g.W, g.Cout, g.Ho, g.Wo, g.KH, g.pad, g.pad_mode, g.post_act, st);
    }
    return launch_Nto1(x, w, bias, y, g.N, g.H, g.W, g.Cin, g.Ho, g.Wo, g.KH, g.pad, g.pad_mode, g.post_act, st);
}

// gradient w.r.t. the PADDED input frame [N][H + 2 pad][W + 2 pad][Cin] (zero-padded correlation of dY with the flipped filter)
int conv_thin_dgrad_frame(const float* dy, const float* w, float* dframe, float* scratch, const ConvGeom& g, cudaStream_t st) {
    const int kind = conv_thin_ok(g.Cin, g.Cout, g.KH, max(g.stride, g.stride_w), g.up);
    const int K = g.KH, Hp = g.H + 2 * g.pad, Wp = g.W + 2 * g.pad;
    if (!kind) {
        affgw_set_error("conv_thin: not a single-channel-sided stride-1 convolution");
        return -1;
    }
    if (kind == 1) {        // forward 1 -> N  =>  input gradient N -> 1; w [N][1][K][K] is already [c][tap]
        thin_weight_kernel<<<4, 256, 0, st>>>(w, scratch, g.Cout, K, 1, 0);
        AFFGW_LAUNCH_CHECK("thin_weight");
        return launch_Nto1(dy, scratch, nullptr, dframe, g.N, g.Ho, g.Wo, g.Cout, Hp, Wp, K, K - 1, PAD_ZERO, ACT_NONE, st);
    }
    // forward N -> 1  =>  input gradient 1 -> N; w [1][N][K][K] -> [tap][c], flipped
    thin_weight_kernel<<<4, 256, 0, st>>>(w, scratch, g.Cin, K, 1, 1);
    AFFGW_LAUNCH_CHECK("thin_weight");
    return launch_1toN(dy, scratch, nullptr, dframe, g.N, g.Ho, g.Wo, g.Cin, Hp, Wp, K, K - 1, PAD_ZERO, ACT_NONE, st);
}

// dw (OIHW, one of O / I is 1) += correlation; the caller zeroes dw
int conv_thin_wgrad(const float* x, const float* dy, float* dw, const ConvGeom& g, cudaStream_t st) {
    const int kind = conv_thin_ok(g.Cin, g.Cout, g.KH, max(g.stride, g.stride_w), g.up);
    if (!kind) {
        affgw_set_error("conv_thin: not a single-channel-sided stride-1 convolution");
        return -1;
    }
    if (kind == 1)          // many = dY (Cout channels), one = padded x
        return launch_corr<false>(dy, x, dw, g.N, g.H, g.W, g.Cout, g.Ho, g.Wo, g.KH, g.pad, g.pad_mode, st);
    return launch_corr<true>(x, dy, dw, g.N, g.H, g.W, g.Cin, g.Ho, g.Wo, g.KH, g.pad, g.pad_mode, st);
}
