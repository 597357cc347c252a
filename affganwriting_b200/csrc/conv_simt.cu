// General implicit-GEMM convolution family on CUDA cores (fp32 FFMA, fp32 accumulate).
//
// This is the "accuracy mode" of the library (fp32 parity at 1e-4 against the CPU oracle is not reachable
// through TF32/bf16 tensor-core MMAs, SURVEY.md §7 hard part 1) and the fully general fallback for shapes the
// tcgen05 kernel does not take (C_in = 1 stem, C_out = 1 output conv, strided head, odd channel counts).
// It replaces the cuDNN/oneDNN calls behind nn.Conv2d at reference blocks.py:148-158 and
// vgg_tro_channel3_modi.py:47.  Layout: NHWC activations, [Cout][KH][KW][Cin] weights (K-major rows).
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

// ------------------------------------------------------------------------------------------------
// forward:  y[m][n] = post_act( sum_k gather(x)[m][k] * w[n][k] + bias[n] + addend[m][n] )
// The same kernel computes dgrad (x := dY, w := flipped/transposed weights, zero pad K-1[-p], zi = stride).
// ------------------------------------------------------------------------------------------------
template <typename TI, typename TW, typename TO>
__global__ void __launch_bounds__(256)
conv_fwd_simt_kernel(const TI* __restrict__ x, const TW* __restrict__ w, const float* __restrict__ bias,
                     const TO* __restrict__ addend, TO* __restrict__ y, const ConvGeom g) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // loader role: one tile row (pixel for A, output channel for B) and 4 consecutive k per thread
    const int lr = tid >> 2, kq = (tid & 3) * 4;
    const long long m = m0 + lr;
    const bool mvalid = m < g.M;
    int n_img = 0, oy = 0, ox = 0;
    if (mvalid) {
        ox = (int)(m % g.Wo);
        long long t = m / g.Wo;
        oy = (int)(t % g.Ho);
        n_img = (int)(t / g.Ho);
    }
    const int vy0 = oy * g.stride - g.pad, vx0 = ox * g.stride_w - g.pad;
    const TI* xn = x + (long long)n_img * g.H * g.W * g.in_pitch;
    const int wn = n0 + lr;
    const bool nvalid = wn < g.Cout;
    const TW* wrow = w + (long long)wn * g.Ktot;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int ty = tid >> 4, tx = tid & 15;

    for (int k0 = 0; k0 < g.Ktot; k0 += BK) {
        int kg = k0 + kq;
        int tap = kg / g.Cin, ci = kg - tap * g.Cin;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float av = 0.f, bv = 0.f;
            if (kg < g.Ktot) {
                if (mvalid) {
                    const int ky = tap / g.KW, kx = tap - ky * g.KW;
                    const int sy = map_coord(vy0 + ky, g.Hv, g.pad_mode, g.up, g.zi);
                    const int sx = map_coord(vx0 + kx, g.Wv, g.pad_mode, g.up, g.zi_w);
                    if (sy >= 0 && sx >= 0)
                        av = act_apply(to_f(xn[((long long)sy * g.W + sx) * g.in_pitch + ci]), g.pre_act);
                }
                if (nvalid) bv = to_f(wrow[kg]);
            }
            As[kq + j][lr] = av;
            Bs[kq + j][lr] = bv;
            ++kg;
            if (++ci == g.Cin) { ci = 0; ++tap; }
        }
        __syncthreads();
        // two-level summation: a 16-term partial per K block, then one add into the running sum, keeps the fp32
        // rounding error growth ~sqrt(K/16) instead of ~sqrt(K) (the 1e-4 image tolerance is tight, K is up to 12800)
        float part[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long mm = m0 + ty * 4 + i;
        if (mm >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nn = n0 + tx * 4 + j;
            if (nn >= g.Cout) continue;
            float v = acc[i][j];
            if (bias) v += bias[nn];
            if (addend) v += to_f(addend[mm * g.out_pitch + nn]);
            y[mm * g.out_pitch + nn] = from_f<TO>(act_apply(v, g.post_act));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// wgrad:  dW[co][ci][ky][kx] (OIHW, fp32, atomically accumulated) += sum_m dY[m][co] * gather(x)[m][k]
// grid = (Cout tiles, K tiles, pixel splits)
// ------------------------------------------------------------------------------------------------
template <typename TI, typename TG>
__global__ void __launch_bounds__(256)
conv_wgrad_simt_kernel(const TI* __restrict__ x, const TG* __restrict__ dy, float* __restrict__ dw, const ConvGeom g,
                       const long long m_per_split) {
    constexpr int BP = 16;
    __shared__ __align__(16) float Ds[BP][64 + 4];
    __shared__ __align__(16) float Xs[BP][64 + 4];
    const int tid = threadIdx.x;
    const int co0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const long long mbeg = (long long)blockIdx.z * m_per_split;
    const long long mend = min(g.M, mbeg + m_per_split);
    const int p = tid >> 4, q = (tid & 15) * 4;

    int ky[4], kx[4], ci[4];
    bool kval[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int kg = k0 + q + j;
        kval[j] = kg < g.Ktot;
        const int tap = kval[j] ? kg / g.Cin : 0;
        ci[j] = kval[j] ? kg - tap * g.Cin : 0;
        ky[j] = tap / g.KW;
        kx[j] = tap - ky[j] * g.KW;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int ty = tid >> 4, tx = tid & 15;

    for (long long mb = mbeg; mb < mend; mb += BP) {
        const long long m = mb + p;
        const bool mvalid = m < mend;
        int n_img = 0, oy = 0, ox = 0;
        if (mvalid) {
            ox = (int)(m % g.Wo);
            long long t = m / g.Wo;
            oy = (int)(t % g.Ho);
            n_img = (int)(t / g.Ho);
        }
        const TI* xn = x + (long long)n_img * g.H * g.W * g.in_pitch;
        const int vy0 = oy * g.stride - g.pad, vx0 = ox * g.stride_w - g.pad;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float dv = 0.f, xv = 0.f;
            if (mvalid) {
                const int co = co0 + q + j;
                if (co < g.Cout) dv = to_f(dy[m * g.out_pitch + co]);
                if (kval[j]) {
                    const int sy = map_coord(vy0 + ky[j], g.Hv, g.pad_mode, g.up, g.zi);
                    const int sx = map_coord(vx0 + kx[j], g.Wv, g.pad_mode, g.up, g.zi_w);
                    if (sy >= 0 && sx >= 0)
                        xv = act_apply(to_f(xn[((long long)sy * g.W + sx) * g.in_pitch + ci[j]]), g.pre_act);
                }
            }
            Ds[p][q + j] = dv;
            Xs[p][q + j] = xv;
        }
        __syncthreads();
        float part[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
        for (int pp = 0; pp < BP; ++pp) {
            const float4 a = *reinterpret_cast<const float4*>(&Ds[pp][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Xs[pp][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
        __syncthreads();
    }
    const int taps = g.KH * g.KW;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= g.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kg = k0 + tx * 4 + j;
            if (kg >= g.Ktot) continue;
            const int tap = kg / g.Cin, c = kg - tap * g.Cin;
            atomicAdd(&dw[((long long)co * g.Cin + c) * taps + tap], acc[i][j]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fold: gradient w.r.t. the padded/upsampled virtual input -> gradient w.r.t. the stored input.
//   dx[n,y,x,c] = act'(x[n,y,x,c]) * sum over virtual padded positions that the forward gather mapped to (y,x)
// ------------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void conv_fold_kernel(const T* __restrict__ dxp, const T* __restrict__ xin, T* __restrict__ dx, int N, int H,
                                 int W, int C, int pad, int pad_mode, int up, int pre_act) {
    const int Hv = H * up, Wv = W * up, Hp = Hv + 2 * pad, Wp = Wv + 2 * pad;
    const int cv = C / VEC;
    const long long total = (long long)N * H * W * cv;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cv) * VEC;
        long long t = idx / cv;
        const int xx = (int)(t % W);
        t /= W;
        const int yy = (int)(t % H);
        const int n = (int)(t / H);
        int pys[12], pxs[12], ny = 0, nx = 0;
        for (int u = 0; u < up; ++u) pys[ny++] = pad + yy * up + u;
        for (int u = 0; u < up; ++u) pxs[nx++] = pad + xx * up + u;
        if (pad_mode != PAD_ZERO) {
            for (int h = 0; h < 2 * pad; ++h) {
                const int py = h < pad ? h : Hv + h, px = h < pad ? h : Wv + h;
                if (ny < 12 && map_coord(py - pad, Hv, pad_mode, up, 1) == yy) pys[ny++] = py;
                if (nx < 12 && map_coord(px - pad, Wv, pad_mode, up, 1) == xx) pxs[nx++] = px;
            }
        }
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        for (int a = 0; a < ny; ++a)
            for (int b = 0; b < nx; ++b) {
                float v[VEC];
                ldv<VEC>(dxp + (((long long)n * Hp + pys[a]) * Wp + pxs[b]) * C + c, v);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] += v[i];
            }
        const long long o = (((long long)n * H + yy) * W + xx) * C + c;
        if (pre_act != ACT_NONE) {
            float xv[VEC];
            ldv<VEC>(xin + o, xv);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] *= act_grad(xv[i], pre_act);
        }
        stv<VEC>(dx + o, acc);
    }
}

// Fast form for zero / reflect padding with up in {1, 2} (every layer of the path): one grid row per stored image row, so the
// row candidates are block-uniform; the UP x UP interior taps are unconditional and the reflected ones (border pixels only)
// predicated.  Reflection: padded coordinate p < pad maps to v = pad - p, p >= V + pad to v = 2 (V - 1) - (p - pad).
template <typename T, int VEC, int UP>
__global__ void __launch_bounds__(256)
conv_fold_fast_kernel(const T* __restrict__ dxp, const T* __restrict__ xin, T* __restrict__ dx, int H, int W, int C, int pad,
                      int reflect, int pre_act) {
    const int Hv = H * UP, Wv = W * UP, Hp = Hv + 2 * pad, Wp = Wv + 2 * pad;
    const int cv = C / VEC;
    const int row = blockIdx.y, n = row / H, yy = row - n * H;
    int ys[3 * UP];
    bool yok[3 * UP];
#pragma unroll
    for (int u = 0; u < UP; ++u) {
        const int v = yy * UP + u;
        ys[3 * u] = pad + v;
        yok[3 * u] = true;
        ys[3 * u + 1] = pad - v;
        yok[3 * u + 1] = reflect && v >= 1 && v <= pad;
        ys[3 * u + 2] = pad + 2 * (Hv - 1) - v;
        yok[3 * u + 2] = reflect && v <= Hv - 2 && v >= Hv - 1 - pad;
    }
    const T* const img = dxp + (long long)n * Hp * Wp * C;
    const int items = W * cv;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < items; i += gridDim.x * 256) {
        const int xx = i / cv, c = (i - xx * cv) * VEC;
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
#pragma unroll
        for (int a = 0; a < 3 * UP; ++a) {
            if (!yok[a]) continue;                                      // block-uniform
            const T* const rp = img + (long long)ys[a] * Wp * C + c;
#pragma unroll
            for (int u = 0; u < UP; ++u) {
                const int v = xx * UP + u;
                float t[VEC];
                ldv<VEC>(rp + (pad + v) * C, t);
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[k] += t[k];
                if (reflect && v >= 1 && v <= pad) {
                    ldv<VEC>(rp + (pad - v) * C, t);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[k] += t[k];
                }
                if (reflect && v <= Wv - 2 && v >= Wv - 1 - pad) {
                    ldv<VEC>(rp + (pad + 2 * (Wv - 1) - v) * C, t);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[k] += t[k];
                }
            }
        }
        const long long o = ((long long)row * W + xx) * C + c;
        if (pre_act != ACT_NONE) {
            float xv[VEC];
            ldv<VEC>(xin + o, xv);
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] *= act_grad(xv[k], pre_act);
        }
        stv<VEC>(dx + o, acc);
    }
}

// column sums of a [M][pitch] matrix (bias gradient): out[c] += sum_m a[m][c]
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ a, float* __restrict__ out, long long M, int C,
                                                     int pitch, long long rows_per_block) {
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int r = threadIdx.x >> 5;
    const long long mbeg = blockIdx.y * rows_per_block, mend = min(M, mbeg + rows_per_block);
    float s = 0.f;
    if (c < C)
        for (long long m = mbeg + r; m < mend; m += 8) s += to_f(a[m * pitch + c]);
    red[r][threadIdx.x & 31] = s;
    __syncthreads();
    if (r == 0 && c < C) {
#pragma unroll
        for (int i = 1; i < 8; ++i) s += red[i][threadIdx.x];
        atomicAdd(&out[c], s);
    }
}

// OIHW fp32 parameter -> [O'][KH][KW][I'pad] in TW.  transpose_flip = 1 builds the dgrad operand
// (O' = Cin, I' = Cout, taps mirrored).
template <typename TW>
__global__ void pack_weight_kernel(const float* __restrict__ w, TW* __restrict__ out, int Cout, int Cin, int KH, int KW,
                                   int ipad, int transpose_flip) {
    const int Od = transpose_flip ? Cin : Cout, Id = transpose_flip ? Cout : Cin;
    const long long total = (long long)Od * KH * KW * ipad;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % ipad);
        long long t = idx / ipad;
        const int kx = (int)(t % KW);
        t /= KW;
        const int ky = (int)(t % KH);
        const int o = (int)(t / KH);
        float v = 0.f;
        if (i < Id) {
            if (transpose_flip)
                v = w[(((long long)i * Cin + o) * KH + (KH - 1 - ky)) * KW + (KW - 1 - kx)];
            else
                v = w[(((long long)o * Cin + i) * KH + ky) * KW + kx];
        }
        out[idx] = from_f<TW>(v);
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// host launchers (called from api.cu)
// ---------------------------------------------------------------------------------------------------
template <typename TI, typename TW, typename TO>
static int launch_fwd(const void* x, const void* w, const float* bias, const void* addend, void* y, const ConvGeom& g,
                      cudaStream_t st) {
    dim3 grid(cdiv(g.M, BM), cdiv(g.Cout, BN));
    conv_fwd_simt_kernel<TI, TW, TO><<<grid, 256, 0, st>>>((const TI*)x, (const TW*)w, bias, (const TO*)addend, (TO*)y, g);
    AFFGW_LAUNCH_CHECK("conv_fwd_simt");
    return 0;
}

int conv_fwd_simt(const void* x, int x_dt, const void* w, int w_dt, const float* bias, const void* addend, void* y,
                  int y_dt, const ConvGeom& g, cudaStream_t st) {
    const int key = x_dt * 4 + w_dt * 2 + y_dt;
    switch (key) {
        case 0: return launch_fwd<float, float, float>(x, w, bias, addend, y, g, st);
        case 1: return launch_fwd<float, float, bf16>(x, w, bias, addend, y, g, st);
        case 2: return launch_fwd<float, bf16, float>(x, w, bias, addend, y, g, st);
        case 3: return launch_fwd<float, bf16, bf16>(x, w, bias, addend, y, g, st);
        case 4: return launch_fwd<bf16, float, float>(x, w, bias, addend, y, g, st);
        case 5: return launch_fwd<bf16, float, bf16>(x, w, bias, addend, y, g, st);
        case 6: return launch_fwd<bf16, bf16, float>(x, w, bias, addend, y, g, st);
        case 7: return launch_fwd<bf16, bf16, bf16>(x, w, bias, addend, y, g, st);
    }
    affgw_set_error("conv_fwd_simt: bad dtype combination");
    return -1;
}

int conv_wgrad_simt(const void* x, int x_dt, const void* dy, int dy_dt, float* dw, const ConvGeom& g, cudaStream_t st) {
    const int gx = cdiv(g.Cout, 64), gy = cdiv(g.Ktot, 64);
    long long splits = (4LL * 148 + gx * gy - 1) / (gx * gy);
    const long long max_splits = (g.M + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    long long mps = (g.M + splits - 1) / splits;
    mps = (mps + 15) / 16 * 16;
    splits = (g.M + mps - 1) / mps;
    dim3 grid(gx, gy, (unsigned)splits);
    if (x_dt == AFFGW_F32 && dy_dt == AFFGW_F32)
        conv_wgrad_simt_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, (const float*)dy, dw, g, mps);
    else if (x_dt == AFFGW_BF16 && dy_dt == AFFGW_BF16)
        conv_wgrad_simt_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, dw, g, mps);
    else if (x_dt == AFFGW_F32 && dy_dt == AFFGW_BF16)
        conv_wgrad_simt_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)x, (const bf16*)dy, dw, g, mps);
    else
        conv_wgrad_simt_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)x, (const float*)dy, dw, g, mps);
    AFFGW_LAUNCH_CHECK("conv_wgrad_simt");
    return 0;
}

int conv_fold(const void* dxp, const void* xin, void* dx, int dt, int N, int H, int W, int C, int pad, int pad_mode,
              int up, int pre_act, cudaStream_t st) {
    const bool v8 = (C % 8 == 0);
    const long long total = (long long)N * H * W * (v8 ? C / 8 : C);
    if (v8 && dt == AFFGW_F32 && (up == 1 || up == 2) && pad_mode != PAD_REPLICATE && (long long)N * H <= 65535 &&
        (long long)(W * up + 2 * pad) * C < (1LL << 30)) {
        const int items = W * (C / 8);
        dim3 grid((unsigned)min(8, (items + 255) / 256), (unsigned)(N * H));
        const int reflect = pad_mode == PAD_REFLECT;
        if (up == 1) conv_fold_fast_kernel<float, 8, 1><<<grid, 256, 0, st>>>((const float*)dxp, (const float*)xin, (float*)dx, H, W, C, pad, reflect, pre_act);
        else conv_fold_fast_kernel<float, 8, 2><<<grid, 256, 0, st>>>((const float*)dxp, (const float*)xin, (float*)dx, H, W, C, pad, reflect, pre_act);
        AFFGW_LAUNCH_CHECK("conv_fold");
        return 0;
    }
    const int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    if (dt == AFFGW_F32) {
        if (v8) conv_fold_kernel<float, 8><<<blocks, 256, 0, st>>>((const float*)dxp, (const float*)xin, (float*)dx, N, H, W, C, pad, pad_mode, up, pre_act);
        else conv_fold_kernel<float, 1><<<blocks, 256, 0, st>>>((const float*)dxp, (const float*)xin, (float*)dx, N, H, W, C, pad, pad_mode, up, pre_act);
    } else {
        if (v8) conv_fold_kernel<bf16, 8><<<blocks, 256, 0, st>>>((const bf16*)dxp, (const bf16*)xin, (bf16*)dx, N, H, W, C, pad, pad_mode, up, pre_act);
        else conv_fold_kernel<bf16, 1><<<blocks, 256, 0, st>>>((const bf16*)dxp, (const bf16*)xin, (bf16*)dx, N, H, W, C, pad, pad_mode, up, pre_act);
    }
    AFFGW_LAUNCH_CHECK("conv_fold");
    return 0;
}

int colsum(const void* a, int dt, float* out, long long M, int C, int pitch, cudaStream_t st) {
    const int gx = cdiv(C, 32);
    long long gy = (2LL * 148 + gx - 1) / gx;
    const long long maxy = (M + 63) / 64;
    if (gy > maxy) gy = maxy;
    if (gy < 1) gy = 1;
    const long long rpb = (M + gy - 1) / gy;
    dim3 grid(gx, (unsigned)gy);
    if (dt == AFFGW_F32) colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)a, out, M, C, pitch, rpb);
    else colsum_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)a, out, M, C, pitch, rpb);
    AFFGW_LAUNCH_CHECK("colsum");
    return 0;
}

int pack_weight(const float* w, void* out, int out_dt, int Cout, int Cin, int KH, int KW, int ipad, int transpose_flip,
                cudaStream_t st) {
    const long long total = (long long)(transpose_flip ? Cin : Cout) * KH * KW * ipad;
    const int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    if (out_dt == AFFGW_F32) pack_weight_kernel<float><<<blocks, 256, 0, st>>>(w, (float*)out, Cout, Cin, KH, KW, ipad, transpose_flip);
    else pack_weight_kernel<bf16><<<blocks, 256, 0, st>>>(w, (bf16*)out, Cout, Cin, KH, KW, ipad, transpose_flip);
    AFFGW_LAUNCH_CHECK("pack_weight");
    return 0;
}
