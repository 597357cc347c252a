// Recogniser-side kernels of the GAN step:
//   * the label-smoothed KL loss the step forms from the recogniser's logits (reference network_tro.py:44-45,92-93 with
//     loss_tro.py:8-35);
//   * the recurrent / attention pieces of RecModel (modules_tro.py:610-638) that are not convolutions or GEMMs: the GRU cell
//     (torch.nn.GRU, encoder_vgg.py:700 / decoder.py:27), the location-attention energy and soft-max + context
//     (attention.py:132-160, decoder.py:36-40), Dropout2d's per-(sample, channel) scaling (encoder_vgg.py:709) and the
//     feature-map -> sequence permutation (encoder_vgg.py:711-713).  The GEMMs (input / hidden projections, attention
//     projections, output layer) run on the tcgen05 kernels through affgw_conv2d_* like every other linear layer.
// Everything here is fp32 and memory- or latency-bound: tensors of a few MB.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------- GRU cell
// r = sigmoid(i_r + h_r), z = sigmoid(i_z + h_z), n = tanh(i_n + r * h_n), h' = (1 - z) * n + z * h
// gi rows have their own pitch (a time slice of the [T*B, 3H] input projection), gh / h / outputs are dense.
__global__ void gru_cell_fwd_kernel(const float* __restrict__ gi, long long gi_pitch, const float* __restrict__ gh,
                                    const float* __restrict__ h, float* __restrict__ hout, int N, int H) {
    const long long total = (long long)N * H;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / H), j = (int)(i % H);
        const float* a = gi + n * gi_pitch;
        const float* b = gh + (long long)n * 3 * H;
        const float r = sigmoidf_(a[j] + b[j]);
        const float z = sigmoidf_(a[H + j] + b[H + j]);
        const float nn = tanhf(a[2 * H + j] + r * b[2 * H + j]);
        hout[i] = (1.f - z) * nn + z * h[i];
    }
}
__global__ void gru_cell_bwd_kernel(const float* __restrict__ dhout, const float* __restrict__ gi, long long gi_pitch,
                                    const float* __restrict__ gh, const float* __restrict__ h, float* __restrict__ dgi,
                                    float* __restrict__ dgh, float* __restrict__ dh, int N, int H) {
    const long long total = (long long)N * H;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / H), j = (int)(i % H);
        const float* a = gi + n * gi_pitch;
        const float* b = gh + (long long)n * 3 * H;
        const float r = sigmoidf_(a[j] + b[j]);
        const float z = sigmoidf_(a[H + j] + b[H + j]);
        const float hn = b[2 * H + j];
        const float nn = tanhf(a[2 * H + j] + r * hn);
        const float g = dhout[i], hp = h[i];
        const float dpn = g * (1.f - z) * (1.f - nn * nn);
        const float dpz = g * (hp - nn) * z * (1.f - z);
        const float dpr = dpn * hn * r * (1.f - r);
        float* da = dgi + (long long)n * 3 * H;
        float* db = dgh + (long long)n * 3 * H;
        da[j] = dpr;         db[j] = dpr;
        da[H + j] = dpz;     db[H + j] = dpz;
        da[2 * H + j] = dpn; db[2 * H + j] = dpn * r;
        dh[i] = g * z;
    }
}

// ---------------------------------------------------------------- small elementwise helpers
// y[n][p][c] = x[n][p][c] * m[n][c]   (nn.Dropout2d with the keep-mask / (1 - p) drawn by the caller)
__global__ void scale_nc_kernel(const float* __restrict__ x, const float* __restrict__ m, float* __restrict__ y, int N,
                                long long P, int C) {
    const long long total = (long long)N * P * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int n = (int)(i / ((long long)P * C));
        y[i] = x[i] * m[(long long)n * C + c];
    }
}
__global__ void mul2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = a[i] * b[i];
}
// feature map [B][H][W][C] (NHWC) <-> sequence [W][B][H*C]  (out.permute(3, 0, 2, 1).reshape(-1, B, H*C), encoder_vgg.py:711-713)
__global__ void map_seq_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int H, int W, int C, int to_seq) {
    const long long total = (long long)B * H * W * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;                       // i indexes the sequence layout
        const int c = (int)(r % C); r /= C;
        const int hh = (int)(r % H); r /= H;
        const int b = (int)(r % B); r /= B;
        const int w = (int)r;
        const long long m = (((long long)b * H + hh) * W + w) * C + c;
        if (to_seq) dst[i] = src[m]; else dst[m] = src[i];
    }
}

// ---------------------------------------------------------------- location attention
// energy[n][t] = v . tanh(e[s(n)][t][:] + hp[n][:] + loc[n][t][:]) + vb        one warp per (n, t)   (attention.py:145-158)
__global__ void attn_energy_fwd_kernel(const float* __restrict__ e, const long long* __restrict__ sidx, const float* __restrict__ hp,
                                       const float* __restrict__ loc, const float* __restrict__ v, const float* __restrict__ vb,
                                       float* __restrict__ energy, int N, int T, int F) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N * T) return;
    const int n = row / T, t = row % T;
    const float* er = e + ((long long)sidx[n] * T + t) * F;
    const float* hr = hp + (long long)n * F;
    const float* lr = loc + (long long)row * F;
    float s = 0.f;
    for (int f = lane; f < F; f += 32) s += v[f] * tanhf(er[f] + hr[f] + lr[f]);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) energy[row] = s + vb[0];
}
// g[f] = denergy * v[f] * (1 - tanh^2): dloc = g (written), de[s(n)][t] += g, dhp[n] += g, dv[f] += denergy * tanh, dvb += denergy
__global__ void attn_energy_bwd_kernel(const float* __restrict__ denergy, const float* __restrict__ e,
                                       const long long* __restrict__ sidx, const float* __restrict__ hp, const float* __restrict__ loc,
                                       const float* __restrict__ v, float* __restrict__ de, float* __restrict__ dhp,
                                       float* __restrict__ dloc, float* __restrict__ dv, float* __restrict__ dvb, int N, int T, int F) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N * T) return;
    const int n = row / T, t = row % T;
    const long long eo = ((long long)sidx[n] * T + t) * F;
    const float* hr = hp + (long long)n * F;
    const float* lr = loc + (long long)row * F;
    const float d = denergy[row];
    for (int f = lane; f < F; f += 32) {
        const float th = tanhf(e[eo + f] + hr[f] + lr[f]);
        const float g = d * v[f] * (1.f - th * th);
        dloc[(long long)row * F + f] = g;
        atomicAdd(de + eo + f, g);
        atomicAdd(dhp + (long long)n * F + f, g);
        atomicAdd(dv + f, d * th);
    }
    if (lane == 0) atomicAdd(dvb, d);
}
// attn[n][:] = softmax_t(energy[n][:]);  ctx[n][f] = sum_t attn[n][t] * enc[s(n)][t][f]        one block per n   (decoder.py:36-40)
constexpr int ATT_MAX_T = 64;
__global__ void attn_ctx_fwd_kernel(const float* __restrict__ energy, const float* __restrict__ enc, const long long* __restrict__ sidx,
                                    float* __restrict__ attn, float* __restrict__ ctx, int T, int F) {
    __shared__ float a[ATT_MAX_T];
    const int n = blockIdx.x;
    if (threadIdx.x == 0) {
        float mx = -INFINITY, s = 0.f;
        for (int t = 0; t < T; ++t) mx = fmaxf(mx, energy[(long long)n * T + t]);
        for (int t = 0; t < T; ++t) { a[t] = expf(energy[(long long)n * T + t] - mx); s += a[t]; }
        for (int t = 0; t < T; ++t) { a[t] /= s; attn[(long long)n * T + t] = a[t]; }
    }
    __syncthreads();
    const float* eb = enc + (long long)sidx[n] * T * F;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float s = 0.f;
        for (int t = 0; t < T; ++t) s += a[t] * eb[(long long)t * F + f];
        ctx[(long long)n * F + f] = s;
    }
}
// da[t] = dattn[n][t] + dctx[n] . enc[s][t];  denergy[t] = attn[t] * (da[t] - sum_s attn[s] da[s]);  denc[s][t][f] += attn[t] dctx[f]
__global__ void attn_ctx_bwd_kernel(const float* __restrict__ dattn, const float* __restrict__ dctx, const float* __restrict__ attn,
                                    const float* __restrict__ enc, const long long* __restrict__ sidx, float* __restrict__ denergy,
                                    float* __restrict__ denc, int T, int F) {
    __shared__ float da[ATT_MAX_T];
    __shared__ float red[8];
    const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long eo = (long long)sidx[n] * T * F;
    for (int t = warp; t < T; t += nw) {                       // one warp per time step: dot(dctx, enc[t])
        float s = 0.f;
        for (int f = lane; f < F; f += 32) s += dctx[(long long)n * F + f] * enc[eo + (long long)t * F + f];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) da[t] = s + (dattn ? dattn[(long long)n * T + t] : 0.f);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float dot = 0.f;
        for (int t = 0; t < T; ++t) dot += attn[(long long)n * T + t] * da[t];
        red[0] = dot;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) denergy[(long long)n * T + t] = attn[(long long)n * T + t] * (da[t] - red[0]);
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const float g = dctx[(long long)n * F + f];
        for (int t = 0; t < T; ++t) atomicAdd(denc + eo + (long long)t * F + f, attn[(long long)n * T + t] * g);
    }
}

// crit(log_softmax(x), target): KLDivLoss(reduction='sum') against the smoothed one-hot
//   t[v] = smoothing / (V - 2), t[target] = 1 - smoothing, t[pad] = 0, whole row 0 when target == pad      (loss_tro.py:19-27)
//   loss += sum_v xlogy(t[v], t[v]) - t[v] * logp[v]
// IEEE semantics are kept on purpose: a NaN logit makes the row's log-sum-exp NaN and with it 0 * logp = NaN, exactly what
// torch.nn.KLDivLoss returns (the recogniser's beam search can emit NaN-free logits only; see oracle/rec_oracle.py).
// One warp per row.
__device__ __forceinline__ float smooth_target(int v, long long tgt, int pad, float smoothing, int V) {
    if (tgt == pad || v == pad) return 0.f;
    return v == (int)tgt ? 1.f - smoothing : smoothing / (float)(V - 2);
}

__global__ void label_smooth_kl_fwd_kernel(const float* __restrict__ x, const long long* __restrict__ y, float* __restrict__ loss,
                                           int rows, int V, int pad, float smoothing, int* __restrict__ err) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long long)row * V;
    const long long t = y[row];
    if (t < 0 || t >= V) { if (lane == 0) *err = 1; return; }
    float mx = -INFINITY;
    bool nan = false;
    for (int c = lane; c < V; c += 32) { const float v = xr[c]; nan |= (v != v); mx = fmaxf(mx, v); }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    nan = __any_sync(0xffffffffu, nan);
    float s = 0.f;
    for (int c = lane; c < V; c += 32) s += expf(xr[c] - mx);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float lse = nan ? __int_as_float(0x7fc00000) : logf(s) + mx;       // fmaxf drops NaNs: restore torch's propagation
    float acc = 0.f;
    for (int c = lane; c < V; c += 32) {
        const float tv = smooth_target(c, t, pad, smoothing, V);
        acc += (tv > 0.f ? tv * logf(tv) : 0.f) - tv * (xr[c] - lse);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(loss, acc);
}

// d loss / d x[v] = g * (softmax[v] * sum_u t[u] - t[v])
__global__ void label_smooth_kl_bwd_kernel(const float* __restrict__ x, const long long* __restrict__ y, const float* __restrict__ gout,
                                           float* __restrict__ dx, int rows, int V, int pad, float smoothing) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long long)row * V;
    const long long t = y[row];
    float mx = -INFINITY;
    bool nan = false;
    for (int c = lane; c < V; c += 32) { const float v = xr[c]; nan |= (v != v); mx = fmaxf(mx, v); }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    nan = __any_sync(0xffffffffu, nan);
    float s = 0.f, tsum = 0.f;
    for (int c = lane; c < V; c += 32) {
        s += expf(xr[c] - mx);
        tsum += smooth_target(c, t, pad, smoothing, V);
    }
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
    }
    const float g = gout[0], inv = nan ? __int_as_float(0x7fc00000) : 1.f / s;
    for (int c = lane; c < V; c += 32)
        dx[(long long)row * V + c] = g * (expf(xr[c] - mx) * inv * tsum - smooth_target(c, t, pad, smoothing, V));
}

}  // namespace

static inline int rec_blocks(long long n) { long long b = (n + 255) / 256; return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b)); }

int gru_cell_fwd(const float* gi, long long gi_pitch, const float* gh, const float* h, float* hout, int N, int H, cudaStream_t st) {
    gru_cell_fwd_kernel<<<rec_blocks((long long)N * H), 256, 0, st>>>(gi, gi_pitch, gh, h, hout, N, H);
    AFFGW_LAUNCH_CHECK("gru_cell_fwd");
    return 0;
}
int gru_cell_bwd(const float* dhout, const float* gi, long long gi_pitch, const float* gh, const float* h, float* dgi, float* dgh,
                 float* dh, int N, int H, cudaStream_t st) {
    gru_cell_bwd_kernel<<<rec_blocks((long long)N * H), 256, 0, st>>>(dhout, gi, gi_pitch, gh, h, dgi, dgh, dh, N, H);
    AFFGW_LAUNCH_CHECK("gru_cell_bwd");
    return 0;
}
int scale_nc(const float* x, const float* m, float* y, int N, long long P, int C, cudaStream_t st) {
    scale_nc_kernel<<<rec_blocks((long long)N * P * C), 256, 0, st>>>(x, m, y, N, P, C);
    AFFGW_LAUNCH_CHECK("scale_nc");
    return 0;
}
int mul2(const float* a, const float* b, float* y, long long n, cudaStream_t st) {
    mul2_kernel<<<rec_blocks(n), 256, 0, st>>>(a, b, y, n);
    AFFGW_LAUNCH_CHECK("mul2");
    return 0;
}
int map_seq(const float* src, float* dst, int B, int H, int W, int C, int to_seq, cudaStream_t st) {
    map_seq_kernel<<<rec_blocks((long long)B * H * W * C), 256, 0, st>>>(src, dst, B, H, W, C, to_seq);
    AFFGW_LAUNCH_CHECK("map_seq");
    return 0;
}
int attn_energy_fwd(const float* e, const long long* sidx, const float* hp, const float* loc, const float* v, const float* vb,
                    float* energy, int N, int T, int F, cudaStream_t st) {
    attn_energy_fwd_kernel<<<cdiv((long long)N * T, 8), 256, 0, st>>>(e, sidx, hp, loc, v, vb, energy, N, T, F);
    AFFGW_LAUNCH_CHECK("attn_energy_fwd");
    return 0;
}
int attn_energy_bwd(const float* denergy, const float* e, const long long* sidx, const float* hp, const float* loc, const float* v,
                    float* de, float* dhp, float* dloc, float* dv, float* dvb, int N, int T, int F, cudaStream_t st) {
    attn_energy_bwd_kernel<<<cdiv((long long)N * T, 8), 256, 0, st>>>(denergy, e, sidx, hp, loc, v, de, dhp, dloc, dv, dvb, N, T, F);
    AFFGW_LAUNCH_CHECK("attn_energy_bwd");
    return 0;
}
int attn_ctx_fwd(const float* energy, const float* enc, const long long* sidx, float* attn, float* ctx, int N, int T, int F,
                 cudaStream_t st) {
    if (T > ATT_MAX_T) { affgw_set_error("attn_ctx: at most %d encoder steps", ATT_MAX_T); return -1; }
    attn_ctx_fwd_kernel<<<N, 256, 0, st>>>(energy, enc, sidx, attn, ctx, T, F);
    AFFGW_LAUNCH_CHECK("attn_ctx_fwd");
    return 0;
}
int attn_ctx_bwd(const float* dattn, const float* dctx, const float* attn, const float* enc, const long long* sidx, float* denergy,
                 float* denc, int N, int T, int F, cudaStream_t st) {
    if (T > ATT_MAX_T) { affgw_set_error("attn_ctx: at most %d encoder steps", ATT_MAX_T); return -1; }
    attn_ctx_bwd_kernel<<<N, 256, 0, st>>>(dattn, dctx, attn, enc, sidx, denergy, denc, T, F);
    AFFGW_LAUNCH_CHECK("attn_ctx_bwd");
    return 0;
}

int label_smooth_kl_fwd(const float* x, const long long* y, float* loss, int rows, int V, int pad, float smoothing, int* err,
                        cudaStream_t st) {
    cudaMemsetAsync(loss, 0, sizeof(float), st);
    label_smooth_kl_fwd_kernel<<<cdiv(rows, 8), 256, 0, st>>>(x, y, loss, rows, V, pad, smoothing, err);
    AFFGW_LAUNCH_CHECK("label_smooth_kl_fwd");
    return 0;
}
int label_smooth_kl_bwd(const float* x, const long long* y, const float* gout, float* dx, int rows, int V, int pad, float smoothing,
                        cudaStream_t st) {
    label_smooth_kl_bwd_kernel<<<cdiv(rows, 8), 256, 0, st>>>(x, y, gout, dx, rows, V, pad, smoothing);
    AFFGW_LAUNCH_CHECK("label_smooth_kl_bwd");
    return 0;
}
