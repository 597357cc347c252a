// Recogniser-side kernels of the GAN step: the label-smoothed KL loss the step forms from the
// recogniser's logits (reference network_tro.py:44-45,92-93 with loss_tro.py:8-35).
#include "common.cuh"

namespace {

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
