// Shared device/host helpers for libaffgw (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

enum { AFFGW_F32 = 0, AFFGW_BF16 = 1 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_TANH = 3 };
enum { PAD_ZERO = 0, PAD_REFLECT = 1, PAD_REPLICATE = 2 };

// ---- error plumbing (thread-local last-error string, see affgw_last_error) ----
void affgw_set_error(const char* fmt, ...);
#define AFFGW_CHECK(cond, ...)                \
    do {                                      \
        if (!(cond)) {                        \
            affgw_set_error(__VA_ARGS__);     \
            return -1;                        \
        }                                     \
    } while (0)
#define AFFGW_LAUNCH_CHECK(name)                                                       \
    do {                                                                               \
        cudaError_t e_ = cudaGetLastError();                                           \
        if (e_ != cudaSuccess) {                                                       \
            affgw_set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));    \
            return -2;                                                                 \
        }                                                                              \
        affgw_count_launch();                                                          \
    } while (0)
void affgw_count_launch();

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- scalar / 8-wide vector element access, fp32 math everywhere ----
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}
// VEC-generic: VEC == 8 uses 128-bit accesses, VEC == 1 is the scalar fallback for odd channel counts.
template <int VEC, typename T> __device__ __forceinline__ void ldv(const T* p, float (&v)[VEC]) {
    if constexpr (VEC == 8) ld8(p, v); else v[0] = to_f(p[0]);
}
template <int VEC, typename T> __device__ __forceinline__ void stv(T* p, const float (&v)[VEC]) {
    if constexpr (VEC == 8) st8(p, v); else p[0] = from_f<T>(v[0]);
}

__device__ __forceinline__ float act_apply(float v, int act) {
    switch (act) {
        case ACT_RELU: return v > 0.f ? v : 0.f;
        case ACT_LRELU: return v > 0.f ? v : 0.2f * v;
        case ACT_TANH: return tanhf(v);
        default: return v;
    }
}
// derivative of act at pre-activation value z (tanh is handled by its caller from the output)
__device__ __forceinline__ float act_grad(float z, int act) {
    switch (act) {
        case ACT_RELU: return z > 0.f ? 1.f : 0.f;
        case ACT_LRELU: return z > 0.f ? 1.f : 0.2f;
        default: return 1.f;
    }
}

// Virtual-input coordinate -> source coordinate (or -1 = contributes zero).
//   V        extent of the virtual input (H*up, or (H-1)*zi+1 when zero-inserting)
//   pad_mode how coordinates outside [0, V) are resolved
//   up       nearest-neighbour upsampling factor folded into the gather (nn.Upsample(scale_factor=2))
//   zi       zero-insertion factor (strided-conv dgrad): only multiples of zi carry data
__device__ __forceinline__ int map_coord(int v, int V, int pad_mode, int up, int zi) {
    if (v < 0 || v >= V) {
        if (pad_mode == PAD_ZERO) return -1;
        if (pad_mode == PAD_REFLECT) v = v < 0 ? -v : 2 * (V - 1) - v;
        else v = v < 0 ? 0 : V - 1;
    }
    if (zi > 1) return (v % zi == 0) ? v / zi : -1;
    return up > 1 ? v / up : v;
}

// q = n / d for 0 <= n < 2^31 with one 32 x 32 -> 64 bit multiply: mul = ceil(2^k / d), k = 31 + ceil(log2 d)
struct FastDiv {
    uint32_t mul, k, d;
};
inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    int s = 0;
    while ((1LL << s) < d) ++s;
    f.k = 31 + s;
    f.mul = (uint32_t)(((1ULL << f.k) + (uint64_t)d - 1) / (uint64_t)d);
    f.d = (uint32_t)d;
    return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv& f) { return (uint32_t)(((uint64_t)n * f.mul) >> f.k); }

// Geometry of one convolution call (all tensors NHWC, channels contiguous).
struct ConvGeom {
    int N, H, W, Cin;        // stored input
    int Cout, KH, KW;
    int stride, pad, pad_mode, up, zi;   // stride / zi: rows
    int stride_w, zi_w;                  // columns (differ from the row values only for Resnet18.py's (2,1) stem)
    int Ho, Wo;              // output extent
    int in_pitch, out_pitch; // elements between consecutive pixels (>= Cin / Cout)
    int pre_act, post_act;
    int Hv, Wv;              // virtual input extent (derived)
    int Ktot;                // KH*KW*Cin
    long long M;             // N*Ho*Wo
};
