// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, fp32 accumulators in TMEM).
//
//   D[128 pixels][BN couts] += A[128 pixels][64 k] * B[BN couts][64 k]^T          (bf16 x bf16 -> fp32)
//
// k runs over the flattened (filter tap, input channel) axis, so any stored channel count that is a multiple of 8
// works (16-byte gather granules): the 1 -> 16 stem (channels padded to 8), the 16/32-channel discriminator blocks and
// the 512-channel VGG / decoder layers all run through this one kernel.
//
// Persistent, warp-specialised CTA (one or two per SM, grid = resident CTAs, static round-robin over output tiles):
//   warps 0-3  producers.  A (activations, NHWC bf16 operand planes) is gathered straight from global memory into the
//              128B-swizzled K-major shared-memory image UMMA expects with 16-byte cp.async copies whose source address
//              already contains the convolution's index map (zero / reflect / replicate padding, nearest x2 upsampling,
//              zero insertion for strided dgrad) - the padded or upsampled tensor is never materialised.  B (weights) is
//              pre-packed once per optimiser step into exactly that shared-memory image (affgw_pack_weight_tc), so one
//              cp.async.bulk (TMA engine, mbarrier complete_tx) moves a whole stage of weight tiles.
//   warp  4    one elected thread issues tcgen05.mma; tcgen05.commit releases the smem stage / publishes the accumulator.
//   warps 5-8  epilogue: tcgen05.ld -> +bias -> +addend -> activation -> vector stores, overlapped with the next tile's
//              main loop through a double-buffered TMEM accumulator.
//
// NPASS = 3 is the split-bf16 product  a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo  (hi = bf16(v), lo = bf16(v - hi)):
// both operands arrive as two bf16 planes, three MMAs accumulate into the same TMEM tile.  It keeps ~16 mantissa bits per
// operand, which is what the 2e-2 image bar needs on this network (the VGG-IN stack amplifies a 2^-9 operand rounding by
// ~1.23x per layer at random init, DESIGN.md "precision").
//
// Replaces the cuDNN convolutions behind reference blocks.py:148 / vgg_tro_channel3_modi.py:47 /
// modules_tro.py:252-259; the same kernel computes dgrad.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr float F16_W_SCALE = 256.f;         // fp16 weight tiles hold w * 2^8 (see conv_shift.cu)
__device__ __forceinline__ uint16_t f16_bits(float v) {
    return __half_as_ushort(__float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)));
}
__device__ __forceinline__ float f16_val(uint16_t b) { return __half2float(__ushort_as_half(b)); }

constexpr int BM = 128, BK = 64;
constexpr int A_PLANE_BYTES = BM * BK * 2;  // 16 KB
constexpr int NUM_PRODUCER_THREADS = 128;
constexpr int MMA_WARP = 4;
constexpr int NUM_THREADS = 32 * 9;         // 4 producer warps + 1 MMA warp + 4 epilogue warps

constexpr int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

template <int BN, int NPASS> struct TcCfg {
    static constexpr int NPL = NPASS == 3 ? 2 : 1;
    static constexpr int A_BYTES = NPL * A_PLANE_BYTES;
    static constexpr int B_PLANE_BYTES = BN * BK * 2;
    static constexpr int B_BYTES = NPL * B_PLANE_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BUDGET = NPASS == 3 ? 200 * 1024 : 100 * 1024;     // 1 or 2 CTAs per SM
    static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
    static constexpr int CTAS_PER_SM = NPASS == 3 ? 1 : 2;
    static constexpr int TMEM_COLS = pow2_cols(2 * BN);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows are 128 B (64 bf16), 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset between 8-row groups
    d |= 1ull << 46;                                    // descriptor version (sm_100)
    d |= 2ull << 61;                                    // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D = f32, A and B both bf16 (fmt 0) or both fp16 (fmt 1), both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n, int fmt = 0) {
    return (1u << 4) | (fmt ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct TcArgs {
    const bf16* x;             // operand planes [NPL][N*H*W][Cin]
    long long x_plane;         // elements between the hi and lo plane
    const bf16* w_tiles;       // [n_tiles][KB][NPL][BN][64] swizzled
    const float* bias;
    const void* addend;
    void* y;
    ConvGeom g;                // g.Cin = stored channels of the planes (== pitch), g.Ktot = taps * g.Cin
    int KB, n_tiles;
    int ksplit, kb_per_split;  // split of the k-block range over CTAs (tiny-M layers); partial sums are reduced with atomics
    long long total_tiles;
    int vec_ok;                // 16-column vector stores allowed (Cout % 16 == 0 and aligned pitch)
    int fmt;                   // 0 = bf16 operands, 1 = fp16 operands
    float alpha;               // result = accumulator * alpha * (alpha_dev ? *alpha_dev : 1): undoes the fp16 operand scales
    const float* alpha_dev;
};

template <typename TO> __device__ __forceinline__ void store16(TO* dst, const float (&v)[16]);
template <> __device__ __forceinline__ void store16<float>(float* dst, const float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
template <> __device__ __forceinline__ void store16<bf16>(bf16* dst, const float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
        reinterpret_cast<uint4*>(dst)[i] = u;
    }
}
template <typename TO> __device__ __forceinline__ void load16(const TO* src, float (&v)[16]);
template <> __device__ __forceinline__ void load16<float>(const float* src, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 f = reinterpret_cast<const float4*>(src)[i];
        v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
    }
}
template <> __device__ __forceinline__ void load16<bf16>(const bf16* src, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float t[8];
        ld8(src + 8 * i, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * i + j] = t[j];
    }
}

// ---------------------------------------------------------------------------------------------------- the kernel
template <int BN, int NPASS, typename TO>
__global__ void __launch_bounds__(NUM_THREADS, TcCfg<BN, NPASS>::CTAS_PER_SM)
conv_igemm_tcgen05_kernel(const __grid_constant__ TcArgs a) {
    using Cfg = TcCfg<BN, NPASS>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int NPL = Cfg::NPL;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base address word
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + i); };
    auto tmem_empty_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + 2 + i); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
    auto a_smem = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
    auto b_smem = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };

    const ConvGeom& g = a.g;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int KB = a.KB;

    if (tid == MMA_WARP * 32) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), NUM_PRODUCER_THREADS);
            mbar_init(empty_bar(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tmem_full_bar(i), 1);
            mbar_init(tmem_empty_bar(i), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        __syncwarp();
        tmem_alloc(tmem_ptr_addr, Cfg::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

    if (warp < 4) {
        // ============================== A/B producer ==============================
        const int chunk = tid & 7;        // 16-byte chunk (8 k elements) of the 128-byte row
        const int rg = tid >> 3;          // rows rg, rg+16, ..., rg+112
        const bool uniform_tap = (g.Cin % BK) == 0;     // all 64 k of a block belong to one filter tap
        // signal stage i - LAG after issuing stage i; one stage of slack (not STAGES - 1) so that issuing the next stage does not
        // have to wait for the MMAs of the stage just signalled (see conv_tc_wgrad.cu)
        constexpr int LAG = STAGES > 2 ? STAGES - 2 : 1;
        uint32_t it = 0;                  // k-block counter across tiles: stage = it % STAGES
        for (long long t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            const int ks = (int)(t % a.ksplit);
            const long long tmn = t / a.ksplit;
            const int nt = (int)(tmn % a.n_tiles);
            const long long m0 = (tmn / a.n_tiles) * BM;
            const int kb0 = ks * a.kb_per_split, kb1 = min(KB, kb0 + a.kb_per_split);
            int vy0[8], vx0[8], nbase[8];
            bool mvalid[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long m = m0 + rg + 16 * i;
                mvalid[i] = m < g.M;
                int ox = 0, oy = 0, n = 0;
                if (mvalid[i]) {
                    ox = (int)(m % g.Wo);
                    const long long q = m / g.Wo;
                    oy = (int)(q % g.Ho);
                    n = (int)(q / g.Ho);
                }
                vy0[i] = oy * g.stride - g.pad;
                vx0[i] = ox * g.stride_w - g.pad;
                nbase[i] = n * g.H * g.W;
            }
            const bf16* wt = a.w_tiles + (size_t)nt * KB * (NPL * BN * BK);
            int ky = 0, kx = 0, c0 = 0;   // uniform-tap cursor, positioned at this split's first k-block
            int sy[8], sx[8];
            int off[8];                   // uniform tap: element offset of row i's source pixel (channel 0), -1 = zero row
            auto refresh = [&]() {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    off[i] = (mvalid[i] && sy[i] >= 0 && sx[i] >= 0) ? (nbase[i] + sy[i] * g.W + sx[i]) * g.Cin : -1;
            };
            if (uniform_tap) {
                const int tap0 = (kb0 * BK) / g.Cin;
                c0 = kb0 * BK - tap0 * g.Cin;
                ky = tap0 / g.KW;
                kx = tap0 - ky * g.KW;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    sy[i] = map_coord(vy0[i] + ky, g.Hv, g.pad_mode, g.up, g.zi);
                    sx[i] = map_coord(vx0[i] + kx, g.Wv, g.pad_mode, g.up, g.zi_w);
                }
                refresh();
            }
            const uint32_t dst_off = rg * 128 + ((chunk ^ (rg & 7)) << 4);     // rows rg + 16 i keep (r & 7)
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                if (tid == 0) {
                    mbar_expect_tx(full_bar(s), Cfg::B_BYTES);
                    bulk_copy_g2s(b_smem(s), wt + (size_t)kb * (NPL * BN * BK), Cfg::B_BYTES, full_bar(s));
                }
                const uint32_t a_base = a_smem(s) + dst_off;
                if (uniform_tap) {
                    const bf16* const xc = a.x + c0 + chunk * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const bool ok = off[i] >= 0;
                        const bf16* src = xc + (ok ? off[i] : 0);
                        const uint32_t dst = a_base + i * 2048;
                        cp_async_16(dst, src, ok ? 16u : 0u);
                        if (NPL == 2) cp_async_16(dst + A_PLANE_BYTES, src + a.x_plane, ok ? 16u : 0u);
                    }
                } else {
                    const int kidx = kb * BK + chunk * 8;
                    const bool kvalid = kidx < g.Ktot;
                    const int tap = kidx / g.Cin;
                    const int coff = kidx - tap * g.Cin;
                    ky = tap / g.KW;
                    kx = tap - ky * g.KW;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int yy = map_coord(vy0[i] + ky, g.Hv, g.pad_mode, g.up, g.zi);
                        const int xx = map_coord(vx0[i] + kx, g.Wv, g.pad_mode, g.up, g.zi_w);
                        const bool ok = mvalid[i] && kvalid && yy >= 0 && xx >= 0;
                        const bf16* src = ok ? a.x + ((size_t)(nbase[i] + yy * g.W + xx) * g.Cin + coff) : a.x;
                        const uint32_t dst = a_base + i * 2048;
                        cp_async_16(dst, src, ok ? 16u : 0u);
                        if (NPL == 2) cp_async_16(dst + A_PLANE_BYTES, ok ? src + a.x_plane : src, ok ? 16u : 0u);
                    }
                }
                cp_async_commit();
                if (it >= (uint32_t)LAG) {
                    cp_async_wait<LAG>();
                    fence_proxy_async();
                    mbar_arrive(full_bar((it - LAG) % STAGES));
                }
                if (uniform_tap) {      // advance (c0, kx, ky); the source pixels change only when the tap does
                    c0 += BK;
                    if (c0 == g.Cin) {
                        c0 = 0;
                        if (++kx == g.KW) {
                            kx = 0;
                            ++ky;
#pragma unroll
                            for (int i = 0; i < 8; ++i) sy[i] = map_coord(vy0[i] + ky, g.Hv, g.pad_mode, g.up, g.zi);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) sx[i] = map_coord(vx0[i] + kx, g.Wv, g.pad_mode, g.up, g.zi_w);
                        refresh();
                    }
                }
            }
        }
        // drain the last LAG stages
        cp_async_wait<0>();
        fence_proxy_async();
        for (uint32_t j = (it > (uint32_t)LAG ? it - LAG : 0u); j < it; ++j) mbar_arrive(full_bar(j % STAGES));
    } else if (warp == MMA_WARP) {
        // ============================== MMA issuer ==============================
        {   // every lane runs the loops (warp-uniform), one elected lane issues: see umma_bf16_elect
            const uint32_t idesc = make_idesc_bf16(BN, a.fmt);
            uint32_t it = 0, tl = 0;
            for (long long t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++tl) {
                const uint32_t acc = tl & 1u, acc_ph = (tl >> 1) & 1u;
                mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                const int ks = (int)(t % a.ksplit);
                const int nkb = min(KB, (ks + 1) * a.kb_per_split) - ks * a.kb_per_split;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint64_t a_hi = make_kmajor_sw128_desc(a_smem(s));
                    const uint64_t b_hi = make_kmajor_sw128_desc(b_smem(s));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)   // +32 bytes (encoded >>4) per 16-element K step inside the swizzle atom
                        umma_bf16_elect(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                    if (NPASS == 3) {
                        const uint64_t a_lo = make_kmajor_sw128_desc(a_smem(s) + A_PLANE_BYTES);
                        const uint64_t b_lo = make_kmajor_sw128_desc(b_smem(s) + Cfg::B_PLANE_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) umma_bf16_elect(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) umma_bf16_elect(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
                    }
                    umma_commit_elect(empty_bar(s));
                }
                umma_commit_elect(tmem_full_bar(acc));
            }
        }
    } else {
        // ============================== epilogue ==============================
        const int q = warp & 3;           // TMEM lane quarter this warp may read
        const float alpha = a.alpha * (a.alpha_dev ? __ldg(a.alpha_dev) : 1.f);
        TO* const y = reinterpret_cast<TO*>(a.y);
        const TO* const addend = reinterpret_cast<const TO*>(a.addend);
        uint32_t tl = 0;
        for (long long t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++tl) {
            const uint32_t acc = tl & 1u, acc_ph = (tl >> 1) & 1u;
            const int ks = (int)(t % a.ksplit);
            const long long tmn = t / a.ksplit;
            const int n0 = (int)(tmn % a.n_tiles) * BN;
            const long long m = (tmn / a.n_tiles) * BM + q * 32 + lane;
            const bool first_split = ks == 0;
            mbar_wait(tmem_full_bar(acc), acc_ph);
            tc_fence_after();
#pragma unroll 1
            for (int j = 0; j < BN / 16; ++j) {
                const int nb = n0 + j * 16;
                if (nb >= g.Cout) break;                       // warp-uniform
                uint32_t raw[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)(j * 16), raw);
                if (m < g.M && a.ksplit > 1) {
                    // split k range: fp32 partial sums into the zeroed output; bias / addend ride with the first split
                    if constexpr (sizeof(TO) == 4) {
                        float* yr = reinterpret_cast<float*>(y) + m * g.out_pitch + nb;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (nb + i < g.Cout) {
                                float r = __uint_as_float(raw[i]) * alpha;
                                if (first_split) {
                                    if (a.bias) r += __ldg(a.bias + nb + i);
                                    if (addend) r += to_f(addend[m * g.out_pitch + nb + i]);
                                }
                                atomicAdd(yr + i, r);
                            }
                        }
                    }
                } else if (m < g.M) {
                    float v[16];
                    if (a.vec_ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(raw[i]), alpha, a.bias ? __ldg(a.bias + nb + i) : 0.f);
                        if (addend) {
                            float ad[16];
                            load16<TO>(addend + m * g.out_pitch + nb, ad);
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] += ad[i];
                        }
                        if (g.post_act != ACT_NONE) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = act_apply(v[i], g.post_act);
                        }
                        store16<TO>(y + m * g.out_pitch + nb, v);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (nb + i < g.Cout) {
                                float r = fmaf(__uint_as_float(raw[i]), alpha, a.bias ? __ldg(a.bias + nb + i) : 0.f);
                                if (addend) r += to_f(addend[m * g.out_pitch + nb + i]);
                                y[m * g.out_pitch + nb + i] = from_f<TO>(act_apply(r, g.post_act));
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// weights: logical [O'][taps * I'pad] -> per (n-tile, k-block, plane) [BN][64] tiles with the 128B swizzle already applied;
// plane 0 = bf16(w), plane 1 = bf16(w - plane0) (written when passes == 3)
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int KH, int KW,
                                      int ipad, int transpose_flip, int BN, int ntiles, int KB, int npl, int fmt) {
    const int Od = transpose_flip ? Cin : Cout, Id = transpose_flip ? Cout : Cin;
    const int taps = KH * KW;
    const long long total = (long long)ntiles * KB * BN * BK;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        // destination-major decode so that writes are coalesced
        const int within = (int)(idx % (BN * BK));
        const long long tile = idx / (BN * BK);
        const int kb = (int)(tile % KB), nt = (int)(tile / KB);
        const int r = within / BK;
        const int pos = within % BK;                   // physical element position inside the 128-byte row
        const int e = (((pos >> 3) ^ (r & 7)) << 3) | (pos & 7);   // logical k element stored there
        const int o = nt * BN + r;
        const int kidx = kb * BK + e;
        const int tap = kidx / ipad, i = kidx - tap * ipad;
        float v = 0.f;
        if (o < Od && i < Id && tap < taps) {
            const int ky = tap / KW, kx = tap % KW;
            if (transpose_flip)
                v = w[(((long long)i * Cin + o) * KH + (KH - 1 - ky)) * KW + (KW - 1 - kx)];
            else
                v = w[(((long long)o * Cin + i) * KH + ky) * KW + kx];
        }
        bf16* dst = out + (tile * npl) * (BN * BK) + within;
        if (fmt) {
            const float sv = v * F16_W_SCALE;
            const uint16_t hb = f16_bits(sv);
            reinterpret_cast<uint16_t*>(dst)[0] = hb;
            if (npl == 2) reinterpret_cast<uint16_t*>(dst)[BN * BK] = f16_bits(sv - f16_val(hb));
        } else {
            const bf16 hi = __float2bfloat16_rn(v);
            dst[0] = hi;
            if (npl == 2) dst[BN * BK] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
    }
}

// x [rows][pitch] (fp32 or bf16) -> operand planes [npl][rows][c_store] bf16: pre-activation applied, channels >= C
// zero-filled, plane 1 = bf16 remainder.
template <typename T>
__global__ void split_planes_kernel(const T* __restrict__ x, bf16* __restrict__ planes, long long rows, int C, int pitch,
                                    int c_store, int npl, int pre_act, int fmt, const float* __restrict__ scale_dev) {
    const float scale = scale_dev ? __ldg(scale_dev) : 1.f;
    const int groups = c_store / 8;
    const long long total = rows * groups;
    const long long plane = rows * (long long)c_store;
    const bool vec = (C % 8 == 0) && (pitch % 8 == 0);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / groups;
        const int c = (int)(idx - r * groups) * 8;
        float v[8];
        if (vec && c + 8 <= C) {
            ld8(x + r * pitch + c, v);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (c + i < C) ? to_f(x[r * pitch + c + i]) : 0.f;
        }
        if (fmt) {
            uint4 hi, lo;
            __half2* hh = reinterpret_cast<__half2*>(&hi);
            __half2* ll = reinterpret_cast<__half2*>(&lo);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float a0 = fminf(fmaxf(act_apply(v[2 * i], pre_act) * scale, -65504.f), 65504.f);
                const float a1 = fminf(fmaxf(act_apply(v[2 * i + 1], pre_act) * scale, -65504.f), 65504.f);
                hh[i] = __floats2half2_rn(a0, a1);
                const float2 f = __half22float2(hh[i]);
                ll[i] = __floats2half2_rn(a0 - f.x, a1 - f.y);
            }
            *reinterpret_cast<uint4*>(planes + r * c_store + c) = hi;
            if (npl == 2) *reinterpret_cast<uint4*>(planes + plane + r * c_store + c) = lo;
            continue;
        }
        float lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = act_apply(v[i], pre_act);
            const float h = __bfloat162float(__float2bfloat16_rn(v[i]));
            lo[i] = v[i] - h;
        }
        st8(planes + r * c_store + c, v);
        if (npl == 2) st8(planes + plane + r * c_store + c, lo);
    }
}

template <int BN, int NPASS, typename TO>
int launch_tc(const TcArgs& args, cudaStream_t st) {
    using Cfg = TcCfg<BN, NPASS>;
    static bool configured = false;
    auto kern = conv_igemm_tcgen05_kernel<BN, NPASS, TO>;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) {
            affgw_set_error("conv_fwd_tc: cannot reserve %d bytes of shared memory", Cfg::SMEM_BYTES);
            return -2;
        }
        configured = true;
    }
    const long long resident = 148LL * Cfg::CTAS_PER_SM;
    const unsigned grid = (unsigned)(args.total_tiles < resident ? args.total_tiles : resident);
    kern<<<grid, NUM_THREADS, Cfg::SMEM_BYTES, st>>>(args);
    AFFGW_LAUNCH_CHECK("conv_igemm_tcgen05");
    return 0;
}

template <int BN, typename TO>
int launch_tc_p(const TcArgs& args, int passes, cudaStream_t st) {
    return passes == 3 ? launch_tc<BN, 3, TO>(args, st) : launch_tc<BN, 1, TO>(args, st);
}
template <typename TO>
int launch_tc_n(const TcArgs& args, int bn, int passes, cudaStream_t st) {
    switch (bn) {
        case 16: return launch_tc_p<16, TO>(args, passes, st);
        case 32: return launch_tc_p<32, TO>(args, passes, st);
        case 64: return launch_tc_p<64, TO>(args, passes, st);
        default: return launch_tc_p<128, TO>(args, passes, st);
    }
}

}  // namespace

// tile width for an output-channel count
int conv_tc_block_n(int cout) { return cout > 64 ? 128 : cout > 32 ? 64 : cout > 16 ? 32 : 16; }

// g.Cin is the STORED channel count of the operand planes
int conv_tc_ok(const ConvGeom& g) {
    if (g.Cin % 8 != 0 || g.in_pitch != g.Cin) return 0;
    if ((long long)g.N * g.H * g.W * g.Cin >= (1LL << 31)) return 0;      // 32-bit element offsets in the gather
    if ((long long)g.N * g.H * g.W >= (1LL << 31)) return 0;
    return conv_tc_block_n(g.Cout);
}

int conv_fwd_tc(const void* x_planes, long long plane_stride, const void* w_tiles, const float* bias, const void* addend,
                void* y, int y_dt, const ConvGeom& g, int passes, int fmt, const float* alpha_dev, cudaStream_t st) {
    const int bn = conv_tc_ok(g);
    if (!bn || (passes != 1 && passes != 3)) {
        affgw_set_error("conv_fwd_tc: unsupported shape (stored Cin %d, pitch %d, passes %d)", g.Cin, g.in_pitch, passes);
        return -1;
    }
    TcArgs a;
    a.x = (const bf16*)x_planes;
    a.x_plane = plane_stride;
    a.w_tiles = (const bf16*)w_tiles;
    a.bias = bias;
    a.addend = addend;
    a.y = y;
    a.g = g;
    a.g.Ktot = g.KH * g.KW * g.Cin;
    a.KB = (a.g.Ktot + BK - 1) / BK;
    a.n_tiles = (g.Cout + bn - 1) / bn;
    const long long mn_tiles = ((g.M + BM - 1) / BM) * a.n_tiles;
    // tiny-M layers (the 2x7 / 4x14 maps of the discriminator) give a handful of output tiles with a long k loop each:
    // spread the k-blocks over the idle SMs
    a.ksplit = 1;
    a.kb_per_split = a.KB;
    if (y_dt == AFFGW_F32 && g.post_act == ACT_NONE && mn_tiles <= 74 && a.KB >= 8) {
        int want = (int)(148 / mn_tiles);
        if (want > a.KB / 4) want = a.KB / 4;
        if (want > 1) {
            a.kb_per_split = (a.KB + want - 1) / want;
            a.ksplit = (a.KB + a.kb_per_split - 1) / a.kb_per_split;
        }
    }
    a.total_tiles = mn_tiles * a.ksplit;
    if (a.ksplit > 1 && cudaMemsetAsync(y, 0, (size_t)g.M * g.out_pitch * sizeof(float), st) != cudaSuccess) {
        affgw_set_error("conv_fwd_tc: memset failed");
        return -2;
    }
    const int esz = y_dt == AFFGW_F32 ? 4 : 2;
    a.vec_ok = (g.Cout % 16 == 0) && ((g.out_pitch * esz) % 16 == 0) && (((uintptr_t)y) % 16 == 0) &&
               (!addend || ((uintptr_t)addend) % 16 == 0);
    a.fmt = fmt;
    a.alpha = fmt ? 1.f / F16_W_SCALE : 1.f;
    a.alpha_dev = alpha_dev;
    return y_dt == AFFGW_F32 ? launch_tc_n<float>(a, bn, passes, st) : launch_tc_n<bf16>(a, bn, passes, st);
}

long long pack_weight_tc_bytes(int Cout, int Cin, int KH, int KW, int ipad, int transpose_flip, int passes) {
    const int Od = transpose_flip ? Cin : Cout, Id = transpose_flip ? Cout : Cin;
    if (ipad % 8 != 0 || ipad < Id || (passes != 1 && passes != 3)) return -1;
    const int bn = conv_tc_block_n(Od);
    const long long ntiles = (Od + bn - 1) / bn;
    const long long KB = ((long long)KH * KW * ipad + BK - 1) / BK;
    return ntiles * KB * (passes == 3 ? 2 : 1) * bn * BK * 2;
}

int pack_weight_tc(const float* w, void* out, int Cout, int Cin, int KH, int KW, int ipad, int transpose_flip, int passes, int fmt,
                   cudaStream_t st) {
    if (pack_weight_tc_bytes(Cout, Cin, KH, KW, ipad, transpose_flip, passes) <= 0) {
        affgw_set_error("pack_weight_tc: bad configuration (i_pad %d, passes %d)", ipad, passes);
        return -1;
    }
    const int Od = transpose_flip ? Cin : Cout;
    const int bn = conv_tc_block_n(Od);
    const int ntiles = (Od + bn - 1) / bn;
    const int KB = (KH * KW * ipad + BK - 1) / BK;
    const long long total = (long long)ntiles * KB * bn * BK;
    const int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    pack_weight_tc_kernel<<<blocks, 256, 0, st>>>(w, (bf16*)out, Cout, Cin, KH, KW, ipad, transpose_flip, bn, ntiles, KB,
                                                  passes == 3 ? 2 : 1, fmt);
    AFFGW_LAUNCH_CHECK("pack_weight_tc");
    return 0;
}

int split_planes(const void* x, int x_dt, void* planes, long long rows, int C, int pitch, int c_store, int passes, int pre_act,
                 int fmt, const float* scale_dev, cudaStream_t st) {
    if (c_store % 8 != 0 || c_store < C || (passes != 1 && passes != 3)) {
        affgw_set_error("split_planes: bad configuration (C %d, c_store %d, passes %d)", C, c_store, passes);
        return -1;
    }
    const long long total = rows * (c_store / 8);
    const int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
    const int npl = passes == 3 ? 2 : 1;
    if (x_dt == AFFGW_F32)
        split_planes_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, (bf16*)planes, rows, C, pitch, c_store, npl, pre_act, fmt, scale_dev);
    else
        split_planes_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)x, (bf16*)planes, rows, C, pitch, c_store, npl, pre_act, fmt, scale_dev);
    AFFGW_LAUNCH_CHECK("split_planes");
    return 0;
}
