// Weight gradient of the convolutions on the tcgen05 tensor cores.
//
//   dW^T[(tap, ci)][co] = sum over output pixels p of  gather(x)[p][(tap, ci)] * dY[p][co]
//
// i.e. a GEMM whose reduction dimension is the pixel index.  Both operands are stored pixel-major with channels
// contiguous (NHWC), which is exactly UMMA's "MN-major" operand form: a 128-byte shared-memory row holds 64 contiguous
// channels (M or N) of ONE pixel (K), eight pixels form one 1024-byte swizzle atom.  So the same 16-byte cp.async gather
// and the same software 128B swizzle as the forward kernel build both operand images; only the descriptors differ
// (MN-major bit set for A and B, SBO = 1024 B between 8-pixel groups, LBO = 8192 B between 64-channel blocks).
//
//   M tile  = 128 rows = two 64-channel blocks of the flattened (tap, cin/64) axis (for Cin = 64 that is two taps)
//   N tile  = 64 or 128 output channels
//   K       = 64 pixels per pipeline stage, a split of the pixel range per CTA (grid.z), fp32 TMEM accumulator
// The epilogue reduces the split partials with coalesced fp32 atomics into an [Cout][taps][Cin] workspace which a small
// kernel then adds into the OIHW parameter gradient.
// Replaces the weight-gradient half of cuDNN's convolution backward behind the reference's loss.backward()
// (network_tro.py:55,102,113,129).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int BP = 64;                       // pixels per stage
constexpr int IMG_BYTES = BP * 128;          // one [64 pixels][64 channels] bf16 image = 8 KB
constexpr int NUM_PRODUCER_THREADS = 128;
constexpr int NUM_THREADS = 160;

template <int BN> struct WgCfg {
    static constexpr int STAGES = (BN == 64) ? 4 : 3;
    static constexpr int A_BYTES = 2 * IMG_BYTES;
    static constexpr int B_BYTES = (BN / 64) * IMG_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// MN-major, SWIZZLE_128B descriptor: start, LBO (between 64-element MN blocks), SBO (between 8-row K groups)
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}
// kind::f16, D = f32, A = B = bf16, both MN-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ int map_fast(int v, int V, int pad_mode, int up) {
    if (pad_mode == PAD_ZERO) {
        if ((unsigned)v >= (unsigned)V) return -1;
    } else if (pad_mode == PAD_REFLECT) {
        v = v < 0 ? -v : v;
        v = v >= V ? 2 * (V - 1) - v : v;
    } else {
        v = max(0, min(v, V - 1));
    }
    return up == 2 ? (v >> 1) : v;
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS)
conv_wgrad_tcgen05_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ ws, const ConvGeom g,
                          const long long m_per_split, const int swap_lbo_sbo) {
    using Cfg = WgCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 1);
    auto a_smem = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
    auto b_smem = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cblocks = g.Cin / 64;                    // g.Cin is the stored (64-aligned) channel count
    const int taps = g.KH * g.KW;
    const int total64 = taps * cblocks;
    const int n0 = blockIdx.y * BN;
    const long long mbeg = (long long)blockIdx.z * m_per_split;
    const long long mend = min(g.M, mbeg + m_per_split);
    const int nstages = (int)((mend - mbeg + BP - 1) / BP);

    // the two 64-channel blocks of this CTA's M tile
    int tap_i[2], cb_i[2];
    bool valid_i[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int kb = 2 * blockIdx.x + i;
        valid_i[i] = kb < total64;
        tap_i[i] = valid_i[i] ? kb / cblocks : 0;
        cb_i[i] = valid_i[i] ? kb % cblocks : 0;
    }

    if (tid == NUM_PRODUCER_THREADS) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), NUM_PRODUCER_THREADS);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        __syncwarp();
        tmem_alloc(tmem_ptr_addr, BN);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

    if (warp < 4) {
        // ============================== producer: gather(x) and dY images ==============================
        const int chunk = tid & 7, slot = tid >> 3;       // rows slot, slot+16, slot+32, slot+48
        int oy[4], ox[4], nimg[4];
        long long mrow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            mrow[j] = mbeg + slot + 16 * j;
            const long long mm = min(mrow[j], g.M - 1);
            ox[j] = (int)(mm % g.Wo);
            const long long t = mm / g.Wo;
            oy[j] = (int)(t % g.Ho);
            nimg[j] = (int)(t / g.Ho);
        }
        int dyo[2], dxo[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            dyo[i] = tap_i[i] / g.KW - g.pad;
            dxo[i] = tap_i[i] % g.KW - g.pad;
        }
        constexpr int LAG = STAGES - 1;
        for (int st = 0; st < nstages; ++st) {
            const int s = st % STAGES;
            const uint32_t ph = (uint32_t)(st / STAGES) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t a_base = a_smem(s), b_base = b_smem(s);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = slot + 16 * j;
                const bool mok = mrow[j] < mend;
                const uint32_t soff = r * 128 + ((chunk ^ (r & 7)) << 4);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    bool ok = mok && valid_i[i];
                    int sy = 0, sx = 0;
                    if (ok) {
                        sy = map_fast(oy[j] * g.stride + dyo[i], g.Hv, g.pad_mode, g.up);
                        sx = map_fast(ox[j] * g.stride + dxo[i], g.Wv, g.pad_mode, g.up);
                        ok = sy >= 0 && sx >= 0;
                    }
                    const bf16* src = ok ? x + ((size_t)((size_t)nimg[j] * g.H * g.W + (size_t)sy * g.W + sx) * g.in_pitch +
                                                cb_i[i] * 64 + chunk * 8)
                                         : x;
                    cp_async_16(a_base + i * IMG_BYTES + soff, src, ok ? 16u : 0u);
                }
#pragma unroll
                for (int i = 0; i < BN / 64; ++i) {
                    const bf16* src = mok ? dy + (size_t)mrow[j] * g.out_pitch + n0 + i * 64 + chunk * 8 : dy;
                    cp_async_16(b_base + i * IMG_BYTES + soff, src, mok ? 16u : 0u);
                }
                // advance this row slot by one stage (64 pixels)
                mrow[j] += BP;
                ox[j] += BP;
                while (ox[j] >= g.Wo) {
                    ox[j] -= g.Wo;
                    if (++oy[j] == g.Ho) { oy[j] = 0; ++nimg[j]; }
                }
            }
            cp_async_commit();
            if (st >= LAG) {
                cp_async_wait<LAG>();
                fence_proxy_async();
                mbar_arrive(full_bar((st - LAG) % STAGES));
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        for (int st = (nstages > LAG ? nstages - LAG : 0); st < nstages; ++st) mbar_arrive(full_bar(st % STAGES));

        // ============================== epilogue: coalesced fp32 reductions into ws[co][tap][ci] ==============================
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
        const int img = row >> 6, ch = row & 63;
        const bool rok = valid_i[img];
        float* wrow = ws + ((size_t)tap_i[img] * g.Cin + cb_i[img] * 64 + ch);
        const size_t co_stride = (size_t)taps * g.Cin;
#pragma unroll 1
        for (int jb = 0; jb < BN / 32; ++jb) {
            uint32_t raw[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(jb * 32), raw);
            if (rok) {
#pragma unroll
                for (int i = 0; i < 32; ++i) atomicAdd(wrow + (size_t)(n0 + jb * 32 + i) * co_stride, __uint_as_float(raw[i]));
            }
        }
        tc_fence_before();
    } else if (lane == 0) {
        // ============================== MMA issuer ==============================
        constexpr uint32_t idesc = make_idesc_bf16_mn(BN);
        const uint32_t lbo = swap_lbo_sbo ? 1024u : (uint32_t)IMG_BYTES;
        const uint32_t sbo = swap_lbo_sbo ? (uint32_t)IMG_BYTES : 1024u;
        for (int st = 0; st < nstages; ++st) {
            const int s = st % STAGES;
            const uint32_t ph = (uint32_t)(st / STAGES) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < BP / 16; ++k) {      // 16 pixels = two 8-row swizzle atoms = 2048 bytes per K step
                const uint64_t adesc = make_mnmajor_sw128_desc(a_smem(s) + k * 2048, lbo, sbo);
                const uint64_t bdesc = make_mnmajor_sw128_desc(b_smem(s) + k * 2048, lbo, sbo);
                umma_bf16(tmem_base, adesc, bdesc, idesc, (uint32_t)((st | k) != 0));
            }
            umma_commit(empty_bar(s));
        }
        umma_commit(tmem_full_bar);
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

// dw_oihw[co][ci][tap] += ws[co][tap][ci]   (ci < Cin_w: channel padding of the stored input is dropped)
__global__ void wgrad_unpack_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Cout, int Cin_w, int Cx, int taps) {
    const long long total = (long long)Cout * Cin_w * taps;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int tap = (int)(idx % taps);
        const long long t = idx / taps;
        const int ci = (int)(t % Cin_w), co = (int)(t / Cin_w);
        dw[idx] += ws[((size_t)co * taps + tap) * Cx + ci];
    }
}

template <int BN>
int launch_wg(const void* x, const void* dy, float* ws, const ConvGeom& g, int swap, cudaStream_t st) {
    using Cfg = WgCfg<BN>;
    static bool configured = false;
    auto kern = conv_wgrad_tcgen05_kernel<BN>;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) {
            affgw_set_error("conv_wgrad_tc: cannot reserve %d bytes of shared memory", Cfg::SMEM_BYTES);
            return -2;
        }
        configured = true;
    }
    const int taps = g.KH * g.KW, total64 = taps * (g.Cin / 64);
    const int gx = (total64 + 1) / 2, gy = g.Cout / BN;
    long long splits = (2LL * 148 + (long long)gx * gy - 1) / ((long long)gx * gy);
    const long long max_splits = (g.M + 511) / 512;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    long long mps = (g.M + splits - 1) / splits;
    mps = (mps + BP - 1) / BP * BP;
    splits = (g.M + mps - 1) / mps;
    dim3 grid(gx, gy, (unsigned)splits);
    kern<<<grid, NUM_THREADS, Cfg::SMEM_BYTES, st>>>((const bf16*)x, (const bf16*)dy, ws, g, mps, swap);
    AFFGW_LAUNCH_CHECK("conv_wgrad_tcgen05");
    return 0;
}

}  // namespace

// g.Cin must be the STORED channel count of x (multiple of 64); cin_w the parameter's input channels (<= g.Cin)
int conv_wgrad_tc_ok(const ConvGeom& g, int x_dt, int dy_dt) {
    if (x_dt != AFFGW_BF16 || dy_dt != AFFGW_BF16) return 0;
    if (g.Cin % 64 != 0 || g.in_pitch != g.Cin || g.Cout % 64 != 0 || g.out_pitch != g.Cout) return 0;
    if (g.pre_act != ACT_NONE || g.zi != 1) return 0;
    if ((long long)g.N * g.H * g.W >= (1LL << 31)) return 0;
    return (g.Cout % 128 == 0) ? 128 : 64;
}

long long conv_wgrad_tc_ws_bytes(const ConvGeom& g) { return (long long)g.Cout * g.KH * g.KW * g.Cin * 4; }

int conv_wgrad_tc(const void* x, const void* dy, float* dw, void* workspace, const ConvGeom& g, int cin_w, cudaStream_t st) {
    const int bn = conv_wgrad_tc_ok(g, AFFGW_BF16, AFFGW_BF16);
    if (!bn || !workspace) {
        affgw_set_error("conv_wgrad_tc: unsupported shape or missing workspace");
        return -1;
    }
    static int swap = -1;
    if (swap < 0) {
        const char* e = getenv("AFFGW_WGRAD_SWAP_LBO_SBO");
        swap = (e && e[0] == '1') ? 1 : 0;
    }
    float* ws = (float*)workspace;
    if (cudaMemsetAsync(ws, 0, (size_t)conv_wgrad_tc_ws_bytes(g), st) != cudaSuccess) {
        affgw_set_error("conv_wgrad_tc: memset failed");
        return -2;
    }
    const int rc = bn == 128 ? launch_wg<128>(x, dy, ws, g, swap, st) : launch_wg<64>(x, dy, ws, g, swap, st);
    if (rc) return rc;
    const int taps = g.KH * g.KW;
    const long long total = (long long)g.Cout * cin_w * taps;
    const int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    wgrad_unpack_kernel<<<blocks, 256, 0, st>>>(ws, dw, g.Cout, cin_w, g.Cin, taps);
    AFFGW_LAUNCH_CHECK("wgrad_unpack");
    return 0;
}
