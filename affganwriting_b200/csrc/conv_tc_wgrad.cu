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
#include "pos_frame.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int BP = 64;                       // pixels per stage
constexpr int IMG_BYTES = BP * 128;          // one [64 pixels][64 channels] bf16 image = 8 KB
constexpr int NUM_PRODUCER_THREADS = 128;
constexpr int NUM_THREADS = 160;

constexpr int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

template <int BN, int NPASS> struct WgCfg {
    static constexpr int NPL = NPASS == 3 ? 2 : 1;
    static constexpr int BNI = BN > 64 ? BN / 64 : 1;      // 64-channel dY images per plane
    static constexpr int A_PLANE = 2 * IMG_BYTES;
    static constexpr int B_PLANE = BNI * IMG_BYTES;
    static constexpr int A_BYTES = NPL * A_PLANE;
    static constexpr int B_BYTES = NPL * B_PLANE;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BUDGET = NPASS == 3 ? 200 * 1024 : 100 * 1024;
    static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 5 ? 5 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
    static constexpr int TMEM_COLS = pow2_cols(BN);
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
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int n, int fmt = 0) {       // fmt 0: bf16 operands, 1: fp16 operands
    return (1u << 4) | (fmt ? 0u : ((1u << 7) | (1u << 10))) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
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

struct WgArgs {
    const bf16* x;      // operand planes [NPL][N*H*W][Cin]   (g.Cin = stored channels)
    long long x_plane;
    const bf16* dy;     // operand planes [NPL][M][g.out_pitch]
    long long dy_plane;
    float* ws;          // [Cout][taps * Cin] fp32 partial sums
    ConvGeom g;
    long long m_per_split;
    FastDiv div_wo, div_ho;
    int fmt;                    // 0 = bf16 planes, 1 = fp16 planes
    const float* alpha_dev;     // 1 / (scale of the fp16 dY planes), device memory; nullptr = 1
};

template <int BN, int NPASS>
__global__ void __launch_bounds__(NUM_THREADS)
conv_wgrad_tcgen05_kernel(const __grid_constant__ WgArgs a) {
    using Cfg = WgCfg<BN, NPASS>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int NPL = Cfg::NPL;
    constexpr int BNI = Cfg::BNI;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 1);
    auto a_smem = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
    auto b_smem = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };

    const ConvGeom& g = a.g;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int ktot = g.KH * g.KW * g.Cin;               // flattened (tap, stored channel) extent = rows of dW^T
    const int n0 = blockIdx.y * BN;
    const long long mbeg = (long long)blockIdx.z * a.m_per_split;
    const long long mend = min(g.M, mbeg + a.m_per_split);
    const int nstages = (int)((mend - mbeg + BP - 1) / BP);

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
        tmem_alloc(tmem_ptr_addr, Cfg::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

    if (warp < 4) {
        // ============================== producer: gather(x) and dY images ==============================
        const int chunk = tid & 7, slot = tid >> 3;       // rows slot, slot+16, slot+32, slot+48
        // this thread's 8 consecutive rows of dW^T inside each of the CTA's two 64-row blocks: one tap, 8 channels
        int dyo[2], dxo[2], coff[2];
        bool kvalid[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int kf = (2 * blockIdx.x + i) * 64 + chunk * 8;
            kvalid[i] = kf < ktot;
            const int tap = kvalid[i] ? kf / g.Cin : 0;
            coff[i] = kvalid[i] ? kf - tap * g.Cin : 0;
            dyo[i] = tap / g.KW - g.pad;
            dxo[i] = tap % g.KW - g.pad;
        }
        bool nvalid[BNI];
#pragma unroll
        for (int i = 0; i < BNI; ++i) nvalid[i] = (chunk * 8 < BN) && (n0 + i * 64 + chunk * 8 < g.out_pitch);
        // 32-bit pixel / element arithmetic throughout (conv_wgrad_tc_ok bounds the tensors below 2^31 elements)
        uint32_t mrow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) mrow[j] = (uint32_t)mbeg + slot + 16 * j;
        const uint32_t m_end = (uint32_t)mend, m_last = (uint32_t)(g.M - 1);
        int ncol[BNI];
#pragma unroll
        for (int i = 0; i < BNI; ++i) ncol[i] = n0 + i * 64 + chunk * 8;
        const int HW = g.H * g.W;
        // signal stage i - LAG after issuing stage i.  LAG = STAGES - 1 would keep every stage either in flight or waiting for its
        // MMAs, so the next issue had to wait for the MMAs of the stage just signalled (issue and MMA alternated: 3260 clocks
        // per 768-clock k-block on the 2 x 7 maps); one stage of slack lets them overlap
        constexpr int LAG = STAGES > 2 ? STAGES - 2 : 1;
        for (int st = 0; st < nstages; ++st) {
            const int s = st % STAGES;
            const uint32_t ph = (uint32_t)(st / STAGES) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t a_base = a_smem(s), b_base = b_smem(s);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = slot + 16 * j;
                const bool mok = mrow[j] < m_end;
                const uint32_t soff = r * 128 + ((chunk ^ (r & 7)) << 4);
                // pixel index -> (image, oy, ox): two multiply-shift divisions (the maps are as small as 2 x 7, so stepping
                // the coordinates by 64 pixels with carries costs far more than recomputing them)
                const uint32_t mm = min(mrow[j], m_last);
                const uint32_t t = fast_div(mm, a.div_wo);
                const int ox = (int)(mm - t * (uint32_t)g.Wo);
                const uint32_t nimg = fast_div(t, a.div_ho);
                const int oy = (int)(t - nimg * (uint32_t)g.Ho);
                const int ybase = oy * g.stride, xbase = ox * g.stride_w;
                const int pix0 = (int)nimg * HW;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int sy = map_fast(ybase + dyo[i], g.Hv, g.pad_mode, g.up);
                    const int sx = map_fast(xbase + dxo[i], g.Wv, g.pad_mode, g.up);
                    const bool ok = mok && kvalid[i] && sy >= 0 && sx >= 0;
                    const bf16* src = a.x + (ok ? (pix0 + sy * g.W + sx) * g.Cin + coff[i] : 0);
                    cp_async_16(a_base + i * IMG_BYTES + soff, src, ok ? 16u : 0u);
                    if (NPL == 2) cp_async_16(a_base + Cfg::A_PLANE + i * IMG_BYTES + soff, src + a.x_plane, ok ? 16u : 0u);
                }
#pragma unroll
                for (int i = 0; i < BNI; ++i) {
                    const bool ok = mok && nvalid[i];
                    const bf16* src = a.dy + (ok ? (int)mm * g.out_pitch + ncol[i] : 0);
                    cp_async_16(b_base + i * IMG_BYTES + soff, src, ok ? 16u : 0u);
                    if (NPL == 2) cp_async_16(b_base + Cfg::B_PLANE + i * IMG_BYTES + soff, src + a.dy_plane, ok ? 16u : 0u);
                }
                mrow[j] += BP;          // next stage: 64 pixels on
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

        // ============================== epilogue: coalesced fp32 reductions into ws[co][tap * Cin + ci] ==============================
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
        const float alpha = a.alpha_dev ? __ldg(a.alpha_dev) : 1.f;
        const int kf = (2 * blockIdx.x) * 64 + row;
        const bool rok = kf < ktot;
        float* wrow = a.ws + kf;
#pragma unroll 1
        for (int jb = 0; jb < BN / 16; ++jb) {
            const int nb = n0 + jb * 16;
            if (nb >= g.Cout) break;
            uint32_t raw[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(jb * 16), raw);
            if (rok) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (nb + i < g.Cout) atomicAdd(wrow + (size_t)(nb + i) * ktot, __uint_as_float(raw[i]) * alpha);
            }
        }
        tc_fence_before();
    } else {
        // ============================== MMA issuer ==============================
        // every lane of warp 4 runs the loops (warp-uniform), one elected lane issues: see umma_bf16_elect
        const uint32_t idesc = make_idesc_bf16_mn(BN, a.fmt);
        constexpr uint32_t lbo = (uint32_t)IMG_BYTES, sbo = 1024u;
        for (int st = 0; st < nstages; ++st) {
            const int s = st % STAGES;
            const uint32_t ph = (uint32_t)(st / STAGES) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < BP / 16; ++k) {      // 16 pixels = two 8-row swizzle atoms = 2048 bytes per K step
                const uint64_t a_hi = make_mnmajor_sw128_desc(a_smem(s) + k * 2048, lbo, sbo);
                const uint64_t b_hi = make_mnmajor_sw128_desc(b_smem(s) + k * 2048, lbo, sbo);
                umma_bf16_elect(tmem_base, a_hi, b_hi, idesc, (uint32_t)((st | k) != 0));
                if (NPASS == 3) {
                    const uint64_t a_lo = make_mnmajor_sw128_desc(a_smem(s) + Cfg::A_PLANE + k * 2048, lbo, sbo);
                    const uint64_t b_lo = make_mnmajor_sw128_desc(b_smem(s) + Cfg::B_PLANE + k * 2048, lbo, sbo);
                    umma_bf16_elect(tmem_base, a_lo, b_hi, idesc, 1u);
                    umma_bf16_elect(tmem_base, a_hi, b_lo, idesc, 1u);
                }
            }
            umma_commit_elect(empty_bar(s));
        }
        umma_commit_elect(tmem_full_bar);
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// dw_oihw[co][ci][tap] += ws[co][tap * Cx + ci]   (ci < Cin_w: channel padding of the stored input is dropped)
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

template <int BN, int NPASS>
int launch_wg(WgArgs& a, cudaStream_t st) {
    using Cfg = WgCfg<BN, NPASS>;
    static bool configured = false;
    auto kern = conv_wgrad_tcgen05_kernel<BN, NPASS>;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) {
            affgw_set_error("conv_wgrad_tc: cannot reserve %d bytes of shared memory", Cfg::SMEM_BYTES);
            return -2;
        }
        configured = true;
    }
    const ConvGeom& g = a.g;
    const int ktot = g.KH * g.KW * g.Cin;
    const int gx = (ktot + 127) / 128, gy = (g.Cout + BN - 1) / BN;
    // split of the pixel range over CTAs.  One round of CTAs costs a fixed part (launch, pipeline fill, and above all the fp32
    // reductions of its 128 x BN tile into the workspace, which contend with the other splits of the same tile) plus its
    // stages; measured on B200: ~20 us fixed, ~1.2 us per 64-pixel stage.  Pick the split count with the lowest modelled time.
    const long long per_wave = 148LL * (NPASS == 3 ? 1 : 2);
    const long long tiles = (long long)gx * gy;
    const long long total_stages = (g.M + BP - 1) / BP;
    long long max_splits = (g.M + 511) / 512;
    if (max_splits > 65535) max_splits = 65535;
    if (max_splits < 1) max_splits = 1;
    long long splits = 1, best = -1;
    for (long long sp = 1; sp <= max_splits; ++sp) {
        const long long rounds = (tiles * sp + per_wave - 1) / per_wave;
        const long long cost = rounds * (100 + 6 * ((total_stages + sp - 1) / sp));
        if (best < 0 || cost < best) { best = cost; splits = sp; }
        if (tiles * sp >= 4 * per_wave) break;
    }
    long long mps = (g.M + splits - 1) / splits;
    mps = (mps + BP - 1) / BP * BP;
    splits = (g.M + mps - 1) / mps;
    a.m_per_split = mps;
    dim3 grid(gx, gy, (unsigned)splits);
    kern<<<grid, NUM_THREADS, Cfg::SMEM_BYTES, st>>>(a);
    AFFGW_LAUNCH_CHECK("conv_wgrad_tcgen05");
    return 0;
}

template <int BN>
int launch_wg_p(WgArgs& a, int passes, cudaStream_t st) {
    return passes == 3 ? launch_wg<BN, 3>(a, st) : launch_wg<BN, 1>(a, st);
}

}  // namespace

int conv_tc_block_n(int cout);

// g.Cin must be the STORED channel count of the x planes, g.out_pitch that of the dY planes (multiples of 8)
int conv_wgrad_tc_ok(const ConvGeom& g) {
    if (g.Cin % 8 != 0 || g.in_pitch != g.Cin || g.out_pitch % 8 != 0 || g.out_pitch < g.Cout) return 0;
    if (g.pre_act != ACT_NONE || g.zi != 1 || g.zi_w != 1) return 0;
    if ((long long)g.N * g.H * g.W * g.Cin >= (1LL << 31) || g.M * g.out_pitch >= (1LL << 31)) return 0;   // 32-bit offsets
    return conv_tc_block_n(g.Cout);
}

long long conv_wgrad_tc_ws_bytes(const ConvGeom& g) { return (long long)g.Cout * g.KH * g.KW * g.Cin * 4; }

int conv_wgrad_tc(const void* x_planes, long long x_plane, const void* dy_planes, long long dy_plane, float* dw, void* workspace,
                  const ConvGeom& g, int cin_w, int passes, int fmt, const float* alpha_dev, cudaStream_t st) {
    const int bn = conv_wgrad_tc_ok(g);
    if (!bn || !workspace || (passes != 1 && passes != 3)) {
        affgw_set_error("conv_wgrad_tc: unsupported shape or missing workspace");
        return -1;
    }
    float* ws = (float*)workspace;
    if (cudaMemsetAsync(ws, 0, (size_t)conv_wgrad_tc_ws_bytes(g), st) != cudaSuccess) {
        affgw_set_error("conv_wgrad_tc: memset failed");
        return -2;
    }
    WgArgs a;
    a.x = (const bf16*)x_planes; a.x_plane = x_plane;
    a.dy = (const bf16*)dy_planes; a.dy_plane = dy_plane;
    a.ws = ws; a.g = g; a.m_per_split = 0;
    a.fmt = fmt; a.alpha_dev = alpha_dev;
    a.div_wo = make_fastdiv(g.Wo);
    a.div_ho = make_fastdiv(g.Ho);
    int rc;
    switch (bn) {
        case 16: rc = launch_wg_p<16>(a, passes, st); break;
        case 32: rc = launch_wg_p<32>(a, passes, st); break;
        case 64: rc = launch_wg_p<64>(a, passes, st); break;
        default: rc = launch_wg_p<128>(a, passes, st); break;
    }
    if (rc) return rc;
    const int taps = g.KH * g.KW;
    const long long total = (long long)g.Cout * cin_w * taps;
    const int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    wgrad_unpack_kernel<<<blocks, 256, 0, st>>>(ws, dw, g.Cout, cin_w, g.Cin, taps);
    AFFGW_LAUNCH_CHECK("wgrad_unpack");
    return 0;
}

// conv_shift.cu
int conv_wgrad_pos_tc(const void* x_planes, const PosFrame& fx, const void* dy_planes, const PosFrame& fy, float* ws, int K,
                      int Cout, int cs, int Ho, int Wo, int passes, int fmt, const float* alpha_dev, cudaStream_t st);

// position-space weight gradient (conv_shift.cu) + the same workspace reduction / unpack as above;
// g.Cin = stored channel count of the ws rows, cin_w = channels of the parameter
int conv_wgrad_pos(const void* x_planes, const PosFrame& fx, const void* dy_planes, const PosFrame& fy, float* dw, void* workspace,
                   const ConvGeom& g, int cin_w, int passes, int fmt, const float* alpha_dev, cudaStream_t st) {
    float* ws = (float*)workspace;
    if (cudaMemsetAsync(ws, 0, (size_t)conv_wgrad_tc_ws_bytes(g), st) != cudaSuccess) {
        affgw_set_error("conv_wgrad_pos: memset failed");
        return -2;
    }
    if (int rc = conv_wgrad_pos_tc(x_planes, fx, dy_planes, fy, ws, g.KH, g.Cout, g.Cin, g.Ho, g.Wo, passes, fmt, alpha_dev, st)) return rc;
    const int taps = g.KH * g.KW;
    const long long total = (long long)g.Cout * cin_w * taps;
    const int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    wgrad_unpack_kernel<<<blocks, 256, 0, st>>>(ws, dw, g.Cout, cin_w, g.Cin, taps);
    AFFGW_LAUNCH_CHECK("wgrad_unpack");
    return 0;
}
