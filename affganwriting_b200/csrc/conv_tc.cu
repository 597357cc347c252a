// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, fp32 accumulators in TMEM).
//
//   D[128 pixels][BN couts] += A[128 pixels][64 cin of one filter tap] * B[BN couts][64 cin]^T       (bf16 x bf16 -> fp32)
//
// A (activations, NHWC bf16) is gathered straight from global memory into the 128B-swizzled K-major shared-memory image
// UMMA expects: four producer warps issue 16-byte cp.async copies whose source address already contains the
// convolution's index map (zero / reflect / replicate padding, nearest x2 upsampling, zero insertion for strided
// dgrad) - the padded or upsampled tensor is never materialised.  B (weights) is pre-packed once per optimiser step
// into exactly that shared-memory image (affgw_pack_weight_tc), so one cp.async.bulk (TMA engine, mbarrier
// complete_tx) moves a whole [BN][64] tile.  One elected thread issues tcgen05.mma; tcgen05.commit releases the
// smem stage and finally publishes the accumulator; the producer warps then turn into the epilogue
// (tcgen05.ld -> +bias -> +addend -> activation -> 128-bit stores).
//
// Replaces the cuDNN convolutions behind reference blocks.py:148 / vgg_tro_channel3_modi.py:47 /
// modules_tro.py:252-259 for every bf16 layer with Cin % 64 == 0 and Cout % 64 == 0; the same kernel computes dgrad.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int BM = 128, BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int NUM_PRODUCER_THREADS = 128;
constexpr int NUM_THREADS = 160;            // 4 producer/epilogue warps + 1 MMA warp

template <int BN> struct TcCfg {
    static constexpr int STAGES = (BN == 64) ? 4 : 3;
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
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
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <typename TO> __device__ __forceinline__ void store_row32(TO* dst, const float (&v)[32]);
template <> __device__ __forceinline__ void store_row32<float>(float* dst, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
template <> __device__ __forceinline__ void store_row32<bf16>(bf16* dst, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
        reinterpret_cast<uint4*>(dst)[i] = u;
    }
}
template <typename TO> __device__ __forceinline__ void load_row32(const TO* src, float (&v)[32]);
template <> __device__ __forceinline__ void load_row32<float>(const float* src, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 f = reinterpret_cast<const float4*>(src)[i];
        v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
    }
}
template <> __device__ __forceinline__ void load_row32<bf16>(const bf16* src, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float t[8];
        ld8(src + 8 * i, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * i + j] = t[j];
    }
}

// ---------------------------------------------------------------------------------------------------- the kernel
template <int BN, typename TO>
__global__ void __launch_bounds__(NUM_THREADS)
conv_igemm_tcgen05_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w_tiles, const float* __restrict__ bias,
                          const TO* __restrict__ addend, TO* __restrict__ y, const ConvGeom g) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;   // full[STAGES], empty[STAGES], tmem_full, tmem_ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 1);
    auto a_smem = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
    auto b_smem = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + A_STAGE_BYTES; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int cblocks = g.Cin / BK;
    const int KB = g.KH * g.KW * cblocks;

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
        // ============================== A/B producer ==============================
        const int chunk = tid & 7;        // 16-byte chunk of the 128-byte row
        const int rg = tid >> 3;          // rows rg, rg+16, ..., rg+112
        int vy0[8], vx0[8], nbase[8];
        bool mvalid[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long m = m0 + rg + 16 * i;
            mvalid[i] = m < g.M;
            int ox = 0, oy = 0, n = 0;
            if (mvalid[i]) {
                ox = (int)(m % g.Wo);
                const long long t = m / g.Wo;
                oy = (int)(t % g.Ho);
                n = (int)(t / g.Ho);
            }
            vy0[i] = oy * g.stride - g.pad;
            vx0[i] = ox * g.stride - g.pad;
            nbase[i] = n * g.H * g.W;
        }
        const bf16* wt = w_tiles + (size_t)blockIdx.y * KB * (BN * BK);
        int ky = 0, kx = 0, cb = 0;
        int sy[8], sx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            sy[i] = map_coord(vy0[i], g.Hv, g.pad_mode, g.up, g.zi);
            sx[i] = map_coord(vx0[i], g.Wv, g.pad_mode, g.up, g.zi);
        }
        constexpr int LAG = STAGES - 1;
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            if (tid == 0) {
                mbar_expect_tx(full_bar(s), Cfg::B_STAGE_BYTES);
                bulk_copy_g2s(b_smem(s), wt + (size_t)kb * (BN * BK), Cfg::B_STAGE_BYTES, full_bar(s));
            }
            const uint32_t a_base = a_smem(s);
            const int coff = cb * BK + chunk * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rg + 16 * i;
                const bool ok = mvalid[i] && sy[i] >= 0 && sx[i] >= 0;
                const bf16* src = ok ? x + ((size_t)(nbase[i] + sy[i] * g.W + sx[i]) * g.in_pitch + coff) : x;
                cp_async_16(a_base + r * 128 + ((chunk ^ (r & 7)) << 4), src, ok ? 16u : 0u);
            }
            cp_async_commit();
            if (kb >= LAG) {
                cp_async_wait<LAG>();
                fence_proxy_async();
                mbar_arrive(full_bar((kb - LAG) % STAGES));
            }
            // advance (cb, kx, ky) and refresh the source coordinates that changed
            if (++cb == cblocks) {
                cb = 0;
                if (++kx == g.KW) {
                    kx = 0;
                    ++ky;
#pragma unroll
                    for (int i = 0; i < 8; ++i) sy[i] = map_coord(vy0[i] + ky, g.Hv, g.pad_mode, g.up, g.zi);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) sx[i] = map_coord(vx0[i] + kx, g.Wv, g.pad_mode, g.up, g.zi);
            }
        }
        // drain the last LAG stages
        cp_async_wait<0>();
        fence_proxy_async();
        for (int kb = (KB > LAG ? KB - LAG : 0); kb < KB; ++kb) mbar_arrive(full_bar(kb % STAGES));

        // ============================== epilogue ==============================
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const long long m = m0 + warp * 32 + lane;
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
            uint32_t raw[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32), raw);
            if (m < g.M) {
                float v[32];
                const int nb = n0 + j * 32;
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) + (bias ? __ldg(bias + nb + i) : 0.f);
                if (addend) {
                    float a[32];
                    load_row32<TO>(addend + m * g.out_pitch + nb, a);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += a[i];
                }
                if (g.post_act != ACT_NONE) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = act_apply(v[i], g.post_act);
                }
                store_row32<TO>(y + m * g.out_pitch + nb, v);
            }
        }
        tc_fence_before();
    } else if (lane == 0) {
        // ============================== MMA issuer (one thread) ==============================
        constexpr uint32_t idesc = make_idesc_bf16(BN);
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint64_t adesc = make_kmajor_sw128_desc(a_smem(s));
            const uint64_t bdesc = make_kmajor_sw128_desc(b_smem(s));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)   // +32 bytes (encoded >>4) per 16-element K step inside the swizzle atom
                umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((kb | k) != 0));
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

// weights: logical [O'][taps][I'pad] -> per (n-tile, k-block) [BN][64] tiles with the 128B swizzle already applied
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int KH, int KW,
                                      int ipad, int transpose_flip, int BN, int ntiles) {
    const int Od = transpose_flip ? Cin : Cout, Id = transpose_flip ? Cout : Cin;
    const int taps = KH * KW, cblocks = ipad / BK, KB = taps * cblocks;
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
        const int tap = kb / cblocks, i = (kb % cblocks) * BK + e;
        const int ky = tap / KW, kx = tap % KW;
        float v = 0.f;
        if (o < Od && i < Id) {
            if (transpose_flip)
                v = w[(((long long)i * Cin + o) * KH + (KH - 1 - ky)) * KW + (KW - 1 - kx)];
            else
                v = w[(((long long)o * Cin + i) * KH + ky) * KW + kx];
        }
        out[idx] = __float2bfloat16_rn(v);
    }
}

template <int BN, typename TO>
int launch_tc(const void* x, const void* w_tiles, const float* bias, const void* addend, void* y, const ConvGeom& g,
              cudaStream_t st) {
    using Cfg = TcCfg<BN>;
    static bool configured = false;
    auto kern = conv_igemm_tcgen05_kernel<BN, TO>;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) {
            affgw_set_error("conv_fwd_tc: cannot reserve %d bytes of shared memory", Cfg::SMEM_BYTES);
            return -2;
        }
        configured = true;
    }
    dim3 grid((unsigned)((g.M + BM - 1) / BM), g.Cout / BN);
    kern<<<grid, NUM_THREADS, Cfg::SMEM_BYTES, st>>>((const bf16*)x, (const bf16*)w_tiles, bias, (const TO*)addend, (TO*)y, g);
    AFFGW_LAUNCH_CHECK("conv_igemm_tcgen05");
    return 0;
}

}  // namespace

int conv_tc_block_n(const ConvGeom& g, int x_dt, int w_dt) {
    if (x_dt != AFFGW_BF16 || w_dt != AFFGW_BF16) return 0;
    if (g.Cin % BK != 0 || g.Cout % 64 != 0) return 0;
    if (g.in_pitch % 8 != 0 || g.out_pitch % 8 != 0) return 0;
    if (g.pre_act != ACT_NONE) return 0;
    if ((long long)g.N * g.H * g.W >= (1LL << 31)) return 0;
    return (g.Cout % 128 == 0) ? 128 : 64;
}

int conv_fwd_tc(const void* x, const void* w_tiles, const float* bias, const void* addend, void* y, int y_dt,
                const ConvGeom& g, cudaStream_t st) {
    const int bn = conv_tc_block_n(g, AFFGW_BF16, AFFGW_BF16);
    if (bn == 128) {
        return y_dt == AFFGW_F32 ? launch_tc<128, float>(x, w_tiles, bias, addend, y, g, st)
                                 : launch_tc<128, bf16>(x, w_tiles, bias, addend, y, g, st);
    } else if (bn == 64) {
        return y_dt == AFFGW_F32 ? launch_tc<64, float>(x, w_tiles, bias, addend, y, g, st)
                                 : launch_tc<64, bf16>(x, w_tiles, bias, addend, y, g, st);
    }
    affgw_set_error("conv_fwd_tc: unsupported shape (Cin %d, Cout %d, pitches %d/%d)", g.Cin, g.Cout, g.in_pitch, g.out_pitch);
    return -1;
}

long long pack_weight_tc_bytes(int Cout, int Cin, int KH, int KW, int ipad, int transpose_flip, int block_n) {
    const int Od = transpose_flip ? Cin : Cout;
    if (block_n <= 0 || ipad % BK != 0) return -1;
    const long long ntiles = (Od + block_n - 1) / block_n;
    return ntiles * KH * KW * (ipad / BK) * block_n * BK * 2;
}

int pack_weight_tc(const float* w, void* out, int Cout, int Cin, int KH, int KW, int ipad, int transpose_flip, int block_n,
                   cudaStream_t st) {
    const int Od = transpose_flip ? Cin : Cout, Id = transpose_flip ? Cout : Cin;
    if ((block_n != 64 && block_n != 128) || ipad % BK != 0 || ipad < Id) {
        affgw_set_error("pack_weight_tc: bad tile configuration (block_n %d, i_pad %d)", block_n, ipad);
        return -1;
    }
    const int ntiles = (Od + block_n - 1) / block_n;
    const long long total = (long long)ntiles * KH * KW * (ipad / BK) * block_n * BK;
    const int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    pack_weight_tc_kernel<<<blocks, 256, 0, st>>>(w, (bf16*)out, Cout, Cin, KH, KW, ipad, transpose_flip, block_n, ntiles);
    AFFGW_LAUNCH_CHECK("pack_weight_tc");
    return 0;
}
