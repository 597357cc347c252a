// Stride-1 convolution as a "shifted" implicit GEMM on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulators).
//
// The im2col kernel in conv_tc.cu re-reads every input element once per filter tap from L2 (9x for 3x3, 25x for 5x5,
// 49x for 7x7); at ~42 B/clk/SM of L2->SM bandwidth that, not the tensor pipe, bounded it.  Here a CTA stages a window of
// the PADDED (and, for the decoder, x2-upsampled) input in shared memory ONCE per 16-channel block and every filter tap
// is just a different start address of the same window:
//
//   virtual input V[n][yp][xp][c], yp < Hp = up*H + 2*pad, xp < Wp; flattened position q = (n*Hp + yp)*Wp + xp
//   output (n, y, x) lives at q = (n*Hp + y)*Wp + x and reads V[q + ky*Wp + kx]           (stride 1)
//
//   shared-memory window, "planar": [hi/lo plane][8-channel group][position] x 16 bytes.  Eight consecutive positions
//   of one channel group are 128 contiguous bytes = one UMMA core matrix of the un-swizzled K-major operand layout,
//   so A for tap (ky, kx) and M-tile mt is the descriptor {start = window + (mt*128 + ky*Wp + kx)*16,
//   LBO = plane pitch (next 8 channels), SBO = 128 (next 8 positions)}.  Padding (zero / reflect / replicate) and
//   nearest x2 upsampling are resolved once per position when the window is filled (16-byte cp.async with the mapped
//   source address or zero fill), never per tap.
//
// CTA tile: 4 M-tiles x 128 positions against BN <= 64 output channels, so a weight stage (taps x 16 channels x BN,
// one cp.async.bulk) is reused by 512 output positions; 2 x 4 x BN TMEM columns double-buffer the accumulators so the
// epilogue of one tile overlaps the main loop of the next.  Positions that fall on padding columns / rows compute junk
// that the epilogue skips (2/Wp of the work).  The 16/32-channel discriminator blocks, the 1->16 and 64->1 7x7 stems
// and the 512-channel VGG / decoder layers all run through this kernel; dgrad is the same kernel on flipped weights.
//
// NPASS = 3: split-bf16 product a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (see conv_tc.cu).
//
// Replaces the cuDNN convolutions behind reference blocks.py:148 / vgg_tro_channel3_modi.py:47 / modules_tro.py:594-603.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int NUM_PRODUCER_THREADS = 128;
constexpr int MMA_WARP = 4;
constexpr int NUM_THREADS = 32 * 9;          // 4 producer warps + 1 MMA warp + 4 epilogue warps
constexpr int MT = 4;                        // M-tiles (of 128 positions) per CTA tile
constexpr int TILE_POS = MT * 128;
constexpr int MAXP = 16;                     // window positions per producer thread pair slot: NP <= 64 * MAXP
constexpr int MAX_STAGES = 4;
constexpr int SMEM_LIMIT = 227 * 1024;

struct ShArgs {
    const bf16* x;             // operand planes [NPL][N*H*W][Cs]
    long long x_plane;         // elements between the hi and lo plane
    const bf16* w;             // [n_tile][ky][cb][kx][plane][cgroup][BN][8]
    const float* bias;
    const float* addend;
    float* y;
    int N, H, W, Cs;           // stored input (before upsampling); Cs = stored channels of the planes
    int up, pad, pad_mode, K;
    int Hv, Wv, Hp, Wp, Ho, Wo;
    int Cout, out_pitch, post_act;
    int CB;                    // 16-channel blocks
    int KYG;                   // kernel rows per pipeline stage: K (whole filter) or 1
    int NP, NPa;               // window positions per stage, plane pitch in positions
    int stages;
    int a_bytes, b_chunk_bytes;   // bytes of one stage's window / of one (ky, cb) weight chunk
    int n_tiles;
    int total_tiles;
    int Q;                     // N * Hp * Wp
    int vec_ok;
};

// K-major, un-swizzled shared-memory matrix descriptor (core matrices of 8 rows x 16 bytes)
__device__ __forceinline__ uint64_t make_kmajor_nosw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;     // between the two 8-element K core matrices
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;     // between 8-row groups along M / N
    d |= 1ull << 46;                                      // descriptor version (sm_100)
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_kk(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int BN, int NPASS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_shift_tcgen05_kernel(const __grid_constant__ ShArgs a) {
    constexpr int NPL = NPASS == 3 ? 2 : 1;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    const int stages = a.stages;
    const int b_bytes = a.KYG * a.b_chunk_bytes;
    const int stage_bytes = a.a_bytes + b_bytes;
    const uint32_t bar_base = smem_base + stages * stage_bytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
    auto tmem_full_bar = [&](int i) { return bar_base + 8u * (2 * MAX_STAGES + i); };
    auto tmem_empty_bar = [&](int i) { return bar_base + 8u * (2 * MAX_STAGES + 2 + i); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * MAX_STAGES + 4);
    auto a_smem = [&](int s) { return smem_base + s * stage_bytes; };
    auto b_smem = [&](int s) { return smem_base + s * stage_bytes + a.a_bytes; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KG = a.K / a.KYG;               // stages per channel block
    const uint32_t plane_pitch = (uint32_t)a.NPa * 16u;

    if (tid == MMA_WARP * 32) {
        for (int s = 0; s < stages; ++s) {
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
        tmem_alloc(tmem_ptr_addr, 2 * MT * BN);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

    if (warp < 4) {
        // ============================== window / weight producer ==============================
        const int cg = tid & 1;               // 8-channel group of the 16-channel block this thread copies
        const int p0 = tid >> 1;              // positions p0, p0 + 64, ...
        const size_t w_tile_elems = (size_t)a.K * a.CB * (a.b_chunk_bytes / 2);
        uint32_t it = 0;
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            const int nt = t % a.n_tiles;
            const int q0 = (t / a.n_tiles) * TILE_POS;
            const bf16* wt = a.w + (size_t)nt * w_tile_elems;
            for (int kg = 0; kg < KG; ++kg) {
                // source pixel of every window position this thread owns (padding / upsampling resolved here, once)
                int off[MAXP];
                const int qb = q0 + kg * a.KYG * a.Wp;
#pragma unroll
                for (int i = 0; i < MAXP; ++i) {
                    const int p = p0 + 64 * i;
                    int o = -2;                                   // -2: outside the window, nothing to copy
                    if (p < a.NP) {
                        const int q = qb + p;
                        o = -1;                                   // -1: contributes zeros
                        if (q < a.Q) {
                            const int xp = q % a.Wp;
                            const int r = q / a.Wp;
                            const int yp = r % a.Hp;
                            const int n = r / a.Hp;
                            const int sy = map_coord(yp - a.pad, a.Hv, a.pad_mode, a.up, 1);
                            const int sx = map_coord(xp - a.pad, a.Wv, a.pad_mode, a.up, 1);
                            if (sy >= 0 && sx >= 0) o = (n * a.H + sy) * a.W + sx;
                        }
                    }
                    off[i] = o;
                }
                for (int cb = 0; cb < a.CB; ++cb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (it / stages) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    if (tid == 0) {
                        mbar_expect_tx(full_bar(s), (uint32_t)b_bytes);
                        for (int r = 0; r < a.KYG; ++r)
                            bulk_copy_g2s(b_smem(s) + r * a.b_chunk_bytes,
                                          wt + ((size_t)(kg * a.KYG + r) * a.CB + cb) * (a.b_chunk_bytes / 2),
                                          (uint32_t)a.b_chunk_bytes, full_bar(s));
                    }
                    const int c = cb * 16 + cg * 8;
                    const bool cok = c < a.Cs;
                    const uint32_t dst0 = a_smem(s) + (uint32_t)cg * plane_pitch + (uint32_t)p0 * 16u;
#pragma unroll
                    for (int i = 0; i < MAXP; ++i) {
                        if (off[i] != -2) {
                            const bool ok = cok && off[i] >= 0;
                            const bf16* src = ok ? a.x + ((size_t)off[i] * a.Cs + c) : a.x;
                            const uint32_t dst = dst0 + (uint32_t)(64 * i) * 16u;
                            cp_async_16(dst, src, ok ? 16u : 0u);
                            if (NPL == 2) cp_async_16(dst + 2u * plane_pitch, ok ? src + a.x_plane : src, ok ? 16u : 0u);
                        }
                    }
                    cp_async_commit();
                    if (it >= 1u) {
                        cp_async_wait<1>();
                        fence_proxy_async();
                        mbar_arrive(full_bar((it - 1u) % stages));
                    }
                }
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        if (it >= 1u) mbar_arrive(full_bar((it - 1u) % stages));
    } else if (warp == MMA_WARP) {
        // ============================== MMA issuer (one thread) ==============================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16_kk(BN);
            const uint32_t b_plane = 2u * BN * 16u;               // bytes of one [cgroup][BN][8] weight image
            uint32_t it = 0, tl = 0;
            for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++tl) {
                const uint32_t acc = tl & 1u, acc_ph = (tl >> 1) & 1u;
                mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u);
                tc_fence_after();
                const uint32_t d0 = tmem_base + acc * (MT * BN);
                for (int st = 0; st < KG * a.CB; ++st, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (it / stages) & 1u;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t ab = a_smem(s), bb = b_smem(s);
                    for (int r = 0; r < a.KYG; ++r) {
                        for (int kx = 0; kx < a.K; ++kx) {
                            const uint32_t tap = (uint32_t)(r * a.K + kx);
                            const uint64_t b_hi = make_kmajor_nosw_desc(bb + tap * NPL * b_plane, BN * 16u, 128u);
                            const uint64_t b_lo = make_kmajor_nosw_desc(bb + (tap * NPL + 1u) * b_plane, BN * 16u, 128u);
                            const uint32_t shift = (uint32_t)(r * a.Wp + kx) * 16u;
                            const uint32_t first = (uint32_t)((st | (int)tap) != 0);
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) {
                                const uint32_t aoff = ab + shift + (uint32_t)mt * 2048u;
                                const uint64_t a_hi = make_kmajor_nosw_desc(aoff, plane_pitch, 128u);
                                umma_bf16(d0 + mt * BN, a_hi, b_hi, idesc, first);
                                if (NPASS == 3) {
                                    const uint64_t a_lo = make_kmajor_nosw_desc(aoff + 2u * plane_pitch, plane_pitch, 128u);
                                    umma_bf16(d0 + mt * BN, a_lo, b_hi, idesc, 1u);
                                    umma_bf16(d0 + mt * BN, a_hi, b_lo, idesc, 1u);
                                }
                            }
                        }
                    }
                    umma_commit(empty_bar(s));
                }
                umma_commit(tmem_full_bar(acc));
            }
        }
    } else {
        // ============================== epilogue ==============================
        const int wq = warp & 3;          // TMEM lane quarter this warp may read
        uint32_t tl = 0;
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++tl) {
            const uint32_t acc = tl & 1u, acc_ph = (tl >> 1) & 1u;
            const int n0 = (t % a.n_tiles) * BN;
            const int q0 = (t / a.n_tiles) * TILE_POS;
            mbar_wait(tmem_full_bar(acc), acc_ph);
            tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const int q = q0 + mt * 128 + wq * 32 + lane;
                const int xp = q % a.Wp;
                const int r = q / a.Wp;
                const int yp = r % a.Hp;
                const int n = r / a.Hp;
                const bool valid = n < a.N && yp < a.Ho && xp < a.Wo;
                const long long m = ((long long)n * a.Ho + yp) * a.Wo + xp;
                if (__ballot_sync(0xffffffffu, valid) == 0u) continue;      // warp-uniform: a run of padding positions
#pragma unroll 1
                for (int j = 0; j < BN / 16; ++j) {
                    const int nb = n0 + j * 16;
                    if (nb >= a.Cout) break;                                // warp-uniform
                    uint32_t raw[16];
                    tmem_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + acc * (MT * BN) + (uint32_t)(mt * BN + j * 16), raw);
                    if (valid) {
                        float* yrow = a.y + m * a.out_pitch + nb;
                        if (a.vec_ok) {
                            float v[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]) + (a.bias ? __ldg(a.bias + nb + i) : 0.f);
                            if (a.addend) {
                                const float4* ad = reinterpret_cast<const float4*>(a.addend + m * a.out_pitch + nb);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float4 f = ad[i];
                                    v[4 * i] += f.x; v[4 * i + 1] += f.y; v[4 * i + 2] += f.z; v[4 * i + 3] += f.w;
                                }
                            }
                            if (a.post_act != ACT_NONE) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] = act_apply(v[i], a.post_act);
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                reinterpret_cast<float4*>(yrow)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                if (nb + i < a.Cout) {
                                    float rr = __uint_as_float(raw[i]) + (a.bias ? __ldg(a.bias + nb + i) : 0.f);
                                    if (a.addend) rr += a.addend[m * a.out_pitch + nb + i];
                                    yrow[i] = act_apply(rr, a.post_act);
                                }
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
        tmem_dealloc(tmem_base, 2 * MT * BN);
    }
}

// weights: OIHW fp32 -> [n_tile][ky][cb][kx][plane][cgroup][BN][8] bf16 (plane 1 = bf16 remainder, when npl == 2)
__global__ void pack_weight_shift_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int K,
                                         int transpose_flip, int BN, int ntiles, int CB, int npl) {
    const int Od = transpose_flip ? Cin : Cout, Id = transpose_flip ? Cout : Cin;
    const long long total = (long long)ntiles * K * CB * K * 2 * BN * 8;      // one plane
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long r = idx;
        const int e = (int)(r % 8); r /= 8;
        const int n = (int)(r % BN); r /= BN;
        const int cgp = (int)(r % 2); r /= 2;
        const int kx = (int)(r % K); r /= K;
        const int cb = (int)(r % CB); r /= CB;
        const int ky = (int)(r % K); r /= K;
        const int nt = (int)r;
        const int o = nt * BN + n, i = cb * 16 + cgp * 8 + e;
        float v = 0.f;
        if (o < Od && i < Id) {
            if (transpose_flip)
                v = w[(((long long)i * Cin + o) * K + (K - 1 - ky)) * K + (K - 1 - kx)];
            else
                v = w[(((long long)o * Cin + i) * K + ky) * K + kx];
        }
        const bf16 hi = __float2bfloat16_rn(v);
        // destination: ((((nt*K + ky)*CB + cb)*K + kx)*npl + pl)*2 + cgp)*BN + n)*8 + e
        const long long base = ((((long long)nt * K + ky) * CB + cb) * K + kx) * npl;
        bf16* dst = out + (((base * 2 + cgp) * BN + n) * 8 + e);
        dst[0] = hi;
        if (npl == 2) dst[2LL * BN * 8] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

struct ShPlan {
    int bn, KYG, NP, NPa, stages, a_bytes, b_chunk_bytes, smem_bytes;
};

int shift_block_n(int cout) { return cout > 32 ? 64 : cout > 16 ? 32 : 16; }

// stage geometry for a convolution; returns 0 if the shifted kernel cannot take it
int make_plan(const ConvGeom& g, int passes, ShPlan& p) {
    if (g.stride != 1 || g.zi != 1 || g.KH != g.KW) return 0;
    if (g.Cin % 8 != 0 || g.in_pitch != g.Cin) return 0;
    const int npl = passes == 3 ? 2 : 1;
    const long long Hp = g.Hv + 2 * g.pad, Wp = g.Wv + 2 * g.pad;
    if (Hp - g.KH + 1 != g.Ho || Wp - g.KW + 1 != g.Wo) return 0;
    // 1x1 filters have no tap reuse to exploit, and on tiny maps the padding positions (computed, then dropped) cost more
    // than the window saves: both stay on the im2col kernel
    if (g.KH == 1 || 2 * Hp * Wp > 3LL * g.Ho * g.Wo) return 0;
    if ((long long)g.N * Hp * Wp + 2048 >= (1LL << 31) / 16) return 0;           // 32-bit position / byte arithmetic
    if ((long long)g.N * g.H * g.W * g.Cin >= (1LL << 31)) return 0;
    p.bn = shift_block_n(g.Cout);
    for (int mode = 0; mode < 2; ++mode) {
        p.KYG = mode == 0 ? g.KH : 1;
        if (mode == 1 && g.KH == 1) break;
        p.NP = TILE_POS + (p.KYG - 1) * (int)Wp + g.KW - 1;
        if (p.NP > 64 * MAXP) continue;
        p.NPa = (p.NP + 7) / 8 * 8;
        p.a_bytes = npl * 2 * p.NPa * 16;
        p.b_chunk_bytes = g.KW * npl * 2 * p.bn * 16;
        const int stage = p.a_bytes + p.KYG * p.b_chunk_bytes;
        const int st = (SMEM_LIMIT - 128 - 256) / stage;
        if (st < 2) continue;
        p.stages = st > MAX_STAGES ? MAX_STAGES : st;
        p.smem_bytes = p.stages * stage + 128 + 256;
        return 1;
    }
    return 0;
}

template <int BN, int NPASS>
int launch_shift(const ShArgs& args, int smem_bytes, cudaStream_t st) {
    auto kern = conv_shift_tcgen05_kernel<BN, NPASS>;
    static int configured = 0;
    if (configured < smem_bytes) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) {
            affgw_set_error("conv_shift: cannot reserve %d bytes of shared memory", SMEM_LIMIT);
            return -2;
        }
        configured = SMEM_LIMIT;
    }
    const int grid = args.total_tiles < 148 ? args.total_tiles : 148;
    kern<<<grid, NUM_THREADS, smem_bytes, st>>>(args);
    AFFGW_LAUNCH_CHECK("conv_shift_tcgen05");
    return 0;
}

}  // namespace

int conv_shift_ok(const ConvGeom& g, int passes) {
    ShPlan p;
    return make_plan(g, passes, p);
}

long long pack_weight_shift_bytes(int Cout, int Cin, int K, int ipad, int transpose_flip, int passes) {
    const int Od = transpose_flip ? Cin : Cout, Id = transpose_flip ? Cout : Cin;
    if (ipad % 8 != 0 || ipad < Id || (passes != 1 && passes != 3)) return -1;
    const int bn = shift_block_n(Od);
    const long long ntiles = (Od + bn - 1) / bn, CB = (ipad + 15) / 16;
    return ntiles * K * CB * K * (passes == 3 ? 2 : 1) * 2 * bn * 8 * 2;
}

int pack_weight_shift(const float* w, void* out, int Cout, int Cin, int K, int ipad, int transpose_flip, int passes,
                      cudaStream_t st) {
    if (pack_weight_shift_bytes(Cout, Cin, K, ipad, transpose_flip, passes) <= 0) {
        affgw_set_error("pack_weight_shift: bad configuration (i_pad %d, passes %d)", ipad, passes);
        return -1;
    }
    const int Od = transpose_flip ? Cin : Cout;
    const int bn = shift_block_n(Od);
    const int ntiles = (Od + bn - 1) / bn, CB = (ipad + 15) / 16;
    const long long total = (long long)ntiles * K * CB * K * 2 * bn * 8;
    const int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    pack_weight_shift_kernel<<<blocks, 256, 0, st>>>(w, (bf16*)out, Cout, Cin, K, transpose_flip, bn, ntiles, CB,
                                                     passes == 3 ? 2 : 1);
    AFFGW_LAUNCH_CHECK("pack_weight_shift");
    return 0;
}

int conv_fwd_shift(const void* x_planes, long long plane_stride, const void* w_packed, const float* bias, const void* addend,
                   void* y, int y_dt, const ConvGeom& g, int passes, cudaStream_t st) {
    ShPlan p;
    if (y_dt != AFFGW_F32 || !make_plan(g, passes, p)) {
        affgw_set_error("conv_shift: unsupported convolution (stride %d, stored Cin %d, passes %d)", g.stride, g.Cin, passes);
        return -1;
    }
    ShArgs a;
    a.x = (const bf16*)x_planes;
    a.x_plane = plane_stride;
    a.w = (const bf16*)w_packed;
    a.bias = bias;
    a.addend = (const float*)addend;
    a.y = (float*)y;
    a.N = g.N; a.H = g.H; a.W = g.W; a.Cs = g.Cin;
    a.up = g.up; a.pad = g.pad; a.pad_mode = g.pad_mode; a.K = g.KH;
    a.Hv = g.Hv; a.Wv = g.Wv; a.Hp = g.Hv + 2 * g.pad; a.Wp = g.Wv + 2 * g.pad; a.Ho = g.Ho; a.Wo = g.Wo;
    a.Cout = g.Cout; a.out_pitch = g.out_pitch; a.post_act = g.post_act;
    a.CB = (g.Cin + 15) / 16;
    a.KYG = p.KYG; a.NP = p.NP; a.NPa = p.NPa; a.stages = p.stages;
    a.a_bytes = p.a_bytes; a.b_chunk_bytes = p.b_chunk_bytes;
    a.n_tiles = (g.Cout + p.bn - 1) / p.bn;
    a.Q = g.N * a.Hp * a.Wp;
    const long long q_last = ((long long)(g.N - 1) * a.Hp + g.Ho - 1) * a.Wp + g.Wo - 1;
    a.total_tiles = (int)((q_last / TILE_POS + 1) * a.n_tiles);
    a.vec_ok = (g.Cout % 16 == 0) && (g.out_pitch % 4 == 0) && (((uintptr_t)y) % 16 == 0) &&
               (!addend || ((uintptr_t)addend) % 16 == 0);
    if (passes == 3) {
        switch (p.bn) {
            case 16: return launch_shift<16, 3>(a, p.smem_bytes, st);
            case 32: return launch_shift<32, 3>(a, p.smem_bytes, st);
            default: return launch_shift<64, 3>(a, p.smem_bytes, st);
        }
    }
    switch (p.bn) {
        case 16: return launch_shift<16, 1>(a, p.smem_bytes, st);
        case 32: return launch_shift<32, 1>(a, p.smem_bytes, st);
        default: return launch_shift<64, 1>(a, p.smem_bytes, st);
    }
}

// =====================================================================================================================
// Weight gradient in the same position space:
//
//   dW[co][ky][kx][ci] = sum_q  V[q + ky*Wp + kx][ci] * dYp[q][co]         (dYp = dY scattered to padded positions)
//
// a GEMM whose reduction runs over positions.  Both operands sit in shared memory in the planar layout above, which is
// UMMA's un-swizzled "MN-major" form (8 positions x 8 channels per core matrix; LBO = 128 B to the next 8 positions,
// SBO = plane pitch to the next 8 channels).  One CTA owns a (128 input channels) x (BN output channels) x (kernel row ky)
// block of dW: for every 16 positions it issues one MMA per kx whose A descriptor starts kx positions further into the
// window - the input window is read once per kernel ROW instead of once per tap, dY once per row.  M is always 128
// (absent channel groups stay zero in shared memory), so thin layers cost N/2 clocks per MMA, not a 128x128 tile.
// The position range is split over grid.y; partials are reduced with coalesced fp32 atomics into ws[co][tap][ci].
// =====================================================================================================================
namespace {

constexpr int WG_THREADS = 160;              // 4 producer / epilogue warps + 1 MMA warp

struct WsArgs {
    const bf16* x; long long x_plane;        // planes [NPL][N*H*W][Cs]
    const bf16* dy; long long dy_plane;      // planes [NPL][N*Ho*Wo][Cys]
    float* ws;                               // [Cout][taps][Cs] fp32
    int N, H, W, Cs;
    int up, pad, pad_mode, K;
    int Hv, Wv, Hp, Wp, Ho, Wo;
    int Cout, Cys;
    int KP, pitchA16, pitchB16;              // positions per stage; plane pitches in 16-byte units
    int n_co_blocks;
    int chunks_total, chunks_per_split;
    int a_bytes, b_bytes, stages;
    int Q;
};

__device__ __forceinline__ uint64_t make_mnmajor_nosw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;     // to the next 8 positions (K)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;     // to the next 8 channels (M / N)
    d |= 1ull << 46;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_mnmn(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int BN, int NPASS>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_shift_kernel(const __grid_constant__ WsArgs a) {
    constexpr int NPL = NPASS == 3 ? 2 : 1;
    constexpr int GB = BN / 8;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    const int stages = a.stages;
    const int stage_bytes = a.a_bytes + a.b_bytes;
    const uint32_t bar_base = smem_base + stages * stage_bytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * MAX_STAGES);
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * MAX_STAGES + 1);
    auto a_smem = [&](int s) { return smem_base + s * stage_bytes; };
    auto b_smem = [&](int s) { return smem_base + s * stage_bytes + a.a_bytes; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ky = blockIdx.x % a.K;
    const int cob = (blockIdx.x / a.K) % a.n_co_blocks;
    const int cib = blockIdx.x / (a.K * a.n_co_blocks);
    const int cbeg = blockIdx.y * a.chunks_per_split;
    const int cend = min(a.chunks_total, cbeg + a.chunks_per_split);
    const int nst = cend - cbeg;
    constexpr int TMEM_COLS = 512;

    // absent channel groups must read as zeros for the whole kernel: clear every stage once
    for (uint32_t o = tid * 16u; o < (uint32_t)(stages * stage_bytes); o += WG_THREADS * 16u)
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(smem_base + o), "r"(0u) : "memory");
    fence_proxy_async();
    if (tid == 128) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar(s), 128);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        __syncwarp();
        tmem_alloc(tmem_ptr_addr, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
    const uint32_t pitchA = (uint32_t)a.pitchA16 * 16u, pitchB = (uint32_t)a.pitchB16 * 16u;

    if (warp < 4) {
        // ============================== producer ==============================
        const int g = tid & 15, jj = tid >> 4;            // channel group, position slot (8 positions per pass)
        const int ca = cib * 128 + g * 8;                 // input channel of this thread's group
        const int cbn = cob * BN + g * 8;                 // output channel of this thread's group
        const bool a_on = ca < a.Cs, b_on = g < GB && cbn < a.Cys;
        const int wlen = a.KP + a.K - 1;
        for (int st = 0; st < nst; ++st) {
            const int s = st % stages;
            const uint32_t ph = (uint32_t)(st / stages) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            const int q0 = (cbeg + st) * a.KP;
            if (a_on) {
                const uint32_t dst0 = a_smem(s) + (uint32_t)g * pitchA;
                // decode the first position with divisions, then walk: 8 positions further per iteration
                const int qf = q0 + ky * a.Wp + jj;
                int xp = qf % a.Wp;
                const int r = qf / a.Wp;
                int yp = r % a.Hp, n = r / a.Hp;
                for (int j = jj; j < wlen; j += 8) {
                    bool ok = n < a.N;
                    const bf16* src = a.x;
                    if (ok) {
                        const int sy = map_coord(yp - a.pad, a.Hv, a.pad_mode, a.up, 1);
                        const int sx = map_coord(xp - a.pad, a.Wv, a.pad_mode, a.up, 1);
                        ok = sy >= 0 && sx >= 0;
                        if (ok) src = a.x + ((size_t)((n * a.H + sy) * a.W + sx) * a.Cs + ca);
                    }
                    const uint32_t dst = dst0 + (uint32_t)j * 16u;
                    cp_async_16(dst, src, ok ? 16u : 0u);
                    if (NPL == 2) cp_async_16(dst + 16u * pitchA, ok ? src + a.x_plane : src, ok ? 16u : 0u);
                    xp += 8;
                    while (xp >= a.Wp) {
                        xp -= a.Wp;
                        if (++yp == a.Hp) { yp = 0; ++n; }
                    }
                }
            }
            if (b_on) {
                const uint32_t dst0 = b_smem(s) + (uint32_t)g * pitchB;
                const int qf = q0 + jj;
                int xp = qf % a.Wp;
                const int r = qf / a.Wp;
                int yp = r % a.Hp, n = r / a.Hp;
                for (int j = jj; j < a.KP; j += 8) {
                    const bool ok = n < a.N && yp < a.Ho && xp < a.Wo;
                    const bf16* src = ok ? a.dy + ((size_t)((n * a.Ho + yp) * a.Wo + xp) * a.Cys + cbn) : a.dy;
                    const uint32_t dst = dst0 + (uint32_t)j * 16u;
                    cp_async_16(dst, src, ok ? 16u : 0u);
                    if (NPL == 2) cp_async_16(dst + (uint32_t)GB * pitchB, ok ? src + a.dy_plane : src, ok ? 16u : 0u);
                    xp += 8;
                    while (xp >= a.Wp) {
                        xp -= a.Wp;
                        if (++yp == a.Hp) { yp = 0; ++n; }
                    }
                }
            }
            cp_async_commit();
            if (st >= 1) {
                cp_async_wait<1>();
                fence_proxy_async();
                mbar_arrive(full_bar((st - 1) % stages));
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        if (nst >= 1) mbar_arrive(full_bar((nst - 1) % stages));

        // ============================== epilogue: fp32 reductions into ws[co][tap][ci] ==============================
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int ci = cib * 128 + warp * 32 + lane;
        const bool rok = ci < a.Cs;
        const int taps = a.K * a.K;
        for (int kx = 0; kx < a.K; ++kx) {
            float* wbase = a.ws + (size_t)(ky * a.K + kx) * a.Cs + ci;
#pragma unroll 1
            for (int jb = 0; jb < BN / 16; ++jb) {
                const int nb = cob * BN + jb * 16;
                if (nb >= a.Cout) break;
                uint32_t raw[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kx * BN + jb * 16), raw);
                if (rok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (nb + i < a.Cout) atomicAdd(wbase + (size_t)(nb + i) * taps * a.Cs, __uint_as_float(raw[i]));
                }
            }
        }
        tc_fence_before();
    } else if (lane == 0) {
        // ============================== MMA issuer ==============================
        constexpr uint32_t idesc = make_idesc_bf16_mnmn(BN);
        for (int st = 0; st < nst; ++st) {
            const int s = st % stages;
            const uint32_t ph = (uint32_t)(st / stages) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t ab = a_smem(s), bb = b_smem(s);
            for (int k = 0; k < a.KP / 16; ++k) {
                const uint64_t b_hi = make_mnmajor_nosw_desc(bb + k * 256u, 128u, pitchB);
                const uint64_t b_lo = make_mnmajor_nosw_desc(bb + (uint32_t)GB * pitchB + k * 256u, 128u, pitchB);
                for (int kx = 0; kx < a.K; ++kx) {
                    const uint32_t astart = ab + (uint32_t)(k * 16 + kx) * 16u;
                    const uint64_t a_hi = make_mnmajor_nosw_desc(astart, 128u, pitchA);
                    const uint32_t accum = (uint32_t)((st | k) != 0);
                    umma_bf16(tmem_base + kx * BN, a_hi, b_hi, idesc, accum);
                    if (NPASS == 3) {
                        const uint64_t a_lo = make_mnmajor_nosw_desc(astart + 16u * pitchA, 128u, pitchA);
                        umma_bf16(tmem_base + kx * BN, a_lo, b_hi, idesc, 1u);
                        umma_bf16(tmem_base + kx * BN, a_hi, b_lo, idesc, 1u);
                    }
                }
            }
            umma_commit(empty_bar(s));
        }
        umma_commit(tmem_full_bar);
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

struct WsPlan {
    int bn, KP, pitchA16, pitchB16, a_bytes, b_bytes, stages, smem_bytes;
};

int make_wg_plan(const ConvGeom& g, int passes, WsPlan& p) {
    if (g.stride != 1 || g.zi != 1 || g.KH != g.KW || g.KH > 7) return 0;
    if (g.Cin % 8 != 0 || g.in_pitch != g.Cin || g.out_pitch % 8 != 0 || g.out_pitch < g.Cout) return 0;
    if (g.pre_act != ACT_NONE) return 0;
    if (g.Cin < 128 || g.KH == 1) return 0;                     // measured: thin layers and 1x1 filters are faster on the im2col wgrad
    const int npl = passes == 3 ? 2 : 1;
    const long long Hp = g.Hv + 2 * g.pad, Wp = g.Wv + 2 * g.pad;
    if (Hp - g.KH + 1 != g.Ho || Wp - g.KW + 1 != g.Wo) return 0;
    if (2 * Hp * Wp > 3LL * g.Ho * g.Wo) return 0;              // tiny maps: padding positions dominate, im2col kernel
    if ((long long)g.N * Hp * Wp + 4096 >= (1LL << 31) / 16) return 0;
    if ((long long)g.N * g.H * g.W * g.Cin >= (1LL << 31) || g.M * g.out_pitch >= (1LL << 31)) return 0;
    const int limit = g.KH <= 4 ? 128 : 64;                     // K * BN TMEM columns <= 512
    p.bn = g.Cout <= 16 ? 16 : g.Cout <= 32 ? 32 : g.Cout <= 64 ? 64 : limit;
    p.KP = (g.Cin <= 32 && p.bn <= 32) ? 128 : 64;
    p.pitchA16 = p.KP + 9;
    p.pitchB16 = p.KP + 1;
    p.a_bytes = npl * 16 * p.pitchA16 * 16;
    p.b_bytes = npl * (p.bn / 8) * p.pitchB16 * 16;
    const int stage = p.a_bytes + p.b_bytes;
    const int st = (SMEM_LIMIT - 128 - 256) / stage;
    if (st < 2) return 0;
    p.stages = st > MAX_STAGES ? MAX_STAGES : st;
    p.smem_bytes = p.stages * stage + 128 + 256;
    return 1;
}

template <int BN, int NPASS>
int launch_wg_shift(const WsArgs& args, dim3 grid, int smem_bytes, cudaStream_t st) {
    auto kern = conv_wgrad_shift_kernel<BN, NPASS>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) {
            affgw_set_error("conv_wgrad_shift: cannot reserve %d bytes of shared memory", SMEM_LIMIT);
            return -2;
        }
        configured = true;
    }
    kern<<<grid, WG_THREADS, smem_bytes, st>>>(args);
    AFFGW_LAUNCH_CHECK("conv_wgrad_shift");
    return 0;
}

}  // namespace

int conv_wgrad_shift_ok(const ConvGeom& g, int passes) {
    WsPlan p;
    return make_wg_plan(g, passes, p);
}

// ws: [Cout][taps][g.Cin] fp32, zeroed by the caller; g.Cin / g.out_pitch are the stored channel counts of the planes
int conv_wgrad_shift(const void* x_planes, long long x_plane, const void* dy_planes, long long dy_plane, float* ws,
                     const ConvGeom& g, int passes, cudaStream_t st) {
    WsPlan p;
    if (!make_wg_plan(g, passes, p)) {
        affgw_set_error("conv_wgrad_shift: unsupported convolution");
        return -1;
    }
    WsArgs a;
    a.x = (const bf16*)x_planes; a.x_plane = x_plane;
    a.dy = (const bf16*)dy_planes; a.dy_plane = dy_plane;
    a.ws = ws;
    a.N = g.N; a.H = g.H; a.W = g.W; a.Cs = g.Cin;
    a.up = g.up; a.pad = g.pad; a.pad_mode = g.pad_mode; a.K = g.KH;
    a.Hv = g.Hv; a.Wv = g.Wv; a.Hp = g.Hv + 2 * g.pad; a.Wp = g.Wv + 2 * g.pad; a.Ho = g.Ho; a.Wo = g.Wo;
    a.Cout = g.Cout; a.Cys = g.out_pitch;
    a.KP = p.KP; a.pitchA16 = p.pitchA16; a.pitchB16 = p.pitchB16;
    a.a_bytes = p.a_bytes; a.b_bytes = p.b_bytes; a.stages = p.stages;
    a.Q = g.N * a.Hp * a.Wp;
    const int n_ci = (g.Cin + 127) / 128;
    a.n_co_blocks = (g.Cout + p.bn - 1) / p.bn;
    const long long q_last = ((long long)(g.N - 1) * a.Hp + g.Ho - 1) * a.Wp + g.Wo - 1;
    a.chunks_total = (int)(q_last / p.KP + 1);
    const int tiles = n_ci * a.n_co_blocks * g.KH;
    // split the position range so that the CTAs fill whole waves of 148 SMs (each split keeps >= 8 stages of work)
    int max_splits = (a.chunks_total + 7) / 8;
    if (max_splits > 320 / tiles + 1) max_splits = 320 / tiles + 1;
    int splits = 1;
    double best = 0.0;
    for (int sp = 1; sp <= max_splits; ++sp) {
        const int ctas = tiles * sp;
        const double eff = (double)ctas / (double)((ctas + 147) / 148 * 148);
        if (eff > best + 0.02) { best = eff; splits = sp; }
    }
    a.chunks_per_split = (a.chunks_total + splits - 1) / splits;
    splits = (a.chunks_total + a.chunks_per_split - 1) / a.chunks_per_split;
    dim3 grid((unsigned)tiles, (unsigned)splits);
    if (passes == 3) {
        switch (p.bn) {
            case 16: return launch_wg_shift<16, 3>(a, grid, p.smem_bytes, st);
            case 32: return launch_wg_shift<32, 3>(a, grid, p.smem_bytes, st);
            case 64: return launch_wg_shift<64, 3>(a, grid, p.smem_bytes, st);
            default: return launch_wg_shift<128, 3>(a, grid, p.smem_bytes, st);
        }
    }
    switch (p.bn) {
        case 16: return launch_wg_shift<16, 1>(a, grid, p.smem_bytes, st);
        case 32: return launch_wg_shift<32, 1>(a, grid, p.smem_bytes, st);
        case 64: return launch_wg_shift<64, 1>(a, grid, p.smem_bytes, st);
        default: return launch_wg_shift<128, 1>(a, grid, p.smem_bytes, st);
    }
}
