// Stride-1 convolutions (forward, input gradient, weight gradient) as "shifted" GEMMs in POSITION SPACE on the
// 5th-generation tensor cores (tcgen05.mma, fp32 accumulators in TMEM).
//
// The im2col kernel in conv_tc.cu re-reads every input element once per filter tap from L2 (9x for 3x3, 25x for 5x5,
// 49x for 7x7); at ~42 B/clk/SM of L2->SM bandwidth that, not the tensor pipe, bounded it.  Here every tensor of one
// convolution lives in the position space of its PADDED (and, for the decoder, x2-upsampled) input:
//
//   frame [N][Hp][Wp], Hp = up*H + 2*pad; flattened position q = (n*Hp + yp)*Wp + xp
//   V[q][ci]   the padded / upsampled / pre-activated input          ("x position planes")
//   Yp[q][co]  the output at its top-left-aligned position, zero where yp >= Ho or xp >= Wo ("dY position planes")
//
//   forward   Y[q]        = sum_{ky,kx} V[q + ky*Wp + kx] . W[ky][kx]
//   dgrad     dV[p]       = sum_{ky,kx} dYp[p - (K-1)(Wp+1) + ky*Wp + kx] . Wflip[ky][kx]      (any padding mode; the
//                           reflect / upsample fold or the zero-pad crop happens on dV afterwards / in the epilogue)
//   wgrad     dW[ky][kx]  = sum_q V[q + ky*Wp + kx]^T . dYp[q]
//
// Operand planes are stored PLANAR in HBM: [hi|lo plane][8-channel group][position] x 16 bytes, written by
// split_positions_kernel (padding mode, upsampling, activation-first LeakyReLU and the split-bf16 remainder are all
// resolved there, once per element).  Eight consecutive positions of one group are 128 contiguous bytes = one UMMA
// core matrix of the un-swizzled layouts (K-major for forward / dgrad, MN-major for wgrad), so
//   * a pipeline stage is filled by a handful of cp.async.bulk copies (TMA engine, mbarrier complete_tx) of CONTIGUOUS
//     runs of positions - no per-thread gather, no index arithmetic in the main loop;
//   * every filter tap is just a different descriptor start address inside the same shared-memory window.
//
// NPASS = 3: split-bf16 product a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (hi = bf16(v), lo = bf16(v - hi)).
//
// Replaces the cuDNN convolutions (forward and backward) behind reference blocks.py:148 /
// vgg_tro_channel3_modi.py:47 / modules_tro.py:594-603 and loss.backward() (network_tro.py:55,102,113,129).
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "pos_frame.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int MMA_WARP = 1;
#ifndef AFFGW_EPI_WARPS
#define AFFGW_EPI_WARPS 8
#endif
constexpr int EPI_WARPS = AFFGW_EPI_WARPS;   // 4 or 8: one or two per TMEM lane quarter (a warp reads the quarter warp_id % 4)
constexpr int NUM_THREADS = 32 * (2 + EPI_WARPS);   // producer warp, MMA warp, epilogue warps
constexpr int MAX_TILE_POS = 512;            // largest CTA tile in positions (frames keep that much tail)
// CTA tile = MT M-tiles of 128 positions x BN output channels; NBUF TMEM accumulator buffers of MT*BN columns.
// Wide N amortises the 4 KB A read of a 128-row MMA over more tensor clocks (the shared-memory pipe delivers 128 B/clk):
// BN = 256 needs 12 KB per 128 clocks, BN = 64 needs 6 KB per 32.
// Split-bf16 mode with BN <= 64 packs [w_hi | w_lo] along N: a_hi x [w_hi | w_lo] is ONE MMA of width 2 BN (two accumulator
// column blocks, summed in the epilogue), a_lo x w_hi a second one - two reads of the A tile per tap instead of three, which
// is what bounds these thin layers (4 KB of A per MMA at 128 B/clk against N/2 tensor clocks).
__host__ __device__ constexpr bool tile_packed(int bn, int npl) { return npl == 2 && bn <= 64; }
__host__ __device__ constexpr int tile_acc_w(int bn, int npl) { return tile_packed(bn, npl) ? 2 * bn : bn; }
__host__ __device__ constexpr int tile_mt(int bn, int npl) { return bn <= 32 ? 4 : (bn == 64 ? (npl == 2 ? 2 : 4) : 2); }
__host__ __device__ constexpr int tile_nbuf(int acc_cols) { return 2 * acc_cols <= 512 ? 2 : 1; }   // accumulator buffers in TMEM
constexpr int MAX_STAGES = 6;
constexpr int SMEM_LIMIT = 227 * 1024;

struct ShArgs {
    const bf16* x;             // position planes [NPL][G][QA][8]
    const bf16* w;             // [n_tile][ky][cb][kx][cgroup][plane][BN][8]
    const float* bias;
    const float* addend;
    float* y;
    int N, Hp, Wp, K;
    int G, lead;
    long long QA;
    int q_shift;               // window start relative to the tile's first position (0 forward, -(K-1)(Wp+1) dgrad)
    int oy0, ox0, OH, OW;      // output pixel of position (n, yp, xp) = (n, yp - oy0, xp - ox0) if inside [OH) x [OW)
    int Cout, out_pitch, post_act;
    int CB;                    // 16-channel blocks
    int KYG;                   // kernel rows per pipeline stage: K (whole filter) or 1
    int NP, NPa;               // window positions per stage, shared-memory plane pitch in positions
    int stages;
    int a_bytes, b_chunk_bytes;   // shared-memory bytes of one stage's window / bytes of one (ky, cb) weight chunk
    int n_tiles;
    int total_tiles;
    int vec_ok;
    int fmt;                   // operand format of the planes / packed weights: 0 = bf16, 1 = fp16
    float alpha;               // result = accumulator * alpha * (alpha_dev ? *alpha_dev : 1): undoes the power-of-two operand
    const float* alpha_dev;    // scales of the fp16 planes (weights x 2^8, dY x a per-tensor scale kept in device memory)
    FastDiv div_wp, div_hp;    // multiply-shift division by the frame's row pitch / height
};

// un-swizzled shared-memory matrix descriptors (core matrices of 8 rows x 16 bytes)
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;                                      // descriptor version (sm_100)
    return d;
}
// kind::f16, D = f32, A and B both bf16 (fmt 0) or both fp16 (fmt 1: the two formats cannot be mixed in one instruction -
// measured: illegal instruction, scripts/probes/mixed_mma_probe.cu), M = 128, N = n; mn_major sets the A and B major bits (wgrad)
__host__ __device__ constexpr uint32_t make_idesc(int n, bool mn_major, int fmt = 0) {
    return (1u << 4) | (fmt ? 0u : ((1u << 7) | (1u << 10))) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ---------------------------------------------------------------------------------------- forward / input gradient
// A stage = the window of one 16-channel block (all kernel rows, or one kernel row when the whole-filter window does not
// fit twice in shared memory) + its weights.
template <int BN, int NPASS, int MT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_shift_tcgen05_kernel(const __grid_constant__ ShArgs a) {
    constexpr int NPL = NPASS == 3 ? 2 : 1;
    constexpr bool PK = tile_packed(BN, NPL);
    constexpr int ACC_W = tile_acc_w(BN, NPL);      // accumulator columns per M-tile
    constexpr int NBUF = tile_nbuf(MT * ACC_W), TILE_POS = MT * 128;
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

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int KG = a.K / a.KYG;               // stages per channel block
    const uint32_t plane_pitch = (uint32_t)a.NPa * 16u;

    if (tid == MMA_WARP * 32) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tmem_full_bar(i), 1);
            mbar_init(tmem_empty_bar(i), EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        __syncwarp();
        tmem_alloc(tmem_ptr_addr, NBUF * MT * ACC_W);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

    if (warp == 0) {
        // ============================== producer: bulk copies only ==============================
        if (lane == 0) {
            const size_t w_tile_elems = (size_t)a.K * a.CB * (a.b_chunk_bytes / 2);
            const uint32_t win_bytes = (uint32_t)a.NP * 16u;
            uint32_t it = 0;
            for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
                const int nt = t % a.n_tiles;
                const long long q0 = (long long)(t / a.n_tiles) * TILE_POS + a.q_shift + a.lead;
                const bf16* wt = a.w + (size_t)nt * w_tile_elems;
                for (int kg = 0; kg < KG; ++kg) {
                    const long long qw = q0 + (long long)kg * a.KYG * a.Wp;
                    for (int cb = 0; cb < a.CB; ++cb, ++it) {
                        const int s = it % stages;
                        const uint32_t ph = (it / stages) & 1u;
                        mbar_wait(empty_bar(s), ph ^ 1u);
                        mbar_arrive_expect_tx(full_bar(s), (uint32_t)b_bytes + 2u * NPL * win_bytes);
                        for (int r = 0; r < a.KYG; ++r)
                            bulk_copy_g2s(b_smem(s) + r * a.b_chunk_bytes,
                                          wt + ((size_t)(kg * a.KYG + r) * a.CB + cb) * (a.b_chunk_bytes / 2),
                                          (uint32_t)a.b_chunk_bytes, full_bar(s));
#pragma unroll
                        for (int pl = 0; pl < NPL; ++pl)
#pragma unroll
                            for (int cg = 0; cg < 2; ++cg)
                                bulk_copy_g2s(a_smem(s) + (uint32_t)(pl * 2 + cg) * plane_pitch,
                                              a.x + (((size_t)pl * a.G + cb * 2 + cg) * a.QA + qw) * 8, win_bytes, full_bar(s));
                    }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ============================== MMA issuer ==============================
        {   // every lane runs the loops (warp-uniform), one elected lane issues: see umma_f16_elect32.  Stage / accumulator
            // cursors are counters (no modulo) and descriptors are (low word, constant high word) pairs so that the whole
            // address arithmetic of the issue loop stays on 32-bit warp-uniform values.
            const uint32_t idesc = make_idesc(BN, false, a.fmt);
            const uint32_t idesc2 = make_idesc(PK ? 2 * BN : BN, false, a.fmt);
            const uint32_t b_tap16 = ((uint32_t)NPL * 2u * BN * 16u) >> 4;   // one tap's [cgroup][plane][BN][8] weight image
            // K-major: LBO (bits 16-29) = next 8 channels, SBO (bits 32-45) = next 8 rows = 128 B, version bit 46
            const uint32_t desc_hi = (128u >> 4) | (1u << 14);
            const uint32_t a_lbo = ((plane_pitch >> 4) & 0x3FFFu) << 16;
            const uint32_t b_lbo = ((((uint32_t)NPL * BN * 16u) >> 4) & 0x3FFFu) << 16;
            const uint32_t a_lo_off = (2u * plane_pitch) >> 4, b_lo_off = (BN * 16u) >> 4;   // lo rows follow the hi rows
            uint32_t s = 0, ph = 0, acc = 0, acc_ph = 0;
            for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
                mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u);
                tc_fence_after();
                const uint32_t d0 = tmem_base + acc * (MT * ACC_W);
                for (int st = 0; st < KG * a.CB; ++st) {
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t a_base = a_lbo | ((a_smem(s) & 0x3FFFFu) >> 4);
                    const uint32_t b_base = b_lbo | ((b_smem(s) & 0x3FFFFu) >> 4);
                    uint32_t b_t = b_base;
                    for (int r = 0; r < a.KYG; ++r) {
                        const uint32_t a_r = a_base + (uint32_t)(r * a.Wp);
#pragma unroll 1
                        for (int kx = 0; kx < a.K; ++kx, b_t += b_tap16) {
                            const uint32_t accum = (uint32_t)((st | r | kx) != 0);
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) {
                                const uint32_t a_t = a_r + (uint32_t)kx + (uint32_t)(mt * 128);
                                const uint32_t d = d0 + mt * ACC_W;
                                if (PK) {       // columns [0, BN): a_hi w_hi + a_lo w_hi, columns [BN, 2 BN): a_hi w_lo
                                    umma_f16_elect32(d, a_t, desc_hi, b_t, desc_hi, idesc2, accum);
                                    umma_f16_elect32(d, a_t + a_lo_off, desc_hi, b_t, desc_hi, idesc, 1u);
                                } else {
                                    umma_f16_elect32(d, a_t, desc_hi, b_t, desc_hi, idesc, accum);
                                    if (NPASS == 3) {
                                        umma_f16_elect32(d, a_t + a_lo_off, desc_hi, b_t, desc_hi, idesc, 1u);
                                        umma_f16_elect32(d, a_t, desc_hi, b_t + b_lo_off, desc_hi, idesc, 1u);
                                    }
                                }
                            }
                        }
                    }
                    umma_commit_elect(empty_bar(s));
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1u; }
                }
                umma_commit_elect(tmem_full_bar(acc));
                if (++acc == (uint32_t)NBUF) { acc = 0; acc_ph ^= 1u; }
            }
        }
    } else {
        // ============================== epilogue ==============================
        // Eight warps, two per TMEM lane quarter: the (M-tile, 16-column block) units of a CTA tile alternate between the two
        // warps of a quarter.  On the thin (16/32-channel) layers the epilogue, not the MMAs, is the critical path (ncu: the
        // issuing warp waits for tmem_empty), so a unit is kept short: multiply-shift position decomposition, both column
        // blocks of a packed tile loaded before one wait, bias as four vector loads.
        const int wq = warp & 3;          // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2; // which of the quarter's two warps
        constexpr int NJ = BN / 16;
        const float alpha = a.alpha * (a.alpha_dev ? __ldg(a.alpha_dev) : 1.f);
        uint32_t acc = 0, acc_ph = 0;
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            const int n0 = (t % a.n_tiles) * BN;
            const int q0 = (t / a.n_tiles) * TILE_POS;
            mbar_wait(tmem_full_bar(acc), acc_ph);
            tc_fence_after();
            int mt_cur = -1;
            bool valid = false, any = false;
            long long m = 0;
#pragma unroll 1
            for (int u = half; u < MT * NJ; u += EPI_WARPS / 4) {
                const int mt = u / NJ, j = u - mt * NJ;
                if (mt != mt_cur) {
                    mt_cur = mt;
                    const int q = q0 + mt * 128 + wq * 32 + lane;
                    const int r = (int)fast_div((uint32_t)q, a.div_wp);
                    const int xp = q - r * a.Wp - a.ox0;
                    const int n = (int)fast_div((uint32_t)r, a.div_hp);
                    const int yp = r - n * a.Hp - a.oy0;
                    valid = n < a.N && (unsigned)yp < (unsigned)a.OH && (unsigned)xp < (unsigned)a.OW;
                    m = ((long long)n * a.OH + yp) * a.OW + xp;
                    any = __ballot_sync(0xffffffffu, valid) != 0u;
                }
                if (!any) continue;                                     // warp-uniform: a run of padding positions
                const int nb = n0 + j * 16;
                if (nb >= a.Cout) continue;                             // warp-uniform
                uint32_t raw[16];
                const uint32_t tcol = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * (MT * ACC_W) + (uint32_t)(mt * ACC_W + j * 16);
                tmem_ld16_issue(tcol, raw);
                if (PK) {
                    uint32_t raw2[16];
                    tmem_ld16_issue(tcol + BN, raw2);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) + __uint_as_float(raw2[i]));
                } else {
                    tmem_ld_wait();
                }
                if (valid) {
                    float* yrow = a.y + m * a.out_pitch + nb;
                    if (a.vec_ok) {
                        float v[16];
                        if (a.bias) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + nb) + i);
                                v[4 * i] = fmaf(__uint_as_float(raw[4 * i]), alpha, b4.x);
                                v[4 * i + 1] = fmaf(__uint_as_float(raw[4 * i + 1]), alpha, b4.y);
                                v[4 * i + 2] = fmaf(__uint_as_float(raw[4 * i + 2]), alpha, b4.z);
                                v[4 * i + 3] = fmaf(__uint_as_float(raw[4 * i + 3]), alpha, b4.w);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]) * alpha;
                        }
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
                                float rr = fmaf(__uint_as_float(raw[i]), alpha, a.bias ? __ldg(a.bias + nb + i) : 0.f);
                                if (a.addend) rr += a.addend[m * a.out_pitch + nb + i];
                                yrow[i] = act_apply(rr, a.post_act);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
            if (++acc == (uint32_t)NBUF) { acc = 0; acc_ph ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, NBUF * MT * ACC_W);
    }
}

// weights: OIHW fp32 -> [n_tile][ky][cb][kx][cgroup][plane][BN][8] bf16 (plane 1 = bf16 remainder, when npl == 2): inside a
// channel group the BN remainder rows follow the BN leading rows, so one K-major descriptor covers [w_hi | w_lo] as 2 BN rows
// fmt 1: fp16 planes of w * 2^8 (typical weights ~1e-2 would leave the remainder plane in fp16's subnormal range; the kernels
// multiply the accumulator by 2^-8, exactly)
constexpr float F16_W_SCALE = 256.f;
__device__ __forceinline__ uint16_t f16_bits(float v) {
    const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    return __half_as_ushort(h);
}
__device__ __forceinline__ float f16_val(uint16_t b) { return __half2float(__ushort_as_half(b)); }

__global__ void pack_weight_shift_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int K,
                                         int transpose_flip, int BN, int ntiles, int CB, int npl, int fmt) {
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
        const long long tap = (((long long)nt * K + ky) * CB + cb) * K + kx;
        bf16* dst = out + ((((tap * 2 + cgp) * npl) * BN + n) * 8 + e);
        if (fmt) {
            const float sv = v * F16_W_SCALE;
            const uint16_t hb = f16_bits(sv);
            reinterpret_cast<uint16_t*>(dst)[0] = hb;
            if (npl == 2) reinterpret_cast<uint16_t*>(dst)[(long long)BN * 8] = f16_bits(sv - f16_val(hb));
        } else {
            const bf16 hi = __float2bfloat16_rn(v);
            dst[0] = hi;
            if (npl == 2) dst[(long long)BN * 8] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
    }
}

// ---------------------------------------------------------------------------------------- position planes
// src [N][Hs][Ws][pitch] (fp32 or bf16) -> planes [npl][G][QA][8] bf16.  Frame position (n, yp, xp) takes the source pixel
// map(yp - oy0), map(xp - ox0) of the (x up) virtual source under pad_mode, or zero; the lead / tail margins are zeros.
// One block transposes a tile of 32 positions x 32 channel groups through shared memory so that both the NHWC reads and
// the planar writes are contiguous.
template <typename T, int TG>
__global__ void __launch_bounds__(256)
split_positions_kernel(const T* __restrict__ src, bf16* __restrict__ planes, int N, int Hs, int Ws, int C, int pitch, int up,
                       int oy0, int ox0, int pad_mode, int pre_act, int Hp, int Wp, int G, int lead, int QA, int npl,
                       float* __restrict__ colsum, const FastDiv div_wp, const FastDiv div_hp, int fmt,
                       const float* __restrict__ scale_dev) {
    constexpr int TP = 1024 / TG;                     // positions per tile: a tile is always 1024 (position, group) items
    __shared__ uint4 tile[2][TG * (TP + 1)];          // [plane][group][position], +1 column against bank conflicts
    __shared__ float red[8][TG * 8];
    const int tid = threadIdx.x;
    const int g0 = blockIdx.y * TG;
    const int Q = N * Hp * Wp;
    const bool vec = (pitch % 8 == 0) && (C % 8 == 0);
    const int n_tiles = (QA + TP - 1) / TP;
    float csum[8];                                    // per-channel sums of this thread's items (bias gradient)
#pragma unroll
    for (int i = 0; i < 8; ++i) csum[i] = 0.f;
    const float scale = scale_dev ? __ldg(scale_dev) : 1.f;       // fp16 dY planes: per-tensor power-of-two scale
    for (int tx = blockIdx.x; tx < n_tiles; tx += gridDim.x) {
        const int p0 = tx * TP;                       // stored position (lead included)
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int item = tid + 256 * it;
            const int gl = item % TG, pi = item / TG;
            const int g = g0 + gl;
            const int q = p0 + pi - lead;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
            if (g * 8 < C && q >= 0 && q < Q) {
                // position -> (image, row, column) with multiply-shift divisions; the generic coordinate map only for
                // replicate padding (nearest x2 upsampling is a shift)
                const int r = (int)fast_div((uint32_t)q, div_wp);
                const int xp = q - r * Wp;
                const int n = (int)fast_div((uint32_t)r, div_hp);
                const int yp = r - n * Hp;
                int sy = yp - oy0, sx = xp - ox0;
                const int Hv = Hs * up, Wv = Ws * up;
                if (pad_mode == PAD_REPLICATE) {
                    sy = max(0, min(sy, Hv - 1));
                    sx = max(0, min(sx, Wv - 1));
                } else if (pad_mode == PAD_REFLECT) {
                    sy = sy < 0 ? -sy : (sy >= Hv ? 2 * (Hv - 1) - sy : sy);
                    sx = sx < 0 ? -sx : (sx >= Wv ? 2 * (Wv - 1) - sx : sx);
                }
                const bool inside = (unsigned)sy < (unsigned)Hv && (unsigned)sx < (unsigned)Wv;
                if (up == 2) { sy >>= 1; sx >>= 1; }
                if (inside) {
                    const T* sp = src + ((size_t)((n * Hs + sy) * Ws + sx) * pitch + g * 8);
                    if (vec) {
                        ld8(sp, v);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = (g * 8 + i < C) ? to_f(sp[i]) : 0.f;
                    }
                    if (pre_act != ACT_NONE) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = act_apply(v[i], pre_act);
                    }
                }
            }
            if (colsum != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) csum[i] += v[i];
            }
            uint4 hi, lo;
            if (fmt) {
                __half2* hh = reinterpret_cast<__half2*>(&hi);
                __half2* ll = reinterpret_cast<__half2*>(&lo);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float a0 = fminf(fmaxf(v[2 * i] * scale, -65504.f), 65504.f);
                    const float a1 = fminf(fmaxf(v[2 * i + 1] * scale, -65504.f), 65504.f);
                    hh[i] = __floats2half2_rn(a0, a1);
                    const float2 f = __half22float2(hh[i]);
                    ll[i] = __floats2half2_rn(a0 - f.x, a1 - f.y);
                }
            } else {
                __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&hi);
                __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&lo);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    hh[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                    const float2 f = __bfloat1622float2(hh[i]);
                    ll[i] = __floats2bfloat162_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
                }
            }
            tile[0][gl * (TP + 1) + pi] = hi;
            tile[1][gl * (TP + 1) + pi] = lo;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int item = tid + 256 * it;
            const int pi = item % TP, gl = item / TP;
            const int g = g0 + gl;
            const int p = p0 + pi;
            if (g < G && p < QA) {
                uint4* dst = reinterpret_cast<uint4*>(planes) + ((size_t)g * QA + p);
                dst[0] = tile[0][gl * (TP + 1) + pi];
                if (npl == 2) dst[(size_t)G * QA] = tile[1][gl * (TP + 1) + pi];
            }
        }
        __syncthreads();
    }
    if (colsum != nullptr) {
        // db[c] = sum over positions of dY[., c] (the conv bias gradient, fused here because this kernel reads dY anyway).
        // Every item of a thread has the same channel group (256 % TG == 0): reduce the lanes of a warp that share it,
        // then the 8 warps through shared memory, then ONE atomic per (block, channel).
#pragma unroll
        for (int off = TG; off < 32; off <<= 1)
#pragma unroll
            for (int i = 0; i < 8; ++i) csum[i] += __shfl_xor_sync(0xffffffffu, csum[i], off);
        const int lane = tid & 31, warp = tid >> 5;
        if (lane < TG) {
#pragma unroll
            for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = csum[i];
        }
        __syncthreads();
        for (int j = tid; j < TG * 8; j += 256) {
            // with TG == 32 a warp's 32 lanes hold 32 different groups, but lane -> group is tid % TG for every warp
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += red[w][j];
            const int c = g0 * 8 + j;
            if (c < C) atomicAdd(colsum + c, t);
        }
    }
}

struct ShPlan {
    int bn, KYG, NP, NPa, stages, a_bytes, b_chunk_bytes, smem_bytes;
};
}  // namespace

int shift_block_n(int cout) { return cout > 128 ? 256 : cout > 64 ? 128 : cout > 32 ? 64 : cout > 16 ? 32 : 16; }

int wgrad_shift_block_n(int cout) { return cout <= 16 ? 16 : cout <= 32 ? 32 : cout <= 64 ? 64 : 128; }

// M-tiles (of 128 positions) per CTA tile of the forward / dgrad kernel for `cout` output channels when the last computed
// position is q_last
int shift_tile_mt(int cout, int npl, long long q_last) {
    const int bn = shift_block_n(cout);
    int mt = tile_mt(bn, npl);
    // (only when 256-position tiles leave a third of the SMs idle: at ~146 tiles the larger tile's weight reuse wins)
    if (bn == 256 && (q_last / (mt * 128) + 1) * ((cout + 255) / 256) < 100) mt = 1;
    return mt;
}

namespace {
// pipeline-stage geometry; K x K filter on a frame of pitch Wp, npl planes, BN-wide weight stage
int make_plan(int K, int Wp, int npl, int bn, ShPlan& p, int mt = 0) {
    p.bn = bn;
    if (mt == 0) mt = tile_mt(bn, npl);
    for (int mode = 0; mode < 2; ++mode) {
        p.KYG = mode == 0 ? K : 1;
        if (mode == 1 && K == 1) break;
        p.NP = mt * 128 + (p.KYG - 1) * Wp + K - 1;
        p.NPa = (p.NP + 7) / 8 * 8;
        p.a_bytes = npl * 2 * p.NPa * 16;
        p.b_chunk_bytes = K * npl * 2 * p.bn * 16;
        const int stage = p.a_bytes + p.KYG * p.b_chunk_bytes;
        const int st = (SMEM_LIMIT - 128 - 256) / stage;
        if (st < 2) continue;
        p.stages = st > MAX_STAGES ? MAX_STAGES : st;
        p.smem_bytes = p.stages * stage + 128 + 256;
        return 1;
    }
    return 0;
}

template <int BN, int NPASS, int MT = tile_mt(BN, NPASS == 3 ? 2 : 1)>
int launch_shift(const ShArgs& args, int smem_bytes, cudaStream_t st) {
    auto kern = conv_shift_tcgen05_kernel<BN, NPASS, MT>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) {
            affgw_set_error("conv_shift: cannot reserve %d bytes of shared memory", SMEM_LIMIT);
            return -2;
        }
        configured = true;
    }
    const int grid = args.total_tiles < 148 ? args.total_tiles : 148;
    kern<<<grid, NUM_THREADS, smem_bytes, st>>>(args);
    AFFGW_LAUNCH_CHECK("conv_shift_tcgen05");
    return 0;
}

}  // namespace

// Is the forward convolution g a position-space convolution?  One rule for forward, dgrad and wgrad so that the operand
// planes of a layer are built once.
int conv_shift_ok(const ConvGeom& g) {
    if (g.stride != 1 || g.stride_w != 1 || g.zi != 1 || g.zi_w != 1 || g.KH != g.KW || g.KH > 7) return 0;
    // 1x1 filters have no tap reuse to exploit; only the thin ones (the 16/32-channel shortcuts of the discriminator at full
    // resolution) come here, because the bulk-copy pipeline beats the per-thread gather of the im2col kernel for them
    if (g.KH == 1 && (g.Cin > 64 || g.Cout > 64 || (long long)g.H * g.W < 1024)) return 0;
    const long long Hp = g.Hv + 2 * g.pad, Wp = g.Wv + 2 * g.pad;
    if (Hp - g.KH + 1 != g.Ho || Wp - g.KW + 1 != g.Wo) return 0;
    // on tiny maps the padding positions (computed, then dropped) cost more than the window saves: im2col kernels
    // ... unless the layer is wide: there the im2col kernel is bound by the L2 -> SM path (its operands are re-read per tap and per
    // output-channel tile) and computing up to 2x the positions on the tensor pipe is still the faster way (4 x 14 maps: 1.7x;
    // measured break-even is between that and the 2.6x of 2 x 7 maps)
    static const int waste_env = [] { const char* e = getenv("AFFGW_SHIFT_WASTE_X10"); return e ? atoi(e) : 0; }();
    const int waste_x10 = waste_env ? waste_env : ((g.Cin >= 128 && g.Cout >= 128) ? 20 : 15);
    if (10 * Hp * Wp > (long long)waste_x10 * g.Ho * g.Wo) return 0;
    if ((long long)g.N * Hp * Wp + 8192 >= (1LL << 31) / 16) return 0;           // 32-bit position arithmetic
    // the forward tile is as wide as Cout allows, the dgrad tile as wide as Cin allows: both windows must fit
    ShPlan p;
    if (!make_plan(g.KH, (int)Wp, 2, shift_block_n(g.Cout), p)) return 0;
    if (!make_plan(g.KH, (int)Wp, 2, shift_block_n(g.Cin), p)) return 0;
    return 1;
}

// frame of the forward convolution g for a tensor with `channels` channels
void conv_shift_frame(const ConvGeom& g, int channels, PosFrame& f) {
    f.N = g.N;
    f.Hp = g.Hv + 2 * g.pad;
    f.Wp = g.Wv + 2 * g.pad;
    f.G = 2 * ((channels + 15) / 16);
    const int span = (g.KH - 1) * (f.Wp + 1);
    f.lead = (span + 7) / 8 * 8;
    const long long Q = (long long)f.N * f.Hp * f.Wp;
    f.QA = f.lead + (Q + MAX_TILE_POS - 1) / MAX_TILE_POS * MAX_TILE_POS + MAX_TILE_POS + span + g.KH + 32;
    f.QA = (f.QA + 31) / 32 * 32;
}

int split_positions(const void* src, int dt, void* planes, const PosFrame& f, int Hs, int Ws, int C, int pitch, int up, int oy0,
                    int ox0, int pad_mode, int pre_act, int passes, float* colsum, int fmt, const float* scale_dev, cudaStream_t st) {
    const int npl = passes == 3 ? 2 : 1;
    if (f.QA >= (1LL << 31) || (long long)f.N * Hs * Ws >= (1LL << 31)) {
        affgw_set_error("split_positions: frame too large for 32-bit position arithmetic");
        return -1;
    }
    const int QA = (int)f.QA;
#define AFFGW_SPLIT(T, TGV)                                                                                               \
    split_positions_kernel<T, TGV><<<dim3((unsigned)min((QA + 1024 / TGV - 1) / (1024 / TGV), 148 * 8), (unsigned)((f.G + TGV - 1) / TGV)), \
                                     256, 0, st>>>((const T*)src, (bf16*)planes, f.N, Hs, Ws, C, pitch, up, oy0, ox0, pad_mode, \
                                                   pre_act, f.Hp, f.Wp, f.G, f.lead, QA, npl, colsum, make_fastdiv(f.Wp), make_fastdiv(f.Hp), \
                                                   fmt, scale_dev)
    if (dt == AFFGW_F32) {
        if (f.G <= 2) AFFGW_SPLIT(float, 2);
        else if (f.G <= 4) AFFGW_SPLIT(float, 4);
        else if (f.G <= 8) AFFGW_SPLIT(float, 8);
        else if (f.G <= 16) AFFGW_SPLIT(float, 16);
        else AFFGW_SPLIT(float, 32);
    } else {
        if (f.G <= 2) AFFGW_SPLIT(bf16, 2);
        else if (f.G <= 4) AFFGW_SPLIT(bf16, 4);
        else if (f.G <= 8) AFFGW_SPLIT(bf16, 8);
        else if (f.G <= 16) AFFGW_SPLIT(bf16, 16);
        else AFFGW_SPLIT(bf16, 32);
    }
#undef AFFGW_SPLIT
    AFFGW_LAUNCH_CHECK("split_positions");
    return 0;
}

long long pack_weight_shift_bytes(int Cout, int Cin, int K, int ipad, int transpose_flip, int passes) {
    const int Od = transpose_flip ? Cin : Cout, Id = transpose_flip ? Cout : Cin;
    if (ipad % 8 != 0 || ipad < Id || (passes != 1 && passes != 3)) return -1;
    const int bn = shift_block_n(Od);
    const long long ntiles = (Od + bn - 1) / bn, CB = (ipad + 15) / 16;
    return ntiles * K * CB * K * (passes == 3 ? 2 : 1) * 2 * bn * 8 * 2;
}

int pack_weight_shift(const float* w, void* out, int Cout, int Cin, int K, int ipad, int transpose_flip, int passes, int fmt,
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
                                                     passes == 3 ? 2 : 1, fmt);
    AFFGW_LAUNCH_CHECK("pack_weight_shift");
    return 0;
}

// One position-space convolution over the planes `x` (frame f): out[(n, yp - oy0, xp - ox0)][co] for the positions whose
// shifted coordinates fall inside [OH) x [OW).
//   forward : q_shift = 0,               (oy0, ox0, OH, OW) = (0, 0, Ho, Wo)
//   dgrad   : q_shift = -(K-1)(Wp+1),    (pad, pad, H, W) for the zero-pad crop or (0, 0, Hp, Wp) for the full frame
int conv_pos_tc(const void* x_planes, const PosFrame& f, const void* w_packed, const float* bias, const void* addend, void* y,
                int K, int q_shift, int oy0, int ox0, int OH, int OW, int Cout, int out_pitch, int post_act, int passes,
                int fmt, const float* alpha_dev, cudaStream_t st) {
    ShPlan p;
    const int npl = passes == 3 ? 2 : 1;
    const int bn0 = shift_block_n(Cout);
    const long long q_last = ((long long)(f.N - 1) * f.Hp + oy0 + OH - 1) * f.Wp + ox0 + OW - 1;
    // tiny maps with wide layers (the 4 x 14 / 2 x 7 blocks of the discriminator): 256-position tiles leave most SMs idle;
    // 128-position tiles double the tile count and double-buffer the 256-column accumulator
    const int mt = shift_tile_mt(Cout, npl, q_last);
    if (!make_plan(K, f.Wp, npl, bn0, p, mt) || -q_shift > f.lead) {
        affgw_set_error("conv_shift: window of a %dx%d filter on a %d-wide frame does not fit", K, K, f.Wp);
        return -1;
    }
    ShArgs a;
    a.x = (const bf16*)x_planes;
    a.w = (const bf16*)w_packed;
    a.bias = bias;
    a.addend = (const float*)addend;
    a.y = (float*)y;
    a.N = f.N; a.Hp = f.Hp; a.Wp = f.Wp; a.K = K;
    a.G = f.G; a.lead = f.lead; a.QA = f.QA;
    a.q_shift = q_shift;
    a.oy0 = oy0; a.ox0 = ox0; a.OH = OH; a.OW = OW;
    a.Cout = Cout; a.out_pitch = out_pitch; a.post_act = post_act;
    a.CB = f.G / 2;
    a.KYG = p.KYG; a.NP = p.NP; a.NPa = p.NPa; a.stages = p.stages;
    a.a_bytes = p.a_bytes; a.b_chunk_bytes = p.b_chunk_bytes;
    a.n_tiles = (Cout + p.bn - 1) / p.bn;
    a.total_tiles = (int)((q_last / (mt * 128) + 1) * a.n_tiles);
    a.vec_ok = (Cout % 16 == 0) && (out_pitch % 4 == 0) && (((uintptr_t)y) % 16 == 0) &&
               (!addend || ((uintptr_t)addend) % 16 == 0) && (!bias || ((uintptr_t)bias) % 16 == 0);
    a.div_wp = make_fastdiv(f.Wp);
    a.div_hp = make_fastdiv(f.Hp);
    a.fmt = fmt;
    a.alpha = fmt ? 1.f / F16_W_SCALE : 1.f;        // the packed fp16 weights carry 2^8
    a.alpha_dev = alpha_dev;
    if (passes == 3) {
        switch (p.bn) {
            case 16: return launch_shift<16, 3>(a, p.smem_bytes, st);
            case 32: return launch_shift<32, 3>(a, p.smem_bytes, st);
            case 64: return launch_shift<64, 3>(a, p.smem_bytes, st);
            case 128: return launch_shift<128, 3>(a, p.smem_bytes, st);
            default: return mt == 1 ? launch_shift<256, 3, 1>(a, p.smem_bytes, st) : launch_shift<256, 3>(a, p.smem_bytes, st);
        }
    }
    switch (p.bn) {
        case 16: return launch_shift<16, 1>(a, p.smem_bytes, st);
        case 32: return launch_shift<32, 1>(a, p.smem_bytes, st);
        case 64: return launch_shift<64, 1>(a, p.smem_bytes, st);
        case 128: return launch_shift<128, 1>(a, p.smem_bytes, st);
        default: return mt == 1 ? launch_shift<256, 1, 1>(a, p.smem_bytes, st) : launch_shift<256, 1>(a, p.smem_bytes, st);
    }
}

// =====================================================================================================================
// Weight gradient:  dW[co][ky][kx][ci] = sum_q V[q + ky*Wp + kx][ci] * dYp[q][co]
//
// A GEMM whose reduction runs over positions.  The planar layout is UMMA's un-swizzled "MN-major" form (8 positions x
// 8 channels per core matrix; LBO = 128 B to the next 8 positions, SBO = plane pitch to the next 8 channels).  One CTA
// owns a (128 input channels) x (BN output channels) x (kernel row ky) block of dW: for every 16 positions it issues one
// MMA per kx whose A descriptor starts kx positions further into the window, so the input is read once per kernel ROW,
// not once per tap.  M is always 128 (absent channel groups stay zero in shared memory), so thin layers cost N/2 clocks
// per MMA.  The position range is split over grid.y; partials are reduced with coalesced fp32 atomics into
// ws[co][tap][ci].
// =====================================================================================================================
namespace {

constexpr int WG_THREADS = 192;              // producer warp, MMA warp, 4 epilogue warps

struct WsArgs {
    const bf16* x; const bf16* dy;           // position planes
    float* ws;                               // [Cout][taps][cs] fp32
    int K, Wp;
    int Gx, Gy, lead;
    long long QA;
    int Cs, Cout;                            // row length of ws / real output channels
    int KP, pitchA16, pitchB16;              // positions per stage; shared-memory plane pitches in 16-byte units
    int rowsA, rowsB;                        // channel groups per TMA box (<= 16, <= BN / 8)
    int KXG, n_kxg;                          // filter columns per CTA (accumulators <= 512 TMEM columns), column groups per row
    int TPM, slots;                          // filter columns packed into the M rows of one MMA (thin layers); A group slots
    int n_co_blocks;
    int chunks_total, chunks_per_split;
    int a_bytes, b_bytes, stages;
    int fmt;                                 // 0 = bf16 planes, 1 = fp16 planes
    const float* alpha_dev;                  // 1 / (scale of the fp16 dY planes), device memory; nullptr = 1
};

template <int BN, int NPASS>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_shift_kernel(const __grid_constant__ WsArgs a, const __grid_constant__ CUtensorMap tmx,
                        const __grid_constant__ CUtensorMap tmdy) {
    constexpr int NPL = NPASS == 3 ? 2 : 1;
    constexpr int GB = BN / 8;
    // thin tiles in split-bf16 mode: x_hi x [dY_hi | dY_lo] is one MMA of width 2 BN (the remainder plane's channel groups follow
    // the leading plane's in shared memory), x_lo x dY_hi a second one; the epilogue adds the two column blocks (see tile_packed)
    constexpr bool PK = tile_packed(BN, NPL);
    constexpr int ACC_W = tile_acc_w(BN, NPL);
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

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int kxg = blockIdx.x % a.n_kxg;
    const int ky = (blockIdx.x / a.n_kxg) % a.K;
    const int cob = (blockIdx.x / (a.n_kxg * a.K)) % a.n_co_blocks;
    const int cib = blockIdx.x / (a.n_kxg * a.K * a.n_co_blocks);
    const int kx0 = kxg * a.KXG;
    const int nkx = min(a.KXG, a.K - kx0);
    const int cbeg = blockIdx.y * a.chunks_per_split;
    const int cend = min(a.chunks_total, cbeg + a.chunks_per_split);
    const int nst = cend - cbeg;
    constexpr int TMEM_COLS = 512;
    // absent channel groups must read as zeros for the whole kernel: clear every stage once
    for (uint32_t o = tid * 16u; o < (uint32_t)(stages * stage_bytes); o += WG_THREADS * 16u)
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(smem_base + o), "r"(0u) : "memory");
    fence_proxy_async();
    if (tid == 32) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        __syncwarp();
        tmem_alloc(tmem_ptr_addr, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
    const uint32_t pitchA = (uint32_t)a.pitchA16 * 16u, pitchB = (uint32_t)a.pitchB16 * 16u;

    if (warp == 0) {
        // ============================== producer: one tiled TMA load per (operand, plane) ==============================
        // tensor maps view the planes as [plane * G + group][position] rows of 16-byte (two uint64) elements; a box is
        // (window positions) x (channel groups) and lands densely: [group][position] x 16 B.  Rows past this CTA's real
        // groups hold other channels' data or TMA zero fill - they only feed accumulator rows / columns nobody reads.
        if (lane == 0) {
            tma_prefetch_desc(&tmx);
            tma_prefetch_desc(&tmdy);
            const uint32_t copies = a.TPM == 1 ? 1u : (uint32_t)nkx;
            const uint32_t tx = (uint32_t)NPL * 16u * (copies * (uint32_t)a.rowsA * (uint32_t)a.pitchA16 + (uint32_t)a.rowsB * (uint32_t)a.pitchB16);
            int s = 0;
            uint32_t ph = 0;
            for (int st = 0; st < nst; ++st) {
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), tx);
                const int q0 = (cbeg + st) * a.KP + a.lead;
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl) {
                    if (a.TPM == 1) {
                        tma_load_2d(a_smem(s) + (uint32_t)(pl * a.slots) * pitchA, &tmx, 2 * (q0 + ky * a.Wp), pl * a.Gx + cib * 16,
                                    full_bar(s));
                    } else {            // thin layer: one pre-shifted copy of the window per filter column, stacked along M
                        for (int j = 0; j < nkx; ++j)
                            tma_load_2d(a_smem(s) + (uint32_t)(pl * a.slots + j * a.Gx) * pitchA, &tmx,
                                        2 * (q0 + ky * a.Wp + kx0 + j), pl * a.Gx, full_bar(s));
                    }
                    tma_load_2d(b_smem(s) + (uint32_t)(pl * GB) * pitchB, &tmdy, 2 * q0, pl * a.Gy + cob * GB, full_bar(s));
                }
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == MMA_WARP) {
        // ============================== MMA issuer ==============================
        {   // every lane runs the loops (warp-uniform), one elected lane issues: see umma_bf16_elect
            const uint32_t idesc = make_idesc(BN, true, a.fmt);
            const uint32_t idesc2 = make_idesc(PK ? 2 * BN : BN, true, a.fmt);
            int s = 0;
            uint32_t ph = 0;
            for (int st = 0; st < nst; ++st) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                // MN-major: LBO = to the next 8 positions (k), SBO = to the next 8 channels (m / n); descriptors differ only
                // in their start address: add (byte offset >> 4) to the low word
                const uint64_t a_base = make_nosw_desc(a_smem(s), 128u, pitchA);
                const uint64_t b_base = make_nosw_desc(b_smem(s), 128u, pitchB);
                const uint32_t a_lo_off = ((uint32_t)a.slots * pitchA) >> 4, b_lo_off = ((uint32_t)GB * pitchB) >> 4;
                // shared window: accumulator j = filter column kx0 + j, A starts j positions further;
                // packed: accumulator j = columns [j*TPM, (j+1)*TPM) stacked along M, A starts at their first copy
                const int nacc = a.TPM == 1 ? nkx : (nkx + a.TPM - 1) / a.TPM;
                const uint32_t a_step = a.TPM == 1 ? 1u : (uint32_t)(a.TPM * a.Gx * a.pitchA16);
                const uint32_t a_first = a.TPM == 1 ? (uint32_t)kx0 : 0u;
#pragma unroll 1
                for (int k = 0; k < a.KP / 16; ++k) {
                    const uint64_t b_hi = b_base + (uint64_t)(k * 16);
                    const uint32_t accum = (uint32_t)((st | k) != 0);
#pragma unroll 1
                    for (int j = 0; j < nacc; ++j) {
                        const uint64_t a_hi = a_base + (uint64_t)((uint32_t)(k * 16) + a_first + (uint32_t)j * a_step);
                        if (PK) {
                            umma_bf16_elect(tmem_base + j * ACC_W, a_hi, b_hi, idesc2, accum);
                            umma_bf16_elect(tmem_base + j * ACC_W, a_hi + a_lo_off, b_hi, idesc, 1u);
                        } else {
                            umma_bf16_elect(tmem_base + j * ACC_W, a_hi, b_hi, idesc, accum);
                            if (NPASS == 3) {
                                umma_bf16_elect(tmem_base + j * ACC_W, a_hi + a_lo_off, b_hi, idesc, 1u);
                                umma_bf16_elect(tmem_base + j * ACC_W, a_hi, b_hi + b_lo_off, idesc, 1u);
                            }
                        }
                    }
                }
                umma_commit_elect(empty_bar(s));
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            umma_commit_elect(tmem_full_bar);
        }
    } else {
        // ============================== epilogue: fp32 reductions into ws[co][tap][ci] ==============================
        const int wq = warp & 3;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int m = wq * 32 + lane;                        // accumulator row
        const int taps = a.K * a.K;
        const float alpha = a.alpha_dev ? __ldg(a.alpha_dev) : 1.f;
        const int nacc = a.TPM == 1 ? nkx : (nkx + a.TPM - 1) / a.TPM;
        if (nst > 0) {
            for (int j = 0; j < nacc; ++j) {
                // which (filter column, input channel) this row holds
                int kx, ci;
                bool rok;
                if (a.TPM == 1) {
                    kx = kx0 + j;
                    ci = cib * 128 + m;
                    rok = ci < a.Cs;
                } else {
                    const int slot = m >> 3, jj = slot / a.Gx;
                    kx = kx0 + j * a.TPM + jj;
                    ci = (slot - jj * a.Gx) * 8 + (m & 7);
                    rok = jj < a.TPM && kx < kx0 + nkx && ci < a.Cs;
                }
                float* wbase = a.ws + (size_t)(ky * a.K + (rok ? kx : 0)) * a.Cs + (rok ? ci : 0);
#pragma unroll 1
                for (int jb = 0; jb < BN / 16; ++jb) {
                    const int nb = cob * BN + jb * 16;
                    if (nb >= a.Cout) break;
                    uint32_t raw[16];
                    const uint32_t tcol = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(j * ACC_W + jb * 16);
                    tmem_ld16(tcol, raw);
                    if (PK) {
                        uint32_t raw2[16];
                        tmem_ld16(tcol + BN, raw2);
#pragma unroll
                        for (int i = 0; i < 16; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) + __uint_as_float(raw2[i]));
                    }
                    if (rok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (nb + i < a.Cout) atomicAdd(wbase + (size_t)(nb + i) * taps * a.Cs, __uint_as_float(raw[i]) * alpha);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// planes [rows = npl * G][QA positions x 16 bytes] as a 2-D tensor of uint64 pairs; box = box_pos positions x box_rows groups
int make_plane_map(CUtensorMap* tm, const void* planes, long long QA, int rows, int box_pos, int box_rows) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) {
        affgw_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return -2;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)QA * 2, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)QA * 16};
    const cuuint32_t box[2] = {(cuuint32_t)box_pos * 2, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(planes), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        affgw_set_error("cuTensorMapEncodeTiled failed (%d): QA %lld rows %d box %d x %d", (int)r, QA, rows, box_pos, box_rows);
        return -2;
    }
    return 0;
}

template <int BN, int NPASS>
int launch_wg_shift(const WsArgs& args, const CUtensorMap& tmx, const CUtensorMap& tmdy, dim3 grid, int smem_bytes, cudaStream_t st) {
    auto kern = conv_wgrad_shift_kernel<BN, NPASS>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) {
            affgw_set_error("conv_wgrad_shift: cannot reserve %d bytes of shared memory", SMEM_LIMIT);
            return -2;
        }
        configured = true;
    }
    kern<<<grid, WG_THREADS, smem_bytes, st>>>(args, tmx, tmdy);
    AFFGW_LAUNCH_CHECK("conv_wgrad_shift");
    return 0;
}

}  // namespace

// x planes: frame fx; dY planes: same frame positions, fy.G groups.
// ws: [Cout][K*K][cs] fp32, zeroed by the caller (cs = row length, >= the real input channels).
int conv_wgrad_pos_tc(const void* x_planes, const PosFrame& fx, const void* dy_planes, const PosFrame& fy, float* ws, int K,
                      int Cout, int cs, int Ho, int Wo, int passes, int fmt, const float* alpha_dev, cudaStream_t st) {
    const int npl = passes == 3 ? 2 : 1;
    const int bn = wgrad_shift_block_n(Cout);
    WsArgs a;
    a.KP = 64;
    // thin layers (<= 64 stored input channels): stack TPM pre-shifted copies of the window along the 128 M rows, one MMA
    // then covers TPM filter columns.  Otherwise one shared window, one accumulator per filter column.
    const bool packed = fx.G <= 8 && K > 1;
    a.TPM = packed ? (K < 16 / fx.G ? K : 16 / fx.G) : 1;
    int acc_max = 512 / tile_acc_w(bn, npl);                    // accumulators that fit in TMEM ...
    if (packed) {                                               // ... and whose window copies fit in ~40 group slots
        const int cap = 24 / (a.TPM * fx.G) > 1 ? 24 / (a.TPM * fx.G) : 1;
        if (acc_max > cap) acc_max = cap;
    }
    const int per_cta = a.TPM * acc_max;                        // filter columns one CTA can own
    a.n_kxg = (K + per_cta - 1) / per_cta;
    a.KXG = (K + a.n_kxg - 1) / a.n_kxg;
    // group slots of one plane: the last accumulator's MMA reads 16 slots from its first copy
    a.slots = packed ? ((a.KXG + a.TPM - 1) / a.TPM - 1) * a.TPM * fx.G + 16 : 16;
    a.rowsA = fx.G < 16 ? fx.G : 16;
    a.rowsB = fy.G < bn / 8 ? fy.G : bn / 8;
    // thin layers are bound by the per-stage hand-shake, not by bytes: 128 positions per stage (the TMA box limit of 256
    // uint64 elements) when three such stages fit
    if (packed && npl * (a.slots + bn / 8) * 128 * 16 * 3 <= SMEM_LIMIT - 1024) a.KP = 128;
    // Group pitches in 16-byte units; TMA boxes are dense (pitch = box width) and land 128-byte aligned.  An MN-major MMA reads,
    // per position, one 16-byte chunk from each of its 16 (A) / BN / 8 (B) group slots, so the pitch decides which banks those
    // chunks start in: the shared window of the wide layers takes the smallest ODD width that holds it (consecutive slots 16
    // bytes apart: vgg 256->256 weight gradient 1297 -> 1404 TFLOP/s against the even 66).  Packed (thin) tiles hold several
    // boxes per stage whose starts must stay 128-byte aligned: multiples of 4.
    a.pitchA16 = a.KP == 128 ? 128 : (packed ? a.KP + 4 : ((a.KP + K - 1) | 1));
    a.pitchB16 = a.KP == 128 ? 128 : a.KP + (bn >= 64 ? 1 : bn == 32 ? 2 : 4);  // planes start 128-byte aligned (TMA destination)
    a.a_bytes = npl * a.slots * a.pitchA16 * 16;
    a.b_bytes = npl * (bn / 8) * a.pitchB16 * 16;
    const int stage = a.a_bytes + a.b_bytes;
    const int stg = (SMEM_LIMIT - 128 - 256) / stage;
    if (stg < 2 || K > 7 || fx.QA != fy.QA || fx.lead != fy.lead || fx.Wp != fy.Wp || fx.QA >= (1LL << 30)) {
        affgw_set_error("conv_wgrad_shift: unsupported configuration");
        return -1;
    }
    CUtensorMap tmx, tmdy;
    if (int rc = make_plane_map(&tmx, x_planes, fx.QA, npl * fx.G, a.pitchA16, a.rowsA)) return rc;
    if (int rc = make_plane_map(&tmdy, dy_planes, fy.QA, npl * fy.G, a.pitchB16, a.rowsB)) return rc;
    a.stages = stg > MAX_STAGES ? MAX_STAGES : stg;
    const int smem_bytes = a.stages * stage + 128 + 256;
    a.x = (const bf16*)x_planes; a.dy = (const bf16*)dy_planes; a.ws = ws;
    a.K = K; a.Wp = fx.Wp; a.Gx = fx.G; a.Gy = fy.G; a.lead = fx.lead; a.QA = fx.QA;
    a.Cs = cs; a.Cout = Cout;
    a.fmt = fmt; a.alpha_dev = alpha_dev;
    const int n_ci = (fx.G + 15) / 16;
    a.n_co_blocks = (Cout + bn - 1) / bn;
    const long long q_last = ((long long)(fx.N - 1) * fx.Hp + Ho - 1) * fx.Wp + Wo - 1;
    a.chunks_total = (int)(q_last / a.KP + 1);
    const int tiles = n_ci * a.n_co_blocks * K * a.n_kxg;
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
        switch (bn) {
            case 16: return launch_wg_shift<16, 3>(a, tmx, tmdy, grid, smem_bytes, st);
            case 32: return launch_wg_shift<32, 3>(a, tmx, tmdy, grid, smem_bytes, st);
            case 64: return launch_wg_shift<64, 3>(a, tmx, tmdy, grid, smem_bytes, st);
            default: return launch_wg_shift<128, 3>(a, tmx, tmdy, grid, smem_bytes, st);
        }
    }
    switch (bn) {
        case 16: return launch_wg_shift<16, 1>(a, tmx, tmdy, grid, smem_bytes, st);
        case 32: return launch_wg_shift<32, 1>(a, tmx, tmdy, grid, smem_bytes, st);
        case 64: return launch_wg_shift<64, 1>(a, tmx, tmdy, grid, smem_bytes, st);
        default: return launch_wg_shift<128, 1>(a, tmx, tmdy, grid, smem_bytes, st);
    }
}

// =====================================================================================================================
// Per-tensor power-of-two scale of an fp16 dY operand (gradients span many orders of magnitude below fp16's range):
// scale = 2^floor(log2(2^14 / amax)), so amax * scale lies in [2^13, 2^14); out[0] = scale, out[1] = 1 / scale.
// =====================================================================================================================
namespace {
// ws[0] = running max of |x| as float bits (non-negative floats order like unsigned ints), ws[1] = blocks finished.  Both are zero
// on entry and are left zero by the last block, which also writes the scale: ONE launch per tensor, no memset.
__global__ void __launch_bounds__(256) amax_scale_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ ws,
                                                         float* __restrict__ out) {
    float m = 0.f;
    const long long n4 = n / 4, stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {          // four 16-byte loads in flight per thread
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(x) + i + stride);
        const float4 c = __ldg(reinterpret_cast<const float4*>(x) + i + 2 * stride);
        const float4 d = __ldg(reinterpret_cast<const float4*>(x) + i + 3 * stride);
        m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                           fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
        m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))),
                           fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w)))));
    }
    for (; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    if (blockIdx.x == 0) for (long long j = n4 * 4 + threadIdx.x; j < n; j += blockDim.x) m = fmaxf(m, fabsf(x[j]));
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f && m < INFINITY) atomicMax(ws, __float_as_uint(m));     // NaN / inf never enter
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(ws + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        const float amax = __uint_as_float(atomicExch(ws, 0u));
        ws[1] = 0u;
        float s = 1.f;
        if (amax > 0.f) {
            int e;
            frexpf(amax, &e);                   // amax = f * 2^e, f in [0.5, 1)  ->  amax * 2^(14 - e) in [2^13, 2^14)
            s = ldexpf(1.f, max(-120, min(120, 14 - e)));
        }
        out[0] = s;
        out[1] = 1.f / s;
    }
}
}  // namespace

int amax_scale(const float* x, long long n, float* out2, unsigned* ws, cudaStream_t st) {
    // ~4 float4 per thread; at most 8 blocks per SM (2048 threads): enough bytes in flight for the HBM latency
    const int blocks = (int)max(1LL, min((long long)148 * 8, (n / 16 + 255) / 256));
    amax_scale_kernel<<<blocks, 256, 0, st>>>(x, n, ws, out2);
    AFFGW_LAUNCH_CHECK("amax_scale");
    return 0;
}
