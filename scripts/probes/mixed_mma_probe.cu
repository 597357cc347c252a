// Probe: does tcgen05.mma kind::f16 accept A = fp16 with B = bf16 (and the reverse) in ONE instruction?
// (the instruction descriptor has separate a_format / b_format fields; wanted for wgrad = x(fp16)^T . dY(bf16)).
// One 128 x 64 x 16 MMA, un-swizzled K-major operands, result read back from TMEM and compared with the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I affganwriting_b200/csrc -o gpurun_out/mixed_mma_probe scripts/probes/mixed_mma_probe.cu
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tc_ptx.cuh"
void affgw_set_error(const char*, ...) {}
void affgw_count_launch() {}
using namespace tcptx;

__device__ __forceinline__ uint64_t nosw_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    return d;
}

// a_fmt / b_fmt: 0 = f16, 1 = bf16
__global__ void probe(const uint16_t* A, const uint16_t* B, float* D, int a_fmt, int b_fmt) {
    __shared__ __align__(128) uint16_t sa[128 * 16];
    __shared__ __align__(128) uint16_t sb[64 * 16];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // K-major no-swizzle: core matrix = 8 rows x 16 bytes (8 k elements), stored contiguously (128 B);
    // layout [k-group (2)][row-group][8 rows][8 elems]: LBO = bytes between k-groups, SBO = 128 between row groups
    for (int i = tid; i < 128 * 16; i += blockDim.x) {
        const int r = i / 16, k = i % 16;
        sa[(k / 8) * (128 * 8) + (r / 8) * 64 + (r % 8) * 8 + (k % 8)] = A[i];
    }
    for (int i = tid; i < 64 * 16; i += blockDim.x) {
        const int r = i / 16, k = i % 16;
        sb[(k / 8) * (64 * 8) + (r / 8) * 64 + (r % 8) * 8 + (k % 8)] = B[i];
    }
    fence_proxy_async();
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) { __syncwarp(); tmem_alloc(smem_u32(&tmem_slot), 64); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t ad = nosw_desc(smem_u32(sa), 128 * 16, 128);
        const uint64_t bd = nosw_desc(smem_u32(sb), 64 * 16, 128);
        umma_bf16_elect(tmem, ad, bd, idesc, 0u);
        umma_commit_elect(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    if (warp < 4) {
        for (int j = 0; j < 4; ++j) {
            uint32_t raw[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + j * 16, raw);
            for (int i = 0; i < 16; ++i) D[(warp * 32 + lane) * 64 + j * 16 + i] = __uint_as_float(raw[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFF + ((u >> 16) & 1); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t f2h(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float h2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

int main() {
    int bad = 0;
    for (int cfg = 0; cfg < 4; ++cfg) {
        const int af = cfg & 1, bfm = cfg >> 1;
        std::vector<uint16_t> A(128 * 16), B(64 * 16);
        std::vector<float> Af(128 * 16), Bf(64 * 16), ref(128 * 64, 0.f), out(128 * 64);
        srand(7);
        for (int i = 0; i < 128 * 16; ++i) { float v = (rand() % 2001 - 1000) / 700.0f; A[i] = af ? f2bf(v) : f2h(v); Af[i] = af ? bf2f(A[i]) : h2f(A[i]); }
        for (int i = 0; i < 64 * 16; ++i) { float v = (rand() % 2001 - 1000) / 900.0f; B[i] = bfm ? f2bf(v) : f2h(v); Bf[i] = bfm ? bf2f(B[i]) : h2f(B[i]); }
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { double s = 0; for (int k = 0; k < 16; ++k) s += (double)Af[m * 16 + k] * Bf[n * 16 + k]; ref[m * 64 + n] = (float)s; }
        uint16_t *dA, *dB; float* dD;
        cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, out.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
        probe<<<1, 128>>>(dA, dB, dD, af, bfm);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0, mag = 0;
        for (int i = 0; i < 128 * 64; ++i) { err = fmax(err, fabs(out[i] - ref[i])); mag = fmax(mag, fabs(ref[i])); }
        printf("A=%s B=%s : cuda %s, max abs err %.3e (max |ref| %.3f) -> %s\n", af ? "bf16" : "f16", bfm ? "bf16" : "f16",
               cudaGetErrorString(e), err, mag, (e == cudaSuccess && err < 1e-4 * mag) ? "OK" : "MISMATCH");
        if (!(e == cudaSuccess && err < 1e-4 * mag)) ++bad;
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    return bad;
}
