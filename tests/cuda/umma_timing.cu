// Micro-timing of small tcgen05.mma shapes on a B200: cycles per instruction for chains of dependent / independent MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_timing umma_timing.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../neural_admixture_b200/csrc/nadm_tc.cuh"
using namespace nadm::tc;

struct Cfg { int kind; int ts; int M, N; int a_mn, b_mn; int reps; int nacc; const char* name; };

template <int KIND, int TS, int NACC>
__global__ void __launch_bounds__(128) timing_kernel(Cfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // finite values
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    if (warp == 0) {
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
        const uint32_t fmt = KIND == 0 ? 0 : kFmtBF16;
        const uint32_t idesc = KIND == 0 ? instr_desc(kAccS32, kFmtU8, kFmtS8, c.a_mn, c.b_mn, c.M, c.N)
                                         : instr_desc(kAccF32, fmt, fmt, c.a_mn, c.b_mn, c.M, c.N);
        const uint64_t ad = smem_desc(sa, c.a_mn ? 1024 : 128, c.a_mn ? 128 : 512);
        const uint64_t bd = smem_desc(sb, c.b_mn ? 512 : 128, c.b_mn ? 128 : 512);
        long long t0 = 0, t1 = 0, t2 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int r0 = 0; r0 < c.reps; r0 += 16) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const uint32_t d = tbase + (r % NACC) * 64;
                    if (TS) mma_f16_ts(d, tbase + 384 + (r & 7) * 8, bd + (uint64_t)((r & 3) * 32), idesc, 1u);
                    else if (KIND == 0) mma_i8_ss(d, ad + (uint64_t)((r & 3) * 16), bd + (uint64_t)((r & 3) * 32), idesc, 1u);
                    else mma_f16_ss(d, ad + (uint64_t)((r & 3) * 128), bd + (uint64_t)((r & 3) * 32), idesc, 1u);
                }
            }
            mma_commit(&bar);
            t1 = clock64();
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        t2 = clock64();
        if (elect_one()) { out[0] = t1 - t0; out[2] = t0; }
        __syncwarp();
        if (tid == 0) out[1] = t2;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tbase);
}

template <int KIND, int TS, int NACC>
static void run(const Cfg& c, long long* d) {
    long long h[3];
    cudaFuncSetAttribute(timing_kernel<KIND, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int it = 0; it < 2; ++it) {
        timing_kernel<KIND, TS, NACC><<<1, 128, 66 * 1024>>>(c, d);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error in %s: %s\n", c.name, cudaGetErrorString(cudaGetLastError())); exit(1); }
    }
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("%-52s issue %.1f cyc/MMA   issue+complete %.1f cyc/MMA\n", c.name, (double)h[0] / c.reps, (double)(h[1] - h[2]) / c.reps);
}

int main() {
    Cfg cfgs[] = {
        {1, 0, 128, 64, 0, 0, 256, 1, "bf16 SS M128 N64 K-major (MMA1), 1 acc"},
        {1, 0, 128, 64, 0, 0, 256, 4, "bf16 SS M128 N64 K-major (MMA1), 4 acc"},
        {1, 1, 128, 32, 0, 1, 256, 1, "bf16 TS M128 N32 (MMA2), 1 acc"},
        {1, 1, 128, 32, 0, 1, 256, 4, "bf16 TS M128 N32 (MMA2), 4 acc"},
        {1, 1, 128, 16, 0, 1, 256, 1, "bf16 TS M128 N16, 1 acc"},
        {1, 0, 64, 24, 1, 1, 256, 1, "bf16 SS M64 N24 A MN-major (MMA3), 1 acc"},
        {1, 0, 64, 24, 1, 1, 256, 4, "bf16 SS M64 N24 A MN-major (MMA3), 4 acc"},
        {1, 0, 128, 32, 1, 1, 256, 1, "bf16 SS M128 N32 A MN-major, 1 acc"},
        {1, 0, 128, 32, 1, 1, 256, 4, "bf16 SS M128 N32 A MN-major, 4 acc"},
        {1, 0, 128, 256, 0, 0, 256, 1, "bf16 SS M128 N256 K-major (GEMM-like), 1 acc"},
        {0, 0, 128, 32, 0, 1, 256, 1, "i8 SS M128 N32 (encoder fwd), 1 acc"},
        {0, 0, 128, 32, 1, 1, 256, 2, "i8 SS M128 N32 A MN-major (encoder bwd), 2 acc"},
    };
    long long* d; cudaMalloc(&d, 64);
    for (auto& c : cfgs) {
        if (c.kind == 0) { if (c.nacc == 1) run<0, 0, 1>(c, d); else run<0, 0, 2>(c, d); }
        else if (c.ts) { if (c.nacc == 1) run<1, 1, 1>(c, d); else run<1, 1, 4>(c, d); }
        else { if (c.nacc == 1) run<1, 0, 1>(c, d); else run<1, 0, 4>(c, d); }
    }
    return 0;
}
