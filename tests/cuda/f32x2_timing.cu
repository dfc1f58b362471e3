// Issue-slot micro-timing of the packed fp32 instructions of sm_100 (FFMA2 / FMUL2) on a B200: does a packed
// instruction free an issue slot for another pipe (LOP3 / MUFU), or does it hold the dispatch port for two cycles?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o f32x2_timing.bin f32x2_timing.cu && ./f32x2_timing.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

// MODE 0: 16 scalar FFMA                      1: 8 FFMA2 (same flops)
//      2: 16 FFMA + 16 LOP3                   3: 8 FFMA2 + 16 LOP3
//      4: 16 FFMA + 8 MUFU.RCP                5: 8 FFMA2 + 8 MUFU.RCP
//      6: 16 LOP3 only                        7: 8 MUFU only
template <int MODE>
__global__ void k(float* out, long long* cyc, float seed, uint32_t iseed) {
    float a[16];
    uint32_t l[16];
    float m[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = seed + i + threadIdx.x; l[i] = iseed * (i + 1) + threadIdx.x; }
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = seed * (i + 2) + 1.5f;
    const float c1 = seed * 0.999f, c2 = seed * 1e-3f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        if (MODE == 0 || MODE == 2 || MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], c1, c2);
        }
        if (MODE == 1 || MODE == 3 || MODE == 5) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 r = __ffma2_rn(make_float2(a[2 * i], a[2 * i + 1]), make_float2(c1, c1), make_float2(c2, c2));
                a[2 * i] = r.x;
                a[2 * i + 1] = r.y;
            }
        }
        if (MODE == 2 || MODE == 3 || MODE == 6) {
#pragma unroll
            for (int i = 0; i < 16; ++i) l[i] = (l[i] & 0x55aa55aau) ^ (l[(i + 1) & 15] | iseed);
        }
        if (MODE == 4 || MODE == 5 || MODE == 7) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { s += a[i]; x ^= l[i]; }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, float* out, long long* cyc) {
    k<MODE><<<148, threads>>>(out, cyc, 1.0001f, 0x9e3779b9u);
    k<MODE><<<148, threads>>>(out, cyc, 1.0001f, 0x9e3779b9u);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s warps/scheduler=%d  cycles/iteration/warp-set = %.2f\n", name, threads / 128, (double)h / kIters);
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, sizeof(long long));
    for (int threads : {128, 384, 512}) {
        run<0>("16 FFMA", threads, out, cyc);
        run<1>("8 FFMA2", threads, out, cyc);
        run<6>("16 LOP3", threads, out, cyc);
        run<7>("8 MUFU.RCP", threads, out, cyc);
        run<2>("16 FFMA + 16 LOP3", threads, out, cyc);
        run<3>("8 FFMA2 + 16 LOP3", threads, out, cyc);
        run<4>("16 FFMA + 8 MUFU", threads, out, cyc);
        run<5>("8 FFMA2 + 8 MUFU", threads, out, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
