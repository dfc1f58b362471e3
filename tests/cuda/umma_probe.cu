// Hardware probe for the tcgen05 building blocks used by libnadm_b200 (run on a B200; see tests/test_gpu_umma_probe.py).
// Each case builds operand images on the host with the layout formulas of csrc/nadm_tc.cuh, runs one CTA that issues
// the MMAs, reads the accumulator back from tensor memory and compares with a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../../neural_admixture_b200/csrc/nadm_tc.cuh"

using namespace nadm::tc;

struct Params {
    int kind;          // 0 = i8, 1 = tf32, 2 = f16(bf16)
    int a_from_tmem;   // TS form
    uint32_t idesc;
    int nk;            // number of MMA instructions (K steps)
    uint32_t a_lbo, a_sbo, a_step;   // bytes (a_step: start-address advance per K step; TS: columns per K step)
    uint32_t b_lbo, b_sbo, b_step;
    int a_bytes, b_bytes;            // image sizes
    int a_tmem_cols;                 // TS: 32-bit columns of the A image ([128][cols])
    int d_cols;
};

__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* a_img, const uint8_t* b_img, const uint32_t* a_tm,
                                                    uint32_t* d_out, Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    uint8_t* sa = smem;
    uint8_t* sb = smem + ((p.a_bytes + 127) / 128) * 128;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < p.a_bytes; i += 128) sa[i] = a_img[i];
    for (int i = tid; i < p.b_bytes; i += 128) sb[i] = b_img[i];
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init_fence();
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t a_col0 = 256;  // A image lives in columns [256, 256 + a_tmem_cols)
    if (p.a_from_tmem) {
        for (int c0 = 0; c0 < p.a_tmem_cols; c0 += 16) {
            uint32_t v[16];
            for (int j = 0; j < 16; ++j) v[j] = (c0 + j < p.a_tmem_cols) ? a_tm[tid * p.a_tmem_cols + c0 + j] : 0u;
            tmem_st16(lane_base + a_col0 + c0, v);
        }
        tmem_wait_st();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
    }
    if (tid == 0) {
        for (int k = 0; k < p.nk; ++k) {
            const uint64_t bd = smem_desc(smem_u32(sb) + k * p.b_step, p.b_lbo, p.b_sbo);
            if (p.a_from_tmem) {
                const uint32_t at = tbase + a_col0 + k * p.a_step;
                if (p.kind == 0) mma_i8_ts(tbase, at, bd, p.idesc, k > 0);
                else if (p.kind == 1) mma_tf32_ts(tbase, at, bd, p.idesc, k > 0);
                else mma_f16_ts(tbase, at, bd, p.idesc, k > 0);
            } else {
                const uint64_t ad = smem_desc(smem_u32(sa) + k * p.a_step, p.a_lbo, p.a_sbo);
                if (p.kind == 0) mma_i8_ss(tbase, ad, bd, p.idesc, k > 0);
                else if (p.kind == 1) mma_tf32_ss(tbase, ad, bd, p.idesc, k > 0);
                else mma_f16_ss(tbase, ad, bd, p.idesc, k > 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < p.d_cols; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(lane_base + c0, v);
        tmem_wait_ld();
        for (int j = 0; j < 8; ++j) d_out[tid * p.d_cols + c0 + j] = v[j];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tbase);
}

// Two threads of DIFFERENT warps issue interleaved K steps into the SAME accumulator (zero-initialised with tcgen05.st,
// every MMA accumulates).  Integer adds commute, so any ordering is fine; what the probe looks for is a lost update
// when read-modify-writes of one accumulator come from two instruction streams.
__global__ void __launch_bounds__(128) probe2_kernel(const uint8_t* a_img, const uint8_t* b_img, uint32_t* d_out, Params p,
                                                     int reps) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    uint8_t* sa = smem;
    uint8_t* sb = smem + ((p.a_bytes + 127) / 128) * 128;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < p.a_bytes; i += 128) sa[i] = a_img[i];
    for (int i = tid; i < p.b_bytes; i += 128) sb[i] = b_img[i];
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) {
        mbar_init(&bar, 2);
        mbar_init_fence();
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    {
        uint32_t z[16];
        for (int j = 0; j < 16; ++j) z[j] = 0u;
        for (int c0 = 0; c0 < p.d_cols; c0 += 16) tmem_st16(lane_base + c0, z);
        tmem_wait_st();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (tid == 0 || tid == 32) {
        const int me = tid >> 5;
        for (int rep = 0; rep < reps; ++rep)
            for (int k = me; k < p.nk; k += 2) {
                const uint64_t ad = smem_desc(smem_u32(sa) + k * p.a_step, p.a_lbo, p.a_sbo);
                const uint64_t bd = smem_desc(smem_u32(sb) + k * p.b_step, p.b_lbo, p.b_sbo);
                mma_i8_ss(tbase, ad, bd, p.idesc, 1u);
            }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < p.d_cols; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(lane_base + c0, v);
        tmem_wait_ld();
        for (int j = 0; j < 8; ++j) d_out[tid * p.d_cols + c0 + j] = v[j];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tbase);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

static std::vector<uint32_t> run(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b,
                                 const std::vector<uint32_t>& atm, Params p) {
    uint8_t *da, *db; uint32_t *dt, *dd;
    CK(cudaMalloc(&da, a.size() + 16)); CK(cudaMalloc(&db, b.size() + 16));
    CK(cudaMalloc(&dt, atm.size() * 4 + 16)); CK(cudaMalloc(&dd, 128 * p.d_cols * 4));
    if (!a.empty()) CK(cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice));
    if (!atm.empty()) CK(cudaMemcpy(dt, atm.data(), atm.size() * 4, cudaMemcpyHostToDevice));
    p.a_bytes = (int)a.size(); p.b_bytes = (int)b.size();
    size_t sm = ((a.size() + 127) / 128) * 128 + b.size() + 256;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    probe_kernel<<<1, 128, sm>>>(da, db, dt, dd, p);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> out(128 * p.d_cols);
    CK(cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dt); cudaFree(dd);
    return out;
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
static uint16_t bf16_bits(float x) { __nv_bfloat16 h = __float2bfloat16(x); uint16_t u; memcpy(&u, &h, 2); return u; }
static uint32_t f2u(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static float u2f(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }

// dlane(r, M): TMEM lane holding accumulator row r
static int dlane(int r, int M) { return M == 128 ? r : (r % 16) + 32 * (r / 16); }

int main() {
    int fails = 0;
    srand(1);
    // ---------------------------------------------------------------- case 1/2: i8, A K-major or MN-major, B MN-major
    for (int a_mn = 0; a_mn < 2; ++a_mn) {
        const int M = 128, N = 32, K = 96;  // 3 instructions of K = 32
        std::vector<int> A(M * K), B(N * K);
        for (auto& x : A) x = rand() % 3;
        for (auto& x : B) x = rand() % 256 - 128;
        // the genotype tile: written K-major as [row r][k] with SBO_t (8-row groups) and LBO_t (16-byte k chunks)
        const uint32_t LBO_t = 128, SBO_t = 128 * (a_mn ? M / 16 : K / 16);
        // when a_mn: the tile is stored as [r' = k index (b)][k' = m index] and read transposed
        std::vector<uint8_t> a(M * K, 0), b(N * K, 0);
        Params p{};
        p.kind = 0; p.nk = K / 32; p.d_cols = N;
        if (!a_mn) {
            for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k)
                a[(r % 8) * 16 + (r / 8) * SBO_t + (k / 16) * LBO_t + (k % 16)] = (uint8_t)A[r * K + k];
            p.a_lbo = LBO_t; p.a_sbo = SBO_t; p.a_step = 2 * LBO_t;   // 32 k = two 16-byte chunks
        } else {
            // tile rows = k (batch rows b), tile columns = M (SNPs m): addr(b, m) = (b%8)*16 + (b/8)*SBO_t + (m/16)*LBO_t + m%16
            for (int i = 0; i < M; ++i) for (int j = 0; j < K; ++j)
                a[(j % 8) * 16 + (j / 8) * SBO_t + (i / 16) * LBO_t + (i % 16)] = (uint8_t)A[i * K + j];
            p.a_sbo = LBO_t; p.a_lbo = SBO_t; p.a_step = 4 * SBO_t;   // 32 k = four 8-row groups
        }
        // B: digits, MN-major: addr(n, k) = n%16 + (n/16)*128 + (k%8)*16 + (k/8)*256
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k)
            b[(n % 16) + (n / 16) * 128 + (k % 8) * 16 + (k / 8) * 256] = (uint8_t)(int8_t)B[n * K + k];
        p.b_sbo = 128; p.b_lbo = 256; p.b_step = 4 * 256;
        p.idesc = instr_desc(kAccS32, kFmtU8, kFmtS8, a_mn != 0, true, M, N);
        auto d = run(a, b, {}, p);
        long bad = 0;
        for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
            int ref = 0;
            for (int k = 0; k < K; ++k) ref += A[r * K + k] * B[n * K + k];
            if ((int)d[r * N + n] != ref) ++bad;
        }
        printf("case i8 A %s-major, B MN-major: %s (%ld mismatches)\n", a_mn ? "MN" : "K", bad ? "FAIL" : "PASS", bad);
        fails += bad != 0;
    }
    // ---------------------------------------------------------------- case 3: tf32 SS, both K-major, hi/lo split accuracy
    {
        const int M = 128, N = 64, K = 8;
        std::vector<float> Q(M * K), P(N * K);
        for (auto& x : Q) x = (float)rand() / RAND_MAX;
        for (auto& x : P) x = (float)rand() / RAND_MAX;
        // chunks per row: [hi_a, hi_b, lo_a, lo_b] (a: k 0-3, b: k 4-7); row r at (r%8)*16 + (r/8)*512, chunk c at c*128
        auto img = [&](const std::vector<float>& X, int R) {
            std::vector<uint8_t> o(R * 64, 0);
            for (int r = 0; r < R; ++r) for (int k = 0; k < K; ++k) {
                float hi = trunc_tf32(X[r * K + k]), lo = X[r * K + k] - hi;
                uint32_t off = (r % 8) * 16 + (r / 8) * 512 + (k % 4) * 4;
                memcpy(&o[off + (k / 4) * 128], &hi, 4);
                memcpy(&o[off + (2 + k / 4) * 128], &lo, 4);
            }
            return o;
        };
        auto a = img(Q, M), b = img(P, N);
        Params p{};
        p.kind = 1; p.d_cols = N; p.a_lbo = p.b_lbo = 128; p.a_sbo = p.b_sbo = 512;
        p.idesc = instr_desc(kAccF32, kFmtTF32, kFmtTF32, false, false, M, N);
        // step 0: Qhi.Phi   step 1: Qlo.Phi   step 2: Qhi.Plo   -> emulate with nk = 1 runs and different bases:
        // (the probe kernel advances A and B together, so run three single-step launches and sum on the host)
        double err1 = 0, err3 = 0, nrm = 0;
        std::vector<double> acc(M * N, 0.0);
        const int aoff[3] = {0, 256, 0}, boff[3] = {0, 0, 256};
        std::vector<std::vector<uint32_t>> parts;
        for (int s = 0; s < 3; ++s) {
            std::vector<uint8_t> a2(a.begin() + aoff[s], a.end()), b2(b.begin() + boff[s], b.end());
            p.nk = 1;
            parts.push_back(run(a2, b2, {}, p));
        }
        for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)Q[r * K + k] * P[n * K + k];
            double one = u2f(parts[0][r * N + n]);
            double three = one + u2f(parts[1][r * N + n]) + u2f(parts[2][r * N + n]);
            err1 += (one - ref) * (one - ref); err3 += (three - ref) * (three - ref); nrm += ref * ref;
        }
        printf("case tf32 SS: rel err single %.3e, 3-term split %.3e -> %s\n", sqrt(err1 / nrm), sqrt(err3 / nrm),
               sqrt(err3 / nrm) < 2e-6 ? "PASS" : "FAIL");
        fails += !(sqrt(err3 / nrm) < 2e-6);
    }
    // ---------------------------------------------------------------- case 4: bf16, A from TMEM (TS), B MN-major
    {
        const int M = 128, N = 32, K = 64;   // 4 instructions of K = 16; A image: 32 columns of packed bf16 pairs
        std::vector<float> A(M * K), B(N * K);
        for (auto& x : A) x = bf16_round((float)rand() / RAND_MAX - 0.5f);
        for (auto& x : B) x = bf16_round((float)rand() / RAND_MAX - 0.5f);
        std::vector<uint32_t> atm(M * (K / 2));
        for (int r = 0; r < M; ++r) for (int k = 0; k < K; k += 2)
            atm[r * (K / 2) + k / 2] = (uint32_t)bf16_bits(A[r * K + k]) | ((uint32_t)bf16_bits(A[r * K + k + 1]) << 16);
        // B MN-major, 16-bit: T = 8: addr(n, k) = (n%8)*2 + (n/8)*SBO + (k%8)*16 + (k/8)*LBO ; SBO = 128, LBO = 128*(N/8)
        std::vector<uint8_t> b(N * K * 2, 0);
        const uint32_t SBO = 128, LBO = 128 * (N / 8);
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
            uint16_t h = bf16_bits(B[n * K + k]);
            memcpy(&b[(n % 8) * 2 + (n / 8) * SBO + (k % 8) * 16 + (k / 8) * LBO], &h, 2);
        }
        Params p{};
        p.kind = 2; p.a_from_tmem = 1; p.nk = K / 16; p.d_cols = N; p.a_tmem_cols = K / 2; p.a_step = 8;
        p.b_sbo = SBO; p.b_lbo = LBO; p.b_step = 2 * LBO;
        p.idesc = instr_desc(kAccF32, kFmtBF16, kFmtBF16, false, true, M, N);
        auto d = run({}, b, atm, p);
        double err = 0, nrm = 0;
        for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[r * K + k] * B[n * K + k];
            double got = u2f(d[r * N + n]);
            err += (got - ref) * (got - ref); nrm += ref * ref;
        }
        printf("case bf16 TS (A in TMEM), B MN-major: rel err %.3e -> %s\n", sqrt(err / nrm), sqrt(err / nrm) < 1e-5 ? "PASS" : "FAIL");
        fails += !(sqrt(err / nrm) < 1e-5);
    }
    // ---------------------------------------------------------------- case 5: bf16 SS, A MN-major with M = 64, N = 24, B MN-major
    {
        const int M = 64, N = 24, K = 128;   // 8 instructions
        std::vector<float> A(M * K), B(N * K);
        for (auto& x : A) x = bf16_round((float)rand() / RAND_MAX - 0.5f);
        for (auto& x : B) x = bf16_round((float)rand() / RAND_MAX - 0.5f);
        // G tile written K-major as [row b (=K index j)][col m (=M index i)], 16-bit: addr(b, m) = (b%8)*16 + (b/8)*SBO_t + (m/8)*LBO_t + (m%8)*2
        const uint32_t LBO_t = 128, SBO_t = 128 * (M / 8);
        std::vector<uint8_t> a(M * K * 2, 0), b(N * K * 2, 0);
        for (int i = 0; i < M; ++i) for (int j = 0; j < K; ++j) {
            uint16_t h = bf16_bits(A[i * K + j]);
            memcpy(&a[(j % 8) * 16 + (j / 8) * SBO_t + (i / 8) * LBO_t + (i % 8) * 2], &h, 2);
        }
        const uint32_t SBOb = 128, LBOb = 128 * (N / 8);
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
            uint16_t h = bf16_bits(B[n * K + k]);
            memcpy(&b[(n % 8) * 2 + (n / 8) * SBOb + (k % 8) * 16 + (k / 8) * LBOb], &h, 2);
        }
        Params p{};
        p.kind = 2; p.nk = K / 16; p.d_cols = 32;
        p.a_sbo = LBO_t; p.a_lbo = SBO_t; p.a_step = 2 * SBO_t;
        p.b_sbo = SBOb; p.b_lbo = LBOb; p.b_step = 2 * LBOb;
        p.idesc = instr_desc(kAccF32, kFmtBF16, kFmtBF16, true, true, M, N);
        auto d = run(a, b, {}, p);
        double err = 0, nrm = 0;
        for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[r * K + k] * B[n * K + k];
            double got = u2f(d[dlane(r, M) * 32 + n]);
            err += (got - ref) * (got - ref); nrm += ref * ref;
        }
        printf("case bf16 SS, A MN-major M=64 N=24, B MN-major: rel err %.3e -> %s\n", sqrt(err / nrm), sqrt(err / nrm) < 1e-5 ? "PASS" : "FAIL");
        fails += !(sqrt(err / nrm) < 1e-5);
    }
    // ---------------------------------------------------------------- case 6: i8, A (u8) from TMEM (TS), B MN-major
    // (not used by the library yet: the layout a forward encoder with its genotype operand in tensor memory would need)
    {
        const int M = 128, N = 32, K = 96;   // 3 instructions of K = 32; A image: 4 consecutive k bytes per 32-bit column
        std::vector<int> A(M * K), B(N * K);
        for (auto& x : A) x = rand() % 3;
        for (auto& x : B) x = rand() % 256 - 128;
        std::vector<uint32_t> atm(M * (K / 4), 0);
        for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k)
            atm[r * (K / 4) + k / 4] |= (uint32_t)A[r * K + k] << (8 * (k % 4));
        std::vector<uint8_t> b(N * K, 0);
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k)
            b[(n % 16) + (n / 16) * 128 + (k % 8) * 16 + (k / 8) * 256] = (uint8_t)(int8_t)B[n * K + k];
        Params p{};
        p.kind = 0; p.a_from_tmem = 1; p.nk = K / 32; p.d_cols = N; p.a_tmem_cols = K / 4; p.a_step = 8;
        p.b_sbo = 128; p.b_lbo = 256; p.b_step = 4 * 256;
        p.idesc = instr_desc(kAccS32, kFmtU8, kFmtS8, false, true, M, N);
        auto d = run({}, b, atm, p);
        long bad = 0;
        for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
            int ref = 0;
            for (int k = 0; k < K; ++k) ref += A[r * K + k] * B[n * K + k];
            if ((int)d[r * N + n] != ref) ++bad;
        }
        printf("case i8 TS (A u8 in TMEM, 4 k per column), B MN-major: %s (%ld mismatches)\n", bad ? "FAIL" : "PASS", bad);
        fails += bad != 0;
    }
    // ---------------------------------------------------------------- case 7: i8 SS, two issuing threads, one accumulator
    {
        const int M = 128, N = 32, K = 512, REPS = 40;   // 16 K steps x 40 repetitions, alternating between two warps
        std::vector<int> A(M * K), B(N * K);
        for (auto& x : A) x = rand() % 3;
        for (auto& x : B) x = rand() % 256 - 128;
        const uint32_t LBO_t = 128, SBO_t = 128 * (K / 16);
        std::vector<uint8_t> a(M * K, 0), b(N * K, 0);
        for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k)
            a[(r % 8) * 16 + (r / 8) * SBO_t + (k / 16) * LBO_t + (k % 16)] = (uint8_t)A[r * K + k];
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k)
            b[(n % 16) + (n / 16) * 128 + (k % 8) * 16 + (k / 8) * 256] = (uint8_t)(int8_t)B[n * K + k];
        Params p{};
        p.kind = 0; p.nk = K / 32; p.d_cols = N;
        p.a_lbo = LBO_t; p.a_sbo = SBO_t; p.a_step = 2 * LBO_t;
        p.b_sbo = 128; p.b_lbo = 256; p.b_step = 4 * 256;
        p.idesc = instr_desc(kAccS32, kFmtU8, kFmtS8, false, true, M, N);
        p.a_bytes = (int)a.size(); p.b_bytes = (int)b.size();
        uint8_t *da, *db; uint32_t* dd;
        CK(cudaMalloc(&da, a.size())); CK(cudaMalloc(&db, b.size())); CK(cudaMalloc(&dd, 128 * N * 4));
        CK(cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        long bad = 0;
        for (int launch = 0; launch < 20; ++launch) {
            probe2_kernel<<<1, 128, a.size() + b.size() + 256>>>(da, db, dd, p, REPS);
            CK(cudaDeviceSynchronize());
            std::vector<uint32_t> d(128 * N);
            CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
            for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
                int ref = 0;
                for (int k = 0; k < K; ++k) ref += A[r * K + k] * B[n * K + k];
                if ((int)d[r * N + n] != ref * REPS) ++bad;
            }
        }
        cudaFree(da); cudaFree(db); cudaFree(dd);
        // informational (the library does not rely on it yet): does not count as a probe failure
        printf("info i8 SS, two issuer threads into ONE accumulator (20 launches x 640 MMAs): %s (%ld mismatches)\n",
               bad ? "LOST UPDATES" : "exact", bad);
    }
    printf(fails ? "PROBE FAILED (%d)\n" : "PROBE OK\n", fails);
    return fails ? 1 : 0;
}
