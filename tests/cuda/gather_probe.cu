// Hardware probe: how fast can one B200 gather the rows of a 2-bit packed genotype minibatch into shared memory?
//
// The encoder / decoder kernels stream, per launch, B = 800 randomly chosen sample rows of an N x pitch byte matrix; CTA c
// of a persistent grid owns a contiguous byte-column range of every row.  This probe times ONLY that movement (each
// thread XORs what arrives into a checksum; no other work) for the ways sm_100a offers to get those bytes on chip:
//
//   ldg        plain 128-bit loads into registers, thread = row (what the fused decoder does)
//   cpasync    16-byte cp.async (LDGSTS) into a shared-memory ring, 64 B or 256 B of every row per tile
//   bulk       cp.async.bulk (1-D TMA copy, one per row and tile) + mbarrier complete_tx, 64 .. 896 B per copy
//   gather4    cp.async.bulk.tensor.2d.tile::gather4 (TMA tensor copy of 4 arbitrary rows per instruction)
//
// and for the two tile orders (row block inner = the round-1 kernels; row block outer = consecutive tiles continue the
// same rows).  Every method must produce the same per-CTA checksum.  Output: one line per variant with us / GB/s.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gather_probe.bin gather_probe.cu   (run: ./gather_probe.bin)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Geo {
    const uint8_t* mat;
    long long pitch;
    const int* rows;   // B row numbers
    int B, nblk, ntt;  // ntt tiles of TILE_BYTES per CTA; CTA c owns bytes [c * ntt * TILE_BYTES, ...)
    unsigned* checksum;  // per CTA
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0, polls = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (!ok && ++polls > (1u << 24)) __trap();
    }
}
template <int HINT = 0>   // HINT: L2 prefetch size qualifier (.L2::128B / .L2::256B) of the copy
__device__ __forceinline__ void cp_async16(void* s, const void* g) {
    if (HINT == 256) asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(smem_u32(s)), "l"(g));
    else if (HINT == 128) asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(smem_u32(s)), "l"(g));
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s)), "l"(g));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }
__device__ __forceinline__ uint32_t x4(uint4 v) { return v.x ^ v.y ^ v.z ^ v.w; }

__global__ void fill_kernel(uint8_t* m, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n / 4; i += (long long)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)(i * 2654435761u) ^ (uint32_t)(i >> 15);
        h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        reinterpret_cast<uint32_t*>(m)[i] = h;
    }
}

// tile index -> (blk, tt): BLK_OUTER: i = blk * ntt + tt, else i = tt * nblk + blk
template <bool BLK_OUTER> __device__ __forceinline__ void tile_of(int i, int nblk, int ntt, int& blk, int& tt) {
    if (BLK_OUTER) { blk = i / ntt; tt = i - blk * ntt; } else { tt = i / nblk; blk = i - tt * nblk; }
}

// ---------------------------------------------------------------------------------------------------------------------
// plain loads, thread = row, 64 B of the row per tile, software prefetch of the next tile
// ---------------------------------------------------------------------------------------------------------------------
template <bool BLK_OUTER>
__global__ void __launch_bounds__(512) k_ldg(Geo g) {
    const int tid = threadIdx.x, grp = tid >> 7, r = tid & 127;
    const int ntile = g.nblk * g.ntt;
    const long long c0 = (long long)blockIdx.x * g.ntt * 64;
    uint32_t acc = 0;
    auto load = [&](int i, uint4 (&v)[4]) {
        int blk, tt; tile_of<BLK_OUTER>(i, g.nblk, g.ntt, blk, tt);
        const int b = blk * 128 + r;
        if (b < g.B) {
            const uint4* p = reinterpret_cast<const uint4*>(g.mat + (long long)g.rows[b] * g.pitch + c0 + tt * 64);
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = __ldg(p + q);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = make_uint4(0, 0, 0, 0);
        }
    };
    uint4 cur[4], nxt[4];
    if (grp < ntile) load(grp, cur);
    for (int i = grp; i < ntile; i += 4) {
        if (i + 4 < ntile) load(i + 4, nxt);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc ^= x4(cur[q]);
#pragma unroll
        for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
    }
    for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) atomicXor(&g.checksum[blockIdx.x], acc);
}

// the fused decoder's pattern: thread = row, ONE 16-byte load per (row, 64-SNP unit), row block inner; HINT as above
template <int HINT>
__global__ void __launch_bounds__(384) k_ldg16(Geo g) {
    const int tid = threadIdx.x, grp = tid >> 7, r = tid & 127;
    const int nunit = g.nblk * g.ntt * 4;
    const long long c0 = (long long)blockIdx.x * g.ntt * 64;
    uint32_t acc = 0;
    auto load = [&](int u) {
        const int sub = u / g.nblk, blk = u - sub * g.nblk, b = blk * 128 + r;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (b < g.B) {
            const uint8_t* p = g.mat + (long long)g.rows[b] * g.pitch + c0 + sub * 16;
            if (HINT == 256) asm volatile("ld.global.nc.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
            else if (HINT == 128) asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
            else asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        }
        return v;
    };
    uint4 cur = make_uint4(0, 0, 0, 0);
    if (grp < nunit) cur = load(grp);
    for (int u = grp; u < nunit; u += 3) {
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (u + 3 < nunit) nxt = load(u + 3);
        acc ^= x4(cur);
        cur = nxt;
    }
    for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) atomicXor(&g.checksum[blockIdx.x], acc);
}

// ---------------------------------------------------------------------------------------------------------------------
// cp.async ring: NG groups of 4 warps, each group owns tiles grp, grp + NG, ...; DEPTH tiles in flight per group;
// tile = 128 rows x TB bytes (TB = 64: lane covers 8 rows x 4 pieces per instruction; TB = 256: 2 rows x 16 pieces)
// ---------------------------------------------------------------------------------------------------------------------
template <bool BLK_OUTER, int TB, int NG, int DEPTH, int HINT = 0>
__global__ void __launch_bounds__(NG * 128) k_cpasync(Geo g) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int PIECES = TB / 16;                  // 16-byte pieces per row and tile
    constexpr int PER_THREAD = PIECES;               // 128 rows * PIECES pieces / 128 threads
    constexpr int ROWS_PER_INSTR = 32 / PIECES;      // rows covered by one warp instruction
    const int tid = threadIdx.x, grp = tid >> 7, t = tid & 127, wl = t >> 5, lane = t & 31;
    const int ntile = g.nblk * (g.ntt * 64 / TB);
    const int ntt = g.ntt * 64 / TB;
    const long long c0 = (long long)blockIdx.x * g.ntt * 64;
    uint8_t* ring = smem + (size_t)grp * DEPTH * 128 * TB;
    uint32_t acc = 0;
    const int q = lane % PIECES, r0 = lane / PIECES;
    auto issue = [&](int i, int slot) {
        if (i < ntile) {
            int blk, tt; tile_of<BLK_OUTER>(i, g.nblk, ntt, blk, tt);
            uint8_t* dst = ring + (size_t)slot * 128 * TB + (size_t)(wl * 32 + lane) * 16;   // piece `it` at + it * 2048
#pragma unroll
            for (int it = 0; it < PER_THREAD; ++it) {
                const int b = blk * 128 + wl * 32 + r0 + it * ROWS_PER_INSTR;
                if (b < g.B) cp_async16<HINT>(dst + it * 2048, g.mat + (long long)g.rows[b] * g.pitch + c0 + (long long)tt * TB + q * 16);
                else *reinterpret_cast<uint4*>(dst + it * 2048) = make_uint4(0, 0, 0, 0);
            }
        }
        cp_commit();
    };
    for (int d = 0; d < DEPTH; ++d) issue(grp + d * NG, d);
    int slot = 0;
    for (int i = grp; i < ntile; i += NG) {
        cp_wait<DEPTH - 1>();
        const uint8_t* src = ring + (size_t)slot * 128 * TB + (size_t)(wl * 32 + lane) * 16;
#pragma unroll
        for (int it = 0; it < PER_THREAD; ++it) acc ^= x4(*reinterpret_cast<const uint4*>(src + it * 2048));
        issue(i + DEPTH * NG, slot);
        slot = (slot + 1 == DEPTH) ? 0 : slot + 1;
    }
    for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicXor(&g.checksum[blockIdx.x], acc);
}

// ---------------------------------------------------------------------------------------------------------------------
// cp.async.bulk: one 1-D bulk copy per (row, tile) of TB bytes, completion on the slot's mbarrier
// ---------------------------------------------------------------------------------------------------------------------
template <bool BLK_OUTER, int TB, int NG, int DEPTH>
__global__ void __launch_bounds__(NG * 128) k_bulk(Geo g) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, grp = tid >> 7, t = tid & 127, lane = tid & 31;
    const int ntt = (g.ntt * 64 + TB - 1) / TB;          // (the last tile of a row may be shorter)
    const int ntile = g.nblk * ntt;
    const long long c0 = (long long)blockIdx.x * g.ntt * 64;
    const int span = g.ntt * 64;
    uint8_t* ring = smem + (size_t)grp * DEPTH * 128 * TB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NG * DEPTH * 128 * TB) + grp * DEPTH;
    if (t == 0) for (int d = 0; d < DEPTH; ++d) mbar_init(&bars[d], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t acc = 0;
    auto issue = [&](int i, int slot) {
        if (i >= ntile) return;
        int blk, tt; tile_of<BLK_OUTER>(i, g.nblk, ntt, blk, tt);
        const int len = min(TB, span - tt * TB);
        const int nrow = min(128, g.B - blk * 128);
        if (t == 0) mbar_expect_tx(&bars[slot], (uint32_t)(nrow * len));
        asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");      // expect_tx before any complete_tx
        const int b = blk * 128 + t;
        if (t < nrow) {
            const uint8_t* src = g.mat + (long long)g.rows[b] * g.pitch + c0 + (long long)tt * TB;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(ring + (size_t)slot * 128 * TB + (size_t)t * TB)),
                         "l"(src), "r"(len), "r"(smem_u32(&bars[slot]))
                         : "memory");
        }
    };
    for (int d = 0; d < DEPTH; ++d) issue(grp + d * NG, d);
    int slot = 0, phase = 0;
    for (int i = grp; i < ntile; i += NG) {
        int blk, tt; tile_of<BLK_OUTER>(i, g.nblk, ntt, blk, tt);
        const int len = min(TB, span - tt * TB);
        mbar_wait(&bars[slot], phase);
        if (blk * 128 + t < g.B) {
            const uint8_t* src = ring + (size_t)slot * 128 * TB + (size_t)t * TB;
            for (int o = 0; o < len; o += 16) acc ^= x4(*reinterpret_cast<const uint4*>(src + o));
        }
        asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");      // everyone has read the slot
        issue(i + DEPTH * NG, slot);
        if (++slot == DEPTH) { slot = 0; phase ^= 1; }
    }
    for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicXor(&g.checksum[blockIdx.x], acc);
}

// ---------------------------------------------------------------------------------------------------------------------
// TMA tensor gather4: 2-D tensor map over the matrix (bytes x rows), box TB bytes; one instruction fetches 4 rows
// ---------------------------------------------------------------------------------------------------------------------
template <bool BLK_OUTER, int TB, int NG, int DEPTH>
__global__ void __launch_bounds__(NG * 128) k_gather4(Geo g, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, grp = tid >> 7, t = tid & 127, lane = tid & 31;
    const int ntt = g.ntt * 64 / TB;
    const int ntile = g.nblk * ntt;
    const long long c0 = (long long)blockIdx.x * g.ntt * 64;
    uint8_t* ring = smem + (size_t)grp * DEPTH * 128 * TB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NG * DEPTH * 128 * TB) + grp * DEPTH;
    if (t == 0) for (int d = 0; d < DEPTH; ++d) mbar_init(&bars[d], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t acc = 0;
    auto issue = [&](int i, int slot) {
        if (i >= ntile) return;
        int blk, tt; tile_of<BLK_OUTER>(i, g.nblk, ntt, blk, tt);
        if (t == 0) mbar_expect_tx(&bars[slot], 128u * TB);                 // rows past the batch re-read row 0 of it
        asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
        if (t < 32) {                                                       // 32 instructions x 4 rows
            int r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int b = blk * 128 + t * 4 + k; r[k] = g.rows[b < g.B ? b : 0]; }
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                ::"r"(smem_u32(ring + (size_t)slot * 128 * TB + (size_t)t * 4 * TB)), "l"(&tmap), "r"((int)(c0 + (long long)tt * TB)),
                "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(smem_u32(&bars[slot]))
                : "memory");
        }
    };
    for (int d = 0; d < DEPTH; ++d) issue(grp + d * NG, d);
    int slot = 0, phase = 0;
    for (int i = grp; i < ntile; i += NG) {
        int blk, tt; tile_of<BLK_OUTER>(i, g.nblk, ntt, blk, tt);
        mbar_wait(&bars[slot], phase);
        if (blk * 128 + t < g.B) {
            const uint8_t* src = ring + (size_t)slot * 128 * TB + (size_t)t * TB;
#pragma unroll
            for (int o = 0; o < TB; o += 16) acc ^= x4(*reinterpret_cast<const uint4*>(src + o));
        }
        asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
        issue(i + DEPTH * NG, slot);
        if (++slot == DEPTH) { slot = 0; phase ^= 1; }
    }
    for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicXor(&g.checksum[blockIdx.x], acc);
}

// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static std::vector<unsigned> g_ref;
static int g_nsets = 1, g_launch = 0;      // row sets; run() rotates through them so that no launch finds its rows in L2
static const int* g_rows_base = nullptr;
static Geo* g_geo = nullptr;
static int g_B = 0;
static void next_rows() { g_geo->rows = g_rows_base + (size_t)(g_launch++ % g_nsets) * g_B; }
static unsigned* d_sum;
static int g_ncta;
static double g_bytes;

template <typename F>
static void run(const char* name, F launch) {
    CK(cudaMemset(d_sum, 0, g_ncta * sizeof(unsigned)));
    g_launch = 0;
    next_rows();
    launch();
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-46s FAILED: %s\n", name, cudaGetErrorString(e)); exit(2); }
    std::vector<unsigned> got(g_ncta);
    CK(cudaMemcpy(got.data(), d_sum, g_ncta * sizeof(unsigned), cudaMemcpyDeviceToHost));
    const char* verdict = "checksum = reference";
    if (g_ref.empty()) { g_ref = got; verdict = "(reference)"; }
    else if (got != g_ref) verdict = "CHECKSUM MISMATCH";
    for (int i = 0; i < 3; ++i) { next_rows(); launch(); }
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f, tot = 0.f;
    for (int i = 0; i < 10; ++i) {
        next_rows();
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        best = std::min(best, ms); tot += ms;
    }
    printf("%-46s avg %7.1f us  best %7.1f us  %7.0f GB/s   %s\n", name, tot * 100.f, best * 1000.f, g_bytes / (tot / 10 * 1e-3) / 1e9, verdict);
    fflush(stdout);
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 20000, B = 800, ncta = 148, ntt = argc > 2 ? atoi(argv[2]) : 16;
    const long long pitch = 125056;                       // ceil(500000 / 4 / 128) * 128 bytes per sample row
    g_ncta = ncta;
    g_bytes = (double)ncta * B * ntt * 64;
    uint8_t* mat;
    CK(cudaMalloc(&mat, (size_t)N * pitch));
    fill_kernel<<<1184, 256>>>(mat, (long long)N * pitch);
    std::vector<int> rows(N);
    for (int i = 0; i < N; ++i) rows[i] = i;
    uint64_t s = 88172645463325252ull;
    const int nsets = std::max(1, std::min(16, N / B));    // disjoint sets of B random rows, one per launch in turn
    for (int i = 0; i < nsets * B; ++i) {                  // partial Fisher-Yates
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        std::swap(rows[i], rows[i + (int)(s % (uint64_t)(N - i))]);
    }
    int* d_rows;
    CK(cudaMalloc(&d_rows, (size_t)nsets * B * sizeof(int)));
    CK(cudaMemcpy(d_rows, rows.data(), (size_t)nsets * B * sizeof(int), cudaMemcpyHostToDevice));
    g_nsets = nsets; g_rows_base = d_rows; g_B = B;
    CK(cudaMalloc(&d_sum, ncta * sizeof(unsigned)));
    Geo g{mat, pitch, d_rows, B, (B + 127) / 128, ntt, d_sum};
    g_geo = &g;
    printf("gather probe: %d x %lld byte matrix (%.2f GB), B = %d random rows, %d CTAs x %d B per row = %.1f MB per launch, "
           "%d disjoint row sets used in turn (%s)\n",
           N, pitch, (double)N * pitch / 1e9, B, ncta, ntt * 64, g_bytes / 1e6, nsets,
           nsets > 1 ? "every launch gathers rows that are not in L2, like a training step" : "the same rows every launch: L2-resident");

    run("ldg 64B/row, row block inner", [&] { k_ldg<false><<<ncta, 512>>>(g); });
    run("ldg 64B/row, row block outer", [&] { k_ldg<true><<<ncta, 512>>>(g); });
    run("ldg 16B/row per unit (decoder pattern)", [&] { k_ldg16<0><<<ncta, 384>>>(g); });
    run("ldg 16B/row per unit, .L2::128B hint", [&] { k_ldg16<128><<<ncta, 384>>>(g); });
    run("ldg 16B/row per unit, .L2::256B hint", [&] { k_ldg16<256><<<ncta, 384>>>(g); });
#define CPA(BO, TB, NG, D, label)                                                                                      \
    {                                                                                                                  \
        const size_t sm = (size_t)NG * D * 128 * TB;                                                                   \
        CK(cudaFuncSetAttribute(k_cpasync<BO, TB, NG, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     \
        run(label, [&] { k_cpasync<BO, TB, NG, D><<<ncta, NG * 128, sm>>>(g); });                                      \
    }
#define CPAH(BO, TB, NG, D, H, label)                                                                                  \
    {                                                                                                                  \
        const size_t sm = (size_t)NG * D * 128 * TB;                                                                   \
        CK(cudaFuncSetAttribute(k_cpasync<BO, TB, NG, D, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));  \
        run(label, [&] { k_cpasync<BO, TB, NG, D, H><<<ncta, NG * 128, sm>>>(g); });                                   \
    }
    CPAH(false, 64, 4, 2, 128, "cp.async 64B/row inner, 4x2, .L2::128B hint")
    CPAH(false, 64, 4, 2, 256, "cp.async 64B/row inner, 4x2, .L2::256B hint")
    CPAH(true, 64, 4, 2, 256, "cp.async 64B/row outer, 4x2, .L2::256B hint")
    CPAH(false, 64, 4, 6, 256, "cp.async 64B/row inner, 4x6, .L2::256B hint")
    CPAH(true, 128, 4, 3, 256, "cp.async 128B/row outer, 4x3, .L2::256B hint")
    CPA(false, 64, 4, 2, "cp.async 64B/row inner, 4 groups x 2 deep")
    CPA(false, 64, 4, 6, "cp.async 64B/row inner, 4 groups x 6 deep")
    CPA(true, 64, 4, 2, "cp.async 64B/row outer, 4 groups x 2 deep")
    CPA(true, 64, 4, 6, "cp.async 64B/row outer, 4 groups x 6 deep")
    CPA(true, 64, 3, 4, "cp.async 64B/row outer, 3 groups x 4 deep")
    CPA(true, 64, 8, 3, "cp.async 64B/row outer, 8 groups x 3 deep")
    CPA(false, 128, 4, 3, "cp.async 128B/row inner, 4 groups x 3 deep")
    CPA(true, 128, 4, 3, "cp.async 128B/row outer, 4 groups x 3 deep")
    CPA(true, 128, 3, 4, "cp.async 128B/row outer, 3 groups x 4 deep")
    if (ntt % 4 == 0) {
        CPA(false, 256, 4, 1, "cp.async 256B/row inner, 4 groups x 1 deep")
        CPA(true, 256, 4, 1, "cp.async 256B/row outer, 4 groups x 1 deep")
        CPA(true, 256, 2, 3, "cp.async 256B/row outer, 2 groups x 3 deep")
        CPA(true, 256, 3, 2, "cp.async 256B/row outer, 3 groups x 2 deep")
    }
    if (ntt % 8 == 0) {
        CPA(true, 512, 3, 1, "cp.async 512B/row outer, 3 groups x 1 deep")
        CPA(true, 512, 1, 3, "cp.async 512B/row outer, 1 group x 3 deep")
    }
    if (ntt % 16 == 0) CPA(true, 1024, 1, 1, "cp.async 1024B/row outer, 1 group x 1 deep")
#define BLK(BO, TB, NG, D, label)                                                                                      \
    {                                                                                                                  \
        const size_t sm = (size_t)NG * D * 128 * TB + 1024;                                                            \
        CK(cudaFuncSetAttribute(k_bulk<BO, TB, NG, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));        \
        run(label, [&] { k_bulk<BO, TB, NG, D><<<ncta, NG * 128, sm>>>(g); });                                         \
    }
    BLK(false, 64, 4, 4, "bulk 64B copies inner, 4 groups x 4 deep")
    BLK(true, 64, 4, 4, "bulk 64B copies outer, 4 groups x 4 deep")
    BLK(true, 128, 4, 3, "bulk 128B copies outer, 4 groups x 3 deep")
    BLK(true, 256, 2, 3, "bulk 256B copies outer, 2 groups x 3 deep")
    if (ntt % 8 == 0) BLK(true, 512, 1, 3, "bulk 512B copies outer, 1 group x 3 deep")
    if (ntt % 16 == 0) BLK(true, 1024, 1, 1, "bulk 1024B copies outer, 1 group x 1 deep")

    // tensor map for gather4: dim0 = bytes of a row, dim1 = rows; box = TB bytes x 1 row
    EncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres) != cudaSuccess || enc == nullptr) {
        printf("gather4: cuTensorMapEncodeTiled not available\n");
        return 0;
    }
#define G4(BO, TB, NG, D, BOXROWS, label)                                                                              \
    {                                                                                                                  \
        CUtensorMap tm;                                                                                                \
        cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)N};                                                       \
        cuuint64_t strides[1] = {(cuuint64_t)pitch};                                                                   \
        cuuint32_t box[2] = {TB, BOXROWS};   /* (which box height tile::gather4 wants is part of what is probed) */                                                                             \
        cuuint32_t estr[2] = {1, 1};                                                                                   \
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)mat, dims, strides, box, estr,                  \
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,  \
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);                                                           \
        if (r != CUDA_SUCCESS) printf("%-46s cuTensorMapEncodeTiled failed (%d)\n", label, (int)r);                    \
        else {                                                                                                         \
            const size_t sm = (size_t)NG * D * 128 * TB + 1024;                                                        \
            CK(cudaFuncSetAttribute(k_gather4<BO, TB, NG, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
            run(label, [&] { k_gather4<BO, TB, NG, D><<<ncta, NG * 128, sm>>>(g, tm); });                              \
        }                                                                                                              \
    }
    G4(false, 64, 4, 4, 1, "gather4 4x64B boxes inner, 4 groups x 4 deep")
    G4(true, 64, 4, 4, 1, "gather4 4x64B boxes outer, 4 groups x 4 deep")
    G4(true, 128, 4, 3, 1, "gather4 4x128B boxes outer, 4 groups x 3 deep")
    G4(true, 128, 2, 3, 1, "gather4 4x128B boxes outer, 2 groups x 3 deep")
    G4(true, 128, 4, 3, 4, "gather4 4x128B, box rows 4, outer, 4 groups x 3")
    if (ntt % 4 == 0) G4(true, 256, 2, 3, 1, "gather4 4x256B boxes outer, 2 groups x 3 deep")
    return 0;
}
