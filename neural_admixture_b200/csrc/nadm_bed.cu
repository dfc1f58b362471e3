// PLINK .bed (SNP-major 2-bit) -> sample-major 2-bit packed genotype matrix, on the device, without ever expanding to
// one byte per genotype.
//
// Reference path being replaced: SNPReader._read_bed (src/snp_reader.py:16-45) -> utils_c.read_bed
// (src/utils_c/utils.pyx:43-68: LUT [2,3,1,0] per 2-bit field into an N x M uint8 host array, 50 GB at 100k x 500k),
// the global allele flip `G if G.mean() < 1 else 2 - G` (snp_reader.py:110, uint8 arithmetic: missing 3 -> 255, whose
// low two bits are 3 again once packed), then pack2bit_cpu_to_gpu (src/utils_c/pack2bit.cu:65-117).
//
// .bed layout: SNP m is a row of ceil(N/4) bytes, sample 4b+i in bits 2i..2i+1 of byte b; field 00 -> 2, 01 -> 3
// (missing), 10 -> 1, 11 -> 0.  Output layout: nadm_b200.h (sample n is a row, SNP 4c+i in bits 2i..2i+1 of byte c).
// The kernel is a 2-bit transpose through shared memory: tile = 128 SNPs x 256 samples (8 KB in, 8 KB out).
#include "nadm_common.cuh"

namespace nadm {

constexpr int kBedSnps = 128, kBedSamples = 256;

// LUT [2,3,1,0] on all 16 fields of a word:  hi' = ~hi ; lo' = lo ^ hi
__device__ __forceinline__ uint32_t bed_recode(uint32_t w) {
    const uint32_t hi = w & 0xAAAAAAAAu, lo = w & 0x55555555u;
    return (~hi & 0xAAAAAAAAu) | ((lo ^ (hi >> 1)) & 0x55555555u);
}
// g -> 2 - g for g in {0,1,2}; 3 stays 3:  lo' = lo ; hi' = ~(hi ^ lo)
__device__ __forceinline__ uint32_t flip_codes(uint32_t w) {
    const uint32_t x = ~((w >> 1) ^ w) & 0x55555555u;
    return (w & 0x55555555u) | (x << 1);
}
__device__ __forceinline__ uint32_t field_mask(int64_t nvalid) {   // low 2*nvalid bits
    return nvalid >= 16 ? 0xFFFFFFFFu : (nvalid <= 0 ? 0u : ((1u << (2 * (int)nvalid)) - 1u));
}

__global__ void __launch_bounds__(kBedSamples)
bed_to_packed_kernel(const uint8_t* __restrict__ bed, int64_t bed_pitch, int64_t N, int64_t M, int flip,
                     uint8_t* __restrict__ dst, int64_t dst_pitch, int64_t dst_byte0,
                     unsigned long long* __restrict__ counts) {
    __shared__ uint8_t tile[kBedSnps][kBedSamples / 4 + 4];      // [SNP][sample byte], +4: rows land on different banks
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * kBedSnps, b0 = (int64_t)blockIdx.y * (kBedSamples / 4);
    for (int i = tid; i < kBedSnps * (kBedSamples / 4); i += kBedSamples) {
        const int r = i / (kBedSamples / 4), c = i % (kBedSamples / 4);
        const int64_t m = m0 + r, b = b0 + c;
        tile[r][c] = (m < M && b < bed_pitch) ? bed[m * bed_pitch + b] : (uint8_t)0xFF;     // 11 -> code 0
    }
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.y * kBedSamples + tid;
    const int sb = tid >> 2, sh = 2 * (tid & 3);
    uint32_t out[kBedSnps / 16];
    uint32_t c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
    for (int w = 0; w < kBedSnps / 16; ++w) {
        uint32_t raw = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) raw |= (uint32_t)((tile[w * 16 + j][sb] >> sh) & 3u) << (2 * j);
        uint32_t g = bed_recode(raw) & field_mask(M - (m0 + w * 16));
        const uint32_t lo = g & 0x55555555u, hi = (g >> 1) & 0x55555555u;
        c1 += __popc(lo & ~hi);
        c2 += __popc(hi & ~lo);
        c3 += __popc(hi & lo);
        if (flip) g = flip_codes(g) & field_mask(M - (m0 + w * 16));
        out[w] = g;
    }
    if (n < N) {
        const int64_t off = dst_byte0 + m0 / 4;
        uint8_t* row = dst + n * dst_pitch + off;
#pragma unroll
        for (int q = 0; q < kBedSnps / 64; ++q)
            if (off + 16 * q + 16 <= dst_pitch)
                *reinterpret_cast<uint4*>(row + 16 * q) = make_uint4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
    } else {
        c1 = c2 = c3 = 0;
    }
    if (counts != nullptr) {
        c1 = __reduce_add_sync(0xffffffffu, c1);
        c2 = __reduce_add_sync(0xffffffffu, c2);
        c3 = __reduce_add_sync(0xffffffffu, c3);
        if ((tid & 31) == 0) {
            if (c1) atomicAdd(counts + 1, (unsigned long long)c1);
            if (c2) atomicAdd(counts + 2, (unsigned long long)c2);
            if (c3) atomicAdd(counts + 3, (unsigned long long)c3);
        }
    }
}

// in-place g -> 2 - g on a packed matrix (missing stays missing, zero tails stay zero)
__global__ void flip_packed_kernel(uint8_t* __restrict__ packed, int64_t pitch, int64_t N, int64_t M) {
    const int64_t words = pitch / 4;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * words) return;
    const int64_t w = i % words;
    uint32_t* p = reinterpret_cast<uint32_t*>(packed) + i;
    *p = flip_codes(*p) & field_mask(M - w * 16);
}

}  // namespace nadm

using namespace nadm;

extern "C" int nadm_bed_to_packed(const uint8_t* bed, int64_t bed_pitch, int64_t N, int64_t M, int64_t snp0, int32_t flip,
                                  uint8_t* dst, int64_t dst_pitch, uint64_t* counts, void* stream) {
    NADM_REQUIRE(bed && dst, "NULL pointer");
    NADM_REQUIRE(N > 0 && M > 0 && bed_pitch >= (N + 3) / 4, "bad .bed shape: N=%lld M=%lld row bytes=%lld", (long long)N,
                 (long long)M, (long long)bed_pitch);
    NADM_REQUIRE(snp0 >= 0 && snp0 % kBedSnps == 0, "snp0=%lld must be a multiple of %d", (long long)snp0, kBedSnps);
    NADM_REQUIRE(dst_pitch % 16 == 0 && ((uintptr_t)dst % 16) == 0, "destination rows must be 16-byte aligned");
    NADM_REQUIRE(dst_pitch * 4 >= snp0 + M, "destination pitch %lld too small for SNPs [%lld, %lld)", (long long)dst_pitch,
                 (long long)snp0, (long long)(snp0 + M));
    const int64_t gy = (N + kBedSamples - 1) / kBedSamples;
    NADM_REQUIRE(gy <= 65535, "N=%lld too large for one call (max %d samples)", (long long)N, 65535 * kBedSamples);
    dim3 grid((unsigned)((M + kBedSnps - 1) / kBedSnps), (unsigned)gy);
    bed_to_packed_kernel<<<grid, kBedSamples, 0, (cudaStream_t)stream>>>(bed, bed_pitch, N, M, flip, dst, dst_pitch,
                                                                        snp0 / 4, (unsigned long long*)counts);
    NADM_CHECK_LAUNCH("bed_to_packed_kernel");
    return NADM_OK;
}

extern "C" int nadm_flip_packed(uint8_t* packed, int64_t pitch, int64_t N, int64_t M, void* stream) {
    NADM_REQUIRE(packed && N > 0 && M > 0, "NULL pointer or empty matrix");
    NADM_REQUIRE(pitch % 4 == 0 && pitch * 4 >= M && ((uintptr_t)packed % 4) == 0, "bad pitch / alignment");
    const int64_t n = N * (pitch / 4);
    flip_packed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(packed, pitch, N, M);
    NADM_CHECK_LAUNCH("flip_packed_kernel");
    return NADM_OK;
}
