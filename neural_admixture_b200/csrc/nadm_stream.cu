// Streaming kernels over the 2-bit packed genotype matrix (fp32 CUDA-core formulation).
//
//   pack2bit / unpack2bit     : bit layout of pack2bit.cu:26-31,53-60
//   enc_fwd_kernel            : Z = X V            (neural_admixture.py:169-172)  thread-per-sample-row, tiles staged
//                               through shared memory with cp.async, V tile broadcast from shared memory
//   dec_kernel                : fused decoder + BCE loss + backward + Adam + clamp for one head
//                               (neural_admixture.py:83-98, :288, :410-412)  lane<->genotype byte, warps split rows
//   enc_bwd_kernel            : dV = X^T dZ + Adam (neural_admixture.py:172 backward, :411)
//   reduce_parts_kernel       : deterministic sum of per-CTA partials (no atomics anywhere on the path)
//   loglik_kernel             : fp64 log-likelihood (utils.pyx:17-40)
#include "nadm_common.cuh"

namespace nadm {

// =================================================================================================================
// pack / unpack
// =================================================================================================================
__global__ void pack2bit_kernel(const uint8_t* __restrict__ src, int64_t rows, int64_t M, int64_t src_pitch,
                                uint8_t* __restrict__ dst, int64_t dst_pitch) {
    // one thread packs 16 SNPs into one 32-bit word; grid.y strides over rows
    const int64_t words = (dst_pitch + 3) / 4;
    const int64_t pc = (M + 3) / 4;
    for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
        const uint8_t* s = src + row * src_pitch;
        uint8_t* d = dst + row * dst_pitch;
        for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < words; w += (int64_t)gridDim.x * blockDim.x) {
            const int64_t m0 = w * 16;
            uint32_t out = 0;
            if (m0 + 16 <= M && ((reinterpret_cast<uintptr_t>(s + m0) & 15) == 0)) {
                uint4 v = *reinterpret_cast<const uint4*>(s + m0);
                uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t x = q[j] & 0x03030303u;  // four codes, one per byte
                    uint32_t b = (x | (x >> 6) | (x >> 12) | (x >> 18)) & 0xFFu;
                    out |= b << (8 * j);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    int64_t m = m0 + i;
                    uint32_t c = (m < M) ? (uint32_t)(s[m] & 3) : 0u;
                    out |= c << (2 * i);
                }
            }
            // store only the bytes that exist in this row (dst_pitch need not be a multiple of 4)
            const int64_t b0 = w * 4;
            if (b0 + 4 <= dst_pitch && ((reinterpret_cast<uintptr_t>(d + b0) & 3) == 0)) {
                *reinterpret_cast<uint32_t*>(d + b0) = out;
            } else {
                for (int j = 0; j < 4; ++j)
                    if (b0 + j < dst_pitch) d[b0 + j] = (uint8_t)(out >> (8 * j));
            }
            (void)pc;
        }
    }
}

__global__ void unpack2bit_kernel(const uint8_t* __restrict__ src, int64_t rows, int64_t M, int64_t src_pitch,
                                  uint8_t* __restrict__ dst, int64_t dst_pitch) {
    const int64_t pc = (M + 3) / 4;
    for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
        const uint8_t* s = src + row * src_pitch;
        uint8_t* d = dst + row * dst_pitch;
        for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < pc; c += (int64_t)gridDim.x * blockDim.x) {
            uint32_t b = s[c];
            uint32_t four = (b & 3u) | ((b & 0xCu) << 6) | ((b & 0x30u) << 12) | ((b & 0xC0u) << 18);
            const int64_t m0 = c * 4;
            if (m0 + 4 <= M && ((reinterpret_cast<uintptr_t>(d + m0) & 3) == 0)) {
                *reinterpret_cast<uint32_t*>(d + m0) = four;
            } else {
                for (int i = 0; i < 4; ++i)
                    if (m0 + i < M) d[m0 + i] = (uint8_t)((four >> (8 * i)) & 3u);
            }
        }
    }
}

// =================================================================================================================
// deterministic partial reduction: out[i*out_ld_scale...] = sum_p part[p][i]
// =================================================================================================================
// part: nparts x (rows x cols_p) ; out: rows x out_ld, written at column offset out_off for cols < cols_out.
// Block = 8 consecutive outputs x 32 part segments; segment sums are combined in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
reduce_parts_kernel(const float* __restrict__ part, int nparts, int rows, int cols_p, int cols_out,
                    float* __restrict__ out, int out_ld, int out_off, float scale,
                    const float* __restrict__ loss_part, float* __restrict__ loss) {
    pdl_prologue();
    __shared__ float red[32][8];
    const int64_t n = (int64_t)rows * cols_p;
    const int o = threadIdx.x & 7, seg = threadIdx.x >> 3;
    const int64_t i = (int64_t)blockIdx.x * 8 + o;
    float acc = 0.f;
    if (i < n)
        for (int q = seg; q < nparts; q += 32) acc += part[(int64_t)q * n + i];
    red[seg][o] = acc;
    __syncthreads();
    if (threadIdx.x < 8 && i < n) {
        float t = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < 32; ++s2) t += red[s2][o];
        const int r = (int)(i / cols_p), c = (int)(i % cols_p);
        if (c < cols_out) out[(int64_t)r * out_ld + out_off + c] = t * scale;
    }
    if (loss_part != nullptr && loss != nullptr && blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 64) {
        double acc2 = 0.0;
        for (int q = threadIdx.x - 32; q < nparts; q += 32) acc2 += (double)loss_part[q];
        acc2 = warp_sum_d(acc2);
        if (threadIdx.x == 32) *loss = (float)((double)*loss + acc2);
    }
}

// =================================================================================================================
// encoder forward: thread per sample row, genotype + V tiles staged through shared memory (2-stage cp.async)
// =================================================================================================================
constexpr int kEncRowBytes = kEncTileSnps / 4;      // 64 genotype bytes per row per tile
constexpr int kEncRowStride = kEncRowBytes + 16;    // padded: conflict-free 16-byte reads across lanes

template <int CP>
__global__ void __launch_bounds__(256, 2)
enc_fwd_kernel(const uint8_t* __restrict__ packed, int64_t pitch, const int64_t* __restrict__ row_idx, int64_t row0,
               int B, int64_t M, const float* __restrict__ V, int C, float* __restrict__ Zpart, int ntiles,
               int nslab) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int RB = blockDim.x;
    const int tid = threadIdx.x;
    const int slab = blockIdx.x;
    const int bbase = blockIdx.y * RB;
    int64_t* rowoff = reinterpret_cast<int64_t*>(smem);
    const int stage_bytes = RB * kEncRowStride + kEncTileSnps * CP * 4;
    uint8_t* stage0 = smem + ((RB * 8 + 15) / 16) * 16;

    {
        const int b = bbase + tid;
        int64_t r = 0;
        if (b < B) r = (row_idx != nullptr) ? row_idx[b] : (row0 + b);
        rowoff[tid] = r * pitch;
    }
    __syncthreads();

    const int t0 = (int)(((int64_t)ntiles * slab) / nslab);
    const int t1 = (int)(((int64_t)ntiles * (slab + 1)) / nslab);
    const bool v_vec = (C == CP) && ((reinterpret_cast<uintptr_t>(V) & 15) == 0);

    auto issue = [&](int t, int s) {
        uint8_t* gs = stage0 + s * stage_bytes;
        float* vs = reinterpret_cast<float*>(gs + RB * kEncRowStride);
        const int64_t byte0 = (int64_t)t * kEncRowBytes;
        for (int i = tid; i < RB * 4; i += RB) {
            const int rr = i >> 2, ch = i & 3;
            const int64_t off = byte0 + ch * 16;
            const bool ok = (bbase + rr < B) && (off + 16 <= pitch);
            const uint8_t* src = ok ? (packed + rowoff[rr] + off) : packed;
            cp_async16(gs + rr * kEncRowStride + ch * 16, src, ok ? 16 : 0);
        }
        const int64_t m0 = (int64_t)t * kEncTileSnps;
        if (v_vec) {
            constexpr int chunks = kEncTileSnps * CP / 4;
            for (int i = tid; i < chunks; i += RB) {
                const int64_t m = m0 + (i * 4) / CP;
                const bool ok = m < M;
                const float* src = ok ? (V + m0 * CP + (int64_t)i * 4) : V;
                cp_async16(vs + i * 4, src, ok ? 16 : 0);
            }
        } else {
            for (int i = tid; i < kEncTileSnps * CP; i += RB) {
                const int s_ = i / CP, c = i % CP;
                const int64_t m = m0 + s_;
                const bool ok = (m < M) && (c < C);
                const float* src = ok ? (V + m * C + c) : V;
                cp_async4(vs + i, src, ok ? 4 : 0);
            }
        }
        cp_async_commit();
    };

    float acc[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[c] = 0.f;

    if (t0 < t1) issue(t0, 0);
    for (int t = t0; t < t1; ++t) {
        const int s = (t - t0) & 1;
        if (t + 1 < t1) {
            issue(t + 1, s ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint8_t* gs = stage0 + s * stage_bytes;
        const float* vs = reinterpret_cast<const float*>(gs + RB * kEncRowStride);
        const uint4* grow = reinterpret_cast<const uint4*>(gs + tid * kEncRowStride);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            const uint4 w4 = grow[ch];
            const uint32_t ws[4] = {clear_missing(w4.x), clear_missing(w4.y), clear_missing(w4.z), clear_missing(w4.w)};
#pragma unroll
            for (int wi = 0; wi < 4; ++wi) {
                const uint32_t w = ws[wi];
                if (w == 0u) continue;  // all-zero genotypes contribute nothing (per-thread skip, no sync inside)
                const float* vrow = vs + (ch * 64 + wi * 16) * CP;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float x2 = __uint_as_float(0x4B000000u | ((w >> (2 * j)) & 3u)) - 8388608.0f;
#pragma unroll
                    for (int c4 = 0; c4 < CP / 4; ++c4) {
                        const float4 v = *reinterpret_cast<const float4*>(vrow + j * CP + c4 * 4);
                        acc[c4 * 4 + 0] = fmaf(x2, v.x, acc[c4 * 4 + 0]);
                        acc[c4 * 4 + 1] = fmaf(x2, v.y, acc[c4 * 4 + 1]);
                        acc[c4 * 4 + 2] = fmaf(x2, v.z, acc[c4 * 4 + 2]);
                        acc[c4 * 4 + 3] = fmaf(x2, v.w, acc[c4 * 4 + 3]);
                    }
                }
            }
        }
        __syncthreads();
    }
    const int b = bbase + tid;
    if (b < B) {
        float* out = Zpart + ((int64_t)slab * B + b) * CP;
#pragma unroll
        for (int c = 0; c < CP; ++c) out[c] = acc[c];
    }
}

// =================================================================================================================
// fused decoder (one head): lane <-> genotype byte (4 SNPs), the CTA's warps split the batch rows
// =================================================================================================================
template <int KP>
__global__ void __launch_bounds__(kStreamWarps * 32)
dec_kernel(const uint8_t* __restrict__ packed, int64_t pitch, const int64_t* __restrict__ row_idx, int64_t row0, int B,
           int64_t M, const float* __restrict__ Q, int q_ld, int q_off, int k, float* __restrict__ P,
           float* __restrict__ Pm, float* __restrict__ Pv, AdamCoef adam_in, float* __restrict__ dP_out,
           float* __restrict__ dQpart, float* __restrict__ loss_part, int ntiles) {
    const AdamCoef adam = adam_resolve(adam_in);
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int W = kStreamWarps;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t* rowoff = reinterpret_cast<int64_t*>(smem);                       // B
    float* Qs = reinterpret_cast<float*>(smem + (size_t)((B * 8 + 15) / 16) * 16);  // B x KP
    float* dQs = Qs + (size_t)B * KP;                                           // B x KP
    float* red = dQs + (size_t)B * KP;                                          // W x 128*KP

    for (int b = tid; b < B; b += blockDim.x) {
        const int64_t r = (row_idx != nullptr) ? row_idx[b] : (row0 + b);
        rowoff[b] = r * pitch;
    }
    for (int i = tid; i < B * KP; i += blockDim.x) {
        const int b = i / KP, kk = i % KP;
        Qs[i] = (kk < k) ? Q[(int64_t)b * q_ld + q_off + kk] : 0.f;
        dQs[i] = 0.f;
    }
    __syncthreads();

    float lossacc = 0.f;
    constexpr int OWN = 32 / KP;  // lanes per reduced element after warp_reduce_vec<KP>

    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t m_lane = (int64_t)t * kTileSnps + lane * 4;
        const int64_t byte_off = (int64_t)t * (kTileSnps / 4) + lane;
        const bool byte_ok = byte_off < pitch;
        float p[4][KP], dp[4][KP];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int kk = 0; kk < KP; ++kk) {
                const int64_t m = m_lane + i;
                p[i][kk] = (m < M && kk < k) ? P[m * k + kk] : 0.f;
                dp[i][kk] = 0.f;
            }

        uint32_t g_next = 0;
        if (warp < B && byte_ok) g_next = packed[rowoff[warp] + byte_off];
        for (int b = warp; b < B; b += W) {
            const uint32_t g = g_next;
            if (b + W < B && byte_ok) g_next = packed[rowoff[b + W] + byte_off];
            float q[KP], dq[KP];
#pragma unroll
            for (int c4 = 0; c4 < KP / 4; ++c4) {
                const float4 v = *reinterpret_cast<const float4*>(Qs + (size_t)b * KP + c4 * 4);
                q[c4 * 4 + 0] = v.x; q[c4 * 4 + 1] = v.y; q[c4 * 4 + 2] = v.z; q[c4 * 4 + 3] = v.w;
            }
#pragma unroll
            for (int kk = 0; kk < KP; ++kk) dq[kk] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t c = (g >> (2 * i)) & 3u;
                const float x = (c == 1u) ? 0.5f : ((c == 2u) ? 1.0f : 0.0f);
                float raw = 0.f;
#pragma unroll
                for (int kk = 0; kk < KP; ++kk) raw = fmaf(q[kk], p[i][kk], raw);
                const float R = fminf(fmaxf(raw, 0.f), 1.f);
                const float omr = 1.f - R;
                const float den = fmaxf(omr * R, 1e-12f);
                const bool inside = (raw >= 0.f) && (raw <= 1.f);
                const float G = inside ? ((R - x) / den) : 0.f;
                // BCE with torch's -100 log clamp; X in {0,.5,1} -> one log per element
                const float arg = (c == 2u) ? R : ((c == 1u) ? (R * omr) : omr);
                const float wgt = (c == 1u) ? 0.5f : 1.0f;
                lossacc -= wgt * fmaxf(logf(arg), -100.f);
#pragma unroll
                for (int kk = 0; kk < KP; ++kk) {
                    dp[i][kk] = fmaf(G, q[kk], dp[i][kk]);
                    dq[kk] = fmaf(G, p[i][kk], dq[kk]);
                }
            }
            const float tot = warp_reduce_vec<KP>(dq, lane);
            if ((lane % OWN) == 0) dQs[(size_t)b * KP + lane / OWN] += tot;
        }

        // cross-warp reduction of dP, then Adam + clamp on this tile's P rows
        float* myred = red + (size_t)warp * (kTileSnps * KP) + (size_t)lane * (4 * KP);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c4 = 0; c4 < KP / 4; ++c4)
                *reinterpret_cast<float4*>(myred + i * KP + c4 * 4) =
                    make_float4(dp[i][c4 * 4], dp[i][c4 * 4 + 1], dp[i][c4 * 4 + 2], dp[i][c4 * 4 + 3]);
        __syncthreads();
        for (int o = tid; o < kTileSnps * KP; o += blockDim.x) {
            float g = 0.f;
#pragma unroll
            for (int w = 0; w < W; ++w) g += red[(size_t)w * (kTileSnps * KP) + o];
            const int sl = o / KP, kk = o % KP;
            const int64_t m = (int64_t)t * kTileSnps + sl;
            if (m < M && kk < k) {
                const int64_t pi = m * k + kk;
                if (dP_out != nullptr) dP_out[pi] = g;
                if (adam.enabled) {
                    float mm = Pm[pi], vv = Pv[pi];
                    float pn = adam_apply(P[pi], g, mm, vv, adam);
                    pn = fminf(fmaxf(pn, 0.f), 1.f);  // restrict_P (neural_admixture.py:179-185)
                    P[pi] = pn; Pm[pi] = mm; Pv[pi] = vv;
                }
            }
        }
        __syncthreads();
    }

    // per-CTA partials out
    float* outp = dQpart + (size_t)blockIdx.x * B * KP;
    for (int i = tid; i < B * KP; i += blockDim.x) outp[i] = dQs[i];
    double l = warp_sum_d((double)lossacc);
    __shared__ double lsh[W];
    if (lane == 0) lsh[warp] = l;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < W; ++w) s += lsh[w];
        loss_part[blockIdx.x] = (float)s;
    }
}

// =================================================================================================================
// encoder backward: dV = X^T dZ (+ Adam on V); same tiling as the decoder
// =================================================================================================================
template <int CP>
__global__ void __launch_bounds__(kStreamWarps * 32)
enc_bwd_kernel(const uint8_t* __restrict__ packed, int64_t pitch, const int64_t* __restrict__ row_idx, int64_t row0,
               int B, int64_t M, const float* __restrict__ dZ, int C, float* __restrict__ V, float* __restrict__ Vm,
               float* __restrict__ Vv, AdamCoef adam_in, float* __restrict__ dV_out, int ntiles) {
    const AdamCoef adam = adam_resolve(adam_in);
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int W = kStreamWarps;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t* rowoff = reinterpret_cast<int64_t*>(smem);
    float* dZs = reinterpret_cast<float*>(smem + (size_t)((B * 8 + 15) / 16) * 16);  // B x CP
    float* red = dZs + (size_t)B * CP;                                             // W x 128*CP

    for (int b = tid; b < B; b += blockDim.x) {
        const int64_t r = (row_idx != nullptr) ? row_idx[b] : (row0 + b);
        rowoff[b] = r * pitch;
    }
    for (int i = tid; i < B * CP; i += blockDim.x) {
        const int b = i / CP, c = i % CP;
        dZs[i] = (c < C) ? dZ[(int64_t)b * C + c] : 0.f;
    }
    __syncthreads();

    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t byte_off = (int64_t)t * (kTileSnps / 4) + lane;
        const bool byte_ok = byte_off < pitch;
        float dv[4][CP];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < CP; ++c) dv[i][c] = 0.f;

        uint32_t g_next = 0;
        if (warp < B && byte_ok) g_next = packed[rowoff[warp] + byte_off];
        for (int b = warp; b < B; b += W) {
            const uint32_t g = g_next;
            if (b + W < B && byte_ok) g_next = packed[rowoff[b + W] + byte_off];
            if (__ballot_sync(0xffffffffu, g != 0u) == 0u) continue;
            float dz[CP];
#pragma unroll
            for (int c4 = 0; c4 < CP / 4; ++c4) {
                const float4 v = *reinterpret_cast<const float4*>(dZs + (size_t)b * CP + c4 * 4);
                dz[c4 * 4 + 0] = v.x; dz[c4 * 4 + 1] = v.y; dz[c4 * 4 + 2] = v.z; dz[c4 * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float x2 = code_to_2x((g >> (2 * i)) & 3u);
#pragma unroll
                for (int c = 0; c < CP; ++c) dv[i][c] = fmaf(x2, dz[c], dv[i][c]);
            }
        }
        float* myred = red + (size_t)warp * (kTileSnps * CP) + (size_t)lane * (4 * CP);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c4 = 0; c4 < CP / 4; ++c4)
                *reinterpret_cast<float4*>(myred + i * CP + c4 * 4) =
                    make_float4(dv[i][c4 * 4], dv[i][c4 * 4 + 1], dv[i][c4 * 4 + 2], dv[i][c4 * 4 + 3]);
        __syncthreads();
        for (int o = tid; o < kTileSnps * CP; o += blockDim.x) {
            float g = 0.f;
#pragma unroll
            for (int w = 0; w < W; ++w) g += red[(size_t)w * (kTileSnps * CP) + o];
            g *= 0.5f;  // x = code/2
            const int sl = o / CP, c = o % CP;
            const int64_t m = (int64_t)t * kTileSnps + sl;
            if (m < M && c < C) {
                const int64_t vi = m * C + c;
                if (dV_out != nullptr) dV_out[vi] = g;
                if (adam.enabled) {
                    float mm = Vm[vi], vv = Vv[vi];
                    V[vi] = adam_apply(V[vi], g, mm, vv, adam);
                    Vm[vi] = mm; Vv[vi] = vv;
                }
            }
        }
        __syncthreads();
    }
}

// =================================================================================================================
// fp64 log-likelihood (utils.pyx:17-40): lane <-> genotype byte, warps split rows, Q rows read through L1
// =================================================================================================================
__global__ void __launch_bounds__(256)
loglik_kernel(const uint8_t* __restrict__ packed, int64_t pitch, int64_t N, int64_t M, const float* __restrict__ Q,
              const float* __restrict__ P, int k, double eps, double* __restrict__ part, int ntiles) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int W = 8;
    double acc = 0.0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t byte_off = (int64_t)t * 32 + lane;
        const int64_t m_lane = (int64_t)t * 128 + lane * 4;
        if (byte_off >= pitch) continue;
        double p[4][NADM_MAX_K];
        for (int i = 0; i < 4; ++i)
            for (int kk = 0; kk < k; ++kk) p[i][kk] = (m_lane + i < M) ? (double)P[(m_lane + i) * k + kk] : 0.0;
        for (int64_t n = warp; n < N; n += W) {
            const uint32_t g = packed[n * pitch + byte_off];
            const float* q = Q + n * k;
            for (int i = 0; i < 4; ++i) {
                const uint32_t c = (g >> (2 * i)) & 3u;
                if (c == 3u || m_lane + i >= M) continue;
                double rec = 0.0;
                for (int kk = 0; kk < k; ++kk) rec += (double)q[kk] * p[i][kk];
                rec = fmax(eps, fmin(rec, 1.0 - eps));
                double gd = fmax(eps, fmin((double)c, 2.0 - eps));
                acc += gd * log(rec) + (2.0 - gd) * log1p(-rec);
            }
        }
    }
    acc = warp_sum_d(acc);
    __shared__ double sh[W];
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < W; ++w) s += sh[w];
        part[blockIdx.x] = s;
    }
}

__global__ void sum_double_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) acc += part[i];
    acc = warp_sum_d(acc);
    if (threadIdx.x == 0) *out = acc;
}

}  // namespace nadm

// =================================================================================================================
// C ABI
// =================================================================================================================
using namespace nadm;

static int check_packed(const uint8_t* packed, int64_t pitch, int64_t M) {
    NADM_REQUIRE(packed != nullptr, "packed is NULL");
    NADM_REQUIRE(pitch >= (M + 3) / 4, "pitch %lld < ceil(M/4) = %lld", (long long)pitch, (long long)((M + 3) / 4));
    NADM_REQUIRE(pitch % 16 == 0, "pitch %lld must be a multiple of 16 bytes", (long long)pitch);
    NADM_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "packed base must be 16-byte aligned");
    return NADM_OK;
}

extern "C" int nadm_pack2bit(const uint8_t* src, int64_t rows, int64_t M, int64_t src_pitch, uint8_t* dst,
                             int64_t dst_pitch, void* stream) {
    NADM_REQUIRE(src && dst, "NULL pointer");
    NADM_REQUIRE(rows >= 0 && M >= 0, "negative shape");
    NADM_REQUIRE(src_pitch >= M && dst_pitch >= (M + 3) / 4, "pitch too small");
    if (rows == 0 || dst_pitch == 0) return NADM_OK;
    const int64_t words = (dst_pitch + 3) / 4;
    dim3 grid((unsigned)std::min<int64_t>((words + 255) / 256, 4096), (unsigned)std::min<int64_t>(rows, 32768));
    pack2bit_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, rows, M, src_pitch, dst, dst_pitch);
    NADM_CHECK_LAUNCH("pack2bit_kernel");
    return NADM_OK;
}

extern "C" int nadm_unpack2bit(const uint8_t* src, int64_t rows, int64_t M, int64_t src_pitch, uint8_t* dst,
                               int64_t dst_pitch, void* stream) {
    NADM_REQUIRE(src && dst, "NULL pointer");
    NADM_REQUIRE(rows >= 0 && M >= 0, "negative shape");
    NADM_REQUIRE(src_pitch >= (M + 3) / 4 && dst_pitch >= M, "pitch too small");
    if (rows == 0 || M == 0) return NADM_OK;
    const int64_t pc = (M + 3) / 4;
    dim3 grid((unsigned)std::min<int64_t>((pc + 255) / 256, 4096), (unsigned)std::min<int64_t>(rows, 32768));
    unpack2bit_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, rows, M, src_pitch, dst, dst_pitch);
    NADM_CHECK_LAUNCH("unpack2bit_kernel");
    return NADM_OK;
}

namespace nadm {
int launch_reduce_parts(const float* part, int nparts, int rows, int cols_p, int cols_out, float* out, int out_ld,
                        int out_off, float scale, const float* loss_part, float* loss, cudaStream_t st) {
    const int64_t n = (int64_t)rows * cols_p;
    launch_pdl(reduce_parts_kernel, dim3((unsigned)((n + 7) / 8)), dim3(256), 0, st, part, nparts, rows, cols_p, cols_out, out,
               out_ld, out_off, scale, loss_part, loss);
    NADM_CHECK_LAUNCH("reduce_parts_kernel");
    return NADM_OK;
}
}  // namespace nadm

static inline int pad_c(int C) { return C <= 8 ? 8 : 16; }
// =================================================================================================================
// device-side step bookkeeping (CUDA-graph replayable steps): see nadm_b200.h
// =================================================================================================================
namespace nadm {
__global__ void step_begin_kernel(const int64_t* __restrict__ order, int64_t order_len, const int64_t* __restrict__ counters,
                                  int64_t stride, int B, int64_t* __restrict__ row_idx_out, float lr, float beta1,
                                  float beta2, float eps, float* __restrict__ coef_out, float* __restrict__ loss_accum) {
    pdl_prologue();
    const int64_t s = counters[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const int64_t j = s * stride + i;
        row_idx_out[i] = (j < order_len) ? order[j] : 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double t = (double)(counters[1] + 1);
        const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
        coef_out[0] = beta1;
        coef_out[1] = beta2;
        coef_out[2] = 1.0f - beta1;
        coef_out[3] = 1.0f - beta2;
        coef_out[4] = (float)((double)lr / bc1);
        coef_out[5] = (float)(1.0 / sqrt(bc2));
        coef_out[6] = eps;
        reinterpret_cast<int*>(coef_out)[7] = 1;
        if (loss_accum != nullptr) *loss_accum = 0.f;
    }
}
__global__ void step_end_kernel(int64_t* __restrict__ counters, const float* __restrict__ loss, float* __restrict__ losses_out) {
    pdl_prologue();
    const int64_t s = counters[0];
    if (loss != nullptr && losses_out != nullptr) losses_out[s] = *loss;
    counters[0] = s + 1;
    counters[1] += 1;
}
// nadm_step_next: finish the pending step (if any), then begin the next one and leave it pending: ONE bookkeeping kernel
// per replayed step instead of two.  Single CTA (the counters are read, then written, by the same block).
__device__ __forceinline__ void step_finish_pending(int64_t* counters, float* losses_out) {
    if (counters[2] != 0) {
        const float* lp = reinterpret_cast<const float*>(static_cast<uintptr_t>(counters[3]));
        if (lp != nullptr && losses_out != nullptr) losses_out[counters[0]] = *lp;
        counters[0] += 1;
        counters[1] += 1;
        counters[2] = 0;
    }
}
__global__ void __launch_bounds__(256)
step_next_kernel(const int64_t* __restrict__ order, int64_t order_len, int64_t* counters, int64_t stride, int B,
                 int64_t* __restrict__ row_idx_out, float lr, float beta1, float beta2, float eps,
                 float* __restrict__ coef_out, float* loss_accum, int record_loss, float* losses_out) {
    pdl_prologue();
    // every thread reads the counters BEFORE thread 0 rewrites them (one round trip instead of finish -> sync -> begin)
    const int64_t c0 = counters[0], c1 = counters[1], pend = counters[2], lptr = counters[3];
    __syncthreads();
    const int64_t s = c0 + (pend != 0 ? 1 : 0), steps_done = c1 + (pend != 0 ? 1 : 0);
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const int64_t j = s * stride + i;
        row_idx_out[i] = (j < order_len) ? order[j] : 0;
    }
    if (threadIdx.x == 0) {
        if (pend != 0) {                                                  // finish the pending step (= nadm_step_end)
            const float* lp = reinterpret_cast<const float*>(static_cast<uintptr_t>(lptr));
            if (lp != nullptr && losses_out != nullptr) losses_out[c0] = *lp;
        }
        if (loss_accum != nullptr) *loss_accum = 0.f;                     // (after the read above: it may be the same word)
        counters[0] = s;
        counters[1] = steps_done;
        counters[2] = 1;
        counters[3] = record_loss ? (int64_t)reinterpret_cast<uintptr_t>(loss_accum) : 0;
    }
    if (threadIdx.x == 32) {
        const double t = (double)(steps_done + 1);
        const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
        coef_out[0] = beta1;
        coef_out[1] = beta2;
        coef_out[2] = 1.0f - beta1;
        coef_out[3] = 1.0f - beta2;
        coef_out[4] = (float)((double)lr / bc1);
        coef_out[5] = (float)(1.0 / sqrt(bc2));
        coef_out[6] = eps;
        reinterpret_cast<int*>(coef_out)[7] = 1;
    }
}
__global__ void step_flush_kernel(int64_t* counters, float* losses_out) {
    pdl_prologue();
    step_finish_pending(counters, losses_out);
}
}  // namespace nadm

extern "C" int nadm_step_next(const int64_t* order, int64_t order_len, int64_t* counters, int64_t stride, int32_t B,
                              int64_t* row_idx_out, const nadm_adam_t* hyper, void* coef_out, float* loss_accum,
                              int32_t record_loss, float* losses_out, void* stream) {
    if (int rc0 = nadm::flush_deferred_apply((cudaStream_t)stream)) return rc0;   // (a pending update adds the supervised loss term)
    NADM_REQUIRE(order && counters && row_idx_out && hyper && coef_out, "NULL pointer");
    NADM_REQUIRE(B > 0 && stride >= B && order_len > 0, "bad minibatch geometry (B=%d, stride=%lld)", B, (long long)stride);
    NADM_REQUIRE(((uintptr_t)coef_out & 15) == 0, "coef_out must be 16-byte aligned");
    NADM_REQUIRE(!record_loss || loss_accum != nullptr, "record_loss needs a loss accumulator");
    nadm::launch_pdl(nadm::step_next_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, order, order_len, counters, stride, B,
                     row_idx_out, hyper->lr, hyper->beta1, hyper->beta2, hyper->eps, (float*)coef_out, loss_accum,
                     (int)record_loss, losses_out);
    NADM_CHECK_LAUNCH("step_next_kernel");
    return NADM_OK;
}

extern "C" int nadm_step_flush(int64_t* counters, float* losses_out, void* stream) {
    if (int rc0 = nadm::flush_deferred_apply((cudaStream_t)stream)) return rc0;   // (a pending update adds the supervised loss term)
    NADM_REQUIRE(counters, "NULL pointer");
    nadm::launch_pdl(nadm::step_flush_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, counters, losses_out);
    NADM_CHECK_LAUNCH("step_flush_kernel");
    return NADM_OK;
}

extern "C" int nadm_step_begin(const int64_t* order, int64_t order_len, int64_t* counters, int64_t stride, int32_t B,
                               int64_t* row_idx_out, const nadm_adam_t* hyper, void* coef_out, float* loss_accum,
                               void* stream) {
    if (int rc0 = nadm::flush_deferred_apply((cudaStream_t)stream)) return rc0;   // (a pending update adds the supervised loss term)
    NADM_REQUIRE(order && counters && row_idx_out && hyper && coef_out, "NULL pointer");
    NADM_REQUIRE(B > 0 && stride >= B && order_len > 0, "bad minibatch geometry (B=%d, stride=%lld)", B, (long long)stride);
    NADM_REQUIRE(((uintptr_t)coef_out & 15) == 0, "coef_out must be 16-byte aligned");
    nadm::launch_pdl(nadm::step_begin_kernel, dim3((B + 255) / 256), dim3(256), 0, (cudaStream_t)stream, order, order_len,
                     counters, stride, B, row_idx_out, hyper->lr, hyper->beta1, hyper->beta2, hyper->eps, (float*)coef_out,
                     loss_accum);
    NADM_CHECK_LAUNCH("step_begin_kernel");
    return NADM_OK;
}

extern "C" int nadm_step_end(int64_t* counters, const float* loss, float* losses_out, void* stream) {
    if (int rc0 = nadm::flush_deferred_apply((cudaStream_t)stream)) return rc0;   // (a pending update adds the supervised loss term)
    NADM_REQUIRE(counters, "NULL pointer");
    nadm::launch_pdl(nadm::step_end_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, counters, loss, losses_out);
    NADM_CHECK_LAUNCH("step_end_kernel");
    return NADM_OK;
}

static inline int pad_k(int k) { return k <= 4 ? 4 : (k <= 8 ? 8 : 16); }

extern "C" size_t nadm_workspace_bytes(int32_t B, int64_t M, int32_t C, int32_t H, int32_t sumK) {
    (void)M;
    size_t enc = (size_t)kMaxParts * (size_t)B * 16 * sizeof(float);
    size_t dec = (size_t)kMaxParts * ((size_t)B * 16 + 1) * sizeof(float);
    size_t mlp = mlp_bwd_workspace_bytes(B, C, H, sumK) + 4096;
    size_t ll = (size_t)kMaxParts * sizeof(double) * 2;
    size_t enc_tc = enc_tc_workspace_bytes(B);
    enc = enc > enc_tc ? enc : enc_tc;
    size_t m = enc > dec + mlp ? enc : dec + mlp;   // (a deferred dQ reduction keeps the decoder's partials alive next to
                                                    //  the MLP backward's slabs)
    m = m > ll ? m : ll;
    return m + 256;
}

template <int CP>
static int launch_enc_fwd(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int B, int64_t M,
                          const float* V, int C, float* Z, float* ws, size_t ws_bytes, cudaStream_t st) {
    const int ngroups = (B + 255) / 256;
    int RB = (B + ngroups - 1) / ngroups;
    RB = ((RB + 31) / 32) * 32;
    const int ntiles = (int)((M + kEncTileSnps - 1) / kEncTileSnps);
    int nslab = std::max(1, (2 * sm_count()) / ngroups);
    nslab = std::min(std::min(nslab, ntiles), kMaxParts);
    NADM_REQUIRE((size_t)nslab * B * CP * sizeof(float) <= ws_bytes, "workspace too small for encoder_fwd");
    const size_t smem = ((RB * 8 + 15) / 16) * 16 + 2 * ((size_t)RB * kEncRowStride + kEncTileSnps * CP * 4);
    static PerDeviceOnce once;                       // (one instance per template instantiation)
    bool* attr = once.slot();
    if (attr == nullptr || !*attr) {
        cudaError_t e = cudaFuncSetAttribute(enc_fwd_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(enc_fwd)");
        if (attr) *attr = true;
    }
    enc_fwd_kernel<CP><<<dim3(nslab, ngroups), RB, smem, st>>>(packed, pitch, row_idx, row0, B, M, V, C, ws, ntiles, nslab);
    NADM_CHECK_LAUNCH("enc_fwd_kernel");
    count_generic();
    const int64_t n = (int64_t)B * CP;
    reduce_parts_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(ws, nslab, B, CP, C, Z, C, 0, 0.5f, nullptr, nullptr);
    NADM_CHECK_LAUNCH("reduce_parts_kernel");
    return NADM_OK;
}

namespace nadm {
DeferredZ& deferred_z() { static thread_local DeferredZ d{}; return d; }
DeferredDQ& deferred_dq() { static thread_local DeferredDQ d{}; return d; }
}
// a reduction left pending for this dQ by an earlier nadm_decoder_step_deferred (another head of the same step) is
// completed before the workspace is reused
static int flush_deferred_dq(const float* dQ, cudaStream_t st) {
    DeferredDQ& r = deferred_dq();
    if (r.dQ == nullptr || r.dQ != dQ) return NADM_OK;
    const DeferredDQ d = r;
    r.dQ = nullptr;
    return launch_reduce_parts(d.part, d.nparts, d.B, d.cols_p, d.k, const_cast<float*>(d.dQ), d.q_ld, d.q_off, 1.0f,
                               d.loss_part, d.loss, st);
}

static int encoder_fwd_impl(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int32_t B,
                            int64_t M, const float* V, int32_t C, float* Z, void* ws, size_t ws_bytes, void* stream,
                            bool defer) {
    if (deferred_z().Z == Z) deferred_z().Z = nullptr;               // (a pending record for this Z is superseded)
    if (int rc = check_packed(packed, pitch, M)) return rc;
    NADM_REQUIRE(B > 0 && M > 0, "empty batch or no SNPs (B=%d, M=%lld)", B, (long long)M);
    NADM_REQUIRE(C >= 1 && C <= NADM_MAX_C, "n_components C=%d unsupported (1..%d)", C, NADM_MAX_C);
    NADM_REQUIRE(V && Z && ws, "NULL pointer");
    NADM_REQUIRE(row0 >= 0 && row0 + B <= (1ll << 32), "row numbers must fit 32 bits (row0=%lld)", (long long)row0);
    if (C <= 8 && !use_generic_kernels() && (reinterpret_cast<uintptr_t>(V) & 15) == 0) {
        // tensor-core path: at most 2048 rows (16 blocks of 128) per launch (1024 for the slab-fed variant)
        const int chunk = enc_fwd_slab() ? 1024 : 2048;
        for (int r0 = 0; r0 < B; r0 += chunk) {
            const int nb = std::min(chunk, B - r0);
            if (int rc = launch_enc_fwd_tc(packed, pitch, row_idx ? row_idx + r0 : nullptr, row0 + r0, nb, M, V, C,
                                           Z + (int64_t)r0 * C, ws, ws_bytes, (cudaStream_t)stream, -1,
                                           defer && B <= chunk && !enc_fwd_slab()))
                return rc;
        }
        return NADM_OK;
    }
    if (pad_c(C) == 8)
        return launch_enc_fwd<8>(packed, pitch, row_idx, row0, B, M, V, C, Z, (float*)ws, ws_bytes, (cudaStream_t)stream);
    return launch_enc_fwd<16>(packed, pitch, row_idx, row0, B, M, V, C, Z, (float*)ws, ws_bytes, (cudaStream_t)stream);
}
extern "C" int nadm_encoder_fwd(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int32_t B,
                                int64_t M, const float* V, int32_t C, float* Z, void* ws, size_t ws_bytes,
                                void* stream) {
    return encoder_fwd_impl(packed, pitch, row_idx, row0, B, M, V, C, Z, ws, ws_bytes, stream, false);
}
extern "C" int nadm_encoder_fwd_deferred(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0,
                                         int32_t B, int64_t M, const float* V, int32_t C, float* Z, void* ws,
                                         size_t ws_bytes, void* stream) {
    return encoder_fwd_impl(packed, pitch, row_idx, row0, B, M, V, C, Z, ws, ws_bytes, stream, true);
}

template <int KP>
static int launch_dec(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int B, int64_t M,
                      const float* Q, float* dQ, int q_ld, int q_off, int k, float* P, float* Pm, float* Pv,
                      const nadm_adam_t* adam, float* dP_out, float* loss, float* ws, size_t ws_bytes,
                      cudaStream_t st) {
    const int ntiles = (int)((M + kTileSnps - 1) / kTileSnps);
    const size_t smem = ((size_t)(B * 8 + 15) / 16) * 16 + (size_t)2 * B * KP * 4 + (size_t)kStreamWarps * kTileSnps * KP * 4;
    NADM_REQUIRE(smem <= (size_t)kMaxDynSmem, "batch B=%d too large for the fused decoder (needs %zu bytes of shared memory)", B, smem);
    const int per_sm = std::max(1, (int)((227 * 1024) / (smem + 2048)));
    int ncta = std::min(std::min(ntiles, sm_count() * std::min(per_sm, 3)), kMaxParts);
    NADM_REQUIRE((size_t)ncta * ((size_t)B * KP + 1) * sizeof(float) <= ws_bytes, "workspace too small for decoder_step");
    static PerDeviceOnce once;
    bool* attr = once.slot();
    if (attr == nullptr || !*attr) {
        cudaError_t e = cudaFuncSetAttribute(dec_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dec)");
        if (attr) *attr = true;
    }
    float* dQpart = ws;
    float* loss_part = ws + (size_t)ncta * B * KP;
    dec_kernel<KP><<<ncta, kStreamWarps * 32, smem, st>>>(packed, pitch, row_idx, row0, B, M, Q, q_ld, q_off, k, P, Pm, Pv,
                                                         make_adam(adam), dP_out, dQpart, loss_part, ntiles);
    NADM_CHECK_LAUNCH("dec_kernel");
    count_generic();
    const int64_t n = (int64_t)B * KP;
    reduce_parts_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(dQpart, ncta, B, KP, k, dQ, q_ld, q_off, 1.0f, loss_part, loss);
    NADM_CHECK_LAUNCH("reduce_parts_kernel");
    return NADM_OK;
}

static int decoder_step_impl(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int32_t B,
                             int64_t M, const float* Q, float* dQ, int32_t q_ld, int32_t q_off, int32_t k, float* P,
                             float* Pm, float* Pv, const nadm_adam_t* adam, float* dP_out, float* loss, void* ws,
                             size_t ws_bytes, void* stream, bool defer) {
    if (int rc = flush_deferred_dq(dQ, (cudaStream_t)stream)) return rc;
    if (int rc = check_packed(packed, pitch, M)) return rc;
    NADM_REQUIRE(B > 0 && M > 0, "empty batch or no SNPs (B=%d, M=%lld)", B, (long long)M);
    NADM_REQUIRE(k >= 1 && k <= NADM_MAX_K, "k=%d unsupported (1..%d)", k, NADM_MAX_K);
    NADM_REQUIRE(q_off >= 0 && q_off + k <= q_ld, "head columns [%d,%d) outside q_ld=%d", q_off, q_off + k, q_ld);
    NADM_REQUIRE(Q && dQ && P && ws, "NULL pointer");
    NADM_REQUIRE(adam == nullptr || (Pm && Pv), "Adam moments are NULL");
    cudaStream_t st = (cudaStream_t)stream;
    float* w = (float*)ws;
    if (!use_generic_kernels() && dec_tc_supported(B, k) && (reinterpret_cast<uintptr_t>(P) & 15) == 0 &&
        (Pm == nullptr || (reinterpret_cast<uintptr_t>(Pm) & 15) == 0) &&
        (Pv == nullptr || (reinterpret_cast<uintptr_t>(Pv) & 15) == 0) &&
        (dP_out == nullptr || (reinterpret_cast<uintptr_t>(dP_out) & 15) == 0))
        return launch_dec_tc(packed, pitch, row_idx, row0, B, M, Q, dQ, q_ld, q_off, k, P, Pm, Pv, adam, dP_out, loss, w,
                             ws_bytes, st, defer);
    switch (pad_k(k)) {
        case 4: return launch_dec<4>(packed, pitch, row_idx, row0, B, M, Q, dQ, q_ld, q_off, k, P, Pm, Pv, adam, dP_out, loss, w, ws_bytes, st);
        case 8: return launch_dec<8>(packed, pitch, row_idx, row0, B, M, Q, dQ, q_ld, q_off, k, P, Pm, Pv, adam, dP_out, loss, w, ws_bytes, st);
        default: return launch_dec<16>(packed, pitch, row_idx, row0, B, M, Q, dQ, q_ld, q_off, k, P, Pm, Pv, adam, dP_out, loss, w, ws_bytes, st);
    }
}
extern "C" int nadm_decoder_step(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int32_t B,
                                 int64_t M, const float* Q, float* dQ, int32_t q_ld, int32_t q_off, int32_t k, float* P,
                                 float* Pm, float* Pv, const nadm_adam_t* adam, float* dP_out, float* loss, void* ws,
                                 size_t ws_bytes, void* stream) {
    return decoder_step_impl(packed, pitch, row_idx, row0, B, M, Q, dQ, q_ld, q_off, k, P, Pm, Pv, adam, dP_out, loss, ws,
                             ws_bytes, stream, false);
}
extern "C" int nadm_decoder_step_deferred(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0,
                                          int32_t B, int64_t M, const float* Q, float* dQ, int32_t q_ld, int32_t q_off,
                                          int32_t k, float* P, float* Pm, float* Pv, const nadm_adam_t* adam,
                                          float* dP_out, float* loss, void* ws, size_t ws_bytes, void* stream) {
    return decoder_step_impl(packed, pitch, row_idx, row0, B, M, Q, dQ, q_ld, q_off, k, P, Pm, Pv, adam, dP_out, loss, ws,
                             ws_bytes, stream, true);
}

template <int CP>
static int launch_enc_bwd(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int B, int64_t M,
                          const float* dZ, int C, float* V, float* Vm, float* Vv, const nadm_adam_t* adam, float* dV_out,
                          cudaStream_t st) {
    const int ntiles = (int)((M + kTileSnps - 1) / kTileSnps);
    const size_t smem = ((size_t)(B * 8 + 15) / 16) * 16 + (size_t)B * CP * 4 + (size_t)kStreamWarps * kTileSnps * CP * 4;
    NADM_REQUIRE(smem <= (size_t)kMaxDynSmem, "batch B=%d too large for encoder_bwd (needs %zu bytes of shared memory)", B, smem);
    const int per_sm = std::max(1, (int)((227 * 1024) / (smem + 2048)));
    const int ncta = std::min(ntiles, sm_count() * std::min(per_sm, 3));
    static PerDeviceOnce once;
    bool* attr = once.slot();
    if (attr == nullptr || !*attr) {
        cudaError_t e = cudaFuncSetAttribute(enc_bwd_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(enc_bwd)");
        if (attr) *attr = true;
    }
    enc_bwd_kernel<CP><<<ncta, kStreamWarps * 32, smem, st>>>(packed, pitch, row_idx, row0, B, M, dZ, C, V, Vm, Vv,
                                                             make_adam(adam), dV_out, ntiles);
    NADM_CHECK_LAUNCH("enc_bwd_kernel");
    count_generic();
    return NADM_OK;
}

extern "C" int nadm_encoder_bwd(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int32_t B,
                                int64_t M, const float* dZ, int32_t C, float* V, float* Vm, float* Vv,
                                const nadm_adam_t* adam, float* dV_out, void* ws, size_t ws_bytes, void* stream) {
    (void)ws; (void)ws_bytes;
    if (int rc = check_packed(packed, pitch, M)) return rc;
    NADM_REQUIRE(B > 0 && M > 0, "empty batch or no SNPs (B=%d, M=%lld)", B, (long long)M);
    NADM_REQUIRE(C >= 1 && C <= NADM_MAX_C, "n_components C=%d unsupported (1..%d)", C, NADM_MAX_C);
    NADM_REQUIRE(dZ && V, "NULL pointer");
    NADM_REQUIRE(adam == nullptr || (Vm && Vv), "Adam moments are NULL");
    NADM_REQUIRE(adam != nullptr || dV_out != nullptr, "nothing to do: neither Adam nor dV_out requested");
    NADM_REQUIRE(row0 >= 0 && row0 + B <= (1ll << 32), "row numbers must fit 32 bits (row0=%lld)", (long long)row0);
    const bool tc = C <= 8 && enc_bwd_tc_supported(B) && !use_generic_kernels() && (reinterpret_cast<uintptr_t>(V) & 15) == 0 &&
                    (dV_out == nullptr || (reinterpret_cast<uintptr_t>(dV_out) & 15) == 0);
    // a parameter update of the network left pending for this dZ (nadm_mlp_bwd_deferred) rides along on the kernel's
    // epilogue warps; in every other case it runs now, as its own kernel
    if (tc && enc_bwd_runs_apply(B) && deferred_apply().dZ == dZ && dZ != nullptr) {
        const ApplyJob job = deferred_apply().job;
        deferred_apply().dZ = nullptr;
        return launch_enc_bwd_tc(packed, pitch, row_idx, row0, B, M, dZ, C, V, Vm, Vv, adam, dV_out, (cudaStream_t)stream, -1, 0, &job);
    }
    if (int rc = flush_deferred_apply((cudaStream_t)stream)) return rc;
    if (tc)
        return launch_enc_bwd_tc(packed, pitch, row_idx, row0, B, M, dZ, C, V, Vm, Vv, adam, dV_out, (cudaStream_t)stream);
    if (pad_c(C) == 8)
        return launch_enc_bwd<8>(packed, pitch, row_idx, row0, B, M, dZ, C, V, Vm, Vv, adam, dV_out, (cudaStream_t)stream);
    return launch_enc_bwd<16>(packed, pitch, row_idx, row0, B, M, dZ, C, V, Vm, Vv, adam, dV_out, (cudaStream_t)stream);
}

// ---- randomized-SVD products on the packed matrix (scope row f2) -------------------------------------------------------
extern "C" int nadm_geno_matmul(const uint8_t* packed, int64_t pitch, int64_t N, int64_t M, const float* Omega, int32_t K,
                                int32_t missing_value, float* Y, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = check_packed(packed, pitch, M)) return rc;
    NADM_REQUIRE(N > 0 && M > 0, "empty matrix");
    NADM_REQUIRE(K >= 1 && K <= 8, "K=%d: at most 8 columns per call (process wider factors in column chunks)", K);
    NADM_REQUIRE(missing_value >= 0 && missing_value <= 255, "missing_value must be a uint8 value");
    NADM_REQUIRE(Omega && Y && ws, "NULL pointer");
    NADM_REQUIRE((reinterpret_cast<uintptr_t>(Omega) & 15) == 0, "Omega must be 16-byte aligned");
    for (int64_t r0 = 0; r0 < N; r0 += 1024) {        // 1024 rows per launch (two accumulator sets in tensor memory)
        const int nb = (int)std::min<int64_t>(1024, N - r0);
        if (int rc = launch_enc_fwd_tc(packed, pitch, nullptr, r0, nb, M, Omega, K, Y + r0 * K, ws, ws_bytes,
                                       (cudaStream_t)stream, missing_value))
            return rc;
    }
    return NADM_OK;
}

extern "C" int nadm_geno_matmul_t(const uint8_t* packed, int64_t pitch, int64_t N, int64_t M, const float* Q, int32_t K,
                                  int32_t missing_value, float* Bt, void* ws, size_t ws_bytes, void* stream) {
    (void)ws; (void)ws_bytes;
    if (int rc = check_packed(packed, pitch, M)) return rc;
    NADM_REQUIRE(N > 0 && M > 0, "empty matrix");
    NADM_REQUIRE(K >= 1 && K <= 8, "K=%d: at most 8 columns per call (process wider factors in column chunks)", K);
    NADM_REQUIRE(missing_value >= 0 && missing_value <= 255, "missing_value must be a uint8 value");
    NADM_REQUIRE(Q && Bt, "NULL pointer");
    NADM_REQUIRE((reinterpret_cast<uintptr_t>(Bt) & 15) == 0, "Bt must be 16-byte aligned");
    int batch = 1024;                                 // largest multiple of 128 rows the backward kernel's smem takes
    while (batch > 128 && !enc_bwd_tc_supported(batch)) batch -= 128;
    for (int64_t r0 = 0; r0 < N; r0 += batch) {       // row batches accumulate into Bt in a fixed order (deterministic)
        const int nb = (int)std::min<int64_t>(batch, N - r0);
        if (int rc = launch_enc_bwd_tc(packed, pitch, nullptr, r0, nb, M, Q + r0 * K, K, Bt, nullptr, nullptr, nullptr, Bt,
                                       (cudaStream_t)stream, missing_value, r0 > 0 ? 1 : 0))
            return rc;
    }
    return NADM_OK;
}

extern "C" int nadm_loglikelihood(const uint8_t* packed, int64_t pitch, int64_t N, int64_t M, const float* Q,
                                  const float* P, int32_t k, double eps, double* out, void* ws, size_t ws_bytes,
                                  void* stream) {
    NADM_REQUIRE(packed && Q && P && out && ws, "NULL pointer");
    NADM_REQUIRE(pitch >= (M + 3) / 4, "pitch too small");
    NADM_REQUIRE(k >= 1 && k <= NADM_MAX_K, "k=%d unsupported (1..%d)", k, NADM_MAX_K);
    NADM_REQUIRE(N > 0 && M > 0, "empty matrix");
    const int ntiles = (int)((M + 127) / 128);
    const int ncta = std::min(std::min(ntiles, 4 * sm_count()), kMaxParts);
    NADM_REQUIRE((size_t)ncta * sizeof(double) <= ws_bytes, "workspace too small for loglikelihood");
    double* part = (double*)ws;
    loglik_kernel<<<ncta, 256, 0, (cudaStream_t)stream>>>(packed, pitch, N, M, Q, P, k, eps, part, ntiles);
    NADM_CHECK_LAUNCH("loglik_kernel");
    sum_double_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(part, ncta, out);
    NADM_CHECK_LAUNCH("sum_double_kernel");
    return NADM_OK;
}
