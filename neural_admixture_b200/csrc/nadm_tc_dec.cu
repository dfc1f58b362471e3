// Fused decoder step on the 5th-generation tensor cores (tcgen05), sm_100a.  One head, k <= 8.
//
//   raw = Q P^T ; R = clamp(raw, 0, 1) ; loss += BCE_sum(R, X) ; G = dLoss/draw (BCELoss backward, 1e-12 floor,
//   inclusive clamp mask) ; dQ = G P ; dP = G^T Q ; Adam(P) ; P <- clamp(P, 0, 1)
//   (neural_admixture.py:83-98 forward, :288/:431 loss, :410 backward, :411 optimizer step, :179-185/:412 restrict_P)
//
// The B x M matrices raw / R / X(float) / G never exist in HBM: per unit (128 batch rows x 64 SNPs)
//   MMA1 (kind::f16, SS)   raw[128 x 64]  = Q_blk . P_sub^T          Q, P split EXACTLY into 3 bf16 terms h + m + l
//                                                                   (8+8+8 bits); 4 MMAs of K=16 cover every product
//                                                                   pair except l.l (2^-32): fp32-faithful
//   CUDA cores             tcgen05.ld raw -> G, loss  (one thread per batch row, 64 SNPs of its own 2-bit packed row)
//                          G split into bf16 hi + lo, written back over raw in tensor memory AND to shared memory
//   MMA2 (kind::f16, TS)   dQ_blk[128 x 32] += G[128 x 64] . [P_h|P_l|P_m|P_h]   A operand straight from tensor memory
//   MMA3 (kind::f16, SS)   dP_sub[64 x 24]  += G^T[64 x 128] . [Q_h|Q_m|Q_l]     A operand = the shared tile, MN-major
// (the K-major bf16 operand tiles of MMA1 are re-read as MN-major B operands by MMA2 / MMA3: nadm_tc.cuh)
// dQ accumulates in tensor memory over all SNP sub-tiles of the CTA (one 128 x 16 accumulator per row block), dP over
// the row blocks of one sub-tile, after which four epilogue warps apply Adam + clamp to the 64 x k slice of P.
//
// Warp roles (576 threads): 0-11 three compute warpgroups (unit u -> warpgroup u % 3, which owns tensor-memory slot
// u % 3), 12 MMA issuer (one elected thread), 13 P-tile producer, 14-17 dP/Adam epilogue.  Pipelines are mbarrier
// based; tcgen05.commit frees operand buffers.
#include "nadm_common.cuh"
#include "nadm_tc.cuh"

#include <cuda_bf16.h>

namespace nadm {
using namespace tc;

constexpr int kMS = 64;                       // SNPs per sub-tile
constexpr int kGtBytes = 2 * 128 * kMS * 2;   // G^T tile: two bf16 terms x 128 rows x 64 SNPs = 32 KB
constexpr int kPTileBytes = 4096;             // P sub-tile: 64 SNPs x 4 bf16 chunks [h | l | m | h] of 8 components
constexpr int kQBlkBytes = 16 * 384;          // Q block: 128 rows x 3 bf16 chunks [h | m | l]
constexpr int kPStages = 3;
constexpr int kWGs = 3;                       // compute warpgroups (4 warps each), one raw/G slot per warpgroup
constexpr int kWarpIssue = 4 * kWGs, kWarpProd = kWarpIssue + 1, kWarpEpi = kWarpIssue + 2;
constexpr int kDecThreads = (4 * kWGs + 2 + 4) * 32;
constexpr int kSlots = 3;                     // raw / G slots of 64 tensor-memory columns, used round-robin by the units
constexpr int kColD3 = 192, kColD2 = 256;     // tensor-memory columns: [0,192) slots, D3 2 x 32, D2 32 per row block
constexpr uint32_t kIdesc1 = instr_desc(kAccF32, kFmtBF16, kFmtBF16, false, false, 128, kMS);
constexpr uint32_t kIdesc2 = instr_desc(kAccF32, kFmtBF16, kFmtBF16, false, true, 128, 32);
constexpr uint32_t kIdesc3 = instr_desc(kAccF32, kFmtBF16, kFmtBF16, true, true, 64, 24);
constexpr float kLog2Clamp = -144.26950408889634f;   // -100 / ln 2: torch clamps log at -100 (BCELoss)

#ifdef NADM_TIMELINE
__device__ long long g_timeline[8][512];
#define TL(row, idx) do { if (blockIdx.x == 0 && (idx) < 512) g_timeline[row][idx] = clock64(); } while (0)
#else
#define TL(row, idx) do { } while (0)
#endif

struct DecSmem {
    uint64_t d1full[kSlots], gready[kSlots], gtfree[4], pfull[kPStages], pempty[kPStages], d3full[2], d3empty[2], alldone;
    uint32_t tmem_base;
    float lossred[16];
};

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {   // low half = a
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// exact 3-term bf16 split of an fp32 value: x = h + m + l, each term the next 8 significant bits (truncation).
// For x >= 0 all terms are >= 0, so every product of the split Q and P is non-negative (raw >= 0).
__device__ __forceinline__ void split3_bf16(float x, uint32_t& h, uint32_t& m, uint32_t& l) {
    const uint32_t hb = __float_as_uint(x) & 0xFFFF0000u;
    const float r = x - __uint_as_float(hb);
    const uint32_t mb = __float_as_uint(r) & 0xFFFF0000u;
    h = hb >> 16;
    m = mb >> 16;
    l = __float_as_uint(r - __uint_as_float(mb)) >> 16;
}
// 8 components -> three 16-byte chunks (bf16 h / m / l of components 0..7)
__device__ __forceinline__ void split3_row(const float (&x)[8], uint4& H, uint4& Mm, uint4& L) {
    uint32_t h[8], m[8], l[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) split3_bf16(x[c], h[c], m[c], l[c]);
    H = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
    Mm = make_uint4(m[0] | (m[1] << 16), m[2] | (m[3] << 16), m[4] | (m[5] << 16), m[6] | (m[7] << 16));
    L = make_uint4(l[0] | (l[1] << 16), l[2] | (l[3] << 16), l[4] | (l[5] << 16), l[6] | (l[7] << 16));
}

// One element: raw (>= 0 by construction), code field f2 = code * 4^jj as an exact float.
//   G and the two loss terms.  x = code / 2.
struct Elem {
    float G, l;
};

// Process 16 consecutive SNPs of one row: v[] holds raw on entry; on exit hi[] / lo[] hold the bf16x2-packed split of G.
// w: the 16 2-bit codes (missing cleared).  Loss accumulators in log2 units (kLoss = false: gradients only).
// kChecked = true is the general path (raw may exceed 1 by rounding: clamp + inclusive mask of the clamp backward);
// kChecked = false is taken when the caller has verified max(raw) <= 1 for the 16 values, where min / mask are no-ops.
template <bool kLoss, bool kChecked>
__device__ __forceinline__ void decode16(const uint32_t (&v)[16], uint32_t w, uint32_t magic, uint32_t (&hi)[8],
                                         uint32_t (&lo)[8], float& acc_all, float& acc_het) {
    const uint32_t wh = w >> 16;
#pragma unroll
    for (int j2 = 0; j2 < 8; ++j2) {
        float g[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int j = 2 * j2 + e;
            const uint32_t wsrc = (j < 8) ? w : wh;
            const int sh = 2 * (j & 7);
            const float raw = __uint_as_float(v[j]);
            // code * 4^(j&7) as an exact float via the 2^23 magic constant (kept in a register: one LOP3); x = code / 2
            const float f = __uint_as_float((wsrc & (3u << sh)) | magic) - 8388608.0f;
            const float Rs = kChecked ? fminf(raw, 1.0f) : raw;
            const float prod = fmaf(-Rs, Rs, Rs);                       // R (1 - R)
            const float inv = rcp_approx(fmaxf(prod, 1e-12f));
            const float num = fmaf(f, -0.5f / (float)(1 << sh), Rs);     // R - x
            float G = num * inv;
            if (kChecked) G = (raw <= 1.0f) ? G : 0.0f;                 // clamp backward mask (raw >= 0 always)
            g[e] = G;
            if (kLoss) {
                // BCE with torch's log clamp; X in {0, .5, 1}: a single log per element.
                // x = 0: 1 - R = 1 - |R - x| ;  x = 1: R = 1 - |R - x| ;  x = .5: weight .5 on log(R (1 - R))
                const bool het = (wsrc >> sh) & 1u;
                const float arg = het ? prod : (1.0f - fabsf(num));
                const float l = fmaxf(lg2_approx(arg), kLog2Clamp);
                acc_all += l;
                acc_het += het ? l : 0.0f;
            }
        }
        const uint32_t h = pack_bf16x2(g[0], g[1]);
        hi[j2] = h;
        lo[j2] = pack_bf16x2(g[0] - __uint_as_float(h << 16), g[1] - __uint_as_float(h & 0xFFFF0000u));
    }
}

template <bool kLoss>
__global__ void __launch_bounds__(kDecThreads, 1)
dec_tc_kernel(const uint8_t* __restrict__ packed, int64_t pitch, const int64_t* __restrict__ row_idx, int64_t row0, int B,
              int64_t M, const float* __restrict__ Q, int q_ld, int q_off, int k, float* __restrict__ P,
              float* __restrict__ Pm, float* __restrict__ Pv, AdamCoef adam, float* __restrict__ dP_out,
              float* __restrict__ dQpart, float* __restrict__ loss_part, int TS, int ngt) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int nblk = (B + 127) / 128;
    uint8_t* QA = smem;                                   // nblk x 6 KB : bf16 h/m/l of Q (MMA1 A K-major, MMA3 B MN-major)
    uint8_t* PT = QA + nblk * kQBlkBytes;                 // kPStages x 4 KB : P sub-tiles (MMA1 B K-major, MMA2 B MN-major)
    uint8_t* GT = PT + kPStages * kPTileBytes;            // ngt x 32 KB : G^T tiles
    int64_t* rowoff = reinterpret_cast<int64_t*>(GT + ngt * kGtBytes);
    DecSmem* S = reinterpret_cast<DecSmem*>(rowoff + nblk * 128);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s0 = (int)(((int64_t)TS * blockIdx.x) / gridDim.x), s1 = (int)(((int64_t)TS * (blockIdx.x + 1)) / gridDim.x);
    const int nsub = s1 - s0, U = nsub * nblk;

    // ---------------- one-time setup: row offsets, Q operands, barriers, tensor memory ----------------
    for (int b = tid; b < nblk * 128; b += blockDim.x)
        rowoff[b] = (b < B) ? ((row_idx != nullptr) ? row_idx[b] : (row0 + b)) * pitch : -1;
    for (int b = tid; b < nblk * 128; b += blockDim.x) {
        float q[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) q[c] = (b < B && c < k) ? Q[(int64_t)b * q_ld + q_off + c] : 0.f;
        uint4 H, Mm, L;
        split3_row(q, H, Mm, L);
        uint8_t* a = QA + (b & 7) * 16 + (b >> 3) * 384;
        *reinterpret_cast<uint4*>(a) = H;
        *reinterpret_cast<uint4*>(a + 128) = Mm;
        *reinterpret_cast<uint4*>(a + 256) = L;
    }
    if (tid == 0) {
        for (int i = 0; i < kSlots; ++i) { mbar_init(&S->d1full[i], 1); mbar_init(&S->gready[i], 4); }
        for (int i = 0; i < 4; ++i) mbar_init(&S->gtfree[i], 1);
        for (int i = 0; i < kPStages; ++i) { mbar_init(&S->pfull[i], 1); mbar_init(&S->pempty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&S->d3full[i], 1); mbar_init(&S->d3empty[i], 4); }
        mbar_init(&S->alldone, 1);
        mbar_init_fence();
    }
    if (warp == kWarpIssue) tmem_alloc<512>(&S->tmem_base);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = S->tmem_base;

    if (warp < kWarpIssue) {
        // =============================== compute warpgroups ===============================
        const int wg = warp >> 2, q = warp & 3;
        const int rb = q * 32 + lane;                                   // row inside the block = tensor-memory lane
        const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
        float acc_all = 0.f, acc_het = 0.f;
        uint4 gw = make_uint4(0u, 0u, 0u, 0u);
        auto load_codes = [&](int u) {
            const int blk = u % nblk;
            const int64_t ro = rowoff[blk * 128 + rb];
            const int64_t off = (int64_t)(s0 + u / nblk) * (kMS / 4);
            uint4 r = make_uint4(0u, 0u, 0u, 0u);
            if (ro >= 0 && off + 16 <= pitch) r = *reinterpret_cast<const uint4*>(packed + ro + off);
            return r;
        };
        uint32_t magic;
        asm volatile("mov.b32 %0, 0x4B000000;" : "=r"(magic));           // opaque to the compiler: stays in a register
        if (wg < U) gw = load_codes(wg);
        for (int u = wg; u < U; u += kWGs) {
            const int blk = u % nblk;
            const int slot = wg, g = u % ngt;
            const uint32_t cw[4] = {clear_missing(gw.x), clear_missing(gw.y), clear_missing(gw.z), clear_missing(gw.w)};
            if (u + kWGs < U) gw = load_codes(u + kWGs);
            const bool active = blk * 128 + q * 32 < B;                 // warp-uniform: any real row in this warp
            if (rb == 0) TL(0, u);                                      // WG starts waiting for raw(u)
            mbar_wait(&S->d1full[slot], ((u / kWGs) & 1));
            tc_fence_after_sync();
            if (rb == 0) TL(1, u);                                      // raw(u) seen
#ifdef NADM_SKIPDECODE
            if (false) {
#else
            if (active) {
#endif
                mbar_wait(&S->gtfree[g], ((u / ngt) & 1) ^ 1);
                if (rb == 0) TL(6, u);                                  // G^T buffer free
                uint8_t* gt = GT + g * kGtBytes + (rb & 7) * 16 + (rb >> 3) * 1024;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {                          // 16 SNPs at a time: raw columns [16c, 16c+16)
                    uint32_t v[16], hi[8], lo[8];
                    tmem_ld16(tlane + slot * 64 + c * 16, v);
                    tmem_wait_ld();
                    const uint32_t w = (c & 2) ? ((c & 1) ? cw[3] : cw[2]) : ((c & 1) ? cw[1] : cw[0]);
                    float mx = __uint_as_float(v[0]);
#pragma unroll
                    for (int j = 1; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                    if (mx <= 1.0f) decode16<kLoss, false>(v, w, magic, hi, lo, acc_all, acc_het);
                    else decode16<kLoss, true>(v, w, magic, hi, lo, acc_all, acc_het);
                    tmem_st8(tlane + slot * 64 + c * 16, hi);          // G hi / lo overwrite their own raw columns
                    tmem_st8(tlane + slot * 64 + c * 16 + 8, lo);
                    *reinterpret_cast<uint4*>(gt + (2 * c) * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(gt + (2 * c + 1) * 128) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    *reinterpret_cast<uint4*>(gt + 16384 + (2 * c) * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<uint4*>(gt + 16384 + (2 * c + 1) * 128) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                }
                tmem_wait_st();
                fence_async_smem();
            }
            tc_fence_before_sync();
            if (rb == 0) TL(2, u);                                      // G(u) written
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->gready[slot]);               // one arrival per warp
        }
        // ---- loss partial of this CTA (log2 units -> nats), dQ partial from tensor memory ----
        float l = -(acc_all - 0.5f * acc_het) * 0.6931471805599453f;
        l = warp_sum(l);
        if (lane == 0) S->lossred[warp] = l;
        mbar_wait(&S->alldone, 0);
        tc_fence_after_sync();
        for (int blk = wg; blk < nblk; blk += kWGs) {
            uint32_t v[32];
            tmem_ld32(tlane + kColD2 + blk * 32, v);
            tmem_wait_ld();
            const int b = blk * 128 + rb;
            if (b < B) {
                float* out = dQpart + ((int64_t)blockIdx.x * B + b) * 8;
                float o[8];
#pragma unroll
                for (int c = 0; c < 8; ++c)   // columns: G.P_h | G.P_l | G.P_m | G.P_h (duplicate, unused)
                    o[c] = (U > 0) ? (__uint_as_float(v[c]) + __uint_as_float(v[16 + c])) + __uint_as_float(v[8 + c]) : 0.f;
                const float4 o0 = make_float4(o[0], o[1], o[2], o[3]), o1 = make_float4(o[4], o[5], o[6], o[7]);
                reinterpret_cast<float4*>(out)[0] = o0;
                reinterpret_cast<float4*>(out)[1] = o1;
            }
        }
    } else if (warp == kWarpIssue) {
        // =============================== MMA issuer ===============================
        // The whole warp walks the unit sequence and waits on the barriers; one elected lane issues.  Descriptors are
        // built once; per instruction only the 14-bit start-address field (16-byte units) is advanced.
        const uint32_t qa = smem_u32(QA), pt = smem_u32(PT), gtb = smem_u32(GT);
        // Q chunks [h m l] (128 B apart, 8-row groups 384 B apart); P chunks [h l m h] (8-row groups 512 B apart).
        // A K=16 instruction multiplies two chunk pairs: start address = first chunk, LBO = distance to the second.
        const uint64_t A_hm = smem_desc(qa, 128, 384), A_hl = smem_desc(qa, 256, 384), A_ml = smem_desc(qa + 128, 128, 384);
        const uint64_t B_hm = smem_desc(pt, 256, 512), B_mh = smem_desc(pt + 256, 128, 512);
        const uint64_t B_lh = smem_desc(pt + 128, 256, 512), B_lm = smem_desc(pt + 128, 128, 512);
        const uint64_t B2 = smem_desc(pt, 512, 128);                       // P tile as MN-major [32 x K] operand
        const uint64_t A3 = smem_desc(gtb, 1024, 128), B3 = smem_desc(qa, 384, 128);
        // Counters are advanced incrementally (no divisions on the single issuing thread's critical path).
        // look-ahead unit (next MMA1): row block, P stage + its phase, slot
        int l_blk = 0, l_stage = 0, l_phase = 0, l_slot = 0, l_left = U;
        auto issue_mma1 = [&]() {
            if (l_blk == 0) {                                           // first unit of a sub-tile: its P tile must be there
                mbar_wait(&S->pfull[l_stage], l_phase);
                tc_fence_after_sync();
            }
            if (elect_one()) {
                const uint64_t ao = (uint64_t)(l_blk * (kQBlkBytes >> 4)), bo = (uint64_t)(l_stage * (kPTileBytes >> 4));
                const uint32_t d = tbase + l_slot * 64;
                mma_f16_ss(d, A_hm + ao, B_hm + bo, kIdesc1, 0u);       // h.h + m.m
                mma_f16_ss(d, A_hm + ao, B_mh + bo, kIdesc1, 1u);       // h.m + m.h
                mma_f16_ss(d, A_hl + ao, B_lh + bo, kIdesc1, 1u);       // h.l + l.h
                mma_f16_ss(d, A_ml + ao, B_lm + bo, kIdesc1, 1u);       // m.l + l.m
                mma_commit(&S->d1full[l_slot]);
            }
            __syncwarp();
            --l_left;
            l_slot = (l_slot == kSlots - 1) ? 0 : l_slot + 1;
            if (++l_blk == nblk) {
                l_blk = 0;
                if (++l_stage == kPStages) { l_stage = 0; l_phase ^= 1; }
            }
        };
        for (int i = 0; i < kSlots && l_left > 0; ++i) issue_mma1();
        const int nks_last = min(8, (B - (nblk - 1) * 128 + 15) / 16);
        int blk = 0, sub = 0, slot = 0, slot_phase = 0, g = 0, stage = 0, dbuf = 0, d3_phase = 1;
        for (int u = 0; u < U; ++u) {
            if (lane == 0) TL(3, u);                                    // issuer starts waiting for G(u)
            mbar_wait(&S->gready[slot], slot_phase);
            if (blk == 0) mbar_wait(&S->d3empty[dbuf], d3_phase);
            tc_fence_after_sync();
            if (lane == 0) TL(4, u);                                    // issuer saw G(u)
            if (elect_one()) {
                // dQ_blk += G . [P_h | P_l | P_m | P_h]   (A from tensor memory: per 16 SNPs, hi in 8 columns, lo in 8)
                const uint64_t b2 = B2 + (uint64_t)(stage * (kPTileBytes >> 4));
                const uint32_t d2 = tbase + kColD2 + blk * 32, a2 = tbase + slot * 64;
                const uint32_t acc2 = sub > 0 ? 1u : 0u;
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int t = 0; t < 2; ++t)
                        mma_f16_ts(d2, a2 + c * 16 + t * 8, b2 + (uint64_t)(c * 2 * 32), kIdesc2, (c + t) ? 1u : acc2);
            }
            __syncwarp();
            // raw of the unit that reuses this slot: queued right behind the MMAs that consume the slot's G
            if (l_left > 0) issue_mma1();
            if (elect_one()) {
                // dP_sub += G^T . [Q_h | Q_m | Q_l]    (A = shared G^T tile, MN-major; only K steps holding real rows)
                const uint64_t a3 = A3 + (uint64_t)(g * (kGtBytes >> 4)), b3 = B3 + (uint64_t)(blk * (kQBlkBytes >> 4));
                const uint32_t d3 = tbase + kColD3 + dbuf * 32;
                const uint32_t acc3 = blk > 0 ? 1u : 0u;
                if (blk != nblk - 1 || nks_last == 8) {
#pragma unroll
                    for (int t = 0; t < 2; ++t)
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            mma_f16_ss(d3, a3 + (uint64_t)(t * 1024 + ks * 128), b3 + (uint64_t)(ks * 48), kIdesc3,
                                       (t + ks) ? 1u : acc3);
                } else {
                    for (int t = 0; t < 2; ++t)
                        for (int ks = 0; ks < nks_last; ++ks)
                            mma_f16_ss(d3, a3 + (uint64_t)(t * 1024 + ks * 128), b3 + (uint64_t)(ks * 48), kIdesc3,
                                       (t + ks) ? 1u : acc3);
                }
                mma_commit(&S->gtfree[g]);
                if (blk == nblk - 1) {
                    mma_commit(&S->d3full[dbuf]);
                    mma_commit(&S->pempty[stage]);
                }
            }
            __syncwarp();
            if (lane == 0) TL(5, u);                                    // issuer done with unit u
            // advance the unit counters
            if (++slot == kSlots) { slot = 0; slot_phase ^= 1; }
            if (++g == ngt) g = 0;
            if (++blk == nblk) {
                blk = 0;
                ++sub;
                if (++stage == kPStages) stage = 0;
                dbuf ^= 1;
                if (dbuf == 0) d3_phase ^= 1;
            }
        }
        if (elect_one()) mma_commit(&S->alldone);
        __syncwarp();
    } else if (warp == kWarpProd) {
        // =============================== P sub-tile producer ===============================
        for (int sub = 0; sub < nsub; ++sub) {
            const int st = sub % kPStages;
            float p[2][8];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t m = (int64_t)(s0 + sub) * kMS + lane + 32 * e;
#pragma unroll
                for (int c = 0; c < 8; ++c) p[e][c] = 0.f;
                if (m < M) {
                    if (k == 8) {
                        const float4 a = reinterpret_cast<const float4*>(P + m * 8)[0];
                        const float4 b = reinterpret_cast<const float4*>(P + m * 8)[1];
                        p[e][0] = a.x; p[e][1] = a.y; p[e][2] = a.z; p[e][3] = a.w;
                        p[e][4] = b.x; p[e][5] = b.y; p[e][6] = b.z; p[e][7] = b.w;
                    } else {
                        for (int c = 0; c < k; ++c) p[e][c] = P[m * k + c];
                    }
                }
            }
            mbar_wait(&S->pempty[st], ((sub / kPStages) & 1) ^ 1);
            uint8_t* tile = PT + st * kPTileBytes;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ml = lane + 32 * e;
                uint8_t* t1 = tile + (ml & 7) * 16 + (ml >> 3) * 512;
                uint4 H, Mm, L;
                split3_row(p[e], H, Mm, L);
                *reinterpret_cast<uint4*>(t1) = H;
                *reinterpret_cast<uint4*>(t1 + 128) = L;
                *reinterpret_cast<uint4*>(t1 + 256) = Mm;
                *reinterpret_cast<uint4*>(t1 + 384) = H;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->pfull[st]);
        }
    } else {
        // =============================== dP epilogue: Adam + clamp on the 64 x k slice of P ===============================
        const int q = warp & 3;
        for (int sub = 0; sub < nsub; ++sub) {
            const int dbuf = sub & 1;
            mbar_wait(&S->d3full[dbuf], (sub >> 1) & 1);
            tc_fence_after_sync();
            uint32_t v[32];
            tmem_ld32(tbase + ((uint32_t)(q * 32) << 16) + kColD3 + dbuf * 32, v);
            tmem_wait_ld();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->d3empty[dbuf]);
            const int64_t m = (int64_t)(s0 + sub) * kMS + q * 16 + lane;       // M = 64 accumulator: lanes 32q + (0..15)
            if (lane < 16 && m < M) {
                float g[8];
#pragma unroll
                for (int c = 0; c < 8; ++c)   // columns: G^T.Q_h | G^T.Q_m | G^T.Q_l
                    g[c] = (__uint_as_float(v[c]) + __uint_as_float(v[8 + c])) + __uint_as_float(v[16 + c]);
                if (k == 8) {
                    if (dP_out != nullptr) {
                        reinterpret_cast<float4*>(dP_out + m * 8)[0] = make_float4(g[0], g[1], g[2], g[3]);
                        reinterpret_cast<float4*>(dP_out + m * 8)[1] = make_float4(g[4], g[5], g[6], g[7]);
                    }
                    if (adam.enabled) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float4 p4 = reinterpret_cast<float4*>(P + m * 8)[h];
                            float4 m4 = reinterpret_cast<float4*>(Pm + m * 8)[h];
                            float4 v4 = reinterpret_cast<float4*>(Pv + m * 8)[h];
                            p4.x = fminf(fmaxf(adam_apply(p4.x, g[4 * h + 0], m4.x, v4.x, adam), 0.f), 1.f);
                            p4.y = fminf(fmaxf(adam_apply(p4.y, g[4 * h + 1], m4.y, v4.y, adam), 0.f), 1.f);
                            p4.z = fminf(fmaxf(adam_apply(p4.z, g[4 * h + 2], m4.z, v4.z, adam), 0.f), 1.f);
                            p4.w = fminf(fmaxf(adam_apply(p4.w, g[4 * h + 3], m4.w, v4.w, adam), 0.f), 1.f);
                            reinterpret_cast<float4*>(P + m * 8)[h] = p4;
                            reinterpret_cast<float4*>(Pm + m * 8)[h] = m4;
                            reinterpret_cast<float4*>(Pv + m * 8)[h] = v4;
                        }
                    }
                } else {
                    for (int c = 0; c < k; ++c) {
                        const int64_t pi = m * k + c;
                        if (dP_out != nullptr) dP_out[pi] = g[c];
                        if (adam.enabled) {
                            float mm = Pm[pi], vv = Pv[pi];
                            const float pn = adam_apply(P[pi], g[c], mm, vv, adam);
                            P[pi] = fminf(fmaxf(pn, 0.f), 1.f);
                            Pm[pi] = mm;
                            Pv[pi] = vv;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < kWarpIssue; ++w) s += S->lossred[w];
        loss_part[blockIdx.x] = s;
    }
    if (warp == kWarpIssue) tmem_dealloc<512>(tbase);
}

// =================================================================================================================
// host launcher
// =================================================================================================================
bool dec_tc_supported(int B, int k) {
    const int nblk = (B + 127) / 128;
    const size_t fixed = (size_t)nblk * (kQBlkBytes + 1024) + kPStages * kPTileBytes + sizeof(DecSmem) + 128;
    return k <= 8 && nblk <= 8 && fixed + 3 * (size_t)kGtBytes <= (size_t)kMaxDynSmem;
}

#ifdef NADM_TIMELINE
extern "C" int nadm_debug_timeline(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(g_timeline));
}
#endif

int launch_dec_tc(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int B, int64_t M,
                  const float* Q, float* dQ, int q_ld, int q_off, int k, float* P, float* Pm, float* Pv,
                  const nadm_adam_t* adam, float* dP_out, float* loss, float* ws, size_t ws_bytes, cudaStream_t st) {
    const bool want_loss = loss != nullptr;
    const int nblk = (B + 127) / 128;
    const int TS = (int)((M + kMS - 1) / kMS);
    const int ncta = std::min(TS, sm_count());
    const size_t fixed = (size_t)nblk * (kQBlkBytes + 1024) + kPStages * kPTileBytes + sizeof(DecSmem) + 128;
    const int ngt = (fixed + 4 * (size_t)kGtBytes <= (size_t)kMaxDynSmem) ? 4 : 3;
    const size_t smem = fixed + (size_t)ngt * kGtBytes;
    NADM_REQUIRE((size_t)ncta * ((size_t)B * 8 + 1) * sizeof(float) <= ws_bytes, "workspace too small for decoder_step");
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(dec_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(dec_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dec_tc)");
        attr = true;
    }
    float* dQpart = ws;
    float* loss_part = ws + (size_t)ncta * B * 8;
    if (want_loss)
        dec_tc_kernel<true><<<ncta, kDecThreads, smem, st>>>(packed, pitch, row_idx, row0, B, M, Q, q_ld, q_off, k, P, Pm, Pv,
                                                            make_adam(adam), dP_out, dQpart, loss_part, TS, ngt);
    else
        dec_tc_kernel<false><<<ncta, kDecThreads, smem, st>>>(packed, pitch, row_idx, row0, B, M, Q, q_ld, q_off, k, P, Pm,
                                                             Pv, make_adam(adam), dP_out, dQpart, loss_part, TS, ngt);
    NADM_CHECK_LAUNCH("dec_tc_kernel");
    return launch_reduce_parts(dQpart, ncta, B, 8, k, dQ, q_ld, q_off, 1.0f, want_loss ? loss_part : nullptr, loss, st);
}

}  // namespace nadm
