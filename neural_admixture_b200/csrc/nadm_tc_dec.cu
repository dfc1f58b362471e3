// Fused decoder step on the 5th-generation tensor cores (tcgen05), sm_100a.  One head, k <= 16 (KH = 1: k <= 8; KH = 2:
// two halves of 8 components each, see "two component halves" below).
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
// Warp roles (WGS = 3: 640 threads; WGS = 4: 768): 0 .. 4 WGS - 1 the compute warpgroups (unit u -> warpgroup u % WGS,
// raw / G slot u % SLOTS), then: tensor memory allocation only, MMA issuer A (MMA2 of a unit, then MMA1 of the unit that
// takes over its slot), MMA issuer B (MMA3), P-tile producer, and four dP/Adam epilogue warps.  Pipelines are mbarrier based; tcgen05.commit frees
// operand buffers.
#include "nadm_common.cuh"
#include "nadm_tc.cuh"

#include <cuda_bf16.h>

namespace nadm {
using namespace tc;

constexpr int kMS = 64;                       // SNPs per sub-tile
constexpr int kGtBytes = 2 * 128 * kMS * 2;   // G^T tile: two bf16 terms x 128 rows x 64 SNPs = 32 KB
// Operand tiles hold, per 8-row (8-SNP) group, the 16-byte chunks (8 components as bf16) of ONE component half after the
// other.  Q: [h | m | l] per half.  P: KH = 1: [h | l | m | h]; KH = 2: [h | l | m | h | m'] per half, m' = the middle
// term ROUNDED to nearest (not truncated), read only by MMA2.
// Two component halves (KH = 2, heads with 9 <= k <= 16): raw = Q_a P_a^T + Q_b P_b^T (MMA1 accumulates the second half
// onto the first), dP_a | dP_b are two accumulators of MMA3, and dQ = G . [P_h | P_m'] per half with N = 16 instead of
// N = 32 — the exact three-term form would need 2 x 32 tensor-memory columns per row block, which 512 columns do not
// hold beside the raw / G slots; h + m' carries P to 2^-17 (unbiased), and dQ is a sum over all the SNPs of the CTA.
__host__ __device__ constexpr int dec_pchunks(int KH) { return KH == 1 ? 4 : 5; }          // chunks per half in a P tile
__host__ __device__ constexpr int dec_ptile_bytes(int KH) { return 8 * 128 * dec_pchunks(KH) * KH; }   // 64 SNPs
__host__ __device__ constexpr int dec_qblk_bytes(int KH) { return 16 * 384 * KH; }                    // 128 rows
__host__ __device__ constexpr int dec_dq_cols(int KH) { return KH == 1 ? 32 : 16; }        // dQ accumulator columns per half
constexpr int kPStages = 3;
// Three compute warpgroups (4 warps each) work on units round-robin; a unit lives in one of SLOTS raw / G slots of 64
// tensor-memory columns (+ one G^T tile in shared memory per slot).  SLOTS = 4 > 3 warpgroups lets a warpgroup start
// its next unit while the issuers are still turning its previous slot around (dQ/dP MMAs, then the next raw).
// Tensor-memory columns: [0, 64 SLOTS) slots, then the dP accumulator(s) (32 each: one for SLOTS = 4, two for 3),
// then one 32-column dQ accumulator per row block:  SLOTS = 4: 256 + 32 + 32 nblk (nblk <= 7, B <= 896);
// SLOTS = 3: 192 + 64 + 32 nblk (nblk <= 8).
// WGS compute warpgroups (template parameter: 3, or 4 = one more warp per SM sub-partition to fill the issue slots the
// others leave while they wait for their slot's turn-around; 768 threads, 80 registers each)
constexpr int kMaxSlots = 4;
__host__ __device__ constexpr int dec_threads(int WGS) { return (4 * WGS + 4 + 4) * 32; }
__host__ __device__ constexpr int dec_nd3(int SLOTS) { return SLOTS == 3 ? 2 : 1; }
constexpr uint32_t kIdesc1 = instr_desc(kAccF32, kFmtBF16, kFmtBF16, false, false, 128, kMS);
constexpr uint32_t kIdesc2 = instr_desc(kAccF32, kFmtBF16, kFmtBF16, false, true, 128, 32);
constexpr uint32_t kIdesc2h = instr_desc(kAccF32, kFmtBF16, kFmtBF16, false, true, 128, 16);
constexpr uint32_t kIdesc3 = instr_desc(kAccF32, kFmtBF16, kFmtBF16, true, true, 64, 24);
constexpr float kLog2Clamp = -144.26950408889634f;   // -100 / ln 2: torch clamps log at -100 (BCELoss)

#ifdef NADM_TIMELINE
__device__ long long g_timeline[8][512];
#define TL(row, idx) do { if (blockIdx.x == 0 && (idx) < 512) g_timeline[row][idx] = clock64(); } while (0)
#else
#define TL(row, idx) do { } while (0)
#endif

struct DecSmem {
    uint64_t d1full[kMaxSlots], gready[kMaxSlots], gtfree[kMaxSlots], pfull[kPStages], pempty[kPStages], d3full[2],
        d3empty[2], alldone;
    uint32_t tmem_base, magic, magic21, pad_;
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

// Process 16 consecutive SNPs of one row: v[] holds raw on entry; on exit hi[] / lo[] hold the bf16x2-packed split of G.
// w: the 16 2-bit codes (missing cleared).  Loss accumulators in log2 units, kept apart for x in {0, 1} (acc_hom) and
// x = 1/2 (acc_het, weight 1/2); kLoss = false: gradients only.
//
// Fast path, taken when every R (1 - R) of the 16 values (prod[], computed by the caller from the unclamped raw) is at
// least kProdFast = 2^-15: then 0 < raw < 1 (no clamp, mask = 1), the 1e-12 floor and the -100 log clamp cannot bind,
// and each BCE argument t (R, 1 - R or R (1 - R)) lies in [2^-15, 1].  The loss is accumulated as PRODUCTS of the
// RECIPROCAL arguments, which the gradient already provides (x in {0,1}: 1/t = |G|; x = 1/2: 1/t = inv), 8 elements
// per product (<= 2^120: no overflow), so 16 elements cost 4 logs and one multiply each instead of 16 logs:
// sum_j log t_j = -log prod_j (1 / t_j).
constexpr float kProdFast = 3.0517578125e-05f;
// General path (kGeneral): any value — raw may exceed 1 by rounding (clamp + inclusive mask of the clamp backward),
// R (1 - R) may fall below the 1e-12 floor of BCELoss' backward, the log may hit torch's -100 clamp — with torch's
// formulas evaluated literally and one log per element, in the SAME packed form as the fast path (+ ~6 instructions
// per element).  It is chosen per WARP, not per lane (thread = row: in late training, when restrict_P has pinned many
// allele frequencies at exactly 0 / 1 and Q is concentrated, nearly every warp holds a row that needs it, and a
// divergent warp pays for every path any of its lanes takes): bench.py's late_training leg measures that regime.
// The fp32 arithmetic of the fast path is issued as PACKED pairs (sm_100 FFMA2 / FADD2 / FMUL2 via __ffma2_rn & co.:
// one issue slot per two elements, results bit-identical to the scalar .rn instructions; negation / |.| fold into
// operand modifiers).  The kernel is bound by instruction issue, not by the fp32 pipe.
template <bool kLoss, bool kGeneral = false>
__device__ __forceinline__ void decode16_fast(const uint32_t (&v)[16], const float2 (&prod_in)[8], uint32_t w,
                                              uint32_t magic, uint32_t magic21, uint32_t (&hi)[8], uint32_t (&lo)[8],
                                              float& acc_hom, float& acc_het) {
    const uint32_t wh = w >> 16;
    float p_hom = 1.0f, p_het = 1.0f;
#pragma unroll
    for (int j2 = 0; j2 < 8; ++j2) {
        const int j = 2 * j2;
        const uint32_t wsrc = (j < 8) ? w : wh;
        const int sh = 2 * (j & 7);                                      // element j: bits sh, sh+1; j+1: sh+2, sh+3
        const float2 raw_u = make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
        // general: R = clamp(raw, 0, 1) (raw >= 0 always: every term of the split Q, P is), floor under R (1 - R)
        const float2 raw = kGeneral ? make_float2(fminf(raw_u.x, 1.0f), fminf(raw_u.y, 1.0f)) : raw_u;
        const float2 prod = kGeneral ? __ffma2_rn(make_float2(-raw.x, -raw.y), raw, raw) : prod_in[j2];
        const float2 inv = kGeneral ? make_float2(rcp_approx(fmaxf(prod.x, 1e-12f)), rcp_approx(fmaxf(prod.y, 1e-12f)))
                                    : make_float2(rcp_approx(prod.x), rcp_approx(prod.y));
        // code * 4^(j&7) as an exact float via the 2^23 magic constant (kept in a register: one LOP3 per element).  The
        // odd element's field sits two bits higher: OR-ing it into 2^21 (ulp 1/4) instead of 2^23 (ulp 1) gives it the
        // SAME scale as the even element, so the pair shares one multiplier (an immediate of the packed FFMA2).
        const float2 fm = make_float2(__uint_as_float((wsrc & (3u << sh)) | magic),
                                      __uint_as_float((wsrc & (12u << sh)) | magic21));
        const float2 f = __fadd2_rn(fm, make_float2(-8388608.0f, -2097152.0f));
        const float cs = -0.5f / (float)(1 << sh);
        const float2 num = __ffma2_rn(f, make_float2(cs, cs), raw);      // R - x
        float2 g = __fmul2_rn(num, inv);
        if (kGeneral) {                                               // clamp backward: inclusive mask 0 <= raw <= 1
            g.x = (raw_u.x <= 1.0f) ? g.x : 0.0f;
            g.y = (raw_u.y <= 1.0f) ? g.y : 0.0f;
        }
        if (kLoss && !kGeneral) {
            // x in {0,1}: |G| = 1 / (1 - |R - x|), the reciprocal of the BCE argument;  x = 1/2: inv = 1 / (R (1 - R))
            if ((wsrc >> sh) & 1u) p_het *= inv.x;
            else p_hom *= fabsf(g.x);
            if ((wsrc >> sh) & 4u) p_het *= inv.y;
            else p_hom *= fabsf(g.y);
        }
        if (kLoss && kGeneral) {
            // one log per element with torch's -100 clamp: x = 0: log(1 - R);  x = 1: log R (from R itself: 1 - |R - 1|
            // would lose an R below 2^-24);  x = 1/2: weight 1/2 on log(R (1 - R))
            const bool het0 = (wsrc >> sh) & 1u, het1 = (wsrc >> sh) & 4u;
            const bool one0 = (wsrc >> sh) & 2u, one1 = (wsrc >> sh) & 8u;
            const float l0 = fmaxf(lg2_approx(het0 ? prod.x : (one0 ? raw.x : 1.0f - raw.x)), kLog2Clamp);
            const float l1 = fmaxf(lg2_approx(het1 ? prod.y : (one1 ? raw.y : 1.0f - raw.y)), kLog2Clamp);
            acc_hom += het0 ? 0.0f : l0;
            acc_het += het0 ? l0 : 0.0f;
            acc_hom += het1 ? 0.0f : l1;
            acc_het += het1 ? l1 : 0.0f;
        }
        const uint32_t h = pack_bf16x2(g.x, g.y);
        hi[j2] = h;
        const float2 l = __fadd2_rn(g, make_float2(-__uint_as_float(h << 16), -__uint_as_float(h & 0xFFFF0000u)));
        lo[j2] = pack_bf16x2(l.x, l.y);
        if (kLoss && !kGeneral && (j2 & 3) == 3) {
            acc_hom -= lg2_approx(p_hom);
            acc_het -= lg2_approx(p_het);
            p_hom = 1.0f;
            p_het = 1.0f;
        }
    }
}

__device__ GridBar g_bar_dec;

template <bool kLoss, int SLOTS, int WGS, int KH>
__global__ void __launch_bounds__(dec_threads(WGS), 1)
dec_tc_kernel(const uint8_t* __restrict__ packed, int64_t pitch, const int64_t* __restrict__ row_idx, int64_t row0, int B,
              int64_t M, const float* __restrict__ Q, int q_ld, int q_off, int k, float* __restrict__ P,
              float* __restrict__ Pm, float* __restrict__ Pv, AdamCoef adam_in, float* __restrict__ dP_out,
              float* __restrict__ dQpart, float* __restrict__ loss_part, int TS, float* __restrict__ dQ_out,
              float* __restrict__ loss_out) {
    // dQ_out != NULL: the CTAs meet at a grid barrier after writing their partials and reduce them themselves
    // (cooperative launch; loss_out += the summed loss); dQ_out == NULL: a separate reduce_parts_kernel follows.
    pdl_prologue();
    constexpr int kWGs = WGS, kWarpIssueA1 = 4 * WGS, kWarpIssueA2 = kWarpIssueA1 + 1, kWarpIssueB = kWarpIssueA1 + 2,
                  kWarpProd = kWarpIssueA1 + 3;
    constexpr int kSlots = SLOTS, kWarpIssue = kWarpIssueA1;
    // unit u is decoded by warpgroup u % WGS in slot u % SLOTS; its wait for raw(u) tests the parity of the slot's phase
    // u / SLOTS, which is only unambiguous if raw(u - SLOTS) is known to be complete: the warpgroup has seen raw(u - WGS),
    // MMA1s complete in unit order, so WGS <= SLOTS is required (4 warpgroups over 3 slots: the fourth's first wait
    // passes on the fresh barrier and it decodes the slot together with the first).
    static_assert(WGS <= SLOTS, "compute warpgroups must not outnumber the raw / G slots");
    const AdamCoef adam = adam_resolve(adam_in);
    constexpr int ngt = SLOTS;   // one G^T tile per slot (see the issuer warps for why not more)
    constexpr int kND3 = (KH == 2) ? 1 : dec_nd3(SLOTS), kColD3 = 64 * SLOTS, kColD2 = kColD3 + 32 * kND3 * KH;
    constexpr int kPTileBytes = dec_ptile_bytes(KH), kQBlkBytes = dec_qblk_bytes(KH), kDQ = dec_dq_cols(KH);
    constexpr int kQGrp = 384 * KH, kPGrp = 128 * dec_pchunks(KH) * KH;   // bytes of one 8-row (8-SNP) group of a Q / P tile
    constexpr int kPHalf = 128 * dec_pchunks(KH);                          // bytes between the two halves inside a P group
    extern __shared__ __align__(1024) uint8_t smem[];
    const int nblk = (B + 127) / 128;
    uint8_t* QA = smem;                                   // nblk x 6 KB x KH : bf16 h/m/l of Q (MMA1 A K-major, MMA3 B MN-major)
    uint8_t* PT = QA + nblk * kQBlkBytes;                 // kPStages P sub-tiles (MMA1 B K-major, MMA2 B MN-major)
    uint8_t* GT = PT + kPStages * kPTileBytes;            // ngt x 32 KB : G^T tiles
    int64_t* rowoff = reinterpret_cast<int64_t*>(GT + ngt * kGtBytes);
    DecSmem* S = reinterpret_cast<DecSmem*>(rowoff + nblk * 128);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s0 = (int)(((int64_t)TS * blockIdx.x) / gridDim.x), s1 = (int)(((int64_t)TS * (blockIdx.x + 1)) / gridDim.x);
    const int nsub = s1 - s0, U = nsub * nblk;
    if (tid == 0) TL(0, 500);                                           // (timeline builds) kernel entry

    // ---------------- one-time setup: row offsets, Q operands, barriers, tensor memory ----------------
    for (int b = tid; b < nblk * 128; b += blockDim.x)
        rowoff[b] = (b < B) ? ((row_idx != nullptr) ? row_idx[b] : (row0 + b)) * pitch : -1;
    for (int b = tid; b < nblk * 128; b += blockDim.x) {
#pragma unroll
        for (int hf = 0; hf < KH; ++hf) {
            float q[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) q[c] = (b < B && 8 * hf + c < k) ? Q[(int64_t)b * q_ld + q_off + 8 * hf + c] : 0.f;
            uint4 H, Mm, L;
            split3_row(q, H, Mm, L);
            uint8_t* a = QA + (b & 7) * 16 + (b >> 3) * kQGrp + hf * 384;
            *reinterpret_cast<uint4*>(a) = H;
            *reinterpret_cast<uint4*>(a + 128) = Mm;
            *reinterpret_cast<uint4*>(a + 256) = L;
        }
    }
    if (tid == 0) {
        for (int i = 0; i < kMaxSlots; ++i) {
            mbar_init(&S->d1full[i], 1);
            mbar_init(&S->gready[i], 4);
            mbar_init(&S->gtfree[i], 1);
        }
        for (int i = 0; i < kPStages; ++i) { mbar_init(&S->pfull[i], 1); mbar_init(&S->pempty[i], 2); }
        for (int i = 0; i < 2; ++i) { mbar_init(&S->d3full[i], 1); mbar_init(&S->d3empty[i], 4); }
        mbar_init(&S->alldone, 1);
        S->magic = 0x4B000000u;
        S->magic21 = 0x4A000000u;
        mbar_init_fence();
    }
    if (warp == kWarpIssue) tmem_alloc<512>(&S->tmem_base);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = S->tmem_base;
    if (tid == 0) TL(0, 501);                                           // setup done
#ifdef NADM_KO_MMA   // knock-out measurement build: no tensor-core work; every raw value is 0.5 (decode-only timing)
    if (warp < 4) {
        uint32_t h[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) h[j] = 0x3F000000u;
        for (int c0 = 0; c0 < 64 * SLOTS; c0 += 16) tmem_st16(tbase + ((uint32_t)(warp * 32) << 16) + c0, h);
        tmem_wait_st();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
#endif

    if (warp < kWarpIssue) {
        // =============================== compute warpgroups ===============================
        const int wg = warp >> 2, q = warp & 3;
        const int rb = q * 32 + lane;                                   // row inside the block = tensor-memory lane
        const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
        float acc_hom = 0.f, acc_het = 0.f;
        uint4 gw = make_uint4(0u, 0u, 0u, 0u);
        auto load_codes = [&](int blk_, int sub_) {
            const int64_t ro = rowoff[blk_ * 128 + rb];
            const int64_t off = (int64_t)(s0 + sub_) * (kMS / 4);
            uint4 r = make_uint4(0u, 0u, 0u, 0u);
            if (ro >= 0 && off + 16 <= pitch) r = *reinterpret_cast<const uint4*>(packed + ro + off);
            return r;
        };
        const uint32_t magic21 = S->magic21;
        const uint32_t magic = S->magic;   // 2^23 as bits, read from shared memory so that it stays in a REGISTER:
                                           // (w & field) | magic is then one LOP3 (an immediate would cost a second)
        // unit counters, advanced incrementally by kWGs units (no divisions in the loop):
        //   u = sub * nblk + blk = uq * kSlots + slot;  (nblk_, nsub_) belong to the prefetched unit u + kWGs
        int blk = wg % nblk, sub = wg / nblk, slot = wg % kSlots, uq = wg / kSlots;
        int nblk_ = blk, nsub_ = sub;
        auto advance = [&](int& b_, int& s_) {
            b_ += kWGs;
            while (b_ >= nblk) { b_ -= nblk; ++s_; }
        };
        if (wg < U) gw = load_codes(blk, sub);
        for (int u = wg; u < U; u += kWGs) {
            const int g = slot;                                         // G^T tile of this unit = its slot
            const uint32_t cw[4] = {clear_missing(gw.x), clear_missing(gw.y), clear_missing(gw.z), clear_missing(gw.w)};
            advance(nblk_, nsub_);
            if (u + kWGs < U) gw = load_codes(nblk_, nsub_);
            const bool active = blk * 128 + q * 32 < B;                 // warp-uniform: any real row in this warp
            if (rb == 0) TL(0, u);                                      // WG starts waiting for raw(u)
            mbar_wait(&S->d1full[slot], (uq & 1));
            tc_fence_after_sync();
            if (rb == 0) TL(1, u);                                      // raw(u) seen
#ifdef NADM_SKIPDECODE
            if (false) {
#else
            if (active) {
#endif
                mbar_wait(&S->gtfree[g], (uq & 1) ^ 1);
                if (rb == 0) TL(6, u);                                  // G^T buffer free
                uint8_t* gt = GT + g * kGtBytes + (rb & 7) * 16 + (rb >> 3) * 1024;
                // 16 SNPs at a time: raw columns [16c, 16c+16) of the slot -> G hi / lo
                auto decode_group = [&](int c, const uint32_t (&v)[16]) {
                    uint32_t hi[8], lo[8];
                    const uint32_t w = (c & 2) ? ((c & 1) ? cw[3] : cw[2]) : ((c & 1) ? cw[1] : cw[0]);
                    float2 prod[8];                                   // R (1 - R), packed pairs (FFMA2)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 r2 = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                        prod[j] = __ffma2_rn(make_float2(-r2.x, -r2.y), r2, r2);
                    }
                    // min over the 16 products as a 3-input tree (depth 3) instead of a chain of 8 dependent FMNMX3
                    const float t0 = fminf(fminf(prod[0].x, prod[0].y), prod[1].x), t1 = fminf(fminf(prod[1].y, prod[2].x), prod[2].y);
                    const float t2 = fminf(fminf(prod[3].x, prod[3].y), prod[4].x), t3 = fminf(fminf(prod[4].y, prod[5].x), prod[5].y);
                    const float t4 = fminf(fminf(prod[6].x, prod[6].y), prod[7].x);
                    const float mn = fminf(fminf(fminf(t0, t1), t2), fminf(fminf(t3, t4), prod[7].y));
                    // one path per warp: all 32 rows in the fast range -> fast, else the general form for all of them
                    if (__all_sync(0xffffffffu, mn >= kProdFast))
                        decode16_fast<kLoss, false>(v, prod, w, magic, magic21, hi, lo, acc_hom, acc_het);
                    else
                        decode16_fast<kLoss, true>(v, prod, w, magic, magic21, hi, lo, acc_hom, acc_het);
#ifndef NADM_KO_MMA   // (knock-out build: raw stays the constant written at setup, so that the decode keeps its fast path)
                    tmem_st8(tlane + slot * 64 + c * 16, hi);          // G hi / lo overwrite their own raw columns
                    tmem_st8(tlane + slot * 64 + c * 16 + 8, lo);
#endif
                    *reinterpret_cast<uint4*>(gt + (2 * c) * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(gt + (2 * c + 1) * 128) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    *reinterpret_cast<uint4*>(gt + 16384 + (2 * c) * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<uint4*>(gt + 16384 + (2 * c + 1) * 128) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                };
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[16];
                    tmem_ld16(tlane + slot * 64 + c * 16, v);
                    tmem_wait_ld();
                    decode_group(c, v);
                }
                tmem_wait_st();
                fence_async_smem();
            }
            tc_fence_before_sync();
            if (rb == 0) TL(2, u);                                      // G(u) written
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->gready[slot]);               // one arrival per warp
            blk = nblk_;
            sub = nsub_;
            slot += kWGs;
            if (slot >= kSlots) { slot -= kSlots; ++uq; }              // (kWGs <= kSlots: wraps at most once)
        }
        // ---- loss partial of this CTA (log2 units -> nats), dQ partial from tensor memory ----
        float l = -(acc_hom + 0.5f * acc_het) * 0.6931471805599453f;
        l = warp_sum(l);
        if (lane == 0) S->lossred[warp] = l;
        if (tid == 0) TL(0, 502);                                       // this warpgroup's units done
        mbar_wait(&S->alldone, 0);
        tc_fence_after_sync();
        if (tid == 0) TL(0, 503);                                       // all MMA2s complete
        for (int blk = wg; blk < nblk; blk += kWGs) {
            uint32_t v[32];
            tmem_ld32(tlane + kColD2 + blk * 32, v);      // KH = 1: one 32-column accumulator; KH = 2: two of 16 columns
            tmem_wait_ld();
            const int b = blk * 128 + rb;
            if (b < B) {
                float* out = dQpart + ((int64_t)blockIdx.x * B + b) * (8 * KH);
#pragma unroll
                for (int hf = 0; hf < KH; ++hf) {
                    float o[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        // KH = 1 columns: G.P_h | G.P_l | G.P_m | G.P_h (duplicate, unused);  KH = 2, per half: G.P_h | G.P_m'
                        const float sum = (KH == 1) ? (__uint_as_float(v[c]) + __uint_as_float(v[16 + c])) + __uint_as_float(v[8 + c])
                                                    : __uint_as_float(v[16 * hf + c]) + __uint_as_float(v[16 * hf + 8 + c]);
                        o[c] = (U > 0) ? sum : 0.f;
                    }
                    reinterpret_cast<float4*>(out + 8 * hf)[0] = make_float4(o[0], o[1], o[2], o[3]);
                    reinterpret_cast<float4*>(out + 8 * hf)[1] = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
        }
    } else if (warp == kWarpIssueA1) {
        // (this warp only allocates and frees tensor memory: MMA1 moved to issuer A below)
    } else if (warp == kWarpIssueA2) {
        // =============================== MMA issuer A: dQ_blk += G . [P_h | P_l | P_m | P_h] (MMA2), then raw of the
        // unit that takes over the slot (MMA1) ===============================
        // MMA2(u) reads the slot's G from tensor memory and MMA1(u + SLOTS) overwrites the same columns with the next
        // raw.  Both are issued by THIS thread, and tcgen05.mma instructions of one thread execute in issue order, so
        // no barrier is needed between them: the slot's turn-around (gready -> raw of its next unit) is the issue and
        // execution time of 12 MMAs.  With the two on different warps (earlier builds) the tcgen05.commit ->
        // mbarrier -> polling-warp hop in between made the turn-around about as long as the stagger between the three
        // warpgroups, which then waited for raw ~20 % of the time (ncu warp-state samples at their d1full wait).
        const uint32_t qa = smem_u32(QA), pt = smem_u32(PT);
        // Q chunks [h m l] (128 B apart, 8-row groups 384 B apart); P chunks [h l m h] (8-row groups 512 B apart).
        // A K=16 instruction multiplies two chunk pairs: start address = first chunk, LBO = distance to the second.
        const uint64_t A_hm = smem_desc(qa, 128, kQGrp), A_hl = smem_desc(qa, 256, kQGrp), A_ml = smem_desc(qa + 128, 128, kQGrp);
        const uint64_t B_hm = smem_desc(pt, 256, kPGrp), B_mh = smem_desc(pt + 256, 128, kPGrp);
        const uint64_t B_lh = smem_desc(pt + 128, 256, kPGrp), B_lm = smem_desc(pt + 128, 128, kPGrp);
        // MMA2's B operand, MN-major: KH = 1: all four chunks [h l m h] (N = 32); KH = 2: chunks [h m'] of a half (N = 16)
        const uint64_t B2 = smem_desc(pt + (KH == 1 ? 0 : 384), kPGrp, 128);
        const uint32_t leader = elect_one() ? 1u : 0u;                  // the one lane that executes the MMAs / commits
        int l_blk = 0, l_stage = 0, l_phase = 0, l_slot = 0;
        auto issue_mma1 = [&]() {
            if (l_blk == 0) mbar_wait(&S->pfull[l_stage], l_phase);     // first unit of a sub-tile: its P tile must be there
            tc_fence_after_sync();
            const uint32_t ao = (uint32_t)(l_blk * (kQBlkBytes >> 4)), bo = (uint32_t)(l_stage * (kPTileBytes >> 4));
            const uint32_t d = tbase + l_slot * 64;
#pragma unroll
            for (int hf = 0; hf < KH; ++hf) {                            // the second component half accumulates onto the first
                const uint32_t ah = ao + (uint32_t)(hf * (384 >> 4)), bh = bo + (uint32_t)(hf * (kPHalf >> 4));
                mma_f16_ss_p(d, desc_add(A_hm, ah), desc_add(B_hm, bh), kIdesc1, hf ? 1u : 0u, leader);   // h.h + m.m
                mma_f16_ss_p(d, desc_add(A_hm, ah), desc_add(B_mh, bh), kIdesc1, 1u, leader);             // h.m + m.h
                mma_f16_ss_p(d, desc_add(A_hl, ah), desc_add(B_lh, bh), kIdesc1, 1u, leader);             // h.l + l.h
                mma_f16_ss_p(d, desc_add(A_ml, ah), desc_add(B_lm, bh), kIdesc1, 1u, leader);             // m.l + l.m
            }
            mma_commit_p(&S->d1full[l_slot], leader);
            if (l_blk == nblk - 1) mma_commit_p(&S->pempty[l_stage], leader); // MMA1's reads of the P stage are issued
            if (++l_slot == kSlots) l_slot = 0;
            if (++l_blk == nblk) {
                l_blk = 0;
                if (++l_stage == kPStages) { l_stage = 0; l_phase ^= 1; }
            }
        };
        for (int l = 0; l < kSlots && l < U; ++l) issue_mma1();         // fill the slots
        int blk = 0, sub = 0, slot = 0, slot_phase = 0, stage = 0;
        for (int u = 0; u < U; ++u) {
            if (lane == 0) TL(3, u);                                    // issuer starts waiting for G(u)
            mbar_wait(&S->gready[slot], slot_phase);
            tc_fence_after_sync();
            if (lane == 0) TL(4, u);                                    // issuer saw G(u)
            const uint64_t b2 = desc_add(B2, (uint32_t)(stage * (kPTileBytes >> 4)));
            const uint32_t a2 = tbase + slot * 64;
            const uint32_t acc2 = sub > 0 ? 1u : 0u;
#pragma unroll
            for (int hf = 0; hf < KH; ++hf) {
                const uint32_t d2 = tbase + kColD2 + (blk * KH + hf) * kDQ;
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int t = 0; t < 2; ++t)
                        mma_f16_ts_p(d2, a2 + c * 16 + t * 8,
                                     desc_add(b2, (uint32_t)(c * 2 * (kPGrp >> 4) + hf * (kPHalf >> 4))),
                                     KH == 1 ? kIdesc2 : kIdesc2h, (c + t) ? 1u : acc2, leader);
            }
            if (blk == nblk - 1) mma_commit_p(&S->pempty[stage], leader); // MMA2's reads of the P stage are issued
            if (u + kSlots < U) issue_mma1();                           // raw of unit u + SLOTS into the slot just consumed
            if (lane == 0) TL(5, u);                                    // issuer done with unit u
            if (++slot == kSlots) { slot = 0; slot_phase ^= 1; }
            if (++blk == nblk) {
                blk = 0;
                ++sub;
                if (++stage == kPStages) stage = 0;
            }
        }
        mma_commit_p(&S->alldone, leader);
        __syncwarp();
    } else if (warp == kWarpIssueB) {
        // =============================== MMA issuer B: dP_sub += G^T . [Q_h | Q_m | Q_l] (MMA3) ===============================
        // A = the shared G^T tile of the unit's slot, MN-major; B = the Q block re-read MN-major; only K steps holding
        // real rows are issued for the last row block.
        const uint64_t A3 = smem_desc(smem_u32(GT), 1024, 128), B3 = smem_desc(smem_u32(QA), kQGrp, 128);
        const int nks_last = min(8, (B - (nblk - 1) * 128 + 15) / 16);
        const uint32_t leader = elect_one() ? 1u : 0u;
        int blk = 0, slot = 0, slot_phase = 0, dbuf = 0, d3_phase = 1;
        for (int u = 0; u < U; ++u) {
            mbar_wait(&S->gready[slot], slot_phase);
            if (blk == 0) mbar_wait(&S->d3empty[dbuf], d3_phase);
            tc_fence_after_sync();
            {
                const uint64_t a3 = desc_add(A3, (uint32_t)(slot * (kGtBytes >> 4))),
                               b3 = desc_add(B3, (uint32_t)(blk * (kQBlkBytes >> 4)));
                const uint32_t acc3 = blk > 0 ? 1u : 0u;
#pragma unroll
                for (int hf = 0; hf < KH; ++hf) {
                    const uint32_t d3 = tbase + kColD3 + (dbuf * KH + hf) * 32;
                    const uint64_t b3h = desc_add(b3, (uint32_t)(hf * (384 >> 4)));
                    if (blk != nblk - 1 || nks_last == 8) {
#pragma unroll
                        for (int t = 0; t < 2; ++t)
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks)
                                mma_f16_ss_p(d3, desc_add(a3, (uint32_t)(t * 1024 + ks * 128)),
                                             desc_add(b3h, (uint32_t)(ks * 2 * (kQGrp >> 4))), kIdesc3, (t + ks) ? 1u : acc3, leader);
                    } else {
                        for (int t = 0; t < 2; ++t)
                            for (int ks = 0; ks < nks_last; ++ks)
                                mma_f16_ss_p(d3, desc_add(a3, (uint32_t)(t * 1024 + ks * 128)),
                                             desc_add(b3h, (uint32_t)(ks * 2 * (kQGrp >> 4))), kIdesc3, (t + ks) ? 1u : acc3, leader);
                    }
                }
                mma_commit_p(&S->gtfree[slot], leader);
                if (blk == nblk - 1) mma_commit_p(&S->d3full[dbuf], leader);
            }
            if (lane == 0) TL(7, u);                                    // issuer B done with unit u
            if (++slot == kSlots) { slot = 0; slot_phase ^= 1; }
            if (++blk == nblk) {
                blk = 0;
                if (++dbuf == kND3) { dbuf = 0; d3_phase ^= 1; }
            }
        }
    } else if (warp == kWarpProd) {
        // =============================== P sub-tile producer ===============================
        for (int sub = 0; sub < nsub; ++sub) {
            const int st = sub % kPStages;
            float p[2][8 * KH];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t m = (int64_t)(s0 + sub) * kMS + lane + 32 * e;
#pragma unroll
                for (int c = 0; c < 8 * KH; ++c) p[e][c] = 0.f;
                if (m < M) {
                    if (k == 8 * KH) {
#pragma unroll
                        for (int c4 = 0; c4 < 2 * KH; ++c4) {
                            const float4 a = reinterpret_cast<const float4*>(P + m * (8 * KH))[c4];
                            p[e][4 * c4] = a.x; p[e][4 * c4 + 1] = a.y; p[e][4 * c4 + 2] = a.z; p[e][4 * c4 + 3] = a.w;
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 8 * KH; ++c)
                            if (c < k) p[e][c] = P[m * k + c];
                    }
                }
            }
            mbar_wait_relaxed(&S->pempty[st], ((sub / kPStages) & 1) ^ 1, 256);
            uint8_t* tile = PT + st * kPTileBytes;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ml = lane + 32 * e;
#pragma unroll
                for (int hf = 0; hf < KH; ++hf) {
                    uint8_t* t1 = tile + (ml & 7) * 16 + (ml >> 3) * kPGrp + hf * kPHalf;
                    float ph[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) ph[c] = p[e][8 * hf + c];
                    uint4 H, Mm, L;
                    split3_row(ph, H, Mm, L);
                    *reinterpret_cast<uint4*>(t1) = H;
                    *reinterpret_cast<uint4*>(t1 + 128) = L;
                    *reinterpret_cast<uint4*>(t1 + 256) = Mm;
                    *reinterpret_cast<uint4*>(t1 + 384) = H;
                    if (KH == 2) {   // m' = bf16_rn(p - h): with h, carries p to 2^-17 (MMA2's two-term operand)
                        uint32_t mr[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float hfl = __uint_as_float(__float_as_uint(ph[c]) & 0xFFFF0000u);
                            const __nv_bfloat16 r = __float2bfloat16_rn(ph[c] - hfl);
                            mr[c] = (uint32_t)(*reinterpret_cast<const unsigned short*>(&r));
                        }
                        *reinterpret_cast<uint4*>(t1 + 512) = make_uint4(mr[0] | (mr[1] << 16), mr[2] | (mr[3] << 16),
                                                                         mr[4] | (mr[5] << 16), mr[6] | (mr[7] << 16));
                    }
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->pfull[st]);
        }
    } else {
        // =============================== dP epilogue: Adam + clamp on the 64 x k slice of P ===============================
        const int q = warp & 3;
        for (int sub = 0; sub < nsub; ++sub) {
            const int dbuf = sub % kND3;
            // one warp polls, the other three sleep on a named barrier.  The poller is the epilogue warp of the SM
            // sub-partition that hosts no busy issuer (warp ids 12..15 -> sub-partitions 0..3; 12 is idle once MMA1 and
            // MMA2 share an issuer): polling costs ~5 % of a sub-partition's issue slots, and the kernel runs at the
            // pace of the busiest one.
            constexpr int kPoller = kWarpProd + 1;
            if (warp == kPoller) mbar_wait_relaxed(&S->d3full[dbuf], (sub / kND3) & 1, 64);
            named_bar_sync(1, 128);
            tc_fence_after_sync();
            float g[8 * KH];
#pragma unroll
            for (int hf = 0; hf < KH; ++hf) {
                uint32_t v[32];
                tmem_ld32(tbase + ((uint32_t)(q * 32) << 16) + kColD3 + (dbuf * KH + hf) * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 8; ++c)   // columns: G^T.Q_h | G^T.Q_m | G^T.Q_l
                    g[8 * hf + c] = (__uint_as_float(v[c]) + __uint_as_float(v[8 + c])) + __uint_as_float(v[16 + c]);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->d3empty[dbuf]);
            const int64_t m = (int64_t)(s0 + sub) * kMS + q * 16 + lane;       // M = 64 accumulator: lanes 32q + (0..15)
            if (lane < 16 && m < M) {
                if (KH == 1 && k == 8) {
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
#pragma unroll
                    for (int c = 0; c < 8 * KH; ++c) {
                        if (c < k) {
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
    }
    if (tid == 0) TL(0, 504);                                           // this warp's dQ partial written
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < kWarpIssue; ++w) s += S->lossred[w];
        loss_part[blockIdx.x] = s;
        TL(0, 505);                                                     // every role done (incl. the last dP epilogue)
    }
    if (warp == kWarpIssue) tmem_dealloc<512>(tbase);
    if (dQ_out != nullptr) {
        // ---- fused reduction of the CTAs' dQ partials: this CTA sums outputs [n cta / nctas, n (cta + 1) / nctas) over all
        // CTAs (deterministic) ----
        grid_barrier(&g_bar_dec, gridDim.x);
        if (tid == 0) TL(0, 506);
        // Latency-bound: a work item = (output, one of 16 interleaved segments of the parts) holds all its loads in
        // flight (<= 10 for 148 CTAs); segment sums are combined in a fixed order.
        float* red = reinterpret_cast<float*>(GT);                      // (the G^T tiles are idle now)
        const int nparts = (int)gridDim.x, colsp = 8 * KH, nthr = (int)blockDim.x;
        const int64_t n = (int64_t)B * colsp;
        const int o0 = (int)((n * blockIdx.x) / gridDim.x), o1 = (int)((n * (blockIdx.x + 1)) / gridDim.x);
        for (int c0 = o0; c0 < o1; c0 += 128) {
            const int cnt = min(128, o1 - c0);
            for (int w = tid; w < cnt * 16; w += nthr) {
                const int o = c0 + (w >> 4), seg = w & 15;
                const float* src = dQpart + o;
                float acc = 0.f;
                if (nparts <= 160) {
                    float v[10];
#pragma unroll
                    for (int i = 0; i < 10; ++i) {
                        const int p = seg + 16 * i;
                        v[i] = (p < nparts) ? __ldcg(src + (int64_t)p * n) : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 10; ++i) acc += v[i];
                } else {
                    for (int p = seg; p < nparts; p += 16) acc += __ldcg(src + (int64_t)p * n);
                }
                red[w] = acc;
            }
            __syncthreads();
            for (int w = tid; w < cnt; w += nthr) {
                const int o = c0 + w, r = o / colsp, c = o - r * colsp;
                float t = 0.f;
#pragma unroll
                for (int sg = 0; sg < 16; ++sg) t += red[16 * w + sg];
                if (c < k) dQ_out[(int64_t)r * q_ld + q_off + c] = t;
            }
            __syncthreads();
        }
        if (kLoss && loss_out != nullptr && blockIdx.x == 0 && warp == 0) {
            double acc2 = 0.0;
            for (int q = lane; q < nparts; q += 32) acc2 += (double)__ldcg(loss_part + q);
            acc2 = warp_sum_d(acc2);
            if (lane == 0) *loss_out = (float)((double)*loss_out + acc2);
        }
        if (tid == 0) TL(0, 507);
    }
}

// =================================================================================================================
// host launcher
// =================================================================================================================
static size_t dec_fixed_smem(int nblk, int KH) {
    return (size_t)nblk * (dec_qblk_bytes(KH) + 1024) + (size_t)kPStages * dec_ptile_bytes(KH) + sizeof(DecSmem) + 128;
}
bool dec_tc_supported(int B, int k) {
    const int nblk = (B + 127) / 128, KH = k <= 8 ? 1 : 2;
    // tensor memory with 3 slots: 192 + dP (KH = 1: 2 x 32; KH = 2: 2 x 32) + dQ (32 nblk either way) <= 512
    return k <= 16 && nblk <= 8 && dec_fixed_smem(nblk, KH) + 3 * (size_t)kGtBytes <= (size_t)kMaxDynSmem;
}

#ifdef NADM_TIMELINE
extern "C" int nadm_debug_timeline(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(g_timeline));
}
#endif

template <bool kLoss, int SLOTS, int WGS, int KH>
static int dec_launch_one(int ncta, size_t smem, cudaStream_t st, const uint8_t* packed, int64_t pitch,
                          const int64_t* row_idx, int64_t row0, int B, int64_t M, const float* Q, int q_ld, int q_off,
                          int k, float* P, float* Pm, float* Pv, const AdamCoef& adam, float* dP_out, float* dQpart,
                          float* loss_part, int TS, float* dQ_out, float* loss_out) {
    static PerDeviceOnce once;                       // one per template instantiation
    bool* attr = once.slot();
    if (attr == nullptr || !*attr) {
        cudaError_t e = cudaFuncSetAttribute(dec_tc_kernel<kLoss, SLOTS, WGS, KH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kMaxDynSmem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dec_tc)");
        if (attr) *attr = true;
    }
    cudaError_t le;
    if (dQ_out != nullptr)   // fused reduction behind a grid barrier: cooperative launch
        le = launch_coop(dec_tc_kernel<kLoss, SLOTS, WGS, KH>, dim3(ncta), dim3(dec_threads(WGS)), smem, st, packed, pitch, row_idx,
                         row0, B, M, Q, q_ld, q_off, k, P, Pm, Pv, adam, dP_out, dQpart, loss_part, TS, dQ_out, loss_out);
    else
        le = launch_pdl(dec_tc_kernel<kLoss, SLOTS, WGS, KH>, dim3(ncta), dim3(dec_threads(WGS)), smem, st, packed, pitch, row_idx,
                        row0, B, M, Q, q_ld, q_off, k, P, Pm, Pv, adam, dP_out, dQpart, loss_part, TS, dQ_out, loss_out);
    if (le != cudaSuccess) return cuda_fail(le, "dec_tc_kernel");
    NADM_CHECK_LAUNCH("dec_tc_kernel");
    return NADM_OK;
}

// raw / G slots: 4 whenever tensor memory (nblk <= 7, i.e. B <= 896) and shared memory allow it, else 3.
// NADM_DEC_SLOTS=3 forces 3 (A/B measurements).
static int dec_pick_slots(int nblk, size_t fixed) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("NADM_DEC_SLOTS");
        forced = e ? atoi(e) : 0;
    }
    int slots = (nblk <= 7 && fixed + 4 * (size_t)kGtBytes <= (size_t)kMaxDynSmem) ? 4 : 3;
    if (forced == 3) slots = 3;
    return slots;
}

// compute warpgroups: NADM_DEC_WGS=3|4 forces one (A/B measurements).  Default, measured on one box
// (profiles/r2_decoder_wgs.txt): gradients only -> 4 (243 vs 248 us: a fourth warp per sub-partition fills issue slots
// the others leave while they wait for their slot's turn-around); with the loss value -> 3 (275 vs 282 us: at the 80
// registers per thread that 768 threads allow, the loss-evaluating decode loop is allocated worse).
static int dec_pick_wgs(bool want_loss) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("NADM_DEC_WGS");
        forced = (e != nullptr && (e[0] == '3' || e[0] == '4')) ? e[0] - '0' : 0;
    }
    return forced ? forced : (want_loss ? 3 : 4);
}

int launch_dec_tc(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int B, int64_t M,
                  const float* Q, float* dQ, int q_ld, int q_off, int k, float* P, float* Pm, float* Pv,
                  const nadm_adam_t* adam, float* dP_out, float* loss, float* ws, size_t ws_bytes, cudaStream_t st,
                  bool defer) {
    const bool want_loss = loss != nullptr;
    const int nblk = (B + 127) / 128;
    const int TS = (int)((M + kMS - 1) / kMS);
    const int ncta = std::min(TS, sm_count());
    const int KH = k <= 8 ? 1 : 2;                                   // component halves (9 <= k <= 16: two)
    const size_t fixed = dec_fixed_smem(nblk, KH);
    const int slots = KH == 2 ? 3 : dec_pick_slots(nblk, fixed);
    const size_t smem = fixed + (size_t)slots * kGtBytes;             // one G^T tile per slot
    NADM_REQUIRE((size_t)ncta * ((size_t)B * 8 * KH + 1) * sizeof(float) <= ws_bytes, "workspace too small for decoder_step");
    float* dQpart = ws;
    float* loss_part = ws + (size_t)ncta * B * 8 * KH;
    const AdamCoef ac = make_adam(adam);
    int rc;
    const int wgs = dec_pick_wgs(want_loss);
    const bool fused = gridbar_enabled();          // the CTAs sum their dQ partials themselves (opt-in, NADM_GRIDBAR=1)
#define NADM_DEC_GO(L, W, G, H)                                                                                        \
    rc = dec_launch_one<L, W, G, H>(ncta, smem, st, packed, pitch, row_idx, row0, B, M, Q, q_ld, q_off, k, P, Pm, Pv, ac, \
                                    dP_out, dQpart, loss_part, TS, fused ? dQ : nullptr, fused ? loss : nullptr)
#define NADM_DEC_GO_WG(L, W, H)                                                                                        \
    do { if (wgs == 4) NADM_DEC_GO(L, W, 4, H); else NADM_DEC_GO(L, W, 3, H); } while (0)
    // 3 slots always run 3 warpgroups: a fourth has no slot to work in, and a warpgroup's first mbarrier wait is only
    // safe when the slot's previous phase is known to be over, i.e. WGS <= SLOTS (static_assert in the kernel).
    if (KH == 2) {
        if (want_loss) NADM_DEC_GO(true, 3, 3, 2); else NADM_DEC_GO(false, 3, 3, 2);
    } else if (slots == 4) {
        if (want_loss) NADM_DEC_GO_WG(true, 4, 1); else NADM_DEC_GO_WG(false, 4, 1);
    } else {
        if (want_loss) NADM_DEC_GO(true, 3, 3, 1); else NADM_DEC_GO(false, 3, 3, 1);
    }
#undef NADM_DEC_GO_WG
#undef NADM_DEC_GO
    if (rc != NADM_OK || fused) return rc;
    if (defer && ncta <= 152) {
        // the consumer (nadm_mlp_bwd on this dQ) sums the partials of its own rows: no reduction kernel
        DeferredDQ& d = deferred_dq();
        d.dQ = dQ; d.part = dQpart; d.loss_part = want_loss ? loss_part : nullptr; d.loss = loss;
        d.nparts = ncta; d.B = B; d.cols_p = 8 * KH; d.k = k; d.q_ld = q_ld; d.q_off = q_off;
        d.bytes = (size_t)ncta * ((size_t)B * 8 * KH + 1) * sizeof(float);
        return NADM_OK;
    }
    return launch_reduce_parts(dQpart, ncta, B, 8 * KH, k, dQ, q_ld, q_off, 1.0f, want_loss ? loss_part : nullptr, loss, st);
}

}  // namespace nadm
