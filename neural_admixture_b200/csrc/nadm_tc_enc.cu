// Encoder projection and its backward on the 5th-generation tensor cores (tcgen05, kind::i8), sm_100a.
//
//   forward   Z[b, c]  = sum_m x[b, m] V[m, c]              (neural_admixture.py:169-172)
//   backward  dV[m, c] = sum_b x[b, m] dZ[b, c] ; Adam(V)     (autograd of :172, optimizer.step :411)
//
// x = code / 2 with code in {0, 1, 2} (missing 3 -> 0) is an exact small integer, so both contractions are done in
// EXACT integer arithmetic: the fp32 factor (V, resp. dZ) is written as a 32-bit fixed-point number relative to a
// power-of-two scale >= max|.| and split into four signed base-256 digits; the tensor core multiplies the unsigned
// genotype bytes with the int8 digit planes (N = 4 planes x 8 components = 32 columns) into int32 accumulators in
// tensor memory, and the epilogue recombines the planes in int64.  The result is independent of summation order
// (bit-reproducible for any grid / sharding) and carries only the fixed-point quantisation of the fp32 factor
// (<= 2^-31 of its largest entry), i.e. it is more accurate than an fp32 FMA chain.
//
// Data movement: the B gathered rows of the 2-bit packed matrix are streamed once per kernel with 16-byte asynchronous
// copies (cp.async, one commit group per tile, 8 tiles = 64 KB in flight per CTA; 64-byte bulk-engine copies were
// measured 2x slower: the TMA engine is bound by copies/s, not bytes/s) into a shared-memory staging ring, widened to
// one byte per genotype with four shift/mask operations per 16 SNPs and stored straight into the UMMA layout; the same tile is the
// K-major operand X of the forward and the MN-major operand X^T of the backward (nadm_tc.cuh).  Inside a group of 16
// SNPs the byte position p holds SNP sigma(p) = 4 (p % 4) + p / 4; the digit operand / epilogue use the same map.
#include "nadm_common.cuh"
#include "nadm_tc.cuh"

namespace nadm {
using namespace tc;

constexpr int kSub = 256;                  // SNPs per genotype tile: 64 packed bytes of every row
constexpr int kATile = 128 * kSub;         // bytes of one widened tile (128 rows x 256 SNPs)
constexpr int kAStages = 4;                // widened tiles: producer group g fills tiles i = g (mod 4) into stage g
constexpr int kDigTile = kSub * 32;        // digit bytes per 256 K positions (4 planes x 8 components each)
constexpr int kProdWarps = 16;             // 4 producer groups of 4 warps; a group widens every 4th tile
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kStTile = 128 * 64;          // packed bytes of one tile in the staging ring
constexpr int kStDepth = 2;                // tiles in flight from HBM per producer group (cp.async commit groups)
#ifndef NADM_ENC_TS_DEPTH
#define NADM_ENC_TS_DEPTH 2
#endif
constexpr int kStDepthTS = NADM_ENC_TS_DEPTH;   // the same for the tensor-memory operand variant (no widened tiles in
                                                // shared memory: room for a deeper ring)
constexpr int kMaxBlkTc = 16;              // 16 blocks of 128 rows per launch (16 x 32 tensor-memory columns)
constexpr uint32_t kIdescFwd = instr_desc(kAccS32, kFmtU8, kFmtS8, /*A MN*/ false, /*B MN*/ true, 128, 32);
constexpr uint32_t kIdescBwd = instr_desc(kAccS32, kFmtU8, kFmtS8, /*A MN*/ true, /*B MN*/ true, 128, 32);

__device__ __forceinline__ int sigma16(int p) { return 4 * (p & 3) + (p >> 2); }

__device__ __forceinline__ uint4 ldg_nc(const uint8_t* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// 16 packed SNPs -> 16 bytes, byte position p <-> SNP sigma16(p).  RAW = false: codes 0,1,2 with missing (3) cleared
// (the training path: x = code / 2, missing -> 0).  RAW = true: the reader's uint8 VALUES 0,1,2 and `missing_value` for
// code 3 (what the reference's randomized-SVD products multiply by, rsvd.pyx:16-50); mvx = 3 ^ missing_value.
template <bool RAW>
__device__ __forceinline__ void widen_store(uint8_t* dst, uint32_t w, uint32_t mvx) {
    uint4 o;
#ifdef NADM_KO_WIDEN   // knock-out measurement build: no widening arithmetic (results are garbage by design)
    *reinterpret_cast<uint4*>(dst) = make_uint4(w, w, w, w);
    return;
#endif
    if (!RAW) {
        const uint32_t c = clear_missing(w);
        o.x = c & 0x03030303u;
        o.y = (c >> 2) & 0x03030303u;
        o.z = (c >> 4) & 0x03030303u;
        o.w = (c >> 6) & 0x03030303u;
    } else {
        uint32_t b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t v = (w >> (2 * k)) & 0x03030303u;
            const uint32_t m3 = v & (v >> 1) & 0x01010101u;          // bytes equal to 3
            b[k] = v ^ (m3 * mvx);
        }
        o = make_uint4(b[0], b[1], b[2], b[3]);
    }
    *reinterpret_cast<uint4*>(dst) = o;
}

// the same widening into registers (tensor-memory operand path): o[0..3] = byte positions 0-3, 4-7, 8-11, 12-15
template <bool RAW>
__device__ __forceinline__ void widen_regs(uint32_t w, uint32_t mvx, uint32_t* o) {
#ifdef NADM_KO_WIDEN   // knock-out measurement build: no widening arithmetic (results are garbage by design)
    o[0] = w; o[1] = w; o[2] = w; o[3] = w;
    return;
#endif
    if (!RAW) {
#ifdef NADM_ENC_TS_SCALED
        // fields left in place: byte positions 4j..4j+3 hold 4^j x code (<= 128, still a u8), one LOP3 per output word
        // and no shifts (integer/logic instructions issue at one per 2 clocks); the digit operand of those K positions
        // is built from V / 4^j (store_digits_scaled), i.e. they carry 2j fewer fixed-point bits (>= 2^-25 of max|V|)
        const uint32_t m = w & (w >> 1) & 0x55555555u;
        const uint32_t keep = ~(m * 3u);
        o[0] = w & 0x03030303u & keep;
        o[1] = w & 0x0C0C0C0Cu & keep;
        o[2] = w & 0x30303030u & keep;
        o[3] = w & 0xC0C0C0C0u & keep;
#else
        const uint32_t c = clear_missing(w);
        o[0] = c & 0x03030303u;
        o[1] = (c >> 2) & 0x03030303u;
        o[2] = (c >> 4) & 0x03030303u;
        o[3] = (c >> 6) & 0x03030303u;
#endif
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t v = (w >> (2 * k)) & 0x03030303u;
            const uint32_t m3 = v & (v >> 1) & 0x01010101u;          // bytes equal to 3
            o[k] = v ^ (m3 * mvx);
        }
    }
}

// |max| over floats as the maximum of their sign-cleared BIT PATTERNS: same order as the values for finite numbers, and
// Inf / NaN (which fmaxf would drop) come out on top, so that fix_scale sees them
__device__ __forceinline__ uint32_t absbits(float x) { return __float_as_uint(x) & 0x7FFFFFFFu; }
__device__ __forceinline__ uint32_t absbits_max4(uint32_t m, const float4& x) {
    return max(max(m, max(absbits(x.x), absbits(x.y))), max(absbits(x.z), absbits(x.w)));
}
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t m) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    return m;
}
// four signed base-256 digits of q, most significant first: q = ((d0*256 + d1)*256 + d2)*256 + d3
__device__ __forceinline__ void digits4(int q, int (&d)[4]) {
#pragma unroll
    for (int i = 3; i > 0; --i) {
        const int low = (q << 24) >> 24;
        d[i] = low;
        q = (q - low) >> 8;
    }
    d[0] = q;
}
__device__ __forceinline__ long long combine4(const uint32_t* v, int c) {
    long long z = (int)v[c];
    z = z * 256 + (int)v[8 + c];
    z = z * 256 + (int)v[16 + c];
    z = z * 256 + (int)v[24 + c];
    return z;
}
// write the digits of 8 components (one K position) into an MN-major digit tile: 2 chunks of 16 bytes, 128 B apart
// The signed base-256 digits of q = sum_i d_i 256^i, d_i in [-128, 127], are the bytes of (q + 0x80808080) with their
// top bits flipped back: adding 128 to every digit makes all of them non-negative, so the carries of the ordinary
// binary addition do the borrowing.  A 4 x 4 byte transpose (8 PRMT) then turns four components' words into the four
// digit planes: 24 instructions per 4 components instead of ~65 with shifts.
__device__ __forceinline__ void store_digits(uint8_t* tile_pos, const float (&v)[8], float inv) {
    uint32_t w[8];  // w[2*p + h]: plane p (0 = most significant digit), components 4h..4h+3
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t t[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) t[c] = ((uint32_t)__float2int_rn(v[4 * h + c] * inv) + 0x80808080u) ^ 0x80808080u;
        const uint32_t lo01 = __byte_perm(t[0], t[1], 0x5140), hi01 = __byte_perm(t[0], t[1], 0x7362);
        const uint32_t lo23 = __byte_perm(t[2], t[3], 0x5140), hi23 = __byte_perm(t[2], t[3], 0x7362);
        w[6 + h] = __byte_perm(lo01, lo23, 0x5410);     // least significant digits of the four components
        w[4 + h] = __byte_perm(lo01, lo23, 0x7632);
        w[2 + h] = __byte_perm(hi01, hi23, 0x5410);
        w[0 + h] = __byte_perm(hi01, hi23, 0x7632);     // most significant
    }
    *reinterpret_cast<uint4*>(tile_pos) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(tile_pos + 128) = make_uint4(w[4], w[5], w[6], w[7]);
}

// ---- genotype feed: HBM -> staging ring (bulk async copies) -> widened UMMA tiles ----------------------------------
// Tile i of a CTA = rows [blk*128, blk*128+128) x SNPs [(t0+tt)*256, +256), order (tt, blk).  Producer thread tid < 128
// issues the copy of row blk*128 + tid; all 256 producer threads widen.
struct Feed {
    const uint8_t* packed;
    int64_t pitch;
    const uint32_t* rowoff;  // row NUMBER of every batch row (< 2^32; the byte offset row x pitch is formed in 64 bits)
    int B, nblk, t0, ntile;
    uint8_t* stage;          // this group's ring: kStDepth x kStTile; thread (wl, lane) owns 4 pieces of 16 bytes per slot
    int c_blk, c_tt, c_i;    // next tile of this group to copy (tile index c_i = g + 4 n)
    int c_slot;
    int depth;               // slots of the ring
};
// every thread of the group: start the asynchronous copy of its four 16-byte pieces of the group's next tile.
// Pieces of rows past the batch (or past the row pitch) are zeroed with a plain shared-memory store instead of a
// zero-fill copy: the variable source size of cp.async costs ~8 address-fixup instructions per copy.
__device__ __forceinline__ void cp_async16_full(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void feed_issue(Feed& f, int wl, int lane) {
    if (f.c_i < f.ntile) {
        const int r = lane & 7, q = lane >> 3;
        const int64_t off = (int64_t)(f.t0 + f.c_tt) * (kSub / 4) + q * 16;
        uint8_t* dst = f.stage + f.c_slot * kStTile + (wl * 32 + lane) * 16;   // piece `it` at + it * 2048 (conflict-free)
        const uint8_t* base = f.packed + off;
        const int b0 = f.c_blk * 128 + wl * 32 + r;                            // piece `it` holds row b0 + 8 it
        const int nrows = (off + 16 <= f.pitch) ? f.B : 0;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int b = b0 + it * 8;
#ifdef NADM_KO_LOAD     // knock-out measurement build: no global loads
            if (b >= nrows) *reinterpret_cast<uint4*>(dst + it * 2048) = make_uint4(0u, 0u, 0u, 0u);
#else
            if (b < nrows) cp_async16_full(dst + it * 2048, base + (uint64_t)f.rowoff[b] * (uint64_t)f.pitch);
            else *reinterpret_cast<uint4*>(dst + it * 2048) = make_uint4(0u, 0u, 0u, 0u);
#endif
        }
        f.c_i += 4;
        f.c_blk += 4;
        while (f.c_blk >= f.nblk) { f.c_blk -= f.nblk; ++f.c_tt; }
    }
    f.c_slot = (f.c_slot + 1 == f.depth) ? 0 : f.c_slot + 1;
    cp_async_commit();       // always one group per call: keeps wait_group counting aligned at the tail
}
// load this thread's four staged pieces of the tile in staging slot `slot` into registers (they have landed once at most
// kStDepth-1 newer copy groups are pending).  After this the slot may be refilled (feed_issue) while the pieces are
// widened: the refill overlaps the wait for the MMAs to release the widened-tile stage.
__device__ __forceinline__ void feed_load(const Feed& f, int slot, int wl, int lane, uint4 (&w)[4]) {
    cp_async_wait<kStDepth - 1>();
    const uint8_t* st = f.stage + slot * kStTile + (wl * 32 + lane) * 16;
#pragma unroll
    for (int it = 0; it < 4; ++it) w[it] = *reinterpret_cast<const uint4*>(st + it * 2048);
}
// widen the four pieces into `tile`; 8-row groups past the batch were zero-filled by the copy and widen to zeros that
// no MMA K step reads
template <bool RAW>
__device__ __forceinline__ void feed_store(uint8_t* tile, int wl, int lane, const uint4 (&w)[4], uint32_t mvx) {
    const int r = lane & 7, q = lane >> 3;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int g8 = wl * 4 + it;                              // 8-row group inside the block
        uint8_t* dst = tile + r * 16 + g8 * 2048 + (q * 4) * 128;
        widen_store<RAW>(dst, w[it].x, mvx);
        widen_store<RAW>(dst + 128, w[it].y, mvx);
        widen_store<RAW>(dst + 256, w[it].z, mvx);
        widen_store<RAW>(dst + 384, w[it].w, mvx);
    }
}

// ---- tensor-memory operand path of the forward kernel (TSA): thread = tensor-memory lane = batch row ----------------
// Thread `lane` of producer warp wl takes the four 16-byte pieces of row wl*32 + lane of the block (they were copied by
// four different lanes of this warp: piece q of row r + 8 it by lane 8 q + r into plane `it`), widens them in registers
// and stores them with tcgen05.st: 4 K positions per 32-bit column, 64 columns per 128 x 256 tile (K order = the byte
// order of the shared-memory tile, so the digit operand is unchanged).
template <int DEPTH>
__device__ __forceinline__ void feed_load_rows(const Feed& f, int slot, int wl, int lane, uint4 (&w)[4]) {
    cp_async_wait<DEPTH - 1>();
    __syncwarp();                                                // the other lanes' copies of my row have landed
    const uint8_t* st = f.stage + slot * kStTile + (wl * 32 + (lane & 7)) * 16 + (lane >> 3) * 2048;
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = *reinterpret_cast<const uint4*>(st + q * 128);
    __syncwarp();                                                // every lane has read before any lane refills the slot
}
template <bool RAW, int Q0 = 0, int Q1 = 4>   // pieces [Q0, Q1) of the row: 64 SNPs = 16 columns each
__device__ __forceinline__ void feed_store_tmem(uint32_t taddr, const uint4 (&w)[4], uint32_t mvx) {
#pragma unroll
    for (int q = Q0; q < Q1; ++q) {
        uint32_t v[16];
        widen_regs<RAW>(w[q].x, mvx, v);
        widen_regs<RAW>(w[q].y, mvx, v + 4);
        widen_regs<RAW>(w[q].z, mvx, v + 8);
        widen_regs<RAW>(w[q].w, mvx, v + 12);
        tmem_st16(taddr + q * 16, v);
    }
}
constexpr int kColA = 256;   // TSA: tensor-memory columns [256, 512) = 4 genotype tiles of 64 columns; [0, 256) accumulators

// =================================================================================================================
// forward
// =================================================================================================================
#ifdef NADM_TIMELINE
__device__ long long g_enc_timeline[8][512];
#define TLE(row, idx) do { if (blockIdx.x == 0 && (idx) < 512) g_enc_timeline[row][idx] = clock64(); } while (0)
extern "C" int nadm_debug_enc_timeline(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_enc_timeline, sizeof(g_enc_timeline));
}
#else
#define TLE(row, idx) do { } while (0)
#endif

constexpr int kZredOut = 64, kZredSeg = 16;   // outputs per chunk x part segments per output (1024 work items per chunk)
__device__ GridBar g_bar_enc_fwd;

// After the grid barrier: Z[b, c] = out_scale * sum over CTAs p of part_p[b, c] * 2^(e_p - 30).  Every CTA's partial is an
// exact integer at the CTA's own power-of-two scale (exactly representable in double, as is the product); the CTAs are
// summed in double in a fixed order (deterministic).  This CTA takes outputs [n cta / nctas, n (cta + 1) / nctas).
// Latency-bound: a work item = (output, segment of the parts) holds all its loads in flight at once (<= 10 for 148
// CTAs); with 4 segments of 37 dependent-looking loads the fused reduction was slower than the kernel it replaced.
__device__ __forceinline__ void enc_fwd_reduce_share(const long long* __restrict__ part, const float* __restrict__ cta_vmax,
                                                     int nparts, int B, int C, float* __restrict__ Z, double out_scale,
                                                     double* back /* shared: nparts doubles */, double* zred /* shared */) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    (void)cta_vmax;                                                     // (the partials are pre-scaled doubles)
    for (int p = tid; p < nparts; p += nthr) back[p] = 1.0;
    __syncthreads();
    const int64_t n = (int64_t)B * 8;
    const int o0 = (int)((n * blockIdx.x) / gridDim.x), o1 = (int)((n * (blockIdx.x + 1)) / gridDim.x);
    for (int c0 = o0; c0 < o1; c0 += kZredOut) {
        const int cnt = min(kZredOut, o1 - c0);
        for (int w = tid; w < cnt * kZredSeg; w += nthr) {
            const int o = c0 + w / kZredSeg, seg = w % kZredSeg;
            const long long* src = part + o;
            if (nparts <= 10 * kZredSeg) {                              // (<= 160 CTAs: every load of the item in flight)
                long long v[10];
#pragma unroll
                for (int i = 0; i < 10; ++i) {
                    const int p = seg + i * kZredSeg;
                    v[i] = (p < nparts) ? __ldcg(src + (int64_t)p * n) : 0ll;
                }
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < 10; ++i) {
                    const int p = seg + i * kZredSeg;
                    if (p < nparts) acc += __longlong_as_double(v[i]) * back[p];
                }
                zred[w] = acc;
            } else {
                double acc = 0.0;
                for (int p = seg; p < nparts; p += kZredSeg) acc += __longlong_as_double(__ldcg(src + (int64_t)p * n)) * back[p];
                zred[w] = acc;
            }
        }
        __syncthreads();
        for (int w = tid; w < cnt; w += nthr) {
            const int o = c0 + w, b = o >> 3, c = o & 7;
            double t = 0.0;
#pragma unroll
            for (int sg = 0; sg < kZredSeg; ++sg) t += zred[w * kZredSeg + sg];
            if (c < C) Z[(int64_t)b * C + c] = (float)(t * out_scale);
        }
        __syncthreads();
    }
}

struct EncSmem {
    uint64_t fullA[kAStages], emptyA[kAStages], fullV[2], emptyV[2], done, fdone;
    uint32_t tmem_base;
    uint32_t finit[2];   // per issuer: bit blk set = its accumulator of row block blk has been written
    float red[32];       // per-warp |max| of this CTA's slice of V
};

// Two MMA-issuer warps: a single issuing thread (waits + descriptor arithmetic + 8 MMAs + commits per tile, ~770 cycles
// measured against 320 cycles of tensor-pipe time) was the kernel's critical path.  Issuer p owns the widened-tile
// stages 2p and 2p+1 (producer groups 2p, 2p+1), i.e. the tiles i with (i >> 1) & 1 == p, and issues all 8 K steps of
// its tiles into its OWN accumulator set (tensor-memory columns (p nblk + blk) 32); the epilogue adds the two int32
// sets.  Each issuer sees every phase of the mbarriers of its own stages.  (Splitting the tiles by row-block parity is
// unsafe — an issuer that skips phases of a stage's mbarrier cannot tell them apart by parity: a timing-dependent hang;
// splitting the K steps of every tile keeps the per-tile overhead on both issuers and gains nothing.)
// With 1 issuer, fewer than 3 or more than 8 row blocks: one accumulator set.
constexpr int kFwdDigWarps = 2, kFwdIssueWarp = kProdWarps + kFwdDigWarps, kFwdThreads = (kFwdIssueWarp + 2) * 32;

// TSA: the genotype operand lives in TENSOR memory (tcgen05.mma TS form) instead of shared memory: the 32 KB widened
// tile is then neither written to nor read from shared memory (the kernel's measured bound).  Both issuers accumulate
// into ONE accumulator set (zero-initialised, every MMA accumulates; integer adds commute; tests/cuda/umma_probe.cu
// case 7): 8 x 32 accumulator columns + 4 x 64 tile columns = 512.
template <int NISS, bool RAW, bool TSA>   // NISS: MMA issuer warps in use (2, or 1: the second then idles); RAW: see widen_store
__global__ void __launch_bounds__(kFwdThreads, 1)
enc_fwd_tc_kernel(const uint8_t* __restrict__ packed, int64_t pitch, const int64_t* __restrict__ row_idx, int64_t row0,
                  int B, int64_t M, const float* __restrict__ V, int C, float* __restrict__ cta_vmax,
                  long long* __restrict__ part, int T, uint32_t mvx, float* __restrict__ Z, double out_scale) {
    // Z != NULL: the CTAs meet at a grid barrier after writing their partials and reduce them themselves (cooperative
    // launch); Z == NULL: a separate enc_fwd_reduce_kernel follows.
    pdl_prologue();
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int kDepth = TSA ? kStDepthTS : kStDepth;
    uint8_t* tilesA = smem;                                         // kAStages x 32 KB (not allocated with TSA)
    uint8_t* tilesV = TSA ? smem : tilesA + kAStages * kATile;      // 2 x 8 KB
    uint8_t* stage = tilesV + 2 * kDigTile;                         // 4 groups x kDepth x 8 KB packed rows
    uint32_t* rowoff = reinterpret_cast<uint32_t*>(stage + 4 * kDepth * kStTile);
    EncSmem* S = reinterpret_cast<EncSmem*>(rowoff + ((B + 3) & ~3));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nblk = (B + 127) / 128;
    const int t0 = (int)(((int64_t)T * blockIdx.x) / gridDim.x), t1 = (int)(((int64_t)T * (blockIdx.x + 1)) / gridDim.x);
    const int ntile = (t1 - t0) * nblk;
    if (tid == 0) TLE(2, 0);                                            // kernel entry

    for (int b = tid; b < B; b += blockDim.x)
        rowoff[b] = (uint32_t)((row_idx != nullptr) ? row_idx[b] : (row0 + b));
    if (tid == 0) {
        for (int s = 0; s < kAStages; ++s) { mbar_init(&S->fullA[s], 4); mbar_init(&S->emptyA[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&S->fullV[s], kFwdDigWarps); mbar_init(&S->emptyV[s], NISS); }
        mbar_init(&S->done, NISS);
        mbar_init(&S->fdone, NISS);
        S->finit[0] = S->finit[1] = 0u;
        mbar_init_fence();
    }
    if (warp == kFwdIssueWarp) tmem_alloc<512>(&S->tmem_base);
    __syncthreads();                                                // row numbers are in shared memory
    // the producers' first copies are in flight while the scale of V is computed (they need the row numbers only)
    Feed f{packed, pitch, rowoff, B, nblk, t0, ntile, stage + (warp >> 2 & 3) * (kDepth * kStTile), (warp >> 2) % nblk,
           (warp >> 2) / nblk, warp >> 2, 0, kDepth};
    if (warp < kProdWarps)
        for (int p = 0; p < kDepth; ++p) feed_issue(f, warp & 3, lane);
    // |max| of THIS CTA's rows of V: the fixed-point scale is per CTA (its partial sums are exact integers at that scale;
    // the reduction kernel converts each CTA's partial with the CTA's own power-of-two scale).  No grid-wide pass over V.
    {
        const int64_t m0 = (int64_t)t0 * kSub, m1 = min((int64_t)t1 * kSub, M);
        const int64_t n = (m1 > m0) ? (m1 - m0) * C : 0;
        const float* v0 = V + m0 * C;
        uint32_t mx = 0u;
        if ((reinterpret_cast<uintptr_t>(v0) & 15) == 0) {
            const int64_t n4 = n / 4;
#pragma unroll 8
            for (int64_t i = tid; i < n4; i += blockDim.x) mx = absbits_max4(mx, reinterpret_cast<const float4*>(v0)[i]);
            for (int64_t i = n4 * 4 + tid; i < n; i += blockDim.x) mx = max(mx, absbits(v0[i]));
        } else {
            for (int64_t i = tid; i < n; i += blockDim.x) mx = max(mx, absbits(v0[i]));
        }
        mx = warp_max_u32(mx);
        if (lane == 0) S->red[warp] = __uint_as_float(mx);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = S->tmem_base;
    float vmax;
    {
        uint32_t m = 0u;
        for (int w = 0; w < kFwdThreads / 32; ++w) m = max(m, __float_as_uint(S->red[w]));
        vmax = __uint_as_float(m);
    }
    if (tid == 0) cta_vmax[blockIdx.x] = vmax;
    if (TSA) {
        if (warp < 4) {                                                 // warps 0..3 = lane quadrants 0..3
            uint32_t z[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] = 0u;
            for (int c0 = 0; c0 < nblk * 32; c0 += 16) tmem_st16(tbase + ((uint32_t)(warp * 32) << 16) + c0, z);
            tmem_wait_st();
        }
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
    }
    if (tid == 0) TLE(2, 1);                                            // setup done

    if (warp < kProdWarps) {
        // ---------------- producers: widened genotype tiles (group g = warp / 4 handles tiles g, g + 4, ...) ----------------
        const int g = warp >> 2, wl = warp & 3;
        int blk = g % nblk, slot = 0, phase = 1;
        for (int i = g; i < ntile; i += 4) {
            uint4 w[4];
            if (tid == 0) TLE(0, i);
            if (TSA) feed_load_rows<kDepth>(f, slot, wl, lane, w);
            else feed_load(f, slot, wl, lane, w);
            feed_issue(f, wl, lane);                                 // refill the staging slot just read
            if (tid == 0) TLE(5, i);
            mbar_wait(&S->emptyA[g], phase);
            if (tid == 0) TLE(1, i);
            if (TSA) {
                const uint32_t ta = tbase + ((uint32_t)(wl * 32) << 16) + kColA + g * 64;
                tc_fence_after_sync();                               // the MMAs that read this stage have completed
                feed_store_tmem<RAW>(ta, w, mvx);
                tmem_wait_st();
                tc_fence_before_sync();
            } else {
                feed_store<RAW>(tilesA + g * kATile, wl, lane, w, mvx);
                fence_async_smem();
            }
            if (tid == 0) TLE(3, i);
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->fullA[g]);                // one arrival per warp of the group
            if (tid == 0) TLE(4, i);
            phase ^= 1;
            slot = (slot + 1 == kDepth) ? 0 : slot + 1;
            blk += 4;
            while (blk >= nblk) blk -= nblk;
        }
        // ---------------- epilogue: recombine the digit planes, write this CTA's exact partial sums ----------------
        if (tid == 0) TLE(2, 2);                                        // producers done
        mbar_wait(&S->done, 0);
        mbar_wait(&S->fdone, 0);
        if (tid == 0) TLE(2, 3);                                        // MMAs done
        tc_fence_after_sync();
        const uint32_t init0 = S->finit[0], init1 = S->finit[1];
        const int q = warp & 3;
        for (int blk = warp >> 2; blk < nblk; blk += kProdWarps / 4) {
            uint32_t v[32];
            tmem_ld32(tbase + ((uint32_t)(q * 32) << 16) + blk * 32, v);
            tmem_wait_ld();
            if (!TSA && !((init0 >> blk) & 1u)) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if (!TSA && NISS == 2 && ((init1 >> blk) & 1u)) {
                uint32_t v1[32];
                tmem_ld32(tbase + ((uint32_t)(q * 32) << 16) + (nblk + blk) * 32, v1);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (uint32_t)((int)v[j] + (int)v1[j]);
            }
            const int b = blk * 128 + q * 32 + lane;
            if (b < B) {
                // the CTA's exact integer sums, times its power-of-two scale: exact in double (|sum| < 2^45), so the
                // consumers just add doubles
                long long* out = part + ((int64_t)blockIdx.x * B + b) * 8;
                const double back = fix_scale(vmax).back;
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    longlong2 z;
                    z.x = __double_as_longlong((double)combine4(v, c) * back);
                    z.y = __double_as_longlong((double)combine4(v, c + 1) * back);
                    *reinterpret_cast<longlong2*>(out + c) = z;
                }
            }
        }
    } else if (warp < kFwdIssueWarp) {
        // ---------------- digit warps: int8 digit planes of V, one 256-SNP sub-tile ahead of the MMAs ----------------
        const FixScale fs = fix_scale(vmax);
        const int dt = tid - kProdThreads;                              // 0..63: K positions dt, dt+64, dt+128, dt+192
        auto load_v = [&](int tt, float (&v)[4][8]) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int pos = dt + 64 * e;
                const int64_t m = (int64_t)(t0 + tt) * kSub + (pos & ~15) + sigma16(pos & 15);
                if (m < M && C == 8) {
                    const float4 a = reinterpret_cast<const float4*>(V + m * 8)[0], b4 = reinterpret_cast<const float4*>(V + m * 8)[1];
                    v[e][0] = a.x; v[e][1] = a.y; v[e][2] = a.z; v[e][3] = a.w;
                    v[e][4] = b4.x; v[e][5] = b4.y; v[e][6] = b4.z; v[e][7] = b4.w;
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[e][c] = (m < M && c < C) ? V[m * C + c] : 0.f;
                }
            }
        };
        float vnext[4][8];
        if (t1 > t0) load_v(0, vnext);
        for (int tt = 0; tt < t1 - t0; ++tt) {
            const int vs = tt & 1;
            float v[4][8];
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int c = 0; c < 8; ++c) v[e][c] = vnext[e][c];
            if (t0 + tt + 1 < t1) load_v(tt + 1, vnext);
            mbar_wait_relaxed(&S->emptyV[vs], ((tt >> 1) & 1) ^ 1, 128);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int pos = dt + 64 * e;
#ifdef NADM_ENC_TS_SCALED
                const float inv_pos = (TSA && !RAW) ? fs.inv * __uint_as_float((uint32_t)(127 - 2 * ((pos & 15) >> 2)) << 23) : fs.inv;
#else
                const float inv_pos = fs.inv;
#endif
                store_digits(tilesV + vs * kDigTile + (pos & 7) * 16 + (pos >> 3) * 256, v[e], inv_pos);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->fullV[vs]);
        }
    } else {
        // ---------------- MMA issuers: the whole warp runs the convergent loop, only the elected lane's MMAs / commits
        // execute; descriptors built once.  Issuer `par` takes the tiles in its two stages ----------------
        const int par = warp - kFwdIssueWarp;
        if (par < NISS) {
            const uint64_t A0 = smem_desc(smem_u32(tilesA), 128, 2048), B0 = smem_desc(smem_u32(tilesV), 256, 128);
            const uint32_t leader = elect_one() ? 1u : 0u;
            uint32_t inited = 0u;
            int i = 0;                                                   // tile index, order (sub-tile, row block)
            for (int tt = 0; tt < t1 - t0; ++tt) {
                const int vs = tt & 1;
                mbar_wait(&S->fullV[vs], (tt >> 1) & 1);                 // every issuer waits for every sub-tile's digits
                for (int blk = 0; blk < nblk; ++blk, ++i) {
                    if (NISS == 2 && ((i >> 1) & 1) != par) continue;
                    const int s = i & (kAStages - 1);
                    mbar_wait(&S->fullA[s], (i >> 2) & 1);
                    tc_fence_after_sync();
                    if (lane == 0) TLE(6, i);
                    const uint64_t a = A0 + (uint64_t)(s * (kATile >> 4)), b = B0 + (uint64_t)(vs * (kDigTile >> 4));
                    const uint32_t d = tbase + ((TSA ? 0 : par * nblk) + blk) * 32, acc0 = (inited >> blk) & 1u;
                    if (TSA) {
                        const uint32_t at = tbase + kColA + s * 64;
#pragma unroll
                        for (int ks = 0; ks < kSub / 32; ++ks)
                            mma_i8_ts_p(d, at + ks * 8, b + (uint64_t)(ks * 64), kIdescFwd, 1u, leader);
                    } else {
#pragma unroll
                        for (int ks = 0; ks < kSub / 32; ++ks)
                            mma_i8_ss_p(d, a + (uint64_t)(ks * 16), b + (uint64_t)(ks * 64), kIdescFwd, ks ? 1u : acc0, leader);
                    }
                    inited |= 1u << blk;
                    mma_commit_p(&S->emptyA[s], leader);
                    if (lane == 0) TLE(7, i);
                }
                mma_commit_p(&S->emptyV[vs], leader);
            }
            if (lane == 0) {
                S->finit[par] = inited;
                __threadfence_block();
                mbar_arrive(&S->fdone);
            }
            mma_commit_p(&S->done, leader);
        }
        __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) TLE(2, 4);                                            // epilogue done
    if (warp == kFwdIssueWarp) tmem_dealloc<512>(tbase);
    if (Z != nullptr) {
        grid_barrier(&g_bar_enc_fwd, gridDim.x);
        if (tid == 0) TLE(2, 5);                                        // every CTA's partial is visible
        // (the staging ring is idle: all copies were consumed by the producers' loop)
        enc_fwd_reduce_share(part, cta_vmax, (int)gridDim.x, B, C, Z, out_scale, reinterpret_cast<double*>(stage),
                             reinterpret_cast<double*>(stage) + kMaxParts);
        if (tid == 0) TLE(2, 6);
    }
}

// =================================================================================================================
// forward, slab feed (OPT-IN: NADM_ENC_FWD_SLAB=1; a measurement switch, not the default — see the end of this comment):
// the same contraction as above, but the rows are streamed in 128- or 256-BYTE runs.
//
// Why it was built: profiles/r2_gather_probe.txt.  Gathering B random sample rows out of HBM at 64 contiguous bytes per
// row and request (one 256-SNP tile) reaches 2.1 TB/s whatever the depth, the tile order, the engine (cp.async, 1-D bulk
// copies, TMA gather4) or an L2 prefetch hint; 128-byte runs reach 3.8 TB/s and 256-byte runs 4.5 TB/s (the same rows out
// of L2: 5.2 TB/s).  So a producer group copies a SLAB = one row block (128 rows) x kSlabSub consecutive sub-tiles with
// one burst of cp.async instructions whose lanes cover contiguous runs of 64 kSlabSub bytes per row, and then widens its
// 256-SNP tiles out of it.  Tile order: slab column outer, row block, sub-tile inner; the accumulators of all row blocks
// stay in tensor memory as before, kSlabSub digit tiles (one slab column) are live and the next ones are produced ahead.
// The genotype operand lives in tensor memory (TS form); widening leaves the 2-bit fields in place (4^j x code, j =
// field index) and the digit operand of those K positions is built from V / 4^j, which removes the shifts.
// What it measured (profiles/r2_encoder_slab_variants.txt): 59 vs 65 us alone, 0.4428 vs 0.4425 ms per step on the same
// box, and -11 % on the forward-only Q pass (1024 instead of 2048 rows per launch) -> the round-1 kernel stays the default.
// =================================================================================================================
#ifndef NADM_SLAB_SUB
#define NADM_SLAB_SUB 2
#endif
#ifndef NADM_SLAB_DEPTH
#define NADM_SLAB_DEPTH 2
#endif
// A slab of kSlabSub sub-tiles = 64 kSlabSub contiguous bytes of every row; kSlabDepth slabs per producer group are in
// flight / being widened.  Measured on one box (profiles/r2_*): 4 sub-tiles x depth 1 (256-byte runs, but nothing in
// flight while a group widens its four tiles) was SLOWER than the round-1 kernel; 2 x 2 (128-byte runs, the next slab
// lands while this one is widened) is the default.  Both use 32 KB of staging per group.
constexpr int kSlabSub = NADM_SLAB_SUB;            // sub-tiles per slab (2 or 4)
constexpr int kSlabDepth = NADM_SLAB_DEPTH;        // slabs per producer group
constexpr int kSlabBytes = 128 * 64 * kSlabSub;    // 16 KB (32 KB)
constexpr int kSlabPieces = 4 * kSlabSub;          // 16-byte pieces per row
constexpr int kSlabRowsPerInstr = 32 / kSlabPieces, kSlabIters = 32 / kSlabRowsPerInstr;
enum : int { kModeTrain = 0, kModeRaw3 = 1, kModeRawAny = 2 };   // what a 2-bit code means (see widen_fields)

// one packed word (16 SNPs) -> 4 words of 4 bytes; byte b of o[k] belongs to SNP 4 b + k of the word (= sigma16(4 k + b)).
//   kModeTrain : 4^k x code, code 3 (missing) -> 0          (x = code / 2, missing trained as 0)
//   kModeRaw3  : 4^k x code, code 3 stays 3                  (randomized-SVD products when the reader's missing value is 3)
//   kModeRawAny: the byte VALUE, code 3 -> any uint8 value   (unscaled: 255 x 64 would not fit a byte)
template <int MODE>
__device__ __forceinline__ void widen_fields(uint32_t w, uint32_t mvx, uint32_t* o) {
#ifdef NADM_KO_WIDEN
    o[0] = w; o[1] = w; o[2] = w; o[3] = w;
    return;
#endif
    if (MODE == kModeRawAny) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t v = (w >> (2 * k)) & 0x03030303u;
            const uint32_t m3 = v & (v >> 1) & 0x01010101u;
            o[k] = v ^ (m3 * mvx);
        }
    } else {
        uint32_t keep = 0xFFFFFFFFu;
        if (MODE == kModeTrain) keep = ~((w & (w >> 1) & 0x55555555u) * 3u);
        o[0] = w & 0x03030303u & keep;
        o[1] = w & 0x0C0C0C0Cu & keep;
        o[2] = w & 0x30303030u & keep;
        o[3] = w & 0xC0C0C0C0u & keep;
    }
}

struct EncSlabSmem {
    uint64_t fullA[4], emptyA[4], fullV[2][kSlabSub], emptyV[2], done;
    uint32_t tmem_base;
    float red[32];
};
constexpr int kSlabThreads = (kProdWarps + 2 + 2) * 32;    // 16 producer warps, 2 digit warps, 2 MMA issuers

// staging address of piece p (16 bytes) of row r inside a slab: piece-major planes, rows XOR-ed with the piece number so
// that both the copy (a quarter warp writes 8 consecutive pieces of one row) and the read-back (a quarter warp reads one
// piece of 8 consecutive rows) touch 8 different 16-byte bank groups
__device__ __forceinline__ uint32_t slab_unit(int r, int p) { return (uint32_t)(p * 128 + (r ^ (p & 7))) << 4; }

template <int NISS, int MODE>
__global__ void __launch_bounds__(kSlabThreads, 1)
enc_fwd_slab_kernel(const uint8_t* __restrict__ packed, int64_t pitch, const int64_t* __restrict__ row_idx, int64_t row0,
                    int B, int64_t M, const float* __restrict__ V, int C, float* __restrict__ cta_vmax,
                    long long* __restrict__ part, int T, uint32_t mvx) {
    pdl_prologue();
    constexpr bool kScaled = MODE != kModeRawAny;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* tilesV = smem;                                          // 2 sets x kSlabSub digit tiles x 8 KB
    uint8_t* slabs = tilesV + 2 * kSlabSub * kDigTile;               // 4 groups x kSlabDepth slabs
    uint32_t* rowoff = reinterpret_cast<uint32_t*>(slabs + 4 * kSlabDepth * kSlabBytes);
    EncSlabSmem* S = reinterpret_cast<EncSlabSmem*>(rowoff + ((B + 3) & ~3));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nblk = (B + 127) / 128;
    const int t0 = (int)(((int64_t)T * blockIdx.x) / gridDim.x), t1 = (int)(((int64_t)T * (blockIdx.x + 1)) / gridDim.x);
    const int ntt = t1 - t0, ncol = (ntt + kSlabSub - 1) / kSlabSub, nslab = ncol * nblk;

    for (int b = tid; b < B; b += blockDim.x)
        rowoff[b] = (uint32_t)((row_idx != nullptr) ? row_idx[b] : (row0 + b));
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) { mbar_init(&S->fullA[s], 4); mbar_init(&S->emptyA[s], 1); }
        for (int s = 0; s < 2; ++s) {
            for (int j = 0; j < kSlabSub; ++j) mbar_init(&S->fullV[s][j], 2);
            mbar_init(&S->emptyV[s], NISS);
        }
        mbar_init(&S->done, NISS);
        mbar_init_fence();
    }
    if (warp == kProdWarps + 2) tmem_alloc<512>(&S->tmem_base);
    __syncthreads();                                                 // row numbers are in shared memory

    // ---- producer side of the feed ----
    const int g = warp >> 2, wl = warp & 3;                          // (producer warps only)
    uint8_t* ring = slabs + (g & 3) * (kSlabDepth * kSlabBytes);
    auto issue_slab = [&](int sl, int slot) {
        if (sl < nslab) {
            const int col = sl / nblk, blk = sl - col * nblk;
            const int npiece = 4 * min(kSlabSub, ntt - kSlabSub * col);
            const int p = lane & (kSlabPieces - 1);
            const int64_t off = (int64_t)(t0 + kSlabSub * col) * (kSub / 4) + p * 16;
            if (p < npiece) {
                const uint8_t* base = packed + off;
                const bool inrow = off + 16 <= pitch;
                uint8_t* slab = ring + slot * kSlabBytes;
#pragma unroll 4
                for (int it = 0; it < kSlabIters; ++it) {
                    const int r = wl * 32 + lane / kSlabPieces + kSlabRowsPerInstr * it, b = blk * 128 + r;
                    if (b < B) {
                        uint8_t* dst = slab + slab_unit(r, p);
#ifndef NADM_KO_LOAD
                        if (inrow) cp_async16_full(dst, base + (uint64_t)rowoff[b] * (uint64_t)pitch);
                        else
#endif
                            *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            }
        }
        cp_async_commit();
    };
    if (warp < kProdWarps)                                           // in flight while the scale of V is computed
        for (int d = 0; d < kSlabDepth; ++d) issue_slab(g + 4 * d, d);

    // ---- |max| of this CTA's rows of V (fixed-point scale per CTA), accumulators zeroed ----
    {
        const int64_t m0 = (int64_t)t0 * kSub, m1 = min((int64_t)t1 * kSub, M);
        const int64_t n = (m1 > m0) ? (m1 - m0) * C : 0;
        const float* v0 = V + m0 * C;
        uint32_t mx = 0u;
        if ((reinterpret_cast<uintptr_t>(v0) & 15) == 0) {
            const int64_t n4 = n / 4;
#pragma unroll 8
            for (int64_t i = tid; i < n4; i += blockDim.x) mx = absbits_max4(mx, reinterpret_cast<const float4*>(v0)[i]);
            for (int64_t i = n4 * 4 + tid; i < n; i += blockDim.x) mx = max(mx, absbits(v0[i]));
        } else {
            for (int64_t i = tid; i < n; i += blockDim.x) mx = max(mx, absbits(v0[i]));
        }
        mx = warp_max_u32(mx);
        if (lane == 0) S->red[warp] = __uint_as_float(mx);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = S->tmem_base;
    float vmax;
    {
        uint32_t m = 0u;
        for (int w = 0; w < kSlabThreads / 32; ++w) m = max(m, __float_as_uint(S->red[w]));
        vmax = __uint_as_float(m);
    }
    if (tid == 0) cta_vmax[blockIdx.x] = vmax;
    if (warp < 4) {                                                  // warps 0..3 = lane quadrants 0..3
        uint32_t z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = 0u;
        for (int c0 = 0; c0 < nblk * 32; c0 += 16) tmem_st16(tbase + ((uint32_t)(warp * 32) << 16) + c0, z);
        tmem_wait_st();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    if (warp < kProdWarps) {
        // ---------------- producers: group g widens the four tiles of slabs g, g + 4, ... ----------------
        const int r = wl * 32 + lane;                                // my row of the block = my tensor-memory lane
        const uint32_t ta = tbase + ((uint32_t)(wl * 32) << 16) + kColA + g * 64;
        uint32_t phase = 1;
        int slot = 0;
        for (int sl = g; sl < nslab; sl += 4) {
            const int col = sl / nblk, blk = sl - col * nblk;
            const int nj = min(kSlabSub, ntt - kSlabSub * col);
            const bool active = blk * 128 + wl * 32 < B;             // warp-uniform: any real row in this warp
            const uint8_t* slab = ring + slot * kSlabBytes;
            cp_async_wait<kSlabDepth - 1>();
            __syncwarp();                                            // the other lanes' copies of my row have landed
#pragma unroll
            for (int j = 0; j < kSlabSub; ++j) {
                if (j < nj) {
                    uint4 w[4];
                    if (active) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) w[q] = *reinterpret_cast<const uint4*>(slab + slab_unit(r, 4 * j + q));
                    }
                    if (j == nj - 1) {
                        __syncwarp();                                // every lane has read the slab: refill it
                        issue_slab(sl + 4 * kSlabDepth, slot);
                        slot = (slot + 1 == kSlabDepth) ? 0 : slot + 1;
                    }
                    mbar_wait(&S->emptyA[g], phase);
                    if (active) {
                        tc_fence_after_sync();                       // the MMAs that read this stage have completed
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint32_t v[16];
                            widen_fields<MODE>(w[q].x, mvx, v);
                            widen_fields<MODE>(w[q].y, mvx, v + 4);
                            widen_fields<MODE>(w[q].z, mvx, v + 8);
                            widen_fields<MODE>(w[q].w, mvx, v + 12);
                            tmem_st16(ta + q * 16, v);
                        }
                        tmem_wait_st();
                        tc_fence_before_sync();
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S->fullA[g]);        // one arrival per warp of the group
                    phase ^= 1;
                }
            }
        }
        // ---------------- epilogue: recombine the digit planes, write this CTA's exact partial sums ----------------
        mbar_wait(&S->done, 0);
        tc_fence_after_sync();
        const int q = warp & 3;
        for (int blk = warp >> 2; blk < nblk; blk += kProdWarps / 4) {
            uint32_t v[32];
            tmem_ld32(tbase + ((uint32_t)(q * 32) << 16) + blk * 32, v);
            tmem_wait_ld();
            const int b = blk * 128 + q * 32 + lane;
            if (b < B) {
                long long* out = part + ((int64_t)blockIdx.x * B + b) * 8;
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    longlong2 z;
                    z.x = combine4(v, c);
                    z.y = combine4(v, c + 1);
                    *reinterpret_cast<longlong2*>(out + c) = z;
                }
            }
        }
    } else if (warp < kProdWarps + 2) {
        // ---------------- digit warps: int8 digit planes of V, one slab column (4 sub-tiles) ahead of the MMAs ----------------
        const FixScale fs = fix_scale(vmax);
        const int dt = tid - kProdThreads;                              // 0..63: K positions dt, dt+64, dt+128, dt+192
        auto load_v = [&](int tt, float (&v)[4][8]) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int pos = dt + 64 * e;
                const int64_t m = (int64_t)(t0 + tt) * kSub + (pos & ~15) + sigma16(pos & 15);
                if (m < M && C == 8) {
                    const float4 a = reinterpret_cast<const float4*>(V + m * 8)[0], b4 = reinterpret_cast<const float4*>(V + m * 8)[1];
                    v[e][0] = a.x; v[e][1] = a.y; v[e][2] = a.z; v[e][3] = a.w;
                    v[e][4] = b4.x; v[e][5] = b4.y; v[e][6] = b4.z; v[e][7] = b4.w;
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[e][c] = (m < M && c < C) ? V[m * C + c] : 0.f;
                }
            }
        };
        float vnext[4][8];
        load_v(0, vnext);
        for (int tt = 0; tt < ntt; ++tt) {
            const int col = tt / kSlabSub, j = tt % kSlabSub, set = col & 1;
            float v[4][8];
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int c = 0; c < 8; ++c) v[e][c] = vnext[e][c];
            if (tt + 1 < ntt) load_v(tt + 1, vnext);
            if (j == 0) mbar_wait_relaxed(&S->emptyV[set], ((col >> 1) & 1) ^ 1, 128);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int pos = dt + 64 * e;
                // scaled widening: K positions 4 k .. 4 k + 3 of every 16 hold 4^k x code -> digits of V / 4^k
                const float inv_pos = kScaled ? fs.inv * __uint_as_float((uint32_t)(127 - 2 * ((pos & 15) >> 2)) << 23) : fs.inv;
                store_digits(tilesV + (set * kSlabSub + j) * kDigTile + (pos & 7) * 16 + (pos >> 3) * 256, v[e], inv_pos);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->fullV[set][j]);
        }
    } else {
        // ---------------- MMA issuers: issuer `par` takes the slabs of producer groups 2 par, 2 par + 1 ----------------
        const int par = warp - (kProdWarps + 2);
        if (par < NISS) {
            const uint64_t B0 = smem_desc(smem_u32(tilesV), 256, 128);
            const uint32_t leader = elect_one() ? 1u : 0u;
            uint32_t phbits = 0u;                                        // phase of fullA[g], bit g
            for (int col = 0; col < ncol; ++col) {
                const int set = col & 1, nj = min(kSlabSub, ntt - kSlabSub * col);
                const uint32_t vph = (uint32_t)((col >> 1) & 1);
                for (int blk = 0; blk < nblk; ++blk) {
                    const int sl = col * nblk + blk;
                    if (NISS == 2 && ((sl >> 1) & 1) != par) continue;
                    const int gg = sl & 3;
                    const uint32_t d = tbase + blk * 32, at = tbase + kColA + gg * 64;
                    for (int j = 0; j < nj; ++j) {
                        mbar_wait(&S->fullV[set][j], vph);
                        mbar_wait(&S->fullA[gg], (phbits >> gg) & 1u);
                        tc_fence_after_sync();
                        const uint64_t b = B0 + (uint64_t)((set * kSlabSub + j) * (kDigTile >> 4));
#pragma unroll
                        for (int ks = 0; ks < kSub / 32; ++ks)
                            mma_i8_ts_p(d, at + ks * 8, b + (uint64_t)(ks * 64), kIdescFwd, 1u, leader);
                        mma_commit_p(&S->emptyA[gg], leader);
                        phbits ^= 1u << gg;
                    }
                }
                mma_commit_p(&S->emptyV[set], leader);
            }
            mma_commit_p(&S->done, leader);
        }
        __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kProdWarps + 2) tmem_dealloc<512>(tbase);
}

// Z[b, c] = 0.5 * sum over CTAs p of part_p[b, c] * 2^(e_p - 30): every CTA's partial is an exact integer at the CTA's own
// power-of-two scale (exactly representable in double, as is the product); the CTAs are summed in double in a fixed
// order (deterministic).  Block = the 8 components of one row x 32 part segments.
__global__ void __launch_bounds__(256)
enc_fwd_reduce_kernel(const long long* __restrict__ part, int nparts, int B, int C,
                      const float* __restrict__ cta_vmax, float* __restrict__ Z, double out_scale) {
    pdl_prologue();
    __shared__ double red[32][8];
    __shared__ double back[kMaxParts];
    const bool scaled = (cta_vmax == nullptr);          // partials already multiplied by their CTA's scale (doubles)
    if (!scaled) {
        for (int p = threadIdx.x; p < nparts; p += blockDim.x) back[p] = fix_scale(cta_vmax[p]).back;
        __syncthreads();
    }
    const int c = threadIdx.x & 7, seg = threadIdx.x >> 3;
    const int b = blockIdx.x;
    const int64_t i = (int64_t)b * 8 + c;
    double acc = 0.0;
    if (scaled)
        for (int p = seg; p < nparts; p += 32) acc += __longlong_as_double(part[(int64_t)p * B * 8 + i]);
    else
        for (int p = seg; p < nparts; p += 32) acc += (double)part[(int64_t)p * B * 8 + i] * back[p];
    red[seg][c] = acc;
    __syncthreads();
    if (threadIdx.x < 8 && c < C) {
        double t = 0.0;
#pragma unroll
        for (int s2 = 0; s2 < 32; ++s2) t += red[s2][c];
        Z[(int64_t)b * C + c] = (float)(t * out_scale);
    }
}

// =================================================================================================================
// backward: dV = X^T dZ, Adam on V
// =================================================================================================================
struct EncBwdSmem {
    uint64_t fullA[kAStages], emptyA[kAStages], dfull[2], dempty[2];
    uint32_t tmem_base;
    float red[32];
};

// warps: 16 producers, 2 MMA issuers (issuer p owns stages 2p, 2p+1 and its own accumulator set, as in the forward), 4 epilogue
__host__ __device__ constexpr int bwd_threads(int EPW) { return kProdThreads + 64 + EPW * 32; }
// NISS: MMA issuer warps in use (2, or 1: issuer 0 then issues both halves).  EPW: epilogue warps, 4 (one thread does both
// 128-SNP halves of a sub-tile) or 8 (one half each: twice the loads of V / m / v in flight)
template <int NISS, bool RAW, int EPW>
__global__ void __launch_bounds__(bwd_threads(EPW), 1)
enc_bwd_tc_kernel(const uint8_t* __restrict__ packed, int64_t pitch, const int64_t* __restrict__ row_idx, int64_t row0,
                  int B, int64_t M, const float* __restrict__ dZ, int C, float* __restrict__ V, float* __restrict__ Vm,
                  float* __restrict__ Vv, AdamCoef adam_in, float* __restrict__ dV_out, int T, uint32_t mvx,
                  double out_scale, int accumulate, ApplyJob job) {
    // job.nslab > 0: the pending parameter update of the small network (see DeferredApply) runs on this kernel's epilogue
    // warps before their first accumulators arrive
    pdl_prologue();
    const AdamCoef adam = adam_resolve(adam_in);
    extern __shared__ __align__(1024) uint8_t smem[];
    const int nblk = (B + 127) / 128;
    uint8_t* tilesA = smem;                                         // kAStages x 32 KB
    uint8_t* digZ = tilesA + kAStages * kATile;                     // nblk x 4 KB: dZ digits, K position = batch row
    uint8_t* stage = digZ + nblk * 4096;                            // kStDepth x 10 KB packed rows
    uint32_t* rowoff = reinterpret_cast<uint32_t*>(stage + 4 * kStDepth * kStTile);
    EncBwdSmem* S = reinterpret_cast<EncBwdSmem*>(rowoff + ((B + 3) & ~3));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t0 = (int)(((int64_t)T * blockIdx.x) / gridDim.x), t1 = (int)(((int64_t)T * (blockIdx.x + 1)) / gridDim.x);
    const int ntile = (t1 - t0) * nblk;

    for (int b = tid; b < B; b += blockDim.x)
        rowoff[b] = (uint32_t)((row_idx != nullptr) ? row_idx[b] : (row0 + b));
    if (tid == 0) {
        for (int s = 0; s < kAStages; ++s) { mbar_init(&S->fullA[s], 4); mbar_init(&S->emptyA[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&S->dfull[s], NISS); mbar_init(&S->dempty[s], EPW); }
        mbar_init_fence();
    }
    if (warp == kProdWarps) tmem_alloc<256>(&S->tmem_base);
    __syncthreads();                                                // row numbers are in shared memory
    // the producers' first copies are in flight while the digits of dZ are set up (they need the row numbers only)
    Feed f{packed, pitch, rowoff, B, nblk, t0, ntile, stage + (warp >> 2 & 3) * (kStDepth * kStTile), (warp >> 2) % nblk,
           (warp >> 2) / nblk, warp >> 2, 0, kStDepth};
    if (warp < kProdWarps)
        for (int p = 0; p < kStDepth; ++p) feed_issue(f, warp & 3, lane);
    // |max| of dZ over the batch (every CTA computes the same value)
    uint32_t mxb = 0u;
    if ((reinterpret_cast<uintptr_t>(dZ) & 15) == 0 && ((B * C) & 3) == 0) {
#pragma unroll 4
        for (int i = tid; i < (B * C) / 4; i += blockDim.x) mxb = absbits_max4(mxb, reinterpret_cast<const float4*>(dZ)[i]);
    } else {
        for (int i = tid; i < B * C; i += blockDim.x) mxb = max(mxb, absbits(dZ[i]));
    }
    mxb = warp_max_u32(mxb);
    if (lane == 0) S->red[warp] = __uint_as_float(mxb);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    mxb = 0u;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mxb = max(mxb, __float_as_uint(S->red[w]));
    const FixScale fs = fix_scale(__uint_as_float(mxb));
    for (int b = tid; b < nblk * 128; b += blockDim.x) {
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = (b < B && c < C) ? dZ[(int64_t)b * C + c] : 0.f;
        store_digits(digZ + (b & 7) * 16 + (b >> 3) * 256, v, fs.inv);
    }
    fence_async_smem();
    __syncthreads();
    const uint32_t tbase = S->tmem_base;

    if (warp < kProdWarps) {
        // ---------------- producers: widened genotype tiles, order (sub-tile, block) ----------------
        const int g = warp >> 2, wl = warp & 3;
        int blk = g % nblk, slot = 0, phase = 1;
        for (int i = g; i < ntile; i += 4) {
            uint4 w[4];
            feed_load(f, slot, wl, lane, w);
            feed_issue(f, wl, lane);                                 // refill the staging slot just read
            mbar_wait(&S->emptyA[g], phase);
            feed_store<RAW>(tilesA + g * kATile, wl, lane, w, mvx);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->fullA[g]);
            phase ^= 1;
            slot = (slot + 1 == kStDepth) ? 0 : slot + 1;
            blk += 4;
            while (blk >= nblk) blk -= nblk;
        }
    } else if (warp <= kProdWarps + 1) {
        // ---------------- MMA issuers: issuer p = warp - kProdWarps takes the tiles of its two stages ((i >> 1) & 1 == p;
        // with >= 3 row blocks every sub-tile has tiles of both) and accumulates both 128-SNP halves into its own set:
        // tensor-memory columns (buf NISS + p) 64 + h 32.  Convergent loop, elected lane issues ----------------
        const int p = warp - kProdWarps;
        if (p < NISS) {
            const uint64_t A0 = smem_desc(smem_u32(tilesA), 2048, 128), B0 = smem_desc(smem_u32(digZ), 256, 128);
            const int nks_last = min(4, (B - (nblk - 1) * 128 + 31) / 32);  // K steps holding real batch rows
            const uint32_t leader = elect_one() ? 1u : 0u;
            int i = 0;
            for (int tt = 0; tt < t1 - t0; ++tt) {
                const int buf = tt & 1;
                mbar_wait(&S->dempty[buf], ((tt >> 1) & 1) ^ 1);
                uint32_t acc0 = 0u;
                for (int blk = 0; blk < nblk; ++blk, ++i) {
                    if (NISS == 2 && ((i >> 1) & 1) != p) continue;
                    const int s = i & (kAStages - 1);
                    mbar_wait(&S->fullA[s], (i >> 2) & 1);
                    tc_fence_after_sync();
                    const uint64_t a = A0 + (uint64_t)(s * (kATile >> 4)), b = B0 + (uint64_t)(blk * 256);
                    const uint32_t d = tbase + (buf * NISS + p) * 64;
                    if (blk != nblk - 1 || nks_last == 4) {
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                mma_i8_ss_p(d + h * 32, a + (uint64_t)(h * 64 + ks * 512), b + (uint64_t)(ks * 64), kIdescBwd,
                                            ks ? 1u : acc0, leader);
                    } else {
                        for (int h = 0; h < 2; ++h)
                            for (int ks = 0; ks < nks_last; ++ks)
                                mma_i8_ss_p(d + h * 32, a + (uint64_t)(h * 64 + ks * 512), b + (uint64_t)(ks * 64), kIdescBwd,
                                            ks ? 1u : acc0, leader);
                    }
                    acc0 = 1u;
                    mma_commit_p(&S->emptyA[s], leader);
                }
                mma_commit_p(&S->dfull[buf], leader);
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue: digit planes -> dV -> Adam on V, one SNP per thread ----------------
        const int q = warp & 3;             // tensor-memory lane quadrant this warp may access (warp id % 4)
        constexpr int kHalves = (EPW == 8) ? 1 : 2;                   // 128-SNP halves of a sub-tile per thread
        const int hsel = (EPW == 8) ? ((warp - (kProdWarps + 2)) >> 2) : 0;
        if (job.nslab > 0)      // this CTA's share of the network's parameter update: nothing here depends on it
            mlp_apply_share(job, (int)blockIdx.x, (int)gridDim.x, tid - (kProdWarps + 2) * 32, EPW * 32);
        for (int tt = 0; tt < t1 - t0; ++tt) {
            const int buf = tt & 1;
            // (measured and kept out, profiles/r2_*: with Adam this kernel takes 85 us against 60 without.  Requesting the
            //  sub-tile's V / m / v ahead of the wait for its accumulators did not help — into registers: 128 us, the 48
            //  extra live registers spill in this 704-thread kernel; as L2 prefetches, this or two sub-tiles ahead: 88 us.)
            if (warp == kProdWarps + 2) mbar_wait_relaxed(&S->dfull[buf], (tt >> 1) & 1, 64);   // one warp polls
            named_bar_sync(1, EPW * 32);
            tc_fence_after_sync();
            // recombine the digit planes of each accumulator set right after loading it (int64, exact) and add the sets:
            // 32 live accumulator registers at a time instead of 96
            long long z[kHalves][8];
#pragma unroll
            for (int hi_ = 0; hi_ < kHalves; ++hi_) {
                const int h = hsel + hi_;
#pragma unroll
                for (int c = 0; c < 8; ++c) z[hi_][c] = 0;
#pragma unroll
                for (int set = 0; set < NISS; ++set) {
                    uint32_t v[32];
                    tmem_ld32(tbase + ((uint32_t)(q * 32) << 16) + (buf * NISS + set) * 64 + h * 32, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 8; ++c) z[hi_][c] += combine4(v, c);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->dempty[buf]);
#pragma unroll
            for (int hi_ = 0; hi_ < kHalves; ++hi_) {
                const int h = hsel + hi_;
                const int pos = q * 32 + lane;
                const int64_t m = (int64_t)(t0 + tt) * kSub + h * 128 + (pos & ~15) + sigma16(pos & 15);
                if (m >= M) continue;
                float g[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) g[c] = (float)((double)z[hi_][c] * fs.back * out_scale);
                if (C == 8) {
                    float4* gv = reinterpret_cast<float4*>(g);
                    if (dV_out != nullptr) {
                        if (accumulate) {                              // += over successive row batches (nadm_geno_matmul_t)
                            const float4 a0 = reinterpret_cast<float4*>(dV_out + m * 8)[0];
                            const float4 a1 = reinterpret_cast<float4*>(dV_out + m * 8)[1];
                            gv[0] = make_float4(gv[0].x + a0.x, gv[0].y + a0.y, gv[0].z + a0.z, gv[0].w + a0.w);
                            gv[1] = make_float4(gv[1].x + a1.x, gv[1].y + a1.y, gv[1].z + a1.z, gv[1].w + a1.w);
                        }
                        reinterpret_cast<float4*>(dV_out + m * 8)[0] = gv[0];
                        reinterpret_cast<float4*>(dV_out + m * 8)[1] = gv[1];
                    }
                    if (adam.enabled) {
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            float4 p4 = reinterpret_cast<float4*>(V + m * 8)[hh];
                            float4 m4 = reinterpret_cast<float4*>(Vm + m * 8)[hh];
                            float4 v4 = reinterpret_cast<float4*>(Vv + m * 8)[hh];
                            p4.x = adam_apply(p4.x, g[hh * 4 + 0], m4.x, v4.x, adam);
                            p4.y = adam_apply(p4.y, g[hh * 4 + 1], m4.y, v4.y, adam);
                            p4.z = adam_apply(p4.z, g[hh * 4 + 2], m4.z, v4.z, adam);
                            p4.w = adam_apply(p4.w, g[hh * 4 + 3], m4.w, v4.w, adam);
                            reinterpret_cast<float4*>(V + m * 8)[hh] = p4;
                            reinterpret_cast<float4*>(Vm + m * 8)[hh] = m4;
                            reinterpret_cast<float4*>(Vv + m * 8)[hh] = v4;
                        }
                    }
                } else {
                    for (int c = 0; c < C; ++c) {
                        const int64_t vi = m * C + c;
                        if (dV_out != nullptr) dV_out[vi] = accumulate ? dV_out[vi] + g[c] : g[c];
                        if (adam.enabled) {
                            float mm = Vm[vi], vv = Vv[vi];
                            V[vi] = adam_apply(V[vi], g[c], mm, vv, adam);
                            Vm[vi] = mm;
                            Vv[vi] = vv;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kProdWarps) tmem_dealloc<256>(tbase);
}

// =================================================================================================================
// backward, slab feed (OPT-IN: NADM_ENC_BWD_SLAB=1; measured 82 vs 60 us for the kernel above — the half-tile stages
// double the hand-overs —, kept as a measurement switch): dV = X^T dZ + Adam with the rows streamed in wide runs (see the
// forward slab kernel for why).  The operand X^T must come from shared memory (SNPs on the M axis), so the widened stages stay
// there — but as HALF tiles (128 rows x 128 SNPs, 16 KB) so that four 32 KB slabs fit beside them.  Order: slab column
// (4 sub-tiles) outer, row block, sub-tile, half inner; the 4 x 2 accumulators (128 SNPs x 32 digit columns) of a
// column are shared by both issuers (zero-initialised by the epilogue, every MMA accumulates: integer adds commute)
// and double-buffered against the epilogue: 2 x 8 x 32 = 512 tensor-memory columns.
// =================================================================================================================
constexpr int kHalfTile = 128 * 128;               // bytes of one widened half tile
struct EncBwdSlabSmem {
    uint64_t fullA[4], emptyA[4], dfull[2], dempty[2];
    uint32_t tmem_base;
    float red[32];
};
constexpr int kBwdSlabThreads = kProdThreads + 64 + 128;

template <int NISS, int MODE>
__global__ void __launch_bounds__(kBwdSlabThreads, 1)
enc_bwd_slab_kernel(const uint8_t* __restrict__ packed, int64_t pitch, const int64_t* __restrict__ row_idx, int64_t row0,
                    int B, int64_t M, const float* __restrict__ dZ, int C, float* __restrict__ V, float* __restrict__ Vm,
                    float* __restrict__ Vv, AdamCoef adam_in, float* __restrict__ dV_out, int T, uint32_t mvx,
                    double out_scale, int accumulate) {
    pdl_prologue();
    constexpr bool kScaled = MODE != kModeRawAny;
    const AdamCoef adam = adam_resolve(adam_in);
    extern __shared__ __align__(1024) uint8_t smem[];
    const int nblk = (B + 127) / 128;
    uint8_t* tilesA = smem;                                          // 4 stages x 16 KB (one per producer group)
    uint8_t* digZ = tilesA + 4 * kHalfTile;                          // nblk x 4 KB: dZ digits, K position = batch row
    uint8_t* slabs = digZ + nblk * 4096;                             // 4 groups x kSlabDepth slabs of packed rows
    uint32_t* rowoff = reinterpret_cast<uint32_t*>(slabs + 4 * kSlabDepth * kSlabBytes);
    EncBwdSlabSmem* S = reinterpret_cast<EncBwdSlabSmem*>(rowoff + ((B + 3) & ~3));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t0 = (int)(((int64_t)T * blockIdx.x) / gridDim.x), t1 = (int)(((int64_t)T * (blockIdx.x + 1)) / gridDim.x);
    const int ntt = t1 - t0, ncol = (ntt + kSlabSub - 1) / kSlabSub, nslab = ncol * nblk;

    for (int b = tid; b < B; b += blockDim.x)
        rowoff[b] = (uint32_t)((row_idx != nullptr) ? row_idx[b] : (row0 + b));
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) { mbar_init(&S->fullA[s], 4); mbar_init(&S->emptyA[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&S->dfull[s], NISS); mbar_init(&S->dempty[s], 4); }
        mbar_init_fence();
    }
    if (warp == kProdWarps) tmem_alloc<512>(&S->tmem_base);
    __syncthreads();                                                 // row numbers are in shared memory

    const int g = warp >> 2, wl = warp & 3;                          // (producer warps only)
    uint8_t* ring = slabs + (g & 3) * (kSlabDepth * kSlabBytes);
    auto issue_slab = [&](int sl, int slot) {
        if (sl < nslab) {
            const int col = sl / nblk, blk = sl - col * nblk;
            const int npiece = 4 * min(kSlabSub, ntt - kSlabSub * col);
            const int p = lane & (kSlabPieces - 1);
            const int64_t off = (int64_t)(t0 + kSlabSub * col) * (kSub / 4) + p * 16;
            if (p < npiece) {
                const uint8_t* base = packed + off;
                const bool inrow = off + 16 <= pitch;
                uint8_t* slab = ring + slot * kSlabBytes;
#pragma unroll 4
                for (int it = 0; it < kSlabIters; ++it) {
                    const int r = wl * 32 + lane / kSlabPieces + kSlabRowsPerInstr * it, b = blk * 128 + r;
                    if (b < B) {
                        uint8_t* dst = slab + slab_unit(r, p);
#ifndef NADM_KO_LOAD
                        if (inrow) cp_async16_full(dst, base + (uint64_t)rowoff[b] * (uint64_t)pitch);
                        else
#endif
                            *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            }
        }
        cp_async_commit();
    };
    if (warp < kProdWarps)                                           // in flight during the set-up below
        for (int d = 0; d < kSlabDepth; ++d) issue_slab(g + 4 * d, d);

    // |max| of dZ over the batch (every CTA computes the same value), digits of dZ
    uint32_t mxb = 0u;
    if ((reinterpret_cast<uintptr_t>(dZ) & 15) == 0 && ((B * C) & 3) == 0) {
#pragma unroll 4
        for (int i = tid; i < (B * C) / 4; i += blockDim.x) mxb = absbits_max4(mxb, reinterpret_cast<const float4*>(dZ)[i]);
    } else {
        for (int i = tid; i < B * C; i += blockDim.x) mxb = max(mxb, absbits(dZ[i]));
    }
    mxb = warp_max_u32(mxb);
    if (lane == 0) S->red[warp] = __uint_as_float(mxb);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    mxb = 0u;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mxb = max(mxb, __float_as_uint(S->red[w]));
    const FixScale fs = fix_scale(__uint_as_float(mxb));
    for (int b = tid; b < nblk * 128; b += blockDim.x) {
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = (b < B && c < C) ? dZ[(int64_t)b * C + c] : 0.f;
        store_digits(digZ + (b & 7) * 16 + (b >> 3) * 256, v, fs.inv);
    }
    const uint32_t tbase = S->tmem_base;
    if (warp < 4) {                                                  // zero both accumulator buffers (512 columns)
        uint32_t z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = 0u;
        for (int c0 = 0; c0 < 512; c0 += 16) tmem_st16(tbase + ((uint32_t)(warp * 32) << 16) + c0, z);
        tmem_wait_st();
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    if (warp < kProdWarps) {
        // ---------------- producers: group g widens the eight half tiles of slabs g, g + 4, ... ----------------
        const int r8 = lane & 7, qq = (lane >> 3) & 1, gsel = lane >> 4;
        uint8_t* tile = tilesA + g * kHalfTile;
        uint32_t phase = 1;
        int slot = 0;
        for (int sl = g; sl < nslab; sl += 4) {
            const int col = sl / nblk, blk = sl - col * nblk;
            const int nj = min(kSlabSub, ntt - kSlabSub * col);
            const bool active = blk * 128 + wl * 32 < B;             // warp-uniform: any real row in this warp
            const uint8_t* slab = ring + slot * kSlabBytes;
            cp_async_wait<kSlabDepth - 1>();
            __syncwarp();                                            // this warp's copies of its 32 rows have landed
#pragma unroll
            for (int j = 0; j < kSlabSub; ++j) {
                if (j < nj) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4 w[2];
                        if (active) {
#pragma unroll
                            for (int it = 0; it < 2; ++it)
                                w[it] = *reinterpret_cast<const uint4*>(
                                    slab + slab_unit((wl * 4 + gsel * 2 + it) * 8 + r8, 4 * j + 2 * h + qq));
                        }
                        if (j == nj - 1 && h == 1) {
                            __syncwarp();                            // every lane has read the slab: refill it
                            issue_slab(sl + 4 * kSlabDepth, slot);
                            slot = (slot + 1 == kSlabDepth) ? 0 : slot + 1;
                        }
                        mbar_wait(&S->emptyA[g], phase);
                        if (active) {
#pragma unroll
                            for (int it = 0; it < 2; ++it) {
                                const int g8 = wl * 4 + gsel * 2 + it;                       // 8-row group of the block
                                uint8_t* dst = tile + r8 * 16 + g8 * 1024 + (qq * 4) * 128;
                                uint32_t o[4];
                                widen_fields<MODE>(w[it].x, mvx, o);
                                *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
                                widen_fields<MODE>(w[it].y, mvx, o);
                                *reinterpret_cast<uint4*>(dst + 128) = make_uint4(o[0], o[1], o[2], o[3]);
                                widen_fields<MODE>(w[it].z, mvx, o);
                                *reinterpret_cast<uint4*>(dst + 256) = make_uint4(o[0], o[1], o[2], o[3]);
                                widen_fields<MODE>(w[it].w, mvx, o);
                                *reinterpret_cast<uint4*>(dst + 384) = make_uint4(o[0], o[1], o[2], o[3]);
                            }
                            fence_async_smem();
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&S->fullA[g]);
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp <= kProdWarps + 1) {
        // ---------------- MMA issuers: issuer p takes the slabs of producer groups 2 p, 2 p + 1 ----------------
        const int par = warp - kProdWarps;
        if (par < NISS) {
            const uint64_t A0 = smem_desc(smem_u32(tilesA), 1024, 128), B0 = smem_desc(smem_u32(digZ), 256, 128);
            const int nks_last = min(4, (B - (nblk - 1) * 128 + 31) / 32);  // K steps holding real batch rows
            const uint32_t leader = elect_one() ? 1u : 0u;
            uint32_t phbits = 0u;
            for (int col = 0; col < ncol; ++col) {
                const int buf = col & 1, nj = min(kSlabSub, ntt - kSlabSub * col);
                mbar_wait(&S->dempty[buf], ((col >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                for (int blk = 0; blk < nblk; ++blk) {
                    const int sl = col * nblk + blk;
                    if (NISS == 2 && ((sl >> 1) & 1) != par) continue;
                    const int gg = sl & 3;
                    const uint64_t a = A0 + (uint64_t)(gg * (kHalfTile >> 4)), b = B0 + (uint64_t)(blk * 256);
                    const int nks = (blk == nblk - 1) ? nks_last : 4;
                    for (int j = 0; j < nj; ++j) {
                        for (int h = 0; h < 2; ++h) {
                            mbar_wait(&S->fullA[gg], (phbits >> gg) & 1u);
                            tc_fence_after_sync();
                            const uint32_t d = tbase + ((buf * kSlabSub + j) * 2 + h) * 32;
                            if (nks == 4) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    mma_i8_ss_p(d, a + (uint64_t)(ks * 256), b + (uint64_t)(ks * 64), kIdescBwd, 1u, leader);
                            } else {
                                for (int ks = 0; ks < nks; ++ks)
                                    mma_i8_ss_p(d, a + (uint64_t)(ks * 256), b + (uint64_t)(ks * 64), kIdescBwd, 1u, leader);
                            }
                            mma_commit_p(&S->emptyA[gg], leader);
                            phbits ^= 1u << gg;
                        }
                    }
                }
                mma_commit_p(&S->dfull[buf], leader);
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue: digit planes -> dV -> Adam on V, one SNP per thread and half tile ----------------
        const int q = warp & 3;             // tensor-memory lane quadrant this warp may access (warp id % 4)
        for (int col = 0; col < ncol; ++col) {
            const int buf = col & 1, nj = min(kSlabSub, ntt - kSlabSub * col);
            if (warp == kProdWarps + 2) mbar_wait_relaxed(&S->dfull[buf], (col >> 1) & 1, 64);   // one warp polls
            named_bar_sync(1, 128);
            tc_fence_after_sync();
            for (int j = 0; j < nj; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(tbase + ((uint32_t)(q * 32) << 16) + ((buf * kSlabSub + j) * 2 + h) * 32, v);
                    tmem_wait_ld();
                    const int pos = q * 32 + lane;
                    const int64_t m = (int64_t)(t0 + kSlabSub * col + j) * kSub + h * 128 + (pos & ~15) + sigma16(pos & 15);
                    if (m >= M) continue;
                    // scaled widening: this SNP's bytes were 4^k x code, k = its field index inside the packed byte
                    const double sc = fs.back * out_scale *
                        (kScaled ? __longlong_as_double((long long)(1023 - 2 * ((pos & 15) >> 2)) << 52) : 1.0);
                    float gr[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) gr[c] = (float)((double)combine4(v, c) * sc);
                    if (C == 8) {
                        float4* gv = reinterpret_cast<float4*>(gr);
                        if (dV_out != nullptr) {
                            if (accumulate) {                              // += over successive row batches (nadm_geno_matmul_t)
                                const float4 a0 = reinterpret_cast<float4*>(dV_out + m * 8)[0];
                                const float4 a1 = reinterpret_cast<float4*>(dV_out + m * 8)[1];
                                gv[0] = make_float4(gv[0].x + a0.x, gv[0].y + a0.y, gv[0].z + a0.z, gv[0].w + a0.w);
                                gv[1] = make_float4(gv[1].x + a1.x, gv[1].y + a1.y, gv[1].z + a1.z, gv[1].w + a1.w);
                            }
                            reinterpret_cast<float4*>(dV_out + m * 8)[0] = gv[0];
                            reinterpret_cast<float4*>(dV_out + m * 8)[1] = gv[1];
                        }
                        if (adam.enabled) {
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                float4 p4 = reinterpret_cast<float4*>(V + m * 8)[hh];
                                float4 m4 = reinterpret_cast<float4*>(Vm + m * 8)[hh];
                                float4 v4 = reinterpret_cast<float4*>(Vv + m * 8)[hh];
                                p4.x = adam_apply(p4.x, gr[hh * 4 + 0], m4.x, v4.x, adam);
                                p4.y = adam_apply(p4.y, gr[hh * 4 + 1], m4.y, v4.y, adam);
                                p4.z = adam_apply(p4.z, gr[hh * 4 + 2], m4.z, v4.z, adam);
                                p4.w = adam_apply(p4.w, gr[hh * 4 + 3], m4.w, v4.w, adam);
                                reinterpret_cast<float4*>(V + m * 8)[hh] = p4;
                                reinterpret_cast<float4*>(Vm + m * 8)[hh] = m4;
                                reinterpret_cast<float4*>(Vv + m * 8)[hh] = v4;
                            }
                        }
                    } else {
                        for (int c = 0; c < C; ++c) {
                            const int64_t vi = m * C + c;
                            if (dV_out != nullptr) dV_out[vi] = accumulate ? dV_out[vi] + gr[c] : gr[c];
                            if (adam.enabled) {
                                float mm = Vm[vi], vv = Vv[vi];
                                V[vi] = adam_apply(V[vi], gr[c], mm, vv, adam);
                                Vm[vi] = mm;
                                Vv[vi] = vv;
                            }
                        }
                    }
                }
            }
            // hand the buffer back zeroed: both issuers accumulate into it from the first MMA on
            {
                uint32_t z[16];
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) z[jj] = 0u;
                for (int c0 = 0; c0 < kSlabSub * 64; c0 += 16)
                    tmem_st16(tbase + ((uint32_t)(q * 32) << 16) + buf * (kSlabSub * 64) + c0, z);
                tmem_wait_st();
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->dempty[buf]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kProdWarps) tmem_dealloc<512>(tbase);
}

bool enc_bwd_slab_supported(int B) {
    const int nblk = (B + 127) / 128;
    return (size_t)4 * kHalfTile + (size_t)nblk * 4096 + (size_t)4 * kSlabDepth * kSlabBytes + (size_t)((B + 3) & ~3) * 4 +
               sizeof(EncBwdSlabSmem) + 64 <= (size_t)kMaxDynSmem;
}

// =================================================================================================================
// host launchers (called from the C ABI in nadm_stream.cu)
// =================================================================================================================
// MMA issuer warps per encoder kernel: NADM_ENC_ISSUERS=1|2 (A/B measurements; default below)
// NADM_ENC_TS=1: forward encoder with its genotype operand in tensor memory (TSA).  Off by default until it has been
// through the full parity suite on the device.
static bool enc_fwd_tmem_operand() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("NADM_ENC_TS");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
// NADM_ENC_FWD_SLAB=1: the slab-fed forward kernel (128-byte runs per row, operand in tensor memory, scaled widening).
// Same-box A/B inside the training step (profiles/r2_*): 0.4428 vs 0.4425 ms per step — no gain where it matters (every
// step gathers rows that are not in L2), and the forward-only Q pass loses (1024 instead of 2048 rows per launch), so it
// stays opt-in for measurements.
bool enc_fwd_slab() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("NADM_ENC_FWD_SLAB");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
// NADM_ENC_BWD_SLAB=1: the slab-fed backward kernel.  Measured SLOWER than the round-1 backward kernel (82 vs 60 us at
// cfg3, profiles/r2_encoder_slab_variants.txt: its half-tile stages double the producer / issuer hand-overs, and that
// costs more than the 128-byte runs gain), so it stays opt-in for measurements.
static bool enc_bwd_slab() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("NADM_ENC_BWD_SLAB");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
bool enc_bwd_runs_apply(int B) { return !(enc_bwd_slab() && enc_bwd_slab_supported(B)); }
static int enc_issuers() {
    static int n = 0;
    if (n == 0) {
        const char* e = getenv("NADM_ENC_ISSUERS");
        n = (e != nullptr && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 2;
    }
    return n;
}

bool enc_bwd_tc_supported(int B) {
    const int nblk = (B + 127) / 128;
    return (size_t)kAStages * kATile + (size_t)nblk * 4096 + (size_t)4 * kStDepth * kStTile + (size_t)((B + 3) & ~3) * 4 +
               sizeof(EncBwdSmem) + 64 <= (size_t)kMaxDynSmem;
}
constexpr size_t kEncWsHeader = 4096;   // per-CTA scales in front of the partial sums
size_t enc_tc_workspace_bytes(int B) { return (size_t)sm_count() * (size_t)B * 8 * sizeof(long long) + kEncWsHeader; }

// raw_mv < 0: training semantics (x = code / 2, missing -> 0).  raw_mv >= 0: raw uint8 values, code 3 -> raw_mv.
int launch_enc_fwd_tc(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int B, int64_t M,
                      const float* V, int C, float* Z, void* ws, size_t ws_bytes, cudaStream_t st, int raw_mv, bool defer) {
    const int T = (int)((M + kSub - 1) / kSub);
    const int ncta = std::min(T, sm_count());
    const size_t need = (size_t)ncta * B * 8 * sizeof(long long) + kEncWsHeader;
    NADM_REQUIRE(need <= ws_bytes, "workspace too small for encoder_fwd (%zu > %zu)", need, ws_bytes);
    float* vmax = reinterpret_cast<float*>(ws);                     // per-CTA |max| of its slice of V (ncta floats)
    long long* part = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(ws) + kEncWsHeader);
    const int nblk_ = (B + 127) / 128;
    if (enc_fwd_slab() && nblk_ <= 8) {
        // ---- slab-fed kernel (default): 256-byte runs per row, genotype operand in tensor memory ----
        NADM_REQUIRE((M + ncta - 1) / ncta <= 65536, "M=%lld: more than 65536 SNPs per CTA would overflow the int32 accumulators",
                     (long long)M);
        const size_t smem_s = (size_t)2 * kSlabSub * kDigTile + (size_t)4 * kSlabDepth * kSlabBytes + (size_t)((B + 3) & ~3) * 4 +
                              sizeof(EncSlabSmem) + 64;
        static PerDeviceOnce once_s;
        bool* as = once_s.slot();
        if (as == nullptr || !*as) {
            cudaError_t e = cudaSuccess;
#define NADM_SLAB_ATTR(N_, M_)                                                                                         \
    if (e == cudaSuccess)                                                                                              \
        e = cudaFuncSetAttribute(enc_fwd_slab_kernel<N_, M_>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)
            NADM_SLAB_ATTR(1, kModeTrain); NADM_SLAB_ATTR(2, kModeTrain); NADM_SLAB_ATTR(1, kModeRaw3);
            NADM_SLAB_ATTR(2, kModeRaw3); NADM_SLAB_ATTR(1, kModeRawAny); NADM_SLAB_ATTR(2, kModeRawAny);
#undef NADM_SLAB_ATTR
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(enc_fwd_slab)");
            if (as) *as = true;
        }
        const bool two_s = enc_issuers() == 2 && nblk_ >= 3;
        const uint32_t mvx_s = raw_mv >= 0 ? (3u ^ (uint32_t)(raw_mv & 0xFF)) : 0u;
        const int mode = raw_mv < 0 ? kModeTrain : (raw_mv == 3 ? kModeRaw3 : kModeRawAny);
#define NADM_SLAB_GO(N_, M_)                                                                                           \
    launch_pdl(enc_fwd_slab_kernel<N_, M_>, dim3(ncta), dim3(kSlabThreads), smem_s, st, packed, pitch, row_idx, row0, B, \
               M, V, C, vmax, part, T, mvx_s)
        if (mode == kModeTrain) { if (two_s) NADM_SLAB_GO(2, kModeTrain); else NADM_SLAB_GO(1, kModeTrain); }
        else if (mode == kModeRaw3) { if (two_s) NADM_SLAB_GO(2, kModeRaw3); else NADM_SLAB_GO(1, kModeRaw3); }
        else { if (two_s) NADM_SLAB_GO(2, kModeRawAny); else NADM_SLAB_GO(1, kModeRawAny); }
#undef NADM_SLAB_GO
        NADM_CHECK_LAUNCH("enc_fwd_slab_kernel");
        launch_pdl(enc_fwd_reduce_kernel, dim3(B), dim3(256), 0, st, part, ncta, B, C, vmax, Z, raw_mv >= 0 ? 1.0 : 0.5);
        NADM_CHECK_LAUNCH("enc_fwd_reduce_kernel");
        return NADM_OK;
    }
    const bool tsa = enc_fwd_tmem_operand() && nblk_ <= 8;           // accumulators + 4 tiles must fit 512 columns
    const size_t smem = (tsa ? (size_t)4 * kStDepthTS * kStTile : (size_t)kAStages * kATile + (size_t)4 * kStDepth * kStTile) +
                        2 * kDigTile + (size_t)((B + 3) & ~3) * 4 + sizeof(EncSmem) + 64;
    static PerDeviceOnce once;
    bool* attr = once.slot();
    if (attr == nullptr || !*attr) {
        cudaError_t e = cudaSuccess;
#define NADM_FWD_ATTR(N_, R_, T_)                                                                                      \
    if (e == cudaSuccess)                                                                                              \
        e = cudaFuncSetAttribute(enc_fwd_tc_kernel<N_, R_, T_>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)
        NADM_FWD_ATTR(1, false, false); NADM_FWD_ATTR(2, false, false); NADM_FWD_ATTR(1, true, false); NADM_FWD_ATTR(2, true, false);
        NADM_FWD_ATTR(1, false, true); NADM_FWD_ATTR(2, false, true); NADM_FWD_ATTR(1, true, true); NADM_FWD_ATTR(2, true, true);
#undef NADM_FWD_ATTR
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(enc_fwd_tc)");
        if (attr) *attr = true;
    }
    const bool two = enc_issuers() == 2 && nblk_ >= 3 && nblk_ <= 8;   // two accumulator sets: 2 x 32 columns per row block
    const uint32_t mvx = raw_mv >= 0 ? (3u ^ (uint32_t)(raw_mv & 0xFF)) : 0u;
    // the CTAs reduce their partials themselves behind a grid barrier (cooperative launch) only with NADM_GRIDBAR=1 (measured slower)
    const bool fused = gridbar_enabled();
    const double out_scale = raw_mv >= 0 ? 1.0 : 0.5;
    float* Zk = fused ? Z : nullptr;
    cudaError_t le = cudaSuccess;
#define NADM_FWD_GO2(N_, R_, T_)                                                                                       \
    do {                                                                                                               \
        if (fused)                                                                                                     \
            le = launch_coop(enc_fwd_tc_kernel<N_, R_, T_>, dim3(ncta), dim3(kFwdThreads), smem, st, packed, pitch, row_idx, \
                             row0, B, M, V, C, vmax, part, T, mvx, Zk, out_scale);                                     \
        else                                                                                                           \
            le = launch_pdl(enc_fwd_tc_kernel<N_, R_, T_>, dim3(ncta), dim3(kFwdThreads), smem, st, packed, pitch, row_idx, \
                            row0, B, M, V, C, vmax, part, T, mvx, Zk, out_scale);                                      \
    } while (0)
#define NADM_FWD_GO(N_, R_) do { if (tsa) NADM_FWD_GO2(N_, R_, true); else NADM_FWD_GO2(N_, R_, false); } while (0)
    if (raw_mv >= 0) { if (two) NADM_FWD_GO(2, true); else NADM_FWD_GO(1, true); }
    else { if (two) NADM_FWD_GO(2, false); else NADM_FWD_GO(1, false); }
#undef NADM_FWD_GO
#undef NADM_FWD_GO2
    if (le != cudaSuccess) return cuda_fail(le, "enc_fwd_tc_kernel");
    NADM_CHECK_LAUNCH("enc_fwd_tc_kernel");
    if (!fused && defer && raw_mv < 0 && ncta <= 152) {
        // the consumer (nadm_mlp_fwd on this Z) sums the partials of its own rows: no reduction kernel
        DeferredZ& d = deferred_z();
        d.Z = Z; d.part = part; d.vmax = nullptr; d.nparts = ncta; d.B = B;
    } else if (!fused) {
        launch_pdl(enc_fwd_reduce_kernel, dim3(B), dim3(256), 0, st, part, ncta, B, C, (const float*)nullptr, Z, out_scale);
        NADM_CHECK_LAUNCH("enc_fwd_reduce_kernel");
    }
    return NADM_OK;
}

int launch_enc_bwd_tc(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int B, int64_t M,
                      const float* dZ, int C, float* V, float* Vm, float* Vv, const nadm_adam_t* adam, float* dV_out,
                      cudaStream_t st, int raw_mv, int accumulate, const ApplyJob* job) {
    const ApplyJob jb = (job != nullptr) ? *job : ApplyJob{};
    const int T = (int)((M + kSub - 1) / kSub);
    const int ncta = std::min(T, sm_count());
    const int nblk = (B + 127) / 128;
    if (enc_bwd_slab() && enc_bwd_slab_supported(B)) {
        // ---- slab-fed kernel (default): 256-byte runs per row, half-tile stages ----
        const size_t smem_s = (size_t)4 * kHalfTile + (size_t)nblk * 4096 + (size_t)4 * kSlabDepth * kSlabBytes +
                              (size_t)((B + 3) & ~3) * 4 + sizeof(EncBwdSlabSmem) + 64;
        static PerDeviceOnce once_s;
        bool* as = once_s.slot();
        if (as == nullptr || !*as) {
            cudaError_t e = cudaSuccess;
#define NADM_BSLAB_ATTR(N_, M_)                                                                                        \
    if (e == cudaSuccess)                                                                                              \
        e = cudaFuncSetAttribute(enc_bwd_slab_kernel<N_, M_>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)
            NADM_BSLAB_ATTR(1, kModeTrain); NADM_BSLAB_ATTR(2, kModeTrain); NADM_BSLAB_ATTR(1, kModeRaw3);
            NADM_BSLAB_ATTR(2, kModeRaw3); NADM_BSLAB_ATTR(1, kModeRawAny); NADM_BSLAB_ATTR(2, kModeRawAny);
#undef NADM_BSLAB_ATTR
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(enc_bwd_slab)");
            if (as) *as = true;
        }
        const bool two_s = enc_issuers() == 2 && nblk >= 3;
        const uint32_t mvx_s = raw_mv >= 0 ? (3u ^ (uint32_t)(raw_mv & 0xFF)) : 0u;
        const double scale_s = raw_mv >= 0 ? 1.0 : 0.5;
        const int mode = raw_mv < 0 ? kModeTrain : (raw_mv == 3 ? kModeRaw3 : kModeRawAny);
#define NADM_BSLAB_GO(N_, M_)                                                                                          \
    launch_pdl(enc_bwd_slab_kernel<N_, M_>, dim3(ncta), dim3(kBwdSlabThreads), smem_s, st, packed, pitch, row_idx, row0, B, M, \
               dZ, C, V, Vm, Vv, make_adam(adam), dV_out, T, mvx_s, scale_s, accumulate)
        if (mode == kModeTrain) { if (two_s) NADM_BSLAB_GO(2, kModeTrain); else NADM_BSLAB_GO(1, kModeTrain); }
        else if (mode == kModeRaw3) { if (two_s) NADM_BSLAB_GO(2, kModeRaw3); else NADM_BSLAB_GO(1, kModeRaw3); }
        else { if (two_s) NADM_BSLAB_GO(2, kModeRawAny); else NADM_BSLAB_GO(1, kModeRawAny); }
#undef NADM_BSLAB_GO
        NADM_CHECK_LAUNCH("enc_bwd_slab_kernel");
        return NADM_OK;
    }
    const size_t smem = (size_t)kAStages * kATile + (size_t)nblk * 4096 + (size_t)4 * kStDepth * kStTile + (size_t)((B + 3) & ~3) * 4 +
                        sizeof(EncBwdSmem) + 64;
    NADM_REQUIRE(smem <= (size_t)kMaxDynSmem, "batch B=%d too large for encoder_bwd", B);
    static PerDeviceOnce once;
    bool* attr = once.slot();
    if (attr == nullptr || !*attr) {
        cudaError_t e = cudaSuccess;
#define NADM_BWD_ATTR(N_, R_, E_)                                                                                      \
    if (e == cudaSuccess)                                                                                              \
        e = cudaFuncSetAttribute(enc_bwd_tc_kernel<N_, R_, E_>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)
        NADM_BWD_ATTR(1, false, 4); NADM_BWD_ATTR(2, false, 4); NADM_BWD_ATTR(1, true, 4); NADM_BWD_ATTR(2, true, 4);
        NADM_BWD_ATTR(1, false, 8); NADM_BWD_ATTR(2, false, 8); NADM_BWD_ATTR(1, true, 8); NADM_BWD_ATTR(2, true, 8);
#undef NADM_BWD_ATTR
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(enc_bwd_tc)");
        if (attr) *attr = true;
    }
    const bool two = enc_issuers() == 2 && nblk >= 3;       // both issuers then have tiles in every sub-tile
    const uint32_t mvx = raw_mv >= 0 ? (3u ^ (uint32_t)(raw_mv & 0xFF)) : 0u;
    const double scale = raw_mv >= 0 ? 1.0 : 0.5;
    static int epw = 0;                                     // NADM_ENC_BWD_EPW=4|8: epilogue warps (A/B measurements)
    if (epw == 0) {
        const char* e = getenv("NADM_ENC_BWD_EPW");
        epw = (e != nullptr && e[0] == '4') ? 4 : 8;      // default 8: 73 instead of 84 us with Adam (profiles/r2_*)
    }
#define NADM_BWD_GO(N_, R_, E_)                                                                                        \
    launch_pdl(enc_bwd_tc_kernel<N_, R_, E_>, dim3(ncta), dim3(bwd_threads(E_)), smem, st, packed, pitch, row_idx, row0, B, M, \
               dZ, C, V, Vm, Vv, make_adam(adam), dV_out, T, mvx, scale, accumulate, jb)
#define NADM_BWD_GO_E(N_, R_) do { if (epw == 8) NADM_BWD_GO(N_, R_, 8); else NADM_BWD_GO(N_, R_, 4); } while (0)
    if (raw_mv >= 0) { if (two) NADM_BWD_GO_E(2, true); else NADM_BWD_GO_E(1, true); }
    else { if (two) NADM_BWD_GO_E(2, false); else NADM_BWD_GO_E(1, false); }
#undef NADM_BWD_GO_E
#undef NADM_BWD_GO
    NADM_CHECK_LAUNCH("enc_bwd_tc_kernel");
    return NADM_OK;
}

}  // namespace nadm
