// The small replicated network between the two M-wide streams, forward and backward, plus Adam on its parameters.
//
//   forward  (neural_admixture.py:173-176): Zn = RMSNorm_C(Z; eps 1e-8) ; Hh = relu(Zn W1^T + b1) ;
//                                           L_k = Hh W2_k^T + b2_k ; Q_k = softmax(L_k)
//   backward (autograd of the above, triggered at :410) and optimizer.step() for these parameters (:411).
//   Supervised term (:293,:473): sup_weight * CrossEntropyLoss(sum)(Q_0, labels), Q_0 fed as logits.
//
// B x H x (C + sumK) is ~26 MFLOP at the default sizes: these kernels are latency-, not bandwidth-bound; they are
// written for few launches and deterministic reductions (no atomics), not for tensor cores.
#include "nadm_common.cuh"
#include <algorithm>

#include <mutex>
#include <string.h>

namespace nadm {

// ---- error / bookkeeping (shared by all translation units) -------------------------------------------------------
static thread_local char g_err[512] = "";
static int64_t g_launches = 0, g_generic = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return NADM_ECUDA;
}
void count_launch(int n) { __atomic_fetch_add(&g_launches, (int64_t)n, __ATOMIC_RELAXED); }
void count_generic(int n) { __atomic_fetch_add(&g_generic, (int64_t)n, __ATOMIC_RELAXED); }
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

bool pdl_enabled() {   // opt-in (NADM_PDL=1): see nadm_common.cuh
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("NADM_PDL");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

bool gridbar_enabled() {   // opt-in (NADM_GRIDBAR=1): measured SLOWER than the separate reduction kernels, see nadm_common.cuh
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("NADM_GRIDBAR");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

bool use_generic_kernels() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("NADM_GENERIC");
        cached = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return cached == 1;
}

constexpr int kMlpRows = 4;       // batch rows per CTA in the per-row kernels
constexpr int kMlpThreads = 256;
#ifndef NADM_BWD_THREADS
#define NADM_BWD_THREADS 256
#endif
constexpr int kBwdThreads = NADM_BWD_THREADS;   // threads of the backward rows kernel
constexpr int kMaxSumK = NADM_MAX_K * NADM_MAX_HEADS;

struct Heads {
    int n;
    int sumK;
    int k[NADM_MAX_HEADS];
    int off[NADM_MAX_HEADS];
};

// =================================================================================================================
// fused peer exchange (SNP-sharded runs): see nadm_xchg_t in nadm_b200.h
// =================================================================================================================
// LL protocol (the low-latency protocol of NCCL, restated): every float travels as an 8-byte {value, sequence number}
// pair written with ONE 8-byte store, which is atomic — over NVLink too —, so the value needs no flag behind a fence: the
// receiver spins on the pair itself until it carries the number of the exchange it is waiting for.  The earlier form of
// this exchange (values, __threadfence_system(), st.release.sys flag, ld.acquire.sys poll, then read the values)
// cost two fenced NVLink round trips per exchange: 0.1479 ms per step at 8 GPUs against 0.112 ms for a rank's kernels
// alone (profiles/r2_scaling_8gpu_flag_exchange.txt).
// Layout of one rank's exchange area: slots [2 phases][2 parities][NADM_MAX_RANKS][slot_floats] of uint2.  Phase 0 =
// partial projection Z (nadm_mlp_fwd), phase 1 = partial dQ | loss (nadm_mlp_bwd).  A CTA's exchange number `seq`
// counts the exchanges that CTA index has done in that phase (from 1: zeroed memory never matches); slots are
// double-buffered by its parity: a peer can be at most one exchange ahead (it cannot finish exchange seq + 1 without
// this rank's contribution to it, which this rank sends only after it has consumed exchange seq).
struct Xchg {
    int world, rank;
    uint8_t* area[NADM_MAX_RANKS];
    uint32_t* seq;
    long long slot_floats;
};
__host__ __device__ inline size_t xchg_area_bytes(long long slot_floats) {
    return (size_t)2 * 2 * NADM_MAX_RANKS * (size_t)slot_floats * sizeof(uint2);
}
__device__ __forceinline__ uint2* xchg_slot(uint8_t* area, long long slot_floats, int phase, int parity, int src) {
    return reinterpret_cast<uint2*>(area) + ((size_t)(phase * 2 + parity) * NADM_MAX_RANKS + src) * (size_t)slot_floats;
}
__device__ __forceinline__ void st_ll(uint2* p, float v, uint32_t flag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(flag) : "memory");
}
__device__ __forceinline__ float ld_ll_wait(const uint2* p, uint32_t flag) {
    uint32_t v, f, polls = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(p) : "memory");
        if (f == flag) break;
        // a peer that never arrives fails the launch instead of hanging it; the bound (about a minute of polling) leaves
        // room for a peer that is merely late (its first launch of a kernel, a graph instantiation, a slower host)
        if (++polls > (1u << 27)) __trap();
    }
    return __uint_as_float(v);
}
// All threads of the CTA call this.  vals (SHARED memory, n floats): this CTA's values of the rank's partial result on
// entry, the sum over the ranks (in rank order: bit-identical on every rank) on return.  `off`: position of vals[0] in
// the slot.  tmp: shared scratch of world * n floats.  `extra` (CTA 0 only, may be NULL): one more scalar in global
// memory exchanged through slot element extra_off (the partial loss).
__device__ void xchg_allreduce_cta(const Xchg& x, int phase, float* vals, long long off, int n, float* tmp, float* extra,
                                   long long extra_off) {
    const int tid = threadIdx.x, cta = blockIdx.x, nthr = blockDim.x;
    const uint32_t seq = x.seq[phase * NADM_XCHG_MAX_CTAS + cta] + 1u;
    const int parity = (int)(seq & 1u);
    const float mine_extra = (extra != nullptr) ? *extra : 0.f;
    // push: my values into slot [rank] of every peer's area
    for (int idx = tid; idx < x.world * n; idx += nthr) {
        const int r = idx / n, i = idx - r * n;
        if (r != x.rank) st_ll(xchg_slot(x.area[r], x.slot_floats, phase, parity, x.rank) + off + i, vals[i], seq);
    }
    if (extra != nullptr && tid < x.world && tid != x.rank)
        st_ll(xchg_slot(x.area[tid], x.slot_floats, phase, parity, x.rank) + extra_off, mine_extra, seq);
    // pull: every peer's values for the same positions out of MY area, as they arrive
    for (int idx = tid; idx < x.world * n; idx += nthr) {
        const int r = idx / n, i = idx - r * n;
        tmp[idx] = (r == x.rank) ? vals[i]
                                 : ld_ll_wait(xchg_slot(x.area[x.rank], x.slot_floats, phase, parity, r) + off + i, seq);
    }
    if (extra != nullptr && tid == 0) {
        float acc = 0.f;
        for (int r = 0; r < x.world; ++r)
            acc += (r == x.rank) ? mine_extra
                                 : ld_ll_wait(xchg_slot(x.area[x.rank], x.slot_floats, phase, parity, r) + extra_off, seq);
        *extra = acc;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nthr) {
        float acc = 0.f;
        for (int r = 0; r < x.world; ++r) acc += tmp[r * n + i];
        vals[i] = acc;
    }
    if (tid == 0) x.seq[phase * NADM_XCHG_MAX_CTAS + cta] = seq;
    __syncthreads();
}
static int make_xchg(const nadm_xchg_t* in, long long need_floats, int nctas, Xchg* out) {
    out->world = 1;
    if (in == nullptr) return NADM_OK;
    NADM_REQUIRE(in->world >= 1 && in->world <= NADM_MAX_RANKS && in->rank >= 0 && in->rank < in->world,
                 "exchange: world=%d rank=%d unsupported (at most %d ranks)", in->world, in->rank, NADM_MAX_RANKS);
    NADM_REQUIRE(in->seq != nullptr, "exchange: seq is NULL");
    NADM_REQUIRE(need_floats <= in->slot_floats, "exchange: slot of %lld floats too small (%lld needed)",
                 (long long)in->slot_floats, need_floats);
    NADM_REQUIRE(nctas <= NADM_XCHG_MAX_CTAS, "exchange: batch too large (%d CTAs > %d)", nctas, NADM_XCHG_MAX_CTAS);
    out->world = in->world;
    out->rank = in->rank;
    for (int r = 0; r < in->world; ++r) {
        NADM_REQUIRE(in->area[r] != nullptr, "exchange: area[%d] is NULL", r);
        out->area[r] = reinterpret_cast<uint8_t*>(in->area[r]);
    }
    out->seq = in->seq;
    out->slot_floats = in->slot_floats;
    return NADM_OK;
}

// =================================================================================================================
// forward
// =================================================================================================================
// Latency is what this kernel costs (10 us for 26 MFLOP): every phase used to start with a round trip to L2.  The
// weights a thread needs first are therefore requested before anything that waits (the peer exchange, the RMSNorm), the
// RMSNorm runs on 16 lanes per row, and a warp's logit column sits in registers (H = 1024: 32 values per lane).
// deferred reduction of the encoder's per-CTA partials (see DeferredZ): nparts == 0 -> Z is complete in global memory
struct ZParts {
    const long long* part;
    const float* vmax;
    int nparts;
};
constexpr int kZSegs = kMlpThreads / (kMlpRows * 8);   // 8 part segments per (row, component) output
constexpr int kZMaxIter = 19;                          // loads in flight per thread: nparts <= 152

__global__ void __launch_bounds__(kMlpThreads, 2)   // 200 CTAs for B = 800: two per SM, one wave
mlp_fwd_kernel(float* Z, int B, int C, int H, const float* __restrict__ w_rms,
               const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
               const float* __restrict__ b2, Heads hd, float* __restrict__ rinv_out, float* __restrict__ Hh,
               float* __restrict__ Q, ZParts zp, Xchg xc) {
    pdl_prologue();
    extern __shared__ __align__(16) float sm[];
    float* Zn = sm;                              // kMlpRows x MAX_C
    float* Zs = Zn + kMlpRows * NADM_MAX_C;      // kMlpRows x C, compact: this rank's / the summed projection of the rows
    float* xt = Zs + kMlpRows * NADM_MAX_C;      // exchange scratch: MAX_RANKS x kMlpRows x MAX_C (also the segment sums)
    float* Hs = xt + NADM_MAX_RANKS * kMlpRows * NADM_MAX_C;   // kMlpRows x H
    float* Ls = Hs + (size_t)kMlpRows * H;       // kMlpRows x sumK
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * kMlpRows;
    const int nrows = min(kMlpRows, B - b0);
    constexpr int kPre = 4;                  // hidden units per thread held in registers (H <= 1024)
    const bool pre1 = (C == 8) && (H <= kPre * kMlpThreads);
    const bool pre2 = (H == 32 * 32) && (warp < hd.sumK);
    float4 wa[kPre], wb[kPre];
    float bj[kPre], w2r[32];
    if (pre1) {
#pragma unroll
        for (int i = 0; i < kPre; ++i) {
            const int j = tid + i * kMlpThreads;
            if (j < H) {
                wa[i] = reinterpret_cast<const float4*>(W1 + (int64_t)j * 8)[0];
                wb[i] = reinterpret_cast<const float4*>(W1 + (int64_t)j * 8)[1];
                bj[i] = b1[j];
            }
        }
    }
    if (pre2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) w2r[i] = W2[(int64_t)warp * H + lane + 32 * i];
    }
    // ---- this rank's projection of the CTA's rows -> Zs ----
    if (zp.nparts > 0) {
        // sum the encoder's per-CTA partials for these rows here (no reduction kernel, no round trip of Z through L2):
        // thread = (output o = row x 8 + component, segment of the parts), every load of the thread in flight at once.
        // Each partial is an exact integer times its CTA's power-of-two scale, stored as a double; summed in a fixed order.
        const int o = tid / kZSegs, seg = tid % kZSegs, r = o >> 3;
        double acc = 0.0;
        if (r < nrows) {
            const long long* src = zp.part + (int64_t)b0 * 8 + o;
            const int64_t n = (int64_t)B * 8;
            long long v[kZMaxIter];
#pragma unroll
            for (int i = 0; i < kZMaxIter; ++i) {
                const int p = seg + kZSegs * i;
                v[i] = (p < zp.nparts) ? __ldcg(src + (int64_t)p * n) : 0ll;
            }
#pragma unroll
            for (int i = 0; i < kZMaxIter; ++i) acc += __longlong_as_double(v[i]);   // (0 bits = +0.0 past the end)
        }
        double* zred = reinterpret_cast<double*>(xt);               // 256 doubles = 2 KB <= the exchange scratch (2 KB)
        zred[tid] = acc;
        __syncthreads();
        if (tid < kMlpRows * 8) {
            const int rr = tid >> 3, c = tid & 7;
            double t = 0.0;
#pragma unroll
            for (int sg = 0; sg < kZSegs; ++sg) t += zred[tid * kZSegs + sg];
            if (rr < nrows && c < C) Zs[rr * C + c] = (float)(t * 0.5);   // x = code / 2
        }
    } else if (tid < kMlpRows * NADM_MAX_C) {
        const int rr = tid / NADM_MAX_C, c = tid % NADM_MAX_C;
        if (rr < nrows && c < C) Zs[rr * C + c] = Z[(int64_t)(b0 + rr) * C + c];
    }
    __syncthreads();
    if (xc.world > 1)   // sum the ranks' partial projections of this CTA's rows
        xchg_allreduce_cta(xc, 0, Zs, (long long)b0 * C, nrows * C, xt, nullptr, 0);
    if ((zp.nparts > 0 || xc.world > 1) && tid < nrows * C) Z[(int64_t)b0 * C + tid] = Zs[tid];   // (kept for the backward)

    if (tid < kMlpRows * 16) {               // RMSNorm: 16 lanes per row (whole warps: full-mask shuffles)
        const int r = tid >> 4, c = tid & 15, b = b0 + r;
        const bool ok = (b < B) && (c < C);
        const float z = ok ? Zs[r * C + c] : 0.f;
        float ss = z * z;
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float rr = 1.0f / sqrtf(ss / (float)C + 1e-8f);  // torch.nn.RMSNorm(C, eps=1e-8)
        if (c == 0 && b < B) rinv_out[b] = rr;
        if (c < C) Zn[r * NADM_MAX_C + c] = ok ? z * rr * w_rms[c] : 0.f;
    }
    __syncthreads();
    if (pre1) {
#pragma unroll
        for (int i = 0; i < kPre; ++i) {
            const int j = tid + i * kMlpThreads;
            if (j < H) {
                const float w[8] = {wa[i].x, wa[i].y, wa[i].z, wa[i].w, wb[i].x, wb[i].y, wb[i].z, wb[i].w};
                float acc[kMlpRows];
#pragma unroll
                for (int r = 0; r < kMlpRows; ++r) acc[r] = bj[i];
#pragma unroll
                for (int c = 0; c < 8; ++c)
#pragma unroll
                    for (int r = 0; r < kMlpRows; ++r) acc[r] = fmaf(Zn[r * NADM_MAX_C + c], w[c], acc[r]);
#pragma unroll
                for (int r = 0; r < kMlpRows; ++r) {
                    const float h = fmaxf(acc[r], 0.f);
                    Hs[(size_t)r * H + j] = h;
                    if (b0 + r < B) Hh[(int64_t)(b0 + r) * H + j] = h;
                }
            }
        }
    } else {
#pragma unroll 4
        for (int j = tid; j < H; j += blockDim.x) {
            float acc[kMlpRows];
            const float bjj = b1[j];
#pragma unroll
            for (int r = 0; r < kMlpRows; ++r) acc[r] = bjj;
            for (int c = 0; c < C; ++c) {
                const float w = W1[(int64_t)j * C + c];
#pragma unroll
                for (int r = 0; r < kMlpRows; ++r) acc[r] = fmaf(Zn[r * NADM_MAX_C + c], w, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < kMlpRows; ++r) {
                const float h = fmaxf(acc[r], 0.f);
                Hs[(size_t)r * H + j] = h;
                if (b0 + r < B) Hh[(int64_t)(b0 + r) * H + j] = h;
            }
        }
    }
    __syncthreads();
    // logits: one warp per output column kk, all kMlpRows rows at once
    for (int kk = warp; kk < hd.sumK; kk += blockDim.x / 32) {
        float acc[kMlpRows];
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) acc[r] = 0.f;
        if (pre2 && kk == warp) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
#pragma unroll
                for (int r = 0; r < kMlpRows; ++r) acc[r] = fmaf(Hs[(size_t)r * H + lane + 32 * i], w2r[i], acc[r]);
        } else {
#pragma unroll 8
            for (int j = lane; j < H; j += 32) {
                const float w = W2[(int64_t)kk * H + j];
#pragma unroll
                for (int r = 0; r < kMlpRows; ++r) acc[r] = fmaf(Hs[(size_t)r * H + j], w, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) {
            const float s = warp_sum(acc[r]);
            if (lane == 0) Ls[r * hd.sumK + kk] = s + b2[kk];
        }
    }
    __syncthreads();
    // softmax per (row, head)
    for (int i = tid; i < kMlpRows * hd.n; i += blockDim.x) {
        const int r = i / hd.n, h = i % hd.n;
        const int b = b0 + r;
        if (b >= B) continue;
        const float* l = Ls + r * hd.sumK + hd.off[h];
        const int k = hd.k[h];
        float mx = l[0];
        for (int kk = 1; kk < k; ++kk) mx = fmaxf(mx, l[kk]);
        float sum = 0.f;
        for (int kk = 0; kk < k; ++kk) sum += expf(l[kk] - mx);
        const float inv = 1.0f / sum;
        for (int kk = 0; kk < k; ++kk) Q[(int64_t)b * hd.sumK + hd.off[h] + kk] = expf(l[kk] - mx) * inv;
    }
}

// =================================================================================================================
// backward, phase 1: one CTA per kBwdRows batch rows does the whole backward of its rows and writes its PARTIAL
// parameter gradients into its own slab (deterministic: no atomics).
//   dL (softmax backward, + supervised CE on head 0) -> dH = relu'(H) . (dL W2) -> dZn = dH W1 -> RMSNorm backward -> dZ
//   partial gradients of the CTA's rows: dW1^T[c][j] (c = C is db1), dW2[kk][j], db2[kk], dw_rms[c], supervised loss.
// Slab layout (floats): [ (C+1) x H | sumK x H | sumK | C | 1 ], see mlp_slab_floats().
// =================================================================================================================
constexpr int kBwdRows = 8;


// deferred reduction of the decoder's per-CTA dQ partials of ONE head (see DeferredDQ): nparts == 0 -> dQ is complete
struct DQParts {
    const float* part;
    const float* loss_part;
    int nparts, cols_p, k, q_off;
};
constexpr int kDqMaxIter = 19;

template <int CP>   // components padded to 8 or 16
__global__ void __launch_bounds__(kBwdThreads)
mlp_bwd_rows_kernel(float* dQ, const float* __restrict__ Q, const float* __restrict__ Hh,
                    const float* __restrict__ Z, const float* __restrict__ rinv, int B, int C, int H, Heads hd,
                    const int64_t* __restrict__ labels, float sup_weight, const float* __restrict__ w_rms,
                    const float* __restrict__ W1, const float* __restrict__ W2, float* __restrict__ part,
                    float* __restrict__ dZ, float* loss, DQParts dp, Xchg xc) {
    pdl_prologue();
    extern __shared__ __align__(16) float sm[];
    float* dLs = sm;                                         // kBwdRows x sumK
    float* Zn = dLs + (size_t)kBwdRows * hd.sumK;             // kBwdRows x MAX_C   (normalised inputs, recomputed)
    float* red = Zn + kBwdRows * CP;                  // nwarps x kBwdRows x MAX_C  (dZn partials per warp)
    float* sups = red + (kBwdThreads / 32) * kBwdRows * CP;   // kBwdRows
    float* qs = sups + kBwdRows;                              // kBwdRows x sumK : the rows' Q
    float* dqs = qs + (size_t)kBwdRows * hd.sumK;             // kBwdRows x sumK : the rows' dQ (summed over CTAs / ranks)
    float* xt = dqs + (size_t)kBwdRows * hd.sumK;             // scratch: max(256, world x kBwdRows x sumK) floats
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * kBwdRows;
    const int sumK = hd.sumK;
    const int nrows = min(kBwdRows, B - b0);
    float* slab = part + (size_t)blockIdx.x * mlp_slab_floats(C, H, sumK);

    // ---- everything the rows need from global memory is requested in ONE round: Q, dQ (or the decoder's per-CTA
    // partials of it), the normalised inputs.  (The per-(row, head) threads below used to walk Q / dQ with k dependent
    // round trips to L2 each.) ----
    {
        const int nvalid = nrows * sumK;
        for (int i = tid; i < kBwdRows * sumK; i += blockDim.x) {
            qs[i] = (i < nvalid) ? Q[(int64_t)b0 * sumK + i] : 0.f;
            const int c = i % sumK;
            const bool from_parts = dp.nparts > 0 && c >= dp.q_off && c < dp.q_off + dp.k;
            dqs[i] = (i < nvalid && !from_parts) ? dQ[(int64_t)b0 * sumK + i] : 0.f;
        }
        for (int i = tid; i < kBwdRows * CP; i += blockDim.x) {
            const int r = i / CP, c = i % CP, b = b0 + r;
            Zn[i] = (b < B && c < C) ? Z[(int64_t)b * C + c] * rinv[b] * w_rms[c] : 0.f;
        }
    }
    if (dp.nparts > 0) {
        // sum the decoder's per-CTA partials of these rows' dQ here (no reduction kernel): thread = (output o = row x
        // cols_p + column, segment of the parts); the rows' partials are contiguous, so a warp's loads are coalesced.
        const int nout = kBwdRows * dp.cols_p, segs = kBwdThreads / nout;      // 64 x 4 or 128 x 2
        const int o = tid % nout, seg = tid / nout;
        const int64_t n = (int64_t)B * dp.cols_p;
        float acc = 0.f;
        if (o < nrows * dp.cols_p) {
            const float* src = dp.part + (int64_t)b0 * dp.cols_p + o;
            for (int p0 = seg; p0 < dp.nparts; p0 += segs * kDqMaxIter) {
                float v[kDqMaxIter];
#pragma unroll
                for (int i = 0; i < kDqMaxIter; ++i) {
                    const int p = p0 + segs * i;
                    v[i] = (p < dp.nparts) ? __ldcg(src + (int64_t)p * n) : 0.f;
                }
#pragma unroll
                for (int i = 0; i < kDqMaxIter; ++i) acc += v[i];
            }
        }
        __syncthreads();                                                       // (xt is not in use yet; keeps the order simple)
        xt[tid] = acc;
        __syncthreads();
        if (tid < nout) {
            float t = 0.f;
            for (int sg = 0; sg < segs; ++sg) t += xt[sg * nout + tid];
            const int r = tid / dp.cols_p, c = tid % dp.cols_p;
            if (r < nrows && c < dp.k) dqs[r * sumK + dp.q_off + c] = t;
        }
        if (blockIdx.x == 0 && warp == 7 && dp.loss_part != nullptr) {       // the head's loss: partials of all CTAs
            double acc2 = 0.0;
            for (int q = lane; q < dp.nparts; q += 32) acc2 += (double)__ldcg(dp.loss_part + q);
            acc2 = warp_sum_d(acc2);
            if (lane == 0) *loss = (float)((double)*loss + acc2);
        }
    }
    __syncthreads();
    if (xc.world > 1)   // sum the ranks' partial dQ of this CTA's rows; CTA 0 also sums the partial losses
        xchg_allreduce_cta(xc, 1, dqs, (long long)b0 * sumK, nrows * sumK, xt, blockIdx.x == 0 ? loss : nullptr,
                           (long long)B * sumK);
    if (dp.nparts > 0 || xc.world > 1)            // (the complete dQ of these rows, for whoever looks at it afterwards)
        for (int i = tid; i < nrows * sumK; i += blockDim.x) dQ[(int64_t)b0 * sumK + i] = dqs[i];

    // ---- softmax backward per (row, head); the supervised cross-entropy acts on head 0 only ----
    for (int i = tid; i < kBwdRows * hd.n; i += blockDim.x) {
        const int r = i / hd.n, h = i % hd.n;
        const int b = b0 + r;
        const int k = hd.k[h], off = hd.off[h];
        float* out = dLs + r * sumK + off;
        if (h == 0) sups[r] = 0.f;
        if (b >= B) {
            for (int kk = 0; kk < k; ++kk) out[kk] = 0.f;
            continue;
        }
        const float* q = qs + r * sumK + off;
        const float* dq = dqs + r * sumK + off;
        const bool sup = (labels != nullptr) && (h == 0);
        float mx = 0.f, lse = 0.f;
        int y = 0;
        if (sup) {
            y = (int)labels[b];
            mx = q[0];
            for (int kk = 1; kk < k; ++kk) mx = fmaxf(mx, q[kk]);
            float se = 0.f;
            for (int kk = 0; kk < k; ++kk) se += expf(q[kk] - mx);
            lse = logf(se);
            sups[r] = sup_weight * (lse + mx - q[y]);
        }
        float dot = 0.f;
        float g[NADM_MAX_K];
        for (int kk = 0; kk < k; ++kk) {
            float gk = dq[kk];
            if (sup) gk += sup_weight * (expf(q[kk] - mx - lse) - (kk == y ? 1.f : 0.f));
            g[kk] = gk;
            dot = fmaf(gk, q[kk], dot);
        }
        for (int kk = 0; kk < k; ++kk) out[kk] = q[kk] * (g[kk] - dot);
    }
    __syncthreads();

    // ---- hidden units j = tid, tid + 256, ...: dH, partial dW2 / dW1 / db1 of these rows, dZn partial sums ----
    float dzn[kBwdRows][CP];
#pragma unroll
    for (int r = 0; r < kBwdRows; ++r)
#pragma unroll
        for (int c = 0; c < CP; ++c) dzn[r][c] = 0.f;
    const bool c8 = (C == 8);
#pragma unroll 2
    for (int j = tid; j < H; j += blockDim.x) {
        float hh[kBwdRows], dH[kBwdRows];
#pragma unroll
        for (int r = 0; r < kBwdRows; ++r) {
            hh[r] = (b0 + r < B) ? Hh[(int64_t)(b0 + r) * H + j] : 0.f;
            dH[r] = 0.f;
        }
        for (int kk = 0; kk < sumK; ++kk) {
            const float w = W2[(int64_t)kk * H + j];
            float g2 = 0.f;
#pragma unroll
            for (int r = 0; r < kBwdRows; ++r) {
                const float d = dLs[r * sumK + kk];
                dH[r] = fmaf(d, w, dH[r]);
                g2 = fmaf(d, hh[r], g2);
            }
            slab[(size_t)(C + 1) * H + (size_t)kk * H + j] = g2;       // partial dW2[kk][j]
        }
        float gb = 0.f;
#pragma unroll
        for (int r = 0; r < kBwdRows; ++r) {
            dH[r] = (hh[r] > 0.f) ? dH[r] : 0.f;                       // ReLU backward
            gb += dH[r];
        }
        slab[(size_t)C * H + j] = gb;                                   // partial db1[j]
        float w1[CP];
        if (c8 && CP == 8) {
            const float4 a = reinterpret_cast<const float4*>(W1 + (int64_t)j * 8)[0];
            const float4 bq = reinterpret_cast<const float4*>(W1 + (int64_t)j * 8)[1];
            w1[0] = a.x; w1[1] = a.y; w1[2] = a.z; w1[3] = a.w; w1[4] = bq.x; w1[5] = bq.y; w1[6] = bq.z; w1[7] = bq.w;
        } else {
#pragma unroll
            for (int c = 0; c < CP; ++c) w1[c] = (c < C) ? W1[(int64_t)j * C + c] : 0.f;
        }
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            if (c < C) {
                float g1 = 0.f;
#pragma unroll
                for (int r = 0; r < kBwdRows; ++r) {
                    g1 = fmaf(dH[r], Zn[r * CP + c], g1);
                    dzn[r][c] = fmaf(dH[r], w1[c], dzn[r][c]);
                }
                slab[(size_t)c * H + j] = g1;                           // partial dW1[j][c], stored transposed
            }
        }
    }
    // ---- dZn: sum over the 256 threads (warp shuffles, then shared memory across warps) ----
#pragma unroll
    for (int r = 0; r < kBwdRows; ++r) {
        float v[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) v[c] = dzn[r][c];
        const float s = warp_reduce_vec<CP>(v, lane);          // lane c * (32 / CP) holds component c
        if (lane % (32 / CP) == 0) red[(warp * kBwdRows + r) * CP + lane / (32 / CP)] = s;
    }
    __syncthreads();
    // ---- RMSNorm backward (y = z * rinv * w) and the small partial gradients; thread (r, c) ----
    if (tid < kBwdRows * CP) {
        const int r = tid / CP, c = tid % CP, b = b0 + r;
        float d = 0.f;
#pragma unroll
        for (int w = 0; w < kBwdThreads / 32; ++w) d += red[(w * kBwdRows + r) * CP + c];
        const bool ok = (b < B) && (c < C);
        const float z = ok ? Z[(int64_t)b * C + c] : 0.f;
        const float wr = ok ? w_rms[c] : 0.f;
        const float ri = (b < B) ? rinv[b] : 0.f;
        float dot = d * wr * z;                                        // sum over c: the 16 lanes of this row
#pragma unroll
        for (int o = CP / 2; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (ok) dZ[(int64_t)b * C + c] = ri * d * wr - z * (ri * ri * ri * dot / (float)C);
        red[(0 * kBwdRows + r) * CP + c] = ok ? d * z * ri : 0.f;   // per-row dw_rms term (warp 0's slots: own element)
    }
    __syncthreads();
    float* small = slab + (size_t)(C + 1) * H + (size_t)sumK * H;
    for (int o = tid; o < sumK + C + 1; o += blockDim.x) {
        float acc = 0.f;
        if (o < sumK) {
            for (int r = 0; r < kBwdRows; ++r) acc += dLs[r * sumK + o];
        } else if (o < sumK + C) {
            for (int r = 0; r < kBwdRows; ++r) acc += red[r * CP + (o - sumK)];
        } else {
            for (int r = 0; r < kBwdRows; ++r) acc += sups[r];
        }
        small[o] = acc;
    }
}

// =================================================================================================================
// backward, phase 2: sum the slabs in a fixed order (deterministic) and apply Adam to W1, b1, W2, b2, w_rms; the
// supervised loss term is added to *loss.
// =================================================================================================================
constexpr int kApplyParams = 64, kApplyGroups = 8;   // a block sums 64 parameters' slabs in 8 interleaved groups

__global__ void __launch_bounds__(kApplyParams * kApplyGroups)
mlp_bwd_apply_kernel(const float* __restrict__ part, int nslab, int C, int H, int sumK, int has_sup,
                     nadm_mlp_params_t prm, AdamCoef adam_in, float* __restrict__ loss) {
    pdl_prologue();
    const AdamCoef adam = adam_resolve(adam_in);
    __shared__ float red[kApplyGroups][kApplyParams];
    const size_t n = mlp_slab_floats(C, H, sumK);
    const int j = threadIdx.x % kApplyParams, grp = threadIdx.x / kApplyParams;
    const size_t i = (size_t)blockIdx.x * kApplyParams + j;
    // slabs grp, grp + 4, ...: two interleaved chains per thread, fixed association -> deterministic
    float g0 = 0.f, g1 = 0.f;
    if (i < n) {
        int z = grp;
        for (; z + kApplyGroups < nslab; z += 2 * kApplyGroups) {
            g0 += part[(size_t)z * n + i];
            g1 += part[(size_t)(z + kApplyGroups) * n + i];
        }
        if (z < nslab) g0 += part[(size_t)z * n + i];
    }
    red[grp][j] = g0 + g1;
    __syncthreads();
    if (grp != 0 || i >= n) return;
    float g = 0.f;
#pragma unroll
    for (int q = 0; q < kApplyGroups; q += 2) g += red[q][j] + red[q + 1][j];      // fixed order
    ApplyJob jb{part, nslab, C, H, sumK, has_sup, prm, adam, loss};
    mlp_apply_param(i, g, jb, adam);
}

DeferredApply& deferred_apply() { static thread_local DeferredApply d{}; return d; }

static int launch_apply(const ApplyJob& jb, cudaStream_t st) {
    const size_t nparam = mlp_slab_floats(jb.C, jb.H, jb.sumK);
    launch_pdl(mlp_bwd_apply_kernel, dim3((unsigned)((nparam + kApplyParams - 1) / kApplyParams)), dim3(kApplyParams * kApplyGroups),
               0, st, jb.part, jb.nslab, jb.C, jb.H, jb.sumK, jb.has_sup, jb.prm, jb.adam, jb.loss);
    NADM_CHECK_LAUNCH("mlp_bwd_apply_kernel");
    return NADM_OK;
}
// a pending update (nadm_mlp_bwd_deferred not followed by the nadm_encoder_bwd that would have run it) as its own kernel
int flush_deferred_apply(cudaStream_t st) {
    DeferredApply& d = deferred_apply();
    if (d.dZ == nullptr) return NADM_OK;
    const ApplyJob jb = d.job;
    d.dZ = nullptr;
    return launch_apply(jb, st);
}

size_t mlp_bwd_workspace_bytes(int B, int C, int H, int sumK) {
    return (size_t)((B + kBwdRows - 1) / kBwdRows) * mlp_slab_floats(C, H, sumK) * sizeof(float);
}

}  // namespace nadm

using namespace nadm;

static int make_heads(const int32_t* ks, int nheads, Heads* hd) {
    NADM_REQUIRE(ks != nullptr && nheads >= 1 && nheads <= NADM_MAX_HEADS, "nheads=%d unsupported (1..%d)", nheads, NADM_MAX_HEADS);
    hd->n = nheads;
    int off = 0;
    for (int i = 0; i < nheads; ++i) {
        NADM_REQUIRE(ks[i] >= 1 && ks[i] <= NADM_MAX_K, "k=%d unsupported (1..%d)", ks[i], NADM_MAX_K);
        hd->k[i] = ks[i];
        hd->off[i] = off;
        off += ks[i];
    }
    hd->sumK = off;
    return NADM_OK;
}

extern "C" int nadm_version(void) { return 100; }
extern "C" const char* nadm_last_error(void) { return g_err; }
extern "C" int64_t nadm_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" int64_t nadm_generic_launch_count(void) { return __atomic_load_n(&g_generic, __ATOMIC_RELAXED); }

extern "C" size_t nadm_xchg_area_bytes(int64_t slot_floats) { return nadm::xchg_area_bytes(slot_floats); }

extern "C" int nadm_ipc_alloc(size_t bytes, void** dev_ptr, void* handle_out) {
    NADM_REQUIRE(dev_ptr != nullptr && handle_out != nullptr && bytes > 0, "nadm_ipc_alloc: bad argument");
    cudaError_t e = cudaMalloc(dev_ptr, bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(exchange area)");
    e = cudaMemset(*dev_ptr, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *dev_ptr);
    if (e != cudaSuccess) { cudaFree(*dev_ptr); *dev_ptr = nullptr; return cuda_fail(e, "cudaIpcGetMemHandle"); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out, &h, sizeof(h));
    return NADM_OK;
}
extern "C" int nadm_ipc_open(const void* handle, void** dev_ptr) {
    NADM_REQUIRE(handle != nullptr && dev_ptr != nullptr, "nadm_ipc_open: NULL pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
    return NADM_OK;
}
extern "C" int nadm_ipc_close(void* dev_ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcCloseMemHandle");
    return NADM_OK;
}
extern "C" int nadm_ipc_free(void* dev_ptr) {
    cudaError_t e = cudaFree(dev_ptr);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFree(exchange area)");
    return NADM_OK;
}

extern "C" int nadm_mlp_fwd(float* Z, int32_t B, int32_t C, int32_t H, const float* w_rms, const float* W1,
                            const float* b1, const float* W2, const float* b2, const int32_t* ks, int32_t nheads,
                            float* rinv, float* Hh, float* Q, const nadm_xchg_t* xchg, void* stream) {
    if (int rc = flush_deferred_apply((cudaStream_t)stream)) return rc;   // (this kernel reads the network's parameters)
    Heads hd;
    if (int rc = make_heads(ks, nheads, &hd)) return rc;
    Xchg xc;
    if (int rc = make_xchg(xchg, (long long)B * C, (B + kMlpRows - 1) / kMlpRows, &xc)) return rc;
    NADM_REQUIRE(B > 0 && H > 0, "empty batch or hidden layer");
    NADM_REQUIRE(C >= 1 && C <= NADM_MAX_C, "n_components C=%d unsupported (1..%d)", C, NADM_MAX_C);
    NADM_REQUIRE(Z && w_rms && W1 && b1 && W2 && b2 && rinv && Hh && Q, "NULL pointer");
    const size_t smem = ((size_t)(2 + NADM_MAX_RANKS) * kMlpRows * NADM_MAX_C + (size_t)kMlpRows * H +
                         (size_t)kMlpRows * hd.sumK) * sizeof(float);
    NADM_REQUIRE(smem <= 200 * 1024, "hidden_size H=%d too large", H);
    // a reduction that nadm_encoder_fwd_deferred left pending for this Z is completed by the kernel itself
    ZParts zp{nullptr, nullptr, 0};
    {
        DeferredZ& d = deferred_z();
        if (d.Z == Z && d.Z != nullptr) {
            NADM_REQUIRE(d.B == B && C <= 8, "deferred projection: batch %d / C=%d do not match the pending reduction (batch %d)",
                         B, C, d.B);
            zp.part = d.part; zp.vmax = d.vmax; zp.nparts = d.nparts;
            d.Z = nullptr;
        }
    }
    static PerDeviceOnce once;
    bool* attr = once.slot();
    if (attr == nullptr || !*attr) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mlp_fwd)");
        if (attr) *attr = true;
    }
    launch_pdl(mlp_fwd_kernel, dim3((B + kMlpRows - 1) / kMlpRows), dim3(kMlpThreads), smem, (cudaStream_t)stream, Z, B, C, H,
               w_rms, W1, b1, W2, b2, hd, rinv, Hh, Q, zp, xc);
    NADM_CHECK_LAUNCH("mlp_fwd_kernel");
    return NADM_OK;
}

static int mlp_bwd_impl(float* dQ, const float* Q, const float* Hh, const float* Z, const float* rinv,
                        int32_t B, int32_t C, int32_t H, const int32_t* ks, int32_t nheads, const int64_t* labels,
                        float sup_weight, const nadm_mlp_params_t* params, const nadm_adam_t* adam, float* dZ,
                        float* loss, void* ws, size_t ws_bytes, const nadm_xchg_t* xchg, void* stream, bool defer_apply) {
    if (int rc = flush_deferred_apply((cudaStream_t)stream)) return rc;
    Heads hd;
    if (int rc = make_heads(ks, nheads, &hd)) return rc;
    Xchg xc;
    if (int rc = make_xchg(xchg, (long long)B * hd.sumK + 1, (B + kBwdRows - 1) / kBwdRows, &xc)) return rc;
    NADM_REQUIRE(B > 0 && H > 0, "empty batch or hidden layer");
    NADM_REQUIRE(C >= 1 && C <= NADM_MAX_C, "n_components C=%d unsupported (1..%d)", C, NADM_MAX_C);
    NADM_REQUIRE(dQ && Q && Hh && Z && rinv && params && dZ && loss && ws, "NULL pointer");
    const nadm_mlp_params_t& p = *params;
    NADM_REQUIRE(p.w_rms && p.W1 && p.b1 && p.W2 && p.b2, "NULL parameter pointer");
    NADM_REQUIRE(adam == nullptr || (p.m_w_rms && p.m_W1 && p.m_b1 && p.m_W2 && p.m_b2 && p.v_w_rms && p.v_W1 &&
                                     p.v_b1 && p.v_W2 && p.v_b2), "NULL Adam moment pointer");
    // workspace: one slab of partial parameter gradients per CTA of kBwdRows rows
    const int nslab = (B + kBwdRows - 1) / kBwdRows;
    const size_t need = mlp_bwd_workspace_bytes(B, C, H, hd.sumK);
    NADM_REQUIRE(need <= ws_bytes, "workspace too small for mlp_bwd (%zu > %zu)", need, ws_bytes);
    float* gpart = (float*)ws;
    cudaStream_t st = (cudaStream_t)stream;
    // a reduction that nadm_decoder_step_deferred left pending for this dQ is completed by the kernel itself; its
    // partials live in the workspace, so the slabs go behind them (or, if they do not fit, the reduction runs now)
    DQParts dp{nullptr, nullptr, 0, 8, 0, 0};
    {
        DeferredDQ& d = deferred_dq();
        if (d.dQ == dQ && d.dQ != nullptr) {
            const DeferredDQ r = d;
            d.dQ = nullptr;
            uint8_t* behind = (uint8_t*)r.part + ((r.bytes + 255) & ~(size_t)255);
            const bool fits = r.B == B && r.q_ld == hd.sumK && (uint8_t*)r.part >= (uint8_t*)ws &&
                              behind + need <= (uint8_t*)ws + ws_bytes;
            if (fits) {
                gpart = (float*)behind;
                dp.part = r.part; dp.loss_part = r.loss_part; dp.nparts = r.nparts; dp.cols_p = r.cols_p; dp.k = r.k;
                dp.q_off = r.q_off;
                NADM_REQUIRE(r.loss_part == nullptr || r.loss == loss, "deferred dQ: the loss accumulator changed");
            } else if (int rc = launch_reduce_parts(r.part, r.nparts, r.B, r.cols_p, r.k, dQ, r.q_ld, r.q_off, 1.0f, r.loss_part,
                                                    r.loss, st)) {
                return rc;
            }
        }
    }
    const int CP = C <= 8 ? 8 : 16;
    const size_t xt_floats = std::max<size_t>(kBwdThreads, xc.world > 1 ? (size_t)xc.world * kBwdRows * hd.sumK : 0);
    const size_t smem = ((size_t)3 * kBwdRows * hd.sumK + (size_t)kBwdRows * CP +
                         (size_t)(kBwdThreads / 32) * kBwdRows * CP + kBwdRows + xt_floats) * sizeof(float);
    if (smem > 48 * 1024) {                          // very wide head sets only (sumK > ~480)
        static PerDeviceOnce once_b;
        bool* ab = once_b.slot();
        if (ab == nullptr || !*ab) {
            cudaError_t e = cudaFuncSetAttribute(mlp_bwd_rows_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(mlp_bwd_rows_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mlp_bwd_rows)");
            if (ab) *ab = true;
        }
    }
    if (CP == 8)
        launch_pdl(mlp_bwd_rows_kernel<8>, dim3(nslab), dim3(kBwdThreads), smem, st, dQ, Q, Hh, Z, rinv, B, C, H, hd, labels,
                   sup_weight, p.w_rms, p.W1, p.W2, gpart, dZ, loss, dp, xc);
    else
        launch_pdl(mlp_bwd_rows_kernel<16>, dim3(nslab), dim3(kBwdThreads), smem, st, dQ, Q, Hh, Z, rinv, B, C, H, hd, labels,
                   sup_weight, p.w_rms, p.W1, p.W2, gpart, dZ, loss, dp, xc);
    NADM_CHECK_LAUNCH("mlp_bwd_rows_kernel");
    const ApplyJob jb{gpart, nslab, C, H, hd.sumK, (int)(labels != nullptr), p, make_adam(adam), loss};
    if (defer_apply) {       // the nadm_encoder_bwd that follows on this dZ runs the update on its epilogue warps
        DeferredApply& d = deferred_apply();
        d.dZ = dZ;
        d.job = jb;
        return NADM_OK;
    }
    return launch_apply(jb, st);
}
extern "C" int nadm_mlp_bwd(float* dQ, const float* Q, const float* Hh, const float* Z, const float* rinv,
                            int32_t B, int32_t C, int32_t H, const int32_t* ks, int32_t nheads, const int64_t* labels,
                            float sup_weight, const nadm_mlp_params_t* params, const nadm_adam_t* adam, float* dZ,
                            float* loss, void* ws, size_t ws_bytes, const nadm_xchg_t* xchg, void* stream) {
    return mlp_bwd_impl(dQ, Q, Hh, Z, rinv, B, C, H, ks, nheads, labels, sup_weight, params, adam, dZ, loss, ws, ws_bytes,
                        xchg, stream, false);
}
extern "C" int nadm_mlp_bwd_deferred(float* dQ, const float* Q, const float* Hh, const float* Z, const float* rinv,
                                     int32_t B, int32_t C, int32_t H, const int32_t* ks, int32_t nheads,
                                     const int64_t* labels, float sup_weight, const nadm_mlp_params_t* params,
                                     const nadm_adam_t* adam, float* dZ, float* loss, void* ws, size_t ws_bytes,
                                     const nadm_xchg_t* xchg, void* stream) {
    return mlp_bwd_impl(dQ, Q, Hh, Z, rinv, B, C, H, ks, nheads, labels, sup_weight, params, adam, dZ, loss, ws, ws_bytes,
                        xchg, stream, true);
}
