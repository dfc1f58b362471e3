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

#include <mutex>

namespace nadm {

// ---- error / bookkeeping (shared by all translation units) -------------------------------------------------------
static thread_local char g_err[512] = "";
static int64_t g_launches = 0;

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
constexpr int kMaxSumK = NADM_MAX_K * NADM_MAX_HEADS;

struct Heads {
    int n;
    int sumK;
    int k[NADM_MAX_HEADS];
    int off[NADM_MAX_HEADS];
};

// =================================================================================================================
// forward
// =================================================================================================================
__global__ void __launch_bounds__(kMlpThreads)
mlp_fwd_kernel(const float* __restrict__ Z, int B, int C, int H, const float* __restrict__ w_rms,
               const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
               const float* __restrict__ b2, Heads hd, float* __restrict__ rinv_out, float* __restrict__ Hh,
               float* __restrict__ Q) {
    extern __shared__ __align__(16) float sm[];
    float* Zn = sm;                         // kMlpRows x C
    float* Hs = Zn + kMlpRows * NADM_MAX_C;  // kMlpRows x H
    float* Ls = Hs + (size_t)kMlpRows * H;   // kMlpRows x sumK
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * kMlpRows;

    if (tid < kMlpRows) {
        const int b = b0 + tid;
        if (b < B) {
            float ss = 0.f;
            for (int c = 0; c < C; ++c) { float z = Z[(int64_t)b * C + c]; ss = fmaf(z, z, ss); }
            const float r = 1.0f / sqrtf(ss / (float)C + 1e-8f);  // torch.nn.RMSNorm(C, eps=1e-8)
            rinv_out[b] = r;
            for (int c = 0; c < C; ++c) Zn[tid * NADM_MAX_C + c] = Z[(int64_t)b * C + c] * r * w_rms[c];
        } else {
            for (int c = 0; c < C; ++c) Zn[tid * NADM_MAX_C + c] = 0.f;
        }
    }
    __syncthreads();
    for (int j = tid; j < H; j += blockDim.x) {
        float acc[kMlpRows];
        const float bj = b1[j];
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) acc[r] = bj;
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
    __syncthreads();
    // logits: one warp per output column kk, all kMlpRows rows at once
    for (int kk = warp; kk < hd.sumK; kk += blockDim.x / 32) {
        float acc[kMlpRows];
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) acc[r] = 0.f;
        for (int j = lane; j < H; j += 32) {
            const float w = W2[(int64_t)kk * H + j];
#pragma unroll
            for (int r = 0; r < kMlpRows; ++r) acc[r] = fmaf(Hs[(size_t)r * H + j], w, acc[r]);
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
// backward A: per-row quantities  dL, dHpre, dZ, per-row dw_rms terms, supervised term
// =================================================================================================================
__global__ void __launch_bounds__(kMlpThreads)
mlp_bwd_rows_kernel(const float* __restrict__ dQ, const float* __restrict__ Q, const float* __restrict__ Hh,
                    const float* __restrict__ Z, const float* __restrict__ rinv, int B, int C, int H, Heads hd,
                    const int64_t* __restrict__ labels, float sup_weight, const float* __restrict__ w_rms,
                    const float* __restrict__ W1, const float* __restrict__ W2, float* __restrict__ dL,
                    float* __restrict__ dHpre, float* __restrict__ dwr, float* __restrict__ suploss,
                    float* __restrict__ dZ) {
    extern __shared__ __align__(16) float sm[];
    float* dLs = sm;                                   // kMlpRows x sumK
    float* dHs = dLs + (size_t)kMlpRows * hd.sumK;      // kMlpRows x H
    float* dZn = dHs + (size_t)kMlpRows * H;            // kMlpRows x MAX_C
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * kMlpRows;

    // softmax backward per (row, head); the supervised cross-entropy acts on head 0 only
    for (int i = tid; i < kMlpRows * hd.n; i += blockDim.x) {
        const int r = i / hd.n, h = i % hd.n;
        const int b = b0 + r;
        const int k = hd.k[h], off = hd.off[h];
        float* out = dLs + r * hd.sumK + off;
        if (b >= B) {
            for (int kk = 0; kk < k; ++kk) out[kk] = 0.f;
            continue;
        }
        const float* q = Q + (int64_t)b * hd.sumK + off;
        const float* dq = dQ + (int64_t)b * hd.sumK + off;
        float sl = 0.f;
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
            sl = sup_weight * (lse + mx - q[y]);
            suploss[b] = sl;
        }
        float dot = 0.f;
        float g[NADM_MAX_K];
        for (int kk = 0; kk < k; ++kk) {
            float gk = dq[kk];
            if (sup) gk += sup_weight * (expf(q[kk] - mx - lse) - (kk == y ? 1.f : 0.f));
            g[kk] = gk;
            dot = fmaf(gk, q[kk], dot);
        }
        for (int kk = 0; kk < k; ++kk) {
            const float v = q[kk] * (g[kk] - dot);
            out[kk] = v;
            dL[(int64_t)b * hd.sumK + off + kk] = v;
        }
    }
    __syncthreads();
    // dHpre[r][j] = relu'(H) * sum_kk dL[r][kk] W2[kk][j]
    for (int j = tid; j < H; j += blockDim.x) {
        float acc[kMlpRows];
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) acc[r] = 0.f;
        for (int kk = 0; kk < hd.sumK; ++kk) {
            const float w = W2[(int64_t)kk * H + j];
#pragma unroll
            for (int r = 0; r < kMlpRows; ++r) acc[r] = fmaf(dLs[r * hd.sumK + kk], w, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) {
            const int b = b0 + r;
            float v = 0.f;
            if (b < B) {
                v = (Hh[(int64_t)b * H + j] > 0.f) ? acc[r] : 0.f;
                dHpre[(int64_t)b * H + j] = v;
            }
            dHs[(size_t)r * H + j] = v;
        }
    }
    __syncthreads();
    // dZn[r][c] = sum_j dHpre[r][j] W1[j][c] : one warp per (r, c)
    for (int o = warp; o < kMlpRows * C; o += blockDim.x / 32) {
        const int r = o / C, c = o % C;
        float acc = 0.f;
        for (int j = lane; j < H; j += 32) acc = fmaf(dHs[(size_t)r * H + j], W1[(int64_t)j * C + c], acc);
        acc = warp_sum(acc);
        if (lane == 0) dZn[r * NADM_MAX_C + c] = acc;
    }
    __syncthreads();
    // RMSNorm backward: y = z * rinv * w
    if (tid < kMlpRows) {
        const int b = b0 + tid;
        if (b < B) {
            const float r = rinv[b];
            float dot = 0.f;
            for (int c = 0; c < C; ++c) dot = fmaf(dZn[tid * NADM_MAX_C + c] * w_rms[c], Z[(int64_t)b * C + c], dot);
            const float coef = r * r * r * dot / (float)C;
            for (int c = 0; c < C; ++c) {
                const float z = Z[(int64_t)b * C + c];
                const float dzn = dZn[tid * NADM_MAX_C + c];
                dZ[(int64_t)b * C + c] = r * dzn * w_rms[c] - z * coef;
                dwr[(int64_t)b * C + c] = dzn * z * r;
            }
        }
    }
}

// =================================================================================================================
// backward B: reductions over the batch -> parameter gradients, then Adam.  grid = (ceil(H/32), 1 + nchunks)
//   blockIdx.y == 0        : dW1[j][:], db1[j]                (+ block (0,0) also does db2, dw_rms, loss terms)
//   blockIdx.y == 1 + q    : dW2[8q .. 8q+8)[j]
// block = 32 hidden units x 8 batch segments; partial sums over segments are combined through shared memory.
// =================================================================================================================
constexpr int kSeg = 8;

__device__ __forceinline__ void adam_store(float* p, float* m, float* v, float* gout, int64_t i, float g,
                                           const AdamCoef& c) {
    if (gout != nullptr) gout[i] = g;
    if (c.enabled) {
        float mm = m[i], vv = v[i];
        p[i] = adam_apply(p[i], g, mm, vv, c);
        m[i] = mm;
        v[i] = vv;
    }
}

constexpr int kRowChunks = 8;   // grid.z: the batch is split into row chunks whose partial gradients are summed in phase 2

// Phase 1: partial parameter gradients of one row chunk.  grid = (ceil(H/32), 1 + ceil(sumK/8), kRowChunks).
//   blockIdx.y == 0     : gW1[z][j][0..C) and gb1 (stored as column C) for 32 hidden units j
//   blockIdx.y == 1 + q : gW2[z][8q .. 8q+8)[j]
// Layout of `part`: [kRowChunks] x ( H x (C+1)  |  sumK x H ).
__global__ void __launch_bounds__(32 * kSeg)
mlp_bwd_params_kernel(const float* __restrict__ dL, const float* __restrict__ dHpre, const float* __restrict__ Hh,
                      const float* __restrict__ Z, const float* __restrict__ rinv, const float* __restrict__ w_rms,
                      int B, int C, int H, int sumK, float* __restrict__ part) {
    __shared__ float red[kSeg][32][NADM_MAX_C + 1];
    const int jl = threadIdx.x & 31, seg = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + jl;
    const bool jok = j < H;
    const int rows_per_chunk = (B + kRowChunks - 1) / kRowChunks;
    const int c0 = blockIdx.z * rows_per_chunk, c1 = min(B, c0 + rows_per_chunk);
    float* mypart = part + (size_t)blockIdx.z * ((size_t)H * (C + 1) + (size_t)sumK * H);

    if (blockIdx.y == 0) {
        float acc[NADM_MAX_C + 1];
#pragma unroll
        for (int c = 0; c <= NADM_MAX_C; ++c) acc[c] = 0.f;
        if (jok) {
            for (int b = c0 + seg; b < c1; b += kSeg) {
                const float d = dHpre[(int64_t)b * H + j];
                if (d == 0.f) continue;
                const float r = rinv[b];
                // Zn is recomputed: Zn[b][c] = Z[b][c] * rinv[b] * w_rms[c]  (w_rms is updated by a later kernel)
#pragma unroll
                for (int c = 0; c < NADM_MAX_C; ++c)
                    if (c < C) acc[c] = fmaf(d, Z[(int64_t)b * C + c] * r * w_rms[c], acc[c]);
                acc[NADM_MAX_C] += d;
            }
        }
#pragma unroll
        for (int c = 0; c <= NADM_MAX_C; ++c) red[seg][jl][c] = acc[c];
        __syncthreads();
        for (int o = threadIdx.x; o < 32 * (C + 1); o += blockDim.x) {
            const int jj = o / (C + 1), c = o % (C + 1);
            const int jg = blockIdx.x * 32 + jj;
            if (jg >= H) continue;
            const int cc = (c < C) ? c : NADM_MAX_C;
            float g = 0.f;
#pragma unroll
            for (int s = 0; s < kSeg; ++s) g += red[s][jj][cc];
            mypart[(size_t)jg * (C + 1) + c] = g;
        }
    } else {
        const int k0 = (blockIdx.y - 1) * 8;
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
        if (jok) {
            for (int b = c0 + seg; b < c1; b += kSeg) {
                const float h = Hh[(int64_t)b * H + j];
                if (h == 0.f) continue;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (k0 + q < sumK) acc[q] = fmaf(dL[(int64_t)b * sumK + k0 + q], h, acc[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) red[seg][jl][q] = acc[q];
        __syncthreads();
        for (int o = threadIdx.x; o < 32 * 8; o += blockDim.x) {
            const int q = o / 32, jj = o % 32;
            const int jg = blockIdx.x * 32 + jj;
            if (jg >= H || k0 + q >= sumK) continue;
            float g = 0.f;
#pragma unroll
            for (int s = 0; s < kSeg; ++s) g += red[s][jj][q];
            mypart[(size_t)H * (C + 1) + (size_t)(k0 + q) * H + jg] = g;
        }
    }
}

// Phase 2: sum the row-chunk partials (fixed order: deterministic) and apply Adam to W1, b1, W2.
__global__ void __launch_bounds__(256)
mlp_bwd_apply_kernel(const float* __restrict__ part, int C, int H, int sumK, nadm_mlp_params_t prm, AdamCoef adam) {
    const size_t n1 = (size_t)H * (C + 1), n = n1 + (size_t)sumK * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = 0.f;
#pragma unroll
    for (int z = 0; z < kRowChunks; ++z) g += part[(size_t)z * n + i];
    if (i < n1) {
        const int jg = (int)(i / (C + 1)), c = (int)(i % (C + 1));
        if (c < C) adam_store(prm.W1, prm.m_W1, prm.v_W1, prm.g_W1, (int64_t)jg * C + c, g, adam);
        else adam_store(prm.b1, prm.m_b1, prm.v_b1, prm.g_b1, jg, g, adam);
    } else {
        adam_store(prm.W2, prm.m_W2, prm.v_W2, prm.g_W2, (int64_t)(i - n1), g, adam);
    }
}

// db2[kk] = sum_b dL[b][kk]; dw_rms[c] = sum_b dwr[b][c]; loss += sum_b suploss[b].  One warp per output.
// Runs AFTER mlp_bwd_params_kernel (which still reads the old w_rms).
__global__ void __launch_bounds__(256)
mlp_bwd_small_kernel(const float* __restrict__ dL, const float* __restrict__ dwr, const float* __restrict__ suploss,
                     int has_sup, int B, int C, int sumK, nadm_mlp_params_t prm, AdamCoef adam,
                     float* __restrict__ loss) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x / 32;
    const int nout = sumK + C + (has_sup ? 1 : 0);
    for (int o = warp; o < nout; o += nwarps) {
        float acc = 0.f;
        if (o < sumK) {
            for (int b = lane; b < B; b += 32) acc += dL[(int64_t)b * sumK + o];
        } else if (o < sumK + C) {
            for (int b = lane; b < B; b += 32) acc += dwr[(int64_t)b * C + (o - sumK)];
        } else {
            for (int b = lane; b < B; b += 32) acc += suploss[b];
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            if (o < sumK) adam_store(prm.b2, prm.m_b2, prm.v_b2, prm.g_b2, o, acc, adam);
            else if (o < sumK + C) adam_store(prm.w_rms, prm.m_w_rms, prm.v_w_rms, prm.g_w_rms, o - sumK, acc, adam);
            else *loss += acc;
        }
    }
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

extern "C" int nadm_mlp_fwd(const float* Z, int32_t B, int32_t C, int32_t H, const float* w_rms, const float* W1,
                            const float* b1, const float* W2, const float* b2, const int32_t* ks, int32_t nheads,
                            float* rinv, float* Hh, float* Q, void* stream) {
    Heads hd;
    if (int rc = make_heads(ks, nheads, &hd)) return rc;
    NADM_REQUIRE(B > 0 && H > 0, "empty batch or hidden layer");
    NADM_REQUIRE(C >= 1 && C <= NADM_MAX_C, "n_components C=%d unsupported (1..%d)", C, NADM_MAX_C);
    NADM_REQUIRE(Z && w_rms && W1 && b1 && W2 && b2 && rinv && Hh && Q, "NULL pointer");
    const size_t smem = ((size_t)kMlpRows * NADM_MAX_C + (size_t)kMlpRows * H + (size_t)kMlpRows * hd.sumK) * sizeof(float);
    NADM_REQUIRE(smem <= 200 * 1024, "hidden_size H=%d too large", H);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mlp_fwd)");
        attr = true;
    }
    mlp_fwd_kernel<<<(B + kMlpRows - 1) / kMlpRows, kMlpThreads, smem, (cudaStream_t)stream>>>(
        Z, B, C, H, w_rms, W1, b1, W2, b2, hd, rinv, Hh, Q);
    NADM_CHECK_LAUNCH("mlp_fwd_kernel");
    return NADM_OK;
}

extern "C" int nadm_mlp_bwd(const float* dQ, const float* Q, const float* Hh, const float* Z, const float* rinv,
                            int32_t B, int32_t C, int32_t H, const int32_t* ks, int32_t nheads, const int64_t* labels,
                            float sup_weight, const nadm_mlp_params_t* params, const nadm_adam_t* adam, float* dZ,
                            float* loss, void* ws, size_t ws_bytes, void* stream) {
    Heads hd;
    if (int rc = make_heads(ks, nheads, &hd)) return rc;
    NADM_REQUIRE(B > 0 && H > 0, "empty batch or hidden layer");
    NADM_REQUIRE(C >= 1 && C <= NADM_MAX_C, "n_components C=%d unsupported (1..%d)", C, NADM_MAX_C);
    NADM_REQUIRE(dQ && Q && Hh && Z && rinv && params && dZ && loss && ws, "NULL pointer");
    const nadm_mlp_params_t& p = *params;
    NADM_REQUIRE(p.w_rms && p.W1 && p.b1 && p.W2 && p.b2, "NULL parameter pointer");
    NADM_REQUIRE(adam == nullptr || (p.m_w_rms && p.m_W1 && p.m_b1 && p.m_W2 && p.m_b2 && p.v_w_rms && p.v_W1 &&
                                     p.v_b1 && p.v_W2 && p.v_b2), "NULL Adam moment pointer");
    // workspace carve-up: dL (B x sumK) | dHpre (B x H) | dwr (B x C) | suploss (B)
    const size_t need = ((size_t)B * ((size_t)hd.sumK + H + C + 1) +
                         (size_t)kRowChunks * ((size_t)H * (C + 1) + (size_t)hd.sumK * H)) * sizeof(float);
    NADM_REQUIRE(need <= ws_bytes, "workspace too small for mlp_bwd (%zu > %zu)", need, ws_bytes);
    float* dL = (float*)ws;
    float* dHpre = dL + (size_t)B * hd.sumK;
    float* dwr = dHpre + (size_t)B * H;
    float* suploss = dwr + (size_t)B * C;
    float* gpart = suploss + B;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = ((size_t)kMlpRows * hd.sumK + (size_t)kMlpRows * H + (size_t)kMlpRows * NADM_MAX_C) * sizeof(float);
    NADM_REQUIRE(smem <= 200 * 1024, "hidden_size H=%d too large", H);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(mlp_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mlp_bwd_rows)");
        attr = true;
    }
    mlp_bwd_rows_kernel<<<(B + kMlpRows - 1) / kMlpRows, kMlpThreads, smem, st>>>(
        dQ, Q, Hh, Z, rinv, B, C, H, hd, labels, sup_weight, p.w_rms, p.W1, p.W2, dL, dHpre, dwr, suploss, dZ);
    NADM_CHECK_LAUNCH("mlp_bwd_rows_kernel");
    const AdamCoef ac = make_adam(adam);
    dim3 grid((H + 31) / 32, 1 + (hd.sumK + 7) / 8, kRowChunks);
    mlp_bwd_params_kernel<<<grid, 32 * kSeg, 0, st>>>(dL, dHpre, Hh, Z, rinv, p.w_rms, B, C, H, hd.sumK, gpart);
    NADM_CHECK_LAUNCH("mlp_bwd_params_kernel");
    const size_t nparam = (size_t)H * (C + 1) + (size_t)hd.sumK * H;
    mlp_bwd_apply_kernel<<<(unsigned)((nparam + 255) / 256), 256, 0, st>>>(gpart, C, H, hd.sumK, p, ac);
    NADM_CHECK_LAUNCH("mlp_bwd_apply_kernel");
    mlp_bwd_small_kernel<<<1, 256, 0, st>>>(dL, dwr, suploss, labels != nullptr, B, C, hd.sumK, p, ac, loss);
    NADM_CHECK_LAUNCH("mlp_bwd_small_kernel");
    return NADM_OK;
}
