// Shared device/host helpers for libnadm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <math.h>
#include <algorithm>

#include "../../include/nadm_b200.h"

namespace nadm {

// ---- error plumbing -------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
void count_generic(int n = 1);   // launches of the first-generation CUDA-core kernels (shapes outside the tensor-core path)
int sm_count();

// cudaFuncSetAttribute is a per-DEVICE setting: remember per device whether a kernel's shared-memory opt-in is done
struct PerDeviceOnce {
    bool done_[64] = {};
    bool* slot() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return nullptr;
        return &done_[d];
    }
};

#define NADM_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            nadm::set_error(__VA_ARGS__);  \
            return NADM_EINVAL;            \
        }                                  \
    } while (0)

#define NADM_CHECK_LAUNCH(what)                                  \
    do {                                                         \
        cudaError_t e__ = cudaGetLastError();                    \
        if (e__ != cudaSuccess) return nadm::cuda_fail(e__, what); \
        nadm::count_launch();                                    \
    } while (0)

// ---- programmatic dependent launch (PDL), opt-in with NADM_PDL=1 --------------------------------------------------
// Every kernel of the step begins with pdl_prologue(): `launch_dependents` lets the NEXT kernel of the stream be
// scheduled as soon as all CTAs of this one have started, `wait` then blocks until the PREVIOUS kernel has completed and
// its memory is visible.  Nothing before the wait touches global memory, so the semantics are those of plain stream
// order.  Without the launch attribute (the default) the two instructions are no-ops.  With NADM_PDL=1 eager launches
// set cudaLaunchAttributeProgrammaticStreamSerialization: measured +3.5 % on the forward-only Q pass (launch and
// CTA-scheduling latency of each kernel boundary hidden); inside a captured CUDA graph the programmatic edges were
// slower than the graph's ordinary edges (0.451 vs 0.441 ms per step), so capturing launches never set it.  It stays
// opt-in until the host-fed loop (two streams, events) has been through the device tests with it.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &capturing) != cudaSuccess) capturing = cudaStreamCaptureStatusActive;
    attr[0].val.programmaticStreamSerializationAllowed =
        (pdl_enabled() && capturing == cudaStreamCaptureStatusNone) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- grid-wide barrier for the persistent kernels (one CTA per SM, grid <= SM count): OPT-IN, NADM_GRIDBAR=1 --------
// The tensor-core kernels end with a sum over their CTAs' partial results, done by a separate small kernel.  The
// experiment behind this switch: the CTAs meet at a barrier after writing their partials and each reduces its share of
// the outputs out of L2, saving a launch.  MEASURED (profiles/r2_gridbar_ab.txt, clock64 phases of CTA 0): the barrier
// costs 13-17 k cycles (the __threadfence() after 51 KB of partials + waiting for the slowest CTA), the share another
// 8-13 k (three dependent round trips to L2 at ~1.3 us each) = 11-15 us against 5-6 us for the separate kernel, whose
// 200 k threads do everything in one round trip: the step was 10 us SLOWER at both M = 500k and M = 62.5k.  Kept as a
// switch so that the result can be reproduced.
// Safe: the grid is launched with cudaLaunchAttributeCooperative (launch_coop below: the launch FAILS if the CTAs
// cannot all be resident, it cannot deadlock), self-resetting (count returns to 0, the generation only grows: replayable
// inside a CUDA graph without a memset node), bounded (a CTA that never arrives -> __trap() after ~seconds).  The two
// words live in a __device__ global of the library (zero at module load, one copy per device); kernels that use the
// same GridBar must not run concurrently on one device (stream order guarantees it for one caller per device, which
// is the ABI's convention).
struct GridBar { unsigned int count, gen; };
#ifdef __CUDACC__
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// All threads of every CTA of the grid call this; global writes made by any thread before it are visible to every thread
// after it (read them with ld.global.cg: L1 is not coherent).
__device__ __forceinline__ void grid_barrier(GridBar* b, unsigned int nctas) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int gen0 = ld_acquire_gpu_u32(&b->gen);        // cannot advance before this CTA has arrived
        __threadfence();
        if (atomicAdd(&b->count, 1u) == nctas - 1u) {
            atomicExch(&b->count, 0u);
            __threadfence();
            st_release_gpu_u32(&b->gen, gen0 + 1u);
        } else {
            unsigned int polls = 0;
            while (ld_acquire_gpu_u32(&b->gen) == gen0) {
                __nanosleep(64);
                if (++polls > (1u << 25)) __trap();
            }
        }
        __threadfence();
    }
    __syncthreads();
}
#endif
// launch with the cooperative attribute (co-residency of the whole grid checked by the driver); never PDL
template <typename... KArgs, typename... Args>
inline cudaError_t launch_coop(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
bool gridbar_enabled();   // NADM_GRIDBAR=1: in-kernel reductions behind a grid barrier (A/B measurements; slower)

// ---- tiling constants shared by kernels and the workspace query --------------------------------------------------
constexpr int kMaxParts = 640;        // upper bound on per-CTA partial slabs (encoder slabs / decoder CTAs)
constexpr int kStreamWarps = 8;       // warps per CTA in the lane<->byte streaming kernels
constexpr int kTileSnps = 128;        // SNPs per CTA tile in those kernels: one byte (4 SNPs) per lane
constexpr int kEncTileSnps = 256;
constexpr int kMaxDynSmem = 226 * 1024;  // dynamic shared memory opt-in (227 KB per CTA minus static + reserve)

struct AdamCoef {
    float beta1, beta2, one_minus_beta1, one_minus_beta2, step_size, inv_bc2_sqrt, eps;
    int enabled;
    const void* dev;   // optional device copy of the first 32 bytes (nadm_adam_t.device_coef): read at kernel start
};

inline void adam_coefficients(float lr, float beta1, float beta2, float eps, long long step, AdamCoef& c) {
    double bc1 = 1.0 - pow((double)beta1, (double)step);
    double bc2 = 1.0 - pow((double)beta2, (double)step);
    c.beta1 = beta1;
    c.beta2 = beta2;
    c.one_minus_beta1 = 1.0f - beta1;
    c.one_minus_beta2 = 1.0f - beta2;
    c.step_size = (float)((double)lr / bc1);
    c.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    c.eps = eps;
    c.enabled = 1;
}

inline AdamCoef make_adam(const nadm_adam_t* a) {
    AdamCoef c{};
    if (!a) return c;
    adam_coefficients(a->lr, a->beta1, a->beta2, a->eps, a->step, c);
    c.dev = a->device_coef;
    return c;
}

// ---- deferred reductions: nadm_encoder_fwd_deferred / nadm_decoder_step_deferred leave the sum over their CTAs'
// partials to the kernel that consumes the result (nadm_mlp_fwd / nadm_mlp_bwd), which saves a kernel boundary and a round
// trip through L2 per reduction.  The hand-over is a host-side record per host thread (one host thread per device is the
// ABI's convention), keyed by the destination pointer (Z, resp. dQ): the consumer looks its argument up.
struct DeferredZ {
    const float* Z;             // destination the record belongs to (NULL: nothing pending)
    const long long* part;      // nparts x B x 8 exact integer partial sums
    const float* vmax;          // nparts per-CTA |max| (-> the CTA's power-of-two scale)
    int nparts, B;
};
struct DeferredDQ {
    const float* dQ;            // destination (NULL: nothing pending)
    const float* part;          // nparts x B x cols_p
    const float* loss_part;     // nparts partial losses, or NULL (gradients only)
    float* loss;                // where the summed loss is added
    int nparts, B, cols_p, k, q_ld, q_off;
    size_t bytes;               // extent of the partials inside the workspace (the consumer's own scratch goes behind)
};
DeferredZ& deferred_z();
DeferredDQ& deferred_dq();
// ---- deferred parameter update of the small network: nadm_mlp_bwd_deferred leaves "sum the CTAs' gradient slabs +
// Adam on W1, b1, W2, b2, w_rms" pending; the nadm_encoder_bwd call that follows on the same dZ runs it on its epilogue
// warps while its producers fill the first tiles (the update depends on nothing that kernel computes), instead of a
// kernel of its own between the two.  Any other library call that reads the network's parameters or the loss first
// runs the pending update as the separate kernel (flush_deferred_apply).
#ifdef __CUDACC__
#define NADM_HD __host__ __device__
#else
#define NADM_HD
#endif
NADM_HD inline size_t mlp_slab_floats(int C, int H, int sumK) {
    return (size_t)(C + 1) * H + (size_t)sumK * H + (size_t)sumK + (size_t)C + 1;
}
struct ApplyJob {
    const float* part;          // nslab slabs of mlp_slab_floats(C, H, sumK) partial gradients (NULL / nslab 0: no job)
    int nslab, C, H, sumK, has_sup;
    nadm_mlp_params_t prm;
    AdamCoef adam;
    float* loss;
};
struct DeferredApply {
    const float* dZ;            // key: the dZ the pending update belongs to (NULL: nothing pending)
    ApplyJob job;
};
DeferredApply& deferred_apply();
int flush_deferred_apply(cudaStream_t st);   // nadm_mlp.cu

// tensor-core (tcgen05) encoder kernels, nadm_tc_enc.cu
int launch_enc_fwd_tc(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int B, int64_t M,
                      const float* V, int C, float* Z, void* ws, size_t ws_bytes, cudaStream_t st, int raw_mv = -1,
                      bool defer = false);
int launch_enc_bwd_tc(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int B, int64_t M,
                      const float* dZ, int C, float* V, float* Vm, float* Vv, const nadm_adam_t* adam, float* dV_out,
                      cudaStream_t st, int raw_mv = -1, int accumulate = 0, const struct ApplyJob* job = nullptr);
bool enc_bwd_runs_apply(int B);   // the default backward kernel can take a pending parameter update along (not the slab variant)
size_t enc_tc_workspace_bytes(int B);
size_t mlp_bwd_workspace_bytes(int B, int C, int H, int sumK);   // nadm_mlp.cu
bool enc_bwd_tc_supported(int B);
bool enc_fwd_slab();                   // NADM_ENC_FWD_SLAB=1 (measurement switch, nadm_tc_enc.cu)
bool enc_bwd_slab_supported(int B);   // the slab-fed backward kernel (256-byte runs per row): B <= 896
// tensor-core fused decoder, nadm_tc_dec.cu
bool dec_tc_supported(int B, int k);
int launch_dec_tc(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int B, int64_t M,
                  const float* Q, float* dQ, int q_ld, int q_off, int k, float* P, float* Pm, float* Pv,
                  const nadm_adam_t* adam, float* dP_out, float* loss, float* ws, size_t ws_bytes, cudaStream_t st,
                  bool defer = false);
// deterministic sum of per-CTA partials (nadm_stream.cu)
int launch_reduce_parts(const float* part, int nparts, int rows, int cols_p, int cols_out, float* out, int out_ld,
                        int out_off, float scale, const float* loss_part, float* loss, cudaStream_t st);
bool use_generic_kernels();   // NADM_GENERIC=1: force the CUDA-core formulation (A/B testing only)

#ifdef __CUDACC__
// power-of-two fixed-point scale for values with absolute maximum `mx`: q = rint(v * inv) fits in [-2^30, 2^30]
struct FixScale {
    float inv;      // 2^(30 - e)
    double back;    // 2^(e - 30)
};
__device__ __forceinline__ FixScale fix_scale(float mx) {
    FixScale s;
    const int E = (int)((__float_as_uint(mx) >> 23) & 0xFF);       // biased exponent: mx < 2^(E - 126)
    if (E == 255) {   // Inf or NaN among the values: the result is NaN, as a floating-point matmul would give
        s.inv = 0.f;
        s.back = __longlong_as_double(0x7FF8000000000000ll);
        return s;
    }
    if (mx == 0.f || E == 0) { s.inv = 0.f; s.back = 0.0; return s; }
    const int e = max(E - 126, -96);                                // keep 2^(30-e) a finite float
    s.inv = __uint_as_float((uint32_t)(127 + 30 - e) << 23);
    s.back = __longlong_as_double((long long)(1023 + e - 30) << 52);
    return s;
}

// torch.optim.Adam (no weight decay / amsgrad), same operation order as torch's fused kernel:
//   m = lerp(m, g, 1-b1); v = b2*v + (1-b2)*g*g; p -= step_size * m / (sqrt(v)/sqrt(bc2) + eps)
// kernels call this once on their by-value copy: coefficients written on the device by nadm_step_begin win
__device__ __forceinline__ AdamCoef adam_resolve(AdamCoef a) {
    if (a.dev != nullptr) {
        const float4 x = reinterpret_cast<const float4*>(a.dev)[0], y = reinterpret_cast<const float4*>(a.dev)[1];
        a.beta1 = x.x; a.beta2 = x.y; a.one_minus_beta1 = x.z; a.one_minus_beta2 = x.w;
        a.step_size = y.x; a.inv_bc2_sqrt = y.y; a.eps = y.z;
        a.enabled = 1;
    }
    return a;
}
__device__ __forceinline__ float adam_apply(float p, float g, float& m, float& v, const AdamCoef& c) {
    m = m + (g - m) * c.one_minus_beta1;
    v = c.beta2 * v + c.one_minus_beta2 * g * g;
    float denom = sqrtf(v) * c.inv_bc2_sqrt + c.eps;
    return p - c.step_size * (m / denom);
}
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
// element i of the network's flat gradient (slab layout: [(C+1) x H | sumK x H | sumK | C | 1]) -> its parameter
__device__ __forceinline__ void mlp_apply_param(size_t i, float g, const ApplyJob& jb, const AdamCoef& adam) {
    const int C = jb.C, H = jb.H, sumK = jb.sumK;
    const nadm_mlp_params_t& prm = jb.prm;
    const size_t n1 = (size_t)(C + 1) * H, n2 = n1 + (size_t)sumK * H;
    if (i < n1) {
        const int c = (int)(i / H), jj = (int)(i % H);
        if (c < C) adam_store(prm.W1, prm.m_W1, prm.v_W1, prm.g_W1, (int64_t)jj * C + c, g, adam);
        else adam_store(prm.b1, prm.m_b1, prm.v_b1, prm.g_b1, jj, g, adam);
    } else if (i < n2) {
        adam_store(prm.W2, prm.m_W2, prm.v_W2, prm.g_W2, (int64_t)(i - n1), g, adam);
    } else if (i < n2 + sumK) {
        adam_store(prm.b2, prm.m_b2, prm.v_b2, prm.g_b2, (int64_t)(i - n2), g, adam);
    } else if (i < n2 + sumK + C) {
        adam_store(prm.w_rms, prm.m_w_rms, prm.v_w_rms, prm.g_w_rms, (int64_t)(i - n2 - sumK), g, adam);
    } else if (jb.has_sup) {
        *jb.loss += g;
    }
}
// Sum over the slabs in EXACTLY the association of mlp_bwd_apply_kernel (8 interleaved groups of two chains each, the
// groups' sums added pairwise in order): the two forms of the update are bit-identical.  Group q of element i:
__device__ __forceinline__ float mlp_apply_group_sum(const float* __restrict__ part, size_t n, size_t i, int nslab, int q) {
    float g0 = 0.f, g1 = 0.f;
    if (nslab <= 13 * 8) {
        // every slab of the group in flight at once (this runs on warps that have better things to wait for); adding the
        // +0 of an absent slab changes nothing, so the association is the loop's below
        float v[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) {
            const int z = q + 8 * k;
            v[k] = (z < nslab) ? __ldcg(part + (size_t)z * n + i) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 13; k += 2) g0 += v[k];
#pragma unroll
        for (int k = 1; k < 13; k += 2) g1 += v[k];
        return g0 + g1;
    }
    int z = q;
    for (; z + 8 < nslab; z += 16) {
        g0 += __ldcg(part + (size_t)z * n + i);
        g1 += __ldcg(part + (size_t)(z + 8) * n + i);
    }
    if (z < nslab) g0 += __ldcg(part + (size_t)z * n + i);
    return g0 + g1;
}
// The share of CTA `cta` of `nctas` of a pending update, run by `nthr` threads (whole warps, t = 0 .. nthr - 1): lane
// pairs (2 p, 2 p + 1) take one parameter, four groups each, and meet through two shuffles.
__device__ __forceinline__ void mlp_apply_share(const ApplyJob& jb, int cta, int nctas, int t, int nthr) {
    const AdamCoef adam = adam_resolve(jb.adam);
    const size_t n = mlp_slab_floats(jb.C, jb.H, jb.sumK);
    const size_t i0 = n * (size_t)cta / (size_t)nctas, i1 = n * (size_t)(cta + 1) / (size_t)nctas;
    const int half = t & 1;
    const size_t cnt = i1 - i0, cnt_pad = (cnt + 15) & ~(size_t)15;     // whole warps stay in the loop (shuffles)
    for (size_t pi = (size_t)(t >> 1); pi < cnt_pad; pi += (size_t)(nthr >> 1)) {
        const bool ok = pi < cnt;
        const size_t i = i0 + (ok ? pi : 0);
        float r[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) r[q] = ok ? mlp_apply_group_sum(jb.part, n, i, jb.nslab, 4 * half + q) : 0.f;
        const float a = r[0] + r[1], b = r[2] + r[3];                  // (red[q] + red[q + 1]) of this half
        const float a1 = __shfl_xor_sync(0xffffffffu, a, 1), b1 = __shfl_xor_sync(0xffffffffu, b, 1);
        if (ok && half == 0) {
            float g = 0.f;
            g += a; g += b; g += a1; g += b1;                           // 0 + (r0+r1) + (r2+r3) + (r4+r5) + (r6+r7)
            mlp_apply_param(i, g, jb, adam);
        }
    }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// 2-bit code -> 2*x as float (0,1,2; missing 3 -> 0).  x = code/2 with missing trained as 0
// (neural_admixture.py:169-170); the factor 1/2 is applied once to the accumulated sums (exact: power of two).
__device__ __forceinline__ float code_to_2x(unsigned c) {
    c = (c == 3u) ? 0u : c;
    return __uint_as_float(0x4B000000u | c) - 8388608.0f;
}
// clear every 2-bit field equal to 3 in a packed word
__device__ __forceinline__ unsigned clear_missing(unsigned w) {
    unsigned m3 = w & (w >> 1) & 0x55555555u;
    return w ^ (m3 | (m3 << 1));
}

// Reduce N (power of two <= 32) per-lane values across the warp.  On return v[0] of lane l holds the warp-wide sum
// of element (l / (32/N)); lanes with l % (32/N) == 0 are the canonical owners.  N-1 + log2(32/N) shuffles.
template <int N>
__device__ __forceinline__ float warp_reduce_vec(float (&v)[N], int lane) {
    int width = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1) {
        const bool up = (lane & width) != 0;
#pragma unroll
        for (int j = 0; j < n / 2; ++j) {
            float send = up ? v[j] : v[j + n / 2];
            float keep = up ? v[j + n / 2] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, width);
        }
        width >>= 1;
    }
#pragma unroll
    for (; width >= 1; width >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], width);
    return v[0];
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double warp_sum_d(double x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
#endif  // __CUDACC__

}  // namespace nadm
