/* nadm_b200.h — C ABI of the B200 (sm_100a) Neural ADMIXTURE hot-path library (libnadm_b200.so).
 *
 * Boundary.  The reference exposes its native code to Python as a pybind11 module JIT-built from
 * neural_admixture/src/utils_c/pack2bit.cu (PYBIND11_MODULE at pack2bit.cu:144-147, loaded at model/train.py:122-125)
 * and does everything else on the hot path through PyTorch eager ops inside
 * neural_admixture/model/neural_admixture.py (Q_P.forward :157-177, NeuralAdmixture._run_epoch :394-417).
 * This header is what a binding for that path binds instead: every entry point is `extern "C"`, takes plain device
 * pointers + sizes + a CUDA stream, allocates nothing, retains no pointer, and is asynchronous on `stream`.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller unless marked [host];
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return value: 0 on success, negative NADM_E* on failure; nadm_last_error() gives the message
 *     (thread-local, valid until the next call on the same thread);
 *   - genotype storage: sample-major 2-bit packed, SNP 4c+i in bits 2i..2i+1 of byte c of a row
 *     (pack2bit.cu:26-31); codes 0,1,2 = genotype, 3 = missing (trained as 0: neural_admixture.py:169-170);
 *     `pitch` = bytes between consecutive sample rows; the streaming kernels require pitch % 16 == 0, a 16-byte
 *     aligned base and zero bits past SNP M-1 in every row (nadm_pack2bit produces that);
 *   - a minibatch is B rows: row b is `row_idx[b]` when row_idx != NULL, else `row0 + b`
 *     (replaces Dataset_admixture.__getitem__ + DataLoader collate, src/loaders.py:33,70-72);
 *   - parameters use the reference's own layouts: V is M x C row-major (neural_admixture.py:129-130), each head's
 *     P is M x k row-major (= decoders.decoders[i].weight, :73-74), W1 is H x C, W2 (all heads concatenated along
 *     rows) is sumK x H;  Adam moments have the layout of their parameter.
 */
#ifndef NADM_B200_H
#define NADM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NADM_OK 0
#define NADM_EINVAL (-1)   /* bad argument (shape, alignment, unsupported C / k) */
#define NADM_ECUDA (-2)    /* a CUDA runtime call or kernel launch failed */
#define NADM_ENOSPC (-3)   /* workspace too small */

#define NADM_MAX_C 16      /* n_components supported by the streaming kernels */
#define NADM_MAX_K 16      /* populations per head */
#define NADM_MAX_HEADS 32

/* Adam hyper-parameters of one optimizer step (torch.optim.Adam, betas (0.9,0.95): neural_admixture.py:197-204).
 * `step` is the 1-based step count AFTER increment (torch's state['step']).
 * `device_coef`: optional DEVICE pointer to the NADM_ADAM_COEF_BYTES coefficient block that nadm_step_begin writes.
 * When it is not NULL the kernels read the step's bias-corrected coefficients from it and ignore lr/betas/eps/step, so
 * the launch arguments of a training step no longer change from step to step and the whole step can be captured
 * into a CUDA graph and replayed. */
typedef struct nadm_adam {
    float lr;
    float beta1;
    float beta2;
    float eps;
    int32_t step;
    int32_t reserved;          /* 0 */
    const void* device_coef;
} nadm_adam_t;
#define NADM_ADAM_COEF_BYTES 32

int nadm_version(void);
const char* nadm_last_error(void);

/* Number of kernels this library has launched from the calling process so far (bench.py's gpu_launches). */
int64_t nadm_launch_count(void);
/* How many of those were the first-generation CUDA-core kernels, which serve the shapes the tensor-core kernels do not
 * take (C > 8, B above the tensor-memory budget, misaligned parameter pointers) and are several times slower: a caller
 * that sees this number move should know it left the fast path (the Python mirror warns once, bench.py reports it). */
int64_t nadm_generic_launch_count(void);

/* ---- 2-bit pack / unpack: device -> device.  Replaces pack2bit_kernel (pack2bit.cu:10-36; the host wrapper
 * pack2bit_cpu_to_gpu :65-117 stages <=1024 unpacked rows to the device and calls it) and unpack2bit_kernel /
 * unpack2bit_gpu_to_gpu (pack2bit.cu:38-62, :120-142).  src: rows x M uint8 codes (pitch src_pitch bytes);
 * dst: rows x ceil(M/4) bytes (pitch dst_pitch >= ceil(M/4)); bytes in [ceil(M/4), dst_pitch) are zeroed. */
int nadm_pack2bit(const uint8_t* src, int64_t rows, int64_t M, int64_t src_pitch,
                  uint8_t* dst, int64_t dst_pitch, void* stream);
int nadm_unpack2bit(const uint8_t* src, int64_t rows, int64_t M, int64_t src_pitch,
                    uint8_t* dst, int64_t dst_pitch, void* stream);

/* Bytes of scratch the step functions below need for a batch of B rows, M SNPs (local slice), C components,
 * hidden width H and sumK total populations over all heads. */
size_t nadm_workspace_bytes(int32_t B, int64_t M, int32_t C, int32_t H, int32_t sumK);

/* ---- encoder projection: Z[b, :] = X[b, :] @ V, X = genotype/2 with missing -> 0.
 * Replaces `X.float()/2`, `where(X==1.5,0,X)`, `X @ self.V` (neural_admixture.py:169-172) and the per-step
 * unpack (neural_admixture.py:404-406).  Also the whole of the inference / post-training Q pass's M-wide work
 * (inference.py:71-77, neural_admixture.py:369-383).   Z: B x C. */
int nadm_encoder_fwd(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int32_t B,
                     int64_t M, const float* V, int32_t C, float* Z, void* ws, size_t ws_bytes, void* stream);
/* Same call, but the sum over the kernel's per-CTA partial results MAY be left pending: the next nadm_mlp_fwd of this
 * host thread that is given the same Z completes it inside its own kernel (each CTA sums the partials of its rows: one
 * kernel boundary and one round trip through L2 fewer per step).  Contract: nothing reads Z and nothing writes `ws`
 * before that nadm_mlp_fwd call.  Whether the reduction was deferred is the library's decision (tensor-core path, one
 * launch); otherwise this is nadm_encoder_fwd. */
int nadm_encoder_fwd_deferred(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int32_t B,
                              int64_t M, const float* V, int32_t C, float* Z, void* ws, size_t ws_bytes, void* stream);

/* ---- SNP-sharded runs: the exchange of the two small per-step messages, fused into the kernels that consume them.
 * Rank r owns a contiguous slice of the SNP axis; the B x C partial projection (after nadm_encoder_fwd) and the
 * B x sumK partial dQ plus the partial loss (after nadm_decoder_step) must be summed over the ranks — the role of DDP's
 * gradient all-reduce in the reference (model/neural_admixture.py:315-319), 25.6 KB instead of 32-96 MB.  Instead of
 * two NCCL all-reduce kernels per step, nadm_mlp_fwd / nadm_mlp_bwd do the exchange themselves when given a
 * nadm_xchg_t: every CTA stores its rows' partial values into all peers' exchange areas over NVLink (peer-mapped device
 * memory) as 8-byte {value, exchange number} pairs (one atomic store each: no fence, no separate flag), spins on the
 * pairs the peers store into its own area until they carry the same exchange number, and sums the values in rank
 * order (so every rank obtains bit-identical sums).  One area per rank, allocated with nadm_ipc_alloc and opened by the
 * peers with nadm_ipc_open (CUDA IPC: one process per GPU on one node).  `seq` is LOCAL device memory of
 * NADM_XCHG_SEQ_WORDS uint32, zero-initialised, private to the rank.  xchg == NULL: no exchange (single GPU, or the
 * caller all-reduces itself, e.g. with NCCL). */
#define NADM_MAX_RANKS 8
#define NADM_XCHG_MAX_CTAS 1024
#define NADM_XCHG_SEQ_WORDS (2 * NADM_XCHG_MAX_CTAS)
typedef struct nadm_xchg {
    int32_t world, rank;
    void* area[NADM_MAX_RANKS];   /* exchange area of every rank, as mapped into THIS process (area[rank] = own) */
    uint32_t* seq;                /* this rank's private exchange counters */
    int64_t slot_floats;          /* capacity of one slot: >= max(B * C, B * sumK + 1) over all calls */
} nadm_xchg_t;
size_t nadm_xchg_area_bytes(int64_t slot_floats);
/* cudaMalloc + zero + cudaIpcGetMemHandle (handle_out: 64 bytes [host]); open / close a peer's area; free one's own. */
int nadm_ipc_alloc(size_t bytes, void** dev_ptr, void* handle_out);
int nadm_ipc_open(const void* handle, void** dev_ptr);
int nadm_ipc_close(void* dev_ptr);
int nadm_ipc_free(void* dev_ptr);

/* ---- the small replicated network: RMSNorm(C, eps 1e-8) -> Linear(C,H)+ReLU -> per-head Linear(H,k) -> softmax
 * (neural_admixture.py:135-144 construction, :173-176 forward).  ks [host]: nheads head sizes; Q: B x sumK
 * (heads side by side); Hh: B x H post-ReLU activations and rinv: B (kept for the backward).
 * xchg != NULL: Z holds this rank's PARTIAL projection on entry and the sum over the ranks on return. */
int nadm_mlp_fwd(float* Z, int32_t B, int32_t C, int32_t H, const float* w_rms, const float* W1,
                 const float* b1, const float* W2, const float* b2, const int32_t* ks, int32_t nheads,
                 float* rinv, float* Hh, float* Q, const nadm_xchg_t* xchg /*[host], may be NULL*/, void* stream);

/* ---- fused decoder for ONE head: R = clamp(Q_k P_k^T, 0, 1); loss += BCE_sum(R, X); G = dLoss/dR through the
 * clamp mask; dQ[:, q_off:q_off+k] = G P_k; dP = G^T Q_k; then Adam on P_k and clamp to [0,1].
 * Replaces NeuralDecoder.forward (neural_admixture.py:83-98), BCELoss(sum) (:288,:431), their autograd backward
 * (:410), the P part of optimizer.step() (:411) and restrict_P (:179-185,:412).  The B x M reconstruction, the
 * float X and G never touch HBM.
 *   Q, dQ : B x q_ld (q_ld = sumK), this head occupies columns [q_off, q_off+k)
 *   P, Pm, Pv : M x k parameter and Adam moments, updated in place when adam != NULL
 *   dP_out : optional M x k raw gradient output (tests); may be NULL
 *   loss : 1 float, the head's loss is ADDED to it (zero it once per step); NULL = gradients only (the loss value
 *          is not needed by the backward: the reference only reports it, neural_admixture.py:414-417). */
int nadm_decoder_step(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int32_t B,
                      int64_t M, const float* Q, float* dQ, int32_t q_ld, int32_t q_off, int32_t k,
                      float* P, float* Pm, float* Pv, const nadm_adam_t* adam /*[host], NULL = no update*/,
                      float* dP_out, float* loss, void* ws, size_t ws_bytes, void* stream);
/* Same call, but the sum over the kernel's per-CTA partials of dQ[:, q_off:q_off+k] (and of the head's loss) MAY be left
 * pending: the next nadm_mlp_bwd of this host thread that is given the same dQ (and the same ws) completes it inside its
 * own kernel; a later nadm_decoder_step(_deferred) on the same dQ completes it first.  Contract: nothing reads those
 * columns of dQ or *loss, and nothing else writes `ws`, before that call.  Use it for the last head of a step. */
int nadm_decoder_step_deferred(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int32_t B,
                               int64_t M, const float* Q, float* dQ, int32_t q_ld, int32_t q_off, int32_t k,
                               float* P, float* Pm, float* Pv, const nadm_adam_t* adam, float* dP_out, float* loss,
                               void* ws, size_t ws_bytes, void* stream);

/* ---- backward of the small network + Adam on its parameters.  dQ: B x sumK (sum of the decoder's and, when
 * labels != NULL, of supervised_loss_weight * CrossEntropyLoss(sum)(Q_0, labels) — neural_admixture.py:293,:473 —
 * which this call adds itself, also adding that loss term to *loss).  Outputs dZ: B x C.  Updates w_rms, W1, b1,
 * W2, b2 and their moments in place when adam != NULL; raw gradients go to the optional g_* buffers.
 * xchg != NULL: dQ and *loss hold this rank's PARTIAL sums on entry and the sums over the ranks on return. */
typedef struct nadm_mlp_params {
    float* w_rms; float* W1; float* b1; float* W2; float* b2;          /* parameters             */
    float* m_w_rms; float* m_W1; float* m_b1; float* m_W2; float* m_b2; /* Adam first moments     */
    float* v_w_rms; float* v_W1; float* v_b1; float* v_W2; float* v_b2; /* Adam second moments    */
    float* g_w_rms; float* g_W1; float* g_b1; float* g_W2; float* g_b2; /* optional raw gradients */
} nadm_mlp_params_t;

int nadm_mlp_bwd(float* dQ, const float* Q, const float* Hh, const float* Z, const float* rinv,
                 int32_t B, int32_t C, int32_t H, const int32_t* ks /*[host]*/, int32_t nheads,
                 const int64_t* labels /*B, or NULL*/, float sup_weight,
                 const nadm_mlp_params_t* params /*[host]*/, const nadm_adam_t* adam /*[host]*/,
                 float* dZ, float* loss, void* ws, size_t ws_bytes, const nadm_xchg_t* xchg /*[host], may be NULL*/,
                 void* stream);
/* Same call, but the parameter update (sum of the per-CTA gradient slabs in `ws` + Adam on W1, b1, W2, b2, w_rms + the
 * supervised loss term) is left pending: the nadm_encoder_bwd of this host thread that follows on the same dZ runs it on
 * its epilogue warps, inside the same kernel (the update depends on nothing that kernel computes; bit-identical to the
 * separate kernel).  Any other call of this library that reads the network's parameters or finishes a step runs the
 * pending update first, as its own kernel.  Contract: nothing else reads the parameters, *loss or writes `ws` in between. */
int nadm_mlp_bwd_deferred(float* dQ, const float* Q, const float* Hh, const float* Z, const float* rinv,
                          int32_t B, int32_t C, int32_t H, const int32_t* ks, int32_t nheads, const int64_t* labels,
                          float sup_weight, const nadm_mlp_params_t* params, const nadm_adam_t* adam, float* dZ,
                          float* loss, void* ws, size_t ws_bytes, const nadm_xchg_t* xchg, void* stream);

/* ---- encoder backward: dV = X^T dZ, then Adam on V.  Replaces the autograd of `X @ self.V`
 * (neural_admixture.py:172,:410) and the V part of optimizer.step() (:411).  dV_out optional. */
int nadm_encoder_bwd(const uint8_t* packed, int64_t pitch, const int64_t* row_idx, int64_t row0, int32_t B,
                     int64_t M, const float* dZ, int32_t C, float* V, float* Vm, float* Vv,
                     const nadm_adam_t* adam /*[host], NULL = no update*/, float* dV_out,
                     void* ws, size_t ws_bytes, void* stream);

/* ---- fp64 binomial log-likelihood of the packed matrix under (Q, P), missing skipped, eps-clamped
 * (src/utils_c/utils.pyx:17-40, called at model/train.py:139,145).  Q: N x k, P: M x k (fp32); out: 1 double. */
int nadm_loglikelihood(const uint8_t* packed, int64_t pitch, int64_t N, int64_t M, const float* Q,
                       const float* P, int32_t k, double eps, double* out, void* ws, size_t ws_bytes,
                       void* stream);

/* ---- PLINK .bed -> sample-major 2-bit packed matrix on the device ("next" row f4 of the scope table).  Replaces
 * SNPReader._read_bed (src/snp_reader.py:16-45) -> utils_c.read_bed (src/utils_c/utils.pyx:43-68, LUT [2,3,1,0] into an
 * N x M uint8 host array), the allele flip `G if G.mean() < 1 else 2 - G` (snp_reader.py:110) and pack2bit_cpu_to_gpu
 * (pack2bit.cu:65-117) without the one-byte-per-genotype intermediate.
 *   bed     : M rows (SNPs) of bed_pitch >= ceil(N/4) bytes, the .bed payload after its 3 magic bytes (device copy)
 *   snp0    : column of the first SNP of this call in the destination (multiple of 128): a file can be converted in
 *             chunks of SNP rows; every chunk but the last must hold a multiple of 128 SNPs
 *   flip    : 0 = codes as read_bed gives them; 1 = additionally g -> 2 - g (missing stays 3)
 *   counts  : optional 4 x uint64; counts[1..3] += number of codes 1, 2, 3 read (BEFORE the flip) in this call, so
 *             the caller can evaluate the reference's `G.mean() < 1` test; zero it first
 * nadm_flip_packed applies g -> 2 - g in place to a packed N x M matrix (missing and the zero row tails unchanged). */
int nadm_bed_to_packed(const uint8_t* bed, int64_t bed_pitch, int64_t N, int64_t M, int64_t snp0, int32_t flip,
                       uint8_t* dst, int64_t dst_pitch, uint64_t* counts, void* stream);
int nadm_flip_packed(uint8_t* packed, int64_t pitch, int64_t N, int64_t M, void* stream);

/* ---- device-side step bookkeeping: what makes a training step replayable as a CUDA graph (the reference's loop
 * body, neural_admixture.py:403-414, re-launches ~25 eager kernels from Python every step).
 * counters: 2 x int64 on the device: [0] = index of the next minibatch inside `order`, [1] = optimizer steps done.
 * nadm_step_begin: row_idx_out[i] = order[counters[0] * stride + i] for i < B (the minibatch's rows: the sampler's
 *   permutation, src/loaders.py:26-33, lives on the device), and the Adam coefficients of step counters[1] + 1
 *   (hyper: lr, betas, eps [host]) -> coef_out (device, NADM_ADAM_COEF_BYTES); *loss_accum = 0 when not NULL (the
 *   step's loss accumulator, which nadm_decoder_step / nadm_mlp_bwd add to).
 * nadm_step_end: losses_out[counters[0]] = *loss when both are given; counters[0] += 1; counters[1] += 1. */
int nadm_step_begin(const int64_t* order, int64_t order_len, int64_t* counters, int64_t stride, int32_t B,
                    int64_t* row_idx_out, const nadm_adam_t* hyper, void* coef_out, float* loss_accum, void* stream);
int nadm_step_end(int64_t* counters, const float* loss, float* losses_out, void* stream);
/* One bookkeeping kernel per replayed step instead of two.  counters: 4 x int64 (zero-initialised): [0], [1] as above,
 * [2] = a step is pending, [3] = device address of the pending step's loss accumulator (0: its loss is not recorded).
 * nadm_step_next: if a step is pending, finish it as nadm_step_end would (losses_out[counters[0]] = its loss when
 *   recorded, both counters += 1); then begin the next one as nadm_step_begin would and leave it pending
 *   (record_loss != 0: its loss accumulator is remembered for the following nadm_step_next / nadm_step_flush).
 * nadm_step_flush: finish the pending step, if any (after the last replay of a run of steps). */
int nadm_step_next(const int64_t* order, int64_t order_len, int64_t* counters, int64_t stride, int32_t B,
                   int64_t* row_idx_out, const nadm_adam_t* hyper, void* coef_out, float* loss_accum,
                   int32_t record_loss, float* losses_out, void* stream);
int nadm_step_flush(int64_t* counters, float* losses_out, void* stream);

/* ---- randomized-SVD products on the packed matrix ("next" row f2).  Replace multiply_A_omega / multiply_QT_A
 * (src/utils_c/rsvd.pyx:16-50, driven by src/svd.py:49-71): naive OpenMP triple loops over the N x M uint8 matrix.
 * They multiply by the reader's uint8 VALUES — 0, 1, 2 and, for a missing genotype, `missing_value`: 3 as read_bed
 * writes it, 255 after the reader's `2 - G` allele flip (snp_reader.py:110) — with no halving and no missing -> 0.
 * Same exact-integer tensor-core contraction as nadm_encoder_fwd / nadm_encoder_bwd (fixed-point digits of the fp32
 * factor); at most 8 columns per call, wider factors go in column chunks.
 *   nadm_geno_matmul  : Y  [N x K] = A Omega     Omega: M x K row-major
 *   nadm_geno_matmul_t: Bt [M x K] = A^T Q       Q: N x K row-major  (the reference's B = Q^T A is Bt^T) */
int nadm_geno_matmul(const uint8_t* packed, int64_t pitch, int64_t N, int64_t M, const float* Omega, int32_t K,
                     int32_t missing_value, float* Y, void* ws, size_t ws_bytes, void* stream);
int nadm_geno_matmul_t(const uint8_t* packed, int64_t pitch, int64_t N, int64_t M, const float* Q, int32_t K,
                       int32_t missing_value, float* Bt, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NADM_B200_H */
