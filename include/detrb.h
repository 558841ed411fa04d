/* detrb.h -- C ABI of libdetrb.so: the B200 (sm_100a) DETR train-step hot path.
 *
 * The reference (Visual-Behavior/detr-tensorflow) is pure Python/TensorFlow and has no
 * FFI / plugin interface; its boundary is the Python API (get_detr_model / get_losses /
 * hungarian_matching / setup_optimizers / training.fit).  This header is the C ABI that
 * sits directly under our Python mirror of that API.  Each entry point names the reference
 * computation (file:line under detr_tf/) that it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers + sizes; every pointer is a DEVICE pointer unless noted.
 *   - the caller owns all memory; the library never allocates or frees device memory and
 *     keeps no pointer past the call.  All work is enqueued on `stream`; no host sync, no
 *     allocation => every call is CUDA-graph capturable.
 *   - return value: 0 ok, <0 error (DETRB_E_*); detrb_last_error() gives a message
 *     (thread-local, host pointer).  No C++ exception crosses the ABI.
 *   - activations are NHWC / row-major bf16; accumulators, losses, gradients of parameters
 *     and optimizer state are fp32; matcher indices are int64 like the reference's.
 *   - there is no CPU fallback: detrb_check_device() fails unless the device is CC 10.x.
 *
 * Parity precision ("split" storage).  The reference computes in fp32 (custom_layers.py:49-50, transformer.py:317,340).
 * Every entry point that reads or writes bf16 ACTIVATIONS takes a plane stride `split` (in elements; 0 = plain bf16, the
 * throughput mode).  With split != 0 a tensor element x lives as a PAIR of bf16 values, hi = bf16(x) at the given pointer and
 * lo = bf16(x - hi) `split` elements further (same shape and strides): 16 significant bits, the same bytes per element as fp32.
 * The GEMM / convolution / weight-gradient kernels then issue three tensor-core passes per product (hi*hi + lo*hi + hi*lo,
 * fp32 accumulation in tensor memory) over the same TMA / tcgen05 pipeline; weights are paired the same way (`wsplit`), the
 * attention core runs in fp32 arithmetic, all other kernels read hi + lo and write the rounded pair.  This is what the
 * fp32-tolerance parity tests run; it is selected per call, not per process.
 */
#ifndef DETRB_H
#define DETRB_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *detrb_stream_t;          /* cudaStream_t */
typedef uint16_t detrb_bf16;           /* raw bfloat16 bits */

enum {
    DETRB_OK = 0,
    DETRB_E_BADARG = -1,
    DETRB_E_SHAPE = -2,
    DETRB_E_ARCH = -3,
    DETRB_E_CUDA = -4,
    DETRB_E_NUMERIC = -5
};

int detrb_version(void);
const char *detrb_last_error(void);
/* fails (DETRB_E_ARCH) unless the current device is compute capability 10.x */
int detrb_check_device(void);
/* programmatic dependent launch between consecutive kernels of the stream (default on); returns the previous setting */
int detrb_set_pdl(int enable);

/* ------------------------------------------------------------------------------------------
 * Handles (one handle per GPU / rank).  The reference keeps its per-model state in the Keras model object
 * (networks/detr.py:116-204); below the Python mirror that state is (a) the device the step runs on and (b) the kernel-policy
 * switches.  A handle owns both; it owns NO device memory (buffers, workspaces and weights stay with the caller, as above).
 *
 *   detrb_create(device, &h)   checks that `device` exists and is compute capability 10.x (DETRB_E_ARCH otherwise: there is no
 *                              fallback path), records it and gives the handle the creating thread's current policy (the defaults below,
 *                              or what the DETRB_* environment variables say, unless the thread called detrb_set_* before).
 *   detrb_handle_set / _get    change / read ONE switch of the handle (nothing else is touched; unknown option: DETRB_E_BADARG).
 *   detrb_bind(h)              makes h current for the CALLING THREAD: cudaSetDevice(h's device) and h's switches become the
 *                              thread's switches.  Every detrb_* compute call of that thread then runs under them.
 *   detrb_destroy(h)           frees the host object (NULL is accepted).
 *
 * The switches are thread-local: the detrb_set_* functions of this header change the calling thread's current value (a thread
 * that never bound a handle runs the defaults, initialised from the DETRB_* environment variables), so two threads that bound two
 * handles -- two GPUs driven from one process -- do not see each other's settings.  A handle is thread-compatible (one thread
 * at a time); detrb_last_error() is the calling thread's message, so it is also "the handle's" last error.
 * ------------------------------------------------------------------------------------------ */
typedef struct detrb_handle detrb_handle_t;
enum {
    DETRB_OPT_PDL = 0,             /* detrb_set_pdl            default 1 */
    DETRB_OPT_TC = 1,              /* detrb_set_tc             default 1 */
    DETRB_OPT_TC_CONV = 2,         /* detrb_set_tc_conv        default 1 */
    DETRB_OPT_TC_TMA_EPILOGUE = 3, /* detrb_set_tc_tma_epilogue default 1 */
    DETRB_OPT_TC_PERSISTENT = 4,   /* detrb_set_tc_persistent  default 1 (auto policy) */
    DETRB_OPT_TC_STREAM = 5,       /* detrb_set_tc_stream      default 1 (auto policy) */
    DETRB_OPT_TC_HALO = 6,         /* detrb_set_tc_halo        default 1 */
    DETRB_OPT_TC_PAIR = 7,         /* detrb_set_tc_pair        default -1 (follow DETRB_PAIR, else off) */
    DETRB_OPT_TC_WGRAD = 8,        /* detrb_set_tc_wgrad       default 1 */
    DETRB_OPT_TC_ATTN = 9,         /* detrb_set_tc_attn        default 1 */
    DETRB_OPT_COUNT = 10
};
int detrb_create(int device, detrb_handle_t **out);
int detrb_destroy(detrb_handle_t *h);
int detrb_handle_device(const detrb_handle_t *h);                       /* the device ordinal, or DETRB_E_BADARG */
int detrb_handle_set(detrb_handle_t *h, int option, int value);
int detrb_handle_get(const detrb_handle_t *h, int option, int *value);
int detrb_bind(const detrb_handle_t *h);

/* ------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear layer:  C[M,N] = epilogue( gather(A)[M,K] * W[N,K]^T )
 * Replaces tf Conv2D+ZeroPadding2D+FrozenBatchNorm2D+ReLU(+residual) (networks/resnet_backbone.py:
 * 20-26, 116-136; custom_layers.py:21-24), Conv2D input_proj (detr.py:44), Linear
 * (custom_layers.py:49-50) and every tf.matmul of the transformer (transformer.py:294-347),
 * plus their data-gradients (tape.gradient, training.py:23).
 *
 *   A     : NHWC tensor [batch, IH, IW, Cin] (bf16, pixel stride lda elements).  A plain matrix
 *           [M,K] is the case batch=1, IH=1, IW=M, KH=KW=1, Cin=K.
 *   mode 0: forward gather   iy = oy*stride - pad + kh
 *   mode 1: transposed gather (data gradient of a strided conv; A is dy at the conv's OUTPUT
 *           resolution IHxIW, the GEMM rows are the conv's INPUT pixels OHxOW):
 *           t = oy + pad - kh ; valid iff t % stride == 0 ; iy = t / stride
 *   W     : [N, K] bf16, K = KH*KW*Cin ordered (kh, kw, c); ldw = row stride.
 *   stem  : Cin == 4 (RGB padded to 4 channels) with KW padded to 8 (7x8 taps, K = 224).
 *   epilogue, in this order:  acc (+bias[n]) (+residual[m,n]) (relu) (*mask: (mask[m,n]>0)*mask_scale, or its 1-bit form mask_bits)
 *           (sigmoid) (dropout keep/(1-p), counter-based on (m,n))  -> C bf16 and/or Cf fp32.
 *           With drop_p > 0 the residual is the un-dropped skip path and is added AFTER the dropout.
 *   scatter: if out_stride > 1 the GEMM row (b,oy,ox) is written to pixel (b, oy*out_stride,
 *           ox*out_stride) of a [batch, SH, SW, N] tensor (data gradient of a 1x1 stride-s conv).
 *   accumulate: C += result (read-modify-write in bf16) instead of C = result.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const detrb_bf16 *A;
    const detrb_bf16 *W;
    int M, N, K;
    int lda, ldw;
    int batch, IH, IW, Cin;
    int OH, OW;
    int KH, KW, stride, pad;
    int mode;
    const float *bias;
    const detrb_bf16 *residual; int ldr;
    const detrb_bf16 *mask; int ldm; float mask_scale;
    int relu, sigmoid;
    float drop_p; uint64_t seed; uint32_t site;
    const uint64_t *seed_ptr;    /* optional device word XOR-ed into seed (lets a captured CUDA graph draw fresh masks) */
    detrb_bf16 *C; int ldc;
    float *Cf; int ldcf;
    int out_stride, SH, SW;
    int accumulate;
    /* sliding-window A operand (0 = ordinary row-major A): k-block j (64 columns) of row m is the 64 contiguous elements at
     * A + (m + j*a_kb_rows)*lda, and lda may be smaller than 64 (overlapping rows).  A zero-padded NHWC tensor with 16-channel
     * pixels read this way IS the im2col matrix of a 4-tap-wide convolution at the pitch of the padded input (the
     * space-to-depth stem: a_kb_rows = padded width).  The caller guarantees (M + (K/64-1)*a_kb_rows)*lda + 64 readable elements.
     * Plain geometry only; tcgen05 path only (DETRB_E_SHAPE otherwise). */
    int a_kb_rows;
    /* parity precision (see the header comment): plane stride of the bf16 pairs of A / residual / mask / C (0 = plain bf16) and
     * of W.  Both zero or both non-zero. */
    int64_t split, wsplit;
    /* 1-bit ReLU masks (the backbone's backward pass; custom_layers / resnet_backbone.py:116-136 under tape.gradient): a [rows, N/8]
     * byte matrix, bit (n % 8) of byte n / 8 of row m belongs to element (m, n); row stride ldmb / ldob bytes.
     *   mask_bits: consumed like `mask` (element kept and scaled by mask_scale iff its bit is set), instead of re-reading the
     *              bf16 activation: 1/16 of the bytes.  Not together with `mask`.
     *   out_bits : produced: bit = (stored result > 0), next to C.  Rows are indexed like C (scatter rows when out_stride > 1).
     * N % 64 == 0, strides multiples of 8 bytes, 8-byte aligned bases; tcgen05 path only (DETRB_E_SHAPE otherwise). */
    const uint8_t *mask_bits; int ldmb;
    uint8_t *out_bits; int ldob;
    /* optional workspace of at least M*N + 64 bf16 elements (16-byte aligned, plain bf16 only).  With it the two stride-2 scatter
     * forms -- the data gradient of a 3x3 / stride-2 convolution (four parity-class sub-convolutions) and out_stride == 2 (the data
     * gradient of a 1x1 / stride-2 shortcut, with or without accumulate) -- write their GEMM results densely through the fast
     * TMA-store kernels and one coalesced pass scatters them (applies mask_bits, adds to C when accumulate); without it the
     * results are scattered by the GEMM's own epilogue, row by row. */
    detrb_bf16 *scratch;
} detrb_igemm_t;

int detrb_igemm(const detrb_igemm_t *p, detrb_stream_t stream);
/* Plain GEMMs (KH=KW=1, stride 1, K % 64 == 0, 16-byte aligned operands) run on the tcgen05 / TMA / TMEM kernel
 * (gemm_tc.cu) when enabled; everything else, and everything when disabled, runs on the mma.sync kernel (igemm.cu).
 * detrb_set_tc returns the previous setting.  detrb_gemm_tc_force runs the tcgen05 kernel or fails (tests; bn = 64|128|0). */
int detrb_set_tc(int enable);
int detrb_set_tc_conv(int enable);
int detrb_set_tc_tma_epilogue(int enable);
int detrb_set_tc_persistent(int enable);     /* persistent tile loop (one CTA per SM, epilogue overlapped with the next main loop) */   /* coalesced TMA-load/TMA-store epilogue of the tcgen05 kernel (default on) */   /* gathered convolutions (TMA im2col) on the tcgen05 kernel too (default on when tc is on) */
int detrb_set_tc_stream(int enable);         /* streaming kernel for the HBM-bound 1x1 layers (weights resident in shared memory): 0 off, 1 auto, 2 wherever supported */
int detrb_set_tc_halo(int enable);           /* halo-reusing row kernel for 3x3 / stride 1 / 64 -> 64 channel convolutions (conv_halo.cu): 0 off, 1 on */
/* CTA-pair persistent GEMM (gemm_pair_kernel: tcgen05 cta_group::2, 256 x BN tiles shared by the two SMs of a TPC).
 * mode 0 = off, 1 = for the 256-wide tiles, 2 = for the 128- and 256-wide tiles, -1 = follow the environment (DETRB_PAIR).
 * Returns the previous mode.  Results are bit-identical to the single-CTA persistent kernel (same products, same order). */
int detrb_set_tc_pair(int mode);
int detrb_gemm_tc_force(const detrb_igemm_t *p, int bn, detrb_stream_t stream);

/* Weight gradient  dW[N,K] (+)= rowscale[n] * sum_m dY[m,n] * gather(A)[m,k]   (fp32 atomics)
 * and optionally dbias[n] += rowscale[n] * sum_m dY[m,n].
 * gather() as in detrb_igemm mode 0 (the forward conv's own gather).  Replaces the
 * kernel/bias gradients of tape.gradient (training.py:23, optimizers.py:115). */
typedef struct {
    const detrb_bf16 *A; int lda;
    const detrb_bf16 *dY; int ldy;
    int M, N, K;
    int batch, IH, IW, Cin;
    int OH, OW;
    int KH, KW, stride, pad;
    const float *rowscale;       /* [N] or NULL */
    float *dW; int ldw;          /* fp32 [N, K] */
    float *dbias;                /* fp32 [N] or NULL */
    int a_kb_rows;               /* sliding-window A operand as in detrb_igemm_t: plain geometry, tcgen05 kernel only */
    int k_mask;                  /* 1: A is the space-to-depth stem operand -- columns of dW that do not exist in the 7x7x3 kernel
                                  * (resnet_backbone.py:11) receive no gradient */
    int64_t split;               /* parity precision: plane stride of the bf16 pairs of A and dY (0 = plain bf16) */
} detrb_wgrad_t;

int detrb_wgrad(const detrb_wgrad_t *p, detrb_stream_t stream);
/* tcgen05 weight-gradient kernel (wgrad_tc.cu; MN-major operands straight from TMA): switch + forced entry for tests */
int detrb_set_tc_wgrad(int enable);
/* tile of the tcgen05 weight-gradient kernel (process-wide developer switch for tests and tuning runs): 0 = the built-in policy,
 * 1 = 128 x 128 everywhere, 2 / 3 / 4 = 128 x 256 / 256 x 128 / 256 x 256 (out channels x k columns) wherever the shape allows.
 * All tiles compute the same sums (the split of the pixel range, hence the fp32 rounding of the merge, differs).  Returns the
 * previous mode. */
int detrb_set_wgrad_tile(int mode);
int detrb_wgrad_tc_force(const detrb_wgrad_t *p, detrb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Multi-head attention core (transformer.py:308-345): softmax(Q K^T) V per (batch, head),
 * head h = channels [32h, 32h+32); the 32^-0.5 query scaling (:307) is applied to the scores.
 * Dropout (p) on the probabilities (:341) is counter-based on (seed, site, b, h, q, k).
 * Q [B, Lq, *] bf16 with row stride ldq, K/V [B, Lk, *] (ldk, ldv); O [B, Lq, H*32] (ldo).
 * lse [B, H, Lq] fp32 = log-sum-exp of the scores (saved for the backward).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const detrb_bf16 *Q, *K, *V; int ldq, ldk, ldv;
    detrb_bf16 *O; int ldo;
    float *lse;
    int B, H, Lq, Lk;
    float scale;                 /* scores = scale * Q K^T  (head_dim^-0.5, transformer.py:307) */
    float drop_p; uint64_t seed; uint32_t site; const uint64_t *seed_ptr;
    int64_t split;               /* parity precision: plane stride of the bf16 pairs of Q / K / V / O; fp32 arithmetic (0 = plain bf16) */
} detrb_attn_fwd_t;
int detrb_attn_fwd(const detrb_attn_fwd_t *p, detrb_stream_t stream);
/* The forward runs on the tcgen05 / TMA / TMEM kernel (attention_tc.cu: S and P.V in tensor memory, thread-per-row softmax) when the
 * strides are multiples of 8 and the operands 16-byte aligned; detrb_set_tc_attn(0) routes it back to the mma.sync kernel
 * (A/B measurements, tests).  Returns the previous setting.  Both kernels draw identical dropout masks. */
int detrb_set_tc_attn(int enable);

typedef struct {
    const detrb_bf16 *Q, *K, *V, *O, *dO; int ldq, ldk, ldv, ldo, lddo;
    const float *lse;
    float *delta;                /* scratch [B, H, Lq] fp32 */
    detrb_bf16 *dQ, *dK, *dV; int lddq, lddk, lddv;
    int B, H, Lq, Lk;
    float scale;
    float drop_p; uint64_t seed; uint32_t site; const uint64_t *seed_ptr;
    int64_t split;               /* parity precision: plane stride of the bf16 pairs of every bf16 operand (0 = plain bf16) */
    int parts;                   /* 0 = everything; else a mask of the launches to enqueue: 1 = delta, 2 = dK/dV, 4 = dQ.  Lets a caller
                                  * put the dK/dV kernel on another stream than the delta -> dQ chain (dK/dV and dQ both need delta) */
} detrb_attn_bwd_t;
int detrb_attn_bwd(const detrb_attn_bwd_t *p, detrb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * LayerNormalization(epsilon=1e-5) over the last dim d == 256 (transformer.py:151-152,200-202,22).
 * fwd:  y = LN(x)*gamma+beta ; optional y2 = y + pos[(row % S)] (the `src + pos_encoding`
 *       of transformer.py:161,209,217) ; saves mean/rstd.
 * bwd:  dx = LN'(dy [+ dy2]) ; optional dx_drop = dx * keep(m,n)/(1-p) (gradient entering the
 *       dropout'ed sub-layer branch, transformer.py:169,176) ; dgamma/dbeta fp32 atomics.
 * ------------------------------------------------------------------------------------------ */
int detrb_layernorm_fwd(const detrb_bf16 *x, const float *gamma, const float *beta,
                        detrb_bf16 *y, detrb_bf16 *y2, const detrb_bf16 *pos, int S,
                        float *mean, float *rstd, int M, int64_t split, detrb_stream_t stream);
int detrb_layernorm_bwd(const detrb_bf16 *dy, const detrb_bf16 *dy2, const detrb_bf16 *x,
                        const float *gamma, const float *mean, const float *rstd,
                        detrb_bf16 *dx, detrb_bf16 *dx_drop, float drop_p, uint64_t seed, uint32_t site,
                        const uint64_t *seed_ptr, float *dgamma, float *dbeta, int M, int64_t split, detrb_stream_t stream);

/* elementwise helpers (bf16, n multiple of 8) */
/* out[r, :] = x[r, :] + pos[r % S, :]          (transformer.py:161: source + pos_encoding) */
int detrb_add_rowbcast(const detrb_bf16 *x, const detrb_bf16 *pos, detrb_bf16 *out,
                       int M, int S, int d, int64_t split, detrb_stream_t stream);
/* out = a + b (b may be NULL -> copy) */
int detrb_add(const detrb_bf16 *a, const detrb_bf16 *b, detrb_bf16 *out, int64_t n, detrb_stream_t stream);
/* fp32 NHWC3 image -> bf16 NHWC4 (4th channel 0): the layout the stem kernel gathers from */
int detrb_image_to_nhwc4(const float *img, detrb_bf16 *out, int64_t npix, detrb_stream_t stream);
/* fp32 NHWC3 image -> bf16 space-to-depth(2) tensor (channel (ry*2+rx)*3+c, 4 zero channels): turns the 7x7/stride-2 stem
 * (resnet_backbone.py:11-12) into a dense 4x4/stride-1 convolution over 16-channel (32-byte) pixels.  out is [B, HP, WP, 16]:
 * the frame's ceil(H/2) x ceil(W/2) pixels start at (pad_top, pad_left), every other position is written as zero -- with
 * pad 2/2 and HP = ceil(H/2)+3, WP = ceil(W/2)+3 the stem's zero padding is explicit and the convolution is a sliding-window
 * GEMM over the flat tensor (detrb_igemm_t.a_kb_rows = WP).  pad 0/0, HP = ceil(H/2), WP = ceil(W/2): the dense tensor. */
int detrb_image_to_s2d16(const float *img, detrb_bf16 *out, int B, int H, int W, int pad_top, int pad_left, int HP, int WP,
                         int64_t split, detrb_stream_t stream);
/* fp32 -> bf16 */
int detrb_f32_to_bf16(const float *x, detrb_bf16 *y, int64_t n, detrb_stream_t stream);
/* column sums: out[n] += scale[n]* sum_m x[m,n]  (bias gradients) */
int detrb_colsum(const detrb_bf16 *x, int ldx, int M, int N, const float *scale, float *out,
                 detrb_stream_t stream);

/* ZeroPadding2D(1)+MaxPool2D(3,2,'valid') on the stem's ReLU output (resnet_backbone.py:16-17,25-26), NHWC, C%8==0, x >= 0.
 * fwd stores the argmax tap (0..8) per output element, or 15 when the window maximum is not positive (no gradient passes the
 * stem's ReLU there); bwd routes dy to the stored tap -- the ReLU mask of the stem is applied without re-reading x.
 * x (and dx) are stored [B, XH, XW, C] with XH >= IH, XW >= IW (the image in the top-left corner; XH = IH, XW = IW: dense):
 * fwd ignores the positions outside IH x IW, bwd writes them as zeros (they are the wrapped-window rows of the sliding-window
 * stem GEMM, whose weight gradient must not see them). */
int detrb_maxpool_fwd(const detrb_bf16 *x, detrb_bf16 *y, uint8_t *argmax,
                      int B, int IH, int IW, int C, int OH, int OW, int XH, int XW, int64_t split, detrb_stream_t stream);
int detrb_maxpool_bwd(const detrb_bf16 *dy, const uint8_t *argmax, detrb_bf16 *dx,
                      int B, int IH, int IW, int C, int OH, int OW, int XH, int XW, int64_t split, detrb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Hungarian matcher (loss/hungarian_matching.py:163-203 + :27-46 -> scipy LSAP) for P = L*B
 * independent problems (L decoder layers x B images), entirely on device.
 *   logits [P, Q, C] fp32 (row stride ldl), boxes [P, Q, 4] fp32 cxcywh, targets in the padded
 *   wire format of data/processing.py:35-55: t_bbox [B,100,4] f32 (row 0 = [n,0,0,0]),
 *   t_class [B,100,1] i64.  Problem p uses image p % B.
 * out: p_indices [P,Q] i64 (ascending query ids, first n valid, rest -1), t_indices [P,Q] i64
 *      (target matched to p_indices[k]), p_selector [P,Q] u8, match [P,Q] i32 (target of query q
 *      or -1), cost [P, Q, 100] fp32 (optional, may be NULL), status [P] i32 (0 ok, 1 NaN/-inf cost).
 * Bit-exact vs scipy.optimize.linear_sum_assignment on the same fp32 cost matrix, ties included.
 * ------------------------------------------------------------------------------------------ */
int detrb_matcher(const float *logits, int ldl, const float *boxes,
                  const float *t_bbox, const int64_t *t_class,
                  int P, int B, int Q, int C,
                  float fcost_class, float fcost_bbox, float fcost_giou,
                  int64_t *p_indices, int64_t *t_indices, uint8_t *p_selector, int32_t *match,
                  float *cost, int32_t *status, detrb_stream_t stream);

/* Set criterion (loss/loss.py:37-179) forward + analytic backward for L layers at once.
 *   sums [L, 8] fp32 scratch (zeroed by the call); losses [L, 6] fp32 in the order
 *   label_cost, true_neg, true_pos, pos_accuracy, giou_loss, l1_loss; total [1] fp32 =
 *   sum_l (1*label_cost + 2*giou_loss + 5*l1_loss) * loss_scale   (loss.py:6-19, training.py:20).
 *   normalisers: device float[2] = {n_matched, sum_w}, the batch-level normalisers (loss.py:66-67, 82, 94);
 *   pass the GLOBAL values under data parallelism; NULL: computed from this batch's targets.
 *   d_logits bf16 [P*Q, ld_dl] (cols >= C zeroed), d_boxpre bf16 [P*Q, ld_db]: gradient wrt the
 *   pre-sigmoid box head output (boxes = sigmoid(pre), detr.py:188), cols >= 4 zeroed.  NULL -> fwd only.
 *   status: the matcher's per-problem status [L*B] i32 or NULL.  The reference raises through scipy ("matrix contains invalid
 *   numeric entries") when a cost matrix holds NaN / -inf; a stream-ordered library cannot raise, so a non-zero status
 *   poisons the result instead: total and all 6*L loss scalars become NaN (visible at the caller's next read-back).
 *   split: plane stride of the bf16 pairs of d_logits / d_boxpre (parity precision; 0 = plain bf16). */
int detrb_set_loss(const float *logits, int ldl, const float *boxes,
                   const float *t_bbox, const int64_t *t_class, const int32_t *match,
                   int L, int B, int Q, int C, int background_class,
                   const float *normalisers, float loss_scale,
                   float *sums, float *losses, float *total,
                   detrb_bf16 *d_logits, int ld_dl, detrb_bf16 *d_boxpre, int ld_db,
                   const int32_t *status, int64_t split, detrb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer (optimizers.py:86-88,137-163): Keras Adam(beta1=.9, beta2=.999, eps=1e-7) with
 * PER-VARIABLE clipnorm, over a flat fp32 parameter arena described by a tensor table.
 *   table: T rows of {offset, numel} (int64 pairs, device); lr_group [T] i32 index into lrs [G] (device, fp32)
 *   so that learning rates can change without re-capturing a graph (training_config.py:66-68).
 *   group_enabled [8] u8 (config.train_<group>, optimizers.py:148), steps [8] i32: per-group Adam iteration
 *   counters (device), incremented by the call for enabled groups.  norms [T] fp32 scratch.
 * ------------------------------------------------------------------------------------------ */
int detrb_adam_clipnorm(float *params, const float *grads, float *m, float *v,
                        const int64_t *table, const int32_t *lr_group, const float *lrs,
                        const uint8_t *group_enabled, int T, int64_t total,
                        float clipnorm, float beta1, float beta2, float eps,
                        int32_t *steps, float *norms, detrb_stream_t stream);

/* Same step over uniform work items: chunks = DEVICE int32 [nchunks][3] = {table row, start offset in the arena (multiple of 4),
 * length <= 8192}; tensors are cut into chunks by the host once.  float4 accesses, balanced blocks.
 * One optimizer step may be issued as several calls over disjoint chunk ranges, each holding whole variables (the engine
 * applies everything but the stem kernel while the stem's weight gradient is still being computed): prologue != 0 on the
 * first call of the step (bumps the enabled groups' counters, zeroes all norms), 0 on the following ones. */
int detrb_adam_clipnorm_chunked(float *params, const float *grads, float *m, float *v, const int32_t *chunks, int nchunks,
                                const int32_t *lr_group, const float *lrs, const uint8_t *group_enabled, int T,
                                float clipnorm, float beta1, float beta2, float eps, int32_t *steps, float *norms,
                                int prologue, detrb_stream_t stream);

/* master fp32 weight [N, taps, Cin] (+ optional per-row fold[n]) -> bf16 forward copy Wf [N, ldf]
 * (K = taps*Cin, zero padded to ldf) and optional data-gradient copy Wd [Cin, taps, ldd] (cols>=N zero). */
int detrb_prep_weight(const float *master, const float *fold, int N, int taps, int Cin,
                      detrb_bf16 *Wf, int ldf, detrb_bf16 *Wd, int ldd, detrb_stream_t stream);

/* The same refresh for ALL weights in one launch.  descs: DEVICE array of nslots descriptors (device pointers inside);
 * tile_begin = exclusive prefix sum of taps * ceil(N/64) * ceil(Cin/64) over the slots; total_tiles = the full sum.
 * Cin, ldf, ldd even (packed bf16 pairs). */
typedef struct {
    const float *master; const float *fold;
    detrb_bf16 *Wf; detrb_bf16 *Wd;
    int N, taps, Cin, ldf, ldd, tile_begin;
} detrb_prep_desc_t;
/* wsplit != 0 (parity precision): Wf / Wd are written as bf16 pairs, the lo plane wsplit elements after the hi plane. */
int detrb_prep_weights_multi(const detrb_prep_desc_t *descs, int nslots, int total_tiles, int64_t wsplit, detrb_stream_t stream);

/* Gradient accumulation over micro-batches (optimizers.py:137-163, aggregate_grad_and_apply): acc[i] = (zero_first ? 0 :
 * acc[i]) + g[i] over n fp32 elements (n % 4 == 0, 16-byte aligned): the reference zeroes its accumulators at
 * step % k == 0 (:150-153) and adds every micro-step's gradient (:155-157). */
int detrb_accumulate(float *acc, const float *g, int64_t n, int zero_first, detrb_stream_t stream);

/* debug/test helper: writes the dropout keep-mask (0/1 bytes) the kernels use for an [M,N] site */
int detrb_dropout_mask(uint8_t *out, int M, int N, float drop_p, uint64_t seed, uint32_t site,
                       const uint64_t *seed_ptr, detrb_stream_t stream);

/* the same for an attention-probability dropout site (detrb_attn_fwd/bwd): M = B*H*Lq rows ((b*H + h)*Lq + q), N = Lk keys */
int detrb_attn_dropout_mask(uint8_t *out, int M, int N, float drop_p, uint64_t seed, uint32_t site,
                            const uint64_t *seed_ptr, detrb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Rows either side of the train step
 * ------------------------------------------------------------------------------------------ */
/* Input normalisation of uint8 frames (data/processing.py:6-23, normalized_images) as a table lookup:
 *   img [npix, 3] u8 (NHWC, any batch), lut [3*256] f32 DEVICE: lut[c*256 + v] = value of byte v in OUTPUT channel c
 *   (built by the host with the reference's float64 arithmetic -> results bit-identical to the reference);
 *   swap_rb != 0 reads input channel 2-c (normalized_method "tf_resnet": RGB -> BGR).  out [npix, 3] f32. */
int detrb_normalize_u8(const uint8_t *img, const float *lut, int swap_rb, float *out, int64_t npix, detrb_stream_t stream);
/* The same normalisation fused into the stem's input layout: img [B,H,W,3] u8 -> bf16 space-to-depth(2) tensor
 * [B, HP, WP, 16] (padding arguments as in detrb_image_to_s2d16); the fp32 image never exists in HBM. */
int detrb_image_u8_to_s2d16(const uint8_t *img, const float *lut, int swap_rb, detrb_bf16 *out, int B, int H, int W,
                            int pad_top, int pad_left, int HP, int WP, int64_t split, detrb_stream_t stream);
/* Inference post-process (inference.py:68-95, get_model_inference) for B images in one launch:
 *   logits [B,Q,C] f32 (row stride ldl), boxes [B,Q,4] f32 cxcywh (16-byte aligned).
 *   Per query: softmax, score = max probability, label = argmax of the softmax (first index on ties); queries whose label
 *   == background_class are dropped; the rest are compacted in ascending query order.
 *   bbox_format: 0 "xy_center" (as predicted), 1 "xyxy", 2 "yxyx" (corners clipped to [0,1], bbox.py:171-183).
 * out: out_boxes [B,Q,4] f32, out_labels [B,Q] i64, out_scores [B,Q] f32, out_query [B,Q] i32 (source query; may be NULL):
 *      first out_count[b] rows of image b valid.  Q <= 1024. */
int detrb_postprocess(const float *logits, int ldl, const float *boxes, int B, int Q, int C, int background_class,
                      int bbox_format, float *out_boxes, int64_t *out_labels, float *out_scores, int32_t *out_query,
                      int32_t *out_count, detrb_stream_t stream);

/* Input geometry on device (data/transformation.py:54-114, detr_aug_seq: Fliplr, Resize / CropToFixedSize / Affine scale and the
 * final Resize to config.image_size collapse into one axis-aligned affine map per image, sampled on the host together with the box
 * transform): B ragged uint8 RGB source frames -> one [B,H,W,3] uint8 batch, bilinear.
 *   src: all frames in one DEVICE buffer, frame b = [src_hw[2b], src_hw[2b+1], 3] bytes at src + src_off[b]
 *   inv [B,4] f32 = {ax, bx, ay, by}: source coordinates of output pixel centre (x+.5, y+.5): xs = ax*(x+.5) + bx - .5, ys likewise
 *   zero_border [B] u8: samples outside the source read 0 (imgaug's constant fill) / the nearest edge pixel (plain resize)
 * The result feeds the model as uint8 frames (normalisation fused into the stem's input layout: detrb_image_u8_to_s2d16). */
int detrb_resize_affine_u8(const uint8_t *src, const int64_t *src_off, const int32_t *src_hw, const float *inv,
                           const uint8_t *zero_border, uint8_t *out, int B, int H, int W, detrb_stream_t stream);

/* mAP matching (loss/compute_map.py:183-272, cal_map 'box' entries; driven per image by eval.py:38-52) for B images in one launch.
 *   detections: pred_boxes [B,Q,4] f32 in yxyx corners, pred_labels [B,Q] i64, pred_scores [B,Q] f32, the first pred_count[b] rows of
 *   image b valid (= the outputs of detrb_postprocess with bbox_format 2).  Q <= 256.
 *   ground truth: t_wire != 0: the padded wire format of data/processing.py:35-55 (t_boxes [B,NT,4] cxcywh with header row, t_labels
 *   [B,NT,1], NT = 100; converted to clipped yxyx corners like bbox.py:171-183 / :125-138); t_wire == 0: t_boxes [B,NT,4] yxyx rows,
 *   t_labels [B,NT], t_count [B].  At most 100 boxes per image.
 *   thresholds [T] f64 (DEVICE; the reference's python floats), T <= 32.
 * out: rank [B,Q] i32 = position of detection i in the stable descending-score order (-1 beyond pred_count); tp [B,T,Q] u8 = 1 when
 *      detection i is a true positive at threshold t: visited in that order it takes the unused ground-truth box of its class with the
 *      highest IoU strictly above the threshold (first maximum wins); IoU in the reference's fp32 operation order, compared in fp64.
 *      gt_count [num_classes] i32 (optional) += ground-truth boxes per class (add_gt_positives, :225). */
int detrb_map_match(const float *pred_boxes, const int64_t *pred_labels, const float *pred_scores, const int32_t *pred_count,
                    int B, int Q, const float *t_boxes, const int64_t *t_labels, const int32_t *t_count, int NT, int t_wire,
                    const double *thresholds, int T, int num_classes, int32_t *rank, uint8_t *tp, int32_t *gt_count,
                    detrb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
