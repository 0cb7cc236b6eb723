/*
 * csi_that.h -- C ABI of libcsi_that.so: the B200 (sm_100a) kernels behind the THAT train step.
 *
 * The reference (amirhosseinmhd/multi_modal_CSI) has no FFI of its own: its hot path is Python calling
 * torch.nn modules.  Each entry point below therefore names the reference call(s) it replaces, with paths
 * relative to the reference checkout (benchmark/wifi_csi/...).  The Python side of the boundary
 * (multi_modal_csi_b200/ops.py) binds these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success, a negative csi_status otherwise; csi_last_error() gives the text
 *     (thread local).  Launches are asynchronous on `stream` (a cudaStream_t passed as void*); nothing here
 *     allocates, frees or synchronises, so every call is CUDA-graph capturable.
 *   - all buffers are device pointers owned by the caller (PyTorch's allocator).
 *   - dtype arguments: CSI_F32 = 0, CSI_BF16 = 1.
 *   - "token buffer": a [B*Lp, ld] row-major matrix holding B samples of L tokens with `halo` zero rows in
 *     front of and behind every sample (Lp = L + 2*halo), ld >= Dp = round_up(d,16) and CSI_GUARD_ROWS
 *     readable rows before row 0 and after the last row.  Token (b,l) lives in row b*Lp + halo + l.  The
 *     halo/guard rows are what turn Conv1d(padding="same") into a GEMM over row-shifted views.
 *   - dropout: element idx of site `drop_site` is kept iff philox(rng[0], rng[1], drop_site, idx) >= p and
 *     then scaled by 1/(1-p).  `rng` points at two device uint64 {seed, step}; it is not read when p == 0.
 */
#ifndef CSI_THAT_H
#define CSI_THAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSI_F32 0
#define CSI_BF16 1
#define CSI_GUARD_ROWS 16
#define CSI_MAX_SEGS 32

typedef enum {
    CSI_OK = 0,
    CSI_ERR_ARG = -1,      /* bad argument (null pointer, unsupported size or dtype) */
    CSI_ERR_CUDA = -2,     /* a CUDA runtime / driver call failed */
    CSI_ERR_ARCH = -3      /* device is not sm_100 */
} csi_status;

/* One K-segment of a row-shifted GEMM: A columns [a_col_off, a_col_off+klen) of row (m + a_row_shift)
 * are contracted with B columns [b_col_off, b_col_off+klen).  klen must be a multiple of 16. */
typedef struct { int a_row_shift, a_col_off, b_col_off, klen; } csi_seg;

/* Output-column band of one segment of csi_gemm_nt_banded: the segment contributes to columns [n_lo, n_hi) only. */
typedef struct { int n_lo, n_hi; } csi_band;

/* One output segment of the weight-gradient GEMM (see csi_gemm_tn). */
typedef struct { int b_row_shift, b_col_off, c_off, nlen; } csi_seg_tn;

/* Up to three per-branch parameter pointers (the three Conv1d->BatchNorm1d branches of an Encoder). */
typedef struct { void* p[3]; } csi_ptr3;

/* Head padding of a feature index: features come in groups (attention heads) of `valid` entries that are stored with
 * a pitch of `pad` entries (27 -> 32, 15 -> 16, 54 -> 64) so that every head starts 16-byte aligned:
 *   padded(i) = (i / valid) * pad + i % valid        compact(i') = (i' / pad) * valid + i' % pad  (if i' % pad < valid)
 * {0, 0} means "no padding". */
typedef struct { int valid, pad; } csi_grp;

/* One tensor of the weight re-layout table (see csi_pack_weights). */
typedef struct {
    long long src_off;   /* element offset of the fp32 [N, C, k] source in the parameter arena        */
    long long dst_off;   /* element offset of the destination matrix in the packed arena            */
    int N, C, k;         /* source shape (k = 1 for Linear weights)                                  */
    int ld;              /* destination leading dimension                                           */
    int mode;            /* 0: dst[n*ld + j*P + c] = src[n,c,j]      (forward operand)               */
                         /* 1: dst[c*ld + (seg_base+j)*P + n] = src[n,c,j]  (data-gradient operand)  */
    int P;               /* padded channel pitch of one tap                                         */
    int seg_base;        /* first segment index of this tensor in mode 1                            */
    csi_grp gn;          /* head padding of the n index (n -> n'), {0,0} = identity                  */
    csi_grp gc;          /* head padding of the c index (c -> c')                                    */
    int reserved;
} csi_pack_entry;

const char* csi_last_error(void);
/* Returns the library ABI version (bumped on any signature change). */
int csi_abi_version(void);
/* Compute capability (major*10+minor) of `device`, or a negative csi_status. */
int csi_device_arch(int device);

/* ---- a1-a5: load_data.py:66-72 (front zero-pad), train.py:65-73 (apply_augmentation),
 *      that.py:257-259,279-280 (AvgPool1d(20,20) of both streams + permutes), that.py:88 (x + PE).
 * x: fp32 [B,T,F] dense when offs == NULL; otherwise sample b is lens[b] <= T rows of F floats starting at
 * element offs[b] of x and is FRONT-padded with zeros to T.  left/right: fp32 token buffers with
 * (L,d) = (T/20, F) and (F, T/20).  pe: fp32 [T/20, ld_pe] added to the left stream (NULL = none).
 * augment != 0 applies (x + 0.1*N(0,1)) * U[0.9,1.1)_b * Bernoulli(0.96) before pooling. */
int csi_pool_dual(const float* x, const long long* offs, const int* lens, int B, int T, int F,
                  const float* pe, int ld_pe, float* left, int ld_left, float* right, int ld_right,
                  int halo, int augment, const unsigned long long* rng, void* stream);

/* ---- a6: that.py:61-90 Gaussian_Position.  w: [L,K] softmax weights (saved for backward). */
int csi_gauss_pe_fwd(const float* pos, const float* mu, const float* sigma, const float* emb,
                     int L, int K, int F, float* w, float* pe, int ld_pe, void* stream);
/* dleft: fp32 token-buffer gradient of the PE output; dpe_ws: fp32 scratch [L, ld_ws].
 * demb [K,F], dmu [K], dsigma [K] are accumulated (caller zeroes). */
int csi_gauss_pe_bwd(const float* dleft, int ld_dleft, int B, int halo, const float* w, const float* pos,
                     const float* mu, const float* sigma, const float* emb, int L, int K, int F,
                     float* dpe_ws, int ld_ws, float* demb, float* dmu, float* dsigma, void* stream);

/* ---- a7: torch.nn.LayerNorm(d, eps=1e-6) at that.py:112,120,206,229.
 * x: fp32 token buffer; y: token buffer of y_dtype (halo rows and pad columns are written as zero). */
int csi_layernorm_fwd(const float* x, int ldx, const float* gamma, const float* beta, void* y, int ldy,
                      int y_dtype, float* mean, float* rstd, int B, int L, int d, int halo, float eps,
                      void* stream);
/* dx = LN'(dy) + dres (dres nullable); optional second output dxm = dropout_mask(dx) in dxm_dtype (the
 * gradient entering the out-projection, that.py:151).  dgamma/dbeta accumulated (caller zeroes). */
int csi_layernorm_bwd(const void* dy, int lddy, int dy_dtype, const float* x, int ldx, const float* gamma,
                      const float* mean, const float* rstd, const float* dres, int lddres, float* dx,
                      int lddx, void* dxm, int lddxm, int dxm_dtype, float drop_p, unsigned drop_site,
                      const unsigned long long* rng, float* dgamma, float* dbeta, int B, int L, int d,
                      int halo, void* stream);

/* ---- a8/a10/a11/a12 contractions: F.linear / Conv1d forward and data-gradient.
 * C[m, n] = sum_s sum_{q<klen_s} A[(m + shift_s)*lda + a_col_off_s + q] * Bw[n*ldb + b_col_off_s + q]
 *           (+ bias[n]) -> dropout(site) -> (+ residual[m*ldr + n]),   m < M, n < N.
 * A and Bw have ab_dtype; C has c_dtype; bias/residual are fp32 and nullable.
 * With ab_dtype == CSI_BF16 this is the tcgen05/TMA kernel; CSI_F32 is the FFMA "fp32 parity" kernel. */
int csi_gemm_nt(const void* A, int lda, const void* Bw, int ldb, int ab_dtype, void* C, int ldc, int c_dtype,
                int M, int N, const csi_seg* segs, int nseg, const float* bias, const float* residual,
                int ldr, float drop_p, unsigned drop_site, const unsigned long long* rng, void* stream);

/* The same contraction where segment s contributes to the output columns [bands[s].n_lo, bands[s].n_hi) only, i.e. the caller
 * guarantees Bw[n, b_col_off_s ...] == 0 for n outside the band.  This is how the three parallel Conv1d branches of an encoder
 * (kernel sizes 1/3/5 or 1/2/3 over the SAME input, that.py:122-135,158-162) run as ONE GEMM with N = 3*Dp: the weights of the
 * branches are stacked along N with their taps aligned on the largest kernel, the taps a branch does not have are zero blocks,
 * and the tcgen05 kernel skips them per column tile (column tiles are cut at the band boundaries).  bands == NULL is csi_gemm_nt.
 * The FFMA kernel ignores the bands (the zero blocks make the result identical). */
int csi_gemm_nt_banded(const void* A, int lda, const void* Bw, int ldb, int ab_dtype, void* C, int ldc, int c_dtype,
                       int M, int N, const csi_seg* segs, const csi_band* bands, int nseg, const float* bias,
                       const float* residual, int ldr, float drop_p, unsigned drop_site, const unsigned long long* rng,
                       void* stream);

/* ---- weight gradients (autograd of the same calls):
 * C[i*ldc + c_off_s + q*c_col_stride] += sum_{m<M} A[m*lda + i] * Bv[(m + b_row_shift_s)*ldb + b_col_off_s + q]
 * for i < Na, q < nlen_s.  C is fp32 and is accumulated with atomics (caller zeroes); with c_col_stride = k
 * it writes Conv1d weight gradients straight into the reference [N, C, k] layout. */
int csi_gemm_tn(const void* A, int lda, const void* Bv, int ldb, int ab_dtype, float* C, int ldc,
                int c_col_stride, int M, int Na, const csi_seg_tn* segs, int nseg, csi_grp i_grp, csi_grp q_grp,
                void* stream);
/* i_grp / q_grp: A's columns (i) / Bv's columns (q) are head-padded; C is indexed with the compact indices and the
 * padding entries are skipped (in-projection and out-projection weight gradients). */

/* Two-stage reduction for csi_gemm_tn (bf16): after this call every csi_gemm_tn issued on `stream` of the current device
 * stores the partial sums of its token chunks in `ws` and adds them to C with a second, fixed-order kernel instead of fp32
 * atomics from every CTA (faster, and the gradient is bit-reproducible).  nfloats >= SMs * 128 * 512 covers every shape;
 * a call that needs more falls back to atomics.  The caller owns `ws` (128-byte aligned) and must keep it alive and use it
 * for one stream only; ws = NULL unregisters.  Replaces nothing in the reference: torch.autograd reduces inside cuBLAS. */
int csi_gemm_tn_workspace(void* stream, float* ws, long long nfloats);

/* Column sums over the valid token rows: out[compact(c)] += sum_{b,l} A[row(b,l), c]  (bias gradients). */
int csi_colsum_tokens(const void* A, int lda, int dtype, int B, int L, int halo, int ncols, csi_grp grp, float* out,
                      void* stream);

/* ---- a8: nn.MultiheadAttention core at that.py:149 (per head softmax(q k^T / sqrt(hd)) v).
 * Head-padded layout: head h of q | k | v lives in columns [w*H*hp + h*hp, +hd) of the token buffer qkv (w = 0,1,2),
 * head h of o / dout in columns [h*hp, +hd); hd = d/H, hp >= hd is the head pitch (padding columns are zero on input
 * and written as zero).  lse: fp32 [B,H,L]. */
int csi_attn_fwd(const void* qkv, int ld3, void* o, int ldo, int dtype, float* lse, int B, int L, int d,
                 int H, int hp, int halo, void* stream);
int csi_attn_bwd(const void* qkv, int ld3, const void* o, int ldo, const void* dout, int lddo, void* dqkv,
                 int lddqkv, int dtype, const float* lse, int B, int L, int d, int H, int hp, int halo, float* dbias,
                 void* stream);
/* dbias (optional, may be NULL): fp32 [3*d], accumulated with the column sums of dq | dk | dv over the valid tokens in the
 * compact (un-padded) channel order -- the gradient of nn.MultiheadAttention.in_proj_bias, fused into the kernel. */

/* ---- a10: BatchNorm1d (train: batch statistics) -> Dropout(.1) -> LeakyReLU, mean of the 3 branches,
 * Dropout(.1), residual (that.py:126-132,160-168).  z: token buffer [rows, ldz] with branch br in columns
 * [br*Dp, br*Dp + d); it holds the convolution WITHOUT its bias (the bias only shifts the batch mean). */
int csi_bn_stats(const void* z, int ldz, int dtype, int B, int L, int halo, int ncols, double* sums,
                 void* stream);                                   /* sums: [2, ncols], accumulated */
int csi_bn_finalize(const double* sums, int Dp, int d, int nbr, long long count, csi_ptr3 conv_bias,
                    csi_ptr3 run_mean, csi_ptr3 run_var, csi_ptr3 num_batches, float momentum, float eps,
                    float* mean, float* invstd, void* stream);
/* eval mode: mean = running_mean - conv_bias, invstd = rsqrt(running_var + eps) */
int csi_bn_eval_prepare(int Dp, int d, int nbr, csi_ptr3 conv_bias, csi_ptr3 run_mean, csi_ptr3 run_var,
                        float eps, float* mean, float* invstd, void* stream);
int csi_bn_act_fwd(const void* z, int ldz, int dtype, const float* mean, const float* invstd, csi_ptr3 gamma,
                   csi_ptr3 beta, const float* t_res, int ldt, float* out, int ldo, int B, int L, int d,
                   int halo, int nbr, float p_branch, unsigned site_branch, float p_out, unsigned site_out,
                   const unsigned long long* rng, unsigned int* masks, void* stream);
/* masks (optional, may be NULL): one 32-bit word per (token row, 8-channel group), index row*(Dp/8) + group, holding
 * the dropout KEEP bits of that group: byte br = branch br, byte 3 = the output dropout.  bn_act_fwd writes it, the two
 * backward kernels read it instead of regenerating the Philox decisions (identical bits either way).
 * red: [2, nbr*Dp] doubles (sum dy, sum dy*zhat), accumulated */
int csi_bn_act_bwd_reduce(const float* dout, int lddo, const void* z, int ldz, int dtype, const float* mean,
                          const float* invstd, csi_ptr3 gamma, csi_ptr3 beta, int B, int L, int d, int halo,
                          int nbr, float p_branch, unsigned site_branch, float p_out, unsigned site_out,
                          const unsigned long long* rng, const unsigned int* masks, double* red, void* stream);
int csi_bn_act_bwd_dz(const float* dout, int lddo, const void* z, int ldz, int dtype, const float* mean,
                      const float* invstd, csi_ptr3 gamma, csi_ptr3 beta, const double* red, int B, int L,
                      int d, int halo, int nbr, float p_branch, unsigned site_branch, float p_out,
                      unsigned site_out, const unsigned long long* rng, const unsigned int* masks, void* dz, int lddz,
                      csi_ptr3 dgamma, csi_ptr3 dbeta, void* stream);

/* ---- a11: head Conv1d (valid) + LeakyReLU + sum over time (that.py:268-272,287-291).
 * p: token buffer [rows, ldp] of conv pre-activations; columns [0,n0) come from a kernel of size k0 and
 * [n0,N) from size k1, so token t contributes iff t <= L - k.  feat: fp32 [B, ldf]. */
int csi_head_reduce_fwd(const void* p, int ldp, int dtype, int B, int L, int halo, int N, int n0, int k0,
                        int k1, float* feat, int ldf, void* stream);
int csi_head_reduce_bwd(const float* dfeat, int ldf, const void* p, int ldp, int dtype, int B, int L, int halo,
                        int N, int n0, int k0, int k1, void* dp, int lddp, void* stream);

/* ---- a12: Dropout(0.5) on the concatenated features (that.py:274-275,293-294); out = mask(in). */
int csi_dropout_rows(const float* in, int ldi, void* out, int ldo, int out_dtype, int rows, int cols, float p,
                     unsigned site, const unsigned long long* rng, void* stream);

/* ---- a13: BCEWithLogitsLoss(pos_weight) mean (that.py:401, train.py:97) and its gradient * grad_scale. */
int csi_bce_logits(const float* z, int ldz, const float* y, int ldy, int rows, int cols, float pos_weight,
                   float grad_scale, float* loss, float* dz, int lddz, void* stream);

/* ---- sibling head: torch.nn.SmoothL1Loss(beta) mean of THAT_COUNT_PRED (that_count_pred.py:399) and its
 * gradient * grad_scale; same calling convention as csi_bce_logits. */
int csi_smooth_l1(const float* z, int ldz, const float* y, int ldy, int rows, int cols, float beta,
                  float grad_scale, float* loss, float* dz, int lddz, void* stream);

/* ---- sibling head: PermutationMatchingLoss of the five-head THAT (model/that_multi_head.py:309-342).
 * z: fp32 [B, ldz], head h's logits at columns [h*cpitch, h*cpitch + classes); y: fp32 [B, ldy] = [B, heads, classes]
 * targets (the class of slot t is argmax(y[b, t])).  Per sample the permutation of the heads with the smallest mean
 * cross-entropy (first minimum in itertools.permutations order) is chosen; loss[0] = mean CE over B*heads of the matched
 * heads; dz (optional, [B, lddz], pad columns written as 0) = its gradient * grad_scale; best_perm (optional, int32
 * [B, heads]) = head matched to each slot. */
int csi_perm_ce(const float* z, int ldz, const float* y, int ldy, int B, int heads, int classes, int cpitch,
                float grad_scale, float* loss, float* dz, int lddz, int* best_perm, void* stream);

/* ---- a15: torch.optim.Adam with coupled L2 (that.py:395-397) over the flat arenas.
 * step: device int64 holding the 1-based step of THIS update.  g is multiplied by grad_scale first. */
int csi_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, const long long* step, float grad_scale, void* stream);
/* rng[1] += 1; step[0] += 1 (one tiny launch so a captured graph advances its own counters) */
int csi_advance_counters(unsigned long long* rng, long long* step, void* stream);

/* ---- per-step re-layout of the fp32 master weights into the GEMM operand copies (table on device). */
int csi_pack_weights(const float* params, void* packed, int dtype, const csi_pack_entry* table, int n_entries,
                     int max_elems, void* stream);

/* Test hook: 1 = route every contraction/attention call to the FFMA kernels, 0 = tensor-core kernels where eligible. */
int csi_set_force_simt(int on);
/* bf16 dispatch is never silent: every bf16 contraction / attention call is counted by the kernel class that served it
 * (tcgen05 / FFMA fallback for a shape the tcgen05 kernel cannot take / mma.sync attention).  In strict mode (on = 1:
 * bench.py, the full-size tests) a bf16 call that would fall back to FFMA returns CSI_ERR_ARG instead. */
int csi_set_strict_tc(int on);
int csi_dispatch_counts(long long* tcgen05_calls, long long* ffma_fallback_calls, long long* mma_sync_calls, int reset);
/* Attention core implementation per direction: 0 = measured-best per shape (default), 1 = tcgen05/TMEM kernels wherever
 * eligible, 2 = mma.sync kernels only (dispatch.cu holds the measurements behind mode 0). */
int csi_set_attn_impl(int fwd_mode, int bwd_mode);

/* ---- a16: the reference decision rule (utils.py:147-183,234-239): sigmoid -> per user the arg-max class counts iff
 * its probability exceeds `threshold` (the reference hard-codes 0.5) -> per-class counts.  logits: fp32 [rows, ldz]
 * with users*classes valid columns; counts: int32 [rows, classes]. */
int csi_predict_counts(const float* logits, int ldz, int rows, int users, int classes, float threshold, int* counts,
                       void* stream);

/* Generic helpers */
int csi_fill_f32(float* p, long long n, float v, void* stream);
int csi_fill_f64(double* p, long long n, double v, void* stream);
int csi_copy_f32(float* dst, const float* src, long long n, void* stream);

/* ---- (f)3 CSI-as-image path: model/cnn_2d.py:23-99 (CNN_2D: BatchNorm2d -> Conv2d -> LeakyReLU -> Dropout(0.2) three
 * times, BatchNorm2d, mean over the image, Linear).  Activations are NHWC matrices [B*H*W, C] (C = 1 for the raw fp32 CSI
 * image, then 32 / 64 / 128, stored in the activation dtype).  A Conv2d(k, stride s) is csi_im2col_bn (the preceding
 * BatchNorm2d applied on the fly) + csi_gemm_nt; its weight gradient csi_gemm_tn over (dZ, patches); its data gradient
 * csi_gemm_nt against the transposed weights + csi_col2im. */
/* load_data.py:66-72 + train.py:65-73 for the image path: the dense FRONT-padded batch out [B,T,F] gathered from a packed
 * arena (offs / lens as in csi_pool_dual; NULL = x is dense already), optionally augmented
 * ((x + 0.1 N(0,1)) * U[0.9,1.1)_b * Bernoulli(0.96)). */
int csi_gather_aug(const float* x, const long long* offs, const int* lens, int B, int T, int F, float* out, int augment,
                   const unsigned long long* rng, void* stream);
/* cnn_2d.py:75,80,85,90 BatchNorm2d statistics: sums[c] += sum x, sums[C + c] += sum x^2 over x [rows, C] contiguous
 * (C = 1, or a multiple of 8 dividing 2048). */
int csi_nhwc_stats(const void* x, int dtype, long long rows, int C, double* sums, void* stream);
/* train (training != 0): batch statistics from sums, running statistics / num_batches_tracked updated (momentum 0.1,
 * unbiased variance); eval: running statistics.  scale = gamma*invstd, shift = beta - mean*scale. */
int csi_bn2d_finalize(const double* sums, int C, long long count, const float* gamma, const float* beta, float* run_mean,
                      float* run_var, long long* nbt, float momentum, float eps, int training, float* mean, float* invstd,
                      float* scale, float* shift, void* stream);
/* cnn_2d.py:76,81,86 patch matrix of Conv2d(k, stride s, no padding) over x [B,H,W,C] with BatchNorm applied:
 * col[(b,oh,ow), (kh*k+kw)*C + c] = x[b, oh*s+kh, ow*s+kw, c]*scale[c] + shift[c]; columns [k*k*C, Kp) are zero. */
int csi_im2col_bn(const void* x, int x_dtype, int B, int H, int W, int C, int k, int s, const float* scale, const float* shift,
                  void* col, int col_dtype, int Kp, void* stream);
/* data gradient: g[b,h,w,c] (fp32) = sum of gcol over the patches covering the pixel (gather, no atomics). */
int csi_col2im(const void* gcol, int dtype, int B, int H, int W, int C, int k, int s, int Kp, float* g, void* stream);
/* cnn_2d.py:77-78: y = Dropout_p(LeakyReLU(z)) over n elements; the keep bits of every 8 elements are stored (1 byte). */
int csi_act_drop_fwd(const void* z, void* y, int dtype, long long n, float p, unsigned site, const unsigned long long* rng,
                     unsigned char* mask, void* stream);
/* BatchNorm2d backward, pass 1: sums[c] += sum g', sums[C+c] += sum g'*(x-mean)*invstd with g' = g_scale * g (row r of g is
 * row r / g_div: the gradient of the spatial mean, cnn_2d.py:91, is shared by the g_div positions it averages). */
int csi_bn2d_bwd_reduce(const void* g, int g_dtype, long long g_div, float g_scale, const void* x, int x_dtype, long long rows, int C,
                        const float* mean, const float* invstd, double* sums, void* stream);
/* pass 2, fused with the Dropout + LeakyReLU backward of the block that produced x = Dropout(LeakyReLU(zprev)):
 * gz = gamma*invstd*(g' - sum_g/n - xhat*sum_gxhat/n) * keep/(1-p) * leaky'(zprev); dgamma += sum_gxhat, dbeta += sum_g. */
int csi_bn2d_bwd_apply(const float* g, long long g_div, float g_scale, const void* x, const void* zprev, int dtype, const unsigned char* mask,
                       float drop_p, long long rows, int C, const float* mean, const float* invstd, const float* gamma,
                       const double* sums, void* gz, float* dgamma, float* dbeta, void* stream);
/* cnn_2d.py:90-91: feat[b,c] = scale[c] * mean_p y[b,p,c] + shift[c] (fp32) and the same in the activation dtype. */
int csi_pool_bn_fwd(const void* y, int dtype, int B, int P, int C, const float* scale, const float* shift, float* feat,
                    void* featd, void* stream);
/* Conv2d / Linear weights: reference layout [N,C,k,k] fp32 -> forward operand wf [N,Kp] ((kh,kw,c) order) and
 * data-gradient operand wb [Kp,Np] (may be NULL); and the weight gradient back: gw[n,c,kh,kw] += gs[n,(kh*k+kw)*C+c]. */
int csi_conv2d_pack(const float* w, int N, int C, int k, void* wf, int Kp, void* wb, int Np, int dtype, void* stream);
int csi_conv2d_unpack_grad(const float* gs, int N, int C, int k, int Kp, float* gw, void* stream);
/* Affine gradients of the single-channel BatchNorm2d in front of conv 0 and the conv-0 bias gradient from the column sums
 * sums = [sum_m gz0, sum_m gz0*z0] (cnn2d.cu explains the identity). */
int csi_bn0_grads(const double* sums, const void* wf, int ldw, int dtype, int N, int K, const float* bias, const float* gamma0,
                  const float* beta0, float* dgamma0, float* dbeta0, float* dbias, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSI_THAT_H */
