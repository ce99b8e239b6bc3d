/* byol_b200.h — C ABI of libbyol_b200.so: the sm_100a kernels behind the BYOL training-step hot path.
 *
 * The reference (jramapuram/BYOL) is pure Python and has no FFI / operator registry of its own; its hot path
 * reaches cuDNN / cuBLAS / ATen / NCCL through torch.  This header is therefore the boundary a maintainer binds
 * from Python (ctypes stub in byol_b200/_lib.py): plain pointers and sizes, one CUDA stream,
 * no torch types, no allocation, no implicit synchronisation, no global mutable state.
 *
 * Conventions
 *   - every function returns 0 on success, < 0 on error; byol_last_error() gives the thread-local message
 *   - all pointers are device pointers unless stated otherwise; `stream` is the caller's cudaStream_t
 *   - activations are NHWC bf16 with the channel count padded to a multiple of 8; "[M, C]" means M = N*H*W rows
 *   - fp32 parameter / gradient tensors use the reference's layouts ([Cout, Cin, KH, KW], [out, in], [C])
 *   - launches are asynchronous; errors detected at launch time are reported, device faults surface at the
 *     caller's next synchronisation
 *
 * Reference lines cited below are relative to /root/reference (jramapuram/BYOL @ 5ea487e).
 */
#ifndef BYOL_B200_H
#define BYOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* byol_stream_t; /* == cudaStream_t */

const char* byol_last_error(void);
int byol_abi_version(void);
int byol_device_sm_count(void);

/* ---- tensor-core convolution / linear: replaces the cuDNN conv fwd/dgrad and cuBLAS Linear calls under
 *      main.py:229-240 (BYOL.prediction -> base_network, head, predictor) and main.py:250-252 (classifier) ----
 * out[M = Nimg*Ho*Wo, Ndim] = gather(src) x Wt^T (+bias) (+resid) (relu); optional fused per-column
 * sum / sum-of-squares of the stored values (BatchNorm statistics).
 *   mode 0 (fprop): src = x [Nimg,Hs,Ws,C], src coordinate = o*stride - pad + k
 *   mode 1 (dgrad): src = dY [Nimg,Hs,Ws,C=Cout], (Ho,Wo) = spatial size of dX, coordinate = (o + pad - k)/stride
 * wt: bf16 [Ndim, ldw] K-major with k = (kh*KW + kw)*C + c (made by byol_prep_weight).
 * resid_mask (optional, uint8 [M*ldc/8], written by byol_bn_apply): resid is added only where its bit is set, i.e.
 * the epilogue applies the ReLU mask of the residual branch (block backward, torchvision resnet.py Bottleneck.forward).
 * resid_up = 1: resid is the COMPACT [Nimg, Ho/2, Wo/2, ldc] gradient of a stride-2 1x1 branch and is added to the
 * even (h, w) pixels only (the dense, 3/4-zero gradient map of the downsample branch is never written). */
int byol_conv_igemm(const void* src, const void* wt, void* dst, const void* resid, const void* resid_mask,
                    int resid_up, const float* bias, float* col_sum, float* col_sqsum, int Nimg, int Hs, int Ws, int C, int Ho, int Wo, int Ndim,
                    int KH, int KW, int stride, int pad, int mode, int ldw, int ldc, int out_fp32, int relu,
                    int force_gather, byol_stream_t stream);

/* dW[Cout][Cin_real][KH][KW] (fp32, accumulated) += dY^T x im2col(x): replaces cuDNN wgrad / cuBLAS in the
 * autograd of main.py:617 (loss.backward()). */
int byol_conv_wgrad(const void* src, const void* dy, float* dw, int Nimg, int Hs, int Ws, int C, int Cin_real,
                    int Ho, int Wo, int Cout, int ldy /* row pitch of dy, 0 = Cout */, int KH, int KW, int stride,
                    int pad, int force_gather, byol_stream_t stream);

/* 1x1 / stride-1 convolution (plain GEMM out[M, Ndim] = src[M, C] x wt[Ndim, C]^T) with a fused BatchNorm epilogue:
 *   t = acc * colscale[n] + bias[n];  out = act(t + resid_colscale[n] * (resid_mask ? resid : 0)) -> dst, mask_out.
 * no_store = 1: only col_sum / col_sqsum (BatchNorm statistics) are produced — pass 1 of "statistics pass +
 * recompute": the block-output BatchNorm of a bottleneck (torchvision Bottleneck.bn3 + residual + ReLU, reached from
 * main.py:237) never materialises the raw conv output.  bwd_reduce = 1: the BatchNorm-backward sums of the recomputed
 * output: with dz = (resid_mask ? resid : 0): col_sum += sum dz, col_sqsum += sum dz * t. */
int byol_conv_igemm_fused(const void* src, const void* wt, void* dst, const void* resid, const void* resid_mask,
                          const float* colscale, const float* bias, const float* resid_colscale, void* mask_out,
                          float* col_sum, float* col_sqsum, int M, int C, int Ndim, int ldw, int ldc, int relu,
                          int no_store, int bwd_reduce, byol_stream_t stream);

/* ---- stem (7x7 / stride 2 / pad 3, <= 4 input channels, 64 output channels, W <= 256): the torchvision ResNet
 *      conv1 reached from main.py:237.  The image is converted once to zero-padded NHWC4 bf16
 *      ([N][H+6][264][4]); the kernel forms the im2col rows with overlapping no-swizzle UMMA descriptors. ---- */
int byol_stem4_supported(int Cin, int Cout, int H, int W, int k, int stride, int pad);
int byol_stem4_row_pixels(void); /* Wp of the padded image tensor [N][H+6][Wp][4] */
int byol_nchw_to_stem4(const float* x, void* xs, int N, int Cin, int H, int W, byol_stream_t stream);
/* w fp32 [64][Cin][7][7] -> ws bf16 [7][4][64][8] (14336 elements) */
int byol_prep_weight_stem4(const float* w, void* ws, int Cin, byol_stream_t stream);
/* y [N, H/2, W/2, 64] bf16; col_sum / col_sqsum (both or none): += per-channel sum / sum of squares of y */
int byol_stem_conv_fprop(const void* xs, const void* ws, void* y, float* col_sum, float* col_sqsum, int N, int H, int W,
                         byol_stream_t stream);

/* dw [64][Cin][7][7] fp32 += dY^T x im2col(xs); dy [N, H/2, W/2, 64] bf16 (autograd of main.py:617 for the stem) */
int byol_stem_conv_wgrad(const void* xs, const void* dy, float* dw, int N, int Cin, int H, int W, byol_stream_t stream);

/* ---- BatchNorm (train / eval, optionally cross-rank): replaces ATen batch_norm and SyncBatchNorm
 *      (main.py:196,202,237,433; torch/nn/modules/_functions.py:10-205) ---- */
int byol_bn_stats(const void* x, float* stats /* zeroed [2C] */, int M, int C, byol_stream_t stream);
/* statistics -> coefficients for L <= 4 lock-step lanes in one launch: stats [L][2C], coeffs [L][4][C] = scale, shift, mean, invstd;
 * running statistics are updated lane after lane (the order of the reference's four forward passes) */
int byol_bn_finalize_lanes(const float* stats, double count, int L, const float* gamma0, const float* beta0,
                           const float* gamma1, const float* beta1, const float* gamma2, const float* beta2,
                           const float* gamma3, const float* beta3, float* running_mean, float* running_var,
                           float momentum, float eps, float* coeffs, int C, byol_stream_t stream);
int byol_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                        float eps, float* scale, float* shift, int C, byol_stream_t stream);
/* y = act(x*scale + shift + (resid | resid*rscale + rshift)); mask_out (optional, uint8 [M*C/8]): bit e of byte i =
 * (y[8*i + e] > 0), the ReLU mask the backward kernels read instead of the activation (mask_mode 3) */
int byol_bn_apply(const void* x, const float* scale, const float* shift, const void* resid, const float* rscale,
                  const float* rshift, void* y, float* y_f32, void* mask_out, int M, int C, int relu,
                  byol_stream_t stream);
/* s12 (zeroed [2C]) += [sum dz, sum dz*xhat]; mask_mode 0 none / 1 relu(x*scale+shift) / 2 act > 0 (act = bf16
 * activation) / 3 mask bits (act = uint8 mask written by byol_bn_apply) */
int byol_bn_bwd_reduce(const void* g, const void* x, const void* act, const float* scale, const float* shift,
                       const float* mean, const float* invstd, float* s12, int M, int C, int mask_mode,
                       byol_stream_t stream);
/* dy = gamma*invstd*(dz - s1/n - xhat*s2/n); dgamma/dbeta (optional) += rank-local sums */
int byol_bn_bwd_apply(const void* g, const void* x, const void* act, const float* scale, const float* shift,
                      const float* mean, const float* invstd, const float* gamma, const float* s12, double count,
                      void* dy, void* dz_out, int M, int C, int mask_mode, const float* s12_local, float* dgamma,
                      float* dbeta, byol_stream_t stream);
/* per-channel vectors for the BatchNorm backward of a recomputed 1x1 convolution (byol_conv_igemm_fused) */
int byol_bn_bwd_prep(const float* mean, const float* invstd, float* out /* [2C] */, int C, byol_stream_t stream);
int byol_bn_bwd_coeffs(const float* s12, const float* s12_local, const float* mean, const float* invstd,
                       const float* gamma, double count, float* out /* [3C] */, float* dgamma, float* dbeta, int C,
                       byol_stream_t stream);
int byol_col_sum(const void* x, float* out, int M, int C, int ld, int is_f32, byol_stream_t stream);

/* ---- layout / pooling: torchvision ResNet stem and tail reached from main.py:237 ---- */
int byol_nchw_to_nhwc8(const float* x, void* y, int N, int Cin, int H, int W, byol_stream_t stream);
int byol_prep_weight(const float* w, void* w_fprop, void* w_dgrad, int Cout, int Cin, int Cpad, int KH, int KW,
                     byol_stream_t stream);
/* stem layout ([Cout][KH*64], column = kh*64 + kw*8 + c) selected by byol_conv_igemm when C == 8 and ldw == KH*64 */
int byol_prep_weight_fold(const float* w, void* w_fprop, int Cout, int Cin, int KH, int KW, byol_stream_t stream);
/* every conv / linear weight of one parameter set in one launch; desc: device int64 [num_units][8] =
 * {src offset in flat, fprop offset in pool_f, dgrad offset in pool_d or -1, Cout, Cin, Cpad, taps, fold (KH*16+KW or 0)} */
int byol_prep_unit_blocks(int Cout, int Cin, int Cpad, int taps, int fold);   /* blocks one unit needs */
int byol_prep_weights_multi(const float* flat, void* pool_f, void* pool_d, const int64_t* desc, int num_units,
                            int num_blocks /* sum of byol_prep_unit_blocks over the units */, byol_stream_t stream);
/* y[n,i,j,:] = x[n,2i,2j,:] (input of a 1x1 / stride-2 downsample conv, compacted for the TMA-fed GEMM) */
int byol_subsample2(const void* x, void* y, int N, int H, int W, int C, byol_stream_t stream);
int byol_cast_f32_bf16(const float* x, void* y, int64_t n, byol_stream_t stream);
/* y[r, c] = bf16(x[r, c]) for c < cols and 0 for cols <= c < ldy (pitched copy, e.g. classifier gradients) */
int byol_cast_f32_bf16_2d(const float* x, void* y, int rows, int cols, int ldx, int ldy, byol_stream_t stream);
int byol_maxpool_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, int k, int s, int p,
                     byol_stream_t stream);
/* stem fusion: y = maxpool(relu(x*scale + shift)); values and argmax indices equal bn_apply + maxpool_fwd exactly */
int byol_bn_relu_maxpool_fwd(const void* x, const float* scale, const float* shift, void* y, void* idx, int N, int H,
                             int W, int C, int k, int s, int p, byol_stream_t stream);
int byol_maxpool_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, int k, int s, int p,
                     byol_stream_t stream);
int byol_avgpool_fwd(const void* x, float* y_f32, void* y_bf16, int N, int HW, int C, byol_stream_t stream);
int byol_avgpool_bwd(const void* g_bf16, const float* g_f32, void* dx, int N, int HW, int C, byol_stream_t stream);

/* ---- objective: replaces objective.py:6-25 (regression_loss / loss_function; Frobenius-normalised, rank-local) ----
 * workspace: 6 doubles; loss: 1 float; saved: 6 floats consumed by byol_loss_bwd. */
int byol_loss_fwd(const float* q1, const float* q2, const float* z1, const float* z2, int rows, int dim,
                  double* workspace, float* loss, float* saved, byol_stream_t stream);
int byol_loss_bwd(const float* q1, const float* q2, const float* z1, const float* z2, const float* saved,
                  const float* grad_out, float* dq1, float* dq2, int rows, int dim, byol_stream_t stream);

/* ---- target network: replaces CosEMA.forward, main.py:159-162.  mean = fl(fl(a*x) + fl(d*mean)), bit-exact ---- */
int byol_ema_update(const float* x, float* mean, float one_minus_decay, float decay, int64_t n,
                    byol_stream_t stream);

/* ---- optimizer: replaces LARS.apply_adaptive_lrs + SGD(momentum).step, optimizers/lars.py:84-127 ----
 * p_ptrs / g_ptrs / m_ptrs: device arrays of num_tensors fp32 pointers (m_ptrs may be NULL: no momentum);
 * chunk tables split the tensors into work items (chunks of one tensor are contiguous:
 * [tensor_first_chunk[t], tensor_first_chunk[t+1])); partial: 2*num_chunks doubles of scratch.  No atomics: the
 * per-tensor norms are bit-reproducible, so data-parallel replicas stay bit-identical (main.py:440). */
int byol_lars_sgd_step(const void* p_ptrs, const void* g_ptrs, const void* m_ptrs, const int64_t* chunk_start,
                       const int* chunk_len, const int* chunk_tensor, int num_chunks, const int* tensor_first_chunk,
                       const float* wd, const float* lr, const int* ignore, int num_tensors, double* partial,
                       float trust_coef, float eps, float momentum, int first_step, byol_stream_t stream);

/* ---- linear-probe objective: replaces F.cross_entropy + helpers.metrics.topk, main.py:596-598 ----
 * logits fp32 [R, C] (row pitch ld), labels int64 [label_rows] (row r uses labels[r % label_rows]: the two views
 * of a sample share its label, main.py:591); scratch: row_lse / row_loss [R] floats, row_rank [R] ints,
 * ticket: one zeroed uint32.  out = [mean loss, top-1 %, top-5 %] (deterministic row-order reduction). */
int byol_ce_topk_fwd(const float* logits, const int64_t* labels, int label_rows, int R, int C, int ld, float* row_lse,
                     float* row_loss, int* row_rank, unsigned int* ticket, float* out, byol_stream_t stream);
/* dlogits[r, c] = grad_out / R * (softmax(logits[r])[c] - [c == labels[r]]) */
int byol_ce_bwd(const float* logits, const int64_t* labels, int label_rows, const float* row_lse, const float* grad_out, int R, int C,
                int ld, float* dlogits, int ldd, byol_stream_t stream);

/* ---- fp32-accurate forward path ("split-bf16", BASELINE.json configs[1]): the reference runs main.py:229-276 in
 *      fp32; here every fp32 operand is split exactly into T bf16 planes (T = 3: ~16 bits, T = 6: 24 bits) that feed
 *      the SAME tensor-core kernels as one GEMM over T*C channels (byol_conv_igemm with C := T*C, out_fp32 = 1);
 *      BatchNorm statistics are accumulated in fp64.  See csrc/split.cu. ---- */
int byol_split_planes(const float* x /* [M, C], pitch ldx */, void* planes /* bf16 [M, T*Cpad] */,
                      void* copy_bf16 /* optional [M, C] */, int64_t M, int C, int Cpad, int ldx, int T,
                      byol_stream_t stream);
int byol_nchw_to_planes(const float* x /* NCHW */, void* planes /* bf16 NHWC [N,H,W,T*Cpad] */, int N, int Cin, int H,
                        int W, int Cpad, int T, byol_stream_t stream);
int byol_prep_weight_planes(const float* w /* [Cout, Cin, taps] */, void* out /* bf16 [Cout, taps*T*Cpad] */, int Cout,
                            int Cin, int Cpad, int taps, int T, byol_stream_t stream);
int byol_stats_f32(const float* y, double* stats /* zeroed [2C] */, int64_t M, int C, byol_stream_t stream);
int byol_bn_finalize_lanes_f64(const double* stats, double count, int L, const float* gamma0, const float* beta0,
                               const float* gamma1, const float* beta1, const float* gamma2, const float* beta2,
                               const float* gamma3, const float* beta3, float* running_mean, float* running_var,
                               float momentum, float eps, float* coeffs, int C, byol_stream_t stream);
int byol_bn_apply_f32(const float* y, const float* scale, const float* shift, const float* resid, const float* rscale,
                      const float* rshift, float* out32, void* planes, void* copy_bf16, void* mask, int64_t M, int C,
                      int relu, int T, byol_stream_t stream);
int byol_maxpool_f32(const float* x, float* y, void* idx, int N, int H, int W, int C, int k, int s, int p,
                     byol_stream_t stream);
int byol_avgpool_f32(const float* x, float* y, int N, int HW, int C, byol_stream_t stream);

/* ---- projector / predictor MLP forward as ONE cooperative kernel (main.py:194-205, 238-239):
 *      Linear -> BatchNorm1d (batch statistics, grid barrier, optional cross-rank exchange) -> ReLU -> Linear.
 *      x [B, K1] bf16, w1 [H, ldw1], w2 [O, ldw2] bf16 (fprop layouts); stats [2H] and out [B, O] fp32 ZEROED by the
 *      caller; coeffs [4, H]; h_save / a_save optional bf16 [B, H] (for the backward pass); grid_bar: 2 zeroed uint32.
 *      See csrc/mlp_fused.cu. ---- */
int byol_mlp_fused_supported(int B, int K1, int H, int O);
int byol_mlp_fused_fwd(const void* x, const void* w1, const float* b1, const float* gamma, const float* beta,
                       const void* w2, const float* b2, float* stats, float* running_mean, float* running_var,
                       float momentum, float eps, double count, float* coeffs, float* out, void* h_save, void* a_save,
                       void* grid_bar, int B, int K1, int H, int O, int ldw1, int ldw2, int train,
                       const uint64_t* peer_ptrs, int world, int rank, int64_t cap_bytes, void* counter,
                       byol_stream_t stream);

/* ---- SyncBatchNorm statistic exchange over NVLink peer memory (main.py:433; replaces the per-layer all_gather /
 *      all_reduce of torch/nn/modules/_functions.py:49-74,158-159): one single-CTA kernel per exchange — publish into
 *      this rank's symmetric buffer, flag every peer, wait, add all peers' values in rank order.  See csrc/xchg.cu. */
int byol_xchg_layout(int* slots, int* max_world, int* flag_bytes);
int byol_xchg_sum(void* vals, void* local_copy, int n, int is_f64, const uint64_t* peer_ptrs /* host, [world] */,
                  int world, int rank, int64_t cap_bytes, void* counter /* device uint32 */, byol_stream_t stream);

/* ---- on-device two-view augmentation (main.py:386-397: RandomResizedCrop, flip, ColorJitter p = 0.8, grayscale
 *      p = 0.2, Gaussian blur p = 0.5) for decoded fp32 NCHW images resident in HBM.  params: fp32 [2, N, 16] records
 *      (view-major; byol_augment_record_floats() floats each: crop top / left / h / w, flip, jitter on, op order x 4,
 *      brightness, contrast, saturation, hue, gray on, blur sigma).  See csrc/augment.cu. ---- */
int byol_augment_record_floats(void);
int byol_augment_params(float* params, int N, int Hs, int Ws, uint64_t seed, uint64_t step, float strength,
                        float p_flip, float p_jitter, float p_gray, float p_blur, byol_stream_t stream);
int byol_augment_apply(const float* src, const float* params, float* out /* [2, N, 3, R, R] */, float* tmp,
                       double* gray_sum /* [2N] */, int N, int Hs, int Ws, int R, int ksize, byol_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BYOL_B200_H */
