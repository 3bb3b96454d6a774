/*
 * vtc_b200.h -- C ABI of the B200-native sparse-coding hot path.
 *
 * The reference (spencerkent/vision-transform-codes) has no FFI: its plug-in interface is "a Python module with a
 * run(...) attribute" (vision_transform_codes/training/sparse_coding.py:389-439, called at :139 and :168). Each
 * entry point below is what the drop-in module of the same dotted name binds through ctypes; the reference function
 * it replaces is cited per entry point. INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *  - all matrix pointers are DEVICE pointers to row-major float32 (the reference's layout: images (b, n),
 *    dictionary (s, n) with atoms as rows, codes (b, s));  ld_* are row pitches in elements;
 *  - the caller owns every buffer, including the workspace (size from the *_workspace_bytes query); the library
 *    never allocates or frees device memory;
 *  - every call is enqueued on `stream` and returns without synchronising, except where stated;
 *  - return value 0 = success; anything else is an error, text via vtc_last_error();
 *  - there is no CPU fallback: without an sm_100 device every compute entry point fails with VTC_ERR_CUDA;
 *  - threading: compute calls may be issued concurrently on DIFFERENT streams with different workspaces (the error
 *    text is per thread). The tuning switches (vtc_set_formulation, vtc_set_fused_iteration,
 *    vtc_set_small_batch_kernel and the VTC_B200_* environment variables), the profile recorder (vtc_profile_*),
 *    the launch counter and the debug trace hook are process-global and not synchronised: set them before the threads
 *    start, and profile from one thread.
 */
#ifndef VTC_B200_H_
#define VTC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VTC_OK 0
#define VTC_ERR_ARG 1        /* bad argument -> AssertionError / ValueError in the Python shim */
#define VTC_ERR_CUDA 2       /* CUDA runtime / driver failure */
#define VTC_ERR_WORKSPACE 3  /* workspace too small or misaligned */
#define VTC_ERR_UNSUPPORTED 4 /* -> NotImplementedError */
#define VTC_ERR_NONFINITE 5  /* dictionary overflowed -> RuntimeError (ista_fista.py:75-79) */

/* Arithmetic of the tensor-core contractions: number of bf16 x bf16 products per fp32 product.
 * 1 = plain bf16 operands; 3 = hi/lo split (error ~2^-17, the default "fp32" path); 6 = three-way split (~fp32). */
#define VTC_PRECISION_BF16 1
#define VTC_PRECISION_BF16X3 3
#define VTC_PRECISION_BF16X6 6

#define VTC_VARIANT_ISTA 0
#define VTC_VARIANT_FISTA 1

typedef void* vtc_stream_t; /* a cudaStream_t */

int vtc_version(void);
const char* vtc_last_error(void);

/* Measurement hooks (bench.py): cumulative number of kernels this library has launched in this process, and
 * CUDA-event timing of the last vtc_fista_fc call, recorded on the call's own stream: setup (step size, splits,
 * Gram-form GEMMs), all iteration launches together, and the mean duration of 16 sampled launches of the fused
 * ISTA/FISTA kernel (plus, in the synthesis form, of the r = y Phi - x launch before it).
 * vtc_profile_last synchronises on the last event. */
long long vtc_launch_count(void);
int vtc_profile_enable(int on);
int vtc_profile_last(float* setup_ms, float* iter_ms, int* iter_launches, int* iters, float* fused_launch_ms,
                     float* first_launch_ms);
/* Every profiled vtc_fista_fc call since vtc_profile_enable(1) (a ring of the last 64), oldest first: device time of the
 * setup (step size, operand splits), of the iteration launches, of the finishing copy, and the idle time of the stream
 * until the next profiled call began (0 for the last). Measurement plumbing of bench.py; no reference counterpart. */
int vtc_profile_history_count(void);
int vtc_profile_history(int index, float* setup_ms, float* iter_ms, float* finish_ms, float* gap_to_next_ms);

/* Contraction used by one ISTA/FISTA iteration: 1 = Gram form (y G - b, G = Phi Phi^T; 2*S*S flops per patch, one
 * launch), 2 = synthesis/analysis form ((y Phi - x) Phi^T as in ista_fista.py:105-106; 4*S*D flops, two launches),
 * 0 = automatic (synthesis when S > 2 D). Both give the same iterates up to rounding. Also VTC_B200_FORMULATION. */
int vtc_set_formulation(int formulation);
int vtc_get_formulation(int64_t S, int64_t D);

/* Schedule of a synthesis-form iteration: 1 (default) = the panel-resident kernel that keeps the operand y_k on chip
 * (panel-resident kernel: r_{k-1} Phi^T -> fused update -> y_k Phi - x accumulated in TMEM), used when D <= 256 and the
 * precision is bf16 or bf16x3 -- all iterations of a call in one persistent launch (VTC_B200_PERSISTENT=0: one launch
 * per iteration; early stopping always launches per iteration); 0 = the two-launch schedule (r = y Phi - x, then r Phi^T with the fused update), which
 * is also what larger D and bf16x6 use. Both give identical iterates for identical arithmetic order per element up to
 * the fp32 accumulation order of the synthesis contraction. Also VTC_B200_FUSED_ITER. vtc_get_fused_iteration reports
 * whether a problem of this shape would run the one-launch schedule. */
int vtc_set_fused_iteration(int on);
int vtc_get_fused_iteration(int64_t S, int64_t D, int precision);
/* Small problems in the Gram form (S <= 256 atoms, at most 2368 patches, scalar threshold, no early stopping, bf16 or
 * bf16x3; e.g. BASELINE configs[0]) run ALL iterations in one launch with G, the drive, both iterates and the operand y
 * resident on chip (csrc/fista_small_kernel.cuh); 0 keeps the tiled schedule (one launch per iteration) for them.
 * Identical results either way. Also VTC_B200_SMALL. */
int vtc_set_small_batch_kernel(int on);
/* Debug aid (tools/iter_trace.py): four threads of CTA 0 of the NEXT one-launch iteration kernel write a timeline into
 * device_buffer (4 regions of 2048 uint64 words: [0] = event count, then event id << 48 | SM clock), zeroed by the
 * caller. One shot: the pointer is dropped after that launch. Only in a library built with -DVTC_TRACE (the shipped
 * kernels carry no trace points); VTC_ERR_ARG otherwise. */
int vtc_debug_iter_trace(void* device_buffer);

/* Number of concurrent half-batch chains vtc_fista_fc uses for this problem (1 unless VTC_B200_CHAINS=2 asks for two;
 * synthesis form and large batches only; measured neutral, hence opt-in). With 2, the tensor-bound and the HBM-bound launch of an iteration overlap across
 * the two halves of the batch, each on half of the SMs, on the caller's stream and an internal side stream that is
 * forked from and joined back into the caller's stream inside the call. */
int vtc_get_chains(int64_t B, int64_t S, int64_t D);

/* Number of SMs / compute capability of the current device (used by bench.py to size workloads). */
int vtc_device_info(int* sm_count, int* cc_major, int* cc_minor);

/*
 * ISTA / FISTA code inference, fully connected, including the subspace (group-shrinkage) prox.
 * Replaces analysis_transforms/fully_connected/ista_fista.py:14-148 (run) and, with group_size > 1,
 * subspace_ista_fista.py:23-192 on a dictionary whose groups are runs of `group_size` adjacent atoms
 * (the drop-in builds that "grouped dictionary" exactly as subspace_ista_fista.py:94-111 does).
 *
 *   images        (B, D) float32, pitch ld_images        -- never written
 *   dictionary    (S, D) float32, dense                  -- never written
 *   initial_codes (B, S) float32, pitch ld_codes, or NULL -- never written (tests/ista_fista_1.py:50-54)
 *   codes_out     (B, S) float32, pitch ld_codes          -- the last thresholded iterate (ista_fista.py:148)
 *   early_stopping_epsilon < 0 disables early stopping; otherwise the call synchronises the stream once per
 *   iteration to test mean(|a_k - a_{k-1}|)/stepsize < eps and k > 1 (ista_fista.py:135-144).
 *   iters_run, lipschitz_out: optional HOST outputs; passing lipschitz_out forces one stream synchronisation and
 *   enables the non-finite-dictionary check (VTC_ERR_NONFINITE).
 */
size_t vtc_fista_workspace_bytes(int64_t B, int64_t S, int64_t D, int precision);
int vtc_fista_fc(const float* images, int64_t ld_images, const float* dictionary, const float* initial_codes,
                 float* codes_out, int64_t ld_codes, int64_t B, int64_t S, int64_t D, float sparsity_weight,
                 int num_iters, int variant, int nonnegative_only, int hard_threshold, int group_size,
                 float early_stopping_epsilon, int precision, void* workspace, size_t workspace_bytes,
                 int* iters_run, float* lipschitz_out, vtc_stream_t stream);

/*
 * Sparse-coding dictionary gradient  grad_sum = codes^T (codes * dictionary - images)   (S, D), NOT divided by B.
 * First half of dict_update_rules/fully_connected/sc_cheap_quadratic_descent.py:43-44 (and sc_steepest_descent.py:
 * 38-39). Kept separate from the apply step so that a data-parallel caller can all-reduce grad_sum in between.
 */
size_t vtc_dict_grad_workspace_bytes(int64_t B, int64_t S, int64_t D, int precision);
int vtc_sc_dict_grad(const float* images, int64_t ld_images, const float* dictionary, const float* codes,
                     int64_t ld_codes, float* grad_sum, int64_t B, int64_t S, int64_t D, int precision,
                     void* workspace, size_t workspace_bytes, vtc_stream_t stream);

/*
 * Apply step, in place on `dictionary`:
 *   U = stepsize * (grad_sum / batch_global [+ alignment_penalty * alignment_grad]);
 *   if hessian_diagonal: U /= (h + lowest_code_val); dictionary -= U; if normalize: every row divided by its L2 norm.
 * sc_cheap_quadratic_descent.py:43-48; hessian_diagonal == NULL gives sc_steepest_descent.py:37-41; alignment_grad
 * (S, D) != NULL adds the regulariser of subspace_sc_cheap_quadratic_descent.py:71-75.
 */
int vtc_sc_dict_apply(float* dictionary, const float* grad_sum, const float* hessian_diagonal,
                      const float* alignment_grad, float alignment_penalty, int64_t S, int64_t D,
                      int64_t batch_global, float stepsize, float lowest_code_val, int normalize,
                      vtc_stream_t stream);

/*
 * Gradient of the within-group alignment penalty (sum over pairs in a group of |cos(phi_i, phi_j)|),
 * subspace_sc_cheap_quadratic_descent.py:59-70 and :91-127. group_slots (num_groups * group_width) int32 on the device:
 * atom of each (group, position) slot, -1 for padding; atoms in several groups accumulate. alignment_grad (S, D) is
 * overwritten. dictionary_is_normalized selects the reference's shortcut for unit-norm atoms (its dict_is_normalized).
 */
int vtc_subspace_alignment_grad(const float* dictionary, int64_t S, int64_t D, const int32_t* group_slots,
                                int64_t num_groups, int64_t group_width, int dictionary_is_normalized,
                                float* alignment_grad, vtc_stream_t stream);

/*
 * Hessian-diagonal running average kept by the trainer (training/sparse_coding.py:154):
 *   h <- 0.99 * h + colmean(codes^2) / 100, with the mean taken over batch_global rows.
 * code_sq_sum (S,) receives sum_rows codes^2 of THIS shard (so it can be all-reduced); pass apply_ema = 1 to fold
 * it into h directly (single GPU).
 */
int vtc_hessian_diag_update(const float* codes, int64_t ld_codes, int64_t B, int64_t S, int64_t batch_global,
                            float* code_sq_sum, float* hessian_diagonal, int apply_ema, vtc_stream_t stream);

/* Second half of the running mean when code_sq_sum has been summed over shards: h <- 0.99 h + (sum/batch)/100. */
int vtc_hessian_ema(float* hessian_diagonal, const float* code_sq_sum, int64_t S, int64_t batch_global,
                    vtc_stream_t stream);

/*
 * Generic contraction exposed for tests and tools: out (M, N) = A (M, K) * B (N, K)^T [- sub (M, N)], float32 in and
 * out, computed on the tcgen05 path with the requested precision.
 */
size_t vtc_matmul_nt_workspace_bytes(int64_t M, int64_t N, int64_t K, int precision);
int vtc_matmul_nt(const float* A, const float* B, const float* sub, float* out, int64_t M, int64_t N, int64_t K,
                  int precision, void* workspace, size_t workspace_bytes, vtc_stream_t stream);

/* Largest eigenvalue of dictionary^T dictionary (ista_fista.py:72-74), written to a DEVICE float. */
size_t vtc_lipschitz_workspace_bytes(int64_t S, int64_t D);
int vtc_lipschitz(const float* dictionary, int64_t S, int64_t D, float* lipschitz_dev, void* workspace,
                  size_t workspace_bytes, vtc_stream_t stream);

/* Subspace helpers (subspace_ista_fista.py:94-111 and :184-190): gather dictionary rows into the padded grouped
 * dictionary, gather initial codes, and scatter-add grouped codes back. index (n_slots,) int32 on the device holds
 * the atom of each slot or -1 for padding. */
int vtc_gather_rows(const float* src, int64_t ld_src, const int32_t* index, int64_t n_slots, int64_t D, float* dst,
                    vtc_stream_t stream);
int vtc_gather_cols(const float* src, int64_t ld_src, const int32_t* index, int64_t B, int64_t n_slots, float* dst,
                    int64_t ld_dst, vtc_stream_t stream);
int vtc_scatter_add_cols(const float* src, int64_t ld_src, const int32_t* index, int64_t B, int64_t n_slots,
                         float* dst, int64_t ld_dst, int64_t S, vtc_stream_t stream);

/*
 * ---- Convolutional sparse coding (BASELINE.json configs[4]) ----
 *
 * ISTA / FISTA code inference with strided convolutional synthesis. Replaces
 * analysis_transforms/convolutional/ista_fista.py:17-197 (run):
 *   codes <- prox(y - eta * conv2d(mask * (conv_transpose2d(y, dictionary, stride) - images_padded), dictionary, stride))
 * with eta = 1 / lambda_max of the Gram matrix of the flattened kernels (:104-113) and mask = create_mask
 * (utils/convolutions.py:17-24). The strided convolutions run as tcgen05 GEMMs over stride-sized image blocks (the
 * (KH/SY)*(KW/SX) kernel taps are row shifts of the same operand), the fused epilogues are those of vtc_fista_fc.
 *
 *   images_padded (B, C, H, W) float32 dense  -- never written; H - KH and W - KW must be multiples of the stride
 *   dictionary    (S, C, KH, KW) float32 dense -- never written; KH % SY == 0 and KW % SX == 0 (else VTC_ERR_UNSUPPORTED)
 *   initial_codes (B, S, SH, SW) or NULL, codes_out (B, S, SH, SW), SH = (H - KH) / SY + 1, SW = (W - KW) / SX + 1
 *   pad_*: rows / columns of the padded border whose reconstruction error is ignored (padding_dims of the reference)
 *   early_stopping_epsilon, iters_run, lipschitz_out: as vtc_fista_fc.
 */
size_t vtc_fista_conv_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W, int64_t S, int64_t KH, int64_t KW,
                                      int64_t SY, int64_t SX, int precision);
int vtc_fista_conv(const float* images_padded, const float* dictionary, const float* initial_codes, float* codes_out,
                   int64_t B, int64_t C, int64_t H, int64_t W, int64_t S, int64_t KH, int64_t KW, int64_t SY,
                   int64_t SX, int pad_top, int pad_bottom, int pad_left, int pad_right, float sparsity_weight,
                   int num_iters, int variant, int nonnegative_only, int hard_threshold,
                   float early_stopping_epsilon, int precision, void* workspace, size_t workspace_bytes,
                   int* iters_run, float* lipschitz_out, vtc_stream_t stream);

/*
 * Convolutional dictionary gradient, summed over the batch and NOT divided by it, in dictionary layout (S, C, KH, KW):
 *   conv2d((mask * (conv_transpose2d(codes, dictionary) - images_padded))^T, codes^T, dilation = stride)^T
 * (dict_update_rules/convolutional/sc_cheap_quadratic_descent.py:65-71, sc_steepest_descent.py:59-65). Separate from
 * the apply step so that a data-parallel caller can all-reduce grad_sum in between.
 */
size_t vtc_conv_dict_grad_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W, int64_t S, int64_t KH,
                                          int64_t KW, int64_t SY, int64_t SX, int precision);
int vtc_sc_conv_dict_grad(const float* images_padded, const float* dictionary, const float* codes, float* grad_sum,
                          int64_t B, int64_t C, int64_t H, int64_t W, int64_t S, int64_t KH, int64_t KW, int64_t SY,
                          int64_t SX, int pad_top, int pad_bottom, int pad_left, int pad_right, int precision,
                          void* workspace, size_t workspace_bytes, vtc_stream_t stream);

/*
 * Apply step, in place on `dictionary` (S, per_kernel = C*KH*KW):
 *   U = grad_sum / batch_global; if hessian_diagonal: U /= (h + lowest_code_val); U *= ||dictionary|| / ||U||;
 *   dictionary -= stepsize * U; if normalize: every kernel divided by its L2 norm.
 * convolutional/sc_cheap_quadratic_descent.py:72-79; hessian_diagonal == NULL gives sc_steepest_descent.py:66-72.
 */
int vtc_sc_conv_dict_apply(float* dictionary, const float* grad_sum, const float* hessian_diagonal, int64_t S,
                           int64_t per_kernel, int64_t batch_global, float stepsize, float lowest_code_val,
                           int normalize, vtc_stream_t stream);

/*
 * Hessian-diagonal running average of the convolutional trainer (training/sparse_coding.py:158-161):
 *   h <- 0.99 * h + mean_over_images(sum_over_positions(codes^2)) / 100.
 * codes (B, S, positions); code_sq_sum (S,) receives the sum over this shard's images and positions.
 */
int vtc_conv_hessian_diag_update(const float* codes, int64_t B, int64_t S, int64_t positions, int64_t batch_global,
                                 float* code_sq_sum, float* hessian_diagonal, int apply_ema, vtc_stream_t stream);

/*
 * Validation metrics of the trainer (training/sparse_coding.py:177-229, compute_metrics) from device-resident tensors,
 * without copying images, codes or reconstructions to the host. totals (8 doubles, device):
 *   [0] sum over items of 0.5 * ||codes*dictionary - image||^2          -> 'Average LASSO L2 component' * items
 *   [1] sum over items of ||codes||_1, or of sum_g ||codes_g||_2 when group_slots is given (:199-210)
 *   [2] sum over items of (number of non-zero codes / codes per item)   -> 'Average Normalized L0' * items
 *   [3] sum over items with a non-zero mean squared error of log10(mse), [4] the number of such items
 *       (pSNR_b = 10 log10(sig^2 / mse_b), utils/plotting.py:17-39, items with mse == 0 are skipped, :223)
 *   [5], [6] min and max pixel of the batch (sig = max - min, :217)     [7] number of items
 * group_slots (num_groups x group_width int32 on the device, -1 = padding) or NULL. Sums have a fixed order.
 */
size_t vtc_sc_metrics_workspace_bytes(int64_t B, int64_t S, int64_t D, int precision);
int vtc_sc_metrics(const float* images, int64_t ld_images, const float* dictionary, const float* codes,
                   int64_t ld_codes, int64_t B, int64_t S, int64_t D, const int32_t* group_slots, int64_t num_groups,
                   int64_t group_width, int precision, double* totals, void* workspace, size_t workspace_bytes,
                   vtc_stream_t stream);
/*
 * The same for the convolutional mode (:185-195): reconstruction = conv_transpose2d(codes, dictionary, stride); both it
 * and the images are cropped to the un-padded region before any metric is taken. Arguments as vtc_fista_conv.
 */
size_t vtc_sc_conv_metrics_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W, int64_t S, int64_t KH,
                                           int64_t KW, int64_t SY, int64_t SX, int precision);
int vtc_sc_conv_metrics(const float* images_padded, const float* dictionary, const float* codes, int64_t B, int64_t C,
                        int64_t H, int64_t W, int64_t S, int64_t KH, int64_t KW, int64_t SY, int64_t SX, int pad_top,
                        int pad_bottom, int pad_left, int pad_right, int precision, double* totals, void* workspace,
                        size_t workspace_bytes, vtc_stream_t stream);
/* mean_abs_change[s] = mean over the element's pixels of |dictionary - previous_dictionary| (:226-228) */
int vtc_dict_change(const float* dictionary, const float* previous_dictionary, int64_t S, int64_t per_kernel,
                    float* mean_abs_change, vtc_stream_t stream);

/*
 * Data feed: B crops of (ph x pw) pixels out of device-resident images (n, h, w, c), each flattened in (y, x, c) order --
 * the patch extraction of utils/dataset_generation.py:207-218 (all_patches[p] = img[v:v+ph, u:u+pw], then reshape(N, -1))
 * without the host loop. corners (B, 3) int32 on the device: image index, top row, left column (the caller draws them;
 * out-of-range corners are the caller's error). patches (B, ph*pw*c) with row pitch ld_patches.
 */
int vtc_extract_patches(const float* images, int64_t n, int64_t h, int64_t w, int64_t c, const int32_t* corners,
                        int64_t B, int64_t ph, int64_t pw, float* patches, int64_t ld_patches, vtc_stream_t stream);

/*
 * Data feed: centre-surround whitening of whole images (utils/image_processing.py:267-308, whiten_center_surround; the
 * low-pass factor is get_low_pass_filter :173-231 with shape 'exponential', the ramp get_whitening_ramp_filter :234-264).
 * vtc_whitening_filter writes the real transfer function of an (h, w) DFT, fp64 arithmetic, float32 result:
 *   max(|f|, cutoff_low) * exp(-(|f| / (0.5 * cutoff_high))^order), |f| from fftfreq; with norm_and_threshold divided by its
 *   maximum and floored at 1e-3 (:300-302). scratch8: 8 bytes of device scratch.
 * vtc_spectrum_filter multiplies a complex64 spectrum (n, h*w, c), in place, by that function (filter_fd :63-92: every
 * colour channel is filtered independently). The DFTs on either side are the caller's (cuFFT through torch.fft).
 */
int vtc_whitening_filter(int64_t h, int64_t w, double cutoff_low, double cutoff_high, double order,
                         int norm_and_threshold, float* filter_out, void* scratch8, vtc_stream_t stream);
int vtc_spectrum_filter(void* spectrum, int64_t n, int64_t hw, int64_t c, const float* filter, vtc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VTC_B200_H_ */
