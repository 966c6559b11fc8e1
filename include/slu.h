/* slu.h -- C ABI of libslu.so: the B200 (sm_100a) hot path of SemanticLiDARUnc.
 *
 * The reference (kav-institute/SemanticLiDARUnc) is pure Python and has no FFI;
 * each entry point below replaces the numpy / eager-torch code cited beside
 * it, and is what a reference-side ctypes binding calls.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller (a torch
 *     tensor's data_ptr()); the library allocates nothing persistent and never
 *     frees caller memory.  h_* pointers are HOST memory read before return.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *     entry point synchronises the device, none touches global device state.
 *   - accumulators (confmat, ece_bins) are ADDED to: the caller zeroes them to
 *     reset, which keeps the reference's update()/compute()/reset() semantics.
 *     They are plain int64 so shards combine with one integer all-reduce and
 *     the N-GPU result is bit-identical to the 1-GPU result.
 *   - return value: 0 ok; <0 bad argument (SLU_E_*); >0 a cudaError_t.
 *     slu_last_error() gives the message for the calling thread.
 *   - image tensors are [.., C, H*W] "planar" exactly as the reference's
 *     [B,C,H,W] contiguous tensors; HW = H*W.
 */
#ifndef SLU_H_
#define SLU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLU_VERSION 100            /* 0.1.0 */

#define SLU_E_ARG      (-1)        /* null / negative / inconsistent argument        */
#define SLU_E_RANGE    (-2)        /* size outside what the kernels support           */
#define SLU_E_ALIGN    (-3)        /* pointer not aligned as documented               */
#define SLU_E_DEVICE   (-4)        /* not an sm_100 device / no device                */
#define SLU_E_IO       (-5)        /* a scan file could not be read / is malformed    */

#define SLU_MAX_CLASSES 32         /* register-resident class axis                    */
#define SLU_MAX_BINS    64         /* reliability bins                                */

/* what the class axis of `d_in` holds (reference: ECEAggregator mode, src/metrics/ece.py:54-64) */
#define SLU_IN_LOGITS 0            /* softmax over classes, per sample                */
#define SLU_IN_PROBS  1            /* already probabilities, per sample               */
#define SLU_IN_ALPHA  2            /* Dirichlet concentrations (T must be 1)          */

/* how the top-label confidence is formed from the mean distribution p_bar */
#define SLU_CONF_RAW     0         /* conf = max_c p_bar                              (ece.py:60-61 'logits') */
#define SLU_CONF_RENORM  1         /* conf = max_c  max(p,0) / max(sum_c max(p,0), eps)  (ece.py:62-63 'probs') */
/* with SLU_IN_ALPHA the distribution is alpha / (alpha0 + eps) (ece.py:57-58) and conf_mode RAW applies to it */

typedef void* slu_stream_t;

int         slu_version(void);
const char* slu_last_error(void);
/* number of CUDA kernels this library has launched in the calling process so far (all threads, all devices):
 * the difference across a region is that region's launch count (bench.py reports it as gpu_launches) */
int64_t     slu_launch_count(void);
/* sm count and compute capability of `device`; fails with SLU_E_DEVICE if it is not cc 10.x */
int         slu_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Stage 3+4 fused: per-pixel reduction over T samples x C classes, plus the evaluation
 * histograms, in ONE pass over d_in.
 * Replaces: src/models/tester.py:412-471 (softmax, mean_T, argmax, H_norm, MI_norm, then
 *           IoUEvaluator.update src/models/evaluator.py:39-53 and ECEAggregator.update
 *           src/metrics/ece.py:66-90 with max_samples=None); T=1 covers the single-pass
 *           branches src/models/trainer.py:1170-1225 and src/utils/mc_dropout.py:121-133.
 *
 *   d_in      [T,B,C,HW] float32
 *   d_labels  [B,HW] int64, or NULL (then no histogram is touched)
 *   eps       clamp inside the entropies (reference default 1e-12)
 *   normalize 1: entropies are divided by log C (the *_norm maps); 0: left in nats
 *   ignore    label value excluded from the ECE bins when has_ignore != 0
 *   h_edges   n_bins+1 float32 bin edges (host), [lo,hi) with a closed last bin
 *   outputs (each may be NULL):
 *     d_pbar   [B,C,HW] float32   mean distribution
 *     d_pred   [B,HW]   int64     argmax_c p_bar (first index on ties)
 *     d_conf   [B,HW]   float32   top-label confidence per conf_mode
 *     d_hnorm  [B,HW]   float32   -sum p_bar log p_bar / log C   (clamped at eps)
 *     d_minorm [B,HW]   float32   max(0, (H[p_bar] - mean_t H[p_t]) / log C)
 *   accumulators (each may be NULL):
 *     d_confmat  [C*C]      int64  rows = label, cols = pred; labels outside [0,C) dropped
 *     d_ece_bins [3*n_bins] int64  n | n_correct | sum(conf) in units of 2^-32
 */
int slu_reduce_metrics(const float* d_in, const int64_t* d_labels,
                       int T, int B, int C, int64_t HW,
                       int in_kind, int conf_mode, float eps, int normalize,
                       int has_ignore, int64_t ignore,
                       int n_bins, const float* h_edges,
                       float* d_pbar, int64_t* d_pred, float* d_conf, float* d_hnorm, float* d_minorm,
                       int64_t* d_confmat, int64_t* d_ece_bins,
                       slu_stream_t stream);

/* Same work with every thread loading straight from global memory (no TMA staging): used for
 * shapes the bulk-copy path cannot take (HW % 4 != 0 or unaligned d_in) and as the A/B
 * comparison in profiles/.  slu_reduce_metrics() dispatches to it on its own when needed. */
int slu_reduce_metrics_direct(const float* d_in, const int64_t* d_labels,
                              int T, int B, int C, int64_t HW,
                              int in_kind, int conf_mode, float eps, int normalize,
                              int has_ignore, int64_t ignore,
                              int n_bins, const float* h_edges,
                              float* d_pbar, int64_t* d_pred, float* d_conf, float* d_hnorm, float* d_minorm,
                              int64_t* d_confmat, int64_t* d_ece_bins,
                              slu_stream_t stream);

/* A/B switch (tests, profiles): 1 = single-sample inputs (T == 1) also take the multi-sample staged kernel instead
 * of the dedicated one-thread-per-pixel kernel slu_reduce_metrics picks for them. */
int slu_debug_reduce_no_single(int on);
/* A/B switch (tests, profiles): 1 = single-sample LOGITS take the general single-sample kernel (shared-atomic histograms)
 * instead of the specialised one with thread-private reliability cells. */
int slu_debug_reduce_no_private(int on);


/* ---------------------------------------------------------------------------------------------
 * Stage 3+4 for the evidential (Dirichlet) head, single pass.
 * Replaces: src/models/tester.py:484-512 = to_alpha_concentrations_from_shape_and_scale
 *           (src/models/probability_helper.py:89-105), get_predictive_entropy (:116-121),
 *           get_aleatoric_uncertainty (:124-130), get_epistemic_uncertainty (:133-136), the Dirichlet
 *           mutual information of src/metrics/auroc.py:55-63, IoUEvaluator.update and
 *           ECEAggregator.update(mode='alpha').
 *   exactly one of:  d_outputs  [B,C+1,HW] float32 head output (C shape logits, then 1 scale logit)
 *                    d_alpha_in [B,C,HW]   float32 concentrations
 *   temperature (reference global 1.0), eps (probability_helper eps, 1e-8), eps_metrics (ECE/AUROC eps, 1e-12)
 *   outputs (each may be NULL):
 *     d_alpha_out [B,C,HW] alpha = 1 + softplus(s/T) * softmax(z) + eps
 *     d_pred [B,HW] int64  argmax softmax(z) (argmax alpha when alpha is the input)
 *     d_conf [B,HW]        max_c alpha / (alpha0 + eps_metrics)
 *     d_h    [B,HW]        predictive entropy (/ log C when normalize)
 *     d_au, d_eu [B,HW]    aleatoric / epistemic uncertainty in nats
 *     d_mi   [B,HW]        Dirichlet mutual information, auroc.py convention (/ log C when normalize)
 *   accumulators as in slu_reduce_metrics.
 */
int slu_evidential_reduce(const float* d_outputs, const float* d_alpha_in, const int64_t* d_labels,
                          int B, int C, int64_t HW, float temperature, float eps, float eps_metrics,
                          int normalize, int has_ignore, int64_t ignore, int n_bins, const float* h_edges,
                          float* d_alpha_out, int64_t* d_pred, float* d_conf, float* d_h, float* d_au,
                          float* d_eu, float* d_mi, int64_t* d_confmat, int64_t* d_ece_bins,
                          slu_stream_t stream);
/* A/B switch (tests, profiles): 1 = one pixel per thread instead of the packed (two adjacent pixels per thread, FFMA2 /
 * FMUL2 / FADD2) kernel slu_evidential_reduce picks for even HW and aligned buffers.  Returns the previous value; a
 * negative argument only queries. */
int slu_debug_no_packed_evidential(int on);

/* ---------------------------------------------------------------------------------------------
 * Evidential loss terms, forward + backward in one pass (training-step data path, config 5).
 * Replaces: DirichletMSELoss (src/losses/dirichlet_losses.py:317-385), KL_offClasses_to_uniform
 *           (src/losses/regularizers.py:291-389, with_conf_weighting=False), _valid_mask (:15-70).
 *   d_alpha  [B,C,HW] float32;  d_target [B,HW] int64 (0 <= target < C on valid pixels)
 *   validity: d_keep_mask [B,HW] uint8 (1 = valid) if given, else target not in h_ignore[0..n_ignore)
 *   d_sums   [3] float64, ADDED to: sum of per-pixel mse | sum of per-pixel kl | number of valid pixels.
 *            loss_term = sums[term] / max(sums[2], 1)
 *   d_grad_* [B,C,HW] float32 or NULL: d(per-pixel term)/d(alpha), 0 on masked pixels; the caller
 *            scales by upstream_grad / max(n_valid, 1).
 */
int slu_dirichlet_loss(const float* d_alpha, const int64_t* d_target, const uint8_t* d_keep_mask,
                       int B, int C, int64_t HW, const int64_t* h_ignore, int n_ignore,
                       float eps_mse, float eps_kl, int want_mse, int want_kl,
                       double* d_sums, float* d_grad_mse, float* d_grad_kl, slu_stream_t stream);

/* Fused training-step loss from the head output (SURVEY.md 8d "L", 176 B/px): alpha = 1 + softplus(l/T) *
 * softmax(z) + eps_alpha, loss = w_mse * MSE + w_kl * KL, and d(loss)/d(head output) with the mean over valid
 * pixels already applied.  Replaces src/models/trainer.py:532-578 plus autograd's backward through the
 * alpha transform.  C >= 3 (the reference's MSE term is defined as 0 for C <= 2).
 *   d_outputs [B,C+1,HW];  d_sums [3] float64 must be ZERO on entry: sum mse | sum kl | n_valid on exit;
 *   loss = (w_mse*sums[0] + w_kl*sums[1]) / max(sums[2],1);  d_grad_outputs [B,C+1,HW] or NULL.
 *   precounted != 0: d_sums[2] already holds the number of valid pixels the mean runs over (d_sums[0..1] zero) and no
 *   count kernel is launched -- the batch-sharded training step: every rank counts its pixels (slu_count_valid), the
 *   counts are all-reduced, and each rank's gradient and loss share come out as fractions of the GLOBAL mean.
 */
int slu_evidential_loss_fused(const float* d_outputs, const int64_t* d_target, const uint8_t* d_keep_mask,
                              int B, int C, int64_t HW, const int64_t* h_ignore, int n_ignore,
                              float temperature, float eps_alpha, float eps_mse, float eps_kl,
                              float w_mse, float w_kl, int precounted, double* d_sums, float* d_grad_outputs, slu_stream_t stream);

/* Number of valid pixels (validity as in slu_dirichlet_loss), ADDED to d_count[0] (float64). */
int slu_count_valid(const int64_t* d_target, const uint8_t* d_keep_mask, int64_t n_px,
                    const int64_t* h_ignore, int n_ignore, double* d_count, slu_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The exchange of the batch-sharded training step over NVLink / NVSwitch peer memory (BASELINE.json
 * configs[4]): slu_count_valid fused with the all-reduce of its result.  No reference counterpart (the reference is a
 * single process; its dormant all_reduce is src/utils/agg.py:75-83).
 *
 *   slu_peer_mailbox_create   allocates this rank's mailbox (264 KB) on the current device (zeroed) and returns its device
 *                             pointer and the 64 bytes of its CUDA IPC handle; the caller passes the handle bytes to the
 *                             other ranks' processes (one all-gather at set-up)
 *   slu_peer_mailbox_open     maps another rank's mailbox (same node) into this process; _close unmaps it
 *   slu_peer_mailbox_destroy  frees the own mailbox (after every peer has closed it)
 *   slu_count_valid_exchange  ONE kernel: counts the valid pixels of (d_target, mask) as slu_count_valid does; its last CTA
 *                             stores (step << 32 | count) into slot [rank] of every rank's mailbox, waits until all
 *                             `world` words of this step have arrived in its own mailbox and writes their sum -- an
 *                             integer, so bit-identical on every rank -- to d_count[0] (float64, OVERWRITTEN).
 *                             h_boxes [world]: HOST array of device pointers, h_boxes[r] = rank r's mailbox as mapped in
 *                             this process (h_boxes[rank] = the own one).  Every rank of the group must make the same
 *                             sequence of calls; the step number lives in the mailbox, so the launch can be captured
 *                             in a CUDA graph and replayed.  n_px < 2^32, world <= 16, n_ignore <= 8.  A peer that does
 *                             not show up within timeout_s (<= 0: 2 s) makes d_count NaN instead of hanging
 *                             (slu_peer_mailbox_timeouts counts such events; it synchronises).
 *   slu_peer_allreduce_i64    ONE single-CTA kernel: d_out[0 .. n_a+n_b) = sum over the ranks of the int64 vector
 *                             (d_a[0..n_a), d_b[0..n_b)) -- the end-of-sweep combination of the confusion matrix and the
 *                             reliability bins (n_a + n_b <= 512, d_b may be NULL).  Payload stores into every
 *                             rank's mailbox, a system-scope fence, a release store of the step number, an acquire wait
 *                             for the peers' step numbers, sums in rank order: integers, identical bits on every rank.
 *                             Inputs are left untouched.  A missing peer (timeout_s) fills d_out with INT64_MIN.
 */
int slu_peer_mailbox_create(void** d_box_out, uint8_t* handle64_out);
int slu_peer_mailbox_open(const uint8_t* handle64, void** d_box_out);
int slu_peer_mailbox_close(void* d_box);
int slu_peer_mailbox_destroy(void* d_box);
int slu_peer_mailbox_timeouts(const void* d_box, uint32_t* h_out);
int slu_count_valid_exchange(const int64_t* d_target, const uint8_t* d_keep_mask, int64_t n_px,
                             const int64_t* h_ignore, int n_ignore,
                             void* const* h_boxes, int rank, int world, double timeout_s,
                             double* d_count, slu_stream_t stream);
int slu_peer_allreduce_i64(const int64_t* d_a, int n_a, const int64_t* d_b, int n_b,
                           void* const* h_boxes, int rank, int world, double timeout_s,
                           int64_t* d_out, slu_stream_t stream);

/* Training-step form of slu_evidential_loss_fused: no host arithmetic and no memset between steps.
 *   d_count  [1] float64  number of valid pixels the masked mean runs over.  precounted == 0: the call counts the valid
 *            pixels of (d_target, mask) into it (it must hold 0 on entry and is reset to 0 by the kernel);
 *            precounted != 0: the caller has written the (all-reduced, GLOBAL) count -- slu_count_valid + one all-reduce,
 *            possibly on another stream, ordered before this call -- and owns the buffer.
 *   d_state  [3] float64  scratch (partial sums, CTA ticket): zero before the FIRST use, left zeroed by every call, so the
 *            same buffers serve every following step and every replay of a captured CUDA graph.
 *   d_loss4  [4] float32  out: loss = w_mse * mse + w_kl * kl | mse | kl | n_valid, written by the last CTA to finish
 *            (with a global count these are this rank's SHARES of the global means).
 *   d_grad_outputs [B,C+1,HW] or NULL: d(loss)/d(outputs), already divided by the count. */
int slu_evidential_loss_step(const float* d_outputs, const int64_t* d_target, const uint8_t* d_keep_mask,
                             int B, int C, int64_t HW, const int64_t* h_ignore, int n_ignore,
                             float temperature, float eps_alpha, float eps_mse, float eps_kl,
                             float w_mse, float w_kl, int precounted, double* d_count, double* d_state,
                             float* d_loss4, float* d_grad_outputs, slu_stream_t stream);
/* A/B switch (tests, profiles): 1 = the loss kernels run one pixel per thread instead of the packed (two adjacent pixels
 * per thread, FFMA2 / FMUL2 / FADD2) variants they pick for even HW and aligned buffers.  Returns the previous value;
 * a negative argument only queries. */
int slu_debug_no_packed_loss(int on);

/* Alternative data-fit terms, one per call, forward + analytic backward.
 * Replaces: NLLDirichletCategorical (src/losses/dirichlet_losses.py:73-119), DigammaDirichletCE (:122-167),
 *           BrierDirichlet (:174-220; s_ref < 0 means "use alpha0").
 *   d_sums [2] float64, ADDED to: sum of per-pixel values | number of valid pixels;  d_grad [B,C,HW] or NULL
 *   holds d(per-pixel value)/d(alpha), 0 on masked pixels; the caller divides by max(n_valid,1).
 */
#define SLU_TERM_NLL        2
#define SLU_TERM_DIGAMMA_CE 3
#define SLU_TERM_BRIER      4
int slu_dirichlet_term(const float* d_alpha, const int64_t* d_target, const uint8_t* d_keep_mask,
                       int B, int C, int64_t HW, const int64_t* h_ignore, int n_ignore,
                       int term, float eps, float s_ref, double* d_sums, float* d_grad, slu_stream_t stream);

/* Remaining evidential terms / regularisers, one per call, forward + analytic backward.
 * Replaces: ComplementKLUniform (src/losses/dirichlet_losses.py:228-314), WrongLowEvidence
 *           (src/losses/regularizers.py:218-289), EvidenceRegBand (:116-147), EvidenceReg (:149-212),
 *           KL_offClasses_to_uniform with_conf_weighting=True (:369-383).
 *   h_params (host floats, n_params must match the term):
 *     SLU_TERM_COMP_KL    7: gamma, tau, sigma, s_target (< 0 = none), normalize (0/1), eps, detach_uncert (0/1)
 *     SLU_TERM_WRONG_LOW  4: s_low, margin, soft_margin_k, eps
 *     SLU_TERM_EVID_BAND  2: s_target, band
 *     SLU_TERM_EVID_REG   4: s_target, mode (0 log_squared | 1 one_sided | 2 l2), margin, scale_correct (0/1)
 *     SLU_TERM_KL_CONF    2: gamma, eps
 *   d_target may be NULL for EVID_BAND / EVID_REG (then every pixel is valid unless d_keep_mask says otherwise).
 *   d_sums [2] float64, ADDED to: sum of per-pixel values | the term's denominator (valid pixels; sum of gates
 *   for WRONG_LOW; sum of weights for KL_CONF);  d_grad [B,C,HW] or NULL = d(per-pixel value)/d(alpha), 0 on masked
 *   pixels.  The caller divides by the reference's clamp of the denominator.
 */
#define SLU_TERM_COMP_KL    5
#define SLU_TERM_WRONG_LOW  6
#define SLU_TERM_EVID_BAND  7
#define SLU_TERM_EVID_REG   8
#define SLU_TERM_KL_CONF    9
int slu_evidence_term(const float* d_alpha, const int64_t* d_target, const uint8_t* d_keep_mask,
                      int B, int C, int64_t HW, const int64_t* h_ignore, int n_ignore,
                      int term, const float* h_params, int n_params,
                      double* d_sums, float* d_grad, slu_stream_t stream);

/* LogitRegularizer (src/losses/regularizers.py:75-110): per element z^2, or relu(z - threshold)^2 when
 * use_threshold != 0, over d_logits [B,Cz,HW] with a per-pixel validity mask (d_target / d_keep_mask may both
 * be NULL = all valid).  d_sums [2] float64 ADDED to: sum over valid elements | number of valid PIXELS (the
 * reference divides the element sum by the pixel count when a mask is given, by the element count otherwise).
 * d_grad [B,Cz,HW] or NULL.
 */
int slu_logit_regularizer(const float* d_logits, const int64_t* d_target, const uint8_t* d_keep_mask,
                          int B, int Cz, int64_t HW, const int64_t* h_ignore, int n_ignore,
                          int use_threshold, float threshold, double* d_sums, float* d_grad, slu_stream_t stream);

/* Diagnostic: d_out[3i..3i+2] = lgamma, digamma, trigamma of d_in[i] (> 0) as the loss kernels evaluate them. */
int slu_diag_special(const float* d_in, int64_t n, float* d_out, slu_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 4 standalone: histograms from already-reduced maps (integer inputs).
 * Replaces: IoUEvaluator.update (src/models/evaluator.py:39-53) and the binning of
 *           ECEAggregator (src/metrics/ece.py:75-90,131-140).
 *   d_pred, d_labels [n] int64;  d_conf [n] float32 or NULL (then only the confusion matrix)
 *   d_confmat [C*C] int64 or NULL; d_ece_bins [3*n_bins] int64 or NULL.
 *   C may be any value up to 1024 here.
 */
int slu_confusion_ece(const int64_t* d_pred, const int64_t* d_labels, const float* d_conf,
                      int64_t n, int C, int has_ignore, int64_t ignore,
                      int n_bins, const float* h_edges,
                      int64_t* d_confmat, int64_t* d_ece_bins, slu_stream_t stream);
/* The same with int32 prediction / label maps (12 B/px instead of 20): for callers that keep their reduced maps in
 * 32 bits.  Counters and semantics are identical; the results are bit-identical to slu_confusion_ece on widened inputs. */
int slu_confusion_ece_i32(const int32_t* d_pred, const int32_t* d_labels, const float* d_conf,
                          int64_t n, int C, int has_ignore, int64_t ignore,
                          int n_bins, const float* h_edges,
                          int64_t* d_confmat, int64_t* d_ece_bins, slu_stream_t stream);
/* A/B switch (tests, profiles): 1 = always run the generic warp-aggregated histogram kernel instead of the
 * streaming one slu_confusion_ece picks for 16-byte aligned inputs.  Results are bit-identical either way. */
int slu_debug_hist_generic(int on);


/* ---------------------------------------------------------------------------------------------
 * Error/score histogram: the sufficient statistic of the reference's AUROC, risk-coverage
 * and accuracy-vs-uncertainty aggregators, which keep every pixel on the host and sort at compute().
 * Replaces the accumulation in AUROCAggregator.update (src/metrics/auroc.py:101-141),
 *           UncertaintyAccuracyAggregator.update (src/models/evaluator.py:660-701),
 *           UncertaintyAggregator.add_batch (src/metrics/aurc.py:273-304).
 *   d_score [n] float32 (clamped to [0,1]; NaN skipped), d_pred / d_labels [n] int64,
 *   h_ignore: up to 4 label values to skip;  d_hist [2, n_score_bins] int64, ADDED to:
 *   hist[0][k] = correct pixels, hist[1][k] = wrong pixels with floor(score * n_score_bins) == k.
 */
int slu_score_hist(const float* d_score, const int64_t* d_pred, const int64_t* d_labels, int64_t n,
                   int n_score_bins, const int64_t* h_ignore, int n_ignore, int64_t* d_hist, slu_stream_t stream);
/* The same accumulation into 2^20 HYBRID bins, d_hist [2, 2^20]: scores pile up at both ends of [0,1] (confident pixels
 * near 0, near-uniform ones near 1), where a uniform grid ties thousands of pixels per bin.  Monotone bin key:
 *   s >= 1/16 : floor(s * 2^20)                                   (step 9.5e-7)
 *   s <  1/16 : 2048 mantissa bins per binary octave, 32 octaves down to 2^-36 (relative step 4.9e-4); smaller -> bin 0.
 * Used by the AUROC aggregator, whose result depends on the ORDER of the scores only. */
int slu_score_hist_hybrid(const float* d_score, const int64_t* d_pred, const int64_t* d_labels, int64_t n,
                          const int64_t* h_ignore, int n_ignore, int64_t* d_hist, slu_stream_t stream);

/* Per-class score histogram: replaces the per-class host arrays of UncertaintyPerClassAggregator.update
 * (src/models/evaluator.py:211-262).  d_hist [C, n_score_bins] int64 and d_sum_fx [C] int64 (sum of scores in
 * units of 2^-32, non-negative scores) are ADDED to; labels outside [0,C) and NaN scores are skipped. */
int slu_class_score_hist(const float* d_score, const int64_t* d_labels, int64_t n, int C, int n_score_bins,
                         int64_t* d_hist, int64_t* d_sum_fx, slu_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 1: spherical range-image projection with a nearest-range depth test, batched.
 * Replaces: spherical_projection (src/dataset/utils.py:288-349) + to_deflection_coordinates
 *           (:61-67) + the KITTI loader glue around it
 *           (src/dataset/dataloader_semantic_KITTI.py:40-49 label remap, :83 range).
 *
 *   d_xyzi      [n_total,4] float32  all scans of the batch, concatenated (16-byte aligned)
 *   d_raw_label [n_total]   uint32   raw labels (semantic id in the low 16 bits) or NULL
 *   d_lut       [65536]     int32    raw id -> train id (missing ids < 0), or NULL = identity
 *   h_offsets   [B+1]       int64    HOST array; scan b owns points [offsets[b], offsets[b+1]); B <= 256
 *   H, W                    image size;  all angle math is float64 (the loaders feed float64)
 *   theta_lo/hi             row range; if use_theta_range == 0 the per-scan min/max of theta is used
 *   farthest_wins           0: nearest point wins (reference default); 1: sort_largest_first=True
 *   d_yaw_cs    [B,2]       float64  (cos a, sin a) of the loaders' yaw augmentation rotate_z
 *                                    (src/dataset/utils.py:4-18), applied in float64 before projection; NULL = none
 *   d_work      workspace of slu_project_workspace_bytes(n_total, B, H*W) bytes, 16-byte aligned
 *   outputs (each may be NULL):
 *     d_img    [B,6,HW] float32  planes x,y,z,range,intensity,label; 0 where empty
 *     d_pix    [n_total] int32   pixel (row*W+col) every point projects to
 *     d_winner [B,HW]   int32    index (within its scan) of the winning point, -1 where empty
 *     d_label  [B,HW]   int64    train id of the winning point (0 where empty): the loaders' `semantics`
 *                                tensor (dataloader_semantic_KITTI.py:97), ready to be the metrics' labels
 *     d_theta  [B,2]    float64  (theta_min, theta_max) used per scan
 *     d_diag   [B,2]    int32    [#points with a raw id missing from the LUT,
 *                                 #points within 4 ulp of a bin edge (index could differ from numpy's)]
 */
int64_t slu_project_workspace_bytes(int64_t n_total, int B, int64_t HW);
/* Test / profiling switch of slu_project_batch: on=1 runs the exact fp64 kernels for every point instead of the
 * fp32-prefiltered ones (bit-identical results, ~2x slower); on=2 runs the three-launch "cell" pipeline (one fused point
 * pass whose depth test is a 128-bit compare-and-swap on a (range, index) cell per pixel; bit-identical results, same
 * speed with an automatic range, slower with a fixed one; needs the larger workspace slu_project_workspace_bytes
 * already reports); on=3 keeps the default kernels but runs their depth test as that 128-bit compare-and-swap (no tie
 * pass; bit-identical, slower: atomics that return a value cost more than reductions); on=0 restores the default; on<0
 * only queries.  Returns the previous setting.  Process-wide host
 * state, not for concurrent use. */
int slu_debug_project_exact(int on);
/* diagnostic: out[i] = the projection prefilter's fp32 arctangent of (y[i], x[i]); NaN for zero / denormal / huge / non-finite
 * inputs (those points always take the fp64 path).  Tests bound its error against float64 atan2. */
int slu_diag_fast_atan2(const float* d_y, const float* d_x, int64_t n, float* d_out, slu_stream_t stream);
int slu_project_batch(const float* d_xyzi, const uint32_t* d_raw_label, const int32_t* d_lut,
                      const int64_t* h_offsets, int64_t n_total, int B, int H, int W,
                      int use_theta_range, double theta_lo, double theta_hi, int farthest_wins,
                      const double* d_yaw_cs, void* d_work,
                      float* d_img, int64_t* d_label, int32_t* d_pix, int32_t* d_winner, double* d_theta,
                      int32_t* d_diag, slu_stream_t stream);

/* Generic form of spherical_projection: pc [N,Cin] float64 (any Cin >= 3), one scan, output in the
 * reference's [H,W,Cin] float32 layout (src/dataset/utils.py:341-344).  Same workspace rule with B=1. */
int slu_project_points(const double* d_pc, int64_t N, int Cin, int H, int W,
                       int use_theta_range, double theta_lo, double theta_hi, int farthest_wins,
                       void* d_work,
                       float* d_img_hwc, int32_t* d_pix, int32_t* d_winner, double* d_theta, int32_t* d_diag,
                       slu_stream_t stream);
/* slu_project_points with caller-supplied ROW edges (the reference's bins_h argument, src/dataset/utils.py:330-338:
 * idx_h = np.digitize(theta, bins_h) - 1).  d_row_edges_ascending [H] float64 holds the edges in ASCENDING order;
 * edges_were_increasing says how the caller's array ran (the reference builds a DEcreasing one: idx = H - 1 - #{e <= theta};
 * for an increasing array idx = #{e <= theta} - 1; -1 wraps to H - 1 either way).  NULL edges = slu_project_points. */
int slu_project_points_bins(const double* d_pc, int64_t N, int Cin, int H, int W,
                            int use_theta_range, double theta_lo, double theta_hi, int farthest_wins,
                            const double* d_row_edges_ascending, int edges_were_increasing, void* d_work,
                            float* d_img_hwc, int32_t* d_pix, int32_t* d_winner, double* d_theta, int32_t* d_diag,
                            slu_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Loader glue behind the projection.
 * Replaces: src/dataset/dataloader_semantic_KITTI.py:60-97 -- cv2.resize INTER_NEAREST (:61-62),
 *           flip + y negation (:71-73), channel split, range (:83), build_normal_xyz (:85,
 *           src/dataset/utils.py:30-59) and the tensor packing (:91-97).
 *   d_img  [B,6,Hs*Ws] planes from slu_project_batch;  (Hd,Wd) output size (nearest-neighbour resize)
 *   h_flip [B] host bytes (1 = mirror columns and negate y) or NULL;  norm_factor: Scharr scale is 1/norm_factor
 *   outputs (each may be NULL): d_range [B,1,Hd*Wd], d_refl [B,1,..], d_xyz [B,3,..], d_normals [B,3,..]
 *   (needs d_xyz), d_sem [B,1,..] int64 -- the five tensors Dataset.__getitem__ returns, stacked over B.
 *   d_rowmap [B, Hs+1] int32 workspace or NULL: when given, image rows without any return are dropped before the
 *   resize (src/dataset/dataloader_semantic_WADS.py:125); rowmap[b][0] returns the number of rows kept.  With
 *   resize_rows == 0 the output keeps Hs rows, of which the first rowmap[b][0] are valid and the rest zero.
 */
int slu_frame_tensors(const float* d_img, int B, int Hs, int Ws, int Hd, int Wd,
                      const uint8_t* h_flip, float norm_factor,
                      float* d_range, float* d_refl, float* d_xyz, float* d_normals, int64_t* d_sem,
                      int32_t* d_rowmap, int resize_rows, slu_stream_t stream);

/* Surface normals straight from planar xyz (no resize / flip): build_normal_xyz, src/dataset/utils.py:30-59, as the
 * loaders call it on the projected image (src/dataset/dataloader_semantic_KITTI.py:85).  The x, y, z planes of scan b start
 * at d_xyz + b * batch_stride elements (3*H*W for a packed [B,3,H,W] tensor; 6*H*W reads the first three planes of the
 * projection image [B,6,H,W] in place).  d_normals [B,3,H*W]. */
int slu_frame_normals(const float* d_xyz, int B, int H, int W, int64_t batch_stride, float norm_factor,
                      float* d_normals, slu_stream_t stream);

/* Organised clouds (Ouster / SemanticTHAB: the sensor delivers H x W points, pixel n = point n, no projection;
 * src/inference_ouster.py:59-62, documentation/dataset.md:109).
 * Replaces: src/dataset/dataloader_semantic_THAB.py:35-66 (label remap, reshape, flip, yaw as image roll +
 *           rotate_z, float64 range) -- writes the same [B,6,HW] planes slu_project_batch writes, so
 *           slu_frame_tensors / slu_backproject (with pix[n] = n) apply unchanged.
 *   d_xyzi [B*H*W,4] float32, d_raw_label [B*H*W] uint32 or NULL, d_lut [65536] or NULL,
 *   h_flip [B] host bytes or NULL, d_col_shift [B] int32 device (np.roll shift) or NULL, d_yaw_cs [B,2] or NULL,
 *   d_missing [B] int32 or NULL: number of raw ids missing from the LUT.
 */
int slu_organized_planes(const float* d_xyzi, const uint32_t* d_raw_label, const int32_t* d_lut,
                         int B, int H, int W, const uint8_t* h_flip, const int32_t* d_col_shift,
                         const double* d_yaw_cs, float* d_img, int32_t* d_missing, slu_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 2: label back-projection, pixels -> points (the reference has no code
 * for it: documentation/dataset.md:109 describes the organised-cloud case only).
 *   point_label[n] = label_img[scan(n)][pix[n]]
 *   d_label_img [B,HW] int64;  d_pix [n_total] int32;  h_offsets [B+1] int64 (host);  d_out [n_total] int64
 */
int slu_backproject(const int64_t* d_label_img, const int32_t* d_pix, const int64_t* h_offsets,
                    int64_t n_total, int B, int64_t HW, int64_t* d_out, slu_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * On-disk scans to HBM: native I/O threads read KITTI-style files into pinned host slots and
 * the consumer copies them to the device on its own stream.
 * Replaces: np.fromfile(frame_path, float32).reshape(-1,4) / np.fromfile(label_path, uint32) at the top of every
 *           loader's __getitem__ (src/dataset/dataloader_semantic_KITTI.py:35-39) and the pageable H2D copy of
 *           the batch (src/models/trainer.py:520-527).
 *   slu_stager_create   n_slots pinned slots of max_points_per_scan points (20 B/point) and n_io_threads readers on
 *                       the CURRENT device; the stager owns them until slu_stager_destroy.  flags: 0, or
 *                       SLU_STAGER_HOST_DEST = pageable slots and HOST destination pointers in slu_stager_fetch
 *                       (no CUDA call at all: exercises the reader threads and the error paths on a box without a GPU).
 *   slu_stager_submit   queue (bin_path, label_path or NULL) -> ticket (0, 1, 2, ... in submission order).
 *   slu_stager_fetch    wait until that scan is in a slot, enqueue H2D copies of its points [n,4] float32 to d_xyzi
 *                       and raw labels [n] uint32 to d_label (skipped if NULL / no label file) on `stream`; the slot
 *                       is recycled when those copies complete.  Fetch tickets in submission order.
 *                       Errors: SLU_E_IO (missing file, size not a multiple of 16 B, label count != point count),
 *                       SLU_E_RANGE (scan larger than capacity_points or than the slot).
 */
#define SLU_STAGER_HOST_DEST 1
int slu_stager_create(int n_slots, int64_t max_points_per_scan, int n_io_threads, int flags, void** out_handle);
int slu_stager_submit(void* handle, const char* bin_path, const char* label_path, int64_t* out_ticket);
int slu_stager_fetch(void* handle, int64_t ticket, float* d_xyzi, uint32_t* d_label, int64_t capacity_points,
                     int64_t* out_n_points, int* out_has_label, slu_stream_t stream);
int slu_stager_destroy(void* handle);

/* ---------------------------------------------------------------------------------------------
 * Diagnostic: stream n float32 from d_in once (16-byte loads, grid = 8 CTAs/SM) and write one
 * float to d_out.  A read-only HBM yardstick for the roofline notes in profiles/; not on any
 * product path.  d_in must be 16-byte aligned, n % 4 == 0.
 */
int slu_diag_read_stream(const float* d_in, int64_t n, float* d_out, slu_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SLU_H_ */
