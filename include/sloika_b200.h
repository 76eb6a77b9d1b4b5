/*
 * sloika_b200.h -- C ABI of the B200 (sm_100a) implementation of Sloika's raw basecall hot path.
 *
 * The reference (nanoporetech/sloika) has no native interface on this path: the network is a Theano
 * graph (`sloika/layers.py`) and the decoder is NumPy (`sloika/decode.py`).  Each entry point below
 * replaces one of those Python-level operators; the citation on each function is the reference code
 * whose result it must reproduce.  A maintainer binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; nothing is allocated, freed or kept
 *     by the library, and no call synchronises the device: work is enqueued on `stream`
 *     (a `cudaStream_t` passed as void*, NULL = legacy default stream) of the CURRENT device
 *   - tensors are float32, `[time, batch, feature]` row-major (`sloika/layers.py:13`); a "row" is one
 *     (time, batch) pair and `ld*` is the distance in floats between consecutive rows
 *     (= feature count when dense), which lets Parallel/birnn outputs be written in place
 *   - `lengths` (nullable, int32[B]) gives the number of valid time steps of each sequence of a padded
 *     ragged batch; NULL means every sequence has the full length
 *   - return value: 0 on success, a negative SLOIKA_ERR_* for rejected arguments, or a positive
 *     `cudaError_t` if the launch failed; `sloika_b200_strerror` describes either
 *   - the functions are re-entrant across streams and devices
 */
#ifndef SLOIKA_B200_H
#define SLOIKA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLOIKA_B200_ABI_VERSION 1

#define SLOIKA_OK               0
#define SLOIKA_ERR_ARG         -1   /* NULL pointer / non-positive size / inconsistent shape      */
#define SLOIKA_ERR_UNSUPPORTED -2   /* valid request with no kernel (e.g. unknown activation)     */
#define SLOIKA_ERR_WORKSPACE   -3   /* workspace pointer missing or too small                     */

/* activation codes: sloika/activation.py:8 (linear), :52 (tanh), :56 (sigmoid), :38-42 (elu) */
#define SLOIKA_ACT_LINEAR  0
#define SLOIKA_ACT_TANH    1
#define SLOIKA_ACT_SIGMOID 2
#define SLOIKA_ACT_ELU     3

int sloika_b200_abi_version(void);
const char *sloika_b200_strerror(int code);

/* Number of SMs / compute capability of the current device (host call, no kernel). */
int sloika_b200_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* Scheduling option (process wide, not a numerical one): the largest number of SMs a tensor-core GEMM launch
 * (sloika_linear_fwd*, sloika_softmax_logits_fwd, the projection inside sloika_gru_fwd) may occupy; 0 = all.  A caller
 * that pipelines several batches on different streams sets it to the SMs its recurrence kernels leave free. */
int sloika_b200_set_gemm_sm_budget(int sms);

/*
 * Convolution.run  -- sloika/layers.py:417-419, sloika/conv.py:66-111
 *   y[t,b,o] = act( bias[o] + sum_i sum_k W[o,i,k] * xpad[t*stride + k, b, i] )
 * cross-correlation (filter_flip=False) over the time axis zero-padded by (pad_l, pad_r);
 * T_out = (T + pad_l + pad_r - winlen) / stride + 1.
 *   x [T,B,Cin] dense; W [Cout,Cin,winlen]; bias [Cout]; y rows of ldy floats, T_out*B rows.
 * With `lengths`, samples t >= lengths[b] of sequence b read as zero (each read is padded with
 * zeros exactly as if it had been convolved on its own).
 */
int sloika_conv1d_fwd(const float *x, const float *W, const float *bias, float *y, long ldy,
                      const int32_t *lengths, int T, int B, int Cin, int Cout, int winlen, int stride,
                      int pad_l, int pad_r, int act, void *stream);
/* Same, and additionally folds max |y| into *absmax (atomic max; the caller zeroes it first; NULL = off): lets the next
 * layer pick the cheaper operand format for outputs that are not bounded by construction (elu).
 * ldy = -1: y is written in the BLOCKED layout of sloika_gru_seq_fwd (sloika_blocked_bytes(T_out, B, Cout) bytes, 16-byte
 * aligned); raw-signal shape only (Cin = 1, winlen = 11, Cout a multiple of 4), SLOIKA_ERR_UNSUPPORTED otherwise. */
int sloika_conv1d_fwd_ex(const float *x, const float *W, const float *bias, float *y, long ldy,
                      const int32_t *lengths, int T, int B, int Cin, int Cout, int winlen, int stride,
                      int pad_l, int pad_r, int act, float *absmax, void *stream);

/*
 * FeedForward.run -- sloika/layers.py:157-158:  y = act( x . W' + bias )
 *   x: M rows of K floats (row distance ldx); W [N,K]; bias [N] (NULL = zeros); y: M rows (ldy).
 * Also the GRU input projection `vI = x iW' + b` (layers.py:1011) for all time steps at once.
 */
int sloika_linear_fwd(const float *x, long ldx, const float *W, const float *bias, float *y, long ldy,
                      long M, int K, int N, int act, void *stream);

/* Same, with an explicit kernel choice: AUTO = tcgen05 tensor-core kernel (3xTF32 split, fp32-equivalent
 * accuracy, TMA fed) when shape/alignment allow, else the fp32 SIMT kernel; SIMT / TC force one
 * (TC returns SLOIKA_ERR_UNSUPPORTED when it cannot run: ldx % 4 != 0, x not 16-byte aligned, K > 256,
 * M < 128). */
#define SLOIKA_GEMM_AUTO 0
#define SLOIKA_GEMM_SIMT 1
#define SLOIKA_GEMM_TC   2
/* Tensor-core kernel on an fp16 hi/lo split (kind::f16) instead of the tf32 one: same fp32-equivalent accuracy
 * (each operand is represented to 2^-22 relative, absolute floor 2^-25) at twice the MMA rate and half the
 * on-chip weight footprint -- but ONLY valid when |x| and |W| stay far below the fp16 range (65504): the caller
 * asserts that, e.g. x produced by a tanh / sigmoid / GRU layer.  Never chosen by SLOIKA_GEMM_AUTO. */
#define SLOIKA_GEMM_TC_F16 3
int sloika_linear_fwd_ex(const float *x, long ldx, const float *W, const float *bias, float *y, long ldy,
                         long M, int K, int N, int act, int algo, void *stream);
/* The same product with the operand format chosen ON THE DEVICE from a range report of the producer: `absmax` points
 * to one float holding max |x| (sloika_conv1d_fwd_ex writes it).  Both tensor-core forms are enqueued; the fp16-split one
 * runs if *absmax < limit, else the tf32-split one -- no host synchronisation.  Returns SLOIKA_ERR_UNSUPPORTED when the
 * tensor-core kernel cannot take the shape (use sloika_linear_fwd then). */
int sloika_linear_fwd_gated(const float *x, long ldx, const float *W, const float *bias, float *y, long ldy,
                            long M, int K, int N, int act, const float *absmax, float limit, void *stream);

/*
 * Softmax.run -- sloika/layers.py:309-314:
 *   t = x . W' + bias ; post = exp(t - max_j t) / sum_j exp(t - max_j t)     (row-wise over N)
 *   post: M rows of N floats (row distance ldp >= N).
 */
int sloika_softmax_fwd(const float *x, long ldx, const float *W, const float *bias, float *post, long ldp,
                       long M, int K, int N, void *stream);

/*
 * Softmax.run split for the fused basecall path (tensor-core kernel only).
 *   sloika_softmax_slices        number of column slices S the kernel uses for (K, N); 0 = use sloika_softmax_fwd
 *   sloika_softmax_logits_fwd    logits t = x . W' + bias (row distance ldl, ldl % 4 == 0 recommended) and
 *                                stats[(row*S + s)*2 + {0,1}] = (max_j t_j, sum_j exp(t_j - max)) over slice s.
 *                                stay_last != 0 writes the columns rotated by one: k-mer states first, the stay /
 *                                blank state (reference column 0) in column N-1 -- the layout the decoder reads.
 *   sloika_softmax_normalise_fwd logits -> posteriors in place: exp(t - M) / S_row with the row maximum M and
 *                                row sum S_row combined from the slice pairs (layers.py:311-314)
 */
int sloika_softmax_slices(int K, int N, int algo);
int sloika_softmax_logits_fwd(const float *x, long ldx, const float *W, const float *bias, float *logits, long ldl,
                              float *stats, long M, int K, int N, int stay_last, int algo, void *stream);
int sloika_softmax_normalise_fwd(float *logits, long ldl, const float *stats, int n_slices, long M, int N,
                                 void *stream);

/*
 * Gru.step scanned over time by RNN.run -- sloika/layers.py:1010-1021, :85-88 (h0 = 0); with
 * `reverse` != 0 it is Reverse(Gru).run (layers.py:1449-1450): each sequence is walked from its own
 * last valid step down to 0 and outputs land at their original time index.
 *   vI = x_t iW' + b ; vS = h sW' ; z = gate(vI[:H]+vS[:H]) ; r = gate(vI[H:2H]+vS[H:])
 *   hbar = act(vI[2H:] + (r*h) sW2') ; h' = z*h + (1-z)*hbar
 *   x: T*B rows of I floats (ldx); iW [3H,I]; sW [2H,H]; sW2 [H,H]; b [3H]; y: T*B rows of H (ldy).
 *   ws: workspace of sloika_gru_workspace_bytes(T,B,H) bytes (holds vI for all steps).
 * Outputs at steps >= lengths[b] are written as zero.  H <= 128 runs on tcgen05 with the weights in tensor memory, H <= 144 as a persistent mma.sync kernel; wider
 * layers run the scan step by step (two GEMMs + two gate kernels per step).
 */
size_t sloika_gru_workspace_bytes(int T, int B, int H);
int sloika_gru_fwd(const float *x, long ldx, const float *iW, const float *sW, const float *sW2,
                   const float *b, float *y, long ldy, void *ws, size_t ws_bytes, const int32_t *lengths,
                   int T, int B, int I, int H, int reverse, int act, int gate_act, void *stream);

/*
 * The same layer with the projection computed INSIDE the recurrence launch (csrc/gru_fused.cu): clusters of three CTAs,
 * two running the recurrence and one the projection, hand vI over through a ring of a few time steps that stays in L2,
 * so the T*B*3H projection never reaches HBM.  tanh / sigmoid, H <= 96, I <= 96, x rows 16-byte aligned (ldx % 4 == 0),
 * |x| far inside the fp16 range (outputs of tanh / sigmoid / GRU layers); SLOIKA_ERR_UNSUPPORTED otherwise (callers fall
 * back to sloika_gru_fwd).  ws: sloika_gru_fused_workspace_bytes(B, H) bytes (the rings).  It packs 32 sequences per
 * recurrence CTA: the throughput form, meant for callers that keep several batches in flight.
 */
size_t sloika_gru_fused_workspace_bytes(int B, int H);
int sloika_gru_fused_fwd(const float *x, long ldx, const float *iW, const float *sW, const float *sW2,
                         const float *b, float *y, long ldy, void *ws, size_t ws_bytes, const int32_t *lengths,
                         int T, int B, int I, int H, int reverse, int act, int gate_act, void *stream);

/*
 * The layer for an input whose range is known only on the device: `absmax` points at max |x| as written by the
 * producing kernel (sloika_conv1d_fwd_ex for an elu / linear convolution).  Both forms are enqueued and that word picks
 * one: below `limit` the fused launch above, otherwise a tf32-split projection GEMM into vI (T*B rows of pitch ldv >= 3H,
 * 16-byte aligned) followed by the recurrence kernel; the CTAs of the form that is ruled out return at once.  Same
 * shape limits as sloika_gru_fused_fwd; `seqs_in_flight` as in sloika_gru_recurrence_fwd_ex.
 */
int sloika_gru_fwd_gated(const float *x, long ldx, const float *iW, const float *sW, const float *sW2,
                         const float *b, float *y, long ldy, float *vI, long ldv, void *ws, size_t ws_bytes,
                         const int32_t *lengths, int T, int B, int I, int H, int reverse, int act, int gate_act,
                         long seqs_in_flight, const float *absmax, float limit, void *stream);

/*
 * The same layer, throughput form with the SEQUENCES on the tensor-memory lanes (csrc/gru_seq.cu): one CTA per 128
 * sequences, activations (x_t, h, r*h) as tensor-memory operands, all three weight matrices resident in shared memory, the
 * projection accumulated in place by MMAs issued a step ahead -- one launch, no workspace, no cluster.  tanh / sigmoid,
 * H <= 96, I <= 96, x rows 16-byte aligned, |x| inside the fp16 range; SLOIKA_ERR_UNSUPPORTED otherwise.  It uses
 * ceil(B / 128) SMs: meant for callers that keep several batches in flight.
 * layout bit 0: x is in the BLOCKED layout, bit 1: y is written in it -- element (t, b, f) at
 * (((t * ceil(B / 128) + b / 128) * ceil(F / 4) + f / 4) * 128 + b % 128) * 4 + f % 4, padding features zero, 16-byte aligned
 * (ldx / ldy are ignored for a blocked tensor): the layout in which a warp's 128-bit accesses are contiguous when its lanes
 * are sequences.  sloika_blocked_bytes gives the size of such a tensor, sloika_block_layout_fwd converts (to_blocked = 1:
 * row-major src with row pitch ld -> blocked dst; 0: blocked src -> row-major dst with row pitch ld).
 * sloika_gru_seq_fwd_gated is the layer for an input whose range only the device knows (absmax / limit as in
 * sloika_gru_fwd_gated) with the output in the blocked layout yb either way.  x is row-major with pitch ldx (x_blocked = 0,
 * `scratch` receives its blocked copy) or already blocked (x_blocked = 1, as sloika_conv1d_fwd_ex writes it with ldy = -1;
 * `scratch` is then a row-major buffer of pitch ldx, written only if the range check fails).  Below the limit: this kernel;
 * otherwise tf32 projection into vI, recurrence into the row-major scratch y, y -> yb.
 */
size_t sloika_blocked_bytes(int T, int B, int F);
int sloika_block_layout_fwd(const float *src, float *dst, long ld, int T, int B, int F, int to_blocked, void *stream);
int sloika_gru_seq_fwd_gated(const float *x, long ldx, int x_blocked, const float *iW, const float *sW,
                             const float *sW2, const float *b, float *yb, float *scratch, float *y, long ldy,
                             float *vI, long ldv, const int32_t *lengths, int T, int B, int I, int H, int reverse,
                             int act, int gate_act, long seqs_in_flight, const float *absmax, float limit,
                             void *stream);
int sloika_gru_seq_fwd(const float *x, long ldx, const float *iW, const float *sW, const float *sW2,
                       const float *b, float *y, long ldy, const int32_t *lengths, int T, int B, int I, int H,
                       int reverse, int act, int gate_act, int layout, void *stream);

/* The recurrence alone, given vI: T*B rows of 3H floats with row pitch ld_vi >= 3H (what sloika_gru_fwd runs
 * after the projection; a pitch that is a multiple of 4 floats keeps every row 16-byte aligned for odd H). */
int sloika_gru_recurrence_fwd(const float *vI, long ld_vi, const float *sW, const float *sW2, float *y, long ldy,
                              const int32_t *lengths, int T, int B, int H, int reverse, int act,
                              int gate_act, void *stream);
/* Same, with a scheduling hint: seqs_in_flight = number of sequences the caller keeps on the device at once (this
 * batch plus the batches it pipelines on other streams; 0 = just this call).  It only chooses how many sequences
 * share a CTA of the tensor-memory kernel (spread over the SMs for latency, packed for throughput); results do
 * not depend on it. */
int sloika_gru_recurrence_fwd_ex(const float *vI, long ld_vi, const float *sW, const float *sW2, float *y, long ldy,
                                 const int32_t *lengths, int T, int B, int H, int reverse, int act,
                                 int gate_act, long seqs_in_flight, void *stream);

/*
 * Lstm.step scanned by Lstm.run -- sloika/layers.py:677-697 (events route), given the input projection
 * vW = x iW' + b for all steps: T*B rows of 4H floats (row pitch ld_vw >= 4H).  sW [4H,H] and the projection rows are
 * in the reference's stored order (row 4*j + g = gate g of unit j: 0 update input, 1 update gate, 2 forget gate,
 * 3 output gate); peep [3,H] (update, forget, output; zeros when the layer has no peepholes).  y: T*B rows of H.
 * Outputs at steps >= lengths[b] are zero.  H <= 256.
 */
int sloika_lstm_recurrence_fwd(const float *vW, long ld_vw, const float *sW, const float *peep, float *y, long ldy,
                               const int32_t *lengths, int T, int B, int H, int reverse, int act, int gate_act,
                               void *stream);

/* Window.run -- sloika/layers.py:346-351: y[t,b,k*F+f] = x[t+k-w/2, b, f], zero outside [0, lengths[b]); w odd. */
int sloika_window_fwd(const float *x, long ldx, float *y, long ldy, const int32_t *lengths, int T, int B, int F, int w,
                      void *stream);

/*
 * olddecode.decode_profile -- sloika/olddecode.py:13-73 (non-transducer models; reached from basecall.decode_post,
 * sloika/basecall.py:47-50), one CTA per read.
 *   post   [T,B,K] posteriors over the K = 4^k k-mer states (K a multiple of 16, <= 4096), element (t,b,j) at
 *          post[t*ld_t + b*ld_b + j]; log_mode != 0: log-probabilities already (`log=True`)
 *   ltrans per-event log weights (stay, step, skip) as float64 [B][T][3] (read b at ltrans + b*ldw_b) with log 4 and
 *          log 16 ALREADY SUBTRACTED from the step and skip columns (olddecode.py:31-32), or NULL for no weights
 *   slip   probability of a slip from the best state (0 -> log(1e-10))
 *   tb_ws  sloika_olddecode_workspace_bytes(T,B,K) bytes (int32 predecessor per state and event)
 *   seq_out int32 [B,T]: ONE state per event (stays included); score_out float64 [B].
 * The recursion is float64 on float32 log-posteriors (the reference's arithmetic under NumPy >= 2).
 */
size_t sloika_olddecode_workspace_bytes(int T, int B, int K);
int sloika_olddecode_fwd(const float *post, long ld_t, long ld_b, const double *ltrans, long ldw_b, const int32_t *lengths,
                         int T, int B, int K, double slip, int log_mode, void *tb_ws, size_t ws_bytes, int32_t *seq_out,
                         double *score_out, void *stream);

/* The per-event sums of olddecode.estimate_transitions -- sloika/olddecode.py:101-109: res[ev-1] = (stay, step, skip)
 * for ev = 1..T-1 (row T-1 is left to the caller), float64 [T,3]; post [T,K] float32 with row pitch ld_t. */
int sloika_transitions_fwd(const float *post, long ld_t, int T, int K, double *res, void *stream);

/* Forward-only scoring -- bin/validate_network.py:46-54: over M rows of S posteriors (row pitch ld) with one int32
 * label each, ADDS sum_m -log post[m][label[m]] to *loss_sum (float64) and the number of rows whose first maximum is the
 * label to *ncorrect (both device words, zeroed by the caller). */
int sloika_score_fwd(const float *post, long ld, const int32_t *labels, long M, int S, double *loss_sum,
                     unsigned long long *ncorrect, void *stream);

/*
 * decode.prepare_post + decode.viterbi -- sloika/decode.py:21-36, :39-93 (called from
 * sloika/basecall.py:44-46), batched: one CTA per read.
 *   post: [T,B,S] with S = nbase^klen + 1, element (t,b,s) at post[t*ld_t + b*ld_b + s]
 *   mode: SLOIKA_VIT_POST     post are probabilities; the kernel applies
 *                                 lpost = log( (min_prob + (1-min_prob)*post) + 1e-10 )   (float32)
 *         SLOIKA_VIT_LOG      post are already log-probabilities (`log=True`, decode.py:56)
 *   tb_ws: workspace of sloika_viterbi_workspace_bytes(T,B,nbase,klen) bytes (traceback: uint16 per quad of states
 *          for nbase 4, klen 5; uint8 per state otherwise)
 *   path_out int32 [B,T]: k-mer states of the best path with stays removed, left-aligned;
 *   path_len int32 [B]; score_out float32 [B] (= max_j v_T[j]).
 * skip_pen and min_prob are the reference's python floats (double); they are cast to float32 where
 * NumPy casts them.  Tie rules follow decode.py exactly: first maximum over predecessors, skip beats step on a tie,
 * stay beats move on a tie.  A sequence with lengths[b] < 1 yields path_len 0 and score 0.
 */
#define SLOIKA_VIT_POST 0
#define SLOIKA_VIT_LOG  1
size_t sloika_viterbi_workspace_bytes(int T, int B, int nbase, int klen);
int sloika_viterbi_fwd(const float *post, long ld_t, long ld_b, const int32_t *lengths, int T, int B,
                       int nbase, int klen, double skip_pen, double min_prob, int mode, void *tb_ws,
                       size_t ws_bytes, int32_t *path_out, int32_t *path_len, float *score_out,
                       void *stream);

/* sloika_softmax_logits_fwd (fp16-split form) with x in the BLOCKED layout of sloika_gru_seq_fwd; M = T * B with B a multiple
 * of 128 (whole blocks).  The last GRU layer's output enters the logits GEMM without a layout conversion. */
int sloika_softmax_logits_blocked_fwd(const float *xb, const float *W, const float *bias, float *logits, long ldl,
                                      float *stats, long M, int K, int N, int stay_last, void *stream);

/*
 * Same decode fed by the un-normalised network output of sloika_softmax_logits_fwd(stay_last = 1):
 * the softmax division (layers.py:314), the min_prob floor (decode.py:36) and the log (decode.py:56) are
 * applied inside the kernel, so the posterior matrix never exists in HBM.  nbase = 4, klen = 5 only.
 *   logits: element (t, b, j) at logits[t*ld_t + b*ld_b + j], j < 1024 k-mer states, j = 1024 stay;
 *   stats: [T*B][n_slices] float pairs as written by sloika_softmax_logits_fwd (rows in (t, b) order: ld_t = B*ld_b);
 *   they are first combined into one float per row, which is WRITTEN INTO THE ROW'S FIRST PADDING COLUMN
 *   (logits[t*ld_t + b*ld_b + 1025]; rows are 16-byte aligned, so ld_b >= 1028) -- the decoder then stages k-mer
 *   logits, stay logit and statistic of an event with one contiguous bulk copy (two launches).
 */
int sloika_viterbi_logits_fwd(float *logits, long ld_t, long ld_b, const float *stats, int n_slices,
                              const int32_t *lengths, int T, int B, int nbase, int klen, double skip_pen,
                              double min_prob, void *tb_ws, size_t ws_bytes, int32_t *path_out, int32_t *path_len,
                              float *score_out, void *stream);

/*
 * Remap decode: transducer.map_to_sequence -- sloika/transducer.py:14-73 -- with its
 * native helper viterbi_helpers.slip_update -- sloika/viterbi_helpers.pyx:12-35 -- batched, one CTA per read
 * (reference callers: sloika/tools/chunkify_raw.py:262-274, sloika/batch.py:141-155, one read per call).
 *   trans: [T,B,nstate] transducer posteriors, element (t,b,s) at trans[t*ld_t + b*ld_b + s], column 0 = stay;
 *          is_log != 0: already log-scaled (`log=True`), else the kernel takes logf (`log=False`)
 *   nev int32 [B] (NULL = T): events per read; seq int32 [B][ld_seq]: state columns of the reference sequence;
 *   npos int32 [B] (NULL = P): positions per read, 3 <= npos <= P <= 65535 (and P small enough for the on-chip
 *          score vectors: 24 P + 8 nstate bytes <= 200 KB, else SLOIKA_ERR_UNSUPPORTED)
 *   slip / has_slip: slip penalty >= 0; has_slip == 0 reproduces the reference's slip=None, which runs the slip
 *          move with a NaN penalty (np.float32(None)) and returns a NaN score
 *   prior_initial / prior_final: float64 [B][ld_prior] or NULL (util.geometric_prior's dtype; added in double)
 *   ws: sloika_remap_workspace_bytes(T,B,P) bytes (uint16 traceback)
 *   path_out int32 [B][T]: position of every event (first nev[b] entries); score_out float32 [B]
 * sloika_slip_update_fwd is slip_update alone (n >= 3; from_pos is int64 like the reference's np.int).
 */
size_t sloika_remap_workspace_bytes(int T, int B, int P);
int sloika_remap_fwd(const float *trans, long ld_t, long ld_b, const int32_t *nev, int T, int B, int nstate,
                     const int32_t *seq, long ld_seq, const int32_t *npos, int P, double slip, int has_slip,
                     const double *prior_initial, const double *prior_final, long ld_prior, int is_log,
                     void *ws, size_t ws_bytes, int32_t *path_out, float *score_out, void *stream);
int sloika_slip_update_fwd(const float *x, int n, float slip, float *from_score, long long *from_pos, void *stream);

/*
 * Best path -> base sequence: bio.kmers_to_sequence -- sloika/bio.py:228-237
 * (max_overlap :160-178, reduce_kmers :208-225) -- as SeqPrinter.write applies it to a read's k-mer states
 * (sloika/basecall.py:141-149), batched, one CTA per read.
 *   path int32 [B][ld_path]: k-mer states (sloika_viterbi_*'s path_out); path_len int32 [B]
 *   always_move != 0: a k-mer followed by itself is a move of klen letters (transducer models), else a stay
 *   alphabet: nbase device bytes ("ACGT"); out: [B][ld_out] bytes, ld_out >= klen * max(path_len); out_len int32 [B]
 */
int sloika_path_to_bases_fwd(const int32_t *path, long ld_path, const int32_t *path_len, int B, int klen, int nbase,
                             int always_move, const char *alphabet, char *out, long ld_out, int32_t *out_len,
                             void *stream);

/*
 * Raw-signal pre-processing: what basecall.raw_worker does to a read before the network --
 * sloika/basecall.py:111-118 -- batch.trim_open_pore (sloika/batch.py:194-220), util.trim_array (util.py:94-99),
 * (signal - median) / mad in float64 (sloika/maths.py:4-45), cast to float32 -- batched, one CTA per read, exact order
 * statistics: the float32 result is bit-identical to the NumPy path.
 *   signals: the reads' float64 samples back to back; offsets long [B+1]; max_len >= longest read
 *   window: trim_open_pore's window (100 in the reference; 2..128), var_method 'mad'
 *   ws: sloika_prepare_workspace_bytes(max_len, B, window) bytes
 *   out: float32 [Tmax][ld_t] time-major batch, column b = read b, zero padded (Tmax >= max_len keeps everything)
 *   out_len int32 [B]: samples kept; 0 = nothing left after trimming ("Read too short"); -1 = no window exceeds the
 *            threshold (the reference raises there: empty `is_read`)
 */
size_t sloika_prepare_workspace_bytes(long max_len, int B, int window);
int sloika_prepare_signal_fwd(const double *signals, const long *offsets, int B, long max_len, int trim_start,
                              int trim_end, double open_pore_fraction, int window, void *ws, size_t ws_bytes,
                              float *out, long ld_t, int Tmax, int32_t *out_len, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SLOIKA_B200_H */
