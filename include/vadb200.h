/*
 * vadb200.h -- C ABI of libvadb200.so, the B200 (sm_100a) implementation of the
 * Self-Attentive VAD inference hot path of voithru/voice-activity-detection.
 *
 * The reference has no native layer (it is pure Python/PyTorch), so there is no existing
 * FFI to mirror: each entry point below names the reference *Python* interface it
 * replaces (paths relative to the reference repository root).  Plain pointers and sizes
 * only; no torch types.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, a negative VADB_E_* code on failure;
 *     vadb_last_error(h) gives the message (the Python wrapper raises RuntimeError).
 *   - "dev" pointers are CUDA device pointers on the handle's device; `stream` is a
 *     cudaStream_t passed as void* (0 = legacy default stream).  Calls taking a stream
 *     are asynchronous on it and never synchronise internally, except where they must
 *     grow the handle-owned workspace (first call with a larger B*T; use vadb_reserve
 *     to take that out of the steady state).
 *   - the caller owns inputs/outputs; the library owns weights, the positional-encoding
 *     table and the workspace inside the handle.  One handle per device; a handle is not
 *     thread-safe, distinct handles are independent.  All calls on one handle share its workspace:
 *     stream-taking calls must be issued on ONE stream at a time (or be ordered by the caller), the
 *     host-buffer calls (vadb_*_host*) run on the library's own streams and order themselves --
 *     a blocking host call first completes any asynchronous one that is still in flight.
 *   - there is NO CPU fallback: without a CUDA device vadb_create fails.
 */
#ifndef VADB200_H_
#define VADB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vadb_handle vadb_handle;

enum { VADB_F32 = 0, VADB_BF16 = 1 };           /* dtypes (input tensors / compute mode) */

enum {
  VADB_OK = 0,
  VADB_E_INVALID = -1,     /* bad argument / unsupported configuration */
  VADB_E_CUDA = -2,        /* CUDA runtime or driver error */
  VADB_E_STATE = -3,       /* e.g. forward before weights were loaded */
  VADB_E_NOMEM = -4
};

/* Model hyper-parameters.  Replaces the arguments of
 * vad/models/self_attention.py:7 SelfAttentiveVAD(feature_size, num_layers, d_model, dropout)
 * as selected by vad/models/model_factory.py:42-48 from the checkpoint's config
 * (vad/configs/model_config.py:8-35).  dropout is an eval-time identity and is dropped.
 * The kernels are specialised for d_model == 128 (d_ff == 512, n_heads == 1), the only
 * values the reference instantiates; anything else is VADB_E_INVALID. */
typedef struct {
  int32_t feature_size;    /* F: n_mels (80 in the reference checkpoint, 64 in BASELINE) */
  int32_t num_layers;      /* L */
  int32_t d_model;         /* must be 128 */
  int32_t compute_dtype;   /* VADB_F32: fp32 CUDA-core path (<=1e-3 parity);
                              VADB_BF16: bf16 tensor-core operands, fp32 accumulate/residual/LN/softmax */
} vadb_config;

/* ---- lifecycle: replaces VADFromScratchPredictor.from_checkpoint's model construction +
 *      load_state_dict + .to(device)  (vad/predictor.py:264-280) ---- */
int vadb_create(vadb_handle** out, const vadb_config* cfg, int device);
void vadb_destroy(vadb_handle* h);
const char* vadb_last_error(const vadb_handle* h);   /* h may be NULL: last create error */

/* Number of fp32 elements of the packed weight blob for a config, and the element offset
 * of a named tensor inside it.  Order = the reference state_dict order
 * (vad/predictor.py:278 load_state_dict; names listed in SURVEY.md section 8 a1), each
 * tensor row-major in nn.Linear [out,in] layout. */
size_t vadb_weight_count(const vadb_config* cfg);

/* Load the packed fp32 blob (host pointer when on_device == 0, device pointer otherwise --
 * e.g. the buffer a rank received from the one-off NCCL weight broadcast) and derive the
 * kernel-side copies (bf16 / fused QKV / classifier).  Synchronises `stream` before return. */
int vadb_load_weights(vadb_handle* h, const float* blob, size_t count, int on_device, void* stream);

/* Multi-GPU load: ONE NCCL broadcast of the packed blob from rank `root` (which has called
 * vadb_load_weights) into every other rank's handle, followed by the same derivation of the kernel-side
 * copies.  This is the only collective of the whole path (there is none in steady state); it replaces the
 * reference's training-time nn.DataParallel replication (vad/training/trainer.py:115-116) for inference.
 *  nccl_comm  an ncclComm_t (passed as void*) whose ranks each own one handle on their own device
 * NCCL is resolved at run time from the calling process (dlsym, else dlopen("libnccl.so.2")): the
 * library has no link-time dependency on it.  Collective: every rank of the communicator must call it. */
int vadb_broadcast_weights(vadb_handle* h, void* nccl_comm, int root, void* stream);

/* Pre-size the workspace (and positional-encoding table) for calls up to B clips x T frames. */
int vadb_reserve(vadb_handle* h, int B, int T);

/* ---- hot path: replaces SelfAttentiveVAD.forward (vad/models/self_attention.py:23-28)
 *      + the caller's softmax(...)[...,1] (vad/predictor.py:225,257-258).
 *  x        dev [B,T,F] contiguous, x_dtype VADB_F32 or VADB_BF16
 *  lengths  dev int32 [B] or NULL; key-padding mask j >= lengths[b] as built by
 *           mask_from_lengths (vad/modeling/transformer.py:432-447, :319-325)
 *  prob     dev float [B,T]   P(speech) = softmax(logp)[1]           (or NULL)
 *  logp     dev float [B,T,2] the module's log_softmax output         (or NULL)        */
int vadb_forward(vadb_handle* h, const void* x, int x_dtype, const int32_t* lengths,
                 int B, int T, float* prob, float* logp, void* stream);

/* Mixed-length batches (BASELINE configs[4]) without the padding work: the same call with the lengths in HOST
 * memory.  Clips are grouped by ceil(length / 128) * 128 processed frames; every group runs as its own dense batch
 * with its own key-padding mask, so per-frame kernels and attention see only the frames each clip needs instead of
 * the batch maximum T (vad/modeling/transformer.py:432-447 builds one [B, T] mask for the padded batch).  Valid
 * frames (t < lengths[b]) get exactly the results of vadb_forward; frames past a clip's processed length, which
 * the reference computes from the padding, are returned as 0.  The groups run concurrently: two library-owned side
 * streams fork from `stream` after the call's prologue and rejoin it before the call returns, so for the caller the
 * call is ordered on `stream` like any other.
 *  x             dev [B,T,F] padded batch        lengths_host  HOST int32 [B]        prob / logp  dev, as above */
int vadb_forward_ragged(vadb_handle* h, const void* x, int x_dtype, const int32_t* lengths_host,
                        int B, int T, float* prob, float* logp, void* stream);

/* Same call with HOST buffers (pageable or pinned): the library stages them through its
 * pinned buffers, copies H2D, runs the forward and copies the results D2H, on its own
 * stream, and returns when the outputs are in host memory.  This is the end-to-end
 * entry the reference-facing wrapper uses for numpy / CPU-tensor inputs
 * (vad/predictor.py:223 .to(device), :247 .cpu()).
 *  x        host [B,T,F], x_dtype VADB_F32 or VADB_BF16 (bf16 features halve the upload; the bf16
 *           compute mode rounds the features to bf16/tf32 operands anyway)
 *  logp     any alignment is accepted (B*T may be odd) */
int vadb_forward_host(vadb_handle* h, const void* x, int x_dtype, const int32_t* lengths, int B, int T,
                      float* prob, float* logp);

/* Asynchronous form for streaming many batches: everything (H2D, forward, D2H) is only enqueued on the
 * library's streams and the call returns at once, so the upload of the next batch overlaps the compute of
 * this one (throughput = max(upload, compute) instead of their pipelined sum).  x, prob and logp must be
 * PINNED host buffers and stay valid/untouched until vadb_host_wait(h, ticket) has returned.  At most four
 * calls may be outstanding: a fifth call before the oldest ticket was waited for fails with VADB_E_STATE
 * (nothing is enqueued).  Calls complete in ticket order; vadb_host_wait never reports success without
 * having synchronised on the ticket's (or a later) completion event.  Same reference interface as
 * vadb_forward_host, called in a loop over batches (vad/predictor.py:182-258 iterates chunks of windows
 * the same way, synchronously). */
int vadb_forward_host_async(vadb_handle* h, const void* x, int x_dtype, const int32_t* lengths, int B, int T,
                            float* prob, float* logp, long* ticket);
int vadb_host_wait(vadb_handle* h, long ticket);

/* ---- replaces VADFromScratchPredictor.predict_probabilities from the feature matrix on
 *      (vad/predictor.py:169-262): window gather (:180-220), batched forward (:221-225)
 *      and boosted aggregation incl. the 0.5 fill of never-written slots (:238-258).
 *  feat       dev float [L,F]  log-mel frames
 *  probs_LW   dev float [L,W]  W = 2*(half-1)/jump + 3           (or NULL)
 *  mean_L     dev float [L]    probs.mean(axis=1) (vad/predictor.py:95)  (or NULL)     */
int vadb_predict_probabilities(vadb_handle* h, const float* feat, int L, int half, int jump,
                               float* probs_LW, float* mean_L, void* stream);
int vadb_predict_probabilities_host(vadb_handle* h, const float* feat, int L, int half,
                                    int jump, float* probs_LW, float* mean_L);

/* ---- log-mel front end on the device: replaces FeatureExtractor.extract_with_postprocessing
 *      (vad/acoustics/feature_extractor.py:71-80) for the log-mel transform
 *      (vad/acoustics/transforms/log_mel_spectrogram.py:19-32:
 *      np.log(librosa.feature.melspectrogram(y, sr, n_mels, n_fft, hop_length, win_length) + 1e-6),
 *      librosa 0.8.0 defaults: periodic hann zero-padded to n_fft, center=True/reflect, power 2, Slaney mel).
 *      librosa is not vendored by the reference: parity is pinned to the NumPy restatement of that
 *      algorithm (oracle/logmel_oracle.py), not to librosa itself.  n_fft: power of two in [32, 4096].
 *  audio   dev float [n_samples] mono PCM in [-1, 1]
 *  feat    dev float [vadb_logmel_frames(n_samples, hop), n_mels]                                   */
long vadb_logmel_frames(long n_samples, int hop);          /* 1 + n_samples / hop */
int vadb_logmel(vadb_handle* h, const float* audio, long n_samples, int sample_rate, int n_fft, int hop,
                int win, int n_mels, float* feat, void* stream);
/* Host-only helper (no device needed): the dense [n_mels, n_fft/2+1] float32 mel filterbank and the
 * [n_fft] float64 window the kernel uses; either output may be NULL. */
int vadb_logmel_tables(int sample_rate, int n_fft, int win, int n_mels, float* fb_dense, double* window);

/* ---- replaces VADFromScratchPredictor.predict_probabilities from the AUDIO on
 *      (vad/predictor.py:159-262 including :160 feature extraction): H2D of the samples, log-mel,
 *      window gather, forward, boosted aggregation, D2H -- one call, nothing but PCM crosses PCIe.
 *  audio     host float [n_samples]; n_mels is the handle's feature_size
 *  feat_out  host float [L, F] log-mel frames, L = vadb_logmel_frames(n_samples, hop)   (or NULL)
 *  probs_LW  host float [L, W]                                                           (or NULL)
 *  mean_L    host float [L]                                                              (or NULL) */
int vadb_predict_audio_host(vadb_handle* h, const float* audio, long n_samples, int sample_rate, int n_fft,
                            int hop, int win, int half, int jump, float* feat_out, float* probs_LW,
                            float* mean_L);

/* ---- individual stages, exported for kernel-level parity tests and the roofline bench ---- */
/* Fused scaled-dot-product attention over the frame axis:
 * O = softmax(Q K^T / sqrt(128) [+ key padding mask]) V   (vad/modeling/transformer.py:351-363,
 * :319-346), one head, d_head = 128.  q,k,v,o dev [B,T,128] contiguous, dtype VADB_F32 or
 * VADB_BF16 (bf16 -> tcgen05 kernel; fp32 -> CUDA-core kernel). */
int vadb_attention(vadb_handle* h, const void* q, const void* k, const void* v, void* o,
                   int dtype, const int32_t* lengths, int B, int T, void* stream);

/* Positional-encoding table PE[t, :]/sqrt(d) the library adds in the front end
 * (vad/modeling/transformer.py:392-414); copies T*128 floats to a host buffer. */
int vadb_positional_table(vadb_handle* h, int T, float* out_host);

/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
int64_t vadb_launch_count(const vadb_handle* h);

/* Library / build identification, e.g. "vadb200 0.1 sm_100a". */
const char* vadb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VADB200_H_ */
