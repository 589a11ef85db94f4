// C ABI of libvadb200.so (see include/vadb200.h): handle, weights, workspace and the
// orchestration of the forward pass / predictor window path on a caller-provided stream.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "vadb_common.cuh"

namespace vadb {

BlobLayout make_layout(int F, int L) {
  // state_dict order of SelfAttentiveVAD (vad/models/self_attention.py:12-21,
  // vad/modeling/transformer.py:40-47,245-248,230,370-375,22)
  BlobLayout lay;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += n; return r; };
  lay.w_in = take((size_t)D * F);
  lay.b_in = take(D);
  lay.layers.resize(L);
  for (int l = 0; l < L; ++l) {
    LayerOffsets& lo = lay.layers[l];
    lo.wq = take((size_t)D * D); lo.bq = take(D);
    lo.wk = take((size_t)D * D); lo.bk = take(D);
    lo.wv = take((size_t)D * D); lo.bv = take(D);
    lo.wo = take((size_t)D * D); lo.bo = take(D);
    lo.ln1_g = take(D); lo.ln1_b = take(D);
    lo.w1 = take((size_t)DFF * D); lo.b1 = take(DFF);
    lo.w2 = take((size_t)D * DFF); lo.b2 = take(D);
    lo.ln2_g = take(D); lo.ln2_b = take(D);
  }
  lay.lnf_g = take(D); lay.lnf_b = take(D);
  lay.wc = take(2 * D); lay.bc = take(2);
  lay.total = o;
  return lay;
}

}  // namespace vadb

using namespace vadb;

struct vadb_handle {
  vadb_config cfg;
  int device = 0;
  BlobLayout lay;
  bool loaded = false;
  std::string err;
  int64_t launches = 0;

  float* w32 = nullptr;      // packed fp32 blob (device)
  float* wqkv = nullptr;     // [L][384*128] fused Q|K|V weight rows
  float* bqkv = nullptr;     // [L][384]
  bf16* wqkv_bf = nullptr;   // bf16 copies for the tensor-core kernels
  bf16* wo_bf = nullptr;     // [L][128*128]
  bf16* w1_bf = nullptr;     // [L][512*128]
  bf16* w2_bf = nullptr;     // [L][128*512]
  bf16* win_bf = nullptr;    // [128, 128] front-end weight, columns >= F zero (tensor-core front end)
  unsigned char* wtail = nullptr;   // [L] packed weight blocks of the fused layer-tail kernel (k_tail_tc.cu)
  float* tail_aux = nullptr;        // [L] folded biases / classifier constants of the same kernel
  unsigned char* head_bf16 = nullptr;   // head of the model (front end + layer 0's q/k/v in the tail kernel's head mode):
  unsigned char* head_tf32 = nullptr;   //   4 weight blocks for bf16 features / for fp32 features as tf32 (F <= 64)
  float* head_aux = nullptr;            //   folded q|k|v biases of layer 0
  float* pe_tiled = nullptr;            // PE / sqrt(d) of pe_tiled_T positions in the tiled layout (head mode)
  int pe_tiled_T = 0;
  bool qkv_ready = false;               // ws_q/k/v already hold layer 0's attention inputs (written by the head kernel)

  float* pe = nullptr;       // [pe_T, 128] = PE / sqrt(d)
  int pe_T = 0;
  int num_sms = 148;
  bool aln_valid = false;    // ws_aln holds LayerNorm(h) for the next Q/K/V GEMM

  // workspace, sized in frames
  size_t cap_frames = 0;
  float* ws_h = nullptr;
  void* ws_q = nullptr; void* ws_k = nullptr; void* ws_v = nullptr; void* ws_o = nullptr;
  void* ws_hid = nullptr;
  bf16* ws_aln = nullptr;    // [frames,128] bf16 LayerNorm(h) handed from kernel to kernel (bf16 mode)
  float* ws_prob = nullptr;

  // host-call staging
  cudaStream_t own_stream = nullptr, own_stream2 = nullptr, own_stream3 = nullptr;     // upload, compute, download
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_out[2] = {nullptr, nullptr};        // downloads of the host call that used device output half i have completed
  long out_seq = 0;                                  // host calls issued so far (selects the output half)
  size_t last_out_half = 0;                          // floats per output half of the previous host call
  cudaEvent_t ev_call[4] = {nullptr, nullptr, nullptr, nullptr};   // completion of the last four asynchronous host calls
  long call_seq = 0, chunk_seq = 0;
  long waited_upto = -1;     // every asynchronous host call with ticket <= waited_upto is known to have completed
  size_t last_chunk_in = 0;
  void* pin_in = nullptr; size_t pin_in_bytes = 0;
  void* pin_out = nullptr; size_t pin_out_bytes = 0;
  void* dev_in = nullptr; size_t dev_in_bytes = 0;
  void* dev_out = nullptr; size_t dev_out_bytes = 0;
  int32_t* dev_len = nullptr; size_t dev_len_n = 0;
  void* win_prob = nullptr; size_t win_prob_bytes = 0;   // [n, W] per-window probabilities
  void* win_proj = nullptr; size_t win_proj_bytes = 0;   // [L, 128] fp32 input projection of the whole clip
  // length-bucketed forward (vadb_forward_ragged): gathered inputs / outputs of one bucket, clip index lists
  void* bk_x = nullptr; size_t bk_x_bytes = 0;
  void* bk_out = nullptr; size_t bk_out_bytes = 0;
  void* bk_idx = nullptr; size_t bk_idx_bytes = 0;
  cudaStream_t bk_stream[2] = {nullptr, nullptr};      // side streams: length buckets run concurrently
  cudaEvent_t bk_fork = nullptr, bk_join[2] = {nullptr, nullptr};

  // log-mel front end (k_logmel.cu): device tables for one (sr, n_fft, win, n_mels) at a time
  int lm_sr = 0, lm_nfft = 0, lm_win = 0, lm_nmels = 0;
  double* lm_window = nullptr; double* lm_twiddle = nullptr;
  int* lm_meta = nullptr;      // [3][n_mels]: first bin, number of bins, offset into lm_w
  float* lm_w = nullptr;       // packed non-zero runs of the mel filterbank
  void* lm_audio = nullptr; size_t lm_audio_bytes = 0;   // device copy of the audio (host entry point)
  void* lm_feat = nullptr; size_t lm_feat_bytes = 0;     // [frames, n_mels] features (host entry point)
};

namespace {

std::string g_create_err;

constexpr size_t MAX_FRAMES_PER_PASS = (size_t)1 << 20;

int fail(vadb_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}

#define CU_TRY(h, expr)                                                                      \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return fail((h), VADB_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

bool is_bf16_mode(const vadb_handle* h) { return h->cfg.compute_dtype == VADB_BF16; }

// bf16 mode: out-projection + FFN + next layer's Q/K/V (or the classifier) in one kernel per layer, residual
// stream in the tiled layout (k_tail_tc.cu).  VADB_FUSE_TAIL=0 keeps the round-1 kernel sequence (A/B runs).
bool fuse_tail(const vadb_handle* h) {
  static const bool on = !(getenv("VADB_FUSE_TAIL") && atoi(getenv("VADB_FUSE_TAIL")) == 0);
  return on && is_bf16_mode(h);
}

template <typename T>
void free_dev(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

int ensure_pe(vadb_handle* h, int T) {
  if (T <= h->pe_T) return VADB_OK;
  int newT = std::max(T, std::max(16, h->pe_T * 2));
  // vad/modeling/transformer.py:403-414, then "/ self.scale" (:390, :401); all in fp32
  std::vector<float> tab((size_t)newT * D);
  const float c = (float)(-(log(10000.0) / (double)D));
  const float scale = (float)sqrt((double)D);
  for (int i = 0; i < D / 2; ++i) {
    const float arg = (float)(2 * i) * c;
    const float div = (float)exp((double)arg);
    for (int t = 0; t < newT; ++t) {
      const float ang = (float)t * div;
      tab[(size_t)t * D + 2 * i] = (float)sin((double)ang) / scale;
      tab[(size_t)t * D + 2 * i + 1] = (float)cos((double)ang) / scale;
    }
  }
  // the table may be in use by work already queued on a caller stream: sync before replacing
  CU_TRY(h, cudaDeviceSynchronize());
  free_dev(h->pe);
  CU_TRY(h, cudaMalloc(&h->pe, tab.size() * sizeof(float)));
  CU_TRY(h, cudaMemcpy(h->pe, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
  h->pe_T = newT;
  return VADB_OK;
}

// PE / sqrt(d) for positions [0, T) in the tiled layout of the residual stream ([T/128 tiles][32 column quads][128 rows]
// [4 floats]): the head kernel bulk-loads one 64 KB tile of it per 128 frames
int ensure_pe_tiled(vadb_handle* h, int T) {
  if (T <= h->pe_tiled_T) return VADB_OK;     // tile j of the table is positions 128 j ..: a longer table serves every shorter T
  if (T % 128 || T > h->pe_T) return fail(h, VADB_E_INVALID, "internal: tiled positional table needs T % 128 == 0");
  std::vector<float> rows((size_t)T * D), tiled((size_t)T * D);
  CU_TRY(h, cudaDeviceSynchronize());
  CU_TRY(h, cudaMemcpy(rows.data(), h->pe, rows.size() * sizeof(float), cudaMemcpyDeviceToHost));
  for (int t = 0; t < T; ++t)
    for (int c = 0; c < D; ++c)
      tiled[(size_t)(t >> 7) * 16384 + (size_t)(c >> 2) * 512 + (size_t)(t & 127) * 4 + (c & 3)] = rows[(size_t)t * D + c];
  free_dev(h->pe_tiled);
  h->pe_tiled_T = 0;
  CU_TRY(h, cudaMalloc(&h->pe_tiled, tiled.size() * sizeof(float)));
  CU_TRY(h, cudaMemcpy(h->pe_tiled, tiled.data(), tiled.size() * sizeof(float), cudaMemcpyHostToDevice));
  h->pe_tiled_T = T;
  return VADB_OK;
}

int ensure_workspace(vadb_handle* h, size_t frames) {
  if (frames <= h->cap_frames) return VADB_OK;
  size_t cap = std::max(frames, h->cap_frames + h->cap_frames / 2);
  CU_TRY(h, cudaDeviceSynchronize());
  free_dev(h->ws_h); free_dev(h->ws_q); free_dev(h->ws_k); free_dev(h->ws_v); free_dev(h->ws_o);
  free_dev(h->ws_hid); free_dev(h->ws_prob); free_dev(h->ws_aln);
  h->cap_frames = 0;
  const size_t act = is_bf16_mode(h) ? sizeof(bf16) : sizeof(float);
  CU_TRY(h, cudaMalloc(&h->ws_h, ((cap + 127) / 128 * 128) * D * sizeof(float)));   // whole 128-row tiles (tiled layout)
  CU_TRY(h, cudaMalloc(&h->ws_q, cap * D * act));
  CU_TRY(h, cudaMalloc(&h->ws_k, cap * D * act));
  CU_TRY(h, cudaMalloc(&h->ws_v, cap * D * act));
  CU_TRY(h, cudaMalloc(&h->ws_o, cap * D * act));
  CU_TRY(h, cudaMalloc(&h->ws_hid, cap * DFF * act));
  CU_TRY(h, cudaMalloc(&h->ws_prob, cap * sizeof(float)));
  CU_TRY(h, cudaMalloc(&h->ws_aln, cap * D * sizeof(bf16)));
  h->cap_frames = cap;
  return VADB_OK;
}

// clips per pass so that one pass stays within the workspace cap
int clips_per_pass(int B, int T) {
  size_t c = MAX_FRAMES_PER_PASS / (size_t)std::max(T, 1);
  if (c < 1) c = 1;
  return (int)std::min<size_t>(c, (size_t)B);
}

int gemm(vadb_handle* h, const GemmArgs& a, cudaStream_t s) {
  cudaError_t e = launch_gemm_f32(a, s);
  if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("gemm_f32: ") + cudaGetErrorString(e));
  h->launches++;
  return VADB_OK;
}

int gemm_tc(vadb_handle* h, const GemmTcArgs& a, cudaStream_t s) {
  std::string err;
  cudaError_t e = launch_gemm_tc(a, h->num_sms, s, &err);
  if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("gemm_tc: ") + cudaGetErrorString(e) + " " + err);
  h->launches++;
  return VADB_OK;
}

int attention(vadb_handle* h, const void* q, const void* k, const void* v, void* o, int dtype,
              const int32_t* lengths, int B, int T, cudaStream_t s) {
  static const bool small_ok = !(getenv("VADB_ATTN_SMALL") && atoi(getenv("VADB_ATTN_SMALL")) == 0);
  if (dtype == VADB_BF16 && small_ok && attn_small_supported(T)) {
    // the reference Predictor's 7-frame windows: a 128-row tcgen05 tile would be 94 % padding
    cudaError_t e = launch_attn_small((const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)o, lengths, B, T, h->num_sms, s);
    if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("attn_small: ") + cudaGetErrorString(e));
  } else if (dtype == VADB_BF16) {
    std::string err;
    cudaError_t e = launch_attn_tc((const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)o, lengths,
                                   B, T, h->num_sms, s, &err);
    if (e != cudaSuccess)
      return fail(h, VADB_E_CUDA, std::string("attn_tc: ") + cudaGetErrorString(e) + " " + err);
  } else {
    cudaError_t e = launch_attn_f32((const float*)q, (const float*)k, (const float*)v, (float*)o,
                                    lengths, B, T, s);
    if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("attn_f32: ") + cudaGetErrorString(e));
  }
  h->launches++;
  return VADB_OK;
}

// encoder layers + classifier over `Bc` clips of `T` frames whose front-end output is in ws_h
int run_encoder(vadb_handle* h, const int32_t* lengths, int Bc, int T, float* prob, float* logp,
                cudaStream_t s) {
  const int M = Bc * T;
  const float* w = h->w32;
  const bool bf = is_bf16_mode(h);
  static const bool fuse_cls = !(getenv("VADB_FUSE_CLS") && atoi(getenv("VADB_FUSE_CLS")) == 0);
  if (fuse_tail(h)) {
    // bf16 mode, fused: [Q/K/V of layer 0] then per layer [attention, layer tail]; the tail kernel of layer l
    // also produces q,k,v of layer l+1, the last one the probabilities (2 + 2L launches per forward)
    const int L = h->cfg.num_layers;
    int rc;
    if (!h->qkv_ready) {
      const LayerOffsets& lo = h->lay.layers[0];
      GemmTcArgs g = {};
      g.M = M; g.N = 3 * D; g.K = D; g.w_bf16 = h->wqkv_bf;
      if (h->aln_valid) g.a_bf16 = h->ws_aln;
      else return fail(h, VADB_E_STATE, "internal: LayerNorm operand of layer 0 missing");
      (void)lo;
      g.bias = h->bqkv;
      g.out[0] = h->ws_q; g.out[1] = h->ws_k; g.out[2] = h->ws_v;
      if ((rc = gemm_tc(h, g, s))) return rc;
    }
    for (int l = 0; l < L; ++l) {
      const LayerOffsets& lo = h->lay.layers[l];
      if ((rc = attention(h, h->ws_q, h->ws_k, h->ws_v, h->ws_o, VADB_BF16, lengths, Bc, T, s))) return rc;
      TailTcArgs t = {};
      t.M = M; t.wpack = h->wtail + (size_t)l * tail_pack_bytes(); t.aux = h->tail_aux + (size_t)l * tail_aux_floats();
      t.o = (const bf16*)h->ws_o; t.h = h->ws_h; t.bo = w + lo.bo; t.b2 = w + lo.b2;
      if (l + 1 < L) { t.q = (bf16*)h->ws_q; t.k = (bf16*)h->ws_k; t.v = (bf16*)h->ws_v; }
      else { t.prob = prob; t.logp = logp; }
      std::string err;
      cudaError_t e = launch_tail_tc(t, h->num_sms, s, &err);
      if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("tail_tc: ") + cudaGetErrorString(e) + " " + err);
      h->launches++;
    }
    h->aln_valid = false;
    h->qkv_ready = false;
    return VADB_OK;
  }
  for (int l = 0; l < h->cfg.num_layers; ++l) {
    const LayerOffsets& lo = h->lay.layers[l];
    int rc;
    if (bf) {
      // bf16 mode: every Linear on the tensor cores (k_gemm_tc.cu / k_ffn_tc.cu), attention in k_attn_tc.cu.
      // Each kernel that produces the residual stream h also emits LayerNorm(h) of the NEXT sublayer in
      // bf16 (ws_aln), so every GEMM reads its A operand by TMA.  ws_aln holds LN1_l(h) on entry
      // (front end / previous layer); it is rebuilt here when the front end took the fp32 fallback.
      const bool have_aln = h->aln_valid;
      {  // q,k,v = LN1(h) W^T + b                (transformer.py:235-236, :281-284)
        GemmTcArgs g = {};
        g.M = M; g.N = 3 * D; g.K = D; g.w_bf16 = h->wqkv_bf + (size_t)l * 3 * D * D;
        if (have_aln) g.a_bf16 = h->ws_aln;
        else { g.a_f32 = h->ws_h; g.ln_g = w + lo.ln1_g; g.ln_b = w + lo.ln1_b; }
        g.bias = h->bqkv + (size_t)l * 3 * D;
        g.out[0] = h->ws_q; g.out[1] = h->ws_k; g.out[2] = h->ws_v;
        if ((rc = gemm_tc(h, g, s))) return rc;
      }
      if ((rc = attention(h, h->ws_q, h->ws_k, h->ws_v, h->ws_o, VADB_BF16, lengths, Bc, T, s))) return rc;
      {  // h += o Wo^T + bo (transformer.py:347, :237); emits LN2(h) for the feed-forward sublayer
        GemmTcArgs g = {};
        g.M = M; g.N = D; g.K = D; g.w_bf16 = h->wo_bf + (size_t)l * D * D;
        g.a_bf16 = (const bf16*)h->ws_o; g.bias = w + lo.bo; g.residual = h->ws_h;
        g.out_f32 = 1; g.out[0] = h->ws_h;
        g.out[1] = h->ws_aln; g.emit_ln_g = w + lo.ln2_g; g.emit_ln_b = w + lo.ln2_b;
        if ((rc = gemm_tc(h, g, s))) return rc;
      }
      {  // h += relu(LN2(h) W1^T + b1) W2^T + b2 (transformer.py:234-238, :370-375), hidden kept on chip;
         // emits LN1 of the next layer
        FfnTcArgs f = {};
        f.M = M; f.h = h->ws_h; f.a_ln = h->ws_aln;
        f.w1_bf16 = h->w1_bf + (size_t)l * DFF * D; f.b1 = w + lo.b1;
        f.w2_bf16 = h->w2_bf + (size_t)l * D * DFF; f.b2 = w + lo.b2;
        if (l + 1 < h->cfg.num_layers) {
          f.emit_out = h->ws_aln;
          f.emit_ln_g = w + h->lay.layers[l + 1].ln1_g; f.emit_ln_b = w + h->lay.layers[l + 1].ln1_b;
        } else if (fuse_cls) {
          // last layer: final LayerNorm + classifier + log-softmax in the FFN epilogue (no classifier pass)
          f.cls_ln_g = w + h->lay.lnf_g; f.cls_ln_b = w + h->lay.lnf_b;
          f.cls_w = w + h->lay.wc; f.cls_bias = w + h->lay.bc;
          f.prob = prob; f.logp = logp;
        }
        std::string err;
        cudaError_t e = launch_ffn_tc(f, h->num_sms, s, &err);
        if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("ffn_tc: ") + cudaGetErrorString(e) + " " + err);
        h->launches++;
        h->aln_valid = true;
      }
      continue;
    }
    {  // a = LN1(h); q,k,v = a W^T + b        (transformer.py:235-236, :281-284)
      GemmArgs g = {};
      g.A = h->ws_h; g.W = h->wqkv + (size_t)l * 3 * D * D; g.bias = h->bqkv + (size_t)l * 3 * D;
      g.M = M; g.N = 3 * D; g.K = D;
      g.ln_g = w + lo.ln1_g; g.ln_b = w + lo.ln1_b;
      g.out[0] = h->ws_q; g.out[1] = h->ws_k; g.out[2] = h->ws_v; g.out_split = D;
      if ((rc = gemm(h, g, s))) return rc;
    }
    if ((rc = attention(h, h->ws_q, h->ws_k, h->ws_v, h->ws_o, VADB_F32, lengths, Bc, T, s)))
      return rc;
    {  // h += o Wo^T + bo                      (transformer.py:347, :237)
      GemmArgs g = {};
      g.A = h->ws_o; g.W = w + lo.wo; g.bias = w + lo.bo;
      g.M = M; g.N = D; g.K = D; g.residual = h->ws_h;
      g.out[0] = h->ws_h; g.out_split = D;
      if ((rc = gemm(h, g, s))) return rc;
    }
    {  // hid = relu(LN2(h) W1^T + b1)          (transformer.py:370-372)
      GemmArgs g = {};
      g.A = h->ws_h; g.W = w + lo.w1; g.bias = w + lo.b1;
      g.M = M; g.N = DFF; g.K = D; g.ln_g = w + lo.ln2_g; g.ln_b = w + lo.ln2_b; g.relu = 1;
      g.out[0] = h->ws_hid; g.out_split = DFF;
      if ((rc = gemm(h, g, s))) return rc;
    }
    {  // h += hid W2^T + b2                    (transformer.py:374, :237)
      GemmArgs g = {};
      g.A = h->ws_hid; g.W = w + lo.w2; g.bias = w + lo.b2;
      g.M = M; g.N = D; g.K = DFF; g.residual = h->ws_h;
      g.out[0] = h->ws_h; g.out_split = D;
      if ((rc = gemm(h, g, s))) return rc;
    }
  }
  if (bf && fuse_cls) return VADB_OK;
  cudaError_t e = launch_classifier(h->ws_h, w + h->lay.lnf_g, w + h->lay.lnf_b, w + h->lay.wc,
                                    w + h->lay.bc, M, prob, logp, 0, s);
  if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("classifier: ") + cudaGetErrorString(e));
  h->launches++;
  return VADB_OK;
}

// h = x W_in^T + b_in + PE/sqrt(d) (self_attention.py:13-14, transformer.py:401), optionally with the
// predictor's window gather folded into the A-row index.  bf16 mode with F <= 128 (F % 4 == 0): tensor
// cores; otherwise the fp32 CUDA-core kernel.
int front_end(vadb_handle* h, const void* x, int x_is_bf16, int M, int pe_T, int win_W, int win_half,
              int win_jump, cudaStream_t s) {
  const int F = h->cfg.feature_size;
  static const bool tf32_ok = !(getenv("VADB_FRONT_TF32") && atoi(getenv("VADB_FRONT_TF32")) == 0);
  if (is_bf16_mode(h) && F <= D && F % 4 == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0) {
    GemmTcArgs g = {};
    g.M = M; g.N = D; g.K = D; g.w_bf16 = h->win_bf;
    if (tf32_ok && !x_is_bf16 && win_W == 0) {
      // fp32 features straight from HBM to the tensor cores (TMA, kind::tf32): no conversion pass, and
      // 10 mantissa bits instead of bf16's 7 on the raw log-mel values
      g.a_f32_tma = (const float*)x; g.w_f32 = h->w32 + h->lay.w_in; g.a_cols = F;
      g.K = 128 * ((F + 63) / 64);
    } else if (tf32_ok && x_is_bf16 && win_W == 0 && F % 8 == 0) {
      // bf16 features: plain TMA-fed GEMM, the tensor map is F columns wide and TMA zero-fills up to K = 128
      g.a_bf16 = (const bf16*)x; g.a_cols = F;
    } else {
      g.a_rows = x; g.a_cols = F; g.a_rows_bf16 = x_is_bf16;
    }
    g.win_W = win_W; g.win_half = win_half; g.win_jump = win_jump;
    g.bias = h->w32 + h->lay.b_in; g.residual = h->pe; g.res_mod = pe_T;
    g.out_f32 = 1; g.out[0] = h->ws_h;
    // also emit LN1 of layer 0 in bf16: the A operand of the first Q/K/V GEMM
    g.out[1] = h->ws_aln;
    g.emit_ln_g = h->w32 + h->lay.layers[0].ln1_g; g.emit_ln_b = h->w32 + h->lay.layers[0].ln1_b;
    g.out_tiled = fuse_tail(h) ? 1 : 0;
    h->aln_valid = true;
    return gemm_tc(h, g, s);
  }
  h->aln_valid = false;
  GemmArgs g = {};
  g.A = x; g.a_is_bf16 = x_is_bf16;
  g.W = h->w32 + h->lay.w_in; g.bias = h->w32 + h->lay.b_in;
  g.M = M; g.N = D; g.K = F; g.pe = h->pe; g.pe_T = pe_T;
  g.win_W = win_W; g.win_half = win_half; g.win_jump = win_jump;
  g.out[0] = h->ws_h; g.out_split = D;
  if (fuse_tail(h)) {
    // CUDA-core front end (feature sizes the tensor-core front end does not take) feeding the fused bf16
    // layers: rows go to the (otherwise unused) hidden workspace, then into the tiled layout + LN1 of layer 0
    g.out[0] = h->ws_hid;
    int rc = gemm(h, g, s);
    if (rc) return rc;
    cudaError_t e = launch_retile_ln((const float*)h->ws_hid, h->ws_h, h->ws_aln, h->w32 + h->lay.layers[0].ln1_g,
                                     h->w32 + h->lay.layers[0].ln1_b, M, s);
    if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("retile: ") + cudaGetErrorString(e));
    h->launches++;
    h->aln_valid = true;
    return VADB_OK;
  }
  return gemm(h, g, s);
}

}  // namespace
namespace vadb {
bool pdl_enabled() {
  static const bool on = !(getenv("VADB_PDL") && atoi(getenv("VADB_PDL")) == 0);
  return on;
}
}  // namespace vadb
namespace {

int check_ready(vadb_handle* h) {
  if (!h) return VADB_E_INVALID;
  if (!h->loaded) return fail(h, VADB_E_STATE, "weights not loaded (call vadb_load_weights first)");
  return VADB_OK;
}

int window_W(int half, int jump) { return 2 * (half - 1) / jump + 3; }

int ensure_bytes(vadb_handle* h, void** p, size_t* have, size_t need, bool pinned) {
  if (need <= *have) return VADB_OK;
  if (*p) {
    cudaDeviceSynchronize();   // the old buffer may still be referenced by queued work
    if (pinned) cudaFreeHost(*p); else cudaFree(*p);
    *p = nullptr; *have = 0;
  }
  size_t n = need + need / 4;
  if (pinned) CU_TRY(h, cudaMallocHost(p, n)); else CU_TRY(h, cudaMalloc(p, n));
  *have = n;
  return VADB_OK;
}

}  // namespace

extern "C" {

const char* vadb_version(void) { return "vadb200 0.1 sm_100a"; }

size_t vadb_weight_count(const vadb_config* cfg) {
  if (!cfg || cfg->feature_size <= 0 || cfg->num_layers <= 0) return 0;
  return make_layout(cfg->feature_size, cfg->num_layers).total;
}

const char* vadb_last_error(const vadb_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int64_t vadb_launch_count(const vadb_handle* h) { return h ? h->launches : 0; }

int vadb_create(vadb_handle** out, const vadb_config* cfg, int device) {
  if (!out || !cfg) return fail(nullptr, VADB_E_INVALID, "null argument");
  *out = nullptr;
  if (cfg->d_model != D) return fail(nullptr, VADB_E_INVALID, "d_model must be 128 (kernels are specialised)");
  if (cfg->feature_size <= 0 || cfg->feature_size > 4096) return fail(nullptr, VADB_E_INVALID, "bad feature_size");
  if (cfg->num_layers <= 0 || cfg->num_layers > 64) return fail(nullptr, VADB_E_INVALID, "bad num_layers");
  if (cfg->compute_dtype != VADB_F32 && cfg->compute_dtype != VADB_BF16)
    return fail(nullptr, VADB_E_INVALID, "compute_dtype must be VADB_F32 or VADB_BF16");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, VADB_E_CUDA, std::string("no CUDA device (there is no CPU fallback): ") +
                                          cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, VADB_E_INVALID, "bad device index");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, VADB_E_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, VADB_E_INVALID, "libvadb200 is built for sm_100a (B200) only");
  vadb_handle* h = new (std::nothrow) vadb_handle();
  if (!h) return fail(nullptr, VADB_E_NOMEM, "out of host memory");
  h->cfg = *cfg;
  h->device = device;
  h->lay = make_layout(cfg->feature_size, cfg->num_layers);
  h->num_sms = prop.multiProcessorCount;
  DeviceGuard dg(device);
  e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream2, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream3, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->ev_call[i], cudaEventDisableTiming);
  if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); delete h; return fail(nullptr, VADB_E_CUDA, m); }
  *out = h;
  return VADB_OK;
}

void vadb_destroy(vadb_handle* h) {
  if (!h) return;
  DeviceGuard dg(h->device);
  cudaDeviceSynchronize();
  free_dev(h->w32); free_dev(h->wqkv); free_dev(h->bqkv);
  free_dev(h->wqkv_bf); free_dev(h->wo_bf); free_dev(h->w1_bf); free_dev(h->w2_bf); free_dev(h->win_bf); free_dev(h->wtail); free_dev(h->tail_aux); free_dev(h->head_bf16); free_dev(h->head_tf32); free_dev(h->head_aux);
  free_dev(h->pe_tiled);
  free_dev(h->pe);
  free_dev(h->ws_h); free_dev(h->ws_q); free_dev(h->ws_k); free_dev(h->ws_v); free_dev(h->ws_o);
  free_dev(h->ws_hid); free_dev(h->ws_prob); free_dev(h->ws_aln);
  if (h->pin_in) cudaFreeHost(h->pin_in);
  if (h->pin_out) cudaFreeHost(h->pin_out);
  if (h->dev_in) cudaFree(h->dev_in);
  if (h->dev_out) cudaFree(h->dev_out);
  if (h->dev_len) cudaFree(h->dev_len);
  if (h->win_prob) cudaFree(h->win_prob);
  if (h->win_proj) cudaFree(h->win_proj);
  if (h->bk_x) cudaFree(h->bk_x);
  if (h->bk_out) cudaFree(h->bk_out);
  if (h->bk_idx) cudaFree(h->bk_idx);
  if (h->lm_window) cudaFree(h->lm_window);
  if (h->lm_twiddle) cudaFree(h->lm_twiddle);
  if (h->lm_meta) cudaFree(h->lm_meta);
  if (h->lm_w) cudaFree(h->lm_w);
  if (h->lm_audio) cudaFree(h->lm_audio);
  if (h->lm_feat) cudaFree(h->lm_feat);
  for (int i = 0; i < 2; ++i) {
    if (h->bk_stream[i]) cudaStreamDestroy(h->bk_stream[i]);
    if (h->bk_join[i]) cudaEventDestroy(h->bk_join[i]);
  }
  if (h->bk_fork) cudaEventDestroy(h->bk_fork);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->own_stream2) cudaStreamDestroy(h->own_stream2);
  if (h->own_stream3) cudaStreamDestroy(h->own_stream3);
  for (int i = 0; i < 2; ++i) if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
    if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
  }
  for (int i = 0; i < 4; ++i)
    if (h->ev_call[i]) cudaEventDestroy(h->ev_call[i]);
  delete h;
}

static int alloc_weights(vadb_handle* h) {
  if (h->w32) return VADB_OK;
  const int L = h->cfg.num_layers;
  CU_TRY(h, cudaMalloc(&h->w32, h->lay.total * sizeof(float)));
  CU_TRY(h, cudaMalloc(&h->wqkv, (size_t)L * 3 * D * D * sizeof(float)));
  CU_TRY(h, cudaMalloc(&h->bqkv, (size_t)L * 3 * D * sizeof(float)));
  CU_TRY(h, cudaMalloc(&h->wqkv_bf, (size_t)L * 3 * D * D * sizeof(bf16)));
  CU_TRY(h, cudaMalloc(&h->wo_bf, (size_t)L * D * D * sizeof(bf16)));
  CU_TRY(h, cudaMalloc(&h->w1_bf, (size_t)L * DFF * D * sizeof(bf16)));
  CU_TRY(h, cudaMalloc(&h->w2_bf, (size_t)L * D * DFF * sizeof(bf16)));
  CU_TRY(h, cudaMalloc(&h->win_bf, (size_t)D * D * sizeof(bf16)));
  CU_TRY(h, cudaMalloc(&h->wtail, (size_t)L * tail_pack_bytes()));
  CU_TRY(h, cudaMalloc(&h->tail_aux, (size_t)L * tail_aux_floats() * sizeof(float)));
  CU_TRY(h, cudaMalloc(&h->head_bf16, 4 * (size_t)32768));
  CU_TRY(h, cudaMalloc(&h->head_tf32, 4 * (size_t)32768));
  CU_TRY(h, cudaMalloc(&h->head_aux, 3 * D * sizeof(float)));
  return VADB_OK;
}

// kernel-side copies of the packed blob in h->w32: fused Q|K|V rows, bf16 operands
static int derive_weights(vadb_handle* h, cudaStream_t s) {
  const int L = h->cfg.num_layers;
  if (h->cfg.feature_size <= D)
    CU_TRY(h, launch_pad_rows_bf16(h->w32 + h->lay.w_in, h->win_bf, D, h->cfg.feature_size, s));
  for (int l = 0; l < L; ++l) {
    const LayerOffsets& lo = h->lay.layers[l];
    const size_t srcw[3] = {lo.wq, lo.wk, lo.wv}, srcb[3] = {lo.bq, lo.bk, lo.bv};
    for (int j = 0; j < 3; ++j) {
      CU_TRY(h, cudaMemcpyAsync(h->wqkv + ((size_t)l * 3 + j) * D * D, h->w32 + srcw[j],
                                (size_t)D * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
      CU_TRY(h, cudaMemcpyAsync(h->bqkv + ((size_t)l * 3 + j) * D, h->w32 + srcb[j],
                                (size_t)D * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    CU_TRY(h, launch_f32_to_bf16(h->wqkv + (size_t)l * 3 * D * D, h->wqkv_bf + (size_t)l * 3 * D * D,
                                 (size_t)3 * D * D, s));
    CU_TRY(h, launch_f32_to_bf16(h->w32 + lo.wo, h->wo_bf + (size_t)l * D * D, (size_t)D * D, s));
    CU_TRY(h, launch_f32_to_bf16(h->w32 + lo.w1, h->w1_bf + (size_t)l * DFF * D, (size_t)DFF * D, s));
    CU_TRY(h, launch_f32_to_bf16(h->w32 + lo.w2, h->w2_bf + (size_t)l * D * DFF, (size_t)D * DFF, s));
  }
  for (int l = 0; l < L; ++l) {      // fused layer tail: Wo, W1, W2 of layer l + the fused Q|K|V weight of layer l+1
    const LayerOffsets& lo = h->lay.layers[l];
    TailPackArgs t = {};
    t.wo = h->w32 + lo.wo; t.w1 = h->w32 + lo.w1; t.b1 = h->w32 + lo.b1; t.w2 = h->w32 + lo.w2;
    t.ln2_g = h->w32 + lo.ln2_g; t.ln2_b = h->w32 + lo.ln2_b;
    if (l + 1 < L) {
      const LayerOffsets& ln = h->lay.layers[l + 1];
      t.wqkv_next = h->wqkv + (size_t)(l + 1) * 3 * D * D; t.bqkv_next = h->bqkv + (size_t)(l + 1) * 3 * D;
      t.ln1n_g = h->w32 + ln.ln1_g; t.ln1n_b = h->w32 + ln.ln1_b;
    } else {
      t.lnf_g = h->w32 + h->lay.lnf_g; t.lnf_b = h->w32 + h->lay.lnf_b; t.wc = h->w32 + h->lay.wc; t.bc = h->w32 + h->lay.bc;
    }
    CU_TRY(h, launch_tail_pack(t, h->wtail + (size_t)l * tail_pack_bytes(), h->tail_aux + (size_t)l * tail_aux_floats(), s));
  }
  CU_TRY(h, launch_head_pack(h->w32 + h->lay.w_in, h->cfg.feature_size, h->wqkv, h->bqkv, h->w32 + h->lay.layers[0].ln1_g,
                             h->w32 + h->lay.layers[0].ln1_b, h->head_bf16, h->head_tf32, h->head_aux, s));
  CU_TRY(h, cudaStreamSynchronize(s));
  h->loaded = true;
  return VADB_OK;
}

int vadb_load_weights(vadb_handle* h, const float* blob, size_t count, int on_device, void* stream) {
  if (!h || !blob) return VADB_E_INVALID;
  if (count != h->lay.total) {
    char buf[160];
    snprintf(buf, sizeof buf, "weight blob has %zu floats, config needs %zu", count, h->lay.total);
    return fail(h, VADB_E_INVALID, buf);
  }
  DeviceGuard dg(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = alloc_weights(h);
  if (rc) return rc;
  CU_TRY(h, cudaMemcpyAsync(h->w32, blob, count * sizeof(float),
                            on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  return derive_weights(h, s);
}

// NCCL is resolved at run time (the process that owns the communicator has it loaded already; otherwise
// libnccl.so.2 is opened), so libvadb200.so carries no link-time dependency on it.
int vadb_broadcast_weights(vadb_handle* h, void* nccl_comm, int root, void* stream) {
  if (!h || !nccl_comm) return VADB_E_INVALID;
  typedef int (*BcastFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  typedef int (*RankFn)(void*, int*);
  static BcastFn bcast = nullptr;
  static RankFn user_rank = nullptr;
  if (!bcast) {
    void* sym = dlsym(RTLD_DEFAULT, "ncclBroadcast");
    void* lib = nullptr;
    if (!sym && (lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL))) sym = dlsym(lib, "ncclBroadcast");
    if (!sym) return fail(h, VADB_E_STATE, "ncclBroadcast not found: load NCCL (libnccl.so.2) into the process first");
    void* sym2 = dlsym(RTLD_DEFAULT, "ncclCommUserRank");
    if (!sym2 && lib) sym2 = dlsym(lib, "ncclCommUserRank");
    if (!sym2) return fail(h, VADB_E_STATE, "ncclCommUserRank not found");
    bcast = reinterpret_cast<BcastFn>(sym);
    user_rank = reinterpret_cast<RankFn>(sym2);
  }
  int rank = -1;
  if (user_rank(nccl_comm, &rank) != 0) return fail(h, VADB_E_INVALID, "ncclCommUserRank failed (bad communicator?)");
  if (rank == root && !h->loaded) return fail(h, VADB_E_STATE, "the root rank must have loaded its weights (vadb_load_weights)");
  DeviceGuard dg(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = alloc_weights(h);
  if (rc) return rc;
  const int nccl_float32 = 7;      // ncclFloat32 (nccl.h, stable since NCCL 2.0)
  const int nr = bcast(h->w32, h->w32, h->lay.total, nccl_float32, root, nccl_comm, s);
  if (nr != 0) return fail(h, VADB_E_CUDA, "ncclBroadcast failed with ncclResult " + std::to_string(nr));
  if (rank == root) { CU_TRY(h, cudaStreamSynchronize(s)); return VADB_OK; }
  return derive_weights(h, s);
}

int vadb_reserve(vadb_handle* h, int B, int T) {
  if (!h || B <= 0 || T <= 0) return VADB_E_INVALID;
  DeviceGuard dg(h->device);
  int rc = ensure_pe(h, T);
  if (rc) return rc;
  return ensure_workspace(h, (size_t)clips_per_pass(B, T) * T);
}

int vadb_positional_table(vadb_handle* h, int T, float* out_host) {
  if (!h || T <= 0 || !out_host) return VADB_E_INVALID;
  DeviceGuard dg(h->device);
  int rc = ensure_pe(h, T);
  if (rc) return rc;
  CU_TRY(h, cudaMemcpy(out_host, h->pe, (size_t)T * D * sizeof(float), cudaMemcpyDeviceToHost));
  return VADB_OK;
}

int vadb_forward(vadb_handle* h, const void* x, int x_dtype, const int32_t* lengths, int B, int T,
                 float* prob, float* logp, void* stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!x || B < 0 || T < 0) return fail(h, VADB_E_INVALID, "bad forward arguments");
  if (x_dtype != VADB_F32 && x_dtype != VADB_BF16) return fail(h, VADB_E_INVALID, "x_dtype must be f32 or bf16");
  if (B == 0 || T == 0) return VADB_OK;
  DeviceGuard dg(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = ensure_pe(h, T))) return rc;
  const int cpp = clips_per_pass(B, T);
  if ((rc = ensure_workspace(h, (size_t)cpp * T))) return rc;
  const int F = h->cfg.feature_size;
  const size_t xsz = x_dtype == VADB_BF16 ? 2 : 4;
  // Head of the model in ONE kernel (front end + layer 0's q/k/v: the tail kernel's head mode) when a 128-frame tile
  // never wraps around the positional table and the features fit its operand tile; VADB_FUSE_HEAD=0: two kernels
  static const bool head_ok = !(getenv("VADB_FUSE_HEAD") && atoi(getenv("VADB_FUSE_HEAD")) == 0);
  const bool use_head = head_ok && fuse_tail(h) && T % 128 == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0 &&
                        (x_dtype == VADB_BF16 ? (F % 8 == 0 && F <= 128) : (F % 4 == 0 && F <= 64));
  if (use_head && (rc = ensure_pe_tiled(h, T))) return rc;
  for (int b0 = 0; b0 < B; b0 += cpp) {
    const int Bc = std::min(cpp, B - b0);
    if (use_head) {
      TailTcArgs t = {};
      t.M = Bc * T; t.x = (const char*)x + (size_t)b0 * T * F * xsz; t.x_is_bf16 = x_dtype == VADB_BF16; t.F = F;
      t.wpack = x_dtype == VADB_BF16 ? h->head_bf16 : h->head_tf32; t.aux = h->head_aux;
      t.bo = h->w32 + h->lay.b_in; t.h = h->ws_h; t.pe_tiled = h->pe_tiled; t.pe_tiles = T / 128;
      t.q = (bf16*)h->ws_q; t.k = (bf16*)h->ws_k; t.v = (bf16*)h->ws_v;
      std::string err;
      cudaError_t e = launch_tail_tc(t, h->num_sms, s, &err);
      if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("head_tc: ") + cudaGetErrorString(e) + " " + err);
      h->launches++;
      h->qkv_ready = true;
    } else if ((rc = front_end(h, (const char*)x + (size_t)b0 * T * F * xsz, x_dtype == VADB_BF16, Bc * T, T, 0, 0, 0, s)))
      return rc;
    if ((rc = run_encoder(h, lengths ? lengths + b0 : nullptr, Bc, T,
                          prob ? prob + (size_t)b0 * T : nullptr,
                          logp ? logp + (size_t)b0 * T * 2 : nullptr, s)))
      return rc;
  }
  return VADB_OK;
}

// Mixed-length batches without the padding work.  Clips are grouped by their length rounded up to whole 128-frame
// tiles; each group is gathered into a dense [n, T', F] batch, run through the ordinary forward with its own
// key-padding mask (lengths), and its results are scattered back to the caller's [B, T] layout.  Per-frame kernels
// and the attention kernel then only see ceil(len/128)*128 frames per clip instead of T.
int vadb_forward_ragged(vadb_handle* h, const void* x, int x_dtype, const int32_t* lengths_host, int B, int T,
                        float* prob, float* logp, void* stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!x || !lengths_host || B < 0 || T < 0) return fail(h, VADB_E_INVALID, "bad forward arguments");
  if (x_dtype != VADB_F32 && x_dtype != VADB_BF16) return fail(h, VADB_E_INVALID, "x_dtype must be f32 or bf16");
  if (B == 0 || T == 0) return VADB_OK;
  DeviceGuard dg(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int F = h->cfg.feature_size;
  const size_t xsz = x_dtype == VADB_BF16 ? 2 : 4;
  // bucket key: frames actually processed for the clip
  std::vector<int> tq(B);
  std::vector<int> keys;
  for (int b = 0; b < B; ++b) {
    const int len = std::min(std::max(lengths_host[b], 0), T);
    tq[b] = std::min(T, std::max(1, (len + 127) / 128) * 128);
    keys.push_back(tq[b]);
  }
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  // device copies of the lengths and of the per-bucket clip lists: [lengths (B) | bucket-ordered clip ids (B) |
  // bucket-ordered lengths (B)]; the source is pageable memory, so the asynchronous copy stages it before returning
  std::vector<int32_t> meta(3 * (size_t)B);
  size_t pos = 0;
  std::vector<size_t> start(keys.size() + 1, 0);
  size_t max_in = 0, max_out = 0;
  for (size_t k = 0; k < keys.size(); ++k) {
    start[k] = pos;
    for (int b = 0; b < B; ++b)
      if (tq[b] == keys[k]) {
        meta[(size_t)B + pos] = b;
        meta[2 * (size_t)B + pos] = std::min(std::max(lengths_host[b], 0), T);
        ++pos;
      }
    const size_t n = pos - start[k];
    max_in = std::max(max_in, n * keys[k] * F * xsz);
    max_out = std::max(max_out, n * keys[k] * 3 * sizeof(float) + 64);
  }
  start[keys.size()] = pos;
  for (int b = 0; b < B; ++b) meta[b] = std::min(std::max(lengths_host[b], 0), T);
  if ((rc = ensure_bytes(h, &h->bk_idx, &h->bk_idx_bytes, meta.size() * sizeof(int32_t), false))) return rc;
  CU_TRY(h, cudaMemcpyAsync(h->bk_idx, meta.data(), meta.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  const int32_t* d_len = (const int32_t*)h->bk_idx;
  const int32_t* d_ids = d_len + B;
  const int32_t* d_blen = d_len + 2 * (size_t)B;
  const bool vec_ok = ((size_t)T * F * xsz) % 16 == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0;
  if ((keys.size() == 1 && keys[0] == T) || !vec_ok)    // nothing to gain (or rows not 16-byte copyable): padded forward
    return vadb_forward(h, x, x_dtype, d_len, B, T, prob, logp, stream);
  // Buckets are independent forwards: each gets its own slice of the gather buffers and of the workspace and they
  // are spread over the caller's stream and two side streams, so the short buckets' small grids run in the SMs the
  // long bucket's last wave leaves idle instead of each costing a chain of launch latencies.
  const size_t K = keys.size();
  std::vector<size_t> in_off(K + 1, 0), out_off(K + 1, 0), ws_off(K + 1, 0);
  for (size_t k = 0; k < K; ++k) {
    const size_t n = start[k + 1] - start[k];
    in_off[k + 1] = in_off[k] + ((n * keys[k] * F * xsz + 255) & ~(size_t)255);
    out_off[k + 1] = out_off[k] + ((n * keys[k] * 3 * sizeof(float) + 64 + 255) & ~(size_t)255);
    ws_off[k + 1] = ws_off[k] + n * keys[k];              // frames; multiples of 128 (whole tiles) unless the key is T itself
    if (ws_off[k + 1] % 128) ws_off[k + 1] += 128 - ws_off[k + 1] % 128;
  }
  const bool concurrent = K > 1 && ws_off[K] <= MAX_FRAMES_PER_PASS &&
                          !(getenv("VADB_BUCKET_STREAMS") && atoi(getenv("VADB_BUCKET_STREAMS")) == 0);
  if ((rc = ensure_bytes(h, &h->bk_x, &h->bk_x_bytes, concurrent ? in_off[K] : max_in, false))) return rc;
  if ((rc = ensure_bytes(h, &h->bk_out, &h->bk_out_bytes, concurrent ? out_off[K] : max_out, false))) return rc;
  if (concurrent) {
    // everything that may (re)allocate happens before the fork
    if ((rc = ensure_pe(h, keys.back()))) return rc;
    if ((rc = ensure_workspace(h, ws_off[K]))) return rc;
    if (fuse_tail(h) && keys.back() % 128 == 0 && (rc = ensure_pe_tiled(h, keys.back()))) return rc;
    for (int i = 0; i < 2; ++i) {
      if (!h->bk_stream[i]) CU_TRY(h, cudaStreamCreateWithFlags(&h->bk_stream[i], cudaStreamNonBlocking));
      if (!h->bk_join[i]) CU_TRY(h, cudaEventCreateWithFlags(&h->bk_join[i], cudaEventDisableTiming));
    }
    if (!h->bk_fork) CU_TRY(h, cudaEventCreateWithFlags(&h->bk_fork, cudaEventDisableTiming));
  }
  // frames past a clip's processed length are not computed: defined as 0
  if (prob) CU_TRY(h, cudaMemsetAsync(prob, 0, (size_t)B * T * sizeof(float), s));
  if (logp) CU_TRY(h, cudaMemsetAsync(logp, 0, (size_t)B * T * 2 * sizeof(float), s));
  if (concurrent) {
    CU_TRY(h, cudaEventRecord(h->bk_fork, s));
    for (int i = 0; i < 2; ++i) CU_TRY(h, cudaStreamWaitEvent(h->bk_stream[i], h->bk_fork, 0));
  }
  // longest bucket first on the caller's stream, the others alternate over the side streams
  struct WsView {            // the handle's workspace pointers shifted to one bucket's slice while its work is enqueued
    vadb_handle* h; bool on; size_t cap; float* hh; void *q, *k, *v, *o, *hid; bf16* aln; float* pr;
    WsView(vadb_handle* hd, bool active, size_t off, size_t frames, size_t act) : h(hd), on(active), cap(hd->cap_frames), hh(hd->ws_h),
        q(hd->ws_q), k(hd->ws_k), v(hd->ws_v), o(hd->ws_o), hid(hd->ws_hid), aln(hd->ws_aln), pr(hd->ws_prob) {
      if (!on) return;       // sequential buckets use (and may grow) the whole workspace
      h->ws_h += off * D; h->ws_q = (char*)q + off * D * act; h->ws_k = (char*)k + off * D * act; h->ws_v = (char*)v + off * D * act;
      h->ws_o = (char*)o + off * D * act; h->ws_hid = (char*)hid + off * DFF * act; h->ws_aln += off * D; h->ws_prob += off;
      h->cap_frames = frames;
    }
    ~WsView() { if (!on) return; h->cap_frames = cap; h->ws_h = hh; h->ws_q = q; h->ws_k = k; h->ws_v = v; h->ws_o = o; h->ws_hid = hid; h->ws_aln = aln; h->ws_prob = pr; }
  };
  const size_t act = is_bf16_mode(h) ? sizeof(bf16) : sizeof(float);
  // the side streams rejoin the caller's stream on EVERY way out once they have been forked (a failed bucket must not
  // leave work behind that races with the next call on the shared workspace)
  auto join = [&]() {
    if (!concurrent) return;
    for (int i = 0; i < 2; ++i)
      if (cudaEventRecord(h->bk_join[i], h->bk_stream[i]) == cudaSuccess) cudaStreamWaitEvent(s, h->bk_join[i], 0);
  };
  int side = 0;
  for (size_t kk = 0; kk < K; ++kk) {
    const size_t k = K - 1 - kk;
    const int n = (int)(start[k + 1] - start[k]), Tk = keys[k];
    cudaStream_t sk = (!concurrent || kk == 0) ? s : h->bk_stream[side++ & 1];
    const int32_t* ids = d_ids + start[k];
    char* xin = (char*)h->bk_x + (concurrent ? in_off[k] : 0);
    cudaError_t e = launch_gather_clips(x, xin, ids, n, T, Tk, (int)(F * xsz), sk);
    if (e != cudaSuccess) { join(); return fail(h, VADB_E_CUDA, std::string("gather: ") + cudaGetErrorString(e)); }
    h->launches++;
    float* pk = (float*)((char*)h->bk_out + (concurrent ? out_off[k] : 0));
    float* lk = pk + (((size_t)n * Tk + 3) & ~(size_t)3);
    {
      WsView view(h, concurrent, ws_off[k], ws_off[k + 1] - ws_off[k], act);
      rc = vadb_forward(h, xin, x_dtype, d_blen + start[k], n, Tk, prob ? pk : nullptr, logp ? lk : nullptr, (void*)sk);
    }
    if (rc) { join(); return rc; }
    e = launch_scatter_clips(prob ? pk : nullptr, logp ? lk : nullptr, prob, logp, ids, n, T, Tk, sk);
    if (e != cudaSuccess) { join(); return fail(h, VADB_E_CUDA, std::string("scatter: ") + cudaGetErrorString(e)); }
    h->launches++;
  }
  join();
  return VADB_OK;
}

// Shared body of vadb_forward_host (ticket == nullptr: returns when the outputs are in host memory) and
// vadb_forward_host_async (ticket != nullptr: everything is only enqueued; pinned buffers required).
static int forward_host_impl(vadb_handle* h, const void* x, int x_dtype, const int32_t* lengths, int B, int T,
                             float* prob, float* logp, long* ticket) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!x || B < 0 || T < 0) return fail(h, VADB_E_INVALID, "bad forward arguments");
  if (x_dtype != VADB_F32 && x_dtype != VADB_BF16) return fail(h, VADB_E_INVALID, "x_dtype must be f32 or bf16");
  if (ticket) *ticket = -1;
  if (B == 0 || T == 0) return VADB_OK;
  // the completion events and the caller's pinned output buffers form a ring of four: a fifth call may
  // only be enqueued once the oldest outstanding ticket has been waited for
  if (ticket && h->call_seq - 4 > h->waited_upto)
    return fail(h, VADB_E_STATE, "four asynchronous host calls are outstanding: vadb_host_wait on a ticket first");
  DeviceGuard dg(h->device);
  // Chunked two-stream pipeline: the H2D copy of clip chunk i+1 overlaps the forward of chunk i
  // (the reference does one blocking .to(device) per 1000-window chunk, vad/predictor.py:223).
  // Three streams: uploads, forwards, downloads.  The download of call i runs beside the forward of call i+1 (two
  // device output halves alternate per call), so the compute stream carries kernels only.
  cudaStream_t s_copy = h->own_stream, s_comp = h->own_stream2, s_down = h->own_stream3;
  const int F = h->cfg.feature_size;
  const size_t clip_in = (size_t)T * F * (x_dtype == VADB_BF16 ? sizeof(bf16) : sizeof(float));
  // >= 8 MB per chunk, at most 4 chunks: enough overlap, few (small, launch-bound) forward passes
  // (measured on B200, 33.5 MB of features: 1 chunk 1.25 ms, 2: 1.06, 3: 1.11, 4: 1.04, 6: 1.38)
  int C = (int)std::max<size_t>(1, ((size_t)8 << 20) / std::max<size_t>(clip_in, 1));
  C = std::max(C, (B + 3) / 4);
  C = std::min(C, B);
  if (ticket) C = B;                                  // asynchronous calls overlap ACROSS calls (two device input slots): one chunk,
                                                      // one full-size forward (two half-size forwards cost ~15 % more compute)
  if (const char* e = getenv("VADB_HOST_CHUNKS")) {    // tuning knob: force the number of chunks
    const int want = atoi(e);
    if (want >= 1) C = std::max(1, (B + want - 1) / want);
  }
  // (sizes ramping up 5:8:9:10 instead of four equal chunks: modelled -5 %, measured +2 % -- not used)
  const int n_chunks = (B + C - 1) / C;
  const size_t n = (size_t)B * T;
  auto is_pinned = [](const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
  const bool in_pinned = is_pinned(x);
  const bool out_pinned = (!prob || is_pinned(prob)) && (!logp || is_pinned(logp));
  if (ticket && !(in_pinned && out_pinned))
    return fail(h, VADB_E_INVALID, "asynchronous host calls need pinned (page-locked) input and output buffers");
  const size_t chunk_in = (size_t)C * clip_in;
  if (chunk_in != h->last_chunk_in) {
    // the two device input slots are laid out by chunk size: a call with another chunk size must not
    // upload over slots an outstanding asynchronous call is still reading
    CU_TRY(h, cudaStreamSynchronize(s_comp));
    h->last_chunk_in = chunk_in;
  }
  if (!in_pinned && (rc = ensure_bytes(h, &h->pin_in, &h->pin_in_bytes, 2 * chunk_in, true))) return rc;
  if (!out_pinned && (rc = ensure_bytes(h, &h->pin_out, &h->pin_out_bytes, n * 3 * sizeof(float), true))) return rc;
  if ((rc = ensure_bytes(h, &h->dev_in, &h->dev_in_bytes, 2 * chunk_in, false))) return rc;
  const size_t out_half = (n * 3 + 4 + 3) & ~(size_t)3;       // floats per output half
  if ((rc = ensure_bytes(h, &h->dev_out, &h->dev_out_bytes, 2 * out_half * sizeof(float), false))) return rc;
  if ((rc = ensure_pe(h, T))) return rc;
  if ((rc = ensure_workspace(h, (size_t)std::min(C, clips_per_pass(C, T)) * T))) return rc;
  int32_t* dlen = nullptr;
  if (lengths) {
    if ((size_t)B > h->dev_len_n) {
      if (h->dev_len) { cudaDeviceSynchronize(); cudaFree(h->dev_len); }
      h->dev_len = nullptr; h->dev_len_n = 0;
      CU_TRY(h, cudaMalloc(&h->dev_len, (size_t)B * sizeof(int32_t)));
      h->dev_len_n = B;
    }
    CU_TRY(h, cudaMemcpyAsync(h->dev_len, lengths, (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, s_comp));
    dlen = h->dev_len;
  }
  const int oslot = (int)(h->out_seq++ & 1);
  float* dprob = (float*)h->dev_out + (size_t)oslot * out_half;
  float* dlogp = dprob + ((n + 3) & ~(size_t)3);       // 16-byte aligned for any B*T (the fused classifier stores float2)
  // this output half was last used two calls ago: its downloads must have finished before a forward writes it again
  // (a call of another size lays the halves out differently: then wait for the previous call's downloads too)
  CU_TRY(h, cudaStreamWaitEvent(s_comp, h->ev_out[oslot], 0));
  if (out_half != h->last_out_half) {
    CU_TRY(h, cudaStreamWaitEvent(s_comp, h->ev_out[oslot ^ 1], 0));
    h->last_out_half = out_half;
  }
  float* hprob = out_pinned ? prob : (float*)h->pin_out;
  float* hlogp = out_pinned ? logp : (float*)h->pin_out + n;
  for (int i = 0; i < n_chunks; ++i) {
    // the two device input slots alternate per chunk ACROSS calls, so an asynchronous call can upload
    // while the previous call still computes
    const int b0 = i * C, Bc = std::min(C, B - b0), slot = (int)((h->chunk_seq + i) & 1);
    const size_t bytes = (size_t)Bc * clip_in;
    const char* src = (const char*)x + (size_t)b0 * clip_in;
    char* dsti = (char*)h->dev_in + (size_t)slot * chunk_in;
    // device slot free once the forward of the chunk that used it last (two chunks ago, possibly in the
    // previous call) has consumed it; a never-recorded event does not block
    CU_TRY(h, cudaStreamWaitEvent(s_copy, h->ev_done[slot], 0));
    if (!in_pinned) {
      char* stage = (char*)h->pin_in + (size_t)slot * chunk_in;
      if (i >= 2) CU_TRY(h, cudaEventSynchronize(h->ev_h2d[slot]));   // staging slot copied out
      memcpy(stage, src, bytes);
      src = stage;
    }
    CU_TRY(h, cudaMemcpyAsync(dsti, src, bytes, cudaMemcpyHostToDevice, s_copy));
    CU_TRY(h, cudaEventRecord(h->ev_h2d[slot], s_copy));
    CU_TRY(h, cudaStreamWaitEvent(s_comp, h->ev_h2d[slot], 0));
    if ((rc = vadb_forward(h, dsti, x_dtype, dlen ? dlen + b0 : nullptr, Bc, T,
                           prob ? dprob + (size_t)b0 * T : nullptr,
                           logp ? dlogp + (size_t)b0 * T * 2 : nullptr, s_comp)))
      return rc;
    CU_TRY(h, cudaEventRecord(h->ev_done[slot], s_comp));
    CU_TRY(h, cudaStreamWaitEvent(s_down, h->ev_done[slot], 0));
    if (prob) CU_TRY(h, cudaMemcpyAsync(hprob + (size_t)b0 * T, dprob + (size_t)b0 * T, (size_t)Bc * T * sizeof(float),
                                        cudaMemcpyDeviceToHost, s_down));
    if (logp) CU_TRY(h, cudaMemcpyAsync(hlogp + (size_t)b0 * T * 2, dlogp + (size_t)b0 * T * 2,
                                        (size_t)Bc * T * 2 * sizeof(float), cudaMemcpyDeviceToHost, s_down));
  }
  h->chunk_seq += n_chunks;
  CU_TRY(h, cudaEventRecord(h->ev_out[oslot], s_down));
  if (ticket) {                      // results land in the caller's pinned buffers; vadb_host_wait(ticket)
    CU_TRY(h, cudaEventRecord(h->ev_call[h->call_seq & 3], s_down));     // downloads complete in call order
    *ticket = h->call_seq++;
    return VADB_OK;
  }
  CU_TRY(h, cudaStreamSynchronize(s_down));
  CU_TRY(h, cudaStreamSynchronize(s_comp));
  CU_TRY(h, cudaStreamSynchronize(s_copy));
  h->waited_upto = h->call_seq - 1;       // earlier asynchronous calls ran on the same streams: all complete
  if (!out_pinned) {
    if (prob) memcpy(prob, hprob, n * sizeof(float));
    if (logp) memcpy(logp, hlogp, 2 * n * sizeof(float));
  }
  return VADB_OK;
}

int vadb_forward_host(vadb_handle* h, const void* x, int x_dtype, const int32_t* lengths, int B, int T,
                      float* prob, float* logp) {
  return forward_host_impl(h, x, x_dtype, lengths, B, T, prob, logp, nullptr);
}

int vadb_forward_host_async(vadb_handle* h, const void* x, int x_dtype, const int32_t* lengths, int B, int T,
                            float* prob, float* logp, long* ticket) {
  if (!ticket) return h ? fail(h, VADB_E_INVALID, "ticket must not be NULL") : VADB_E_INVALID;
  return forward_host_impl(h, x, x_dtype, lengths, B, T, prob, logp, ticket);
}

int vadb_host_wait(vadb_handle* h, long ticket) {
  if (!h) return VADB_E_INVALID;
  if (ticket < 0) return VADB_OK;                                   // empty call: nothing was enqueued
  if (ticket >= h->call_seq) return fail(h, VADB_E_INVALID, "unknown ticket");
  if (ticket <= h->waited_upto) return VADB_OK;                     // already known to be complete
  DeviceGuard dg(h->device);
  // Calls complete in ticket order (one compute stream).  A ticket whose own event slot has been reused
  // by a later call (it cannot be, given the bound in vadb_forward_host_async, but stay safe) waits for
  // the NEWEST call instead -- never "success" without a synchronisation.
  const bool slot_reused = ticket + 4 < h->call_seq;       // call ticket+4 (recorded when call_seq was ticket+4) took the slot
  const long on = slot_reused ? h->call_seq - 1 : ticket;
  CU_TRY(h, cudaEventSynchronize(h->ev_call[on & 3]));
  if (on > h->waited_upto) h->waited_upto = on;
  return VADB_OK;
}

int vadb_predict_probabilities(vadb_handle* h, const float* feat, int L, int half, int jump,
                               float* probs_LW, float* mean_L, void* stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!feat || L < 0 || half < 1 || jump < 1) return fail(h, VADB_E_INVALID, "bad window arguments");
  const int W = window_W(half, jump);
  const int nl = (half + jump - 1) / jump;
  if (2 * nl + 1 != W)   // the reference's scatter (vad/predictor.py:254) has mismatching shapes here
    return fail(h, VADB_E_INVALID, "context_window half/jump inconsistent with 2*(half-1)//jump+3 window slots");
  if (L == 0) return VADB_OK;
  DeviceGuard dg(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int n = L - 2 * half;                         // vad/predictor.py:169
  const int F = h->cfg.feature_size;
  if (n > 0) {
    if ((rc = ensure_pe(h, W))) return rc;
    const int wpp = clips_per_pass(n, W);             // windows per pass
    // per-window probabilities for all n windows live in one [n, W] buffer
    if ((rc = ensure_workspace(h, (size_t)wpp * W))) return rc;
    float* prob_all = nullptr;
    size_t need = (size_t)n * W * sizeof(float);
    if ((rc = ensure_bytes(h, &h->win_prob, &h->win_prob_bytes, need, false))) return rc;
    prob_all = (float*)h->win_prob;
    // bf16 mode: project-then-gather.  The input Linear is per frame, so the L frames are projected
    // once (tensor cores) and each window row is proj[src] + PE[slot]: W times fewer MMAs and no
    // register-staged gather inside the GEMM.  fp32 mode keeps the gather folded into the GEMM.
    static const bool ptg_ok = !(getenv("VADB_WINDOW_PTG") && atoi(getenv("VADB_WINDOW_PTG")) == 0);
    const bool ptg = ptg_ok && is_bf16_mode(h) && F <= D && F % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) % 16) == 0;
    if (ptg) {
      if ((rc = ensure_bytes(h, &h->win_proj, &h->win_proj_bytes, (size_t)L * D * sizeof(float), false))) return rc;
      GemmTcArgs g = {};
      g.M = L; g.N = D; g.K = D; g.w_bf16 = h->win_bf;
      static const bool tf32_ok = !(getenv("VADB_FRONT_TF32") && atoi(getenv("VADB_FRONT_TF32")) == 0);
      if (tf32_ok) { g.a_f32_tma = feat; g.w_f32 = h->w32 + h->lay.w_in; g.a_cols = F; g.K = 128 * ((F + 63) / 64); }
      else { g.a_rows = feat; g.a_cols = F; g.a_rows_bf16 = 0; }
      g.bias = h->w32 + h->lay.b_in;
      g.out_f32 = 1; g.out[0] = h->win_proj;
      if ((rc = gemm_tc(h, g, s))) return rc;
    }
    for (int c0 = 0; c0 < n; c0 += wpp) {
      const int nc = std::min(wpp, n - c0);
      if (ptg) {
        cudaError_t e = launch_window_gather_ln((const float*)h->win_proj + (size_t)c0 * D, h->pe, h->ws_h, h->ws_aln,
                                                h->w32 + h->lay.layers[0].ln1_g, h->w32 + h->lay.layers[0].ln1_b,
                                                (long)nc * W, W, half, jump, fuse_tail(h) ? 1 : 0, s);
        if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("window gather: ") + cudaGetErrorString(e));
        h->aln_valid = true;
        h->launches++;
      } else {
        // window gather folded into the front-end GEMM's A-row index (vad/predictor.py:182-218);
        // the positional slot of row m is m % W (the window is the model's whole sequence)
        if ((rc = front_end(h, feat + (size_t)c0 * F, 0, nc * W, W, W, half, jump, s))) return rc;
      }
      if ((rc = run_encoder(h, nullptr, nc, W, prob_all + (size_t)c0 * W, nullptr, s))) return rc;
    }
    cudaError_t e = launch_boost(prob_all, L, half, jump, W, probs_LW, mean_L, s);
    if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("boost: ") + cudaGetErrorString(e));
  } else {
    cudaError_t e = launch_boost(nullptr, L, half, jump, W, probs_LW, mean_L, s);
    if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("boost: ") + cudaGetErrorString(e));
  }
  h->launches++;
  return VADB_OK;
}

// The blocking host entry points below share the staging buffers and the workspace with vadb_forward_host_async: any
// asynchronous call that is still in flight is completed first.
static int drain_async_host_calls(vadb_handle* h) {
  if (h->call_seq - 1 <= h->waited_upto) return VADB_OK;
  DeviceGuard dg(h->device);
  CU_TRY(h, cudaStreamSynchronize(h->own_stream2));
  CU_TRY(h, cudaStreamSynchronize(h->own_stream3));
  h->waited_upto = h->call_seq - 1;
  return VADB_OK;
}

int vadb_predict_probabilities_host(vadb_handle* h, const float* feat, int L, int half, int jump,
                                    float* probs_LW, float* mean_L) {
  int rc = check_ready(h);
  if (rc) return rc;
  if ((rc = drain_async_host_calls(h))) return rc;
  if (!feat || L < 0 || half < 1 || jump < 1) return fail(h, VADB_E_INVALID, "bad window arguments");
  if (L == 0) return VADB_OK;
  DeviceGuard dg(h->device);
  cudaStream_t s = h->own_stream;
  const int W = window_W(half, jump);
  const int F = h->cfg.feature_size;
  const size_t in_bytes = (size_t)L * F * sizeof(float);
  const size_t out_floats = (size_t)L * (W + 1);
  if ((rc = ensure_bytes(h, &h->pin_in, &h->pin_in_bytes, in_bytes, true))) return rc;
  if ((rc = ensure_bytes(h, &h->pin_out, &h->pin_out_bytes, out_floats * sizeof(float), true))) return rc;
  if ((rc = ensure_bytes(h, &h->dev_in, &h->dev_in_bytes, in_bytes + out_floats * sizeof(float), false))) return rc;
  memcpy(h->pin_in, feat, in_bytes);
  CU_TRY(h, cudaMemcpyAsync(h->dev_in, h->pin_in, in_bytes, cudaMemcpyHostToDevice, s));
  float* d_lw = (float*)((char*)h->dev_in + in_bytes);
  float* d_mean = d_lw + (size_t)L * W;
  if ((rc = vadb_predict_probabilities(h, (const float*)h->dev_in, L, half, jump, d_lw, d_mean, s))) return rc;
  CU_TRY(h, cudaMemcpyAsync(h->pin_out, d_lw, out_floats * sizeof(float), cudaMemcpyDeviceToHost, s));
  CU_TRY(h, cudaStreamSynchronize(s));
  if (probs_LW) memcpy(probs_LW, h->pin_out, (size_t)L * W * sizeof(float));
  if (mean_L) memcpy(mean_L, (float*)h->pin_out + (size_t)L * W, (size_t)L * sizeof(float));
  return VADB_OK;
}

/* ---- log-mel front end ---- */
long vadb_logmel_frames(long n_samples, int hop) { return (n_samples < 0 || hop <= 0) ? 0 : 1 + n_samples / hop; }

int vadb_logmel_tables(int sample_rate, int n_fft, int win, int n_mels, float* fb_dense, double* window) {
  if (sample_rate <= 0 || n_fft < 2 || (n_fft & (n_fft - 1)) || win <= 0 || win > n_fft || n_mels <= 0) return VADB_E_INVALID;
  std::vector<float> fb;
  std::vector<double> w;
  logmel_tables(sample_rate, n_fft, win, n_mels, &fb, &w);
  if (fb_dense) memcpy(fb_dense, fb.data(), fb.size() * sizeof(float));
  if (window) memcpy(window, w.data(), w.size() * sizeof(double));
  return VADB_OK;
}

static int ensure_logmel(vadb_handle* h, int sr, int n_fft, int win, int n_mels, cudaStream_t s) {
  if (sr <= 0 || n_fft < 32 || n_fft > 4096 || (n_fft & (n_fft - 1)) || win <= 0 || win > n_fft || n_mels <= 0 || n_mels > 1024)
    return fail(h, VADB_E_INVALID, "log-mel: n_fft must be a power of two in [32, 4096], 0 < win <= n_fft");
  if (h->lm_sr == sr && h->lm_nfft == n_fft && h->lm_win == win && h->lm_nmels == n_mels) return VADB_OK;
  std::vector<float> fb;
  std::vector<double> window;
  logmel_tables(sr, n_fft, win, n_mels, &fb, &window);
  const int n_bins = n_fft / 2 + 1;
  std::vector<int> meta(3 * (size_t)n_mels);
  std::vector<float> packed;
  for (int m = 0; m < n_mels; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < n_bins; ++k)
      if (fb[(size_t)m * n_bins + k] != 0.f) { if (first < 0) first = k; last = k; }
    meta[m] = first < 0 ? 0 : first;
    meta[n_mels + m] = first < 0 ? 0 : last - first + 1;
    meta[2 * n_mels + m] = (int)packed.size();
    for (int k = meta[m]; k < meta[m] + meta[n_mels + m]; ++k) packed.push_back(fb[(size_t)m * n_bins + k]);
  }
  if (packed.empty()) packed.push_back(0.f);
  std::vector<double> tw((size_t)n_fft);      // [n_fft/2] (cos, -sin)
  const double two_pi = 6.283185307179586476925286766559;
  for (int k = 0; k < n_fft / 2; ++k) { tw[2 * k] = cos(two_pi * k / n_fft); tw[2 * k + 1] = -sin(two_pi * k / n_fft); }
  cudaStreamSynchronize(s);                    // queued kernels may still read the old tables
  if (h->lm_window) cudaFree(h->lm_window);
  if (h->lm_twiddle) cudaFree(h->lm_twiddle);
  if (h->lm_meta) cudaFree(h->lm_meta);
  if (h->lm_w) cudaFree(h->lm_w);
  h->lm_window = nullptr; h->lm_twiddle = nullptr; h->lm_meta = nullptr; h->lm_w = nullptr; h->lm_sr = 0;
  CU_TRY(h, cudaMalloc(&h->lm_window, window.size() * sizeof(double)));
  CU_TRY(h, cudaMalloc(&h->lm_twiddle, tw.size() * sizeof(double)));
  CU_TRY(h, cudaMalloc(&h->lm_meta, meta.size() * sizeof(int)));
  CU_TRY(h, cudaMalloc(&h->lm_w, packed.size() * sizeof(float)));
  CU_TRY(h, cudaMemcpy(h->lm_window, window.data(), window.size() * sizeof(double), cudaMemcpyHostToDevice));
  CU_TRY(h, cudaMemcpy(h->lm_twiddle, tw.data(), tw.size() * sizeof(double), cudaMemcpyHostToDevice));
  CU_TRY(h, cudaMemcpy(h->lm_meta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice));
  CU_TRY(h, cudaMemcpy(h->lm_w, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
  h->lm_sr = sr; h->lm_nfft = n_fft; h->lm_win = win; h->lm_nmels = n_mels;
  return VADB_OK;
}

int vadb_logmel(vadb_handle* h, const float* audio, long n_samples, int sample_rate, int n_fft, int hop,
                int win, int n_mels, float* feat, void* stream) {
  if (!h) return VADB_E_INVALID;
  if (!audio || !feat || n_samples <= 0 || hop <= 0) return fail(h, VADB_E_INVALID, "bad log-mel arguments");
  DeviceGuard dg(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = ensure_logmel(h, sample_rate, n_fft, win, n_mels, s);
  if (rc) return rc;
  const long n_frames = vadb_logmel_frames(n_samples, hop);
  cudaError_t e = launch_logmel(audio, n_samples, n_fft, hop, n_frames, h->lm_window, h->lm_twiddle, h->lm_meta,
                                h->lm_meta + n_mels, h->lm_meta + 2 * n_mels, h->lm_w, n_mels, feat, s);
  if (e != cudaSuccess) return fail(h, VADB_E_CUDA, std::string("logmel: ") + cudaGetErrorString(e));
  h->launches++;
  return VADB_OK;
}

int vadb_predict_audio_host(vadb_handle* h, const float* audio, long n_samples, int sample_rate, int n_fft,
                            int hop, int win, int half, int jump, float* feat_out, float* probs_LW,
                            float* mean_L) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!audio || n_samples <= 0 || hop <= 0 || half < 1 || jump < 1) return fail(h, VADB_E_INVALID, "bad audio arguments");
  if ((rc = drain_async_host_calls(h))) return rc;
  DeviceGuard dg(h->device);
  cudaStream_t s = h->own_stream;
  const int F = h->cfg.feature_size, W = window_W(half, jump);
  const long L = vadb_logmel_frames(n_samples, hop);
  if (L > 0x7fffffffL / (W + 1)) return fail(h, VADB_E_INVALID, "audio too long for one call (split it: split_max_seconds)");
  const size_t a_bytes = (size_t)n_samples * sizeof(float);
  if ((rc = ensure_bytes(h, &h->lm_audio, &h->lm_audio_bytes, a_bytes, false))) return rc;
  if ((rc = ensure_bytes(h, &h->lm_feat, &h->lm_feat_bytes, (size_t)L * F * sizeof(float), false))) return rc;
  cudaPointerAttributes pa;
  const bool audio_pinned = cudaPointerGetAttributes(&pa, audio) == cudaSuccess && pa.type == cudaMemoryTypeHost;
  if (!audio_pinned) cudaGetLastError();
  if (audio_pinned) {
    // caller's buffer is page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory): one DMA, no staging copy
    CU_TRY(h, cudaMemcpyAsync(h->lm_audio, audio, a_bytes, cudaMemcpyHostToDevice, s));
  } else {
    // upload in chunks through the two pinned staging slots: the memcpy of chunk i+1 overlaps the DMA of chunk i
    const size_t CH = (size_t)4 << 20;
    if ((rc = ensure_bytes(h, &h->pin_in, &h->pin_in_bytes, 2 * CH, true))) return rc;
    int n_ch = 0;
    for (size_t off = 0; off < a_bytes; off += CH, ++n_ch) {
      const size_t nb = std::min(CH, a_bytes - off);
      const int slot = n_ch & 1;
      char* stage = (char*)h->pin_in + (size_t)slot * CH;
      if (n_ch >= 2) CU_TRY(h, cudaEventSynchronize(h->ev_h2d[slot]));
      memcpy(stage, (const char*)audio + off, nb);
      CU_TRY(h, cudaMemcpyAsync((char*)h->lm_audio + off, stage, nb, cudaMemcpyHostToDevice, s));
      CU_TRY(h, cudaEventRecord(h->ev_h2d[slot], s));
    }
  }
  if ((rc = vadb_logmel(h, (const float*)h->lm_audio, n_samples, sample_rate, n_fft, hop, win, F,
                        (float*)h->lm_feat, s)))
    return rc;
  const size_t out_floats = (size_t)L * (W + 1);
  const size_t feat_floats = feat_out ? (size_t)L * F : 0;
  if ((rc = ensure_bytes(h, &h->pin_out, &h->pin_out_bytes, (out_floats + feat_floats) * sizeof(float), true))) return rc;
  if ((rc = ensure_bytes(h, &h->dev_out, &h->dev_out_bytes, out_floats * sizeof(float), false))) return rc;
  float* d_lw = (float*)h->dev_out;
  float* d_mean = d_lw + (size_t)L * W;
  if ((rc = vadb_predict_probabilities(h, (const float*)h->lm_feat, (int)L, half, jump, d_lw, d_mean, s))) return rc;
  CU_TRY(h, cudaMemcpyAsync(h->pin_out, d_lw, out_floats * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (feat_out)
    CU_TRY(h, cudaMemcpyAsync((float*)h->pin_out + out_floats, h->lm_feat, feat_floats * sizeof(float),
                              cudaMemcpyDeviceToHost, s));
  CU_TRY(h, cudaStreamSynchronize(s));
  if (probs_LW) memcpy(probs_LW, h->pin_out, (size_t)L * W * sizeof(float));
  if (mean_L) memcpy(mean_L, (float*)h->pin_out + (size_t)L * W, (size_t)L * sizeof(float));
  if (feat_out) memcpy(feat_out, (float*)h->pin_out + out_floats, feat_floats * sizeof(float));
  return VADB_OK;
}

int vadb_attention(vadb_handle* h, const void* q, const void* k, const void* v, void* o, int dtype,
                   const int32_t* lengths, int B, int T, void* stream) {
  if (!h || !q || !k || !v || !o || B < 0 || T < 0) return VADB_E_INVALID;
  if (dtype != VADB_F32 && dtype != VADB_BF16) return fail(h, VADB_E_INVALID, "dtype must be f32 or bf16");
  if (B == 0 || T == 0) return VADB_OK;
  DeviceGuard dg(h->device);
  return attention(h, q, k, v, o, dtype, lengths, B, T, (cudaStream_t)stream);
}

}  // extern "C"
