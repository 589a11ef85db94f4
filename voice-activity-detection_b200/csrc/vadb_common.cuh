// Shared declarations for libvadb200 (B200 / sm_100a Self-Attentive VAD forward path).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/vadb200.h"

namespace vadb {

constexpr int D = 128;       // d_model == d_head (n_heads = 1): vad/models/self_attention.py:17-19
constexpr int DFF = 512;     // d_ff = 4*d_model: vad/models/self_attention.py:10
constexpr float LN_EPS = 1e-5f;

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  One forward is 14 back-to-back kernels, most of them persistent
// one-CTA-per-SM grids whose prologue (mbarrier init, TMEM allocation, tensor-map fetch, weight loads)
// does not depend on the previous kernel.  Every kernel is launched with the programmatic-stream-
// serialization attribute: its CTAs may start as soon as all CTAs of the previous kernel have called
// pdl_launch_dependents() (first thing they do) and an SM has room, run their prologue, and block in
// pdl_wait() until the previous grid has completed and its memory is visible.  Nothing produced by an
// earlier kernel may be touched before pdl_wait().  VADB_PDL=0 falls back to plain stream order.
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// float4 index of (row m, column quad c4) of the fp32 residual stream: tiled [tile][32 c4][128 rows] or row-major
__device__ __forceinline__ size_t h_quad_index(long m, int c4, int tiled) {
  return tiled ? (size_t)(m >> 7) * 4096 + (size_t)c4 * 128 + (size_t)(m & 127) : (size_t)m * 32 + (size_t)c4;
}
#endif
bool pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Offsets (in floats) of every tensor inside the packed fp32 blob; order == state_dict order.
struct LayerOffsets {
  size_t wq, bq, wk, bk, wv, bv, wo, bo, ln1_g, ln1_b, w1, b1, w2, b2, ln2_g, ln2_b;
};
struct BlobLayout {
  size_t w_in, b_in;
  std::vector<LayerOffsets> layers;
  size_t lnf_g, lnf_b, wc, bc, total;
};
BlobLayout make_layout(int F, int L);

// ---------------------------------------------------------------------------------------
// generic fp32 CUDA-core GEMM with fused prologue/epilogue (k_gemm_f32.cu)
//   C[m, n] = epi( sum_k pro(A)[m, k] * W[n, k] + bias[n] )
// ---------------------------------------------------------------------------------------
struct GemmArgs {
  const void* A;          // [M, K] row-major, fp32 or bf16
  int a_is_bf16;
  const float* W;         // [N, K] row-major (nn.Linear layout)
  const float* bias;      // [N]
  int M, N, K;
  // prologue: LayerNorm over K (requires K == 128)
  const float* ln_g;      // nullptr -> no LayerNorm
  const float* ln_b;
  // epilogue
  int relu;
  const float* residual;  // [M, N] fp32 or nullptr (may alias out[0])
  const float* pe;        // [T, N] already divided by sqrt(d), or nullptr
  int pe_T;               // row m uses pe[m % pe_T]
  // window gather (vad/predictor.py:182-218) folded into the A-row index: when win_W > 0,
  // output row m = i*win_W + k reads A row  half + i + rel[k]
  int win_W, win_half, win_jump;
  // outputs: columns [j*out_split, (j+1)*out_split) go to out[j] (row stride out_split)
  void* out[3];
  int out_split;          // == N when a single output
  int out_is_bf16;
};
cudaError_t launch_gemm_f32(const GemmArgs& a, cudaStream_t s);

// fp32 flash attention on CUDA cores (k_attn_f32.cu)
cudaError_t launch_attn_f32(const float* q, const float* k, const float* v, float* o,
                            const int32_t* lengths, int B, int T, cudaStream_t s);

// bf16 tcgen05 attention (k_attn_tc.cu)
cudaError_t launch_attn_tc(const bf16* q, const bf16* k, const bf16* v, bf16* o,
                           const int32_t* lengths, int B, int T, int num_sms, cudaStream_t s,
                           std::string* err);

// bf16 attention for T <= 8 (the 7-frame windows of the reference Predictor): one warp per pair of
// windows on mma.sync, no shared memory (k_attn_small.cu)
bool attn_small_supported(int T);
cudaError_t launch_attn_small(const bf16* q, const bf16* k, const bf16* v, bf16* o,
                              const int32_t* lengths, int B, int T, int num_sms, cudaStream_t s);

// tcgen05 GEMM for the per-frame Linears in bf16 mode (k_gemm_tc.cu)
struct GemmTcArgs {
  int M, N, K;
  const bf16* w_bf16;      // [N, K]
  const bf16* a_bf16;      // [M, K] (when ln_g == nullptr)
  const float* a_f32;      // [M, 128] fp32 rows through LayerNorm (when ln_g != nullptr)
  const float* a_f32_tma;  // front end, tf32 mode: [M, a_cols] fp32 rows loaded by TMA and multiplied as tf32 with
  const float* w_f32;      //   the fp32 weight [N, a_cols]; K must be 128 * ceil(a_cols / 64) (64-float stages)
  const void* a_rows;      // front end: [*, a_cols] fp32/bf16 feature rows, converted + zero-padded to K = 128
  int a_cols, a_rows_bf16;
  int win_W, win_half, win_jump;   // front end: window gather folded into the A-row index (win_W > 0)
  const float* ln_g;
  const float* ln_b;
  const float* bias;       // [N] fp32
  int relu;
  const float* residual;   // fp32 [M,128] or nullptr (may alias out[0])
  int res_mod;             // > 0: residual row = m % res_mod (positional-encoding table [T,128])
  int out_f32;             // fp32 [M,N] output, else bf16
  void* out[3];            // out[1], out[2] set -> columns split in 128-wide blocks (q, k, v)
  // fp32 N == 128 outputs: also write LayerNorm(out_row) (gamma/beta of the NEXT sublayer) as bf16 to out[1]
  const float* emit_ln_g;
  const float* emit_ln_b;
  // fp32 N == 128 output written in the TILED residual layout [tile][32 column quads][128 rows][4 floats]
  // (what k_tail_tc.cu reads and writes) instead of row-major
  int out_tiled;
};
cudaError_t launch_gemm_tc(const GemmTcArgs& a, int num_sms, cudaStream_t s, std::string* err);

// fused LN2 + FFN1 + ReLU + FFN2 + residual (k_ffn_tc.cu); h is updated in place
struct FfnTcArgs {
  int M;
  float* h;                // [M,128] fp32 residual stream (in/out)
  const bf16* a_ln;        // [M,128] bf16 LayerNorm(h) written by the previous kernel (A operand, via TMA)
  bf16* emit_out;          // optional: LayerNorm of the OUTPUT rows for the next layer's Q/K/V GEMM
  const float* emit_ln_g;
  const float* emit_ln_b;
  const bf16* w1_bf16;     // [512,128]
  const float* b1;         // [512]
  const bf16* w2_bf16;     // [128,512]
  const float* b2;         // [128]
  // last layer only: final LayerNorm + classifier + log-softmax / sigmoid fused into the epilogue
  // (h is then NOT written back); cls_ln_g == nullptr -> ordinary layer
  const float* cls_ln_g;
  const float* cls_ln_b;
  const float* cls_w;      // [2,128]
  const float* cls_bias;   // [2]
  float* prob;             // [M] or nullptr
  float* logp;             // [M,2] or nullptr
};
cudaError_t launch_ffn_tc(const FfnTcArgs& a, int num_sms, cudaStream_t s, std::string* err);

// fused layer tail: out-projection + residual + LN2 + FFN + residual + (LN1_next + Q/K/V of the next layer |
// final LayerNorm + classifier) in one kernel (k_tail_tc.cu); h is the TILED fp32 residual stream, updated in place
struct TailTcArgs {
  int M;
  const unsigned char* wpack;   // packed weight blocks of this layer (launch_tail_pack)
  const float* aux;             // folded biases / classifier constants of this layer (launch_tail_pack)
  const bf16* o;                // [M,128] attention output
  float* h;                     // tiled residual stream
  const float* bo; const float* b2;
  bf16* q; bf16* k; bf16* v;    // next layer's attention inputs; q == nullptr -> last layer: classifier epilogue
  float* prob; float* logp;
  // head of the model (x != nullptr): h0 = x W_in^T + b_in + PE/sqrt(d) and layer 0's q,k,v in the same kernel;
  // wpack = launch_head_pack's blocks for the feature dtype, aux = its folded q|k|v biases, bo = b_in, h = output
  const void* x; int x_is_bf16; int F;
  const float* pe_tiled; int pe_tiles;     // PE/sqrt(d) in the tiled layout, T / 128 tiles (needs T % 128 == 0)
};
// Load-time preparation of one layer: weight blocks in consumption order as the swizzled shared-memory image,
// LayerNorm gammas folded into the following weights (W diag(gamma)), betas into the following biases
// (b + W beta), W2 scaled by 1/2 (the kernel computes 2 ReLU), classifier folded with the final LayerNorm.
struct TailPackArgs {
  const float* wo; const float* w1; const float* b1; const float* w2;
  const float* ln2_g; const float* ln2_b;                       // this layer's feed-forward pre-LN
  const float* wqkv_next; const float* bqkv_next;               // fused [384,128] / [384] of layer l+1, nullptr: last layer
  const float* ln1n_g; const float* ln1n_b;                     // layer l+1's attention pre-LN
  const float* lnf_g; const float* lnf_b; const float* wc; const float* bc;   // last layer: final LN + classifier
};
size_t tail_pack_bytes();       // bytes of one layer's packed blocks
size_t tail_aux_floats();       // floats of one layer's folded biases / constants
cudaError_t launch_tail_pack(const TailPackArgs& a, unsigned char* dst, float* aux, cudaStream_t s);
cudaError_t launch_tail_tc(const TailTcArgs& a, int num_sms, cudaStream_t s, std::string* err);
cudaError_t launch_head_pack(const float* w_in, int F, const float* wqkv0, const float* bqkv0, const float* ln1_g,
                             const float* ln1_b, unsigned char* dst_bf16, unsigned char* dst_tf32, float* aux384,
                             cudaStream_t s);

// final LayerNorm + classifier + sigmoid/log-softmax (k_classifier.cu); tiled: h in the tiled residual layout
cudaError_t launch_classifier(const float* h, const float* g, const float* b, const float* wc,
                              const float* bc, int M, float* prob, float* logp, int tiled, cudaStream_t s);

// window path helpers (k_window.cu)
cudaError_t launch_boost(const float* prob_nW, int L, int half, int jump, int W,
                         float* probs_LW, float* mean_L, cudaStream_t s);

// rows of the projected clip gathered into context windows + PE + LayerNorm_1 emit (k_window.cu)
cudaError_t launch_window_gather_ln(const float* proj, const float* pe, float* out_h, bf16* out_ln,
                                    const float* ln_g, const float* ln_b, long n_rows, int W, int half,
                                    int jump, int tiled, cudaStream_t s);
// row-major fp32 residual rows -> tiled layout + LayerNorm(row) in bf16 (fallback front ends, k_window.cu)
cudaError_t launch_retile_ln(const float* h_rows, float* h_tiled, bf16* out_ln, const float* ln_g,
                             const float* ln_b, long n_rows, cudaStream_t s);

// log-mel front end (k_logmel.cu): host-built tables and the per-frame FFT + mel + log kernel
void logmel_tables(int sr, int n_fft, int win, int n_mels, std::vector<float>* fb_dense, std::vector<double>* window);
cudaError_t launch_logmel(const float* audio, long n_samples, int n_fft, int hop, long n_frames, const double* window,
                          const double* twiddle, const int* fb_start, const int* fb_len, const int* fb_off,
                          const float* fb_w, int n_mels, float* out, cudaStream_t s);

// length-bucketed forward (k_window.cu): clip rows [ids[i], 0:Tk] of a [B, T, row_bytes] batch -> dense [n, Tk, row_bytes],
// and the bucket's outputs back to the [B, T] / [B, T, 2] layout
cudaError_t launch_gather_clips(const void* x, void* out, const int32_t* ids, int n, int T, int Tk, int row_bytes,
                                cudaStream_t s);
cudaError_t launch_scatter_clips(const float* prob_k, const float* logp_k, float* prob, float* logp, const int32_t* ids,
                                 int n, int T, int Tk, cudaStream_t s);

// misc element-wise (k_window.cu)
cudaError_t launch_pad_rows_bf16(const float* in, bf16* out, int rows, int cols, cudaStream_t s);
cudaError_t launch_f32_to_bf16(const float* in, bf16* out, size_t n, cudaStream_t s);

}  // namespace vadb
