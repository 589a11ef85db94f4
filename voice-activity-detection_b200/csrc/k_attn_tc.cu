// Fused scaled-dot-product attention over the frame axis on the 5th-gen tensor cores (sm_100a).
//
//   O[b] = softmax(Q[b] K[b]^T / sqrt(128) + keymask(lengths[b])) V[b]        bf16 in, bf16 out
//
// Reference semantics: vad/modeling/transformer.py:351-363 (scaled_dot_product), :319-325 (key
// padding mask), :333 (softmax over keys), :338-346 (P V and the head merge) -- the three
// [B,1,T,T] fp32 score tensors the reference materialises never leave the SM.
//
// Persistent CTAs (one per SM, 11 warps) walk a static list of work items; one item = one clip x 256 query
// rows = two 128-row query tiles (the last, partial round is split into single-tile items, and odd CTAs
// run theirs first so that the CTAs do not all hit their item boundaries -- and the memory system --
// together, see ItemWalk):
//   warp 8      TMA producer: Q (once per item), 3-stage rings of 64-key K tiles and V tiles, all
//               128B-swizzled; K runs two steps ahead of V
//   warps 9,10  one MMA-issuing warp per query tile (one elected lane): S_t = Q_t K^T and O_t += P_t V via
//               tcgen05.mma with accumulators in TMEM -- S double-buffered per query tile (4 x 64
//               columns), O 2 x 128 columns.  Q K^T of step j+2 is issued right behind P V of step j, so
//               the scores of the next step are already in TMEM when a softmax warpgroup finishes a step.
//   warps 0-3   softmax warpgroup of query tile 0: one thread per query row (TMEM lane),
//   warps 4-7   softmax warpgroup of query tile 1   tcgen05.ld S -> online softmax (fp32, exp2, lazy
//               rescale of O in TMEM) -> P (bf16) written back over the S buffer in TMEM and consumed
//               by the P V MMA from there (TS form); steps are software-pipelined (the scores of step
//               j+1 are loaded while the P hand-off of step j is in flight) and the two warpgroups take
//               turns on the SFU through a named-barrier token
// The two query tiles share every K/V tile and ping-pong on the tensor pipe.
//
// Algorithmic traffic per launch: read Q,K,V once + write O once = 4*B*T*128*2 bytes (SURVEY.md
// section 8d); K/V re-reads by the other query pairs of a clip are served from L2.
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include <utility>
#include <vector>

#include "tc_common.cuh"
#include "vadb_common.cuh"

namespace vadb {
namespace {

using namespace tc;

constexpr int BM = 128;          // query rows per tile (UMMA M)
constexpr int BKV = 64;          // keys per K/V tile
constexpr int NK = 3;            // K ring depth (K(j+2) is requested once Q K^T(j-1) retired)
constexpr int NV = 3;            // V ring depth (V(j) is requested once P V(j-3) retired: three steps of latency cover)
constexpr int NTHREADS = 352;    // 8 softmax warps + TMA producer + one MMA-issuing warp per query tile

constexpr uint32_t Q_HALF_BYTES = BM * 128;          // [128 rows x 64 d] bf16 = 16 KB
constexpr uint32_t Q_TILE_BYTES = 2 * Q_HALF_BYTES;  // two d-halves
constexpr uint32_t KV_HALF_BYTES = BKV * 128;        // [64 keys x 64 d] = 8 KB
constexpr uint32_t KV_TILE_BYTES = 2 * KV_HALF_BYTES;
constexpr uint32_t P_TILE_BYTES = BM * 128;          // [128 rows x 64 keys] bf16 = 16 KB

constexpr uint32_t OFF_Q = 0;
constexpr uint32_t OFF_K = OFF_Q + 2 * Q_TILE_BYTES;
constexpr uint32_t OFF_V = OFF_K + NK * KV_TILE_BYTES;
constexpr uint32_t OFF_OST = OFF_V + NV * KV_TILE_BYTES;        // O staging: [tile][d half] x 16 KB
constexpr uint32_t OFF_BAR = OFF_OST + 4 * P_TILE_BYTES;
constexpr uint32_t OFF_SCR = OFF_BAR + 256;        // 1 KB unused, then 64 bytes: zero word (+0), TMEM base (+4),
                                                   // 8 dummy words (token pinning, +32)
constexpr uint32_t SMEM_BYTES = OFF_SCR + 1024 + 64;
constexpr uint32_t SMEM_ALLOC = SMEM_BYTES + 1024;   // slack for 1024-byte alignment

// barrier slots (8 bytes each) at OFF_BAR
enum { B_QFULL = 0, B_KFULL = 1, B_KEMPTY = 5, B_VFULL = 9, B_VEMPTY = 12, B_SFULL = 15 /* [t][buf] */,
       B_PFULL = 19, B_OFULL = 21, B_QEMPTY = 23, B_QFULL1 = 24, B_QEMPTY1 = 25, B_COUNT = 26 };
static_assert(8 * B_COUNT <= 256, "barrier region");
static_assert(8 * B_COUNT <= 256, "barrier region");
static_assert(NK <= 4 && NV <= 3, "barrier slots");

constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_S = 0;      // S_t[buf] at columns (2 t + buf) * 64; P_t (bf16 pairs, 32 columns)
                                  // overwrites the head of the S buffer it was computed from
constexpr uint32_t TM_O = 256;    // O_t at columns 256 + 128 t

constexpr uint32_t IDESC_QK = idesc_bf16(128, BKV, 0, 0);   // A = Q (K-major), B = K (K-major)
constexpr uint32_t IDESC_PV = idesc_bf16(128, 128, 0, 1);   // A = P (K-major), B = V (MN-major)

constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units: rescale O only when the max grew > 2^8

// TRACE: developer instrumentation -- CTA 0 records clock64() stamps of the pipeline events of its
// second work item into `trace` (VADB_ATTN_TRACE=1); compiled out of the production instantiation.
#define TR(slot) do { if (TRACE && blockIdx.x == 0 && n_item == 1 && trace) trace[(slot)] = clock64(); } while (0)

struct Item { int b, q0, len, nkv, ntile; bool valid; };

__device__ __forceinline__ void nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// named barriers: 1/2 = MUFU token (softmax warpgroup 0 / 1 may run its exponentials), 3+t = warpgroup t
enum { NB_TOKEN0 = 1, NB_TOKEN1 = 2, NB_WG = 3 /* +t */ };

// Work items.  idx < n_full: one clip x 256 query rows (two tiles sharing K/V).  The pairs of the last,
// partially filled round are split into single-tile items (idx >= n_full) so that the tail of the launch
// spreads over twice as many SMs.
__device__ __forceinline__ Item get_item(int idx, int n_full, int npairs, int T, const int32_t* lengths) {
  Item it;
  int pair = idx, tile = -1;
  if (idx >= n_full) { pair = n_full + ((idx - n_full) >> 1); tile = (idx - n_full) & 1; }
  it.b = pair / npairs;
  it.q0 = (pair % npairs) * 2 * BM;
  it.ntile = (it.q0 + BM < T) ? 2 : 1;        // second query tile entirely past T: skip it
  it.valid = true;
  if (tile >= 0) {
    if (tile == 1 && it.ntile == 1) it.valid = false;
    it.q0 += tile * BM;
    it.ntile = 1;
  }
  int len = lengths ? lengths[it.b] : T;
  it.len = min(max(len, 0), T);
  it.nkv = (it.len + BKV - 1) / BKV;
  return it;
}

// VAR: softmax-schedule variant bits (selected on the host, launch_attn_tc):
//   bit 0  hand the SFU token to the other warpgroup after element EARLY_IDX of a step instead of after
//          the last one: its barrier / LDS / first-FFMA start-up latency overlaps our last exponentials
//   (bit 1, P stored in two 32-key halves, was measured without effect and removed: the rare redo of a
//          step re-reads the scores from TMEM, which the early half-store would have overwritten)
//   bits 2-3  fraction of the exponentials evaluated on the FMA pipe (Cody-Waite + degree-3 minimax,
//          rel. error 7.5e-5, far below the bf16 rounding of P) instead of the SFU: 0, 1/4, 3/8, 1/2
//   bit 4  no SFU token (the warpgroups run free)
//   bits 5-6  element index of the early hand-over: 46, 30, 16
#define VADB_ATTN_VARIANTS(X) X(0) X(1)
constexpr int ATTN_DEFAULT_VARIANT = 0;      // re-timed at the end of round 2: 0 is 0.5 % faster than the early hand-over (1)
constexpr int ATTN_DEFAULT_STAGGER = 1;
template <int VAR> __device__ __forceinline__ bool use_poly(int i) {
  constexpr int f = (VAR >> 2) & 3;
  return f == 1 ? (i & 3) == 3 : f == 2 ? ((i & 7) == 1 || (i & 7) == 4 || (i & 7) == 7) : f == 3 ? (i & 1) == 1 : false;
}

// Order in which a CTA walks its work items.  Every CTA has the same amount of work per item, so with the
// plain order (round r -> item blockIdx + r * grid) all 148 CTAs reach their item boundaries together and
// their Q/K/V requests for the next item (96 KB each) arrive as one 14 MB burst.  When the last round
// consists of the shorter single-tile items, odd CTAs run theirs FIRST: odd and even CTAs are then offset
// by the difference between a single-tile and a two-tile item for the rest of the launch.
struct ItemWalk {
  int first, n, rot;    // first item index, number of items of this CTA, 1 -> the last item goes first
  __device__ __forceinline__ ItemWalk(int n_items, int n_full, int stagger) {
    first = blockIdx.x;
    n = first < n_items ? (n_items - first + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int last = first + (n - 1) * (int)gridDim.x;
    rot = (stagger && n > 1 && (blockIdx.x & 1) && last >= n_full) ? 1 : 0;
  }
  __device__ __forceinline__ int idx(int r) const {
    const int rr = rot ? (r == 0 ? n - 1 : r - 1) : r;
    return first + rr * (int)gridDim.x;
  }
};

// Rare path of a softmax step (the running max grew by more than 2^RESCALE_THRESHOLD): redo the step
// against the new reference max from the scores that are still in TMEM (P has not been stored yet), eight
// keys per trip; P chunk c (4 columns) lands on score columns that were read in trips <= c.  Deliberately
// NOT inlined: a second inlined copy of the unrolled exponential code sat as ~7 KB of cold instructions in
// the middle of every step (a shared copy with the redo looping back spills ~30 registers: +25 %).
// Returns the row sum of the new P.
static __device__ __noinline__ float redo_step_from_tmem(uint32_t tsb, float m_new, float c, int kbase, int len) {
  float ps = 0.f;
#pragma unroll 1
  for (int cch = 0; cch < 8; ++cch) {
    uint32_t sx[8], px[4];
    tmem_ld8(tsb + cch * 8, sx);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const int key = kbase + cch * 8 + i;
      const float s0 = key < len ? __uint_as_float(sx[i]) : -CUDART_INF_F;
      const float s1 = key + 1 < len ? __uint_as_float(sx[i + 1]) : -CUDART_INF_F;
      const float p0 = fast_exp2(fmaf(s0, c, -m_new)), p1 = fast_exp2(fmaf(s1, c, -m_new));
      ps += p0 + p1;
      px[i >> 1] = pack_bf16(p0, p1);
    }
    tmem_st4(tsb + cch * 4, px);
  }
  return ps;
}

template <int VAR> constexpr int early_idx() { return ((VAR >> 5) & 3) == 0 ? 46 : ((VAR >> 5) & 3) == 1 ? 30 : 16; }

template <bool TRACE, int VAR>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
               const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
               bf16* __restrict__ O, const int32_t* __restrict__ lengths, int T, int npairs, int n_items,
               int n_full, int stagger, long long* trace) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + OFF_BAR;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_SCR + 1024 + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (TRACE && blockIdx.x == 0 && threadIdx.x == 0 && trace) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    trace[2040] = clock64();
    trace[2042] = (long long)gt;
  }
  if (threadIdx.x == 0) {
    // the Q buffer of each query tile is its own single-slot channel: tile 0's Q of the next item is
    // requested as soon as tile 0's last Q K^T has retired, without waiting for tile 1
    mbar_init(BAR(B_QFULL), 1); mbar_init(BAR(B_QFULL1), 1);
    mbar_init(BAR(B_QEMPTY), 1); mbar_init(BAR(B_QEMPTY1), 1);
    for (int s = 0; s < NK; ++s) { mbar_init(BAR(B_KFULL + s), 1); mbar_init(BAR(B_KEMPTY + s), 2); }
    for (int s = 0; s < NV; ++s) { mbar_init(BAR(B_VFULL + s), 1); mbar_init(BAR(B_VEMPTY + s), 2); }
    for (int i = 0; i < 4; ++i) mbar_init(BAR(B_SFULL + i), 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(BAR(B_PFULL + t), 4);          // one arrival per softmax warp
      mbar_init(BAR(B_OFULL + t), 1);
    }
    mbar_fence_init();
    *reinterpret_cast<volatile float*>(smem_gen + OFF_SCR + 1024) = 0.f;
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
  }
  if (warp == 9) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();                 // the set-up above overlaps the previous kernel's tail (vadb_common.cuh)

  if (warp == 8) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int kc = 0, vc = 0, nq[2] = {0, 0};      // K tiles / V tiles / Q loads (per query tile) issued so far
      const ItemWalk walk(n_items, n_full, stagger);
      for (int r = 0; r < walk.n; ++r) {
        const int idx = walk.idx(r);
        const Item it = get_item(idx, n_full, npairs, T, lengths);
        if (!it.valid || it.nkv == 0) continue;
        auto load_k = [&](int j) {
          const int ks = (kc + j) % NK;
          mbar_wait(BAR(B_KEMPTY + ks), (((kc + j) / NK) & 1) ^ 1, 2);
          const uint32_t kdst = smem_base + OFF_K + ks * KV_TILE_BYTES;
          mbar_arrive_expect_tx(BAR(B_KFULL + ks), KV_TILE_BYTES);
          tma_load_3d(kdst, &tm_k, BAR(B_KFULL + ks), 0, j * BKV, it.b);
          tma_load_3d(kdst + KV_HALF_BYTES, &tm_k, BAR(B_KFULL + ks), 64, j * BKV, it.b);
        };
        auto load_q = [&](int t) {
          const uint32_t qe = BAR(t == 0 ? B_QEMPTY : B_QEMPTY1), qf = BAR(t == 0 ? B_QFULL : B_QFULL1);
          mbar_wait(qe, (nq[t] & 1) ^ 1, 1);                  // every Q_t K^T of the previous item retired
          mbar_arrive_expect_tx(qf, Q_TILE_BYTES);
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(smem_base + OFF_Q + t * Q_TILE_BYTES + hf * Q_HALF_BYTES, &tm_q, qf, hf * 64,
                        it.q0 + t * BM, it.b);
          ++nq[t];
        };
        // Q of tile 0 and the first K tile first (all the first Q K^T needs), then tile 1's Q; K runs two
        // steps ahead of V, in the order the MMA warps consume them
        load_q(0);
        load_k(0);
        if (it.ntile > 1) load_q(1);
        if (it.nkv > 1) load_k(1);
        for (int j = 0; j < it.nkv; ++j) {
          if (j + 2 < it.nkv) load_k(j + 2);
          const int vs = (vc + j) % NV;
          mbar_wait(BAR(B_VEMPTY + vs), (((vc + j) / NV) & 1) ^ 1, 3);
          const uint32_t vdst = smem_base + OFF_V + vs * KV_TILE_BYTES;
          mbar_arrive_expect_tx(BAR(B_VFULL + vs), KV_TILE_BYTES);
          tma_load_3d(vdst, &tm_v, BAR(B_VFULL + vs), 0, j * BKV, it.b);
          tma_load_3d(vdst + KV_HALF_BYTES, &tm_v, BAR(B_VFULL + vs), 64, j * BKV, it.b);
        }
        kc += it.nkv; vc += it.nkv;
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ======================= MMA issuers (one warp per query tile) =======================
    // Each query tile has its own issuing warp, so the chain  P_t(j) ready -> P V_t(j), Q K^T_t(j+2)
    // of one tile never queues behind the (blocking, tensor-pipe back-pressured) issue of the other
    // tile; with a single issuer the pipe idled ~600 of every ~1950 cycles (profiles/r1_attention_notes.md).
    // The whole warp walks the schedule (waits included) and one elected lane issues, so the
    // descriptor arithmetic stays warp-uniform and costs an add or two per MMA.  The K/V "empty"
    // barriers count one arrival per issuing warp; on single-tile items the idle warp keeps the
    // protocol uniform with plain arrivals (always behind its own wait on the matching "full" barrier,
    // so it cannot run a phase ahead).  Each query tile's Q buffer has its own full/empty pair.
    const int t = warp - 9;
    const uint32_t q_lo = desc_lo(smem_base + OFF_Q, 16);
    const uint32_t k_lo = desc_lo(smem_base + OFF_K, 16);
    const uint32_t v_lo = desc_lo(smem_base + OFF_V, KV_HALF_BYTES);   // LBO = stride between d halves
    auto issue_qk = [&](int ks, int buf) {
      // S_t[buf][128 x 64] = Q_t[128 x 128] K[64 x 128]^T : 8 MMAs of K = 16 (two 64-wide d halves)
      const uint32_t qa = q_lo + (uint32_t)t * (Q_TILE_BYTES >> 4);
      const uint32_t kb = k_lo + (uint32_t)ks * (KV_TILE_BYTES >> 4);
      const uint32_t d = tmem_base + TM_S + (uint32_t)(2 * t + buf) * 64;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss_lh(d, qa + hf * (Q_HALF_BYTES >> 4) + kk * 2, kb + hf * (KV_HALF_BYTES >> 4) + kk * 2,
                     DESC_HI_SW128, IDESC_QK, (hf | kk) != 0 ? 1u : 0u);
    };
    auto issue_pv = [&](int vs, int buf, uint32_t accumulate) {
      // O_t[128 x 128] += P_t[128 x 64] V[64 x 128] : 4 MMAs of K = 16 keys, N = 128 (two d halves);
      // P_t is read from TMEM (the head of S_t[buf]), V from shared memory
      const uint32_t pa = tmem_base + TM_S + (uint32_t)(2 * t + buf) * 64;
      const uint32_t vb = v_lo + (uint32_t)vs * (KV_TILE_BYTES >> 4);
      const uint32_t d = tmem_base + TM_O + 128u * t;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_ts_lh(d, pa + kk * 8, vb + kk * (2048 >> 4), DESC_HI_SW128, IDESC_PV,
                   kk != 0 ? 1u : accumulate);
    };
    // up to three barrier waits in parallel (lanes 0..2 spin on one barrier each): the latency of a
    // step's formal waits is their maximum, not their sum
    auto wait_par = [&](uint32_t b0, uint32_t p0, uint32_t b1, uint32_t p1, uint32_t b2, uint32_t p2, int n) {
      if (lane < n) mbar_wait(lane == 0 ? b0 : lane == 1 ? b1 : b2, lane == 0 ? p0 : lane == 1 ? p1 : p2, 6);
      __syncwarp();
    };
    int kc = 0, vc = 0, nq = 0, n_item = 0;
    int g = 0;                                  // softmax steps issued so far for this query tile
    const ItemWalk walk(n_items, n_full, stagger);
    for (int r = 0; r < walk.n; ++r, ++n_item) {
      const int idx = walk.idx(r);
      const Item it = get_item(idx, n_full, npairs, T, lengths);
      if (!it.valid || it.nkv == 0) continue;
      const int nkv = it.nkv;
      const bool live = t < it.ntile;           // this warp's query tile exists in the item
      if (t == 0) TR(0);
      const uint32_t q_empty = BAR(t == 0 ? B_QEMPTY : B_QEMPTY1);
      if (live) {
        mbar_wait(BAR(t == 0 ? B_QFULL : B_QFULL1), nq & 1, 4);
        ++nq;
      }
      // prologue: scores of steps 0 and 1
      for (int j = 0; j < 2 && j < nkv; ++j) {
        const int ks = (kc + j) % NK;
        mbar_wait(BAR(B_KFULL + ks), ((kc + j) / NK) & 1, 5);
        tc_fence_after();
        if (live) {
          if (elect_one()) {
            issue_qk(ks, (g + j) & 1);
            umma_commit(BAR(B_SFULL + 2 * t + ((g + j) & 1)));
            umma_commit(BAR(B_KEMPTY + ks));
            if (j == nkv - 1) umma_commit(q_empty);
          }
        } else if (lane == 0) {
          mbar_arrive(BAR(B_KEMPTY + ks));
        }
        __syncwarp();
      }
      if (t == 0) TR(1);
      for (int j = 0; j < nkv; ++j) {
        const int vs = (vc + j) % NV, kn = (kc + j + 2) % NK;
        const bool more = j + 2 < nkv;
        TR(64 + j * 16 + t * 4 + 0);
        if (live) {
          // V(j) and K(j+2) landed (formal: requested three steps ago; checked while P_t(j) is still
          // being computed), then the critical wait alone: P_t(j) in TMEM (S_t consumed, O_t corrected)
          wait_par(BAR(B_VFULL + vs), ((vc + j) / NV) & 1, BAR(B_KFULL + kn), ((kc + j + 2) / NK) & 1, 0, 0,
                   more ? 2 : 1);
          mbar_wait(BAR(B_PFULL + t), (g + j) & 1, 8);
          tc_fence_after();
          TR(64 + j * 16 + t * 4 + 1);
          if (elect_one()) {
            issue_pv(vs, (g + j) & 1, j > 0 ? 1u : 0u);
            umma_commit(BAR(B_OFULL + t));
            umma_commit(BAR(B_VEMPTY + vs));
            if (more) {
              issue_qk(kn, (g + j) & 1);
              umma_commit(BAR(B_SFULL + 2 * t + ((g + j) & 1)));
              umma_commit(BAR(B_KEMPTY + kn));
              if (j + 2 == nkv - 1) umma_commit(q_empty);   // last Q_t K^T of this item issued
            }
          }
        } else {
          wait_par(BAR(B_VFULL + vs), ((vc + j) / NV) & 1, BAR(B_KFULL + kn), ((kc + j + 2) / NK) & 1, 0, 0,
                   more ? 2 : 1);
          if (lane == 0) {
            mbar_arrive(BAR(B_VEMPTY + vs));
            if (more) {
              mbar_arrive(BAR(B_KEMPTY + kn));
            }
          }
        }
        __syncwarp();
        TR(64 + j * 16 + t * 4 + 3);
      }
      kc += nkv; vc += nkv;
      if (live) g += nkv;
    }
  } else {
    // ======================= softmax warpgroups =======================
    const int t = warp >> 2;                       // query tile of this warpgroup
    const int row = (warp & 3) * 32 + lane;        // row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t ts = tmem_base + lane_addr + TM_S + 128 * t;
    const uint32_t to = tmem_base + lane_addr + TM_O + 128 * t;
    unsigned char* ost = smem_gen + OFF_OST + t * 2 * P_TILE_BYTES;
    volatile float* scratch = reinterpret_cast<volatile float*>(smem_gen + OFF_SCR + 1024);   // [0] = 0.f, [8..15] dummies
    // softmax in the exp2 domain: p = 2^(s*c - m), c = log2(e)/sqrt(d_head)  (transformer.py:362)
    const float c = 1.4426950408889634f * 0.08838834764831845f;
    int g = 0, n_item = 0;                         // softmax steps done so far by this warpgroup
    // The two warpgroups take turns on the SFU: exp2 throughput (16/clk/SM) is the scarce resource of
    // this kernel, so their exponential phases are serialised with a token and everything else
    // (TMEM traffic, row max, barrier traffic) of one overlaps the exponentials of the other.
    if (t == 1) nbar_arrive(NB_TOKEN0, 256);       // warpgroup 0 goes first

#define TS(k) do { if (TRACE && (threadIdx.x & 127) == 0) TR(512 + j * 32 + t * 16 + (k)); } while (0)
    const ItemWalk walk(n_items, n_full, stagger);
    for (int r = 0; r < walk.n; ++r, ++n_item) {
      const int idx = walk.idx(r);
      const Item it = get_item(idx, n_full, npairs, T, lengths);
      const int nkv = it.nkv, len = it.len;
      if (!it.valid) continue;
      if (nkv == 0) {
        // every key masked: softmax of all -inf -> NaN rows, as the reference produces
        const int qrow = it.q0 + t * BM + row;
        bf16* orow = O + ((long)it.b * T + qrow) * D;
        if (qrow < T)
          for (int cgi = 0; cgi < D / 8; ++cgi)
            *reinterpret_cast<uint4*>(orow + cgi * 8) = make_uint4(0x7FC07FC0u, 0x7FC07FC0u, 0x7FC07FC0u, 0x7FC07FC0u);
        continue;
      }
      if (t >= it.ntile) continue;
      const bool pingpong = it.ntile == 2 && !(VAR & 16);
      float m_ref = -CUDART_INF_F;     // reference max (log2 domain) the stored P/O are relative to
      float l_sum = 0.f;

      // Software-pipelined steps.  The scores of step j+1 are pulled from TMEM into registers while the
      // P hand-off of step j (tcgen05.st -> wait -> fence -> arrive) is still in flight, so the serial
      // chain of a warpgroup per step is  token -> exponentials -> P store -> arrive  and nothing else;
      // every other latency (S barrier probe, TMEM load, O barrier probe) sits in the shadow of the
      // other warpgroup's exponentials.  One load site only (sv is written in exactly one place).
      uint32_t sv[2][32];
      for (int j = -1; j < nkv; ++j) {
        const bool work = j >= 0, has_next = j + 1 < nkv;
        const uint32_t tsb = ts + (g & 1) * 64;
        uint32_t pk[32];
        bool o_ready = true, rescale = false, s_ready = false;
        if (work) {
          TS(0);
          // (the two non-blocking barrier probes of a step -- "P V of the previous step retired", "scores of the next
          // step are in TMEM" -- are issued half way through the exponentials, in the shadow of the SFU, instead of
          // on the serial chain in front of / behind them: see the exps lambda)
          const int kbase = j * BKV;
          if (kbase + BKV > len) {           // warp-uniform: only the last tile holds masked keys
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (kbase + h2 * 32 + i >= len) sv[h2][i] = 0xFF800000u;   // -inf
          }
          float psum, p_last, mxl;
          auto exps = [&](float m_use, bool pinned) {
            // exponentials, running sum, bf16 packing and (on otherwise idle issue slots in the shadow of
            // the SFU) the row max of this tile, which decides about a rescale afterwards
            float ps4[4] = {0.f, 0.f, 0.f, 0.f};
            float mx4[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
            for (int i = 0; i < 64; i += 2) {
              const float s0 = __uint_as_float(sv[i >> 5][i & 31]), s1 = __uint_as_float(sv[(i + 1) >> 5][(i + 1) & 31]);
              const float a0 = fmaf(s0, c, -m_use);
              const float a1 = fmaf(s1, c, -m_use);
              const float p0 = use_poly<VAR>(i) ? poly_exp2(a0) : fast_exp2_pinned(a0);
              const float p1 = use_poly<VAR>(i + 1) ? poly_exp2(a1) : fast_exp2_pinned(a1);
              ps4[(i >> 1) & 3] += p0 + p1;
              mx4[(i >> 1) & 3] = fmaxf(mx4[(i >> 1) & 3], fmaxf(s0, s1));
              pk[i >> 1] = pack_bf16(p0, p1);
              if (i == 62) p_last = p1;
              if (i == 32 && pinned) {
                o_ready = (j == 0) || mbar_test_wait(BAR(B_OFULL + t), (g - 1) & 1);
                s_ready = has_next && mbar_test_wait(BAR(B_SFULL + 2 * t + ((g + 1) & 1)), ((g + 1) >> 1) & 1);
              }
              if ((VAR & 1) && pinned && i == early_idx<VAR>() && pingpong) {
                scratch[8 + (threadIdx.x & 7)] = p0;
                nbar_arrive(t == 0 ? NB_TOKEN1 : NB_TOKEN0, 256);
              }
            }
            psum = (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
            mxl = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * c;
          };
          // The exponentials of step j > 0 start from the PREVIOUS reference max; only if the max grew by
          // more than 2^8 is the step redone against the new reference (rare).
          if (j == 0) {
            float mx4[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              mx4[(i >> 1) & 3] = fmaxf(mx4[(i >> 1) & 3], fmaxf(__uint_as_float(sv[0][i]), __uint_as_float(sv[0][i + 1])));
              mx4[((i >> 1) + 2) & 3] = fmaxf(mx4[((i >> 1) + 2) & 3], fmaxf(__uint_as_float(sv[1][i]), __uint_as_float(sv[1][i + 1])));
            }
            m_ref = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * c;
          }
          TS(1);
          float m_use = m_ref;
          if (pingpong) {
            nbar_sync(t == 0 ? NB_TOKEN0 : NB_TOKEN1, 256);
            // data dependence on a load issued after the barrier: keeps ptxas from hoisting the
            // exponentials above the token wait
            m_use += scratch[0];
          }
          TS(2);
          float alpha = 1.f;
          exps(m_use, true);
          if (pingpong && !(VAR & 1)) {
            // ... and the token is handed over once the last exponential has been through the SFU
            scratch[8 + (threadIdx.x & 7)] = p_last;
            nbar_arrive(t == 0 ? NB_TOKEN1 : NB_TOKEN0, 256);
          }
          rescale = (j > 0) && __any_sync(0xffffffffu, mxl > m_ref + RESCALE_THRESHOLD);
          if (rescale) {                               // warp-uniform, rare
            // redo against the new reference max, out of line (see redo_step_from_tmem)
            const float m_new = fmaxf(m_ref, mxl);
            alpha = fast_exp2(m_ref - m_new);
            m_ref = m_new;
            psum = redo_step_from_tmem(tsb, m_new, c, j * BKV, len);
          }
          TS(3);
          l_sum = l_sum * alpha + psum;
          // previous P V must have retired before O is rescaled (P itself lives in this step's S buffer)
          if (rescale) {
            if (!o_ready) mbar_wait(BAR(B_OFULL + t), (g - 1) & 1, 10);
            tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < 4; ++cb) {
              uint32_t ov[32];
              tmem_ld32(to + cb * 32, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
              tmem_st32(to + cb * 32, ov);
            }
          }
          if (!rescale) tmem_st32(tsb, pk);            // P_t (bf16 pairs) over the head of S_t[buf]
          TS(4);
        }
        // scores of the next step: TMEM -> registers.  If they are already there the load is issued now and
        // its latency overlaps the P hand-off; if not, P is handed over first (it must never wait for S:
        // the MMA warp needs P_t(j) to get to Q K^T(j+2)) and the scores are awaited afterwards.
        const int gn = work ? g + 1 : g;
        const uint32_t tsn = ts + (gn & 1) * 64;
        bool s_loaded = false;
        if (has_next && work && s_ready) {
          tc_fence_after();
          tmem_ld32(tsn, sv[0]);
          tmem_ld32(tsn + 32, sv[1]);
          s_loaded = true;
        }
        if (work) {
          TS(5);
          // An mbarrier may run at most one phase ahead of its waiter: do not signal P_t(j) before the
          // MMA warp has consumed P_t(j-1) (it has once P V(j-1) retired).  The probe was issued at the
          // top of the step, so this is normally free.
          if (!o_ready && !rescale) mbar_wait(BAR(B_OFULL + t), (g - 1) & 1, 12);
          tmem_st_wait();
          TS(6);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_PFULL + t));
          ++g;
        }
        if (has_next && !s_loaded) {
          mbar_wait(BAR(B_SFULL + 2 * t + (gn & 1)), (gn >> 1) & 1, 9);
          tc_fence_after();
          tmem_ld32(tsn, sv[0]);
          tmem_ld32(tsn + 32, sv[1]);
        }
        if (has_next) tmem_ld_wait();
        if (work) TS(7);
      }

      // epilogue: O_t / l -> bf16 -> two 128B-swizzled staging tiles (64 columns each) -> TMA stores in
      // full lines; rows past T are clipped by the tensor map.  The next item's first P V (which
      // overwrites O_t) is ordered behind these TMEM reads by this warp's own next p_full arrival.
      if (TRACE && (threadIdx.x & 127) == 0) TR(32 + t * 4 + 0);
      mbar_wait(BAR(B_OFULL + t), (g - 1) & 1, 11);
      tc_fence_after();
      if (TRACE && (threadIdx.x & 127) == 0) TR(32 + t * 4 + 1);
      const float inv = 1.0f / l_sum;
      // the staging tiles were handed to the TMA engine one whole item ago: formal guarantee only
      if ((threadIdx.x & 127) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      nbar_sync(NB_WG + t, 128);
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t ov[2][32];
        tmem_ld32(to + hf * 64, ov[0]);
        tmem_ld32(to + hf * 64 + 32, ov[1]);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint4 o4;
          const uint32_t* src = &ov[ch >> 2][(ch & 3) * 8];
          o4.x = pack_bf16_alu(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
          o4.y = pack_bf16_alu(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
          o4.z = pack_bf16_alu(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
          o4.w = pack_bf16_alu(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
          *reinterpret_cast<uint4*>(ost + hf * P_TILE_BYTES + sw128_offset(row, ch)) = o4;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      nbar_sync(NB_WG + t, 128);
      if ((threadIdx.x & 127) == 0) {
        const uint32_t src = smem_base + OFF_OST + t * 2 * P_TILE_BYTES;
        tma_store_3d(&tm_o, src, 0, it.q0 + t * BM, it.b);
        tma_store_3d(&tm_o, src + P_TILE_BYTES, 64, it.q0 + t * BM, it.b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (TRACE && (threadIdx.x & 127) == 0) TR(32 + t * 4 + 2);
    }
  }

  if (warp < 8 && (threadIdx.x & 127) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (TRACE && blockIdx.x == 0 && threadIdx.x == 0 && trace) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    trace[2041] = clock64();
    trace[2043] = (long long)gt;
  }
}

}  // namespace

CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* base,
                      const cuuint64_t* gdim, const cuuint64_t* gstride, const cuuint32_t* box,
                      const cuuint32_t* estr, CUtensorMapSwizzle swz) {
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
      return CUDA_ERROR_NOT_FOUND;
    fn = reinterpret_cast<Fn>(p);
  }
  // Descriptors are pure functions of their arguments and the workspace pointers are stable across
  // calls: memoise them (a forward issues ~60 encodes otherwise, ~0.1 ms of host time).
  struct Key { uintptr_t base; int dt, rank, swz; cuuint64_t gd[3], gs[2]; cuuint32_t bx[3]; };
  static thread_local std::vector<std::pair<Key, CUtensorMap>> cache;
  Key k = {};
  k.base = reinterpret_cast<uintptr_t>(base); k.dt = (int)dt; k.rank = (int)rank; k.swz = (int)swz;
  for (cuuint32_t i = 0; i < rank; ++i) { k.gd[i] = gdim[i]; k.bx[i] = box[i]; if (i + 1 < rank) k.gs[i] = gstride[i]; }
  for (const auto& e : cache)
    if (memcmp(&e.first, &k, sizeof(Key)) == 0) { *map = e.second; return CUDA_SUCCESS; }
  CUresult r = fn(map, dt, rank, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) {
    if (cache.size() >= 512) cache.clear();
    cache.emplace_back(k, *map);
  }
  return r;
}

CUresult make_tmap_bt128(CUtensorMap* map, const void* base, int B, int T, int box_rows) {
  cuuint64_t gdim[3] = {128, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstride[2] = {256, (cuuint64_t)T * 256};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box,
                      estr, CU_TENSOR_MAP_SWIZZLE_128B);
}

cudaError_t launch_attn_tc(const bf16* q, const bf16* k, const bf16* v, bf16* o, const int32_t* lengths,
                           int B, int T, int num_sms, cudaStream_t s, std::string* err) {
  if (B <= 0 || T <= 0) return cudaSuccess;
  CUtensorMap tq, tk, tv, to;
  CUresult r;
  if ((r = make_tmap_bt128(&to, o, B, T, BM)) != CUDA_SUCCESS ||
      (r = make_tmap_bt128(&tq, q, B, T, BM)) != CUDA_SUCCESS ||
      (r = make_tmap_bt128(&tk, k, B, T, BKV)) != CUDA_SUCCESS ||
      (r = make_tmap_bt128(&tv, v, B, T, BKV)) != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
    return cudaErrorInvalidValue;
  }
  const int npairs = (T + 2 * BM - 1) / (2 * BM);
  const long n_items_l = (long)B * npairs;
  if (n_items_l > 0x7fffffffL) return cudaErrorInvalidValue;
  const int n_pairs = (int)n_items_l;
  const long grid = n_pairs < num_sms ? n_pairs : num_sms;
  int n_full = n_pairs, n_items = n_pairs;
  {
    const int rem = n_pairs % (int)grid;        // pairs in the last, partially filled round
    if (n_pairs > grid && rem > 0 && 2 * rem <= grid) { n_full = n_pairs - rem; n_items = n_full + 2 * rem; }
  }
  static const int stagger = getenv("VADB_ATTN_STAGGER") ? atoi(getenv("VADB_ATTN_STAGGER")) : ATTN_DEFAULT_STAGGER;
  (void)stagger;   // captured by the generic lambda below
  static const bool want_trace = getenv("VADB_ATTN_TRACE") != nullptr;
  static const int variant = getenv("VADB_ATTN_VARIANT") ? atoi(getenv("VADB_ATTN_VARIANT")) : ATTN_DEFAULT_VARIANT;
  auto run = [&](auto kern_trace, auto kern) -> cudaError_t {
    if (want_trace) {
      cudaError_t e = cudaFuncSetAttribute(kern_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC);
      if (e != cudaSuccess) return e;
      long long* dtrace = nullptr;
      const int NTR = 2048;
      cudaMalloc(&dtrace, NTR * sizeof(long long));
      cudaMemsetAsync(dtrace, 0, NTR * sizeof(long long), s);
      kern_trace<<<(unsigned)grid, NTHREADS, SMEM_ALLOC, s>>>(tq, tk, tv, to, o, lengths, T, npairs, n_items, n_full, stagger, dtrace);
      std::vector<long long> ht(NTR);
      cudaMemcpyAsync(ht.data(), dtrace, NTR * sizeof(long long), cudaMemcpyDeviceToHost, s);
      cudaStreamSynchronize(s);
      cudaFree(dtrace);
      const long long t0 = ht[0];
      auto rel = [&](int i) { return ht[i] ? (long long)(ht[i] - t0) : -1LL; };
      fprintf(stderr, "[attn trace] variant %d: mma start=0 prologue issued=%lld\n", variant, rel(1));
      // CTA 0, whole kernel: clock64 and globaltimer at entry / exit -> SM clock and cycles per CTA
      if (ht[2043] > ht[2042])
        fprintf(stderr, "[attn trace] CTA0 lifetime: %lld cycles, %lld ns -> %.0f MHz; item 1 started at cycle %lld\n",
                ht[2041] - ht[2040], ht[2043] - ht[2042], 1e3 * (double)(ht[2041] - ht[2040]) / (double)(ht[2043] - ht[2042]),
                t0 - ht[2040]);
      const int nkv = (T + BKV - 1) / BKV;
      for (int j = 0; j < nkv && j < 16; ++j) {
        fprintf(stderr, "[attn trace] j=%d mma t0: waitP %lld->%lld issued %lld | t1: waitP %lld->%lld issued %lld\n", j,
                rel(64 + j * 16 + 0), rel(64 + j * 16 + 1), rel(64 + j * 16 + 3),
                rel(64 + j * 16 + 4), rel(64 + j * 16 + 5), rel(64 + j * 16 + 7));
        for (int t = 0; t < 2; ++t) {
          const int b0 = 512 + j * 32 + t * 16;
          fprintf(stderr, "[attn trace]   sm%d: waitS %lld->%lld ld %lld max %lld exp %lld waitO %lld corr+sts %lld arrive %lld\n",
                  t, rel(b0 + 0), rel(b0 + 1), rel(b0 + 2), rel(b0 + 3), rel(b0 + 4), rel(b0 + 5), rel(b0 + 6),
                  rel(b0 + 7));
        }
      }
      for (int t = 0; t < 2; ++t)
        fprintf(stderr, "[attn trace] epilogue t%d: waitO %lld->%lld done %lld\n", t, rel(32 + t * 4), rel(32 + t * 4 + 1),
                rel(32 + t * 4 + 2));
      return cudaGetLastError();
    }
    // the attribute is per function and per device; setting it on every launch costs ~1 us of host time
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC);
    if (e != cudaSuccess) return e;
    return launch_k(kern, (unsigned)grid, NTHREADS, SMEM_ALLOC, s, tq, tk, tv, to, o, lengths, T, npairs, n_items, n_full,
                    stagger, (long long*)nullptr);
  };
  switch (variant) {
#define VADB_ATTN_CASE(V) case V: return run(attn_tc_kernel<true, V>, attn_tc_kernel<false, V>);
    VADB_ATTN_VARIANTS(VADB_ATTN_CASE)
#undef VADB_ATTN_CASE
    default:
      if (err) *err = "unknown VADB_ATTN_VARIANT " + std::to_string(variant);
      return cudaErrorInvalidValue;
  }
}

}  // namespace vadb
