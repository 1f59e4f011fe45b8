// Fused forward of one MLP (all hidden layers + scalar head) per 128-row batch tile, hidden width 256.
// sm_100a only.  Replaces the per-layer forward launches of the tcgen05 path (first_fwd + hidden_fwd):
// the hidden activations of a tile never leave the SM between layers.
//
//   layer 0   D0[128,256] = X W0^T as 3xTF32 (Xhi Whi + Xlo Whi + Xhi Wlo), operands by TMA into the ring
//   epilogue  H1 = drop(relu(D0 + b0)), rounded to TF32, written BACK into the same 256 TMEM columns
//             (tcgen05.st) and, for the passes that train, to global memory for the backward pass
//   layer l   Dl = Hl Wl^T with A = Hl read straight from tensor memory (tcgen05.mma, A in TMEM) and only
//             the weight k-blocks streamed through the ring; Dl lands in the other 256 columns
//   ...       the two 256-column regions alternate, so any depth fits the 512 columns
//   last      H_L epilogue + the FP32 scalar head y = H_L w + b (Q, V passes) or the store of H_L for the
//             policy head kernel (actor pass)
//
// Per tile this removes the write + read of every intermediate H_l of the forward-only passes (V(s'), target
// Q) and the read of H_l of the training passes, and the activation operand traffic L2 -> SM of every layer
// but the first (128 KB of the 384 KB a 128x256x256 tile used to load).
//
// Roles: warp 0 lane 0 TMA producer (runs ahead across layers and tiles), warp 1 lane 0 MMA issuer, warps
// 2-17 epilogue (4 per TMEM lane quarter x 4 column groups), warp 18 stages the NEXT tile's biases, head
// weights and scalars in shared memory so that no epilogue waits on a dependent global load.
// Barriers: ring full / empty; tfull[2] (accumulator of forward event e = tile * L + l complete; alternating so
// a barrier is never lapped); achunk[c] (the four lane quarters have written columns [32c, +32) of H_{l+1} to
// tensor memory: the next layer's k-block c may be issued while the rest of the epilogue still runs); elast
// (every epilogue warp has read the last layer's accumulator); recfull / recempty[2] for the per-tile records.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tcgen05.cuh"
#include "umma_gemm.h"

namespace iql {

namespace {

constexpr int FT_M = 128, FT_N = 256, FT_K = 32, F_UMMA_K = 8;
constexpr int F_STAGE_A = FT_M * FT_K * 4;  // 16 KB (layer 0 only)
constexpr int F_STAGE_B = FT_N * FT_K * 4;  // 32 KB
constexpr int F_STAGE = F_STAGE_A + F_STAGE_B;
constexpr int F_STAGES = 3;
// CTA pair: the ring is cut into 16 KB granules; a layer-0 k-block takes two (own rows of X, own half of W0),
// a hidden-layer k-block one (own half of the W_l k-block): the whole of W_l (8 granules) can be in flight
constexpr int F_GRAN = 16 * 1024, F_NGRAN = 9, F_MAXBAR = 9;
constexpr int F_NA = 4, F_NB = F_NGRAN - F_NA;  // pair: granules of the layer-0 ring / of the hidden-layer weight ring
// 16 epilogue warps in two groups of 8 (2 per TMEM lane quarter = 2 column groups); group g owns the tiles with
// tile_it % 2 == g, so the first-layer epilogue of tile t+1 overlaps the last-layer epilogue of tile t
constexpr int F_EPI_WARPS = 16, F_EGROUPS = 2, F_GROUP_WARPS = F_EPI_WARPS / F_EGROUPS, F_CGROUPS = 2, F_CHUNKS = FT_N / 32;
constexpr int F_STG_FLOATS = 32 * 32;  // per-warp staging: one 32 x 32 chunk, rows of 128 B, 128-byte swizzled (TMA store box)
constexpr int F_RING = F_STAGES * F_STAGE;
constexpr int F_STG_BYTES = F_EPI_WARPS * F_STG_FLOATS * 4;
// per-tile record written by the prefetch warp one tile ahead: biases of every layer, head weights, scalars
constexpr int F_REC_FLOATS = FUSED_MAX_LAYERS * FT_N + FT_N + 64;
constexpr int F_REC_BYTES = F_REC_FLOATS * 4;
constexpr int F_YPART_BYTES = F_EGROUPS * 2 * F_CGROUPS * FT_M * 4;  // [group][2 buffers][column group][row]
constexpr int F_SMEM = F_RING + F_STG_BYTES + 2 * F_REC_BYTES + F_YPART_BYTES + 1024 /*align*/ + 512 /*barriers*/;
constexpr int F_THREADS = 608;  // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-17 epilogue, warp 18 per-tile prefetch
constexpr int FUSED_POL_MAX = 8;
constexpr int FUSED_TRACE_TILES = 8;
constexpr int FUSED_TRACE_ROLES = 5;  // TMA, MMA, epilogue (layer level), epilogue chunk level x 2
constexpr int FUSED_TRACE_WORDS = FUSED_TRACE_ROLES * FUSED_TRACE_TILES * FUSED_MAX_LAYERS * 4;

struct TileRec {  // the 64-float tail of a record
  float* C[FUSED_MAX_LAYERS];
  uint32_t* bits[FUSED_MAX_LAYERS];  // ReLU sign bits of the layer's output for the dgrad mask (null: not needed)
  float* head_out;
  unsigned long long seed, dstep;
  float head_b, drop_scale;
  uint32_t drop_thr;
  int head_ldc, store, fuse, member;
  int drop_layer[FUSED_MAX_LAYERS];
  int ldc[FUSED_MAX_LAYERS];
  // policy head (actor pass, act_dim <= FUSED_POL_MAX): z = H_L Wp^T + bp, FP32, in the last epilogue
  const float* pol_w;
  float* pol_out;
  int pol_a, pol_ldw, pol_ldc, pad_;
  float pol_b[FUSED_POL_MAX];
};
static_assert(sizeof(TileRec) <= 64 * 4, "TileRec must fit the record tail");

struct FusedParams {
  const GemmProb* probs[FUSED_MAX_LAYERS];   // hidden layer l: problem table of forward phase l (device)
  const CUtensorMap* maps[FUSED_MAX_LAYERS]; // l = 0: 4 per problem (Xhi, W0hi, Xlo, W0lo); l >= 1: 2 per problem, [1] = W_l
  const CUtensorMap* smaps[FUSED_MAX_LAYERS]; // output H_{l+1} of layer l, 1 per problem (TMA store, box 32 x 32)
  const GemmProb* probs_out;                 // output-layer problems (scalar heads of the first fuse_count)
  int L, nprob, tiles_m, units, fuse_count, nkb0;
  int fuse_policy;  // the problems >= fuse_count (actor pass) get their N = act_dim head fused as well
  int dbg;  // IQL_FUSED_DBG timing probes (results are wrong when set): 1 no sign bits, 2 no activation stores, 4 no staging wait, 8 no head / policy math, 16 reversed tile order
  int ks_last0;  // UMMA_K steps that carry data in the last k-block of layer 0 (observation widths <= 24: 3 of 4)
  uint32_t idesc;
  int wide;  // every CTA (pair) has at most ONE tile: both epilogue groups work on it, 2 chunks per warp and layer instead of 4
  unsigned trace_cta;  // IQL_FUSED_TRACE_CTA: which CTA records (default 0)
  long long* trace;  // IQL_FUSED_TRACE: clock64 stamps of CTA 0, [3 roles][FUSED_TRACE_TILES][FUSED_MAX_LAYERS][4]
};

__device__ __forceinline__ void trace_put(const FusedParams& fp, int role, uint32_t tile_it, int l, int slot) {
  if (fp.trace && blockIdx.x == fp.trace_cta && tile_it < (uint32_t)FUSED_TRACE_TILES)
    fp.trace[((role * FUSED_TRACE_TILES + tile_it) * FUSED_MAX_LAYERS + l) * 4 + slot] = clock64();
}

// Work order: table order (forward-only and critic passes first, the actor pass last).  Handing the actor tiles
// (policy head, dropout: the longest epilogues) out first measured slower (97 vs 92.5 us, same box, 64-member ensemble).
__device__ __forceinline__ int unit_prob(const FusedParams& fp, int u) { return ((fp.dbg & 16) ? fp.units - 1 - u : u) / fp.tiles_m; }

template <bool CTA2>
__global__ void __launch_bounds__(F_THREADS, 1) fused_fwd_kernel(FusedParams fp, StepCtx ctx) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  float* stg_all = reinterpret_cast<float*>(smem + F_RING);
  float* rec_s = reinterpret_cast<float*>(smem + F_RING + F_STG_BYTES);                        // [2][F_REC_FLOATS]
  float* ypart_s = reinterpret_cast<float*>(smem + F_RING + F_STG_BYTES + 2 * F_REC_BYTES);    // [2 groups][2][2][128]
  constexpr int BAR_OFF = F_RING + F_STG_BYTES + 2 * F_REC_BYTES + F_YPART_BYTES;
  const uint32_t bars = base + BAR_OFF;
  // full[9] empty[9] tfull[2 groups][2] elast achunk[8] recfull[2] recempty[2] | tmem slot
  const uint32_t full0 = bars, empty0 = bars + 8 * F_MAXBAR, tfull0 = bars + 16 * F_MAXBAR, elast = tfull0 + 32;
  const uint32_t achunk0 = elast + 8, recfull0 = achunk0 + 8 * F_CHUNKS, recempty0 = recfull0 + 16;
  const uint32_t tslot = recempty0 + 16;
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + BAR_OFF + (tslot - bars));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = fp.L;
  auto life_stamp = [&](int slot) {  // kernel entry / set-up done / predecessor done / exit, CTA 0 thread 0
    if (fp.trace && blockIdx.x == fp.trace_cta && threadIdx.x == 0)
      fp.trace[((4 * FUSED_TRACE_TILES + (FUSED_TRACE_TILES - 1)) * FUSED_MAX_LAYERS + (FUSED_MAX_LAYERS - 1)) * 4 + slot] = clock64();
  };
  life_stamp(0);
  stamp_begin(ctx.stamps, ST_FWD);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < F_MAXBAR; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int i = 0; i < 2 * F_EGROUPS; ++i) mbar_init(tfull0 + 8 * i, 1);
    // pair: the leader's barriers collect the epilogue warps of both CTAs
    mbar_init(elast, CTA2 ? 2 * F_GROUP_WARPS : F_GROUP_WARPS);  // the warps of the group that owns the tile
    for (int c = 0; c < F_CHUNKS; ++c) mbar_init(achunk0 + 8 * c, CTA2 ? 8 : 4);  // the lane quarters of chunk c
    for (int b = 0; b < 2; ++b) {
      mbar_init(recfull0 + 8 * b, 32);
      mbar_init(recempty0 + 8 * b, F_GROUP_WARPS);  // record buffer b belongs to epilogue group b
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 1) {
    if (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CTA2) cluster_sync();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tslot_ptr;
  life_stamp(1);
  pdl_wait();  // everything above overlapped the tail of the previous launch; the gathered rows and weights are read below
  life_stamp(2);
  const int nkb_h = FT_N / FT_K;  // 8 k-blocks of a hidden layer (K = 256)
  // pair: rank 0 leads (issues the MMAs, owns full / achunk / elast); work is dealt to pairs, 256 rows per tile
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const int worker = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tile_rows = CTA2 ? 2 * FT_M : FT_M;
  const int m_off = (int)rank * FT_M;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, bphase = 0, tile_it = 0;  // pair: phase = ring A, (stage, bphase) = ring B
      for (int u = worker; u < fp.units; u += n_workers, ++tile_it) {
        const int prob = unit_prob(fp, u);
        const int m0 = (u % fp.tiles_m) * tile_rows + m_off;
        {  // descriptors of the next tile's problem
          const int un = u + n_workers;
          if (un < fp.units) {
            const int pn = unit_prob(fp, un);
            if (pn != prob) {
              for (int i = 0; i < 4; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(fp.maps[0] + 4 * pn + i) : "memory");
              for (int l = 1; l < L; ++l) asm volatile("prefetch.tensormap [%0];" ::"l"(fp.maps[l] + 2 * pn + 1) : "memory");
              for (int l = 0; l < L; ++l) asm volatile("prefetch.tensormap [%0];" ::"l"(fp.smaps[l] + pn) : "memory");
            }
          }
        }
        if (CTA2) {
          const uint32_t full_lead = mapa_u32(full0, 0);
          // ring A (granules 0..3) holds ONE layer-0 k-block: Xhi, W0hi, Xlo, W0lo (each operand loaded once for the three
          // 3xTF32 passes); ring B (granules 4..8) streams the hidden-layer weight k-blocks.  Two rings instead of one
          // FIFO: the layer-0 operands of tile t+1 are requested as soon as the layer-0 MMAs of tile t are done instead
          // of queueing behind the hidden-layer k-blocks of tile t
          auto put = [&](uint32_t slot, uint32_t ph, const CUtensorMap* map, int c0, int c1) {  // one 16 KB granule of this CTA
            mbar_wait(empty0 + 8 * slot, ph ^ 1);
            if (rank == 0) mbar_expect_tx(full0 + 8 * slot, 2 * F_GRAN);  // the peer's granule lands on the same barrier
            tma_load_2d_cg2(base + slot * F_GRAN, map, full_lead + 8 * slot, c0, c1);
          };
          // layer-0 operands of tile `uu`; requested one tile ahead, in the middle of the previous tile's last
          // hidden layer: ring A is free as soon as that tile's layer-0 MMAs are done, long before ring B has room
          // for the rest of its k-blocks
          // ring A holds one k-block at a time, so only the FIRST layer-0 k-block of a tile is requested ahead; the
          // others (observation widths > 32) follow at the top of the tile itself -- requesting them ahead would wait
          // for layer-0 MMAs of a tile whose predecessor still lacks hidden-layer k-blocks this thread has yet to send
          auto put_l0 = [&](int uu, uint32_t ti, int kb_lo, int kb_hi) {
            const CUtensorMap* pm = fp.maps[0] + 4 * unit_prob(fp, uu);
            const int mm = (uu % fp.tiles_m) * tile_rows + m_off;
            if (kb_lo == 0) trace_put(fp, 0, ti, 0, 0);
            for (int kb = kb_lo; kb < kb_hi; ++kb) {
              put(0, phase, pm + 0, kb * FT_K, mm);                // own 128 rows of Xhi
              put(1, phase, pm + 1, kb * FT_K, (int)rank * 128);  // own half of W0hi
              put(2, phase, pm + 2, kb * FT_K, mm);                // Xlo
              put(3, phase, pm + 3, kb * FT_K, (int)rank * 128);  // W0lo
              phase ^= 1;
            }
            if (kb_hi == fp.nkb0) trace_put(fp, 0, ti, 0, 1);
          };
          if (tile_it == 0) put_l0(u, 0, 0, 1);
          if (fp.nkb0 > 1) put_l0(u, tile_it, 1, fp.nkb0);
          const int un = u + n_workers;
          bool ahead = un >= fp.units;  // nothing to request ahead after the last tile
          for (int l = 1; l < L; ++l) {
            trace_put(fp, 0, tile_it, l, 0);
            for (int kb = 0; kb < nkb_h; ++kb) {
              if (l == L - 1 && kb == F_NB && !ahead) {
                put_l0(un, tile_it + 1, 0, 1);
                ahead = true;
              }
              put(F_NA + stage, bphase, fp.maps[l] + 2 * prob + 1, kb * FT_K, (int)rank * 128);
              if (++stage == F_NB) { stage = 0; bphase ^= 1; }
            }
            trace_put(fp, 0, tile_it, l, 1);
          }
          if (!ahead) put_l0(un, tile_it + 1, 0, 1);  // L == 1
          continue;
        }
        for (int l = 0; l < L; ++l) {
          trace_put(fp, 0, tile_it, l, 0);
          if (l == 0) {
            const CUtensorMap* pm = fp.maps[0] + 4 * prob;  // passes: Xhi Whi, Xlo Whi, Xhi Wlo
            for (int sj = 0; sj < 3; ++sj) {
              const CUtensorMap* mapA = pm + (sj == 1 ? 2 : 0);
              const CUtensorMap* mapB = pm + (sj == 2 ? 3 : 1);
              for (int kb = 0; kb < fp.nkb0; ++kb) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t sa = base + stage * F_STAGE, sb = sa + F_STAGE_A, fb = full0 + 8 * stage;
                mbar_expect_tx(fb, F_STAGE);
                tma_load_2d(sa, mapA, fb, kb * FT_K, m0);
                tma_load_2d(sb, mapB, fb, kb * FT_K, 0);
                if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
              }
            }
          } else {
            const CUtensorMap* mapB = fp.maps[l] + 2 * prob + 1;
            for (int kb = 0; kb < nkb_h; ++kb) {
              mbar_wait(empty0 + 8 * stage, phase ^ 1);
              const uint32_t sb = base + stage * F_STAGE + F_STAGE_A, fb = full0 + 8 * stage;
              mbar_expect_tx(fb, F_STAGE_B);
              tma_load_2d(sb, mapB, fb, kb * FT_K, 0);
              if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
            }
          }
          trace_put(fp, 0, tile_it, l, 1);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0, bphase = 0, ev = 0, tile_it = 0, aev = 0;  // pair: phase = ring A, (stage, bphase) = ring B
      // the last layer's accumulator region (L-1)&1 is still being drained by the previous tile's last epilogue
      // when this tile reaches the first layer that writes the same region
      const int l_guard = (L - 1) & 1;
      for (int u = worker; u < fp.units; u += n_workers, ++tile_it) {
        for (int l = 0; l < L; ++l, ++ev) {
          trace_put(fp, 1, tile_it, l, 0);
          if (l == l_guard && tile_it > 0) {
            mbar_wait(elast, (tile_it - 1) & 1);
            tc_fence_after();
          }
          trace_put(fp, 1, tile_it, l, 1);
          const uint32_t tacc = tmem_base + (uint32_t)(l & 1) * 256u;
          const uint32_t ta = tmem_base + (uint32_t)((l - 1) & 1) * 256u;
          const int nkb = (l == 0) ? 3 * fp.nkb0 : nkb_h;
          // accumulator-full barriers: a pair per epilogue group, alternating over the group's own events
          const uint32_t tf_bar = (tile_it & 1) * 2 + (((tile_it >> 1) * (uint32_t)L + (uint32_t)l) & 1);
          if (CTA2) {
            if (l == 0) {
              for (int kb = 0; kb < fp.nkb0; ++kb) {
                const int nks = (kb == fp.nkb0 - 1) ? fp.ks_last0 : FT_K / F_UMMA_K;  // skip all-zero K steps
#pragma unroll
                for (int sj = 0; sj < 3; ++sj) {  // Xhi W0hi, Xlo W0hi, Xhi W0lo
                  const uint32_t ga = (sj == 1) ? 2u : 0u, gb = (sj == 2) ? 3u : 1u;
                  if (sj == 0) mbar_wait(full0, phase);
                  mbar_wait(full0 + 8 * (sj + 1), phase);
                  tc_fence_after();
                  if (kb == 0 && sj == 0) trace_put(fp, 1, tile_it, l, 2);
                  const uint64_t adesc0 = make_desc(base + ga * F_GRAN, 1, 1024 >> 4, 2);
                  const uint64_t bdesc0 = make_desc(base + gb * F_GRAN, 1, 1024 >> 4, 2);
#pragma unroll
                  for (int ks = 0; ks < FT_K / F_UMMA_K; ++ks)
                    if (ks < nks)
                      umma_tf32_cg2(tacc, adesc0 + (uint64_t)(ks * 2), bdesc0 + (uint64_t)(ks * 2), fp.idesc, (kb | sj | ks) != 0);
                }
                for (uint32_t g = 0; g < (uint32_t)F_NA; ++g) umma_commit_cg2(empty0 + 8 * g);
                phase ^= 1;
              }
            } else {
              for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(achunk0 + 8 * kb, aev & 1);  // both CTAs have columns [32 kb, +32) of H_l in tensor memory
                mbar_wait(full0 + 8 * (F_NA + stage), bphase);
                tc_fence_after();
                if (kb == 0) trace_put(fp, 1, tile_it, l, 2);
                const uint64_t bdesc0 = make_desc(base + (F_NA + stage) * F_GRAN, 1, 1024 >> 4, 2);
#pragma unroll
                for (int ks = 0; ks < FT_K / F_UMMA_K; ++ks)
                  umma_tf32_ts_cg2(tacc, ta + (uint32_t)(kb * FT_K + ks * F_UMMA_K), bdesc0 + (uint64_t)(ks * 2), fp.idesc,
                                   (kb | ks) != 0);
                umma_commit_cg2(empty0 + 8 * (F_NA + stage));
                if (++stage == F_NB) { stage = 0; bphase ^= 1; }
              }
            }
            if (l > 0) ++aev;
            umma_commit_cg2(tfull0 + 8 * tf_bar);
            trace_put(fp, 1, tile_it, l, 3);
            continue;
          }
          for (int kb = 0; kb < nkb; ++kb) {
            if (l > 0) {  // columns [32 kb, +32) of H_l are in tensor memory (all four lane quarters)
              mbar_wait(achunk0 + 8 * kb, aev & 1);
              tc_fence_after();
            }
            mbar_wait(full0 + 8 * stage, phase);
            tc_fence_after();
            if (kb == 0) trace_put(fp, 1, tile_it, l, 2);
            const uint32_t sa = base + stage * F_STAGE, sb = sa + F_STAGE_A;
            const uint64_t bdesc0 = make_desc(sb, 1, 1024 >> 4, 2);
            if (l == 0) {
              const uint64_t adesc0 = make_desc(sa, 1, 1024 >> 4, 2);
              const int nks = ((kb % fp.nkb0) == fp.nkb0 - 1) ? fp.ks_last0 : FT_K / F_UMMA_K;  // skip all-zero K steps
#pragma unroll
              for (int ks = 0; ks < FT_K / F_UMMA_K; ++ks)
                if (ks < nks) umma_tf32(tacc, adesc0 + (uint64_t)(ks * 2), bdesc0 + (uint64_t)(ks * 2), fp.idesc, (kb | ks) != 0);
            } else {
#pragma unroll
              for (int ks = 0; ks < FT_K / F_UMMA_K; ++ks)
                umma_tf32_ts(tacc, ta + (uint32_t)(kb * FT_K + ks * F_UMMA_K), bdesc0 + (uint64_t)(ks * 2), fp.idesc,
                             (kb | ks) != 0);
            }
            umma_commit(empty0 + 8 * stage);
            if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
          }
          if (l > 0) ++aev;  // one completion of every achunk barrier per non-last epilogue
          umma_commit(tfull0 + 8 * tf_bar);
          trace_put(fp, 1, tile_it, l, 3);
        }
      }
    }
  } else if (warp == 18) {
    // ===================== per-tile prefetch (one tile ahead of the epilogue) =====================
    uint32_t tile_it = 0;
    for (int u = worker; u < fp.units; u += n_workers, ++tile_it) {
      const int prob = unit_prob(fp, u);
      const uint32_t b = tile_it & 1;
      float* rec = rec_s + b * F_REC_FLOATS;
      const bool fuse = prob < fp.fuse_count;
      // lane l < L: problem of layer l; lane L: the head problem
      GemmProb g;
      memset(&g, 0, sizeof(g));
      const bool pol = !fuse && fp.fuse_policy;
      if (lane < L) g = fp.probs[lane][prob];
      else if (lane == L && (fuse || pol)) g = fp.probs_out[prob];
      mbar_wait(recempty0 + 8 * b, ((tile_it >> 1) & 1) ^ 1);  // the epilogue is done with the tile that used this buffer
      for (int l = 0; l < L; ++l) {
        const float* bias = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)g.bias, l));
        const float4* b4 = reinterpret_cast<const float4*>(bias);
        reinterpret_cast<float4*>(rec + l * FT_N)[lane] = __ldg(b4 + lane);
        reinterpret_cast<float4*>(rec + l * FT_N)[lane + 32] = __ldg(b4 + lane + 32);
      }
      const float* hw = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)g.B, L));
      if (fuse) {
        reinterpret_cast<float4*>(rec + FUSED_MAX_LAYERS * FT_N)[lane] = __ldg(reinterpret_cast<const float4*>(hw) + lane);
        reinterpret_cast<float4*>(rec + FUSED_MAX_LAYERS * FT_N)[lane + 32] = __ldg(reinterpret_cast<const float4*>(hw) + lane + 32);
      }
      TileRec* tr = reinterpret_cast<TileRec*>(rec + FUSED_MAX_LAYERS * FT_N + FT_N);
      if (lane < L) {
        tr->C[lane] = g.C;
        tr->bits[lane] = g.bits;
        tr->ldc[lane] = g.ldc;
        tr->drop_layer[lane] = g.drop_layer;
      }
      if (lane == L - 1) {
        tr->store = g.no_store == 0;  // forward-only passes keep nothing
        tr->member = g.member;
        tr->fuse = fuse ? 1 : 0;
        const MemberScalars* sc = ctx.scalars + g.member;
        tr->drop_thr = sc->drop_threshold;
        tr->drop_scale = sc->drop_scale;
        tr->seed = sc->seed;
        tr->dstep = (unsigned long long)(ctx.counters[g.member].actor_step + ctx.k);
      }
      if (lane == L && fuse) {
        tr->head_out = g.C;
        tr->head_ldc = g.ldc;
        tr->head_b = __ldg(g.bias);
      }
      {
        const float* pb = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)g.bias, L));
        const int pa = __shfl_sync(0xffffffffu, g.N, L);
        if (lane < FUSED_POL_MAX) tr->pol_b[lane] = (pol && lane < pa) ? __ldg(pb + lane) : 0.f;
        if (pol) {  // pull the act_dim x 256 policy weights towards this SM: the epilogue reads them with uniform loads
          const float* pw = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)g.B, L));
          const int ldw = __shfl_sync(0xffffffffu, g.ldb, L);
          for (int idx = lane; idx < pa * (FT_N / 32); idx += 32)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pw + (int64_t)(idx / (FT_N / 32)) * ldw + (idx % (FT_N / 32)) * 32) : "memory");
        }
        if (lane == L) {
          tr->pol_a = pol ? g.N : 0;
          tr->pol_w = g.B;
          tr->pol_ldw = g.ldb;
          tr->pol_out = g.C;
          tr->pol_ldc = g.ldc;
        }
      }
      mbar_arrive(recfull0 + 8 * b);  // every lane releases its own writes
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int eg = (warp - 2) / F_GROUP_WARPS;         // epilogue group: tiles with tile_it % 2 == eg
    const int ch = ((warp - 2) % F_GROUP_WARPS) >> 2;  // column group within the group: chunks ch, ch + 2, ch + 4, ch + 6
    // wide: a launch with one tile per CTA (single learners, few members) would leave the other group idle; instead all
    // 16 warps take the tile, as four column groups of two chunks each (the head partials then add up in four terms)
    const bool wide = fp.wide != 0;
    const int cg = wide ? eg * F_CGROUPS + ch : ch;        // column group of this warp
    const int n_cg = wide ? F_EGROUPS * F_CGROUPS : F_CGROUPS;
    const int tr_id = wide ? (int)threadIdx.x - 64 : (int)threadIdx.x - 64 - eg * (F_GROUP_WARPS * 32);  // thread index within the tile's warps
    const int bar_id = wide ? 3 : 1 + eg, bar_n = wide ? F_EPI_WARPS * 32 : F_GROUP_WARPS * 32;
    const bool tracer = ((warp - 2) % F_GROUP_WARPS) == 0 && lane == 0 && (!wide || eg == 0);
    float* stg = stg_all + (warp - 2) * F_STG_FLOATS;
    uint32_t ev = 0, tile_it = 0;
    const uint32_t elast_lead = CTA2 ? mapa_u32(elast, 0) : elast;
    const uint32_t achunk_lead = CTA2 ? mapa_u32(achunk0, 0) : achunk0;
    for (int u = worker; u < fp.units; u += n_workers, ++tile_it) {
      if (!wide && (int)(tile_it & 1) != eg) continue;  // the other group's tile
      const int prob = unit_prob(fp, u);
      const int m0 = (u % fp.tiles_m) * tile_rows + m_off;
      const uint32_t b = tile_it & 1;
      const float* rec = rec_s + b * F_REC_FLOATS;
      ev = (tile_it >> 1) * (uint32_t)L;  // this group's event counter at the tile's first layer
      mbar_wait(recfull0 + 8 * b, (tile_it >> 1) & 1);
      const TileRec* tr = reinterpret_cast<const TileRec*>(rec + FUSED_MAX_LAYERS * FT_N + FT_N);
      const bool store = tr->store != 0;
      const int row = m0 + q * 32 + lane;  // the accumulator row this lane holds
      for (int l = 0; l < L; ++l, ++ev) {
        const bool last = (l == L - 1);
        const bool fuse = last && tr->fuse;
        const int drop_layer = tr->drop_layer[l];
        const uint32_t drop_thr = (drop_layer >= 0) ? tr->drop_thr : 0u;
        const bool drop = drop_thr != 0u;
        const float drop_scale = tr->drop_scale;
        const float* bs = rec + l * FT_N;
        const float* ws = rec + FUSED_MAX_LAYERS * FT_N;
        float* ypart = wide ? ypart_s : ypart_s + (eg * 2 + ((tile_it >> 1) & 1)) * (F_CGROUPS * FT_M);
        if (tracer) trace_put(fp, 2, tile_it, l, 0);
        mbar_wait(tfull0 + 8 * ((wide ? 0 : eg * 2) + (ev & 1)), (ev >> 1) & 1);  // wide: tile 0 belongs to group 0's barriers
        tc_fence_after();
        if (tracer) trace_put(fp, 2, tile_it, l, 1);
        const uint32_t region = tmem_base + (uint32_t)(l & 1) * 256u + ((uint32_t)(q * 32) << 16);
        float yacc = 0.f;
        const int pol_a = last ? tr->pol_a : 0;  // > 0: actor pass, policy head fused
        float pacc[FUSED_POL_MAX];
#pragma unroll
        for (int a = 0; a < FUSED_POL_MAX; ++a) pacc[a] = 0.f;
#pragma unroll 1
        for (int c = cg; c < F_CHUNKS; c += n_cg) {
          uint32_t r[32];
          tmem_ld32(region + (uint32_t)(c * 32), r);
          tmem_ld_wait();
          if (tracer && c == cg) trace_put(fp, 3, tile_it, l, 0);
          if (tracer && c == cg + n_cg) trace_put(fp, 4, tile_it, l, 0);
          if (last && !wide && c + F_CGROUPS >= F_CHUNKS) {  // this warp's last read of the tile: the region may be overwritten (wide: no next tile)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CTA2) mbar_arrive_remote(elast_lead);
              else mbar_arrive(elast);
            }
          }
          const float4* bc = reinterpret_cast<const float4*>(bs + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = bc[j];
            r[4 * j + 0] = __float_as_uint(fmaxf(__uint_as_float(r[4 * j + 0]) + b4.x, 0.f));
            r[4 * j + 1] = __float_as_uint(fmaxf(__uint_as_float(r[4 * j + 1]) + b4.y, 0.f));
            r[4 * j + 2] = __float_as_uint(fmaxf(__uint_as_float(r[4 * j + 2]) + b4.z, 0.f));
            r[4 * j + 3] = __float_as_uint(fmaxf(__uint_as_float(r[4 * j + 3]) + b4.w, 0.f));
          }
          if (drop) {
            if (ctx.dropout_masks) {
              const uint8_t* mkb = ctx.dropout_masks +
                                   ((((int64_t)tr->member * ctx.K + ctx.k) * ctx.L + drop_layer) * ctx.B + row) * (int64_t)ctx.H + c * 32;
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const uint32_t mk4 = *reinterpret_cast<const uint32_t*>(mkb + 4 * j4);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float v = __uint_as_float(r[4 * j4 + t]);
                  r[4 * j4 + t] = __float_as_uint(((mk4 >> (8 * t)) & 0xFFu) ? v * drop_scale : 0.f);
                }
              }
            } else {
              const uint64_t drop_seed = tr->seed, dstep = tr->dstep;
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const uint32_t quad = (uint32_t)(((int64_t)row * FT_N + c * 32 + 4 * j4) >> 2);
                const Philox4 ph = philox_dropout_quad(drop_seed, dstep, (uint32_t)drop_layer, quad);
                r[4 * j4 + 0] = __float_as_uint((ph.x >= drop_thr) ? __uint_as_float(r[4 * j4 + 0]) * drop_scale : 0.f);
                r[4 * j4 + 1] = __float_as_uint((ph.y >= drop_thr) ? __uint_as_float(r[4 * j4 + 1]) * drop_scale : 0.f);
                r[4 * j4 + 2] = __float_as_uint((ph.z >= drop_thr) ? __uint_as_float(r[4 * j4 + 2]) * drop_scale : 0.f);
                r[4 * j4 + 3] = __float_as_uint((ph.w >= drop_thr) ? __uint_as_float(r[4 * j4 + 3]) * drop_scale : 0.f);
              }
            }
          }
          if (fuse && !(fp.dbg & 8)) {  // FP32 head on the unrounded activations
            const float4* wc = reinterpret_cast<const float4*>(ws + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 w4 = wc[j];
              yacc = fmaf(__uint_as_float(r[4 * j + 0]), w4.x, yacc);
              yacc = fmaf(__uint_as_float(r[4 * j + 1]), w4.y, yacc);
              yacc = fmaf(__uint_as_float(r[4 * j + 2]), w4.z, yacc);
              yacc = fmaf(__uint_as_float(r[4 * j + 3]), w4.w, yacc);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(round_tf32(__uint_as_float(r[j])));
          if (tracer && c == cg) trace_put(fp, 3, tile_it, l, 1);
          if (!last) {  // columns [32 c, +32) of the next layer's operand A, in place; the MMAs of k-block c may go
            tmem_st32(region + (uint32_t)(c * 32), r);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CTA2) mbar_arrive_remote(achunk_lead + 8 * c);
              else mbar_arrive(achunk0 + 8 * c);
            }
          }
          if (tracer && c == cg) trace_put(fp, 3, tile_it, l, 2);
          if (store && !(fp.dbg & 2)) {
            // every lane holds one full 128-byte row of the chunk: park it (16-byte units XOR-swizzled by the row,
            // the layout a SWIZZLE_128B tensor map expects) and let one TMA store write the 32 x 32 box -- no
            // transposition, and no store instruction of this warp waits for the memory system
            if (lane == 0 && !(fp.dbg & 4)) bulk_wait_read0();  // the previous box has left the staging tile
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(&stg[lane * 32 + 4 * (j ^ (lane & 7))]) =
                  make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                              __uint_as_float(r[4 * j + 3]));
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(fp.smaps[l] + prob, smem_u32(stg), c * 32, m0 + q * 32);
              bulk_commit();
            }
            if (pol_a > 0 && !(fp.dbg & 8)) {
              // Policy head z += H_L[:, 32c..32c+31] Wp[:, 32c..]^T on the parked (TF32-exact) chunk with warp-level
              // tensor-core MMAs: m16n8k8, A fragments straight from the swizzled staging tile (conflict-free), B = the
              // act_dim <= 8 weight rows split hi + lo on the fly (two accumulating passes: FP32-accurate weights).
              // 16 MMAs + 32 LDS per chunk instead of 256 shuffles + 256 FMAs (1.4-1.8 us -> ~0.3 us per chunk).
              const int g = lane >> 2, t = lane & 3;
              const float* wrow = tr->pol_w + (int64_t)g * tr->pol_ldw + c * 32 + t;
              const bool wvalid = g < pol_a;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const float w0 = wvalid ? __ldg(wrow + 8 * ks) : 0.f, w1 = wvalid ? __ldg(wrow + 8 * ks + 4) : 0.f;
                const float w0h = round_tf32(w0), w1h = round_tf32(w1);
                const uint32_t bh0 = __float_as_uint(w0h), bh1 = __float_as_uint(w1h);
                const uint32_t bl0 = __float_as_uint(round_tf32(w0 - w0h)), bl1 = __float_as_uint(round_tf32(w1 - w1h));
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                  const int r0 = 16 * mt + g, r1 = r0 + 8;  // r0 & 7 == r1 & 7 == g
                  const int u0 = 4 * ((2 * ks) ^ g) + t, u1 = 4 * ((2 * ks + 1) ^ g) + t;
                  const uint32_t a0 = __float_as_uint(stg[r0 * 32 + u0]), a1 = __float_as_uint(stg[r1 * 32 + u0]);
                  const uint32_t a2 = __float_as_uint(stg[r0 * 32 + u1]), a3 = __float_as_uint(stg[r1 * 32 + u1]);
                  float* d = pacc + 4 * mt;
                  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bh0), "r"(bh1));
                  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bl0), "r"(bl1));
                }
              }
            }
          }
          if (tracer && c == cg) trace_put(fp, 3, tile_it, l, 3);
          uint32_t* const bits = tr->bits[l];
          if (store && bits != nullptr && !(fp.dbg & 1)) {  // 1 bit per element for the dgrad mask (of the ROUNDED value, like the store)
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) word |= (__uint_as_float(r[j]) > 0.f ? 1u : 0u) << j;
            bits[(int64_t)row * F_CHUNKS + c] = word;
          }
          if (tracer && c == cg) trace_put(fp, 4, tile_it, l, 1);
        }
        if (tracer) trace_put(fp, 2, tile_it, l, 2);
        if (fuse) {
          ypart[cg * FT_M + q * 32 + lane] = yacc;
          asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_n) : "memory");
          if (tr_id < FT_M) {
            float y = ypart[tr_id] + ypart[FT_M + tr_id];
            if (wide) y += ypart[2 * FT_M + tr_id] + ypart[3 * FT_M + tr_id];
            tr->head_out[(int64_t)(m0 + tr_id) * tr->head_ldc] = y + tr->head_b;
          }
        }
        if (pol_a > 0) {
          // the four column groups park their partial sums in their (idle) staging tiles; the warps of group 0
          // add them up in a fixed order and write z
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
          // accumulator fragment of m-tile mt: rows 16 mt + (lane >> 2) (+ 8), actions 2 (lane & 3) (+ 1)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              stg[(16 * mt + (lane >> 2) + 8 * (i >> 1)) * FUSED_POL_MAX + 2 * (lane & 3) + (i & 1)] = pacc[4 * mt + i];
          asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_n) : "memory");
          if (cg == 0) {
            float* zrow = tr->pol_out + (int64_t)row * tr->pol_ldc;
#pragma unroll
            for (int a = 0; a < FUSED_POL_MAX; ++a) {
              if (a < pol_a) {
                const float* p0 = stg + lane * FUSED_POL_MAX + a;  // this warp is group 0 of its quarter; group g4 is 4 g4 tiles further
                float z = 0.f;
#pragma unroll
                for (int g4 = 0; g4 < F_EGROUPS * F_CGROUPS; ++g4)
                  if (g4 < n_cg) z += p0[g4 * 4 * F_STG_FLOATS];
                zrow[a] = z + tr->pol_b[a];
              }
            }
          }
          asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_n) : "memory");  // the staging tiles go back to the activation stores
        }
        if (tracer) trace_put(fp, 2, tile_it, l, 3);
      }
      __syncwarp();
      if (lane == 0 && !wide) mbar_arrive(recempty0 + 8 * b);  // this warp no longer reads the tile's record (wide: never reused)
    }
    bulk_wait0();  // all activation stores of this thread have completed
  }
  tc_fence_before();
  __syncwarp();
  if (CTA2) cluster_sync();
  else __syncthreads();
  if (warp == 1) {
    if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  life_stamp(3);
  stamp_end(ctx.stamps, ST_FWD);
}

}  // namespace

static long long* fused_trace_buffer() {  // allocated once when IQL_FUSED_TRACE is set
  static long long* buf = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    if (dbg_getenv("IQL_FUSED_TRACE") && cudaMalloc(&buf, sizeof(long long) * FUSED_TRACE_WORDS) == cudaSuccess)
      cudaMemset(buf, 0, sizeof(long long) * FUSED_TRACE_WORDS);
    else
      buf = nullptr;
  }
  return buf;
}

bool fused_fwd_supported(int batch, int hidden, int n_hidden, int k0) {
  fused_trace_buffer();  // allocated here (state binding), never inside a stream capture
  return dbg_getenv("IQL_B200_NO_FUSED_FWD") == nullptr && umma_phase_supported(0, batch, hidden) && hidden == FT_N &&
         n_hidden >= 1 && n_hidden <= FUSED_MAX_LAYERS && k0 >= 1;
}

// CTA pairs: one pair per 256 batch rows of a problem, each CTA staging half of every weight k-block
bool fused_fwd_pair(int batch) {
  return dbg_getenv("IQL_B200_NO_FUSED_PAIR") == nullptr && dbg_getenv("IQL_B200_NO_CTA2") == nullptr && batch % (2 * FT_M) == 0;
}

bool fused_fwd_policy_head(int act_dim) { return act_dim >= 1 && act_dim <= FUSED_POL_MAX && dbg_getenv("IQL_B200_NO_FUSED_POLICY") == nullptr; }

void launch_fused_fwd(const FusedFwdArgs& a, const StepCtx& ctx, cudaStream_t st) {
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(fused_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM);
    cudaFuncSetAttribute(fused_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM);
  }
  const bool pair = a.pair != 0;
  const int tile_rows = pair ? 2 * FT_M : FT_M;
  FusedParams fp;
  memset(&fp, 0, sizeof(fp));
  for (int l = 0; l < a.L; ++l) {
    fp.probs[l] = a.probs[l];
    fp.maps[l] = (const CUtensorMap*)a.maps[l];
    fp.smaps[l] = (const CUtensorMap*)a.store_maps[l];
  }
  fp.probs_out = a.probs_out;
  fp.L = a.L;
  fp.nprob = a.nprob;
  fp.tiles_m = (a.batch + tile_rows - 1) / tile_rows;
  fp.units = a.nprob * fp.tiles_m;
  fp.fuse_count = a.probs_out ? a.fuse_count : 0;
  fp.fuse_policy = (a.probs_out && a.fuse_policy) ? 1 : 0;
  {
    static const int dbg = dbg_getenv("IQL_FUSED_DBG") ? atoi(dbg_getenv("IQL_FUSED_DBG")) : 0;
    fp.dbg = dbg;
  }
  fp.nkb0 = (a.k0_max + FT_K - 1) / FT_K;
  fp.ks_last0 = (a.k0_max - (fp.nkb0 - 1) * FT_K + F_UMMA_K - 1) / F_UMMA_K;
  // c = F32, a = b = TF32, both K-major, N = 256, M = 128 (256 for a CTA pair)
  fp.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(FT_N >> 3) << 17) | ((uint32_t)(tile_rows >> 4) << 24);
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  fp.trace = fused_trace_buffer();
  {
    static const unsigned tc = dbg_getenv("IQL_FUSED_TRACE_CTA") ? (unsigned)atoi(dbg_getenv("IQL_FUSED_TRACE_CTA")) : 0u;
    fp.trace_cta = tc;
  }
  static const bool no_wide = dbg_getenv("IQL_B200_NO_WIDE_EPILOGUE") != nullptr;
  if (!pair) {
    const int grid = fp.units < n_sm ? fp.units : n_sm;
    fp.wide = (fp.units <= n_sm && !no_wide) ? 1 : 0;
    launch_pdl(fused_fwd_kernel<false>, dim3(grid), dim3(F_THREADS), F_SMEM, st, 1, fp, ctx);
    return;
  }
  const int workers = fp.units < n_sm / 2 ? fp.units : n_sm / 2;
  fp.wide = (fp.units <= n_sm / 2 && !no_wide) ? 1 : 0;
  launch_pdl(fused_fwd_kernel<true>, dim3(2 * workers), dim3(F_THREADS), F_SMEM, st, 2, fp, ctx);
}

// measurement hook (tools/fused_trace.py): copies the clock64 stamps of the last fused_fwd launch to the host
extern "C" int iql_debug_fused_trace(long long* out, int32_t max_words) {
  long long* buf = fused_trace_buffer();
  if (!buf || !out || max_words < FUSED_TRACE_WORDS) return -1;
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  if (cudaMemcpy(out, buf, sizeof(long long) * FUSED_TRACE_WORDS, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  return FUSED_TRACE_WORDS;
}

}  // namespace iql
