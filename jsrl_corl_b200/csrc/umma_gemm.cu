// tcgen05 (UMMA) TF32 grouped GEMM for the hidden-layer GEMMs of the IQL step
// (IQL_MATH_TF32_TCGEN05).  sm_100a only.
//
//   mode 0  NT  C[M,N] = A[M,K] * B[N,K]^T   forward  H' = relu(H W^T + b)      A,B K-major
//   mode 1  NN  C[M,N] = A[M,K] * B[K,N]     dgrad    G' = (G W) * [H > 0]      A K-major, B MN-major
//   mode 2  TN  C[M,N] = A[K,M]^T * B[K,N]   wgrad    dW = G^T H                A,B MN-major
//
// Persistent kernel, one CTA per SM, looping over 128 x tile_n output tiles (tile_n <= 256).  The FP32
// accumulator of a tile is one M=128 `tcgen05.mma.cta_group::1.kind::tf32` block of 256 TMEM columns; the 512
// columns hold TWO such accumulators, so the epilogue of tile i overlaps the TMA + MMA main loop of tile
// i+1.  Warp 0 lane 0 issues TMA (cp.async.bulk.tensor, 128-byte swizzle) into a 3-stage ring and runs
// ahead across tile boundaries; warp 1 lane 0 issues the MMAs, `tcgen05.commit`s every stage to the
// ring's empty barrier and the last one of a tile to the accumulator-full barrier; warps 2-9 (two per
// TMEM lane quarter, splitting the columns) drain the accumulator with tcgen05.ld, release it through
// the accumulator-empty barrier, transpose through a private staging tile and apply the fused epilogue
// (bias + ReLU + dropout, ReLU mask, or plain store; optionally the FP32 output heads) with coalesced
// 16-byte global accesses.
//
// CTA pairs (CTA2 = true): when M is a multiple of 256 and the N tile is 256 wide the kernel is launched as
// 2-CTA clusters and one `tcgen05.mma.cta_group::2` covers a 256 x 256 tile: each CTA of the pair stages its
// own 128 rows of A and HALF of the B tile (the tensor cores of both SMs read both halves), so a tile costs
// 32 KB instead of 48 KB of operand delivery per k-block and SM.  Both CTAs issue TMA (`.cta_group::2`,
// completing on the leader's full barrier), only the leader (cluster rank 0) issues the MMAs and commits
// with `.multicast::cluster` to the ring/accumulator barriers of both CTAs; each CTA drains its own 128 TMEM
// lanes and the epilogue warps of both release the accumulator on the leader's barrier.
//
// Shared-memory / descriptor conventions (cute/atom/mma_traits_sm100.hpp):
//   K-major  : rows of 128 B, SWIZZLE_128B (16 B atoms), SBO = 1024 B between
//              8-row groups, K advance of one UMMA (8 tf32) = +32 B.
//   MN-major : tf32 requires SWIZZLE_128B_BASE32B; slabs of 32 MN-elements
//              ([32 k-rows][128 B]), LBO = slab stride, SBO = 512 B between
//              4-row groups, K advance of one UMMA (8 k-rows) = +1024 B.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"
#include "umma_gemm.h"

namespace iql {

constexpr int TILE_M = 128, TILE_N = 256 /* maximum; the N tile is a launch parameter */, TILE_K = 32, UMMA_K = 8, N_STAGES = 3 /* ring size in MAX-size stages */,
              MAX_STAGES = 8;  // narrow N tiles use more, smaller stages: same bytes in flight
constexpr int STAGE_A_BYTES = TILE_M * TILE_K * 4;  // 16 KB
constexpr int STAGE_B_BYTES = TILE_N * TILE_K * 4;  // 32 KB
constexpr int STAGE_BYTES = STAGE_A_BYTES + STAGE_B_BYTES;
constexpr int N_EPI_WARPS = 16;                                       // 4 column groups x 4 TMEM lane quarters
constexpr int N_CGROUPS = N_EPI_WARPS / 4;
constexpr int STG_FLOATS = 32 * 36;                                   // per-warp transpose tile
constexpr int AUX_FLOATS = 2 * N_CGROUPS * TILE_M;                    // head partials [2 parities][4 groups][128]; also
                                                                      // the bias-gradient sums [4 quarters][256]
constexpr int AUX_BYTES = (AUX_FLOATS > 4 * TILE_N ? AUX_FLOATS : 4 * TILE_N) * 4;
constexpr int PEER_CSUM_SLOTS = 3;                                    // CTA pair: ring of bias-gradient partials sent by the peer
constexpr int EPI_SMEM_BYTES = N_EPI_WARPS * STG_FLOATS * 4 + AUX_BYTES + PEER_CSUM_SLOTS * TILE_N * 4;
constexpr int SMEM_BYTES = N_STAGES * STAGE_BYTES + EPI_SMEM_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
constexpr int N_THREADS = 576;  // warp0 TMA, warp1 MMA + TMEM alloc, warps 2-17 epilogue (4 per TMEM lane quarter)

struct UmmaParams {
  // per-operand descriptor fields, host-computed (16-byte units unless noted)
  uint32_t a_lbo, a_sbo, a_layout, a_kstep;
  uint32_t b_lbo, b_sbo, b_layout, b_kstep;
  uint32_t idesc;
  int a_mn, b_mn;  // operand majors (0 K-major, 1 MN-major)
  int tile_n;      // UMMA N of this launch: multiple of 32, <= 256 (B tile = tile_n * 128 B per stage)
  int tile_m;      // rows of one work tile: 128, or 256 for a CTA pair (each CTA owns 128 of them)
  int tiles_m, tiles_n, total_tiles;  // tile grid per problem and over the whole launch
  // Work units handed to CTAs round-robin.  Default: one unit = one tile.  prob_major (dgrad with fused bias
  // gradient): one unit = all M tiles of one (problem, N tile), so that the column sums of the produced
  // gradient are completed inside one CTA in a fixed order -- no atomics, no zero-fill.
  int prob_major, units, tiles_per_unit;
  // 3xTF32 input layer: the K loop runs n_split passes over K, pass j taking A from map a_sel[j] and B from
  // map b_sel[j] of the problem's maps_per_prob tensor maps: (Xhi,Whi), (Xlo,Whi), (Xhi,Wlo).
  int n_split, maps_per_prob, a_sel[3], b_sel[3];
  int fuse_count;  // FUSE_OUT: problems [0, fuse_count) have a scalar head evaluated in the epilogue
  int n_stages, stage_bytes;  // TMA ring geometry for this tile_n
  int k_max;  // largest K of the phase: every problem runs ceil(k_max / 32) k-blocks (TMA zero-fills beyond its own K)
  // IQL_UMMA_DBG, profiling experiments only (tools/umma_probe.sh; the results of the step are WRONG):
  // 1 no global stores, 2 operands of 4 problems only (L2 hits), 4 no epilogue, 8 no MMAs, 16 no TMA loads,
  // 32 no tensor-map prefetch, 64 no loads of the A operand
  int dbg;
  int stamp_slot;  // step-timeline slot of this launch (ST_WGRAD / ST_DGRAD / ST_FWGRAD / ST_POLHEAD)
};

__device__ __forceinline__ void decode_tile(const UmmaParams& up, int unit, int j, int& prob, int& m0, int& n0) {
  if (up.prob_major) {
    prob = unit / up.tiles_n;
    n0 = (unit - prob * up.tiles_n) * up.tile_n;
    m0 = j * up.tile_m;
  } else {
    const int per = up.tiles_m * up.tiles_n;
    prob = unit / per;
    const int rem = unit - prob * per;
    m0 = (rem / up.tiles_n) * up.tile_m;
    n0 = (rem % up.tiles_n) * up.tile_n;
  }
}

// FUSE_OUT (forward, last hidden layer): the scalar heads y = H_L w^T + b (Q, V: N = 1) of the first
// `up.fuse_count` problems are evaluated in the epilogue, in FP32, on the FP32 accumulators -- Q and V never see
// TF32 rounding and H_L makes no extra trip through HBM (for the forward-only passes it is never stored).
// `probs_out` is the problem table of the output-layer phase; the policy head (N = act_dim) of the remaining
// problems is left to the FP32 output-layer kernel.
// ROWEPI (backward phases, EPI_NONE / EPI_DRELU): the epilogue keeps the tcgen05.ld layout (one accumulator row per
// lane), masks with one 32-bit word of ReLU sign bits per (row, 32-column chunk) instead of 32 FP32 activations,
// parks the chunk 128-byte-swizzled in shared memory and lets ONE TMA store per chunk write it (`cmaps`, one
// output map per problem): no transposition, no store instruction waits for the memory system.
template <int EPI, bool FUSE_OUT, bool CTA2, bool ROWEPI>
__global__ void __launch_bounds__(N_THREADS, 1)
umma_gemm_kernel(const GemmProb* __restrict__ probs, const CUtensorMap* __restrict__ maps,
                 const GemmProb* __restrict__ probs_out, const CUtensorMap* __restrict__ cmaps, UmmaParams up, StepCtx ctx) {
  extern __shared__ uint8_t smem_raw[];
  stamp_begin(ctx.stamps, up.stamp_slot);
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle atoms need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  constexpr int RING = N_STAGES * STAGE_BYTES;
  float* epi_smem = reinterpret_cast<float*>(smem + RING);
  // barriers: full[8], empty[8], tmem_full[2], tmem_empty[2], tmem slot, peer-csum full[3]
  const uint32_t bars = base + RING + EPI_SMEM_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * MAX_STAGES, tfull0 = bars + 16 * MAX_STAGES, tempty0 = tfull0 + 16;
  const uint32_t tslot = tempty0 + 16;
  const uint32_t csfull0 = tslot + 8;
  // CTA pair: rank 0 leads (issues the MMAs, owns the full / accumulator-empty barriers); work is dealt to pairs
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const int worker = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_off = (int)rank * TILE_M;  // this CTA's rows inside the pair's 256-row tile
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + RING + EPI_SMEM_BYTES + 16 * MAX_STAGES + 32);
  const int n_stages = up.n_stages;                      // RING / stage_bytes, capped at MAX_STAGES
  const uint32_t stage_bytes = (uint32_t)up.stage_bytes;  // 16 KB of A + tile_n * 128 B of B, 1024-byte multiple

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_n = up.tile_n;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, CTA2 ? 2 * N_EPI_WARPS : N_EPI_WARPS);  // pair: the epilogue warps of both CTAs
    }
    for (int b = 0; b < PEER_CSUM_SLOTS; ++b) mbar_init(csfull0 + 8 * b, up.tile_n);  // one arrival per column of the N tile
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 1) {  // whole warp: allocate all 512 TMEM columns (1 CTA per SM) = two accumulators
    if (CTA2) {     // the same warp of BOTH CTAs of the pair allocates
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CTA2) cluster_sync();  // the peer's barriers and TMEM must exist before anything is sent to them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tslot_ptr;
  pdl_wait();  // set-up above overlaps the previous launch; operands, masks and outputs are touched below

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const int b_rows = CTA2 ? (tile_n >> 1) : tile_n;  // pair: this CTA stages half of the B tile
      const bool skip_a = (up.dbg & 64) != 0;  // probe: what the phase would cost if its A operand needed no load
      const uint32_t tx_bytes = (uint32_t)(CTA2 ? 2 : 1) * (uint32_t)((skip_a ? 0 : STAGE_A_BYTES) + b_rows * TILE_K * 4);
      const uint32_t full0_lead = CTA2 ? mapa_u32(full0, 0) : full0;
      for (int u = worker; u < up.units; u += n_workers)
      for (int j = 0; j < up.tiles_per_unit; ++j) {
        int prob, m0, n0;
        decode_tile(up, u, j, prob, m0, n0);
        m0 += m_off;
        n0 += (int)rank * b_rows;
        const CUtensorMap* pmaps = maps + up.maps_per_prob * ((up.dbg & 2) ? (prob & 3) : prob);
        const int nkb = (up.k_max + TILE_K - 1) / TILE_K;  // TMA zero-fills the K tail
        const int num_kb = nkb * up.n_split;
        if (up.dbg & 16) continue;
        if (!(up.dbg & 32)) {  // descriptors of the NEXT tile: fetched while this tile's k-blocks stream in
          int un = u, jn = j + 1;
          if (jn == up.tiles_per_unit) { jn = 0; un += n_workers; }
          if (un < up.units) {
            int prob2, m2, n2;
            decode_tile(up, un, jn, prob2, m2, n2);
            if (prob2 != prob) {
              const CUtensorMap* nm = maps + up.maps_per_prob * ((up.dbg & 2) ? (prob2 & 3) : prob2);
              for (int i = 0; i < up.maps_per_prob; ++i)
                asm volatile("prefetch.tensormap [%0];" ::"l"(nm + i) : "memory");
            }
          }
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          const int sj = kb / nkb;
          const CUtensorMap* mapA = pmaps + up.a_sel[sj];
          const CUtensorMap* mapB = pmaps + up.b_sel[sj];
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t sa = base + stage * stage_bytes, sb = sa + STAGE_A_BYTES;
          const int k0 = (kb - sj * nkb) * TILE_K;
          if (CTA2) {  // the leader arms its barrier for the loads of both CTAs; the peer's bytes land on it too
            const uint32_t fb = full0_lead + 8 * stage;
            if (rank == 0) mbar_expect_tx(full0 + 8 * stage, tx_bytes);
            if (skip_a) {}
            else if (up.a_mn) tma_load_3d_cg2(sa, mapA, fb, 0, k0, m0 >> 5);
            else tma_load_2d_cg2(sa, mapA, fb, k0, m0);
            if (up.b_mn) tma_load_3d_cg2(sb, mapB, fb, 0, k0, n0 >> 5);
            else tma_load_2d_cg2(sb, mapB, fb, k0, n0);
          } else {
            const uint32_t fb = full0 + 8 * stage;
            mbar_expect_tx(fb, tx_bytes);
            if (skip_a) {}
            else if (up.a_mn) tma_load_3d(sa, mapA, fb, 0, k0, m0 >> 5);
            else tma_load_2d(sa, mapA, fb, k0, m0);
            if (up.b_mn) tma_load_3d(sb, mapB, fb, 0, k0, n0 >> 5);
            else tma_load_2d(sb, mapB, fb, k0, n0);
          }
          if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int u = worker; u < up.units; u += n_workers)
      for (int j = 0; j < up.tiles_per_unit; ++j, ++it) {
        int prob, m0, n0;
        decode_tile(up, u, j, prob, m0, n0);
        const int num_kb = ((up.k_max + TILE_K - 1) / TILE_K) * up.n_split;
        const uint32_t buf = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(tempty0 + 8 * buf, acc_phase ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!(up.dbg & 16)) mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = base + stage * stage_bytes, sb = sa + STAGE_A_BYTES;
          const uint64_t adesc0 = make_desc(sa, up.a_lbo, up.a_sbo, up.a_layout);
          const uint64_t bdesc0 = make_desc(sb, up.b_lbo, up.b_sbo, up.b_layout);
#pragma unroll
          for (int ks = 0; ks < TILE_K / UMMA_K; ++ks) {
            if (up.dbg & 8) break;
            if (CTA2) umma_tf32_cg2(tacc, adesc0 + (uint64_t)(ks * up.a_kstep), bdesc0 + (uint64_t)(ks * up.b_kstep), up.idesc,
                                    (kb | ks) != 0);
            else umma_tf32(tacc, adesc0 + (uint64_t)(ks * up.a_kstep), bdesc0 + (uint64_t)(ks * up.b_kstep), up.idesc,
                           (kb | ks) != 0);
          }
          // frees the smem slot (of both CTAs of a pair) when these MMAs have read it
          if (CTA2) umma_commit_cg2(empty0 + 8 * stage);
          else umma_commit(empty0 + 8 * stage);
          if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
        }
        if (CTA2) umma_commit_cg2(tfull0 + 8 * buf);  // accumulator complete (both halves)
        else umma_commit(tfull0 + 8 * buf);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // warp w may only touch TMEM lanes [32*(w%4), +32); two warps share a quarter and split the columns.
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;  // column group (0..3) handled by this warp: chunks ch, ch+4
    float* stg = epi_smem + (warp - 2) * STG_FLOATS;
    float* ypart_all = epi_smem + N_EPI_WARPS * STG_FLOATS;  // [2 tile parities][4 column groups][128 rows]
    const int lr = lane >> 3;        // row within a group of 4
    const int lc = (lane & 7) * 4;   // first of this lane's 4 columns
    uint32_t it = 0, ucount = 0;
    const uint32_t tempty0_lead = CTA2 ? mapa_u32(tempty0, 0) : tempty0;
    float* peer_csum = epi_smem + N_EPI_WARPS * STG_FLOATS + AUX_BYTES / 4;  // [PEER_CSUM_SLOTS][TILE_N], written by the peer
    auto acc_release = [&](uint32_t buf) {  // this warp has read its part of the accumulator
      if (CTA2) mbar_arrive_cluster(tempty0_lead + 8 * buf);
      else mbar_arrive(tempty0 + 8 * buf);
    };
    for (int u = worker; u < up.units; u += n_workers, ++ucount)
    for (int j = 0; j < up.tiles_per_unit; ++j, ++it) {
      int prob, m0, n0;
      decode_tile(up, u, j, prob, m0, n0);
      m0 += m_off;
      const GemmProb p = probs[prob];
      const uint32_t buf = it & 1, acc_phase = (it >> 1) & 1;
      // per-member scalars are read ONCE per tile into registers: inside the store loop the compiler would
      // have to reload them after every global store (possible aliasing), exposing a DRAM latency per row
      const MemberScalars* sc = ctx.scalars + p.member;
      const uint32_t drop_thr = (p.drop_layer >= 0) ? sc->drop_threshold : 0u;
      const float drop_scale = sc->drop_scale;
      const uint64_t drop_seed = sc->seed;
      float dscale = 1.0f;
      bool drop = false;
      uint64_t dstep = 0;
      if (EPI == EPI_DRELU) dscale = (drop_thr != 0u) ? drop_scale : 1.0f;
      if (EPI == EPI_RELU) {
        drop = drop_thr != 0u;
        if (drop) dstep = (uint64_t)(ctx.counters[p.member].actor_step + ctx.k);
      }
      GemmProb po;
      const bool fuse = FUSE_OUT && prob < up.fuse_count;
      float* ypart = ypart_all + (it & 1) * (N_CGROUPS * TILE_M);
      if (fuse) po = probs_out[prob];
      const bool skip_store = fuse && p.no_store;
      float yacc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) yacc[i] = 0.f;
      const int row_base = m0 + q * 32;
      const int n_chunks = tile_n >> 5;
      // fused bias gradient (dgrad): column sums of the produced G over all rows of the problem
      const bool do_csum = (EPI == EPI_DRELU) && up.prob_major && p.dbias != nullptr;
      float* csum_s = ypart_all;  // [4 quarters][TILE_N]; the head-partial area is unused by dgrad launches
      // Operand prefetch for the epilogue, one chunk ahead: the ReLU mask (dgrad), the bias (forward) and the
      // head weights do not depend on the accumulator, so the first chunk's loads are in flight while this
      // warp still waits for the MMAs, and chunk c+2's loads are issued before chunk c is processed.
      float4 mk[8];
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), b4_n = b4;
      float4 w4 = b4, w4_n = b4;
      auto prefetch = [&](int c, float4& bd, float4& wd) {
        const int col = n0 + c * 32 + lc;
        if (EPI == EPI_RELU) bd = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        if (EPI == EPI_LINEAR) {  // output layer: N = 1 or act_dim, guarded
          bd.x = col < p.N ? p.bias[col] : 0.f;
          bd.y = col + 1 < p.N ? p.bias[col + 1] : 0.f;
          bd.z = col + 2 < p.N ? p.bias[col + 2] : 0.f;
          bd.w = col + 3 < p.N ? p.bias[col + 3] : 0.f;
        }
        if (fuse) wd = __ldg(reinterpret_cast<const float4*>(po.B + col));
      };
      if (ROWEPI) {
        // ---- row layout: lane = accumulator row; sign-bit mask; TMA store ----
        float* const stgr = epi_smem + (warp - 2) * 1024;  // 4 KB per warp, 1024-byte aligned (swizzle atoms)
        const int row = row_base + lane;
        const int nw = p.ldmask >> 5;  // words per row of the sign-bit matrix
        uint32_t wbits[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
        if (EPI == EPI_DRELU) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int c = ch + i * N_CGROUPS;
            if (c < n_chunks) wbits[i] = __ldg(p.bits + (int64_t)row * nw + (n0 >> 5) + c);
          }
        }
        mbar_wait(tfull0 + 8 * buf, acc_phase);
        tc_fence_after();
        if (up.dbg & 4) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) acc_release(buf);
          continue;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c = ch + i * N_CGROUPS;
          if (c >= n_chunks) break;
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + c * 32), r);
          tmem_ld_wait();
          if (c + N_CGROUPS >= n_chunks) {  // last TMEM read of this warp for this tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) acc_release(buf);
          }
          if (EPI == EPI_DRELU) {
            const uint32_t w = wbits[i];
#pragma unroll
            for (int j2 = 0; j2 < 32; ++j2)
              r[j2] = ((w >> j2) & 1u) ? __float_as_uint(round_tf32(__uint_as_float(r[j2]) * dscale)) : 0u;
          }
          if (lane == 0) bulk_wait_read0();  // the previous chunk has left the staging tile
          __syncwarp();
#pragma unroll
          for (int j2 = 0; j2 < 8; ++j2)
            *reinterpret_cast<float4*>(&stgr[lane * 32 + 4 * (j2 ^ (lane & 7))]) =
                make_float4(__uint_as_float(r[4 * j2]), __uint_as_float(r[4 * j2 + 1]), __uint_as_float(r[4 * j2 + 2]),
                            __uint_as_float(r[4 * j2 + 3]));
          fence_async_smem();
          __syncwarp();
          if (!(up.dbg & 1) && lane == 0) {
            tma_store_2d(cmaps + prob, smem_u32(stgr), n0 + c * 32, row_base);
            bulk_commit();
          }
          if (do_csum) {  // column `lane` of the parked chunk, summed over its 32 rows
            float cs = 0.f;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) cs += stgr[rr * 32 + 4 * ((lane >> 2) ^ (rr & 7)) + (lane & 3)];
            float* dst = &csum_s[q * TILE_N + c * 32 + lane];
            *dst = (j == 0) ? cs : *dst + cs;
          }
        }
      } else {
      if (ch < n_chunks) prefetch(ch, b4_n, w4_n);
      mbar_wait(tfull0 + 8 * buf, acc_phase);
      tc_fence_after();
      if (up.dbg & 4) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) acc_release(buf);
        continue;
      }
#pragma unroll 1
      for (int c = ch; c < n_chunks; c += N_CGROUPS) {
        const int col = n0 + c * 32 + lc;
        // full 16-byte accesses when this lane's 4 columns exist and rows are 16-byte aligned
        const bool vec = (col + 3 < p.N) && ((p.ldc & 3) == 0);
        b4 = b4_n;
        w4 = w4_n;
        if (c + N_CGROUPS < n_chunks) prefetch(c + N_CGROUPS, b4_n, w4_n);
        if (EPI == EPI_DRELU) {  // ReLU mask of this chunk: in flight during the TMEM read and the transpose
#pragma unroll
          for (int i = 0; i < 8; ++i)
            mk[i] = __ldg(reinterpret_cast<const float4*>(p.mask + (int64_t)(row_base + i * 4 + lr) * p.ldmask + col));
        }
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + c * 32), r);
        tmem_ld_wait();
        if (c + N_CGROUPS >= n_chunks) {  // last TMEM read of this warp for this tile: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) acc_release(buf);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(&stg[lane * 36 + 4 * j]) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        float4 cs4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float* const cbase = p.C + (int64_t)(row_base + lr) * p.ldc + col;  // row i*4+lr is 4*i*ldc floats further
        const int cstep = 4 * p.ldc;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + lr;
          const int row = row_base + rl;
          float4 v = *reinterpret_cast<const float4*>(&stg[rl * 36 + lc]);
          if (EPI == EPI_RELU) {
            v.x = fmaxf(v.x + b4.x, 0.f); v.y = fmaxf(v.y + b4.y, 0.f);
            v.z = fmaxf(v.z + b4.z, 0.f); v.w = fmaxf(v.w + b4.w, 0.f);
            if (drop) {
              if (ctx.dropout_masks) {
                const uint8_t* mkb = ctx.dropout_masks +
                                     ((((int64_t)p.member * ctx.K + ctx.k) * ctx.L + p.drop_layer) * ctx.B + row) * (int64_t)ctx.H + col;
                v.x = mkb[0] ? v.x * drop_scale : 0.f; v.y = mkb[1] ? v.y * drop_scale : 0.f;
                v.z = mkb[2] ? v.z * drop_scale : 0.f; v.w = mkb[3] ? v.w * drop_scale : 0.f;
              } else {
                const uint32_t quad = (uint32_t)(((int64_t)row * p.N + col) >> 2);
                const Philox4 ph = philox_dropout_quad(drop_seed, dstep, (uint32_t)p.drop_layer, quad);
                v.x = (ph.x >= drop_thr) ? v.x * drop_scale : 0.f;
                v.y = (ph.y >= drop_thr) ? v.y * drop_scale : 0.f;
                v.z = (ph.z >= drop_thr) ? v.z * drop_scale : 0.f;
                v.w = (ph.w >= drop_thr) ? v.w * drop_scale : 0.f;
              }
            }
          } else if (EPI == EPI_LINEAR) {
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          } else if (EPI == EPI_DRELU) {
            v.x = mk[i].x > 0.f ? v.x * dscale : 0.f; v.y = mk[i].y > 0.f ? v.y * dscale : 0.f;
            v.z = mk[i].z > 0.f ? v.z * dscale : 0.f; v.w = mk[i].w > 0.f ? v.w * dscale : 0.f;
            cs4.x += v.x; cs4.y += v.y; cs4.z += v.z; cs4.w += v.w;
          }
          if (fuse) {
            yacc[i] = fmaf(v.x, w4.x, fmaf(v.y, w4.y, fmaf(v.z, w4.z, fmaf(v.w, w4.w, yacc[i]))));
            if (skip_store) continue;
          }
          if (EPI == EPI_RELU || EPI == EPI_DRELU) {  // these outputs are operands of later tcgen05 GEMMs
            v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
          }
          float* crow = cbase + i * cstep;
          if (up.dbg & 1) continue;
          if (vec) {
            *reinterpret_cast<float4*>(crow) = v;
          } else {
            if (col < p.N) crow[0] = v.x;
            if (col + 1 < p.N) crow[1] = v.y;
            if (col + 2 < p.N) crow[2] = v.z;
            if (col + 3 < p.N) crow[3] = v.w;
          }
        }
        if (do_csum) {  // 32-row partial of this warp: add the 4 row groups, accumulate across the M tiles
          cs4.x += __shfl_xor_sync(0xffffffffu, cs4.x, 8);  cs4.y += __shfl_xor_sync(0xffffffffu, cs4.y, 8);
          cs4.z += __shfl_xor_sync(0xffffffffu, cs4.z, 8);  cs4.w += __shfl_xor_sync(0xffffffffu, cs4.w, 8);
          cs4.x += __shfl_xor_sync(0xffffffffu, cs4.x, 16); cs4.y += __shfl_xor_sync(0xffffffffu, cs4.y, 16);
          cs4.z += __shfl_xor_sync(0xffffffffu, cs4.z, 16); cs4.w += __shfl_xor_sync(0xffffffffu, cs4.w, 16);
          if (lane < 8) {
            float4* dst = reinterpret_cast<float4*>(&csum_s[q * TILE_N + c * 32 + lc]);
            float4 acc = (j == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : *dst;
            acc.x += cs4.x; acc.y += cs4.y; acc.z += cs4.z; acc.w += cs4.w;
            *dst = acc;
          }
        }
        __syncwarp();
      }
      }  // !ROWEPI
      if (do_csum && j == up.tiles_per_unit - 1) {  // all rows of the problem seen: combine the quarters
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const int tc = threadIdx.x - 64;  // 0..511; the first tile_n threads own one column each
        if (tc < tile_n) {
          float own = ((csum_s[tc] + csum_s[TILE_N + tc]) + csum_s[2 * TILE_N + tc]) + csum_s[3 * TILE_N + tc];
          if (CTA2) {
            // the peer sends the sums of its 128 rows into the leader's ring slot and arrives (release.cluster)
            // on the slot's barrier.  Slot reuse is safe with 3 slots: the peer can only reach unit i+3 after the
            // MMAs of i+3, which wait for every leader epilogue warp to have started unit i+1, i.e. finished i.
            const uint32_t slot = ucount % PEER_CSUM_SLOTS, par = (ucount / PEER_CSUM_SLOTS) & 1u;
            if (rank != 0) {
              st_cluster_f32(mapa_u32(smem_u32(&peer_csum[slot * TILE_N + tc]), 0), own);
              mbar_arrive_cluster(mapa_u32(csfull0 + 8 * slot, 0));
            } else {
              mbar_wait_cluster(csfull0 + 8 * slot, par);
              own += peer_csum[slot * TILE_N + tc];
            }
          }
          if (rank == 0 && n0 + tc < p.N) p.dbias[n0 + tc] = own;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");  // csum_s is reused by the next unit
      }
      if (n_chunks <= ch) {  // this warp had no chunk (tile_n < 128): still release the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) acc_release(buf);
      }
      if (FUSE_OUT) {
        // scalar heads: reduce the 8 lanes that share a row, park the per-column-group partials, combine them.
        // Every epilogue warp takes part in the barrier (tiles without a fused head just pass through).
        if (fuse) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float tsum = yacc[i];
            tsum += __shfl_xor_sync(0xffffffffu, tsum, 1);
            tsum += __shfl_xor_sync(0xffffffffu, tsum, 2);
            tsum += __shfl_xor_sync(0xffffffffu, tsum, 4);
            if ((lane & 7) == 0) ypart[ch * TILE_M + q * 32 + i * 4 + lr] = tsum;
          }
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");  // the 16 epilogue warps
        const int tr = threadIdx.x - 64;                // 0..511
        if (fuse && tr < TILE_M)
          po.C[(int64_t)(m0 + tr) * po.ldc] =
              (((ypart[tr] + ypart[TILE_M + tr]) + ypart[2 * TILE_M + tr]) + ypart[3 * TILE_M + tr]) + po.bias[0];
      }
    }
  }
  if (ROWEPI && warp >= 2) bulk_wait0();  // this thread's TMA stores have completed
  tc_fence_before();
  __syncwarp();
  if (CTA2) cluster_sync();  // neither CTA may exit (or free TMEM) while the pair's MMAs / remote arrivals are in flight
  else __syncthreads();
  if (warp == 1) {
    if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  stamp_end(ctx.stamps, up.stamp_slot);
}

// bias gradient of the wgrad phases: dbias[m] = sum_k A[k][m]  (A = G, [K = batch rows][M] row-major).
// 1024 threads = 32 columns x 32 row groups: coalesced 128-byte row segments, 16 / 8 independent loads in flight per
// thread, fixed summation order.  (256 threads with 4 loads in flight walked 512 rows per thread at batch 4096:
// 75 us of pure load latency per launch on the stress shape.)
__global__ void __launch_bounds__(1024) colsum_kernel(const GemmProb* __restrict__ probs) {
  __shared__ float part[32][33];
  const GemmProb p = probs[blockIdx.y];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int m = blockIdx.x * 32 + c;
  float s = 0.f;
  if (m < p.M && p.dbias != nullptr) {
    const float* col = p.A + m;
    int k = g;
    for (; k + 15 * 32 < p.K; k += 16 * 32) {
      float a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = col[(int64_t)(k + 32 * j) * p.lda];
      s += (((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]))) +
           (((a[8] + a[9]) + (a[10] + a[11])) + ((a[12] + a[13]) + (a[14] + a[15])));
    }
    for (; k + 7 * 32 < p.K; k += 8 * 32) {
      float a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = col[(int64_t)(k + 32 * j) * p.lda];
      s += ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    }
    for (; k < p.K; k += 32) s += col[(int64_t)k * p.lda];
  }
  part[g][c] = s;
  __syncthreads();
  if (g == 0 && m < p.M && p.dbias != nullptr) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = (part[4 * j][c] + part[4 * j + 1][c]) + (part[4 * j + 2][c] + part[4 * j + 3][c]);
    p.dbias[m] = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static uint32_t env_u32(const char* name, uint32_t dflt) {
  const char* s = dbg_getenv(name);
  return s ? (uint32_t)strtoul(s, nullptr, 0) : dflt;
}

bool umma_phase_supported(int mode, int batch, int hidden) {
  (void)mode;  // every tcgen05 phase has M in {batch, hidden}: both must be multiples of the 256-row tile
  static int disabled = -1;
  if (disabled < 0) disabled = dbg_getenv("IQL_B200_NO_UMMA") ? 1 : 0;
  return !disabled && batch >= 128 && batch % 128 == 0 && hidden >= 256 && hidden % 256 == 0;
}

// K-major operand [rows][K] (ld floats): 2-D map {K, rows}, box {32, box_rows}, SWIZZLE_128B.
// Rows / K beyond the extents are zero-filled by TMA (K tails, output layers with N = 1 or act_dim).
static int encode_kmajor(CUtensorMap* out, const float* ptr, int rows, int K, int ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 1;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {TILE_K, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

// MN-major operand [K rows][MN] (ld floats): 3-D map {32, K, ceil(MN/32)}, box {32, 32, slabs},
// SWIZZLE_128B_ATOM_32B.  MN < 32 shrinks the inner extent so the missing columns are zero-filled.
static int encode_mnmajor(CUtensorMap* out, const float* ptr, int mn, int K, int ld, int slabs) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 1;
  cuuint64_t gdim[3] = {(cuuint64_t)(mn < 32 ? mn : 32), (cuuint64_t)K, (cuuint64_t)((mn + 31) / 32)};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, 128};
  cuuint32_t box[3] = {32, TILE_K, (cuuint32_t)slabs};
  cuuint32_t es[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = (CUtensorMapSwizzle)env_u32("IQL_UMMA_MN_TMA_SWIZZLE", (uint32_t)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

int umma_encode_store_map(void* h_map_out, const float* ptr, int rows, int cols, int ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 1;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t es[2] = {1, 1};
  return enc((CUtensorMap*)h_map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstr, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

int umma_tile_n(int maxN) {
  const int t = (maxN + 31) / 32 * 32;
  return t > TILE_N ? TILE_N : t;
}

// CTA pairs: every M of the phase a multiple of 256 (two 128-row halves), a full 256-wide N tile; the fused
// bias gradient of the dgrad epilogue is exchanged inside the pair, which covers one 256-row tile per problem.
bool umma_dgrad_writes_dbias(int batch);

bool umma_cta2_ok(int maxM, int maxN) {
  return dbg_getenv("IQL_B200_NO_CTA2") == nullptr && maxM > 0 && (maxM % (2 * TILE_M)) == 0 && maxN >= TILE_N &&
         (maxN % TILE_N) == 0;
}

// Where pairs pay (measured, profiles/r01_cta_pair.md): long K loops (hidden >= 512 forward, batch >= 512 weight
// gradients) where operand delivery per MMA matters, and dgrad launches with too few problems to give every SM a
// tile of its own.  At 2x256 / batch 256 with 64 members the phases are bound by HBM traffic, which pairs do not
// change, and the coupling of the two epilogues costs ~8 %.
bool umma_cta2(int mode, int nprob, int maxM, int maxN, int maxK) {
  if (!umma_cta2_ok(maxM, maxN)) return false;
  if (dbg_getenv("IQL_B200_FORCE_CTA2")) return true;  // tests: run every eligible phase on pairs
  if (mode == 1) return umma_dgrad_writes_dbias(maxM) && nprob * (maxN / TILE_N) <= 74;
  return maxK >= 512;
}

int umma_encode_maps(int mode, const GemmProb* h_probs, int nprob, int tile_n, void* h_maps_out, bool cta2) {
  CUtensorMap* maps = (CUtensorMap*)h_maps_out;
  if (cta2) tile_n >>= 1;  // each CTA of a pair stages half of the B tile
  for (int i = 0; i < nprob; ++i) {
    const GemmProb& p = h_probs[i];
    int rc;
    if (mode == 2) rc = encode_mnmajor(&maps[2 * i], p.A, p.M, p.K, p.lda, TILE_M / 32);  // 4 slabs = 128 rows
    else rc = encode_kmajor(&maps[2 * i], p.A, p.M, p.K, p.lda, TILE_M);
    if (rc) return rc;
    if (mode == 0) rc = encode_kmajor(&maps[2 * i + 1], p.B, p.N, p.K, p.ldb, tile_n);
    else rc = encode_mnmajor(&maps[2 * i + 1], p.B, p.N, p.K, p.ldb, tile_n / 32);
    if (rc) return rc;
  }
  return 0;
}

static UmmaParams make_params(int mode, int tile_n, bool cta2) {
  UmmaParams u;
  u.tile_n = tile_n;
  u.tile_m = cta2 ? 2 * TILE_M : TILE_M;
  u.n_split = 1;
  u.fuse_count = 0;
  u.maps_per_prob = 2;
  u.a_sel[0] = u.a_sel[1] = u.a_sel[2] = 0;
  u.b_sel[0] = u.b_sel[1] = u.b_sel[2] = 1;
  const int a_mn = (mode == 2), b_mn = (mode != 0);
  u.a_mn = a_mn;
  u.b_mn = b_mn;
  // K-major SW128: LBO unused (1), SBO 1024 B, layout 2, K step 32 B, second half = 128 rows * 128 B
  // MN-major SW128/32B: LBO = slab stride 4096 B, SBO 512 B, layout 1, K step 1024 B, second half = 4 slabs
  u.a_lbo = a_mn ? env_u32("IQL_UMMA_MN_LBO", 4096 >> 4) : 1;
  u.a_sbo = a_mn ? env_u32("IQL_UMMA_MN_SBO", 512 >> 4) : (1024 >> 4);
  u.a_layout = a_mn ? env_u32("IQL_UMMA_MN_LAYOUT", 1) : 2;
  u.a_kstep = a_mn ? env_u32("IQL_UMMA_MN_KSTEP", 1024 >> 4) : (32 >> 4);
  u.b_lbo = b_mn ? env_u32("IQL_UMMA_MN_LBO", 4096 >> 4) : 1;
  u.b_sbo = b_mn ? env_u32("IQL_UMMA_MN_SBO", 512 >> 4) : (1024 >> 4);
  u.b_layout = b_mn ? env_u32("IQL_UMMA_MN_LAYOUT", 1) : 2;
  u.b_kstep = b_mn ? env_u32("IQL_UMMA_MN_KSTEP", 1024 >> 4) : (32 >> 4);
  // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=TF32 [7,10), b=TF32 [10,13),
  // a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
  u.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
            ((uint32_t)(tile_n >> 3) << 17) | ((uint32_t)(u.tile_m >> 4) << 24);
  return u;
}

// 3xTF32 forward of the input layer: 4 maps per problem = [Xhi, Whi, Xlo, Wlo] (all K-major)
int umma_encode_maps_split(const GemmProb* h_hi, const GemmProb* h_lo, int nprob, int tile_n, void* h_maps_out, bool cta2) {
  CUtensorMap* maps = (CUtensorMap*)h_maps_out;
  if (cta2) tile_n >>= 1;
  for (int i = 0; i < nprob; ++i) {
    const GemmProb& a = h_hi[i];
    const GemmProb& b = h_lo[i];
    if (encode_kmajor(&maps[4 * i + 0], a.A, a.M, a.K, a.lda, TILE_M)) return 1;
    if (encode_kmajor(&maps[4 * i + 1], a.B, a.N, a.K, a.ldb, tile_n)) return 1;
    if (encode_kmajor(&maps[4 * i + 2], b.A, b.M, b.K, b.lda, TILE_M)) return 1;
    if (encode_kmajor(&maps[4 * i + 3], b.B, b.N, b.K, b.ldb, tile_n)) return 1;
  }
  return 0;
}

bool umma_dgrad_writes_dbias(int batch) { return (batch + TILE_M - 1) / TILE_M <= 2; }

bool umma_can_fuse_out(int act_dim) { (void)act_dim; return dbg_getenv("IQL_B200_NO_FUSE_OUT") == nullptr; }

template <int EPI, bool FUSE_OUT, bool ROWEPI>
static void launch_variant(bool cta2, int workers, const GemmProb* probs, const CUtensorMap* maps, const GemmProb* probs_out,
                           const CUtensorMap* cmaps, const UmmaParams& up, const StepCtx& ctx, cudaStream_t st) {
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(umma_gemm_kernel<EPI, FUSE_OUT, false, ROWEPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(umma_gemm_kernel<EPI, FUSE_OUT, true, ROWEPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  }
  if (!cta2)
    launch_pdl(umma_gemm_kernel<EPI, FUSE_OUT, false, ROWEPI>, dim3(workers), dim3(N_THREADS), SMEM_BYTES, st, 1, probs, maps,
               probs_out, cmaps, up, ctx);
  else
    launch_pdl(umma_gemm_kernel<EPI, FUSE_OUT, true, ROWEPI>, dim3(2 * workers), dim3(N_THREADS), SMEM_BYTES, st, 2, probs, maps,
               probs_out, cmaps, up, ctx);
}

void launch_umma_gemm(int mode, const GemmProb* probs, const void* maps, const GemmProb* probs_out, int epi, int nprob,
                      int maxM, int maxN, const StepCtx& ctx, cudaStream_t st, int split3, int fuse_count, bool cta2,
                      int maxK, const void* cmaps, int tile_n_arg) {
  const int tile_n = tile_n_arg > 0 ? tile_n_arg : umma_tile_n(maxN);
  UmmaParams up = make_params(mode, tile_n, cta2);
  up.tiles_m = (maxM + up.tile_m - 1) / up.tile_m;
  up.tiles_n = (maxN + tile_n - 1) / tile_n;
  up.total_tiles = nprob * up.tiles_m * up.tiles_n;
  if (split3 == 2) {  // A is TF32-exact already (a stored activation): passes A Whi, A Wlo
    up.n_split = 2;
    up.maps_per_prob = 4;
    up.a_sel[0] = 0; up.b_sel[0] = 1;
    up.a_sel[1] = 2; up.b_sel[1] = 3;
  } else if (split3) {  // passes: Xhi Whi, Xlo Whi, Xhi Wlo
    up.n_split = 3;
    up.maps_per_prob = 4;
    up.a_sel[0] = 0; up.b_sel[0] = 1;
    up.a_sel[1] = 2; up.b_sel[1] = 1;
    up.a_sel[2] = 0; up.b_sel[2] = 3;
  }
  up.fuse_count = probs_out ? fuse_count : 0;
  // per-CTA stage: 16 KB of A + the B rows this CTA stages (multiples of 4 KB: swizzle-atom alignment holds)
  up.stage_bytes = STAGE_A_BYTES + (cta2 ? tile_n / 2 : tile_n) * TILE_K * 4;
  up.n_stages = (N_STAGES * STAGE_BYTES) / up.stage_bytes;
  if (up.n_stages > MAX_STAGES) up.n_stages = MAX_STAGES;
  // fused bias gradient needs all rows of a problem in one CTA (or CTA pair): only worth it while that leaves
  // enough independent units to fill the GPU (batch <= 256); larger batches use tile-major order + colsum_kernel
  up.k_max = maxK;
  up.stamp_slot = mode == 1 ? ST_DGRAD : (mode == 0 ? ST_POLHEAD : (maxN >= TILE_N ? ST_WGRAD : ST_FWGRAD));
  up.dbg = (int)env_u32("IQL_UMMA_DBG", 0);
  up.prob_major = (epi == EPI_DRELU && umma_dgrad_writes_dbias(maxM)) ? 1 : 0;
  up.tiles_per_unit = up.prob_major ? up.tiles_m : 1;
  up.units = up.total_tiles / up.tiles_per_unit;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  const int max_workers = cta2 ? n_sm / 2 : n_sm;  // persistent: one CTA per SM
  const int workers = up.units < max_workers ? up.units : max_workers;
  const CUtensorMap* m = (const CUtensorMap*)maps;
  const CUtensorMap* cm = (const CUtensorMap*)cmaps;
  if (epi == EPI_RELU && probs_out) launch_variant<EPI_RELU, true, false>(cta2, workers, probs, m, probs_out, nullptr, up, ctx, st);
  else if (epi == EPI_RELU) launch_variant<EPI_RELU, false, false>(cta2, workers, probs, m, nullptr, nullptr, up, ctx, st);
  else if (epi == EPI_DRELU && cm) launch_variant<EPI_DRELU, false, true>(cta2, workers, probs, m, nullptr, cm, up, ctx, st);
  else if (epi == EPI_DRELU) launch_variant<EPI_DRELU, false, false>(cta2, workers, probs, m, nullptr, nullptr, up, ctx, st);
  else if (epi == EPI_LINEAR) launch_variant<EPI_LINEAR, false, false>(cta2, workers, probs, m, nullptr, nullptr, up, ctx, st);
  else if (cm) launch_variant<EPI_NONE, false, true>(cta2, workers, probs, m, nullptr, cm, up, ctx, st);
  else launch_variant<EPI_NONE, false, false>(cta2, workers, probs, m, nullptr, nullptr, up, ctx, st);
}

void launch_colsum(const GemmProb* probs, int nprob, int maxM, cudaStream_t st) {
  dim3 grid((maxM + 31) / 32, nprob);
  colsum_kernel<<<grid, 1024, 0, st>>>(probs);
}

}  // namespace iql
