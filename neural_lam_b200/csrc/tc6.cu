// Edge kernel for H = 64, receiver-sorted edge sets whose EDGE FEATURES ARE BATCH-BROADCAST (stride-0 batch: the
// reference's expand_to_batch of a static edge embedding, graph/base.py:298-306) and that do not update their edges
// (update_edges=False: grid->mesh encoder) while the SENDER set is large next to the edge set (every grid node feeds
// 1-2 mesh nodes, so a per-node projection pass would cost more than the edge work itself).
//
//   z_e,b = W1e·e_e + W1s·x_s[b, src(e)] + (W1r·x_r + b1)[b, dst(e)]
//           \_ Ze: the same for every batch _/   \_ GEMM on gathered RAW sender rows _/   \_ receiver projection _/
//
//   * work item = (tile t, batch b), TILE-MAJOR: a CTA runs a contiguous range of items, i.e. all batches of a tile
//     back to back.  The edge term Ze_t = e_t·W1eᵀ is computed ONCE per tile into TMEM and added by epilogue 1 for
//     every batch: no per-(tile, batch) edge traffic at all;
//   * shared memory is a ring of FIVE 32 KB slots.  An item takes one slot: the 128 x 64 tile of gathered sender rows
//     (TMA tile::gather4 straight from the (B·Ns, 64) sender tensor, one row per edge), which holds the messages after
//     the first GEMM has consumed it (segmented sum over the tile's CSR receiver segments).  When the tile changes
//     the e tile travels through the same ring as a pseudo entry in front of the tile's first item;
//   * PING-PONG EPILOGUE GROUPS: two groups of 8 warps (thread = tile row x 32 columns); group g runs epilogue 1,
//     the segmented sum of its previous item and epilogue 2 for the items of parity g.  Measured before (one group per
//     epilogue stage): every epilogue instruction stream is latency-bound (5 warps per scheduler), epilogue 1 + sum
//     took 3.3 k cycles per item on one group while the other idled half the time; two independent item pipelines
//     balance whatever the split of work between the stages is.  All synchronisation is by mbarrier (no group-wide
//     named barriers: the slowest of 8 warps no longer sets the pace three times per item);
//   * three TMEM stages (D | hidden) + the Ze accumulator; LayerNorm exchange scratch in the stage's dead hidden
//     columns; 640 threads: 2 x 8 epilogue warps, MMA issue (uniform control flow, one elected lane), 3 loader warps
//     (a tile::gather4 costs its warp ~100 issue cycles per active lane).
// Same packed-fp32 / halved-weights SiLU arithmetic as tc5.cu.  Replaces, for this call shape, the per-edge work of
// InteractionNet.forward (reference gnn_layers.py:144-189): index_select gathers, cat, edge_mlp, scatter-sum.
#include "tc_ptx.cuh"

namespace nlam {

namespace e6 {
#ifndef NLAM_E6_NG
#define NLAM_E6_NG 2
#endif
#ifndef NLAM_E6_RNOW
#define NLAM_E6_RNOW 0
#endif
constexpr int NG = NLAM_E6_NG;       // epilogue groups (items it % NG)
constexpr int THREADS = 32 * (8 * NG + 4);
constexpr int EPI = 256;              // threads of one epilogue group
constexpr int LD_WARPS = 3;
constexpr int LD_THREADS = 32 * LD_WARPS;
constexpr int W_MMA = 8 * NG, W_LD = 8 * NG + 1;
constexpr int NR = 5;   // ring slots
constexpr int NT = 3;   // TMEM stages (D | hidden)
constexpr int PF = 4;   // sender rows are prefetched into L2 this many batches ahead
constexpr uint32_t BLK = 16384;
constexpr uint32_t WBLK = 8192;
constexpr uint32_t OFF_W1E = 0;
constexpr uint32_t OFF_W1S = 2 * WBLK;
constexpr uint32_t OFF_W2 = 4 * WBLK;
constexpr uint32_t OFF_RING = 6 * WBLK;          // slot r: [block 0 | block 1]
constexpr uint32_t OFF_MISC = OFF_RING + NR * 2 * BLK;
constexpr uint32_t OFF_LNX = OFF_MISC + 4096;    // LayerNorm exchange scratch: [NG][item parity][half][128 rows] float2
constexpr uint32_t SMEM = OFF_LNX + NG * 4096;
constexpr bool REDUCE_NOW = NLAM_E6_RNOW != 0;                // segmented sum of an item right after its epilogue 2
constexpr uint32_t TM_ZE = 384;                  // TMEM columns of the per-tile edge term
}  // namespace e6

struct Edge6Params {
  const int32_t* src;
  const int32_t* dst;
  int send_rows;  // rows per batch of the sender tensor (0: sender rows are batch-broadcast)
  const float* xs;  // the sender tensor (for the L2 prefetch of the batches to come)
  const float* pr;   // (n_rec, 64) receiver projection W1r·x_r + b1 (batch-broadcast receivers)
  const float* b2;
  const float* gamma;
  const float* beta;
  float eps;
  float* aggr;
  int mean;
  long long n_edges;
  long long n_rec;
  int B;
  int n_tiles;
  const int32_t* tile_e0;
  const int4* tile_meta;
  const int32_t* rowptr;
  int items_per_cta;
  long long* dbg;
};

#define E6_DBG(slot, it)                                                                        \
  do {                                                                                          \
    if (p.dbg && blockIdx.x == 1 && (it) >= 40 && (it) < 56) p.dbg[((it) - 40) * 16 + (slot)] = clock64(); \
  } while (0)

__global__ void __launch_bounds__(e6::THREADS, 1)
tc_edge_bcast_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmXs,
                     const Edge6Params p) {
  using namespace e6;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if ((sbase & 1023u) != 0) {
    if (tid == 0) printf("nlam tc_edge_bcast: dynamic shared memory not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t mb = sbase + OFF_MISC;
  const uint32_t bar_w = mb + 0;
  const uint32_t bar_wscaled = mb + 8;     // W1e / W1s halved in place (256 arrivals)
  const uint32_t bar_full = mb + 16;       // [NR] ring slot filled (tx bytes)
  const uint32_t bar_free = mb + 56;       // [NR] ring slot released (256 arrivals of ONE epilogue group)
  const uint32_t bar_staged = mb + 96;     // [6] messages of item it (index it % 6) written to its slot (256 arrivals);
                                           // per ITEM, not per slot: e entries pass through the slots without staging
  const uint32_t bar_d1_full = mb + 144;   // [NT]
  const uint32_t bar_hb_full = mb + 168;   // [NT] 256 arrivals
  const uint32_t bar_d2_full = mb + 192;   // [NT]
  const uint32_t bar_d_free = mb + 216;    // [NT] accumulators of the TMEM stage drained by epilogue 2 (256 arrivals)
  const uint32_t bar_zfix = mb + 240;      // receiver term added into Ze for the current tile (256 arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 252);
  int* lp_all = reinterpret_cast<int*>(smem + OFF_MISC + 256);      // [NR][132] local CSR offsets of the item in slot r
  float* sprm = reinterpret_cast<float*>(smem + OFF_MISC + 3072);  // b2 | gamma | beta

  if (warp == W_MMA) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_wscaled, EPI);
      for (int s = 0; s < NR; ++s) {
        mbar_init(bar_full + 8 * s, 1);
        mbar_init(bar_free + 8 * s, EPI);
      }
      for (int s = 0; s < 6; ++s) mbar_init(bar_staged + 8 * s, EPI);
      mbar_init(bar_zfix, EPI);
      for (int s = 0; s < NT; ++s) {
        mbar_init(bar_d1_full + 8 * s, 1);
        mbar_init(bar_hb_full + 8 * s, EPI);
        mbar_init(bar_d2_full + 8 * s, 1);
        mbar_init(bar_d_free + 8 * s, EPI);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == W_LD && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmE) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmXs) : "memory");
  }
  pdl_launch_dependents();
  pdl_wait();  // everything below may read what the previous kernel in the stream wrote
  if (tid < 64) {
    sprm[tid] = p.b2[tid];
    sprm[64 + tid] = p.gamma[tid];
    sprm[128 + tid] = p.beta[tid];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  // TMEM columns: stage s: D at s*128 (first GEMM, then second), hidden at +64 (LayerNorm scratch there after the
  // second GEMM); Ze at 384
  const long long n_work = (long long)p.n_tiles * p.B;
  const long long w_begin = (long long)blockIdx.x * p.items_per_cta;
  const long long w_end = min(n_work, w_begin + p.items_per_cta);
  const int n_my = w_end > w_begin ? (int)(w_end - w_begin) : 0;
  // Ring positions: every role walks the same sequence.  Item `it` (tile t, batch b) sits at ring position
  // pos_x = it + (number of tile changes among items 0..it); a tile's e entry sits right in front of its first item.

  if (warp >= W_LD) {
    // =============================== loaders (3 warps) ===============================
    const uint64_t pol_keep = policy_evict_last();
    const uint64_t pol_stream = policy_evict_normal();  // grid rows are read ~1.6 times (neighbouring receivers)
    const int lw = warp - W_LD;
    if (lw == 0 && lane == 0) {
      mbar_expect_tx(bar_w, 6u * WBLK);
      for (int j = 0; j < 4; ++j) tma_load_2d(sbase + OFF_W1E + j * WBLK, &tmW1, bar_w, 32 * j, 0);  // W1e | W1s
      for (int j = 0; j < 2; ++j) tma_load_2d(sbase + OFF_W2 + j * WBLK, &tmW2, bar_w, 32 * j, 0);
    }
    // the 64 gathers of an item (32 row groups x 2 column blocks) are spread over the 96 loader lanes: op = 22*lw + lane
    const int op = 22 * lw + lane;
    const bool issuer = lane < 22 && op < 64;
    const int cb = op >> 5;
    const int r_first = 4 * (op & 31);
    int4 ids = make_int4(0, 0, 0, 0);
    int t_cur = -1, pos = 0;
    bool fresh = false;
    int t = (int)(w_begin / p.B), b = (int)(w_begin - (long long)t * p.B) - 1;
    for (int it = 0; it < n_my; ++it) {
      if (++b == p.B) {
        b = 0;
        ++t;
      }
      if (t != t_cur) {  // new tile: sender ids of its rows (rows past the edge set read row 0: never used)
        const int e0 = __ldg(p.tile_e0 + t);
        if (issuer) {
          int v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const long long e = (long long)e0 + r_first + j;
            v[j] = (e < p.n_edges) ? __ldg(p.src + e) : 0;
          }
          ids = make_int4(v[0], v[1], v[2], v[3]);
        }
        if (lw == 0 && lane == 0) {  // the tile's edge features: one ring entry
          const int r = pos % NR;
          mbar_wait(bar_free + 8 * r, (uint32_t)(((pos / NR) & 1) ^ 1));
          mbar_expect_tx(bar_full + 8 * r, 2u * BLK);
          tma_load_3d(sbase + OFF_RING + r * 2 * BLK, &tmE, bar_full + 8 * r, 0, e0, 0, pol_keep);
          tma_load_3d(sbase + OFF_RING + r * 2 * BLK + BLK, &tmE, bar_full + 8 * r, 32, e0, 0, pol_keep);
        }
        t_cur = t;
        fresh = true;
        ++pos;
      }
      const int r = pos % NR;
      const uint32_t full = bar_full + 8 * r;
      if (lw == 0) {
        if (lane == 0) {
          mbar_wait(bar_free + 8 * r, (uint32_t)(((pos / NR) & 1) ^ 1));
          E6_DBG(0, it);
          mbar_expect_tx(full, 2u * BLK);
        }
        __syncwarp();
      }
      named_bar_sync(1, LD_THREADS);
      const int boff = p.send_rows * b;
      if (issuer)
        tma_gather4(sbase + OFF_RING + r * 2 * BLK + cb * BLK + r_first * 128, &tmXs, full, 32 * cb, ids.x + boff, ids.y + boff,
                    ids.z + boff, ids.w + boff, pol_stream);
      // The same rows of the batches to come are pulled into L2 now: a slot's life starts with its gather, 4.3 k cycles from DRAM
      // (measured); this lane's four 128-byte row pieces of batch b + PF cost four prefetch instructions.
      if (issuer && p.send_rows > 0) {
        // (the first item of a tile also covers the batches in between)
        for (int j = fresh ? 1 : PF; j <= PF; ++j) {
          if (b + j >= p.B) break;
          const float* base = p.xs + ((long long)p.send_rows * (b + j)) * 64 + 32 * cb;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (long long)ids.x * 64));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (long long)ids.y * 64));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (long long)ids.z * 64));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (long long)ids.w * 64));
        }
      }
      fresh = false;
      ++pos;
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issue (uniform control flow, one elected lane) ===============================
    const uint32_t idesc = umma_idesc_tf32(128, 64);
    mbar_wait(bar_w, 0);
    mbar_wait(bar_wscaled, 0);
    tc_fence_after();
    const uint64_t desc_w1e = umma_desc(sbase + OFF_W1E);
    const uint64_t desc_w1s = umma_desc(sbase + OFF_W1S);
    const uint64_t desc_w2 = umma_desc(sbase + OFF_W2);
    const uint64_t desc_ring = umma_desc(sbase + OFF_RING);
    int g1 = 0, g2 = 0, t_cur = -1, pos = 0;  // pos: ring position of the next entry to consume
    int t = (int)(w_begin / p.B), b_left = p.B - (int)(w_begin - (long long)t * p.B);  // tile of item g1, items left in it
    uint32_t idle = 0;
    while (g2 < n_my) {
      bool progress = false;
      if (g1 < n_my && g1 <= g2 + 2) {
        const int ts = g1 % NT;
        const bool new_tile = t != t_cur;
        const int pos_x = pos + (new_tile ? 1 : 0);
        const int rx = pos_x % NR;
        bool ready = mbar_test_u(bar_full + 8 * rx, (uint32_t)((pos_x / NR) & 1));
        // the TMEM stage must have been drained by epilogue 2 of item g1 - NT
        if (ready && g1 >= NT) ready = mbar_test_u(bar_d_free + 8 * ts, (uint32_t)(((g1 / NT) - 1) & 1));
        if (ready && new_tile) {
          // every earlier item must have left epilogue 1 (it reads Ze) before Ze is overwritten: their second GEMMs
          // have all been issued (g2 == g1), and those waited for epilogue 1
          ready = (g2 == g1) && mbar_test_u(bar_full + 8 * (pos % NR), (uint32_t)((pos / NR) & 1));
        }
        if (ready) {
          tc_fence_after();
          if (new_tile) {
            const uint64_t ae = desc_ring + (uint64_t)(((pos % NR) * 2 * BLK) >> 4);
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_tf32(tmem_base + TM_ZE, ae + (uint64_t)((j * BLK) >> 4) + 2 * k,
                            desc_w1e + (uint64_t)((j * WBLK) >> 4) + 2 * k, idesc, (uint32_t)((j | k) != 0));
            }
            __syncwarp();
            t_cur = t;
          }
          if (lane == 0) E6_DBG(1, g1);
          const uint32_t dd = tmem_base + ts * 128;
          const uint64_t a0 = desc_ring + (uint64_t)((rx * 2 * BLK) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32(dd, a0 + (uint64_t)((j * BLK) >> 4) + 2 * k, desc_w1s + (uint64_t)((j * WBLK) >> 4) + 2 * k, idesc,
                          (uint32_t)((j | k) != 0));
            umma_commit(bar_d1_full + 8 * ts);  // covers the Ze GEMM issued just before, too
          }
          __syncwarp();
          pos = pos_x + 1;
          ++g1;
          if (--b_left == 0) {
            b_left = p.B;
            ++t;
          }
          progress = true;
        }
      }
      if (g2 < g1) {
        const int ts = g2 % NT;
        if (mbar_test_u(bar_hb_full + 8 * ts, (uint32_t)((g2 / NT) & 1))) {
          tc_fence_after();
          if (lane == 0) E6_DBG(2, g2);
          const uint32_t dd = tmem_base + ts * 128;
          const uint32_t ht = dd + 64;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_ts(dd, ht + (uint32_t)(j * 32 + k * 8), desc_w2 + (uint64_t)((j * WBLK) >> 4) + 2 * k, idesc,
                             (uint32_t)((j | k) != 0));
            umma_commit(bar_d2_full + 8 * ts);
          }
          __syncwarp();
          ++g2;
          progress = true;
        }
      }
      if (progress) idle = 0;
      else if (__nanosleep(32), ++idle > (1u << 24)) {
        if (lane == 0) printf("nlam tc_edge_bcast: MMA issuer timeout (block %d g1 %d g2 %d)\n", blockIdx.x, g1, g2);
        __trap();
      }
    }
  } else {
    // =============================== epilogue groups (ping-pong over the items) ===============================
    const int grp = warp >> 3;            // group g: items g, g + NG, g + 2 NG, ...
    const int gw = warp & 7;              // warp within the group
    const int gt = tid & 255;             // thread within the group
    const int q = gw & 3;
    const int half = gw >> 2;
    const int row = q * 32 + lane;
    const int c0 = half * 32;
    const int rxs = row & 7;
    const uint32_t rsw = (uint32_t)(row * 128);
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    const int pbar = 2 + 4 * grp + q;
    // SiLU(z) = h + h*tanh(h), h = z/2: W1e and W1s are halved in place once (exact); the receiver term is halved in
    // the FMA that adds it
    if (grp == 0) {
      mbar_wait(bar_w, 0);
      float4* wq = reinterpret_cast<float4*>(smem + OFF_W1E) + gt;
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // 32 KB = 2048 float4 over 256 threads
        float4 x = wq[i * EPI];
        x.x *= 0.5f;
        x.y *= 0.5f;
        x.z *= 0.5f;
        x.w *= 0.5f;
        wq[i * EPI] = x;
      }
      fence_proxy_async();
      mbar_arrive(bar_wscaled);
    }

    // segmented sum of one of this group's items: slot r (use number u of that slot), batch br, receivers r0.. (nrec)
    auto reduce_item = [&](int itr, int r, int br, int r0, int nrec) {
      const int* lp = lp_all + r * 132;
      // messages staged by this group's epilogue 2; the offsets were written by its threads before epilogue 1 of the
      // item released the hidden activations, i.e. before the second GEMM the staged messages came from
      if (lane == 0) mbar_wait(bar_staged + 8 * (itr % 6), (uint32_t)((itr / 6) & 1));
      __syncwarp();
      if (gt == 0) E6_DBG(8, itr);
      const int cg = gt & 15, g = gt >> 4;
      const uint8_t* mbase = smem + OFF_RING + r * 2 * BLK + (cg >> 3) * BLK;
      const int chq = cg & 7;
      // LayerNorm's affine part is applied here, once per receiver and column, instead of once per edge in epilogue 2:
      // sum_e (gamma*n_e + beta) = gamma * sum_e n_e + deg * beta  (this kernel never writes the messages themselves)
      const float4 gam = *reinterpret_cast<const float4*>(sprm + 64 + 4 * cg);
      const float4 bet = *reinterpret_cast<const float4*>(sprm + 128 + 4 * cg);
      for (int j = g; j < nrec; j += EPI / 16) {
        const int k0 = lp[j], k1 = lp[j + 1];
        // rows k0..k1-1 in CSR order, as two interleaved chains (even / odd position) of packed adds
        float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
        int k = k0;
        for (; k + 2 <= k1; k += 2) {
          const float4 x = *reinterpret_cast<const float4*>(mbase + swz(k, chq));
          const float4 y = *reinterpret_cast<const float4*>(mbase + swz(k + 1, chq));
          a0 = add2(a0, make_float2(x.x, x.y));
          a1 = add2(a1, make_float2(x.z, x.w));
          b0 = add2(b0, make_float2(y.x, y.y));
          b1 = add2(b1, make_float2(y.z, y.w));
        }
        if (k < k1) {
          const float4 x = *reinterpret_cast<const float4*>(mbase + swz(k, chq));
          a0 = add2(a0, make_float2(x.x, x.y));
          a1 = add2(a1, make_float2(x.z, x.w));
        }
        a0 = add2(a0, b0);
        a1 = add2(a1, b1);
        float4 acc = make_float4(a0.x, a0.y, a1.x, a1.y);
        float cnt = (float)(k1 - k0);
        if (p.mean) {
          const float sc = 1.0f / (float)max(k1 - k0, 1);
          acc.x *= sc;
          acc.y *= sc;
          acc.z *= sc;
          acc.w *= sc;
          cnt = k1 > k0 ? 1.f : 0.f;
        }
        acc.x = fmaf(acc.x, gam.x, cnt * bet.x);
        acc.y = fmaf(acc.y, gam.y, cnt * bet.y);
        acc.z = fmaf(acc.z, gam.z, cnt * bet.z);
        acc.w = fmaf(acc.w, gam.w, cnt * bet.w);
        *reinterpret_cast<float4*>(p.aggr + ((long long)br * p.n_rec + r0 + j) * 64 + cg * 4) = acc;
      }
      mbar_arrive(bar_free + 8 * r);  // this thread is done with the slot (and its offsets)
      if (gt == 0) E6_DBG(7, itr);
    };

    int t_seen = -1, pos = 0, n_tile = 0;   // ring walk (all items); n_tile: tiles seen so far
    int t_mine = -1, lp_cur = 0, r0_cur = 0, nrec_cur = 0;  // tile data of this group's current item
    int prev_it = -1, prev_r = 0, prev_b = 0, prev_r0 = 0, prev_nrec = 0;
    int t = (int)(w_begin / p.B), b = (int)(w_begin - (long long)t * p.B) - 1;
    for (int it = 0; it < n_my; ++it) {
      if (++b == p.B) {
        b = 0;
        ++t;
      }
      const bool new_tile = t != t_seen;
      int pos_e = -1;
      if (new_tile) {
        pos_e = pos;
        ++pos;
        ++n_tile;
        t_seen = t;
      }
      const int pos_x = pos;
      ++pos;
      if (it % NG != grp) continue;
      const int ts = it % NT;
      const int r = pos_x % NR;
      const bool first_of_mine = t != t_mine;
      int my_dst = 0;
      if (first_of_mine) {
        const int4 m0 = __ldg(p.tile_meta + t);
        r0_cur = m0.z;
        nrec_cur = m0.w;
        if (new_tile) my_dst = (m0.x + row < p.n_edges) ? __ldg(p.dst + m0.x + row) : 0;
        lp_cur = (gt <= m0.w) ? __ldg(p.rowptr + m0.z + gt) - m0.x : 0;
        t_mine = t;
      }
      if (lane == 0) mbar_wait(bar_d1_full + 8 * ts, (uint32_t)((it / NT) & 1));  // implies the tile's Ze GEMM completed
      __syncwarp();
      tc_fence_after();
      // the first GEMM of this item ran, so its slot had been released by all 256 threads of the group that used it
      // before: the slot's offset table may be rewritten; it is read again only after the second GEMM, which waits
      // for every thread's arrival below.  The e entry of a new tile was consumed by the Ze GEMM: release it.
      if (gt <= nrec_cur) lp_all[r * 132 + gt] = lp_cur;
      const uint32_t d1 = tmem_base + ts * 128 + t_lane + c0;
      const uint32_t ze = tmem_base + TM_ZE + t_lane + c0;
      if (new_tile) {
        // This group runs the tile's first item: fold the receiver term into the tile's edge term once,
        // Z = (W1e·e + W1r·x_r[dst] + b1) / 2 — the receiver rows are batch-broadcast, so Z serves every batch of the
        // tile and no epilogue holds receiver rows in registers.
        mbar_arrive(bar_free + 8 * (pos_e % NR));
        const float4* prow = reinterpret_cast<const float4*>(p.pr + (long long)my_dst * 64 + c0);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float4 q4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) q4[k] = __ldg(prow + 4 * c + k);
          float z[16];
          tmem_ld16(ze + 16 * c, z);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            z[4 * k + 0] = fmaf(q4[k].x, 0.5f, z[4 * k + 0]);
            z[4 * k + 1] = fmaf(q4[k].y, 0.5f, z[4 * k + 1]);
            z[4 * k + 2] = fmaf(q4[k].z, 0.5f, z[4 * k + 2]);
            z[4 * k + 3] = fmaf(q4[k].w, 0.5f, z[4 * k + 3]);
          }
          tmem_st16(ze + 16 * c, z);
        }
        tc_fence_before();
        mbar_arrive(bar_zfix);
        tc_fence_after();
      } else if (first_of_mine) {
        // the other group ran the tile's first item and folds the receiver term into Ze
        if (lane == 0) mbar_wait(bar_zfix, (uint32_t)((n_tile - 1) & 1));
        __syncwarp();
        tc_fence_after();
      }
      if (gt == 0) E6_DBG(3, it);
      {
        // one TMEM round trip in, one out (a dependent TMEM access costs ~300 cycles)
        float v[32], z[32];
        tmem_ld32_nowait(d1, v);
        tmem_ld32_nowait(ze, z);
        tmem_ld_wait();
        reg_fence32(v);
        reg_fence32(z);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float2 h0 = add2(make_float2(v[4 * k + 0], v[4 * k + 1]), make_float2(z[4 * k + 0], z[4 * k + 1]));
          const float2 h1 = add2(make_float2(v[4 * k + 2], v[4 * k + 3]), make_float2(z[4 * k + 2], z[4 * k + 3]));
          const float2 o0 = fma2(h0, make_float2(tanh_fast(h0.x), tanh_fast(h0.y)), h0);
          const float2 o1 = fma2(h1, make_float2(tanh_fast(h1.x), tanh_fast(h1.y)), h1);
          v[4 * k + 0] = o0.x;
          v[4 * k + 1] = o0.y;
          v[4 * k + 2] = o1.x;
          v[4 * k + 3] = o1.y;
        }
        tmem_st32_nowait(d1 + 64, v);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(bar_hb_full + 8 * ts);
      if (gt == 0) E6_DBG(4, it);
      if (!REDUCE_NOW && prev_it >= 0) reduce_item(prev_it, prev_r, prev_b, prev_r0, prev_nrec);

      // ---- epilogue 2: bias, LayerNorm -> messages into the item's slot (the gathered rows were consumed by GEMM 1)
      if (lane == 0) mbar_wait(bar_d2_full + 8 * ts, (uint32_t)((it / NT) & 1));
      __syncwarp();
      tc_fence_after();
      if (gt == 0) E6_DBG(5, it);
      {
        float vf[32];
        tmem_ld32(tmem_base + ts * 128 + t_lane + c0, vf);
        tc_fence_before();
        mbar_arrive(bar_d_free + 8 * ts);  // the accumulators of this TMEM stage are in registers
        float2 v[16];
        float2 sm2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 bb = *reinterpret_cast<const float4*>(sprm + c0 + 4 * k);
          v[2 * k] = add2(make_float2(vf[4 * k], vf[4 * k + 1]), make_float2(bb.x, bb.y));
          v[2 * k + 1] = add2(make_float2(vf[4 * k + 2], vf[4 * k + 3]), make_float2(bb.z, bb.w));
        }
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          sm2 = add2(sm2, add2(v[i], v[i + 1]));
          sq2 = fma2(v[i], v[i], sq2);
          sq2 = fma2(v[i + 1], v[i + 1], sq2);
        }
        // the two column halves of a row exchange (sum, sum of squares) through shared memory: one 64-thread
        // barrier per item (double-buffered by item parity so that no second barrier is needed)
        float2* lnx = reinterpret_cast<float2*>(smem + OFF_LNX) + grp * 512 + ((it / NG) & 1) * 256;
        const float my_s = sm2.x + sm2.y, my_q = sq2.x + sq2.y;
        lnx[half * 128 + row] = make_float2(my_s, my_q);
        named_bar_sync(pbar, 64);
        const float2 other = lnx[(half ^ 1) * 128 + row];
        const float mu = (my_s + other.x) * (1.0f / 64.0f);
        const float ex2 = (my_q + other.y) * (1.0f / 64.0f);
        const float rstd = rsqrtf(fmaxf(ex2 - mu * mu, 0.f) + p.eps);
        const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mu * rstd, -mu * rstd);
        uint8_t* mrow = smem + OFF_RING + r * 2 * BLK + half * BLK + rsw;
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // normalised values only: gamma / beta follow the segmented sum
          const float2 m0 = fma2(v[2 * k], rs2, nm2);
          const float2 m1 = fma2(v[2 * k + 1], rs2, nm2);
          *reinterpret_cast<float4*>(mrow + ((k ^ rxs) << 4)) = make_float4(m0.x, m0.y, m1.x, m1.y);
        }
      }
      mbar_arrive(bar_staged + 8 * (it % 6));
      if (gt == 0) E6_DBG(6, it);
      if (REDUCE_NOW) {
        // the ring is the scarce resource (a slot lives ~13 k cycles from the gather to its release): sum the item
        // right away instead of one item later, at the price of idling while the second GEMM runs
        reduce_item(it, r, b, r0_cur, nrec_cur);
        continue;
      }
      prev_it = it;
      prev_r = r;
      prev_b = b;
      prev_r0 = r0_cur;
      prev_nrec = nrec_cur;
    }
    if (!REDUCE_NOW && prev_it >= 0) reduce_item(prev_it, prev_r, prev_b, prev_r0, prev_nrec);
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == W_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host
bool tc_edge_bcast_supported(const NlamGraph* g, const NlamMlp* edge_mlp, int flags, const float* send, int64_t send_bs,
                             const float* rec, int64_t rec_bs, const float* edge, int64_t edge_bs, int B, bool has_edge_out) {
  static int on = -1;
  if (on < 0) on = getenv("NLAM_TC_NO_BCAST") ? 0 : 1;
  if (!on || has_edge_out || !tc_edge_supported(g, edge_mlp, flags)) return false;
  if (!(edge_bs == 0 || B == 1)) return false;  // edge features must be the same for every batch
  if (!(rec_bs == 0 || B == 1)) return false;   // ... and so must the receiver rows (folded into the per-tile term)
  if (!(aligned16(send) && aligned16(rec) && aligned16(edge))) return false;
  // gathered raw sender rows: one (B*Ns, 64) row space
  if (!(send_bs == 0 || B == 1 || send_bs % 64 == 0)) return false;
  const int64_t send_rows = (send_bs == 0 || B == 1) ? g->n_send : send_bs / 64;
  if (send_rows < g->n_send || send_rows * (int64_t)B >= (1LL << 31)) return false;
  return true;
}

size_t tc_edge_bcast_workspace_floats(const NlamGraph* g, int B, int64_t rec_bs) {
  return (size_t)((rec_bs == 0 || B == 1) ? 1 : B) * (size_t)g->n_rec * 64 + 64;
}

int tc_edge_bcast(const NlamGraph* g, const NlamMlp* edge_mlp, const float* send, int64_t send_bs, const float* rec,
                  int64_t rec_bs, const float* edge, float* aggr_out, int B, int flags, cudaStream_t st, float* ws) {
  NLAM_REQUIRE(aligned16(aggr_out) && aligned16(ws), NLAM_E_INVALID, "tc_edge_bcast: pointers must be 16-byte aligned");
  const int Bs = (send_bs == 0 || B == 1) ? 1 : B;
  const int Br = (rec_bs == 0 || B == 1) ? 1 : B;
  const int64_t nr = g->n_rec;
  const int64_t send_rows = Bs > 1 ? send_bs / 64 : g->n_send;
  const float* w1 = edge_mlp->w[0];  // (64, 192): columns [e | sender | receiver]
  float* Pr = ws;
  RowLinProblem pr = {rec, rec_bs, nr, Br, w1 + 128, 192, edge_mlp->b[0], Pr};
  int rc = rowlinear_multi(&pr, 1, st);
  if (rc) return rc;

  CUtensorMap me, mw1, mw2, mxs;
  rc = make_map(&me, edge, 64, (uint64_t)g->n_edges, 1, 64, (uint64_t)g->n_edges * 64, 128, true);
  if (rc) return rc;
  rc = make_map(&mw1, w1, 128, 64, 1, 192, 0, 64, false);  // the e and sender columns of W1
  if (rc) return rc;
  rc = make_map(&mw2, edge_mlp->w[1], 64, 64, 1, 64, 0, 64, false);
  if (rc) return rc;
  rc = make_map(&mxs, send, 64, (uint64_t)send_rows * Bs, 1, 64, 0, 1, false);
  if (rc) return rc;
  Edge6Params p;
  memset(&p, 0, sizeof(p));
  p.src = g->src;
  p.dst = g->dst;
  p.send_rows = Bs > 1 ? (int)send_rows : 0;
  p.xs = send;
  p.pr = Pr;
  p.b2 = edge_mlp->b[1];
  p.gamma = edge_mlp->ln_gamma;
  p.beta = edge_mlp->ln_beta;
  p.eps = edge_mlp->ln_eps;
  p.aggr = aggr_out;
  p.mean = (flags & NLAM_AGGR_MEAN) ? 1 : 0;
  p.n_edges = g->n_edges;
  p.n_rec = g->n_rec;
  p.B = B;
  p.n_tiles = g->n_tiles;
  p.tile_e0 = g->tile_e0;
  p.tile_meta = reinterpret_cast<const int4*>(g->tile_meta);
  p.rowptr = g->rowptr;
  static unsigned attr_mask = 0;
  int dev = 0;
  NLAM_CUDA_OK(cudaGetDevice(&dev));
  if (!(attr_mask & (1u << (dev & 31)))) {
    NLAM_CUDA_OK(cudaFuncSetAttribute(tc_edge_bcast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e6::SMEM));
    attr_mask |= 1u << (dev & 31);
  }
  const long long n_work = (long long)p.n_tiles * p.B;
  NLAM_REQUIRE(n_work < (1LL << 30), NLAM_E_UNSUPPORTED, "tc_edge_bcast: too many work items");
  const int sms = num_sms();
  p.items_per_cta = (int)((n_work + sms - 1) / sms);
  const int grid = (int)((n_work + p.items_per_cta - 1) / p.items_per_cta);
  static long long* dbg_buf = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) dbg_on = getenv("NLAM_TC_TIMELINE") ? 1 : 0;
  if (dbg_on) {
    if (!dbg_buf) NLAM_CUDA_OK(cudaMalloc(&dbg_buf, 256 * sizeof(long long)));
    NLAM_CUDA_OK(cudaMemsetAsync(dbg_buf, 0, 256 * sizeof(long long), st));
    p.dbg = dbg_buf;
  }
  {
    ProfScope ps("tc_edge_bcast_kernel", st, edge_algorithmic_bytes(g, B, send_bs, rec_bs, 0, false, 64));
    NLAM_CUDA_OK(launch_pdl(tc_edge_bcast_kernel, grid, e6::THREADS, e6::SMEM, st, me, mw1, mw2, mxs, p));
  }
  count_launch();
  if (dbg_on) {
    long long h[256];
    NLAM_CUDA_OK(cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
    NLAM_CUDA_OK(cudaStreamSynchronize(st));
    long long t0 = h[0];
    fprintf(stderr, "[nlam tc_edge_bcast timeline] grid=%d items/cta=%d (CTA 1, items 40..55; cycles rel. to the first)\n", grid,
            p.items_per_cta);
    fprintf(stderr, " it  ld_iss  g1_iss  g2_iss e1_start e1_done e2_start e2_done reduced red_beg e1_last\n");
    for (int it = 0; it < 16; ++it) {
      fprintf(stderr, "%3d ", it + 40);
      for (int k = 0; k < 10; ++k) fprintf(stderr, "%7lld ", h[it * 16 + k] ? h[it * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

}  // namespace nlam
