// Edge kernel for H = 64, general receiver-sorted edge sets, for the calls that do not need a SEPARATE edge output:
// the middle layers of a processor stack update their edge tensor IN PLACE (e += m, a TMA reduce-add of the staged
// message tile: cp.reduce.async.bulk.tensor ... .add), the last layer writes no edge output at all (nobody reads
// it, reference graph_lam.py:185).  Same math as tc5.cu,
//   z = W1e·e + (W1s·x)_src + (W1r·x + b1)_dst ,  m = LN(W2·SiLU(z) + b2) ,  aggr[r] = sum_{dst(e)=r} m_e ,  e' = e + m,
// but the e tile leaves shared memory as soon as the first GEMM has consumed it, because nothing adds to it on the
// SM any more.  Measured on tc5.cu: a shared-memory stage (e tile + sender window) lives ~14 k cycles per tile and
// three of them fit, which sets the tile period (4.4 k cycles) while both epilogue groups wait most of the time.
// Here:
//   * two rings: TWO e slots (load -> first GEMM, ~3 k cycles) and FOUR window slots (gathered sender projections ->
//     messages -> segmented sum + reduce-add store): four items in flight in the same 192 KB;
//   * ping-pong epilogue groups as in tc6.cu (group g runs epilogue 1, the segmented sum of its previous item and
//     epilogue 2 for the items of parity g), all hand-overs by mbarrier;
//   * work items (tile, batch) in TILE-MAJOR order, a CTA runs a contiguous range: the per-tile index tables (window
//     ids, window rows, receivers, CSR offsets) are loaded once per tile instead of once per item;
//   * rows of the 128-row message tile past the tile's own edges are written as zeros, so the reduce-add of the full TMA
//     box leaves the following tile's rows untouched (tiles hold whole receivers and are shorter than their box).
// In-place contract: edge_out == edge (dense batches).  Values are bit-identical to the out-of-place kernel (one fp32 add
// per element, done by the L2 instead of the SM).
#include "tc_ptx.cuh"

namespace nlam {

namespace e8 {
constexpr int NG = 2;
constexpr int THREADS = 640;
constexpr int EPI = 256;
constexpr int LD_THREADS = 64;
constexpr int W_MMA = 16, W_LD = 17, W_ST = 19;
constexpr int NE = 2;  // e slots
constexpr int NW = 4;  // window / message slots
constexpr int NT = 3;  // TMEM stages (D | hidden)
constexpr uint32_t BLK = 16384;
constexpr uint32_t WBLK = 8192;
constexpr uint32_t OFF_W1 = 0;
constexpr uint32_t OFF_W2 = 2 * WBLK;
constexpr uint32_t OFF_E = 4 * WBLK;                  // e slot s: [cols 0-31 | cols 32-63]
constexpr uint32_t OFF_WIN = OFF_E + NE * 2 * BLK;    // window slot s: [cols 0-31 | cols 32-63]
constexpr uint32_t OFF_MISC = OFF_WIN + NW * 2 * BLK;
constexpr uint32_t SMEM = OFF_MISC + 3072;  // 232448
constexpr uint32_t TM_LN = 384;             // LayerNorm exchange scratch columns
}  // namespace e8

struct Edge8Params {
  const int32_t* win_u;
  const int32_t* win_nu;
  const uint8_t* win_loc;
  const int32_t* dst;
  int ps_rows;
  const float* pr;
  long long pr_bs;
  const float* b2;
  const float* gamma;
  const float* beta;
  float eps;
  float* aggr;
  int has_out;
  int e_batched;
  int mean;
  long long n_edges;
  long long n_rec;
  int B;
  int n_tiles;
  const int32_t* tile_e0;
  const int4* tile_meta;
  const int32_t* rowptr;
  int items_per_cta;
  int e_policy;  // in place: L2 policy of the e tile loads (0 normal, 1 evict_last: the lines wait for the reduce-add)
  long long* dbg;
};

#define E8_DBG(slot, it)                                                                                   \
  do {                                                                                                     \
    if (p.dbg && blockIdx.x == 1 && (it) >= 40 && (it) < 56) p.dbg[((it) - 40) * 16 + (slot)] = clock64(); \
  } while (0)

// HAS_OUT: the edge tensor is updated in place (else: no edge output, and LayerNorm's affine part follows the segmented sum)
template <bool HAS_OUT>
__global__ void __launch_bounds__(e8::THREADS, 1)
tc_edge_rmw_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmW1,
                   const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut,
                   const __grid_constant__ CUtensorMap tmPs, const Edge8Params p) {
  using namespace e8;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if ((sbase & 1023u) != 0) {
    if (tid == 0) printf("nlam tc_edge_rmw: dynamic shared memory not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t mb = sbase + OFF_MISC;
  const uint32_t bar_w = mb + 0;
  const uint32_t bar_wscaled = mb + 8;     // W1e halved in place (256 arrivals)
  const uint32_t bar_e_full = mb + 16;     // [NE] e tile landed (tx bytes)
  const uint32_t bar_e_free = mb + 32;     // [NE] first GEMM has consumed the e tile (tcgen05.commit)
  const uint32_t bar_w_full = mb + 48;     // [NW] sender window landed (tx bytes)
  const uint32_t bar_w_free = mb + 80;     // [NW] segmented sum done (256 arrivals) [+ the reduce-add store has read the slot]
  const uint32_t bar_staged = mb + 112;    // [6] messages of item it (index it % 6) written to its slot (256 arrivals)
  const uint32_t bar_d1_full = mb + 160;   // [NT]
  const uint32_t bar_hb_full = mb + 184;   // [NT] 256 arrivals
  const uint32_t bar_d2_full = mb + 208;   // [NT]
  const uint32_t bar_d_free = mb + 232;    // [NT] accumulators drained by epilogue 2 (256 arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 256);
  uint16_t* lp_all = reinterpret_cast<uint16_t*>(smem + OFF_MISC + 264);  // [NW][132] local CSR offsets of the item in slot s
  float* sprm = reinterpret_cast<float*>(smem + OFF_MISC + 1344);         // b2 | gamma | beta

  if (warp == W_MMA) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_wscaled, EPI);
      for (int s = 0; s < NE; ++s) {
        mbar_init(bar_e_full + 8 * s, 1);
        mbar_init(bar_e_free + 8 * s, 1);
      }
      for (int s = 0; s < NW; ++s) {
        mbar_init(bar_w_full + 8 * s, 1);
        mbar_init(bar_w_free + 8 * s, EPI + (HAS_OUT ? 1 : 0));
      }
      for (int s = 0; s < 6; ++s) mbar_init(bar_staged + 8 * s, EPI);
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
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmPs) : "memory");
    if (HAS_OUT) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
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
  // TMEM columns: stage s: D at s*128 (first GEMM, then second), hidden at +64; LayerNorm exchange scratch at 384
  const long long n_work = (long long)p.n_tiles * p.B;
  const long long w_begin = (long long)blockIdx.x * p.items_per_cta;
  const long long w_end = min(n_work, w_begin + p.items_per_cta);
  const int n_my = w_end > w_begin ? (int)(w_end - w_begin) : 0;
  const int t_first = (int)(w_begin / p.B);
  const int b_first = (int)(w_begin - (long long)t_first * p.B);

  if (warp == W_ST) {
    // =============================== e += m (TMA reduce-add of the staged messages) ===============================
    if (HAS_OUT && lane == 0) {
      int t = t_first, b = b_first - 1, t_cur = -1, e0 = 0;
      for (int it = 0; it < n_my; ++it) {
        if (++b == p.B) {
          b = 0;
          ++t;
        }
        if (t != t_cur) {
          e0 = __ldg(p.tile_e0 + t);
          t_cur = t;
        }
        const int sw = it % NW;
        mbar_wait(bar_staged + 8 * (it % 6), (uint32_t)((it / 6) & 1));
        const uint32_t slot = sbase + OFF_WIN + sw * 2 * BLK;
        tma_reduce_add_3d(&tmOut, slot, 0, e0, b);
        tma_reduce_add_3d(&tmOut, slot + BLK, 32, e0, b);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar_w_free + 8 * sw);
        E8_DBG(11, it);
      }
      bulk_wait0();
    }
  } else if (warp >= W_LD) {
    // =============================== loaders (2 warps) ===============================
    const uint64_t pol_keep = policy_evict_last();
    // in place: the lines of the e tile should stay in L2 until the reduce-add reaches them
    const uint64_t pol_e = HAS_OUT ? (p.e_policy ? policy_evict_last() : policy_evict_normal()) : policy_evict_first();
    const int lw = warp - W_LD;
    if (lw == 0 && lane == 0) {
      mbar_expect_tx(bar_w, 4u * WBLK);
      for (int j = 0; j < 2; ++j) tma_load_2d(sbase + OFF_W1 + j * WBLK, &tmW1, bar_w, 32 * j, 0);
      for (int j = 0; j < 2; ++j) tma_load_2d(sbase + OFF_W2 + j * WBLK, &tmW2, bar_w, 32 * j, 0);
    }
    // lane l of loader warp lw issues the gather of window rows 4l..4l+3, column block lw
    int4 ids = make_int4(0, 0, 0, 0);
    int ngrp = 0, e0 = 0, t_cur = -1;
    int t = t_first, b = b_first - 1;
    for (int it = 0; it < n_my; ++it) {
      if (++b == p.B) {
        b = 0;
        ++t;
      }
      if (t != t_cur) {
        ngrp = __ldg(p.win_nu + t) >> 2;
        if (lane < ngrp) ids = __ldg(reinterpret_cast<const int4*>(p.win_u + (size_t)t * 128) + lane);
        e0 = __ldg(p.tile_e0 + t);
        t_cur = t;
      }
      const int se = it % NE, sw = it % NW;
      const uint32_t wfull = bar_w_full + 8 * sw;
      if (lw == 0) {
        if (lane == 0) {
          mbar_wait(bar_e_free + 8 * se, (uint32_t)(((it / NE) & 1) ^ 1));
          const uint32_t efull = bar_e_full + 8 * se;
          mbar_expect_tx(efull, 2u * BLK);
          const uint32_t es = sbase + OFF_E + se * 2 * BLK;
          tma_load_3d(es, &tmE, efull, 0, e0, p.e_batched ? b : 0, pol_e);
          tma_load_3d(es + BLK, &tmE, efull, 32, e0, p.e_batched ? b : 0, pol_e);
          mbar_wait(bar_w_free + 8 * sw, (uint32_t)(((it / NW) & 1) ^ 1));
          E8_DBG(0, it);
          mbar_expect_tx(wfull, (uint32_t)ngrp * 1024u);
        }
        __syncwarp();
      }
      named_bar_sync(12, LD_THREADS);
      if (lane < ngrp) {
        const int boff = p.ps_rows * b;
        tma_gather4(sbase + OFF_WIN + sw * 2 * BLK + lw * BLK + lane * 512, &tmPs, wfull, 32 * lw, ids.x + boff, ids.y + boff,
                    ids.z + boff, ids.w + boff, pol_keep);
      }
      if (lw == 1 && lane == 0 && p.e_batched && it + 3 < n_my) {  // pull the e tile three items ahead into L2
        int b3 = b + 3, t3 = t;
        while (b3 >= p.B) {
          b3 -= p.B;
          ++t3;
        }
        const int r3 = (t3 == t) ? e0 : __ldg(p.tile_e0 + t3);
        tma_prefetch_3d(&tmE, 0, r3, b3);
        tma_prefetch_3d(&tmE, 32, r3, b3);
      }
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issue (uniform control flow, one elected lane) ===============================
    const uint32_t idesc = umma_idesc_tf32(128, 64);
    mbar_wait(bar_w, 0);
    mbar_wait(bar_wscaled, 0);
    tc_fence_after();
    const uint64_t desc_w1 = umma_desc(sbase + OFF_W1);
    const uint64_t desc_w2 = umma_desc(sbase + OFF_W2);
    const uint64_t desc_e = umma_desc(sbase + OFF_E);
    int g1 = 0, g2 = 0;
    uint32_t idle = 0;
    while (g2 < n_my) {
      bool progress = false;
      if (g1 < n_my && g1 <= g2 + 2) {
        const int ts = g1 % NT, se = g1 % NE;
        bool ready = mbar_test_u(bar_e_full + 8 * se, (uint32_t)((g1 / NE) & 1));
        // the TMEM stage must have been drained by epilogue 2 of item g1 - NT
        if (ready && g1 >= NT) ready = mbar_test_u(bar_d_free + 8 * ts, (uint32_t)(((g1 / NT) - 1) & 1));
        if (ready) {
          tc_fence_after();
          if (lane == 0) E8_DBG(1, g1);
          const uint32_t dd = tmem_base + ts * 128;
          const uint64_t a0 = desc_e + (uint64_t)((se * 2 * BLK) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32(dd, a0 + (uint64_t)((j * BLK) >> 4) + 2 * k, desc_w1 + (uint64_t)((j * WBLK) >> 4) + 2 * k, idesc,
                          (uint32_t)((j | k) != 0));
            umma_commit(bar_d1_full + 8 * ts);
            umma_commit(bar_e_free + 8 * se);  // the e tile is dead once these MMAs have read it
          }
          __syncwarp();
          ++g1;
          progress = true;
        }
      }
      if (g2 < g1) {
        const int ts = g2 % NT;
        if (mbar_test_u(bar_hb_full + 8 * ts, (uint32_t)((g2 / NT) & 1))) {
          tc_fence_after();
          if (lane == 0) E8_DBG(2, g2);
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
        if (lane == 0) printf("nlam tc_edge_rmw: MMA issuer timeout (block %d g1 %d g2 %d)\n", blockIdx.x, g1, g2);
        __trap();
      }
    }
  } else {
    // =============================== epilogue groups (ping-pong over the items) ===============================
    const int grp = warp >> 3;  // group g: items g, g + NG, ...
    const int gw = warp & 7;
    const int gt = tid & 255;
    const int q = gw & 3;
    const int half = gw >> 2;
    const int row = q * 32 + lane;
    const int c0 = half * 32;
    const int rxs = row & 7;
    const uint32_t rsw = (uint32_t)(row * 128);
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    const int pbar = 2 + 4 * grp + q;
    const uint32_t ln_col = tmem_base + TM_LN + 8 * grp + t_lane;
    // SiLU(z) = h + h*tanh(h), h = z/2: W1e is halved in place once (exact); the gathered node terms are halved in the
    // FMA that adds them
    if (grp == 0) {
      mbar_wait(bar_w, 0);
      float4* wq = reinterpret_cast<float4*>(smem + OFF_W1) + gt;
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // 16 KB = 1024 float4 over 256 threads
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
    const float2 half2 = make_float2(0.5f, 0.5f);

    // segmented sum of one of this group's items: window slot sw, batch br, receivers r0.. (nrec)
    auto reduce_item = [&](int itr, int sw, int br, int r0, int nrec) {
      const uint16_t* lp = lp_all + sw * 132;
      if (lane == 0) mbar_wait(bar_staged + 8 * (itr % 6), (uint32_t)((itr / 6) & 1));
      __syncwarp();
      if (gt == 0) E8_DBG(8, itr);
      const int cg = gt & 15, g = gt >> 4;
      const uint8_t* mbase = smem + OFF_WIN + sw * 2 * BLK + (cg >> 3) * BLK;
      const int chq = cg & 7;
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
        if (!HAS_OUT) {
          const float4 gam = *reinterpret_cast<const float4*>(sprm + 64 + 4 * cg);
          const float4 bet = *reinterpret_cast<const float4*>(sprm + 128 + 4 * cg);
          acc.x = fmaf(acc.x, gam.x, cnt * bet.x);
          acc.y = fmaf(acc.y, gam.y, cnt * bet.y);
          acc.z = fmaf(acc.z, gam.z, cnt * bet.z);
          acc.w = fmaf(acc.w, gam.w, cnt * bet.w);
        }
        *reinterpret_cast<float4*>(p.aggr + ((long long)br * p.n_rec + r0 + j) * 64 + cg * 4) = acc;
      }
      mbar_arrive(bar_w_free + 8 * sw);  // this thread is done with the slot (and its offsets)
      if (gt == 0) E8_DBG(7, itr);
    };

    int t_mine = -1, lp_cur = 0, r0_cur = 0, nrec_cur = 0, ne_cur = 0, my_dst = 0, loc = 0;
    int prev_it = -1, prev_sw = 0, prev_b = 0, prev_r0 = 0, prev_nrec = 0;
    int t = t_first, b = b_first - 1;
    for (int it = 0; it < n_my; ++it) {
      if (++b == p.B) {
        b = 0;
        ++t;
      }
      if (it % NG != grp) continue;
      const int ts = it % NT, sw = it % NW;
      if (t != t_mine) {  // per-tile index data of this thread's row
        const int4 m0 = __ldg(p.tile_meta + t);
        r0_cur = m0.z;
        nrec_cur = m0.w;
        ne_cur = m0.y;
        my_dst = (m0.x + row < p.n_edges) ? __ldg(p.dst + m0.x + row) : 0;
        loc = __ldg(p.win_loc + (size_t)t * 128 + row);
        lp_cur = (gt <= m0.w) ? __ldg(p.rowptr + m0.z + gt) - m0.x : 0;
        t_mine = t;
      }
      // receiver projection row of this edge (rows of one CSR segment share it: the lanes' loads coalesce)
      const float4* prow = reinterpret_cast<const float4*>(p.pr + (long long)b * p.pr_bs + (long long)my_dst * 64 + c0);
      float4 pr[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) pr[k] = __ldg(prow + k);
      if (lane == 0) {
        mbar_wait(bar_w_full + 8 * sw, (uint32_t)((it / NW) & 1));  // sender window visible
        mbar_wait(bar_d1_full + 8 * ts, (uint32_t)((it / NT) & 1));
      }
      __syncwarp();
      tc_fence_after();
      // The window slot was refilled, so all 256 threads of this group had released it (their sum of item it - NW is
      // done): its offset table may be rewritten.  It is read after the second GEMM, which waits for every thread's
      // arrival below.
      if (gt <= nrec_cur) lp_all[sw * 132 + gt] = (uint16_t)lp_cur;
      if (gt == 0) E8_DBG(3, it);
      {
        const uint8_t* ps = smem + OFF_WIN + sw * 2 * BLK + half * BLK + loc * 128;
        const int rx = loc & 7;
        const uint32_t d1 = tmem_base + ts * 128 + t_lane + c0;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[16];
          tmem_ld16(d1 + 16 * c, v);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 s4 = *reinterpret_cast<const float4*>(ps + (((4 * c + k) ^ rx) << 4));
            const float4 q4 = pr[4 * c + k];
            const float2 h0 = fma2(add2(make_float2(s4.x, s4.y), make_float2(q4.x, q4.y)), half2,
                                   make_float2(v[4 * k + 0], v[4 * k + 1]));
            const float2 h1 = fma2(add2(make_float2(s4.z, s4.w), make_float2(q4.z, q4.w)), half2,
                                   make_float2(v[4 * k + 2], v[4 * k + 3]));
            const float2 o0 = fma2(h0, make_float2(tanh_fast(h0.x), tanh_fast(h0.y)), h0);
            const float2 o1 = fma2(h1, make_float2(tanh_fast(h1.x), tanh_fast(h1.y)), h1);
            v[4 * k + 0] = o0.x;
            v[4 * k + 1] = o0.y;
            v[4 * k + 2] = o1.x;
            v[4 * k + 3] = o1.y;
          }
          tmem_st16(d1 + 64 + 16 * c, v);
        }
      }
      tc_fence_before();
      mbar_arrive(bar_hb_full + 8 * ts);
      if (gt == 0) E8_DBG(4, it);
      if (prev_it >= 0) reduce_item(prev_it, prev_sw, prev_b, prev_r0, prev_nrec);

      // ---- epilogue 2: bias, LayerNorm -> messages into the item's window slot (epilogue 1 has consumed the window)
      if (lane == 0) mbar_wait(bar_d2_full + 8 * ts, (uint32_t)((it / NT) & 1));
      __syncwarp();
      tc_fence_after();
      if (gt == 0) E8_DBG(5, it);
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
        // the two column halves of a row exchange (sum, sum of squares) through spare TMEM columns of the row's lane;
        // scratch double-buffered by item parity: one 64-thread barrier per item
        const uint32_t scr = ln_col + 4 * ((it / NG) & 1);
        tmem_st2(scr + 2 * half, sm2.x + sm2.y, sq2.x + sq2.y);
        tc_fence_before();
        named_bar_sync(pbar, 64);
        tc_fence_after();
        float st4[4];
        tmem_ld4(scr, st4);
        const float mu = (st4[0] + st4[2]) * (1.0f / 64.0f);
        const float ex2 = (st4[1] + st4[3]) * (1.0f / 64.0f);
        const float rstd = rsqrtf(fmaxf(ex2 - mu * mu, 0.f) + p.eps);
        const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mu * rstd, -mu * rstd);
        uint8_t* mrow = smem + OFF_WIN + sw * 2 * BLK + half * BLK + rsw;
        if (!HAS_OUT) {
          // no edge output: only the aggregate is needed, and sum_e (gamma*n_e + beta) = gamma * sum_e n_e + deg * beta — the
          // affine part is applied once per receiver after the segmented sum (rows past the tile's edges are never summed)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float2 m0 = fma2(v[2 * k], rs2, nm2);
            const float2 m1 = fma2(v[2 * k + 1], rs2, nm2);
            *reinterpret_cast<float4*>(mrow + ((k ^ rxs) << 4)) = make_float4(m0.x, m0.y, m1.x, m1.y);
          }
        } else if (row < ne_cur) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 g4 = *reinterpret_cast<const float4*>(sprm + 64 + c0 + 4 * k);
            const float4 b4 = *reinterpret_cast<const float4*>(sprm + 128 + c0 + 4 * k);
            const float2 m0 = fma2(fma2(v[2 * k], rs2, nm2), make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
            const float2 m1 = fma2(fma2(v[2 * k + 1], rs2, nm2), make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
            *reinterpret_cast<float4*>(mrow + ((k ^ rxs) << 4)) = make_float4(m0.x, m0.y, m1.x, m1.y);
          }
        } else {
          // rows past the tile's own edges (they belong to the following tiles): zeros, so that the reduce-add of the full
          // 128-row box leaves them as they are
#pragma unroll
          for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(mrow + (k << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (HAS_OUT) fence_proxy_async();
      mbar_arrive(bar_staged + 8 * (it % 6));
      if (gt == 0) E8_DBG(6, it);
      prev_it = it;
      prev_sw = sw;
      prev_b = b;
      prev_r0 = r0_cur;
      prev_nrec = nrec_cur;
    }
    if (prev_it >= 0) reduce_item(prev_it, prev_sw, prev_b, prev_r0, prev_nrec);
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == W_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host
int edge_projections(const float* send, int64_t send_bs, int64_t ns, int Bs, const float* rec, int64_t rec_bs, int64_t nr,
                     int Br, const float* w1, const float* b1, float* Ps, float* Pr, cudaStream_t st);  // tc2.cu

// the call shapes of this kernel: no edge output, or the edge tensor (dense batches) updated in place
bool tc_edge_rmw_supported(const NlamGraph* g, const float* edge, int64_t edge_bs, const float* edge_out, int B) {
  static int on = -1;
  if (on < 0) on = getenv("NLAM_TC_NO_RMW") ? 0 : 1;
  if (!on) return false;
  if (!edge_out) return true;
  return edge_out == edge && (B == 1 || edge_bs == (int64_t)g->n_edges * 64);
}

int tc_edge_rmw(const NlamGraph* g, const NlamMlp* edge_mlp, const float* send, int64_t send_bs, const float* rec,
                int64_t rec_bs, const float* edge, int64_t edge_bs, float* edge_out, float* aggr_out, int B, int flags,
                cudaStream_t st, float* ws, bool have_proj) {
  NLAM_REQUIRE(aligned16(edge) && aligned16(aggr_out) && edge_bs % 4 == 0 && aligned16(ws), NLAM_E_INVALID,
               "tc_edge_rmw: pointers / strides must be 16-byte aligned");
  NLAM_REQUIRE(!edge_out || edge_out == edge, NLAM_E_INVALID, "tc_edge_rmw: the edge output must alias the edge input");
  const int Bs = (send_bs == 0 || B == 1) ? 1 : B;
  const int Br = (rec_bs == 0 || B == 1) ? 1 : B;
  const int64_t ns = g->n_send, nr = g->n_rec;
  float* Ps = ws;
  float* Pr = ws + (size_t)Bs * ns * 64;
  const float* w1 = edge_mlp->w[0];  // (64, 192): columns [e | sender | receiver]
  // have_proj: ws already holds [P_s | P_r] of this edge MLP (written by the previous layer's node kernel, tc10.cu)
  int rc = have_proj ? NLAM_OK : edge_projections(send, send_bs, ns, Bs, rec, rec_bs, nr, Br, w1, edge_mlp->b[0], Ps, Pr, st);
  if (rc) return rc;

  CUtensorMap me, mw1, mw2, mps;
  const bool batched = edge_bs != 0 && B > 1;
  rc = make_map(&me, edge, 64, (uint64_t)g->n_edges, batched ? (uint64_t)B : 1, 64,
                batched ? (uint64_t)edge_bs : (uint64_t)g->n_edges * 64, 128, true);
  if (rc) return rc;
  rc = make_map(&mw1, w1, 64, 64, 1, 192, 0, 64, false);
  if (rc) return rc;
  rc = make_map(&mw2, edge_mlp->w[1], 64, 64, 1, 64, 0, 64, false);
  if (rc) return rc;
  rc = make_map(&mps, Ps, 64, (uint64_t)ns * Bs, 1, 64, 0, 1, false);
  if (rc) return rc;
  Edge8Params p;
  memset(&p, 0, sizeof(p));
  p.win_u = g->win_u;
  p.win_nu = g->win_nu;
  p.win_loc = g->win_loc;
  p.dst = g->dst;
  p.ps_rows = Bs > 1 ? (int)ns : 0;
  p.pr = Pr;
  p.pr_bs = Br > 1 ? (long long)g->n_rec * 64 : 0;
  p.b2 = edge_mlp->b[1];
  p.gamma = edge_mlp->ln_gamma;
  p.beta = edge_mlp->ln_beta;
  p.eps = edge_mlp->ln_eps;
  p.aggr = aggr_out;
  p.has_out = edge_out ? 1 : 0;
  p.e_batched = batched;
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
    NLAM_CUDA_OK(cudaFuncSetAttribute(tc_edge_rmw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e8::SMEM));
    NLAM_CUDA_OK(cudaFuncSetAttribute(tc_edge_rmw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e8::SMEM));
    attr_mask |= 1u << (dev & 31);
  }
  const long long n_work = (long long)p.n_tiles * p.B;
  NLAM_REQUIRE(n_work < (1LL << 30), NLAM_E_UNSUPPORTED, "tc_edge_rmw: too many work items");
  const int sms = num_sms();
  p.items_per_cta = (int)((n_work + sms - 1) / sms);
  static int epol = -1;
  if (epol < 0) {
    const char* e = getenv("NLAM_E8_EPOL");
    epol = e ? atoi(e) : 1;  // measured: 242 vs 254 us, 691 vs 708 MB read from DRAM
  }
  p.e_policy = epol;
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
    ProfScope ps("tc_edge_rmw_kernel", st, edge_algorithmic_bytes(g, B, send_bs, rec_bs, edge_bs, edge_out != nullptr, 64));
    if (edge_out)
      NLAM_CUDA_OK(launch_pdl(tc_edge_rmw_kernel<true>, grid, e8::THREADS, e8::SMEM, st, me, mw1, mw2, me, mps, p));
    else
      NLAM_CUDA_OK(launch_pdl(tc_edge_rmw_kernel<false>, grid, e8::THREADS, e8::SMEM, st, me, mw1, mw2, me, mps, p));
  }
  count_launch();
  if (dbg_on) {
    long long h[256];
    NLAM_CUDA_OK(cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
    NLAM_CUDA_OK(cudaStreamSynchronize(st));
    long long t0 = h[0];
    fprintf(stderr, "[nlam tc_edge_rmw timeline] grid=%d items/cta=%d (CTA 1, items 40..55; cycles rel. to the first)\n", grid,
            p.items_per_cta);
    fprintf(stderr, " it  ld_iss  g1_iss  g2_iss e1_start e1_done e2_start e2_done reduced red_beg       -       -  stored\n");
    for (int it = 0; it < 16; ++it) {
      fprintf(stderr, "%3d ", it + 40);
      for (int k = 0; k < 12; ++k) fprintf(stderr, "%7lld ", h[it * 16 + k] ? h[it * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

}  // namespace nlam
