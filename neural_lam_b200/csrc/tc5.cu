// Edge kernel v3 for H = 64 (general receiver-sorted edge sets, with or without edge update): the split first
// Linear of tc2.cu,   W1·[e; x_s; x_r] + b1 = W1e·e + (W1s·x)_src + (W1r·x + b1)_dst ,   restructured after the
// measurements of round 1:
//   * SENDER WINDOWS — a 128-edge tile of the mesh graph reads ~60-70 distinct senders, so the loaders gather
//     each distinct P_s row once per tile (graph handle: win_u / win_nu / win_loc) instead of one row per edge;
//     the TMA gather engine (~1 tile::gather4 per 100 cycles and per issuing lane) is no longer the bottleneck;
//   * 640 threads (96 registers each, no spills) instead of 896 (72 registers, spilling): one epilogue-1 group
//     and one epilogue-2 group of 8 warps (thread = row x 32 columns), packed fp32 arithmetic (FADD2/FFMA2),
//     SiLU through h = z/2 with W1e halved in place;
//   * the segmented sum over the tile's receivers is done by the epilogue-1 group (which has slack); the second
//     epilogue does bias + LayerNorm, stages the messages and writes e' = e + m in place over the e tile; a store
//     warp moves e' out with TMA; no group-wide barrier or TMA issue sits in the second epilogue's path;
//   * three tiles in flight: shared-memory stage = TMEM stage = tile mod 3; the MMA warp runs in uniform
//     control flow with one elected lane and sleeps while idle.
// Per tile and stage: e tile 32 KB (TMA; becomes e' in place and is stored by TMA) + sender window 32 KB
// (tile::gather4; later holds the messages for the segmented sum).
// Receiver term: (x_r·W1rᵀ + b1)[dst] read from global/L2 by epilogue 1 (rows of one CSR segment share it).
// Node projections P_s / P_r: tc_rowlinear_kernel (tc2.cu).
#include "tc_ptx.cuh"

namespace nlam {

namespace e5 {
constexpr int THREADS = 640;
constexpr int EPI = 256;
constexpr int LD_THREADS = 64;
constexpr int W_E1 = 8, W_MMA = 16, W_LD = 17, W_ST = 19;
constexpr int NS = 3;
constexpr uint32_t BLK = 16384;
constexpr uint32_t WBLK = 8192;
constexpr uint32_t OFF_W1 = 0;
constexpr uint32_t OFF_W2 = 2 * WBLK;
constexpr uint32_t OFF_ST = 4 * WBLK;  // stage s: [e0 e1 win0 win1]
constexpr uint32_t OFF_MISC = OFF_ST + NS * 4 * BLK;
constexpr uint32_t SMEM = OFF_MISC + 3072;  // 232448
}  // namespace e5

struct Edge3Params {
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
  long long* dbg;
};

#define E5_DBG(slot, it)                                                                  \
  do {                                                                                    \
    if (p.dbg && blockIdx.x == 0 && (it) < 16) p.dbg[(it) * 16 + (slot)] = clock64();     \
  } while (0)

__global__ void __launch_bounds__(e5::THREADS, 1)
tc_edge3_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut,
                const __grid_constant__ CUtensorMap tmPs, const Edge3Params p) {
  using namespace e5;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if ((sbase & 1023u) != 0) {
    if (tid == 0) printf("nlam tc_edge3: dynamic shared memory not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t mb = sbase + OFF_MISC;
  const uint32_t bar_w = mb + 0;
  const uint32_t bar_wscaled = mb + 8;    // W1e halved in place (256 arrivals)
  const uint32_t bar_full = mb + 16;      // [3] stage filled: 1 arrival + e tile + window bytes
  const uint32_t bar_free = mb + 40;      // [3] stage released: segmented sum done + e' store has read the tile (2 arrivals)
  const uint32_t bar_d1_full = mb + 64;   // [3]
  const uint32_t bar_hb_full = mb + 88;   // [3] 256 arrivals
  const uint32_t bar_d2_full = mb + 112;  // [3]
  const uint32_t bar_staged = mb + 136;   // [3] messages (and e') written to the stage (256 arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 192);
  int* lp = reinterpret_cast<int*>(smem + OFF_MISC + 256);          // [132] local CSR offsets of the tile being reduced
  float* sprm = reinterpret_cast<float*>(smem + OFF_MISC + 1024);  // b2 | gamma | beta

  if (warp == W_MMA) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_wscaled, EPI);
      for (int s = 0; s < NS; ++s) {
        mbar_init(bar_full + 8 * s, 1);
        mbar_init(bar_free + 8 * s, 2);
        mbar_init(bar_d1_full + 8 * s, 1);
        mbar_init(bar_hb_full + 8 * s, EPI);
        mbar_init(bar_d2_full + 8 * s, 1);
        mbar_init(bar_staged + 8 * s, EPI);
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
    if (p.has_out) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
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
  // TMEM columns: stage s: D at s*128 (first GEMM, then second), hidden at +64; LayerNorm scratch at 384
  const int n_work = p.n_tiles * p.B;
  int n_my = 0;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) ++n_my;

  if (warp == W_ST) {
    // =============================== e' stores + stage release ===============================
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int b = w / p.n_tiles, t = w - b * p.n_tiles;
        const int s = it % NS;
        mbar_wait(bar_staged + 8 * s, (uint32_t)((it / NS) & 1));  // e' = e + m written over the e tile
        if (p.has_out) {
          const int e0 = p.tile_e0[t];
          tma_store_3d(&tmOut, sbase + OFF_ST + s * 4 * BLK, 0, e0, b);
          tma_store_3d(&tmOut, sbase + OFF_ST + s * 4 * BLK + BLK, 32, e0, b);
          bulk_commit();
          bulk_wait_read0();
        }
        mbar_arrive(bar_free + 8 * s);
        E5_DBG(11, it);
      }
      if (p.has_out) bulk_wait0();
    }
  } else if (warp >= W_LD) {
    // =============================== loaders (2 warps) ===============================
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_keep = policy_evict_last();
    const int lw = warp - W_LD;
    if (lw == 0 && lane == 0) {
      mbar_expect_tx(bar_w, 4u * WBLK);
      for (int j = 0; j < 2; ++j) tma_load_2d(sbase + OFF_W1 + j * WBLK, &tmW1, bar_w, 32 * j, 0);
      for (int j = 0; j < 2; ++j) tma_load_2d(sbase + OFF_W2 + j * WBLK, &tmW2, bar_w, 32 * j, 0);
    }
    // lane l of loader warp lw issues gather4 number op = 32*lw + l: window rows 4*l.., column block lw; only the
    // groups the window has are issued.  Ids are prefetched one tile ahead.
    const int grp4 = lane, jb = lw;
    const bool issuer = true;
    int4 ids = make_int4(0, 0, 0, 0);
    int ngrp = 0;
    if ((int)blockIdx.x < n_work) {
      const int t0 = (int)blockIdx.x % p.n_tiles;
      ngrp = __ldg(p.win_nu + t0) >> 2;
      if (issuer && grp4 < ngrp) ids = __ldg(reinterpret_cast<const int4*>(p.win_u + (size_t)t0 * 128) + grp4);
    }
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      const int b = w / p.n_tiles, t = w - b * p.n_tiles;
      const int s = it % NS;
      const uint32_t full = bar_full + 8 * s;
      const uint32_t stg = sbase + OFF_ST + s * 4 * BLK;
      if (lw == 0) {
        if (lane == 0) {
          mbar_wait(bar_free + 8 * s, (uint32_t)(((it / NS) & 1) ^ 1));
          E5_DBG(0, it);
          mbar_expect_tx(full, 2u * BLK + (uint32_t)ngrp * 1024u);
          const int e0 = p.tile_e0[t];
          tma_load_3d(stg, &tmE, full, 0, e0, p.e_batched ? b : 0, pol_stream);
          tma_load_3d(stg + BLK, &tmE, full, 32, e0, p.e_batched ? b : 0, pol_stream);
        }
        __syncwarp();
      }
      named_bar_sync(12, LD_THREADS);
      if (issuer && grp4 < ngrp) {
        const int boff = p.ps_rows * b;
        tma_gather4(stg + (2 + jb) * BLK + grp4 * 512, &tmPs, full, 32 * jb, ids.x + boff, ids.y + boff, ids.z + boff,
                    ids.w + boff, pol_keep);
      }
      if (lw == 0 && lane == 0) {  // pull the e tile after next into L2
        const int w2 = w + 2 * (int)gridDim.x;
        if (w2 < n_work) {
          const int b2 = w2 / p.n_tiles, t2 = w2 - b2 * p.n_tiles;
          const int r2 = p.tile_e0[t2];
          tma_prefetch_3d(&tmE, 0, r2, p.e_batched ? b2 : 0);
          tma_prefetch_3d(&tmE, 32, r2, p.e_batched ? b2 : 0);
        }
      }
      const int wn = w + (int)gridDim.x;
      if (wn < n_work) {
        const int tn = wn % p.n_tiles;
        ngrp = __ldg(p.win_nu + tn) >> 2;
        if (issuer && grp4 < ngrp) ids = __ldg(reinterpret_cast<const int4*>(p.win_u + (size_t)tn * 128) + grp4);
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
    const uint64_t desc_st = umma_desc(sbase + OFF_ST);
    int g1 = 0, g2 = 0;
    uint32_t idle = 0;
    while (g2 < n_my) {
      bool progress = false;
      if (g1 < n_my && g1 <= g2 + 2) {
        const int s = g1 % NS;
        // a filled stage implies that tile g1-3 has left it, i.e. its accumulators were drained long ago
        if (mbar_test_u(bar_full + 8 * s, (uint32_t)((g1 / NS) & 1))) {
          tc_fence_after();
          if (lane == 0) E5_DBG(1, g1);
          const uint32_t dd = tmem_base + s * 128;
          const uint64_t a0 = desc_st + (uint64_t)((s * 4 * BLK) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32(dd, a0 + (uint64_t)((j * BLK) >> 4) + 2 * k, desc_w1 + (uint64_t)((j * WBLK) >> 4) + 2 * k, idesc,
                          (uint32_t)((j | k) != 0));
            umma_commit(bar_d1_full + 8 * s);
          }
          __syncwarp();
          ++g1;
          progress = true;
        }
      }
      if (g2 < g1) {
        const int s = g2 % NS;
        if (mbar_test_u(bar_hb_full + 8 * s, (uint32_t)((g2 / NS) & 1))) {
          tc_fence_after();
          if (lane == 0) E5_DBG(2, g2);
          const uint32_t dd = tmem_base + s * 128;
          const uint32_t ht = dd + 64;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_ts(dd, ht + (uint32_t)(j * 32 + k * 8), desc_w2 + (uint64_t)((j * WBLK) >> 4) + 2 * k, idesc,
                             (uint32_t)((j | k) != 0));
            umma_commit(bar_d2_full + 8 * s);
          }
          __syncwarp();
          ++g2;
          progress = true;
        }
      }
      if (progress) idle = 0;
      else if (__nanosleep(40), ++idle > (1u << 24)) {
        if (lane == 0) printf("nlam tc_edge3: MMA issuer timeout (block %d g1 %d g2 %d)\n", blockIdx.x, g1, g2);
        __trap();
      }
    }
  } else if (warp >= W_E1) {
    // =============================== epilogue 1 (+ segmented sum of the previous tile) ===============================
    const bool lead = warp == W_E1;
    const int gt = tid - W_E1 * 32;  // 0..255
    const int q = warp & 3;
    const int half = (warp - W_E1) >> 2;
    const int row = q * 32 + lane;
    const int c0 = half * 32;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    // SiLU(z) = h + h*tanh(h), h = z/2: W1e is halved in place once (exact); the gathered node terms are halved
    // in the FMA that adds them
    {
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
    // per-tile indices are loaded one tile ahead: a dependent chain of L2 round trips must not sit in front of
    // the barrier waits (there is no L1 next to 227 KB of shared memory)
    int dst_next = 0, loc_next = 0, lp_next = 0, r0_next = 0, nrec_next = 0;
    if ((int)blockIdx.x < n_work) {
      const int t0 = (int)blockIdx.x % p.n_tiles;
      const int4 m0 = __ldg(p.tile_meta + t0);
      r0_next = m0.z;
      nrec_next = m0.w;
      dst_next = (m0.x + row < p.n_edges) ? __ldg(p.dst + m0.x + row) : 0;
      loc_next = __ldg(p.win_loc + (size_t)t0 * 128 + row);
      lp_next = (gt <= m0.w) ? __ldg(p.rowptr + m0.z + gt) - m0.x : 0;
    }
    // segmented sum of tile `itr` (stage sr, batch br, tile meta mr): thread = (float4 column group, receiver group)
    auto reduce_tile = [&](int itr, int br, int r0, int nrec, int lp_val) {
      const int sr = itr % NS;
      if (gt <= nrec) lp[gt] = lp_val;  // CSR offsets of the tile's receivers, loaded a tile ahead
      if (lead) mbar_wait(bar_staged + 8 * sr, (uint32_t)((itr / NS) & 1));
      named_bar_sync(1, EPI);  // messages staged, offsets visible
      if (gt == 0) E5_DBG(12, itr);
      const int cg = gt & 15, g = gt >> 4;
      const uint8_t* mbase = smem + OFF_ST + sr * 4 * BLK + (2 + (cg >> 3)) * BLK;
      const int chq = cg & 7;
      for (int j = g; j < nrec; j += EPI / 16) {
        const int k0 = lp[j], k1 = lp[j + 1];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int k = k0;
        for (; k + 4 <= k1; k += 4) {  // four independent loads in flight
          const float4 a = *reinterpret_cast<const float4*>(mbase + swz(k, chq));
          const float4 b4 = *reinterpret_cast<const float4*>(mbase + swz(k + 1, chq));
          const float4 c = *reinterpret_cast<const float4*>(mbase + swz(k + 2, chq));
          const float4 d4 = *reinterpret_cast<const float4*>(mbase + swz(k + 3, chq));
          acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
          acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
          acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
          acc.x += d4.x; acc.y += d4.y; acc.z += d4.z; acc.w += d4.w;
        }
        for (; k < k1; ++k) {
          const float4 m4 = *reinterpret_cast<const float4*>(mbase + swz(k, chq));
          acc.x += m4.x;
          acc.y += m4.y;
          acc.z += m4.z;
          acc.w += m4.w;
        }
        if (p.mean) {
          const float sc = 1.0f / (float)max(k1 - k0, 1);
          acc.x *= sc;
          acc.y *= sc;
          acc.z *= sc;
          acc.w *= sc;
        }
        *reinterpret_cast<float4*>(p.aggr + ((long long)br * p.n_rec + r0 + j) * 64 + cg * 4) = acc;
      }
      named_bar_sync(1, EPI);  // every access of the stage (and of lp) is done
      if (gt == 0) {
        mbar_arrive(bar_free + 8 * sr);
        E5_DBG(7, itr);
      }
    };
    int it = 0;
    int b_prev = 0, lp_prev = 0, r0_prev = 0, nrec_prev = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      const int b = w / p.n_tiles;
      const int s = it % NS;
      const int my_dst = dst_next, loc = loc_next;
      const int lp_cur = lp_next, r0_cur = r0_next, nrec_cur = nrec_next;
      const int wn = w + (int)gridDim.x;
      // receiver projection row of this edge (rows of one CSR segment share it: the lanes' loads coalesce)
      const float4* prow = reinterpret_cast<const float4*>(p.pr + (long long)b * p.pr_bs + (long long)my_dst * 64 + c0);
      float4 pr[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) pr[k] = __ldg(prow + k);
      int4 mn = make_int4(0, 0, 0, 0);
      if (wn < n_work) {
        const int tn = wn % p.n_tiles;
        mn = __ldg(p.tile_meta + tn);
        loc_next = __ldg(p.win_loc + (size_t)tn * 128 + row);
      }
      if (lead) {
        mbar_wait(bar_full + 8 * s, (uint32_t)((it / NS) & 1));  // sender window visible
        mbar_wait(bar_d1_full + 8 * s, (uint32_t)((it / NS) & 1));
      }
      named_bar_sync(1, EPI);
      tc_fence_after();
      if (lead && lane == 0) E5_DBG(3, it);
      const uint8_t* ps = smem + OFF_ST + s * 4 * BLK + (2 + half) * BLK + loc * 128;
      const int rx = loc & 7;
      const uint32_t d1 = tmem_base + s * 128 + t_lane + c0;
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
      tc_fence_before();
      mbar_arrive(bar_hb_full + 8 * s);
      if (lead && lane == 0) E5_DBG(4, it);
      if (wn < n_work) {  // the next tile's meta has landed long ago: these loads do not stall
        r0_next = mn.z;
        nrec_next = mn.w;
        dst_next = (mn.x + row < p.n_edges) ? __ldg(p.dst + mn.x + row) : 0;
        lp_next = (gt <= mn.w) ? __ldg(p.rowptr + mn.z + gt) - mn.x : 0;
      }
      // while the tensor core and the second epilogue work on this tile, sum the previous one
      if (it > 0) reduce_tile(it - 1, b_prev, r0_prev, nrec_prev, lp_prev);
      b_prev = b;
      r0_prev = r0_cur;
      nrec_prev = nrec_cur;
      lp_prev = lp_cur;
    }
    if (it > 0) reduce_tile(it - 1, b_prev, r0_prev, nrec_prev, lp_prev);
  } else {
    // =============================== epilogue 2: bias, LayerNorm -> messages, e' = e + m ===============================
    const int q = warp & 3;
    const int half = warp >> 2;
    const int row = q * 32 + lane;
    const int c0 = half * 32;
    const int rx = row & 7;
    const uint32_t rsw = (uint32_t)(row * 128);
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    const int pbar = 4 + q;
    const uint32_t ln_col = tmem_base + 384 + t_lane;
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      const int s = it % NS;
      if (warp == 0) mbar_wait(bar_d2_full + 8 * s, (uint32_t)((it / NS) & 1));
      named_bar_sync(2, EPI);
      tc_fence_after();
      if (tid == 0) E5_DBG(5, it);
      float vf[32];
      tmem_ld32(tmem_base + s * 128 + t_lane + c0, vf);
      if (tid == 0) E5_DBG(8, it);
      float2 v[16];
      float2 sm2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 bb = *reinterpret_cast<const float4*>(sprm + c0 + 4 * k);
        v[2 * k] = add2(make_float2(vf[4 * k], vf[4 * k + 1]), make_float2(bb.x, bb.y));
        v[2 * k + 1] = add2(make_float2(vf[4 * k + 2], vf[4 * k + 3]), make_float2(bb.z, bb.w));
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        sm2 = add2(sm2, v[i]);
        sq2 = fma2(v[i], v[i], sq2);
      }
      // the two column halves of a row exchange (sum, sum of squares) through spare TMEM columns of the row's lane;
      // scratch double-buffered by tile parity: one 64-thread barrier per tile
      const uint32_t scr = ln_col + 4 * (it & 1);
      tmem_st2(scr + 2 * half, sm2.x + sm2.y, sq2.x + sq2.y);
      tc_fence_before();
      named_bar_sync(pbar, 64);
      tc_fence_after();
      float st4[4];
      tmem_ld4(scr, st4);
      if (tid == 0) E5_DBG(9, it);
      const float mu = (st4[0] + st4[2]) * (1.0f / 64.0f);
      const float ex2 = (st4[1] + st4[3]) * (1.0f / 64.0f);
      const float rstd = rsqrtf(fmaxf(ex2 - mu * mu, 0.f) + p.eps);
      const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mu * rstd, -mu * rstd);
      uint8_t* mrow = smem + OFF_ST + s * 4 * BLK + (2 + half) * BLK + rsw;  // the sender window was consumed by epilogue 1
      uint8_t* erow = smem + OFF_ST + s * 4 * BLK + half * BLK + rsw;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 g4 = *reinterpret_cast<const float4*>(sprm + 64 + c0 + 4 * k);
        const float4 b4 = *reinterpret_cast<const float4*>(sprm + 128 + c0 + 4 * k);
        const float2 m0 = fma2(fma2(v[2 * k], rs2, nm2), make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
        const float2 m1 = fma2(fma2(v[2 * k + 1], rs2, nm2), make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
        *reinterpret_cast<float4*>(mrow + ((k ^ rx) << 4)) = make_float4(m0.x, m0.y, m1.x, m1.y);
        if (p.has_out) {  // e' = e + m in place over the e tile (stored by the store warp)
          float4* ptr = reinterpret_cast<float4*>(erow + ((k ^ rx) << 4));
          const float4 r = *ptr;
          const float2 o0 = add2(make_float2(r.x, r.y), m0);
          const float2 o1 = add2(make_float2(r.z, r.w), m1);
          *ptr = make_float4(o0.x, o0.y, o1.x, o1.y);
        }
      }
      if (p.has_out) fence_proxy_async();
      mbar_arrive(bar_staged + 8 * s);
      if (tid == 0) E5_DBG(6, it);
    }
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

int tc_edge3(const NlamGraph* g, const NlamMlp* edge_mlp, const float* send, int64_t send_bs, const float* rec,
             int64_t rec_bs, const float* edge, int64_t edge_bs, float* edge_out, float* aggr_out, int B, int flags,
             cudaStream_t st, float* ws, bool have_proj) {
  NLAM_REQUIRE(aligned16(edge) && aligned16(aggr_out) && (!edge_out || aligned16(edge_out)) && edge_bs % 4 == 0 &&
                   aligned16(ws),
               NLAM_E_INVALID, "tc_edge3: pointers / strides must be 16-byte aligned");
  const int Bs = (send_bs == 0 || B == 1) ? 1 : B;
  const int Br = (rec_bs == 0 || B == 1) ? 1 : B;
  const int64_t ns = g->n_send, nr = g->n_rec;
  float* Ps = ws;
  float* Pr = ws + (size_t)Bs * ns * 64;
  const float* w1 = edge_mlp->w[0];  // (64, 192): columns [e | sender | receiver]
  // have_proj: ws already holds [P_s | P_r] of this edge MLP (written by the previous layer's node kernel, tc10.cu)
  int rc = have_proj ? NLAM_OK : edge_projections(send, send_bs, ns, Bs, rec, rec_bs, nr, Br, w1, edge_mlp->b[0], Ps, Pr, st);
  if (rc) return rc;

  CUtensorMap me, mw1, mw2, mo, mps;
  const bool batched = edge_bs != 0 && B > 1;
  rc = make_map(&me, edge, 64, (uint64_t)g->n_edges, batched ? (uint64_t)B : 1, 64,
                batched ? (uint64_t)edge_bs : (uint64_t)g->n_edges * 64, 128, true);
  if (rc) return rc;
  rc = make_map(&mw1, w1, 64, 64, 1, 192, 0, 64, false);
  if (rc) return rc;
  rc = make_map(&mw2, edge_mlp->w[1], 64, 64, 1, 64, 0, 64, false);
  if (rc) return rc;
  if (edge_out) {
    rc = make_map(&mo, edge_out, 64, (uint64_t)g->n_edges, (uint64_t)B, 64, (uint64_t)g->n_edges * 64, 128, true);
    if (rc) return rc;
  } else {
    mo = me;
  }
  rc = make_map(&mps, Ps, 64, (uint64_t)ns * Bs, 1, 64, 0, 1, false);
  if (rc) return rc;
  Edge3Params p;
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
    NLAM_CUDA_OK(cudaFuncSetAttribute(tc_edge3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e5::SMEM));
    attr_mask |= 1u << (dev & 31);
  }
  const long long n_work = (long long)p.n_tiles * p.B;
  NLAM_REQUIRE(n_work < (1LL << 30), NLAM_E_UNSUPPORTED, "tc_edge3: too many work items");
  const int grid = (int)std::min<long long>(n_work, num_sms());
  static long long* dbg_buf = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) dbg_on = getenv("NLAM_TC_TIMELINE") ? 1 : 0;
  if (dbg_on) {
    if (!dbg_buf) NLAM_CUDA_OK(cudaMalloc(&dbg_buf, 256 * sizeof(long long)));
    NLAM_CUDA_OK(cudaMemsetAsync(dbg_buf, 0, 256 * sizeof(long long), st));
    p.dbg = dbg_buf;
  }
  {
    // the node-side terms arrive as projections (same row counts as the raw node rows)
    ProfScope ps("tc_edge3_kernel", st, edge_algorithmic_bytes(g, B, send_bs, rec_bs, edge_bs, edge_out != nullptr, 64));
    NLAM_CUDA_OK(launch_pdl(tc_edge3_kernel, grid, e5::THREADS, e5::SMEM, st, me, mw1, mw2, mo, mps, p));
  }
  count_launch();
  if (dbg_on) {
    long long h[256];
    NLAM_CUDA_OK(cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
    NLAM_CUDA_OK(cudaStreamSynchronize(st));
    long long t0 = h[0];
    fprintf(stderr, "[nlam tc_edge3 timeline] grid=%d work=%lld (cycles rel. to first load)\n", grid, n_work);
    fprintf(stderr, " it  ld_iss  g1_iss  g2_iss e1_start e1_done e2_start e2_done reduced   e2_ld   e2_ln e2_stgd e2_wtrd red_beg\n");
    for (int it = 0; it < 14; ++it) {
      fprintf(stderr, "%3d ", it);
      for (int k = 0; k < 13; ++k) fprintf(stderr, "%7lld ", h[it * 16 + k] ? h[it * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

}  // namespace nlam
