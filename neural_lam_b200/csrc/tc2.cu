// Node projection kernel of the split first Linear + the FIRST split edge kernel (kept as a fallback).
//
// The first Linear of the edge MLP acts on [e | x_src | x_dst] (reference gnn_layers.py:168-172).  It is
// linear, so   W1·[e; x_s; x_r] + b1 = W1e·e + (W1s·x)_src + (W1r·x + b1)_dst .
// The two node-side terms are computed once per NODE by `tc_rowlinear_kernel` (P_s = x_send·W1sᵀ,
// P_r = x_rec·W1rᵀ + b1; both problems of a call in ONE launch) instead of once per EDGE, and the edge kernel only
// runs the K=64 GEMM e·W1eᵀ on the tensor cores and adds the gathered projections in its first epilogue.
// Summation order differs from the reference (three TF32 GEMMs summed in fp32) — inside the stated TF32 tolerance.
//
// The edge kernel in use is tc_edge3_kernel (tc5.cu); `tc_edge2_kernel` below is its predecessor (896 threads at
// 72 registers — it spills —, two alternating groups per epilogue, segmented sum inside epilogue 2), reachable
// with NLAM_TC_NO_EDGE3=1 and covered by tests/test_tc_kernels.py.  It shares the sender windows of the graph
// handle (one tile::gather4 per 4 DISTINCT sender rows of a tile) and the packed-fp32 epilogue arithmetic.
//
// tc_edge2_kernel: 896 threads, 1 CTA/SM, persistent over (batch, tile):
//   warps 0-7   epilogue-2 group 0 (tiles 0,2,4,..)   } D2 -> bias, LayerNorm, messages staged in the
//   warps 8-15  epilogue-2 group 1 (tiles 1,3,5,..)   } tile's window buffer, e' = e + m in place + TMA
//                                                       store, CSR segmented sum -> aggr
//   warps 16-23 epilogue-1, two groups alternating tiles: D1 + P_s[window row] (smem) + P_r[dst] (global) ->
//               SiLU -> hidden in TMEM
//   warp 24     tcgen05.mma issue (GEMM1 SS form K=64, GEMM2 TS form, A = hidden in TMEM)
//   warps 25-27 loaders: weights once; per tile the e tile (TMA) + the window's gather4 operations
// Shared memory: W1e 16 KB | W2 16 KB | 3 stages x (e 32 KB + window 32 KB) | misc = 227 KB.
// TMEM: 2 stages x (D 64 cols [D1, later D2] + hidden 64 cols) + LayerNorm scratch.
#include "tc_ptx.cuh"

namespace nlam {

namespace e2 {
constexpr int THREADS = 896;
constexpr int LD_THREADS = 96;   // 4 loader warps: a warp issues ~1 TMA operation per 100 cycles, so the
                                 // 64 gather4 of a tile are spread over 4 warps x 16 lanes
constexpr int G2_THREADS = 256;  // per epilogue-2 group
constexpr int E1_THREADS = 128;
constexpr int W_E1 = 16, W_MMA = 24, W_TMA = 25;  // E1: two groups of 4 warps (16-19, 20-23); loaders 25-27
constexpr int BM = 128;
constexpr int NS = 3;  // shared-memory stages
constexpr uint32_t BLK = 16384;
constexpr uint32_t WBLK = 8192;
constexpr uint32_t OFF_W1 = 0;
constexpr uint32_t OFF_W2 = 2 * WBLK;
constexpr uint32_t OFF_ST = 4 * WBLK;           // stage s at OFF_ST + s*4*BLK: [e0 e1 ps0 ps1]
constexpr uint32_t OFF_MISC = OFF_ST + NS * 4 * BLK;
constexpr uint32_t SMEM = OFF_MISC + 3072;       // 232448
}  // namespace e2

struct Edge2Params {
  const int32_t* win_u;     // per tile: 128 distinct sender ids of its edge window
  const int32_t* win_nu;    // per tile: window rows (multiple of 4)
  const uint8_t* win_loc;   // per tile: window row of each of the 128 edge rows
  const int32_t* dst;       // CSR-ordered receiver ids
  int ps_rows;              // rows of one batch in the P_s gather map (0: batch-broadcast)
  const float* pr;          // P_r (B_r, n_rec, 64)
  long long pr_bs;          // batch stride of P_r in floats (0: broadcast)
  const float* b2;
  const float* gamma;
  const float* beta;
  float eps;
  float* aggr;              // (B, n_rec, 64)
  int has_out;              // edge update: e' = e + m stored through tmOut
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

#define E2_DBG(slot, it)                                                        \
  do {                                                                          \
    if (p.dbg && blockIdx.x == 0 && (it) < 16) p.dbg[(it) * 16 + (slot)] = clock64(); \
  } while (0)

__global__ void __launch_bounds__(e2::THREADS, 1)
tc_edge2_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut,
                const __grid_constant__ CUtensorMap tmPs, const Edge2Params p) {
  using namespace e2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  if ((sbase & 1023u) != 0) {
    if (tid == 0) printf("nlam tc_edge2: dynamic shared memory not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t mb = sbase + OFF_MISC;
  const uint32_t bar_w = mb + 0;
  const uint32_t bar_full = mb + 8;       // [3] stage filled (1 arrival + 64 KB of TMA transactions)
  const uint32_t bar_epi_done = mb + 32;  // [3] stage released by the second epilogue
  const uint32_t bar_d1_full = mb + 56;   // [2]
  const uint32_t bar_hb_full = mb + 72;   // [2] hidden written (128 arrivals)
  const uint32_t bar_d2_full = mb + 88;   // [2]
  const uint32_t bar_d_free = mb + 104;   // [2] D2 drained into registers (256 arrivals)
  const uint32_t bar_wscaled = mb + 1192;  // W1e halved in place (256 arrivals); after the CSR offset arrays
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 120);
  int* lp_all = reinterpret_cast<int*>(smem + OFF_MISC + 128);  // [2 groups][132] local CSR offsets
  float* sprm = reinterpret_cast<float*>(smem + OFF_MISC + 2048);  // b2 | gamma | beta (64 floats each):
                                                                      // there is no L1 next to 227 KB of smem,
                                                                      // so per-tile __ldg of constants costs an
                                                                      // L2 round trip each

  if (warp == W_MMA) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_wscaled, 2 * E1_THREADS);
      for (int s = 0; s < NS; ++s) {
        mbar_init(bar_full + 8 * s, 1);
        mbar_init(bar_epi_done + 8 * s, 1);
      }
      for (int t = 0; t < 2; ++t) {
        mbar_init(bar_d1_full + 8 * t, 1);
        mbar_init(bar_hb_full + 8 * t, E1_THREADS);
        mbar_init(bar_d2_full + 8 * t, 1);
        mbar_init(bar_d_free + 8 * t, G2_THREADS);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == W_TMA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmE) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmPs) : "memory");
    if (p.has_out) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
  }
  if (tid < 64) {
    sprm[tid] = p.b2[tid];
    sprm[64 + tid] = p.gamma[tid];
    sprm[128 + tid] = p.beta[tid];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int n_work = p.n_tiles * p.B;

  if (warp >= W_TMA) {
    // =============================== loaders (4 warps) ===============================
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_keep = policy_evict_last();
    const int lw = warp - W_TMA;  // 0..3
    if (lw == 0 && lane == 0) {
      mbar_expect_tx(bar_w, 4u * WBLK);
      for (int j = 0; j < 2; ++j) tma_load_2d(sbase + OFF_W1 + j * WBLK, &tmW1, bar_w, 32 * j, 0);
      for (int j = 0; j < 2; ++j) tma_load_2d(sbase + OFF_W2 + j * WBLK, &tmW2, bar_w, 32 * j, 0);
    }
    // lanes 0-15 of loader warp lw issue gather4 number op = lw*16 + lane: row group op&31 (window rows
    // 4*grp..4*grp+3), column block op>>5.  Sender ids are prefetched one tile ahead.
    const int op = lw * 22 + lane;
    const int grp4 = op & 31, jb = (op >> 5) & 1;
    const bool issuer = lane < 22 && op < 64;
    // the window's distinct senders (prefetched one tile ahead): ngrp groups of 4 rows
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
      const uint32_t sph = (uint32_t)((it / NS) & 1);
      const uint32_t full = bar_full + 8 * s;
      const uint32_t stg = sbase + OFF_ST + s * 4 * BLK;
      if (lw == 0) {
        if (lane == 0) {
          mbar_wait(bar_epi_done + 8 * s, sph ^ 1);  // tile it-3 released the stage
          E2_DBG(0, it);
          mbar_expect_tx(full, 2u * BLK + (uint32_t)ngrp * 1024u);
          const int e0 = p.tile_e0[t];
          tma_load_3d(stg, &tmE, full, 0, e0, p.e_batched ? b : 0, pol_stream);
          tma_load_3d(stg + BLK, &tmE, full, 32, e0, p.e_batched ? b : 0, pol_stream);
        }
        __syncwarp();
      }
      named_bar_sync(12, LD_THREADS);  // stage free + transaction count armed
      if (issuer && grp4 < ngrp) {
        const int boff = p.ps_rows * b;
        tma_gather4(stg + (2 + jb) * BLK + grp4 * 512, &tmPs, full, 32 * jb, ids.x + boff, ids.y + boff, ids.z + boff,
                    ids.w + boff, pol_keep);
      }
      {
        const int wn = w + (int)gridDim.x;
        if (wn < n_work) {
          const int tn = wn % p.n_tiles;
          ngrp = __ldg(p.win_nu + tn) >> 2;
          if (issuer && grp4 < ngrp) ids = __ldg(reinterpret_cast<const int4*>(p.win_u + (size_t)tn * 128) + grp4);
        }
      }
      if (lw == 0 && lane == 0) {
        const int w2 = w + 2 * (int)gridDim.x;
        if (w2 < n_work) {
          const int b2 = w2 / p.n_tiles, t2 = w2 - b2 * p.n_tiles;
          const int r2 = p.tile_e0[t2];
          tma_prefetch_3d(&tmE, 0, r2, p.e_batched ? b2 : 0);
          tma_prefetch_3d(&tmE, 32, r2, p.e_batched ? b2 : 0);
        }
      }
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(BM, 64);
      int n_my = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) ++n_my;
      mbar_wait(bar_w, 0);
      mbar_wait(bar_wscaled, 0);
      const uint64_t desc_w1 = umma_desc(sbase + OFF_W1);
      const uint64_t desc_w2 = umma_desc(sbase + OFF_W2);
      const uint64_t desc_st = umma_desc(sbase + OFF_ST);
      int g1 = 0, g2 = 0;
      uint32_t idle = 0;
      while (g2 < n_my) {
        bool progress = false;
        if (g1 < n_my && g1 <= g2 + 1) {
          const int it = g1, ts = it & 1, s = it % NS;
          // operands landed, and epilogue 2 of tile it-2 has drained this TMEM stage
          if (mbar_test(bar_full + 8 * s, (uint32_t)((it / NS) & 1)) &&
              mbar_test(bar_d_free + 8 * ts, (uint32_t)(((it >> 1) & 1) ^ 1))) {
            tc_fence_after();
            E2_DBG(3, it);
            const uint32_t d = tmem_base + ts * 128;
            const uint64_t a0 = desc_st + (uint64_t)((s * 4 * BLK) >> 4);
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32(d, a0 + (uint64_t)((j * BLK) >> 4) + 2 * k, desc_w1 + (uint64_t)((j * WBLK) >> 4) + 2 * k, idesc,
                          (uint32_t)((j | k) != 0));
            umma_commit(bar_d1_full + 8 * ts);
            ++g1;
            progress = true;
          }
        }
        if (g2 < g1) {
          const int it = g2, ts = it & 1;
          if (mbar_test(bar_hb_full + 8 * ts, (uint32_t)((it >> 1) & 1))) {
            tc_fence_after();
            E2_DBG(4, it);
            const uint32_t d = tmem_base + ts * 128;  // D2 overwrites D1 (consumed by epilogue 1)
            const uint32_t ht = d + 64;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_ts(d, ht + (uint32_t)(j * 32 + k * 8), desc_w2 + (uint64_t)((j * WBLK) >> 4) + 2 * k, idesc,
                             (uint32_t)((j | k) != 0));
            umma_commit(bar_d2_full + 8 * ts);
            ++g2;
            progress = true;
          }
        }
        if (progress) idle = 0;
        else if (__nanosleep(40), ++idle > (1u << 24)) {  // idle polling must not take issue slots from the epilogues
          printf("nlam tc_edge2: MMA issuer timeout (block %d g1 %d g2 %d)\n", blockIdx.x, g1, g2);
          __trap();
        }
      }
    }
  } else if (warp >= W_E1) {
    // =============================== epilogue 1 ===============================
    // two groups of 4 warps alternate tiles (group = TMEM stage), like the second epilogue
    const int g1g = (warp - W_E1) >> 2;
    const bool g1lead = ((warp - W_E1) & 3) == 0;
    const int g1bar = g1g ? 13 : 1;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    const int stride1 = 2 * (int)gridDim.x;
    const int w_first1 = (int)blockIdx.x + g1g * (int)gridDim.x;
    // SiLU(z) = h + h*tanh(h), h = z/2: W1e is halved in place once (exact); the gathered node terms are halved
    // in the FMA that adds them
    {
      mbar_wait(bar_w, 0);
      float4* wq = reinterpret_cast<float4*>(smem + OFF_W1) + (tid - W_E1 * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // 16 KB = 1024 float4 over 256 threads
        float4 x = wq[i * 2 * E1_THREADS];
        x.x *= 0.5f;
        x.y *= 0.5f;
        x.z *= 0.5f;
        x.w *= 0.5f;
        wq[i * 2 * E1_THREADS] = x;
      }
      fence_proxy_async();
      mbar_arrive(bar_wscaled);
    }
    const float2 half2 = make_float2(0.5f, 0.5f);
    int dst_next = 0, loc_next = 0;
    if (w_first1 < n_work) {
      const int t0 = w_first1 % p.n_tiles;
      const int e0 = p.tile_e0[t0];
      dst_next = (e0 + row < p.n_edges) ? __ldg(p.dst + e0 + row) : 0;
      loc_next = __ldg(p.win_loc + (size_t)t0 * 128 + row);
    }
    int it = g1g;
    for (int w = w_first1; w < n_work; w += stride1, it += 2) {
      const int b = w / p.n_tiles;
      const int ts = it & 1, s = it % NS;
      const int my_dst = dst_next, loc = loc_next;
      const int wn = w + stride1;
      if (wn < n_work) {
        const int tn = wn % p.n_tiles;
        const int e0n = p.tile_e0[tn];
        dst_next = (e0n + row < p.n_edges) ? __ldg(p.dst + e0n + row) : 0;
        loc_next = __ldg(p.win_loc + (size_t)tn * 128 + row);
      }
      // receiver projection row of this edge: rows of one CSR segment share it, so the lanes' loads
      // coalesce to one request per distinct receiver
      const float4* prow = reinterpret_cast<const float4*>(p.pr + (long long)b * p.pr_bs + (long long)my_dst * 64);
      float4 pr_cur[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) pr_cur[k] = __ldg(prow + k);
      if (g1lead) {
        mbar_wait(bar_full + 8 * s, (uint32_t)((it / NS) & 1));  // P_s rows of this tile visible
        mbar_wait(bar_d1_full + 8 * ts, (uint32_t)((it >> 1) & 1));
      }
      named_bar_sync(g1bar, E1_THREADS);
      tc_fence_after();
      if (g1lead && lane == 0) E2_DBG(6, it);
      const uint8_t* ps = smem + OFF_ST + s * 4 * BLK + 2 * BLK + loc * 128;  // this edge's row of the sender window
      const int rx = loc & 7;
      const uint32_t d1 = tmem_base + ts * 128 + t_lane;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 pr_nxt[4];
        if (c < 3) {
#pragma unroll
          for (int k = 0; k < 4; ++k) pr_nxt[k] = __ldg(prow + 4 * (c + 1) + k);
        }
        float v[16];
        tmem_ld16(d1 + c * 16, v);
        const uint8_t* psb = ps + (c >> 1) * BLK;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 s4 = *reinterpret_cast<const float4*>(psb + ((((c & 1) * 4 + k) ^ rx) << 4));
          const float2 h0 = fma2(add2(make_float2(s4.x, s4.y), make_float2(pr_cur[k].x, pr_cur[k].y)), half2,
                                 make_float2(v[4 * k + 0], v[4 * k + 1]));
          const float2 h1 = fma2(add2(make_float2(s4.z, s4.w), make_float2(pr_cur[k].z, pr_cur[k].w)), half2,
                                 make_float2(v[4 * k + 2], v[4 * k + 3]));
          const float2 o0 = fma2(h0, make_float2(tanh_fast(h0.x), tanh_fast(h0.y)), h0);
          const float2 o1 = fma2(h1, make_float2(tanh_fast(h1.x), tanh_fast(h1.y)), h1);
          v[4 * k + 0] = o0.x;
          v[4 * k + 1] = o0.y;
          v[4 * k + 2] = o1.x;
          v[4 * k + 3] = o1.y;
        }
        tmem_st16(d1 + 64 + c * 16, v);
        if (c < 3) {
#pragma unroll
          for (int k = 0; k < 4; ++k) pr_cur[k] = pr_nxt[k];
        }
      }
      tc_fence_before();
      mbar_arrive(bar_hb_full + 8 * ts);
      if (g1lead && lane == 0) E2_DBG(7, it);
    }
  } else {
    // =============================== epilogue 2 (two groups, alternating tiles) ===============================
    const int grp = warp >> 3;              // also the TMEM stage this group works on
    const int gt = tid - grp * G2_THREADS;  // thread index inside the group
    const int gw = warp & 7;
    const int q = gw & 3;
    const int half = gw >> 2;
    const int row = q * 32 + lane;
    const int c0 = half * 32;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    const uint32_t rsw = (uint32_t)(row * 128);
    const int rx = row & 7;
    const int gbar = 2 + grp;           // group barrier
    const int pbar = 4 + 4 * grp + q;   // barrier of the two warps sharing this lane quarter
    int* lp = lp_all + grp * 132;
    const uint32_t ln_col = tmem_base + 256 + 8 * grp + t_lane;
    const int stride = 2 * (int)gridDim.x;

    int4 meta = make_int4(0, 0, 0, 0), meta_n = meta;
    int lp_val = 0;
    const int w_first = (int)blockIdx.x + grp * (int)gridDim.x;
    if (w_first < n_work) {
      meta = __ldg(p.tile_meta + (w_first % p.n_tiles));
      meta_n = meta;
      if (w_first + stride < n_work) meta_n = __ldg(p.tile_meta + ((w_first + stride) % p.n_tiles));
      if (gt <= meta.w) lp_val = __ldg(p.rowptr + meta.z + gt) - meta.x;
    }
    int it = grp;
    for (int w = w_first; w < n_work; w += stride, it += 2) {
      const int b = w / p.n_tiles;
      const int s = it % NS;
      const uint32_t tph = (uint32_t)((it >> 1) & 1);
      const int row0 = meta.x, r0 = meta.z, nrec = meta.w;
      const int wn = w + stride, wnn = wn + stride;
      int4 meta_nn = meta_n;
      if (wnn < n_work) meta_nn = __ldg(p.tile_meta + (wnn % p.n_tiles));
      if (gt <= nrec) lp[gt] = lp_val;  // previous tile of this group has passed its final barrier
      if (wn < n_work && gt <= meta_n.w) lp_val = __ldg(p.rowptr + meta_n.z + gt) - meta_n.x;

      if (gt == 0) E2_DBG(5, it);
      if (gw == 0) mbar_wait(bar_d2_full + 8 * grp, tph);
      named_bar_sync(gbar, G2_THREADS);
      tc_fence_after();
      if (gt == 0) E2_DBG(8, it);
      float vf[32];
      tmem_ld32(tmem_base + grp * 128 + t_lane + c0, vf);
      tc_fence_before();
      mbar_arrive(bar_d_free + 8 * grp);  // the TMEM stage may take the next tile's first GEMM
      float2 v[16];
      {
        // bias, then LayerNorm over 64 columns held by two threads (column halves): partial (sum, sum of
        // squares) parked in spare TMEM columns of the row's lane, one 64-thread barrier, read both back
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
        tmem_st2(ln_col + 2 * half, sm2.x + sm2.y, sq2.x + sq2.y);
        tc_fence_before();
        named_bar_sync(pbar, 64);
        tc_fence_after();
        float st4[4];
        tmem_ld4(ln_col, st4);
        const float mu = (st4[0] + st4[2]) * (1.0f / 64.0f);
        const float ex2 = (st4[1] + st4[3]) * (1.0f / 64.0f);
        const float rstd = rsqrtf(fmaxf(ex2 - mu * mu, 0.f) + p.eps);
        const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mu * rstd, -mu * rstd);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 g4 = *reinterpret_cast<const float4*>(sprm + 64 + c0 + 4 * k);
          const float4 b4 = *reinterpret_cast<const float4*>(sprm + 128 + c0 + 4 * k);
          v[2 * k] = fma2(fma2(v[2 * k], rs2, nm2), make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
          v[2 * k + 1] = fma2(fma2(v[2 * k + 1], rs2, nm2), make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
        }
      }
      if (gt == 0) E2_DBG(9, it);
      // messages -> the tile's sender-window buffer (consumed by epilogue 1 long ago); e' = e + m in place
      uint8_t* stg = smem + OFF_ST + s * 4 * BLK;
      {
        uint8_t* mrow = stg + (2 + half) * BLK + rsw;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4*>(mrow + ((k ^ rx) << 4)) = make_float4(v[2 * k].x, v[2 * k].y, v[2 * k + 1].x, v[2 * k + 1].y);
        if (p.has_out) {
          uint8_t* erow = stg + half * BLK + rsw;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float4* ptr = reinterpret_cast<float4*>(erow + ((k ^ rx) << 4));
            const float4 r = *ptr;
            const float2 o0 = add2(make_float2(r.x, r.y), v[2 * k]);
            const float2 o1 = add2(make_float2(r.z, r.w), v[2 * k + 1]);
            *ptr = make_float4(o0.x, o0.y, o1.x, o1.y);
          }
        }
      }
      fence_proxy_async();
      named_bar_sync(gbar, G2_THREADS);
      if (gt == 0) E2_DBG(10, it);
      if (p.has_out && gt == 0) {
        const uint32_t src = sbase + OFF_ST + s * 4 * BLK;
        tma_store_3d(&tmOut, src, 0, row0, b);
        tma_store_3d(&tmOut, src + BLK, 32, row0, b);
        bulk_commit();
      }
      {
        // segmented sum over the tile's receivers (CSR order): thread = (float4 column group, receiver group)
        const int cg = gt & 15, g = gt >> 4;
        const uint8_t* mbase = stg + (2 + (cg >> 3)) * BLK;
        const int chq = cg & 7;
        for (int j = g; j < nrec; j += G2_THREADS / 16) {
          const int k0 = lp[j], k1 = lp[j + 1];
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int k = k0; k < k1; ++k) {
            const float4 m4 = *reinterpret_cast<const float4*>(mbase + swz(k, chq));
            acc.x += m4.x; acc.y += m4.y; acc.z += m4.z; acc.w += m4.w;
          }
          if (p.mean) {
            const float sc = 1.0f / (float)max(k1 - k0, 1);
            acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
          }
          *reinterpret_cast<float4*>(p.aggr + ((long long)b * p.n_rec + r0 + j) * 64 + cg * 4) = acc;
        }
      }
      if (gt == 0) E2_DBG(12, it);
      if (gt == 0 && p.has_out) bulk_wait_read0();
      named_bar_sync(gbar, G2_THREADS);
      if (gt == 0) E2_DBG(13, it);
      if (gt == 0) mbar_arrive(bar_epi_done + 8 * s);
      meta = meta_n;
      meta_n = meta_nn;
    }
    if (gt == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == W_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// Node projections: out_i[b, r, :] = x_i[b, r, :] · Wslice_iᵀ (+ bias_i), Wslice = 64 x 64 block of a (64, ldw)
// weight, for ONE or TWO independent problems in a single launch (the sender and the receiver projection of an
// edge MLP: one launch and one prologue instead of two, and the tiles of both fill the SMs together).
// Persistent; per 128-row tile: TMA load (3 stages) -> 8 tcgen05.mma (K = 64) -> epilogue (8 warps: TMEM ->
// registers, + bias, staged in place in the input tile) -> TMA store.
// ---------------------------------------------------------------------------------------------------
namespace rl {
constexpr int THREADS = 320;  // warps 0-7 epilogue, 8 MMA, 9 TMA
constexpr int EPI = 256;
constexpr int NS = 3;
constexpr uint32_t BLK = 16384;
constexpr uint32_t WBLK = 8192;
constexpr uint32_t OFF_W = 0;  // 2 problems x 2 blocks
constexpr uint32_t OFF_ST = 4 * WBLK;
constexpr uint32_t OFF_MISC = OFF_ST + NS * 2 * BLK;
constexpr uint32_t SMEM = OFF_MISC + 1024;
}  // namespace rl

struct RowLinParams {
  const float* bias[2];  // may be null
  int batched[2];
  int n_tiles[2];
  int n_work0;  // tile-works of problem 0 (tiles x batches); works >= n_work0 belong to problem 1
  int n_work;
};

__global__ void __launch_bounds__(rl::THREADS, 1)
tc_rowlinear_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmW0,
                    const __grid_constant__ CUtensorMap tmOut0, const __grid_constant__ CUtensorMap tmX1,
                    const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmOut1,
                    const RowLinParams p) {
  using namespace rl;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t mb = sbase + OFF_MISC;
  const uint32_t bar_w = mb, bar_full = mb + 8, bar_free = mb + 32, bar_d_full = mb + 56, bar_d_free = mb + 72;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 96);
  if (warp == 8) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      for (int s = 0; s < NS; ++s) {
        mbar_init(bar_full + 8 * s, 1);
        mbar_init(bar_free + 8 * s, 1);
      }
      for (int t = 0; t < 2; ++t) {
        mbar_init(bar_d_full + 8 * t, 1);
        mbar_init(bar_d_free + 8 * t, EPI);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int n_work = p.n_work;
  const bool two = p.n_work0 < n_work;

  if (warp == 9) {
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      mbar_expect_tx(bar_w, (two ? 4u : 2u) * WBLK);
      for (int j = 0; j < 2; ++j) tma_load_2d(sbase + OFF_W + j * WBLK, &tmW0, bar_w, 32 * j, 0);
      if (two)
        for (int j = 0; j < 2; ++j) tma_load_2d(sbase + OFF_W + (2 + j) * WBLK, &tmW1, bar_w, 32 * j, 0);
      int it = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int pr = w >= p.n_work0;
        const int wl = pr ? w - p.n_work0 : w;
        const int b = wl / p.n_tiles[pr], t = wl - b * p.n_tiles[pr];
        const int s = it % NS;
        mbar_wait(bar_free + 8 * s, (uint32_t)(((it / NS) & 1) ^ 1));
        mbar_expect_tx(bar_full + 8 * s, 2u * BLK);
        for (int j = 0; j < 2; ++j)
          tma_load_3d(sbase + OFF_ST + (s * 2 + j) * BLK, pr ? &tmX1 : &tmX0, bar_full + 8 * s, 32 * j, t * 128,
                      p.batched[pr] ? b : 0, pol);
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, 64);
      mbar_wait(bar_w, 0);
      const uint64_t desc_w = umma_desc(sbase + OFF_W);
      const uint64_t desc_st = umma_desc(sbase + OFF_ST);
      int it = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int pr = w >= p.n_work0;
        const int s = it % NS, ts = it & 1;
        mbar_wait(bar_full + 8 * s, (uint32_t)((it / NS) & 1));
        mbar_wait(bar_d_free + 8 * ts, (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(tmem_base + ts * 64, desc_st + (uint64_t)(((s * 2 + j) * BLK) >> 4) + 2 * k,
                      desc_w + (uint64_t)(((pr * 2 + j) * WBLK) >> 4) + 2 * k, idesc, (uint32_t)((j | k) != 0));
        umma_commit(bar_d_full + 8 * ts);
      }
    }
  } else {
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane, c0 = half * 32;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    const int rx = row & 7;
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      const int pr = w >= p.n_work0;
      const int wl = pr ? w - p.n_work0 : w;
      const int b = wl / p.n_tiles[pr], t = wl - b * p.n_tiles[pr];
      const int s = it % NS, ts = it & 1;
      const float* bias = p.bias[pr];
      if (warp == 0) mbar_wait(bar_d_full + 8 * ts, (uint32_t)((it >> 1) & 1));
      named_bar_sync(1, EPI);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ts * 64 + t_lane + c0, v);
      tc_fence_before();
      mbar_arrive(bar_d_free + 8 * ts);
      uint8_t* orow = smem + OFF_ST + (s * 2 + half) * BLK + row * 128;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float4 o = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        if (bias) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4 * k));
          o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
        }
        *reinterpret_cast<float4*>(orow + ((k ^ rx) << 4)) = o;  // the MMAs have consumed the input tile
      }
      fence_proxy_async();
      named_bar_sync(1, EPI);
      if (tid == 0) {
        const CUtensorMap* mo = pr ? &tmOut1 : &tmOut0;
        tma_store_3d(mo, sbase + OFF_ST + (s * 2) * BLK, 0, t * 128, b);
        tma_store_3d(mo, sbase + OFF_ST + (s * 2 + 1) * BLK, 32, t * 128, b);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar_free + 8 * s);
      }
    }
    if (tid == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host
// one launch for 1 or 2 projection problems
int rowlinear_multi(const RowLinProblem* pr, int n_prob, cudaStream_t st) {
  NLAM_REQUIRE(n_prob == 1 || n_prob == 2, NLAM_E_INVALID, "rowlinear_multi: 1 or 2 problems");
  CUtensorMap mx[2], mw[2], mo[2];
  RowLinParams p;
  memset(&p, 0, sizeof(p));
  long long works[2] = {0, 0};
  for (int i = 0; i < n_prob; ++i) {
    const bool batched = pr[i].B_eff > 1;
    int rc = make_map(&mx[i], pr[i].x, 64, (uint64_t)pr[i].n_rows, batched ? (uint64_t)pr[i].B_eff : 1, 64,
                      batched ? (uint64_t)pr[i].x_bs : (uint64_t)pr[i].n_rows * 64, 128, true);
    if (rc) return rc;
    rc = make_map(&mw[i], pr[i].wslice, 64, 64, 1, (uint64_t)pr[i].ldw, 0, 64, false);
    if (rc) return rc;
    rc = make_map(&mo[i], pr[i].out, 64, (uint64_t)pr[i].n_rows, (uint64_t)pr[i].B_eff, 64, (uint64_t)pr[i].n_rows * 64, 128,
                  true);
    if (rc) return rc;
    p.bias[i] = pr[i].bias;
    p.batched[i] = batched;
    p.n_tiles[i] = (int)((pr[i].n_rows + 127) / 128);
    works[i] = (long long)p.n_tiles[i] * pr[i].B_eff;
  }
  if (n_prob == 1) {
    mx[1] = mx[0];
    mw[1] = mw[0];
    mo[1] = mo[0];
    p.n_tiles[1] = 1;
  }
  NLAM_REQUIRE(works[0] + works[1] < (1LL << 30), NLAM_E_UNSUPPORTED, "rowlinear: too many work items");
  p.n_work0 = (int)works[0];
  p.n_work = (int)(works[0] + works[1]);
  static unsigned attr_mask = 0;
  int dev = 0;
  NLAM_CUDA_OK(cudaGetDevice(&dev));
  if (!(attr_mask & (1u << (dev & 31)))) {
    NLAM_CUDA_OK(cudaFuncSetAttribute(tc_rowlinear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rl::SMEM));
    attr_mask |= 1u << (dev & 31);
  }
  const int grid = (int)std::min<long long>(p.n_work, num_sms());
  {
    double bytes = 0;
    for (int i = 0; i < n_prob; ++i) bytes += 2.0 * 256.0 * (double)pr[i].n_rows * pr[i].B_eff + 4.0 * (64 * 64 + 64);
    ProfScope ps("tc_rowlinear_kernel", st, bytes);
    tc_rowlinear_kernel<<<grid, rl::THREADS, rl::SMEM, st>>>(mx[0], mw[0], mo[0], mx[1], mw[1], mo[1], p);
  }
  count_launch();
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

int rowlinear(const float* x, int64_t x_bs, int64_t n_rows, int B_eff, const float* wslice, int ldw,
              const float* bias, float* out, cudaStream_t st) {
  RowLinProblem pr = {x, x_bs, n_rows, B_eff, wslice, ldw, bias, out};
  return rowlinear_multi(&pr, 1, st);
}

// sender and receiver projections of an edge MLP in one launch
int edge_projections(const float* send, int64_t send_bs, int64_t ns, int Bs, const float* rec, int64_t rec_bs, int64_t nr,
                     int Br, const float* w1, const float* b1, float* Ps, float* Pr, cudaStream_t st) {
  RowLinProblem pr[2] = {{send, send_bs, ns, Bs, w1 + 64, 192, nullptr, Ps}, {rec, rec_bs, nr, Br, w1 + 128, 192, b1, Pr}};
  return rowlinear_multi(pr, 2, st);
}

bool tc_edge2_supported(const NlamGraph* g, const NlamMlp* edge_mlp, int flags, const float* send, int64_t send_bs,
                        const float* rec, int64_t rec_bs, int B, int64_t send_rows) {
  (void)send_rows;
  // NLAM_TC_EDGE=v1 | v2 forces one formulation; default: the split kernels where the per-node projection
  // passes are small next to the edge work (mesh graph: ~9 edges per node), the K=192 kernel otherwise
  // (grid->mesh: projecting every grid node would cost more than gathering raw rows)
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("NLAM_TC_EDGE");
    mode = (e && e[0] == 'v' && e[1] == '1') ? 1 : (e && e[0] == 'v' && e[1] == '2') ? 2 : 0;
  }
  if (mode == 1) return false;
  if (!tc_edge_supported(g, edge_mlp, flags)) return false;
  if (!(aligned16(send) && aligned16(rec) && send_bs % 4 == 0 && rec_bs % 4 == 0)) return false;
  if (mode == 2) return true;
  const double edge_rows = (double)g->n_edges * B;
  const double node_rows = (double)g->n_send * ((send_bs == 0 || B == 1) ? 1 : B) +
                           (double)g->n_rec * ((rec_bs == 0 || B == 1) ? 1 : B);
  return edge_rows >= 2.5 * node_rows;
}

size_t tc_edge2_workspace_floats(const NlamGraph* g, int B, int64_t send_rows_max) {
  (void)send_rows_max;
  return (size_t)B * (size_t)(g->n_send + g->n_rec) * 64 + 64;
}

int tc_edge2(const NlamGraph* g, const NlamMlp* edge_mlp, const float* send, int64_t send_bs, const float* rec,
             int64_t rec_bs, const float* edge, int64_t edge_bs, float* edge_out, float* aggr_out, int B, int flags,
             cudaStream_t st, int64_t send_rows, float* ws) {
  (void)send_rows;
  NLAM_REQUIRE(aligned16(edge) && aligned16(aggr_out) && (!edge_out || aligned16(edge_out)) && edge_bs % 4 == 0 &&
                   aligned16(ws),
               NLAM_E_INVALID, "tc_edge2: pointers / strides must be 16-byte aligned");
  // node projections over exactly the rows the edges reference; batch-broadcast inputs are projected once
  const int Bs = (send_bs == 0 || B == 1) ? 1 : B;
  const int Br = (rec_bs == 0 || B == 1) ? 1 : B;
  const int64_t ns = g->n_send, nr = g->n_rec;
  float* Ps = ws;
  float* Pr = ws + (size_t)Bs * ns * 64;
  const float* w1 = edge_mlp->w[0];  // (64, 192): columns [e | sender | receiver]
  int rc = edge_projections(send, send_bs, ns, Bs, rec, rec_bs, nr, Br, w1, edge_mlp->b[0], Ps, Pr, st);
  if (rc) return rc;

  CUtensorMap me, mw1, mw2, mo, mps;
  const bool batched = edge_bs != 0 && B > 1;
  rc = make_map(&me, edge, 64, (uint64_t)g->n_edges, batched ? (uint64_t)B : 1, 64,
                batched ? (uint64_t)edge_bs : (uint64_t)g->n_edges * 64, 128, true);
  if (rc) return rc;
  rc = make_map(&mw1, w1, 64, 64, 1, 192, 0, 64, false);  // the e columns of W1
  if (rc) return rc;
  rc = make_map(&mw2, edge_mlp->w[1], 64, 64, 1, 64, 0, 64, false);
  if (rc) return rc;
  memset(&mo, 0, sizeof(mo));
  if (edge_out) {
    rc = make_map(&mo, edge_out, 64, (uint64_t)g->n_edges, (uint64_t)B, 64, (uint64_t)g->n_edges * 64, 128, true);
    if (rc) return rc;
  } else {
    mo = me;
  }
  rc = make_map(&mps, Ps, 64, (uint64_t)ns * Bs, 1, 64, 0, 1, false);
  if (rc) return rc;
  Edge2Params p;
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
    NLAM_CUDA_OK(cudaFuncSetAttribute(tc_edge2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e2::SMEM));
    attr_mask |= 1u << (dev & 31);
  }
  const long long n_work = (long long)p.n_tiles * p.B;
  NLAM_REQUIRE(n_work < (1LL << 31) - 4096, NLAM_E_UNSUPPORTED, "tc_edge2: too many work items");
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
    ProfScope ps("tc_edge2_kernel", st, edge_algorithmic_bytes(g, B, send_bs, rec_bs, edge_bs, edge_out != nullptr, 64));
    tc_edge2_kernel<<<grid, e2::THREADS, e2::SMEM, st>>>(me, mw1, mw2, mo, mps, p);
  }
  count_launch();
  if (dbg_on) {
    long long h[256];
    NLAM_CUDA_OK(cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
    NLAM_CUDA_OK(cudaStreamSynchronize(st));
    long long t0 = h[0];
    fprintf(stderr, "[nlam tc_edge2 timeline] grid=%d work=%lld (cycles rel. to first TMA issue)\n", grid, n_work);
    fprintf(stderr, " it    tma       -       -  g1_iss  g2_iss  e2wait  d1_rdy e1_done  d2_rdy  ln_done  staged       - reduced     end\n");
    for (int it = 0; it < 12; ++it) {
      fprintf(stderr, "%3d ", it);
      for (int k = 0; k < 14; ++k) fprintf(stderr, "%7lld ", h[it * 16 + k] ? h[it * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

}  // namespace nlam
