// Node update of an InteractionNet layer chained with the node projections of the NEXT layer's edge MLP (H = 64, a stack
// of layers over one node set: the mesh processor, reference graph_lam.py:117-126 / hi_lam.py same-level stacks):
//   x'   = x + LN(W2·SiLU(W1·[x | aggr] + b1) + b2)                 node update       (reference gnn_layers.py:148-151)
//   P_s  = W1s'·x' ,  P_r = W1r'·x' + b1'                            what the next layer's edge kernel gathers (tc5.cu / tc8.cu:
//                                                                   W1'·[e; x_s; x_r] + b1' = W1e'·e + P_s[src] + P_r[dst])
// As two kernels (tc4.cu node update, tc2.cu projections) x' is written, read again by the projection pass, and two persistent
// launches of ~40 µs each process a node set of a few thousand rows (11 tiles per CTA at the bench batch).  Here the x' tile
// that epilogue 2 leaves in its shared-memory slot (K-major, 128B-swizzled) is the A operand of a third GEMM with the 128 rows
// [W1s'; W1r'] as B; the 128 accumulator columns overwrite the stage's (dead) D | hidden columns and leave as two TMA tensor stores.
// Pipeline as tc9.cu: four... three GEMMs per tile, the SiLU group also finishes the projections (whichever is ready first).
#include "tc_ptx.cuh"

namespace nlam {

namespace r10 {
constexpr int THREADS = 640;
constexpr int EPI = 256;
constexpr int W_E1 = 8, W_MMA = 16, W_RING = 17, W_ST = 18;
constexpr int NRS = 2;  // node-tile slots
constexpr int NT = 2;   // TMEM stages
constexpr uint32_t BLK = 16384;
constexpr uint32_t WBLK = 8192;
constexpr uint32_t OFF_W1 = 0;              // 4 blocks (K = 128)
constexpr uint32_t OFF_W2 = 4 * WBLK;       // 2 blocks
constexpr uint32_t OFF_WP = 6 * WBLK;       // 2 k-blocks of 128 rows (16 KB each): rows 0-63 W1s', rows 64-127 W1r'
constexpr uint32_t OFF_R = 10 * WBLK;       // NRS x 32 KB: x tile -> x' -> P_s tile
constexpr uint32_t OFF_G = OFF_R + NRS * 2 * BLK;  // aggregate tile
constexpr uint32_t OFF_P = OFF_G + 2 * BLK;        // P_r tile
constexpr uint32_t OFF_MISC = OFF_P + 2 * BLK;
constexpr uint32_t SMEM = OFF_MISC + 2048;
}  // namespace r10

struct NodeProjParams {
  int rec_batched;
  const float* b1;
  const float* b2;
  const float* gamma;
  const float* beta;
  float eps;
  const float* b1n;  // first-layer bias of the next layer's edge MLP (goes into P_r)
  long long n_rows;
  int B;
  int n_tiles;
};

__global__ void __launch_bounds__(r10::THREADS, 1)
tc_node_proj_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmG,
                    const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                    const __grid_constant__ CUtensorMap tmWp, const __grid_constant__ CUtensorMap tmOut,
                    const __grid_constant__ CUtensorMap tmPs, const __grid_constant__ CUtensorMap tmPr,
                    const NodeProjParams p) {
  using namespace r10;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if ((sbase & 1023u) != 0) {
    if (tid == 0) printf("nlam tc_node_proj: dynamic shared memory not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t mb = sbase + OFF_MISC;
  const uint32_t bar_w = mb + 0;
  const uint32_t bar_wscaled = mb + 8;      // W1 halved in place (256 arrivals)
  const uint32_t bar_r_full = mb + 16;      // [2] node tile landed
  const uint32_t bar_r_free = mb + 32;      // [2] P_s tile stored, slot reusable
  const uint32_t bar_g_full = mb + 48;      // aggregate tile landed
  const uint32_t bar_g_free = mb + 56;      // first GEMM has consumed it
  const uint32_t bar_d1_full = mb + 64;     // [2]
  const uint32_t bar_hb_full = mb + 80;     // [2] 256 arrivals
  const uint32_t bar_d2_full = mb + 96;     // [2]
  const uint32_t bar_mid = mb + 112;        // [2] x' written over the x tile (256 arrivals)
  const uint32_t bar_d3_full = mb + 128;    // [2] projections accumulated
  const uint32_t bar_xdone = mb + 144;      // [2] the x' store has read the slot
  const uint32_t bar_pstaged = mb + 160;    // [2] P_s / P_r tiles written (256 arrivals)
  const uint32_t bar_d_free = mb + 176;     // [2] TMEM stage drained (256 arrivals)
  const uint32_t bar_p_free = mb + 192;     // P_r staging tile stored
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 200);
  volatile int* sel = reinterpret_cast<volatile int*>(smem + OFF_MISC + 208);  // [2]
  float* sprm = reinterpret_cast<float*>(smem + OFF_MISC + 256);  // gamma | beta | b1/2 | b1n (64 each)
  float *s_gamma = sprm, *s_beta = sprm + 64, *s_b1h = sprm + 128, *s_b1n = sprm + 192;

  if (warp == W_MMA) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_wscaled, EPI);
      mbar_init(bar_g_full, 1);
      mbar_init(bar_g_free, 1);
      mbar_init(bar_p_free, 1);
      for (int t = 0; t < 2; ++t) {
        mbar_init(bar_r_full + 8 * t, 1);
        mbar_init(bar_r_free + 8 * t, 1);
        mbar_init(bar_d1_full + 8 * t, 1);
        mbar_init(bar_hb_full + 8 * t, EPI);
        mbar_init(bar_d2_full + 8 * t, 1);
        mbar_init(bar_mid + 8 * t, EPI);
        mbar_init(bar_d3_full + 8 * t, 1);
        mbar_init(bar_xdone + 8 * t, 1);
        mbar_init(bar_pstaged + 8 * t, EPI);
        mbar_init(bar_d_free + 8 * t, EPI);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == W_RING && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWp) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmPs) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmPr) : "memory");
  }
  pdl_launch_dependents();
  pdl_wait();  // everything below may read what the previous kernel in the stream wrote
  if (tid < 64) {
    s_gamma[tid] = p.gamma[tid];
    s_beta[tid] = p.beta[tid];
    s_b1h[tid] = 0.5f * p.b1[tid];
    s_b1n[tid] = p.b1n[tid];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  // TMEM columns: stage ts: D at ts*128 (first and second GEMM), hidden at +64; the third GEMM writes all 128 columns
  // (P_s | P_r); LayerNorm scratch at 384
  const int n_work = p.n_tiles * p.B;
  int n_my = 0;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) ++n_my;

  if (warp == W_RING) {
    // =============================== loads ===============================
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();
      mbar_expect_tx(bar_w, 10u * WBLK);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + OFF_W1 + kb * WBLK, &tmW1, bar_w, 32 * kb, 0);
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(sbase + OFF_W2 + kb * WBLK, &tmW2, bar_w, 32 * kb, 0);
      for (int kb = 0; kb < 2; ++kb) {  // W1' columns [e | sender | receiver]: the sender and receiver column blocks
        tma_load_2d(sbase + OFF_WP + kb * 2 * WBLK, &tmWp, bar_w, 64 + 32 * kb, 0);
        tma_load_2d(sbase + OFF_WP + kb * 2 * WBLK + WBLK, &tmWp, bar_w, 128 + 32 * kb, 0);
      }
      for (int ti = 0; ti < n_my; ++ti) {
        const int w = blockIdx.x + ti * gridDim.x;
        const int b = w / p.n_tiles, t = w - b * p.n_tiles;
        const int rs = ti % NRS;
        mbar_wait(bar_r_free + 8 * rs, (uint32_t)(((ti / NRS) & 1) ^ 1));
        const uint32_t rfull = bar_r_full + 8 * rs;
        mbar_expect_tx(rfull, 2u * BLK);
        const uint32_t rdst = sbase + OFF_R + rs * 2 * BLK;
        tma_load_3d(rdst, &tmR, rfull, 0, t * 128, p.rec_batched ? b : 0, pol_stream);
        tma_load_3d(rdst + BLK, &tmR, rfull, 32, t * 128, p.rec_batched ? b : 0, pol_stream);
        mbar_wait(bar_g_free, (uint32_t)((ti & 1) ^ 1));
        mbar_expect_tx(bar_g_full, 2u * BLK);
        tma_load_3d(sbase + OFF_G, &tmG, bar_g_full, 0, t * 128, b, pol_stream);
        tma_load_3d(sbase + OFF_G + BLK, &tmG, bar_g_full, 32, t * 128, b, pol_stream);
      }
    }
  } else if (warp == W_ST) {
    // =============================== stores: x', then P_s / P_r ===============================
    if (lane == 0) {
      for (int ti = 0; ti < n_my; ++ti) {
        const int w = blockIdx.x + ti * gridDim.x;
        const int b = w / p.n_tiles, t = w - b * p.n_tiles;
        const int ts = ti % NT, rs = ti % NRS;
        const uint32_t rsl = sbase + OFF_R + rs * 2 * BLK;
        mbar_wait(bar_mid + 8 * ts, (uint32_t)((ti / NT) & 1));
        tma_store_3d(&tmOut, rsl, 0, t * 128, b);
        tma_store_3d(&tmOut, rsl + BLK, 32, t * 128, b);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar_xdone + 8 * ts);
        mbar_wait(bar_pstaged + 8 * ts, (uint32_t)((ti / NT) & 1));
        tma_store_3d(&tmPs, rsl, 0, t * 128, b);
        tma_store_3d(&tmPs, rsl + BLK, 32, t * 128, b);
        tma_store_3d(&tmPr, sbase + OFF_P, 0, t * 128, b);
        tma_store_3d(&tmPr, sbase + OFF_P + BLK, 32, t * 128, b);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar_r_free + 8 * rs);
        mbar_arrive(bar_p_free);
      }
      bulk_wait0();
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issue (uniform control flow, one elected lane) ===============================
    const uint32_t idesc = umma_idesc_tf32(128, 64);
    const uint32_t idesc_p = umma_idesc_tf32(128, 128);
    mbar_wait(bar_w, 0);
    mbar_wait(bar_wscaled, 0);
    tc_fence_after();
    const uint64_t desc_w1 = umma_desc(sbase + OFF_W1);
    const uint64_t desc_w2 = umma_desc(sbase + OFF_W2);
    const uint64_t desc_wp = umma_desc(sbase + OFF_WP);
    const uint64_t desc_r = umma_desc(sbase + OFF_R);
    const uint64_t desc_g = umma_desc(sbase + OFF_G);
    int g1 = 0, g2 = 0, g3 = 0;
    uint32_t idle = 0;
    while (g3 < n_my) {
      bool progress = false;
      if (g3 < g2) {  // projections: D[0:128] = x' · [W1s'; W1r']ᵀ
        const int ts = g3 % NT, rs = g3 % NRS;
        if (mbar_test_u(bar_mid + 8 * ts, (uint32_t)((g3 / NT) & 1))) {
          tc_fence_after();
          const uint32_t dd = tmem_base + ts * 128;
          const uint64_t a0 = desc_r + (uint64_t)((rs * 2 * BLK) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_tf32(dd, a0 + (uint64_t)((jj * BLK) >> 4) + 2 * kk, desc_wp + (uint64_t)((jj * 2 * WBLK) >> 4) + 2 * kk,
                          idesc_p, (uint32_t)((jj | kk) != 0));
            umma_commit(bar_d3_full + 8 * ts);
          }
          __syncwarp();
          ++g3;
          progress = true;
        }
      }
      if (g2 < g1) {  // node MLP, second Linear
        const int ts = g2 % NT;
        if (mbar_test_u(bar_hb_full + 8 * ts, (uint32_t)((g2 / NT) & 1))) {
          tc_fence_after();
          const uint32_t dd = tmem_base + ts * 128;
          if (elect_one()) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_tf32_ts(dd, dd + 64 + (uint32_t)(jj * 32 + kk * 8), desc_w2 + (uint64_t)((jj * WBLK) >> 4) + 2 * kk, idesc,
                             (uint32_t)((jj | kk) != 0));
            umma_commit(bar_d2_full + 8 * ts);
          }
          __syncwarp();
          ++g2;
          progress = true;
        }
      }
      if (g1 < n_my && g1 < g3 + NT) {  // node MLP, first Linear (K = 128)
        const int ts = g1 % NT, rs = g1 % NRS;
        bool ready = mbar_test_u(bar_r_full + 8 * rs, (uint32_t)((g1 / NRS) & 1));
        if (ready) ready = mbar_test_u(bar_g_full, (uint32_t)(g1 & 1));
        if (ready && g1 >= NT) ready = mbar_test_u(bar_d_free + 8 * ts, (uint32_t)(((g1 / NT) - 1) & 1));
        if (ready) {
          tc_fence_after();
          const uint32_t dd = tmem_base + ts * 128;
          const uint64_t ar = desc_r + (uint64_t)((rs * 2 * BLK) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              const uint64_t a0 = s ? desc_g : ar;
              const uint64_t b0 = desc_w1 + (uint64_t)((s * 2 * WBLK) >> 4);
#pragma unroll
              for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_tf32(dd, a0 + (uint64_t)((jj * BLK) >> 4) + 2 * kk, b0 + (uint64_t)((jj * WBLK) >> 4) + 2 * kk, idesc,
                            (uint32_t)((s | jj | kk) != 0));
            }
            umma_commit(bar_d1_full + 8 * ts);
            umma_commit(bar_g_free);
          }
          __syncwarp();
          ++g1;
          progress = true;
        }
      }
      if (progress) idle = 0;
      else if (__nanosleep(40), ++idle > (1u << 24)) {
        if (lane == 0) printf("nlam tc_node_proj: MMA issuer timeout (block %d g %d %d %d of %d)\n", blockIdx.x, g1, g2, g3, n_my);
        __trap();
      }
    }
  } else if (warp >= W_E1 && warp < W_MMA) {
    // =============================== epilogue group 1: SiLU of the node MLP, and the projection tiles ===============================
    const bool lead = warp == W_E1;
    const int q = warp & 3;
    const int half = (warp - W_E1) >> 2;
    const int row = q * 32 + lane;
    const int c0 = half * 32;
    const int rx = row & 7;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    {
      mbar_wait(bar_w, 0);
      float4* wq = reinterpret_cast<float4*>(smem + OFF_W1) + (tid - W_E1 * 32);
      for (int i = 0; i < 8; ++i) {  // 32 KB = 2048 float4 over 256 threads
        float4 x = wq[i * EPI];
        x.x *= 0.5f; x.y *= 0.5f; x.z *= 0.5f; x.w *= 0.5f;
        wq[i * EPI] = x;
      }
      fence_proxy_async();
      mbar_arrive(bar_wscaled);
    }
    int round = 0;
    for (int i = 0; i <= n_my; ++i) {
      bool pend_a = i < n_my, pend_b = i >= 1;
      const int ti = i - 1;
      while (pend_a || pend_b) {
        if (lead) {
          int pick = -1;
          uint32_t spins = 0;
          while (pick < 0) {
            if (pend_b && mbar_test_u(bar_d3_full + 8 * (ti % NT), (uint32_t)((ti / NT) & 1)) &&
                mbar_test_u(bar_xdone + 8 * (ti % NT), (uint32_t)((ti / NT) & 1)) &&
                mbar_test_u(bar_p_free, (uint32_t)((ti & 1) ^ 1)))
              pick = 1;
            else if (pend_a && mbar_test_u(bar_d1_full + 8 * (i % NT), (uint32_t)((i / NT) & 1))) pick = 0;
            else if (__nanosleep(20), ++spins > (1u << 24)) {
              if (lane == 0) printf("nlam tc_node_proj: epilogue group 1 timeout (block %d tile %d)\n", blockIdx.x, i);
              __trap();
            }
          }
          if (lane == 0) sel[round & 1] = pick;
        }
        named_bar_sync(1, EPI);
        const int pick = sel[round & 1];
        ++round;
        tc_fence_after();
        if (pick == 0) {  // SiLU of tile i
          const int ts = i % NT;
          const uint32_t d1 = tmem_base + ts * 128 + t_lane + c0;
          float v[32];
          tmem_ld32(d1, v);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 bb = *reinterpret_cast<const float4*>(s_b1h + c0 + 4 * k);
            const float2 h0 = add2(make_float2(v[4 * k], v[4 * k + 1]), make_float2(bb.x, bb.y));
            const float2 h1 = add2(make_float2(v[4 * k + 2], v[4 * k + 3]), make_float2(bb.z, bb.w));
            const float2 o0 = fma2(h0, make_float2(tanh_fast(h0.x), tanh_fast(h0.y)), h0);
            const float2 o1 = fma2(h1, make_float2(tanh_fast(h1.x), tanh_fast(h1.y)), h1);
            v[4 * k] = o0.x;
            v[4 * k + 1] = o0.y;
            v[4 * k + 2] = o1.x;
            v[4 * k + 3] = o1.y;
          }
          tmem_st32(d1 + 64, v);
          tc_fence_before();
          mbar_arrive(bar_hb_full + 8 * ts);
          pend_a = false;
        } else {  // projections of tile i - 1: P_s over the (stored) x' tile, P_r + b1' into the staging tile
          const int ts = ti % NT, rs = ti % NRS;
          uint8_t* ps_row = smem + OFF_R + rs * 2 * BLK + half * BLK + row * 128;
          uint8_t* pr_row = smem + OFF_P + half * BLK + row * 128;
          float v[32];
          tmem_ld32(tmem_base + ts * 128 + t_lane + c0, v);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(ps_row + ((k ^ rx) << 4)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          tmem_ld32(tmem_base + ts * 128 + 64 + t_lane + c0, v);
          tc_fence_before();
          mbar_arrive(bar_d_free + 8 * ts);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 bb = *reinterpret_cast<const float4*>(s_b1n + c0 + 4 * k);
            *reinterpret_cast<float4*>(pr_row + ((k ^ rx) << 4)) =
                make_float4(v[4 * k] + bb.x, v[4 * k + 1] + bb.y, v[4 * k + 2] + bb.z, v[4 * k + 3] + bb.w);
          }
          fence_proxy_async();
          mbar_arrive(bar_pstaged + 8 * ts);
          pend_b = false;
        }
      }
    }
  } else if (warp < W_E1) {
    // =============================== epilogue group 2: bias, LayerNorm, + x -> x' in place over the x tile ===============
    const int q = warp & 3;
    const int half = warp >> 2;
    const int row = q * 32 + lane;
    const int c0 = half * 32;
    const int rx = row & 7;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    const int pbar = 4 + q;
    const uint32_t ln_col = tmem_base + 384 + t_lane;
    float2 b2r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) b2r[i] = make_float2(__ldg(p.b2 + c0 + 2 * i), __ldg(p.b2 + c0 + 2 * i + 1));
    for (int i = 0; i < n_my; ++i) {
      const int ts = i % NT, rs = i % NRS;
      if (warp == 0) mbar_wait(bar_d2_full + 8 * ts, (uint32_t)((i / NT) & 1));
      named_bar_sync(2, EPI);
      tc_fence_after();
      float vf[32];
      tmem_ld32(tmem_base + ts * 128 + t_lane + c0, vf);
      float2 v[16];
      float2 sm2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        v[k] = add2(make_float2(vf[2 * k], vf[2 * k + 1]), b2r[k]);
        sm2 = add2(sm2, v[k]);
        sq2 = fma2(v[k], v[k], sq2);
      }
      const uint32_t scr = ln_col + 4 * (i & 1);
      tmem_st2(scr + 2 * half, sm2.x + sm2.y, sq2.x + sq2.y);
      tc_fence_before();
      named_bar_sync(pbar, 64);
      tc_fence_after();
      float st4[4];
      tmem_ld4(scr, st4);
      const float mu = (st4[0] + st4[2]) * (1.0f / 64.0f);
      const float ex2 = (st4[1] + st4[3]) * (1.0f / 64.0f);
      const float rstd = rsqrtf(fmaxf(ex2 - mu * mu, 0.f) + p.eps);
      const float2 rs2 = make_float2(rstd, rstd);
      const float2 nm2 = make_float2(-mu * rstd, -mu * rstd);
      uint8_t* orow = smem + OFF_R + rs * 2 * BLK + half * BLK + row * 128;
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        const float4 g4 = *reinterpret_cast<const float4*>(s_gamma + c0 + 4 * k8);
        const float4 b4 = *reinterpret_cast<const float4*>(s_beta + c0 + 4 * k8);
        float4* ptr = reinterpret_cast<float4*>(orow + ((k8 ^ rx) << 4));
        const float4 r4v = *ptr;
        const float2 o0 = add2(fma2(fma2(v[2 * k8], rs2, nm2), make_float2(g4.x, g4.y), make_float2(b4.x, b4.y)),
                               make_float2(r4v.x, r4v.y));
        const float2 o1 = add2(fma2(fma2(v[2 * k8 + 1], rs2, nm2), make_float2(g4.z, g4.w), make_float2(b4.z, b4.w)),
                               make_float2(r4v.z, r4v.w));
        *ptr = make_float4(o0.x, o0.y, o1.x, o1.y);
      }
      fence_proxy_async();  // generic writes -> tcgen05.mma operand reads and the TMA store
      tc_fence_before();
      mbar_arrive(bar_mid + 8 * ts);
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
bool tc_node_proj_supported(const NlamMlp* node_mlp, const NlamMlp* next_edge_mlp, const float* rec, int64_t rec_bs,
                            const float* aggr, int64_t n_rows, const float* out, const float* proj_out) {
  static int on = -1;
  if (on < 0) on = getenv("NLAM_TC_NO_NODE_PROJ") ? 0 : 1;
  if (!on) return false;
  int no1 = 0, no2 = 0;
  if (n_rows < 1 || n_rows >= (1LL << 31) - 256) return false;
  if (!mlp_shape_ok(node_mlp, &no1) || no1 != 64 || node_mlp->in_dim != 128 || !node_mlp->ln_gamma || !node_mlp->ln_beta) return false;
  if (!mlp_shape_ok(next_edge_mlp, &no2) || no2 != 64 || next_edge_mlp->in_dim != 192) return false;
  return aligned16(rec) && aligned16(aggr) && aligned16(out) && aligned16(proj_out) && rec_bs % 4 == 0;
}

// rec' (B, n_rows, 64) -> out; P_s (B, n_rows, 64) -> proj_out, P_r (B, n_rows, 64) -> proj_out + B*n_rows*64
int tc_node_proj(const NlamMlp* node_mlp, const NlamMlp* next_edge_mlp, const float* rec, int64_t rec_bs, const float* aggr,
                 int64_t n_rows, int B, float* out, float* proj_out, cudaStream_t st) {
  NLAM_REQUIRE(tc_node_proj_supported(node_mlp, next_edge_mlp, rec, rec_bs, aggr, n_rows, out, proj_out), NLAM_E_UNSUPPORTED,
               "tc_node_proj: unsupported shapes");
  CUtensorMap mr, mg, w1, w2, wp, mo, mps, mpr;
  const bool batched = rec_bs != 0 && B > 1;
  int rc = make_map(&mr, rec, 64, (uint64_t)n_rows, batched ? (uint64_t)B : 1, 64, batched ? (uint64_t)rec_bs : (uint64_t)n_rows * 64,
                    128, true);
  if (rc) return rc;
  rc = make_map(&mg, aggr, 64, (uint64_t)n_rows, (uint64_t)B, 64, (uint64_t)n_rows * 64, 128, true);
  if (rc) return rc;
  rc = make_map(&w1, node_mlp->w[0], 128, 64, 1, 128, 0, 64, false);
  if (rc) return rc;
  rc = make_map(&w2, node_mlp->w[1], 64, 64, 1, 64, 0, 64, false);
  if (rc) return rc;
  rc = make_map(&wp, next_edge_mlp->w[0], 192, 64, 1, 192, 0, 64, false);
  if (rc) return rc;
  rc = make_map(&mo, out, 64, (uint64_t)n_rows, (uint64_t)B, 64, (uint64_t)n_rows * 64, 128, true);
  if (rc) return rc;
  rc = make_map(&mps, proj_out, 64, (uint64_t)n_rows, (uint64_t)B, 64, (uint64_t)n_rows * 64, 128, true);
  if (rc) return rc;
  rc = make_map(&mpr, proj_out + (size_t)B * n_rows * 64, 64, (uint64_t)n_rows, (uint64_t)B, 64, (uint64_t)n_rows * 64, 128, true);
  if (rc) return rc;
  NodeProjParams p;
  memset(&p, 0, sizeof(p));
  p.rec_batched = batched;
  p.b1 = node_mlp->b[0];
  p.b2 = node_mlp->b[1];
  p.gamma = node_mlp->ln_gamma;
  p.beta = node_mlp->ln_beta;
  p.eps = node_mlp->ln_eps;
  p.b1n = next_edge_mlp->b[0];
  p.n_rows = n_rows;
  p.B = B;
  p.n_tiles = (int)((n_rows + 127) / 128);
  static unsigned attr_mask = 0;
  int dev = 0;
  NLAM_CUDA_OK(cudaGetDevice(&dev));
  if (!(attr_mask & (1u << (dev & 31)))) {
    NLAM_CUDA_OK(cudaFuncSetAttribute(tc_node_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r10::SMEM));
    attr_mask |= 1u << (dev & 31);
  }
  const long long n_work = (long long)p.n_tiles * p.B;
  NLAM_REQUIRE(n_work < (1LL << 30), NLAM_E_UNSUPPORTED, "tc_node_proj: too many work items");
  const int grid = (int)std::min<long long>(n_work, num_sms());
  {
    const long long rows = (long long)B * n_rows;
    const long long nb = 4LL * 64 * ((batched ? rows : n_rows) + rows) + 4LL * 64 * 3 * rows + 4LL * (64 * 128 + 64 * 64 + 128 * 64 + 64 * 5);
    ProfScope ps("tc_node_proj_kernel", st, nb);
    NLAM_CUDA_OK(launch_pdl(tc_node_proj_kernel, grid, r10::THREADS, r10::SMEM, st, mr, mg, w1, w2, wp, mo, mps, mpr, p));
  }
  count_launch();
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

}  // namespace nlam
