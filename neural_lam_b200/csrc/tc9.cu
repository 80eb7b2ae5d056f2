// Tail of the grid side of a forecast step in ONE kernel (H = 64):
//   grid' = grid + LN(W2·SiLU(W1·[grid | aggr] + b1) + b2)        node update of the mesh->grid InteractionNet
//                                                                   (reference gnn_layers.py:148-151)
//   y     = W4·SiLU(W3·grid' + b3) + b4                            output_map (reference graph/base.py:322)
//   new   = mask*boundary + (1-mask)*(prev + y*std + mean)          rescale + residual + boundary mix
//                                                                   (graph/base.py:339-342, autoregressive.py:128-131)
// As two kernels (tc4.cu) the updated grid representation goes to HBM and comes straight back: 2 x 4·64·B·G bytes
// (1.04 GB per step at the bench batch) that nobody else reads.  Here the 128-row tile of grid' stays in the shared-
// memory slot the grid tile arrived in — epilogue 2 writes it in place in the K-major 128B-swizzled layout, which IS
// the A operand of the third GEMM — and only the 17-column result leaves.
//
// Per tile: four tcgen05 GEMMs (TF32, accumulators in TMEM) and four epilogues; the two epilogue groups of tc4.cu each
// take one epilogue of either MLP, software-pipelined: group 1 runs SiLU of the node MLP for tile i, then SiLU of the
// output MLP for tile i-1; group 2 runs the LayerNorm/residual epilogue of tile i, then the narrow step epilogue of tile
// i-1 (prev / boundary / mask slabs arrive by 1-D bulk copies in the tile's slot once the third GEMM has consumed it,
// the result slab leaves by one bulk store).  Shared memory: 72 KB of weights, three grid-tile slots (a slot lives for
// the whole tile), one aggregate slot (free again after the first GEMM).  Three TMEM stages (D | hidden).
#include "tc_ptx.cuh"

namespace nlam {

namespace r9 {
constexpr int THREADS = 640;
constexpr int EPI = 256;
constexpr int W_E1 = 8, W_MMA = 16, W_RING = 17, W_ST = 18;
constexpr int NRS = 3;  // grid-tile slots
constexpr int NT = 3;   // TMEM stages
constexpr uint32_t BLK = 16384;
constexpr uint32_t WBLK = 8192;
constexpr uint32_t OFF_W1 = 0;              // 4 blocks (K = 128)
constexpr uint32_t OFF_W2 = 4 * WBLK;       // 2 blocks
constexpr uint32_t OFF_W3 = 6 * WBLK;       // 2 blocks
constexpr uint32_t OFF_W4 = 8 * WBLK;       // 2 blocks of 32 rows (4 KB each)
constexpr uint32_t OFF_R = 9 * WBLK;        // NRS x 32 KB
constexpr uint32_t OFF_G = OFF_R + NRS * 2 * BLK;
constexpr uint32_t OFF_MISC = OFF_G + 2 * BLK;
constexpr uint32_t SMEM = OFF_MISC + 3072;
// layout of a grid-tile slot once the third GEMM has consumed it
constexpr uint32_t S_PREV = 0, S_BND = 9216, S_OUT = 18432, S_MASK = 27648;
}  // namespace r9

struct NodeOutParams {
  int rec_batched;
  const float* b1;
  const float* b2;
  const float* gamma;
  const float* beta;
  float eps;
  const float* b3;
  const float* b4;
  int nout;
  long long n_rows;
  int B;
  int n_tiles;
  float* out;
  const float* ep_prev;
  const float* ep_bnd;
  const float* ep_mask;
  const float* ep_std;
  const float* ep_mean;
  long long* dbg;
};

#define R9_DBG(slot, it)                                                                  \
  do {                                                                                    \
    if (p.dbg && blockIdx.x == 0 && (it) < 16) p.dbg[(it) * 16 + (slot)] = clock64();     \
  } while (0)

__global__ void __launch_bounds__(r9::THREADS, 1)
tc_node_out_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmG,
                   const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                   const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ CUtensorMap tmW4,
                   const NodeOutParams p) {
  using namespace r9;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if ((sbase & 1023u) != 0) {
    if (tid == 0) printf("nlam tc_node_out: dynamic shared memory not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t mb = sbase + OFF_MISC;
  const uint32_t bar_w = mb + 0;
  const uint32_t bar_wscaled = mb + 8;      // W1 and W3 halved in place (256 arrivals)
  const uint32_t bar_r_full = mb + 16;      // [3] grid tile landed
  const uint32_t bar_r_free = mb + 40;      // [3] result slab stored, slot reusable
  const uint32_t bar_g_full = mb + 64;      // aggregate tile landed
  const uint32_t bar_g_free = mb + 72;      // first GEMM has consumed it (tcgen05.commit)
  const uint32_t bar_d1_full = mb + 80;     // [3]
  const uint32_t bar_hb_full = mb + 104;    // [3] 256 arrivals
  const uint32_t bar_d2_full = mb + 128;    // [3]
  const uint32_t bar_mid = mb + 152;        // [3] grid' written over the grid tile (256 arrivals)
  const uint32_t bar_d3_full = mb + 176;    // [3]
  const uint32_t bar_hb2_full = mb + 200;   // [3] 256 arrivals
  const uint32_t bar_d4_full = mb + 224;    // [3]
  const uint32_t bar_d_free = mb + 248;     // [3] accumulators of the TMEM stage drained by the last epilogue (256 arrivals)
  const uint32_t bar_staged = mb + 272;     // [3] result slab written (256 arrivals)
  const uint32_t bar_ep_full = mb + 296;    // [3] prev / boundary / mask slabs landed (tx bytes)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 320);
  volatile int* sel = reinterpret_cast<volatile int*>(smem + OFF_MISC + 328);  // [2 groups][2]: which stage a group runs next
  float* sprm = reinterpret_cast<float*>(smem + OFF_MISC + 512);  // gamma | beta | b1/2 | b3/2 | std | mean | b4 (64 each)
  float* s_gamma = sprm, *s_beta = sprm + 64, *s_b1h = sprm + 128, *s_b3h = sprm + 192, *s_std = sprm + 256,
        *s_mean = sprm + 320, *s_b4 = sprm + 384;

  if (warp == W_MMA) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_wscaled, EPI);
      mbar_init(bar_g_full, 1);
      mbar_init(bar_g_free, 1);
      for (int t = 0; t < NT; ++t) {
        mbar_init(bar_r_full + 8 * t, 1);
        mbar_init(bar_r_free + 8 * t, 1);
        mbar_init(bar_d1_full + 8 * t, 1);
        mbar_init(bar_hb_full + 8 * t, EPI);
        mbar_init(bar_d2_full + 8 * t, 1);
        mbar_init(bar_mid + 8 * t, EPI);
        mbar_init(bar_d3_full + 8 * t, 1);
        mbar_init(bar_hb2_full + 8 * t, EPI);
        mbar_init(bar_d4_full + 8 * t, 1);
        mbar_init(bar_d_free + 8 * t, EPI);
        mbar_init(bar_staged + 8 * t, EPI);
        mbar_init(bar_ep_full + 8 * t, 1);
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
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW3) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW4) : "memory");
  }
  pdl_launch_dependents();
  pdl_wait();  // everything below may read what the previous kernel in the stream wrote
  if (tid < 64) {
    s_gamma[tid] = p.gamma[tid];
    s_beta[tid] = p.beta[tid];
    s_b1h[tid] = 0.5f * p.b1[tid];
    s_b3h[tid] = 0.5f * p.b3[tid];
    s_std[tid] = (p.ep_prev && tid < p.nout) ? p.ep_std[tid] : 1.f;
    s_mean[tid] = (p.ep_prev && tid < p.nout) ? p.ep_mean[tid] : 0.f;
    s_b4[tid] = tid < p.nout ? p.b4[tid] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  // TMEM columns: stage ts: D at ts*128 (all four GEMMs of the tile in turn), hidden at +64; LayerNorm scratch at 384
  const int n_work = p.n_tiles * p.B;
  int n_my = 0;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) ++n_my;

  if (warp == W_RING) {
    // =============================== loads ===============================
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();
      mbar_expect_tx(bar_w, 8u * WBLK + WBLK);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + OFF_W1 + kb * WBLK, &tmW1, bar_w, 32 * kb, 0);
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(sbase + OFF_W2 + kb * WBLK, &tmW2, bar_w, 32 * kb, 0);
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(sbase + OFF_W3 + kb * WBLK, &tmW3, bar_w, 32 * kb, 0);
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(sbase + OFF_W4 + kb * (WBLK / 2), &tmW4, bar_w, 32 * kb, 0);
      for (int ti = 0; ti < n_my; ++ti) {
        const int w = blockIdx.x + ti * gridDim.x;
        const int b = w / p.n_tiles, t = w - b * p.n_tiles;
        const int rs = ti % NRS;
        mbar_wait(bar_r_free + 8 * rs, (uint32_t)(((ti / NRS) & 1) ^ 1));
        const uint32_t rfull = bar_r_full + 8 * rs;
        mbar_expect_tx(rfull, 2u * BLK);
        R9_DBG(0, ti);
        const uint32_t rdst = sbase + OFF_R + rs * 2 * BLK;
        tma_load_3d(rdst, &tmR, rfull, 0, t * 128, p.rec_batched ? b : 0, pol_stream);
        tma_load_3d(rdst + BLK, &tmR, rfull, 32, t * 128, p.rec_batched ? b : 0, pol_stream);
        mbar_wait(bar_g_free, (uint32_t)((ti & 1) ^ 1));
        mbar_expect_tx(bar_g_full, 2u * BLK);
        tma_load_3d(sbase + OFF_G, &tmG, bar_g_full, 0, t * 128, b, pol_stream);
        tma_load_3d(sbase + OFF_G + BLK, &tmG, bar_g_full, 32, t * 128, b, pol_stream);
        // pull the tiles after next into L2 (a slot's life starts with its load: 3.4 k cycles from DRAM, measured)
        for (int a = (ti == 0 ? 1 : 2); a <= 2; ++a) {
          if (ti + a >= n_my) break;
          const int w2 = blockIdx.x + (ti + a) * gridDim.x;
          const int b2 = w2 / p.n_tiles, t2 = w2 - b2 * p.n_tiles;
          tma_prefetch_3d(&tmR, 0, t2 * 128, p.rec_batched ? b2 : 0);
          tma_prefetch_3d(&tmR, 32, t2 * 128, p.rec_batched ? b2 : 0);
          tma_prefetch_3d(&tmG, 0, t2 * 128, b2);
          tma_prefetch_3d(&tmG, 32, t2 * 128, b2);
        }
      }
    }
  } else if (warp == W_ST) {
    // =============================== result stores ===============================
    if (lane == 0) {
      for (int ti = 0; ti < n_my; ++ti) {
        const int w = blockIdx.x + ti * gridDim.x;
        const int b = w / p.n_tiles, t = w - b * p.n_tiles;
        const int ts = ti % NT, rs = ti % NRS;
        const int nrows = (int)min(128LL, p.n_rows - (long long)t * 128);
        mbar_wait(bar_staged + 8 * ts, (uint32_t)((ti / NT) & 1));
        bulk_store_1d(p.out + ((long long)b * p.n_rows + (long long)t * 128) * p.nout, sbase + OFF_R + rs * 2 * BLK + S_OUT,
                      (uint32_t)(nrows * p.nout * 4));
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar_r_free + 8 * rs);
        R9_DBG(9, ti);
      }
      bulk_wait0();
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issue (uniform control flow, one elected lane) ===============================
    const uint32_t idesc = umma_idesc_tf32(128, 64);
    const uint32_t idesc4 = umma_idesc_tf32(128, 32);
    mbar_wait(bar_w, 0);
    mbar_wait(bar_wscaled, 0);
    tc_fence_after();
    const uint64_t desc_w1 = umma_desc(sbase + OFF_W1);
    const uint64_t desc_w2 = umma_desc(sbase + OFF_W2);
    const uint64_t desc_w3 = umma_desc(sbase + OFF_W3);
    const uint64_t desc_w4 = umma_desc(sbase + OFF_W4);
    const uint64_t desc_r = umma_desc(sbase + OFF_R);
    const uint64_t desc_g = umma_desc(sbase + OFF_G);
    int g1 = 0, g2 = 0, g3 = 0, g4 = 0;
    uint32_t idle = 0;
    while (g4 < n_my) {
      bool progress = false;
      if (g4 < g3) {  // output MLP, second Linear (N = 32): D = hidden · W4ᵀ
        const int ts = g4 % NT;
        if (mbar_test_u(bar_hb2_full + 8 * ts, (uint32_t)((g4 / NT) & 1))) {
          tc_fence_after();
          const uint32_t dd = tmem_base + ts * 128;
          if (elect_one()) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_tf32_ts(dd, dd + 64 + (uint32_t)(jj * 32 + kk * 8), desc_w4 + (uint64_t)((jj * (WBLK / 2)) >> 4) + 2 * kk,
                             idesc4, (uint32_t)((jj | kk) != 0));
            umma_commit(bar_d4_full + 8 * ts);
          }
          __syncwarp();
          if (lane == 0) R9_DBG(4, g4);
          ++g4;
          progress = true;
        }
      }
      if (g3 < g2) {  // output MLP, first Linear: D = grid' · W3ᵀ (grid' sits in the tile's slot)
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
                umma_tf32(dd, a0 + (uint64_t)((jj * BLK) >> 4) + 2 * kk, desc_w3 + (uint64_t)((jj * WBLK) >> 4) + 2 * kk, idesc,
                          (uint32_t)((jj | kk) != 0));
            umma_commit(bar_d3_full + 8 * ts);
          }
          __syncwarp();
          if (lane == 0) R9_DBG(3, g3);
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
          if (lane == 0) R9_DBG(2, g2);
          ++g2;
          progress = true;
        }
      }
      if (g1 < n_my && g1 < g4 + NT) {  // node MLP, first Linear (K = 128): D = [grid | aggr] · W1ᵀ
        const int ts = g1 % NT, rs = g1 % NRS;
        bool ready = mbar_test_u(bar_r_full + 8 * rs, (uint32_t)((g1 / NRS) & 1));
        if (ready) ready = mbar_test_u(bar_g_full, (uint32_t)(g1 & 1));
        // the TMEM stage must have been drained by the last epilogue of tile g1 - NT
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
          if (lane == 0) R9_DBG(1, g1);
          ++g1;
          progress = true;
        }
      }
      if (progress) idle = 0;
      else if (__nanosleep(40), ++idle > (1u << 24)) {
        if (lane == 0)
          printf("nlam tc_node_out: MMA issuer timeout (block %d g %d %d %d %d of %d)\n", blockIdx.x, g1, g2, g3, g4, n_my);
        __trap();
      }
    }
  } else if (warp >= W_E1 && warp < W_MMA) {
    // =============================== epilogue group 1: the SiLU of either MLP ===============================
    const bool lead = warp == W_E1;
    const int q = warp & 3;
    const int half = (warp - W_E1) >> 2;
    const int c0 = half * 32;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    {
      // SiLU(z) = h + h*tanh(h), h = z/2: W1 and W3 are halved in place once (exact)
      mbar_wait(bar_w, 0);
      float4* wq = reinterpret_cast<float4*>(smem + OFF_W1) + (tid - W_E1 * 32);
      for (int i = 0; i < 8; ++i) {  // 32 KB = 2048 float4 over 256 threads
        float4 x = wq[i * EPI];
        x.x *= 0.5f; x.y *= 0.5f; x.z *= 0.5f; x.w *= 0.5f;
        wq[i * EPI] = x;
      }
      float4* w3 = reinterpret_cast<float4*>(smem + OFF_W3) + (tid - W_E1 * 32);
      for (int i = 0; i < 4; ++i) {  // 16 KB
        float4 x = w3[i * EPI];
        x.x *= 0.5f; x.y *= 0.5f; x.z *= 0.5f; x.w *= 0.5f;
        w3[i * EPI] = x;
      }
      fence_proxy_async();
      mbar_arrive(bar_wscaled);
    }
    auto silu_stage = [&](int ts, const float* bias_h, uint32_t bar_done) {
      const uint32_t d1 = tmem_base + ts * 128 + t_lane + c0;
      float v[32];
      tmem_ld32(d1, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 bb = *reinterpret_cast<const float4*>(bias_h + c0 + 4 * i);
        const float2 h0 = add2(make_float2(v[4 * i], v[4 * i + 1]), make_float2(bb.x, bb.y));
        const float2 h1 = add2(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(bb.z, bb.w));
        const float2 o0 = fma2(h0, make_float2(tanh_fast(h0.x), tanh_fast(h0.y)), h0);
        const float2 o1 = fma2(h1, make_float2(tanh_fast(h1.x), tanh_fast(h1.y)), h1);
        v[4 * i] = o0.x;
        v[4 * i + 1] = o0.y;
        v[4 * i + 2] = o1.x;
        v[4 * i + 3] = o1.y;
      }
      tmem_st32(d1 + 64, v);
      tc_fence_before();
      mbar_arrive(bar_done + 8 * ts);
    };
    // Per iteration two stages are pending: the node MLP of tile i and the output MLP of tile i - 1; the group takes
    // whichever accumulator is ready first (the lead warp polls both barriers and publishes the choice).
    int round = 0;
    for (int i = 0; i <= n_my; ++i) {
      bool pend_a = i < n_my, pend_b = i >= 1;
      const int ti = i - 1;
      while (pend_a || pend_b) {
        if (lead) {
          int pick = -1;
          uint32_t spins = 0;
          while (pick < 0) {
            if (pend_b && mbar_test_u(bar_d3_full + 8 * (ti % NT), (uint32_t)((ti / NT) & 1))) pick = 1;
            else if (pend_a && mbar_test_u(bar_d1_full + 8 * (i % NT), (uint32_t)((i / NT) & 1))) pick = 0;
            else if (__nanosleep(20), ++spins > (1u << 24)) {
              if (lane == 0) printf("nlam tc_node_out: epilogue group 1 timeout (block %d tile %d)\n", blockIdx.x, i);
              __trap();
            }
          }
          if (lane == 0) sel[round & 1] = pick;
        }
        named_bar_sync(1, EPI);
        const int pick = sel[round & 1];
        ++round;
        tc_fence_after();
        if (pick == 0) {  // node MLP of tile i
          if (lead && lane == 0) R9_DBG(5, i);
          silu_stage(i % NT, s_b1h, bar_hb_full);
          pend_a = false;
        } else {  // output MLP of tile i - 1
          const int ts = ti % NT;
          if (lead && lane == 0 && p.ep_prev) {
            // the third GEMM has consumed grid': the slot takes the prev / boundary / mask slabs of the tile (contiguous
            // in global memory: one bulk copy each)
            const int w = blockIdx.x + ti * gridDim.x;
            const int b = w / p.n_tiles, t = w - b * p.n_tiles;
            const int nrows = (int)min(128LL, p.n_rows - (long long)t * 128);
            const uint32_t bytes = (uint32_t)(nrows * p.nout * 4);
            const uint32_t dst = sbase + OFF_R + (ti % NRS) * 2 * BLK;
            const long long g0 = ((long long)b * p.n_rows + (long long)t * 128) * p.nout;
            const uint32_t bar = bar_ep_full + 8 * ts;
            mbar_expect_tx(bar, p.ep_bnd ? 2u * bytes + (uint32_t)(nrows * 4) : bytes);
            bulk_load_1d(dst + S_PREV, p.ep_prev + g0, bytes, bar);
            if (p.ep_bnd) {
              bulk_load_1d(dst + S_BND, p.ep_bnd + g0, bytes, bar);
              bulk_load_1d(dst + S_MASK, p.ep_mask + (long long)t * 128, (uint32_t)(nrows * 4), bar);
            }
          }
          if (lead && lane == 0) R9_DBG(6, ti);
          silu_stage(ts, s_b3h, bar_hb2_full);
          pend_b = false;
        }
      }
    }
  } else if (warp < W_E1) {
    // =============================== epilogue group 2: LayerNorm + residual, then the narrow step epilogue ===============
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
    const int nout = p.nout;
    int round = 0;
    for (int i = 0; i <= n_my; ++i) {
     bool pend_a = i < n_my, pend_b = i >= 1;
     while (pend_a || pend_b) {
      {
        const int tb = i - 1;
        if (warp == 0) {  // the result epilogue of tile i - 1 first when it is ready: it returns the tile's slot
          int pick = -1;
          uint32_t spins = 0;
          while (pick < 0) {
            if (pend_b && mbar_test_u(bar_d4_full + 8 * (tb % NT), (uint32_t)((tb / NT) & 1)) &&
                (!p.ep_prev || mbar_test_u(bar_ep_full + 8 * (tb % NT), (uint32_t)((tb / NT) & 1))))
              pick = 1;
            else if (pend_a && mbar_test_u(bar_d2_full + 8 * (i % NT), (uint32_t)((i / NT) & 1))) pick = 0;
            else if (__nanosleep(20), ++spins > (1u << 24)) {
              if (lane == 0) printf("nlam tc_node_out: epilogue group 2 timeout (block %d tile %d)\n", blockIdx.x, i);
              __trap();
            }
          }
          if (lane == 0) sel[2 + (round & 1)] = pick;
        }
        named_bar_sync(2, EPI);
      }
      const int pick = sel[2 + (round & 1)];
      ++round;
      tc_fence_after();
      if (pick == 0) {
        // ---- node MLP of tile i: bias, LayerNorm, + grid (residual) -> grid' in place over the grid tile
        pend_a = false;
        const int ts = i % NT, rs = i % NRS;
        if (tid == 0) R9_DBG(7, i);
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
        // the two column halves of a row exchange (sum, sum of squares) through spare TMEM columns of the row's lane
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
        fence_proxy_async();  // generic writes -> tcgen05.mma operand reads
        tc_fence_before();
        mbar_arrive(bar_mid + 8 * ts);
      } else {
        // ---- output MLP of tile i - 1: this thread finishes columns [9*half, 9*half + 9) of its row
        pend_b = false;
        const int ti = i - 1;
        const int ts = ti % NT, rs = ti % NRS;
        if (tid == 0) R9_DBG(8, ti);
        const float* s_prev = reinterpret_cast<const float*>(smem + OFF_R + rs * 2 * BLK + S_PREV);
        const float* s_bnd = reinterpret_cast<const float*>(smem + OFF_R + rs * 2 * BLK + S_BND);
        float* s_out = reinterpret_cast<float*>(smem + OFF_R + rs * 2 * BLK + S_OUT);
        const float* s_mask = reinterpret_cast<const float*>(smem + OFF_R + rs * 2 * BLK + S_MASK);
        float vv[16];
        tmem_ld16(tmem_base + ts * 128 + t_lane + (half ? 8 : 0), vv);  // half 1: columns 8..23, it uses 9..17
        tc_fence_before();
        mbar_arrive(bar_d_free + 8 * ts);
        const float mk = (p.ep_prev && p.ep_bnd) ? s_mask[row] : 0.f;
        const float om = 1.0f - mk;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const int c = 9 * half + j;
          if (c < nout) {
            const float y = (half ? vv[j + 1] : vv[j]) + s_b4[c];
            float o = y;
            if (p.ep_prev) {
              const float pv = s_prev[row * nout + c];  // row pitch nout floats: conflict-free for odd nout
              const float bd = p.ep_bnd ? s_bnd[row * nout + c] : 0.f;
              o = fmaf(om * s_std[c], y, mk * bd + om * (pv + s_mean[c]));
            }
            s_out[row * nout + c] = o;
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_staged + 8 * ts);
      }
     }
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
bool tc_node_out_supported(const NlamMlp* node_mlp, const NlamMlp* out_mlp, const float* rec, int64_t rec_bs, const float* aggr,
                           int64_t n_rows, int B, const float* out, const StepEpilogue* ep) {
  static int on = -1;
  if (on < 0) on = getenv("NLAM_TC_NO_NODE_OUT") ? 0 : 1;
  if (!on) return false;
  int no1 = 0, no2 = 0;
  if (n_rows < 1 || n_rows >= (1LL << 31) - 256 || n_rows % 4 != 0) return false;
  if (!mlp_shape_ok(node_mlp, &no1) || no1 != 64 || node_mlp->in_dim != 128 || !node_mlp->ln_gamma || !node_mlp->ln_beta) return false;
  if (!mlp_shape_ok(out_mlp, &no2) || no2 > 18 || out_mlp->in_dim != 64 || out_mlp->ln_gamma) return false;
  if (!(aligned16(rec) && aligned16(aggr) && aligned16(out) && rec_bs % 4 == 0)) return false;
  if (ep && !(ep->prev && ep->std && ep->mean && aligned16(ep->prev) &&
              (!ep->boundary || (ep->mask && aligned16(ep->boundary) && aligned16(ep->mask)))))
    return false;
  (void)B;
  return true;
}

int tc_node_out(const NlamMlp* node_mlp, const NlamMlp* out_mlp, const float* rec, int64_t rec_bs, const float* aggr,
                int64_t n_rows, int B, float* out, const StepEpilogue* ep, cudaStream_t st) {
  NLAM_REQUIRE(tc_node_out_supported(node_mlp, out_mlp, rec, rec_bs, aggr, n_rows, B, out, ep), NLAM_E_UNSUPPORTED,
               "tc_node_out: unsupported shapes");
  const int nout = out_mlp->out_dim[1];
  CUtensorMap mr, mg, w1, w2, w3, w4;
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
  rc = make_map(&w3, out_mlp->w[0], 64, 64, 1, 64, 0, 64, false);
  if (rc) return rc;
  rc = make_map(&w4, out_mlp->w[1], 64, (uint64_t)nout, 1, 64, 0, 32, false);
  if (rc) return rc;
  NodeOutParams p;
  memset(&p, 0, sizeof(p));
  p.rec_batched = batched;
  p.b1 = node_mlp->b[0];
  p.b2 = node_mlp->b[1];
  p.gamma = node_mlp->ln_gamma;
  p.beta = node_mlp->ln_beta;
  p.eps = node_mlp->ln_eps;
  p.b3 = out_mlp->b[0];
  p.b4 = out_mlp->b[1];
  p.nout = nout;
  p.n_rows = n_rows;
  p.B = B;
  p.n_tiles = (int)((n_rows + 127) / 128);
  p.out = out;
  if (ep) {
    p.ep_prev = ep->prev;
    p.ep_bnd = ep->boundary;
    p.ep_mask = ep->mask;
    p.ep_std = ep->std;
    p.ep_mean = ep->mean;
  }
  static unsigned attr_mask = 0;
  int dev = 0;
  NLAM_CUDA_OK(cudaGetDevice(&dev));
  if (!(attr_mask & (1u << (dev & 31)))) {
    NLAM_CUDA_OK(cudaFuncSetAttribute(tc_node_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r9::SMEM));
    attr_mask |= 1u << (dev & 31);
  }
  const long long n_work = (long long)p.n_tiles * p.B;
  NLAM_REQUIRE(n_work < (1LL << 30), NLAM_E_UNSUPPORTED, "tc_node_out: too many work items");
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
    // algorithmic bytes: grid + aggregate rows in, the narrow result out, prev / boundary in, mask, weights
    const long long rows = (long long)B * n_rows;
    long long nb = 4LL * 64 * ((batched ? rows : n_rows) + rows) + 4LL * nout * rows;
    if (ep) nb += 4LL * nout * rows * (ep->boundary ? 2 : 1) + (ep->boundary ? 4LL * n_rows : 0);
    nb += 4LL * (64 * 128 + 64 * 64 + 64 * 64 + nout * 64 + 64 * 5 + nout);
    ProfScope ps("tc_node_out_kernel", st, nb);
    NLAM_CUDA_OK(launch_pdl(tc_node_out_kernel, grid, r9::THREADS, r9::SMEM, st, mr, mg, w1, w2, w3, w4, p));
  }
  count_launch();
  if (dbg_on) {
    long long h[256];
    NLAM_CUDA_OK(cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
    NLAM_CUDA_OK(cudaStreamSynchronize(st));
    long long t0 = h[0];
    fprintf(stderr, "[nlam tc_node_out timeline] grid=%d tiles=%lld (cycles rel. to first load)\n", grid, n_work);
    fprintf(stderr, " ti  ld_iss  g1_iss  g2_iss  g3_iss  g4_iss  siluA   siluB  lnA_beg outB_beg  stored\n");
    for (int it = 0; it < 16; ++it) {
      fprintf(stderr, "%3d ", it);
      for (int k = 0; k < 10; ++k) fprintf(stderr, "%7lld ", h[it * 16 + k] ? h[it * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

}  // namespace nlam
