// Row MLP, H = 64, dense 64-wide inputs:  out = [res +] LN(W2·SiLU(W1·[x0 | x1] + b1) + b2)
// (node update aggr_mlp([rec | aggr]) + rec of InteractionNet, reference gnn_layers.py:119-136; the grid
// encoder of BaseGraphModel.predict_step, reference base_graph_model.py:143-147).  Rows are independent, so
// this is a pure streaming kernel: one pass over the inputs, one over the output.
//
// Persistent CTAs, 128-row tiles.  A 5-slot ring of 32 KB operand tiles (one per source and tile) is filled
// by TMA; GEMM1 (K = 64 or 128, tcgen05 kind::tf32, accumulators in TMEM) reads the slots directly, epilogue 1
// (SiLU) writes the hidden tile back to TMEM as the A operand of GEMM2, epilogue 2 (bias, LayerNorm, residual)
// reads the residual from the source tile still sitting in its ring slot and writes the output tile IN PLACE
// over it, from where one TMA store moves it out (rows past the end are clipped) — no separate staging buffer.
// Three TMEM stages (D | hidden) keep three tiles in flight.
//
// 640 threads: warps 0-7 epilogue 2, warps 8-15 epilogue 1 (thread = row x 32 columns each), warp 16 MMA issue
// (whole warp in uniform control flow, one elected lane issues), warp 17 ring loader, warp 18 output stores.
//
// Two further modes of the same pipeline cover the rest of the grid side of a forecast step:
//   * NARROW INPUTS (grid embedder: prev | prev_prev | forcing | static = 17|17|18|4 columns, reference
//     graph/base.py:275-283): the 128-row slab of every source is contiguous in global memory, so one 1-D bulk
//     copy each lands it in a flat staging slot (ring slots 0/1); warps 17 and 19 repack it into the K-major,
//     128B-swizzled operand tile (ring slots 2-4) — the torch.cat is never materialised and no thread waits on a
//     global load;
//   * NARROW OUTPUT with the forecast-step epilogue (output_map 64 -> 64 -> 17, base.py:322-342 +
//     forecasters/autoregressive.py:128-131): second GEMM with N = 32, no LayerNorm; epilogue 2 prefetches
//     prev / boundary / mask for its elements BEFORE it waits for the accumulators, stages y as flat [128][17] in
//     the operand slot and writes  m*boundary + (1-m)*(prev + y*std + mean)  with coalesced stores.
#include "tc_ptx.cuh"

namespace nlam {

namespace r4 {
constexpr int THREADS = 640;
constexpr int EPI = 256;
constexpr int W_E1 = 8, W_MMA = 16, W_RING = 17, W_ST = 18;  // warp 19 idle
constexpr int NR = 5;   // ring slots
constexpr int NT = 3;   // TMEM stages
constexpr uint32_t BLK = 16384;
constexpr uint32_t WBLK = 8192;
constexpr uint32_t OFF_W1 = 0;                    // up to 4 blocks (K = 128)
constexpr uint32_t OFF_W2 = 4 * WBLK;             // 2 blocks
constexpr uint32_t OFF_RING = 6 * WBLK;           // NR x 32 KB
constexpr uint32_t OFF_MISC = OFF_RING + NR * 2 * BLK;
constexpr uint32_t SMEM = OFF_MISC + 2048;
}  // namespace r4

struct Row64Params {
  int n_src;      // 1 or 2
  int out_src;    // ring slot (source) that receives the output tile; also the residual when has_res
  int has_res;
  int batched[2];
  const float* b1;
  const float* b2;
  const float* gamma;
  const float* beta;
  float eps;
  long long n_rows;
  int B;
  int n_tiles;
  // narrow-input mode
  int in_narrow;  // number of narrow sources (0 = dense 64-wide sources through the tensor maps)
  const float* esrc[NLAM_MAX_SRC];
  long long ebs[NLAM_MAX_SRC];
  int edim[NLAM_MAX_SRC];
  int k_real;
  // narrow-output mode
  int out_narrow;  // 0 or the output width (< 64): no LayerNorm, no residual, plain stores
  float* out;
  const float* ep_prev;  // fused step epilogue (NULL: store y)
  const float* ep_bnd;
  const float* ep_mask;
  const float* ep_std;
  const float* ep_mean;
  long long* dbg;
};

#define R4_DBG(slot, it)                                                                  \
  do {                                                                                    \
    if (p.dbg && blockIdx.x == 0 && (it) < 16) p.dbg[(it) * 16 + (slot)] = clock64();     \
  } while (0)

__global__ void __launch_bounds__(r4::THREADS, 1)
tc_rowmlp64_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                   const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                   const __grid_constant__ CUtensorMap tmOut, const Row64Params p) {
  using namespace r4;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if ((sbase & 1023u) != 0) {
    if (tid == 0) printf("nlam tc_rowmlp64: dynamic shared memory not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t mb = sbase + OFF_MISC;
  const uint32_t bar_w = mb + 0;
  const uint32_t bar_wscaled = mb + 8;     // W1 halved in place (256 arrivals)
  const uint32_t bar_ring_full = mb + 16;  // [5]
  const uint32_t bar_ring_free = mb + 56;  // [5] GEMM1 commit, or the store thread for the output slot
  const uint32_t bar_d1_full = mb + 96;    // [3]
  const uint32_t bar_hb_full = mb + 120;   // [3] 256 arrivals
  const uint32_t bar_d2_full = mb + 144;   // [3]
  const uint32_t bar_d_free = mb + 168;    // [3] 256 arrivals
  const uint32_t bar_staged = mb + 192;    // [3] output tile written over its source slot (256 arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 216);
  const uint32_t bar_st_full = mb + 224;   // [2] narrow-input staging slot filled by the bulk copies
  const uint32_t bar_st_free = mb + 240;   // [2] ... and repacked (64 arrivals)
  float* sprm = reinterpret_cast<float*>(smem + OFF_MISC + 256);  // gamma | beta   (narrow output: std | mean)
  int2* ctab = reinterpret_cast<int2*>(smem + OFF_MISC + 1024);    // narrow inputs: column cc of the operand tile =
                                                                   // staging[ctab.x + row * ctab.y] (ctab.x < 0: zero)
  float* sb2 = reinterpret_cast<float*>(smem + OFF_MISC + 768);     // narrow output: b2 (zero past the output width)
  const uint32_t bar_ep_full = mb + 1536;  // [3] narrow output: prev / boundary / mask slabs of the tile landed (tx bytes)
  const int n_src = p.n_src;
  const int nb1 = 2 * n_src;

  if (warp == W_MMA) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_wscaled, EPI);
      for (int t = 0; t < NR; ++t) {
        mbar_init(bar_ring_full + 8 * t, p.in_narrow ? 2 * EPI : 1);  // repack threads / TMA transactions
        mbar_init(bar_ring_free + 8 * t, 1);
      }
      for (int t = 0; t < 2; ++t) {
        mbar_init(bar_st_full + 8 * t, 1);
        mbar_init(bar_st_free + 8 * t, 2 * EPI);
      }
      for (int t = 0; t < NT; ++t) {
        mbar_init(bar_d1_full + 8 * t, 1);
        mbar_init(bar_hb_full + 8 * t, EPI);
        mbar_init(bar_d2_full + 8 * t, 1);
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
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
  }
  pdl_launch_dependents();
  pdl_wait();  // everything below may read what the previous kernel in the stream wrote
  if (tid < 64 && p.in_narrow) {
    int base = 0, col = 0, found = 0;
    int2 e = make_int2(-1, 0);
    for (int sidx = 0; sidx < p.in_narrow; ++sidx) {
      const int d = p.edim[sidx];
      if (!found && tid < col + d) {
        e = make_int2(base + (tid - col), d);
        found = 1;
      }
      col += d;
      base += 128 * d;
    }
    ctab[tid] = e;
  }
  if (tid < 64) sb2[tid] = (p.out_narrow && tid < p.out_narrow) ? p.b2[tid] : 0.f;
  if (tid < 64) {
    if (p.out_narrow) {
      sprm[tid] = (p.ep_prev && tid < p.out_narrow) ? p.ep_std[tid] : 1.f;
      sprm[64 + tid] = (p.ep_prev && tid < p.out_narrow) ? p.ep_mean[tid] : 0.f;
    } else {
      sprm[tid] = p.gamma[tid];
      sprm[64 + tid] = p.beta[tid];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  // TMEM columns: stage ts: D at ts*128 (first, then second GEMM), hidden at +64; LayerNorm scratch at 384
  const int n_work = p.n_tiles * p.B;
  int n_my = 0;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) ++n_my;
  // ring slot / phase of source s of tile ti (narrow inputs: one operand tile per tile in slots 2-4)
  auto op_slot = [&](int ti, int s) -> int { return p.in_narrow ? 2 + ti % 3 : (ti * n_src + s) % NR; };
  auto op_phase = [&](int ti, int s) -> uint32_t {
    return (uint32_t)(p.in_narrow ? (ti / 3) & 1 : ((ti * n_src + s) / NR) & 1);
  };
  // narrow inputs: BOTH epilogue groups repack tile ti from the flat staging slot into the K-major swizzled operand tile: group
  // grp fills the 32-column block grp (thread = row x four 16-byte chunks; columns k_real..63 are zero).  History: two dedicated
  // warps (450 instructions per thread and tile in series, 6.2 k cycles per tile: 351 us), then the epilogue-1 group alone
  // (SiLU 1.3 k + repack 1.4 k cycles per tile while epilogue 2 idled half the time: 219 us).
  auto repack_tile = [&](int ti, int gt, bool lead, int grp) {
    const int st = ti & 1;
    const int slot = op_slot(ti, 0);
    if (lead) {
      mbar_wait(bar_st_full + 8 * st, (uint32_t)((ti >> 1) & 1));
      mbar_wait(bar_ring_free + 8 * slot, op_phase(ti, 0) ^ 1u);
    }
    named_bar_sync(1 + grp, EPI);
    uint8_t* tile = smem + OFF_RING + slot * 2 * BLK;
    const float* stg = reinterpret_cast<const float*>(smem + OFF_RING + st * 2 * BLK);
    const int row = gt & 127, hh = grp, c4 = 4 * (gt >> 7);
    const int rx = row & 7;
#pragma unroll
    for (int cq = 0; cq < 4; ++cq) {  // one 16-byte chunk of the operand row per store
      const int ch = c4 + cq;
      float4 o;
      float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int2 e = ctab[32 * hh + 4 * ch + u];  // broadcast read
        ov[u] = (e.x >= 0) ? stg[e.x + row * e.y] : 0.f;  // lane stride d floats: conflict-free for odd d
      }
      *reinterpret_cast<float4*>(tile + hh * BLK + row * 128 + ((ch ^ rx) << 4)) = o;
    }
    fence_proxy_async();  // generic writes -> tcgen05.mma operand reads
    mbar_arrive(bar_ring_full + 8 * slot);
    mbar_arrive(bar_st_free + 8 * st);
  };

  if (warp == W_RING) {
    // =============================== operand ring ===============================
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();
      const uint64_t pol_keep = policy_evict_last();
      mbar_expect_tx(bar_w, (uint32_t)nb1 * WBLK + (p.out_narrow ? WBLK : 2u * WBLK));  // W2 box: 32 or 64 rows
      for (int kb = 0; kb < nb1; ++kb) tma_load_2d(sbase + OFF_W1 + kb * WBLK, &tmW1, bar_w, 32 * kb, 0);
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(sbase + OFF_W2 + kb * WBLK, &tmW2, bar_w, 32 * kb, 0);
      int i = 0;
      for (int ti = 0; !p.in_narrow && ti < n_my; ++ti) {
        const int w = blockIdx.x + ti * gridDim.x;
        const int b = w / p.n_tiles, t = w - b * p.n_tiles;
        for (int s = 0; s < n_src; ++s, ++i) {
          const int slot = i % NR;
          const uint32_t dst = sbase + OFF_RING + slot * 2 * BLK;
          const uint32_t full = bar_ring_full + 8 * slot;
          mbar_wait(bar_ring_free + 8 * slot, (uint32_t)(((i / NR) & 1) ^ 1));
          mbar_expect_tx(full, 2u * BLK);
          if (s == 0) R4_DBG(0, ti);
          const CUtensorMap* map = s ? &tmA1 : &tmA0;
          const uint64_t pol = p.batched[s] ? pol_stream : pol_keep;
          tma_load_3d(dst, map, full, 0, t * 128, p.batched[s] ? b : 0, pol);
          tma_load_3d(dst + BLK, map, full, 32, t * 128, p.batched[s] ? b : 0, pol);
        }
      }
    }
    if (p.in_narrow && lane == 0) {
      // ---- narrow inputs: bulk copies into the two staging slots (ring slots 0/1), as far ahead as they are free
      for (int ti = 0; ti < n_my; ++ti) {
        const int w = blockIdx.x + ti * gridDim.x;
        const int b = w / p.n_tiles, t = w - b * p.n_tiles;
        const int st = ti & 1;
        const int nrows = (int)min(128LL, p.n_rows - (long long)t * 128);
        mbar_wait(bar_st_free + 8 * st, (uint32_t)(((ti >> 1) & 1) ^ 1));
        mbar_expect_tx(bar_st_full + 8 * st, (uint32_t)(nrows * p.k_real * 4));
        uint32_t dst = sbase + OFF_RING + st * 2 * BLK;
        for (int sidx = 0; sidx < p.in_narrow; ++sidx) {
          const int d = p.edim[sidx];
          bulk_load_1d(dst, p.esrc[sidx] + (long long)b * p.ebs[sidx] + (long long)t * 128 * d, (uint32_t)(nrows * d * 4),
                       bar_st_full + 8 * st);
          dst += (uint32_t)(128 * d * 4);
        }
        R4_DBG(0, ti);
      }
    }
  } else if (warp == W_ST) {
    // =============================== output stores ===============================
    if (lane == 0 && !p.out_narrow) {
      for (int ti = 0; ti < n_my; ++ti) {
        const int w = blockIdx.x + ti * gridDim.x;
        const int b = w / p.n_tiles, t = w - b * p.n_tiles;
        const int ts = ti % NT;
        const int slot = op_slot(ti, p.out_src);
        mbar_wait(bar_staged + 8 * ts, (uint32_t)((ti / NT) & 1));
        const uint32_t src = sbase + OFF_RING + slot * 2 * BLK;
        tma_store_3d(&tmOut, src, 0, t * 128, b);
        tma_store_3d(&tmOut, src + BLK, 32, t * 128, b);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar_ring_free + 8 * slot);
        R4_DBG(7, ti);
      }
      bulk_wait0();
    }
    if (lane == 0 && p.out_narrow) {
      for (int ti = 0; ti < n_my; ++ti) {
        const int w = blockIdx.x + ti * gridDim.x;
        const int b = w / p.n_tiles, t = w - b * p.n_tiles;
        const int ts = ti % NT;
        const int slot = op_slot(ti, p.out_src);
        const int nrows = (int)min(128LL, p.n_rows - (long long)t * 128);
        mbar_wait(bar_staged + 8 * ts, (uint32_t)((ti / NT) & 1));
        bulk_store_1d(p.out + ((long long)b * p.n_rows + (long long)t * 128) * p.out_narrow,
                      sbase + OFF_RING + slot * 2 * BLK + 18432, (uint32_t)(nrows * p.out_narrow * 4));
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar_ring_free + 8 * slot);
        R4_DBG(7, ti);
      }
      bulk_wait0();
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issue (uniform control flow, one elected lane) ===============================
    const uint32_t idesc = umma_idesc_tf32(128, 64);
    const uint32_t idesc2 = p.out_narrow ? umma_idesc_tf32(128, 32) : idesc;
    mbar_wait(bar_w, 0);
    mbar_wait(bar_wscaled, 0);
    tc_fence_after();
    const uint64_t desc_w1 = umma_desc(sbase + OFF_W1);
    const uint64_t desc_w2 = umma_desc(sbase + OFF_W2);
    const uint64_t desc_ring = umma_desc(sbase + OFF_RING);
    int g1 = 0, g2 = 0;
    uint32_t idle = 0;
    while (g2 < n_my) {
      bool progress = false;
      if (g1 < n_my && g1 <= g2 + 2) {
        const int ts = g1 % NT;
        bool ready = mbar_test_u(bar_d_free + 8 * ts, (uint32_t)(((g1 / NT) & 1) ^ 1));
        for (int s = 0; s < n_src && ready; ++s) ready = mbar_test_u(bar_ring_full + 8 * op_slot(g1, s), op_phase(g1, s));
        if (ready) {
          tc_fence_after();
          if (lane == 0) R4_DBG(1, g1);
          const uint32_t dd = tmem_base + ts * 128;
          if (elect_one()) {
            for (int s = 0; s < n_src; ++s) {
              const int slot = op_slot(g1, s);
              const uint64_t a0 = desc_ring + (uint64_t)((slot * 2 * BLK) >> 4);
              const uint64_t b0 = desc_w1 + (uint64_t)((s * 2 * WBLK) >> 4);
#pragma unroll
              for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_tf32(dd, a0 + (uint64_t)((jj * BLK) >> 4) + 2 * kk, b0 + (uint64_t)((jj * WBLK) >> 4) + 2 * kk, idesc,
                            (uint32_t)((s | jj | kk) != 0));
            }
            umma_commit(bar_d1_full + 8 * ts);
            for (int s = 0; s < n_src; ++s)
              if (s != p.out_src) umma_commit(bar_ring_free + 8 * op_slot(g1, s));
          }
          __syncwarp();
          if (lane == 0) R4_DBG(2, g1);
          ++g1;
          progress = true;
        }
      }
      if (g2 < g1) {
        const int ts = g2 % NT;
        if (mbar_test_u(bar_hb_full + 8 * ts, (uint32_t)((g2 / NT) & 1))) {
          tc_fence_after();
          const uint32_t dd = tmem_base + ts * 128;
          const uint32_t ht = dd + 64;
          if (elect_one()) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_tf32_ts(dd, ht + (uint32_t)(jj * 32 + kk * 8), desc_w2 + (uint64_t)((jj * WBLK) >> 4) + 2 * kk, idesc2,
                             (uint32_t)((jj | kk) != 0));
            umma_commit(bar_d2_full + 8 * ts);
          }
          __syncwarp();
          if (lane == 0) R4_DBG(3, g2);
          ++g2;
          progress = true;
        }
      }
      if (progress) idle = 0;
      else if (__nanosleep(40), ++idle > (1u << 24)) {
        if (lane == 0) printf("nlam tc_rowmlp64: MMA issuer timeout (block %d g1 %d g2 %d)\n", blockIdx.x, g1, g2);
        __trap();
      }
    }
  } else if (warp >= W_E1 && warp < W_MMA) {
    // =============================== epilogue 1: hidden = SiLU(D1 + b1) ===============================
    const bool lead = warp == W_E1;
    const int q = warp & 3;
    const int half = (warp - W_E1) >> 2;
    const int c0 = half * 32;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    // SiLU(z) = h + h*tanh(h), h = z/2: W1 is halved in place once (exact), b1/2 lives in registers
    float2 bh[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bh[i] = make_float2(0.5f * __ldg(p.b1 + c0 + 2 * i), 0.5f * __ldg(p.b1 + c0 + 2 * i + 1));
    {
      mbar_wait(bar_w, 0);
      float4* wq = reinterpret_cast<float4*>(smem + OFF_W1) + (tid - W_E1 * 32);
      for (int i = 0; i < 2 * nb1; ++i) {  // nb1 x 8 KB = nb1 x 512 float4 over 256 threads
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
    const int gt1 = tid - W_E1 * 32;
    if (p.in_narrow) {  // operand tiles are repacked two tiles ahead of their SiLU
      if (n_my > 0) repack_tile(0, gt1, lead, 0);
      if (n_my > 1) repack_tile(1, gt1, lead, 0);
    }
    for (int ti = 0; ti < n_my; ++ti) {
      const int ts = ti % NT;
      if (lead) mbar_wait(bar_d1_full + 8 * ts, (uint32_t)((ti / NT) & 1));
      if (lead && lane == 0 && p.out_narrow && p.ep_prev) {
        // narrow output + step epilogue: the first GEMM has consumed the operand tile, so its slot can already take
        // the prev / boundary / mask slabs of this tile (contiguous in global memory: one bulk copy each)
        const int w = blockIdx.x + ti * gridDim.x;
        const int b = w / p.n_tiles, t = w - b * p.n_tiles;
        const int nrows = (int)min(128LL, p.n_rows - (long long)t * 128);
        const uint32_t bytes = (uint32_t)(nrows * p.out_narrow * 4);
        const uint32_t dst = sbase + OFF_RING + op_slot(ti, p.out_src) * 2 * BLK;
        const long long g0 = ((long long)b * p.n_rows + (long long)t * 128) * p.out_narrow;
        const uint32_t bar = bar_ep_full + 8 * ts;
        mbar_expect_tx(bar, p.ep_bnd ? 2u * bytes + (uint32_t)(nrows * 4) : bytes);
        bulk_load_1d(dst, p.ep_prev + g0, bytes, bar);
        if (p.ep_bnd) {
          bulk_load_1d(dst + 9216, p.ep_bnd + g0, bytes, bar);
          bulk_load_1d(dst + 27648, p.ep_mask + (long long)t * 128, (uint32_t)(nrows * 4), bar);
        }
      }
      named_bar_sync(1, EPI);
      tc_fence_after();
      if (lead && lane == 0) R4_DBG(4, ti);
      const uint32_t d1 = tmem_base + ts * 128 + t_lane + c0;
      float v[32];
      tmem_ld32(d1, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 h = add2(make_float2(v[2 * i], v[2 * i + 1]), bh[i]);
        const float2 o = fma2(h, make_float2(tanh_fast(h.x), tanh_fast(h.y)), h);
        v[2 * i] = o.x;
        v[2 * i + 1] = o.y;
      }
      tmem_st32(d1 + 64, v);
      tc_fence_before();
      mbar_arrive(bar_hb_full + 8 * ts);
      if (p.in_narrow && ti + 2 < n_my) repack_tile(ti + 2, gt1, lead, 0);
    }
  } else if (warp < W_E1) {
    // =============================== epilogue 2: bias, LayerNorm, residual; output in place over the source tile ===
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
    for (int i = 0; i < 16; ++i) {
      const int c = c0 + 2 * i;
      const int nb = p.out_narrow ? p.out_narrow : 64;
      b2r[i] = make_float2(c < nb ? __ldg(p.b2 + c) : 0.f, c + 1 < nb ? __ldg(p.b2 + c + 1) : 0.f);
    }
    if (p.in_narrow) {  // this group's column block of the first two operand tiles
      if (n_my > 0) repack_tile(0, tid, warp == 0, 1);
      if (n_my > 1) repack_tile(1, tid, warp == 0, 1);
    }
    for (int ti = 0; ti < n_my; ++ti) {
      const int ts = ti % NT;
      const int slot = op_slot(ti, p.out_src);
      if (p.out_narrow) {
        // ---- narrow output (+ forecast-step epilogue), bulk-copy form: prev / boundary / mask of the tile sit in the
        // (dead) operand slot, this thread finishes columns [9*half, 9*half + 9) of its row into the flat output slab,
        // which leaves by one bulk store — no global access and no index arithmetic in the epilogue threads (the
        // element-wise form below executes 8 k warp instructions per tile, this one ~1 k)
        const int nout = p.out_narrow;
        if (warp == 0) {
          mbar_wait(bar_d2_full + 8 * ts, (uint32_t)((ti / NT) & 1));
          if (p.ep_prev) mbar_wait(bar_ep_full + 8 * ts, (uint32_t)((ti / NT) & 1));
        }
        named_bar_sync(2, EPI);
        tc_fence_after();
        if (tid == 0) R4_DBG(5, ti);
        const float* s_prev = reinterpret_cast<const float*>(smem + OFF_RING + slot * 2 * BLK);
        const float* s_bnd = s_prev + 2304;
        float* s_out = const_cast<float*>(s_prev) + 4608;
        const float* s_mask = s_prev + 6912;
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
            const float y = (half ? vv[j + 1] : vv[j]) + sb2[c];
            float o = y;
            if (p.ep_prev) {
              const float pv = s_prev[row * nout + c];  // row pitch nout floats: conflict-free for odd nout
              const float bd = p.ep_bnd ? s_bnd[row * nout + c] : 0.f;
              o = fmaf(om * sprm[c], y, mk * bd + om * (pv + sprm[64 + c]));
            }
            s_out[row * nout + c] = o;
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_staged + 8 * ts);
        if (tid == 0) R4_DBG(6, ti);
        if (p.in_narrow && ti + 2 < n_my) repack_tile(ti + 2, tid, warp == 0, 1);
        continue;
      }
      if (warp == 0) mbar_wait(bar_d2_full + 8 * ts, (uint32_t)((ti / NT) & 1));
      named_bar_sync(2, EPI);
      tc_fence_after();
      if (tid == 0) R4_DBG(5, ti);
      float vf[32];
      tmem_ld32(tmem_base + ts * 128 + t_lane + c0, vf);
      tc_fence_before();
      mbar_arrive(bar_d_free + 8 * ts);
      float2 v[16];
      float2 sm2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = add2(make_float2(vf[2 * i], vf[2 * i + 1]), b2r[i]);
        sm2 = add2(sm2, v[i]);
        sq2 = fma2(v[i], v[i], sq2);
      }
      // the two column halves of a row exchange (sum, sum of squares) through spare TMEM columns of the row's lane;
      // scratch is double-buffered by tile parity so one 64-thread barrier per tile suffices
      const uint32_t scr = ln_col + 4 * (ti & 1);
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
      uint8_t* orow = smem + OFF_RING + slot * 2 * BLK + half * BLK + row * 128;
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        const float4 g4 = *reinterpret_cast<const float4*>(sprm + c0 + 4 * k8);
        const float4 b4 = *reinterpret_cast<const float4*>(sprm + 64 + c0 + 4 * k8);
        float4* ptr = reinterpret_cast<float4*>(orow + ((k8 ^ rx) << 4));
        float2 o0 = fma2(fma2(v[2 * k8], rs2, nm2), make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
        float2 o1 = fma2(fma2(v[2 * k8 + 1], rs2, nm2), make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
        if (p.has_res) {
          const float4 r4v = *ptr;
          o0 = add2(o0, make_float2(r4v.x, r4v.y));
          o1 = add2(o1, make_float2(r4v.z, r4v.w));
        }
        *ptr = make_float4(o0.x, o0.y, o1.x, o1.y);
      }
      fence_proxy_async();
      mbar_arrive(bar_staged + 8 * ts);
      if (tid == 0) R4_DBG(6, ti);
      if (p.in_narrow && ti + 2 < n_my) repack_tile(ti + 2, tid, warp == 0, 1);
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
static bool row64_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("NLAM_TC_ROW");
    on = (e && e[0] == 'v' && e[1] == '1') ? 0 : 1;
  }
  return on != 0;
}

// input side: 1-2 dense 64-wide sources (TMA ring), or up to 4 narrow sources whose widths sum to in_dim <= 64 (bulk
// copies: 16-byte aligned slabs).  Returns 0 (unsupported), 1 (dense), 2 (narrow).
static int row64_input_kind(const NlamMlp* mlp, const NlamRowSrc* srcs, int n_src, int64_t n_rows) {
  if (n_src < 1 || n_src > NLAM_MAX_SRC) return 0;
  bool wide = n_src <= 2 && mlp->in_dim == 64 * n_src;
  for (int s = 0; s < n_src && wide; ++s)
    wide = srcs[s].dim == 64 && !srcs[s].idx && aligned16(srcs[s].ptr) && srcs[s].bstride % 4 == 0;
  if (wide) return 1;
  if (mlp->in_dim > 64 || n_rows % 4 != 0) return 0;
  int k = 0;
  for (int s = 0; s < n_src; ++s) {
    if (srcs[s].idx || !aligned16(srcs[s].ptr) || (srcs[s].bstride * 4) % 16 != 0 || srcs[s].dim < 1) return 0;
    k += srcs[s].dim;
  }
  return k == mlp->in_dim ? 2 : 0;
}

bool tc_rowmlp64_supported(const NlamMlp* mlp, const NlamRowSrc* srcs, int n_src, const NlamRowSrc* res, int64_t n_rows) {
  if (!row64_enabled()) return false;
  int nout = 0;
  if (n_rows < 1 || n_rows >= (1LL << 31) - 256) return false;
  if (!mlp_shape_ok(mlp, &nout) || nout != 64 || !mlp->ln_gamma || !mlp->ln_beta) return false;
  const int kind = row64_input_kind(mlp, srcs, n_src, n_rows);
  if (!kind) return false;
  if (res) {
    if (res->idx || kind != 1) return false;
    bool match = false;
    for (int s = 0; s < n_src; ++s) match = match || (srcs[s].ptr == res->ptr && srcs[s].bstride == res->bstride);
    if (!match) return false;
  }
  return true;
}

// narrow output (<= 18 columns, no LayerNorm, no residual), optionally with the forecast-step epilogue.  The output slab
// of a tile (and prev / boundary / mask) travels by 1-D bulk copies: rows a multiple of 4, 16-byte aligned tensors.
bool tc_rowmlp_narrow_out_supported(const NlamMlp* mlp, const NlamRowSrc* srcs, int n_src, int64_t n_rows, const float* out,
                                    const StepEpilogue* ep) {
  if (!row64_enabled()) return false;
  int nout = 0;
  if (n_rows < 1 || n_rows >= (1LL << 31) - 256) return false;
  if (!mlp_shape_ok(mlp, &nout) || nout > 18 || mlp->ln_gamma) return false;
  if (n_rows % 4 != 0 || !aligned16(out)) return false;
  if (ep && !(aligned16(ep->prev) && (!ep->boundary || (aligned16(ep->boundary) && aligned16(ep->mask))))) return false;
  return row64_input_kind(mlp, srcs, n_src, n_rows) != 0;
}

int tc_rowmlp64(const NlamMlp* mlp, const NlamRowSrc* srcs, int n_src, const NlamRowSrc* res, float* out, int64_t n_rows,
                int B, cudaStream_t st, const StepEpilogue* ep) {
  int nout = 0;
  NLAM_REQUIRE(mlp_shape_ok(mlp, &nout), NLAM_E_UNSUPPORTED, "tc_rowmlp64: unsupported MLP shape");
  const bool narrow_out = nout < 64;
  NLAM_REQUIRE(narrow_out || aligned16(out), NLAM_E_INVALID, "tc_rowmlp64: output not 16-byte aligned");
  NLAM_REQUIRE(!ep || narrow_out, NLAM_E_UNSUPPORTED, "tc_rowmlp64: the step epilogue needs a narrow output");
  const int kind = row64_input_kind(mlp, srcs, n_src, n_rows);
  NLAM_REQUIRE(kind != 0, NLAM_E_UNSUPPORTED, "tc_rowmlp64: unsupported inputs");
  Row64Params p;
  memset(&p, 0, sizeof(p));
  CUtensorMap a[2], w1, w2, om;
  memset(a, 0, sizeof(a));
  memset(&om, 0, sizeof(om));
  p.out_src = 0;
  int rc;
  if (kind == 1) {
    p.n_src = n_src;
    for (int s = 0; s < n_src; ++s) {
      const bool batched = srcs[s].bstride != 0 && B > 1;
      p.batched[s] = batched;
      rc = make_map(&a[s], srcs[s].ptr, 64, (uint64_t)n_rows, batched ? (uint64_t)B : 1, 64,
                    batched ? (uint64_t)srcs[s].bstride : (uint64_t)n_rows * 64, 128, true);
      if (rc) return rc;
    }
    if (n_src == 1) a[1] = a[0];
  } else {
    p.n_src = 1;  // one repacked operand tile per row tile
    p.in_narrow = n_src;
    p.k_real = mlp->in_dim;
    for (int s = 0; s < n_src; ++s) {
      p.esrc[s] = srcs[s].ptr;
      p.ebs[s] = (B > 1) ? srcs[s].bstride : 0;
      p.edim[s] = srcs[s].dim;
    }
  }
  if (res) {
    p.has_res = 1;
    for (int s = n_src - 1; s >= 0; --s)
      if (srcs[s].ptr == res->ptr && srcs[s].bstride == res->bstride) p.out_src = s;
  }
  // W1: (64, in_dim) row-major; K blocks past in_dim (narrow inputs, in_dim < 64) are zero-filled by TMA
  rc = make_map(&w1, mlp->w[0], (uint64_t)mlp->in_dim, 64, 1, (uint64_t)mlp->in_dim, 0, 64, false);
  if (rc) return rc;
  rc = make_map(&w2, mlp->w[1], 64, (uint64_t)nout, 1, 64, 0, narrow_out ? 32 : 64, false);
  if (rc) return rc;
  if (kind == 2) a[0] = a[1] = w1;  // unused tensor-map slots must still be valid objects
  if (!narrow_out) {
    rc = make_map(&om, out, 64, (uint64_t)n_rows, (uint64_t)B, 64, (uint64_t)n_rows * 64, 128, true);
    if (rc) return rc;
  } else {
    om = w1;
    p.out_narrow = nout;
    p.out = out;
    if (ep) {
      p.ep_prev = ep->prev;
      p.ep_bnd = ep->boundary;
      p.ep_mask = ep->mask;
      p.ep_std = ep->std;
      p.ep_mean = ep->mean;
    }
  }
  p.b1 = mlp->b[0];
  p.b2 = mlp->b[1];
  p.gamma = mlp->ln_gamma;
  p.beta = mlp->ln_beta;
  p.eps = mlp->ln_eps;
  p.n_rows = n_rows;
  p.B = B;
  p.n_tiles = (int)((n_rows + 127) / 128);
  static unsigned attr_mask = 0;
  int dev = 0;
  NLAM_CUDA_OK(cudaGetDevice(&dev));
  if (!(attr_mask & (1u << (dev & 31)))) {
    NLAM_CUDA_OK(cudaFuncSetAttribute(tc_rowmlp64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r4::SMEM));
    attr_mask |= 1u << (dev & 31);
  }
  const long long n_work = (long long)p.n_tiles * p.B;
  NLAM_REQUIRE(n_work < (1LL << 30), NLAM_E_UNSUPPORTED, "tc_rowmlp64: too many work items");
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
    ProfScope ps(ep ? "tc_rowmlp64_kernel(step)" : (kind == 2 ? "tc_rowmlp64_kernel(narrow-in)" : "tc_rowmlp64_kernel"), st,
                 rowmlp_algorithmic_bytes(mlp, srcs, n_src, res, nullptr, n_rows, B, false, ep));
    NLAM_CUDA_OK(launch_pdl(tc_rowmlp64_kernel, grid, r4::THREADS, r4::SMEM, st, a[0], a[1], w1, w2, om, p));
  }
  count_launch();
  if (dbg_on) {
    long long h[256];
    NLAM_CUDA_OK(cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
    NLAM_CUDA_OK(cudaStreamSynchronize(st));
    long long t0 = h[0];
    fprintf(stderr, "[nlam tc_rowmlp64 timeline] grid=%d tiles=%lld n_src=%d (cycles rel. to first load)\n", grid, n_work, n_src);
    fprintf(stderr, " ti  ld_iss  g1_beg  g1_iss  g2_iss e1_start e2_start e2_done  stored\n");
    for (int it = 0; it < 16; ++it) {
      fprintf(stderr, "%3d ", it);
      for (int k = 0; k < 8; ++k) fprintf(stderr, "%7lld ", h[it * 16 + k] ? h[it * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

}  // namespace nlam
