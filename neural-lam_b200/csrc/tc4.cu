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
#include "tc_ptx.cuh"

namespace nlam {

namespace r4 {
constexpr int THREADS = 640;
constexpr int EPI = 256;
constexpr int W_E1 = 8, W_MMA = 16, W_RING = 17, W_ST = 18;
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
  long long* dbg;
};

#define R4_DBG(slot, it)                                                                  \
  do {                                                                                    \
    if (p.dbg && blockIdx.x == 0 && (it) < 16) p.dbg[(it) * 16 + (slot)] = clock64();     \
  } while (0)

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// warp-uniform barrier test (every lane tests; lane 0's answer is taken so the compiler sees a uniform value)
__device__ __forceinline__ bool mbar_test_u(uint32_t bar, uint32_t parity) {
  return __shfl_sync(0xffffffffu, (int)mbar_test(bar, parity), 0) != 0;
}

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
  float* sprm = reinterpret_cast<float*>(smem + OFF_MISC + 256);  // gamma | beta
  const int n_src = p.n_src;
  const int nb1 = 2 * n_src;

  if (warp == W_MMA) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_wscaled, EPI);
      for (int t = 0; t < NR; ++t) {
        mbar_init(bar_ring_full + 8 * t, 1);
        mbar_init(bar_ring_free + 8 * t, 1);
      }
      for (int t = 0; t < NT; ++t) {
        mbar_init(bar_d1_full + 8 * t, 1);
        mbar_init(bar_hb_full + 8 * t, EPI);
        mbar_init(bar_d2_full + 8 * t, 1);
        mbar_init(bar_d_free + 8 * t, EPI);
        mbar_init(bar_staged + 8 * t, EPI);
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
  if (tid < 64) {
    sprm[tid] = p.gamma[tid];
    sprm[64 + tid] = p.beta[tid];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  // TMEM columns: stage ts: D at ts*128 (first, then second GEMM), hidden at +64; LayerNorm scratch at 384
  const int n_work = p.n_tiles * p.B;
  int n_my = 0;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) ++n_my;

  if (warp == W_RING) {
    // =============================== operand ring ===============================
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();
      const uint64_t pol_keep = policy_evict_last();
      mbar_expect_tx(bar_w, (uint32_t)(nb1 + 2) * WBLK);
      for (int kb = 0; kb < nb1; ++kb) tma_load_2d(sbase + OFF_W1 + kb * WBLK, &tmW1, bar_w, 32 * kb, 0);
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(sbase + OFF_W2 + kb * WBLK, &tmW2, bar_w, 32 * kb, 0);
      int i = 0;
      for (int ti = 0; ti < n_my; ++ti) {
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
  } else if (warp == W_ST) {
    // =============================== output stores ===============================
    if (lane == 0) {
      for (int ti = 0; ti < n_my; ++ti) {
        const int w = blockIdx.x + ti * gridDim.x;
        const int b = w / p.n_tiles, t = w - b * p.n_tiles;
        const int ts = ti % NT;
        const int slot = (ti * n_src + p.out_src) % NR;
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
  } else if (warp == W_MMA) {
    // =============================== MMA issue (uniform control flow, one elected lane) ===============================
    const uint32_t idesc = umma_idesc_tf32(128, 64);
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
        const int i0 = g1 * n_src;
        bool ready = mbar_test_u(bar_d_free + 8 * ts, (uint32_t)(((g1 / NT) & 1) ^ 1));
        for (int s = 0; s < n_src && ready; ++s)
          ready = mbar_test_u(bar_ring_full + 8 * ((i0 + s) % NR), (uint32_t)(((i0 + s) / NR) & 1));
        if (ready) {
          tc_fence_after();
          if (lane == 0) R4_DBG(1, g1);
          const uint32_t dd = tmem_base + ts * 128;
          if (elect_one()) {
            for (int s = 0; s < n_src; ++s) {
              const int slot = (i0 + s) % NR;
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
              if (s != p.out_src) umma_commit(bar_ring_free + 8 * ((i0 + s) % NR));
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
                umma_tf32_ts(dd, ht + (uint32_t)(jj * 32 + kk * 8), desc_w2 + (uint64_t)((jj * WBLK) >> 4) + 2 * kk, idesc,
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
    for (int ti = 0; ti < n_my; ++ti) {
      const int ts = ti % NT;
      if (lead) mbar_wait(bar_d1_full + 8 * ts, (uint32_t)((ti / NT) & 1));
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
    for (int i = 0; i < 16; ++i) b2r[i] = make_float2(__ldg(p.b2 + c0 + 2 * i), __ldg(p.b2 + c0 + 2 * i + 1));
    for (int ti = 0; ti < n_my; ++ti) {
      const int ts = ti % NT;
      const int slot = (ti * n_src + p.out_src) % NR;
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
bool tc_rowmlp64_supported(const NlamMlp* mlp, const NlamRowSrc* srcs, int n_src, const NlamRowSrc* res, int64_t n_rows) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("NLAM_TC_ROW");
    on = (e && e[0] == 'v' && e[1] == '1') ? 0 : 1;
  }
  if (!on) return false;
  int nout = 0;
  if (n_src < 1 || n_src > 2 || n_rows < 1 || n_rows >= (1LL << 31) - 256) return false;
  if (!mlp_shape_ok(mlp, &nout) || nout != 64 || !mlp->ln_gamma || !mlp->ln_beta || mlp->in_dim != 64 * n_src) return false;
  for (int s = 0; s < n_src; ++s)
    if (srcs[s].dim != 64 || srcs[s].idx || !aligned16(srcs[s].ptr) || srcs[s].bstride % 4 != 0) return false;
  if (res) {
    if (res->idx) return false;
    bool match = false;
    for (int s = 0; s < n_src; ++s) match = match || (srcs[s].ptr == res->ptr && srcs[s].bstride == res->bstride);
    if (!match) return false;
  }
  return true;
}

int tc_rowmlp64(const NlamMlp* mlp, const NlamRowSrc* srcs, int n_src, const NlamRowSrc* res, float* out, int64_t n_rows,
                int B, cudaStream_t st) {
  NLAM_REQUIRE(aligned16(out), NLAM_E_INVALID, "tc_rowmlp64: output not 16-byte aligned");
  Row64Params p;
  memset(&p, 0, sizeof(p));
  CUtensorMap a[2], w1, w2, om;
  memset(a, 0, sizeof(a));
  p.n_src = n_src;
  p.out_src = 0;
  for (int s = 0; s < n_src; ++s) {
    const bool batched = srcs[s].bstride != 0 && B > 1;
    p.batched[s] = batched;
    int rc = make_map(&a[s], srcs[s].ptr, 64, (uint64_t)n_rows, batched ? (uint64_t)B : 1, 64,
                      batched ? (uint64_t)srcs[s].bstride : (uint64_t)n_rows * 64, 128, true);
    if (rc) return rc;
  }
  if (res) {
    p.has_res = 1;
    for (int s = n_src - 1; s >= 0; --s)
      if (srcs[s].ptr == res->ptr && srcs[s].bstride == res->bstride) p.out_src = s;
  }
  if (n_src == 1) a[1] = a[0];
  int rc = make_map(&w1, mlp->w[0], (uint64_t)mlp->in_dim, 64, 1, (uint64_t)mlp->in_dim, 0, 64, false);
  if (rc) return rc;
  rc = make_map(&w2, mlp->w[1], 64, 64, 1, 64, 0, 64, false);
  if (rc) return rc;
  rc = make_map(&om, out, 64, (uint64_t)n_rows, (uint64_t)B, 64, (uint64_t)n_rows * 64, 128, true);
  if (rc) return rc;
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
  tc_rowmlp64_kernel<<<grid, r4::THREADS, r4::SMEM, st>>>(a[0], a[1], w1, w2, om, p);
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
