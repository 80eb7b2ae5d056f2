// Node projection kernel of the split first Linear.
//
// The first Linear of the edge MLP acts on [e | x_src | x_dst] (reference gnn_layers.py:168-172).  It is
// linear, so   W1·[e; x_s; x_r] + b1 = W1e·e + (W1s·x)_src + (W1r·x + b1)_dst .
// The two node-side terms are computed once per NODE by `tc_rowlinear_kernel` (P_s = x_send·W1sᵀ,
// P_r = x_rec·W1rᵀ + b1; both problems of a call in ONE launch) instead of once per EDGE, and the edge kernels
// (tc5.cu: sender windows; tc3.cu: uniform in-degree) only run the K=64 GEMM e·W1eᵀ on the tensor cores and add the
// gathered projections in their first epilogue.  Summation order differs from the reference (three TF32 GEMMs summed
// in fp32) — inside the stated TF32 tolerance.
#include "tc_ptx.cuh"

namespace nlam {

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
  pdl_launch_dependents();
  pdl_wait();  // everything below may read what the previous kernel in the stream wrote
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
    NLAM_CUDA_OK(launch_pdl(tc_rowlinear_kernel, grid, rl::THREADS, rl::SMEM, st, mx[0], mw[0], mo[0], mx[1], mw[1], mo[1], p));
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

}  // namespace nlam
