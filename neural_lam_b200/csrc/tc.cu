// tcgen05 / TMEM / TMA kernels (sm_100a) for hidden width H = 64, TF32 inputs, fp32 accumulate.
//
// One persistent, warp-specialised kernel `tc_mlp_kernel` evaluates, per 128-row tile,
//     y = [LayerNorm]( W2 · SiLU(W1 · a + b1) + b2 )            (make_mlp, networks.py:27-40)
// with the two Linear layers as tcgen05.mma (kind::tf32, M=128, N=64|32, accumulators in
// TMEM) and everything around them fused:
//   edge mode (InteractionNet message + aggregate, gnn_layers.py:144-189):
//     a = [e | x[src] | x[dst]] : e tile by TMA, sender/receiver rows gathered with cp.async
//     straight into the swizzled UMMA operand layout; epilogue adds the edge residual
//     (e' = e + m, coalesced store) and segment-sums m over the CSR receiver segments of the
//     tile (tiles hold whole receivers, so no atomics and a deterministic order).
//   row mode (node update gnn_layers.py:148-151, embedders / grid MLPs base.py:286-322):
//     a = up to two 64-wide row blocks by TMA (e.g. [rec | aggr]) or a generic concatenation
//     of narrow inputs (grid features 17|17|18|4); optional residual, optional LayerNorm.
//
// Warp roles (832 threads, 1 CTA / SM, persistent over (batch, tile) work items):
//   warps 0-15  epilogue 2: D2 from TMEM (tcgen05.ld 32x32b), bias / LayerNorm / residual,
//               segmented reduction, TMA stores.  warp%4 selects the TMEM lane quarter (hardware
//               restriction), warp/4 the 16-column quarter.
//   warps 16-19 epilogue 1: D1 -> bias + SiLU -> hidden, TMEM to TMEM (thread = tile row); its own
//               group so that tile i+1's first epilogue overlaps tile i's second epilogue
//   warps 20-23 producers: index-driven gathers (cp.async 16 B, manual 128B swizzle)
//   warp 24     TMEM allocation + single-thread tcgen05.mma issue
//   warp 25     TMA loads (weights once, A tiles per work item)
//
// The hidden activations never touch shared memory: the first epilogue writes SiLU(D1+b1) back to
// TMEM (tcgen05.st) and the second GEMM reads its A operand from TMEM (tcgen05.mma "TS" form).
// Shared memory (dynamic, 1024-byte aligned): W1 6x8 KB | W2 2x8 KB | A 8x16 KB | HB 2x16 KB |
// barriers + LayerNorm exchange + local CSR offsets = 227 KB.  TMA-loaded A blocks are double
// buffered (edge mode: e tile stage 0 = blocks 0-1, stage 1 = blocks 6-7, gathered sender /
// receiver rows = blocks 2-5; row mode: stage s = blocks [s*nb1, (s+1)*nb1)), and so are the
// TMEM buffers (stage s: D1 at column s*192, hidden at +64, D2 at +128), so that the TMA loads, the
// gathers and the first GEMM of tile i+1 overlap the epilogue of tile i.
#include "tc_ptx.cuh"

namespace nlam {

constexpr int TC_THREADS = 832;
constexpr int EPI_THREADS = 512;
constexpr int EPI_WARPS = EPI_THREADS / 32;
constexpr int PROD_THREADS = 128;
constexpr int E1_THREADS = 128;
constexpr int W_E1 = EPI_WARPS;        // first warp of the epilogue-1 group (4 warps)
constexpr int W_PROD = EPI_WARPS + 4;   // producers (4 warps)
constexpr int W_MMA = EPI_WARPS + 8;
constexpr int W_TMA = EPI_WARPS + 9;
constexpr int BM = 128;
constexpr uint32_t A_BLOCK = 16384;  // 128 rows x 128 B
constexpr uint32_t W_BLOCK = 8192;   // 64 rows x 128 B
constexpr uint32_t OFF_W1 = 0;
constexpr uint32_t OFF_W2 = 6 * W_BLOCK;
constexpr uint32_t OFF_A = 8 * W_BLOCK;
constexpr uint32_t OFF_HB = OFF_A + 8 * A_BLOCK;
constexpr uint32_t OFF_MISC = OFF_HB + 2 * A_BLOCK;
constexpr uint32_t MISC_BYTES = 3072;  // total = 232448 B = the 227 KB opt-in maximum
constexpr uint32_t TC_SMEM = OFF_MISC + MISC_BYTES;

struct TcParams {
  int mode_edge;
  int nb1;        // K blocks (32 floats each) of the first Linear's input
  int a0_blocks;  // blocks loaded by TMA from source 0 / source 1
  int a1_blocks;
  int a0_batched, a1_batched;
  // gather sources (edge mode): blocks [2,4) <- gsrc[0][gidx[0][row]], [4,6) <- gsrc[1][..]
  const float* gsrc[2];
  long long gbs[2];
  const int32_t* gidx[2];
  int use_g4;      // gather with TMA tile::gather4 (one instruction per 4 rows x 32 columns)
  int g4_rows[2];  // rows of one batch in the gather4 tensor maps (0 for batch-broadcast sources)
  // generic element-wise sources (row mode): concatenated into blocks [0, nb1)
  int n_elem;
  const float* esrc[NLAM_MAX_SRC];
  long long ebs[NLAM_MAX_SRC];
  int edim[NLAM_MAX_SRC];
  int k_real;
  int elem_bulk;  // narrow sources are staged by 1-D bulk copies (alignment checked by the host)
  const float* b1;
  const float* b2;
  const float* gamma;
  const float* beta;
  float eps;
  int n2;    // padded N of the second Linear (64 or 32)
  int nout;  // real output width
  int res_block;  // first A block holding the residual rows, or -1
  float* out;     // (B, n_rows, nout) dense; NULL in edge mode without edge update
  float* aggr;    // edge mode: (B, n_rec, 64)
  int mean;
  long long n_rows;
  int B;
  int n_tiles;
  const int32_t* tile_rec;
  const int32_t* tile_e0;  // first CSR edge of each tile (n_tiles+1)
  const int4* tile_meta;   // {first edge, #edges, first receiver, #receivers} per tile
  const int32_t* rowptr;
  long long n_rec;
  // optional fused forecast-step epilogue on a narrow output (output_map): instead of the MLP output y the kernel
  // stores  m*boundary + (1-m)*(prev + (y*std + mean))  (graph/base.py:339-342, forecasters/autoregressive.py:128-131)
  const float* ep_prev;   // (B, n_rows, nout); NULL = no epilogue
  const float* ep_bnd;    // (B, n_rows, nout) or NULL
  const float* ep_mask;   // (n_rows) when ep_bnd
  const float* ep_std;    // (nout)
  const float* ep_mean;   // (nout)
  long long* dbg;  // optional per-phase clock64 timeline of block 0 (bring-up / profiling aid)
};

#define NLAM_DBG(slot, it)                                                      \
  do {                                                                          \
    if (p.dbg && blockIdx.x == 0 && (it) < 16) p.dbg[(it) * 16 + (slot)] = clock64(); \
  } while (0)

// ------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_mlp_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
              const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
              const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmG0,
              const __grid_constant__ CUtensorMap tmG1, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  // misc region: mbarriers | tmem ptr | LayerNorm exchange | local CSR offsets
  const uint32_t mb = sbase + OFF_MISC;
  const uint32_t bar_w = mb + 0;
  const uint32_t bar_a_gat_full = mb + 8;   // produced blocks (cp.async / element-wise loaders): 128 arrivals
  const uint32_t bar_a_free_g = mb + 16;    // ... and their release by the first GEMM's commit
  // gather4 mode, per gathered source (+8*s): 1 arrival + 32 KB of TMA transactions / release by commit
  const uint32_t bar_a_g4_full = mb + 24;
  const uint32_t bar_a_g4_free = mb + 40;
  // stage-indexed (+8*st)
  const uint32_t bar_a_tma_full = mb + 56;
  const uint32_t bar_epi_done = mb + 72;
  const uint32_t bar_d1_full = mb + 88;
  const uint32_t bar_d2_full = mb + 104;
  const uint32_t bar_hb_full = mb + 120;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 136);
  int* lp = reinterpret_cast<int*>(smem + OFF_MISC + 144);  // local CSR offsets, <= 129 entries
  // b1 | b2 (zero past nout) | gamma | beta: there is no L1 next to 227 KB of shared memory, so a per-tile __ldg
  // of these constants would put an L2 round trip into every epilogue
  float* sprm = reinterpret_cast<float*>(smem + OFF_MISC + 1024);
  const uint32_t bar_st_full = mb + 896;  // [2] narrow-source staging filled by bulk copies
  const uint32_t bar_st_free = mb + 912;  // [2] ... and repacked (128 arrivals)

  if ((sbase & 1023u) != 0) {
    if (tid == 0) printf("nlam tc kernel: dynamic shared memory not 1024-byte aligned\n");
    __trap();
  }

  const bool has_tma_a = (p.a0_blocks + p.a1_blocks) > 0;
  const bool has_prod = p.mode_edge || p.n_elem > 0;

  if (warp == W_MMA) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_a_gat_full, PROD_THREADS);
      mbar_init(bar_a_free_g, 1);
      mbar_init(bar_a_g4_full, 1);
      mbar_init(bar_a_g4_full + 8, 1);
      mbar_init(bar_a_g4_free, 1);
      mbar_init(bar_a_g4_free + 8, 1);
      for (int st = 0; st < 2; ++st) {
        mbar_init(bar_st_full + 8 * st, 1);
        mbar_init(bar_st_free + 8 * st, PROD_THREADS);
      }
      for (int st = 0; st < 2; ++st) {
        mbar_init(bar_a_tma_full + 8 * st, 1);
        mbar_init(bar_epi_done + 8 * st, 1);
        mbar_init(bar_d1_full + 8 * st, 1);
        mbar_init(bar_d2_full + 8 * st, 1);
        mbar_init(bar_hb_full + 8 * st, E1_THREADS);
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
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
    if (p.a0_blocks) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    if (p.a1_blocks) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    if (p.out && p.nout == 64) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
    if (p.use_g4) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG0) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG1) : "memory");
    }
  }
  if (tid < 64) {
    sprm[tid] = p.b1[tid];
    sprm[64 + tid] = (tid < p.nout) ? p.b2[tid] : 0.f;
    sprm[128 + tid] = p.gamma ? p.gamma[tid] : (p.ep_prev && tid < p.nout ? p.ep_std[tid] : 1.f);
    sprm[192 + tid] = p.gamma ? p.beta[tid] : (p.ep_prev && tid < p.nout ? p.ep_mean[tid] : 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // first A block of TMA stage st (edge mode keeps the gathered blocks 2-5 single buffered)
  const int stage_blk1 = p.mode_edge ? 6 : p.nb1;

  const int n_work = p.n_tiles * p.B;  // host guarantees < 2^31

  if (warp == W_TMA) {
    // =============================== TMA loader ===============================
    if (lane == 0) {
      const uint32_t w2_block_bytes = (uint32_t)p.n2 * 128u;
      mbar_expect_tx(bar_w, (uint32_t)p.nb1 * W_BLOCK + 2u * w2_block_bytes);
      for (int j = 0; j < p.nb1; ++j) tma_load_2d(sbase + OFF_W1 + j * W_BLOCK, &tmW1, bar_w, 32 * j, 0);
      for (int j = 0; j < 2; ++j) tma_load_2d(sbase + OFF_W2 + j * W_BLOCK, &tmW2, bar_w, 32 * j, 0);
      if (p.elem_bulk) {
        // narrow sources (e.g. prev | prev_prev | forcing | static): the 128-row slab of every source is contiguous
        // in global memory -> one 1-D bulk copy each into a flat staging area (A blocks 2-4 / 5-7, unused in this
        // mode), a tile ahead of the producers that repack it into the K-major operand tile
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
          const int b = w / p.n_tiles;
          const int t = w - b * p.n_tiles;
          const int st = it & 1;
          const int nrows = (int)min((long long)BM, p.n_rows - (long long)t * BM);
          mbar_wait(bar_st_free + 8 * st, (uint32_t)(((it >> 1) & 1) ^ 1));
          NLAM_DBG(0, it);
          mbar_expect_tx(bar_st_full + 8 * st, (uint32_t)(nrows * p.k_real * 4));
          uint32_t dst = sbase + OFF_A + (2 + 3 * st) * A_BLOCK;
          for (int sidx = 0; sidx < p.n_elem; ++sidx) {
            const int d = p.edim[sidx];
            const float* src = p.esrc[sidx] + (long long)b * p.ebs[sidx] + (long long)t * BM * d;
            bulk_load_1d(dst, src, (uint32_t)(nrows * d * 4), bar_st_full + 8 * st);
            dst += (uint32_t)(BM * d * 4);
          }
        }
      }
      if (has_tma_a) {
        const uint64_t pol_stream = policy_evict_first();
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
          const int b = w / p.n_tiles;
          const int t = w - b * p.n_tiles;
          const int st = it & 1;
          const int row0 = p.mode_edge ? p.tile_e0[t] : t * BM;
          const uint32_t abase = sbase + OFF_A + (st ? stage_blk1 : 0) * A_BLOCK;
          const uint32_t full = bar_a_tma_full + 8 * st;
          mbar_wait(bar_epi_done + 8 * st, (uint32_t)(((it >> 1) & 1) ^ 1));  // tile it-2 released the stage
          NLAM_DBG(0, it);
          mbar_expect_tx(full, (uint32_t)(p.a0_blocks + p.a1_blocks) * A_BLOCK);
          for (int j = 0; j < p.a0_blocks; ++j)
            tma_load_3d(abase + j * A_BLOCK, &tmA0, full, 32 * j, row0, p.a0_batched ? b : 0, pol_stream);
          for (int j = 0; j < p.a1_blocks; ++j)
            tma_load_3d(abase + (p.a0_blocks + j) * A_BLOCK, &tmA1, full, 32 * j, row0, p.a1_batched ? b : 0, pol_stream);
          // pull the tile two iterations ahead into L2 so that its TMA load is an L2 hit
          const int w2 = w + 2 * (int)gridDim.x;
          if (w2 < n_work) {
            const int b2 = w2 / p.n_tiles, t2 = w2 - b2 * p.n_tiles;
            const int r2 = p.mode_edge ? p.tile_e0[t2] : t2 * BM;
            for (int j = 0; j < p.a0_blocks; ++j) tma_prefetch_3d(&tmA0, 32 * j, r2, p.a0_batched ? b2 : 0);
            for (int j = 0; j < p.a1_blocks; ++j) tma_prefetch_3d(&tmA1, 32 * j, r2, p.a1_batched ? b2 : 0);
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issuer ===============================
    // Software pipelined: GEMM1 of tile i+1 is issued as soon as its operands have landed,
    // GEMM2 of tile i as soon as the epilogue has produced the hidden activations.
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc_tf32(BM, 64);
      const uint32_t idesc2 = umma_idesc_tf32(BM, p.n2);
      int n_my = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) ++n_my;
      mbar_wait(bar_w, 0);
      const uint64_t desc_a0 = umma_desc(sbase + OFF_A);
      const uint64_t desc_w1 = umma_desc(sbase + OFF_W1);
      const uint64_t desc_w2 = umma_desc(sbase + OFF_W2);
      int g1 = 0, g2 = 0, g1_phase = 0;
      uint32_t idle = 0;
      while (g2 < n_my) {
        bool progress = false;
        if (g1 < n_my && g1 <= g2 + 1) {
          // (g1 <= g2+1 also guarantees that epilogue 1 of tile g1-2 has drained this stage's D1.)
          const int it = g1, st = it & 1;
          const uint32_t sph = (uint32_t)((it >> 1) & 1);
          const uint32_t d1 = tmem_base + st * 192;
          if (p.mode_edge && p.use_g4) {
            // operands arrive independently: sender rows, receiver rows, e tile.  Issue each K-slice as
            // soon as it has landed and release its blocks right away, so that the next tile's gathers
            // overlap the rest of this GEMM.
            if (g1_phase < 2) {
              if (mbar_test(bar_a_g4_full + 8 * g1_phase, (uint32_t)(it & 1))) {
                tc_fence_after();
                if (g1_phase == 0) NLAM_DBG(3, it);
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                  const int j = 2 + 2 * g1_phase + jj;
                  const uint64_t ad0 = desc_a0 + (uint64_t)((j * A_BLOCK) >> 4);
                  const uint64_t bd0 = desc_w1 + (uint64_t)((j * W_BLOCK) >> 4);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_tf32(d1, ad0 + 2 * k, bd0 + 2 * k, idesc1, (uint32_t)((g1_phase | jj | k) != 0));
                }
                umma_commit(bar_a_g4_free + 8 * g1_phase);
                ++g1_phase;
                progress = true;
              }
            } else if (mbar_test(bar_a_tma_full + 8 * st, sph)) {
              tc_fence_after();
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const int blk = st ? 6 + j : j;
                const uint64_t ad0 = desc_a0 + (uint64_t)((blk * A_BLOCK) >> 4);
                const uint64_t bd0 = desc_w1 + (uint64_t)((j * W_BLOCK) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(d1, ad0 + 2 * k, bd0 + 2 * k, idesc1, 1u);
              }
              umma_commit(bar_d1_full + 8 * st);
              g1_phase = 0;
              ++g1;
              progress = true;
            }
          } else {
            bool ready = true;
            if (has_tma_a) ready = mbar_test(bar_a_tma_full + 8 * st, sph);
            else ready = mbar_test(bar_epi_done + 8 * st, sph ^ 1);  // accumulators of tile it-2 drained
            if (ready && has_prod) ready = mbar_test(bar_a_gat_full, (uint32_t)(it & 1));
            if (ready) {
              tc_fence_after();
              NLAM_DBG(3, it);
              for (int j = 0; j < p.nb1; ++j) {
                // A block j: TMA blocks come from the stage, produced blocks (edge gathers) are fixed
                int blk;
                if (p.mode_edge) blk = (j < 2) ? (st ? 6 + j : j) : j;
                else blk = has_tma_a ? (st ? stage_blk1 : 0) + j : j;
                // descriptor = base + (byte offset >> 4) in the 14-bit start-address field (no carry:
                // all operands live below 256 KB)
                const uint64_t ad0 = desc_a0 + (uint64_t)((blk * A_BLOCK) >> 4);
                const uint64_t bd0 = desc_w1 + (uint64_t)((j * W_BLOCK) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(d1, ad0 + 2 * k, bd0 + 2 * k, idesc1, (uint32_t)((j | k) != 0));
              }
              umma_commit(bar_d1_full + 8 * st);
              if (has_prod) umma_commit(bar_a_free_g);
              ++g1;
              progress = true;
            }
          }
        }
        if (g2 < g1) {
          const int it = g2, st = it & 1;
          if (mbar_test(bar_hb_full + 8 * st, (uint32_t)((it >> 1) & 1))) {
            tc_fence_after();
            NLAM_DBG(4, it);
            const uint32_t ht = tmem_base + st * 192 + 64;  // hidden activations (A operand in TMEM)
            const uint32_t d2 = tmem_base + st * 192 + 128;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_ts(d2, ht + (uint32_t)(j * 32 + k * 8), desc_w2 + (uint64_t)((j * W_BLOCK) >> 4) + 2 * k, idesc2,
                             (uint32_t)((j | k) != 0));
            }
            umma_commit(bar_d2_full + 8 * st);
            ++g2;
            progress = true;
          }
        }
        if (progress) idle = 0;
        else if (++idle > (1u << 26)) {
          printf("nlam tc kernel: MMA issuer timeout (block %d g1 %d g2 %d)\n", blockIdx.x, g1, g2);
          __trap();
        }
      }
    }
  } else if (warp >= W_PROD) {
    // =============================== producers ===============================
    if (has_prod) {
      const int pt = tid - W_PROD * 32;  // 0..127
      if (!p.mode_edge) {
        for (int cc = p.k_real; cc < p.nb1 * 32; ++cc)
          *reinterpret_cast<float*>(smem + OFF_A + (cc >> 5) * A_BLOCK + swz(pt, (cc & 31) >> 2) + (cc & 3) * 4) = 0.f;
      }
      const uint64_t pol_keep = policy_evict_last();
      int pf_ne = 0;
      int pf_idx[2] = {0, 0};
      if (p.mode_edge && (int)blockIdx.x < n_work) {
        const int t0 = (int)blockIdx.x % p.n_tiles;
        const int e00 = p.tile_e0[t0];
        pf_ne = (int)min((long long)BM, p.n_rows - e00);
        const int my_row = (pt >> 5) * 32 + lane;
        pf_idx[0] = (my_row < pf_ne) ? __ldg(p.gidx[0] + e00 + my_row) : 0;
        pf_idx[1] = (my_row < pf_ne) ? __ldg(p.gidx[1] + e00 + my_row) : 0;
      }
      int pf4[4] = {0, 0, 0, 0};
      if (p.mode_edge && p.use_g4 && (int)blockIdx.x < n_work) {
        const int e00 = p.tile_e0[(int)blockIdx.x % p.n_tiles];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          pf4[u] = (4 * (pt & 31) + u < pf_ne) ? __ldg(p.gidx[pt >> 6] + e00 + 4 * (pt & 31) + u) : 0;
      }
      int it = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int b = w / p.n_tiles;
        const int t = w - b * p.n_tiles;
        if (p.use_g4) {
          // producer threads 0-63 gather the sender rows, 64-127 the receiver rows; each half waits for
          // ITS blocks to be released (the first GEMM commits after each source's MMAs) and arms its own
          // transaction barrier before any of its threads can issue
          const int half = pt >> 6;
          if ((pt & 63) < 32) {  // leader warp of the half
            mbar_wait(bar_a_g4_free + 8 * half, (uint32_t)((it & 1) ^ 1));
            if (lane == 0) mbar_expect_tx(bar_a_g4_full + 8 * half, 2u * A_BLOCK);
          }
          named_bar_sync(8 + half, 64);
        } else {
          group_wait(warp == W_PROD, bar_a_free_g, (uint32_t)((it & 1) ^ 1), 8, PROD_THREADS);
        }
        if (pt == 0) NLAM_DBG(1, it);
        if (p.mode_edge && p.use_g4) {
          // TMA tile::gather4: producer thread pt issues ONE instruction = 4 gathered rows x 32 columns
          // (512 B, hardware 128B swizzle) of source (pt>>6), column block (pt>>5)&1, row group pt&31.
          // The four row indices were prefetched during the previous iteration.
          const int sidx = pt >> 6, jb = (pt >> 5) & 1, grp = pt & 31;
          const int boff = p.g4_rows[sidx] * b;
          int r[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) r[u] = (4 * grp + u < pf_ne) ? pf4[u] + boff : 0;  // past the end: any valid row (never stored)
          tma_gather4(sbase + OFF_A + (2 + 2 * sidx + jb) * A_BLOCK + grp * 512, sidx ? &tmG1 : &tmG0,
                      bar_a_g4_full + 8 * sidx, 32 * jb, r[0], r[1], r[2], r[3], pol_keep);
          if (pt == 0) NLAM_DBG(14, it);
          {
            const int wn = w + (int)gridDim.x;
            if (wn < n_work) {
              const int e0n = p.tile_e0[wn % p.n_tiles];
              pf_ne = (int)min((long long)BM, p.n_rows - e0n);
#pragma unroll
              for (int u = 0; u < 4; ++u) pf4[u] = (4 * grp + u < pf_ne) ? __ldg(p.gidx[sidx] + e0n + 4 * grp + u) : 0;
            }
          }
          if (pt == 0) NLAM_DBG(2, it);
          continue;  // completion is signalled by the TMA transactions themselves
        } else if (p.mode_edge) {
          // each warp owns 32 tile rows; a half-warp copies one 256-byte row per instruction.  The
          // row indices of this tile were loaded while waiting (prefetched in the previous iteration).
          const int ne = pf_ne;
          const int pw = pt >> 5;
#pragma unroll
          for (int sidx = 0; sidx < 2; ++sidx) {
            const int my_idx = pf_idx[sidx];
            const float* base = p.gsrc[sidx] + (long long)b * p.gbs[sidx];
            const uint32_t blk0 = sbase + OFF_A + (2 + 2 * sidx) * A_BLOCK;
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
              const int rl = 2 * i + (lane >> 4);  // row within the warp's 32
              const int row = pw * 32 + rl;
              const int src_row = __shfl_sync(0xffffffffu, my_idx, rl);
              const int ch = lane & 15;  // 16-byte chunk of the 256-byte row
              const float* g = base + (long long)src_row * 64 + ch * 4;
              const uint32_t dst = blk0 + (ch >> 3) * A_BLOCK + swz(row, ch & 7);
              if (row < ne) cp_async_16(dst, g, pol_keep);
            }
          }
          // prefetch the indices of the next tile while the copies are in flight
          {
            const int wn = w + (int)gridDim.x;
            if (wn < n_work) {
              const int tn = wn % p.n_tiles;
              const int e0n = p.tile_e0[tn];
              pf_ne = (int)min((long long)BM, p.n_rows - e0n);
              const int my_row = pw * 32 + lane;
              pf_idx[0] = (my_row < pf_ne) ? __ldg(p.gidx[0] + e0n + my_row) : 0;
              pf_idx[1] = (my_row < pf_ne) ? __ldg(p.gidx[1] + e0n + my_row) : 0;
            }
          }
          cp_async_wait_all();
        } else {
          // generic concatenation of narrow inputs (e.g. prev | prev_prev | forcing | static grid
          // features, graph/base.py:275-283): producer thread pt owns tile row pt and walks its row
          // of every source; the zero padding up to nb1*32 columns was written once before the loop
          const long long gr = (long long)t * BM + pt;
          if (p.elem_bulk) {
            const int st = it & 1;
            group_wait(warp == W_PROD, bar_st_full + 8 * st, (uint32_t)((it >> 1) & 1), 8, PROD_THREADS);
            const float* stg = reinterpret_cast<const float*>(smem + OFF_A + (2 + 3 * st) * A_BLOCK);
            int col = 0;
#pragma unroll
            for (int sidx = 0; sidx < NLAM_MAX_SRC; ++sidx) {
              if (sidx < p.n_elem) {
                const int d = p.edim[sidx];
                const float* srow = stg + pt * d;  // lane stride d floats: conflict-free for odd d
                for (int c = 0; c < d; ++c) {
                  const int cc = col + c;
                  *reinterpret_cast<float*>(smem + OFF_A + (cc >> 5) * A_BLOCK + swz(pt, (cc & 31) >> 2) + (cc & 3) * 4) = srow[c];
                }
                col += d;
                stg += BM * d;
              }
            }
            mbar_arrive(bar_st_free + 8 * st);
          } else if (gr < p.n_rows) {
            int col = 0;
#pragma unroll
            for (int sidx = 0; sidx < NLAM_MAX_SRC; ++sidx) {
              if (sidx < p.n_elem) {
                const int d = p.edim[sidx];
                const float* src = p.esrc[sidx] + (long long)b * p.ebs[sidx] + gr * d;
                int c = 0;
                for (; c + 4 <= d; c += 4) {
                  const float v0 = __ldg(src + c), v1 = __ldg(src + c + 1), v2 = __ldg(src + c + 2), v3 = __ldg(src + c + 3);
                  const float vv[4] = {v0, v1, v2, v3};
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const int cc = col + c + u;
                    *reinterpret_cast<float*>(smem + OFF_A + (cc >> 5) * A_BLOCK + swz(pt, (cc & 31) >> 2) + (cc & 3) * 4) = vv[u];
                  }
                }
                for (; c < d; ++c) {
                  const int cc = col + c;
                  *reinterpret_cast<float*>(smem + OFF_A + (cc >> 5) * A_BLOCK + swz(pt, (cc & 31) >> 2) + (cc & 3) * 4) = __ldg(src + c);
                }
                col += d;
              }
            }
          }
        }
        fence_proxy_async();
        if (pt == 0) NLAM_DBG(2, it);
        mbar_arrive(bar_a_gat_full);
      }
    }
  } else if (warp >= W_E1) {
    // =============================== epilogue 1 (warps 16-19) ===============================
    // hidden = SiLU(D1 + b1), TMEM -> registers -> TMEM (A operand of the second GEMM).  Runs on
    // its own warp group so that it overlaps the second epilogue of the previous tile.
    const int q = warp & 3;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t sph = (uint32_t)((it >> 1) & 1);
      const uint32_t tmem_d1 = tmem_base + st * 192;
      const uint32_t tmem_ht = tmem_d1 + 64;
      if (tid == W_E1 * 32) NLAM_DBG(5, it);
      if (warp == W_E1) {
        mbar_wait(bar_d2_full + 8 * st, sph ^ 1);  // GEMM2 of tile it-2 has consumed this hidden buffer
        mbar_wait(bar_d1_full + 8 * st, sph);
      }
      named_bar_sync(7, E1_THREADS);
      tc_fence_after();
      if (tid == W_E1 * 32) NLAM_DBG(6, it);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        float v[32];
        tmem_ld32(tmem_d1 + t_lane + cc * 32, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 bb = lds128(sprm + cc * 32 + 4 * k);
          v[4 * k + 0] = silu_fast(v[4 * k + 0] + bb.x);
          v[4 * k + 1] = silu_fast(v[4 * k + 1] + bb.y);
          v[4 * k + 2] = silu_fast(v[4 * k + 2] + bb.z);
          v[4 * k + 3] = silu_fast(v[4 * k + 3] + bb.w);
        }
        tmem_st32(tmem_ht + t_lane + cc * 32, v);
      }
      tc_fence_before();
      mbar_arrive(bar_hb_full + 8 * st);
      if (tid == W_E1 * 32) NLAM_DBG(7, it);
    }
  } else {
    // =============================== epilogue 2 (warps 0-15) ===============================
    // thread = (tile row, 16-column quarter): warp%4 selects the TMEM lane quarter (hardware
    // restriction of tcgen05.ld), warp/4 the column quarter.
    const int q = warp & 3;
    const int cq = warp >> 2;
    const int row = q * 32 + lane;
    const int c0 = cq * 16;
    const int blk = cq >> 1;             // 32-column block holding this thread's columns
    const int ch0 = (cq & 1) * 4;        // first 16-byte chunk inside that block
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    const uint32_t rsw = (uint32_t)(row * 128);
    const int rx = row & 7;
    const int lnbar = 2 + q;  // LayerNorm exchange only couples the 4 warps of one lane quarter

    // per-tile metadata {first row, #rows, first receiver, #receivers}, prefetched one tile ahead
    auto load_meta = [&](int w) -> int4 {
      const int t = w % p.n_tiles;
      if (p.mode_edge) return __ldg(p.tile_meta + t);
      const long long r0 = (long long)t * BM;
      return make_int4((int)r0, (int)min((long long)BM, p.n_rows - r0), 0, 0);
    };
    int4 meta = make_int4(0, 0, 0, 0);
    int4 meta_n = meta;
    int lp_val = 0;
    if ((int)blockIdx.x < n_work) {
      meta = load_meta(blockIdx.x);
      meta_n = meta;
      if ((int)(blockIdx.x + gridDim.x) < n_work) meta_n = load_meta(blockIdx.x + gridDim.x);
      if (p.mode_edge && tid <= meta.w) lp_val = __ldg(p.rowptr + meta.z + tid) - meta.x;
    }
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t sph = (uint32_t)((it >> 1) & 1);
      const uint32_t tmem_d2 = tmem_base + st * 192 + 128;
      const int b = w / p.n_tiles;
      const int row0 = meta.x, nrows = meta.y, r0 = meta.z, nrec = meta.w;
      const int cur_lp = lp_val;
      // metadata is prefetched two tiles ahead so that the dependent CSR-offset load of the next
      // tile never waits on it
      const int wn = w + (int)gridDim.x;
      const int wnn = wn + (int)gridDim.x;
      int4 meta_nn = meta_n;
      if (wnn < n_work) meta_nn = load_meta(wnn);

      float v[16];
      if (p.mode_edge) {
        if (tid <= nrec) lp[tid] = cur_lp;  // the previous tile's reduction has passed its final barrier
        if (wn < n_work && tid <= meta_n.w) lp_val = __ldg(p.rowptr + meta_n.z + tid) - meta_n.x;
      }

      // ---- E2: y = D2 + b2, LayerNorm, residual ----
      group_wait(warp == 0, bar_d2_full + 8 * st, sph, 6, EPI_THREADS);
      tc_fence_after();
      if (tid == 0) NLAM_DBG(8, it);
      const bool active = c0 < p.n2;
      const bool wide = (p.nout == 64);
      if (active) tmem_ld16(tmem_d2 + t_lane + c0, v);
      if (wide) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 bb = lds128(sprm + 64 + c0 + 4 * k);
          v[4 * k + 0] += bb.x;
          v[4 * k + 1] += bb.y;
          v[4 * k + 2] += bb.z;
          v[4 * k + 3] += bb.w;
        }
      } else if (active) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < p.nout) v[i] += sprm[64 + c0 + i];
      }
      if (p.gamma) {
        // LayerNorm over 64 columns split across the four column quarters: each thread parks its
        // partial (sum, sum of squares) in two spare TMEM columns of its lane, one barrier scoped to
        // the 4 warps of the lane quarter, then every thread reads the 4 pairs of its row back.
        float s = 0.f, sq = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          s += v[i];
          sq = fmaf(v[i], v[i], sq);
        }
        const uint32_t lnx = tmem_base + 384 + t_lane;
        tmem_st2(lnx + 2 * cq, s, sq);
        tc_fence_before();
        named_bar_sync(lnbar, 128);
        tc_fence_after();
        float st8[8];
        tmem_ld8(lnx, st8);
        const float mu = (st8[0] + st8[2] + st8[4] + st8[6]) * (1.0f / 64.0f);
        const float ex2 = (st8[1] + st8[3] + st8[5] + st8[7]) * (1.0f / 64.0f);
        const float rstd = rsqrtf(fmaxf(ex2 - mu * mu, 0.f) + p.eps);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 g4 = lds128(sprm + 128 + c0 + 4 * k);
          const float4 b4 = lds128(sprm + 192 + c0 + 4 * k);
          v[4 * k + 0] = (v[4 * k + 0] - mu) * rstd * g4.x + b4.x;
          v[4 * k + 1] = (v[4 * k + 1] - mu) * rstd * g4.y + b4.y;
          v[4 * k + 2] = (v[4 * k + 2] - mu) * rstd * g4.z + b4.z;
          v[4 * k + 3] = (v[4 * k + 3] - mu) * rstd * g4.w + b4.w;
        }
      }
      if (tid == 0) NLAM_DBG(9, it);

      // staging: messages (edge mode) -> HB; output rows -> residual blocks in place, or HB
      uint32_t stage = 0;  // shared address of the staged 64-wide output tile (swizzled block pair)
      if (wide) {
        if (p.mode_edge) {
          uint8_t* hb = smem + OFF_HB + blk * A_BLOCK + rsw;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<float4*>(hb + (((ch0 + k) ^ rx) << 4)) =
                make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
        if (p.res_block >= 0) {
          const uint32_t soff = OFF_A + ((st ? stage_blk1 : 0) + p.res_block) * A_BLOCK;
          stage = sbase + soff;
          if (p.out) {
            if (has_tma_a) mbar_wait(bar_a_tma_full + 8 * st, sph);  // TMA-written rows visible to this thread
            uint8_t* rb = smem + soff + blk * A_BLOCK + rsw;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float4* ptr = reinterpret_cast<float4*>(rb + (((ch0 + k) ^ rx) << 4));
              float4 r = *ptr;
              r.x += v[4 * k];
              r.y += v[4 * k + 1];
              r.z += v[4 * k + 2];
              r.w += v[4 * k + 3];
              *ptr = r;
            }
          }
        } else if (!p.mode_edge) {
          stage = sbase + OFF_HB;
          uint8_t* hb = smem + OFF_HB + blk * A_BLOCK + rsw;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<float4*>(hb + (((ch0 + k) ^ rx) << 4)) =
                make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
      } else if (active) {
        // narrow output (e.g. output_map 64 -> 17): plain [128][nout] floats in HB, odd pitch
        float* flat = reinterpret_cast<float*>(smem + OFF_HB);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < p.nout) flat[row * p.nout + c0 + i] = v[i];
      }
      fence_proxy_async();  // staged tile -> async proxy (TMA store)
      named_bar_sync(1, EPI_THREADS);
      if (tid == 0) NLAM_DBG(10, it);

      // ---- output: one TMA tensor store per 32-column block (full 128-row box, clipped at the
      // tensor end; in edge mode the rows past the tile's own edges belong to the following
      // tiles and carry the identical values those tiles store) ----
      if (p.out) {
        if (wide) {
          if (tid == 0 && stage) {
            tma_store_3d(&tmOut, stage, 0, row0, b);
            tma_store_3d(&tmOut, stage + A_BLOCK, 32, row0, b);
            bulk_commit();
          }
        } else {
          const float* flat = reinterpret_cast<const float*>(smem + OFF_HB);
          float* og = p.out + ((long long)b * p.n_rows + row0) * p.nout;
          const int n = nrows * p.nout;
          if (!p.ep_prev) {
            for (int i = tid; i < n; i += EPI_THREADS) og[i] = flat[i];
          } else {
            const long long g0 = ((long long)b * p.n_rows + row0) * p.nout;
            for (int i = tid; i < n; i += EPI_THREADS) {
              const int r = i / p.nout, c = i - r * p.nout;
              float v1 = __ldg(p.ep_prev + g0 + i) + (flat[i] * sprm[128 + c] + sprm[192 + c]);
              if (p.ep_bnd) {
                const float m = __ldg(p.ep_mask + row0 + r);
                v1 = m * __ldg(p.ep_bnd + g0 + i) + (1.0f - m) * v1;
              }
              og[i] = v1;
            }
          }
        }
      }
      if (tid == 0) NLAM_DBG(11, it);
      // ---- segmented sum of the messages over the tile's receivers (CSR order) ----
      if (p.mode_edge) {
        // thread = (float4 column group cg, receiver group g): 16 x 32
        const int cg = tid & 15, g = tid >> 4;
        const uint8_t* mbase = smem + OFF_HB + (cg >> 3) * A_BLOCK;
        const int chq = cg & 7;
        for (int j = g; j < nrec; j += EPI_THREADS / 16) {
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
      if (tid == 0) {
        NLAM_DBG(12, it);
        if (p.out && wide && stage) bulk_wait_read0();  // staged tile fully read by the TMA engine
      }
      tc_fence_before();
      named_bar_sync(1, EPI_THREADS);
      if (tid == 0) {
        NLAM_DBG(13, it);
        mbar_arrive(bar_epi_done + 8 * st);
      }
      meta = meta_n;
      meta_n = meta_nn;
    }
    if (tid == 0) bulk_wait0();  // all output stores complete before the CTA retires
  }

  // teardown
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == W_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


static int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w1, const CUtensorMap& w2,
                  const CUtensorMap& om, const TcParams& p, cudaStream_t st, double bytes, const CUtensorMap* g0 = nullptr,
                  const CUtensorMap* g1 = nullptr) {
  static unsigned attr_mask = 0;  // per device
  int dev = 0;
  NLAM_CUDA_OK(cudaGetDevice(&dev));
  if (!(attr_mask & (1u << (dev & 31)))) {
    NLAM_CUDA_OK(cudaFuncSetAttribute(tc_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    attr_mask |= 1u << (dev & 31);
  }
  long long n_work = (long long)p.n_tiles * p.B;
  NLAM_REQUIRE(n_work < (1LL << 31) - 4096, NLAM_E_UNSUPPORTED, "tc kernel: too many work items (%lld)", n_work);
  int grid = (int)std::min<long long>(n_work, num_sms());
  TcParams pp = p;
  static long long* dbg_buf = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) dbg_on = getenv("NLAM_TC_TIMELINE") ? 1 : 0;
  if (dbg_on) {
    if (!dbg_buf) NLAM_CUDA_OK(cudaMalloc(&dbg_buf, 256 * sizeof(long long)));
    NLAM_CUDA_OK(cudaMemsetAsync(dbg_buf, 0, 256 * sizeof(long long), st));
    pp.dbg = dbg_buf;
  }
  {
    ProfScope ps(p.mode_edge ? "tc_mlp_kernel(edge)" : "tc_mlp_kernel(row)", st, bytes);
    tc_mlp_kernel<<<grid, TC_THREADS, TC_SMEM, st>>>(a0, a1, w1, w2, om, g0 ? *g0 : om, g1 ? *g1 : om, pp);
  }
  if (dbg_on) {
    long long h[256];
    NLAM_CUDA_OK(cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
    NLAM_CUDA_OK(cudaStreamSynchronize(st));
    long long t0 = h[0] ? h[0] : h[1];
    fprintf(stderr, "[nlam tc timeline] mode_edge=%d nb1=%d grid=%d work=%lld (cycles rel. to first event)\n", p.mode_edge, p.nb1, grid, n_work);
    fprintf(stderr, " it    tma  g_start g_done   g1_iss  g2_iss | e_wait  d1_rdy  e1_done d2_rdy  ln_done staged  copied  reduced end | g_issued g_pref\n");
    for (int it = 0; it < 8; ++it) {
      fprintf(stderr, "%3d ", it);
      for (int k = 0; k < 16; ++k) fprintf(stderr, "%7lld ", h[it * 16 + k] ? h[it * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  count_launch();
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

static int weight_maps(const NlamMlp* m, int nb1, int n2, int nout, CUtensorMap* w1, CUtensorMap* w2) {
  (void)nb1;
  int rc = make_map(w1, m->w[0], (uint64_t)m->in_dim, 64, 1, (uint64_t)m->in_dim, 0, 64, false);
  if (rc) return rc;
  return make_map(w2, m->w[1], 64, (uint64_t)nout, 1, 64, 0, (uint32_t)n2, false);
}

bool tc_rowmlp_supported(const NlamMlp* mlp, const NlamRowSrc* srcs, int n_src, const NlamRowSrc* res,
                         const NlamRowSrc* res2, int64_t n_rows) {
  int nout = 0;
  if (res2 || n_rows < 1 || n_rows >= (1LL << 31) - 256) return false;
  if (!mlp_shape_ok(mlp, &nout)) return false;
  bool all_wide = n_src <= 2;
  for (int s = 0; s < n_src; ++s)
    all_wide = all_wide && srcs[s].dim == 64 && !srcs[s].idx && aligned16(srcs[s].ptr) && (srcs[s].bstride % 4 == 0);
  if (all_wide) {
    if (res) {
      if (nout != 64 || res->idx) return false;
      bool match = false;
      for (int s = 0; s < n_src; ++s) match = match || (srcs[s].ptr == res->ptr && srcs[s].bstride == res->bstride);
      if (!match) return false;
    }
    return true;
  }
  // generic narrow concatenation (grid embedder): total width <= 64, no gathers, no residual
  if (res || mlp->in_dim > 64) return false;
  for (int s = 0; s < n_src; ++s)
    if (srcs[s].idx) return false;
  return true;
}

int tc_rowmlp(const NlamMlp* mlp, const NlamRowSrc* srcs, int n_src, const NlamRowSrc* res, float* out, int64_t n_rows,
              int B, cudaStream_t st, const StepEpilogue* ep) {
  int nout = 0;
  NLAM_REQUIRE(mlp_shape_ok(mlp, &nout), NLAM_E_UNSUPPORTED, "tc_rowmlp: unsupported MLP shape");
  NLAM_REQUIRE(aligned16(out), NLAM_E_INVALID, "tc_rowmlp: output not 16-byte aligned");
  // dense 64-wide inputs, LayerNorm output: the streaming kernel of tc4.cu
  if (!ep && tc_rowmlp64_supported(mlp, srcs, n_src, res, n_rows)) return tc_rowmlp64(mlp, srcs, n_src, res, out, n_rows, B, st);
  if (!res && tc_rowmlp_narrow_out_supported(mlp, srcs, n_src, n_rows, out, ep)) return tc_rowmlp64(mlp, srcs, n_src, nullptr, out, n_rows, B, st, ep);
  TcParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap a0, a1, w1, w2;
  memset(&a0, 0, sizeof(a0));
  memset(&a1, 0, sizeof(a1));
  bool all_wide = n_src <= 2;
  for (int s = 0; s < n_src; ++s)
    all_wide = all_wide && srcs[s].dim == 64 && !srcs[s].idx && aligned16(srcs[s].ptr) && (srcs[s].bstride % 4 == 0);
  p.res_block = -1;
  if (all_wide) {
    p.nb1 = 2 * n_src;
    for (int s = 0; s < n_src; ++s) {
      const bool batched = srcs[s].bstride != 0 && B > 1;
      int rc = make_map(s == 0 ? &a0 : &a1, srcs[s].ptr, 64, (uint64_t)n_rows, batched ? (uint64_t)B : 1, 64,
                        batched ? (uint64_t)srcs[s].bstride : (uint64_t)n_rows * 64, BM, true);
      if (rc) return rc;
      if (s == 0) { p.a0_blocks = 2; p.a0_batched = batched; } else { p.a1_blocks = 2; p.a1_batched = batched; }
      if (res && res->ptr == srcs[s].ptr && res->bstride == srcs[s].bstride && p.res_block < 0) p.res_block = 2 * s;
    }
    NLAM_REQUIRE(!res || p.res_block >= 0, NLAM_E_UNSUPPORTED, "tc_rowmlp: residual must be one of the inputs");
  } else {
    p.nb1 = (mlp->in_dim + 31) / 32;
    p.n_elem = n_src;
    for (int s = 0; s < n_src; ++s) {
      p.esrc[s] = srcs[s].ptr;
      p.ebs[s] = srcs[s].bstride;
      p.edim[s] = srcs[s].dim;
    }
    p.k_real = mlp->in_dim;
    // 1-D bulk copies need 16-byte aligned slab addresses and sizes: rows % 4 == 0 makes every (128-row or tail)
    // slab a multiple of 16 bytes; batch strides must keep the alignment
    bool bulk = (n_rows % 4 == 0) && !getenv("NLAM_TC_NO_BULK");
    for (int s = 0; s < n_src; ++s)
      bulk = bulk && aligned16(srcs[s].ptr) && ((srcs[s].bstride * 4) % 16 == 0) && ((128LL * srcs[s].dim * 4) % 16 == 0);
    p.elem_bulk = bulk ? 1 : 0;
  }
  NLAM_REQUIRE(p.nb1 * 32 >= mlp->in_dim, NLAM_E_INVALID, "tc_rowmlp: width mismatch");
  p.n2 = nout <= 32 ? 32 : 64;
  p.nout = nout;
  int rc = weight_maps(mlp, p.nb1, p.n2, nout, &w1, &w2);
  if (rc) return rc;
  p.b1 = mlp->b[0];
  p.b2 = mlp->b[1];
  p.gamma = mlp->ln_gamma;
  p.beta = mlp->ln_beta;
  p.eps = mlp->ln_eps;
  p.out = out;
  p.n_rows = n_rows;
  p.B = B;
  p.n_tiles = (int)((n_rows + BM - 1) / BM);
  if (ep) {
    NLAM_REQUIRE(nout < 64 && !mlp->ln_gamma && !res && ep->prev && ep->std && ep->mean && (!ep->boundary || ep->mask),
                 NLAM_E_UNSUPPORTED, "tc_rowmlp: the fused step epilogue needs a narrow output without LayerNorm / residual");
    p.ep_prev = ep->prev;
    p.ep_bnd = ep->boundary;
    p.ep_mask = ep->mask;
    p.ep_std = ep->std;
    p.ep_mean = ep->mean;
  }
  CUtensorMap om;
  memset(&om, 0, sizeof(om));
  if (nout == 64) {
    rc = make_map(&om, out, 64, (uint64_t)n_rows, (uint64_t)B, 64, (uint64_t)n_rows * 64, BM, true);
    if (rc) return rc;
  }
  return launch(a0, a1, w1, w2, om, p, st, rowmlp_algorithmic_bytes(mlp, srcs, n_src, res, nullptr, n_rows, B, false, ep));
}

bool tc_edge_supported(const NlamGraph* g, const NlamMlp* edge_mlp, int flags) {
  int nout = 0;
  if (flags & NLAM_PROPAGATION) return false;
  if (!g || g->n_tiles <= 0) return false;
  if (!mlp_shape_ok(edge_mlp, &nout) || nout != 64 || edge_mlp->in_dim != 192 || !edge_mlp->ln_gamma) return false;
  return true;
}

int tc_edge(const NlamGraph* g, const NlamMlp* edge_mlp, const float* send, int64_t send_bs, const float* rec,
            int64_t rec_bs, const float* edge, int64_t edge_bs, float* edge_out, float* aggr_out, int B, int flags,
            cudaStream_t st, int64_t send_rows) {
  NLAM_REQUIRE(tc_edge_supported(g, edge_mlp, flags), NLAM_E_UNSUPPORTED, "tc_edge: unsupported shape");
  NLAM_REQUIRE(aligned16(send) && aligned16(rec) && aligned16(edge) && aligned16(aggr_out) &&
                   (!edge_out || aligned16(edge_out)) && send_bs % 4 == 0 && rec_bs % 4 == 0 && edge_bs % 4 == 0,
               NLAM_E_INVALID, "tc_edge: pointers / strides must be 16-byte aligned");
  TcParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap a0, a1, w1, w2;
  memset(&a1, 0, sizeof(a1));
  const bool batched = edge_bs != 0 && B > 1;
  int rc = make_map(&a0, edge, 64, (uint64_t)g->n_edges, batched ? (uint64_t)B : 1, 64,
                    batched ? (uint64_t)edge_bs : (uint64_t)g->n_edges * 64, BM, true);
  if (rc) return rc;
  rc = weight_maps(edge_mlp, 6, 64, 64, &w1, &w2);
  if (rc) return rc;
  p.mode_edge = 1;
  p.nb1 = 6;
  p.a0_blocks = 2;
  p.a0_batched = batched;
  p.gsrc[0] = send; p.gbs[0] = send_bs; p.gidx[0] = g->src;
  p.gsrc[1] = rec;  p.gbs[1] = rec_bs;  p.gidx[1] = g->dst;
  p.b1 = edge_mlp->b[0];
  p.b2 = edge_mlp->b[1];
  p.gamma = edge_mlp->ln_gamma;
  p.beta = edge_mlp->ln_beta;
  p.eps = edge_mlp->ln_eps;
  p.n2 = 64;
  p.nout = 64;
  p.res_block = 0;  // the e tile (needed for e' = e + m; harmless when edge_out is NULL)
  p.out = edge_out;
  p.aggr = aggr_out;
  p.mean = (flags & NLAM_AGGR_MEAN) ? 1 : 0;
  p.n_rows = g->n_edges;
  p.B = B;
  p.n_tiles = g->n_tiles;
  p.tile_rec = g->tile_rec;
  p.tile_e0 = g->tile_e0;
  p.tile_meta = reinterpret_cast<const int4*>(g->tile_meta);
  p.rowptr = g->rowptr;
  p.n_rec = g->n_rec;
  CUtensorMap om;
  memset(&om, 0, sizeof(om));
  if (edge_out) {
    rc = make_map(&om, edge_out, 64, (uint64_t)g->n_edges, (uint64_t)B, 64, (uint64_t)g->n_edges * 64, BM, true);
    if (rc) return rc;
  }
  // TMA gather4 maps over the sender / receiver node tensors: 2-D (64 x rows), box 32 x 1.  Usable when
  // the batches are dense (stride = rows*64) or broadcast (stride 0); otherwise the cp.async path runs.
  CUtensorMap g0, g1;
  static int g4_env = -1;
  if (g4_env < 0) g4_env = getenv("NLAM_TC_NO_GATHER4") ? 0 : 1;
  const int64_t ns = g->n_send, nr = g->n_rec;
  const bool dense0 = (send_bs == 0 || B == 1 || send_bs == send_rows * 64);
  const bool dense1 = (rec_bs == 0 || B == 1 || rec_bs == nr * 64);
  if (g4_env && dense0 && dense1 && send_rows >= ns) {
    const uint64_t rows0 = (uint64_t)send_rows * ((send_bs == 0 || B == 1) ? 1 : B);
    const uint64_t rows1 = (uint64_t)nr * ((rec_bs == 0 || B == 1) ? 1 : B);
    rc = make_map(&g0, send, 64, rows0, 1, 64, 0, 1, false);
    if (rc) return rc;
    rc = make_map(&g1, rec, 64, rows1, 1, 64, 0, 1, false);
    if (rc) return rc;
    p.use_g4 = 1;
    p.g4_rows[0] = (send_bs == 0 || B == 1) ? 0 : (int)send_rows;
    p.g4_rows[1] = (rec_bs == 0 || B == 1) ? 0 : (int)nr;
    return launch(a0, a1, w1, w2, om, p, st, edge_algorithmic_bytes(g, B, send_bs, rec_bs, edge_bs, edge_out != nullptr, 64), &g0, &g1);
  }
  return launch(a0, a1, w1, w2, om, p, st, edge_algorithmic_bytes(g, B, send_bs, rec_bs, edge_bs, edge_out != nullptr, 64));
}

}  // namespace nlam
