// C-ABI entry points that dispatch between the tcgen05 (TF32) kernels and the exact fp32
// kernels.  See include/nlam_b200.h for the contract and the reference interfaces replaced.
#include "common.cuh"

using namespace nlam;

static bool want_tf32(int flags) { return !(flags & NLAM_MATH_FP32); }

extern "C" int nlam_rowmlp_fwd(const NlamMlp* mlp, const NlamRowSrc* srcs, int n_src, const NlamRowSrc* res,
                               const NlamRowSrc* res2, float* out, float* out2, int64_t n_rows, int B, int flags,
                               void* stream) {
  NLAM_REQUIRE(mlp && srcs && out, NLAM_E_INVALID, "nlam_rowmlp_fwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (want_tf32(flags) && tc_rowmlp_supported(mlp, srcs, n_src, res, res2, n_rows))
    return tc_rowmlp(mlp, srcs, n_src, res, out, n_rows, B, st);
  NLAM_REQUIRE(!(flags & NLAM_MATH_TF32), NLAM_E_UNSUPPORTED,
               "nlam_rowmlp_fwd: shape not supported by the tcgen05 kernels (in=%d)", mlp->in_dim);
  return rowmlp_simt(mlp, srcs, n_src, res, res2, out, out2, n_rows, B, st);
}

extern "C" int nlam_rowmlp_step_fwd(const NlamMlp* mlp, const NlamRowSrc* srcs, int n_src, const float* prev,
                                    const float* boundary, const float* bmask, const float* diff_std,
                                    const float* diff_mean, float* new_state, int64_t n_rows, int B, int flags,
                                    void* stream) {
  NLAM_REQUIRE(mlp && srcs && prev && diff_std && diff_mean && new_state, NLAM_E_INVALID, "nlam_rowmlp_step_fwd: null argument");
  NLAM_REQUIRE((boundary == nullptr) || bmask, NLAM_E_INVALID, "nlam_rowmlp_step_fwd: boundary without mask");
  const int nout = mlp->out_dim[mlp->n_linear - 1];
  // only the tensor-core path fuses the epilogue; callers fall back to nlam_rowmlp_fwd + nlam_step_epilogue
  NLAM_REQUIRE(want_tf32(flags) && nout < 64 && !mlp->ln_gamma && tc_rowmlp_supported(mlp, srcs, n_src, nullptr, nullptr, n_rows),
               NLAM_E_UNSUPPORTED, "nlam_rowmlp_step_fwd: shape / math mode not covered by the fused kernel");
  StepEpilogue ep = {prev, boundary, bmask, diff_std, diff_mean};
  return tc_rowmlp(mlp, srcs, n_src, nullptr, new_state, n_rows, B, (cudaStream_t)stream, &ep);
}

extern "C" size_t nlam_inet_workspace_bytes(const NlamGraph* g, int B, int H, int flags) {
  (void)flags;
  if (!g) return 0;
  // message buffer (exact path) + aggregate buffer; 256-byte aligned sections
  size_t msg = ((size_t)B * g->n_edges * H * sizeof(float) + 255) / 256 * 256;
  size_t agg = ((size_t)B * g->n_rec * H * sizeof(float) + 255) / 256 * 256;
  size_t proj = (H == 64) ? (tc_edge2_workspace_floats(g, B, 0) * sizeof(float) + 255) / 256 * 256 : 0;
  return msg + agg + proj;
}

extern "C" int nlam_inet_fwd(const NlamGraph* g, const NlamMlp* edge_mlp, const NlamMlp* aggr_mlp,
                             const float* send, int64_t send_bs, const float* rec, int64_t rec_bs,
                             const float* edge, int64_t edge_bs, float* rec_out, float* edge_out,
                             float* aggr_out, int B, int flags, void* workspace, size_t ws_bytes,
                             void* stream) {
  NLAM_REQUIRE(g && edge_mlp && aggr_mlp && send && rec && edge && rec_out, NLAM_E_INVALID, "nlam_inet_fwd: null argument");
  NLAM_REQUIRE(B >= 1, NLAM_E_INVALID, "nlam_inet_fwd: bad batch");
  const int H = edge_mlp->out_dim[edge_mlp->n_linear - 1];
  NLAM_REQUIRE(edge_mlp->in_dim == 3 * H && aggr_mlp->in_dim == 2 * H &&
                   aggr_mlp->out_dim[aggr_mlp->n_linear - 1] == H,
               NLAM_E_INVALID, "nlam_inet_fwd: MLP widths inconsistent with H=%d", H);
  cudaStream_t st = (cudaStream_t)stream;
  const int mean = (flags & (NLAM_AGGR_MEAN | NLAM_PROPAGATION)) ? 1 : 0;  // PropagationNet forces mean
  const bool prop = flags & NLAM_PROPAGATION;
  const size_t msg_bytes = ((size_t)B * g->n_edges * H * sizeof(float) + 255) / 256 * 256;
  const size_t agg_bytes = ((size_t)B * g->n_rec * H * sizeof(float) + 255) / 256 * 256;

  const bool use_tc = want_tf32(flags) && tc_edge_supported(g, edge_mlp, flags);
  NLAM_REQUIRE(use_tc || !(flags & NLAM_MATH_TF32), NLAM_E_UNSUPPORTED,
               "nlam_inet_fwd: shape (H=%d, max in-degree %d, hidden_layers=%d) not supported by the tcgen05 kernels",
               H, g->max_in_degree, edge_mlp->n_linear - 1);

  float* aggr = aggr_out;
  if (!aggr) {
    NLAM_REQUIRE(workspace && ws_bytes >= msg_bytes + agg_bytes, NLAM_E_WORKSPACE, "nlam_inet_fwd: workspace too small");
    aggr = (float*)((char*)workspace + msg_bytes);
  }
  const int64_t aggr_bs = (int64_t)g->n_rec * H;

  if (use_tc) {
    // rows of the sender tensor: known exactly when the batches are dense, else at least n_send
    const int64_t send_rows = (B > 1 && send_bs > 0) ? send_bs / H : g->n_send;
    const size_t proj_bytes = tc_edge2_workspace_floats(g, B, 0) * sizeof(float);
    int rc;
    if (tc_ell_supported(g, edge_mlp, flags, send, send_bs, rec, rec_bs, edge_out != nullptr) && workspace &&
        ws_bytes >= msg_bytes + agg_bytes + proj_bytes) {
      // uniform in-degree: receiver-tiled ELL kernel, aggregation in registers (tc3.cu)
      rc = tc_ell_edge(g, edge_mlp, send, send_bs, rec, rec_bs, edge, edge_bs, aggr, B, flags, st,
                       (float*)((char*)workspace + msg_bytes + agg_bytes));
    } else if (tc_edge2_supported(g, edge_mlp, flags, send, send_bs, rec, rec_bs, B, send_rows) && workspace &&
        ws_bytes >= msg_bytes + agg_bytes + proj_bytes) {
      // split first Linear: node projections + K=64 edge kernel (tc2.cu)
      float* proj = (float*)((char*)workspace + msg_bytes + agg_bytes);
      rc = tc_edge3_enabled()
               ? tc_edge3(g, edge_mlp, send, send_bs, rec, rec_bs, edge, edge_bs, edge_out, aggr, B, flags, st, proj)
               : tc_edge2(g, edge_mlp, send, send_bs, rec, rec_bs, edge, edge_bs, edge_out, aggr, B, flags, st, send_rows, proj);
    } else if (tc_edge_bcast_supported(g, edge_mlp, flags, send, send_bs, rec, rec_bs, edge, edge_bs, B, edge_out != nullptr) &&
               workspace && ws_bytes >= msg_bytes + agg_bytes + proj_bytes) {
      // batch-broadcast edge features, large sender set, no edge update (grid -> mesh): tile-major kernel with the
      // edge term resident in TMEM and raw sender rows gathered (tc6.cu)
      rc = tc_edge_bcast(g, edge_mlp, send, send_bs, rec, rec_bs, edge, aggr, B, flags, st,
                         (float*)((char*)workspace + msg_bytes + agg_bytes));
    } else {
      rc = tc_edge(g, edge_mlp, send, send_bs, rec, rec_bs, edge, edge_bs, edge_out, aggr, B, flags, st, send_rows);
    }
    if (rc) return rc;
  } else {
    NLAM_REQUIRE(workspace && ws_bytes >= msg_bytes, NLAM_E_WORKSPACE, "nlam_inet_fwd: workspace too small");
    float* msg = (float*)workspace;
    NlamRowSrc srcs[3] = {{edge, nullptr, edge_bs, H, 0}, {send, g->src, send_bs, H, 0}, {rec, g->dst, rec_bs, H, 0}};
    NlamRowSrc res = {send, g->src, send_bs, H, 0};
    NlamRowSrc res2 = {edge, nullptr, edge_bs, H, 0};
    int rc = rowmlp_simt(edge_mlp, srcs, 3, prop ? &res : nullptr, edge_out ? &res2 : nullptr, msg, edge_out,
                         g->n_edges, B, st);
    if (rc) return rc;
    rc = nlam_segment_sum(g->rowptr, nullptr, g->n_rec, msg, (int64_t)g->n_edges * H, aggr, aggr_bs, B, H, mean, st);
    if (rc) return rc;
  }
  // node update: rec' = base + aggr_mlp(cat(rec, aggr)); base = rec (InteractionNet) or aggr (PropagationNet)
  NlamRowSrc nsrcs[2] = {{rec, nullptr, rec_bs, H, 0}, {aggr, nullptr, aggr_bs, H, 0}};
  NlamRowSrc nres = prop ? nsrcs[1] : nsrcs[0];
  if (use_tc && tc_rowmlp_supported(aggr_mlp, nsrcs, 2, &nres, nullptr, g->n_rec))
    return tc_rowmlp(aggr_mlp, nsrcs, 2, &nres, rec_out, g->n_rec, B, st);
  return rowmlp_simt(aggr_mlp, nsrcs, 2, &nres, nullptr, rec_out, nullptr, g->n_rec, B, st);
}
