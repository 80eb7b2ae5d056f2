// placeholder: tcgen05 kernels not built yet
#include "common.cuh"
namespace nlam {
bool tc_rowmlp_supported(const NlamMlp*, const NlamRowSrc*, int, const NlamRowSrc*, const NlamRowSrc*, int64_t) { return false; }
int tc_rowmlp(const NlamMlp*, const NlamRowSrc*, int, const NlamRowSrc*, float*, int64_t, int, cudaStream_t) { set_error("tc_rowmlp: not built"); return NLAM_E_UNSUPPORTED; }
bool tc_edge_supported(const NlamGraph*, const NlamMlp*, int) { return false; }
int tc_edge(const NlamGraph*, const NlamMlp*, const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, float*, int, int, cudaStream_t) { set_error("tc_edge: not built"); return NLAM_E_UNSUPPORTED; }
}
