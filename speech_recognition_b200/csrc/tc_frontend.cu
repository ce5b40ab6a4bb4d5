// Tensor-core tier of the STFT front end (DFT-as-GEMM on tcgen05).  Until the split-fp16
// tcgen05 DFT kernel lands, the TC tier runs the fp32 CUDA-core GEMM chain of frontend.cu
// (still a CUDA path; there is no CPU fallback anywhere).
#include "common.cuh"
namespace kws {
int frontend_build_tc(kws_handle*, const std::vector<float>&) { return KWS_OK; }
int launch_features_tc(kws_handle* h, const float* wav, int B, int kind, float* out, cudaStream_t st) {
  return launch_features_f32(h, wav, B, kind, out, st);
}
}  // namespace kws
