// Host-side entry points of the conv_chain_kernel instantiations (chain_inst.cu, one object per GroupNorm width).
#pragma once
#include <cuda_runtime.h>

namespace dad {

struct ChainArgs;

struct ChainOps {
  cudaError_t (*launch)(int mh, int ns, int grid, int smem, cudaStream_t st, const ChainArgs &a);
  cudaError_t (*prepare)(int mh, int ns, int max_optin);          // opt-in shared memory size, once per process
  cudaError_t (*max_clusters)(int mh, int ns, int smem, int *out); // co-resident 2-CTA clusters (cudaOccupancyMaxActiveClusters)
};

extern const ChainOps chain_ops_16, chain_ops_32, chain_ops_64, chain_ops_128, chain_ops_256;

inline const ChainOps *chain_ops(int gw) {
  switch (gw) {
    case 16: return &chain_ops_16;
    case 32: return &chain_ops_32;
    case 64: return &chain_ops_64;
    case 128: return &chain_ops_128;
    case 256: return &chain_ops_256;
  }
  return nullptr;
}

}  // namespace dad
