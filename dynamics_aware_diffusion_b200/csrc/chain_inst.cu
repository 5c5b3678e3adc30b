// Instantiations of conv_chain_kernel for ONE GroupNorm width (-DCHAIN_GW=16|32|64|128|256): one translation unit per
// width so that the build compiles them in parallel.  dad_api.cu dispatches through chain_ops(gw).
#include <cuda.h>
#include <cuda_runtime.h>

#include "chain_host.h"
#include "conv_chain.cuh"
#include "launch.cuh"

#ifndef CHAIN_GW
#error "compile with -DCHAIN_GW=<GroupNorm width>"
#endif

namespace dad {
namespace {

template <int MH, int NS>
cudaError_t launch_one(int grid, int smem, cudaStream_t st, const ChainArgs &a) {
  // the 9 KB argument block is passed by reference to the runtime (no by-value copies on the host stack)
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(ch_threads(CHAIN_GW));
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  launch_attrs(at, n, 2);
  cfg.attrs = at;
  cfg.numAttrs = (unsigned)n;
  void *params[1] = {const_cast<ChainArgs *>(&a)};
  return cudaLaunchKernelExC(&cfg, reinterpret_cast<const void *>(conv_chain_kernel<CHAIN_GW, MH, NS>), params);
}

template <int MH, int NS>
cudaError_t prepare_one(int max_optin) {
  return cudaFuncSetAttribute(conv_chain_kernel<CHAIN_GW, MH, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin);
}

template <int MH, int NS>
cudaError_t clusters_one(int smem, int *out) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(ch_threads(CHAIN_GW));
  cfg.dynamicSmemBytes = (size_t)smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaOccupancyMaxActiveClusters(out, conv_chain_kernel<CHAIN_GW, MH, NS>, &cfg);
}

#if CHAIN_GW == 256
#define CHAIN_DISPATCH(fn, mh, ns, ...) ((mh) == 1 && (ns) == 2 ? fn<1, 2>(__VA_ARGS__) : cudaErrorInvalidValue)
#else
#define CHAIN_DISPATCH(fn, mh, ns, ...)                                  \
  ((mh) == 1 && (ns) == 1   ? fn<1, 1>(__VA_ARGS__)                      \
   : (mh) == 2 && (ns) == 1 ? fn<2, 1>(__VA_ARGS__)                      \
   : (mh) == 1 && (ns) == 2 ? fn<1, 2>(__VA_ARGS__)                      \
                            : cudaErrorInvalidValue)
#endif

cudaError_t do_launch(int mh, int ns, int grid, int smem, cudaStream_t st, const ChainArgs &a) {
  return CHAIN_DISPATCH(launch_one, mh, ns, grid, smem, st, a);
}
cudaError_t do_prepare(int mh, int ns, int max_optin) { return CHAIN_DISPATCH(prepare_one, mh, ns, max_optin); }
cudaError_t do_clusters(int mh, int ns, int smem, int *out) { return CHAIN_DISPATCH(clusters_one, mh, ns, smem, out); }

}  // namespace

#define CHAIN_CAT2(a, b) a##b
#define CHAIN_CAT(a, b) CHAIN_CAT2(a, b)
const ChainOps CHAIN_CAT(chain_ops_, CHAIN_GW) = {do_launch, do_prepare, do_clusters};

}  // namespace dad
