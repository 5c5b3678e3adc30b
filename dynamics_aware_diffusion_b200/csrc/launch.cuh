// Kernel launch helper shared by the translation units of the library.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>

namespace dad {

// Ablation / tuning switches exist only in -DDAD_TUNING builds (python -m dynamics_aware_diffusion_b200.build -DDAD_TUNING
// --out=...): the shipping library never reads the environment, so a stray variable cannot change its kernels.
inline int tuning_env(const char *name, int dflt) {
#ifdef DAD_TUNING
  const char *v = getenv(name);
  return v ? atoi(v) : dflt;
#else
  (void)name;
  return dflt;
#endif
}

// Every kernel of the sampling step is launched with programmatic stream serialization: it may start (and run
// its set-up) while its predecessor drains, and calls griddepcontrol.wait before touching global data.
// (DAD_PDL=0 in a -DDAD_TUNING build disables the attribute: plain stream order.)
inline bool pdl_enabled() {
  static const bool on = tuning_env("DAD_PDL", 1) != 0;
  return on;
}

inline void launch_attrs(cudaLaunchAttribute *at, int &n, int cluster) {
  n = 0;
  if (cluster > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
}

template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                     Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  launch_attrs(at, n, cluster);
  cfg.attrs = at;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace dad
