// Host-callable launchers of the phase-function table kernels (tables.cu).
#pragma once
#include <cuda_runtime.h>

namespace i3rc {

struct PhaseTableDev {
  int kind;                      // 1 Legendre, 2 tabulated
  int nEntries;
  const int* offsets;            // device [nEntries+1]
  const float* coefs;            // device
  int nAngles;
  const float* angles;           // device [nAngles]
  const float* values;           // device [nEntries][nAngles] (normalised)
  int maxNodes;                  // max quadrature nodes of any entry (stride of the scratch arrays)
  const int* hostNodesPerEntry;  // host [nEntries]
};

cudaError_t normalize_tabulated(const float* dAngles, float* dValues, int nAngles, int nEntries, int repeats,
                                cudaStream_t st);
cudaError_t build_inverse_table(const PhaseTableDev& t, int nSteps, float* dOut, cudaStream_t st);
cudaError_t build_forward_table(const PhaseTableDev& t, int nSteps, float* dOut, cudaStream_t st);
cudaError_t build_hybrid_table(const float* dOrig, float* dOut, int nSteps, int nEntries, float widthDeg,
                               cudaStream_t st);

}  // namespace i3rc
