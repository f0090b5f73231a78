// cge_landmarks.cuh -- interface of cge_landmarks.cu
#pragma once
#include <cuda_runtime.h>

namespace cge {
// all pointers are device pointers; oa / ob / ow have room for m cells
cudaError_t landmarks_aggregate_device(int n, int d, int N, int base, const long long *lm,
                                       const double *vw, const long long *comm, const double *x,
                                       long long m, const long long *src, const long long *dst,
                                       const double *ew, int directed, double *embed, double *lweight,
                                       double *dii, long long *cluster, long long *oa, long long *ob,
                                       double *ow, int *n_cells, int *n_bad, cudaStream_t st);
}  // namespace cge
