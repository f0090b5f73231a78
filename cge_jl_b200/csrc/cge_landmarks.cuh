// cge_landmarks.cuh -- interface of cge_landmarks.cu
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace cge {
// all pointers are device pointers; oa / ob / ow have room for m cells
cudaError_t landmarks_aggregate_device(int n, int d, int N, int base, const long long *lm,
                                       const double *vw, const long long *comm, const double *x,
                                       long long m, const long long *src, const long long *dst,
                                       const double *ew, int directed, double *embed, double *lweight,
                                       double *dii, long long *cluster, long long *oa, long long *ob,
                                       double *ow, int *n_cells, int *n_bad, cudaStream_t st);
// SURVEY.md 8(f) F4 (cge_select.cu): runsplit with the cuts on the device.  Host pointers; clusters as
// CSR over 0-based vertex ids in the order of sort(initial_clusters); rule 0 rss, 2 size, 3 diameter.
// Returns 0, -1 (CUDA failure) or -2 (an error the reference raises); msg says which.
// eig (optional): fills v with the principal axis of the symmetric row-major d x d matrix c, returns 0.
typedef int (*SelectEigFn)(const double *c, long long d, double *v, void *user);
int landmarks_select_device(int device, cudaStream_t st, long long n, int d, const double *x_rowmajor,
                            const double *vweights, long long n_clusters, const long long *cl_ptr,
                            const int *cl_members, long long land, long long forced, int rule,
                            SelectEigFn eig, void *eig_user, long long *out_group, long long *out_cuts,
                            std::string &msg);
// landmarks.jl:369 -- size(unique(embedding, dims=1), 1) on the device (host pointer in, count out)
int count_unique_rows_device(cudaStream_t st, long long n, int d, const double *x_rowmajor,
                             long long *out_count, std::string &msg);
// unit-length eigenvector of the largest eigenvalue of a symmetric d x d matrix (upper triangle read),
// largest-magnitude component positive
void sym_top_eigvec(const double *a, int d, double *v_out, double *lambda_out);
}  // namespace cge
