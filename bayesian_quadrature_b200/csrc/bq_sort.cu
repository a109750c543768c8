// Device-side sort of a query vector (CUB radix sort of (x, index) pairs).
//
// Band skipping (bq_score.cu) works on the hull of the 32 query points a warp owns: for a sorted vector (a grid) that
// hull is tiny and only a short band of observations is relevant; for points in arbitrary order the hull of every warp
// spans the whole domain and nothing can be skipped -- measured 2.2x / 2.7x / 5.1x slower at ns = 64 / 128 / 256.
// Sorting 10^6 doubles takes 0.2 ms, so the host entry points sort vectors that do not look sorted, score them in
// ascending order and write every result back to its original position through the permutation (ScoreArgs::perm).
#include <cub/device/device_radix_sort.cuh>

#include "bq_common.cuh"

namespace bqb {

__global__ void iota_kernel(int *v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// temp == nullptr: only the size query.  perm[p] = original index of the p-th smallest point.
cudaError_t sort_points(const double *d_x, int n, double *d_x_sorted, int *d_iota, int *d_perm, void *temp, size_t *temp_bytes,
                        cudaStream_t s) {
    if (!temp) return cub::DeviceRadixSort::SortPairs(nullptr, *temp_bytes, d_x, d_x_sorted, d_iota, d_perm, n, 0, 64, s);
    iota_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_iota, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cub::DeviceRadixSort::SortPairs(temp, *temp_bytes, d_x, d_x_sorted, d_iota, d_perm, n, 0, 64, s);
}

}  // namespace bqb
