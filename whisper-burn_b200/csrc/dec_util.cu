// Beam ancestry tables of the decoding session: beams address the self K/V history of their ancestors through
// anc[row][position] = cache row that holds that position, instead of copying caches when the search reorders beams
// (what the reference gets for free by re-running the whole prefix per beam, src/transcribe.rs:253-270).
#include "decoder.h"

namespace wb {

namespace {

// anc_new[r][0..p) = anc_old[parent[r]][0..p), anc_new[r][p] = r
__global__ void dec_reorder_kernel(const int* __restrict__ anc_old, int* __restrict__ anc_new, const int* __restrict__ parent,
                                   const int* __restrict__ pos_ptr, int t_max) {
    const int r = blockIdx.x;
    const int p = *pos_ptr;
    const int* src = anc_old + (int64_t)parent[r] * t_max;
    int* dst = anc_new + (int64_t)r * t_max;
    for (int j = threadIdx.x; j < p; j += blockDim.x) dst[j] = src[j];
    if (threadIdx.x == 0) dst[p] = r;
}

__global__ void dec_anc_identity_kernel(int* __restrict__ anc, int t_max) {
    for (int j = threadIdx.x; j < t_max; j += blockDim.x) anc[(int64_t)blockIdx.x * t_max + j] = blockIdx.x;
}

}  // namespace

void launch_dec_anc_identity(int* anc, int R, int t_max, cudaStream_t st) {
    dec_anc_identity_kernel<<<R, 128, 0, st>>>(anc, t_max);
    WB_LAUNCH_CHECK();
}

void launch_dec_reorder(const int* anc_old, int* anc_new, const int* parent, const int* pos_ptr, int R, int t_max, cudaStream_t st) {
    dec_reorder_kernel<<<R, 128, 0, st>>>(anc_old, anc_new, parent, pos_ptr, t_max);
    WB_LAUNCH_CHECK();
}

}  // namespace wb
