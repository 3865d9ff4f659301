// Lock-free union-find on a label array (label[p] == p for roots, -1 for pixels outside every component), shared by
// the two speckle filters (postproc.cu: RemoveSpeckles, PostProcessing.h:250; display.cu: cv::filterSpeckles).
#pragma once
#include <cstddef>

__device__ __forceinline__ int uf_find(volatile int* L, int a) {
    int b = L[a];
    while (b != a) {
        a = b;
        b = L[a];
    }
    return a;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) {
            int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// sizes of the components, accumulated at their roots; labels are flattened to roots on the way
static __global__ void __launch_bounds__(256) uf_count_kernel(int* __restrict__ label, size_t n, int* __restrict__ size) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int root = -1;
    if (p < n && label[p] >= 0) {
        root = uf_find(label, (int)p);
        label[p] = root;  // roots keep label[root]==root, so concurrent finds stay valid
    }
    // one atomic per distinct root per warp (large regions would otherwise serialise on a single address)
    const unsigned peers = __match_any_sync(0xffffffffu, root);
    if (root >= 0 && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&size[root], __popc(peers));
}
