// Micro-benchmark: how fast can "one warp walks one line of a [H][W][D] volume" stream on this GPU, as a function of
// the request width (bytes per lane) and of the bytes each warp keeps in flight?  (Development aid for DESIGN.md.)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

// MODE 0: every line walks x = 0..W-1; 1: odd lines walk backwards; 2: line l starts at x = (l * 37) % W and wraps
// (2 only probes the memory system: a prefix cannot start mid-line)
__device__ int g_mode = 0;
__device__ long long g_pitch = 0;  // row pitch in floats (0 = dense W*D)
template <int VW, int BATCH>  // VW floats per lane per position (1,2,4), BATCH positions in flight per warp
__global__ void __launch_bounds__(256) walk_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int D,
                                                   int smem_pad) {
    extern __shared__ float pad[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dchunk = 32 * VW, nchunk = D / dchunk;
    const long long wg = (long long)blockIdx.x * 8 + warp;
    const int line = (int)(wg / nchunk), chunk = (int)(wg % nchunk);
    if (line >= H) return;
    const long long pitch = g_pitch ? g_pitch : (long long)W * D;
    const float* src = in + (long long)line * pitch + chunk * dchunk + lane * VW;
    float* dst = out + (long long)line * pitch + chunk * dchunk + lane * VW;
    if (smem_pad < 0) pad[threadIdx.x] = 0;  // keep the allocation alive
    const int mode = g_mode;
    for (int x0 = 0; x0 < W; x0 += BATCH) {
        float v[BATCH][VW];
        int xs[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            int x = x0 + k;
            if (mode == 1 && (line & 1)) x = W - 1 - x;
            if (mode == 2) x = (x + line * 37) % W;
            xs[k] = x;
            const float* p = src + (long long)x * D;
            if (VW == 4) { float4 t = __ldcs((const float4*)p); v[k][0] = t.x; v[k][1 % VW] = t.y; v[k][2 % VW] = t.z; v[k][3 % VW] = t.w; }
            else if (VW == 2) { float2 t = __ldcs((const float2*)p); v[k][0] = t.x; v[k][1 % VW] = t.y; }
            else v[k][0] = __ldcs(p);
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            float* q = dst + (long long)xs[k] * D;
            if (VW == 4) __stcs((float4*)q, make_float4(v[k][0] + 1.f, v[k][1 % VW], v[k][2 % VW], v[k][3 % VW]));
            else if (VW == 2) __stcs((float2*)q, make_float2(v[k][0] + 1.f, v[k][1 % VW]));
            else __stcs(q, v[k][0] + 1.f);
        }
    }
}

template <int VW, int BATCH>
void run(const float* in, float* out, int H, int W, int D, int smem, const char* tag) {
    const int nchunk = D / (32 * VW);
    const long long nw = (long long)H * nchunk;
    const unsigned grid = (unsigned)((nw + 7) / 8);
    cudaFuncSetAttribute(walk_kernel<VW, BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    walk_kernel<VW, BATCH><<<grid, 256, smem>>>(in, out, H, W, D, smem);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) walk_kernel<VW, BATCH><<<grid, 256, smem>>>(in, out, H, W, D, smem);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    const double gb = 2.0 * H * (double)W * D * 4 / 1e9;
    printf("%-28s VW=%d BATCH=%2d smem/CTA=%3d KB grid=%6u : %.3f ms  %.0f GB/s (R+W)  err=%s\n", tag, VW, BATCH, smem / 1024, grid, ms, gb / ms * 1e3,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int H = 1080, W = 1920, D = 256;
    const size_t n = (size_t)H * W * D;
    float *in, *out;
    cudaMalloc(&in, n * 4 + (size_t)H * 16416 * 4); cudaMalloc(&out, n * 4 + (size_t)H * 16416 * 4);
    cudaMemset(in, 0, n * 4);
    // occupancy limited to 1 CTA (8 warps) per SM by a 200 KB shared allocation, like agg_seq_kernel
    for (int m = 0; m < 3; ++m) {
        cudaMemcpyToSymbol(g_mode, &m, sizeof(int));
        printf("-- mode %d\n", m);
        run<1, 32>(in, out, H, W, D, 200 * 1024, "8 warps/SM, 128B requests");
        run<4, 8>(in, out, H, W, D, 24 * 1024, "64 warps/SM, 512B requests");
    }
    { int m = 0; cudaMemcpyToSymbol(g_mode, &m, sizeof(int)); }
    for (long long extra : {0ll, 32ll, 256ll, 1184ll, 4128ll, 16416ll}) {
        long long pitch = (long long)W * D + extra;
        cudaMemcpyToSymbol(g_pitch, &pitch, sizeof(long long));
        printf("-- row pitch = W*D + %lld floats\n", extra);
        run<1, 32>(in, out, H, W, D, 200 * 1024, "8 warps/SM, 128B requests");
    }
    { long long z = 0; cudaMemcpyToSymbol(g_pitch, &z, sizeof(long long)); }
    run<1, 32>(in, out, H, W, D, 200 * 1024, "8 warps/SM, 128B requests");
    run<1, 64>(in, out, H, W, D, 200 * 1024, "8 warps/SM, 128B requests");
    run<2, 32>(in, out, H, W, D, 200 * 1024, "8 warps/SM, 256B requests");
    run<4, 16>(in, out, H, W, D, 200 * 1024, "8 warps/SM, 512B requests");
    run<4, 32>(in, out, H, W, D, 200 * 1024, "8 warps/SM, 512B requests");
    run<1, 32>(in, out, H, W, D, 100 * 1024, "16 warps/SM, 128B requests");
    run<1, 32>(in, out, H, W, D, 48 * 1024, "32 warps/SM, 128B requests");
    run<1, 16>(in, out, H, W, D, 24 * 1024, "64 warps/SM, 128B requests");
    run<4, 8>(in, out, H, W, D, 24 * 1024, "64 warps/SM, 512B requests");
    cudaMemcpy(out, in, n * 4, cudaMemcpyDeviceToDevice);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) cudaMemcpyAsync(out, in, n * 4, cudaMemcpyDeviceToDevice);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    printf("cudaMemcpy D2D: %.3f ms %.0f GB/s (R+W)\n", ms, 2.0 * n * 4 / 1e9 / ms * 1e3);
    return 0;
}
