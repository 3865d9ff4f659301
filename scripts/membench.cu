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
    if (mode == 3) {
        // "tiled" order: blocks in launch order take adjacent BATCH-position tiles of the volume (rows are contiguous, so
        // this is a linear sweep of memory by the whole GPU) instead of one far-apart line each
        const long long ntile_row = (W + BATCH - 1) / BATCH;
        for (long long t = blockIdx.x; t < (long long)H * ntile_row; t += gridDim.x) {
            const int row = (int)(t / ntile_row), x0 = (int)(t % ntile_row) * BATCH;
            const int ch = warp % nchunk;
            if (warp >= nchunk) break;
            const float* s2 = in + (long long)row * pitch + ch * dchunk + lane * VW;
            float* d2 = out + (long long)row * pitch + ch * dchunk + lane * VW;
            float v[BATCH][VW];
#pragma unroll
            for (int k = 0; k < BATCH; ++k) {
                const float* p = s2 + (long long)min(x0 + k, W - 1) * D;
                if (VW == 4) { float4 t4 = __ldcs((const float4*)p); v[k][0] = t4.x; v[k][1 % VW] = t4.y; v[k][2 % VW] = t4.z; v[k][3 % VW] = t4.w; }
                else if (VW == 2) { float2 t2 = __ldcs((const float2*)p); v[k][0] = t2.x; v[k][1 % VW] = t2.y; }
                else v[k][0] = __ldcs(p);
            }
#pragma unroll
            for (int k = 0; k < BATCH; ++k) {
                float* q = d2 + (long long)min(x0 + k, W - 1) * D;
                if (VW == 4) __stcs((float4*)q, make_float4(v[k][0] + 1.f, v[k][1 % VW], v[k][2 % VW], v[k][3 % VW]));
                else if (VW == 2) __stcs((float2*)q, make_float2(v[k][0] + 1.f, v[k][1 % VW]));
                else __stcs(q, v[k][0] + 1.f);
            }
        }
        return;
    }
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

// L2 residency probe: `lines` warps walk lines of W positions x 32 floats of ONE buffer in place (read position x, write it
// back), `reps` launches back to back.  Small buffers stay in the 126 MB L2 between launches.
template <int BATCH, int HINT>
__global__ void __launch_bounds__(256) walk_inplace_kernel(float* __restrict__ buf, int lines, int W) {
    extern __shared__ float pad[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int line = blockIdx.x * 8 + warp;
    if (line >= lines) return;
    if (W < 0) pad[threadIdx.x] = 0;
    float* p = buf + (long long)line * W * 32 + lane;
    for (int x0 = 0; x0 < W; x0 += BATCH) {
        float v[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const float* q = p + (long long)min(x0 + k, W - 1) * 32;
            v[k] = HINT ? __ldcs(q) : __ldcg(q);
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            float* q = p + (long long)min(x0 + k, W - 1) * 32;
            if (HINT) __stcs(q, v[k] + 1.f); else __stcg(q, v[k] + 1.f);
        }
    }
}
template <int BATCH, int HINT>
void run_inplace(float* buf, double mb, int smem, const char* tag) {
    const int lines = 148 * 8 * (200 * 1024 / smem);
    const int W = (int)(mb * 1e6 / (lines * 128.0));
    cudaFuncSetAttribute(walk_inplace_kernel<BATCH, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const unsigned grid = (lines + 7) / 8;
    for (int r = 0; r < 2; ++r) walk_inplace_kernel<BATCH, HINT><<<grid, 256, smem>>>(buf, lines, W);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 8; ++r) walk_inplace_kernel<BATCH, HINT><<<grid, 256, smem>>>(buf, lines, W);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 8;
    const double gb = 2.0 * lines * (double)W * 128 / 1e9;
    printf("in-place %-22s hint=%s buffer=%6.1f MB lines=%d W=%d : %.4f ms/launch  %.0f GB/s (R+W)  err=%s\n", tag, HINT ? "cs" : "cg",
           lines * (double)W * 128 / 1e6, lines, W, ms, gb / ms * 1e3, cudaGetErrorString(cudaGetLastError()));
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

int main(int argc, char** argv) {
    const int H = 1080, W = 1920, D = 256;
    const size_t n = (size_t)H * W * D;
    float *in, *out;
    cudaMalloc(&in, n * 4 + (size_t)H * 16416 * 4); cudaMalloc(&out, n * 4 + (size_t)H * 16416 * 4);
    cudaMemset(in, 0, n * 4);
    if (argc > 1 && argv[1][0] == 'p') {  // in-place vs out-of-place at the aggregation's geometry (rows, then columns as lines)
        for (int rep = 0; rep < 2; ++rep) {
            run<1, 32>(in, out, H, W, D, 200 * 1024, "rows  out-of-place 8w");
            run<1, 32>(in, in, H, W, D, 200 * 1024, "rows  in-place     8w");
            run<1, 32>(in, out, W, H, D, 200 * 1024, "cols* out-of-place 8w");
            run<1, 32>(in, in, W, H, D, 200 * 1024, "cols* in-place     8w");
            run<1, 32>(in, out, H, W, D, 100 * 1024, "rows  out-of-place 16w");
            run<1, 32>(in, in, H, W, D, 100 * 1024, "rows  in-place     16w");
        }
        return 0;
    }
    if (argc > 1 && argv[1][0] == 'l') {  // "l2": residency probe only
        for (double mb : {8.0, 16.0, 24.0, 32.0, 48.0, 64.0, 80.0, 96.0, 128.0, 192.0, 512.0, 2000.0}) {
            run_inplace<32, 0>(in, mb, 200 * 1024, "8 warps/SM");
            run_inplace<32, 1>(in, mb, 200 * 1024, "8 warps/SM");
            run_inplace<32, 0>(in, mb, 100 * 1024, "16 warps/SM");
            run_inplace<16, 0>(in, mb, 48 * 1024, "32 warps/SM");
        }
        return 0;
    }
    // occupancy limited to 1 CTA (8 warps) per SM by a 200 KB shared allocation, like agg_seq_kernel
    for (int m = 0; m < 4; ++m) {
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
