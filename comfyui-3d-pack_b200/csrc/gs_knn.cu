// distCUDA2 replacement: mean squared distance to the 3 nearest neighbours.
// Replaces simple_knn._C.distCUDA2 (called once per run from
// main_3DGS_renderer.py:408,419).
//
// v1: exact tiled all-pairs search (every thread keeps its 3 smallest squared
// distances while tiles of 1024 points stream through shared memory).  O(N^2)
// but pure FMA work: ~0.3 s at N = 1e6 on a B200, run once at initialisation.
#include "gs_common.cuh"

namespace {
constexpr int KT = 256;
constexpr int KTILE = 1024;

__global__ void __launch_bounds__(KT) knn3_kernel(const float* __restrict__ pts, int N, float* __restrict__ out) {
    __shared__ float sx[KTILE], sy[KTILE], sz[KTILE];
    const int i = blockIdx.x * KT + threadIdx.x;
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < N) { x = pts[3 * i]; y = pts[3 * i + 1]; z = pts[3 * i + 2]; }
    float b0 = 3.0e38f, b1 = 3.0e38f, b2 = 3.0e38f;   // b0 <= b1 <= b2
    for (int base = 0; base < N; base += KTILE) {
        __syncthreads();
        for (int j = threadIdx.x; j < KTILE; j += KT) {
            const int g = base + j;
            if (g < N) { sx[j] = pts[3 * g]; sy[j] = pts[3 * g + 1]; sz[j] = pts[3 * g + 2]; }
            else { sx[j] = 1.0e18f; sy[j] = 1.0e18f; sz[j] = 1.0e18f; }
        }
        __syncthreads();
        const int self = i - base;
#pragma unroll 8
        for (int j = 0; j < KTILE; j++) {
            const float dx = sx[j] - x, dy = sy[j] - y, dz = sz[j] - z;
            float d = dx * dx + dy * dy + dz * dz;
            if (j == self) d = 3.0e38f;
            if (d < b2) {
                if (d < b1) { b2 = b1; if (d < b0) { b1 = b0; b0 = d; } else b1 = d; }
                else b2 = d;
            }
        }
    }
    if (i < N) {
        // fewer than 3 other points: average what exists (matches an all-pairs definition)
        float s = 0.f; int c = 0;
        if (b0 < 1.0e37f) { s += b0; c++; }
        if (b1 < 1.0e37f) { s += b1; c++; }
        if (b2 < 1.0e37f) { s += b2; c++; }
        out[i] = c ? s / 3.0f : 0.f;
    }
}
}  // namespace

int gs_launch_knn(const float* points, int N, float* out, cudaStream_t s) {
    if (N <= 0) return 0;
    if (!points || !out) { gs_set_error("knn: NULL"); return 1; }
    knn3_kernel<<<(N + KT - 1) / KT, KT, 0, s>>>(points, N, out);
    gs_count_launches(1);
    GS_CUDA_CHECK(cudaGetLastError());
    return 0;
}
