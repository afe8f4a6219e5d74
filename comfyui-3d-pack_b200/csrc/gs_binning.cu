// Binning stage: tile-key duplication (in depth order), per-tile ranges, and the
// debug reconstruction of the package's 64-bit keys.
//
// Replaces duplicateWithKeys / identifyTileRanges of diff_gaussian_rasterization
// (SURVEY.md App. A.1.5).  B200-first restructuring: the package sorts K*N
// 64-bit (tile|depth) keys in one go; here the N Gaussians are depth-sorted
// FIRST (N 32-bit keys), pairs are emitted in that order, and only the tile id
// (ceil(log2(tiles)) bits) is sorted over the K*N pairs.  Both sorts are stable,
// so the resulting per-tile order (depth, then Gaussian index) and therefore
// point_list / ranges are bit-identical to the single 64-bit sort, at ~1/3 of
// its memory traffic.
#include "gs_common.cuh"

namespace {

// One thread per depth-sorted Gaussian; writes its run of tiles row-major.
__global__ void __launch_bounds__(256)
emit_kernel(const SplatRec* __restrict__ recs, const uint32_t* __restrict__ sorted_ids,
            const uint32_t* __restrict__ offsets, int N, int tiles_x, int tiles_y,
            uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ vals) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const uint32_t idx = sorted_ids[s];
    const float4 g = __ldg(&recs[idx].g);
    const int radius = __float_as_int(g.w);
    if (radius <= 0) return;
    int x0, y0, x1, y1;
    get_rect(g.x, g.y, radius, tiles_x, tiles_y, x0, y0, x1, y1);
    uint32_t off = offsets[s];
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            tile_keys[off] = (uint32_t)(y * tiles_x + x);
            vals[off] = idx;
            off++;
        }
}

__global__ void __launch_bounds__(256)
ranges_kernel(const uint32_t* __restrict__ keys, int64_t P, uint32_t* __restrict__ ranges) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t t = keys[i];
    if (i == 0) ranges[2 * t] = 0;
    else {
        const uint32_t prev = keys[i - 1];
        if (prev != t) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * t] = (uint32_t)i; }
    }
    if (i == P - 1) ranges[2 * t + 1] = (uint32_t)P;
}

__global__ void __launch_bounds__(256)
sorted_keys_kernel(const SplatRec* __restrict__ recs, const uint32_t* __restrict__ point_list,
                   const uint32_t* __restrict__ tile_keys, int64_t P, uint64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float depth = recs[point_list[i]].g.z;
    out[i] = ((uint64_t)tile_keys[i] << 32) | (uint64_t)__float_as_uint(depth);
}

}  // namespace

int gs_launch_emit(const SplatRec* recs, const uint32_t* sorted_ids, const uint32_t* offsets, int N,
                   int tiles_x, int tiles_y, uint32_t* tile_keys, uint32_t* vals, cudaStream_t s) {
    if (N <= 0) return 0;
    emit_kernel<<<(N + 255) / 256, 256, 0, s>>>(recs, sorted_ids, offsets, N, tiles_x, tiles_y, tile_keys, vals);
    gs_count_launches(1);
    GS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int gs_launch_ranges(const uint32_t* sorted_tile_keys, int64_t P, uint32_t* ranges, cudaStream_t s) {
    if (P <= 0) return 0;
    ranges_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(sorted_tile_keys, P, ranges);
    gs_count_launches(1);
    GS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int gs_launch_sorted_keys(const SplatRec* recs, const uint32_t* point_list, const uint32_t* tile_keys, int64_t P,
                          uint64_t* keys_out, cudaStream_t s) {
    if (P <= 0) return 0;
    sorted_keys_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(recs, point_list, tile_keys, P, keys_out);
    gs_count_launches(1);
    GS_CUDA_CHECK(cudaGetLastError());
    return 0;
}
