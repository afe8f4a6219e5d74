// Densification as stream compaction in the packed layout.
//
// Replaces the Python surgery of GaussianModel.densify_and_prune / densify_and_clone / densify_and_split /
// prune_points / cat_tensors_to_optimizer (MVs_Algorithms/GaussianSplatting/main_3DGS_renderer.py:543-688,752-781):
// the reference builds boolean masks, rebuilds every nn.Parameter and patches the Adam state dicts.  Here
//   plan : one pass classifies every Gaussian (clone / split / keep, and whether its children survive the prune), four
//          exclusive scans turn the flags into destination rows; the host reads the four counts (the only sync, 32 B);
//   apply: one pass scatters parameters AND both Adam moments from the source buffers into destination buffers laid
//          out for the new N — survivors in order, then the kept clones, then the kept split children (first copies,
//          then second copies: the order torch.cat / repeat(2, 1) / mask produce in the reference) — new rows with zero
//          moments (cat_tensors_to_optimizer, :606-625).
// Order of the rules (reproduced, pinned by tests/golden/ref_training.npz):  grads = accum / denom, NaN -> 0;
// clone = grads >= thr and max(exp(scaling)) <= percent_dense * extent; split = the same with >;  split uses
// samples = N(0, exp(scaling)) rotated by the normalised quaternion, new scaling = log(exp(scaling) / (0.8 * 2));
// prune = split parents | sigmoid(opacity) < min_opacity | max(exp(scaling)) > 0.1 * extent, evaluated on the old AND
// the new points (the screen-size rule is inert in the reference: max_radii2D is zeroed before it is read).
#include "../../include/gs_b200.h"
#include "gs_common.cuh"
#include <algorithm>

namespace {

constexpr int DN_THREADS = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// flags[0][i] keep_old, flags[1][i] clone kept, flags[2][i] split parent, flags[3][i] split children kept
__global__ void __launch_bounds__(DN_THREADS)
densify_classify_kernel(int N, const float* __restrict__ raw_opacity, const float* __restrict__ raw_scaling,
                        const float* __restrict__ grad_accum, const float* __restrict__ denom, float max_grad,
                        float min_opacity, float extent, float percent_dense, uint32_t* __restrict__ flags,
                        unsigned long long* __restrict__ clone_total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool clone = false;
    if (i < N) {
        float g = grad_accum[i] / denom[i];
        if (g != g) g = 0.f;
        const float s0 = expf(raw_scaling[3 * (size_t)i]), s1 = expf(raw_scaling[3 * (size_t)i + 1]), s2 = expf(raw_scaling[3 * (size_t)i + 2]);
        const float scal = fmaxf(s0, fmaxf(s1, s2));
        const bool big = scal > percent_dense * extent;
        const bool sel = g >= max_grad;
        clone = sel && !big;
        const bool split = sel && big;
        const float op = sigmoidf_(raw_opacity[i]);
        const bool bad = (op < min_opacity) || (scal > 0.1f * extent);
        // children: scaling <- log(exp(scaling) / 1.6); the prune reads exp() of that
        const float c0 = expf(logf(s0 / (0.8f * 2.f))), c1 = expf(logf(s1 / (0.8f * 2.f))), c2 = expf(logf(s2 / (0.8f * 2.f)));
        const bool child_bad = (op < min_opacity) || (fmaxf(c0, fmaxf(c1, c2)) > 0.1f * extent);
        flags[i] = !(split || bad);
        flags[(size_t)N + i] = clone && !bad;
        flags[2 * (size_t)N + i] = split;
        flags[3 * (size_t)N + i] = split && !child_bad;
    }
    const unsigned n = __syncthreads_count(clone);
    if (threadIdx.x == 0 && n) atomicAdd(clone_total, (unsigned long long)n);
}

struct Groups { size_t xyz, shs, opac, scal, rot; };     // float offsets of the groups inside a packed buffer of n rows
__host__ __device__ inline Groups groups_of(size_t n, size_t M) {
    Groups g;
    g.xyz = 0; g.shs = 3 * n; g.opac = g.shs + 3 * M * n; g.scal = g.opac + n; g.rot = g.scal + 3 * n;
    return g;
}

// one warp per source Gaussian: lanes copy the SH row (3M floats), lanes 0..10 the 11 small-group floats
__device__ __forceinline__ void copy_row(const float* __restrict__ src, float* __restrict__ dst, const Groups& gs, const Groups& gd,
                                         size_t i, size_t d, int M, int lane, bool zero) {
    const int row = 3 * M;
    for (int k = lane; k < row; k += 32) dst[gd.shs + d * row + k] = zero ? 0.f : src[gs.shs + i * row + k];
    if (lane < 3) dst[gd.xyz + 3 * d + lane] = zero ? 0.f : src[gs.xyz + 3 * i + lane];
    else if (lane == 3) dst[gd.opac + d] = zero ? 0.f : src[gs.opac + i];
    else if (lane < 7) dst[gd.scal + 3 * d + (lane - 4)] = zero ? 0.f : src[gs.scal + 3 * i + (lane - 4)];
    else if (lane < 11) dst[gd.rot + 4 * d + (lane - 7)] = zero ? 0.f : src[gs.rot + 4 * i + (lane - 7)];
}

__global__ void __launch_bounds__(DN_THREADS)
densify_apply_kernel(int N, int M, const float* __restrict__ src_raw, const float* __restrict__ src_m1,
                     const float* __restrict__ src_m2, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ offs,
                     int n_keep, int n_clone, int n_split_parents, int n_split_keep, const float* __restrict__ z,
                     float* __restrict__ dst_raw, float* __restrict__ dst_m1, float* __restrict__ dst_m2) {
    const int lane = threadIdx.x & 31;
    const size_t nnew = (size_t)n_keep + n_clone + 2 * (size_t)n_split_keep;
    const Groups gs = groups_of((size_t)N, (size_t)M), gd = groups_of(nnew, (size_t)M);
    for (int i = blockIdx.x * (DN_THREADS / 32) + (threadIdx.x >> 5); i < N; i += gridDim.x * (DN_THREADS / 32)) {
        const uint32_t f0 = flags[i], f1 = flags[(size_t)N + i], f3 = flags[3 * (size_t)N + i];
        if (f0) {                                   // survivor: parameters and moments move together
            const size_t d = offs[i];
            copy_row(src_raw, dst_raw, gs, gd, i, d, M, lane, false);
            copy_row(src_m1, dst_m1, gs, gd, i, d, M, lane, false);
            copy_row(src_m2, dst_m2, gs, gd, i, d, M, lane, false);
        }
        if (f1) {                                   // clone: same parameters, zero Adam state
            const size_t d = (size_t)n_keep + offs[(size_t)N + i];
            copy_row(src_raw, dst_raw, gs, gd, i, d, M, lane, false);
            copy_row(src_m1, dst_m1, gs, gd, i, d, M, lane, true);
            copy_row(src_m2, dst_m2, gs, gd, i, d, M, lane, true);
        }
        if (f3) {                                   // two children sampled inside the parent
            const size_t pr = offs[2 * (size_t)N + i], r = offs[3 * (size_t)N + i];
            const float* q = src_raw + gs.rot + 4 * (size_t)i;
            float qr = q[0], qx = q[1], qy = q[2], qz = q[3];
            const float inv = 1.0f / sqrtf(qr * qr + qx * qx + qy * qy + qz * qz);       // build_rotation normalises (:81-102)
            qr *= inv; qx *= inv; qy *= inv; qz *= inv;
            const float R[9] = {1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - qr * qz), 2.f * (qx * qz + qr * qy),
                                2.f * (qx * qy + qr * qz), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - qr * qx),
                                2.f * (qx * qz - qr * qy), 2.f * (qy * qz + qr * qx), 1.f - 2.f * (qx * qx + qy * qy)};
            const float s0 = expf(src_raw[gs.scal + 3 * (size_t)i]), s1 = expf(src_raw[gs.scal + 3 * (size_t)i + 1]),
                        s2 = expf(src_raw[gs.scal + 3 * (size_t)i + 2]);
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const size_t d = (size_t)n_keep + n_clone + (size_t)c * n_split_keep + r;
                copy_row(src_raw, dst_raw, gs, gd, i, d, M, lane, false);      // shs, opacity, rotation stay; xyz / scaling overwritten below
                copy_row(src_m1, dst_m1, gs, gd, i, d, M, lane, true);
                copy_row(src_m2, dst_m2, gs, gd, i, d, M, lane, true);
                __syncwarp();
                const float* zz = z + 3 * ((size_t)c * n_split_parents + pr);
                const float e0 = s0 * zz[0], e1 = s1 * zz[1], e2 = s2 * zz[2];         // torch.normal(0, std) = std * z
                if (lane < 3) {
                    const float o0 = R[0] * e0 + R[1] * e1 + R[2] * e2, o1 = R[3] * e0 + R[4] * e1 + R[5] * e2,
                                o2 = R[6] * e0 + R[7] * e1 + R[8] * e2;                                        // bmm(R, samples)
                    const float off = lane == 0 ? o0 : (lane == 1 ? o1 : o2);
                    dst_raw[gd.xyz + 3 * d + lane] = off + src_raw[gs.xyz + 3 * (size_t)i + lane];
                    const float sl = lane == 0 ? s0 : (lane == 1 ? s1 : s2);
                    dst_raw[gd.scal + 3 * d + lane] = logf(sl / (0.8f * 2.f));
                }
            }
        }
    }
}

}  // namespace

extern "C" {

size_t gs_b200_densify_scratch_bytes(int32_t N) { return gs_scan_scratch_bytes(N) + 256; }

int32_t gs_b200_densify_plan(int32_t N, const float* raw_opacity, const float* raw_scaling, const float* grad_accum,
                             const float* denom, float max_grad, float min_opacity, float extent, float percent_dense,
                             uint32_t* work, uint64_t* counts_dev, void* scratch, void* stream_) {
    cudaStream_t s = (cudaStream_t)stream_;
    if (N <= 0 || !raw_opacity || !raw_scaling || !grad_accum || !denom || !work || !counts_dev || !scratch) {
        gs_set_error("densify_plan: bad argument"); return 1; }
    uint32_t* flags = work;
    uint32_t* offs = work + 4 * (size_t)N;
    unsigned long long* counts = (unsigned long long*)counts_dev;
    GS_CUDA_CHECK(cudaMemsetAsync(counts, 0, 5 * sizeof(unsigned long long), s));
    densify_classify_kernel<<<(N + DN_THREADS - 1) / DN_THREADS, DN_THREADS, 0, s>>>(N, raw_opacity, raw_scaling, grad_accum, denom,
                                                                                    max_grad, min_opacity, extent, percent_dense, flags, counts + 4);
    gs_count_launches(1);
    GS_CUDA_CHECK(cudaGetLastError());
    for (int k = 0; k < 4; k++)
        if (gs_scan_gather_u32(flags + (size_t)k * N, nullptr, offs + (size_t)k * N, counts + k, N, scratch, s)) return 1;
    gs_count_launches(12);
    return 0;
}

int32_t gs_b200_densify_apply(int32_t N, int32_t M, const float* src_raw, const float* src_m1, const float* src_m2,
                              const uint32_t* work, int32_t n_keep, int32_t n_clone, int32_t n_split_parents,
                              int32_t n_split_keep, const float* z, float* dst_raw, float* dst_m1, float* dst_m2, void* stream_) {
    cudaStream_t s = (cudaStream_t)stream_;
    if (N <= 0 || M <= 0 || !src_raw || !src_m1 || !src_m2 || !work || !dst_raw || !dst_m1 || !dst_m2 || (n_split_keep > 0 && !z) ||
        n_keep < 0 || n_clone < 0 || n_split_parents < n_split_keep || n_split_keep < 0) { gs_set_error("densify_apply: bad argument"); return 1; }
    const int warps_per_block = DN_THREADS / 32;
    const int blocks = std::min((N + warps_per_block - 1) / warps_per_block, 148 * 32);
    densify_apply_kernel<<<blocks, DN_THREADS, 0, s>>>(N, M, src_raw, src_m1, src_m2, work, work + 4 * (size_t)N, n_keep, n_clone,
                                                       n_split_parents, n_split_keep, z, dst_raw, dst_m1, dst_m2);
    gs_count_launches(1);
    GS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
