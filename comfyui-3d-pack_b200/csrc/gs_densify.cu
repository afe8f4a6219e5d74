// Densification as stream compaction in the packed layout.
//
// Replaces the Python surgery of GaussianModel.densify_and_prune / densify_and_clone / densify_and_split /
// prune_points / cat_tensors_to_optimizer (MVs_Algorithms/GaussianSplatting/main_3DGS_renderer.py:543-688,752-781):
// the reference builds boolean masks, rebuilds every nn.Parameter and patches the Adam state dicts.  Here
//   plan : one pass classifies every Gaussian (clone / split / keep, and whether its children survive the prune), four
//          exclusive scans turn the flags into destination rows; the host reads the four counts (the only sync, 32 B);
//   apply: a destination map (source row + kind per new row) is scattered once, then every group of the parameters
//          AND of both Adam moments is gathered through it into buffers laid out for the new N, fully coalesced on the
//          writing side — survivors in order, then the kept clones, then the kept split children (first copies, then
//          second copies: the order torch.cat / repeat(2, 1) / mask produce in the reference) — new rows with zero
//          moments (cat_tensors_to_optimizer, :606-625); a last small kernel samples the children.
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

// Destination map: for every row of the new packing its source row and what it is (bit 31: new row = zero Adam
// moments; bits 30..29: 0 survivor / clone, 1 first child, 2 second child of a split parent).
constexpr uint32_t MAP_NEW = 0x80000000u, MAP_CHILD_SHIFT = 29, MAP_ROW_MASK = 0x1FFFFFFFu;

__global__ void __launch_bounds__(DN_THREADS)
densify_map_kernel(int N, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ offs, int n_keep, int n_clone,
                   int n_split_keep, uint32_t* __restrict__ dst_map, uint32_t* __restrict__ child_rank) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (flags[i]) dst_map[offs[i]] = (uint32_t)i;
    if (flags[(size_t)N + i]) dst_map[(size_t)n_keep + offs[(size_t)N + i]] = (uint32_t)i | MAP_NEW;
    if (flags[3 * (size_t)N + i]) {
        const size_t r = offs[3 * (size_t)N + i];
        const size_t d0 = (size_t)n_keep + n_clone + r, d1 = d0 + n_split_keep;
        dst_map[d0] = (uint32_t)i | MAP_NEW | (1u << MAP_CHILD_SHIFT);
        dst_map[d1] = (uint32_t)i | MAP_NEW | (2u << MAP_CHILD_SHIFT);
        child_rank[r] = offs[2 * (size_t)N + i];            // rank of the parent among ALL split parents: row of its z samples
    }
}

// one thread per float of the destination group: coalesced writes, reads in long ascending runs
__global__ void __launch_bounds__(DN_THREADS)
densify_gather_kernel(size_t total, int width, const uint32_t* __restrict__ dst_map, const float* __restrict__ src,
                      float* __restrict__ dst, int zero_new) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const size_t d = e / (size_t)width;
    const int k = (int)(e - d * (size_t)width);
    const uint32_t m = dst_map[d];
    dst[e] = (zero_new && (m & MAP_NEW)) ? 0.f : src[(size_t)(m & MAP_ROW_MASK) * width + k];
}

// split children: xyz = R(q/|q|) (exp(scaling) * z) + xyz ; scaling = log(exp(scaling) / 1.6)   (densify_and_split, :641-668)
__global__ void __launch_bounds__(DN_THREADS)
densify_children_kernel(int n_children, int n_split_keep, int n_split_parents, size_t first_row, const uint32_t* __restrict__ dst_map,
                        const uint32_t* __restrict__ child_rank, const float* __restrict__ src_xyz, const float* __restrict__ src_scal,
                        const float* __restrict__ src_rot, const float* __restrict__ z, float* __restrict__ dst_xyz,
                        float* __restrict__ dst_scal) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_children) return;
    const size_t d = first_row + j;
    const uint32_t m = dst_map[d];
    const size_t i = m & MAP_ROW_MASK;
    const int c = (int)((m >> MAP_CHILD_SHIFT) & 3u) - 1;
    const int r = j - c * n_split_keep;
    float qr = src_rot[4 * i], qx = src_rot[4 * i + 1], qy = src_rot[4 * i + 2], qz = src_rot[4 * i + 3];
    const float inv = 1.0f / sqrtf(qr * qr + qx * qx + qy * qy + qz * qz);       // build_rotation normalises (:81-102)
    qr *= inv; qx *= inv; qy *= inv; qz *= inv;
    const float s0 = expf(src_scal[3 * i]), s1 = expf(src_scal[3 * i + 1]), s2 = expf(src_scal[3 * i + 2]);
    const float* zz = z + 3 * ((size_t)c * n_split_parents + child_rank[r]);
    const float e0 = s0 * zz[0], e1 = s1 * zz[1], e2 = s2 * zz[2];               // torch.normal(0, std) = std * z
    dst_xyz[3 * d + 0] = (1.f - 2.f * (qy * qy + qz * qz)) * e0 + (2.f * (qx * qy - qr * qz)) * e1 + (2.f * (qx * qz + qr * qy)) * e2 + src_xyz[3 * i + 0];
    dst_xyz[3 * d + 1] = (2.f * (qx * qy + qr * qz)) * e0 + (1.f - 2.f * (qx * qx + qz * qz)) * e1 + (2.f * (qy * qz - qr * qx)) * e2 + src_xyz[3 * i + 1];
    dst_xyz[3 * d + 2] = (2.f * (qx * qz - qr * qy)) * e0 + (2.f * (qy * qz + qr * qx)) * e1 + (1.f - 2.f * (qx * qx + qy * qy)) * e2 + src_xyz[3 * i + 2];
    dst_scal[3 * d + 0] = logf(s0 / (0.8f * 2.f)); dst_scal[3 * d + 1] = logf(s1 / (0.8f * 2.f)); dst_scal[3 * d + 2] = logf(s2 / (0.8f * 2.f));
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
    const size_t nnew = (size_t)n_keep + n_clone + 2 * (size_t)n_split_keep;
    if (nnew == 0) return 0;
    if (nnew > MAP_ROW_MASK || (size_t)N > MAP_ROW_MASK) { gs_set_error("densify_apply: more than 2^29 rows"); return 1; }
    // the destination map and the child ranks live in the flag / offset workspace's tail?  No: they must not alias the
    // flags and offsets the map kernel reads -> a stream-ordered scratch allocation
    uint32_t* dst_map = nullptr;
    GS_CUDA_CHECK(cudaMallocAsync((void**)&dst_map, (nnew + (size_t)std::max(n_split_keep, 1)) * sizeof(uint32_t), s));
    uint32_t* child_rank = dst_map + nnew;
    densify_map_kernel<<<(N + DN_THREADS - 1) / DN_THREADS, DN_THREADS, 0, s>>>(N, work, work + 4 * (size_t)N, n_keep, n_clone, n_split_keep,
                                                                               dst_map, child_rank);
    const Groups gs = groups_of((size_t)N, (size_t)M), gd = groups_of(nnew, (size_t)M);
    const size_t so[5] = {gs.xyz, gs.shs, gs.opac, gs.scal, gs.rot}, dofs[5] = {gd.xyz, gd.shs, gd.opac, gd.scal, gd.rot};
    const int width[5] = {3, 3 * M, 1, 3, 4};
    const float* srcs[3] = {src_raw, src_m1, src_m2};
    float* dsts[3] = {dst_raw, dst_m1, dst_m2};
    int launches = 1;
    for (int b = 0; b < 3; b++)
        for (int g = 0; g < 5; g++) {
            const size_t total = nnew * (size_t)width[g];
            densify_gather_kernel<<<(unsigned)((total + DN_THREADS - 1) / DN_THREADS), DN_THREADS, 0, s>>>(
                total, width[g], dst_map, srcs[b] + so[g], dsts[b] + dofs[g], b > 0 ? 1 : 0);
            launches++;
        }
    if (n_split_keep > 0) {
        const int nc = 2 * n_split_keep;
        densify_children_kernel<<<(nc + DN_THREADS - 1) / DN_THREADS, DN_THREADS, 0, s>>>(
            nc, n_split_keep, n_split_parents, (size_t)n_keep + n_clone, dst_map, child_rank, src_raw + gs.xyz, src_raw + gs.scal,
            src_raw + gs.rot, z, dst_raw + gd.xyz, dst_raw + gd.scal);
        launches++;
    }
    gs_count_launches(launches);
    GS_CUDA_CHECK(cudaGetLastError());
    GS_CUDA_CHECK(cudaFreeAsync(dst_map, s));
    return 0;
}

}  // extern "C"
