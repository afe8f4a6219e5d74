// C ABI of the rasterizer (include/gs_b200.h): orchestration of the kernels.
// Replaces RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA of
// diff_gaussian_rasterization (reached from main_3DGS_renderer.py:927-936 and
// main_3DGS.py:205).
#include "../../include/gs_b200.h"
#include "gs_common.cuh"

#include <stdarg.h>
#include <string.h>
#include <vector>

static thread_local char g_err[512] = "";

void gs_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#include <atomic>
#include <mutex>
static std::atomic<long long> g_launches{0};
void gs_count_launches(int n) { g_launches += n; }

namespace {

// ---- per-stage CUDA-event timing (off by default; bench.py turns it on for a profiling pass)
struct StageEv { int stage; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<StageEv> g_prof;
struct StageTimer {
    cudaStream_t s; int idx = -1;
    StageTimer(int stage, cudaStream_t s_) : s(s_) {
        if (!g_prof_on) return;
        std::lock_guard<std::mutex> l(g_prof_mu);
        StageEv e; e.stage = stage;
        if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
        cudaEventRecord(e.a, s);
        g_prof.push_back(e); idx = (int)g_prof.size() - 1;
    }
    ~StageTimer() {
        if (idx < 0) return;
        std::lock_guard<std::mutex> l(g_prof_mu);
        if (idx < (int)g_prof.size()) cudaEventRecord(g_prof[idx].b, s);
    }
};

struct Allocator {
    gs_b200_alloc_fn fn;
    void* user;
    cudaStream_t stream;
    std::vector<void*> scratch;     // internally allocated scratch, freed at scope end
    bool failed = false;

    void* get(int tag, size_t bytes, void** owned_slot = nullptr) {
        if (bytes == 0) bytes = 256;
        bytes = (bytes + 255) & ~(size_t)255;
        void* p = nullptr;
        if (fn) {
            p = fn(user, tag, bytes);
            if (!p) { failed = true; gs_set_error("allocator callback returned NULL for %zu bytes (tag %d)", bytes, tag); }
            return p;
        }
        static thread_local int pool_dev = -1;      // keep freed blocks in the pool across stream syncs
        int devn = 0;
        cudaGetDevice(&devn);
        if (pool_dev != devn) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, devn) == cudaSuccess) {
                unsigned long long thr = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
            pool_dev = devn;
        }
        cudaError_t e = cudaMallocAsync(&p, bytes, stream);
        if (e != cudaSuccess) { failed = true; gs_set_error("cudaMallocAsync(%zu) -> %s", bytes, cudaGetErrorString(e)); return nullptr; }
        if (tag == GS_B200_BUF_SCRATCH) scratch.push_back(p);
        else if (owned_slot) *owned_slot = p;
        return p;
    }
    ~Allocator() {
        for (void* p : scratch) cudaFreeAsync(p, stream);
    }
};

// carve sub-buffers out of one allocation
struct Carver {
    char* base; size_t off = 0;
    explicit Carver(void* b) : base((char*)b) {}
    template <typename T> T* take(size_t count) {
        T* p = (T*)(base + off);
        off += (count * sizeof(T) + 255) & ~(size_t)255;
        return p;
    }
    static size_t need(size_t count, size_t elem) { return (count * elem + 255) & ~(size_t)255; }
};

int make_view_args(const gs_b200_view* v, ViewArgs& va) {
    if (!v || !v->bg || !v->viewmatrix || !v->projmatrix || !v->campos) { gs_set_error("view: NULL field"); return 1; }
    if (v->image_height <= 0 || v->image_width <= 0) { gs_set_error("view: bad image size"); return 1; }
    if (v->sh_degree < 0 || v->sh_degree > 3) { gs_set_error("view: sh_degree must be 0..3"); return 1; }
    va.view = v->viewmatrix; va.proj = v->projmatrix; va.campos = v->campos; va.bg = v->bg;
    va.tanfovx = v->tanfovx; va.tanfovy = v->tanfovy;
    va.W = v->image_width; va.H = v->image_height;
    // fp32 focal, as the oracle: S / (2 tan)
    va.focal_x = (float)va.W / (2.0f * va.tanfovx);
    va.focal_y = (float)va.H / (2.0f * va.tanfovy);
    va.scale_modifier = v->scale_modifier;
    va.tiles_x = (va.W + GS_TILE - 1) / GS_TILE;
    va.tiles_y = (va.H + GS_TILE - 1) / GS_TILE;
    va.sh_degree = v->sh_degree;
    return 0;
}

int check_inputs(int N, int M, int sh_degree, const float* means3D, const float* shs, const float* colors_precomp,
                 const float* opacities, const float* scales, const float* rotations, const float* cov3D_precomp) {
    if (N < 0) { gs_set_error("N < 0"); return 1; }
    if (N > 0 && (!means3D || !opacities)) { gs_set_error("means3D/opacities NULL"); return 1; }
    if ((shs == nullptr) == (colors_precomp == nullptr)) {
        gs_set_error("Please provide excatly one of either SHs or precomputed colors!"); return 1;
    }
    const bool sr = scales != nullptr || rotations != nullptr;
    if ((sr && cov3D_precomp) || (!sr && !cov3D_precomp) || (sr && (!scales || !rotations))) {
        gs_set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!"); return 1;
    }
    if (shs && M < (sh_degree + 1) * (sh_degree + 1)) { gs_set_error("shs has %d coefficients, sh_degree %d needs %d", M, sh_degree, (sh_degree + 1) * (sh_degree + 1)); return 1; }
    if (shs && M > 16) { gs_set_error("at most 16 SH coefficients (degree 3) supported, got %d", M); return 1; }
    return 0;
}

int bits_for(int n) { int b = 0; while ((1 << b) < n) b++; return b < 1 ? 1 : b; }

unsigned long long* pinned_u64() {
    static thread_local unsigned long long* p = nullptr;
    if (!p) { if (cudaHostAlloc((void**)&p, 64, cudaHostAllocDefault) != cudaSuccess) p = nullptr; }
    return p;
}

#define STAGE_CHECK(dbg, s, what)                                                                     \
    do {                                                                                              \
        if (dbg) {                                                                                    \
            cudaError_t _e = cudaStreamSynchronize(s);                                                \
            if (_e == cudaSuccess) _e = cudaGetLastError();                                           \
            if (_e != cudaSuccess) { gs_set_error("stage %s failed: %s", what, cudaGetErrorString(_e)); return 1; } \
        }                                                                                             \
    } while (0)

}  // namespace

extern "C" {

int32_t gs_b200_abi_version(void) { return GS_B200_ABI_VERSION; }
int64_t gs_b200_launch_count(void) { return (int64_t)g_launches.load(); }
void gs_b200_profile_enable(int32_t on) {
    std::lock_guard<std::mutex> l(g_prof_mu);
    for (auto& e : g_prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    g_prof.clear();
    g_prof_on = on != 0;
}
int32_t gs_b200_profile_read(float* ms_out, int32_t* calls_out) {
    std::lock_guard<std::mutex> l(g_prof_mu);
    for (int i = 0; i < GS_B200_NSTAGES; i++) { if (ms_out) ms_out[i] = 0.f; if (calls_out) calls_out[i] = 0; }
    for (auto& e : g_prof) {
        if (cudaEventSynchronize(e.b) != cudaSuccess) { gs_set_error("profile_read: event sync failed"); return 1; }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) != cudaSuccess) { gs_set_error("profile_read: elapsed failed"); return 1; }
        if (ms_out) ms_out[e.stage] += ms;
        if (calls_out) calls_out[e.stage] += 1;
    }
    return 0;
}
const char* gs_b200_last_error(void) { return g_err; }

size_t gs_b200_sort_scratch_bytes(int64_t n) { return gs_sort_scratch_bytes(n); }

int32_t gs_b200_sort_pairs_u32(uint32_t* keys, uint32_t* keys_alt, uint32_t* vals, uint32_t* vals_alt, int64_t n,
                               int32_t begin_bit, int32_t end_bit, void* scratch, int32_t* result_in_alt,
                               void* stream) {
    int r = 0;
    int rc = gs_sort_pairs_u32(keys, keys_alt, vals, vals_alt, n, begin_bit, end_bit, scratch, &r, (cudaStream_t)stream);
    if (result_in_alt) *result_in_alt = r;
    return rc;
}

int32_t gs_b200_rasterize_forward(const gs_b200_view* view, int32_t N, int32_t M, const float* means3D,
                                  const float* shs, const float* colors_precomp, const float* opacities,
                                  const float* scales, const float* rotations, const float* cov3D_precomp,
                                  float* out_color, float* out_depth, float* out_alpha, int32_t* radii,
                                  gs_b200_alloc_fn alloc, void* alloc_user, gs_b200_state* state, void* stream_) {
    cudaStream_t s = (cudaStream_t)stream_;
    ViewArgs va;
    if (make_view_args(view, va)) return 1;
    if (check_inputs(N, M, view->sh_degree, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp)) return 1;
    if (!out_color || !out_depth || !out_alpha || (N > 0 && !radii) || !state) { gs_set_error("forward: NULL output"); return 1; }
    const int dbg = view->debug;
    memset(state, 0, sizeof(*state));
    state->num_gaussians = N; state->tiles_x = va.tiles_x; state->tiles_y = va.tiles_y;
    Allocator A{alloc, alloc_user, s};
    const size_t npix = (size_t)va.W * va.H;
    const int ntiles = va.tiles_x * va.tiles_y;

    // ---- image state -----------------------------------------------------------
    const size_t img_bytes = Carver::need((size_t)ntiles * 2, 4) + Carver::need(npix, 4) * 2;
    void* img = A.get(GS_B200_BUF_IMAGE, img_bytes, &state->owned[GS_B200_BUF_IMAGE]);
    if (A.failed) return 1;
    Carver ci(img);
    state->ranges = ci.take<uint32_t>((size_t)ntiles * 2);
    state->n_contrib = ci.take<uint32_t>(npix);
    state->final_T = ci.take<float>(npix);
    GS_CUDA_CHECK(cudaMemsetAsync(state->ranges, 0, (size_t)ntiles * 2 * 4, s));

    // ---- per-Gaussian stage ------------------------------------------------------
    SplatRec* recs = (SplatRec*)A.get(GS_B200_BUF_GEOM, (size_t)N * sizeof(SplatRec), &state->owned[GS_B200_BUF_GEOM]);
    if (A.failed) return 1;
    state->geom = recs;
    unsigned long long P = 0;
    uint32_t *sorted_ids = nullptr, *offsets = nullptr;
    if (N > 0) {
        const size_t sort_b = gs_sort_scratch_bytes(N), scan_b = gs_scan_scratch_bytes(N);
        const size_t sc_bytes = Carver::need(N, 4) * 6 + Carver::need(1, 8) + Carver::need(sort_b, 1) + Carver::need(scan_b, 1);
        void* sc = A.get(GS_B200_BUF_SCRATCH, sc_bytes);
        if (A.failed) return 1;
        Carver c(sc);
        uint32_t* tiles = c.take<uint32_t>(N);
        uint32_t* dkeys = c.take<uint32_t>(N);
        uint32_t* ids = c.take<uint32_t>(N);
        uint32_t* dkeys_alt = c.take<uint32_t>(N);
        uint32_t* ids_alt = c.take<uint32_t>(N);
        offsets = c.take<uint32_t>(N);
        unsigned long long* total = c.take<unsigned long long>(1);
        void* sort_scratch = c.take<char>(sort_b);
        void* scan_scratch = c.take<char>(scan_b);

        { StageTimer t(0, s);
        if (gs_launch_preprocess(va, N, M, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                 recs, radii, tiles, dkeys, ids, s)) return 1; }
        STAGE_CHECK(dbg, s, "preprocess");
        int in_alt = 0;
        { StageTimer t(1, s);
        if (gs_sort_pairs_u32(dkeys, dkeys_alt, ids, ids_alt, N, 0, 32, sort_scratch, &in_alt, s)) return 1; }
        STAGE_CHECK(dbg, s, "depth sort");
        sorted_ids = in_alt ? ids_alt : ids;
        { StageTimer t(2, s);
        if (gs_scan_gather_u32(tiles, sorted_ids, offsets, total, N, scan_scratch, s)) return 1; }
        unsigned long long* hp = pinned_u64();
        if (!hp) { gs_set_error("cudaHostAlloc failed"); return 1; }
        GS_CUDA_CHECK(cudaMemcpyAsync(hp, total, 8, cudaMemcpyDeviceToHost, s));
        GS_CUDA_CHECK(cudaStreamSynchronize(s));
        P = *hp;
    }
    if (P >= (1ull << 30)) { gs_set_error("too many (tile,splat) pairs: %llu", P); return 1; }
    state->num_rendered = (int64_t)P;

    // ---- binning -----------------------------------------------------------------
    const size_t bin_bytes = Carver::need(P, 4) * 2;
    void* bin = A.get(GS_B200_BUF_BINNING, bin_bytes, &state->owned[GS_B200_BUF_BINNING]);
    if (A.failed) return 1;
    Carver cb(bin);
    state->point_list = cb.take<uint32_t>(P);
    state->tile_keys = cb.take<uint32_t>(P);
    if (P > 0) {
        const size_t sort_b = gs_sort_scratch_bytes((int64_t)P);
        void* sc2 = A.get(GS_B200_BUF_SCRATCH, Carver::need(P, 4) * 2 + Carver::need(sort_b, 1));
        if (A.failed) return 1;
        Carver c2(sc2);
        uint32_t* keys_b = c2.take<uint32_t>(P);
        uint32_t* vals_b = c2.take<uint32_t>(P);
        void* sort_scratch = c2.take<char>(sort_b);
        const int tbits = bits_for(ntiles);
        const int npasses = (tbits + 7) / 8;
        // start in the buffer that makes the last pass land in the saved (A) pair
        uint32_t *k0 = (npasses & 1) ? keys_b : state->tile_keys, *v0 = (npasses & 1) ? vals_b : state->point_list;
        uint32_t *k1 = (npasses & 1) ? state->tile_keys : keys_b, *v1 = (npasses & 1) ? state->point_list : vals_b;
        { StageTimer t(3, s);
        if (gs_launch_emit(recs, sorted_ids, offsets, N, va.tiles_x, va.tiles_y, k0, v0, s)) return 1; }
        STAGE_CHECK(dbg, s, "emit");
        int in_alt = 0;
        { StageTimer t(4, s);
        if (gs_sort_pairs_u32(k0, k1, v0, v1, (int64_t)P, 0, tbits, sort_scratch, &in_alt, s)) return 1; }
        STAGE_CHECK(dbg, s, "tile sort");
        { StageTimer t(5, s);
        if (gs_launch_ranges(state->tile_keys, (int64_t)P, state->ranges, s)) return 1; }
        STAGE_CHECK(dbg, s, "ranges");
    }
    { StageTimer t(6, s);
    if (gs_launch_render_forward(va, recs, state->point_list, state->ranges, out_color, out_depth, out_alpha,
                                 state->n_contrib, state->final_T, s)) return 1; }
    STAGE_CHECK(dbg, s, "render");
    return 0;
}

int32_t gs_b200_rasterize_backward(const gs_b200_view* view, int32_t N, int32_t M, const float* means3D,
                                   const float* shs, const float* colors_precomp, const float* opacities,
                                   const float* scales, const float* rotations, const float* cov3D_precomp,
                                   const int32_t* radii, const gs_b200_state* state, const float* dL_dcolor,
                                   const float* dL_ddepth, const float* dL_dalpha, float* dL_dmeans3D,
                                   float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors, float* dL_dopacities,
                                   float* dL_dscales, float* dL_drotations, float* dL_dcov3D, int32_t accumulate,
                                   gs_b200_alloc_fn alloc, void* alloc_user, void* stream_) {
    cudaStream_t s = (cudaStream_t)stream_;
    ViewArgs va;
    if (make_view_args(view, va)) return 1;
    if (check_inputs(N, M, view->sh_degree, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp)) return 1;
    if (!state || !dL_dcolor || !dL_ddepth || !dL_dalpha) { gs_set_error("backward: NULL state/upstream gradient"); return 1; }
    if (N == 0) return 0;
    if (!radii || !dL_dmeans3D || !dL_dmeans2D || !dL_dopacities || (shs && !dL_dshs) || (colors_precomp && !dL_dcolors) ||
        (scales && (!dL_dscales || !dL_drotations)) || (cov3D_precomp && !dL_dcov3D)) {
        gs_set_error("backward: NULL gradient output"); return 1;
    }
    if (state->num_gaussians != N) { gs_set_error("backward: state is for N=%d, got %d", state->num_gaussians, N); return 1; }
    const int dbg = view->debug;
    Allocator A{alloc, alloc_user, s};
    SplatGrad* sg = (SplatGrad*)A.get(GS_B200_BUF_SCRATCH, (size_t)N * sizeof(SplatGrad));
    if (A.failed) return 1;
    GS_CUDA_CHECK(cudaMemsetAsync(sg, 0, (size_t)N * sizeof(SplatGrad), s));
    if (state->num_rendered > 0) {
        { StageTimer t(7, s);
        if (gs_launch_render_backward(va, (const SplatRec*)state->geom, state->point_list, state->ranges,
                                      state->n_contrib, state->final_T, dL_dcolor, dL_ddepth, dL_dalpha, sg, s)) return 1; }
        STAGE_CHECK(dbg, s, "render backward");
    }
    { StageTimer t(8, s);
    if (gs_launch_preprocess_backward(va, N, M, means3D, shs, colors_precomp, opacities, scales, rotations,
                                      cov3D_precomp, radii, sg, dL_dmeans3D, dL_dmeans2D, dL_dshs, dL_dcolors,
                                      dL_dopacities, dL_dscales, dL_drotations, dL_dcov3D, accumulate, s)) return 1; }
    STAGE_CHECK(dbg, s, "preprocess backward");
    return 0;
}

int32_t gs_b200_state_free(gs_b200_state* state, void* stream_) {
    if (!state) return 0;
    for (int i = 0; i < 4; i++)
        if (state->owned[i]) { cudaFreeAsync(state->owned[i], (cudaStream_t)stream_); state->owned[i] = nullptr; }
    return 0;
}

int32_t gs_b200_debug_sorted_keys(const gs_b200_state* state, uint64_t* keys_out, void* stream_) {
    if (!state || !keys_out) { gs_set_error("debug_sorted_keys: NULL"); return 1; }
    return gs_launch_sorted_keys((const SplatRec*)state->geom, state->point_list, state->tile_keys,
                                 state->num_rendered, keys_out, (cudaStream_t)stream_);
}

int32_t gs_b200_knn_mean_dist2(const float* points, int32_t N, float* out, void* stream_) {
    return gs_launch_knn(points, N, out, (cudaStream_t)stream_);
}

// ---------------------------------------------------------------------------------
// Multi-view optimisation step with host buffers (e2e entry).
// ---------------------------------------------------------------------------------
namespace {
struct HostStepCache {
    void* dev = nullptr; size_t bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t up_done[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr}, img_ready[2] = {nullptr, nullptr},
                img_copied[2] = {nullptr, nullptr};
};
thread_local HostStepCache g_hs;
}

static int32_t step_host_impl(int32_t V, int32_t H, int32_t W, int32_t sh_degree, float scale_modifier,
                          const float* views_host, int32_t N, int32_t M, const float* means3D_host,
                          const float* shs_host, const float* opacities_host, const float* scales_host,
                          const float* rotations_host, const float* dL_dout_host, float* grads_host, float* grads_dev,
                          float* images_host, int64_t* num_rendered_out, void* stream_) {
    cudaStream_t s = (cudaStream_t)stream_;
    if (V <= 0 || N <= 0 || !views_host || !means3D_host || !shs_host || !opacities_host || !scales_host ||
        !rotations_host || !dL_dout_host || (!grads_host && !grads_dev)) { gs_set_error("step_host: bad argument"); return 1; }
    HostStepCache& C = g_hs;
    if (!C.copy_stream) {
        GS_CUDA_CHECK(cudaStreamCreateWithFlags(&C.copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            GS_CUDA_CHECK(cudaEventCreateWithFlags(&C.up_done[i], cudaEventDisableTiming));
            GS_CUDA_CHECK(cudaEventCreateWithFlags(&C.consumed[i], cudaEventDisableTiming));
            GS_CUDA_CHECK(cudaEventCreateWithFlags(&C.img_ready[i], cudaEventDisableTiming));
            GS_CUDA_CHECK(cudaEventCreateWithFlags(&C.img_copied[i], cudaEventDisableTiming));
        }
    }
    const size_t npix = (size_t)H * W;
    const size_t n_par = (size_t)N * (3 + 3 * (size_t)M + 1 + 3 + 4);
    const size_t n_grad = n_par + (size_t)N * 3;
    const size_t need = Carver::need(n_par, 4) + Carver::need(n_grad, 4) + Carver::need((size_t)V * 40, 4) +
                        Carver::need(5 * npix, 4) * 4 + Carver::need(N, 4);
    if (C.bytes < need) {
        if (C.dev) cudaFree(C.dev);
        GS_CUDA_CHECK(cudaMalloc(&C.dev, need));
        C.bytes = need;
    }
    Carver c(C.dev);
    float* d_par = c.take<float>(n_par);
    float* d_grad = c.take<float>(n_grad);
    float* d_views = c.take<float>((size_t)V * 40);
    float* d_up[2] = {c.take<float>(5 * npix), c.take<float>(5 * npix)};
    float* d_img[2] = {c.take<float>(5 * npix), c.take<float>(5 * npix)};
    int32_t* d_radii = c.take<int32_t>(N);

    float* d_means = d_par;
    float* d_shs = d_means + (size_t)N * 3;
    float* d_opac = d_shs + (size_t)N * 3 * M;
    float* d_scales = d_opac + N;
    float* d_rots = d_scales + (size_t)N * 3;
    float* g_means = d_grad;
    float* g_shs = g_means + (size_t)N * 3;
    float* g_opac = g_shs + (size_t)N * 3 * M;
    float* g_scales = g_opac + N;
    float* g_rots = g_scales + (size_t)N * 3;
    float* g_m2d = g_rots + (size_t)N * 4;

    GS_CUDA_CHECK(cudaMemcpyAsync(d_means, means3D_host, (size_t)N * 12, cudaMemcpyHostToDevice, s));
    GS_CUDA_CHECK(cudaMemcpyAsync(d_shs, shs_host, (size_t)N * 12 * M, cudaMemcpyHostToDevice, s));
    GS_CUDA_CHECK(cudaMemcpyAsync(d_opac, opacities_host, (size_t)N * 4, cudaMemcpyHostToDevice, s));
    GS_CUDA_CHECK(cudaMemcpyAsync(d_scales, scales_host, (size_t)N * 12, cudaMemcpyHostToDevice, s));
    GS_CUDA_CHECK(cudaMemcpyAsync(d_rots, rotations_host, (size_t)N * 16, cudaMemcpyHostToDevice, s));
    GS_CUDA_CHECK(cudaMemcpyAsync(d_views, views_host, (size_t)V * 160, cudaMemcpyHostToDevice, s));
    GS_CUDA_CHECK(cudaMemsetAsync(d_grad, 0, n_grad * 4, s));

    // upstream-gradient uploads run ahead on the copy stream (double buffered)
    GS_CUDA_CHECK(cudaMemcpyAsync(d_up[0], dL_dout_host, 5 * npix * 4, cudaMemcpyHostToDevice, C.copy_stream));
    GS_CUDA_CHECK(cudaEventRecord(C.up_done[0], C.copy_stream));
    int64_t rendered = 0;
    for (int v = 0; v < V; v++) {
        const int b = v & 1;
        if (v + 1 < V) {
            const int nb = (v + 1) & 1;
            if (v >= 1) GS_CUDA_CHECK(cudaStreamWaitEvent(C.copy_stream, C.consumed[nb], 0));
            GS_CUDA_CHECK(cudaMemcpyAsync(d_up[nb], dL_dout_host + (size_t)(v + 1) * 5 * npix, 5 * npix * 4,
                                          cudaMemcpyHostToDevice, C.copy_stream));
            GS_CUDA_CHECK(cudaEventRecord(C.up_done[nb], C.copy_stream));
        }
        const float* vh = views_host + (size_t)v * 40;
        const float* vd = d_views + (size_t)v * 40;
        gs_b200_view view;
        view.image_height = H; view.image_width = W; view.tanfovx = vh[38]; view.tanfovy = vh[39];
        view.bg = vd + 35; view.scale_modifier = scale_modifier; view.viewmatrix = vd; view.projmatrix = vd + 16;
        view.sh_degree = sh_degree; view.campos = vd + 32; view.prefiltered = 0; view.debug = 0;
        if (images_host && v >= 2) GS_CUDA_CHECK(cudaStreamWaitEvent(s, C.img_copied[b], 0));
        gs_b200_state st;
        float* img = d_img[b];
        if (gs_b200_rasterize_forward(&view, N, M, d_means, d_shs, nullptr, d_opac, d_scales, d_rots, nullptr, img,
                                      img + 3 * npix, img + 4 * npix, d_radii, nullptr, nullptr, &st, s)) return 1;
        rendered += st.num_rendered;
        if (images_host) {
            GS_CUDA_CHECK(cudaEventRecord(C.img_ready[b], s));
            GS_CUDA_CHECK(cudaStreamWaitEvent(C.copy_stream, C.img_ready[b], 0));
            GS_CUDA_CHECK(cudaMemcpyAsync(images_host + (size_t)v * 5 * npix, img, 5 * npix * 4, cudaMemcpyDeviceToHost, C.copy_stream));
            GS_CUDA_CHECK(cudaEventRecord(C.img_copied[b], C.copy_stream));
        }
        GS_CUDA_CHECK(cudaStreamWaitEvent(s, C.up_done[b], 0));
        const float* up = d_up[b];
        int rc = gs_b200_rasterize_backward(&view, N, M, d_means, d_shs, nullptr, d_opac, d_scales, d_rots, nullptr,
                                            d_radii, &st, up, up + 3 * npix, up + 4 * npix, g_means, g_m2d, g_shs,
                                            nullptr, g_opac, g_scales, g_rots, nullptr, 1, nullptr, nullptr, s);
        GS_CUDA_CHECK(cudaEventRecord(C.consumed[b], s));
        gs_b200_state_free(&st, s);
        if (rc) return 1;
    }
    if (grads_dev) GS_CUDA_CHECK(cudaMemcpyAsync(grads_dev, d_grad, n_grad * 4, cudaMemcpyDeviceToDevice, s));
    if (grads_host) GS_CUDA_CHECK(cudaMemcpyAsync(grads_host, d_grad, n_grad * 4, cudaMemcpyDeviceToHost, s));
    GS_CUDA_CHECK(cudaStreamSynchronize(C.copy_stream));
    if (grads_host) GS_CUDA_CHECK(cudaStreamSynchronize(s));
    if (num_rendered_out) *num_rendered_out = rendered;
    return 0;
}

int32_t gs_b200_step_host(int32_t V, int32_t H, int32_t W, int32_t sh_degree, float scale_modifier,
                          const float* views_host, int32_t N, int32_t M, const float* means3D_host,
                          const float* shs_host, const float* opacities_host, const float* scales_host,
                          const float* rotations_host, const float* dL_dout_host, float* grads_host,
                          float* images_host, int64_t* num_rendered_out, void* stream_) {
    return step_host_impl(V, H, W, sh_degree, scale_modifier, views_host, N, M, means3D_host, shs_host,
                          opacities_host, scales_host, rotations_host, dL_dout_host, grads_host, nullptr,
                          images_host, num_rendered_out, stream_);
}

int32_t gs_b200_step_host_dev_grads(int32_t V, int32_t H, int32_t W, int32_t sh_degree, float scale_modifier,
                          const float* views_host, int32_t N, int32_t M, const float* means3D_host,
                          const float* shs_host, const float* opacities_host, const float* scales_host,
                          const float* rotations_host, const float* dL_dout_host, float* grads_dev,
                          float* images_host, int64_t* num_rendered_out, void* stream_) {
    return step_host_impl(V, H, W, sh_degree, scale_modifier, views_host, N, M, means3D_host, shs_host,
                          opacities_host, scales_host, rotations_host, dL_dout_host, nullptr, grads_dev,
                          images_host, num_rendered_out, stream_);
}

}  // extern "C"
