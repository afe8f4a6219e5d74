// Tile compositing: front-to-back alpha blend (forward) and its back-to-front
// replay (backward).
//
// Replaces renderCUDA (forward.cu / backward.cu) of diff_gaussian_rasterization,
// semantics per SURVEY.md App. A.1.6:  alpha = min(0.99, o*exp(power)); skip
// power>0 or alpha<1/255; stop WITHOUT blending when T(1-alpha) < 1e-4;
// color = C + T*bg, depth = sum(depth*alpha*T), alpha_out = sum(alpha*T).
//
// Design (sm_100a):
//  * one CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel sub-rect;
//  * a batch of 256 splat records is gathered with 3 x LDG.128 per record into
//    shared memory; while it is staged, the loading thread computes for its splat
//    an 8-bit mask "can reach alpha>=1/255 inside warp w's sub-rect" from the
//    exact minimum of the quadratic form over the rectangle (conservative by a
//    slack), so each warp only walks the splats that can touch its 32 pixels —
//    the results are unchanged because a skipped splat would have failed the
//    per-pixel alpha test on every lane anyway;
//  * backward: per-splat partial gradients are reduced over the warp's 32
//    pixels with shuffles and leave the SM as one red.global.add per value.
#include "gs_common.cuh"

namespace {

constexpr int RB = 256;   // threads per CTA == splats per batch
constexpr float ALPHA_MIN = 1.0f / 255.0f;

// 8-bit mask over the CTA's warps: bit w set if the splat may contribute in
// sub-rect w (x in [ox+8*(w&1), +7], y in [oy+4*(w>>1), +3]).
__device__ __forceinline__ uint32_t subrect_mask(const float4 g, const float4 c, int ox, int oy) {
    const float o = c.w;
    if (!(o * 255.0f >= 1.0f) || __float_as_int(g.w) <= 0) return 0u;   // can never reach 1/255 (also NaN)
    // contributes iff q(d) = 0.5(A dx^2 + C dy^2) + B dx dy <= tau ;  slack covers fp32 rounding
    const float tau = __logf(o * 255.0f) * 1.001f + 0.02f;
    const float A = c.x, B = c.y, C = c.z;
    if (!(A > 0.f && A * C - B * B > 0.f)) return 0xFFu;   // not positive definite: no culling, per-pixel test decides
    const float invA = 1.0f / A, invC = 1.0f / C;
    uint32_t mask = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const float ux0 = (float)(ox + 8 * (w & 1)) - g.x, ux1 = ux0 + 7.0f;
        const float uy0 = (float)(oy + 4 * (w >> 1)) - g.y, uy1 = uy0 + 3.0f;
        const bool outx = (ux0 > 0.f) || (ux1 < 0.f);
        const bool outy = (uy0 > 0.f) || (uy1 < 0.f);
        float q = 0.f;
        if (outx || outy) {
            q = 3.0e38f;
            if (outx) {
                const float ue = (ux0 > 0.f) ? ux0 : ux1;
                const float uy = fminf(fmaxf(-B * ue * invC, uy0), uy1);
                q = 0.5f * (A * ue * ue + C * uy * uy) + B * ue * uy;
            }
            if (outy) {
                const float ue = (uy0 > 0.f) ? uy0 : uy1;
                const float ux = fminf(fmaxf(-B * ue * invA, ux0), ux1);
                q = fminf(q, 0.5f * (A * ux * ux + C * ue * ue) + B * ux * ue);
            }
        }
        if (!(q > tau)) mask |= 1u << w;    // NaN -> keep (conservative)
    }
    return mask;
}

__global__ void __launch_bounds__(RB)
render_forward_kernel(ViewArgs va, const SplatRec* __restrict__ recs, const uint32_t* __restrict__ point_list,
                      const uint32_t* __restrict__ ranges, float* __restrict__ out_color,
                      float* __restrict__ out_depth, float* __restrict__ out_alpha,
                      uint32_t* __restrict__ n_contrib, float* __restrict__ final_T) {
    __shared__ float4 s_g[RB], s_c[RB], s_k[RB];
    __shared__ uint32_t s_mask[RB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ox = blockIdx.x * GS_TILE, oy = blockIdx.y * GS_TILE;
    const int px = ox + 8 * (warp & 1) + (lane & 7);
    const int py = oy + 4 * (warp >> 1) + (lane >> 3);
    const bool inside = px < va.W && py < va.H;
    const float pxf = (float)px, pyf = (float)py;
    const int tile = blockIdx.y * va.tiles_x + blockIdx.x;
    const uint32_t start = ranges[2 * tile], end = ranges[2 * tile + 1];

    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f, Aacc = 0.f;
    uint32_t last = 0;
    bool done = !inside;

    for (uint32_t base = start; base < end; base += RB) {
        if (__syncthreads_count(done) == RB) break;
        const int n = min((uint32_t)RB, end - base);
        if (tid < n) {
            const uint32_t id = __ldg(point_list + base + tid);
            const float4 g = __ldg(&recs[id].g), c = __ldg(&recs[id].c), k = __ldg(&recs[id].k);
            s_g[tid] = g; s_c[tid] = c; s_k[tid] = k;
            s_mask[tid] = subrect_mask(g, c, ox, oy);
        }
        __syncthreads();
        for (int k0 = 0; k0 < n; k0 += 32) {
            const uint32_t mine = (k0 + lane < n) ? ((s_mask[k0 + lane] >> warp) & 1u) : 0u;
            uint32_t bal = __ballot_sync(0xFFFFFFFFu, mine);
            while (bal) {
                const int j = k0 + __ffs(bal) - 1;
                bal &= bal - 1;
                if (!done) {
                    const float4 g = s_g[j], c = s_c[j];
                    const float dx = g.x - pxf, dy = g.y - pyf;
                    const float power = -0.5f * (c.x * dx * dx + c.z * dy * dy) - c.y * dx * dy;
                    if (power <= 0.f) {
                        const float alpha = fminf(0.99f, c.w * __expf(power));
                        if (alpha >= ALPHA_MIN) {
                            const float test_T = T * (1.f - alpha);
                            if (test_T < 0.0001f) {
                                done = true;
                            } else {
                                const float4 k = s_k[j];
                                const float w = alpha * T;
                                C0 += k.x * w; C1 += k.y * w; C2 += k.z * w;
                                D += g.z * w; Aacc += w;
                                T = test_T;
                                last = base - start + j + 1;
                            }
                        }
                    }
                }
            }
            if (__all_sync(0xFFFFFFFFu, done)) break;
        }
    }
    if (inside) {
        const size_t pix = (size_t)py * va.W + px;
        const size_t plane = (size_t)va.W * va.H;
        const float bg0 = __ldg(va.bg), bg1 = __ldg(va.bg + 1), bg2 = __ldg(va.bg + 2);
        out_color[pix] = C0 + T * bg0;
        out_color[plane + pix] = C1 + T * bg1;
        out_color[2 * plane + pix] = C2 + T * bg2;
        out_depth[pix] = D;
        out_alpha[pix] = Aacc;
        n_contrib[pix] = last;
        final_T[pix] = T;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

__global__ void __launch_bounds__(RB)
render_backward_kernel(ViewArgs va, const SplatRec* __restrict__ recs, const uint32_t* __restrict__ point_list,
                       const uint32_t* __restrict__ ranges, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ final_T, const float* __restrict__ dL_dcolor,
                       const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha,
                       SplatGrad* __restrict__ sg) {
    __shared__ float4 s_g[RB], s_c[RB], s_k[RB];
    __shared__ uint32_t s_mask[RB], s_id[RB];
    __shared__ uint32_t s_max;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ox = blockIdx.x * GS_TILE, oy = blockIdx.y * GS_TILE;
    const int px = ox + 8 * (warp & 1) + (lane & 7);
    const int py = oy + 4 * (warp >> 1) + (lane >> 3);
    const bool inside = px < va.W && py < va.H;
    const float pxf = (float)px, pyf = (float)py;
    const int tile = blockIdx.y * va.tiles_x + blockIdx.x;
    const uint32_t start = ranges[2 * tile], end = ranges[2 * tile + 1];
    if (end <= start) return;

    const size_t pix = (size_t)py * va.W + px;
    const size_t plane = (size_t)va.W * va.H;
    const float T_final = inside ? final_T[pix] : 0.f;
    const uint32_t last_contrib = inside ? n_contrib[pix] : 0u;
    float gC0 = 0.f, gC1 = 0.f, gC2 = 0.f, gD = 0.f, gA = 0.f;
    if (inside) {
        gC0 = dL_dcolor[pix]; gC1 = dL_dcolor[plane + pix]; gC2 = dL_dcolor[2 * plane + pix];
        gD = dL_ddepth[pix]; gA = dL_dalpha[pix];
    }
    const float bg_dot = __ldg(va.bg) * gC0 + __ldg(va.bg + 1) * gC1 + __ldg(va.bg + 2) * gC2;

    if (tid == 0) s_max = 0;
    __syncthreads();
    {
        uint32_t m = last_contrib;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        if (lane == 0 && m) atomicMax(&s_max, m);
    }
    __syncthreads();
    const uint32_t nproc = s_max;        // only positions 1..nproc were blended by some pixel
    if (nproc == 0) return;
    // warp-level bound as well: nothing above this warp's own max was blended by its pixels
    uint32_t wmax = last_contrib;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xFFFFFFFFu, wmax, o));

    float T = T_final;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accD = 0.f, accA = 0.f;
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, lD = 0.f;

    // batches from the back: batch b covers list positions (0-based) [lo, hi)
    for (int hi = (int)nproc; hi > 0; hi -= RB) {
        const int lo = max(0, hi - RB);
        const int n = hi - lo;
        __syncthreads();
        if (tid < n) {
            const uint32_t id = __ldg(point_list + start + lo + tid);
            const float4 g = __ldg(&recs[id].g), c = __ldg(&recs[id].c), k = __ldg(&recs[id].k);
            s_g[tid] = g; s_c[tid] = c; s_k[tid] = k; s_id[tid] = id;
            s_mask[tid] = subrect_mask(g, c, ox, oy);
        }
        __syncthreads();
        for (int k0 = ((n - 1) >> 5) << 5; k0 >= 0; k0 -= 32) {
            if ((uint32_t)(lo + k0) >= wmax) continue;     // whole group is past this warp's last contributor
            const uint32_t mine = (k0 + lane < n) ? ((s_mask[k0 + lane] >> warp) & 1u) : 0u;
            uint32_t bal = __ballot_sync(0xFFFFFFFFu, mine);
            while (bal) {
                const int jb = 31 - __clz(bal);
                bal &= ~(1u << jb);
                const int j = k0 + jb;
                const uint32_t pos1 = (uint32_t)(lo + j) + 1u;       // 1-based list position
                const float4 g = s_g[j], c = s_c[j];
                const float dx = g.x - pxf, dy = g.y - pyf;
                const float power = -0.5f * (c.x * dx * dx + c.z * dy * dy) - c.y * dx * dy;
                const float G = __expf(power);
                const float alpha = fminf(0.99f, c.w * G);
                const bool active = (pos1 <= last_contrib) && (power <= 0.f) && (alpha >= ALPHA_MIN);
                if (!__any_sync(0xFFFFFFFFu, active)) continue;
                float v_mx = 0.f, v_my = 0.f, v_ca = 0.f, v_cb = 0.f, v_cc = 0.f, v_op = 0.f;
                float v_r = 0.f, v_g = 0.f, v_b = 0.f, v_d = 0.f;
                if (active) {
                    const float4 k = s_k[j];
                    T = T / (1.f - alpha);
                    const float dchan = alpha * T;
                    float dL_da = 0.f;
                    acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0; lc0 = k.x;
                    dL_da += (k.x - acc0) * gC0; v_r = dchan * gC0;
                    acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1; lc1 = k.y;
                    dL_da += (k.y - acc1) * gC1; v_g = dchan * gC1;
                    acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2; lc2 = k.z;
                    dL_da += (k.z - acc2) * gC2; v_b = dchan * gC2;
                    accD = last_alpha * lD + (1.f - last_alpha) * accD; lD = g.z;
                    dL_da += (g.z - accD) * gD; v_d = dchan * gD;
                    accA = last_alpha + (1.f - last_alpha) * accA;
                    dL_da += (1.f - accA) * gA;
                    dL_da *= T;
                    last_alpha = alpha;
                    dL_da += (-T_final / (1.f - alpha)) * bg_dot;
                    // straight-through min(0.99, .): gradient as if unclamped (App. A.1.6)
                    const float dL_dG = c.w * dL_da;
                    const float gdx = G * dx, gdy = G * dy;
                    v_mx = dL_dG * (-gdx * c.x - gdy * c.y);
                    v_my = dL_dG * (-gdy * c.z - gdx * c.y);
                    v_ca = -0.5f * gdx * dx * dL_dG;
                    v_cb = -gdx * dy * dL_dG;
                    v_cc = -0.5f * gdy * dy * dL_dG;
                    v_op = G * dL_da;
                }
                v_mx = warp_sum(v_mx); v_my = warp_sum(v_my); v_d = warp_sum(v_d);
                v_ca = warp_sum(v_ca); v_cb = warp_sum(v_cb); v_cc = warp_sum(v_cc); v_op = warp_sum(v_op);
                v_r = warp_sum(v_r); v_g = warp_sum(v_g); v_b = warp_sum(v_b);
                if (lane == 0) {
                    float* dst = reinterpret_cast<float*>(sg + s_id[j]);
                    atomicAdd(dst + 0, v_mx); atomicAdd(dst + 1, v_my); atomicAdd(dst + 2, v_d);
                    atomicAdd(dst + 4, v_ca); atomicAdd(dst + 5, v_cb); atomicAdd(dst + 6, v_cc); atomicAdd(dst + 7, v_op);
                    atomicAdd(dst + 8, v_r); atomicAdd(dst + 9, v_g); atomicAdd(dst + 10, v_b);
                }
            }
        }
    }
}

}  // namespace

int gs_launch_render_forward(const ViewArgs& va, const SplatRec* recs, const uint32_t* point_list,
                             const uint32_t* ranges, float* out_color, float* out_depth, float* out_alpha,
                             uint32_t* n_contrib, float* final_T, cudaStream_t s) {
    dim3 grid(va.tiles_x, va.tiles_y);
    render_forward_kernel<<<grid, RB, 0, s>>>(va, recs, point_list, ranges, out_color, out_depth, out_alpha,
                                              n_contrib, final_T);
    gs_count_launches(1);
    GS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int gs_launch_render_backward(const ViewArgs& va, const SplatRec* recs, const uint32_t* point_list,
                              const uint32_t* ranges, const uint32_t* n_contrib, const float* final_T,
                              const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                              SplatGrad* sg, cudaStream_t s) {
    dim3 grid(va.tiles_x, va.tiles_y);
    render_backward_kernel<<<grid, RB, 0, s>>>(va, recs, point_list, ranges, n_contrib, final_T, dL_dcolor,
                                               dL_ddepth, dL_dalpha, sg);
    gs_count_launches(1);
    GS_CUDA_CHECK(cudaGetLastError());
    return 0;
}
