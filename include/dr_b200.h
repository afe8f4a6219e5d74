/*
 * dr_b200.h — C ABI of the B200-native mesh rasterisation ops (the nvdiffrast surface ComfyUI-3D-Pack uses).
 *
 * Replaces, for the reference's call sites, the torch ops of the un-vendored package `nvdiffrast.torch`
 * (0.3.3, my-reqs.txt:74):
 *   rasterize   MVs_Algorithms/DiffRastMesh/diff_mesh_renderer.py:97, FlexiCubes/flexicubes_renderer.py:49,
 *               mesh_processer/mesh_utils.py:531
 *   interpolate diff_mesh_renderer.py:104,110,131; FlexiCubes/util.py:90-93; mesh_utils.py:534
 *   texture     diff_mesh_renderer.py:105 (filter_mode='linear')
 *   antialias   diff_mesh_renderer.py:101,138; flexicubes_renderer.py:55
 * Plain device pointers + sizes, fp32 / int32, row-major contiguous; every function returns 0 or non-zero with
 * the message in gs_b200_last_error() (shared with gs_b200.h); work is enqueued on `stream`.
 * Layouts: pos[B,V,4] clip space; tri[F,3]; rast[B,H,W,4] = (u, v, z/w, id+1), row 0 = bottom row;
 * rast_db[B,H,W,4] = (du/dX, du/dY, dv/dX, dv/dY) per pixel.
 */
#ifndef DR_B200_H
#define DR_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* scratch for rasterize: B*H*W 64-bit depth/id words + a queue for large triangles */
size_t dr_b200_rasterize_scratch_bytes(int32_t B, int32_t F, int32_t H, int32_t W);
int32_t dr_b200_rasterize_fwd(const float* pos, const int32_t* tri, int32_t B, int32_t V, int32_t F, int32_t H,
                              int32_t W, float* rast, float* rast_db, void* scratch, void* stream);
/* dL_dpos[B,V,4] is ADDED to (caller zeroes it); gradients flow through u,v only */
int32_t dr_b200_rasterize_bwd(const float* pos, const int32_t* tri, int32_t B, int32_t V, int32_t F, int32_t H,
                              int32_t W, const float* rast, const float* dL_drast, float* dL_dpos, void* stream);

/* attr[attr_B,V,A] with attr_B in {1,B}; rast_db/out_da may be NULL (diff_attrs=None) else all attributes are
 * differentiated: out_da[B,H,W,2A] = (da/dX, da/dY) per attribute */
int32_t dr_b200_interpolate_fwd(const float* attr, int32_t attr_B, const float* rast, const int32_t* tri,
                                const float* rast_db, int32_t B, int32_t V, int32_t F, int32_t H, int32_t W,
                                int32_t A, float* out, float* out_da, void* stream);
/* dL_dattr is ADDED to (caller zeroes); dL_drast[B,H,W,4] and dL_drast_db[B,H,W,4] (may be NULL) are written */
int32_t dr_b200_interpolate_bwd(const float* attr, int32_t attr_B, const float* rast, const int32_t* tri,
                                const float* rast_db, int32_t B, int32_t V, int32_t F, int32_t H, int32_t W,
                                int32_t A, const float* dL_dout, const float* dL_dout_da, float* dL_dattr,
                                float* dL_drast, float* dL_drast_db, void* stream);

/* bilinear texture lookup; tex[tex_B,Ht,Wt,C] (tex_B in {1,B}); uv[B,H,W,2]; boundary 0 = wrap, 1 = clamp */
int32_t dr_b200_texture_fwd(const float* tex, int32_t tex_B, int32_t Ht, int32_t Wt, int32_t C, const float* uv,
                            int32_t B, int32_t H, int32_t W, int32_t boundary, float* out, void* stream);
/* dL_dtex is ADDED to (caller zeroes); dL_duv written */
int32_t dr_b200_texture_bwd(const float* tex, int32_t tex_B, int32_t Ht, int32_t Wt, int32_t C, const float* uv,
                            int32_t B, int32_t H, int32_t W, int32_t boundary, const float* dL_dout,
                            float* dL_dtex, float* dL_duv, void* stream);

/* topology for antialias: opp[F,3] = opposite vertex of the other triangle across edge e (edge e joins vertices
 * e+1,e+2 of the triangle), -1 if none.  scratch: dr_b200_topology_scratch_bytes(F) */
size_t dr_b200_topology_scratch_bytes(int32_t F);
int32_t dr_b200_edge_opposites(const int32_t* tri, int32_t F, int32_t V, int32_t* opp, void* scratch, void* stream);

int32_t dr_b200_antialias_fwd(const float* color, const float* rast, const float* pos, const int32_t* tri,
                              const int32_t* opp, int32_t B, int32_t V, int32_t F, int32_t H, int32_t W, int32_t C,
                              float* out, void* stream);
/* dL_dcolor written; dL_dpos[B,V,4] ADDED to (caller zeroes) */
int32_t dr_b200_antialias_bwd(const float* color, const float* rast, const float* pos, const int32_t* tri,
                              const int32_t* opp, int32_t B, int32_t V, int32_t F, int32_t H, int32_t W, int32_t C,
                              const float* dL_dout, float* dL_dcolor, float* dL_dpos, void* stream);
#ifdef __cplusplus
}
#endif
#endif
