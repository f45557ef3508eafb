/*
 * b2a.h - C ABI of libb2a.so: the B200-native (sm_100a) reconstruction hot path of 3DAnimals.
 *
 * Drop-in boundary.  Every entry point takes plain device pointers + sizes + a cudaStream_t and
 * returns 0 on success / non-zero on error (message: b2a_last_error_string(), thread-local).  The library never
 * allocates, frees or retains device memory; the caller owns inputs, outputs and workspaces and keeps them alive
 * until the stream work completes.  No host synchronisation happens inside any call.  All floating point is fp32,
 * contiguous row-major; index buffers are int32 unless a parameter says i64.
 *
 * Each group cites the reference interface it replaces (paths relative to the reference repo root).
 */
#ifndef B2A_H
#define B2A_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b2a_stream_t; /* cudaStream_t */

#define B2A_VERSION 100

int b2a_version(void);
const char* b2a_last_error_string(void);

/* ------------------------------------------------------------------------------------------------------------
 * Marching tetrahedra.  Replaces DMTet.__call__ (model/geometry/dmtet.py:104-155) and its autograd backward.
 * Static per-grid tables (built once at load_tets time, dmtet.py:214-226 / generate_edges :283-288):
 *   edge_start [Vg+1], edge_b [E]  = CSR of the unique sorted (min,max) grid edges, lexicographic.
 * Two phases so the caller can size exact outputs with ONE device->host read of `counts`:
 *   count: counts[0]=V (crossing edges), counts[1]=N1 (tets with 1 triangle), counts[2]=N2 (tets with 2).
 *   emit : verts [V,3], vert_edge [V,2] (grid-vertex pair of each output vertex, for the backward),
 *          faces ordered [1-triangle tets in tet order][2-triangle tets in tet order] (dmtet.py:140-143),
 *          uv_idx = (4t, 4t+k+1, 4t+k+2) for tet t / k-th triangle (map_uv :69-98).  Any of the face outputs
 *          may be NULL.
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_mt_workspace_bytes(int64_t Vg, int64_t E, int64_t T, size_t* bytes);
/* tile_words (nullable): static per-grid table [ceil(T/tile_tets), tile_words] int32 of the occupancy words (vertex >> 5)
 * touched by each tile of tile_tets consecutive tets, padded by repeating an entry; -1 in slot 0 = "too many, never skip".
 * Tiles whose words are uniformly inside or outside are not read at all.  Shape constants: b2a_mt_tile_shape. */
int b2a_mt_tile_shape(int* tile_tets, int* tile_words);
int b2a_mt_count(const float* sdf, const int32_t* tets, const int32_t* edge_start, const int32_t* edge_b,
                 const int32_t* tile_words, int64_t Vg, int64_t E, int64_t T, void* workspace, size_t workspace_bytes,
                 int32_t* counts /* device [4] */, b2a_stream_t stream);
int b2a_mt_emit(const float* pos, const float* sdf, const int32_t* tets, const int32_t* edge_start,
                const int32_t* edge_b, int64_t Vg, int64_t E, int64_t T, void* workspace, size_t workspace_bytes,
                int64_t V, int64_t N1, int64_t N2,
                float* verts, int32_t* vert_edge, int32_t* faces_i32, int64_t* faces_i64, int64_t* uv_idx_i64,
                b2a_stream_t stream);
/* Static per-grid tables, once per loaded grid (DMTetGeometry.generate_edges, dmtet.py:283-288, + the tile skip table).
 * build_edges: the unique sorted (min,max) edges of the 6 T tet edges inside `workspace` (b2a_mt_tables_workspace_bytes),
 * edge_start [Vg+1] written, *num_edges (device or pinned-host int64) = E; then the caller allocates edge_b [E] and calls
 * emit_edges with the same workspace.  tets: int32 or int64 [T,4].  build_tile_words: int32 tets -> [ceil(T/tile), words]. */
int b2a_mt_tables_workspace_bytes(int64_t Vg, int64_t T, size_t* bytes);
int b2a_mt_build_edges(const void* tets, int tets_are_i64, int64_t Vg, int64_t T, void* workspace, size_t workspace_bytes,
                       int32_t* edge_start, int64_t* num_edges, b2a_stream_t stream);
int b2a_mt_emit_edges(const void* workspace, size_t workspace_bytes, int64_t Vg, int64_t T, int64_t E, int32_t* edge_b,
                      b2a_stream_t stream);
int b2a_mt_build_tile_words(const int32_t* tets, int64_t T, int32_t* tile_words, b2a_stream_t stream);
/* d_sdf [Vg] and d_pos [Vg,3] (nullable) must be zero-initialised by the caller; gradients are accumulated. */
int b2a_mt_bwd(const float* pos, const float* sdf, const int32_t* vert_edge, const float* d_verts, int64_t V,
               float* d_sdf, float* d_pos, b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Linear blend skinning.  Replaces skinning() (model/geometry/skinning.py:369-439) incl.
 * _compute_vertices_to_bones_weights (:16-22), line_segment_distance (geometry/util.py:30-53),
 * _estimate_bone_rotation (:251-270), euler_angles_to_matrix 'XYZ' (:315-340).
 * B = batch*frames.  bones [Bb,K,2,3] (Bb in {1,B}), angles [B,K,3] rad, v_pos [Bv,V,3] (Bv in {1,B}).
 * chain_ptr [K+1], chain_ids: for bone k the bones whose local transforms are multiplied, root first, k last.
 * G, T_local: [B,K,12] row-major 3x4 affine.  weights (nullable): [K,Bw,V], Bw = max(Bv,Bb).
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_lbs_bone_transforms(const float* bones, const float* angles, const int32_t* chain_ptr,
                            const int32_t* chain_ids, int B, int Bb, int K, float* T_local, float* G,
                            float* posed_bones /* nullable [B,K,2,3] */, b2a_stream_t stream);
int b2a_lbs_fwd(const float* v_pos, const float* bones, const float* G, int B, int Bv, int Bb, int K, int64_t V,
                float inv_temperature, float* out /* [B,V,3] */, float* weights /* nullable */, b2a_stream_t stream);
/* d_v_pos [Bv,V,3] zero-initialised when Bv==1<B (accumulated), else written; d_G [B,K,12] zero-initialised. */
int b2a_lbs_bwd(const float* v_pos, const float* bones, const float* G, const float* d_out, int B, int Bv, int Bb,
                int K, int64_t V, float inv_temperature, float* d_v_pos /* nullable */, float* d_G,
                b2a_stream_t stream);
/* d_G is consumed (d_posed_bones, nullable, is folded into it); d_T_local [B,K,12] is zero-initialised scratch. */
int b2a_lbs_bone_transforms_bwd(const float* bones, const float* angles, const int32_t* chain_ptr,
                                const int32_t* chain_ids, const float* T_local, float* d_G,
                                const float* d_posed_bones, int B, int Bb, int K, float* d_T_local,
                                float* d_angles /* [B,K,3] */, b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Bone placement heuristic.  Replaces estimate_bones (model/geometry/skinning.py:49-248) for body_bones_mode
 * 'z_minmax' (mode 0) / 'z_minmax_y+' (mode 1), resample=False.  bone_y_threshold = 0: the MagicPony / Ponymation leg
 * quadrants (bone_y_threshold=None, skinning.py:156-161); in (0,1]: the 3D-Fauna variant (InstancePredictorFauna.py:20,90;
 * skinning.py:163-175): quadrants centred on the medians of the vertices below that quantile of y, with margins from their
 * 5 % / 95 % quantiles in x and z - seven exact whole-batch quantiles, six of them over a data-dependent subset.
 * verts [N,V,3] (N = batch*frames) -> bones [N,K,2,3], K = n_body_bones + 4*n_leg_bones (legs only if n_leg_bones>0).
 * attach0..3: index of the body bone whose end joint each leg attaches to (aux['legs'][i]['body_bone_idx']); -1 = pick
 * the body bone closest in z to that instance's own foot (skinning.py:190-192) - attach_out [4] (nullable) receives
 * instance 0's choice, which is the one the reference then reuses for every instance.
 * The 5 %/95 % x-quantiles are taken over the WHOLE batch (skinning.py:157) by exact radix select, no sort, no sync.
 * stats_out (nullable) [N,8]: x_margin, mean xyz, bit-cast arg-max/arg-min vertex ids, x0, z0 (test hook).
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_estimate_bones_workspace_bytes(size_t* bytes);
int b2a_estimate_bones(const float* verts, int N, int64_t V, int n_body_bones, int n_leg_bones, int mode,
                       float bone_y_threshold, int attach0, int attach1, int attach2, int attach3, void* workspace,
                       size_t workspace_bytes, float* bones, int32_t* attach_out, float* stats_out, b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Smooth vertex normals.  Replaces mesh.auto_normals (model/render/mesh.py:276-304).
 * nsum [B,V,4] is the un-normalised area-weighted sum (xyz, 0), rows padded to 16 bytes (kept for the backward).
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_vertex_normals_fwd(const float* v_pos, const int32_t* tri, int B, int64_t V, int64_t F, float* nsum,
                           float* v_nrm, b2a_stream_t stream);
/* scratch [B,V,4]; d_v_pos [B,V,3] zero-initialised by the caller (accumulated). */
int b2a_vertex_normals_bwd(const float* v_pos, const int32_t* tri, const float* nsum, const float* d_v_nrm, int B,
                           int64_t V, int64_t F, float* scratch, float* d_v_pos, b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Clip-space transform.  Replaces ru.xfm_points(use_python=True) (model/render/renderutils/ops.py:524-525).
 * pts [Bp,V,3] (Bp in {1,B}), mtx [B,4,4] -> out [B,V,4].  Backward: upstream gradient d_out (+ d_out2, nullable: a second
 * contribution summed on the fly); d_pts accumulated when Bp==1<B (zero-init) or when `accumulate` is set, else written;
 * d_mtx [B,16] accumulated (zero-init); either may be NULL.
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_xfm_points_fwd(const float* pts, const float* mtx, int B, int Bp, int64_t V, float* out, b2a_stream_t stream);
int b2a_xfm_points_bwd(const float* pts, const float* mtx, const float* d_out, const float* d_out2, int accumulate,
                       int B, int Bp, int64_t V, float* d_pts, float* d_mtx, b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Rasterizer.  Replaces nvdiffrast.torch.rasterize / DepthPeeler first layer (call sites model/render/render.py:
 * 292-294, :351).  pos [B,V,4] clip space, tri [F,3]; rast [B,H,W,4] = (u, v, z/w, id+1).  Fill rule: oracle/
 * raster_ref.c header.  workspace: b2a_rasterize_workspace_bytes (z-buffer keys + large-triangle queue).
 * Backward: d_rast[...,0:2] -> d_pos (x,y,w), accumulated into zero-initialised d_pos [B,V,4].
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_rasterize_workspace_bytes(int B, int64_t F, int H, int W, size_t* bytes);
/* cov_list [B*H*W,4] / cov_count [1] (both nullable): compact, unordered list of covered pixels for b2a_gbuffer_bwd;
 * entry = (flat pixel index b*H*W + y*W + x, vertex ids i0, i1, i2 of the visible triangle); 16-byte aligned. */
int b2a_rasterize_fwd(const float* pos, const int32_t* tri, int B, int64_t V, int64_t F, int H, int W,
                      void* workspace, size_t workspace_bytes, float* rast, int32_t* cov_list, int32_t* cov_count,
                      b2a_stream_t stream);
int b2a_rasterize_bwd(const float* pos, const int32_t* tri, const float* rast, const float* d_rast, int B, int64_t V,
                      int64_t F, int H, int W, float* d_pos, b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Attribute interpolation.  Replaces nvdiffrast.torch.interpolate (render.py:24, :182-209).
 * attr [Ba,V,C] (Ba in {1,B}); out [B,H,W,C].  Backward: d_attr accumulated (zero-init, nullable),
 * d_rast [B,H,W,4] written (u,v grads; z,w = 0; nullable).
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_interpolate_fwd(const float* attr, const float* rast, const int32_t* tri, int B, int Ba, int64_t V, int64_t F,
                        int H, int W, int C, float* out, b2a_stream_t stream);
int b2a_interpolate_bwd(const float* attr, const float* rast, const int32_t* tri, const float* d_out, int B, int Ba,
                        int64_t V, int64_t F, int H, int W, int C, float* d_attr, float* d_rast, b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Antialiasing (+ optional fused composite).  Replaces nvdiffrast.torch.antialias and the lerp composite in
 * render_mesh.composite_buffer (model/render/render.py:258-268).
 *   edge adjacency: opp [F,3] = vertex opposite edge e in the neighbouring triangle, -1 if none
 *                   (e=0:(v1,v2), 1:(v2,v0), 2:(v0,v1)); built once per topology.
 *   composite = 0 : `color` is [B,H,W,C]; out [B,H,W,C].
 *   composite = 1 : `color` is [B,H,W,C-1] (no alpha, C >= 2); the blended input is
 *                   lerp(bg, [color,1], id>0) with bg [Bg,H,W,C] (Bg in {1,B}) or NULL (zeros).
 * Backward: d_out has Cg <= C channels (the channels the caller sliced off carry zero gradient) and arbitrary
 * element strides (sb,sy,sx,sc) (NCHW or NHWC views both fine); d_color (nullable) has the layout of `color`;
 * d_pos [B,V,4] accumulated (zero-init, nullable).
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_edge_adjacency_workspace_bytes(int64_t F, size_t* bytes);
int b2a_edge_adjacency(const int32_t* tri, int64_t F, int64_t V, void* workspace, size_t workspace_bytes,
                       int32_t* opp, b2a_stream_t stream);
/* Optional per-render context (composite mode, H*W % 32 == 0): one pass over rast builds 1-bit/pixel coverage and
 * silhouette masks and the compact list of silhouette pixels, a second small launch runs the pixel-pair analysis once
 * per silhouette pair (needs pos [B,V,4], tri, opp).  Every antialias launch of that render (each key, fwd and bwd) is
 * then a single streaming pass that never touches rast / pos / tri.  aa_ctx may be NULL (generic kernels). */
int b2a_antialias_workspace_bytes(int B, int H, int W, size_t* bytes);
int b2a_antialias_prepare(const float* rast, const float* pos, const int32_t* tri, const int32_t* opp, int B, int64_t V,
                          int64_t F, int H, int W, void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream);
int b2a_antialias_fwd(const float* color, const float* bg, int Bg, int composite, const float* rast, const float* pos,
                      const int32_t* tri, const int32_t* opp, int B, int64_t V, int64_t F, int H, int W, int C,
                      float* out, const void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream);
int b2a_antialias_bwd(const float* color, const float* bg, int Bg, int composite, const float* rast, const float* pos,
                      const int32_t* tri, const int32_t* opp, const float* d_out, int64_t d_sb, int64_t d_sy,
                      int64_t d_sx, int64_t d_sc, int Cg, int B, int64_t V, int64_t F, int H, int W, int C,
                      float* d_color, float* d_pos, const void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream);
/* Fused fast path of render_mesh's composite_buffer loop (model/render/render.py:311-331) for the training pair of keys:
 * 'dino_pred' (wide, Cw = D+1 = 17 channels) and 'shaded' (narrow, Cn = 4) composited + antialiased in ONE launch per
 * direction, sharing the render's prepared context.  Arguments as in b2a_antialias_fwd/bwd with composite = 1, one set
 * per key (suffix _w / _n); d_pos [B,V,4] receives the SUM of both keys' position gradients (zero-init, nullable).
 * Only (Cw, Cn) = (17, 4) with Cgw = 16, Cgn in {3, 4} is instantiated: other combinations return an error and the
 * caller issues the two single-key calls. */
int b2a_antialias_pair_fwd(const float* color_w, const float* bg_w, int Bg_w, int Cw, float* out_w, const float* color_n,
                           const float* bg_n, int Bg_n, int Cn, float* out_n, int B, int H, int W, const void* aa_ctx,
                           size_t aa_ctx_bytes, b2a_stream_t stream);
int b2a_antialias_pair_bwd(const float* color_w, const float* bg_w, int Bg_w, int Cw, const float* d_out_w, int64_t w_sb,
                           int64_t w_sy, int64_t w_sx, int64_t w_sc, int Cgw, float* d_color_w, const float* color_n,
                           const float* bg_n, int Bg_n, int Cn, const float* d_out_n, int64_t n_sb, int64_t n_sy,
                           int64_t n_sx, int64_t n_sc, int Cgn, float* d_color_n, int B, int64_t V, int H, int W,
                           float* d_pos, const void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream);

/* msaa / logging fast path of the composite loop (model/render/render.py:217-219 + :258-268, :311-331): `color` is the
 * shaded buffer at the g-buffer resolution [B,H/up,W/up,C-1] and is nearest-upsampled IN PLACE while it is composited
 * (and, with antialias = 1, antialiased) at the raster resolution [B,H,W]; antialias = 0 composites only (kd, normal,
 * geo_normal).  Narrow keys (C in 2..4) with a prepared context.  Backward: d_color at LOW resolution (the sum over each
 * up x up block), d_pos [B,V,4] accumulated (zero-init, nullable; untouched when antialias = 0). */
int b2a_composite_up_fwd(const float* color, int up, const float* bg, int Bg, int antialias, int B, int H, int W, int C,
                         float* out, const void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream);
int b2a_composite_up_bwd(const float* color, int up, const float* bg, int Bg, int antialias, const float* d_out,
                         int64_t d_sb, int64_t d_sy, int64_t d_sx, int64_t d_sc, int Cg, int B, int64_t V, int H, int W,
                         int C, float* d_color, float* d_pos, const void* aa_ctx, size_t aa_ctx_bytes,
                         b2a_stream_t stream);
/* The same, followed by the spp x spp average of the result (render.py:322-323 util.avg_pool_nhwc) in the same kernel:
 * out [B,H/up,W/up,keep] (the leading `keep` of the C channels); the raster-resolution image is never materialised.
 * d_out: gradient of that average, strides over [B,H/up,W/up,Cg].  Bit-identical to b2a_composite_up_fwd + avg_pool2d. */
int b2a_composite_up_pool_fwd(const float* color, int up, const float* bg, int Bg, int antialias, int B, int H, int W, int C,
                              int keep, float* out, const void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream);
int b2a_composite_up_pool_bwd(const float* color, int up, const float* bg, int Bg, int antialias, const float* d_out,
                              int64_t d_sb, int64_t d_sy, int64_t d_sx, int64_t d_sc, int Cg, int B, int64_t V, int H, int W,
                              int C, float* d_color, float* d_pos, const void* aa_ctx, size_t aa_ctx_bytes,
                              b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Directional-light diffuse shading (DirectionalLight.shade, model/render/light.py:186-193):
 *   shading = ambient + diffuse * clamp(dot(light_dir, normal), min=0);  shaded = shading * kd.
 * kd rows are kd_stride floats apart (3 = dense [B,HW,3]; 9 = the leading channels of the texture field's [B,HW,9]
 * output, read in place); normal [B,HW,3]; light [Bl,5] = (dir.xyz, ambient, diffuse), Bl in {1,B}.
 * Outputs shaded [B,HW,3], shading [B,HW] (nullable).  Backward: d_shaded [B,HW,3], d_shading [B,HW] (nullable) ->
 * d_kd [B,HW,3] dense, d_normal [B,HW,3], d_light [Bl,5] accumulated (zero-init); each nullable.
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_shade_directional_fwd(const float* kd, int64_t kd_stride, const float* normal, const float* light, int Bl, int B,
                              int64_t HW, float* shaded, float* shading, b2a_stream_t stream);
int b2a_shade_directional_bwd(const float* kd, int64_t kd_stride, const float* normal, const float* light, int Bl,
                              const float* d_shaded, const float* d_shading, int B, int64_t HW, float* d_kd,
                              float* d_normal, float* d_light, b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Benchmark stand-in for the field MLPs - NOT a reference interface.  SURVEY.md §8d defines M1a with the texture and
 * DINO CoordMLPs replaced by a fixed analytic function of the canonical position x [N,3]:  squash = 1: out [N,3C] =
 * three copies of sigmoid(x W) (kd|ks|normal);  squash = 0: out [N,C] = sin(x W);  W [3,C], C <= 16.  One kernel per
 * direction so the stand-in costs as little as possible on the number that measures the hot-path kernels.
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_analytic_field_fwd(const float* x, const float* weight, int C, int squash, int64_t N, float* out,
                           b2a_stream_t stream);
int b2a_analytic_field_bwd(const float* x, const float* weight, int C, int squash, int64_t N, const float* d_out,
                           float* d_x, b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused g-buffer pass (fast path of render_layer + shade's geometry part, model/render/render.py:160-209, :72-75):
 * interpolate v_pos, v_nrm, prior v_pos and the per-face normal at each covered pixel; prepare_shading_normal
 * (renderutils/ops.py:194-227 -> bsdf.py:46-51 with perturbed normal (0,0,1)); camera-space normal
 * safe_normalize(n . R_w2c^T) (render.py:75).
 * rast is the rasterizer output at [B,H*spp,W*spp,4]; the g-buffer is shaded at [H,W] from the nearest-downscaled
 * rast (pixel (x*spp, y*spp)), which is what render_layer does when msaa is on (render.py:170-172).
 * Outputs (each nullable): gb_pos, gb_geo_nrm, gb_shading_nrm, gb_cam_nrm, gb_tex_pos [B,H,W,3].
 * v_pos, v_nrm [B,V,3]; prior_pos [Bq,V,3] (Bq in {1,B}); w2c [B,4,4]; campos [B,3].
 * Backward consumes the same inputs, pos_clip [B,V,4] and the (nullable) output grads.  Vertex gradients are summed with
 * 16-byte vector reductions into a [B,V,12] accumulator (workspace, zeroed inside the call) and then WRITTEN (not
 * accumulated) to the nullable d_v_pos [B,V,3], d_v_nrm [B,V,3], d_prior_pos [Bq,V,3], d_clip [B,V,4] (x,y,w; z = 0);
 * d_w2c [B,16] and d_campos [B,3] (nullable) are accumulated into zero-initialised buffers.
 * workspace_is_zero = 1: the caller guarantees the accumulator is all zeros on entry (no memset is issued) and gets
 * it back zeroed - for callers that keep one accumulator per device across calls.
 * cov_list / cov_count (nullable, spp == 1 only): the rasterizer's compact covered-pixel list - dense warps.
 * ---------------------------------------------------------------------------------------------------------- */
/* packed (nullable): caller-owned scratch of b2a_gbuffer_pack_bytes bytes; the forward re-lays the vertex attributes out
 * as 16-byte records there (the per-pixel gather is bound by load instructions, not bytes) and the backward re-uses it. */
int b2a_gbuffer_pack_bytes(int B, int Bq, int64_t V, size_t* bytes);
int b2a_gbuffer_fwd(const float* rast, int spp, const int32_t* tri, const float* v_pos, const float* v_nrm,
                    const float* prior_pos, int Bq, const float* w2c, const float* campos, int two_sided, int B,
                    int64_t V, int64_t F, int H, int W, void* packed, size_t packed_bytes, float* gb_pos,
                    float* gb_geo_nrm, float* gb_shading_nrm, float* gb_cam_nrm, float* gb_tex_pos, b2a_stream_t stream);
int b2a_gbuffer_bwd_workspace_bytes(int B, int64_t V, size_t* bytes);
int b2a_gbuffer_bwd(const float* rast, int spp, const float* pos_clip, const int32_t* tri, const float* v_pos,
                    const float* v_nrm, const float* prior_pos, int Bq, const float* w2c, const float* campos,
                    int two_sided, int B, int64_t V, int64_t F, int H, int W, const void* packed, size_t packed_bytes,
                    const int32_t* cov_list, const int32_t* cov_count, const float* d_gb_pos, const float* d_gb_geo_nrm,
                    const float* d_gb_shading_nrm, const float* d_gb_cam_nrm, const float* d_gb_tex_pos,
                    void* workspace, size_t workspace_bytes, int workspace_is_zero, float* d_v_pos, float* d_v_nrm,
                    float* d_prior_pos, float* d_clip, float* d_w2c, float* d_campos, b2a_stream_t stream);

/* Fused geometry half of render_mesh, one call per direction (the head of render_mesh, model/render/render.py:270-296:
 * ru.xfm_points + dr.DepthPeeler(...).rasterize_next_layer(); render_layer's interpolations :160-209 + the shading / camera
 * normal of shade :72-75; and the discontinuity analysis nvdiffrast's antialias re-runs inside every call, :264).
 * fwd = b2a_xfm_points_fwd -> b2a_rasterize_fwd at [H*spp, W*spp] -> b2a_gbuffer_fwd at [H,W] -> b2a_antialias_prepare
 * (skipped when aa_ctx is NULL); every buffer is the caller's: raster_ws (b2a_rasterize_workspace_bytes(B,F,H*spp,W*spp)),
 * packed (b2a_gbuffer_pack_bytes), clip [B,V,4], rast [B,H*spp,W*spp,4], cov_list / cov_count (nullable, spp == 1),
 * the nullable g-buffers [B,H,W,3], aa_ctx (b2a_antialias_workspace_bytes(B,H*spp,W*spp)).
 * bwd = the g-buffer / rasterize adjoint over the covered-pixel list with the clip-space positions RECOMPUTED from v_pos and mtx
 * (no pos_clip gather), then ONE per-vertex pass that adds the upstream clip gradient d_clip_up [B,V,4] (nullable: the
 * antialias position gradient), applies the clip-transform adjoint and writes d_v_pos, d_v_nrm [B,V,3], d_prior_pos [Bq,V,3]
 * (each nullable, WRITTEN) and accumulates d_mtx [B,16], d_w2c [B,16], d_campos [B,3] (nullable, zero-initialised by the
 * caller).  workspace / workspace_is_zero as in b2a_gbuffer_bwd.  With every d_gb_* NULL only the clip-transform adjoint runs. */
int b2a_render_geometry_fwd(const float* v_pos, const float* v_nrm, const float* prior_pos, int Bq, const float* mtx,
                            const float* w2c, const float* campos, const int32_t* tri, const int32_t* opp, int two_sided,
                            int B, int64_t V, int64_t F, int H, int W, int spp, void* raster_ws, size_t raster_ws_bytes,
                            void* packed, size_t packed_bytes, float* clip, float* rast, int32_t* cov_list,
                            int32_t* cov_count, float* gb_pos, float* gb_geo_nrm, float* gb_shading_nrm, float* gb_cam_nrm,
                            float* gb_tex_pos, void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream);
int b2a_render_geometry_bwd(const float* rast, int spp, const float* mtx, const int32_t* tri, const float* v_pos,
                            const float* v_nrm, const float* prior_pos, int Bq, const float* w2c, const float* campos,
                            int two_sided, int B, int64_t V, int64_t F, int H, int W, const void* packed, size_t packed_bytes,
                            const int32_t* cov_list, const int32_t* cov_count, const float* d_gb_pos,
                            const float* d_gb_geo_nrm, const float* d_gb_shading_nrm, const float* d_gb_cam_nrm,
                            const float* d_gb_tex_pos, const float* d_clip_up, void* workspace, size_t workspace_bytes,
                            int workspace_is_zero, float* d_v_pos, float* d_v_nrm, float* d_prior_pos, float* d_mtx,
                            float* d_w2c, float* d_campos, b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Covered-row gather / scatter between a dense shaded image [n, C] and the compact rows [N, C] the field networks are
 * evaluated on (the reference samples material / dino_net densely, model/render/render.py:54,61; uncovered pixels never
 * reach an output).  idx [N] int64 unique row numbers.  scatter writes dst[idx[r]] = src[r]; zero_fill != 0 clears dst
 * [dst_rows, C] first.
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_rows_gather(const float* src, const int64_t* idx, int64_t N, int C, float* out, b2a_stream_t stream);
int b2a_rows_scatter(const float* src, const int64_t* idx, int64_t N, int C, float* dst, int64_t dst_rows, int zero_fill,
                     b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Field MLPs on the tensor cores (tcgen05 + tensor memory).  Replaces the fp32 GEMMs of CoordMLP.forward
 * (model/networks/MLPs.py:34-101) behind material.sample / dino_net.sample (model/render/render.py:54,61) on the covered
 * rows.  Every fp32 operand is split into two bf16 terms and each product is three MMAs with fp32 accumulation
 * (passes = 3; fp32-grade) or one (passes = 1; the arithmetic of the reference's autocast configs).
 * pack_weights: W [N,K] fp32 row-major with row stride ldw (transpose = 1: W is [K,N] and its transpose is packed) ->
 * `packed` (b2a_mlp_packed_bytes), the shared-memory image the GEMM fetches with bulk async copies; N <= 256.
 * rows_gemm: out[rows,N] (row stride ldo) = epilogue(op(A)[rows,K] . W^T) with A fp32 (row stride lda, lda % 4 == 0);
 * relu_on_load: op = max(., 0).  epilogue 0: + bias (nullable; [N], or row bias_rows[r] of a [*,N] table when bias_rows is
 * given); 1: zero where mask_src[r,n] <= 0 (row stride ldm) - or, when mask_bits [rows, ceil16(N)/32 words] is given, where its
 * bit n is clear (the sign bits an epilogue-0 call wrote through bits_out: 1/32 of the mask bytes); 2: sigmoid(. + bias).
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_mlp_packed_bytes(int N, int K, size_t* bytes);
int b2a_mlp_pack_weights(const float* W, int64_t ldw, int N, int K, int transpose, void* packed, size_t packed_bytes,
                         b2a_stream_t stream);
/* the same for n <= 32 weight matrices in one launch: jobs = HOST array [n][6] of int64 {W (device pointer), ldw, N, K,
 * transpose, packed (device pointer)} */
int b2a_mlp_pack_weights_many(const int64_t* jobs, int n, b2a_stream_t stream);
int b2a_mlp_rows_gemm(const float* A, int64_t lda, int64_t rows, int K, const void* packed, int N, int relu_on_load,
                      int passes, int epilogue, const float* bias, const int32_t* bias_rows, const float* mask_src,
                      int64_t ldm, const uint32_t* mask_bits, uint32_t* bits_out, float* out, int64_t ldo,
                      b2a_stream_t stream);
/* wgrad: out[m*ldo + n] (transpose_out: out[n*ldo + m]) += sum_r op(P)[r,m] * op(Q)[r,n] over all rows (the weight gradient
 * dz^T . a of a Linear layer); out is ACCUMULATED with reductions and must be zero-initialised by the caller; M, N <= 256.
 * embed fwd / bwd: the harmonic embedding in front of CoordMLP.in_layer (model/networks/HarmonicEmbedding.py; MLPs.py:75-84):
 * E[r] = [x', sin(x' f), cos(x' f)] padded with zeros to ldE floats, x' = (|x|, y, z) when symmetrize; and its adjoint.
 * colsum_segments: out[s,n] = sum of G[r,n] over seg_start[s] <= r < seg_start[s+1] (adjoint of a per-image bias). */
int b2a_mlp_wgrad(const float* P, int64_t ldp, int relu_p, const float* Q, int64_t ldq, int relu_q, int64_t rows, int M,
                  int N, int passes, float* out, int64_t ldo, int transpose_out, b2a_stream_t stream);
int b2a_mlp_embed_fwd(const float* x, int64_t ldx, int64_t rows, int n_harmonic, float scalar, int symmetrize,
                      int concat_pts, float* E, int64_t ldE, b2a_stream_t stream);
int b2a_mlp_embed_bwd(const float* x, int64_t ldx, int64_t rows, int n_harmonic, float scalar, int symmetrize,
                      int concat_pts, const float* dE, int64_t ldE, float* d_x, int64_t lddx, b2a_stream_t stream);
int b2a_mlp_colsum_segments(const float* G, int64_t ldg, const int64_t* seg_start, int n_seg, int N, float* out,
                            b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Mesh export: the text of a Wavefront OBJ file.  Replaces the per-line loops of write_obj
 * (model/render/obj.py:128-177); byte-identical output, including the reference's number format: every coordinate is
 * '{}'.format(np.float32) = Python's repr of the value widened to double, the texcoord v is flipped in float32 first
 * (obj.py:148).  HOST pointers (the only entry points of this library that take them; no CUDA work, no stream).
 * Text: "mtllib <name>.mtl\ng default\n", 'v' lines, 'vt' lines (n_tex of them), 'vn' lines, "s 1 \ng pMesh1\n
 * usemtl defaultMat\n", 'f' lines "f  a/b/c a/b/c a/b/c" with 1-based indices; the b / c columns are empty when v_tex /
 * v_nrm is NULL (obj.py:166).  n_tex = 0 with a non-NULL v_tex is the reference's save_material=False case (:145).
 * `out` must hold at least b2a_obj_text_bound bytes; lines are formatted by `threads` host threads (0 = all) at
 * worst-case offsets and compacted in place; *written = size of the text.
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_obj_text_bound(int64_t n_pos, int64_t n_tex, int64_t n_nrm, int64_t n_faces, int64_t name_bytes, size_t* bytes);
int b2a_obj_format(const float* v_pos, int64_t n_pos, const float* v_tex, int64_t n_tex, const float* v_nrm,
                   int64_t n_nrm, const int64_t* t_pos_idx, const int64_t* t_tex_idx, const int64_t* t_nrm_idx,
                   int64_t n_faces, const char* mtl_name, int64_t name_bytes, char* out, size_t capacity,
                   size_t* written, int threads);

/* ------------------------------------------------------------------------------------------------------------
 * Articulation-angle constraints (csrc/articulation.cu).  Replaces InstancePredictorBase.apply_articulation_constraints
 * (model/predictors/InstancePredictorBase.py:435-511) and Fauna's form (InstancePredictorFauna.py:149-212 + :225-228):
 * out = post_S(..post_1(tanh(pre_P(..pre_1(x))))), every stage one fp32 multiplication (division where bit j of
 * post_div_mask is set) by table[stage][bone][axis].  x, out, d_out, d_x: [rows, K, 3]; pre [n_pre,K,3], post [n_post,K,3].
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_articulation_constraints_fwd(const float* x, const float* pre, int n_pre, const float* post, int n_post,
                                     int post_div_mask, int64_t rows, int K, float* out, b2a_stream_t stream);
int b2a_articulation_constraints_bwd(const float* x, const float* pre, int n_pre, const float* post, int n_post,
                                     int post_div_mask, int64_t rows, int K, const float* d_out, float* d_x,
                                     b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Gradient all-reduce over NVLink peer memory (csrc/allreduce_p2p.cu).  Replaces, for the ranks of one node, the
 * DistributedDataParallel / accelerate gradient all-reduce the reference's trainer relies on (model/trainer/Trainer.py:170-180,
 * accelerator.prepare + accelerator.backward): average of n floats over `world` ranks in ONE kernel launch per rank.
 * b2a_p2p_alloc: a zero-filled device buffer other processes can map, and its 64-byte IPC handle; b2a_p2p_open maps a
 * peer's buffer into this process on the current device.  b2a_allreduce_p2p: bufs / flags are host arrays of `world`
 * device pointers (own buffer at index `rank`); flags: 16 zero-initialised 32-bit words per rank per channel;
 * local_counter: two zero-initialised device words per channel (grid-barrier counter; error word, nonzero after a peer
 * failed to arrive within ~10 s); epoch = 1, 2, 3, ... per channel, equal on all ranks.
 * ---------------------------------------------------------------------------------------------------------- */
int b2a_p2p_alloc(size_t bytes, void** ptr, void* handle64);
int b2a_p2p_open(const void* handle64, void** ptr);
int b2a_p2p_close(void* ptr);
int b2a_p2p_free(void* ptr);
int b2a_allreduce_p2p(const void* const* bufs, const void* const* flags, int rank, int world, int64_t n, int epoch,
                      void* local_counter, b2a_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B2A_H */
