/*
 * homan_b200 — C ABI of the B200-native hand-object fitting kernels.
 *
 * The reference (hassony2/homan) has no FFI/plugin interface: its seams for this path are the
 * Python call signatures of four third-party packages and of its own modules.  Every entry point
 * below names the reference interface it replaces (file:line relative to the reference root).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns all memory,
 *     the library never allocates;
 *   - every launch goes on the `stream` argument (a cudaStream_t passed as void*), no entry point
 *     synchronises, so a sequence of calls can be captured in a CUDA graph;
 *   - return value: 0 on success, negative hm_status otherwise; hm_last_error() gives the message
 *     of the last failure on the calling thread;
 *   - layouts are row-major, fp32 / int32 unless stated; B = number of images (problems x frames).
 */
#ifndef HOMAN_B200_H
#define HOMAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HM_OK 0
#define HM_ERR_INVALID (-1)
#define HM_ERR_CUDA (-2)
#define HM_ERR_UNSUPPORTED (-3)

#define HM_VERSION 100

int hm_version(void);
const char *hm_last_error(void);

/* ---------------------------------------------------------------- projection
 * nr.projection(vertices, K, R, t, dist_coeffs, orig_size)  (neural_renderer, un-vendored;
 * call sites homan/losses.py:34-41 and inside Renderer.render_silhouettes, homan/losses.py:187).
 * verts [B,V,3] -> ndc [B,V,3] = (u, v, z).  K [K_batch,3,3] (K_batch 1 or B); R [9], t [3],
 * dist [dist_batch,5] may be NULL (identity / zero). */
int hm_project_fwd(const float *verts, const float *K, int K_batch, const float *R, const float *t,
                   const float *dist, int dist_batch, float orig_size, float eps, int B, int V, float *ndc,
                   void *stream);
/* backward of the above w.r.t. verts (zero distortion only); grad_verts += or = per `accumulate`. */
int hm_project_bwd(const float *verts, const float *K, int K_batch, const float *R, const float *t,
                   float orig_size, float eps, int B, int V, const float *grad_ndc, float *grad_verts,
                   int accumulate, void *stream);

/* ---------------------------------------------------------------- silhouette rasteriser
 * nr.Renderer(...)(vertices, faces, mode="silhouettes")  (homan/losses.py:73-77,172-176,187;
 * homan/homan.py:168-176): fill_back face doubling + vertices_to_faces + rasterize_silhouettes
 * (z-buffered nearest front face, vertical flip, 2x2 average pool when anti-aliasing) and its
 * hand-crafted backward (backward_pixel_map).  image_size is the OUTPUT size R; the raster size is
 * S = 2R with anti-aliasing.  S must be a multiple of 64. */
#define HM_FACE_RECORD_BYTES 224
#define HM_FACE_BBOX_BYTES 8
/* ndc [B,V,3], faces [faces_batch,F,3] -> records (B*F*HM_FACE_RECORD_BYTES bytes: a plane of 128-byte forward
 * records followed by a plane of 96-byte backward records) + bboxes (HM_FACE_BBOX_BUFFER_BYTES(B, F) bytes: the pixel
 * box of every face [B,F,8 B], then, 16-byte aligned, the pixel box of every image [B,16 B]: tiles outside it are
 * skipped). Front-facing winding of every face; never materialises the doubled face array. */
#define HM_FACE_BBOX_BUFFER_BYTES(B, F) \
    ((((size_t)(B) * (size_t)(F) * HM_FACE_BBOX_BYTES + 15) & ~(size_t)15) + (size_t)(B) * 16)
int hm_raster_setup(const float *ndc, const int32_t *faces, int faces_batch, int B, int V, int F,
                    int image_size, int anti_aliasing, int fill_back, void *records, void *bboxes,
                    void *stream);
/* -> face_index [B,S,S] (doubled numbering: f or F+f, -1 background; raster frame, row 0 = y -1),
 *    alpha [B,R,R] (after flip + pool), coverage bitmaps cov_row / cov_col [B,S,S/32] (may be NULL),
 *    face_vis [B,HM_FACE_VIS_WORDS(F)] (may be NULL): bit fn set iff face fn of the doubled numbering owns a pixel,
 *    cov_blocks [B,S/8,S/32] bytes (may be NULL): bit j of [band][w] set iff the 8x8 block 4w + j of the 8-row band has
 *    an uncovered pixel.  The last two are what hm_raster_sil_bwd culls hidden faces and in-sweeps with. */
#define HM_FACE_VIS_WORDS(F) ((2 * (F) + 31) / 32)
int hm_raster_sil_fwd(const void *records, const void *bboxes, int B, int F, int image_size,
                      int anti_aliasing, float near_, float far_, int32_t *face_index, float *alpha,
                      uint32_t *cov_row, uint32_t *cov_col, uint32_t *face_vis, uint8_t *cov_blocks, void *stream);
/* grad_alpha [B,R,R] + coverage -> sweep masks m_row / m_col [B,2,S,S/32]
 * (0: uncovered & grad<0, 1: covered & grad>0) and their run-length form:
 * runs [B,4,S,HM_RASTER_RUN_CAP] (8 B each), run_counts [B,4,S] (count | first pixel << 4 | last pixel << 16). */
#define HM_RASTER_RUN_CAP 8
int hm_raster_grad_prep(const float *grad_alpha, const uint32_t *cov_row, const uint32_t *cov_col, int B,
                        int image_size, int anti_aliasing, uint32_t *m_row, uint32_t *m_col, void *runs,
                        uint32_t *run_counts, void *stream);
/* backward_pixel_map fused with the vertices_to_faces scatter-add:
 * grad_ndc [B,V,3] += d loss / d (u, v) of every vertex (z receives nothing in silhouette mode). */
int hm_raster_sil_bwd(const void *records, const void *bboxes, const int32_t *face_index,
                      const float *grad_alpha, const uint32_t *cov_row, const uint32_t *cov_col,
                      const uint32_t *face_vis, const uint8_t *cov_blocks,
                      const uint32_t *m_row, const uint32_t *m_col, const void *runs, const uint32_t *run_counts,
                      int B, int V, int F, int image_size, int anti_aliasing, float eps, float *grad_ndc,
                      unsigned long long *grad_fixed, void *stream);
/* Order-independent accumulation (test mode, SURVEY.md section 5 "deterministic reduction"): when an entry point is
 * given a `grad_fixed` buffer (same shape as the float gradient it shadows, 64-bit, zeroed by the caller) its scattered
 * contributions are rounded to 2^-44 fixed point and summed with integer atomics instead of float atomics, so the sum
 * does not depend on the arrival order of CTAs / warps (|contribution| < 2^19). hm_fold_fixed: dst[i] += fixed[i] * 2^-44. */
int hm_fold_fixed(const unsigned long long *fixed, int n, float *dst, void *stream);

/* Forward of nr.Renderer.render / rasterize_rgbad for texture_size 1 (visualisation: homan/homan.py:510-613,
 * homan/visualize.py:44-128, homan/utils/nmr_renderer.py:71,164,209), on top of hm_raster_setup + hm_raster_sil_fwd:
 * colours [colours_batch, 2F, 3] = lit colour of every face in the doubled numbering (f, F + f);
 * rgb [B,3,R,R] (may be NULL) = colour of the winning face or the background, depth [B,R,R] (may be NULL) = the
 * reference's interpolated depth or `far`; both after the vertical flip and the 2x2 average when anti-aliasing. */
/* nr.lighting for flat per-face colours (the lighting step of Renderer.render, texture_size 1): verts [B,V,3] (3-D, the
 * space the light direction lives in), faces [faces_batch,F,3], colours [colours_batch,F,3]; the three light vectors
 * are HOST pointers to 3 floats. -> lit [B,2F,3] in the doubled numbering hm_raster_shade reads (F + f = the reversed
 * copy, lit from its flipped normal; zeros when fill_back is 0). */
int hm_face_lighting(const float *verts, const int32_t *faces, int faces_batch, const float *colours, int colours_batch,
                     int B, int V, int F, int fill_back, float intensity_ambient, float intensity_directional,
                     const float *color_ambient, const float *color_directional, const float *direction, float *lit,
                     void *stream);
int hm_raster_shade(const void *records, const int32_t *face_index, const float *colours, int colours_batch, int B,
                    int F, int image_size, int anti_aliasing, float far_, float bg_r, float bg_g, float bg_b,
                    float *rgb, float *depth, void *stream);
/* Losses.compute_sil_loss_object (homan/losses.py:183-197) on a rendered alpha:
 * target int8 [B,R,R] in {-1 occluded, 0, 1}; norm [B] = 1 / (sum keep * T) of the image's problem;
 * loss_img[b*loss_stride] = norm * sum (keep*alpha - ref)^2 ; iou_img[b*iou_stride] (mask IoU, metric);
 * grad_alpha = weight * 2 * norm * keep * (keep*alpha - ref). */
int hm_sil_loss_fwd_bwd(const float *alpha, const int8_t *target, const float *norm, float weight, int B,
                        int image_size, float *loss_img, int loss_stride, float *iou_img, int iou_stride,
                        float *grad_alpha, void *stream);
/* hm_sil_loss_fwd_bwd followed by hm_raster_grad_prep in one kernel (same outputs, bit for bit): the loss gradient of
 * a silhouette rendered by hm_raster_sil_fwd takes nine values per image, so it is kept as one byte per pixel in shared
 * memory and the sweep masks / run lists are derived from it without re-reading grad_alpha. Precondition: alpha is the
 * output of hm_raster_sil_fwd (multiples of 1/4). image_size a multiple of 32, at most 384. */
int hm_sil_loss_prep(const float *alpha, const int8_t *target, const float *norm, float weight, int B, int image_size,
                     int anti_aliasing, float *loss_img, int loss_stride, float *iou_img, int iou_stride,
                     float *grad_alpha, const uint32_t *cov_row, const uint32_t *cov_col, uint32_t *m_row,
                     uint32_t *m_col, void *runs, uint32_t *run_counts, void *stream);

/* ---------------------------------------------------------------- MANO + rigid placement of the hand
 * ManoModel.forward_pca (homan/manomodel.py:84-151) -> mano layer LBS (un-vendored `mano` package,
 * manomodel.py:119-123,136-140) -> + mano_trans -> rot6d_to_matrix (homan/utils/geometry.py:9-27) ->
 * compute_transformation_persp (homan/utils/camera.py:108-139), i.e. HOMan.get_verts_hand
 * (homan/homan.py:341-382), and its autograd.
 * `model` is one fp32 blob (offsets in floats below); J_template = J_regressor @ v_template and
 * J_shapedirs = J_regressor @ shapedirs are folded by the caller; comps holds ncomps rows of 45.
 * pca [B,pca_stride] (first ncomps used), rot [B,3] (mano global orient), betas [B,10] (NULL = 0),
 * mano_trans [B,3] (NULL = 0), rot6d [B,3,2] (NULL = no rigid placement: plain MANO layer output),
 * trans [B,3], scale [1] (NULL = 1).  verts [B,778,3]; joints [B,16,3] (mano frame, may be NULL). */
#define HM_MANO_NV 778
#define HM_MANO_NJ 16
#define HM_MANO_OFF_VTEMPLATE 0
#define HM_MANO_OFF_SHAPEDIRS (HM_MANO_OFF_VTEMPLATE + 778 * 3)        /* [778,3,10] */
#define HM_MANO_OFF_POSEDIRS (HM_MANO_OFF_SHAPEDIRS + 778 * 3 * 10)    /* [135,2334] */
#define HM_MANO_OFF_JTEMPLATE (HM_MANO_OFF_POSEDIRS + 135 * 2334)      /* [16,3] */
#define HM_MANO_OFF_JSHAPEDIRS (HM_MANO_OFF_JTEMPLATE + 48)            /* [16,3,10] */
#define HM_MANO_OFF_WEIGHTS (HM_MANO_OFF_JSHAPEDIRS + 480)             /* [778,16] */
#define HM_MANO_OFF_MEAN (HM_MANO_OFF_WEIGHTS + 778 * 16)              /* [45] (+3 pad) */
#define HM_MANO_OFF_COMPS (HM_MANO_OFF_MEAN + 48)                      /* [ncomps,45] */
#define HM_MANO_BLOB_FLOATS(ncomps) (HM_MANO_OFF_COMPS + (ncomps) * 45)
int hm_mano_fwd(const float *model, int ncomps, int left, const float *pca, int pca_stride, const float *rot,
                const float *betas, const float *mano_trans, const float *rot6d, const float *trans,
                const float *scale, int B, float *verts, float *joints, float *vposed, void *stream);
/* vposed [B,778*3] (may be NULL in both calls): the posed template (v_template + shape + pose blend shapes) that
 * hm_mano_fwd writes and hm_mano_bwd of the same iteration reads instead of streaming posedirs a second time.
 * grad_verts [B,778,3]; grad_centroid_det [B,3] (may be NULL) is d loss / d mean_v(verts) through the
 * mesh-detached twin of compute_transformation_persp (reaches rot6d / trans only; homan/homan.py:484-490).
 * All outputs are accumulated (+=) and may be NULL. */
int hm_mano_bwd(const float *model, int ncomps, int left, const float *pca, int pca_stride, const float *rot,
                const float *betas, const float *mano_trans, const float *rot6d, const float *trans,
                const float *scale, int B, const float *vposed, const float *grad_verts,
                const float *grad_centroid_det, float *grad_pca, float *grad_rot, float *grad_betas,
                float *grad_mano_trans, float *grad_rot6d, float *grad_trans, void *stream);

/* ---------------------------------------------------------------- rigid placement of the object
 * HOMan.get_verts_object (homan/homan.py:298-307): verts = (|scale| * mesh) @ rot6d_to_matrix(rot6d) + trans.
 * mesh [mesh_batch,V,3] (mesh_batch 1 or B), rot6d [B,3,2], trans [B,3], scale [1] (NULL = 1). */
int hm_rigid_fwd(const float *mesh, int mesh_batch, const float *rot6d, const float *trans, const float *scale,
                 int B, int V, float *verts, void *stream);
/* grad_rot6d [B,6] / grad_trans [B,3] += ; the scale is a buffer on this path (optimize_object_scale=0). */
int hm_rigid_bwd(const float *mesh, int mesh_batch, const float *rot6d, const float *scale, int B, int V,
                 const float *grad_verts, float *grad_rot6d, float *grad_trans, void *stream);

/* ---------------------------------------------------------------- vertex-space losses (fused fwd + bwd)
 * compute_smooth_loss (homan/lossutils.py:18-36), compute_verts2d_loss_hand (homan/losses.py:141-164),
 * compute_interaction_loss + assign_interaction_pairs + project_bbox (homan/losses.py:20-49,98-139,199-242),
 * compute_pca_loss (homan/lossutils.py:39-40), evaluated per image b = p*T + t with per-problem normalisers.
 * partials [B, HM_NPART] receives the unweighted per-image contributions (summed per problem by
 * hm_finalize_losses); gradients are weighted by `w` and accumulated (+=).
 * w = {smooth_hand, smooth_obj, v2d_hand, inter, pca} ; weight 0 with loss_on 0 skips the term. */
#define HM_PART_SMOOTH_HAND 0
#define HM_PART_SMOOTH_OBJ 1
#define HM_PART_V2D 2
#define HM_PART_V2D_PX 3      /* metric: sum of pixel distances / 778 */
#define HM_PART_INTER 4
#define HM_PART_PCA 5
#define HM_PART_SIL_OBJ 6
#define HM_PART_IOU_OBJ 7     /* metric */
#define HM_PART_SIL_HAND 8
#define HM_PART_IOU_HAND 9    /* metric */
#define HM_PART_COLLISION 10
#define HM_PART_CONTACT 11
#define HM_PART_MINDIST 12    /* metric: min hand-object vertex distance of the image */
#define HM_PART_INTER_FLAG 13
#define HM_PART_OFFSCREEN 14  /* object-pose initialiser only (hm_offscreen_loss_fwd_bwd) */
#define HM_NPART 16
int hm_vertex_losses(const float *verts_hand, const float *verts_obj, const float *camintr,
                     const float *ref_verts2d, const float *pca, int pca_dim, int B, int T, int Vo,
                     float image_size, float w_smooth_hand, float w_smooth_obj, float w_v2d, float w_inter,
                     float w_pca, int flags, float *partials, float *grad_verts_hand, float *grad_verts_obj,
                     float *grad_centroid_det, float *grad_pca, void *stream);
#define HM_VL_SMOOTH 1
#define HM_VL_V2D 2
#define HM_VL_INTER 4
#define HM_VL_PCA 8

/* compute_contact_loss, default-argument path (homan/interactions/contactloss.py:149-309): nearest object
 * vertex of every hand vertex (first minimum of |h|^2 + |o|^2 - 2 h.o), loss = mean 0.02 tanh(d / 0.02);
 * gradient to both meshes, weighted by `weight / (T * 778)`.  partials as above (CONTACT, MINDIST). */
int hm_contact_fwd_bwd(const float *verts_hand, const float *verts_obj, int B, int T, int Vo, float thresh,
                       float weight, float *partials, float *grad_verts_hand, float *grad_verts_obj,
                       unsigned long long *grad_fixed_obj, void *stream);

/* Nearest point of b [B,M,3] for every point of a [B,N,3]: squared distance [B,N] and (optionally) index [B,N] (ties:
 * lowest index).  The search behind the evaluation's point metrics (homan/eval/pointmetrics.py:17-45,95-99:
 * pytorch3d chamfer_distance and scipy cKDTree.query(k=1) for ADD-S), which the reference runs on the fitted meshes. */
int hm_nearest_point(const float *a, const float *b, int B, int N, int M, float *dist2, int32_t *index, void *stream);

/* ---------------------------------------------------------------- SDF interpenetration
 * SDFSceneLoss.forward for (hand, object) (homan/interactions/scenesdf.py:77-148, called from
 * compute_collision_loss, homan/lossutils.py:43-64): phi = clamp(SDF, 0) of the grid mesh in its own
 * normalised bbox cube, sampled trilinearly (grid_sample, zeros padding, align_corners=False) at the other
 * mesh's vertices.  Evaluated sparsely: only the voxels that samples touch.  One call = one ordered pair:
 * grid mesh (verts_g [B,Vg,3], faces_g [faces_batch,Fg,3] with faces_batch 1 or B) sampled at verts_s [B,Vs,3].
 * partials[b*HM_NPART + HM_PART_COLLISION] += sum of samples; grad_verts_s += weight * d/d verts_s (may be NULL);
 * dist_values [B,Vs] (may be NULL) = the samples in scene units, the `dist_values[(g, s)]` of the reference, which its
 * penetration metric reads (homan/eval/pointmetrics.py:102-124).
 * workspace: phi [B, G^3] fp32 scratch (written sparsely). */
int hm_sdf_pair(const float *verts_g, const int32_t *faces_g, int faces_batch, const float *verts_s, int B, int Vg,
                int Fg, int Vs, int grid, float scale_factor, float weight, float *phi_scratch, float *partials,
                float *grad_verts_s, float *dist_values, void *stream);
/* sdf.SDF()(faces, vertices) (un-vendored `sdf` package; homan/interactions/scenesdf.py:32,119): dense
 * signed distance grid phi [B,G,G,G] (inside positive) of vertices already normalised to [-1,1]^3. */
int hm_sdf_grid(const int32_t *faces, const float *verts, int B, int V, int F, int grid, float *phi,
                void *stream);

/* ---------------------------------------------------------------- per-problem reduction, Adam, argmin
 * partials [B,HM_NPART] -> losses [P,HM_NPART] (sums over the T frames of each problem; metrics: mean,
 * MINDIST -> max) and total[P] = sum_k w[k] * losses[p,k]; also advances the device step counter. */
int hm_finalize_losses(const float *partials, const float *weights_part, int P, int T, float *losses,
                       float *total, int *step_counter, void *stream);
/* torch.optim.Adam step (homan/jointopt.py:138-151,192) on one flat parameter buffer:
 * lr_per_elem [n] (0 = frozen parameter, e.g. mano_rot / mano_trans which match no optimiser group). */
int hm_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                 const float *lr_per_elem, int n, float beta1, float beta2, float eps,
                 const int *step_counter, void *stream);
/* best init of every clip: total [C, I] -> best_index [C], best_loss [C] (first minimum). */
int hm_argmin_over_inits(const float *total, int C, int I, int32_t *best_index, float *best_loss,
                         void *stream);

/* ---------------------------------------------------------------- object-pose multi-init fitter
 * PoseOptimizer.compute_offscreen_loss (homan/pose_optimization.py:112-134) on projected vertices
 * ndc [B,V,3] = nr.projection output: partials[b*HM_NPART + HM_PART_OFFSCREEN] = sum over vertices of
 * relu(u-1) + relu(v-1) + relu(-1-u) + relu(-1-v) + relu(-z) + relu(z-far); grad_ndc += weight * d/d ndc
 * (may be NULL). The reference weights the term by 100000 (pose_optimization.py:148). */
int hm_offscreen_loss_fwd_bwd(const float *ndc, int B, int V, float far_, float weight, float *partials,
                              float *grad_ndc, void *stream);
/* Best candidate ever of find_optimal_pose (homan/pose_optimization.py:349-353): if min_i total[i] < best[0],
 * best = {that loss, rot6d[i] (6), trans[i] (3)} with the CURRENT parameters (the reference reads them after
 * optimizer.step()); best_index (may be NULL) receives i. best[0] starts at +inf. */
int hm_track_best(const float *total, int N, const float *rot6d, const float *trans, float *best,
                  int32_t *best_index, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HOMAN_B200_H */
