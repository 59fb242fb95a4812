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
#define HM_FACE_RECORD_BYTES 64
#define HM_FACE_BBOX_BYTES 8
/* ndc [B,V,3], faces [faces_batch,F,3] -> records [B,F,64 B] + bboxes [B,F,8 B] (front-facing
 * winding of every face; never materialises the doubled face array). */
int hm_raster_setup(const float *ndc, const int32_t *faces, int faces_batch, int B, int V, int F,
                    int image_size, int anti_aliasing, int fill_back, void *records, void *bboxes,
                    void *stream);
/* -> face_index [B,S,S] (doubled numbering: f or F+f, -1 background; raster frame, row 0 = y -1),
 *    alpha [B,R,R] (after flip + pool), coverage bitmaps cov_row / cov_col [B,S,S/32] (may be NULL). */
int hm_raster_sil_fwd(const void *records, const void *bboxes, int B, int F, int image_size,
                      int anti_aliasing, float near_, float far_, int32_t *face_index, float *alpha,
                      uint32_t *cov_row, uint32_t *cov_col, void *stream);
/* grad_alpha [B,R,R] + coverage -> sweep masks m_row / m_col [B,2,S,S/32]
 * (0: uncovered & grad<0, 1: covered & grad>0). */
int hm_raster_grad_prep(const float *grad_alpha, const uint32_t *cov_row, const uint32_t *cov_col, int B,
                        int image_size, int anti_aliasing, uint32_t *m_row, uint32_t *m_col, void *stream);
/* backward_pixel_map fused with the vertices_to_faces scatter-add:
 * grad_ndc [B,V,3] += d loss / d (u, v) of every vertex (z receives nothing in silhouette mode). */
int hm_raster_sil_bwd(const void *records, const void *bboxes, const int32_t *face_index,
                      const float *grad_alpha, const uint32_t *cov_row, const uint32_t *cov_col,
                      const uint32_t *m_row, const uint32_t *m_col, int B, int V, int F, int image_size,
                      int anti_aliasing, float eps, float *grad_ndc, void *stream);

/* Losses.compute_sil_loss_object (homan/losses.py:183-197) on a rendered alpha:
 * target int8 [B,R,R] in {-1 occluded, 0, 1}; norm [B] = 1 / (sum keep * T) of the image's problem;
 * loss_img[b*loss_stride] = norm * sum (keep*alpha - ref)^2 ; iou_img[b*iou_stride] (mask IoU, metric);
 * grad_alpha = weight * 2 * norm * keep * (keep*alpha - ref). */
int hm_sil_loss_fwd_bwd(const float *alpha, const int8_t *target, const float *norm, float weight, int B,
                        int image_size, float *loss_img, int loss_stride, float *iou_img, int iou_stride,
                        float *grad_alpha, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HOMAN_B200_H */
