/* TEST INFRASTRUCTURE ONLY -- the CPU oracle.  Never linked, imported or called by
 * the product path (slam-toolkit_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, and only as the
 * checker / reported baseline.
 *
 * Dependency-free C restatement of the reference's ORB front end:
 *   src/orb_extractor.cpp:72-147,410-853,1034-1132, include/orb_extractor.h:87-103,
 *   src/matcher.cpp:54-209, src/camera.cpp:26-36,50-79 (paths in geonuklee/slam-toolkit).
 * The OpenCV 3.4 primitives the reference delegates to (cv::FAST, cv::resize,
 * cv::GaussianBlur, cv::fastAtan2, cvRound) are replaced by closed-form integer /
 * float32 models that are pinned bit-for-bit against cv2 4.13 by
 * tests/test_oracle_vs_cv2.py and the fixtures in tests/golden/.
 *
 * Parity status: the reference ships no tests / golden vectors, and its own
 * keypoint order depends on heap addresses (src/orb_extractor.cpp:684), so parity
 * is defined against THIS canonical oracle with the declared tie rules T1-T4
 * (SURVEY.md §8c).  "Parity unpinned by the reference's own tests."
 */
#ifndef ORB_ORACLE_H_
#define ORB_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* byte-identical to cv::KeyPoint (28 B) */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orc_keypoint;

typedef struct orc_extractor orc_extractor;

orc_extractor *orc_extractor_create(int nfeatures, float scale_factor, int nlevels,
                                    int ini_th_fast, int min_th_fast);
void orc_extractor_destroy(orc_extractor *ex);
/* tables: out arrays of nlevels entries each (any may be NULL) */
void orc_extractor_tables(const orc_extractor *ex, float *scale, float *inv_scale, float *sigma2,
                          float *inv_sigma2, int *features_per_level, int *umax16);
void orc_level_size(const orc_extractor *ex, int w, int h, int level, int *lw, int *lh);

/* full extraction; returns keypoint count (<= cap) or <0 on error */
int orc_extract(orc_extractor *ex, const uint8_t *img, int w, int h, int stride,
                orc_keypoint *kps, uint8_t *desc, int cap);

/* stage taps, valid after orc_extract on the same handle */
int orc_get_level(const orc_extractor *ex, int level, uint8_t *out /* lw*lh */);
int orc_get_blur(const orc_extractor *ex, int level, uint8_t *out /* lw*lh */);
int orc_get_score(const orc_extractor *ex, int level, uint8_t *out /* lw*lh, FAST best map */);
int orc_get_candidates(const orc_extractor *ex, int level, float *xyr /* cap*3 */, int cap);
int orc_get_distributed(const orc_extractor *ex, int level, float *xyr /* cap*3 */, int cap);

/* 1 when level's quadtree stopped between two equally-full nodes (the reference's heap-address order decides, T1) */
int orc_get_tie_cut(const orc_extractor *ex, int level);
int orc_distribute_ex(const float *xyr, int n, int min_x, int max_x, int min_y, int max_y,
                      int n_want, float *out_xyr, int cap, int *tie_cut);

/* primitive models (pinned to cv2) */
void orc_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride,
                          uint8_t *dst, int dw, int dh, int dstride);
void orc_gaussian7_s2_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride);
void orc_fast_score_u8(const uint8_t *img, int w, int h, int stride, uint8_t *score /* w*h */);
int orc_fast_nms(const uint8_t *img, int w, int h, int stride, int threshold,
                 float *xyr /* cap*3 */, int cap);
float orc_fast_atan2(float y, float x);
int orc_distribute(const float *xyr, int n, int min_x, int max_x, int min_y, int max_y,
                   int n_want, float *out_xyr, int cap);

/* matching */
int orc_hamming256(const void *a, const void *b);
void orc_stereo_match(const orc_keypoint *kl, const uint8_t *dl, int nl,
                      const orc_keypoint *kr, const uint8_t *dr, int nr,
                      double y_thr, double max_dx, double ratio, int *out_idx, int *out_dist);
typedef struct {
    double fx, fy, cx, cy;
    double d[4];
    int width, height;
} orc_camera;
/* kp_to_query[m] = winning map-point index or -1; kp_dist[m] = its distance */
void orc_projection_match(const double *xw, const uint8_t *mp_desc, const uint8_t *skip, int n,
                          const double rt[12], const orc_camera *cam,
                          const orc_keypoint *kps, const uint8_t *kp_desc, int m, double radius,
                          double ratio, int *kp_to_query, int *kp_dist);
/* the same with a 32-px bucket grid over the keypoints as candidate accelerator (CPU-baseline stand-in for FLANN) */
void orc_projection_match_grid(const double *xw, const uint8_t *mp_desc, const uint8_t *skip, int n,
                               const double rt[12], const orc_camera *cam,
                               const orc_keypoint *kps, const uint8_t *kp_desc, int m, double radius,
                               double ratio, int *kp_to_query, int *kp_dist);
/* The same three with the pose as the reference holds it (g2o::SE3Quat, src/matcher.cpp:151): qt = {qx, qy, qz, qw, tx, ty, tz},
 * Xc = t + q * Xw with Eigen's quaternion-vector product.  This is the form pinned bit-for-bit to oracle/_ref. */
void orc_se3_apply(const double qt[7], const double *x, int n, double *out);
void orc_projection_match_se3(const double *xw, const uint8_t *mp_desc, const uint8_t *skip, int n,
                              const double qt[7], const orc_camera *cam,
                              const orc_keypoint *kps, const uint8_t *kp_desc, int m, double radius,
                              double ratio, int *kp_to_query, int *kp_dist);
void orc_projection_match_grid_se3(const double *xw, const uint8_t *mp_desc, const uint8_t *skip, int n,
                                   const double qt[7], const orc_camera *cam,
                                   const orc_keypoint *kps, const uint8_t *kp_desc, int m, double radius,
                                   double ratio, int *kp_to_query, int *kp_dist);
/* out[q*4] = {idx0, dist0, idx1, dist1}; lexicographic (dist, idx) top-2 */
void orc_knn2(const uint8_t *queries, int q, const uint8_t *db, int64_t m, int64_t idx_base,
              int32_t *out);
/* the same with the queries split over nthreads worker threads (database walked in cache-sized tiles) */
void orc_knn2_mt(const uint8_t *queries, int q, const uint8_t *db, int64_t m, int64_t idx_base,
                 int32_t *out, int nthreads);

/* frame glue: src/camera.cpp:95-109, src/frame.cpp:50-56,157-193,391-409.
 * T6 (canonical): SearchNeareast resolves equal distances towards the smaller keypoint index; SearchRadius lists
 * indices in ascending order (FLANN's own order is by distance and irrelevant to every caller). */
void orc_normalized_undistort(const orc_camera *cam, const orc_keypoint *kps, int n, double *xy);
void orc_stereo_depth(const orc_camera *cam, double baseline, const orc_keypoint *kps_l, const double *norm_xy, int n,
                      const orc_keypoint *kps_r, const int *stereo_idx, double *xc, uint8_t *valid);
void orc_reprojection_error(const orc_camera *cam, const double rt[12], const orc_keypoint *kps, int n, const double *xw,
                            const uint8_t *has_mp, double *err);
void orc_reprojection_error_se3(const orc_camera *cam, const double qt[7], const orc_keypoint *kps, int n, const double *xw,
                                const uint8_t *has_mp, double *err);
int orc_search_radius(const orc_keypoint *kps, int m, double u, double v, double radius, int *idx, int cap);
void orc_search_nearest(const orc_keypoint *kps, int m, double u, double v, int *kpt_index, double *dist2);

/* BoW transform, thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1259; T7: nid = 0 when the descent
 * ends above level L - levelsup (the reference leaves it unwritten). */
void orc_vocab_transform(int n_nodes, const int32_t *parent, const uint8_t *is_leaf, const uint8_t *node_desc,
                         const double *node_weight, int L, const uint8_t *features, int n, int levelsup,
                         int32_t *word_id, double *weight, int32_t *node_id);

/* CPU baseline helper: extract L + extract R + stereo match for `count` stereo
 * frames with `nthreads` worker threads (one frame per thread at a time).
 * left/right: count contiguous w*h images.  Returns total matches. */
int64_t orc_stereo_frames(const uint8_t *left, const uint8_t *right, int count, int w, int h,
                          int nthreads, int nfeatures, float scale_factor, int nlevels,
                          int ini_th, int min_th, int64_t *total_kps);

/* tracking step between consecutive stereo frames (src/frame.cpp:391-409 + src/matcher.cpp:134-209): GetDepth of the
 * previous frame's stereo keypoints, ProjectionMatch into the current frame; track_idx[j] = previous keypoint or -1 */
void orc_track_pair(const orc_camera *cam, double baseline, const double rt[12], double radius, double ratio,
                    const orc_keypoint *kl_prev, const uint8_t *dl_prev, int nl_prev, const orc_keypoint *kr_prev,
                    const int *sidx_prev, const orc_keypoint *kps, const uint8_t *desc, int n, int *track_idx,
                    int *track_dist, int use_grid);
/* CPU baseline helper for the sequence workload: orc_stereo_frames + orc_track_pair over consecutive frames */
int64_t orc_stereo_sequence(const uint8_t *left, const uint8_t *right, int count, int w, int h, int nthreads, int nfeatures,
                            float scale_factor, int nlevels, int ini_th, int min_th, const orc_camera *cam, double baseline,
                            double radius, int64_t *total_kps, int64_t *total_tracked);

#ifdef __cplusplus
}
#endif
#endif
