/* mini_cv_ops.h — the C callback table behind mini_cv (TEST INFRASTRUCTURE ONLY).
 * oracle/ref_lib.py registers one callback per OpenCV function; each is implemented with the cv2 4.13 wheel,
 * i.e. with the real OpenCV library.  Every callback returns 0 on success, non-zero when cv2 raised. */
#ifndef MINI_CV_OPS_H
#define MINI_CV_OPS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mini_cv_mat {
    unsigned char *data;
    int rows, cols, type; /* OpenCV type code: depth + ((cn-1) << 3) */
    size_t step;
} mini_cv_mat;

typedef struct mini_cv_ops {
    int (*resize)(const mini_cv_mat *src, mini_cv_mat *dst, int interp);
    int (*cvt_color)(const mini_cv_mat *src, mini_cv_mat *dst, int code);
    int (*gftt)(const mini_cv_mat *img, const mini_cv_mat *mask, int max_corners, double quality, double min_dist, int block, int harris,
                double k, float *xy, int cap, int *n);
    int (*pyr_lk)(const mini_cv_mat *prev, const mini_cv_mat *next, const float *prev_xy, int n, float *next_xy, unsigned char *status,
                  float *err, int win_w, int win_h, int max_level, int crit_type, int crit_count, double crit_eps, int flags,
                  double min_eig);
    int (*estimate_affine_partial)(const float *from_xy, const float *to_xy, int n, int method, double thresh, int max_iters,
                                   double confidence, int refine_iters, double *m6, unsigned char *mask, int *ok);
    int (*warp_affine)(const mini_cv_mat *src, mini_cv_mat *dst, const double *m6, int m_is_f32, int flags, int border_mode,
                       const double *border_value);
    int (*copy_make_border)(const mini_cv_mat *src, mini_cv_mat *dst, int top, int bottom, int left, int right, int border_type,
                            const double *value);
    int (*add_weighted)(const mini_cv_mat *a, double alpha, const mini_cv_mat *b, double beta, double gamma, mini_cv_mat *dst);
    int (*threshold)(const mini_cv_mat *src, mini_cv_mat *dst, double thresh, double maxval, int type);
    int (*find_contours)(const mini_cv_mat *img, int mode, int method, int *pts_xy, int pts_cap, int *lens, int lens_cap, int *ncont);
    int (*kalman_create)(int dp, int mp, int cp, int *handle);
    int (*kalman_predict)(int handle, mini_cv_mat *state9);
    int (*kalman_correct)(int handle, const mini_cv_mat *measurement, mini_cv_mat *state9);
    int (*kalman_release)(int handle);
    /* RollCorrection / AutoZoomCrop (cv::cuda:: in the reference; the CPU equivalents of the same OpenCV build) */
    int (*canny)(const mini_cv_mat *src, mini_cv_mat *dst, double low, double high, int aperture, int l2);
    int (*hough_lines)(const mini_cv_mat *edges, double rho, double theta, int threshold, int max_lines, float *rho_theta, int *votes, int cap, int *n);
    int (*gaussian_blur)(const mini_cv_mat *src, mini_cv_mat *dst, int kw, int kh, double sx, double sy);
    int (*remap)(const mini_cv_mat *src, mini_cv_mat *dst, const mini_cv_mat *mapx, const mini_cv_mat *mapy, int interp, int border_mode);
    int (*morphology)(const mini_cv_mat *src, mini_cv_mat *dst, int op, const mini_cv_mat *kernel);
    int (*structuring_element)(int shape, mini_cv_mat *dst);
    int (*draw_contours)(mini_cv_mat *img, const int *pts_xy, const int *lens, int ncont, int idx, const double *color, int thickness);
    int (*rotation_matrix)(double cx, double cy, double angle, double scale, double *m6);
    int (*sobel)(const mini_cv_mat *src, mini_cv_mat *dst, int dx, int dy, int ksize);
} mini_cv_ops;

void mini_cv_set_ops(const mini_cv_ops *ops);
const mini_cv_ops *mini_cv_get_ops(void);

#ifdef __cplusplus
}
#endif
#endif
