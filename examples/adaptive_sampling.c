/* examples/adaptive_sampling.rs of the reference (BASELINE config 1), as a plain-C host of the C ABI.
 *
 *   f64::random_approximate_low_rank_matrix((500, 200), 1.0, 1e-10)        examples/adaptive_sampling.rs:24
 *   mat.sample_range_adaptive(1e-5, 5, &mut rng)                           :30
 *   exact residual ||A - Q_i Q_i^T A|| / ||A|| for i = 0, 5, 10, ...        :63-72  (the red curve of residuals.png)
 *   QR::compute_from_range_estimate(q, &mat); rel_diff_fro(to_mat, mat)    :80-84
 *
 * The plot itself (plotters, :36-75) is out of scope; the two curves are printed as a table.
 * Build: gcc -std=c99 -Iinclude examples/adaptive_sampling.c -Lrusty_compression_b200 -lrc_b200 -lm
 * No CPU fallback: without a B200 rc_ctx_create fails and the program exits non-zero. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rc_api.h"

#define CHECK(call)                                                                              \
    do {                                                                                         \
        rc_status st_ = (call);                                                                  \
        if (st_ != RC_OK) {                                                                      \
            fprintf(stderr, "%s failed with status %d: %s\n", #call, (int)st_,                   \
                    ctx ? rc_last_error_string(ctx) : "(no context)");                           \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

int main(int argc, char** argv) {
    rc_ctx* ctx = NULL;
    const int64_t m = 500, n = 200, sample_size = 5;
    const double rel_tol = 1e-5;
    const uint64_t seed = (argc > 1) ? strtoull(argv[1], NULL, 10) : 0;
    CHECK(rc_ctx_create(0, &ctx));

    rc_matrix* mat = NULL;
    CHECK(rc_random_approximate_low_rank_matrix(ctx, RC_F64, m, n, 1.0, 1e-10, seed, &mat));

    rc_matrix* q = NULL;
    uint64_t hist_rank[1024];
    double hist_res[1024];
    size_t hist_len = 0;
    CHECK(rc_sample_range_adaptive(ctx, mat, rel_tol, sample_size, NULL, seed + 1, 0, &q, hist_rank, hist_res, 1024,
                                   &hist_len));
    const int64_t rank = rc_matrix_cols(q);

    /* exact residuals: host copy of q, leading column blocks re-uploaded as strided views */
    double* qh = (double*)malloc(sizeof(double) * (size_t)m * (size_t)rank);
    double* bh = (double*)malloc(sizeof(double) * (size_t)n * (size_t)rank);
    if (!qh || !bh) return 1;
    CHECK(rc_matrix_to_host(ctx, q, qh));
    printf("%8s %22s %22s\n", "rank", "estimated residual", "exact residual");
    for (size_t i = 0; i < hist_len; ++i) {
        const int64_t r = (int64_t)hist_rank[i];
        rc_matrix *qi = NULL, *b = NULL, *bt = NULL, *proj = NULL;
        double exact = 0.0;
        CHECK(rc_matrix_from_host(ctx, RC_F64, qh, m, r, rank, 1, &qi));     /* q[:, 0..r] as a strided view */
        CHECK(rc_conj_matmat(ctx, mat, qi, &b));                             /* b = A^T Q_r        (n x r) */
        CHECK(rc_matrix_to_host(ctx, b, bh));
        CHECK(rc_matrix_from_host(ctx, RC_F64, bh, r, n, 1, r, &bt));        /* b^T = Q_r^T A      (r x n), transposed view */
        CHECK(rc_matmat(ctx, qi, bt, &proj));                                /* Q_r Q_r^T A        (m x n) */
        CHECK(rc_rel_diff_fro(ctx, proj, mat, &exact));                      /* ||Q Q^T A - A|| / ||A|| */
        printf("%8lld %22.6e %22.6e\n", (long long)r, hist_res[i], exact);
        rc_matrix_free(qi); rc_matrix_free(b); rc_matrix_free(bt); rc_matrix_free(proj);
    }
    printf("Rank: %lld\n", (long long)rank);

    rc_qr* qr = NULL;
    rc_matrix* approx = NULL;
    double rel_diff = 0.0;
    CHECK(rc_qr_compute_from_range_estimate(ctx, q, mat, &qr));
    CHECK(rc_qr_to_mat(ctx, qr, &approx));
    CHECK(rc_rel_diff_fro(ctx, approx, mat, &rel_diff));
    printf("The relative difference of the compressed and original matrix is %1.2E\n", rel_diff);

    int64_t launches = 0;
    CHECK(rc_ctx_get_counter(ctx, "kernel_launches", &launches));
    printf("kernel launches: %lld\n", (long long)launches);

    free(qh);
    free(bh);
    rc_matrix_free(approx); rc_qr_free(qr); rc_matrix_free(q); rc_matrix_free(mat);
    rc_ctx_destroy(ctx);
    return (rel_diff < 10.0 * rel_tol && launches > 0) ? 0 : 2;
}
