/* examples/interpolative_decomposition.rs of the reference, as a plain-C host of the C ABI.
 *
 *   f64::random_approximate_low_rank_matrix((500, 100), 1.0, 1e-10)        examples/interpolative_decomposition.rs:22
 *   QR::compute_from(mat)                                                  :25
 *   qr.compress(CompressionType::RANK(20))                                 :29
 *   column_id() -> two_sided_id() -> to_mat() -> rel_diff_fro              :32-46
 *
 * Build: gcc -std=c99 -Iinclude examples/interpolative_decomposition.c -Lrusty_compression_b200 -lrc_b200 -lm */
#include <stdio.h>
#include <stdlib.h>
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
    const int64_t m = 500, n = 100, k = 20;
    const uint64_t seed = (argc > 1) ? strtoull(argv[1], NULL, 10) : 0;
    CHECK(rc_ctx_create(0, &ctx));

    rc_matrix* mat = NULL;
    CHECK(rc_random_approximate_low_rank_matrix(ctx, RC_F64, m, n, 1.0, 1e-10, seed, &mat));

    rc_qr *qr = NULL, *qr_compressed = NULL;
    rc_column_id* col_int_decomp = NULL;
    rc_two_sided_id* two_sided_int_decomp = NULL;
    rc_matrix* mat_approx = NULL;
    double rel_diff = 0.0;
    CHECK(rc_qr_compute_from(ctx, mat, &qr));
    CHECK(rc_qr_compress_rank(ctx, qr, k, &qr_compressed));
    CHECK(rc_qr_column_id(ctx, qr_compressed, &col_int_decomp));
    CHECK(rc_column_id_two_sided_id(ctx, col_int_decomp, &two_sided_int_decomp));
    CHECK(rc_two_sided_id_to_mat(ctx, two_sided_int_decomp, &mat_approx));
    CHECK(rc_rel_diff_fro(ctx, mat, mat_approx, &rel_diff));
    printf("The relative difference of the compressed and original matrix is %1.2E\n", rel_diff);

    /* the skeleton: first k entries of the full-length index vectors (quirk Q8) */
    uint64_t col_ind[100], row_ind[500];
    CHECK(rc_two_sided_id_get_col_ind(two_sided_int_decomp, col_ind, (size_t)n));
    CHECK(rc_two_sided_id_get_row_ind(two_sided_int_decomp, row_ind, (size_t)m));
    printf("skeleton columns:");
    for (int i = 0; i < k; ++i) printf(" %llu", (unsigned long long)col_ind[i]);
    printf("\nskeleton rows:");
    for (int i = 0; i < k; ++i) printf(" %llu", (unsigned long long)row_ind[i]);
    printf("\n");

    rc_matrix_free(mat_approx);
    rc_two_sided_id_free(two_sided_int_decomp);
    rc_column_id_free(col_int_decomp);
    rc_qr_free(qr_compressed);
    rc_qr_free(qr);
    rc_matrix_free(mat);
    rc_ctx_destroy(ctx);
    /* sigma_20 / sigma_0 of the geometric spectrum is 10^(-10 * 20 / 99) ~ 1e-2 */
    return (rel_diff > 0.0 && rel_diff < 0.2) ? 0 : 2;
}
