// rusty_compression_b200.hpp -- C++17 host-side mirror of the rusty-compression public surface over the C ABI.
//
// The reference is a Rust crate and this image has no Rust toolchain, so the compiled-language host side above
// include/rc_api.h is this header (the Rust equivalent is shipped source-only under rust/).  It mirrors the
// crate's types and methods name for name, with the same argument meaning and error behaviour:
//
//   reference (file:line)                                              here
//   ------------------------------------------------------------------ ------------------------------------------
//   RustyCompressionError {LinalgError, CompressionError, LayoutError,  rcb200::Error + LinalgError, CompressionError,
//     PivotedQRError}; assert! panics (src/types.rs:11-23)                LayoutError, PivotedQRError; InvalidArgument
//   CompressionType::{ADAPTIVE(tol), RANK(k)} (src/lib.rs:82-87)        rcb200::CompressionType::ADAPTIVE / RANK
//   MatVec / MatMat / ConjMatVec / ConjMatMat (src/types.rs:40-101)     Matrix<A>::matvec / matmat / conj_matvec /
//                                                                          conj_matmat; Operator<A> for matrix-free
//   SampleRange / SampleRangePowerIteration / AdaptiveSampling          sample_range_by_rank / _power_iteration /
//     (src/random_sampling.rs:58-98, 202-218)                             _adaptive (free functions over any operator)
//   MaxColNorm (src/random_sampling.rs:175-199), RelDiff (types.rs:162) max_col_norm, rel_diff_fro, rel_diff_l2
//   QR / QRTraits, LQ / LQTraits (src/qr.rs:31-237)                     QR<A>, LQ<A>
//   SVD / SVDTraits (src/svd.rs:13-122)                                 SVD<A>
//   ColumnID, RowID, TwoSidedID + *Traits + Apply                       ColumnID<A>, RowID<A>, TwoSidedID<A> (dot = Apply)
//   invert_permutation_vector, ApplyPermutationTo{Matrix,Vector}        invert_permutation_vector, apply_permutation
//     (src/permutation.rs:7-75)
//   RandomMatrix (src/random_matrix.rs:21-93)                           Matrix<A>::random_gaussian / random_orthogonal_matrix /
//                                                                          random_approximate_low_rank_matrix
//
// Differences that follow from the device: matrices live in HBM (`Matrix<A>` is an owning handle; `from_host` /
// `to_host` are the only transfers), the RNG argument `&mut R` becomes a 64-bit seed (or an ingested Omega for
// parity runs), and results are returned by value as handles.  Index vectors are std::vector<size_t>, always of
// full length (quirk Q8).  There is no CPU fallback: every call ends in librc_b200.so.
#ifndef RUSTY_COMPRESSION_B200_HPP
#define RUSTY_COMPRESSION_B200_HPP

#include <complex>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "rc_api.h"

namespace rcb200 {

using c32 = std::complex<float>;    // src/types.rs:9 (num::Complex is #[repr(C)] (re, im), like std::complex)
using c64 = std::complex<double>;

// ---------------------------------------------------------------------------------------------- errors
struct Error : std::runtime_error {
    rc_status status;
    Error(rc_status st, const std::string& what) : std::runtime_error(what), status(st) {}
};
struct LinalgError : Error { using Error::Error; };         // RustyCompressionError::LinalgError
struct CompressionError : Error { using Error::Error; };    // ::CompressionError (tolerance not reached, src/qr.rs:196-199)
struct LayoutError : Error { using Error::Error; };         // ::LayoutError
struct PivotedQRError : Error { using Error::Error; };      // ::PivotedQRError
struct InvalidArgument : Error { using Error::Error; };     // where the crate panics through assert!

template <class A> struct ScalarTraits;
template <> struct ScalarTraits<float> { static constexpr rc_dtype dtype = RC_F32; using Real = float; };
template <> struct ScalarTraits<double> { static constexpr rc_dtype dtype = RC_F64; using Real = double; };
template <> struct ScalarTraits<c32> { static constexpr rc_dtype dtype = RC_C32; using Real = float; };
template <> struct ScalarTraits<c64> { static constexpr rc_dtype dtype = RC_C64; using Real = double; };

// ---------------------------------------------------------------------------------------------- context
class Context {
  public:
    Context(std::nullptr_t) {}                  // empty (no device context); what default-constructed handles hold
    explicit Context(int device = 0) {
        rc_ctx* raw = nullptr;
        rc_status st = rc_ctx_create(device, &raw);
        if (st != RC_OK) throw Error(st, "rc_ctx_create failed (no B200 visible? there is no CPU fallback)");
        ctx_ = std::shared_ptr<rc_ctx>(raw, [](rc_ctx* c) { rc_ctx_destroy(c); });
    }
    rc_ctx* raw() const { return ctx_.get(); }
    void check(rc_status st) const {
        if (st == RC_OK) return;
        const std::string msg = rc_last_error_string(ctx_.get());
        switch (st) {
            case RC_LINALG_ERROR: throw LinalgError(st, msg);
            case RC_COMPRESSION_ERROR: throw CompressionError(st, msg);
            case RC_LAYOUT_ERROR: throw LayoutError(st, msg);
            case RC_PIVOTED_QR_ERROR: throw PivotedQRError(st, msg);
            case RC_INVALID_ARGUMENT: throw InvalidArgument(st, msg);
            default: throw Error(st, msg);
        }
    }
    void synchronize() const { check(rc_ctx_synchronize(ctx_.get())); }
    // page-lock / unlock host memory the caller owns (full-rate, truly asynchronous uploads from it)
    void pin(void* host, size_t bytes) const { check(rc_host_register(ctx_.get(), host, bytes)); }
    void unpin(void* host) const { check(rc_host_unregister(ctx_.get(), host)); }
    void set_option(const char* key, int64_t value) const { check(rc_ctx_set_option(ctx_.get(), key, value)); }
    int64_t counter(const char* key) const {
        int64_t v = 0;
        check(rc_ctx_get_counter(ctx_.get(), key, &v));
        return v;
    }

  private:
    std::shared_ptr<rc_ctx> ctx_;
};

// CompressionType (src/lib.rs:82-87)
struct CompressionType {
    enum Kind { ADAPTIVE_, RANK_ } kind;
    double tol;
    size_t rank;
    static CompressionType ADAPTIVE(double tol) { return {ADAPTIVE_, tol, 0}; }
    static CompressionType RANK(size_t k) { return {RANK_, 0.0, k}; }
};

enum class MatrixPermutationMode { COL = RC_PERM_COL, ROW = RC_PERM_ROW, COLINV = RC_PERM_COLINV, ROWINV = RC_PERM_ROWINV };
enum class VectorPermutationMode { INV = RC_VPERM_INV, NOINV = RC_VPERM_NOINV };

namespace detail {
inline std::vector<size_t> to_usize(const std::vector<uint64_t>& v) { return std::vector<size_t>(v.begin(), v.end()); }
inline std::vector<uint64_t> to_u64(const std::vector<size_t>& v) { return std::vector<uint64_t>(v.begin(), v.end()); }
}  // namespace detail

// ---------------------------------------------------------------------------------------------- Matrix<A>
// Device-resident Array2<A>; for A itself also the operator (MatVec / MatMat / ConjMatVec / ConjMatMat,
// src/types.rs:40-133).  `owned == false` wraps a handle that belongs to a decomposition (get_q() ...).
// Such a borrowed view must not outlive the decomposition it came from.
template <class A>
class Matrix {
  public:
    using Real = typename ScalarTraits<A>::Real;
    Matrix() = default;
    Matrix(Context ctx, rc_matrix* h, bool owned = true) : ctx_(std::move(ctx)), h_(h, Deleter{owned}) {}

    // Upload a (possibly strided) host view, strides in elements (ndarray views, src/pivoted_qr.rs:25-31).
    static Matrix from_host(const Context& ctx, const A* data, size_t rows, size_t cols, ptrdiff_t row_stride, ptrdiff_t col_stride) {
        rc_matrix* h = nullptr;
        ctx.check(rc_matrix_from_host(ctx.raw(), ScalarTraits<A>::dtype, data, (int64_t)rows, (int64_t)cols,
                                      (int64_t)row_stride, (int64_t)col_stride, &h));
        return Matrix(ctx, h);
    }
    static Matrix from_host(const Context& ctx, const std::vector<A>& row_major, size_t rows, size_t cols) {
        if (row_major.size() != rows * cols) throw InvalidArgument(RC_INVALID_ARGUMENT, "from_host: size mismatch");
        return from_host(ctx, row_major.data(), rows, cols, (ptrdiff_t)cols, 1);
    }
    // Pipelined upload of a row-major host buffer (pinned memory for a truly asynchronous copy): returns at once, the
    // copy runs on the context's copy stream under whatever the context stream is computing.  The result must go through
    // await_upload() before any other use; `data` stays valid and unmodified until await_upload(true) returns (or any
    // later blocking call on the context).  No counterpart in the reference (a host library has no transfer to hide).
    static Matrix from_host_async(const Context& ctx, const A* data, size_t rows, size_t cols, ptrdiff_t row_stride) {
        rc_matrix* h = nullptr;
        ctx.check(rc_matrix_from_host_async(ctx.raw(), ScalarTraits<A>::dtype, data, (int64_t)rows, (int64_t)cols,
                                            (int64_t)row_stride, &h));
        return Matrix(ctx, h);
    }
    Matrix& await_upload(bool block_host = false) {
        ctx_.check(rc_matrix_await(ctx_.raw(), h_.get(), block_host ? 1 : 0));
        return *this;
    }
    // RandomMatrix (src/random_matrix.rs:21-93), seeded
    static Matrix random_gaussian(const Context& ctx, size_t rows, size_t cols, uint64_t seed) {
        rc_matrix* h = nullptr;
        ctx.check(rc_random_gaussian(ctx.raw(), ScalarTraits<A>::dtype, (int64_t)rows, (int64_t)cols, seed, 0, 0, &h));
        return Matrix(ctx, h);
    }
    static Matrix random_orthogonal_matrix(const Context& ctx, size_t rows, size_t cols, uint64_t seed) {
        rc_matrix* h = nullptr;
        ctx.check(rc_random_orthogonal_matrix(ctx.raw(), ScalarTraits<A>::dtype, (int64_t)rows, (int64_t)cols, seed, 0, &h));
        return Matrix(ctx, h);
    }
    static Matrix random_approximate_low_rank_matrix(const Context& ctx, size_t rows, size_t cols, double sigma_max,
                                                     double sigma_min, uint64_t seed) {
        rc_matrix* h = nullptr;
        ctx.check(rc_random_approximate_low_rank_matrix(ctx.raw(), ScalarTraits<A>::dtype, (int64_t)rows, (int64_t)cols,
                                                        sigma_max, sigma_min, seed, &h));
        return Matrix(ctx, h);
    }

    size_t nrows() const { return (size_t)rc_matrix_rows(h_.get()); }
    size_t ncols() const { return (size_t)rc_matrix_cols(h_.get()); }
    rc_matrix* raw() const { return h_.get(); }
    const Context& context() const { return ctx_; }
    explicit operator bool() const { return (bool)h_; }

    std::vector<A> to_host() const {           // dense row-major copy
        std::vector<A> out(nrows() * ncols());
        ctx_.check(rc_matrix_to_host(ctx_.raw(), h_.get(), out.data()));
        return out;
    }
    // MatMat::matmat / ConjMatMat::conj_matmat (src/types.rs:58-71, 88-101)
    Matrix matmat(const Matrix& x) const {
        rc_matrix* y = nullptr;
        ctx_.check(rc_matmat(ctx_.raw(), h_.get(), x.raw(), &y));
        return Matrix(ctx_, y);
    }
    Matrix conj_matmat(const Matrix& x) const {
        rc_matrix* z = nullptr;
        ctx_.check(rc_conj_matmat(ctx_.raw(), h_.get(), x.raw(), &z));
        return Matrix(ctx_, z);
    }
    // MatVec::matvec / ConjMatVec::conj_matvec (src/types.rs:40-51, 77-81): host vector in, host vector out
    std::vector<A> matvec(const std::vector<A>& x) const { return matmat(from_host(ctx_, x, x.size(), 1)).to_host(); }
    std::vector<A> conj_matvec(const std::vector<A>& x) const { return conj_matmat(from_host(ctx_, x, x.size(), 1)).to_host(); }

  private:
    struct Deleter {
        bool owned;
        void operator()(rc_matrix* m) const { if (owned && m) rc_matrix_free(m); }
    };
    Context ctx_{nullptr};
    std::shared_ptr<rc_matrix> h_;
};

// Matrix-free operator: the plugin API proper (a user type implementing MatMat / ConjMatMat on DEVICE buffers).
// Op must provide   int matmat(const A* x, int64_t ldx, int64_t ncols, A* y, int64_t ldy, void* cuda_stream) const
//             and   int conj_matmat(...)  with the same signature (z = A^H x).
template <class A, class Op>
class Operator {
  public:
    Operator(const Context& ctx, size_t rows, size_t cols, Op op) : op_(std::make_unique<Op>(std::move(op))) {
        rc_matrix* h = nullptr;
        ctx.check(rc_operator_create(ctx.raw(), ScalarTraits<A>::dtype, (int64_t)rows, (int64_t)cols, &Operator::mm, &Operator::cmm,
                                     op_.get(), &h));
        m_ = Matrix<A>(ctx, h);
    }
    const Matrix<A>& as_matrix() const { return m_; }     // accepted wherever an operator is (samplers, *_from_range_estimate)

  private:
    static int mm(void* user, const void* x, int64_t ldx, int64_t ncols, void* y, int64_t ldy, void* stream) {
        return static_cast<const Op*>(user)->matmat(static_cast<const A*>(x), ldx, ncols, static_cast<A*>(y), ldy, stream);
    }
    static int cmm(void* user, const void* x, int64_t ldx, int64_t ncols, void* z, int64_t ldz, void* stream) {
        return static_cast<const Op*>(user)->conj_matmat(static_cast<const A*>(x), ldx, ncols, static_cast<A*>(z), ldz, stream);
    }
    std::unique_ptr<Op> op_;
    Matrix<A> m_;
};

// ---------------------------------------------------------------------------------------------- free functions
// RelDiff (src/types.rs:162-204): ||first - second|| / ||second||
template <class A>
double rel_diff_fro(const Matrix<A>& first, const Matrix<A>& second) {
    double out = 0.0;
    first.context().check(rc_rel_diff_fro(first.context().raw(), first.raw(), second.raw(), &out));
    return out;
}
template <class A>
double rel_diff_l2(const Matrix<A>& first, const Matrix<A>& second) {      // 1 x n or n x 1 matrices
    double out = 0.0;
    first.context().check(rc_rel_diff_l2(first.context().raw(), first.raw(), second.raw(), &out));
    return out;
}
// MaxColNorm::max_col_norm (src/random_sampling.rs:175-199)
template <class A>
double max_col_norm(const Matrix<A>& m) {
    double out = 0.0;
    m.context().check(rc_max_col_norm(m.context().raw(), m.raw(), &out));
    return out;
}

// src/permutation.rs:28-38
inline std::vector<size_t> invert_permutation_vector(const std::vector<size_t>& perm) {
    std::vector<uint64_t> in = detail::to_u64(perm), out(perm.size());
    if (rc_invert_permutation_vector(in.data(), in.size(), out.data()) != RC_OK)
        throw InvalidArgument(RC_INVALID_ARGUMENT, "invert_permutation_vector: not a permutation");
    return detail::to_usize(out);
}
// ApplyPermutationToMatrix::apply_permutation (src/permutation.rs:40-56, 77-145)
template <class A>
Matrix<A> apply_permutation(const Matrix<A>& m, const std::vector<size_t>& index_array, MatrixPermutationMode mode) {
    std::vector<uint64_t> idx = detail::to_u64(index_array);
    rc_matrix* out = nullptr;
    m.context().check(rc_apply_permutation_matrix(m.context().raw(), m.raw(), idx.data(), idx.size(), (rc_perm_mode)mode, &out));
    return Matrix<A>(m.context(), out);
}
// ApplyPermutationToVector::apply_permutation (src/permutation.rs:58-75, 147-184); v is 1 x n or n x 1
template <class A>
Matrix<A> apply_permutation(const Matrix<A>& v, const std::vector<size_t>& index_array, VectorPermutationMode mode) {
    std::vector<uint64_t> idx = detail::to_u64(index_array);
    rc_matrix* out = nullptr;
    v.context().check(rc_apply_permutation_vector(v.context().raw(), v.raw(), idx.data(), idx.size(), (rc_vperm_mode)mode, &out));
    return Matrix<A>(v.context(), out);
}

// SampleRange::sample_range_by_rank (src/random_sampling.rs:58-72): Omega from Philox(seed), or ingested.
template <class A>
Matrix<A> sample_range_by_rank(const Matrix<A>& op, size_t k, size_t p, uint64_t seed, const Matrix<A>* omega = nullptr) {
    rc_matrix* q = nullptr;
    op.context().check(rc_sample_range_by_rank(op.context().raw(), op.raw(), (int64_t)k, (int64_t)p, omega ? omega->raw() : nullptr,
                                               seed, &q));
    return Matrix<A>(op.context(), q);
}
// SampleRangePowerIteration::sample_range_power_iteration (src/random_sampling.rs:82-98), quirk Q1 included.
template <class A>
Matrix<A> sample_range_power_iteration(const Matrix<A>& op, size_t k, size_t p, size_t it_count, uint64_t seed,
                                       const Matrix<A>* omega = nullptr) {
    rc_matrix* q = nullptr;
    op.context().check(rc_sample_range_power_iteration(op.context().raw(), op.raw(), (int64_t)k, (int64_t)p, (int64_t)it_count,
                                                       omega ? omega->raw() : nullptr, seed, &q));
    return Matrix<A>(op.context(), q);
}
// AdaptiveSampling::sample_range_adaptive (src/random_sampling.rs:202-218) -> (Q, Vec<(rank, rel_res)>)
template <class A>
std::pair<Matrix<A>, std::vector<std::pair<size_t, double>>> sample_range_adaptive(const Matrix<A>& op, double rel_tol,
                                                                                  size_t sample_size, uint64_t seed,
                                                                                  const Matrix<A>* omega_blocks = nullptr,
                                                                                  size_t max_rank = 0) {
    rc_matrix* q = nullptr;
    std::vector<uint64_t> ranks(4096);
    std::vector<double> res(4096);
    size_t len = 0;
    op.context().check(rc_sample_range_adaptive(op.context().raw(), op.raw(), rel_tol, (int64_t)sample_size,
                                                omega_blocks ? omega_blocks->raw() : nullptr, seed, (int64_t)max_rank, &q,
                                                ranks.data(), res.data(), ranks.size(), &len));
    std::vector<std::pair<size_t, double>> hist;
    for (size_t i = 0; i < len && i < ranks.size(); ++i) hist.emplace_back((size_t)ranks[i], res[i]);
    return {Matrix<A>(op.context(), q), std::move(hist)};
}

// ---------------------------------------------------------------------------------------------- decompositions
template <class A> class ColumnID;
template <class A> class RowID;
template <class A> class TwoSidedID;

// TwoSidedID / TwoSidedIDTraits / Apply (src/two_sided_interp_decomp.rs:19-173)
template <class A>
class TwoSidedID {
  public:
    TwoSidedID(Context ctx, rc_two_sided_id* h) : ctx_(std::move(ctx)), h_(h, [](rc_two_sided_id* p) { rc_two_sided_id_free(p); }) {}
    // argument order of the crate: (x, r, c, col_ind, row_ind) (:89-95)
    static TwoSidedID make(const Matrix<A>& x, const Matrix<A>& r, const Matrix<A>& c, const std::vector<size_t>& col_ind,
                           const std::vector<size_t>& row_ind) {
        std::vector<uint64_t> ci = detail::to_u64(col_ind), ri = detail::to_u64(row_ind);
        rc_two_sided_id* h = nullptr;
        x.context().check(rc_two_sided_id_new(x.context().raw(), x.raw(), r.raw(), c.raw(), ci.data(), ci.size(), ri.data(), ri.size(), &h));
        return TwoSidedID(x.context(), h);
    }
    Matrix<A> get_c() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_two_sided_id_get_c(h_.get())), false); }
    Matrix<A> get_x() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_two_sided_id_get_x(h_.get())), false); }
    Matrix<A> get_r() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_two_sided_id_get_r(h_.get())), false); }
    std::vector<size_t> get_col_ind() const {
        std::vector<uint64_t> v(rc_two_sided_id_col_ind_len(h_.get()));
        ctx_.check(rc_two_sided_id_get_col_ind(h_.get(), v.data(), v.size()));
        return detail::to_usize(v);
    }
    std::vector<size_t> get_row_ind() const {
        std::vector<uint64_t> v(rc_two_sided_id_row_ind_len(h_.get()));
        ctx_.check(rc_two_sided_id_get_row_ind(h_.get(), v.data(), v.size()));
        return detail::to_usize(v);
    }
    size_t nrows() const { return get_c().nrows(); }
    size_t ncols() const { return get_r().ncols(); }
    size_t rank() const { return get_x().nrows(); }
    Matrix<A> to_mat() const {                      // c (x r), src/two_sided_interp_decomp.rs:62-64
        rc_matrix* out = nullptr;
        ctx_.check(rc_two_sided_id_to_mat(ctx_.raw(), h_.get(), &out));
        return Matrix<A>(ctx_, out);
    }
    Matrix<A> dot(const Matrix<A>& rhs) const {     // Apply (:154-171)
        rc_matrix* out = nullptr;
        ctx_.check(rc_two_sided_id_apply(ctx_.raw(), h_.get(), rhs.raw(), &out));
        return Matrix<A>(ctx_, out);
    }

  private:
    Context ctx_;
    std::shared_ptr<rc_two_sided_id> h_;
};

// ColumnID / ColumnIDTraits / Apply (src/col_interp_decomp.rs:23-156)
template <class A>
class ColumnID {
  public:
    ColumnID(Context ctx, rc_column_id* h) : ctx_(std::move(ctx)), h_(h, [](rc_column_id* p) { rc_column_id_free(p); }) {}
    static ColumnID make(const Matrix<A>& c, const Matrix<A>& z, const std::vector<size_t>& col_ind) {      // ::new (:113)
        std::vector<uint64_t> ci = detail::to_u64(col_ind);
        rc_column_id* h = nullptr;
        c.context().check(rc_column_id_new(c.context().raw(), c.raw(), z.raw(), ci.data(), ci.size(), &h));
        return ColumnID(c.context(), h);
    }
    Matrix<A> get_c() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_column_id_get_c(h_.get())), false); }
    Matrix<A> get_z() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_column_id_get_z(h_.get())), false); }
    std::vector<size_t> get_col_ind() const {
        std::vector<uint64_t> v(rc_column_id_col_ind_len(h_.get()));
        ctx_.check(rc_column_id_get_col_ind(h_.get(), v.data(), v.size()));
        return detail::to_usize(v);
    }
    size_t nrows() const { return get_c().nrows(); }
    size_t ncols() const { return get_z().ncols(); }
    size_t rank() const { return get_c().ncols(); }
    Matrix<A> to_mat() const {
        rc_matrix* out = nullptr;
        ctx_.check(rc_column_id_to_mat(ctx_.raw(), h_.get(), &out));
        return Matrix<A>(ctx_, out);
    }
    Matrix<A> dot(const Matrix<A>& rhs) const {
        rc_matrix* out = nullptr;
        ctx_.check(rc_column_id_apply(ctx_.raw(), h_.get(), rhs.raw(), &out));
        return Matrix<A>(ctx_, out);
    }
    TwoSidedID<A> two_sided_id() const {            // pivoted LQ of C -> row_id (:116-125)
        rc_two_sided_id* out = nullptr;
        ctx_.check(rc_column_id_two_sided_id(ctx_.raw(), h_.get(), &out));
        return TwoSidedID<A>(ctx_, out);
    }

  private:
    Context ctx_;
    std::shared_ptr<rc_column_id> h_;
};

// RowID / RowIDTraits / Apply (src/row_interp_decomp.rs:25-156)
template <class A>
class RowID {
  public:
    RowID(Context ctx, rc_row_id* h) : ctx_(std::move(ctx)), h_(h, [](rc_row_id* p) { rc_row_id_free(p); }) {}
    static RowID make(const Matrix<A>& x, const Matrix<A>& r, const std::vector<size_t>& row_ind) {          // ::new (:116)
        std::vector<uint64_t> ri = detail::to_u64(row_ind);
        rc_row_id* h = nullptr;
        x.context().check(rc_row_id_new(x.context().raw(), x.raw(), r.raw(), ri.data(), ri.size(), &h));
        return RowID(x.context(), h);
    }
    Matrix<A> get_x() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_row_id_get_x(h_.get())), false); }
    Matrix<A> get_r() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_row_id_get_r(h_.get())), false); }
    std::vector<size_t> get_row_ind() const {
        std::vector<uint64_t> v(rc_row_id_row_ind_len(h_.get()));
        ctx_.check(rc_row_id_get_row_ind(h_.get(), v.data(), v.size()));
        return detail::to_usize(v);
    }
    size_t nrows() const { return get_x().nrows(); }
    size_t ncols() const { return get_r().ncols(); }
    size_t rank() const { return get_r().nrows(); }
    Matrix<A> to_mat() const {
        rc_matrix* out = nullptr;
        ctx_.check(rc_row_id_to_mat(ctx_.raw(), h_.get(), &out));
        return Matrix<A>(ctx_, out);
    }
    Matrix<A> dot(const Matrix<A>& rhs) const {
        rc_matrix* out = nullptr;
        ctx_.check(rc_row_id_apply(ctx_.raw(), h_.get(), rhs.raw(), &out));
        return Matrix<A>(ctx_, out);
    }
    TwoSidedID<A> two_sided_id() const {            // pivoted QR of R -> column_id (:120-130)
        rc_two_sided_id* out = nullptr;
        ctx_.check(rc_row_id_two_sided_id(ctx_.raw(), h_.get(), &out));
        return TwoSidedID<A>(ctx_, out);
    }

  private:
    Context ctx_;
    std::shared_ptr<rc_row_id> h_;
};

// QR / QRTraits (src/qr.rs:31-40, 141-237)
template <class A>
class QR {
  public:
    QR(Context ctx, rc_qr* h) : ctx_(std::move(ctx)), h_(h, [](rc_qr* p) { rc_qr_free(p); }) {}
    // `QR { q, r, ind }` from parts: the crate's struct has pub fields (:31-40)
    static QR make(const Matrix<A>& q, const Matrix<A>& r, const std::vector<size_t>& ind) {
        std::vector<uint64_t> iv = detail::to_u64(ind);
        rc_qr* h = nullptr;
        q.context().check(rc_qr_new(q.context().raw(), q.raw(), r.raw(), iv.data(), iv.size(), &h));
        return QR(q.context(), h);
    }
    static QR compute_from(const Matrix<A>& arr) {                                           // :214, 251-253
        rc_qr* h = nullptr;
        arr.context().check(rc_qr_compute_from(arr.context().raw(), arr.raw(), &h));
        return QR(arr.context(), h);
    }
    static QR compute_from_range_estimate(const Matrix<A>& range, const Matrix<A>& op) {     // :221-224, 311-323
        rc_qr* h = nullptr;
        op.context().check(rc_qr_compute_from_range_estimate(op.context().raw(), range.raw(), op.raw(), &h));
        return QR(op.context(), h);
    }
    Matrix<A> get_q() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_qr_get_q(h_.get())), false); }
    Matrix<A> get_r() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_qr_get_r(h_.get())), false); }
    std::vector<size_t> get_ind() const {
        std::vector<uint64_t> v((size_t)rc_qr_ncols(h_.get()));
        ctx_.check(rc_qr_get_ind(h_.get(), v.data(), v.size()));
        return detail::to_usize(v);
    }
    size_t nrows() const { return (size_t)rc_qr_nrows(h_.get()); }
    size_t ncols() const { return (size_t)rc_qr_ncols(h_.get()); }
    size_t rank() const { return (size_t)rc_qr_rank(h_.get()); }
    Matrix<A> to_mat() const {                                                               // :160-166
        rc_matrix* out = nullptr;
        ctx_.check(rc_qr_to_mat(ctx_.raw(), h_.get(), &out));
        return Matrix<A>(ctx_, out);
    }
    QR compress_qr_rank(size_t max_rank) const {                                             // :169-184
        rc_qr* out = nullptr;
        ctx_.check(rc_qr_compress_rank(ctx_.raw(), h_.get(), (int64_t)max_rank, &out));
        return QR(ctx_, out);
    }
    QR compress_qr_tolerance(double tol) const {                                             // :187-200 (throws CompressionError)
        rc_qr* out = nullptr;
        ctx_.check(rc_qr_compress_tolerance(ctx_.raw(), h_.get(), tol, &out));
        return QR(ctx_, out);
    }
    QR compress(const CompressionType& t) const {                                            // :203-208
        return t.kind == CompressionType::RANK_ ? compress_qr_rank(t.rank) : compress_qr_tolerance(t.tol);
    }
    ColumnID<A> column_id() const {                                                          // :270-309
        rc_column_id* out = nullptr;
        ctx_.check(rc_qr_column_id(ctx_.raw(), h_.get(), &out));
        return ColumnID<A>(ctx_, out);
    }

  private:
    Context ctx_;
    std::shared_ptr<rc_qr> h_;
};

// LQ / LQTraits (src/qr.rs:42-51, 54-139, 326-405)
template <class A>
class LQ {
  public:
    LQ(Context ctx, rc_lq* h) : ctx_(std::move(ctx)), h_(h, [](rc_lq* p) { rc_lq_free(p); }) {}
    // `LQ { l, q, ind }` from parts (pub fields, :42-51)
    static LQ make(const Matrix<A>& l, const Matrix<A>& q, const std::vector<size_t>& ind) {
        std::vector<uint64_t> iv = detail::to_u64(ind);
        rc_lq* h = nullptr;
        l.context().check(rc_lq_new(l.context().raw(), l.raw(), q.raw(), iv.data(), iv.size(), &h));
        return LQ(l.context(), h);
    }
    static LQ compute_from(const Matrix<A>& arr) {                                           // :135, 354-362
        rc_lq* h = nullptr;
        arr.context().check(rc_lq_compute_from(arr.context().raw(), arr.raw(), &h));
        return LQ(arr.context(), h);
    }
    Matrix<A> get_l() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_lq_get_l(h_.get())), false); }
    Matrix<A> get_q() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_lq_get_q(h_.get())), false); }
    std::vector<size_t> get_ind() const {
        std::vector<uint64_t> v((size_t)rc_lq_nrows(h_.get()));
        ctx_.check(rc_lq_get_ind(h_.get(), v.data(), v.size()));
        return detail::to_usize(v);
    }
    size_t nrows() const { return (size_t)rc_lq_nrows(h_.get()); }
    size_t ncols() const { return (size_t)rc_lq_ncols(h_.get()); }
    size_t rank() const { return (size_t)rc_lq_rank(h_.get()); }
    Matrix<A> to_mat() const {                                                               // :73-78
        rc_matrix* out = nullptr;
        ctx_.check(rc_lq_to_mat(ctx_.raw(), h_.get(), &out));
        return Matrix<A>(ctx_, out);
    }
    LQ compress_lq_rank(size_t max_rank) const {                                             // :80-95
        rc_lq* out = nullptr;
        ctx_.check(rc_lq_compress_rank(ctx_.raw(), h_.get(), (int64_t)max_rank, &out));
        return LQ(ctx_, out);
    }
    LQ compress_lq_tolerance(double tol) const {                                             // :98-111
        rc_lq* out = nullptr;
        ctx_.check(rc_lq_compress_tolerance(ctx_.raw(), h_.get(), tol, &out));
        return LQ(ctx_, out);
    }
    LQ compress(const CompressionType& t) const {                                            // :114-119
        return t.kind == CompressionType::RANK_ ? compress_lq_rank(t.rank) : compress_lq_tolerance(t.tol);
    }
    RowID<A> row_id() const {                                                                // :363-403
        rc_row_id* out = nullptr;
        ctx_.check(rc_lq_row_id(ctx_.raw(), h_.get(), &out));
        return RowID<A>(ctx_, out);
    }

  private:
    Context ctx_;
    std::shared_ptr<rc_lq> h_;
};

// SVD / SVDTraits (src/svd.rs:13-20, 23-186)
template <class A>
class SVD {
  public:
    using Real = typename ScalarTraits<A>::Real;
    SVD(Context ctx, rc_svd* h) : ctx_(std::move(ctx)), h_(h, [](rc_svd* p) { rc_svd_free(p); }) {}
    // `SVD { u, s, vt }` from parts (pub fields, :13-20)
    static SVD make(const Matrix<A>& u, const std::vector<Real>& s, const Matrix<A>& vt) {
        std::vector<double> sv(s.begin(), s.end());
        rc_svd* h = nullptr;
        u.context().check(rc_svd_new(u.context().raw(), u.raw(), sv.data(), sv.size(), vt.raw(), &h));
        return SVD(u.context(), h);
    }
    static SVD compute_from(const Matrix<A>& arr) {                                          // :103, 165-169
        rc_svd* h = nullptr;
        arr.context().check(rc_svd_compute_from(arr.context().raw(), arr.raw(), &h));
        return SVD(arr.context(), h);
    }
    static SVD compute_from_range_estimate(const Matrix<A>& range, const Matrix<A>& op) {    // :110-113, 171-183
        rc_svd* h = nullptr;
        op.context().check(rc_svd_compute_from_range_estimate(op.context().raw(), range.raw(), op.raw(), &h));
        return SVD(op.context(), h);
    }
    Matrix<A> get_u() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_svd_get_u(h_.get())), false); }
    Matrix<A> get_vt() const { return Matrix<A>(ctx_, const_cast<rc_matrix*>(rc_svd_get_vt(h_.get())), false); }
    std::vector<Real> get_s() const {                                                        // descending, A::Real
        std::vector<double> s((size_t)rc_svd_rank(h_.get()));
        ctx_.check(rc_svd_get_s(h_.get(), s.data(), s.size()));
        return std::vector<Real>(s.begin(), s.end());
    }
    size_t nrows() const { return get_u().nrows(); }
    size_t ncols() const { return get_vt().ncols(); }
    size_t rank() const { return (size_t)rc_svd_rank(h_.get()); }
    Matrix<A> to_mat() const {                                                               // :42-54
        rc_matrix* out = nullptr;
        ctx_.check(rc_svd_to_mat(ctx_.raw(), h_.get(), &out));
        return Matrix<A>(ctx_, out);
    }
    QR<A> to_qr() const {                                                                    // :57, 150-163
        rc_qr* out = nullptr;
        ctx_.check(rc_svd_to_qr(ctx_.raw(), h_.get(), &out));
        return QR<A>(ctx_, out);
    }
    SVD compress_svd_rank(size_t max_rank) const {                                           // :68-84
        rc_svd* out = nullptr;
        ctx_.check(rc_svd_compress_rank(ctx_.raw(), h_.get(), (int64_t)max_rank, &out));
        return SVD(ctx_, out);
    }
    SVD compress_svd_tolerance(double tol) const {                                           // :87-101
        rc_svd* out = nullptr;
        ctx_.check(rc_svd_compress_tolerance(ctx_.raw(), h_.get(), tol, &out));
        return SVD(ctx_, out);
    }
    SVD compress(const CompressionType& t) const {                                           // :60-65
        return t.kind == CompressionType::RANK_ ? compress_svd_rank(t.rank) : compress_svd_tolerance(t.tol);
    }

  private:
    Context ctx_;
    std::shared_ptr<rc_svd> h_;
};

}  // namespace rcb200

#endif  // RUSTY_COMPRESSION_B200_HPP
