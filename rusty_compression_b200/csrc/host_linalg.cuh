// Host-side orchestration of the reference's algorithms over the CUDA kernels (templates shared
// by host_api.cu).  Every routine cites the reference lines whose behaviour it reproduces.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <memory>
#include "rc_internal.cuh"

// ----------------------------------------------------------------------------- matrices
inline int64_t rc_pad_ld(int dtype, int64_t cols) {
    int64_t e = (int64_t)(16 / rc_dtype_size(dtype));   // elements per 16 bytes (TMA row pitch)
    if (e < 1) e = 1;
    return std::max<int64_t>((cols + e - 1) / e * e, e);
}

inline uint64_t rc_next_matrix_id() {
    static std::atomic<uint64_t> next{1};
    return next.fetch_add(1);
}
inline rc_matrix* mat_new(rc_ctx* c, int dtype, int64_t rows, int64_t cols) {
    RC_REQUIRE(rows >= 0 && cols >= 0, "negative matrix dimension");
    std::unique_ptr<rc_matrix> m(new rc_matrix());
    m->ctx = c; m->dtype = dtype; m->rows = rows; m->cols = cols; m->ld = rc_pad_ld(dtype, cols);
    m->owns = true;
    m->id = rc_next_matrix_id();
    size_t bytes = (size_t)std::max<int64_t>(rows, 1) * m->ld * rc_dtype_size(dtype);
    m->data = rc_dev_alloc(c, bytes);
    return m.release();
}
inline void mat_free(rc_matrix* m) {
    if (!m) return;
    if (m->upload_done) {                    // never awaited: the copy must not outlive the buffer it writes
        DeviceGuard dg(m->ctx->device);
        cudaStreamWaitEvent(m->ctx->stream, m->upload_done, 0);
        cudaEventDestroy(m->upload_done);
        m->upload_done = nullptr;
    }
    if (m->owns && m->data) {
        DeviceGuard dg(m->ctx->device);      // the *_free entry points are not routed through guard()
        rc_dev_free(m->ctx, m->data);
    }
    if (m->companion) mat_free(m->companion);
    delete m;
}
struct MatPtr {   // RAII owner used while a routine can still throw
    rc_matrix* m = nullptr;
    MatPtr() {}
    explicit MatPtr(rc_matrix* p) : m(p) {}
    ~MatPtr() { mat_free(m); }
    MatPtr(const MatPtr&) = delete;
    MatPtr& operator=(const MatPtr&) = delete;
    void reset(rc_matrix* p) { mat_free(m); m = p; }
    rc_matrix* release() { rc_matrix* r = m; m = nullptr; return r; }
    rc_matrix* operator->() const { return m; }
    rc_matrix* get() const { return m; }
};
template <class T> inline T* P(const rc_matrix* m) {
    if (m->op_matmat) RC_THROW(RC_INVALID_ARGUMENT, "a matrix-free operator handle was passed where a dense matrix is required");
    return reinterpret_cast<T*>(m->data);
}
inline bool mat_sharded(const rc_matrix* m) { return m->global_rows > 0 && m->ctx->nranks > 1; }
inline void inherit_shard(rc_matrix* dst, const rc_matrix* src) {
    dst->global_rows = src->global_rows;
    dst->row_offset = src->row_offset;
}

template <class T>
rc_matrix* mat_clone(rc_ctx* c, const rc_matrix* a) {
    MatPtr o(mat_new(c, a->dtype, a->rows, a->cols));
    k_copy<T>(c, P<T>(o.get()), o->ld, P<T>(a), a->ld, a->rows, a->cols);
    inherit_shard(o.get(), a);
    return o.release();
}
// rows [r0, r1) x cols [c0, c1) copied out
template <class T>
rc_matrix* mat_slice(rc_ctx* c, const rc_matrix* a, int64_t r0, int64_t r1, int64_t c0, int64_t c1) {
    MatPtr o(mat_new(c, a->dtype, r1 - r0, c1 - c0));
    k_copy<T>(c, P<T>(o.get()), o->ld, P<T>(a) + r0 * a->ld + c0, a->ld, r1 - r0, c1 - c0);
    if (r0 == 0 && r1 == a->rows) inherit_shard(o.get(), a);
    return o.release();
}
template <class T>
rc_matrix* mat_conj_transpose(rc_ctx* c, const rc_matrix* a) {
    MatPtr o(mat_new(c, a->dtype, a->cols, a->rows));
    k_transpose<T>(c, P<T>(o.get()), o->ld, P<T>(a), a->ld, a->rows, a->cols, true);
    return o.release();
}

inline std::vector<int> to_int_index(const std::vector<uint64_t>& v) {
    std::vector<int> o(v.size());
    for (size_t i = 0; i < v.size(); ++i) o[i] = (int)v[i];
    return o;
}
inline std::vector<uint64_t> invert_perm(const std::vector<uint64_t>& p) {
    std::vector<uint64_t> inv(p.size());
    for (size_t i = 0; i < p.size(); ++i) {
        RC_REQUIRE(p[i] < p.size(), "index array is not a permutation");
        inv[p[i]] = i;
    }
    return inv;
}
struct DevIndex {   // host index vector uploaded as int
    DevBuf<int> d;
    DevIndex(rc_ctx* c, const std::vector<uint64_t>& v) {
        std::vector<int> h = to_int_index(v);
        d.alloc(c, std::max<size_t>(h.size(), 1));
        if (!h.empty()) {
            RC_CUDA(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
            RC_CUDA(cudaStreamSynchronize(c->stream));   // h goes out of scope
        }
    }
};

// ----------------------------------------------------------------------------- GEMM dispatch
template <class T>
void gemm(rc_ctx* c, RcOp opa, RcOp opb, int64_t M, int64_t N, int64_t K, const T* A, int64_t lda,
          const T* B, int64_t ldb, T* C, int64_t ldc, T alpha, T beta);

template <class T>
rc_matrix* mat_mul(rc_ctx* c, RcOp opa, const rc_matrix* a, RcOp opb, const rc_matrix* b) {
    int64_t M = (opa == RC_OP_N) ? a->rows : a->cols, K = (opa == RC_OP_N) ? a->cols : a->rows;
    int64_t Kb = (opb == RC_OP_N) ? b->rows : b->cols, N = (opb == RC_OP_N) ? b->cols : b->rows;
    RC_REQUIRE(K == Kb, "matrix product: inner dimensions differ (%lld vs %lld)", (long long)K, (long long)Kb);
    MatPtr o(mat_new(c, a->dtype, M, N));
    gemm<T>(c, opa, opb, M, N, K, P<T>(a), a->ld, P<T>(b), b->ld, P<T>(o.get()), o->ld, rc_one<T>(), rc_zero<T>());
    return o.release();
}

// ----------------------------------------------------------------------------- results
struct QrParts {
    MatPtr q, r;
    std::vector<uint64_t> ind;
    bool want_ind = true;     // false: the caller only needs q / r (the samplers); skips the device-to-host copy of the
                              // pivot vector and its host synchronisation
};
struct SvdParts {
    MatPtr u, vt;
    std::vector<double> s;
};

template <class T> void pivoted_qr_impl(rc_ctx* c, const rc_matrix* arr, bool input_is_conj_transposed, int64_t ncq,
                                        bool may_destroy, QrParts& out);
template <class T> void svd_impl(rc_ctx* c, const rc_matrix* arr, bool input_is_conj_transposed, SvdParts& out);
