/*
 * fastad_b200.h — C ABI of libfastad_b200.so: the B200 (sm_100a) implementation of FastAD's
 * reverse-mode hot path (ad::autodiff = feval + beval over Var/VarView expressions).
 *
 * The reference (JamesYang007/FastAD, header-only C++17) has no FFI surface: its hot path
 * sits behind a C++ template concept (feval/beval/bind_cache).  Each entry point below
 * names the reference interface it replaces, `file:line` relative to
 * /root/reference/include/fastad_bits/ .  The C++17 host layer in include/fastad (ad::Var,
 * ad::bind, ad::autodiff ...) is a thin front-end over this ABI; so is the ctypes
 * front-end in fastad_b200/__init__.py.
 *
 * Conventions
 *   - plain pointers and sizes only; every `double*` below is a DEVICE pointer unless the
 *     parameter name starts with `host_`;
 *   - matrices are column-major (value_view.hpp:141-173);
 *   - adjoints ACCUMULATE (`adj += seed`, var_view.hpp:91-94) unless FADB_OVERWRITE_ADJ;
 *   - every function returns 0 on success or a negative FADB_ERR_* code; the message is
 *     available from fadb_last_error() (thread-local).  No C++ exception crosses the ABI;
 *   - calls are asynchronous on the library stream of the current CUDA device, except
 *     those that return a host scalar (they synchronise that stream);
 *   - there is NO CPU fallback: without a usable CUDA device every call fails with
 *     FADB_ERR_CUDA.
 */
#ifndef FASTAD_B200_H
#define FASTAD_B200_H
#ifndef __CUDACC_RTC__
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define FADB_OK 0
#define FADB_ERR_ARG (-1)
#define FADB_ERR_CUDA (-2)
#define FADB_ERR_UNSUPPORTED (-3)
#define FADB_ERR_STATE (-4)

/* flags */
#define FADB_OVERWRITE_ADJ 1u  /* adjoint outputs are stored, not accumulated ("fused reset") */
#define FADB_TRUST_SIGMA 2u    /* caller guarantees every element is in range (sigma > 0 ...): no validity guard at all */
#define FADB_STRICT_SIGMA 4u   /* decide validity in a read-only pre-pass BEFORE any adjoint is written (bit-exact "untouched"
                                  on invalid input, +8 B/element) instead of the default optimistic pass + undo         */

/* shapes: ad::scl / ad::vec / ad::mat (util/shape_traits.hpp:7-9) */
enum { FADB_SCL = 0, FADB_VEC = 1, FADB_MAT = 2 };

/* unary functors (reverse/core/unary.hpp:190-296) */
enum {
    FADB_NEG = 0, FADB_SIN, FADB_COS, FADB_TAN, FADB_ASIN, FADB_ACOS, FADB_ATAN, FADB_EXP,
    FADB_LOG, FADB_SQRT, FADB_ERF, FADB_SIGMOID, FADB_SINH, FADB_COSH, FADB_TANH,
    FADB_SQUARE, /* x*x with adjoint seed*2*x: the map inside NormNode (reverse/core/norm.hpp:47-57) */
    FADB_UNARY_COUNT,
    FADB_MOCK2X = 100 /* test functor f(x)=2x (test/testutil/base_fixture.hpp:9-22) */
};
/* binary functors (reverse/core/binary.hpp:265-338) */
enum { FADB_ADD = 0, FADB_SUB, FADB_MUL, FADB_DIV,
       /* comparison / logical functors: value 0 or 1, no adjoint (binary.hpp:347-470) */
       FADB_LT, FADB_LE, FADB_GT, FADB_GE, FADB_EQ, FADB_NE, FADB_AND, FADB_OR, FADB_BINARY_COUNT,
       FADB_MOCKB = 100 /* test functor f(x,y)=x-2y (base_fixture.hpp:24-46) */ };

#ifndef __CUDACC_RTC__   /* device code (NVRTC, csrc/jit.cu) needs only the constants above */
/* ------------------------------------------------------------------ lifecycle / runtime */
/* Binds the library to the current CUDA device (cudaSetDevice(device)), creates its stream
 * and workspace.  Idempotent. */
int fadb_init(int device);
void fadb_shutdown(void);
int fadb_device_count(void);
int fadb_get_device(void);                     /* the current CUDA device (>= 0) or a negative FADB_ERR_* code */
/* Use a caller-owned cudaStream_t (e.g. torch's current stream) for all launches on the
 * current device; NULL restores the library's own stream (pass cudaStreamLegacy, (void*)1, to
 * name the legacy default stream). */
int fadb_set_stream(void* cuda_stream);
int fadb_get_stream(void** out_cuda_stream);   /* the stream launches currently go to */
int fadb_sync(void);
const char* fadb_last_error(void);
/* number of kernels this library has launched on the current device since fadb_init */
unsigned long long fadb_launch_count(void);
const char* fadb_version(void);

/* ------------------------------------------------------------------ value/adjoint storage
 * Replaces the host-heap value/adjoint storage of Var (var.hpp:23-220), the ExprBind caches
 * (bind.hpp:38-40) and util::PtrPack (util/ptr_pack.hpp:13-25): buffers now live in HBM.
 * The caller owns every buffer. */
int fadb_malloc(size_t nbytes, void** out_dev);
int fadb_free(void* dev);
int fadb_h2d(void* dev_dst, const void* host_src, size_t nbytes);
int fadb_d2h(void* host_dst, const void* dev_src, size_t nbytes);
int fadb_memset_zero(void* dev, size_t nbytes);           /* reset_adj(), value_adj_view.hpp:57 */
int fadb_host_alloc_pinned(size_t nbytes, void** out_host);
int fadb_host_free_pinned(void* host);
/* page-lock + map an existing host array (caller keeps ownership; unregister before freeing it): the *_host
 * entry points then run ONE launch over PCIe on it instead of staging copies */
int fadb_host_register(void* host, size_t nbytes);
int fadb_host_unregister(void* host);

/* ------------------------------------------------------------------ K1: Black-Scholes
 * The expression of example/black_scholes.cpp:16-39 (call and put, each bound and
 * autodiff'ed as in :43-70) evaluated for n independent options, S/sigma/r as variables:
 * forward pass + adjoint sweep in registers, SoA in / SoA out.
 * out[8] = call, put, dcall/dS, dcall/dsigma, dcall/dr, dput/dS, dput/dsigma, dput/dr.
 * Prices are stored; the six adjoints accumulate unless FADB_OVERWRITE_ADJ. */
int fadb_black_scholes(size_t n, const double* S, const double* K, const double* sigma,
                       const double* tau, const double* r, double* const out[8],
                       unsigned flags);
/* `count` batches of n options in ONE host call: batch b runs fadb_black_scholes on operand set (first_set + b) % nsets,
 * in[5*s + {0..4}] = S, K, sigma, tau, r and out[8*s + {0..7}] of set s.  One launch per batch, back to back on the
 * library stream: the per-option loop of example/black_scholes.cpp:43-70 over count x n options without host work
 * between the batches.  The operand sets must not alias each other (no array of one set inside another set): with
 * FADB_OVERWRITE_ADJ and nsets >= 2 the batches are independent, and batch b+1 starts on the SMs batch b frees
 * instead of waiting for its slowest CTA.  Towards the rest of the stream the call is ordered like count ordinary
 * launches: nothing starts before the work enqueued ahead of the call has completed, and work enqueued behind it
 * sees every batch complete (FADB_BS_EARLY=0 turns the overlap off). */
int fadb_black_scholes_batches(size_t count, size_t n, size_t nsets, size_t first_set,
                               const double* const* in, double* const* out, unsigned flags);
/* Same call through HOST buffers (host↔device copies inside, chunked and overlapped). */
int fadb_black_scholes_host(size_t n, const double* host_S, const double* host_K,
                            const double* host_sigma, const double* host_tau,
                            const double* host_r, double* const host_out[8]);
/* the transfer scheme the call above currently uses on this device: 1 = one launch over PCIe on page-locked arrays, 0 = staged
 * through the copy engines, -1 = still measuring (it times both on its first six calls and keeps the faster;
 * FADB_BS_HOST_MODE=0/1 pins it); out_ms[2] (may be NULL) = best call time seen for {staged, direct}, -1 if not measured */
int fadb_black_scholes_host_mode(double out_ms[2]);

/* ------------------------------------------------------------------ K2: sum of a unary map
 * autodiff of ad::sum(f(x)) for a VarView x (sum.hpp:129-203 over unary.hpp:32-120 /
 * pow.hpp:67-196): value = sum_i f(x_i); x_adj_i += seed * f'(x_i); one pass.
 * op = FADB_* unary code, or 1000+n for pow<n>.  *host_value receives the value. */
int fadb_sum_unary(int op, size_t n, const double* x, double* x_adj, double seed,
                   unsigned flags, double* host_value);
/* same, value left in a device double (no host synchronisation) */
int fadb_sum_unary_async(int op, size_t n, const double* x, double* x_adj, double seed,
                         unsigned flags, double* dev_value);

/* ------------------------------------------------------------------ K3: product
 * autodiff of ad::prod(x) (prod.hpp:143-245) including the exact-zero cases. */
int fadb_prod(size_t n, const double* x, double* x_adj, double seed, unsigned flags,
              double* host_value);

/* ------------------------------------------------------------------ K4: normal log-pdf
 * autodiff of ad::normal_adj_log_pdf(x, mu, sigma) with leaf operands
 * (reverse/stat/normal.hpp:108-578; builder :785-797).
 *   shape_code: bit0 = mu is a vector, bit1 = sigma is a vector (x is a vector of n;
 *               n == 1 with shape_code 0 is the sss case);
 *   an operand is a CONSTANT (no adjoint) iff its *_adj pointer is NULL;
 *   scalar mu/sigma are device scalars; their adjoints are device scalars too.
 * Returns -inf in *host_value and leaves all adjoints untouched when any sigma <= 0
 * (normal.hpp:147,160,440-442,457).  For a vector sigma "untouched" is reached by an optimistic single pass plus, in the
 * invalid case only, an undo pass that takes the contributions back: exact for zero-initialised or FADB_OVERWRITE_ADJ
 * adjoints (they read 0), to within one rounding of (previous + contribution) for accumulated ones; FADB_STRICT_SIGMA
 * selects the bit-exact pre-pass form. */
int fadb_normal_lpdf(int shape_code, size_t n, const double* x, const double* mu,
                     const double* sigma, double* x_adj, double* mu_adj, double* sigma_adj,
                     double seed, unsigned flags, double* host_value);
/* Multi-GPU building blocks (element-range shards).  Vector adjoints only need
 * the seed, so each shard runs the same fused pass; what must cross GPUs are the partial sums
 * dev_partial4 = {sum d^2 | z^2, sum log sigma_i, sum d | d/sigma_i^2, #sigma<=0}:
 *   fadb_normal_lpdf_partial                 (fused value sums + vector adjoints of the shard; dev_invalid_in = NULL:
 *                                             optimistic -- elements with sigma_i <= 0 add nothing and are counted)
 *   all-reduce dev_partial4 (ncclSum)        (the caller's collective)
 *   fadb_normal_lpdf_finish                  (value + adjoints of scalar operands, n_total global)
 *   fadb_normal_lpdf_undo                    (vector sigma without FADB_TRUST_SIGMA: takes the shard's vector adjoints back
 *                                             when the global count dev_partial4[3] != 0; returns at once otherwise)
 * fadb_normal_lpdf is exactly this chain on one device.  The strict form instead starts with
 *   fadb_sigma_check -> all-reduce count -> fadb_normal_lpdf_partial(dev_invalid_in = count) and needs no undo. */
int fadb_sigma_check(size_t n, const double* sigma, double* dev_count);
int fadb_normal_lpdf_partial(int shape_code, size_t n, const double* x, const double* mu,
                             const double* sigma, double* x_adj, double* mu_adj,
                             double* sigma_adj, double seed, unsigned flags,
                             const double* dev_invalid_in, double* dev_partial4);
int fadb_normal_lpdf_finish(int shape_code, size_t n_total, const double* dev_partial4,
                            const double* x, const double* mu, const double* sigma,
                            double* x_adj, double* mu_adj, double* sigma_adj, double seed,
                            unsigned flags, double* host_value);
/* dev_invalid_count: device double holding the GLOBAL number of sigma_i <= 0 (dev_partial4 + 3 after the all-reduce);
 * NULL = the count published by the last fadb_normal_lpdf_finish / _finish_pull on this device. */
int fadb_normal_lpdf_undo(int shape_code, size_t n, const double* x, const double* mu, const double* sigma,
                          double* x_adj, double* mu_adj, double* sigma_adj, double seed, unsigned flags,
                          const double* dev_invalid_count);

/* ------------------------------------------------------------------ K7: dense-covariance normal
 * autodiff of ad::normal_adj_log_pdf(x, mu, Sigma) with Sigma an n x n matrix (vsm / vvm,
 * reverse/stat/normal.hpp:581-773): blocked Cholesky of the LOWER triangle of Sigma (Eigen::LLT<Lower>),
 * log-det, explicit inverse, z = inv (x - mu); Sigma_adj += -0.5 seed (inv - z z^T), mu_adj += seed z
 * (its sum if mu is a scalar: mu_vec = 0), x_adj += -seed z.  Not positive definite => -inf and no
 * adjoint update (normal.hpp:627-629, 644).  `factor` = 0 reuses the factorisation of the previous
 * call on this device (constant Sigma, normal.hpp:605-607). */
int fadb_normal_dense(size_t n, const double* x, const double* mu, int mu_vec, const double* Sigma,
                      double* x_adj, double* mu_adj, double* Sigma_adj, double seed, unsigned flags,
                      double* host_value);
int fadb_normal_dense_async(size_t n, const double* x, const double* mu, int mu_vec, const double* Sigma,
                            double* x_adj, double* mu_adj, double* Sigma_adj, double seed, unsigned flags,
                            int factor, double* dev_value);

/* ------------------------------------------------------------------ det / log_det / wishart
 * autodiff of ad::det<Policy>(A) / ad::log_det<Policy>(A) (reverse/core/det.hpp:25-92, 211-231;
 * log_det.hpp): value det A (resp. log|det A|), A_adj += seed * det * A^{-T} (resp. seed * A^{-T});
 * an invalid decomposition (singular / not positive definite) leaves the adjoint untouched
 * (det.hpp:58-60).  policy selects the reference's decomposition functor (det.hpp:97-202). */
enum { FADB_DET_FULLPIVLU = 0, FADB_DET_LDLT = 1, FADB_DET_LLT = 2 };
int fadb_det(size_t n, const double* A, double* A_adj, double seed, unsigned flags, int policy,
             int take_log, double* host_value);
int fadb_det_async(size_t n, const double* A, double* A_adj, double seed, unsigned flags, int policy,
                   int take_log, int factor, double* dev_value);
/* autodiff of ad::wishart_adj_log_pdf(X, V, n) (reverse/stat/wishart.hpp:92-231, builder :242-257):
 * X, V p x p column-major (lower triangles factored, X V^-1 formed from the full X), df constant */
int fadb_wishart(size_t p, const double* X, const double* V, double df, double* X_adj, double* V_adj,
                 double seed, unsigned flags, double* host_value);
int fadb_wishart_async(size_t p, const double* X, const double* V, double df, double* X_adj,
                       double* V_adj, double seed, unsigned flags, int factor, double* dev_value);

/* ------------------------------------------------------------------ other diagonal log-pdfs
 * autodiff of ad::cauchy_adj_log_pdf(x, loc, scale) / ad::uniform_adj_log_pdf(x, min, max) /
 * ad::bernoulli_adj_log_pdf(x, p) with leaf operands (reverse/stat/cauchy.hpp:102-451,
 * uniform.hpp:101-536, bernoulli.hpp:94-280).  dist = FADB_CAUCHY | FADB_UNIFORM | FADB_BERNOULLI
 * (declared with the expression builder below); shape_code bit0: b is a vector, bit1: c is a
 * vector; an operand is constant iff its *_adj pointer is NULL; out-of-range parameters give
 * -inf (bernoulli: bernoulli.hpp:127-133) and leave every adjoint untouched.  fadb_logpdf_async
 * leaves {value, #invalid, scratch, scratch} in 4 device doubles instead of synchronising. */
int fadb_logpdf(int dist, int shape_code, size_t n, const double* x, const double* b, const double* c,
                double* x_adj, double* b_adj, double* c_adj, double seed, unsigned flags,
                double* host_value);
int fadb_logpdf_async(int dist, int shape_code, size_t n, const double* x, const double* b,
                      const double* c, double* x_adj, double* b_adj, double* c_adj, double seed,
                      unsigned flags, double* dev_out4);

/* ------------------------------------------------------------------ K5: regression
 * autodiff of normal_adj_log_pdf(y, dot(X, w), sigma) (dot.hpp:54-129 feeding
 * normal.hpp:304-372): X rows x cols column-major constant, y constant, w variable vector,
 * sigma variable scalar.  One pass over X computes r = y - Xw, sum r^2 and X^T r.
 * Row shards: fadb_linreg_partial -> all-reduce (cols+1 doubles: {sum r^2, X^T r}) ->
 * fadb_linreg_finish with the global row count. */
int fadb_linreg_partial(size_t rows, size_t cols, const double* X, const double* y,
                        const double* w, double* dev_partial);
int fadb_linreg_finish(size_t rows_total, size_t cols, const double* dev_partial,
                       const double* sigma, double* w_adj, double* sigma_adj, double seed,
                       unsigned flags, double* host_value);
int fadb_linreg_nll(size_t rows, size_t cols, const double* X, const double* y,
                    const double* w, double* w_adj, const double* sigma, double* sigma_adj,
                    double seed, unsigned flags, double* host_value);
/* Least squares sum((y - X w)^2): value + w_adj += seed * (-2) X^T (y - X w) from one pass over X -- the
 * batched form of the reference's regression example, bind(pow<2>(yi - dot(xi, theta))) looped over rows
 * (README.md:640-662).  ad::bind recognises sum(pow<2>(y - dot(X, w))) and routes it here. */
int fadb_lsq(size_t rows, size_t cols, const double* X, const double* y, const double* w, double* w_adj,
             double seed, unsigned flags, double* host_value);
int fadb_lsq_finish(size_t cols, const double* dev_partial, double* w_adj, double seed, unsigned flags,
                    double* dev_value);

/* ------------------------------------------------------------------ K6: 3-layer network
 * The expression of README.md:676-726 / example/ceres_3layer_neural_net/src.cpp:38-40
 * (2->3->3->1 with jump connection, squared error), summed over `batch` rows with the 25
 * parameter adjoints accumulated (the per-row loop of src.cpp:48-55 as one launch).
 * X is batch x 2 column-major, parm/grad are 25 device doubles. */
int fadb_mlp3(size_t batch, const double* X, const double* y, const double* parm,
              double* grad, unsigned flags, double* host_loss);
int fadb_mlp3_async(size_t batch, const double* X, const double* y, const double* parm,
                    double* grad, unsigned flags, double* dev_loss);

/* ------------------------------------------------------------------ generic expressions
 * The run-time image of a FastAD expression template.  The C++17 host layer (include/fastad)
 * lowers the static expression type into these calls at ad::bind() time; node constructors
 * mirror the reference's builders one to one and return a node id >= 0 (or a negative error):
 *   fadb_expr_var          VarView<T,Shape>(val, adj, rows, cols)     var_view.hpp:120-194
 *   fadb_expr_const_*      ad::constant / ad::constant_view           constant.hpp:158-200
 *   fadb_expr_unary        ad::sin ... ad::tanh, operator-            unary.hpp:299-330
 *   fadb_expr_binary       operator+ - * /                            binary.hpp:472-498
 *   fadb_expr_pow          ad::pow<n>                                 pow.hpp:205-231
 *   fadb_expr_sum[_iter]   ad::sum(expr) / ad::sum(begin,end,f)       sum.hpp:216-261
 *   fadb_expr_prod[_iter]  ad::prod(expr) / ad::prod(begin,end,f)     prod.hpp:259-306
 *   fadb_expr_dot          ad::dot                                    dot.hpp:133-163
 *   fadb_expr_normal       ad::normal_adj_log_pdf                     stat/normal.hpp:785-797
 *   fadb_expr_eq           placeholder `w = expr`                     var_view.hpp:54-63, eq.hpp
 *   fadb_expr_glue         operator,                                  glue.hpp:124-139
 * Scalar constant operands are folded eagerly on the host like the reference does
 * (unary.hpp:182-184, binary.hpp:251-256).  Leaf pointers are DEVICE pointers.
 *   fadb_expr_plan         host-only: cut into fused stages (no CUDA call); cache sizes play the
 *                          role of bind_cache_size() (bind.hpp:28) -- only stage boundaries count
 *   fadb_expr_bind         ad::bind: plan + allocate boundary buffers in HBM   bind.hpp:24-33
 *   fadb_expr_feval / _beval / _autodiff   ad::evaluate / evaluate_adj / autodiff   eval.hpp:18-153
 * seed: scalar `seed` broadcast, or `seed_dev` (device array of the root's size) when non-NULL.
 * host_value receives the root value (root-size doubles) when non-NULL. */
typedef struct fadb_expr fadb_expr;
int fadb_expr_create(fadb_expr** out);
void fadb_expr_destroy(fadb_expr*);
int fadb_expr_var(fadb_expr*, int shape, size_t rows, size_t cols, double* val, double* adj);
int fadb_expr_const_scalar(fadb_expr*, double c);
int fadb_expr_const_view(fadb_expr*, int shape, size_t rows, size_t cols, const double* val);
int fadb_expr_rebind(fadb_expr*, int leaf, double* val, double* adj);
int fadb_expr_unary(fadb_expr*, int op, int child);
int fadb_expr_binary(fadb_expr*, int op, int lhs, int rhs);
int fadb_expr_pow(fadb_expr*, long n, int child);
int fadb_expr_sum(fadb_expr*, int child);
int fadb_expr_sum_iter(fadb_expr*, int k, const int* children);
int fadb_expr_prod(fadb_expr*, int child);
int fadb_expr_prod_iter(fadb_expr*, int k, const int* children);
int fadb_expr_dot(fadb_expr*, int lhs, int rhs);
int fadb_expr_normal(fadb_expr*, int x, int mu, int sigma);
int fadb_expr_eq(fadb_expr*, int var, int child);
int fadb_expr_glue(fadb_expr*, int lhs, int rhs);
/* SURVEY.md 8f "next" rows:
 *   fadb_expr_norm       ad::norm(expr), squared L2 / Frobenius norm          norm.hpp:24-108
 *   fadb_expr_transpose  ad::transpose(expr)                                 transpose.hpp:17-81
 *   fadb_expr_if_else    ad::if_else(cond, a, b), scalar condition           if_else.hpp:29-165
 *   fadb_expr_opeq       w += / -= / *= / /= expr (op = FADB_ADD..FADB_DIV)   eq.hpp:194-413
 * (ad::for_each is a chain of fadb_expr_glue, for_each.hpp:19-131.) */
/* the other diagonal adjusted log-pdfs (reverse/stat/{cauchy,uniform,bernoulli}.hpp);
 * operands are scalars or vectors of one length: cauchy(x, loc, scale), uniform(x, min, max),
 * bernoulli(x, p) (c ignored).  Out-of-range parameters give -inf (bernoulli: see bernoulli.hpp:127-133)
 * and no adjoint update. */
enum { FADB_CAUCHY = 0, FADB_UNIFORM = 1, FADB_BERNOULLI = 2 };
int fadb_expr_logpdf(fadb_expr*, int dist, int a, int b, int c);
/* DetNode / LogDetNode (det.hpp:25-92, log_det.hpp:25-92) and WishartAdjLogPDFNode (wishart.hpp:92-231) */
int fadb_expr_det(fadb_expr*, int policy, int take_log, int child);
int fadb_expr_wishart(fadb_expr*, int x, int v, double df);
int fadb_expr_norm(fadb_expr*, int child);
int fadb_expr_transpose(fadb_expr*, int child);
int fadb_expr_if_else(fadb_expr*, int cond, int if_node, int else_node);
int fadb_expr_opeq(fadb_expr*, int op, int var, int child);
int fadb_expr_node_size(fadb_expr*, int node, size_t* rows, size_t* cols);
int fadb_expr_is_const(fadb_expr*, int node);
int fadb_expr_plan(fadb_expr*, int root, size_t out_cache_size[2]);
int fadb_expr_describe(fadb_expr*, char* buf, size_t cap);
int fadb_expr_bind(fadb_expr*, int root, unsigned flags, size_t out_cache_size[2]);
int fadb_expr_feval(fadb_expr*, double* host_value);
int fadb_expr_beval(fadb_expr*, double seed, const double* seed_dev);
int fadb_expr_autodiff(fadb_expr*, double seed, const double* seed_dev, double* host_value);
int fadb_expr_value_ptr(fadb_expr*, const double** dev_value, size_t* size);

/* ---- the one exchange of a sharded reduction, over NVLink peer memory (csrc/p2p.cu) ------------
 * element-range shards exchange 1-66 partial sums per evaluation.  Every rank (one process
 * per GPU of one node) owns a mailbox in its HBM; fadb_p2p_create returns its CUDA IPC handle, the caller
 * all-gathers the handles with whatever it has (torch.distributed, MPI) and passes them to
 * fadb_p2p_connect.  fadb_p2p_allreduce then sums dev_buf[0..n), n <= 128, over the ranks in place with ONE
 * small kernel per rank: peer stores + system-scope fence + flags, a bounded wait, a rank-ordered sum (all
 * ranks get identical bits).  Asynchronous on the library stream; no NCCL, no host synchronisation. */
int fadb_p2p_create(int world, int rank, void* out_handle64);
int fadb_p2p_connect(const void* all_handles /* world x 64 bytes, rank order */);
/* ONE process driving several GPUs: rank r = devices[r], mailboxes connected through plain peer access (no IPC, no handle
 * exchange).  The caller then selects a device and issues that rank's calls; a pull is enqueued after every rank's push. */
int fadb_p2p_create_local(int world, const int* devices);
int fadb_p2p_allreduce(double* dev_buf, int n);
/* Failure mode: a wait that exceeds FADB_P2P_TIMEOUT_MS (default 30000) -- dead peer, mismatched rank count -- sets a STICKY
 * error: that exchange and every later one deliver NaN (never a plausible partial sum), and every later host entry of the
 * exchange (fadb_p2p_allreduce, *_push, *_pull, *_allreduce_async) returns FADB_ERR_STATE until fadb_p2p_destroy +
 * fadb_p2p_create.  fadb_p2p_status synchronises the stream and returns the failed sequence number (0 = healthy). */
int fadb_p2p_status(unsigned long long* out_error_seq);
/* the caller barriers all ranks between their last exchange and this call (peers store into this mailbox) */
int fadb_p2p_destroy(void);
/* the normal log-pdf and regression passes with the exchange fused in: the producing kernel's last CTA pushes {sum r^2, X^T r} into every
 * rank's mailbox, the finish kernel waits, reduces in rank order and forms value + adjoints (cols <= 127) */
int fadb_normal_lpdf_partial_push(int shape_code, size_t n, const double* x, const double* mu, const double* sigma,
                                  double* x_adj, double* mu_adj, double* sigma_adj, double seed, unsigned flags,
                                  const double* dev_invalid_in, double* dev_partial4);
int fadb_normal_lpdf_finish_pull(int shape_code, size_t n_total, const double* x, const double* mu, const double* sigma,
                                 double* x_adj, double* mu_adj, double* sigma_adj, double seed, unsigned flags,
                                 double* host_value);
int fadb_linreg_partial_push(size_t rows, size_t cols, const double* X, const double* y, const double* w, double* dev_partial);
int fadb_linreg_finish_pull(size_t rows_total, size_t cols, const double* sigma, double* w_adj, double* sigma_adj, double seed,
                            unsigned flags, double* host_value);
/* sum(f(x)) (sum.hpp:156-173) and the network (example/ceres_3layer_neural_net/src.cpp:48-55) over a shard with the WHOLE all-reduce
 * fused into the one launch: the last CTA pushes its 1 / 26 sums into every rank's mailbox, waits for the other ranks and writes the
 * rank-ordered totals (dev_value; grad[25] and dev_loss) - same arguments as fadb_sum_unary_async / fadb_mlp3_async, identical bits on
 * every rank, and with one rank the bits of the plain calls.  Vector adjoints stay shard-local (the seed is known up front). */
int fadb_sum_unary_allreduce_async(int op, size_t n, const double* x, double* x_adj, double seed, unsigned flags, double* dev_value);
int fadb_mlp3_allreduce_async(size_t batch, const double* X, const double* y, const double* parm, double* grad, unsigned flags,
                              double* dev_loss);

/* ---- bind-time kernel specialisation (csrc/jit.cu) ------------------------------------------
 * The reference's expression templates hand every expression its own fused evaluator at compile
 * time (bind.hpp / eval.hpp instantiate feval/beval per expression type).  For host code built
 * without nvcc the library does the same at bind time: the stage program the planner produced is
 * spliced, as compile-time constants, into the stage-kernel source embedded in this library and
 * compiled for sm_100a with NVRTC on first use (cached per device and program).  Stages NVRTC
 * cannot take (if_else control flow, programs over 64 nodes) and machines without libnvrtc run
 * the same source as a tape interpreter.  On by default; FADB_NO_JIT=1 or fadb_jit_enable(0)
 * select the interpreter.
 *   fadb_jit_enable         switch at run time, returns the previous setting
 *   fadb_jit_stats          out = {kernels compiled, cache hits, compile failures, launches}
 *   fadb_jit_diagnostics    why the layer is unavailable / the last NVRTC log
 *   fadb_expr_jit_program   the generated constants of one planned stage (host only)
 *   fadb_jit_compile_only   NVRTC-compile such a text to a cubin; needs no device (build check) */
int fadb_jit_enable(int on);
int fadb_jit_stats(unsigned long long out[4]);
const char* fadb_jit_diagnostics(void);
int fadb_expr_jit_program(fadb_expr*, int stage, int mode, char* buf, size_t cap);
/* sum(f(dot(X, w), per-row constants, scalars)) with X constant and <= 64 columns is evaluated in ONE pass over X (the map runs
 * inside the tile kernel, csrc/glm_kernel.cuh; `fast=6` in fadb_expr_describe) instead of gemv -> map -> gemv^T.
 * fadb_glm_enable(0) turns the route off (returns the previous setting); fadb_expr_jit_program(.., mode 6) is its program. */
int fadb_glm_enable(int on);
/* A root that is a SUM of terms (a posterior: log-likelihood + log-priors; +, -, unary minus, * or / by constants over the term
 * values) hands every term a seed that is known before anything is evaluated: each term is then swept as a root of its own --
 * forward and adjoint sweep fused in one pass, on the single-pass kernels where its shape allows (`additive root` in
 * fadb_expr_describe) -- instead of a forward pass before and a backward pass after the root stage.
 * fadb_additive_root_enable(0) restores the one-block order (returns the previous setting). */
int fadb_additive_root_enable(int on);
/* Per-element stages with placeholder stores (`c = expr` over vectors, reverse/core/eq.hpp:68-125; the vectorised
 * example/black_scholes.cpp) run one element per thread; their specialised kernel fetches every operand, old adjoint and
 * per-element seed one trip ahead with cp.async into a per-thread shared-memory column instead of ten serialised global
 * loads per element (csrc/stage_kernel.cuh, JIT_RING).  Taken when no written operand of the stage overlaps another operand;
 * fadb_jit_ring_enable(0) keeps the plain kernel (returns the previous setting); fadb_expr_jit_program(.., mode | 8) is
 * its program text. */
int fadb_jit_ring_enable(int on);
int fadb_jit_compile_only(const char* program, size_t* cubin_bytes, char* log, size_t log_cap);

#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* FASTAD_B200_H */
