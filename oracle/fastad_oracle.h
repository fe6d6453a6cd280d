/*
 * fastad_oracle.h — TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * CPU restatement (plain loops, no Eigen, single-threaded per graph) of the reverse-mode
 * hot path of JamesYang007/FastAD (the reference lives at /root/reference; it needs
 * Eigen 3.3, which is absent from this image, so it cannot be compiled here — see
 * DESIGN.md "Oracle").  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Parity is PINNED: tests/test_oracle_goldens.py checks this restatement against the
 * reference's own golden vectors (test/reverse/stat/normal_unittest.cpp:105-297,
 * test/reverse/core/unary_unittest.cpp:320-324, bind_unittest.cpp:13-35,
 * eval_unittest.cpp:32-48, test/integration/node_inttest.cpp, README.md:582-662,
 * test/util/GenTestData.nb:691-700).
 *
 * The oracle is a *run-time* node graph (the reference is a compile-time expression
 * template); each node evaluates array-at-a-time and materialises its value and
 * adjoint caches exactly as the reference's nodes do, in the same child order.
 */
#ifndef FASTAD_ORACLE_H
#define FASTAD_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_graph orc_graph;

/* shapes follow ad::scl / ad::vec / ad::mat (util/shape_traits.hpp:7-9) */
enum { ORC_SCL = 0, ORC_VEC = 1, ORC_MAT = 2 };

/* unary functors, reverse/core/unary.hpp:190-296 */
enum {
    ORC_NEG = 0, ORC_SIN, ORC_COS, ORC_TAN, ORC_ASIN, ORC_ACOS, ORC_ATAN, ORC_EXP, ORC_LOG,
    ORC_SQRT, ORC_ERF, ORC_SIGMOID, ORC_SINH, ORC_COSH, ORC_TANH,
    ORC_MOCK2X = 100 /* MockUnary f(x)=2x, test/testutil/base_fixture.hpp:9-22 */
};
/* binary functors, reverse/core/binary.hpp:265-338 */
enum { ORC_ADD = 0, ORC_SUB, ORC_MUL, ORC_DIV,
       /* comparison / logical functors, binary.hpp:347-470: value 0/1, beval is a no-op */
       ORC_LT, ORC_LE, ORC_GT, ORC_GE, ORC_EQ, ORC_NE, ORC_AND, ORC_OR,
       ORC_MOCKB = 100 /* MockBinary f(x,y)=x-2y, base_fixture.hpp:24-46 */ };

orc_graph* orc_new(void);
void orc_free(orc_graph*);

/* leaves: pointers are HOST pointers owned by the caller (column-major for mat) */
int orc_var(orc_graph*, int shape, size_t rows, size_t cols, double* val, double* adj);
int orc_const_scalar(orc_graph*, double c);
int orc_const_view(orc_graph*, int shape, size_t rows, size_t cols, const double* val);
/* re-point a leaf (ValueAdjView::bind, value_adj_view.hpp:47-52) */
void orc_rebind(orc_graph*, int leaf, double* val, double* adj, size_t rows, size_t cols);

int orc_unary(orc_graph*, int op, int child);
int orc_binary(orc_graph*, int op, int lhs, int rhs);
int orc_pow(orc_graph*, long n, int child);
int orc_sum(orc_graph*, int child);                       /* SumElemNode  */
int orc_sum_iter(orc_graph*, int k, const int* children); /* SumIterNode  */
int orc_prod(orc_graph*, int child);                      /* ProdElemNode */
int orc_prod_iter(orc_graph*, int k, const int* children);/* ProdIterNode */
int orc_dot(orc_graph*, int lhs, int rhs);
int orc_normal(orc_graph*, int x, int mu, int sigma);     /* normal_adj_log_pdf */
int orc_eq(orc_graph*, int var, int child);               /* placeholder w = expr */
int orc_glue(orc_graph*, int lhs, int rhs);               /* (lhs, rhs) */
int orc_norm(orc_graph*, int child);                      /* NormNode, norm.hpp:24-108 */
int orc_transpose(orc_graph*, int child);                 /* TransposeNode, transpose.hpp:17-81 */
int orc_if_else(orc_graph*, int cond, int a, int b);      /* IfElseNode, if_else.hpp:29-165 */
int orc_opeq(orc_graph*, int op, int var, int child);     /* OpEqNode, eq.hpp:194-413 */
/* other diagonal adjusted log-pdfs (reverse/stat): dist = ORC_CAUCHY (x, loc, scale; cauchy.hpp),
 * ORC_UNIFORM (x, min, max; uniform.hpp), ORC_BERNOULLI (x, p, unused; bernoulli.hpp) */
enum { ORC_CAUCHY = 0, ORC_UNIFORM = 1, ORC_BERNOULLI = 2 };
int orc_logpdf(orc_graph*, int dist, int a, int b, int c);

/* DetNode / LogDetNode (reverse/core/det.hpp:25-92, log_det.hpp): policy 0 FullPivLU (default),
 * 1 LDLT, 2 LLT (det.hpp:97-202); take_log = 1 for log_det */
int orc_det(orc_graph*, int policy, int take_log, int child);
/* WishartAdjLogPDFNode (reverse/stat/wishart.hpp:92-231): x, v square mat, df a constant */
int orc_wishart(orc_graph*, int x, int v, double df);

/* ad::bind: size query + allocate + post-order bind. out2 = {#val, #adj} as SizePack */
void orc_bind(orc_graph*, int root, size_t out2[2]);
/* ad::evaluate / evaluate_adj / autodiff (eval.hpp:18-153). Return pointer to root value. */
const double* orc_feval(orc_graph*, int root);
void orc_beval(orc_graph*, int root, double seed_scalar, const double* seed_array);
size_t orc_size(orc_graph*, int node);
int orc_is_const(orc_graph*, int node);

/* ---- whole workloads built on the graph above (bench cpu_baseline + parity) ---- */

/* C1: example/black_scholes.cpp:16-39 over vec leaves, S/sigma/r variables.
 * out = 8 SoA arrays of n: call, put, dcall_dS, dcall_dsigma, dcall_dr, dput_dS,
 * dput_dsigma, dput_dr.  nthreads>1 splits the batch in contiguous slices, each thread
 * owning a private graph evaluated over L2-sized chunks. */
void orc_black_scholes(size_t n, const double* S, const double* K, const double* sigma,
                       const double* tau, const double* r, double* const out[8],
                       int nthreads);

/* C2..C3 helpers: sum(f(x)) value + adjoint accumulate, threaded element ranges */
double orc_sum_unary(int op, size_t n, const double* x, double* x_adj, double seed,
                     int nthreads);
/* normal log-pdf with leaves only. shape_code: bit0 mu is vec, bit1 sigma is vec.
 * const_mask: bit0 x const, bit1 mu const, bit2 sigma const. adj may be NULL if const. */
double orc_normal_lpdf(int shape_code, int const_mask, size_t n, const double* x,
                       const double* mu, const double* sigma, double* x_adj,
                       double* mu_adj, double* sigma_adj, double seed, int nthreads);
/* C4: normal_adj_log_pdf(y, dot(X,w), sigma), X rows x cols column-major const */
double orc_linreg(size_t rows, size_t cols, const double* X, const double* y,
                  const double* w, double* w_adj, const double* sigma, double* sigma_adj,
                  double seed, int nthreads);
/* C5: example/ceres_3layer_neural_net/src.cpp:17-57. X is batch x 2 column-major,
 * parm/grad 25 doubles. Returns summed loss; grad accumulates. */
double orc_mlp3(size_t batch, const double* X, const double* y, const double* parm,
                double* grad, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
