/*
 * fastad_oracle.cpp — TEST INFRASTRUCTURE ONLY (see fastad_oracle.h).
 *
 * Eigen-free CPU restatement of FastAD's reverse-mode path.  Every routine cites the
 * reference file:line (relative to /root/reference/include/fastad_bits/) it follows.
 * The reference is a compile-time expression template; this is a run-time node graph that
 * performs the same array-at-a-time passes with the same materialised caches, in the same
 * child order.  Build: see oracle/Makefile (parity build: -O2 -ffp-contract=off).
 */
#include "fastad_oracle.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace {

enum Kind { K_VAR, K_CONST, K_UNARY, K_BINARY, K_POW, K_SUM, K_SUM_ITER, K_PROD,
            K_PROD_ITER, K_DOT, K_NORMAL, K_EQ, K_GLUE, K_NORM, K_TRANSPOSE, K_IF_ELSE, K_OPEQ, K_LOGPDF, K_DET, K_WISHART };

/* A seed is either a scalar (broadcast) or an array, as in the reference where the
 * seed type is a template parameter (var_view.hpp:91-94). */
struct Seed {
    const double* p;
    double s;
    double at(size_t i) const { return p ? p[i] : s; }
};

struct Node {
    Kind kind;
    int op = 0;
    long pw = 0;
    int shape = ORC_SCL;
    size_t rows = 1, cols = 1;
    std::vector<int> ch;
    double* val = nullptr;   /* bound value storage  (ValueView)    */
    double* adj = nullptr;   /* bound adjoint storage (ValueAdjView) */
    std::vector<double> cval;          /* owned constant payload (Constant<T,Shape>) */
    std::vector<double> tmp0, tmp1;    /* child-seed staging (Eigen lazy exprs in the ref) */
    std::vector<double> extra;         /* ProdElemNode adj_cache_ (prod.hpp:244)      */
    /* NormalAdjLogPDFNode cached statistics (normal.hpp:292-300, 487-493) */
    double log_sigma = 0, z_sq = 0, x_mean = 0, x_var = 0;
    double sq_term = 0, lin_term = 0, const_term = 0;
    bool pos_def = false;
    size_t size() const { return rows * cols; }
};

inline double unary_f(int op, double x)
{   /* fmap bodies, unary.hpp:190-296 */
    switch (op) {
    case ORC_NEG: return -x;
    case ORC_SIN: return std::sin(x);
    case ORC_COS: return std::cos(x);
    case ORC_TAN: return std::tan(x);
    case ORC_ASIN: return std::asin(x);
    case ORC_ACOS: return std::acos(x);
    case ORC_ATAN: return std::atan(x);
    case ORC_EXP: return std::exp(x);
    case ORC_LOG: return std::log(x);
    case ORC_SQRT: return std::sqrt(x);
    case ORC_ERF: return std::erf(x);
    case ORC_SIGMOID: return 1 / (1 + std::exp(-x));
    case ORC_SINH: return std::sinh(x);
    case ORC_COSH: return std::cosh(x);
    case ORC_TANH: return std::tanh(x);
    case ORC_MOCK2X: return 2 * x;
    }
    return std::numeric_limits<double>::quiet_NaN();
}

inline double unary_b(int op, double seed, double x, double f)
{   /* bmap bodies, unary.hpp:190-296 */
    switch (op) {
    case ORC_NEG: return -seed;
    case ORC_SIN: return seed * std::cos(x);
    case ORC_COS: return -seed * std::sin(x);
    case ORC_TAN: { double t = std::cos(x); return seed / (t * t); }
    case ORC_ASIN: return seed / std::sqrt(1. - x * x);
    case ORC_ACOS: return -(seed / std::sqrt(1. - x * x));
    case ORC_ATAN: return seed / (1. + x * x);
    case ORC_EXP: return seed * f;
    case ORC_LOG: return seed / x;
    case ORC_SQRT: return 0.5 * seed / f;
    case ORC_ERF: return 1.1283791670955126 * seed * std::exp(-x * x);
    case ORC_SIGMOID: return seed * std::exp(-x) / ((std::exp(-x) + 1) * (std::exp(-x) + 1));
    case ORC_SINH: return seed * std::cosh(x);
    case ORC_COSH: return seed * std::sinh(x);
    case ORC_TANH: return seed * (1 - f * f);
    case ORC_MOCK2X: return seed * 2;
    }
    return std::numeric_limits<double>::quiet_NaN();
}

inline double binary_f(int op, double x, double y)
{   /* binary.hpp:265-338 */
    switch (op) {
    case ORC_ADD: return x + y;
    case ORC_SUB: return x - y;
    case ORC_MUL: return x * y;
    case ORC_DIV: return x / y;
    case ORC_MOCKB: return x - 2 * y;
    case ORC_LT: return x < y; case ORC_LE: return x <= y; case ORC_GT: return x > y; case ORC_GE: return x >= y;
    case ORC_EQ: return x == y; case ORC_NE: return x != y;
    case ORC_AND: return x != 0 && y != 0; case ORC_OR: return x != 0 || y != 0;
    }
    return std::numeric_limits<double>::quiet_NaN();
}
inline bool is_comparison(int op) { return op >= ORC_LT && op <= ORC_OR; }
inline double binary_bl(int op, double s, double x, double y, double f)
{
    (void)x; (void)f;
    switch (op) {
    case ORC_ADD: return s;
    case ORC_SUB: return s;
    case ORC_MUL: return s * y;
    case ORC_DIV: return s / y;
    case ORC_MOCKB: return s;
    }
    return 0;
}
inline double binary_br(int op, double s, double x, double y, double f)
{
    switch (op) {
    case ORC_ADD: return s;
    case ORC_SUB: return -s;
    case ORC_MUL: return s * x;
    case ORC_DIV: return -s * f / y;
    case ORC_MOCKB: return -2. * s;
    }
    return 0;
}

/* PowFunc, pow.hpp:20-56 (repeated squaring, used for scalars) */
double pow_func(long n, double base)
{
    if (n < 0) {
        return (base == 0) ? std::numeric_limits<double>::infinity()
                           : pow_func(-n, 1. / base);
    }
    if (n == 0) return 1.;
    if (n % 2 == 0) { double t = pow_func(n / 2, base); return t * t; }
    return base * pow_func(n - 1, base);
}

} // namespace

struct orc_graph {
    std::vector<Node> nodes;
    std::vector<double> val_cache, adj_cache;   /* ExprBind caches, bind.hpp:38-40 */

    int add(Node&& n) { nodes.push_back(std::move(n)); return (int)nodes.size() - 1; }
    bool is_const(int id) const { return nodes[id].kind == K_CONST; }

    /* ---- bind_cache_size protocol (size_pack.hpp:9; e.g. unary.hpp:107-116) ---- */
    void single_size(int id, size_t out[2]) const
    {
        const Node& n = nodes[id];
        switch (n.kind) {
        case K_VAR: case K_CONST: case K_EQ: case K_GLUE: case K_IF_ELSE: out[0] = out[1] = 0; break;
        case K_BINARY:
            out[0] = n.size(); out[1] = is_comparison(n.op) ? 0 : n.size(); break;   /* binary.hpp:166-174 */
        case K_UNARY: case K_POW: case K_DOT: case K_SUM_ITER: case K_PROD_ITER: case K_TRANSPOSE:
        case K_OPEQ:   /* OpEqNode's cache_ holds the previous lhs value + adjoint, eq.hpp:289-292 */
            out[0] = n.size(); out[1] = n.size(); break;
        case K_SUM: case K_NORMAL: case K_NORM: case K_LOGPDF: case K_DET: case K_WISHART:
            out[0] = n.size(); out[1] = 0; break; /* sum.hpp:196-199, det.hpp:81-84, wishart.hpp single_bind_cache_size */
        case K_PROD: out[0] = n.size(); out[1] = nodes[n.ch[0]].size(); break; /* prod.hpp:236-239 */
        }
    }
    void total_size(int id, size_t out[2]) const
    {
        const Node& n = nodes[id];
        size_t s[2];
        single_size(id, s);
        out[0] = s[0]; out[1] = s[1];
        for (int c : n.ch) {
            if ((n.kind == K_EQ || n.kind == K_OPEQ) && c == n.ch[0]) continue;  /* the placeholder itself */
            size_t t[2];
            total_size(c, t);
            out[0] += t[0]; out[1] += t[1];
        }
        if (n.kind == K_EQ) {  /* eq.hpp:141-145: gives the root's slot back */
            size_t r[2];
            single_size(n.ch[1], r);
            out[0] -= r[0]; out[1] -= r[1];
        }
    }

    /* ---- post-order bind (bind.hpp:24-33 + each node's bind_cache) ---- */
    void bind_node(int id, double*& v, double*& a)
    {
        Node& n = nodes[id];
        switch (n.kind) {
        case K_VAR: case K_CONST: return;            /* var_view.hpp:97-102 */
        case K_EQ: {                                 /* eq.hpp:119-136 */
            Node& w = nodes[n.ch[0]];
            n.val = w.val; n.adj = w.adj;
            bind_node(n.ch[1], v, a);
            size_t r[2];
            single_size(n.ch[1], r);
            v -= r[0]; a -= r[1];
            Node& e = nodes[n.ch[1]];
            if (e.kind != K_VAR && e.kind != K_CONST) { e.val = w.val; if (r[1]) e.adj = w.adj; }
            return;
        }
        case K_IF_ELSE:                              /* if_else.hpp:80-85: no storage of its own */
            for (int c : n.ch) bind_node(c, v, a);
            return;
        case K_OPEQ: {                               /* eq.hpp:273-279 */
            Node& w = nodes[n.ch[0]];
            n.val = w.val; n.adj = w.adj;
            bind_node(n.ch[1], v, a);
            n.extra.assign(2 * n.size(), 0.);        /* cache_: previous value | its adjoint */
            v += n.size(); a += n.size();
            return;
        }
        case K_GLUE: {                               /* glue.hpp:94-100 */
            bind_node(n.ch[0], v, a);
            bind_node(n.ch[1], v, a);
            n.val = nodes[n.ch[1]].val; n.adj = nodes[n.ch[1]].adj;
            return;
        }
        default: break;
        }
        for (int c : n.ch) bind_node(c, v, a);
        if (n.kind == K_PROD) {                      /* prod.hpp:224-233 */
            n.extra.assign(nodes[n.ch[0]].size(), 0.);
            a += nodes[n.ch[0]].size();
        }
        size_t s[2];
        single_size(id, s);
        n.val = v; v += s[0];
        if (n.kind == K_SUM || n.kind == K_NORMAL || n.kind == K_PROD || n.kind == K_NORM || n.kind == K_LOGPDF ||
            n.kind == K_DET || n.kind == K_WISHART ||
            (n.kind == K_BINARY && is_comparison(n.op))) n.adj = nullptr;
        else { n.adj = a; a += n.size(); }
    }

    void bind(int root, size_t out[2])
    {
        total_size(root, out);
        val_cache.assign(out[0] + 1, 0.);
        adj_cache.assign(out[1] + 1, 0.);
        double* v = val_cache.data();
        double* a = adj_cache.data();
        bind_node(root, v, a);
    }

    void normal_update_cache(Node& n)
    {
        const Node& sg = nodes[n.ch[2]];
        if (sg.shape == ORC_SCL) {
            n.log_sigma = std::log(sg.val[0]);          /* normal.hpp:173-175 */
        } else {                                       /* normal.hpp:480-485, 567-574 */
            bool ok = true;
            for (size_t i = 0; i < sg.size(); ++i) ok = ok && (sg.val[i] > 0);
            n.pos_def = ok;
            if (ok) {
                double acc = 0;
                for (size_t i = 0; i < sg.size(); ++i) acc += std::log(sg.val[i]);
                n.log_sigma = acc;
            }
        }
    }

    /* NormalAdjLogPDFNode ctor-time precomputation (normal.hpp:210-222, 412-427) */
    void normal_ctor(Node& n)
    {
        const Node& x = nodes[n.ch[0]];
        const Node& sg = nodes[n.ch[2]];
        const Node& mu = nodes[n.ch[1]];
        bool xc = x.kind == K_CONST, sc = sg.kind == K_CONST;
        if (is_dense(sg, x)) return;
        if (sc) normal_update_cache(n);
        if (x.shape != ORC_SCL && mu.shape == ORC_SCL && sg.shape == ORC_SCL && xc) {
            size_t N = x.size();
            double m = 0;
            for (size_t i = 0; i < N; ++i) m += x.val[i];
            n.x_mean = m / (double)N;
            double v = 0;
            for (size_t i = 0; i < N; ++i) { double d = x.val[i] - n.x_mean; v += d * d; }
            n.x_var = v;
        }
        if (x.shape != ORC_SCL && mu.shape == ORC_SCL && sg.shape != ORC_SCL && xc && sc) {
            size_t N = x.size();
            double a = 0, b = 0, c = 0;
            for (size_t i = 0; i < N; ++i) {
                double xs = x.val[i] / sg.val[i];
                a += xs * xs;
                b += x.val[i] / (sg.val[i] * sg.val[i]);
                double is = 1. / sg.val[i];
                c += is * is;
            }
            n.sq_term = a; n.lin_term = b; n.const_term = c;
        }
    }

    /* ---- forward: expr.feval() ---- */
    const double* feval(int id)
    {
        Node& n = nodes[id];
        const size_t N = n.size();
        switch (n.kind) {
        case K_VAR: case K_CONST:
            return n.val;                                /* var_view.hpp:69, constant.hpp:100 */
        case K_UNARY: {                                  /* unary.hpp:63-68 */
            const double* x = feval(n.ch[0]);
            for (size_t i = 0; i < N; ++i) n.val[i] = unary_f(n.op, x[i]);
            return n.val;
        }
        case K_BINARY: {                                 /* binary.hpp:100-106 */
            const double* l = feval(n.ch[0]);
            const double* r = feval(n.ch[1]);
            const bool ls = nodes[n.ch[0]].size() == 1 && N > 1;
            const bool rs = nodes[n.ch[1]].size() == 1 && N > 1;
            for (size_t i = 0; i < N; ++i)
                n.val[i] = binary_f(n.op, l[ls ? 0 : i], r[rs ? 0 : i]);
            return n.val;
        }
        case K_POW: {                                    /* pow.hpp:91-99 */
            const double* x = feval(n.ch[0]);
            if (n.shape == ORC_SCL) n.val[0] = pow_func(n.pw, x[0]);
            else for (size_t i = 0; i < N; ++i) n.val[i] = std::pow(x[i], (double)n.pw);
            return n.val;
        }
        case K_SUM: {                                    /* sum.hpp:156-164 */
            const double* x = feval(n.ch[0]);
            const size_t M = nodes[n.ch[0]].size();
            double acc = 0;
            for (size_t i = 0; i < M; ++i) acc += x[i];
            n.val[0] = acc;
            return n.val;
        }
        case K_SUM_ITER: {                               /* sum.hpp:57-64 */
            for (size_t i = 0; i < N; ++i) n.val[i] = 0;
            for (int c : n.ch) {
                const double* x = feval(c);
                for (size_t i = 0; i < N; ++i) n.val[i] += x[i];
            }
            return n.val;
        }
        case K_PROD: {                                   /* prod.hpp:172-180 */
            const double* x = feval(n.ch[0]);
            const size_t M = nodes[n.ch[0]].size();
            double acc = 1;
            for (size_t i = 0; i < M; ++i) acc *= x[i];
            n.val[0] = acc;
            return n.val;
        }
        case K_PROD_ITER: {                              /* prod.hpp:55-62 */
            for (size_t i = 0; i < N; ++i) n.val[i] = 1;
            for (int c : n.ch) {
                const double* x = feval(c);
                for (size_t i = 0; i < N; ++i) n.val[i] *= x[i];
            }
            return n.val;
        }
        case K_DOT: {                                    /* dot.hpp:89-94, column-major */
            const double* L = feval(n.ch[0]);
            const double* R = feval(n.ch[1]);
            const size_t m = n.rows, p = n.cols, k = nodes[n.ch[0]].cols;
            for (size_t j = 0; j < p; ++j) {
                for (size_t i = 0; i < m; ++i) n.val[i + j * m] = 0;
                for (size_t q = 0; q < k; ++q) {
                    const double rq = R[q + j * k];
                    const double* Lq = L + q * m;
                    for (size_t i = 0; i < m; ++i) n.val[i + j * m] += Lq[i] * rq;
                }
            }
            return n.val;
        }
        case K_EQ: {                                     /* eq.hpp:89-92 */
            const double* e = feval(n.ch[1]);
            Node& w = nodes[n.ch[0]];
            if (e != w.val) for (size_t i = 0; i < N; ++i) w.val[i] = e[i];
            return n.val;
        }
        case K_GLUE: {                                   /* glue.hpp:67-71 */
            feval(n.ch[0]);
            return feval(n.ch[1]);
        }
        case K_NORMAL: return normal_feval(n);
        case K_LOGPDF: return logpdf_feval(n);
        case K_DET: return det_feval(n);
        case K_WISHART: return wishart_feval(n);
        case K_NORM: {                                   /* norm.hpp:47-51 */
            const double* x = feval(n.ch[0]);
            const size_t M = nodes[n.ch[0]].size();
            double acc = 0;
            for (size_t i = 0; i < M; ++i) acc += x[i] * x[i];
            n.val[0] = acc;
            return n.val;
        }
        case K_TRANSPOSE: {                              /* transpose.hpp:32-35 */
            const double* x = feval(n.ch[0]);
            const size_t m = nodes[n.ch[0]].rows, k = nodes[n.ch[0]].cols;
            for (size_t j = 0; j < k; ++j)
                for (size_t i = 0; i < m; ++i) n.val[j + i * k] = x[i + j * m];
            return n.val;
        }
        case K_IF_ELSE: {                                /* if_else.hpp:64-68 */
            const double c = feval(n.ch[0])[0];
            n.pw = c != 0 ? 1 : 2;
            n.val = const_cast<double*>(feval(n.ch[n.pw]));
            return n.val;
        }
        case K_OPEQ: {                                   /* eq.hpp:240-247 */
            Node& w = nodes[n.ch[0]];
            for (size_t i = 0; i < N; ++i) n.extra[i] = w.val[i];
            const double* e = feval(n.ch[1]);
            const bool es = nodes[n.ch[1]].size() == 1 && N > 1;
            for (size_t i = 0; i < N; ++i) w.val[i] = binary_f(n.op, w.val[i], e[es ? 0 : i]);
            return w.val;
        }
        }
        return nullptr;
    }


    /* ---- decompositions behind det / log_det / wishart.  Eigen 3.3.7 (absent here) supplies
     * FullPivLU, LDLT and LLT; these are plain-loop restatements of the same factorizations
     * (same pivoting rules), not of Eigen's blocking, so results agree to rounding. ---- */
    /* Eigen::LLT<Lower>: reads the lower triangle; returns false when a pivot is not positive.
     * prodL = det(matrixL()); inv = llt.solve(I). */
    static bool llt_decomp(size_t d, const double* A, double& prodL, std::vector<double>& inv)
    {
        std::vector<double> L(d * d, 0.);
        for (size_t j = 0; j < d; ++j) {
            double s = A[j + j * d];
            for (size_t k = 0; k < j; ++k) s -= L[j + k * d] * L[j + k * d];
            if (!(s > 0)) return false;
            const double ljj = std::sqrt(s);
            L[j + j * d] = ljj;
            for (size_t i = j + 1; i < d; ++i) {
                double t = A[i + j * d];
                for (size_t k = 0; k < j; ++k) t -= L[i + k * d] * L[j + k * d];
                L[i + j * d] = t / ljj;
            }
        }
        prodL = 1;
        for (size_t j = 0; j < d; ++j) prodL *= L[j + j * d];
        inv.assign(d * d, 0.);
        std::vector<double> y(d);
        for (size_t c = 0; c < d; ++c) {
            for (size_t i = 0; i < d; ++i) {
                double t = (i == c) ? 1. : 0.;
                for (size_t k = 0; k < i; ++k) t -= L[i + k * d] * y[k];
                y[i] = t / L[i + i * d];
            }
            for (size_t ii = d; ii-- > 0;) {
                double t = y[ii];
                for (size_t k = ii + 1; k < d; ++k) t -= L[k + ii * d] * inv[k + c * d];
                inv[ii + c * d] = t / L[ii + ii * d];
            }
        }
        return true;
    }
    /* Eigen::FullPivLU: complete pivoting (largest |a_ij| of the trailing block, first hit in
     * column-major order), determinant = sign * prod(diag U), isInvertible() = every pivot larger
     * than epsilon * n * |max pivot|, inverse() = solve(I). */
    static void fullpivlu_decomp(size_t d, const double* A, double& det, bool& invertible, std::vector<double>& inv)
    {
        std::vector<double> M(A, A + d * d);
        std::vector<size_t> rp(d), cp(d);
        for (size_t i = 0; i < d; ++i) rp[i] = cp[i] = i;
        int sign = 1;
        double maxpivot = 0;
        size_t nonzero = d;
        for (size_t k = 0; k < d; ++k) {
            size_t br = k, bc = k;
            double big = -1;
            for (size_t j = k; j < d; ++j)
                for (size_t i = k; i < d; ++i)
                    if (std::fabs(M[i + j * d]) > big) { big = std::fabs(M[i + j * d]); br = i; bc = j; }
            if (big == 0) { nonzero = k; break; }
            if (big > maxpivot) maxpivot = big;
            if (br != k) { for (size_t j = 0; j < d; ++j) std::swap(M[k + j * d], M[br + j * d]); std::swap(rp[k], rp[br]); sign = -sign; }
            if (bc != k) { for (size_t i = 0; i < d; ++i) std::swap(M[i + k * d], M[i + bc * d]); std::swap(cp[k], cp[bc]); sign = -sign; }
            for (size_t i = k + 1; i < d; ++i) M[i + k * d] /= M[k + k * d];
            for (size_t j = k + 1; j < d; ++j)
                for (size_t i = k + 1; i < d; ++i) M[i + j * d] -= M[i + k * d] * M[k + j * d];
        }
        det = sign;
        for (size_t k = 0; k < d; ++k) det *= M[k + k * d];
        const double thr = std::numeric_limits<double>::epsilon() * (double)d * maxpivot;
        size_t rank = 0;
        for (size_t k = 0; k < nonzero; ++k) rank += std::fabs(M[k + k * d]) > thr;
        invertible = rank == d;
        inv.assign(d * d, 0.);
        if (!invertible) return;
        std::vector<double> y(d);
        for (size_t c = 0; c < d; ++c) {               /* P A Q = L U  =>  A^-1 e_c = Q U^-1 L^-1 P e_c */
            for (size_t i = 0; i < d; ++i) y[i] = rp[i] == c ? 1. : 0.;
            for (size_t i = 0; i < d; ++i)
                for (size_t k = 0; k < i; ++k) y[i] -= M[i + k * d] * y[k];
            for (size_t ii = d; ii-- > 0;) {
                for (size_t k = ii + 1; k < d; ++k) y[ii] -= M[ii + k * d] * y[k];
                y[ii] /= M[ii + ii * d];
            }
            for (size_t i = 0; i < d; ++i) inv[cp[i] + c * d] = y[i];
        }
    }
    /* Eigen::LDLT<Lower>: reads the lower triangle, pivots on the largest |diagonal| entry;
     * vectorD().prod() is the determinant; solve(I) = P^T L^-T D^-1 L^-1 P. */
    static void ldlt_decomp(size_t d, const double* A, double& det, bool& success, std::vector<double>& inv)
    {
        std::vector<double> M(d * d);
        for (size_t j = 0; j < d; ++j)
            for (size_t i = 0; i < d; ++i) M[i + j * d] = i >= j ? A[i + j * d] : A[j + i * d];
        std::vector<size_t> perm(d);
        for (size_t i = 0; i < d; ++i) perm[i] = i;
        success = true;
        for (size_t k = 0; k < d; ++k) {
            size_t p = k;
            double big = -1;
            for (size_t i = k; i < d; ++i)
                if (std::fabs(M[i + i * d]) > big) { big = std::fabs(M[i + i * d]); p = i; }
            if (p != k) {
                for (size_t j = 0; j < d; ++j) std::swap(M[k + j * d], M[p + j * d]);
                for (size_t i = 0; i < d; ++i) std::swap(M[i + k * d], M[i + p * d]);
                std::swap(perm[k], perm[p]);
            }
            const double dk = M[k + k * d];
            if (dk == 0) {
                for (size_t i = k + 1; i < d; ++i) if (M[i + k * d] != 0) success = false;
                continue;
            }
            for (size_t i = k + 1; i < d; ++i) M[i + k * d] /= dk;
            for (size_t j = k + 1; j < d; ++j)
                for (size_t i = j; i < d; ++i) M[i + j * d] -= M[i + k * d] * dk * M[j + k * d];
            for (size_t j = k + 1; j < d; ++j)
                for (size_t i = k + 1; i < j; ++i) M[i + j * d] = M[j + i * d];
        }
        det = 1;
        for (size_t k = 0; k < d; ++k) det *= M[k + k * d];
        inv.assign(d * d, 0.);
        std::vector<double> y(d);
        for (size_t c = 0; c < d; ++c) {
            for (size_t i = 0; i < d; ++i) y[i] = perm[i] == c ? 1. : 0.;
            for (size_t i = 0; i < d; ++i)
                for (size_t k = 0; k < i; ++k) y[i] -= M[i + k * d] * y[k];
            for (size_t i = 0; i < d; ++i) y[i] = M[i + i * d] != 0 ? y[i] / M[i + i * d] : 0.;
            for (size_t ii = d; ii-- > 0;)
                for (size_t k = ii + 1; k < d; ++k) y[ii] -= M[k + ii * d] * y[k];
            for (size_t i = 0; i < d; ++i) inv[perm[i] + c * d] = y[i];
        }
    }

    /* DetNode / LogDetNode, det.hpp:52-64, log_det.hpp:52-64; policies det.hpp:97-202.
     * n.op = policy (0 FullPivLU, 1 LDLT, 2 LLT), n.pw = 1 for log_det; n.extra = inverse */
    const double* det_feval(Node& n)
    {
        const Node& e = nodes[n.ch[0]];
        const double* A = feval(n.ch[0]);
        const size_t d = e.rows;
        double det = 0;
        if (n.op == 2) {
            double prodL = 0;
            n.pos_def = llt_decomp(d, A, prodL, n.extra);
            if (!n.pos_def) prodL = std::numeric_limits<double>::quiet_NaN();     /* Eigen leaves garbage; flagged invalid */
            n.val[0] = n.pw ? 2. * std::log(std::fabs(prodL)) : prodL * prodL;
        } else if (n.op == 1) {
            bool ok;
            ldlt_decomp(d, A, det, ok, n.extra);
            n.val[0] = n.pw ? std::log(std::fabs(det)) : det;
            n.pos_def = ok && (n.pw ? std::isfinite(n.val[0]) : det != 0);
        } else {
            bool inv_ok;
            fullpivlu_decomp(d, A, det, inv_ok, n.extra);
            n.val[0] = n.pw ? std::log(std::fabs(det)) : det;
            n.pos_def = inv_ok;
        }
        return n.val;
    }
    void det_beval(Node& n, double seed)
    {
        if (seed == 0 || !n.pos_def) return;
        const size_t d = nodes[n.ch[0]].rows;
        const double f = n.pw ? seed : seed * n.val[0];
        n.tmp0.resize(d * d);
        for (size_t j = 0; j < d; ++j)
            for (size_t i = 0; i < d; ++i) n.tmp0[i + j * d] = f * n.extra[j + i * d];     /* inverse().transpose() */
        beval(n.ch[0], Seed{n.tmp0.data(), 0});
    }

    /* WishartAdjLogPDFNode, wishart.hpp:140-214.  n.z_sq holds the (constant) degrees of freedom;
     * n.extra = [x_inv | v_inv | xv_inv], log_sigma = log det L_x, x_mean = log det L_v. */
    const double* wishart_feval(Node& n)
    {
        Node& xn = nodes[n.ch[0]];
        Node& vn = nodes[n.ch[1]];
        const double* X = feval(n.ch[0]);
        const double* V = feval(n.ch[1]);
        const size_t p = vn.rows, pp = p * p;
        n.extra.resize(3 * pp);
        std::vector<double> xi, vi;
        double pl = 0;
        bool v_ok = llt_decomp(p, V, pl, vi), x_ok = false;
        if (v_ok) {
            n.x_mean = std::log(pl);
            std::copy(vi.begin(), vi.end(), n.extra.begin() + pp);
            x_ok = llt_decomp(p, X, pl, xi);
            if (x_ok) {
                n.log_sigma = std::log(pl);
                std::copy(xi.begin(), xi.end(), n.extra.begin());
                for (size_t j = 0; j < p; ++j)
                    for (size_t i = 0; i < p; ++i) {
                        double acc = 0;
                        for (size_t k = 0; k < p; ++k) acc += X[i + k * p] * vi[k + j * p];
                        n.extra[2 * pp + i + j * p] = acc;
                    }
            }
        }
        (void)xn;
        const double df = n.z_sq;
        n.pos_def = v_ok && x_ok && (df + 1 > (double)p);
        if (!n.pos_def) { n.val[0] = -std::numeric_limits<double>::infinity(); return n.val; }
        double tr = 0;
        for (size_t i = 0; i < p; ++i) tr += n.extra[2 * pp + i + i * p];
        n.val[0] = (df - (double)p - 1.) * n.log_sigma - 0.5 * tr - df * n.x_mean;
        return n.val;
    }
    void wishart_beval(Node& n, double seed)
    {
        if (seed == 0 || !n.pos_def) return;
        const size_t p = nodes[n.ch[1]].rows, pp = p * p;
        const double df = n.z_sq, pd = (double)p;
        const double* xi = n.extra.data();
        const double* vi = xi + pp;
        const double* xv = vi + pp;
        n.tmp0.resize(pp);
        n.tmp1.resize(pp);
        for (size_t j = 0; j < p; ++j)
            for (size_t i = 0; i < p; ++i) {
                double acc = 0;
                for (size_t k = 0; k < p; ++k) acc += vi[i + k * p] * xv[k + j * p];
                n.tmp1[i + j * p] = (0.5 * seed) * (acc - df * vi[i + j * p]);
                n.tmp0[i + j * p] = (0.5 * seed) * ((df - pd - 1) * xi[i + j * p] - vi[i + j * p]);
            }
        beval(n.ch[1], Seed{n.tmp1.data(), 0});
        beval(n.ch[0], Seed{n.tmp0.data(), 0});
    }

    /* ---- dense covariance: vsm / vvm, normal.hpp:581-773.  Eigen::LLT<Lower> reads the lower
     * triangle only; log_det_ = log(det L); inv_ = llt.solve(I) (normal.hpp:657-665, 756-763). */
    static bool is_dense(const Node& sn, const Node& xn) { return sn.shape == ORC_MAT && sn.rows == sn.cols && sn.rows == xn.size() && sn.rows > 1; }
    void dense_update_cache(Node& n)
    {
        const Node& sn = nodes[n.ch[2]];
        const size_t d = sn.rows;
        std::vector<double> L(d * d, 0.);
        n.pos_def = true;
        for (size_t j = 0; j < d && n.pos_def; ++j) {                 /* column Cholesky */
            double s = sn.val[j + j * d];
            for (size_t k = 0; k < j; ++k) s -= L[j + k * d] * L[j + k * d];
            if (!(s > 0)) { n.pos_def = false; break; }
            const double ljj = std::sqrt(s);
            L[j + j * d] = ljj;
            for (size_t i = j + 1; i < d; ++i) {
                double t = sn.val[i + j * d];
                for (size_t k = 0; k < j; ++k) t -= L[i + k * d] * L[j + k * d];
                L[i + j * d] = t / ljj;
            }
        }
        if (!n.pos_def) return;
        double det = 1;
        for (size_t j = 0; j < d; ++j) det *= L[j + j * d];
        n.log_sigma = std::log(det);
        /* inv = (L L^T)^{-1}: solve L Y = I, then L^T X = Y, column by column */
        n.extra.assign(d * d + d, 0.);
        std::vector<double> y(d);
        for (size_t c = 0; c < d; ++c) {
            for (size_t i = 0; i < d; ++i) {
                double t = (i == c) ? 1. : 0.;
                for (size_t k = 0; k < i; ++k) t -= L[i + k * d] * y[k];
                y[i] = t / L[i + i * d];
            }
            for (size_t ii = d; ii-- > 0;) {
                double t = y[ii];
                for (size_t k = ii + 1; k < d; ++k) t -= L[k + ii * d] * n.extra[k + c * d];
                n.extra[ii + c * d] = t / L[ii + ii * d];
            }
        }
    }
    const double* dense_feval(Node& n)
    {
        Node& xn = nodes[n.ch[0]];
        Node& mn = nodes[n.ch[1]];
        Node& sn = nodes[n.ch[2]];
        const double* x = feval(n.ch[0]);
        const double* m = feval(n.ch[1]);
        feval(n.ch[2]);
        if (sn.kind != K_CONST || n.extra.empty()) dense_update_cache(n);
        const size_t d = xn.size();
        if (!n.pos_def) { n.val[0] = -std::numeric_limits<double>::infinity(); return n.val; }
        const bool mv = mn.size() > 1;
        double* z = n.extra.data() + d * d;
        double sq = 0;
        for (size_t i = 0; i < d; ++i) {
            double acc = 0;
            for (size_t k = 0; k < d; ++k) acc += n.extra[i + k * d] * (x[k] - m[mv ? k : 0]);
            z[i] = acc;
        }
        for (size_t i = 0; i < d; ++i) sq += (x[i] - m[mv ? i : 0]) * z[i];
        n.val[0] = -0.5 * sq - n.log_sigma;
        return n.val;
    }
    void dense_beval(Node& n, double seed)
    {
        if (seed == 0 || !n.pos_def) return;                          /* normal.hpp:644, 742 */
        Node& xn = nodes[n.ch[0]];
        Node& mn = nodes[n.ch[1]];
        Node& sn = nodes[n.ch[2]];
        const size_t d = xn.size();
        const double* z = n.extra.data() + d * d;
        if (sn.kind != K_CONST) {
            n.tmp1.resize(d * d);
            for (size_t j = 0; j < d; ++j)
                for (size_t i = 0; i < d; ++i) n.tmp1[i + j * d] = (-0.5 * seed) * (n.extra[i + j * d] - z[i] * z[j]);
            beval(n.ch[2], Seed{n.tmp1.data(), 0});
        }
        n.tmp0.resize(d);
        if (mn.size() > 1) { for (size_t i = 0; i < d; ++i) n.tmp0[i] = seed * z[i]; beval(n.ch[1], Seed{n.tmp0.data(), 0}); }
        else { double acc = 0; for (size_t i = 0; i < d; ++i) acc += z[i]; beval(n.ch[1], Seed{nullptr, seed * acc}); }
        for (size_t i = 0; i < d; ++i) n.tmp0[i] = (-seed) * z[i];
        beval(n.ch[0], Seed{n.tmp0.data(), 0});
    }

    const double* normal_feval(Node& n)
    {
        Node& xn = nodes[n.ch[0]];
        Node& mn = nodes[n.ch[1]];
        Node& sn = nodes[n.ch[2]];
        if (is_dense(sn, xn)) return dense_feval(n);
        const double* x = feval(n.ch[0]);
        const double* m = feval(n.ch[1]);
        const double* s = feval(n.ch[2]);
        const bool xc = xn.kind == K_CONST, sc = sn.kind == K_CONST;
        const size_t N = xn.size();
        const double ninf = -std::numeric_limits<double>::infinity();
        if (sn.shape == ORC_SCL) {
            if (s[0] <= 0) { n.val[0] = ninf; return n.val; }   /* normal.hpp:147, 231, 344 */
            if (!sc) normal_update_cache(n);
            if (xn.shape == ORC_SCL) {                          /* sss normal.hpp:141-156 */
                double z = (x[0] - m[0]) / s[0];
                n.val[0] = -0.5 * z * z - n.log_sigma;
            } else if (mn.shape == ORC_SCL && xc) {              /* vss const-x :237-242 */
                double centered = m[0] - n.x_mean;
                double inv_s_sq = 1. / (s[0] * s[0]);
                n.val[0] = -0.5 * inv_s_sq * (n.x_var + (double)N * centered * centered)
                           - (double)N * n.log_sigma;
            } else {                                            /* vss/vvs :243-247, 350-353 */
                double acc = 0;
                const bool mv = mn.shape != ORC_SCL;
                for (size_t i = 0; i < N; ++i) { double d = x[i] - m[mv ? i : 0]; acc += d * d; }
                n.z_sq = acc / (s[0] * s[0]);
                n.val[0] = -0.5 * n.z_sq - (double)N * n.log_sigma;
            }
        } else {
            if (!sc) normal_update_cache(n);                    /* normal.hpp:436-442, 538-544 */
            if (!n.pos_def) { n.val[0] = ninf; return n.val; }
            if (mn.shape == ORC_SCL && xc && sc) {               /* vsv const :444-448 */
                n.val[0] = -0.5 * (n.sq_term - 2 * m[0] * n.lin_term + m[0] * m[0] * n.const_term)
                           - n.log_sigma;
            } else {                                            /* vsv/vvv :449-452, 546-548 */
                double acc = 0;
                const bool mv = mn.shape != ORC_SCL;
                for (size_t i = 0; i < N; ++i) { double z = (x[i] - m[mv ? i : 0]) / s[i]; acc += z * z; }
                n.val[0] = -0.5 * acc - n.log_sigma;
            }
        }
        return n.val;
    }

    /* ---- backward: expr.beval(seed) ---- */
    void beval(int id, Seed seed)
    {
        Node& n = nodes[id];
        const size_t N = n.size();
        switch (n.kind) {
        case K_VAR:                                       /* var_view.hpp:91-94 */
            for (size_t i = 0; i < N; ++i) n.adj[i] += seed.at(i);
            return;
        case K_CONST: return;                             /* constant.hpp:55-56, 102-103 */
        case K_UNARY: {                                   /* unary.hpp:81-89 */
            if (seed.p != n.adj) for (size_t i = 0; i < N; ++i) n.adj[i] = seed.at(i);
            const double* x = nodes[n.ch[0]].val;
            n.tmp0.resize(N);
            for (size_t i = 0; i < N; ++i) n.tmp0[i] = unary_b(n.op, n.adj[i], x[i], n.val[i]);
            beval(n.ch[0], Seed{n.tmp0.data(), 0});
            return;
        }
        case K_BINARY: {                                  /* binary.hpp:120-136 */
            if (is_comparison(n.op)) return;              /* binary.hpp:124 */
            if (seed.p != n.adj) for (size_t i = 0; i < N; ++i) n.adj[i] = seed.at(i);
            Node& L = nodes[n.ch[0]];
            Node& R = nodes[n.ch[1]];
            const bool ls = L.size() == 1 && N > 1;
            const bool rs = R.size() == 1 && N > 1;
            n.tmp0.resize(N); n.tmp1.resize(N);
            for (size_t i = 0; i < N; ++i)
                n.tmp1[i] = binary_br(n.op, n.adj[i], L.val[ls ? 0 : i], R.val[rs ? 0 : i], n.val[i]);
            for (size_t i = 0; i < N; ++i)
                n.tmp0[i] = binary_bl(n.op, n.adj[i], L.val[ls ? 0 : i], R.val[rs ? 0 : i], n.val[i]);
            /* scalar side of a scalar-vs-array node receives .sum() (binary.hpp:266-338) */
            if (rs) { double a = 0; for (size_t i = 0; i < N; ++i) a += n.tmp1[i]; beval(n.ch[1], Seed{nullptr, a}); }
            else beval(n.ch[1], Seed{n.tmp1.data(), 0});
            if (ls) { double a = 0; for (size_t i = 0; i < N; ++i) a += n.tmp0[i]; beval(n.ch[0], Seed{nullptr, a}); }
            else beval(n.ch[0], Seed{n.tmp0.data(), 0});
            return;
        }
        case K_POW: {                                     /* pow.hpp:101-162 */
            if (n.pw == 0) { beval(n.ch[0], Seed{nullptr, 0.}); return; }
            if (n.pw == 1) { beval(n.ch[0], seed); return; }
            if (seed.p != n.adj) for (size_t i = 0; i < N; ++i) n.adj[i] = seed.at(i);
            const double* x = nodes[n.ch[0]].val;
            n.tmp0.resize(N);
            if (n.pw > 1) {
                for (size_t i = 0; i < N; ++i)
                    n.tmp0[i] = (double)n.pw * n.adj[i] * (x[i] == 0 ? 0 : n.val[i] / x[i]);
            } else {
                for (size_t i = 0; i < N; ++i)
                    n.tmp0[i] = x[i] == 0 ? -std::numeric_limits<double>::infinity()
                                          : (double)n.pw * n.adj[i] * n.val[i] / x[i];
            }
            beval(n.ch[0], Seed{n.tmp0.data(), 0});
            return;
        }
        case K_SUM:                                       /* sum.hpp:170-173 */
            beval(n.ch[0], Seed{nullptr, seed.at(0)});
            return;
        case K_SUM_ITER: {                                /* sum.hpp:75-85 */
            if (seed.p != n.adj) for (size_t i = 0; i < N; ++i) n.adj[i] = seed.at(i);
            for (size_t k = n.ch.size(); k-- > 0;) beval(n.ch[k], Seed{n.adj, 0});
            return;
        }
        case K_PROD: {                                    /* prod.hpp:188-212 */
            Node& e = nodes[n.ch[0]];
            const size_t M = e.size();
            const double sd = seed.at(0);
            for (size_t i = 0; i < M; ++i) {
                double a = sd;
                if (e.val[i] == 0) {
                    for (size_t j = 0; j < M; ++j) if (j != i) a *= e.val[j];
                } else {
                    a *= n.val[0] / e.val[i];
                }
                n.extra[i] = a;
            }
            beval(n.ch[0], Seed{n.extra.data(), 0});
            return;
        }
        case K_PROD_ITER: {                               /* prod.hpp:67-104 */
            if (n.ch.empty()) return;
            if (seed.p != n.adj) for (size_t i = 0; i < N; ++i) n.adj[i] = seed.at(i);
            n.tmp0.resize(N);
            for (size_t idx = n.ch.size(); idx-- > 0;) {
                Node& e = nodes[n.ch[idx]];
                bool any_zero = false;
                for (size_t i = 0; i < N; ++i) any_zero = any_zero || (e.val[i] == 0);
                if (any_zero) {
                    for (size_t i = 0; i < N; ++i) n.val[i] = 1;
                    for (size_t k = 0; k < n.ch.size(); ++k)
                        if (k != idx) for (size_t i = 0; i < N; ++i) n.val[i] *= nodes[n.ch[k]].val[i];
                    for (size_t i = 0; i < N; ++i) n.tmp0[i] = n.adj[i] * n.val[i];
                    beval(n.ch[idx], Seed{n.tmp0.data(), 0});
                    for (size_t i = 0; i < N; ++i) n.val[i] *= e.val[i];
                } else {
                    for (size_t i = 0; i < N; ++i) n.tmp0[i] = n.adj[i] * n.val[i] / e.val[i];
                    beval(n.ch[idx], Seed{n.tmp0.data(), 0});
                }
            }
            return;
        }
        case K_DOT: {                                     /* dot.hpp:96-104 */
            if (seed.p != n.adj) for (size_t i = 0; i < N; ++i) n.adj[i] = seed.at(i);
            Node& L = nodes[n.ch[0]];
            Node& R = nodes[n.ch[1]];
            const size_t m = n.rows, p = n.cols, k = L.cols;
            n.tmp0.assign(m * k, 0.);   /* adj * R^T  (m x k) */
            n.tmp1.assign(k * p, 0.);   /* L^T * adj  (k x p) */
            for (size_t j = 0; j < p; ++j)
                for (size_t q = 0; q < k; ++q) {
                    const double rq = R.val[q + j * k];
                    for (size_t i = 0; i < m; ++i) n.tmp0[i + q * m] += n.adj[i + j * m] * rq;
                }
            for (size_t j = 0; j < p; ++j)
                for (size_t q = 0; q < k; ++q) {
                    double acc = 0;
                    for (size_t i = 0; i < m; ++i) acc += L.val[i + q * m] * n.adj[i + j * m];
                    n.tmp1[q + j * k] = acc;
                }
            beval(n.ch[1], Seed{n.tmp1.data(), 0});
            beval(n.ch[0], Seed{n.tmp0.data(), 0});
            return;
        }
        case K_EQ: {                                      /* eq.hpp:101-107 */
            beval(n.ch[0], seed);
            beval(n.ch[1], Seed{nodes[n.ch[0]].adj, 0});
            return;
        }
        case K_GLUE:                                      /* glue.hpp:81-86 */
            beval(n.ch[1], seed);
            beval(n.ch[0], Seed{nullptr, 0.});
            return;
        case K_NORMAL: normal_beval(n, seed.at(0)); return;
        case K_LOGPDF: logpdf_beval(n, seed.at(0)); return;
        case K_DET: det_beval(n, seed.at(0)); return;
        case K_WISHART: wishart_beval(n, seed.at(0)); return;
        case K_NORM: {                                    /* norm.hpp:53-57 */
            Node& e = nodes[n.ch[0]];
            n.tmp0.resize(e.size());
            for (size_t i = 0; i < e.size(); ++i) n.tmp0[i] = seed.at(0) * 2. * e.val[i];
            beval(n.ch[0], Seed{n.tmp0.data(), 0});
            return;
        }
        case K_TRANSPOSE: {                               /* transpose.hpp:37-41 */
            if (seed.p != n.adj) for (size_t i = 0; i < N; ++i) n.adj[i] = seed.at(i);
            const size_t m = nodes[n.ch[0]].rows, k = nodes[n.ch[0]].cols;
            n.tmp0.resize(N);
            for (size_t j = 0; j < k; ++j)
                for (size_t i = 0; i < m; ++i) n.tmp0[i + j * m] = n.adj[j + i * k];
            beval(n.ch[0], Seed{n.tmp0.data(), 0});
            return;
        }
        case K_IF_ELSE:                                   /* if_else.hpp:70-78 */
            beval(n.ch[n.pw == 1 ? 1 : 2], seed);
            return;
        case K_OPEQ: {                                    /* eq.hpp:249-271 */
            Node& w = nodes[n.ch[0]];
            Node& e = nodes[n.ch[1]];
            const bool es = e.size() == 1 && N > 1;
            for (size_t i = 0; i < N; ++i) w.adj[i] += seed.at(i);          /* var_view_.beval(seed) */
            for (size_t i = 0; i < N; ++i) w.val[i] = n.extra[i];           /* restore previous lhs */
            double* cadj = n.extra.data() + N;
            for (size_t i = 0; i < N; ++i) { cadj[i] = w.adj[i]; w.adj[i] = 0; }
            n.tmp0.resize(N); n.tmp1.resize(N);
            for (size_t i = 0; i < N; ++i) {
                const double s = cadj[i], x = w.val[i], y = e.val[es ? 0 : i];
                double ls, rs;
                switch (n.op) {                                             /* eq.hpp:305-413 */
                case ORC_ADD: ls = s; rs = s; break;
                case ORC_SUB: ls = s; rs = -s; break;
                case ORC_MUL: ls = s * y; rs = s * x; break;
                default: ls = s / y; rs = -s * x / (y * y); break;
                }
                n.tmp0[i] = ls; n.tmp1[i] = rs;
            }
            if (es) { double acc = 0; for (size_t i = 0; i < N; ++i) acc += n.tmp1[i]; beval(n.ch[1], Seed{nullptr, acc}); }
            else beval(n.ch[1], Seed{n.tmp1.data(), 0});
            beval(n.ch[0], Seed{n.tmp0.data(), 0});
            return;
        }
        }
    }

    /* ---- cauchy / uniform / bernoulli adjusted log-pdfs (reverse/stat/{cauchy,uniform,bernoulli}.hpp).
     * Operands are scalars or vectors of one length; `at(p, k, i)` reads operand k at element i. */
    static double at(const Node& o, size_t i) { return o.val[o.size() == 1 ? 0 : i]; }
    bool logpdf_valid(const Node& n) const
    {
        const Node& A = nodes[n.ch[0]];
        const Node& B = nodes[n.ch[1]];
        size_t N = std::max(A.size(), B.size());
        if (n.op != ORC_BERNOULLI) N = std::max(N, nodes[n.ch[2]].size());
        for (size_t i = 0; i < N; ++i) {
            if (n.op == ORC_CAUCHY) { if (!(at(nodes[n.ch[2]], i) > 0)) return false; }          /* cauchy.hpp:168-170, 309-311 */
            else if (n.op == ORC_UNIFORM) {                                                       /* uniform.hpp:170-173 */
                if (!(at(B, i) < at(A, i) && at(A, i) < at(nodes[n.ch[2]], i))) return false;
            } else { if (!(0 < at(B, i) && at(B, i) < 1)) return false; }                         /* bernoulli.hpp:165-168 */
        }
        return true;
    }
    const double* logpdf_feval(Node& n)
    {
        feval(n.ch[0]); feval(n.ch[1]);
        if (n.op != ORC_BERNOULLI) feval(n.ch[2]);
        const Node& A = nodes[n.ch[0]];
        const Node& B = nodes[n.ch[1]];
        size_t N = std::max(A.size(), B.size());
        if (n.op != ORC_BERNOULLI) N = std::max(N, nodes[n.ch[2]].size());
        const double ninf = -std::numeric_limits<double>::infinity();
        n.pos_def = logpdf_valid(n);
        if (n.op == ORC_CAUCHY) {
            const Node& G = nodes[n.ch[2]];
            if (!n.pos_def) { n.val[0] = ninf; return n.val; }
            if (N == 1) {                                                                         /* sss, cauchy.hpp:134-147 */
                const double d = A.val[0] - B.val[0], g = G.val[0];
                n.z_sq = g + (d * d) / g;
                n.val[0] = -std::log(n.z_sq);
            } else {                                                                              /* cauchy.hpp:216-217 etc. */
                double acc = 0;
                for (size_t i = 0; i < N; ++i) {
                    const double d = at(A, i) - at(B, i), g = at(G, i);
                    acc += std::log(g + (1. / g) * (d * d));
                }
                n.val[0] = -acc;
            }
        } else if (n.op == ORC_UNIFORM) {
            const Node& MX = nodes[n.ch[2]];
            if (!n.pos_def) { n.val[0] = ninf; return n.val; }
            if (B.size() == 1 && MX.size() == 1) {
                n.val[0] = -(double)N * std::log(MX.val[0] - B.val[0]);                          /* uniform.hpp:154, 240 */
                if (N == 1) n.val[0] = -std::log(MX.val[0] - B.val[0]);
            } else {
                double acc = 0;
                for (size_t i = 0; i < N; ++i) acc += std::log(at(MX, i) - at(B, i));             /* uniform.hpp:348, 441, 529 */
                n.val[0] = -acc;
            }
        } else {                                                                                  /* bernoulli.hpp:119-141, 215-241 */
            bool zero_one = true;
            double x_sum = 0;
            for (size_t i = 0; i < A.size(); ++i) { zero_one = zero_one && (A.val[i] == 0 || A.val[i] == 1); x_sum += A.val[i]; }
            n.x_var = zero_one ? 1 : 0;
            if (B.size() == 1) {
                const double p = B.val[0];
                if (!n.pos_def) {
                    if (p <= 0) n.val[0] = (x_sum == 0 && zero_one) ? 0 : ninf;
                    else n.val[0] = (x_sum == (double)A.size() && zero_one) ? 0 : ninf;
                } else if (!zero_one) n.val[0] = ninf;
                else n.val[0] = x_sum * std::log(p) + ((double)A.size() - x_sum) * std::log(1 - p);
            } else {
                double acc = 0;
                for (size_t i = 0; i < N; ++i) {
                    const double x = at(A, i), p = at(B, i);
                    if (!(0 < p && p < 1)) acc += (p <= 0) ? (x == 0 ? 0 : ninf) : (x == 1 ? 0 : ninf);
                    else acc += x == 0 ? std::log(1 - p) : x == 1 ? std::log(p) : ninf;
                }
                n.val[0] = acc;
            }
        }
        return n.val;
    }
    void logpdf_beval(Node& n, double seed)
    {
        if (seed == 0 || !n.pos_def) return;
        Node& A = nodes[n.ch[0]];
        Node& B = nodes[n.ch[1]];
        size_t N = std::max(A.size(), B.size());
        if (n.op != ORC_BERNOULLI) N = std::max(N, nodes[n.ch[2]].size());
        auto send = [&](int child, std::vector<double>& per_elem) {     /* scalar operands receive the sum */
            Node& c = nodes[child];
            if (c.size() == 1 && N > 1) { double a = 0; for (double v : per_elem) a += v; beval(child, Seed{nullptr, a}); }
            else beval(child, Seed{per_elem.data(), 0});
        };
        std::vector<double> ga(N), gb(N), gc(N);
        if (n.op == ORC_CAUCHY) {
            Node& G = nodes[n.ch[2]];
            if (N == 1) {                                                                         /* cauchy.hpp:149-165 */
                const double d = A.val[0] - B.val[0], g = G.val[0];
                const double x0_adj = 2. * d / (g * n.z_sq);
                beval(n.ch[2], Seed{nullptr, seed * (1. / g * (x0_adj * d - 1))});
                beval(n.ch[1], Seed{nullptr, seed * x0_adj});
                beval(n.ch[0], Seed{nullptr, seed * -x0_adj});
                return;
            }
            /* vector cases, cauchy.hpp:220-237, 290-306, 359-376, 429-445 (note: the reference
             * multiplies by `seed` twice in dx0 and dgamma; restated as written) */
            for (size_t i = 0; i < N; ++i) {
                const double d = at(A, i) - at(B, i), g = at(G, i);
                const double dx = (-2. * seed) * d / (g * g + d * d);
                ga[i] = dx;
                gb[i] = (-seed) * dx;
                gc[i] = (-seed / g) * (dx * d + 1);
            }
            send(n.ch[2], gc); send(n.ch[1], gb); send(n.ch[0], ga);
        } else if (n.op == ORC_UNIFORM) {                                                         /* uniform.hpp:157-163, 243-250, 336-344 */
            Node& MX = nodes[n.ch[2]];
            for (size_t i = 0; i < N; ++i) { const double a = seed / (at(MX, i) - at(B, i)); gc[i] = -a; gb[i] = a; }
            send(n.ch[2], gc); send(n.ch[1], gb);
        } else {                                                                                  /* bernoulli.hpp:143-149, 243-250 */
            if (n.x_var == 0) return;
            for (size_t i = 0; i < N; ++i) { const double x = at(A, i), p = at(B, i); gb[i] = x == 0 ? -seed / (1 - p) : seed / p; }
            send(n.ch[1], gb);
        }
    }

    void normal_beval(Node& n, double seed)
    {
        Node& xn = nodes[n.ch[0]];
        Node& mn = nodes[n.ch[1]];
        Node& sn = nodes[n.ch[2]];
        if (is_dense(sn, xn)) { dense_beval(n, seed); return; }
        const bool xc = xn.kind == K_CONST, sc = sn.kind == K_CONST;
        const size_t N = xn.size();
        const double* x = xn.val;
        const double* m = mn.val;
        const double* s = sn.val;
        if (sn.shape == ORC_SCL) {
            if (seed == 0 || s[0] <= 0) return;               /* normal.hpp:160, 252, 358 */
            const double inv_s = 1. / s[0];
            if (xn.shape == ORC_SCL) {                        /* sss :158-171 */
                double z = (x[0] - m[0]) * inv_s;
                if (!sc) beval(n.ch[2], Seed{nullptr, seed * (z * z - 1) * inv_s});
                double adj = seed * z * inv_s;
                beval(n.ch[1], Seed{nullptr, adj});
                beval(n.ch[0], Seed{nullptr, -adj});
                return;
            }
            const double inv_s_sq = inv_s * inv_s;
            if (mn.shape == ORC_SCL) {                        /* vss :250-286 */
                if (xc) {
                    if (!sc) {
                        double c = m[0] - n.x_mean;
                        double sigma_adj = ((n.x_var + (double)N * c * c) * inv_s_sq - (double)N) * inv_s;
                        beval(n.ch[2], Seed{nullptr, seed * sigma_adj});
                    }
                    double mean_adj = (double)N * (n.x_mean - m[0]) * inv_s_sq;
                    beval(n.ch[1], Seed{nullptr, seed * mean_adj});
                } else {
                    if (!sc) beval(n.ch[2], Seed{nullptr, seed * (n.z_sq - (double)N) * inv_s});
                    double acc = 0;
                    for (size_t i = 0; i < N; ++i) acc += x[i] - m[0];
                    beval(n.ch[1], Seed{nullptr, seed * (acc * inv_s_sq)});
                    n.tmp0.resize(N);
                    const double c = seed * inv_s_sq;
                    for (size_t i = 0; i < N; ++i) n.tmp0[i] = c * (m[0] - x[i]);
                    beval(n.ch[0], Seed{n.tmp0.data(), 0});
                }
            } else {                                          /* vvs :356-372 */
                if (!sc) beval(n.ch[2], Seed{nullptr, seed * (n.z_sq - (double)N) * inv_s});
                n.tmp0.resize(N);
                const double c = seed * inv_s_sq;
                for (size_t i = 0; i < N; ++i) n.tmp0[i] = c * (x[i] - m[i]);
                beval(n.ch[1], Seed{n.tmp0.data(), 0});
                for (size_t i = 0; i < N; ++i) n.tmp0[i] = c * (m[i] - x[i]);
                beval(n.ch[0], Seed{n.tmp0.data(), 0});
            }
            return;
        }
        if (seed == 0 || !n.pos_def) return;                  /* normal.hpp:457, 553 */
        n.tmp0.resize(N);
        if (mn.shape == ORC_SCL) {                            /* vsv :455-478 */
            if (xc && sc) {
                double mean_adj = n.lin_term - m[0] * n.const_term;
                beval(n.ch[1], Seed{nullptr, seed * mean_adj});
                return;
            }
            if (!sc) {
                for (size_t i = 0; i < N; ++i) {
                    double z = (x[i] - m[0]) / s[i];
                    n.tmp0[i] = (seed / s[i]) * (z * z - 1.);
                }
                beval(n.ch[2], Seed{n.tmp0.data(), 0});
            }
            double acc = 0;
            for (size_t i = 0; i < N; ++i) acc += (x[i] - m[0]) / (s[i] * s[i]);
            beval(n.ch[1], Seed{nullptr, seed * acc});
            for (size_t i = 0; i < N; ++i) n.tmp0[i] = (seed / (s[i] * s[i])) * (m[0] - x[i]);
            beval(n.ch[0], Seed{n.tmp0.data(), 0});
        } else {                                              /* vvv :551-565 */
            if (!sc) {
                for (size_t i = 0; i < N; ++i) {
                    double z = (x[i] - m[i]) / s[i];
                    n.tmp0[i] = (seed / s[i]) * (z * z - 1.);
                }
                beval(n.ch[2], Seed{n.tmp0.data(), 0});
            }
            for (size_t i = 0; i < N; ++i) n.tmp0[i] = seed * (x[i] - m[i]) / (s[i] * s[i]);
            beval(n.ch[1], Seed{n.tmp0.data(), 0});
            for (size_t i = 0; i < N; ++i) n.tmp0[i] = seed * (m[i] - x[i]) / (s[i] * s[i]);
            beval(n.ch[0], Seed{n.tmp0.data(), 0});
        }
    }
};

/* ===================== C API: graph construction ===================== */

extern "C" {

orc_graph* orc_new(void) { return new orc_graph(); }
void orc_free(orc_graph* g) { delete g; }

int orc_var(orc_graph* g, int shape, size_t rows, size_t cols, double* val, double* adj)
{
    Node n; n.kind = K_VAR; n.shape = shape; n.rows = rows; n.cols = cols; n.val = val; n.adj = adj;
    return g->add(std::move(n));
}

static int make_const(orc_graph* g, int shape, size_t rows, size_t cols, std::vector<double>&& v)
{
    Node n; n.kind = K_CONST; n.shape = shape; n.rows = rows; n.cols = cols;
    n.cval = std::move(v);
    int id = g->add(std::move(n));
    g->nodes[id].val = g->nodes[id].cval.data();
    return id;
}

int orc_const_scalar(orc_graph* g, double c) { return make_const(g, ORC_SCL, 1, 1, {c}); }

int orc_const_view(orc_graph* g, int shape, size_t rows, size_t cols, const double* val)
{   /* ConstantView, constant.hpp:82-151: views caller memory */
    Node n; n.kind = K_CONST; n.shape = shape; n.rows = rows; n.cols = cols;
    n.val = const_cast<double*>(val);
    return g->add(std::move(n));
}

void orc_rebind(orc_graph* g, int leaf, double* val, double* adj, size_t rows, size_t cols)
{
    Node& n = g->nodes[leaf];
    n.val = val; n.adj = adj; n.rows = rows; n.cols = cols;
}

static void fix_const_ptrs(orc_graph* g)
{   /* std::vector<Node> may reallocate; owned constants are re-pointed */
    for (auto& n : g->nodes) if (n.kind == K_CONST && !n.cval.empty()) n.val = n.cval.data();
}

int orc_unary(orc_graph* g, int op, int child)
{
    const Node& c = g->nodes[child];
    if (c.kind == K_CONST) {                        /* eager folding, unary.hpp:182-184 */
        std::vector<double> v(c.size());
        for (size_t i = 0; i < v.size(); ++i) v[i] = unary_f(op, c.val[i]);
        int id = make_const(g, c.shape, c.rows, c.cols, std::move(v));
        fix_const_ptrs(g);
        return id;
    }
    Node n; n.kind = K_UNARY; n.op = op; n.shape = c.shape; n.rows = c.rows; n.cols = c.cols;
    n.ch = {child};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

int orc_binary(orc_graph* g, int op, int lhs, int rhs)
{
    const Node& l = g->nodes[lhs];
    const Node& r = g->nodes[rhs];
    const Node& big = (l.shape >= r.shape) ? l : r;  /* max_shape_t, shape_traits.hpp:102-110 */
    if (l.kind == K_CONST && r.kind == K_CONST) {   /* binary.hpp:251-256 */
        std::vector<double> v(big.size());
        const bool ls = l.size() == 1 && v.size() > 1, rs = r.size() == 1 && v.size() > 1;
        for (size_t i = 0; i < v.size(); ++i) v[i] = binary_f(op, l.val[ls ? 0 : i], r.val[rs ? 0 : i]);
        int id = make_const(g, big.shape, big.rows, big.cols, std::move(v));
        fix_const_ptrs(g);
        return id;
    }
    Node n; n.kind = K_BINARY; n.op = op; n.shape = big.shape;
    n.rows = std::max(l.rows, r.rows); n.cols = std::max(l.cols, r.cols);
    n.ch = {lhs, rhs};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

int orc_pow(orc_graph* g, long pw, int child)
{
    const Node& c = g->nodes[child];
    if (c.kind == K_CONST) {                        /* pow.hpp:219-227 */
        std::vector<double> v(c.size());
        for (size_t i = 0; i < v.size(); ++i)
            v[i] = c.shape == ORC_SCL ? pow_func(pw, c.val[i]) : std::pow(c.val[i], (double)pw);
        int id = make_const(g, c.shape, c.rows, c.cols, std::move(v));
        fix_const_ptrs(g);
        return id;
    }
    Node n; n.kind = K_POW; n.pw = pw; n.shape = c.shape; n.rows = c.rows; n.cols = c.cols;
    n.ch = {child};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

int orc_sum(orc_graph* g, int child)
{
    const Node& c = g->nodes[child];
    if (c.kind == K_CONST) {                        /* sum.hpp:250-256 */
        if (c.shape == ORC_SCL) return child;
        double acc = 0;
        for (size_t i = 0; i < c.size(); ++i) acc += c.val[i];
        int id = make_const(g, ORC_SCL, 1, 1, {acc}); fix_const_ptrs(g); return id;
    }
    Node n; n.kind = K_SUM; n.ch = {child};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

static int make_iter(orc_graph* g, Kind kind, int k, const int* children)
{
    Node n; n.kind = kind;
    if (k > 0) {
        const Node& c = g->nodes[children[0]];
        n.shape = c.shape; n.rows = c.rows; n.cols = c.cols;
    }
    n.ch.assign(children, children + k);
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}
int orc_sum_iter(orc_graph* g, int k, const int* ch) { return make_iter(g, K_SUM_ITER, k, ch); }
int orc_prod_iter(orc_graph* g, int k, const int* ch) { return make_iter(g, K_PROD_ITER, k, ch); }

int orc_prod(orc_graph* g, int child)
{
    const Node& c = g->nodes[child];
    if (c.kind == K_CONST) {                        /* prod.hpp:292-300 */
        if (c.shape == ORC_SCL) return child;
        double acc = 1;
        for (size_t i = 0; i < c.size(); ++i) acc *= c.val[i];
        int id = make_const(g, ORC_SCL, 1, 1, {acc}); fix_const_ptrs(g); return id;
    }
    Node n; n.kind = K_PROD; n.ch = {child};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

int orc_dot(orc_graph* g, int lhs, int rhs)
{
    const Node& l = g->nodes[lhs];
    const Node& r = g->nodes[rhs];
    assert(l.cols == r.rows);
    Node n; n.kind = K_DOT; n.shape = r.shape == ORC_VEC ? ORC_VEC : ORC_MAT;  /* dot.hpp:19-33 */
    n.rows = l.rows; n.cols = r.cols; n.ch = {lhs, rhs};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

int orc_normal(orc_graph* g, int x, int mu, int sigma)
{
    Node n; n.kind = K_NORMAL; n.ch = {x, mu, sigma};
    int id = g->add(std::move(n)); fix_const_ptrs(g);
    g->normal_ctor(g->nodes[id]);
    return id;
}

int orc_eq(orc_graph* g, int var, int child)
{
    const Node& w = g->nodes[var];
    Node n; n.kind = K_EQ; n.shape = w.shape; n.rows = w.rows; n.cols = w.cols; n.ch = {var, child};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

int orc_glue(orc_graph* g, int lhs, int rhs)
{
    const Node& r = g->nodes[rhs];
    Node n; n.kind = K_GLUE; n.shape = r.shape; n.rows = r.rows; n.cols = r.cols; n.ch = {lhs, rhs};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

int orc_logpdf(orc_graph* g, int dist, int a, int b, int c)
{
    Node n; n.kind = K_LOGPDF; n.op = dist;
    n.ch = {a, b};
    if (dist != ORC_BERNOULLI) n.ch.push_back(c);
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}


int orc_det(orc_graph* g, int policy, int take_log, int child)
{
    const Node& c = g->nodes[child];
    if (c.kind == K_CONST) {     /* det.hpp:222-227: constants fold through FullPivLU's determinant */
        double det; bool ok; std::vector<double> inv;
        orc_graph::fullpivlu_decomp(c.rows, c.val, det, ok, inv);
        int id = make_const(g, ORC_SCL, 1, 1, {take_log ? std::log(std::fabs(det)) : det}); fix_const_ptrs(g); return id;
    }
    Node n; n.kind = K_DET; n.op = policy; n.pw = take_log; n.ch = {child};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

int orc_wishart(orc_graph* g, int x, int v, double df)
{
    Node n; n.kind = K_WISHART; n.z_sq = df; n.ch = {x, v};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

int orc_norm(orc_graph* g, int child)
{
    const Node& c = g->nodes[child];
    if (c.kind == K_CONST) {
        double acc = 0;
        for (size_t i = 0; i < c.size(); ++i) acc += c.val[i] * c.val[i];
        int id = make_const(g, ORC_SCL, 1, 1, {acc}); fix_const_ptrs(g); return id;
    }
    Node n; n.kind = K_NORM; n.ch = {child};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

int orc_transpose(orc_graph* g, int child)
{
    const Node& c = g->nodes[child];
    Node n; n.kind = K_TRANSPOSE; n.shape = ORC_MAT; n.rows = c.cols; n.cols = c.rows; n.ch = {child};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

int orc_if_else(orc_graph* g, int cond, int a, int b)
{
    const Node& c = g->nodes[cond];
    if (c.kind == K_CONST && c.size() == 1) return c.val[0] != 0 ? a : b;     /* if_else.hpp:141-155 */
    const Node& x = g->nodes[a];
    Node n; n.kind = K_IF_ELSE; n.shape = x.shape; n.rows = x.rows; n.cols = x.cols; n.ch = {cond, a, b};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

int orc_opeq(orc_graph* g, int op, int var, int child)
{
    const Node& w = g->nodes[var];
    Node n; n.kind = K_OPEQ; n.op = op; n.shape = w.shape; n.rows = w.rows; n.cols = w.cols; n.ch = {var, child};
    int id = g->add(std::move(n)); fix_const_ptrs(g); return id;
}

void orc_bind(orc_graph* g, int root, size_t out2[2])
{
    size_t o[2];
    g->bind(root, o);
    if (out2) { out2[0] = o[0]; out2[1] = o[1]; }
}

const double* orc_feval(orc_graph* g, int root) { return g->feval(root); }

void orc_beval(orc_graph* g, int root, double seed_scalar, const double* seed_array)
{
    g->beval(root, Seed{seed_array, seed_scalar});
}

size_t orc_size(orc_graph* g, int node) { return g->nodes[node].size(); }
int orc_is_const(orc_graph* g, int node) { return g->is_const(node) ? 1 : 0; }

} // extern "C"

/* ===================== whole workloads ===================== */

namespace {

template <class F>
void parallel_ranges(size_t n, int nthreads, F&& f)
{
    if (nthreads <= 1 || n < (size_t)nthreads) { f(0, (size_t)0, n); return; }
    std::vector<std::thread> th;
    const size_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        size_t b = std::min(n, per * t), e = std::min(n, per * (t + 1));
        th.emplace_back([&f, t, b, e] { if (e > b) f(t, b, e); });
    }
    for (auto& x : th) x.join();
}

/* The Black-Scholes expression of example/black_scholes.cpp:16-39 over vec leaves.
 * S, sigma, r are variables (so delta/vega/rho exist); K, tau and sqrt(tau) are constant
 * views (the example computes std::sqrt(tau) on the host as a plain double, :28-30). */
struct BSGraph {
    orc_graph* g;
    int S, K, sigma, tau, sqrt_tau, r, c0, c1, c2, root;
    std::vector<double> c0v, c0a, c1v, c1a, c2v, c2a, sq;
    size_t cap;

    int phi(int x)
    {   /* Phi(x) = 0.5 * (erf(x / sqrt(2)) + 1), black_scholes.cpp:9-13 */
        int e = orc_unary(g, ORC_ERF, orc_binary(g, ORC_DIV, x, orc_const_scalar(g, std::sqrt(2.))));
        return orc_binary(g, ORC_MUL, orc_const_scalar(g, 0.5),
                          orc_binary(g, ORC_ADD, e, orc_const_scalar(g, 1.)));
    }
    BSGraph(size_t chunk, bool put) : cap(chunk)
    {
        g = orc_new();
        c0v.resize(cap); c0a.resize(cap); c1v.resize(cap); c1a.resize(cap); c2v.resize(cap); c2a.resize(cap);
        sq.resize(cap);
        S = orc_var(g, ORC_VEC, cap, 1, nullptr, nullptr);
        sigma = orc_var(g, ORC_VEC, cap, 1, nullptr, nullptr);
        r = orc_var(g, ORC_VEC, cap, 1, nullptr, nullptr);
        K = orc_const_view(g, ORC_VEC, cap, 1, nullptr);
        tau = orc_const_view(g, ORC_VEC, cap, 1, nullptr);
        sqrt_tau = orc_const_view(g, ORC_VEC, cap, 1, sq.data());
        c0 = orc_var(g, ORC_VEC, cap, 1, c0v.data(), c0a.data());
        c1 = orc_var(g, ORC_VEC, cap, 1, c1v.data(), c1a.data());
        c2 = orc_var(g, ORC_VEC, cap, 1, c2v.data(), c2a.data());
        /* node construction never dereferences a non-constant leaf, and no constant-constant
         * pair is combined below, so the views can be (re)bound afterwards */
        int two = orc_const_scalar(g, 2.);
        int e0 = orc_eq(g, c0, orc_unary(g, ORC_LOG, orc_binary(g, ORC_DIV, S, K)));
        int drift = orc_binary(g, ORC_MUL,
                        orc_binary(g, ORC_ADD, r,
                            orc_binary(g, ORC_DIV, orc_binary(g, ORC_MUL, sigma, sigma), two)),
                        tau);
        int e1 = orc_eq(g, c1, orc_binary(g, ORC_DIV, orc_binary(g, ORC_ADD, c0, drift),
                                          orc_binary(g, ORC_MUL, sigma, sqrt_tau)));
        int e2 = orc_eq(g, c2, orc_binary(g, ORC_SUB, c1, orc_binary(g, ORC_MUL, sigma, sqrt_tau)));
        int common = orc_glue(g, orc_glue(g, e0, e1), e2);
        /* PV = K * exp(-r * tau) with r a variable */
        int pv = orc_binary(g, ORC_MUL, K,
                     orc_unary(g, ORC_EXP, orc_binary(g, ORC_MUL, orc_unary(g, ORC_NEG, r), tau)));
        if (!put) {         /* Phi(c1)*S - Phi(c2)*PV */
            root = orc_glue(g, common,
                orc_binary(g, ORC_SUB, orc_binary(g, ORC_MUL, phi(c1), S),
                                       orc_binary(g, ORC_MUL, phi(c2), pv)));
        } else {            /* Phi(-c2)*PV - Phi(-c1)*S */
            root = orc_glue(g, common,
                orc_binary(g, ORC_SUB,
                    orc_binary(g, ORC_MUL, phi(orc_unary(g, ORC_NEG, c2)), pv),
                    orc_binary(g, ORC_MUL, phi(orc_unary(g, ORC_NEG, c1)), S)));
        }
        orc_bind(g, root, nullptr);
    }
    ~BSGraph() { orc_free(g); }
};

} // namespace

extern "C" {

void orc_black_scholes(size_t n, const double* S, const double* K, const double* sigma,
                       const double* tau, const double* r, double* const out[8], int nthreads)
{
    /* out: call, put, dcall/dS, dcall/dsigma, dcall/dr, dput/dS, dput/dsigma, dput/dr */
    const size_t CH = 1024;  /* chunk stays cache resident: the best way to batch the reference */
    parallel_ranges(n, nthreads, [&](int, size_t b, size_t e) {
        std::vector<double> ones(CH, 1.0);
        size_t off = b;
        while (off < e) {
            const size_t m = std::min(CH, e - off);
            /* one bound expression per option type, reused for every full chunk
             * (ad::bind once, autodiff many times) */
            BSGraph call(m, false), put(m, true);
            for (; off + m <= e; off += m) {
                for (int which = 0; which < 2; ++which) {
                    BSGraph& bs = which ? put : call;
                    orc_graph* g = bs.g;
                    double* dS = out[which ? 5 : 2] + off;
                    double* dsg = out[which ? 6 : 3] + off;
                    double* dr = out[which ? 7 : 4] + off;
                    std::memset(dS, 0, m * sizeof(double));
                    std::memset(dsg, 0, m * sizeof(double));
                    std::memset(dr, 0, m * sizeof(double));
                    std::memset(bs.c0a.data(), 0, m * sizeof(double));   /* reset_adj() */
                    std::memset(bs.c1a.data(), 0, m * sizeof(double));
                    std::memset(bs.c2a.data(), 0, m * sizeof(double));
                    for (size_t i = 0; i < m; ++i) bs.sq[i] = std::sqrt(tau[off + i]);
                    orc_rebind(g, bs.S, const_cast<double*>(S + off), dS, m, 1);
                    orc_rebind(g, bs.sigma, const_cast<double*>(sigma + off), dsg, m, 1);
                    orc_rebind(g, bs.r, const_cast<double*>(r + off), dr, m, 1);
                    orc_rebind(g, bs.K, const_cast<double*>(K + off), nullptr, m, 1);
                    orc_rebind(g, bs.tau, const_cast<double*>(tau + off), nullptr, m, 1);
                    const double* v = orc_feval(g, bs.root);
                    std::memcpy(out[which] + off, v, m * sizeof(double));
                    orc_beval(g, bs.root, 0., ones.data());
                }
                if (m < CH) break;
            }
            if (m < CH) { off = e; }
        }
    });
}

double orc_sum_unary(int op, size_t n, const double* x, double* x_adj, double seed, int nthreads)
{
    std::vector<double> part(std::max(nthreads, 1), 0.);
    parallel_ranges(n, nthreads, [&](int t, size_t b, size_t e) {
        orc_graph* g = orc_new();
        int xv = orc_var(g, ORC_VEC, e - b, 1, const_cast<double*>(x + b), x_adj + b);
        int root = orc_sum(g, op >= 900 ? orc_pow(g, op - 1000, xv) : orc_unary(g, op, xv));  /* 1000+n = pow<n> */
        orc_bind(g, root, nullptr);
        part[t] = orc_feval(g, root)[0];
        orc_beval(g, root, seed, nullptr);
        orc_free(g);
    });
    double acc = 0;
    for (double p : part) acc += p;
    return acc;
}

double orc_normal_lpdf(int shape_code, int const_mask, size_t n, const double* x,
                       const double* mu, const double* sigma, double* x_adj, double* mu_adj,
                       double* sigma_adj, double seed, int nthreads)
{
    const bool mv = shape_code & 1, sv = shape_code & 2;
    const int T = std::max(nthreads, 1);
    std::vector<double> part(T, 0.), pm(T, 0.), ps(T, 0.);
    parallel_ranges(n, nthreads, [&](int t, size_t b, size_t e) {
        orc_graph* g = orc_new();
        const size_t m = e - b;
        int xn = (const_mask & 1) ? orc_const_view(g, ORC_VEC, m, 1, x + b)
                                  : orc_var(g, ORC_VEC, m, 1, const_cast<double*>(x + b), x_adj + b);
        int mn, sn;
        if (mv) mn = (const_mask & 2) ? orc_const_view(g, ORC_VEC, m, 1, mu + b)
                                      : orc_var(g, ORC_VEC, m, 1, const_cast<double*>(mu + b), mu_adj + b);
        else mn = (const_mask & 2) ? orc_const_scalar(g, mu[0])
                                   : orc_var(g, ORC_SCL, 1, 1, const_cast<double*>(mu), &pm[t]);
        if (sv) sn = (const_mask & 4) ? orc_const_view(g, ORC_VEC, m, 1, sigma + b)
                                      : orc_var(g, ORC_VEC, m, 1, const_cast<double*>(sigma + b), sigma_adj + b);
        else sn = (const_mask & 4) ? orc_const_scalar(g, sigma[0])
                                   : orc_var(g, ORC_SCL, 1, 1, const_cast<double*>(sigma), &ps[t]);
        int root = orc_normal(g, xn, mn, sn);
        orc_bind(g, root, nullptr);
        part[t] = orc_feval(g, root)[0];
        orc_beval(g, root, seed, nullptr);
        orc_free(g);
    });
    double acc = 0;
    for (int t = 0; t < T; ++t) {
        acc += part[t];
        if (!mv && !(const_mask & 2) && mu_adj) mu_adj[0] += pm[t];
        if (!sv && !(const_mask & 4) && sigma_adj) sigma_adj[0] += ps[t];
    }
    return acc;
}

double orc_linreg(size_t rows, size_t cols, const double* X, const double* y, const double* w,
                  double* w_adj, const double* sigma, double* sigma_adj, double seed, int nthreads)
{
    /* normal_adj_log_pdf(y, dot(X, w), sigma): y const vec, X const mat (col-major), w vec var,
     * sigma scalar var.  Threads own row ranges; X's row block is gathered into a private
     * column-major panel because a row range of a column-major matrix is strided. */
    const int T = std::max(nthreads, 1);
    std::vector<double> part(T, 0.), ps(T, 0.);
    std::vector<std::vector<double>> pw(T, std::vector<double>(cols, 0.));
    parallel_ranges(rows, nthreads, [&](int t, size_t b, size_t e) {
        const size_t m = e - b;
        const size_t CH = 4096;
        std::vector<double> panel(CH * cols);
        for (size_t off = 0; off < m; off += CH) {
            const size_t mm = std::min(CH, m - off);
            for (size_t j = 0; j < cols; ++j)
                std::memcpy(&panel[j * mm], X + j * rows + b + off, mm * sizeof(double));
            orc_graph* g = orc_new();
            int Xn = orc_const_view(g, ORC_MAT, mm, cols, panel.data());
            int wn = orc_var(g, ORC_VEC, cols, 1, const_cast<double*>(w), pw[t].data());
            int yn = orc_const_view(g, ORC_VEC, mm, 1, y + b + off);
            int sn = orc_var(g, ORC_SCL, 1, 1, const_cast<double*>(sigma), &ps[t]);
            int root = orc_normal(g, yn, orc_dot(g, Xn, wn), sn);
            orc_bind(g, root, nullptr);
            part[t] += orc_feval(g, root)[0];
            orc_beval(g, root, seed, nullptr);
            orc_free(g);
        }
    });
    double acc = 0;
    for (int t = 0; t < T; ++t) {
        acc += part[t];
        if (sigma_adj) sigma_adj[0] += ps[t];
        for (size_t j = 0; j < cols; ++j) w_adj[j] += pw[t][j];
    }
    return acc;
}

double orc_mlp3(size_t batch, const double* X, const double* y, const double* parm, double* grad,
                int nthreads)
{
    /* README.md:676-726 (the Mathematica-confirmed form of the network in
     * example/ceres_3layer_neural_net/src.cpp:17-57, written WITHOUT placeholders, so the
     * x1 sub-tree appears twice): per-row autodiff, parameter gradients accumulate.
     * The src.cpp variant assigns x1..residual_norm2 to placeholder Vars whose adjoints it
     * never resets between rows, so from the second row on its gradients are polluted
     * (eq.hpp:101-107 accumulates); the README form is the one GenTestData.nb:691-700 pins. */
    const int T = std::max(nthreads, 1);
    std::vector<double> part(T, 0.);
    std::vector<std::vector<double>> pg(T, std::vector<double>(25, 0.));
    parallel_ranges(batch, nthreads, [&](int t, size_t b, size_t e) {
        orc_graph* g = orc_new();
        double* p = const_cast<double*>(parm);
        double* gr = pg[t].data();
        int A1 = orc_var(g, ORC_MAT, 3, 2, p, gr);
        int b1 = orc_var(g, ORC_MAT, 3, 1, p + 6, gr + 6);
        int A2 = orc_var(g, ORC_MAT, 3, 3, p + 9, gr + 9);
        int b2 = orc_var(g, ORC_MAT, 3, 1, p + 18, gr + 18);
        int A3 = orc_var(g, ORC_MAT, 1, 3, p + 21, gr + 21);
        int b3 = orc_var(g, ORC_MAT, 1, 1, p + 24, gr + 24);
        double xrow[2], yrow[1];
        int xi = orc_const_view(g, ORC_MAT, 2, 1, xrow);
        int yi = orc_const_view(g, ORC_MAT, 1, 1, yrow);
        int x1a = orc_binary(g, ORC_ADD, orc_dot(g, A1, xi), b1);
        int x1b = orc_binary(g, ORC_ADD, orc_dot(g, A1, xi), b1);
        int y1 = orc_unary(g, ORC_SIGMOID, x1a);
        int y2 = orc_binary(g, ORC_ADD,
                     orc_unary(g, ORC_SIGMOID, orc_binary(g, ORC_ADD, orc_dot(g, A2, y1), b2)), x1b);
        int y3 = orc_binary(g, ORC_ADD, orc_dot(g, A3, y2), b3);
        int root = orc_pow(g, 2, orc_binary(g, ORC_SUB, yi, y3));
        orc_bind(g, root, nullptr);
        const double one = 1.;
        double loss = 0;
        for (size_t i = b; i < e; ++i) {
            xrow[0] = X[i]; xrow[1] = X[i + batch];
            yrow[0] = y[i];
            loss += orc_feval(g, root)[0];
            orc_beval(g, root, 0., &one);
        }
        part[t] = loss;
        orc_free(g);
    });
    double acc = 0;
    for (int t = 0; t < T; ++t) {
        acc += part[t];
        for (int j = 0; j < 25; ++j) grad[j] += pg[t][j];
    }
    return acc;
}

} // extern "C"
