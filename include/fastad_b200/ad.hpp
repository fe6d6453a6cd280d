// fastad_b200/ad.hpp — C++17 host layer: the ad:: expression-template API of FastAD, lowered
// onto the C ABI of libfastad_b200.so (include/fastad_b200.h).
//
// What stays as in the reference (file:line relative to /root/reference/include/fastad_bits/):
//   ad::Var / ad::VarView (reverse/core/var.hpp, var_view.hpp), ad::constant / constant_view
//   (constant.hpp:158-200), unary builders ad::sin ... ad::tanh and operator- (unary.hpp:299-330),
//   operator+ - * / (binary.hpp:472-498), ad::pow<n> (pow.hpp:205-231), ad::sum / ad::prod
//   (sum.hpp:216-261, prod.hpp:259-306), ad::dot (dot.hpp:133-163), ad::normal_adj_log_pdf
//   (stat/normal.hpp:785-797), placeholders `w = expr` (var_view.hpp:54-63) and operator,
//   (glue.hpp:124-139), ad::bind (bind.hpp:45-49), ad::evaluate / evaluate_adj / autodiff
//   (eval.hpp:18-153).  Expressions are static types built by value; constants fold eagerly.
//
// What changes: values and adjoints live in HBM.  ad::Var owns device buffers plus a lazily
// synchronised host mirror (get()/get_adj() read and write the mirror; it is uploaded before the
// next evaluation and refreshed after one).  ad::VarView(val, adj, ...) wraps DEVICE pointers.
// ad::bind() lowers the expression TYPE into the run-time builder calls of the C ABI and the
// library fuses it into stages; nodes have no feval()/beval() members of their own -- use
// ad::evaluate / ad::evaluate_adj / ad::autodiff on the bound expression (or on the raw
// expression, which binds a temporary).  Eigen is not required (nor present in this image):
// vector/matrix accessors return ad::HostView, seeds are std::vector<double> or device pointers.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../fastad_b200.h"

namespace ad {

// shape tags (util/shape_traits.hpp:7-9)
struct scl {};
struct vec {};
struct mat {};

namespace util {
template <class T> inline constexpr bool is_scl_tag_v = std::is_same_v<T, scl>;
template <class A, class B>
using max_shape_t = std::conditional_t<std::is_same_v<A, mat> || std::is_same_v<B, mat>, mat,
                    std::conditional_t<std::is_same_v<A, vec> || std::is_same_v<B, vec>, vec, scl>>;
template <class S> constexpr int shape_code() { return std::is_same_v<S, scl> ? FADB_SCL : std::is_same_v<S, vec> ? FADB_VEC : FADB_MAT; }

// SizePack / PtrPack look-alikes (util/size_pack.hpp:9, util/ptr_pack.hpp:13-25)
struct SizePack {
    size_t v[2];
    size_t operator()(int i) const { return v[i]; }
};
template <class T>
struct PtrPack {
    T* val;
    T* adj;
};
}  // namespace util

// the library never throws across the C ABI; the host layer turns error codes into exceptions
struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
namespace detail {
inline void check(int rc)
{
    if (rc < 0) throw Error(std::string("fastad_b200: ") + fadb_last_error());
}
inline int checked_id(int rc) { check(rc); return rc; }

// ---------------------------------------------------------------------------------------------
// Storage: one value buffer + one adjoint buffer in HBM with host mirrors.
// Scalars come from a slab pool (a cudaMalloc per ad::Var<double> would cost ~100 us each).
// ---------------------------------------------------------------------------------------------
struct Slab {
    double* base = nullptr;
    size_t used = 0;
    static constexpr size_t cap = 8192;
};
// Per (thread, device) pool with one free list per block size: a Var that dies returns its block, so a loop that
// creates temporaries per iteration (as the reference's examples do) reuses them instead of growing without bound,
// and a block carved on device A is never handed out on device B.
struct Pool {
    std::vector<Slab> slabs;
    std::vector<double*> free_[33];      // by block size in doubles (2 .. 32)
};
inline Pool& pool_of(int dev)
{
    static thread_local std::map<int, Pool> pools;
    return pools[dev];
}
inline double* pool_take(int dev, size_t n)
{
    Pool& P = pool_of(dev);
    if (!P.free_[n].empty()) { double* p = P.free_[n].back(); P.free_[n].pop_back(); return p; }
    if (P.slabs.empty() || P.slabs.back().used + n > Slab::cap) {
        Slab s;
        void* p = nullptr;
        check(fadb_malloc(Slab::cap * sizeof(double), &p));
        s.base = static_cast<double*>(p);
        P.slabs.push_back(s);
    }
    double* r = P.slabs.back().base + P.slabs.back().used;
    P.slabs.back().used += n;
    return r;
}
inline void pool_give(int dev, double* p, size_t n) { pool_of(dev).free_[n].push_back(p); }

struct Storage {
    double* dval = nullptr;
    double* dadj = nullptr;
    size_t n = 0;
    int own = 0;   // 0: caller's device pointers, 1: fadb_malloc, 2: pooled
    int dev = 0;   // device the pooled block was carved on
    std::vector<double> hval, hadj;
    bool host_val_newer = false, host_adj_newer = false, dev_val_newer = false, dev_adj_newer = false;

    Storage(size_t n_, bool owning) : n(n_)
    {
        if (!owning) return;
        if (n <= 16) {
            dev = fadb_get_device();
            check(dev);
            dval = pool_take(dev, 2 * n);
            dadj = dval + n;
            own = 2;
        } else {
            void* p = nullptr;
            check(fadb_malloc(2 * n * sizeof(double), &p));
            dval = static_cast<double*>(p);
            dadj = dval + n;
            own = 1;
        }
        hval.assign(n, 0.0);
        hadj.assign(n, 0.0);
        host_val_newer = host_adj_newer = true;   // zero-initialised like var.hpp:33-36
    }
    Storage(double* v, double* a, size_t n_) : dval(v), dadj(a), n(n_) {}
    ~Storage()
    {
        if (own == 1) fadb_free(dval);
        else if (own == 2) pool_give(dev, dval, 2 * n);
    }
    Storage(const Storage&) = delete;
    Storage& operator=(const Storage&) = delete;

    void ensure_mirror()
    {
        if (hval.size() != n) { hval.assign(n, 0.0); dev_val_newer = dval != nullptr; }
        if (hadj.size() != n) { hadj.assign(n, 0.0); dev_adj_newer = dadj != nullptr; }
    }
    void to_device()
    {
        if (host_val_newer && dval) { check(fadb_h2d(dval, hval.data(), n * sizeof(double))); host_val_newer = false; }
        if (host_adj_newer && dadj) { check(fadb_h2d(dadj, hadj.data(), n * sizeof(double))); host_adj_newer = false; }
    }
    double* host_val(bool will_write)
    {
        ensure_mirror();
        if (dev_val_newer && !host_val_newer) check(fadb_d2h(hval.data(), dval, n * sizeof(double)));
        dev_val_newer = false;
        if (will_write) host_val_newer = true;
        return hval.data();
    }
    double* host_adj(bool will_write)
    {
        ensure_mirror();
        if (dev_adj_newer && !host_adj_newer && dadj) check(fadb_d2h(hadj.data(), dadj, n * sizeof(double)));
        dev_adj_newer = false;
        if (will_write) host_adj_newer = true;
        return hadj.data();
    }
    void zero_adj()
    {
        if (dadj) check(fadb_memset_zero(dadj, n * sizeof(double)));
        if (hadj.size() == n) std::fill(hadj.begin(), hadj.end(), 0.0);
        host_adj_newer = dev_adj_newer = false;
    }
};

// lowering context: the C-ABI builder handle + every Storage the expression touches
struct Lowering {
    fadb_expr* e = nullptr;
    bool device = true;    // false: host-only planning (bind_cache_size), no CUDA call is made
    std::vector<std::shared_ptr<Storage>> storages;
    void touch(const std::shared_ptr<Storage>& s)
    {
        for (auto& t : storages) if (t == s) return;
        storages.push_back(s);
    }
};
}  // namespace detail

// ---------------------------------------------------------------------------------------------
// HostView: what get()/get_adj() return for vec / mat instead of Eigen::Map (column-major).
// ---------------------------------------------------------------------------------------------
struct HostView {
    double* p;
    size_t r, c;
    double* data() const { return p; }
    size_t rows() const { return r; }
    size_t cols() const { return c; }
    size_t size() const { return r * c; }
    double& operator()(size_t i) const { assert(i < r * c); return p[i]; }
    double& operator[](size_t i) const { return p[i]; }
    double& operator()(size_t i, size_t j) const { assert(i < r && j < c); return p[i + j * r]; }
    double* begin() const { return p; }
    double* end() const { return p + r * c; }
    void setZero() const { std::fill(p, p + r * c, 0.0); }
    void setOnes() const { std::fill(p, p + r * c, 1.0); }
    const HostView& operator=(const std::vector<double>& v) const
    {
        assert(v.size() == r * c);
        std::copy(v.begin(), v.end(), p);
        return *this;
    }
    const HostView& operator=(std::initializer_list<double> v) const
    {
        assert(v.size() == r * c);
        std::copy(v.begin(), v.end(), p);
        return *this;
    }
    double sum() const { double a = 0; for (size_t i = 0; i < r * c; ++i) a += p[i]; return a; }
    std::vector<double> to_vector() const { return std::vector<double>(p, p + r * c); }
};

namespace core {

// CRTP base (reverse/core/expr_base.hpp:16-23)
struct ExprTag {};
template <class Derived>
struct ExprBase;
template <class ExprType> struct ExprBind;
template <class D> ExprBind<D>& own_binding(const ExprBase<D>&);
template <class Derived>
struct ExprBase : ExprTag {
    Derived& self() { return static_cast<Derived&>(*this); }
    const Derived& self() const { return static_cast<const Derived&>(*this); }
    // The reference's manual binding protocol (e.g. unary.hpp:95-116, README.md:284-306; example/main.cpp):
    //   auto sz = expr.bind_cache_size();  std::vector<double> v(sz(0)), a(sz(1));
    //   expr.bind_cache({v.data(), a.data()});  autodiff(expr);
    // Temporaries live in registers and stage boundaries in HBM buffers the library allocates itself, so the
    // sizes are what the planner reports and the caller's host buffers are accepted but not used; the
    // returned pack is advanced past them like the reference's.
    util::SizePack bind_cache_size() const;
    util::PtrPack<double> bind_cache(util::PtrPack<double> begin) const;
    // The node-level protocol of the reference (unary.hpp:63-89, binary.hpp:100-136; README.md:284-306):
    //   expr.bind_cache({v, a});  auto f = expr.feval();  expr.beval(seed);
    // feval() = ad::evaluate, beval(seed) = ad::evaluate_adj on a binding the node makes on first use and keeps
    // (copies of the node share it, like copies of a reference node share the cache they were bound to).
    // Returns double for a scalar expression, std::vector<double> (column-major) otherwise.
    auto feval() const;
    void beval(double seed = 1.) const;
    void beval(const std::vector<double>& seed) const;

private:
    mutable std::shared_ptr<void> own_binding_;      // type-erased ExprBind<Derived>
    template <class D> friend ExprBind<D>& own_binding(const ExprBase<D>&);
};

// Var<T,S> derives from ExprBase<VarView<T,S>>: anything below an ExprBase is an expression
template <class T> inline constexpr bool is_expr_v = std::is_base_of_v<ExprTag, std::decay_t<T>>;

template <class ValueType, class ShapeType> struct Constant;
template <class ValueType, class ShapeType> struct ConstantView;

template <class T> struct is_constant : std::false_type {};
template <class V, class S> struct is_constant<Constant<V, S>> : std::true_type {};
template <class V, class S> struct is_constant<ConstantView<V, S>> : std::true_type {};
template <class T> inline constexpr bool is_constant_v = is_constant<std::decay_t<T>>::value;
template <class T> inline constexpr bool is_scl_constant_v = false;
template <class V> inline constexpr bool is_scl_constant_v<Constant<V, scl>> = true;

template <class VarViewType, class ExprType> struct EqNode;

}  // namespace core

template <class ValueType, class ShapeType = scl> struct VarView;
template <class ValueType, class ShapeType = scl> struct Var;

namespace util {
template <class T> struct is_var_view : std::false_type {};
template <class V, class S> struct is_var_view<VarView<V, S>> : std::true_type {};
template <class T> inline constexpr bool is_var_view_v = is_var_view<std::decay_t<T>>::value;
template <class T> struct is_var : std::false_type {};
template <class V, class S> struct is_var<Var<V, S>> : std::true_type {};
template <class T> inline constexpr bool is_var_v = is_var<std::decay_t<T>>::value;

// convert_to_ad_t (util/type_traits.hpp:196-246): Var -> VarView, arithmetic -> Constant<scl>
template <class T, class = void> struct convert_to_ad;
template <class T> struct convert_to_ad<T, std::enable_if_t<is_var_v<T>>> {
    using type = VarView<typename T::value_t, typename T::shape_t>;
};
template <class T> struct convert_to_ad<T, std::enable_if_t<core::is_expr_v<T> && !is_var_v<T>>> { using type = T; };
template <class T> struct convert_to_ad<T, std::enable_if_t<std::is_arithmetic_v<T>>> {
    using type = core::Constant<double, scl>;
};
template <class T> using convert_to_ad_t = typename convert_to_ad<std::decay_t<T>>::type;
template <class T, class = void> struct is_convertible_to_ad : std::false_type {};
template <class T> struct is_convertible_to_ad<T, std::void_t<convert_to_ad_t<T>>> : std::true_type {};
template <class T> inline constexpr bool is_convertible_to_ad_v = is_convertible_to_ad<T>::value;
template <class... Ts> inline constexpr bool any_ad_v = (core::is_expr_v<Ts> || ...);

template <class T> struct expr_traits { using value_t = typename T::value_t; using shape_t = typename T::shape_t; };
template <class T> inline constexpr bool is_scl_v = std::is_same_v<typename std::decay_t<T>::shape_t, scl>;
template <class T> inline constexpr bool is_vec_v = std::is_same_v<typename std::decay_t<T>::shape_t, vec>;
template <class T> inline constexpr bool is_mat_v = std::is_same_v<typename std::decay_t<T>::shape_t, mat>;
}  // namespace util

// ---------------------------------------------------------------------------------------------
// Leaves
// ---------------------------------------------------------------------------------------------
template <class ValueType, class ShapeType>
struct VarView : core::ExprBase<VarView<ValueType, ShapeType>> {
    static_assert(std::is_same_v<ValueType, double>, "fastad_b200 computes in IEEE double only");
    using value_t = ValueType;
    using shape_t = ShapeType;
    using var_view_t = VarView<ValueType, ShapeType>;

    VarView() : VarView(nullptr, nullptr, 1, 1) {}
    // DEVICE pointers (var_view.hpp:130-188 takes host pointers; storage now lives in HBM)
    VarView(value_t* dev_val, value_t* dev_adj, size_t rows = 1, size_t cols = 1)
        : st_(std::make_shared<detail::Storage>(dev_val, dev_adj, rows * (util::is_scl_tag_v<ShapeType> ? 1 : cols))),
          off_(0), rows_(util::is_scl_tag_v<ShapeType> ? 1 : rows),
          cols_(std::is_same_v<ShapeType, mat> ? cols : 1) {}
    VarView(std::shared_ptr<detail::Storage> st, size_t off, size_t rows, size_t cols)
        : st_(std::move(st)), off_(off), rows_(rows), cols_(cols) {}

    size_t rows() const { return rows_; }
    size_t cols() const { return cols_; }
    size_t size() const { return rows_ * cols_; }
    value_t* data() const { return st_->dval ? st_->dval + off_ : nullptr; }          // device pointer
    value_t* data_adj() const { return st_->dadj ? st_->dadj + off_ : nullptr; }      // device pointer

    // host mirror access (value_view.hpp / value_adj_view.hpp get(), get_adj())
    decltype(auto) get() { return view(st_->host_val(true) + off_); }
    decltype(auto) get() const { return view(st_->host_val(false) + off_); }
    decltype(auto) get_adj() { return view(st_->host_adj(true) + off_); }
    decltype(auto) get_adj() const { return view(st_->host_adj(false) + off_); }
    value_t& get(size_t i, size_t j) { return (st_->host_val(true) + off_)[i + j * rows_]; }
    value_t& get_adj(size_t i, size_t j) { return (st_->host_adj(true) + off_)[i + j * rows_]; }
    value_t get(size_t i, size_t j) const { return (st_->host_val(false) + off_)[i + j * rows_]; }
    value_t get_adj(size_t i, size_t j) const { return (st_->host_adj(false) + off_)[i + j * rows_]; }
    void reset_adj()
    {
        if (off_ == 0 && size() == st_->n) { st_->zero_adj(); return; }
        double* h = st_->host_adj(true) + off_;
        std::fill(h, h + size(), 0.0);
    }
    void zero_adj() { reset_adj(); }
    // re-point the view (ValueAdjView::bind, value_adj_view.hpp:47-52); device pointers
    util::PtrPack<value_t> bind(util::PtrPack<value_t> begin)
    {
        st_ = std::make_shared<detail::Storage>(begin.val, begin.adj, size());
        off_ = 0;
        return {begin.val ? begin.val + size() : nullptr, begin.adj ? begin.adj + size() : nullptr};
    }
    // push the host mirror to HBM now / mark device contents as changed by foreign kernels
    void upload() const { st_->to_device(); }
    void invalidate_host() const { st_->dev_val_newer = st_->dev_adj_newer = true; st_->host_val_newer = st_->host_adj_newer = false; }

    // placeholder expression  w = expr  (var_view.hpp:54-63)
    template <class Derived, class = std::enable_if_t<util::is_convertible_to_ad_v<Derived> &&
                                                      !std::is_same_v<std::decay_t<Derived>, var_view_t>>>
    auto operator=(const Derived& x) const
    {
        using expr_t = util::convert_to_ad_t<Derived>;
        expr_t expr = x;
        return core::EqNode<var_view_t, expr_t>(*this, expr);
    }
    VarView(const VarView&) = default;
    VarView(VarView&&) = default;
    VarView& operator=(const VarView&) = default;
    VarView& operator=(VarView&&) = default;

    // vec element / sub-views (var_view.hpp:154-172)
    template <class S = ShapeType, class = std::enable_if_t<std::is_same_v<S, vec>>>
    auto operator()(size_t i) const { assert(i < size()); return VarView<value_t, scl>(st_, off_ + i, 1, 1); }
    template <class S = ShapeType, class = std::enable_if_t<std::is_same_v<S, vec>>>
    auto operator[](size_t i) const { return operator()(i); }
    template <class S = ShapeType, class = std::enable_if_t<std::is_same_v<S, vec>>>
    auto head(size_t n) const { assert(n <= size()); return VarView<value_t, vec>(st_, off_, n, 1); }
    template <class S = ShapeType, class = std::enable_if_t<std::is_same_v<S, vec>>>
    auto tail(size_t n) const { assert(n <= size()); return VarView<value_t, vec>(st_, off_ + size() - n, n, 1); }

    int lower(detail::Lowering& L) const
    {
        if (L.device) st_->to_device();
        L.touch(st_);
        return detail::checked_id(fadb_expr_var(L.e, util::shape_code<ShapeType>(), rows_, cols_, data(), data_adj()));
    }
    const std::shared_ptr<detail::Storage>& storage() const { return st_; }

protected:
    decltype(auto) view(double* p) const
    {
        if constexpr (util::is_scl_tag_v<ShapeType>) return (*p);
        else return HostView{p, rows_, cols_};
    }
    std::shared_ptr<detail::Storage> st_;
    size_t off_, rows_, cols_;
};

// Var owns its value and adjoint (var.hpp:23-220); copies are deep, zero-initialised.
template <class ValueType, class ShapeType>
struct Var : VarView<ValueType, ShapeType> {
    using base_t = VarView<ValueType, ShapeType>;
    using value_t = ValueType;
    using shape_t = ShapeType;
    using base_t::operator=;

    template <class S = ShapeType, class = std::enable_if_t<util::is_scl_tag_v<S>>>
    Var() : base_t(std::make_shared<detail::Storage>(1, true), 0, 1, 1) {}
    template <class S = ShapeType, class = std::enable_if_t<util::is_scl_tag_v<S>>>
    explicit Var(value_t v) : base_t(std::make_shared<detail::Storage>(1, true), 0, 1, 1) { this->get() = v; }
    template <class S = ShapeType, class = std::enable_if_t<std::is_same_v<S, vec>>>
    explicit Var(size_t n) : base_t(std::make_shared<detail::Storage>(n, true), 0, n, 1) {}
    template <class S = ShapeType, class = std::enable_if_t<std::is_same_v<S, mat>>>
    Var(size_t rows, size_t cols) : base_t(std::make_shared<detail::Storage>(rows * cols, true), 0, rows, cols) {}

    Var(const Var& v) : base_t(std::make_shared<detail::Storage>(v.size(), true), 0, v.rows(), v.cols()) { copy_from(v); }
    Var(Var&& v) = default;
    Var& operator=(const Var& v)
    {
        if (this == &v) return *this;
        assert(v.rows() == this->rows() && v.cols() == this->cols());
        copy_from(v);
        return *this;
    }
    Var& operator=(Var&& v) = default;

private:
    void copy_from(const Var& v)
    {
        const double* sv = v.storage()->host_val(false);
        const double* sa = v.storage()->host_adj(false);
        std::copy(sv, sv + v.size(), this->st_->host_val(true));
        std::copy(sa, sa + v.size(), this->st_->host_adj(true));
    }
};

namespace core {

// Constant<T, scl> holds a host double and folds eagerly; Constant<T, vec|mat> owns a device copy
// of host data; ConstantView views caller-owned DEVICE memory (constant.hpp:26-151).
template <class ValueType>
struct Constant<ValueType, scl> : ExprBase<Constant<ValueType, scl>> {
    using value_t = double;
    using shape_t = scl;
    Constant(double c) : c_(c) {}
    double feval() const { return c_; }
    double get() const { return c_; }
    size_t rows() const { return 1; }
    size_t cols() const { return 1; }
    size_t size() const { return 1; }
    int lower(detail::Lowering& L) const { return detail::checked_id(fadb_expr_const_scalar(L.e, c_)); }
private:
    double c_;
};

template <class ValueType, class ShapeType>
struct Constant : ExprBase<Constant<ValueType, ShapeType>> {
    using value_t = double;
    using shape_t = ShapeType;
    Constant(const double* host, size_t rows, size_t cols)
        : st_(std::make_shared<detail::Storage>(rows * cols, true)), rows_(rows), cols_(cols)
    {
        std::copy(host, host + rows * cols, st_->host_val(true));
    }
    HostView get() const { return HostView{st_->host_val(false), rows_, cols_}; }
    size_t rows() const { return rows_; }
    size_t cols() const { return cols_; }
    size_t size() const { return rows_ * cols_; }
    int lower(detail::Lowering& L) const
    {
        if (L.device) st_->to_device();
        L.touch(st_);
        return detail::checked_id(fadb_expr_const_view(L.e, util::shape_code<ShapeType>(), rows_, cols_, st_->dval));
    }
    const std::shared_ptr<detail::Storage>& storage() const { return st_; }
private:
    std::shared_ptr<detail::Storage> st_;
    size_t rows_, cols_;
};

template <class ValueType, class ShapeType>
struct ConstantView : ExprBase<ConstantView<ValueType, ShapeType>> {
    static_assert(!util::is_scl_tag_v<ShapeType>, "ConstantView is for vec/mat (constant.hpp:31-33)");
    using value_t = double;
    using shape_t = ShapeType;
    ConstantView(const double* dev, size_t rows, size_t cols) : p_(dev), rows_(rows), cols_(cols) {}
    const double* data() const { return p_; }
    void bind(const double* dev) { p_ = dev; }
    size_t rows() const { return rows_; }
    size_t cols() const { return cols_; }
    size_t size() const { return rows_ * cols_; }
    int lower(detail::Lowering& L) const
    {
        return detail::checked_id(fadb_expr_const_view(L.e, util::shape_code<ShapeType>(), rows_, cols_, p_));
    }
private:
    const double* p_;
    size_t rows_, cols_;
};

// ---------------------------------------------------------------------------------------------
// Nodes.  Each holds its children BY VALUE like the reference and knows how to lower itself.
// ---------------------------------------------------------------------------------------------
template <int Code> struct UnaryTag { static constexpr int code = Code; };
using UnaryMinus = UnaryTag<FADB_NEG>;   using Sin = UnaryTag<FADB_SIN>;     using Cos = UnaryTag<FADB_COS>;
using Tan = UnaryTag<FADB_TAN>;          using Arcsin = UnaryTag<FADB_ASIN>; using Arccos = UnaryTag<FADB_ACOS>;
using Arctan = UnaryTag<FADB_ATAN>;      using Exp = UnaryTag<FADB_EXP>;     using Log = UnaryTag<FADB_LOG>;
using Sqrt = UnaryTag<FADB_SQRT>;        using Erf = UnaryTag<FADB_ERF>;     using Sigmoid = UnaryTag<FADB_SIGMOID>;
using Sinh = UnaryTag<FADB_SINH>;        using Cosh = UnaryTag<FADB_COSH>;   using Tanh = UnaryTag<FADB_TANH>;
using MockUnary = UnaryTag<FADB_MOCK2X>;   // f(x) = 2x, test/testutil/base_fixture.hpp:9-22

inline double fold_unary(int code, double x)
{   // host fmap for eager constant folding (unary.hpp:182-184)
    switch (code) {
    case FADB_NEG: return -x; case FADB_SIN: return std::sin(x); case FADB_COS: return std::cos(x);
    case FADB_TAN: return std::tan(x); case FADB_ASIN: return std::asin(x); case FADB_ACOS: return std::acos(x);
    case FADB_ATAN: return std::atan(x); case FADB_EXP: return std::exp(x); case FADB_LOG: return std::log(x);
    case FADB_SQRT: return std::sqrt(x); case FADB_ERF: return std::erf(x);
    case FADB_SIGMOID: return 1 / (1 + std::exp(-x)); case FADB_SINH: return std::sinh(x);
    case FADB_COSH: return std::cosh(x); case FADB_TANH: return std::tanh(x); default: return 2 * x;
    }
}

template <class Unary, class ExprType>
struct UnaryNode : ExprBase<UnaryNode<Unary, ExprType>> {
    using value_t = double;
    using shape_t = typename ExprType::shape_t;
    UnaryNode(const ExprType& e) : expr_(e) {}
    size_t rows() const { return expr_.rows(); }
    size_t cols() const { return expr_.cols(); }
    size_t size() const { return expr_.size(); }
    int lower(detail::Lowering& L) const { return detail::checked_id(fadb_expr_unary(L.e, Unary::code, expr_.lower(L))); }
    const ExprType& child() const { return expr_; }
private:
    ExprType expr_;
};

template <int Code> struct BinaryTag { static constexpr int code = Code; };
using Add = BinaryTag<FADB_ADD>; using Sub = BinaryTag<FADB_SUB>; using Mul = BinaryTag<FADB_MUL>;
using Div = BinaryTag<FADB_DIV>; using MockBinary = BinaryTag<FADB_MOCKB>;
// comparison / logical functors (binary.hpp:347-470): value 0 or 1, no adjoint
using LessThan = BinaryTag<FADB_LT>; using LessThanEq = BinaryTag<FADB_LE>; using GreaterThan = BinaryTag<FADB_GT>;
using GreaterThanEq = BinaryTag<FADB_GE>; using Equal = BinaryTag<FADB_EQ>; using NotEqual = BinaryTag<FADB_NE>;
using LogicalAnd = BinaryTag<FADB_AND>; using LogicalOr = BinaryTag<FADB_OR>;

inline double fold_binary(int code, double x, double y)
{
    switch (code) {
    case FADB_ADD: return x + y; case FADB_SUB: return x - y; case FADB_MUL: return x * y;
    case FADB_DIV: return x / y;
    case FADB_LT: return x < y; case FADB_LE: return x <= y; case FADB_GT: return x > y; case FADB_GE: return x >= y;
    case FADB_EQ: return x == y; case FADB_NE: return x != y; case FADB_AND: return x != 0 && y != 0;
    case FADB_OR: return x != 0 || y != 0;
    default: return x - 2 * y;
    }
}

template <class Binary, class L, class R>
struct BinaryNode : ExprBase<BinaryNode<Binary, L, R>> {
    using value_t = double;
    using shape_t = util::max_shape_t<typename L::shape_t, typename R::shape_t>;
    static_assert(util::is_scl_v<L> || util::is_scl_v<R> || std::is_same_v<typename L::shape_t, typename R::shape_t>,
                  "binary node: one side scalar, or both vec, or both mat (binary.hpp:63-68)");
    BinaryNode(const L& l, const R& r) : l_(l), r_(r)
    {
        if constexpr (!util::is_scl_v<L> && !util::is_scl_v<R>) assert(l.rows() == r.rows() && l.cols() == r.cols());
    }
    size_t rows() const { return std::max(l_.rows(), r_.rows()); }
    size_t cols() const { return std::max(l_.cols(), r_.cols()); }
    size_t size() const { return rows() * cols(); }
    int lower(detail::Lowering& Lw) const
    {
        int a = l_.lower(Lw);
        int b = r_.lower(Lw);
        return detail::checked_id(fadb_expr_binary(Lw.e, Binary::code, a, b));
    }
    const L& lhs() const { return l_; }
    const R& rhs() const { return r_; }
private:
    L l_;
    R r_;
};

template <int64_t N, class ExprType>
struct PowNode : ExprBase<PowNode<N, ExprType>> {
    using value_t = double;
    using shape_t = typename ExprType::shape_t;
    PowNode(const ExprType& e) : expr_(e) {}
    size_t rows() const { return expr_.rows(); }
    size_t cols() const { return expr_.cols(); }
    size_t size() const { return expr_.size(); }
    int lower(detail::Lowering& L) const { return detail::checked_id(fadb_expr_pow(L.e, (long)N, expr_.lower(L))); }
    const ExprType& child() const { return expr_; }
private:
    ExprType expr_;
};

template <bool IsProd, class ExprType>
struct ReduceElemNode : ExprBase<ReduceElemNode<IsProd, ExprType>> {
    using value_t = double;
    using shape_t = scl;
    ReduceElemNode(const ExprType& e) : expr_(e) {}
    size_t rows() const { return 1; }
    size_t cols() const { return 1; }
    size_t size() const { return 1; }
    int lower(detail::Lowering& L) const
    {
        int c = expr_.lower(L);
        return detail::checked_id(IsProd ? fadb_expr_prod(L.e, c) : fadb_expr_sum(L.e, c));
    }
    const ExprType& child() const { return expr_; }
private:
    ExprType expr_;
};
template <class E> using SumElemNode = ReduceElemNode<false, E>;
template <class E> using ProdElemNode = ReduceElemNode<true, E>;

template <bool IsProd, class ExprType>
struct ReduceIterNode : ExprBase<ReduceIterNode<IsProd, ExprType>> {
    using value_t = double;
    using shape_t = typename ExprType::shape_t;
    ReduceIterNode(std::vector<ExprType> exprs) : exprs_(std::move(exprs)) {}
    size_t rows() const { return exprs_.empty() ? 1 : exprs_[0].rows(); }
    size_t cols() const { return exprs_.empty() ? 1 : exprs_[0].cols(); }
    size_t size() const { return rows() * cols(); }
    int lower(detail::Lowering& L) const
    {
        std::vector<int> ids;
        ids.reserve(exprs_.size());
        for (auto& x : exprs_) ids.push_back(x.lower(L));
        return detail::checked_id(IsProd ? fadb_expr_prod_iter(L.e, (int)ids.size(), ids.data())
                                         : fadb_expr_sum_iter(L.e, (int)ids.size(), ids.data()));
    }
private:
    std::vector<ExprType> exprs_;
};
template <class E> using SumIterNode = ReduceIterNode<false, E>;
template <class E> using ProdIterNode = ReduceIterNode<true, E>;

template <class L, class R>
struct DotNode : ExprBase<DotNode<L, R>> {
    static_assert(util::is_mat_v<L> && (util::is_vec_v<R> || util::is_mat_v<R>), "dot: mat x vec or mat x mat (dot.hpp:19-33)");
    using value_t = double;
    using shape_t = std::conditional_t<util::is_vec_v<R>, vec, mat>;
    DotNode(const L& l, const R& r) : l_(l), r_(r) { assert(l.cols() == r.rows()); }
    size_t rows() const { return l_.rows(); }
    size_t cols() const { return r_.cols(); }
    size_t size() const { return rows() * cols(); }
    int lower(detail::Lowering& Lw) const
    {
        int a = l_.lower(Lw);
        int b = r_.lower(Lw);
        return detail::checked_id(fadb_expr_dot(Lw.e, a, b));
    }
private:
    L l_;
    R r_;
};

template <class X, class M, class S>
struct NormalAdjLogPDFNode : ExprBase<NormalAdjLogPDFNode<X, M, S>> {
    using value_t = double;
    using shape_t = scl;
    NormalAdjLogPDFNode(const X& x, const M& m, const S& s) : x_(x), m_(m), s_(s) {}
    size_t rows() const { return 1; }
    size_t cols() const { return 1; }
    size_t size() const { return 1; }
    int lower(detail::Lowering& L) const
    {
        int a = x_.lower(L);
        int b = m_.lower(L);
        int c = s_.lower(L);
        return detail::checked_id(fadb_expr_normal(L.e, a, b, c));
    }
    const X& x() const { return x_; }
    const M& mean() const { return m_; }
    const S& sigma() const { return s_; }
private:
    X x_;
    M m_;
    S s_;
};

// Cauchy / Uniform / Bernoulli adjusted log-pdf nodes (reverse/stat/{cauchy,uniform,bernoulli}.hpp)
template <int Dist, class A, class B, class C>
struct LogPDFNode : ExprBase<LogPDFNode<Dist, A, B, C>> {
    using value_t = double;
    using shape_t = scl;
    LogPDFNode(const A& a, const B& b, const C& c) : a_(a), b_(b), c_(c) {}
    size_t rows() const { return 1; }
    size_t cols() const { return 1; }
    size_t size() const { return 1; }
    int lower(detail::Lowering& L) const
    {
        int a = a_.lower(L);
        int b = b_.lower(L);
        int c = Dist == FADB_BERNOULLI ? -1 : c_.lower(L);
        return detail::checked_id(fadb_expr_logpdf(L.e, Dist, a, b, c));
    }
    const A& first() const { return a_; }
    const B& second() const { return b_; }
    const C& third() const { return c_; }
private:
    A a_;
    B b_;
    C c_;
};

template <class VarViewType, class ExprType>
struct EqNode : ExprBase<EqNode<VarViewType, ExprType>> {
    using value_t = double;
    using shape_t = typename VarViewType::shape_t;
    static_assert(std::is_same_v<typename VarViewType::shape_t, typename ExprType::shape_t>,
                  "placeholder and expression must have the same shape (eq.hpp:57-60)");
    EqNode(const VarViewType& v, const ExprType& e) : v_(v), e_(e) { assert(v.rows() == e.rows() && v.cols() == e.cols()); }
    size_t rows() const { return v_.rows(); }
    size_t cols() const { return v_.cols(); }
    size_t size() const { return v_.size(); }
    int lower(detail::Lowering& L) const
    {
        int c = e_.lower(L);
        int w = v_.lower(L);
        return detail::checked_id(fadb_expr_eq(L.e, w, c));
    }
    const VarViewType& var() const { return v_; }
    const ExprType& expr() const { return e_; }
private:
    VarViewType v_;
    ExprType e_;
};

template <class L, class R>
struct GlueNode : ExprBase<GlueNode<L, R>> {
    using value_t = double;
    using shape_t = typename R::shape_t;
    GlueNode(const L& l, const R& r) : l_(l), r_(r) {}
    size_t rows() const { return r_.rows(); }
    size_t cols() const { return r_.cols(); }
    size_t size() const { return r_.size(); }
    int lower(detail::Lowering& Lw) const
    {
        int a = l_.lower(Lw);
        int b = r_.lower(Lw);
        return detail::checked_id(fadb_expr_glue(Lw.e, a, b));
    }
    const L& lhs() const { return l_; }
    const R& rhs() const { return r_; }
private:
    L l_;
    R r_;
};

// DetNode / LogDetNode (det.hpp:25-92, log_det.hpp:25-92): determinant resp. log|determinant| of a
// square matrix expression; DecompType carries the reference's decomposition policy as a code
template <class DecompType, class ExprType, bool TakeLog>
struct DetNodeBase : ExprBase<DetNodeBase<DecompType, ExprType, TakeLog>> {
    static_assert(!util::is_scl_v<ExprType>, "det of a scalar (det.hpp:36)");
    using value_t = double;
    using shape_t = scl;
    DetNodeBase(const ExprType& e) : expr_(e) { assert(e.rows() == e.cols()); }
    size_t rows() const { return 1; }
    size_t cols() const { return 1; }
    size_t size() const { return 1; }
    int lower(detail::Lowering& L) const
    {
        return detail::checked_id(fadb_expr_det(L.e, DecompType::code, TakeLog ? 1 : 0, expr_.lower(L)));
    }
    const ExprType& child() const { return expr_; }
private:
    ExprType expr_;
};
template <class DecompType, class ExprType> using DetNode = DetNodeBase<DecompType, ExprType, false>;
template <class DecompType, class ExprType> using LogDetNode = DetNodeBase<DecompType, ExprType, true>;

// NormNode (norm.hpp:24-108): squared L2 / Frobenius norm
template <class ExprType>
struct NormNode : ExprBase<NormNode<ExprType>> {
    static_assert(!util::is_scl_v<ExprType>, "norm of a scalar (norm.hpp:34)");
    using value_t = double;
    using shape_t = scl;
    NormNode(const ExprType& e) : expr_(e) {}
    size_t rows() const { return 1; }
    size_t cols() const { return 1; }
    size_t size() const { return 1; }
    int lower(detail::Lowering& L) const { return detail::checked_id(fadb_expr_norm(L.e, expr_.lower(L))); }
    const ExprType& child() const { return expr_; }
private:
    ExprType expr_;
};

// TransposeNode (transpose.hpp:17-60)
template <class ExprType>
struct TransposeNode : ExprBase<TransposeNode<ExprType>> {
    static_assert(!util::is_scl_v<ExprType>, "transpose of a scalar (transpose.hpp:25)");
    using value_t = double;
    using shape_t = mat;
    TransposeNode(const ExprType& e) : expr_(e) {}
    size_t rows() const { return expr_.cols(); }
    size_t cols() const { return expr_.rows(); }
    size_t size() const { return expr_.size(); }
    int lower(detail::Lowering& L) const { return detail::checked_id(fadb_expr_transpose(L.e, expr_.lower(L))); }
private:
    ExprType expr_;
};

// IfElseNode (if_else.hpp:29-114): scalar condition, branches of one shape
template <class C, class I, class E>
struct IfElseNode : ExprBase<IfElseNode<C, I, E>> {
    static_assert(util::is_scl_v<C> && std::is_same_v<typename I::shape_t, typename E::shape_t>,
                  "if_else: scalar condition, same-shape branches (if_else.hpp:52-55)");
    using value_t = double;
    using shape_t = typename I::shape_t;
    IfElseNode(const C& c, const I& i, const E& e) : c_(c), i_(i), e_(e) { assert(i.rows() == e.rows() && i.cols() == e.cols()); }
    size_t rows() const { return i_.rows(); }
    size_t cols() const { return i_.cols(); }
    size_t size() const { return i_.size(); }
    int lower(detail::Lowering& L) const
    {
        int c = c_.lower(L);
        int a = i_.lower(L);
        int b = e_.lower(L);
        return detail::checked_id(fadb_expr_if_else(L.e, c, a, b));
    }
private:
    C c_;
    I i_;
    E e_;
};

// OpEqNode (eq.hpp:194-303): w += / -= / *= / /= expr
template <class Op, class VarViewType, class ExprType>
struct OpEqNode : ExprBase<OpEqNode<Op, VarViewType, ExprType>> {
    using value_t = double;
    using shape_t = typename VarViewType::shape_t;
    OpEqNode(const VarViewType& v, const ExprType& e) : v_(v), e_(e)
    {
        if constexpr (!util::is_scl_v<ExprType>) assert(v.rows() == e.rows() && v.cols() == e.cols());
    }
    size_t rows() const { return v_.rows(); }
    size_t cols() const { return v_.cols(); }
    size_t size() const { return v_.size(); }
    int lower(detail::Lowering& L) const
    {
        int w = v_.lower(L);
        int c = e_.lower(L);
        return detail::checked_id(fadb_expr_opeq(L.e, Op::code, w, c));
    }
private:
    VarViewType v_;
    ExprType e_;
};
using AddEq = BinaryTag<FADB_ADD>; using SubEq = BinaryTag<FADB_SUB>; using MulEq = BinaryTag<FADB_MUL>; using DivEq = BinaryTag<FADB_DIV>;

// ---- operators live in ad::core for ADL (binary.hpp:472-498, unary.hpp:297, glue.hpp:124-139)
#define FASTAD_B200_BINARY_OP(name, tag)                                                              \
    template <class A, class B,                                                                      \
              class = std::enable_if_t<util::is_convertible_to_ad_v<A> && util::is_convertible_to_ad_v<B> && \
                                       util::any_ad_v<A, B>>>                                         \
    inline auto name(const A& a, const B& b)                                                          \
    {                                                                                                 \
        using ea_t = util::convert_to_ad_t<A>;                                                        \
        using eb_t = util::convert_to_ad_t<B>;                                                        \
        ea_t ea = a;                                                                                  \
        eb_t eb = b;                                                                                  \
        if constexpr (is_scl_constant_v<ea_t> && is_scl_constant_v<eb_t>)                             \
            return Constant<double, scl>(fold_binary(tag::code, ea.feval(), eb.feval()));             \
        else                                                                                          \
            return BinaryNode<tag, ea_t, eb_t>(ea, eb);                                               \
    }
FASTAD_B200_BINARY_OP(operator+, Add)
FASTAD_B200_BINARY_OP(operator-, Sub)
FASTAD_B200_BINARY_OP(operator*, Mul)
FASTAD_B200_BINARY_OP(operator/, Div)
FASTAD_B200_BINARY_OP(operator<, LessThan)
FASTAD_B200_BINARY_OP(operator<=, LessThanEq)
FASTAD_B200_BINARY_OP(operator>, GreaterThan)
FASTAD_B200_BINARY_OP(operator>=, GreaterThanEq)
FASTAD_B200_BINARY_OP(operator==, Equal)
FASTAD_B200_BINARY_OP(operator!=, NotEqual)
FASTAD_B200_BINARY_OP(operator&&, LogicalAnd)
FASTAD_B200_BINARY_OP(operator||, LogicalOr)
#undef FASTAD_B200_BINARY_OP

template <class A, class B, class = std::enable_if_t<is_expr_v<A> && is_expr_v<B>>>
inline auto operator,(const A& a, const B& b)
{
    using ea_t = util::convert_to_ad_t<A>;
    using eb_t = util::convert_to_ad_t<B>;
    ea_t ea = a;
    eb_t eb = b;
    return GlueNode<ea_t, eb_t>(ea, eb);
}

#define FASTAD_B200_UNARY_FN(name, tag)                                                               \
    template <class A, class = std::enable_if_t<util::is_convertible_to_ad_v<A> && util::any_ad_v<A>>> \
    inline auto name(const A& a)                                                                      \
    {                                                                                                 \
        using ea_t = util::convert_to_ad_t<A>;                                                        \
        ea_t ea = a;                                                                                  \
        if constexpr (is_scl_constant_v<ea_t>) return Constant<double, scl>(fold_unary(tag::code, ea.feval())); \
        else return UnaryNode<tag, ea_t>(ea);                                                         \
    }
FASTAD_B200_UNARY_FN(operator-, UnaryMinus)
}  // namespace core

#define FASTAD_B200_UNARY_FN2(name, tag)                                                              \
    template <class A, class = std::enable_if_t<util::is_convertible_to_ad_v<A> && util::any_ad_v<A>>> \
    inline auto name(const A& a)                                                                      \
    {                                                                                                 \
        using ea_t = util::convert_to_ad_t<A>;                                                        \
        ea_t ea = a;                                                                                  \
        if constexpr (core::is_scl_constant_v<ea_t>)                                                  \
            return core::Constant<double, scl>(core::fold_unary(core::tag::code, ea.feval()));        \
        else                                                                                          \
            return core::UnaryNode<core::tag, ea_t>(ea);                                              \
    }
FASTAD_B200_UNARY_FN2(sin, Sin) FASTAD_B200_UNARY_FN2(cos, Cos) FASTAD_B200_UNARY_FN2(tan, Tan)
FASTAD_B200_UNARY_FN2(asin, Arcsin) FASTAD_B200_UNARY_FN2(acos, Arccos) FASTAD_B200_UNARY_FN2(atan, Arctan)
FASTAD_B200_UNARY_FN2(exp, Exp) FASTAD_B200_UNARY_FN2(log, Log) FASTAD_B200_UNARY_FN2(sqrt, Sqrt)
FASTAD_B200_UNARY_FN2(erf, Erf) FASTAD_B200_UNARY_FN2(sigmoid, Sigmoid) FASTAD_B200_UNARY_FN2(sinh, Sinh)
FASTAD_B200_UNARY_FN2(cosh, Cosh) FASTAD_B200_UNARY_FN2(tanh, Tanh)
#undef FASTAD_B200_UNARY_FN2
#undef FASTAD_B200_UNARY_FN

// ---- builders ----------------------------------------------------------------------------------
inline auto constant(double x) { return core::Constant<double, scl>(x); }
inline auto constant(const std::vector<double>& x) { return core::Constant<double, vec>(x.data(), x.size(), 1); }
// column-major host data of a rows x cols matrix
inline auto constant(const std::vector<double>& x, size_t rows, size_t cols)
{
    assert(x.size() == rows * cols);
    return core::Constant<double, mat>(x.data(), rows, cols);
}
// DEVICE pointers (constant.hpp:158-170)
inline auto constant_view(const double* dev, size_t rows) { return core::ConstantView<double, vec>(dev, rows, 1); }
template <class ShapeType = mat>
inline auto constant_view(const double* dev, size_t rows, size_t cols) { return core::ConstantView<double, ShapeType>(dev, rows, cols); }

namespace core {
inline double fold_pow(int64_t n, double b)
{   // PowFunc, pow.hpp:20-56
    if (n < 0) return b == 0 ? std::numeric_limits<double>::infinity() : fold_pow(-n, 1. / b);
    if (n == 0) return 1.;
    if (n % 2 == 0) { double t = fold_pow(n / 2, b); return t * t; }
    return b * fold_pow(n - 1, b);
}
}  // namespace core

template <int64_t N, class A, class = std::enable_if_t<util::is_convertible_to_ad_v<A> && util::any_ad_v<A>>>
inline auto pow(const A& a)
{
    using ea_t = util::convert_to_ad_t<A>;
    ea_t ea = a;
    if constexpr (core::is_scl_constant_v<ea_t>) return core::Constant<double, scl>(core::fold_pow(N, ea.feval()));
    else return core::PowNode<N, ea_t>(ea);
}

template <class A, class = std::enable_if_t<util::is_convertible_to_ad_v<A> && util::any_ad_v<A>>>
inline auto sum(const A& a)
{
    using ea_t = util::convert_to_ad_t<A>;
    ea_t ea = a;
    if constexpr (core::is_scl_constant_v<ea_t>) return ea;      // sum.hpp:250-252
    else return core::SumElemNode<ea_t>(ea);
}
template <class A, class = std::enable_if_t<util::is_convertible_to_ad_v<A> && util::any_ad_v<A>>>
inline auto prod(const A& a)
{
    using ea_t = util::convert_to_ad_t<A>;
    ea_t ea = a;
    if constexpr (core::is_scl_constant_v<ea_t>) return ea;      // prod.hpp:292-294
    else return core::ProdElemNode<ea_t>(ea);
}
// ad::sum(begin, end, f) / ad::prod(begin, end, f)  (sum.hpp:216-241, prod.hpp:259-285)
template <class Iter, class Lmda>
inline auto sum(Iter begin, Iter end, Lmda&& f)
{
    using expr_t = util::convert_to_ad_t<std::decay_t<decltype(f(*begin))>>;
    if constexpr (core::is_scl_constant_v<expr_t>) {
        double acc = 0;
        for (auto it = begin; it != end; ++it) acc += expr_t(f(*it)).feval();
        return core::Constant<double, scl>(acc);
    } else {
        std::vector<expr_t> exprs;
        exprs.reserve(std::distance(begin, end));
        for (auto it = begin; it != end; ++it) exprs.emplace_back(f(*it));
        return core::SumIterNode<expr_t>(std::move(exprs));
    }
}
template <class Iter, class Lmda>
inline auto prod(Iter begin, Iter end, Lmda&& f)
{
    using expr_t = util::convert_to_ad_t<std::decay_t<decltype(f(*begin))>>;
    if constexpr (core::is_scl_constant_v<expr_t>) {
        double acc = 1;
        for (auto it = begin; it != end; ++it) acc *= expr_t(f(*it)).feval();
        return core::Constant<double, scl>(acc);
    } else {
        std::vector<expr_t> exprs;
        exprs.reserve(std::distance(begin, end));
        for (auto it = begin; it != end; ++it) exprs.emplace_back(f(*it));
        return core::ProdIterNode<expr_t>(std::move(exprs));
    }
}

template <class A, class B, class = std::enable_if_t<util::is_convertible_to_ad_v<A> && util::is_convertible_to_ad_v<B>>>
inline auto dot(const A& a, const B& b)
{
    using ea_t = util::convert_to_ad_t<A>;
    using eb_t = util::convert_to_ad_t<B>;
    ea_t ea = a;
    eb_t eb = b;
    return core::DotNode<ea_t, eb_t>(ea, eb);
}

template <class X, class M, class S,
          class = std::enable_if_t<util::is_convertible_to_ad_v<X> && util::is_convertible_to_ad_v<M> &&
                                   util::is_convertible_to_ad_v<S> && util::any_ad_v<X, M, S>>>
inline auto normal_adj_log_pdf(const X& x, const M& m, const S& s)
{
    using x_t = util::convert_to_ad_t<X>;
    using m_t = util::convert_to_ad_t<M>;
    using s_t = util::convert_to_ad_t<S>;
    x_t ex = x;
    m_t em = m;
    s_t es = s;
    return core::NormalAdjLogPDFNode<x_t, m_t, s_t>(ex, em, es);
}

// ---- ad::cauchy_adj_log_pdf / uniform_adj_log_pdf / bernoulli_adj_log_pdf ------
#define FASTAD_B200_LOGPDF3(name, dist)                                                               \
    template <class X, class P1, class P2,                                                            \
              class = std::enable_if_t<util::is_convertible_to_ad_v<X> && util::is_convertible_to_ad_v<P1> && \
                                       util::is_convertible_to_ad_v<P2> && util::any_ad_v<X, P1, P2>>> \
    inline auto name(const X& x, const P1& p1, const P2& p2)                                          \
    {                                                                                                 \
        using x_t = util::convert_to_ad_t<X>;                                                         \
        using a_t = util::convert_to_ad_t<P1>;                                                        \
        using b_t = util::convert_to_ad_t<P2>;                                                        \
        x_t ex = x;                                                                                   \
        a_t ea = p1;                                                                                  \
        b_t eb = p2;                                                                                  \
        return core::LogPDFNode<dist, x_t, a_t, b_t>(ex, ea, eb);                                     \
    }
FASTAD_B200_LOGPDF3(cauchy_adj_log_pdf, FADB_CAUCHY)     // cauchy.hpp: (x, loc, scale)
FASTAD_B200_LOGPDF3(uniform_adj_log_pdf, FADB_UNIFORM)   // uniform.hpp: (x, min, max)
#undef FASTAD_B200_LOGPDF3
template <class X, class P, class = std::enable_if_t<util::is_convertible_to_ad_v<X> && util::is_convertible_to_ad_v<P> &&
                                                     util::any_ad_v<X, P>>>
inline auto bernoulli_adj_log_pdf(const X& x, const P& p)   // bernoulli.hpp: (x, p)
{
    using x_t = util::convert_to_ad_t<X>;
    using p_t = util::convert_to_ad_t<P>;
    x_t ex = x;
    p_t ep = p;
    return core::LogPDFNode<FADB_BERNOULLI, x_t, p_t, core::Constant<double, scl>>(ex, ep, core::Constant<double, scl>(0.0));
}

// ---- ad::det / ad::log_det / ad::wishart_adj_log_pdf ------------------------------
// decomposition policies (det.hpp:97-202, log_det.hpp:97-202); here they only select the device path
template <class ValueType> struct DetFullPivLU { using value_t = ValueType; static constexpr int code = FADB_DET_FULLPIVLU; };
template <class ValueType> struct DetLDLT { using value_t = ValueType; static constexpr int code = FADB_DET_LDLT; };
template <class ValueType> struct DetLLT { using value_t = ValueType; static constexpr int code = FADB_DET_LLT; };
template <class ValueType> struct LogDetFullPivLU { using value_t = ValueType; static constexpr int code = FADB_DET_FULLPIVLU; };
template <class ValueType> struct LogDetLDLT { using value_t = ValueType; static constexpr int code = FADB_DET_LDLT; };
template <class ValueType> struct LogDetLLT { using value_t = ValueType; static constexpr int code = FADB_DET_LLT; };

// det.hpp:211-231.  A constant matrix operand is NOT folded at build time here (the reference
// returns a Constant): its data lives in HBM, so the node is evaluated like any other.
template <template <class> class DecompType = DetFullPivLU, class T,
          class = std::enable_if_t<util::is_convertible_to_ad_v<T> && util::any_ad_v<T>>>
inline auto det(const T& x)
{
    using expr_t = util::convert_to_ad_t<T>;
    expr_t expr = x;
    return core::DetNode<DecompType<double>, expr_t>(expr);
}
template <template <class> class DecompType = LogDetFullPivLU, class T,
          class = std::enable_if_t<util::is_convertible_to_ad_v<T> && util::any_ad_v<T>>>
inline auto log_det(const T& x)       // log_det.hpp:211-231
{
    using expr_t = util::convert_to_ad_t<T>;
    expr_t expr = x;
    return core::LogDetNode<DecompType<double>, expr_t>(expr);
}

namespace stat {
// WishartAdjLogPDFNode (wishart.hpp:92-231): x, v square matrices, n a scalar constant
template <class X, class V, class N>
struct WishartAdjLogPDFNode : core::ExprBase<WishartAdjLogPDFNode<X, V, N>> {
    static_assert(util::is_mat_v<X> && util::is_mat_v<V> && util::is_scl_v<N>, "wishart: mat, mat, scl (wishart.hpp:25-27)");
    static_assert(core::is_scl_constant_v<N>, "wishart: n must be a constant (wishart.hpp:28)");
    using value_t = double;
    using shape_t = scl;
    WishartAdjLogPDFNode(const X& x, const V& v, const N& n) : x_(x), v_(v), n_(n) {}
    size_t rows() const { return 1; }
    size_t cols() const { return 1; }
    size_t size() const { return 1; }
    int lower(detail::Lowering& L) const
    {
        int a = x_.lower(L);
        int b = v_.lower(L);
        return detail::checked_id(fadb_expr_wishart(L.e, a, b, (double)n_.feval()));
    }
private:
    X x_;
    V v_;
    N n_;
};
}  // namespace stat

template <class XType, class VType, class NType,
          class = std::enable_if_t<util::is_convertible_to_ad_v<XType> && util::is_convertible_to_ad_v<VType> &&
                                   util::is_convertible_to_ad_v<NType> && util::any_ad_v<XType, VType, NType>>>
inline auto wishart_adj_log_pdf(const XType& x, const VType& v, const NType& n)      // wishart.hpp:242-257
{
    using x_t = util::convert_to_ad_t<XType>;
    using v_t = util::convert_to_ad_t<VType>;
    using n_t = util::convert_to_ad_t<NType>;
    x_t ex = x;
    v_t ev = v;
    n_t en = n;
    return stat::WishartAdjLogPDFNode<x_t, v_t, n_t>(ex, ev, en);
}

// ---- SURVEY.md 8f "next" rows ---------------------------------------------------------------
template <class A, class = std::enable_if_t<util::is_convertible_to_ad_v<A> && util::any_ad_v<A>>>
inline auto norm(const A& a)          // norm.hpp:91-106
{
    using ea_t = util::convert_to_ad_t<A>;
    ea_t ea = a;
    return core::NormNode<ea_t>(ea);
}
template <class A, class = std::enable_if_t<util::is_convertible_to_ad_v<A> && util::any_ad_v<A>>>
inline auto transpose(const A& a)     // transpose.hpp:64-79
{
    using ea_t = util::convert_to_ad_t<A>;
    ea_t ea = a;
    return core::TransposeNode<ea_t>(ea);
}
template <class C, class I, class E,
          class = std::enable_if_t<util::is_convertible_to_ad_v<C> && util::is_convertible_to_ad_v<I> &&
                                   util::is_convertible_to_ad_v<E> && util::any_ad_v<C, I, E>>>
inline auto if_else(const C& c, const I& i, const E& e)   // if_else.hpp:118-163
{
    using c_t = util::convert_to_ad_t<C>;
    using i_t = util::convert_to_ad_t<I>;
    using e_t = util::convert_to_ad_t<E>;
    c_t ec = c;
    i_t ei = i;
    e_t ee = e;
    return core::IfElseNode<c_t, i_t, e_t>(ec, ei, ee);
}
// ad::for_each(begin, end, f): evaluate every f(x) in order, value of the last (for_each.hpp:19-131);
// lowered as a left-to-right chain of glue nodes
namespace core {
template <class ExprType>
struct ForEachIterNode : ExprBase<ForEachIterNode<ExprType>> {
    using value_t = double;
    using shape_t = typename ExprType::shape_t;
    ForEachIterNode(std::vector<ExprType> v) : vec_(std::move(v)) { assert(!vec_.empty()); }
    size_t rows() const { return vec_.back().rows(); }
    size_t cols() const { return vec_.back().cols(); }
    size_t size() const { return vec_.back().size(); }
    int lower(detail::Lowering& L) const
    {
        int acc = vec_[0].lower(L);
        for (size_t k = 1; k < vec_.size(); ++k) acc = detail::checked_id(fadb_expr_glue(L.e, acc, vec_[k].lower(L)));
        return acc;
    }
private:
    std::vector<ExprType> vec_;
};
}  // namespace core
template <class Iter, class Lmda>
inline auto for_each(Iter begin, Iter end, Lmda&& f)
{
    using expr_t = util::convert_to_ad_t<std::decay_t<decltype(f(*begin))>>;
    std::vector<expr_t> exprs;
    exprs.reserve(std::distance(begin, end));
    for (auto it = begin; it != end; ++it) exprs.emplace_back(f(*it));
    return core::ForEachIterNode<expr_t>(std::move(exprs));
}

// w += expr, w -= expr, w *= expr, w /= expr  (var_view.hpp:199-217)
#define FASTAD_B200_OPEQ(name, tag)                                                                   \
    template <class V, class S, class D, class = std::enable_if_t<util::is_convertible_to_ad_v<D>>>   \
    inline auto name(const VarView<V, S>& var, const D& x)                                            \
    {                                                                                                 \
        using expr_t = util::convert_to_ad_t<D>;                                                      \
        expr_t expr = x;                                                                              \
        return core::OpEqNode<core::tag, VarView<V, S>, expr_t>(var, expr);                           \
    }
FASTAD_B200_OPEQ(operator+=, AddEq)
FASTAD_B200_OPEQ(operator-=, SubEq)
FASTAD_B200_OPEQ(operator*=, MulEq)
FASTAD_B200_OPEQ(operator/=, DivEq)
#undef FASTAD_B200_OPEQ

// ---------------------------------------------------------------------------------------------
// bind / evaluate / evaluate_adj / autodiff  (bind.hpp:18-49, eval.hpp:18-153)
// ---------------------------------------------------------------------------------------------
struct device_seed {
    const double* ptr;   // device array of the root's size
};

namespace core {

template <class ExprType>
struct ExprBind {
    using expr_t = ExprType;
    using value_t = double;
    using shape_t = typename ExprType::shape_t;

    explicit ExprBind(const expr_t& expr, unsigned flags = 0) : expr_(expr)
    {
        fadb_expr* h = nullptr;
        detail::check(fadb_expr_create(&h));
        L_.e = h;
        handle_.reset(h, [](fadb_expr* p) { fadb_expr_destroy(p); });
        root_ = expr_.lower(L_);
        size_t cs[2] = {0, 0};
        detail::check(fadb_expr_bind(h, root_, flags, cs));
        cache_ = {{cs[0], cs[1]}};
        size_ = expr_.size();
    }
    expr_t& get() { return expr_; }
    size_t size() const { return size_; }
    util::SizePack cache_size() const { return cache_; }
    fadb_expr* handle() const { return handle_.get(); }

    void before()
    {
        for (auto& s : L_.storages) s->to_device();
    }
    void after(bool vals, bool adjs)
    {
        for (auto& s : L_.storages) {
            if (vals) s->dev_val_newer = true;
            if (adjs && s->dadj) s->dev_adj_newer = true;
        }
    }
    const double* stage_seed(const std::vector<double>& seed)
    {
        assert(seed.size() == size_);
        if (!seed_buf_) {
            void* p = nullptr;
            detail::check(fadb_malloc(size_ * sizeof(double), &p));
            seed_buf_.reset(static_cast<double*>(p), [](double* q) { fadb_free(q); });
        }
        detail::check(fadb_h2d(seed_buf_.get(), seed.data(), size_ * sizeof(double)));
        return seed_buf_.get();
    }
    // device pointer of the root value after an evaluation (no host copy)
    const double* value_device() const
    {
        const double* p = nullptr;
        detail::check(fadb_expr_value_ptr(handle_.get(), &p, nullptr));
        return p;
    }

private:
    expr_t expr_;
    detail::Lowering L_;
    std::shared_ptr<fadb_expr> handle_;
    std::shared_ptr<double> seed_buf_;
    int root_ = -1;
    size_t size_ = 1;
    util::SizePack cache_{{0, 0}};
};

template <class T> struct is_bound : std::false_type {};
template <class E> struct is_bound<ExprBind<E>> : std::true_type {};
template <class T> inline constexpr bool is_bound_v = is_bound<std::decay_t<T>>::value;

template <class B>
inline auto fetch(B&, const std::vector<double>& v)
{
    if constexpr (util::is_scl_v<B>) return v[0];
    else return v;
}

}  // namespace core

template <class Derived>
inline auto bind(const core::ExprBase<Derived>& expr, unsigned flags = 0)
{
    return core::ExprBind<Derived>(expr.self(), flags);
}

// ---- on bound expressions
template <class E>
inline auto evaluate(core::ExprBind<E>& b)
{
    b.before();
    std::vector<double> v(b.size());
    detail::check(fadb_expr_feval(b.handle(), v.data()));
    b.after(true, false);
    return core::fetch(b, v);
}
template <class E> inline auto evaluate(core::ExprBind<E>&& b) { return evaluate(b); }

template <class E, class = std::enable_if_t<util::is_scl_v<E>>>
inline void evaluate_adj(core::ExprBind<E>& b, double seed = 1.)
{
    b.before();
    detail::check(fadb_expr_beval(b.handle(), seed, nullptr));
    b.after(true, true);      // values too: the sweep of `w op= expr` puts the previous value back into w (eq.hpp:262-271)
}
template <class E>
inline void evaluate_adj(core::ExprBind<E>& b, const std::vector<double>& seed)
{
    b.before();
    detail::check(fadb_expr_beval(b.handle(), 0., b.stage_seed(seed)));
    b.after(true, true);      // values too: the sweep of `w op= expr` puts the previous value back into w (eq.hpp:262-271)
}
template <class E>
inline void evaluate_adj(core::ExprBind<E>& b, device_seed seed)
{
    b.before();
    detail::check(fadb_expr_beval(b.handle(), 0., seed.ptr));
    b.after(true, true);      // values too: the sweep of `w op= expr` puts the previous value back into w (eq.hpp:262-271)
}

template <class E, class = std::enable_if_t<util::is_scl_v<E>>>
inline double autodiff(core::ExprBind<E>& b, double seed = 1.)
{
    b.before();
    double v = 0;
    detail::check(fadb_expr_autodiff(b.handle(), seed, nullptr, &v));
    b.after(true, true);
    return v;
}
template <class E, class = std::enable_if_t<util::is_scl_v<E>>>
inline double autodiff(core::ExprBind<E>&& b, double seed = 1.) { return autodiff(b, seed); }

template <class E>
inline auto autodiff(core::ExprBind<E>& b, const std::vector<double>& seed)
{
    b.before();
    std::vector<double> v(b.size());
    detail::check(fadb_expr_autodiff(b.handle(), 0., b.stage_seed(seed), v.data()));
    b.after(true, true);
    return core::fetch(b, v);
}
// device-resident seed, no host copy of the value (read it through value_device())
template <class E>
inline void autodiff(core::ExprBind<E>& b, device_seed seed)
{
    b.before();
    detail::check(fadb_expr_autodiff(b.handle(), 0., seed.ptr, nullptr));
    b.after(true, true);
}

// ---- on raw expressions: the reference requires a manual bind_cache first (README.md:284-306);
//      here a temporary binding is made (costs a planning pass per call: bind once in loops).
template <class E, class = std::enable_if_t<core::is_expr_v<E> && !core::is_bound_v<E>>>
inline auto evaluate(const E& expr)
{
    auto b = bind(expr);
    return evaluate(b);
}
template <class E, class = std::enable_if_t<core::is_expr_v<E> && !core::is_bound_v<E> && util::is_scl_v<E>>>
inline double autodiff(const E& expr, double seed = 1.)
{
    auto b = bind(expr);
    return autodiff(b, seed);
}
template <class E, class = std::enable_if_t<core::is_expr_v<E> && !core::is_bound_v<E>>>
inline auto autodiff(const E& expr, const std::vector<double>& seed)
{
    auto b = bind(expr);
    return autodiff(b, seed);
}

// bind_cache_size() of the reference (e.g. unary.hpp:107-116): only stage boundaries are
// materialised here, so this is what the planner reports.  Host-only (no CUDA call).
template <class Derived>
inline util::SizePack bind_cache_size(const core::ExprBase<Derived>& expr)
{
    fadb_expr* h = nullptr;
    detail::check(fadb_expr_create(&h));
    std::shared_ptr<fadb_expr> guard(h, [](fadb_expr* p) { fadb_expr_destroy(p); });
    detail::Lowering L;
    L.e = h;
    L.device = false;
    int root = expr.self().lower(L);
    size_t cs[2] = {0, 0};
    detail::check(fadb_expr_plan(h, root, cs));
    return {{cs[0], cs[1]}};
}

namespace core {
template <class D>
inline ExprBind<D>& own_binding(const ExprBase<D>& e)
{
    if (!e.own_binding_) {
        auto b = std::make_shared<ExprBind<D>>(e.self());      // the copy inside holds no binding yet: no cycle
        e.own_binding_ = b;
    }
    return *static_cast<ExprBind<D>*>(e.own_binding_.get());
}
}  // namespace core
template <class Derived>
inline auto core::ExprBase<Derived>::feval() const { return ad::evaluate(core::own_binding(*this)); }
template <class Derived>
inline void core::ExprBase<Derived>::beval(double seed) const { ad::evaluate_adj(core::own_binding(*this), seed); }
template <class Derived>
inline void core::ExprBase<Derived>::beval(const std::vector<double>& seed) const { ad::evaluate_adj(core::own_binding(*this), seed); }

template <class Derived>
inline util::SizePack core::ExprBase<Derived>::bind_cache_size() const { return ad::bind_cache_size(*this); }
template <class Derived>
inline util::PtrPack<double> core::ExprBase<Derived>::bind_cache(util::PtrPack<double> begin) const
{
    const util::SizePack sz = ad::bind_cache_size(*this);
    return {begin.val ? begin.val + sz(0) : nullptr, begin.adj ? begin.adj + sz(1) : nullptr};
}

}  // namespace ad
