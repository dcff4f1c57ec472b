// TEST INFRASTRUCTURE (oracle/): a minimal, eagerly evaluated stand-in for the part of Eigen's dense API that the
// reference's own sources use on the hot path (lib/PLS/src/pls.cpp, lib/PLS/include/PLS/pls.h, the path functions of
// src/AbcUtil.cpp). Eigen itself (un-vendored submodule lib/PLS/lib/eigen @ 23e1541) is absent from this image, so the
// reference cannot be built as shipped; with this header on the include path (as <Eigen/Core>, <Eigen/Dense>,
// <Eigen/Eigenvalues>) the reference's UNMODIFIED source files compile where they lie under /root/reference into
// oracle/_ref/ (oracle/Makefile, target ref). What runs is then the reference's own statements — loop structure, quirks,
// operation order at the statement level — on top of naive loops for Eigen's kernels:
//   * products, reductions: plain sequential loops (Eigen's are blocked / vectorised: same values to rounding);
//   * EigenSolver<MatrixXd>: only for symmetric input (what pls.cpp:406 feeds it, XY^T XY): Householder tridiagonalisation
//     + implicit QL (the classic tred2 / tql2 pair) — a different method from the cyclic Jacobi of oracle/abc_oracle.cpp on
//     purpose, so the two are independent checks of each other; eigenvectors unit-norm, sign arbitrary as in Eigen;
//   * vector <-> transposed-vector assignment is accepted (Eigen transposes vectors implicitly on assignment).
// Nothing in the product (abcsmc_b200/, include/) includes this file.
#pragma once
// Real Eigen pulls <stdlib.h> in through the SSE intrinsics headers (<emmintrin.h> -> <mm_malloc.h>), and libstdc++'s
// <stdlib.h> / <math.h> wrappers put the floating-point overloads of abs() into the global namespace: pls.cpp:157 calls an
// unqualified abs(z) on a double and depends on that (with <cmath> alone it would be the int overload).
#include <stdlib.h>
#include <math.h>
#include <cmath>
#include <complex>
#include <vector>
#include <cassert>
#include <cstddef>
#include <climits>   // real Eigen/Core includes it (AbcUtil.cpp:225 names INT_MIN)
#include <cstdio>
#include <ostream>
#include <algorithm>
#include <type_traits>

namespace Eigen {

typedef std::ptrdiff_t Index;
const int Dynamic = -1;
namespace placeholders { struct all_t {}; static const all_t all = all_t(); }

template <class T, int R, int C> class Matrix;
template <class T, int R, int C> class Block;
template <class M> class ArrayWrap;

template <class S> struct is_scalar : std::is_arithmetic<S> {};
template <class S> struct is_scalar<std::complex<S>> : std::true_type {};
template <class S> struct real_of { typedef S type; };
template <class S> struct real_of<std::complex<S>> { typedef S type; };

namespace internal {
template <class S> inline typename real_of<S>::type real_part(const S& x) { return x; }
template <class S> inline S real_part(const std::complex<S>& x) { return x.real(); }
template <class S> inline typename real_of<S>::type abs2(const S& x) { return x * x; }
template <class S> inline S abs2(const std::complex<S>& x) { return std::norm(x); }
}

// ------------------------------------------------------------------------------------------------------------------
// Base: everything that only READS a matrix-shaped thing. D provides eval() -> (const ref to | value of) Matrix.
// ------------------------------------------------------------------------------------------------------------------
template <class D> struct traits;
template <class T, int R, int C> struct traits<Matrix<T, R, C>> { typedef T Scalar; enum { Rows = R, Cols = C }; };
template <class T, int R, int C> struct traits<Block<T, R, C>> { typedef T Scalar; enum { Rows = R, Cols = C }; };

template <class M> class ColwiseOp;
template <class M> class RowwiseOp;

template <class D>
struct Base {
    typedef typename traits<D>::Scalar Scalar;
    enum { Rows = traits<D>::Rows, Cols = traits<D>::Cols };
    typedef Matrix<Scalar, Rows, Cols> Plain;
    typedef typename real_of<Scalar>::type RealScalar;

    const D& derived() const { return *static_cast<const D*>(this); }
    Plain plain() const { return Plain(derived().eval()); }

    Matrix<Scalar, Cols, Rows> transpose() const {
        const auto& a = derived().eval();
        Matrix<Scalar, Cols, Rows> t; t.resize(a.cols(), a.rows());
        for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) t(j, i) = a(i, j);
        return t;
    }
    Plain cwiseAbs() const {
        Plain r = plain();
        for (Index k = 0; k < r.size(); k++) r.data()[k] = std::abs(r.data()[k]);
        return r;
    }
    Plain cwiseSqrt() const {
        Plain r = plain();
        for (Index k = 0; k < r.size(); k++) r.data()[k] = std::sqrt(r.data()[k]);
        return r;
    }
    template <class E> Plain cwiseProduct(const Base<E>& o) const {
        Plain r = plain(); const auto& b = o.derived().eval();
        assert(r.rows() == b.rows() && r.cols() == b.cols());
        for (Index j = 0; j < r.cols(); j++) for (Index i = 0; i < r.rows(); i++) r(i, j) *= b(i, j);
        return r;
    }
    Matrix<RealScalar, Rows, Cols> real() const {
        const auto& a = derived().eval();
        Matrix<RealScalar, Rows, Cols> r; r.resize(a.rows(), a.cols());
        for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) r(i, j) = internal::real_part(a(i, j));
        return r;
    }
    template <class U> Matrix<U, Rows, Cols> cast() const {
        const auto& a = derived().eval();
        Matrix<U, Rows, Cols> r; r.resize(a.rows(), a.cols());
        for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) r(i, j) = U(a(i, j));
        return r;
    }
    ArrayWrap<Plain> array() const { return ArrayWrap<Plain>(plain()); }
    ColwiseOp<Plain> colwise() const { return ColwiseOp<Plain>(plain()); }
    RowwiseOp<Plain> rowwise() const { return RowwiseOp<Plain>(plain()); }

    Scalar sum() const {
        const auto& a = derived().eval(); Scalar s = Scalar(0);
        for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) s += a(i, j);
        return s;
    }
    RealScalar squaredNorm() const {
        const auto& a = derived().eval(); RealScalar s = 0;
        for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) s += internal::abs2(a(i, j));
        return s;
    }
    RealScalar norm() const { return std::sqrt(squaredNorm()); }
    Scalar mean() const { const auto& a = derived().eval(); return sum() / Scalar(a.size()); }

    // first extremum wins, like Eigen's visitors (strict comparison while scanning in storage order)
    template <class I> Scalar minCoeff(I* idx) const {
        const auto& a = derived().eval(); assert(a.size() > 0 && (a.rows() == 1 || a.cols() == 1));
        Index best = 0; Scalar m = a(0);
        for (Index k = 1; k < a.size(); k++) if (a(k) < m) { m = a(k); best = k; }
        *idx = static_cast<I>(best); return m;
    }
    Scalar minCoeff() const {
        const auto& a = derived().eval(); assert(a.size() > 0); Scalar m = a.data()[0];
        for (Index k = 1; k < a.size(); k++) if (a.data()[k] < m) m = a.data()[k];
        return m;
    }
    Scalar maxCoeff() const {
        const auto& a = derived().eval(); assert(a.size() > 0); Scalar m = a.data()[0];
        for (Index k = 1; k < a.size(); k++) if (a.data()[k] > m) m = a.data()[k];
        return m;
    }

    // row gather: M(indices, placeholders::all)
    template <class I> Matrix<Scalar, Dynamic, Cols> operator()(const std::vector<I>& rows_, placeholders::all_t) const {
        const auto& a = derived().eval();
        Matrix<Scalar, Dynamic, Cols> r; r.resize(static_cast<Index>(rows_.size()), a.cols());
        for (Index j = 0; j < a.cols(); j++)
            for (size_t i = 0; i < rows_.size(); i++) r(static_cast<Index>(i), j) = a(static_cast<Index>(rows_[i]), j);
        return r;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// Matrix: column-major owning storage; R, C in {Dynamic, 1}
// ------------------------------------------------------------------------------------------------------------------
template <class T, int R, int C>
class Matrix : public Base<Matrix<T, R, C>> {
    static_assert((R == Dynamic || R == 1) && (C == Dynamic || C == 1), "mini_eigen: only dynamic matrices and vectors");
    Index r_, c_;
    std::vector<T> d_;
  public:
    typedef T Scalar;
    typedef T value_type;
    Matrix() : r_(R == 1 ? 1 : 0), c_(C == 1 ? 1 : 0), d_(static_cast<size_t>((R == 1 && C == 1) ? 1 : 0)) {}
    // vector of n entries (vectors only) / r x c matrix
    template <class I, class = typename std::enable_if<std::is_integral<I>::value>::type>
    explicit Matrix(I n) { static_assert(R == 1 || C == 1, "size constructor is for vectors"); if (R == 1) resize(1, Index(n)); else resize(Index(n), 1); }
    template <class I, class J, class = typename std::enable_if<std::is_integral<I>::value && std::is_integral<J>::value>::type>
    Matrix(I r, J c) { resize(Index(r), Index(c)); }
    Matrix(const Matrix&) = default;
    Matrix(Matrix&&) = default;
    Matrix& operator=(const Matrix&) = default;
    Matrix& operator=(Matrix&&) = default;

    // from any matrix-shaped thing of the same scalar type; a vector may be assigned from its transposed kind
    template <class E, class = typename std::enable_if<std::is_same<typename traits<E>::Scalar, T>::value>::type>
    Matrix(const Base<E>& o) { assign(o.derived().eval()); }
    template <class M> Matrix(const ArrayWrap<M>& a) { assign(a.matrix()); }
    template <class E, class = typename std::enable_if<std::is_same<typename traits<E>::Scalar, T>::value>::type>
    Matrix& operator=(const Base<E>& o) { Matrix<T, traits<E>::Rows, traits<E>::Cols> tmp(o.derived().eval()); assign(tmp); return *this; }
    template <class M> Matrix& operator=(const ArrayWrap<M>& a) { assign(a.matrix()); return *this; }

    template <int R2, int C2> void assign(const Matrix<T, R2, C2>& o) {
        if ((R == 1 && o.cols() == 1 && o.rows() != 1) || (C == 1 && o.rows() == 1 && o.cols() != 1)) {
            resize(o.cols(), o.rows());                      // implicit transposition of a vector
        } else {
            assert((R != 1 || o.rows() == 1) && (C != 1 || o.cols() == 1));
            resize(o.rows(), o.cols());
        }
        std::copy(o.data(), o.data() + o.size(), d_.begin());
    }

    const Matrix& eval() const { return *this; }
    Matrix& noalias() { return *this; }

    void resize(Index r, Index c) { assert(r >= 0 && c >= 0); r_ = r; c_ = c; d_.assign(static_cast<size_t>(r * c), T(0)); }
    void resize(Index n) { static_assert(R == 1 || C == 1, "vectors"); if (R == 1) resize(1, n); else resize(n, 1); }
    void setZero(Index r, Index c) { resize(r, c); }
    void setZero() { std::fill(d_.begin(), d_.end(), T(0)); }
    static Matrix Zero(Index r, Index c) { Matrix m; m.resize(r, c); return m; }
    static Matrix Zero(Index n) { static_assert(R == 1 || C == 1, "vectors"); return Matrix(n); }
    static Matrix Constant(Index r, Index c, const T& v) { Matrix m; m.resize(r, c); std::fill(m.d_.begin(), m.d_.end(), v); return m; }
    static Matrix Constant(Index n, const T& v) { Matrix m(n); std::fill(m.d_.begin(), m.d_.end(), v); return m; }

    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    Index outerStride() const { return r_; }
    T* data() { return d_.data(); }
    const T* data() const { return d_.data(); }
    T& operator()(Index i, Index j) { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return d_[static_cast<size_t>(i + j * r_)]; }
    const T& operator()(Index i, Index j) const { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return d_[static_cast<size_t>(i + j * r_)]; }
    T& operator()(Index k) { assert(k >= 0 && k < size()); return d_[static_cast<size_t>(k)]; }
    const T& operator()(Index k) const { assert(k >= 0 && k < size()); return d_[static_cast<size_t>(k)]; }
    T& operator[](Index k) { return (*this)(k); }
    const T& operator[](Index k) const { return (*this)(k); }
    typename std::vector<T>::iterator begin() { return d_.begin(); }
    typename std::vector<T>::iterator end() { return d_.end(); }
    typename std::vector<T>::const_iterator begin() const { return d_.begin(); }
    typename std::vector<T>::const_iterator end() const { return d_.end(); }

    // blocks (views into this storage)
    Block<T, 1, C> row(Index i) { return Block<T, 1, C>(data() + i, r_, 1, c_); }
    Block<T, 1, C> row(Index i) const { return Block<T, 1, C>(const_cast<T*>(data()) + i, r_, 1, c_); }
    Block<T, R, 1> col(Index j) { return Block<T, R, 1>(data() + j * r_, r_, r_, 1); }
    Block<T, R, 1> col(Index j) const { return Block<T, R, 1>(const_cast<T*>(data()) + j * r_, r_, r_, 1); }
    Block<T, R, C> block_(Index i0, Index j0, Index nr, Index nc) const {
        assert(i0 >= 0 && j0 >= 0 && nr >= 0 && nc >= 0 && i0 + nr <= r_ && j0 + nc <= c_);
        return Block<T, R, C>(const_cast<T*>(data()) + i0 + j0 * r_, r_, nr, nc);
    }
    Block<T, R, C> topRows(Index n) const { return block_(0, 0, n, c_); }
    Block<T, R, C> bottomRows(Index n) const { return block_(r_ - n, 0, n, c_); }
    Block<T, R, C> middleRows(Index i0, Index n) const { return block_(i0, 0, n, c_); }
    Block<T, R, C> leftCols(Index n) const { return block_(0, 0, r_, n); }
    Block<T, R, C> rightCols(Index n) const { return block_(0, c_ - n, r_, n); }
    Block<T, R, C> head(Index n) const { return R == 1 ? block_(0, 0, 1, n) : block_(0, 0, n, 1); }
    using Base<Matrix<T, R, C>>::operator();                 // the row gather
    // column scatter: M(placeholders::all, indices) = other
    template <class I> struct ColScatter {
        Matrix& m; const std::vector<I>& cols_;
        template <class E> void operator=(const Base<E>& o) const {
            const auto& x = o.derived().eval(); assert(x.rows() == m.rows() && x.cols() == static_cast<Index>(cols_.size()));
            for (size_t j = 0; j < cols_.size(); j++) for (Index i = 0; i < m.rows(); i++) m(i, static_cast<Index>(cols_[j])) = x(i, static_cast<Index>(j));
        }
    };
    template <class I> ColScatter<I> operator()(placeholders::all_t, const std::vector<I>& cols_) { return ColScatter<I>{*this, cols_}; }

    template <class E> Matrix& operator+=(const Base<E>& o) { row_or_all() += o; return *this; }
    template <class E> Matrix& operator-=(const Base<E>& o) { row_or_all() -= o; return *this; }
    template <class S, class = typename std::enable_if<is_scalar<S>::value>::type>
    Matrix& operator/=(const S& s) { for (auto& x : d_) x /= s; return *this; }
    template <class S, class = typename std::enable_if<is_scalar<S>::value>::type>
    Matrix& operator*=(const S& s) { for (auto& x : d_) x *= s; return *this; }

    // Eigen (Core/Dot.h): RealScalar z = squaredNorm(); if (z > RealScalar(0)) derived() /= numext::sqrt(z);
    // so a zero vector AND a vector holding a NaN (z > 0 is false) are both left as they are
    void normalize() { const auto z = this->squaredNorm(); if (z > decltype(z)(0)) { const auto n = std::sqrt(z); for (auto& x : d_) x /= n; } }
    Matrix normalized() const { Matrix m(*this); m.normalize(); return m; }
  private:
    Block<T, R, C> row_or_all() { return Block<T, R, C>(data(), r_, r_, c_); }
};

// ------------------------------------------------------------------------------------------------------------------
// Block: a writable view (pointer, leading dimension, extent). Reads go through eval() -> Matrix.
// ------------------------------------------------------------------------------------------------------------------
template <class T, int R, int C>
class Block : public Base<Block<T, R, C>> {
    T* p_; Index ld_, r_, c_;
  public:
    typedef T Scalar;
    Block(T* p, Index ld, Index r, Index c) : p_(p), ld_(ld), r_(r), c_(c) {}
    Block(const Block&) = default;
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    T& operator()(Index i, Index j) const { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return p_[i + j * ld_]; }
    T& operator()(Index k) const { assert(r_ == 1 || c_ == 1); return r_ == 1 ? (*this)(0, k) : (*this)(k, 0); }
    T& operator[](Index k) const { return (*this)(k); }
    Matrix<T, R, C> eval() const {
        Matrix<T, R, C> m; m.resize(r_, c_);
        for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) m(i, j) = p_[i + j * ld_];
        return m;
    }

    struct iterator {
        T* p; Index step;
        T& operator*() const { return *p; }
        iterator& operator++() { p += step; return *this; }
        bool operator!=(const iterator& o) const { return p != o.p; }
    };
    iterator begin() const { assert(r_ == 1 || c_ == 1); return iterator{p_, c_ == 1 ? 1 : ld_}; }
    iterator end() const { const Index step = (c_ == 1 ? 1 : ld_); return iterator{p_ + step * size(), step}; }
    Block<T, 1, C> row(Index i) const { assert(i >= 0 && i < r_); return Block<T, 1, C>(p_ + i, ld_, 1, c_); }
    Block<T, R, 1> col(Index j) const { assert(j >= 0 && j < c_); return Block<T, R, 1>(p_ + j * ld_, ld_, r_, 1); }
    Block middleRows(Index i0, Index n) const { assert(i0 >= 0 && i0 + n <= r_); return Block(p_ + i0, ld_, n, c_); }
    Block topRows(Index n) const { return middleRows(0, n); }
    Block bottomRows(Index n) const { return middleRows(r_ - n, n); }
    Block leftCols(Index n) const { assert(n <= c_); return Block(p_, ld_, r_, n); }

    // writes: element-wise when the shapes agree, linear when a vector meets its transposed kind
    template <class M> void store(const M& m, int op) const {
        const bool same = (m.rows() == r_ && m.cols() == c_);
        const bool flipped = !same && (m.rows() == c_ && m.cols() == r_) && (r_ == 1 || c_ == 1);
        assert(same || flipped); (void)flipped;
        for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) {
            const T v = same ? m(i, j) : m(j, i);
            T& x = p_[i + j * ld_];
            if (op == 0) x = v; else if (op == 1) x += v; else x -= v;
        }
    }
    template <class E> const Block& operator=(const Base<E>& o) const { const auto m = o.plain(); store(m, 0); return *this; }
    const Block& operator=(const Block& o) const { const auto m = o.eval(); store(m, 0); return *this; }
    template <class E> const Block& operator+=(const Base<E>& o) const { const auto m = o.plain(); store(m, 1); return *this; }
    template <class E> const Block& operator-=(const Base<E>& o) const { const auto m = o.plain(); store(m, 2); return *this; }
};

// ------------------------------------------------------------------------------------------------------------------
// arithmetic
// ------------------------------------------------------------------------------------------------------------------
template <class A, class B> typename Base<A>::Plain operator+(const Base<A>& a, const Base<B>& b) {
    typename Base<A>::Plain r = a.plain(); const auto& y = b.derived().eval();
    if (y.rows() == r.rows() && y.cols() == r.cols()) { for (Index k = 0; k < r.size(); k++) r.data()[k] += y.data()[k]; }
    else { assert(y.size() == r.size() && (r.rows() == 1 || r.cols() == 1)); for (Index k = 0; k < r.size(); k++) r.data()[k] += y.data()[k]; }
    return r;
}
template <class A, class B> typename Base<A>::Plain operator-(const Base<A>& a, const Base<B>& b) {
    typename Base<A>::Plain r = a.plain(); const auto& y = b.derived().eval();
    assert(y.size() == r.size() && ((y.rows() == r.rows()) || r.rows() == 1 || r.cols() == 1));
    for (Index k = 0; k < r.size(); k++) r.data()[k] -= y.data()[k];
    return r;
}
template <class A> typename Base<A>::Plain operator-(const Base<A>& a) {
    typename Base<A>::Plain r = a.plain(); for (Index k = 0; k < r.size(); k++) r.data()[k] = -r.data()[k]; return r;
}
// matrix product: sequential inner products, k ascending
template <class A, class B>
Matrix<decltype(typename Base<A>::Scalar() * typename Base<B>::Scalar()), Base<A>::Rows, Base<B>::Cols>
operator*(const Base<A>& a, const Base<B>& b) {
    typedef decltype(typename Base<A>::Scalar() * typename Base<B>::Scalar()) S;
    const auto& x = a.derived().eval(); const auto& y = b.derived().eval();
    assert(x.cols() == y.rows());
    Matrix<S, Base<A>::Rows, Base<B>::Cols> r; r.resize(x.rows(), y.cols());
    for (Index j = 0; j < y.cols(); j++)
        for (Index k = 0; k < x.cols(); k++) {
            const auto ykj = y(k, j);
            for (Index i = 0; i < x.rows(); i++) r(i, j) += x(i, k) * ykj;
        }
    return r;
}
// scalar on one side: the result type is only formed when S really is a scalar (keeps these out of matrix x matrix)
template <class A, class S, bool = is_scalar<S>::value> struct scaled {};
template <class A, class S> struct scaled<A, S, true> {
    typedef Matrix<decltype(typename traits<A>::Scalar() * S()), traits<A>::Rows, traits<A>::Cols> type;
};
template <class A, class S> typename scaled<A, S>::type operator*(const Base<A>& a, const S& s) {
    const auto& x = a.derived().eval();
    typename scaled<A, S>::type r; r.resize(x.rows(), x.cols());
    for (Index j = 0; j < x.cols(); j++) for (Index i = 0; i < x.rows(); i++) r(i, j) = x(i, j) * s;
    return r;
}
template <class A, class S> typename scaled<A, S>::type operator*(const S& s, const Base<A>& a) {
    const auto& x = a.derived().eval();
    typename scaled<A, S>::type r; r.resize(x.rows(), x.cols());
    for (Index j = 0; j < x.cols(); j++) for (Index i = 0; i < x.rows(); i++) r(i, j) = s * x(i, j);
    return r;
}
template <class A, class S> typename scaled<A, S>::type operator/(const Base<A>& a, const S& s) {
    const auto& x = a.derived().eval();
    typename scaled<A, S>::type r; r.resize(x.rows(), x.cols());
    for (Index j = 0; j < x.cols(); j++) for (Index i = 0; i < x.rows(); i++) r(i, j) = x(i, j) / s;
    return r;
}
// comma initialiser (`m << 1, 2, 3, ...;`): fills row by row, as Eigen does (the reference's own tests/abcutil.cpp uses it)
template <class T, int R, int C>
struct CommaInit {
    Matrix<T, R, C>& m; Index k;
    template <class S> CommaInit& operator,(const S& v) { assert(k < m.size()); m(k / m.cols(), k % m.cols()) = T(v); k++; return *this; }
};
template <class T, int R, int C, class S, class = typename std::enable_if<is_scalar<S>::value>::type>
CommaInit<T, R, C> operator<<(Matrix<T, R, C>& m, const S& v) { assert(m.size() > 0); m(0, 0) = T(v); return CommaInit<T, R, C>{m, 1}; }
template <class D> std::ostream& operator<<(std::ostream& os, const Base<D>& b) {
    const auto& a = b.derived().eval();
    for (Index i = 0; i < a.rows(); i++) {
        for (Index j = 0; j < a.cols(); j++) { if (j) os << ' '; os << a(i, j); }
        if (i + 1 < a.rows()) os << '\n';
    }
    return os;
}

// ------------------------------------------------------------------------------------------------------------------
// partial reductions and broadcasting
// ------------------------------------------------------------------------------------------------------------------
template <class M>
class ColwiseOp {
    M m_;
    typedef typename M::Scalar T;
    typedef Matrix<T, 1, Dynamic> RowT;
    typedef Matrix<typename real_of<T>::type, 1, Dynamic> RealRowT;
  public:
    explicit ColwiseOp(const M& m) : m_(m) {}
    RowT sum() const { RowT r(m_.cols()); for (Index j = 0; j < m_.cols(); j++) { T s = T(0); for (Index i = 0; i < m_.rows(); i++) s += m_(i, j); r(j) = s; } return r; }
    RowT mean() const { RowT r = sum(); for (Index j = 0; j < m_.cols(); j++) r(j) /= T(m_.rows()); return r; }
    RealRowT squaredNorm() const { RealRowT r(m_.cols()); for (Index j = 0; j < m_.cols(); j++) { typename real_of<T>::type s = 0; for (Index i = 0; i < m_.rows(); i++) s += internal::abs2(m_(i, j)); r(j) = s; } return r; }
    RealRowT norm() const { RealRowT r = squaredNorm(); for (Index j = 0; j < r.size(); j++) r(j) = std::sqrt(r(j)); return r; }
};
template <class M>
class RowwiseOp {
    M m_;
    typedef typename M::Scalar T;
    typedef Matrix<T, Dynamic, 1> ColT;
    typedef Matrix<typename real_of<T>::type, Dynamic, 1> RealColT;
  public:
    explicit RowwiseOp(const M& m) : m_(m) {}
    template <class E> M operator-(const Base<E>& v) const { const auto& x = v.derived().eval(); assert(x.size() == m_.cols()); M r = m_; for (Index j = 0; j < r.cols(); j++) for (Index i = 0; i < r.rows(); i++) r(i, j) -= x(j); return r; }
    template <class E> M operator+(const Base<E>& v) const { const auto& x = v.derived().eval(); assert(x.size() == m_.cols()); M r = m_; for (Index j = 0; j < r.cols(); j++) for (Index i = 0; i < r.rows(); i++) r(i, j) += x(j); return r; }
    ColT sum() const { ColT r(m_.rows()); for (Index i = 0; i < m_.rows(); i++) { T s = T(0); for (Index j = 0; j < m_.cols(); j++) s += m_(i, j); r(i) = s; } return r; }
    RealColT squaredNorm() const { RealColT r(m_.rows()); for (Index i = 0; i < m_.rows(); i++) { typename real_of<T>::type s = 0; for (Index j = 0; j < m_.cols(); j++) s += internal::abs2(m_(i, j)); r(i) = s; } return r; }
    RealColT norm() const { RealColT r = squaredNorm(); for (Index i = 0; i < r.size(); i++) r(i) = std::sqrt(r(i)); return r; }
};

template <class M> class ArrayRowwise;
template <class M>
class ArrayWrap {
    M m_;
    typedef typename M::Scalar T;
  public:
    typedef T Scalar;
    explicit ArrayWrap(const M& m) : m_(m) {}
    const M& matrix() const { return m_; }
    Index rows() const { return m_.rows(); }
    Index cols() const { return m_.cols(); }
    Index size() const { return m_.size(); }
    ArrayWrap square() const { M r = m_; for (Index k = 0; k < r.size(); k++) r.data()[k] = r.data()[k] * r.data()[k]; return ArrayWrap(r); }
    ArrayWrap sqrt() const { M r = m_; for (Index k = 0; k < r.size(); k++) r.data()[k] = std::sqrt(r.data()[k]); return ArrayWrap(r); }
    ArrayWrap abs() const { M r = m_; for (Index k = 0; k < r.size(); k++) r.data()[k] = std::abs(r.data()[k]); return ArrayWrap(r); }
    ArrayWrap log() const { M r = m_; for (Index k = 0; k < r.size(); k++) r.data()[k] = std::log(r.data()[k]); return ArrayWrap(r); }
    template <class S> ArrayWrap pow(const S& e) const { M r = m_; for (Index k = 0; k < r.size(); k++) r.data()[k] = std::pow(r.data()[k], e); return ArrayWrap(r); }
    T sum() const { return m_.sum(); }
    T mean() const { return m_.mean(); }
    ColwiseOp<M> colwise() const { return ColwiseOp<M>(m_); }
    ArrayRowwise<M> rowwise() const { return ArrayRowwise<M>(m_); }
    template <class M2> ArrayWrap operator/(const ArrayWrap<M2>& o) const { assert(o.size() == size()); M r = m_; for (Index k = 0; k < r.size(); k++) r.data()[k] /= o.matrix().data()[k]; return ArrayWrap(r); }
    template <class M2> ArrayWrap operator*(const ArrayWrap<M2>& o) const { assert(o.size() == size()); M r = m_; for (Index k = 0; k < r.size(); k++) r.data()[k] *= o.matrix().data()[k]; return ArrayWrap(r); }
    template <class M2> ArrayWrap operator-(const ArrayWrap<M2>& o) const { assert(o.size() == size()); M r = m_; for (Index k = 0; k < r.size(); k++) r.data()[k] -= o.matrix().data()[k]; return ArrayWrap(r); }
    template <class M2> ArrayWrap operator+(const ArrayWrap<M2>& o) const { assert(o.size() == size()); M r = m_; for (Index k = 0; k < r.size(); k++) r.data()[k] += o.matrix().data()[k]; return ArrayWrap(r); }
};
template <class S, class M, class = typename std::enable_if<is_scalar<S>::value>::type>
ArrayWrap<M> operator-(const S& s, const ArrayWrap<M>& a) { M r = a.matrix(); for (Index k = 0; k < r.size(); k++) r.data()[k] = s - r.data()[k]; return ArrayWrap<M>(r); }
template <class S, class M, class = typename std::enable_if<is_scalar<S>::value>::type>
ArrayWrap<M> operator-(const ArrayWrap<M>& a, const S& s) { M r = a.matrix(); for (Index k = 0; k < r.size(); k++) r.data()[k] -= s; return ArrayWrap<M>(r); }
template <class S, class M, class = typename std::enable_if<is_scalar<S>::value>::type>
ArrayWrap<M> operator+(const ArrayWrap<M>& a, const S& s) { M r = a.matrix(); for (Index k = 0; k < r.size(); k++) r.data()[k] += s; return ArrayWrap<M>(r); }
template <class S, class M, class = typename std::enable_if<is_scalar<S>::value>::type>
ArrayWrap<M> operator/(const ArrayWrap<M>& a, const S& s) { M r = a.matrix(); for (Index k = 0; k < r.size(); k++) r.data()[k] /= s; return ArrayWrap<M>(r); }
template <class M>
class ArrayRowwise {
    M m_;
  public:
    explicit ArrayRowwise(const M& m) : m_(m) {}
    template <class M2> ArrayWrap<M> operator/(const ArrayWrap<M2>& v) const {
        assert(v.size() == m_.cols()); M r = m_;
        for (Index j = 0; j < r.cols(); j++) for (Index i = 0; i < r.rows(); i++) r(i, j) /= v.matrix().data()[j];
        return ArrayWrap<M>(r);
    }
};

typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, 1, Dynamic> RowVectorXd;
typedef Matrix<std::complex<double>, Dynamic, Dynamic> MatrixXcd;
typedef Matrix<std::complex<double>, Dynamic, 1> VectorXcd;
typedef Matrix<int, Dynamic, 1> VectorXi;
typedef Matrix<int, 1, Dynamic> RowVectorXi;

// ------------------------------------------------------------------------------------------------------------------
// EigenSolver stand-in: symmetric real input only (see the header comment). tred2 + tql2.
// ------------------------------------------------------------------------------------------------------------------
template <class MatrixType>
class EigenSolver {
    VectorXcd vals_;
    MatrixXcd vecs_;
  public:
    template <class E> explicit EigenSolver(const Base<E>& m) { compute(MatrixXd(m.derived().eval())); }
    const VectorXcd& eigenvalues() const { return vals_; }
    const MatrixXcd& eigenvectors() const { return vecs_; }
  private:
    void compute(const MatrixXd& a) {
        const Index n = a.rows(); assert(a.cols() == n);
        double amax = 0; for (Index k = 0; k < a.size(); k++) amax = std::max(amax, std::fabs(a.data()[k]));
        for (Index i = 0; i < n; i++) for (Index j = 0; j < i; j++)
            if (std::fabs(a(i, j) - a(j, i)) > 1e-12 * amax) { std::fprintf(stderr, "mini_eigen: EigenSolver stand-in needs a symmetric matrix\n"); std::abort(); }
        std::vector<std::vector<double>> V(n, std::vector<double>(n));
        for (Index i = 0; i < n; i++) for (Index j = 0; j < n; j++) V[i][j] = 0.5 * (a(i, j) + a(j, i));
        std::vector<double> d(n), e(n);
        tridiagonalise(V, d, e); ql_implicit(V, d, e);
        vals_ = VectorXcd(n); vecs_ = MatrixXcd(n, n);
        for (Index j = 0; j < n; j++) {
            vals_(j) = std::complex<double>(d[j], 0.0);
            double nrm = 0; for (Index i = 0; i < n; i++) nrm += V[i][j] * V[i][j];
            nrm = std::sqrt(nrm);
            for (Index i = 0; i < n; i++) vecs_(i, j) = std::complex<double>(V[i][j] / nrm, 0.0);
        }
    }
    // Householder reduction of a symmetric matrix to tridiagonal form, accumulating the transformation in V
    static void tridiagonalise(std::vector<std::vector<double>>& V, std::vector<double>& d, std::vector<double>& e) {
        const Index n = static_cast<Index>(d.size());
        for (Index j = 0; j < n; j++) d[j] = V[n - 1][j];
        for (Index i = n - 1; i > 0; i--) {
            double scale = 0, h = 0;
            for (Index k = 0; k < i; k++) scale += std::fabs(d[k]);
            if (scale == 0.0) {
                e[i] = d[i - 1];
                for (Index j = 0; j < i; j++) { d[j] = V[i - 1][j]; V[i][j] = 0; V[j][i] = 0; }
            } else {
                for (Index k = 0; k < i; k++) { d[k] /= scale; h += d[k] * d[k]; }
                double f = d[i - 1], g = std::sqrt(h);
                if (f > 0) g = -g;
                e[i] = scale * g; h -= f * g; d[i - 1] = f - g;
                for (Index j = 0; j < i; j++) e[j] = 0;
                for (Index j = 0; j < i; j++) {
                    f = d[j]; V[j][i] = f; g = e[j] + V[j][j] * f;
                    for (Index k = j + 1; k <= i - 1; k++) { g += V[k][j] * d[k]; e[k] += V[k][j] * f; }
                    e[j] = g;
                }
                f = 0;
                for (Index j = 0; j < i; j++) { e[j] /= h; f += e[j] * d[j]; }
                const double hh = f / (h + h);
                for (Index j = 0; j < i; j++) e[j] -= hh * d[j];
                for (Index j = 0; j < i; j++) {
                    f = d[j]; g = e[j];
                    for (Index k = j; k <= i - 1; k++) V[k][j] -= (f * e[k] + g * d[k]);
                    d[j] = V[i - 1][j]; V[i][j] = 0;
                }
            }
            d[i] = h;
        }
        for (Index i = 0; i < n - 1; i++) {
            V[n - 1][i] = V[i][i]; V[i][i] = 1.0;
            const double h = d[i + 1];
            if (h != 0.0) {
                for (Index k = 0; k <= i; k++) d[k] = V[k][i + 1] / h;
                for (Index j = 0; j <= i; j++) {
                    double g = 0;
                    for (Index k = 0; k <= i; k++) g += V[k][i + 1] * V[k][j];
                    for (Index k = 0; k <= i; k++) V[k][j] -= g * d[k];
                }
            }
            for (Index k = 0; k <= i; k++) V[k][i + 1] = 0;
        }
        for (Index j = 0; j < n; j++) { d[j] = V[n - 1][j]; V[n - 1][j] = 0; }
        V[n - 1][n - 1] = 1.0; e[0] = 0;
    }
    // implicit QL iteration on the tridiagonal (d, e), rotating the columns of V
    static void ql_implicit(std::vector<std::vector<double>>& V, std::vector<double>& d, std::vector<double>& e) {
        const Index n = static_cast<Index>(d.size());
        for (Index i = 1; i < n; i++) e[i - 1] = e[i];
        e[n - 1] = 0;
        double f = 0, tst1 = 0; const double eps = std::pow(2.0, -52.0);
        for (Index l = 0; l < n; l++) {
            tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
            Index m = l;
            while (m < n) { if (std::fabs(e[m]) <= eps * tst1) break; m++; }
            if (m > l) {
                int iter = 0;
                do {
                    if (++iter > 200) { std::fprintf(stderr, "mini_eigen: QL did not converge\n"); std::abort(); }
                    double g = d[l], p = (d[l + 1] - g) / (2.0 * e[l]), r = std::hypot(p, 1.0);
                    if (p < 0) r = -r;
                    d[l] = e[l] / (p + r); d[l + 1] = e[l] * (p + r);
                    const double dl1 = d[l + 1]; double h = g - d[l];
                    for (Index i = l + 2; i < n; i++) d[i] -= h;
                    f += h;
                    p = d[m]; double c = 1, c2 = c, c3 = c; const double el1 = e[l + 1]; double s = 0, s2 = 0;
                    for (Index i = m - 1; i >= l; i--) {
                        c3 = c2; c2 = c; s2 = s;
                        g = c * e[i]; h = c * p; r = std::hypot(p, e[i]);
                        e[i + 1] = s * r; s = e[i] / r; c = p / r; p = c * d[i] - s * g;
                        d[i + 1] = h + s * (c * g + s * d[i]);
                        for (Index k = 0; k < n; k++) { h = V[k][i + 1]; V[k][i + 1] = s * V[k][i] + c * h; V[k][i] = c * V[k][i] - s * h; }
                    }
                    p = -s * s2 * c3 * el1 * e[l] / dl1; e[l] = s * p; d[l] = c * p;
                } while (std::fabs(e[l]) > eps * tst1);
            }
            d[l] += f; e[l] = 0;
        }
    }
};

}  // namespace Eigen
