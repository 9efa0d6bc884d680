// ublas_min.hpp -- a small stand-in for the part of Boost uBLAS that the reference's LDPC
// blocks use (lib/ldpc_decoder_cb_impl.{h,cc}, lib/ldpc_encoder_bc_impl.{h,cc}).
//
// TEST INFRASTRUCTURE ONLY (oracle/): it exists so that the reference's own, unmodified block
// sources can be compiled where they lie under /root/reference into oracle/_ref/ and be used
// as the checker for oracle/ldpc_oracle.c and for the CUDA path.  Boost is not in this image.
// Written from the uBLAS interface the call sites need, not from Boost's sources:
//
//   vector<T>, matrix<T> (dense, row-major), zero_matrix<T>
//   row(m, i), column(m, j), subrange(m, r0, r1, c0, c1)          (read/write proxies)
//   inner_prod, element_prod, prod(matrix expr, vector expr), unary -, binary +
//
// uBLAS builds expression templates; here every operator evaluates eagerly into a concrete
// vector.  For the reference's call sites (element-wise sums/products over int and double,
// each evaluated once into a named vector) the values are identical: the element expressions
// and the ascending index order of inner_prod / prod are the same.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <limits>
#include <type_traits>
#include <utility>
#include <vector>

namespace boost { namespace numeric { namespace ublas {

template <class E> struct vector_expression {
    const E &self() const { return static_cast<const E &>(*this); }
    E &self() { return static_cast<E &>(*this); }
};
template <class E> struct matrix_expression {
    const E &self() const { return static_cast<const E &>(*this); }
    E &self() { return static_cast<E &>(*this); }
};

template <class T> class vector : public vector_expression<vector<T> >
{
    std::vector<T> d_;

public:
    typedef T value_type;
    typedef std::size_t size_type;

    vector() {}
    explicit vector(size_type n) : d_(n) {}
    vector(size_type n, const T &v) : d_(n, v) {}
    vector(const vector &o) : vector_expression<vector<T> >(), d_(o.d_) {}
    template <class E> vector(const vector_expression<E> &e) { assign(e.self()); }

    vector &operator=(const vector &o) { d_ = o.d_; return *this; }
    template <class E> vector &operator=(const vector_expression<E> &e)
    {
        assign(e.self());
        return *this;
    }

    size_type size() const { return d_.size(); }
    void resize(size_type n, bool = true) { d_.resize(n); }
    T &operator()(size_type i) { return d_[i]; }
    const T &operator()(size_type i) const { return d_[i]; }
    T &operator[](size_type i) { return d_[i]; }
    const T &operator[](size_type i) const { return d_[i]; }

private:
    template <class E> void assign(const E &e)
    {
        std::vector<T> t(e.size());                 // via a temporary: e may alias *this
        for (size_type i = 0; i < t.size(); i++) t[i] = e(i);
        d_.swap(t);
    }
};

template <class T> class matrix : public matrix_expression<matrix<T> >
{
    std::size_t r_, c_;
    std::vector<T> d_;

public:
    typedef T value_type;
    typedef std::size_t size_type;

    matrix() : r_(0), c_(0) {}
    matrix(size_type r, size_type c) : r_(r), c_(c), d_(r * c) {}
    matrix(const matrix &o) : matrix_expression<matrix<T> >(), r_(o.r_), c_(o.c_), d_(o.d_) {}
    template <class E> matrix(const matrix_expression<E> &e) : r_(0), c_(0) { assign(e.self()); }

    matrix &operator=(const matrix &o)
    {
        r_ = o.r_; c_ = o.c_; d_ = o.d_;
        return *this;
    }
    template <class E> matrix &operator=(const matrix_expression<E> &e)
    {
        assign(e.self());
        return *this;
    }

    size_type size1() const { return r_; }
    size_type size2() const { return c_; }
    T &operator()(size_type i, size_type j) { return d_[i * c_ + j]; }
    const T &operator()(size_type i, size_type j) const { return d_[i * c_ + j]; }

private:
    template <class E> void assign(const E &e)
    {
        const size_type r = e.size1(), c = e.size2();
        std::vector<T> t(r * c);
        for (size_type i = 0; i < r; i++)
            for (size_type j = 0; j < c; j++) t[i * c + j] = e(i, j);
        r_ = r; c_ = c;
        d_.swap(t);
    }
};

template <class T> class zero_matrix : public matrix_expression<zero_matrix<T> >
{
    std::size_t r_, c_;

public:
    typedef T value_type;
    typedef std::size_t size_type;
    zero_matrix(size_type r, size_type c) : r_(r), c_(c) {}
    size_type size1() const { return r_; }
    size_type size2() const { return c_; }
    T operator()(size_type, size_type) const { return T(); }
};

// ---- proxies -------------------------------------------------------------------------------
template <class M> struct proxy_ref {
    typedef typename M::value_type value_type;
    typedef typename std::conditional<std::is_const<M>::value, const value_type &,
                                      value_type &>::type type;
};

template <class M> class matrix_row : public vector_expression<matrix_row<M> >
{
    M &m_;
    std::size_t i_;

public:
    typedef typename M::value_type value_type;
    typedef std::size_t size_type;
    matrix_row(M &m, size_type i) : m_(m), i_(i) {}
    matrix_row(const matrix_row &o) : vector_expression<matrix_row<M> >(), m_(o.m_), i_(o.i_) {}
    size_type size() const { return m_.size2(); }
    typename proxy_ref<M>::type operator()(size_type j) const { return m_(i_, j); }

    matrix_row &operator=(const matrix_row &o) { return assign(o); }
    template <class E> matrix_row &operator=(const vector_expression<E> &e)
    {
        return assign(e.self());
    }

private:
    template <class E> matrix_row &assign(const E &e)
    {
        const vector<value_type> t(e);              // snapshot: e may read this row
        assert(t.size() == size());
        for (size_type j = 0; j < t.size(); j++) m_(i_, j) = t(j);
        return *this;
    }
};

template <class M> class matrix_column : public vector_expression<matrix_column<M> >
{
    M &m_;
    std::size_t j_;

public:
    typedef typename M::value_type value_type;
    typedef std::size_t size_type;
    matrix_column(M &m, size_type j) : m_(m), j_(j) {}
    matrix_column(const matrix_column &o)
        : vector_expression<matrix_column<M> >(), m_(o.m_), j_(o.j_) {}
    size_type size() const { return m_.size1(); }
    typename proxy_ref<M>::type operator()(size_type i) const { return m_(i, j_); }

    matrix_column &operator=(const matrix_column &o) { return assign(o); }
    template <class E> matrix_column &operator=(const vector_expression<E> &e)
    {
        return assign(e.self());
    }

private:
    template <class E> matrix_column &assign(const E &e)
    {
        const vector<value_type> t(e);
        assert(t.size() == size());
        for (size_type i = 0; i < t.size(); i++) m_(i, j_) = t(i);
        return *this;
    }
};

template <class M> class matrix_range : public matrix_expression<matrix_range<M> >
{
    M &m_;
    std::size_t r0_, r1_, c0_, c1_;

public:
    typedef typename M::value_type value_type;
    typedef std::size_t size_type;
    matrix_range(M &m, size_type r0, size_type r1, size_type c0, size_type c1)
        : m_(m), r0_(r0), r1_(r1), c0_(c0), c1_(c1) {}
    matrix_range(const matrix_range &o)
        : matrix_expression<matrix_range<M> >(), m_(o.m_), r0_(o.r0_), r1_(o.r1_), c0_(o.c0_),
          c1_(o.c1_) {}
    size_type size1() const { return r1_ - r0_; }
    size_type size2() const { return c1_ - c0_; }
    typename proxy_ref<M>::type operator()(size_type i, size_type j) const
    {
        return m_(r0_ + i, c0_ + j);
    }

    matrix_range &operator=(const matrix_range &o) { return assign(o); }
    template <class E> matrix_range &operator=(const matrix_expression<E> &e)
    {
        return assign(e.self());
    }

private:
    template <class E> matrix_range &assign(const E &e)
    {
        const matrix<value_type> t(e);
        assert(t.size1() == size1() && t.size2() == size2());
        for (size_type i = 0; i < t.size1(); i++)
            for (size_type j = 0; j < t.size2(); j++) m_(r0_ + i, c0_ + j) = t(i, j);
        return *this;
    }
};

template <class M> matrix_row<M> row(M &m, std::size_t i) { return matrix_row<M>(m, i); }
template <class M> matrix_column<M> column(M &m, std::size_t j) { return matrix_column<M>(m, j); }
template <class M>
matrix_range<M> subrange(M &m, std::size_t r0, std::size_t r1, std::size_t c0, std::size_t c1)
{
    return matrix_range<M>(m, r0, r1, c0, c1);
}

// ---- operations (eager) -----------------------------------------------------------------------
template <class A, class B> struct promote {
    typedef decltype(std::declval<A>() * std::declval<B>()) type;
};

template <class E1, class E2>
typename promote<typename E1::value_type, typename E2::value_type>::type
inner_prod(const vector_expression<E1> &a, const vector_expression<E2> &b)
{
    typedef typename promote<typename E1::value_type, typename E2::value_type>::type R;
    const E1 &x = a.self();
    const E2 &y = b.self();
    assert(x.size() == y.size());
    R s = R();
    for (std::size_t i = 0; i < x.size(); i++) s += x(i) * y(i);
    return s;
}

template <class E1, class E2>
vector<typename promote<typename E1::value_type, typename E2::value_type>::type>
element_prod(const vector_expression<E1> &a, const vector_expression<E2> &b)
{
    typedef typename promote<typename E1::value_type, typename E2::value_type>::type R;
    const E1 &x = a.self();
    const E2 &y = b.self();
    assert(x.size() == y.size());
    vector<R> v(x.size());
    for (std::size_t i = 0; i < x.size(); i++) v(i) = x(i) * y(i);
    return v;
}

template <class E1, class E2>
vector<typename promote<typename E1::value_type, typename E2::value_type>::type>
prod(const matrix_expression<E1> &a, const vector_expression<E2> &b)
{
    typedef typename promote<typename E1::value_type, typename E2::value_type>::type R;
    const E1 &m = a.self();
    const E2 &x = b.self();
    assert(m.size2() == x.size());
    vector<R> v(m.size1());
    for (std::size_t i = 0; i < m.size1(); i++) {
        R s = R();
        for (std::size_t j = 0; j < m.size2(); j++) s += m(i, j) * x(j);
        v(i) = s;
    }
    return v;
}

template <class E1, class E2>
vector<typename promote<typename E1::value_type, typename E2::value_type>::type>
operator+(const vector_expression<E1> &a, const vector_expression<E2> &b)
{
    typedef typename promote<typename E1::value_type, typename E2::value_type>::type R;
    const E1 &x = a.self();
    const E2 &y = b.self();
    assert(x.size() == y.size());
    vector<R> v(x.size());
    for (std::size_t i = 0; i < x.size(); i++) v(i) = x(i) + y(i);
    return v;
}

template <class E>
vector<typename E::value_type> operator-(const vector_expression<E> &a)
{
    const E &x = a.self();
    vector<typename E::value_type> v(x.size());
    for (std::size_t i = 0; i < x.size(); i++) v(i) = -x(i);
    return v;
}

}}}  // namespace boost::numeric::ublas
