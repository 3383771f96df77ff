/*
 * nb_vec2.h -- the reference's 2-vector types for hosts that switch to this framework but keep
 * their own body arithmetic: `Vec2f` (include/vec2f.h:13-98, host + device) and `Vec2<T>`
 * (include/vec2.h:6-96, host only) as two aliases of ONE template.
 *
 * Same public surface: members X, Y and Element[2]; an uninitialised default constructor, a
 * broadcasting scalar constructor (so `force = 0.f` zeroes both lanes, src/nbody.cu:153), element-wise
 * `*`, `+`, `-`, unary `-`, scalar `*` on both sides, the compound forms, and length().
 * Same arithmetic where it is observable: `v / s` multiplies by the reciprocal `1.0f / s`, rounded once in T
 * (for Vec2<double> the float literal is promoted and the division is a double division, exactly as in the
 * reference, vec2.h:49) -- reciprocal-then-multiply, not a division per component -- and length() is
 * sqrt(X*X + Y*Y) evaluated in T.  Dividing by zero asserts (vec2f.h:51) unless NDEBUG is set.
 *
 * The SoA device store of the library (float4 {x,y,m,r}, float2 v) does not use this type; the
 * BodiesData block that nb_upload / nb_download exchange is layout-compatible with arrays of it
 * (sizeof(Vec2f) == 8, no padding).
 */
#ifndef NB_VEC2_H
#define NB_VEC2_H

#include <assert.h>
#include <math.h>

#ifdef __CUDACC__
#define NB_HD __host__ __device__
#else
#define NB_HD
#endif

namespace nbvec {

template <typename T>
struct Vec2T {
    union {
        T Element[2];
        struct {
            T X, Y;
        };
    };

    NB_HD Vec2T() {}                                             /* deliberately uninitialised */
    NB_HD Vec2T(T both) : X(both), Y(both) {}
    NB_HD Vec2T(T x, T y) : X(x), Y(y) {}
    NB_HD Vec2T(const Vec2T &o) : X(o.X), Y(o.Y) {}
    NB_HD Vec2T &operator=(const Vec2T &o) { X = o.X; Y = o.Y; return *this; }

    NB_HD T operator[](int k) const { return Element[k]; }
    NB_HD T &operator[](int k) { return Element[k]; }

    NB_HD Vec2T operator*(T s) const { return Vec2T(s * X, s * Y); }
    NB_HD Vec2T operator*(const Vec2T &o) const { return Vec2T(o.X * X, o.Y * Y); }
    NB_HD Vec2T operator+(const Vec2T &o) const { return Vec2T(X + o.X, Y + o.Y); }
    NB_HD Vec2T operator-(const Vec2T &o) const { return Vec2T(X - o.X, Y - o.Y); }
    NB_HD Vec2T operator-() const { return Vec2T(-X, -Y); }
    NB_HD Vec2T operator/(T s) const
    {
        assert(s != 0.0);
        return *this * (T)(1.0f / s);                            /* reciprocal first, evaluated in T */
    }

    NB_HD Vec2T &operator*=(T s) { return *this = *this * s; }
    NB_HD Vec2T &operator*=(const Vec2T &o) { return *this = *this * o; }
    NB_HD Vec2T &operator/=(T s) { return *this = *this / s; }
    NB_HD Vec2T &operator+=(const Vec2T &o) { return *this = *this + o; }
    NB_HD Vec2T &operator-=(const Vec2T &o) { return *this = *this - o; }

    NB_HD T length() const { return (T)sqrt(X * X + Y * Y); }
};

template <typename T>
NB_HD inline Vec2T<T> operator*(T s, const Vec2T<T> &v) { return Vec2T<T>(s * v.X, s * v.Y); }

}  // namespace nbvec

typedef nbvec::Vec2T<float> Vec2f;
template <typename T>
using Vec2 = nbvec::Vec2T<T>;

#endif /* NB_VEC2_H */
