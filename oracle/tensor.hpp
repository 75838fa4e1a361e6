// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the value/tensor semantics of the reference
// (power_grid_model/common/three_phase_tensor.hpp:30-143, 200-350, 380-392 and common/common.hpp:24-107).
// The reference uses Eigen fixed-size arrays; Eigen is not available here, so the semantics are restated
// with plain arrays.  B = 1 is the symmetric (positive-sequence) calculation, B = 3 the asymmetric (abc) one.
#pragma once

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <limits>
#include <numbers>
#include <stdexcept>
#include <string>
#include <vector>

namespace pgm_oracle {

using Idx = int64_t;
using ID = int32_t;
using IntS = int8_t;
using cplx = std::complex<double>;
using IdxVector = std::vector<Idx>;

// common/common.hpp:73-107
constexpr double sqrt3 = std::numbers::sqrt3;
constexpr double inv_sqrt3 = std::numbers::inv_sqrt3;
constexpr double pi = std::numbers::pi;
constexpr cplx a2{-0.5, -sqrt3 / 2.0};
constexpr cplx a{-0.5, sqrt3 / 2.0};
constexpr double deg_30 = (1.0 / 6.0) * pi;
constexpr double numerical_tolerance = 1e-8;
constexpr double nan = std::numeric_limits<double>::quiet_NaN();
constexpr IntS na_IntS = std::numeric_limits<IntS>::min();
constexpr ID na_IntID = std::numeric_limits<ID>::min();
constexpr double base_power_3p = 1e6;
constexpr double base_power_1p = base_power_3p / 3.0;
constexpr double default_source_sk = 1e10;
constexpr double default_source_rx_ratio = 0.1;
constexpr double default_source_z01_ratio = 1.0;
constexpr double transformer_low_susceptance_ratio = 1e-8;
template <int B> constexpr double u_scale = (B == 1) ? 1.0 : inv_sqrt3;
template <int B> constexpr double base_power = (B == 1) ? base_power_3p : base_power_1p;

// error classes mirroring common/exception.hpp:89-110
struct SparseMatrixError : std::runtime_error {
    SparseMatrixError()
        : std::runtime_error{
              "Sparse matrix error, possibly singular matrix!\n"
              "If you get this error from state estimation, "
              "it might mean the system is not fully observable, i.e. not enough measurements.\n"
              "It might also mean that you are running into a corner case where PGM cannot resolve yet.\n"
              "See https://github.com/PowerGridModel/power-grid-model/issues/864."} {}
};
struct IterationDiverge : std::runtime_error {
    Idx num_iter;
    double max_dev;
    IterationDiverge(Idx n, double dev, double tol)
        : std::runtime_error{"Iteration failed to converge after " + std::to_string(n) +
                             " iterations! Max deviation: " + std::to_string(dev) +
                             ", error tolerance: " + std::to_string(tol) + ".\n"},
          num_iter{n},
          max_dev{dev} {}
    explicit IterationDiverge(std::string const& msg) : std::runtime_error{msg}, num_iter{0}, max_dev{0.0} {}
};
// common/exception.hpp:113-117 (thrown by the tap position optimizer; caught wherever IterationDiverge is)
struct MaxIterationReached : IterationDiverge {
    explicit MaxIterationReached(std::string const& msg = "") : IterationDiverge{"Maximum number of iterations reached! " + msg + "\n"} {}
};
struct PgmError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline double cabs(double x) { return std::abs(x); }
inline double cabs(cplx const& x) { return std::sqrt(std::norm(x)); } // three_phase_tensor.hpp cabs(complex)
inline double abs2(double x) { return x * x; }
inline double abs2(cplx const& x) { return std::norm(x); }
inline bool is_nan(double x) { return std::isnan(x); }

// three_phase_tensor.hpp:380-392
inline bool is_normal(double v) { return std::isnormal(v); }
inline bool is_normal(cplx const& v) {
    if (v.real() == 0.0) {
        return is_normal(v.imag());
    }
    if (v.imag() == 0.0) {
        return is_normal(v.real());
    }
    return is_normal(v.real()) && is_normal(v.imag());
}

// ---- vectors / tensors over phases -------------------------------------------------------------------------
template <class T, int B> struct Vec {
    T v[B]{};
    T& operator()(int i) { return v[i]; }
    T const& operator()(int i) const { return v[i]; }
};
template <class T, int B> struct Mat { // m[r][c]
    T m[B][B]{};
    T& operator()(int r, int c) { return m[r][c]; }
    T const& operator()(int r, int c) const { return m[r][c]; }
};
template <int B> using RVec = Vec<double, B>;
template <int B> using CVec = Vec<cplx, B>;
template <int B> using RMat = Mat<double, B>;
template <int B> using CMat = Mat<cplx, B>;

// ComplexValue<sym>{complex scalar}: asym = (x, x a^2, x a)  (three_phase_tensor.hpp:47-53)
template <int B> inline CVec<B> cvec_rotated(cplx const& x) {
    CVec<B> r;
    if constexpr (B == 1) {
        r.v[0] = x;
    } else {
        r.v[0] = x;
        r.v[1] = x * a2;
        r.v[2] = x * a;
    }
    return r;
}
// piecewise_complex_value / RealValue{double}: repeated without rotation
template <class T, int B> inline Vec<T, B> vec_piecewise(T const& x) {
    Vec<T, B> r;
    for (int i = 0; i < B; ++i) {
        r.v[i] = x;
    }
    return r;
}
// ComplexTensor<asym>{s, m}: diagonal s, off-diagonal m; {x}: diagonal only (three_phase_tensor.hpp:66-78)
template <int B> inline CMat<B> cmat_sm(cplx const& s, cplx const& m) {
    CMat<B> r;
    for (int i = 0; i < B; ++i) {
        for (int j = 0; j < B; ++j) {
            r.m[i][j] = (i == j) ? s : m;
        }
    }
    return r;
}
template <int B> inline CMat<B> cmat_diag(cplx const& x) { return cmat_sm<B>(x, cplx{0.0}); }

template <class T, int B> inline Vec<T, B> operator+(Vec<T, B> x, Vec<T, B> const& y) {
    for (int i = 0; i < B; ++i) x.v[i] += y.v[i];
    return x;
}
template <class T, int B> inline Vec<T, B> operator-(Vec<T, B> x, Vec<T, B> const& y) {
    for (int i = 0; i < B; ++i) x.v[i] -= y.v[i];
    return x;
}
template <class T, int B> inline Vec<T, B>& operator+=(Vec<T, B>& x, Vec<T, B> const& y) {
    for (int i = 0; i < B; ++i) x.v[i] += y.v[i];
    return x;
}
template <class T, int B> inline Mat<T, B>& operator+=(Mat<T, B>& x, Mat<T, B> const& y) {
    for (int i = 0; i < B; ++i)
        for (int j = 0; j < B; ++j) x.m[i][j] += y.m[i][j];
    return x;
}
template <class T, int B> inline Mat<T, B> operator-(Mat<T, B> x) {
    for (int i = 0; i < B; ++i)
        for (int j = 0; j < B; ++j) x.m[i][j] = -x.m[i][j];
    return x;
}
template <int B> inline CVec<B> conj(CVec<B> x) {
    for (int i = 0; i < B; ++i) x.v[i] = std::conj(x.v[i]);
    return x;
}
template <int B> inline CVec<B> operator*(CVec<B> x, CVec<B> const& y) { // element-wise (Eigen Array semantics)
    for (int i = 0; i < B; ++i) x.v[i] = x.v[i] * y.v[i];
    return x;
}
template <int B> inline CVec<B> operator/(CVec<B> x, CVec<B> const& y) {
    for (int i = 0; i < B; ++i) x.v[i] = x.v[i] / y.v[i];
    return x;
}
template <int B> inline CVec<B> operator*(CVec<B> x, double s) {
    for (int i = 0; i < B; ++i) x.v[i] = x.v[i] * s;
    return x;
}
template <int B> inline CVec<B> operator*(CVec<B> x, RVec<B> const& s) {
    for (int i = 0; i < B; ++i) x.v[i] = x.v[i] * s.v[i];
    return x;
}
template <int B> inline RVec<B> cabs(CVec<B> const& x) {
    RVec<B> r;
    for (int i = 0; i < B; ++i) r.v[i] = cabs(x.v[i]);
    return r;
}
template <int B> inline RVec<B> abs2(CVec<B> const& x) {
    RVec<B> r;
    for (int i = 0; i < B; ++i) r.v[i] = std::norm(x.v[i]);
    return r;
}
template <int B> inline double max_val(RVec<B> const& x) {
    double r = x.v[0];
    for (int i = 1; i < B; ++i) r = std::max(r, x.v[i]);
    return r;
}
// dot(tensor, vector): matrix-vector product (three_phase_tensor.hpp:245-273)
template <int B> inline CVec<B> dot(CMat<B> const& y, CVec<B> const& u) {
    CVec<B> r;
    for (int i = 0; i < B; ++i) {
        cplx s = y.m[i][0] * u.v[0];
        for (int k = 1; k < B; ++k) s += y.m[i][k] * u.v[k];
        r.v[i] = s;
    }
    return r;
}
template <int B> inline CMat<B> dot(CMat<B> const& x, CMat<B> const& y) {
    CMat<B> r;
    for (int i = 0; i < B; ++i)
        for (int j = 0; j < B; ++j) {
            cplx s = x.m[i][0] * y.m[0][j];
            for (int k = 1; k < B; ++k) s += x.m[i][k] * y.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
template <int B> inline void add_diag(CMat<B>& x, CVec<B> const& y) {
    for (int i = 0; i < B; ++i) x.m[i][i] += y.v[i];
}
template <int B> inline bool is_nan(RVec<B> const& x) { // Eigen isNaN().all()
    for (int i = 0; i < B; ++i)
        if (!std::isnan(x.v[i])) return false;
    return true;
}

// three_phase_tensor.hpp:449-459
inline CMat<3> get_sym_matrix() {
    CMat<3> m;
    cplx const v[9] = {1.0, 1.0, 1.0, 1.0, a2, a, 1.0, a, a2};
    for (int i = 0; i < 9; ++i) m.m[i / 3][i % 3] = v[i];
    return m;
}
inline CMat<3> get_sym_matrix_inv() {
    CMat<3> m;
    cplx const v[9] = {1.0, 1.0, 1.0, 1.0, a, a2, 1.0, a2, a};
    for (int i = 0; i < 9; ++i) m.m[i / 3][i % 3] = v[i] / 3.0;
    return m;
}

// three_phase_tensor.hpp:214-221  phase_shift(x) = x/|x| (1 if zero)
inline cplx phase_shift(cplx const& x) {
    double const ax = cabs(x);
    return ax > 0.0 ? x / ax : cplx{1.0};
}

} // namespace pgm_oracle
