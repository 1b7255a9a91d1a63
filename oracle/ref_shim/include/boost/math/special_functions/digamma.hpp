// TEST INFRASTRUCTURE ONLY -- shim: boost::math special functions via libm (not on the likelihood path).
#pragma once
#include <cmath>
namespace boost { namespace math {
#ifndef REF_SHIM_BOOST_MATH
#define REF_SHIM_BOOST_MATH
inline double tgamma(double x) { return std::tgamma(x); }
inline double lgamma(double x) { return std::lgamma(x); }
inline double digamma(double x) {  // asymptotic series with recurrence; adequate for a stub
    double r = 0;
    while (x < 6) { r -= 1 / x; x += 1; }
    const double f = 1 / (x * x);
    return r + std::log(x) - 0.5 / x - f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f / 132))));
}
#endif
}}
