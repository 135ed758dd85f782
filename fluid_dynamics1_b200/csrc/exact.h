// exact.h -- individually rounded fp64 arithmetic shared by device kernels and host code.
//
// Parity with the reference (an x86-64 -O2 build without FMA) requires that every product
// and sum is a separately rounded IEEE-754 binary64 operation in the association order of
// the reference expression.  On the device the __d*_rn intrinsics are never contracted
// into DFMA by nvcc; host translation units are built with -ffp-contract=off.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define CNV_HD __host__ __device__ __forceinline__
#else
#define CNV_HD inline
#endif

namespace cnv {

CNV_HD double xadd(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
CNV_HD double xsub(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
CNV_HD double xmul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
CNV_HD double xfma(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}

// Correctly rounded a / d for a loop-invariant divisor d with rd = RN(1/d) computed once on
// the host.  q0 = RN(a*rd) is within one ulp of a/d; the residual r = a - q0*d is exact in
// one FMA; q0 + r*rd rounds to RN(a/d) (Markstein's theorem).  The host refuses this path
// (falls back to a true division) for the one divisor class where q0 may be further than
// one ulp away: an all-ones significand.  tests/test_host_logic.py checks it against the
// hardware division on random operands.
CNV_HD double xdiv_const(double a, double d, double rd)
{
    double q0 = xmul(a, rd);
    double r = xfma(-q0, d, a);
    return xfma(r, rd, q0);
}
CNV_HD double xdiv(double a, double d)
{
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, d);
#else
    return a / d;
#endif
}

// Constants of the 5-point relaxation, all evaluated on the host with the reference's own
// expressions (src/poisson.c:246): cyy = dy*dy, cxx = dx*dx, cf = dx*dx*dy*dy,
// D = 2*(dx*dx+dy*dy), omb = 1-beta.
struct RelaxConsts {
    double cxx, cyy, cf, D, rD, beta, omb;
    double bb;      // pow2 path: 0.25*beta
    double pscale;  // what the right-hand side is pre-multiplied with: cf (general) or dx*dx (pow2 path)
    int pow2;       // dx == dy == 2^-k: every scaling by cxx, cyy, cf, D is exact (see relax<true>)
    int true_div;   // general path: use the hardware division instead of xdiv_const
};

// One SOR cell update, src/poisson.c:246 / :257:
//   u = beta * (dy*dy*(uN+uS) + dx*dx*(uE+uW) - dx*dx*dy*dy*f) / (2*(dx*dx+dy*dy)) + (1-beta)*u0
// N,S = u[i+1][j], u[i-1][j]; E,W = u[i][j+1], u[i][j-1]; P = pscale*f (rounded once, like the
// reference's product); own = u0[i][j].
//
// POW2 (dx == dy == 2^-k): cxx = cyy = c, cf = c*c, D = 4c are powers of two, so
//   RN(c*t1 + c*t2) = c*RN(t1+t2),  RN(c*X' - c*c*f) = c*RN(X' - c*f),  RN(beta*c*X)/(4c) = RN((beta/4)*X)
// (scaling by a power of two commutes with rounding, barring underflow), i.e. the literal
// expression equals  RN( RN((beta/4) * RN(RN(t1+t2) - c*f)) + RN((1-beta)*u0) )  bit for bit.
template <bool POW2>
CNV_HD double relax(double N, double S, double E, double W, double own, double P, const RelaxConsts &c)
{
    double t1 = xadd(N, S), t2 = xadd(E, W);
    if (POW2) {
        double X = xsub(xadd(t1, t2), P);
        return xadd(xmul(c.bb, X), xmul(c.omb, own));
    } else {
        double A = xsub(xadd(xmul(c.cyy, t1), xmul(c.cxx, t2)), P);
        double m = xmul(c.beta, A);
        double q = c.true_div ? xdiv(m, c.D) : xdiv_const(m, c.D, c.rD);
        return xadd(q, xmul(c.omb, own));
    }
}

}  // namespace cnv
