// Shape factors of every deposition order (0..3) and derivative type (0..2), host + device.
//
// The reference tabulates them case by case (src/particles/particles_utils/ShapeFactors.H:40-466).
// All of those tables are one formula: with S_n the centred uniform B-spline of degree n and
// x = xmid - cell,
//     weight                       S_n(x)
//     type 0 (analytic)  "-sdx" = -S_n'(x)
//     type 1 (nodal)     "-sdx" =  S_n(x - 1/2) - S_n(x + 1/2)      on the stencil of order n+1
//     type 2 (centred)   "-sdx" = (S_n(x - 1) - S_n(x + 1)) / 2     on the order's stencil +- 1 cell
// so this file evaluates the B-spline pieces by the Cox-de Boor recursion (fully unrolled at compile
// time: straight-line polynomial code) instead of restating the tables.  Everything is
// __host__ __device__: tests/test_device_math_host.py runs these very functions on the CPU and
// holds them to the reference's own header (oracle/ref_headers.cpp) for every order and type.
//
// The order-2 kernels of particles.cu keep their hand-specialised shape2 / dshape2_* (common.cuh).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HPB_HD __host__ __device__ __forceinline__
#else
#define HPB_HD inline
#endif

// piece i (knot interval [i, i+1)) of the uniform B-spline of degree N on the knots 0..N+1, at the
// local coordinate f in [0, 1)
template <int N>
struct BsplinePiece {
    static HPB_HD double eval(int i, double f)
    {
        if (i < 0 || i > N) return 0.;
        return ((i + f) * BsplinePiece<N - 1>::eval(i, f)
                + ((N + 1 - i) - f) * BsplinePiece<N - 1>::eval(i - 1, f)) * (1.0 / N);
    }
    // d/dt B_N = B_{N-1}(t) - B_{N-1}(t - 1)
    static HPB_HD double deriv(int i, double f)
    {
        if (i < 0 || i > N) return 0.;
        return BsplinePiece<N - 1>::eval(i, f) - BsplinePiece<N - 1>::eval(i - 1, f);
    }
};
template <>
struct BsplinePiece<0> {
    static HPB_HD double eval(int i, double) { return i == 0 ? 1. : 0.; }
    static HPB_HD double deriv(int, double) { return 0.; }
};

HPB_HD int hpb_floor_half(int a) { return a >= 0 ? a / 2 : -((1 - a) / 2); }

// B_N at t = xint + halves/2 with 0 <= xint < 1.  A whole number of cells selects the piece
// outright; a half-cell offset needs the one comparison the reference's nodal tables have.
template <int N, bool DERIV>
HPB_HD double hpb_bspline_at(int halves, double xint)
{
    if ((halves & 1) == 0) {
        const int i = hpb_floor_half(halves);
        return DERIV ? BsplinePiece<N>::deriv(i, xint) : BsplinePiece<N>::eval(i, xint);
    }
    const int ilo = hpb_floor_half(halves - 1), ihi = ilo + 1;
    if (xint < 0.5)
        return DERIV ? BsplinePiece<N>::deriv(ilo, xint + 0.5) : BsplinePiece<N>::eval(ilo, xint + 0.5);
    return DERIV ? BsplinePiece<N>::deriv(ihi, xint - 0.5) : BsplinePiece<N>::eval(ihi, xint - 0.5);
}

// leftmost cell of an order-M stencil and the in-cell coordinate it is counted from: even M hang
// on the nearest cell (floor(xmid + 1/2)), odd M on floor(xmid)
template <int M>
HPB_HD int hpb_stencil_origin(double xmid, double &xint)
{
    const double xm = (M % 2 == 0) ? xmid + 0.5 : xmid;
    const double xf = floor(xm);
    xint = xm - xf;
    return (int)xf - ((M % 2 == 0) ? M / 2 : (M - 1) / 2);
}

// compute_shape_factor<ORDER>: ORDER+1 weights, returns the leftmost cell
template <int ORDER>
HPB_HD int hpb_shape(double xmid, double *w)
{
    double xint;
    const int cell = hpb_stencil_origin<ORDER>(xmid, xint);
#pragma unroll
    for (int k = 0; k <= ORDER; ++k) w[k] = hpb_bspline_at<ORDER, false>(2 * (ORDER - k), xint);
    return cell;
}

// single_derivative_shape_factor<DTYPE, ORDER>: ORDER+DTYPE+1 weights s[] and derivative weights
// ds[] (the sign convention of the reference: ds = "-sdx"), returns the leftmost cell
template <int DTYPE, int ORDER>
HPB_HD int hpb_dshape(double xmid, double *s, double *ds)
{
    constexpr int M = (DTYPE == 1) ? ORDER + 1 : ORDER;     // the stencil the cells hang on
    constexpr int E = (DTYPE == 2) ? 1 : 0;                 // grown by one cell on both sides
    double xint;
    const int cell = hpb_stencil_origin<M>(xmid, xint) - E;
#pragma unroll
    for (int k = 0; k <= ORDER + DTYPE; ++k) {
        const int h = M + ORDER + 2 * E - 2 * k;            // x + (ORDER+1)/2 = xint + h/2
        s[k] = hpb_bspline_at<ORDER, false>(h, xint);
        if (DTYPE == 0) ds[k] = -hpb_bspline_at<ORDER, true>(h, xint);
        else if (DTYPE == 1) ds[k] = hpb_bspline_at<ORDER, false>(h - 1, xint) - hpb_bspline_at<ORDER, false>(h + 1, xint);
        else ds[k] = 0.5 * (hpb_bspline_at<ORDER, false>(h - 2, xint) - hpb_bspline_at<ORDER, false>(h + 2, xint));
    }
    return cell;
}
