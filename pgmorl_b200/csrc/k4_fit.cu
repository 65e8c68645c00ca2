// K4 -- batched fit of the hyperbolic prediction model (one warp per fit, FLOAT64).
//
// Replaces the scipy call of predict_hyperbolic (morl/population_2d.py:56-108, population_3d.py:51-103):
//     least_squares(fun, ones(4), loss='soft_l1', f_scale=20, jac=jac,
//                   bounds=([0, .1, -5, -500], [A_ub, 20, 5, 500]))
// i.e. scipy's bounded Trust Region Reflective algorithm with the 'exact' trust-region solver
// (scipy/optimize/_lsq/least_squares.py:900-1030, trf.py:129-412, common.py). The port follows the
// numpy restatement in oracle/selection_oracle.py::trf_fit (which reproduces scipy bit for bit with
// LAPACK's SVD) line by line; the one substitution is the SVD of the (K+4) x 4 augmented Jacobian:
// Householder QR, then a one-sided Jacobi (Hestenes) SVD of the 4 x 4 triangular factor in registers
// (oracle/selection_oracle.py::qr_jacobi_svd), accurate to ~1e-16 relative but not bit-identical to gesdd.
//
//   model     f(x) = A (e^{a(x-b)} - 1) / (e^{a(x-b)} + 1) + c,   residual_i = (f(x_i) - y_i) w_i
//   per fit   K weighted points (x, y, w), upper bounds ub[4]; start ones(4); max_nfev = 400
//
// One warp owns one fit: lane l holds rows l, l+32, ...; every dot product over rows is a butterfly
// all-reduce, so all lanes carry identical copies of the 4-vectors and take identical branches.
#include "common.cuh"

namespace pgm {

constexpr int K4_WARPS = 1;          // one fit per CTA: a slow fit (400 evaluations) does not pin the warp slots of finished neighbours
constexpr double K4_EPS = 2.220446049250313e-16;
constexpr double K4_FSCALE = 20.0;

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ bool wall(bool p) { return __all_sync(0xffffffffu, p); }

struct Fit {
    const double *x, *y, *w;   // global [K]
    double *f, *J, *A, *fn;    // shared: f[K], J[K][4], A[K+4][4], fn[K+4]
    int K, lane;
    double lb[4], ub[4];
};

// residuals of the model at p into out[]; returns true if all finite
__device__ inline bool fit_fun(const Fit &q, const double p[4], double *out) {
    bool ok = true;
    for (int r = q.lane; r < q.K; r += 32) {
        const double e = exp(p[1] * (q.x[r] - p[2]));
        const double v = (p[0] * (e - 1.0) / (e + 1.0) + p[3] - q.y[r]) * q.w[r];
        out[r] = v;
        ok = ok && isfinite(v);
    }
    return wall(ok);
}

__device__ inline void fit_jac(const Fit &q, const double p[4]) {
    for (int r = q.lane; r < q.K; r += 32) {
        const double xb = q.x[r] - p[2], w = q.w[r];
        const double e = exp(p[1] * xb);
        const double e1 = e + 1.0, e12 = e1 * e1;
        q.J[4 * r + 0] = ((e - 1.0) / e1) * w;
        q.J[4 * r + 1] = (p[0] * xb * (2.0 * e) / e12) * w;
        q.J[4 * r + 2] = (p[0] * (-p[1]) * (2.0 * e) / e12) * w;
        q.J[4 * r + 3] = w;
    }
}

// soft_l1 with f_scale C: cost only
__device__ inline double fit_cost(const Fit &q, const double *f) {
    double s = 0.0;
    for (int r = q.lane; r < q.K; r += 32) {
        const double z = (f[r] / K4_FSCALE) * (f[r] / K4_FSCALE);
        s += 2.0 * (sqrt(1.0 + z) - 1.0);
    }
    return 0.5 * K4_FSCALE * K4_FSCALE * wsum(s);
}

// rho = soft_l1(f); cost = 0.5*sum(rho0); scale J and f for the robust loss (common.py:720-731); g = J^T f
__device__ inline double fit_robust_scale(const Fit &q, double g[4]) {
    double c = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
    for (int r = q.lane; r < q.K; r += 32) {
        const double fr = q.f[r];
        const double z = (fr / K4_FSCALE) * (fr / K4_FSCALE);
        const double t = 1.0 + z;
        const double rho0 = 2.0 * (sqrt(t) - 1.0) * (K4_FSCALE * K4_FSCALE);
        const double rho1 = 1.0 / sqrt(t);
        const double rho2 = (-0.5 * pow(t, -1.5)) / (K4_FSCALE * K4_FSCALE);
        c += rho0;
        double js = rho1 + 2.0 * rho2 * fr * fr;
        if (js < K4_EPS) js = K4_EPS;
        js = sqrt(js);
        const double fs = fr * (rho1 / js);
        q.f[r] = fs;
        const double j0 = q.J[4 * r] * js, j1 = q.J[4 * r + 1] * js, j2 = q.J[4 * r + 2] * js, j3 = q.J[4 * r + 3] * js;
        q.J[4 * r] = j0; q.J[4 * r + 1] = j1; q.J[4 * r + 2] = j2; q.J[4 * r + 3] = j3;
        g0 += j0 * fs; g1 += j1 * fs; g2 += j2 * fs; g3 += j3 * fs;
    }
    g[0] = wsum(g0); g[1] = wsum(g1); g[2] = wsum(g2); g[3] = wsum(g3);
    return 0.5 * wsum(c);
}

__device__ inline void cl_scaling(const Fit &q, const double x[4], const double g[4], double v[4], double dv[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = 1.0; dv[i] = 0.0;
        if (g[i] < 0.0) { v[i] = q.ub[i] - x[i]; dv[i] = -1.0; }
        if (g[i] > 0.0) { v[i] = x[i] - q.lb[i]; dv[i] = 1.0; }
    }
}

__device__ inline void strictly_feasible(const Fit &q, double x[4], double rstep) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double lo = q.lb[i], hi = q.ub[i], xi = x[i];
        double xn = xi;
        if (rstep == 0.0) {
            if (xi <= lo) xn = nextafter(lo, hi);
            if (xi >= hi) xn = nextafter(hi, lo);
        } else {
            const double ld = xi - lo, ud = hi - xi;
            const double lt = rstep * fmax(1.0, fabs(lo)), ut = rstep * fmax(1.0, fabs(hi));
            if (ld <= fmin(ud, lt)) xn = lo + lt;
            if (ud <= fmin(ld, ut)) xn = hi - ut;
        }
        if (xn < lo || xn > hi) xn = 0.5 * (lo + hi);
        x[i] = xn;
    }
}

__device__ inline double step_to_bound(const Fit &q, const double x[4], const double s[4], int hits[4]) {
    double st[4], mn = INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        st[i] = INFINITY;
        if (s[i] != 0.0) st[i] = fmax((q.lb[i] - x[i]) / s[i], (q.ub[i] - x[i]) / s[i]);
        mn = fmin(mn, st[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) hits[i] = (st[i] == mn) ? (s[i] > 0.0 ? 1 : (s[i] < 0.0 ? -1 : 0)) : 0;
    return mn;
}

__device__ inline double dot4(const double a[4], const double b[4]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]; }
__device__ inline double norm4(const double a[4]) { return sqrt(dot4(a, a)); }

// ||J_h s||^2 with J_h = J * d (rows in smem), and optionally (J_h s0).(J_h s), ||J_h s0||^2
__device__ inline void jh_products(const Fit &q, const double d[4], const double s[4], const double *s0, double &vv,
                                   double &uv, double &uu) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int r = q.lane; r < q.K; r += 32) {
        const double j0 = q.J[4 * r] * d[0], j1 = q.J[4 * r + 1] * d[1], j2 = q.J[4 * r + 2] * d[2], j3 = q.J[4 * r + 3] * d[3];
        const double v = j0 * s[0] + j1 * s[1] + j2 * s[2] + j3 * s[3];
        a += v * v;
        if (s0) {
            const double u = j0 * s0[0] + j1 * s0[1] + j2 * s0[2] + j3 * s0[3];
            b += u * v; c += u * u;
        }
    }
    vv = wsum(a);
    if (s0) { uv = wsum(b); uu = wsum(c); }
}

__device__ inline double eval_quad(const Fit &q, const double d[4], const double gh[4], const double s[4], const double diag[4]) {
    double vv, uv, uu;
    jh_products(q, d, s, nullptr, vv, uv, uu);
    double qd = vv;
    qd += s[0] * diag[0] * s[0] + s[1] * diag[1] * s[1] + s[2] * diag[2] * s[2] + s[3] * diag[3] * s[3];
    return 0.5 * qd + dot4(s, gh);
}

__device__ inline void min_quad1d(double a, double b, double lo, double hi, double c, double &t_out, double &y_out) {
    double tb = lo, yb = lo * (a * lo + b) + c;
    const double yh = hi * (a * hi + b) + c;
    if (yh < yb) { tb = hi; yb = yh; }          // np.argmin keeps the first minimum: order lo, hi, extremum
    if (a != 0.0) {
        const double ex = -0.5 * b / a;
        if (lo < ex && ex < hi) {
            const double ye = ex * (a * ex + b) + c;
            if (ye < yb) { tb = ex; yb = ye; }
        }
    }
    t_out = tb; y_out = yb;
}

// all-reduce of four values at once (the butterflies interleave, one latency instead of four)
__device__ __forceinline__ void wsum4(double &a, double &b, double &c, double &d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
        const double tc = __shfl_xor_sync(0xffffffffu, c, o), td = __shfl_xor_sync(0xffffffffu, d, o);
        a += ta; b += tb; c += tc; d += td;
    }
}

// Jacobi rotation that orthogonalises two columns with squared norms al, be and inner product ga
// (identity when they already are orthogonal to 1e-16); returns whether it rotated.
__device__ __forceinline__ bool jacobi_angle(double al, double be, double ga, double &cs, double &sn) {
    cs = 1.0; sn = 0.0;
    if (ga == 0.0 || ga * ga <= 1e-32 * (al * be)) return false;      // |ga| <= 1e-16 sqrt(al be), without the square root
    // tan of the rotation angle, t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)) with zeta = (be - al) / (2 ga),
    // written with one square root and one division: t = sign(d) h / (|d| + sqrt(d^2 + h^2)), d = be - al, h = 2 ga
    const double d = be - al, h = 2.0 * ga;
    const double t = copysign(1.0, d) * h / (fabs(d) + sqrt(d * d + h * h));
    cs = rsqrt(1.0 + t * t);
    sn = cs * t;
    return true;
}

// SVD of the augmented Jacobian A [(K+4) x 4] (rows across lanes, in smem) and uf = U^T f_aug, the two things
// trf.py:312-318 takes from numpy's svd. Like LAPACK's driver for tall matrices it first reduces A to its 4 x 4
// triangular factor: Householder QR with f_aug (q.fn, overwritten) carried as a fifth column, R and (Q^T f)[:4]
// replicated in every lane; then a one-sided Jacobi SVD of R entirely in registers (no shuffles), the two
// independent rotations of each round-robin round side by side. Singular values are left unsorted.
__device__ inline void qr_jacobi_svd(const Fit &q, double V[4][4], double s[4], double uf[4]) {
    const int rows = q.K + 4;
    double B[4][4], qtf[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double n2 = 0.0, e0 = 0.0, e1 = 0.0, e2 = 0.0;
        for (int r = q.lane; r < rows; r += 32)
            if (r >= j) { const double a = q.A[4 * r + j]; n2 += a * a; }
        n2 = wsum(n2);
        const double x0 = q.A[4 * j + j];
        double rowj[4], fj = q.fn[j];
#pragma unroll
        for (int c = 0; c < 4; ++c) rowj[c] = q.A[4 * j + c];
        __syncwarp();
        const double nx = sqrt(n2);
        if (nx == 0.0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) B[j][c] = (c < j) ? 0.0 : rowj[c];
            qtf[j] = fj;
            continue;
        }
        const double alpha = -copysign(nx, x0), v0 = x0 - alpha;
        const double beta = 1.0 / (nx * (nx + fabs(x0)));          // 2 / (v.v)
        // w_c = beta * v . A[:, c] for the columns right of j and for f
        double df = 0.0;
        for (int r = q.lane; r < rows; r += 32)
            if (r >= j) {
                const double vr = (r == j) ? v0 : q.A[4 * r + j];
                if (j < 3) e0 += vr * q.A[4 * r + (j + 1 < 4 ? j + 1 : 3)];
                if (j < 2) e1 += vr * q.A[4 * r + (j + 2 < 4 ? j + 2 : 3)];
                if (j < 1) e2 += vr * q.A[4 * r + 3];
                df += vr * q.fn[r];
            }
        wsum4(e0, e1, e2, df);
        e0 *= beta; e1 *= beta; e2 *= beta; df *= beta;
        for (int r = q.lane; r < rows; r += 32)
            if (r > j) {
                const double vr = q.A[4 * r + j];
                if (j < 3) q.A[4 * r + (j + 1 < 4 ? j + 1 : 3)] -= vr * e0;
                if (j < 2) q.A[4 * r + (j + 2 < 4 ? j + 2 : 3)] -= vr * e1;
                if (j < 1) q.A[4 * r + 3] -= vr * e2;
                q.fn[r] -= vr * df;
            }
#pragma unroll
        for (int c = 0; c < 4; ++c) B[j][c] = (c < j) ? 0.0 : rowj[c];
        B[j][j] = alpha;
        if (j < 3) B[j][j + 1 < 4 ? j + 1 : 3] = rowj[j + 1 < 4 ? j + 1 : 3] - v0 * e0;
        if (j < 2) B[j][j + 2 < 4 ? j + 2 : 3] = rowj[j + 2 < 4 ? j + 2 : 3] - v0 * e1;
        if (j < 1) B[j][3] = rowj[3] - v0 * e2;
        qtf[j] = fj - v0 * df;
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int rnd = 0; rnd < 3; ++rnd) {
            // round-robin pairs: (0,1)(2,3) | (0,2)(1,3) | (0,3)(1,2)
            const int p0 = 0, c0 = rnd + 1;
            const int p1 = (rnd == 0) ? 2 : 1, c1 = (rnd == 2) ? 2 : 3;
            double al0 = 0.0, be0 = 0.0, ga0 = 0.0, al1 = 0.0, be1 = 0.0, ga1 = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                al0 += B[i][p0] * B[i][p0]; be0 += B[i][c0] * B[i][c0]; ga0 += B[i][p0] * B[i][c0];
                al1 += B[i][p1] * B[i][p1]; be1 += B[i][c1] * B[i][c1]; ga1 += B[i][p1] * B[i][c1];
            }
            double cs0, sn0, cs1, sn1;
            const bool r0 = jacobi_angle(al0, be0, ga0, cs0, sn0);
            const bool r1 = jacobi_angle(al1, be1, ga1, cs1, sn1);
            rotated = rotated || r0 || r1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (r0) {
                    const double bp = B[i][p0], bq = B[i][c0], vp = V[i][p0], vq = V[i][c0];
                    B[i][p0] = cs0 * bp - sn0 * bq; B[i][c0] = sn0 * bp + cs0 * bq;
                    V[i][p0] = cs0 * vp - sn0 * vq; V[i][c0] = sn0 * vp + cs0 * vq;
                }
                if (r1) {
                    const double bp = B[i][p1], bq = B[i][c1], vp = V[i][p1], vq = V[i][c1];
                    B[i][p1] = cs1 * bp - sn1 * bq; B[i][c1] = sn1 * bp + cs1 * bq;
                    V[i][p1] = cs1 * vp - sn1 * vq; V[i][c1] = sn1 * vp + cs1 * vq;
                }
            }
        }
        if (!rotated) break;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double n2 = 0.0, a = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { n2 += B[i][c] * B[i][c]; a += B[i][c] * qtf[i]; }
        s[c] = sqrt(n2);
        uf[c] = s[c] > 0.0 ? a / s[c] : 0.0;
    }
}

// solve_lsq_trust_region (common.py:57-168); s unsorted (smax/smin passed), returns p_h and updates alpha
__device__ inline void solve_tr(int m, const double uf[4], const double s[4], const double V[4][4], double Delta,
                                double &alpha, double ph[4]) {
    double suf[4], smax = 0.0, smin = INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) { suf[i] = s[i] * uf[i]; smax = fmax(smax, s[i]); smin = fmin(smin, s[i]); }
    const bool full_rank = (m >= 4) && (smin > K4_EPS * m * smax);
    if (full_rank) {
        double t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = uf[i] / s[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) ph[i] = -(V[i][0] * t[0] + V[i][1] * t[1] + V[i][2] * t[2] + V[i][3] * t[3]);
        if (norm4(ph) <= Delta) { alpha = 0.0; return; }
    }
    auto phi_d = [&](double al, double &phi, double &phip) {
        double pn2 = 0.0, sp = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double den = s[i] * s[i] + al;
            const double qv = suf[i] / den;
            pn2 += qv * qv;
            sp += suf[i] * suf[i] / (den * den * den);
        }
        const double pn = sqrt(pn2);
        phi = pn - Delta; phip = -sp / pn;
    };
    double alpha_upper = norm4(suf) / Delta, alpha_lower = 0.0;
    if (full_rank) {
        double phi, phip;
        phi_d(0.0, phi, phip);
        alpha_lower = -phi / phip;
    }
    double al = alpha;
    if (!full_rank && alpha == 0.0) al = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
    for (int it = 0; it < 10; ++it) {
        if (al < alpha_lower || al > alpha_upper) al = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
        double phi, phip;
        phi_d(al, phi, phip);
        if (phi < 0.0) alpha_upper = al;
        const double ratio = phi / phip;
        alpha_lower = fmax(alpha_lower, al - ratio);
        al -= (phi + Delta) * ratio / Delta;
        if (fabs(phi) < 0.01 * Delta) break;
    }
    double t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = suf[i] / (s[i] * s[i] + al);
#pragma unroll
    for (int i = 0; i < 4; ++i) ph[i] = -(V[i][0] * t[0] + V[i][1] * t[1] + V[i][2] * t[2] + V[i][3] * t[3]);
    const double sc = Delta / norm4(ph);
#pragma unroll
    for (int i = 0; i < 4; ++i) ph[i] *= sc;
    alpha = al;
}

// select_step (trf.py:129-203); returns predicted reduction, fills step / step_h
__device__ inline double select_step(const Fit &q, const double x[4], const double d[4], const double diag[4],
                                     const double gh[4], double p[4], double ph[4], double Delta, double theta,
                                     double step[4], double step_h[4]) {
    bool inb = true;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const double xn = x[i] + p[i]; inb = inb && xn >= q.lb[i] && xn <= q.ub[i]; }
    if (inb) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { step[i] = p[i]; step_h[i] = ph[i]; }
        return -eval_quad(q, d, gh, ph, diag);
    }
    int hits[4], dummy[4];
    const double p_stride = step_to_bound(q, x, p, hits);
    double rh[4], r[4], xon[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { rh[i] = hits[i] ? -ph[i] : ph[i]; r[i] = d[i] * rh[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { p[i] *= p_stride; ph[i] *= p_stride; xon[i] = x[i] + p[i]; }
    // intersect_trust_region(p_h, r_h, Delta): positive root
    double to_tr;
    {
        const double a = dot4(rh, rh), b = dot4(ph, rh), c = dot4(ph, ph) - Delta * Delta;
        const double dd = sqrt(b * b - a * c);
        const double qq = -(b + copysign(dd, b));
        const double t1 = qq / a, t2 = c / qq;
        to_tr = fmax(t1, t2);
    }
    const double to_bound = step_to_bound(q, xon, r, dummy);
    double r_stride = fmin(to_bound, to_tr), lo, hi;
    if (r_stride > 0.0) {
        lo = (1.0 - theta) * p_stride / r_stride;
        hi = (r_stride == to_bound) ? theta * to_bound : to_tr;
    } else { lo = 0.0; hi = -1.0; }
    double r_value = INFINITY;
    if (lo <= hi) {
        double vv, uv, uu;
        jh_products(q, d, rh, ph, vv, uv, uu);
        double a = vv + (rh[0] * diag[0] * rh[0] + rh[1] * diag[1] * rh[1] + rh[2] * diag[2] * rh[2] + rh[3] * diag[3] * rh[3]);
        a *= 0.5;
        double b = dot4(gh, rh) + uv;
        double c = 0.5 * uu + dot4(gh, ph);
        b += ph[0] * diag[0] * rh[0] + ph[1] * diag[1] * rh[1] + ph[2] * diag[2] * rh[2] + ph[3] * diag[3] * rh[3];
        c += 0.5 * (ph[0] * diag[0] * ph[0] + ph[1] * diag[1] * ph[1] + ph[2] * diag[2] * ph[2] + ph[3] * diag[3] * ph[3]);
        min_quad1d(a, b, lo, hi, c, r_stride, r_value);
#pragma unroll
        for (int i = 0; i < 4; ++i) { rh[i] = rh[i] * r_stride + ph[i]; r[i] = rh[i] * d[i]; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { p[i] *= theta; ph[i] *= theta; }
    const double p_value = eval_quad(q, d, gh, ph, diag);
    double agh[4], ag[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { agh[i] = -gh[i]; ag[i] = d[i] * agh[i]; }
    const double to_tr2 = Delta / norm4(agh);
    const double to_b2 = step_to_bound(q, x, ag, dummy);
    double ag_stride = (to_b2 < to_tr2) ? theta * to_b2 : to_tr2, ag_value;
    {
        double vv, uv, uu;
        jh_products(q, d, agh, nullptr, vv, uv, uu);
        double a = vv + (agh[0] * diag[0] * agh[0] + agh[1] * diag[1] * agh[1] + agh[2] * diag[2] * agh[2] + agh[3] * diag[3] * agh[3]);
        a *= 0.5;
        const double b = dot4(gh, agh);
        min_quad1d(a, b, 0.0, ag_stride, 0.0, ag_stride, ag_value);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { agh[i] *= ag_stride; ag[i] *= ag_stride; }
    if (p_value < r_value && p_value < ag_value) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { step[i] = p[i]; step_h[i] = ph[i]; }
        return -p_value;
    }
    if (r_value < p_value && r_value < ag_value) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { step[i] = r[i]; step_h[i] = rh[i]; }
        return -r_value;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { step[i] = ag[i]; step_h[i] = agh[i]; }
    return -ag_value;
}

__global__ void __launch_bounds__(32 * K4_WARPS) k4_fit_kernel(const double *__restrict__ gx, const double *__restrict__ gy,
                                                               const double *__restrict__ gw, const int *__restrict__ klen,
                                                               const double *__restrict__ gub, double *__restrict__ theta,
                                                               int *__restrict__ status, int *__restrict__ nfev_out,
                                                               double *__restrict__ cost_out, int F, int Kmax, int max_nfev) {
    extern __shared__ __align__(16) double sm4[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fit = blockIdx.x * K4_WARPS + warp;
    if (fit >= F) return;
    const int per = 10 * Kmax + 20;
    Fit q;
    q.lane = lane; q.K = klen[fit];
    q.x = gx + (size_t)fit * Kmax; q.y = gy + (size_t)fit * Kmax; q.w = gw + (size_t)fit * Kmax;
    double *base = sm4 + (size_t)warp * per;
    q.f = base; q.J = base + Kmax; q.A = base + 5 * Kmax; q.fn = base + 9 * Kmax + 16;
    q.lb[0] = 0.0; q.lb[1] = 0.1; q.lb[2] = -5.0; q.lb[3] = -500.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) q.ub[i] = gub[4 * fit + i];
    const double ftol = 1e-8, xtol = 1e-8, gtol = 1e-8;
    const int m = q.K;

    double x[4] = {1.0, 1.0, 1.0, 1.0};
    strictly_feasible(q, x, 1e-10);
    fit_fun(q, x, q.f);
    int nfev = 1;
    fit_jac(q, x);
    __syncwarp();
    double g[4];
    double cost = fit_robust_scale(q, g);
    __syncwarp();
    double v[4], dv[4];
    cl_scaling(q, x, g, v, dv);
    double Delta;
    {
        double t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = x[i] / sqrt(v[i]);
        Delta = norm4(t);
        if (Delta == 0.0) Delta = 1.0;
    }
    double alpha = 0.0;
    int st = -1;     // -1 = None
    while (true) {
        cl_scaling(q, x, g, v, dv);
        double g_norm = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) g_norm = fmax(g_norm, fabs(g[i] * v[i]));
        if (g_norm < gtol) st = 1;
        if (st != -1 || nfev == max_nfev) break;
        double d[4], diag[4], gh[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { d[i] = sqrt(v[i]); diag[i] = g[i] * dv[i]; gh[i] = d[i] * g[i]; }
        // augmented Jacobian: rows [J * d ; diag(sqrt(diag_h))]
        for (int r = lane; r < m; r += 32) {
#pragma unroll
            for (int c = 0; c < 4; ++c) q.A[4 * r + c] = q.J[4 * r + c] * d[c];
        }
        if (lane < 4) {
#pragma unroll
            for (int c = 0; c < 4; ++c) q.A[4 * (m + lane) + c] = (c == lane) ? sqrt(diag[lane]) : 0.0;
        }
        __syncwarp();
        for (int r = lane; r < m + 4; r += 32) q.fn[r] = (r < m) ? q.f[r] : 0.0;     // f_augmented
        __syncwarp();
        double V[4][4], s[4], uf[4];
        qr_jacobi_svd(q, V, s, uf);
        const double theta_sb = fmax(0.995, 1.0 - g_norm);
        double actual = -1.0, cost_new = cost, xn[4];
        while (actual <= 0.0 && nfev < max_nfev) {
            double ph[4], p[4], step[4], step_h[4];
            solve_tr(m, uf, s, V, Delta, alpha, ph);
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = d[i] * ph[i];
            const double predicted = select_step(q, x, d, diag, gh, p, ph, Delta, theta_sb, step, step_h);
#pragma unroll
            for (int i = 0; i < 4; ++i) xn[i] = x[i] + step[i];
            strictly_feasible(q, xn, 0.0);
            const bool finite = fit_fun(q, xn, q.fn);
            ++nfev;
            __syncwarp();
            const double shn = norm4(step_h);
            if (!finite) { Delta = 0.25 * shn; continue; }
            cost_new = fit_cost(q, q.fn);
            actual = cost - cost_new;
            double ratio;
            if (predicted > 0.0) ratio = actual / predicted;
            else if (predicted == 0.0 && actual == 0.0) ratio = 1.0;
            else ratio = 0.0;
            double Delta_new = Delta;
            if (ratio < 0.25) Delta_new = 0.25 * shn;
            else if (ratio > 0.75 && shn > 0.95 * Delta) Delta_new = Delta * 2.0;
            const double step_norm = norm4(step);
            const bool ft = (actual < ftol * cost) && (ratio > 0.25);
            const bool xt = step_norm < xtol * (xtol + norm4(x));
            st = (ft && xt) ? 4 : (ft ? 2 : (xt ? 3 : -1));
            if (st != -1) break;
            alpha *= Delta / Delta_new;
            Delta = Delta_new;
        }
        if (actual > 0.0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = xn[i];
            for (int r = lane; r < m; r += 32) q.f[r] = q.fn[r];
            cost = cost_new;
            fit_jac(q, x);
            __syncwarp();
            (void)fit_robust_scale(q, g);      // robust-loss scaling of J, f and the new gradient; cost stays cost_new
            __syncwarp();
        }
    }
    if (st == -1) st = 0;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) theta[4 * fit + i] = x[i];
        status[fit] = st; nfev_out[fit] = nfev; cost_out[fit] = cost;
    }
}

}  // namespace pgm

using namespace pgm;

extern "C" int pgm_fit_hyperbolic_f64(const double *x, const double *y, const double *w, const int32_t *k_len,
                                      const double *ub, double *theta, int32_t *status, int32_t *nfev, double *cost,
                                      int F, int Kmax, void *stream) {
    PGM_REQUIRE(x && y && w && k_len && ub && theta && status && nfev && cost, "pgm_fit_hyperbolic_f64: null pointer");
    PGM_REQUIRE(F >= 0 && Kmax >= 1, "pgm_fit_hyperbolic_f64: bad sizes F=%d Kmax=%d", F, Kmax);
    if (F == 0) return PGM_OK;
    const size_t smem = (size_t)K4_WARPS * (10 * (size_t)Kmax + 20) * sizeof(double);
    PGM_REQUIRE(smem <= 200 * 1024, "pgm_fit_hyperbolic_f64: Kmax=%d exceeds the shared-memory budget (max 2558 points per fit)", Kmax);
    PGM_CUDA(cudaFuncSetAttribute(k4_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k4_fit_kernel<<<(F + K4_WARPS - 1) / K4_WARPS, 32 * K4_WARPS, smem, (cudaStream_t)stream>>>(
        x, y, w, k_len, ub, theta, status, nfev, cost, F, Kmax, 400);
    PGM_CUDA(cudaGetLastError());
    return PGM_OK;
}
