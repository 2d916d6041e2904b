// match_core.h -- residual / Jacobian of one beam endpoint against the distance map and the 3x3
// Gauss-Newton / Levenberg-Marquardt step, shared by the fused sm_100a solver kernel and host code.
//
// Reference: MatchSurface2D::eval src/match_surface_2d.cpp:42-90, DynamicDistanceMap::distance
// src/sdm/dynamic_distance_map.cpp:66-92,140-147, robust weights src/nlls/robust_cost.cpp:36-82,
// GaussNewton::step/valid src/nlls/gauss_newton.cpp:53-86, LevenbergMarquard
// src/nlls/levenberg_marquardt.cpp:57-102, Solver::solve src/nlls/solver.cpp:53-117.
#pragma once

#include "lama_core.h"

namespace lama_b200 {

enum RobustKind : int { kRobustUnit = 0, kRobustCauchy = 1, kRobustHuber = 2, kRobustTukey = 3, kRobustTDist = 4 };
enum StrategyKind : int { kStrategyGN = 0, kStrategyLM = 1 };

LAMA_HD double robust_weight(int kind, double param, double x)
{
    switch (kind) {
    case kRobustCauchy: { double c = 1.0 / mul_rn(param, param); return 1.0 / add_rn(1.0, mul_rn(mul_rn(x, x), c)); }
    case kRobustHuber: return (x < param) ? 1.0 : (param / fabs(x));
    case kRobustTukey: { double bb = mul_rn(param, param), xx = mul_rn(x, x); if (xx <= bb) { double w = add_rn(1.0, -(xx / bb)); return mul_rn(w, w); } return 0.0; }
    case kRobustTDist: return add_rn(param, 1.0) / add_rn(param, mul_rn(x, x));
    default: return 1.0;
    }
}

// Bilinear distance + gradient from the four cell values (dynamic_distance_map.cpp:66-92).
// v[k] are already metric: sqrt(sqdist) * resolution.
struct BeamEval {
    double dist, gx, gy, hx, hy;
};
LAMA_HD BeamEval bilinear(const double v[4], double mu0, double mu1, double scale, double hx, double hy)
{
    const double nu0 = add_rn(1.0, -mu0), nu1 = add_rn(1.0, -mu1);
    BeamEval e;
    e.dist = add_rn(add_rn(add_rn(mul_rn(mul_rn(v[0], nu0), nu1), mul_rn(mul_rn(v[1], nu1), mu0)), mul_rn(mul_rn(v[2], nu0), mu1)),
                    mul_rn(mul_rn(v[3], mu0), mu1));
    e.gx = mul_rn(-add_rn(mul_rn(add_rn(v[0], -v[1]), nu1), mul_rn(add_rn(v[2], -v[3]), mu1)), scale);
    e.gy = mul_rn(-add_rn(mul_rn(add_rn(v[0], -v[2]), nu0), mul_rn(add_rn(v[1], -v[3]), mu0)), scale);
    e.hx = hx;
    e.hy = hy;
    return e;
}

// The ten sums of the weighted normal equations + the unweighted sum of squared distances.
//   s[0..5] = A00 A01 A02 A11 A12 A22   s[6..8] = g   s[9] = chi2   s[10] = sum d^2   s[11] = likelihood terms
constexpr int kNumSums = 12;
LAMA_HD void accumulate(double s[kNumSums], const BeamEval& e, int robust_kind, double robust_param, double meas_sigma)
{
    const double w  = sqrt(robust_weight(robust_kind, robust_param, e.dist));
    const double r  = mul_rn(e.dist, w);
    const double j0 = mul_rn(e.gx, w), j1 = mul_rn(e.gy, w);
    const double j2 = mul_rn(add_rn(mul_rn(e.gy, e.hx), -mul_rn(e.gx, e.hy)), w);  // match_surface_2d.cpp:88
    s[0] += j0 * j0; s[1] += j0 * j1; s[2] += j0 * j2;
    s[3] += j1 * j1; s[4] += j1 * j2; s[5] += j2 * j2;
    s[6] += j0 * r;  s[7] += j1 * r;  s[8] += j2 * r;
    s[9] += r * r;
    const double d2 = mul_rn(e.dist, e.dist);
    s[10] += d2;
    s[11] += -d2 / meas_sigma;  // pf_slam2d.cpp:410
}

// Solve A h = b, A symmetric 3x3 given by its upper triangle {A00,A01,A02,A11,A12,A22}, with an
// LDL^T that pivots on the largest remaining diagonal entry (Eigen::LDLT's strategy).
//
// Written with compile-time indices only: the pivot is applied by swapping NAMED scalars inside `if (piv == ...)`, the permutation is applied with
// selects, so that on the device every value lives in a register.  (The first version indexed M[piv][j] and b[perm[i]] at run time, which put the
// matrices into local memory: in k_match the one warp that solves -- while 16 others wait at the barrier -- then chased local-memory loads
// that missed the L1 half of the time.)  The floating-point operations and their order are those of ldlt_solve3_indexed below, which stays as
// the executable specification: tests/emu compares the two bit for bit.
LAMA_HD void ldlt_swap(double& x, double& y) { const double t = x; x = y; y = t; }
LAMA_HD void ldlt_solve3(const double a[6], const double b[3], double h[3])
{
    // the full matrix: both triangles are carried, because the trailing update computes M[i][j] and M[j][i] with different operand orders and a later
    // pivot swap can move an upper element into the lower triangle
    double m00 = a[0], m01 = a[1], m02 = a[2], m10 = a[1], m11 = a[3], m12 = a[4], m20 = a[2], m21 = a[4], m22 = a[5];
    int p0 = 0, p1 = 1, p2 = 2;
    // ---- k = 0 ----
    {
        int piv = 0;
        double best = fabs(m00);
        if (fabs(m11) > best) { best = fabs(m11); piv = 1; }
        if (fabs(m22) > best) { best = fabs(m22); piv = 2; }
        if (piv == 1) {          // rows 0 <-> 1, then columns 0 <-> 1
            const int t = p0; p0 = p1; p1 = t;
            ldlt_swap(m00, m10); ldlt_swap(m01, m11); ldlt_swap(m02, m12);
            ldlt_swap(m00, m01); ldlt_swap(m10, m11); ldlt_swap(m20, m21);
        } else if (piv == 2) {   // rows 0 <-> 2, then columns 0 <-> 2
            const int t = p0; p0 = p2; p2 = t;
            ldlt_swap(m00, m20); ldlt_swap(m01, m21); ldlt_swap(m02, m22);
            ldlt_swap(m00, m02); ldlt_swap(m10, m12); ldlt_swap(m20, m22);
        }
    }
    const double d0 = m00;
    double l10 = (d0 != 0.0) ? m10 / d0 : 0.0;
    double l20 = (d0 != 0.0) ? m20 / d0 : 0.0;
    m11 -= l10 * d0 * l10; m12 -= l10 * d0 * l20;
    m21 -= l20 * d0 * l10; m22 -= l20 * d0 * l20;
    // ---- k = 1 ----
    {
        if (fabs(m22) > fabs(m11)) {   // rows 1 <-> 2, columns 1 <-> 2, and the finished column of L
            const int t = p1; p1 = p2; p2 = t;
            ldlt_swap(m10, m20); ldlt_swap(m11, m21); ldlt_swap(m12, m22);
            ldlt_swap(m01, m02); ldlt_swap(m11, m12); ldlt_swap(m21, m22);
            ldlt_swap(l10, l20);
        }
    }
    const double d1 = m11;
    const double l21 = (d1 != 0.0) ? m21 / d1 : 0.0;
    m22 -= l21 * d1 * l21;
    // ---- k = 2 ----
    const double d2 = m22;
    // forward substitution on the permuted right-hand side, diagonal scaling, back substitution
    double y0 = p0 == 0 ? b[0] : (p0 == 1 ? b[1] : b[2]);
    double y1 = p1 == 0 ? b[0] : (p1 == 1 ? b[1] : b[2]);
    double y2 = p2 == 0 ? b[0] : (p2 == 1 ? b[1] : b[2]);
    y1 -= l10 * y0;
    y2 -= l20 * y0;
    y2 -= l21 * y1;
    y0 = (d0 != 0.0) ? y0 / d0 : 0.0;
    y1 = (d1 != 0.0) ? y1 / d1 : 0.0;
    y2 = (d2 != 0.0) ? y2 / d2 : 0.0;
    const double z2 = y2;
    double z1 = y1;
    z1 -= l21 * z2;
    double z0 = y0;
    z0 -= l10 * z1;
    z0 -= l20 * z2;
    h[0] = p0 == 0 ? z0 : (p1 == 0 ? z1 : z2);
    h[1] = p0 == 1 ? z0 : (p1 == 1 ? z1 : z2);
    h[2] = p0 == 2 ? z0 : (p1 == 2 ? z1 : z2);
}

// the same with run-time indices (the original formulation): executable specification of ldlt_solve3, used by the host tests only
LAMA_HD void ldlt_solve3_indexed(const double a[6], const double b[3], double h[3])
{
    double M[3][3] = {{a[0], a[1], a[2]}, {a[1], a[3], a[4]}, {a[2], a[4], a[5]}};
    int perm[3] = {0, 1, 2};
    double L[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, D[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        double best = fabs(M[k][k]);
        for (int i = k + 1; i < 3; ++i)
            if (fabs(M[i][i]) > best) { best = fabs(M[i][i]); piv = i; }
        if (piv != k) {
            int tp = perm[k]; perm[k] = perm[piv]; perm[piv] = tp;
            for (int j = 0; j < 3; ++j) { double t = M[k][j]; M[k][j] = M[piv][j]; M[piv][j] = t; }
            for (int i = 0; i < 3; ++i) { double t = M[i][k]; M[i][k] = M[i][piv]; M[i][piv] = t; }
            for (int j = 0; j < k; ++j) { double t = L[k][j]; L[k][j] = L[piv][j]; L[piv][j] = t; }
        }
        D[k] = M[k][k];
        for (int i = k + 1; i < 3; ++i) L[i][k] = (D[k] != 0.0) ? M[i][k] / D[k] : 0.0;
        for (int i = k + 1; i < 3; ++i)
            for (int j = k + 1; j < 3; ++j) M[i][j] -= L[i][k] * D[k] * L[j][k];
    }
    double y[3], z[3];
    for (int i = 0; i < 3; ++i) {
        y[i] = b[perm[i]];
        for (int j = 0; j < i; ++j) y[i] -= L[i][j] * y[j];
    }
    for (int i = 0; i < 3; ++i) y[i] = (D[i] != 0.0) ? y[i] / D[i] : 0.0;
    for (int i = 2; i >= 0; --i) {
        z[i] = y[i];
        for (int j = i + 1; j < 3; ++j) z[i] -= L[j][i] * z[j];
    }
    for (int i = 0; i < 3; ++i) h[perm[i]] = z[i];
}

// Iteration control of Solver::solve fused so that ONE evaluation per iteration suffices: the
// evaluation at the updated state provides the chi2 for the validity test of the previous step and,
// when the step is accepted, the normal equations of the next one (solver.cpp:66-104 recomputes the
// same values twice).  Drive it as:
//     ctl.begin(opts); state = initial;
//     loop { sums = evaluate(state); act = ctl.advance(sums, state); if (act == Done) break; }
struct SolverOptions {
    int strategy;      // StrategyKind
    int robust_kind;   // RobustKind
    double robust_param;
    uint32_t max_iterations;
    double eps1, eps2, tau;  // gauss_newton.cpp:40-41, levenberg_marquardt.cpp:41-43
};

struct SolverControl {
    SolverOptions o;
    uint32_t iter;
    int phase;        // 0 = need (r,J) at current state; 1 = waiting for the trial evaluation
    double chi2, mu, v;
    double g[3], h[3], a[6];
    uint32_t evals_ref;  // evaluations the reference's two-pass loop would have performed
    bool state_dirty;    // the state changed after the evaluation passed to the last advance()

    LAMA_HD void begin(const SolverOptions& opts)
    {
        o = opts;
        iter = 0;
        phase = 0;
        chi2 = 0;
        mu = -1;
        v = 2.0;
        evals_ref = 0;
        state_dirty = false;
        for (int i = 0; i < 3; ++i) g[i] = h[i] = 0;
        for (int i = 0; i < 6; ++i) a[i] = 0;
    }

    // strategy->step(): returns true when the step must be applied, false on stop.
    LAMA_HD bool step()
    {
        double mg = fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2])));
        if (mg < o.eps1) return false;
        double A[6] = {a[0], a[1], a[2], a[3], a[4], a[5]};
        if (o.strategy == kStrategyLM) {
            if (mu < 0) mu = o.tau * fmax(A[0], fmax(A[3], A[5]));
            A[0] += mu; A[3] += mu; A[5] += mu;
        }
        double ng[3] = {-g[0], -g[1], -g[2]};
        ldlt_solve3(A, ng, h);
        double mh = fmax(fabs(h[0]), fmax(fabs(h[1]), fabs(h[2])));
        if (mh < o.eps2) return false;  // stop raised; solver.cpp:84-86 discards the step
        return true;
    }

    // returns true when finished.  `state` is updated in place.
    LAMA_HD bool advance(const double s[kNumSums], SE2& state)
    {
        state_dirty = false;
        if (phase == 0) {
            if (iter >= o.max_iterations) return true;  // while-condition of solver.cpp:66
            for (int i = 0; i < 6; ++i) a[i] = s[i];
            g[0] = s[6]; g[1] = s[7]; g[2] = s[8];
            chi2 = s[9];
            ++evals_ref;
        } else {
            // trial evaluation: validity test of the step taken (gauss_newton.cpp:75-86, LM :83-102)
            ++evals_ref;
            double dF = chi2 - s[9];
            bool ok;
            if (o.strategy == kStrategyGN) {
                ok = dF > 0;
            } else {
                double dL = 0;
                for (int i = 0; i < 3; ++i) dL += h[i] * (mu * h[i] - g[i]);
                dL *= 0.5;
                ok = dL > 0.0 && dF > 0.0;
                if (ok) {
                    double q = 2 * (dF / dL) - 1;
                    mu = mu * fmax(1.0 / 3.0, 1 - q * q * q);
                    v  = 2.0;
                } else {
                    mu = mu * v;
                    v  = 2 * v;
                }
            }
            ++iter;
            if (ok) {
                // accepted: this evaluation is also the next iteration's (r, J)
                if (iter >= o.max_iterations) return true;
                for (int i = 0; i < 6; ++i) a[i] = s[i];
                g[0] = s[6]; g[1] = s[7]; g[2] = s[8];
                chi2 = s[9];
                ++evals_ref;
            } else {
                double nh[3] = {-h[0], -h[1], -h[2]};
                state = se2_mul(se2_exp(nh), state);  // problem.update(-h), solver.cpp:100-101
                state_dirty = true;
                if (o.strategy == kStrategyGN) return true;  // GaussNewton::valid raised stop
                if (iter >= o.max_iterations) return true;
                // LM retries with the larger damping from the SAME (r, J) (valid == false skips eval)
            }
        }
        if (!step()) return true;
        state = se2_mul(se2_exp(h), state);  // match_surface_2d.cpp:118-122
        state_dirty = true;
        phase = 1;
        return false;
    }
};

}  // namespace lama_b200
