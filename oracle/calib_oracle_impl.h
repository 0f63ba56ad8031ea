/*
 * calib_oracle_impl.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar, one-problem-at-a-time CPU restatement of the reference's batched calibration solve.
 * Included twice by calib_oracle.c with REAL = float / double and SUF = f32 / f64.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * use it, and only as the checker or the timed CPU arm.
 *
 * Each function cites the reference lines it follows, relative to
 * /root/reference/deep_attention_visual_odometry/.  The reference is batched and mask driven;
 * no operation couples problems, so running this per problem is equivalent (SURVEY.md
 * Appendix A).  Comparison operators are kept literally as written in the reference (every one
 * is false on NaN).  Compile with -ffp-contract=off so that no FMA is formed: the reference's
 * ATen CPU kernels round after every multiply.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)
#define R(x) ((REAL)(x))
#ifndef ORACLE_LANES
#define ORACLE_LANES 32 /* interleaved partial sums per reduction (power of two), see objective() */
#endif

typedef struct {
    int model, n, N, V, has_weights;
    const REAL* data0; /* DISTORT10: staged [N,4];  JOINT: world [N,3];  ANGLE_BA: obs [V,N,2] */
    const REAL* data1; /* JOINT: obs [V,N,2]                           */
    const REAL* w;     /* [V,N] or NULL                                */
} FN(problem);

/* Intrinsics + distortion of one match: camera_model/distorted_camera_model.py:59-86 (forward),
 * solvers/least_squares_utils.py:4-28 (residual, squared error), gradient = 2 J^T r derived from
 * that forward model (the reference's own Jacobian is wrong in 8 columns).
 * a = x'/z', b = y'/z'.  acc[0..9] += d(0.5 cost)/d(cx,cy,k1,k2,k3,p1,p2,fx,s,fy); returns
 * w (ru^2 + rv^2) and optionally (gu, gv) = d(0.5 cost)/d(u,v) for the pose chain rule. */
static inline REAL FN(match_intrinsics)(const REAL* th, REAL a, REAL b, REAL us, REAL vs, REAL w,
                                        REAL* acc, REAL* gu_out, REAL* gv_out) {
    const REAL cx = th[0], cy = th[1], k1 = th[2], k2 = th[3], k3 = th[4], p1 = th[5], p2 = th[6],
               fx = th[7], s = th[8], fy = th[9];
    REAL u = fx * a + s * b;                        /* :59-61 */
    REAL v = fy * b;                                /* :62    */
    REAL r2 = u * u + v * v;                        /* :64    */
    REAL uv = u * v;                                /* :65    */
    REAL rad = R(1.0) + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2; /* :66-74 */
    REAL up = u * rad + R(2.0) * p1 * uv + p2 * (r2 + R(2.0) * u * u) + cx; /* :75-80 */
    REAL vp = v * rad + R(2.0) * p2 * uv + p1 * (r2 + R(2.0) * v * v) + cy; /* :81-86 */
    REAL ru = up - us, rv = vp - vs;                /* least_squares_utils.py:13 */
    REAL cost = ru * ru + rv * rv;                  /* :24-28 */
    if (w != R(1.0)) cost = w * cost;
    if (acc) {
        REAL wru = w * ru, wrv = w * rv;
        REAL r4 = r2 * r2, r6 = r4 * r2;
        REAL radp = k1 + R(2.0) * k2 * r2 + R(3.0) * k3 * r4;
        REAL uv2 = R(2.0) * uv;
        REAL A = r2 + R(2.0) * u * u, Bv = r2 + R(2.0) * v * v;
        REAL Duu = rad + R(2.0) * u * u * radp + R(2.0) * p1 * v + R(6.0) * p2 * u;
        REAL Dvv = rad + R(2.0) * v * v * radp + R(6.0) * p1 * v + R(2.0) * p2 * u;
        REAL Duv = uv2 * radp + R(2.0) * p1 * u + R(2.0) * p2 * v;
        REAL gu = wru * Duu + wrv * Duv;
        REAL gv = wru * Duv + wrv * Dvv;
        REAL t = wru * u + wrv * v;
        acc[0] += wru;
        acc[1] += wrv;
        acc[2] += t * r2;
        acc[3] += t * r4;
        acc[4] += t * r6;
        acc[5] += wru * uv2 + wrv * Bv;
        acc[6] += wru * A + wrv * uv2;
        acc[7] += gu * a;
        acc[8] += gu * b;
        acc[9] += gv * b;
        if (gu_out) {
            *gu_out = gu;
            *gv_out = gv;
        }
    }
    return cost;
}

/* R = Rz Ry Rx and the three d R / d r_k: distorted_camera_model.py:29-55. */
static void FN(euler)(REAL rx, REAL ry, REAL rz, REAL Rm[9], REAL dRx[9], REAL dRy[9], REAL dRz[9]) {
    REAL sx = R(sin((double)rx)), cx = R(cos((double)rx));
    REAL sy = R(sin((double)ry)), cy = R(cos((double)ry));
    REAL sz = R(sin((double)rz)), cz = R(cos((double)rz));
    Rm[0] = cy * cz; Rm[1] = sx * sy * cz - cx * sz; Rm[2] = cx * sy * cz + sx * sz;
    Rm[3] = cy * sz; Rm[4] = sx * sy * sz + cx * cz; Rm[5] = cx * sy * sz - sx * cz;
    Rm[6] = -sy;     Rm[7] = sx * cy;                Rm[8] = cx * cy;
    if (dRx) {
        dRx[0] = 0; dRx[1] = cx * sy * cz + sx * sz;  dRx[2] = -sx * sy * cz + cx * sz;
        dRx[3] = 0; dRx[4] = cx * sy * sz - sx * cz;  dRx[5] = -sx * sy * sz - cx * cz;
        dRx[6] = 0; dRx[7] = cx * cy;                 dRx[8] = -sx * cy;
        dRy[0] = -sy * cz; dRy[1] = sx * cy * cz; dRy[2] = cx * cy * cz;
        dRy[3] = -sy * sz; dRy[4] = sx * cy * sz; dRy[5] = cx * cy * sz;
        dRy[6] = -cy;      dRy[7] = -sx * sy;     dRy[8] = -cx * sy;
        dRz[0] = -cy * sz; dRz[1] = -sx * sy * sz - cx * cz; dRz[2] = -cx * sy * sz + sx * cz;
        dRz[3] = cy * cz;  dRz[4] = sx * sy * cz - cx * sz;  dRz[5] = cx * sy * cz + sx * sz;
        dRz[6] = 0; dRz[7] = 0; dRz[8] = 0;
    }
}

/* Extrinsic transform of one point, distorted_camera_model.py:38-57 (z' == 0 -> += 1e-8). */
static inline void FN(transform)(const REAL Rm[9], const REAL t[3], const REAL X[3], REAL Xp[3]) {
    Xp[0] = X[0] * Rm[0] + X[1] * Rm[1] + X[2] * Rm[2] + t[0];
    Xp[1] = X[0] * Rm[3] + X[1] * Rm[4] + X[2] * Rm[5] + t[1];
    Xp[2] = X[0] * Rm[6] + X[1] * Rm[7] + X[2] * Rm[8] + t[2];
    if (Xp[2] == R(0.0)) Xp[2] += R(1e-8);
}

/* sin(a)/a, (1 - cos a)/a^2 and their derivatives as the reference's custom autograd Functions compute them:
 * utils/func_sin_x_on_x.py:5-43 (Taylor below 0.01, backward = x * (cos x / x^2 - sin x / x^3) with that
 * function's own Taylor branch below 0.01, :45-75) and utils/func_one_minus_cos_x_on_x_squared.py:6-58
 * (Taylor below 0.05, backward = (1/x) (sin x / x - 2 result), 1/x := 0 at x == 0). */
static void FN(sinc_terms)(REAL a, REAL* s, REAL* oc, REAL* ds, REAL* doc) {
    const REAL a2 = a * a;
    REAL k;
    if (R(fabs((double)a)) < R(0.01)) {
        const REAL a4 = a2 * a2, a6 = a4 * a2;
        *s = R(1.0) - a2 / R(6.0) + a4 / R(120.0) - a6 / R(5040.0);
        k = R(-1.0) / R(3.0) + a2 / R(30.0) - a4 / R(840.0) + a6 / R(45360.0);
    } else {
        const REAL sn = R(sin((double)a)), cs = R(cos((double)a));
        *s = sn / a;
        k = cs / a2 - sn / (a * a2);
    }
    *ds = a * k;
    if (R(fabs((double)a)) < R(0.05)) {
        const REAL a4 = a2 * a2, a6 = a4 * a2;
        *oc = R(0.5) - a2 / R(24.0) + a4 / R(720.0) - a6 / R(40320.0);
    } else {
        *oc = (R(1.0) - R(cos((double)a))) / a2;
    }
    const REAL rec = (a == R(0.0)) ? R(0.0) : R(1.0) / a;
    *doc = rec * (*s - R(2.0) * *oc);
}

/* The entry script's objective, networks/calibration_network.py:58-67, and its gradient (reverse mode written
 * out by hand; matches torch.autograd of the reference functions to 1e-15 relative in float64, including the
 * zero sub-gradients autograd uses for |x| at 0, vector_norm at 0 and the clamp in the normalisation).
 * x = (f, cx, cy | X[N][3] | t[V-1][3] | w[V-1][3]); obs [V,N,2]; vis [V,N] or NULL. */
static REAL FN(angle_ba)(const FN(problem)* p, const REAL* x, REAL* g) {
    const int N = p->N, V = p->V, n = p->n;
    const REAL EPS = R(2.220446049250313e-16); /* projective_plane_angle_distance.py:48,51 */
    const REAL* X = x + 3;
    const REAL* t = X + 3 * N;
    const REAL* w = t + 3 * (V - 1);
    /* geometry/homogeneous_projection.py:37: f' = elu(f) + 1 */
    const REAL fp = x[0] > R(0.0) ? x[0] + R(1.0) : R(exp((double)x[0]));
    const REAL dfp = x[0] > R(0.0) ? R(1.0) : R(exp((double)x[0]));
    /* calibration_pinhole_camera_model.py:98-104: one scale for points and translations */
    REAL ps = 0, cs = 0;
    for (int i = 0; i < 3 * N; ++i) ps += R(fabs((double)X[i]));
    for (int i = 0; i < 3 * (V - 1); ++i) cs += R(fabs((double)t[i]));
    ps /= R(3 * N);
    cs /= R(3 * (V - 1));
    const REAL sig = (ps * R(N) + cs * R(V)) / R(N + V);
    REAL gsig_num = 0; /* sum gXs . Xs + sum gts . ts */
    REAL cost = 0;
    if (g) for (int j = 0; j < n; ++j) g[j] = 0;
    REAL* gX = g ? g + 3 : NULL;           /* accumulates d/dXs first */
    REAL* gt = g ? gX + 3 * N : NULL;      /* d/dts */
    REAL* gw = g ? gt + 3 * (V - 1) : NULL;
    for (int m = 0; m < V; ++m) {
        REAL om[3] = {0, 0, 0}, ts[3] = {0, 0, 0}, ang = 0, c = 1, sn = 0, s = 1, oc = R(0.5), ds = 0, doc = 0;
        if (m > 0) {
            for (int k = 0; k < 3; ++k) { om[k] = w[3 * (m - 1) + k]; ts[k] = t[3 * (m - 1) + k] / sig; }
            ang = R(sqrt((double)(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]))); /* axis_angle_rotation.py:37 */
            c = R(cos((double)ang));
            sn = R(sin((double)ang));
            FN(sinc_terms)(ang, &s, &oc, &ds, &doc);
        }
        for (int j = 0; j < N; ++j) {
            const REAL xs[3] = {X[3 * j] / sig, X[3 * j + 1] / sig, X[3 * j + 2] / sig};
            REAL P[3], cr[3] = {0, 0, 0}, dot = 0;
            if (m == 0) {
                P[0] = xs[0]; P[1] = xs[1]; P[2] = xs[2];
            } else { /* axis_angle_rotation.py:38-48, then + translation (calibration_pinhole_camera_model.py:110) */
                dot = xs[0] * om[0] + xs[1] * om[1] + xs[2] * om[2];
                cr[0] = om[1] * xs[2] - om[2] * xs[1];
                cr[1] = om[2] * xs[0] - om[0] * xs[2];
                cr[2] = om[0] * xs[1] - om[1] * xs[0];
                for (int k = 0; k < 3; ++k) P[k] = xs[k] * c + oc * dot * om[k] + cr[k] * s + ts[k];
            }
            const REAL* ob = p->data0 + 2 * ((size_t)m * N + j);
            const REAL vis = p->w ? p->w[(size_t)m * N + j] : R(1.0);
            const REAL h[3] = {ob[0] - x[1], ob[1] - x[2], fp}; /* homogeneous_projection.py:38-44 */
            REAL nh = R(sqrt((double)(h[0] * h[0] + h[1] * h[1] + h[2] * h[2])));
            REAL nP = R(sqrt((double)(P[0] * P[0] + P[1] * P[1] + P[2] * P[2])));
            if (nh < EPS) nh = EPS;
            if (nP < EPS) nP = EPS;
            REAL a[3], b[3], sv[3], dv[3];
            for (int k = 0; k < 3; ++k) {
                a[k] = h[k] / nh; b[k] = P[k] / nP;
                sv[k] = a[k] + b[k]; dv[k] = a[k] - b[k];
            }
            const REAL S = R(sqrt((double)(sv[0] * sv[0] + sv[1] * sv[1] + sv[2] * sv[2])));
            const REAL D = R(sqrt((double)(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2])));
            cost += R(2.0) * R(atan2((double)D, (double)S)) * vis; /* projective_plane_angle_distance.py:53-60 */
            if (!g) continue;
            const REAL den = S * S + D * D;
            const REAL gD = R(2.0) * vis * S / den, gS = R(-2.0) * vis * D / den;
            REAL ga[3], gb[3], gaa = 0, gbb = 0;
            for (int k = 0; k < 3; ++k) {
                const REAL uD = (D != R(0.0)) ? dv[k] / D : R(0.0);
                const REAL uS = (S != R(0.0)) ? sv[k] / S : R(0.0);
                ga[k] = gD * uD + gS * uS;
                gb[k] = gS * uS - gD * uD;
                gaa += ga[k] * a[k];
                gbb += gb[k] * b[k];
            }
            REAL gh[3], gP[3];
            for (int k = 0; k < 3; ++k) {
                gh[k] = (ga[k] - a[k] * gaa) / nh;
                gP[k] = (gb[k] - b[k] * gbb) / nP;
            }
            g[1] -= gh[0];
            g[2] -= gh[1];
            g[0] += gh[2] * dfp;
            REAL gxs[3];
            if (m == 0) {
                gxs[0] = gP[0]; gxs[1] = gP[1]; gxs[2] = gP[2];
            } else {
                const REAL wg = om[0] * gP[0] + om[1] * gP[1] + om[2] * gP[2];
                const REAL xg = xs[0] * gP[0] + xs[1] * gP[1] + xs[2] * gP[2];
                const REAL crg = cr[0] * gP[0] + cr[1] * gP[1] + cr[2] * gP[2];
                const REAL gxo[3] = {gP[1] * om[2] - gP[2] * om[1], gP[2] * om[0] - gP[0] * om[2],
                                     gP[0] * om[1] - gP[1] * om[0]};                       /* gP x w */
                const REAL xxg[3] = {xs[1] * gP[2] - xs[2] * gP[1], xs[2] * gP[0] - xs[0] * gP[2],
                                     xs[0] * gP[1] - xs[1] * gP[0]};                       /* Xs x gP */
                const REAL gang = -sn * xg + doc * dot * wg + ds * crg;
                for (int k = 0; k < 3; ++k) {
                    gxs[k] = c * gP[k] + oc * wg * om[k] + s * gxo[k];
                    REAL gwk = oc * (dot * gP[k] + wg * xs[k]) + s * xxg[k];
                    if (ang != R(0.0)) gwk += gang * om[k] / ang;
                    gw[3 * (m - 1) + k] += gwk;
                    gt[3 * (m - 1) + k] += gP[k];
                    gsig_num += gP[k] * ts[k];
                }
            }
            for (int k = 0; k < 3; ++k) {
                gX[3 * j + k] += gxs[k];
                gsig_num += gxs[k] * xs[k];
            }
        }
    }
    if (g) {
        const REAL gsig = -gsig_num / sig;
        const REAL cX = gsig / R(3 * (N + V)), cT = gsig * R(V) / R(3 * (V - 1) * (N + V));
        for (int i = 0; i < 3 * N; ++i) {
            const REAL sg = X[i] > R(0.0) ? R(1.0) : (X[i] < R(0.0) ? R(-1.0) : R(0.0));
            gX[i] = gX[i] / sig + cX * sg;
        }
        for (int i = 0; i < 3 * (V - 1); ++i) {
            const REAL sg = t[i] > R(0.0) ? R(1.0) : (t[i] < R(0.0) ? R(-1.0) : R(0.0));
            gt[i] = gt[i] / sig + cT * sg;
        }
    }
    return cost;
}

/* cost and gradient of one problem.  g may be NULL (cost only). */
static REAL FN(objective)(const FN(problem)* p, const REAL* x, REAL* g) {
    const int n = p->n;
    if (p->model == DAVO_MODEL_ANGLE_BA) return FN(angle_ba)(p, x, g);
    /* Sums over matches.  The reference sums with ATen's vectorised cascade kernels (torch.sum and the
     * sum_to_size reductions of autograd's broadcast backward: SIMD lanes x 4 interleaved accumulators, then a
     * tree over the partial sums), NOT one running scalar sum.  In float32 that difference is visible at
     * population level on ill-conditioned problems (config 4, 2048 problems: the reference takes 143.5 steps on
     * average, a running-sum restatement 135.6, this lane-strided form 145.0), so the restatement keeps
     * ORACLE_LANES interleaved partial sums per quantity and combines them pairwise. */
    if (p->model == DAVO_MODEL_DISTORT10) {
        REAL part[ORACLE_LANES][11];
        memset(part, 0, sizeof(part));
        for (int m = 0; m < p->N; ++m) {
            const REAL* q = p->data0 + 4 * (size_t)m;
            REAL w = p->w ? p->w[m] : R(1.0);
            REAL* acc = part[m % ORACLE_LANES];
            acc[10] += FN(match_intrinsics)(x, q[0], q[1], q[2], q[3], w, g ? acc : NULL, NULL, NULL);
        }
        for (int s = ORACLE_LANES / 2; s > 0; s >>= 1)
            for (int l = 0; l < s; ++l)
                for (int j = 0; j < 11; ++j) part[l][j] += part[l + s][j];
        if (g)
            for (int j = 0; j < 10; ++j) g[j] = R(2.0) * part[0][j]; /* least_squares_utils.py:43 */
        return part[0][10];
    }
    if (p->model == DAVO_MODEL_JOINT) {
        REAL acc[10] = {0};
        REAL cost = 0;
        for (int v = 0; v < p->V; ++v) {
            const REAL* pose = x + 10 + 6 * v;
            REAL Rm[9], dRx[9], dRy[9], dRz[9];
            FN(euler)(pose[0], pose[1], pose[2], Rm, g ? dRx : NULL, dRy, dRz);
            /* per lane: [0..9] intrinsic sums, [10] cost, [11..19] M = sum gX' (x) X, [20..22] sum gX' */
            REAL part[ORACLE_LANES][23];
            memset(part, 0, sizeof(part));
            for (int m = 0; m < p->N; ++m) {
                const REAL* X = p->data0 + 3 * (size_t)m;
                const REAL* ob = p->data1 + 2 * ((size_t)v * p->N + m);
                REAL w = p->w ? p->w[(size_t)v * p->N + m] : R(1.0);
                REAL* pl = part[m % ORACLE_LANES];
                REAL Xp[3];
                FN(transform)(Rm, pose + 3, X, Xp);
                REAL a = Xp[0] / Xp[2], b = Xp[1] / Xp[2]; /* :59-62 */
                REAL gu, gv;
                pl[10] += FN(match_intrinsics)(x, a, b, ob[0], ob[1], w, g ? pl : NULL, &gu, &gv);
                if (g) {
                    REAL iz = R(1.0) / Xp[2];
                    REAL gA = gu * x[7];
                    REAL gB = gu * x[8] + gv * x[9];
                    REAL gX[3] = {gA * iz, gB * iz, -(gA * a + gB * b) * iz};
                    for (int r = 0; r < 3; ++r) {
                        pl[20 + r] += gX[r];
                        for (int c = 0; c < 3; ++c) pl[11 + 3 * r + c] += gX[r] * X[c];
                    }
                }
            }
            for (int s = ORACLE_LANES / 2; s > 0; s >>= 1)
                for (int l = 0; l < s; ++l)
                    for (int j = 0; j < 23; ++j) part[l][j] += part[l + s][j];
            const REAL* M = part[0] + 11;
            const REAL* gt = part[0] + 20;
            cost += part[0][10];
            for (int j = 0; j < 10; ++j) acc[j] += part[0][j];
            if (g) {
                REAL drx = 0, dry = 0, drz = 0;
                for (int e = 0; e < 9; ++e) {
                    drx += dRx[e] * M[e];
                    dry += dRy[e] * M[e];
                    drz += dRz[e] * M[e];
                }
                REAL* gp = g + 10 + 6 * v;
                gp[0] = R(2.0) * drx; gp[1] = R(2.0) * dry; gp[2] = R(2.0) * drz;
                gp[3] = R(2.0) * gt[0]; gp[4] = R(2.0) * gt[1]; gp[5] = R(2.0) * gt[2];
            }
        }
        if (g)
            for (int j = 0; j < 10; ++j) g[j] = R(2.0) * acc[j];
        return cost;
    }
    /* analytic objectives: tests/autograd_solvers/reference_functions.py:20-62,
     * tests/autograd_solvers/test_bfgs_solver.py:33-46,
     * tests/autograd_solvers/line_search/test_wolffe_conditions.py:214-305 */
    if (p->model == DAVO_MODEL_DISTANCE) {
        REAL dd = 0;
        for (int j = 0; j < n; ++j) dd += (x[j] - p->data0[j]) * (x[j] - p->data0[j]);
        REAL nrm = R(sqrt((double)dd));
        /* torch's vector_norm backward uses the zero subgradient at the origin */
        if (g) for (int j = 0; j < n; ++j) g[j] = nrm == R(0.0) ? R(0.0) : (x[j] - p->data0[j]) / nrm;
        return nrm;
    }
    REAL ss = 0;
    for (int j = 0; j < n; ++j) ss += x[j] * x[j];
    switch (p->model) {
        case DAVO_MODEL_SPHERE:
        case DAVO_MODEL_SPHERE_OFFSET:
            if (g) for (int j = 0; j < n; ++j) g[j] = R(2.0) * x[j];
            return p->model == DAVO_MODEL_SPHERE ? ss : ss + R(10.0);
        case DAVO_MODEL_LOG_SPHERE: {
            REAL d = ss + R(1.0);
            if (g) for (int j = 0; j < n; ++j) g[j] = R(2.0) * x[j] / d;
            return R(log((double)d));
        }
        case DAVO_MODEL_ROSENBROCK: {
            REAL a = R(1.0) - x[0], b = x[1] - x[0] * x[0];
            if (g) {
                g[0] = R(-2.0) * a - R(400.0) * x[0] * b;
                g[1] = R(200.0) * b;
            }
            return a * a + R(100.0) * b * b;
        }
        case DAVO_MODEL_COSINE: {
            REAL nrm = R(sqrt((double)ss));
            REAL d = R(1.0) - nrm;
            if (g)
                for (int j = 0; j < n; ++j) {
                    /* d/dx_j [1 - x0/|x|] = x0 x_j/|x|^3 - [j==0]/|x| ; d/dx_j (1-|x|)^2 = -2(1-|x|) x_j/|x| */
                    REAL t = x[0] * x[j] / (nrm * nrm * nrm) - (j == 0 ? R(1.0) / nrm : R(0.0));
                    g[j] = t - R(2.0) * d * x[j] / nrm;
                }
            return (R(1.0) - x[0] / nrm) + d * d;
        }
        case DAVO_MODEL_X2_SINE: {
            REAL nrm = R(sqrt((double)ss));
            REAL sn = R(sin((double)nrm)), cs = R(cos((double)nrm));
            /* f = r^2 (sin r + 2); df/dr = 2 r (sin r + 2) + r^2 cos r ; dr/dx_j = x_j / r */
            if (g)
                for (int j = 0; j < n; ++j)
                    g[j] = (R(2.0) * (sn + R(2.0)) + nrm * cs) * x[j];
            return nrm * nrm * (sn + R(2.0));
        }
        default:
            if (g) for (int j = 0; j < n; ++j) g[j] = R(NAN);
            return R(NAN);
    }
}

/* utils/func_inverse_curvature.py:8-11 : 1/(y.s), forced to 0 when y.s <= 0. */
static inline REAL FN(inverse_curvature)(const REAL* s, const REAL* y, int n) {
    REAL c = 0;
    for (int j = 0; j < n; ++j) c += s[j] * y[j];
    REAL inv = R(1.0) / c;
    if (c <= R(0.0)) inv = R(0.0);
    return inv;
}

/* autograd_solvers/bfgs_solver.py:217-233 (eq. 6.20). */
static inline REAL FN(initial_scale)(const REAL* s, const REAL* y, int n) {
    REAL den = 0, num = 0;
    for (int j = 0; j < n; ++j) den += y[j] * y[j];
    if (den < R(1e-5)) den = R(1e-5);
    for (int j = 0; j < n; ++j) num += s[j] * y[j];
    REAL sc = num / den;
    if (sc < R(1e-4)) sc = R(1e-4);
    return sc;
}

/* autograd_solvers/bfgs_solver.py:235-303 (eq. 6.17), same association order:
 * rho first, then H + (s rho) s^T (1+q) - (s rho)(y^T H) - (H y)(s rho)^T, old H on the right. */
static void FN(update_inverse_hessian)(REAL* H, const REAL* s, const REAL* y, int n, REAL* tmp /* 3n */) {
    REAL rho = FN(inverse_curvature)(s, y, n);
    REAL* yH = tmp;
    REAL* Hy = tmp + n;
    REAL* sr = tmp + 2 * n;
    for (int j = 0; j < n; ++j) {
        REAL a = 0;
        for (int i = 0; i < n; ++i) a += y[i] * H[i * n + j]; /* :268-270 */
        yH[j] = a;
    }
    REAL q = 0;
    for (int j = 0; j < n; ++j) q += yH[j] * (y[j] * rho); /* :271-274 */
    for (int i = 0; i < n; ++i) {
        REAL a = 0;
        for (int j = 0; j < n; ++j) a += H[i * n + j] * y[j]; /* :293-295 */
        Hy[i] = a;
        sr[i] = s[i] * rho; /* :277 */
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            REAL sop = (sr[i] * s[j]) * (R(1.0) + q); /* :278-284 */
            REAL sgp = sr[i] * yH[j];                 /* :287-289 */
            REAL gsp = Hy[i] * sr[j];                 /* :296-298 */
            H[i * n + j] = H[i * n + j] + sop - sgp - gsp; /* :299-303 */
        }
}

/* utils/func_interpolate_alpha.py:15-33, one element.  `nonlinear` (may be NULL) receives the bisection mask. */
static REAL FN(interpolate_alpha)(REAL a1, REAL a2, REAL v1, REAL v2, int* nonlinear) {
    const REAL lo = a1 < a2 ? a1 : a2, hi = a1 < a2 ? a2 : a1; /* :15-16 */
    const REAL diff = v2 - v1;                                 /* :18 */
    const REAL inv_gradient = (a2 - a1) / diff;                /* :19 */
    REAL cand = a1 - v1 * inv_gradient;                        /* :20 */
    const int nl = (diff == R(0.0)) || (cand < lo + R(1e-3)) || (cand > hi - R(1e-3)); /* :23-29 */
    if (nl) cand = (a1 + a2) / R(2.0);                         /* :30-32 */
    if (nonlinear) *nonlinear = nl;
    return cand;
}

/* forward and the custom backward (:42-79), elementwise; any output may be NULL */
int FN(davo_oracle_interpolate_alpha)(long long k, const REAL* a1, const REAL* a2, const REAL* v1, const REAL* v2,
                                      REAL* out, const REAL* grad_out, REAL* g_a1, REAL* g_a2, REAL* g_v1,
                                      REAL* g_v2) {
    for (long long i = 0; i < k; ++i) {
        int nl;
        const REAL c = FN(interpolate_alpha)(a1[i], a2[i], v1[i], v2[i], &nl);
        if (out) out[i] = c;
        if (grad_out) {
            const REAL go = grad_out[i], diff = v2[i] - v1[i];
            const REAL inv_gradient = (a2[i] - a1[i]) / diff;
            const REAL one_on_diff = nl ? R(0.0) : R(1.0) / diff;                         /* :34-35 */
            if (g_a1) g_a1[i] = nl ? R(0.5) * go : one_on_diff * v2[i] * go;             /* :56-61 */
            if (g_a2) g_a2[i] = nl ? R(0.5) * go : R(-1.0) * one_on_diff * v1[i] * go;   /* :62-67 */
            if (g_v1) g_v1[i] = nl ? R(0.0) : R(-1.0) * v2[i] * inv_gradient * one_on_diff * go; /* :68-73 */
            if (g_v2) g_v2[i] = nl ? R(0.0) : v1[i] * inv_gradient * one_on_diff * go;   /* :74-79 */
        }
    }
    return DAVO_OK;
}

/* autograd_solvers/line_search/wolfe_conditions.py:23-239, one problem.
 * secant != 0: the zoom step (:128-131) is interpolate_alpha(lo, hi, phi'(lo), phi'(hi)) as in the older
 * solvers/line_search_strong_wolfe_conditions.py:147-155 (bisection when the interpolant is not finite); this
 * composition has no reference implementation: PARITY UNPINNED for the composition, pinned for interpolate_alpha.
 * Returns `upper_alpha` (:239).  g_at_hi (n values) receives the gradient at x + hi d when the
 * final hi is the last probe (valid flag), so the caller can check the reuse identity. */
static REAL FN(line_search)(const FN(problem)* p, const REAL* x, const REAL* d, REAL f0, const REAL* g,
                            REAL c1, REAL c2, int strong, int secant, int max_probes, int* probes, REAL* xt,
                            REAL* gt) {
    const int n = p->n;
    REAL g0 = 0;
    for (int j = 0; j < n; ++j) g0 += d[j] * g[j]; /* :77 */
    int widening = 1, zooming = 0;                /* :80-82 */
    REAL lo = 0, hi = 0, cand = 1;                /* :97-108 */
    REAL lo_f = f0, hi_f = f0, cand_f = f0;       /* :109-111 */
    REAL lo_d = g0, hi_d = g0, cand_d = g0;       /* secant zoom: phi' at the bracket ends / last probe */
    for (int i = 0; i < max_probes; ++i) {        /* :116 */
        if (!(widening || zooming)) break;        /* :119-121 */
        if (i > 0) {
            if (widening) { hi = cand; hi_f = cand_f; hi_d = cand_d; cand = R(2.0) * cand; } /* :125-127 */
            if (zooming) {
                cand = R(0.5) * (lo + hi);                                     /* :128-131, :242-253 */
                if (secant) {
                    const REAL c = FN(interpolate_alpha)(lo, hi, lo_d, hi_d, NULL);
                    if (isfinite((double)c)) cand = c;
                }
            }
        }
        for (int j = 0; j < n; ++j) xt[j] = x[j] + cand * d[j]; /* :139 */
        cand_f = FN(objective)(p, xt, gt);                      /* :134-143 */
        REAL dphi = 0;
        for (int j = 0; j < n; ++j) dphi += d[j] * gt[j];
        cand_d = dphi;
        ++*probes;
        int D = cand_f > f0 + c1 * cand * g0;                 /* :146-150 */
        if (zooming) D = D || (cand_f >= lo_f);               /* :151-153 */
        if (widening && i > 0) D = D || (cand_f >= hi_f);     /* :154-157 */
        davo_oracle_stat[0] += 1; davo_oracle_stat[1] += D;   /* unsynchronised statistics, single-thread use */
        int C;
        if (strong) C = (dphi < 0 ? -dphi : dphi) <= R(-1.0) * c2 * g0; /* :160-164 */
        else        C = R(-1.0) * dphi <= R(-1.0) * c2 * g0;            /* :165-169 */
        int G = widening ? (dphi >= R(0.0)) : (dphi * (hi - lo) >= R(0.0)); /* :174-180 */
        if (zooming) {                                        /* :187-207 */
            if (D) { hi = cand; hi_f = cand_f; hi_d = dphi; }
            else if (C) { hi = lo = cand; hi_f = lo_f = cand_f; hi_d = lo_d = dphi; zooming = 0; }
            else {
                if (G) { hi = lo; hi_f = lo_f; hi_d = lo_d; }
                lo = cand; lo_f = cand_f; lo_d = dphi;
            }
        } else if (widening) {                                /* :216-237 */
            if (D) { lo = hi; lo_f = hi_f; lo_d = hi_d; hi = cand; hi_f = cand_f; hi_d = dphi; widening = 0; zooming = 1; }
            else if (C) { hi = lo = cand; hi_f = lo_f = cand_f; hi_d = lo_d = dphi; widening = 0; }
            else if (G) { lo = cand; lo_f = cand_f; lo_d = dphi; widening = 0; zooming = 1; }
        }
        if (zooming && !(lo != hi)) zooming = 0;              /* :236 */
    }
    return hi;
}

typedef struct {
    REAL cost;
    int iters, fevals, reason, converged;
} FN(result);

/* autograd_solvers/bfgs_solver.py:80-215, eval mode, one problem.  scratch: n*n + 10n REALs. */
static FN(result) FN(solve_one)(const FN(problem)* p, REAL* x, const davo_problem_desc* d, REAL* scratch) {
    const int n = p->n;
    REAL* H = scratch;
    REAL* g = H + n * n;
    REAL* gprev = g + n;
    REAL* dir = gprev + n;
    REAL* s = dir + n;
    REAL* y = s + n;
    REAL* xt = y + n;
    REAL* gt = xt + n;
    REAL* tmp = gt + n; /* 3n */
    const REAL thr = R(d->error_threshold), min_step = R(d->minimum_step);
    const REAL c1 = R(d->sufficient_decrease), c2 = R(d->curvature);
    for (int i = 0; i < n * n; ++i) H[i] = 0;
    for (int i = 0; i < n; ++i) { H[i * n + i] = 1; s[i] = 0; g[i] = 0; }
    FN(result) r = {0, 0, 0, DAVO_REASON_CAP, 0};
    REAL f = 0;
    int have_f_at_x = 0;
    for (int k = 0; k < d->max_iters; ++k) {                /* :118 */
        for (int j = 0; j < n; ++j) gprev[j] = g[j];        /* :119 */
        f = FN(objective)(p, x, g);                         /* :128-135 */
        r.fevals++;
        have_f_at_x = 1;
        if (!(f > thr)) {                                   /* :143 */
            r.reason = (f <= thr) ? DAVO_REASON_THRESHOLD : DAVO_REASON_NAN;
            break;
        }
        if (k == 0) {
            for (int j = 0; j < n; ++j) dir[j] = R(-1.0) * g[j]; /* :152-155 */
        } else {
            for (int j = 0; j < n; ++j) y[j] = g[j] - gprev[j];  /* :157 */
            if (k == 1) {                                        /* :159-167 */
                REAL sc = FN(initial_scale)(s, y, n);
                for (int i = 0; i < n * n; ++i) H[i] = sc * H[i];
            }
            FN(update_inverse_hessian)(H, s, y, n, tmp);         /* :168-172 */
            for (int i = 0; i < n; ++i) {                        /* :173-176 */
                REAL a = 0;
                for (int j = 0; j < n; ++j) a += H[i * n + j] * g[j];
                dir[i] = R(-1.0) * a;
            }
        }
        int probes = 0;
        REAL alpha = FN(line_search)(p, x, dir, f, g, c1, c2, d->strong, d->zoom_interpolation, d->max_ls_iters, &probes, xt, gt); /* :181-190 */
        r.fevals += probes;
        r.iters++;
        REAL nrm = 0;
        for (int j = 0; j < n; ++j) {                        /* :191-199 */
            s[j] = alpha * dir[j];
            x[j] = x[j] + s[j];
            nrm += s[j] * s[j];
        }
        have_f_at_x = 0;
        nrm = R(sqrt((double)nrm));
        if (!(nrm > min_step)) { r.reason = DAVO_REASON_STEP; break; } /* :203-207 */
    }
    if (!have_f_at_x) f = FN(objective)(p, x, NULL); /* what calibration_network.py:71 re-evaluates */
    r.cost = f;
    r.converged = (f <= thr);
    return r;
}

static void FN(bind)(FN(problem)* p, const davo_problem_desc* d, const REAL* data0, const REAL* data1,
                     const REAL* w, int b) {
    p->model = d->model; p->n = d->n; p->N = d->N; p->V = d->V; p->has_weights = d->has_weights;
    p->data0 = NULL; p->data1 = NULL;
    p->w = (d->has_weights && w) ? w + (size_t)b * d->V * d->N : NULL;
    if (d->model == DAVO_MODEL_DISTORT10) p->data0 = data0 + (size_t)b * d->N * 4;
    else if (d->model == DAVO_MODEL_DISTANCE) p->data0 = data0 + (size_t)b * d->n;
    else if (d->model == DAVO_MODEL_ANGLE_BA) p->data0 = data0 + (size_t)b * d->V * d->N * 2;
    else if (d->model == DAVO_MODEL_JOINT) {
        p->data0 = data0 + (size_t)b * d->N * 3;
        p->data1 = data1 + (size_t)b * d->V * d->N * 2;
    }
}

static int FN(check_desc)(const davo_problem_desc* d) {
    if (!d) return DAVO_ERR_NULL_POINTER;
    if (d->B < 0 || d->n < 1) return DAVO_ERR_BAD_SHAPE;
    if (d->model == DAVO_MODEL_DISTORT10 && (d->n != 10 || d->V != 1)) return DAVO_ERR_BAD_SHAPE;
    if (d->model == DAVO_MODEL_JOINT && (d->V < 1 || d->n != 10 + 6 * d->V)) return DAVO_ERR_BAD_SHAPE;
    if (d->model == DAVO_MODEL_ROSENBROCK && d->n != 2) return DAVO_ERR_BAD_SHAPE;
    if (d->model == DAVO_MODEL_ANGLE_BA && (d->V < 2 || d->N < 1 || d->n != 3 + 3 * d->N + 6 * (d->V - 1)))
        return DAVO_ERR_BAD_SHAPE;
    return DAVO_OK;
}

int FN(davo_oracle_solve)(const davo_problem_desc* d, const REAL* data0, const REAL* data1, const REAL* w,
                          const REAL* x0, REAL* x_out, REAL* cost_out, uint8_t* converged_out,
                          int32_t* iters_out, int32_t* fevals_out, int32_t* reason_out, int nthreads) {
    int st = FN(check_desc)(d);
    if (st) return st;
    const int n = d->n;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#endif
#pragma omp parallel num_threads(nthreads)
    {
        REAL* scratch = (REAL*)malloc(sizeof(REAL) * ((size_t)n * n + 12 * (size_t)n));
#pragma omp for schedule(dynamic, 4)
        for (int b = 0; b < d->B; ++b) {
            FN(problem) p;
            FN(bind)(&p, d, data0, data1, w, b);
            REAL* x = x_out + (size_t)b * n;
            if (x != x0 + (size_t)b * n)
                for (int j = 0; j < n; ++j) x[j] = x0[(size_t)b * n + j];
            FN(result) r = FN(solve_one)(&p, x, d, scratch);
            if (cost_out) cost_out[b] = r.cost;
            if (converged_out) converged_out[b] = (uint8_t)r.converged;
            if (iters_out) iters_out[b] = r.iters;
            if (fevals_out) fevals_out[b] = r.fevals;
            if (reason_out) reason_out[b] = r.reason;
        }
        free(scratch);
    }
    return DAVO_OK;
}

int FN(davo_oracle_eval)(const davo_problem_desc* d, const REAL* data0, const REAL* data1, const REAL* w,
                         const REAL* x, REAL* cost, REAL* grad) {
    int st = FN(check_desc)(d);
    if (st) return st;
    for (int b = 0; b < d->B; ++b) {
        FN(problem) p;
        FN(bind)(&p, d, data0, data1, w, b);
        REAL f = FN(objective)(&p, x + (size_t)b * d->n, grad ? grad + (size_t)b * d->n : NULL);
        if (cost) cost[b] = f;
    }
    return DAVO_OK;
}

int FN(davo_oracle_line_search)(const davo_problem_desc* d, const REAL* data0, const REAL* data1,
                                const REAL* w, const REAL* x, const REAL* dir, const REAL* f0,
                                const REAL* g, REAL* alpha_out, int32_t* fevals_out) {
    int st = FN(check_desc)(d);
    if (st) return st;
    const int n = d->n;
    REAL* tmp = (REAL*)malloc(sizeof(REAL) * 2 * (size_t)n);
    for (int b = 0; b < d->B; ++b) {
        FN(problem) p;
        FN(bind)(&p, d, data0, data1, w, b);
        int probes = 0;
        alpha_out[b] = FN(line_search)(&p, x + (size_t)b * n, dir + (size_t)b * n, f0[b], g + (size_t)b * n,
                                       R(d->sufficient_decrease), R(d->curvature), d->strong,
                                       d->zoom_interpolation, d->max_ls_iters, &probes, tmp, tmp + n);
        if (fevals_out) fevals_out[b] = probes;
    }
    free(tmp);
    return DAVO_OK;
}

int FN(davo_oracle_bfgs_update)(int k, int n, REAL* H, const REAL* s, const REAL* y) {
    REAL* tmp = (REAL*)malloc(sizeof(REAL) * 3 * (size_t)n);
    for (int b = 0; b < k; ++b)
        FN(update_inverse_hessian)(H + (size_t)b * n * n, s + (size_t)b * n, y + (size_t)b * n, n, tmp);
    free(tmp);
    return DAVO_OK;
}

int FN(davo_oracle_bfgs_initial_scale)(int k, int n, const REAL* s, const REAL* y, REAL* scale) {
    for (int b = 0; b < k; ++b) scale[b] = FN(initial_scale)(s + (size_t)b * n, y + (size_t)b * n, n);
    return DAVO_OK;
}

/* compute_distorted_camera_model (distorted_camera_model.py:24-111): full 16-parameter forward. */
int FN(davo_oracle_project)(int B, int N, const REAL* pts, const REAL* th16, REAL* u_out, REAL* v_out,
                            REAL* J /* [B,2N,16] or NULL */) {
    for (int b = 0; b < B; ++b) {
        const REAL* th = th16 + 16 * (size_t)b;
        REAL Rm[9], dRx[9], dRy[9], dRz[9];
        FN(euler)(th[DAVO_RX], th[DAVO_RY], th[DAVO_RZ], Rm, dRx, dRy, dRz);
        const REAL cx = th[0], cy = th[1], k1 = th[2], k2 = th[3], k3 = th[4], p1 = th[5], p2 = th[6],
                   fx = th[7], s = th[8], fy = th[9];
        for (int m = 0; m < N; ++m) {
            const REAL* X = pts + 3 * ((size_t)b * N + m);
            REAL Xp[3];
            FN(transform)(Rm, th + DAVO_TX, X, Xp);
            REAL a = Xp[0] / Xp[2], bb = Xp[1] / Xp[2];
            REAL u = fx * a + s * bb, v = fy * bb;
            REAL r2 = u * u + v * v, uv = u * v;
            REAL rad = R(1.0) + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2;
            REAL up = u * rad + R(2.0) * p1 * uv + p2 * (r2 + R(2.0) * u * u) + cx;
            REAL vp = v * rad + R(2.0) * p2 * uv + p1 * (r2 + R(2.0) * v * v) + cy;
            u_out[(size_t)b * N + m] = up;
            v_out[(size_t)b * N + m] = vp;
            if (J) {
                REAL r4 = r2 * r2, r6 = r4 * r2;
                REAL radp = k1 + R(2.0) * k2 * r2 + R(3.0) * k3 * r4;
                REAL Duu = rad + R(2.0) * u * u * radp + R(2.0) * p1 * v + R(6.0) * p2 * u;
                REAL Dvv = rad + R(2.0) * v * v * radp + R(6.0) * p1 * v + R(2.0) * p2 * u;
                REAL Duv = R(2.0) * uv * radp + R(2.0) * p1 * u + R(2.0) * p2 * v;
                REAL iz = R(1.0) / Xp[2];
                REAL* Ju = J + 16 * ((size_t)b * 2 * N + m);
                REAL* Jv = J + 16 * ((size_t)b * 2 * N + N + m);
                Ju[0] = 1; Jv[0] = 0;
                Ju[1] = 0; Jv[1] = 1;
                Ju[2] = u * r2; Jv[2] = v * r2;
                Ju[3] = u * r4; Jv[3] = v * r4;
                Ju[4] = u * r6; Jv[4] = v * r6;
                Ju[5] = R(2.0) * uv; Jv[5] = r2 + R(2.0) * v * v;
                Ju[6] = r2 + R(2.0) * u * u; Jv[6] = R(2.0) * uv;
                Ju[7] = Duu * a;  Jv[7] = Duv * a;
                Ju[8] = Duu * bb; Jv[8] = Duv * bb;
                Ju[9] = Duv * bb; Jv[9] = Dvv * bb;
                /* d(u',v')/dX' */
                REAL ux = Duu * fx * iz, uy = (Duu * s + Duv * fy) * iz, uz = -(Duu * u + Duv * v) * iz;
                REAL vx = Duv * fx * iz, vy = (Duv * s + Dvv * fy) * iz, vz = -(Duv * u + Dvv * v) * iz;
                const REAL* dR[3] = {dRx, dRy, dRz};
                for (int k = 0; k < 3; ++k) {
                    REAL dX = dR[k][0] * X[0] + dR[k][1] * X[1] + dR[k][2] * X[2];
                    REAL dY = dR[k][3] * X[0] + dR[k][4] * X[1] + dR[k][5] * X[2];
                    REAL dZ = dR[k][6] * X[0] + dR[k][7] * X[1] + dR[k][8] * X[2];
                    Ju[10 + k] = ux * dX + uy * dY + uz * dZ;
                    Jv[10 + k] = vx * dX + vy * dY + vz * dZ;
                }
                Ju[13] = ux; Ju[14] = uy; Ju[15] = uz;
                Jv[13] = vx; Jv[14] = vy; Jv[15] = vz;
            }
        }
    }
    return DAVO_OK;
}

/* Staging for DISTORT10: {a, b, u*, v*} per match (last paragraph). */
int FN(davo_oracle_stage)(int B, int N, const REAL* pts, const REAL* obs, const REAL* pose, REAL* staged) {
    for (int b = 0; b < B; ++b) {
        REAL Rm[9];
        REAL t[3] = {0, 0, 0};
        if (pose) {
            const REAL* ps = pose + 6 * (size_t)b;
            FN(euler)(ps[0], ps[1], ps[2], Rm, NULL, NULL, NULL);
            t[0] = ps[3]; t[1] = ps[4]; t[2] = ps[5];
        } else {
            FN(euler)(0, 0, 0, Rm, NULL, NULL, NULL);
        }
        for (int m = 0; m < N; ++m) {
            size_t i = (size_t)b * N + m;
            REAL Xp[3];
            FN(transform)(Rm, t, pts + 3 * i, Xp);
            staged[4 * i + 0] = Xp[0] / Xp[2];
            staged[4 * i + 1] = Xp[1] / Xp[2];
            staged[4 * i + 2] = obs[2 * i + 0];
            staged[4 * i + 3] = obs[2 * i + 1];
        }
    }
    return DAVO_OK;
}

/* solvers/least_squares_utils.py:16-48 on explicit residuals / jacobian. */
int FN(davo_oracle_least_squares)(int B, int Rn, int P, const REAL* res, const REAL* jac, const REAL* w,
                                  REAL* err, REAL* grad) {
    for (int b = 0; b < B; ++b) {
        REAL e = 0;
        if (grad) for (int p = 0; p < P; ++p) grad[(size_t)b * P + p] = 0;
        for (int r = 0; r < Rn; ++r) {
            size_t i = (size_t)b * Rn + r;
            REAL ww = w ? w[i] : R(1.0);
            REAL sq = res[i] * res[i];
            e += w ? ww * sq : sq;
            if (grad && jac) {
                REAL gr = R(2.0) * res[i];
                if (w) gr = ww * gr;
                for (int p = 0; p < P; ++p) grad[(size_t)b * P + p] += gr * jac[i * P + p];
            }
        }
        if (err) err[b] = e;
    }
    return DAVO_OK;
}

#undef CAT_
#undef CAT
#undef FN
#undef R
