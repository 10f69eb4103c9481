/* ORACLE (test infrastructure, NOT product code): plain-C restatement of the reference's GCFM path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * the library built from this file.  The product (optimal_crowds_b200/) never links or calls it.
 *
 * Follows, line by line:
 *   - step sweep         : /root/reference/simulations.py:271-339
 *   - pair force         : /root/reference/pedestrians.py:216-280   (agents_repulsion)
 *   - cutoff distance    : /root/reference/pedestrians.py:336-356   (distance)
 *   - wall force         : /root/reference/pedestrians.py:282-334   (wall_repulsion)
 *   - door test          : /root/reference/pedestrians.py:121-136   (check_status)
 *   - field sampler      : /root/reference/optimals.py:212-250      (choose_optimal_velocity)
 *   - density            : /root/reference/simulations.py:453-487   (gaussian_density)
 *   - room rasteriser    : /root/reference/simulations.py:516-576   (create_potential)
 *
 * Pinning ("two-oracle protocol"): the reference's GCFM update is chaotic at
 * round-off level and numpy's exp/arctan2 are CPU-feature dependent, so the unmodified reference
 * (O1) is not bit-reproducible even against itself on another CPU.  This file is O2: the same
 * formulae in the same evaluation order with fully specified arithmetic (IEEE + - * / sqrt fma and
 * the elementary functions of optimal_crowds_b200/csrc/oc_math.h -- the ONLY product header
 * included here).  O2 is checked against O1 teacher-forced (one step from identical state,
 * <= 1e-12, identical exits) on golden vectors produced by running the reference in the build
 * container (tests/golden/gcfm_*.npz, oracle/make_goldens.py).  The CUDA path must match O2 bit for
 * bit over whole runs.
 *
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../optimal_crowds_b200/csrc/oc_math.h"

/* vectorised wrappers so tests can pin the elementary functions against numpy */
void oco_math_exp(const double *x, double *o, int n) { for (int i = 0; i < n; i++) o[i] = ocm_exp(x[i]); }
void oco_math_sin(const double *x, double *o, int n) { for (int i = 0; i < n; i++) o[i] = ocm_sin(x[i]); }
void oco_math_cos(const double *x, double *o, int n) { for (int i = 0; i < n; i++) o[i] = ocm_cos(x[i]); }
void oco_math_atan2(const double *y, const double *x, double *o, int n) {
    for (int i = 0; i < n; i++) o[i] = ocm_atan2(y[i], x[i]);
}

/* shared-work variants used by the CUDA kernels: must return the bits of the separate calls */
void oco_math_sincos(const double *x, double *s, double *c, int n) { for (int i = 0; i < n; i++) ocm_sincos(x[i], &s[i], &c[i]); }
void oco_math_atan2_both(const double *y, const double *x, double *p, double *q, int n) {
    for (int i = 0; i < n; i++) ocm_atan2_both(y[i], x[i], &p[i], &q[i]);
}

/* Parameters computed on the host by the SAME Python expressions the reference evaluates
 * (e.g. dt**2, noise_intensity/2, np.cos(0.7*np.pi)), so constants are bit-identical. */
typedef struct {
    double dt, dt2;            /* dt, dt**2                            simulations.py:314 */
    double half_noise;         /* noise_intensity/2                    simulations.py:303 */
    double relaxation, v_max;  /*                                       simulations.py:307,323 */
    double cutoff;             /* repulsion_cutoff                     simulations.py:291 */
    double a_min, tau_a, b_min, b_max, eta, eta_walls; /* pedestrians.py:73-79 (a_min == b_min) */
    double cos_fov, one_minus_cos_fov; /* np.cos(0.7*np.pi), 1 - that  pedestrians.py:262 */
    double dx, dy;             /* grid_step                            optimals.py:58-59 */
    double room_length, room_height;
    int Ny, Nx;
} oco_gcfm_params;

/* one HJB key (target set): potential, field slices, doors */
typedef struct {
    const double *V;            /* (Ny,Nx) remapped {-100,0,1}          optimals.py:89-91 */
    const double *vx_opt, *vy_opt; /* (nt_alloc-1, Ny-2, Nx-2)          optimals.py:80-81 */
    int nt_opt;                 /* current nt_opt                        optimals.py:140 */
    int n_slices;               /* allocated slices (bounds check only) */
    const double *doors;        /* (n_doors,4) cx,cy,w,h                 pedestrians.py:132-135 */
    int n_doors;
} oco_key;

/* CPython/numpy float floor-division (fmod based; Objects/floatobject.c float_divmod, numpy npy_divmod) */
static double py_floordiv(double vx, double wx) {
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    if (mod != 0.0) {
        if ((wx < 0) != (mod < 0)) div -= 1.0;
    }
    if (div != 0.0) {
        double fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
        return fl;
    }
    return copysign(0.0, vx / wx);
}

/* optimals.py:212-250.  Returns 0 ok, 1 if the reference would raise IndexError (App. C #7). */
int oco_choose_velocity(const oco_gcfm_params *p, const oco_key *k, double x, double y, int t,
                        double *ox, double *oy) {
    if (t >= k->nt_opt - 1) { *ox = 0.0; *oy = 0.0; return 0; }
    long j0, j1, i0, i1;
    if (x < p->room_length - p->dx) {
        j0 = (long)py_floordiv(x, p->dx);
        j1 = (x > p->dx) ? j0 + 1 : j0;
    } else { j0 = j1 = p->Nx - 3; }
    if (y < p->room_height - p->dy) {
        i0 = (long)py_floordiv(y, p->dy);
        i1 = (y > p->dy) ? i0 + 1 : i0;
    } else { i0 = i1 = p->Ny - 3; }
    int W = p->Nx - 2, H = p->Ny - 2;
    /* numpy wraps negative indices (x < 0 or y < 0; then the index list has one element) */
    if (j0 < 0) { j0 += W; j1 = j0; }
    if (i0 < 0) { i0 += H; i1 = i0; }
    if (j0 < 0 || j1 >= W || i0 < 0 || i1 >= H || t < 0 || t >= k->n_slices) { *ox = 0; *oy = 0; return 1; }
    const double *sx = k->vx_opt + (size_t)t * W * H, *sy = k->vy_opt + (size_t)t * W * H;
    if (j0 == j1 && i0 == i1) { /* scalar: np.mean of a 0-d value */
        *ox = sx[i0 * W + j0];
        *oy = sy[i0 * W + j0];
    } else { /* two selected nodes (i0,j0),(i1,j1): lists pair up element-wise => diagonal */
        *ox = (sx[i0 * W + j0] + sx[i1 * W + j1]) / 2.0;
        *oy = (sy[i0 * W + j0] + sy[i1 * W + j1]) / 2.0;
    }
    return 0;
}

/* pedestrians.py:237-280; i = self, j = other (current state) */
static void pair_force(const oco_gcfm_params *p, double xi, double yi, double vxi, double vyi,
                       double v_des_i, double xj, double yj, double vxj, double vyj, double *fx,
                       double *fy) {
    double ni = ocm_norm2(vxi, vyi), nj = ocm_norm2(vxj, vyj);
    double a_i = p->a_min + p->tau_a * ni;
    double b_i = p->b_max - (p->b_max - p->b_min) * ocm_npmin(ni / v_des_i, 1.0);
    double a_j = p->a_min + p->tau_a * nj;
    double b_j = p->b_max - (p->b_max - p->b_min) * ocm_npmin(nj / v_des_i, 1.0);
    double Rx = xj - xi, Ry = yj - yi;
    double nR = ocm_norm2(Rx, Ry);
    double ex = Rx / nR, ey = Ry / nR;
    double wx = vxj - vxi, wy = vyj - vyi;
    double d = wx * (-ex) + wy * (-ey);
    double v_rel = 0.5 * (d + fabs(d));
    double k = 0.0;
    if (ni > 0) k = ocm_npmax((vxi * ex + vyi * ey) / ni - p->cos_fov, 0.0) / p->one_minus_cos_fov;
    double alpha_i = ocm_atan2(Ry, Rx), beta_i = ocm_atan2(vyi, vxi);
    double alpha_j = ocm_atan2(-Ry, -Rx), beta_j = ocm_atan2(vyj, vxj);
    double ci = ocm_cos(alpha_i - beta_i) / a_i, si = ocm_sin(alpha_i - beta_i) / b_i;
    double q_i = sqrt(1.0 / (ci * ci + si * si));
    double cj = ocm_cos(alpha_j - beta_j) / a_j, sj = ocm_sin(alpha_j - beta_j) / b_j;
    double q_j = sqrt(1.0 / (cj * cj + sj * sj));
    double dist = nR - q_i - q_j;
    double rep = ocm_npmin(k * ocm_exp(-dist / (p->eta * (1.0 + v_rel))), 1.0);
    *fx = -rep * Rx;
    *fy = -rep * Ry;
}

/* pedestrians.py:311-313: first flat index minimising sqrt((X-x)^2+(Y-y)^2) + V*10e3 over the WHOLE grid.
 * X[j], Y[i] are np.linspace coordinates supplied by the caller. */
long oco_wall_argmin(const double *X, const double *Y, const double *V, int Ny, int Nx, double x, double y) {
    double best = INFINITY;
    long bi = 0;
    for (int i = 0; i < Ny; i++) {
        double dy = Y[i] - y, dy2 = dy * dy;
        for (int j = 0; j < Nx; j++) {
            double dx = X[j] - x;
            double key = sqrt(dx * dx + dy2) + V[(size_t)i * Nx + j] * 10e3;
            if (key < best) { best = key; bi = (long)i * Nx + j; }
        }
    }
    return bi;
}

/* pedestrians.py:315-334 given the nearest wall node */
static void wall_force(const oco_gcfm_params *p, double xi, double yi, double vxi, double vyi,
                       double v_des_i, double wxp, double wyp, double *fx, double *fy) {
    double Rx = wxp - xi, Ry = wyp - yi;
    double nR = ocm_norm2(Rx, Ry);
    double ex = Rx / nR, ey = Ry / nR;
    double d = vxi * ex + vyi * ey;
    double v_rel = 0.5 * (d + fabs(d));
    double alpha = ocm_atan2(Ry, Rx), beta = ocm_atan2(vyi, vxi);
    double ni = ocm_norm2(vxi, vyi);
    double a_i = p->a_min + p->tau_a * ni;
    double b_i = p->b_max - (p->b_max - p->b_min) * ocm_npmin(ni / v_des_i, 1.0);
    double c = ocm_cos(alpha - beta) / a_i, s = ocm_sin(alpha - beta) / b_i;
    double q_i = 1.0 / (c * c + s * s); /* no sqrt: pedestrians.py:328 */
    double dist = nR - q_i;
    double rep = ocm_npmin(ocm_exp(-dist / (p->eta_walls * (1.0 + v_rel))), 1.0);
    *fx = -3.0 * rep * Rx;
    *fy = -3.0 * rep * Ry;
}

void oco_wall_force(const oco_gcfm_params *p, const double *X, const double *Y, const double *V, double xi,
                    double yi, double vxi, double vyi, double v_des_i, double *fx, double *fy, long *ind) {
    long k = oco_wall_argmin(X, Y, V, p->Ny, p->Nx, xi, yi);
    if (ind) *ind = k;
    wall_force(p, xi, yi, vxi, vyi, v_des_i, X[k % p->Nx], Y[k / p->Nx], fx, fy);
}

void oco_pair_force(const oco_gcfm_params *p, double xi, double yi, double vxi, double vyi, double v_des_i,
                    double xj, double yj, double vxj, double vyj, double *fx, double *fy) {
    pair_force(p, xi, yi, vxi, vyi, v_des_i, xj, yj, vxj, vyj, fx, fy);
}

/* One call of simulation.step (simulations.py:252-339), sequential in-place sweep.
 *  state (in/out): x,y,vx,vy,time (N doubles each), status (N uint8: 1 inside / 0 exited)
 *  perm : this step's np.random.choice(arange(N),N,replace=False)           (:271)
 *  noise: (n_active_at_start,2) standard normals in consumption order       (:303)
 *  exit_log (cap N): agent ids in the order they left this step; returns their count via *n_exit
 *  wall_ind (nullable, N): flat argmin index per processed agent (diagnostic)
 *  returns 0, or 1 if any sampler access would have been out of range in the reference */
int oco_gcfm_step(const oco_gcfm_params *p, int N, double *x, double *y, double *vx, double *vy,
                  double *tim, uint8_t *status, const double *v_des, const int *key_id, const oco_key *keys,
                  const double *X, const double *Y, const int *perm, const double *noise, int simu_step,
                  int *exit_log, int *n_exit, long *wall_ind) {
    int nz = 0, ne = 0, bad = 0;
    for (int r = 0; r < N; r++) {
        int i = perm[r];
        if (!status[i]) continue;
        const oco_key *K = &keys[key_id[i]];
        double ux, uy;
        bad |= oco_choose_velocity(p, K, x[i], y[i], simu_step, &ux, &uy);
        double des_x = v_des[i] * ux, des_y = v_des[i] * uy;
        double rx = 0.0, ry = 0.0;
        for (int j = 0; j < N; j++) {
            if (!status[j] || j == i) continue;
            double ddx = x[j] - x[i], ddy = y[j] - y[i];
            if (sqrt(ddx * ddx + ddy * ddy) < p->cutoff) { /* pedestrians.py:354 */
                double fx, fy;
                pair_force(p, x[i], y[i], vx[i], vy[i], v_des[i], x[j], y[j], vx[j], vy[j], &fx, &fy);
                rx = rx + fx;
                ry = ry + fy;
            }
        }
        double wfx, wfy;
        long wi;
        oco_wall_force(p, X, Y, K->V, x[i], y[i], vx[i], vy[i], v_des[i], &wfx, &wfy, &wi);
        if (wall_ind) wall_ind[i] = wi;
        double cx = vx[i] + p->half_noise * noise[2 * nz] + rx + wfx;      /* :303 */
        double cy = vy[i] + p->half_noise * noise[2 * nz + 1] + ry + wfy;
        nz++;
        double ax = (des_x - cx) / p->relaxation, ay = (des_y - cy) / p->relaxation; /* :307-308 */
        double nx_ = x[i] + cx * p->dt + 0.5 * ax * p->dt2;                 /* :314-315 */
        double ny_ = y[i] + cy * p->dt + 0.5 * ay * p->dt2;
        double nvx = cx + ax * p->dt, nvy = cy + ay * p->dt;                /* :316-317 */
        double nr = sqrt(nvx * nvx + nvy * nvy);                            /* :321 */
        if (!(nr < p->v_max)) {                                             /* :323-326 */
            double sc = p->v_max / nr;
            nvx = nvx * sc;
            nvy = nvy * sc;
        }
        x[i] = nx_; y[i] = ny_; vx[i] = nvx; vy[i] = nvy;
        tim[i] += p->dt;                                                    /* pedestrians.py:191 */
        for (int d = 0; d < K->n_doors; d++) {                              /* pedestrians.py:132-136 */
            const double *door = K->doors + 4 * d;
            if (fabs(nx_ - door[0]) < door[2] * 0.5 && fabs(ny_ - door[1]) < door[3] * 0.5) status[i] = 0;
        }
        if (!status[i]) exit_log[ne++] = i;                                 /* simulations.py:331-332 */
    }
    *n_exit = ne;
    return bad;
}

/* simulations.py:469-487 */
void oco_density(const double *X, const double *Y, const double *Vglobal, int Ny, int Nx, int N,
                 const double *x, const double *y, const uint8_t *status, double sigma, double C, double *d) {
    size_t n = (size_t)Ny * Nx;
    for (size_t c = 0; c < n; c++) d[c] = 0.0;
    double two_s2 = 2 * (sigma * sigma); /* 2*sigma**2 */
    for (int a = 0; a < N; a++) {
        if (!status[a]) continue;
        for (int i = 0; i < Ny; i++) {
            double cy = Y[i] - y[a], cy2 = cy * cy;
            for (int j = 0; j < Nx; j++) {
                double cx = X[j] - x[a];
                d[(size_t)i * Nx + j] += ocm_exp(-(cx * cx + cy2) / two_s2) / C;
            }
        }
    }
    for (size_t c = 0; c < n; c++)
        if (Vglobal[c] < 0) d[c] = 0.0;
}

/* simulations.py:516-576.  shapes are flat (n,4) / (n,3) arrays in JSON order. */
void oco_create_potential(const double *X, const double *Y, int Ny, int Nx, const double *walls, int n_walls,
                          const double *holes, int n_holes, const double *cyls, int n_cyls,
                          const double *targets, int n_targets, double *V) {
    for (int i = 0; i < Ny; i++)
        for (int j = 0; j < Nx; j++) {
            double v = 0.0;
            for (int w = 0; w < n_walls; w++) {
                const double *s = walls + 4 * w;
                if (fabs(X[j] - s[0]) < s[2] / 2 && fabs(Y[i] - s[1]) < s[3] / 2) v += -1.0;
            }
            for (int w = 0; w < n_holes; w++) {
                const double *s = holes + 4 * w;
                if (fabs(X[j] - s[0]) < s[2] / 2 && fabs(Y[i] - s[1]) < s[3] / 2) v = 0.0;
            }
            for (int w = 0; w < n_cyls; w++) {
                const double *s = cyls + 3 * w;
                double ddx = X[j] - s[0], ddy = Y[i] - s[1];
                if (sqrt(ddx * ddx + ddy * ddy) < s[2]) v += -1.0;
            }
            if (j == 0 || j == Nx - 1 || i == 0 || i == Ny - 1) v = -1.0;
            for (int w = 0; w < n_targets; w++) {
                const double *s = targets + 4 * w;
                if (fabs(X[j] - s[0]) < s[2] / 2 && fabs(Y[i] - s[1]) < s[3] / 2) v = 1.0;
            }
            V[(size_t)i * Nx + j] = v;
        }
}
