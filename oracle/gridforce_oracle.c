/* TEST INFRASTRUCTURE — see gridforce_oracle.h. Plain C99 restatement of
 * /root/reference/platforms/reference/src/ReferenceGridForceKernels.cpp:646-1121 (trilinear branch, and the cubic
 * B-spline branch :727-795 when gfo_grid.interp_method == 1, the tricubic Hermite branch :796-893 when it is 2).
 * Build with -ffp-contract=off so a*b+c stays two roundings, as in the reference build. */
#include "gridforce_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* Cubic B-spline basis and derivative (:54-63), same expressions. */
static double bs_b0(double t) { return (1.0 - t) * (1.0 - t) * (1.0 - t) / 6.0; }
static double bs_b1(double t) { return (3.0 * t * t * t - 6.0 * t * t + 4.0) / 6.0; }
static double bs_b2(double t) { return (-3.0 * t * t * t + 3.0 * t * t + 3.0 * t + 1.0) / 6.0; }
static double bs_b3(double t) { return t * t * t / 6.0; }
static double bs_d0(double t) { return -(1.0 - t) * (1.0 - t) / 2.0; }
static double bs_d1(double t) { return (3.0 * t * t - 4.0 * t) / 2.0; }
static double bs_d2(double t) { return (-3.0 * t * t + 2.0 * t + 1.0) / 2.0; }
static double bs_d3(double t) { return t * t / 2.0; }
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* interpolation method 1 (:727-795): 4x4x4 points ix-1..ix+2 with indices clamped into the grid, weights
 * bx[i]*by[j]*bz[k], summed in i, j, k order. Unlike the trilinear branch the clamping makes the upper face
 * (idx == n-1, fraction 0) well defined, so it is evaluated exactly as the reference does. */
static double gfo_interp_bspline(const gfo_grid* g, double scale, const double* pi, double* f, gfo_class* cls) {
    const int nz = g->counts[2];
    const int nyz = g->counts[1] * nz;
    int i, j, k, idx[3];
    double fr[3], b[3][4], d[3][4];
    double interpolated = 0.0, dvdx = 0.0, dvdy = 0.0, dvdz = 0.0, grd[3];
    for (k = 0; k < 3; k++) {
        idx[k] = (int)(pi[k] / g->spacing[k]);                   /* :708-710 */
        fr[k] = (pi[k] / g->spacing[k]) - idx[k];                /* :713-715 */
        b[k][0] = bs_b0(fr[k]); b[k][1] = bs_b1(fr[k]); b[k][2] = bs_b2(fr[k]); b[k][3] = bs_b3(fr[k]);   /* :741-743 */
        d[k][0] = bs_d0(fr[k]); d[k][1] = bs_d1(fr[k]); d[k][2] = bs_d2(fr[k]); d[k][3] = bs_d3(fr[k]);   /* :746-748 */
    }
    if (cls) { cls->cell[0] = idx[0]; cls->cell[1] = idx[1]; cls->cell[2] = idx[2]; }
    for (i = 0; i < 4; i++) {                                                                /* :754-775 */
        const int gx = clampi(idx[0] - 1 + i, 0, g->counts[0] - 1);
        for (j = 0; j < 4; j++) {
            const int gy = clampi(idx[1] - 1 + j, 0, g->counts[1] - 1);
            for (k = 0; k < 4; k++) {
                const int gz = clampi(idx[2] - 1 + k, 0, nz - 1);
                const double val = g->vals[gx * nyz + gy * nz + gz];
                const double weight = b[0][i] * b[1][j] * b[2][k];
                interpolated += weight * val;
                dvdx += d[0][i] * b[1][j] * b[2][k] * val;
                dvdy += b[0][i] * d[1][j] * b[2][k] * val;
                dvdz += b[0][i] * b[1][j] * d[2][k] * val;
            }
        }
    }
    if (g->inv_power > 0.0) {                                                                /* :778-787 */
        const double base = interpolated;
        const double pf = g->inv_power * pow(base, g->inv_power - 1.0);
        interpolated = pow(interpolated, g->inv_power);
        dvdx *= pf; dvdy *= pf; dvdz *= pf;
    }
    grd[0] = dvdx / g->spacing[0]; grd[1] = dvdy / g->spacing[1]; grd[2] = dvdz / g->spacing[2];   /* :790 */
    if (f) for (k = 0; k < 3; k++) f[k] -= scale * grd[k];                                   /* :794 */
    return scale * interpolated;                                                             /* :793 */
}

/* Cubic Hermite basis and derivative (:68-77), same expressions. */
static double he_h00(double t) { return (1.0 + 2.0 * t) * (1.0 - t) * (1.0 - t); }
static double he_h10(double t) { return t * (1.0 - t) * (1.0 - t); }
static double he_h01(double t) { return t * t * (3.0 - 2.0 * t); }
static double he_h11(double t) { return t * t * (t - 1.0); }
static double he_d00(double t) { return 6.0 * t * t - 6.0 * t; }
static double he_d10(double t) { return 3.0 * t * t - 4.0 * t + 1.0; }
static double he_d01(double t) { return -6.0 * t * t + 6.0 * t; }
static double he_d11(double t) { return 3.0 * t * t - 2.0 * t; }

/* A grid value by FLAT index, as the reference's tricubic branch forms them (:801-861). Neighbour indices past an axis
 * end are not clamped there: iy+2 == ny or iz+2 == nz lands in the next row / x-slab (deterministic, reproduced here),
 * and in the last x layer (ix == nx-2) ix+2 == nx lands past the end of the vector (UB in the reference). A read past the
 * end returns 0 here — the CUDA layout carries a zero-filled guard slab for the same purpose. */
static double gfo_flat(const gfo_grid* g, long long i) {
    const long long n = (long long)g->counts[0] * g->counts[1] * g->counts[2];
    return (i >= 0 && i < n) ? g->vals[i] : 0.0;
}

/* interpolation method 2 (:796-893): "tricubic Hermite" — cubic Hermite in x on the 4 cell edges with centred-difference
 * x-derivatives, then in y and in z with one-sided differences of the partly interpolated values. Restated statement by
 * statement, including what a derivation from scratch would not do: dvdy is formed on the z = iz plane only (:862), the
 * y- and z-differences use the value basis alone for the neighbour rows (:852-855, :866-867), and the derivative
 * estimates are switched off (0) in the first cell layer of an axis and in a one-point-wide band only (:817-832). */
static double gfo_interp_tricubic(const gfo_grid* g, double scale, const double* pi, double* f, gfo_class* cls) {
    const int nz = g->counts[2];
    const int nyz = g->counts[1] * nz;
    const double sx = g->spacing[0], sy = g->spacing[1], sz = g->spacing[2];
    int k, idx[3];
    double fr[3];
    for (k = 0; k < 3; k++) {
        idx[k] = (int)(pi[k] / g->spacing[k]);                   /* :708-710 */
        fr[k] = (pi[k] / g->spacing[k]) - idx[k];                /* :713-715 */
        if (idx[k] > g->counts[k] - 2) {                         /* quirk Q2, as in the trilinear branch */
            idx[k] = g->counts[k] - 2;
            fr[k] = (pi[k] / g->spacing[k]) - idx[k];
        }
    }
    if (cls) { cls->cell[0] = idx[0]; cls->cell[1] = idx[1]; cls->cell[2] = idx[2]; }
    {
        const int ix = idx[0], iy = idx[1], iz = idx[2];
        const double fx = fr[0], fy = fr[1], fz = fr[2];
        const long long im = (long long)ix * nyz + (long long)iy * nz + iz;                  /* :801-804 */
        const long long imp = im + nz, ip = im + nyz, ipp = ip + nz;
        const double f000 = gfo_flat(g, im), f001 = gfo_flat(g, im + 1);                     /* :806-813 */
        const double f010 = gfo_flat(g, imp), f011 = gfo_flat(g, imp + 1);
        const double f100 = gfo_flat(g, ip), f101 = gfo_flat(g, ip + 1);
        const double f110 = gfo_flat(g, ipp), f111 = gfo_flat(g, ipp + 1);
        const int xin = ix > 0 && ix < g->counts[0] - 1;                                     /* :817-832 */
        const int yin = iy > 0 && iy < g->counts[1] - 1;                                     /* :852-855 */
        const int zin = iz > 0 && iz < g->counts[2] - 1;                                     /* :866-867 */
        const double dx000 = xin ? (gfo_flat(g, im + nyz) - gfo_flat(g, im - nyz)) / (2.0 * sx) : 0.0;
        const double dx001 = xin ? (gfo_flat(g, im + 1 + nyz) - gfo_flat(g, im + 1 - nyz)) / (2.0 * sx) : 0.0;
        const double dx010 = xin ? (gfo_flat(g, imp + nyz) - gfo_flat(g, imp - nyz)) / (2.0 * sx) : 0.0;
        const double dx011 = xin ? (gfo_flat(g, imp + 1 + nyz) - gfo_flat(g, imp + 1 - nyz)) / (2.0 * sx) : 0.0;
        const double dx100 = xin ? (gfo_flat(g, im + 2 * (long long)nyz) - gfo_flat(g, im)) / (2.0 * sx) : 0.0;
        const double dx101 = xin ? (gfo_flat(g, im + 1 + 2 * (long long)nyz) - gfo_flat(g, im + 1)) / (2.0 * sx) : 0.0;
        const double dx110 = xin ? (gfo_flat(g, imp + 2 * (long long)nyz) - gfo_flat(g, imp)) / (2.0 * sx) : 0.0;
        const double dx111 = xin ? (gfo_flat(g, imp + 1 + 2 * (long long)nyz) - gfo_flat(g, imp + 1)) / (2.0 * sx) : 0.0;
        const double h00x = he_h00(fx), h01x = he_h01(fx), h10x = he_h10(fx), h11x = he_h11(fx);   /* :835-836 */
        const double d00x = he_d00(fx), d01x = he_d01(fx), d10x = he_d10(fx), d11x = he_d11(fx);
        const double v00 = h00x * f000 + h01x * f100 + h10x * dx000 * sx + h11x * dx100 * sx;      /* :838-841 */
        const double v01 = h00x * f001 + h01x * f101 + h10x * dx001 * sx + h11x * dx101 * sx;
        const double v10 = h00x * f010 + h01x * f110 + h10x * dx010 * sx + h11x * dx110 * sx;
        const double v11 = h00x * f011 + h01x * f111 + h10x * dx011 * sx + h11x * dx111 * sx;
        const double dv00 = d00x * f000 + d01x * f100 + d10x * dx000 * sx + d11x * dx100 * sx;     /* :843-846 */
        const double dv01 = d00x * f001 + d01x * f101 + d10x * dx001 * sx + d11x * dx101 * sx;
        const double dv10 = d00x * f010 + d01x * f110 + d10x * dx010 * sx + d11x * dx110 * sx;
        const double dv11 = d00x * f011 + d01x * f111 + d10x * dx011 * sx + d11x * dx111 * sx;
        const double dy00 = yin ? (v10 - (h00x * gfo_flat(g, im - nz) + h01x * gfo_flat(g, ip - nz))) / sy : 0.0;          /* :849-852 */
        const double dy01 = yin ? (v11 - (h00x * gfo_flat(g, im + 1 - nz) + h01x * gfo_flat(g, ip + 1 - nz))) / sy : 0.0;
        const double dy10 = yin ? ((h00x * gfo_flat(g, im + 2 * nz) + h01x * gfo_flat(g, ip + 2 * nz)) - v00) / sy : 0.0;
        const double dy11 = yin ? ((h00x * gfo_flat(g, im + 1 + 2 * nz) + h01x * gfo_flat(g, ip + 1 + 2 * nz)) - v01) / sy : 0.0;
        const double h00y = he_h00(fy), h01y = he_h01(fy), h10y = he_h10(fy), h11y = he_h11(fy);   /* :855-856 */
        const double d00y = he_d00(fy), d01y = he_d01(fy), d10y = he_d10(fy), d11y = he_d11(fy);
        const double v0 = h00y * v00 + h01y * v10 + h10y * dy00 * sy + h11y * dy10 * sy;           /* :858-859 */
        const double v1 = h00y * v01 + h01y * v11 + h10y * dy01 * sy + h11y * dy11 * sy;
        const double dvdx_0 = h00y * dv00 + h01y * dv10;                                            /* :861-862 */
        const double dvdx_1 = h00y * dv01 + h01y * dv11;
        double dvdy = (d00y * v00 + d01y * v10 + d10y * dy00 * sy + d11y * dy10 * sy);              /* :863 */
        const double dz0 = zin ? (v1 - (h00y * (h00x * gfo_flat(g, im - 1) + h01x * gfo_flat(g, ip - 1)) +
                                        h01y * (h00x * gfo_flat(g, imp - 1) + h01x * gfo_flat(g, ipp - 1)))) / sz : 0.0;   /* :866 */
        const double dz1 = zin ? ((h00y * (h00x * gfo_flat(g, im + 2) + h01x * gfo_flat(g, ip + 2)) +
                                   h01y * (h00x * gfo_flat(g, imp + 2) + h01x * gfo_flat(g, ipp + 2))) - v0) / sz : 0.0;   /* :867 */
        const double h00z = he_h00(fz), h01z = he_h01(fz), h10z = he_h10(fz), h11z = he_h11(fz);   /* :870-871 */
        const double d00z = he_d00(fz), d01z = he_d01(fz), d10z = he_d10(fz), d11z = he_d11(fz);
        double interpolated = h00z * v0 + h01z * v1 + h10z * dz0 * sz + h11z * dz1 * sz;            /* :873 */
        double dvdx = h00z * dvdx_0 + h01z * dvdx_1;                                                /* :875 */
        double dvdz = d00z * v0 + d01z * v1 + d10z * dz0 * sz + d11z * dz1 * sz;                    /* :876 */
        double grd[3];
        if (g->inv_power > 0.0) {                                                                   /* :879-886 */
            const double base = interpolated;
            const double pf = g->inv_power * pow(base, g->inv_power - 1.0);
            interpolated = pow(interpolated, g->inv_power);
            dvdx *= pf; dvdy *= pf; dvdz *= pf;
        }
        grd[0] = dvdx / sx; grd[1] = dvdy / sy; grd[2] = dvdz / sz;                                 /* :889 */
        if (f) for (k = 0; k < 3; k++) f[k] -= scale * grd[k];                                      /* :893 */
        return scale * interpolated;                                                                /* :892 */
    }
}

/* The inside branch for one atom (:706-1084). pi = position - origin. Returns scale * V and
 * subtracts scale * grad V from f[3]. */
static double gfo_interp(const gfo_grid* g, double scale, const double* pi, double* f, gfo_class* cls) {
    const int nz = g->counts[2];
    const int nyz = g->counts[1] * nz;                           /* :653 */
    int k, idx[3];
    double fr[3];
    if (g->interp_method == 1) return gfo_interp_bspline(g, scale, pi, f, cls);               /* :727 */
    if (g->interp_method == 2) return gfo_interp_tricubic(g, scale, pi, f, cls);              /* :796 */
    for (k = 0; k < 3; k++) {
        idx[k] = (int)(pi[k] / g->spacing[k]);                   /* :708-710 */
        fr[k] = (pi[k] / g->spacing[k]) - idx[k];                /* :713-715 */
        /* pi == hCorner gives idx == n-1 and the reference reads past the grid (UB, quirk Q2).
         * The restatement evaluates the same cell the limit from inside would: idx = n-2, f = 1. */
        if (idx[k] > g->counts[k] - 2) {
            idx[k] = g->counts[k] - 2;
            fr[k] = (pi[k] / g->spacing[k]) - idx[k];
        }
    }
    if (cls) { cls->cell[0] = idx[0]; cls->cell[1] = idx[1]; cls->cell[2] = idx[2]; }
    {
        {
            const double* v = g->vals;
            const int im = idx[0] * nyz + idx[1] * nz + idx[2];  /* :1022 */
            const int imp = im + nz, ip = im + nyz, ipp = ip + nz; /* :1023-1025 */
            const double vmmm = v[im], vmmp = v[im + 1], vmpm = v[imp], vmpp = v[imp + 1];   /* :1028-1031 */
            const double vpmm = v[ip], vpmp = v[ip + 1], vppm = v[ipp], vppp = v[ipp + 1];   /* :1033-1036 */
            const double fx = fr[0], fy = fr[1], fz = fr[2];
            const double ax = 1.0 - fx, ay = 1.0 - fy, az = 1.0 - fz;                       /* :1039-1041 */
            const double vmm = az * vmmm + fz * vmmp;                                       /* :1044-1047 */
            const double vmp = az * vmpm + fz * vmpp;
            const double vpm = az * vpmm + fz * vpmp;
            const double vpp = az * vppm + fz * vppp;
            const double vm = ay * vmm + fy * vmp;                                          /* :1049-1050 */
            const double vp = ay * vpm + fy * vpp;
            double interpolated = ax * vm + fx * vp;                                        /* :1053 */
            double dvdx, dvdy, dvdz, grd[3], enr;
            if (g->inv_power > 0.0) interpolated = pow(interpolated, g->inv_power);         /* :1057-1059 */
            enr = scale * interpolated;                                                     /* :1061 */
            dvdx = -vm + vp;                                                                /* :1066 */
            dvdy = (-vmm + vmp) * ax + (-vpm + vpp) * fx;                                   /* :1068 */
            dvdz = ((-vmmm + vmmp) * ay + (-vmpm + vmpp) * fy) * ax +
                   ((-vpmm + vpmp) * ay + (-vppm + vppp) * fy) * fx;                        /* :1070-1071 */
            grd[0] = dvdx / g->spacing[0];                                                  /* :1072 */
            grd[1] = dvdy / g->spacing[1];
            grd[2] = dvdz / g->spacing[2];
            if (g->inv_power > 0.0) {                                                       /* :1076-1080 */
                const double base = ax * vm + fx * vp;
                const double pf = g->inv_power * pow(base, g->inv_power - 1.0);
                grd[0] = grd[0] * pf; grd[1] = grd[1] * pf; grd[2] = grd[2] * pf;
            }
            if (f) for (k = 0; k < 3; k++) f[k] -= scale * grd[k];                          /* :1082 */
            return enr;
        }
    }
}

double gfo_execute(const gfo_grid* grid, const double* scaling, int n_scaling, const int* ligand_atoms,
                   const double* pos, double* forces, gfo_class* cls) {
    double energy = 0.0;
    int ia;
    for (ia = 0; ia < n_scaling; ia++) {                                                     /* :682 */
        const int particle = ligand_atoms ? ligand_atoms[ia] : ia;                           /* :684 */
        const double* p = pos + 3 * (size_t)particle;
        const double scale = scaling[ia];
        double h[3], pi[3];
        int k, inside = 1;
        for (k = 0; k < 3; k++) {
            h[k] = grid->spacing[k] * (grid->counts[k] - 1);                                 /* :654-656 hCorner */
            pi[k] = p[k] - grid->origin[k];                                                  /* :687-688 */
            if (!(pi[k] >= 0.0 && pi[k] <= h[k])) inside = 0;                                /* :690-696, <= */
        }
        if (inside && scale != 0.0) {
            if (cls) cls[ia].inside = 1;
            energy += gfo_interp(grid, scale, pi, forces ? forces + 3 * (size_t)ia : 0, cls ? cls + ia : 0);
        } else {
            /* :1093-1117 — unscaled harmonic wall. Three separate terms are added to the running
             * energy (:1112); keep that association so the sum is bit-identical to the reference. */
            if (cls) { cls[ia].inside = inside; cls[ia].cell[0] = cls[ia].cell[1] = cls[ia].cell[2] = -1; }
            for (k = 0; k < 3; k++) {
                double dev = 0.0;
                if (pi[k] < 0.0) dev = pi[k];
                else if (pi[k] > h[k]) dev = pi[k] - h[k];
                energy += 0.5 * grid->oob_k * dev * dev;                                     /* :1112 */
                if (forces) forces[3 * (size_t)ia + k] -= grid->oob_k * dev;                 /* :1113-1116 */
            }
        }
    }
    return energy;                                                                           /* :1120 */
}

typedef struct {
    const gfo_grid* grids;
    int n_grids;
    const double* scaling;
    int r0, r1, n_atoms;
    const double* pos;
    double* forces;
    double* energies;
} gfo_job;

static void* gfo_worker(void* arg) {
    gfo_job* j = (gfo_job*)arg;
    int r, g;
    for (r = j->r0; r < j->r1; r++) {
        const double* p = j->pos + (size_t)r * j->n_atoms * 3;
        double* f = j->forces ? j->forces + (size_t)r * j->n_atoms * 3 : 0;
        if (f) memset(f, 0, sizeof(double) * 3 * (size_t)j->n_atoms);
        for (g = 0; g < j->n_grids; g++) {
            const double e = gfo_execute(&j->grids[g], j->scaling + (size_t)g * j->n_atoms, j->n_atoms, 0, p, f, 0);
            if (j->energies) j->energies[(size_t)r * j->n_grids + g] = e;
        }
    }
    return 0;
}

void gfo_execute_batched(const gfo_grid* grids, int n_grids, const double* scaling, int n_replicas, int n_atoms,
                         const double* pos, double* forces, double* energies, int n_threads) {
    int t;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_replicas) n_threads = n_replicas > 0 ? n_replicas : 1;
    {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
        gfo_job* jobs = (gfo_job*)malloc(sizeof(gfo_job) * n_threads);
        for (t = 0; t < n_threads; t++) {
            gfo_job j;
            j.grids = grids; j.n_grids = n_grids; j.scaling = scaling; j.n_atoms = n_atoms;
            j.r0 = (int)((long long)n_replicas * t / n_threads);
            j.r1 = (int)((long long)n_replicas * (t + 1) / n_threads);
            j.pos = pos; j.forces = forces; j.energies = energies;
            jobs[t] = j;
            if (t > 0) pthread_create(&th[t], 0, gfo_worker, &jobs[t]);
        }
        gfo_worker(&jobs[0]);
        for (t = 1; t < n_threads; t++) pthread_join(th[t], 0);
        free(th);
        free(jobs);
    }
}

/* ---- grid generation (ReferenceGridForceKernels.cpp:465-544) ------------------------------------------------- */
typedef struct {
    const int* counts;
    const double *spacing, *origin, *pos, *charges, *sigmas, *epsilons;
    int grid_type, n_atoms, i0, i1;
    double grid_cap;
    double* out;
} gfo_gen_job;

static void* gfo_gen_worker(void* arg) {
    gfo_gen_job* j = (gfo_gen_job*)arg;
    const int ny = j->counts[1], nz = j->counts[2];
    const double COULOMB_CONST = 138.935456;                                                  /* :493 */
    const double U_MAX = j->grid_cap;                                                         /* :494 */
    int i, jj, k, a;
    for (i = j->i0; i < j->i1; i++)
        for (jj = 0; jj < ny; jj++)
            for (k = 0; k < nz; k++) {
                const double gx = j->origin[0] + i * j->spacing[0];                           /* :502-504 */
                const double gy = j->origin[1] + jj * j->spacing[1];
                const double gz = j->origin[2] + k * j->spacing[2];
                double v = 0.0;
                for (a = 0; a < j->n_atoms; a++) {                                            /* :508 */
                    const double dx = gx - j->pos[3 * a], dy = gy - j->pos[3 * a + 1], dz = gz - j->pos[3 * a + 2];
                    const double r2 = dx * dx + dy * dy + dz * dz;                            /* :516 */
                    double r = sqrt(r2);
                    if (r < 1e-6) r = 1e-6;                                                   /* :520-522 */
                    if (j->grid_type == 1) {
                        v += COULOMB_CONST * j->charges[a] / r;                               /* :527 */
                    } else if (j->grid_type == 2) {
                        const double diameter = 2.0 * j->sigmas[a];
                        v += sqrt(j->epsilons[a]) * pow(diameter, 6.0) / pow(r, 12.0);        /* :531 */
                    } else if (j->grid_type == 3) {
                        const double diameter = 2.0 * j->sigmas[a];
                        v += -2.0 * sqrt(j->epsilons[a]) * pow(diameter, 3.0) / pow(r, 6.0);  /* :535 */
                    }
                }
                j->out[((size_t)i * ny + jj) * nz + k] = U_MAX * tanh(v / U_MAX);             /* :540-541 */
            }
    return 0;
}

void gfo_generate_grid(const int counts[3], const double spacing[3], const double origin[3], int grid_type, int n_atoms,
                       const double* pos, const double* charges, const double* sigmas, const double* epsilons,
                       double grid_cap, double* out, int n_threads) {
    int t;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > counts[0]) n_threads = counts[0];
    {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
        gfo_gen_job* jobs = (gfo_gen_job*)malloc(sizeof(gfo_gen_job) * n_threads);
        for (t = 0; t < n_threads; t++) {
            gfo_gen_job j;
            j.counts = counts; j.spacing = spacing; j.origin = origin; j.pos = pos;
            j.charges = charges; j.sigmas = sigmas; j.epsilons = epsilons;
            j.grid_type = grid_type; j.n_atoms = n_atoms; j.grid_cap = grid_cap; j.out = out;
            j.i0 = (int)((long long)counts[0] * t / n_threads);
            j.i1 = (int)((long long)counts[0] * (t + 1) / n_threads);
            jobs[t] = j;
            if (t > 0) pthread_create(&th[t], 0, gfo_gen_worker, &jobs[t]);
        }
        gfo_gen_worker(&jobs[0]);
        for (t = 1; t < n_threads; t++) pthread_join(th[t], 0);
        free(th);
        free(jobs);
    }
}

/* GridForce::applyInvPowerTransformation (openmmapi/src/GridForce.cpp:262-268): G -> sign(G) * |G|^(1/n), zeros kept. */
void gfo_inv_power_transform(double* vals, size_t n, double inv_power) {
    size_t i;
    for (i = 0; i < n; i++) {
        if (vals[i] != 0.0) {
            const double sign = (vals[i] >= 0.0) ? 1.0 : -1.0;
            vals[i] = sign * pow(fabs(vals[i]), 1.0 / inv_power);
        }
    }
}
