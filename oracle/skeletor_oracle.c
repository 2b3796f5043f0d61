/* skeletor_oracle.c — CPU restatement of skeletor's particle hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's CPU-baseline legs may load this library; the product path
 * (skeletor_b200/ + libskeletor_b200.so) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit (particle
 * work) / to summation order (deposit is bit-exact too, the loop order is the
 * reference's) against the unmodified reference compiled into oracle/_ref by
 * oracle/build_ref.py (tests/test_oracle_vs_reference.py) and against the
 * committed fixtures tests/golden/*.npz generated from that build by
 * oracle/make_golden.py (tests/test_oracle_golden.py).
 *
 * Each function cites the reference file:line whose arithmetic it restates
 * (paths relative to the reference root).  Operation order is kept exactly so
 * that gcc -O2 -ffp-contract=off reproduces the reference's SSE2 results.
 *
 * Layouts (the reference's): particles AoS {x,y,vx,vy,vz} (types.pxd:10-11);
 * fields C-order [myp][mx] of interleaved structs, nc = 3 (E,B) or 4 (sources:
 * rho,Jx,Jy,Jz) (types.pyx:9-10, field.py:6-9).
 */
#include <math.h>
#include <string.h>

typedef struct {
  int nx, ny, nyp, noff, lbx, lby, ubx, uby;
  double dx, dy, Lx, Ly, x0, y0, edges[2];
} ogrid_t;

#define F3(F, mx, iy, ix, c) (F)[((long)(iy) * (mx) + (ix)) * 3 + (c)]
#define F4(F, mx, iy, ix, c) (F)[((long)(iy) * (mx) + (ix)) * 4 + (c)]

/* --- gather ------------------------------------------------------------- */

/* particle_push.pxd:3-27 */
static void gather1(const double *p, const double *F, int mx, double ox,
                    double oy, double f[3]) {
  double x = p[0] + ox, y = p[1] + oy;
  int ix = (int)x, iy = (int)y;
  double dx = x - (double)ix, dy = y - (double)iy;
  double tx = 1.0 - dx, ty = 1.0 - dy;
  for (int c = 0; c < 3; c++)
    f[c] = dy * (dx * F3(F, mx, iy + 1, ix + 1, c) + tx * F3(F, mx, iy + 1, ix, c)) +
           ty * (dx * F3(F, mx, iy, ix + 1, c) + tx * F3(F, mx, iy, ix, c));
}

/* particle_push.pxd:29-67 */
static void tsc_weights(double x, int *i, double w[3]) {
  /* caller has already added the +0.5 */
  *i = (int)x;
  double d = x - (double)*i - 0.5;
  w[1] = 0.75 - d * d;
  w[2] = 0.5 * ((0.5 + d) * (0.5 + d)); /* pow(0.5+d, 2.0) == exact square */
  w[0] = 1.0 - (w[1] + w[2]);
}

static void gather2(const double *p, const double *F, int mx, double ox,
                    double oy, double f[3]) {
  double x = p[0] + ox, y = p[1] + oy;
  x = x + 0.5;
  y = y + 0.5;
  int ix, iy;
  double wx[3], wy[3];
  tsc_weights(x, &ix, wx);
  tsc_weights(y, &iy, wy);
  for (int c = 0; c < 3; c++) {
    double r0 = wx[0] * F3(F, mx, iy - 1, ix - 1, c) + wx[1] * F3(F, mx, iy - 1, ix, c) +
                wx[2] * F3(F, mx, iy - 1, ix + 1, c);
    double r1 = wx[0] * F3(F, mx, iy, ix - 1, c) + wx[1] * F3(F, mx, iy, ix, c) +
                wx[2] * F3(F, mx, iy, ix + 1, c);
    double r2 = wx[0] * F3(F, mx, iy + 1, ix - 1, c) + wx[1] * F3(F, mx, iy + 1, ix, c) +
                wx[2] * F3(F, mx, iy + 1, ix + 1, c);
    f[c] = wy[0] * r0 + wy[1] * r1 + wy[2] * r2;
  }
}

/* particle_push.pxd:69-86 (kick_particle) */
static void boris_kick(double *p, const double e[3], const double b[3]) {
  double vmx = p[2] + e[0], vmy = p[3] + e[1], vmz = p[4] + e[2];
  double vpx = vmx + (vmy * b[2] - vmz * b[1]);
  double vpy = vmy + (vmz * b[0] - vmx * b[2]);
  double vpz = vmz + (vmx * b[1] - vmy * b[0]);
  double fac = 2. / (1. + b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
  p[2] = vmx + fac * (vpy * b[2] - vpz * b[1]) + e[0];
  p[3] = vmy + fac * (vpz * b[0] - vpx * b[2]) + e[1];
  p[4] = vmz + fac * (vpx * b[1] - vpy * b[0]) + e[2];
}

/* fields at the particle, rescaled; particle_push.pyx:28-35 (+ :75-76 modified) */
static void fields_at(const double *p, const double *E, const double *B,
                      const ogrid_t *g, int order, double qtmh, int modified,
                      double dt, double Omega, double S, double e[3], double b[3]) {
  int mx = g->nx + 2 * g->lbx;
  double obx = g->lbx, oby = g->lby - g->noff; /* particle_push.pyx:16-17 */
  double oex = obx - 0.5, oey = oby - 0.5;     /* :18-19 */
  if (order == 1) {
    gather1(p, E, mx, oex, oey, e);
    gather1(p, B, mx, obx, oby, b);
  } else {
    gather2(p, E, mx, oex, oey, e);
    gather2(p, B, mx, obx, oby, b);
  }
  for (int c = 0; c < 3; c++) { /* rescale, particle_push.pxd:93-97 */
    e[c] = e[c] * qtmh;
    b[c] = b[c] * qtmh;
  }
  if (modified) { /* particle_push.pyx:75-76 */
    b[2] = b[2] + Omega * dt;
    e[1] = e[1] - S * (g->y0 + p[1] * g->dy) * b[2];
  }
}

/* boris_push_cic/tsc, modified_boris_push_cic/tsc — particle_push.pyx:4-156 */
void orc_push(double *part, long np, const double *E, const double *B,
              const ogrid_t *g, int order, double qtmh, double dt, int modified,
              double Omega, double S) {
  double dtdsx = dt / g->dx, dtdsy = dt / g->dy;
  for (long ip = 0; ip < np; ip++) {
    double *p = part + 5 * ip, e[3], b[3];
    fields_at(p, E, B, g, order, qtmh, modified, dt, Omega, S, e, b);
    boris_kick(p, e, b);
    p[0] = p[0] + p[2] * dtdsx; /* drift_particle, particle_push.pxd:88-91 */
    p[1] = p[1] + p[3] * dtdsy;
  }
}

/* drift — particle_push.pyx:159-169 */
void orc_drift(double *part, long np, const ogrid_t *g, double dt) {
  double dtdsx = dt / g->dx, dtdsy = dt / g->dy;
  for (long ip = 0; ip < np; ip++) {
    double *p = part + 5 * ip;
    p[0] = p[0] + p[2] * dtdsx;
    p[1] = p[1] + p[3] * dtdsy;
  }
}

/* --- particle boundary conditions --------------------------------------- */

/* periodic_x — particle_boundary.pxd:3-7, pyx:5-11 */
static void wrap_x(double *p, double nx) {
  while (p[0] < 0.0) p[0] = p[0] + nx;
  while (p[0] >= nx) p[0] = p[0] - nx;
}
void orc_periodic_x(double *part, long np, const ogrid_t *g) {
  double nx = (double)g->nx;
  for (long ip = 0; ip < np; ip++) wrap_x(part + 5 * ip, nx);
}

/* shear_periodic_y — particle_boundary.pyx:26-49 */
void orc_shear_periodic_y(double *part, long np, const ogrid_t *g, double S,
                          double t) {
  double ny = (double)g->ny;
  double vx_boost = S * g->Ly;
  double x_boost = vx_boost * t / g->dx;
  for (long ip = 0; ip < np; ip++) {
    double *p = part + 5 * ip;
    if (p[1] < 0.0) {
      p[0] = p[0] - x_boost;
      p[2] = p[2] - vx_boost;
    }
    if (p[1] >= ny) {
      p[0] = p[0] + x_boost;
      p[2] = p[2] + vx_boost;
    }
  }
}

/* calculate_ihole — particle_boundary.pxd:10-22, pyx:14-23.
 * ihole has ntmax+1 entries; ihole[0] = count, or -count on overflow. */
static int note_hole(const double *p, int *ihole, int ntmax, const ogrid_t *g,
                     int ih, long ip) {
  if (p[1] < g->edges[0] || p[1] >= g->edges[1]) {
    if (ih < ntmax)
      ihole[ih + 1] = (int)ip + 1;
    else
      ihole[0] = -ih;
    ih += 1;
  }
  return ih;
}
void orc_calculate_ihole(const double *part, long np, int *ihole, int ntmax,
                         const ogrid_t *g) {
  int ih = 0;
  for (long ip = 0; ip < np; ip++)
    ih = note_hole(part + 5 * ip, ihole, ntmax, g, ih, ip);
  if (ihole[0] >= 0) ihole[0] = ih;
}

/* --- deposit ------------------------------------------------------------ */

/* deposit_particle_cic — deposit.pxd:3-44 */
static void scatter1(const double *p, double *cur, int mx, const ogrid_t *g,
                     double S, double ox, double oy) {
  double x = p[0] + ox, y = p[1] + oy;
  int ix = (int)x, iy = (int)y;
  double dx = x - (double)ix, dy = y - (double)iy;
  double tx = 1.0 - dx, ty = 1.0 - dy;
  double vx = p[2] + S * (p[1] * g->dy + g->y0);
  const double wy[2] = {ty, dy}, wx[2] = {tx, dx};
  for (int a = 0; a < 2; a++)
    for (int c = 0; c < 2; c++) {
      double *q = &F4(cur, mx, iy + a, ix + c, 0);
      q[0] += wy[a] * wx[c];
      q[1] += wy[a] * wx[c] * vx;
      q[2] += wy[a] * wx[c] * p[3];
      q[3] += wy[a] * wx[c] * p[4];
    }
}

/* deposit_particle_tsc — deposit.pxd:46-118 */
static void scatter2(const double *p, double *cur, int mx, const ogrid_t *g,
                     double S, double ox, double oy) {
  double x = p[0] + ox + 0.5, y = p[1] + oy + 0.5;
  int ix, iy;
  double wx[3], wy[3];
  tsc_weights(x, &ix, wx);
  tsc_weights(y, &iy, wy);
  double vx = p[2] + S * (p[1] * g->dy + g->y0);
  for (int a = 0; a < 3; a++)
    for (int c = 0; c < 3; c++) {
      double *q = &F4(cur, mx, iy - 1 + a, ix - 1 + c, 0);
      q[0] += wy[a] * wx[c];
      q[1] += wy[a] * wx[c] * vx;
      q[2] += wy[a] * wx[c] * p[3];
      q[3] += wy[a] * wx[c] * p[4];
    }
}

/* deposit_cic / deposit_tsc — deposit.pyx:6-34 */
void orc_deposit(const double *part, long np, double *cur, const ogrid_t *g,
                 int order, double S) {
  int mx = g->nx + 2 * g->lbx;
  double ox = g->lbx - 0.5, oy = g->lby - 0.5 - g->noff; /* deposit.pyx:14-15 */
  for (long ip = 0; ip < np; ip++) {
    if (order == 1)
      scatter1(part + 5 * ip, cur, mx, g, S, ox, oy);
    else
      scatter2(part + 5 * ip, cur, mx, g, S, ox, oy);
  }
}

/* push_and_deposit_cic/tsc — push_and_deposit.pyx:10-170 */
void orc_push_and_deposit(double *part, long np, const double *E,
                          const double *B, const ogrid_t *g, int order,
                          double qtmh, double dt, int *ihole, int ntmax,
                          double *cur, double S, int update) {
  int mx = g->nx + 2 * g->lbx;
  double oex = (double)g->lbx - 0.5, oey = (double)(g->lby - g->noff) - 0.5;
  double d2x = 0.5 * dt / g->dx, d2y = 0.5 * dt / g->dy; /* :37-38 */
  double nx = (double)g->nx;
  int ih = 0;
  for (long ip = 0; ip < np; ip++) {
    double q[5], e[3], b[3];
    memcpy(q, part + 5 * ip, sizeof q);
    fields_at(part + 5 * ip, E, B, g, order, qtmh, 0, dt, 0.0, 0.0, e, b);
    boris_kick(q, e, b);
    q[0] = q[0] + q[2] * d2x;
    q[1] = q[1] + q[3] * d2y;
    if (fabs(q[0] - part[5 * ip]) > 0.5 || fabs(q[1] - part[5 * ip + 1]) > 0.5)
      ihole[0] = -1; /* :66-68 */
    if (order == 1)
      scatter1(q, cur, mx, g, S, oex, oey); /* offsetE reused, :71 */
    else
      scatter2(q, cur, mx, g, S, oex, oey);
    if (update) {
      q[0] = q[0] + q[2] * d2x;
      q[1] = q[1] + q[3] * d2y;
      wrap_x(q, nx);
      ih = note_hole(q, ihole, ntmax, g, ih, ip);
      memcpy(part + 5 * ip, q, sizeof q);
    }
  }
  if (update && ihole[0] >= 0) ihole[0] = ih;
}

/* --- finite differences — finite_difference.pyx:5-85 -------------------- */
/* f* are scalar planes with row stride `mx` and element stride `es` doubles
 * (es = 3 or 4 when the plane is a component of an interleaved field). */
#define P(f, iy, ix) (f)[((long)(iy) * mx + (ix)) * es]

void orc_gradient(const double *f, int es, double *grad, const ogrid_t *g) {
  int mx = g->nx + 2 * g->lbx;
  for (int iy = g->lby; iy < g->uby; iy++)
    for (int ix = g->lbx; ix < g->ubx; ix++) {
      F3(grad, mx, iy, ix, 0) = 0.5 / g->dx * (P(f, iy, ix + 1) - P(f, iy, ix - 1));
      F3(grad, mx, iy, ix, 1) = 0.5 / g->dy * (P(f, iy + 1, ix) - P(f, iy - 1, ix));
      F3(grad, mx, iy, ix, 2) = 0.0;
    }
}

static double ddyup(const double *f, int es, int mx, int ix, int iy, const ogrid_t *g) {
  return 0.5 / g->dy * (P(f, iy + 1, ix + 1) + P(f, iy + 1, ix) - P(f, iy, ix + 1) - P(f, iy, ix));
}
static double ddxup(const double *f, int es, int mx, int ix, int iy, const ogrid_t *g) {
  return 0.5 / g->dx * (P(f, iy + 1, ix + 1) + P(f, iy, ix + 1) - P(f, iy + 1, ix) - P(f, iy, ix));
}
static double ddydn(const double *f, int es, int mx, int ix, int iy, const ogrid_t *g) {
  return 0.5 / g->dy * (P(f, iy, ix) + P(f, iy, ix - 1) - P(f, iy - 1, ix) - P(f, iy - 1, ix - 1));
}
static double ddxdn(const double *f, int es, int mx, int ix, int iy, const ogrid_t *g) {
  return 0.5 / g->dx * (P(f, iy, ix) + P(f, iy - 1, ix) - P(f, iy, ix - 1) - P(f, iy - 1, ix - 1));
}

/* curl_up (down=0) / curl_down (down=1): finite_difference.pyx:15-35 */
void orc_curl(const double *fx, const double *fy, const double *fz, int es,
              double *curl, const ogrid_t *g, int down) {
  int mx = g->nx + 2 * g->lbx;
  for (int iy = g->lby; iy < g->uby; iy++)
    for (int ix = g->lbx; ix < g->ubx; ix++) {
      if (down) {
        F3(curl, mx, iy, ix, 0) = ddydn(fz, es, mx, ix, iy, g);
        F3(curl, mx, iy, ix, 1) = -ddxdn(fz, es, mx, ix, iy, g);
        F3(curl, mx, iy, ix, 2) = ddxdn(fy, es, mx, ix, iy, g) - ddydn(fx, es, mx, ix, iy, g);
      } else {
        F3(curl, mx, iy, ix, 0) = ddyup(fz, es, mx, ix, iy, g);
        F3(curl, mx, iy, ix, 1) = -ddxup(fz, es, mx, ix, iy, g);
        F3(curl, mx, iy, ix, 2) = ddxup(fy, es, mx, ix, iy, g) - ddyup(fx, es, mx, ix, iy, g);
      }
    }
}

/* divergence: finite_difference.pyx:37-45; div is a plain scalar plane */
void orc_divergence(const double *fx, const double *fy, int es, double *div,
                    const ogrid_t *g) {
  int mx = g->nx + 2 * g->lbx;
  for (int iy = g->lby; iy < g->uby; iy++)
    for (int ix = g->lbx; ix < g->ubx; ix++)
      div[(long)iy * mx + ix] = ddxdn(fx, es, mx, ix, iy, g) + ddydn(fy, es, mx, ix, iy, g);
}

/* unstagger (up=0, inter_dn) / stagger (up=1, inter_up): pyx:59-85 */
void orc_interp(const double *fx, const double *fy, const double *fz, int es,
                double *out, const ogrid_t *g, int up) {
  int mx = g->nx + 2 * g->lbx;
  const double *f[3] = {fx, fy, fz};
  for (int iy = g->lby; iy < g->uby; iy++)
    for (int ix = g->lbx; ix < g->ubx; ix++)
      for (int c = 0; c < 3; c++) {
        const double *h = f[c];
        F3(out, mx, iy, ix, c) =
            up ? 0.25 * (P(h, iy + 1, ix + 1) + P(h, iy + 1, ix) + P(h, iy, ix + 1) + P(h, iy, ix))
               : 0.25 * (P(h, iy, ix) + P(h, iy - 1, ix) + P(h, iy, ix - 1) + P(h, iy - 1, ix - 1));
      }
}
