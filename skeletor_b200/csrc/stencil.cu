// stencil.cu — second-order finite differences on the extended grid, and the
// fused Ohm / Faraday updates.
//
// Replaces gradient, curl_up, curl_down, divergence, unstagger, stagger
// (reference skeletor/cython/finite_difference.pyx:5-85) and, fused into one pass
// each, the whole-array NumPy sequences of Ohm.__call__ (skeletor/ohm.py:35-75)
// and Faraday.__call__ (skeletor/faraday.py:16-30).  Active cells only; guards of
// the inputs must be set (reference second_order.py:25,36,51).  Operation order
// is the reference's; the one transcendental (log rho) may differ from NumPy's in
// the last bit.
#include "common.cuh"

#define ST 256

// scalar plane accessor: element stride es doubles, row stride mx*es
struct Plane {
  const double *p;
  int es, mx;
  __device__ __forceinline__ double operator()(int iy, int ix) const {
    return p[((size_t)iy * mx + ix) * es];
  }
};

// finite_difference.pyx:47-57
__device__ __forceinline__ double ddyup(const Plane &f, int ix, int iy, double dy) {
  return 0.5 / dy * (f(iy + 1, ix + 1) + f(iy + 1, ix) - f(iy, ix + 1) - f(iy, ix));
}
__device__ __forceinline__ double ddxup(const Plane &f, int ix, int iy, double dx) {
  return 0.5 / dx * (f(iy + 1, ix + 1) + f(iy, ix + 1) - f(iy + 1, ix) - f(iy, ix));
}
__device__ __forceinline__ double ddydn(const Plane &f, int ix, int iy, double dy) {
  return 0.5 / dy * (f(iy, ix) + f(iy, ix - 1) - f(iy - 1, ix) - f(iy - 1, ix - 1));
}
__device__ __forceinline__ double ddxdn(const Plane &f, int ix, int iy, double dx) {
  return 0.5 / dx * (f(iy, ix) + f(iy - 1, ix) - f(iy, ix - 1) - f(iy - 1, ix - 1));
}
// finite_difference.pyx:81-85
__device__ __forceinline__ double inter_up(const Plane &f, int ix, int iy) {
  return 0.25 * (f(iy + 1, ix + 1) + f(iy + 1, ix) + f(iy, ix + 1) + f(iy, ix));
}
__device__ __forceinline__ double inter_dn(const Plane &f, int ix, int iy) {
  return 0.25 * (f(iy, ix) + f(iy - 1, ix) + f(iy, ix - 1) + f(iy - 1, ix - 1));
}

__device__ __forceinline__ bool active_cell(const DevGrid &g, int &iy, int &ix) {
  long long idx = (long long)blockIdx.x * ST + threadIdx.x;
  if (idx >= (long long)g.nx * g.nyp) return false;
  iy = (int)(idx / g.nx);
  ix = (int)(idx - (long long)iy * g.nx) + g.lbx;
  iy += g.lby;
  return true;
}

__global__ void __launch_bounds__(ST)
gradient_kernel(Plane f, double *grad, DevGrid g) {
  int iy, ix;
  if (!active_cell(g, iy, ix)) return;
  double *o = grad + ((size_t)iy * g.mx + ix) * 3;
  o[0] = 0.5 / g.dx * (f(iy, ix + 1) - f(iy, ix - 1));
  o[1] = 0.5 / g.dy * (f(iy + 1, ix) - f(iy - 1, ix));
  o[2] = 0.0;
}

__global__ void __launch_bounds__(ST)
curl_kernel(Plane fx, Plane fy, Plane fz, double *curl, DevGrid g, int down) {
  int iy, ix;
  if (!active_cell(g, iy, ix)) return;
  double *o = curl + ((size_t)iy * g.mx + ix) * 3;
  if (down) {
    o[0] = ddydn(fz, ix, iy, g.dy);
    o[1] = -ddxdn(fz, ix, iy, g.dx);
    o[2] = ddxdn(fy, ix, iy, g.dx) - ddydn(fx, ix, iy, g.dy);
  } else {
    o[0] = ddyup(fz, ix, iy, g.dy);
    o[1] = -ddxup(fz, ix, iy, g.dx);
    o[2] = ddxup(fy, ix, iy, g.dx) - ddyup(fx, ix, iy, g.dy);
  }
}

__global__ void __launch_bounds__(ST)
divergence_kernel(Plane fx, Plane fy, double *div, DevGrid g) {
  int iy, ix;
  if (!active_cell(g, iy, ix)) return;
  div[(size_t)iy * g.mx + ix] = ddxdn(fx, ix, iy, g.dx) + ddydn(fy, ix, iy, g.dy);
}

__global__ void __launch_bounds__(ST)
interp_kernel(Plane fx, Plane fy, Plane fz, double *out, DevGrid g, int up) {
  int iy, ix;
  if (!active_cell(g, iy, ix)) return;
  double *o = out + ((size_t)iy * g.mx + ix) * 3;
  if (up) {
    o[0] = inter_up(fx, ix, iy); o[1] = inter_up(fy, ix, iy); o[2] = inter_up(fz, ix, iy);
  } else {
    o[0] = inter_dn(fx, ix, iy); o[1] = inter_dn(fy, ix, iy); o[2] = inter_dn(fz, ix, iy);
  }
}

// ohm.py:35-75 in one pass.  src = Float4 (rho, Jx, Jy, Jz), B = Float3.
__global__ void __launch_bounds__(ST)
ohm_kernel(const double *__restrict__ src, const double *__restrict__ B, double *E,
           double *Je_out, double *Bc_out, DevGrid g, double alpha, double eta,
           const int *skip) {
  int iy, ix;
  if (skip && *skip) return;                  // (device-side loop control of the steppers)
  if (!active_cell(g, iy, ix)) return;
  const Plane rho{src, 4, g.mx};
  const Plane bx{B, 3, g.mx}, by{B + 1, 3, g.mx}, bz{B + 2, 3, g.mx};
  // electron pressure: gradient(log(rho)) then *= -alpha        (ohm.py:38-40)
  double ex = 0.5 / g.dx * (log(rho(iy, ix + 1)) - log(rho(iy, ix - 1)));
  double ey = 0.5 / g.dy * (log(rho(iy + 1, ix)) - log(rho(iy - 1, ix)));
  double ez = 0.0;
  ex = ex * (-alpha);
  ey = ey * (-alpha);
  // total current J = curl_down(B)                                (ohm.py:43)
  double jx = ddydn(bz, ix, iy, g.dy);
  double jy = -ddxdn(bz, ix, iy, g.dx);
  double jz = ddxdn(by, ix, iy, g.dx) - ddydn(bx, ix, iy, g.dy);
  // resistive part                                                (ohm.py:46-48)
  ex = ex + eta * jx;
  ey = ey + eta * jy;
  ez = ez + eta * jz;
  // negative electron fluid velocity (J - J_i)/rho                (ohm.py:57-61)
  const double *s = src + ((size_t)iy * g.mx + ix) * 4;
  jx = jx - s[1]; jx = jx / s[0];
  jy = jy - s[2]; jy = jy / s[0];
  jz = jz - s[3]; jz = jz / s[0];
  // B at the location of E                                         (ohm.py:64)
  const double cx = inter_dn(bx, ix, iy), cy = inter_dn(by, ix, iy), cz = inter_dn(bz, ix, iy);
  // J_e x B                                                        (ohm.py:67-69)
  ex = ex + (jy * cz - jz * cy);
  ey = ey + (jz * cx - jx * cz);
  ez = ez + (jx * cy - jy * cx);
  double *o = E + ((size_t)iy * g.mx + ix) * 3;
  o[0] = ex; o[1] = ey; o[2] = ez;
  if (Je_out) { double *q = Je_out + ((size_t)iy * g.mx + ix) * 3; q[0] = jx; q[1] = jy; q[2] = jz; }
  if (Bc_out) { double *q = Bc_out + ((size_t)iy * g.mx + ix) * 3; q[0] = cx; q[1] = cy; q[2] = cz; }
}

// faraday.py:16-30: B -= curl_up(E)*dt.  E (with guards) is read-only, B is updated
// cell by cell: in place (Bin == B) or from a second field (B = Bin - curl_up(E)*dt, the
// "B3 = B; faraday(E2, B3, dt)" of horowitz.py:147-148 in one pass).
__global__ void __launch_bounds__(ST)
faraday_kernel(const double *__restrict__ E, const double *Bin, double *B, double *dB_out,
               DevGrid g, double dt, const int *skip) {
  int iy, ix;
  if (skip && *skip) return;
  if (!active_cell(g, iy, ix)) return;
  const Plane ex{E, 3, g.mx}, ey{E + 1, 3, g.mx}, ez{E + 2, 3, g.mx};
  const double cx = ddyup(ez, ix, iy, g.dy);
  const double cy = -ddxup(ez, ix, iy, g.dx);
  const double cz = ddxup(ey, ix, iy, g.dx) - ddyup(ex, ix, iy, g.dy);
  double *b = B + ((size_t)iy * g.mx + ix) * 3;
  const double *bi = Bin + ((size_t)iy * g.mx + ix) * 3;
  b[0] = bi[0] - cx * dt;
  b[1] = bi[1] - cy * dt;
  b[2] = bi[2] - cz * dt;
  if (dB_out) { double *q = dB_out + ((size_t)iy * g.mx + ix) * 3; q[0] = cx; q[1] = cy; q[2] = cz; }
}

static inline unsigned sblk(const DevGrid &g) {
  return (unsigned)(((long long)g.nx * g.nyp + ST - 1) / ST);
}

extern "C" int skb_gradient(const double *f, int es, double *grad, const skb_grid_t *grid,
                            void *stream) {
  DevGrid g = make_grid(grid);
  gradient_kernel<<<sblk(g), ST, 0, (cudaStream_t)stream>>>(Plane{f, es, g.mx}, grad, g);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_curl(const double *fx, const double *fy, const double *fz, int es,
                        double *curl, const skb_grid_t *grid, int down, void *stream) {
  DevGrid g = make_grid(grid);
  curl_kernel<<<sblk(g), ST, 0, (cudaStream_t)stream>>>(
      Plane{fx, es, g.mx}, Plane{fy, es, g.mx}, Plane{fz, es, g.mx}, curl, g, down);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_divergence(const double *fx, const double *fy, int es, double *div,
                              const skb_grid_t *grid, void *stream) {
  DevGrid g = make_grid(grid);
  divergence_kernel<<<sblk(g), ST, 0, (cudaStream_t)stream>>>(
      Plane{fx, es, g.mx}, Plane{fy, es, g.mx}, div, g);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_interp(const double *fx, const double *fy, const double *fz, int es,
                          double *out, const skb_grid_t *grid, int up, void *stream) {
  DevGrid g = make_grid(grid);
  interp_kernel<<<sblk(g), ST, 0, (cudaStream_t)stream>>>(
      Plane{fx, es, g.mx}, Plane{fy, es, g.mx}, Plane{fz, es, g.mx}, out, g, up);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_ohm(const double *sources, const double *B, double *E, double *Je_out,
                       double *Bc_out, const skb_grid_t *grid, double alpha, double eta,
                       void *stream) {
  DevGrid g = make_grid(grid);
  ohm_kernel<<<sblk(g), ST, 0, (cudaStream_t)stream>>>(sources, B, E, Je_out, Bc_out, g,
                                                       alpha, eta, nullptr);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_faraday(const double *E, double *B, double *dB_out,
                           const skb_grid_t *grid, double dt, void *stream) {
  DevGrid g = make_grid(grid);
  faraday_kernel<<<sblk(g), ST, 0, (cudaStream_t)stream>>>(E, B, B, dB_out, g, dt, nullptr);
  SKB_CHECK_LAUNCH();
  return 0;
}

// ---- field algebra of the time steppers on the device ---------------------------------
// The Horowitz iteration (horowitz.py:138-169) is a handful of whole-field averages
// around one Faraday and one Ohm solve, repeated until the electric field stops changing.
// In the reference every step is a NumPy pass and the convergence test a host decision.
// Here the averages are fused kernels, the residual is reduced on the device, and every
// kernel of an iteration takes a `skip` flag (state[0]) that skb_converged raises - so
// several iterations can be queued without a host round trip; the ones behind the
// converged one do nothing.
__global__ void __launch_bounds__(256)
combine_kernel(double *out, const double *__restrict__ x, const double *__restrict__ y,
               long long n, double c, int mode, const int *skip) {
  if (skip && *skip) return;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  // mode 0: c*(x + y) (E2 = 0.5*(E3 + E), horowitz.py:142); mode 1: -x + c*y
  // (E3 = -E + 2*E2, :159)
  out[i] = mode == 0 ? c * (x[i] + y[i]) : (-x[i]) + c * y[i];
}

// E3 <- -E + 2*E2 over the whole array (horowitz.py:159) and acc[0] += sum over the active
// cells and the three components of (E3_new - E3_old)^2 (calculate_diff, :112-121, before
// the mean / allreduce / sqrt)
__global__ void __launch_bounds__(256)
horowitz_update_kernel(double *E3, const double *__restrict__ E, const double *__restrict__ E2,
                       DevGrid g, double *acc, const int *skip) {
  if (skip && *skip) return;
  const long long ncell = (long long)g.mx * g.myp;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  double d2 = 0.0;
  if (i < ncell) {
    const int iy = (int)(i / g.mx), ix = (int)(i - (long long)iy * g.mx);
    const bool active = iy >= g.lby && iy < g.uby && ix >= g.lbx && ix < g.ubx;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double old = E3[i * 3 + k];
      const double nw = (-E[i * 3 + k]) + 2.0 * E2[i * 3 + k];
      E3[i * 3 + k] = nw;
      if (active) { const double d = nw - old; d2 += d * d; }
    }
  }
  // block sum -> one atomic per block
  __shared__ double part[8];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) d2 += __shfl_down_sync(0xffffffffu, d2, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = d2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; k++) t += part[k];
    if (t != 0.0) atomicAdd(acc, t);
  }
}

__global__ void converged_kernel(const double *acc, double scale, double tol, int iter,
                                 int *state) {
  if (state[0]) return;
  if (sqrt(acc[0] * scale) < tol) { state[0] = 1; state[1] = iter; }
}

extern "C" int skb_field_combine(double *out, const double *x, const double *y, long long n,
                                 double c, int mode, const int *skip, void *stream) {
  if (n <= 0) return 0;
  combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, x, y, n, c,
                                                                                 mode, skip);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_faraday_to(const double *E, const double *Bin, double *Bout, double *dB_out,
                              const skb_grid_t *grid, double dt, const int *skip,
                              void *stream) {
  DevGrid g = make_grid(grid);
  faraday_kernel<<<sblk(g), ST, 0, (cudaStream_t)stream>>>(E, Bin, Bout, dB_out, g, dt, skip);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_ohm_if(const double *sources, const double *B, double *E, double *Je_out,
                          double *Bc_out, const skb_grid_t *grid, double alpha, double eta,
                          const int *skip, void *stream) {
  DevGrid g = make_grid(grid);
  ohm_kernel<<<sblk(g), ST, 0, (cudaStream_t)stream>>>(sources, B, E, Je_out, Bc_out, g,
                                                       alpha, eta, skip);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_horowitz_update(double *E3, const double *E, const double *E2,
                                   const skb_grid_t *grid, double *acc, const int *skip,
                                   void *stream) {
  DevGrid g = make_grid(grid);
  const long long ncell = (long long)g.mx * g.myp;
  horowitz_update_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      E3, E, E2, g, acc, skip);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_converged(const double *acc, double scale, double tol, int iter, int *state,
                             void *stream) {
  converged_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(acc, scale, tol, iter, state);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_version(void) { return 100; }

extern "C" const char *skb_error_string(int err) {
  return cudaGetErrorString((cudaError_t)err);
}
