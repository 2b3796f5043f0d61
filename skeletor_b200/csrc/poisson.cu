// poisson.cu — k-space part of the electrostatic field solve E = grad del^-2 rho.
//
// Restates calc_form_factors + grad_inv_del (reference skeletor/cython/operators.pyx:13-135;
// ppic2's cppois22, ppic2_wrapper.pyx:100-116) on the layout of a cuFFT real-to-complex
// transform: q[ky][kx], kx = 0 .. nx/2, ky = 0 .. ny-1 (wrapped), complex128.  The forward
// and inverse transforms (ppic2's cwppfft2r / cwppfft2r2 with their MPI transposes in the
// reference) are cuFFT calls made by the caller.
//
// Quirk Q3: operators.pyx reads the form factors and the charge
// spectrum through crealf / cimagf, i.e. truncated to float32; `float32_quirk` reproduces
// that (parity with the reference at its own ~1e-7 level), 0 keeps full double precision.
#include "common.cuh"

__global__ void __launch_bounds__(256)
poisson_kspace_kernel(const double2 *__restrict__ q, double2 *fx, double2 *fy, int nx, int ny,
                      double dnx, double dny, double ax, double ay, double affp, int quirk,
                      double *we) {
  const int nxh = nx / 2, nyh = max(1, ny / 2);
  const int ncol = nxh + 1;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  double w = 0.0;
  if (i < (long long)ncol * ny) {
    const int k = (int)(i / ncol), j = (int)(i - (long long)k * ncol);
    const int ks = k <= ny / 2 ? k : k - ny;            // signed mode number along y
    const double dkx = dnx * (double)j, dky = dny * (double)ks;
    // form factors, operators.pyx:38-49
    const double at3 = dky * dky + dkx * dkx;
    const double dax = dkx * ax, day = dky * ay;
    const double at4 = exp(-.5 * (day * day + dax * dax));
    double re = at3 == 0.0 ? affp : affp * at4 / at3;
    double im = at3 == 0.0 ? 1.0 : at4;
    double at1 = quirk ? (double)((float)re * (float)im) : re * im;     // :86, 97
    double2 qq = q[i];
    if (quirk) { qq.x = (double)(float)qq.x; qq.y = (double)(float)qq.y; }
    // modes the reference zeroes: kx = ky = 0, kx = nx/2, ky = ny/2 (:100-101, 116-133)
    const bool dead = (j == 0 && k == 0) || j == nxh || (ny > 1 && k == nyh);
    if (dead) at1 = 0.0;
    // zt = imag(q) - i real(q), E_k = k S(k)/k^2 zt   (:88-96)
    const double2 zt = make_double2(qq.y, -qq.x);
    const double a2 = dkx * at1, a3 = dky * at1;
    fx[i] = make_double2(a2 * zt.x, a2 * zt.y);
    fy[i] = make_double2(a3 * zt.x, a3 * zt.y);
    // field energy: every mode of the half spectrum the reference visits once
    // (kx > 0: all ky; kx = 0: ky > 0), :97-98, 108, 123
    if (!dead && (j > 0 || ks > 0)) w = at1 * (double)(float)(qq.x * qq.x + qq.y * qq.y);
  }
  __shared__ double part[8];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0 && we) {
    double t = 0.0;
    for (int a = 0; a < 8; a++) t += part[a];
    // operators.pyx:135 returns wp*nx*ny on ppic2's spectrum, which carries the 1/(nx ny)
    // of its forward transform; on the unnormalised cuFFT spectrum that is sum/(nx ny)
    if (t != 0.0) atomicAdd(we, t / ((double)nx * (double)ny));
  }
}

// q, fx, fy: [ny][nx/2 + 1] complex128 (interleaved re, im); we: device double, zeroed by
// the caller (may be NULL)
extern "C" int skb_poisson_kspace(const double *q, double *fx, double *fy, int nx, int ny,
                                  double Lx, double Ly, double ax, double ay, double affp,
                                  int float32_quirk, double *we, void *stream) {
  if (nx < 2 || ny < 1) return (int)cudaErrorInvalidValue;
  const long long n = (long long)(nx / 2 + 1) * ny;
  const double pi = 3.14159265358979323846;
  poisson_kspace_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const double2 *)q, (double2 *)fx, (double2 *)fy, nx, ny, 2.0 * pi / Lx, 2.0 * pi / Ly, ax,
      ay, affp, float32_quirk, we);
  SKB_CHECK_LAUNCH();
  return 0;
}
