// ewald_kernels.cu -- smooth particle-mesh Ewald reciprocal-space sum on the device.
//
// Replaces ewald_recip.f90:30-470 for an orthorhombic box (the only kind set_periodic.f90 builds):
//   1. ew_spread_kernel     fractional coordinates, order-5 B-splines (bsplgen.f90, values and first
//                           derivatives) and charge spreading onto the nfft^3 grid (:110-330; the
//                           chunk tables of setchunk.f90 / ewald_adjust.f90 reduce to one chunk)
//   2. ew_fft_lines_kernel  dfftw_execute_dft(planf) (:369): the complex 3-D DFT as three sweeps of 1-D lines, each
//                           line a mixed-radix Stockham transform in shared memory (below); no library FFT
//   3. ew_influence_kernel  exp(-pi^2 h^2 / a^2) / (pi V h^2 B(m)) on every grid point but the
//                           origin, energy = 1/2 sum expterm |S|^2 (:371-415)
//   4. ew_fft_lines_kernel  dfftw_execute_dft(planb) (:417), sign +1, unnormalised like FFTW
//   5. ew_gather_kernel     gradient of every site from the convolved grid (:419-468)
// The reference never reaches this routine (ff_nonb.f90:337 sets ewald=.false.): it is
// exported as its own entry point (crcl_ewald_recip) and validated against the oracle's restatement
// and the plain Ewald reciprocal sum.
//
// Mapping: one warp per (image, atom) for spreading and gathering -- 25 lanes own a (y,z) column of
// the 5x5x5 stencil, FP64 red.global.add to the grid; all images of a call share one batched FFT.
// HBM-bound: the grid (16 B per point) is written by the spread, read+written by each FFT pass and
// by the influence kernel, read by the gather.
#include <cmath>
#include <cstring>
#include <vector>
#include "ewald.cuh"

namespace crcl {

constexpr int BSO = 5;   // bsorder of set_periodic.f90:208

// bsplgen.f90:30-98 for bsorder = 5, level = 2: th[i] = value, dth[i] = first derivative, i = 0..4
__device__ __forceinline__ void bsplgen5(double w, double th[BSO], double dth[BSO])
{
    double b[BSO + 1][BSO + 1];   // bsbuild(i,j), 1-based
    b[2][2] = w;
    b[2][1] = 1.0 - w;
    b[3][3] = 0.5 * w * b[2][2];
    b[3][2] = 0.5 * ((1.0 + w) * b[2][1] + (2.0 - w) * b[2][2]);
    b[3][1] = 0.5 * (1.0 - w) * b[2][1];
#pragma unroll
    for (int i = 4; i <= BSO; i++) {
        const int k = i - 1;
        const double denom = 1.0 / (double)k;
        b[i][i] = denom * w * b[k][k];
#pragma unroll
        for (int j = 1; j <= i - 2; j++)
            b[i][i - j] = denom * ((w + (double)j) * b[k][i - j - 1] + ((double)(i - j) - w) * b[k][i - j]);
        b[i][1] = denom * (1.0 - w) * b[k][1];
    }
    constexpr int k = BSO - 1;
    b[k][BSO] = b[k][BSO - 1];
#pragma unroll
    for (int i = BSO - 1; i >= 2; i--) b[k][i] = b[k][i - 1] - b[k][i];
    b[k][1] = -b[k][1];
#pragma unroll
    for (int i = 1; i <= BSO; i++) {
        th[i - 1] = b[BSO][i];
        dth[i - 1] = b[BSO - 1][i];
    }
}

// grid index (0-based, wrapped) of stencil point it = 0..4 of an atom whose igrid value is g0:
// the reference's 1-based index is g0 + it + 2 for g0 + it + 1 >= 0 and nfft more otherwise
// (ewald_recip.f90:436-448; the spreading loop :300-325 addresses the same points)
__device__ __forceinline__ int wrap(int g0, int it, int nfft)
{
    const int k0 = g0 + it + 1;
    return (k0 >= 0) ? k0 : k0 + nfft;
}

__global__ void __launch_bounds__(128) ew_spread_kernel(int n, int natot, int nfft, double rx, double ry, double rz,
                                                        const double* __restrict__ xyz, const double* __restrict__ q,
                                                        double2* __restrict__ grid,
                                                        double* __restrict__ theta, int* __restrict__ igrid)
{
    const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (a >= natot) return;
    const int img = a / n, at = a - img * n;
    const double eps = 1.0e-8;
    double th[3][BSO], dth[3][BSO];
    int g0[3];
    const double rec[3] = {rx, ry, rz};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double w = xyz[3 * (size_t)a + d] * rec[d];
        const double fr = (double)nfft * (w - round(w) + 0.5);   // anint
        const int ifr = (int)(fr - eps);
        w = fr - (double)ifr;
        g0[d] = ifr - BSO;
        bsplgen5(w, th[d], dth[d]);
    }
    if (lane < 3 * BSO) {
        const int d = lane / BSO, i = lane - d * BSO;
        double* t = theta + (size_t)a * 3 * BSO * 2;
        t[(d * BSO + i) * 2] = (d == 0) ? th[0][i] : (d == 1 ? th[1][i] : th[2][i]);
        t[(d * BSO + i) * 2 + 1] = (d == 0) ? dth[0][i] : (d == 1 ? dth[1][i] : dth[2][i]);
    }
    if (lane < 3) igrid[3 * (size_t)a + lane] = (lane == 0) ? g0[0] : (lane == 1 ? g0[1] : g0[2]);
    if (lane < BSO * BSO) {
        const int kz = lane / BSO, jy = lane - kz * BSO;
        double v0 = 0.0, u0 = 0.0;
#pragma unroll
        for (int i = 0; i < BSO; i++) {
            if (i == kz) v0 = th[2][i];
            if (i == jy) u0 = th[1][i];
        }
        const double term = (v0 * q[at]) * u0;
        const int k = wrap(g0[2], kz, nfft), j = wrap(g0[1], jy, nfft);
        double2* row = grid + (((size_t)img * nfft + k) * nfft + j) * nfft;
#pragma unroll
        for (int i = 0; i < BSO; i++) atomicAdd(&row[wrap(g0[0], i, nfft)].x, term * th[0][i]);
    }
}

__global__ void __launch_bounds__(256) ew_influence_kernel(int nfft, int nimg, double rx, double ry, double rz,
                                                           double pterm, double volterm,
                                                           const double* __restrict__ bsmod,
                                                           double2* __restrict__ grid,
                                                           double* __restrict__ energy)
{
    const size_t npoint = (size_t)nfft * nfft * nfft;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int img = blockIdx.y;
    double e = 0.0;
    if (t >= 1 && t < npoint) {   // the origin keeps its value, as in the reference (:380)
        const int nff = nfft * nfft, nf = (nfft + 1) / 2;
        const int k3 = (int)(t / nff), jr = (int)(t - (size_t)k3 * nff), k2 = jr / nfft, k1 = jr - k2 * nfft;
        const int m1 = (k1 + 1 > nf) ? k1 - nfft : k1, m2 = (k2 + 1 > nf) ? k2 - nfft : k2,
                  m3 = (k3 + 1 > nf) ? k3 - nfft : k3;
        const double h1 = rx * (double)m1, h2 = ry * (double)m2, h3 = rz * (double)m3;
        const double hsq = h1 * h1 + h2 * h2 + h3 * h3;
        const double term = -pterm * hsq;
        double expterm = 0.0;
        double2* p = grid + (size_t)img * npoint + t;
        double2 v = *p;
        if (term > -50.0) {
            const double denom = volterm * hsq * bsmod[k1] * bsmod[nfft + k2] * bsmod[2 * nfft + k3];
            expterm = exp(term) / denom;
            e = 0.5 * expterm * (v.x * v.x + v.y * v.y);
        }
        v.x *= expterm;
        v.y *= expterm;
        *p = v;
    }
    __shared__ double sh[8];
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += sh[w];
        if (s != 0.0) atomicAdd(&energy[img], s);
    }
}

__global__ void __launch_bounds__(128) ew_gather_kernel(int n, int natot, int nfft, double rx, double ry, double rz,
                                                        const double* __restrict__ q,
                                                        const double2* __restrict__ grid,
                                                        const double* __restrict__ theta, const int* __restrict__ igrid,
                                                        double* __restrict__ grad)
{
    const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (a >= natot) return;
    const int img = a / n, at = a - img * n;
    const double* t = theta + (size_t)a * 3 * BSO * 2;
    const int g0x = igrid[3 * (size_t)a], g0y = igrid[3 * (size_t)a + 1], g0z = igrid[3 * (size_t)a + 2];
    const double dn = (double)nfft;
    double de1 = 0.0, de2 = 0.0, de3 = 0.0;
    if (lane < BSO * BSO) {
        const int kz = lane / BSO, jy = lane - kz * BSO;
        const double t3 = t[(2 * BSO + kz) * 2], dt3 = dn * t[(2 * BSO + kz) * 2 + 1];
        const double t2 = t[(BSO + jy) * 2], dt2 = dn * t[(BSO + jy) * 2 + 1];
        const double2* row =
            grid + (((size_t)img * nfft + wrap(g0z, kz, nfft)) * nfft + wrap(g0y, jy, nfft)) * nfft;
#pragma unroll
        for (int i = 0; i < BSO; i++) {
            const double t1 = t[i * 2], dt1 = dn * t[i * 2 + 1];
            const double term = row[wrap(g0x, i, nfft)].x;
            de1 += term * dt1 * t2 * t3;
            de2 += term * dt2 * t1 * t3;
            de3 += term * dt3 * t1 * t2;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        de1 += __shfl_xor_sync(0xffffffffu, de1, o);
        de2 += __shfl_xor_sync(0xffffffffu, de2, o);
        de3 += __shfl_xor_sync(0xffffffffu, de3, o);
    }
    if (lane == 0) {
        const double fi = q[at];
        grad[3 * (size_t)a] = fi * rx * de1;
        grad[3 * (size_t)a + 1] = fi * ry * de2;
        grad[3 * (size_t)a + 2] = fi * rz * de3;
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Complex 3-D DFT of the charge grid (the two dfftw_execute_dft calls, ewald_recip.f90:369,417): three sweeps, one
// per axis; a CTA owns a tile of L lines that are adjacent in memory along x (for the y and z sweeps: L consecutive x
// at one (z, img) or (y, img), so a global access covers L*16 contiguous bytes; for the x sweep: L consecutive
// lines), keeps them in shared memory as sh[t*LP + line] and runs a decimation-in-frequency Stockham autosort
// transform over the factors of nfft (any factorisation: a thread produces ONE output of a radix-r stage from its r
// inputs, twiddles from a table of the nfft-th roots of unity computed in double on the host).  HBM traffic: one read
// and one write of the grid per sweep.
__global__ void __launch_bounds__(256) ew_fft_lines_kernel(double2* __restrict__ grid, int nf, int nimg, int dim, int L,
                                                           int LP, const double2* __restrict__ tw_g, FftPlan plan)
{
    extern __shared__ double2 ew_sh[];
    double2* tw = ew_sh;                       // [nf]
    double2* x = ew_sh + nf;                   // [nf][LP]
    double2* y = x + (size_t)nf * LP;          // [nf][LP]
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t nf2 = (size_t)nf * nf;
    // tile -> first element and number of valid lines
    size_t base, lstride, tstride;
    int nl;
    if (dim == 0) {
        const size_t line0 = (size_t)blockIdx.x * L;          // lines of the whole batch are consecutive
        const size_t nlines = nf2 * nimg;
        if (line0 >= nlines) return;
        nl = (int)((nlines - line0 < (size_t)L) ? nlines - line0 : L);
        base = line0 * nf;
        lstride = nf;
        tstride = 1;
    } else {
        const int tiles_i = (nf + L - 1) / L;
        const int i0 = (blockIdx.x % tiles_i) * L, o = blockIdx.x / tiles_i;   // o = z (dim 1) or y (dim 2)
        nl = (nf - i0 < L) ? nf - i0 : L;
        base = (dim == 1) ? (size_t)o * nf2 + i0 : (size_t)o * nf + i0;
        lstride = 1;
        tstride = (dim == 1) ? (size_t)nf : nf2;
    }
    if (dim != 0) base += (size_t)blockIdx.y * nf2 * nf;       // image
    for (int t = tid; t < nf; t += nt) tw[t] = tw_g[t];
    if (dim == 0) {
        for (int e = tid; e < nl * nf; e += nt) {
            const int line = e / nf, t = e - line * nf;
            x[(size_t)t * LP + line] = grid[base + (size_t)line * lstride + t];
        }
    } else {
        for (int e = tid; e < nl * nf; e += nt) {
            const int t = e / nl, line = e - t * nl;
            x[(size_t)t * LP + line] = grid[base + (size_t)t * tstride + line];
        }
    }
    __syncthreads();
    int n = nf, st = 1;
    for (int f = 0; f < plan.nfac; f++) {
        const int r = plan.r[f], m = n / r, wr = nf / r;
        for (int e = tid; e < nl * nf; e += nt) {
            const int o = e / nl, line = e - o * nl;
            const int q = o % st, tt = o / st, j = tt % r, pp = tt / r;
            double re = 0.0, im = 0.0;
            int jk = 0;                                        // (j*k) mod r
            for (int k = 0; k < r; k++) {
                const double2 a = x[(size_t)(q + st * (pp + k * m)) * LP + line];
                const double2 w = tw[jk * wr];
                re += a.x * w.x - a.y * w.y;
                im += a.x * w.y + a.y * w.x;
                jk += j;
                if (jk >= r) jk -= r;
            }
            const double2 w = tw[pp * j * st];                 // < nf: pp < m, j < r, m*r*st = nf
            y[(size_t)o * LP + line] = make_double2(re * w.x - im * w.y, re * w.y + im * w.x);
        }
        __syncthreads();
        double2* z = x;
        x = y;
        y = z;
        n = m;
        st *= r;
    }
    if (dim == 0) {
        for (int e = tid; e < nl * nf; e += nt) {
            const int line = e / nf, t = e - line * nf;
            grid[base + (size_t)line * lstride + t] = x[(size_t)t * LP + line];
        }
    } else {
        for (int e = tid; e < nl * nf; e += nt) {
            const int t = e / nl, line = e - t * nl;
            grid[base + (size_t)t * tstride + line] = x[(size_t)t * LP + line];
        }
    }
}

// one unnormalised 3-D transform of every image's grid; sgn = -1 (FFTW_FORWARD) or +1 (FFTW_BACKWARD)
static int ew_fft3(EwaldDev* E, int nimg, int sgn, cudaStream_t s, long long* launches)
{
    const int nf = E->nfft;
    const double2* tw = E->tw + (sgn < 0 ? 0 : nf);
    for (int dim = 0; dim < 3; dim++) {
        const int L = E->fft_L, LP = L | 1;
        const size_t shm = ((size_t)nf + 2 * (size_t)nf * LP) * sizeof(double2);
        // x sweep: the lines of the whole batch are consecutive; y / z sweeps: blockIdx.y is the image
        const dim3 grid = (dim == 0) ? dim3((unsigned)(((size_t)nf * nf * nimg + L - 1) / L), 1)
                                     : dim3((unsigned)(nf * ((nf + L - 1) / L)), nimg);
        ew_fft_lines_kernel<<<grid, 256, shm, s>>>(E->grid, nf, nimg, dim, L, LP, tw, E->plan);
        if (launches) *launches += 1;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int ewald_recip(EwaldDev* E, int n, int nimg, const double* d_xyz, const double* d_q, double* d_energy, double* d_grad,
                cudaStream_t s, long long* launches, const char** err)
{
    if (n <= 0 || nimg <= 0) return CRCL_OK;
    const int nfft = E->nfft;
    const size_t npoint = (size_t)nfft * nfft * nfft, natot = (size_t)n * nimg;
    if (natot > (size_t)1 << 26) {
        *err = "crcl_ewald_recip: more than 2^26 (image, atom) sites in one call";
        return CRCL_EINVAL;
    }
    if (npoint * nimg > E->grid_cap) {
        if (E->grid) cudaFree(E->grid);
        E->grid = nullptr;
        E->grid_cap = 0;
        if (cudaMalloc(&E->grid, npoint * nimg * sizeof(double2)) != cudaSuccess) {
            *err = "crcl_ewald_recip: grid allocation failed";
            return CRCL_ENOMEM;
        }
        E->grid_cap = npoint * nimg;
    }
    if (natot > E->atom_cap) {
        if (E->theta) cudaFree(E->theta);
        if (E->igrid) cudaFree(E->igrid);
        E->theta = nullptr;
        E->igrid = nullptr;
        E->atom_cap = 0;
        if (cudaMalloc(&E->theta, natot * 3 * BSO * 2 * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&E->igrid, natot * 3 * sizeof(int)) != cudaSuccess) {
            *err = "crcl_ewald_recip: B-spline work space allocation failed";
            return CRCL_ENOMEM;
        }
        E->atom_cap = natot;
    }
    const double volbox = E->box[0] * E->box[1] * E->box[2];
    // recip(1,1) = (br2*cr3)/volbox etc. (ewald_recip.f90:97-105): the division is kept as written
    const double rx = (E->box[1] * E->box[2]) / volbox, ry = (E->box[2] * E->box[0]) / volbox,
                 rz = (E->box[0] * E->box[1]) / volbox;
    const double PI = 3.1415926535897932384626433832795029;
    const double pterm = (PI / E->a_ewald) * (PI / E->a_ewald), volterm = PI * volbox;
    cudaMemsetAsync(E->grid, 0, npoint * nimg * sizeof(double2), s);
    cudaMemsetAsync(d_energy, 0, nimg * sizeof(double), s);
    const unsigned wblocks = (unsigned)((natot * 32 + 127) / 128);
    ew_spread_kernel<<<wblocks, 128, 0, s>>>(n, (int)natot, nfft, rx, ry, rz, d_xyz, d_q, E->grid, E->theta, E->igrid);
    if (ew_fft3(E, nimg, -1, s, launches)) {
        *err = "crcl_ewald_recip: forward FFT launch failed";
        return CRCL_ECUDA;
    }
    ew_influence_kernel<<<dim3((unsigned)((npoint + 255) / 256), nimg), 256, 0, s>>>(nfft, nimg, rx, ry, rz, pterm, volterm,
                                                                                    E->bsmod, E->grid, d_energy);
    if (ew_fft3(E, nimg, +1, s, launches)) {
        *err = "crcl_ewald_recip: backward FFT launch failed";
        return CRCL_ECUDA;
    }
    ew_gather_kernel<<<wblocks, 128, 0, s>>>(n, (int)natot, nfft, rx, ry, rz, d_q, E->grid, E->theta, E->igrid, d_grad);
    if (launches) *launches += 3;   // spread, influence, gather; ew_fft3 counts its six sweeps
    if (cudaGetLastError() != cudaSuccess) {
        *err = "crcl_ewald_recip: kernel launch failed";
        return CRCL_ECUDA;
    }
    return CRCL_OK;
}

int ewald_upload(const crcl_ewald_params* P, EwaldDev** out, const char** err)
{
    *out = nullptr;
    if (P->bsorder != BSO) {
        *err = "SPME: bsorder must be 5 (set_periodic.f90:208)";
        return CRCL_ENOSUP;
    }
    if (P->nfft < 2 * BSO || P->nfft > 864 || !(P->a_ewald > 0.0) || !P->bsmod1 || !P->bsmod2 || !P->bsmod3 ||
        !(P->box[0] > 0.0) || !(P->box[1] > 0.0) || !(P->box[2] > 0.0)) {
        *err = "SPME: nfft in 10..864, a_ewald > 0, box > 0 and the three bsmod tables are required";
        return CRCL_EINVAL;
    }
    EwaldDev* E = new EwaldDev();
    memset(E, 0, sizeof(*E));
    E->nfft = P->nfft;
    E->bsorder = P->bsorder;
    E->a_ewald = P->a_ewald;
    for (int d = 0; d < 3; d++) E->box[d] = P->box[d];
    std::vector<double> bs(3 * (size_t)P->nfft);
    for (int i = 0; i < P->nfft; i++) {
        bs[i] = P->bsmod1[i];
        bs[P->nfft + i] = P->bsmod2[i];
        bs[2 * P->nfft + i] = P->bsmod3[i];
    }
    if (cudaMalloc(&E->bsmod, bs.size() * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(E->bsmod, bs.data(), bs.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
        ewald_free(E);
        *err = "SPME: device allocation / upload failed";
        return CRCL_ENOMEM;
    }
    // roots of unity e^{-2 pi i t / nfft} (forward) and their conjugates (backward), the factors of nfft (largest
    // first: fewer stages), and the number of lines a CTA holds in its two shared-memory line buffers
    {
        const int nf = P->nfft;
        const double PI = 3.1415926535897932384626433832795029;
        std::vector<double2> tw(2 * (size_t)nf);
        for (int t = 0; t < nf; t++) {
            const double a = 2.0 * PI * (double)t / (double)nf;
            tw[t] = make_double2(cos(a), -sin(a));
            tw[nf + t] = make_double2(cos(a), sin(a));
        }
        int rest = nf, nfac = 0;
        const int pref[4] = {5, 4, 3, 2};
        for (int i = 0; i < 4; i++)
            while (rest % pref[i] == 0 && nfac < 16) {
                E->plan.r[nfac++] = pref[i];
                rest /= pref[i];
            }
        for (int r = 7; rest > 1 && nfac < 16; r += 2)
            while (rest % r == 0 && nfac < 16) {
                E->plan.r[nfac++] = r;
                rest /= r;
            }
        if (rest != 1) {
            ewald_free(E);
            *err = "SPME: nfft has more than 16 prime factors";
            return CRCL_EINVAL;
        }
        E->plan.nfac = nfac;
        int L = 8;
        while (L > 1 && ((size_t)nf + 2 * (size_t)nf * (L | 1)) * sizeof(double2) > 200 * 1024) L >>= 1;
        E->fft_L = L;
        const size_t shm = ((size_t)nf + 2 * (size_t)nf * (L | 1)) * sizeof(double2);
        if (cudaFuncSetAttribute(ew_fft_lines_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm) !=
                cudaSuccess ||
            cudaMalloc(&E->tw, tw.size() * sizeof(double2)) != cudaSuccess ||
            cudaMemcpy(E->tw, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice) != cudaSuccess) {
            ewald_free(E);
            *err = "SPME: FFT set-up failed";
            return CRCL_ENOMEM;
        }
    }
    *out = E;
    return CRCL_OK;
}

void ewald_free(EwaldDev* E)
{
    if (!E) return;
    cudaFree(E->bsmod);
    cudaFree(E->tw);
    cudaFree(E->grid);
    cudaFree(E->theta);
    cudaFree(E->igrid);
    cudaFree(E->dq);
    delete E;
}

}  // namespace crcl
