// ewald_kernels.cu -- smooth particle-mesh Ewald reciprocal-space sum on the device.
//
// Replaces ewald_recip.f90:30-470 for an orthorhombic box (the only kind set_periodic.f90 builds):
//   1. ew_spread_kernel     fractional coordinates, order-5 B-splines (bsplgen.f90, values and first
//                           derivatives) and charge spreading onto the nfft^3 grid (:110-330; the
//                           chunk tables of setchunk.f90 / ewald_adjust.f90 reduce to one chunk)
//   2. cuFFT Z2Z forward    dfftw_execute_dft(planf) (:369); cuFFT is a library FFT like FFTW there
//   3. ew_influence_kernel  exp(-pi^2 h^2 / a^2) / (pi V h^2 B(m)) on every grid point but the
//                           origin, energy = 1/2 sum expterm |S|^2 (:371-415)
//   4. cuFFT Z2Z inverse    dfftw_execute_dft(planb) (:417), unnormalised like FFTW
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
                                                        cufftDoubleComplex* __restrict__ grid,
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
        cufftDoubleComplex* row = grid + (((size_t)img * nfft + k) * nfft + j) * nfft;
#pragma unroll
        for (int i = 0; i < BSO; i++) atomicAdd(&row[wrap(g0[0], i, nfft)].x, term * th[0][i]);
    }
}

__global__ void __launch_bounds__(256) ew_influence_kernel(int nfft, int nimg, double rx, double ry, double rz,
                                                           double pterm, double volterm,
                                                           const double* __restrict__ bsmod,
                                                           cufftDoubleComplex* __restrict__ grid,
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
        cufftDoubleComplex* p = grid + (size_t)img * npoint + t;
        cufftDoubleComplex v = *p;
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
                                                        const cufftDoubleComplex* __restrict__ grid,
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
        const cufftDoubleComplex* row =
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
        if (cudaMalloc(&E->grid, npoint * nimg * sizeof(cufftDoubleComplex)) != cudaSuccess) {
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
    if (E->plan_batch != nimg) {
        if (E->plan_batch) cufftDestroy(E->plan);
        E->plan_batch = 0;
        int dims[3] = {nfft, nfft, nfft};
        if (cufftPlanMany(&E->plan, 3, dims, nullptr, 1, (int)npoint, nullptr, 1, (int)npoint, CUFFT_Z2Z, nimg) !=
            CUFFT_SUCCESS) {
            *err = "crcl_ewald_recip: cufftPlanMany failed";
            return CRCL_ECUDA;
        }
        E->plan_batch = nimg;
    }
    if (cufftSetStream(E->plan, s) != CUFFT_SUCCESS) {
        *err = "crcl_ewald_recip: cufftSetStream failed";
        return CRCL_ECUDA;
    }
    const double volbox = E->box[0] * E->box[1] * E->box[2];
    // recip(1,1) = (br2*cr3)/volbox etc. (ewald_recip.f90:97-105): the division is kept as written
    const double rx = (E->box[1] * E->box[2]) / volbox, ry = (E->box[2] * E->box[0]) / volbox,
                 rz = (E->box[0] * E->box[1]) / volbox;
    const double PI = 3.1415926535897932384626433832795029;
    const double pterm = (PI / E->a_ewald) * (PI / E->a_ewald), volterm = PI * volbox;
    cudaMemsetAsync(E->grid, 0, npoint * nimg * sizeof(cufftDoubleComplex), s);
    cudaMemsetAsync(d_energy, 0, nimg * sizeof(double), s);
    const unsigned wblocks = (unsigned)((natot * 32 + 127) / 128);
    ew_spread_kernel<<<wblocks, 128, 0, s>>>(n, (int)natot, nfft, rx, ry, rz, d_xyz, d_q, E->grid, E->theta, E->igrid);
    if (cufftExecZ2Z(E->plan, E->grid, E->grid, CUFFT_FORWARD) != CUFFT_SUCCESS) {
        *err = "crcl_ewald_recip: forward FFT failed";
        return CRCL_ECUDA;
    }
    ew_influence_kernel<<<dim3((unsigned)((npoint + 255) / 256), nimg), 256, 0, s>>>(nfft, nimg, rx, ry, rz, pterm, volterm,
                                                                                    E->bsmod, E->grid, d_energy);
    if (cufftExecZ2Z(E->plan, E->grid, E->grid, CUFFT_INVERSE) != CUFFT_SUCCESS) {
        *err = "crcl_ewald_recip: backward FFT failed";
        return CRCL_ECUDA;
    }
    ew_gather_kernel<<<wblocks, 128, 0, s>>>(n, (int)natot, nfft, rx, ry, rz, d_q, E->grid, E->theta, E->igrid, d_grad);
    if (launches) *launches += 3;   // own kernels; the two FFTs are library launches
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
    *out = E;
    return CRCL_OK;
}

void ewald_free(EwaldDev* E)
{
    if (!E) return;
    if (E->plan_batch) cufftDestroy(E->plan);
    cudaFree(E->bsmod);
    cudaFree(E->grid);
    cudaFree(E->theta);
    cudaFree(E->igrid);
    cudaFree(E->dq);
    delete E;
}

}  // namespace crcl
