// water_kernels.cu -- flexible SPC water box (pes WATER_SPC) on the device.
//
// Replaces egrad_water.f90:36-333: per molecule two Morse O-H stretches, a harmonic H-H term and the two
// stretch-stretch / stretch-bend couplings (:64-205); between molecules the Coulomb term (plain, or Zahn's damped
// shifted form under periodic boundaries, :270-276) and the O-O Lennard-Jones term (:295-301) over all atom pairs
// i < j of different molecules -- O(N^2) in the reference, once per bead and step.
//
// B200 mapping (HBM / FP64-latency bound integer-free pair work, no GEMM shape anywhere):
//   wat_intra_kernel   one thread per (image, molecule); it OWNS the three gradient rows of its molecule and
//                      stores them (no memset of g needed), energy by warp reduction + one FP64 red per warp;
//   wat_inter_kernel   one warp per (image, atom i): the lanes sweep j > i in chunks of 32 (coalesced rows of the
//                      AoS structure), minimum image by the reference's own loop, exact `r < cut` test; the force on
//                      i stays in registers and is reduced over the warp once, the reaction goes to j with FP64
//                      red.global.add.  Heavy rows (small i) are issued first.
// Reproduced as written: the Lennard-Jones block tests name(i) twice (:295), so every pair whose first atom is an
// oxygen gets the O-O term; the Zahn gradient reuses e0 / r^2 (:279); no cut-off without periodic boundaries (:268).
#include <cmath>
#include "water.cuh"

namespace crcl {

__device__ __forceinline__ double wat_image(double v, double L, double L2)
{
    while (fabs(v) > L2) v = v - (v >= 0 ? L : -L);   // box_image.f90
    return v;
}

__global__ void __launch_bounds__(128) wat_intra_kernel(const WaterDev W, const double* __restrict__ xyz, int nimg,
                                                        double* __restrict__ V, double* __restrict__ g)
{
    const int nw = W.n / 3, m = blockIdx.x * blockDim.x + threadIdx.x, img = blockIdx.y;   // a warp never straddles images
    double e = 0.0;
    if (m < nw) {
        const double* x = xyz + ((size_t)img * W.n + 3 * m) * 3;
        double* gm = g + ((size_t)img * W.n + 3 * m) * 3;
        const double r_zero = W.pars[0], r_0HH = W.pars[1], d_e = W.pars[2], a_par = W.pars[3], k_theta = W.pars[4],
                     k_rtheta = W.pars[5], k_rr = W.pars[6];
        double a[3], b[3], c[3];   // O-H1, O-H2, H1-H2
#pragma unroll
        for (int d = 0; d < 3; d++) {
            a[d] = x[d] - x[3 + d];
            b[d] = x[d] - x[6 + d];
            c[d] = x[3 + d] - x[6 + d];
            if (W.periodic) {
                a[d] = wat_image(a[d], W.box[d], 0.5 * W.box[d]);
                b[d] = wat_image(b[d], W.box[d], 0.5 * W.box[d]);
                c[d] = wat_image(c[d], W.box[d], 0.5 * W.box[d]);
            }
        }
        const double r1 = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), r2 = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]),
                     rh = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        const double d1 = r1 - r_zero, d2 = r2 - r_zero, dh = rh - r_0HH;
        const double x1 = exp(a_par * d1), x2 = exp(a_par * d2);
        e = d_e * ((1.0 - x1) * (1.0 - x1)) + d_e * ((1.0 - x2) * (1.0 - x2)) + 0.5 * k_theta * (dh * dh) +
            k_rtheta * dh * (d1 + d2) + k_rr * d1 * d2;
        // dE/dr1, dE/dr2, dE/drHH collected (:75-205 adds them term by term), then one pass over the three atoms
        const double f1 = (-2.0 * a_par * d_e * x1 * (1.0 - x1) + k_rtheta * dh + k_rr * d2) / r1;
        const double f2 = (-2.0 * a_par * d_e * x2 * (1.0 - x2) + k_rtheta * dh + k_rr * d1) / r2;
        const double fh = (k_theta * dh + k_rtheta * (d1 + d2)) / rh;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            gm[d] = f1 * a[d] + f2 * b[d];
            gm[3 + d] = -f1 * a[d] + fh * c[d];
            gm[6 + d] = -f2 * b[d] - fh * c[d];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&V[img], e);
}

// One pair (i, j) inside the cut-off: Coulomb (+ Lennard-Jones when atom i is an oxygen); adds to the warp lane's
// partial force on i and energy, sends the reaction to j.
__device__ __forceinline__ void wat_pair(const WaterDev& W, const double* __restrict__ x, double* __restrict__ gi, int j,
                                         double xi0, double xi1, double xi2, double qi, bool oi, double& f0, double& f1,
                                         double& f2, double& e)
{
    double d0 = xi0 - x[3 * j], d1 = xi1 - x[3 * j + 1], d2 = xi2 - x[3 * j + 2];
    if (W.periodic) {
        d0 = wat_image(d0, W.box[0], 0.5 * W.box[0]);
        d1 = wat_image(d1, W.box[1], 0.5 * W.box[1]);
        d2 = wat_image(d2, W.box[2], 0.5 * W.box[2]);
    }
    const double r = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    const double oner = 1.0 / r;
    double e0;
    if (W.zahn)
        e0 = qi * W.q[j] * ((erfc(W.zahn_a * r) * oner) - W.zahn_par * (r - W.coul_cut));
    else
        e0 = qi * W.q[j] * oner;
    double s = e0 * oner * oner;
    e += e0;
    if (oi) {
        const double sig = W.pars[9], eps = W.pars[10];
        const double s2 = (sig * oner) * (sig * oner), s6 = s2 * s2 * s2;
        e += 4.0 * eps * (s6 * s6 - s6);
        s += 24.0 * eps * oner * oner * s6 * (2.0 * s6 - 1.0);
    }
    const double g0 = s * d0, g1 = s * d1, g2 = s * d2;
    f0 -= g0;
    f1 -= g1;
    f2 -= g2;
    atomicAdd(&gi[3 * j], g0);
    atomicAdd(&gi[3 * j + 1], g1);
    atomicAdd(&gi[3 * j + 2], g2);
}

// The sweep over j > i is split in a cheap test (minimum image, r < cut: ~14 % of the pairs of a 31 A box survive) and
// the expensive evaluation (erfc, divisions).  Survivors are compacted into a per-warp queue (ballot + popc) and
// evaluated 32 at a time with all lanes busy; without the queue nearly every warp iteration paid for the expensive
// path because some lane is always inside the cut-off (same scheme as qm_inter_kernel).
__global__ void __launch_bounds__(128) wat_inter_kernel(const WaterDev W, const double* __restrict__ xyz, int nimg,
                                                        double* __restrict__ V, double* __restrict__ g)
{
    __shared__ int queue[4][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int i = blockIdx.x * (blockDim.x >> 5) + wib, img = blockIdx.y;
    if (i >= W.n - 1) return;                       // whole warps leave together; the last atom has no j > i
    const double* x = xyz + (size_t)img * W.n * 3;
    double* gi = g + (size_t)img * W.n * 3;
    const double xi0 = x[3 * i], xi1 = x[3 * i + 1], xi2 = x[3 * i + 2], qi = W.q[i];
    const bool oi = W.is_O[i] != 0;
    const int mi = i / 3;
    int* q = queue[wib];
    int cnt = 0;                                    // warp-uniform
    double f0 = 0.0, f1 = 0.0, f2 = 0.0, e = 0.0;
    const double c2 = W.coul_cut * W.coul_cut, c2lo = c2 * (1.0 - 1e-12), c2hi = c2 * (1.0 + 1e-12);
    for (int j0 = i + 1; j0 < W.n; j0 += 32) {
        const int j = j0 + lane;
        bool in = false;
        if (j < W.n && j / 3 != mi) {
            double d0 = xi0 - x[3 * j], d1 = xi1 - x[3 * j + 1], d2 = xi2 - x[3 * j + 2];
            if (W.periodic) {
                d0 = wat_image(d0, W.box[0], 0.5 * W.box[0]);
                d1 = wat_image(d1, W.box[1], 0.5 * W.box[1]);
                d2 = wat_image(d2, W.box[2], 0.5 * W.box[2]);
            }
            // egrad_water.f90:268 `r < cut`, exact: sqrt is monotone and correctly rounded, so only squared
            // distances within 1e-12 of cut^2 need the FP64 square root of the reference's own test
            const double r2 = d0 * d0 + d1 * d1 + d2 * d2;
            in = !W.periodic || r2 < c2lo || (r2 < c2hi && sqrt(r2) < W.coul_cut);
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (in) q[cnt + __popc(m & ((1u << lane) - 1u))] = j;
        cnt += __popc(m);
        __syncwarp();
        if (cnt >= 32) {
            wat_pair(W, x, gi, q[lane], xi0, xi1, xi2, qi, oi, f0, f1, f2, e);
            __syncwarp();
            cnt -= 32;
            if (lane < cnt) q[lane] = q[32 + lane];
            __syncwarp();
        }
    }
    if (lane < cnt) wat_pair(W, x, gi, q[lane], xi0, xi1, xi2, qi, oi, f0, f1, f2, e);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        f0 += __shfl_xor_sync(0xffffffffu, f0, o);
        f1 += __shfl_xor_sync(0xffffffffu, f1, o);
        f2 += __shfl_xor_sync(0xffffffffu, f2, o);
        e += __shfl_xor_sync(0xffffffffu, e, o);
    }
    if (lane == 0) {
        atomicAdd(&gi[3 * i], f0);
        atomicAdd(&gi[3 * i + 1], f1);
        atomicAdd(&gi[3 * i + 2], f2);
        atomicAdd(&V[img], e);
    }
}

cudaError_t water_egrad(const WaterDev* D, const double* d_xyz, int nimg, double* d_V, double* d_g, cudaStream_t s,
                        long long* launches)
{
    if (nimg <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_V, 0, (size_t)nimg * sizeof(double), s);
    if (e != cudaSuccess) return e;
    wat_intra_kernel<<<dim3((D->n / 3 + 127) / 128, nimg), 128, 0, s>>>(*D, d_xyz, nimg, d_V, d_g);
    if (D->n > 3) wat_inter_kernel<<<dim3((D->n - 1 + 3) / 4, nimg), 128, 0, s>>>(*D, d_xyz, nimg, d_V, d_g);
    if (launches) *launches += (D->n > 3) ? 2 : 1;
    return cudaGetLastError();
}

int water_upload(const crcl_water_params* P, WaterDev** out, const char** err)
{
    *out = nullptr;
    if (P->n <= 0 || P->n % 3 != 0 || !P->q || !P->is_O) {
        *err = "crcl_set_water: n must be a positive multiple of three (O,H,H per molecule) and q, is_O given";
        return CRCL_EINVAL;
    }
    for (int i = 0; i < P->n; i++)
        if ((P->is_O[i] != 0) != (i % 3 == 0)) {
            *err = "crcl_set_water: atoms must be ordered O,H,H per molecule (water_init.f90:58-69)";
            return CRCL_EINVAL;
        }
    WaterDev* D = new WaterDev();
    D->n = P->n;
    D->periodic = P->periodic;
    D->zahn = P->zahn;
    for (int d = 0; d < 3; d++) D->box[d] = P->box[d];
    D->coul_cut = P->coul_cut;
    D->zahn_a = P->zahn_a;
    D->zahn_par = P->zahn_par;
    for (int k = 0; k < 11; k++) D->pars[k] = P->pars[k];
    D->q = nullptr;
    D->is_O = nullptr;
    bool ok = cudaMalloc(&D->q, P->n * sizeof(double)) == cudaSuccess && cudaMalloc(&D->is_O, P->n * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMemcpy(D->q, P->q, P->n * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(D->is_O, P->is_O, P->n * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        water_free(D);
        *err = "crcl_set_water: device allocation failed";
        return CRCL_ENOMEM;
    }
    *out = D;
    return CRCL_OK;
}

void water_free(WaterDev* D)
{
    if (!D) return;
    if (D->q) cudaFree(D->q);
    if (D->is_O) cudaFree(D->is_O);
    delete D;
}

}  // namespace crcl
