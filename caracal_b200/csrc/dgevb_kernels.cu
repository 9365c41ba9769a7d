// dgevb_kernels.cu -- DG-EVB coupling and two-state mixing on the device.
//
// Replaces the dg_evb branch of gradient.f90:365-537 (analytic path, num_grad = .false.):
//   E  = (E1+E2)/2 - sqrt(((E1-E2)/2)^2 + V12sq)
//   g  = ( g1 + g2 - (dE (g1-g2) + 2 grad V12sq) / sqrt(dE^2 + 4 V12sq) ) / 2
// with V12sq the distributed-Gaussian expansion in internal coordinates (sum_v12.f90, sum_dv12.f90,
// modes 1-3), internal coordinates from xyz_2int.f90 (dist/ang/dihed/oop.f90) and the Cartesian
// gradient through the Wilson B matrix (int2grad.f90).  The reference ALWAYS builds B numerically
// (init_int.f90:142 sets num_wilson=.true.): central differences with shift 1e-3 bohr of the
// internal-coordinate functions (calc_wilson.f90:114-178) -- reproduced, including the
// x-h, (x-h)+2h evaluation points.  E1, g1 / E2, g2 come from the QMDFF kernels run on the two
// table sets (the second with the *_two semantics: never periodic, list terms only, Coulomb
// without cut-off; ff_eg_two.f90, ff_nonb_two.f90, ff_hb_two.f90).
//
// One CTA (64 threads) per image: thread l owns internal coordinate l for sum_dv12 (the
// O(points * nat6^3) part of mode 3), B entries are spread over all threads.
#include <cmath>
#include <vector>
#include "dgevb.cuh"

namespace crcl {

__device__ __forceinline__ double d3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void c3(const double a[3], const double b[3], double c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
// internal coordinate of type `ty` on the (up to four) atom positions p[0..3]
__device__ double ic_eval(int ty, const double p[4][3])
{
    if (ty == 1) {  // dist.f90
        const double dx = p[1][0] - p[0][0], dy = p[1][1] - p[0][1], dz = p[1][2] - p[0][2];
        return sqrt(dx * dx + (dy * dy + dz * dz));
    } else if (ty == 2) {  // ang.f90: angle at atom 2
        double a[3], b[3];
        for (int c = 0; c < 3; c++) {
            a[c] = p[0][c] - p[1][c];
            b[c] = p[2][c] - p[1][c];
        }
        return acos(d3(a, b) / (sqrt(d3(a, a)) * sqrt(d3(b, b))));
    } else if (ty == 3) {  // dihed.f90
        double u[3], v[3], w[3], uxw[3], vxw[3];
        for (int c = 0; c < 3; c++) {
            u[c] = p[0][c] - p[1][c];
            v[c] = p[3][c] - p[2][c];
            w[c] = p[2][c] - p[1][c];
        }
        const double ul = sqrt(d3(u, u)), vl = sqrt(d3(v, v)), wl = sqrt(d3(w, w));
        for (int c = 0; c < 3; c++) {
            u[c] = u[c] / ul;
            v[c] = v[c] / vl;
            w[c] = w[c] / wl;
        }
        c3(u, w, uxw);
        c3(v, w, vxw);
        const double uw = d3(u, w), vw = d3(v, w);
        double cv = d3(uxw, vxw) / (sqrt(1.0 - uw * uw) * sqrt(1.0 - vw * vw));
        cv = (cv >= 1.0) ? 1.0 : ((cv <= -1.0) ? -1.0 : cv);
        return acos(cv);
    } else {  // oop.f90
        double v41[3], v42[3], v43[3], c12[3], c23[3], c31[3], nv[3];
        for (int c = 0; c < 3; c++) {
            v41[c] = p[3][c] - p[0][c];
            v42[c] = p[3][c] - p[1][c];
            v43[c] = p[3][c] - p[2][c];
        }
        const double l1 = sqrt(d3(v41, v41)), l2 = sqrt(d3(v42, v42)), l3 = sqrt(d3(v43, v43));
        for (int c = 0; c < 3; c++) {
            v41[c] = v41[c] / l1;
            v42[c] = v42[c] / l2;
            v43[c] = v43[c] / l3;
        }
        c3(v41, v42, c12);
        c3(v42, v43, c23);
        c3(v43, v41, c31);
        for (int c = 0; c < 3; c++) nv[c] = c12[c] + c23[c] + c31[c];
        return d3(v41, c23) / sqrt(d3(nv, nv));
    }
}
__device__ __forceinline__ void ic_load(const DgevbDev& P, const double* x, int i, int& ty, int& nact, int at[4],
                                        double p[4][3])
{
    const int* cd = P.coord_def + 5 * i;
    ty = cd[0];
    nact = (ty == 1) ? 2 : (ty == 2 ? 3 : 4);
    for (int a = 0; a < 4; a++) {
        at[a] = (a < nact) ? cd[1 + a] : cd[1];
        for (int c = 0; c < 3; c++) p[a][c] = x[3 * at[a] + c];
    }
}

__global__ void __launch_bounds__(64) dgevb_mix_kernel(const DgevbDev P, int natoms, const double* __restrict__ xyz,
                                                      const double* __restrict__ V1, const double* __restrict__ G1,
                                                      const double* __restrict__ V2, const double* __restrict__ G2,
                                                      double* __restrict__ V, double* __restrict__ G)
{
    extern __shared__ double sm[];
    const int nat6 = P.nat6, n3 = 3 * natoms, img = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    double* internal = sm;           // [nat6]
    double* gq = internal + nat6;    // [nat6]
    double* qq = gq + nat6;          // [nat6]
    double* gv = qq + nat6;          // [3n]
    double* red = gv + n3;           // [2]: d_p, V12
    const double* x = xyz + (size_t)img * n3;
    for (int i = tid; i < nat6; i += nth) {
        int ty, nact, at[4];
        double p[4][3];
        ic_load(P, x, i, ty, nact, at, p);
        internal[i] = ic_eval(ty, p);
        gq[i] = 0.0;
    }
    for (int c = tid; c < n3; c += nth) gv[c] = 0.0;
    if (tid == 0) red[1] = 0.0;
    __syncthreads();
    const int mode = P.mode;
    const int block = 1 + nat6 + nat6 * (nat6 + 1) / 2, first = 1 + nat6;
    const double* b = P.b_vec - 1;   // 1-based as in the reference
    for (int j = 1; j <= P.npoints; j++) {
        const double* pt = P.point_int + (size_t)(j - 1) * nat6;
        const double al = P.alph[j - 1];
        for (int k = tid; k < nat6; k += nth) qq[k] = internal[k] - pt[k];
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int k = 0; k < nat6; k++) s += qq[k] * qq[k];
            red[0] = s;
        }
        __syncthreads();
        const double d_p = red[0];
        const double expo = exp(-0.5 * al * d_p);
        if (!(expo < P.g_thres)) {
            // ---- sum_v12: thread 0 (O(nat6^2) at most) ----
            if (tid == 0) {
                double v = 0.0;
                if (mode == 1) {
                    v = b[j] * (1 + 0.5 * al * d_p) * expo;
                } else if (mode == 2) {
                    v = b[(j - 1) * nat6 + j] * (1 + 0.5 * al * d_p) * expo;
                    for (int k = 1; k <= nat6; k++) v += b[j + (j - 1) * nat6 + k] * qq[k - 1] * expo;
                } else {
                    v = b[(j - 1) * (block - 1) + j] * (1 + 0.5 * al * d_p) * expo;
                    for (int k = 1; k <= nat6; k++) v += b[k + (j - 1) * block + 1] * qq[k - 1] * expo;
                    int inc = 0;
                    for (int k = 1; k <= nat6; k++)
                        for (int l = k; l <= nat6; l++) {
                            inc++;
                            const double bb = b[inc + (j - 1) * block + first];
                            v += (k == l) ? bb * 0.5 * qq[k - 1] * qq[l - 1] * expo : bb * qq[k - 1] * qq[l - 1] * expo;
                        }
                }
                red[1] += v;
            }
            // ---- sum_dv12: thread l owns g_V12(l) ----
            for (int l = tid + 1; l <= nat6; l += nth) {
                double a = 0.0;
                const double ql = qq[l - 1];
                if (mode == 1) {
                    a -= 0.5 * al * al * b[j] * d_p * ql * expo;
                } else if (mode == 2) {
                    a -= 0.5 * al * al * b[(j - 1) * nat6 + j] * d_p * ql * expo;
                    for (int k = 1; k <= nat6; k++) {
                        const double bb = b[j + (j - 1) * nat6 + k];
                        a -= (k == l) ? bb * (al * qq[k - 1] * ql - 1.0) * expo : bb * al * ql * qq[k - 1] * expo;
                    }
                } else {
                    a -= 0.5 * al * al * b[(j - 1) * (block - 1) + j] * d_p * ql * expo;
                    for (int k = 1; k <= nat6; k++) {
                        const double bb = b[k + (j - 1) * block + 1];
                        a -= (k == l) ? bb * (al * qq[k - 1] * ql - 1.0) * expo : bb * al * ql * qq[k - 1] * expo;
                    }
                    int inc = 0;
                    for (int k = 1; k <= nat6; k++) {
                        const double qk = qq[k - 1];
                        for (int m = k; m <= nat6; m++) {
                            inc++;
                            const double bb = b[inc + (j - 1) * block + first], qm = qq[m - 1];
                            if (m == k) {
                                a -= (l == k) ? 0.5 * bb * qk * (al * qk * qk - 2.0) * expo
                                              : 0.5 * bb * al * qk * qk * ql * expo;
                            } else {
                                if (l == k)
                                    a -= bb * qm * (al * qk * qk - 1.0) * expo;
                                else if (l == m)
                                    a -= bb * qk * (al * qm * qm - 1.0) * expo;
                                else
                                    a -= bb * al * qm * qk * ql * expo;
                            }
                        }
                    }
                }
                gq[l - 1] += a;
            }
        }
        __syncthreads();
    }
    // ---- numeric Wilson B (calc_wilson.f90:114-178) contracted with gq: gv = B^T gq ----
    for (int w = tid; w < nat6 * 12; w += nth) {
        const int i = w / 12, slot = w - 12 * i;
        int ty, nact, at[4];
        double p[4][3];
        ic_load(P, x, i, ty, nact, at, p);
        const int a = slot / 3, m = slot - 3 * a;
        if (a < nact) {
            const double shift = 0.001;
            p[a][m] = p[a][m] - shift;
            // an atom may appear once only in a coordinate definition, so perturbing slot a is
            // perturbing atom at[a] as the reference does
            const double lo = ic_eval(ty, p);
            p[a][m] = p[a][m] + 2 * shift;
            const double hi = ic_eval(ty, p);
            atomicAdd(&gv[3 * at[a] + m], (hi - lo) / (2 * shift) * gq[i]);
        }
    }
    __syncthreads();
    const double e1 = V1[img], e2 = V2[img];
    const double ediff = e1 - e2, V12 = red[1], off4 = 4.0 * V12;
    const bool unset = (ediff * ediff + off4 < 0.0);
    const double root2 = unset ? 1.0 : sqrt(ediff * ediff + off4);
    const double* g1 = G1 + (size_t)img * n3;
    const double* g2 = G2 + (size_t)img * n3;
    double* g = G + (size_t)img * n3;
    for (int c = tid; c < n3; c += nth) {
        const double deldiscr = ediff * (g1[c] - g2[c]) + 2.0 * gv[c];
        const double delsqrt = unset ? 0.0 : deldiscr / root2;
        g[c] = 0.5 * (g1[c] + g2[c] - delsqrt);
    }
    if (tid == 0) {
        const double root = (0.5 * ediff) * (0.5 * ediff) + V12;
        V[img] = (root <= 0) ? 0.5 * (e1 + e2) : 0.5 * (e1 + e2) - sqrt(root);
    }
}

cudaError_t dgevb_mix(const DgevbDev* P, int natoms, const double* xyz, int nimg, const double* V1, const double* G1,
                      const double* V2, const double* G2, double* V, double* G, cudaStream_t s)
{
    if (nimg <= 0) return cudaSuccess;
    const size_t smem = sizeof(double) * (3 * (size_t)P->nat6 + 3 * (size_t)natoms + 2);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(dgevb_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dgevb_mix_kernel<<<nimg, 64, smem, s>>>(*P, natoms, xyz, V1, G1, V2, G2, V, G);
    return cudaGetLastError();
}

int dgevb_upload(const crcl_dgevb_params* E, int natoms, DgevbDev** out, const char** err)
{
    *out = nullptr;
    if (E->mode < 1 || E->mode > 3 || E->npoints < 1 || E->nat6 < 1) {
        *err = "DG-EVB: mode must be 1..3, npoints and nat6 positive";
        return CRCL_EINVAL;
    }
    const int nat6 = E->nat6;
    std::vector<int> cd(5 * (size_t)nat6);
    for (int i = 0; i < nat6; i++) {
        const int ty = E->coord_def[5 * i];
        if (ty < 1 || ty > 4) {
            *err = "DG-EVB: coordinate type must be 1 (dist), 2 (angle), 3 (dihedral) or 4 (oop)";
            return CRCL_EINVAL;
        }
        const int nact = (ty == 1) ? 2 : (ty == 2 ? 3 : 4);
        cd[5 * i] = ty;
        for (int a = 0; a < 4; a++) {
            const int v = E->coord_def[5 * i + 1 + a];
            if (a < nact && (v < 1 || v > natoms)) {
                *err = "DG-EVB: coord_def atom index out of range";
                return CRCL_EINVAL;
            }
            cd[5 * i + 1 + a] = (a < nact) ? v - 1 : 0;
        }
    }
    const size_t mat = (E->mode == 1) ? (size_t)E->npoints
                                      : (E->mode == 2 ? (size_t)E->npoints * (1 + nat6)
                                                      : (size_t)E->npoints * (1 + nat6 + (size_t)nat6 * (nat6 + 1) / 2));
    DgevbDev* P = new DgevbDev();
    P->mode = E->mode;
    P->npoints = E->npoints;
    P->nat6 = nat6;
    P->g_thres = E->g_thres;
    bool ok = cudaMalloc(&P->coord_def, cd.size() * sizeof(int)) == cudaSuccess &&
              cudaMalloc(&P->point_int, (size_t)E->npoints * nat6 * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&P->alph, E->npoints * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&P->b_vec, mat * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMemcpy(P->coord_def, cd.data(), cd.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(P->point_int, E->point_int, (size_t)E->npoints * nat6 * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(P->alph, E->alph, E->npoints * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(P->b_vec, E->b_vec, mat * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        dgevb_free(P);
        *err = "DG-EVB: device allocation / upload failed";
        return CRCL_ENOMEM;
    }
    *out = P;
    return CRCL_OK;
}

void dgevb_free(DgevbDev* P)
{
    if (!P) return;
    cudaFree(P->coord_def);
    cudaFree(P->point_int);
    cudaFree(P->alph);
    cudaFree(P->b_vec);
    delete P;
}

}  // namespace crcl
