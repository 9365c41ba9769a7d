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
// One warp per image (see dgevb_mix_kernel).
#include <cmath>
#include <vector>
#include "dgevb.cuh"
#include "crcl_common.cuh"

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
        return CRCL_SQRT(dx * dx + (dy * dy + dz * dz));
    } else if (ty == 2) {  // ang.f90: angle at atom 2
        double a[3], b[3];
        for (int c = 0; c < 3; c++) {
            a[c] = p[0][c] - p[1][c];
            b[c] = p[2][c] - p[1][c];
        }
        return CRCL_ACOS(d3(a, b) * CRCL_RSQRT(d3(a, a)) * CRCL_RSQRT(d3(b, b)));
    } else if (ty == 3) {  // dihed.f90
        double u[3], v[3], w[3], uxw[3], vxw[3];
        for (int c = 0; c < 3; c++) {
            u[c] = p[0][c] - p[1][c];
            v[c] = p[3][c] - p[2][c];
            w[c] = p[2][c] - p[1][c];
        }
        const double iul = CRCL_RSQRT(d3(u, u)), ivl = CRCL_RSQRT(d3(v, v)), iwl = CRCL_RSQRT(d3(w, w));
        for (int c = 0; c < 3; c++) {
            u[c] = u[c] * iul;
            v[c] = v[c] * ivl;
            w[c] = w[c] * iwl;
        }
        c3(u, w, uxw);
        c3(v, w, vxw);
        const double uw = d3(u, w), vw = d3(v, w);
        double cv = d3(uxw, vxw) * CRCL_RSQRT(1.0 - uw * uw) * CRCL_RSQRT(1.0 - vw * vw);
        cv = (cv >= 1.0) ? 1.0 : ((cv <= -1.0) ? -1.0 : cv);
        return CRCL_ACOS(cv);
    } else {  // oop.f90
        double v41[3], v42[3], v43[3], c12[3], c23[3], c31[3], nv[3];
        for (int c = 0; c < 3; c++) {
            v41[c] = p[3][c] - p[0][c];
            v42[c] = p[3][c] - p[1][c];
            v43[c] = p[3][c] - p[2][c];
        }
        const double i1 = CRCL_RSQRT(d3(v41, v41)), i2 = CRCL_RSQRT(d3(v42, v42)), i3 = CRCL_RSQRT(d3(v43, v43));
        for (int c = 0; c < 3; c++) {
            v41[c] = v41[c] * i1;
            v42[c] = v42[c] * i2;
            v43[c] = v43[c] * i3;
        }
        c3(v41, v42, c12);
        c3(v42, v43, c23);
        c3(v43, v41, c31);
        for (int c = 0; c < 3; c++) nv[c] = c12[c] + c23[c] + c31[c];
        return d3(v41, c23) * CRCL_RSQRT(d3(nv, nv));
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

// packed index (1-based, as sum_dv12.f90 counts `inc`) of coefficient (k,m), k <= m, of the upper triangle
__device__ __forceinline__ int tri(int k, int m, int n) { return (k - 1) * (2 * n - k + 2) / 2 + (m - k + 1); }

// One WARP per image.  Per Gaussian the mode-3 gradient of sum_dv12.f90:124-176 is evaluated in its
// closed form -- with Q = 1/2 q^T B q (B the symmetric matrix of the second-order coefficients),
//   g_l = -expo [ 1/2 a^2 b0 d q_l + a q_l (b1.q) - b1_l + a q_l Q - (B q)_l ],
// O(nat6^2) per point instead of the reference's O(nat6^3) loop nest over (l,k,m); the two differ by
// rounding only.  Lane l owns g_V12(l) and row l of B q; the numeric Wilson B entries are spread over
// all lanes.
__global__ void __launch_bounds__(128) dgevb_mix_kernel(const DgevbDev P, int natoms, int nimg,
                                                       const double* __restrict__ xyz, const double* __restrict__ V1,
                                                       const double* __restrict__ G1, const double* __restrict__ V2,
                                                       const double* __restrict__ G2, double* __restrict__ V,
                                                       double* __restrict__ G)
{
    extern __shared__ double sm[];
    const int nat6 = P.nat6, n3 = 3 * natoms, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int img = blockIdx.x * (blockDim.x >> 5) + wib;
    if (img >= nimg) return;                                  // whole warps leave together
    const int per = 4 * nat6;
    double* internal = sm + (size_t)wib * per;   // [nat6]
    double* gq = internal + nat6;                // [nat6]
    double* qq = gq + nat6;                      // [nat6]
    double* Bq = qq + nat6;                      // [nat6]
    const double* x = xyz + (size_t)img * n3;
    for (int i = lane; i < nat6; i += 32) {
        int ty, nact, at[4];
        double p[4][3];
        ic_load(P, x, i, ty, nact, at, p);
        internal[i] = ic_eval(ty, p);
        gq[i] = 0.0;
    }
    __syncwarp();
    const int mode = P.mode;
    const int block = 1 + nat6 + nat6 * (nat6 + 1) / 2, first = 1 + nat6;
    const double* b = P.b_vec - 1;   // 1-based as in the reference
    double V12 = 0.0;
    for (int j = 1; j <= P.npoints; j++) {
        const double* pt = P.point_int + (size_t)(j - 1) * nat6;
        const double al = P.alph[j - 1];
        for (int k = lane; k < nat6; k += 32) qq[k] = internal[k] - pt[k];
        __syncwarp();
        double d_p = 0.0;
        for (int k = 0; k < nat6; k++) d_p += qq[k] * qq[k];   // every lane: same order as dot_product
        const double expo = CRCL_EXP(-0.5 * al * d_p);
        if (!(expo < P.g_thres)) {
            // offsets of the coefficient blocks of point j (sum_v12.f90 / sum_dv12.f90 index arithmetic)
            const double b0 = (mode == 1) ? b[j] : (mode == 2 ? b[(j - 1) * nat6 + j] : b[(j - 1) * (block - 1) + j]);
            const double* b1 = (mode == 2) ? b + j + (j - 1) * nat6 : b + (j - 1) * block + 1;   // b1[k], k = 1..nat6
            const double* b2 = b + (j - 1) * block + first;                                      // b2[inc]
            double b1q = 0.0, Q = 0.0;
            if (mode >= 2)
                for (int k = 1; k <= nat6; k++) b1q += b1[k] * qq[k - 1];
            if (mode == 3) {
                for (int l = lane + 1; l <= nat6; l += 32) {
                    double acc = 0.0;
                    for (int m = 1; m <= nat6; m++) acc += b2[(l <= m) ? tri(l, m, nat6) : tri(m, l, nat6)] * qq[m - 1];
                    Bq[l - 1] = acc;
                }
                __syncwarp();
                for (int k = 0; k < nat6; k++) Q += qq[k] * Bq[k];
                Q *= 0.5;
            }
            V12 += (b0 * (1 + 0.5 * al * d_p) + b1q + Q) * expo;
            for (int l = lane + 1; l <= nat6; l += 32) {
                const double ql = qq[l - 1];
                double a = 0.5 * al * al * b0 * d_p * ql;
                if (mode >= 2) a += al * ql * b1q - b1[l];
                if (mode == 3) a += al * ql * Q - Bq[l - 1];
                gq[l - 1] -= a * expo;
            }
        }
        __syncwarp();
    }
    // ---- mixing (gradient.f90:497-537) and the numeric Wilson B (calc_wilson.f90:114-178) contracted with gq ----
    // g = 1/2 (g1 + g2 - [ediff (g1 - g2) + 2 B^T gq] / root2) is linear in the Wilson term: lane c first stores the part
    // without it, then every finite-difference slot adds its share -(dq_i/dx_c) gq_i / root2 straight into g with a
    // fire-and-forget FP64 reduction (RED.ADD at the L2).  The first version accumulated B^T gq in shared memory, where an
    // FP64 atomicAdd is a compare-and-swap loop: 45 % of a DG-EVB step went there (profiles/r2q_launches_c4.csv).
    const double e1 = V1[img], e2 = V2[img];
    const double ediff = e1 - e2, off4 = 4.0 * V12;
    const bool unset = (ediff * ediff + off4 < 0.0);
    const double root2 = unset ? 1.0 : sqrt(ediff * ediff + off4);
    const double iroot2 = unset ? 0.0 : 1.0 / root2;
    const double* g1 = G1 + (size_t)img * n3;
    const double* g2 = G2 + (size_t)img * n3;
    double* g = G + (size_t)img * n3;
    for (int c = lane; c < n3; c += 32) g[c] = 0.5 * (g1[c] + g2[c] - ediff * (g1[c] - g2[c]) * iroot2);
    __syncwarp();   // the stores above are ordered before the reductions of the other lanes below
    for (int w = lane; w < nat6 * 12; w += 32) {
        const int i = w / 12, slot = w - 12 * i;
        int ty, nact, at[4];
        double p[4][3];
        ic_load(P, x, i, ty, nact, at, p);
        const int a = slot / 3, m = slot - 3 * a;
        if (a < nact) {
            const double shift = 0.001;
            // an atom may appear once only in a coordinate definition, so perturbing slot a is perturbing atom at[a] as the
            // reference does: x - shift, then (x - shift) + 2 shift.  The slot is picked by selects, not by indexing p
            // with the run-time (a, m), which would put the twelve coordinates in local memory.
            double pl[4][3], ph[4][3];
#pragma unroll
            for (int aa = 0; aa < 4; aa++)
#pragma unroll
                for (int mm = 0; mm < 3; mm++) {
                    const bool hit = (aa == a) && (mm == m);
                    pl[aa][mm] = hit ? p[aa][mm] - shift : p[aa][mm];
                    ph[aa][mm] = hit ? pl[aa][mm] + 2 * shift : p[aa][mm];
                }
            const double lo = ic_eval(ty, pl);
            const double hi = ic_eval(ty, ph);
            atomicAdd(&g[3 * at[a] + m], -((hi - lo) / (2 * shift) * gq[i]) * iroot2);
        }
    }
    if (lane == 0) {
        const double root = (0.5 * ediff) * (0.5 * ediff) + V12;
        V[img] = (root <= 0) ? 0.5 * (e1 + e2) : 0.5 * (e1 + e2) - sqrt(root);
    }
}

cudaError_t dgevb_mix(const DgevbDev* P, int natoms, const double* xyz, int nimg, const double* V1, const double* G1,
                      const double* V2, const double* G2, double* V, double* G, cudaStream_t s)
{
    if (nimg <= 0) return cudaSuccess;
    const size_t per = sizeof(double) * 4 * (size_t)P->nat6;
    int wpb = 4;                                  // warps (images) per CTA
    while (wpb > 1 && per * wpb > 96 * 1024) wpb >>= 1;
    const size_t smem = per * wpb;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(dgevb_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dgevb_mix_kernel<<<(nimg + wpb - 1) / wpb, 32 * wpb, smem, s>>>(*P, natoms, nimg, xyz, V1, G1, V2, G2, V, G);
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
