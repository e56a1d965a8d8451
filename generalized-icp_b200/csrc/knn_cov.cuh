// K2: fixed-radius k-NN over the uniform grid + plane-to-plane regularised covariance.
// Replaces compute_covariance_matrix / compute_covariance_matrix_single_point
// (reference gicp.py:19-35, 5-17): tree.query(k, distance_upper_bound) per point,
// np.cov, np.linalg.eig, R diag(100,10) R^T.
//
// Work decomposition: one warp = 32 consecutive points of the cell-sorted order (so the
// 32 queries sit in a handful of neighbouring cells).  The warp stages the candidate runs
// of the cells around its queries' bounding box into shared memory with TMA bulk copies
// (cp.async.bulk, one per grid row, completion on an mbarrier) and every lane scans the
// staged candidates (broadcast LDS.128) keeping its own sorted top-k in registers.
// Selection is exact: an fp32 distance is only a conservative filter; the key that is
// ranked is the float64 squared distance (dx*dx + dy*dy) + dz*dz of the oracle, ties broken
// by the lower point index.  The search grows ring by ring until every lane's k-th
// distance is covered by the searched box (or the radius is).
#pragma once
#include "common.cuh"

namespace gicp {

constexpr int KNN_WARPS = 4;
constexpr int KNN_THREADS = KNN_WARPS * 32;
constexpr int KNN_STAGE_BYTES = 8192;  // per warp

template <typename Real> struct KnnArgs {
    const CloudMeta* meta;
    const int* cell_start;
    const PRec<Real>* spts;
    const Real* raw;     // caller's (n_total, D) array: neighbour coordinates for the covariance
    Real* cov_sorted;    // [n_total][NS], cell-sorted order
    int* knn_idx;        // optional [n_total][k], input order, cloud-local, -1 missing
    double* knn_dist;    // optional [n_total][k]
    int k;
    double radius;
    double lam_t, lam_n;
    int slice_begin, slice_end;  // sorted-position slice handled by this rank (<0: whole clouds)
};

__device__ __forceinline__ bool key_less(double d, int i, double dd, int ii) {
    return d < dd || (d == dd && i < ii);
}

template <int D, typename Real, int KCAP>
__global__ void __launch_bounds__(KNN_THREADS) knn_cov_kernel(const KnnArgs<Real> a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    PRec<Real>* stage = reinterpret_cast<PRec<Real>*>(smem_raw + 128 + warp * KNN_STAGE_BYTES);
    uint64_t* bar = bars + warp;
    constexpr int CAP = KNN_STAGE_BYTES / (int)sizeof(PRec<Real>);

    const CloudMeta m = a.meta[blockIdx.y];
    int begin = m.pt_begin, end = m.pt_end;
    if (a.slice_begin >= 0) { begin = max(begin, a.slice_begin); end = min(end, a.slice_end); }
    const int base = begin + (blockIdx.x * KNN_WARPS + warp) * 32;
    if (base >= end) return;

    if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncwarp();
    uint32_t phase = 0;

    const bool valid = base + lane < end;
    const PRec<Real> me = a.spts[valid ? base + lane : end - 1];
    const int my_idx = (int)me.idx;
    const int nx = m.dims[0], ny = m.dims[1], nz = m.dims[2];
    int cx = min(max(cell_coord((double)me.x, m.origin[0], m.inv_h), 0), nx - 1);
    int cy = min(max(cell_coord((double)me.y, m.origin[1], m.inv_h), 0), ny - 1);
    int cz = (D == 3) ? min(max(cell_coord((double)me.z, m.origin[2], m.inv_h), 0), nz - 1) : 0;
    const int xa = warp_min(cx), xb = warp_max(cx);
    const int ya = warp_min(cy), yb = warp_max(cy);
    const int za = warp_min(cz), zb = warp_max(cz);

    // sorted top-k: slots [KCAP-k, KCAP) are live, the ones before hold -1 sentinels
    double ad[KCAP];
    int ai[KCAP];
    const double r2cap = a.radius * a.radius * (1.0 + 1e-12);
#pragma unroll
    for (int s = 0; s < KCAP; ++s) {
        const bool live = s >= KCAP - a.k;
        ad[s] = live ? r2cap : -1.0;
        ai[s] = live ? INT_MAX : -1;
    }
    float thr32 = __double2float_ru(r2cap * (1.0 + 1e-6));
    const Real mx = me.x, my = me.y, mz = me.z;

    const int rho_max = max(1, (int)ceil(a.radius / (m.h * (1.0 - 1e-9))));
    for (int rho = 1; rho <= rho_max; ++rho) {
        const int X0 = max(xa - rho, 0), X1 = min(xb + rho, nx - 1);
        const int Y0 = max(ya - rho, 0), Y1 = min(yb + rho, ny - 1);
        const int Z0 = max(za - rho, 0), Z1 = min(zb + rho, nz - 1);
        const int PX0 = max(xa - rho + 1, 0), PX1 = min(xb + rho - 1, nx - 1);
        const int PY0 = max(ya - rho + 1, 0), PY1 = min(yb + rho - 1, ny - 1);
        const int PZ0 = max(za - rho + 1, 0), PZ1 = min(zb + rho - 1, nz - 1);
        const int nyb = Y1 - Y0 + 1, nzb = Z1 - Z0 + 1;
        const int n_slots = 2 * nyb * nzb;
        const int n_sweeps = (rho == 1) ? 2 : 1;
        for (int sweep = 0; sweep < n_sweeps; ++sweep) {
            for (int g0 = 0; g0 < n_slots; g0 += 32) {
                const int e = g0 + lane;
                int start = 0, len = 0;
                if (e < n_slots) {
                    const int row = e >> 1, side = e & 1;
                    const int y = Y0 + row % nyb, z = Z0 + row / nyb;
                    int xlo = 0, xhi = -1;
                    if (rho == 1) {
                        const bool core = (y >= ya && y <= yb && z >= za && z <= zb);
                        if (side == 0 && core == (sweep == 0)) { xlo = X0; xhi = X1; }
                    } else {
                        const bool in_prev = (y >= PY0 && y <= PY1 && z >= PZ0 && z <= PZ1);
                        if (!in_prev) {
                            if (side == 0) { xlo = X0; xhi = X1; }
                        } else if (side == 0) { xlo = X0; xhi = PX0 - 1; }
                        else { xlo = PX1 + 1; xhi = X1; }
                    }
                    if (xhi >= xlo) {
                        const int rowbase = m.cell_base + (z * ny + y) * nx;
                        start = __ldg(a.cell_start + rowbase + xlo);
                        len = __ldg(a.cell_start + rowbase + xhi + 1) - start;
                    }
                }
                const int incl = warp_incl_scan(len, lane);
                const int excl = incl - len;
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                for (int w0 = 0; w0 < total; w0 += CAP) {
                    const int n_win = min(CAP, total - w0);
                    if (lane == 0) mbar_expect_tx(bar, (uint32_t)(n_win * sizeof(PRec<Real>)));
                    __syncwarp();
                    const int lo = max(excl, w0), hi = min(excl + len, w0 + CAP);
                    if (hi > lo)
                        tma_load_1d(stage + (lo - w0), a.spts + start + (lo - excl),
                                    (uint32_t)((hi - lo) * sizeof(PRec<Real>)), bar);
                    mbar_wait(bar, phase);
                    phase ^= 1u;
                    for (int j = 0; j < n_win; ++j) {
                        const PRec<Real> c = stage[j];
                        bool pass;
                        if (sizeof(Real) == 4) {
                            const float dx = (float)c.x - (float)mx, dy = (float)c.y - (float)my,
                                        dz = (float)c.z - (float)mz;
                            pass = fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= thr32;
                        } else {
                            pass = true;
                        }
                        if (pass) {
                            const double e2 = exact_d2((double)c.x - (double)mx, (double)c.y - (double)my,
                                                       (double)c.z - (double)mz);
                            const int ci = (int)c.idx;
                            bool lt_s = key_less(e2, ci, ad[KCAP - 1], ai[KCAP - 1]);
                            if (lt_s) {
#pragma unroll
                                for (int s = KCAP - 1; s > 0; --s) {
                                    const bool lt_prev = key_less(e2, ci, ad[s - 1], ai[s - 1]);
                                    if (lt_prev) { ad[s] = ad[s - 1]; ai[s] = ai[s - 1]; }
                                    else if (lt_s) { ad[s] = e2; ai[s] = ci; }
                                    lt_s = lt_prev;
                                }
                                if (lt_s) { ad[0] = e2; ai[0] = ci; }
                                thr32 = __double2float_ru(ad[KCAP - 1] * (1.0 + 1e-6));
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
        // done when the searched box covers every lane's k-th distance, or the radius
        const double cover = rho * m.h * (1.0 - 1e-9);
        const bool mine_done = (ai[KCAP - 1] != INT_MAX) && (ad[KCAP - 1] <= cover * cover);
        if (__all_sync(0xffffffffu, mine_done) || cover >= a.radius) break;
    }

    if (!valid) return;
    // ---- covariance of the surviving neighbours (gicp.py:25-34, 5-17) ----
    const size_t cloud_row0 = (size_t)m.pt_begin;
    int cnt = 0;
    double mean[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int s = 0; s < KCAP; ++s) {
        const bool ok = (s >= KCAP - a.k) && ai[s] != INT_MAX && sqrt(ad[s]) < a.radius;
        if (!ok && s >= KCAP - a.k) ai[s] = INT_MAX;
        if (ok) {
            ++cnt;
            const Real* q = a.raw + (cloud_row0 + ai[s]) * D;
#pragma unroll
            for (int c = 0; c < D; ++c) mean[c] += (double)q[c];
        }
    }
    if (a.knn_idx) {
        int o = 0;
        int* out = a.knn_idx + (cloud_row0 + my_idx) * (size_t)a.k;
        double* outd = a.knn_dist ? a.knn_dist + (cloud_row0 + my_idx) * (size_t)a.k : nullptr;
#pragma unroll
        for (int s = 0; s < KCAP; ++s) {
            if (s >= KCAP - a.k && ai[s] != INT_MAX) {
                out[o] = ai[s];
                if (outd) outd[o] = sqrt(ad[s]);
                ++o;
            }
        }
        for (; o < a.k; ++o) {
            out[o] = -1;
            if (outd) outd[o] = INFINITY;
        }
    }
    constexpr int NS = Dim<D>::NS;
    double C[NS];
    bool ident = cnt <= 1;
    if (!ident) {
        const double inv = 1.0 / cnt;
#pragma unroll
        for (int c = 0; c < D; ++c) mean[c] *= inv;
        double S[6] = {0, 0, 0, 0, 0, 0};  // 00 01 02 11 12 22
#pragma unroll
        for (int s = 0; s < KCAP; ++s) {
            if (s >= KCAP - a.k && ai[s] != INT_MAX) {
                const Real* q = a.raw + (cloud_row0 + ai[s]) * D;
                const double d0 = (double)q[0] - mean[0], d1 = (double)q[1] - mean[1];
                const double d2 = (D == 3) ? (double)q[D - 1] - mean[2] : 0.0;
                S[0] += d0 * d0; S[1] += d0 * d1; S[2] += d0 * d2;
                S[3] += d1 * d1; S[4] += d1 * d2; S[5] += d2 * d2;
            }
        }
        const double f = 1.0 / (cnt - 1);  // ddof = 1 (np.cov default, gicp.py:12)
#pragma unroll
        for (int i = 0; i < 6; ++i) S[i] *= f;
        bool finite = true;
#pragma unroll
        for (int i = 0; i < 6; ++i) finite = finite && isfinite(S[i]);
        if (!finite) {
            ident = true;  // gicp.py:31-32
        } else if constexpr (D == 2) {
            // eigenvector of the largest eigenvalue (gicp.py:14-16): C = lam_n I + (lam_t-lam_n) v v^T
            const double phi = 0.5 * atan2(2.0 * S[1], S[0] - S[3]);
            double sn, cs;
            sincos(phi, &sn, &cs);
            const double dl = a.lam_t - a.lam_n;
            C[0] = a.lam_n + dl * cs * cs;
            C[1] = dl * cs * sn;
            C[2] = a.lam_n + dl * sn * sn;
        } else {
            // normal = eigenvector of the smallest eigenvalue: C = lam_t I - (lam_t-lam_n) n n^T
            double n[3];
            smallest_eigvec3(S[0], S[1], S[2], S[3], S[4], S[5], n);
            const double dl = a.lam_t - a.lam_n;
            C[0] = a.lam_t - dl * n[0] * n[0];
            C[1] = -dl * n[0] * n[1];
            C[2] = -dl * n[0] * n[2];
            C[3] = a.lam_t - dl * n[1] * n[1];
            C[4] = -dl * n[1] * n[2];
            C[5] = a.lam_t - dl * n[2] * n[2];
        }
    }
    if (ident) {
        if constexpr (D == 2) { C[0] = 1.0; C[1] = 0.0; C[2] = 1.0; }
        else { C[0] = 1.0; C[1] = 0.0; C[2] = 0.0; C[3] = 1.0; C[4] = 0.0; C[5] = 1.0; }
    }
    Real* out = a.cov_sorted + (size_t)(base + lane) * NS;
#pragma unroll
    for (int i = 0; i < NS; ++i) out[i] = (Real)C[i];
}

}  // namespace gicp
