// K2: fixed-radius k-NN over the uniform grid + plane-to-plane regularised covariance.
// Replaces compute_covariance_matrix / compute_covariance_matrix_single_point
// (reference gicp.py:19-35, 5-17): tree.query(k, distance_upper_bound) per point,
// np.cov, np.linalg.eig, R diag(100,10) R^T.
//
// Work decomposition: one warp = 32 consecutive points of the Morton-sorted order (a compact
// blob of a few cells).  The warp stages the cell blocks around its queries with TMA bulk copies
// (stream.cuh) and every lane scans the staged candidates (broadcast LDS.128) keeping its own
// sorted top-k in registers.  Selection is exact: an fp32 distance is only a conservative
// filter; the ranked key is the float64 squared distance (dx*dx + dy*dy) + dz*dz of the oracle,
// ties broken by the lower point index.  Survivors of the filter are parked in a per-lane queue in
// shared memory and merged into the sorted top-k in batches, so that the long unrolled insertion
// network runs with many lanes active.  The search grows ring by ring (ring 0 = the queries' own
// cells) until every lane's k-th distance is covered by the searched box (or the radius is).
#pragma once
#include "common.cuh"
#include "stream.cuh"

namespace gicp {

constexpr int KNN_WARPS = 4;
constexpr int KNN_THREADS = KNN_WARPS * 32;
constexpr int KNN_STAGE_BYTES = 8192;  // per warp
constexpr int KNN_QUEUE = 24;          // parked survivors per lane
constexpr int KNN_WARP_SMEM = KNN_STAGE_BYTES + KNN_QUEUE * 32 * 12;
constexpr int KNN_GROUP_REACH = 4;

template <typename Real> struct KnnArgs {
    const CloudMeta* meta;
    const int* cell_start;
    const int* lut;
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
    unsigned char* wbase = smem_raw + 128 + warp * KNN_WARP_SMEM;
    WarpStage<Real> ws;
    ws.buf = reinterpret_cast<PRec<Real>*>(wbase);
    ws.bar = bars + warp;
    ws.phase = 0;
    ws.cap = KNN_STAGE_BYTES / (int)sizeof(PRec<Real>);
    double* qd = reinterpret_cast<double*>(wbase + KNN_STAGE_BYTES);             // [KNN_QUEUE][32]
    int* qi = reinterpret_cast<int*>(wbase + KNN_STAGE_BYTES + KNN_QUEUE * 32 * 8);  // [KNN_QUEUE][32]

    const CloudMeta m = a.meta[blockIdx.y];
    int begin = m.pt_begin, end = m.pt_end;
    if (a.slice_begin >= 0) { begin = max(begin, a.slice_begin); end = min(end, a.slice_end); }
    const int base = begin + (blockIdx.x * KNN_WARPS + warp) * 32;
    if (base >= end) return;

    if (lane == 0) { mbar_init(ws.bar, 1); mbar_fence_init(); }
    __syncwarp();

    const bool valid = base + lane < end;
    const PRec<Real> me = a.spts[valid ? base + lane : end - 1];
    const int my_idx = (int)me.idx;
    const int cx = min(max(cell_coord((double)me.x, m.origin[0], m.inv_h), 0), m.dims[0] - 1);
    const int cy = min(max(cell_coord((double)me.y, m.origin[1], m.inv_h), 0), m.dims[1] - 1);
    const int cz = (D == 3) ? min(max(cell_coord((double)me.z, m.origin[2], m.inv_h), 0), m.dims[2] - 1) : 0;

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
    int qn = 0;

    auto drain = [&]() {
        const int nmax = warp_max(qn);
        for (int i = 0; i < nmax; ++i) {
            if (i < qn) {
                const double e2 = qd[i * 32 + lane];
                const int ci = qi[i * 32 + lane];
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
                }
            }
        }
        qn = 0;
        thr32 = __double2float_ru(ad[KCAP - 1] * (1.0 + 1e-6));
    };

    const int rho_max = max(1, (int)ceil(a.radius / (m.h * (1.0 - 1e-9))));
    unsigned pending = 0xffffffffu;
    while (pending) {
        const unsigned grp = next_group(pending, cx, cy, cz, KNN_GROUP_REACH);
        pending &= ~grp;
        const bool mine = (grp >> lane) & 1u;
        const int mycell[3] = {cx, cy, cz};
        int qlo[3], qhi[3];
        group_union(grp, lane, mycell, mycell, m, qlo, qhi);
        int plo[3] = {0, 0, 0}, phi[3] = {-1, -1, -1};
        for (int rho = 0; rho <= rho_max; ++rho) {
            int lo[3], hi[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) { lo[c] = max(qlo[c] - rho, 0); hi[c] = min(qhi[c] + rho, m.dims[c] - 1); }
            stream_cells<Real>(m, a.cell_start, a.lut, a.spts, lo, hi, plo, phi, rho > 0, ws, lane,
                               [&](const PRec<Real>& c) {
                bool pass = mine;
                if (sizeof(Real) == 4) {
                    const float dx = (float)c.x - (float)mx, dy = (float)c.y - (float)my, dz = (float)c.z - (float)mz;
                    pass = pass && (fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= thr32);
                }
                if (pass) {
                    const double e2 = exact_d2((double)c.x - (double)mx, (double)c.y - (double)my,
                                               (double)c.z - (double)mz);
                    const int ci = (int)c.idx;
                    if (key_less(e2, ci, ad[KCAP - 1], ai[KCAP - 1])) {
                        qd[qn * 32 + lane] = e2;
                        qi[qn * 32 + lane] = ci;
                        ++qn;
                    }
                }
                if (__any_sync(0xffffffffu, qn == KNN_QUEUE)) drain();
            });
            drain();
#pragma unroll
            for (int c = 0; c < 3; ++c) { plo[c] = lo[c]; phi[c] = hi[c]; }
            if (rho == 0) continue;
            // done when the searched box covers every lane's k-th distance, or the radius
            const double cover = rho * m.h * (1.0 - 1e-9);
            const bool mine_done = !mine || ((ai[KCAP - 1] != INT_MAX) && (ad[KCAP - 1] <= cover * cover));
            if (__all_sync(0xffffffffu, mine_done) || cover >= a.radius) break;
        }
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
