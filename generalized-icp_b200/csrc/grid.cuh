// K1: uniform-grid build.  Replaces scipy's KDTree(points) (reference gicp.py:21,127).
//   bbox -> per-cloud grid geometry -> Morton cell key per point (+ histogram) -> radix sort by key
//   -> exclusive scan of the histogram = cell_start -> gather points into sorted records.
#pragma once
#include "common.cuh"

namespace gicp {

constexpr int BBOX_THREADS = 256;
constexpr int BBOX_ITEMS = 16;  // points per thread

// grid (chunks, n_clouds): per-chunk min/max of every coordinate -> part[cloud][chunk][2*3]
template <int D, typename Real>
__global__ void __launch_bounds__(BBOX_THREADS) bbox_partial_kernel(const Real* __restrict__ pts,
                                                                    const int* __restrict__ offsets,
                                                                    double* __restrict__ part, int chunks) {
    const int cloud = blockIdx.y;
    const int b = offsets[cloud], e = offsets[cloud + 1];
    const int c0 = b + blockIdx.x * (BBOX_THREADS * BBOX_ITEMS);
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (c0 < e) {
#pragma unroll 4
        for (int i = 0; i < BBOX_ITEMS; ++i) {
            const int g = c0 + i * BBOX_THREADS + threadIdx.x;
            if (g < e) {
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const double v = (double)pts[(size_t)g * D + c];
                    if (isfinite(v)) { lo[c] = fmin(lo[c], v); hi[c] = fmax(hi[c], v); }
                }
            }
        }
    }
    __shared__ double s_lo[BBOX_THREADS / 32][3], s_hi[BBOX_THREADS / 32][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        for (int c = 0; c < 3; ++c) { s_lo[w][c] = lo[c]; s_hi[w][c] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        double l = s_lo[0][c], h = s_hi[0][c];
        for (int i = 1; i < BBOX_THREADS / 32; ++i) { l = fmin(l, s_lo[i][c]); h = fmax(h, s_hi[i][c]); }
        double* o = part + ((size_t)cloud * chunks + blockIdx.x) * 6;
        o[c] = l;
        o[3 + c] = h;
    }
}

// Geometry of one cloud's grid from its bounding box (run by one thread): the cell edge is the requested one,
// enlarged until the Morton-padded table (2^(sum bits) entries) fits the budget and no axis needs more than
// GICP_MAX_AXIS_BITS bits.  lo/hi are sanitised in place (unused axes and empty clouds -> 0).
template <int D>
__device__ inline CloudMeta choose_grid(double lo[3], double hi[3], double h_target, long long budget, int cloud,
                                        int pt_begin, int pt_end) {
    for (int c = 0; c < 3; ++c) {
        if (c >= D || !(lo[c] <= hi[c])) { lo[c] = 0.0; hi[c] = 0.0; }
    }
    CloudMeta m;
    double h = h_target;
    for (int guard = 0; guard < 400; ++guard) {
        m.h = h;
        m.inv_h = 1.0 / h;
        int sum_bits = 0;
        bool ok = true;
        for (int c = 0; c < 3; ++c) {
            m.dims[c] = (c < D) ? cell_coord(hi[c], lo[c], m.inv_h) + 1 : 1;
            int bts = 0;
            while ((1 << bts) < m.dims[c] && bts < 31) ++bts;
            m.bits[c] = bts;
            sum_bits += bts;
            if (bts > GICP_MAX_AXIS_BITS) ok = false;
        }
        if (ok && sum_bits <= 30 && (1LL << sum_bits) <= budget) break;
        // smallest enlargement that saves one Morton bit: the axis whose cell count is closest above a power of
        // two gives way first (extent / 2^(bits - 1) is the edge at which that axis fits one bit less)
        double h_next = INFINITY;
        for (int c = 0; c < D; ++c) {
            if (m.bits[c] < 1) continue;
            const double hc = (hi[c] - lo[c]) / (double)(1 << (m.bits[c] - 1)) * (1.0 + 1e-9);
            if (hc > h) h_next = fmin(h_next, hc);
        }
        h = (h_next < h * 1.2599210498948732) ? h_next : h * 1.2599210498948732;
    }
    for (int c = 0; c < 3; ++c) m.origin[c] = lo[c];
    m.cell_base = (int)((long long)cloud * budget);
    m.pt_begin = pt_begin;
    m.pt_end = pt_end;
    m.lut_base = cloud * 3 * GICP_LUT_N;
    return m;
}

// entry e (axis = e / GICP_LUT_N, v = e % GICP_LUT_N) of a cloud's per-axis Morton spread tables
__device__ inline int morton_lut_entry(const CloudMeta& m, int e) {
    const int axis = e / GICP_LUT_N, v = e % GICP_LUT_N;
    int code = 0, pos = 0;
    for (int b = 0; b < GICP_MAX_AXIS_BITS; ++b) {
        for (int ax = 0; ax < 3; ++ax) {
            if (b < m.bits[ax]) {
                if (ax == axis) code |= ((v >> b) & 1) << pos;
                ++pos;
            }
        }
    }
    return code;
}

// one WARP per cloud: reduce the partial boxes (lanes stride over the chunks: a 16.7 M-point cloud has 4096 of them),
// then lane 0 chooses the cell edge and lays out the cell table.
// bbox_out[cloud][6] keeps (lo, hi) for later use (centring point of the reduced form).
template <int D>
__global__ void grid_meta_kernel(const double* __restrict__ part, int chunks, const int* __restrict__ offsets,
                                 int n_clouds, double h_target, long long budget, CloudMeta* __restrict__ meta,
                                 double* __restrict__ bbox_out) {
    const int cloud = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (cloud >= n_clouds) return;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    const int n = offsets[cloud + 1] - offsets[cloud];
    const int used = min(chunks, (n + BBOX_THREADS * BBOX_ITEMS - 1) / (BBOX_THREADS * BBOX_ITEMS));
    for (int k = lane; k < used; k += 32) {
        const double* p = part + ((size_t)cloud * chunks + k) * 6;
        for (int c = 0; c < D; ++c) { lo[c] = fmin(lo[c], p[c]); hi[c] = fmax(hi[c], p[3 + c]); }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    }
    if (lane != 0) return;
    const CloudMeta m = choose_grid<D>(lo, hi, h_target, budget, cloud, offsets[cloud], offsets[cloud + 1]);
    meta[cloud] = m;
    if (bbox_out) {
        for (int c = 0; c < 3; ++c) { bbox_out[cloud * 6 + c] = lo[c]; bbox_out[cloud * 6 + 3 + c] = hi[c]; }
    }
}

// grid (blocks, n_clouds): cell key of every point, histogram of cell populations.
template <int D, typename Real>
__global__ void __launch_bounds__(256) cell_key_kernel(const Real* __restrict__ pts, const CloudMeta* __restrict__ meta,
                                                       unsigned* __restrict__ keys, int* __restrict__ vals,
                                                       int* __restrict__ cell_count) {
    const CloudMeta m = meta[blockIdx.y];
    const int g = m.pt_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= m.pt_end) return;
    int c[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const int v = cell_coord((double)pts[(size_t)g * D + a], m.origin[a], m.inv_h);
        c[a] = min(max(v, 0), m.dims[a] - 1);
    }
    const int cell = m.cell_base + morton_code(c[0], c[1], c[2], m.bits[0], m.bits[1], m.bits[2]);
    keys[g] = (unsigned)cell;
    vals[g] = g;
    atomicAdd(&cell_count[cell], 1);
}

// grid (blocks, n_clouds): sorted position s -> record {coords, idx}; idx is the cloud-local index of the point in
// the caller's array (the tie-break key of every search, and it addresses the raw coordinates).
// inv_perm[global row] = s maps an input row back to its sorted position.
template <int D, typename Real>
__global__ void __launch_bounds__(256) gather_sorted_kernel(const Real* __restrict__ pts,
                                                            const CloudMeta* __restrict__ meta,
                                                            const int* __restrict__ sorted_vals,
                                                            PRec<Real>* __restrict__ spts, int* __restrict__ inv_perm) {
    const CloudMeta m = meta[blockIdx.y];
    const int s = m.pt_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m.pt_end) return;
    const int g = sorted_vals[s];
    PRec<Real> r;
    r.x = pts[(size_t)g * D + 0];
    r.y = pts[(size_t)g * D + 1];
    r.z = (D == 3) ? pts[(size_t)g * D + (D - 1)] : Real(0);
    r.idx = g - m.pt_begin;
    spts[s] = r;
    inv_perm[g] = s;
}

// grid (3 * GICP_LUT_N / 256, n_clouds): per-axis Morton spread tables, code(x,y,z) = lx[x] | ly[y] | lz[z]
__global__ void morton_lut_kernel(const CloudMeta* __restrict__ meta, int* __restrict__ lut) {
    const CloudMeta m = meta[blockIdx.y];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 3 * GICP_LUT_N) return;
    lut[m.lut_base + e] = morton_lut_entry(m, e);
}

// ================================================================================================
// Small clouds (<= SMALL_GRID_MAX points each): the whole grid build of one cloud in ONE block - bounding box,
// geometry, Morton tables, keys, a bitonic sort of (cell code, input index) in shared memory, the cell table by
// binary search in the sorted keys, and the gather.  Produces exactly what the multi-kernel build produces (same
// geometry code, same stable order), with one launch instead of ten: a registration of two 360-beam scans is
// bound by launch latency, not by work.
// ================================================================================================
constexpr int SMALL_GRID_MAX = 2048;
constexpr int SMALL_GRID_THREADS = 256;

template <int D, typename Real>
__global__ void __launch_bounds__(SMALL_GRID_THREADS) small_grid_kernel(
    const Real* __restrict__ pts, const int* __restrict__ offsets, double h_target, long long budget,
    CloudMeta* __restrict__ meta, double* __restrict__ bbox_out, int* __restrict__ lut, int* __restrict__ cell_start,
    PRec<Real>* __restrict__ spts, int* __restrict__ inv_perm, int is_last_cloud_total_cells, int inline_n) {
    __shared__ unsigned s_key[SMALL_GRID_MAX];   // (cell code << 11) | input index
    __shared__ double s_lo[SMALL_GRID_THREADS / 32][3], s_hi[SMALL_GRID_THREADS / 32][3];
    __shared__ CloudMeta s_meta;
    const int cloud = blockIdx.x;
    // inline_n >= 0: a single cloud of that many points (the offsets array has not been uploaded)
    const int b = inline_n >= 0 ? 0 : offsets[cloud], e = inline_n >= 0 ? inline_n : offsets[cloud + 1];
    const int n = e - b;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // ---- bounding box ----
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < n; i += SMALL_GRID_THREADS) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const double v = (double)pts[(size_t)(b + i) * D + c];
            if (isfinite(v)) { lo[c] = fmin(lo[c], v); hi[c] = fmax(hi[c], v); }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    }
    if (lane == 0) {
        for (int c = 0; c < 3; ++c) { s_lo[w][c] = lo[c]; s_hi[w][c] = hi[c]; }
    }
    __syncthreads();
    if (tid == 0) {
        for (int c = 0; c < 3; ++c) {
            for (int i = 1; i < SMALL_GRID_THREADS / 32; ++i) { lo[c] = fmin(lo[c], s_lo[i][c]); hi[c] = fmax(hi[c], s_hi[i][c]); }
        }
        s_meta = choose_grid<D>(lo, hi, h_target, budget, cloud, b, e);
        meta[cloud] = s_meta;
        if (bbox_out) {
            for (int c = 0; c < 3; ++c) { bbox_out[cloud * 6 + c] = lo[c]; bbox_out[cloud * 6 + 3 + c] = hi[c]; }
        }
    }
    __syncthreads();
    const CloudMeta m = s_meta;
    // ---- Morton tables ----
    // (only the entries below the grid's extent are ever used in an address; the rest is zeroed, not computed:
    // computing all 3 x 1024 entries was 40 % of this kernel's instructions)
    for (int i = tid; i < 3 * GICP_LUT_N; i += SMALL_GRID_THREADS) {
        const int axis = i / GICP_LUT_N, v = i % GICP_LUT_N;
        lut[m.lut_base + i] = (v <= m.dims[axis] + 1) ? morton_lut_entry(m, i) : 0;
    }
    // ---- keys (padded to a power of two with keys that sort last) ----
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = tid; i < np2; i += SMALL_GRID_THREADS) {
        unsigned key = 0xffffffffu;
        if (i < n) {
            int c[3] = {0, 0, 0};
#pragma unroll
            for (int a = 0; a < D; ++a) {
                const int v = cell_coord((double)pts[(size_t)(b + i) * D + a], m.origin[a], m.inv_h);
                c[a] = min(max(v, 0), m.dims[a] - 1);
            }
            key = ((unsigned)morton_code(c[0], c[1], c[2], m.bits[0], m.bits[1], m.bits[2]) << 11) | (unsigned)i;
        }
        s_key[i] = key;
    }
    __syncthreads();
    // ---- bitonic sort: ascending (cell, input index) = the stable order of the radix sort ----
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += SMALL_GRID_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned a0 = s_key[i], a1 = s_key[ixj];
                    const bool up = (i & k) == 0;
                    if ((a0 > a1) == up) { s_key[i] = a1; s_key[ixj] = a0; }
                }
            }
            __syncthreads();
        }
    }
    // ---- gather ----
    for (int i = tid; i < n; i += SMALL_GRID_THREADS) {
        const int g = (int)(s_key[i] & 2047u);
        PRec<Real> r;
        r.x = pts[(size_t)(b + g) * D + 0];
        r.y = pts[(size_t)(b + g) * D + 1];
        r.z = (D == 3) ? pts[(size_t)(b + g) * D + (D - 1)] : Real(0);
        r.idx = g;
        spts[b + i] = r;
        inv_perm[b + g] = b + i;
    }
    // ---- cell table: cell_start[cell_base + c] = b + (number of points with a cell code < c) ----
    const int n_cells = (int)budget + (is_last_cloud_total_cells && cloud == gridDim.x - 1 ? 1 : 0);
    for (int c = tid; c < n_cells; c += SMALL_GRID_THREADS) {
        int l = 0, r = n;   // first position whose cell code is >= c
        while (l < r) {
            const int mid = (l + r) >> 1;
            if ((int)(s_key[mid] >> 11) < c) l = mid + 1; else r = mid;
        }
        cell_start[m.cell_base + c] = b + l;
    }
}

}  // namespace gicp
