// Warp-level candidate streaming shared by K2 (k-NN) and K3 (1-NN).
//
// A warp owns 32 queries that are consecutive in a Morton-sorted cloud, i.e. spatially compact.
// Every lane knows the box of grid cells that can hold what it is looking for (all cells that
// intersect the ball around its query).  The warp takes the union box of a group of lanes,
// enumerates the aligned 2x2x2 cell blocks that intersect it - such a block is ONE contiguous run
// of the Morton-sorted point array - lets every lane fetch the [start, end) range of one block,
// prefix-sums the lengths, moves the runs into the warp's shared-memory stage with TMA bulk
// copies (cp.async.bulk, completion on the warp's mbarrier) and calls the consumer once per
// staged candidate on all lanes (the candidate is a broadcast LDS.128).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace gicp {

template <typename Real> struct WarpStage {
    PRec<Real>* buf;   // capacity `cap` records
    uint64_t* bar;
    uint32_t phase;
    int cap;
};

struct NeedAll {};   // tag: no per-block culling

// Streams every aligned cell block intersecting the cell box [lo, hi] (already clamped to the grid,
// extent <= 64 cells per axis) - except the blocks that already intersected [plo, phi] when
// has_prev - through the stage and hands each staged window to consume(records, count)
// warp-synchronously (records are read as broadcast LDS.128).
// need(x0, y0, z0, x1, y1, z1): optional per-lane predicate on a block's cell range; a non-empty
// block that no lane needs is not staged.  All 32 lanes must call this together with identical boxes.
template <typename Real, typename F, typename Need = NeedAll>
__device__ __forceinline__ void stream_cells(const CloudMeta& m, const int* __restrict__ cell_start,
                                             const int* __restrict__ lut, const PRec<Real>* __restrict__ spts,
                                             const int lo[3], const int hi[3], const int plo[3], const int phi[3],
                                             bool has_prev, WarpStage<Real>& ws, int lane, F&& consume,
                                             Need need = Need()) {
    if (hi[0] < lo[0] || hi[1] < lo[1] || hi[2] < lo[2]) return;
    // block = 2 cells along every axis that has at least one Morton bit
    const int s0 = m.bits[0] > 0, s1 = m.bits[1] > 0, s2 = m.bits[2] > 0;
    const int b0 = lo[0] >> s0, b1 = lo[1] >> s1, b2 = lo[2] >> s2;
    const int nbx = (hi[0] >> s0) - b0 + 1, nby = (hi[1] >> s1) - b1 + 1, nbz = (hi[2] >> s2) - b2 + 1;
    const int cells_per_block = 1 << (s0 + s1 + s2);
    // previous block box (block coordinates); empty when !has_prev
    const int pb0 = has_prev ? plo[0] >> s0 : 1, pb1 = has_prev ? plo[1] >> s1 : 1, pb2 = has_prev ? plo[2] >> s2 : 1;
    const int pe0 = has_prev ? phi[0] >> s0 : 0, pe1 = has_prev ? phi[1] >> s1 : 0, pe2 = has_prev ? phi[2] >> s2 : 0;
    // per-axis Morton codes of the block origins, one entry per lane (block extents are <= 32)
    const int* L = lut + m.lut_base;
    const int tx = __ldg(L + min((b0 + lane) << s0, GICP_LUT_N - 1));
    const int ty = __ldg(L + GICP_LUT_N + min((b1 + lane) << s1, GICP_LUT_N - 1));
    const int tz = __ldg(L + 2 * GICP_LUT_N + min((b2 + lane) << s2, GICP_LUT_N - 1));
    const int n_blocks = nbx * nby * nbz;
    const float inv_x = 1.0f / (float)nbx, inv_y = 1.0f / (float)nby;
    for (int g0 = 0; g0 < n_blocks; g0 += 32) {
        const int e = min(g0 + lane, n_blocks - 1);
        const int q1 = (int)(((float)e + 0.5f) * inv_x);   // e / nbx (exact for these ranges)
        const int ix = e - q1 * nbx;
        const int iz = (int)(((float)q1 + 0.5f) * inv_y);  // q1 / nby
        const int iy = q1 - iz * nby;
        const int code = __shfl_sync(0xffffffffu, tx, ix) | __shfl_sync(0xffffffffu, ty, iy) |
                         __shfl_sync(0xffffffffu, tz, iz);
        int start = 0, len = 0;
        const bool seen = (b0 + ix >= pb0 && b0 + ix <= pe0) && (b1 + iy >= pb1 && b1 + iy <= pe1) &&
                          (b2 + iz >= pb2 && b2 + iz <= pe2);
        if (g0 + lane < n_blocks && !seen) {
            const int* cs = cell_start + m.cell_base + code;
            start = __ldg(cs);
            len = __ldg(cs + cells_per_block) - start;
        }
        if constexpr (!std::is_same<Need, NeedAll>::value) {
            // cull: every lane tests every non-empty block of this round against its own ball
            unsigned nonempty = __ballot_sync(0xffffffffu, len > 0);
            unsigned keep = 0;
            while (nonempty) {
                const int src = __ffs(nonempty) - 1;
                nonempty &= nonempty - 1;
                const int bx = (b0 + __shfl_sync(0xffffffffu, ix, src)) << s0;
                const int by = (b1 + __shfl_sync(0xffffffffu, iy, src)) << s1;
                const int bz = (b2 + __shfl_sync(0xffffffffu, iz, src)) << s2;
                if (__any_sync(0xffffffffu, need(bx, by, bz, bx + s0, by + s1, bz + s2))) keep |= 1u << src;
            }
            if (!((keep >> lane) & 1u)) len = 0;
        }
        const int incl = warp_incl_scan(len, lane);
        const int excl = incl - len;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        for (int w0 = 0; w0 < total; w0 += ws.cap) {
            const int n_win = min(ws.cap, total - w0);
            if (lane == 0) mbar_expect_tx(ws.bar, (uint32_t)(n_win * sizeof(PRec<Real>)));
            __syncwarp();
            const int l = max(excl, w0), h = min(excl + len, w0 + ws.cap);
            if (h > l)
                tma_load_1d(ws.buf + (l - w0), spts + start + (l - excl), (uint32_t)((h - l) * sizeof(PRec<Real>)),
                            ws.bar);
            mbar_wait(ws.bar, ws.phase);
            ws.phase ^= 1u;
            consume(ws.buf, n_win);
            __syncwarp();
        }
    }
}

// fp32 rounding slack of the cell boxes below: cell membership is decided in double, the box is
// rebuilt in float, so it is widened by a few ulps of the largest coordinate of the grid
__device__ __forceinline__ float cell_box_pad(const CloudMeta& m) {
    float c = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
        c = fmaxf(c, fmaxf(fabsf((float)m.origin[a]), fabsf((float)(m.origin[a] + m.dims[a] * m.h))));
    return 4e-6f * c + 1e-6f * (float)m.h;
}

// lower bound on the squared distance from a point to the box of cells [x0..x1] x [y0..y1] x [z0..z1]
__device__ __forceinline__ float cell_box_dist2(const CloudMeta& m, float pad, float px, float py, float pz, int x0,
                                                int y0, int z0, int x1, int y1, int z1) {
    const float h = (float)m.h;
    const float ox = (float)m.origin[0], oy = (float)m.origin[1], oz = (float)m.origin[2];
    const float lx = ox + x0 * h - pad, hx = ox + (x1 + 1) * h + pad;
    const float ly = oy + y0 * h - pad, hy = oy + (y1 + 1) * h + pad;
    const float lz = oz + z0 * h - pad, hz = oz + (z1 + 1) * h + pad;
    const float dx = fmaxf(0.f, fmaxf(lx - px, px - hx));
    const float dy = fmaxf(0.f, fmaxf(ly - py, py - hy));
    const float dz = fmaxf(0.f, fmaxf(lz - pz, pz - hz));
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// Greedy spatial grouping of a warp's queries: returns the mask of still-pending lanes whose cell
// lies within `reach` cells (Chebyshev) of the first pending lane's cell.  Morton-sorted chunks
// almost always form a single group; the split only bounds the union box when a chunk straddles a
// high-level Z-curve boundary (or a sparse region).
__device__ __forceinline__ unsigned next_group(unsigned pending, int cx, int cy, int cz, int reach) {
    const int leader = __ffs(pending) - 1;
    const int lx = __shfl_sync(0xffffffffu, cx, leader), ly = __shfl_sync(0xffffffffu, cy, leader),
              lz = __shfl_sync(0xffffffffu, cz, leader);
    const bool in = abs(cx - lx) <= reach && abs(cy - ly) <= reach && abs(cz - lz) <= reach;
    return __ballot_sync(0xffffffffu, in) & pending;
}

// union of the lanes' cell boxes over the lanes of `mask`, clamped to the grid
__device__ __forceinline__ void group_union(unsigned mask, int lane, const int mylo[3], const int myhi[3],
                                            const CloudMeta& m, int lo[3], int hi[3]) {
    const bool in = (mask >> lane) & 1u;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = max(warp_min(in ? mylo[a] : INT_MAX), 0);
        hi[a] = min(warp_max(in ? myhi[a] : INT_MIN), m.dims[a] - 1);
    }
}

}  // namespace gicp
