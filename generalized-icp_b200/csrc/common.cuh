// Shared types and device helpers for libgicp_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include <math.h>

namespace gicp {

// One uniform grid per cloud.  Cells are numbered along a Morton (Z-order) curve whose bit
// budget per axis follows the grid's extent (bits[a] = ceil(log2(dims[a]))), so that points that
// are consecutive in the sorted array are spatially compact in all axes.  Each cell is one
// contiguous run of the sorted point array (cell_start[code] .. cell_start[code + 1]).
struct CloudMeta {
    double origin[3];
    double h;        // cell edge actually used (>= requested; enlarged to fit the cell budget)
    double inv_h;
    int dims[3];
    int bits[3];     // Morton bits per axis
    int cell_base;   // first entry of this cloud in the shared cell_start table
    int pt_begin;    // row range of this cloud in the concatenated arrays
    int pt_end;
    int lut_base;    // first entry of this cloud's per-axis Morton tables: lut[lut_base + axis * GICP_LUT_N + v]
};
constexpr int GICP_LUT_N = 1024;   // per axis (GICP_MAX_AXIS_BITS bits)

// anisotropic Morton code: bit b of every axis that still has bits, x first
__host__ __device__ __forceinline__ int morton_code(int x, int y, int z, int bx, int by, int bz) {
    int code = 0, pos = 0;
#pragma unroll
    for (int b = 0; b < 10; ++b) {
        if (b < bx) { code |= ((x >> b) & 1) << pos; ++pos; }
        if (b < by) { code |= ((y >> b) & 1) << pos; ++pos; }
        if (b < bz) { code |= ((z >> b) & 1) << pos; ++pos; }
    }
    return code;
}
constexpr int GICP_MAX_AXIS_BITS = 10;  // <= 1024 cells per axis

// Sorted point record: coordinates + cloud-local index of the point in the
// caller's array.  16 B (f32) / 32 B (f64): multiples of 16 B so that runs can be
// moved with cp.async.bulk (TMA) without padding.
template <typename Real> struct PRec;
template <> struct __align__(16) PRec<float> { float x, y, z; int idx; };
template <> struct __align__(32) PRec<double> { double x, y, z; long long idx; };

enum PairStatus { PAIR_ACTIVE = 0, PAIR_CONVERGED = 1, PAIR_MAXITER = 2 };

struct PairState {
    double R[9];        // current rotation (row-major, top-left dim x dim used)
    double t[3];        // current translation
    double theta;       // dim 2: accumulated angle (the reference rebuilds T from it, gicp.py:166)
    double last_loss;   // gicp.py:110,165
    double mu[3];       // centring point of the reduced form (target bbox centre)
    double Rp[9];       // transform of the previous outer iteration (how far each point moved since)
    double tp[3];
    int iter;           // outer iterations completed
    int status;
    int converged_at;   // gicp.py:161 or -1
    int pad;
};

template <int D> struct Dim {
    static constexpr int NP = D + 1;                 // p~ = (1, p)
    static constexpr int NS = D * (D + 1) / 2;       // symmetric W entries
    static constexpr int NAB = NP * (NP + 1) / 2;    // symmetric p~ p~^T entries
    static constexpr int NH = NAB * NS;              // 60 / 18
    static constexpr int NG = D * NP;                // 12 / 6
    static constexpr int NQ = NH + NG;               // accumulators besides loss and count
    static constexpr int NRED = (D == 3) ? 80 : 32;  // padded row written per pair
    static constexpr int NPAR = (D == 3) ? 6 : 3;    // (t, rotation)
};

__host__ __device__ inline int symidx(int n, int a, int b) {
    // index of (a,b), a<=b, in the row-major upper triangle of an n x n symmetric matrix
    if (a > b) { int t = a; a = b; b = t; }
    return a * n - a * (a - 1) / 2 + (b - a);
}

__device__ __forceinline__ int cell_coord(double x, double origin, double inv_h) {
    double u = floor((x - origin) * inv_h);
    u = fmin(fmax(u, -1073741824.0), 1073741824.0);
    return (int)u;
}

// exact squared distance, in the association the oracle uses: (dx*dx + dy*dy) + dz*dz,
// every operation rounded separately (no FMA contraction).
__device__ __forceinline__ double exact_d2(double dx, double dy, double dz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
__device__ __forceinline__ double exact_d2(double dx, double dy) {
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// ---- TMA (bulk async copy) + mbarrier helpers: raw PTX, sm_90+/sm_100a ----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// make the barrier's initialisation visible to the async proxy (the TMA unit): the documented pattern
// (CUTLASS fence_barrier_init).  A lighter fence.proxy.async was tried; it measured no faster and one bench
// run in ~10 died with "unspecified launch failure" while it was in, so the documented fence stays.
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (SASS: UBLKCP), completion signalled on the mbarrier
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- small closed-form linear algebra (fp64) -------------------------------------------------
// Unit eigenvector of the smallest eigenvalue of the symmetric 3x3 matrix
// [a00 a01 a02; . a11 a12; . . a22].  Degenerate (isotropic) input -> (1,0,0), which is
// what LAPACK's eigh returns for a multiple of the identity.
__host__ __device__ inline void smallest_eigvec3(double a00, double a01, double a02, double a11, double a12, double a22,
                                        double n[3]) {
    const double scale = fmax(fmax(fabs(a00), fabs(a11)), fmax(fabs(a22), fmax(fabs(a01), fmax(fabs(a02), fabs(a12)))));
    n[0] = 1.0; n[1] = 0.0; n[2] = 0.0;
    if (!(scale > 0.0) || !isfinite(scale)) return;
    const double is = 1.0 / scale;
    a00 *= is; a01 *= is; a02 *= is; a11 *= is; a12 *= is; a22 *= is;
    const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
    const double q = (a00 + a11 + a22) / 3.0;
    const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
    const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
    if (!(p2 > 1e-30)) return;  // multiple of the identity
    const double p = sqrt(p2 / 6.0);
    const double ip = 1.0 / p;
    const double c00 = b00 * ip, c01 = a01 * ip, c02 = a02 * ip, c11 = b11 * ip, c12 = a12 * ip, c22 = b22 * ip;
    double r = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02));
    r = fmin(1.0, fmax(-1.0, r));
    const double phi = acos(r) / 3.0;
    // smallest eigenvalue
    const double lam = q + 2.0 * p * cos(phi + 2.0943951023931954923);
    // eigenvector: largest cross product of two rows of (A - lam I)
    const double r0[3] = {a00 - lam, a01, a02}, r1[3] = {a01, a11 - lam, a12}, r2[3] = {a02, a12, a22 - lam};
    double v0[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
    double v1[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
    double v2[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    const double n0 = v0[0] * v0[0] + v0[1] * v0[1] + v0[2] * v0[2];
    const double n1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
    const double n2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
    const double* best = v0;
    double nb = n0;
    if (n1 > nb) { best = v1; nb = n1; }
    if (n2 > nb) { best = v2; nb = n2; }
    if (!(nb > 1e-300)) return;  // rank <= 1: smallest eigenspace is 2-dimensional, keep the fallback
    double v[3] = {best[0], best[1], best[2]};
    // normalise; the largest of the three cross products is the best conditioned one in fp64
    const double inv = rsqrt(nb);
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
    n[0] = v[0]; n[1] = v[1]; n[2] = v[2];
}

// inverse of a symmetric 3x3 (entries 00,01,02,11,12,22) -> same layout
template <typename T> __device__ __forceinline__ void sym_inv3(const T m[6], T w[6]) {
    const T c00 = m[3] * m[5] - m[4] * m[4];
    const T c01 = m[2] * m[4] - m[1] * m[5];
    const T c02 = m[1] * m[4] - m[2] * m[3];
    const T det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const T id = T(1) / det;
    w[0] = c00 * id; w[1] = c01 * id; w[2] = c02 * id;
    w[3] = (m[0] * m[5] - m[2] * m[2]) * id;
    w[4] = (m[1] * m[2] - m[0] * m[4]) * id;
    w[5] = (m[0] * m[3] - m[1] * m[1]) * id;
}
// inverse of a symmetric 2x2 (00,01,11)
template <typename T> __device__ __forceinline__ void sym_inv2(const T m[3], T w[3]) {
    const T det = m[0] * m[2] - m[1] * m[1];
    const T id = T(1) / det;
    w[0] = m[2] * id; w[1] = -m[1] * id; w[2] = m[0] * id;
}

}  // namespace gicp
