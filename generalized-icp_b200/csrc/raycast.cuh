// 8f row 2: batched 2-D LiDAR ray caster - the input generator of the reference's robot demo
// (robot-visualization.py:42-120 cast_ray / ray_line_intersection / ray_circle_intersection,
// 222-237 scan loop), one thread per ray, all poses of a trajectory in one launch.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gicp {

struct RayCastArgs {
    const double* poses;     // [n_poses][3]  x, y, yaw (degrees; the demo keeps yaw an int)
    int n_poses, num_rays;
    const double* segments;  // [n_seg][4]    x3, y3, x4, y4 (rectangle edges, robot-visualization.py:52-57)
    int n_seg;
    const double* circles;   // [n_circ][3]   cx, cy, r
    int n_circ;
    double max_range;        // MAX_RAY_RANGE
    const double* noise;     // optional [n_poses][num_rays]: additive range noise (random.uniform(-NOISE, NOISE))
    double* rel_xy;          // [n_poses][num_rays][2] robot-relative hit point (robot-visualization.py:227-230)
    int* hit;                // [n_poses][num_rays]   1 if the ray hit something (and the range is non-zero)
};

__global__ void raycast_kernel(const RayCastArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_poses * a.num_rays) return;
    const int pose = i / a.num_rays, ray = i % a.num_rays;
    const double x1 = a.poses[pose * 3], y1 = a.poses[pose * 3 + 1], yaw = a.poses[pose * 3 + 2];
    const double angle = yaw + (double)(ray * (360 / a.num_rays));        // range(yaw, yaw + 360, 360 // NUM_RAYS)
    const double rad = angle * (M_PI / 180.0);
    const double x2 = x1 + a.max_range * cos(rad), y2 = y1 + a.max_range * sin(rad);
    double best = INFINITY;
    for (int s = 0; s < a.n_seg; ++s) {                                    // ray_line_intersection, :79-93
        const double x3 = a.segments[s * 4], y3 = a.segments[s * 4 + 1], x4 = a.segments[s * 4 + 2], y4 = a.segments[s * 4 + 3];
        const double denom = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
        if (denom == 0.0) continue;
        const double t = ((x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)) / denom;
        const double u = -((x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3)) / denom;
        if (t >= 0.0 && t <= 1.0 && u >= 0.0 && u <= 1.0) {
            const double d = hypot(t * (x2 - x1), t * (y2 - y1));
            best = fmin(best, d);
        }
    }
    for (int c = 0; c < a.n_circ; ++c) {                                   // ray_circle_intersection, :95-120
        const double cx = a.circles[c * 3], cy = a.circles[c * 3 + 1], r = a.circles[c * 3 + 2];
        const double dx = x2 - x1, dy = y2 - y1, fx = x1 - cx, fy = y1 - cy;
        const double A = dx * dx + dy * dy, B = 2.0 * (fx * dx + fy * dy), C = (fx * fx + fy * fy) - r * r;
        double disc = B * B - 4.0 * A * C;
        if (disc >= 0.0) {
            disc = sqrt(disc);
            const double t1 = (-B - disc) / (2.0 * A), t2 = (-B + disc) / (2.0 * A);
            if (t1 >= 0.0 && t1 <= 1.0) best = fmin(best, hypot(t1 * dx, t1 * dy));
            if (t2 >= 0.0 && t2 <= 1.0) best = fmin(best, hypot(t2 * dx, t2 * dy));
        }
    }
    bool ok = best < INFINITY;
    double d = best;
    if (ok && a.noise) d += a.noise[i];                                    // :73-75
    ok = ok && d != 0.0;                                                   // "if distance:" :226
    const double rel = (angle - yaw) * (M_PI / 180.0);
    a.rel_xy[(size_t)i * 2] = ok ? d * cos(rel) : 0.0;
    a.rel_xy[(size_t)i * 2 + 1] = ok ? d * sin(rel) : 0.0;
    a.hit[i] = ok ? 1 : 0;
}

}  // namespace gicp
