// Device-side pieces of the zoom-in engine around the network:
//   * rasterize_triangles: the rendering half of triangulate_corr (inference_helper.py:293-308) - the reference draws
//     the Delaunay triangles of the source points with OpenGL, vertex colour = target coordinates; here one thread
//     block per triangle walks the triangle's bounding box and writes the barycentric interpolation at every pixel
//     centre it covers (top-left fill rule, so shared edges are written exactly once and the result is deterministic).
#include "common.cuh"

namespace cotr {

namespace {

struct Vtx { float x, y, u, v; };

// edge function of (a -> b) at p, in double: the coordinates are fp32 pixel positions, so the products are exact
__device__ __forceinline__ double edge_fn(double ax, double ay, double bx, double by, double px, double py) {
    return (bx - ax) * (py - ay) - (by - ay) * (px - ax);
}
// top-left rule for a counter-clockwise triangle in a y-down image: an edge owns the pixels exactly on it when it is a
// "left" edge (going down) or a horizontal "top" edge (going left)
__device__ __forceinline__ bool owns_edge(double ax, double ay, double bx, double by) {
    const double dx = bx - ax, dy = by - ay;
    return dy > 0.0 || (dy == 0.0 && dx < 0.0);
}

__global__ void __launch_bounds__(256) rasterize_triangles_kernel(const Vtx* __restrict__ tris, int n_tri, int H, int W, float2* __restrict__ out) {
    for (int t = blockIdx.x; t < n_tri; t += gridDim.x) {
        Vtx a = tris[3 * t], b = tris[3 * t + 1], c = tris[3 * t + 2];
        double area = edge_fn(a.x, a.y, b.x, b.y, c.x, c.y);
        if (area == 0.0) continue;                         // degenerate
        if (area < 0.0) { const Vtx tmp = b; b = c; c = tmp; area = -area; }
        const float minx = fminf(a.x, fminf(b.x, c.x)), maxx = fmaxf(a.x, fmaxf(b.x, c.x));
        const float miny = fminf(a.y, fminf(b.y, c.y)), maxy = fmaxf(a.y, fmaxf(b.y, c.y));
        // pixel (ix, iy) is sampled at its centre (ix + 0.5, iy + 0.5)
        int x0 = (int)floorf(minx - 0.5f), x1 = (int)ceilf(maxx - 0.5f);
        int y0 = (int)floorf(miny - 0.5f), y1 = (int)ceilf(maxy - 0.5f);
        x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, W - 1); y1 = min(y1, H - 1);
        if (x1 < x0 || y1 < y0) continue;
        const int bw = x1 - x0 + 1, n = bw * (y1 - y0 + 1);
        const bool own_ab = owns_edge(a.x, a.y, b.x, b.y), own_bc = owns_edge(b.x, b.y, c.x, c.y), own_ca = owns_edge(c.x, c.y, a.x, a.y);
        const double inv = 1.0 / area;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int iy = y0 + i / bw, ix = x0 + i % bw;
            const double px = ix + 0.5, py = iy + 0.5;
            const double wa = edge_fn(b.x, b.y, c.x, c.y, px, py);      // weight of vertex a
            const double wb = edge_fn(c.x, c.y, a.x, a.y, px, py);
            const double wc = edge_fn(a.x, a.y, b.x, b.y, px, py);
            const bool in = (wa > 0.0 || (wa == 0.0 && own_bc)) && (wb > 0.0 || (wb == 0.0 && own_ca)) && (wc > 0.0 || (wc == 0.0 && own_ab));
            if (!in) continue;
            const double la = wa * inv, lb = wb * inv, lc = wc * inv;
            out[(size_t)iy * W + ix] = make_float2((float)(la * a.u + lb * b.u + lc * c.u), (float)(la * a.v + lb * b.v + lc * c.v));
        }
    }
}

}  // namespace

int rasterize_triangles_launch(const float* tris, int n_tri, int H, int W, float* out, cudaStream_t s) {
    COTR_CHECK(out != nullptr && H > 0 && W > 0 && n_tri >= 0 && (n_tri == 0 || tris != nullptr), "cotr_rasterize_triangles: bad arguments");
    COTR_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)H * W * 2 * sizeof(float), s));
    if (n_tri == 0) return 0;
    const int grid = n_tri < 148 * 8 ? n_tri : 148 * 8;
    rasterize_triangles_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const Vtx*>(tris), n_tri, H, W, reinterpret_cast<float2*>(out));
    COTR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cotr
