// wlsqm_grid.h -- device-side view of the uniform search grid over a point set (wlsqm_grid.cu).
//
// The grid replaces, on the device, what the reference's callers and ExpertSolver do with
// scipy.spatial.cKDTree on the host:
//   neighbourhood construction  examples/expertsolver_example.py:51-66, examples/wlsqm_example.py:100-132
//   nearest-model search        wlsqm/fitter/expert.pyx:676-681 (prep_interpolate), :837 (tree.query(x, k=1))
//   ball search                 wlsqm/fitter/expert.pyx:898-911 (query_ball_tree, mode='continuous')
#pragma once
#include <cuda_runtime.h>

namespace wlsqm {

struct GridView {
    int dim;
    long long n;                 // points
    double lo[3];                // lower corner of the bounding box
    double h, inv_h;             // cell edge
    int dims[3];                 // cells per axis (1 on unused axes)
    const int* cell_start;       // [ncells + 1] first sorted position of every cell
    const int* sorted_idx;       // [n] original index of the point at each sorted position (ascending inside a cell)
    const double* sorted_x;      // [n][dim] coordinates in sorted order
};

__host__ __device__ inline int grid_cell_coord(const GridView& g, int axis, double x) {
    const double t = (x - g.lo[axis]) * g.inv_h;
    int c = t > 0.0 ? (t < 2.0e9 ? (int)t : 2000000000) : 0;     // NaN and negatives -> 0
    return c < g.dims[axis] ? c : g.dims[axis] - 1;
}

}  // namespace wlsqm
