/*
 * ORACLE — test infrastructure only. Never imported, linked or executed by the product path
 * (efg_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may use it.
 *
 * CPU restatement (plain C, fp32) of the reference's first-come hard voxelizer and its dynamic
 * variant.  Follows, statement by statement:
 *   efg/operators/src/voxelize/voxelization_cpu.cpp:8-40   dynamic_voxelize_kernel
 *   efg/operators/src/voxelize/voxelization_cpu.cpp:44-96  hard_voxelize_kernel
 *   efg/operators/src/voxelize/voxelization_cpu.cpp:105-142 hard_voxelize_cpu (grid size, dense map)
 * which is the same loop as efg/geometry/point_cloud_ops.py:6-53 (numba), the voxelizer the
 * playground configs actually run (efg/data/augmentations/extend_3d.py:256-283).
 * Pinned against both reference implementations run in the build container: tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* voxelization_cpu.cpp:119-122: grid_size[i] = round((range[3+i] - range[i]) / voxel_size[i]) in fp32 */
void oracle_grid_size(const float* voxel_size, const float* coors_range, int32_t* grid_xyz) {
  for (int i = 0; i < 3; ++i) {
    float q = (coors_range[3 + i] - coors_range[i]) / voxel_size[i];
    grid_xyz[i] = (int32_t)roundf(q);
  }
}

/* voxelization_cpu.cpp:8-40 — coors[i] = (z,y,x) or (-1,-1,-1) */
void oracle_dynamic_voxelize(const float* points, int64_t num_points, int num_features, const float* voxel_size,
                             const float* coors_range, int32_t* coors) {
  int32_t grid[3];
  oracle_grid_size(voxel_size, coors_range, grid);
  for (int64_t i = 0; i < num_points; ++i) {
    int failed = 0;
    int32_t coor[3];
    for (int j = 0; j < 3; ++j) {
      /* fp32 subtraction, fp32 true division, floor (:23) */
      volatile float diff = points[i * num_features + j] - coors_range[j];
      volatile float quot = diff / voxel_size[j];
      float f = floorf(quot);
      if (!(f >= 0.0f) || !(f < (float)grid[j])) { /* (:25) c < 0 || c >= grid; NaN treated as out of range */
        failed = 1;
        break;
      }
      coor[2 - j] = (int32_t)f; /* reversed: (z,y,x) (:29) */
    }
    for (int k = 0; k < 3; ++k) coors[i * 3 + k] = failed ? -1 : coor[k];
  }
}

/*
 * voxelization_cpu.cpp:44-96.  Outputs must hold max_voxels rows (or num_points rows when
 * max_voxels == -1); voxels and num_points_per_voxel must be zero-initialised by the caller,
 * exactly as efg/operators/voxelize.py:39-41 does.  Returns voxel_num.
 */
int64_t oracle_hard_voxelize(const float* points, int64_t num_points, int num_features, const float* voxel_size,
                             const float* coors_range, int max_points, int max_voxels, float* voxels, int32_t* coors,
                             int32_t* num_points_per_voxel) {
  int32_t grid[3];
  oracle_grid_size(voxel_size, coors_range, grid);
  int32_t* temp_coors = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)(num_points > 0 ? num_points : 1));
  oracle_dynamic_voxelize(points, num_points, num_features, voxel_size, coors_range, temp_coors);

  /* dense coor_to_voxelidx[grid_z][grid_y][grid_x] = -1 (:127-128) */
  const size_t cells = (size_t)grid[2] * grid[1] * grid[0];
  int32_t* coor_to_voxelidx = (int32_t*)malloc(sizeof(int32_t) * cells);
  memset(coor_to_voxelidx, 0xFF, sizeof(int32_t) * cells);

  int64_t voxel_num = 0;
  for (int64_t i = 0; i < num_points; ++i) {
    const int32_t* c = temp_coors + i * 3;
    if (c[0] == -1) continue; /* (:71) */
    const size_t lin = ((size_t)c[0] * grid[1] + c[1]) * grid[0] + c[2];
    int32_t voxelidx = coor_to_voxelidx[lin];
    if (voxelidx == -1) { /* (:76-86) */
      voxelidx = (int32_t)voxel_num;
      if (max_voxels != -1 && voxel_num >= max_voxels) break;
      voxel_num += 1;
      coor_to_voxelidx[lin] = voxelidx;
      for (int k = 0; k < 3; ++k) coors[(size_t)voxelidx * 3 + k] = c[k];
    }
    const int32_t num = num_points_per_voxel[voxelidx]; /* (:89-95) */
    if (max_points == -1 || num < max_points) {
      for (int k = 0; k < num_features; ++k)
        voxels[((size_t)voxelidx * max_points + num) * num_features + k] = points[i * num_features + k];
      num_points_per_voxel[voxelidx] += 1;
    }
  }
  free(coor_to_voxelidx);
  free(temp_coors);
  return voxel_num;
}
