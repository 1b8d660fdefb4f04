/*
 * sw_types.h — plain-C value types shared by the CUDA library (swgpu.h), the CPU oracle
 * (oracle/) and the host adapters.  No torch / CUDA types appear here.
 *
 * Reference concepts mirrored (paths relative to /root/reference/schwarzwald):
 *   sw_sampling        <-> names accepted by make_sampling_strategy_from_name,
 *                          core/tiling/Sampling.h:774-791
 *   sw_tiling          <-> TilingStrategy {Accurate, Fast}, core/process/Tiler.cpp:189-198
 *   sw_params          <-> TilerMetaParameters, core/process/Tiler.h:64-75 (+ dataset bounds,
 *                          Tiler.cpp:185-187) and the indexing thread count that
 *                          TilingAlgorithmV3::estimate_start_node_level_in_octree depends on
 *                          (core/tiling/TilingAlgorithms.cpp:1473-1535)
 *   sw_node            <-> OctreeNodeIndex64 {index, levels}
 *                          (core/datastructures/OctreeNodeIndex.h:118-579) plus the slice of the
 *                          node-major point-id array that persist_points() would receive
 *                          (core/io/PointsPersistence.h:23-31)
 */
#ifndef SW_TYPES_H
#define SW_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum sw_sampling {
  SW_RANDOM_GRID = 0,  /* RandomSortedGridSampling, Sampling.h:187-308 */
  SW_GRID_CENTER = 1,  /* GridCenterSampling,       Sampling.h:314-416 */
  SW_MIN_DISTANCE = 2, /* PoissonDiskSampling,      Sampling.h:421-471 */
  SW_JITTERED = 3,     /* JitteredSampling,         Sampling.h:598-759 */
  SW_MIN_DISTANCE_FAST = 4 /* AdaptivePoissonDiskSampling, Sampling.h:477-542, with the CLI's density
                            * function (process/TilerProcess.cpp:500-508): every 4th point of a node is
                            * analysed at the root, every 2nd at node level 0, every point below */
} sw_sampling;

typedef enum sw_tiling {
  SW_ACCURATE = 0, /* TilingAlgorithmV1, TilingAlgorithms.cpp:577-626 */
  SW_FAST = 1      /* TilingAlgorithmV3, TilingAlgorithms.cpp:1207-1784 */
} sw_tiling;

/* node flags */
#define SW_NODE_TAKE_ALL 1u      /* all points of the visit were taken (count <= max_points_per_node) */
#define SW_NODE_TERMINAL 2u      /* tile_terminal_node: level >= max_level, stored unsampled */
#define SW_NODE_RECONSTRUCTED 4u /* FAST finalize: re-sampled copy of the children's points */
#define SW_NODE_DEEP 8u          /* swgpu_set_deep_node_policy(h, 1): the node sits where the reference would re-index its
                                  * points with the node as the new root (TilingAlgorithms.cpp:444-483); it holds all
                                  * remaining points unsampled (set together with SW_NODE_TERMINAL) */

typedef struct sw_params {
  int32_t sampling;             /* sw_sampling */
  int32_t tiling;               /* sw_tiling */
  float spacing_at_root;        /* TilerMetaParameters::spacing_at_root (float!) */
  uint32_t max_depth;           /* TilerMetaParameters::max_depth; CLI effectively uses 100 */
  uint64_t max_points_per_node; /* default 20000 */
  double bounds_min[3];         /* cubic dataset bounds (AABB::makeCubic, math/AABB.h:50-61); the tiling entry
                                 * points refuse boxes whose extents differ (SW_ERR_INVALID_ARGUMENT) */
  double bounds_max[3];
  uint32_t concurrency;         /* FAST: num_indexing_threads of the reference run being matched */
  uint32_t reserved;
} sw_params;

typedef struct sw_node {
  uint64_t index;  /* octants packed 3 bits per level, deepest level in the low bits */
  uint32_t levels; /* 0 = root "r"; node level in the reference's sense is levels-1 */
  uint32_t flags;
  uint64_t first;  /* offset of the node's first point in the node-major id array */
  uint64_t count;
} sw_node;

/* LAS record -> PointBuffer position, as the reference's reader and tiler transformation compute it
 * (SURVEY.md section 8 f2):
 *   position_from_las_point, core/io/LASFile.cpp:79-94:
 *       p = offset + X * scale;  p = std::min(header_max, std::max(header_min, p))
 *   TilerProcess's point transformation, core/process/TilerProcess.cpp:552-559 (3DTILES output with an
 *   identity SRS transform): p -= center; p = (double)(float)p     when shift_to_center != 0
 * center = cubic_bounds.getCenter() = min + extent / 2 (core/math/AABB.h:70), computed by the caller. */
typedef struct sw_las_transform {
  double scale[3];      /* laszip_header::x/y/z_scale_factor */
  double offset[3];     /* laszip_header::x/y/z_offset */
  double header_min[3]; /* laszip_header::min_x/y/z */
  double header_max[3]; /* laszip_header::max_x/y/z */
  double center[3];     /* only read when shift_to_center != 0 */
  int32_t shift_to_center;
  int32_t reserved;
} sw_las_transform;

/* Header values LASPersistence::persist_points writes for one node (core/io/LASPersistence.h:119-131):
 * offset = min = node bounds min, max = node bounds max, one scale for the three axes
 * (compute_las_scale_from_bounds, core/io/LASPersistence.cpp:17-28). */
typedef struct sw_las_node_header {
  double offset[3];
  double scale;
  double max[3];
  double reserved;
} sw_las_node_header;

/* error codes shared by swgpu_* and swo_* */
#define SW_OK 0
#define SW_ERR_INVALID_ARGUMENT 1
#define SW_ERR_CUDA 2
#define SW_ERR_OUT_OF_MEMORY 3
#define SW_ERR_JITTER_GRID_TOO_SMALL 4 /* Sampling.h:632-635 "Grids smaller than 16x16 ..." */
#define SW_ERR_JITTER_NODE_TOO_SMALL 5 /* Sampling.h:641-653 grid_level >= 21 */
#define SW_ERR_DEEP_REROOT 6           /* TilingAlgorithms.cpp:444-483 path, not supported */
#define SW_ERR_EMPTY_NODE 7            /* TilingAlgorithms.cpp:253-259 */
#define SW_ERR_STATE 8
#define SW_ERR_TOO_FEW_POINTS 9        /* Parallel.h:181-186: fewer points than indexing threads */
#define SW_ERR_COLLECTIVE 10           /* multi-GPU: the caller-supplied all-reduce hook failed */

#ifdef __cplusplus
}
#endif
#endif
