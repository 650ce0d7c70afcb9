// pcl::VoxelGrid<PointXYZI> on an arbitrary cloud (laserMapping.cpp:543-549: the scan stacks) -- general GPU
// implementation (stable LSD radix sort on the voxel index + ordered float centroids).
#include "ctx.h"
extern "C" int aloam_voxel_filter_impl(aloam_ctx* c, aloam_cloud_view in, float leaf, aloam_cloud_view* out);
#include "voxel_impl.inc"
