// scan-to-map kernels (K0 grid build, K5 5-NN, K6 line / plane fits) -- under construction in this commit
#include "../../include/aloam_b200.h"
#include "kernels.h"
extern "C" {
int aloam_map_upload_impl(aloam_ctx*, aloam_cloud_view, aloam_cloud_view) { return ALOAM_ERR_STATE; }
int aloam_mapping_register_impl(aloam_ctx*, aloam_cloud_view, aloam_cloud_view, double*, aloam_stats*) { return ALOAM_ERR_STATE; }
int aloam_voxel_filter_impl(aloam_ctx*, aloam_cloud_view, float, aloam_cloud_view*) { return ALOAM_ERR_STATE; }
}
