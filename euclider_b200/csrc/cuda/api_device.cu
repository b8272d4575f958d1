// temporary stub
#include "euclider_b200.h"
#include "error.h"
extern "C" {
int eucl_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
int eucl_scene_create(const EuclFlatScene*, int, EuclScene**) { return eucl::fail(EUCL_ERR_NO_DEVICE, "stub"); }
void eucl_scene_destroy(EuclScene*) {}
int eucl_render(EuclScene*, const EuclCamera*, const EuclRenderOpts*, uint8_t*, int32_t*, EuclStats*) { return eucl::fail(EUCL_ERR_NO_DEVICE, "stub"); }
int eucl_render_device(EuclScene*, const EuclCamera*, const EuclRenderOpts*, void*, void*, EuclStats*) { return eucl::fail(EUCL_ERR_NO_DEVICE, "stub"); }
int eucl_ipc_export(void*, uint8_t*) { return eucl::fail(EUCL_ERR_NO_DEVICE, "stub"); }
int eucl_ipc_open(const uint8_t*, int, void**) { return eucl::fail(EUCL_ERR_NO_DEVICE, "stub"); }
int eucl_ipc_close(void*) { return eucl::fail(EUCL_ERR_NO_DEVICE, "stub"); }
int eucl_fp64_peak(int, double*, double*, double*) { return eucl::fail(EUCL_ERR_NO_DEVICE, "stub"); }
}
