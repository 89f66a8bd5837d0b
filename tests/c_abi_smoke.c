/* Plain-C consumer of include/yolopost_b200.h: proves the boundary is a C ABI (no C++ / torch types) and that the host-only
 * entry points (versioning, workspace sizing, argument validation) work without a device.  Built and run by
 * tests/test_host_logic.py::test_c_program_links_against_the_library. */
#include <stdio.h>
#include <string.h>

#include "yolopost_b200.h"

int main(void) {
  if (ypb_abi_version() != YPB_ABI_VERSION) { printf("abi %d != %d\n", ypb_abi_version(), YPB_ABI_VERSION); return 1; }
  size_t small = ypb_nms_workspace_bytes(1, 8400, 8400, 300, 30000, YPB_NMS_GREEDY);
  size_t big = ypb_nms_workspace_bytes(64, 8400, 8400, 300, 30000, YPB_NMS_GREEDY);
  if (!(small > 0 && big > small)) { printf("workspace sizing\n"); return 2; }
  ypb_head_desc h;
  memset(&h, 0, sizeof h);
  if (ypb_decode_dense(&h, NULL, 0, 0, 0, NULL, YPB_F32, 0, 0, NULL) != YPB_ERR_INVALID_ARGUMENT) { printf("validation\n"); return 3; }
  if (strlen(ypb_last_error_string()) == 0) { printf("no error text\n"); return 4; }
  ypb_scale_xform xf = {1.f, 0.f, 0.f, 640.f, 640.f, 0.f, 0.f, 0.f};
  if (ypb_scale_rows(NULL, 0, 4, 1, 0, NULL, NULL, &xf, YPB_BOXES_XYXY, YPB_SCALE_PADDING, 0, NULL, 0, 0, 0, 0, NULL) != YPB_OK) return 5;
  if (ypb_peer_wait(NULL, 2, NULL, 0, 3, NULL, 0, NULL, NULL) != YPB_ERR_INVALID_ARGUMENT) return 6;
  if (ypb_dist2bbox(NULL, 0, 0, NULL, 0, 0, 0, NULL, 0, YPB_F32, 1, 4, 1, NULL, 0, 0, NULL) != YPB_ERR_INVALID_ARGUMENT) return 8;
  if (ypb_dfl_expectation(NULL, YPB_F32, 1, 8, 4, 0, 0, NULL, 0, 0, NULL) != YPB_ERR_UNSUPPORTED) return 9;
  if (ypb_peer_wait_copy(NULL, 2, NULL, 0, 3, NULL, 0, NULL, NULL, 0, NULL, NULL) != YPB_ERR_INVALID_ARGUMENT) return 10;
  if (sizeof(ypb_scale_xform) != 32 || YPB_MAX_PEERS != 8) return 7;
  printf("c abi ok, version %d, workspace C2 = %zu bytes\n", ypb_abi_version(), big);
  return 0;
}
