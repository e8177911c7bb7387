// tcgen05 edge kernel -- placeholder until the 3xTF32 kernel lands.
#include "egnn.cuh"

namespace b200md {
bool tc_edge_supported(const b200md_egnn_config&) { return false; }
int64_t tc_packed_floats(const b200md_egnn_config&) { return 0; }
int tc_pack_layer(const b200md_egnn_config&, const EgnnLayout&, int, float*, cudaStream_t) { return B200MD_ERR_UNSUPPORTED; }
int tc_launch_edge_layer(const b200md_egnn_config&, const EgnnLayout&, int, const float*, const EgnnWorkspace&, cudaStream_t) {
  return B200MD_ERR_UNSUPPORTED;
}
}  // namespace b200md
