// placeholder (replaced by the tcgen05 kernel)
#include "common.cuh"
namespace evpp {
int launch_mlp_tc(const TcWeights&, const RayBatch&, float*, int, cudaStream_t) { return -1; }
int launch_mlp_tc_points(const TcWeights&, const float*, const float*, int64_t, float*, int, cudaStream_t) { return -1; }
size_t tc_weight_image_bytes() { return 16; }
void tc_pack_weights(const std::vector<std::vector<float>>&, int, std::vector<uint8_t>& image) { image.assign(16, 0); }
int tc_configure() { return 0; }
}
