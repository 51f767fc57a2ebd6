// tcgen05 implicit-GEMM convolution (placeholder until the kernel lands; every layer uses the direct kernel).
#include "net.cuh"
namespace pcls {
int Net::tc_prepare() { return PCLS_OK; }
int Net::tc_launch(ConvLayer&, const ConvParams&, int, cudaStream_t) { set_error("tc path not built"); return PCLS_ERR_STATE; }
void Net::tc_release() {}
}  // namespace pcls
