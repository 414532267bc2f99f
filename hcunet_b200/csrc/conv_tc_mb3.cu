// Specialised instantiations of conv_tc_kernel for 3 M-blocks per CTA (see conv_tc_kernel.cuh: struct Var).
#include "conv_tc_kernel.cuh"

namespace hcu {
namespace tc {
HCU_TC_DEFINE_VARIANTS(3)
}  // namespace tc
}  // namespace hcu
