// Specialised instantiations of conv_tc_kernel for 1 M-block per CTA (see conv_tc_kernel.cuh: struct Var).
#include "conv_tc_kernel.cuh"

namespace hcu {
namespace tc {
HCU_TC_DEFINE_VARIANTS(1)
}  // namespace tc
}  // namespace hcu
