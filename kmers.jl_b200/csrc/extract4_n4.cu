// 4-bit source (FourToTwo) instantiations for N = 4 limbs: strict Fw/FwRv/Canonical with the
// uncertain-symbol check.
#include "extract_kernels.cuh"
namespace kmc {
KMC_DEFINE_FOURBIT_TABLES(get_strict4_launcher_n4, 4)
}
