// 4-bit / ASCII source (FourToTwo, AsciiEncode) instantiations for N = 1 limbs: strict
// Fw/FwRv/Canonical with the uncertain-symbol check, and the ordered compaction of UnambiguousKmers.
#include "compact_kernels.cuh"
#include "lincompact.cuh"
namespace kmc {
KMC_DEFINE_FOURBIT_TABLES(get_strict4_launcher_n1, 1)
KMC_DEFINE_COMPACT_TABLE(get_compact_launcher_n1, 1)
KMC_DEFINE_LIN_TABLE(get_lin_launcher_n1, 1)
}
