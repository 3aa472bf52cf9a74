// Instantiations of the extraction kernels for N = 4 limbs (K in [97, 128]).
#include "extract_kernels.cuh"
namespace kmc {
KMC_DEFINE_LAUNCHER_TABLE(get_extract_launcher_n4, 4)
}
