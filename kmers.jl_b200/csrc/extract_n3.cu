// Instantiations of the extraction kernels for N = 3 limbs (K in [65, 96]).
#include "extract_kernels.cuh"
namespace kmc {
KMC_DEFINE_LAUNCHER_TABLE(get_extract_launcher_n3, 3)
}
