// Instantiations of the extraction kernels for N = 1 limbs (K in [1, 32]).
#include "extract_kernels.cuh"
namespace kmc {
KMC_DEFINE_LAUNCHER_TABLE(get_extract_launcher_n1, 1)
KMC_DEFINE_DIGEST_TABLE(get_digest_launcher_n1, 1)
}
