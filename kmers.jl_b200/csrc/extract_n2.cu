// Instantiations of the extraction kernels for N = 2 limbs (K in [33, 64]).
#include "extract_kernels.cuh"
namespace kmc {
KMC_DEFINE_LAUNCHER_TABLE(get_extract_launcher_n2, 2)
KMC_DEFINE_DIGEST_TABLE(get_digest_launcher_n2, 2)
}
