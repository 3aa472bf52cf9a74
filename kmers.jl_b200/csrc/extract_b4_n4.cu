// Instantiations of the extraction kernels for k-mers over 4-bit alphabets, N = 4 limbs (K in [49, 64]).
#include "extract_kernels.cuh"
namespace kmc {
KMC_DEFINE_KMER4_TABLE(get_kmer4_launcher_n4, 4)
}
