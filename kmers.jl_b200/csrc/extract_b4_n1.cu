// Instantiations of the extraction kernels for k-mers over 4-bit alphabets, N = 1 limbs (K in [1, 16]).
#include "extract_kernels.cuh"
namespace kmc {
KMC_DEFINE_KMER4_TABLE(get_kmer4_launcher_n1, 1)
}
