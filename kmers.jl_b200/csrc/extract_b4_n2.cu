// Instantiations of the extraction kernels for k-mers over 4-bit alphabets, N = 2 limbs (K in [17, 32]).
#include "extract_kernels.cuh"
namespace kmc {
KMC_DEFINE_KMER4_TABLE(get_kmer4_launcher_n2, 2)
}
