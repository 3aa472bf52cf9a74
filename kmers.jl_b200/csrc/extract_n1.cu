// Instantiations of the extraction kernels for N = 1 limbs (K in [1, 32]).
#include "extract_kernels.cuh"
namespace kmc {
KMC_DEFINE_LAUNCHER_TABLE(get_extract_launcher_n1, 1)
KMC_DEFINE_DIGEST_TABLE(get_digest_launcher_n1, 1)

template <int NX> static AosLaunchFn pick_aos(bool fwrv, bool hash)
{
    if (fwrv) return hash ? &launch_extract_aos<NX, AOS_FWRV, true> : &launch_extract_aos<NX, AOS_FWRV, false>;
    return hash ? &launch_extract_aos<NX, AOS_INDEX, true> : &launch_extract_aos<NX, AOS_INDEX, false>;
}
AosLaunchFn get_aos_launcher_n1(int nx, bool fwrv, bool hash)
{
    switch (nx) { // ceil((2 K + 2 kAosGroup - 2) / 32) for K = 1 .. 32
    case 1: return pick_aos<1>(fwrv, hash);
    case 2: return pick_aos<2>(fwrv, hash);
    case 3: return pick_aos<3>(fwrv, hash);
    }
    return nullptr;
}
}
