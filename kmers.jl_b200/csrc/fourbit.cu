// fourbit.cu -- 4-bit sources (FourToTwo).  Filled in after the 2-bit path is parity-green.
#include "fourbit.h"

namespace kmc {

int32_t extract_device_4bit(kmc_ctx *ctx, const kmc_seqs *, int32_t, int32_t, uint32_t, const kmc_out *, kmc_result *,
                            cudaStream_t, bool)
{
    ctx->last_error = "4-bit sources are not implemented yet";
    return KMC_E_UNSUPPORTED;
}
int32_t count_unambiguous_4bit(kmc_ctx *ctx, const kmc_seqs *, int32_t, uint64_t *, cudaStream_t)
{
    ctx->last_error = "4-bit sources are not implemented yet";
    return KMC_E_UNSUPPORTED;
}
int32_t extract_host_4bit(kmc_ctx *ctx, const kmc_seqs *, int32_t, int32_t, uint32_t, const kmc_out *, kmc_result *)
{
    ctx->last_error = "4-bit sources are not implemented yet";
    return KMC_E_UNSUPPORTED;
}

} // namespace kmc
