// fourbit.h -- 4-bit (DNAAlphabet{4}) sources: FourToTwo recoding scheme (construction.jl:85-86).
#pragma once
#include "kmc_internal.h"

namespace kmc {

// FwKmers / FwRvIterator / CanonicalKmers over a 4-bit source (strict: an uncertain symbol is an
// EncodeError, FwKmers.jl:104-115, CanonicalKmers.jl:131-144) and UnambiguousKmers (skip/restart,
// UnambiguousKmers.jl:134-148).  Device-resident descriptors.
int32_t extract_device_4bit(kmc_ctx *ctx, const kmc_seqs *s, int32_t k, int32_t mode, uint32_t flags,
                            const kmc_out *out, kmc_result *res, cudaStream_t stream, bool sync);
int32_t count_unambiguous_4bit(kmc_ctx *ctx, const kmc_seqs *s, int32_t k, uint64_t *n_out, cudaStream_t stream);
int32_t extract_host_4bit(kmc_ctx *ctx, const kmc_seqs *hs, int32_t k, int32_t mode, uint32_t flags,
                          const kmc_out *ho, kmc_result *res);

} // namespace kmc
