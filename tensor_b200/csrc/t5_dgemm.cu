// FP64 tensor-core (DMMA) batched product — placeholder until the kernel lands.
#include "t5_internal.h"
namespace t5 {
int dmma_matmul(const Shard&, int, int, int, const void*, const void*, void*) { return T5I_NOT_SPECIALISED; }
}  // namespace t5
