// Probe kernel for tests/studies/nvls/nvls_probe.py: every thread stores one 8-byte layout row through
// the MULTICAST mapping; NVSwitch replicates the store into the bound memory of every device of the multicast object.
#include <stdint.h>
extern "C" __global__ void k_multimem_store_rows(float2 *mc_rows, uint64_t first, uint64_t count, float tag)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float2 v = make_float2(tag, (float)(first + i));
    const uint64_t bits = ((uint64_t)__float_as_uint(v.y) << 32) | (uint64_t)__float_as_uint(v.x);
    asm volatile("multimem.st.weak.global.b64 [%0], %1;" ::"l"(mc_rows + first + i), "l"(bits) : "memory");
}
// the same rows written with plain stores into one peer replica (what k_epoch_in does today, once per peer)
extern "C" __global__ void k_plain_store_rows(float2 *rows, uint64_t first, uint64_t count, float tag)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) rows[first + i] = make_float2(tag, (float)(first + i));
}
