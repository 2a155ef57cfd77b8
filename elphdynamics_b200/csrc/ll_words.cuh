// Self-validating words for flag-free exchange through global / peer memory (the "LL" idea of NCCL): a double travels
// as ONE 16-byte store of two 64-bit words, each carrying half of the value and a 32-bit tag naming the step that
// produced it.  The reader spins on the element itself until both tags match: no flag, no fence, and a torn store
// cannot pass for a complete one.  Shared by cg_p2p.cu and cg_pipe.cu.
#pragma once

#include <cuda_runtime.h>

namespace ll {

__device__ __forceinline__ void st2(unsigned long long* p, unsigned long long w0, unsigned long long w1) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ void ld2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void push(unsigned long long* p, double v, unsigned int tag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    st2(p, (bits & 0xffffffffull) | ((unsigned long long)tag << 32), (bits >> 32) | ((unsigned long long)tag << 32));
}
__device__ __forceinline__ double unpack(unsigned long long a, unsigned long long b) {
    return __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
}
__device__ __forceinline__ bool tag_ok(unsigned long long a, unsigned long long b, unsigned int tag) {
    return ((unsigned int)(a >> 32) == tag) && ((unsigned int)(b >> 32) == tag);
}

}  // namespace ll
