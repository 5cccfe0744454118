// Shared pieces of the CTA-pair group-convolution policies (conv_f16_policies.cuh).
#pragma once
#include "conv_policies.cuh"
#include "tc_gemm2.cuh"

namespace tvae {

// 3-D TMA load for the pair kernel (coordinates innermost first)
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

}  // namespace tvae
