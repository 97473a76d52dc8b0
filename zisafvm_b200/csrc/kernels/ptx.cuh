// PTX wrappers used by the streaming kernels: mbarrier, TMA bulk copies (cp.async.bulk, UBLKCP in SASS), L2 policies.
#pragma once
#include <cstdint>

#include "common.cuh"

namespace zfvm {

namespace ptx {
ZFVM_DEVICE std::uint32_t smem_u32(const void *p) { return (std::uint32_t)__cvta_generic_to_shared(p); }
ZFVM_DEVICE void mbar_init(std::uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
ZFVM_DEVICE void mbar_expect_tx(std::uint64_t *bar, std::uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
ZFVM_DEVICE void mbar_arrive(std::uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
ZFVM_DEVICE void mbar_wait(std::uint64_t *bar, int parity) {
  const std::uint32_t a = smem_u32(bar);
  std::uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"((std::uint32_t)parity)
        : "memory");
  } while (!done);
}
ZFVM_DEVICE std::uint64_t policy_evict_first() {
  std::uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
ZFVM_DEVICE std::uint64_t policy_evict_normal() {
  std::uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
/// TMA bulk copy global -> shared, completion counted in bytes on `bar`.
ZFVM_DEVICE void bulk_g2s(void *dst, const void *src, std::uint32_t bytes, std::uint64_t *bar, std::uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
ZFVM_DEVICE void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
ZFVM_DEVICE void named_bar_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
}  // namespace ptx

}  // namespace zfvm
