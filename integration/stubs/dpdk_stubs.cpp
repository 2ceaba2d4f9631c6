// Link-time stand-ins for the DPDK-backed functions the reference's hwacc benchmark refers to once DPDK_FOUND is defined
// (it has to be, or the benchmark times nothing: its clock is rte_rdtsc_precise). The ACC100 path they serve is never
// taken here (-T cuda); every factory answers "not available", and the cycle counter is the steady clock in nanoseconds.
#include "srsran/hal/dpdk/bbdev/bbdev_acc_factory.h"
#include "srsran/hal/dpdk/dpdk_eal_factory.h"
#include "srsran/hal/phy/upper/channel_processors/pusch/ext_harq_buffer_context_repository_factory.h"
#include <chrono>

extern "C" {
uint64_t rte_rdtsc_precise(void)
{
  return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}
uint64_t rte_get_tsc_hz(void)
{
  return 1000000000ULL;
}
void rte_mempool_free(struct rte_mempool*) {}
int  rte_eal_cleanup(void)
{
  return 0;
}
}

namespace srsran {
namespace dpdk {
std::unique_ptr<dpdk_eal> create_dpdk_eal(const std::string&, srslog::basic_logger&)
{
  return nullptr;
}
std::shared_ptr<bbdev_acc> create_bbdev_acc(const bbdev_acc_configuration&, srslog::basic_logger&)
{
  return nullptr;
}
} // namespace dpdk
namespace hal {
std::shared_ptr<ext_harq_buffer_context_repository> create_ext_harq_buffer_context_repository(unsigned, uint64_t, bool)
{
  return nullptr;
}
} // namespace hal
} // namespace srsran
