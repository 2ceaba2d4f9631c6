/* Build-only stand-ins for the DPDK headers the reference's hal layer includes unconditionally (DPDK is not installed in
 * this image and the ACC100 accelerator is never built here: ENABLE_PUSCH_HWACC stays undefined). They declare just the
 * names the hal HEADERS mention, so that hw_accelerator_factories.cpp with the "cuda" branch can be compiled; nothing in
 * them is ever called. */
#ifndef PDC_INTEGRATION_DPDK_STUB_H
#define PDC_INTEGRATION_DPDK_STUB_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
struct rte_mempool;
struct rte_mbuf;
struct rte_bbdev_driver_info {
  uint32_t harq_buffer_size;
};
struct rte_bbdev_info {
  int                          socket_id;
  struct rte_bbdev_driver_info drv;
};
enum rte_bbdev_op_type { RTE_BBDEV_OP_NONE, RTE_BBDEV_OP_LDPC_DEC, RTE_BBDEV_OP_LDPC_ENC, RTE_BBDEV_OP_FFT };
struct rte_bbdev_dec_op {
  int stub;
};
struct rte_bbdev_enc_op {
  int stub;
};
#define RTE_BBDEV_LDPC_E_MAX_MBUF 64000
void     rte_mempool_free(struct rte_mempool* mp);
int      rte_eal_cleanup(void);
uint64_t rte_rdtsc_precise(void);
uint64_t rte_get_tsc_hz(void);
#ifdef __cplusplus
}
#endif
#endif
