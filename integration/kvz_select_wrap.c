/*
 * kvz_select_wrap.c -- registers the "cuda" strategies THROUGH the reference's own selector.
 *
 * INTEGRATION.md 2 shows the one line per group a maintainer adds to src/strategies/strategies-<group>.c.  The reference
 * tree may not be edited here, so the same effect is produced at link time: the group registrars that
 * kvz_strategyselector_init calls (src/strategyselector.c:66-104) are wrapped (-Wl,--wrap=kvz_strategy_register_<group>,
 * oracle/Makefile target `sel`), each wrapper runs the reference's registrar and then the cuda registrar of the group.
 * Every cuda entry therefore goes through kvz_strategyselector_register (strategyselector.c:233-273) and is chosen
 * -- or not -- by strategyselector_choose_for (:275-341): priority 50 beats AVX2's 40, KVAZAAR_OVERRIDE_<type> forces
 * any registered name.  tests/test_abi_and_dropin.py::test_selection_through_the_reference_selector.
 */
#include <stdint.h>

#define WRAP(group, cuda_fn)                                                     \
  int __real_kvz_strategy_register_##group(void *opaque, uint8_t bitdepth);      \
  int cuda_fn(void *opaque, uint8_t bitdepth);                                   \
  int __wrap_kvz_strategy_register_##group(void *opaque, uint8_t bitdepth)       \
  {                                                                              \
    return __real_kvz_strategy_register_##group(opaque, bitdepth) & cuda_fn(opaque, bitdepth); \
  }

WRAP(picture, kvz_strategy_register_picture_all_cuda)
WRAP(nal, kvz_strategy_register_nal_cuda)
WRAP(dct, kvz_strategy_register_dct_cuda)
WRAP(ipol, kvz_strategy_register_ipol_cuda)
WRAP(quant, kvz_strategy_register_quant_cuda)
WRAP(intra, kvz_strategy_register_intra_cuda)
WRAP(sao, kvz_strategy_register_sao_cuda)
