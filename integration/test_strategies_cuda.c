/*
 * test_strategies_cuda.c -- drop-in replacement for the reference's tests/test_strategies.c (ref: :41-65): the
 * same global `strategies` list the greatest suites iterate over, with the cuda registrars appended after the
 * host's own, so the UNMODIFIED suites (sad_tests, intra_sad_tests, satd_tests, dct_tests, coeff_sum_tests,
 * speed_tests) exercise every "cuda" entry exactly like they exercise "generic"/"avx2".
 */
#include <stdio.h>

#include "src/strategyselector.h"
#include "tests/test_strategies.h"

int kvz_strategy_register_picture_all_cuda(void *opaque, uint8_t bitdepth);
int kvz_strategy_register_dct_cuda(void *opaque, uint8_t bitdepth);
int kvz_strategy_register_quant_cuda(void *opaque, uint8_t bitdepth);

strategy_list_t strategies;

void init_test_strategies()
{
  strategies.allocated = 0;
  strategies.count = 0;
  strategies.strategies = NULL;
  kvz_strategyselector_init(1, KVZ_BIT_DEPTH, 1);
  if (!kvz_strategy_register_picture(&strategies, KVZ_BIT_DEPTH) || !kvz_strategy_register_dct(&strategies, KVZ_BIT_DEPTH) ||
      !kvz_strategy_register_quant(&strategies, KVZ_BIT_DEPTH)) {
    fprintf(stderr, "reference strategy registration failed\n");
    return;
  }
  const unsigned before = strategies.count;
  if (!kvz_strategy_register_picture_all_cuda(&strategies, KVZ_BIT_DEPTH) || !kvz_strategy_register_dct_cuda(&strategies, KVZ_BIT_DEPTH) ||
      !kvz_strategy_register_quant_cuda(&strategies, KVZ_BIT_DEPTH)) {
    fprintf(stderr, "cuda strategy registration failed\n");
    return;
  }
  fprintf(stderr, "test_strategies_cuda: %u cuda entries appended to %u reference entries\n", strategies.count - before, before);
}
