/* The drop-in boundary must be consumable from plain C (cgo / JNI / ctypes generators read it as C99):
 *   gcc -std=c99 -pedantic -Wall -Werror -fsyntax-only -Iinclude tests/native/abi_c_check.c
 * and, linked against libcpmpc.so, every entry point the header declares must resolve. */
#include <stddef.h>
#include <stdio.h>

#include "cpmpc.h"

int main(void) {
  cpmpc_params p;
  cpmpc_plant_params q;
  cpmpc_handle h = NULL;
  cpmpc_stats s;
  cpmpc_default_params(&p);
  cpmpc_plant_default_params(CPMPC_PLANT_DOUBLE, &q);
  (void)s;
  (void)h;
  printf("ABI_C_OK version %d window %ld plant_nx %d flags %u %u\n", cpmpc_version(), (long)p.window_length,
         cpmpc_plant_state_dim(CPMPC_PLANT_DOUBLE), CPMPC_FLAG_LOW_LATENCY, CPMPC_FLAG_MIN_LATENCY);
  /* without a device the library must fail loudly, never fall back */
  if (cpmpc_create(&p, 4, 0, &h) != CPMPC_OK) {
    printf("create failed as expected without a GPU: %s\n", cpmpc_last_error());
  } else {
    cpmpc_destroy(h);
  }
  return 0;
}
