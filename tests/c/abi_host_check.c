/* Plain-C consumer of include/ofspmm.h: proves the header is C-clean (no C++/CUDA/torch types)
 * and that the host-side entry points link and behave from C.  No GPU needed. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ofspmm.h"

int main(void) {
  /* 4 rows: lengths 2, 0, 3, 1 */
  int32_t crow[5] = {0, 2, 2, 5, 6};
  int64_t rows_out[4], nz_out[4];
  int rc = ofspmm_partition_host(crow, OFSPMM_DTYPE_INT32, 4, 6, 3, rows_out, nz_out);
  if (rc != OFSPMM_OK) { printf("partition_host rc=%d\n", rc); return 1; }
  /* total = 10 items, 3 parts -> 4 items each: diagonals 0,4,8,10 */
  const int64_t want_sum[4] = {0, 4, 8, 10};
  for (int k = 0; k < 4; ++k)
    if (rows_out[k] + nz_out[k] != want_sum[k]) { printf("diag %d: %lld+%lld\n", k, (long long)rows_out[k], (long long)nz_out[k]); return 2; }
  if (rows_out[3] != 4 || nz_out[3] != 6 || rows_out[0] != 0) return 3;
  if (ofspmm_partition_host(NULL, OFSPMM_DTYPE_INT32, 4, 6, 3, rows_out, nz_out) != OFSPMM_ERR_INVALID_ARG) return 4;
  if (ofspmm_partition_host(crow, OFSPMM_DTYPE_FLOAT, 4, 6, 3, rows_out, nz_out) != OFSPMM_ERR_UNSUPPORTED_DTYPE) return 5;
  if (ofspmm_fwd_workspace_bytes(4, 4, 6, 128, OFSPMM_DTYPE_FLOAT) == 0) return 6;
  if (ofspmm_bwd_b_workspace_bytes(4, 4, 6, 128, OFSPMM_DTYPE_BFLOAT16, 0) <=
      ofspmm_bwd_b_workspace_bytes(4, 4, 6, 128, OFSPMM_DTYPE_FLOAT, 0)) return 7;   /* bf16 needs the fp32 accumulator */
  ofspmm_csr a;
  memset(&a, 0, sizeof a);
  a.rows = 4; a.cols = 4; a.nnz = 6; a.crow = crow; a.idx_dtype = OFSPMM_DTYPE_INT32; a.val_dtype = OFSPMM_DTYPE_FLOAT;
  /* col / val missing with nnz > 0 -> invalid argument, before any CUDA call */
  if (ofspmm_fwd(&a, NULL, NULL, 8, OFSPMM_DTYPE_FLOAT, NULL, 0, NULL) != OFSPMM_ERR_INVALID_ARG) return 8;
  if (strcmp(ofspmm_strerror(OFSPMM_OK), "ok") != 0) return 9;
  if (ofspmm_version() != 100) return 10;
  printf("abi_host_check ok: version %d, launches so far %llu\n", ofspmm_version(),
         (unsigned long long)ofspmm_launch_count());
  return 0;
}
