/* integration/femto_b200_shim.c -- the reference-side binding of INTEGRATION.md section 2.
 *
 * Compiled INTO femto (against the reference's own headers) in place of the bodies of
 * parallel_count / parallel_locate / parallel_locate_range in src/main/femto.c:275-536.  With it
 * the reference's batch tools -- femto_multiquery (src/main/query_tool.c) and index_test
 * (src/main/index_test.c:351-434) -- link unmodified and run their count/locate batches on the
 * GPU through libfemto_b200.so.  Everything else of femto keeps its own implementation.
 *
 * oracle/Makefile builds _ref/femto_multiquery_b200 from the unmodified query_tool.c + this
 * file (tests/test_gpu_dropin.py runs it next to the stock femto_multiquery on the same index).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "femto_internal.h"   /* reference prototypes: femto_server_t, index_locator_t, error_t */
#include "server.h"           /* shared_server_state: path_to_id */
#include "block_storage.h"    /* path_translator_path_for_id */
#include "femto_b200.h"

#define FM_SHIM_MAX_INDEXES 64

static fm_index_t* handle_for(femto_server_t* srv, index_locator_t loc)
{
  /* one fm_index_t per index id, opened on first use; FEMTO_B200_DEVICE selects the GPU.  The
   * reference lets several caller threads run batches at once (src/main/server.c:3732-3793), so the
   * table is guarded; the fm_* calls themselves are thread-safe per handle. */
  static struct { intptr_t id; fm_index_t* ix; } cache[FM_SHIM_MAX_INDEXES];
  static pthread_mutex_t cache_mu = PTHREAD_MUTEX_INITIALIZER;
  path_translator_t* t = &srv->state->path_to_id;
  const char* path = NULL;
  fm_index_t* found = NULL;
  int i;
  pthread_mutex_lock(&cache_mu);
  for (i = 0; i < FM_SHIM_MAX_INDEXES; i++) if (cache[i].id == loc.id && cache[i].ix) { found = cache[i].ix; break; }
  if (!found) {
    /* id -> path: block_storage.h declares path_translator_path_for_id() but the reference never
     * defines it (only the static _unlocked helper, block_storage.c:237), so read the table here */
    pthread_rwlock_rdlock(&t->rwlock);
    if (loc.id > 0 && loc.id < (intptr_t) t->next_id) path = t->id_to_path[loc.id].path;
    pthread_rwlock_unlock(&t->rwlock);
    for (i = 0; path && i < FM_SHIM_MAX_INDEXES; i++) {
      if (!cache[i].ix) {
        const char* dev = getenv("FEMTO_B200_DEVICE");
        if (fm_open(path, dev ? atoi(dev) : 0, &cache[i].ix) == FM_OK) {
          cache[i].id = loc.id;
          found = cache[i].ix;
        } else {
          cache[i].ix = NULL;
        }
        break;
      }
    }
  }
  pthread_mutex_unlock(&cache_mu);
  return found;
}

static error_t to_error(int rc)
{
  return rc ? ERR_MAKE_STR((err_code_t) rc, fm_last_error()) : ERR_NOERR;
}

error_t parallel_count(femto_server_t* srv, index_locator_t loc, int npats, int* plen,
                       alpha_t** pats, int64_t* first, int64_t* last)
{
  fm_index_t* ix;
  if (!srv) return ERR_PARAM;
  ix = handle_for(srv, loc);
  return to_error(ix ? fm_count(ix, npats, plen, (const uint16_t* const*) pats, first, last) : FM_ERR_IO);
}

error_t parallel_locate(femto_server_t* srv, index_locator_t loc, int npats, int* plen,
                        alpha_t** pats, int max_occs_each, int* noccs, int64_t** offsets)
{
  fm_index_t* ix;
  if (!srv) return ERR_PARAM;
  ix = handle_for(srv, loc);
  return to_error(ix ? fm_locate(ix, npats, plen, (const uint16_t* const*) pats, max_occs_each, noccs, offsets)
                     : FM_ERR_IO);
}

error_t parallel_locate_range(femto_server_t* srv, index_locator_t loc, int64_t first, int64_t last,
                              int64_t* offsets)
{
  fm_index_t* ix;
  if (!srv) return ERR_PARAM;
  ix = handle_for(srv, loc);
  return to_error(ix ? fm_locate_range(ix, first, last, offsets) : FM_ERR_IO);
}
