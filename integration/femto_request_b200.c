/* integration/femto_request_b200.c -- the reference's femto_handle_request tool
 * (src/main/handle_request.c: "Usage: <index> <request>") on the B200 engine: the same three
 * sections on stdout, the response produced by fm_generic_request instead of a femto server.
 * Serves the string_rows* requests (femto.h:75-137).
 *   cc -Iinclude integration/femto_request_b200.c -Lfemto_b200 -lfemto_b200 -o femto_request_b200
 */
#include <stdio.h>
#include <stdlib.h>

#include "femto_b200.h"

int main(int argc, char** argv)
{
  fm_index_t* ix = NULL;
  char* response = NULL;
  const char* dev = getenv("FEMTO_B200_DEVICE");
  int rc;
  if (argc != 3) {
    printf("Usage: %s <index> <request>\n", argv[0]);
    return -1;
  }
  rc = fm_open(argv[1], dev ? atoi(dev) : 0, &ix);
  if (rc) {
    fprintf(stderr, "%s\n", fm_last_error());
    return rc;
  }
  printf("Index:%s\n", argv[1]);
  printf("Request:\n%s\n", argv[2]);
  rc = fm_generic_request(ix, argv[2], &response);
  if (rc) {
    fprintf(stderr, "%s\n", fm_last_error());
    fm_close(ix);
    return rc;
  }
  printf("Response:\n%s\n", response);
  free(response);
  fm_close(ix);
  return 0;
}
