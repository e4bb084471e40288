// fm_sa.cc -- host suffix sort (prefix doubling) for tests and small corpora.
//
// Order: plain lexicographic order of the suffixes of the prepared text, a suffix that is a
// proper prefix of another sorting first -- the order the reference's builders produce
// (in-memory: src/main/bwt_qsufsort.c:331-352 over src/utils/suffix_sort.c; external:
// src/dcx_cc).  Large corpora are sorted on the GPU by femto_b200/build_gpu.py instead.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "../../include/femto_b200.h"

extern "C" int fm_suffix_sort_host(const uint16_t* text, int64_t n, int64_t* sa) {
  if (n < 0 || (n && (!text || !sa))) return FM_ERR_PARAM;
  if (n == 0) return FM_OK;
  try {
    std::vector<int64_t> rank(static_cast<size_t>(n)), tmp(static_cast<size_t>(n));
    std::iota(sa, sa + n, int64_t(0));
    for (int64_t i = 0; i < n; i++) rank[size_t(i)] = text[i];
    for (int64_t k = 1;; k <<= 1) {
      auto key2 = [&](int64_t i) { return i + k < n ? rank[size_t(i + k)] : int64_t(-1); };
      auto less = [&](int64_t a, int64_t b) {
        if (rank[size_t(a)] != rank[size_t(b)]) return rank[size_t(a)] < rank[size_t(b)];
        return key2(a) < key2(b);
      };
      std::sort(sa, sa + n, less);
      tmp[size_t(sa[0])] = 0;
      for (int64_t i = 1; i < n; i++) tmp[size_t(sa[i])] = tmp[size_t(sa[i - 1])] + (less(sa[i - 1], sa[i]) ? 1 : 0);
      rank.swap(tmp);
      if (rank[size_t(sa[n - 1])] == n - 1) break;
      if (k > n) break;
    }
  } catch (const std::bad_alloc&) {
    return FM_ERR_MEM;
  }
  return FM_OK;
}
