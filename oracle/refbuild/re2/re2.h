/* oracle/refbuild/re2/re2.h -- TEST INFRASTRUCTURE.  Stand-in for the RE2 header that the reference's
 * femto_search (src/main_cc/search_tool.cc) includes: RE2 is only behind its --filter-results option,
 * which the checker never passes, so the vendored src/re2 tree need not be compiled. */
#pragma once
#include <string>
namespace re2 {
struct StringPiece {
  StringPiece(const char*, long) {}
};
}  // namespace re2
class RE2 {
 public:
  explicit RE2(const char*) {}
  bool ok() const { return false; }
  std::string error() const { return "RE2 is not built into this checker"; }
  static bool PartialMatch(const re2::StringPiece&, const RE2&) { return false; }
};
