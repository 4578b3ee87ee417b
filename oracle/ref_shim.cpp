// oracle/ref_shim.cpp — thin C wrapper that is compiled TOGETHER WITH the reference's own
// src/Grouping.cpp and src/Reference.cpp (taken where they lie under /root/reference, never
// copied) into oracle/_ref/libmsweep_ref_grouping.so.  Those two translation units are the only
// part of the reference that builds without its eight un-vendored dependencies.
// TEST INFRASTRUCTURE ONLY: used to pin oracle::read_grouping against the real reference.
#include "Reference.hpp"

#include <cstring>
#include <fstream>
#include <limits>
#include <string>

namespace {
thread_local std::string g_err;
struct Handle { std::unique_ptr<mSWEEP::Reference> ref; };

// The reference picks the stored integer width from the counts (src/Reference.cpp:45-54,
// src/Grouping.cpp:46-87) and its callers static_cast accordingly (src/mSWEEP.cpp:336-344).
template <typename T> void indicators_as(const mSWEEP::Reference &r, uint32_t *out) {
  const auto &v = static_cast<const mSWEEP::AdaptiveReference<T>&>(r).get_group_indicators(0);
  for (size_t i = 0; i < v.size(); ++i) out[i] = (uint32_t)v[i];
}
template <typename U, typename T> void sizes_as(const mSWEEP::Reference &r, uint64_t *out) {
  const auto &v = static_cast<const mSWEEP::AdaptiveReference<T>&>(r).template group_sizes<U>(0);
  for (size_t i = 0; i < v.size(); ++i) out[i] = (uint64_t)v[i];
}
template <typename T> void sizes_dispatch(const mSWEEP::Reference &r, uint64_t *out) {
  const size_t m = r.get_grouping(0).max_group_size();
  if (m <= std::numeric_limits<uint8_t>::max()) sizes_as<uint8_t, T>(r, out);
  else if (m <= std::numeric_limits<uint16_t>::max()) sizes_as<uint16_t, T>(r, out);
  else if (m <= std::numeric_limits<uint32_t>::max()) sizes_as<uint32_t, T>(r, out);
  else sizes_as<uint64_t, T>(r, out);
}
} // namespace

extern "C" {
const char *ref_last_error() { return g_err.c_str(); }
int ref_grouping_read(const char *path, void **out) {
  try {
    std::ifstream in(path);
    Handle *h = new Handle;
    h->ref = mSWEEP::ConstructAdaptiveReference(&in, '\t');
    *out = h;
    return 0;
  } catch (const std::exception &e) { g_err = e.what(); return 1; }
}
uint64_t ref_grouping_n_targets(const void *h) { return ((const Handle*)h)->ref->get_n_refs(); }
uint32_t ref_grouping_n_groups(const void *h) { return (uint32_t)((const Handle*)h)->ref->n_groups(0); }
const char *ref_grouping_name(const void *h, uint32_t g) { return ((const Handle*)h)->ref->group_names(0)[g].c_str(); }
void ref_grouping_export(const void *h, uint32_t *group_of_target, uint64_t *sizes) {
  const mSWEEP::Reference &r = *((const Handle*)h)->ref;
  const size_t K = r.n_groups(0);
  if (K <= std::numeric_limits<uint8_t>::max()) { indicators_as<uint8_t>(r, group_of_target); sizes_dispatch<uint8_t>(r, sizes); }
  else if (K <= std::numeric_limits<uint16_t>::max()) { indicators_as<uint16_t>(r, group_of_target); sizes_dispatch<uint16_t>(r, sizes); }
  else if (K <= std::numeric_limits<uint32_t>::max()) { indicators_as<uint32_t>(r, group_of_target); sizes_dispatch<uint32_t>(r, sizes); }
  else { indicators_as<uint64_t>(r, group_of_target); sizes_dispatch<uint64_t>(r, sizes); }
}
void ref_grouping_free(void *h) { delete (Handle*)h; }
}
