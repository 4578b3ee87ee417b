// input.hpp — host-side readers: the -i group indicator file and Themisto plaintext pseudoalignments.
#pragma once
#include "msweep_b200.hpp"

#include <string>
#include <vector>

namespace b200 {

// include/Grouping.hpp:62-83 + include/Reference.hpp:67-94: one line per reference sequence, optional
// tab-separated further groupings (only `column` is read); ids in order of first appearance.
struct Grouping {
  std::vector<std::string> names;
  std::vector<uint64_t> sizes;
  std::vector<uint32_t> group_of_target;
  size_t n_groupings = 1;
};
Grouping read_grouping(const std::string &path, char delimiter = '\t', size_t column = 0);

// include/mSWEEP_alignment.hpp:54-135: "<read_id> <t0> <t1> ...", paired strands merged by
// "intersection" or "union".  Multi-threaded tokeniser over the whole file in memory.
ReadTable read_themisto(const std::vector<std::string> &paths, uint64_t n_targets, const std::string &merge_mode,
                        int n_threads);

} // namespace b200
