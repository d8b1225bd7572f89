// -m (per-read) and -M (per-interval) statistics files.
//
// The element set of every hit and the intervals behind it are computed on the GPU (mma_annotate_hits /
// mma_annotate_intervals); what is left here is text: grouping the hits of a read NAME, which only the host knows, and
// formatting.  The bookkeeping follows what Counter::addCount does with its name-keyed maps (mmannot.cpp:1665-1739,
// end-of-file flush 1783-1800), the line formats are printReadStats (mmannot.cpp:474-493) and Counter::dump
// (mmannot.cpp:1819-1850).  The order of the -m lines written at end of file is the iteration order of the reference's
// std::unordered_map<std::string, ...> (mmannot.cpp:1656): the same container, fed the same insert/erase sequence, is used
// here, so the order is the same with the same libstdc++.
#pragma once
#include <cstdint>
#include <map>
#include <ostream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "annotation.hpp"
#include "config.hpp"

namespace mmb {

class StatsWriters {
 public:
  StatsWriters(const Config &config, const FeatureTable &features, int strategy, float rescueThreshold, std::ostream *readStats,
               bool intervalStats);
  // one visited hit, in file order (under -y unique the caller only passes records with NH == 1, mm:1773)
  void addHit(const std::string &name, uint32_t nHits, uint64_t elements, const uint32_t *intervals, size_t nIntervals);
  void endOfFile();                              // mm:1783-1800
  void dumpIntervals(std::ostream &out) const;   // mm:1819-1850

 private:
  struct Open {
    uint32_t remaining = 0, rawNh = 0;
    std::vector<uint32_t> elements;   // with repeats, in arrival order
  };
  void printRead(const std::string &name, uint32_t nHits, std::vector<uint32_t> &elements);
  void countIntervals(std::vector<uint32_t> &intervals);

  const Config &config_;
  const FeatureTable &features_;
  int strategy_;
  float rescueThreshold_;
  std::ostream *readStats_;
  bool intervalStats_;
  std::unordered_map<std::string, Open> open_;                           // readCounts + rawCounts
  std::unordered_map<std::string, std::vector<uint32_t>> openIntervals_; // readsIntervals
  std::map<std::vector<uint32_t>, unsigned int> intervalCounts_;
  std::unordered_map<std::string, uint32_t> chosenId_, numberSeen_;      // -y random
  std::unordered_set<std::string> seen_;
};

}  // namespace mmb
