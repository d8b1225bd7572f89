#include "stats_writers.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "mmannot_b200.h"

namespace mmb {

StatsWriters::StatsWriters(const Config &config, const FeatureTable &features, int strategy, float rescueThreshold,
                           std::ostream *readStats, bool intervalStats)
    : config_(config), features_(features), strategy_(strategy), rescueThreshold_(rescueThreshold), readStats_(readStats),
      intervalStats_(intervalStats) {}

// One -m line: "name \tNH" then "\telement: multiplicity" per distinct element in ascending order, then "\tRescued"
// when one element holds at least ceil(n * threshold) of the n entries (mm:474-509; only with -e < 100).
void StatsWriters::printRead(const std::string &name, uint32_t nHits, std::vector<uint32_t> &elements) {
  if (!readStats_) return;
  std::sort(elements.begin(), elements.end());
  std::ostream &o = *readStats_;
  o << name << " \t" << nHits;
  for (size_t i = 0; i < elements.size();) {
    size_t j = i;
    while (j < elements.size() && elements[j] == elements[i]) ++j;
    o << "\t" << config_.getName(elements[i]) << ": " << (j - i);
    i = j;
  }
  bool rescued = false;
  if (rescueThreshold_ < 1.0f && elements.size() != 1) {
    const size_t t = static_cast<size_t>(std::ceil(static_cast<float>(elements.size()) * rescueThreshold_));
    for (size_t i = 0; i < elements.size() && !rescued;) {
      size_t j = i;
      while (j < elements.size() && elements[j] == elements[i]) ++j;
      if (j - i >= t) rescued = true;  // the smallest element reaching the threshold wins (the list is sorted)
      i = j;
    }
  }
  if (rescued) o << "\tRescued";
  o << "\n";
}

void StatsWriters::countIntervals(std::vector<uint32_t> &intervals) {
  if (intervals.empty()) return;
  std::sort(intervals.begin(), intervals.end());
  ++intervalCounts_[intervals];
}

void StatsWriters::addHit(const std::string &name, uint32_t nHits, uint64_t elementMask, const uint32_t *intervals, size_t nIntervals) {
  std::vector<uint32_t> elements;
  for (uint32_t e = 0; e < 64; ++e)
    if ((elementMask >> e) & 1ull) elements.push_back(e);
  if (nHits > 1 && strategy_ == MMA_STRATEGY_DEFAULT) {  // the hits of a multi-mapping read are gathered by name
    auto pos = open_.find(name);
    if (pos == open_.end()) {
      Open &g = open_[name];
      g.remaining = nHits - 1;
      g.rawNh = nHits;
      g.elements = elements;
      if (intervalStats_) openIntervals_[name].assign(intervals, intervals + nIntervals);
      return;
    }
    Open &g = pos->second;
    --g.remaining;
    g.elements.insert(g.elements.end(), elements.begin(), elements.end());
    if (intervalStats_) {
      std::vector<uint32_t> &v = openIntervals_[name];
      v.insert(v.end(), intervals, intervals + nIntervals);
    }
    if (g.remaining == 0) {
      if (!g.elements.empty()) {
        printRead(name, nHits, g.elements);
        if (intervalStats_) {
          auto pri = openIntervals_.find(name);
          countIntervals(pri->second);
          openIntervals_.erase(pri);
        }
      }
      open_.erase(pos);
    }
    return;
  }
  if (elements.empty()) return;
  if (strategy_ == MMA_STRATEGY_RANDOM) {  // the i-th annotated hit of the name, i drawn once per name (mm:1706-1726)
    if (seen_.count(name)) return;
    auto p = chosenId_.find(name);
    uint32_t i;
    if (p == chosenId_.end()) {
      i = static_cast<uint32_t>(rand()) % nHits;
      chosenId_[name] = i;
      numberSeen_[name] = 0;
    } else {
      i = p->second;
      ++numberSeen_[name];
    }
    if (numberSeen_[name] != i) return;
    chosenId_.erase(name);
    numberSeen_.erase(name);
    seen_.insert(name);
  }
  printRead(name, nHits, elements);
  if (intervalStats_ && nIntervals) {
    std::vector<uint32_t> v(intervals, intervals + nIntervals);
    countIntervals(v);
  }
}

void StatsWriters::endOfFile() {
  for (auto &e : open_)
    if (!e.second.elements.empty()) printRead(e.first, e.second.rawNh, e.second.elements);
  open_.clear();
  if (intervalStats_)
    for (auto &e : openIntervals_) countIntervals(e.second);
  openIntervals_.clear();
}

// One line per distinct multiset of intervals: "id (element)" strings sorted and joined by " -- ", lines sorted,
// equal lines summed (mm:1819-1850).
void StatsWriters::dumpIntervals(std::ostream &out) const {
  std::vector<std::pair<std::string, unsigned int>> lines;
  for (const auto &p : intervalCounts_) {
    std::vector<std::string> names;
    for (uint32_t i : p.first) names.push_back(features_.id[i] + " (" + config_.getName(features_.type[i]) + ")");
    std::sort(names.begin(), names.end());
    std::string line;
    for (size_t k = 0; k < names.size(); ++k) { if (k) line += " -- "; line += names[k]; }
    lines.emplace_back(line, p.second);
  }
  std::sort(lines.begin(), lines.end());
  std::string current;
  unsigned int count = 0;
  for (const auto &l : lines) {
    if (l.first == current) count += l.second;
    else {
      if (!current.empty()) out << current << "\t" << count << "\n";
      current = l.first;
      count = l.second;
    }
  }
  if (!current.empty()) out << current << "\t" << count << "\n";
}

}  // namespace mmb
