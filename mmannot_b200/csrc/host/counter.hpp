// Host mirror of the reference's Counter / TableCount (mmannot.cpp:1653-1901) on top of the
// device C ABI.  Same method names and meaning: Counter::read(file) annotates one input,
// dump() prints its statistics, TableCount::addCounter()/dump() build and print the table.
// The per-hit and per-read work happens on the GPU (mma_submit_hits / mma_finish_sample).
#pragma once
#include <cmath>
#include <map>
#include <ostream>
#include <string>
#include <vector>

#include "annotation.hpp"
#include "config.hpp"
#include "mmannot_b200.h"
#include "stats_writers.hpp"
#include "xam.hpp"

namespace mmb {

struct RunOptions {
  Strandedness strandedness = Strandedness::F;  // mm:1938
  ReadsFormat format = ReadsFormat::UNKNOWN;
  int strategy = MMA_STRATEGY_DEFAULT;
  float overlap = -1.0f;          // mm:1932
  float rescueThreshold = 1.0f;   // mm:1933
  bool readStats = false, intervalStats = false, progress = false;
  uint32_t batchHits = 1u << 22;
};

std::string withThousands(uint64_t n);  // the comma_numpunct locale of mm:111-115, 2092-2093

class Counter {
 public:
  Counter(mma_ctx *ctx, const FeatureTable &features, const Config &config, const RunOptions &opt);
  // One input spread over several GPUs (one context each, all created with the same parameters): the hits of a batch are dealt
  // out by read name (hash of the read key mod #GPUs, so that every record of a name -- both mates, every repeat -- stays on
  // one GPU in file order), and the tables are summed on the devices at the end of the file (mma_allreduce).  Not with -m / -M.
  Counter(const std::vector<mma_ctx *> &ctxs, const FeatureTable &features, const Config &config, const RunOptions &opt);
  ~Counter();
  // Annotates one SAM/BAM file as sample `column`; false + message on a fatal problem.
  bool read(const std::string &fileName, uint32_t column, std::string &err, std::ostream &log);
  void dump(std::ostream &log) const;  // mm:1806-1818
  // element set (bitmask) -> the reference's regionCounts value
  const std::map<uint64_t, double> &getCounts() const { return counts_; }
  const mma_sample_stats &getStats() const { return stats_; }
  // -m / -M: the per-hit element sets (and interval ids) of every batch are also handed to `w` in file order
  void setStatsWriters(StatsWriters *w) { writers_ = w; }

 private:
  mma_ctx *ctx_;
  const FeatureTable &features_;
  const Config &config_;
  RunOptions opt_;
  std::string fileName_;
  mma_sample_stats stats_;
  std::map<uint64_t, double> counts_;
  HitBuffers pinned_[2];
  // compact transfer format of each slot (mma_pack_hits): what actually crosses PCIe
  struct PackedBuffers {
    uint32_t *packed = nullptr, *tileRunBase = nullptr, *escIndex = nullptr, *escEnd = nullptr, *escNh = nullptr;
    uint64_t *runKey = nullptr;
    uint64_t escCapacity = 0;
  } packedBuf_[2];
  StatsWriters *writers_ = nullptr;
  // several GPUs: shards_[g] holds GPU g's two page-locked slots (pinned_ / packedBuf_ above are those of GPU 0's reader batch)
  struct Shard {
    mma_ctx *ctx = nullptr;
    HitBuffers pinned[2];
    PackedBuffers packed[2];
    size_t n = 0;
  };
  std::vector<Shard> shards_;
  bool readSharded(XamReader &reader, uint32_t column, std::string &err, std::ostream &log);
};

class TableCount {
 public:
  TableCount(const Config &config, uint32_t nInputs) : config_(config), nInputs_(nInputs), nColumns_(0) {}
  void addCounter(const Counter &counter) { addCounts(counter.getCounts()); }   // mm:1861-1876
  void addCounts(const std::map<uint64_t, double> &counts);
  void dump(std::ostream &out, const std::vector<std::string> &samples) const;  // mm:1877-1900

 private:
  const Config &config_;
  uint32_t nInputs_, nColumns_;
  std::map<std::vector<size_t>, std::vector<unsigned int> > rows_;  // ordered like the reference's sorted lineNames
};

}  // namespace mmb
