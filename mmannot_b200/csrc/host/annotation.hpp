// GTF/GFF -> typed feature table in REFERENCE ORDER.
// Behaviour follows IntervalList::IntervalList / Gene / Transcript / GtfLineParser
// (mmannot.cpp:515-580, 708-829, 911-990, 1094-1290).  The output is the
// struct-of-arrays feature buffer handed to mma_load_features().
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "config.hpp"

namespace mmb {

struct FeatureTable {
  // one entry per typed interval, sorted by (chr, start) with the reference's tie order
  std::vector<uint32_t> chr, start, end;
  std::vector<uint8_t>  type;    // flattened Order element index
  std::vector<uint8_t>  strand;  // 1 = '+', 2 = anything else (mm:530)
  std::vector<std::string> id;   // interval id (gene id + suffix), for -M
  std::vector<std::string> chromosomes;  // annotation chromosome names, id = position
  std::vector<uint8_t> chrHasFeatures;   // reads on a chromosome without features are "unknown" (mm:1293)
  size_t nGenes = 0;
  size_t nLines = 0;
  size_t size() const { return start.size(); }
};

struct AnnotationOptions {
  uint64_t upstreamSize = 1000;    // -d (mm:80)
  uint64_t downstreamSize = 1000;  // -D (mm:81)
};

// Returns false with the reference's message in `err` on fatal problems; non-fatal
// "Warning, cannot deduce ..." lines are appended to `warnings`.
bool buildFeatureTable(const std::string &gtfFile, const Config &config, const AnnotationOptions &opt,
                       FeatureTable &out, std::string &err, std::string &warnings);

}  // namespace mmb
