// Feeds a BAM file to the device decoder (mma_submit_bam): the BGZF members are read as they lie in the file into page-locked
// chunks, the (small) BAM header is inflated and parsed here to resolve the reference names against the annotation, and
// everything else -- inflate, record parse, annotation -- happens on the GPU.  When the file holds something that route
// leaves to the host decoder (see mmannot_b200.h), run() says so and the caller falls back to XamReader for the whole file.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "annotation.hpp"
#include "mmannot_b200.h"
#include "xam.hpp"

namespace mmb {

class DeviceBamFeeder {
 public:
  enum class Result { DONE, FALLBACK, FAILED };
  DeviceBamFeeder(mma_ctx *ctx, const FeatureTable &features, Strandedness strandedness);
  ~DeviceBamFeeder();
  // DONE: every record of the file has been submitted to sample `column` (nRecords = their number, warnings = the reference's
  // "chromosome not present" lines in order of appearance).  FALLBACK: nothing usable was counted; reset the sample and decode
  // on the host (why = the reason, for the log at -p).  FAILED: err.
  Result run(const std::string &fileName, uint32_t column, uint64_t &nRecords, std::string &warnings, std::string &why, std::string &err);

 private:
  mma_ctx *ctx_;
  const FeatureTable &features_;
  Strandedness strandedness_;
  unsigned char *buf_[2] = {nullptr, nullptr};
  size_t cap_ = 0;
};

}  // namespace mmb
