// Synthetic annotation + alignment generator for the benchmark shapes named in BASELINE.json
// (TAIR10-, GRCh38/Ensembl- and FlyBase6-shaped annotations; sRNA-Seq / RNA-Seq multi-mapping
// reads).  Deterministic: every read is drawn from a counter-based generator keyed by
// (seed, read index), so any range of reads can be produced independently and in parallel, and
// the BAM written for the reference holds exactly the hits of the packed buffers fed to the GPU.
// Only the element names come from the reference's config files; genome shapes are ours.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "annotation.hpp"
#include "xam.hpp"

namespace mmb {

struct SynthGene {
  uint32_t chr;
  uint64_t start, end;
  bool forward;
  uint32_t cls;  // index into the shape's class table
  std::vector<std::pair<uint64_t, uint64_t> > exons;  // of the union transcript
  uint64_t cdsStart = 0, cdsEnd = 0;                  // 0 = non coding
  uint32_t nTranscripts = 1;
  uint64_t serial;
};

struct SynthReadSpec {
  uint32_t maxNH = 20;
  bool paired = false;      // two records per placement sharing the read name (mate flags set)
  bool flipMate2 = false;   // store mate 2 with its strand bit flipped (turns -s FR into -s F for the reference)
  bool rnaSeq = false;      // read length 50..150 instead of 18..30
  double pInFeature = 0.5;  // a placement falls inside a random gene with this probability
  double pSameClass = 0.6;  // all hits of a multi-mapping read fall into genes of one class
};

struct SynthRecord {
  uint32_t chr;       // index into SynthGenome::chrNames
  uint64_t pos;       // 1-based
  uint32_t len;
  bool forward;
  uint32_t nh;
  uint32_t flag;      // SAM flag bits other than 0x10
};

class SynthGenome {
 public:
  // shape: "tair10" | "hs38" | "flybase6".  geneScale scales the number of genes (1.0 = full size).
  SynthGenome(const std::string &shape, uint64_t seed, double geneScale);
  bool ok() const { return ok_; }
  void writeAnnotation(const std::string &path) const;

  // records of read `r` (file order); returns the read name in `name`
  void readRecords(uint64_t r, const SynthReadSpec &spec, std::string &name, std::vector<SynthRecord> &out) const;

  // reads [first, first + nReads): BAM for the reference ...
  // (headerless: records only -- BGZF files can be concatenated, so that a large BAM can be written as parts side by side)
  bool writeBam(const std::string &path, uint64_t first, uint64_t nReads, const SynthReadSpec &spec, bool coordinateSorted, bool headerless = false,
                bool straddle = false) const;  // straddle: records may cross BGZF members (htslib never writes that)
  // ... number of hits of the range, and the same hits as packed buffers (chromosome ids of `features`)
  uint64_t countHits(uint64_t first, uint64_t nReads, const SynthReadSpec &spec) const;
  uint64_t fillHits(const FeatureTable &features, Strandedness s, uint64_t first, uint64_t nReads, const SynthReadSpec &spec,
                    const HitBuffers &dst) const;

  std::vector<std::string> chrNames;
  std::vector<uint64_t> chrLen;
  std::vector<SynthGene> genes;

 private:
  struct GeneClass {
    std::string biotype;   // column 2 for Ensembl style, column 3 of the child line for GFF3 style
    double weight;
    bool coding, spliced;
    uint32_t minLen, maxLen;
    bool inConfig;
  };
  std::string shape_;
  uint64_t seed_;
  bool ok_ = false, gff3_ = false;
  std::string source_;
  std::vector<GeneClass> classes_;
  std::vector<std::vector<uint32_t> > genesOfClass_;
  std::vector<uint64_t> chrCum_;
  uint64_t genomeLen_ = 0;
};

}  // namespace mmb
