// SAM / BAM decode -> packed struct-of-arrays hit batches (the buffers of mma_submit_hits()).
// Hit semantics follow Reader / SamReader / BamReader / Read of the reference
// (mmannot.cpp:846-903, 1339-1650) with XamRecord::setFlags repaired (mm:606: the FLAG
// is honoured).  One hit = one record or one XA alternative (one iteration of mm:1772).
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "annotation.hpp"
#include "bgzf.hpp"

namespace mmb {

enum class ReadsFormat { UNKNOWN, SAM, BAM };
enum class Strandedness { U, F, R };

static const uint32_t HIT_CHR_MASK = 0x00FFFFFFu;  // meta bits 0..23: annotation chromosome id
static const uint32_t HIT_CHR_NONE = 0x00FFFFFFu;  // chromosome without features / unmapped
static const uint32_t HIT_STRAND_BIT = 0x80000000u;  // meta bit 31: read strand after -s mapping (mm:836-844)

// Caller-owned struct-of-arrays destination (typically pinned memory).
struct HitBuffers {
  uint32_t *start = nullptr, *end = nullptr, *meta = nullptr, *nh = nullptr;
  uint64_t *key = nullptr;
  size_t capacity = 0;
};

class XamReader {
 public:
  XamReader(const std::string &fileName, ReadsFormat format, Strandedness strandedness, const FeatureTable &features);
  ~XamReader();

  // false + message when the file cannot be opened / is not what it claims to be
  bool open(std::string &err);
  // the checks of open() that need no decoding: the file exists and its format is known (isBam() is valid afterwards)
  bool probe(std::string &err);
  bool isBam() const { return bam_; }

  // Decodes hits into `dst` until it is full or the input ends.  A batch never ends in
  // the middle of a run of records sharing a read name (unless one run fills the whole
  // buffer).  Returns the number of hits written; 0 = end of input.
  // If `names` is given, the name of every hit is appended (for -m).
  size_t nextBatch(const HitBuffers &dst, std::vector<std::string> *names = nullptr);

  // -y unique: the reference only looks at hits with NH = 1 (mm:1773), so a chromosome the annotation does not know is only
  // reported (mm:1297) when such a hit lies on it
  void warnOnlyForUniqueHits(bool on) { uniqueOnly_ = on; }
  uint64_t recordsRead() const { return nRecords_; }  // reference's "lines read" (= hits, mm:1772)
  std::string takeWarnings();                         // unknown chromosomes, CIGAR problems, XA problems
  // Read-key verification.  The device tells reads apart by the 64-bit key of the name, the reference by the name string
  // (mm:1656-1662, mm:1671): whenever two neighbouring records carry the same key their names are compared here.  Empty = every
  // such pair had equal names; else the first pair of DIFFERENT names sharing a key ("'a' and 'b'"), which the device would count
  // as one read.  (Name-grouped input: every pair that could be merged is a neighbouring pair.  Coordinate-sorted input: the
  // pairs that are not neighbours are not seen; DESIGN.md section 7 gives their probability.)
  const std::string &keyCollision() const { return keyCollision_; }

 private:
  struct Hit { uint32_t start, end, meta, nh; uint64_t key; };
  struct Alt { uint32_t chrMeta; bool strand; uint64_t start; std::vector<std::pair<char, int> > cigar; };

  bool openBam(std::string &err);
  bool fillRaw(size_t need);  // make `need` bytes available at rawPos_ (BAM)
  XamReader(const XamReader &parent, int);  // parser clone (decodeBamChunkParallel)
  bool decodeBamRecord();
  void parseBamRecord(const unsigned char *p, uint32_t blockSize);
  size_t decodeBamChunkParallel();
  bool decodeSamRecord();
  void pushRecordHits(const std::string &name, uint32_t chrMeta, uint64_t start, bool strand,
                      const std::vector<std::pair<char, int> > &cigar, bool cigarIsStar, uint32_t nHits);
  uint64_t cigarEnd(uint64_t start, uint64_t prevEnd, const std::vector<std::pair<char, int> > &cigar);
  uint32_t chrMetaOf(const std::string &name, bool quiet = false);
  void checkKey(const std::string &name, uint64_t key);
  void parseAlternatives(const std::string &xa);

  std::string fileName_;
  ReadsFormat format_;
  Strandedness strandedness_;
  const FeatureTable &features_;
  std::unordered_map<std::string, uint32_t> chrByName_;
  std::vector<std::string> unknownChr_;
  bool bam_ = false, over_ = false;
  gzFile gz_ = nullptr;
  BgzfSource bgzf_;        // multi-threaded inflate when the file is BGZF (every BAM is); gz_ otherwise
  bool useBgzf_ = false;
  long readRaw(unsigned char *dst, size_t cap);
  std::ifstream sam_;
  std::vector<unsigned char> raw_;
  size_t rawPos_ = 0, rawEnd_ = 0;
  std::vector<uint32_t> bamChrMeta_;  // per BAM refID
  std::vector<std::string> bamChrName_;
  uint32_t nMismatches_ = 0;          // persists across records like XamRecord::nMismatches (mm:596)
  std::vector<Alt> alts_;
  std::string nameBuf_, lastZBuf_;                  // per-record scratch, reused
  std::vector<std::pair<char, int> > cigarBuf_;
  std::vector<Hit> pending_;          // decoded but not yet handed out
  std::vector<std::string> pendingNames_;
  size_t pendingPos_ = 0;
  bool keepNames_ = false;
  uint64_t nRecords_ = 0;
  unsigned parseThreads_ = 1;   // record parsers working side by side on a chunk (BAM)
  bool clone_ = false;
  bool uniqueOnly_ = false;
  std::vector<size_t> recOff_;
  std::string warnings_;
  // read-key verification (keyCollision): the last name seen with its key; a clone also keeps the first of its range
  std::string prevName_, firstName_, keyCollision_;
  uint64_t prevKey_ = 0, firstKey_ = 0;
  bool havePrev_ = false;
  uint64_t keyMask_ = ~0ull;  // test knob MMANNOT_B200_KEY_MASK: keys cut down to a few bits so that collisions can be staged
};

}  // namespace mmb
