#include "xam.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

namespace mmb {

namespace {

const char CIGAR_OPS[] = "MIDNSHP=X";  // BAM op codes 0..8
const uint64_t COORD_LIMIT = 0xFFFFFFFEull;

inline uint32_t le32(const unsigned char *p) {
  return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) |
         (static_cast<uint32_t>(p[3]) << 24);
}
inline uint32_t le16(const unsigned char *p) { return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8); }

// text CIGAR -> (op, length) pairs; every non-digit closes a pair (mm:1347-1359)
void textCigar(const std::string &s, std::vector<std::pair<char, int> > &out) {
  out.clear();
  int v = 0;
  for (char c : s) {
    if (c >= '0' && c <= '9') v = v * 10 + (c - '0');
    else { out.push_back(std::make_pair(c, v)); v = 0; }
  }
}

inline bool strandMap(Strandedness s, bool forward) {
  // strandF / strandR / strandU, mm:836-844
  return s == Strandedness::F ? forward : s == Strandedness::R ? !forward : true;
}

}  // namespace

XamReader::XamReader(const std::string &fileName, ReadsFormat format, Strandedness strandedness, const FeatureTable &features)
    : fileName_(fileName), format_(format), strandedness_(strandedness), features_(features) {
  for (size_t i = 0; i < features.chromosomes.size(); ++i)
    if (features.chrHasFeatures[i]) chrByName_[features.chromosomes[i]] = static_cast<uint32_t>(i);
  if (const char *m = std::getenv("MMANNOT_B200_KEY_MASK")) keyMask_ = std::strtoull(m, nullptr, 0);
}

// Neighbouring records with one key must carry one name (see keyCollision() in the header).
void XamReader::checkKey(const std::string &name, uint64_t key) {
  if (!havePrev_) {
    havePrev_ = true;
    firstName_ = name; firstKey_ = key;
    prevName_ = name; prevKey_ = key;
  } else if (key != prevKey_) {
    prevName_ = name; prevKey_ = key;
  } else if (name != prevName_ && keyCollision_.empty()) {
    keyCollision_ = "'" + prevName_ + "' and '" + name + "'";
  }
}

XamReader::~XamReader() {
  if (gz_) gzclose(gz_);
}

std::string XamReader::takeWarnings() {
  std::string w;
  w.swap(warnings_);
  return w;
}

uint32_t XamReader::chrMetaOf(const std::string &name, bool quiet) {
  auto it = chrByName_.find(name);
  if (it != chrByName_.end()) return it->second;
  if (quiet) return HIT_CHR_NONE;  // (a hit the reference never looks at: no warning, and the name stays unreported)
  if (clone_) {  // the owning reader decides, in record order, whether the name is new (decodeBamChunkParallel)
    if (std::find(unknownChr_.begin(), unknownChr_.end(), name) == unknownChr_.end()) {
      warnings_ += '\x01' + name + "\n";
      unknownChr_.push_back(name);
    }
    return HIT_CHR_NONE;
  }
  if (std::find(unknownChr_.begin(), unknownChr_.end(), name) == unknownChr_.end()) {
    if (name != "*")
      warnings_ += "\t\tWarning!  Chromosome '" + name + "' (found in your reads) is not present in your annotation file.\n";
    unknownChr_.push_back(name);
  }
  return HIT_CHR_NONE;
}

bool XamReader::open(std::string &err) {
  if (!probe(err)) return false;
  if (!bam_) {
    sam_.open(fileName_.c_str());
    return true;
  }
  return openBam(err);
}

bool XamReader::probe(std::string &err) {
  {
    std::ifstream probe(fileName_.c_str());
    if (!probe.good()) {
      err = "Error, file '" + fileName_ + "' does not exists!";
      return false;
    }
  }
  ReadsFormat f = format_;
  if (f == ReadsFormat::UNKNOWN) {
    std::string suffix = fileName_.size() >= 4 ? lowered(fileName_.substr(fileName_.size() - 4)) : "";
    if (suffix == ".bam") f = ReadsFormat::BAM;
    else if (suffix == ".sam") f = ReadsFormat::SAM;
    else {
      err = "Cannot deduce type from file name '" + fileName_ + "'.  Should be a .sam or .bam file.  Please specify it using the '-f' option.";
      return false;
    }
  }
  bam_ = (f == ReadsFormat::BAM);
  return true;
}

bool XamReader::openBam(std::string &err) {
  {
    unsigned threads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (const char *e = std::getenv("MMANNOT_B200_DECODE_THREADS")) threads = static_cast<unsigned>(std::max(1, std::atoi(e)));
    useBgzf_ = bgzf_.open(fileName_, threads);
    parseThreads_ = threads;
  }
  if (!useBgzf_) {  // plain gzip (or uncompressed): the reference's own route, gzread
    gz_ = gzopen(fileName_.c_str(), "rb");
    if (!gz_) {
      err = "Cannot open file '" + fileName_ + "'.";
      return false;
    }
    gzbuffer(gz_, 1 << 20);
  }
  raw_.resize(8 << 20);
  if (!fillRaw(12) || std::memcmp(&raw_[rawPos_], "BAM\1", 4) != 0) {
    err = "Problem with file '" + fileName_ + "': file does not look like a BAM file (missing magic string).";
    return false;
  }
  uint32_t lText = le32(&raw_[rawPos_ + 4]);
  rawPos_ += 8;
  if (!fillRaw(static_cast<size_t>(lText) + 4)) { err = "Problem with file '" + fileName_ + "': truncated BAM header."; return false; }
  rawPos_ += lText;
  uint32_t nRef = le32(&raw_[rawPos_]);
  rawPos_ += 4;
  for (uint32_t i = 0; i < nRef; ++i) {
    if (!fillRaw(4)) { err = "Problem with file '" + fileName_ + "': truncated BAM header."; return false; }
    uint32_t lName = le32(&raw_[rawPos_]);
    rawPos_ += 4;
    if (!fillRaw(static_cast<size_t>(lName) + 4)) { err = "Problem with file '" + fileName_ + "': truncated BAM header."; return false; }
    std::string name(reinterpret_cast<const char *>(&raw_[rawPos_]), lName);
    name = name.c_str();  // the reference keeps the bytes up to the first NUL (mm:1510)
    rawPos_ += lName + 4;
    bamChrName_.push_back(name);
  }
  // chromosome ids are resolved lazily so that the "unknown chromosome" warnings appear
  // only for references that actually carry reads, like the reference does (mm:1293-1300)
  bamChrMeta_.assign(nRef, 0xFFFFFFFFu);
  return true;
}

long XamReader::readRaw(unsigned char *dst, size_t cap) {
  if (useBgzf_) return bgzf_.read(dst, cap);
  return gzread(gz_, dst, static_cast<unsigned>(cap));
}

bool XamReader::fillRaw(size_t need) {
  if (rawEnd_ - rawPos_ >= need) return true;
  if (rawPos_ > 0) {
    std::memmove(&raw_[0], &raw_[rawPos_], rawEnd_ - rawPos_);
    rawEnd_ -= rawPos_;
    rawPos_ = 0;
  }
  if (raw_.size() < need) raw_.resize(std::max(need, raw_.size() * 2));
  while (rawEnd_ < need) {
    long got = readRaw(&raw_[rawEnd_], std::min<size_t>(raw_.size() - rawEnd_, 1u << 30));
    if (got <= 0) return false;
    rawEnd_ += static_cast<size_t>(got);
  }
  // opportunistically top the buffer up so that most records need no further call
  if (rawEnd_ < raw_.size()) {
    long got = readRaw(&raw_[rawEnd_], std::min<size_t>(raw_.size() - rawEnd_, 1u << 30));
    if (got > 0) rawEnd_ += static_cast<size_t>(got);
  }
  return true;
}

uint64_t XamReader::cigarEnd(uint64_t start, uint64_t prevEnd, const std::vector<std::pair<char, int> > &cigar) {
  // Read::parseCigar, mm:852-875
  if (cigar.size() == 1 && cigar.front().first == '*') return prevEnd;
  uint64_t end = start;
  for (const auto &part : cigar) {
    switch (part.first) {
      case 'M': case 'D': case '=': case 'X':
        end += static_cast<uint64_t>(static_cast<int64_t>(part.second));
        break;
      case 'I': case 'S': case 'H': case 'P':
        break;
      default:
        warnings_ += std::string("Problem in the cigar: do not understand char ") + part.first + "\n";
    }
  }
  return end - 1;
}

void XamReader::parseAlternatives(const std::string &xa) {
  // Reader::parseAlternativeHit, mm:1360-1399 (a missing field re-uses the last field that was read)
  if (xa == "0") return;
  std::vector<std::string> pieces, f;
  split_getline(xa, ';', pieces);
  for (const std::string &piece : pieces) {
    if (piece.empty()) continue;
    split_getline(piece, ',', f);
    auto field = [&f](size_t i) -> const std::string & { return f[std::min(i, f.size() - 1)]; };
    bool ok = true;
    Alt alt;
    const std::string &chrName = field(0);
    const std::string &posField = field(1);
    alt.strand = (!posField.empty() && posField[0] == '+');
    if (posField.empty()) ok = false;
    if (ok) alt.start = parse_ulong(posField.substr(1), ok);
    unsigned long nm = 0;
    if (ok) nm = parse_ulong(field(3), ok);
    if (!ok) {
      warnings_ += "Warning!  Problem while parsing an \"XA\" tag, which is probably too long:\n" + xa + "\n";
      continue;
    }
    if (nm == nMismatches_) {
      alt.chrMeta = chrMetaOf(chrName);
      textCigar(field(2), alt.cigar);
      alts_.push_back(alt);
    }
  }
}

void XamReader::pushRecordHits(const std::string &name, uint32_t chrMeta, uint64_t start, bool strand,
                               const std::vector<std::pair<char, int> > &cigar, bool, uint32_t nHits) {
  const uint64_t key = name_key(name.data(), name.size()) & keyMask_;
  checkKey(name, key);
  uint64_t end = start;  // Read::reset, mm:878-879
  auto emit = [&](uint32_t cm, uint64_t s, uint64_t e, bool fwd) {
    Hit h;
    uint32_t chr = cm;
    if (s > COORD_LIMIT) { chr = HIT_CHR_NONE; s = 0; e = 0; }         // cannot touch any feature
    else if (e == ~0ull) e = 0xFFFFFFFFull;                             // start 0, empty CIGAR: wraps like the reference's unsigned long
    else if (e > COORD_LIMIT) e = COORD_LIMIT;
    h.start = static_cast<uint32_t>(s);
    h.end = static_cast<uint32_t>(e);
    h.meta = (chr & HIT_CHR_MASK) | (strandMap(strandedness_, fwd) ? HIT_STRAND_BIT : 0u);
    h.nh = nHits;
    h.key = key;
    pending_.push_back(h);
    if (keepNames_) pendingNames_.push_back(name);
    ++nRecords_;
  };
  end = cigarEnd(start, end, cigar);
  emit(chrMeta, start, end, strand);
  for (const Alt &a : alts_) {  // Read::setNextAlternativeHit, mm:891-896
    end = cigarEnd(a.start, end, a.cigar);
    emit(a.chrMeta, a.start, end, a.strand);
  }
}

bool XamReader::decodeBamRecord() {
  if (decodeBamChunkParallel() > 0) return true;
  if (!fillRaw(4)) return false;
  uint32_t blockSize = le32(&raw_[rawPos_]);
  if (!fillRaw(static_cast<size_t>(blockSize) + 4)) {
    warnings_ += "Warning!  Truncated BAM record at the end of '" + fileName_ + "'.\n";
    return false;
  }
  const unsigned char *p = &raw_[rawPos_ + 4];
  rawPos_ += static_cast<size_t>(blockSize) + 4;
  parseBamRecord(p, blockSize);
  return true;
}

// One BAM alignment record (the bytes after its block_size field) -> hits.  Touches only the parser state of `this`
// (scratch buffers, NM carried from record to record, chromosome cache, pending hits, warnings), so that the records of a
// chunk can be parsed by several parser clones side by side (decodeBamChunkParallel).
void XamReader::parseBamRecord(const unsigned char *p, uint32_t blockSize) {
  const unsigned char *recEnd = p + blockSize;
  if (blockSize < 32) return;  // malformed, skip
  int32_t refId = static_cast<int32_t>(le32(p));
  int32_t pos = static_cast<int32_t>(le32(p + 4));
  uint32_t lReadName = le32(p + 8) & 0xff;
  uint32_t flagNc = le32(p + 12);
  uint32_t flag = flagNc >> 16, nCigar = flagNc & 0xffff;
  uint32_t lSeq = le32(p + 16);
  const unsigned char *q = p + 32;
  if (q + lReadName + 4ull * nCigar + (lSeq + 1ull) / 2 + lSeq > recEnd) return;  // malformed, skip
  // (buffers reused from record to record: no allocation on the hot path)
  std::string &name = nameBuf_;
  name.assign(reinterpret_cast<const char *>(q), strnlen(reinterpret_cast<const char *>(q), lReadName));  // up to the first NUL (mm:1545)
  q += lReadName;
  std::vector<std::pair<char, int> > &cigar = cigarBuf_;
  cigar.clear();
  for (uint32_t i = 0; i < nCigar; ++i, q += 4) {
    uint32_t v = le32(q);
    uint32_t op = v & 15;
    cigar.push_back(std::make_pair(op < 9 ? CIGAR_OPS[op] : '?', static_cast<int>(v >> 4)));
  }
  q += (lSeq + 1) / 2 + lSeq;
  uint32_t nHits = 1;
  alts_.clear();
  std::string &lastZ = lastZBuf_;
  lastZ.clear();
  while (q + 3 <= recEnd) {  // aux fields (mm:1563-1648); unsigned integer types only feed NH / NM (mm:1596-1618)
    char t0 = static_cast<char>(q[0]), t1 = static_cast<char>(q[1]), ty = static_cast<char>(q[2]);
    q += 3;
    uint32_t vU = 0;
    bool bad = false;
    switch (ty) {
      case 'A': case 'c': q += 1; break;
      case 'C': if (q + 1 <= recEnd) vU = q[0]; q += 1; break;
      case 's': q += 2; break;
      case 'S': if (q + 2 <= recEnd) vU = le16(q); q += 2; break;
      case 'i': case 'f': q += 4; break;
      case 'I': if (q + 4 <= recEnd) vU = le32(q); q += 4; break;
      case 'Z': case 'H': {
        const unsigned char *z = q;
        while (z < recEnd && *z) ++z;
        if (ty == 'Z') lastZ.assign(reinterpret_cast<const char *>(q), z - q);
        q = z + 1;
        break;
      }
      case 'B': {
        if (q + 5 > recEnd) { bad = true; break; }
        char sub = static_cast<char>(q[0]);
        uint64_t cnt = le32(q + 1);
        uint64_t width = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        q += 5 + cnt * width;
        break;
      }
      default:
        warnings_ += std::string("Problem with tag type '") + ty + "'\n";
        bad = true;
    }
    if (bad || q > recEnd) break;
    if (t0 == 'N' && t1 == 'H') { if (alts_.empty()) nHits = vU; }
    else if (t0 == 'N' && t1 == 'M') nMismatches_ = vU;
    else if (t0 == 'X' && t1 == 'A') {
      if (lastZ != "0") { parseAlternatives(lastZ); nHits = static_cast<uint32_t>(alts_.size()) + 1; }
    }
  }
  uint32_t chrMeta;
  const bool quiet = uniqueOnly_ && nHits != 1;
  if (refId < 0 || static_cast<size_t>(refId) >= bamChrMeta_.size()) chrMeta = chrMetaOf("*");
  else if (bamChrMeta_[refId] != 0xFFFFFFFFu) chrMeta = bamChrMeta_[refId];
  else {
    chrMeta = chrMetaOf(bamChrName_[refId], quiet);
    if (!quiet || chrMeta != HIT_CHR_NONE) bamChrMeta_[refId] = chrMeta;
  }
  uint64_t start = static_cast<uint64_t>(static_cast<int64_t>(pos) + 1);  // ++pos then widened (mm:1536-1537)
  pushRecordHits(name, chrMeta, start, (flag & 0x10) == 0, cigar, false, nHits);
}

namespace {
// NM as the record parser leaves it after this record (mm:1596-1618: only the unsigned integer types carry a value, any
// other type of an NM tag gives 0); false when the record has no NM tag
bool lastNmOfRecord(const unsigned char *p, uint32_t blockSize, uint32_t &nm) {
  if (blockSize < 32) return false;
  const unsigned char *recEnd = p + blockSize;
  const uint32_t lReadName = le32(p + 8) & 0xff, nCigar = le32(p + 12) & 0xffff, lSeq = le32(p + 16);
  const unsigned char *q = p + 32;
  if (q + lReadName + 4ull * nCigar + (lSeq + 1ull) / 2 + lSeq > recEnd) return false;
  q += lReadName + 4ull * nCigar + (lSeq + 1ull) / 2 + lSeq;
  bool found = false;
  while (q + 3 <= recEnd) {
    const char t0 = static_cast<char>(q[0]), t1 = static_cast<char>(q[1]), ty = static_cast<char>(q[2]);
    q += 3;
    uint32_t vU = 0;
    bool bad = false;
    switch (ty) {
      case 'A': case 'c': q += 1; break;
      case 'C': if (q + 1 <= recEnd) vU = q[0]; q += 1; break;
      case 's': q += 2; break;
      case 'S': if (q + 2 <= recEnd) vU = le16(q); q += 2; break;
      case 'i': case 'f': q += 4; break;
      case 'I': if (q + 4 <= recEnd) vU = le32(q); q += 4; break;
      case 'Z': case 'H': { while (q < recEnd && *q) ++q; ++q; break; }
      case 'B': {
        if (q + 5 > recEnd) { bad = true; break; }
        const char sub = static_cast<char>(q[0]);
        const uint64_t cnt = le32(q + 1), width = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        q += 5 + cnt * width;
        break;
      }
      default: bad = true;
    }
    if (bad || q > recEnd) break;
    if (t0 == 'N' && t1 == 'M') { nm = vU; found = true; }
  }
  return found;
}
}  // namespace

// Parses every complete record currently in the raw buffer with several parser clones, each on a contiguous range of
// records, and appends their hits, warnings and counters in record order.  Returns the number of records consumed (0: not
// enough data for this to pay off, the caller parses one record the usual way).
size_t XamReader::decodeBamChunkParallel() {
  if (parseThreads_ < 2 || keepNames_) return 0;
  if (raw_.size() < (32u << 20)) raw_.resize(32u << 20);
  fillRaw(std::min<size_t>(raw_.size(), 4));  // tops the buffer up (no-op at end of input)
  // index the complete records
  recOff_.clear();
  size_t pos = rawPos_;
  while (pos + 4 <= rawEnd_) {
    const size_t bs = le32(&raw_[pos]);
    if (pos + 4 + bs > rawEnd_) break;
    recOff_.push_back(pos);
    pos += 4 + bs;
  }
  const size_t nRec = recOff_.size();
  if (nRec < 4096) return 0;
  const unsigned nT = static_cast<unsigned>(std::min<size_t>(parseThreads_, nRec / 1024));
  std::vector<std::unique_ptr<XamReader> > clones;
  for (unsigned t = 0; t < nT; ++t) clones.emplace_back(new XamReader(*this, 0));
  auto work = [&](unsigned t) {
    XamReader &c = *clones[t];
    const size_t lo = nRec * t / nT, hi = nRec * (t + 1) / nT;
    // NM carried into the range: from the nearest earlier record of the chunk that has the tag, else the reader's own
    c.nMismatches_ = nMismatches_;
    for (size_t k = lo; k-- > 0;) {
      uint32_t nm;
      if (lastNmOfRecord(&raw_[recOff_[k] + 4], le32(&raw_[recOff_[k]]), nm)) { c.nMismatches_ = nm; break; }
    }
    c.pending_.reserve((hi - lo) + (hi - lo) / 8);
    for (size_t k = lo; k < hi; ++k) c.parseBamRecord(&raw_[recOff_[k] + 4], le32(&raw_[recOff_[k]]));
  };
  {
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nT; ++t) pool.emplace_back(work, t);
    work(0);
    for (std::thread &th : pool) th.join();
  }
  for (unsigned t = 0; t < nT; ++t) {  // merge in record order
    XamReader &c = *clones[t];
    pending_.insert(pending_.end(), c.pending_.begin(), c.pending_.end());
    nRecords_ += c.nRecords_;
    // read-key verification across the border of two ranges, then the clone's own finding
    if (c.havePrev_) {
      if (havePrev_ && c.firstKey_ == prevKey_ && c.firstName_ != prevName_ && keyCollision_.empty())
        keyCollision_ = "'" + prevName_ + "' and '" + c.firstName_ + "'";
      if (keyCollision_.empty()) keyCollision_ = c.keyCollision_;
      if (!havePrev_) { firstName_ = c.firstName_; firstKey_ = c.firstKey_; }
      havePrev_ = true;
      prevName_ = c.prevName_; prevKey_ = c.prevKey_;
    }
    // warnings: lines starting with \x01 announce a chromosome the clone did not know
    size_t a = 0;
    while (a < c.warnings_.size()) {
      size_t b = c.warnings_.find('\n', a);
      if (b == std::string::npos) b = c.warnings_.size() - 1;
      if (c.warnings_[a] == '\x01') (void)chrMetaOf(c.warnings_.substr(a + 1, b - a - 1));
      else warnings_.append(c.warnings_, a, b - a + 1);
      a = b + 1;
    }
    for (size_t r = 0; r < bamChrMeta_.size(); ++r)
      if (bamChrMeta_[r] == 0xFFFFFFFFu && c.bamChrMeta_[r] != 0xFFFFFFFFu) bamChrMeta_[r] = c.bamChrMeta_[r];
  }
  nMismatches_ = clones[nT - 1]->nMismatches_;
  rawPos_ = pos;
  return nRec;
}

// parser clone: the read-only tables of `parent`, fresh parser state
XamReader::XamReader(const XamReader &parent, int)
    : fileName_(parent.fileName_), format_(parent.format_), strandedness_(parent.strandedness_), features_(parent.features_),
      chrByName_(parent.chrByName_), unknownChr_(), bam_(true), bamChrMeta_(parent.bamChrMeta_), bamChrName_(parent.bamChrName_),
      clone_(true), uniqueOnly_(parent.uniqueOnly_), keyMask_(parent.keyMask_) {}

bool XamReader::decodeSamRecord() {
  std::string line;
  do {
    if (!std::getline(sam_, line)) return false;
  } while (line.empty() || line[0] == '@' || line[0] == '#');
  std::vector<std::string> col;
  split_getline(line, '\t', col);
  if (col.size() < 12) {
    warnings_ += "Error, SAM line with fewer than 12 columns (mmannot needs at least one optional tag, e.g. NH): '" + line + "'\n";
    over_ = true;
    return false;
  }
  bool ok;
  unsigned long flag = parse_ulong(col[1], ok);
  uint64_t start = parse_ulong(col[3], ok);
  uint32_t nHits = 1;
  alts_.clear();
  std::vector<std::pair<char, int> > cigar;
  textCigar(col[5], cigar);
  for (size_t i = 11; i < col.size(); ++i) {  // mm:1461-1477
    const std::string &part = col[i];
    size_t p1 = part.find(':');
    std::string key = part.substr(0, p1);
    size_t p2 = (p1 == std::string::npos) ? part.find(':') : part.find(':', p1 + 1);
    std::string value = (p2 == std::string::npos) ? part : part.substr(p2 + 1);
    if (key == "NH") {
      if (alts_.empty()) { unsigned long v = parse_ulong(value, ok); if (ok) nHits = static_cast<uint32_t>(v); }
    } else if (key == "NM") {
      unsigned long v = parse_ulong(value, ok);
      if (ok) nMismatches_ = static_cast<uint32_t>(v);
    } else if (key == "XA") {
      if (value != "0") { parseAlternatives(value); nHits = static_cast<uint32_t>(alts_.size()) + 1; }
    }
  }
  pushRecordHits(col[0], chrMetaOf(col[2], uniqueOnly_ && nHits != 1), start, (flag & 0x10) == 0, cigar, false, nHits);
  return true;
}

size_t XamReader::nextBatch(const HitBuffers &dst, std::vector<std::string> *names) {
  keepNames_ = (names != nullptr);
  const size_t cap = dst.capacity;
  // decode until strictly more than `cap` hits are pending (so that the hit after the cut is known) or input ends
  while (!over_ && pending_.size() - pendingPos_ <= cap) {
    bool more = bam_ ? decodeBamRecord() : decodeSamRecord();
    if (!more) over_ = true;
  }
  size_t avail = pending_.size() - pendingPos_;
  size_t n = std::min(avail, cap);
  if (n < avail) {  // cut on a read-name boundary when there is one
    size_t m = n;
    while (m > 0 && pending_[pendingPos_ + m].key == pending_[pendingPos_ + m - 1].key) --m;
    if (m > 0) n = m;
  }
  for (size_t i = 0; i < n; ++i) {
    const Hit &h = pending_[pendingPos_ + i];
    dst.start[i] = h.start; dst.end[i] = h.end; dst.meta[i] = h.meta; dst.nh[i] = h.nh; dst.key[i] = h.key;
  }
  if (names) {
    names->clear();
    names->insert(names->end(), pendingNames_.begin() + pendingPos_, pendingNames_.begin() + pendingPos_ + n);
  }
  pendingPos_ += n;
  if (pendingPos_ == pending_.size()) {
    pending_.clear(); pendingNames_.clear(); pendingPos_ = 0;
  } else if (pendingPos_ > (1u << 16)) {
    pending_.erase(pending_.begin(), pending_.begin() + pendingPos_);
    if (keepNames_) pendingNames_.erase(pendingNames_.begin(), pendingNames_.begin() + pendingPos_);
    pendingPos_ = 0;
  }
  return n;
}

}  // namespace mmb
