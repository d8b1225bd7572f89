#include "counter.hpp"

#include "bam_device.hpp"

#include <algorithm>
#include <cstdio>
#include <iomanip>
#include <sstream>

namespace mmb {

std::string withThousands(uint64_t n) {
  std::string digits = std::to_string(n), out;
  for (size_t i = 0; i < digits.size(); ++i) {
    if (i && (digits.size() - i) % 3 == 0) out += ',';
    out += digits[i];
  }
  return out;
}

namespace {

// printStats, mm:139-143
void printStats(std::ostream &log, uint64_t n, const char *label, uint64_t total) {
  unsigned int size = static_cast<unsigned int>(std::log10(static_cast<double>(total)) + 1);
  size += static_cast<unsigned int>(size / 3.0);
  char pct[64];
  std::snprintf(pct, sizeof(pct), "%5.1f", static_cast<double>(static_cast<float>(n) / total * 100));
  log << "\t" << label << std::setw(static_cast<int>(size)) << withThousands(n) << " (" << pct << "%)\n";
}

bool allocPinned(HitBuffers &b, size_t cap) {
  b.capacity = cap;
  b.start = static_cast<uint32_t *>(mma_alloc_pinned(cap * 4));
  b.end = static_cast<uint32_t *>(mma_alloc_pinned(cap * 4));
  b.meta = static_cast<uint32_t *>(mma_alloc_pinned(cap * 4));
  b.nh = static_cast<uint32_t *>(mma_alloc_pinned(cap * 4));
  b.key = static_cast<uint64_t *>(mma_alloc_pinned(cap * 8));
  return b.start && b.end && b.meta && b.nh && b.key;
}
void freePinned(HitBuffers &b) {
  mma_free_pinned(b.start); mma_free_pinned(b.end); mma_free_pinned(b.meta); mma_free_pinned(b.nh); mma_free_pinned(b.key);
  b = HitBuffers();
}

}  // namespace

Counter::Counter(mma_ctx *ctx, const FeatureTable &features, const Config &config, const RunOptions &opt)
    : ctx_(ctx), features_(features), config_(config), opt_(opt), stats_() {
  allocPinned(pinned_[0], opt.batchHits);
  allocPinned(pinned_[1], opt.batchHits);
  for (PackedBuffers &p : packedBuf_) {
    const size_t cap = opt.batchHits;
    p.escCapacity = cap / 64 + 16;
    p.packed = static_cast<uint32_t *>(mma_alloc_pinned(cap * 4));
    p.runKey = static_cast<uint64_t *>(mma_alloc_pinned(cap * 8));
    p.tileRunBase = static_cast<uint32_t *>(mma_alloc_pinned((cap / MMA_PACK_TILE + 1) * 4));
    p.escIndex = static_cast<uint32_t *>(mma_alloc_pinned(p.escCapacity * 4));
    p.escEnd = static_cast<uint32_t *>(mma_alloc_pinned(p.escCapacity * 4));
    p.escNh = static_cast<uint32_t *>(mma_alloc_pinned(p.escCapacity * 4));
  }
}

Counter::Counter(const std::vector<mma_ctx *> &ctxs, const FeatureTable &features, const Config &config, const RunOptions &opt)
    : Counter(ctxs.empty() ? nullptr : ctxs[0], features, config, opt) {
  if (ctxs.size() < 2) return;
  shards_.resize(ctxs.size());
  for (size_t g = 0; g < ctxs.size(); ++g) {
    Shard &sh = shards_[g];
    sh.ctx = ctxs[g];
    for (int k = 0; k < 2; ++k) {
      allocPinned(sh.pinned[k], opt.batchHits);
      PackedBuffers &p = sh.packed[k];
      const size_t cap = opt.batchHits;
      p.escCapacity = cap / 64 + 16;
      p.packed = static_cast<uint32_t *>(mma_alloc_pinned(cap * 4));
      p.runKey = static_cast<uint64_t *>(mma_alloc_pinned(cap * 8));
      p.tileRunBase = static_cast<uint32_t *>(mma_alloc_pinned((cap / MMA_PACK_TILE + 1) * 4));
      p.escIndex = static_cast<uint32_t *>(mma_alloc_pinned(p.escCapacity * 4));
      p.escEnd = static_cast<uint32_t *>(mma_alloc_pinned(p.escCapacity * 4));
      p.escNh = static_cast<uint32_t *>(mma_alloc_pinned(p.escCapacity * 4));
    }
  }
}

Counter::~Counter() {
  for (Shard &sh : shards_)
    for (int k = 0; k < 2; ++k) {
      freePinned(sh.pinned[k]);
      PackedBuffers &p = sh.packed[k];
      mma_free_pinned(p.packed); mma_free_pinned(p.runKey); mma_free_pinned(p.tileRunBase);
      mma_free_pinned(p.escIndex); mma_free_pinned(p.escEnd); mma_free_pinned(p.escNh);
    }
  freePinned(pinned_[0]);
  freePinned(pinned_[1]);
  for (PackedBuffers &p : packedBuf_) {
    mma_free_pinned(p.packed); mma_free_pinned(p.runKey); mma_free_pinned(p.tileRunBase);
    mma_free_pinned(p.escIndex); mma_free_pinned(p.escEnd); mma_free_pinned(p.escNh);
  }
}

bool Counter::read(const std::string &fileName, uint32_t column, std::string &err, std::ostream &log) {
  fileName_ = fileName;
  counts_.clear();
  stats_ = mma_sample_stats();
  if (!pinned_[0].start || !pinned_[1].key) { err = "Cannot allocate page-locked hit buffers."; return false; }
  XamReader reader(fileName, opt_.format, opt_.strandedness, features_);
  if (!reader.probe(err)) return false;
  reader.warnOnlyForUniqueHits(opt_.strategy == MMA_STRATEGY_UNIQUE);
  log << (reader.isBam() ? "Reading BAM file " : "Reading SAM file ") << fileName << std::endl;
  // BAM on one GPU without per-read statistics: the compressed file goes to the device as it is (inflate + record parse there);
  // files that route does not take (XA alternative hits, ...) are decoded on the host below, from the start
  bool onDevice = false;
  uint64_t deviceRecords = 0;
  if (reader.isBam() && !writers_ && shards_.empty() && !std::getenv("MMANNOT_B200_HOST_DECODE")) {
    if (mma_reset_sample(ctx_, column) != MMA_OK) { err = mma_last_error(ctx_); return false; }
    DeviceBamFeeder feeder(ctx_, features_, opt_.strandedness);
    std::string w, why;
    const DeviceBamFeeder::Result r = feeder.run(fileName, column, deviceRecords, w, why, err);
    if (r == DeviceBamFeeder::Result::FAILED) return false;
    if (r == DeviceBamFeeder::Result::DONE) {
      onDevice = true;
      if (!w.empty()) log << w;
    } else if (std::getenv("MMANNOT_B200_VERBOSE")) {
      log << "\t(device BAM decoder not used: " << why << ")" << std::endl;
    }
  }
  if (onDevice) {
    log << "\t" << withThousands(deviceRecords) << " lines read, done." << std::endl;
  } else {
  if (!reader.open(err)) return false;
  if (!shards_.empty()) {
    if (writers_) { err = "Read / interval statistics need the whole input on one GPU."; return false; }
    if (!readSharded(reader, column, err, log)) return false;
  } else {
  if (mma_reset_sample(ctx_, column) != MMA_OK) { err = mma_last_error(ctx_); return false; }
  // decode batch k+1 on the host while batch k is copied and annotated on the device
  std::vector<std::string> names;
  std::vector<uint64_t> masks, offsets;
  for (unsigned which = 0;; which ^= 1) {
    const HitBuffers &buf = pinned_[which];
    size_t n = reader.nextBatch(buf, writers_ ? &names : nullptr);
    std::string w = reader.takeWarnings();
    if (!w.empty()) log << w;
    if (n == 0) break;
    mma_hit_batch b;
    b.n = n; b.start = buf.start; b.end = buf.end; b.meta = buf.meta; b.nh = buf.nh; b.read_key = buf.key;
    // compact format over PCIe when the batch fits it (short reads, NH < 255 with a few escapes), else the wide arrays
    const PackedBuffers &pk = packedBuf_[which];
    mma_packed_batch pb;
    const bool canPack = pk.packed && pk.runKey && pk.tileRunBase && pk.escIndex && pk.escEnd && pk.escNh &&
                         mma_pack_hits(&b, pk.packed, pk.runKey, pk.tileRunBase, pk.escIndex, pk.escEnd, pk.escNh, pk.escCapacity, &pb) == MMA_OK;
    const int src = canPack ? mma_submit_hits_packed(ctx_, column, &pb) : mma_submit_hits(ctx_, column, &b);
    if (src != MMA_OK) { err = mma_last_error(ctx_); return false; }
    if (writers_) {  // -m / -M: scan alone for the same hits, then the name-keyed text bookkeeping on the host
      masks.resize(n);
      offsets.assign(n + 1, 0);
      const uint32_t *ids = nullptr;
      const int rc = opt_.intervalStats ? mma_annotate_intervals(ctx_, &b, masks.data(), offsets.data(), &ids)
                                        : mma_annotate_hits(ctx_, &b, masks.data());
      if (rc != MMA_OK) { err = mma_last_error(ctx_); return false; }
      for (size_t i = 0; i < n; ++i) {
        if (opt_.strategy == MMA_STRATEGY_UNIQUE && buf.nh[i] != 1) continue;  // mm:1773
        writers_->addHit(names[i], buf.nh[i], masks[i], ids ? ids + offsets[i] : nullptr, ids ? offsets[i + 1] - offsets[i] : 0);
      }
    }
    if (opt_.progress) log << "\t" << withThousands(reader.recordsRead()) << " lines read.\r" << std::flush;
  }
  }
  log << "\t" << withThousands(reader.recordsRead()) << " lines read, done." << std::endl;
  // the strategies that group records by read (mm:1669-1702, mm:1706-1726) rely on the 64-bit key standing for the name
  if ((opt_.strategy == MMA_STRATEGY_DEFAULT || opt_.strategy == MMA_STRATEGY_RANDOM) && !reader.keyCollision().empty()) {
    err = "Read names " + reader.keyCollision() + " of '" + fileName + "' share one 64-bit read key: their hits would be counted as one read.";
    return false;
  }
  }
  if (writers_) writers_->endOfFile();
  mma_sample_result res;
  if (mma_finish_sample(ctx_, column, &res) != MMA_OK) { err = mma_last_error(ctx_); return false; }
  stats_ = res.stats;
  // regionCounts value of every element set (mm:1658, 1730): count * 1/NH summed over NH under -y ratio
  std::vector<size_t> order(res.n_rows);
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&res](size_t a, size_t b) {
    return res.row_mask[a] != res.row_mask[b] ? res.row_mask[a] < res.row_mask[b] : res.row_nh[a] < res.row_nh[b];
  });
  for (size_t i : order) {
    const double w = res.row_nh[i] ? 1.0 / res.row_nh[i] : 1.0;
    counts_[res.row_mask[i]] += static_cast<double>(res.row_count[i]) * w;
  }
  return true;
}

// One input over several GPUs: every decoded batch is dealt out by read name, each GPU gets its share in file order
bool Counter::readSharded(XamReader &reader, uint32_t column, std::string &err, std::ostream &log) {
  const size_t nG = shards_.size();
  for (Shard &sh : shards_)
    if (mma_reset_sample(sh.ctx, column) != MMA_OK) { err = mma_last_error(sh.ctx); return false; }
  for (unsigned which = 0;; which ^= 1) {
    const HitBuffers &buf = pinned_[which & 1];  // (only a staging area here: nothing is copied to a GPU from it)
    const size_t n = reader.nextBatch(buf, nullptr);
    std::string w = reader.takeWarnings();
    if (!w.empty()) log << w;
    if (n == 0) break;
    for (Shard &sh : shards_) sh.n = 0;
    for (size_t i = 0; i < n; ++i) {
      uint64_t k = buf.key[i];
      k = (k ^ (k >> 33)) * 0xff51afd7ed558ccdull;
      k ^= k >> 33;
      Shard &sh = shards_[k % nG];
      const HitBuffers &d = sh.pinned[which];
      const size_t at = sh.n++;
      d.start[at] = buf.start[i]; d.end[at] = buf.end[i]; d.meta[at] = buf.meta[i]; d.nh[at] = buf.nh[i]; d.key[at] = buf.key[i];
    }
    for (Shard &sh : shards_) {
      const HitBuffers &d = sh.pinned[which];
      mma_hit_batch b;
      b.n = sh.n; b.start = d.start; b.end = d.end; b.meta = d.meta; b.nh = d.nh; b.read_key = d.key;
      if (b.n == 0) {  // nothing for this GPU in this batch: the call still waits for its previous copy, which frees the other slot
        if (mma_submit_hits(sh.ctx, column, &b) != MMA_OK) { err = mma_last_error(sh.ctx); return false; }
        continue;
      }
      const PackedBuffers &pk = sh.packed[which];
      mma_packed_batch pb;
      const bool canPack = pk.packed && pk.runKey && pk.tileRunBase && pk.escIndex && pk.escEnd && pk.escNh &&
                           mma_pack_hits(&b, pk.packed, pk.runKey, pk.tileRunBase, pk.escIndex, pk.escEnd, pk.escNh, pk.escCapacity, &pb) == MMA_OK;
      const int rc = canPack ? mma_submit_hits_packed(sh.ctx, column, &pb) : mma_submit_hits(sh.ctx, column, &b);
      if (rc != MMA_OK) { err = mma_last_error(sh.ctx); return false; }
    }
    if (opt_.progress) log << "\t" << withThousands(reader.recordsRead()) << " lines read.\r" << std::flush;
  }
  // the sum TableCount::addCounter would form had one Counter seen everything (mm:1861-1876), on the devices
  std::vector<mma_ctx *> ctxs;
  for (Shard &sh : shards_) ctxs.push_back(sh.ctx);
  if (mma_allreduce(ctxs.data(), static_cast<uint32_t>(ctxs.size()), column) != MMA_OK) { err = mma_last_error(ctxs[0]); return false; }
  return true;
}

void Counter::dump(std::ostream &log) const {
  log << "Results for " << fileName_ << ":" << std::endl;
  if (stats_.n_hits == 0) {
    log << "\tNo hit." << std::endl;
    return;
  }
  log << "\t# reads:                       " << withThousands(stats_.n_reads) << "\n";
  printStats(log, stats_.n_unique, "# uniquely mapped reads:       ", stats_.n_reads);
  printStats(log, stats_.n_rescued, "# multi-mapping rescued reads: ", stats_.n_reads);
  log << "\t# hits:                        " << withThousands(stats_.n_hits) << "\n";
  printStats(log, stats_.n_ambiguous, "# ambiguous hits:              ", stats_.n_hits);
  printStats(log, stats_.n_unassigned, "# unassigned hits:             ", stats_.n_hits);
}

void TableCount::addCounts(const std::map<uint64_t, double> &counts) {
  for (const auto &kv : counts) {
    std::vector<size_t> elements;
    for (size_t i = 0; i < 64; ++i) if ((kv.first >> i) & 1) elements.push_back(i);
    std::vector<unsigned int> &row = rows_[elements];
    if (row.empty()) row.assign(nInputs_, 0);
    row[nColumns_] = static_cast<unsigned int>(std::round(kv.second));
  }
  ++nColumns_;
}

void TableCount::dump(std::ostream &out, const std::vector<std::string> &samples) const {
  out << "Type";
  for (const std::string &s : samples) out << "\t" << s;
  out << "\n";
  for (const auto &row : rows_) {
    for (size_t k = 0; k < row.first.size(); ++k) out << (k ? "--" : "") << config_.getName(row.first[k]);
    for (unsigned int v : row.second) out << "\t" << v;
    out << "\n";
  }
}

}  // namespace mmb
