#include "bam_device.hpp"

#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace mmb {

namespace {

inline uint32_t rd32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24); }
inline uint32_t rd16(const unsigned char *p) { return p[0] | (p[1] << 8); }

// total size of the BGZF member starting at p (0: not a whole BGZF member header in `avail` bytes / not BGZF); *hdr = header length
size_t bgzfMember(const unsigned char *p, size_t avail, size_t *hdr) {
  if (avail < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || p[3] != 4) return 0;
  const size_t xlen = rd16(p + 10);
  if (avail < 12 + xlen) return 0;
  for (size_t at = 12; at + 4 <= 12 + xlen;) {
    const size_t slen = rd16(p + at + 2);
    if (p[at] == 'B' && p[at + 1] == 'C' && slen == 2 && at + 6 <= 12 + xlen) {
      *hdr = 12 + xlen;
      return static_cast<size_t>(rd16(p + at + 4)) + 1;
    }
    at += 4 + slen;
  }
  return 0;
}

bool inflateRaw(const unsigned char *src, size_t n, unsigned char *dst, size_t want) {
  z_stream zs;
  std::memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, -15) != Z_OK) return false;
  zs.next_in = const_cast<unsigned char *>(src); zs.avail_in = static_cast<uInt>(n);
  zs.next_out = dst; zs.avail_out = static_cast<uInt>(want);
  const int rc = inflate(&zs, Z_FINISH);
  const bool ok = (rc == Z_STREAM_END) && zs.total_out == want;
  inflateEnd(&zs);
  return ok;
}

// `want` bytes of the file from `offset` into dst, by several threads (one thread copies out of the page cache at 2-4 GB/s on the
// hosts this runs on: less than the link to the device takes).  Returns the bytes read (less than `want` only at the end of the file).
size_t readAt(int fd, uint64_t fileSize, uint64_t offset, unsigned char *dst, size_t want) {
  if (offset >= fileSize) return 0;
  const size_t n = static_cast<size_t>(std::min<uint64_t>(want, fileSize - offset));
  const size_t slice = 8u << 20;
  unsigned nThreads = static_cast<unsigned>(std::min<size_t>((n + slice - 1) / slice, 8));
  const unsigned hw = std::thread::hardware_concurrency();
  if (hw && nThreads > std::max(1u, hw / 2)) nThreads = std::max(1u, hw / 2);
  std::vector<size_t> got(std::max(1u, nThreads), 0);
  auto work = [&](unsigned t) {
    const size_t per = (n + nThreads - 1) / nThreads;
    const size_t a = std::min(n, per * t), b = std::min(n, a + per);
    size_t done = a;
    while (done < b) {
      const ssize_t r = ::pread(fd, dst + done, b - done, static_cast<off_t>(offset + done));
      if (r <= 0) break;
      done += static_cast<size_t>(r);
    }
    got[t] = done - a;
  };
  if (nThreads <= 1) { nThreads = 1; work(0); return got[0]; }
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nThreads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto &x : th) x.join();
  // (a short slice in the middle would leave a hole: report only the contiguous part)
  size_t total = 0;
  const size_t per = (n + nThreads - 1) / nThreads;
  for (unsigned t = 0; t < nThreads; ++t) {
    total += got[t];
    if (got[t] < std::min(n, per * (t + 1)) - std::min(n, per * t)) break;
  }
  return total;
}

}  // namespace

DeviceBamFeeder::DeviceBamFeeder(mma_ctx *ctx, const FeatureTable &features, Strandedness strandedness)
    : ctx_(ctx), features_(features), strandedness_(strandedness) {
  size_t mb = 32;  // (page-locking costs ~0.45 ms per MB on the hosts this runs on: two small buffers, refilled by several threads)
  if (const char *e = std::getenv("MMANNOT_B200_BAM_CHUNK_MB")) mb = static_cast<size_t>(std::max(1, std::atoi(e)));
  cap_ = mb << 20;
}

DeviceBamFeeder::~DeviceBamFeeder() {
  mma_free_pinned(buf_[0]);
  mma_free_pinned(buf_[1]);
}

DeviceBamFeeder::Result DeviceBamFeeder::run(const std::string &fileName, uint32_t column, uint64_t &nRecords, std::string &warnings,
                                             std::string &why, std::string &err) {
  nRecords = 0;
  const int fd = ::open(fileName.c_str(), O_RDONLY);
  if (fd < 0) { why = "cannot open the file"; return Result::FALLBACK; }
  struct Closer { int fd; ~Closer() { ::close(fd); } } closer{fd};
  struct stat stt;
  if (::fstat(fd, &stt) != 0 || !S_ISREG(stt.st_mode)) { why = "not a regular file"; return Result::FALLBACK; }
  const uint64_t fileSize = static_cast<uint64_t>(stt.st_size);
  uint64_t filePos = 0;
  const auto tRun0 = std::chrono::steady_clock::now();
  for (int k = 0; k < 2; ++k)
    if (!buf_[k]) {
      buf_[k] = static_cast<unsigned char *>(mma_alloc_pinned(cap_));
      if (!buf_[k]) { why = "no page-locked memory for the file chunks"; return Result::FALLBACK; }
    }
  const bool timing = std::getenv("MMANNOT_B200_TIMING") != nullptr;
  const double msPinned = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tRun0).count();
  double msRead = 0, msSubmit = 0, msStage = 0, msReserve = 0;
  auto now = []() { return std::chrono::steady_clock::now(); };
  auto since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  if (timing) mma_timing_enable(ctx_, 1);
  int cur = 0;
  auto tr0 = now();
  size_t have = readAt(fd, fileSize, filePos, buf_[cur], cap_);
  filePos += have;
  msRead += since(tr0);
  bool eof = have < cap_;
  // ---- the BAM header, inflated here: magic, text, reference names (mm:1487-1520)
  std::vector<unsigned char> head;
  std::vector<uint32_t> headMemberEnd;  // inflated bytes up to the end of each member looked at
  std::vector<std::string> refNames;
  size_t headerBytes = 0, scanned = 0;
  bool headerDone = false;
  while (!headerDone) {
    size_t hdr = 0;
    const size_t total = bgzfMember(buf_[cur] + scanned, have - scanned, &hdr);
    if (total == 0 || scanned + total > have || total < hdr + 8) { why = "not a BGZF file (or a header larger than a chunk)"; return Result::FALLBACK; }
    const unsigned char *m = buf_[cur] + scanned;
    const uint32_t isize = rd32(m + total - 4);
    const size_t at = head.size();
    head.resize(at + isize);
    if (isize && !inflateRaw(m + hdr, total - hdr - 8, head.data() + at, isize)) { why = "corrupt BGZF member in the header"; return Result::FALLBACK; }
    scanned += total;
    headMemberEnd.push_back(static_cast<uint32_t>(head.size()));
    // complete?
    if (head.size() < 12) continue;
    if (std::memcmp(head.data(), "BAM\1", 4) != 0) { why = "missing BAM magic"; return Result::FALLBACK; }
    const size_t lText = rd32(head.data() + 4);
    size_t p = 8 + lText;
    if (head.size() < p + 4) continue;
    const uint32_t nRef = rd32(head.data() + p);
    p += 4;
    refNames.clear();
    bool complete = true;
    for (uint32_t i = 0; i < nRef; ++i) {
      if (head.size() < p + 4) { complete = false; break; }
      const size_t lName = rd32(head.data() + p);
      if (head.size() < p + 4 + lName + 4) { complete = false; break; }
      std::string name(reinterpret_cast<const char *>(head.data() + p + 4), lName);
      refNames.push_back(name.c_str());  // up to the first NUL, like the reference (mm:1510)
      p += 4 + lName + 4;
    }
    if (!complete) continue;
    headerBytes = p;
    headerDone = true;
  }
  // the member that holds the first record, and how much of it is still header
  size_t firstMember = 0;
  while (firstMember < headMemberEnd.size() && headMemberEnd[firstMember] <= headerBytes) ++firstMember;
  size_t pos = 0;  // byte offset in buf_[cur] of member `firstMember`
  {
    size_t off = 0;
    for (size_t k = 0; k < firstMember; ++k) { size_t hdr; off += bgzfMember(buf_[cur] + off, have - off, &hdr); }
    pos = off;
  }
  uint32_t skipFirst = (firstMember < headMemberEnd.size()) ? static_cast<uint32_t>(headerBytes - (firstMember ? headMemberEnd[firstMember - 1] : 0)) : 0;
  // reference -> annotation chromosome (only chromosomes that carry features, like XamReader)
  std::unordered_map<std::string, uint32_t> chrByName;
  for (size_t i = 0; i < features_.chromosomes.size(); ++i)
    if (features_.chrHasFeatures[i]) chrByName[features_.chromosomes[i]] = static_cast<uint32_t>(i);
  std::vector<uint32_t> refToChr(refNames.size(), MMA_HIT_CHR_NONE);
  for (size_t i = 0; i < refNames.size(); ++i) {
    auto it = chrByName.find(refNames[i]);
    if (it != chrByName.end()) refToChr[i] = it->second;
  }
  const int strand = strandedness_ == Strandedness::U ? 0 : strandedness_ == Strandedness::F ? 1 : 2;
  if (mma_bam_begin(ctx_, column, refToChr.data(), static_cast<uint32_t>(refToChr.size()), strand) != MMA_OK) { err = mma_last_error(ctx_); return Result::FAILED; }
  // ---- the file goes to the device in pieces of one page-locked buffer (while the next piece is read); whole members pile up
  //      in the staging area until there are enough of them for one inflate launch (or the file ends)
  uint64_t kMaxComp = 1536ull << 20;
  const uint64_t kMaxOut = 3584ull << 20;
  // a launch as soon as this much is staged (checked between buffers): the device inflates while the file is still being read
  uint64_t kLaunch = 256ull << 20;
  if (kLaunch > kMaxComp) kLaunch = kMaxComp;
  if (const char *e = std::getenv("MMANNOT_B200_BAM_LAUNCH_MB")) kMaxComp = static_cast<uint64_t>(std::max(1, std::atoi(e))) << 20;  // (tests: several launches per file)
  {
    auto tv0 = now();
    if (fileSize > 0) mma_bam_reserve(ctx_, std::min<uint64_t>(fileSize, kMaxComp) + 65536);
    msReserve = since(tv0);
  }
  std::vector<uint32_t> memberOff, memberIsize;
  uint64_t staged = 0, inflated = 0;
  // A launch is started for what is staged and finished only before the next one starts (or at the end of the file): the device
  // inflates chunk k while chunk k + 1 is read and staged.
  bool pending = false;
  unsigned nLaunches = 0;
  auto finishPending = [&]() -> int {  // 0 ok, 1 fallback, 2 failed
    if (!pending) return 0;
    pending = false;
    uint64_t n = 0;
    uint32_t flags = 0;
    auto ts0 = now();
    if (mma_submit_bam_finish(ctx_, &n, &flags) != MMA_OK) { err = mma_last_error(ctx_); return 2; }
    msSubmit += since(ts0);
    if (flags) {
      why = std::string("the file needs the host decoder:") + ((flags & MMA_BAM_HAS_XA) ? " XA tags" : "") + ((flags & MMA_BAM_STRADDLE) ? " records across BGZF members" : "") +
            ((flags & MMA_BAM_ODD_CIGAR) ? " CIGAR operations with warnings" : "") + ((flags & MMA_BAM_ODD_AUX) ? " unknown aux types" : "") +
            ((flags & MMA_BAM_BAD_DEFLATE) ? " deflate data this decoder rejects" : "") + ((flags & MMA_BAM_MALFORMED) ? " malformed records" : "") +
            ((flags & MMA_BAM_KEY_COLLISION) ? " read names sharing a 64-bit key" : "");
      return 1;
    }
    nRecords += n;
    return 0;
  };
  auto submit = [&]() -> int {
    if (memberIsize.empty()) return 0;
    const int rcp = finishPending();
    if (rcp) return rcp;
    memberOff.push_back(static_cast<uint32_t>(staged));
    mma_bam_chunk c;
    c.data = nullptr; c.n_bytes = staged;
    c.member_offset = memberOff.data(); c.member_isize = memberIsize.data();
    c.n_members = static_cast<uint32_t>(memberIsize.size());
    c.skip_first = skipFirst;
    auto ts0 = now();
    if (mma_submit_bam_start(ctx_, column, &c) != MMA_OK) { err = mma_last_error(ctx_); return 2; }
    msSubmit += since(ts0);
    pending = true;
    ++nLaunches;
    skipFirst = 0;
    memberOff.clear(); memberIsize.clear();
    staged = 0; inflated = 0;
    return 0;
  };
  for (;;) {
    size_t at = pos, from = pos;  // [from, at) = scanned, not staged yet
    auto stageScanned = [&]() -> bool {
      if (at == from) return true;
      auto tg0 = now();
      if (mma_bam_stage(ctx_, buf_[cur] + from, at - from, staged - (at - from)) != MMA_OK) { err = mma_last_error(ctx_); return false; }
      msStage += since(tg0);
      from = at;
      return true;
    };
    while (at < have) {
      size_t hdr = 0;
      const size_t total = bgzfMember(buf_[cur] + at, have - at, &hdr);
      if (total == 0) {
        if (have - at >= 18 || eof) { why = "not BGZF all the way (or a truncated file)"; return Result::FALLBACK; }
        break;  // the member's header is cut by the end of the buffer
      }
      if (at + total > have) {
        if (eof) { why = "truncated BGZF member at the end of the file"; return Result::FALLBACK; }
        break;
      }
      const uint32_t isize = rd32(buf_[cur] + at + total - 4);
      if (staged + total > kMaxComp || inflated + isize > kMaxOut) {  // enough for one launch
        if (!stageScanned()) return Result::FAILED;
        const int rc = submit();
        if (rc == 1) return Result::FALLBACK;
        if (rc == 2) return Result::FAILED;
      }
      memberOff.push_back(static_cast<uint32_t>(staged));
      memberIsize.push_back(isize);
      staged += total;
      inflated += isize;
      at += total;
    }
    if (!stageScanned()) return Result::FAILED;
    if (at == pos && !(eof && at >= have)) { why = "a BGZF member larger than a buffer"; return Result::FALLBACK; }
    if (eof && at >= have) break;
    // the rest of this buffer (a member cut by its end) moves to the front of the other one, the file continues behind it (read
    // by a helper thread while this one may be waiting for a launch); the copy out of this buffer is waited for by the next
    // mma_bam_stage call, before this buffer is filled again
    const size_t rest = have - at;
    const int nxt = cur ^ 1;
    std::memcpy(buf_[nxt], buf_[cur] + at, rest);
    auto tr1 = now();
    const uint64_t readFrom = filePos;
    std::future<size_t> reading = std::async(std::launch::async, [&, nxt, rest, readFrom]() { return readAt(fd, fileSize, readFrom, buf_[nxt] + rest, cap_ - rest); });
    int rc = 0;
    if (staged >= (nLaunches == 0 ? std::min<uint64_t>(kLaunch, 64ull << 20) : kLaunch)) rc = submit();  // (a small first launch: the device starts early)
    const size_t got = reading.get();
    filePos += got;
    msRead += since(tr1);
    if (rc == 1) return Result::FALLBACK;
    if (rc == 2) return Result::FAILED;
    eof = got < cap_ - rest;
    have = rest + got;
    cur = nxt;
    pos = 0;
    if (have == 0) break;
  }
  {
    int rc = submit();
    if (rc == 0) rc = finishPending();
    if (rc == 1) return Result::FALLBACK;
    if (rc == 2) return Result::FAILED;
  }
  // ---- the reference's warnings for chromosomes the annotation does not know, in order of first appearance (mm:1297)
  if (!refNames.empty()) {
    std::vector<uint64_t> first(refNames.size());
    if (mma_bam_ref_first(ctx_, first.data(), static_cast<uint32_t>(first.size())) != MMA_OK) { err = mma_last_error(ctx_); return Result::FAILED; }
    std::vector<std::pair<uint64_t, size_t> > seen;
    for (size_t i = 0; i < refNames.size(); ++i)
      if (refToChr[i] == MMA_HIT_CHR_NONE && first[i] != ~0ull && refNames[i] != "*") seen.push_back(std::make_pair(first[i], i));
    std::sort(seen.begin(), seen.end());
    std::vector<std::string> said;
    for (const auto &s : seen) {
      const std::string &name = refNames[s.second];
      if (std::find(said.begin(), said.end(), name) != said.end()) continue;
      said.push_back(name);
      warnings += "\t\tWarning!  Chromosome '" + name + "' (found in your reads) is not present in your annotation file.\n";
    }
  }
  if (timing) {
    mma_timing t;
    if (mma_timing_get(ctx_, &t) == MMA_OK)
      std::fprintf(stderr, "[timing] bam file read %.1f ms, mma_submit_bam %.1f ms (host side), device: inflate %.1f ms, record walk + scan %.1f ms, parse %.1f ms, batch kernels %.1f ms; "
                           "host: page-locked buffers %.1f ms, reserve %.1f ms, staging calls %.1f ms, feeder total %.1f ms\n",
                   msRead, msSubmit, t.ms_bam_inflate, t.ms_bam_index, t.ms_bam_parse, t.ms_batch, msPinned, msReserve, msStage, since(tRun0));
  }
  return Result::DONE;
}

}  // namespace mmb
