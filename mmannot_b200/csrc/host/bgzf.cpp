#include "bgzf.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstring>
#include <thread>

namespace mmb {

namespace {

const size_t kTargetOut = 24u << 20;   // inflated bytes per chunk (about 400 members)
const size_t kReadStep = 8u << 20;     // compressed bytes per fread

struct Member {
  size_t compOff, compLen;  // deflate payload inside comp_
  size_t outOff, outLen;
  uint32_t crc;
};

inline uint32_t rd32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24); }
inline uint32_t rd16(const unsigned char *p) { return p[0] | (p[1] << 8); }

// total size of the BGZF member starting at p (0 = not a BGZF member, needs >= 18 bytes); *hdr = header length
size_t memberSize(const unsigned char *p, size_t avail, size_t *hdr) {
  if (avail < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
  const size_t xlen = rd16(p + 10);
  if (avail < 12 + xlen) return 0;
  size_t at = 12;
  while (at + 4 <= 12 + xlen) {
    const size_t slen = rd16(p + at + 2);
    if (p[at] == 'B' && p[at + 1] == 'C' && slen == 2 && at + 6 <= 12 + xlen) {
      *hdr = 12 + xlen;
      return static_cast<size_t>(rd16(p + at + 4)) + 1;
    }
    at += 4 + slen;
  }
  return 0;
}

}  // namespace

BgzfSource::~BgzfSource() {
  if (ahead_.valid()) ahead_.wait();
  if (file_) std::fclose(file_);
}

bool BgzfSource::open(const std::string &path, unsigned threads) {
  file_ = std::fopen(path.c_str(), "rb");
  if (!file_) return false;
  threads_ = std::max(1u, threads);
  comp_.resize(kReadStep * 2);
  compEnd_ = std::fread(comp_.data(), 1, kReadStep, file_);
  if (compEnd_ < kReadStep) fileEof_ = true;
  size_t hdr = 0;
  if (memberSize(comp_.data(), compEnd_, &hdr) == 0) {
    std::fclose(file_);
    file_ = nullptr;
    return false;
  }
  return true;
}

void BgzfSource::produce(Chunk &out) {
  out.size = 0; out.pos = 0; out.last = false; out.error.clear();
  std::vector<Member> members;
  size_t outBytes = 0;
  // collect whole members until the chunk is large enough
  for (;;) {
    if (compEnd_ - compPos_ < 18 + 65536 && !fileEof_) {  // top up the compressed buffer
      if (compPos_ > 0 && members.empty()) {
        std::memmove(comp_.data(), comp_.data() + compPos_, compEnd_ - compPos_);
        compEnd_ -= compPos_;
        compPos_ = 0;
      }
      if (comp_.size() - compEnd_ < kReadStep) comp_.resize(comp_.size() + kReadStep);
      const size_t got = std::fread(comp_.data() + compEnd_, 1, kReadStep, file_);
      compEnd_ += got;
      if (got < kReadStep) fileEof_ = true;
    }
    const size_t avail = compEnd_ - compPos_;
    if (avail == 0) { out.last = true; break; }
    size_t hdr = 0;
    const size_t total = memberSize(comp_.data() + compPos_, avail, &hdr);
    if (total == 0 || total > avail || total < hdr + 8) {
      if (total != 0 && total > avail && !fileEof_) continue;  // (cannot happen: a member is at most 64 KB)
      out.error = "truncated or corrupt BGZF block";
      out.last = true;
      break;
    }
    const unsigned char *m = comp_.data() + compPos_;
    Member mb;
    mb.compOff = compPos_ + hdr;
    mb.compLen = total - hdr - 8;
    mb.crc = rd32(m + total - 8);
    mb.outLen = rd32(m + total - 4);
    mb.outOff = outBytes;
    if (mb.outLen > 65536) { out.error = "corrupt BGZF block (size)"; out.last = true; break; }
    outBytes += mb.outLen;
    members.push_back(mb);
    compPos_ += total;
    if (outBytes >= kTargetOut) break;
  }
  if (out.data.size() < outBytes) out.data.resize(outBytes);
  out.size = outBytes;
  if (members.empty()) return;
  // inflate the members side by side
  const unsigned nT = static_cast<unsigned>(std::min<size_t>(threads_, members.size()));
  std::vector<std::string> errs(nT);
  auto work = [&](unsigned t) {
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) { errs[t] = "inflateInit2 failed"; return; }
    for (size_t k = t; k < members.size(); k += nT) {
      const Member &mb = members[k];
      inflateReset(&zs);
      zs.next_in = comp_.data() + mb.compOff;
      zs.avail_in = static_cast<uInt>(mb.compLen);
      zs.next_out = out.data.data() + mb.outOff;
      zs.avail_out = static_cast<uInt>(mb.outLen);
      const int rc = mb.outLen ? inflate(&zs, Z_FINISH) : Z_STREAM_END;
      if ((rc != Z_STREAM_END && !(rc == Z_OK && zs.avail_out == 0) && !(rc == Z_BUF_ERROR && zs.avail_out == 0)) || zs.avail_out != 0) {
        errs[t] = "corrupt BGZF block (inflate)";
        break;
      }
      if (crc32(crc32(0L, Z_NULL, 0), out.data.data() + mb.outOff, static_cast<uInt>(mb.outLen)) != mb.crc) {
        errs[t] = "corrupt BGZF block (CRC)";
        break;
      }
    }
    inflateEnd(&zs);
  };
  if (nT == 1) work(0);
  else {
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nT; ++t) pool.emplace_back(work, t);
    work(0);
    for (std::thread &th : pool) th.join();
  }
  for (const std::string &e : errs)
    if (!e.empty() && out.error.empty()) out.error = e;
}

long BgzfSource::read(unsigned char *dst, size_t cap) {
  if (!file_ || cap == 0) return 0;
  if (!started_) {
    produce(chunk_[0]);
    cur_ = 0;
    started_ = true;
    if (!chunk_[0].last) ahead_ = std::async(std::launch::async, [this]() { produce(chunk_[1]); });
  }
  size_t copied = 0;
  while (copied < cap) {
    Chunk &c = chunk_[cur_];
    if (c.pos < c.size) {
      const size_t n = std::min(cap - copied, c.size - c.pos);
      std::memcpy(dst + copied, c.data.data() + c.pos, n);
      c.pos += n;
      copied += n;
      continue;
    }
    if (!c.error.empty()) { error_ = c.error; return copied ? static_cast<long>(copied) : -1; }
    if (c.last || done_) { done_ = true; break; }
    ahead_.get();  // the next chunk is being (or has been) produced
    cur_ ^= 1;
    Chunk &n = chunk_[cur_];
    if (!n.last && n.error.empty()) {
      Chunk &other = chunk_[cur_ ^ 1];
      ahead_ = std::async(std::launch::async, [this, &other]() { produce(other); });
    }
  }
  return static_cast<long>(copied);
}

}  // namespace mmb
