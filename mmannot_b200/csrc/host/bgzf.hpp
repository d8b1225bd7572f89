// Multi-threaded BGZF inflate for the BAM decoder.
//
// The reference reads its BAM through gzread (mmannot.cpp:1487-1500), one deflate stream after the other on one core:
// ~27 % of its wall time.  A BGZF file is a series of independent gzip members of at most 64 KB, each announcing its own
// compressed size in the "BC" extra field, so the members of a chunk of the file can be inflated side by side.  This
// source hands the decoder the same byte stream gzread would, one chunk ahead of the parser.
// Files that are gzip but not BGZF are refused by open() (the caller falls back to gzread).
#pragma once
#include <cstdint>
#include <cstdio>
#include <future>
#include <string>
#include <vector>

namespace mmb {

class BgzfSource {
 public:
  ~BgzfSource();
  // false when the file does not start with a BGZF member (nothing is consumed that the caller needs)
  bool open(const std::string &path, unsigned threads);
  // Copies up to `cap` bytes of the inflated stream into dst; 0 = end of file, -1 = corrupt file (message in error()).
  long read(unsigned char *dst, size_t cap);
  const std::string &error() const { return error_; }

 private:
  struct Chunk {
    std::vector<unsigned char> data;
    size_t size = 0, pos = 0;
    bool last = false;
    std::string error;
  };
  void produce(Chunk &out);  // reads and inflates the next chunk of members

  FILE *file_ = nullptr;
  unsigned threads_ = 1;
  std::vector<unsigned char> comp_;  // compressed bytes read but not yet consumed
  size_t compPos_ = 0, compEnd_ = 0;
  bool fileEof_ = false;
  Chunk chunk_[2];
  int cur_ = 0;
  bool started_ = false, done_ = false;
  std::future<void> ahead_;
  std::string error_;
};

}  // namespace mmb
