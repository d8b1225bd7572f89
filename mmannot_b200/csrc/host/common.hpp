// Host-side helpers shared by the front-end (config / annotation / alignment decode).
// Part of the drop-in host layer that feeds the device hot path; nothing here runs
// on the GPU and nothing here annotates reads.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace mmb {

static const size_t NO_ID = static_cast<size_t>(-1);

inline bool is_space(char c) {
  // same set as C isspace() in the "C" locale
  return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r';
}

inline std::string trimmed(const std::string &s) {
  size_t b = 0, e = s.size();
  while (b < e && is_space(s[b])) ++b;
  while (e > b && is_space(s[e - 1])) --e;
  return s.substr(b, e - b);
}
inline void ltrim_inplace(std::string &s) {
  size_t b = 0;
  while (b < s.size() && is_space(s[b])) ++b;
  if (b) s.erase(0, b);
}
inline void rtrim_inplace(std::string &s) {
  size_t e = s.size();
  while (e > 0 && is_space(s[e - 1])) --e;
  s.resize(e);
}

// getline()-style split: a trailing empty piece is dropped, an empty input gives no pieces.
inline void split_getline(const std::string &s, char delim, std::vector<std::string> &out) {
  out.clear();
  size_t pos = 0;
  while (pos < s.size()) {
    size_t q = s.find(delim, pos);
    if (q == std::string::npos) {
      out.push_back(s.substr(pos));
      return;
    }
    out.push_back(s.substr(pos, q - pos));
    pos = q + 1;
  }
}

// split at the first delimiter, both halves trimmed; false when the delimiter is absent
inline bool split_first(const std::string &s, char delim, std::string &a, std::string &b) {
  size_t p = s.find(delim);
  if (p == std::string::npos) return false;
  a = trimmed(s.substr(0, p));
  b = trimmed(s.substr(p + 1));
  return true;
}

inline std::string lowered(std::string s) {
  for (char &c : s) c = static_cast<char>(::tolower(static_cast<unsigned char>(c)));
  return s;
}

// strtoul with std::stoul's acceptance rules; ok=false where stoul would throw.
inline unsigned long parse_ulong(const std::string &s, bool &ok) {
  const char *p = s.c_str();
  char *endp = nullptr;
  errno = 0;
  unsigned long v = std::strtoul(p, &endp, 10);
  ok = (endp != p) && (errno != ERANGE);
  return v;
}

// 64-bit read-name key (the device groups the hits of a read by this key; the
// reference groups by the name string itself, mm:1656-1662, mm:1671).
inline uint64_t name_key(const char *p, size_t n) {
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (static_cast<uint64_t>(n) * 0xD6E8FEB86659FD93ull);
  while (n >= 8) {
    uint64_t w;
    std::memcpy(&w, p, 8);
    h = (h ^ w) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
    p += 8;
    n -= 8;
  }
  uint64_t w = 0;
  if (n) std::memcpy(&w, p, n);
  h = (h ^ w) * 0xC4CEB9FE1A85EC53ull;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return h;
}

}  // namespace mmb
