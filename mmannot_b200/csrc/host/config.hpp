// Configuration file (Synonyms / Introns / Vicinity / Order) -> flattened element table.
// Behaviour follows the reference's Config class (mmannot.cpp:219-471); the data model
// here is a flat element array (what the device wants), not the reference's nested one.
#pragma once
#include <regex>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.hpp"

namespace mmb {

enum ElemStrand : uint8_t { ES_ALL = 0, ES_F = 1, ES_R = 2 };
enum ElemVicinity : uint8_t { EV_NONE = 0, EV_UP = 1, EV_DOWN = 2 };

struct OrderElement {
  std::string source;   // as written in the file (before '*' expansion)
  std::regex  matcher;  // source with its first '*' turned into ".*" (mm:313-316)
  std::string type;     // "" = any type (mm:231)
  ElemStrand  strand;   // ' +' / ' -' suffix (mm:303-311)
  uint32_t    line;     // Order line this element sits on (priority rank)
};

class Config {
 public:
  // Parses the file; on failure returns false and `err` holds the reference's message.
  bool parse(const std::string &fileName, std::string &err);

  // First matching synonym, else the input (mm:384-391).
  std::string translate(const std::string &s) const;

  // Flattened index of the first element matching (source, type), NO_ID if none (mm:414-424).
  size_t getOrder(const std::string &source, const std::string &type) const;
  size_t checkIntrons(const std::string &source, const std::string &type) const;     // mm:393-398
  size_t checkUpstream(const std::string &source, const std::string &type) const;    // mm:400-405
  size_t checkDownstream(const std::string &source, const std::string &type) const;  // mm:407-412

  size_t getNElements() const { return elements_.size(); }
  size_t getNLines() const { return nLines_; }
  const OrderElement &getElement(size_t i) const { return elements_[i]; }
  std::string getName(size_t i) const;  // mm:445-462
  bool isUpstream(size_t i) const { return elements_[i].type == "upstream"; }
  bool isDownstream(size_t i) const { return elements_[i].type == "downstream"; }

  // The "Order:" echo the reference prints on stderr after parsing (mm:375-381).
  std::string orderEcho() const;

  // Packed per-element tables for the device (mma_params.elem_*).
  void deviceTables(std::vector<uint16_t> &line, std::vector<uint8_t> &strand, std::vector<uint8_t> &vicinity) const;

 private:
  struct Synonym { std::regex matcher; std::string value; };
  struct IntronRule { std::string source, type; size_t element; };
  struct VicinityRule { std::string source, type; size_t up, down; };
  mutable std::unordered_map<std::string, std::string> translateCache_;  // memoised regex lookups (see config.cpp)
  mutable std::unordered_map<std::string, size_t> orderCache_;
  std::vector<Synonym>      synonyms_;
  std::vector<IntronRule>   introns_;
  std::vector<VicinityRule> vicinity_;
  std::vector<OrderElement> elements_;
  size_t nLines_ = 0;
};

}  // namespace mmb
