#include "config.hpp"

#include <fstream>
#include <sstream>

namespace mmb {

namespace {

enum Section { SEC_NONE, SEC_SYNONYMS, SEC_INTRONS, SEC_VICINITY, SEC_ORDER };

// only the FIRST '*' becomes ".*" (mm:273, mm:314, mm:325)
std::string starToRegex(std::string key) {
  size_t p = key.find('*');
  if (p != std::string::npos) key.replace(p, 1, ".*");
  return key;
}

bool ruleMatches(const std::string &ruleSource, const std::string &ruleType, const std::string &source, const std::string &type) {
  return (ruleSource == "*" || ruleSource == source) && (ruleType == "*" || ruleType == type);
}

const char *strandSuffix(ElemStrand s) { return s == ES_F ? "(+)" : s == ES_R ? "(-)" : ""; }

}  // namespace

bool Config::parse(const std::string &fileName, std::string &err) {
  std::ifstream file(fileName.c_str());
  if (!file.good()) {
    err = "Error, configuration file '" + fileName + "' does not exists!";
    return false;
  }
  synonyms_.clear(); introns_.clear(); vicinity_.clear(); elements_.clear();
  nLines_ = 0;
  Section section = SEC_NONE;
  std::string raw, key, value;
  while (std::getline(file, raw)) {
    std::string line = trimmed(raw);
    if (line.empty() || line[0] == '#') continue;
    if (line == "Synonyms:") { section = SEC_SYNONYMS; continue; }
    if (line == "Introns:")  { section = SEC_INTRONS;  continue; }
    if (line == "Vicinity:") { section = SEC_VICINITY; continue; }
    if (line == "Order:")    { section = SEC_ORDER;    continue; }
    switch (section) {
      case SEC_SYNONYMS: {
        if (!split_first(line, ':', key, value)) {
          err = "Error, cannot parse line '" + line + "' in the 'Synonyms' section of the configuration file!";
          return false;
        }
        std::string re = starToRegex(key);
        try {
          synonyms_.push_back(Synonym{std::regex(re), value});
        } catch (const std::regex_error &) {
          err = "Error, cannot parse regular expression '" + re + "' in line '" + line + "' in the 'Synonyms' section of the configuration file!";
          return false;
        }
        break;
      }
      case SEC_INTRONS: {
        if (!split_first(line, ':', key, value)) {
          err = "Error, cannot parse line '" + line + "' in the 'Introns' section of the configuration file!";
          return false;
        }
        introns_.push_back(IntronRule{key, value, NO_ID});
        break;
      }
      case SEC_VICINITY: {
        if (!split_first(line, ':', key, value)) {
          err = "Error, cannot parse line '" + line + "' in the 'Vicinity' section of the configuration file!";
          return false;
        }
        vicinity_.push_back(VicinityRule{key, value, NO_ID, NO_ID});
        break;
      }
      case SEC_ORDER: {
        std::vector<std::string> fields;
        split_getline(line, ',', fields);
        for (std::string field : fields) {
          ElemStrand strand = ES_ALL;
          std::string head, tail;
          if (split_first(field, ' ', head, tail)) {
            if (tail == "+") strand = ES_F;
            else if (tail == "-") strand = ES_R;
            else {
              err = "Error, cannot parse line '" + line + "' in the 'Order' section of the configuration file (last item item should be the strand: '+' or '-')!";
              return false;
            }
            field = head;
          }
          std::string source = field, type;
          if (split_first(field, ':', key, value)) { source = key; type = value; }
          try {
            elements_.push_back(OrderElement{source, std::regex(starToRegex(source)), type, strand, static_cast<uint32_t>(nLines_)});
          } catch (const std::regex_error &) {
            err = "Error, cannot parse regular expression '" + source + "' in line '" + line + "' in the 'Order' section of the configuration file!";
            return false;
          }
        }
        ++nLines_;
        break;
      }
      default:
        err = "Error, line '" + line + "' is not in the 'Synonyms', 'Introns', 'Vicinity', nor 'Order' section !";
        return false;
    }
  }
  if (nLines_ == 0) {
    err = "Error, the 'Order' section is empty!  Please provide annotations.";
    return false;
  }
  for (IntronRule &r : introns_) {
    r.element = getOrder(r.source, "intron");
    if (r.element == NO_ID) {
      err = "Error, type '" + r.source + ":intron' (of '" + r.source + ":" + r.type + "') should be included in the 'Order:' section.";
      return false;
    }
  }
  for (VicinityRule &r : vicinity_) {
    r.up = getOrder(r.source, "upstream");
    if (r.up == NO_ID) {
      err = "Error, type '" + r.source + ":upstream' (of '" + r.source + ":" + r.type + "') should be included in the 'Order:' section.";
      return false;
    }
    r.down = getOrder(r.source, "downstream");
    if (r.down == NO_ID) {
      err = "Error, type '" + r.source + ":downstream' (of '" + r.source + ":" + r.type + "') should be included in the 'Order:' section.";
      return false;
    }
  }
  return true;
}

// Both lookups run std::regex_match per configuration entry; an annotation asks for the same few (source, type) strings
// hundreds of thousands of times, so the answers are memoised (the annotation is built by one thread).
std::string Config::translate(const std::string &s) const {
  auto hit = translateCache_.find(s);
  if (hit != translateCache_.end()) return hit->second;
  std::string out = s;
  for (const Synonym &syn : synonyms_)
    if (std::regex_match(s, syn.matcher)) { out = syn.value; break; }
  if (translateCache_.size() < 65536) translateCache_.emplace(s, out);
  return out;
}

size_t Config::getOrder(const std::string &source, const std::string &type) const {
  std::string key;
  key.reserve(source.size() + type.size() + 1);
  key.append(source).push_back('\t');
  key.append(type);
  auto hit = orderCache_.find(key);
  if (hit != orderCache_.end()) return hit->second;
  size_t out = NO_ID;
  for (size_t i = 0; i < elements_.size(); ++i) {
    const OrderElement &e = elements_[i];
    if (std::regex_match(source, e.matcher) && (e.type.empty() || e.type == type)) { out = i; break; }
  }
  if (orderCache_.size() < 65536) orderCache_.emplace(std::move(key), out);
  return out;
}

size_t Config::checkIntrons(const std::string &source, const std::string &type) const {
  for (const IntronRule &r : introns_)
    if (ruleMatches(r.source, r.type, source, type)) return r.element;
  return NO_ID;
}
size_t Config::checkUpstream(const std::string &source, const std::string &type) const {
  for (const VicinityRule &r : vicinity_)
    if (ruleMatches(r.source, r.type, source, type)) return r.up;
  return NO_ID;
}
size_t Config::checkDownstream(const std::string &source, const std::string &type) const {
  for (const VicinityRule &r : vicinity_)
    if (ruleMatches(r.source, r.type, source, type)) return r.down;
  return NO_ID;
}

std::string Config::getName(size_t i) const {
  if (i >= elements_.size()) return "";
  const OrderElement &e = elements_[i];
  std::string s = e.source;
  if (!e.type.empty()) s += ":" + e.type;
  if (e.strand == ES_F) s += " (+)";
  else if (e.strand == ES_R) s += " (-)";
  return s;
}

std::string Config::orderEcho() const {
  std::ostringstream os;
  os << "Order:\n";
  uint32_t line = 0;
  for (size_t i = 0; i < elements_.size(); ++i) {
    const OrderElement &e = elements_[i];
    if (e.line != line) { os << "\n"; line = e.line; }
    os << e.source << ":" << e.type << " " << strandSuffix(e.strand) << "\t";
  }
  os << "\n";
  return os.str();
}

void Config::deviceTables(std::vector<uint16_t> &line, std::vector<uint8_t> &strand, std::vector<uint8_t> &vicinity) const {
  line.clear(); strand.clear(); vicinity.clear();
  for (size_t i = 0; i < elements_.size(); ++i) {
    line.push_back(static_cast<uint16_t>(elements_[i].line));
    strand.push_back(elements_[i].strand);
    vicinity.push_back(isUpstream(i) ? EV_UP : isDownstream(i) ? EV_DOWN : EV_NONE);
  }
}

}  // namespace mmb
