// Fortran namelist reader for the drivers' input files (Example-Drivers/monteCarloDriver.f95:90-103, 143-150;
// planeParallel.f95:55-112, 279-291): groups "&name ... /", "name = value[, value...]", logicals (T, .true., F,
// .false.), quoted strings, "!" comments, repeat counts "3*0.5".  Names are case-insensitive, as in Fortran.
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace i3rc_host {

class Namelist {
 public:
  bool load(const std::string& path) {
    std::ifstream in(path);
    if (!in) return false;
    std::string text, line;
    while (std::getline(in, line)) {
      bool q1 = false, q2 = false;
      for (size_t i = 0; i < line.size(); i++) {  // strip comments outside quotes
        if (line[i] == '\'' && !q2) q1 = !q1;
        if (line[i] == '"' && !q1) q2 = !q2;
        if (line[i] == '!' && !q1 && !q2) {
          line.erase(i);
          break;
        }
      }
      text += line + "\n";
    }
    size_t pos = 0;
    while ((pos = text.find('&', pos)) != std::string::npos) {
      size_t e = pos + 1;
      while (e < text.size() && (isalnum((unsigned char)text[e]) || text[e] == '_')) e++;
      std::string group = lower(text.substr(pos + 1, e - pos - 1));
      // the group ends at a "/" outside quotes
      size_t end = e;
      bool q1 = false, q2 = false;
      for (; end < text.size(); end++) {
        if (text[end] == '\'' && !q2) q1 = !q1;
        if (text[end] == '"' && !q1) q2 = !q2;
        if (text[end] == '/' && !q1 && !q2) break;
      }
      parse_group(group, text.substr(e, end - e));
      pos = end;
    }
    return true;
  }
  bool has(const std::string& g, const std::string& n) const {
    auto it = groups_.find(lower(g));
    return it != groups_.end() && it->second.count(lower(n));
  }
  std::vector<std::string> raw(const std::string& g, const std::string& n) const {
    auto it = groups_.find(lower(g));
    if (it == groups_.end()) return {};
    auto jt = it->second.find(lower(n));
    return jt == it->second.end() ? std::vector<std::string>{} : jt->second;
  }
  double real(const std::string& g, const std::string& n, double dflt) const {
    auto v = raw(g, n);
    return v.empty() ? dflt : to_real(v[0]);
  }
  long integer(const std::string& g, const std::string& n, long dflt) const {
    auto v = raw(g, n);
    return v.empty() ? dflt : (long)to_real(v[0]);
  }
  bool logical(const std::string& g, const std::string& n, bool dflt) const {
    auto v = raw(g, n);
    if (v.empty()) return dflt;
    std::string s = lower(v[0]);
    if (!s.empty() && s[0] == '.') s = s.substr(1);
    return !s.empty() && s[0] == 't';
  }
  std::string str(const std::string& g, const std::string& n, const std::string& dflt) const {
    auto v = raw(g, n);
    return v.empty() ? dflt : v[0];
  }
  std::vector<double> reals(const std::string& g, const std::string& n) const {
    std::vector<double> out;
    for (auto& s : raw(g, n)) out.push_back(to_real(s));
    return out;
  }

 private:
  std::map<std::string, std::map<std::string, std::vector<std::string>>> groups_;
  static std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)tolower(c); });
    return s;
  }
  static double to_real(std::string s) {
    for (auto& c : s)
      if (c == 'd' || c == 'D') c = 'e';
    return atof(s.c_str());
  }
  void parse_group(const std::string& group, const std::string& body) {
    // tokens: names followed by '=', values separated by commas / blanks
    std::vector<std::string> tok;
    std::string cur;
    char quote = 0;
    auto flush = [&]() {
      if (!cur.empty()) tok.push_back(cur);
      cur.clear();
    };
    for (char c : body) {
      if (quote) {
        if (c == quote) {
          tok.push_back("\x01" + cur);  // marks a string value (possibly empty)
          cur.clear();
          quote = 0;
        } else {
          cur += c;
        }
      } else if (c == '\'' || c == '"') {
        flush();
        quote = c;
      } else if (c == '=') {
        flush();
        tok.push_back("=");
      } else if (c == ',' || isspace((unsigned char)c)) {
        flush();
      } else {
        cur += c;
      }
    }
    flush();
    std::string name;
    for (size_t i = 0; i < tok.size(); i++) {
      if (i + 1 < tok.size() && tok[i + 1] == "=") {
        name = lower(tok[i]);
        groups_[group][name].clear();
        i++;
        continue;
      }
      if (name.empty() || tok[i] == "=") continue;
      std::string v = tok[i];
      if (!v.empty() && v[0] == '\x01') {
        groups_[group][name].push_back(v.substr(1));
        continue;
      }
      size_t star = v.find('*');
      if (star != std::string::npos && star > 0 && isdigit((unsigned char)v[0])) {  // r*c
        int r = atoi(v.substr(0, star).c_str());
        for (int k = 0; k < r; k++) groups_[group][name].push_back(v.substr(star + 1));
      } else {
        groups_[group][name].push_back(v);
      }
    }
  }
};

}  // namespace i3rc_host
