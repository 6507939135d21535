// Minimal classic netCDF (CDF-1 / CDF-2, "netCDF-3") reader and writer: what the reference's netCDF-Fortran calls
// produce and consume for domain files, phase-function tables and result files (Code/opticalProperties.f95:554-844,
// Code/scatteringPhaseFunctions.f95:899-1252, Example-Drivers/monteCarloDriver.f95:609-854).  Fixed-size variables
// only (the reference never defines a record dimension).  Format: "The NetCDF Classic Format Specification".
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace i3rc_host {

enum NcType { NC_BYTE = 1, NC_CHAR = 2, NC_SHORT = 3, NC_INT = 4, NC_FLOAT = 5, NC_DOUBLE = 6 };
inline size_t nc_size(int t) { return t == NC_BYTE || t == NC_CHAR ? 1 : t == NC_SHORT ? 2 : t == NC_DOUBLE ? 8 : 4; }

struct NcAtt {
  int type = NC_CHAR;
  std::vector<uint8_t> data;  // host byte order, nc_size(type) bytes per element
  size_t count() const { return data.size() / nc_size(type); }
  std::string text() const { return std::string(data.begin(), data.end()); }
  double number(size_t i = 0) const {
    const uint8_t* p = data.data() + i * nc_size(type);
    switch (type) {
      case NC_BYTE: return (double)*(const int8_t*)p;
      case NC_SHORT: { int16_t v; memcpy(&v, p, 2); return v; }
      case NC_INT: { int32_t v; memcpy(&v, p, 4); return v; }
      case NC_FLOAT: { float v; memcpy(&v, p, 4); return v; }
      case NC_DOUBLE: { double v; memcpy(&v, p, 8); return v; }
      default: return (double)*p;
    }
  }
};
struct NcVar {
  std::string name;
  std::vector<int> dimids;
  std::map<std::string, NcAtt> atts;
  int type = NC_FLOAT;
  uint64_t begin = 0, vsize = 0;
  std::vector<uint8_t> data;  // writer: host byte order
};

inline void swap_to_host(uint8_t* p, size_t n, size_t w) {  // big-endian file <-> little-endian host
  if (w == 1) return;
  for (size_t i = 0; i < n; i++)
    for (size_t a = 0, b = w - 1; a < b; a++, b--) std::swap(p[i * w + a], p[i * w + b]);
}

class NcFile {
 public:
  std::vector<std::pair<std::string, uint32_t>> dims;
  std::map<std::string, NcAtt> gatts;
  std::vector<std::string> gatt_order;  // writer: attributes in definition order
  std::vector<NcVar> vars;

  // ---------------------------------------------------------------- reading
  bool open(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    buf_.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
    pos_ = 0;
    if (buf_.size() < 8 || buf_[0] != 'C' || buf_[1] != 'D' || buf_[2] != 'F' || (buf_[3] != 1 && buf_[3] != 2)) return false;
    version_ = buf_[3];
    pos_ = 4;
    try {
      u32();  // numrecs
      uint32_t tag = u32(), n = u32();
      if (tag == 0x0A)
        for (uint32_t i = 0; i < n; i++) {
          std::string nm = name();
          dims.emplace_back(nm, u32());
        }
      read_atts(gatts);
      tag = u32();
      n = u32();
      if (tag == 0x0B)
        for (uint32_t i = 0; i < n; i++) {
          NcVar v;
          v.name = name();
          uint32_t nd = u32();
          for (uint32_t k = 0; k < nd; k++) v.dimids.push_back((int)u32());
          read_atts(v.atts);
          v.type = (int)u32();
          v.vsize = u32();
          v.begin = version_ == 1 ? u32() : u64();
          vars.push_back(v);
        }
    } catch (const std::exception&) {
      return false;
    }
    return true;
  }
  int dim_id(const std::string& n) const {
    for (size_t i = 0; i < dims.size(); i++)
      if (dims[i].first == n) return (int)i;
    return -1;
  }
  long dim_len(const std::string& n) const {
    int i = dim_id(n);
    return i < 0 ? -1 : (long)dims[i].second;
  }
  const NcVar* var(const std::string& n) const {
    for (auto& v : vars)
      if (v.name == n) return &v;
    return nullptr;
  }
  size_t var_count(const NcVar& v) const {
    size_t n = 1;
    for (int d : v.dimids) n *= dims[d].second;
    return n;
  }
  template <typename T>
  bool get(const std::string& n, std::vector<T>& out) const {  // converts the stored type to T
    const NcVar* v = var(n);
    if (!v) return false;
    size_t cnt = var_count(*v), w = nc_size(v->type);
    if (v->begin + cnt * w > buf_.size()) return false;
    out.resize(cnt);
    std::vector<uint8_t> tmp(buf_.begin() + v->begin, buf_.begin() + v->begin + cnt * w);
    swap_to_host(tmp.data(), cnt, w);
    for (size_t i = 0; i < cnt; i++) {
      const uint8_t* p = tmp.data() + i * w;
      switch (v->type) {
        case NC_BYTE: out[i] = (T) * (const int8_t*)p; break;
        case NC_CHAR: out[i] = (T)*p; break;
        case NC_SHORT: { int16_t x; memcpy(&x, p, 2); out[i] = (T)x; break; }
        case NC_INT: { int32_t x; memcpy(&x, p, 4); out[i] = (T)x; break; }
        case NC_FLOAT: { float x; memcpy(&x, p, 4); out[i] = (T)x; break; }
        default: { double x; memcpy(&x, p, 8); out[i] = (T)x; break; }
      }
    }
    return true;
  }

  // ---------------------------------------------------------------- writing
  int def_dim(const std::string& n, uint32_t len) {
    dims.emplace_back(n, len);
    return (int)dims.size() - 1;
  }
  void put_att_text(const std::string& n, const std::string& s) {
    NcAtt a;
    a.type = NC_CHAR;
    a.data.assign(s.begin(), s.end());
    set_gatt(n, a);
  }
  template <typename T>
  void put_att(const std::string& n, int type, T value) {
    NcAtt a;
    a.type = type;
    a.data.resize(nc_size(type));
    store(a.data.data(), type, (double)value);
    set_gatt(n, a);
  }
  template <typename T>
  void def_var(const std::string& n, int type, const std::vector<int>& dimids, const T* values) {
    NcVar v;
    v.name = n;
    v.type = type;
    v.dimids = dimids;
    size_t cnt = var_count(v), w = nc_size(type);
    v.data.resize(cnt * w);
    for (size_t i = 0; i < cnt; i++) store(v.data.data() + i * w, type, (double)values[i]);
    vars.push_back(std::move(v));
  }
  bool write(const std::string& path) {
    // header size first, then the begin offsets (CDF-1: 32-bit offsets, as nf90_create(nf90_Clobber) makes)
    auto padded = [](size_t n) { return (n + 3) & ~(size_t)3; };
    auto name_size = [&](const std::string& s) { return 4 + padded(s.size()); };
    auto atts_size = [&](const std::map<std::string, NcAtt>& m, const std::vector<std::string>& order) {
      size_t n = 8;
      for (auto& k : order) {
        const NcAtt& a = m.at(k);
        n += name_size(k) + 8 + padded(a.data.size());
      }
      return n;
    };
    size_t hdr = 4 + 4 + 8;
    for (auto& d : dims) hdr += name_size(d.first) + 4;
    hdr += atts_size(gatts, gatt_order);
    hdr += 8;
    for (auto& v : vars) hdr += name_size(v.name) + 4 + 4 * v.dimids.size() + 8 /* no variable attributes */ + 4 + 4 + 4;
    uint64_t off = hdr;
    for (auto& v : vars) {
      v.begin = off;
      v.vsize = padded(v.data.size());
      off += v.vsize;
    }
    if (off > 0x7fffffffULL) return false;  // would need CDF-2
    out_.clear();
    out_.reserve(off);
    const char magic[4] = {'C', 'D', 'F', 1};
    out_.insert(out_.end(), magic, magic + 4);
    w32(0);
    if (dims.empty()) { w32(0); w32(0); } else { w32(0x0A); w32((uint32_t)dims.size()); }
    for (auto& d : dims) { wname(d.first); w32(d.second); }
    watts(gatts, gatt_order);
    if (vars.empty()) { w32(0); w32(0); } else { w32(0x0B); w32((uint32_t)vars.size()); }
    for (auto& v : vars) {
      wname(v.name);
      w32((uint32_t)v.dimids.size());
      for (int d : v.dimids) w32((uint32_t)d);
      w32(0); w32(0);
      w32((uint32_t)v.type);
      w32((uint32_t)v.vsize);
      w32((uint32_t)v.begin);
    }
    for (auto& v : vars) {
      std::vector<uint8_t> tmp = v.data;
      swap_to_host(tmp.data(), tmp.size() / nc_size(v.type), nc_size(v.type));
      out_.insert(out_.end(), tmp.begin(), tmp.end());
      out_.resize(out_.size() + (v.vsize - tmp.size()), 0);
    }
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    f.write((const char*)out_.data(), (std::streamsize)out_.size());
    return (bool)f;
  }

 private:
  std::vector<uint8_t> buf_, out_;
  size_t pos_ = 0;
  int version_ = 1;
  void need(size_t n) const {
    if (pos_ + n > buf_.size()) throw std::runtime_error("truncated netCDF header");
  }
  uint32_t u32() {
    need(4);
    uint32_t v = ((uint32_t)buf_[pos_] << 24) | ((uint32_t)buf_[pos_ + 1] << 16) | ((uint32_t)buf_[pos_ + 2] << 8) | buf_[pos_ + 3];
    pos_ += 4;
    return v;
  }
  uint64_t u64() {
    uint64_t hi = u32();
    return (hi << 32) | u32();
  }
  std::string name() {
    uint32_t n = u32();
    need(n);
    std::string s((const char*)&buf_[pos_], n);
    pos_ += (n + 3) & ~3u;
    return s;
  }
  void read_atts(std::map<std::string, NcAtt>& m) {
    uint32_t tag = u32(), n = u32();
    if (tag != 0x0C) return;
    for (uint32_t i = 0; i < n; i++) {
      std::string nm = name();
      NcAtt a;
      a.type = (int)u32();
      uint32_t cnt = u32();
      size_t bytes = cnt * nc_size(a.type);
      need(bytes);
      a.data.assign(buf_.begin() + pos_, buf_.begin() + pos_ + bytes);
      swap_to_host(a.data.data(), cnt, nc_size(a.type));
      pos_ += (bytes + 3) & ~(size_t)3;
      m[nm] = a;
    }
  }
  void set_gatt(const std::string& n, const NcAtt& a) {
    if (!gatts.count(n)) gatt_order.push_back(n);
    gatts[n] = a;
  }
  static void store(uint8_t* p, int type, double v) {
    switch (type) {
      case NC_BYTE: { int8_t x = (int8_t)v; memcpy(p, &x, 1); break; }
      case NC_CHAR: { uint8_t x = (uint8_t)v; memcpy(p, &x, 1); break; }
      case NC_SHORT: { int16_t x = (int16_t)v; memcpy(p, &x, 2); break; }
      case NC_INT: { int32_t x = (int32_t)v; memcpy(p, &x, 4); break; }
      case NC_FLOAT: { float x = (float)v; memcpy(p, &x, 4); break; }
      default: memcpy(p, &v, 8);
    }
  }
  void w32(uint32_t v) {
    uint8_t b[4] = {(uint8_t)(v >> 24), (uint8_t)(v >> 16), (uint8_t)(v >> 8), (uint8_t)v};
    out_.insert(out_.end(), b, b + 4);
  }
  void wname(const std::string& s) {
    w32((uint32_t)s.size());
    out_.insert(out_.end(), s.begin(), s.end());
    out_.resize((out_.size() + 3) & ~(size_t)3, 0);
  }
  void watts(const std::map<std::string, NcAtt>& m, const std::vector<std::string>& order) {
    if (order.empty()) { w32(0); w32(0); return; }
    w32(0x0C);
    w32((uint32_t)order.size());
    for (auto& k : order) {
      const NcAtt& a = m.at(k);
      wname(k);
      w32((uint32_t)a.type);
      w32((uint32_t)a.count());
      std::vector<uint8_t> tmp = a.data;
      swap_to_host(tmp.data(), a.count(), nc_size(a.type));
      out_.insert(out_.end(), tmp.begin(), tmp.end());
      out_.resize((out_.size() + 3) & ~(size_t)3, 0);
    }
  }
};

}  // namespace i3rc_host
