// Domains and phase-function tables on the host side of the C ABI: the objects of Code/opticalProperties.f95 and
// Code/scatteringPhaseFunctions.f95 reduced to what new_Integrator needs, and their netCDF files
// (read_Domain / write_Domain, opticalProperties.f95:554-844; add_/read_PhaseFunctionTable,
// scatteringPhaseFunctions.f95:928-1252).  Arrays are in the reference's Fortran order (x fastest).
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "../include/i3rc_b200.h"
#include "netcdf3.hpp"

namespace i3rc_host {

struct PhaseTable {
  int kind = 0;  // 1 Legendre, 2 angle-value on one angle set
  std::vector<float> key, extinction, ssa;
  std::vector<int32_t> offsets;  // [n+1], Legendre
  std::vector<float> coefs;
  std::vector<float> angles, values;  // values[entry][angle]
  std::string description;
  int entries() const { return (int)key.size(); }
  i3rc_phase_table as_c() const {
    i3rc_phase_table t{};
    t.kind = kind;
    t.n_entries = entries();
    t.coef_offsets = offsets.data();
    t.coefs = coefs.data();
    t.n_angles = (int)angles.size();
    t.angles = angles.data();
    t.values = values.data();
    return t;
  }
};

struct Component {
  std::string name;
  int zLevelBase = 1, nz = 0;
  bool uniform = false;
  std::vector<float> ext, ssa;
  std::vector<int32_t> pfi;
  PhaseTable table;
};

struct Domain {
  std::vector<float> x, y, z;  // cell edges
  std::vector<Component> comps;
  int nx() const { return (int)x.size() - 1; }
  int ny() const { return (int)y.size() - 1; }
  int nz() const { return (int)z.size() - 1; }
};

inline std::string comp_prefix(int i) { return "Component" + std::to_string(i) + "_"; }  // makePrefix, :1006-1016

inline bool read_phase_table(const NcFile& f, const std::string& p, PhaseTable& t, std::string& err) {
  auto it = f.gatts.find(p + "phaseFunctionStorageType");
  if (it == f.gatts.end() || !f.get(p + "phaseFunctionKeyT", t.key)) {
    err = "read_PhaseFunctionTable: file doesn't look like a phase function table.";
    return false;
  }
  f.get(p + "extinctionT", t.extinction);
  f.get(p + "singleScatteringAlbedoT", t.ssa);
  auto d = f.gatts.find(p + "description");
  if (d != f.gatts.end()) t.description = d->second.text();
  std::string kind = it->second.text();
  if (kind.rfind("Angle-Value", 0) == 0) {
    t.kind = 2;
    if (!f.get(p + "scatteringAngle", t.angles) || !f.get(p + "phaseFunctionValues", t.values)) {
      err = "read_PhaseFunctionTable: Error reading phase function table.";
      return false;
    }
  } else if (kind.rfind("LegendreCoefficients", 0) == 0) {
    t.kind = 1;
    std::vector<int32_t> start, length;
    if (!f.get(p + "start", start) || !f.get(p + "length", length) || !f.get(p + "legendreCoefficients", t.coefs)) {
      err = "read_PhaseFunctionTable: Error reading phase function table.";
      return false;
    }
    t.offsets.assign(start.size() + 1, 0);
    for (size_t i = 0; i < start.size(); i++) {
      t.offsets[i] = start[i] - 1;  // 1-based in the file
      t.offsets[i + 1] = start[i] - 1 + length[i];
    }
  } else {
    err = "read_PhaseFunctionTable: Unknown phase function table format.";
    return false;
  }
  return true;
}

inline bool read_domain(const std::string& path, Domain& d, std::string& err) {
  NcFile f;
  if (!f.open(path)) {
    err = "read_Domain: Can't open file " + path;
    return false;
  }
  if (!f.get("x-Edges", d.x) || !f.get("y-Edges", d.y) || !f.get("z-Edges", d.z) || f.dim_id("z-Grid") < 0) {
    err = "read_Domain: " + path + " doesn't look an optical properties file.";
    return false;
  }
  auto nc = f.gatts.find("numberOfComponents");
  int ncomp = nc == f.gatts.end() ? 0 : (int)nc->second.number();
  const size_t ncol = (size_t)d.nx() * d.ny();
  for (int i = 1; i <= ncomp; i++) {
    const std::string p = comp_prefix(i);
    Component c;
    auto nm = f.gatts.find(p + "Name"), zb = f.gatts.find(p + "zLevelBase");
    const NcVar* v = f.var(p + "Extinction");
    if (nm == f.gatts.end() || zb == f.gatts.end() || !v) {
      err = "read_Domain: Error reading scalar fields from file " + path;
      return false;
    }
    c.name = nm->second.text();
    c.zLevelBase = (int)zb->second.number();
    c.uniform = v->dimids.size() == 1;
    c.nz = (int)f.dims[v->dimids[0]].second;  // file order is [z][y][x]
    if (!f.get(p + "Extinction", c.ext) || !f.get(p + "SingleScatteringAlbedo", c.ssa) || !f.get(p + "PhaseFunctionIndex", c.pfi) ||
        c.ext.size() != (c.uniform ? 1 : ncol) * (size_t)c.nz) {
      err = "read_Domain: Error reading scalar fields from file " + path;
      return false;
    }
    if (!read_phase_table(f, p, c.table, err)) {
      err = "read_Domain: Error reading phase function table.";
      return false;
    }
    d.comps.push_back(std::move(c));
  }
  return true;
}

inline void add_phase_table(NcFile& f, const std::string& p, const PhaseTable& t) {  // add_PhaseFunctionTable
  int ed = f.def_dim(p + "phaseFunctionNumber", (uint32_t)t.entries());
  std::vector<float> zeros(t.entries(), 0.0f);
  f.def_var(p + "phaseFunctionKeyT", NC_FLOAT, {ed}, t.key.data());
  f.def_var(p + "extinctionT", NC_FLOAT, {ed}, t.extinction.empty() ? zeros.data() : t.extinction.data());
  f.def_var(p + "singleScatteringAlbedoT", NC_FLOAT, {ed}, t.ssa.empty() ? zeros.data() : t.ssa.data());
  if (!t.description.empty()) f.put_att_text(p + "description", t.description);
  if (t.kind == 2) {
    int ad = f.def_dim(p + "scatteringAngle", (uint32_t)t.angles.size());
    f.def_var(p + "scatteringAngle", NC_FLOAT, {ad}, t.angles.data());
    f.def_var(p + "phaseFunctionValues", NC_FLOAT, {ed, ad}, t.values.data());
    f.put_att_text(p + "phaseFunctionStorageType", "Angle-Value");
  } else {
    std::vector<int32_t> start(t.entries()), length(t.entries());
    for (int i = 0; i < t.entries(); i++) {
      start[i] = t.offsets[i] + 1;
      length[i] = t.offsets[i + 1] - t.offsets[i];
    }
    int cd = f.def_dim(p + "coefficents", (uint32_t)t.coefs.size());  // (sic)
    f.def_var(p + "start", NC_INT, {ed}, start.data());
    f.def_var(p + "length", NC_INT, {ed}, length.data());
    f.def_var(p + "legendreCoefficients", NC_FLOAT, {cd}, t.coefs.data());
    f.put_att_text(p + "phaseFunctionStorageType", "LegendreCoefficients");
  }
}

inline bool regular(const std::vector<float>& e, float tol_ulps) {
  const float d = e[1] - e[0];
  for (size_t i = 1; i < e.size(); i++) {
    const float sp = std::nextafter(std::fabs(e[i]), INFINITY) - std::fabs(e[i]);
    if (!(std::fabs((e[i] - e[i - 1]) - d) <= tol_ulps * sp)) return false;
  }
  return true;
}

inline bool write_domain(const Domain& d, const std::string& path) {  // write_Domain, :554-706
  NcFile f;
  int xe = f.def_dim("x-Edges", (uint32_t)d.x.size()), ye = f.def_dim("y-Edges", (uint32_t)d.y.size()),
      ze = f.def_dim("z-Edges", (uint32_t)d.z.size());
  int xg = f.def_dim("x-Grid", (uint32_t)d.nx()), yg = f.def_dim("y-Grid", (uint32_t)d.ny()), zg = f.def_dim("z-Grid", (uint32_t)d.nz());
  f.def_var("x-Edges", NC_FLOAT, {xe}, d.x.data());
  f.def_var("y-Edges", NC_FLOAT, {ye}, d.y.data());
  f.def_var("z-Edges", NC_FLOAT, {ze}, d.z.data());
  f.put_att("xyRegularlySpaced", NC_BYTE, (int)(regular(d.x, 2.0f) && regular(d.y, 2.0f)));
  f.put_att("zRegularlySpaced", NC_BYTE, (int)regular(d.z, 2.0f));
  if (!d.comps.empty()) f.put_att("numberOfComponents", NC_INT, (int)d.comps.size());
  for (size_t i = 0; i < d.comps.size(); i++) {
    const Component& c = d.comps[i];
    const std::string p = comp_prefix((int)i + 1);
    f.put_att_text(p + "Name", c.name);
    f.put_att(p + "zLevelBase", NC_INT, c.zLevelBase);
    int zdim = zg;
    if (!(c.zLevelBase == 1 && c.nz == d.nz())) zdim = f.def_dim(p + "z-Grid", (uint32_t)c.nz);
    std::vector<int> dims = c.uniform ? std::vector<int>{zdim} : std::vector<int>{zdim, yg, xg};
    f.def_var(p + "Extinction", NC_FLOAT, dims, c.ext.data());
    f.def_var(p + "SingleScatteringAlbedo", NC_FLOAT, dims, c.ssa.data());
    f.def_var(p + "PhaseFunctionIndex", NC_SHORT, dims, c.pfi.data());
    add_phase_table(f, p, c.table);
  }
  return f.write(path);
}

}  // namespace i3rc_host
