// Test helper (no CUDA): exercises the host-side readers and writers so that tests/ can compare them with the Python
// mirror (fileIO.py) file by file.
//   hostio_check domain  in.dom out.dom     read a domain file, print a digest, write it back
//   hostio_check nml     file.nml           print the parsed namelists as "group.name = v1 | v2 ..."
//   hostio_check fmt                        print a few Fortran edit-descriptor renderings
#include <cstdio>
#include <cstring>

#include "domain_io.hpp"
#include "namelist.hpp"
#include "results_io.hpp"

using namespace i3rc_host;

int main(int argc, char** argv) {
  if (argc >= 4 && !strcmp(argv[1], "domain")) {
    Domain d;
    std::string err;
    if (!read_domain(argv[2], d, err)) {
      printf("ERROR %s\n", err.c_str());
      return 1;
    }
    printf("nx %d ny %d nz %d ncomp %zu\n", d.nx(), d.ny(), d.nz(), d.comps.size());
    for (auto& c : d.comps) {
      double se = 0, ss = 0;
      long sp = 0;
      for (float v : c.ext) se += v;
      for (float v : c.ssa) ss += v;
      for (int v : c.pfi) sp += v;
      double st = 0;
      for (float v : c.table.kind == 1 ? c.table.coefs : c.table.values) st += v;
      printf("component '%s' zbase %d nz %d uniform %d kind %d entries %d sums %.9g %.9g %ld %.9g\n", c.name.c_str(), c.zLevelBase, c.nz,
             (int)c.uniform, c.table.kind, c.table.entries(), se, ss, sp, st);
    }
    return write_domain(d, argv[3]) ? 0 : 1;
  }
  if (argc >= 3 && !strcmp(argv[1], "nml")) {
    Namelist n;
    if (!n.load(argv[2])) return 1;
    const char* keys[][2] = {{"radiativeTransfer", "solarMu"}, {"radiativeTransfer", "intensityMus"}, {"radiativeTransfer", "intensityPhis"},
                             {"monteCarlo", "numPhotonsPerBatch"}, {"algorithms", "useRayTracing"}, {"algorithms", "zetaMin"},
                             {"fileNames", "domainFileName"}, {"fileNames", "outputNetcdfFile"}, {"output", "reportAbsorptionProfile"}};
    for (auto& k : keys) {
      printf("%s.%s =", k[0], k[1]);
      for (auto& v : n.raw(k[0], k[1])) printf(" %s |", v.c_str());
      printf("\n");
    }
    printf("logical %d %d real %.6f\n", (int)n.logical("algorithms", "useRayTracing", false), (int)n.logical("output", "reportVolumeAbsorption", true),
           n.real("algorithms", "zetaMin", -1));
    return 0;
  }
  if (argc >= 2 && !strcmp(argv[1], "fmt")) {
    printf("[%s][%s][%s][%s][%s]\n", F(0.5, 7, 3).c_str(), F(-0.25, 9, 4).c_str(), F(12.3456, 5, 2).c_str(), F(0.85, 5, 2).c_str(), F(1234567.0, 7, 3).c_str());
    printf("[%s][%s][%s]\n", E13_6(1.0).c_str(), E13_6(1365.5).c_str(), E13_6(0.0123).c_str());
    return 0;
  }
  return 2;
}
